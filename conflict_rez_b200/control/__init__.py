"""Planner layer mirroring ``confrez/control`` (only the OBCA hot path and its callers)."""
