"""conflict_rez_b200 -- B200-native batched OBCA solver behind the conflict_rez planner API.

Only the strategy-guided OBCA hot path of XuShenLZ/conflict_rez is rebuilt here
(SURVEY.md section 8).  The data model mirrors ``confrez.pytypes`` /
``confrez.vehicle_types`` / ``confrez.obstacle_types``; the CasADi+IPOPT call
is replaced by hand-written sm_100a CUDA reached through a C ABI
(``include/obca.h``, loaded by :mod:`conflict_rez_b200.solver`).
"""

__version__ = "0.1.0"
