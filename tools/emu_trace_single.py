"""Developer script: iteration trace of one single-vehicle solve of the bench workload (host emulation)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from conflict_rez_b200 import solver as _s
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import random_init_offsets
from conflict_rez_b200.control.scenario import build_problem, build_guess
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

agent, ids = sys.argv[1], [int(a) for a in sys.argv[2:]]
ia = int(agent.split("_")[1])
lib = _s.load_library(os.path.join(ROOT, "tools/host_emu/libobca_hostemu.so"))
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
opts = SolveOptions(tol=1e-2, constr_viol_tol=1e-2, max_iter=600)
offs = random_init_offsets(512, 4, seed=0)[ids]
p1 = build_problem(fn, [agent], init_offsets=offs[:, ia:ia + 1])
g1 = build_guess(p1, fn, [agent])
sv = ObcaSolver(p1, opts, device="cpu", lib=lib)
os.environ["OBCA_TRACE"] = "1"
r = sv.solve(g1)
print("status", r.status, "iters", r.iters, "obj", r.obj, "dt", r.dt)
