"""Developer script: elastic magnitude / statuses of the first closed-loop steps (host loop)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.vehicle_follower import MultiDistributedFollower
from conflict_rez_b200.pytypes import VehicleState
AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
HEADINGS = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
np.random.seed(0)
m = MultiDistributedFollower(fn, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, HEADINGS, device=sys.argv[2] if len(sys.argv) > 2 else "cuda:0",
                             lib=None)
m.setup_multi_vehicles()
orig = m.solver.solve_step
def wrap(*a, **k):
    r = orig(*a, **k)
    print("status", r.status.tolist(), "iters", r.iters.tolist(), "elastic", np.round(r.elastic, 4).tolist(), "cviol", np.round(r.cviol, 4).tolist())
    return r
m.solver.solve_step = wrap
m.solve(num_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 6)
for v in m.vehicles:
    print(v.agent, "state", round(v.state.x.x, 3), round(v.state.x.y, 3), round(v.state.e.psi, 3))
