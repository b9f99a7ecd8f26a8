"""Developer script: per-step statuses / iterations of the 4-vehicle closed loop (device-resident loop)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.vehicle_follower import DeviceMpcLoop, MultiDistributedFollower
from conflict_rez_b200.pytypes import VehicleState
from conflict_rez_b200.solver import RETURN_STATUS

AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
HEADINGS = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
np.random.seed(0)
m = MultiDistributedFollower(fn, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, HEADINGS, device="cuda:0")
m.setup_multi_vehicles()
loop = DeviceMpcLoop(m)
loop.run(int(sys.argv[1]) if len(sys.argv) > 1 else 250)
ex = loop.export()
st, it = ex["status"], ex["iters"]
for k in np.unique(st): print(RETURN_STATUS[int(k)], int((st == k).sum()))
bad = np.argwhere(st < 0)
print("failed (step, vehicle, status, iters):", [(int(s), int(v), int(st[s, v]), int(it[s, v])) for s, v in bad][:60])
print("iters per step (max over vehicles):", it.max(1).tolist())
