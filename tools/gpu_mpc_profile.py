"""Developer script: where a distributed-MPC control step spends its time (kernel vs host, iterations, phases)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["OBCA_PROFILE"] = "1"
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.vehicle_follower import MultiDistributedFollower
from conflict_rez_b200.pytypes import VehicleState
from conflict_rez_b200 import solver as S

AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
mdf = MultiDistributedFollower(fn, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, heads)
mdf.setup_multi_vehicles()

rec = []
orig_run = S.ObcaSolver.run


def run(self):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig_run(self); e1.record(); torch.cuda.synchronize()
    rec.append(e0.elapsed_time(e1))


S.ObcaSolver.run = run
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
mdf.solve(num_iter=n)
t = 1e3 * np.array(mdf.step_time[5:]); k = np.array(rec[5:])
its = np.array(getattr(mdf, "step_iters", []))
print("steps %d  step p50 %.2f ms p99 %.2f | k_solve p50 %.2f ms p99 %.2f | host share p50 %.2f ms" % (len(t), np.percentile(t, 50), np.percentile(t, 99), np.percentile(k, 50), np.percentile(k, 99), np.percentile(t - k[:len(t)], 50)))
if len(its):
    print("iterations per step (max over vehicles): p50 %d p99 %d max %d" % (np.percentile(its, 50), np.percentile(its, 99), its.max()))
prof = mdf.solver.debug_profile(); tot = sum(prof.values())
for kk, v in prof.items():
    if v:
        print("  %-18s %6.2f %%" % (kk, 100.0 * v / tot))
