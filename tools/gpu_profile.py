"""Developer script: per-phase cycle breakdown of k_solve on the 4-vehicle joint problem (OBCA_PROFILE=1)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OBCA_PROFILE", "1")
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
opts = SolveOptions(max_iter=600)
plan = prepare_joint_batch(fn, agents, random_init_offsets(B, 4), opts)
sv = plan.solver
d = sv.upload(plan.guess); sv.set_inputs(d); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sv.run(); e1.record(); torch.cuda.synchronize()
st, it, dbl = sv.fetch_stats(); it = it.cpu().numpy()
prof = sv.debug_profile(); tot = sum(prof.values())
print("B=%d  %.1f ms  iters sum %d med %d max %d" % (B, e0.elapsed_time(e1), it.sum(), np.median(it), it.max()))
for k, v in prof.items():
    print("  %-18s %6.2f %%   %9.1f kcycles/iteration" % (k, 100.0 * v / tot, v / it.sum() / 1e3))
print("  total %.1f kcycles/iteration" % (tot / it.sum() / 1e3))
