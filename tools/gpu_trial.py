"""Developer script: first contact with the GPU (smoke + small batched joint solve with timings)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

g.smoke()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
opts = SolveOptions(max_iter=600)
t0 = time.time()
plan = prepare_joint_batch(fn, agents, random_init_offsets(B, 4), opts)
print("prepare %.2fs" % (time.time() - t0))
for a, r in zip(agents, plan.singles):
    print(a, "status", np.unique(r.status, return_counts=True), "iters med %d max %d" % (np.median(r.iters), r.iters.max()))
sv = plan.solver
d = sv.upload(plan.guess)
for rep in range(2):
    sv.set_inputs(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sv.run(); e1.record(); torch.cuda.synchronize()
    st, it, dbl = sv.fetch_stats(); torch.cuda.synchronize()
    st, it = st.cpu().numpy(), it.cpu().numpy()
    ms = e0.elapsed_time(e1)
    print("joint B=%d: %.1f ms, status %s, iters med %d max %d, obj[0] %.6f -> %.1f solves/s" % (B, ms, np.unique(st, return_counts=True), np.median(it), it.max(), dbl[0][0].item(), (st >= 0).sum() / ms * 1e3))
