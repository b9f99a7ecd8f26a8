"""Developer script: how much of the bench step is straggler tail and how much is per-iteration time under full occupancy."""
import os, sys, tempfile, heapq
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
opts = SolveOptions(tol=1e-2, constr_viol_tol=1e-2, max_iter=600)
plan = prepare_joint_batch(fn, agents, random_init_offsets(B, 4, seed=0), opts)
sv = plan.solver
d = sv.upload(plan.guess)


def timed():
    sv.set_inputs(d); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sv.run(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for a, r in zip(agents, plan.singles):
    bad = np.nonzero(r.status < 0)[0]
    print(a, "single failures", bad.tolist(), r.status[bad].tolist(), "iters p50 %d max %d" % (np.median(r.iters), r.iters.max()))
print("warm start timing", plan.timing)
timed()
ms = timed()
st, it, dbl = sv.fetch_stats(); it = it.cpu().numpy(); st = st.cpu().numpy()
np.save(os.path.join(ROOT, "gpurun_out", "iters_b%d.npy" % B), it)
sms = torch.cuda.get_device_properties(0).multi_processor_count
h = [0] * min(sms, B)
for n in it:  # the device queue hands out instances in index order to whichever CTA is free
    heapq.heappush(h, heapq.heappop(h) + int(n))
mk = max(h)
print("B=%d %.1f ms; iters sum %d mean %.1f med %d p90 %d p99 %d max %d; status ok %d" % (
    B, ms, it.sum(), it.mean(), np.median(it), np.percentile(it, 90), np.percentile(it, 99), it.max(), (st >= 0).sum()))
bad = np.unique(np.concatenate([np.nonzero(r.status < 0)[0] for r in plan.singles]))
print("joint iterations of the instances with a failed single:", it[bad].tolist(), " slowest 12:", np.argsort(-it)[:12].tolist(), np.sort(it)[::-1][:12].tolist())
print("queue makespan %d iterations vs ideal %.1f (efficiency %.2f); ms per iteration on the critical CTA %.2f" % (
    mk, it.sum() / len(h), it.sum() / len(h) / mk, ms / mk))
