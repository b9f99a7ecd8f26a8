"""Developer script: per-phase cycle breakdown of k_solve on a single-vehicle problem (OBCA_PROFILE=1)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OBCA_PROFILE", "1")
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.scenario import build_problem, build_guess
from conflict_rez_b200.control.batch_planner import random_init_offsets
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
agent = sys.argv[2] if len(sys.argv) > 2 else "vehicle_0"
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
offs = random_init_offsets(B, 4, seed=0)[:, [int(agent[-1])]]
prob = build_problem(fn, [agent], init_offsets=offs, final_headings=heads)
guess = build_guess(prob, fn, [agent])
sv = ObcaSolver(prob, SolveOptions(max_iter=600))
d = sv.upload(guess); sv.set_inputs(d); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sv.run(); e1.record(); torch.cuda.synchronize()
st, it, dbl = sv.fetch_stats(); it = it.cpu().numpy()
prof = sv.debug_profile(); tot = sum(prof.values())
print("%s B=%d  %.1f ms  iters sum %d med %d max %d  status ok %d" % (agent, B, e0.elapsed_time(e1), it.sum(), np.median(it), it.max(), int((st.cpu().numpy() >= 0).sum())))
for k, v in prof.items():
    if v: print("  %-18s %6.2f %%   %9.1f kcycles/iteration" % (k, 100.0 * v / tot, v / it.sum() / 1e3))
print("  total %.1f kcycles/iteration" % (tot / it.sum() / 1e3))
