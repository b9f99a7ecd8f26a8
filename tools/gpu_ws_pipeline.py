"""Developer script: the two state warm starts of the batched pipeline side by side (python tools/gpu_ws_pipeline.py [batch])."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
from conflict_rez_b200.solver import SolveOptions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
offs = random_init_offsets(B, 4, seed=0)
for rep in range(2):
    for mode in ("collocation", "euler"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan = prepare_joint_batch(fn, agents, offs, SolveOptions(max_iter=600), device="cuda:0", final_headings=heads, state_ws=mode)
        torch.cuda.synchronize()
        t_ws = time.perf_counter() - t0
        sv = plan.solver
        sv.set_order(np.sum([r.iters for r in plan.singles], axis=0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sv.set_inputs(plan.dev_guess)
        e0.record(); sv.run(); e1.record()
        st, it, _ = sv.fetch_stats()
        torch.cuda.synchronize()
        st, it = st.cpu().numpy(), it.cpu().numpy()
        print("%-11s B=%d warm start %.2f s (%s) | singles: iters p50 %s max %s fail %s | joint: %.1f ms, %d/%d converged, iters p50 %d max %d sum %d" % (
            mode, B, t_ws, {k: round(v, 2) if isinstance(v, float) else v for k, v in plan.timing.items()},
            [int(np.median(r.iters)) for r in plan.singles], [int(r.iters.max()) for r in plan.singles], [int((r.status < 0).sum()) for r in plan.singles],
            e0.elapsed_time(e1), (st >= 0).sum(), B, np.median(it), it.max(), it.sum()), flush=True)
        sv.close()
