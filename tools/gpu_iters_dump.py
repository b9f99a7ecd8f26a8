"""Developer script: per-instance iteration counts of the joint solve and of its single-vehicle warm-start solves (predictor study)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
from conflict_rez_b200.solver import SolveOptions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
offs = random_init_offsets(B, 4, seed=0)
plan = prepare_joint_batch(fn, agents, offs, SolveOptions(max_iter=600), device="cuda:0", final_headings=heads)
sv = plan.solver
sv.set_inputs(plan.dev_guess); sv.run()
st, it, dbl = sv.fetch_stats()
torch.cuda.synchronize()
np.savez(os.path.join(ROOT, "gpurun_out", "iters_dump.npz"), joint=it.cpu().numpy(), status=st.cpu().numpy(), singles=np.stack([r.iters for r in plan.singles]),
         single_obj=np.stack([r.obj for r in plan.singles]), offs=offs, obj=dbl[0].cpu().numpy())
print("dumped", B)
