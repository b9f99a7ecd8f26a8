"""Developer script: geometric check of a larger randomized batch (every converged plan must satisfy the reference problem
statement: collocation, tubes, terminal conditions, obstacle and vehicle clearances by plain geometry)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cases import check_solution_properties
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.batch_planner import solve_joint_batch, random_init_offsets
from conflict_rez_b200.solver import SolveOptions

B, seed = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 5
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
plan = solve_joint_batch(fn, agents, random_init_offsets(B, 4, seed=seed), SolveOptions(max_iter=600))
res = plan.result
ok = res.status >= 0
worst = check_solution_properties(plan.problem, res.z[ok], res.dt[ok])
print("B=%d seed=%d converged %d/%d iterations med %d max %d" % (B, seed, ok.sum(), B, np.median(res.iters), res.iters.max()))
print("worst violations over all converged plans:", {k: float("%.3e" % v) for k, v in worst.items()}, " dmin =", plan.problem.dmin)
print("timing", plan.timing)
