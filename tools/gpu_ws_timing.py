import os, sys, tempfile, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control import batch_planner as bp
from conflict_rez_b200.solver import SolveOptions, ObcaSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
offs = bp.random_init_offsets(B, 4, seed=0)
T = {}
def timed(name, fn_):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn_(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0
        return r
    return w
bp.build_problem = timed("build_problem", bp.build_problem)
bp.pose_guess = timed("pose_guess", bp.pose_guess)
bp.kinematic_paths = timed("kinematic_paths", bp.kinematic_paths)
bp.euler_state_ws = timed("euler_state_ws(total)", bp.euler_state_ws)
oi = ObcaSolver.__init__; ObcaSolver.__init__ = timed("ObcaSolver.__init__", oi)
ObcaSolver.run = timed("run(k_solve)", ObcaSolver.run)
ObcaSolver.dual_ws = timed("dual_ws", ObcaSolver.dual_ws)
ObcaSolver.joint_dual_ws = timed("joint_dual_ws", ObcaSolver.joint_dual_ws)
ObcaSolver.fetch_solution = timed("fetch_solution", ObcaSolver.fetch_solution)
ObcaSolver.fetch_stats = timed("fetch_stats", ObcaSolver.fetch_stats)
ObcaSolver._to_dev = timed("_to_dev", ObcaSolver._to_dev)
ObcaSolver.set_inputs = timed("set_inputs", ObcaSolver.set_inputs)
ObcaSolver.close = timed("close", ObcaSolver.close)
for rep in range(2):
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan = bp.prepare_joint_batch(fn, agents, offs, SolveOptions(max_iter=600), device="cuda:0", final_headings=heads)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    plan.solver.close()
    print("total %.2f s" % tot, {k: round(v, 2) for k, v in sorted(T.items(), key=lambda kv: -kv[1])})
