"""Developer script: distributed MPC, one process per GPU (torchrun), NCCL all-gather of the predictions per control step.
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_mpc_distributed.py [steps]"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from conflict_rez_b200.control.strategy import write_strategy
from conflict_rez_b200.control.vehicle_follower import DistributedFollowerNode
from conflict_rez_b200.pytypes import VehicleState

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
np.random.seed(0)
node = DistributedFollowerNode(fn, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, heads, device="cuda:%d" % local)
node.setup_multi_vehicles()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
node.solve(num_iter=steps)
t = 1e3 * np.array(node.step_time[5:]); e = 1e3 * np.array(node.exchange_time[5:])
stats = torch.tensor([np.percentile(t, 50), np.percentile(t, 99), np.percentile(e, 50), np.percentile(e, 99)], dtype=torch.float64, device="cuda")
dist.all_reduce(stats, op=dist.ReduceOp.MAX)
final = {a: [tr.x[-1], tr.y[-1]] for a, tr in node.final_results.items()}
allf = [None] * world
dist.all_gather_object(allf, final)
if rank == 0:
    s = stats.cpu().tolist()
    print(json.dumps({"world": world, "vehicles_per_rank": len(node.vehicles), "steps": len(t), "solve_p50_ms": s[0], "solve_p99_ms": s[1],
                      "allgather_p50_ms": s[2], "allgather_p99_ms": s[3], "final_xy": {k: v for d in allf for k, v in d.items()}}))
dist.destroy_process_group()
