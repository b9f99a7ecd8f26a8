"""Developer script: the scaling-sweep leg of bench.py on its own (python tools/gpu_sweep.py [batch])."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from conflict_rez_b200.control.strategy import write_strategy

fn = os.path.join(tempfile.mkdtemp(), "4v"); write_strategy(fn)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = bench.sweep_leg(fn, torch.device("cuda", 0), 1e-2, batch)
for c in out:
    print(json.dumps(c))
