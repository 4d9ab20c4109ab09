import os, sys, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "diff-qp-mpc_b200"))
import torch, bench
out = bench.bench_mpc_shapes(torch.device("cuda:0"))
for k, v in out.items():
    print(f"{k:32s} {v['ms_per_call']:8.3f} ms  {v['rollouts_per_s']:10.0f} rollouts/s")
