import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "diff-qp-mpc_b200"))
import torch, bench
dev = torch.device("cuda:0")
if len(sys.argv) > 1:
    print("bench_mpc first:", bench.bench_mpc(dev)["cuda_graph_rollouts_per_s"])
for rep in range(3):
    out = bench.bench_mpc_shapes(dev)
    print(rep, {k: round(v["ms_per_call"], 2) for k, v in out.items()}, flush=True)
