"""MPC shape leg of bench.py on its own (fused AL-MPC solve + backward at the BASELINE config shapes)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
import bench
out = bench.bench_mpc_shapes(torch.device("cuda:0"))
for k, v in out.items():
    print(k, "ms_per_call %.3f" % v["ms_per_call"], "rollouts/s %.0f" % v["rollouts_per_s"], "forward_ms %.3f" % v["forward_ms"], "frac %.4f" % v["roofline"]["frac"])
