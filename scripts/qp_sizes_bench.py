"""The larger sizes of the BASELINE configs[4] sweep on their own (bench.py:bench_qp_sizes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
import bench
for k, v in bench.bench_qp_sizes(torch.device("cuda:0")).items():
    print(k, "ms/step %.2f" % v["ms_per_step"], "solves/s %.0f" % v["solves_per_s"], "n_iter", v["n_iter"], "TF/s %.2f" % v["tflops_algorithmic"])
