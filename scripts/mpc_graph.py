import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))); sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))), "diff-qp-mpc_b200"))
import torch, bench
print(json.dumps(bench.bench_mpc(torch.device("cuda:0")), indent=0))
