#!/bin/bash
# first GPU run of round 2: resident route tests + variant benches
cd /root/repo
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
timeout 900 python -m pytest tests/test_qp_resident_gpu.py -x -q -m gpu > gpurun_out/r2a/pytest_res.log 2>&1
echo "pytest_res rc=$?" >> gpurun_out/r2a/summary.txt
timeout 600 python -m pytest tests/test_qp_parity_gpu.py -x -q -m gpu > gpurun_out/r2a/pytest_qp.log 2>&1
echo "pytest_qp rc=$?" >> gpurun_out/r2a/summary.txt
for cfg in "RES=0" "RES=1 CH=4 P=1 S=1" "RES=1 CH=4 P=0 S=0" "RES=1 CH=4 P=1 S=0" "RES=1 CH=4 P=0 S=1" "RES=1 CH=2 P=1 S=1" "RES=1 CH=5 P=1 S=1" "RES=1 CH=10 P=1 S=1" "RES=1 CH=20 P=1 S=1"; do
  eval "$cfg"
  tag=$(echo "$cfg" | tr ' =' '__')
  B200QP_RES=$RES B200QP_RES_CH=${CH:-4} B200QP_RES_PANEL=${P:-1} B200QP_RES_SWEEP=${S:-1} timeout 300 python bench.py --steps 5 --warmup 3 --quick --no-e2e --no-cpu > gpurun_out/r2a/bench_$tag.json 2> gpurun_out/r2a/bench_$tag.err
  echo "bench $cfg rc=$?" >> gpurun_out/r2a/summary.txt
  python - <<PY >> gpurun_out/r2a/summary.txt
import json
try:
    d=json.loads(open("gpurun_out/r2a/bench_$tag.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  value %.0f solves/s  ms/step %.2f  frac %.4f  by_kernel %s  n_iter %s nan_onset %s rerun %s" % (d["value"], d["ms_per_step"], r["frac"], {k: round(v,3) for k,v in r["whole_solve"]["ms_per_step_by_kernel"].items()}, d["config"]["pdipm_iterations"], d["config"].get("nan_onset_iteration"), d["config"].get("exact_rerun")))
except Exception as e:
    print("  parse error", e)
PY
done
tail -5 gpurun_out/r2a/pytest_res.log >> gpurun_out/r2a/summary.txt
tail -5 gpurun_out/r2a/pytest_qp.log >> gpurun_out/r2a/summary.txt
cat gpurun_out/r2a/summary.txt
