#!/bin/bash
# first GPU check of the warp-per-QP resident kernel: resident tests, parity tests, A/B bench
cd /root/repo
O=gpurun_out/r2w1; mkdir -p $O; rm -f $O/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/smi.txt
timeout 900 python -m pytest tests/test_qp_resident_gpu.py -q -m gpu -s -x > $O/pytest_res.log 2>&1
echo "pytest_res rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_qp_parity_gpu.py -q -m gpu -s > $O/pytest_qp.log 2>&1
echo "pytest_qp rc=$?" >> $O/summary.txt
for cfg in "RES=1 W=1 CH=4" "RES=1 W=0 CH=4" "RES=0 W=0 CH=4" "RES=1 W=1 CH=10" "RES=1 W=1 CH=2"; do
  eval "$cfg"
  tag=$(echo "$cfg" | tr ' =' '__')
  B200QP_RES=$RES B200QP_RES_WARP=$W B200QP_RES_CH=${CH:-4} timeout 300 python bench.py --steps 5 --warmup 3 --quick --no-e2e --no-cpu > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "bench $cfg rc=$?" >> $O/summary.txt
  python - <<PY >> $O/summary.txt
import json
try:
    d=json.loads(open("$O/bench_$tag.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  value %.0f solves/s  ms/step %.2f  frac %.4f  by_kernel %s  n_iter %s nan_onset %s rerun %s" % (d["value"], d["ms_per_step"], r["frac"], {k: round(v,3) for k,v in r["whole_solve"]["ms_per_step_by_kernel"].items()}, d["config"]["pdipm_iterations"], d["config"].get("nan_onset_iteration"), d["config"].get("exact_rerun")))
except Exception as e:
    print("  parse error", e)
PY
done
grep -E "passed|failed|^FAILED|^ERROR" $O/pytest_res.log | tail -8 >> $O/summary.txt
grep -E "passed|failed|^FAILED|^ERROR" $O/pytest_qp.log | tail -8 >> $O/summary.txt
cat $O/summary.txt
