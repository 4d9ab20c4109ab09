#!/bin/bash
# warp-per-QP resident kernel: resident tests + bench variants given as arguments "W CH" pairs
cd /root/repo
O=gpurun_out/${TAG:-r2w3}; mkdir -p $O; rm -f $O/summary.txt
if [ -z "$NOTEST" ]; then
timeout 900 python -m pytest tests/test_qp_resident_gpu.py -q -m gpu -s -x --deselect tests/test_qp_resident_gpu.py::test_bench_batch_against_oracle > $O/pytest_res.log 2>&1
echo "pytest_res rc=$?" >> $O/summary.txt
grep -E "passed|failed|^FAILED|^ERROR" $O/pytest_res.log | tail -8 >> $O/summary.txt
fi
for cfg in "$@"; do
  set -- $cfg
  W=$1; CH=$2
  tag=W${W}_CH${CH}
  B200QP_RES=1 B200QP_RES_WARP=$W B200QP_RES_CH=$CH timeout 300 python bench.py --steps 5 --warmup 3 --quick --no-e2e --no-cpu > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "bench W=$W CH=$CH rc=$?" >> $O/summary.txt
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
cat $O/summary.txt
