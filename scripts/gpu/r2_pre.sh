#!/bin/bash
# warp-per-QP prefactor + 3-slot host pipeline: QP tests, then bench variants
cd /root/repo
O=gpurun_out/${TAG:-r2pre}; mkdir -p $O; rm -f $O/summary.txt
timeout 1200 python -m pytest tests/test_qp_resident_gpu.py tests/test_qp_parity_gpu.py -q -m gpu -x > $O/pytest_qp.log 2>&1
echo "pytest_qp rc=$?" >> $O/summary.txt
tail -4 $O/pytest_qp.log >> $O/summary.txt
for PRE in 1 0; do
  B200QP_RES_PRE=$PRE timeout 300 python bench.py --steps 6 --warmup 3 --quick --no-cpu > $O/bench_pre$PRE.json 2> $O/bench_pre$PRE.err
  echo "bench PRE=$PRE rc=$?" >> $O/summary.txt
  python - <<PY >> $O/summary.txt
import json
try:
    d=json.loads(open("$O/bench_pre$PRE.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  value %.0f solves/s  e2e %.0f (%.2f ms; single %.2f ms) ms/step %.2f  frac %.4f  by_kernel %s  n_iter %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["single_call_ms_per_step"], d["ms_per_step"], r["frac"], {k: round(v,3) for k,v in r["whole_solve"]["ms_per_step_by_kernel"].items()}, d["config"]["pdipm_iterations"]))
except Exception as e:
    print("  parse error", e)
PY
done
cat $O/summary.txt
