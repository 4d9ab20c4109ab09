#!/bin/bash
# gpurun helper: parity tests + bench line (no profiler).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
r=d['roofline']
print('value %.0f solves/s  ms/step %.2f  e2e %s  iter_kernel %.3f ms frac_fp64 %.3f hbm_frac %.3f share %.2f  cfg1 %s cpu %s clocks %s' % (
 d['value'], d['ms_per_step'], d['e2e'] and round(d['e2e']['value']), r['avg_launch_ms'], r['frac'], r['hbm']['frac'], r['whole_solve']['kernel_share_of_step'], d['cfg1_nb128'], d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['clocks']))
PY
tail -3 gpurun_out/bench.err
