#!/bin/bash
# gpurun helper: A/B of the two fast-path variants (128-thread CTA vs one warp per QP).
mkdir -p gpurun_out
for NT in 128 32; do
  export B200QP_FAST_NT=$NT
  echo "== FAST_NT=$NT"
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_nt$NT.json 2> gpurun_out/bench_nt$NT.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_nt$NT.json'))
r=d['roofline']
print('value %.0f solves/s ms/step %.2f iter_kernel %.3f ms frac %.3f cfg1 %s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['cfg1_nb128']))
PY
done
