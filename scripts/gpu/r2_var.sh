#!/bin/bash
# A/B of prebuilt library variants (build/variants/libb200qp_<tag>.so) on the headline bench; "base" = the in-tree build
cd /root/repo
O=gpurun_out/${TAG:-r2var}; mkdir -p $O; rm -f $O/summary.txt
cp diff-qp-mpc_b200/b200qp/libb200qp.so /tmp/libb200qp_base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/libb200qp_base.so diff-qp-mpc_b200/b200qp/libb200qp.so; else cp build/variants/libb200qp_$v.so diff-qp-mpc_b200/b200qp/libb200qp.so; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --quick --no-e2e --no-cpu > $O/bench_$v.json 2> $O/bench_$v.err
  echo "bench $v rc=$?" >> $O/summary.txt
  python - <<PY >> $O/summary.txt
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  value %.0f solves/s  ms/step %.2f  frac %.4f  by_kernel %s  n_iter %s" % (d["value"], d["ms_per_step"], r["frac"], {k: round(v,3) for k,v in r["whole_solve"]["ms_per_step_by_kernel"].items()}, d["config"]["pdipm_iterations"]))
except Exception as e:
    print("  parse error", e)
PY
done
cp /tmp/libb200qp_base.so diff-qp-mpc_b200/b200qp/libb200qp.so
cat $O/summary.txt
