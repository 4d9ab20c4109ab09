#!/bin/bash
# A/B of prebuilt library variants on the MPC shapes
cd /root/repo
O=gpurun_out/${TAG:-r2varm}; mkdir -p $O; rm -f $O/summary.txt
cp diff-qp-mpc_b200/b200qp/libb200qp.so /tmp/libb200qp_base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/libb200qp_base.so diff-qp-mpc_b200/b200qp/libb200qp.so; else cp build/variants/libb200qp_$v.so diff-qp-mpc_b200/b200qp/libb200qp.so; fi
  echo "== $v" >> $O/summary.txt
  timeout 200 python scripts/mpc_shapes_bench.py 2>&1 | tail -4 >> $O/summary.txt
done
cp /tmp/libb200qp_base.so diff-qp-mpc_b200/b200qp/libb200qp.so
cat $O/summary.txt
