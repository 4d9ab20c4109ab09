#!/bin/bash
# ncu --set full of one k_res_chunk launch (second launch of the second call: iterations 4..7, every problem alive)
cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
for v in "0 0" "1 1"; do
  set -- $v
  export B200QP_RES_PANEL=$1 B200QP_RES_SWEEP=$2 B200QP_RES_CH=4
  python scripts/run_qp_once.py 4096 2 > $O/plain_$1$2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_res_chunk -s 7 -c 1 -o $O/prof_res_$1$2 python scripts/run_qp_once.py 4096 2 > $O/ncu_$1$2.log 2>&1
  echo "variant $v rc=$?" >> $O/summary.txt
done
cat $O/summary.txt; tail -3 $O/plain_*.log
