#!/bin/bash
# ncu --set full of one k_wres_chunk launch (second launch of the second call: iterations 4..7, every problem alive)
cd /root/repo
O=gpurun_out/${1:-r2w2}; mkdir -p $O; rm -f $O/summary.txt
export B200QP_RES_CH=4
python scripts/run_qp_once.py 8192 2 > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wres_chunk -s 7 -c 1 -f -o $O/prof_wres python scripts/run_qp_once.py 8192 2 > $O/ncu.log 2>&1
echo "rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/plain.log
