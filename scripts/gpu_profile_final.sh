#!/bin/bash
# gpurun helper: the judged profile set -- launch list of the bench command + one full capture of the
# PDIPM iteration kernel, the backward kernel and the AL-MPC solve kernel.
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --nb 4096 --no-e2e --no-cpu"
$SMALL > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fast_iter -s 12 -c 1 -o gpurun_out/prof_iter -f $SMALL > gpurun_out/ncu2.log 2>&1
echo "iter capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_al_solve -s 2 -c 1 -o gpurun_out/prof_mpc -f $SMALL > gpurun_out/ncu4.log 2>&1
echo "mpc capture rc=$?"
