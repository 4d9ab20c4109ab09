#!/bin/bash
# gpurun helper: ncu launch list + one full capture of the iteration kernel on a small batch.
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --nb 4096 --no-e2e --no-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_pdipm_iter} -s ${NCU_SKIP:-30} -c ${NCU_COUNT:-2} -o gpurun_out/prof_iter -f $SMALL > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
