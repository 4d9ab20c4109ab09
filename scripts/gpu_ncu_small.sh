#!/bin/bash
# ncu capture of the iteration kernel at a latency-bound grid (nb=128): stall reasons = critical path.
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --nb 128 --no-e2e --no-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_fast_iter} -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-2} -o gpurun_out/prof_small -f $SMALL > gpurun_out/ncu3.log 2>&1
echo "ncu full rc=$?"
