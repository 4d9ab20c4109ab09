#!/bin/bash
# gpurun helper: smoke, bench, ncu launch list + one full capture of the iteration kernel.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SMALL="python bench.py --steps 2 --warmup 1 --nb 4096 --no-e2e --no-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pdipm_iter -s 30 -c 3 -o gpurun_out/prof_iter $SMALL > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
