#!/bin/bash
# gpurun helper: the round's judged evidence in one call -- GPU tests, bench (both arms), launch list,
# full captures of the PDIPM iteration kernel (headline shape), the large-QP iteration kernel and the AL-MPC solve.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python scripts/qp_sweep.py > gpurun_out/qp_sweep.txt 2>&1; echo "sweep rc=$?"
SMALL="python bench.py --steps 2 --warmup 1 --nb 4096 --no-e2e --no-cpu"
$SMALL > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fast_iter -s 12 -c 1 -o gpurun_out/prof_iter -f $SMALL > gpurun_out/ncu2.log 2>&1
echo "iter capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_al_solve -s 2 -c 1 -o gpurun_out/prof_mpc -f $SMALL > gpurun_out/ncu4.log 2>&1
echo "mpc capture rc=$?"
python scripts/qp_one.py 100 888 4 > gpurun_out/one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pdipm_iter -s 2 -c 1 -o gpurun_out/prof_big -f python scripts/qp_one.py 100 888 4 > gpurun_out/ncu_big.log 2>&1
echo "big capture rc=$?"
