#!/bin/bash
# round-2 evidence: full GPU suite, bench (both arms), ncu launch list, ncu --set full of the warp-per-QP kernels
cd /root/repo
O=gpurun_out/${TAG:-r2prof}; mkdir -p $O; rm -f $O/summary.txt
if [ -z "$NOTEST" ]; then
timeout 1500 python -m pytest tests -q -m gpu -x -o faulthandler_timeout=400 --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/summary.txt
tail -15 $O/pytest_gpu.log >> $O/summary.txt
fi
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
echo "bench_ref rc=$?" >> $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> $O/summary.txt
python scripts/run_qp_once.py 32768 2 > $O/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wres -s 5 -c 5 -f -o $O/prof_wres python scripts/run_qp_once.py 32768 2 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?" >> $O/summary.txt
cat $O/summary.txt
