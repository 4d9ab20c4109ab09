#!/bin/bash
# full GPU suite + default bench (both arms)
cd /root/repo
O=gpurun_out/${TAG:-r2full}; mkdir -p $O; rm -f $O/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/summary.txt
tail -5 $O/pytest_gpu.log >> $O/summary.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
if [ -z "$NOREF" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
echo "bench_ref rc=$?" >> $O/summary.txt
fi
cat $O/summary.txt
