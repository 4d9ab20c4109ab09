#!/bin/bash
# round-2 evidence: full GPU suite, bench (both arms), ncu launch list, ncu --set full of the warp-per-QP kernels
cd /root/repo
O=gpurun_out/${TAG:-r2prof}; mkdir -p $O; rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/summary.txt
if [ -z "$NOTEST" ]; then
timeout 1500 python -m pytest tests -q -m gpu -x -o faulthandler_timeout=400 --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/summary.txt
tail -15 $O/pytest_gpu.log >> $O/summary.txt
fi
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
echo "bench_ref rc=$?" >> $O/summary.txt
timeout 600 python bench.py --impl reference --ref-nb 32768 --steps 1 --warmup 0 > $O/bench_ref_nb32768.json 2> $O/bench_ref_nb32768.err
echo "bench_ref nb=32768 rc=$?" >> $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> $O/summary.txt
python scripts/run_qp_once.py 32768 2 > $O/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:k_wres -s 5 -c 5 -f -o $O/prof_wres python scripts/run_qp_once.py 32768 2 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?" >> $O/summary.txt
# per-line listing of the dominant kernel only (the report with source is ~15 MB; gpurun_out/ is capped at 64 MiB)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wres_chunk -s 3 -c 1 -f -o $O/prof_wres_chunk_src python scripts/run_qp_once.py 32768 2 > $O/ncu_src.log 2>&1
echo "ncu chunk source rc=$?" >> $O/summary.txt
python scripts/run_rex_once.py > $O/plain_rex.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:k_al_solve -s 1 -c 1 -f -o $O/prof_rex python scripts/run_rex_once.py > $O/ncu_rex.log 2>&1
echo "ncu rex rc=$?" >> $O/summary.txt
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | head -25 >> $O/topo.txt 2>&1
# gpurun_out/ is capped at 64 MiB: keep the CSV pages of the reports, not the reports
for r in prof_wres prof_rex; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null && rm -f $O/$r.ncu-rep
done
if [ -f $O/prof_wres_chunk_src.ncu-rep ]; then
  ncu -i $O/prof_wres_chunk_src.ncu-rep --page raw --csv > $O/prof_wres_chunk_src_raw.csv 2>/dev/null
  ncu -i $O/prof_wres_chunk_src.ncu-rep --page source --csv > $O/wres_chunk_src.csv 2>/dev/null
  ncu -i $O/prof_wres_chunk_src.ncu-rep --page source --print-source cuda,sass --csv > $O/wres_chunk_src2.csv 2>/dev/null
  rm -f $O/prof_wres_chunk_src.ncu-rep
fi
du -sh $O >> $O/summary.txt
cat $O/summary.txt
