cd $GRAFT_REPO_ROOT
( time timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -2 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vcfx_scan_kernel -s 8 -c 2 -o gpurun_out/prof_r1_c2 $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
