cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python bench_ops.py --reps 3 --configs C4 > gpurun_out/ops_r1.jsonl 2> gpurun_out/ops_r1.err; tail -3 gpurun_out/ops_r1.err; cat gpurun_out/ops_r1.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'],d['tool'],'in %.2fGB out %.2fGB'%(d['input_GB'],d['output_GB']),'%.3f ms'%d['kernel_ms'],'%.0f GB/s'%d['algorithmic_GB_per_s'],'%.1f%%'%(100*d['frac_of_measured_hbm_peak']),d['rows'],d['flagged'])"
