cd $GRAFT_REPO_ROOT
for T in 32768 65536 131072 262144 524288; do
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --tile $T > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; python -c "
import json;d=json.load(open('gpurun_out/bench_t.json'));print($T, d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['roofline']['other_kernels'])"
done
