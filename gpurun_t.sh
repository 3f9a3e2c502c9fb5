cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for st in 1 0; do
SVOF_DENSE_STAGED=$st timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_r1l_$st.json 2> gpurun_out/bench_r1l.err; tail -3 gpurun_out/bench_r1l.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1l_$st.json')); print('staged=$st ms/step', d['ms_per_step'], 'Gcu/s', d['value']/1e9, 'dense ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'step_frac', d['roofline']['step_frac']); print(d['roofline']['note'][:90])"
done
