cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -3 gpurun_out/bench_r1g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1g.json')); print('ms/step', d['ms_per_step'], 'Gcu/s', d['value']/1e9, 'dense ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'step_frac', d['roofline']['step_frac'], d['gpu_launches'])"
SVOF_OVERLAP=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_r1g_ov.json 2> gpurun_out/bench_r1g_ov.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1g_ov.json')); print('OVERLAP ms/step', d['ms_per_step'], 'Gcu/s', d['value']/1e9, 'dense ms', d['roofline']['kernel_ms'])"
