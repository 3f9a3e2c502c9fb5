cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; tail -3 gpurun_out/bench_r1m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1m.json')); print('ms/step', d['ms_per_step'], 'Gcu/s', d['value']/1e9, 'dense ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'step_frac', d['roofline']['step_frac']); print(d['roofline']['note'][:90])"
SVOF_OVERLAP=0 SVOF_PROFILE=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 2>&1 >/dev/null | grep -E "us "
