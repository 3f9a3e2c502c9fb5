cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 10 > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; tail -3 gpurun_out/bench_r1p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1p.json')); print('ms/step', d['ms_per_step'], 'Gcu/s', d['value']/1e9); e=d['e2e']; print(e['ms_per_step'], e['value']/1e9, e['h2d_bytes_per_step'], e['d2h_bytes_per_step'])"
