cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2
nproc
timeout 1300 python bench.py --size 512 --steps 10 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/bench_r1p_512.json 2> gpurun_out/bench_r1p_512.err; echo rc=$?
tail -c 2500 gpurun_out/bench_r1p_512.json; tail -5 gpurun_out/bench_r1p_512.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
