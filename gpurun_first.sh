set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -3 gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1a.json 2> gpurun_out/bench_ref_r1a.err; tail -3 gpurun_out/bench_ref_r1a.err; cat gpurun_out/bench_ref_r1a.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
