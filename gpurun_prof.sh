set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_dense_update|k_plic' -s 6 -c 4 -o gpurun_out/prof_r1a python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
