cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
SVOF_OVERLAP=0 SVOF_PROFILE=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 2>&1 >/dev/null | grep -E "k_bound|last step"
