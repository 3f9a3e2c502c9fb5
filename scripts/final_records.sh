#!/bin/bash
# Round records on one B200: GPU tests, the bench line, the reference arm, the ncu launch list of the bench command and
# one full ncu capture of the streaming kernel and the plane-positioning kernel.  Usage: bash scripts/final_records.sh <tag>
T=${1:-rX}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_gpu_pytest.txt 2>&1; tail -2 $O/${T}_gpu_pytest.txt
python bench.py --steps 20 --warmup 3 > $O/${T}_bench_256.json 2> $O/${T}_bench_256.err; tail -c 600 $O/${T}_bench_256.json
python bench.py --impl reference --steps 20 --warmup 3 > $O/${T}_bench_256_reference_arm.json 2>> $O/${T}_bench_256.err; tail -c 300 $O/${T}_bench_256_reference_arm.json
# (the ncu passes keep the default schedule: --no-sched-auto; the shipped run-time selection is in the bench line's schedule_selected)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_256.csv python bench.py --steps 2 --warmup 1 --late-steps 0 --no-cpu --e2e-steps 1 --no-sched-auto > $O/${T}_ncu_bench.log 2>&1
python scripts/launch_summary.py $O/${T}_launches_256.csv > $O/${T}_launch_summary.txt 2>&1; head -30 $O/${T}_launch_summary.txt
ncu --set full --import-source on --clock-control none -k regex:'^k_dense_update$|k_plic_warp' -c 2 -o $O/${T}_dense_plic_256 python bench.py --steps 2 --warmup 1 --late-steps 0 --no-cpu --e2e-steps 1 --no-sched-auto > $O/${T}_ncu_full.log 2>&1
ls -la $O/${T}_dense_plic_256.ncu-rep
