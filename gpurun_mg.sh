cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.err; tail -2 gpurun_out/bench_r1_n$N.err; tail -1 gpurun_out/bench_r1_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['config']['timing'], d['config']['volume'])"
done
