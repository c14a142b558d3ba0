cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_multi_${N}.log 2>&1
grep '^{' gpurun_out/bench_multi_${N}.log | cut -c1-140
