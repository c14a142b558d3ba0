set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_multi_${N}.log 2>&1; tail -1 gpurun_out/bench_multi_${N}.log | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --workload synthA > gpurun_out/bench_multi_synthA_${N}.log 2>&1; tail -1 gpurun_out/bench_multi_synthA_${N}.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 1 --warmup 1 > gpurun_out/bench_multi_ref_${N}.log 2>&1; tail -1 gpurun_out/bench_multi_ref_${N}.log | cut -c1-300
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_single_full.log 2>&1; tail -1 gpurun_out/bench_single_full.log | cut -c1-1500
