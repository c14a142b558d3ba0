cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; local t0=$SECONDS; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name ($((SECONDS-t0)) s): $(grep '^{' gpurun_out/$name.log | cut -c1-110)"; }
run bench_multi_${N} $T --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3
run bench_multi_synthM_${N} $T --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --workload synthM --no-e2e
run bench_multi_synthA_${N} $T --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --workload synthA --no-e2e
run bench_multi_ref_${N} $T --master-port 29514 bench.py --gpus $N --impl reference --steps 1 --warmup 1
