# multi-GPU evidence: bash scripts/r02/multi.sh N  (under gpurun --gpus N): distributed tests on N devices, then bench.py on 1 and N GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -$N
timeout 600 python -m pytest tests/test_gpu_distributed.py -q --timeout 500 -p no:cacheprovider > gpurun_out/r02_pytest_multi$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_multi$N.log
tail -4 gpurun_out/r02_pytest_multi$N.log
for G in 1 $N; do
  if [ $G -eq 1 ]; then CMD="python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline"; fi
  t0=$SECONDS; timeout 900 $CMD > gpurun_out/r02_bench_multi${N}_g$G.log 2> gpurun_out/r02_bench_multi${N}_g$G.err; echo "bench g$G exit $? in $((SECONDS-t0)) s"; tail -2 gpurun_out/r02_bench_multi${N}_g$G.err | cut -c1-300
  python - $N $G <<'PY'
import json, sys
n, g = sys.argv[1:3]
lines = [l for l in open("gpurun_out/r02_bench_multi%s_g%s.log" % (n, g)) if l.startswith("{")]
if lines:
    d = json.loads(lines[-1])
    def show(name, w):
        print("g%s %-9s value %.4e e2e %.4e ms/step %.3f" % (g, name, w["value"], (w.get("e2e") or {}).get("value", 0), w["ms_per_step"]))
    show("headline", d)
    for k, w in d.get("workloads", {}).items(): show(k, w)
PY
done
