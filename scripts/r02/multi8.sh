# 8-GPU evidence of round 2 (gpurun --gpus 8 -- bash scripts/r02/multi8.sh): the bench line at N = 8 (weak scaling of the Monte Carlo
# workloads, strong scaling of the 2^32 / 2^36 enumerations, e2e through pcetk_b200.distributed) and ONE run of BASELINE configs[4] at
# its full shape: 1000 sites / 3000 instances, 65536 chains x 41 pH values (8192 chains per pH value and GPU), nequil 500, nprod 100.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
t0=$SECONDS; timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_g$N.log 2> gpurun_out/r02_bench_g$N.err; echo "bench g$N exit $? in $((SECONDS-t0)) s"
t0=$SECONDS; timeout 600 $TR bench.py --gpus $N --workload synthM --chains $((65536 / N)) --nequil 500 --nprod 100 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r02_synthM_full_shape_g$N.log 2> gpurun_out/r02_synthM_full_shape_g$N.err; echo "synthM full shape exit $? in $((SECONDS-t0)) s"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for name in ("r02_bench_g%s.log" % n, "r02_synthM_full_shape_g%s.log" % n):
    lines = [l for l in open("gpurun_out/" + name) if l.startswith("{")]
    if not lines:
        print(name, "no JSON line"); continue
    d = json.loads(lines[-1])
    print(name, "headline value %.4e e2e %.4e ms/step %.3f" % (d["value"], (d.get("e2e") or {}).get("value", 0), d["ms_per_step"]), d["config"]["workload"][:110])
    for k, w in d.get("workloads", {}).items():
        print("   %-9s value %.4e e2e %.4e ms/step %.3f" % (k, w["value"], (w.get("e2e") or {}).get("value", 0), w["ms_per_step"]))
PY
