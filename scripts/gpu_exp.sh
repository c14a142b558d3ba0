# experiment: k_mc_thread v10 (accepting lane hands out the difference-row address; second site of a double move in a separate pass)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 --nprod 4000 --nequil 100"
run() { name=$1; shift; timeout 120 "$@" > gpurun_out/$name.log 2>&1; python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
for line in open("gpurun_out/%s.log" % name):
    if line.startswith("{"):
        d = json.loads(line); print(name, "%.4g" % d["value"], d["config"].get("launch"))
        break
else:
    print(name, "no JSON line:", open("gpurun_out/%s.log" % name).read()[-300:])
PY
}
run v10_l $B
run v10_d $B --workload defensin
