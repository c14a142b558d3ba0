# round 2, call 7: new parity tests (synthetic configs, ifp20 whole curve, compensated oracle) + probes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
t0=$SECONDS; timeout 1500 python -m pytest tests/test_gpu_parity_synthetic.py tests/test_gpu_analytic.py -q --timeout 900 -p no:cacheprovider --durations=12 > gpurun_out/r02_pytest7.log 2>&1; echo "pytest exit $? in $((SECONDS-t0)) s" >> gpurun_out/r02_pytest7.log
tail -40 gpurun_out/r02_pytest7.log
python - <<'PY'
import ctypes as C
from pcetk_b200 import _native
lib = _native.library()
for kind, name in ((0, "dfma 1e9/s"), (1, "shared GB/s"), (2, "issue 1e9 warp-instr/s")):
    v = C.c_double(0)
    _native.check(lib.pce_probe_on_chip(0, kind, C.byref(v)))
    print(name, v.value)
PY
