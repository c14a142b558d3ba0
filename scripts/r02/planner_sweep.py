"""Round-2 experiment: MC kernel time against ensemble size, kernel and block size (input of the occupancy-aware planner).

    python scripts/r02/planner_sweep.py [model] > gpurun_out/planner_sweep_<model>.txt

For every chain count per pH value (41 pH values) the chain-per-thread kernel is timed with forced block sizes and the
chain-per-warp kernel with 2/3/4 resident CTAs; prints one line per configuration: moves/s and ms per launch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pcetk_b200 as P
from pcetk_b200 import _native

name = sys.argv[1] if len(sys.argv) > 1 else "lysozyme"
nprod = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
record = P.ModelRecord.from_npz_dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
model = P.MEADModel(record, log=None)
em = model.energyModel
stream = torch.cuda.current_stream()
em.SetStream(stream.cuda_stream)
lib = _native.library()
grid = np.arange(0.0, 20.0 + 1e-9, 0.5)
nph = len(grid)
dev = torch.device("cuda", 0)
d_counts = torch.zeros((nph, record.ninstances), dtype=torch.int64, device=dev)
d_counters = torch.zeros((nph, 4), dtype=torch.int64, device=dev)
d_esum = torch.zeros(nph, dtype=torch.float64, device=dev)


def run(sampler):
    _native.check(lib.pce_mc_run_device(sampler._handle, grid, nph, 0, 0, C.c_void_p(d_counts.data_ptr()), C.c_void_p(d_counters.data_ptr()),
                                        C.c_void_p(d_esum.data_ptr())))


def timed(sampler, reps=2):
    run(sampler)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        run(sampler)
        b.record(stream)
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


print("# model %s nsites %d ninst %d nprod %d nequil 500" % (name, record.nsites, record.ninstances, nprod))
for nchains in [int(x) for x in os.environ.get("SWEEP_CHAINS", "1,4,16,64,128,256,512,1024,2772").split(",")]:
    sampler = P.MCModelDefault(nequil=500, nprod=nprod, randomSeed=5, nchains=nchains)
    model.DefineMCModel(sampler, log=None)
    nmoves = record.nsites + sampler.npairs
    units = nph * nchains * (500 + nprod) * nmoves
    configs = [("auto", 0, {})]
    for block in (32, 64, 96, 128, 192, 256, 384, 512, 768):
        if block * 148 <= 4 * nph * nchains or block == 32:
            configs.append(("thread b%d" % block, 1, {"PCE_MC_BLOCK": str(block)}))
    for ctas in (2, 3, 4):
        configs.append(("warp c%d" % ctas, 2, {"PCE_MC_CTAS": str(ctas)}))
    for label, kernel, env in configs:
        for key in ("PCE_MC_BLOCK", "PCE_MC_CTAS"):
            os.environ.pop(key, None)
        os.environ.update(env)
        sampler.kernel = kernel
        sampler._configure()
        try:
            ms = timed(sampler)
            plan = sampler.LaunchPlan(nph)
            print("chains/pH %5d total %7d  %-12s kernel %d block %4d ctas %d : %9.3f ms  %.3e moves/s" % (
                nchains, nph * nchains, label, plan["kernel"], plan["block"], plan["ctas_per_sm"], ms, units / ms * 1e3), flush=True)
        except Exception as error:      # a forced plan that does not fit
            print("chains/pH %5d %-12s failed: %s" % (nchains, label, error), flush=True)
