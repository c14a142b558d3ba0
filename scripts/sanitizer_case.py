"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / synccheck): defensin fixture, few chains,
few scans, both MC kernels, the enumeration and the energy kernel.  Usage: compute-sanitizer --tool racecheck python scripts/sanitizer_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcetk_b200 as P

record = P.ModelRecord.from_npz_dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "defensin.npz")))
model = P.MEADModel(record, log=None)
states = np.stack([np.array(record.site_first) for _ in range(4)]).astype(np.int32)
print("energies", model.energyModel.CalculateMicrostateEnergies(states, pH=[7.0])[:, 0])
print("analytic", model.CalculateProbabilitiesBatch([7.0])[0][:4])
for kernel in (1, 2):
    sampler = P.MCModelDefault(nequil=5, nprod=40, randomSeed=3, nchains=40, kernel=kernel)
    model.DefineMCModel(sampler, log=None)
    out = sampler.RunChains([3.0, 7.0])
    counts, counters, esum = sampler.RunCounts([3.0, 7.0])
    assert np.array_equal(counts, out["counts"].sum(axis=1))
    print("kernel", kernel, "counters", counters.tolist())
print("done")
