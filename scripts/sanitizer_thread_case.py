"""compute-sanitizer case for the chain-per-thread MC kernel: both short-field paths (defensin: two chains per round on
half-warps; lysozyme: one chain per round on 32 lanes), per-chain and aggregate entry points.
Usage: compute-sanitizer --tool racecheck python scripts/sanitizer_thread_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcetk_b200 as P

for name in ("defensin", "lysozyme"):
    record = P.ModelRecord.from_npz_dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    model = P.MEADModel(record, log=None)
    sampler = P.MCModelDefault(nequil=5, nprod=40, randomSeed=3, nchains=40, kernel=1)
    model.DefineMCModel(sampler, log=None)
    out = sampler.RunChains([3.0, 7.0])
    counts, counters, esum = sampler.RunCounts([3.0, 7.0])
    assert np.array_equal(counts, out["counts"].sum(axis=1))
    print(name, "counters", counters.tolist())
print("done")
