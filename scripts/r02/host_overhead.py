"""Where the host time of one titration curve goes (TitrationCurves.CalculateCurves on lysozyme, 64 chains): cProfile of 50 curves at
the reference's statistical effort, where the kernel takes ~2.4 ms."""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, ".")
import bench
import pcetk_b200 as P

record = bench.load_record("lysozyme")
cm = P.MEADModel(record, log=None)
cm.DefineMCModel(P.MCModelDefault(nequil=500, nprod=313, randomSeed=7, nchains=64), log=None)
for _ in range(3):
    P.TitrationCurves(cm, log=None, curveSampling=0.5, curveStart=0.0, curveStop=20.0).CalculateCurves(log=None)


def run(n=50):
    for _ in range(n):
        tc = P.TitrationCurves(cm, log=None, curveSampling=0.5, curveStart=0.0, curveStop=20.0)
        tc.CalculateCurves(log=None)
        tc.CalculateHalfpKs()


t0 = time.perf_counter(); run(); print("wall per curve %.3f ms" % ((time.perf_counter() - t0) * 20.0))
profile = cProfile.Profile(); profile.enable(); run(); profile.disable()
pstats.Stats(profile).sort_stats("cumulative").print_stats(22)
