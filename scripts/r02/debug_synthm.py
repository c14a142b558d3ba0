"""Debug: where do GPU chains and reference-sampler seeds disagree on the 1000-site model (same protocol)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import pcetk_b200 as P
from pcetk_b200.synthetic import synth_m
from test_gpu_parity_synthetic import reference_seeds

nsites = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nequil, nprod, nseeds = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rec = synth_m(nsites)
seeds = reference_seeds(rec, [7.0], nequil, nprod, [31 + 5 * k for k in range(nseeds)])[:, 0, :]
model = P.MEADModel(rec, log=None)
sampler = P.MCModelDefault(nequil=nequil, nprod=nprod, randomSeed=99, nchains=1024)
model.DefineMCModel(sampler, log=None)
out = sampler.RunChains([7.0])
per_chain = out["counts"][0] / float(nprod)
p = per_chain.mean(axis=0); sigma = per_chain.std(axis=0, ddof=1)
mean = seeds.mean(axis=0)
sem = sigma * np.sqrt(1.0 / nseeds + 1.0 / 1024)
z = (p - mean) / np.sqrt(sem ** 2 + 1e-10)
site_of = np.repeat(np.arange(rec.nsites), rec.site_last - rec.site_first + 1)
print(os.environ.get("PCE_DW_BUDGET_MB"), os.environ.get("PCE_CROSS_TABLE"), "max|z| %.2f rms z %.3f  rms dp %.4f" % (np.abs(z).max(), np.sqrt((z ** 2).mean()), np.sqrt(((p - mean) ** 2).mean())))
print("z histogram:", np.histogram(np.abs(z), bins=[0, 1, 2, 3, 4, 5, 6, 8, 100])[0])
for k in np.argsort(-np.abs(z))[:8]:
    s = site_of[k]
    print("inst %4d site %3d (n=%d, protons %d) gpu %.4f sigma %.4f | ref mean %.4f seeds %s | z %.2f" % (
        k, s, rec.site_last[s] - rec.site_first[s] + 1, rec.protons[k], p[k], sigma[k], mean[k], np.round(seeds[:, k], 3).tolist(), z[k]))
# acceptance statistics
c = out["counters"][0].sum(axis=0).astype(float)
print("gpu acceptance single %.4f double %.5f" % (c[1] / c[0], c[3] / c[2]))
