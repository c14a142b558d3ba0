"""Parity of the two synthetic BASELINE configs at (or near) their named shapes, and of ifp20 over its whole curve.

* synth-A (32 coupled sites, 2^32 states) has no CPU oracle: (1) the answer must not depend on the order of the sites (a
  permutation puts other sites into the kernel's M / I / O / F roles); (2) an independent brute force on the GPU -- kernel 1
  energies of every state of the coupled 28-site model, summed by torch in fp64 -- must agree to 1e-12; (3) a Monte Carlo
  ensemble must agree with the enumeration to 0.01.
* synth-M (1000 sites) and a 200-site slice of the same generator: GPU ensembles against the mean of independent seeds of the
  REFERENCE sampler (oracle/_ref = the reference's MCModelDefault.c compiled unmodified; the C port when absent).
  Tolerance: north_star's |dp| <= 0.01 plus the statistical error of the reference mean (it is what it is: the reference does
  ~20 scans per second and core on the 1000-site model), and a z-score bound.
* ifp20: all 29 pH values of the fixture's curve against 16 seeds of the reference sampler, |dp| <= 0.01."""

import concurrent.futures
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle
import pcetk_b200 as P
from pcetk_b200 import _native
from pcetk_b200.formats import ModelRecord
from pcetk_b200.synthetic import synth_a, synth_m
from conftest import load_fixture

pytestmark = pytest.mark.gpu


def permuted(rec, order):
    """The same model with its sites listed in another order (instances move with their sites)."""
    inst = np.concatenate([np.arange(rec.site_first[s], rec.site_last[s] + 1) for s in order])
    sizes = [int(rec.site_last[s] - rec.site_first[s] + 1) for s in order]
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    last = (first + np.array(sizes) - 1).astype(np.int32)
    new = ModelRecord(name=rec.name + "_perm", site_first=first, site_last=last, protons=rec.protons[inst].copy(), gmodel=rec.gmodel[inst].copy(),
                      gintr=rec.gintr[inst].copy(), interactions=rec.interactions[np.ix_(inst, inst)].copy(), temperature=rec.temperature,
                      site_labels=[rec.site_labels[s] for s in order], inst_labels=[rec.inst_labels[k] for k in inst])
    return new, inst


def test_synthetic_32_sites_coupled_does_not_depend_on_the_site_order():
    rec = synth_a(32)
    grid = [0.0, 3.5, 7.0, 10.5, 14.0, 20.0]
    model = P.MEADModel(rec, log=None)
    model.energyModel.analyticStatesLimit = 2 ** 32
    p = model.CalculateProbabilitiesBatch(grid)
    assert np.abs(p.sum(axis=1) - rec.nsites).max() < 1e-9
    rng = np.random.default_rng(5)
    for order in (rng.permutation(32), np.arange(32)[::-1], np.roll(np.arange(32), 11)):
        other, inst = permuted(rec, order)
        m2 = P.MEADModel(other, log=None)
        m2.energyModel.analyticStatesLimit = 2 ** 32
        q = m2.CalculateProbabilitiesBatch(grid)
        mask = p[:, inst] > 1e-250
        assert np.abs(q[mask] / p[:, inst][mask] - 1.0).max() <= 1e-12


def test_synthetic_28_sites_coupled_vs_brute_force_with_kernel_1():
    """2^28 states, every site coupled to every other: Boltzmann sums over kernel-1 energies (reference summation order,
    bit-exact against the oracle in test_gpu_energy.py) of ALL states, accumulated by torch in fp64, batches combined with fsum."""
    import torch

    nsites = 28
    rec = synth_a(nsites)
    model = P.MEADModel(rec, log=None)
    model.energyModel.analyticStatesLimit = 2 ** 32
    lib, handle = _native.library(), model.energyModel._handle
    model.energyModel.Upload()
    beta = 1.0 / (0.001987165392 * rec.temperature)
    first = torch.tensor(np.asarray(rec.site_first), dtype=torch.int32, device="cuda")
    shifts = torch.arange(nsites, dtype=torch.int64, device="cuda")
    batch = 1 << 22
    for ph in (7.0, 2.0):
        d_ph = torch.tensor([ph], dtype=torch.float64, device="cuda")
        d_out = torch.zeros(batch, dtype=torch.float64, device="cuda")
        zs, s1, s0 = [], [], []
        eref = None
        for start in range(0, 1 << nsites, batch):
            index = torch.arange(start, start + batch, dtype=torch.int64, device="cuda")
            bits = ((index[:, None] >> shifts[None, :]) & 1)
            states = (first[None, :] + bits.to(torch.int32)).contiguous()
            _native.check(lib.pce_energy_device(handle, C.c_void_p(states.data_ptr()), batch, C.c_void_p(d_ph.data_ptr()), 1, 0, C.c_void_p(d_out.data_ptr())))
            model.energyModel.Synchronize()
            if eref is None:
                eref = float(d_out.min().item())
            w = torch.exp(-beta * (d_out - eref))
            fb = bits.to(torch.float64)
            zs.append(float(w.sum().item()))
            s1.append((fb.T @ w).cpu().numpy())
            s0.append(((1.0 - fb).T @ w).cpu().numpy())
        z = math.fsum(zs)
        brute = np.zeros(rec.ninstances)
        for i in range(nsites):
            brute[2 * i] = math.fsum(v[i] for v in s0) / z
            brute[2 * i + 1] = math.fsum(v[i] for v in s1) / z
        p = model.CalculateProbabilitiesBatch([ph])[0]
        mask = brute > 1e-250
        assert np.abs(p[mask] / brute[mask] - 1.0).max() <= 1e-12


def test_synthetic_32_sites_monte_carlo_vs_enumeration():
    rec = synth_a(32)
    model = P.MEADModel(rec, log=None)
    model.energyModel.analyticStatesLimit = 2 ** 32
    grid = [3.0, 7.0, 11.0]
    exact = model.CalculateProbabilitiesBatch(grid)
    sampler = P.MCModelDefault(nequil=500, nprod=20000, randomSeed=77, nchains=256)
    model.DefineMCModel(sampler, log=None)
    p = model.CalculateProbabilitiesBatch(grid)
    assert np.abs(p - exact).max() <= 0.01


# ---- reference-sampler ensembles ----------------------------------------------------------------------------------------
def reference_seeds(rec, ph_values, nequil, nprod, seeds):
    """p[seed, npH, ninst] from independent runs of the reference sampler, one host thread per seed (ctypes releases the GIL)."""
    use_ref = oracle.have_ref()

    def one(seed):
        if use_ref:
            ref = oracle.Ref(rec)
            mc = ref.mc_create(2.0, nequil, nprod, seed)
            out = np.array([ref.mc_run(mc, float(ph)) for ph in ph_values])
            ref.close()
            return out
        port = oracle.Port()
        wsym = port.symmetrize(rec.interactions)
        pairs = port.find_pairs(rec, wsym, 2.0)[:2]
        return port.mc_mt(rec, wsym, pairs, nequil, nprod, seed, ph_values)[0]

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(seeds), os.cpu_count() or 1)) as pool:
        return np.array(list(pool.map(one, seeds)))


def compare_with_reference_seeds(chains, p_seeds, nprod, dp_max, label):
    """GPU ensemble against the mean over seeds of the reference sampler, both run with the SAME protocol (equilibration and
    production scans per chain, random initial states), so that incomplete equilibration biases both alike: a two-sample test.
    `chains` = per-chain occupancy counts of the GPU run [npH, nchains, ninst]; their chain-to-chain scatter is the measured
    standard deviation of ONE chain's estimate (whatever the autocorrelation time is), so the reference mean over n seeds
    carries sigma / sqrt(n) and the GPU mean sigma / sqrt(nchains)."""
    per_chain = chains / float(nprod)
    p_gpu = per_chain.mean(axis=1)
    sigma = per_chain.std(axis=1, ddof=1)
    nseeds, nchains = p_seeds.shape[0], per_chain.shape[1]
    mean = p_seeds.mean(axis=0)
    dp = p_gpu - mean
    sem = sigma * np.sqrt(1.0 / nseeds + 1.0 / nchains)
    # rare visits (an instance seen in 3 of the reference's 6400 scans and hardly ever by a GPU chain) are counting noise, not
    # Gaussian scatter: the resolution of the reference mean, a few scans out of nseeds * nprod, floors the standard error
    z = dp / np.sqrt(sem ** 2 + (3.0 / (nseeds * nprod)) ** 2)
    assert np.abs(dp).max() <= dp_max, "%s: max |dp| %.4f" % (label, np.abs(dp).max())
    assert np.abs(dp).max() <= 0.01 + 4.0 * sem.max(), "%s: max |dp| %.4f against 0.01 + 4 sem (%.4f)" % (label, np.abs(dp).max(), sem.max())
    assert np.abs(z).max() <= 6.0, "%s: max |z| %.2f" % (label, np.abs(z).max())
    assert np.sqrt(np.mean(z ** 2)) <= 1.3, "%s: rms z %.2f" % (label, np.sqrt(np.mean(z ** 2)))
    # the seed-to-seed scatter of the reference must look like the chain-to-chain scatter of the GPU sampler
    visited = sigma > 0.02
    ratio = p_seeds.std(axis=0, ddof=1)[visited] / sigma[visited]
    assert 0.8 <= np.median(ratio) <= 1.25, "%s: median scatter ratio %.2f" % (label, np.median(ratio))
    return p_gpu


def test_synthetic_200_site_slice_vs_reference_sampler_seeds():
    """synth_m(200): 200 sites / 600 instances / 933 double-move pairs, the generator of BASELINE configs[4] at a size where
    the reference does ~900 scans per second: 16 seeds x (500 + 8000) scans at two pH values against 512 GPU chains of the
    same protocol.  |dp| <= 0.01 (+ 4 standard errors of the difference), every z-score within 6, rms z ~ 1."""
    rec = synth_m(200)
    grid = [4.0, 8.0]
    seeds = reference_seeds(rec, grid, 500, 8000, [1000 + 7 * k for k in range(16)])
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=500, nprod=8000, randomSeed=424242, nchains=512)
    model.DefineMCModel(sampler, log=None)
    out = sampler.RunChains(grid)
    compare_with_reference_seeds(out["counts"], seeds, 8000, 0.02, "synth_m(200)")


def test_synthetic_1000_sites_vs_reference_sampler_seeds():
    """BASELINE configs[4]'s model itself (1000 sites / 3000 instances / 5526 pairs) at pH 7: the reference manages ~20 scans per
    second and core, so it gets 8 seeds x (100 + 400) scans; the GPU runs 1024 chains of the same protocol (a chain 100 scans
    from a random start is not equilibrated on this model: both sides carry the same bias).  The reference mean carries up to
    ~0.1 of noise per instance, so the bounds that bite are the z-scores over the 3000 instances and the scatter ratio."""
    rec = synth_m(1000)
    seeds = reference_seeds(rec, [7.0], 100, 400, [31 + 5 * k for k in range(8)])
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=100, nprod=400, randomSeed=99, nchains=1024)
    model.DefineMCModel(sampler, log=None)
    out = sampler.RunChains([7.0])
    p = compare_with_reference_seeds(out["counts"], seeds, 400, 0.3, "synth_m(1000)")
    assert np.abs(p.sum() - rec.nsites) < 1e-6


def test_ifp20_whole_curve_vs_reference_sampler_seeds():
    """ifp20 (73 sites, 171 instances, 49 pairs; cannot be enumerated): all 29 pH values of the fixture's curve, 256 GPU chains x
    20000 scans against the mean of 16 seeds of the reference sampler at its defaults (500 + 20000 scans): |dp| <= 0.01."""
    rec, gold = load_fixture("ifp20")
    grid = np.asarray(gold["golden_curve_ph"], dtype=np.float64)
    assert len(grid) == 29
    seeds = reference_seeds(rec, grid, 500, 20000, [2000 + 3 * k for k in range(16)])
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=500, nprod=20000, randomSeed=5150, nchains=256)
    model.DefineMCModel(sampler, log=None)
    p = model.CalculateProbabilitiesBatch(grid)
    dp = p - seeds.mean(axis=0)
    assert np.abs(dp).max() <= 0.01
    assert np.sqrt(np.mean(dp ** 2)) <= 0.002
