"""Kernel 2 (exhaustive enumeration) through the C ABI: <= 1e-12 relative against the oracle (the north-star bar),
<= 5e-7 against the reference's curve files, plus the size-independent properties used at sizes the CPU cannot reach."""

import ctypes as C

import numpy as np
import pytest

import pcetk_b200 as P
from pcetk_b200 import _native
from pcetk_b200.synthetic import synth_a
from conftest import load_fixture

pytestmark = pytest.mark.gpu

RTOL = 1e-12      # BASELINE.json north_star: analytic probabilities within 1e-12 relative


def rel_err(p, ref):
    mask = ref > 1e-250
    return np.abs(p[mask] / ref[mask] - 1.0).max()


@pytest.mark.parametrize("name", ["twosites", "defensin"])
def test_analytic_all_ph_vs_oracle_and_golden(port, name):
    rec, gold = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    grid = gold["golden_curve_ph"]
    p = model.CalculateProbabilitiesBatch(grid)
    assert np.abs(p - gold["golden_curve_analytic"]).max() <= 5.1e-7
    for row, ph in zip(p, grid):
        ref, _z, _g = port.analytic(rec, wsym, float(ph))
        assert rel_err(row, ref) <= RTOL
    assert np.abs(p.reshape(len(grid), -1).sum(axis=1) - rec.nsites).max() < 1e-9


@pytest.mark.parametrize("name,picks", [("lysozyme_full_test", [0, 9, 14, 28]), ("rubredoxin", [0, 9, 14, 28]), ("lysozyme", [4, 14])])
def test_analytic_million_states_vs_oracle_and_golden(port, name, picks):
    rec, gold = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    grid = gold["golden_curve_ph"]
    p = model.CalculateProbabilitiesBatch(grid)
    assert np.abs(p - gold["golden_curve_analytic"]).max() <= 5.1e-7      # all 29 points against the curve files
    for k in picks:
        ref, _z, _g = port.analytic(rec, wsym, float(grid[k]))
        assert rel_err(p[k], ref) <= RTOL


def test_wide_ph_grid_in_one_pass(port):
    """pH 0..20 (BASELINE config 2 grid): every proton-count bin carries its own scale, so ONE enumeration pass serves the
    whole grid (round 1 needed two); results must not depend on what else is in the pH list."""
    rec, _ = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    grid = np.arange(0.0, 20.5, 0.5)
    assert len(model.energyModel._phGroups(grid, 15)) == 1
    p = model.CalculateProbabilitiesBatch(grid)
    for k in (0, 13, 27, 40):
        ref, _z, _g = port.analytic(rec, wsym, float(grid[k]))
        assert rel_err(p[k], ref) <= RTOL
    single = model.CalculateProbabilitiesBatch([grid[27]])
    assert rel_err(single[0], p[27]) <= 1e-13
    # a grid wider than one pass can serve is split by the host layer and still matches
    wide = np.array([-30.0, 0.0, 7.0, 14.0, 45.0])
    assert len(model.energyModel._phGroups(wide, 15)) > 1
    pw = model.CalculateProbabilitiesBatch(wide)
    ref, _z, _g = port.analytic(rec, wsym, 7.0)
    assert rel_err(pw[2], ref) <= RTOL


@pytest.mark.parametrize("name", ["defensin", "lysozyme"])
def test_generic_accumulation_path_matches(port, name, monkeypatch):
    """PCE_ENUM_GENERIC=1 forces the kernel variant for non-canonical sites (run-time proton groups, bins in shared memory) on
    models that normally take the compile-time variant: same answers."""
    rec, _ = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    fast = P.MEADModel(rec, log=None).CalculateProbabilitiesBatch([2.0, 7.0, 12.0])
    monkeypatch.setenv("PCE_ENUM_GENERIC", "1")
    generic = P.MEADModel(rec, log=None).CalculateProbabilitiesBatch([2.0, 7.0, 12.0])
    ref, _z, _g = port.analytic(rec, wsym, 7.0)
    assert rel_err(generic[1], ref) <= RTOL
    mask = fast > 1e-250
    assert np.abs(generic[mask] / fast[mask] - 1.0).max() <= 1e-12


def test_unfolded_and_partition_functions(port):
    rec, _ = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    for ph in (2.0, 7.0):
        pu = model.energyModel.CalculateProbabilitiesAnalyticallyBatch([ph], unfolded=True)[0]
        ref, zu, _g = port.analytic(rec, wsym, ph, unfolded=True)
        assert rel_err(pu, ref) <= RTOL
        _p, zf, _gf = port.analytic(rec, wsym, ph)
        assert abs(model.energyModel.CalculateZfolded(pH=ph) / zf - 1.0) <= 1e-11
        assert abs(model.energyModel.CalculateZunfolded(pH=ph) / zu - 1.0) <= 1e-11
    # single-pH entry points keep the reference's contract: return nstates, leave p in the model
    assert model.energyModel.CalculateProbabilitiesAnalytically(pH=7.0) == 32768
    ref, _z, _g = port.analytic(rec, wsym, 7.0)
    assert rel_err(model.energyModel.GetProbabilities(), ref) <= RTOL
    sites = model.CalculateProbabilities(pH=7.0, isCalculateCurves=True, log=None)
    assert len(sites) == rec.nsites and abs(sum(sites[0]) - 1.0) < 1e-12


def test_state_limit_and_preconditions():
    rec, _ = load_fixture("ifp20")
    model = P.MEADModel(rec, log=None)
    with pytest.raises(P.CLibraryError, match="Maximum number of states"):
        model.energyModel.CalculateProbabilitiesAnalytically(pH=7.0)
    rec, _ = load_fixture("lysozyme_full_test")
    model = P.MEADModel(rec, log=None)
    model.energyModel.analyticStatesLimit = 1000
    with pytest.raises(P.CLibraryError, match="Maximum number of states"):
        model.CalculateProbabilities(pH=7.0, log=None)


def test_shards_add_up():
    """Partial bins of disjoint shards sum to the single-shard bins: the only exchange the multi-GPU path needs."""
    import torch

    rec, _ = load_fixture("rubredoxin")
    model = P.MEADModel(rec, log=None)
    lib, handle = _native.library(), model.energyModel._handle
    n = lib.pce_analytic_bins_size(handle)
    grid = np.array([5.0, 7.0, 9.0])

    def bins(shard, nshards):
        out = torch.zeros(n, dtype=torch.float64, device="cuda")
        _native.check(lib.pce_analytic_bins(handle, grid, len(grid), 0, shard, nshards, 0.0, C.c_void_p(out.data_ptr())))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    whole = bins(0, 1)
    parts = sum(bins(s, 3) for s in range(3))
    mask = whole > 0
    assert np.abs(parts[mask] / whole[mask] - 1.0).max() < 1e-13
    p = np.zeros((len(grid), rec.ninstances))
    _native.check(lib.pce_analytic_finish(handle, parts, grid, len(grid), 0, 0.0, _native.ptr(p), None))
    q = model.CalculateProbabilitiesBatch(grid)
    assert np.abs(p / q - 1.0).max() < 1e-12


def test_synthetic_26_sites_vs_compensated_oracle(port):
    """2^26 states (the reference's own cap): GPU against the reference's algorithm with compensated (long double) sums,
    held to north_star's 1e-12 -- the reference's naive serial double sum over 6.7e7 terms carries ~1e-12 of rounding itself."""
    rec = synth_a(26)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    grid = [4.0, 7.0, 10.0]
    p = model.CalculateProbabilitiesBatch(grid)
    ref, _z, _g = port.analytic_compensated(rec, wsym, 7.0)
    assert rel_err(p[1], ref) <= RTOL
    assert np.abs(p.sum(axis=1) - rec.nsites).max() < 1e-9


def test_synthetic_32_sites_factorised():
    """2^32 states (BASELINE config 4).  The last 6 sites are decoupled from the rest, so their probabilities are the
    closed-form two-state values and the first 26 sites must reproduce the 26-site model exactly."""
    rec26 = synth_a(26)
    rec32 = synth_a(32, decouple_from=26)
    small = P.MEADModel(rec26, log=None)
    big = P.MEADModel(rec32, log=None)
    big.energyModel.analyticStatesLimit = 2 ** 32
    grid = [3.0, 7.0, 11.0]
    p26 = small.CalculateProbabilitiesBatch(grid)
    p32 = big.CalculateProbabilitiesBatch(grid)
    assert np.abs(p32[:, :52] / p26 - 1.0).max() <= 1e-11
    c = 0.001987165392 * 300.0 * 2.302585092994
    beta = 1.0 / (0.001987165392 * 300.0)
    for q, ph in enumerate(grid):
        for s in range(26, 32):
            g0 = rec32.gintr[2 * s] + rec32.protons[2 * s] * c * ph
            g1 = rec32.gintr[2 * s + 1] + rec32.protons[2 * s + 1] * c * ph
            expect = 1.0 / (1.0 + np.exp(-beta * (g1 - g0)))
            assert abs(p32[q, 2 * s] / expect - 1.0) <= 1e-11


def test_half_pks_match_reference_table():
    """lysozyme.out prints pK1/2 (1 decimal) from its analytic curves; same post-processing on the GPU curves."""
    rec, gold = load_fixture("lysozyme")
    model = P.MEADModel(rec, log=None)
    curves = P.TitrationCurves(model, curveSampling=0.5, curveStart=0.0, curveStop=14.0, log=None)
    curves.CalculateCurves(log=None)
    curves.CalculateHalfpKs()
    golden = gold["golden_halfpk"]
    for site in model.sites:
        for instance in site.instances:
            mine = curves.halves[site.siteIndex][instance.instIndex]
            expect = golden[instance._instIndexGlobal]
            if np.isnan(expect):
                assert mine == []
            else:
                assert abs(mine[0] - expect) <= 0.051
