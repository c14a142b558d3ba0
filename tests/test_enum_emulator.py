"""Kernel 2 on the CPU: the CTA programs of the enumeration kernel (pcetk_b200/csrc/pce_enum_core.cuh) and its host planner
(pce_enum_plan.h) compiled for the host by tests/emu and run phase by phase.  What is checked here without a GPU is what the
GPU executes: roles, tables, index arithmetic, bin scales and the finishing step -- against the oracle, <= 1e-12 relative
(BASELINE.json north_star), on every fixture the reference can enumerate, over the whole pH 0..20 grid in ONE pass."""

import numpy as np
import pytest

import pcetk_b200 as P
from pcetk_b200 import _native
from pcetk_b200.synthetic import synth_a
from conftest import load_fixture
from emu import analytic_bins, finish

RTOL = 1e-12
GRID = np.arange(0.0, 20.5, 0.5)


def rel_err(p, ref):
    mask = ref > 1e-250
    return np.abs(p[mask] / ref[mask] - 1.0).max()


def library_finish(rec, bins, grid, unfolded=False):
    """The product's own finishing step (host code of the library: needs no GPU) on bins from the emulator."""
    model = P.MEADModel(rec, log=None)
    handle, lib = model.energyModel._handle, _native.library()
    sigma = np.zeros(bins.shape[1])
    _native.check(lib.pce_analytic_bin_scales(handle, np.ascontiguousarray(grid, dtype=np.float64), len(grid), int(unfolded), sigma))
    p = np.zeros((len(grid), rec.ninstances))
    _native.check(lib.pce_analytic_finish(handle, np.ascontiguousarray(bins.ravel()), np.ascontiguousarray(grid, dtype=np.float64), len(grid), int(unfolded), 0.0,
                                          _native.ptr(p), None))
    return p, sigma


@pytest.mark.parametrize("name", ["twosites", "defensin"])
@pytest.mark.parametrize("mode", [2, 1, 0])
def test_emulated_kernel_vs_oracle_every_ph(port, name, mode):
    rec, _gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    bins, sigma, info = analytic_bins(rec, wsym, GRID, max_mode=mode)
    assert info["runs"] <= 3 and max(info["modes"]) <= mode
    p, library_sigma = library_finish(rec, bins, GRID)
    assert np.array_equal(sigma, library_sigma)                     # the emulator and the library plan the same scales
    assert np.abs(p - finish(rec, bins, sigma, GRID)).max() < 1e-15   # numpy longdouble restatement of the finishing step
    for row, ph in zip(p, GRID):
        ref, _z, _g = port.analytic(rec, wsym, float(ph))
        assert rel_err(row, ref) <= RTOL


@pytest.mark.parametrize("name,picks", [("lysozyme", [0, 14, 28, 40]), ("rubredoxin", [9, 27])])
def test_emulated_kernel_million_states(port, name, picks):
    """4.2e6 / 1.0e6 states: F sites (chunks over several emulated CTAs), O2 loop, two role rotations."""
    rec, _gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    bins, sigma, info = analytic_bins(rec, wsym, GRID, grid=5, grid_slots=64)
    assert info["runs"] <= 3 and info["modes"][0] == 2
    p = finish(rec, bins, sigma, GRID)
    for k in picks:
        ref, _z, _g = port.analytic(rec, wsym, float(GRID[k]))
        assert rel_err(p[k], ref) <= RTOL
    assert 1e-30 < bins[bins > 0].max() < 1e6                      # the bin scales keep the sums near 1 over 20 pH units


def test_emulated_shards_add_up(port):
    """Disjoint shards of the chunk range use common scales and add to the unsharded bins (the multi-GPU exchange)."""
    rec = synth_a(18)
    wsym = port.symmetrize(rec.interactions)
    grid = np.array([2.0, 7.0, 12.0])
    whole, sigma, _ = analytic_bins(rec, wsym, grid, grid=4, grid_slots=32)
    parts = sum(analytic_bins(rec, wsym, grid, shard=s, nshards=3, grid=3, grid_slots=32)[0] for s in range(3))
    mask = whole > 0
    assert np.abs(parts[mask] / whole[mask] - 1.0).max() < 1e-13
    p = finish(rec, parts, sigma, grid)
    ref, _z, _g = port.analytic(rec, wsym, 7.0)
    assert rel_err(p[1], ref) <= RTOL


def test_emulated_unfolded_and_generic_radices(port):
    """Unfolded state (no interactions) and a model whose sites are all three-instance (no canonical sites at all)."""
    rec, _gold = load_fixture("defensin")
    wsym = port.symmetrize(rec.interactions)
    bins, sigma, _ = analytic_bins(rec, wsym, [2.0, 7.0], unfolded=True)
    p = finish(rec, bins, sigma, np.array([2.0, 7.0]))
    ref, _z, _g = port.analytic(rec, wsym, 7.0, unfolded=True)
    assert rel_err(p[1], ref) <= RTOL
    rng = np.random.default_rng(3)
    nsites, per = 9, 3
    first = np.arange(0, nsites * per, per, dtype=np.int32)
    protons = np.tile([2, 1, 0], nsites).astype(np.int32)
    gintr = rng.normal(0.0, 3.0, nsites * per)
    w = rng.normal(0.0, 1.0, (nsites * per, nsites * per))
    w = 0.5 * (w + w.T)
    site_of = np.repeat(np.arange(nsites), per)
    w[site_of[:, None] == site_of[None, :]] = 0.0
    rec3 = P.ModelRecord(name="three", site_first=first, site_last=first + per - 1, protons=protons, gmodel=gintr.copy(), gintr=gintr, interactions=w,
                         temperature=300.0, site_labels=[("T", "S", k) for k in range(nsites)], inst_labels=["a", "b", "c"] * nsites)
    wsym3 = port.symmetrize(rec3.interactions)
    bins, sigma, info = analytic_bins(rec3, wsym3, GRID[::4], grid=3, grid_slots=4)
    assert info["modes"] == [0] * info["runs"]
    p = finish(rec3, bins, sigma, GRID[::4])
    for k in (0, 3, 7):
        ref, _z, _g = port.analytic(rec3, wsym3, float(GRID[::4][k]))
        assert rel_err(p[k], ref) <= RTOL
