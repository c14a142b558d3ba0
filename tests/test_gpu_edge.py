"""Edge cases of the hot path through the C ABI: ragged instance counts, single-instance sites, no pairs, tiny and
degenerate models, chain ensembles that straddle pH values inside one block.  Each case is checked against the oracle."""

import numpy as np
import pytest

import pcetk_b200 as P
from conftest import load_fixture
from pcetk_b200.formats import ModelRecord
from pcetk_b200.synthetic import synth_a, synth_m

pytestmark = pytest.mark.gpu


def rel_err(p, ref):
    mask = ref > 1e-250
    return np.abs(p[mask] / ref[mask] - 1.0).max()


def ragged_model(nsites=11, seed=3, tautomers=True):
    """2/3/5-instance sites from the synth-M generator; optionally make some instances tautomers (equal protons)."""
    rec = synth_m(nsites, seed=seed)
    if tautomers:
        protons = rec.protons.copy()
        for s in range(rec.nsites):
            if rec.site_last[s] - rec.site_first[s] + 1 == 3:
                protons[rec.site_first[s] + 1] = protons[rec.site_first[s]]     # two forms with the same proton count
        rec = ModelRecord(rec.name + "t", rec.site_first, rec.site_last, protons, rec.gmodel, rec.gintr * 0.2, rec.interactions * 0.2,
                          rec.temperature, rec.site_labels, rec.inst_labels)
    return rec


def single_instance_model():
    """A site with one instance in the middle of the model (the reference's MC would spin forever on it)."""
    base = synth_a(6)
    first = np.array([0, 2, 4, 5, 7, 9], dtype=np.int32)
    last = np.array([1, 3, 4, 6, 8, 10], dtype=np.int32)
    keep = [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11]
    w = base.interactions[np.ix_(keep, keep)]
    return ModelRecord("single", first, last, base.protons[keep], base.gmodel[keep], base.gintr[keep], w, 300.0,
                       [("SYN", "SIT", k + 1) for k in range(6)], ["i%d" % k for k in range(11)])


@pytest.mark.parametrize("nsites,tautomers", [(11, True), (11, False), (7, True), (1, False), (2, False)])
def test_analytic_ragged_models(port, nsites, tautomers):
    rec = ragged_model(nsites, tautomers=tautomers)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    grid = [0.0, 4.5, 7.0, 9.5, 14.0]
    p = model.CalculateProbabilitiesBatch(grid)
    for row, ph in zip(p, grid):
        ref, _z, _g = port.analytic(rec, wsym, ph)
        assert rel_err(row, ref) <= 1e-12
    pu = model.energyModel.CalculateProbabilitiesAnalyticallyBatch(grid, unfolded=True)
    ref, _z, _g = port.analytic(rec, wsym, 7.0, unfolded=True)
    assert rel_err(pu[2], ref) <= 1e-12


def test_analytic_single_instance_site(port):
    rec = single_instance_model()
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    p = model.CalculateProbabilitiesBatch([3.0, 7.0])
    for row, ph in zip(p, [3.0, 7.0]):
        ref, _z, _g = port.analytic(rec, wsym, ph)
        assert rel_err(row, ref) <= 1e-12
    assert abs(p[0, 4] - 1.0) < 1e-14                      # the lone instance is always occupied


@pytest.mark.parametrize("kernel", [1, 2])
def test_mc_single_instance_site_and_ragged(port, kernel):
    for rec in (single_instance_model(), ragged_model(9)):
        model = P.MEADModel(rec, log=None)
        wsym = port.symmetrize(rec.interactions)
        sampler = P.MCModelDefault(nequil=20, nprod=300, randomSeed=31337, nchains=37, kernel=kernel, doubleFlip=0.5)
        model.DefineMCModel(sampler, log=None)
        pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
        out = sampler.RunChains([5.0, 8.0])
        for q, ph in enumerate([5.0, 8.0]):
            for chain in (0, 17, 36):
                spec = port.mc_philox_chain(rec, wsym, pairs, 20, 300, sampler.seed, chain, q, ph)
                assert np.array_equal(out["counts"][q, chain], spec["counts"])
                assert out["counters"][q, chain].tolist() == spec["counters"].tolist()
        counts, _c, _e = sampler.RunCounts([5.0, 8.0])
        assert np.array_equal(counts, out["counts"].sum(axis=1))


@pytest.mark.parametrize("kernel", [1, 2])
def test_mc_model_without_any_choice(port, kernel):
    """Every site has a single instance: nothing can move (the reference's sampler would spin forever); every trial is a
    counted no-op and the single microstate has probability 1."""
    base = synth_a(4)
    keep = [0, 2, 4, 6]
    first = np.arange(4, dtype=np.int32)
    rec = ModelRecord("frozen", first, first.copy(), base.protons[keep], base.gmodel[keep], base.gintr[keep],
                      base.interactions[np.ix_(keep, keep)], 300.0, [("SYN", "SIT", k + 1) for k in range(4)], ["i%d" % k for k in range(4)])
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=5, nprod=50, randomSeed=3, nchains=33, kernel=kernel, doubleFlip=0.0)
    model.DefineMCModel(sampler, log=None)
    counts, counters, _esum = sampler.RunCounts([7.0])
    assert counts[0].tolist() == [33 * 50] * 4
    assert counters[0][1] == 0 and counters[0][3] == 0 and counters[0][0] + counters[0][2] == 33 * 50 * (4 + sampler.npairs)
    wsym = port.symmetrize(rec.interactions)
    pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
    out = sampler.RunChains([7.0])
    spec = port.mc_philox_chain(rec, wsym, pairs, 5, 50, sampler.seed, 7, 0, 7.0)
    assert np.array_equal(out["counts"][0, 7], spec["counts"]) and out["counters"][0, 7].tolist() == spec["counters"].tolist()


def test_mc_long_runs_use_32_bit_residence_stamps(port):
    """nprod >= 65536 switches the chain-per-thread kernel from 16-bit to 32-bit residence stamps: same chain, count for count."""
    rec, _gold = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    sampler = P.MCModelDefault(nequil=10, nprod=70000, randomSeed=99, nchains=3, kernel=1)
    model.DefineMCModel(sampler, log=None)
    pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
    out = sampler.RunChains([6.5])
    spec = port.mc_philox_chain(rec, wsym, pairs, 10, 70000, sampler.seed, 2, 0, 6.5)
    assert np.array_equal(out["counts"][0, 2], spec["counts"]) and out["counters"][0, 2].tolist() == spec["counters"].tolist()
    assert out["counts"][0, 2].sum() == 70000 * rec.nsites and out["esum"][0, 2] == spec["esum"]


def test_mc_no_pairs_and_all_pairs(port):
    rec = ragged_model(8)
    wsym = port.symmetrize(rec.interactions)
    for limit, expect in ((1e9, 0), (0.0, 28)):
        model = P.MEADModel(rec, log=None)
        sampler = P.MCModelDefault(nequil=10, nprod=200, randomSeed=5, nchains=33, doubleFlip=limit)
        model.DefineMCModel(sampler, log=None)
        assert sampler.npairs == expect
        pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
        out = sampler.RunChains([6.0])
        spec = port.mc_philox_chain(rec, wsym, pairs, 10, 200, sampler.seed, 32, 0, 6.0)
        assert np.array_equal(out["counts"][0, 32], spec["counts"])


def test_mc_few_chains_many_ph_in_one_block(port):
    """3 chains x 41 pH values: every block of the chain-major grid straddles many pH values."""
    rec = ragged_model(9)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    sampler = P.MCModelDefault(nequil=10, nprod=150, randomSeed=77, nchains=3)
    model.DefineMCModel(sampler, log=None)
    pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
    grid = np.arange(0.0, 20.5, 0.5)
    counts, counters, esum = sampler.RunCounts(grid)
    for q in (0, 13, 40):
        expect = sum(port.mc_philox_chain(rec, wsym, pairs, 10, 150, sampler.seed, c, q, float(grid[q]))["counts"] for c in range(3))
        assert np.array_equal(counts[q], expect)
    assert (counts.sum(axis=1) == 3 * 150 * rec.nsites).all()
    assert (counters[:, 0] + counters[:, 2] == 3 * 150 * (rec.nsites + sampler.npairs)).all()


def test_energy_large_model_bit_exact(port):
    rec = synth_m(60)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    rng = np.random.default_rng(9)
    states = np.stack([[rng.integers(f, l + 1) for f, l in zip(rec.site_first, rec.site_last)] for _ in range(300)]).astype(np.int32)
    energies = model.energyModel.CalculateMicrostateEnergies(states, pH=[1.0, 13.0])
    for s in (0, 150, 299):
        assert energies[s, 0] == port.energy(rec, wsym, states[s], 1.0)
        assert energies[s, 1] == port.energy(rec, wsym, states[s], 13.0)


@pytest.mark.parametrize("tpl,budget,cross", [("1", None, None), ("2", None, None), ("1", "0", None), ("1", None, "0")])
def test_warp_kernel_on_the_1000_site_model_matches_specification(port, monkeypatch, tpl, budget, cross):
    """BASELINE config 5 shape (1000 sites / 3000 instances, W = 72 MB): two short chains, count for count."""
    rec = synth_m(1000)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    sampler = P.MCModelDefault(nequil=1, nprod=3, randomSeed=2024, nchains=2)
    model.DefineMCModel(sampler, log=None)
    # trials per lane and round (1, 2) and the two ways of applying an accepted move (difference rows / W rows)
    monkeypatch.setenv("PCE_MC_TPL", tpl)
    if budget is not None:
        monkeypatch.setenv("PCE_DW_BUDGET_MB", budget)
    if cross is not None:
        monkeypatch.setenv("PCE_CROSS_TABLE", cross)         # gather the four W entries instead of the precomputed cross term
    assert sampler.LaunchPlan(1)["kernel"] == 2          # auto: W (72 MB) is not on chip -> chain per warp
    pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
    out = sampler.RunChains([7.0])
    for chain in (0, 1):
        spec = port.mc_philox_chain(rec, wsym, pairs, 1, 3, sampler.seed, chain, 0, 7.0)
        assert np.array_equal(out["counts"][0, chain], spec["counts"])
        assert np.array_equal(out["state"][0, chain], spec["state"])
    # the quiet build of the large-field kernel (no energy tracking) walks the same chains
    quiet, counters, esum = sampler.RunCounts([7.0], energy=False)
    assert np.array_equal(quiet[0], out["counts"][0].sum(axis=0)) and np.array_equal(counters[0], out["counters"][0].sum(axis=0)) and not esum.any()


@pytest.mark.parametrize("kernel", [1, 2])
def test_mc_one_site_model(port, kernel):
    """A model of ONE site: a scan is a single trial, so a round of the chain-per-warp kernel crosses up to 32 scan ends (and the
    production boundary somewhere inside a round)."""
    rec = ragged_model(1, tautomers=False)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    sampler = P.MCModelDefault(nequil=45, nprod=700, randomSeed=77, nchains=35, kernel=kernel)
    model.DefineMCModel(sampler, log=None)
    assert sampler.npairs == 0
    out = sampler.RunChains([4.0, 9.0])
    for q, ph in enumerate([4.0, 9.0]):
        for chain in (0, 34):
            spec = port.mc_philox_chain(rec, wsym, ([], []), 45, 700, sampler.seed, chain, q, ph)
            assert np.array_equal(out["counts"][q, chain], spec["counts"]) and out["counters"][q, chain].tolist() == spec["counters"].tolist()
            assert abs(out["esum"][q, chain] - spec["esum"]) <= 1e-9 * max(1.0, abs(spec["esum"]))
    counts, counters, _e = sampler.RunCounts([4.0, 9.0], energy=False)
    assert np.array_equal(counts, out["counts"].sum(axis=1)) and np.array_equal(counters, out["counters"].sum(axis=1))


def test_same_site_interactions_are_ignored(port):
    """W entries between instances of ONE site: the reference's energy function never reads them (EnergyModel.c:204-224), its
    Move does, its DoubleMove does not.  Here they are ignored throughout: a model with junk in those blocks walks exactly the
    chains of the clean model and enumerates to the same probabilities."""
    rec, _gold = load_fixture("defensin")
    junk = rec.interactions.copy()
    rng = np.random.default_rng(8)
    for f, l in zip(rec.site_first, rec.site_last):
        junk[f:l + 1, f:l + 1] = rng.normal(0.0, 3.0, (l - f + 1, l - f + 1))
    dirty = P.ModelRecord(name="defensin_junk", site_first=rec.site_first, site_last=rec.site_last, protons=rec.protons, gmodel=rec.gmodel,
                          gintr=rec.gintr, interactions=junk, temperature=rec.temperature, site_labels=rec.site_labels, inst_labels=rec.inst_labels)
    results = []
    for record in (rec, dirty):
        model = P.MEADModel(record, log=None)
        exact = model.CalculateProbabilitiesBatch([3.0, 7.0])
        for kernel in (1, 2):
            sampler = P.MCModelDefault(nequil=10, nprod=150, randomSeed=77, nchains=40, kernel=kernel)
            model.DefineMCModel(sampler, log=None)
            results.append((exact, sampler.RunChains([3.0, 7.0])["counts"]))
    for (exact_a, counts_a), (exact_b, counts_b) in zip(results[:2], results[2:]):
        assert np.array_equal(counts_a, counts_b)
        assert np.abs(exact_a / exact_b - 1.0).max() < 1e-13


def test_too_many_trials_per_chain_are_refused():
    """A chain counts its trials in 32 bits: (nequil + nprod) x (nsites + npairs) >= 2^32 is refused, not wrapped."""
    rec, _gold = load_fixture("lysozyme")
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=0, nprod=2 ** 31 - 1, randomSeed=1, nchains=1)
    model.DefineMCModel(sampler, log=None)
    with pytest.raises(P.CLibraryError, match="too many trials per chain"):
        sampler.RunCounts([7.0])


def test_reloading_the_same_model_keeps_its_tables_and_a_changed_model_does_not(port):
    """The tables derived from a model (enumeration plan, difference rows, cross terms) are keyed on the uploaded CONTENT:
    re-loading identical arrays reuses them, changing one energy rebuilds them -- the results must follow the model either way."""
    rec, _gold = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    em = model.energyModel
    wsym = port.symmetrize(rec.interactions)
    grid = [3.0, 7.0, 11.0]
    first = model.CalculateProbabilitiesBatch(grid)
    em.Load(gintr=rec.gintr, gmodel=rec.gmodel, protons=rec.protons, interactions=rec.interactions)
    em.SymmetrizeInteractions(log=None)
    assert np.array_equal(model.CalculateProbabilitiesBatch(grid), first)
    changed = ModelRecord(rec.name + "c", rec.site_first, rec.site_last, rec.protons, rec.gmodel, rec.gintr.copy(), rec.interactions.copy(),
                          rec.temperature, rec.site_labels, rec.inst_labels)
    changed.gintr[3] += 1.5
    changed.interactions[2, 9] += 0.75
    changed.interactions[9, 2] += 0.75
    em.Load(gintr=changed.gintr, interactions=changed.interactions)
    em.SymmetrizeInteractions(log=None)
    wchanged = port.symmetrize(changed.interactions)
    p = model.CalculateProbabilitiesBatch(grid)
    assert not np.array_equal(p, first)
    for row, ph in zip(p, grid):
        ref, _z, _g = port.analytic(changed, wchanged, ph)
        assert rel_err(row, ref) <= 1e-12
    # the same for the sampler's tables (both kernels): chains of the changed model, then of the original one again
    for kernel in (1, 2):
        sampler = P.MCModelDefault(nequil=10, nprod=150, randomSeed=4242, nchains=33, kernel=kernel)
        model.DefineMCModel(sampler, log=None)
        pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
        for current, w in ((changed, wchanged), (rec, wsym), (rec, wsym)):
            em.Load(gintr=current.gintr, gmodel=current.gmodel, protons=current.protons, interactions=current.interactions)
            em.SymmetrizeInteractions(log=None)
            out = sampler.RunChains([6.0], phIndexOffset=0)
            spec = port.mc_philox_chain(current, w, pairs, 10, 150, sampler.seed, 32, 0, 6.0)
            assert np.array_equal(out["counts"][0, 32], spec["counts"])
