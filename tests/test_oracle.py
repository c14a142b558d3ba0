"""Pin the CPU oracle: (1) against the reference's own golden vectors (tests/golden, extracted from the
fixtures' results.tgz), (2) against oracle/_ref -- the reference's C compiled unmodified -- bit for bit,
where that library is present (it is built in the build container and shipped to the GPU box)."""

import numpy as np
import pytest

import oracle
from oracle import oracle_np
from conftest import load_fixture

ANALYTIC_FIXTURES = ["twosites", "defensin", "lysozyme_full_test", "rubredoxin", "lysozyme"]
ALL_FIXTURES = ANALYTIC_FIXTURES + ["ifp20"]


def test_philox_known_answers(port):
    # Random123 known-answer vectors for philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, expect in kat:
        assert port.philox(ctr, key).tolist() == expect


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_symmetry_check_matches_logs(port, name):
    rec, _ = load_fixture(name)
    ok, dev = port.check_symmetric(rec.interactions, 0.05)
    # ifp20.out: "WARNING: Maximum deviation of interactions is 0.0533"; all other logs: symmetric within 0.05
    assert ok == (name != "ifp20")
    if name == "ifp20":
        assert round(dev, 4) == 0.0533
    if name == "twosites":
        assert round(dev, 4) == 0.0406


@pytest.mark.parametrize("name", ["defensin", "lysozyme_full_test", "rubredoxin"])
def test_symmetrised_w_matches_job_inter(port, name):
    rec, gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    assert np.abs(wsym - gold["golden_job_wsym"]).max() <= 5.1e-9     # 8-decimal file
    assert np.abs(rec.gintr - gold["golden_job_gint"]).max() <= 5.1e-9
    assert (rec.protons == gold["golden_job_protons"]).all()


@pytest.mark.parametrize("name", ["twosites", "defensin", "lysozyme_full_test", "rubredoxin"])
def test_gmicro_golden(port, name):
    rec, gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    golden = gold["golden_gmicro"]
    energies = port.enumerate_energies(rec, wsym, 7.0, len(golden))
    assert np.abs(energies - golden).max() <= 5.1e-7                  # "%f" precision of the logs


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_pairs_golden(port, name):
    rec, gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    a, b, wmax = port.find_pairs(rec, wsym, 2.0)
    assert np.stack([a, b], axis=1).tolist() == gold["golden_pairs"].tolist()
    assert np.abs(wmax - gold["golden_pairs_wmax"]).max() <= 5.1e-7


@pytest.mark.parametrize("name", ["twosites", "defensin"])
def test_analytic_curves_golden_all_ph(port, name):
    rec, gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    for ph, row in zip(gold["golden_curve_ph"], gold["golden_curve_analytic"]):
        p, _z, _gmin = port.analytic(rec, wsym, float(ph))
        assert np.abs(p - row).max() <= 5.1e-7


@pytest.mark.parametrize("name", ["lysozyme_full_test", "rubredoxin", "lysozyme"])
def test_analytic_curves_golden_some_ph(port, name):
    rec, gold = load_fixture(name)
    wsym = port.symmetrize(rec.interactions)
    ph_grid = gold["golden_curve_ph"]
    picks = [0, 14] if name == "lysozyme" else [0, 9, 14, 28]          # lysozyme: 4.2M states per point
    for k in picks:
        p, _z, _gmin = port.analytic(rec, wsym, float(ph_grid[k]))
        assert np.abs(p - gold["golden_curve_analytic"][k]).max() <= 5.1e-7


def test_ph7_tables_golden(port):
    rec, gold = load_fixture("twosites")
    wsym = port.symmetrize(rec.interactions)
    p, _z, _gmin = port.analytic(rec, wsym, 7.0)
    assert np.abs(p - gold["golden_prob_tables_ph7"][0]).max() <= 5.1e-5   # 4-decimal table


def test_numpy_restatement_agrees(port):
    rec, gold = load_fixture("defensin")
    wsym = port.symmetrize(rec.interactions)
    ph = [0.0, 7.0, 12.5]
    p_np = oracle_np.analytic(rec, wsym, ph)
    for q, value in enumerate(ph):
        p, _z, _gmin = port.analytic(rec, wsym, value)
        assert np.abs(p_np[q] / p - 1.0).max() <= 1e-12


def test_mc_mt_statistics_vs_analytic(port):
    """Reference-semantics MC (MT19937 port) against the exact answer: the reference's own accuracy class."""
    rec, _ = load_fixture("defensin")
    wsym = port.symmetrize(rec.interactions)
    a, b, _ = port.find_pairs(rec, wsym, 2.0)
    p_exact, _z, _g = port.analytic(rec, wsym, 7.0)
    p_mc, counters, _ = port.mc_mt(rec, wsym, (a, b), 500, 20000, 12345, [7.0])
    assert np.abs(p_mc[0] - p_exact).max() <= 0.02
    assert counters[0, 0] + counters[0, 2] == 20000 * (rec.nsites + len(a))


def test_philox_chain_spec_statistics(port):
    """The GPU chain specification samples the same distribution (many short chains vs exact)."""
    rec, _ = load_fixture("twosites")
    wsym = port.symmetrize(rec.interactions)
    a, b, _ = port.find_pairs(rec, wsym, 2.0)
    p_exact, _z, _g = port.analytic(rec, wsym, 4.0)
    total = np.zeros(rec.ninstances)
    nchains, nprod = 64, 2000
    for chain in range(nchains):
        out = port.mc_philox_chain(rec, wsym, (a, b), 100, nprod, 2024, chain, 0, 4.0)
        assert out["counts"].sum() == nprod * rec.nsites
        total += out["counts"]
    assert np.abs(total / (nchains * nprod) - p_exact).max() <= 0.01


# ------------------------------------------------------------------------------------------------
# against the compiled reference (present in the build container and, prebuilt, on the GPU box)
# ------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_port_equals_reference_energy_and_pairs(port, name):
    rec, _ = load_fixture(name)
    ref = oracle.Ref(rec)
    wsym = port.symmetrize(rec.interactions)
    assert np.array_equal(ref.wsym(), wsym)
    assert ref.check_symmetric(0.05) == port.check_symmetric(rec.interactions, 0.05)
    rng = np.random.default_rng(5)
    for _ in range(20):
        state = np.array([rng.integers(f, l + 1) for f, l in zip(rec.site_first, rec.site_last)], dtype=np.int32)
        ph = float(rng.uniform(0, 14))
        assert ref.energy(state, ph) == port.energy(rec, wsym, state, ph)
        assert ref.energy(state, ph, unfolded=True) == port.energy(rec, wsym, state, ph, unfolded=True)
    mc = ref.mc_create(2.0, 10, 10, 1)
    ra, rb, rw = ref.mc_pairs(mc)
    pa, pb, pw = port.find_pairs(rec, wsym, 2.0)
    assert ra.tolist() == pa.tolist() and rb.tolist() == pb.tolist() and np.array_equal(rw, pw)
    ref.close()


@needs_ref
@pytest.mark.parametrize("name", ["twosites", "defensin", "rubredoxin"])
def test_port_equals_reference_analytic_bitwise(port, name):
    rec, _ = load_fixture(name)
    ref = oracle.Ref(rec)
    wsym = port.symmetrize(rec.interactions)
    for ph in (0.0, 3.5, 7.0, 14.0):
        p, z, _gmin = port.analytic(rec, wsym, ph)
        assert np.array_equal(ref.analytic(ph), p)
        assert ref.z(ph) == z
        pu, zu, _ = port.analytic(rec, wsym, ph, unfolded=True)
        assert np.array_equal(ref.analytic(ph, unfolded=True), pu)
        assert ref.z(ph, unfolded=True) == zu
    ref.close()


@needs_ref
@pytest.mark.parametrize("name", ["twosites", "lysozyme", "ifp20"])
def test_port_equals_reference_mc_bitwise(port, name):
    """Same MT19937 seed -> identical probabilities, per-scan energies and move counters."""
    rec, _ = load_fixture(name)
    ref = oracle.Ref(rec)
    wsym = port.symmetrize(rec.interactions)
    pairs = port.find_pairs(rec, wsym, 2.0)[:2]
    nequil, nprod, seed = 50, 400, 777
    mc = ref.mc_create(2.0, nequil, nprod, seed)
    p_ref, e_ref, c_ref = ref.mc_run_trajectory(mc, 6.5, nprod)
    p_port, c_port, e_port = port.mc_mt(rec, wsym, pairs, nequil, nprod, seed, [6.5], want_energies=True)
    assert np.array_equal(p_ref, p_port[0])
    assert np.array_equal(e_ref, e_port[0])
    assert c_ref.tolist() == c_port[0].tolist()
    ref.close()


@needs_ref
def test_reference_mc_mean_energy_matches_stat_mc_fixture(port):
    """lysozyme_traj/stat_mc.dat: 20000 per-scan energies at pH 7; its mean (-17.3828) and minimum are a
    scalar known answer for the sampler.  The compiled reference reproduces them statistically."""
    rec, _ = load_fixture("lysozyme")
    gold = np.load("tests/golden/lysozyme_traj.npz")
    ref = oracle.Ref(rec)
    mc = ref.mc_create(2.0, 500, 20000, 4242)
    _p, energies, _c = ref.mc_run_trajectory(mc, 7.0, 20000)
    assert abs(energies.mean() - gold["stat_mc"].mean()) < 0.05
    assert abs(energies.min() - gold["stat_mc"].min()) < 1e-5
    ref.close()


COMPARE_FIXTURES = ["lysozyme_compare_7LYZ", "lysozyme_compare_2LZT"]


@pytest.mark.parametrize("name", COMPARE_FIXTURES)
def test_lysozyme_compare_models_are_pinned(port, name):
    """tests/lysozyme_compare (two lysozyme structures in one run): the models rebuilt from its MEAD logs reproduce the
    Gintr table, the symmetry verdict and the pair list of its log, and the exact probabilities of the oracle lie within
    single-run Monte Carlo noise of the reference's own curve files (two 20000-scan runs of one model differ by 0.021)."""
    rec, gold = load_fixture(name)
    assert np.abs(np.round(rec.gintr, 4) - gold["golden_gintr_4dp"]).max() < 1e-9
    ok, _dev = port.check_symmetric(rec.interactions, 0.05)
    assert ok
    wsym = port.symmetrize(rec.interactions)
    a, b, wmax = port.find_pairs(rec, wsym, 2.0)
    assert np.stack([a, b], axis=1).tolist() == gold["golden_pairs"].tolist()
    assert np.abs(wmax - gold["golden_pairs_wmax"]).max() <= 5.1e-7
    for pick in (4, 13, 22):
        p, _z, _g = port.analytic(rec, wsym, float(gold["golden_curve_ph"][pick]))
        assert np.abs(p - gold["golden_curve_mc_curves"][pick]).max() <= 0.03


def test_two_reference_runs_of_one_model_size_the_mc_tolerance():
    """lysozyme/curves_mc and lysozyme_compare/curves/7LYZ are two runs of the reference's sampler on the SAME model (same
    MEAD outputs, clock-seeded generator): their largest difference is what 'statistical parity' with one reference run means."""
    (rec_a, gold_a), (rec_b, gold_b) = load_fixture("lysozyme"), load_fixture("lysozyme_compare_7LYZ")
    assert np.array_equal(rec_a.gintr, rec_b.gintr) and np.array_equal(rec_a.interactions, rec_b.interactions)
    diff = np.abs(gold_a["golden_curve_mc_curves_mc"] - gold_b["golden_curve_mc_curves"])
    assert 0.015 < diff.max() < 0.03 and np.sqrt((diff ** 2).mean()) < 0.004
