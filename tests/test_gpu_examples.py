"""The example scripts (the reference's tests/<name>/<name>.py from the point where MEAD has finished) run end to end
on the GPU and their output files reproduce the reference's own result files."""

import importlib
import io
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_fixture

pytestmark = pytest.mark.gpu

EXAMPLES = os.path.join(ROOT, "examples")


def run_example(name, directory):
    from pcetk_b200.log import TextLog

    if EXAMPLES not in sys.path:
        sys.path.insert(0, EXAMPLES)
    module = importlib.import_module(name)
    out = io.StringIO()
    result = module.main(str(directory), log=TextLog(out))
    return result, out.getvalue()


def read_curves(directory, record):
    from pcetk_b200 import formats

    return formats.read_curves_directory(str(directory), record)


def test_example_twosites(tmp_path):
    rec, gold = load_fixture("twosites")
    result, text = run_example("twosites", tmp_path)
    assert np.abs(np.array(result["energies"]) - gold["golden_gmicro"]).max() <= 5.1e-7
    assert text.count("Gmicro = ") == 8 and "Found 1 pair of strongly interacting sites" in text
    ph, analytic = read_curves(tmp_path / "curves_analytic", rec)
    assert np.allclose(ph, gold["golden_curve_ph"])
    assert np.abs(analytic - gold["golden_curve_analytic"]).max() <= 1.01e-6     # two %f roundings
    _ph, mc = read_curves(tmp_path / "curves_mc", rec)
    assert np.abs(mc - gold["golden_curve_analytic"]).max() <= 0.01


def test_example_lysozyme(tmp_path):
    rec, gold = load_fixture("lysozyme")
    result, text = run_example("lysozyme", tmp_path)
    _ph, analytic = read_curves(tmp_path / "curves_analytic", rec)
    assert np.abs(analytic - gold["golden_curve_analytic"]).max() <= 1.01e-6
    _ph, mc = read_curves(tmp_path / "curves_mc", rec)
    assert np.abs(mc - gold["golden_curve_analytic"]).max() <= 0.01
    assert "Found 3 pairs of strongly interacting sites" in text
    with open(str(tmp_path / "his_repl.sed")) as handle:
        sed = handle.read()
    assert "s/H../HS" in sed                                          # HIS15 gets its most probable tautomer
    back = __import__("pcetk_b200").formats.model_from_gmct_tables("lysozyme", str(tmp_path / "gintr.dat"), str(tmp_path / "W.dat"))
    assert np.abs(back.gintr - rec.gintr).max() <= 5.1e-5 and np.abs(back.interactions - rec.interactions).max() <= 5.1e-5


def test_example_ifp20(tmp_path):
    rec, gold = load_fixture("ifp20")
    result, text = run_example("ifp20", tmp_path)
    _ph, mc = read_curves(tmp_path / "curves", rec)
    assert np.abs(mc - gold["golden_curve_mc_curves"]).max() <= 0.03      # against one 20 000-scan run of the reference
    assert "Found 49 pairs of strongly interacting sites" in text
    substate = result["substate"]
    energies = np.array([e - substate.zeroEnergy for e, _ in substate.substates])
    assert np.abs(energies - gold["golden_substate_energy"]).max() <= 0.0051
    with open(str(tmp_path / "table.tex")) as handle:
        assert len(handle.read().splitlines()) == 80 + 5


def test_example_lyso_mc_traj(tmp_path):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "lysozyme_traj.npz"))
    result, text = run_example("lyso_mc_traj", tmp_path)
    energies = np.loadtxt(str(tmp_path / "stat_mc.dat"))
    assert energies.shape == gold["stat_mc"].shape
    assert abs(energies.min() - gold["stat_mc"].min()) < 1e-5           # both runs find the global minimum
    assert abs(energies.mean() - gold["stat_mc"].mean()) < 0.08
    rows = result["sampler"].lastLogRows
    assert rows.shape == gold["counters"].shape and np.array_equal(rows[:, 0], gold["counters"][:, 0])
    assert "** Equilibration phase done **" in text and "** Production phase done **" in text
