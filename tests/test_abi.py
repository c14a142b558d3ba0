"""The C-ABI library loads on a CPU-only box and exports every symbol include/pcetk_b200.h declares; host-only entry
points (setters, getters, symmetry, pair finding) work without a GPU; compute entry points fail loudly without one."""

import ctypes as C
import os
import re

import numpy as np
import pytest

import pcetk_b200 as P
from pcetk_b200 import _native
from conftest import ROOT, load_fixture


def header_symbols():
    with open(os.path.join(ROOT, "include", "pcetk_b200.h")) as handle:
        text = re.sub(r"/\*.*?\*/", "", handle.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pce_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_native.LIBRARY_PATH)
    names = header_symbols()
    assert len(names) >= 50
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_native.SIGNATURES) == names       # the ctypes table and the header agree


def test_model_setters_getters_and_errors():
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    em = model.energyModel
    assert em.GetGintr(0) == rec.gintr[0] and em.GetProtons(0) == 2 and em.GetGmodel(1) == rec.gmodel[1]
    assert em.GetInteraction(0, 5) == rec.interactions[0, 5]
    assert em.GetInteractionSymmetric(0, 5) == 0.5 * (rec.interactions[0, 5] + rec.interactions[5, 0])
    assert em.GetDeviation(0, 5) == (rec.interactions[0, 5] + rec.interactions[5, 0]) * .5 - rec.interactions[0, 5]
    assert em.CheckIfSymmetric(0.05) == (True, pytest.approx(0.04062, abs=1e-5))
    assert em.CheckIfSymmetric(0.01)[0] is False
    em.SetProbability(3, 0.25)
    assert em.GetProbability(3) == 0.25 and model.sites[0].instances[3].probability == 0.25
    with pytest.raises(P.CLibraryError, match="out of range"):
        em.GetGintr(6)
    with pytest.raises(P.CLibraryError, match="out of range"):
        em.SetInteraction(0, -1, 1.0)
    em.ScaleInteractions(2.0)
    assert em.GetInteractionSymmetric(0, 5) == rec.interactions[0, 5] + rec.interactions[5, 0]
    em.ResetInteractions()
    assert em.GetInteractionSymmetric(0, 5) == 0.0
    assert em.NumberOfStates() == 8.0 and em.nstates == 8


def test_pairs_found_on_host_match_reference_logs():
    for name in ("twosites", "lysozyme", "defensin", "ifp20"):
        rec, gold = load_fixture(name)
        model = P.MEADModel(rec, log=None)
        sampler = P.MCModelDefault(nprod=10)
        model.DefineMCModel(sampler, log=None)
        assert [[a, b] for a, b, _w in sampler.GetPairs()] == gold["golden_pairs"].tolist()


def test_compute_fails_loudly_without_gpu():
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    with pytest.raises(P.CLibraryError):
        model.CalculateMicrostateEnergy(P.StateVector(model))
    with pytest.raises(P.CLibraryError):
        model.CalculateProbabilities(pH=7.0, log=None)
    sampler = P.MCModelDefault(nprod=10)
    model.DefineMCModel(sampler, log=None)
    with pytest.raises(P.CLibraryError):
        model.CalculateProbabilities(pH=7.0, log=None)


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import pcetk_b200, pcetk_b200.synthetic; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for base, _dirs, files in os.walk(os.path.join(ROOT, "pcetk_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(base, name)) as handle:
                    text = handle.read()
                assert "import oracle" not in text and "from oracle" not in text and "libpce_oracle" not in text, name
