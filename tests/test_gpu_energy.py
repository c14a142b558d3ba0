"""Kernel 1 (microstate energy) through the C ABI against the oracle and the reference's golden Gmicro values."""

import numpy as np
import pytest

import pcetk_b200 as P
from pcetk_b200 import _native
from conftest import load_fixture

pytestmark = pytest.mark.gpu

ALL = ["twosites", "defensin", "lysozyme_full_test", "rubredoxin", "lysozyme", "ifp20"]


def random_states(rec, n, seed):
    rng = np.random.default_rng(seed)
    return np.stack([[rng.integers(f, l + 1) for f, l in zip(rec.site_first, rec.site_last)] for _ in range(n)]).astype(np.int32)


def test_device_philox_known_answers(port):
    lib = _native.library()
    ctr = np.array([0, 0, 0, 0, 0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], dtype=np.uint32)
    out = np.zeros(8, dtype=np.uint32)
    _native.check(lib.pce_test_philox(0, ctr[:4].copy(), 1, np.array([0, 0], dtype=np.uint32), out[:4]))
    assert out[:4].tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    key = np.array([0xA4093822, 0x299F31D0], dtype=np.uint32)
    both = np.zeros(8, dtype=np.uint32)
    _native.check(lib.pce_test_philox(0, ctr, 2, key, both))
    assert both[4:].tolist() == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    rng = np.random.default_rng(3)
    many = rng.integers(0, 2 ** 32, size=4 * 257, dtype=np.uint64).astype(np.uint32)
    out = np.zeros_like(many)
    _native.check(lib.pce_test_philox(0, many, 257, key, out))
    for k in (0, 100, 256):
        assert out[4 * k:4 * k + 4].tolist() == port.philox(many[4 * k:4 * k + 4], key).tolist()


@pytest.mark.parametrize("name", ["twosites", "defensin", "lysozyme_full_test", "rubredoxin"])
def test_gmicro_golden(name):
    """Energies of the first states in odometer order, as printed by the reference's scripts (1e-6 files)."""
    rec, gold = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    vector = P.StateVector(model)
    states = []
    for _ in range(len(gold["golden_gmicro"])):
        states.append(vector.GlobalIndices())
        vector.Increment()
    energies = model.energyModel.CalculateMicrostateEnergies(np.stack(states), pH=7.0)
    assert np.abs(energies - gold["golden_gmicro"]).max() <= 5.1e-7
    vector.Reset()
    assert abs(model.CalculateMicrostateEnergy(vector, pH=7.0) - gold["golden_gmicro"][0]) <= 5.1e-7


@pytest.mark.parametrize("name", ALL)
def test_energy_bit_exact_vs_oracle(port, name):
    rec, _ = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    wsym = port.symmetrize(rec.interactions)
    assert np.array_equal(model.energyModel.GetSymmetricMatrix(), wsym)
    states = random_states(rec, 97, 11)
    grid = [0.0, 3.3, 7.0, 11.9, 20.0]
    energies = model.energyModel.CalculateMicrostateEnergies(states, pH=grid)
    unfolded = model.energyModel.CalculateMicrostateEnergies(states, pH=grid, unfolded=True)
    for s in (0, 1, 50, 96):
        for q, ph in enumerate(grid):
            assert energies[s, q] == port.energy(rec, wsym, states[s], ph)
            assert unfolded[s, q] == port.energy(rec, wsym, states[s], ph, unfolded=True)


def test_energy_input_validation():
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    with pytest.raises(P.CLibraryError):
        model.energyModel.CalculateMicrostateEnergies(np.array([[0, 3]], dtype=np.int32))   # instance 3 belongs to site 0
    empty = model.energyModel.CalculateMicrostateEnergies(np.zeros((0, 2), dtype=np.int32), pH=[7.0])
    assert empty.shape == (0, 1)


def test_substate_matches_reference_table(port):
    """defensin.out substate table: energies relative to the lowest, 2 decimals (Substate.py)."""
    rec, _ = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    substate = P.Substate(model, (("PRTA", "ASP", 2), ("PRTA", "GLU", 14), ("PRTA", "ARG", 15)), pH=7.0, log=None)
    substate.CalculateSubstateEnergies(log=None)
    assert len(substate.substates) == 8
    wsym = port.symmetrize(rec.interactions)
    vector = substate.vector
    for energy, indices in substate.substates:
        for site_index, local in zip(substate.indicesOfSites, indices):
            vector[site_index] = local
        assert energy == port.energy(rec, wsym, vector.GlobalIndices(), 7.0)
    assert substate.substates[0][0] == min(e for e, _ in substate.substates)


@pytest.mark.parametrize("name,use_mc", [("defensin", False), ("ifp20", True)])
def test_substate_tables_golden(name, use_mc):
    """The substate tables printed in defensin.out (8 rows) and ifp20.out (80 rows: 4 x 5 x 2 x 2 forms of HIS260, the BLF
    chromophore and two propionates): relative energies to the 2 decimals of the log, proton counts and instance labels."""
    rec, gold = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    if use_mc:      # ifp20 cannot be enumerated; the reference also locates the base state from MC probabilities
        model.DefineMCModel(P.MCModelDefault(nequil=500, nprod=5000, randomSeed=3, nchains=256), log=None)
    sites = [tuple(s.split("|")) for s in gold["golden_substate_sites"].tolist()]
    substate = P.Substate(model, [(a, b, int(c)) for a, b, c in sites], pH=7.0, log=None)
    substate.CalculateSubstateEnergies(log=None)
    energies = np.array([e - substate.zeroEnergy for e, _ in substate.substates])
    assert len(energies) == len(gold["golden_substate_energy"])
    assert np.abs(energies - gold["golden_substate_energy"]).max() <= 0.0051
    for (energy, indices), labels, protons in zip(substate.substates, gold["golden_substate_labels"].tolist(), gold["golden_substate_protons"].tolist()):
        got, nprot = [], 0
        for site_index, local in zip(substate.indicesOfSites, indices):
            instance = model.sites[site_index].instances[local]
            got.append(instance.label)
            nprot += instance.protons
        assert "|".join(got) == labels and nprot == protons
