"""The native state vector (pce_state, include/pcetk_b200.h) against the reference's own StateVector.c, function by
function: same return values, same error class, same state after every call."""

import ctypes as C

import numpy as np
import pytest

import oracle
import pcetk_b200 as P
from pcetk_b200 import _native
from conftest import load_fixture

# pce_status of the native call <-> Status of the reference (oracle/shim/Status.h)
STATUS = {_native.PCE_OK: "Continue", _native.PCE_ERR_INDEX: "IndexOutOfRange", _native.PCE_ERR_VALUE: "ValueError",
          _native.PCE_ERR_ALLOC: "MemoryAllocationFailure"}


class Native(object):
    """pce_state through ctypes, returning (status name, value) like ``oracle.Ref.vector_call``."""

    def __init__(self, model):
        self.lib = _native.library()
        self.handle = C.c_void_p()
        _native.check(self.lib.pce_state_create_for_model(model.energyModel._handle, C.byref(self.handle)))

    def close(self):
        self.lib.pce_state_destroy(self.handle)

    def call(self, name, *args):
        fn = getattr(self.lib, "pce_state_" + name)
        if name in ("get_item", "get_actual_item", "is_substate", "get_substate_item", "increment", "increment_substate"):
            value = C.c_int(0)
            status = fn(self.handle, *args, C.byref(value))
            return STATUS[status], value.value
        return STATUS[fn(self.handle, *args)], None

    def state(self):
        out = np.zeros(self.lib.pce_state_nsites(self.handle), dtype=np.int32)
        _native.check(self.lib.pce_state_get_active(self.handle, out))
        return out


needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (reference sources absent)")


@needs_ref
@pytest.mark.parametrize("name", ["twosites", "defensin", "ifp20"])
def test_random_call_sequences_match_the_reference(name):
    rec, _ = load_fixture(name)
    model = P.MEADModel(rec, log=None)
    ref, nat = oracle.Ref(rec), Native(model)
    names = oracle.Ref.STATUS_NAMES
    rng = np.random.default_rng(20241020)
    nsites, ninst = rec.nsites, rec.ninstances
    # a substate of three sites (one of them the last site), defined through the same calls on both sides
    chosen = sorted(rng.choice(nsites - 1, size=min(2, nsites - 1), replace=False).tolist()) + [nsites - 1]
    assert names[ref.vector_call("allocate_substate", len(chosen))] == nat.call("allocate_substate", len(chosen))[0] == "Continue"
    assert names[ref.vector_call("allocate_substate", 1)] == nat.call("allocate_substate", 1)[0] == "MemoryAllocationFailure"
    for slot, site in enumerate(chosen):
        assert names[ref.vector_call("set_substate_item", site, slot)] == nat.call("set_substate_item", site, slot)[0] == "Continue"
    assert names[ref.vector_call("set_substate_item", 0, len(chosen))] == nat.call("set_substate_item", 0, len(chosen))[0] == "IndexOutOfRange"
    assert names[ref.vector_call("set_substate_item", nsites, 0)] == nat.call("set_substate_item", nsites, 0)[0] == "ValueError"
    for slot in range(-1, len(chosen) + 1):
        status, value = ref.vector_call("get_substate_item", slot)
        assert (names[status], value) == nat.call("get_substate_item", slot)
    for step in range(4000):
        op = int(rng.integers(0, 11))
        site = int(rng.integers(-1, nsites + 1))
        if op == 0:
            local = int(rng.integers(-1, 6))
            assert names[ref.vector_call("set_item", site, local)] == nat.call("set_item", site, local)[0]
        elif op == 1:
            value = int(rng.integers(-1, ninst + 1))
            assert names[ref.vector_call("set_actual_item", site, value)] == nat.call("set_actual_item", site, value)[0]
        elif op in (2, 3, 4):
            call = ("get_item", "get_actual_item", "is_substate")[op - 2]
            status, value = ref.vector_call(call, site)
            assert (names[status], value) == nat.call(call, site)
        elif op in (5, 6):
            for _ in range(int(rng.integers(1, 40))):
                assert ref.vector_call("increment") == nat.call("increment")[1]
        elif op == 7:
            for _ in range(int(rng.integers(1, 12))):
                assert ref.vector_call("increment_substate") == nat.call("increment_substate")[1]
        elif op == 8:
            ref.vector_call("reset_substate"); nat.call("reset_substate")
        elif op == 9 and step % 7 == 0:
            ref.vector_call("reset_to_maximum"); nat.call("reset_to_maximum")
        elif op == 10 and step % 11 == 0:
            ref.vector_call("reset"); nat.call("reset")
        assert np.array_equal(ref.vector_state(), nat.state()), "state differs after step %d (op %d)" % (step, op)
    nat.close()
    ref.close()


@needs_ref
def test_full_odometer_cycle_matches_the_reference():
    """All 32768 states of defensin in the reference's order, the wrap-around and the state after it (StateVector.c:370-384)."""
    rec, _ = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    ref, nat = oracle.Ref(rec), Native(model)
    count = 0
    while True:
        count += 1
        if count % 97 == 0 or count > 32760:
            assert np.array_equal(ref.vector_state(), nat.state())
        r, n = ref.vector_call("increment"), nat.call("increment")[1]
        assert r == n
        if not r:
            break
    assert count == rec.nstates == 32768
    assert np.array_equal(ref.vector_state(), nat.state()) and np.array_equal(nat.state(), rec.site_first)
    nat.close()
    ref.close()


def test_state_index_is_the_number_of_increments():
    rec, _ = load_fixture("lysozyme")
    model = P.MEADModel(rec, log=None)
    v, w = P.StateVector(model), P.StateVector(model)
    for index in (0, 1, 2, 3, 4, 11, 12, 4095, 4096):
        w.SetIndex(index)
        assert w.GetIndex() == index
    v.Reset()
    for index in range(1, 300):
        assert v.Increment()
        w.SetIndex(index)
        assert np.array_equal(v.GlobalIndices(), w.GlobalIndices()) and v.GetIndex() == index
    w.SetIndex(rec.nstates - 1)
    v.ResetToMaximum()
    assert np.array_equal(v.GlobalIndices(), w.GlobalIndices()) and v.GetIndex() == rec.nstates - 1
    with pytest.raises(P.CLibraryError, match="beyond the number of states"):
        w.SetIndex(rec.nstates)
    assert np.array_equal(w.GlobalIndices(), rec.site_first)
    with pytest.raises(P.CLibraryError, match="negative"):
        w.SetIndex(-1)
    # the 1000-site synthetic model has more states than 63 bits hold: the index of its last state is refused, small ones work
    from pcetk_b200.synthetic import synth_m
    big = P.StateVector(P.MEADModel(synth_m(1000), log=None))
    big.SetIndex(12345678901)
    assert big.GetIndex() == 12345678901
    big.ResetToMaximum()
    with pytest.raises(P.CLibraryError, match="63 bits"):
        big.GetIndex()


def test_substate_rows_equal_the_increment_loop():
    rec, gold = load_fixture("ifp20")
    model = P.MEADModel(rec, log=None)
    sites = [tuple(s.split("|")) for s in gold["golden_substate_sites"].tolist()]
    pairs = [(seg, int(serial)) for seg, _res, serial in sites]
    a, b = P.StateVector(model), P.StateVector(model)
    for vector in (a, b):
        vector.SetIndex(987654321)                   # some state outside the substate: it must be preserved
        vector.DefineSubstate(pairs)
    with pytest.raises(P.CLibraryError, match="Memory allocation failure"):
        a.DefineSubstate(pairs)
    rows = a.SubstateRows()
    b.ResetSubstate()
    expect = []
    more = True
    while more:
        expect.append(b.GlobalIndices())
        more = b.IncrementSubstate()
    assert rows.shape == (80, rec.nsites) and np.array_equal(rows, np.stack(expect))
    assert np.array_equal(a.GlobalIndices(), b.GlobalIndices())          # both back at the first substate
    assert [a.GetSubstateItem(k) for k in range(4)] == [i for i, s in enumerate(model.sites) if (s.segName, s.resSerial) in pairs]
    c = P.StateVector(model)
    a.CopyTo(c)
    assert np.array_equal(c.GlobalIndices(), a.GlobalIndices()) and c.IsSubstate(a.GetSubstateItem(0))
    with pytest.raises(P.CLibraryError, match="different lengths"):
        a.CopyTo(P.StateVector(P.MEADModel(load_fixture("twosites")[0], log=None)))


def test_randomize_and_bulk_setters():
    rec, _ = load_fixture("ifp20")
    model = P.MEADModel(rec, log=None)
    v = P.StateVector(model)
    seen = np.zeros((rec.nsites, 6), dtype=int)
    for seed in range(400):
        v.Randomize(seed)
        row = v.GlobalIndices()
        assert (row >= rec.site_first).all() and (row <= rec.site_last).all()
        seen[np.arange(rec.nsites), row - rec.site_first] += 1
    span = rec.site_last - rec.site_first + 1
    for i in range(rec.nsites):
        assert (seen[i, :span[i]] > 400 / span[i] * 0.55).all() and seen[i, span[i]:].sum() == 0      # every instance is drawn
    v.Randomize(7)
    first = v.GlobalIndices()
    v.Randomize(7)
    assert np.array_equal(first, v.GlobalIndices())                    # a seed names a state
    v.Randomize()
    v.SetGlobalIndices(rec.site_last)
    assert np.array_equal(v.GlobalIndices(), rec.site_last)
    bad = rec.site_first.copy()
    bad[3] = rec.site_last[3] + 1
    with pytest.raises(P.CLibraryError, match=r"Invalid value \(%d\)" % bad[3]):
        v.SetGlobalIndices(bad)
    assert np.array_equal(v.GlobalIndices(), rec.site_last)            # a refused row changes nothing
