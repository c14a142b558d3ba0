"""StateVector: the microstate, one active instance per titratable site.

Mirrors the Python-visible class of ``extensions/cython/ContinuumElectrostatics.StateVector.pyx:14-207``
over the native object ``pce_state`` (``pcetk_b200/csrc/pce_state.cu``), which replaces
``extensions/csource/StateVector.c`` (odometer increment :370-384, substate increment :389-404, item access
:236-330) behind the same entry points as ``StateVector.h:53-84``.  The state lives on the host -- it is a few
integers -- and reaches the GPU as the int32 row of global instance indices the native object holds
(:meth:`GlobalIndices`); :class:`EnergyModel` ships such rows in batches.
"""

import ctypes as C

import numpy as np

from . import _native
from .errors import CLibraryError
from .log import logFile, LogFileActive


class StateVector(object):
    """A class defining the state vector."""

    def __init__(self, ceModel):
        """Constructor: site ranges come from the model's instances (StateVector.pyx:80-103)."""
        self._lib = _native.library()
        self._handle = None
        nsites = ceModel.nsites
        handle = C.c_void_p()
        try:
            _native.check(self._lib.pce_state_create(nsites, C.byref(handle)))
        except CLibraryError:
            raise CLibraryError("Cannot allocate state vector.")
        self._handle = handle
        for indexSite, site in enumerate(ceModel.sites):
            indices = [instance._instIndexGlobal for instance in site.instances]
            _native.check(self._lib.pce_state_set_site(handle, indexSite, min(indices), max(indices)))
        self.isOwner = False
        self.owner = ceModel

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle:
            self._lib.pce_state_destroy(handle)
            self._handle = None

    def __len__(self):
        return self._lib.pce_state_nsites(self._handle)

    def __getitem__(self, index):
        """Local index of the active instance (StateVector_GetItem, StateVector.c:236-249)."""
        item = C.c_int(-1)
        _native.check(self._lib.pce_state_get_item(self._handle, int(index), C.byref(item)))
        return item.value

    def __setitem__(self, index, value):
        """StateVector_SetItem, StateVector.c:254-274."""
        _native.check(self._lib.pce_state_set_item(self._handle, int(index), int(value)))

    def GetActualItem(self, index):
        """Global index of the active instance (StateVector_GetActualItem, StateVector.c:279-292)."""
        item = C.c_int(-1)
        _native.check(self._lib.pce_state_get_actual_item(self._handle, int(index), C.byref(item)))
        return item.value

    def SetActualItem(self, index, value):
        """StateVector_SetActualItem, StateVector.c:297-313."""
        _native.check(self._lib.pce_state_set_actual_item(self._handle, int(index), int(value)))

    def CopyTo(self, other):
        """In-place copy (a no-op ``pass`` in the reference's Python class, StateVector.pyx:106-108; StateVector_CopyTo here)."""
        _native.check(self._lib.pce_state_copy_to(self._handle, other._handle))

    def Increment(self):
        """Odometer step, site 0 fastest; False (and the all-first state) after the last state
        (StateVector_Increment, StateVector.c:370-384)."""
        more = C.c_int(0)
        _native.check(self._lib.pce_state_increment(self._handle, C.byref(more)))
        return bool(more.value)

    def Reset(self):
        """Set all sites to their first instances (StateVector_Reset, StateVector.c:126-133)."""
        _native.check(self._lib.pce_state_reset(self._handle))

    def ResetToMaximum(self):
        """Set all sites to their last instances (StateVector_ResetToMaximum, StateVector.c:150-157)."""
        _native.check(self._lib.pce_state_reset_to_maximum(self._handle))

    def Randomize(self, seed=None):
        """StateVector_Randomize (StateVector.c:163-170): a uniformly drawn instance per site, from Philox words keyed by
        ``seed`` (drawn from NumPy's entropy pool when not given; the reference seeds its generator from the clock)."""
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).view(np.uint64)[0])
        _native.check(self._lib.pce_state_randomize(self._handle, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def DefineSubstate(self, selectedSites):
        """|selectedSites| is a sequence of (segmentName, residueSerial) (StateVector.pyx:163-191)."""
        chosen = []
        for selectedSegment, selectedSerial in selectedSites:
            for siteIndex, site in enumerate(self.owner.sites):
                if site.segName == selectedSegment and site.resSerial == selectedSerial:
                    chosen.append(siteIndex)
                    break
            else:
                raise CLibraryError("Site %s %d not found." % (selectedSegment, selectedSerial))
        try:
            _native.check(self._lib.pce_state_allocate_substate(self._handle, len(chosen)))     # refuses a second call, like the reference
        except CLibraryError:
            raise CLibraryError("Memory allocation failure.")
        for substateIndex, siteIndex in enumerate(chosen):
            _native.check(self._lib.pce_state_set_substate_item(self._handle, siteIndex, substateIndex))

    def IncrementSubstate(self):
        """StateVector_IncrementSubstate, StateVector.c:389-404."""
        more = C.c_int(0)
        _native.check(self._lib.pce_state_increment_substate(self._handle, C.byref(more)))
        return bool(more.value)

    def ResetSubstate(self):
        """StateVector_ResetSubstate, StateVector.c:138-145."""
        _native.check(self._lib.pce_state_reset_substate(self._handle))

    def IsSubstate(self, index):
        flag = C.c_int(0)
        _native.check(self._lib.pce_state_is_substate(self._handle, int(index), C.byref(flag)))
        return bool(flag.value)

    def GetSubstateItem(self, index):
        """Index of the site in slot ``index`` of the substate (StateVector_GetSubstateItem, StateVector.c:318-327)."""
        site = C.c_int(-1)
        _native.check(self._lib.pce_state_get_substate_item(self._handle, int(index), C.byref(site)))
        return site.value

    # ---- bulk forms: rows for the kernels ---------------------------------------------------------------------
    def GlobalIndices(self):
        """The state as an int32 array of global instance indices (what the C ABI takes)."""
        out = np.zeros(len(self), dtype=np.int32)
        _native.check(self._lib.pce_state_get_active(self._handle, out))
        return out

    def SetGlobalIndices(self, indices):
        """Set every site at once from global instance indices (each checked against its site's range)."""
        row = np.ascontiguousarray(indices, dtype=np.int32)
        if row.shape != (len(self), ):
            raise CLibraryError("Invalid value (%d)." % row.size)
        _native.check(self._lib.pce_state_set_active(self._handle, row))

    def SetIndex(self, stateIndex):
        """Go to the state ``stateIndex`` increments after the all-first state (mixed-radix decode, site 0 fastest)."""
        _native.check(self._lib.pce_state_set_index(self._handle, int(stateIndex)))

    def GetIndex(self):
        index = C.c_longlong(0)
        _native.check(self._lib.pce_state_get_index(self._handle, C.byref(index)))
        return index.value

    def SubstateRows(self):
        """Every state of the substate in ``IncrementSubstate`` order as rows of global indices -> int32[nsubstates, nsites]
        (the loop of Substate.py:99-108, batched for kernel 1).  Leaves the substate at its first state."""
        count = C.c_longlong(0)
        _native.check(self._lib.pce_state_enumerate_substate(self._handle, 0, None, C.byref(count)))
        rows = np.zeros((count.value, len(self)), dtype=np.int32)
        _native.check(self._lib.pce_state_enumerate_substate(self._handle, count.value, _native.ptr(rows), C.byref(count)))
        return rows

    def Print(self, verbose=True, title=None, log=logFile):
        """StateVector.pyx:131-159."""
        if LogFileActive(log):
            rows = []
            for indexSite, site in enumerate(self.owner.sites):
                local = self[indexSite]
                instance = site.instances[local]
                rows.append((site.segName, site.resName, "%d" % site.resSerial, instance.label, "%d" % instance.instIndex,
                             "@" if self.IsSubstate(indexSite) else " "))
            log.Table(rows, title=title if title else "State vector")
