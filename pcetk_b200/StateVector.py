"""StateVector: the microstate, one active instance per titratable site.

Mirrors the Python-visible class of ``extensions/cython/ContinuumElectrostatics.StateVector.pyx:14-207``
over the semantics of ``extensions/csource/StateVector.c`` (odometer increment :370-384, substate
increment :389-404, item access :236-330).  It is pure host logic: the state is a few integers that
are shipped to the GPU in batches by :class:`EnergyModel` -- keeping a device-resident copy of a
twenty-byte object would be slower than passing it.
"""

import numpy as np

from .errors import CLibraryError
from .log import logFile, LogFileActive


class StateVector(object):
    """A class defining the state vector."""

    def __init__(self, ceModel):
        """Constructor: site ranges come from the model's instances (StateVector.pyx:80-103)."""
        nsites = ceModel.nsites
        self._first = np.zeros(nsites, dtype=np.int32)
        self._last = np.zeros(nsites, dtype=np.int32)
        for indexSite, site in enumerate(ceModel.sites):
            indices = [instance._instIndexGlobal for instance in site.instances]
            self._first[indexSite] = min(indices)
            self._last[indexSite] = max(indices)
        self._active = self._first.copy()
        self._isSubstate = np.zeros(nsites, dtype=bool)
        self._substateSites = None
        self.isOwner = False
        self.owner = ceModel

    def __len__(self):
        return int(self._active.shape[0])

    def _checkIndex(self, index):
        if index < 0 or index >= len(self):
            raise CLibraryError("Index (%d) out of range." % index)

    def __getitem__(self, index):
        """Local index of the active instance (StateVector_GetItem, StateVector.c:236-249)."""
        self._checkIndex(index)
        return int(self._active[index] - self._first[index])

    def __setitem__(self, index, value):
        """StateVector_SetItem, StateVector.c:254-274."""
        self._checkIndex(index)
        actual = int(value) + int(self._first[index])
        if actual < self._first[index] or actual > self._last[index]:
            raise CLibraryError("Invalid value (%d)." % value)
        self._active[index] = actual

    def GetActualItem(self, index):
        """Global index of the active instance (StateVector_GetActualItem, StateVector.c:279-292)."""
        self._checkIndex(index)
        return int(self._active[index])

    def SetActualItem(self, index, value):
        """StateVector_SetActualItem, StateVector.c:297-313."""
        self._checkIndex(index)
        if value < self._first[index] or value > self._last[index]:
            raise CLibraryError("Invalid value (%d)." % value)
        self._active[index] = value

    def CopyTo(self, other):
        """In-place copy (a no-op ``pass`` in the reference, StateVector.pyx:106-108; implemented here)."""
        other._active[:] = self._active

    def Increment(self):
        """Odometer step, site 0 fastest; False (and the all-first state) after the last state
        (StateVector_Increment, StateVector.c:370-384)."""
        for i in range(len(self)):
            if self._active[i] < self._last[i]:
                self._active[i] += 1
                return True
            self._active[i] = self._first[i]
        return False

    def Reset(self):
        self._active[:] = self._first

    def ResetToMaximum(self):
        self._active[:] = self._last

    def Randomize(self, generator=None):
        """StateVector_Randomize (StateVector.c:163-170) with a NumPy generator."""
        generator = generator if generator is not None else np.random.default_rng()
        span = self._last - self._first + 1
        self._active[:] = self._first + generator.integers(0, span)

    def DefineSubstate(self, selectedSites):
        """|selectedSites| is a sequence of (segmentName, residueSerial) (StateVector.pyx:163-191)."""
        if self._substateSites is not None:
            raise CLibraryError("Memory allocation failure.")        # StateVector_AllocateSubstate refuses a second call
        sites = self.owner.sites
        chosen = []
        for selectedSegment, selectedSerial in selectedSites:
            for siteIndex, site in enumerate(sites):
                if site.segName == selectedSegment and site.resSerial == selectedSerial:
                    chosen.append(siteIndex)
                    self._isSubstate[siteIndex] = True
                    break
            else:
                raise CLibraryError("Site %s %d not found." % (selectedSegment, selectedSerial))
        self._substateSites = chosen

    def IncrementSubstate(self):
        """StateVector_IncrementSubstate, StateVector.c:389-404."""
        for i in (self._substateSites or []):
            if self._active[i] < self._last[i]:
                self._active[i] += 1
                return True
            self._active[i] = self._first[i]
        return False

    def ResetSubstate(self):
        for i in (self._substateSites or []):
            self._active[i] = self._first[i]

    def IsSubstate(self, index):
        self._checkIndex(index)
        return bool(self._isSubstate[index])

    def GlobalIndices(self):
        """The state as an int32 array of global instance indices (what the C ABI takes)."""
        return self._active.copy()

    def Print(self, verbose=True, title=None, log=logFile):
        """StateVector.pyx:131-159."""
        if LogFileActive(log):
            rows = []
            for indexSite, site in enumerate(self.owner.sites):
                local = self[indexSite]
                instance = site.instances[local]
                rows.append((site.segName, site.resName, "%d" % site.resSerial, instance.label, "%d" % instance.instIndex,
                             "@" if self._isSubstate[indexSite] else " "))
            log.Table(rows, title=title if title else "State vector")
