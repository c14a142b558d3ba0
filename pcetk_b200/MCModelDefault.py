"""MCModelDefault: Metropolis Monte Carlo sampling of protonation states on the GPU.

Mirrors ``extensions/cython/ContinuumElectrostatics.MCModelDefault.pyx:15-194`` (constructor
arguments, ``Initialize``, ``CalculateOwnerProbabilities``, ``Summary``, ``PrintPairs``).  The sampler runs
``nchains`` independent chains per pH value in one kernel launch (the reference runs one chain); all pH
values of a titration curve can go into the same launch (``CalculateOwnerProbabilitiesBatch``).
"""

import ctypes as C

import numpy as np

from . import _native
from .constants import DEFAULT_DOUBLE_FLIP, DEFAULT_EQUILIBRATION_SCANS, DEFAULT_PRODUCTION_SCANS
from .errors import CLibraryError
from .log import logFile, LogFileActive

_DefaultDoubleFlip = DEFAULT_DOUBLE_FLIP
_DefaultProductionScans = DEFAULT_PRODUCTION_SCANS
_DefaultEquilibrationScans = DEFAULT_EQUILIBRATION_SCANS
_DefaultChains = 64


class MCModelDefault(object):
    """A class defining the default Monte Carlo model."""

    def __init__(self, doubleFlip=_DefaultDoubleFlip, nequil=_DefaultEquilibrationScans, nprod=_DefaultProductionScans,
                 randomSeed=-1, nchains=_DefaultChains, kernel=0):
        if nequil < 0 or nprod < 1 or nchains < 1:
            raise CLibraryError("Cannot allocate Monte Carlo model.")
        self.doubleFlip = float(doubleFlip)
        self.nequil = int(nequil)
        self.nprod = int(nprod)
        self.randomSeed = int(randomSeed)
        self.nchains = int(nchains)
        self.kernel = int(kernel)
        self.chainOffset = 0          # first chain of this process in a multi-GPU ensemble
        self.isOwner = True
        self.owner = None
        self._handle = None
        self._lib = _native.library()
        self.lastCounters = None
        self.lastEnergySums = None
        self.lastLogRows = None
        # Random streams are addressed by (seed, pH index, chain, trial).  The reference's Mersenne Twister keeps advancing from
        # call to call; here every call of the reference-compatible entry points takes the next unused pH indices, so that
        # repeated calls (and the pH points of a serial titration loop) draw independent samples.  ``lastPhIndexOffset`` is the
        # first index the last call used: pass it as ``phIndexOffset`` to replay that call exactly.
        self._nextStream = 0
        self.lastPhIndexOffset = 0

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle:
            self._lib.pce_mc_destroy(handle)
            self._handle = None

    def Initialize(self, ceModel):
        """Link Monte Carlo and energy models and find pairs for double moves (MCModelDefault.pyx:37-54)."""
        if self._handle:
            self._lib.pce_mc_destroy(self._handle)
            self._handle = None
        handle = C.c_void_p()
        try:
            _native.check(self._lib.pce_mc_create(ceModel.energyModel._handle, self.doubleFlip, self.nequil, self.nprod,
                                                  self.randomSeed, self.nchains, C.byref(handle)))
        except CLibraryError as error:
            if getattr(error, "status", 0) == _native.PCE_ERR_STATE:
                raise
            raise CLibraryError("Cannot initialize Monte Carlo model.")
        self._handle = handle
        _native.check(self._lib.pce_mc_set_kernel(handle, self.kernel))
        seed = C.c_longlong(0)
        _native.check(self._lib.pce_mc_get_info(handle, None, None, None, None, C.byref(seed)))
        self.seed = seed.value
        self.isOwner = False
        self.owner = ceModel

    # ---- pairs ------------------------------------------------------------------------------------------
    @property
    def npairs(self):
        return self._lib.pce_mc_npairs(self._handle)

    def GetPairs(self):
        """-> list of (siteIndexA, siteIndexB, Wmax) in the reference's order (MCModelDefault.c:256-288)."""
        pairs = []
        for index in range(self.npairs):
            a, b, w = C.c_int(0), C.c_int(0), C.c_double(0.0)
            _native.check(self._lib.pce_mc_get_pair(self._handle, index, C.byref(a), C.byref(b), C.byref(w)))
            pairs.append((a.value, b.value, w.value))
        return pairs

    def LaunchPlan(self, nph=1):
        """Launch geometry for ``nph`` pH values: dict(kernel, block, ctas_per_sm, chains_per_wave, smem_bytes)."""
        self._requireOwner()
        self._configure()
        k, b, c, w, s = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_longlong(0)
        _native.check(self._lib.pce_mc_plan(self._handle, int(nph), C.byref(k), C.byref(b), C.byref(c), C.byref(w), C.byref(s)))
        return dict(kernel=k.value, block=b.value, ctas_per_sm=c.value, chains_per_wave=w.value, smem_bytes=s.value)

    # ---- sampling -----------------------------------------------------------------------------------------
    def _requireOwner(self):
        if self.owner is None or not self._handle:
            raise CLibraryError("Cannot initialize Monte Carlo model.")
        if not self.owner.isCalculated:
            raise CLibraryError("First calculate electrostatic energies.")

    def _configure(self):
        _native.check(self._lib.pce_mc_set_scans(self._handle, self.nequil, self.nprod))
        _native.check(self._lib.pce_mc_set_chains(self._handle, self.nchains))
        _native.check(self._lib.pce_mc_set_kernel(self._handle, self.kernel))

    def RunCounts(self, pHs, phIndexOffset=0, energy=True):
        """Raw results of one launch over a pH list: (counts[npH, ninst] int64, counters[npH, 4], esum[npH]).
        ``energy=False`` skips the tracked chain energies (``esum`` comes back as zeros): the reference's quiet path returns
        occupancies only, and the chain-per-thread kernel then fits more chains per SM."""
        self._requireOwner()
        self._configure()
        ph = np.ascontiguousarray(np.atleast_1d(pHs), dtype=np.float64)
        ninst = self.owner.ninstances
        counts = np.zeros((ph.shape[0], ninst), dtype=np.int64)
        counters = np.zeros((ph.shape[0], 4), dtype=np.int64)
        esum = np.zeros(ph.shape[0], dtype=np.float64)
        _native.check(self._lib.pce_mc_run(self._handle, ph, ph.shape[0], int(phIndexOffset), int(self.chainOffset),
                                           _native.ptr(counts), _native.ptr(counters), _native.ptr(esum) if energy else None))
        self.lastCounters = counters
        self.lastEnergySums = esum if energy else None
        return counts, counters, esum

    def _takeStreams(self, count, phIndexOffset=None):
        """First pH index of a call over ``count`` pH values: the next unused ones, or the explicit (replayable) offset."""
        if phIndexOffset is None:
            phIndexOffset = self._nextStream
            self._nextStream += int(count)
        self.lastPhIndexOffset = int(phIndexOffset)
        return int(phIndexOffset)

    def CalculateOwnerProbabilitiesBatch(self, pHs, phIndexOffset=None):
        """Probabilities for a list of pH values in one launch -> p[npH, ninst]; the model keeps the last row.
        Successive calls draw fresh random streams unless ``phIndexOffset`` addresses them explicitly."""
        count = len(np.atleast_1d(pHs))
        counts, _counters, _esum = self.RunCounts(pHs, self._takeStreams(count, phIndexOffset), energy=False)
        return counts / float(self.nchains * self.nprod)

    def CalculateOwnerProbabilities(self, pH=7.0, logFrequency=0, trajectoryFilename="", log=logFile):
        """Calculate probabilities of the owner (MCModelDefault.pyx:56-159).

        Quiet path: all chains in one launch.  With ``logFrequency`` > 0 and/or ``trajectoryFilename`` one chain is run, like
        the reference: the table of moves / accepted moves / double moves / accepted double moves per ``logFrequency``
        scans is printed (and kept in ``lastLogRows``), and the energy after every production scan is written to the
        file, one ``%f`` per line (:127-135)."""
        self._requireOwner()
        self._configure()
        if trajectoryFilename != "" or logFrequency > 0:
            # logging / trajectory branch (MCModelDefault.pyx:79-159): one chain, like the reference
            energies = np.zeros(self.nprod, dtype=np.float64)
            counts = np.zeros(self.owner.ninstances, dtype=np.int64)
            counters = np.zeros(4, dtype=np.int64)
            nrows = (self.nequil // logFrequency + self.nprod // logFrequency) if logFrequency > 0 else 0
            rows = np.zeros((max(nrows, 1), 5), dtype=np.int64)
            _native.check(self._lib.pce_mc_run_logged(self._handle, float(pH), self._takeStreams(1), int(max(logFrequency, 0)), _native.ptr(energies),
                                                      _native.ptr(counts), _native.ptr(counters), _native.ptr(rows) if nrows else None))
            if trajectoryFilename != "":
                with open(trajectoryFilename, "w") as output:
                    for value in energies:
                        output.write("%f\n" % value)
            self.lastCounters = counters[None, :]
            self.lastLogRows = rows[:nrows]
            if LogFileActive(log) and nrows:
                nEquilRows = self.nequil // logFrequency
                table = [("%8s" % "Scan", "%10s" % "Moves", "%10s" % "M-Acc", "%10s" % "Flips", "%10s" % "F-Acc")]
                for k, row in enumerate(rows[:nrows]):
                    if k == nEquilRows:
                        table.append(("** Equilibration phase done **", "", "", "", ""))
                    table.append(tuple(("%8d" if c == 0 else "%10d") % x for c, x in enumerate(row)))
                table.append(("** Production phase done **", "", "", "", ""))
                log.Table(table)
        else:
            self.CalculateOwnerProbabilitiesBatch([float(pH)])
            if LogFileActive(log):
                log.Text("\nCompleted %d equilibration scans.\n" % self.nequil)
                log.Text("\nCompleted %d production scans.\n" % self.nprod)

    def RunChains(self, pHs, phIndexOffset=0):
        """Per-chain outputs (parity tests): dict of counts[npH, nchains, ninst], state, energy, esum, counters."""
        self._requireOwner()
        self._configure()
        ph = np.ascontiguousarray(np.atleast_1d(pHs), dtype=np.float64)
        nph, nc, ninst, nsites = ph.shape[0], self.nchains, self.owner.ninstances, self.owner.nsites
        out = dict(counts=np.zeros((nph, nc, ninst), dtype=np.int64), state=np.zeros((nph, nc, nsites), dtype=np.int32),
                   energy=np.zeros((nph, nc), dtype=np.float64), esum=np.zeros((nph, nc), dtype=np.float64),
                   counters=np.zeros((nph, nc, 4), dtype=np.int64))
        _native.check(self._lib.pce_mc_run_chains(self._handle, ph, nph, int(phIndexOffset), int(self.chainOffset), _native.ptr(out["counts"]),
                                                  _native.ptr(out["state"]), _native.ptr(out["energy"]), _native.ptr(out["esum"]),
                                                  _native.ptr(out["counters"])))
        return out

    # ---- reporting ------------------------------------------------------------------------------------------
    def Summary(self, log=logFile):
        if LogFileActive(log):
            log.Table([("Equilibration scans", "%d" % self.nequil), ("Production scans", "%d" % self.nprod),
                       ("Limit for double moves", "%.1f" % self.doubleFlip), ("Chains per pH value", "%d" % self.nchains)],
                      title="Default Monte Carlo sampling model")

    def PrintPairs(self, log=logFile):
        """Summary of strongly interacting pairs (MCModelDefault.pyx:176-194)."""
        if LogFileActive(log):
            pairs = self.GetPairs()
            log.Text("\nFound %d pair%s of strongly interacting sites:\n" % (len(pairs), "s" if len(pairs) != 1 else ""))
            sites = self.owner.sites
            for indexSiteA, indexSiteB, Wmax in pairs:
                first, second = sites[indexSiteA], sites[indexSiteB]
                log.Text("%4s %4s %4d -- %4s %4s %4d : %f\n" % (first.segName, first.resName, first.resSerial,
                                                               second.segName, second.resName, second.resSerial, Wmax))
