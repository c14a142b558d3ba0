"""EnergyModel: Gintr, protons, interactions and probabilities of a model, plus the three compute entry
points of the hot path, on the GPU.

Mirrors ``extensions/cython/ContinuumElectrostatics.EnergyModel.pyx:15-217`` (same method names, argument
meaning and error messages) over the C ABI in ``include/pcetk_b200.h``.  Additions are batched forms
(``CalculateMicrostateEnergies``, ``CalculateProbabilitiesAnalyticallyBatch``) that the reference has no
equivalent for: it loops over states and pH values in Python.
"""

import ctypes as C

import numpy as np

from . import _native
from .constants import ANALYTIC_STATES
from .errors import CLibraryError
from .log import logFile, LogFileActive
from .StateVector import StateVector


class EnergyModel(object):
    """A class defining the energy model."""

    def __init__(self, ceModel, totalSites, totalInstances, device=0):
        lib = _native.library()
        handle = C.c_void_p()
        try:
            _native.check(lib.pce_model_create(int(totalSites), int(totalInstances), float(ceModel.temperature), int(device), C.byref(handle)))
        except CLibraryError:
            raise CLibraryError("Cannot allocate energy model.")
        self._lib = lib
        self._handle = handle
        self.device = device
        self.isOwner = False
        self.owner = ceModel
        self.nstates = 0
        self.analyticStatesLimit = ANALYTIC_STATES   # the reference's guard (EnergyModel.pyx:11); raise it to use the GPU's range

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle:
            self._lib.pce_model_destroy(handle)
            self._handle = None

    # ---- setters / getters (EnergyModel.pyx:27-68) ----------------------------------------------
    def SetGmodel(self, instIndexGlobal, value):
        _native.check(self._lib.pce_model_set_gmodel(self._handle, instIndexGlobal, value))

    def SetGintr(self, instIndexGlobal, value):
        _native.check(self._lib.pce_model_set_gintr(self._handle, instIndexGlobal, value))

    def SetProtons(self, instIndexGlobal, value):
        _native.check(self._lib.pce_model_set_protons(self._handle, instIndexGlobal, value))

    def SetProbability(self, instIndexGlobal, value):
        _native.check(self._lib.pce_model_set_probability(self._handle, instIndexGlobal, value))

    def SetInteraction(self, instIndexGlobalA, instIndexGlobalB, value):
        _native.check(self._lib.pce_model_set_interaction(self._handle, instIndexGlobalA, instIndexGlobalB, value))

    def _getDouble(self, function, *indices):
        value = C.c_double(0.0)
        _native.check(function(self._handle, *indices, C.byref(value)))
        return value.value

    def GetGmodel(self, instIndexGlobal):
        return self._getDouble(self._lib.pce_model_get_gmodel, instIndexGlobal)

    def GetGintr(self, instIndexGlobal):
        return self._getDouble(self._lib.pce_model_get_gintr, instIndexGlobal)

    def GetProtons(self, instIndexGlobal):
        value = C.c_int(0)
        _native.check(self._lib.pce_model_get_protons(self._handle, instIndexGlobal, C.byref(value)))
        return value.value

    def GetProbability(self, instIndexGlobal):
        return self._getDouble(self._lib.pce_model_get_probability, instIndexGlobal)

    def GetInteraction(self, instIndexGlobalA, instIndexGlobalB):
        return self._getDouble(self._lib.pce_model_get_interaction, instIndexGlobalA, instIndexGlobalB)

    def GetInteractionSymmetric(self, instIndexGlobalA, instIndexGlobalB):
        return self._getDouble(self._lib.pce_model_get_interaction_symmetric, instIndexGlobalA, instIndexGlobalB)

    def GetDeviation(self, instIndexGlobalA, instIndexGlobalB):
        return self._getDouble(self._lib.pce_model_get_deviation, instIndexGlobalA, instIndexGlobalB)

    # ---- bulk forms ---------------------------------------------------------------------------------
    def Load(self, gintr=None, gmodel=None, protons=None, interactions=None):
        """Set whole arrays at once (the reference sets them element by element from Python)."""
        def prepared(array, dtype):
            return None if array is None else np.ascontiguousarray(array, dtype=dtype)
        gintr, gmodel = prepared(gintr, np.float64), prepared(gmodel, np.float64)
        protons, interactions = prepared(protons, np.int32), prepared(interactions, np.float64)
        _native.check(self._lib.pce_model_load(self._handle, _native.ptr(gintr), _native.ptr(gmodel), _native.ptr(protons), _native.ptr(interactions)))

    def GetProbabilities(self):
        out = np.zeros(self._lib.pce_model_ninstances(self._handle), dtype=np.float64)
        _native.check(self._lib.pce_model_get_probabilities(self._handle, out))
        return out

    def SetProbabilities(self, probabilities):
        _native.check(self._lib.pce_model_set_probabilities(self._handle, np.ascontiguousarray(probabilities, dtype=np.float64)))

    def GetSymmetricMatrix(self):
        n = self._lib.pce_model_ninstances(self._handle)
        out = np.zeros((n, n), dtype=np.float64)
        _native.check(self._lib.pce_model_get_symmetric_matrix(self._handle, out))
        return out

    def SetStream(self, cuda_stream):
        """Launch on this ``cudaStream_t`` (an int, e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        _native.check(self._lib.pce_set_stream(self._handle, C.c_void_p(cuda_stream)))

    def Synchronize(self):
        _native.check(self._lib.pce_synchronize(self._handle))

    def Upload(self):
        _native.check(self._lib.pce_model_upload(self._handle))

    # ---- EnergyModel.pyx:86-136 -------------------------------------------------------------------
    def Initialize(self):
        """Set up sites and calculate the number of possible states."""
        ceModel = self.owner
        nstates = 1
        for indexSite, site in enumerate(ceModel.sites):
            indices = [instance._instIndexGlobal for instance in site.instances]
            _native.check(self._lib.pce_model_set_site(self._handle, indexSite, min(indices), max(indices)))
            if nstates <= ANALYTIC_STATES:          # the reference stops multiplying past the limit (:108-109)
                nstates = nstates * len(indices)
        self.nstates = nstates

    def NumberOfStates(self):
        """Exact number of microstates as a float (no cap)."""
        value = C.c_double(0.0)
        _native.check(self._lib.pce_model_nstates(self._handle, C.byref(value)))
        return value.value

    def CheckIfSymmetric(self, tolerance=0.05):
        """-> (isSymmetric, maxDeviation)"""
        deviation, flag = C.c_double(0.0), C.c_int(0)
        _native.check(self._lib.pce_model_check_symmetric(self._handle, tolerance, C.byref(deviation), C.byref(flag)))
        return (bool(flag.value), deviation.value)

    def SymmetrizeInteractions(self, log=logFile):
        _native.check(self._lib.pce_model_symmetrize(self._handle))
        if LogFileActive(log):
            log.Text("\nSymmetrizing interactions complete.\n")

    def ResetInteractions(self):
        _native.check(self._lib.pce_model_reset_interactions(self._handle))

    def ScaleInteractions(self, scale):
        _native.check(self._lib.pce_model_scale_interactions(self._handle, scale))

    # ---- kernel 1 -------------------------------------------------------------------------------------
    def _requireCalculated(self):
        if not self.owner.isCalculated:
            raise CLibraryError("First calculate electrostatic energies.")

    def CalculateMicrostateEnergies(self, states, pH=7.0, unfolded=False):
        """Batched kernel-1 call: ``states`` is (nstates, nsites) GLOBAL instance indices, ``pH`` a scalar
        or a sequence -> array (nstates, npH) (or (nstates,) for a scalar pH)."""
        self._requireCalculated()
        states = np.ascontiguousarray(states, dtype=np.int32)
        if states.ndim == 1:
            states = states[None, :]
        scalar = np.isscalar(pH)
        ph = np.ascontiguousarray(np.atleast_1d(pH), dtype=np.float64)
        out = np.zeros((states.shape[0], ph.shape[0]), dtype=np.float64)
        _native.check(self._lib.pce_energy(self._handle, states, states.shape[0], ph, ph.shape[0], int(bool(unfolded)), out))
        return out[:, 0] if scalar else out

    def CalculateMicrostateEnergy(self, vector, pH=7.0):
        """Calculate energy of a protonation state (=microstate) (EnergyModel.pyx:139-147)."""
        self._requireCalculated()
        return float(self.CalculateMicrostateEnergies(vector.GlobalIndices(), pH=float(pH))[0])

    def CalculateMicrostateEnergyUnfolded(self, vector, pH=7.0):
        self._requireCalculated()
        return float(self.CalculateMicrostateEnergies(vector.GlobalIndices(), pH=float(pH), unfolded=True)[0])

    # ---- kernel 2 -------------------------------------------------------------------------------------
    def _checkStates(self):
        limit = self.analyticStatesLimit
        if self.NumberOfStates() > limit:
            raise CLibraryError("Maximum number of states for analytic treatment (%d) exceeded." % limit)

    @staticmethod
    def _phGroups(ph, max_protons):
        """Split a pH list into groups one enumeration pass can serve (kernel 2).  Every proton-count bin carries
        its own scale, so only the protons of a partial sum (at most 16) see the half-width of a group: pH 0..20 is one
        group for any model; a grid wider than ~32 pH units is split."""
        order = np.argsort(ph, kind="stable")
        halfwidth = 590.0 / (2.302585092994 * max(min(max_protons, 16), 1))
        groups, current = [], []
        for index in order:
            if current and ph[index] - ph[current[0]] > 2.0 * halfwidth:
                groups.append(current)
                current = []
            current.append(int(index))
        if current:
            groups.append(current)
        return groups

    def CalculateProbabilitiesAnalyticallyBatch(self, pHs, unfolded=False, Gzero=0.0, want_lnz=False):
        """Analytic probabilities for a whole pH grid from one enumeration -> p[npH, ninst] (and lnZ[npH]).

        The last pH's probabilities are left in the model, like a loop over
        ``CalculateProbabilitiesAnalytically`` would."""
        self._checkStates()
        self._requireCalculated()
        ph = np.ascontiguousarray(np.atleast_1d(pHs), dtype=np.float64)
        ninst = self._lib.pce_model_ninstances(self._handle)
        out = np.zeros((ph.shape[0], ninst), dtype=np.float64)
        lnz = np.zeros(ph.shape[0], dtype=np.float64)
        max_protons = int(self._lib.pce_analytic_bins_size(self._handle) // (1 + ninst)) - 1
        groups = self._phGroups(ph, max_protons)
        last = int(ph.shape[0]) - 1
        groups.sort(key=lambda g: last in g)         # the group holding the last pH runs last (model probabilities)
        for group in groups:
            if last in group:
                group = [k for k in group if k != last] + [last]
            sub = np.ascontiguousarray(ph[group])
            p = np.zeros((len(group), ninst), dtype=np.float64)
            z = np.zeros(len(group), dtype=np.float64)
            _native.check(self._lib.pce_analytic(self._handle, sub, len(group), int(bool(unfolded)), float(Gzero),
                                                 float(self.analyticStatesLimit), _native.ptr(p), _native.ptr(z)))
            out[group] = p
            lnz[group] = z
        return (out, lnz) if want_lnz else out

    def CalculateProbabilitiesAnalytically(self, pH=7.0):
        """Calculate probabilities of protonation states analytically (EnergyModel.pyx:150-164)."""
        self._checkStates()
        self._requireCalculated()
        self.CalculateProbabilitiesAnalyticallyBatch([float(pH)])
        return int(self.NumberOfStates())

    def CalculateProbabilitiesAnalyticallyUnfolded(self, pH=7.0):
        self._checkStates()
        self._requireCalculated()
        self.CalculateProbabilitiesAnalyticallyBatch([float(pH)], unfolded=True)
        return int(self.NumberOfStates())

    def _z(self, Gzero, pH, unfolded):
        # The reference returns sum_s exp(-(G_s - Gmin)/RT), Gmin = min_s (G_s - Gzero): independent of
        # Gzero and relative to an unreturned Gmin (EnergyModel.c:252-281).  Reproduced for parity.
        self._checkStates()
        _p, lnz = self.CalculateProbabilitiesAnalyticallyBatch([float(pH)], unfolded=unfolded, Gzero=0.0, want_lnz=True)
        gmin = C.c_double(0.0)
        _native.check(self._lib.pce_analytic_gmin(self._handle, float(pH), int(bool(unfolded)), float(self.analyticStatesLimit), C.byref(gmin)))
        beta = 1.0 / (0.001987165392 * self._lib.pce_model_temperature(self._handle))
        return float(np.exp(lnz[0] + beta * gmin.value))

    def CalculateZfolded(self, Gzero=0.0, pH=7.0):
        """Calculate partition function of a folded protein (EnergyModel.pyx:194-204)."""
        return self._z(Gzero, pH, False)

    def CalculateZunfolded(self, Gzero=0.0, pH=7.0):
        """Calculate partition function of an unfolded protein (EnergyModel.pyx:207-217)."""
        return self._z(Gzero, pH, True)

    def CalculateLnZ(self, pHs, unfolded=False, Gzero=0.0):
        """ln sum_s exp(-(G_s - Gzero)/RT) per pH: the quantity folding free energies need
        (the reference's Z is only defined up to its hidden Gmin shift)."""
        _p, lnz = self.CalculateProbabilitiesAnalyticallyBatch(pHs, unfolded=unfolded, Gzero=Gzero, want_lnz=True)
        return lnz

    def StateVectorFromProbabilities(self):
        """Most probable instance of every site as a StateVector (EnergyModel.c:107-140, loop bound fixed)."""
        vector = StateVector(self.owner)
        state = np.zeros(len(vector), dtype=np.int32)
        _native.check(self._lib.pce_model_state_from_probabilities(self._handle, state))
        vector.SetGlobalIndices(state)
        return vector
