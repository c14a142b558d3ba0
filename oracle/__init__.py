"""CPU oracles for the microstate path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package.  ``pcetk_b200`` never does; the product has no CPU fallback.

Two oracles live here:

``Port``   ``oracle/pce_oracle.c`` -- an independent plain-C restatement of the reference algorithms
           (EnergyModel.c / StateVector.c / MCModelDefault.c, cited per function) plus the executable
           specification of the GPU Monte Carlo chain (Philox stream, cached fields).
``Ref``    ``oracle/_ref/libpcetk_ref.so`` -- the reference's own three C files compiled unmodified
           from ``/root/reference`` against a small pCore shim (``oracle/shim``), driven through
           ``oracle/ref_driver.c``.  Built in the build container; travels to the GPU box as a binary.

Pinning (tests/test_oracle.py): Port == golden vectors of the reference's fixtures (5e-7, the ``%f``
precision of the files) and, where ``_ref`` is present, Port == Ref bit for bit (energies, analytic
probabilities, pair lists, and the MT19937-driven MC at equal seeds).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "_build", "libpce_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libpcetk_ref.so")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """Compile the port always, and ``_ref`` when ``/root/reference`` is mounted."""
    proc = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout, proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _model_args(rec, wsym):
    return (
        C.c_int(rec.nsites), C.c_int(rec.ninstances), C.c_double(rec.temperature),
        np.ascontiguousarray(rec.site_first, dtype=np.int32), np.ascontiguousarray(rec.site_last, dtype=np.int32),
        np.ascontiguousarray(rec.protons, dtype=np.int32), np.ascontiguousarray(rec.gintr, dtype=np.float64),
        np.ascontiguousarray(rec.gmodel, dtype=np.float64), np.ascontiguousarray(wsym, dtype=np.float64),
    )


_MODEL_SIG = [C.c_int, C.c_int, C.c_double, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p]


class Port:
    """ctypes face of ``pce_oracle.c``."""

    def __init__(self):
        if not os.path.exists(PORT_LIB):
            build()
        lib = C.CDLL(PORT_LIB)
        lib.oracle_symmetrize.argtypes = [C.c_int, _f64p, _f64p]
        lib.oracle_check_symmetric.argtypes = [C.c_int, _f64p, C.c_double, C.POINTER(C.c_double)]
        lib.oracle_energy.restype = C.c_double
        lib.oracle_energy.argtypes = _MODEL_SIG + [_i32p, C.c_double, C.c_int]
        lib.oracle_analytic.argtypes = _MODEL_SIG + [C.c_double, C.c_double, C.c_int, C.c_longlong, C.c_void_p,
                                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.oracle_analytic_compensated.argtypes = _MODEL_SIG + [C.c_double, C.c_double, C.c_int, C.c_longlong, C.c_int, C.c_void_p,
                                                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.oracle_enumerate_energies.argtypes = _MODEL_SIG + [C.c_double, C.c_longlong, _f64p]
        lib.oracle_find_pairs.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_double, C.c_int, _i32p, _i32p, _f64p]
        lib.oracle_mc_mt.argtypes = _MODEL_SIG + [C.c_int, _i32p, _i32p, C.c_int, C.c_int, C.c_uint32, C.c_int, _f64p,
                                                  _f64p, C.c_void_p, C.c_void_p]
        lib.oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        lib.oracle_mc_philox_chain.argtypes = _MODEL_SIG + [C.c_int, _i32p, _i32p, C.c_int, C.c_int, C.c_uint64, C.c_uint32,
                                                            C.c_uint32, C.c_double, _i64p, _i64p, C.POINTER(C.c_double),
                                                            _i32p, C.POINTER(C.c_double)]
        self.lib = lib

    def symmetrize(self, wraw):
        wraw = np.ascontiguousarray(wraw, dtype=np.float64)
        out = np.empty_like(wraw)
        self.lib.oracle_symmetrize(wraw.shape[0], wraw, out)
        return out

    def check_symmetric(self, wraw, tolerance=0.05):
        wraw = np.ascontiguousarray(wraw, dtype=np.float64)
        dev = C.c_double(0.0)
        ok = self.lib.oracle_check_symmetric(wraw.shape[0], wraw, tolerance, C.byref(dev))
        return bool(ok), dev.value

    def energy(self, rec, wsym, state, pH=7.0, unfolded=False):
        state = np.ascontiguousarray(state, dtype=np.int32)
        return self.lib.oracle_energy(*_model_args(rec, wsym), state, pH, int(unfolded))

    def enumerate_energies(self, rec, wsym, pH, n):
        out = np.empty(n, dtype=np.float64)
        self.lib.oracle_enumerate_energies(*_model_args(rec, wsym), pH, n, out)
        return out

    def analytic(self, rec, wsym, pH=7.0, unfolded=False, gzero=0.0, want_p=True):
        """-> (p[ninst] or None, Z, Gmin) exactly as EnergyModel.c:252-354 computes them."""
        p = np.empty(rec.ninstances, dtype=np.float64) if want_p else None
        z, gmin = C.c_double(0.0), C.c_double(0.0)
        status = self.lib.oracle_analytic(*_model_args(rec, wsym), pH, gzero, int(unfolded), rec.nstates,
                                          p.ctypes.data if want_p else None, C.byref(z), C.byref(gmin))
        if status != 0:
            raise MemoryError("cannot allocate Boltzmann factors")
        return p, z.value, gmin.value

    def analytic_compensated(self, rec, wsym, pH=7.0, unfolded=False, gzero=0.0, threads=0):
        """-> (p[ninst], Z, Gmin): the reference's algorithm (EnergyModel.c:252-354) with long-double sums, state range split
        over ``threads`` host threads (default: all).  For state counts where the reference's serial double sum is the noise."""
        p = np.empty(rec.ninstances, dtype=np.float64)
        z, gmin = C.c_double(0.0), C.c_double(0.0)
        status = self.lib.oracle_analytic_compensated(*_model_args(rec, wsym), pH, gzero, int(unfolded), rec.nstates,
                                                      int(threads or os.cpu_count() or 1), p.ctypes.data, C.byref(z), C.byref(gmin))
        if status != 0:
            raise MemoryError("cannot allocate accumulators")
        return p, z.value, gmin.value

    def find_pairs(self, rec, wsym, limit=2.0):
        cap = rec.nsites * (rec.nsites - 1) // 2 + 1
        a = np.zeros(cap, dtype=np.int32)
        b = np.zeros(cap, dtype=np.int32)
        wmax = np.zeros(cap, dtype=np.float64)
        n = self.lib.oracle_find_pairs(rec.nsites, rec.ninstances, np.ascontiguousarray(rec.site_first, dtype=np.int32),
                                       np.ascontiguousarray(rec.site_last, dtype=np.int32),
                                       np.ascontiguousarray(wsym, dtype=np.float64), limit, cap, a, b, wmax)
        return a[:n].copy(), b[:n].copy(), wmax[:n].copy()

    def mc_mt(self, rec, wsym, pairs, nequil, nprod, seed, ph, want_energies=False):
        """Reference-semantics MC (MT19937) over a list of pH values on one continuing stream."""
        pa, pb = (np.ascontiguousarray(x, dtype=np.int32) for x in pairs)
        ph = np.ascontiguousarray(np.atleast_1d(ph), dtype=np.float64)
        p = np.zeros((len(ph), rec.ninstances), dtype=np.float64)
        counters = np.zeros((len(ph), 4), dtype=np.int64)
        energies = np.zeros((len(ph), nprod), dtype=np.float64) if want_energies else None
        if len(pa) == 0:
            pa = pb = np.zeros(1, dtype=np.int32)
            npairs = 0
        else:
            npairs = len(pa)
        self.lib.oracle_mc_mt(*_model_args(rec, wsym), npairs, pa, pb, nequil, nprod, seed, len(ph), ph, p,
                              energies.ctypes.data if want_energies else None, counters.ctypes.data)
        return p, counters, energies

    def philox(self, ctr, key):
        out = np.zeros(4, dtype=np.uint32)
        self.lib.oracle_philox4x32_10(np.ascontiguousarray(ctr, dtype=np.uint32), np.ascontiguousarray(key, dtype=np.uint32), out)
        return out

    def mc_philox_chain(self, rec, wsym, pairs, nequil, nprod, seed, chain, ph_index, pH):
        """One chain of the GPU sampler's specification -> dict(counts, counters, esum, state, energy)."""
        pa, pb = (np.ascontiguousarray(x, dtype=np.int32) for x in pairs)
        npairs = len(pa)
        if npairs == 0:
            pa = pb = np.zeros(1, dtype=np.int32)
        counts = np.zeros(rec.ninstances, dtype=np.int64)
        counters = np.zeros(4, dtype=np.int64)
        esum, energy = C.c_double(0.0), C.c_double(0.0)
        state = np.zeros(rec.nsites, dtype=np.int32)
        self.lib.oracle_mc_philox_chain(*_model_args(rec, wsym), npairs, pa, pb, nequil, nprod, seed, chain, ph_index, pH,
                                        counts, counters, C.byref(esum), state, C.byref(energy))
        return dict(counts=counts, counters=counters, esum=esum.value, state=state, energy=energy.value)


class Ref:
    """ctypes face of ``oracle/_ref/libpcetk_ref.so`` (the reference's own C, unmodified)."""

    def __init__(self, rec, nstates_cap=2 ** 26):
        if not have_ref():
            raise FileNotFoundError(REF_LIB)
        lib = C.CDLL(REF_LIB)
        lib.ref_model_create.restype = C.c_void_p
        lib.ref_model_create.argtypes = [C.c_int, C.c_int, C.c_double, _i32p, _i32p, _f64p, _f64p, _i32p, _f64p, C.c_longlong]
        lib.ref_model_destroy.argtypes = [C.c_void_p]
        lib.ref_check_symmetric.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        lib.ref_get_w.restype = C.c_double
        lib.ref_get_w.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ref_energy.restype = C.c_double
        lib.ref_energy.argtypes = [C.c_void_p, _i32p, C.c_double, C.c_int]
        lib.ref_enumerate_energies.argtypes = [C.c_void_p, C.c_double, C.c_longlong, _f64p]
        lib.ref_analytic.argtypes = [C.c_void_p, C.c_double, C.c_int, _f64p]
        lib.ref_z.restype = C.c_double
        lib.ref_z.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        lib.ref_mc_create.restype = C.c_void_p
        lib.ref_mc_create.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        lib.ref_mc_destroy.argtypes = [C.c_void_p]
        lib.ref_mc_npairs.argtypes = [C.c_void_p]
        lib.ref_mc_get_pair.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        lib.ref_mc_run.argtypes = [C.c_void_p, C.c_double, _f64p]
        lib.ref_mc_run_trajectory.argtypes = [C.c_void_p, C.c_double, _f64p, C.c_void_p, C.c_void_p]
        pint = C.POINTER(C.c_int)
        for name, argtypes in (("ref_vector_nsites", []), ("ref_vector_increment", []), ("ref_vector_increment_substate", []),
                               ("ref_vector_get_item", [C.c_int, pint]), ("ref_vector_set_item", [C.c_int, C.c_int]),
                               ("ref_vector_get_actual_item", [C.c_int, pint]), ("ref_vector_set_actual_item", [C.c_int, C.c_int]),
                               ("ref_vector_is_substate", [C.c_int, pint]), ("ref_vector_allocate_substate", [C.c_int]),
                               ("ref_vector_set_substate_item", [C.c_int, C.c_int]), ("ref_vector_get_substate_item", [C.c_int, pint])):
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = [C.c_void_p] + argtypes
        for name in ("ref_vector_reset", "ref_vector_reset_to_maximum", "ref_vector_reset_substate"):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.ref_vector_state.restype = None
        lib.ref_vector_state.argtypes = [C.c_void_p, _i32p]
        self.lib = lib
        self.rec = rec
        self.handle = lib.ref_model_create(
            rec.nsites, rec.ninstances, rec.temperature,
            np.ascontiguousarray(rec.site_first, dtype=np.int32), np.ascontiguousarray(rec.site_last, dtype=np.int32),
            np.ascontiguousarray(rec.gintr, dtype=np.float64), np.ascontiguousarray(rec.gmodel, dtype=np.float64),
            np.ascontiguousarray(rec.protons, dtype=np.int32), np.ascontiguousarray(rec.interactions, dtype=np.float64),
            nstates_cap)
        if not self.handle:
            raise MemoryError("ref_model_create failed")
        self._mc = []

    def close(self):
        for mc in self._mc:
            self.lib.ref_mc_destroy(mc)
        self._mc = []
        if self.handle:
            self.lib.ref_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's StateVector, function by function (StateVector.h:53-84) -> (status, value) ----
    STATUS_NAMES = ("Continue", "ArrayNonConformableSizes", "IndexOutOfRange", "MemoryAllocationFailure", "ValueError")

    def vector_call(self, name, *args):
        """Call ``ref_vector_<name>``; functions with an out-parameter return ``(status, value)``, the others their result."""
        fn = getattr(self.lib, "ref_vector_" + name)
        if name in ("get_item", "get_actual_item", "is_substate", "get_substate_item"):
            value = C.c_int(0)
            status = fn(self.handle, *args, C.byref(value))
            return status, value.value
        return fn(self.handle, *args)

    def vector_state(self):
        out = np.zeros(self.lib.ref_vector_nsites(self.handle), dtype=np.int32)
        self.lib.ref_vector_state(self.handle, out)
        return out

    def check_symmetric(self, tolerance=0.05):
        dev = C.c_double(0.0)
        ok = self.lib.ref_check_symmetric(self.handle, tolerance, C.byref(dev))
        return bool(ok), dev.value

    def wsym(self):
        n = self.rec.ninstances
        return np.array([[self.lib.ref_get_w(self.handle, a, b) for b in range(n)] for a in range(n)])

    def energy(self, state, pH=7.0, unfolded=False):
        return self.lib.ref_energy(self.handle, np.ascontiguousarray(state, dtype=np.int32), pH, int(unfolded))

    def enumerate_energies(self, pH, n):
        out = np.empty(n, dtype=np.float64)
        self.lib.ref_enumerate_energies(self.handle, pH, n, out)
        return out

    def analytic(self, pH=7.0, unfolded=False):
        p = np.empty(self.rec.ninstances, dtype=np.float64)
        status = self.lib.ref_analytic(self.handle, pH, int(unfolded), p)
        if status != 0:
            raise MemoryError("cannot allocate Boltzmann factors")
        return p

    def z(self, pH=7.0, gzero=0.0, unfolded=False):
        return self.lib.ref_z(self.handle, pH, gzero, int(unfolded))

    def mc_create(self, limit=2.0, nequil=500, nprod=20000, seed=1):
        mc = self.lib.ref_mc_create(self.handle, limit, nequil, nprod, seed)
        self._mc.append(mc)
        return mc

    def mc_pairs(self, mc):
        n = self.lib.ref_mc_npairs(mc)
        a, b, w = [], [], []
        for k in range(n):
            sa, sb, wm = C.c_int(0), C.c_int(0), C.c_double(0.0)
            self.lib.ref_mc_get_pair(mc, k, C.byref(sa), C.byref(sb), C.byref(wm))
            a.append(sa.value), b.append(sb.value), w.append(wm.value)
        return np.array(a, dtype=np.int32), np.array(b, dtype=np.int32), np.array(w)

    def mc_run(self, mc, pH=7.0):
        p = np.empty(self.rec.ninstances, dtype=np.float64)
        self.lib.ref_mc_run(mc, pH, p)
        return p

    def mc_run_trajectory(self, mc, pH, nprod):
        p = np.empty(self.rec.ninstances, dtype=np.float64)
        energies = np.empty(nprod, dtype=np.float64)
        counters = np.zeros(4, dtype=np.int64)
        self.lib.ref_mc_run_trajectory(mc, pH, p, energies.ctypes.data, counters.ctypes.data)
        return p, energies, counters
