"""CPU emulator of the enumeration kernel's CTA programs (test infrastructure; see enum_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libenum_emu.so")
SOURCES = [os.path.join(HERE, "enum_emu.cpp"), os.path.join(ROOT, "pcetk_b200", "csrc", "pce_enum_core.cuh"),
           os.path.join(ROOT, "pcetk_b200", "csrc", "pce_enum_plan.h")]

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lib = None


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SOURCES):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", LIB, SOURCES[0]], check=True)


def library():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emu_analytic_bins.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, _f64p, C.c_double, C.c_int, _f64p, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, _f64p, _f64p, _i32p, C.c_char_p, C.c_int]
    return _lib


def analytic_bins(rec, wsym, ph, unfolded=False, shard=0, nshards=1, grid=3, grid_slots=0, max_mode=2):
    """-> (bins[(1 + ninst), Nb] scaled by exp(sigma[n]), sigma[Nb], info dict); info['modes'] = kernel variant of every run
    (0 generic, 1 canonical I and O1 sites, 2 with a J site)"""
    lib = library()
    ph = np.ascontiguousarray(np.atleast_1d(ph), dtype=np.float64)
    nb_max = int(np.sum([rec.protons[f:l + 1].max() for f, l in zip(rec.site_first, rec.site_last)])) + 1
    bins = np.zeros((1 + rec.ninstances, nb_max), dtype=np.float64)
    sigma = np.zeros(nb_max, dtype=np.float64)
    info = np.zeros(4, dtype=np.int32)
    error = C.create_string_buffer(256)
    g = rec.gmodel if unfolded else rec.gintr
    status = lib.emu_analytic_bins(rec.nsites, rec.ninstances, np.ascontiguousarray(rec.site_first, dtype=np.int32),
                                   np.ascontiguousarray(rec.site_last, dtype=np.int32), np.ascontiguousarray(rec.protons, dtype=np.int32),
                                   np.ascontiguousarray(g, dtype=np.float64), np.ascontiguousarray(wsym, dtype=np.float64), float(rec.temperature),
                                   int(bool(unfolded)), ph, len(ph), shard, nshards, grid, grid_slots, int(max_mode), bins, sigma, info, error, 256)
    if status != 0:
        raise RuntimeError("emulator: " + error.value.decode())
    assert info[1] == nb_max
    return bins, sigma, dict(runs=int(info[0]), modes=[(int(info[2]) >> (2 * r)) & 3 for r in range(int(info[0]))], P_O2=int(info[3]))


def finish(rec, bins, sigma, ph):
    """Probabilities p[npH, ninst] from scaled bins (extended precision; the library's pce_analytic_finish does the same)."""
    ph = np.atleast_1d(ph)
    beta = np.longdouble(1.0) / (np.longdouble(0.001987165392) * np.longdouble(rec.temperature))
    ph_star = 0.5 * (ph.min() + ph.max())
    mu = lambda x: -np.longdouble(0.001987165392) * np.longdouble(rec.temperature) * np.longdouble(2.302585092994) * np.longdouble(x)
    n = np.arange(bins.shape[1]).astype(np.longdouble)
    out = np.zeros((len(ph), rec.ninstances))
    for q, value in enumerate(ph):
        logw = -sigma.astype(np.longdouble) + beta * n * (mu(value) - mu(ph_star))
        z = bins[0].astype(np.longdouble)
        shift = np.max(np.where(z > 0, np.log(np.where(z > 0, z, 1)) + logw, -np.inf))
        w = np.exp(logw - shift)
        out[q] = ((bins[1:].astype(np.longdouble) * w).sum(axis=1) / (z * w).sum()).astype(np.float64)
    return out
