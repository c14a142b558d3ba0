"""ctypes binding of ``libpcetk_b200.so`` (the C ABI declared in ``include/pcetk_b200.h``).

This is the only place the Python host layer touches native code.  There is no fallback: if the
shared library is missing, importing any compute class raises :class:`NativeLibraryMissing`; if no CUDA
device is present, every compute call raises :class:`CLibraryError` with the CUDA error text.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .errors import CLibraryError, NativeLibraryMissing

HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.environ.get("PCETK_B200_LIBRARY") or os.path.join(HERE, "libpcetk_b200.so")      # override: tuning builds

PCE_OK = 0
PCE_ERR_ALLOC, PCE_ERR_INDEX, PCE_ERR_VALUE, PCE_ERR_SHAPE, PCE_ERR_STATE, PCE_ERR_CUDA, PCE_ERR_LIMIT = range(1, 8)

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_void = C.c_void_p
_int, _dbl, _ll = C.c_int, C.c_double, C.c_longlong
_pint, _pdbl, _pll = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_longlong)

# name -> (restype, argtypes); every symbol include/pcetk_b200.h declares appears here
SIGNATURES = {
    "pce_last_error": (C.c_char_p, []),
    "pce_version": (C.c_char_p, []),
    "pce_device_count": (_int, [_pint]),
    "pce_device_name": (_int, [_int, C.c_char_p, _int]),
    "pce_set_stream": (_int, [_void, _void]),
    "pce_synchronize": (_int, [_void]),
    "pce_launch_count": (_ll, [_int]),
    "pce_probe_read_bandwidth": (_int, [_int, _ll, _int, C.POINTER(C.c_double)]),
    "pce_probe_on_chip": (_int, [_int, _int, C.POINTER(C.c_double)]),
    "pce_model_create": (_int, [_int, _int, _dbl, _int, C.POINTER(_void)]),
    "pce_model_destroy": (None, [_void]),
    "pce_model_nsites": (_int, [_void]),
    "pce_model_ninstances": (_int, [_void]),
    "pce_model_temperature": (_dbl, [_void]),
    "pce_model_set_site": (_int, [_void, _int, _int, _int]),
    "pce_model_set_sites": (_int, [_void, _i32p, _i32p]),
    "pce_model_get_site": (_int, [_void, _int, _pint, _pint]),
    "pce_model_nstates": (_int, [_void, _pdbl]),
    "pce_model_set_gmodel": (_int, [_void, _int, _dbl]),
    "pce_model_set_gintr": (_int, [_void, _int, _dbl]),
    "pce_model_set_protons": (_int, [_void, _int, _int]),
    "pce_model_set_probability": (_int, [_void, _int, _dbl]),
    "pce_model_set_interaction": (_int, [_void, _int, _int, _dbl]),
    "pce_model_get_gmodel": (_int, [_void, _int, _pdbl]),
    "pce_model_get_gintr": (_int, [_void, _int, _pdbl]),
    "pce_model_get_protons": (_int, [_void, _int, _pint]),
    "pce_model_get_probability": (_int, [_void, _int, _pdbl]),
    "pce_model_get_interaction": (_int, [_void, _int, _int, _pdbl]),
    "pce_model_get_interaction_symmetric": (_int, [_void, _int, _int, _pdbl]),
    "pce_model_get_deviation": (_int, [_void, _int, _int, _pdbl]),
    "pce_model_load": (_int, [_void, _void, _void, _void, _void]),
    "pce_model_get_probabilities": (_int, [_void, _f64p]),
    "pce_model_set_probabilities": (_int, [_void, _f64p]),
    "pce_model_get_symmetric_matrix": (_int, [_void, _f64p]),
    "pce_model_check_symmetric": (_int, [_void, _dbl, _pdbl, _pint]),
    "pce_model_symmetrize": (_int, [_void]),
    "pce_model_reset_interactions": (_int, [_void]),
    "pce_model_scale_interactions": (_int, [_void, _dbl]),
    "pce_model_state_from_probabilities": (_int, [_void, _i32p]),
    "pce_model_upload": (_int, [_void]),
    "pce_state_create": (_int, [_int, C.POINTER(_void)]),
    "pce_state_create_for_model": (_int, [_void, C.POINTER(_void)]),
    "pce_state_destroy": (None, [_void]),
    "pce_state_clone": (_int, [_void, C.POINTER(_void)]),
    "pce_state_copy_to": (_int, [_void, _void]),
    "pce_state_nsites": (_int, [_void]),
    "pce_state_nsubstate": (_int, [_void]),
    "pce_state_reset": (_int, [_void]),
    "pce_state_reset_substate": (_int, [_void]),
    "pce_state_reset_to_maximum": (_int, [_void]),
    "pce_state_randomize": (_int, [_void, C.c_ulonglong]),
    "pce_state_set_site": (_int, [_void, _int, _int, _int]),
    "pce_state_get_site": (_int, [_void, _int, _pint, _pint]),
    "pce_state_is_substate": (_int, [_void, _int, _pint]),
    "pce_state_get_item": (_int, [_void, _int, _pint]),
    "pce_state_set_item": (_int, [_void, _int, _int]),
    "pce_state_get_actual_item": (_int, [_void, _int, _pint]),
    "pce_state_set_actual_item": (_int, [_void, _int, _int]),
    "pce_state_allocate_substate": (_int, [_void, _int]),
    "pce_state_get_substate_item": (_int, [_void, _int, _pint]),
    "pce_state_set_substate_item": (_int, [_void, _int, _int]),
    "pce_state_increment": (_int, [_void, _pint]),
    "pce_state_increment_substate": (_int, [_void, _pint]),
    "pce_state_active": (C.POINTER(C.c_int32), [_void]),
    "pce_state_get_active": (_int, [_void, _i32p]),
    "pce_state_set_active": (_int, [_void, _i32p]),
    "pce_state_set_index": (_int, [_void, _ll]),
    "pce_state_get_index": (_int, [_void, _pll]),
    "pce_state_enumerate_substate": (_int, [_void, _ll, _void, _pll]),
    "pce_energy": (_int, [_void, _i32p, _ll, _f64p, _int, _int, _f64p]),
    "pce_energy_device": (_int, [_void, _void, _ll, _void, _int, _int, _void]),
    "pce_analytic_bins_size": (_ll, [_void]),
    "pce_analytic_bins": (_int, [_void, _f64p, _int, _int, _int, _int, _dbl, _void]),
    "pce_analytic_finish": (_int, [_void, _f64p, _f64p, _int, _int, _dbl, _void, _void]),
    "pce_analytic_bin_scales": (_int, [_void, _f64p, _int, _int, _f64p]),
    "pce_analytic": (_int, [_void, _f64p, _int, _int, _dbl, _dbl, _void, _void]),
    "pce_analytic_gmin": (_int, [_void, _dbl, _int, _dbl, _pdbl]),
    "pce_mc_create": (_int, [_void, _dbl, _int, _int, _ll, _int, C.POINTER(_void)]),
    "pce_mc_destroy": (None, [_void]),
    "pce_mc_npairs": (_int, [_void]),
    "pce_mc_get_pair": (_int, [_void, _int, _pint, _pint, _pdbl]),
    "pce_mc_set_scans": (_int, [_void, _int, _int]),
    "pce_mc_set_chains": (_int, [_void, _int]),
    "pce_mc_set_kernel": (_int, [_void, _int]),
    "pce_mc_get_info": (_int, [_void, _pint, _pint, _pint, _pint, _pll]),
    "pce_mc_plan": (_int, [_void, _int, _pint, _pint, _pint, _pint, _pll]),
    "pce_mc_run": (_int, [_void, _f64p, _int, _int, _ll, _void, _void, _void]),
    "pce_mc_run_device": (_int, [_void, _f64p, _int, _int, _ll, _void, _void, _void]),
    "pce_mc_run_chains": (_int, [_void, _f64p, _int, _int, _ll, _void, _void, _void, _void, _void]),
    "pce_mc_run_trajectory": (_int, [_void, _dbl, _int, _void, _void, _void]),
    "pce_mc_run_logged": (_int, [_void, _dbl, _int, _int, _void, _void, _void, _void]),
    "pce_test_philox": (_int, [_int, _u32p, _int, _u32p, _u32p]),
}

_lib = None


def library():
    """Load the shared library once; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBRARY_PATH):
            raise NativeLibraryMissing(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  pcetk_b200 has no CPU fallback." % LIBRARY_PATH)
        lib = C.CDLL(LIBRARY_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)        # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    return library().pce_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Raise ``CLibraryError`` for a non-zero status, like the Cython layer does for ``Status``."""
    if status != PCE_OK:
        error = CLibraryError(last_error())
        error.status = status
        raise error


def ptr(array):
    """Data pointer of a numpy array (or None) as c_void_p."""
    return None if array is None else C.c_void_p(array.ctypes.data)


def device_count() -> int:
    n = C.c_int(0)
    status = library().pce_device_count(C.byref(n))
    return n.value if status == PCE_OK else 0


def device_name(device: int = 0) -> str:
    buffer = C.create_string_buffer(256)
    check(library().pce_device_name(device, buffer, 256))
    return buffer.value.decode()


def launch_count(reset: bool = False) -> int:
    return int(library().pce_launch_count(1 if reset else 0))
