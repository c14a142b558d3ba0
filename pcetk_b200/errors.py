"""Exceptions mirroring the reference's two error types.

``ContinuumElectrostaticsError`` is ``ContinuumElectrostatics/Error.py:13``; ``CLibraryError``
stands in for ``pCore.CLibraryError``, which the Cython layer raises when the C ``Status``
comes back non-zero (e.g. ``ContinuumElectrostatics.EnergyModel.pyx:156-159``).
"""


class ContinuumElectrostaticsError(Exception):
    pass


class CLibraryError(Exception):
    pass


class NativeLibraryMissing(ImportError):
    """Raised when the CUDA shared library is absent.  There is no CPU fallback."""
