"""pcetk_b200 -- B200-native engine for the microstate hot path of efliks/pcetk.

Same names as ``ContinuumElectrostatics/__init__.py`` exports for that path.  Importing the package does not
need a GPU; constructing an ``EnergyModel`` needs ``libpcetk_b200.so`` and computing needs a CUDA device --
there is no CPU fallback.
"""

from .errors import ContinuumElectrostaticsError, CLibraryError, NativeLibraryMissing
from .constants import ANALYTIC_STATES, CONSTANT_LN10, CONSTANT_MOLAR_GAS_KCAL_MOL
from .formats import ModelRecord
from .StateVector import StateVector
from .EnergyModel import EnergyModel
from .MCModelDefault import MCModelDefault
from .CEModel import CEModel, MEADModel, Site, Instance
from .TitrationCurves import TitrationCurves
from .Substate import Substate, MEADSubstate, StateVector_FromProbabilities

__all__ = [
    "ContinuumElectrostaticsError", "CLibraryError", "NativeLibraryMissing", "ANALYTIC_STATES", "CONSTANT_LN10",
    "CONSTANT_MOLAR_GAS_KCAL_MOL", "ModelRecord", "StateVector", "EnergyModel", "MCModelDefault", "CEModel", "MEADModel",
    "Site", "Instance", "TitrationCurves", "Substate", "MEADSubstate", "StateVector_FromProbabilities",
]
