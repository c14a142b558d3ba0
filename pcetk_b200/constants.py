"""Physical constants and limits of the microstate path.

Values follow the reference bit for bit: ``EnergyModel.h:31-32`` and
``ContinuumElectrostatics/Constants.py:14-15`` (the truncated ln10 is intentional:
parity with the reference's energies depends on it).
"""

CONSTANT_MOLAR_GAS_KCAL_MOL = 0.001987165392
CONSTANT_LN10 = 2.302585092994

# Reference limit for analytic enumeration (EnergyModel.pyx:11, Constants.py:24-25).
# The reference needs it because it materialises an nstates-long array of Boltzmann
# factors and counts states in a 32-bit Integer; this engine streams, so the limit
# here is only the default guard and can be raised per call (see EnergyModel).
ANALYTIC_SITES = 26
ANALYTIC_STATES = 2 ** ANALYTIC_SITES

# Hard limit of this engine's enumeration kernel (64-bit state index).
ANALYTIC_STATES_B200 = 2 ** 44

DEFAULT_TEMPERATURE = 300.0
DEFAULT_DOUBLE_FLIP = 2.0
DEFAULT_PRODUCTION_SCANS = 20000
DEFAULT_EQUILIBRATION_SCANS = 500
