"""NumPy restatement of the analytic path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, vectorised opinion with a *different* summation order from both the reference and the C
port (pairwise sums; energies accumulated site by site over all states at once).  Follows
EnergyModel.c:204-224 (energy), StateVector.c:370-384 (state order: site 0 fastest) and
EnergyModel.c:252-281, 318-337 (Boltzmann factors relative to the minimum, per-instance sums / Z).
Practical up to a few million states.
"""

import numpy as np

R = 0.001987165392
LN10 = 2.302585092994


def digits(rec):
    """Local instance index of every site for every state, shape (nsites, nstates), site 0 fastest."""
    radices = (rec.site_last - rec.site_first + 1).astype(np.int64)
    index = np.arange(rec.nstates, dtype=np.int64)
    out = np.empty((rec.nsites, rec.nstates), dtype=np.int32)
    for i, n in enumerate(radices):
        out[i] = index % n
        index //= n
    return out


def state_energy_parts(rec, wsym, unfolded=False):
    """pH-independent part A_s and proton count n_s of every state: G_s(pH) = A_s + n_s * R T ln10 pH."""
    d = digits(rec)
    g = d + rec.site_first[:, None].astype(np.int32)
    base = rec.gmodel if unfolded else rec.gintr
    a = np.zeros(rec.nstates, dtype=np.float64)
    nprot = np.zeros(rec.nstates, dtype=np.int64)
    for i in range(rec.nsites):
        a += base[g[i]]
        nprot += rec.protons[g[i]]
        if not unfolded:
            for j in range(i):
                a += wsym[g[i], g[j]]
    return a, nprot, g


def analytic(rec, wsym, ph_values, unfolded=False):
    """-> p[npH, ninst] for a list of pH values."""
    a, nprot, g = state_energy_parts(rec, wsym, unfolded)
    beta = 1.0 / (R * rec.temperature)
    out = np.zeros((len(ph_values), rec.ninstances), dtype=np.float64)
    for q, ph in enumerate(ph_values):
        energy = a + nprot * (R * rec.temperature * LN10 * ph)
        b = np.exp(-(energy - energy.min()) * beta)
        z = b.sum()
        for i in range(rec.nsites):
            out[q] += np.bincount(g[i], weights=b, minlength=rec.ninstances)
        out[q] /= z
    return out
