"""Synthetic models of the shapes BASELINE.json names (generator specs frozen from SURVEY.md section 8d).

``synth_a(nsites)``  : nsites two-instance sites (32 -> 2^32 microstates) for the analytic enumeration benchmark.
``synth_m(nsites)``  : 1000 sites / 3000 instances (pattern [2,2,3,5]) for the Monte Carlo benchmark.
Both use a screened-Coulomb W between point sites; same-site entries are zero, W is exactly symmetric.
"""

import numpy as np

from .formats import ModelRecord

PK_UNIT = 1.3727   # kcal/mol per pK unit at 300 K (R T ln10 = 1.3727)


def _screened_coulomb(positions, charges_a, charges_b):
    d = np.linalg.norm(positions[:, None, :] - positions[None, :, :], axis=-1)
    d = np.maximum(d, 3.0)
    return charges_a * charges_b * 332.06 / (20.0 * d) * np.exp(-d / 9.74)


def synth_a(nsites=32, seed=7, decouple_from=None):
    """nsites x 2 instances.  ``decouple_from=k`` zeroes every interaction of sites >= k (their answer factorises)."""
    rng = np.random.default_rng(seed)
    nmax = max(nsites, 32)
    positions = rng.normal(0.0, 10.0, size=(nmax, 3))
    pk = rng.uniform(2.0, 12.0, size=nmax)
    flag = rng.integers(0, 2, size=nmax)          # 1 = base (protonated form charged +1), 0 = acid (deprotonated form -1)
    positions, pk, flag = positions[:nsites], pk[:nsites], flag[:nsites]
    ninst = 2 * nsites
    gintr = np.zeros(ninst)
    protons = np.zeros(ninst, dtype=np.int32)
    charge = np.zeros(ninst)
    site_of = np.repeat(np.arange(nsites), 2)
    gintr[0::2] = -PK_UNIT * pk
    protons[0::2] = 1
    charge[0::2] = flag
    charge[1::2] = -1.0 + flag
    d = np.linalg.norm(positions[site_of][:, None, :] - positions[site_of][None, :, :], axis=-1)
    d = np.maximum(d, 3.0)
    w = charge[:, None] * charge[None, :] * 332.06 / (20.0 * d) * np.exp(-d / 9.74)
    w[site_of[:, None] == site_of[None, :]] = 0.0
    if decouple_from is not None:
        cut = site_of >= decouple_from
        w[cut, :] = 0.0
        w[:, cut] = 0.0
    first = np.arange(0, ninst, 2, dtype=np.int32)
    return ModelRecord(name="synthA%d" % nsites, site_first=first, site_last=first + 1, protons=protons, gmodel=gintr.copy(),
                       gintr=gintr, interactions=w, temperature=300.0,
                       site_labels=[("SYNA", "SIT", k + 1) for k in range(nsites)], inst_labels=["p", "d"] * nsites)


def synth_m(nsites=1000, seed=2024):
    """Instance counts from the shuffled pattern [2,2,3,5]; protons n-1..0 within a site."""
    rng = np.random.default_rng(seed)
    pattern = np.array([2, 2, 3, 5] * ((nsites + 3) // 4))[:nsites]
    rng.shuffle(pattern)
    radius = 40.9 * (nsites / 1000.0) ** (1.0 / 3.0)
    direction = rng.normal(size=(nsites, 3))
    direction /= np.linalg.norm(direction, axis=1)[:, None]
    positions = direction * (radius * rng.uniform(size=nsites) ** (1.0 / 3.0))[:, None]
    pk = rng.uniform(2.0, 12.0, size=nsites)
    flag = rng.integers(0, 2, size=nsites)
    first = np.concatenate([[0], np.cumsum(pattern)[:-1]]).astype(np.int32)
    last = (first + pattern - 1).astype(np.int32)
    ninst = int(pattern.sum())
    site_of = np.repeat(np.arange(nsites), pattern)
    local = np.arange(ninst) - first[site_of]
    n_of = pattern[site_of]
    protons = (n_of - 1 - local).astype(np.int32)
    noise = rng.normal(size=ninst) * (local > 0)
    gintr = -protons * PK_UNIT * pk[site_of] + noise
    # charge of an instance: fully deprotonated form is -1 (acid) or 0 (base); each proton adds +1
    charge = (-1.0 + flag[site_of]) + protons
    pos = positions[site_of]
    w = np.zeros((ninst, ninst))
    block = 500
    for start in range(0, ninst, block):
        d = np.linalg.norm(pos[start:start + block, None, :] - pos[None, :, :], axis=-1)
        d = np.maximum(d, 3.0)
        w[start:start + block] = charge[start:start + block, None] * charge[None, :] * 332.06 / (20.0 * d) * np.exp(-d / 9.74)
    w[site_of[:, None] == site_of[None, :]] = 0.0
    w = 0.5 * (w + w.T)
    labels = []
    for n in pattern:
        labels += ["i%d" % k for k in range(n)]
    return ModelRecord(name="synthM%d" % nsites, site_first=first, site_last=last, protons=protons, gmodel=gintr.copy(), gintr=gintr,
                       interactions=w, temperature=300.0, site_labels=[("SYNM", "SIT", k + 1) for k in range(nsites)], inst_labels=labels)
