"""Shared helper of the example scripts: a model from the precomputed energies of one of the reference's fixtures.

The reference's scripts start from a CHARMM structure and run MEAD (``tests/<name>/<name>.py``); the hot path starts where
MEAD ends, so here the model is built from ``tests/golden/<name>.npz`` -- Gintr, protons and the raw interaction matrix
rebuilt from the MEAD logs cached in the reference's ``results.tgz`` (``tests/golden/make_golden.py``)."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load(name):
    """-> (ModelRecord, golden arrays) of fixture ``name``."""
    from pcetk_b200 import ModelRecord

    data = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    return ModelRecord.from_npz_dict(data), data


def output_directory(argv, default):
    directory = argv[1] if len(argv) > 1 else default
    os.makedirs(directory, exist_ok=True)
    return directory
