"""Seeded inputs of the golden op cases (shared by make_golden.py and the tests; outputs are committed)."""
import numpy as np


def golden_inputs():
    rng = np.random.RandomState(123)
    g = {}
    g["c1_x"] = rng.normal(size=(1, 6890, 3)).astype(np.float32)
    g["c1_W"] = np.clip(rng.normal(0, 0.1, size=(18, 64)), -0.2, 0.2).astype(np.float32)
    g["cnp_x"] = rng.normal(size=(2, 6890, 16)).astype(np.float32)
    g["cnp_W"] = np.clip(rng.normal(0, 0.1, size=(32, 32)), -0.2, 0.2).astype(np.float32)
    g["cnp_b"] = rng.normal(0, 0.1, size=(32,)).astype(np.float32)
    g["up_x"] = rng.normal(size=(2, 3445, 8)).astype(np.float32)
    return g
