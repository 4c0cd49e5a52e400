"""Synthetic inputs for the dev tools (PCG64, uniform bases).  Kept here so that nothing under tools/
touches oracle/, which is test infrastructure."""
import numpy as np


def synthetic_codes(n, L=101, seed=3):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 4, size=(n, L), dtype=np.uint8)
