"""Minimal jax.random-style key handling for the host side (PRNGKey only); the Threefry
generator that consumes keys runs on the device (csrc/rpchol.cu, csrc/lanczos.cu)."""
import numpy as np


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def as_key(key) -> np.ndarray:
    k = np.asarray(key)
    if k.shape != (2,):
        raise ValueError("a PRNG key is two uint32 words, e.g. gpx_b200.random.PRNGKey(2023)")
    return k.astype(np.uint32)
