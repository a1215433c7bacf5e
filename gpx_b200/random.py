"""jax.random-style key handling for the host side: PRNGKey and split.  Bulk sampling
(Rademacher probes, RPCholesky pivot draws) runs on the device (csrc/threefry.cuh); the host
only derives sub-keys, a handful of Threefry-2x32-20 blocks (Random123 algorithm)."""
import numpy as np

_M = 0xFFFFFFFF


def _rotl(v, r):
    return ((v << r) | (v >> (32 - r))) & _M


def _threefry2x32(k0, k1, x0, x1):
    ks = (k0, k1, (k0 ^ k1 ^ 0x1BD11BDA) & _M)
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    x0 = (x0 + ks[0]) & _M
    x1 = (x1 + ks[1]) & _M
    for i in range(5):
        for r in rot[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & _M], dtype=np.uint32)


def as_key(key) -> np.ndarray:
    k = np.asarray(key)
    if k.shape != (2,):
        raise ValueError("a PRNG key is two uint32 words, e.g. gpx_b200.random.PRNGKey(2023)")
    return k.astype(np.uint32)


def split(key, num: int = 2, partitionable: bool = True) -> np.ndarray:
    """jax.random.split(key, num) -> (num, 2) uint32, for either jax_threefry_partitionable mode."""
    k0, k1 = (int(v) for v in as_key(key))
    out = np.zeros((num, 2), dtype=np.uint32)
    if partitionable:
        for i in range(num):
            out[i] = _threefry2x32(k0, k1, 0, i)
        return out
    flat = []
    n = 2 * num
    ys = [_threefry2x32(k0, k1, i, n // 2 + i) for i in range(n // 2)]
    flat = [y[0] for y in ys] + [y[1] for y in ys]
    return np.array(flat, dtype=np.uint32).reshape(num, 2)


# ---- bulk draws on the host (small shapes only: the sampling of gpx/models/utils.py:52-75; the big draws --
# Rademacher probes, RPCholesky pivots -- run on the device) ----
def _threefry2x32_np(k0, k1, x0, x1):
    """Vectorised Threefry-2x32-20 on uint32 arrays (same rounds as _threefry2x32 above)."""
    m = np.uint64(_M)
    x0 = np.asarray(x0, dtype=np.uint64)
    x1 = np.asarray(x1, dtype=np.uint64)
    ks = (np.uint64(k0), np.uint64(k1), np.uint64((int(k0) ^ int(k1) ^ 0x1BD11BDA) & _M))
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    x0 = (x0 + ks[0]) & m
    x1 = (x1 + ks[1]) & m
    for i in range(5):
        for r in rot[i % 2]:
            x0 = (x0 + x1) & m
            x1 = (((x1 << np.uint64(r)) | (x1 >> np.uint64(32 - r))) & m) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & m
        x1 = (x1 + ks[(i + 2) % 3] + np.uint64(i + 1)) & m
    return x0, x1


def uniform(key, shape=(), partitionable: bool = True) -> np.ndarray:
    """jax.random.uniform(key, shape, float64) in [0, 1): 64 random bits per element (counter (hi, lo) = element
    index when partitionable; first / second half of arange(2 size) in the legacy layout), top 52 bits as mantissa."""
    k0, k1 = (int(v) for v in as_key(key))
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    idx = np.arange(size, dtype=np.uint64)
    if partitionable:
        hi, lo = _threefry2x32_np(k0, k1, idx >> np.uint64(32), idx & np.uint64(_M))
    else:
        hi, lo = _threefry2x32_np(k0, k1, idx & np.uint64(_M), (idx + np.uint64(size)) & np.uint64(_M))
    bits = (hi << np.uint64(32)) | lo
    f = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    return f.reshape(shape)


def normal(key, shape, partitionable: bool = True) -> np.ndarray:
    """jax.random.normal(key, shape, float64) = sqrt(2) erfinv(u), u uniform on (-1, 1) (minval nextafter(-1, 0))."""
    from scipy.special import erfinv

    lo = np.nextafter(-1.0, 0.0)
    u = np.maximum(lo, uniform(key, shape, partitionable) * (1.0 - lo) + lo)
    return np.sqrt(2.0) * erfinv(u)
