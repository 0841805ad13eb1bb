"""The BVH refit (madrona_b200/csrc/kernels_physics.cu: atomicMinFloat / atomicMaxFloat) grows
node bounds with ONE native integer atomic on the float's bit pattern: value >= 0 -> signed
min / max, value < 0 -> unsigned max / min.  CPU model over every sign combination, zeros,
infinities, denormals: the stored result must equal the float min / max (+0 and -0 are the same
bound for every comparison the broadphase makes)."""
import numpy as np


def _bits(x):
    return np.array(x, dtype=np.float32).view(np.uint32)


def _as_float(b):
    return np.array(b, dtype=np.uint32).view(np.float32)


def atomic_min_float(old, value):
    if value >= 0:
        return _as_float(np.minimum(_bits(old).view(np.int32), _bits(value).view(np.int32)).view(np.uint32))
    return _as_float(np.maximum(_bits(old), _bits(value)))


def atomic_max_float(old, value):
    if value >= 0:
        return _as_float(np.maximum(_bits(old).view(np.int32), _bits(value).view(np.int32)).view(np.uint32))
    return _as_float(np.minimum(_bits(old), _bits(value)))


def test_integer_atomics_on_bit_patterns_order_like_floats():
    rng = np.random.default_rng(3)
    special = np.array([0.0, -0.0, 1e-42, -1e-42, 1.0, -1.0, 3.4e38, -3.4e38, np.inf, -np.inf, 1e-7, -1e-7],
                       dtype=np.float32)
    pool = np.concatenate([special, rng.normal(scale=100, size=400).astype(np.float32),
                           (rng.normal(size=100) * 1e-30).astype(np.float32)])
    for old in pool:
        for value in pool[::7]:
            lo, hi = atomic_min_float(old, value), atomic_max_float(old, value)
            assert lo == min(old, value), (old, value, lo)       # == treats +0 / -0 alike
            assert hi == max(old, value), (old, value, hi)
