"""TEST INFRASTRUCTURE ONLY -- numpy / pure-Python restatements of pieces of the
reference's algorithm that the compiled CPU backend (oracle/_ref) does not
cover or that tests need in closed form.  Never imported by madrona_b200/.

Pinned against the reference: tests/test_oracle_kat.py checks threefry against
the reference's known-answer values (tests/rand.cpp:131-141) and against the
reference headers compiled here (oracle/_ref/kat_probe_ref); the sort
restatement is checked against the reference CPU backend's compaction through
the gridworld golden traces.
"""
from __future__ import annotations

import numpy as np

_ROT = (13, 15, 26, 6, 17, 29, 16, 24)
_M32 = 0xFFFFFFFF


def _rotl(v, d):
    return ((v << d) | (v >> (32 - d))) & _M32


def split_i(key, idx, idx_upper=0):
    """threefry2x32-20 block function (include/madrona/rand.inl:31-103)."""
    a, b = key
    ks = (a, b, 0x1BD11BDA ^ a ^ b)
    x0 = (idx + ks[0]) & _M32
    x1 = (idx_upper + ks[1]) & _M32
    for g in range(5):
        rots = _ROT[4:] if g & 1 else _ROT[:4]
        for r in rots:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & _M32
    return (x0, x1)


def init_key(seed, seed_upper=0):
    return split_i((seed & _M32, seed_upper & _M32), 0)


def bits32(key):
    return key[0] ^ key[1]


def bits_to_float01(bits):
    return np.float32(bits >> 8) * np.float32(2.0 ** -24)


def sample_uniform(key):
    return bits_to_float01(bits32(key))


def sample_i32(key, a, b):
    """Lemire's unbiased bounded integer (rand.inl:110-160)."""
    s = (b - a) & _M32
    m = bits32(key) * s
    lo = m & _M32
    if lo < s:
        t = ((-s) & _M32) % s
        while lo < t:
            key = split_i(key, 0)
            m = bits32(key) * s
            lo = m & _M32
    hi = m >> 32
    v = (hi + a) & _M32
    return v - (1 << 32) if v >= (1 << 31) else v


def sample_i32_biased(key, a, b):
    return (bits32(key) * ((b - a) & _M32)) >> 32


class RNG:
    """madrona::RNG (rand.inl:226-277): key + running counter."""

    def __init__(self, seed):
        self.key = init_key(seed)
        self.count = 0

    def _advance(self):
        k = split_i(self.key, self.count)
        self.count += 1
        return k

    def sample_uniform(self):
        return sample_uniform(self._advance())

    def sample_i32(self, a, b):
        return sample_i32(self._advance(), a, b)


# ---- archetype sort semantics (src/mw/device/sort_archetype.cpp:977-1551) ----

def world_sort_passes(num_worlds: int) -> int:
    bits = int(num_worlds).bit_length()
    return max(1, (bits + 7) // 8)


def sort_archetype(keys: np.ndarray, num_worlds: int, world_sort: bool = True):
    """Returns (perm, new_num_rows, world_offsets, world_counts): stable sort of
    rows by the low 8*P key bits; rows with key -1 (destroyed) sort last and are
    dropped; empty worlds get offset = new_num_rows, count = 0."""
    keys = np.asarray(keys).astype(np.uint32)
    passes = world_sort_passes(num_worlds) if world_sort else 4
    mask = np.uint32((1 << (8 * passes)) - 1) if passes < 4 else np.uint32(0xFFFFFFFF)
    perm = np.argsort(keys & mask, kind="stable")
    if not world_sort:
        return perm, len(keys), None, None
    new_n = int((keys != np.uint32(0xFFFFFFFF)).sum())
    sorted_keys = keys[perm][:new_n]
    offsets = np.full(num_worlds, new_n, dtype=np.int32)
    counts = np.zeros(num_worlds, dtype=np.int32)
    if new_n:
        uniq, first, cnt = np.unique(sorted_keys, return_index=True, return_counts=True)
        offsets[uniq] = first
        counts[uniq] = cnt
    return perm[:new_n], new_n, offsets, counts
