"""Shared input builders for the parity tests (seeded, reproducible; the reference's own tests use
thread_rng and are not reproducible -- tests/msm/mod.rs:66,82,86)."""
import numpy as np

from oracle import capi
from oracle.py import curves, ec

CURVE_BY_NAME = {"BLS12_381": curves.BLS12_381, "BLS12_377": curves.BLS12_377, "BN254": curves.BN254}


def seed_points(c, seed):
    """Two deterministic subgroup points (P0, Q) as wire bytes."""
    g = ec.encode_point(c, (c.gx, c.gy))
    k0 = (0x9E3779B97F4A7C15 * (seed + 1)) % c.r
    k1 = (0xC2B2AE3D27D4EB4F * (seed + 7) + 12345) % c.r
    return capi.point_mul(c.name, g, k0), capi.point_mul(c.name, g, k1)


def random_scalars(c, n, seed):
    """n canonical scalars, uniform in [0, r), as a uint8 array of n*32 bytes (rejection sampling on
    r.bit_length()-bit draws, vectorised)."""
    rng = np.random.default_rng(seed)
    bits = c.r.bit_length()
    top_mask = np.uint64((1 << (bits - 192)) - 1)
    rw = [np.uint64((c.r >> (64 * k)) & 0xFFFFFFFFFFFFFFFF) for k in range(4)]

    def draw(m):
        w = rng.integers(0, 1 << 64, size=(m, 4), dtype=np.uint64)
        w[:, 3] &= top_mask
        return w

    def ge_r(w):   # lexicographic (most significant word first) w >= r
        ge = np.ones(len(w), dtype=bool)        # equal so far
        res = np.zeros(len(w), dtype=bool)
        for k in (3, 2, 1, 0):
            res |= ge & (w[:, k] > rw[k])
            ge &= w[:, k] == rw[k]
        return res | ge

    w = draw(n)
    bad = np.nonzero(ge_r(w))[0]
    while len(bad):
        w[bad] = draw(len(bad))
        bad = bad[ge_r(w[bad])]
    return w.view(np.uint8).reshape(-1)


def chain_points(c, n, seed=0):
    """n distinct points P0 + i*Q (wire format, factor 1) + (p0, q)."""
    p0, q = seed_points(c, seed)
    return capi.chain_points(c.name, p0, q, n), p0, q


def precompute_bases(c, pts: np.ndarray, n: int, factor: int) -> np.ndarray:
    """tests/msm/mod.rs:360-380: record k = P_k || 2^32 P_k || ... || 2^(32(f-1)) P_k."""
    ps = c.point_size
    out = np.zeros(n * ps * factor, dtype=np.uint8)
    for k in range(n):
        p = bytes(pts[k * ps:(k + 1) * ps])
        rec = p
        for i in range(1, factor):
            rec += capi.point_mul(c.name, p, pow(2, 32 * i, c.r))
        out[k * ps * factor:(k + 1) * ps * factor] = np.frombuffer(rec, dtype=np.uint8)
    return out


def tile(arr: np.ndarray, rec: int, n_block: int, n: int) -> np.ndarray:
    """The reference's large-input construction: a 256-element block repeated (mod.rs:92-109)."""
    reps = n // n_block
    rest = n % n_block
    parts = [arr[:n_block * rec]] * reps + [arr[:rest * rec]]
    return np.concatenate(parts)
