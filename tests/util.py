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
    """n canonical scalars (< r) as a uint8 array of n*32 bytes."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    top_bits = c.r.bit_length() - 1            # force < 2^(bits-1) <= r
    keep = top_bits - 8 * 31
    raw[:, 31] &= (1 << keep) - 1 if keep > 0 else 0
    return raw.reshape(-1)


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
