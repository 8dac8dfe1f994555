"""CPU: the per-segment routine of the batched-affine bucket accumulation (blaze_b200/csrc/msm_ba2.cuh) -- the code each
CUDA lane runs, compiled with g++ against the carry-flag emulation -- against big-integer bucket sums, over bucket shapes
that hit every branch: runs cut by segment boundaries, odd tails, empty buckets, one huge bucket, many tiny runs (scratch
fallback), duplicates (tangent), P + (-P) (identity results), identity table entries.  (Test vehicle only.)"""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle.py import curves, ec

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hb():
    so = os.path.join(HERE, "libhostba.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "host_ba_check.cpp"), "-o", so])
    return ctypes.CDLL(so)


def run_case(hb, c, pts, bucket_entries, L, rounds, cap):
    """pts: list of affine points or None (identity record); bucket_entries: list (per bucket) of lists of (index, negate)."""
    ps = c.point_size
    wire = b"".join(ec.encode_point(c, p) if p is not None else bytes(ps) for p in pts)
    sorted_, goff = [], [0]
    for ent in bucket_entries:
        for idx, neg in ent:
            sorted_.append(idx | (0x80000000 if neg else 0))
        goff.append(len(sorted_))
    total, ngoff = len(sorted_), len(bucket_entries)
    s_arr = np.array(sorted_ + [0], dtype=np.uint32)
    g_arr = np.array(goff, dtype=np.uint32)
    out = ctypes.create_string_buffer(ngoff * ps)
    rc = hb.hc_ba_accumulate(c.code, wire, len(pts), s_arr.ctypes.data_as(ctypes.c_void_p), total,
                             g_arr.ctypes.data_as(ctypes.c_void_p), ngoff, L, rounds, cap, out)
    assert rc == 0
    for g, ent in enumerate(bucket_entries):
        acc = None
        for idx, neg in ent:
            p = pts[idx]
            if p is None:
                continue
            acc = ec.add(c, acc, ec.neg(c, p) if neg else p)
        got = out.raw[g * ps:(g + 1) * ps]
        exp = ec.encode_point(c, acc) if acc is not None else bytes(ps)
        assert got == exp, (g, len(ent))


@pytest.mark.parametrize("name", ["BLS12_381", "BN254", "BLS12_377"])
def test_ba_segments_vs_bigint(hb, name):
    c = curves.CURVES[name]
    rng = random.Random(7 + c.code)
    G = (c.gx, c.gy)
    base = [ec.scalar_mul(c, rng.randrange(1, c.r), G) for _ in range(24)]
    pts = list(base) + [None]                       # index 24: identity record (zero-filled HBM)
    npt = len(base)

    def rnd_bucket(k):
        return [(rng.randrange(npt), rng.random() < 0.5) for _ in range(k)]

    # 1. typical: runs of varied length, empty buckets, cut by segment boundaries; several (L, rounds)
    shape = [0, 37, 1, 0, 0, 64, 5, 129, 2, 3, 0, 41, 16, 7]
    buckets = [rnd_bucket(k) for k in shape]
    for L, rounds, cap in ((32, 2, 24), (64, 3, 40), (16, 1, 12), (64, 5, 40), (256, 4, 160)):
        run_case(hb, c, pts, buckets, L, rounds, cap)
    # 2. one huge bucket of duplicates (the reference's tiled vectors): tangent at every pair, spans many segments
    run_case(hb, c, pts, [[(3, False)] * 200, [(4, True)] * 3], 32, 3, 24)
    # 3. cancellations and identity operands inside the tree
    b = [(1, False), (1, True), (2, False), (2, True), (5, False), (24, False), (24, True), (6, False), (6, False), (6, True),
         (24, False), (24, False), (7, True)]
    run_case(hb, c, pts, [b, list(reversed(b)) + b, [(24, False)] * 5], 16, 3, 12)
    run_case(hb, c, pts, [b * 5], 64, 4, 40)
    # 4. many tiny runs: more half-slots than the scratch holds -> the segment folds directly (no tree)
    run_case(hb, c, pts, [rnd_bucket(1) for _ in range(90)], 32, 2, 10)
    # 5. rounds = 0 (pure XYZZ fold) and a single entry
    run_case(hb, c, pts, [rnd_bucket(9), rnd_bucket(1)], 8, 0, 8)
    run_case(hb, c, pts, [[(0, True)]], 64, 3, 40)
