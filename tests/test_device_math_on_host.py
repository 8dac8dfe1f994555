"""CPU: the DEVICE field/curve code (blaze_b200/csrc/ff.cuh, ec.cuh) compiled with g++ against the
carry-flag emulation in bz_common.cuh -- the exact limb schedules of the GPU kernels, checked
bit-for-bit against big-integer arithmetic.  (Test vehicle only; see tests/host_ff_check.cpp.)"""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.py import curves, ec

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HERE, "libhostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "host_ff_check.cpp"), "-o", so])
    return ctypes.CDLL(so)


def fop(hc, fid, op, a, b, nb):
    out = ctypes.create_string_buffer(nb)
    assert hc.hc_field_op(fid, op, a.to_bytes(nb, "little"), b.to_bytes(nb, "little"), out) == 0
    return int.from_bytes(out.raw, "little")


@pytest.mark.parametrize("name", ["BLS12_381", "BLS12_377", "BN254"])
def test_field_ops(hc, name):
    c = curves.CURVES[name]
    rng = random.Random(3)
    for fid, p, nb in ((c.code, c.q, c.fq_bytes), (10 + c.code, c.r, 32)):
        vals = [0, 1, 2, p - 1, p - 2, 1 << (p.bit_length() - 1), p >> 1] + [rng.randrange(p) for _ in range(150)]
        for i, a in enumerate(vals):
            b = vals[(i * 7 + 3) % len(vals)]
            assert fop(hc, fid, 0, a, b, nb) == a * b % p
            assert fop(hc, fid, 1, a, b, nb) == (a + b) % p
            assert fop(hc, fid, 2, a, b, nb) == (a - b) % p
            assert fop(hc, fid, 3, a, b, nb) == a * a % p
            assert fop(hc, fid, 5, a, b, nb) == (-a) % p
            assert fop(hc, fid, 6, a, b, nb) == 2 * a % p
            assert fop(hc, fid, 7, a, b, nb) == (a * b - (a + b) * (a - b)) % p      # fused two-product reduction
            assert fop(hc, fid, 8, a, b, nb) == (a * b + (a + b) * (a - b)) % p
            assert fop(hc, fid, 9, a, b, nb) == a * b % p                               # Karatsuba + reduce-only
            assert fop(hc, fid, 10, a, b, nb) == (a * b + (a + b) * (a - b)) % p
            assert fop(hc, fid, 13, a, b, nb) == a * a % p                              # split-call forms (BZ_SPLIT_MUL)
            assert fop(hc, fid, 14, a, b, nb) == a * b % p
            assert fop(hc, fid, 15, a, b, nb) == (a * b + (a + b) * (a - b)) % p
        for a in vals[1:10]:
            assert fop(hc, fid, 4, a, 0, nb) == pow(a, -1, p)
            assert fop(hc, fid, 12, a, 0, nb) == pow(a, -1, p)                        # Fermat ladder
        for a in vals[1:]:
            assert fop(hc, fid, 11, a, 0, nb) == pow(a, -1, p)                        # division steps (what inv() uses)
        assert fop(hc, fid, 11, 0, 0, nb) == 0 and fop(hc, fid, 4, 0, 0, nb) == 0


@pytest.mark.parametrize("name", ["BLS12_381", "BLS12_377", "BN254"])
def test_xyzz_group_law(hc, name):
    c = curves.CURVES[name]
    rng = random.Random(4)
    G = (c.gx, c.gy)
    P = ec.scalar_mul(c, rng.randrange(c.r), G)
    Q = ec.scalar_mul(c, rng.randrange(c.r), G)

    def cop(op, P, Q, k1, k2):
        out = ctypes.create_string_buffer(c.point_size)
        rc = hc.hc_curve_op(c.code, op, ec.encode_point(c, P), ec.encode_point(c, Q), k1, k2, out)
        return None if rc == 1 else ec.decode_point(c, out.raw)

    for k1, k2 in ((1, 0), (2, 0), (3, 1), (5, 7), (0, 1), (1, 1)):
        exp = ec.add(c, ec.scalar_mul(c, k1, P), ec.scalar_mul(c, k2, Q))
        assert cop(0, P, Q, k1, k2) == exp          # repeated madd: P+P hits the doubling branch
        assert cop(1, P, Q, k1, k2) == exp          # mul_small + full add
    assert cop(0, P, ec.neg(c, P), 1, 1) is None    # P + (-P)
    assert cop(0, P, ec.neg(c, P), 2, 1) == P
    assert cop(2, P, Q, 0, 0) == ec.scalar_mul(c, 2, ec.add(c, P, ec.neg(c, Q)))
    assert cop(2, P, P, 0, 0) is None
    assert cop(1, P, P, 3, 3) == ec.scalar_mul(c, 6, P)     # full add with equal inputs -> dbl
    assert cop(1, P, Q, 123456789, 987654321) == ec.add(c, ec.scalar_mul(c, 123456789, P),
                                                        ec.scalar_mul(c, 987654321, Q))
