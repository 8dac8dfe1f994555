"""CPU: the C++ oracle (oracle/cpp) against the big-integer Python oracle, the golden fixtures
and algebraic identities.  This is what "pins" the oracle (see oracle/cpp/oracle.cpp header)."""
import json
import os
import random

import numpy as np
import pytest

from oracle.py import curves, ec, ntt as pyntt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["BLS12_381", "BLS12_377", "BN254"]


def test_curve_constants_self_check():
    assert curves.self_check()


@pytest.mark.parametrize("name", NAMES)
def test_golden_msm(oracle, name):
    g = json.load(open(os.path.join(GOLD, name.lower() + ".json")))
    c = curves.CURVES[name]
    for case in g["msm"]:
        bases, scal = bytes.fromhex(case["bases"]), bytes.fromhex(case["scalars"])
        exp = bytes.fromhex(case["result"])
        assert oracle.msm_naive(name, bases, scal, case["n"], case["factor"]) == exp
        if case["factor"] == 1:
            assert oracle.msm_pippenger(name, bases, scal, case["n"]) == exp
        # and the python restatement of the literal wire semantics agrees with the fixture
        assert ec.encode_result(c, ec.msm_wire(c, bases, scal, case["n"], case["factor"])) == exp


@pytest.mark.parametrize("name", NAMES)
def test_golden_generator_multiples(oracle, name):
    g = json.load(open(os.path.join(GOLD, name.lower() + ".json")))
    c = curves.CURVES[name]
    G = ec.encode_point(c, (c.gx, c.gy))
    for e in g["generator_multiples"]:
        assert oracle.point_mul(name, G, int(e["k"], 16)) == bytes.fromhex(e["point"])
    assert oracle.point_mul(name, G, c.r) is None           # r*G = infinity
    assert oracle.on_curve(name, G)


@pytest.mark.parametrize("name", NAMES)
def test_golden_ntt(oracle, name):
    g = json.load(open(os.path.join(GOLD, name.lower() + ".json")))
    for case in g["ntt"]:
        d = np.frombuffer(bytes.fromhex(case["in"]), dtype=np.uint8).copy()
        oracle.ntt(name, d, case["log_n"])
        assert bytes(d) == bytes.fromhex(case["out"])
        oracle.ntt(name, d, case["log_n"], inverse=True)
        assert bytes(d) == bytes.fromhex(case["in"])


@pytest.mark.parametrize("name", NAMES)
def test_group_law_vs_bigint(oracle, name):
    c = curves.CURVES[name]
    rng = random.Random(42 + c.code)
    G = (c.gx, c.gy)
    P = ec.scalar_mul(c, rng.randrange(c.r), G)
    Q = ec.scalar_mul(c, rng.randrange(c.r), G)
    eP, eQ = ec.encode_point(c, P), ec.encode_point(c, Q)
    assert ec.decode_point(c, oracle.point_add(name, eP, eQ)) == ec.add(c, P, Q)
    assert ec.decode_point(c, oracle.point_add(name, eP, eP)) == ec.add(c, P, P)        # doubling branch
    assert oracle.point_add(name, eP, ec.encode_point(c, ec.neg(c, P))) is None         # inverse branch
    for k in (0, 1, 2, rng.randrange(c.r), c.r - 1):
        got = oracle.point_mul(name, eP, k)
        exp = ec.scalar_mul(c, k, P)
        assert (got is None and exp is None) or ec.decode_point(c, got) == exp
    a, b = rng.randrange(c.q), rng.randrange(c.q)
    assert int.from_bytes(oracle.fq_mul(name, ec.encode_fq(c, a), ec.encode_fq(c, b)), "little") == a * b % c.q


@pytest.mark.parametrize("name", NAMES)
def test_pippenger_vs_naive_and_closed_form(oracle, name):
    """n = 700 -> arkworks window c = 10*69/100+2 = 8; chain workload closed form agrees."""
    c = curves.CURVES[name]
    n = 700
    from util import chain_points, random_scalars
    pts, p0, q = chain_points(c, n, seed=3)
    sc = random_scalars(c, n, seed=4)
    a = oracle.msm_naive(name, bytes(pts), bytes(sc), n, 1)
    assert oracle.msm_pippenger(name, pts, sc, n) == a
    assert oracle.chain_expected(name, p0, q, sc, n) == a
    # linearity: MSM(2s) = 2 MSM(s)
    two = np.frombuffer(b"".join(((2 * int.from_bytes(bytes(sc[32 * i:32 * i + 32]), "little")) % c.r)
                                 .to_bytes(32, "little") for i in range(n)), dtype=np.uint8).copy()
    r2 = ec.decode_result(c, oracle.msm_pippenger(name, pts, two, n))
    r1 = ec.decode_result(c, a)
    assert r2 == ec.add(c, r1, r1)


def test_ntt_properties(oracle):
    c = curves.BLS12_381
    rng = random.Random(5)
    log_n = 10
    n = 1 << log_n
    v = [rng.randrange(c.r) for _ in range(n)]
    d = np.frombuffer(pyntt.encode(v), dtype=np.uint8).copy()
    oracle.ntt("BLS12_381", d, log_n)
    assert pyntt.decode(bytes(d)) == pyntt.ntt(c, v)
    # NTT(delta_1)[k] = w^k
    e = [0] * n
    e[1] = 1
    d = np.frombuffer(pyntt.encode(e), dtype=np.uint8).copy()
    oracle.ntt("BLS12_381", d, log_n)
    w = curves.root_of_unity(c, log_n)
    assert pyntt.decode(bytes(d)) == [pow(w, k, c.r) for k in range(n)]
