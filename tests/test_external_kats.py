"""CPU: both oracles against values PUBLISHED OUTSIDE this repository (tests/golden/external_kats.json).

The reference holds no golden vectors for this path (random thread_rng inputs, tests/msm/mod.rs:66-90; NTT /
Poseidon golden files are external), so these published constants are what pins the oracle -- and, through
tests/test_external_kats_gpu.py, the CUDA path -- to the outside world: curve generators and their small
multiples, the arkworks 2-adic roots of unity that define the NTT, the Poseidon reference test vector."""
import json
import os

import numpy as np

from oracle.py import curves, ec, ntt as pyntt, poseidon as P

KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "external_kats.json")))


def bls381_compressed(c, pt):
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if y > (c.q - 1) // 2 else 0)
    return b.hex()


def test_bls12_381_generator_multiples_both_oracles(oracle):
    c = curves.BLS12_381
    G = (c.gx, c.gy)
    for v in KATS["bls12_381_g1_pubkeys"]["vectors"]:
        assert bls381_compressed(c, ec.scalar_mul(c, v["k"], G)) == v["compressed"]
        got = oracle.point_mul("BLS12_381", ec.encode_point(c, G), v["k"])
        assert bls381_compressed(c, ec.decode_point(c, got)) == v["compressed"]
        # and as a one-element MSM through the arkworks-style Pippenger
        rec = oracle.msm_pippenger("BLS12_381", ec.encode_point(c, G), v["k"].to_bytes(32, "little"), 1)
        assert bls381_compressed(c, ec.decode_result(c, rec)) == v["compressed"]


def test_bn254_double_both_oracles(oracle):
    c = curves.BN254
    k = KATS["bn254_g1_double"]
    exp = (int(k["x"], 16), int(k["y"], 16))
    assert (c.gx, c.gy) == (1, 2)
    assert ec.scalar_mul(c, 2, (1, 2)) == exp
    assert ec.decode_point(c, oracle.point_mul("BN254", ec.encode_point(c, (1, 2)), 2)) == exp
    assert ec.decode_point(c, oracle.point_add("BN254", ec.encode_point(c, (1, 2)), ec.encode_point(c, (1, 2)))) == exp


def test_bls12_377_generator(oracle):
    c = curves.BLS12_377
    k = KATS["bls12_377_g1_generator"]
    assert (c.gx, c.gy) == (int(k["x"]), int(k["y"]))
    assert oracle.on_curve("BLS12_377", ec.encode_point(c, (c.gx, c.gy)))
    assert oracle.point_mul("BLS12_377", ec.encode_point(c, (c.gx, c.gy)), c.r) is None


def test_two_adic_roots_pin_the_ntt(oracle):
    """The published arkworks root constants equal what both oracles derive, and the oracle NTT is the DFT at the powers
    of that root in natural order (ark-poly Radix2EvaluationDomain::fft semantics)."""
    for name, k in KATS["two_adic_roots"].items():
        if name == "source":
            continue
        c = curves.CURVES[name]
        root = int(k["root_hex"], 16) if "root_hex" in k else int(k["root_dec"])
        s = k["two_adicity"]
        assert pow(k["generator"], (c.r - 1) >> s, c.r) == root
        assert pow(root, 1 << s, c.r) == 1 and pow(root, 1 << (s - 1), c.r) == c.r - 1
        assert curves.root_of_unity(c, s) == root
        if "root_montgomery_limbs_le64" in k:
            raw = sum(int(l, 16) << (64 * i) for i, l in enumerate(k["root_montgomery_limbs_le64"]))
            assert raw * pow(1 << 256, -1, c.r) % c.r == root
        # C++ oracle: NTT(delta_1)[j] = w^j with w = root^(2^(s - log_n)), natural order
        log_n = 6
        n = 1 << log_n
        w = pow(root, 1 << (s - log_n), c.r)
        d = np.frombuffer(pyntt.encode([0, 1] + [0] * (n - 2)), dtype=np.uint8).copy()
        oracle.ntt(name, d, log_n)
        assert pyntt.decode(bytes(d)) == [pow(w, j, c.r) for j in range(n)]
        v = np.frombuffer(pyntt.encode([3, 1, 4, 1, 5, 9, 2, 6]), dtype=np.uint8).copy()
        w8 = pow(root, 1 << (s - 3), c.r)
        exp = [sum(x * pow(w8, j * i, c.r) for j, x in enumerate([3, 1, 4, 1, 5, 9, 2, 6])) % c.r for i in range(8)]
        oracle.ntt(name, v, 3)
        assert pyntt.decode(bytes(v)) == exp
        assert oracle.ntt_eval(name, np.frombuffer(pyntt.encode([3, 1, 4, 1, 5, 9, 2, 6]), dtype=np.uint8).copy(), 3,
                               list(range(8))) == exp


def test_poseidon_reference_vector():
    k = KATS["poseidon_x5_255_3"]
    assert (P.R_F, P.R_P[k["t"]]) == (k["r_f"], k["r_p"])
    out = P.permute([int(x) for x in k["input"]], P.MDS_GRAIN)
    assert [format(x, "064x") for x in out] == k["output"]
    # the product's instances share generator and round structure and differ only in the (Cauchy) MDS
    rc9, mds9 = P.params(9)
    assert mds9[2][3] * (2 + 9 + 3) % P.R_ == 1
