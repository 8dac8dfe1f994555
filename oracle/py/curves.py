"""Curve and field constants for the three curves blaze's MSM core supports.

TEST INFRASTRUCTURE (oracle). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product path
(blaze_b200/) never does.

Source of the numbers: the reference depends on arkworks 0.3.0 (`ark-bls12-381`,
`ark-bls12-377`, `ark-bn254`; /root/reference/Cargo.toml:14-19), which is not
vendored.  The values below are the published curve parameters (SURVEY.md §8(c)
lists them, verified there with sympy); `self_check()` re-derives every property
this repo relies on (primality is checked probabilistically, generator on curve,
subgroup order, two-adicity and the order of the roots of unity).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class CurveParams:
    name: str
    code: int           # id used across the C ABI (matches include/blaze_b200.h)
    q: int              # base field modulus
    r: int              # scalar field modulus (prime subgroup order of G1)
    b: int              # y^2 = x^3 + b   (a = 0 for all three)
    gx: int
    gy: int
    fq_bytes: int       # wire size of one base-field element (msm_cfg.rs:44-92)
    fr_gen: int         # multiplicative generator of Fr used by arkworks
    fr_two_adicity: int

    @property
    def point_size(self):      # affine x||y
        return 2 * self.fq_bytes

    @property
    def result_point_size(self):   # Z||Y||X
        return 3 * self.fq_bytes

    @property
    def fq_bits(self):
        return self.q.bit_length()

    @property
    def fr_bits(self):
        return self.r.bit_length()


BLS12_377 = CurveParams(
    name="BLS12_377", code=0,
    q=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
    r=0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
    b=1,
    gx=81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
    gy=241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030,
    fq_bytes=48, fr_gen=22, fr_two_adicity=47,
)

BN254 = CurveParams(
    name="BN254", code=1,
    q=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    b=3, gx=1, gy=2,
    fq_bytes=32, fr_gen=5, fr_two_adicity=28,
)

BLS12_381 = CurveParams(
    name="BLS12_381", code=2,
    q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    b=4,
    gx=0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    gy=0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
    fq_bytes=48, fr_gen=7, fr_two_adicity=32,
)

# curve codes follow the reference's image-parameter printer
# (/root/reference/src/ingo_msm/msm_api.rs:359-364): 0=BLS12_377, 1=BN254, 2=BLS12_381
CURVES = {c.name: c for c in (BLS12_377, BN254, BLS12_381)}
BY_CODE = {c.code: c for c in CURVES.values()}


def root_of_unity(curve: CurveParams, log_n: int) -> int:
    """arkworks `FftField::get_root_of_unity(2^log_n)`:
    TWO_ADIC_ROOT_OF_UNITY^(2^(TWO_ADICITY-log_n)), TWO_ADIC_ROOT = g^((r-1)/2^s)."""
    s = curve.fr_two_adicity
    assert 0 <= log_n <= s
    w = pow(curve.fr_gen, (curve.r - 1) >> s, curve.r)
    return pow(w, 1 << (s - log_n), curve.r)


def _is_probable_prime(n, rounds=16):
    import random
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rnd = random.Random(0xB1A2E)
    for _ in range(rounds):
        a = rnd.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def self_check():
    from . import ec
    for c in CURVES.values():
        assert _is_probable_prime(c.q), c.name
        assert _is_probable_prime(c.r), c.name
        assert (c.gy * c.gy - c.gx ** 3 - c.b) % c.q == 0, c.name
        assert (c.r - 1) % (1 << c.fr_two_adicity) == 0
        assert (c.r - 1) % (1 << (c.fr_two_adicity + 1)) != 0
        w = root_of_unity(c, c.fr_two_adicity)
        assert pow(w, 1 << c.fr_two_adicity, c.r) == 1
        assert pow(w, 1 << (c.fr_two_adicity - 1), c.r) == c.r - 1
        assert ec.scalar_mul(c, c.r, (c.gx, c.gy)) is None, c.name   # r*G = infinity
    # the two constants SURVEY.md §8(c) quotes for BLS12-381 Fr
    assert root_of_unity(BLS12_381, 32) == \
        10238227357739495823651030575849232062558860180284477541189508159991286009131
    assert root_of_unity(BLS12_381, 27) == \
        15932505959375582308231798849995567447410469395474322018100309999481287547373
    return True
