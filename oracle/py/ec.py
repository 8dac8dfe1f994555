"""Big-integer affine group law, naive MSM and the reference's wire formats.

TEST INFRASTRUCTURE (oracle) — see curves.py header.  Deliberately the most
obvious implementation possible (affine chord/tangent with `pow(x,-1,q)`), used
to pin the faster C++ oracle (oracle/cpp) and, through it, the CUDA path.

Restates /root/reference/tests/msm/mod.rs:
  * expected MSM value  = sum_k aff_k.mul(scalar_k)               (mod.rs:81-90, 327-334)
  * base encoding       = x.to_bytes_le() || y.to_bytes_le(), canonical
                          (non-Montgomery), followed by 2^(32 i) * P for
                          i = 1..factor-1 in the same encoding        (mod.rs:360-380)
  * scalar encoding     = Fr::into_repr().to_bytes_le(), 32 bytes    (mod.rs:331-332)
  * result decoding     = bytes [0,S)=Z, [S,2S)=Y, [2S,3S)=X, little-endian,
                          homogeneous projective x=X/Z, y=Y/Z         (mod.rs:397-405)
A point is `None` (infinity) or an `(x, y)` tuple of Python ints.
"""
from .curves import CurveParams


def is_on_curve(c: CurveParams, P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - c.b) % c.q == 0


def neg(c, P):
    if P is None:
        return None
    return (P[0], (-P[1]) % c.q)


def add(c: CurveParams, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    q = c.q
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, q) % q        # a = 0
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, q) % q
    x3 = (lam * lam - x1 - x2) % q
    y3 = (lam * (x1 - x3) - y1) % q
    return (x3, y3)


def scalar_mul(c: CurveParams, k: int, P):
    if k < 0:
        return scalar_mul(c, -k, neg(c, P))
    R = None
    A = P
    while k:
        if k & 1:
            R = add(c, R, A)
        A = add(c, A, A)
        k >>= 1
    return R


def msm_naive(c: CurveParams, scalars, points):
    acc = None
    for s, P in zip(scalars, points):
        acc = add(c, acc, scalar_mul(c, s, P))
    return acc


# ---------------------------------------------------------------- wire formats
def encode_fq(c, v):
    return int(v).to_bytes(c.fq_bytes, "little")


def encode_point(c, P):
    return encode_fq(c, P[0]) + encode_fq(c, P[1])


def decode_point(c, b):
    s = c.fq_bytes
    return (int.from_bytes(b[:s], "little"), int.from_bytes(b[s:2 * s], "little"))


def encode_base(c, P, precompute_factor=1):
    """tests/msm/mod.rs:360-380 `precompute_base_*`."""
    out = encode_point(c, P)
    for i in range(1, precompute_factor):
        out += encode_point(c, scalar_mul(c, pow(2, 32 * i, c.r), P))
    return out


def encode_scalar(s):
    return int(s).to_bytes(32, "little")


def decode_result(c, b):
    """Z||Y||X homogeneous projective -> affine tuple or None (tests/msm/mod.rs:397-405).
    Like `from_le_bytes_mod_order` the coordinates are reduced mod q."""
    s = c.fq_bytes
    Z = int.from_bytes(b[0:s], "little") % c.q
    Y = int.from_bytes(b[s:2 * s], "little") % c.q
    X = int.from_bytes(b[2 * s:3 * s], "little") % c.q
    if Z == 0:
        return None
    zi = pow(Z, -1, c.q)
    return (X * zi % c.q, Y * zi % c.q)


def encode_result(c, P):
    """Canonical (Z = 1) result record; infinity is Z=0, Y=1, X=0."""
    if P is None:
        return encode_fq(c, 0) + encode_fq(c, 1) + encode_fq(c, 0)
    return encode_fq(c, 1) + encode_fq(c, P[1]) + encode_fq(c, P[0])


def msm_wire(c, bases: bytes, scalars: bytes, n: int, precompute_factor=1):
    """What the MSM core computes from the byte streams, literally:
    sum_k sum_j limb_j(s_k) * B_{k,j} with 32-bit limbs when factor == 8
    (SURVEY.md §8(a) M4), plain sum_k s_k * B_k when factor == 1."""
    ps = c.point_size
    acc = None
    for k in range(n):
        s = int.from_bytes(scalars[32 * k:32 * k + 32], "little")
        rec = bases[k * ps * precompute_factor:(k + 1) * ps * precompute_factor]
        if precompute_factor == 1:
            acc = add(c, acc, scalar_mul(c, s, decode_point(c, rec)))
        else:
            width = 256 // precompute_factor
            for j in range(precompute_factor):
                limb = (s >> (width * j)) & ((1 << width) - 1)
                acc = add(c, acc, scalar_mul(c, limb, decode_point(c, rec[j * ps:(j + 1) * ps])))
    return acc


# ---------------------------------------------------------------- input helpers
def random_point(c, rng):
    """A pseudo-random point of the prime-order subgroup: k*G."""
    return scalar_mul(c, rng.randrange(1, c.r), (c.gx, c.gy))
