"""Big-integer NTT with arkworks `Radix2EvaluationDomain::fft` semantics.

TEST INFRASTRUCTURE (oracle) — see curves.py header.

The reference never names the field / root / ordering of its 2^27 NTT
(/root/reference/src/ingo_ntt/ntt_api.rs:20-23,110-124 only fix the I/O as a
flat little-endian vector of 32-byte elements); BASELINE.json fixes BLS12-381 Fr
and arkworks semantics: natural order in, natural order out,
    out[k] = sum_j in[j] * w^(j k),  w = Fr::get_root_of_unity(n).
**Parity unpinned** against the reference's own golden files (they are not in the
repository); pinned here by the O(n^2) definition below.
"""
from .curves import CurveParams, root_of_unity


def dft_definition(c: CurveParams, a, inverse=False):
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = root_of_unity(c, log_n)
    if inverse:
        w = pow(w, -1, c.r)
    out = []
    for k in range(n):
        wk = pow(w, k, c.r)
        acc, t = 0, 1
        for j in range(n):
            acc += a[j] * t
            t = t * wk % c.r
        out.append(acc % c.r)
    if inverse:
        ninv = pow(n, -1, c.r)
        out = [x * ninv % c.r for x in out]
    return out


def ntt(c: CurveParams, a, inverse=False):
    """Iterative radix-2 (bit-reverse then DIT butterflies); same values as dft_definition."""
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    r = c.r
    a = list(a)
    for i in range(n):
        j = int(format(i, "0%db" % log_n)[::-1], 2) if log_n else 0
        if i < j:
            a[i], a[j] = a[j], a[i]
    w_n = root_of_unity(c, log_n)
    if inverse:
        w_n = pow(w_n, -1, r)
    m = 1
    while m < n:
        w_m = pow(w_n, n // (2 * m), r)
        for s in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                u = a[s + j]
                t = a[s + j + m] * w % r
                a[s + j] = (u + t) % r
                a[s + j + m] = (u - t) % r
                w = w * w_m % r
        m *= 2
    if inverse:
        ninv = pow(n, -1, r)
        a = [x * ninv % r for x in a]
    return a


def encode(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def decode(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]
