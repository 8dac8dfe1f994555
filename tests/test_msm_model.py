"""CPU: integer models of the MSM bookkeeping that the CUDA kernels implement (blaze_b200/csrc/msm_sort.cu,
msm_curve.cuh, msm_client.cu) -- the group is replaced by the integers, so every identity the kernels rely on can be
checked exactly without a GPU:
  * signed-digit recoding (k_digits): sum_w d_w 2^(c w) == s, |d_w| <= 2^(c-1), top digit unsigned;
  * sort key / bucket slot permutation (sort_key, k_final, k_reduce_level's slot()): a bijection, partition levels
    consume the LOW bits of b-1 first;
  * multi-level running-sum reduction with pre-scaled chunk sums (k_reduce_level): V_top == sum_i (i+1) A[slot(i)];
  * window-merged table: sum_i s_i P_i == sum over (w, i) of d_{w,i} * (2^(c w) P_i) with ONE bucket set.
"""
import random

import pytest

from oracle.py import curves


def plan_windows(smax, sbits, c):
    """msm_client.cu plan_windows: smallest W >= ceil(sbits / c) whose top digit of the largest scalar is <= 2^(c-1)."""
    W = max(1, (sbits + c - 1) // c)
    while True:
        K = sum(1 << (c - 1 + c * w) for w in range(W - 1))
        if ((smax + K) >> (c * (W - 1))) <= (1 << (c - 1)):
            return W, K
        W += 1


def digits(s, c, W, K):
    """k_digits: branch-free recoding s' = s + K; digit_w = field_w(s') - 2^(c-1), top window takes the rest."""
    sp = s + K
    half = 1 << (c - 1)
    out = []
    for w in range(W):
        v = sp >> (c * w)
        if w < W - 1:
            out.append((v & ((1 << c) - 1)) - half)
        else:
            assert v <= half
            out.append(v)
    return out


@pytest.mark.parametrize("name", ["BLS12_381", "BLS12_377", "BN254"])
@pytest.mark.parametrize("c", [4, 11, 16, 20, 22, 24, 26])
def test_signed_digits_recompose(name, c):
    r = curves.CURVES[name].r
    W, K = plan_windows(r - 1, r.bit_length(), c)
    rng = random.Random(c)
    for s in [0, 1, r - 1, r - 2, (1 << (r.bit_length() - 1)), (1 << c) - 1, 1 << (c - 1)] + [rng.randrange(r) for _ in range(200)]:
        d = digits(s, c, W, K)
        assert sum(dw << (c * w) for w, dw in enumerate(d)) == s
        assert all(abs(dw) <= (1 << (c - 1)) for dw in d) and d[-1] >= 0


def sort_levels(kb, Ms):
    """msm_client.cu make_plan: bits of the partition levels (rest) and of the final in-CTA level (fb)."""
    rest_t = 0
    while rest_t < 30 and Ms / (1 << rest_t) > 8192.0:
        rest_t += 1
    lo, hi = max(kb - 8, 0), kb
    if lo <= 16:
        hi = min(hi, 16)
    rest = max(lo, min(rest_t, hi))
    fb = kb - rest
    nlev = max(1, (rest + 7) // 8)
    lbits, left = [], rest
    for l in range(nlev):
        b = (left + (nlev - l) - 1) // (nlev - l)
        lbits.append(b)
        left -= b
    return rest, fb, lbits


def sort_key(b, rest, fb):
    v = b - 1
    return ((v & ((1 << rest) - 1)) << fb) | (v >> rest)


@pytest.mark.parametrize("kb,Ms", [(3, 10), (10, 3000), (15, 1 << 16), (19, 1 << 26), (21, 12 << 23), (23, 11 << 26), (25, 10 << 20)])
def test_sort_key_is_a_bijection_and_levels_cover_the_key(kb, Ms):
    rest, fb, lbits = sort_levels(kb, Ms)
    assert 0 <= fb <= 8 and sum(lbits) == rest and all(0 <= b <= 8 for b in lbits) and len(lbits) <= 4
    n = 1 << kb
    if kb <= 15:
        keys = sorted(sort_key(b, rest, fb) for b in range(1, n + 1))
        assert keys == list(range(n))
    # reduction side: value i (= b - 1) is read at slot (i mod 2^rest) * 2^fb + (i >> rest)
    rng = random.Random(kb)
    for _ in range(200):
        b = rng.randrange(1, n + 1)
        i = b - 1
        assert sort_key(b, rest, fb) == (i & ((1 << rest) - 1)) * (1 << fb) + (i >> rest)
    # small digits (low bits only) still spread over the partition parents
    parents = {sort_key(b, rest, fb) >> fb for b in range(1, min(n, 1 << min(rest, 12)) + 1)}
    assert len(parents) == min(n, 1 << min(rest, 12))


def reduce_levels(A, slot, chunk0=16):
    """k_reduce_level recursion on integers: returns V_top."""
    n = len(A)
    arr = [A[slot(i)] for i in range(n)]      # level 0 reads through the slot permutation
    Vin, level = None, 0
    while True:
        s = chunk0 if level == 0 else 4
        nch = (n + s - 1) // s
        Sout, Vout = [], []
        for k in range(nch):
            lo, hi = k * s, min(k * s + s, n)
            S = R = 0
            for i in range(hi - 1, lo, -1):
                S += arr[i]
                R += S
            S += arr[lo]
            if level == 0:
                R += S                          # weights t + 1 at level 0 (slot i holds bucket value i + 1)
            if Vin is not None:
                R += sum(Vin[lo:hi])
            Vout.append(R)
            Sout.append(S * (s if nch > 1 else 1))   # next level gets the chunk sums already multiplied by s
        if nch == 1:
            return Vout[0]
        arr, Vin, n, level = Sout, Vout, nch, level + 1


@pytest.mark.parametrize("kb", [1, 4, 7, 10, 13])
def test_bucket_reduction_recursion(kb):
    rest, fb, _ = sort_levels(kb, 1 << 20)
    n = 1 << kb
    rng = random.Random(kb)
    A = [rng.randrange(1 << 40) if rng.random() < 0.8 else 0 for _ in range(n)]
    slot = lambda i: (i & ((1 << rest) - 1)) * (1 << fb) + (i >> rest)
    exp = sum((i + 1) * A[slot(i)] for i in range(n))
    for chunk0 in (8, 16):
        assert reduce_levels(A, slot, chunk0) == exp


def test_window_merged_table_identity():
    """One bucket set for all windows: sum_i s_i P_i == sum_b b * (sum of +-T[w][i] with |d_{w,i}| == b), T[w][i] = 2^(c w) P_i."""
    r = curves.BLS12_381.r
    rng = random.Random(5)
    n, c = 200, 11
    W, K = plan_windows(r - 1, r.bit_length(), c)
    P = [rng.randrange(1 << 60) for _ in range(n)]             # "points" = integers
    s = [rng.randrange(r) for _ in range(n)]
    table = [[p << (c * w) for p in P] for w in range(W)]
    buckets = {}
    for i in range(n):
        for w, d in enumerate(digits(s[i], c, W, K)):
            if d:
                buckets[abs(d)] = buckets.get(abs(d), 0) + (table[w][i] if d > 0 else -table[w][i])
    assert sum(b * v for b, v in buckets.items()) == sum(si * pi for si, pi in zip(s, P))
