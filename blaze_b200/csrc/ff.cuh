// Prime-field arithmetic on 32-bit limbs, Montgomery form, for sm_100a.
//
// One thread owns one field element in registers (12 limbs for the 381/377-bit base
// fields, 8 limbs for the 254/255-bit fields).  All values are kept fully reduced in
// [0, p); "Montgomery form" means a*R mod p with R = 2^(32 N).
//
// mul() is a word-serial Montgomery product arranged as two interleaved accumulators
// ("even"/"odd" columns) so that every 32x32->64 partial product and its carry lands in one
// IMAD.WIDE.U32(.X): per multiplier limb b[i] the schedule is
//     odd  = (odd >> 64) + a[1,3,5..]*b[i]      (carry-in from the even[0]+odd[1] fold)
//     even =  even       + a[0,2,4..]*b[i]
//     m    = even[0] * (-p^-1 mod 2^32)
//     odd +=  p[1,3,5..]*m ;  even += p[0,2,4..]*m        => even[0] == 0
// and the roles of the two arrays swap for the next limb (that swap IS the divide by 2^32).
// N*(2N+1) multiply instructions per product (300 for N=12, 136 for N=8).
//
// What the reference does here: nothing -- the arithmetic of /root/reference lives in an FPGA
// bitstream; the semantics (arkworks Fp Montgomery arithmetic) are restated by oracle/.
#pragma once
#include "bz_common.cuh"
#include "field_constants.h"

namespace bz {

template <class F>
struct Fe {
  uint32_t v[F::N];
};

template <class F>
struct ff {
  static constexpr int N = F::N;
  typedef Fe<F> E;

  BZ_HDI static E zero() {
    E r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  BZ_HDI static E one() {
    E r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = F::one()[i];
    return r;
  }
  BZ_HDI static bool is_zero(const E& a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= a.v[i];
    return t == 0;
  }
  BZ_HDI static bool eq(const E& a, const E& b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
  }

  // r = a - p if a >= p else a      (a < 2p)
  BZ_HDI static void final_sub(uint32_t* a) {
    uint32_t t[N];
    t[0] = cc::sub_cc(a[0], F::mod()[0]);
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = cc::subc_cc(a[i], F::mod()[i]);
    uint32_t borrow = cc::subc(0u, 0u);   // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
  }

  BZ_HDI static E add(const E& a, const E& b) {
    E r;
    r.v[0] = cc::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = cc::addc_cc(a.v[i], b.v[i]);
    r.v[N - 1] = cc::addc(a.v[N - 1], b.v[N - 1]);   // 2p < 2^(32N): no carry out
    final_sub(r.v);
    return r;
  }
  BZ_HDI static E dbl(const E& a) { return add(a, a); }

  BZ_HDI static E sub(const E& a, const E& b) {
    E r;
    r.v[0] = cc::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = cc::subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = cc::subc(0u, 0u);   // 0 or 0xffffffff
    r.v[0] = cc::add_cc(r.v[0], F::mod()[0] & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = cc::addc_cc(r.v[i], F::mod()[i] & borrow);
    r.v[N - 1] = cc::addc(r.v[N - 1], F::mod()[N - 1] & borrow);
    return r;
  }
  BZ_HDI static E neg(const E& a) { return sub(zero(), a); }

  // ---- Montgomery product building blocks (see header) ----
  // acc[0..N) = sum_{j even} a[j]*bi * 2^(32 j)            (fresh, no carries needed)
  BZ_HDI static void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      acc[j] = cc::mul_lo(a[j], bi);
      acc[j + 1] = cc::mul_hi(a[j], bi);
    }
  }
  // acc[0..N) += sum_{j even} a[j]*bi * 2^(32 j); carry-out left in CC
  BZ_HDI static void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    acc[0] = cc::mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = cc::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      acc[j] = cc::madc_lo_cc(a[j], bi, acc[j]);
      acc[j + 1] = cc::madc_hi_cc(a[j], bi, acc[j + 1]);
    }
  }
  // odd = (odd >> 64) + sum_{j even} a[j]*bi * 2^(32 j), carry-in from CC, no carry-out
  BZ_HDI static void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      odd[j] = cc::madc_lo_cc(a[j], bi, odd[j + 2]);
      odd[j + 1] = cc::madc_hi_cc(a[j], bi, odd[j + 3]);
    }
    odd[N - 2] = cc::madc_lo_cc(a[N - 2], bi, 0u);
    odd[N - 1] = cc::madc_hi(a[N - 2], bi, 0u);
  }
  // one multiplier limb: T = (T + a*bi + m*p) / 2^32 in the even/odd representation
  template <bool FIRST>
  BZ_HDI static void mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    if (FIRST) {
      mul_n(odd, a + 1, bi);
      mul_n(even, a, bi);
    } else {
      even[0] = cc::add_cc(even[0], odd[1]);
      madc_n_rshift(odd, a + 1, bi);
      cmad_n(even, a, bi);
      odd[N - 1] = cc::addc(odd[N - 1], 0u);
    }
    uint32_t mi = even[0] * F::INV;
    cmad_n(odd, F::mod() + 1, mi);
    cmad_n(even, F::mod(), mi);
    odd[N - 1] = cc::addc(odd[N - 1], 0u);
  }

  // r = a*b/R mod p
  BZ_HDI static E mul(const E& a, const E& b) {
    uint32_t even[N], odd[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      if (i == 0) mad_n_redc<true>(even, odd, a.v, b.v[0]);
      else        mad_n_redc<false>(even, odd, a.v, b.v[i]);
      mad_n_redc<false>(odd, even, a.v, b.v[i + 1]);
    }
    // T = even + (odd >> 32)
    E r;
    r.v[0] = cc::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = cc::addc_cc(even[i], odd[i + 1]);
    r.v[N - 1] = cc::addc(even[N - 1], 0u);
    final_sub(r.v);
    return r;
  }
  BZ_HDI static E sqr(const E& a) { return mul(a, a); }

  BZ_HDI static E to_mont(const E& a) {
    E r2;
#pragma unroll
    for (int i = 0; i < N; i++) r2.v[i] = F::r2()[i];
    return mul(a, r2);
  }
  BZ_HDI static E from_mont(const E& a) {
    E o = zero();
    o.v[0] = 1;
    return mul(a, o);
  }

  // a^(p-2): only used once per MSM (result normalisation) and in input generators
  BZ_HDI static E inv(const E& a) {
    uint32_t e[N];
    // e = p - 2   (p odd and > 2, so only limb 0 changes... unless limb 0 < 2)
    e[0] = cc::sub_cc(F::mod()[0], 2u);
#pragma unroll
    for (int i = 1; i < N; i++) e[i] = cc::subc_cc(F::mod()[i], 0u);
    E r = one();
    for (int i = N - 1; i >= 0; i--) {
      for (int bit = 31; bit >= 0; bit--) {
        r = sqr(r);
        if ((e[i] >> bit) & 1) r = mul(r, a);
      }
    }
    return r;
  }
};

}  // namespace bz
