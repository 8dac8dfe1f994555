// Prime-field arithmetic on 32-bit limbs, Montgomery form, for sm_100a.
//
// One thread owns one field element in registers (12 limbs for the 381/377-bit base
// fields, 8 limbs for the 254/255-bit fields).  All values are kept fully reduced in
// [0, p); "Montgomery form" means a*R mod p with R = 2^(32 N).
//
// mul() is a word-serial Montgomery product arranged as two interleaved sets of 64-bit
// accumulator columns ("even" = limb pairs (2k,2k+1), "odd" = pairs (2k+1,2k+2)) so that every
// 32x32->64 partial product plus its 64-bit accumulate-with-carry is ONE IMAD.WIDE.U32(.X)
// (ptxas fuses mul.wide.u32 + add(c).cc.u64).  Per multiplier limb b[i]:
//     odd  = (odd >> 64) + a[1,3,5..]*b[i]
//     even =  even + hi32(old odd[0]) + a[0,2,4..]*b[i]
//     m    = lo32(even[0]) * (-p^-1 mod 2^32)
//     odd +=  p[1,3,5..]*m ;  even += p[0,2,4..]*m        => lo32(even[0]) == 0
// and the roles of the two arrays swap for the next limb (that swap IS the divide by 2^32).
// About N*(2N+3) multiplier-pipe instructions per product.
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

  // m * p[K] for the two lowest modulus limbs: 1 and 2^32 - 1 (Fr of BLS12-381 has both, several fields the
  // first) need no multiplier instruction
  template <int K>
  BZ_HDI static uint64_t mod_times(uint32_t m) {
    constexpr uint32_t c = K == 0 ? F::MOD0 : F::MOD1;
    if constexpr (c == 1u) return (uint64_t)m;
    else if constexpr (c == 0xffffffffu) return cc::mul_2p32m1(m);
    else return cc::mul_wide(F::mod()[K], m);
  }

  // ---- Montgomery product (see header) ----
  // One multiplier limb.  E ("even") holds 64-bit columns (2k, 2k+1), O ("odd") columns
  // (2k+1, 2k+2):  T = E + 2^32 * O.  On entry of a non-first step O is the previous step's E,
  // whose low 32 bits are zero after the reduction, so  T_prev / 2^32 = E + hi32(O[0]) + 2^32 * (O >> 64).
  template <bool FIRST>
  BZ_HDI static void step(uint64_t* Ev, uint64_t* Ov, const uint32_t* a, uint32_t bi) {
    constexpr int NW = N / 2;
    if (FIRST) {
#pragma unroll
      for (int k = 0; k < NW; k++) {
        Ov[k] = cc::mul_wide(a[2 * k + 1], bi);
        Ev[k] = cc::mul_wide(a[2 * k], bi);
      }
    } else {
      uint64_t h = Ov[0] >> 32;
      // O' = (O >> 64) + a[odd] * bi
      Ov[0] = cc::add_cc64(Ov[1], cc::mul_wide(a[1], bi));
#pragma unroll
      for (int k = 1; k < NW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k + 1], cc::mul_wide(a[2 * k + 1], bi));
      Ov[NW - 1] = cc::addc64(0ull, cc::mul_wide(a[N - 1], bi));
      // E' = E + h + a[even] * bi
      Ev[0] = cc::add_cc64(Ev[0], cc::mad_wide(a[0], bi, h));
#pragma unroll
      for (int k = 1; k < NW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(a[2 * k], bi));
      Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);   // carry out of the even chain = bit 32 of the top odd word
    }
    uint32_t m = (uint32_t)Ev[0] * F::INV;
    Ov[0] = cc::add_cc64(Ov[0], mod_times<1>(m));
#pragma unroll
    for (int k = 1; k < NW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k], cc::mul_wide(F::mod()[2 * k + 1], m));
    Ov[NW - 1] = cc::addc64(Ov[NW - 1], cc::mul_wide(F::mod()[N - 1], m));
    Ev[0] = cc::add_cc64(Ev[0], mod_times<0>(m));
#pragma unroll
    for (int k = 1; k < NW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(F::mod()[2 * k], m));
    Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);
  }

  // T += x * y: one more row pair on the E/O columns (the same schedule as the m * p row of step()); the
  // carry out of the even chain lands in bit 32 of the top odd word
  BZ_HDI static void add_row(uint64_t* Ev, uint64_t* Ov, const uint32_t* x, uint32_t y) {
    constexpr int NW = N / 2;
    Ov[0] = cc::add_cc64(Ov[0], cc::mul_wide(x[1], y));
#pragma unroll
    for (int k = 1; k < NW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k], cc::mul_wide(x[2 * k + 1], y));
    Ov[NW - 1] = cc::addc64(Ov[NW - 1], cc::mul_wide(x[N - 1], y));
    Ev[0] = cc::add_cc64(Ev[0], cc::mul_wide(x[0], y));
#pragma unroll
    for (int k = 1; k < NW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(x[2 * k], y));
    Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);
  }
  // One step of the FUSED sum of two products a*b + c*d: both multiplier rows are accumulated before the
  // single reduction row, so the pair costs 3 N^2 instead of 4 N^2 multiplier instructions.
  // Bounds: after every division T < 3p + 1, before it T < 3p (1 + 2^32): fits the N+1 limb window when
  // p < 2^(32N-2); the final value is < p (2p/R + 1) < 2p, so one conditional subtraction normalises it.
  template <bool FIRST>
  BZ_HDI static void step2(uint64_t* Ev, uint64_t* Ov, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di) {
    static_assert(F::BITS + 2 <= 32 * N, "fused product sum needs two spare bits");
    constexpr int NW = N / 2;
    if (FIRST) {
#pragma unroll
      for (int k = 0; k < NW; k++) {
        Ov[k] = cc::mul_wide(a[2 * k + 1], bi);
        Ev[k] = cc::mul_wide(a[2 * k], bi);
      }
    } else {
      uint64_t h = Ov[0] >> 32;
      Ov[0] = cc::add_cc64(Ov[1], cc::mul_wide(a[1], bi));
#pragma unroll
      for (int k = 1; k < NW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k + 1], cc::mul_wide(a[2 * k + 1], bi));
      Ov[NW - 1] = cc::addc64(0ull, cc::mul_wide(a[N - 1], bi));
      Ev[0] = cc::add_cc64(Ev[0], cc::mad_wide(a[0], bi, h));
#pragma unroll
      for (int k = 1; k < NW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(a[2 * k], bi));
      Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);
    }
    add_row(Ev, Ov, c, di);
    uint32_t m = (uint32_t)Ev[0] * F::INV;
    add_row(Ev, Ov, F::mod(), m);
  }
  // r = (a*b + c*d)/R mod p with ONE Montgomery reduction
  BZ_HDI static E mul2_inline(const E& a, const E& b, const E& c, const E& d) {
    constexpr int NW = N / 2;
    uint64_t Ev[NW], Ov[NW];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      if (i == 0) step2<true>(Ev, Ov, a.v, b.v[0], c.v, d.v[0]);
      else        step2<false>(Ev, Ov, a.v, b.v[i], c.v, d.v[i]);
      step2<false>(Ov, Ev, a.v, b.v[i + 1], c.v, d.v[i + 1]);
    }
    E r;
    r.v[0] = cc::add_cc((uint32_t)Ev[0], (uint32_t)(Ov[0] >> 32));
#pragma unroll
    for (int j = 1; j < N - 1; j++) {
      uint32_t e = (j & 1) ? (uint32_t)(Ev[j / 2] >> 32) : (uint32_t)Ev[j / 2];
      uint32_t o = ((j + 1) & 1) ? (uint32_t)(Ov[(j + 1) / 2] >> 32) : (uint32_t)Ov[(j + 1) / 2];
      r.v[j] = cc::addc_cc(e, o);
    }
    r.v[N - 1] = cc::addc((uint32_t)(Ev[NW - 1] >> 32), 0u);
    final_sub(r.v);
    return r;
  }
  // a*b + c*d and a*b - c*d (fields with two spare bits only; the others fall back to two products)
  BZ_HDI static E mul2(const E& a, const E& b, const E& c, const E& d) {
    if constexpr (F::BITS + 2 > 32 * N) {
      return add(mul(a, b), mul(c, d));
    } else {
#if defined(BZ_NOINLINE_MUL) && defined(__CUDACC__)
      if constexpr (kSplit) return mul2_split(a, b, c, d);
      else return mul2_call(a, b, c, d);
#else
      return mul2_sel(a, b, c, d);
#endif
    }
  }
  BZ_HDI static E mul_sub2(const E& a, const E& b, const E& c, const E& d) {
#ifdef BZ_NO_FUSED_MUL2
    return sub(mul(a, b), mul(c, d));
#else
    return mul2(a, b, neg(c), d);
#endif
  }

  // ---- Karatsuba product + reduction-only Montgomery (BZ_KARATSUBA; fields with N % 4 == 0) -----------------
  // The interleaved product above spends N^2 multiplier instructions on a*b and N^2 + N on the reduction.
  // Splitting the two lets the product be one level of Karatsuba -- three H x H products, H = N/2, i.e.
  // 3/4 N^2 -- at the price of ~10 N plain adds, which go to the otherwise idle ALU pipe; the sum of two
  // products (mul2) shares one reduction as before.
  //
  // z[0..2H) = a[0..H) * b[0..H): schoolbook on the E/O columns, one output limb per row
  template <int H>
  BZ_HDI static void prod_half(const uint32_t* a, const uint32_t* b, uint32_t* z) {
    static_assert(H % 2 == 0, "half width must be even");
    constexpr int HW = H / 2;
    uint64_t X[HW], Y[HW];   // X starts in the even role; the roles swap every row (that is the shift by 32 bits)
#pragma unroll
    for (int i = 0; i < H; i++) {
      uint64_t* Ev = (i & 1) ? Y : X;
      uint64_t* Ov = (i & 1) ? X : Y;
      const uint32_t bi = b[i];
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < HW; k++) {
          Ov[k] = cc::mul_wide(a[2 * k + 1], bi);
          Ev[k] = cc::mul_wide(a[2 * k], bi);
        }
      } else {
        uint64_t h = Ov[0] >> 32;
        if (HW > 1) {
          Ov[0] = cc::add_cc64(Ov[1], cc::mul_wide(a[1], bi));
#pragma unroll
          for (int k = 1; k < HW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k + 1], cc::mul_wide(a[2 * k + 1], bi));
          Ov[HW - 1] = cc::addc64(0ull, cc::mul_wide(a[H - 1], bi));
        } else {
          Ov[0] = cc::mul_wide(a[1], bi);
        }
        Ev[0] = cc::add_cc64(Ev[0], cc::mad_wide(a[0], bi, h));
#pragma unroll
        for (int k = 1; k < HW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(a[2 * k], bi));
        Ov[HW - 1] = cc::addc_hi32(Ov[HW - 1]);
      }
      z[i] = (uint32_t)Ev[0];
    }
    // H is even: the last row had Y in the even role; what is left is (Y >> 32) + X  (H limbs, no carry out)
    z[H] = cc::add_cc((uint32_t)(Y[0] >> 32), (uint32_t)X[0]);
#pragma unroll
    for (int j = 1; j < H - 1; j++) {
      uint32_t y = ((j + 1) & 1) ? (uint32_t)(Y[(j + 1) / 2] >> 32) : (uint32_t)Y[(j + 1) / 2];
      uint32_t x = (j & 1) ? (uint32_t)(X[j / 2] >> 32) : (uint32_t)X[j / 2];
      z[H + j] = cc::addc_cc(y, x);
    }
    z[2 * H - 1] = cc::addc((uint32_t)(X[HW - 1] >> 32), 0u);
  }
  // T[0..2N) = a * b, one level of Karatsuba
  BZ_HDI static void prod_kara(const uint32_t* a, const uint32_t* b, uint32_t* T) {
    constexpr int H = N / 2;
    uint32_t sa[H], sb[H], z1[2 * H + 1];
    prod_half<H>(a, b, T);                   // z0
    prod_half<H>(a + H, b + H, T + 2 * H);   // z2
    sa[0] = cc::add_cc(a[0], a[H]);
#pragma unroll
    for (int i = 1; i < H; i++) sa[i] = cc::addc_cc(a[i], a[H + i]);
    const uint32_t ca = cc::addc(0u, 0u);
    sb[0] = cc::add_cc(b[0], b[H]);
#pragma unroll
    for (int i = 1; i < H; i++) sb[i] = cc::addc_cc(b[i], b[H + i]);
    const uint32_t cb = cc::addc(0u, 0u);
    prod_half<H>(sa, sb, z1);
    // (sa + ca 2^(32H)) (sb + cb 2^(32H)) = sa sb + (ca sb + cb sa) 2^(32H) + ca cb 2^(64H)
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    z1[H] = cc::add_cc(z1[H], sb[0] & ma);
#pragma unroll
    for (int i = 1; i < H; i++) z1[H + i] = cc::addc_cc(z1[H + i], sb[i] & ma);
    z1[2 * H] = cc::addc(ca & cb, 0u);
    z1[H] = cc::add_cc(z1[H], sa[0] & mb);
#pragma unroll
    for (int i = 1; i < H; i++) z1[H + i] = cc::addc_cc(z1[H + i], sa[i] & mb);
    z1[2 * H] = cc::addc(z1[2 * H], 0u);
    // z1 -= z0 + z2  (the result is the non-negative middle term)
    z1[0] = cc::sub_cc(z1[0], T[0]);
#pragma unroll
    for (int i = 1; i < 2 * H; i++) z1[i] = cc::subc_cc(z1[i], T[i]);
    z1[2 * H] = cc::subc(z1[2 * H], 0u);
    z1[0] = cc::sub_cc(z1[0], T[2 * H]);
#pragma unroll
    for (int i = 1; i < 2 * H; i++) z1[i] = cc::subc_cc(z1[i], T[2 * H + i]);
    z1[2 * H] = cc::subc(z1[2 * H], 0u);
    // T += z1 << (32 H)
    T[H] = cc::add_cc(T[H], z1[0]);
#pragma unroll
    for (int i = 1; i <= 2 * H; i++) T[H + i] = cc::addc_cc(T[H + i], z1[i]);
#pragma unroll
    for (int i = 3 * H + 1; i < 4 * H - 1; i++) T[i] = cc::addc_cc(T[i], 0u);
    T[4 * H - 1] = cc::addc(T[4 * H - 1], 0u);
  }
  // r = T / R mod p for T < 2 p^2 (2N limbs): N reduction rows on the low half, then the high half is added
  BZ_HDI static E redc(const uint32_t* T) {
    constexpr int NW = N / 2;
    uint64_t X[NW], Y[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) { X[k] = (uint64_t)T[2 * k] | ((uint64_t)T[2 * k + 1] << 32); Y[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint64_t* Ev = (i & 1) ? Y : X;
      uint64_t* Ov = (i & 1) ? X : Y;
      uint64_t h = 0;
      if (i != 0) {
        h = Ov[0] >> 32;
#pragma unroll
        for (int k = 0; k < NW - 1; k++) Ov[k] = Ov[k + 1];
        Ov[NW - 1] = 0;
      }
      const uint32_t m = ((uint32_t)Ev[0] + (uint32_t)h) * F::INV;
      Ov[0] = cc::add_cc64(Ov[0], mod_times<1>(m));
#pragma unroll
      for (int k = 1; k < NW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k], cc::mul_wide(F::mod()[2 * k + 1], m));
      Ov[NW - 1] = cc::addc64(Ov[NW - 1], cc::mul_wide(F::mod()[N - 1], m));
      Ev[0] = cc::add_cc64(Ev[0], mod_times<0>(m) + h);   // m p0 + h < 2^64
#pragma unroll
      for (int k = 1; k < NW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(F::mod()[2 * k], m));
      Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);
    }
    // N is even: the last row had Y in the even role: (T_lo + M p) / R = X + (Y >> 32); then + T_hi
    E r;
    r.v[0] = cc::add_cc((uint32_t)X[0], (uint32_t)(Y[0] >> 32));
#pragma unroll
    for (int j = 1; j < N - 1; j++) {
      uint32_t e = (j & 1) ? (uint32_t)(X[j / 2] >> 32) : (uint32_t)X[j / 2];
      uint32_t o = ((j + 1) & 1) ? (uint32_t)(Y[(j + 1) / 2] >> 32) : (uint32_t)Y[(j + 1) / 2];
      r.v[j] = cc::addc_cc(e, o);
    }
    r.v[N - 1] = cc::addc((uint32_t)(X[NW - 1] >> 32), 0u);
    r.v[0] = cc::add_cc(r.v[0], T[N]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r.v[j] = cc::addc_cc(r.v[j], T[N + j]);
    r.v[N - 1] = cc::addc(r.v[N - 1], T[2 * N - 1]);
    final_sub(r.v);
    return r;
  }
  BZ_HDI static E mul_kara(const E& a, const E& b) {
    uint32_t T[2 * N];
    prod_kara(a.v, b.v, T);
    return redc(T);
  }
  BZ_HDI static E mul2_kara(const E& a, const E& b, const E& c, const E& d) {
    static_assert(F::BITS + 2 <= 32 * N, "sum of two products needs two spare bits");
    uint32_t T[2 * N], U[2 * N];
    prod_kara(a.v, b.v, T);
    prod_kara(c.v, d.v, U);
    T[0] = cc::add_cc(T[0], U[0]);
#pragma unroll
    for (int i = 1; i < 2 * N - 1; i++) T[i] = cc::addc_cc(T[i], U[i]);
    T[2 * N - 1] = cc::addc(T[2 * N - 1], U[2 * N - 1]);
    return redc(T);
  }

  // T[0..2N) = a^2 without reduction: N(N+1)/2 partial products.  Row i adds a_i * V_i with V_i as in the dedicated
  // squaring below (diagonal term once, the limbs above i doubled through d = a << 1); limb i of T is final after
  // row i.  Needs 2a < 2^(32N) (one spare bit: every field here).
  template <int I>
  BZ_HDI static void sqr_wide_rows(uint64_t* Ev, uint64_t* Ov, const uint32_t* a, const uint32_t* d, uint32_t* T) {
    if constexpr (I < N) {
      constexpr int NW = N / 2;
      const uint32_t bi = a[I];
      if constexpr (I == 0) {
#pragma unroll
        for (int k = 0; k < NW; k++) {
          Ov[k] = cc::mul_wide(sq_limb(a, d, 0, 2 * k + 1), bi);
          Ev[k] = cc::mul_wide(sq_limb(a, d, 0, 2 * k), bi);
        }
      } else {
        uint64_t h = Ov[0] >> 32;
        constexpr int K0 = I / 2;          // first odd-position product of this row: column 2 K0 + 1 >= I
        constexpr int KE = (I + 1) / 2;    // first even-position product: column 2 KE >= I
#pragma unroll
        for (int k = 0; k < K0 && k < NW - 1; k++) Ov[k] = Ov[k + 1];
        if constexpr (K0 < NW - 1) {
          Ov[K0] = cc::add_cc64(Ov[K0 + 1], cc::mul_wide(sq_limb(a, d, I, 2 * K0 + 1), bi));
#pragma unroll
          for (int k = K0 + 1; k < NW - 1; k++)
            Ov[k] = cc::addc_cc64(Ov[k + 1], cc::mul_wide(sq_limb(a, d, I, 2 * k + 1), bi));
          Ov[NW - 1] = cc::addc64(0ull, cc::mul_wide(sq_limb(a, d, I, N - 1), bi));
        } else {
          Ov[NW - 1] = cc::mul_wide(sq_limb(a, d, I, N - 1), bi);
        }
        Ev[0] = cc::add_cc64(Ev[0], h);
#pragma unroll
        for (int k = 1; k < NW; k++)
          Ev[k] = cc::addc_cc64(Ev[k], k >= KE ? cc::mul_wide(sq_limb(a, d, I, 2 * k), bi) : 0ull);
        Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);
      }
      T[I] = (uint32_t)Ev[0];
      sqr_wide_rows<I + 1>(Ov, Ev, a, d, T);
    }
  }
  BZ_HDI static void sqr_wide(const uint32_t* a, uint32_t* T) {
    static_assert(F::BITS + 1 <= 32 * N, "doubled operand needs one spare bit");
    constexpr int NW = N / 2;
    uint64_t X[NW], Y[NW];
    uint32_t d[N];
    d[0] = a[0] << 1;
#pragma unroll
    for (int j = 1; j < N; j++) d[j] = (a[j] << 1) | (a[j - 1] >> 31);
    sqr_wide_rows<0>(X, Y, a, d, T);
    // N is even: the last row had Y in the even role; what is left is (Y >> 32) + X
    T[N] = cc::add_cc((uint32_t)(Y[0] >> 32), (uint32_t)X[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) {
      uint32_t y = ((j + 1) & 1) ? (uint32_t)(Y[(j + 1) / 2] >> 32) : (uint32_t)Y[(j + 1) / 2];
      uint32_t x = (j & 1) ? (uint32_t)(X[j / 2] >> 32) : (uint32_t)X[j / 2];
      T[N + j] = cc::addc_cc(y, x);
    }
    T[2 * N - 1] = cc::addc((uint32_t)(X[NW - 1] >> 32), 0u);
  }
  BZ_HDI static E sqr_split(const E& a) {
    uint32_t T[2 * N];
    sqr_wide(a.v, T);
    return redc(T);
  }

  // ---- BZ_SPLIT_MUL: product, squaring product and reduction as THREE shared routines ---------------------------
  // mul = redc(prod), sqr = redc(sqr_wide), a*b + c*d = redc(prod + prod): the 2N-limb intermediate travels in
  // registers through the call ABI (checked in SASS: no local memory).  The Karatsuba product above then costs
  // 3/4 N^2 multiplier instructions where the interleaved routines spend N^2, and -- unlike BZ_KARATSUBA, which
  // inlines product + reduction into each of mul / sqr / mul2 (31 KB) -- the three bodies together are ~11 KB:
  // the accumulation loop stays inside the 32 KB instruction cache.
  struct Wide {
    uint32_t v[2 * N];
  };
#ifdef __CUDACC__
  static __device__ __noinline__ Wide prodw_call(const E a, const E b) {
    Wide t;
    prod_kara(a.v, b.v, t.v);
    return t;
  }
  static __device__ __noinline__ Wide sqrw_call(const E a) {
    Wide t;
    sqr_wide(a.v, t.v);
    return t;
  }
  static __device__ __noinline__ E redc_call(const Wide t) { return redc(t.v); }
#else
  static Wide prodw_call(const E& a, const E& b) {
    Wide t;
    prod_kara(a.v, b.v, t.v);
    return t;
  }
  static Wide sqrw_call(const E& a) {
    Wide t;
    sqr_wide(a.v, t.v);
    return t;
  }
  static E redc_call(const Wide& t) { return redc(t.v); }
#endif
#if defined(__CUDACC__) && defined(BZ_SPLIT_NESTED)
  // nested form: the product is inlined into the routine the caller sees, only the reduction is shared
  static __device__ __noinline__ E mul_nested(const E a, const E b) {
    Wide t;
    prod_kara(a.v, b.v, t.v);
    return redc_call(t);
  }
  static __device__ __noinline__ E sqr_nested(const E a) {
    Wide t;
    sqr_wide(a.v, t.v);
    return redc_call(t);
  }
  BZ_HDI static E mul_split(const E& a, const E& b) { return mul_nested(a, b); }
  BZ_HDI static E sqr_splitc(const E& a) { return sqr_nested(a); }
#else
  BZ_HDI static E mul_split(const E& a, const E& b) { return redc_call(prodw_call(a, b)); }
  BZ_HDI static E sqr_splitc(const E& a) { return redc_call(sqrw_call(a)); }
#endif
  BZ_HDI static E mul2_split(const E& a, const E& b, const E& c, const E& d) {
    static_assert(F::BITS + 2 <= 32 * N, "sum of two products needs two spare bits");
    Wide t = prodw_call(a, b);
    const Wide u = prodw_call(c, d);
    t.v[0] = cc::add_cc(t.v[0], u.v[0]);
#pragma unroll
    for (int i = 1; i < 2 * N - 1; i++) t.v[i] = cc::addc_cc(t.v[i], u.v[i]);
    t.v[2 * N - 1] = cc::addc(t.v[2 * N - 1], u.v[2 * N - 1]);
    return redc_call(t);
  }
#if defined(BZ_SPLIT_MUL) && defined(__CUDACC__)
  static constexpr bool kSplit = (N % 4 == 0) && (N >= BZ_SPLIT_MUL);
#else
  static constexpr bool kSplit = false;
#endif

  // r = a*b/R mod p.  With BZ_NOINLINE_MUL the product and the square are real function calls (operands
  // and result travel in registers, ~35 MOVs per call): the unrolled bodies are 400 / 330 instructions, so
  // a mixed add with ten of them inlined is ~62 KB of code per loop iteration -- twice the SM's 32 KB
  // instruction cache (ncu: no_instruction stalls) -- while the called form keeps the loop under 24 KB.
  BZ_HDI static E mul(const E& a, const E& b) {
#if defined(BZ_NOINLINE_MUL) && defined(__CUDACC__)
    if constexpr (kSplit) return mul_split(a, b);
    else return mul_call(a, b);
#else
    return mul_sel(a, b);
#endif
  }
  BZ_HDI static E sqr(const E& a) {
#if defined(BZ_NOINLINE_MUL) && defined(__CUDACC__)
    if constexpr (kSplit) return sqr_splitc(a);
    else return sqr_call(a);
#else
    return sqr_inline(a);
#endif
  }
#ifdef __CUDACC__
  static __device__ __noinline__ E mul_call(const E a, const E b) { return mul_sel(a, b); }
  static __device__ __noinline__ E sqr_call(const E a) { return sqr_inline(a); }
  static __device__ __noinline__ E mul2_call(const E a, const E b, const E c, const E d) { return mul2_sel(a, b, c, d); }
  // out-of-line add / sub / inverse for code whose loops would otherwise outgrow the instruction cache (msm_ba2.cuh)
  static __device__ __noinline__ E sub_call(const E a, const E b) { return sub(a, b); }
  static __device__ __noinline__ E add_call(const E a, const E b) { return add(a, b); }
  static __device__ __noinline__ E inv_gcd_call(const E a) { return inv_gcd(a); }
#else
  static E sub_call(const E& a, const E& b) { return sub(a, b); }
  static E add_call(const E& a, const E& b) { return add(a, b); }
  static E inv_gcd_call(const E& a) { return inv_gcd(a); }
#endif
  BZ_HDI static E mul_sel(const E& a, const E& b) {
#ifdef BZ_KARATSUBA
    if constexpr (N % 4 == 0 && N >= BZ_KARATSUBA) return mul_kara(a, b);
    else return mul_inline(a, b);
#else
    return mul_inline(a, b);
#endif
  }
  BZ_HDI static E mul2_sel(const E& a, const E& b, const E& c, const E& d) {
#ifdef BZ_KARATSUBA
    if constexpr (N % 4 == 0 && N >= BZ_KARATSUBA) return mul2_kara(a, b, c, d);
    else return mul2_inline(a, b, c, d);
#else
    return mul2_inline(a, b, c, d);
#endif
  }
  BZ_HDI static E mul_inline(const E& a, const E& b) {
    constexpr int NW = N / 2;
    uint64_t Ev[NW], Ov[NW];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      if (i == 0) step<true>(Ev, Ov, a.v, b.v[0]);
      else        step<false>(Ev, Ov, a.v, b.v[i]);
      step<false>(Ov, Ev, a.v, b.v[i + 1]);
    }
    // last step had Ov in the "even" role: T = Ev + (Ov >> 32)
    E r;
    r.v[0] = cc::add_cc((uint32_t)Ev[0], (uint32_t)(Ov[0] >> 32));
#pragma unroll
    for (int j = 1; j < N - 1; j++) {
      uint32_t e = (j & 1) ? (uint32_t)(Ev[j / 2] >> 32) : (uint32_t)Ev[j / 2];
      uint32_t o = ((j + 1) & 1) ? (uint32_t)(Ov[(j + 1) / 2] >> 32) : (uint32_t)Ov[(j + 1) / 2];
      r.v[j] = cc::addc_cc(e, o);
    }
    r.v[N - 1] = cc::addc((uint32_t)(Ev[NW - 1] >> 32), 0u);
    final_sub(r.v);
    return r;
  }
  // ---- dedicated squaring: N(N+1)/2 instead of N^2 partial products in the multiplication half.
  // Row i adds a_i * V_i with V_i = a_i 2^(32 i) + 2 * (limbs above i of a) 2^(32(i+1)) -- the diagonal
  // term once and every off-diagonal product doubled, so rows need no products below column i.  The
  // doubled tail comes from d = a << 1 (limbwise funnel shift); its lowest limb drops the bit that
  // would have carried in from a_i.  Doubling front-loads the partial sums: before the division of step
  // i the accumulator is < 2^(32N) + 3 * 2^32 p, which fits the N+1-limb window only if p < 2^(32N-2).
  // That holds for the three base fields and for Fr(BN254) / Fr(BLS12-377); Fr(BLS12-381) (255 bits in
  // 256) keeps the generic product.
  BZ_HDI static uint32_t sq_limb(const uint32_t* a, const uint32_t* d, int i, int j) {
    return j == i ? a[i] : (j == i + 1 ? (d[j] & 0xfffffffeu) : d[j]);
  }
  template <int I>
  BZ_HDI static void step_sqr(uint64_t* Ev, uint64_t* Ov, const uint32_t* a, const uint32_t* d) {
    static_assert(F::BITS + 2 <= 32 * N, "front-loaded doubling needs two spare bits");
    constexpr int NW = N / 2;
    const uint32_t bi = a[I];
    if (I == 0) {
#pragma unroll
      for (int k = 0; k < NW; k++) {
        Ov[k] = cc::mul_wide(sq_limb(a, d, 0, 2 * k + 1), bi);
        Ev[k] = cc::mul_wide(sq_limb(a, d, 0, 2 * k), bi);
      }
    } else {
      uint64_t h = Ov[0] >> 32;
      constexpr int K0 = I / 2;          // first odd-position product of this row: column 2 K0 + 1 >= I
      constexpr int KE = (I + 1) / 2;    // first even-position product: column 2 KE >= I
#pragma unroll
      for (int k = 0; k < K0 && k < NW - 1; k++) Ov[k] = Ov[k + 1];
      if (K0 < NW - 1) {
        Ov[K0] = cc::add_cc64(Ov[K0 + 1], cc::mul_wide(sq_limb(a, d, I, 2 * K0 + 1), bi));
#pragma unroll
        for (int k = K0 + 1; k < NW - 1; k++)
          Ov[k] = cc::addc_cc64(Ov[k + 1], cc::mul_wide(sq_limb(a, d, I, 2 * k + 1), bi));
        Ov[NW - 1] = cc::addc64(0ull, cc::mul_wide(sq_limb(a, d, I, N - 1), bi));
      } else {
        Ov[NW - 1] = cc::mul_wide(sq_limb(a, d, I, N - 1), bi);
      }
      Ev[0] = cc::add_cc64(Ev[0], h);
#pragma unroll
      for (int k = 1; k < NW; k++)
        Ev[k] = cc::addc_cc64(Ev[k], k >= KE ? cc::mul_wide(sq_limb(a, d, I, 2 * k), bi) : 0ull);
      Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);   // carry out of the even chain = bit 32 of the top odd word
    }
    uint32_t m = (uint32_t)Ev[0] * F::INV;
    Ov[0] = cc::add_cc64(Ov[0], mod_times<1>(m));
#pragma unroll
    for (int k = 1; k < NW - 1; k++) Ov[k] = cc::addc_cc64(Ov[k], cc::mul_wide(F::mod()[2 * k + 1], m));
    Ov[NW - 1] = cc::addc64(Ov[NW - 1], cc::mul_wide(F::mod()[N - 1], m));
    Ev[0] = cc::add_cc64(Ev[0], mod_times<0>(m));
#pragma unroll
    for (int k = 1; k < NW; k++) Ev[k] = cc::addc_cc64(Ev[k], cc::mul_wide(F::mod()[2 * k], m));
    Ov[NW - 1] = cc::addc_hi32(Ov[NW - 1]);
  }
  template <int I>
  BZ_HDI static void sqr_rows(uint64_t* Ev, uint64_t* Ov, const uint32_t* a, const uint32_t* d) {
    if constexpr (I < N) {
      step_sqr<I>(Ev, Ov, a, d);
      step_sqr<I + 1>(Ov, Ev, a, d);
      sqr_rows<I + 2>(Ev, Ov, a, d);
    }
  }
  BZ_HDI static E sqr_inline(const E& a) {
    if constexpr (F::BITS + 2 > 32 * N) {
      return mul_inline(a, a);
    } else {
      return sqr_dedicated(a);
    }
  }
  BZ_HDI static E sqr_dedicated(const E& a) {
    constexpr int NW = N / 2;
    uint64_t Ev[NW], Ov[NW];
    uint32_t d[N];
    d[0] = a.v[0] << 1;
#pragma unroll
    for (int j = 1; j < N; j++) d[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
    sqr_rows<0>(Ev, Ov, a.v, d);
    E r;
    r.v[0] = cc::add_cc((uint32_t)Ev[0], (uint32_t)(Ov[0] >> 32));
#pragma unroll
    for (int j = 1; j < N - 1; j++) {
      uint32_t e = (j & 1) ? (uint32_t)(Ev[j / 2] >> 32) : (uint32_t)Ev[j / 2];
      uint32_t o = ((j + 1) & 1) ? (uint32_t)(Ov[(j + 1) / 2] >> 32) : (uint32_t)Ov[(j + 1) / 2];
      r.v[j] = cc::addc_cc(e, o);
    }
    r.v[N - 1] = cc::addc((uint32_t)(Ev[NW - 1] >> 32), 0u);
    final_sub(r.v);
    return r;
  }

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

  // ---- inversion by division steps (Bernstein & Yang, "Fast constant-time gcd computation and modular inversion",
  // CHES 2019), on signed 30-bit limbs.  The Fermat inverse below costs ~1.5 N x 32 field products, all on the integer
  // multiplier pipe; this one costs ~(26 + ...)  batches of 30 division steps, each batch being ~400 plain ALU
  // instructions plus two small matrix-vector updates (~10 N multiplier instructions) -- about 25x fewer multiplier
  // instructions.  That is what makes ONE inversion per thread and batch affordable in the batched-affine bucket
  // accumulation (msm_ba2.cuh): no cross-thread product trees, no extra kernels.
  //
  //   divstep(delta, f, g) = (1 - delta, g, (g - f) / 2)              if delta > 0 and g odd
  //                          (1 + delta, f, (g + (g mod 2) f) / 2)    otherwise
  // starting from (1, p, x) reaches g = 0 with f = +-1 after at most (49 bits + 57) / 17 steps.  30 steps at a time
  // are taken on the low words only and collected in a 2x2 matrix T (entries |.| <= 2^30) with
  // (f, g) <- T (f, g) / 2^30; the same matrix keeps (d, e) with  f = d x,  g = e x  (mod p), the division by 2^30
  // being made exact by adding the right multiple of p.  At the end x^-1 = f d.
  static constexpr int NL30 = (32 * N + 2 + 29) / 30;   // 30-bit limbs for values in (-2p, 2p): 9 (256 bits), 13 (384)
  static constexpr int GCD_MAX_BATCHES = (49 * F::BITS + 57) / 17 / 30 + 1;

  BZ_HDI static void to_signed30(const uint32_t* a, int32_t* o) {
#pragma unroll
    for (int j = 0; j < NL30; j++) {
      const int lo = 30 * j, w = lo >> 5, sh = lo & 31;
      uint32_t v = w < N ? (a[w] >> sh) : 0u;
      if (sh > 2 && w + 1 < N) v |= a[w + 1] << (32 - sh);
      o[j] = (int32_t)(v & 0x3fffffffu);
    }
  }

  // 30 division steps on the low words; returns the new delta, T = (u v; q r)
  BZ_HDI static int32_t divsteps30(int32_t delta, uint32_t f, uint32_t g, int32_t& u, int32_t& v, int32_t& q, int32_t& r) {
    uint32_t uu = 1, vv = 0, qq = 0, rr = 1;
#pragma unroll 1
    for (int i = 0; i < 30; i++) {
      const uint32_t odd = 0u - (g & 1u);                                  // all ones when g is odd
      const uint32_t sw = odd & (uint32_t)((int32_t)(0 - delta) >> 31);    // ... and delta > 0: swap
      // swap: (f, g) <- (g, -f), rows likewise, delta <- -delta
      const uint32_t nf = (f ^ sw) - sw, nu = (uu ^ sw) - sw, nv = (vv ^ sw) - sw;   // -f, -u, -v when swapping
      const uint32_t tf = sw ? g : f, tu = sw ? qq : uu, tv = sw ? rr : vv;
      const uint32_t tg = sw ? nf : g, tq = sw ? nu : qq, tr = sw ? nv : rr;
      delta = (int32_t)(((uint32_t)delta ^ sw) - sw) + 1;
      // g <- (g + odd f) / 2, row_g += odd row_f; the f row is doubled instead of halving the g row
      g = (tg + (tf & odd)) >> 1;
      qq = tq + (tu & odd);
      rr = tr + (tv & odd);
      f = tf;
      uu = tu << 1;
      vv = tv << 1;
    }
    u = (int32_t)uu; v = (int32_t)vv; q = (int32_t)qq; r = (int32_t)rr;
    return delta;
  }

  BZ_HDI static E inv_gcd(const E& a) {
    constexpr int L = NL30;
    constexpr int32_t M30 = 0x3fffffff;
    constexpr uint32_t PINV30 = (0u - F::INV) & 0x3fffffffu;   // p^-1 mod 2^30
    int32_t f[L], g[L], d[L], e[L], pm[L];
    {
      uint32_t pw[N];
#pragma unroll
      for (int i = 0; i < N; i++) pw[i] = F::mod()[i];
      to_signed30(pw, pm);
    }
    to_signed30(a.v, g);
#pragma unroll
    for (int i = 0; i < L; i++) { f[i] = pm[i]; d[i] = 0; e[i] = 0; }
    e[0] = 1;
    int32_t delta = 1;
#pragma unroll 1
    for (int it = 0; it < GCD_MAX_BATCHES; it++) {
      int32_t gz = 0;
#pragma unroll
      for (int i = 0; i < L; i++) gz |= g[i];
      if (gz == 0) break;
      int32_t u, v, q, r;
      delta = divsteps30(delta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), u, v, q, r);
      // (d, e) <- T (d, e) / 2^30 mod p, kept in (-2p, p)
      {
        const int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
        int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
        int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
        int64_t ce = (int64_t)q * d[0] + (int64_t)r * e[0];
        md -= (int32_t)((PINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
        me -= (int32_t)((PINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
        cd += (int64_t)md * pm[0];
        ce += (int64_t)me * pm[0];
        cd >>= 30;
        ce >>= 30;
#pragma unroll
        for (int i = 1; i < L; i++) {
          cd += (int64_t)u * d[i] + (int64_t)v * e[i] + (int64_t)md * pm[i];
          ce += (int64_t)q * d[i] + (int64_t)r * e[i] + (int64_t)me * pm[i];
          d[i - 1] = (int32_t)cd & M30;
          e[i - 1] = (int32_t)ce & M30;
          cd >>= 30;
          ce >>= 30;
        }
        d[L - 1] = (int32_t)cd;
        e[L - 1] = (int32_t)ce;
      }
      // (f, g) <- T (f, g) / 2^30 (exact)
      {
        int64_t cf = (int64_t)u * f[0] + (int64_t)v * g[0];
        int64_t cg = (int64_t)q * f[0] + (int64_t)r * g[0];
        cf >>= 30;
        cg >>= 30;
#pragma unroll
        for (int i = 1; i < L; i++) {
          cf += (int64_t)u * f[i] + (int64_t)v * g[i];
          cg += (int64_t)q * f[i] + (int64_t)r * g[i];
          f[i - 1] = (int32_t)cf & M30;
          g[i - 1] = (int32_t)cg & M30;
          cf >>= 30;
          cg >>= 30;
        }
        f[L - 1] = (int32_t)cf;
        g[L - 1] = (int32_t)cg;
      }
    }
    // x^-1 = f d with f = +-1; bring it to [0, 2p) on the 30-bit limbs, then to 32-bit words
    const int32_t fneg = f[L - 1] >> 31;
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {   // d <- -d when f = -1
      int32_t t = ((d[i] ^ fneg) - fneg) + carry;
      if (i < L - 1) { carry = t >> 30; t &= M30; }
      d[i] = t;
    }
    const int32_t dneg = d[L - 1] >> 31;   // d in (-2p, 2p): add p once when negative ...
    carry = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
      int32_t t = d[i] + (pm[i] & dneg) + carry;
      if (i < L - 1) { carry = t >> 30; t &= M30; }
      d[i] = t;
    }
    const int32_t dneg2 = d[L - 1] >> 31;  // ... and once more (d was below -p)
    carry = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
      int32_t t = d[i] + (pm[i] & dneg2) + carry;
      if (i < L - 1) { carry = t >> 30; t &= M30; }
      d[i] = t;
    }
    E y;
#pragma unroll
    for (int k = 0; k < N; k++) {   // word k = bits [32k, 32k + 32)
      const int lo = 32 * k, j = lo / 30, sh = lo - 30 * j;
      uint32_t w = (uint32_t)d[j] >> sh;
      if (j + 1 < L) w |= (uint32_t)d[j + 1] << (30 - sh);
      if (30 - sh + 30 < 32 && j + 2 < L) w |= (uint32_t)d[j + 2] << (60 - sh);
      y.v[k] = w;
    }
    final_sub(y.v);
    // y = (a R)^-1 as an integer = a^-1 R^-1: two Montgomery products by R^2 give a^-1 R
    E r2;
#pragma unroll
    for (int i = 0; i < N; i++) r2.v[i] = F::r2()[i];
    return mul(mul(y, r2), r2);
  }

  // a^-1 (0 -> 0).  Division steps (inv_gcd) everywhere: the window-table build, the result normalisation of every
  // task and the multi-GPU combine all sit on latency-critical single-thread paths where the Fermat ladder's ~1.5 * 32 N
  // dependent field products cost ~0.5 ms.
  BZ_HDI static E inv(const E& a) {
#if defined(__CUDACC__)
    return inv_gcd_call(a);
#else
    return inv_gcd(a);
#endif
  }
  // a^(p-2): the textbook inverse, kept as the independent check of inv_gcd (tests/test_device_math_on_host.py)
  BZ_HDI static E inv_fermat(const E& a) {
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
