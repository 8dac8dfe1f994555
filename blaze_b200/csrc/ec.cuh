// Short-Weierstrass (a = 0) group law in extended-Jacobian ("XYZZ") coordinates.
//
//   affine (x, y)          <->  XYZZ (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2
//   infinity               <->  ZZ == 0
//   affine (0, 0)          is treated as the identity (padding record; not on any of the
//                          three curves since b != 0)
//
// Formulas: EFD "xyzz" for short Weierstrass curves -- madd-2008-s (8M+2S), add-2008-s
// (12M+2S), dbl-2008-s-1 (6M+4S+...); a = 0 for BLS12-377, BLS12-381 and BN254.
// All three operations are COMPLETE here: equal inputs fall through to doubling, opposite
// inputs give infinity.  The reference's own test vectors need this: tiling a 256-element
// block (tests/msm/mod.rs:92-109 of the reference) puts thousands of copies of one point in
// a single bucket.
#pragma once
#include "ff.cuh"

namespace bz {

template <class C>
struct Affine {
  Fe<typename C::Fq> x, y;   // Montgomery form
};

template <class C>
struct XYZZ {
  Fe<typename C::Fq> X, Y, ZZ, ZZZ;   // Montgomery form
};

template <class C>
struct ec {
  typedef typename C::Fq Fq;
  typedef ff<Fq> F;
  typedef Fe<Fq> E;
  typedef Affine<C> A;
  typedef XYZZ<C> P;

  BZ_HDI static P infinity() {
    P r;
    r.X = F::zero(); r.Y = F::zero(); r.ZZ = F::zero(); r.ZZZ = F::zero();
    return r;
  }
  BZ_HDI static bool is_inf(const P& p) { return F::is_zero(p.ZZ); }
  BZ_HDI static bool is_identity(const A& a) { return F::is_zero(a.x) && F::is_zero(a.y); }

  BZ_HDI static P from_affine(const A& a) {
    if (is_identity(a)) return infinity();
    P r;
    r.X = a.x; r.Y = a.y; r.ZZ = F::one(); r.ZZZ = F::one();
    return r;
  }

  // 2*(x, y) for an affine point (mdbl-2008-s-1)
  BZ_HDI static P dbl_affine(const A& a) {
    P r;
    E U = F::dbl(a.y);
    E V = F::sqr(U);
    E W = F::mul(U, V);
    E S = F::mul(a.x, V);
    E x2 = F::sqr(a.x);
    E M = F::add(F::dbl(x2), x2);
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::mul_sub2(M, F::sub(S, r.X), W, a.y);
    r.ZZ = V;
    r.ZZZ = W;
    return r;   // y == 0 gives ZZ = 0 = infinity (2-torsion), consistent
  }

  BZ_HDI static P dbl(const P& p) {
    if (is_inf(p)) return p;
    P r;
    E U = F::dbl(p.Y);
    E V = F::sqr(U);
    E W = F::mul(U, V);
    E S = F::mul(p.X, V);
    E x2 = F::sqr(p.X);
    E M = F::add(F::dbl(x2), x2);
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::mul_sub2(M, F::sub(S, r.X), W, p.Y);
    r.ZZ = F::mul(V, p.ZZ);
    r.ZZZ = F::mul(W, p.ZZZ);
    return r;
  }

  // acc += a   (mixed addition, complete)
  BZ_HDI static void madd(P& acc, const A& a) {
    if (is_identity(a)) return;
    if (is_inf(acc)) {
      acc.X = a.x; acc.Y = a.y; acc.ZZ = F::one(); acc.ZZZ = F::one();
      return;
    }
    E Pd = F::sub(F::mul(a.x, acc.ZZ), acc.X);    // U2 - X1
    E Rd = F::sub(F::mul(a.y, acc.ZZZ), acc.Y);   // S2 - Y1
    if (F::is_zero(Pd)) {
      if (F::is_zero(Rd)) acc = dbl_affine(a);
      else acc = infinity();
      return;
    }
    E PP = F::sqr(Pd);
    E PPP = F::mul(Pd, PP);
    E Q = F::mul(acc.X, PP);
    E X3 = F::sub(F::sub(F::sqr(Rd), PPP), F::dbl(Q));
    E Y3 = F::mul_sub2(Rd, F::sub(Q, X3), acc.Y, PPP);   // one reduction for both products
    acc.ZZ = F::mul(acc.ZZ, PP);
    acc.ZZZ = F::mul(acc.ZZZ, PPP);
    acc.X = X3;
    acc.Y = Y3;
  }

  // acc += b   (full addition, complete)
  BZ_HDI static void add(P& acc, const P& b) {
    if (is_inf(b)) return;
    if (is_inf(acc)) { acc = b; return; }
    E U1 = F::mul(acc.X, b.ZZ);
    E U2 = F::mul(b.X, acc.ZZ);
    E S1 = F::mul(acc.Y, b.ZZZ);
    E S2 = F::mul(b.Y, acc.ZZZ);
    E Pd = F::sub(U2, U1);
    E Rd = F::sub(S2, S1);
    if (F::is_zero(Pd)) {
      if (F::is_zero(Rd)) acc = dbl(acc);
      else acc = infinity();
      return;
    }
    E PP = F::sqr(Pd);
    E PPP = F::mul(Pd, PP);
    E Q = F::mul(U1, PP);
    E X3 = F::sub(F::sub(F::sqr(Rd), PPP), F::dbl(Q));
    E Y3 = F::mul_sub2(Rd, F::sub(Q, X3), S1, PPP);
    acc.ZZ = F::mul(F::mul(acc.ZZ, b.ZZ), PP);
    acc.ZZZ = F::mul(F::mul(acc.ZZZ, b.ZZZ), PPP);
    acc.X = X3;
    acc.Y = Y3;
  }

#ifdef __CUDACC__
  // ---- acc += b by FOUR lanes (an aligned quad of a warp, all holding the same acc and b).  The 14 field products
  // of the XYZZ addition form four dependency levels of <= 4 independent products: every lane computes one product of
  // the level and the quad exchanges the results with shuffles, so the latency of an addition is ~4 product latencies
  // instead of 14.  For the latency-bound tails of the pipeline (upper levels of the bucket reduction and of the
  // partial-merge tree: a few thousand additions on chains that one thread walks serially), not for throughput.
  __device__ __forceinline__ static E quad_get(const E& v, int src, unsigned mask) {
    E r;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) r.v[i] = __shfl_sync(mask, v.v[i], src, 4);
    return r;
  }
  __device__ __forceinline__ static E sel4(int q, const E& a, const E& b, const E& c, const E& d) {
    E r;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) r.v[i] = q == 0 ? a.v[i] : (q == 1 ? b.v[i] : (q == 2 ? c.v[i] : d.v[i]));
    return r;
  }
  __device__ static void add_quad(P& acc, const P& b) {
    const int lane = threadIdx.x & 31, q = lane & 3;
    const unsigned mask = 0xFu << (lane & ~3);
    if (is_inf(b)) return;                    // uniform within the quad: all four lanes hold the same values
    if (is_inf(acc)) { acc = b; return; }
    const E m1 = F::mul(sel4(q, acc.X, b.X, acc.Y, b.Y), sel4(q, b.ZZ, acc.ZZ, b.ZZZ, acc.ZZZ));
    const E U1 = quad_get(m1, 0, mask), U2 = quad_get(m1, 1, mask), S1 = quad_get(m1, 2, mask), S2 = quad_get(m1, 3, mask);
    const E Pd = F::sub(U2, U1), Rd = F::sub(S2, S1);
    if (F::is_zero(Pd)) {
      if (F::is_zero(Rd)) acc = dbl(acc);
      else acc = infinity();
      return;
    }
    const E m2 = F::mul(sel4(q, Pd, Rd, acc.ZZ, acc.ZZZ), sel4(q, Pd, Rd, b.ZZ, b.ZZZ));
    const E PP = quad_get(m2, 0, mask), RR = quad_get(m2, 1, mask), ZZ12 = quad_get(m2, 2, mask), ZZZ12 = quad_get(m2, 3, mask);
    const E m3 = F::mul(sel4(q, Pd, U1, ZZ12, ZZ12), PP);
    const E PPP = quad_get(m3, 0, mask), Q = quad_get(m3, 1, mask), ZZ3 = quad_get(m3, 2, mask);
    const E X3 = F::sub(F::sub(RR, PPP), F::dbl(Q));
    const E m4 = F::mul(sel4(q, Rd, S1, ZZZ12, ZZZ12), sel4(q, F::sub(Q, X3), PPP, PPP, PPP));
    const E t1 = quad_get(m4, 0, mask), t2 = quad_get(m4, 1, mask), ZZZ3 = quad_get(m4, 2, mask);
    acc.X = X3;
    acc.Y = F::sub(t1, t2);
    acc.ZZ = ZZ3;
    acc.ZZZ = ZZZ3;
  }
#endif

  // ---- affine + affine with a shared ("batched") inversion --------------------------------------
  // ba_classify picks the denominator whose inverse the addition needs and says which formula applies;
  // ba_finish completes the addition given that inverse.  Denominators are never zero, so thousands of
  // them can be inverted with one field inversion (Montgomery's trick, see msm_ba.cuh).
  //   kind 0: generic chord          den = x2 - x1     lambda = (y2 - y1) / den
  //   kind 1: tangent (p1 == p2)     den = 2 y1        lambda = 3 x1^2 / den
  //   kind 2: p1 is the identity     den = 1           result = p2
  //   kind 3: p2 is the identity     den = 1           result = p1
  //   kind 4: p1 == -p2              den = 1           result = identity (encoded (0, 0))
  BZ_HDI static int ba_classify(const A& p1, const A& p2, E& den) {
    if (is_identity(p1)) { den = F::one(); return 2; }
    if (is_identity(p2)) { den = F::one(); return 3; }
    E dx = F::sub(p2.x, p1.x);
    if (!F::is_zero(dx)) { den = dx; return 0; }
    if (F::is_zero(F::add(p1.y, p2.y))) { den = F::one(); return 4; }   // also covers y == 0
    den = F::dbl(p1.y);
    return 1;
  }
  BZ_HDI static A ba_finish(int kind, const A& p1, const A& p2, const E& inv_den) {
    if (kind == 2) return p2;
    if (kind == 3) return p1;
    A r;
    if (kind == 4) { r.x = F::zero(); r.y = F::zero(); return r; }
    E lam;
    if (kind == 0) {
      lam = F::mul(F::sub(p2.y, p1.y), inv_den);
    } else {
      E x2 = F::sqr(p1.x);
      lam = F::mul(F::add(F::dbl(x2), x2), inv_den);
    }
    r.x = F::sub(F::sub(F::sqr(lam), p1.x), p2.x);
    r.y = F::sub(F::mul(lam, F::sub(p1.x, r.x)), p1.y);
    return r;
  }

  BZ_HDI static A neg(const A& a) {
    A r;
    r.x = a.x;
    r.y = F::neg(a.y);
    return r;
  }

  // k * p for a small non-negative k (double-and-add, MSB first)
  BZ_HDI static P mul_small(const P& p, uint32_t k) {
    P r = infinity();
    for (int bit = 31; bit >= 0; bit--) {
      r = dbl(r);
      if ((k >> bit) & 1) add(r, p);
    }
    return r;
  }

  // XYZZ -> affine (Montgomery form); returns false for infinity
  BZ_HDI static bool to_affine(const P& p, A& out) {
    if (is_inf(p)) return false;
    // x = X/ZZ, y = Y/ZZZ; 1/ZZZ = t, 1/ZZ = t^2 * ZZZ... use one inversion of ZZZ*ZZ
    E zi = F::inv(F::mul(p.ZZ, p.ZZZ));           // 1/(ZZ*ZZZ)
    out.x = F::mul(p.X, F::mul(zi, p.ZZZ));       // X/ZZ
    out.y = F::mul(p.Y, F::mul(zi, p.ZZ));        // Y/ZZZ
    return true;
  }
};

}  // namespace bz
