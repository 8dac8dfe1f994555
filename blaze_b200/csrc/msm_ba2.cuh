// Bucket accumulation by batched AFFINE addition, fused into one kernel ("k_accumulate_ba").
//
// Black box being replaced: the FPGA MSM core's "bucket accumulation" phase
// (/root/reference/src/ingo_msm/msm_hw_code.rs:33).  Drop-in alternative to k_accumulate (msm_curve.cuh): same inputs
// (table, sorted refs, bucket offsets), same outputs (bucket sums + the two partial entries of every segment), so the
// merge tree and the bucket reduction behind it do not change.
//
// Why.  The sweep is bound by the integer multiplier.  An XYZZ mixed add costs ~9 field products; an
// affine add with the inversion shared by a batch costs 5 products + 1 squaring (1 for the running product of the
// denominators, 2 to peel a denominator's inverse off, 3 for lambda, x3, y3).  The batch is THREAD-LOCAL: a thread owns
// a segment of L sorted entries, sums the entries of every bucket run in the segment as a balanced tree
//     round r: entries (2j, 2j+1) of a run -> entry j (an odd tail is copied)
// so that round r offers L / 2^(r+1) independent additions, and inverts the product of their denominators ITSELF with
// the division-step inverse of ff.cuh (ALU pipe, ~25x fewer multiplier instructions than Fermat).  No cross-thread
// product trees, no separate forward / inverse / backward kernels, no global prefix arrays (round 1's msm_ba.cuh moved
// ~580 B per addition through separate memory-bound phases and broke even).  After `rounds` tree rounds (batches get
// too small to pay for an inversion) what is left of a run is folded with XYZZ mixed adds.
//
// Per-warp scratch (global memory, reused by every segment the warp processes, so it stays in L2 / L1):
//   sx, sy   affine list of the current round, in place (round r+1 overwrites the front of round r),
//   sp       running products of the denominators (the k-th pair's slot holds the product of the pairs after it),
// laid out [slot][16-byte piece][lane] so that a warp-wide access is a contiguous 512-byte run.
//
// All special cases are complete: equal operands (tangent), opposite operands (identity, encoded (0,0)), identity
// operands -- ec<C>::ba_classify / ba_finish.  The reference's tiled test vectors put thousands of copies of one point
// in a bucket (tests/msm/mod.rs:92-109): every pair of the first round is then a doubling.
#pragma once
#include "ec.cuh"
#include "msm_types.cuh"

namespace bz {

#ifdef __CUDACC__
#define BA2_LDG(p) __ldg(p)
#else
#define BA2_LDG(p) (*(p))
struct ba2_u4 { uint32_t x, y, z, w; };
#endif

template <class C>
struct ba2 {
  typedef typename C::Fq Fq;
  typedef ff<Fq> F;
  typedef Fe<Fq> E;
  typedef ec<C> G;
  static constexpr int N = Fq::N;
  static constexpr int NQ = N / 4;          // 16-byte pieces per field element
#ifdef __CUDACC__
  typedef uint4 Q;
#else
  typedef ba2_u4 Q;
#endif

  struct Ctx {
    const AffineT<C>* table;
    const uint32_t* sorted;
    const uint32_t* goff;
    uint32_t ngoff;
    uint32_t s, e;          // the segment: sorted positions [s, e)
    uint32_t g_first;       // bucket that contains position s
    Q *sx, *sy, *sp;        // lane-adjusted scratch bases
    uint32_t cap;           // scratch slots per lane
  };

  // ---- scratch element (slot o of this lane): piece k at ((o NQ + k) 32) 16-byte units from the lane's base
  BZ_HDI static E ld_s(const Q* base, uint32_t o) {
    E r;
#pragma unroll
    for (int k = 0; k < NQ; k++) {
      Q v = base[((size_t)o * NQ + k) * 32];
      r.v[4 * k] = v.x; r.v[4 * k + 1] = v.y; r.v[4 * k + 2] = v.z; r.v[4 * k + 3] = v.w;
    }
    return r;
  }
  BZ_HDI static void st_s(Q* base, uint32_t o, const E& r) {
#pragma unroll
    for (int k = 0; k < NQ; k++) {
      Q v;
      v.x = r.v[4 * k]; v.y = r.v[4 * k + 1]; v.z = r.v[4 * k + 2]; v.w = r.v[4 * k + 3];
      base[((size_t)o * NQ + k) * 32] = v;
    }
  }
  BZ_HDI static E ld_t(const uint32_t* p) {   // table coordinate (16-byte aligned)
    E r;
    const Q* q = reinterpret_cast<const Q*>(p);
#pragma unroll
    for (int k = 0; k < NQ; k++) {
      Q v = BA2_LDG(q + k);
      r.v[4 * k] = v.x; r.v[4 * k + 1] = v.y; r.v[4 * k + 2] = v.z; r.v[4 * k + 3] = v.w;
    }
    return r;
  }

  // operand i of the current round's input list: round 0 reads the table through the sorted refs, later rounds the scratch
  BZ_HDI static E in_x(const Ctx& c, int r, uint32_t i) {
    if (r == 0) return ld_t(c.table[BA2_LDG(c.sorted + c.s + i) & 0x7fffffffu].x);
    return ld_s(c.sx, i);
  }
  BZ_HDI static Affine<C> in_pt(const Ctx& c, int r, uint32_t i) {
    Affine<C> a;
    if (r == 0) {
      const uint32_t ent = BA2_LDG(c.sorted + c.s + i);
      const uint32_t* rec = c.table[ent & 0x7fffffffu].x;
      a.x = ld_t(rec);
      a.y = ld_t(rec + N);
      if (ent & 0x80000000u) a.y = F::neg(a.y);
    } else {
      a.x = ld_s(c.sx, i);
      a.y = ld_s(c.sy, i);
    }
    return a;
  }

  // entries of bucket g inside the segment before round r
  BZ_HDI static uint32_t run_len(const Ctx& c, uint32_t g, int r) {
    const uint32_t b0 = BA2_LDG(c.goff + g), b1 = BA2_LDG(c.goff + g + 1);
    const uint32_t lo = b0 > c.s ? b0 : c.s, hi = b1 < c.e ? b1 : c.e;
    const uint32_t n0 = hi > lo ? hi - lo : 0;
    return (n0 + ((1u << r) - 1)) >> r;
  }

  // the XYZZ sweep of k_accumulate for one segment: used when a segment has too many tiny runs for the scratch
  // (and as the tail after the tree rounds, through fold_run)
  BZ_HDI static void store_run(const Ctx& c, uint64_t t, uint32_t g, const XYZZ<C>& acc, XyzzM<C>* buckets, XyzzM<C>* part_pt,
                               uint32_t& id0, uint32_t& id1) {
    const uint32_t b0 = BA2_LDG(c.goff + g), b1 = BA2_LDG(c.goff + g + 1);
    XyzzM<C>* dst;
    if (b0 >= c.s && b1 <= c.e) dst = buckets + g;
    else if (b0 < c.s) { id0 = g; dst = part_pt + 2 * t; }
    else { id1 = g; dst = part_pt + 2 * t + 1; }
#pragma unroll
    for (int k = 0; k < N; k++) { dst->X[k] = acc.X.v[k]; dst->Y[k] = acc.Y.v[k]; dst->ZZ[k] = acc.ZZ.v[k]; dst->ZZZ[k] = acc.ZZZ.v[k]; }
  }

  // ---- affine + affine given the inverse of the denominator (ec<C>::ba_classify / ba_finish with the additions and
  // subtractions out of line: the backward loop stays small enough for the instruction cache)
  BZ_HDI static E sub_(const E& a, const E& b) { return F::sub_call(a, b); }
  BZ_HDI static int classify(const Affine<C>& p1, const Affine<C>& p2, E& den) {
    if (G::is_identity(p1)) { den = F::one(); return 2; }
    if (G::is_identity(p2)) { den = F::one(); return 3; }
    den = sub_(p2.x, p1.x);
    if (!F::is_zero(den)) return 0;
    if (F::is_zero(F::add_call(p1.y, p2.y))) { den = F::one(); return 4; }   // also covers y == 0
    den = F::add_call(p1.y, p1.y);
    return 1;
  }
  BZ_HDI static Affine<C> finish(int kind, const Affine<C>& p1, const Affine<C>& p2, const E& inv_den) {
    if (kind == 2) return p2;
    if (kind == 3) return p1;
    Affine<C> r;
    if (kind == 4) { r.x = F::zero(); r.y = F::zero(); return r; }
    E lam;
    if (kind == 0) {
      lam = F::mul(sub_(p2.y, p1.y), inv_den);
    } else {
      const E x2 = F::sqr(p1.x);
      lam = F::mul(F::add_call(F::add_call(x2, x2), x2), inv_den);
    }
    r.x = sub_(sub_(F::sqr(lam), p1.x), p2.x);
    r.y = sub_(F::mul(lam, sub_(p1.x, r.x)), p1.y);
    return r;
  }

  // bring operand i of round r's input list towards the SM ahead of its use (table line, or the lane's scratch pieces)
  BZ_HDI static void prefetch_in(const Ctx& c, int r, uint32_t i, bool with_y) {
#ifdef __CUDACC__
    if (r == 0) {
      const uint32_t ent = __ldg(c.sorted + c.s + i);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(c.table[ent & 0x7fffffffu].x));
    } else {
#pragma unroll
      for (int k = 0; k < NQ; k++) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(c.sx + ((size_t)i * NQ + k) * 32));
        if (with_y) asm volatile("prefetch.global.L1 [%0];" ::"l"(c.sy + ((size_t)i * NQ + k) * 32));
      }
    }
#endif
  }

  // one segment.  `rounds` tree rounds of batched affine additions, then an XYZZ fold of what is left of every run.
  BZ_HDI static void segment(Ctx& c, uint64_t t, int rounds, XyzzM<C>* buckets, uint32_t* part_id, XyzzM<C>* part_pt) {
    uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu;
    // last bucket with entries in the segment, and whether round 0's output fits the scratch
    uint32_t g_last = c.g_first;
    {
      uint32_t g = c.g_first, need = 0;
      while (g < c.ngoff && BA2_LDG(c.goff + g) < c.e) {
        const uint32_t n = run_len(c, g, 0);
        if (n) g_last = g;
        need += (n + 1) >> 1;
        g++;
      }
      if (need > c.cap) rounds = 0;   // many tiny runs: nothing to batch, fold directly
    }
    for (int r = 0; r < rounds; r++) {
      // ---- forward, last slot first: running product of the pair denominators
      uint32_t I = 0, O = 0;
      for (uint32_t g = c.g_first; g <= g_last; g++) { const uint32_t n = run_len(c, g, r); I += n; O += (n + 1) >> 1; }
      if (I == O) { rounds = r; break; }   // every run is down to one entry: the list stays where round r-1 left it
      E prod = F::one();
      uint32_t npairs = 0;
      {
        uint32_t g = g_last + 1, n = 0, Ib = I, Ob = O;
        int32_t j = -1;
        for (;;) {
          if (j < 0) {
            if (g == c.g_first) break;
            g--;
            n = run_len(c, g, r);
            Ib -= n;
            Ob -= (n + 1) >> 1;
            j = (int32_t)((n + 1) >> 1) - 1;
            continue;
          }
          const uint32_t i1 = Ib + 2 * j;
          if (i1 >= 4) { prefetch_in(c, r, i1 - 3, false); prefetch_in(c, r, i1 - 4, false); }   // two slots ahead (descending)
          if (2u * (uint32_t)j + 1 < n) {
            const E x1 = in_x(c, r, i1), x2 = in_x(c, r, i1 + 1);
            E den = sub_(x2, x1);
            if (F::is_zero(x1) || F::is_zero(x2) || F::is_zero(den)) {   // identity operand, tangent or cancellation
              const Affine<C> p1 = in_pt(c, r, i1), p2 = in_pt(c, r, i1 + 1);
              classify(p1, p2, den);
            }
            st_s(c.sp, Ob + j, prod);
            prod = F::mul(prod, den);
            npairs++;
          }
          j--;
        }
      }
      E inv = npairs ? F::inv_gcd_call(prod) : F::one();
      // ---- backward, first slot first: peel the inverses off and finish the additions, in place
      {
        uint32_t g = c.g_first, n = run_len(c, g, r), Ib = 0, Ob = 0, j = 0, cnt = (n + 1) >> 1;
        for (;;) {
          if (j == cnt) {
            if (g == g_last) break;
            Ib += n;
            Ob += cnt;
            g++;
            n = run_len(c, g, r);
            cnt = (n + 1) >> 1;
            j = 0;
            continue;
          }
          const uint32_t i1 = Ib + 2 * j;
          if (i1 + 5 < I) { prefetch_in(c, r, i1 + 4, true); prefetch_in(c, r, i1 + 5, true); }   // two slots ahead
          const Affine<C> p1 = in_pt(c, r, i1);
          if (2 * j + 1 < n) {
            const Affine<C> p2 = in_pt(c, r, i1 + 1);
            E den;
            const int kind = classify(p1, p2, den);
            const E dinv = F::mul(inv, ld_s(c.sp, Ob + j));
            inv = F::mul(inv, den);
            const Affine<C> sum = finish(kind, p1, p2, dinv);
            st_s(c.sx, Ob + j, sum.x);
            st_s(c.sy, Ob + j, sum.y);
          } else {
            st_s(c.sx, Ob + j, p1.x);
            st_s(c.sy, Ob + j, p1.y);
          }
          j++;
        }
      }
    }
    // ---- what is left of every run: XYZZ mixed adds, then the bucket (or the segment's head / tail partial).  ONE flat
    // loop over the remaining entries (lanes of a warp hold different run structures: nested loops would serialise them)
    {
      uint32_t g = c.g_first, n = run_len(c, g, rounds), Ib = 0, j = 0;
      XYZZ<C> acc = G::infinity();
      for (;;) {
        if (j == n) {
          if (n) store_run(c, t, g, acc, buckets, part_pt, id0, id1);
          if (g == g_last) break;
          Ib += n;
          g++;
          n = run_len(c, g, rounds);
          j = 0;
          acc = G::infinity();
          continue;
        }
        const Affine<C> a = in_pt(c, rounds, Ib + j);
        G::madd(acc, a);
        j++;
      }
    }
    part_id[2 * t] = id0;
    part_id[2 * t + 1] = id1;
  }
};

#ifdef __CUDACC__
#ifndef BZ_BA2_MINBLOCKS
#define BZ_BA2_MINBLOCKS 3
#endif
// Persistent grid: every warp takes groups of 32 consecutive segments (lane = segment) and reuses ITS scratch.
template <class C>
__global__ void __launch_bounds__(128, BZ_BA2_MINBLOCKS)
k_accumulate_ba(const AffineT<C>* __restrict__ table, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ goff,
                XyzzM<C>* __restrict__ buckets, uint32_t* __restrict__ part_id, XyzzM<C>* __restrict__ part_pt, uint64_t nseg,
                uint32_t L, uint32_t ngoff, int rounds, uint4* __restrict__ scratch, uint32_t cap) {
  typedef ba2<C> B;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t total = __ldg(goff + ngoff);   // zero digits are dropped by the sort: data dependent
  const size_t plane = (size_t)cap * B::NQ * 32;   // 16-byte units of one scratch array of one warp
  typename B::Ctx c;
  c.table = table;
  c.sorted = sorted;
  c.goff = goff;
  c.ngoff = ngoff;
  c.cap = cap;
  c.sx = scratch + warp * 3 * plane + lane;
  c.sy = c.sx + plane;
  c.sp = c.sy + plane;
  for (uint64_t item = warp; item * 32 < nseg; item += nwarps) {
    const uint64_t t = item * 32 + lane;
    if (t < nseg) {
      const uint64_t s64 = t * L;
      if (s64 >= total) {
        part_id[2 * t] = 0xffffffffu;
        part_id[2 * t + 1] = 0xffffffffu;
      } else {
        c.s = (uint32_t)s64;
        c.e = (uint32_t)(s64 + L < total ? s64 + L : total);
        uint32_t lo = 0, hi = ngoff;   // largest g with goff[g] <= s
        while (hi - lo > 1) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          if (__ldg(goff + mid) <= c.s) lo = mid; else hi = mid;
        }
        c.g_first = lo;
        B::segment(c, t, rounds, buckets, part_id, part_pt);
      }
    }
    __syncwarp();
  }
}
#endif

}  // namespace bz
