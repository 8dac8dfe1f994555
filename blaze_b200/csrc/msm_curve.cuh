// MSM back end, one instantiation per curve: bucket accumulation over the sorted index
// lists, partial-bucket merge, bucket reduction (running sums), window combine (Horner) and
// result serialisation.
//
// Black box being replaced: the FPGA MSM core's "bucket accumulation" and "final accumulation"
// phases (/root/reference/src/ingo_msm/msm_hw_code.rs:33-34) and its RESULT register window
// (msm_api.rs:240-274: result_point_size bytes, Z||Y||X per tests/msm/mod.rs:397-403).
//
// Accumulation is load-balanced by SEGMENT, not by bucket: thread t owns sorted entries
// [t*L, (t+1)*L) of the concatenation of all windows and walks the bucket boundaries (goff)
// as it goes, so 2^18 copies of one point in one bucket (the reference's tiled test vectors,
// tests/msm/mod.rs:92-109) cost the same as uniformly random scalars.  Buckets that lie inside
// one segment are stored directly; the (at most two) buckets cut by a segment boundary go to a
// partial list that k_merge_partials folds afterwards.
#pragma once
#include <cuda_runtime.h>

#include "ec.cuh"
#include "msm_internal.h"
#include "msm_types.cuh"
#include "msm_ba2.cuh"

namespace bz {

template <class C>
struct dev {
  typedef typename C::Fq Fq;
  static constexpr int N = Fq::N;
  typedef ff<Fq> F;
  typedef ec<C> G;

  // rec points at x[0] of an AffineM / AffineT record (x then y, contiguous)
  __device__ __forceinline__ static Affine<C> load_affine(const uint32_t* rec) {
    Affine<C> a;
    const uint4* q = reinterpret_cast<const uint4*>(rec);
#pragma unroll
    for (int k = 0; k < N / 4; k++) {
      uint4 v = __ldg(q + k);
      a.x.v[4 * k] = v.x; a.x.v[4 * k + 1] = v.y; a.x.v[4 * k + 2] = v.z; a.x.v[4 * k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < N / 4; k++) {
      uint4 v = __ldg(q + N / 4 + k);
      a.y.v[4 * k] = v.x; a.y.v[4 * k + 1] = v.y; a.y.v[4 * k + 2] = v.z; a.y.v[4 * k + 3] = v.w;
    }
    return a;
  }
  __device__ __forceinline__ static void ld_limbs(uint32_t* dst, const uint4* q) {
#pragma unroll
    for (int k = 0; k < N / 4; k++) {
      uint4 v = q[k];
      dst[4 * k] = v.x; dst[4 * k + 1] = v.y; dst[4 * k + 2] = v.z; dst[4 * k + 3] = v.w;
    }
  }
  __device__ __forceinline__ static void st_limbs(uint4* q, const uint32_t* src) {
#pragma unroll
    for (int k = 0; k < N / 4; k++) q[k] = make_uint4(src[4 * k], src[4 * k + 1], src[4 * k + 2], src[4 * k + 3]);
  }
  __device__ __forceinline__ static XYZZ<C> load_xyzz(const XyzzM<C>* p) {
    XYZZ<C> r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    ld_limbs(r.X.v, q);
    ld_limbs(r.Y.v, q + N / 4);
    ld_limbs(r.ZZ.v, q + 2 * (N / 4));
    ld_limbs(r.ZZZ.v, q + 3 * (N / 4));
    return r;
  }
  __device__ __forceinline__ static void store_xyzz(XyzzM<C>* p, const XYZZ<C>& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    st_limbs(q, r.X.v);
    st_limbs(q + N / 4, r.Y.v);
    st_limbs(q + 2 * (N / 4), r.ZZ.v);
    st_limbs(q + 3 * (N / 4), r.ZZZ.v);
  }
  // canonical little-endian bytes (4-byte aligned) -> Montgomery
  __device__ __forceinline__ static Fe<Fq> load_canonical(const uint8_t* b) {
    Fe<Fq> r;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(b);
#pragma unroll
    for (int k = 0; k < N; k++) r.v[k] = w[k];
    return F::to_mont(r);
  }
  __device__ __forceinline__ static void store_canonical(uint8_t* b, const Fe<Fq>& a) {
    Fe<Fq> r = F::from_mont(a);
    uint32_t* w = reinterpret_cast<uint32_t*>(b);
#pragma unroll
    for (int k = 0; k < N; k++) w[k] = r.v[k];
  }
  // Z||Y||X homogeneous projective WITHOUT the normalising inversion (x = X/Z, y = Y/Z -- the reference's own
  // result format, tests/msm/mod.rs:397-403): X' = X ZZZ, Y' = Y ZZ, Z' = ZZ ZZZ.  For shards of a multi-GPU
  // MSM, whose records are summed (and normalised once) by k_combine_results.
  __device__ static void store_result_raw(uint8_t* out, const XYZZ<C>& p) {
    uint32_t* w = reinterpret_cast<uint32_t*>(out);
    for (int k = 0; k < 3 * N; k++) w[k] = 0;
    if (G::is_inf(p)) { w[N] = 1; return; }
    store_canonical(out, F::mul(p.ZZ, p.ZZZ));
    store_canonical(out + 4 * N, F::mul(p.Y, p.ZZ));
    store_canonical(out + 8 * N, F::mul(p.X, p.ZZZ));
  }
  // Z||Y||X with Z = 1 (infinity: Z=0, Y=1, X=0), canonical LE
  __device__ static void store_result(uint8_t* out, const XYZZ<C>& p) {
    Affine<C> a;
    uint32_t* w = reinterpret_cast<uint32_t*>(out);
    for (int k = 0; k < 3 * N; k++) w[k] = 0;
    if (!G::to_affine(p, a)) { w[N] = 1; return; }
    w[0] = 1;
    store_canonical(out + 4 * N, a.y);
    store_canonical(out + 8 * N, a.x);
  }
};

// ---------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) k_points_to_mont(const uint8_t* __restrict__ raw, AffineT<C>* __restrict__ table,
                                                        uint64_t n) {
  typedef dev<C> D;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = raw + i * (2 * C::FQ_BYTES);
  Fe<typename C::Fq> x = D::load_canonical(p), y = D::load_canonical(p + C::FQ_BYTES);
#pragma unroll
  for (int k = 0; k < D::N; k++) { table[i].x[k] = x.v[k]; table[i].y[k] = y.v[k]; }
}

// ---------------------------------------------------------------------------------------------
// Window-merged table ("precomputed points resident in HBM", the HBM-mode path): level w holds
// 2^(c w) * P_i, so digit window w of scalar i can use table entry w*n + i and ALL windows share one
// bucket set -- one running-sum reduction instead of W, which makes wider windows (fewer mixed adds per
// scalar) affordable.  A level is made from the one below: c Jacobian doublings per point (a = 0:
// dbl-2009-l, 2M + 5S), then back to affine with ONE field inversion per thread shared by K points
// (Montgomery's trick).  Built once per resident point set, reused by every MSM on it.
template <class C, int K>
__global__ void __launch_bounds__(128) k_wtable_level(const AffineT<C>* __restrict__ prev, AffineT<C>* __restrict__ next,
                                                      uint64_t n, int c) {
  typedef dev<C> D;
  typedef ff<typename C::Fq> F;
  typedef Fe<typename C::Fq> E;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i0 = t * K;
  if (i0 >= n) return;
  E Xs[K], Ys[K], Zs[K], pre[K];   // local memory (indexed dynamically): this kernel runs once per point set
  E run = F::one();
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    Zs[k] = F::zero();
    pre[k] = run;
    if (i0 + k >= n) continue;
    Affine<C> a = D::load_affine(prev[i0 + k].x);
    if (ec<C>::is_identity(a)) continue;
    E X = a.x, Y = a.y, Z = F::one();
#pragma unroll 1
    for (int d = 0; d < c; d++) {
      E A = F::sqr(X), B = F::sqr(Y), Cc = F::sqr(B);
      E t0 = F::sub(F::sub(F::sqr(F::add(X, B)), A), Cc);
      E Dd = F::dbl(t0);
      E Ee = F::add(F::dbl(A), A);
      E Ff = F::sqr(Ee);
      E Z3 = F::dbl(F::mul(Y, Z));
      X = F::sub(Ff, F::dbl(Dd));
      E c8 = F::dbl(F::dbl(F::dbl(Cc)));
      Y = F::sub(F::mul(Ee, F::sub(Dd, X)), c8);
      Z = Z3;
    }
    if (F::is_zero(Z)) continue;   // 2-torsion input (possible on curves with even cofactor): identity from here on
    Xs[k] = X; Ys[k] = Y; Zs[k] = Z;
    run = F::mul(run, Z);
  }
  E inv = F::inv(run);
#pragma unroll 1
  for (int k = K - 1; k >= 0; k--) {
    if (i0 + k >= n) continue;
    uint32_t* ox = next[i0 + k].x;
    uint32_t* oy = next[i0 + k].y;
    if (F::is_zero(Zs[k])) {
#pragma unroll
      for (int j = 0; j < D::N; j++) { ox[j] = 0; oy[j] = 0; }
      continue;
    }
    E zi = F::mul(inv, pre[k]);
    inv = F::mul(inv, Zs[k]);
    E zi2 = F::sqr(zi);
    E x = F::mul(Xs[k], zi2);
    E y = F::mul(Ys[k], F::mul(zi2, zi));
#pragma unroll
    for (int j = 0; j < D::N; j++) { ox[j] = x.v[j]; oy[j] = y.v[j]; }
  }
}

// level-major Montgomery table (entry w*n + i) -> the reference's precomputed wire format (tests/msm/mod.rs:360-380):
// record i = level 0 .. levels-1 of point i, affine x||y canonical little-endian
template <class C>
__global__ void __launch_bounds__(128) k_table_to_wire(const AffineT<C>* __restrict__ table, uint64_t n, int levels,
                                                       uint8_t* __restrict__ out) {
  typedef dev<C> D;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (uint64_t)levels) return;
  const uint64_t i = t / levels;
  const int w = (int)(t % levels);
  Affine<C> a = D::load_affine(table[(uint64_t)w * n + i].x);
  uint8_t* o = out + t * (2 * C::FQ_BYTES);
  D::store_canonical(o, a.x);
  D::store_canonical(o + C::FQ_BYTES, a.y);
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation (the dominant kernel)
// CTAs of 128 threads per SM.  3 (168 registers: the 12-limb curves spill ~170 B/thread around the field calls)
// beats 2 (228 registers, no spills) by 5% since the field products became calls: the third warp per scheduler
// covers the fixed-latency waits of the carry chains (measured at 2^24: 68.6 vs 72.2 ms).
#ifndef BZ_ACC_MINBLOCKS
#define BZ_ACC_MINBLOCKS 3
#endif
#ifndef BZ_ACC_MINBLOCKS_N8   // 8-limb base field (BN254)
#define BZ_ACC_MINBLOCKS_N8 4   // 128 registers, 72 B of spills: 33.2 -> 32.5 ms at 2^24
#endif
template <class C>
__global__ void __launch_bounds__(128, (C::Fq::N == 8 ? BZ_ACC_MINBLOCKS_N8 : BZ_ACC_MINBLOCKS))
k_accumulate(const AffineT<C>* __restrict__ table, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ goff, XyzzM<C>* __restrict__ buckets, uint32_t* __restrict__ part_id,
             XyzzM<C>* __restrict__ part_pt, uint64_t nseg, uint32_t L, uint32_t nb, uint32_t ngoff) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  const uint64_t total = __ldg(goff + ngoff);   // zero digits are dropped by the sort: data dependent
  uint64_t s64 = t * L;
  if (s64 >= total) {
    part_id[2 * t] = 0xffffffffu;
    part_id[2 * t + 1] = 0xffffffffu;
    return;
  }
  const uint32_t s = (uint32_t)s64;
  const uint32_t e = (uint32_t)(s64 + L < total ? s64 + L : total);
  // largest g with goff[g] <= s
  uint32_t lo = 0, hi = ngoff;
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(goff + mid) <= s) lo = mid; else hi = mid;
  }
  uint32_t g = lo;
  uint32_t bstart = __ldg(goff + g), bend = __ldg(goff + g + 1);
  const bool skip = false;   // every bucket slot is a real bucket (zero digits never reach the sort)
  uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu;
  XYZZ<C> acc = G::infinity();
  // software pipeline: the point of entry pos+1 and the index of entry pos+2 are in flight while the
  // mixed add of entry pos runs (a gather miss costs ~1 us, an add ~8 us)
  uint32_t ent_n = __ldg(sorted + s);
  Affine<C> a_n = D::load_affine(table[ent_n & 0x7fffffffu].x);
  uint32_t ent_n2 = s + 1 < e ? __ldg(sorted + s + 1) : 0;

  for (uint32_t pos = s; pos < e; pos++) {
    const uint32_t ent = ent_n;
    Affine<C> a = a_n;
    ent_n = ent_n2;
    if (pos + 1 < e) a_n = D::load_affine(table[ent_n & 0x7fffffffu].x);
    if (pos + 2 < e) ent_n2 = __ldg(sorted + pos + 2);
    if (pos == bend) {
      // bucket g is finished (bend <= e here)
      if (!skip) {
        if (bstart >= s) D::store_xyzz(buckets + g, acc);
        else { id0 = g; D::store_xyzz(part_pt + 2 * t, acc); }
      }
      do {
        g++;
        bstart = bend;
        bend = __ldg(goff + g + 1);
      } while (bend == pos);
      acc = G::infinity();
    }
    if (!skip) {
      if (ent & 0x80000000u) a.y = ff<typename C::Fq>::neg(a.y);
      G::madd(acc, a);
    }
  }
  if (!skip) {
    if (bstart >= s && bend <= e) D::store_xyzz(buckets + g, acc);
    else if (bstart < s) { id0 = g; D::store_xyzz(part_pt + 2 * t, acc); }
    else { id1 = g; D::store_xyzz(part_pt + 2 * t + 1, acc); }
  }
  part_id[2 * t] = id0;
  part_id[2 * t + 1] = id1;
}

// ---------------------------------------------------------------------------------------------
// k_accumulate with the gathered points STAGED THROUGH SHARED MEMORY by the bulk-copy engine:
// every lane issues `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` (SASS: UBLKCP) for
// the point of its NEXT sorted entry into a per-warp two-stage ring and the warp's mbarrier of that
// stage collects the 32 arrivals + transferred bytes; the mixed add of the current entry runs meanwhile.
// Compared with the register-prefetch version above this frees the 24 registers of the in-flight point.
// Ring slot stride is 112 B (7 x 16): 8 consecutive lanes reading 16-byte pieces hit 8 distinct banks.
#define BZ_ACC_RING_STRIDE 112
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <class C>
__global__ void __launch_bounds__(128, (C::Fq::N == 8 ? BZ_ACC_MINBLOCKS_N8 : BZ_ACC_MINBLOCKS))
k_accumulate_tma(const AffineT<C>* __restrict__ table, const uint32_t* __restrict__ sorted,
                 const uint32_t* __restrict__ goff, XyzzM<C>* __restrict__ buckets, uint32_t* __restrict__ part_id,
                 XyzzM<C>* __restrict__ part_pt, uint64_t nseg, uint32_t L, uint32_t nb, uint32_t ngoff) {
  typedef dev<C> D;
  typedef ec<C> G;
  constexpr int N = C::Fq::N;
  constexpr uint32_t PT_BYTES = 2 * N * 4;
  extern __shared__ __align__(16) uint8_t ring_raw[];
  // layout per warp: [2 stages][32 lanes][112 B], then (after all rings) [4 warps][2] mbarriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  uint8_t* ring = ring_raw + (size_t)warp * 2 * 32 * BZ_ACC_RING_STRIDE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_raw + (size_t)nwarps * 2 * 32 * BZ_ACC_RING_STRIDE) + 2 * warp;
  const uint32_t bar0 = smem_u32(bars), bar1 = bar0 + 8;
  if (lane == 0) { mbar_init(bar0, 32); mbar_init(bar1, 32); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();

  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = __ldg(goff + ngoff);
  const uint64_t s64 = t * L;
  const bool active = t < nseg && s64 < total;
  const uint32_t s = active ? (uint32_t)s64 : 0;
  const uint32_t e = active ? (uint32_t)(s64 + L < total ? s64 + L : total) : 0;
  uint32_t g = 0, bstart = 0, bend = 0;
  bool skip = true;
  if (active) {
    uint32_t lo = 0, hi = ngoff;
    while (hi - lo > 1) {
      uint32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(goff + mid) <= s) lo = mid; else hi = mid;
    }
    g = lo;
    bstart = __ldg(goff + g);
    bend = __ldg(goff + g + 1);
    skip = false;
  }
  uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu;
  XYZZ<C> acc = G::infinity();

  // stage 0 <- entry s; index of entry s+1 in flight
  uint32_t ent_n = active ? __ldg(sorted + s) : 0;
  uint32_t ent_n2 = (active && s + 1 < e) ? __ldg(sorted + s + 1) : 0;
  {
    const uint32_t dst = smem_u32(ring + (size_t)lane * BZ_ACC_RING_STRIDE);
    if (active) {
      mbar_arrive_expect_tx(bar0, PT_BYTES);
      bulk_g2s(dst, table[ent_n & 0x7fffffffu].x, PT_BYTES, bar0);
    } else {
      mbar_arrive(bar0);
    }
  }
  // every lane of the warp runs the same number of iterations (the mbarriers count 32 arrivals)
  for (uint32_t it = 0; it < L; it++) {
    const uint32_t pos = s + it;
    const bool have = active && pos < e;
    const uint32_t ent = ent_n;
    ent_n = ent_n2;
    // refill the other stage with the point of entry pos+1 (its previous contents were consumed at it-1)
    {
      const uint32_t stage = (it + 1) & 1;
      const uint32_t bar = stage ? bar1 : bar0;
      if (it + 1 < L) {
        if (active && pos + 1 < e) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const uint32_t dst = smem_u32(ring + ((size_t)stage * 32 + lane) * BZ_ACC_RING_STRIDE);
          mbar_arrive_expect_tx(bar, PT_BYTES);
          bulk_g2s(dst, table[ent_n & 0x7fffffffu].x, PT_BYTES, bar);
        } else {
          mbar_arrive(bar);
        }
      }
    }
    if (active && pos + 2 < e) ent_n2 = __ldg(sorted + pos + 2);
    // wait for this iteration's stage (k-th use of a stage completes phase k)
    {
      const uint32_t stage = it & 1;
      const uint32_t bar = stage ? bar1 : bar0;
      const uint32_t parity = (it >> 1) & 1;
      while (!mbar_try_wait(bar, parity)) {}
    }
    if (!have) continue;
    if (pos == bend) {
      if (!skip) {
        if (bstart >= s) D::store_xyzz(buckets + g, acc);
        else { id0 = g; D::store_xyzz(part_pt + 2 * t, acc); }
      }
      do {
        g++;
        bstart = bend;
        bend = __ldg(goff + g + 1);
      } while (bend == pos);
      acc = G::infinity();
    }
    if (!skip) {
      const uint4* q = reinterpret_cast<const uint4*>(ring + ((size_t)(it & 1) * 32 + lane) * BZ_ACC_RING_STRIDE);
      Affine<C> a;
#pragma unroll
      for (int k = 0; k < N / 4; k++) {
        uint4 v = q[k];
        a.x.v[4 * k] = v.x; a.x.v[4 * k + 1] = v.y; a.x.v[4 * k + 2] = v.z; a.x.v[4 * k + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < N / 4; k++) {
        uint4 v = q[N / 4 + k];
        a.y.v[4 * k] = v.x; a.y.v[4 * k + 1] = v.y; a.y.v[4 * k + 2] = v.z; a.y.v[4 * k + 3] = v.w;
      }
      if (ent & 0x80000000u) a.y = ff<typename C::Fq>::neg(a.y);
      G::madd(acc, a);
    }
  }
  if (t >= nseg) return;
  if (active && !skip) {
    if (bstart >= s && bend <= e) D::store_xyzz(buckets + g, acc);
    else if (bstart < s) { id0 = g; D::store_xyzz(part_pt + 2 * t, acc); }
    else { id1 = g; D::store_xyzz(part_pt + 2 * t + 1, acc); }
  }
  part_id[2 * t] = id0;
  part_id[2 * t + 1] = id1;
}

// Fold the partial sums of buckets that straddle segment boundaries -- as a tree, so that one
// bucket holding millions of entries (all scalars equal, or the reference's tiled test vectors) does
// not serialise on one thread.  Level k: a thread owns `group` consecutive child units (segments at
// level 1, groups of the level below afterwards), i.e. positions [g*span, (g+1)*span) of the sorted
// array; it walks the children's (head, tail) partial entries in order, sums equal ids, and flushes
// a finished run to its bucket if the bucket lies inside the span, else to its own head / tail slot.
// The top level spans everything, so every remaining run lands in its bucket.
// QUAD: four lanes per group (ec<C>::add_quad): the same walk with ~3.5x shorter addition latency, for the levels that
// are too small to fill the machine (every level but the first of a large MSM)
template <class C, bool QUAD>
__device__ __forceinline__ void merge_walk(uint64_t g, bool writer, const uint32_t* goff, uint64_t total, XyzzM<C>* buckets,
                                           const uint32_t* in_id, const XyzzM<C>* in_pt, uint64_t n_children, uint32_t group,
                                           uint64_t span, uint32_t* out_id, XyzzM<C>* out_pt, uint64_t n_groups) {
  typedef dev<C> D;
  typedef ec<C> G;
  const uint64_t lo = g * span;
  uint64_t hi = lo + span;
  if (hi > total || n_groups == 1) hi = total;
  uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu;
  uint32_t cur = 0xffffffffu;
  XYZZ<C> acc = G::infinity();
  uint64_t c0 = g * group, c1 = c0 + group;
  if (c1 > n_children) c1 = n_children;
  for (uint64_t e = 2 * c0; e <= 2 * c1; e++) {
    uint32_t id = e < 2 * c1 ? in_id[e] : 0xfffffffeu;   // sentinel flushes the last run
    if (id == 0xffffffffu) continue;
    if (id != cur) {
      if (cur != 0xffffffffu) {
        bool left_open = (uint64_t)goff[cur] < lo, right_open = (uint64_t)goff[cur + 1] > hi;
        if (!left_open && !right_open) { if (writer) D::store_xyzz(buckets + cur, acc); }
        else if (left_open) { id0 = cur; if (writer) D::store_xyzz(out_pt + 2 * g, acc); }
        else { id1 = cur; if (writer) D::store_xyzz(out_pt + 2 * g + 1, acc); }
      }
      if (id == 0xfffffffeu) break;
      cur = id;
      acc = D::load_xyzz(in_pt + e);
    } else {
      XYZZ<C> o = D::load_xyzz(in_pt + e);
      if (QUAD) G::add_quad(acc, o);
      else G::add(acc, o);
    }
  }
  if (writer) {
    out_id[2 * g] = id0;
    out_id[2 * g + 1] = id1;
  }
}

template <class C, bool QUAD>
__global__ void __launch_bounds__(128) k_merge_level(const uint32_t* __restrict__ goff, uint32_t ngoff,
                                                     XyzzM<C>* __restrict__ buckets,
                                                     const uint32_t* __restrict__ in_id,
                                                     const XyzzM<C>* __restrict__ in_pt, uint64_t n_children,
                                                     uint32_t group, uint64_t span /* positions per group */,
                                                     uint32_t* __restrict__ out_id, XyzzM<C>* __restrict__ out_pt,
                                                     uint64_t n_groups) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool writer = !QUAD || (threadIdx.x & 3) == 0;
  if (QUAD) g >>= 2;
  if (g >= n_groups) return;   // whole quads leave together (n_groups * 4 threads are launched in multiples of 4)
  merge_walk<C, QUAD>(g, writer, goff, __ldg(goff + ngoff), buckets, in_id, in_pt, n_children, group, span, out_id, out_pt, n_groups);
}

// ALL remaining levels of the merge tree in one launch (one CTA, four lanes per walk, __syncthreads between levels): from
// the level with <= 16 groups on
template <class C>
__global__ void __launch_bounds__(256, 1)
k_merge_top(const uint32_t* goff, uint32_t ngoff, XyzzM<C>* buckets, const uint32_t* in_id, const XyzzM<C>* in_pt, uint64_t n_children,
            uint32_t group, uint64_t child_span, uint32_t* out_id, XyzzM<C>* out_pt) {
  const uint32_t quad = threadIdx.x >> 2, nquads = blockDim.x >> 2;
  const bool writer = (threadIdx.x & 3) == 0;
  const uint64_t total = goff[ngoff];
  for (;;) {
    const uint64_t span = child_span * group;
    const uint64_t n_groups = (n_children + group - 1) / group;
    for (uint64_t g = quad; g < n_groups; g += nquads)
      merge_walk<C, true>(g, writer, goff, total, buckets, in_id, in_pt, n_children, group, span, out_id, out_pt, n_groups);
    __syncthreads();
    if (n_groups == 1) break;
    in_id = out_id;
    in_pt = out_pt;
    out_id += 2 * n_groups;
    out_pt += 2 * n_groups;
    n_children = n_groups;
    child_span = span;
    group = MERGE_GROUP_UPPER;
  }
}

// ---------------------------------------------------------------------------------------------
// bucket reduction.  Slot i of the bucket array holds the bucket of VALUE i + 1 (zero digits never reach the
// sort), so per window we need T = sum_i (i+1) A[i].  Chunks of s entries:
//     T = sum_k R_k + s * sum_k k * S_k,   R_k = sum_{t<s} (t+1) * A[ks+t],  S_k = sum_{t<s} A[ks+t]
// and the second term is the plain weighted sum (weights k = 0, 1, ...) of the array s * S_k (one entry per
// chunk): a recursion of log_s(n) levels, one launch each, every level with thousands of independent running
// sums.  Every level hands the NEXT level its chunk sums already multiplied by its own chunk size (log2 s
// doublings per thread), so the weighted sums R need no rescaling, and the level results are folded as they go:
//     V^l_k = R^l_k + sum_{j in chunk k} V^{l-1}_j
// so the window total is the single V of the top level.
// one chunk of one level (see above); t = chunk index over all windows.  QUAD: executed by an aligned quad of lanes
// (ec<C>::add_quad), the stores by its first lane
template <class C, bool QUAD>
__device__ __forceinline__ void reduce_chunk(uint32_t t, bool writer, const XyzzM<C>* A, const XyzzM<C>* Vin,   // no __restrict__: the
                                             uint32_t n, uint32_t a_stride, int cbits, uint32_t nfine, uint32_t s,   // fused kernel reads what
                                             uint32_t nchunks, int first, int out_shift, XyzzM<C>* Sout,           // its previous level wrote
                                             XyzzM<C>* Vout) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint32_t w = t / nchunks, k = t % nchunks;
  const XyzzM<C>* a = A + (uint64_t)w * a_stride;
  const uint32_t cmask = cbits >= 0 ? (1u << cbits) - 1 : 0;
  auto slot = [&](uint32_t i) { return cbits >= 0 ? (i & cmask) * nfine + (i >> cbits) : i; };
  uint32_t lo = k * s, hi = lo + s;
  if (hi > n) hi = n;
  auto plus = [&](XYZZ<C>& x, const XYZZ<C>& y) { if (QUAD) G::add_quad(x, y); else G::add(x, y); };
  XYZZ<C> S = G::infinity(), R = G::infinity();
  for (uint32_t i = hi - 1; i > lo; i--) {
    XYZZ<C> v = D::load_xyzz(a + slot(i));
    plus(S, v);
    plus(R, S);
  }
  {
    XYZZ<C> v = D::load_xyzz(a + slot(lo));
    plus(S, v);
  }
  if (first) plus(R, S);
  if (Vin) {
    const XyzzM<C>* vin = Vin + (uint64_t)w * n;
    for (uint32_t i = lo; i < hi; i++) {
      XYZZ<C> v = D::load_xyzz(vin + i);
      plus(R, v);
    }
  }
  if (writer) D::store_xyzz(Vout + t, R);
  for (int d = 0; d < out_shift; d++) S = G::dbl(S);
  if (writer) D::store_xyzz(Sout + t, S);
}

template <class C, bool QUAD>
__global__ void __launch_bounds__(128, 2)
k_reduce_level(const XyzzM<C>* __restrict__ A, const XyzzM<C>* __restrict__ Vin, uint32_t n, uint32_t a_stride,
               int cbits /* >= 0: entry i lives in slot (i & (2^cbits - 1)) * nfine + (i >> cbits) */, uint32_t nfine,
               uint32_t s, uint32_t nchunks, int W, int first /* level 0: weights t+1 instead of t */,
               int out_shift /* log2(s), 0 at the top level: Sout = 2^out_shift * chunk sum */,
               XyzzM<C>* __restrict__ Sout, XyzzM<C>* __restrict__ Vout) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool writer = !QUAD || (threadIdx.x & 3) == 0;
  if (QUAD) t >>= 2;   // four lanes per chunk (ec<C>::add_quad): the latency-bound upper levels
  if (t >= (uint32_t)W * nchunks) return;
  reduce_chunk<C, QUAD>(t, writer, A, Vin, n, a_stride, cbits, nfine, s, nchunks, first, out_shift, Sout, Vout);
}

template <class C>
__device__ void finish_windows(const XyzzM<C>* win, int W, int c, int raw, uint8_t* result);

// ALL remaining levels of the reduction in ONE launch (one CTA, four lanes per chunk, __syncthreads between levels),
// followed by the window combine + serialisation: used from the level whose chunk count fits the CTA's quads.
template <class C>
__global__ void __launch_bounds__(256, 1)
k_reduce_top(const XyzzM<C>* A, const XyzzM<C>* Vin, uint32_t n, uint32_t a_stride, int cbits, uint32_t nfine, uint32_t s, int W, int first,
             XyzzM<C>* buf0, XyzzM<C>* buf1, int parity, int c, int raw, uint8_t* result) {
  const uint32_t quad = threadIdx.x >> 2, nquads = blockDim.x >> 2;
  const bool writer = (threadIdx.x & 3) == 0;
  const XyzzM<C>* top;
  for (;;) {
    int log_s = 0;
    while ((1u << log_s) < s) log_s++;
    const uint32_t nch = (n + s - 1) / s;
    XyzzM<C>* Sout = parity ? buf1 : buf0;
    XyzzM<C>* Vout = Sout + (size_t)W * nch;
    for (uint32_t t = quad; t < (uint32_t)W * nch; t += nquads)
      reduce_chunk<C, true>(t, writer, A, Vin, n, a_stride, cbits, nfine, s, nch, first, nch == 1 ? 0 : log_s, Sout, Vout);
    __syncthreads();
    top = Vout;
    if (nch == 1) break;
    A = Sout;
    Vin = Vout;
    n = nch;
    a_stride = nch;
    cbits = -1;
    s = 4;
    first = 0;
    parity ^= 1;
  }
  if (threadIdx.x == 0) finish_windows<C>(top, W, c, raw, result);
}

// Horner over the window sums, normalise, serialise
template <class C>
__device__ void finish_windows(const XyzzM<C>* win, int W, int c, int raw, uint8_t* result) {
  typedef dev<C> D;
  typedef ec<C> G;
  XYZZ<C> acc = D::load_xyzz(win + (W - 1));
  for (int w = W - 2; w >= 0; w--) {
    for (int d = 0; d < c; d++) acc = G::dbl(acc);
    XYZZ<C> v = D::load_xyzz(win + w);
    G::add(acc, v);
  }
  if (raw) D::store_result_raw(result, acc);
  else D::store_result(result, acc);
}

template <class C>
__global__ void k_finish(const XyzzM<C>* __restrict__ win, int W, int c, int raw, uint8_t* __restrict__ result) {
  if (blockIdx.x || threadIdx.x) return;
  finish_windows<C>(win, W, c, raw, result);
}

// sum n canonical result records
template <class C>
__global__ void k_combine_results(const uint8_t* __restrict__ recs, int n, int raw, uint8_t* __restrict__ out) {
  typedef dev<C> D;
  typedef ec<C> G;
  if (blockIdx.x || threadIdx.x) return;
  XYZZ<C> acc = G::infinity();
  for (int i = 0; i < n; i++) {
    const uint8_t* r = recs + (size_t)i * 3 * C::FQ_BYTES;
    const uint32_t* zw = reinterpret_cast<const uint32_t*>(r);
    bool zzero = true;
    for (int k = 0; k < D::N; k++) zzero &= zw[k] == 0;
    if (zzero) continue;
    // homogeneous (X:Y:Z) -> XYZZ (X*Z, Y*Z^2, Z^2, Z^3)
    Fe<typename C::Fq> Z = D::load_canonical(r), Y = D::load_canonical(r + C::FQ_BYTES),
                       X = D::load_canonical(r + 2 * C::FQ_BYTES);
    XYZZ<C> p;
    p.ZZ = ff<typename C::Fq>::sqr(Z);
    p.ZZZ = ff<typename C::Fq>::mul(p.ZZ, Z);
    p.X = ff<typename C::Fq>::mul(X, Z);
    p.Y = ff<typename C::Fq>::mul(Y, p.ZZ);
    G::add(acc, p);
  }
  if (raw) D::store_result_raw(out, acc);
  else D::store_result(out, acc);
}

// ---------------------------------------------------------------------------------------------
// bench / test input generator: out[i] = P0 + (first + i) * Q   (affine wire format)
template <class C>
__global__ void __launch_bounds__(128) k_gen_chain(const uint8_t* __restrict__ p0q, uint64_t first, uint64_t n,
                                                   uint8_t* __restrict__ out) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> p0, q;
  p0.x = D::load_canonical(p0q);
  p0.y = D::load_canonical(p0q + C::FQ_BYTES);
  q.x = D::load_canonical(p0q + 2 * C::FQ_BYTES);
  q.y = D::load_canonical(p0q + 3 * C::FQ_BYTES);
  uint64_t k = first + i;
  XYZZ<C> acc = G::infinity();
  for (int bit = 63; bit >= 0; bit--) {
    acc = G::dbl(acc);
    if ((k >> bit) & 1) G::madd(acc, q);
  }
  G::madd(acc, p0);
  Affine<C> a;
  uint8_t* o = out + i * (2 * C::FQ_BYTES);
  if (!G::to_affine(acc, a)) {
    uint32_t* w = reinterpret_cast<uint32_t*>(o);
    for (int j = 0; j < 2 * D::N; j++) w[j] = 0;
    return;
  }
  D::store_canonical(o, a.x);
  D::store_canonical(o + C::FQ_BYTES, a.y);
}

// field self-test: out[i] = a[i] (op) b[i] on canonical little-endian elements of Fq
template <class C>
__global__ void k_field_selftest(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                 uint8_t* __restrict__ out, int n, int op) {
  typedef dev<C> D;
  typedef ff<typename C::Fq> F;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<typename C::Fq> x = D::load_canonical(a + (size_t)i * C::FQ_BYTES),
                     y = D::load_canonical(b + (size_t)i * C::FQ_BYTES), r;
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::sqr(x); break;
    case 4: r = F::inv(x); break;
    default: r = F::neg(x); break;
  }
  D::store_canonical(out + (size_t)i * C::FQ_BYTES, r);
}

// ---------------------------------------------------------------------------------------------
template <class C>
static void ba_bucket_phase_fwd(const MsmPlan& p, const MsmWorkspace& ws, const void* table, cudaStream_t st);   // msm_ba.cuh

template <class C>
struct CurveLaunch {
  static void points_to_mont(const uint8_t* raw, void* table, uint64_t n, cudaStream_t st) {
    if (!n) return;
    k_points_to_mont<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(raw, (AffineT<C>*)table, n);
    g_kernel_launches += 1;
  }
  static cudaStream_t bucket_phase(const MsmPlan& p, const MsmWorkspace& ws, const void* table, cudaStream_t st) {
    const uint32_t ngoff = (uint32_t)p.W * p.nb;
    XyzzM<C>* buckets = (XyzzM<C>*)ws.buckets;
    if (p.batch_affine != 1) cudaMemsetAsync(buckets, 0, (size_t)ngoff * sizeof(XyzzM<C>), st);
    if (p.batch_affine == 1) {
      ba_bucket_phase_fwd<C>(p, ws, table, st);
    } else {
    if (ws.ev_acc0) cudaEventRecord(ws.ev_acc0, st);
    if (p.batch_affine == 2) {
      k_accumulate_ba<C><<<p.ba_ctas, 128, 0, st>>>((const AffineT<C>*)table, ws.sorted, ws.goff, buckets, ws.part_id,
                                                   (XyzzM<C>*)ws.part_pt, p.nseg, p.seg_len, ngoff, p.ba_rounds,
                                                   (uint4*)ws.ba2_scratch, p.ba_cap);
    } else if (p.tma_stage) {
      const size_t smem = (size_t)4 * 2 * 32 * BZ_ACC_RING_STRIDE + 4 * 2 * sizeof(uint64_t);
      k_accumulate_tma<C><<<(unsigned)((p.nseg + 127) / 128), 128, smem, st>>>(
          (const AffineT<C>*)table, ws.sorted, ws.goff, buckets, ws.part_id, (XyzzM<C>*)ws.part_pt, p.nseg, p.seg_len,
          p.nb, ngoff);
    } else {
      k_accumulate<C><<<(unsigned)((p.nseg + 127) / 128), 128, 0, st>>>(
          (const AffineT<C>*)table, ws.sorted, ws.goff, buckets, ws.part_id, (XyzzM<C>*)ws.part_pt, p.nseg, p.seg_len,
          p.nb, ngoff);
    }
    if (ws.ev_acc1) cudaEventRecord(ws.ev_acc1, st);
    {
      // merge tree over the partial list (see k_merge_level)
      const uint32_t* in_id = ws.part_id;
      const XyzzM<C>* in_pt = (const XyzzM<C>*)ws.part_pt;
      uint32_t* out_id = ws.part2_id;
      XyzzM<C>* out_pt = (XyzzM<C>*)ws.part2_pt;
      uint64_t n_children = p.nseg, child_span = p.seg_len;
      for (int level = 0;; level++) {
        const uint32_t group = merge_group(level, p.nseg);
        const uint64_t span = child_span * group;   // sorted positions covered by one group
        uint64_t n_groups = (n_children + group - 1) / group;
        if (n_groups <= 16) {   // the rest of the tree in one launch
          k_merge_top<C><<<1, 256, 0, st>>>(ws.goff, ngoff, buckets, in_id, in_pt, n_children, group, child_span, out_id, out_pt);
          g_kernel_launches += 1;
          break;
        }
        if (n_groups <= 16384)   // too few walks to fill the machine: four lanes per walk
          k_merge_level<C, true><<<(unsigned)((4 * n_groups + 127) / 128), 128, 0, st>>>(ws.goff, ngoff, buckets, in_id, in_pt, n_children,
                                                                                         group, span, out_id, out_pt, n_groups);
        else
          k_merge_level<C, false><<<(unsigned)((n_groups + 127) / 128), 128, 0, st>>>(ws.goff, ngoff, buckets, in_id, in_pt, n_children,
                                                                                      group, span, out_id, out_pt, n_groups);
        g_kernel_launches += 1;
        if (n_groups == 1) break;
        in_id = out_id;
        in_pt = out_pt;
        out_id += 2 * n_groups;
        out_pt += 2 * n_groups;
        n_children = n_groups;
        child_span = span;
      }
    }
    }   // !batch_affine
    // multi-level running-sum reduction (see k_reduce_level); scratch: red_a = S / V of even levels, red_b = odd
    g_kernel_launches += 1;   // accumulate
    // level 0 uses p.chunk entries per thread (throughput); the upper levels are tiny and latency-bound,
    // so they use chunks of 4 (total chain length ~ s log_s n is shortest for small s)
    const XyzzM<C>* A = buckets;
    const XyzzM<C>* Vin = nullptr;
    uint32_t n = p.nvalues, a_stride = p.nb;
    int perm_bits = p.rest;   // level 0 reads value i at its sort slot (i mod 2^rest) * 2^fb + (i >> rest)
    XyzzM<C>* scratch[2] = {(XyzzM<C>*)ws.red_a, (XyzzM<C>*)ws.red_b};
    int level = 0;
    const XyzzM<C>* top = nullptr;
    // the previous task's tail (other stream) may still be reading the reduction scratch and ws.result
    if (ws.tail && ws.tail_busy) cudaStreamWaitEvent(st, ws.ev_tail_done, 0);
    while (true) {
      const uint32_t s = level == 0 ? p.chunk : 4;
      int log_s = 0;
      while ((1u << log_s) < s) log_s++;
      uint32_t nch = (n + s - 1) / s;
      XyzzM<C>* Sout = scratch[level & 1];
      XyzzM<C>* Vout = Sout + (size_t)p.W * nch;
      uint32_t nt = (uint32_t)p.W * nch;
      if (nt <= 64) {   // one chunk per quad of one CTA: all remaining levels + the window combine in one launch (measured:
                        // no faster than separate launches at 2^21 buckets -- the levels are chain-latency bound, ~0.12 ms
                        // each -- but it saves the launches; with more chunks than quads it is slower)
        k_reduce_top<C><<<1, 256, 0, st>>>(A, Vin, n, a_stride, perm_bits, 1u << p.fb, s, p.W, level == 0 ? 1 : 0, scratch[0], scratch[1],
                                           level & 1, p.c, p.raw_result, ws.result);
        g_kernel_launches += 1;
        return st;
      }
      if (nt <= 8192)   // latency-bound level: four lanes per chunk (measured at 2^21 buckets: 32768 chunks are still
                        // throughput-bound -- 0.22 ms with one thread per chunk, 0.54 ms with four)
        k_reduce_level<C, true><<<(4 * nt + 127) / 128, 128, 0, st>>>(A, Vin, n, a_stride, perm_bits, 1u << p.fb, s, nch, p.W,
                                                                      level == 0 ? 1 : 0, nch == 1 ? 0 : log_s, Sout, Vout);
      else
        k_reduce_level<C, false><<<(nt + 127) / 128, 128, 0, st>>>(A, Vin, n, a_stride, perm_bits, 1u << p.fb, s, nch, p.W,
                                                                   level == 0 ? 1 : 0, nch == 1 ? 0 : log_s, Sout, Vout);
      g_kernel_launches += 1;
      top = Vout;
      if (nch == 1) break;
      if (level == 0 && ws.tail) {
        // everything from here on is a few thousand additions on long dependency chains: continue on the tail stream
        // (higher priority: its few CTAs are placed as soon as slots free up) and let the work stream start the next task
        cudaEventRecord(ws.ev_fork, st);
        cudaStreamWaitEvent(ws.tail, ws.ev_fork, 0);
        st = ws.tail;
      }
      A = Sout;
      Vin = Vout;
      n = nch;
      a_stride = nch;
      perm_bits = -1;
      level++;
    }
    k_finish<C><<<1, 32, 0, st>>>(top, p.W, p.c, p.raw_result, ws.result);
    g_kernel_launches += 1;
    return st;
  }
  static void build_wtable(void* wtable, uint64_t n, int levels, int c, cudaStream_t st) {
    constexpr int K = 16;
    AffineT<C>* t = (AffineT<C>*)wtable;
    const uint64_t threads = (n + K - 1) / K;
    for (int w = 1; w < levels; w++) {
      k_wtable_level<C, K><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(t + (uint64_t)(w - 1) * n, t + (uint64_t)w * n, n, c);
      g_kernel_launches += 1;
    }
  }
  static void table_to_wire(const void* table, uint64_t n, int levels, uint8_t* out, cudaStream_t st) {
    const uint64_t total = n * (uint64_t)levels;
    if (!total) return;
    k_table_to_wire<C><<<(unsigned)((total + 127) / 128), 128, 0, st>>>((const AffineT<C>*)table, n, levels, out);
    g_kernel_launches += 1;
  }
  static void combine_results(const uint8_t* recs, int n, uint8_t* out, int raw, cudaStream_t st) {
    k_combine_results<C><<<1, 32, 0, st>>>(recs, n, raw, out);
  }
  static void gen_chain_points(const uint8_t* p0q, uint64_t first, uint64_t n, uint8_t* out, cudaStream_t st) {
    if (!n) return;
    k_gen_chain<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p0q, first, n, out);
  }
  static void field_selftest(const uint8_t* a, const uint8_t* b, uint8_t* out, int n, int op, cudaStream_t st) {
    k_field_selftest<C><<<(n + 63) / 64, 64, 0, st>>>(a, b, out, n, op);
  }
  static const CurveOps* ops() {
    static const CurveOps o = {C::CODE,
                               C::FQ_BYTES,
                               C::SCALAR_BITS,
                               fr_mod_host(),
                               sizeof(AffineT<C>),
                               sizeof(AffineM<C>),
                               sizeof(XyzzM<C>),
                               &points_to_mont,
                               &bucket_phase,
                               &build_wtable,
                               &table_to_wire,
                               &combine_results,
                               &gen_chain_points,
                               &field_selftest};
    return &o;
  }
  static const uint32_t* fr_mod_host();
};

}  // namespace bz
