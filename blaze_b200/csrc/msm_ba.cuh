// Bucket accumulation by BATCHED AFFINE addition (alternative to k_accumulate's XYZZ sweep).
//
// An affine addition costs 1 inversion + 2M + 1S; with Montgomery's trick the inversion is shared by
// hundreds of millions of independent additions and amortises to 3 extra products, i.e. ~6 field
// products per addition instead of the 10 of an XYZZ mixed add -- the accumulate phase is bound by
// the integer multiplier, so that is the lever that matters.
//
// Independence comes from summing every bucket as a balanced tree.  Round r halves every bucket:
// entries (2j, 2j+1) of the round-r list of bucket g are added into entry j of its round-(r+1) list
// (an odd leftover is copied).  Lists live in closed-form slots, no per-round prefix sums:
//     k_r[g] = ceil(k_0[g] / 2^r)                         entries of bucket g at round r
//     s_r[g] = (goff[g] >> r) + g                         its first slot (strictly increasing in g)
// Round 0 reads the points through the sorted index list (k_0, goff as produced by msm_sort.cu);
// later rounds ping-pong between two affine buffers.  After r* = ceil(log2(max k_0)) rounds every
// bucket holds at most one point.
//
// One round = three phases over the S = (total >> (r+1)) + G output slots; thread t of a warp owns
// slots base + i*32 + lane (coalesced), i < BA_L:
//   k_ba_forward   denominators (x2-x1, or 2y / 1 for the special cases) -> running products
//                  pref[slot], thread totals tot[t]
//   batch inverse  of tot[]: hierarchical (groups of 256 per thread, top level by Fermat) -> itot[]
//   k_ba_backward  walks the thread's slots in reverse, peeling 1/den off the running inverse, and
//                  finishes each addition (lambda, x3, y3) into the output list
#pragma once
#include <cuda_runtime.h>

#include "msm_curve.cuh"

namespace bz {

#define BA_L 16          // slots per thread
#define BA_GROUP 256     // elements per thread in the batch-inverse tree

template <class C>
struct ba {
  typedef typename C::Fq Fq;
  typedef ff<Fq> F;
  typedef Fe<Fq> E;
  typedef ec<C> G;
  typedef dev<C> D;
  static constexpr int N = Fq::N;

  struct Round {
    const AffineT<C>* table;    // round 0 inputs: Montgomery table + sorted refs
    const uint32_t* sorted;
    const uint32_t* goff;       // bucket offsets of round 0 (ngoff + 1 entries)
    uint32_t ngoff;
    const AffineM<C>* in;       // round >= 1 inputs (list r)
    AffineM<C>* out;            // list r + 1
    int r;
    uint32_t nslots;            // (total >> (r+1)) + ngoff
  };

  __device__ __forceinline__ static E ld_fe(const uint32_t* p) {
    E r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int k = 0; k < N / 4; k++) {
      uint4 v = q[k];
      r.v[4 * k] = v.x; r.v[4 * k + 1] = v.y; r.v[4 * k + 2] = v.z; r.v[4 * k + 3] = v.w;
    }
    return r;
  }
  __device__ __forceinline__ static void st_fe(uint32_t* p, const E& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int k = 0; k < N / 4; k++) q[k] = make_uint4(r.v[4 * k], r.v[4 * k + 1], r.v[4 * k + 2], r.v[4 * k + 3]);
  }
  __device__ __forceinline__ static void st_affine(AffineM<C>* p, const Affine<C>& a) {
    st_fe(p->x, a.x);
    st_fe(p->y, a.y);
  }

  // first slot of bucket g in list rr (rr >= 1)
  __device__ __forceinline__ static uint32_t slot0(const Round& R, uint32_t g, int rr) {
    return (__ldg(R.goff + g) >> rr) + g;
  }
  // largest g with slot0(g, r+1) <= p
  __device__ static uint32_t find_bucket(const Round& R, uint32_t p) {
    uint32_t lo = 0, hi = R.ngoff;
    while (hi - lo > 1) {
      uint32_t mid = lo + ((hi - lo) >> 1);
      if (slot0(R, mid, R.r + 1) <= p) lo = mid; else hi = mid;
    }
    return lo;
  }

  // records are addressed by the pointer to their x[0] (table entries are padded, list entries packed)
  __device__ __forceinline__ static E load_x(const uint32_t* rec) { return ld_fe(rec); }
  __device__ __forceinline__ static E load_y(const uint32_t* rec, bool negate) {
    E y = ld_fe(rec + N);
    return negate ? F::neg(y) : y;
  }
  // Forward-pass variant: where the two inputs of output slot p live (null src2 = no pair); returns like fetch.
  __device__ __forceinline__ static int locate(const Round& R, uint32_t p, uint32_t g, const uint32_t*& s1, bool& n1,
                                               const uint32_t*& s2, bool& n2) {
    const uint32_t g0 = __ldg(R.goff + g), g1 = __ldg(R.goff + g + 1);
    const uint32_t k0 = g1 - g0;
    const uint32_t j = p - ((g0 >> (R.r + 1)) + g);
    const uint32_t kr = (uint32_t)(((uint64_t)k0 + ((1ull << R.r) - 1)) >> R.r);
    const uint32_t kr1 = (kr + 1) >> 1;
    if (j >= kr1) return 0;
    const bool pair = 2 * j + 1 < kr;
    n1 = n2 = false;
    if (R.r == 0) {
      uint32_t e = __ldg(R.sorted + g0 + 2 * j);
      s1 = R.table[e & 0x7fffffffu].x;
      n1 = (e >> 31) != 0;
      if (pair) {
        e = __ldg(R.sorted + g0 + 2 * j + 1);
        s2 = R.table[e & 0x7fffffffu].x;
        n2 = (e >> 31) != 0;
      }
    } else {
      s1 = R.in[((g0 >> R.r) + g) + 2 * j].x;
      s2 = s1 + 2 * N;
    }
    return pair ? 2 : 1;
  }

  // What output slot p of bucket g computes: returns 0 = nothing (gap), 1 = copy of p1, 2 = p1 + p2.
  __device__ __forceinline__ static int fetch(const Round& R, uint32_t p, uint32_t g, Affine<C>& p1, Affine<C>& p2) {
    const uint32_t g0 = __ldg(R.goff + g), g1 = __ldg(R.goff + g + 1);
    const uint32_t k0 = g1 - g0;
    const uint32_t j = p - ((g0 >> (R.r + 1)) + g);
    const uint32_t kr = (uint32_t)(((uint64_t)k0 + ((1ull << R.r) - 1)) >> R.r);
    const uint32_t kr1 = (kr + 1) >> 1;
    if (j >= kr1) return 0;
    const bool pair = 2 * j + 1 < kr;
    if (R.r == 0) {
      uint32_t e = __ldg(R.sorted + g0 + 2 * j);
      p1 = D::load_affine(R.table[e & 0x7fffffffu].x);
      if (e & 0x80000000u) p1.y = F::neg(p1.y);
      if (pair) {
        e = __ldg(R.sorted + g0 + 2 * j + 1);
        p2 = D::load_affine(R.table[e & 0x7fffffffu].x);
        if (e & 0x80000000u) p2.y = F::neg(p2.y);
      }
    } else {
      const AffineM<C>* src = R.in + ((g0 >> R.r) + g) + 2 * j;
      p1 = D::load_affine(src->x);
      if (pair) p2 = D::load_affine(src[1].x);
    }
    return pair ? 2 : 1;
  }
};

// ---------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) k_ba_forward(typename ba<C>::Round R, uint32_t* __restrict__ pref,
                                                    uint32_t* __restrict__ tot) {
  typedef ba<C> B;
  typedef typename B::F F;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = t >> 5, lane = t & 31;
  const uint32_t base = warp * (32 * BA_L) + lane;
  typename B::E run = F::one();
  if (base < R.nslots) {
    uint32_t g = B::find_bucket(R, base);
    for (int i = 0; i < BA_L; i++) {
      uint32_t p = base + i * 32;
      if (p >= R.nslots) break;
      while (g + 1 < R.ngoff && B::slot0(R, g + 1, R.r + 1) <= p) g++;
      const uint32_t *s1 = nullptr, *s2 = nullptr;
      bool n1, n2;
      int what = B::locate(R, p, g, s1, n1, s2, n2);
      if (what == 2) {
        // the x coordinates decide the generic case (both non-zero and different => chord, den = x2 - x1,
        // exactly what ba_classify returns); only the rare rest needs the y coordinates as well
        typename B::E x1 = B::load_x(s1), x2 = B::load_x(s2);
        typename B::E den = F::sub(x2, x1);
        if (F::is_zero(x1) || F::is_zero(x2) || F::is_zero(den)) {
          Affine<C> p1, p2;
          p1.x = x1; p1.y = B::load_y(s1, n1);
          p2.x = x2; p2.y = B::load_y(s2, n2);
          ec<C>::ba_classify(p1, p2, den);
        }
        run = F::mul(run, den);
      }
      B::st_fe(pref + (size_t)p * B::N, run);
    }
  }
  B::st_fe(tot + (size_t)t * B::N, run);
}

template <class C>
__global__ void __launch_bounds__(128) k_ba_backward(typename ba<C>::Round R, const uint32_t* __restrict__ pref,
                                                     const uint32_t* __restrict__ itot) {
  typedef ba<C> B;
  typedef typename B::F F;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = t >> 5, lane = t & 31;
  const uint32_t base = warp * (32 * BA_L) + lane;
  if (base >= R.nslots) return;
  int last = BA_L - 1;
  while (base + (uint32_t)last * 32 >= R.nslots) last--;
  typename B::E inv = B::ld_fe(itot + (size_t)t * B::N);   // 1 / (product of this thread's denominators)
  uint32_t g = B::find_bucket(R, base + (uint32_t)last * 32);
  for (int i = last; i >= 0; i--) {
    uint32_t p = base + i * 32;
    while (B::slot0(R, g, R.r + 1) > p) g--;
    Affine<C> p1, p2;
    int what = B::fetch(R, p, g, p1, p2);
    if (what == 0) continue;
    if (what == 1) { B::st_affine(R.out + p, p1); continue; }
    typename B::E den;
    int kind = ec<C>::ba_classify(p1, p2, den);
    // 1/den = inv * (product of the denominators before this one)
    typename B::E dinv = i > 0 ? F::mul(inv, B::ld_fe(pref + (size_t)(p - 32) * B::N)) : inv;
    inv = F::mul(inv, den);
    B::st_affine(R.out + p, ec<C>::ba_finish(kind, p1, p2, dinv));
  }
}

// ---- batch inverse of n field elements: up-sweep (group products), top level by Fermat, down-sweep ----
template <class C>
__global__ void __launch_bounds__(128) k_binv_up(const uint32_t* __restrict__ X, uint32_t n, uint32_t* __restrict__ P,
                                                 uint32_t* __restrict__ Gout, uint32_t ngroups) {
  typedef ba<C> B;
  typedef typename B::F F;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ngroups) return;
  uint32_t lo = j * BA_GROUP, hi = lo + BA_GROUP < n ? lo + BA_GROUP : n;
  typename B::E run = F::one();
  for (uint32_t i = lo; i < hi; i++) {
    run = F::mul(run, B::ld_fe(X + (size_t)i * B::N));
    B::st_fe(P + (size_t)i * B::N, run);
  }
  B::st_fe(Gout + (size_t)j * B::N, run);
}
template <class C>
__global__ void __launch_bounds__(128) k_binv_direct(uint32_t* __restrict__ X, uint32_t n) {
  typedef ba<C> B;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  B::st_fe(X + (size_t)j * B::N, B::F::inv(B::ld_fe(X + (size_t)j * B::N)));
}
// Y (written over P) [i] = 1 / X[i], given IG[j] = 1 / (product of group j)
template <class C>
__global__ void __launch_bounds__(128) k_binv_down(const uint32_t* __restrict__ X, uint32_t n, uint32_t* __restrict__ P,
                                                   const uint32_t* __restrict__ IG, uint32_t ngroups) {
  typedef ba<C> B;
  typedef typename B::F F;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ngroups) return;
  uint32_t lo = j * BA_GROUP, hi = lo + BA_GROUP < n ? lo + BA_GROUP : n;
  typename B::E run = B::ld_fe(IG + (size_t)j * B::N);
  for (uint32_t i = hi; i-- > lo;) {
    typename B::E x = B::ld_fe(X + (size_t)i * B::N);
    typename B::E y = i > lo ? F::mul(run, B::ld_fe(P + (size_t)(i - 1) * B::N)) : run;
    run = F::mul(run, x);
    B::st_fe(P + (size_t)i * B::N, y);
  }
}

// largest bucket size (bucket 0 of every window is empty: zero digits are dropped by the sort)
static __global__ void k_ba_maxk(const uint32_t* __restrict__ goff, uint32_t ngoff, uint32_t* __restrict__ out) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t k = g < ngoff ? goff[g + 1] - goff[g] : 0;
  for (int o = 16; o; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
  if ((threadIdx.x & 31) == 0 && k) atomicMax(out, k);
}

// buckets[g] (XYZZ) = the single remaining point of bucket g after r* rounds
template <class C>
__global__ void __launch_bounds__(128) k_ba_gather(const AffineT<C>* __restrict__ table, const uint32_t* __restrict__ sorted,
                                                   const uint32_t* __restrict__ goff, uint32_t ngoff,
                                                   const AffineM<C>* __restrict__ list, int rstar,
                                                   XyzzM<C>* __restrict__ buckets) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngoff) return;
  uint32_t g0 = goff[g], k0 = goff[g + 1] - g0;
  XYZZ<C> r = G::infinity();
  if (k0) {
    Affine<C> a;
    if (rstar == 0) {
      uint32_t e = sorted[g0];
      a = D::load_affine(table[e & 0x7fffffffu].x);
      if (e & 0x80000000u) a.y = ff<typename C::Fq>::neg(a.y);
    } else {
      a = D::load_affine(list[(g0 >> rstar) + g].x);
    }
    r = G::from_affine(a);
  }
  D::store_xyzz(buckets + g, r);
}

// host side: all rounds of the batched-affine accumulation.  Needs two small device->host reads
// (total number of sorted entries, largest bucket) once the sort has finished.
template <class C>
static void ba_bucket_phase(const MsmPlan& p, const MsmWorkspace& ws, const void* table, cudaStream_t st) {
  typedef ba<C> B;
  const uint32_t ngoff = (uint32_t)p.W * p.nb;
  XyzzM<C>* buckets = (XyzzM<C>*)ws.buckets;
  uint32_t* d_maxk = ws.ba_scalars;   // [0] = max k, device
  cudaMemsetAsync(d_maxk, 0, 4, st);
  k_ba_maxk<<<(ngoff + 255) / 256, 256, 0, st>>>(ws.goff, ngoff, d_maxk);
  uint32_t h[2] = {0, 0};
  cudaMemcpyAsync(&h[0], d_maxk, 4, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&h[1], ws.goff + ngoff, 4, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  const uint32_t maxk = h[0], total = h[1];
  int rstar = 0;
  while ((1ull << rstar) < maxk) rstar++;
  g_kernel_launches += 2;
  AffineM<C>* buf[2] = {(AffineM<C>*)ws.ba_buf0, (AffineM<C>*)ws.ba_buf1};   // list r lives in buf[r & 1]; list 1 is the largest
  if (ws.ev_acc0) cudaEventRecord(ws.ev_acc0, st);
  for (int r = 0; r < rstar; r++) {
    typename B::Round R;
    R.table = (const AffineT<C>*)table;
    R.sorted = ws.sorted;
    R.goff = ws.goff;
    R.ngoff = ngoff;
    R.in = buf[r & 1];              // list r   (unused for r = 0)
    R.out = buf[(r + 1) & 1];       // list r+1
    R.r = r;
    R.nslots = (total >> (r + 1)) + ngoff;
    const uint32_t nthreads = ((R.nslots + 32 * BA_L - 1) / (32 * BA_L)) * 32;
    const uint32_t nblocks = (nthreads + 127) / 128;
    const uint32_t ntot = nblocks * 128;   // every launched thread writes its total
    k_ba_forward<C><<<nblocks, 128, 0, st>>>(R, ws.ba_pref, ws.ba_tot);
    // batch inverse of ba_tot[0..ntot) -> ba_itot
    uint32_t* X[4] = {ws.ba_tot, ws.ba_lvl[0], ws.ba_lvl[1], ws.ba_lvl[2]};
    uint32_t* Pp[3] = {ws.ba_itot, ws.ba_lvlp[0], ws.ba_lvlp[1]};
    uint32_t n[4];
    n[0] = ntot;
    int levels = 0;
    while (n[levels] > 1024 && levels < 3) {
      n[levels + 1] = (n[levels] + BA_GROUP - 1) / BA_GROUP;
      k_binv_up<C><<<(n[levels + 1] + 127) / 128, 128, 0, st>>>(X[levels], n[levels], Pp[levels], X[levels + 1], n[levels + 1]);
      levels++;
    }
    k_binv_direct<C><<<(n[levels] + 127) / 128, 128, 0, st>>>(X[levels], n[levels]);
    for (int l = levels - 1; l >= 0; l--) {
      // inverses of the group products of level l: the top array was inverted in place, lower ones
      // come out of the previous down-sweep (written over that level's prefix array)
      const uint32_t* IG = (l + 1 == levels) ? X[l + 1] : Pp[l + 1];
      k_binv_down<C><<<(n[l + 1] + 127) / 128, 128, 0, st>>>(X[l], n[l], Pp[l], IG, n[l + 1]);
    }
    const uint32_t* itot = levels ? Pp[0] : X[0];
    k_ba_backward<C><<<nblocks, 128, 0, st>>>(R, ws.ba_pref, itot);
    g_kernel_launches += 3 + 2 * levels;
  }
  if (ws.ev_acc1) cudaEventRecord(ws.ev_acc1, st);
  k_ba_gather<C><<<(ngoff + 127) / 128, 128, 0, st>>>((const AffineT<C>*)table, ws.sorted, ws.goff, ngoff, buf[rstar & 1], rstar,
                                                    buckets);
  g_kernel_launches += 1;
}

template <class C>
static void ba_bucket_phase_fwd(const MsmPlan& p, const MsmWorkspace& ws, const void* table, cudaStream_t st) {
  ba_bucket_phase<C>(p, ws, table, st);
}

}  // namespace bz
