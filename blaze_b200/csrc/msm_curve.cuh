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

namespace bz {

template <class C>
struct alignas(16) AffineM {
  uint32_t x[C::Fq::N], y[C::Fq::N];
};
template <class C>
struct alignas(16) XyzzM {
  uint32_t X[C::Fq::N], Y[C::Fq::N], ZZ[C::Fq::N], ZZZ[C::Fq::N];
};

template <class C>
struct dev {
  typedef typename C::Fq Fq;
  static constexpr int N = Fq::N;
  typedef ff<Fq> F;
  typedef ec<C> G;

  __device__ __forceinline__ static Affine<C> load_affine(const AffineM<C>* p) {
    Affine<C> a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
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
__global__ void __launch_bounds__(128) k_points_to_mont(const uint8_t* __restrict__ raw, AffineM<C>* __restrict__ table,
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
// bucket accumulation (the dominant kernel)
template <class C>
__global__ void __launch_bounds__(128, 2)
k_accumulate(const AffineM<C>* __restrict__ table, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ goff, XyzzM<C>* __restrict__ buckets, uint32_t* __restrict__ part_id,
             XyzzM<C>* __restrict__ part_pt, uint64_t total, uint32_t L, uint32_t nb, uint32_t ngoff) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s64 = t * L;
  if (s64 >= total) return;
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
  bool skip = (g % nb) == 0;
  uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu;
  XYZZ<C> acc = G::infinity();

  for (uint32_t pos = s; pos < e; pos++) {
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
      skip = (g % nb) == 0;
      acc = G::infinity();
    }
    if (!skip) {
      uint32_t ent = __ldg(sorted + pos);
      Affine<C> a = D::load_affine(table + (ent & 0x7fffffffu));
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

// fold the partial sums of buckets that straddle segment boundaries
template <class C>
__global__ void __launch_bounds__(128) k_merge_partials(const uint32_t* __restrict__ goff,
                                                        XyzzM<C>* __restrict__ buckets,
                                                        const uint32_t* __restrict__ part_id,
                                                        const XyzzM<C>* __restrict__ part_pt, uint64_t nseg,
                                                        uint32_t L) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  uint32_t g = part_id[2 * t + 1];
  if (g == 0xffffffffu) return;        // runs start at a tail entry
  XYZZ<C> acc = D::load_xyzz(part_pt + 2 * t + 1);
  uint32_t bend = goff[g + 1];
  for (uint64_t u = t + 1; u < nseg; u++) {
    if (part_id[2 * u] != g) break;
    XYZZ<C> o = D::load_xyzz(part_pt + 2 * u);
    G::add(acc, o);
    if ((uint64_t)bend <= (u + 1) * L) break;   // bucket ends inside segment u
  }
  D::store_xyzz(buckets + g, acc);
}

// ---------------------------------------------------------------------------------------------
// bucket reduction: per window sum_b b * B_b, buckets 1..nb-1 split into chunks of `chunk`
template <class C>
__global__ void __launch_bounds__(128, 2)
k_reduce_chunks(const XyzzM<C>* __restrict__ buckets, XyzzM<C>* __restrict__ out, uint32_t nb, uint32_t chunk,
                uint32_t nchunks, int W) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint32_t)W * nchunks) return;
  uint32_t w = t / nchunks, j = t % nchunks;
  const XyzzM<C>* B = buckets + (uint64_t)w * nb;
  uint32_t lo = j * chunk + 1, hi = lo + chunk - 1;
  XYZZ<C> S = G::infinity(), R = G::infinity();
  for (uint32_t b = hi; b >= lo; b--) {
    XYZZ<C> v = D::load_xyzz(B + b);
    G::add(S, v);
    G::add(R, S);
  }
  // sum_{b in chunk} b*B_b = R + (lo-1)*S
  if (lo > 1) {
    XYZZ<C> m = G::mul_small(S, lo - 1);
    G::add(R, m);
  }
  D::store_xyzz(out + t, R);
}

// out[i] = sum_{k<group} in[i*group + k]
template <class C>
__global__ void __launch_bounds__(128) k_sum_groups(const XyzzM<C>* __restrict__ in, XyzzM<C>* __restrict__ out,
                                                    uint32_t nout, uint32_t group) {
  typedef dev<C> D;
  typedef ec<C> G;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  XYZZ<C> acc = G::infinity();
  for (uint32_t k = 0; k < group; k++) {
    XYZZ<C> v = D::load_xyzz(in + (uint64_t)i * group + k);
    G::add(acc, v);
  }
  D::store_xyzz(out + i, acc);
}

// Horner over the window sums, normalise, serialise
template <class C>
__global__ void k_finish(const XyzzM<C>* __restrict__ win, int W, int c, uint8_t* __restrict__ result) {
  typedef dev<C> D;
  typedef ec<C> G;
  if (blockIdx.x || threadIdx.x) return;
  XYZZ<C> acc = D::load_xyzz(win + (W - 1));
  for (int w = W - 2; w >= 0; w--) {
    for (int d = 0; d < c; d++) acc = G::dbl(acc);
    XYZZ<C> v = D::load_xyzz(win + w);
    G::add(acc, v);
  }
  D::store_result(result, acc);
}

// sum n canonical result records
template <class C>
__global__ void k_combine_results(const uint8_t* __restrict__ recs, int n, uint8_t* __restrict__ out) {
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
  D::store_result(out, acc);
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
struct CurveLaunch {
  static void points_to_mont(const uint8_t* raw, void* table, uint64_t n, cudaStream_t st) {
    if (!n) return;
    k_points_to_mont<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(raw, (AffineM<C>*)table, n);
    g_kernel_launches += 1;
  }
  static void bucket_phase(const MsmPlan& p, const MsmWorkspace& ws, const void* table, cudaStream_t st) {
    const uint64_t total = (uint64_t)p.W * p.M;
    const uint32_t ngoff = (uint32_t)p.W * p.nb;
    XyzzM<C>* buckets = (XyzzM<C>*)ws.buckets;
    cudaMemsetAsync(buckets, 0, (size_t)ngoff * sizeof(XyzzM<C>), st);
    if (ws.ev_acc0) cudaEventRecord(ws.ev_acc0, st);
    k_accumulate<C><<<(unsigned)((p.nseg + 127) / 128), 128, 0, st>>>(
        (const AffineM<C>*)table, ws.sorted, ws.goff, buckets, ws.part_id, (XyzzM<C>*)ws.part_pt, total, p.seg_len,
        p.nb, ngoff);
    if (ws.ev_acc1) cudaEventRecord(ws.ev_acc1, st);
    k_merge_partials<C><<<(unsigned)((p.nseg + 127) / 128), 128, 0, st>>>(ws.goff, buckets, ws.part_id,
                                                                          (const XyzzM<C>*)ws.part_pt, p.nseg,
                                                                          p.seg_len);
    uint32_t n = (uint32_t)p.W * p.nchunks;
    XyzzM<C>* a = (XyzzM<C>*)ws.red_a;
    XyzzM<C>* b = (XyzzM<C>*)ws.red_b;
    k_reduce_chunks<C><<<(n + 127) / 128, 128, 0, st>>>(buckets, a, p.nb, p.chunk, p.nchunks, p.W);
    // tree-sum the chunk results of each window down to one point per window
    g_kernel_launches += 4;   // accumulate, merge, reduce_chunks, finish
    uint32_t per = p.nchunks;
    while (per > 1) {
      g_kernel_launches += 1;
      uint32_t group = per > 64 ? 64 : per;
      while (per % group) group--;   // per is a power of two, so this never iterates
      uint32_t nper = per / group;
      uint32_t nout = (uint32_t)p.W * nper;
      k_sum_groups<C><<<(nout + 127) / 128, 128, 0, st>>>(a, b, nout, group);
      XyzzM<C>* tmp = a; a = b; b = tmp;
      per = nper;
    }
    k_finish<C><<<1, 32, 0, st>>>(a, p.W, p.c, ws.result);
  }
  static void combine_results(const uint8_t* recs, int n, uint8_t* out, cudaStream_t st) {
    k_combine_results<C><<<1, 32, 0, st>>>(recs, n, out);
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
                               sizeof(AffineM<C>),
                               sizeof(XyzzM<C>),
                               &points_to_mont,
                               &bucket_phase,
                               &combine_results,
                               &gen_chain_points,
                               &field_selftest};
    return &o;
  }
  static const uint32_t* fr_mod_host();
};

}  // namespace bz
