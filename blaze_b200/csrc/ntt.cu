// Number-theoretic transform over the BLS12-381 scalar field (and the other two Fr), sm_100a.
//
// Black box being replaced: the FPGA NTT core behind NTTClient
// (/root/reference/src/ingo_ntt/ntt_api.rs:58-124; fixed size 2^27, ntt_data.rs:65-66).  The
// reference never states field / root / ordering (golden files are external); BASELINE.json fixes
// BLS12-381 Fr with arkworks `Radix2EvaluationDomain::fft` semantics: natural order in and out,
// out[k] = sum_j in[j] w^(jk), w = g^((r-1)/2^32)^(2^(32-log n)).
//
// Algorithm: Stockham autosort, one global-memory pass per radix R = 2^lr (lr <= 9), ping-pong
// buffers.  Pass with sub-transform size Ns:  for j in [0, L/R):
//     x[r]   = in[j + r L/R] * w_{Ns R}^{r (j mod Ns)}          r in [0, R)
//     X      = DFT_R(x)
//     out[(j / Ns) Ns R + (j mod Ns) + k Ns] = X[k]
// A CTA owns V = 8 adjacent work items (adjacent in memory => 256-byte runs), stages the R x V
// tile in shared memory and runs the R-point DFT as in-place decimation-in-frequency rounds of
// radix 8 (then 4 or 2), eight points per thread in registers; the digit reversal of the
// in-place DIF is undone for free in the store addressing.
//
// No Montgomery conversion of the data: a Montgomery product of a CANONICAL value with a twiddle
// stored in Montgomery form (w R) is x w R R^-1 = x w, canonical again, and the butterflies' add/sub
// do not care.  So the 32-byte wire elements are transformed as they are.
//
// Twiddles: w^e for any e < 2^log_root from two tables (w^x, x < 2^14; w^(x 2^14)) -- one extra
// product -- built once per (log_root, direction) on the device from the field's 2-adic root.
// Field products INLINED (111 KB of code for the pass kernel): for the 8-limb scalar fields the call marshalling of a
// called product (16-24 IMAD.MOV on the multiplier pipe per 112 IMAD.WIDE) costs more than the instruction-cache misses
// of the unrolled body -- measured on B200: 2^27 37.38 ms with calls, 35.73 ms inlined, 35.98 ms with only the butterflies'
// internal or only the inter-round twiddle products inlined (profiles/r2_ntt_variants.txt).
// The 12-limb base fields of the MSM are the other way round (ff.cuh).  -DNTT_CALL_MUL restores the called form.
#ifdef NTT_CALL_MUL
#define BZ_NOINLINE_MUL 1
#endif
#include <cuda_runtime.h>

#include <cstdint>

#include "ff.cuh"
#include "msm_internal.h"
#include "ntt_internal.h"

namespace bz {

template <class F>
struct nt {
  typedef ff<F> A;
  typedef Fe<F> E;
  static_assert(F::N == 8, "scalar fields are 8 x 32-bit limbs");

  __device__ __forceinline__ static E ld(const uint4* p) {   // element = 2 consecutive uint4
    uint4 a = p[0], b = p[1];
    E r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ static void st(uint4* p, const E& r) {
    p[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    p[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
  // shared-memory tile: lo halves then hi halves, so 8 adjacent lanes hit 8 distinct 16-byte banks
  __device__ __forceinline__ static E lds(const uint4* lo, const uint4* hi, int idx) {
    uint4 a = lo[idx], b = hi[idx];
    E r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ static void sts(uint4* lo, uint4* hi, int idx, const E& r) {
    lo[idx] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[idx] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
  // w^e from the two-level table (Montgomery form)
  __device__ __forceinline__ static E tw(const NttTables& t, uint64_t e) {
    uint32_t lo = (uint32_t)(e & ((1u << t.lo_bits) - 1));
    uint32_t hi = (uint32_t)(e >> t.lo_bits);
    if (lo == 0) return ld(t.hi + 2 * (size_t)hi);   // hi[0] = 1; exponents that are multiples of 2^lo_bits (the
                                                      // per-CTA w_R table of a large transform) need no product
    E a = ld(t.lo + 2 * (size_t)lo);
    if (hi == 0) return a;
    E b = ld(t.hi + 2 * (size_t)hi);
    return A::mul(a, b);
  }
};

// ---------------------------------------------------------------------------------------------
// table generation: lo[x] = w^x (x < 2^lo_bits), hi[x] = w^(x 2^lo_bits), ninv = (2^log_n)^-1,
// all in Montgomery form; w = ROOT^(2^(TWO_ADICITY - log_root)) (or ROOT_INV for the inverse).
template <class F>
__global__ void k_ntt_tables(uint4* lo, uint4* hi, uint4* ninv, int log_root, int lo_bits, int inverse) {
  typedef ff<F> A;
  typedef Fe<F> E;
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nlo = 1u << (log_root < lo_bits ? log_root : lo_bits);
  uint32_t nhi = log_root > lo_bits ? 1u << (log_root - lo_bits) : 1u;
  if (x >= nlo + nhi + 1) return;
  E w;
#pragma unroll
  for (int i = 0; i < 8; i++) w.v[i] = inverse ? F::root_inv()[i] : F::root()[i];
  for (int i = 0; i < F::TWO_ADICITY - log_root; i++) w = A::sqr(w);
  if (x == nlo + nhi) {   // (2^log_n)^-1: 2^-1 = (p+1)/2, raised log_root times
    E two = A::add(A::one(), A::one());
    E i2 = A::inv(two);
    E r = A::one();
    for (int i = 0; i < log_root; i++) r = A::mul(r, i2);
    nt<F>::st(ninv, r);
    return;
  }
  uint64_t e;
  uint4* dst;
  if (x < nlo) { e = x; dst = lo + 2 * (size_t)x; }
  else { e = (uint64_t)(x - nlo) << lo_bits; dst = hi + 2 * (size_t)(x - nlo); }
  E r = A::one();
  for (int bit = 31; bit >= 0; bit--) {
    r = A::sqr(r);
    if ((e >> bit) & 1) r = A::mul(r, w);
  }
  nt<F>::st(dst, r);
}

// ---------------------------------------------------------------------------------------------
template <class F, int B>   // one in-place DIF round of radix 2^B on the shared tile
__device__ __forceinline__ void dif_round(uint4* lo, uint4* hi, const uint4* trl, const uint4* trh, int lr, int lv,
                                          int lblk /*log2 of the current block size*/, int tid, int nthreads) {
  typedef ff<F> A;
  typedef Fe<F> E;
  constexpr int RHO = 1 << B;
  const int V = 1 << lv;
  const int R = 1 << lr;
  const int lsub = lblk - B;
  const int sub = 1 << lsub;            // stride between the points of one butterfly group
  const int items = V * (R / RHO);
  const int tws = R >> lblk;            // w_blk = w_R^tws
  for (int w = tid; w < items; w += nthreads) {
    int v = w & (V - 1), grp = w >> lv;
    int b0 = grp >> lsub, u = grp & (sub - 1);
    int base = (b0 << lblk) + u;
    E x[RHO];
#pragma unroll
    for (int i = 0; i < RHO; i++) x[i] = nt<F>::lds(lo, hi, (base + i * sub) * V + v);
    // radix-2 DIF stages inside the group; w_RHO = w_R^(R/RHO)
#pragma unroll
    for (int span = RHO / 2; span >= 1; span >>= 1) {
#pragma unroll
      for (int h = 0; h < RHO; h += 2 * span) {
#pragma unroll
        for (int i = 0; i < span; i++) {
          E a = A::add(x[h + i], x[h + i + span]);
          E d = A::sub(x[h + i], x[h + i + span]);
          if (i != 0) {
            int e = i * (R / (2 * span));   // w_{2 span}^i
            d = A::mul(d, nt<F>::lds(trl, trh, e));
          }
          x[h + i] = a;
          x[h + i + span] = d;
        }
      }
    }
    // slot i now holds Y[bitrev_B(i)]; twiddle by w_blk^(u m) and store Y[m] at base + m*sub
#pragma unroll
    for (int i = 0; i < RHO; i++) {
      int m = 0;
#pragma unroll
      for (int bb = 0; bb < B; bb++) m |= ((i >> bb) & 1) << (B - 1 - bb);
      E y = x[i];
      if (m != 0 && u != 0 && sub > 1) y = A::mul(y, nt<F>::lds(trl, trh, u * m * tws));
      nt<F>::sts(lo, hi, (base + m * sub) * V + v, y);
    }
  }
}

// CTA shape: 256 threads x 2 CTAs per SM (122 registers).  Measured alternatives on B200 (scripts/ab_probe_ntt.sh):
// see DESIGN.md 7b.
#ifndef NTT_THREADS
#define NTT_THREADS 256
#endif
#ifndef NTT_MINBLOCKS
#define NTT_MINBLOCKS 2
#endif
template <class F>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MINBLOCKS) k_ntt_pass(NttPassParams P) {
  typedef ff<F> A;
  typedef Fe<F> E;
  extern __shared__ uint4 smem[];
  const int R = 1 << P.lr, V = 1 << P.lv;
  uint4* lo = smem;
  uint4* hi = smem + R * V;
  uint4* trl = smem + 2 * R * V;   // w_R^e table, lo halves
  uint4* trh = trl + R;
  const int tid = threadIdx.x, nth = blockDim.x;
  const uint64_t qbase = (uint64_t)blockIdx.x << P.lv;

  // w_R^e = w^(e * root/R)
  for (int e = tid; e < R; e += nth) {
    E t = nt<F>::tw(P.tab, (uint64_t)e << (P.tab.log_root - P.lr));
    nt<F>::sts(trl, trh, e, t);
  }
  // load (lane fastest: V adjacent work items are adjacent in memory), apply the pass twiddle
  for (int idx = tid; idx < R * V; idx += nth) {
    int v = idx & (V - 1), r = idx >> P.lv;
    uint64_t q = qbase + v;
    if (q >= P.Q) continue;
    uint64_t q0 = q & (P.Q0 - 1), qr = q >> P.lq0;
    uint64_t q1 = qr & (P.Q1 - 1), q2 = qr >> P.lq1;
    const uint64_t off = q0 * P.in_s0 + q1 * P.in_s1 + q2 * P.in_s2 + (uint64_t)r * P.in_sr;
    if (P.tw_full_out) {   // table generation: twiddle of this element, in the input layout
      uint64_t tq = P.tw_sel == 0 ? q0 : (P.tw_sel == 1 ? q1 : q2);
      nt<F>::st(P.tw_full_out + 2 * off, nt<F>::tw(P.tab, (uint64_t)r * tq * P.tw_scale));
      continue;
    }
    E x = nt<F>::ld(P.in + 2 * off);
    if (P.err) {   // wire elements must be canonical (< r): the butterflies' add / sub assume it
      cc::sub_cc(x.v[0], F::mod()[0]);
#pragma unroll
      for (int i = 1; i < 8; i++) cc::subc_cc(x.v[i], F::mod()[i]);
      if (cc::subc(0u, 0u) == 0u) atomicExch(P.err, 1);   // no borrow: x >= r
    }
    if (P.tw_sel >= 0 && r != 0) {
      uint64_t tq = P.tw_sel == 0 ? q0 : (P.tw_sel == 1 ? q1 : q2);
      if (tq != 0) {
        if (P.tw_full) x = A::mul(x, nt<F>::ld(P.tw_full + 2 * off));
        else x = A::mul(x, nt<F>::tw(P.tab, (uint64_t)r * tq * P.tw_scale));
      }
    }
    nt<F>::sts(lo, hi, r * V + v, x);
  }
  if (P.tw_full_out) return;
  __syncthreads();
  // DIF rounds: radix 8 while possible, then 4 or 2
  int lblk = P.lr, rem = P.lr;
  while (rem >= 3) { dif_round<F, 3>(lo, hi, trl, trh, P.lr, P.lv, lblk, tid, nth); lblk -= 3; rem -= 3; __syncthreads(); }
  if (rem == 2) { dif_round<F, 2>(lo, hi, trl, trh, P.lr, P.lv, lblk, tid, nth); __syncthreads(); }
  if (rem == 1) { dif_round<F, 1>(lo, hi, trl, trh, P.lr, P.lv, lblk, tid, nth); __syncthreads(); }
  // store: position p holds X[k], k = digit reversal of p over the round radices
  const int n8 = P.lr / 3, last = P.lr % 3;
  for (int idx = tid; idx < R * V; idx += nth) {
    int v, k;
    if (P.store_k_fastest) { k = idx & (R - 1); v = idx >> P.lr; } else { v = idx & (V - 1); k = idx >> P.lv; }
    uint64_t q = qbase + v;
    if (q >= P.Q) continue;
    // p from k: k = m1 + 8 m2 + 64 m3 (+ ...), p = m1 R/8 + m2 R/64 + ...
    int p = 0, kk = k, sub = R;
    for (int t = 0; t < n8; t++) { sub >>= 3; p += (kk & 7) * sub; kk >>= 3; }
    if (last) { sub >>= last; p += (kk & ((1 << last) - 1)) * sub; }
    E x = nt<F>::lds(lo, hi, p * V + v);
    uint64_t q0 = q & (P.Q0 - 1), qr = q >> P.lq0;
    uint64_t q1 = qr & (P.Q1 - 1), q2 = qr >> P.lq1;
    if (P.otw_rsel >= 0) {
      uint64_t row = (P.otw_rsel == 0 ? q0 : (P.otw_rsel == 1 ? q1 : q2)) * P.otw_ra + (uint64_t)k * P.otw_rb;
      uint64_t col = P.otw_base + q0;
      uint64_t e = ((row * col) & (((uint64_t)1 << P.tab.log_root) - 1)) * P.otw_scale;
      if (e) x = A::mul(x, nt<F>::tw(P.tab, e));
    }
    if (P.scale_ninv) x = A::mul(x, nt<F>::ld(P.tab.ninv));
    uint64_t off = q0 * P.out_s0 + q1 * P.out_s1 + q2 * P.out_s2;
    uint4* dst;
    if (P.peer_rows) {   // power of two
      int h = k >> P.lpeer_rows;
      dst = P.peer_out[h] + 2 * (off + (uint64_t)(k & (P.peer_rows - 1)) * P.out_sr);
    } else {
      dst = P.out + 2 * (off + (uint64_t)k * P.out_sr);
    }
    nt<F>::st(dst, x);
  }
}

// ---------------------------------------------------------------------------------------------
template <class F>
static void gen_tables_t(NttTables& t, int log_root, int inverse, cudaStream_t st) {
  uint32_t nlo = 1u << (log_root < t.lo_bits ? log_root : t.lo_bits);
  uint32_t nhi = log_root > t.lo_bits ? 1u << (log_root - t.lo_bits) : 1u;
  uint32_t n = nlo + nhi + 1;
  k_ntt_tables<F><<<(n + 63) / 64, 64, 0, st>>>(const_cast<uint4*>(t.lo), const_cast<uint4*>(t.hi),
                                                 const_cast<uint4*>(t.ninv), log_root, t.lo_bits, inverse);
}

template <class F>
static cudaError_t launch_pass_t(const NttPassParams& Pin, cudaStream_t st) {
  NttPassParams P = Pin;
  int R = 1 << P.lr;
  int V = NTT_TILE / R < 4 ? 4 : NTT_TILE / R;
  P.lv = 0;
  while ((1 << P.lv) < V) P.lv++;
  size_t smem = ((size_t)2 * R * V + 2 * R) * sizeof(uint4);
  {   // per device and cheap: set on every launch (a process may drive several devices)
    cudaError_t e = cudaFuncSetAttribute(k_ntt_pass<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  uint64_t ctas = (P.Q + V - 1) / V;
  k_ntt_pass<F><<<(unsigned)ctas, NTT_THREADS, smem, st>>>(P);
  g_kernel_launches += 1;
  return cudaGetLastError();
}

void ntt_gen_tables(int field, NttTables& t, int log_root, int inverse, cudaStream_t st) {
  switch (field) {
    case 0: gen_tables_t<Fr377>(t, log_root, inverse, st); break;
    case 1: gen_tables_t<Fr254>(t, log_root, inverse, st); break;
    default: gen_tables_t<Fr381>(t, log_root, inverse, st); break;
  }
}
cudaError_t ntt_launch_pass(int field, const NttPassParams& P, cudaStream_t st) {
  switch (field) {
    case 0: return launch_pass_t<Fr377>(P, st);
    case 1: return launch_pass_t<Fr254>(P, st);
    default: return launch_pass_t<Fr381>(P, st);
  }
}
int ntt_two_adicity(int field) { return field == 0 ? Fr377::TWO_ADICITY : field == 1 ? Fr254::TWO_ADICITY : Fr381::TWO_ADICITY; }

}  // namespace bz
