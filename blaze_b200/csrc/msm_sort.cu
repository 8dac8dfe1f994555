// MSM front end, curve independent: scalar windowing (signed digits) and the two-level
// counting sort of (bucket, point-index) pairs, one sort per window.
//
// Black box being replaced: the FPGA MSM core's ingest + "bucket accumulation" scheduling
// (/root/reference/src/ingo_msm/msm_hw_code.rs:33-54 only exposes its phase counters).
//
// Data flow (M = number of (sub)scalars, W windows of c bits, nb = 2^(c-1)+1 buckets/window):
//   scalars (32 B or 4 B each, read ONCE, coalesced)
//     -> k_digits      dig[w][i]   = sign<<31 | |digit|                       (4 B x W x M)
//     -> k_hist1       hmat[w][tile][coarse] tile histograms of the LOW bucket bits (no atomics to HBM)
//     -> k_colscan1 / k_binscan1  exclusive prefix over tiles, then over coarse bins
//     -> k_scatter1    l1[w][pos]  = {bucket, sign|idx} grouped by coarse bin (8 B x W x M)
//     -> k_sort2       sorted[w*M + pos] = sign|idx grouped by bucket, and
//                      goff[w*nb + b] = start of bucket b of window w in `sorted`
// All of it is HBM-bound streaming / scatter; the histograms live in shared memory.
#include <cuda_runtime.h>

#include <cstdint>

#include "msm_internal.h"

namespace bz {

// ---------------------------------------------------------------------------------------------
// digits.  Signed-digit recoding without a carry chain: s' = s + K with
// K = sum_{w < W-1} 2^(c-1) * 2^(c w); digit_w = field_w(s') - 2^(c-1) for w < W-1 and the top
// window takes the remaining bits unsigned.  Host code picks W so that the top digit of the
// largest legal scalar is <= 2^(c-1) (msm_engine.cu: plan_windows).
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars, int words_per_scalar, uint64_t M,
                                                int W, int c, DigitConst dc, uint32_t* __restrict__ dig,
                                                int* __restrict__ err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  uint32_t s[10];
#pragma unroll
  for (int k = 0; k < 10; k++) s[k] = 0;
  if (words_per_scalar == 8) {
    const uint4* p = reinterpret_cast<const uint4*>(scalars) + 2 * i;
    uint4 a = __ldg(p), b = __ldg(p + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if (dc.check_mod) {   // scalar must be canonical (< r), like Fr::into_repr() output
      bool lt = false, decided = false;
#pragma unroll
      for (int k = 7; k >= 0; k--) {
        if (!decided && s[k] != dc.mod[k]) { lt = s[k] < dc.mod[k]; decided = true; }
      }
      if (!lt) atomicExch(err, BZ_ERR_SCALAR_RANGE);
    }
  } else {
    s[0] = __ldg(scalars + i);
  }
  // s' = s + K
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    uint64_t t = (uint64_t)s[k] + dc.K[k] + carry;
    s[k] = (uint32_t)t;
    carry = t >> 32;
  }
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (c >= 32) ? 0xffffffffu : ((1u << c) - 1);
  for (int w = 0; w < W; w++) {
    int o = w * c;
    int limb = o >> 5, sh = o & 31;
    uint64_t two = (uint64_t)s[limb] | ((uint64_t)s[limb + 1] << 32);
    uint32_t v = (uint32_t)(two >> sh);
    uint32_t out;
    if (w < W - 1) {
      v &= mask;
      int32_t d = (int32_t)v - (int32_t)half;
      uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      out = a | (d < 0 ? 0x80000000u : 0u);
    } else {
      // top window: everything that is left (host guarantees <= half for legal scalars)
      if (v > half) { atomicExch(err, BZ_ERR_SCALAR_RANGE); v = 0; }
      out = v;
    }
    dig[(uint64_t)w * M + i] = out;
  }
}

// ---------------------------------------------------------------------------------------------
// level 1: group by coarse bin = bucket & (ncoarse - 1).  Low bits, not high bits: a window whose digits
// only span a few bits (short top window, small scalars) would otherwise land in ONE coarse bin.
__global__ void __launch_bounds__(512) k_hist1(const uint32_t* __restrict__ dig, uint64_t M, uint32_t cmask, int ncoarse,
                                               uint32_t tile, uint32_t ntiles, uint32_t* __restrict__ hmat) {
  extern __shared__ uint32_t sh[];
  int w = blockIdx.y;
  uint32_t t = blockIdx.x;
  for (int k = threadIdx.x; k < ncoarse; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  uint64_t lo = (uint64_t)t * tile, hi = lo + tile;
  if (hi > M) hi = M;
  const uint32_t* d = dig + (uint64_t)w * M;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    uint32_t b = __ldg(d + i) & 0x7fffffffu;
    if (b) atomicAdd(&sh[b & cmask], 1u);   // zero digits contribute nothing: dropped here
  }
  __syncthreads();
  uint32_t* out = hmat + ((uint64_t)w * ntiles + t) * ncoarse;
  for (int k = threadIdx.x; k < ncoarse; k += blockDim.x) out[k] = sh[k];
}

// exclusive prefix down each column (over tiles); column totals to tot[w][bin]
__global__ void k_colscan1(uint32_t* __restrict__ hmat, uint32_t ntiles, int ncoarse, int W,
                           uint32_t* __restrict__ tot) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * ncoarse) return;
  int w = idx / ncoarse, bin = idx % ncoarse;
  uint32_t run = 0;
  uint32_t* col = hmat + (uint64_t)w * ntiles * ncoarse + bin;
  for (uint32_t t = 0; t < ntiles; t++) {
    uint32_t v = col[(uint64_t)t * ncoarse];
    col[(uint64_t)t * ncoarse] = run;
    run += v;
  }
  tot[idx] = run;
}

// per window: base1[w][bin] = exclusive prefix of tot[w][*]; base1[w][ncoarse] = M
__global__ void __launch_bounds__(1024) k_binscan1(const uint32_t* __restrict__ tot, int ncoarse,
                                                   uint32_t* __restrict__ base1) {
  __shared__ uint32_t part[1024];
  int w = blockIdx.x;
  const uint32_t* t = tot + (uint64_t)w * ncoarse;
  uint32_t* b = base1 + (uint64_t)w * (ncoarse + 1);
  int per = (ncoarse + blockDim.x - 1) / blockDim.x;
  int lo = threadIdx.x * per, hi = min(lo + per, ncoarse);
  uint32_t sum = 0;
  for (int k = lo; k < hi; k++) sum += t[k];
  part[threadIdx.x] = sum;
  __syncthreads();
  // inclusive Hillis-Steele over the 1024 partials
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (int k = lo; k < hi; k++) {
    b[k] = run;
    run += t[k];
  }
  if (threadIdx.x == blockDim.x - 1) b[ncoarse] = part[blockDim.x - 1];
}

// wbase[w] = number of non-zero digits in windows < w; wbase[W] = total length of `sorted`
__global__ void k_wbase(const uint32_t* __restrict__ base1, int ncoarse, int W, uint32_t* __restrict__ wbase,
                        uint32_t* __restrict__ goff_end) {
  if (blockIdx.x || threadIdx.x) return;
  uint32_t run = 0;
  for (int w = 0; w < W; w++) {
    wbase[w] = run;
    run += base1[(uint64_t)w * (ncoarse + 1) + ncoarse];
  }
  wbase[W] = run;
  *goff_end = run;
}

__global__ void __launch_bounds__(512) k_scatter1(const uint32_t* __restrict__ dig, uint64_t M, uint32_t cmask, int ncoarse,
                                                  uint32_t tile, uint32_t ntiles, const uint32_t* __restrict__ hmat,
                                                  const uint32_t* __restrict__ base1, uint2* __restrict__ l1) {
  extern __shared__ uint32_t sh[];
  int w = blockIdx.y;
  uint32_t t = blockIdx.x;
  const uint32_t* hm = hmat + ((uint64_t)w * ntiles + t) * ncoarse;
  const uint32_t* b1 = base1 + (uint64_t)w * (ncoarse + 1);
  for (int k = threadIdx.x; k < ncoarse; k += blockDim.x) sh[k] = b1[k] + hm[k];
  __syncthreads();
  uint64_t lo = (uint64_t)t * tile, hi = lo + tile;
  if (hi > M) hi = M;
  const uint32_t* d = dig + (uint64_t)w * M;
  uint2* out = l1 + (uint64_t)w * M;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    uint32_t e = __ldg(d + i);
    uint32_t b = e & 0x7fffffffu;
    if (!b) continue;
    uint32_t pos = atomicAdd(&sh[b & cmask], 1u);
    out[pos] = make_uint2(b, (e & 0x80000000u) | (uint32_t)i);
  }
}

// ---------------------------------------------------------------------------------------------
// level 2: one block per (coarse bin, window): counting sort by the high bits (bucket >> cbits)
__global__ void __launch_bounds__(256) k_sort2(const uint2* __restrict__ l1, uint64_t M, int cbits, int ncoarse,
                                               uint32_t nfine_, uint32_t nb, int W, const uint32_t* __restrict__ base1,
                                               const uint32_t* __restrict__ wbase_arr,
                                               uint32_t* __restrict__ sorted, uint32_t* __restrict__ goff) {
  extern __shared__ uint32_t sh[];   // [nfine] counters, then [256] scan scratch
  const int nfine = (int)nfine_;
  uint32_t* cnt = sh;
  uint32_t* scratch = sh + nfine;
  int w = blockIdx.y;
  int cb = blockIdx.x;
  const uint32_t* b1 = base1 + (uint64_t)w * (ncoarse + 1);
  uint32_t lo = b1[cb], hi = b1[cb + 1];
  const uint2* in = l1 + (uint64_t)w * M;
  for (int k = threadIdx.x; k < nfine; k += blockDim.x) cnt[k] = 0;
  __syncthreads();
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&cnt[in[i].x >> cbits], 1u);
  __syncthreads();
  // exclusive scan of cnt[0..nfine): each thread owns a contiguous run
  int per = (nfine + blockDim.x - 1) / blockDim.x;
  int klo = threadIdx.x * per, khi = min(klo + per, nfine);
  uint32_t sum = 0;
  for (int k = klo; k < khi; k++) sum += cnt[k];
  scratch[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    uint32_t v = threadIdx.x >= off ? scratch[threadIdx.x - off] : 0;
    __syncthreads();
    scratch[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = lo + (threadIdx.x ? scratch[threadIdx.x - 1] : 0);
  uint64_t wbase = wbase_arr[w];   // positions are global over the concatenated (compacted) windows
  for (int k = klo; k < khi; k++) {
    uint32_t v = cnt[k];
    cnt[k] = run;   // becomes the scatter cursor
    goff[(uint64_t)w * nb + (uint32_t)cb * nfine + k] = (uint32_t)(wbase + run);   // slot = coarse * nfine + fine
    run += v;
  }
  __syncthreads();
  uint32_t* out = sorted + wbase;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    uint2 e = in[i];
    uint32_t pos = atomicAdd(&cnt[e.x >> cbits], 1u);
    out[pos] = e.y;
  }
}

// ---------------------------------------------------------------------------------------------
void launch_msm_sort(const MsmPlan& p, const MsmWorkspace& ws, const uint32_t* scalars_dev, cudaStream_t st) {
  const uint64_t M = p.M;     // scalars
  const uint64_t Ms = p.Ms;   // entries per bucket window (merged: all Wd digit windows form ONE list, index = w*M + i)
  {
    unsigned blocks = (unsigned)((M + 255) / 256);
    k_digits<<<blocks, 256, 0, st>>>(scalars_dev, p.words_per_scalar, M, p.Wd, p.c, p.dc, ws.dig, ws.err);
  }
  dim3 g1(p.ntiles, p.W);
  size_t sh1 = (size_t)p.ncoarse * sizeof(uint32_t);
  k_hist1<<<g1, 512, sh1, st>>>(ws.dig, Ms, (uint32_t)p.ncoarse - 1, p.ncoarse, p.tile, p.ntiles, ws.hmat);
  {
    int n = p.W * p.ncoarse;
    k_colscan1<<<(n + 255) / 256, 256, 0, st>>>(ws.hmat, p.ntiles, p.ncoarse, p.W, ws.tot);
    k_binscan1<<<p.W, 1024, 0, st>>>(ws.tot, p.ncoarse, ws.base1);
    k_wbase<<<1, 32, 0, st>>>(ws.base1, p.ncoarse, p.W, ws.wbase, ws.goff + (size_t)p.W * p.nb);
  }
  k_scatter1<<<g1, 512, sh1, st>>>(ws.dig, Ms, (uint32_t)p.ncoarse - 1, p.ncoarse, p.tile, p.ntiles, ws.hmat, ws.base1, ws.l1);
  dim3 g2(p.ncoarse, p.W);
  size_t sh2 = ((size_t)p.nfine + 256) * sizeof(uint32_t);
  k_sort2<<<g2, 256, sh2, st>>>(ws.l1, Ms, p.cbits, p.ncoarse, p.nfine, p.nb, p.W, ws.base1, ws.wbase, ws.sorted, ws.goff);
  g_kernel_launches += 7;
}

}  // namespace bz
