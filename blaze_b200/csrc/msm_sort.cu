// MSM front end, curve independent: scalar windowing (signed digits) and the bucket sort of
// (bucket, point-reference) pairs.
//
// Black box being replaced: the FPGA MSM core's ingest + "bucket accumulation" scheduling
// (/root/reference/src/ingo_msm/msm_hw_code.rs:33-54 only exposes its phase counters).
//
// Data flow (M scalars, Wd digit windows of c bits, W bucket windows -- W = Wd, or 1 when the windows
// share one bucket set (MsmPlan::merged) -- Ms = entries per bucket window, kb = c-1 key bits):
//   scalars (32 B or 4 B each, read ONCE, coalesced)
//     -> k_digits        dig[w][i] = sign<<31 | |digit|                          (4 B per entry)
//     -> partition level 1..nlev (most significant SORT-KEY bits first, <= 8 bits per level):
//          k_part_hist     per-parent child histogram (shared-memory atomics, one global add per child and tile)
//          k_part_scan     exclusive prefix -> child offsets (= next level's parent ranges), scatter cursors, tile map
//          k_part_scatter  a tile of 8192 entries of ONE parent is ranked in shared memory, every child run of the
//                          tile reserves its output range with one atomicAdd and is written at once (runs of
//                          256-512 B issued together: full-sector DRAM writes)         (8 B per entry and level)
//     -> k_final         one CTA per final parent (~8-16 K entries): counting sort by the last fb bits in shared
//                        memory, payloads staged in shared memory and written coalesced:
//                        sorted[pos] = sign|ref grouped by bucket, goff[slot] = start of the bucket
// Sort key of bucket value b (1 .. 2^kb; zero digits are dropped): b' = b - 1,
//     K = (b' mod 2^rest) << fb  |  b' >> rest
// i.e. the partition levels consume the LOW bits of b' first: windows whose digits span only a few bits
// (short top window, small scalars) still spread over all parents.  Bucket slot = K, so the reduction reads
// value i at slot (i mod 2^rest) * 2^fb + (i >> rest).  The order of the entries inside a bucket is
// irrelevant (the group sum is unique), so ranks come from atomics.
#include <cuda_runtime.h>

#include <cstdint>

#include "msm_internal.h"

namespace bz {

// ---------------------------------------------------------------------------------------------
// digits.  Signed-digit recoding without a carry chain: s' = s + K with
// K = sum_{w < W-1} 2^(c-1) * 2^(c w); digit_w = field_w(s') - 2^(c-1) for w < W-1 and the top
// window takes the remaining bits unsigned.  Host code picks W so that the top digit of the
// largest legal scalar is <= 2^(c-1) (msm_client.cu: plan_windows).
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
// partition levels
struct PartArgs {
  const uint32_t* dig;         // FIRST level input: [W][Ms] digit entries (payload index = position in the window)
  const uint2* in;             // later levels: {K, sign|ref} pairs grouped by parent
  uint64_t Ms;                 // FIRST: entries per window
  uint32_t tiles_per_win;      // FIRST: ceil(Ms / PART_TILE)
  const uint32_t* parent_off;  // later levels: [nparents + 1] ranges in `in` (previous level's child offsets)
  const uint32_t* tpref;       // later levels: [nparents + 1] tile-count prefix
  uint32_t nparents;
  int kshift;                  // child = (K >> kshift) & (2^bits - 1)
  int bits;
  int rest, fb;                // sort-key construction (FIRST only)
  int gs;                      // log2 of the cursor groups per child: tile t uses group t mod 2^gs, so the tiles of a
                               // level with few children do not all hammer the same few atomic counters
  uint32_t* hist;              // [(nparents << bits) << gs]  child-major, group-minor
  uint32_t* cursor;            // [(nparents << bits) << gs] scatter cursors (absolute output positions)
  uint2* out;
};

#ifndef PART_TILE
#define PART_TILE 8192
#endif
#ifndef PART_THREADS
#define PART_THREADS 512
#endif
#define PART_EPT (PART_TILE / PART_THREADS)

__device__ __forceinline__ uint32_t sort_key(uint32_t b, int rest, int fb) {   // b >= 1
  const uint32_t v = b - 1;
  return ((v & ((1u << rest) - 1)) << fb) | (v >> rest);
}

// tile -> (parent, [lo, hi)) ; returns false when the tile is past the end
template <bool FIRST>
__device__ __forceinline__ bool part_tile(const PartArgs& a, uint32_t& parent, uint64_t& lo, uint64_t& hi) {
  const uint32_t t = blockIdx.x;
  if (FIRST) {
    parent = t / a.tiles_per_win;
    if (parent >= a.nparents) return false;
    const uint64_t l = (uint64_t)(t - parent * a.tiles_per_win) * PART_TILE;
    lo = (uint64_t)parent * a.Ms + l;
    hi = (uint64_t)parent * a.Ms + (l + PART_TILE < a.Ms ? l + PART_TILE : a.Ms);
    return true;
  } else {
    if (t >= __ldg(a.tpref + a.nparents)) return false;
    uint32_t l = 0, h = a.nparents;   // largest p with tpref[p] <= t  (parents without tiles share a value: take the last)
    while (h - l > 1) {
      uint32_t mid = l + ((h - l) >> 1);
      if (__ldg(a.tpref + mid) <= t) l = mid; else h = mid;
    }
    parent = l;
    const uint64_t p0 = __ldg(a.parent_off + parent), p1 = __ldg(a.parent_off + parent + 1);
    lo = p0 + (uint64_t)(t - __ldg(a.tpref + parent)) * PART_TILE;
    hi = lo + PART_TILE < p1 ? lo + PART_TILE : p1;
    return lo < hi;
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(PART_THREADS) k_part_hist(PartArgs a) {
  __shared__ uint32_t cnt[256];
  uint32_t parent;
  uint64_t lo, hi;
  if (!part_tile<FIRST>(a, parent, lo, hi)) return;
  const int nchild = 1 << a.bits;
  for (int k = threadIdx.x; k < nchild; k += blockDim.x) cnt[k] = 0;
  __syncthreads();
  const uint32_t mask = (uint32_t)nchild - 1;
#pragma unroll 4
  for (uint64_t i = lo + threadIdx.x; i < hi; i += PART_THREADS) {
    uint32_t K;
    if (FIRST) {
      uint32_t b = __ldg(a.dig + i) & 0x7fffffffu;
      if (!b) continue;   // zero digits contribute nothing: dropped here
      K = sort_key(b, a.rest, a.fb);
    } else {
      K = a.in[i].x;
    }
    atomicAdd(&cnt[(K >> a.kshift) & mask], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nchild; k += blockDim.x) {
    uint32_t v = cnt[k];
    if (v) atomicAdd(a.hist + (((((uint64_t)parent << a.bits) + k) << a.gs) | (blockIdx.x & ((1u << a.gs) - 1))), v);
  }
}

// single CTA.  hist holds n children x 2^gs group counters (child-major).  cursor = exclusive prefix over all
// counters (absolute output positions of every (child, group) run), off[child] = start of the child
// (off[n] = total), tpref_next = exclusive prefix of ceil(child total / PART_TILE): the children are the next
// level's parents.
__global__ void __launch_bounds__(1024) k_part_scan(const uint32_t* __restrict__ hist, uint32_t n, int gs,
                                                    uint32_t* __restrict__ off, uint32_t* __restrict__ cursor,
                                                    uint32_t* __restrict__ tpref_next, uint32_t* __restrict__ total_out) {
  __shared__ uint32_t pa[1024], pb[1024];
  const uint32_t G = 1u << gs;
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per < n ? threadIdx.x * per : n, hi = lo + per < n ? lo + per : n;
  uint32_t sa = 0, sb = 0;
  for (uint32_t k = lo; k < hi; k++) {
    uint32_t v = 0;
    for (uint32_t g = 0; g < G; g++) v += hist[((uint64_t)k << gs) + g];
    sa += v;
    sb += (v + PART_TILE - 1) / PART_TILE;
  }
  pa[threadIdx.x] = sa;
  pb[threadIdx.x] = sb;
  __syncthreads();
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {
    uint32_t va = (int)threadIdx.x >= o ? pa[threadIdx.x - o] : 0, vb = (int)threadIdx.x >= o ? pb[threadIdx.x - o] : 0;
    __syncthreads();
    pa[threadIdx.x] += va;
    pb[threadIdx.x] += vb;
    __syncthreads();
  }
  uint32_t ra = threadIdx.x ? pa[threadIdx.x - 1] : 0, rb = threadIdx.x ? pb[threadIdx.x - 1] : 0;
  for (uint32_t k = lo; k < hi; k++) {
    off[k] = ra;
    tpref_next[k] = rb;
    uint32_t v = 0;
    for (uint32_t g = 0; g < G; g++) {
      uint32_t h = hist[((uint64_t)k << gs) + g];
      cursor[((uint64_t)k << gs) + g] = ra + v;
      v += h;
    }
    ra += v;
    rb += (v + PART_TILE - 1) / PART_TILE;
  }
  if (threadIdx.x == blockDim.x - 1) {
    off[n] = pa[blockDim.x - 1];
    tpref_next[n] = pb[blockDim.x - 1];
    if (total_out) *total_out = pa[blockDim.x - 1];
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(PART_THREADS) k_part_scatter(PartArgs a) {
  // the tile is regrouped by child in shared memory, then copied out linearly: consecutive lanes write
  // consecutive 8-byte pairs of one child run (full 32-byte sectors, 4x fewer L2 write requests than scattering
  // from registers)
  extern __shared__ uint2 stage[];   // [PART_TILE]
  __shared__ uint32_t cnt[256], lbase[256], gdelta[256], wsum[8];
  uint32_t parent;
  uint64_t lo, hi;
  if (!part_tile<FIRST>(a, parent, lo, hi)) return;
  const int nchild = 1 << a.bits;
  for (int k = threadIdx.x; k < 256; k += blockDim.x) cnt[k] = 0;
  __syncthreads();
  const uint32_t mask = (uint32_t)nchild - 1;
  uint32_t K[PART_EPT], pay[PART_EPT], rk[PART_EPT];   // rk = child << 16 | rank (rank < 8192), 0xffffffff = no entry
  const uint64_t win_base = FIRST ? (uint64_t)parent * a.Ms : 0;
  // all loads of the tile first (independent: 16 in flight per thread), then the ranking
#pragma unroll
  for (int j = 0; j < PART_EPT; j++) {
    const uint64_t i = lo + (uint64_t)j * PART_THREADS + threadIdx.x;
    K[j] = 0;
    pay[j] = 0;
    if (i < hi) {
      if (FIRST) {
        pay[j] = __ldg(a.dig + i);   // raw digit entry for now
      } else {
        uint2 e = a.in[i];
        K[j] = e.x;
        pay[j] = e.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < PART_EPT; j++) {
    const uint64_t i = lo + (uint64_t)j * PART_THREADS + threadIdx.x;
    rk[j] = 0xffffffffu;
    bool have = i < hi;
    if (FIRST && have) {
      const uint32_t e = pay[j];
      const uint32_t b = e & 0x7fffffffu;
      have = b != 0;   // zero digits are dropped
      if (have) {
        K[j] = sort_key(b, a.rest, a.fb);
        pay[j] = (e & 0x80000000u) | (uint32_t)(i - win_base);
      }
    }
    if (have) {
      uint32_t ch = (K[j] >> a.kshift) & mask;
      rk[j] = (ch << 16) | atomicAdd(&cnt[ch], 1u);
    }
  }
  __syncthreads();
  // exclusive prefix of the 256 child counts (tile-local positions) + one global reservation per child run
  uint32_t v = 0, inc = 0;
  if (threadIdx.x < 256) {
    v = cnt[threadIdx.x];
    inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)(threadIdx.x & 31) >= o) inc += u;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    uint32_t pre = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) pre += wsum[w];
    const uint32_t lb = pre + inc - v;
    lbase[threadIdx.x] = lb;
    uint32_t g = 0;
    if (v && (int)threadIdx.x < nchild)
      g = atomicAdd(a.cursor + (((((uint64_t)parent << a.bits) + threadIdx.x) << a.gs) | (blockIdx.x & ((1u << a.gs) - 1))), v);
    gdelta[threadIdx.x] = g - lb;   // output position = gdelta[child] + tile-local position
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PART_EPT; j++) {
    if (rk[j] != 0xffffffffu) stage[lbase[rk[j] >> 16] + (rk[j] & 0xffffu)] = make_uint2(K[j], pay[j]);
  }
  __syncthreads();
  uint32_t n = 0;
  for (int w = 0; w < 8; w++) n += wsum[w];
  for (uint32_t idx = threadIdx.x; idx < n; idx += PART_THREADS) {
    uint2 e = stage[idx];
    a.out[gdelta[(e.x >> a.kshift) & mask] + idx] = e;
  }
}

// ---------------------------------------------------------------------------------------------
// final level: one CTA per parent P (all entries share the top `rest` sort-key bits); counting sort by the
// low fb bits.  Bucket slots of P are [P << fb, (P+1) << fb) in goff (windows included: slot = w * 2^kb + K).
#define FINAL_THREADS 256
#define FINAL_STAGE 10240   // payloads staged in shared memory (40 KB) when the parent is not larger
__global__ void __launch_bounds__(FINAL_THREADS) k_final(const uint2* __restrict__ in, const uint32_t* __restrict__ parent_off,
                                                         uint32_t nparents, int fb, uint32_t* __restrict__ sorted,
                                                         uint32_t* __restrict__ goff) {
  __shared__ uint32_t cnt[256];
  __shared__ uint32_t stage[FINAL_STAGE];
  const uint32_t P = blockIdx.x;
  if (P >= nparents) return;
  const uint32_t lo = __ldg(parent_off + P), hi = __ldg(parent_off + P + 1);
  const int nf = 1 << fb;
  const uint32_t fmask = (uint32_t)nf - 1;
  for (int k = threadIdx.x; k < nf; k += blockDim.x) cnt[k] = 0;
  __syncthreads();
  // four independent loads in flight per thread, then the shared-memory atomics
  for (uint32_t i0 = lo + threadIdx.x; i0 < hi; i0 += 4 * FINAL_THREADS) {
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t i = i0 + u * FINAL_THREADS;
      k[u] = i < hi ? in[i].x : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * FINAL_THREADS < hi) atomicAdd(&cnt[k[u] & fmask], 1u);
  }
  __syncthreads();
  // exclusive scan of <= 256 counters by warp 0 (<= 8 per lane)
  if (threadIdx.x < 32) {
    const int per = (nf + 31) / 32;
    const int klo = (int)threadIdx.x * per < nf ? (int)threadIdx.x * per : nf, khi = klo + per < nf ? klo + per : nf;
    uint32_t sum = 0;
    for (int k = klo; k < khi; k++) sum += cnt[k];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)threadIdx.x >= o) inc += v;
    }
    uint32_t run = inc - sum;
    for (int k = klo; k < khi; k++) {
      uint32_t v = cnt[k];
      cnt[k] = run;   // becomes the (parent-relative) scatter cursor
      goff[((uint64_t)P << fb) + k] = lo + run;
      run += v;
    }
  }
  __syncthreads();
  const uint32_t n = hi - lo;
  if (n <= FINAL_STAGE) {
    for (uint32_t i0 = lo + threadIdx.x; i0 < hi; i0 += 4 * FINAL_THREADS) {
      uint2 e[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t i = i0 + u * FINAL_THREADS;
        e[u] = i < hi ? in[i] : make_uint2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (i0 + u * FINAL_THREADS < hi) stage[atomicAdd(&cnt[e[u].x & fmask], 1u)] = e[u].y;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += FINAL_THREADS) sorted[lo + i] = stage[i];
  } else {
    for (uint32_t i = lo + threadIdx.x; i < hi; i += FINAL_THREADS) {
      uint2 e = in[i];
      sorted[lo + atomicAdd(&cnt[e.x & fmask], 1u)] = e.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
void launch_msm_sort(const MsmPlan& p, const MsmWorkspace& ws, const uint32_t* scalars_dev, cudaStream_t st) {
  const uint64_t M = p.M;     // scalars
  const uint64_t Ms = p.Ms;   // entries per bucket window (merged: all Wd digit windows form ONE list, index = w*M + i)
  const uint64_t total = (uint64_t)p.W * Ms;
  {
    unsigned blocks = (unsigned)((M + 255) / 256);
    k_digits<<<blocks, 256, 0, st>>>(scalars_dev, p.words_per_scalar, M, p.Wd, p.c, p.dc, ws.dig, ws.err);
  }
  cudaMemsetAsync(ws.lvl_hist[0], 0, p.lvl_hist_words * sizeof(uint32_t), st);   // all levels are one allocation
  int consumed = 0;
  const uint2* in = nullptr;
  uint64_t launches = 2;
  const size_t stage_bytes = (size_t)PART_TILE * sizeof(uint2);   // 64 KB: above the static limit
  // per device and cheap: set on every call (a process may drive several devices)
  cudaFuncSetAttribute(k_part_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
  cudaFuncSetAttribute(k_part_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
  for (int l = 0; l < p.nlev; l++) {
    PartArgs a{};
    a.dig = ws.dig;
    a.in = in;
    a.Ms = Ms;
    a.tiles_per_win = (uint32_t)((Ms + PART_TILE - 1) / PART_TILE);
    a.parent_off = l ? ws.lvl_off[l - 1] : nullptr;
    a.tpref = l ? ws.lvl_tpref[l - 1] : nullptr;
    a.nparents = (uint32_t)p.W << consumed;
    a.bits = p.lbits[l];
    a.kshift = p.kb - consumed - p.lbits[l];
    a.rest = p.rest;
    a.fb = p.fb;
    a.gs = p.lgs[l];
    a.hist = ws.lvl_hist[l];
    a.cursor = ws.lvl_cursor[l];
    a.out = (l & 1) ? ws.pairB : ws.pairA;
    // tiles: level 0 is arithmetic; later levels waste less than one tile per parent
    const uint64_t ntiles = l == 0 ? (uint64_t)a.tiles_per_win * p.W : total / PART_TILE + a.nparents;
    const uint32_t nchildren = a.nparents << a.bits;
    if (l == 0) k_part_hist<true><<<(unsigned)ntiles, PART_THREADS, 0, st>>>(a);
    else        k_part_hist<false><<<(unsigned)ntiles, PART_THREADS, 0, st>>>(a);
    k_part_scan<<<1, 1024, 0, st>>>(ws.lvl_hist[l], nchildren, a.gs, ws.lvl_off[l], ws.lvl_cursor[l], ws.lvl_tpref[l],
                                    l + 1 == p.nlev ? ws.goff + (size_t)p.W * p.nb : nullptr);
    if (l == 0) k_part_scatter<true><<<(unsigned)ntiles, PART_THREADS, stage_bytes, st>>>(a);
    else        k_part_scatter<false><<<(unsigned)ntiles, PART_THREADS, stage_bytes, st>>>(a);
    launches += 3;
    in = a.out;
    consumed += p.lbits[l];
  }
  const uint32_t nparents = (uint32_t)p.W << p.rest;
  k_final<<<nparents, FINAL_THREADS, 0, st>>>(in, ws.lvl_off[p.nlev - 1], nparents, p.fb, ws.sorted, ws.goff);
  g_kernel_launches += launches;
}

}  // namespace bz
