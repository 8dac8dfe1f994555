// MSMClient, one device ("leaf"): the task state machine that drives the CUDA pipeline of msm_sort.cu / msm_curve.cuh.
// The C ABI names of include/blaze_b200.h are defined in msm_api.cu and land here (directly, or once per member device
// for a multi-device client).
//
// Mirrors (behaviour, not code) /root/reference/src/ingo_msm/msm_api.rs: same call order tolerance
// (initialize -> start_process -> set_data -> wait_result -> result), same mode selection
// ((mem_type, hbm_point_addr) -> DMA / HBM, msm_api.rs:75-95,163-216), same wire sizes
// (msm_cfg.rs:44-92).  Register polling becomes CUDA events.  No CPU fallback.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/blaze_b200.h"
#include "api_common.h"
#include "client_internal.h"
#include "msm_client.h"
#include "msm_internal.h"

using namespace bz;

#define fail bz_fail

#ifndef BZ_MSM_BA2_DEFAULT
#define BZ_MSM_BA2_DEFAULT 0   // measured slower than the XYZZ sweep on B200 (profiles/r2_ncu_k_accumulate_ba_2p22.txt): opt-in
#endif
#ifndef BZ_MSM_SEG_MAX
#define BZ_MSM_SEG_MAX 256
#endif
#ifndef BZ_BA2_CTAS_PER_SM
#define BZ_BA2_CTAS_PER_SM 3     // = BZ_BA2_MINBLOCKS of msm_ba2.cuh
#endif

// ------------------------------------------------------------------------------------ MSM planning
static const CurveOps* ops_for(int curve) {
  switch (curve) {
    case BZ_CURVE_BLS377: return curve_ops_bls12_377();
    case BZ_CURVE_BN254: return curve_ops_bn254();
    case BZ_CURVE_BLS381: return curve_ops_bls12_381();
  }
  return nullptr;
}

// 288-bit little-endian helper for window planning
struct Big9 {
  uint32_t v[9];
};
static void big_add_pow2(Big9& a, int bit) {   // a += 2^bit
  uint64_t carry = 1ull << (bit & 31);
  for (int k = bit >> 5; k < 9 && carry; k++) {
    uint64_t t = (uint64_t)a.v[k] + carry;
    a.v[k] = (uint32_t)t;
    carry = t >> 32;
  }
}
static uint64_t big_shr(const Big9& a, int bit) {   // (a >> bit) truncated to 64 bits
  uint64_t r = 0;
  for (int i = 0; i < 64; i++) {
    int b = bit + i;
    if (b >= 288) break;
    if ((a.v[b >> 5] >> (b & 31)) & 1) r |= 1ull << i;
  }
  return r;
}

// Number of windows for c-bit signed digits so that the TOP (unsigned) digit of the largest legal
// scalar `smax` stays <= 2^(c-1); fills K = sum_{w<W-1} 2^(c-1+cw).
static int plan_windows(const uint32_t smax[8], int sbits, int c, DigitConst& dc) {
  for (int W = std::max(1, (sbits + c - 1) / c);; W++) {
    Big9 k;
    memset(&k, 0, sizeof(k));
    for (int w = 0; w < W - 1; w++) big_add_pow2(k, c - 1 + c * w);
    Big9 s = k;   // s = smax + K
    uint64_t carry = 0;
    for (int i = 0; i < 9; i++) {
      uint64_t t = (uint64_t)s.v[i] + (i < 8 ? smax[i] : 0) + carry;
      s.v[i] = (uint32_t)t;
      carry = t >> 32;
    }
    // with W >= ceil(sbits/c) the shifted value has at most c+2 bits, so 64 bits are enough
    if (big_shr(s, c * (W - 1)) <= (1ull << (c - 1))) {
      memcpy(dc.K, k.v, sizeof(dc.K));
      return W;
    }
  }
}

static void wtable_free(bz_msm* m) {
  if (m->wtable) cudaFree(m->wtable);
  m->wtable = nullptr;
  m->wtable_bytes = 0;
  m->wtable_n = 0;
  m->wtable_c = m->wtable_levels = 0;
}

static void ws_free(bz_msm* m) {
  for (void* p : m->ws_allocs) cudaFree(p);
  m->ws_allocs.clear();
  memset(&m->ws, 0, sizeof(m->ws));
  m->have_plan = false;
}

template <class T>
static cudaError_t ws_alloc(bz_msm* m, T** p, size_t bytes) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
  if (e == cudaSuccess) { m->ws_allocs.push_back(q); *p = (T*)q; }
  return e;
}

static int ilog2_floor(uint64_t v) { int l = 0; while (v >>= 1) l++; return l; }

static size_t ws_bytes_estimate(const bz_msm* m, uint64_t total, int c, int W) {
  const size_t xb = m->ops->xyzz_bytes;
  const uint64_t nseg = total / 256 + 1;
  return (size_t)(total * 24 + (uint64_t)W * ((1ull << (c - 1)) + (1ull << 13)) * xb * 9 / 8 + nseg * 2 * (xb + 4) * 33 / 32 + (64ull << 20));
}

static int32_t make_plan(bz_msm* m, uint64_t M, int words_per_scalar, bool merged) {
  if (m->have_plan && m->plan.M == M && m->plan.words_per_scalar == words_per_scalar && (m->plan.merged != 0) == merged &&
      (m->forced_c == 0 || m->forced_c == m->plan.c))
    return BZ_OK;
  ws_free(m);
  size_t mem_free = 0, mem_total = 0;
  if (merged) {
    cudaMemGetInfo(&mem_free, &mem_total);
    mem_free += m->wtable_bytes;   // an existing merged table is reused or replaced
    double reserve_gb = 20.0;      // left for the caller's other clients (e.g. the 16 GiB of NTT slots at 2^27)
    if (const char* e = getenv("BZ_MSM_PRECOMP_RESERVE_GB")) reserve_gb = atof(e);
    size_t reserve = (size_t)(std::min(reserve_gb * 1073741824.0, (double)mem_total * 0.5));
    mem_free = mem_free > reserve ? mem_free - reserve : 0;
  }
  MsmPlan p{};
  p.M = M;
  p.words_per_scalar = words_per_scalar;
  uint32_t smax[8];
  int sbits;
  if (words_per_scalar == 8) {
    // largest canonical scalar is r - 1
    memcpy(smax, m->ops->fr_mod, 32);
    smax[0] -= 1;   // r is odd
    sbits = m->ops->scalar_bits;
    memcpy(p.dc.mod, m->ops->fr_mod, 32);
    p.dc.check_mod = 1;
  } else {
    memset(smax, 0, sizeof(smax));
    smax[0] = 0xffffffffu;
    sbits = 32;
    p.dc.check_mod = 0;
  }
  // window size: minimise (#mixed adds in accumulation) + (cost of the running-sum reduction)
  int best_c = 0;
  double best = 1e300;
  int env_c = 0;
  if (const char* e = getenv("BZ_MSM_C")) env_c = atoi(e);
  int force = m->forced_c ? m->forced_c : env_c;
  for (int c = 4; c <= (merged ? 26 : 23); c++) {
    if (force && c != force) continue;
    DigitConst dcx{};
    int W = plan_windows(smax, sbits, c, dcx);
    if ((uint64_t)W * M >= (1ull << 32)) continue;
    if (merged) {
      // one bucket set: the reduction is paid once, the table costs W * M entries of HBM
      if ((uint64_t)W * M >= (1ull << 31)) continue;   // entry index = w*M + i must leave bit 31 for the sign
      if ((size_t)W * M * m->ops->affine_bytes + ws_bytes_estimate(m, (uint64_t)W * M, c, 1) > mem_free) continue;
      double cost = (double)W * (double)M * 1.06 + 4.0 * (double)(1ull << (c - 1));
      if (cost < best) { best = cost; best_c = c; }
      continue;
    }
    // measured on B200 (perf_probe, 2^24..2^26): per (scalar, window) the sort costs about 0.06 of a mixed add;
    // the running-sum reduction costs about 4 mixed-add equivalents per bucket (11 ms for 2^23 buckets)
    double sort_w = 0.06;
    double cost = (double)W * ((double)M * (1.0 + sort_w) + 4.0 * (double)(1ull << (c - 1)));
    // a top window with only a few bits funnels all M entries into a handful of buckets of one
    // coarse bin (one CTA sorts them, long merge chains): avoid such c unless the problem is tiny
    int top_bits = sbits - c * (W - 1);
    if (top_bits < 6 && M > (1u << 16)) cost *= 1.5;
    if (cost < best) { best = cost; best_c = c; }
  }
  if (!best_c) {
    if (merged) return BZ_ERR_NO_RESULT;   // does not fit in HBM (caller falls back to the plain table); not an error
    return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "no feasible window size for %llu elements", (unsigned long long)M);
  }
  p.batch_affine = 0;
  if (const char* e = getenv("BZ_MSM_BA")) p.batch_affine = (atoi(e) && !merged) ? 1 : 0;

  p.tma_stage = 0;
  if (const char* e = getenv("BZ_MSM_TMA")) p.tma_stage = atoi(e) ? 1 : 0;
  p.c = best_c;
  p.Wd = plan_windows(smax, sbits, p.c, p.dc);
  p.merged = merged ? 1 : 0;
  p.W = merged ? 1 : p.Wd;
  p.Ms = merged ? (uint64_t)p.Wd * M : M;
  const uint64_t Ms = p.Ms;
  // sort levels (msm_sort.cu): final in-CTA level of fb <= 8 bits over parents of ~8K entries; the other
  // `rest` key bits go to partition levels of <= 8 bits each
  {
    p.kb = p.c - 1;
    int rest_t = 0;
    while (rest_t < 30 && ((double)Ms / (double)(1ull << rest_t)) > 8192.0) rest_t++;
    int lo = std::max(p.kb - 8, 0), hi = p.kb;
    if (lo <= 16) hi = std::min(hi, 16);   // two partition levels whenever they suffice
    p.rest = std::max(lo, std::min(rest_t, hi));
    p.fb = p.kb - p.rest;
    p.nlev = std::max(1, (p.rest + 7) / 8);
    int left = p.rest;
    for (int l = 0; l < p.nlev; l++) {
      p.lbits[l] = (left + (p.nlev - l) - 1) / (p.nlev - l);
      left -= p.lbits[l];
    }
    p.nb = 1u << p.kb;
    p.nvalues = p.nb;
  }
  uint64_t total = (uint64_t)p.W * Ms;
  // segment length: enough threads to fill the machine several times over, at most BZ_MSM_SEG_MAX entries each (measured
  // at 2^26: 256 -> 248.6 ms accumulate + 14.3 ms tail, 1024 -> 255.5 + 13.2)
  uint32_t L = BZ_MSM_SEG_MAX;
  while (L > 16 && total / L < 148ull * 256 * 8) L >>= 1;
  if (const char* e = getenv("BZ_MSM_SEG")) L = (uint32_t)std::max(1, atoi(e));
  p.seg_len = L;
  p.nseg = (total + L - 1) / L;
  // reduction chunk (k_reduce_level): 16 buckets per thread, 8 when that leaves the machine underfilled
  p.chunk = ((uint64_t)p.W * p.nvalues / 16 < 148ull * 256) ? 8 : 16;
  if (const char* e = getenv("BZ_MSM_CHUNK")) {
    uint32_t ch = (uint32_t)atoi(e);
    if (ch >= 2 && (ch & (ch - 1)) == 0 && ch <= 1024) p.chunk = ch;
  }
  p.nchunks = (p.nvalues + p.chunk - 1) / p.chunk;
  uint32_t nch1 = (p.nchunks + 3) / 4;   // upper reduction levels use chunks of 4 (msm_curve.cuh)

  size_t xb = m->ops->xyzz_bytes;
  cudaError_t e = cudaSuccess;
  auto A = [&](auto** ptr, size_t bytes) { if (e == cudaSuccess) e = ws_alloc(m, ptr, bytes); };
  A(&m->ws.dig, total * 4);
  {
    size_t words = 0;
    int consumed = 0;
    for (int l = 0; l < p.nlev; l++) {
      consumed += p.lbits[l];
      // cursor groups: aim at >= 2^16 counters per level so that ~10^5 tiles do not serialise on a few addresses
      int lg = 0;
      while (lg < 8 && (((uint64_t)p.W << consumed) << lg) < 65536 && (total >> 13) > (((uint64_t)p.W << consumed) << lg)) lg++;
      p.lgs[l] = lg;
      words += (((size_t)p.W << consumed) << lg) + 64;
    }
    p.lvl_hist_words = words;
    uint32_t *h = nullptr, *o = nullptr, *cu = nullptr, *tp = nullptr;
    A(&h, words * 4);
    A(&o, words * 4);
    A(&cu, words * 4);
    A(&tp, words * 4);
    size_t at = 0;
    consumed = 0;
    for (int l = 0; l < p.nlev && e == cudaSuccess; l++) {
      consumed += p.lbits[l];
      m->ws.lvl_hist[l] = h + at; m->ws.lvl_off[l] = o + at; m->ws.lvl_cursor[l] = cu + at; m->ws.lvl_tpref[l] = tp + at;
      at += (((size_t)p.W << consumed) << p.lgs[l]) + 64;
    }
  }
  A(&m->ws.pairA, total * 8);
  if (p.nlev > 1) A(&m->ws.pairB, total * 8);
  A(&m->ws.sorted, total * 4);
  A(&m->ws.goff, ((size_t)p.W * p.nb + 1) * 4);
  A((uint8_t**)&m->ws.buckets, (size_t)p.W * p.nb * xb);
  A(&m->ws.part_id, (size_t)p.nseg * 2 * 4);
  A((uint8_t**)&m->ws.part_pt, (size_t)p.nseg * 2 * xb);
  {
    uint64_t n = p.nseg, tot = 0;
    for (int lvl = 0;; lvl++) { const uint64_t g = merge_group(lvl, p.nseg); n = (n + g - 1) / g; tot += n; if (n == 1) break; }
    A(&m->ws.part2_id, (size_t)tot * 2 * 4);
    A((uint8_t**)&m->ws.part2_pt, (size_t)tot * 2 * xb);
  }
  if (p.batch_affine) {
    const size_t fb = (size_t)m->ops->fq_bytes, ab = m->ops->affine_list_bytes;
    const uint64_t ng = (uint64_t)p.W * p.nb;
    const uint64_t S1 = (total >> 1) + ng, S2 = (total >> 2) + ng;
    const uint64_t n0 = ((S1 + 511) / 512) * 32 + 256, n1 = n0 / 256 + 2, n2 = n1 / 256 + 2, n3 = n2 / 256 + 2;
    A(&m->ws.ba_scalars, 16);
    A(&m->ws.ba_pref, S1 * fb);
    A(&m->ws.ba_tot, n0 * fb);
    A(&m->ws.ba_itot, n0 * fb);
    A(&m->ws.ba_lvl[0], n1 * fb);
    A(&m->ws.ba_lvlp[0], n1 * fb);
    A(&m->ws.ba_lvl[1], n2 * fb);
    A(&m->ws.ba_lvlp[1], n2 * fb);
    A(&m->ws.ba_lvl[2], n3 * fb);
    A((uint8_t**)&m->ws.ba_buf1, S1 * ab);
    A((uint8_t**)&m->ws.ba_buf0, S2 * ab);
  }
  A((uint8_t**)&m->ws.red_a, (size_t)2 * p.W * p.nchunks * xb);        // S and V of the even levels
  A((uint8_t**)&m->ws.red_b, (size_t)2 * p.W * (nch1 + 1) * xb);       // ... of the odd levels
  A(&m->ws.err, 16);
  A(&m->ws.result, 256);
  if (e != cudaSuccess) {
    ws_free(m);
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    return fail(BZ_ERR_WRITE, "workspace allocation failed for M=%llu c=%d merged=%d: %s (%zu MiB of %zu MiB free)", (unsigned long long)M, p.c,
                (int)merged, cudaGetErrorString(e), f >> 20, t >> 20);
  }
  (void)ilog2_floor;
  m->plan = p;
  m->have_plan = true;
  return BZ_OK;
}

// ------------------------------------------------------------------------------------ MSM client (leaf)
int32_t bz::leaf_new(bz_dclient* dc, int32_t curve, int32_t mem_type, int32_t is_precompute, bz_msm** out) {
  *out = nullptr;
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  const CurveOps* ops = ops_for(curve);
  if (!ops) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown curve %d", curve);
  if (mem_type != BZ_MEM_HBM && mem_type != BZ_MEM_DMA) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown memory type %d", mem_type);
  bz_msm* m = new bz_msm();
  m->dc = dc;
  m->ops = ops;
  m->curve = curve;
  m->mem_type = mem_type;
  m->factor = is_precompute ? 8 : 1;   // PRECOMPUTE_FACTOR / PRECOMPUTE_FACTOR_BASE, msm_api.rs:39-40
  if (const char* e = getenv("BZ_MSM_PRECOMP")) { int v = atoi(e); if (v >= 0 && v <= 2) m->precomp_mode = v; }
  for (auto& slot : m->tev) for (auto& e : slot) cudaEventCreate(&e);
  cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
  {
    const char* e = getenv("BZ_MSM_TAIL");
    if (!e || atoi(e) != 0) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = greatest priority
      if (cudaStreamCreateWithPriority(&m->tail, cudaStreamNonBlocking, hi) != cudaSuccess) { m->tail = nullptr; cudaGetLastError(); }
      if (m->tail) {
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&m->ev_tail_done, cudaEventDisableTiming);
      }
    }
  }
  for (int b = 0; b < 2; b++) {
    cudaEventCreateWithFlags(&m->ev_copied[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&m->ev_consumed[b], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&m->ev_points_copied, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&m->ev_points_consumed, cudaEventDisableTiming);
  if (cudaHostAlloc((void**)&m->pinned, (size_t)RESULT_SLOTS * RESULT_SLOT_BYTES, cudaHostAllocPortable) != cudaSuccess) {
    leaf_free(m);
    return fail(BZ_ERR_NO_DEVICE, "pinned allocation failed");
  }
  *out = m;
  return BZ_OK;
}

static void result_release(MsmTaskResult& r) {
  if (r.done) cudaEventDestroy(r.done);
  r.done = nullptr; r.host_slot = nullptr; r.host_err = nullptr;
}

int32_t bz::leaf_free(bz_msm* m) {
  if (!m) return BZ_OK;
  cudaSetDevice(m->dc->device);
  cudaStreamSynchronize(m->dc->stream);
  if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
  if (m->tail) cudaStreamSynchronize(m->tail);
  for (auto& r : m->results) result_release(r);
  ws_free(m);
  if (m->table) cudaFree(m->table);
  wtable_free(m);
  if (m->dma_points) cudaFree(m->dma_points);
  if (m->comb_dev) cudaFree(m->comb_dev);
  if (m->ba2_scratch) cudaFree(m->ba2_scratch);
  for (int b = 0; b < 2; b++) {
    if (m->scalars_dev[b]) cudaFree(m->scalars_dev[b]);
    if (m->ev_copied[b]) cudaEventDestroy(m->ev_copied[b]);
    if (m->ev_consumed[b]) cudaEventDestroy(m->ev_consumed[b]);
  }
  if (m->ev_points_copied) cudaEventDestroy(m->ev_points_copied);
  if (m->ev_points_consumed) cudaEventDestroy(m->ev_points_consumed);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->tail) cudaStreamDestroy(m->tail);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_tail_done) cudaEventDestroy(m->ev_tail_done);
  if (m->pinned) cudaFreeHost(m->pinned);
  for (auto& slot : m->tev) for (auto& e : slot) if (e) cudaEventDestroy(e);
  for (auto& e : m->ev_part) if (e) cudaEventDestroy(e);
  delete m;
  return BZ_OK;
}

int32_t bz::leaf_initialize(bz_msm* m, uint32_t nof_elements, int32_t has_hbm_addr, uint64_t hbm_addr, uint64_t hbm_offset) {
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->mem_type == BZ_MEM_DMA && !has_hbm_addr) {
    m->hbm_mode = false;                      // BASES_SOURCE = 0, msm_api.rs:75-81
  } else {
    if (!has_hbm_addr)                        // the reference unwrap()-panics here (msm_api.rs:84)
      return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "HBM point memory needs hbm_point_addr");
    m->hbm_mode = true;                       // BASES_SOURCE = 1 + start address, msm_api.rs:82-95
    m->hbm_addr = hbm_addr;
    m->hbm_off = hbm_offset;
  }
  if (nof_elements == 0) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "nof_elements must be > 0");
  m->nof_elements = nof_elements;             // NUMBER_OF_MSM_ELEMENTS, msm_api.rs:103-108
  return BZ_OK;
}

// make sure the window-merged table matches the current plan (c, Wd) and the resident point set
static int32_t ensure_wtable(bz_msm* m, uint64_t n) {
  const MsmPlan& p = m->plan;
  if (m->wtable && m->wtable_n == n && m->wtable_c == p.c && m->wtable_levels == p.Wd && m->wtable_gen == m->table_gen)
    return BZ_OK;
  const size_t bytes = (size_t)p.Wd * n * m->ops->affine_bytes;
  if (m->wtable_bytes < bytes) {
    wtable_free(m);
    if (cudaMalloc(&m->wtable, bytes) != cudaSuccess) {
      cudaGetLastError();
      m->wtable = nullptr;
      return BZ_ERR_NO_RESULT;
    }
    m->wtable_bytes = bytes;
  }
  cudaStream_t st = m->dc->stream;
  nvtxRangePushA("blaze_b200: window-merged table build");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(m->wtable, m->table, n * m->ops->affine_bytes, cudaMemcpyDeviceToDevice, st));
  m->ops->build_wtable(m->wtable, n, p.Wd, p.c, st);
  cudaEventRecord(e1, st);
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
  // one-off and long (seconds at 2^26): wait for it here so that its duration can be reported (bz_msm_table_build_ms)
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&m->wtable_build_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  nvtxRangePop();
  m->wtable_n = n;
  m->wtable_c = p.c;
  m->wtable_levels = p.Wd;
  m->wtable_gen = m->table_gen;
  return BZ_OK;
}

static int32_t build_table(bz_msm* m, const uint8_t* raw_dev, uint64_t n_points);

int32_t bz::leaf_comb_reserve(bz_msm* m, size_t bytes) {
  if (m->comb_cap >= bytes) return BZ_OK;
  if (m->comb_dev) cudaFree(m->comb_dev);   // cudaFree waits for outstanding work
  m->comb_dev = nullptr;
  m->comb_cap = 0;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&m->comb_dev, bytes));
  m->comb_cap = bytes;
  return BZ_OK;
}

// Which bucket-accumulation kernel this task uses (decided per task: it does not touch the workspace, only the scratch of
// the batched-affine sweep, which is allocated on first use).  Fused batched-affine sweep (msm_ba2.cuh): pays when the
// bucket runs inside a segment are long enough to offer a few dozen independent additions per tree round.
static int32_t plan_accumulate(bz_msm* m) {
  MsmPlan& p = m->plan;
  if (p.batch_affine == 1) return BZ_OK;   // round 1's multi-kernel phases (BZ_MSM_BA=1): workspace-bound, left alone
  int want = -1, rounds = -1;
  if (const char* e = getenv("BZ_MSM_BA2")) want = atoi(e) ? 1 : 0;
  if (const char* e = getenv("BZ_MSM_BA2_ROUNDS")) rounds = atoi(e);
  if (m->acc_mode >= 0) want = m->acc_mode == 2 ? 1 : 0;
  if (m->acc_rounds >= 0) rounds = m->acc_rounds;
  const uint32_t L = p.seg_len;
  const uint64_t total = (uint64_t)p.W * p.Ms;
  const double avg = (double)total / ((double)p.W * (double)p.nb);   // entries per bucket (upper bound: zero digits drop out)
  const bool on = want >= 0 ? want != 0 : (BZ_MSM_BA2_DEFAULT && avg >= 16.0 && L >= 64);
  p.batch_affine = 0;
  if (!on) return BZ_OK;
  // tree rounds: while the runs still hold pairs (avg / 2^r >= 2) and a round still offers >= 16 additions per thread
  int r = 0;
  while (r < 6 && avg / (double)(1u << r) >= 2.0 && (double)L / (double)(2u << r) >= 16.0) r++;
  r = std::max(1, r);
  if (rounds >= 0) r = std::min(rounds, 8);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->dc->device);
  const uint64_t want_ctas = (uint64_t)sms * BZ_BA2_CTAS_PER_SM;
  p.ba_rounds = r;
  p.ba_cap = L / 2 + 32;
  p.ba_ctas = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want_ctas, (p.nseg + 127) / 128));
  // three arrays (x, y, running products) of ba_cap slots per lane, one set per resident warp
  const size_t bytes = (size_t)3 * p.ba_cap * m->ops->fq_bytes * 32 * p.ba_ctas * 4;
  if (m->ba2_cap < bytes) {
    if (m->ba2_scratch) cudaFree(m->ba2_scratch);   // waits for outstanding work
    m->ba2_scratch = nullptr;
    m->ba2_cap = 0;
    if (cudaMalloc(&m->ba2_scratch, bytes) != cudaSuccess) {
      cudaGetLastError();
      return BZ_OK;   // no room for the scratch: the XYZZ sweep needs none
    }
    m->ba2_cap = bytes;
  }
  m->ws.ba2_scratch = m->ba2_scratch;
  p.batch_affine = 2;
  return BZ_OK;
}

// enqueue the whole pipeline for one task on the client's stream
static int32_t launch_task(bz_msm* m) {
  bz_dclient* dc = m->dc;
  const uint64_t M = m->data_M;
  const int wps = m->factor == 8 ? 1 : 8;
  cudaStream_t st = dc->stream;
  // nothing is touched before we know the task can be queued (a refused task must not burn a label or a slot)
  if ((int)m->results.size() >= RESULT_SLOTS) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "more than %d results pending; pop them with result()", RESULT_SLOTS);
  // window-merged table: only for a resident (arena) point set with full-width scalars, and by default only once
  // the same point set is used a second time (building it costs about as much as a dozen plain MSMs)
  bool merged = false;
  int32_t rc;
  if (m->table_from_arena && wps == 8 && !m->precomp_failed &&
      (m->precomp_mode == 2 || (m->precomp_mode == 1 && m->table_uses >= 1))) {
    rc = make_plan(m, M, wps, true);
    if (rc == BZ_OK) {
      rc = ensure_wtable(m, M);
      if (rc == BZ_OK) merged = true;
    }
    if (!merged) m->precomp_failed = true;
  }
  if (!merged) {
    rc = make_plan(m, M, wps, false);
    if (rc) return rc;
  }
  rc = plan_accumulate(m);
  if (rc) return rc;
  const size_t rs = 3 * (size_t)m->ops->fq_bytes;
  const bool ranked = dc->comm != nullptr && dc->world > 1;
  if (ranked) { rc = leaf_comb_reserve(m, rs * (dc->world + 1)); if (rc) return rc; }
  MsmTaskResult r;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaEventCreateWithFlags(&r.done, cudaEventBlockingSync | cudaEventDisableTiming));
  if (m->table_from_arena) m->table_uses++;
  r.label = m->next_label++;
  m->last_label = r.label;
  r.slot = (int)(r.label % RESULT_SLOTS);   // labels in flight are consecutive and at most RESULT_SLOTS: slots are distinct
  r.host_slot = m->pinned + (size_t)r.slot * RESULT_SLOT_BYTES;
  r.host_err = reinterpret_cast<int*>(r.host_slot + 256);
  cudaEvent_t* ev = m->tev[r.slot];
  nvtxRangePushA("blaze_b200: MSM task enqueue");
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaMemsetAsync(m->ws.err, 0, 4, st));
  m->ws.ev_acc0 = ev[2];
  m->ws.ev_acc1 = ev[3];
  if (m->stage_cur >= 0) CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamWaitEvent(st, m->ev_copied[m->stage_cur], 0));
  cudaEventRecord(ev[0], st);
  launch_msm_sort(m->plan, m->ws, m->scalars_src, st);
  // the error word is final once the digits are extracted (k_digits is its only writer): read it back here, on the work
  // stream, before the next task's memset can reach it -- the rest of this task may finish on the tail stream
  CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(r.host_err, m->ws.err, 4, cudaMemcpyDeviceToHost, st));
  if (m->stage_cur >= 0) {   // k_digits (the only reader of the staging buffer) is queued: mark it consumed
    cudaEventRecord(m->ev_consumed[m->stage_cur], st);
    m->consumed_valid[m->stage_cur] = true;
  }
  if (m->table_pending) {   // DMA mode: the points were copied behind the scalars; convert them now
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamWaitEvent(st, m->ev_points_copied, 0));
    rc = build_table(m, m->dma_points, m->table_pending_n);
    if (rc) return rc;
    cudaEventRecord(m->ev_points_consumed, st);
    m->points_consumed_valid = true;
    m->table_pending = false;
  }
  cudaEventRecord(ev[1], st);
  m->plan.raw_result = (m->raw_result || ranked) ? 1 : 0;
  m->ws.tail = m->tail;
  m->ws.ev_fork = m->ev_fork;
  m->ws.ev_tail_done = m->ev_tail_done;
  m->ws.tail_busy = m->tail_busy ? 1 : 0;
  const cudaStream_t work = st;
  st = m->ops->bucket_phase(m->plan, m->ws, merged ? m->wtable : m->table, st);   // from here on: work or tail stream
  m->result_stream = st;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
  const uint8_t* final_rec = m->ws.result;
  if (ranked) {
    // final exchange of the point-sharded MSM, device side: all-gather of the ranks' projective records straight into
    // the combine kernel on the same stream (no host bounce); every rank ends with the same normalised (or raw) sum
    rc = comm_allgather(dc, m->ws.result, m->comb_dev, rs, st);
    if (rc) return rc;
    m->ops->combine_results(m->comb_dev, dc->world, m->comb_dev + rs * dc->world, m->raw_result, st);
    g_kernel_launches += 1;
    final_rec = m->comb_dev + rs * dc->world;
  }
  cudaEventRecord(ev[4], st);
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
  CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(r.host_slot, final_rec, rs, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaEventRecord(r.done, st));
  if (st != work) {
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaEventRecord(m->ev_tail_done, st));
    m->tail_busy = true;
  }
  nvtxRangePop();
  m->results.push_back(r);
  m->pending_tasks--;
  m->data_ready = false;
  m->launched++;
  return BZ_OK;
}

int32_t bz::leaf_start_process(bz_msm* m) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  m->pending_tasks++;                          // PUSH_MSM_TASK_TO_QUEUE, msm_api.rs:113-120
  if (m->data_ready) {
    rc = launch_task(m);
    if (rc) m->pending_tasks--;                // the refused task is not left queued
    return rc;
  }
  return BZ_OK;
}

// make the Montgomery table for `n_points` wire points starting at raw (device pointer)
static int32_t build_table(bz_msm* m, const uint8_t* raw_dev, uint64_t n_points) {
  if (m->table_cap < n_points) {
    if (m->table) cudaFree(m->table);
    m->table = nullptr;
    m->table_cap = 0;
    CUDA_TRY(BZ_ERR_WRITE, cudaMalloc(&m->table, n_points * m->ops->affine_bytes));
    m->table_cap = n_points;
  }
  m->ops->points_to_mont(raw_dev, m->table, n_points, m->dc->stream);
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
  m->table_n = n_points;
  m->table_gen++;
  return BZ_OK;
}

static int32_t ensure_arena_table(bz_msm* m, uint64_t addr, uint64_t n_points) {
  bz_dclient* dc = m->dc;
  const uint64_t bytes = n_points * 2ull * m->ops->fq_bytes;
  if (addr & 15) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "hbm point address must be 16-byte aligned");
  if (addr + bytes < addr) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "hbm point range overflows");
  // the card lock is held across the table build: nothing can write the range between the check and the launch, and
  // the address space never moves (virtual reservation), so the kernel's pointer stays valid after the unlock
  std::lock_guard<std::mutex> lk(dc->mu);
  if (m->table_from_arena && m->table_addr == addr && m->table_n == n_points &&
      !arena_dirty_since(dc, m->table_epoch, addr, addr + bytes)) {
    m->table_epoch = dc->epoch;   // nothing overlapping was written: keep the window of the write log short
    return BZ_OK;
  }
  int32_t rc = arena_map(dc, addr, addr + bytes);   // unwritten HBM reads as zeros = identity padding
  if (rc) return rc;
  wtable_free(m);   // stale: derived from the previous point set
  m->precomp_failed = false;
  m->table_uses = 0;
  rc = build_table(m, dc->arena + addr, n_points);
  if (rc) return rc;
  m->table_from_arena = true;
  m->table_addr = addr;
  m->table_epoch = dc->epoch;
  return BZ_OK;
}

static int32_t stage_scalars(bz_msm* m, const uint8_t* scalars, size_t len) {
  const int b = m->stage_next;
  m->stage_next ^= 1;
  if (m->scalars_cap[b] < len) {
    if (m->scalars_dev[b]) cudaFree(m->scalars_dev[b]);   // cudaFree waits for outstanding work
    m->scalars_dev[b] = nullptr;
    m->scalars_cap[b] = 0;
    m->consumed_valid[b] = false;
    CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&m->scalars_dev[b], len));
    m->scalars_cap[b] = len;
  }
  // do not overwrite the buffer before the task that used it last has read it
  if (m->consumed_valid[b]) CUDA_TRY(BZ_ERR_WRITE, cudaStreamWaitEvent(m->copy_stream, m->ev_consumed[b], 0));
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(m->scalars_dev[b], scalars, len, cudaMemcpyHostToDevice, m->copy_stream));
  CUDA_TRY(BZ_ERR_WRITE, cudaEventRecord(m->ev_copied[b], m->copy_stream));
  // (the host blocks on the copy stream at the end of set_data, after the task has been queued)
  m->scalars_src = m->scalars_dev[b];
  m->stage_cur = b;
  return BZ_OK;
}

int32_t bz::leaf_sync_copies(bz_msm* m) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  cudaError_t ce = cudaStreamSynchronize(m->copy_stream);
  if (ce != cudaSuccess) return fail(BZ_ERR_WRITE, "host-to-device copy failed: %s", cudaGetErrorString(ce));
  return BZ_OK;
}

int32_t bz::leaf_set_data(bz_msm* m, const uint8_t* points, size_t points_len, const uint8_t* scalars_host,
                          uint64_t scalars_dev_ptr, size_t scalars_len, uint32_t nof_elements, int32_t has_hbm_addr,
                          uint64_t hbm_addr, uint64_t hbm_offset, bool sync_host) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  const uint64_t n = nof_elements;
  const uint64_t ps = 2ull * m->ops->fq_bytes * m->factor;   // bytes per element record
  if (n == 0) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "nof_elements must be > 0");
  // The reference trusts lengths (last chunk is "whatever is left", msm_api.rs:166-170); we validate.
  if (scalars_len != n * 32) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "scalars length %zu != %llu*32", scalars_len, (unsigned long long)n);
  if (points && points_len != n * ps) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "points length %zu != %llu*%llu", points_len, (unsigned long long)n, (unsigned long long)ps);
  if (!points && !has_hbm_addr) return BZ_OK;   // (None, None): the reference silently does nothing (msm_api.rs:163-216)
  const uint64_t npts = n * m->factor;
  if (npts >= (1ull << 31)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "too many points");
  bool dma_points_now = false;
  if (points && !has_hbm_addr) {
    // DMA mode: points streamed with the call (msm_api.rs:175-202)
    if (m->dma_points_cap < points_len) {
      if (m->dma_points) cudaFree(m->dma_points);
      m->dma_points = nullptr;
      m->dma_points_cap = 0;
      CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&m->dma_points, points_len));
      m->dma_points_cap = points_len;
    }
    m->table_from_arena = false;
    dma_points_now = true;   // copied below, behind the scalars
  } else {
    uint64_t a = hbm_addr + hbm_offset;
    if (points) {
      // preload to HBM, then scalars (msm_api.rs:203-216)
      rc = bz_dclient_dma_write(m->dc, hbm_addr, hbm_offset, points, points_len);
      if (rc) return rc;
      m->hbm_mode = true; m->hbm_addr = hbm_addr; m->hbm_off = hbm_offset;
    }
    // an HBM-mode input REPLACES a streamed (DMA) point set that no task has consumed yet: its deferred table build
    // must not run over the table derived from HBM below
    m->table_pending = false;
    rc = ensure_arena_table(m, a, npts);
    if (rc) return rc;
  }
  if (scalars_host) {
    rc = stage_scalars(m, scalars_host, scalars_len);
    if (rc) return rc;
  } else {
    m->stage_cur = -1;
    m->scalars_src = reinterpret_cast<const uint32_t*>(scalars_dev_ptr);
  }
  if (dma_points_now) {
    // do not overwrite the point buffer before the table of the previous task has been built from it
    if (m->points_consumed_valid) CUDA_TRY(BZ_ERR_WRITE, cudaStreamWaitEvent(m->copy_stream, m->ev_points_consumed, 0));
    CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(m->dma_points, points, points_len, cudaMemcpyHostToDevice, m->copy_stream));
    CUDA_TRY(BZ_ERR_WRITE, cudaEventRecord(m->ev_points_copied, m->copy_stream));
    m->table_pending = true;
    m->table_pending_n = npts;
  }
  m->data_M = npts;
  m->data_ready = true;
  rc = BZ_OK;
  if (m->pending_tasks > 0) rc = launch_task(m);
  // the caller may drop its buffers when we return (move-in semantics): block the HOST on the copy stream only --
  // the work stream keeps running (the previous task, or this task's digits + sort while the points still travel)
  if (sync_host && (scalars_host || dma_points_now)) {
    cudaError_t ce = cudaStreamSynchronize(m->copy_stream);
    if (ce != cudaSuccess && rc == BZ_OK) rc = fail(BZ_ERR_WRITE, "host-to-device copy failed: %s", cudaGetErrorString(ce));
  }
  return rc;
}

static int32_t collect(bz_msm* m, MsmTaskResult& r) {
  if (r.collected) return r.status;
  CUDA_TRY(BZ_ERR_READ, cudaEventSynchronize(r.done));
  r.bytes.assign(r.host_slot, r.host_slot + 3 * m->ops->fq_bytes);
  int err = *r.host_err;
  r.status = err == BZ_ERR_NONE ? BZ_OK : BZ_ERR_INVALID_PRIMITIVE_PARAM;
  r.collected = true;
  {
    cudaEvent_t* ev = m->tev[r.slot];
    float a = 0, b = 0, k = 0;
    cudaEventElapsedTime(&a, ev[0], ev[4]);
    cudaEventElapsedTime(&b, ev[0], ev[1]);
    cudaEventElapsedTime(&k, ev[2], ev[3]);
    m->last_ms[0] = a;
    m->last_ms[1] = b;
    m->last_ms[2] = k;           // bucket accumulation kernel(s) alone
    m->last_ms[3] = a - b - k;   // memset + merge + reduce + finish (+ exchange)
    m->tasks_done++;
  }
  if (r.status) return fail(r.status, "device flagged a non-canonical scalar (>= r)");
  return BZ_OK;
}

int32_t bz::leaf_wait_result(bz_msm* m) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->results.empty()) return fail(BZ_ERR_NO_RESULT, "no task in flight (the reference would spin forever on RESULT_VALID)");
    MsmTaskResult& r = m->results.front();
    if (r.collected) return collect(m, r);
    ev = r.done;
  }
  // block WITHOUT the client lock: a producer thread may keep queueing tasks (set_data / start_process) meanwhile
  CUDA_TRY(BZ_ERR_READ, cudaEventSynchronize(ev));
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->results.empty() || m->results.front().done != ev) return BZ_OK;   // another thread has already popped it
  return collect(m, m->results.front());
}

int32_t bz::leaf_result(bz_msm* m, uint8_t* out, size_t out_len, uint32_t* result_label) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->results.empty()) return fail(BZ_ERR_NO_RESULT, "result queue is empty");
  size_t need = 3 * (size_t)m->ops->fq_bytes;
  if (out && out_len < need) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "result buffer too small (%zu < %zu)", out_len, need);
  MsmTaskResult& r = m->results.front();
  rc = collect(m, r);
  if (rc == BZ_OK) {
    if (out) memcpy(out, r.bytes.data(), need);
    if (result_label) *result_label = r.label;
  }
  result_release(r);
  m->results.pop_front();                      // POP_RESULT, msm_api.rs:265-269
  return rc;
}

int32_t bz::leaf_load_data_to_hbm(bz_msm* m, const uint8_t* points, size_t len, uint64_t addr, uint64_t offset) {
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->hbm_mode = true;                        // BASES_SOURCE = 1 + start address, msm_api.rs:299-311
    m->hbm_addr = addr;
  }
  return bz_dclient_dma_write(m->dc, addr, offset, points, len);
}

int32_t bz::leaf_phase_times(bz_msm* m, float ms[4]) {
  std::lock_guard<std::mutex> lk(m->mu);
  memcpy(ms, m->last_ms, sizeof(m->last_ms));
  return BZ_OK;
}

int32_t bz::leaf_combine_results(bz_msm* m, const uint8_t* records, int32_t n, uint8_t* out, size_t out_len) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  size_t rs = 3 * (size_t)m->ops->fq_bytes;
  if (out_len < rs) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "result buffer too small");
  std::lock_guard<std::mutex> lk(m->mu);
  rc = leaf_comb_reserve(m, rs * (n + 1));   // small scratch kept across calls (no cudaMalloc / cudaFree per step)
  if (rc) return rc;
  if (m->tail) cudaStreamSynchronize(m->tail);   // a ranked task's tail uses the same scratch
  uint8_t* d = m->comb_dev;
  cudaStream_t st = m->dc->stream;
  cudaMemcpyAsync(d, records, rs * n, cudaMemcpyHostToDevice, st);
  m->ops->combine_results(d, n, d + rs * n, 0, st);
  g_kernel_launches += 1;
  cudaMemcpyAsync(out, d + rs * n, rs, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(BZ_ERR_UNKNOWN, "combine failed: %s", cudaGetErrorString(e));
  return BZ_OK;
}

int32_t bz::leaf_generate_chain_points(bz_msm* m, const uint8_t* p0q, size_t p0q_len, uint64_t first, uint64_t n,
                                       uint64_t addr, uint64_t offset) {
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  size_t ps = 2 * (size_t)m->ops->fq_bytes;
  if (p0q_len != 2 * ps) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "p0q must be two wire points");
  if (m->factor != 1) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "chain generator writes factor-1 bases");
  bz_dclient* dc = m->dc;
  uint64_t a = addr + offset;
  if (a & 15) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "address must be 16-byte aligned");
  std::lock_guard<std::mutex> lk(dc->mu);
  rc = arena_map(dc, a, a + n * ps);
  if (rc) return rc;
  uint8_t* d = nullptr;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&d, 2 * ps));
  cudaMemcpyAsync(d, p0q, 2 * ps, cudaMemcpyHostToDevice, dc->stream);
  m->ops->gen_chain_points(d, first, n, dc->arena + a, dc->stream);
  cudaError_t e = cudaStreamSynchronize(dc->stream);
  cudaFree(d);
  arena_note_write(dc, a, a + n * ps);
  if (e != cudaSuccess) return fail(BZ_ERR_UNKNOWN, "generator failed: %s", cudaGetErrorString(e));
  return BZ_OK;
}
