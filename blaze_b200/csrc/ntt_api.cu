// NTTClient half of the C ABI (include/blaze_b200.h): buffer slots, pass planning, launches.
//
// Mirrors the behaviour of /root/reference/src/ingo_ntt/ntt_api.rs: two buffer slots
// (ntt_data.rs:42,54-56), set_data(buf_host) fills a slot, start_process(buf_kernel) transforms a
// slot IN PLACE, result(buf) reads a slot back (integration_ntt.rs:48-55); the flat vector of
// 32-byte little-endian elements is the public I/O format (ntt_api.rs:20-23, README.md:118).  The
// host-side bank permutation (NTTBanks::preprocess/postprocess, ntt_data.rs:80-157) is the FPGA's
// internal HBM layout and has no counterpart here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/blaze_b200.h"
#include "api_common.h"
#include "client_internal.h"
#include "ntt_internal.h"

using namespace bz;

struct bz_ntt {
  bz_dclient* dc = nullptr;
  int field = 2;          // curve code of the scalar field (2 = BLS12-381 Fr)
  int log_n = 27;
  int inverse = 0;
  uint64_t n = 0;
  // each slot has two physical buffers (Stockham ping-pong); cur[s] says which one holds the data
  uint4* buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int cur[2] = {0, 0};
  uint4* tab_mem = nullptr;
  NttTables tab{};
  bool tables_ready = false;
  std::vector<int> radices;
  std::vector<uint4*> tw_full;   // per pass: twiddles in the pass' input layout (null for the first pass)
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t done = nullptr;
  bool launched = false;
  float last_ms = 0;
  int* err_dev = nullptr;    // device flag: a transform saw a non-canonical input element
  int* err_host = nullptr;   // pinned copy, refreshed behind every transform
  // Copies run on their own streams so that the H2D of one slot and the D2H of the other overlap the transform
  // (the reference's double-buffer pipeline, integration_ntt.rs:103-136).  Per slot: input landed / transform
  // finished / output read -- each stream waits only for what it needs.
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  bool in_valid[2] = {false, false}, cmp_valid[2] = {false, false}, out_valid[2] = {false, false};
  std::mutex mu;
};

static std::vector<int> plan_radices(int log_n) {
  std::vector<int> r;
  if (log_n == 0) return r;
  int npass = (log_n + 8) / 9, base = log_n / npass, extra = log_n % npass;
  for (int i = 0; i < npass; i++) r.push_back(i < extra ? base + 1 : base);
  return r;
}

static NttPassParams ntt_pass_params(bz_ntt* t, size_t p, uint64_t Ns);

static int32_t ntt_alloc_slot(bz_ntt* t, int s) {
  bool fresh = false;
  for (int k = 0; k < 2; k++)
    if (!t->buf[s][k]) {
      CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&t->buf[s][k], std::max<uint64_t>(t->n, 1) * 32));
      CUDA_TRY(BZ_ERR_WRITE, cudaMemsetAsync(t->buf[s][k], 0, std::max<uint64_t>(t->n, 1) * 32, dc_stream(t->dc)));
      fresh = true;
    }
  if (fresh) CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(dc_stream(t->dc)));   // zero fill before any copy stream touches it
  return BZ_OK;
}

static int32_t ntt_new_common(bz_dclient* dc, int field, int log_n, int inverse, bz_ntt** out) {
  if (!out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  *out = nullptr;
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  if (field < 0 || field > 2) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown field %d", field);
  if (log_n < 0 || log_n > ntt_two_adicity(field) || log_n > 30)
    return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "log size %d unsupported for this field (two-adicity %d)", log_n, ntt_two_adicity(field));
  bz_ntt* t = new bz_ntt();
  t->dc = dc;
  t->field = field;
  t->log_n = log_n;
  t->inverse = inverse ? 1 : 0;
  t->n = 1ull << log_n;
  t->radices = plan_radices(log_n);
  cudaEventCreate(&t->ev[0]);
  cudaEventCreate(&t->ev[1]);
  cudaEventCreateWithFlags(&t->done, cudaEventBlockingSync | cudaEventDisableTiming);
  cudaStreamCreateWithFlags(&t->h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&t->d2h, cudaStreamNonBlocking);
  cudaMalloc((void**)&t->err_dev, 16);
  cudaHostAlloc((void**)&t->err_host, 16, cudaHostAllocPortable);
  if (t->err_host) *t->err_host = 0;
  for (int s = 0; s < 2; s++) {
    cudaEventCreateWithFlags(&t->ev_in[s], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&t->ev_cmp[s], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&t->ev_out[s], cudaEventDisableTiming);
  }
  *out = t;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_new(bz_dclient* dc, int32_t ntt_type, bz_ntt** out) {
  if (ntt_type != 0) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown NTT type %d", ntt_type);
  return ntt_new_common(dc, BZ_CURVE_BLS381, 27, 0, out);   // the reference core: fixed 2^27 (ntt_data.rs:65-66)
}
extern "C" int32_t bz_ntt_new_ex(bz_dclient* dc, int32_t field, int32_t log_size, int32_t inverse, bz_ntt** out) {
  return ntt_new_common(dc, field, log_size, inverse, out);
}

extern "C" int32_t bz_ntt_free(bz_ntt* t) {
  if (!t) return BZ_OK;
  cudaSetDevice(dc_device(t->dc));
  cudaStreamSynchronize(dc_stream(t->dc));
  if (t->h2d) { cudaStreamSynchronize(t->h2d); cudaStreamDestroy(t->h2d); }
  if (t->d2h) { cudaStreamSynchronize(t->d2h); cudaStreamDestroy(t->d2h); }
  for (int s = 0; s < 2; s++)
    for (cudaEvent_t e : {t->ev_in[s], t->ev_cmp[s], t->ev_out[s]}) if (e) cudaEventDestroy(e);
  for (auto& s : t->buf) for (auto& b : s) if (b) cudaFree(b);
  if (t->tab_mem) cudaFree(t->tab_mem);
  if (t->err_dev) cudaFree(t->err_dev);
  if (t->err_host) cudaFreeHost(t->err_host);
  for (auto p : t->tw_full) if (p) cudaFree(p);
  for (auto& e : t->ev) if (e) cudaEventDestroy(e);
  if (t->done) cudaEventDestroy(t->done);
  delete t;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_loaded_binary_parameters(bz_ntt* t, uint32_t out[2]) {
  // `todo!()` in the reference (ntt_api.rs:33-35); we report {field code | direction, log size}
  if (!t || !out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  out[0] = 0xB2000100u | (uint32_t)t->field | ((uint32_t)t->inverse << 4);
  out[1] = (uint32_t)t->log_n;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_initialize(bz_ntt* t) {
  // The reference programs a debug-program register block (ntt_api.rs:37-56); here: build the
  // twiddle tables once.  Callable every iteration (integration_ntt.rs:47) or once (:100).
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null NTTClient");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(t->mu);
  if (t->tables_ready) return BZ_OK;
  const int lo_bits = 14;
  uint32_t nlo = 1u << std::min(t->log_n, lo_bits);
  uint32_t nhi = t->log_n > lo_bits ? 1u << (t->log_n - lo_bits) : 1u;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&t->tab_mem, ((size_t)nlo + nhi + 1) * 32));
  t->tab.lo = t->tab_mem;
  t->tab.hi = t->tab_mem + 2 * (size_t)nlo;
  t->tab.ninv = t->tab_mem + 2 * ((size_t)nlo + nhi);
  t->tab.lo_bits = lo_bits;
  t->tab.log_root = t->log_n;
  ntt_gen_tables(t->field, t->tab, t->log_n, t->inverse, dc_stream(t->dc));
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
  // Precompute the inter-pass twiddles in each pass' input layout ("twiddles precomputed", one
  // coalesced 32-B read per element instead of a two-level table product).  Costs n x 32 B per twiddled
  // pass; skipped (two-level tables only) when BZ_NTT_FULL_TW=0 or the allocation fails.
  const char* env = getenv("BZ_NTT_FULL_TW");
  if (!(env && atoi(env) == 0) && t->log_n >= 10) {
    t->tw_full.assign(t->radices.size(), nullptr);
    uint64_t Ns = 1;
    for (size_t p = 0; p < t->radices.size(); p++) {
      if (Ns > 1) {
        uint4* buf = nullptr;
        if (cudaMalloc((void**)&buf, t->n * 32) != cudaSuccess) { cudaGetLastError(); break; }
        NttPassParams P = ntt_pass_params(t, p, Ns);
        P.tw_full_out = buf;
        cudaError_t e = ntt_launch_pass(t->field, P, dc_stream(t->dc));
        if (e != cudaSuccess) { cudaFree(buf); return bz_fail(BZ_ERR_UNKNOWN, "twiddle table pass failed: %s", cudaGetErrorString(e)); }
        t->tw_full[p] = buf;
      }
      Ns <<= t->radices[p];
    }
  }
  t->tables_ready = true;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_set_data(bz_ntt* t, size_t buf_host, const uint8_t* data, size_t len) {
  if (!t || !data) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (buf_host > 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "buf_host must be 0 or 1");
  if (len != t->n * 32) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "data length %zu != %llu*32", len, (unsigned long long)t->n);
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::unique_lock<std::mutex> lk(t->mu);
  rc = ntt_alloc_slot(t, (int)buf_host);
  if (rc) return rc;
  // the slot must be idle: its last transform finished, its last result read out (stream-side waits only)
  cudaStream_t st = t->h2d;
  if (t->cmp_valid[buf_host]) CUDA_TRY(BZ_ERR_WRITE, cudaStreamWaitEvent(st, t->ev_cmp[buf_host], 0));
  if (t->out_valid[buf_host]) CUDA_TRY(BZ_ERR_WRITE, cudaStreamWaitEvent(st, t->ev_out[buf_host], 0));
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(t->buf[buf_host][t->cur[buf_host]], data, len, cudaMemcpyHostToDevice, st));
  CUDA_TRY(BZ_ERR_WRITE, cudaEventRecord(t->ev_in[buf_host], st));
  t->in_valid[buf_host] = true;
  // The host blocks on the copy WITHOUT the client lock: another thread may queue result() of the other slot meanwhile
  // (PCIe is full duplex: with a feeder and a drainer thread the two directions overlap, bench.py ntt.e2e.threaded_ms).
  lk.unlock();
  CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(st));   // caller may drop `data` (move-in semantics); the work stream keeps running
  return BZ_OK;
}

// parameters of pass p (sub-transform size Ns) of the single-GPU transform
static NttPassParams ntt_pass_params(bz_ntt* t, size_t p, uint64_t Ns) {
  const uint64_t L = t->n;
  const int lr = t->radices[p];
  const uint64_t R = 1ull << lr;
  NttPassParams P;
  memset(&P, 0, sizeof(P));
  P.lr = lr;
  P.Q = L / R;
  P.in_sr = L / R;
  P.otw_rsel = -1;
  P.tab = t->tab;
  if (Ns == 1) {
    P.Q0 = P.Q; P.Q1 = 1;
    P.in_s0 = 1;
    P.out_s0 = R; P.out_sr = 1;
    P.tw_sel = -1;
    P.store_k_fastest = 1;
  } else {
    P.Q0 = Ns; P.Q1 = L / (R * Ns);
    P.in_s0 = 1; P.in_s1 = Ns;
    P.out_s0 = 1; P.out_s1 = Ns * R; P.out_sr = Ns;
    P.tw_sel = 0;
    P.tw_scale = (1ull << t->log_n) / (Ns * R);
    P.store_k_fastest = 0;
  }
  P.scale_ninv = (t->inverse && p + 1 == t->radices.size()) ? 1 : 0;
  P.lq0 = 0; while ((1ull << P.lq0) < P.Q0) P.lq0++;
  P.lq1 = 0; while ((1ull << P.lq1) < P.Q1) P.lq1++;
  return P;
}

// enqueue all passes of one transform of slot s
static int32_t ntt_enqueue(bz_ntt* t, int s) {
  cudaStream_t st = dc_stream(t->dc);
  uint64_t Ns = 1;
  if (t->in_valid[s]) cudaStreamWaitEvent(st, t->ev_in[s], 0);     // the slot's input has landed
  if (t->out_valid[s]) cudaStreamWaitEvent(st, t->ev_out[s], 0);   // nobody is still reading the slot out
  cudaMemsetAsync(t->err_dev, 0, 4, st);
  cudaEventRecord(t->ev[0], st);
  for (size_t p = 0; p < t->radices.size(); p++) {
    NttPassParams P = ntt_pass_params(t, p, Ns);
    if (p == 0) P.err = t->err_dev;
    P.in = t->buf[s][t->cur[s]];
    P.out = t->buf[s][t->cur[s] ^ 1];
    P.tw_full = p < t->tw_full.size() ? t->tw_full[p] : nullptr;
    cudaError_t e = ntt_launch_pass(t->field, P, st);
    if (e != cudaSuccess) return bz_fail(BZ_ERR_UNKNOWN, "NTT pass launch failed: %s", cudaGetErrorString(e));
    t->cur[s] ^= 1;
    Ns <<= t->radices[p];
  }
  cudaEventRecord(t->ev[1], st);
  cudaMemcpyAsync(t->err_host, t->err_dev, 4, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(t->done, st);
  cudaEventRecord(t->ev_cmp[s], st);
  t->cmp_valid[s] = true;
  t->launched = true;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_start_process(bz_ntt* t, size_t buf_kernel) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null NTTClient");
  if (buf_kernel > 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "buf_kernel must be 0 or 1");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> lk(t->mu);
    if (!t->tables_ready) { /* tolerate a missing initialize(): the reference's is register pokes only */ }
  }
  if (!t->tables_ready) { rc = bz_ntt_initialize(t); if (rc) return rc; }
  std::lock_guard<std::mutex> lk(t->mu);
  rc = ntt_alloc_slot(t, (int)buf_kernel);
  if (rc) return rc;
  return ntt_enqueue(t, (int)buf_kernel);
}

extern "C" int32_t bz_ntt_wait_result(bz_ntt* t) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null NTTClient");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::unique_lock<std::mutex> lk(t->mu);
  if (!t->launched) return bz_fail(BZ_ERR_NO_RESULT, "no transform in flight");
  cudaEvent_t done = t->done;
  lk.unlock();   // block without the client lock (the reference busy-polls a register here): copies of the other slot go on
  CUDA_TRY(BZ_ERR_READ, cudaEventSynchronize(done));
  lk.lock();
  cudaEventElapsedTime(&t->last_ms, t->ev[0], t->ev[1]);
  if (t->err_host && *t->err_host)
    return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "the transformed vector held a non-canonical element (>= r); its output is undefined");
  return BZ_OK;
}

extern "C" int32_t bz_ntt_result(bz_ntt* t, size_t buf_num, uint8_t* out, size_t out_len) {
  if (!t || !out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (buf_num > 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "buf_num must be 0 or 1");
  if (out_len < t->n * 32) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "result buffer too small");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::unique_lock<std::mutex> lk(t->mu);
  rc = ntt_alloc_slot(t, (int)buf_num);
  if (rc) return rc;
  cudaStream_t st = t->d2h;
  if (t->cmp_valid[buf_num]) CUDA_TRY(BZ_ERR_READ, cudaStreamWaitEvent(st, t->ev_cmp[buf_num], 0));
  if (t->in_valid[buf_num]) CUDA_TRY(BZ_ERR_READ, cudaStreamWaitEvent(st, t->ev_in[buf_num], 0));
  CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(out, t->buf[buf_num][t->cur[buf_num]], t->n * 32, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(BZ_ERR_READ, cudaEventRecord(t->ev_out[buf_num], st));
  t->out_valid[buf_num] = true;
  lk.unlock();
  CUDA_TRY(BZ_ERR_READ, cudaStreamSynchronize(st));   // blocks the host on this copy only, without the client lock
  return BZ_OK;
}

extern "C" int32_t bz_ntt_phase_times(bz_ntt* t, float* total_ms, uint32_t* passes) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null NTTClient");
  if (total_ms) *total_ms = t->last_ms;
  if (passes) *passes = (uint32_t)t->radices.size();
  return BZ_OK;
}

extern "C" int32_t bz_ntt_slot_device_ptr(bz_ntt* t, size_t buf_num, uint64_t* dev_ptr) {
  if (!t || !dev_ptr || buf_num > 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(t->mu);
  rc = ntt_alloc_slot(t, (int)buf_num);
  if (rc) return rc;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(dc_stream(t->dc)));
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(t->h2d));
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(t->d2h));
  *dev_ptr = (uint64_t)(uintptr_t)t->buf[buf_num][t->cur[buf_num]];
  return BZ_OK;
}

// =====================================================================================================
// Multi-GPU four-step NTT (B200 addition).  One process per GPU; rank g of G.
//   N = N1 x N2, input index j = j1 N2 + j2, output index k = k1 + N1 k2.
//   rank g holds the column slab  A_g[j1][c] = in[j1 N2 + g C + c],  C = N2 / G        (N1 x C)
//   step 1: length-N1 transforms down the columns (1-2 passes); the LAST pass multiplies by the
//           four-step twiddle w^((g C + c) k1) and stores row k1 straight into the owner's exchange
//           buffer B_h[k1 mod T][g C + c], h = k1 / T, T = N1 / G -- peer stores over NVLink: the
//           all-to-all transpose is fused into the kernel epilogue, there is no separate collective.
//   (host barrier between the steps: every rank must have finished writing every B)
//   step 3: length-N2 transforms along the rows of B_h (1-2 passes); result O_h[k2][t] = X[(h T + t) + N1 k2].
// In/out layouts are the strided slabs above (cudaMemcpy2D from / to a natural-order host vector).
struct bz_ntt_dist {
  bz_dclient* dc = nullptr;
  int field = 2, log_n = 0, inverse = 0, rank = 0, world = 1;
  int l1 = 0, l2 = 0;
  uint64_t N1 = 0, N2 = 0, C = 0, T = 0, per = 0;   // per = N / G elements per rank
  uint4 *A = nullptr, *B = nullptr, *I = nullptr, *O = nullptr;
  uint4* peerB[NTT_MAX_PEERS] = {nullptr};
  bool peers_open = false;
  bool use_comm = false;   // handles / barriers go through the DriverClient's communicator
  uint4* tab_mem = nullptr;
  NttTables tab{};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms[2] = {0, 0};
};

static int ilog2u(uint64_t v) { int l = 0; while ((1ull << l) < v) l++; return l; }

extern "C" int32_t bz_ntt_dist_new(bz_dclient* dc, int32_t field, int32_t log_size, int32_t inverse, int32_t rank,
                                   int32_t world, bz_ntt_dist** out) {
  if (!out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  *out = nullptr;
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  if (field < 0 || field > 2) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown field %d", field);
  if (world < 1 || world > NTT_MAX_PEERS || (world & (world - 1)) || rank < 0 || rank >= world)
    return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad rank/world %d/%d (world must be a power of two <= %d)", rank, world, NTT_MAX_PEERS);
  if (log_size > ntt_two_adicity(field) || log_size > 30) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "log size %d unsupported", log_size);
  int lw = ilog2u(world);
  // N = N1 x N2 = 2^l1 x 2^l2: the split with the fewest global passes (radix <= 2^9 each), e.g. 27 = 9 + 18 -> 1 + 2
  // passes instead of 13 + 14 -> 2 + 2; ties go to the more balanced split
  int l1 = log_size / 2, l2 = log_size - l1;
  {
    auto passes = [](int l) { return l == 0 ? 0 : (l + 8) / 9; };
    int best = passes(l1) + passes(l2);
    for (int a = std::max(lw, 1); a <= log_size - std::max(lw, 1); a++) {   // both factors >= max(world, 2)
      int b = log_size - a, np = passes(a) + passes(b);
      if (np < best || (np == best && std::abs(a - b) < std::abs(l1 - l2))) { best = np; l1 = a; l2 = b; }
    }
  }
  // the last column pass needs radix >= world, and each rank at least one row / column
  if (l1 < lw || l2 < lw || l1 < 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "2^%d is too small for %d ranks", log_size, world);
  bz_ntt_dist* t = new bz_ntt_dist();
  t->dc = dc; t->field = field; t->log_n = log_size; t->inverse = inverse ? 1 : 0; t->rank = rank; t->world = world;
  t->l1 = l1; t->l2 = l2;
  t->N1 = 1ull << l1; t->N2 = 1ull << l2; t->C = t->N2 / world; t->T = t->N1 / world;
  t->per = (1ull << log_size) / world;
  size_t bytes = t->per * 32;
  cudaStream_t st = dc_stream(dc);
  for (uint4** b : {&t->A, &t->B, &t->I, &t->O}) {
    if (cudaMalloc((void**)b, bytes) != cudaSuccess) { bz_ntt_dist_free(t); return bz_fail(BZ_ERR_WRITE, "device allocation of %zu bytes failed", bytes); }
    cudaMemsetAsync(*b, 0, bytes, st);
  }
  const int lo_bits = 14;
  uint32_t nlo = 1u << std::min(log_size, lo_bits);
  uint32_t nhi = log_size > lo_bits ? 1u << (log_size - lo_bits) : 1u;
  if (cudaMalloc((void**)&t->tab_mem, ((size_t)nlo + nhi + 1) * 32) != cudaSuccess) { bz_ntt_dist_free(t); return bz_fail(BZ_ERR_WRITE, "table allocation failed"); }
  t->tab.lo = t->tab_mem;
  t->tab.hi = t->tab_mem + 2 * (size_t)nlo;
  t->tab.ninv = t->tab_mem + 2 * ((size_t)nlo + nhi);
  t->tab.lo_bits = lo_bits;
  t->tab.log_root = log_size;
  ntt_gen_tables(field, t->tab, log_size, t->inverse, st);
  for (auto& e : t->ev) cudaEventCreate(&e);
  if (world == 1) { t->peerB[0] = t->B; t->peers_open = true; }
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(st));
  if (world > 1 && dc->comm && dc->world == world && dc->rank == rank) {
    // a ranked DriverClient: the IPC handles of the exchange buffers travel through its communicator
    uint8_t mine[64];
    int32_t rc2 = bz_ntt_dist_ipc_handle(t, mine);
    uint8_t* d = nullptr;
    std::vector<uint8_t> all((size_t)world * 64);
    if (!rc2 && cudaMalloc((void**)&d, (size_t)(world + 1) * 64) != cudaSuccess) rc2 = bz_fail(BZ_ERR_WRITE, "device allocation failed");
    if (!rc2) {
      cudaMemcpyAsync(d + (size_t)world * 64, mine, 64, cudaMemcpyHostToDevice, st);
      rc2 = comm_allgather(dc, d + (size_t)world * 64, d, 64, st);
    }
    if (!rc2) {
      cudaMemcpyAsync(all.data(), d, (size_t)world * 64, cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) rc2 = bz_fail(BZ_ERR_UNKNOWN, "handle exchange failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (d) cudaFree(d);
    if (!rc2) rc2 = bz_ntt_dist_open_peers(t, all.data());
    if (rc2) { bz_ntt_dist_free(t); return rc2; }
    t->use_comm = true;
  }
  *out = t;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_free(bz_ntt_dist* t) {
  if (!t) return BZ_OK;
  cudaSetDevice(dc_device(t->dc));
  cudaStreamSynchronize(dc_stream(t->dc));
  if (t->world > 1 && t->peers_open)
    for (int h = 0; h < t->world; h++) if (h != t->rank && t->peerB[h]) cudaIpcCloseMemHandle(t->peerB[h]);
  for (uint4* b : {t->A, t->B, t->I, t->O, t->tab_mem}) if (b) cudaFree(b);
  for (auto& e : t->ev) if (e) cudaEventDestroy(e);
  delete t;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_ipc_handle(bz_ntt_dist* t, uint8_t out[64]) {
  if (!t || !out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaIpcGetMemHandle(&h, t->B));
  memcpy(out, &h, 64);
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_open_peers(bz_ntt_dist* t, const uint8_t* handles /* world x 64 bytes, rank order */) {
  if (!t || !handles) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  for (int h = 0; h < t->world; h++) {
    if (h == t->rank) { t->peerB[h] = t->B; continue; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handles + (size_t)h * 64, 64);
    void* p = nullptr;
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    t->peerB[h] = (uint4*)p;
  }
  t->peers_open = true;
  return BZ_OK;
}

// host natural-order vector -> this rank's column slab A (strided copy)
extern "C" int32_t bz_ntt_dist_set_input(bz_ntt_dist* t, const uint8_t* full_input, size_t len) {
  if (!t || !full_input) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (len != ((size_t)32 << t->log_n)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "input must be the full 2^%d x 32 B vector", t->log_n);
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpy2DAsync(t->A, t->C * 32, full_input + (size_t)t->rank * t->C * 32, t->N2 * 32, t->C * 32, t->N1,
                                           cudaMemcpyHostToDevice, st));
  CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(st));
  return BZ_OK;
}

// this rank's output block O[k2][t] -> natural-order host vector positions (rank T + t) + N1 k2
extern "C" int32_t bz_ntt_dist_get_output(bz_ntt_dist* t, uint8_t* full_output, size_t len) {
  if (!t || !full_output) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (len != ((size_t)32 << t->log_n)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "output must be the full 2^%d x 32 B vector", t->log_n);
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  CUDA_TRY(BZ_ERR_READ, cudaMemcpy2DAsync(full_output + (size_t)t->rank * t->T * 32, t->N1 * 32, t->O, t->T * 32, t->T * 32, t->N2,
                                          cudaMemcpyDeviceToHost, st));
  CUDA_TRY(BZ_ERR_READ, cudaStreamSynchronize(st));
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_buffers(bz_ntt_dist* t, uint64_t* slab_in_dev, uint64_t* block_out_dev, uint64_t* elems_per_rank) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (slab_in_dev) *slab_in_dev = (uint64_t)(uintptr_t)t->A;
  if (block_out_dev) *block_out_dev = (uint64_t)(uintptr_t)t->O;
  if (elems_per_rank) *elems_per_rank = t->per;
  return BZ_OK;
}

static void fill_logs(NttPassParams& P) {
  P.lq0 = ilog2u(P.Q0);
  P.lq1 = ilog2u(P.Q1);
  P.lpeer_rows = P.peer_rows ? ilog2u(P.peer_rows) : 0;
}

extern "C" int32_t bz_ntt_dist_step1(bz_ntt_dist* t) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (!t->peers_open) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "open_peers has not been called");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  const uint64_t N = 1ull << t->log_n, N1 = t->N1, N2 = t->N2, C = t->C;
  std::vector<int> rad = plan_radices(t->l1);
  if (rad.empty()) rad.push_back(0);
  cudaEventRecord(t->ev[0], st);
  const uint4* cur = t->A;
  uint64_t Ns = 1;
  for (size_t p = 0; p < rad.size(); p++) {
    int lr = rad[p];
    uint64_t R = 1ull << lr;
    bool last = p + 1 == rad.size();
    NttPassParams P;
    memset(&P, 0, sizeof(P));
    P.tab = t->tab;
    P.lr = lr;
    P.in = cur;
    P.Q = (N1 / R) * C;
    P.Q0 = C;
    P.in_s0 = 1; P.out_s0 = 1;
    P.tw_sel = -1;
    P.otw_rsel = -1;
    if (Ns == 1) {
      P.Q1 = 1;
      P.in_s2 = C; P.in_sr = (N1 / R) * C;
      P.out_s2 = R * C; P.out_sr = C;
    } else {
      P.Q1 = Ns;
      P.in_s1 = C; P.in_s2 = Ns * C; P.in_sr = (N1 / R) * C;
      P.out_s1 = C; P.out_s2 = Ns * R * C; P.out_sr = Ns * C;
      P.tw_sel = 1;
      P.tw_scale = N / (Ns * R);
    }
    if (last) {
      if (R < (uint64_t)t->world) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "last column radix %llu < world %d", (unsigned long long)R, t->world);
      P.otw_rsel = 1; P.otw_ra = 1; P.otw_rb = Ns; P.otw_base = (uint64_t)t->rank * C; P.otw_scale = 1;
      P.peer_rows = (uint32_t)(R / t->world);
      P.out_s0 = 1; P.out_s1 = N2; P.out_s2 = 0; P.out_sr = Ns * N2;
      for (int h = 0; h < t->world; h++) P.peer_out[h] = t->peerB[h] + 2 * ((uint64_t)t->rank * C);
      P.out = nullptr;
    } else {
      P.out = (cur == t->A) ? t->I : const_cast<uint4*>(t->A);
    }
    fill_logs(P);
    cudaError_t e = ntt_launch_pass(t->field, P, st);
    if (e != cudaSuccess) return bz_fail(BZ_ERR_UNKNOWN, "NTT column pass failed: %s", cudaGetErrorString(e));
    cur = P.out;
    Ns *= R;
  }
  cudaEventRecord(t->ev[1], st);
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_sync(bz_ntt_dist* t) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(dc_stream(t->dc)));
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_step3(bz_ntt_dist* t) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  const uint64_t N = 1ull << t->log_n, N2 = t->N2, T = t->T;
  std::vector<int> rad = plan_radices(t->l2);
  cudaEventRecord(t->ev[2], st);
  const uint4* cur = t->B;
  uint64_t Ns = 1;
  for (size_t p = 0; p < rad.size(); p++) {
    int lr = rad[p];
    uint64_t R = 1ull << lr;
    bool last = p + 1 == rad.size();
    NttPassParams P;
    memset(&P, 0, sizeof(P));
    P.tab = t->tab;
    P.lr = lr;
    P.in = cur;
    P.tw_sel = -1;
    P.otw_rsel = -1;
    P.Q = (N2 / R) * T;
    if (p == 0 && !last) {          // rows in (contiguous), transposed out: I[pos][t]
      P.Q0 = N2 / R; P.Q1 = T;
      P.in_s0 = 1; P.in_s1 = N2; P.in_sr = N2 / R;
      P.out_s0 = R * T; P.out_s1 = 1; P.out_sr = T;
    } else if (p == 0 && last) {    // single pass: rows in, O[k2][t] out
      P.Q0 = T; P.Q1 = 1;
      P.in_s0 = N2; P.in_sr = 1;
      P.out_s0 = 1; P.out_sr = T;
    } else {                        // lanes along t on I[pos][t]
      P.Q0 = T; P.Q1 = Ns;
      P.in_s0 = 1; P.in_s1 = T; P.in_s2 = Ns * T; P.in_sr = (N2 / R) * T;
      P.out_s0 = 1; P.out_s1 = T; P.out_s2 = Ns * R * T; P.out_sr = Ns * T;
      P.tw_sel = 1;
      P.tw_scale = N / (Ns * R);
    }
    P.out = last ? t->O : ((cur == t->I) ? t->A : t->I);
    P.scale_ninv = (t->inverse && last) ? 1 : 0;
    fill_logs(P);
    cudaError_t e = ntt_launch_pass(t->field, P, st);
    if (e != cudaSuccess) return bz_fail(BZ_ERR_UNKNOWN, "NTT row pass failed: %s", cudaGetErrorString(e));
    cur = P.out;
    Ns *= R;
  }
  cudaEventRecord(t->ev[3], st);
  return BZ_OK;
}

// The whole transform, stream-ordered, no host synchronisation: the two cross-rank barriers (every exchange buffer is
// free again / every rank has finished storing into every exchange buffer) are 4-byte NCCL all-reduces on the client's
// stream.  Needs a ranked DriverClient (bz_dclient_comm_init); otherwise drive step1 / sync / barrier / step3 yourself.
extern "C" int32_t bz_ntt_dist_run(bz_ntt_dist* t) {
  if (!t) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  if (t->world > 1) {
    if (!t->use_comm) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bz_ntt_dist_run needs a DriverClient with a communicator");
    rc = comm_barrier(t->dc, st);
    if (rc) return rc;
  }
  rc = bz_ntt_dist_step1(t);
  if (rc) return rc;
  if (t->world > 1) { rc = comm_barrier(t->dc, st); if (rc) return rc; }
  return bz_ntt_dist_step3(t);
}

extern "C" int32_t bz_ntt_dist_plan(bz_ntt_dist* t, int32_t out[4]) {
  if (!t || !out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  out[0] = t->l1;
  out[1] = t->l2;
  out[2] = (int32_t)std::max<size_t>(1, plan_radices(t->l1).size());
  out[3] = (int32_t)plan_radices(t->l2).size();
  return BZ_OK;
}

extern "C" int32_t bz_ntt_dist_times(bz_ntt_dist* t, float ms[2]) {
  if (!t || !ms) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(dc_stream(t->dc)));
  cudaEventElapsedTime(&ms[0], t->ev[0], t->ev[1]);
  cudaEventElapsedTime(&ms[1], t->ev[2], t->ev[3]);
  return BZ_OK;
}
