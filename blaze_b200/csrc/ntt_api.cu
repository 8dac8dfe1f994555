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
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/blaze_b200.h"
#include "api_common.h"
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
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t done = nullptr;
  bool launched = false;
  float last_ms = 0;
  std::mutex mu;
};

static std::vector<int> plan_radices(int log_n) {
  std::vector<int> r;
  if (log_n == 0) return r;
  int npass = (log_n + 8) / 9, base = log_n / npass, extra = log_n % npass;
  for (int i = 0; i < npass; i++) r.push_back(i < extra ? base + 1 : base);
  return r;
}

static int32_t ntt_alloc_slot(bz_ntt* t, int s) {
  for (int k = 0; k < 2; k++)
    if (!t->buf[s][k]) {
      CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&t->buf[s][k], std::max<uint64_t>(t->n, 1) * 32));
      CUDA_TRY(BZ_ERR_WRITE, cudaMemsetAsync(t->buf[s][k], 0, std::max<uint64_t>(t->n, 1) * 32, dc_stream(t->dc)));
    }
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
  for (auto& s : t->buf) for (auto& b : s) if (b) cudaFree(b);
  if (t->tab_mem) cudaFree(t->tab_mem);
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
  t->tables_ready = true;
  return BZ_OK;
}

extern "C" int32_t bz_ntt_set_data(bz_ntt* t, size_t buf_host, const uint8_t* data, size_t len) {
  if (!t || !data) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (buf_host > 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "buf_host must be 0 or 1");
  if (len != t->n * 32) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "data length %zu != %llu*32", len, (unsigned long long)t->n);
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(t->mu);
  rc = ntt_alloc_slot(t, (int)buf_host);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(t->buf[buf_host][t->cur[buf_host]], data, len, cudaMemcpyHostToDevice, st));
  CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(st));   // caller may drop `data` (move-in semantics)
  return BZ_OK;
}

// enqueue all passes of one transform of slot s
static int32_t ntt_enqueue(bz_ntt* t, int s) {
  cudaStream_t st = dc_stream(t->dc);
  uint64_t L = t->n, Ns = 1;
  cudaEventRecord(t->ev[0], st);
  for (size_t p = 0; p < t->radices.size(); p++) {
    int lr = t->radices[p];
    uint64_t R = 1ull << lr;
    NttPassParams P;
    memset(&P, 0, sizeof(P));
    P.in = t->buf[s][t->cur[s]];
    P.out = t->buf[s][t->cur[s] ^ 1];
    P.lr = lr;
    P.Q = L / R;
    P.in_sr = L / R;
    P.otw_sel = -1;
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
    cudaError_t e = ntt_launch_pass(t->field, P, st);
    if (e != cudaSuccess) return bz_fail(BZ_ERR_UNKNOWN, "NTT pass launch failed: %s", cudaGetErrorString(e));
    t->cur[s] ^= 1;
    Ns *= R;
  }
  cudaEventRecord(t->ev[1], st);
  cudaEventRecord(t->done, st);
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
  std::lock_guard<std::mutex> lk(t->mu);
  if (!t->launched) return bz_fail(BZ_ERR_NO_RESULT, "no transform in flight");
  CUDA_TRY(BZ_ERR_READ, cudaEventSynchronize(t->done));
  cudaEventElapsedTime(&t->last_ms, t->ev[0], t->ev[1]);
  return BZ_OK;
}

extern "C" int32_t bz_ntt_result(bz_ntt* t, size_t buf_num, uint8_t* out, size_t out_len) {
  if (!t || !out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (buf_num > 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "buf_num must be 0 or 1");
  if (out_len < t->n * 32) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "result buffer too small");
  int32_t rc = dc_select(t->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(t->mu);
  rc = ntt_alloc_slot(t, (int)buf_num);
  if (rc) return rc;
  cudaStream_t st = dc_stream(t->dc);
  CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(out, t->buf[buf_num][t->cur[buf_num]], t->n * 32, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(BZ_ERR_READ, cudaStreamSynchronize(st));
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
  *dev_ptr = (uint64_t)(uintptr_t)t->buf[buf_num][t->cur[buf_num]];
  return BZ_OK;
}
