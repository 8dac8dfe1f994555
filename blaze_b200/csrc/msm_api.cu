// MSMClient half of the C ABI (include/blaze_b200.h): dispatch to the one-device pipeline (msm_client.cu) or, for a
// multi-device DriverClient (id "0,1,.."), to one such pipeline per member device.
//
// Multi-device MSM (point-sharded): element i of the MSM lives on member i / per, per =
// ceil(nof_elements / members); `load_data_to_hbm` and `set_data` split the caller's buffers accordingly (the H2D
// copies of all members are in flight together), every member runs the whole single-device pipeline on its shard and
// leaves a projective partial record in ITS HBM; the records are peer-copied to the first device on the members' own
// streams and summed there by k_combine_results -- the host never sees the partial results.  To the caller this is
// the unchanged MSMClient of /root/reference/src/ingo_msm/msm_api.rs: one result record and one label per task.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/blaze_b200.h"
#include "api_common.h"
#include "client_internal.h"
#include "msm_client.h"

using namespace bz;

#define fail bz_fail
void bz_set_last_error(const std::string& s);

static inline bool is_group(const bz_msm* m) { return !m->parts.empty(); }

// run fn(g) for every member with a non-empty shard; members 1.. on their own host threads when `parallel`
template <class F>
static int32_t for_parts(bz_msm* m, bool parallel, F fn) {
  const int G = (int)m->parts.size();
  std::vector<int32_t> rc(G, BZ_OK);
  std::vector<std::string> msg(G);
  auto run = [&](int g) {
    if (m->part_n.size() == (size_t)G && m->part_n[g] == 0) return;
    rc[g] = fn(g);
    if (rc[g]) msg[g] = bz_last_error();
  };
  if (parallel && G > 1) {
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(run, g);
    run(0);
    for (auto& t : th) t.join();
  } else {
    for (int g = 0; g < G; g++) run(g);
  }
  for (int g = 0; g < G; g++)
    if (rc[g]) { bz_set_last_error("device member " + std::to_string(g) + ": " + msg[g]); return rc[g]; }
  return BZ_OK;
}

static void group_shard(bz_msm* m, uint64_t n) {
  const uint64_t G = m->parts.size();
  const uint64_t per = (n + G - 1) / G;
  m->part_per = (uint32_t)per;
  m->part_n.assign(G, 0);
  for (uint64_t g = 0; g < G; g++) {
    const uint64_t lo = g * per;
    m->part_n[g] = lo >= n ? 0 : (uint32_t)std::min<uint64_t>(per, n - lo);
  }
}

// ------------------------------------------------------------------------------------ construction
extern "C" int32_t bz_msm_new(bz_dclient* dc, int32_t curve, int32_t mem_type, int32_t is_precompute, bz_msm** out) {
  if (!out) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  *out = nullptr;
  if (!dc) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient");
  if (dc_members(dc) == 1) return leaf_new(dc, curve, mem_type, is_precompute, out);
  // group handle: a leaf-shaped object on the first device (result slots, combine scratch) + one leaf per member
  bz_msm* g = nullptr;
  int32_t rc = leaf_new(dc, curve, mem_type, is_precompute, &g);
  if (rc) return rc;
  const int G = dc_members(dc);
  for (int k = 0; k < G; k++) {
    bz_msm* p = nullptr;
    rc = leaf_new(dc_member(dc, k), curve, mem_type, is_precompute, &p);
    if (rc) { bz_msm_free(g); return rc; }
    p->raw_result = 1;   // members leave projective partial records; the combine normalises once
    g->parts.push_back(p);
    cudaSetDevice(dc_member(dc, k)->device);
    cudaEventCreateWithFlags(&g->ev_part[k], cudaEventDisableTiming);
  }
  g->part_seen.assign(G, 0);
  cudaSetDevice(dc->device);
  const size_t rs = 3 * (size_t)g->ops->fq_bytes;
  rc = leaf_comb_reserve(g, (size_t)RESULT_SLOTS * (G + 1) * rs);
  if (rc) { bz_msm_free(g); return rc; }
  *out = g;
  return BZ_OK;
}

extern "C" int32_t bz_msm_free(bz_msm* m) {
  if (!m) return BZ_OK;
  for (bz_msm* p : m->parts) leaf_free(p);
  m->parts.clear();
  return leaf_free(m);
}

extern "C" int32_t bz_msm_sizes(bz_msm* m, uint32_t* scalar_size, uint32_t* point_size, uint32_t* result_point_size, uint32_t* precompute_factor) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (scalar_size) *scalar_size = 32;
  if (point_size) *point_size = 2 * m->ops->fq_bytes;
  if (result_point_size) *result_point_size = 3 * m->ops->fq_bytes;
  if (precompute_factor) *precompute_factor = m->factor;
  return BZ_OK;
}

static uint32_t image_parameters(bz_msm* m) {
  // The reference decodes the word as reverse_bits().to_be_bytes() unpacked msb0 (msm_api.rs:333-354): after the
  // reversal, LSB-first: bits 0..3 placeholder, 4..7 #segments, 8..15 bucket addr width, 16..19 #ec adders,
  // 20..27 curve, 28..31 is_stub.  We synthesise: segments = member devices, addr width = current c-1 (or 0),
  // ec adders = 0xF (saturated: 148 SMs do not fit 4 bits), curve code, is_stub = 0.
  const bz_msm* p = is_group(m) ? m->parts[0] : m;
  uint32_t c = p->have_plan ? (uint32_t)(p->plan.c - 1) : 0;
  uint32_t seg = is_group(m) ? (uint32_t)std::min<size_t>(m->parts.size(), 15) : 1;
  return (seg << 4) | ((c & 0xff) << 8) | (0xFu << 16) | (((uint32_t)m->curve & 0xff) << 20);
}

extern "C" int32_t bz_msm_loaded_binary_parameters(bz_msm* m, uint32_t out[2]) {
  if (!m || !out) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  out[0] = 0xB2000000u | (uint32_t)m->curve;   // image id
  out[1] = image_parameters(m);
  return BZ_OK;
}

// ------------------------------------------------------------------------------------ the 7-method trait
extern "C" int32_t bz_msm_initialize(bz_msm* m, uint32_t nof_elements, int32_t has_hbm_addr, uint64_t hbm_addr, uint64_t hbm_offset) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  int32_t rc = leaf_initialize(m, nof_elements, has_hbm_addr, hbm_addr, hbm_offset);
  if (rc || !is_group(m)) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  group_shard(m, nof_elements);
  return for_parts(m, false, [&](int g) { return leaf_initialize(m->parts[g], m->part_n[g], has_hbm_addr, hbm_addr, hbm_offset); });
}

// after a call that may have made the members launch a task: enqueue the device-side sum of their partial records
static int32_t group_after_launch(bz_msm* m) {
  const int G = (int)m->parts.size();
  int launched = 0, active = 0;
  for (int g = 0; g < G; g++) {
    if (m->part_n[g] == 0) continue;
    active++;
    if (m->parts[g]->launched > m->part_seen[g]) launched++;
  }
  if (launched == 0) return BZ_OK;
  if (launched != active) return fail(BZ_ERR_UNKNOWN, "internal: %d of %d member devices launched", launched, active);
  if ((int)m->results.size() >= RESULT_SLOTS) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "more than %d results pending; pop them with result()", RESULT_SLOTS);
  const size_t rs = 3 * (size_t)m->ops->fq_bytes;
  MsmTaskResult r;
  cudaSetDevice(m->dc->device);
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaEventCreateWithFlags(&r.done, cudaEventBlockingSync | cudaEventDisableTiming));
  r.label = m->next_label++;
  m->last_label = r.label;
  r.slot = (int)(r.label % RESULT_SLOTS);
  r.host_slot = m->pinned + (size_t)r.slot * RESULT_SLOT_BYTES;
  r.host_err = reinterpret_cast<int*>(r.host_slot + 256);
  *r.host_err = 0;
  uint8_t* comb = m->comb_dev + (size_t)r.slot * (G + 1) * rs;
  int a = 0;
  for (int g = 0; g < G; g++) {
    if (m->part_n[g] == 0) continue;
    bz_msm* p = m->parts[g];
    cudaSetDevice(p->dc->device);
    // behind the member's pipeline on the member's stream: its record travels device-to-device over NVLink
    cudaStream_t pst = p->result_stream ? p->result_stream : p->dc->stream;   // the member's pipeline may end on its tail stream
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaMemcpyPeerAsync(comb + (size_t)a * rs, m->dc->device, p->ws.result, p->dc->device, rs, pst));
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaEventRecord(m->ev_part[g], pst));
    m->part_seen[g] = p->launched;
    a++;
  }
  cudaSetDevice(m->dc->device);
  cudaStream_t st = m->dc->stream;
  for (int g = 0; g < G; g++)
    if (m->part_n[g]) CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamWaitEvent(st, m->ev_part[g], 0));
  m->ops->combine_results(comb, a, comb + (size_t)G * rs, m->raw_result, st);
  g_kernel_launches += 1;
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
  CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(r.host_slot, comb + (size_t)G * rs, rs, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaEventRecord(r.done, st));
  m->results.push_back(r);
  if (m->pending_tasks > 0) m->pending_tasks--;
  m->launched++;
  return BZ_OK;
}

extern "C" int32_t bz_msm_start_process(bz_msm* m) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (!is_group(m)) return leaf_start_process(m);
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->part_n.empty()) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: initialize() first (the shard sizes follow nof_elements)");
  if ((int)m->results.size() >= RESULT_SLOTS) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "more than %d results pending; pop them with result()", RESULT_SLOTS);
  m->pending_tasks++;
  int32_t rc = for_parts(m, false, [&](int g) { return leaf_start_process(m->parts[g]); });
  if (rc) return rc;
  return group_after_launch(m);
}

extern "C" int32_t bz_msm_set_data(bz_msm* m, const uint8_t* points, size_t points_len, const uint8_t* scalars, size_t scalars_len,
                                   uint32_t nof_elements, int32_t has_hbm_addr, uint64_t hbm_addr, uint64_t hbm_offset) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (!scalars) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null scalars");
  if (!is_group(m))
    return leaf_set_data(m, points, points_len, scalars, 0, scalars_len, nof_elements, has_hbm_addr, hbm_addr, hbm_offset, true);
  std::lock_guard<std::mutex> lk(m->mu);
  const uint64_t n = nof_elements;
  const uint64_t ps = 2ull * m->ops->fq_bytes * m->factor;
  if (n == 0) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "nof_elements must be > 0");
  if (scalars_len != n * 32) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "scalars length %zu != %llu*32", scalars_len, (unsigned long long)n);
  if (points && points_len != n * ps) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "points length %zu != %llu*%llu", points_len, (unsigned long long)n, (unsigned long long)ps);
  if (!points && !has_hbm_addr) return BZ_OK;   // (None, None): nothing happens (msm_api.rs:163-216)
  if (m->part_n.empty() || m->nof_elements != nof_elements) {
    // MSMParams travels twice (initialize and MSMInput, msm_api.rs:22-32); the copy in MSMInput decides
    group_shard(m, n);
    m->nof_elements = nof_elements;
    for (size_t g = 0; g < m->parts.size(); g++) {
      std::lock_guard<std::mutex> lp(m->parts[g]->mu);
      m->parts[g]->nof_elements = m->part_n[g];
    }
  }
  const uint64_t per = m->part_per;
  // every member's copies are enqueued (pinned host memory: truly asynchronous; pageable: one host thread per member)
  int32_t rc = for_parts(m, true, [&](int g) {
    const uint64_t lo = (uint64_t)g * per, cnt = m->part_n[g];
    return leaf_set_data(m->parts[g], points ? points + lo * ps : nullptr, points ? cnt * ps : 0, scalars + lo * 32, 0, cnt * 32,
                         (uint32_t)cnt, has_hbm_addr, hbm_addr, hbm_offset, false);
  });
  int32_t rc2 = for_parts(m, false, [&](int g) { return leaf_sync_copies(m->parts[g]); });   // move-in semantics
  if (rc) return rc;
  if (rc2) return rc2;
  return group_after_launch(m);
}

extern "C" int32_t bz_msm_set_scalars_device(bz_msm* m, uint64_t scalars_dev_ptr, uint32_t nof_elements, int32_t has_hbm_addr,
                                             uint64_t hbm_addr, uint64_t hbm_offset) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (is_group(m)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "device-resident scalars are per device: not available on a multi-device client (use set_data)");
  if (!scalars_dev_ptr || (scalars_dev_ptr & 15)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "device scalar pointer must be 16-byte aligned");
  if (!has_hbm_addr) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "device scalars need HBM-resident points");
  return leaf_set_data(m, nullptr, 0, nullptr, scalars_dev_ptr, (size_t)nof_elements * 32, nof_elements, has_hbm_addr, hbm_addr, hbm_offset, true);
}

// group: wait for the members' tasks (their device error flags) and for the combined record
static int32_t group_collect(bz_msm* m, MsmTaskResult& r) {
  if (r.collected) return r.status;
  int32_t rc = for_parts(m, false, [&](int g) { return leaf_wait_result(m->parts[g]); });
  cudaSetDevice(m->dc->device);
  CUDA_TRY(BZ_ERR_READ, cudaEventSynchronize(r.done));
  r.bytes.assign(r.host_slot, r.host_slot + 3 * m->ops->fq_bytes);
  r.status = rc;
  r.collected = true;
  float ms[4] = {0, 0, 0, 0};
  for (size_t g = 0; g < m->parts.size(); g++) {
    if (m->part_n[g] == 0) continue;
    float t[4];
    leaf_phase_times(m->parts[g], t);
    for (int i = 0; i < 4; i++) ms[i] = std::max(ms[i], t[i]);
  }
  memcpy(m->last_ms, ms, sizeof(ms));
  m->tasks_done++;
  return rc;
}

extern "C" int32_t bz_msm_wait_result(bz_msm* m) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (!is_group(m)) return leaf_wait_result(m);
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->results.empty()) return fail(BZ_ERR_NO_RESULT, "no task in flight (the reference would spin forever on RESULT_VALID)");
  return group_collect(m, m->results.front());
}

extern "C" int32_t bz_msm_result(bz_msm* m, uint8_t* out, size_t out_len, uint32_t* result_label) {
  if (!m || !out) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (!is_group(m)) return leaf_result(m, out, out_len, result_label);
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->results.empty()) return fail(BZ_ERR_NO_RESULT, "result queue is empty");
  const size_t need = 3 * (size_t)m->ops->fq_bytes;
  if (out_len < need) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "result buffer too small (%zu < %zu)", out_len, need);
  MsmTaskResult& r = m->results.front();
  int32_t rc = group_collect(m, r);
  std::string msg = rc ? bz_last_error() : "";
  for_parts(m, false, [&](int g) { leaf_result(m->parts[g], nullptr, 0, nullptr); return BZ_OK; });   // pop the members' queues
  if (rc == BZ_OK) {
    memcpy(out, r.bytes.data(), need);
    if (result_label) *result_label = r.label;
  } else {
    bz_set_last_error(msg);
  }
  if (r.done) cudaEventDestroy(r.done);
  m->results.pop_front();                      // POP_RESULT, msm_api.rs:265-269
  return rc;
}

// ------------------------------------------------------------------------------------ inherent methods
extern "C" int32_t bz_msm_task_label(bz_msm* m, uint32_t* label) {
  if (!m || !label) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  *label = m->last_label;
  return BZ_OK;
}
extern "C" int32_t bz_msm_nof_elements(bz_msm* m, uint32_t* n) {
  if (!m || !n) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  *n = m->nof_elements;
  return BZ_OK;
}
extern "C" int32_t bz_msm_is_msm_engine_ready(bz_msm* m, uint32_t* ready) {
  if (!m || !ready) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  *ready = 1;
  return BZ_OK;
}

// which elements of [first, first + cnt) live on member g, and where
static bool shard_overlap(const bz_msm* m, int g, uint64_t first, uint64_t cnt, uint64_t& lo, uint64_t& hi) {
  const uint64_t s0 = (uint64_t)g * m->part_per, s1 = s0 + m->part_n[g];
  lo = std::max(first, s0);
  hi = std::min(first + cnt, s1);
  return lo < hi;
}

extern "C" int32_t bz_msm_load_data_to_hbm(bz_msm* m, const uint8_t* points, size_t len, uint64_t addr, uint64_t offset) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (!is_group(m)) return leaf_load_data_to_hbm(m, points, len, addr, offset);
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->part_n.empty()) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: initialize() before load_data_to_hbm (the shards follow nof_elements)");
  const uint64_t rec = 2ull * m->ops->fq_bytes * m->factor;
  if (offset % rec || len % rec) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: offset and length must be whole %llu-byte base records", (unsigned long long)rec);
  const uint64_t first = offset / rec, cnt = len / rec;
  if (first + cnt > m->nof_elements) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: bases [%llu, %llu) exceed nof_elements %u", (unsigned long long)first, (unsigned long long)(first + cnt), m->nof_elements);
  m->hbm_mode = true;
  m->hbm_addr = addr;
  return for_parts(m, true, [&](int g) {
    uint64_t lo, hi;
    if (!shard_overlap(m, g, first, cnt, lo, hi)) return (int32_t)BZ_OK;
    return leaf_load_data_to_hbm(m->parts[g], points + (lo - first) * rec, (hi - lo) * rec, addr, (lo - (uint64_t)g * m->part_per) * rec);
  });
}

extern "C" int32_t bz_msm_get_data_from_hbm(bz_msm* m, uint8_t* out, size_t len, uint64_t addr, uint64_t offset) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (!is_group(m)) return bz_dclient_dma_read(m->dc, addr, offset, out, len);
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->part_n.empty()) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: initialize() first");
  const uint64_t rec = 2ull * m->ops->fq_bytes * m->factor;
  if (offset % rec || len % rec) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: offset and length must be whole %llu-byte base records", (unsigned long long)rec);
  const uint64_t first = offset / rec, cnt = len / rec;
  if (first + cnt > m->nof_elements) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "multi-device client: bases [%llu, %llu) exceed nof_elements %u", (unsigned long long)first, (unsigned long long)(first + cnt), m->nof_elements);
  return for_parts(m, false, [&](int g) {
    uint64_t lo, hi;
    if (!shard_overlap(m, g, first, cnt, lo, hi)) return (int32_t)BZ_OK;
    return bz_dclient_dma_read(m->parts[g]->dc, addr, (lo - (uint64_t)g * m->part_per) * rec, out + (lo - first) * rec, (hi - lo) * rec);
  });
}

// ------------------------------------------------------------------------------------ B200 additions
extern "C" int32_t bz_msm_phase_times(bz_msm* m, float ms[4]) {
  if (!m || !ms) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  return leaf_phase_times(m, ms);
}
extern "C" int32_t bz_msm_set_window_bits(bz_msm* m, int32_t c) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (c != 0 && (c < 4 || c > 26)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "window bits must be 0 or in [4, 26]");
  std::lock_guard<std::mutex> lk(m->mu);
  m->forced_c = c;
  for (bz_msm* p : m->parts) { std::lock_guard<std::mutex> lp(p->mu); p->forced_c = c; }
  return BZ_OK;
}
extern "C" int32_t bz_msm_set_precompute(bz_msm* m, int32_t mode) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (mode < 0 || mode > 2) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "precompute mode must be 0 (never), 1 (on reuse) or 2 (always)");
  std::lock_guard<std::mutex> lk(m->mu);
  m->precomp_mode = mode;
  m->precomp_failed = false;
  for (bz_msm* p : m->parts) { std::lock_guard<std::mutex> lp(p->mu); p->precomp_mode = mode; p->precomp_failed = false; }
  return BZ_OK;
}
extern "C" int32_t bz_msm_set_accumulate_mode(bz_msm* m, int32_t mode, int32_t rounds) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (mode != -1 && mode != 0 && mode != 2) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "accumulate mode: -1 automatic, 0 XYZZ sweep, 2 batched-affine sweep");
  if (rounds < -1 || rounds > 8) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "tree rounds must be -1 (automatic) or 0..8");
  std::lock_guard<std::mutex> lk(m->mu);
  m->acc_mode = mode;
  m->acc_rounds = rounds;
  for (bz_msm* p : m->parts) { std::lock_guard<std::mutex> lp(p->mu); p->acc_mode = mode; p->acc_rounds = rounds; }
  return BZ_OK;
}
extern "C" int32_t bz_msm_set_raw_result(bz_msm* m, int32_t raw) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  std::lock_guard<std::mutex> lk(m->mu);
  m->raw_result = raw ? 1 : 0;   // a group's members always stay raw; this is the flag of the combined record
  return BZ_OK;
}
extern "C" int32_t bz_msm_plan_info_ex(bz_msm* m, uint32_t out[8]) {
  if (!m || !out) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  memset(out, 0, 8 * sizeof(uint32_t));
  bz_msm* p = is_group(m) ? m->parts[0] : m;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!p->have_plan) return BZ_OK;
  out[0] = p->plan.c;
  out[1] = p->plan.Wd;        // digit windows = mixed adds per scalar
  out[2] = p->plan.nvalues;
  out[3] = p->plan.seg_len;
  out[4] = p->plan.W;         // bucket sets (1 when the windows are merged)
  out[5] = p->plan.merged;
  out[6] = (uint32_t)(p->wtable_bytes >> 20);   // MiB held by the window-merged table
  out[7] = p->plan.fb | (p->plan.rest << 8) | (p->plan.nlev << 16) | (p->plan.batch_affine << 24) |
           ((p->plan.batch_affine == 2 ? p->plan.ba_rounds : 0) << 28);
  return BZ_OK;
}
extern "C" int32_t bz_msm_plan_info(bz_msm* m, uint32_t out[4]) {
  uint32_t ex[8];
  int32_t rc = bz_msm_plan_info_ex(m, ex);
  if (rc) return rc;
  memcpy(out, ex, 4 * sizeof(uint32_t));
  return BZ_OK;
}
extern "C" int32_t bz_msm_table_build_ms(bz_msm* m, float* ms) {
  if (!m || !ms) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  float v = 0;
  if (is_group(m)) { for (bz_msm* p : m->parts) v = std::max(v, p->wtable_build_ms); }
  else v = m->wtable_build_ms;
  *ms = v;
  return BZ_OK;
}

extern "C" int32_t bz_msm_combine_results(bz_msm* m, const uint8_t* records, int32_t n, uint8_t* out, size_t out_len) {
  if (!m || !records || !out || n <= 0) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad argument");
  return leaf_combine_results(is_group(m) ? m->parts[0] : m, records, n, out, out_len);
}

extern "C" int32_t bz_msm_generate_chain_points(bz_msm* m, const uint8_t* p0q, size_t p0q_len, uint64_t first, uint64_t n,
                                                uint64_t addr, uint64_t offset) {
  if (!m || !p0q) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (!is_group(m)) return leaf_generate_chain_points(m, p0q, p0q_len, first, n, addr, offset);
  std::lock_guard<std::mutex> lk(m->mu);
  if (n == 0 || n > 0xffffffffull) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad element count");
  group_shard(m, n);   // the same split initialize(n) makes
  m->nof_elements = (uint32_t)n;
  return for_parts(m, true, [&](int g) {
    return leaf_generate_chain_points(m->parts[g], p0q, p0q_len, first + (uint64_t)g * m->part_per, m->part_n[g], addr, offset);
  });
}

// The reference's x8 "precomputed" base records (tests/msm/mod.rs:360-380: P, 2^32 P, .., 2^224 P per base) derived ON
// THE DEVICE from n factor-1 bases at src_addr into dst_addr of the card address space -- the test helper's host loop of
// seven scalar multiplications per base takes hours at the reference's 2^26 "precompute max" size.
extern "C" int32_t bz_msm_expand_precompute(bz_msm* m, uint64_t src_addr, uint64_t n, uint64_t dst_addr) {
  if (!m) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null MSMClient");
  if (is_group(m)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "expand_precompute works on one device's address space");
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  if (n == 0 || n > (1ull << 28)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad element count");
  const uint64_t ps = 2ull * m->ops->fq_bytes;
  if ((src_addr & 15) || (dst_addr & 15)) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "addresses must be 16-byte aligned");
  const uint64_t s1 = src_addr + n * ps, d1 = dst_addr + n * ps * 8;
  if (src_addr < d1 && dst_addr < s1) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "source and destination ranges overlap");
  bz_dclient* dc = m->dc;
  std::lock_guard<std::mutex> lk(dc->mu);
  if ((rc = arena_map(dc, src_addr, s1)) || (rc = arena_map(dc, dst_addr, d1))) return rc;
  void* t = nullptr;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc(&t, (size_t)(8 * n) * m->ops->affine_bytes));
  cudaStream_t st = dc->stream;
  m->ops->points_to_mont(dc->arena + src_addr, t, n, st);
  m->ops->build_wtable(t, n, 8, 32, st);
  m->ops->table_to_wire(t, n, 8, dc->arena + dst_addr, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(t);
  arena_note_write(dc, dst_addr, d1);
  if (e != cudaSuccess) return fail(BZ_ERR_UNKNOWN, "expand_precompute failed: %s", cudaGetErrorString(e));
  return BZ_OK;
}

extern "C" int32_t bz_msm_field_selftest(bz_msm* m, const uint8_t* a, const uint8_t* b, uint8_t* out, int32_t n, int32_t op) {
  if (!m || !a || !b || !out || n <= 0) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad argument");
  int32_t rc = dc_select(m->dc);
  if (rc) return rc;
  size_t bytes = (size_t)n * m->ops->fq_bytes;
  uint8_t* d = nullptr;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&d, 3 * bytes));
  cudaStream_t st = m->dc->stream;
  cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + bytes, b, bytes, cudaMemcpyHostToDevice, st);
  m->ops->field_selftest(d, d + bytes, d + 2 * bytes, n, op, st);
  cudaMemcpyAsync(out, d + 2 * bytes, bytes, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(BZ_ERR_UNKNOWN, "selftest failed: %s", cudaGetErrorString(e));
  return BZ_OK;
}

// ------------------------------------------------------------------------------------ get_api(): the register file
// The reference's get_api() / log_api_values() read every INGO_MSM_ADDR register (msm_api.rs:324-330,
// msm_hw_code.rs:6-55).  Here the same offsets are filled from the client state; the per-phase clock counters
// (msm_hw_code.rs:35-46) come from the CUDA events of the last completed task, converted to SM clocks:
//   PHASE1 total = ingest + digits + sort + bucket accumulation, PHASE1 busy EC adder = the accumulation kernel alone,
//   PHASE2 total = busy = partial merge + bucket reduction + window combine (+ the multi-GPU exchange), PHASE3 = 0.
extern "C" int32_t bz_msm_get_api(bz_msm* m, uint32_t* regs, size_t n_words) {
  if (!m || !regs) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if (n_words < 82) return fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "register image needs 82 words (0x000 .. 0x144)");
  std::lock_guard<std::mutex> lk(m->mu);
  memset(regs, 0, n_words * sizeof(uint32_t));
  auto W = [&](uint32_t off) -> uint32_t& { return regs[off / 4]; };
  W(0x0) = 0xB2000000u | (uint32_t)m->curve;
  W(0x4) = image_parameters(m);
  W(0x8) = 1;
  W(0xc) = m->last_label;
  W(0x10) = (uint32_t)m->hbm_addr;
  W(0x14) = (uint32_t)(m->hbm_addr >> 32);
  W(0x18) = m->hbm_mode ? 1 : 0;
  W(0x28) = m->nof_elements;
  uint32_t ready = 0;
  for (auto& r : m->results) {
    cudaError_t q = r.collected ? cudaSuccess : cudaEventQuery(r.done);
    if (q != cudaSuccess) { cudaGetLastError(); break; }
    ready++;
  }
  W(0x30) = ready ? 1 : 0;
  if (!m->results.empty()) {
    const MsmTaskResult& r = m->results.front();
    W(0x34) = r.label;
    if (ready) memcpy(&W(0x38), r.collected ? r.bytes.data() : r.host_slot, 3 * (size_t)m->ops->fq_bytes);
  }
  W(0xcc) = (uint32_t)std::max(0, m->pending_tasks) + (uint32_t)(m->results.size() - ready);
  W(0xd0) = ready;
  const bz_msm* p = is_group(m) ? m->parts[0] : m;
  W(0xd4) = (uint32_t)p->precomp_mode | (p->have_plan && p->plan.merged ? 0x10u : 0u);
  W(0xf0) = (uint32_t)m->tasks_done;
  W(0xf4) = (uint32_t)m->tasks_done;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, m->dc->device);
  auto clocks = [&](float ms) { return (uint64_t)((double)ms * (double)khz); };
  auto W64 = [&](uint32_t off, uint64_t v) { W(off) = (uint32_t)v; W(off + 4) = (uint32_t)(v >> 32); };
  W64(0xf8, clocks(m->last_ms[1] + m->last_ms[2]));
  W64(0x100, clocks(m->last_ms[2]));
  W64(0x108, clocks(m->last_ms[3]));
  W64(0x110, clocks(m->last_ms[3]));
  return BZ_OK;
}
