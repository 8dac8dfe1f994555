// Internal (non-ABI) view of the MSMClient handle.
//
// A bz_msm is either a LEAF -- one device, the whole pipeline of msm_client.cu -- or a GROUP over the member devices of
// a multi-device DriverClient (id "0,1,..": msm_api.cu shards points and scalars over one leaf per member and sums the
// members' partial results on the first device).  The extern "C" entry points of include/blaze_b200.h live in
// msm_api.cu and dispatch on `parts.empty()`.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <deque>
#include <mutex>
#include <vector>

#include "client_internal.h"
#include "msm_internal.h"

namespace bz {

static const int RESULT_SLOTS = 16;         // max tasks in flight per client
static const int RESULT_SLOT_BYTES = 272;   // 3*48 result bytes + error word, 16-byte aligned

struct MsmTaskResult {
  std::vector<uint8_t> bytes;   // filled from the pinned slot when the event has completed
  uint32_t label = 0;
  int slot = 0;                 // pinned result slot + timing-event set of this task
  cudaEvent_t done = nullptr;
  uint8_t* host_slot = nullptr;   // pinned (slot of bz_msm::pinned)
  int* host_err = nullptr;        // pinned
  bool collected = false;
  int32_t status = BZ_OK;
};

}  // namespace bz

struct bz_msm {
  bz_dclient* dc = nullptr;
  const bz::CurveOps* ops = nullptr;
  int curve = 0, mem_type = BZ_MEM_DMA, factor = 1;
  // "registers" of the reference core
  uint32_t nof_elements = 0;
  bool hbm_mode = false;
  uint64_t hbm_addr = 0, hbm_off = 0;
  uint32_t next_label = 0, last_label = 0;
  int forced_c = 0;
  int acc_mode = -1, acc_rounds = -1;   // bz_msm_set_accumulate_mode: -1 automatic
  // task state machine
  int pending_tasks = 0;     // start_process() calls not yet matched with data
  bool data_ready = false;   // set_data() arrived, not yet consumed by a task
  uint64_t data_M = 0;
  uint64_t launched = 0;     // tasks enqueued so far (the group watches this)
  std::deque<bz::MsmTaskResult> results;
  // device state
  bz::MsmPlan plan{};
  bool have_plan = false;
  bz::MsmWorkspace ws{};
  std::vector<void*> ws_allocs;
  void* table = nullptr;
  uint64_t table_cap = 0, table_n = 0, table_addr = ~0ull, table_epoch = 0, table_gen = 0;
  bool table_from_arena = false;
  // window-merged table (SURVEY 8(a) "HBM-resident precomputed points"): entry w*n + i = 2^(c w) * P_i.  Built on
  // the second MSM over an unchanged resident point set (precomp_mode 1), immediately (2) or never (0).
  void* wtable = nullptr;
  size_t wtable_bytes = 0;
  uint64_t wtable_n = 0, wtable_gen = 0;
  int wtable_c = 0, wtable_levels = 0;
  float wtable_build_ms = 0;
  int precomp_mode = 1;
  int raw_result = 0;            // bz_msm_set_raw_result
  bool precomp_failed = false;   // allocation failed for this point set: stay on the plain table
  uint64_t table_uses = 0;       // MSMs launched on the current arena table
  uint8_t* comb_dev = nullptr;   // scratch of the result combine (ranked / group / bz_msm_combine_results)
  size_t comb_cap = 0;
  void* ba2_scratch = nullptr;   // per-warp scratch of the batched-affine sweep (allocated on first use)
  size_t ba2_cap = 0;
  uint8_t* dma_points = nullptr;
  size_t dma_points_cap = 0;
  // DMA mode: the points travel on the copy stream BEHIND the scalars, so digits + sort of the task run while
  // the (larger) point copy is still in flight; the table is built on the work stream once they have landed
  bool table_pending = false;
  uint64_t table_pending_n = 0;
  cudaEvent_t ev_points_copied = nullptr, ev_points_consumed = nullptr;
  bool points_consumed_valid = false;
  // scalar ingest: two staging buffers filled on a dedicated copy stream, so the H2D of task k+1
  // overlaps the kernels of task k (the reference's task queue allows exactly that pipelining)
  uint32_t* scalars_dev[2] = {nullptr, nullptr};
  size_t scalars_cap[2] = {0, 0};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};     // copy stream: staging buffer b is filled
  cudaEvent_t ev_consumed[2] = {nullptr, nullptr};   // work stream: k_digits has read staging buffer b
  bool consumed_valid[2] = {false, false};
  int stage_next = 0, stage_cur = -1;
  const uint32_t* scalars_src = nullptr;   // where the pending task reads its scalars from
  // tail stream (see MsmWorkspace::tail): higher priority than the work stream; BZ_MSM_TAIL=0 disables it
  cudaStream_t tail = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_tail_done = nullptr;
  bool tail_busy = false;
  cudaStream_t result_stream = nullptr;   // stream on which the last launched task leaves ws.result (work or tail)
  uint8_t* pinned = nullptr;   // RESULT_SLOTS result slots
  // per result slot: start, sorted, accumulate begin, accumulate end, done -- a task's phase times are read from ITS
  // events, so two tasks in flight do not clobber each other's timers
  cudaEvent_t tev[bz::RESULT_SLOTS][5] = {};
  float last_ms[4] = {0, 0, 0, 0};
  uint64_t tasks_done = 0;
  // group over the members of a multi-device DriverClient (empty for a leaf)
  std::vector<bz_msm*> parts;
  std::vector<uint32_t> part_n;        // elements of the current shard of each part (0 = idle)
  uint32_t part_per = 0;               // elements per shard = ceil(nof_elements / parts)
  std::vector<uint64_t> part_seen;     // parts[g]->launched already combined
  cudaEvent_t ev_part[16] = {};        // per member: its partial record has landed on the first device
  std::mutex mu;
};

namespace bz {

// leaf implementations (msm_client.cu); the extern "C" names of include/blaze_b200.h dispatch to these
int32_t leaf_new(bz_dclient* dc, int32_t curve, int32_t mem_type, int32_t is_precompute, bz_msm** out);
int32_t leaf_free(bz_msm* m);
int32_t leaf_initialize(bz_msm* m, uint32_t nof_elements, int32_t has_hbm_addr, uint64_t hbm_addr, uint64_t hbm_offset);
int32_t leaf_start_process(bz_msm* m);
// sync_host = false: the H2D copies are only enqueued; the caller waits for them with leaf_sync_copies before returning
int32_t leaf_set_data(bz_msm* m, const uint8_t* points, size_t points_len, const uint8_t* scalars_host,
                      uint64_t scalars_dev_ptr, size_t scalars_len, uint32_t nof_elements, int32_t has_hbm_addr,
                      uint64_t hbm_addr, uint64_t hbm_offset, bool sync_host);
int32_t leaf_sync_copies(bz_msm* m);
int32_t leaf_wait_result(bz_msm* m);
int32_t leaf_result(bz_msm* m, uint8_t* out, size_t out_len, uint32_t* result_label);
int32_t leaf_load_data_to_hbm(bz_msm* m, const uint8_t* points, size_t len, uint64_t addr, uint64_t offset);
int32_t leaf_generate_chain_points(bz_msm* m, const uint8_t* p0q, size_t p0q_len, uint64_t first, uint64_t n, uint64_t addr,
                                   uint64_t offset);
int32_t leaf_phase_times(bz_msm* m, float ms[4]);

int32_t leaf_combine_results(bz_msm* m, const uint8_t* records, int32_t n, uint8_t* out, size_t out_len);
// make sure m->comb_dev holds at least `bytes`
int32_t leaf_comb_reserve(bz_msm* m, size_t bytes);

}  // namespace bz
