// Internal (non-ABI) view of the DriverClient handle, shared by the C-ABI translation units.
//
// A bz_dclient is the B200 stand-in for /root/reference/src/driver_client/dclient.rs:50-86: one "card" = one CUDA
// device, one work stream, and the card's flat HBM address space (dma_write / dma_read, dclient.rs:456-517).
//   * The address space is a VIRTUAL reservation of the whole window backed by physical chunks mapped on first touch
//     (CUDA virtual memory management): device pointers into it never move, growing costs no copy, and HBM that was never
//     written reads back as zeros.
//   * Writes are logged as (epoch, range) so that a client caching something derived from a range (the MSM's Montgomery
//     / window-merged tables) is invalidated only by writes that overlap it.
//   * id "0,1,2,3" makes a multi-device client: the handle itself drives the first device and owns one member client per
//     further device (MSMClient shards its points / scalars over the members, msm_group.cu).
//   * bz_dclient_comm_init turns the handle into rank r of a world of processes (one per GPU): the final exchange of a
//     sharded MSM and the barrier of the four-step NTT then run through NCCL on the client's stream.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/blaze_b200.h"

struct bz_dclient {
  int device = 0;
  int card_type = BZ_CARD_B200;
  cudaStream_t stream = nullptr;
  // card address space
  uint8_t* arena = nullptr;        // base of the virtual reservation (stable for the life of the client)
  uint64_t arena_va = 0;           // bytes usable (whole chunks)
  uint64_t arena_reserved = 0;     // bytes of the reservation itself
  size_t chunk = 0;                // mapping granule
  std::vector<uint8_t> mapped;     // per chunk: physical memory mapped
  std::vector<unsigned long long> handles;   // per chunk: CUmemGenericAllocationHandle
  // write log
  struct Write { uint64_t epoch, lo, hi; };
  std::deque<Write> wlog;
  uint64_t epoch = 1;              // bumped on every write into the address space
  uint64_t wlog_floor = 0;         // writes with epoch <= floor have been dropped from the log
  // multi-device client: members for devices 1.. (this handle is member 0)
  std::vector<bz_dclient*> peers;
  bool is_member = false;
  // rank of a multi-process world
  int rank = 0, world = 1;
  void* comm = nullptr;            // ncclComm_t
  int* comm_scratch = nullptr;     // device word used by comm_barrier
  std::mutex mu;
};

namespace bz {

// map (and zero) every chunk that intersects [lo, hi); mu must be held
int32_t arena_map(bz_dclient* dc, uint64_t lo, uint64_t hi);
// record a write into [lo, hi); mu must be held
void arena_note_write(bz_dclient* dc, uint64_t lo, uint64_t hi);
// has anything overlapping [lo, hi) been written after `epoch`?  mu must be held
bool arena_dirty_since(bz_dclient* dc, uint64_t epoch, uint64_t lo, uint64_t hi);

// number of member devices (1 for a plain client) and member g (0 = the handle itself)
inline int dc_members(bz_dclient* dc) { return 1 + (int)dc->peers.size(); }
inline bz_dclient* dc_member(bz_dclient* dc, int g) { return g == 0 ? dc : dc->peers[g - 1]; }

// NCCL (resolved with dlopen at first use; absent library => BZ_ERR_NO_DEVICE with a message)
int32_t comm_allgather(bz_dclient* dc, const void* send, void* recv, size_t bytes, cudaStream_t st);
int32_t comm_barrier(bz_dclient* dc, cudaStream_t st);   // stream-ordered: later work on st starts after every rank got here

}  // namespace bz
