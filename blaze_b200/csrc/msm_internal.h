// Internal (non-ABI) declarations shared by the MSM translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include <atomic>

namespace bz {

// number of kernels this library has launched in this process (bz_kernel_launch_count)
extern std::atomic<uint64_t> g_kernel_launches;

// device-side error flags (written with atomicExch into MsmWorkspace::err)
enum : int {
  BZ_ERR_NONE = 0,
  BZ_ERR_SCALAR_RANGE = 1,   // a scalar was not canonical (>= r) / top digit overflow
};

// children per thread in the partial-merge tree.  A level costs about `group` sequential additions of latency
// whatever its size: 64 at the first level only when that still fills the machine (millions of segments), else 16;
// 8 above (the upper levels are tiny).
#define MERGE_GROUP 64
#define MERGE_GROUP_SMALL 16
#define MERGE_GROUP_UPPER 8
static inline unsigned merge_group(int level, unsigned long long nseg) {
  if (level > 0) return MERGE_GROUP_UPPER;
  return nseg / MERGE_GROUP >= 148ull * 384 ? MERGE_GROUP : MERGE_GROUP_SMALL;
}

struct DigitConst {
  uint32_t K[9];      // sum of the half-window offsets (see k_digits)
  uint32_t mod[8];    // scalar-field modulus, for the canonical-scalar check
  int check_mod;
};

// Everything the kernels need to know about one MSM launch.
struct MsmPlan {
  uint64_t M;              // number of (sub)scalars == number of table points used
  int words_per_scalar;    // 8: full 256-bit scalars; 1: 32-bit limbs against a x8 precomputed table
  int c;                   // window bits
  int W;                   // number of BUCKET windows (bucket sets): Wd, or 1 when the windows are merged
  int Wd;                  // number of DIGIT windows per scalar (mixed adds per scalar)
  int merged;              // 1: window-merged table -- digit window w of scalar i uses table entry w*M + i
                           //    (= 2^(c w) P_i, precomputed and resident in HBM) and all windows share one bucket set
  uint64_t Ms;             // sort entries per bucket window: M, or Wd*M when merged (dig[][] read as one flat list)
  // Bucket VALUES b per window are 1 .. 2^kb (zero digits are dropped by the sort), kb = c - 1.  The sort key of
  // b is K = ((b-1) mod 2^rest) << fb | (b-1) >> rest: `rest` bits are consumed by the partition levels (LOW
  // bits of b-1 first, so skewed digit ranges still spread over all parents), the last fb bits by the final
  // in-CTA counting sort (msm_sort.cu).  Bucket SLOT of b is K; nb = 2^kb slots per window.
  int kb;                  // key bits = c - 1
  int fb;                  // bits of the final level (<= 8)
  int rest;                // bits of the partition levels = kb - fb
  int nlev;                // partition levels (>= 1; the first one converts digits to {K, ref} pairs)
  int lbits[4];            // bits per partition level (sum = rest)
  int lgs[4];              // log2 of the cursor groups per child at each level (spreads the scatter atomics)
  size_t lvl_hist_words;   // total words of the per-level child histograms (one allocation, zeroed per task)
  uint32_t nvalues;        // 2^kb bucket values per window
  uint32_t nb;             // bucket slots per window = 2^kb
  uint32_t seg_len;        // sorted entries per accumulate thread
  uint64_t nseg;           // number of accumulate segments = ceil(W*M / seg_len)
  int raw_result;          // 1: result record left projective (Z != 1): shard of a multi-GPU MSM, normalised by the combine
  int tma_stage;           // 1: k_accumulate_tma (points staged through shared memory by cp.async.bulk), 0: register prefetch
  int batch_affine;        // 0: XYZZ sweep (k_accumulate); 2: fused batched-affine sweep (msm_ba2.cuh, k_accumulate_ba);
                           // 1: round 1's multi-kernel batched-affine phases (msm_ba.cuh, opt-in, kept for comparison)
  int ba_rounds;           // batch_affine == 2: tree rounds per segment before the XYZZ fold
  uint32_t ba_cap;         //                    scratch slots per lane
  uint32_t ba_ctas;        //                    persistent grid size
  uint32_t chunk;          // entries per reduce thread at every level of the running-sum recursion (power of two)
  uint32_t nchunks;        // ceil(nb / chunk): level-0 chunk count
  DigitConst dc;
};

struct MsmWorkspace {
  uint32_t* dig;       // [Wd][M]  (merged: one window of Wd*M entries)
  uint32_t* lvl_hist[4];    // per partition level: child counts        [W << (bits consumed so far)]
  uint32_t* lvl_off[4];     //   exclusive prefix = child offsets (+1 entry: total)
  uint32_t* lvl_cursor[4];  //   scatter cursors
  uint32_t* lvl_tpref[4];   //   tile-count prefix of the children (next level's tile map)
  uint2* pairA;             // [W*Ms] {K, sign|ref} pairs, output of the odd partition levels (1st, 3rd)
  uint2* pairB;             // [W*Ms] ... of the even ones (null when nlev == 1)
  uint32_t* sorted;    // [W*M]
  uint32_t* goff;      // [W*nb + 1]
  void* buckets;       // [W*nb] XYZZ
  uint32_t* part_id;   // [nseg][2]
  void* part_pt;       // [nseg][2] XYZZ
  uint32_t* part2_id;  // merge-tree levels above the segments: [sum over levels of the group counts][2]
  void* part2_pt;
  void* red_a;         // [2][W*nchunks] XYZZ: S and V outputs of the even reduction levels
  void* red_b;         // ... of the odd levels
  int* err;            // device error flag
  uint8_t* result;     // 3*FQ_BYTES result record (device)
  // batched-affine accumulation (msm_ba.cuh); null when that mode is off
  uint32_t* ba_scalars;           // [4] device scratch words
  uint32_t* ba_pref;              // running denominator products, one field element per output slot
  uint32_t* ba_tot;               // per-thread totals, then ba_itot = their inverses
  uint32_t* ba_itot;
  uint32_t* ba_lvl[3];            // batch-inverse tree: group products per level
  uint32_t* ba_lvlp[2];           //                      prefix / inverse arrays per level
  void* ba_buf0;                  // affine lists of the even rounds (>= 2): (total/4 + W nb) entries
  void* ba_buf1;                  // affine lists of the odd rounds: (total/2 + W nb) entries
  void* ba2_scratch;              // batch_affine == 2: per-warp affine lists + running products
  cudaEvent_t ev_acc0, ev_acc1;   // bracket the accumulate kernel alone (roofline timing); may be null
  // Side ("tail") stream for the latency-bound end of the pipeline -- the small levels of the running-sum reduction, the
  // window combine, the final exchange and the result copy -- so that it overlaps the NEXT task's windowing / sort /
  // accumulation instead of idling 140 SMs (set per task by the client; null = everything on the work stream).
  cudaStream_t tail;
  cudaEvent_t ev_fork;        // work stream: first reduction level done, the tail may start
  cudaEvent_t ev_tail_done;   // tail stream: the previous task's tail has finished with the reduction scratch
  int tail_busy;              // ev_tail_done has been recorded at least once
};

void launch_msm_sort(const MsmPlan& p, const MsmWorkspace& ws, const uint32_t* scalars_dev, cudaStream_t st);

// Per-curve kernels live in msm_curve_*.cu; the engine reaches them through this table.
struct CurveOps {
  int code;              // 0 BLS12_377, 1 BN254, 2 BLS12_381 (reference: msm_api.rs:359-364)
  int fq_bytes;          // 48 / 32
  int scalar_bits;       // 253 / 254 / 255
  const uint32_t* fr_mod;   // 8 limbs, host copy
  size_t affine_bytes;   // Montgomery affine TABLE entry (padded for single-line gathers)
  size_t affine_list_bytes;   // packed affine point (batched-affine lists)
  size_t xyzz_bytes;
  void (*points_to_mont)(const uint8_t* raw, void* table, uint64_t n, cudaStream_t st);
  // returns the stream on which ws.result becomes ready: `st`, or ws.tail when the end of the pipeline was forked there
  cudaStream_t (*bucket_phase)(const MsmPlan& p, const MsmWorkspace& ws, const void* table, cudaStream_t st);
  // window-merged table: level w (entries [w*n, (w+1)*n)) = 2^c * level w-1; level 0 must already be in place
  void (*build_wtable)(void* wtable, uint64_t n, int levels, int c, cudaStream_t st);
  // level-major table (entry w*n + i) -> wire records (levels affine points per base, canonical LE)
  void (*table_to_wire)(const void* table, uint64_t n, int levels, uint8_t* out, cudaStream_t st);
  // sum `n` result records (Z||Y||X, any Z) into one record: canonical (Z = 1), or left projective when raw != 0
  void (*combine_results)(const uint8_t* recs, int n, uint8_t* out, int raw, cudaStream_t st);
  // test / bench helpers
  void (*gen_chain_points)(const uint8_t* p0q_raw, uint64_t first, uint64_t n, uint8_t* out_raw, cudaStream_t st);
  void (*field_selftest)(const uint8_t* a, const uint8_t* b, uint8_t* out, int n, int op, cudaStream_t st);
};

const CurveOps* curve_ops_bls12_377();
const CurveOps* curve_ops_bn254();
const CurveOps* curve_ops_bls12_381();

}  // namespace bz
