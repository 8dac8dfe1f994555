/* blaze_b200 -- C ABI of the B200-native drop-in for ingonyama-zk/blaze's primitive clients.
 *
 * Every entry point below replaces one method of the reference's Rust surface
 * (`DriverClient`, `DriverPrimitive<T,P,I,O>` as implemented by `MSMClient`, `NTTClient`,
 * `PoseidonClient`); the reference file:line each one stands for is cited next to it.
 * INTEGRATION.md shows the Rust `extern "C"` block and the `impl DriverPrimitive` a blaze
 * maintainer would add on top of this header.
 *
 * Conventions
 *   - opaque handles, plain pointers and sizes, no C++/torch types;
 *   - return value: 0 (BZ_OK) or a negative bz_status that maps 1:1 onto the reference's
 *     `DriverClientError` variants (/root/reference/src/error.rs:6-32); nothing panics or aborts
 *     across the ABI (the reference `unwrap()`s / `todo!()`s in several places -- those become
 *     BZ_ERR_INVALID_PRIMITIVE_PARAM here);
 *   - `bz_last_error()` returns a thread-local human-readable message for the last failure;
 *   - the caller owns every buffer; inputs are copied (or consumed by the device) before the
 *     call returns, matching the reference's move-in `Vec<u8>` semantics (msm_api.rs:28-32);
 *   - all data is byte-exact wire format: 32-byte little-endian canonical scalars, affine
 *     x||y little-endian canonical bases (x8 with `2^(32 i) P` when precomputed), Z||Y||X results
 *     (/root/reference/tests/msm/mod.rs:331-332, 360-380, 397-403).
 *   - there is NO CPU fallback: without a CUDA device every constructor returns
 *     BZ_ERR_NO_DEVICE.
 */
#ifndef BLAZE_B200_H
#define BLAZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bz_status {
  BZ_OK = 0,
  BZ_ERR_WRITE = -1,                   /* DriverClientError::WriteError            error.rs:8-13  */
  BZ_ERR_READ = -2,                    /* DriverClientError::ReadError             error.rs:14-19 */
  BZ_ERR_HBICAP_NOT_READY = -3,        /* DriverClientError::HBICAPNotReady        error.rs:20-21 */
  BZ_ERR_INVALID_PRIMITIVE_PARAM = -4, /* DriverClientError::InvalidPrimitiveParam error.rs:22-23 */
  BZ_ERR_CSV = -5,                     /* DriverClientError::CsvError              error.rs:24-25 */
  BZ_ERR_LOAD_FAILED = -6,             /* DriverClientError::LoadFailed            error.rs:26-27 */
  BZ_ERR_FILE = -7,                    /* DriverClientError::FileError             error.rs:28-29 */
  BZ_ERR_UNKNOWN = -8,                 /* DriverClientError::Unknown               error.rs:30-31 */
  BZ_ERR_NO_DEVICE = -9,               /* no usable CUDA device (the reference panics in open_channel, utils.rs:74) */
  BZ_ERR_NO_RESULT = -10               /* result()/wait_result() with an empty task queue */
} bz_status;

/* Curve (msm_cfg.rs:4-8); numeric codes follow the image-parameter word (msm_api.rs:359-364). */
typedef enum bz_curve { BZ_CURVE_BLS377 = 0, BZ_CURVE_BN254 = 1, BZ_CURVE_BLS381 = 2 } bz_curve;
/* PointMemoryType (msm_cfg.rs:10-14) */
typedef enum bz_point_memory_type { BZ_MEM_HBM = 0, BZ_MEM_DMA = 1 } bz_point_memory_type;
/* CardType (dclient_cfg.rs:1-3) plus the card this library drives */
typedef enum bz_card_type { BZ_CARD_C1100 = 0, BZ_CARD_B200 = 1 } bz_card_type;

typedef struct bz_dclient bz_dclient;
typedef struct bz_msm bz_msm;
typedef struct bz_ntt bz_ntt;
typedef struct bz_poseidon bz_poseidon;

const char* bz_last_error(void);
/* library / build identification: "blaze_b200 <version> sm_100a" */
const char* bz_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t bz_kernel_launch_count(void);

/* ------------------------------------------------------------------ DriverClient
 * `id` is the reference's FPGA slot string (dclient.rs:79-86, env ID) = CUDA device ordinal.  A comma-separated list
 * ("0,1,2,3,4,5,6,7") opens ONE client over several devices: an MSMClient created on it shards its bases and scalars
 * over the devices and returns one result per task (the final sum runs device-side over NVLink); NTTClient,
 * PoseidonClient and dma_write / dma_read use the first device of the list. */
int32_t bz_dclient_new(const char* id, int32_t card_type, bz_dclient** out);          /* dclient.rs:79-86   */
/* number of devices behind the handle (1 for a plain id) */
int32_t bz_dclient_device_count(bz_dclient* dc, uint32_t* n);
/* One process per GPU (torchrun / MPI style): turn a single-device client into rank `rank` of `world` processes.
 * `unique_id` is the 128-byte NCCL id made by bz_comm_unique_id() on one rank and handed to all of them by the caller.
 * An MSMClient on such a client treats its inputs as THIS rank's shard of a point-sharded MSM; the ranks' partial
 * results are all-gathered with NCCL and summed on the client's stream, so result() returns the same full sum on every
 * rank.  All ranks must issue the same sequence of tasks.  bz_ntt_dist_* uses the communicator for its handle exchange
 * and its barrier. */
int32_t bz_comm_unique_id(uint8_t out[128]);
int32_t bz_dclient_comm_init(bz_dclient* dc, int32_t rank, int32_t world, const uint8_t unique_id[128]);
int32_t bz_dclient_comm_info(bz_dclient* dc, int32_t* rank, int32_t* world);
int32_t bz_dclient_free(bz_dclient* dc);
int32_t bz_dclient_reset(bz_dclient* dc);                                              /* dclient.rs:88-93   */
/* flat card address space (dclient.rs:456-517): bytes land in / come from the device arena */
int32_t bz_dclient_dma_write(bz_dclient* dc, uint64_t base, uint64_t offset, const uint8_t* data, size_t len);
int32_t bz_dclient_dma_read(bz_dclient* dc, uint64_t base, uint64_t offset, uint8_t* out, size_t len);
/* FPGA-shell management calls the reference's tests make (integration_msm.rs:41-49); there is no
 * bitstream or AXI firewall on a GPU, so these validate the handle and report "healthy". */
int32_t bz_dclient_firewalls_status(bz_dclient* dc, uint32_t* blocked_mask);           /* dclient.rs:566-579 */
int32_t bz_dclient_unblock_firewalls(bz_dclient* dc);                                  /* dclient.rs:258-279 */
int32_t bz_dclient_initialize_cms(bz_dclient* dc);                                     /* dclient.rs:115-131 */
int32_t bz_dclient_reset_sensor_data(bz_dclient* dc);                                  /* dclient.rs:133-151 */
int32_t bz_dclient_setup_before_load_binary(bz_dclient* dc);                           /* dclient.rs:176-187 */
int32_t bz_dclient_load_binary(bz_dclient* dc, const uint8_t* image, size_t len);      /* dclient.rs:213-236 */
/* device introspection (B200 addition): name into buf, total/free HBM bytes */
int32_t bz_dclient_device_info(bz_dclient* dc, char* name, size_t name_len, uint64_t* hbm_total, uint64_t* hbm_free);

/* pinned host memory for callers that want full-rate H2D (B200 addition; optional) */
int32_t bz_host_alloc(size_t bytes, void** out);
int32_t bz_host_free(void* p);

/* ------------------------------------------------------------------ MSMClient (src/ingo_msm/msm_api.rs)
 * Sizes: scalar 32 B; point 96 B (BLS) / 64 B (BN254); result 144 B / 96 B (msm_cfg.rs:44-92). */
int32_t bz_msm_new(bz_dclient* dc, int32_t curve, int32_t mem_type, int32_t is_precompute, bz_msm** out); /* :44-55 */
int32_t bz_msm_free(bz_msm* m);
int32_t bz_msm_loaded_binary_parameters(bz_msm* m, uint32_t out[2]);                   /* :57-70, 333-364 */
/* MSMParams{nof_elements, hbm_point_addr: Option<(u64,u64)>} (:22-26) */
int32_t bz_msm_initialize(bz_msm* m, uint32_t nof_elements, int32_t has_hbm_addr, uint64_t hbm_addr,
                          uint64_t hbm_offset);                                          /* :72-111  */
int32_t bz_msm_start_process(bz_msm* m);                                                /* :113-120 */
/* MSMInput{points: Option<Vec<u8>>, scalars, params} (:28-32); points == NULL means None */
int32_t bz_msm_set_data(bz_msm* m, const uint8_t* points, size_t points_len, const uint8_t* scalars,
                        size_t scalars_len, uint32_t nof_elements, int32_t has_hbm_addr, uint64_t hbm_addr,
                        uint64_t hbm_offset);                                            /* :155-220 */
int32_t bz_msm_wait_result(bz_msm* m);                                                  /* :222-238 */
/* MSMResult{result, result_label} (:33-37); pops the result queue like POP_RESULT (:265-269) */
int32_t bz_msm_result(bz_msm* m, uint8_t* out, size_t out_len, uint32_t* result_label);  /* :240-274 */
int32_t bz_msm_task_label(bz_msm* m, uint32_t* label);                                  /* :278-283 */
int32_t bz_msm_nof_elements(bz_msm* m, uint32_t* n);                                    /* :285-290 */
int32_t bz_msm_is_msm_engine_ready(bz_msm* m, uint32_t* ready);                         /* :292-297 */
int32_t bz_msm_load_data_to_hbm(bz_msm* m, const uint8_t* points, size_t len, uint64_t addr, uint64_t offset); /* :299-313 */
int32_t bz_msm_get_data_from_hbm(bz_msm* m, uint8_t* out, size_t len, uint64_t addr, uint64_t offset);         /* :315-322 */
/* wire sizes of this client's curve (msm_cfg.rs:17-29) */
int32_t bz_msm_sizes(bz_msm* m, uint32_t* scalar_size, uint32_t* point_size, uint32_t* result_point_size,
                     uint32_t* precompute_factor);

/* --- B200 additions (not in the reference) ---
 * phase timers, the analogue of the core's per-phase clock counters (msm_hw_code.rs:33-54, get_api
 * :324-330): milliseconds of the last completed task, measured with CUDA events on the client's stream:
 * [0] total, [1] ingest+digits+sort, [2] bucket accumulation kernel, [3] merge+reduce+finish */
int32_t bz_msm_phase_times(bz_msm* m, float ms[4]);
/* override the window size chosen by the cost model (0 = automatic) */
int32_t bz_msm_set_window_bits(bz_msm* m, int32_t c);
/* get_api() (msm_api.rs:324-330): the register file of the MSM core, indexed by INGO_MSM_ADDR offset / 4
 * (msm_hw_code.rs:6-55; 82 words, 0x000..0x144).  Task / result queue registers reflect the client's queues; the
 * LAST_TASK_PHASE* clock counters are the CUDA-event times of the last completed task in SM clocks (phase 1 = ingest,
 * sort and bucket accumulation, "busy EC adder" = the accumulation kernel alone; phase 2 = final accumulation). */
int32_t bz_msm_get_api(bz_msm* m, uint32_t* regs, size_t n_words);
/* milliseconds the last build of the window-merged table took (0 if none was built) */
int32_t bz_msm_table_build_ms(bz_msm* m, float* ms);
/* bucket-accumulation kernel: -1 (default) chosen per task from the bucket sizes, 0 the XYZZ mixed-add sweep, 2 the
 * fused batched-affine sweep; `rounds` = tree rounds of the latter (-1 automatic).  Same results in every mode. */
int32_t bz_msm_set_accumulate_mode(bz_msm* m, int32_t mode, int32_t rounds);
/* plan of the last launched task: c, W, buckets/window, segment length */
int32_t bz_msm_plan_info(bz_msm* m, uint32_t out[4]);
/* extended plan: c, digit windows (mixed adds per scalar), buckets/set, segment length, bucket sets,
 * merged flag, MiB held by the window-merged table, [7] = final sort bits | partition bits << 8 | levels << 16 |
 * accumulate mode << 24 | batched-affine tree rounds << 28 */
int32_t bz_msm_plan_info_ex(bz_msm* m, uint32_t out[8]);
/* Window-merged table for HBM-resident point sets ("precomputed points resident in HBM"): the client
 * derives 2^(c w) * P_i for every digit window w from the points written with load_data_to_hbm
 * (msm_api.rs:299-313) so that all windows share one bucket set.  mode 0: never; 1 (default): from the second
 * MSM over an unchanged point set; 2: immediately.  Results are identical in every mode.  Env BZ_MSM_PRECOMP
 * sets the default. */
int32_t bz_msm_set_precompute(bz_msm* m, int32_t mode);
/* raw = 1: result records stay homogeneous projective with Z != 1 (x = X/Z, y = Y/Z -- the reference's own result
 * format, tests/msm/mod.rs:397-403) instead of being normalised to Z = 1: saves the field inversion per task for
 * the shards of a multi-GPU MSM, whose records bz_msm_combine_results sums and normalises once.  Default 0. */
int32_t bz_msm_set_raw_result(bz_msm* m, int32_t raw);
/* like set_data(points=None) but the scalars already live in device memory (device pointer) */
int32_t bz_msm_set_scalars_device(bz_msm* m, uint64_t scalars_dev_ptr, uint32_t nof_elements, int32_t has_hbm_addr,
                                  uint64_t hbm_addr, uint64_t hbm_offset);
/* sum `n` result records (host memory, result_point_size each) on the device into one canonical
 * record: the final exchange step of a point-sharded multi-GPU MSM */
int32_t bz_msm_combine_results(bz_msm* m, const uint8_t* records, int32_t n, uint8_t* out, size_t out_len);
/* synthetic-input helper: writes points P0 + (first+i)*Q, i in [0,n), in wire format (factor 1)
 * into the card address space at addr+offset; p0q = P0 || Q in wire format */
int32_t bz_msm_generate_chain_points(bz_msm* m, const uint8_t* p0q, size_t p0q_len, uint64_t first, uint64_t n,
                                     uint64_t addr, uint64_t offset);
/* derive the reference's x8 precomputed records (P, 2^32 P, .., 2^224 P per base; tests/msm/mod.rs:360-380) on the device:
 * n factor-1 bases at src_addr -> n x 8 wire points at dst_addr of the card address space */
int32_t bz_msm_expand_precompute(bz_msm* m, uint64_t src_addr, uint64_t n, uint64_t dst_addr);
/* device self-test of the base-field arithmetic: out[i] = a[i] (op) b[i], canonical LE elements;
 * op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv, 5 neg */
int32_t bz_msm_field_selftest(bz_msm* m, const uint8_t* a, const uint8_t* b, uint8_t* out, int32_t n, int32_t op);

/* ------------------------------------------------------------------ NTTClient (src/ingo_ntt/ntt_api.rs)
 * I/O: flat vector of 2^log_size field elements, 32 bytes little-endian canonical each
 * (ntt_api.rs:20-23, README.md:118).  Semantics fixed by BASELINE.json: BLS12-381 Fr, arkworks
 * Radix2EvaluationDomain::fft -- natural order in and out, out[k] = sum_j in[j] w^(jk).
 * Two buffer slots (ntt_data.rs:42,54-56); start_process transforms a slot in place. */
int32_t bz_ntt_new(bz_dclient* dc, int32_t ntt_type /* NTT::Ntt = 0 */, bz_ntt** out);   /* :26-31, size 2^27 */
int32_t bz_ntt_free(bz_ntt* t);
int32_t bz_ntt_loaded_binary_parameters(bz_ntt* t, uint32_t out[2]);                     /* :33-35 (todo!() there) */
int32_t bz_ntt_initialize(bz_ntt* t);                                                     /* :37-56   */
int32_t bz_ntt_set_data(bz_ntt* t, size_t buf_host, const uint8_t* data, size_t len);     /* :72-87   */
int32_t bz_ntt_start_process(bz_ntt* t, size_t buf_kernel);                               /* :58-70   */
int32_t bz_ntt_wait_result(bz_ntt* t);                                                    /* :89-108  */
int32_t bz_ntt_result(bz_ntt* t, size_t buf_num, uint8_t* out, size_t out_len);           /* :110-124 */
/* --- B200 additions --- */
/* any size 2^log_size (<= 2^30 and <= the field's two-adicity), any of the three scalar fields
 * (bz_curve code), forward or inverse (inverse includes the 1/n scaling) */
int32_t bz_ntt_new_ex(bz_dclient* dc, int32_t field, int32_t log_size, int32_t inverse, bz_ntt** out);
/* device time of the last transform (CUDA events on the client's stream) and its number of passes */
int32_t bz_ntt_phase_times(bz_ntt* t, float* total_ms, uint32_t* passes);
/* device address of the buffer currently holding slot `buf_num` (device-resident use) */
int32_t bz_ntt_slot_device_ptr(bz_ntt* t, size_t buf_num, uint64_t* dev_ptr);

/* Multi-GPU four-step NTT, one process per GPU (B200 addition; config 4 of BASELINE.json).  N = N1 x N2:
 * rank g of G holds the column slab in[j1 N2 + g C + c] (C = N2/G); step1 transforms the columns and
 * its last pass stores every output row, already multiplied by the four-step twiddle, straight into
 * the owning rank's exchange buffer (peer stores over NVLink -- the transpose is fused into the kernel);
 * after a host barrier step3 transforms the rows; rank h ends with X[(h T + t) + N1 k2] (T = N1/G).
 * Exchange buffers are shared through CUDA IPC handles (64 bytes each) passed around by the caller. */
typedef struct bz_ntt_dist bz_ntt_dist;
int32_t bz_ntt_dist_new(bz_dclient* dc, int32_t field, int32_t log_size, int32_t inverse, int32_t rank, int32_t world,
                        bz_ntt_dist** out);
int32_t bz_ntt_dist_free(bz_ntt_dist* t);
int32_t bz_ntt_dist_ipc_handle(bz_ntt_dist* t, uint8_t out[64]);
int32_t bz_ntt_dist_open_peers(bz_ntt_dist* t, const uint8_t* handles /* world x 64 bytes in rank order */);
int32_t bz_ntt_dist_set_input(bz_ntt_dist* t, const uint8_t* full_input, size_t len);    /* natural-order host vector in */
int32_t bz_ntt_dist_get_output(bz_ntt_dist* t, uint8_t* full_output, size_t len);        /* fills this rank's part */
int32_t bz_ntt_dist_buffers(bz_ntt_dist* t, uint64_t* slab_in_dev, uint64_t* block_out_dev, uint64_t* elems_per_rank);
int32_t bz_ntt_dist_step1(bz_ntt_dist* t);
int32_t bz_ntt_dist_sync(bz_ntt_dist* t);   /* drains the rank's stream; the cross-rank barrier is the caller's */
int32_t bz_ntt_dist_step3(bz_ntt_dist* t);
/* whole transform on the client's stream with device-side (NCCL) barriers; needs a ranked DriverClient
 * (bz_dclient_comm_init), which also exchanges the IPC handles inside bz_ntt_dist_new */
int32_t bz_ntt_dist_run(bz_ntt_dist* t);
int32_t bz_ntt_dist_times(bz_ntt_dist* t, float ms[2]);   /* CUDA-event ms of the last step1 / step3 */
int32_t bz_ntt_dist_plan(bz_ntt_dist* t, int32_t out[4]);   /* log2 N1, log2 N2, column passes, row passes */

/* ------------------------------------------------------------------ PoseidonClient (src/ingo_hash/poseidon_api.rs)
 * Stream of 32-byte little-endian BLS12-381 Fr elements in, 64-byte records out:
 * hash[32] || meta[32], meta = LE(hash_id | layer_id << 30) (poseidon_api.rs:42-71).
 * Tree: base node = Poseidon of 11 elements (TreeC) / 8 (TreeD), upper layers arity 8
 * (ingo_hash/utils.rs:2-30); height h gives (8^h - 1)/7 records.  The reference's constants CSV is
 * not in its repository: constants are generated by the library (see oracle/py/poseidon.py for the
 * parameter set) -- hash VALUES are therefore not pinned to the FPGA image, counts and format are. */
typedef enum bz_tree_mode { BZ_TREE_C = 0, BZ_TREE_D = 1 } bz_tree_mode;              /* utils.rs:16-30 */
int32_t bz_poseidon_new(bz_dclient* dc, int32_t hash_type /* Hash::Poseidon = 0 */, bz_poseidon** out); /* :77-79 */
int32_t bz_poseidon_free(bz_poseidon* p);
int32_t bz_poseidon_loaded_binary_parameters(bz_poseidon* p, uint32_t out[2]);          /* :81-94   */
/* PoseidonInitializeParameters{tree_height, tree_mode, instruction_path} (:19-24); a non-empty path must
 * be readable (else BZ_ERR_LOAD_FAILED like :100-103) but its FPGA instruction words are not interpreted */
int32_t bz_poseidon_initialize(bz_poseidon* p, uint32_t tree_height, int32_t tree_mode, const char* instruction_path); /* :96-111 */
/* one element of <= 32 bytes (zero-extended), or a whole number of 32-byte elements */
int32_t bz_poseidon_set_data(bz_poseidon* p, const uint8_t* input, size_t len);         /* :117-122 */
int32_t bz_poseidon_start_process(bz_poseidon* p);                                       /* :113-115 (todo!() there) */
int32_t bz_poseidon_wait_result(bz_poseidon* p);                                         /* :124-126 (todo!() there) */
/* result(Some(expected)): drains records until `expected` were written to out (64 bytes each) */
int32_t bz_poseidon_result(bz_poseidon* p, size_t expected, uint8_t* out, size_t out_cap_records, size_t* n_out); /* :128-146 */
int32_t bz_poseidon_get_last_element_sent_to_ring(bz_poseidon* p, uint32_t* id);        /* :149-154 */
int32_t bz_poseidon_get_num_of_pending_results(bz_poseidon* p, uint32_t* n);            /* :156-161 */
int32_t bz_poseidon_get_raw_results(bz_poseidon* p, uint32_t num_of_results, uint8_t* out); /* :191-196 */
int32_t bz_poseidon_get_last_hash_sent_to_host(bz_poseidon* p, uint32_t* id);           /* :198-203 */
/* --- B200 additions --- */
/* kernel milliseconds spent since initialize() (the analogue of the core's clock counters, hash_hw_code.rs:16-24) */
int32_t bz_poseidon_device_ms(bz_poseidon* p, float* ms);
/* n bare permutations of width t (3, 9 or 12) over full states of canonical elements; mds_mode 0 = the Cauchy matrix
 * the client uses, 1 = the Grain-sampled matrix of the published Poseidon reference vectors (known-answer tests) */
int32_t bz_poseidon_permute(bz_poseidon* p, int32_t t, int32_t mds_mode, const uint8_t* in, size_t n, uint8_t* out);
/* host-only: the preprocessed constants of width t (canonical bytes, kernel order: first-half round constants, MDS,
 * pre-sparse matrix, R_P x (c0, row0[t], col0[t-1]), second-half round constants); out may be NULL to query the size */
int32_t bz_poseidon_optimized_constants(int32_t t, int32_t mds_mode, uint8_t* out, size_t out_cap, size_t* n_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BLAZE_B200_H */
