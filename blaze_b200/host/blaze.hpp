// blaze.hpp -- header-only C++ mirror of the reference's Rust surface over the C ABI
// (include/blaze_b200.h).  Same names, argument meaning and error behaviour as
// /root/reference/src/{driver_client/dclient.rs, ingo_msm/msm_api.rs, ingo_ntt/ntt_api.rs, ingo_hash/poseidon_api.rs,
// ingo_hash/utils.rs, error.rs};
// `Result<T>` becomes "returns T or throws DriverClientError".
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/blaze_b200.h"

namespace ingo_blaze {

// ---- error.rs:6-32
struct DriverClientError : std::runtime_error {
  enum Variant { WriteError, ReadError, HBICAPNotReady, InvalidPrimitiveParam, CsvError, LoadFailed, FileError, Unknown, NoDevice, NoResult };
  Variant variant;
  DriverClientError(Variant v, const std::string& m) : std::runtime_error(m), variant(v) {}
};
inline void check(int32_t rc) {
  if (rc == BZ_OK) return;
  static const DriverClientError::Variant map[] = {
      DriverClientError::Unknown,       DriverClientError::WriteError, DriverClientError::ReadError,
      DriverClientError::HBICAPNotReady, DriverClientError::InvalidPrimitiveParam, DriverClientError::CsvError,
      DriverClientError::LoadFailed,    DriverClientError::FileError,  DriverClientError::Unknown,
      DriverClientError::NoDevice,      DriverClientError::NoResult};
  int i = -rc;
  throw DriverClientError(i >= 1 && i <= 10 ? map[i] : DriverClientError::Unknown, bz_last_error());
}

// ---- driver_client: dclient_cfg.rs:1-47, dclient.rs:50-592
enum class CardType { C1100 = BZ_CARD_C1100, B200 = BZ_CARD_B200 };
struct DriverConfig {
  CardType card_type = CardType::B200;
  static DriverConfig driver_client_cfg(CardType c) { return DriverConfig{c}; }
};
class DriverClient {
 public:
  DriverClient(const std::string& id, DriverConfig cfg) : cfg(cfg) { check(bz_dclient_new(id.c_str(), (int32_t)cfg.card_type, &h_)); }
  DriverClient(DriverClient&& o) noexcept : cfg(o.cfg), h_(o.h_) { o.h_ = nullptr; }
  DriverClient(const DriverClient&) = delete;
  ~DriverClient() { if (h_) bz_dclient_free(h_); }
  void reset() { check(bz_dclient_reset(h_)); }                                                       // dclient.rs:88-93
  void dma_write(uint64_t base, uint64_t off, const std::vector<uint8_t>& d) { check(bz_dclient_dma_write(h_, base, off, d.data(), d.size())); }
  void dma_read(uint64_t base, uint64_t off, std::vector<uint8_t>& out) { check(bz_dclient_dma_read(h_, base, off, out.data(), out.size())); }
  void firewalls_status() { uint32_t m; check(bz_dclient_firewalls_status(h_, &m)); }                 // dclient.rs:566-579
  void unblock_firewalls() { check(bz_dclient_unblock_firewalls(h_)); }
  void initialize_cms() { check(bz_dclient_initialize_cms(h_)); }
  void reset_sensor_data() { check(bz_dclient_reset_sensor_data(h_)); }
  void setup_before_load_binary() { check(bz_dclient_setup_before_load_binary(h_)); }
  uint32_t load_binary(const std::vector<uint8_t>& b) { check(bz_dclient_load_binary(h_, b.data(), b.size())); return 0; }
  // ---- multi-GPU (B200 additions).  `id` may be a device list ("0,1,2,3"): one client over several GPUs.
  uint32_t device_count() { uint32_t n; check(bz_dclient_device_count(h_, &n)); return n; }
  // one process per GPU: make the 128-byte id on one rank, hand it to all ranks, then comm_init on each
  static std::vector<uint8_t> comm_unique_id() { std::vector<uint8_t> v(128); check(bz_comm_unique_id(v.data())); return v; }
  void comm_init(int rank, int world, const std::vector<uint8_t>& unique_id) {
    if (unique_id.size() != 128) throw DriverClientError(DriverClientError::InvalidPrimitiveParam, "unique id must be 128 bytes");
    check(bz_dclient_comm_init(h_, rank, world, unique_id.data()));
  }
  bz_dclient* raw() const { return h_; }
  DriverConfig cfg;

 private:
  bz_dclient* h_ = nullptr;
};

// ---- ingo_msm: msm_cfg.rs:4-14, msm_api.rs:16-40
enum class Curve { BLS377 = BZ_CURVE_BLS377, BN254 = BZ_CURVE_BN254, BLS381 = BZ_CURVE_BLS381 };
enum class PointMemoryType { HBM = BZ_MEM_HBM, DMA = BZ_MEM_DMA };
constexpr uint32_t PRECOMPUTE_FACTOR_BASE = 1, PRECOMPUTE_FACTOR = 8;
struct MSMInit { PointMemoryType mem_type; bool is_precompute; Curve curve; };
struct MSMParams { uint32_t nof_elements; std::optional<std::pair<uint64_t, uint64_t>> hbm_point_addr; };
struct MSMInput { std::optional<std::vector<uint8_t>> points; std::vector<uint8_t> scalars; MSMParams params; };
struct MSMResult { std::vector<uint8_t> result; uint32_t result_label; };

class MSMClient {   // impl DriverPrimitive<MSMInit, MSMParams, MSMInput, MSMResult>, msm_api.rs:42-274
 public:
  MSMClient(MSMInit init, DriverClient dclient) : driver_client(std::move(dclient)) {
    check(bz_msm_new(driver_client.raw(), (int32_t)init.curve, (int32_t)init.mem_type, init.is_precompute, &h_));
    uint32_t s, p, f;
    check(bz_msm_sizes(h_, &s, &p, &result_point_size_, &f));
  }
  MSMClient(const MSMClient&) = delete;
  ~MSMClient() { if (h_) bz_msm_free(h_); }
  std::vector<uint32_t> loaded_binary_parameters() { uint32_t v[2]; check(bz_msm_loaded_binary_parameters(h_, v)); return {v[0], v[1]}; }
  void initialize(MSMParams p) {
    check(bz_msm_initialize(h_, p.nof_elements, p.hbm_point_addr.has_value(), p.hbm_point_addr ? p.hbm_point_addr->first : 0,
                            p.hbm_point_addr ? p.hbm_point_addr->second : 0));
  }
  void start_process(std::optional<size_t> = std::nullopt) { check(bz_msm_start_process(h_)); }
  void set_data(MSMInput d) {   // by value: move-in, like the reference
    const auto& a = d.params.hbm_point_addr;
    check(bz_msm_set_data(h_, d.points ? d.points->data() : nullptr, d.points ? d.points->size() : 0, d.scalars.data(),
                          d.scalars.size(), d.params.nof_elements, a.has_value(), a ? a->first : 0, a ? a->second : 0));
  }
  void wait_result() { check(bz_msm_wait_result(h_)); }
  std::optional<MSMResult> result(std::optional<size_t> = std::nullopt) {
    MSMResult r{std::vector<uint8_t>(result_point_size_), 0};
    check(bz_msm_result(h_, r.result.data(), r.result.size(), &r.result_label));
    return r;
  }
  uint32_t task_label() { uint32_t v; check(bz_msm_task_label(h_, &v)); return v; }
  uint32_t nof_elements() { uint32_t v; check(bz_msm_nof_elements(h_, &v)); return v; }
  uint32_t is_msm_engine_ready() { uint32_t v; check(bz_msm_is_msm_engine_ready(h_, &v)); return v; }
  void load_data_to_hbm(const std::vector<uint8_t>& pts, uint64_t addr, uint64_t off) { check(bz_msm_load_data_to_hbm(h_, pts.data(), pts.size(), addr, off)); }
  std::vector<uint8_t> get_data_from_hbm(size_t len, uint64_t addr, uint64_t off) {
    std::vector<uint8_t> v(len);
    check(bz_msm_get_data_from_hbm(h_, v.data(), len, addr, off));
    return v;
  }
  // get_api(), msm_api.rs:324-330: the register file (msm_hw_code.rs:6-55), word index = offset / 4
  std::vector<uint32_t> get_api() { std::vector<uint32_t> v(82); check(bz_msm_get_api(h_, v.data(), v.size())); return v; }
  // ---- B200 additions (all optional)
  float table_build_ms() { float v; check(bz_msm_table_build_ms(h_, &v)); return v; }
  // -1 automatic / 0 XYZZ sweep / 2 fused batched-affine sweep
  void set_accumulate_mode(int mode, int rounds = -1) { check(bz_msm_set_accumulate_mode(h_, mode, rounds)); }
  // n factor-1 bases at src -> the reference's x8 precomputed records at dst, derived on the device
  void expand_precompute(uint64_t src_addr, uint64_t n, uint64_t dst_addr) { check(bz_msm_expand_precompute(h_, src_addr, n, dst_addr)); }
  struct PhaseTimes { float total, sort, accumulate, reduce; };
  PhaseTimes phase_times() { float v[4]; check(bz_msm_phase_times(h_, v)); return {v[0], v[1], v[2], v[3]}; }
  struct PlanInfo { uint32_t c, windows, buckets_per_set, segment, bucket_sets; bool merged_table; uint32_t merged_table_mib; };
  PlanInfo plan_info() { uint32_t v[8]; check(bz_msm_plan_info_ex(h_, v)); return {v[0], v[1], v[2], v[3], v[4], v[5] != 0, v[6]}; }
  void set_window_bits(int c) { check(bz_msm_set_window_bits(h_, c)); }
  // 0 never / 1 on reuse (default) / 2 always: table of window multiples for an HBM-resident point set
  void set_precompute(int mode) { check(bz_msm_set_precompute(h_, mode)); }
  void set_scalars_device(uint64_t dev_ptr, const MSMParams& p) {
    const auto& a = p.hbm_point_addr;
    check(bz_msm_set_scalars_device(h_, dev_ptr, p.nof_elements, a.has_value(), a ? a->first : 0, a ? a->second : 0));
  }
  std::vector<uint8_t> combine_results(const std::vector<uint8_t>& records, int n) {
    std::vector<uint8_t> out(result_point_size_);
    check(bz_msm_combine_results(h_, records.data(), n, out.data(), out.size()));
    return out;
  }
  void generate_chain_points(const std::vector<uint8_t>& p0q, uint64_t first, uint64_t n, uint64_t addr, uint64_t off) {
    check(bz_msm_generate_chain_points(h_, p0q.data(), p0q.size(), first, n, addr, off));
  }
  uint32_t result_point_size() const { return result_point_size_; }
  DriverClient driver_client;   // pub field in the reference (msm_api.rs:13)

 private:
  bz_msm* h_ = nullptr;
  uint32_t result_point_size_ = 0;
};

// ---- ingo_ntt: ntt_api.rs:8-125
enum class NTT { Ntt = 0 };
struct NttInit {};
struct NTTInput { size_t buf_host; std::vector<uint8_t> data; };
class NTTClient {
 public:
  NTTClient(NTT t, DriverClient dclient) : driver_client(std::move(dclient)) { check(bz_ntt_new(driver_client.raw(), (int32_t)t, &h_)); }
  // B200 addition: any size / scalar field / direction
  NTTClient(DriverClient dclient, Curve field, int log_size, bool inverse) : driver_client(std::move(dclient)), log_size_(log_size) {
    check(bz_ntt_new_ex(driver_client.raw(), (int32_t)field, log_size, inverse, &h_));
  }
  NTTClient(const NTTClient&) = delete;
  ~NTTClient() { if (h_) bz_ntt_free(h_); }
  void initialize(NttInit) { check(bz_ntt_initialize(h_)); }
  void set_data(NTTInput in) { check(bz_ntt_set_data(h_, in.buf_host, in.data.data(), in.data.size())); }
  void start_process(std::optional<size_t> buf_kernel) {
    if (!buf_kernel) throw DriverClientError(DriverClientError::InvalidPrimitiveParam, "start_process(None)");   // unwrap() in the reference
    check(bz_ntt_start_process(h_, *buf_kernel));
  }
  void wait_result() { check(bz_ntt_wait_result(h_)); }
  std::optional<std::vector<uint8_t>> result(std::optional<size_t> buf_num) {
    if (!buf_num) throw DriverClientError(DriverClientError::InvalidPrimitiveParam, "result(None)");
    std::vector<uint8_t> v((size_t)32 << log_size_);   // 2^27 elements x 32 B for the reference constructor
    check(bz_ntt_result(h_, *buf_num, v.data(), v.size()));
    return v;
  }
  float phase_ms() { float ms; uint32_t passes; check(bz_ntt_phase_times(h_, &ms, &passes)); return ms; }
  DriverClient driver_client;

 private:
  bz_ntt* h_ = nullptr;
  int log_size_ = 27;
};

// ---- ingo_hash: utils.rs:2-30, poseidon_api.rs:11-253
inline uint32_t num_of_elements_oct_tree(uint32_t tree_height) {        // utils.rs:2-10
  uint32_t n = 0, layer = 1;
  for (uint32_t i = 0; i < tree_height; i++) { n += layer; layer *= 8; }
  return n;
}
inline uint32_t num_of_elements_in_base_layer(uint32_t tree_height) {   // utils.rs:12-14
  uint32_t layer = 1;
  for (uint32_t i = 1; i < tree_height; i++) layer *= 8;
  return layer;
}
enum class TreeMode { TreeC = 0, TreeD = 1 };                            // utils.rs:16-30
enum class Hash { Poseidon = 0 };                                        // poseidon_api.rs:11-13
struct PoseidonInitializeParameters { uint32_t tree_height; TreeMode tree_mode; std::string instruction_path; };   // :19-24
struct PoseidonResult {                                                  // :26-30
  std::vector<uint8_t> hash_byte;
  uint32_t hash_id, layer_id;
  // 64-byte records: hash[32] || meta[32]; hash_id = LE32(meta[0..4]) & 0x3fffffff, layer_id = LE32(meta[3..5],0,0) >> 6  (:42-71)
  static std::vector<PoseidonResult> parse_poseidon_hash_results(const std::vector<uint8_t>& data) {
    std::vector<PoseidonResult> out;
    for (size_t i = 0; i + 64 <= data.size(); i += 64) {
      const uint8_t* el = data.data() + i;
      uint32_t id, ly = (uint32_t)el[35] | ((uint32_t)el[36] << 8);
      memcpy(&id, el + 32, 4);
      out.push_back({std::vector<uint8_t>(el, el + 32), id & 0x3fffffffu, ly >> 6});
    }
    return out;
  }
};
class PoseidonClient {   // impl DriverPrimitive<Hash, PoseidonInitializeParameters, &[u8], Vec<PoseidonResult>>, :74-146
 public:
  PoseidonClient(Hash t, DriverClient dclient) : dclient(std::move(dclient)) { check(bz_poseidon_new(this->dclient.raw(), (int32_t)t, &h_)); }
  PoseidonClient(const PoseidonClient&) = delete;
  ~PoseidonClient() { if (h_) bz_poseidon_free(h_); }
  std::vector<uint32_t> loaded_binary_parameters() { uint32_t v[2]; check(bz_poseidon_loaded_binary_parameters(h_, v)); return {v[0], v[1]}; }
  void initialize(const PoseidonInitializeParameters& p) {
    check(bz_poseidon_initialize(h_, p.tree_height, (int32_t)p.tree_mode, p.instruction_path.c_str()));
  }
  void set_data(const uint8_t* input, size_t len) { check(bz_poseidon_set_data(h_, input, len)); }
  void set_data(const std::vector<uint8_t>& input) { set_data(input.data(), input.size()); }
  void start_process(std::optional<size_t> = std::nullopt) { check(bz_poseidon_start_process(h_)); }   // todo!() in the reference
  void wait_result() { check(bz_poseidon_wait_result(h_)); }                                           // todo!() in the reference
  std::optional<std::vector<PoseidonResult>> result(std::optional<size_t> expected_result) {
    if (!expected_result) throw DriverClientError(DriverClientError::InvalidPrimitiveParam, "result(None)");
    size_t cap = std::max<size_t>(*expected_result, get_num_of_pending_results()) + 8, got = 0;
    std::vector<uint8_t> buf(cap * 64);
    check(bz_poseidon_result(h_, *expected_result, buf.data(), cap, &got));
    buf.resize(got * 64);
    return PoseidonResult::parse_poseidon_hash_results(buf);
  }
  uint32_t get_last_element_sent_to_ring() { uint32_t v; check(bz_poseidon_get_last_element_sent_to_ring(h_, &v)); return v; }
  uint32_t get_num_of_pending_results() { uint32_t v; check(bz_poseidon_get_num_of_pending_results(h_, &v)); return v; }
  std::vector<uint8_t> get_raw_results(uint32_t n) { std::vector<uint8_t> v((size_t)n * 64); check(bz_poseidon_get_raw_results(h_, n, v.data())); return v; }
  uint32_t get_last_hash_sent_to_host() { uint32_t v; check(bz_poseidon_get_last_hash_sent_to_host(h_, &v)); return v; }
  // ---- B200 additions
  float device_ms() { float v; check(bz_poseidon_device_ms(h_, &v)); return v; }
  std::vector<uint8_t> permute(const std::vector<uint8_t>& states, int t, int mds_mode = 0) {   // bare permutation (KATs)
    std::vector<uint8_t> out(states.size());
    check(bz_poseidon_permute(h_, t, mds_mode, states.data(), states.size() / (32 * (size_t)t), out.data()));
    return out;
  }
  DriverClient dclient;   // pub field `dclient` in the reference (poseidon_api.rs:15-17)

 private:
  bz_poseidon* h_ = nullptr;
};

}  // namespace ingo_blaze
