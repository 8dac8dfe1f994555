// blaze.hpp -- header-only C++ mirror of the reference's Rust surface over the C ABI
// (include/blaze_b200.h).  Same names, argument meaning and error behaviour as
// /root/reference/src/{driver_client/dclient.rs, ingo_msm/msm_api.rs, ingo_ntt/ntt_api.rs, error.rs};
// `Result<T>` becomes "returns T or throws DriverClientError".
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
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
    std::vector<uint8_t> v((size_t)1 << 32);   // 2^27 elements x 32 B
    check(bz_ntt_result(h_, *buf_num, v.data(), v.size()));
    return v;
  }
  DriverClient driver_client;

 private:
  bz_ntt* h_ = nullptr;
};

}  // namespace ingo_blaze
