// PoseidonClient half of the C ABI (include/blaze_b200.h): element stream in, (hash, id, layer)
// records out, 8-ary Merkle tree bookkeeping.
//
// Mirrors the behaviour of /root/reference/src/ingo_hash/poseidon_api.rs:
//   initialize(tree_height, tree_mode, instruction_path)   :96-111
//   set_data(&[u8])  one field element per call             :117-122 (tests send <= 32 bytes LE)
//   result(Some(expected)) -> Vec<PoseidonResult>           :128-146, record format :42-71
//   get_num_of_pending_results / get_raw_results / get_last_element_sent_to_ring /
//   get_last_hash_sent_to_host                              :149-203
// Tree shape: base node = hash of 11 elements (TreeC; tests/integration_poseidon.rs:109-116) or 8
// (TreeD start layer, ingo_hash/utils.rs:16-30), upper layers arity 8 (utils.rs:2-14).
// The reference's constants CSV is absent: constants are generated here (Grain LFSR), parity UNPINNED.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/blaze_b200.h"
#include "api_common.h"
#include "poseidon_internal.h"

using namespace bz;

static const int P_RF = 8, P_RP = 57;
// BLS12-381 Fr, big-endian bytes for the rejection test
static const uint8_t FR381_BE[32] = {0x73, 0xed, 0xa7, 0x53, 0x29, 0x9d, 0x7d, 0x48, 0x33, 0x39, 0xd8, 0x08, 0x09, 0xa1, 0xd8, 0x05,
                                     0x53, 0xbd, 0xa4, 0x02, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0xff, 0xff, 0xff, 0x00, 0x00, 0x00, 0x01};

// Grain LFSR of the Poseidon reference parameter generator (self-shrinking output)
struct Grain {
  uint8_t s[80];
  int pos = 0;
  static void put(uint8_t*& p, uint32_t v, int w) { for (int i = w - 1; i >= 0; i--) *p++ = (v >> i) & 1; }
  Grain(int n, int t, int rf, int rp) {
    uint8_t* p = s;
    put(p, 1, 2); put(p, 0, 4); put(p, n, 12); put(p, t, 12); put(p, rf, 10); put(p, rp, 10);
    for (int i = 0; i < 30; i++) *p++ = 1;
    for (int i = 0; i < 160; i++) step();
  }
  int at(int i) const { return s[(pos + i) % 80]; }
  int step() {
    int b = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
    s[pos] = (uint8_t)b;   // overwrite the oldest bit: it becomes the newest
    pos = (pos + 1) % 80;
    return b;
  }
  int bit() {
    for (;;) {
      int a = step(), b = step();
      if (a) return b;
    }
  }
  // 255-bit candidates, most significant bit first, rejection-sampled below r; out = 32-byte LE
  void element(uint8_t out[32]) {
    for (;;) {
      uint8_t be[32];
      memset(be, 0, 32);
      for (int i = 0; i < 255; i++) {
        int bitpos = 254 - i;   // value bit index
        if (bit()) be[31 - bitpos / 8] |= (uint8_t)(1u << (bitpos % 8));
      }
      if (memcmp(be, FR381_BE, 32) < 0) {
        for (int i = 0; i < 32; i++) out[i] = be[31 - i];
        return;
      }
    }
  }
};

struct PoseidonParams {
  uint4* rc = nullptr;    // (RF+RP)*t elements, Montgomery
  uint4* mds = nullptr;   // t*t
  bool ready = false;
};

struct bz_poseidon {
  bz_dclient* dc = nullptr;
  bool initialized = false;
  uint32_t height = 0;
  int tree_mode = 0;
  int in_arity = 11;
  PoseidonParams par[2];   // [0]: t = 9 (arity 8), [1]: t = 12 (arity 11)
  // element stream
  std::vector<uint8_t> staged;      // host staging of elements not yet on the device (32 B each)
  uint64_t elems_total = 0;         // elements received since initialize (ring id = elems_total - 1)
  uint64_t elems_tree = 0;          // elements of the current tree
  uint64_t elems_on_device = 0;
  uint8_t* d_inputs = nullptr;      // base-layer inputs of the current tree
  uint64_t d_inputs_cap = 0;
  std::vector<uint8_t*> d_layer;    // node digests per layer
  std::vector<uint64_t> layer_size, done;
  std::deque<std::vector<uint8_t>> pending;   // 64-byte records
  uint32_t last_hash_id = 0;
  std::mutex mu;
};

static int32_t ensure_params(bz_poseidon* p, int t) {
  PoseidonParams& q = p->par[t == 9 ? 0 : 1];
  if (q.ready) return BZ_OK;
  int n_rc = (P_RF + P_RP) * t;
  std::vector<uint8_t> host((size_t)n_rc * 32);
  Grain g(255, t, P_RF, P_RP);
  for (int i = 0; i < n_rc; i++) g.element(&host[(size_t)i * 32]);
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&q.rc, (size_t)n_rc * 32));
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&q.mds, (size_t)t * t * 32));
  cudaStream_t st = dc_stream(p->dc);
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(q.rc, host.data(), host.size(), cudaMemcpyHostToDevice, st));
  poseidon_prepare(q.rc, n_rc, q.mds, t, st);
  CUDA_TRY(BZ_ERR_UNKNOWN, cudaStreamSynchronize(st));
  q.ready = true;
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_new(bz_dclient* dc, int32_t hash_type, bz_poseidon** out) {
  if (!out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  *out = nullptr;
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  if (hash_type != 0) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown hash type %d", hash_type);
  bz_poseidon* p = new bz_poseidon();
  p->dc = dc;
  *out = p;
  return BZ_OK;
}

static void free_tree(bz_poseidon* p) {
  for (auto b : p->d_layer) if (b) cudaFree(b);
  p->d_layer.clear();
  if (p->d_inputs) cudaFree(p->d_inputs);
  p->d_inputs = nullptr;
  p->d_inputs_cap = 0;
}

extern "C" int32_t bz_poseidon_free(bz_poseidon* p) {
  if (!p) return BZ_OK;
  cudaSetDevice(dc_device(p->dc));
  cudaStreamSynchronize(dc_stream(p->dc));
  free_tree(p);
  for (auto& q : p->par) { if (q.rc) cudaFree(q.rc); if (q.mds) cudaFree(q.mds); }
  delete p;
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_loaded_binary_parameters(bz_poseidon* p, uint32_t out[2]) {
  if (!p || !out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  out[0] = 0xB2000200u;                                   // image id
  out[1] = (uint32_t)P_RF | ((uint32_t)P_RP << 8) | (2u << 20);   // rounds | curve code (BLS12-381)
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_initialize(bz_poseidon* p, uint32_t tree_height, int32_t tree_mode, const char* instruction_path) {
  if (!p) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null PoseidonClient");
  int32_t rc = dc_select(p->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(p->mu);
  // The reference streams an instruction/constant CSV into the core (poseidon_api.rs:205-243) and
  // maps any failure to LoadFailed (:100-103).  That file's contents are FPGA-specific and not in
  // the repository; if a path is given it must at least be readable, the constants used are ours.
  if (instruction_path && *instruction_path) {
    FILE* f = fopen(instruction_path, "rb");
    if (!f) return bz_fail(BZ_ERR_LOAD_FAILED, "failed to load instruction set from: %s", instruction_path);
    fclose(f);
  }
  if (tree_height < 1 || tree_height > 10) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "tree height %u unsupported", tree_height);
  if (tree_mode != 0 && tree_mode != 1) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "unknown tree mode %d", tree_mode);
  free_tree(p);
  p->height = tree_height;
  p->tree_mode = tree_mode;
  p->in_arity = tree_mode == 0 ? 11 : 8;     // TreeC: column hash of 11; TreeD: start one layer up
  rc = ensure_params(p, 9);
  if (rc) return rc;
  if (p->in_arity == 11) { rc = ensure_params(p, 12); if (rc) return rc; }
  p->layer_size.assign(tree_height, 0);
  p->done.assign(tree_height, 0);
  uint64_t sz = 1;
  for (int l = (int)tree_height - 1; l >= 0; l--) { p->layer_size[l] = sz; sz *= 8; }
  p->d_layer.assign(tree_height, nullptr);
  for (uint32_t l = 0; l < tree_height; l++) CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&p->d_layer[l], p->layer_size[l] * 32));
  p->d_inputs_cap = p->layer_size[0] * p->in_arity;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&p->d_inputs, p->d_inputs_cap * 32));
  p->staged.clear();
  p->pending.clear();
  p->elems_total = p->elems_tree = p->elems_on_device = 0;
  p->last_hash_id = 0;
  p->initialized = true;
  return BZ_OK;
}

// hash every node whose inputs are complete; append the records
static int32_t flush(bz_poseidon* p) {
  if (!p->initialized) return BZ_OK;
  cudaStream_t st = dc_stream(p->dc);
  if (!p->staged.empty()) {
    CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(p->d_inputs + p->elems_on_device * 32, p->staged.data(), p->staged.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(st));
    p->elems_on_device += p->staged.size() / 32;
    p->staged.clear();
  }
  std::vector<uint8_t> host;
  for (uint32_t l = 0; l < p->height; l++) {
    uint64_t avail = l == 0 ? p->elems_on_device / p->in_arity : p->done[l - 1] / 8;
    if (avail <= p->done[l]) break;
    uint64_t n = avail - p->done[l];
    int t = (l == 0 && p->in_arity == 11) ? 12 : 9;
    const PoseidonParams& q = p->par[t == 9 ? 0 : 1];
    const uint8_t* in = l == 0 ? p->d_inputs + p->done[0] * (uint64_t)p->in_arity * 32 : p->d_layer[l - 1] + p->done[l] * 8 * 32;
    uint8_t* out = p->d_layer[l] + p->done[l] * 32;
    poseidon_hash(t, (const uint4*)in, n, q.rc, q.mds, P_RF, P_RP, (uint4*)out, st);
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
    host.resize(n * 32);
    CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(host.data(), out, n * 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(BZ_ERR_READ, cudaStreamSynchronize(st));
    for (uint64_t i = 0; i < n; i++) {
      std::vector<uint8_t> rec(64, 0);
      memcpy(rec.data(), &host[i * 32], 32);
      uint64_t id = p->done[l] + i;
      uint64_t meta = (id & 0x3fffffffull) | ((uint64_t)l << 30);   // poseidon_api.rs:50-61
      memcpy(rec.data() + 32, &meta, 8);
      p->pending.push_back(std::move(rec));
    }
    p->done[l] = avail;
  }
  // tree complete: the next element starts a new tree
  if (p->done[p->height - 1] == 1) {
    std::fill(p->done.begin(), p->done.end(), 0);
    p->elems_on_device = 0;
    p->elems_tree = 0;
  }
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_set_data(bz_poseidon* p, const uint8_t* input, size_t len) {
  if (!p || (!input && len)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(p->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!p->initialized) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "set_data before initialize");
  // <= 32 bytes: ONE little-endian element, zero-extended (the tests send BigUint::to_bytes_le() and
  // u32::to_le_bytes(), integration_poseidon.rs:49-51,109-116); longer: a whole number of 32-byte elements
  size_t n_el = len <= 32 ? 1 : len / 32;
  if (len > 32 && len % 32) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bulk set_data length must be a multiple of 32");
  for (size_t e = 0; e < n_el; e++) {
    if (p->elems_tree == p->d_inputs_cap) {   // previous tree's inputs are full: finish it first
      rc = flush(p);
      if (rc) return rc;
      if (p->elems_tree == p->d_inputs_cap) return bz_fail(BZ_ERR_WRITE, "input ring full");
    }
    uint8_t el[32];
    memset(el, 0, 32);
    memcpy(el, input + e * 32, len <= 32 ? len : 32);
    p->staged.insert(p->staged.end(), el, el + 32);
    p->elems_total++;
    p->elems_tree++;
  }
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_get_num_of_pending_results(bz_poseidon* p, uint32_t* n) {
  if (!p || !n) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  int32_t rc = dc_select(p->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(p->mu);
  rc = flush(p);
  if (rc) return rc;
  *n = (uint32_t)p->pending.size();
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_get_raw_results(bz_poseidon* p, uint32_t num_of_results, uint8_t* out) {
  if (!p || (!out && num_of_results)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(p->mu);
  if (num_of_results > p->pending.size()) return bz_fail(BZ_ERR_READ, "only %zu results pending", p->pending.size());
  for (uint32_t i = 0; i < num_of_results; i++) {
    const std::vector<uint8_t>& rec = p->pending.front();
    memcpy(out + (size_t)i * 64, rec.data(), 64);
    uint32_t id;
    memcpy(&id, rec.data() + 32, 4);
    p->last_hash_id = id & 0x3fffffffu;
    p->pending.pop_front();
  }
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_result(bz_poseidon* p, size_t expected, uint8_t* out, size_t out_cap_records, size_t* n_out) {
  if (!p || !out || !n_out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  *n_out = 0;
  // poseidon_api.rs:128-146: keep draining until `expected` records were collected.  The FPGA may
  // still be hashing; here everything that CAN be hashed is hashed by flush(), so if fewer than
  // `expected` records exist after it the reference would spin forever -- we report NoResult.
  size_t got = 0;
  for (int guard = 0; got < expected && guard < 2; guard++) {
    uint32_t n = 0;
    int32_t rc = bz_poseidon_get_num_of_pending_results(p, &n);
    if (rc) return rc;
    if (got + n > out_cap_records) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "result buffer too small (%zu records)", out_cap_records);
    rc = bz_poseidon_get_raw_results(p, n, out + got * 64);
    if (rc) return rc;
    got += n;
  }
  *n_out = got;
  if (got < expected) return bz_fail(BZ_ERR_NO_RESULT, "%zu of %zu expected records available (not enough input elements)", got, expected);
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_get_last_element_sent_to_ring(bz_poseidon* p, uint32_t* id) {
  if (!p || !id) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  *id = (uint32_t)(p->elems_total - 1);   // u32 wrap like a hardware counter before the first element
  return BZ_OK;
}
extern "C" int32_t bz_poseidon_get_last_hash_sent_to_host(bz_poseidon* p, uint32_t* id) {
  if (!p || !id) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  *id = p->last_hash_id;
  return BZ_OK;
}
// start_process / wait_result are `todo!()` in the reference (poseidon_api.rs:113-115,124-126);
// here: "hash whatever is complete now" and "nothing to wait for".
extern "C" int32_t bz_poseidon_start_process(bz_poseidon* p) {
  if (!p) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null PoseidonClient");
  int32_t rc = dc_select(p->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(p->mu);
  return flush(p);
}
extern "C" int32_t bz_poseidon_wait_result(bz_poseidon* p) { return bz_poseidon_start_process(p); }
