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
  // one 255-bit draw, NOT rejection-sampled (the caller reduces it mod r): how the reference generator samples the MDS
  void element_mod(uint8_t out[32]) {
    memset(out, 0, 32);
    for (int i = 0; i < 255; i++) {
      int bitpos = 254 - i;
      if (bit()) out[bitpos / 8] |= (uint8_t)(1u << (bitpos % 8));
    }
  }
};

// ---- host-side Fr(BLS12-381) for the one-off constant preprocessing (4 x 64-bit limbs, Montgomery, R = 2^256)
namespace {
typedef unsigned __int128 u128;
struct Fr {
  uint64_t v[4];
};
static const uint64_t FR_P[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const uint64_t FR_PINV = 0xfffffffeffffffffull;   // -p^-1 mod 2^64
static bool fr_geq_p(const uint64_t* a) {
  for (int i = 3; i >= 0; i--) { if (a[i] != FR_P[i]) return a[i] > FR_P[i]; }
  return true;
}
static void fr_sub_p(uint64_t* a) {
  u128 bo = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - FR_P[i] - bo; a[i] = (uint64_t)t; bo = (t >> 64) & 1; }
}
static Fr fr_add(const Fr& a, const Fr& b) {
  Fr r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
  if (c || fr_geq_p(r.v)) fr_sub_p(r.v);
  return r;
}
static Fr fr_sub(const Fr& a, const Fr& b) {
  Fr r;
  u128 bo = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a.v[i] - b.v[i] - bo; r.v[i] = (uint64_t)t; bo = (t >> 64) & 1; }
  if (bo) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + FR_P[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
  return r;
}
static Fr fr_mul(const Fr& a, const Fr& b) {   // a b / R mod p
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a.v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * FR_PINV;
    c = (u128)m * FR_P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * FR_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  Fr r;
  memcpy(r.v, t, 32);
  if (t[4] || fr_geq_p(r.v)) fr_sub_p(r.v);
  return r;
}
static Fr fr_raw(uint64_t x) { Fr r = {{x, 0, 0, 0}}; return r; }
static Fr FR_R2;   // R^2 mod p, computed once
static void fr_init() {
  static bool done = false;
  if (done) return;
  // R mod p by 256 doublings of 1, then R^2 = mont_mul-free: 256 more doublings of R
  Fr x = fr_raw(1);
  for (int i = 0; i < 512; i++) x = fr_add(x, x);
  FR_R2 = x;
  done = true;
}
static Fr fr_to_mont(const Fr& a) { return fr_mul(a, FR_R2); }
static Fr fr_from_mont(const Fr& a) { return fr_mul(a, fr_raw(1)); }
static Fr fr_from_u64(uint64_t x) { return fr_to_mont(fr_raw(x)); }
static Fr fr_from_le(const uint8_t b[32]) { Fr r; memcpy(r.v, b, 32); while (fr_geq_p(r.v)) fr_sub_p(r.v); return fr_to_mont(r); }
static void fr_to_le(const Fr& a, uint8_t out[32]) { Fr r = fr_from_mont(a); memcpy(out, r.v, 32); }
static bool fr_is_zero(const Fr& a) { return !(a.v[0] | a.v[1] | a.v[2] | a.v[3]); }
static Fr fr_inv(const Fr& a) {   // a^(p-2)
  uint64_t e[4];
  memcpy(e, FR_P, 32);
  e[0] -= 2;
  Fr r = fr_from_u64(1);
  for (int i = 3; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      r = fr_mul(r, r);
      if ((e[i] >> b) & 1) r = fr_mul(r, a);
    }
  return r;
}
typedef std::vector<std::vector<Fr>> Mat;
static Mat mat_mul(const Mat& a, const Mat& b) {
  const size_t t = a.size();
  Mat c(t, std::vector<Fr>(t, fr_raw(0)));
  for (size_t i = 0; i < t; i++)
    for (size_t j = 0; j < t; j++) {
      Fr acc = fr_raw(0);
      for (size_t k = 0; k < t; k++) acc = fr_add(acc, fr_mul(a[i][k], b[k][j]));
      c[i][j] = acc;
    }
  return c;
}
static bool mat_inv(Mat a, Mat& out) {   // Gauss-Jordan
  const size_t n = a.size();
  out.assign(n, std::vector<Fr>(n, fr_raw(0)));
  for (size_t i = 0; i < n; i++) out[i][i] = fr_from_u64(1);
  for (size_t col = 0; col < n; col++) {
    size_t piv = col;
    while (piv < n && fr_is_zero(a[piv][col])) piv++;
    if (piv == n) return false;
    std::swap(a[col], a[piv]);
    std::swap(out[col], out[piv]);
    const Fr inv = fr_inv(a[col][col]);
    for (size_t j = 0; j < n; j++) { a[col][j] = fr_mul(a[col][j], inv); out[col][j] = fr_mul(out[col][j], inv); }
    for (size_t r = 0; r < n; r++) {
      if (r == col || fr_is_zero(a[r][col])) continue;
      const Fr f = a[r][col];
      for (size_t j = 0; j < n; j++) {
        a[r][j] = fr_sub(a[r][j], fr_mul(f, a[col][j]));
        out[r][j] = fr_sub(out[r][j], fr_mul(f, out[col][j]));
      }
    }
  }
  return true;
}
}  // namespace

// Constants of width t in the order the kernel reads them (canonical 32-byte LE; the device converts to Montgomery):
//   [half t] first-half round constants | [t t] MDS | [t t] pre-sparse matrix | R_P x (c0, row0[t], col0[t-1]) |
//   [half t] second-half round constants
// The optimised form (constants pushed forward, sparse partial-round matrices) evaluates the SAME permutation as
// round = add constants, S-box, MDS -- oracle/py/poseidon.py: optimized_params / permute_optimized is the big-integer
// statement of this function and tests/test_poseidon_oracle.py checks the two against each other.
static bool poseidon_constants(int t, int mds_mode, std::vector<uint8_t>& out) {
  fr_init();
  const int rp = P_RP, half = P_RF / 2, nr = P_RF + P_RP;
  Grain g(255, t, P_RF, P_RP);
  std::vector<std::vector<Fr>> c(nr, std::vector<Fr>(t));
  for (int r = 0; r < nr; r++)
    for (int i = 0; i < t; i++) { uint8_t le[32]; g.element(le); c[r][i] = fr_from_le(le); }
  Mat mds(t, std::vector<Fr>(t));
  if (mds_mode == 0) {   // Cauchy 1 / (i + t + j): Filecoin neptune's matrix
    for (int i = 0; i < t; i++)
      for (int j = 0; j < t; j++) mds[i][j] = fr_inv(fr_from_u64((uint64_t)(i + t + j)));
  } else {               // x_i, y_j from the Grain stream behind the constants (Poseidon reference generator)
    for (;;) {
      std::vector<Fr> xy(2 * t);
      for (auto& e : xy) { uint8_t le[32]; g.element_mod(le); e = fr_from_le(le); }
      bool ok = true;
      for (int i = 0; i < 2 * t && ok; i++)
        for (int j = i + 1; j < 2 * t && ok; j++) ok = memcmp(xy[i].v, xy[j].v, 32) != 0;
      for (int i = 0; i < t && ok; i++)
        for (int j = 0; j < t && ok; j++) ok = !fr_is_zero(fr_add(xy[i], xy[t + j]));
      if (!ok) continue;
      for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) mds[i][j] = fr_inv(fr_add(xy[i], xy[t + j]));
      break;
    }
  }
  // constants of the partial rounds pushed forward through the MDS
  std::vector<Fr> a(rp);
  std::vector<Fr> k = c[half];
  for (int r = 0; r < rp; r++) {
    a[r] = k[0];
    std::vector<Fr> nxt = c[half + r + 1];
    for (int i = 0; i < t; i++)
      for (int j = 1; j < t; j++) nxt[i] = fr_add(nxt[i], fr_mul(mds[i][j], k[j]));
    k = nxt;
  }
  // sparse factorisation, last partial round first
  std::vector<std::vector<Fr>> row0(rp), col0(rp);
  Mat d = mds;
  for (int it = 0; it < rp; it++) {
    Mat dh(t - 1, std::vector<Fr>(t - 1)), dhi;
    for (int i = 1; i < t; i++)
      for (int j = 1; j < t; j++) dh[i - 1][j - 1] = d[i][j];
    if (!mat_inv(dh, dhi)) return false;
    const int r = rp - 1 - it;
    row0[r].assign(t, fr_raw(0));
    row0[r][0] = d[0][0];
    for (int j = 0; j < t - 1; j++) {
      Fr acc = fr_raw(0);
      for (int q = 0; q < t - 1; q++) acc = fr_add(acc, fr_mul(d[0][1 + q], dhi[q][j]));
      row0[r][1 + j] = acc;
    }
    col0[r].resize(t - 1);
    for (int i = 1; i < t; i++) col0[r][i - 1] = d[i][0];
    Mat p(t, std::vector<Fr>(t, fr_raw(0)));
    p[0][0] = fr_from_u64(1);
    for (int i = 1; i < t; i++)
      for (int j = 1; j < t; j++) p[i][j] = dh[i - 1][j - 1];
    d = mat_mul(p, mds);
  }
  out.clear();
  auto put = [&](const Fr& x) { uint8_t le[32]; fr_to_le(x, le); out.insert(out.end(), le, le + 32); };
  for (int r = 0; r < half; r++) for (int i = 0; i < t; i++) put(c[r][i]);
  for (int i = 0; i < t; i++) for (int j = 0; j < t; j++) put(mds[i][j]);
  for (int i = 0; i < t; i++) for (int j = 0; j < t; j++) put(d[i][j]);
  for (int r = 0; r < rp; r++) {
    put(a[r]);
    for (int j = 0; j < t; j++) put(row0[r][j]);
    for (int j = 0; j < t - 1; j++) put(col0[r][j]);
  }
  for (int i = 0; i < t; i++) put(k[i]);
  for (int r = 1; r < half; r++) for (int i = 0; i < t; i++) put(c[half + rp + r][i]);
  return true;
}

// host-only (no device needed): the optimised constants as canonical bytes, for the CPU test that compares them with
// the oracle's big-integer derivation
extern "C" int32_t bz_poseidon_optimized_constants(int32_t t, int32_t mds_mode, uint8_t* out, size_t out_cap, size_t* n_bytes) {
  if (!n_bytes) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  if ((t != 3 && t != 9 && t != 12) || (mds_mode != 0 && mds_mode != 1)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "width must be 3, 9 or 12; mds_mode 0 or 1");
  std::vector<uint8_t> v;
  if (!poseidon_constants(t, mds_mode, v)) return bz_fail(BZ_ERR_UNKNOWN, "singular sub-matrix while factoring the MDS");
  *n_bytes = v.size();
  if (out) {
    if (out_cap < v.size()) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "buffer too small (%zu < %zu)", out_cap, v.size());
    memcpy(out, v.data(), v.size());
  }
  return BZ_OK;
}

struct PoseidonParams {
  uint4* consts = nullptr;   // Montgomery form, layout of poseidon_constants()
  bool ready = false;
};

struct bz_poseidon {
  bz_dclient* dc = nullptr;
  bool initialized = false;
  uint32_t height = 0;
  int tree_mode = 0;
  int in_arity = 11;
  PoseidonParams par[3][2];   // [width index: 0 -> t = 3, 1 -> t = 9, 2 -> t = 12][mds mode]
  // element stream
  std::vector<uint8_t> staged;      // host staging of elements not yet on the device (32 B each)
  uint64_t elems_total = 0;         // elements received since initialize (ring id = elems_total - 1)
  uint64_t elems_tree = 0;          // elements of the current tree
  uint64_t elems_on_device = 0;
  uint8_t* d_inputs = nullptr;      // base-layer inputs of the current tree
  uint64_t d_inputs_cap = 0;
  std::vector<uint8_t*> d_layer;    // node digests per layer
  std::vector<uint64_t> layer_size, done;
  std::vector<uint8_t> pending;     // 64-byte records, FIFO: [pending_head, pending.size() / 64)
  size_t pending_head = 0;
  uint32_t last_hash_id = 0;
  cudaEvent_t ev_layer[10][3] = {};  // per layer: hash begin, hash end, digests copied to the host
  uint8_t* stage_host = nullptr;    // pinned staging of one flush's digests
  size_t stage_cap = 0;
  double device_ms = 0;             // kernel time since initialize (CUDA events around every hash launch)
  std::mutex mu;
};

static int width_index(int t) { return t == 3 ? 0 : (t == 9 ? 1 : 2); }

static int32_t ensure_params(bz_poseidon* p, int t, int mds_mode = 0) {
  PoseidonParams& q = p->par[width_index(t)][mds_mode];
  if (q.ready) return BZ_OK;
  std::vector<uint8_t> host;
  if (!poseidon_constants(t, mds_mode, host)) return bz_fail(BZ_ERR_UNKNOWN, "singular sub-matrix while factoring the MDS");
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&q.consts, host.size()));
  cudaStream_t st = dc_stream(p->dc);
  CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(q.consts, host.data(), host.size(), cudaMemcpyHostToDevice, st));
  poseidon_prepare(q.consts, (int)(host.size() / 32), st);
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
  for (auto& l : p->ev_layer) for (auto& e : l) cudaEventCreate(&e);
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
  for (auto& w : p->par) for (auto& q : w) if (q.consts) cudaFree(q.consts);
  for (auto& l : p->ev_layer) for (auto& e : l) if (e) cudaEventDestroy(e);
  if (p->stage_host) cudaFreeHost(p->stage_host);
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
  p->pending_head = 0;
  p->device_ms = 0;
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
  // Every layer that has new complete nodes is hashed now.  The launches go out back to back (layer l reads what layer
  // l-1 wrote: stream order), each followed by the copy of its digests into pinned staging; the host then builds the
  // 64-byte records of layer l as soon as ITS copy has landed, i.e. while the layers above are still being hashed.
  struct Batch { uint32_t l; uint64_t first, n; size_t stage_off; };
  std::vector<Batch> batches;
  uint64_t avail_prev = 0;
  size_t stage_need = 0;
  for (uint32_t l = 0; l < p->height; l++) {
    const uint64_t avail = l == 0 ? p->elems_on_device / p->in_arity : avail_prev / 8;
    avail_prev = avail;
    if (avail <= p->done[l]) break;
    batches.push_back({l, p->done[l], avail - p->done[l], stage_need});
    stage_need += (size_t)(avail - p->done[l]) * 32;
  }
  if (batches.empty()) return BZ_OK;
  if (p->stage_cap < stage_need) {
    if (p->stage_host) cudaFreeHost(p->stage_host);
    p->stage_host = nullptr;
    p->stage_cap = 0;
    CUDA_TRY(BZ_ERR_READ, cudaHostAlloc((void**)&p->stage_host, stage_need, cudaHostAllocDefault));
    p->stage_cap = stage_need;
  }
  for (size_t b = 0; b < batches.size(); b++) {
    const Batch& B = batches[b];
    const uint32_t l = B.l;
    int t = (l == 0 && p->in_arity == 11) ? 12 : 9;
    const PoseidonParams& q = p->par[width_index(t)][0];
    const uint8_t* in = l == 0 ? p->d_inputs + B.first * (uint64_t)p->in_arity * 32 : p->d_layer[l - 1] + B.first * 8 * 32;
    uint8_t* out = p->d_layer[l] + B.first * 32;
    cudaEventRecord(p->ev_layer[l][0], st);
    poseidon_hash(t, (const uint4*)in, B.n, q.consts, P_RF, P_RP, 0, (uint4*)out, st);
    cudaEventRecord(p->ev_layer[l][1], st);
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaGetLastError());
    CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(p->stage_host + B.stage_off, out, B.n * 32, cudaMemcpyDeviceToHost, st));
    cudaEventRecord(p->ev_layer[l][2], st);
  }
  if (p->pending_head && p->pending_head * 64 == p->pending.size()) { p->pending.clear(); p->pending_head = 0; }
  for (const Batch& B : batches) {
    CUDA_TRY(BZ_ERR_READ, cudaEventSynchronize(p->ev_layer[B.l][2]));
    float ms = 0;
    cudaEventElapsedTime(&ms, p->ev_layer[B.l][0], p->ev_layer[B.l][1]);
    p->device_ms += ms;
    const size_t at = p->pending.size();
    p->pending.resize(at + B.n * 64);
    const uint8_t* src = p->stage_host + B.stage_off;
    for (uint64_t i = 0; i < B.n; i++) {
      uint8_t* rec = &p->pending[at + i * 64];
      memcpy(rec, src + i * 32, 32);
      memset(rec + 32, 0, 32);
      uint64_t id = B.first + i;
      uint64_t meta = (id & 0x3fffffffull) | ((uint64_t)B.l << 30);   // poseidon_api.rs:50-61
      memcpy(rec + 32, &meta, 8);
    }
    p->done[B.l] = B.first + B.n;
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
  size_t e = 0;
  while (e < n_el) {
    if (p->elems_tree == p->d_inputs_cap) {   // previous tree's inputs are full: finish it first
      rc = flush(p);
      if (rc) return rc;
      if (p->elems_tree == p->d_inputs_cap) return bz_fail(BZ_ERR_WRITE, "input ring full");
    }
    if (len <= 32) {
      uint8_t el[32];
      memset(el, 0, 32);
      memcpy(el, input, len);
      p->staged.insert(p->staged.end(), el, el + 32);
      e = 1;
      p->elems_total++;
      p->elems_tree++;
    } else {   // bulk: as many whole elements as the current tree still takes
      const size_t take = (size_t)std::min<uint64_t>(n_el - e, p->d_inputs_cap - p->elems_tree);
      if (take >= 4096) {
        // large burst: straight from the caller's buffer to the device, behind whatever single elements are staged
        cudaStream_t st = dc_stream(p->dc);
        if (!p->staged.empty()) {
          CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(p->d_inputs + p->elems_on_device * 32, p->staged.data(), p->staged.size(), cudaMemcpyHostToDevice, st));
          CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(st));
          p->elems_on_device += p->staged.size() / 32;
          p->staged.clear();
        }
        CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(p->d_inputs + p->elems_on_device * 32, input + e * 32, take * 32, cudaMemcpyHostToDevice, st));
        CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(st));   // the caller may reuse `input` on return
        p->elems_on_device += take;
      } else {
        p->staged.insert(p->staged.end(), input + e * 32, input + (e + take) * 32);
      }
      e += take;
      p->elems_total += take;
      p->elems_tree += take;
    }
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
  *n = (uint32_t)(p->pending.size() / 64 - p->pending_head);
  return BZ_OK;
}

extern "C" int32_t bz_poseidon_get_raw_results(bz_poseidon* p, uint32_t num_of_results, uint8_t* out) {
  if (!p || (!out && num_of_results)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(p->mu);
  const size_t have = p->pending.size() / 64 - p->pending_head;
  if (num_of_results > have) return bz_fail(BZ_ERR_READ, "only %zu results pending", have);
  if (num_of_results) {
    const uint8_t* src = &p->pending[p->pending_head * 64];
    memcpy(out, src, (size_t)num_of_results * 64);
    uint32_t id;
    memcpy(&id, src + (size_t)(num_of_results - 1) * 64 + 32, 4);
    p->last_hash_id = id & 0x3fffffffu;
    p->pending_head += num_of_results;
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

// ---- B200 additions
// kernel milliseconds since initialize() (CUDA events around every hash launch): the device-side cost of the tree
extern "C" int32_t bz_poseidon_device_ms(bz_poseidon* p, float* ms) {
  if (!p || !ms) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(p->mu);
  *ms = (float)p->device_ms;
  return BZ_OK;
}

// n raw permutations of width t (3, 9 or 12) on full states (n x t canonical elements in and out), with the Cauchy
// (mds_mode 0: the PoseidonClient instances) or the Grain-sampled MDS (1: the parameter set of the published
// Poseidon reference vectors): lets tests run known-answer vectors and random states through the CUDA kernel
extern "C" int32_t bz_poseidon_permute(bz_poseidon* p, int32_t t, int32_t mds_mode, const uint8_t* in, size_t n, uint8_t* out) {
  if (!p || !in || !out || !n) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad argument");
  if ((t != 3 && t != 9 && t != 12) || (mds_mode != 0 && mds_mode != 1)) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "width must be 3, 9 or 12; mds_mode 0 or 1");
  int32_t rc = dc_select(p->dc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(p->mu);
  rc = ensure_params(p, t, mds_mode);
  if (rc) return rc;
  const size_t bytes = n * (size_t)t * 32;
  uint8_t* d = nullptr;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&d, 2 * bytes));
  cudaStream_t st = dc_stream(p->dc);
  cudaMemcpyAsync(d, in, bytes, cudaMemcpyHostToDevice, st);
  poseidon_hash(t, (const uint4*)d, n, p->par[width_index(t)][mds_mode].consts, P_RF, P_RP, 1, (uint4*)(d + bytes), st);
  cudaMemcpyAsync(out, d + bytes, bytes, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return bz_fail(BZ_ERR_UNKNOWN, "permutation failed: %s", cudaGetErrorString(e));
  return BZ_OK;
}
