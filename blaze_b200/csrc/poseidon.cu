// Poseidon (x^5) over BLS12-381 Fr: many independent hashes, one thread per hash, sm_100a.
//
// Black box being replaced: the FPGA Poseidon core behind PoseidonClient
// (/root/reference/src/ingo_hash/poseidon_api.rs:96-146): streaming hasher + 8-ary Merkle tree.
// Parameter set (the reference's constants CSV is not in the repository; the permutation itself is pinned by the
// published Poseidon reference vector, tests/golden/external_kats.json): width t = arity+1 in {9, 12}, R_F = 8,
// R_P = 57, Grain-LFSR round constants, Cauchy MDS (Filecoin neptune), state = [2^arity - 1, inputs...],
// output = state[1]; see oracle/py/poseidon.py for the same definition in big integers.
//
// Layout: constants in global memory in Montgomery form (every thread of a warp reads the same
// address: one broadcast transaction per load); the state lives in registers (t x 8 limbs).
// field products as real calls: keeps the hot loops inside the 32 KB instruction cache (measured: ff.cuh)
#ifndef BZ_INLINE_MUL_TU   // A/B switch (scripts/build_variant.sh): inlined products measured slower here, profiles/r2_inline_call_ab.txt
#define BZ_NOINLINE_MUL 1
#endif
#include <cuda_runtime.h>

#include <cstdint>

#include "ff.cuh"
#include "msm_internal.h"
#include "poseidon_internal.h"

namespace bz {

#ifndef POSEIDON_COOP_BELOW
#define POSEIDON_COOP_BELOW 16384   // hashes per launch below which the sixteen-lane kernel is used
#endif

typedef ff<Fr381> PF;
typedef Fe<Fr381> PE;

__device__ __forceinline__ PE p_ld(const uint4* p) {
  uint4 a = __ldg(p), b = __ldg(p + 1);
  PE r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void p_st(uint4* p, const PE& r) {
  p[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  p[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ PE p_sbox(const PE& x) {
  PE x2 = PF::sqr(x);
  PE x4 = PF::sqr(x2);
  return PF::mul(x4, x);
}

// canonical constants -> Montgomery form, in place
__global__ void k_poseidon_prepare(uint4* consts, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PE v = p_ld(consts + 2 * i);
  p_st(consts + 2 * i, PF::to_mont(v));
}

// One thread per hash, state in registers, the OPTIMISED evaluation of the permutation (oracle/py/poseidon.py:
// permute_optimized is the same computation in big integers): R_F dense rounds + one pre-sparse matrix, and R_P
// partial rounds that add one constant, pass cell 0 through the S-box and multiply by a sparse matrix with 2t-1
// non-trivial entries (t products for the new cell 0, one product-accumulate for each other cell) instead of t^2.
// Constants: [half T] rc | [T T] mds | [T T] pre-sparse | R_P x (c0, row0[T], col0[T-1]) | [half T] rc, Montgomery form,
// read with the same address by every lane (broadcast).
//   raw = 0: in = n x (T-1) canonical elements; state = [2^(T-1) - 1, in...]; out = n canonical digests (cell 1)
//   raw = 1: in / out = n x T canonical state cells (the bare permutation: known-answer tests)
template <int T>
__device__ __forceinline__ void p_dense(PE (&s)[T], const uint4* __restrict__ m) {
  PE n[T];
#pragma unroll
  for (int i = 0; i < T; i++) {
    PE acc = PF::mul(p_ld(m + 2 * (i * T)), s[0]);
#pragma unroll
    for (int j = 1; j < T; j++) acc = PF::add(acc, PF::mul(p_ld(m + 2 * (i * T + j)), s[j]));
    n[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < T; i++) s[i] = n[i];
}

template <int T>
__global__ void __launch_bounds__(128) k_poseidon_hash(const uint4* __restrict__ in, uint64_t n_hashes,
                                                       const uint4* __restrict__ consts, int r_f, int r_p, int raw,
                                                       uint4* __restrict__ out) {
  uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hashes) return;
  constexpr int ARITY = T - 1;
  const int half = r_f / 2;
  const uint4* rc1 = consts;
  const uint4* mds = rc1 + 2 * (half * T);
  const uint4* pre = mds + 2 * (T * T);
  const uint4* part = pre + 2 * (T * T);
  const uint4* rc2 = part + 2 * ((size_t)r_p * 2 * T);
  PE s[T];
  if (raw) {
#pragma unroll
    for (int i = 0; i < T; i++) s[i] = PF::to_mont(p_ld(in + 2 * (h * T + i)));
  } else {
    PE tag = PF::zero();
    tag.v[0] = (1u << ARITY) - 1;
    s[0] = PF::to_mont(tag);
#pragma unroll
    for (int i = 0; i < ARITY; i++) s[i + 1] = PF::to_mont(p_ld(in + 2 * (h * ARITY + i)));
  }
  for (int r = 0; r < half; r++) {
#pragma unroll
    for (int i = 0; i < T; i++) s[i] = p_sbox(PF::add(s[i], p_ld(rc1 + 2 * (r * T + i))));
    p_dense<T>(s, r == half - 1 ? pre : mds);
  }
  for (int r = 0; r < r_p; r++) {
    const uint4* q = part + 2 * ((size_t)r * 2 * T);
    const PE x0 = p_sbox(PF::add(s[0], p_ld(q)));
    PE n0 = PF::mul(p_ld(q + 2), x0);
#pragma unroll
    for (int j = 1; j < T; j++) n0 = PF::add(n0, PF::mul(p_ld(q + 2 * (1 + j)), s[j]));
#pragma unroll
    for (int i = 1; i < T; i++) s[i] = PF::add(s[i], PF::mul(p_ld(q + 2 * (T + i)), x0));
    s[0] = n0;
  }
  for (int r = 0; r < half; r++) {
#pragma unroll
    for (int i = 0; i < T; i++) s[i] = p_sbox(PF::add(s[i], p_ld(rc2 + 2 * (r * T + i))));
    p_dense<T>(s, mds);
  }
  if (raw) {
#pragma unroll
    for (int i = 0; i < T; i++) p_st(out + 2 * (h * T + i), PF::from_mont(s[i]));
  } else {
    p_st(out + 2 * h, PF::from_mont(s[1]));
  }
}

// The same permutation by SIXTEEN lanes per hash (lane i < T holds state cell i), for layers too small to fill the
// machine with one thread per hash (the upper layers of a tree: 8^k hashes): a dense round is the S-box on every lane and
// one output row per lane (3 + T product latencies instead of 3 T + T^2), a partial round is the S-box on lane 0, one
// product per lane for the new cell 0 (summed by a shuffle tree) and one product-accumulate per lane for the other
// cells (~5.5 product latencies instead of 2 T + 2).  ~5x lower latency per hash, ~2x more issued work: latency-bound
// layers only.
__device__ __forceinline__ PE p_bcast(const PE& v, int src, unsigned mask) {
  PE r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(mask, v.v[k], src, 16);
  return r;
}
__device__ __forceinline__ PE p_xor(const PE& v, int m, unsigned mask) {
  PE r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = __shfl_xor_sync(mask, v.v[k], m, 16);
  return r;
}

template <int T>
__global__ void __launch_bounds__(128) k_poseidon_hash_coop(const uint4* __restrict__ in, uint64_t n_hashes,
                                                            const uint4* __restrict__ consts, int r_f, int r_p, int raw,
                                                            uint4* __restrict__ out) {
  static_assert(T <= 16, "one state cell per lane of a half warp");
  constexpr int ARITY = T - 1;
  const int l16 = threadIdx.x & 15;
  const unsigned mask = 0xFFFFu << (threadIdx.x & 16);
  const uint64_t h = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const bool live = h < n_hashes;            // uniform within the half warp
  const bool cell = l16 < T;
  const int i = cell ? l16 : 0;              // idle lanes shadow cell 0; their results are never used
  const int half = r_f / 2;
  const uint4* rc1 = consts;
  const uint4* mds = rc1 + 2 * (half * T);
  const uint4* pre = mds + 2 * (T * T);
  const uint4* part = pre + 2 * (T * T);
  const uint4* rc2 = part + 2 * ((size_t)r_p * 2 * T);
  PE s = PF::zero();
  if (live && cell) {
    if (raw) {
      s = PF::to_mont(p_ld(in + 2 * (h * T + i)));
    } else if (i == 0) {
      s.v[0] = (1u << ARITY) - 1;
      s = PF::to_mont(s);
    } else {
      s = PF::to_mont(p_ld(in + 2 * (h * ARITY + i - 1)));
    }
  }
  auto dense = [&](const uint4* m) {
    PE acc = PF::mul(p_ld(m + 2 * (i * T)), p_bcast(s, 0, mask));
#pragma unroll 1
    for (int j = 1; j < T; j++) acc = PF::add(acc, PF::mul(p_ld(m + 2 * (i * T + j)), p_bcast(s, j, mask)));
    s = acc;
  };
  for (int r = 0; r < half; r++) {
    s = p_sbox(PF::add(s, p_ld(rc1 + 2 * (r * T + i))));
    dense(r == half - 1 ? pre : mds);
  }
#pragma unroll 1
  for (int r = 0; r < r_p; r++) {
    const uint4* q = part + 2 * ((size_t)r * 2 * T);
    const PE x0 = p_bcast(p_sbox(PF::add(s, p_ld(q))), 0, mask);       // lane 0's value is the real one
    PE p = PF::mul(p_ld(q + 2 * (1 + i)), l16 == 0 ? x0 : s);          // row0[i] * (new cell 0 | cell i)
    if (!cell) p = PF::zero();
    const PE snew = PF::add(s, PF::mul(p_ld(q + 2 * (T + i)), x0));    // cell i + col0[i-1] * x0   (unused on lane 0)
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) p = PF::add(p, p_xor(p, m, mask));
    s = l16 == 0 ? p : snew;
  }
  for (int r = 0; r < half; r++) {
    s = p_sbox(PF::add(s, p_ld(rc2 + 2 * (r * T + i))));
    dense(mds);
  }
  if (!live || !cell) return;
  if (raw) p_st(out + 2 * (h * T + i), PF::from_mont(s));
  else if (i == 1) p_st(out + 2 * h, PF::from_mont(s));
}

void poseidon_prepare(uint4* consts, int n, cudaStream_t st) {
  k_poseidon_prepare<<<(n + 63) / 64, 64, 0, st>>>(consts, n);
  g_kernel_launches += 1;
}

void poseidon_hash(int t, const uint4* in, uint64_t n_hashes, const uint4* consts, int r_f, int r_p, int raw, uint4* out,
                   cudaStream_t st) {
  if (!n_hashes) return;
  if (n_hashes < POSEIDON_COOP_BELOW) {   // latency-bound layer: sixteen lanes per hash
    unsigned cb = (unsigned)((n_hashes * 16 + 127) / 128);
    if (t == 3) k_poseidon_hash_coop<3><<<cb, 128, 0, st>>>(in, n_hashes, consts, r_f, r_p, raw, out);
    else if (t == 9) k_poseidon_hash_coop<9><<<cb, 128, 0, st>>>(in, n_hashes, consts, r_f, r_p, raw, out);
    else k_poseidon_hash_coop<12><<<cb, 128, 0, st>>>(in, n_hashes, consts, r_f, r_p, raw, out);
    g_kernel_launches += 1;
    return;
  }
  unsigned blocks = (unsigned)((n_hashes + 127) / 128);
  if (t == 3) k_poseidon_hash<3><<<blocks, 128, 0, st>>>(in, n_hashes, consts, r_f, r_p, raw, out);
  else if (t == 9) k_poseidon_hash<9><<<blocks, 128, 0, st>>>(in, n_hashes, consts, r_f, r_p, raw, out);
  else k_poseidon_hash<12><<<blocks, 128, 0, st>>>(in, n_hashes, consts, r_f, r_p, raw, out);
  g_kernel_launches += 1;
}

}  // namespace bz
