// Poseidon (x^5) over BLS12-381 Fr: many independent hashes, one thread per hash, sm_100a.
//
// Black box being replaced: the FPGA Poseidon core behind PoseidonClient
// (/root/reference/src/ingo_hash/poseidon_api.rs:96-146): streaming hasher + 8-ary Merkle tree.
// Parameter set (the reference's constants CSV is not in the repository -- parity UNPINNED):
// width t = arity+1 in {9, 12}, R_F = 8, R_P = 57, Grain-LFSR round constants, Cauchy MDS,
// state = [2^arity - 1, inputs...], output = state[1]; see oracle/py/poseidon.py for the same
// definition in big integers.
//
// Layout: constants in global memory in Montgomery form (every thread of a warp reads the same
// address: one broadcast transaction per load); the state lives in registers (t x 8 limbs).
// field products as real calls: keeps the hot loops inside the 32 KB instruction cache (measured: ff.cuh)
#define BZ_NOINLINE_MUL 1
#include <cuda_runtime.h>

#include <cstdint>

#include "ff.cuh"
#include "msm_internal.h"
#include "poseidon_internal.h"

namespace bz {

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

// canonical constants -> Montgomery; mds[i][j] = 1/(i + t + j)
__global__ void k_poseidon_prepare(uint4* rc, int n_rc, uint4* mds, int t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rc) {
    PE v = p_ld(rc + 2 * i);
    p_st(rc + 2 * i, PF::to_mont(v));
  } else if (i < n_rc + t * t) {
    int k = i - n_rc;
    int row = k / t, col = k % t;
    PE v = PF::zero();
    v.v[0] = (uint32_t)(row + t + col);
    p_st(mds + 2 * k, PF::inv(PF::to_mont(v)));
  }
}

// in: n_hashes x arity canonical elements (32 B LE); out: n_hashes canonical digests
template <int T>
__global__ void __launch_bounds__(128) k_poseidon_hash(const uint4* __restrict__ in, uint64_t n_hashes,
                                                       const uint4* __restrict__ rc, const uint4* __restrict__ mds,
                                                       int r_f, int r_p, uint4* __restrict__ out) {
  uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hashes) return;
  constexpr int ARITY = T - 1;
  PE s[T];
  {
    PE tag = PF::zero();
    tag.v[0] = (1u << ARITY) - 1;
    s[0] = PF::to_mont(tag);
  }
#pragma unroll
  for (int i = 0; i < ARITY; i++) s[i + 1] = PF::to_mont(p_ld(in + 2 * (h * ARITY + i)));
  const int nr = r_f + r_p;
  for (int r = 0; r < nr; r++) {
#pragma unroll
    for (int i = 0; i < T; i++) s[i] = PF::add(s[i], p_ld(rc + 2 * (r * T + i)));
    bool full = r < r_f / 2 || r >= r_f / 2 + r_p;
    s[0] = p_sbox(s[0]);
    if (full) {
#pragma unroll
      for (int i = 1; i < T; i++) s[i] = p_sbox(s[i]);
    }
    PE n[T];
#pragma unroll
    for (int i = 0; i < T; i++) {
      PE acc = PF::mul(p_ld(mds + 2 * (i * T)), s[0]);
#pragma unroll
      for (int j = 1; j < T; j++) acc = PF::add(acc, PF::mul(p_ld(mds + 2 * (i * T + j)), s[j]));
      n[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < T; i++) s[i] = n[i];
  }
  p_st(out + 2 * h, PF::from_mont(s[1]));
}

void poseidon_prepare(uint4* rc, int n_rc, uint4* mds, int t, cudaStream_t st) {
  int n = n_rc + t * t;
  k_poseidon_prepare<<<(n + 63) / 64, 64, 0, st>>>(rc, n_rc, mds, t);
  g_kernel_launches += 1;
}

void poseidon_hash(int t, const uint4* in, uint64_t n_hashes, const uint4* rc, const uint4* mds, int r_f, int r_p,
                   uint4* out, cudaStream_t st) {
  if (!n_hashes) return;
  unsigned blocks = (unsigned)((n_hashes + 127) / 128);
  if (t == 9) k_poseidon_hash<9><<<blocks, 128, 0, st>>>(in, n_hashes, rc, mds, r_f, r_p, out);
  else k_poseidon_hash<12><<<blocks, 128, 0, st>>>(in, n_hashes, rc, mds, r_f, r_p, out);
  g_kernel_launches += 1;
}

}  // namespace bz
