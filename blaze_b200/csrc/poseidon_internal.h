// Internal (non-ABI) declarations shared by poseidon.cu and poseidon_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace bz {
// n canonical 32-byte constants, converted in place to Montgomery form
void poseidon_prepare(uint4* consts, int n, cudaStream_t st);
// n_hashes independent hashes of (t-1) canonical elements each -> canonical digests (raw = 0), or n_hashes bare
// permutations of t-cell states (raw = 1); consts: the layout of poseidon_constants() (poseidon_api.cu)
void poseidon_hash(int t, const uint4* in, uint64_t n_hashes, const uint4* consts, int r_f, int r_p, int raw, uint4* out,
                   cudaStream_t st);
}  // namespace bz
