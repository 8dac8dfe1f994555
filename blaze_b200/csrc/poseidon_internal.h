// Internal (non-ABI) declarations shared by poseidon.cu and poseidon_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace bz {
// rc: n_rc canonical 32-byte constants (converted in place to Montgomery); mds: t*t elements (computed)
void poseidon_prepare(uint4* rc, int n_rc, uint4* mds, int t, cudaStream_t st);
// n_hashes independent hashes of (t-1) canonical elements each -> canonical digests
void poseidon_hash(int t, const uint4* in, uint64_t n_hashes, const uint4* rc, const uint4* mds, int r_f, int r_p,
                   uint4* out, cudaStream_t st);
}  // namespace bz
