// Record layouts of the MSM back end in HBM (shared by the kernels and by the host-side check vehicle).
#pragma once
#include <cstdint>

#include "ec.cuh"

namespace bz {

template <class C>
struct alignas(16) AffineM {   // packed affine point (lists that are read sequentially)
  uint32_t x[C::Fq::N], y[C::Fq::N];
};
// Entry of the resident Montgomery point table, which is only ever GATHERED: 96-byte records are padded to
// 128 B so that one gather touches exactly one 128-byte DRAM line (unpadded they straddle two lines half
// of the time: measured 197 B of DRAM traffic per 96-B record, profiles/r1_traffic.json).
template <class C>
struct alignas((sizeof(uint32_t) * 2 * C::Fq::N == 96) ? 128 : 16) AffineT {
  uint32_t x[C::Fq::N], y[C::Fq::N];
};
template <class C>
struct alignas(16) XyzzM {
  uint32_t X[C::Fq::N], Y[C::Fq::N], ZZ[C::Fq::N], ZZZ[C::Fq::N];
};

}  // namespace bz
