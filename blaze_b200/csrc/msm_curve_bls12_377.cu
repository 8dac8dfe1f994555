// Bls12_377 instantiation of the MSM back end (see msm_curve.cuh).
// field products as real calls: keeps the hot loops inside the 32 KB instruction cache (measured: ff.cuh)
#ifndef BZ_INLINE_MUL_TU   // A/B switch (scripts/build_variant.sh): inlined products measured slower here, profiles/r2_inline_call_ab.txt
#define BZ_NOINLINE_MUL 1
#endif
#include "msm_curve.cuh"
#include "msm_ba.cuh"

namespace bz {
template <>
const uint32_t* CurveLaunch<Bls12_377>::fr_mod_host() { return FR377_MOD_H; }
const CurveOps* curve_ops_bls12_377() { return CurveLaunch<Bls12_377>::ops(); }
}  // namespace bz
