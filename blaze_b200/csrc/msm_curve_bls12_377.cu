// Bls12_377 instantiation of the MSM back end (see msm_curve.cuh).
#include "msm_curve.cuh"
#include "msm_ba.cuh"

namespace bz {
template <>
const uint32_t* CurveLaunch<Bls12_377>::fr_mod_host() { return FR377_MOD_H; }
const CurveOps* curve_ops_bls12_377() { return CurveLaunch<Bls12_377>::ops(); }
}  // namespace bz
