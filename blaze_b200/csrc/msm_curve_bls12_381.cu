// Bls12_381 instantiation of the MSM back end (see msm_curve.cuh).
#include "msm_curve.cuh"
#include "msm_ba.cuh"

namespace bz {
template <>
const uint32_t* CurveLaunch<Bls12_381>::fr_mod_host() { return FR381_MOD_H; }
const CurveOps* curve_ops_bls12_381() { return CurveLaunch<Bls12_381>::ops(); }
}  // namespace bz
