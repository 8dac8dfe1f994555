// Helpers shared by the C-ABI translation units (dclient.cu, msm_client.cu, msm_api.cu, ntt_api.cu, poseidon_api.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/blaze_b200.h"

// records a thread-local message for bz_last_error() and returns `code`
int32_t bz_fail(int32_t code, const char* fmt, ...);

#define CUDA_TRY(code, expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return bz_fail((code), "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

int32_t dc_select(bz_dclient* dc);       // cudaSetDevice(dc's device); validates the handle
cudaStream_t dc_stream(bz_dclient* dc);  // the client's work stream
int dc_device(bz_dclient* dc);
