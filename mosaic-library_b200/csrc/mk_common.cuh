// mk_common.cuh -- shared host/device helpers for libmosaic_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mosaic_b200.h"

namespace mk {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MK_ERR_CUDA;
}

#define MK_CUDA_TRY(expr)                                         \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return mk::cuda_fail(_e, #expr);   \
    } while (0)

#define MK_LAUNCH_CHECK(name)                                     \
    do {                                                          \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return mk::cuda_fail(_e, name);    \
        mk::g_launches.fetch_add(1, std::memory_order_relaxed);   \
    } while (0)

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// 3x3 inverse exactly as cv::invert does for 3x3 double matrices (closed-form adjugate, no FMA
// contraction: this translation unit's host code is built with -ffp-contract=off).
// cv::warpPerspective inverts H unless WARP_INVERSE_MAP is set (mosaic.py:114 does not set it).
bool invert3x3(const double *S, double *D);

// SMs available to work launched on `stream` (its green-context partition, mk_sm_partition_create, or the whole device)
int stream_sm_count(cudaStream_t stream);

// Tensor-core kNN flags, written on the DEVICE (prep pass + TC kernel) and read by the guarded SIMT fallback:
//   flags[0] == flags[1]  some value is not an integer in [0,255]: bf16 operands would round it, the TC kernel does nothing
//                  (flags[1] = this call's epoch, flags[0] is set to it by the prep pass: nothing to zero before the call);
//   flags[3] != 0  some row's best-two squared distance reached 2^22: the TC path ranks by the integer d^2, which equals
//                  cv2's ranking by sqrtf((float)d^2) only while sqrtf is injective (d^2 < 2^22), so the whole call is
//                  recomputed by the bit-exact SIMT kernel.  (Never happens for uint8 rows <= 64 bytes or SIFT.)
#ifdef __CUDACC__
__device__ __forceinline__ bool knn_tc_inputs_ok(const int *flags) { return flags[0] != flags[1]; }
__device__ __forceinline__ bool knn_tc_result_ok(const int *flags) { return flags[0] != flags[1] && flags[3] == 0; }
#endif

}  // namespace mk
