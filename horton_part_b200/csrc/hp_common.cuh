// Shared device/host helpers for the hp_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "hp_b200.h"

#ifndef __CUDA_ARCH__
#define HP_HOST_ONLY 1
#endif

namespace hp {

constexpr int kWarp = 32;
constexpr double kFourPi = 12.566370614359172;   // 4*pi rounded to nearest, == numpy's 4*np.pi
constexpr double kEightPi = 25.132741228718345;  // 8*pi, == numpy's 8*np.pi

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t err, const char* what);

#define HP_REQUIRE(cond, msg)                           \
    do {                                                \
        if (!(cond)) {                                  \
            ::hp::set_error("%s: %s", __func__, (msg)); \
            return HP_ERR_ARG;                          \
        }                                               \
    } while (0)

#define HP_LAUNCH_CHECK(what)                                            \
    do {                                                                 \
        int _rc = ::hp::check_cuda(cudaGetLastError(), (what));          \
        if (_rc != HP_OK) return _rc;                                    \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// Function attributes (dynamic shared-memory limits) are per device: `flags` is a per-kernel-family array
// indexed by the current device; returns true the first time it is called for that device.
bool first_use_on_device(bool (&flags)[64]);

// Butterfly sum: every lane ends with the same value, summation order fixed by lane id.
__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Block-wide sum for blockDim.x a multiple of 32 (<= 1024); result valid in thread 0.
// `scratch` needs 32 doubles of shared memory.  Fixed order => bit-reproducible.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_allsum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double total = 0.0;
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        total = lane < nw ? scratch[lane] : 0.0;
        total = warp_allsum(total);
    }
    return total;
}

}  // namespace hp
