// Shared helpers for libmdt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mdt_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmdt_b200 is written for sm_100a only"
#endif

namespace mdt {

extern unsigned long long g_launch_count;  // defined in capi.cu

inline int launch_status() {
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MDT_OK : (int)e;
}

inline int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// The opt-in dynamic shared-memory size is a PER-DEVICE function attribute: a process that drives several GPUs must set it on each.
// `done` is the call site's table (one per kernel).
constexpr int kMaxDevices = 64;
template <typename K>
inline bool ensure_smem_attr(K kernel, int bytes, bool (&done)[kMaxDevices]) {
    int dev = 0;
    const bool tracked = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices;
    if (tracked && done[dev]) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    if (tracked) done[dev] = true;
    return true;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace mdt
