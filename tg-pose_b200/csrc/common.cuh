// common.cuh -- shared helpers for the sm_100a kernels of libtgpose_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/tgpose_b200.h"

#ifndef TGP_NUM_SMS
#define TGP_NUM_SMS 148  // B200: 2 dies x 74 SMs
#endif

namespace tgp {

extern thread_local char g_err[256];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return TGP_OK;
}

inline cudaStream_t as_stream(tgp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// index tensors arrive as int64 (torch.topk, gcn3d.py:21) or int32 (internal)
template <typename IdxT>
__device__ __forceinline__ int ld_idx(const IdxT* p, size_t i) {
    return (int)__ldg(p + i);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// F.normalize(v, dim=-1) with eps 1e-12 (gcn3d.py:54): v / max(||v||, eps); zero stays zero
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
    float n = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
    float inv = 1.0f / fmaxf(n, 1e-12f);
    x *= inv; y *= inv; z *= inv;
}

#define TGP_DISPATCH_IDX(bits, ...)                                   \
    do {                                                              \
        if ((bits) == 64) { using IdxT = int64_t; __VA_ARGS__; }      \
        else if ((bits) == 32) { using IdxT = int32_t; __VA_ARGS__; } \
        else return tgp::fail(TGP_EINVAL, "idx_bits must be 32 or 64"); \
    } while (0)

}  // namespace tgp
