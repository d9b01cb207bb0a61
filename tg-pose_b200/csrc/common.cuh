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

// true the first time it is called with the current device for this guard (cudaFuncSetAttribute is per device)
inline bool first_on_device(std::atomic<unsigned long long>& seen) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (seen.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
}

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

// MIXED tensor-core operand (tgp_gemm_args.mixed): a row is 4*Kp 16-bit slots,
//   [fp16(x) x Kp | bf16(x) x Kp | bf16(x - fp16(x)) x Kp | unused x Kp].
// fp16(x) saturates to +-65504 instead of overflowing, the residual then carries the rest (reduced precision, never inf);
// below the fp16 normal range the residual keeps the relative accuracy.
__device__ __forceinline__ uint16_t mixed_hi16(float v, float& hi) {
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    asm("cvt.f32.f16 %0, %1;" : "=f"(hi) : "h"(h));
    return h;
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
    uint16_t h;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
}
// one value -> slot `col` of the row starting at `row16`
// nb: leave the bf16(x) slot unwritten -- the operand of a contraction whose WEIGHTS carry their residual in fp16
// (tgp_gemm_args.mixed == 2: the third pass is fp16(a).fp16lo(b) and never reads bf16(a))
__device__ __forceinline__ void mixed_store1(uint16_t* row16, int Kp, int col, float v, bool nb = false) {
    float hi;
    row16[col] = mixed_hi16(v, hi);
    if (!nb) row16[Kp + col] = bf16_bits(v);
    row16[2 * Kp + col] = bf16_bits(v - hi);
}
// two values -> packed fp16x2 (saturating; x in the low half) and their fp32 read-back
__device__ __forceinline__ uint32_t mixed_hi16x2(float x, float y, float& hx, float& hy) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(hx), "=f"(hy) : "r"(d));
    return d;
}
__device__ __forceinline__ uint32_t f16x2_bits(float x, float y) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_bits(float x, float y) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
    return d;
}
// four consecutive values (col % 4 == 0, row 16-byte aligned): three 8-byte stores, packed conversions
__device__ __forceinline__ void mixed_store4(uint16_t* row16, int Kp, int col, float4 v) {
    float h0, h1, h2, h3;
    const uint32_t a01 = mixed_hi16x2(v.x, v.y, h0, h1), a23 = mixed_hi16x2(v.z, v.w, h2, h3);
    *reinterpret_cast<uint2*>(row16 + col) = make_uint2(a01, a23);
    *reinterpret_cast<uint2*>(row16 + Kp + col) = make_uint2(bf16x2_bits(v.x, v.y), bf16x2_bits(v.z, v.w));
    *reinterpret_cast<uint2*>(row16 + 2 * Kp + col) = make_uint2(bf16x2_bits(v.x - h0, v.y - h1), bf16x2_bits(v.z - h2, v.w - h3));
}
// the same with streaming (evict-first) stores: a GEMM epilogue that writes an operand larger than L2 must not push the
// weight tiles, which every row block re-reads, out of the cache
__device__ __forceinline__ void st_cs_u2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void mixed_store4_cs(uint16_t* row16, int Kp, int col, float4 v, bool nb = false) {
    float h0, h1, h2, h3;
    const uint32_t a01 = mixed_hi16x2(v.x, v.y, h0, h1), a23 = mixed_hi16x2(v.z, v.w, h2, h3);
    st_cs_u2(row16 + col, a01, a23);
    if (!nb) st_cs_u2(row16 + Kp + col, bf16x2_bits(v.x, v.y), bf16x2_bits(v.z, v.w));
    st_cs_u2(row16 + 2 * Kp + col, bf16x2_bits(v.x - h0, v.y - h1), bf16x2_bits(v.z - h2, v.w - h3));
}

#define TGP_DISPATCH_IDX(bits, ...)                                   \
    do {                                                              \
        if ((bits) == 64) { using IdxT = int64_t; __VA_ARGS__; }      \
        else if ((bits) == 32) { using IdxT = int32_t; __VA_ARGS__; } \
        else return tgp::fail(TGP_EINVAL, "idx_bits must be 32 or 64"); \
    } while (0)

}  // namespace tgp
