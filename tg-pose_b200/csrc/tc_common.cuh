// tc_common.cuh -- tcgen05 / TMEM / TMA / mbarrier helpers shared by the tensor-core kernels (gemm_tc.cu, knn_tc.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace tgp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // 32 fp32 = one 128-byte swizzle span
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(s_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 27)) __trap();   // a broken pipeline must fault, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(s_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
// one lane of a converged warp (the MMA issuer runs its loop with all 32 lanes so that descriptors and loop state stay
// warp-uniform -- uniform registers, no per-instruction election / broadcast -- and elects a lane only for the issue)
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                  // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset
    d |= (uint64_t)1 << 46;                  // version
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}

// MN-major, 128B-swizzled operand tile (the contraction index is the SLOW one: a row-major (rows = contraction, cols = M or N)
// matrix read in place): rows of 128 B = 64 consecutive 16-bit M/N elements for ONE contraction index; 8 consecutive
// contraction indices form a 1024-byte swizzle atom; the next group of 8 lies SBO = 1024 B further, the next 64 M/N
// elements LBO bytes further (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset: between 64-element blocks along M/N
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: between groups of 8 contraction indices
    d |= (uint64_t)1 << 46;                             // version
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}


#define TMEM_LD_32x32(taddr, r)                                                                         \
    asm volatile(                                                                                       \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                       \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                       \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"       \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),          \
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),        \
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),        \
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                             \
        : "r"(taddr))

// wait for the outstanding tcgen05.ld of `r`: the registers are in-out operands, so no value read before the wait can be
// used after it (the loads complete asynchronously; other work may be scheduled between the load and this wait)
#define TMEM_WAIT_LD(r)                                                                                 \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                       \
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),   \
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),          \
          "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),        \
          "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),        \
          "+r"(r[29]), "+r"(r[30]), "+r"(r[31])                                                             \
        :: "memory")

#define TMEM_ST_32x32(taddr, r)                                                                         \
    asm volatile(                                                                                       \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                 \
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                      \
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"              \
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),   \
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),         \
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),       \
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])        \
        : "memory")

// D[tmem] (+)= A[tmem] * B[smem]: the A tile (M = 128 rows on the 128 lanes, one 32-bit column per tf32 element of K)
// is read from tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

}  // namespace tgp

// ---- host side: tensor maps over a split operand (rows, 2*Kp) fp32, box = (TC_BK columns, box_rows rows), 128B swizzle
typedef CUresult (*tgp_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline tgp_encode_fn tgp_get_encode() {
    static tgp_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tgp_encode_fn>(p);
    }
    return fn;
}

static inline int tgp_make_map(CUtensorMap* tm, const float* ptr, long rows, long Kp, int box_rows) {
    tgp_encode_fn enc = tgp_get_encode();
    if (!enc) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled unavailable");
    cuuint64_t gdim[2] = {(cuuint64_t)(2 * Kp), (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)(2 * Kp) * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)tgp::TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled failed");
    return TGP_OK;
}


// bf16 view of a MIXED operand (rows, 8*Kp bytes per row): box = (64 bf16 = 128 bytes, box_rows rows), 128B swizzle.
// Columns [2Kp, 3Kp) hold bf16(x), [3Kp, 4Kp) hold bf16(x - tf32(x)).
static inline int tgp_make_map_bf16(CUtensorMap* tm, const float* ptr, long rows, long Kp, int box_rows) {
    tgp_encode_fn enc = tgp_get_encode();
    if (!enc) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled unavailable");
    cuuint64_t gdim[2] = {(cuuint64_t)(4 * Kp), (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)(8 * Kp)};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled (bf16) failed");
    return TGP_OK;
}

// 16-bit view of a K-BLOCKED transposed mixed operand (tgp_split_mixed_t): `total_rows` dense rows of 64 slots (128 bytes).
static inline int tgp_make_map_bf16_blocked(CUtensorMap* tm, const float* ptr, long total_rows, int box_rows) {
    tgp_encode_fn enc = tgp_get_encode();
    if (!enc) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled unavailable");
    cuuint64_t gdim[2] = {64u, (cuuint64_t)total_rows};
    cuuint64_t gstride[1] = {128u};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled (blocked bf16) failed");
    return TGP_OK;
}

// fp32 view of a K-BLOCKED transposed [tf32 | residual] operand (tgp_split_tf32_t): dense rows of 32 floats (128 bytes).
static inline int tgp_make_map_f32_blocked(CUtensorMap* tm, const float* ptr, long total_rows, int box_rows) {
    tgp_encode_fn enc = tgp_get_encode();
    if (!enc) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled unavailable");
    cuuint64_t gdim[2] = {32u, (cuuint64_t)total_rows};
    cuuint64_t gstride[1] = {128u};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tgp::fail(TGP_EINVAL, "tgp_gemm: cuTensorMapEncodeTiled (blocked fp32) failed");
    return TGP_OK;
}
