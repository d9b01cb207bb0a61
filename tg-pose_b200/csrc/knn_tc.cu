// knn_tc.cu -- feature-space kNN (RF-F, gcn3d.py:14-23 as used by :201-206) with the pairwise inner products
// on the 5th-gen tensor cores and the top-(k+1) selection by warp shuffles, fused in one persistent kernel.
//
// The reference computes the (B,N,N) matrix with torch.bmm (a dense contraction of depth D = 128/256), adds the
// norms in two more passes and calls torch.topk; nothing here ever reaches HBM except the (B,N,k) indices.
//
//   unit  = (cloud b, tile of <=128 queries); persistent CTAs stride over the units
//   warp 0      TMA producer: query tile and candidate tile of the [tf32(x) | x - tf32(x)] operand
//               (the same split operand the projection GEMM reads), 128B swizzle, 4-stage mbarrier ring
//   warp 1      tcgen05.mma kind::tf32 128x128x8, three passes lo.hi + hi.lo + hi.hi (3xTF32: fp32-level inner
//               products, error << the reference formula's own rounding bound, SURVEY 8c rule 3) into a
//               double-buffered TMEM accumulator: the MMA of candidate tile t+1 overlaps the selection of tile t
//   warps 2-17  tcgen05.ld -> d = ((inner * -2) + q_j) + q_i (gcn3d.py:20, same rounding order) -> 128x128 distance
//               tile in shared memory -> each warp owns 8 queries (two at a time, interleaved) whose sorted (distance, index) lists live in
//               registers across all candidate tiles (knn_select.cuh), ascending (distance, index).
// Padding: rows of a tile that belong to the next cloud / lie past the end are masked through q_j = +inf.
#include "knn_select.cuh"
#include "tc_common.cuh"

namespace tgp {

constexpr int KT_STAGES = 4;
constexpr int KT_BN = 128;
constexpr int KT_EPI_WARPS = 16;
constexpr int KT_THREADS = 64 + 32 * KT_EPI_WARPS;
constexpr int KT_QPW = TC_BM / KT_EPI_WARPS;            // queries per selection warp (16)
constexpr int KT_LDD = KT_BN + 4;                       // distance tile pitch (floats)
constexpr int KT_STAGE_BYTES = TC_A_BYTES + KT_BN * TC_BK * 4;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * KT_EPI_WARPS) : "memory"); }

__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC,
              const float* __restrict__ qn, int N, int Kp, int k, int q_tiles, int q_step, int num_units,
              int64_t* __restrict__ idx64, int32_t* __restrict__ idx32) {
    extern __shared__ __align__(1024) unsigned char kt_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)kt_smem + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + KT_STAGES * KT_STAGE_BYTES);
    uint64_t* empty = full + KT_STAGES;
    uint64_t* tmem_full = empty + KT_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* Ds = reinterpret_cast<float*>(base + KT_STAGES * KT_STAGE_BYTES + 256);        // [128][KT_LDD]
    float* qns = Ds + TC_BM * KT_LDD;                                                       // [2][128] candidate norms, per tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = Kp / TC_BK;
    const int c_tiles = (N + KT_BN - 1) / KT_BN;
    const int K = k + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KT_STAGES; ++s) { tc_mbar_init(full + s, 1); tc_mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { tc_mbar_init(tmem_full + s, 1); tc_mbar_init(tmem_empty + s, KT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int b = u / q_tiles, qt = u - b * q_tiles;
                const int qrow = b * N + qt * q_step;
                for (int ct = 0; ct < c_tiles; ++ct) {
                    const int crow = b * N + ct * KT_BN;
                    for (int seg = 0; seg < 3; ++seg) {
                        const int a_off = (seg == 0) ? Kp : 0, b_off = (seg == 1) ? Kp : 0;
                        for (int kb = 0; kb < kblocks; ++kb) {
                            tc_mbar_wait(empty + stage, phase ^ 1);
                            unsigned char* sa = base + stage * KT_STAGE_BYTES;
                            tc_mbar_expect_tx(full + stage, KT_STAGE_BYTES);
                            tma_load_2d(sa, &tmQ, a_off + kb * TC_BK, qrow, full + stage);
                            tma_load_2d(sa + TC_A_BYTES, &tmC, b_off + kb * TC_BK, crow, full + stage);
                            if (++stage == KT_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                for (int ct = 0; ct < c_tiles; ++ct, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (it >> 1) & 1;
                    tc_mbar_wait(tmem_empty + acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * KT_BN;
                    uint32_t accum = 0;
                    for (int kb = 0; kb < 3 * kblocks; ++kb) {
                        tc_mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t sa = s_u32(base + stage * KT_STAGE_BYTES);
                        const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 8; ++kk) {
                            umma_tf32(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accum);
                            accum = 1;
                        }
                        umma_commit(empty + stage);
                        if (++stage == KT_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(tmem_full + acc);
                }
            }
        }
    } else {
        // ===================== distance tile + selection (warps 2..9) =====================
        const int e = warp - 2;                    // selection warp 0..7: queries e, e+8, ...
        const int quarter = warp & 3;              // TMEM lane quarter this warp may read
        const int chunk = e >> 2;                  // which 32-column chunk of the accumulator it converts (4 warps / quarter)
        const int et = threadIdx.x - 64;           // 0..255
        const int my_row = quarter * 32 + lane;    // query row of this thread in the TMEM layout
        int it = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int b = u / q_tiles, qt = u - b * q_tiles;
            const int q0 = qt * q_step;
            const int nq = min(q_step, N - q0);    // valid queries of this unit
            const float* qb = qn + (size_t)b * N;
            // candidate norms of tile 0, +inf beyond N (masks the padding columns of the last tile); the norms of
            // tile t+1 are fetched into the other buffer before tile t's closing barrier
            // (every warp passed the last tile's closing barrier, so nobody still reads the previous unit's norms)
            if (et < KT_BN) qns[(it & 1) * KT_BN + et] = et < N ? __ldg(qb + et) : CUDART_INF_F;
            const float qi = my_row < nq ? __ldg(qb + q0 + my_row) : 0.f;
            float td[KT_QPW];
            int ti[KT_QPW];
            float th[KT_QPW];
#pragma unroll
            for (int i = 0; i < KT_QPW; ++i) { td[i] = CUDART_INF_F; ti[i] = -1; th[i] = CUDART_INF_F; }
            epi_bar_sync();
            for (int ct = 0; ct < c_tiles; ++ct, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                const int c0 = ct * KT_BN;
                tc_mbar_wait(tmem_full + acc, acc_phase);
                tc_fence_after();
                {
                    const int cc = chunk * 32;
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * KT_BN + cc);
                    TMEM_LD_32x32(taddr, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float* drow = Ds + my_row * KT_LDD + cc;
                    const float4* qj4 = reinterpret_cast<const float4*>(qns + (it & 1) * KT_BN + cc);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 qj = qj4[j >> 2];
                        float4 o;
                        o.x = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[j]), -2.0f), qj.x), qi);
                        o.y = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 1]), -2.0f), qj.y), qi);
                        o.z = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 2]), -2.0f), qj.z), qi);
                        o.w = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 3]), -2.0f), qj.w), qi);
                        *reinterpret_cast<float4*>(drow + j) = o;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(tmem_empty + acc);
                epi_bar_sync();                    // distance tile complete
#pragma unroll
                for (int i = 0; i < KT_QPW; i += 2) {
                    const int qa = e + KT_EPI_WARPS * i, qb2 = qa + KT_EPI_WARPS;
                    if (qa < nq) {
                        // rows past nq hold other clouds' queries: harmless, their lists are never stored
                        WarpTopPair pr;
                        pr.dA = td[i]; pr.iA = ti[i]; pr.thA = th[i];
                        pr.dB = td[i + 1]; pr.iB = ti[i + 1]; pr.thB = th[i + 1];
                        const float* ra = Ds + qa * KT_LDD;
                        const float* rb = Ds + qb2 * KT_LDD;
#pragma unroll
                        for (int j0 = 0; j0 < KT_BN; j0 += 32) pr.admit2(ra[j0 + lane], rb[j0 + lane], c0 + j0, lane, K);
                        td[i] = pr.dA; ti[i] = pr.iA; th[i] = pr.thA;
                        td[i + 1] = pr.dB; ti[i + 1] = pr.iB; th[i + 1] = pr.thB;
                    }
                }
                if (et < KT_BN && ct + 1 < c_tiles) {
                    const int j = c0 + KT_BN + et;
                    qns[((it + 1) & 1) * KT_BN + et] = j < N ? __ldg(qb + j) : CUDART_INF_F;
                }
                epi_bar_sync();                    // selection done: Ds may be overwritten
            }
#pragma unroll
            for (int i = 0; i < KT_QPW; ++i) {
                const int ql = e + KT_EPI_WARPS * i;
                if (ql < nq && lane >= 1 && lane <= k) {
                    const size_t o = ((size_t)b * N + q0 + ql) * k + lane - 1;
                    // (a rank can only stay unfilled if a distance was NaN; never hand -1 to the gathers)
                    const int v = ti[i] < 0 ? 0 : ti[i];
                    if (idx64) idx64[o] = v;
                    if (idx32) idx32[o] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
    }
}

}  // namespace tgp

using namespace tgp;

// true when the tensor-core kernel covers this shape (otherwise the fp32 FMA kernel in knn.cu runs)
// (must not depend on B: a cloud's result may not change with the batch it sits in)
bool tgp_knn_tc_eligible(int B, int N, int D, int k) {
    (void)B;
    return k + 1 <= 32 && N >= 64 && D >= 16;
}

int tgp_knn_tc(const float* x_split, const float* qn, int B, int N, int D, int k, int64_t* idx64, int32_t* idx32,
               cudaStream_t st) {
    const int Kp = tgp_split_kpad(D);
    CUtensorMap tmQ, tmC;
    int rc = tgp_make_map(&tmQ, x_split, (long)B * N, Kp, TC_BM);
    if (rc) return rc;
    rc = tgp_make_map(&tmC, x_split, (long)B * N, Kp, KT_BN);
    if (rc) return rc;
    const int q_tiles = (N + TC_BM - 1) / TC_BM;
    const int q_step = (N + q_tiles - 1) / q_tiles;          // balanced query tiles (1028 -> 9 x 115, not 8 x 128 + 4)
    const int num_units = B * q_tiles;
    const size_t smem = (size_t)KT_STAGES * KT_STAGE_BYTES + 1024 + 256 + sizeof(float) * (TC_BM * KT_LDD + 2 * KT_BN);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    if (smem > 227 * 1024) return fail(TGP_EINVAL, "tgp_knn_feat: N too large for the tensor-core kernel");
    const int grid = num_units < TGP_NUM_SMS ? num_units : TGP_NUM_SMS;
    knn_tc_kernel<<<grid, KT_THREADS, smem, st>>>(tmQ, tmC, qn, N, Kp, k, q_tiles, q_step, num_units, idx64, idx32);
    return check_launch("knn_tc_kernel");
}
