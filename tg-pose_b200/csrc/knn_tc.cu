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
#include <stdlib.h>

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
              const int* __restrict__ unit_list, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32) {
    // unit_list != nullptr: fix-up launch behind knn_tc2_kernel -- only the units it listed ([0] = count) are redone
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
    if (unit_list) {
        num_units = min(__ldg(unit_list), num_units);
        if (num_units == 0) return;
        ++unit_list;
    }

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
            for (int ui = blockIdx.x; ui < num_units; ui += gridDim.x) {
                const int u = unit_list ? __ldg(unit_list + ui) : ui;
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
        for (int ui = blockIdx.x; ui < num_units; ui += gridDim.x) {
            const int u = unit_list ? __ldg(unit_list + ui) : ui;
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


// ------------------------------------------------------------------------------------------------------------
// Threshold-selection variant (knn_select.cuh): the unit's distance tiles are produced TWICE by the tensor cores.
//   pass 1  every selection thread owns one query row x 32 candidate columns of the TMEM accumulator, turns them into
//           distances in registers and folds them into 64 group minima of its query (groups of GS consecutive
//           candidates, group g -> slot g % 64; each slot is only ever touched by one thread: no atomics, no barriers
//           between tiles); then one warp per query takes the K-th smallest of the 64 minima: T >= K-th distance.
//   pass 2  the same tiles again; a candidate with d <= T (~K + 4 per query) is appended to the query's buffer with one
//           shared-memory atomic; finally every survivor ranks itself against the others.
// ~7x fewer selection instructions than the insertion lists of knn_tc_kernel at 2x its (cheap) tensor work; no distance
// tile in shared memory.  A query with more than KT2_CAP survivors (massive ties) or fewer than K (non-finite input)
// sends its unit to the fix-up list, which knn_tc_kernel redoes right behind this launch.
// NS group-minimum slots per query: 64 for K = k + 1 <= 32, 128 for K <= 64 (expected survivors -NS ln(1 - K/NS):
// 42 at K = 31 / NS = 64; 65 at K = 51, 89 at K = 64 / NS = 128).  The bigger survivor buffers of NS = 128 are paid for
// with a 2-stage operand ring instead of 4 (the selection, not the feed, bounds this kernel).
template <int NS> struct Kt2 {
    static constexpr int CAP = NS == 64 ? 80 : 144;      // survivor buffer entries per query
    static constexpr int GM_LD = NS + 1;                 // pitch of the group-minimum rows (bank-conflict free)
    static constexpr int STAGES = NS == 64 ? KT_STAGES : 2;
    static constexpr int RANK_SLOTS = (CAP + 31) / 32;
};

// 32 accumulator columns of one query row -> distances -> pass 0: group minima (groups of GS consecutive candidates, slot =
// group % 64, exclusive to this thread); pass 1: survivors (d <= T) appended with one predicated shared-memory atomic each
template <int GS, int NS>
__device__ __forceinline__ void kt2_consume(const uint32_t (&r)[32], const float* wqs, float qi, int pass, int c0, float* gmr,
                                            float T, uint32_t cnt_addr, uint32_t buf_row) {
    const float4* qj4 = reinterpret_cast<const float4*>(wqs);
#pragma unroll
    for (int h = 0; h < 32; h += 16) {                    // 16 columns at a time: bounded register pressure
        float d[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 qj = qj4[(h + j) >> 2];
            d[j] = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[h + j]), -2.0f), qj.x), qi);
            d[j + 1] = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[h + j + 1]), -2.0f), qj.y), qi);
            d[j + 2] = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[h + j + 2]), -2.0f), qj.z), qi);
            d[j + 3] = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(r[h + j + 3]), -2.0f), qj.w), qi);
        }
        if (pass == 0) {
#pragma unroll
            for (int w = 1; w < GS; w <<= 1)
#pragma unroll
                for (int j = 0; j < 16; j += 2 * w) d[j] = fminf(d[j], d[j + w]);
            const int g0 = (c0 + h) / GS;
#pragma unroll
            for (int g = 0; g < 16 / GS; ++g) {
                float* slot = gmr + ((g0 + g) & (NS - 1));
                if (d[g * GS] < *slot) *slot = d[g * GS];  // (+inf padding never writes: GS = 1 shares slots across chunks only for N <= 64)
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                asm volatile(
                    "{\n\t.reg .pred p, q;\n\t.reg .u32 pos, addr;\n\t"
                    "setp.le.f32 p, %0, %1;\n\t"
                    "@p atom.shared.add.u32 pos, [%2], 1;\n\t"
                    "setp.lt.and.u32 q, pos, %3, p;\n\t"
                    "mad.lo.u32 addr, pos, 8, %4;\n\t"
                    "@q st.shared.v2.b32 [addr], {%5, %6};\n\t}"
                    :: "f"(d[j]), "f"(T), "r"(cnt_addr), "n"(Kt2<NS>::CAP), "r"(buf_row), "r"(__float_as_uint(d[j])), "r"(c0 + h + j)
                    : "memory");
            }
        }
    }
}

// ATM (D <= 128): the unit's query tile [tf32 | residual] is written ONCE into the 256 tensor-memory columns the
// accumulators leave free and every MMA reads its A operand from there; only candidate tiles stream through shared
// memory (hi and lo k-blocks of a stage loaded once and used by all three products).  3x less L2 -> SM traffic than
// re-fetching the query tile for every candidate tile and segment -- that traffic, not the tensor pipe, bounded the
// streaming variant (ncu: tensor pipe 34 % busy, long-scoreboard stalls).
template <int GS, bool ATM, int NS>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC,
               const float* __restrict__ x_split, const float* __restrict__ qn, int N, int Kp, int k, int q_tiles,
               int q_step, int num_units, int* __restrict__ fix_list, int64_t* __restrict__ idx64,
               int32_t* __restrict__ idx32) {
    extern __shared__ __align__(1024) unsigned char kt_smem[];
    constexpr int KT2_CAP = Kt2<NS>::CAP, KT2_GM_LD = Kt2<NS>::GM_LD, STG = Kt2<NS>::STAGES;
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)kt_smem + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + STG * KT_STAGE_BYTES);
    uint64_t* empty = full + STG;
    uint64_t* tmem_full = empty + STG;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* a_ready = tmem_empty + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 1);
    int* unit_flag = reinterpret_cast<int*>(tmem_ptr + 1);
    constexpr int TMEM_COLS = ATM ? 512 : 256;
    constexpr uint32_t A_COL = 256;                                                 // first tensor-memory column of the query tile
    unsigned char* sel = base + STG * KT_STAGE_BYTES + 256;
    uint2* buf = reinterpret_cast<uint2*>(sel);                                     // [128][KT2_CAP] survivors (pass 2)
    float* gm = reinterpret_cast<float*>(sel);                                      // [128][KT2_GM_LD] group minima (pass 1), same bytes
    int* cnt = reinterpret_cast<int*>(sel + TC_BM * KT2_CAP * 8);                   // [128]
    float* Ts = reinterpret_cast<float*>(cnt + TC_BM);                              // [128]
    float* wq = Ts + TC_BM;                                                         // [KT_EPI_WARPS][32] candidate norms of a chunk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = Kp / TC_BK;
    const int c_tiles = (N + KT_BN - 1) / KT_BN;
    const int K = k + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; ++s) { tc_mbar_init(full + s, 1); tc_mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { tc_mbar_init(tmem_full + s, 1); tc_mbar_init(tmem_empty + s, KT_EPI_WARPS); }
        tc_mbar_init(a_ready, KT_EPI_WARPS);
        *unit_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer: both passes load the same tiles =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int b = u / q_tiles, qt = u - b * q_tiles;
                const int qrow = b * N + qt * q_step;
                for (int ct2 = 0; ct2 < 2 * c_tiles; ++ct2) {
                    const int ct = ct2 < c_tiles ? ct2 : ct2 - c_tiles;
                    const int crow = b * N + ct * KT_BN;
                    if (ATM) {
                        // a stage = the hi and the lo k-block of the candidate tile
                        for (int kb = 0; kb < kblocks; ++kb) {
                            tc_mbar_wait(empty + stage, phase ^ 1);
                            unsigned char* sa = base + stage * KT_STAGE_BYTES;
                            tc_mbar_expect_tx(full + stage, KT_STAGE_BYTES);
                            tma_load_2d(sa, &tmC, kb * TC_BK, crow, full + stage);
                            tma_load_2d(sa + TC_A_BYTES, &tmC, Kp + kb * TC_BK, crow, full + stage);
                            if (++stage == STG) { stage = 0; phase ^= 1; }
                        }
                        continue;
                    }
                    for (int seg = 0; seg < 3; ++seg) {
                        const int a_off = (seg == 0) ? Kp : 0, b_off = (seg == 1) ? Kp : 0;
                        for (int kb = 0; kb < kblocks; ++kb) {
                            tc_mbar_wait(empty + stage, phase ^ 1);
                            unsigned char* sa = base + stage * KT_STAGE_BYTES;
                            tc_mbar_expect_tx(full + stage, KT_STAGE_BYTES);
                            tma_load_2d(sa, &tmQ, a_off + kb * TC_BK, qrow, full + stage);
                            tma_load_2d(sa + TC_A_BYTES, &tmC, b_off + kb * TC_BK, crow, full + stage);
                            if (++stage == STG) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged; one elected lane issues) =====================
        {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            uint32_t a_phase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                if (ATM) {
                    tc_mbar_wait(a_ready, a_phase);        // the selection warps have written this unit's query tile
                    a_phase ^= 1;
                    tc_fence_after();
                }
                for (int ct2 = 0; ct2 < 2 * c_tiles; ++ct2, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (it >> 1) & 1;
                    tc_mbar_wait(tmem_empty + acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * KT_BN;
                    uint32_t accum = 0;
                    if (ATM) {
                        for (int kb = 0; kb < kblocks; ++kb) {
                            tc_mbar_wait(full + stage, phase);
                            tc_fence_after();
                            const uint32_t sb = s_u32(base + stage * KT_STAGE_BYTES);
                            const uint64_t bhi = make_sw128_desc(sb), blo = make_sw128_desc(sb + TC_A_BYTES);
                            const uint32_t ahi = tmem_base + A_COL + kb * TC_BK, alo = ahi + Kp;
                            if (tc_elect_one()) {
#pragma unroll
                                for (int kk = 0; kk < TC_BK / 8; ++kk) {
                                    umma_tf32_ts(d_tmem, alo + 8 * kk, bhi + (uint64_t)(2 * kk), idesc, kk ? 1u : accum);
                                    umma_tf32_ts(d_tmem, ahi + 8 * kk, blo + (uint64_t)(2 * kk), idesc, 1);
                                    umma_tf32_ts(d_tmem, ahi + 8 * kk, bhi + (uint64_t)(2 * kk), idesc, 1);
                                }
                                umma_commit(empty + stage);
                            }
                            __syncwarp();
                            accum = 1;
                            if (++stage == STG) { stage = 0; phase ^= 1; }
                        }
                        if (tc_elect_one()) umma_commit(tmem_full + acc);
                        __syncwarp();
                        continue;
                    }
                    for (int kb = 0; kb < 3 * kblocks; ++kb) {
                        tc_mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t sa = s_u32(base + stage * KT_STAGE_BYTES);
                        const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + TC_A_BYTES);
                        if (tc_elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk)
                                umma_tf32(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk ? 1u : accum);
                            umma_commit(empty + stage);
                        }
                        __syncwarp();
                        accum = 1;
                        if (++stage == STG) { stage = 0; phase ^= 1; }
                    }
                    if (tc_elect_one()) umma_commit(tmem_full + acc);
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== selection warps 2..17 =====================
        const int e = warp - 2;                    // selection warp: queries e, e+16, ... in the threshold / rank phases
        const int quarter = warp & 3;              // TMEM lane quarter this warp may read
        const int chunk = e >> 2;                  // 32-column chunk of the accumulator it owns
        const int et = threadIdx.x - 64;           // 0..511
        const int my_row = quarter * 32 + lane;    // query row of this thread in the TMEM layout
        const int cc = chunk * 32;
        float* wqs = wq + e * 32;
        float* gmr = gm + my_row * KT2_GM_LD;
        const uint32_t cnt_addr = s_u32(cnt + my_row);
        const uint32_t buf_row = s_u32(buf + my_row * KT2_CAP);
        int it = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int b = u / q_tiles, qt = u - b * q_tiles;
            const int q0 = qt * q_step;
            const int nq = min(q_step, N - q0);    // valid queries of this unit
            const float* qb = qn + (size_t)b * N;
            if (ATM) {
                // query tile -> tensor memory: this thread's row, columns [chunk * 2Kp/4, +2Kp/4) of [tf32 | residual].
                // (every MMA that read the previous unit's tile completed before its last tmem_full arrival was seen)
                const int per = (2 * Kp) / 4;                              // 64 at D = 128; a multiple of 16 (Kp % 32 == 0)
                const float4* src = reinterpret_cast<const float4*>(x_split + ((size_t)b * N + q0 + min(my_row, nq - 1)) * (2 * Kp) + chunk * per);
                for (int c = 0; c < per; c += 32) {
                    uint32_t r[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c + j < per) v = __ldg(src + ((c + j) >> 2));
                        r[j] = __float_as_uint(v.x); r[j + 1] = __float_as_uint(v.y);
                        r[j + 2] = __float_as_uint(v.z); r[j + 3] = __float_as_uint(v.w);
                    }
                    if (c + 32 <= per) {
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + A_COL + (uint32_t)(chunk * per + c);
                        TMEM_ST_32x32(taddr, r);
                    } else {
                        // 16-column tail (Kp = 32 or 96): two x8 stores would need another macro; Kp % 64 == 0 is required
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(a_ready);
            }
            for (int i = et; i < TC_BM * NS; i += 32 * KT_EPI_WARPS) gm[(i / NS) * KT2_GM_LD + (i % NS)] = CUDART_INF_F;
            const float qi = my_row < nq ? __ldg(qb + q0 + my_row) : 0.f;
            epi_bar_sync();
            float T = 0.f;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                for (int ct = 0; ct < c_tiles; ++ct, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (it >> 1) & 1;
                    const int c0 = ct * KT_BN + cc;                       // first candidate of this thread's chunk
                    // candidate norms of the chunk (+inf past N masks the padding columns), via the warp's own staging row
                    __syncwarp();
                    wqs[lane] = c0 + lane < N ? __ldg(qb + c0 + lane) : CUDART_INF_F;
                    __syncwarp();
                    tc_mbar_wait(tmem_full + acc, acc_phase);
                    tc_fence_after();
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * KT_BN + cc);
                    TMEM_LD_32x32(taddr, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(tmem_empty + acc);      // accumulator is in registers: the next tile may start
                    kt2_consume<GS, NS>(r, wqs, qi, pass, c0, gmr, T, cnt_addr, buf_row);
                }
                epi_bar_sync();
                if (pass == 0) {
                    // thresholds: K-th smallest of the NS group minima of each query
#pragma unroll 2
                    for (int i = 0; i < KT_QPW; ++i) {
                        const int ql = e + KT_EPI_WARPS * i;
                        const float* gq = gm + ql * KT2_GM_LD;
                        const float t = NS == 64 ? warp_kth_of_64(gq[lane], gq[32 + lane], K, lane)
                                                 : warp_kth_of_128(gq[lane], gq[32 + lane], gq[(64 + lane) % NS], gq[(96 + lane) % NS], K, lane);
                        if (lane == 0) Ts[ql] = fminf(t, 3.0e38f);
                    }
                    epi_bar_sync();                                        // group minima are dead: the survivor buffers take their place
                    T = my_row < nq ? Ts[my_row] : -CUDART_INF_F;
                    if (et < TC_BM) cnt[et] = 0;
                    epi_bar_sync();
                }
            }
            // ranks: (distance, index) ascending, rank 0 dropped
            bool bad = false;
            for (int i = 0; i < KT_QPW; ++i) {
                const int ql = e + KT_EPI_WARPS * i;
                if (ql >= nq) break;
                const int n = cnt[ql];
                if (n < K || n > KT2_CAP) { bad = true; continue; }
                uint2* qbuf = buf + ql * KT2_CAP;
#pragma unroll
                for (int s = 0; s < (KT2_CAP + 31) / 32; ++s)
                    if (s * 32 + lane < n) qbuf[s * 32 + lane].x = ordered_key(__uint_as_float(qbuf[s * 32 + lane].x));
                __syncwarp();
                const size_t o = ((size_t)b * N + q0 + ql) * k;
                warp_rank_store<(KT2_CAP + 31) / 32>(qbuf, n, k, lane, idx64 ? idx64 + o : nullptr, idx32 ? idx32 + o : nullptr);
            }
            if (bad && lane == 0) atomicOr(unit_flag, 1);
            epi_bar_sync();
            if (et == 0 && *unit_flag) {
                *unit_flag = 0;
                fix_list[1 + atomicAdd(fix_list, 1)] = u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

}  // namespace tgp

using namespace tgp;

// true when the tensor-core kernel covers this shape (otherwise the fp32 FMA kernel in knn.cu runs)
// (must not depend on B: a cloud's result may not change with the batch it sits in)
bool tgp_knn_tc_eligible(int B, int N, int D, int k) {
    (void)B;
    return k + 1 <= 64 && N >= 64 && D >= 16;
}

size_t tgp_knn_tc_fix_bytes(int B, int N) {
    const size_t units = (size_t)B * ((N + TC_BM - 1) / TC_BM);
    return ((units + 1) * sizeof(int) + 255) & ~(size_t)255;
}

// fix_list: tgp_knn_tc_fix_bytes(B, N) bytes of workspace ([0] = count, then unit ids), or nullptr to run the insertion-list
// kernel alone
int tgp_knn_tc(const float* x_split, const float* qn, int B, int N, int D, int k, int64_t* idx64, int32_t* idx32,
               int* fix_list, cudaStream_t st) {
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
    const bool big = k + 1 > 32;                             // 128 group minima, bigger survivor buffers, 2-stage ring
    const size_t smem2 = (size_t)(big ? Kt2<128>::STAGES : Kt2<64>::STAGES) * KT_STAGE_BYTES + 1024 + 256 +
                         (size_t)TC_BM * (big ? Kt2<128>::CAP : Kt2<64>::CAP) * 8 + TC_BM * 8 + KT_EPI_WARPS * 32 * 4;
    static std::atomic<unsigned long long> attr_set{0};   // one bit per device: function attributes are per device
    if (first_on_device(attr_set)) {
        cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
#define KT2_ATTR(GS, ATM, NS) cudaFuncSetAttribute(knn_tc2_kernel<GS, ATM, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
        KT2_ATTR(1, false, 64); KT2_ATTR(2, false, 64); KT2_ATTR(4, false, 64);
        KT2_ATTR(1, true, 64); KT2_ATTR(2, true, 64); KT2_ATTR(4, true, 64);
        KT2_ATTR(1, false, 128); KT2_ATTR(2, false, 128); KT2_ATTR(4, false, 128);
        KT2_ATTR(1, true, 128); KT2_ATTR(2, true, 128); KT2_ATTR(4, true, 128);
#undef KT2_ATTR
    }
    if (smem > 227 * 1024 || smem2 > 227 * 1024) return fail(TGP_EINVAL, "tgp_knn_feat: shared memory budget exceeded");
    const int grid = num_units < TGP_NUM_SMS ? num_units : TGP_NUM_SMS;
    if (!fix_list) {
        if (big) return fail(TGP_EINVAL, "tgp_knn_feat: the insertion-list kernel covers k <= 31 only");
        knn_tc_kernel<<<grid, KT_THREADS, smem, st>>>(tmQ, tmC, qn, N, Kp, k, q_tiles, q_step, num_units, nullptr, idx64, idx32);
        return check_launch("knn_tc_kernel");
    }
    cudaMemsetAsync(fix_list, 0, sizeof(int), st);
    // query tile in tensor memory when [tf32 | residual] fits the 256 free columns in whole 32-column stores per chunk
    const bool atm = Kp <= 128 && Kp % 64 == 0;
#define KT2_LAUNCH(GS, ATM, NS) knn_tc2_kernel<GS, ATM, NS><<<grid, KT_THREADS, smem2, st>>>(tmQ, tmC, x_split, qn, N, Kp, k, q_tiles, q_step, num_units, fix_list, idx64, idx32)
    // group size: the slots (group % NS) must hold at least K finite minima -> groups of 1 / 2 / 4 consecutive candidates
    if (!big) {
        if (atm) { if (N <= 64) KT2_LAUNCH(1, true, 64); else if (N < 256) KT2_LAUNCH(2, true, 64); else KT2_LAUNCH(4, true, 64); }
        else { if (N <= 64) KT2_LAUNCH(1, false, 64); else if (N < 256) KT2_LAUNCH(2, false, 64); else KT2_LAUNCH(4, false, 64); }
    } else {
        if (atm) { if (N < 256) KT2_LAUNCH(1, true, 128); else if (N < 512) KT2_LAUNCH(2, true, 128); else KT2_LAUNCH(4, true, 128); }
        else { if (N < 256) KT2_LAUNCH(1, false, 128); else if (N < 512) KT2_LAUNCH(2, false, 128); else KT2_LAUNCH(4, false, 128); }
    }
#undef KT2_LAUNCH
    rc = check_launch("knn_tc2_kernel");
    if (rc) return rc;
    if (big) return TGP_OK;      // k > 31: the caller redoes the listed units with the fp32 tile kernel (knn.cu)
    // fix-up: the units knn_tc2_kernel listed (normally none: the launch returns at once)
    const int fgrid = grid < 32 ? grid : 32;
    knn_tc_kernel<<<fgrid, KT_THREADS, smem, st>>>(tmQ, tmC, qn, N, Kp, k, q_tiles, q_step, num_units, fix_list, idx64, idx32);
    return check_launch("knn_tc_kernel (fix-up)");
}
