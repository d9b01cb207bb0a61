// gemm_tc.cu -- the dense contractions of the path on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces `feature_map @ self.weights + self.bias` (gcn3d.py:170), STE_layer / conv2 (gcn3d.py:70-71,
// 130,132) and the 1x1 Conv1d head layers at fp32-equivalent accuracy: 3xTF32.
//
// Why 3xTF32: a single TF32 pass misses the rel-1e-4 tolerance on 41.7 % of the final features and flips
// 5-9 % of the downstream feature-space kNN rows.  Each operand is split once into
// hi = tf32(x) and lo = x - hi (both exactly representable), stored side by side as [hi | lo] (2*Kp
// columns, Kp = K rounded up to 32), and the K loop runs three segments into ONE TMEM accumulator:
//     lo(A).hi(B) + hi(A).lo(B) + hi(A).hi(B)          (small terms first)
// so the 3xTF32 product is a plain TF32 GEMM over a 3x longer K -- no operand transform inside the
// kernel, TMA feeds the tensor core directly.
//
// Kernel: persistent, one CTA per SM, 192 threads:
//   warp 0    TMA producer   cp.async.bulk.tensor.2d (128B swizzle) -> 4-stage smem ring, mbarrier tx-count
//   warp 1    MMA issuer     tcgen05.mma.cta_group::1.kind::tf32, M=128 x N=BN x K=8, fp32 accum in TMEM,
//                            tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2-9 epilogue       tcgen05.ld 32x32b.x32 -> registers -> smem transpose -> fused epilogue -> coalesced global
//                            (two warps per TMEM lane quarter, interleaved 32-column chunks)
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile t overlap the mainloop of t+1.
// Tile order is n-fastest: the ~148 tiles in flight cover a band of a few m-tiles x all n-tiles, so each A tile
// is fetched from HBM once and re-read from L2 by the CTAs working on its other column blocks (ncu, round 1:
// m-fastest order streamed the A operand from DRAM once per column block -- 8.2 GB for the 4096-wide head GEMM).
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdlib.h>

namespace tgp {

// operand ring depth: as many stages as fit in ~192 KB, so that narrow tiles (short, latency-bound mainloops) keep
// as many bytes in flight as wide ones: BN 256 -> 4 x 48 KB, BN 128 -> 6 x 32 KB, BN 64 -> 8 x 24 KB
template <int BN> struct TcStages { static constexpr int value = (192 * 1024) / (TC_BM * TC_BK * 4 + BN * TC_BK * 4); };
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // TMA warp + MMA warp + epilogue warps

struct GemmDev {
    tgp_gemm_args a;
    int blocked;       // 1: operands are the K-blocked transposed splits of the weight-gradient contraction (tgp_gemm_tn_tc)
                       // 2: row-major mixed operands read in place as MN-major tiles (tgp_gemm_tn_tc_rm)
    int kp_a, kp_b;    // blocked == 2: 16-bit slots per part of the A / B operand rows
};

// one output destination of a column: pointer to (row0, col), row stride, and (split mode) the lo-half offset
struct EpiDst {
    float* p;
    long rs;
    int lo;      // > 0: split mode, offset of the residual half;  -1: per-group column max (p -> encoded int cell of group 0)
    int mix;     // > 0: MIXED operand, mix = Kp and `rel` = column inside the operand (p points at the fp32 slot)
    int rel;
};

// MIXED operand row (common.cuh); p = address of the fp32 slot `rel` of the row, i.e. row base + rel floats
__device__ __forceinline__ void store_mixed(float* p, int Kp, int rel, float v, bool nb = false) {
    mixed_store1(reinterpret_cast<uint16_t*>(p - rel), Kp, rel, v, nb);
}

// order-preserving float -> int map for atomicMax (mode 3): signed-int order == float order
__device__ __forceinline__ int enc_ordered(float v) {
    const int i = __float_as_int(v);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

__device__ __forceinline__ void epi_store(EpiDst& d, float v, bool nb) {
    if (d.p) {
        if (d.mix) {
            store_mixed(d.p, d.mix, d.rel, v, nb);
        } else if (d.lo) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float hi = __uint_as_float(hb);
            d.p[0] = hi;
            d.p[d.lo] = v - hi;
        } else d.p[0] = v;
        d.p += d.rs;
    }
}

// rows of one 32-column chunk: lane = column.  Loads (shared-memory staging + residuals) for 16 rows are issued
// before any of their stores so that the global loads overlap (ncu round 1: the row-at-a-time loop stalled on
// long_scoreboard for every residual load).
template <bool D1, bool MX>
__device__ __forceinline__ void epi_rows(uint32_t stg_addr, int lane, int nrows, float bias, float sc, float sh,
                                         float slope, float gbv0, float gbv1, int gb_switch, const float* r1p,
                                         long ld1, const float* r2p, long ld2, EpiDst d0, EpiDst d1, bool nb) {
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
    const bool is_max = MX && d0.lo < 0;
#pragma unroll 1
    for (int rr0 = 0; rr0 < nrows; rr0 += 16) {
        float a[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) a[u] = lds_f32(stg_addr + (uint32_t)(((rr0 + u) * 33 + lane) * 4)) + bias;
        if (r1p) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (rr0 + u < nrows) a[u] += __ldg(r1p + (long)(rr0 + u) * ld1);
        }
        if (r2p) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (rr0 + u < nrows) a[u] += __ldg(r2p + (long)(rr0 + u) * ld2);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (rr0 + u < nrows) {
                float v = a[u] + ((rr0 + u) >= gb_switch ? gbv1 : gbv0);
                v = fmaf(v, sc, sh);
                v = v > 0.f ? v : v * slope;
                if (MX && is_max) {
                    if ((rr0 + u) >= gb_switch) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
                } else {
                    epi_store(d0, v, nb);
                    if (D1) epi_store(d1, v, nb);
                }
            }
        }
    }
    if (MX && is_max && nrows > 0) {
        int* cell = reinterpret_cast<int*>(d0.p);
        if (gb_switch > 0) atomicMax(cell, enc_ordered(mx0));
        if (gb_switch < nrows) atomicMax(cell + d0.rs, enc_ordered(mx1));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// vectorised epilogue of one 32-row x 32-column chunk (the common case: every column count, segment boundary and
// leading dimension a multiple of 4 and every pointer 16-byte aligned -- checked on the host, `vec_ok`).
// ncu, round 1: the scalar epilogue (lane = column, one row per instruction) ran ~3000 latency-bound instructions per
// chunk and warp -- 35-40 us per output tile with only 8 epilogue warps resident, which bounded every small GEMM.
// Here the chunk is staged with an XOR swizzle (conflict-free 128-bit writes by row and reads by (row, 4 columns)),
// a lane owns 4 consecutive columns, a warp instruction covers 4 rows, and all global traffic is 128-bit.
struct VDst {
    float* p;      // (row0, first column of this lane)
    long rs;       // row stride
    int kind;      // 0 none, 1 raw, 2 split, 3 column max, 4 mixed operand
    int lo;        // split: offset of the residual half; mixed: Kp
    int rel;       // mixed: column inside the operand
};

__device__ __forceinline__ float4 ldg4s(const float* p) {   // 4 scalar loads: parameter vectors may be unaligned views
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float act1(float v, float sc, float sh, float sl) {
    v = fmaf(v, sc, sh);
    return v > 0.f ? v : v * sl;
}
__device__ __forceinline__ void vstore(const VDst& d, int row, float4 v, bool nb) {
    if (d.kind == 1) {
        *reinterpret_cast<float4*>(d.p + row * d.rs) = v;
    } else if (d.kind == 2) {
        float4 hi, lo;
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.x)); hi.x = __uint_as_float(hb); lo.x = v.x - hi.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.y)); hi.y = __uint_as_float(hb); lo.y = v.y - hi.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.z)); hi.z = __uint_as_float(hb); lo.z = v.z - hi.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.w)); hi.w = __uint_as_float(hb); lo.w = v.w - hi.w;
        float* q = d.p + row * d.rs;
        *reinterpret_cast<float4*>(q) = hi;
        *reinterpret_cast<float4*>(q + d.lo) = lo;
    } else if (d.kind == 4) {
        mixed_store4_cs(reinterpret_cast<uint16_t*>(d.p + row * d.rs - d.rel), d.lo, d.rel, v, nb);
    }
}

// LEAN (host-checked: no residual, no per-cloud bias, segments do not overlap): the epilogue of the
// projection / head GEMMs that only add a bias, optionally apply the BatchNorm affine + activation and write ONE
// destination per column -- ~4x fewer instructions per chunk than the general path, which matters because only 8 epilogue
// warps are resident and the K = 128..256 encoder GEMMs are bound by exactly this code.
// stage one chunk: thread = row writes its 32 columns as 8 float4, slot j ^ (row & 7) of its 128-byte line
__device__ __forceinline__ void epi_stage_vec(uint32_t stg_addr, const uint32_t (&r)[32], int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t ad = stg_addr + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                     "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
    }
}

template <bool LEAN>
__device__ __forceinline__ void epi_chunk_vec(const tgp_gemm_args& g, uint32_t stg_addr, int lane,
                                              long row0, int nrows, int colbase, long grp0, int gb_switch, long zoff, int dbg = 0,
                                              int ri1 = 0, int ri2 = 0, bool nb = false) {
    const int c4i = lane & 7, rsub = lane >> 3;
    const int col = colbase + c4i * 4;
    const bool live = col < g.Ncols && nrows > 0;     // row blocks past M must not touch group_bias / residual rows
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = bias, gb0 = bias, gb1 = bias;
    const float s0 = g.relu ? 0.f : 1.f;
    float4 sl = make_float4(s0, s0, s0, s0);
    VDst d0 = {nullptr, 0, 0, 0, 0}, d1 = {nullptr, 0, 0, 0, 0};
    const float* r1p = nullptr;
    const float* r2p = nullptr;
    if (live) {
        if (g.bias) bias = ldg4s(g.bias + col);
        if (g.scale) { sc = ldg4s(g.scale + col); sh = ldg4s(g.shift + col); }
        if (g.neg_slope) sl = ldg4s(g.neg_slope + col);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s < g.nseg && col >= g.seg[s].col_begin && col < g.seg[s].col_end) {
                const int rel = col - g.seg[s].col_begin;
                VDst d;
                d.lo = 0;
                d.rel = rel;
                if (g.seg[s].mode == 3) {
                    d.kind = 3;
                    d.rs = g.seg[s].col_end - g.seg[s].col_begin;
                    d.p = g.seg[s].ptr + grp0 * d.rs + rel;
                } else if (g.seg[s].mode == 1) {
                    const int w = g.seg[s].slab_width;
                    const int cg = rel / w, rr = rel - cg * w;
                    d.kind = 1;
                    d.rs = w;
                    d.p = g.seg[s].ptr + ((long)cg * g.M + row0) * w + rr;
                } else {
                    d.kind = g.seg[s].mode == 2 ? 2 : (g.seg[s].mode >= 4 ? 4 : 1);
                    d.rs = g.seg[s].ld;
                    d.p = g.seg[s].ptr + zoff + row0 * d.rs + rel;
                    d.lo = g.seg[s].slab_width;
                }
                if (LEAN || !d0.kind) d0 = d; else d1 = d;
            }
        }
        if (!LEAN) {
            if (g.group_bias) {
                gb0 = ldg4s(g.group_bias + grp0 * g.Ncols + col);
                if (gb_switch < nrows) gb1 = ldg4s(g.group_bias + (grp0 + 1) * g.Ncols + col);
            }
            // residual rows: lane L of the warp holds the residual row of output row row0 + L (ri1 / ri2: the row itself, or
            // res*_idx[row] for GATHERED residuals)
            if (g.res1) r1p = g.res1 + col;
            if (g.res2) r2p = g.res2 + col;
        }
    }
    float4 mx0 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F), mx1 = mx0;
    const bool any_max = d0.kind == 3 || (!LEAN && d1.kind == 3);
    const bool has_act = g.scale != nullptr || g.neg_slope != nullptr || g.relu != 0;       // warp-uniform
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float4 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = (half * 4 + u) * 4 + rsub;
            const uint32_t ad = stg_addr + (uint32_t)(row * 128 + ((c4i ^ (row & 7)) << 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "r"(ad));
        }
        if (LEAN) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = (half * 4 + u) * 4 + rsub;
                if (live && row < nrows) {
                    float4 v = f4add(a[u], bias);
                    if (has_act) {
                        v.x = act1(v.x, sc.x, sh.x, sl.x); v.y = act1(v.y, sc.y, sh.y, sl.y);
                        v.z = act1(v.z, sc.z, sh.z, sl.z); v.w = act1(v.w, sc.w, sh.w, sl.w);
                    }
                    if (d0.kind == 3) {            // per-cloud column max: nothing is stored, the running maxima go out below
                        float4& m = row >= gb_switch ? mx1 : mx0;
                        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                    } else if (!(dbg & 8)) vstore(d0, row, v, nb);
                    else if (v.x == 1.2345e33f) d0.p[0] = v.y;       // (profiling switch: keep the arithmetic, drop the stores)
                }
            }
            continue;
        }
        float4 q1[4], q2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = (half * 4 + u) * 4 + rsub;
            q1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            q2[u] = q1[u];
            const int i1 = __shfl_sync(0xffffffffu, ri1, row), i2 = __shfl_sync(0xffffffffu, ri2, row);
            if (row < nrows) {
                if (r1p) q1[u] = __ldg(reinterpret_cast<const float4*>(r1p + (long)i1 * g.ld_res1));
                if (r2p) q2[u] = __ldg(reinterpret_cast<const float4*>(r2p + (long)i2 * g.ld_res2));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = (half * 4 + u) * 4 + rsub;
            if (live && row < nrows) {
                const bool second = row >= gb_switch;
                float4 v = f4add(f4add(f4add(a[u], bias), f4add(q1[u], q2[u])), second ? gb1 : gb0);
                v.x = act1(v.x, sc.x, sh.x, sl.x); v.y = act1(v.y, sc.y, sh.y, sl.y);
                v.z = act1(v.z, sc.z, sh.z, sl.z); v.w = act1(v.w, sc.w, sh.w, sl.w);
                vstore(d0, row, v, nb);
                vstore(d1, row, v, nb);
                if (any_max) {
                    float4& m = second ? mx1 : mx0;
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, any_max)) {
        // combine the 4 row sub-groups (lanes differing in bits 3, 4), then one atomicMax per column and group
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            mx0.x = fmaxf(mx0.x, __shfl_xor_sync(0xffffffffu, mx0.x, o)); mx0.y = fmaxf(mx0.y, __shfl_xor_sync(0xffffffffu, mx0.y, o));
            mx0.z = fmaxf(mx0.z, __shfl_xor_sync(0xffffffffu, mx0.z, o)); mx0.w = fmaxf(mx0.w, __shfl_xor_sync(0xffffffffu, mx0.w, o));
            mx1.x = fmaxf(mx1.x, __shfl_xor_sync(0xffffffffu, mx1.x, o)); mx1.y = fmaxf(mx1.y, __shfl_xor_sync(0xffffffffu, mx1.y, o));
            mx1.z = fmaxf(mx1.z, __shfl_xor_sync(0xffffffffu, mx1.z, o)); mx1.w = fmaxf(mx1.w, __shfl_xor_sync(0xffffffffu, mx1.w, o));
        }
        if (any_max && live && rsub == 0 && nrows > 0) {
            const VDst& dm = (LEAN || d0.kind == 3) ? d0 : d1;
            int* cell = reinterpret_cast<int*>(dm.p);
            atomicMax(cell, enc_ordered(mx0.x)); atomicMax(cell + 1, enc_ordered(mx0.y));
            atomicMax(cell + 2, enc_ordered(mx0.z)); atomicMax(cell + 3, enc_ordered(mx0.w));
            if (gb_switch < nrows) {
                cell += dm.rs;
                atomicMax(cell, enc_ordered(mx1.x)); atomicMax(cell + 1, enc_ordered(mx1.y));
                atomicMax(cell + 2, enc_ordered(mx1.z)); atomicMax(cell + 3, enc_ordered(mx1.w));
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// FAST lean chunk (vec_ok == 3: lean + 16-byte aligned parameter vectors; all 32 rows live; the chunk's 32 columns lie in
// ONE segment of kind raw / slab / split / mixed).  The 8 epilogue warps are latency-bound, not issue-bound (halving them
// takes 90 -> 141 us on the level-0 projection), so what counts is the length of the per-chunk dependency chain: the
// destination of a lane is computed ONCE per tile (before the accumulator wait) and fetched by shuffles, parameters come as
// 128-bit loads, shared-memory addresses are loop invariants, all eight LDS.128 are issued before the first store, and the
// kind / activation switches are compile-time.
struct FastDst {
    float* p;      // (row0, this lane's first column)
    int rs;        // row stride in floats
    int kind;      // 0 none, 1 raw / slab, 2 split, 3 column max (-> general lean path), 4 mixed operand
    int lo;        // split: offset of the residual half; mixed: Kp
    int rel;       // column inside the segment
};

__device__ __forceinline__ FastDst fast_entry(const tgp_gemm_args& g, int col, long row0, long zoff, long grp0) {
    FastDst d = {nullptr, 0, 0, 0, 0};
    if (col >= g.Ncols) return d;
    // pick the segment first (cheap compares; segments are disjoint on the fast paths), then do the arithmetic for it alone:
    // this runs once per tile and epilogue warp, in front of the accumulator wait, and short-K launches feel its length
    int sidx = -1;
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (s < g.nseg && col >= g.seg[s].col_begin && col < g.seg[s].col_end) sidx = s;
    if (sidx < 0) return d;
    const tgp_out_seg& sg = g.seg[sidx];
    const int rel = col - sg.col_begin;
    d.rel = rel;
    if (sg.mode == 3) {
        d.kind = 3;      // (the lean fast path leaves column maxima to the general lean chunk; the residual one uses p / rs)
        d.rs = sg.col_end - sg.col_begin;
        d.p = sg.ptr + grp0 * d.rs + rel;
    } else if (sg.mode == 1) {
        const unsigned w = (unsigned)sg.slab_width;
        const unsigned cg = (unsigned)rel / w, rr = (unsigned)rel - cg * w;
        d.kind = 1;
        d.rs = (int)w;
        d.p = sg.ptr + ((long)cg * g.M + row0) * w + rr;
    } else {
        d.kind = sg.mode == 2 ? 2 : (sg.mode >= 4 ? 4 : 1);
        d.rs = (int)sg.ld;
        d.p = sg.ptr + zoff + row0 * sg.ld + rel;
        d.lo = sg.slab_width;
    }
    return d;
}

__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool ACT, int KIND>
__device__ __forceinline__ void epi_fast_rows(uint32_t ld_even, uint32_t ld_odd, float* p_r, int rs, int lo, int rel,
                                              float4 bias, float4 sc, float4 sh, float4 sl, bool nb) {
    float4 a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const uint32_t ad = ((u & 1) ? ld_odd : ld_even) + (uint32_t)(u * 512);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "r"(ad));
    }
    const long step = 4L * rs;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float4 v = f4add(a[u], bias);
        if (ACT) {
            v.x = act1(v.x, sc.x, sh.x, sl.x); v.y = act1(v.y, sc.y, sh.y, sl.y);
            v.z = act1(v.z, sc.z, sh.z, sl.z); v.w = act1(v.w, sc.w, sh.w, sl.w);
        }
        float* q = p_r + u * step;
        if (KIND == 1) {
            *reinterpret_cast<float4*>(q) = v;
        } else if (KIND == 2) {
            float4 hi, lw;
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.x)); hi.x = __uint_as_float(hb); lw.x = v.x - hi.x;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.y)); hi.y = __uint_as_float(hb); lw.y = v.y - hi.y;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.z)); hi.z = __uint_as_float(hb); lw.z = v.z - hi.z;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.w)); hi.w = __uint_as_float(hb); lw.w = v.w - hi.w;
            *reinterpret_cast<float4*>(q) = hi;
            *reinterpret_cast<float4*>(q + lo) = lw;
        } else {
            mixed_store4_cs(reinterpret_cast<uint16_t*>(q - rel), lo, rel, v, nb);
        }
    }
}

template <bool ACT>
__device__ __forceinline__ void epi_fast_kind(int kind, uint32_t ld_even, uint32_t ld_odd, float* p_r, int rs, int lo, int rel,
                                              float4 bias, float4 sc, float4 sh, float4 sl, bool nb) {
    if (kind == 1) epi_fast_rows<ACT, 1>(ld_even, ld_odd, p_r, rs, lo, rel, bias, sc, sh, sl, nb);
    else if (kind == 2) epi_fast_rows<ACT, 2>(ld_even, ld_odd, p_r, rs, lo, rel, bias, sc, sh, sl, nb);
    else epi_fast_rows<ACT, 4>(ld_even, ld_odd, p_r, rs, lo, rel, bias, sc, sh, sl, nb);
}


// FAST chunk WITH residuals / per-cloud bias (vec_ok == 4: one destination per column, 16-byte aligned parameter vectors, all 32
// rows live): the ORL contractions (conv2(cat[f, g]) + f + f_STE, gcn3d.py:110-112,184-186: two residuals + the per-cloud term)
// and the heads' first layers with GATHERED residuals (tgp_gemm_args.res1_idx).  Same structure as the lean fast chunk; the
// residual rows of the 32 output rows sit one per lane (ri1 / ri2) and are fetched by shuffles; 4 rows (a 128-bit staging
// read + two 128-bit residual loads each) are in flight before the first store.
// RESIDUAL PREFETCH (vec_ok == 5: the fast residual chunk of a 256-column-tile launch).  ncu / K sweep: the two residual loads
// of a row cost the chunk two exposed L2 round trips (~1500 cycles each, 170 of the 436 us of the heads' first layer) and the
// register budget (168 with 10 warps) has no room to keep more of them in flight.  These launches are epilogue-bound, so
// they run their operand ring two stages deep and the freed 96 KB hold, per epilogue warp, the residual pieces of ONE chunk
// ([half][residual][row quad][lane] x 16 B = 8 KB), fetched by cp.async one half-chunk ahead: every lane copies exactly the
// 16 bytes it will add, so no barrier is involved -- only its own cp.async groups (one per half-chunk, in order).
__device__ __forceinline__ void cp_async16_g2s(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void res_prefetch_half(uint32_t rb_lane, int h, const float* r1c, long ld1, int ri1, const float* r2c,
                                                  long ld2, int ri2, int rsub, bool on) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int row = 4 * (h * 4 + u) + rsub;
        const int i1 = __shfl_sync(0xffffffffu, ri1, row), i2 = __shfl_sync(0xffffffffu, ri2, row);
        if (on && r1c) cp_async16_g2s(rb_lane + (uint32_t)(((h * 2 + 0) * 4 + u) * 512), r1c + (long)i1 * ld1);
        if (on && r2c) cp_async16_g2s(rb_lane + (uint32_t)(((h * 2 + 1) * 4 + u) * 512), r2c + (long)i2 * ld2);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");      // (an empty group when nothing was issued: the counts stay uniform)
}

// BG: the launch has a bias or a per-cloud bias (false: both adds and the per-row group select drop out of the chunk -- the heads'
// factored first layers carry their bias in the BatchNorm shift)
template <int KIND, bool BUF, bool BG>
__device__ __forceinline__ void epi_fast_res_rows(uint32_t ld_even, uint32_t ld_odd, float* p_r, int rs, int lo, int rel, int rsub,
                                                  float4 bias, float4 sc, float4 sh, float4 sl, const float* r1c, long ld1, int ri1,
                                                  const float* r2c, long ld2, int ri2, float4 gb0, float4 gb1, int gb_switch,
                                                  bool nb, uint32_t rb_lane = 0, bool has_next = false) {
    float4 mx0 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F), mx1 = mx0;
    const long step = 4L * rs;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 a[4], q1[4], q2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int uu = h * 4 + u;
            const uint32_t ad = ((uu & 1) ? ld_odd : ld_even) + (uint32_t)(uu * 512);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "r"(ad));
        }
        if (BUF) {
            // this half's pieces were requested one half-chunk ago; at most the following half's group may still be pending
            asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q1[u] = zero4; q2[u] = zero4;
                if (r1c) { const uint32_t ad = rb_lane + (uint32_t)(((h * 2 + 0) * 4 + u) * 512);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q1[u].x), "=f"(q1[u].y), "=f"(q1[u].z), "=f"(q1[u].w) : "r"(ad)); }
                if (r2c) { const uint32_t ad = rb_lane + (uint32_t)(((h * 2 + 1) * 4 + u) * 512);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q2[u].x), "=f"(q2[u].y), "=f"(q2[u].z), "=f"(q2[u].w) : "r"(ad)); }
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = 4 * (h * 4 + u) + rsub;
                const int i1 = __shfl_sync(0xffffffffu, ri1, row), i2 = __shfl_sync(0xffffffffu, ri2, row);
                q1[u] = r1c ? ldg128(r1c + (long)i1 * ld1) : zero4;
                q2[u] = r2c ? ldg128(r2c + (long)i2 * ld2) : zero4;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int uu = h * 4 + u, row = 4 * uu + rsub;
            const bool second = row >= gb_switch;
            float4 v = BG ? f4add(f4add(f4add(a[u], bias), f4add(q1[u], q2[u])), second ? gb1 : gb0)
                          : f4add(a[u], f4add(q1[u], q2[u]));
            v.x = act1(v.x, sc.x, sh.x, sl.x); v.y = act1(v.y, sc.y, sh.y, sl.y);
            v.z = act1(v.z, sc.z, sh.z, sl.z); v.w = act1(v.w, sc.w, sh.w, sl.w);
            float* q = p_r + uu * step;
            if (KIND == 1) {
                *reinterpret_cast<float4*>(q) = v;
            } else if (KIND == 2) {
                float4 hi, lw;
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.x)); hi.x = __uint_as_float(hb); lw.x = v.x - hi.x;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.y)); hi.y = __uint_as_float(hb); lw.y = v.y - hi.y;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.z)); hi.z = __uint_as_float(hb); lw.z = v.z - hi.z;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.w)); hi.w = __uint_as_float(hb); lw.w = v.w - hi.w;
                *reinterpret_cast<float4*>(q) = hi;
                *reinterpret_cast<float4*>(q + lo) = lw;
            } else if (KIND == 4) {
                mixed_store4_cs(reinterpret_cast<uint16_t*>(q - rel), lo, rel, v, nb);
            } else {
                float4& m = second ? mx1 : mx0;
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        // the values of this half are consumed: its buffer takes the same half of the NEXT chunk (64 columns further)
        if (BUF) res_prefetch_half(rb_lane, h, r1c ? r1c + 64 : nullptr, ld1, ri1, r2c ? r2c + 64 : nullptr, ld2, ri2, rsub, has_next);
    }
    if (KIND == 3) {
        // per-cloud column max: combine the 4 row sub-groups (lane bits 3, 4), one atomicMax per column and cloud
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            mx0.x = fmaxf(mx0.x, __shfl_xor_sync(0xffffffffu, mx0.x, o)); mx0.y = fmaxf(mx0.y, __shfl_xor_sync(0xffffffffu, mx0.y, o));
            mx0.z = fmaxf(mx0.z, __shfl_xor_sync(0xffffffffu, mx0.z, o)); mx0.w = fmaxf(mx0.w, __shfl_xor_sync(0xffffffffu, mx0.w, o));
            mx1.x = fmaxf(mx1.x, __shfl_xor_sync(0xffffffffu, mx1.x, o)); mx1.y = fmaxf(mx1.y, __shfl_xor_sync(0xffffffffu, mx1.y, o));
            mx1.z = fmaxf(mx1.z, __shfl_xor_sync(0xffffffffu, mx1.z, o)); mx1.w = fmaxf(mx1.w, __shfl_xor_sync(0xffffffffu, mx1.w, o));
        }
        if (rsub == 0) {
            // p_r = cell of (first cloud of the block, this lane's first column) for rsub == 0; rs = cells per cloud
            int* cell = reinterpret_cast<int*>(p_r);
            if (gb_switch > 0) {
                atomicMax(cell, enc_ordered(mx0.x)); atomicMax(cell + 1, enc_ordered(mx0.y));
                atomicMax(cell + 2, enc_ordered(mx0.z)); atomicMax(cell + 3, enc_ordered(mx0.w));
            }
            if (gb_switch < 32) {
                cell += rs;
                atomicMax(cell, enc_ordered(mx1.x)); atomicMax(cell + 1, enc_ordered(mx1.y));
                atomicMax(cell + 2, enc_ordered(mx1.z)); atomicMax(cell + 3, enc_ordered(mx1.w));
            }
        }
    }
}

// NB: the launch's mixed destinations leave out the bf16(x) slot (output mode 5; see tgp_gemm_args.mixed == 2)
template <int BN, bool NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmB16,
               const __grid_constant__ GemmDev P, int Kp, int num_n_tiles, int num_tiles, int dbg,
               int tiles_mn, int kb_per, long zstride, int vec_ok, int nst) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    constexpr int B_BYTES = BN * TC_BK * 4;
    constexpr int TC_STAGES = TcStages<BN>::value;
    constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    // carve: [stages x (A | B)] [barriers] [tmem ptr]
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + TC_STAGES * STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* stage_all = reinterpret_cast<float*>(base + TC_STAGES * STAGE_BYTES + 256);   // [epi warps][32][33]

    const tgp_gemm_args& g = P.a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = P.a.mixed ? Kp / 64 : Kp / TC_BK;      // K blocks: 64 columns (mixed) or 32 (tf32)

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(full + s, 1); tc_mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { tc_mbar_init(tmem_full + s, 1); tc_mbar_init(tmem_empty + s, TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // one full warp allocates all 512 TMEM columns (1 CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int z = t / tiles_mn, tt = t - z * tiles_mn;       // split-K slice z (weight-gradient shapes)
                const int m0 = (tt / num_n_tiles) * TC_BM, n0 = (tt % num_n_tiles) * BN;
                const int kb0 = z * kb_per, kb1 = min(kblocks, kb0 + kb_per);
                if (g.mixed) {
                    // per 64 columns of K three 16-bit stages (64 columns = one 128-byte swizzle span each):
                    //   fp16(a).fp16(b), lo(a).bf16(b), bf16(a).lo(b).  Every stage is 16 KB + BN*128 B.
                    for (int k64 = kb0; k64 < kb1; ++k64) {
                        for (int u = 0; u < 3; ++u) {
                            tc_mbar_wait(empty + stage, phase ^ 1);
                            unsigned char* sa = base + stage * STAGE_BYTES;
                            tc_mbar_expect_tx(full + stage, STAGE_BYTES);
                            // (mixed == 2: the weights' residual slot is fp16, so the third pass pairs it with fp16(a) again)
                            const int a_part = u == 0 ? 0 : (u == 1 ? 2 : (g.mixed == 2 ? 0 : 1)), b_part = u == 0 ? 0 : (u == 1 ? 1 : 2);
                            if (P.blocked == 2) {
                                // ROW-MAJOR mixed operands read in place as MN-major tiles (tgp_gemm_tn_tc_rm): the contraction
                                // runs over the ROWS; a stage = 64 rows x (128 | BN) columns of one 16-bit part, fetched as
                                // 64-column boxes that land 8 KB apart (the descriptors' leading byte offset)
                                for (int h = 0; h < TC_BM / 64; ++h)
                                    tma_load_2d(sa + h * 8192, &tmA16, a_part * P.kp_a + m0 + h * 64, k64 * 64, full + stage);
                                for (int h = 0; h < BN / 64; ++h)
                                    tma_load_2d(sa + TC_A_BYTES + h * 8192, &tmB16, b_part * P.kp_b + n0 + h * 64, k64 * 64, full + stage);
                            } else if (P.blocked) {
                                // K-blocked transposed operands (tgp_split_mixed_t): a tile is one contiguous block
                                const int nblk = Kp / 64;
                                const long RA = ((long)g.M + 255) / 256 * 256, RB = ((long)g.Ncols + 255) / 256 * 256;
                                tma_load_2d(sa, &tmA16, 0, (int)(((long)a_part * nblk + k64) * RA + m0), full + stage);
                                tma_load_2d(sa + TC_A_BYTES, &tmB16, 0, (int)(((long)b_part * nblk + k64) * RB + n0), full + stage);
                            } else {
                                // (grouped contraction: the n-tile's group selects a Kp-wide column block of a wider A operand)
                                const int a_col = g.a_group_cols > 0 ? a_part * g.a_kp + (n0 / g.a_group_cols) * Kp : a_part * Kp;
                                tma_load_2d(sa, &tmA16, a_col + k64 * 64, m0, full + stage);
                                tma_load_2d(sa + TC_A_BYTES, &tmB16, b_part * Kp + k64 * 64, n0, full + stage);
                            }
                            if (++stage == nst) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
                for (int seg = 0; seg < 3; ++seg) {
                    // segment order: lo.hi, hi.lo, hi.hi
                    const int a_off = (seg == 0) ? Kp : 0, b_off = (seg == 1) ? Kp : 0;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        tc_mbar_wait(empty + stage, phase ^ 1);
                        unsigned char* sa = base + stage * STAGE_BYTES;
                        if (dbg & 4) { tc_mbar_arrive(full + stage); }
                        else {
                            tc_mbar_expect_tx(full + stage, STAGE_BYTES);
                            if (P.blocked) {
                                // K-blocked transposed splits (tgp_split_tf32_t): [part][block of 32][ceil256(rows)][32]
                                const long RA = ((long)g.M + 255) / 256 * 256, RB = ((long)g.Ncols + 255) / 256 * 256;
                                tma_load_2d(sa, &tmA, 0, (int)(((long)(a_off ? 1 : 0) * kblocks + kb) * RA + m0), full + stage);
                                tma_load_2d(sa + TC_A_BYTES, &tmB, 0, (int)(((long)(b_off ? 1 : 0) * kblocks + kb) * RB + n0), full + stage);
                            } else {
                                tma_load_2d(sa, &tmA, a_off + kb * TC_BK, m0, full + stage);
                                tma_load_2d(sa + TC_A_BYTES, &tmB, b_off + kb * TC_BK, n0, full + stage);
                            }
                        }
                        if (++stage == nst) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues: descriptors and loop state
        // stay in uniform registers -- ~2 instructions per MMA instead of ~8 with a single-lane branch) =====================
        {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            // kind::f16, fp32 accumulate: bf16 operands (format 1) for the cross terms, fp16 operands (format 0) for hi.hi
            constexpr uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            constexpr uint32_t idesc_h = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                tc_mbar_wait(tmem_empty + acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accum = 0;
                const int zz = t / tiles_mn;
                const int nkb = 3 * (min(kblocks, zz * kb_per + kb_per) - zz * kb_per);
                int u3 = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    tc_mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = s_u32(base + stage * STAGE_BYTES);
                    const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + TC_A_BYTES);
                    const uint32_t id = (u3 == 0 || (u3 == 2 && g.mixed == 2)) ? idesc_h : idesc16;
                    if (++u3 == 3) u3 = 0;
                    if (g.mixed && P.blocked == 2) {
                        // MN-major A and B (instruction-descriptor bits 15 / 16); 64 contraction rows per stage = 4 x K16,
                        // each K16 step = two 8-row swizzle atoms = 2048 B further
                        const uint64_t ad = make_sw128_mn_desc(sa, 8192), bd = make_sw128_mn_desc(sa + TC_A_BYTES, 8192);
                        if (tc_elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d_tmem, ad + (uint64_t)(128 * k), bd + (uint64_t)(128 * k), id | (1u << 15) | (1u << 16), k ? 1u : accum);
                            umma_commit(empty + stage);
                        }
                        __syncwarp();
                        accum = 1;
                        if (++stage == nst) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (tc_elect_one()) {
                        if (g.mixed) {
                            // 16-bit stage: 64 columns of K, 16 per instruction = the same +32 B descriptor step
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), id, k ? 1u : accum);
                        } else if (!(dbg & 2)) {
#pragma unroll
                            for (int k = 0; k < TC_BK / 8; ++k)
                                // +32 B along K inside the swizzle span = +2 in the descriptor's 16-byte address units
                                umma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : accum);
                        }
                        umma_commit(empty + stage);
                    }
                    __syncwarp();
                    accum = 1;
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
                if (tc_elect_one()) umma_commit(tmem_full + acc);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 -> TMEM lane quarters 2,3,0,1,2,3,0,1) =====================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;            // which interleaved set of 32-column chunks
        float* stg = stage_all + (warp - 2) * (32 * 33);
        const uint32_t stg_addr = s_u32(stg);
        const int c4i = lane & 7, rsub = lane >> 3;
        // loop-invariant staging read addresses of the fast path: row = 4 u + rsub, slot c4i ^ (row & 7)
        const uint32_t ld_even = stg_addr + (uint32_t)(rsub * 128 + ((c4i ^ rsub) << 4));
        const uint32_t ld_odd = stg_addr + (uint32_t)(rsub * 128 + ((c4i ^ (4 + rsub)) << 4));
        const bool gb_slow = g.group_bias && g.rows_per_group > 0 && g.rows_per_group < 32;
        const int path = (dbg & 1) ? 0 : ((vec_ok == 2 || vec_ok == 3) ? 2 : ((vec_ok && !gb_slow) ? 1 : 0));   // 2 lean, 1 vector, 0 scalar
        const bool fast_ok = (vec_ok >= 3 && vec_ok <= 5) && !(dbg & 9);
        const bool fast_res = vec_ok == 4 || vec_ok == 5;            // residuals / per-cloud bias on the fast chunk
        const bool res_buf = vec_ok == 5;             // ... with the residual pieces prefetched into shared memory (2-stage ring)
        const uint32_t rb_lane = s_u32(base + 2 * STAGE_BYTES) + (uint32_t)((warp - 2) * 8192 + lane * 16);
        const bool has_act = g.scale != nullptr || g.neg_slope != nullptr || g.relu != 0;
        int it = 0;
        // tile coordinates (split-K slice, row tile, column tile) advanced incrementally: three integer divisions per tile sat
        // in front of every tile's epilogue
        const int num_m_tiles = tiles_mn / num_n_tiles;
        int z = (int)blockIdx.x / tiles_mn, tm, tn;
        { const int tt0 = (int)blockIdx.x - z * tiles_mn; tm = tt0 / num_n_tiles; tn = tt0 - tm * num_n_tiles; }
        const int dz = (int)gridDim.x / tiles_mn, dr = (int)gridDim.x - dz * tiles_mn;
        const int dm = dr / num_n_tiles, dn = dr - dm * num_n_tiles;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it,
                 tn += dn, tm += dm + (tn >= num_n_tiles), tn -= (tn >= num_n_tiles) ? num_n_tiles : 0,
                 z += dz + (tm >= num_m_tiles), tm -= (tm >= num_m_tiles) ? num_m_tiles : 0) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = tm * TC_BM, n0 = tn * BN;
            // TMEM gives each thread one ROW (32 consecutive columns per load); a shared-memory
            // transpose turns that into lane = COLUMN so that every global access of the epilogue
            // (residual loads, output stores) is a full, coalesced 128-byte row segment.
            const int row0 = m0 + quarter * 32;          // (32-bit: the row count fits the TMA's int coordinates)
            const int nrows = max(0, min(32, (int)g.M - row0));
            // per-cloud bias: at most one group boundary inside these 32 rows when rows_per_group >= 32
            int grp0 = 0;
            int gb_switch = 64;
            if (g.rows_per_group > 0) {
                grp0 = (int)((unsigned)row0 / (unsigned)g.rows_per_group);
                const int nxt = (grp0 + 1) * g.rows_per_group - row0;
                gb_switch = nxt < 64 ? nxt : 64;
            }
            // fast path: lane L owns the destination of (chunk L / 8 of this warp, column quad L % 8), computed while the
            // mainloop of this tile is still running
            FastDst fd = {nullptr, 0, 0, 0, 0};
            if (fast_ok) {
                const int fc0 = half * 32 + 64 * (lane >> 3);
                if (fc0 < BN) fd = fast_entry(g, n0 + fc0 + c4i * 4, row0, z * zstride, grp0);
            }
            // residual row of output row row0 + lane (the row itself, or the gathered one)
            int ri1 = 0, ri2 = 0;
            if (row0 + lane < g.M) {
                if (g.res1) ri1 = g.res1_idx ? __ldg(g.res1_idx + row0 + lane) : (int)(row0 + lane);
                if (g.res2) ri2 = g.res2_idx ? __ldg(g.res2_idx + row0 + lane) : (int)(row0 + lane);
            }
            int pf = -1;          // chunk of this tile whose residual pieces are in flight / in the buffer
            if (res_buf) {
                // first chunk of the tile: requested before the accumulator wait
                const int col = n0 + half * 32 + c4i * 4;
                const bool on = col < g.Ncols;
                const float* r1c = g.res1 ? g.res1 + col : nullptr;
                const float* r2c = g.res2 ? g.res2 + col : nullptr;
                res_prefetch_half(rb_lane, 0, r1c, g.ld_res1, ri1, r2c, g.ld_res2, ri2, rsub, on);
                res_prefetch_half(rb_lane, 1, r1c, g.ld_res1, ri1, r2c, g.ld_res2, ri2, rsub, on);
                pf = 0;
            }
            tc_mbar_wait(tmem_full + acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
            uint32_t r[32];
            TMEM_LD_32x32(taddr0 + (uint32_t)(half * 32), r);
            int ci = 0;
#pragma unroll 1
            for (int c0 = half * 32; c0 < BN; c0 += 64, ++ci) {
                TMEM_WAIT_LD(r);
                if (path) {
                    epi_stage_vec(stg_addr, r, lane);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sts_f32(stg_addr + (uint32_t)((lane * 33 + j) * 4), __uint_as_float(r[j]));
                }
                __syncwarp();
                // the chunk is in shared memory: fetch the next one while this one is processed, or hand the accumulator
                // stage back to the MMA warp right away
                if (c0 + 64 < BN) {
                    TMEM_LD_32x32(taddr0 + (uint32_t)(c0 + 64), r);
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(tmem_empty + acc);
                }
                bool done = false;
                if (fast_ok) {
                    const int src = ci * 8 + c4i;
                    const unsigned long long pp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)fd.p, src);
                    const int rs = __shfl_sync(0xffffffffu, fd.rs, src), kind = __shfl_sync(0xffffffffu, fd.kind, src);
                    const int lo = __shfl_sync(0xffffffffu, fd.lo, src), rel = __shfl_sync(0xffffffffu, fd.rel, src);
                    const int k0 = __shfl_sync(0xffffffffu, kind, 0);
                    if (fast_res) {
                        if (nrows == 32 && k0 >= 1 && k0 <= 4 && __all_sync(0xffffffffu, kind == k0)) {
                            const int col = n0 + c0 + c4i * 4;
                            float* p_r = reinterpret_cast<float*>((uintptr_t)pp) + (k0 == 3 ? 0L : (long)rsub * rs);
                            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            const float s0 = g.relu ? 0.f : 1.f;
                            const float4 bias = g.bias ? ldg128(g.bias + col) : zero4;
                            const float4 sc = g.scale ? ldg128(g.scale + col) : make_float4(1.f, 1.f, 1.f, 1.f);
                            const float4 sh = g.scale ? ldg128(g.shift + col) : zero4;
                            const float4 sl = g.neg_slope ? ldg128(g.neg_slope + col) : make_float4(s0, s0, s0, s0);
                            float4 gb0 = zero4, gb1 = zero4;
                            if (g.group_bias) {
                                gb0 = ldg128(g.group_bias + grp0 * g.Ncols + col);
                                if (gb_switch < 32) gb1 = ldg128(g.group_bias + (grp0 + 1) * g.Ncols + col);
                            }
                            const float* r1c = g.res1 ? g.res1 + col : nullptr;
                            const float* r2c = g.res2 ? g.res2 + col : nullptr;
#define TGP_FAST_RES(KD, BF, BGF, ...) epi_fast_res_rows<KD, BF, BGF>(ld_even, ld_odd, p_r, rs, lo, rel, rsub, bias, sc, sh, sl, r1c, g.ld_res1, \
                                                                      ri1, r2c, g.ld_res2, ri2, gb0, gb1, gb_switch, NB, ##__VA_ARGS__)
                            if (res_buf) {
                                if (pf != ci) {      // (the chunk before this one did not take the fast path: request now)
                                    res_prefetch_half(rb_lane, 0, r1c, g.ld_res1, ri1, r2c, g.ld_res2, ri2, rsub, true);
                                    res_prefetch_half(rb_lane, 1, r1c, g.ld_res1, ri1, r2c, g.ld_res2, ri2, rsub, true);
                                }
                                const bool has_next = c0 + 64 < BN && col + 64 < g.Ncols;
                                if (g.bias || g.group_bias) {
                                    if (k0 == 1) TGP_FAST_RES(1, true, true, rb_lane, has_next);
                                    else if (k0 == 2) TGP_FAST_RES(2, true, true, rb_lane, has_next);
                                    else if (k0 == 3) TGP_FAST_RES(3, true, true, rb_lane, has_next);
                                    else TGP_FAST_RES(4, true, true, rb_lane, has_next);
                                } else {
                                    if (k0 == 1) TGP_FAST_RES(1, true, false, rb_lane, has_next);
                                    else if (k0 == 2) TGP_FAST_RES(2, true, false, rb_lane, has_next);
                                    else if (k0 == 3) TGP_FAST_RES(3, true, false, rb_lane, has_next);
                                    else TGP_FAST_RES(4, true, false, rb_lane, has_next);
                                }
                                pf = has_next ? ci + 1 : -1;
                            }
                            else if (k0 == 1) TGP_FAST_RES(1, false, true);
                            else if (k0 == 2) TGP_FAST_RES(2, false, true);
                            else if (k0 == 3) TGP_FAST_RES(3, false, true);
                            else TGP_FAST_RES(4, false, true);
#undef TGP_FAST_RES
                            done = true;
                        }
                    } else if (nrows == 32 && (k0 == 1 || k0 == 2 || k0 == 4) && __all_sync(0xffffffffu, kind == k0)) {
                        const int col = n0 + c0 + c4i * 4;
                        float* p_r = reinterpret_cast<float*>((uintptr_t)pp) + (long)rsub * rs;
                        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 bias = g.bias ? ldg128(g.bias + col) : zero4;
                        if (has_act) {
                            const float s0 = g.relu ? 0.f : 1.f;
                            const float4 sc = g.scale ? ldg128(g.scale + col) : make_float4(1.f, 1.f, 1.f, 1.f);
                            const float4 sh = g.scale ? ldg128(g.shift + col) : zero4;
                            const float4 sl = g.neg_slope ? ldg128(g.neg_slope + col) : make_float4(s0, s0, s0, s0);
                            epi_fast_kind<true>(k0, ld_even, ld_odd, p_r, rs, lo, rel, bias, sc, sh, sl, NB);
                        } else {
                            epi_fast_kind<false>(k0, ld_even, ld_odd, p_r, rs, lo, rel, bias, zero4, zero4, zero4, NB);
                        }
                        done = true;
                    }
                }
                if (done) {
                } else if (path == 2) {
                    epi_chunk_vec<true>(g, stg_addr, lane, row0, nrows, n0 + c0, grp0, gb_switch, z * zstride, dbg, 0, 0, NB);
                } else if (path == 1) {
                    epi_chunk_vec<false>(g, stg_addr, lane, row0, nrows, n0 + c0, grp0, gb_switch, z * zstride, 0, ri1, ri2, NB);
                } else {
                    const int col = n0 + c0 + lane;
                    const bool live = col < g.Ncols && nrows > 0 && !(dbg & 1);
                    // per-column constants of this lane
                    float bias = 0.f, sc = 1.f, sh = 0.f, slope = g.relu ? 0.f : 1.f, gbv0 = 0.f, gbv1 = 0.f;
                    EpiDst d0 = {nullptr, 0, 0, 0, 0}, d1 = {nullptr, 0, 0, 0, 0};
                    const float* r1p = nullptr;
                    const float* r2p = nullptr;
                    if (live) {
                        if (g.bias) bias = __ldg(g.bias + col);
                        if (g.scale) { sc = __ldg(g.scale + col); sh = __ldg(g.shift + col); }
                        if (g.neg_slope) slope = __ldg(g.neg_slope + col);
    #pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            if (s < g.nseg && col >= g.seg[s].col_begin && col < g.seg[s].col_end) {
                                const int rel = col - g.seg[s].col_begin;
                                EpiDst d;
                                d.mix = 0; d.rel = 0;
                                if (g.seg[s].mode == 3) {
                                    // per-group column max: rows_per_group >= 32, so a 32-row block touches <= 2 groups
                                    d.rs = g.seg[s].col_end - g.seg[s].col_begin;
                                    d.lo = -1;
                                    d.p = g.seg[s].ptr + grp0 * d.rs + rel;
                                } else if (g.seg[s].mode == 1) {
                                    const int w = g.seg[s].slab_width;
                                    const int cg = rel / w, rr = rel - cg * w;
                                    d.rs = w; d.lo = 0;
                                    d.p = g.seg[s].ptr + ((long)cg * g.M + row0) * w + rr;
                                } else {
                                    d.rs = g.seg[s].ld;
                                    d.p = g.seg[s].ptr + z * zstride + row0 * d.rs + rel;
                                    d.lo = g.seg[s].mode == 2 ? g.seg[s].slab_width : 0;
                                    if (g.seg[s].mode >= 4) { d.mix = g.seg[s].slab_width; d.rel = rel; }
                                }
                                if (!d0.p) d0 = d; else d1 = d;     // at most two destinations per column (raw + split)
                            }
                        }
                        if (g.group_bias && !gb_slow) {
                            gbv0 = __ldg(g.group_bias + grp0 * g.Ncols + col);
                            if (gb_switch < nrows) gbv1 = __ldg(g.group_bias + (grp0 + 1) * g.Ncols + col);
                        }
                        if (g.res1) r1p = g.res1 + row0 * g.ld_res1 + col;
                        if (g.res2) r2p = g.res2 + row0 * g.ld_res2 + col;
                    }
                    if (gb_slow) {
                        // tiny groups (< 32 rows per cloud): fold the per-row group bias into the staged values first
                        if (live)
                            for (int rr = 0; rr < nrows; ++rr) {
                                const uint32_t ad = stg_addr + (uint32_t)((rr * 33 + lane) * 4);
                                sts_f32(ad, lds_f32(ad) + __ldg(g.group_bias + ((row0 + rr) / g.rows_per_group) * g.Ncols + col));
                            }
                    }
                    const int nr = live ? nrows : 0;
                    if (__any_sync(0xffffffffu, d0.p != nullptr && d0.lo < 0))
                        epi_rows<true, true>(stg_addr, lane, nr, bias, sc, sh, slope, gbv0, gbv1, gb_switch, r1p, g.ld_res1, r2p,
                                             g.ld_res2, d0, d1, NB);
                    else if (__any_sync(0xffffffffu, d1.p != nullptr))
                        epi_rows<true, false>(stg_addr, lane, nr, bias, sc, sh, slope, gbv0, gbv1, gb_switch, r1p, g.ld_res1, r2p,
                                       g.ld_res2, d0, d1, NB);
                    else
                        epi_rows<false, false>(stg_addr, lane, nr, bias, sc, sh, slope, gbv0, gbv1, gb_switch, r1p, g.ld_res1, r2p,
                                        g.ld_res2, d0, d1, NB);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// operand split: dst[r, 0:Kp) = tf32(src[r,:]) ; dst[r, Kp:2Kp) = src - hi ; zero padding beyond K.
// src_is_kn: src is stored (K, rows) (HS_layer.weights, gcn3d.py:125) and is transposed on the fly.
__global__ void split_tf32_kernel(const float* __restrict__ src, long rows, int K, long ld, int src_is_kn, int Kp,
                                  float* __restrict__ dst) {
    // one CTA per row chunk: no per-element division, coalesced along K
    for (long r = blockIdx.x; r < rows; r += gridDim.x) {
        float* d = dst + r * 2 * Kp;
        for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
            float v = 0.f;
            if (k < K) v = src_is_kn ? __ldg(src + (long)k * ld + r) : __ldg(src + r * ld + k);
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float hi = __uint_as_float(hb);
            d[k] = hi;
            d[Kp + k] = v - hi;
        }
    }
}

// MIXED operand of a row-major (rows, K) matrix: [fp16(x) | bf16(x) | bf16(x - fp16(x)) | unused], zero padded to Kp (multiple of 64)
__global__ void split_mixed_kernel(const float* __restrict__ src, long rows, int K, long ld, int Kp, float* __restrict__ dst) {
    for (long r = blockIdx.x; r < rows; r += gridDim.x) {
        float* d = dst + r * 2 * Kp;
        for (int k = threadIdx.x; k < Kp; k += blockDim.x) store_mixed(d + k, Kp, k, k < K ? __ldg(src + r * ld + k) : 0.f);
    }
}

// 128-bit variant (K % 4 == 0, 16-byte aligned rows): one float4 load and three 8-byte stores per 4 elements instead of
// twelve 2-byte stores -- the scalar kernel ran at a third of the HBM rate on the training step's activations
__global__ void __launch_bounds__(256)
split_mixed_vec_kernel(const float* __restrict__ src, long rows, int K, long ld, int Kp, float* __restrict__ dst) {
    for (long r = blockIdx.x; r < rows; r += gridDim.x) {
        uint16_t* d = reinterpret_cast<uint16_t*>(dst + r * 2 * Kp);
        const float4* s4 = reinterpret_cast<const float4*>(src + r * ld);
        for (int k = threadIdx.x * 4; k < Kp; k += blockDim.x * 4)
            mixed_store4(d, Kp, k, k < K ? __ldg(s4 + (k >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

// transposing split for large (K, rows) sources (activations as weight-gradient operands): 32x32 tiles through
// shared memory, reads coalesced along rows, writes coalesced along K.
__global__ void __launch_bounds__(256)
split_tf32_transpose_kernel(const float* __restrict__ src, long rows, long K, long ld, long Kp, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const long k0 = (long)blockIdx.x * 32, r0 = (long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const long k = k0 + i, r = r0 + tx;
        tile[i][tx] = (k < K && r < rows) ? __ldg(src + k * ld + r) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const long r = r0 + i, k = k0 + tx;
        if (r < rows && k < Kp) {
            const float v = tile[tx][i];
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float hi = __uint_as_float(hb);
            dst[r * 2 * Kp + k] = hi;
            dst[r * 2 * Kp + Kp + k] = v - hi;
        }
    }
}

// K-BLOCKED transposed [tf32 | residual] split of a row-major (Kdim, rows) matrix -- the 3xTF32 operand of the encoder's
// weight-gradient contractions: floats [part hi|lo][block of 32 source rows][R = ceil256(rows)][32]; a TMA tile is one
// contiguous piece (see split_mixed_transpose_kernel for why).
__global__ void __launch_bounds__(256)
split_tf32_t_blocked_kernel(const float* __restrict__ src, long Kdim, long rows, long ld, long nblk, long R, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const long kb = blockIdx.x, k0 = kb * 32, r0 = (long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const long k = k0 + i, r = r0 + tx;
        tile[i][tx] = (k < Kdim && r < rows) ? __ldg(src + k * ld + r) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const long r = r0 + i;
        if (r < rows) {
            const float v = tile[tx][i];
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float hi = __uint_as_float(hb);
            dst[(kb * R + r) * 32 + tx] = hi;
            dst[((nblk + kb) * R + r) * 32 + tx] = v - hi;
        }
    }
}

// out[r, c] = sum_z partial[z][r][c]  (fixed order: deterministic)
__global__ void tc_splitk_reduce_kernel(const float* __restrict__ partial, int nsplit, long rows, int cols,
                                        float* __restrict__ out, long ldo) {
    const long total = rows * cols;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[(long)sp * total + e];
        const long r = e / cols;
        out[r * ldo + (e - r * cols)] = s;
    }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_split_kpad(int K) { return (K + TC_BK - 1) / TC_BK * TC_BK; }

// TRANSPOSED mixed operand of a row-major (rows, K) matrix, K-BLOCKED: the operand of the weight-gradient contraction
// x^T . dY (contraction over the rows).  Layout: 16-bit slots [part 0..2][row block of 64][R = ceil256(K)][64], i.e. for every
// block of 64 source rows a dense (R x 128 B) matrix per part (fp16 hi | bf16 x | bf16 lo).  A TMA tile (128 or 256 operand
// rows x one 64-wide K block) is then ONE contiguous 16 / 32 KB piece of memory.  (The first version kept each operand row
// contiguous over all source rows: 2 MB between the rows of a tile -- every 128-byte row segment in its own page -- and
// ran the contraction at 120 TFLOP/s against 430 for the forward shape; TLB / DRAM-page thrash.)
// Tile = 64 rows x 32 columns through shared memory; a lane packs two consecutive source rows per 32-bit store, a warp
// writes one full 128-byte operand row.
__global__ void __launch_bounds__(256)
split_mixed_transpose_kernel(const float* __restrict__ src, long rows, int K, long ld, long nblk, long R, float* __restrict__ dst) {
    __shared__ float tile[64][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long blk = blockIdx.y, r0 = blk * 64;
    const int k0 = blockIdx.x * 32;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 8) {
        const long r = r0 + rr + ty;
        tile[rr + ty][tx] = (r < rows && k0 + tx < K) ? __ldg(src + r * ld + k0 + tx) : 0.f;
    }
    __syncthreads();
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
    for (int kk = 0; kk < 32; kk += 8) {
        const int k = k0 + kk + ty;
        if (k >= K) continue;
        const float v0 = tile[2 * tx][kk + ty], v1 = tile[2 * tx + 1][kk + ty];
        float h0, h1;
        const long w = (blk * R + k) * 32 + tx;                   // 32-bit word of (block, operand row k), part 0
        const long part = nblk * R * 32;
        d32[w] = mixed_hi16x2(v0, v1, h0, h1);
        d32[part + w] = bf16x2_bits(v0, v1);
        d32[2 * part + w] = bf16x2_bits(v0 - h0, v1 - h1);
    }
}

// WEIGHT operand of a mixed == 2 contraction: [fp16(w) | bf16(w) | fp16(w - fp16(w)) | unused] -- the residual in fp16 (fp16
// subnormals keep |w - hi - lo| <= 2^-25, i.e. ~2^-20 of a weight of size 0.03), so that the third pass can pair it with the
// activations' fp16 slot and the activations need no bf16(x) slot at all
__global__ void __launch_bounds__(256)
split_mixed_w16_kernel(const float* __restrict__ src, long rows, int K, long ld, int Kp, float* __restrict__ dst) {
    for (long r = blockIdx.x; r < rows; r += gridDim.x) {
        uint16_t* d = reinterpret_cast<uint16_t*>(dst + r * 2 * Kp);
        for (int k = threadIdx.x * 2; k < Kp; k += blockDim.x * 2) {
            const float x = k < K ? __ldg(src + r * ld + k) : 0.f, y = k + 1 < K ? __ldg(src + r * ld + k + 1) : 0.f;
            float hx, hy;
            *reinterpret_cast<uint32_t*>(d + k) = mixed_hi16x2(x, y, hx, hy);
            *reinterpret_cast<uint32_t*>(d + Kp + k) = bf16x2_bits(x, y);
            *reinterpret_cast<uint32_t*>(d + 2 * Kp + k) = f16x2_bits(x - hx, y - hy);
        }
    }
}

extern "C" int tgp_split_mixed_w16(const float* src, long rows, int K, long ld, float* dst, tgp_stream_t stream) {
    if (!src || !dst) return fail(TGP_EINVAL, "tgp_split_mixed_w16: null pointer");
    if (rows <= 0 || K <= 0) return fail(TGP_EINVAL, "tgp_split_mixed_w16: sizes must be positive");
    if ((uintptr_t)dst % 16) return fail(TGP_EINVAL, "tgp_split_mixed_w16: dst must be 16-byte aligned");
    const int Kp = (K + 63) / 64 * 64;
    long nb = rows < (long)TGP_NUM_SMS * 64 ? rows : (long)TGP_NUM_SMS * 64;
    split_mixed_w16_kernel<<<(unsigned)nb, 256, 0, as_stream(stream)>>>(src, rows, K, ld, Kp, dst);
    return check_launch("split_mixed_w16_kernel");
}

static long mixed_t_rows(int K) { return ((long)K + 255) / 256 * 256; }

extern "C" size_t tgp_split_mixed_t_bytes(long rows, int K) {
    if (rows <= 0 || K <= 0) return 0;
    return (size_t)3 * ((rows + 63) / 64) * mixed_t_rows(K) * 128;
}

extern "C" int tgp_mixed_kpad(int K) { return (K + 63) / 64 * 64; }

extern "C" int tgp_split_mixed_t(const float* src, long rows, int K, long ld, float* dst, tgp_stream_t stream) {
    if (!src || !dst) return fail(TGP_EINVAL, "tgp_split_mixed_t: null pointer");
    if (rows <= 0 || rows > 0x7fffffffL - 64 || K <= 0) return fail(TGP_EINVAL, "tgp_split_mixed_t: bad sizes");
    if ((uintptr_t)dst % 128) return fail(TGP_EINVAL, "tgp_split_mixed_t: dst must be 128-byte aligned");
    const long nblk = (rows + 63) / 64, R = mixed_t_rows(K);
    if (3 * nblk * R > 0x7fffffffL) return fail(TGP_EINVAL, "tgp_split_mixed_t: operand too large for one tensor map");
    dim3 grid((unsigned)((K + 31) / 32), (unsigned)nblk);
    if (grid.y > 65535) return fail(TGP_EINVAL, "tgp_split_mixed_t: too many rows");
    split_mixed_transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, rows, K, ld, nblk, R, dst);
    return check_launch("split_mixed_transpose_kernel");
}

extern "C" int tgp_split_mixed(const float* src, long rows, int K, long ld, float* dst, tgp_stream_t stream) {
    if (!src || !dst) return fail(TGP_EINVAL, "tgp_split_mixed: null pointer");
    if (rows <= 0 || K <= 0) return fail(TGP_EINVAL, "tgp_split_mixed: sizes must be positive");
    if ((uintptr_t)dst % 16) return fail(TGP_EINVAL, "tgp_split_mixed: dst must be 16-byte aligned");
    const int Kp = tgp_mixed_kpad(K);
    long nb = rows < (long)TGP_NUM_SMS * 64 ? rows : (long)TGP_NUM_SMS * 64;
    if (K % 4 == 0 && ld % 4 == 0 && (uintptr_t)src % 16 == 0) {
        split_mixed_vec_kernel<<<(unsigned)nb, Kp >= 1024 ? 256 : 64, 0, as_stream(stream)>>>(src, rows, K, ld, Kp, dst);
        return check_launch("split_mixed_vec_kernel");
    }
    split_mixed_kernel<<<(unsigned)nb, 256, 0, as_stream(stream)>>>(src, rows, K, ld, Kp, dst);
    return check_launch("split_mixed_kernel");
}

extern "C" size_t tgp_split_tf32_t_bytes(long Kdim, int rows) {
    if (Kdim <= 0 || rows <= 0) return 0;
    return (size_t)2 * ((Kdim + 31) / 32) * (((long)rows + 255) / 256 * 256) * 128;
}

extern "C" int tgp_split_tf32_t(const float* src, long Kdim, int rows, long ld, float* dst, tgp_stream_t stream) {
    if (!src || !dst) return fail(TGP_EINVAL, "tgp_split_tf32_t: null pointer");
    if (Kdim <= 0 || Kdim > 0x7fffffffL - 64 || rows <= 0) return fail(TGP_EINVAL, "tgp_split_tf32_t: bad sizes");
    if ((uintptr_t)dst % 128) return fail(TGP_EINVAL, "tgp_split_tf32_t: dst must be 128-byte aligned");
    const long nblk = (Kdim + 31) / 32, R = ((long)rows + 255) / 256 * 256;
    if (2 * nblk * R > 0x7fffffffL) return fail(TGP_EINVAL, "tgp_split_tf32_t: operand too large for one tensor map");
    dim3 grid((unsigned)nblk, (unsigned)((rows + 31) / 32));
    if (grid.y > 65535) return fail(TGP_EINVAL, "tgp_split_tf32_t: too many rows");
    split_tf32_t_blocked_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, Kdim, rows, ld, nblk, R, dst);
    return check_launch("split_tf32_t_blocked_kernel");
}

extern "C" int tgp_split_tf32(const float* src, long rows, int K, long ld, int src_is_kn, float* dst,
                              tgp_stream_t stream) {
    if (!src || !dst) return fail(TGP_EINVAL, "tgp_split_tf32: null pointer");
    if (rows <= 0 || K <= 0) return fail(TGP_EINVAL, "tgp_split_tf32: sizes must be positive");
    if ((uintptr_t)dst % 16) return fail(TGP_EINVAL, "tgp_split_tf32: dst must be 16-byte aligned");
    const int Kp = tgp_split_kpad(K);
    if (src_is_kn && (long)rows * K >= 4096) {
        dim3 grid((unsigned)(Kp / 32), (unsigned)((rows + 31) / 32));
        if (grid.y > 65535) return fail(TGP_EINVAL, "tgp_split_tf32: too many rows for the transposing split");
        split_tf32_transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, rows, K, ld, Kp, dst);
        return check_launch("split_tf32_transpose_kernel");
    }
    const int threads = Kp >= 256 ? 256 : (Kp >= 128 ? 128 : 64);
    long nb = rows < (long)TGP_NUM_SMS * 64 ? rows : (long)TGP_NUM_SMS * 64;
    split_tf32_kernel<<<(unsigned)nb, threads, 0, as_stream(stream)>>>(src, rows, K, ld, src_is_kn, Kp, dst);
    return check_launch("split_tf32_kernel");
}

template <int BN>
static int launch_tc(const tgp_gemm_args* a, cudaStream_t st, int ksplit = 1, int blocked = 0, int kp_a = 0, int kp_b = 0) {
    const int Kp = a->mixed ? tgp_mixed_kpad(a->K) : tgp_split_kpad(a->K);
    CUtensorMap tmA, tmB, tmA16, tmB16;
    int rc = 0;
    if (!a->mixed && blocked) {
        const long nblk = Kp / TC_BK;
        rc = tgp_make_map_f32_blocked(&tmA, a->A_split, 2 * nblk * (((long)a->M + 255) / 256 * 256), TC_BM);
        if (rc) return rc;
        rc = tgp_make_map_f32_blocked(&tmB, a->B_split, 2 * nblk * (((long)a->Ncols + 255) / 256 * 256), BN);
        if (rc) return rc;
    } else if (!a->mixed) {
        rc = tgp_make_map(&tmA, a->A_split, a->M, Kp, TC_BM);
        if (rc) return rc;
        rc = tgp_make_map(&tmB, a->B_split, a->Ncols, Kp, BN);
        if (rc) return rc;
    }
    if (a->mixed) {
        if (blocked == 2) {
            // row-major mixed operands (rows = contraction index a->K, 4*kp 16-bit slots per row), boxes of 64 rows x 64 slots
            rc = tgp_make_map_bf16(&tmA16, a->A_split, a->K, kp_a, 64);
            if (rc) return rc;
            rc = tgp_make_map_bf16(&tmB16, a->B_split, a->K, kp_b, 64);
            if (rc) return rc;
        } else if (blocked) {
            const long nblk = Kp / 64;
            rc = tgp_make_map_bf16_blocked(&tmA16, a->A_split, 3 * nblk * (((long)a->M + 255) / 256 * 256), TC_BM);
            if (rc) return rc;
            rc = tgp_make_map_bf16_blocked(&tmB16, a->B_split, 3 * nblk * (((long)a->Ncols + 255) / 256 * 256), BN);
            if (rc) return rc;
        } else {
            if (a->a_group_cols > 0 && (a->a_group_cols % 256 || a->a_kp < Kp || a->a_kp % 64 || a->K % 64 ||
                                        (long)((a->Ncols + a->a_group_cols - 1) / a->a_group_cols) * Kp > a->a_kp || ksplit != 1))
                return fail(TGP_EINVAL, "tgp_gemm: bad grouped-contraction arguments");
            rc = tgp_make_map_bf16(&tmA16, a->A_split, a->M, a->a_group_cols > 0 ? a->a_kp : Kp, TC_BM);
            if (rc) return rc;
            rc = tgp_make_map_bf16(&tmB16, a->B_split, a->Ncols, Kp, BN);
            if (rc) return rc;
        }
        tmA = tmA16;      // (the fp32 maps are not used by mixed launches)
        tmB = tmB16;
    } else {
        tmA16 = tmA;
        tmB16 = tmB;
    }
    GemmDev P;
    P.a = *a;
    P.blocked = blocked;
    P.kp_a = kp_a;
    P.kp_b = kp_b;
    const int num_m_tiles = (int)((a->M + TC_BM - 1) / TC_BM);
    const int num_n_tiles = (a->Ncols + BN - 1) / BN;
    const int tiles_mn = num_m_tiles * num_n_tiles;
    const int kblocks = a->mixed ? Kp / 64 : Kp / TC_BK;
    const int kb_per = (kblocks + ksplit - 1) / ksplit;
    ksplit = (kblocks + kb_per - 1) / kb_per;          // no empty slices
    const int num_tiles = tiles_mn * ksplit;
    const long zstride = ksplit > 1 ? a->M * (long)a->seg[0].ld : 0;
    const size_t smem = (size_t)TcStages<BN>::value * (TC_A_BYTES + BN * TC_BK * 4) + 1024 + 256 + TC_EPI_WARPS * 32 * 33 * sizeof(float);
    // mixed destinations without the bf16(x) slot (mode 5) select the NB instantiation; the two kinds do not mix in one launch
    bool nb = false, m4 = false;
    for (int s = 0; s < a->nseg; ++s) { nb = nb || a->seg[s].mode == 5; m4 = m4 || a->seg[s].mode == 4; }
    if (nb && m4) return fail(TGP_EINVAL, "tgp_gemm: output modes 4 and 5 cannot be combined in one launch");
    static std::atomic<unsigned long long> attr_set[2];   // one bit per device: function attributes are per device
    if (first_on_device(attr_set[nb ? 1 : 0])) {
        if (nb) cudaFuncSetAttribute(gemm_tc_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else cudaFuncSetAttribute(gemm_tc_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int grid = num_tiles < TGP_NUM_SMS ? num_tiles : TGP_NUM_SMS;
    // 128-bit epilogue: every width / boundary / leading dimension a multiple of 4 floats, every pointer 16-byte aligned
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    int vec_ok = (a->Ncols % 4 == 0);
    if (a->res1) vec_ok = vec_ok && al16(a->res1) && a->ld_res1 % 4 == 0;
    if (a->res2) vec_ok = vec_ok && al16(a->res2) && a->ld_res2 % 4 == 0;
    for (int s = 0; s < a->nseg; ++s) {
        const tgp_out_seg& sg = a->seg[s];
        vec_ok = vec_ok && sg.col_begin % 4 == 0 && sg.col_end % 4 == 0;
        if (sg.mode == 3) continue;
        vec_ok = vec_ok && al16(sg.ptr);
        if (sg.mode == 1) vec_ok = vec_ok && sg.slab_width % 4 == 0;
        else vec_ok = vec_ok && sg.ld % 4 == 0 && ((sg.mode != 2 && sg.mode < 4) || sg.slab_width % 4 == 0);
    }
    if (vec_ok && !a->res1 && !a->res2 && !a->group_bias) {
        // lean epilogue: one destination per column (segments disjoint; raw / slab / split / mixed / column max)
        bool lean = true;
        for (int s = 0; s < a->nseg && lean; ++s) {
            for (int t = 0; t < s && lean; ++t)
                if (a->seg[s].col_begin < a->seg[t].col_end && a->seg[t].col_begin < a->seg[s].col_end) lean = false;
        }
        { const char* e = getenv("TGP_TC_NO_LEAN"); if (e && e[0] == '1') lean = false; }
        if (lean) {
            vec_ok = 2;
            // fast lean path: the per-column parameter vectors are read with 128-bit loads
            if ((!a->bias || al16(a->bias)) && (!a->scale || (al16(a->scale) && al16(a->shift))) && (!a->neg_slope || al16(a->neg_slope))) {
                bool small = true;
                for (int s = 0; s < a->nseg; ++s) small = small && a->seg[s].ld < (1L << 31);
                const char* e = getenv("TGP_TC_NO_FAST");
                if (small && !(e && e[0] == '1')) vec_ok = 3;
            }
        }
    }
    if (vec_ok == 1) {
        // fast chunk with residuals / per-cloud bias: one destination per column, aligned parameter vectors, groups >= 32 rows
        bool ok = (a->rows_per_group == 0 || a->rows_per_group >= 32) && (!a->group_bias || al16(a->group_bias));
        ok = ok && (!a->bias || al16(a->bias)) && (!a->scale || (al16(a->scale) && al16(a->shift))) && (!a->neg_slope || al16(a->neg_slope));
        for (int s = 0; s < a->nseg && ok; ++s) {
            ok = ok && a->seg[s].ld < (1L << 31);
            for (int t = 0; t < s && ok; ++t)
                if (a->seg[s].col_begin < a->seg[t].col_end && a->seg[t].col_begin < a->seg[s].col_end) ok = false;
        }
        const char* e = getenv("TGP_TC_NO_FAST");
        if (ok && !(e && e[0] == '1')) {
            vec_ok = 4;
            // 256-column tiles: residual pieces prefetched into the shared memory of ring stages 2..3 (the ring runs 2 deep)
            const char* e2 = getenv("TGP_TC_NO_RESBUF");
            if (BN == 256 && (a->res1 || a->res2) && TcStages<BN>::value >= 4 && !(e2 && e2[0] == '1')) vec_ok = 5;
        }
    }
    if ((a->res1_idx || a->res2_idx) && (!vec_ok || (a->rows_per_group > 0 && a->rows_per_group < 32 && a->group_bias)))
        return fail(TGP_EINVAL, "tgp_gemm: gathered residuals need the 128-bit epilogue (widths / leading dimensions multiples of 4, 16-byte aligned pointers)");
    if ((a->res1_idx && !a->res1) || (a->res2_idx && !a->res2)) return fail(TGP_EINVAL, "tgp_gemm: res*_idx without res*");
    { const char* e = getenv("TGP_TC_SCALAR_EPI"); if (e && e[0] == '1') vec_ok = 0; }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TGP_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    int nst = TcStages<BN>::value;      // operand ring depth in use (<= the depth the shared memory is carved for)
    { const char* e = getenv("TGP_TC_NST"); if (e && atoi(e) >= 2 && atoi(e) < nst) nst = atoi(e); }
    if (vec_ok == 5) nst = 2;
    if (nb) gemm_tc_kernel<BN, true><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmA16, tmB16, P, Kp, num_n_tiles, num_tiles, dbg, tiles_mn, kb_per, zstride, vec_ok, nst);
    else gemm_tc_kernel<BN, false><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmA16, tmB16, P, Kp, num_n_tiles, num_tiles, dbg, tiles_mn, kb_per, zstride, vec_ok, nst);
    return check_launch("gemm_tc_kernel");
}


// tensor-core path: both operands pre-split ([hi | lo], tgp_split_tf32)
int tgp_gemm_tc(const tgp_gemm_args* a, cudaStream_t st) {
    if ((uintptr_t)a->A_split % 16 || (uintptr_t)a->B_split % 16)
        return fail(TGP_EINVAL, "tgp_gemm: split operands must be 16-byte aligned");
    if (a->M > 0x7fffffffL - 256) return fail(TGP_EINVAL, "tgp_gemm: M exceeds the tensor-core path's 32-bit row coordinates");
    // widest column block that still gives every SM a tile (small problems: more, narrower tiles)
    const long mt = (a->M + TC_BM - 1) / TC_BM;
    int bn = a->Ncols > 128 ? 256 : (a->Ncols > 64 ? 128 : 64);
    while (bn > 64 && mt * ((a->Ncols + bn - 1) / bn) < TGP_NUM_SMS) bn >>= 1;
    if (bn == 256) return launch_tc<256>(a, st);
    if (bn == 128) return launch_tc<128>(a, st);
    return launch_tc<64>(a, st);
}

// ---------------------------------------------------------------------------------------------------------
// weight-gradient contraction out (K1,K2) = A^T B over M rows on the tensor cores: both operands arrive as
// TRANSPOSED splits (tgp_split_tf32 with src_is_kn): At_split (K1, 2*Mp), Bt_split (K2, 2*Mp).  The long
// contraction is cut into split-K slices (one TMEM accumulator each, partials in `workspace`) that a second
// kernel adds in a fixed order.
static int tn_plan(long M, int K1, int K2, int* bn_out, int mixed = 0) {
    const long mt = (K1 + TC_BM - 1) / TC_BM;
    int bn = K2 > 128 ? 256 : (K2 > 64 ? 128 : 64);
    const long tiles = mt * ((K2 + bn - 1) / bn);
    const int kblocks = mixed ? tgp_mixed_kpad((int)M) / 64 : tgp_split_kpad((int)M) / TC_BK;
    long ks = (2L * TGP_NUM_SMS + tiles - 1) / tiles;
    const long cap = kblocks / 16 > 0 ? kblocks / 16 : 1;     // at least 16 K blocks (x3 passes) per slice
    if (ks > cap) ks = cap;
    if (ks < 1) ks = 1;
    if (ks > 512) ks = 512;
    const int kb_per = (int)((kblocks + ks - 1) / ks);
    ks = (kblocks + kb_per - 1) / kb_per;
    *bn_out = bn;
    return (int)ks;
}

extern "C" size_t tgp_gemm_tn_tc_workspace(long M, int K1, int K2) {
    if (M <= 0 || K1 <= 0 || K2 <= 0) return 0;
    int bn;
    const int ks = tn_plan(M, K1, K2, &bn), ksm = tn_plan(M, K1, K2, &bn, 1);
    return (size_t)(ks > ksm ? ks : ksm) * K1 * K2 * sizeof(float);
}

extern "C" int tgp_gemm_tn_tc(const float* At_split, const float* Bt_split, long M, int K1, int K2, float* out,
                              long ldo, int mixed, void* workspace, size_t workspace_bytes, tgp_stream_t stream) {
    if (!At_split || !Bt_split || !out || !workspace) return fail(TGP_EINVAL, "tgp_gemm_tn_tc: null pointer");
    if (M <= 0 || M > 0x7fffffffL - 64 || K1 <= 0 || K2 <= 0) return fail(TGP_EINVAL, "tgp_gemm_tn_tc: bad sizes");
    if ((uintptr_t)At_split % 16 || (uintptr_t)Bt_split % 16) return fail(TGP_EINVAL, "tgp_gemm_tn_tc: operands must be 16-byte aligned");
    if (workspace_bytes < tgp_gemm_tn_tc_workspace(M, K1, K2)) return fail(TGP_ENOSPACE, "tgp_gemm_tn_tc: workspace too small");
    int bn;
    const int ks = tn_plan(M, K1, K2, &bn, mixed);
    tgp_gemm_args a = {};
    a.mixed = mixed ? 1 : 0;
    a.A_split = At_split;
    a.B_split = Bt_split;
    a.M = K1;
    a.K = (int)M;
    a.Ncols = K2;
    a.nseg = 1;
    a.seg[0].col_begin = 0;
    a.seg[0].col_end = K2;
    a.seg[0].mode = 0;
    a.seg[0].ld = ks > 1 ? K2 : ldo;
    a.seg[0].ptr = ks > 1 ? static_cast<float*>(workspace) : out;
    cudaStream_t st = as_stream(stream);
    int rc;
    if (bn == 256) rc = launch_tc<256>(&a, st, ks, 1);
    else if (bn == 128) rc = launch_tc<128>(&a, st, ks, 1);
    else rc = launch_tc<64>(&a, st, ks, 1);
    if (rc || ks == 1) return rc;
    const long total = (long)K1 * K2;
    long nb = (total + 255) / 256;
    if (nb > TGP_NUM_SMS * 8) nb = TGP_NUM_SMS * 8;
    tc_splitk_reduce_kernel<<<(unsigned)nb, 256, 0, st>>>(static_cast<const float*>(workspace), ks, K1, K2, out, ldo);
    return check_launch("tc_splitk_reduce_kernel");
}

// The same contraction on ROW-MAJOR mixed operands (tgp_split_mixed / epilogue mode 4 / tgp_affine_act(mixed)), read in place:
// A (M, K1) and B (M, K2) both have the contraction index M as their slow dimension, i.e. they are MN-major operands of the
// MMA (tcgen05 instruction-descriptor bits a_major / b_major).  No transposing split: the forward pass's operand of x and the
// backward pass's operand of dY (which the dX contraction needs anyway) are all this launch reads.
extern "C" int tgp_gemm_tn_tc_rm(const float* A_mixed, int kp_a, const float* B_mixed, int kp_b, long M, int K1, int K2,
                                 float* out, long ldo, void* workspace, size_t workspace_bytes, tgp_stream_t stream) {
    if (!A_mixed || !B_mixed || !out || !workspace) return fail(TGP_EINVAL, "tgp_gemm_tn_tc_rm: null pointer");
    if (M <= 0 || M > 0x7fffffffL - 64 || K1 <= 0 || K2 <= 0) return fail(TGP_EINVAL, "tgp_gemm_tn_tc_rm: bad sizes");
    if (kp_a < K1 || kp_b < K2 || kp_a % 64 || kp_b % 64) return fail(TGP_EINVAL, "tgp_gemm_tn_tc_rm: kp must be a multiple of 64 and >= the operand width");
    if ((uintptr_t)A_mixed % 16 || (uintptr_t)B_mixed % 16) return fail(TGP_EINVAL, "tgp_gemm_tn_tc_rm: operands must be 16-byte aligned");
    if (workspace_bytes < tgp_gemm_tn_tc_workspace(M, K1, K2)) return fail(TGP_ENOSPACE, "tgp_gemm_tn_tc_rm: workspace too small");
    int bn;
    const int ks = tn_plan(M, K1, K2, &bn, 1);
    tgp_gemm_args a = {};
    a.mixed = 1;
    a.A_split = A_mixed;
    a.B_split = B_mixed;
    a.M = K1;
    a.K = (int)M;
    a.Ncols = K2;
    a.nseg = 1;
    a.seg[0].col_begin = 0;
    a.seg[0].col_end = K2;
    a.seg[0].mode = 0;
    a.seg[0].ld = ks > 1 ? K2 : ldo;
    a.seg[0].ptr = ks > 1 ? static_cast<float*>(workspace) : out;
    cudaStream_t st = as_stream(stream);
    int rc;
    if (bn == 256) rc = launch_tc<256>(&a, st, ks, 2, kp_a, kp_b);
    else if (bn == 128) rc = launch_tc<128>(&a, st, ks, 2, kp_a, kp_b);
    else rc = launch_tc<64>(&a, st, ks, 2, kp_a, kp_b);
    if (rc || ks == 1) return rc;
    const long total = (long)K1 * K2;
    long nb = (total + 255) / 256;
    if (nb > TGP_NUM_SMS * 8) nb = TGP_NUM_SMS * 8;
    tc_splitk_reduce_kernel<<<(unsigned)nb, 256, 0, st>>>(static_cast<const float*>(workspace), ks, K1, K2, out, ldo);
    return check_launch("tc_splitk_reduce_kernel");
}
