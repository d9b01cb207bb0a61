// knn.cu -- k-nearest-neighbour index kernels for sm_100a.
//
// Replaces get_neighbor_index / get_nearest_index (reference network/fs_net_repo/gcn3d.py:14-35):
// the reference materialises the (B,N,N) distance matrix with torch.bmm, two broadcast adds
// and torch.topk; here distances are produced tile by tile in shared memory / registers and
// the top-(k+1) list of a query lives in the registers of one warp (one or two ranks per
// lane, kept sorted with shuffles).  Nothing of size N*N ever reaches HBM.
//
// Ordering contract: ascending by (distance, index); rank 0 is dropped
// positionally (gcn3d.py:22), never by testing j == i.
#include "knn_select.cuh"
#include <stdlib.h>

namespace tgp {

// ------------------------------------------------------------------------------------------
// xyz-space kNN.  Distance recipe reproduces CPU torch bit for bit:
//   q = (x0*x0 + x1*x1) + x2*x2 ; inner = fma(a2,b2, fma(a1,b1, a0*b0)) ; d = ((inner*-2)+q_j)+q_i
// Candidates of one cloud are staged once per CTA as (x,y,z,q) float4 in shared memory; each
// warp owns a query and scans 32 candidates per step (conflict-free LDS.128).
constexpr int KNN_QPC = 32;        // queries per CTA (4 per warp): enough CTAs to fill 148 SMs at small batch
constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE_MAX = 4096; // candidates resident in shared memory at a time

template <int SLOTS>
__global__ void __launch_bounds__(KNN_THREADS)
knn_xyz_kernel(const float* __restrict__ xyz, int N, int k, int tile, int64_t* __restrict__ idx64,
               int32_t* __restrict__ idx32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* pts = reinterpret_cast<float4*>(smem_raw);
    float* st_d = reinterpret_cast<float*>(pts + tile);          // [KNN_QPC][32*SLOTS], only if multi-tile
    int* st_i = reinterpret_cast<int*>(st_d + KNN_QPC * 32 * SLOTS);

    const int b = blockIdx.y;
    const int q0 = blockIdx.x * KNN_QPC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* cloud = xyz + (size_t)b * N * 3;
    const int K = k + 1;
    const bool multi = N > tile;

    for (int t0 = 0; t0 < N; t0 += tile) {
        const int nt = min(tile, N - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += KNN_THREADS) {
            const float x0 = cloud[(t0 + j) * 3], x1 = cloud[(t0 + j) * 3 + 1], x2 = cloud[(t0 + j) * 3 + 2];
            const float q = __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2));
            pts[j] = make_float4(x0, x1, x2, q);
        }
        __syncthreads();
        for (int ql = warp; ql < KNN_QPC; ql += KNN_THREADS / 32) {
            const int qi = q0 + ql;
            if (qi >= N) break;
            const float a0 = __ldg(cloud + qi * 3), a1 = __ldg(cloud + qi * 3 + 1), a2 = __ldg(cloud + qi * 3 + 2);
            const float qq = __fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2));
            WarpTopList<SLOTS> top;
            if (t0 == 0) top.init();
            else {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    top.d[s] = st_d[(ql * SLOTS + s) * 32 + lane];
                    top.i[s] = st_i[(ql * SLOTS + s) * 32 + lane];
                }
            }
            float th = top.thresh(K);
            for (int j0 = 0; j0 < nt; j0 += 32) {
                const int j = j0 + lane;
                float dd = CUDART_INF_F;
                if (j < nt) {
                    const float4 p = pts[j];
                    const float inner = __fmaf_rn(a2, p.z, __fmaf_rn(a1, p.y, __fmul_rn(a0, p.x)));
                    dd = __fadd_rn(__fadd_rn(__fmul_rn(inner, -2.0f), p.w), qq);
                }
                top.admit(dd, t0 + j0, lane, th, K);
            }
            if (t0 + nt >= N) {
                if (idx64) top.store_ranks(idx64 + ((size_t)b * N + qi) * k, k, lane);
                if (idx32) top.store_ranks(idx32 + ((size_t)b * N + qi) * k, k, lane);
            } else if (multi) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    st_d[(ql * SLOTS + s) * 32 + lane] = top.d[s];
                    st_i[(ql * SLOTS + s) * 32 + lane] = top.i[s];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// xyz-space kNN by threshold selection (knn_select.cuh), k + 1 <= 32: each warp owns 4 queries and advances them
// together, so one LDS.128 of a candidate feeds four distance evaluations.  Pass 1 keeps two strided minima per
// lane and query (64 groups), pass 2 recomputes the (bit-identical) distances and compacts the ~K + 4 survivors
// with warp ballots, then every survivor ranks itself.  ~6x fewer instructions per query than the insertion list.
constexpr int KS_QPW = 4;                                  // queries per warp
#ifndef KS_THREADS
#define KS_THREADS 128
#endif
constexpr int KS_QPC = KS_QPW * (KS_THREADS / 32);         // queries per CTA
constexpr int KS_SLOTS_MAX = 8;                            // survivor slots per lane and query (lane-private lists):
constexpr int KS_CAP = 128;                                //   6 for k <= 21, 8 above; compacted survivors per query
constexpr int KS_TILE_MAX = 8192;                          // candidates resident in shared memory at a time
static inline int ks_slots(int k) { return k + 1 <= 22 ? 6 : KS_SLOTS_MAX; }
static inline int ks_warp_smem(int k) { return (KS_QPW * ks_slots(k) * 32 + KS_CAP) * 8; }

__device__ __forceinline__ float xyz_dist(float a0, float a1, float a2, float qq, const float4& p) {
    const float inner = __fmaf_rn(a2, p.z, __fmaf_rn(a1, p.y, __fmul_rn(a0, p.x)));
    return __fadd_rn(__fadd_rn(__fmul_rn(inner, -2.0f), p.w), qq);
}

__device__ __forceinline__ void stage_xyz_tile(float4* pts, const float* cloud, int t0, int nt, int ntp) {
    for (int j = threadIdx.x; j < ntp; j += KS_THREADS) {
        float4 v = make_float4(0.f, 0.f, 0.f, CUDART_INF_F);          // padding: distance +inf
        if (j < nt) {
            const float x0 = cloud[(t0 + j) * 3], x1 = cloud[(t0 + j) * 3 + 1], x2 = cloud[(t0 + j) * 3 + 2];
            v = make_float4(x0, x1, x2, __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
        }
        pts[j] = v;
    }
}

// G4 (k + 1 in 33..64): 128 group minima (four per lane) instead of 64 -- the K-th smallest of them still bounds the
// K-th distance, and ~ -128 ln(1 - K/128) candidates (65 at K = 51, 89 at K = 64) survive it
template <bool G4>
__global__ void __launch_bounds__(KS_THREADS)
knn_xyz_sel_kernel(const float* __restrict__ xyz, int N, int k, int tile, int slots, int64_t* __restrict__ idx64,
                   int32_t* __restrict__ idx32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* pts = reinterpret_cast<float4*>(smem_raw);                                  // [tile] (tile % 32 == 0)
    const int qstride = slots * 256;                                                    // bytes of one query's lists
    unsigned char* wsm = reinterpret_cast<unsigned char*>(pts + tile) + (size_t)(threadIdx.x >> 5) * (KS_QPW * qstride + KS_CAP * 8);
    uint2* compact = reinterpret_cast<uint2*>(wsm + KS_QPW * qstride);                  // [KS_CAP] (key, index)

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* cloud = xyz + (size_t)b * N * 3;
    const int K = k + 1;
    const int qbase = blockIdx.x * KS_QPC + warp * KS_QPW;
    const bool single = N <= tile;

    float a0[KS_QPW], a1[KS_QPW], a2[KS_QPW], qq[KS_QPW], m0[KS_QPW], m1[KS_QPW], m2[KS_QPW], m3[KS_QPW];
#pragma unroll
    for (int q = 0; q < KS_QPW; ++q) {
        const int qi = min(qbase + q, N - 1);              // clamped: idle warps still walk the barriers
        a0[q] = __ldg(cloud + qi * 3); a1[q] = __ldg(cloud + qi * 3 + 1); a2[q] = __ldg(cloud + qi * 3 + 2);
        qq[q] = __fadd_rn(__fadd_rn(__fmul_rn(a0[q], a0[q]), __fmul_rn(a1[q], a1[q])), __fmul_rn(a2[q], a2[q]));
        m0[q] = CUDART_INF_F; m1[q] = CUDART_INF_F; m2[q] = CUDART_INF_F; m3[q] = CUDART_INF_F;
    }
    // ---- pass 1: 64 strided group minima per query
    for (int t0 = 0; t0 < N; t0 += tile) {
        const int nt = min(tile, N - t0), ntp = (nt + 63) & ~63;
        if (t0) __syncthreads();
        stage_xyz_tile(pts, cloud, t0, nt, ntp);
        __syncthreads();
#pragma unroll 2
        for (int j0 = 0; j0 < ntp; j0 += 64) {
            const float4 p = pts[j0 + lane], r = pts[j0 + 32 + lane];
            const bool odd = G4 && (j0 & 64);              // G4: blocks of 64 alternate between the minima pairs (0,1) and (2,3)
#pragma unroll
            for (int q = 0; q < KS_QPW; ++q) {
                const float dp = xyz_dist(a0[q], a1[q], a2[q], qq[q], p), dr = xyz_dist(a0[q], a1[q], a2[q], qq[q], r);
                if (odd) { m2[q] = fminf(m2[q], dp); m3[q] = fminf(m3[q], dr); }
                else { m0[q] = fminf(m0[q], dp); m1[q] = fminf(m1[q], dr); }
            }
        }
    }
    float T[KS_QPW];
#pragma unroll
    for (int q = 0; q < KS_QPW; ++q)
        T[q] = fminf(G4 ? warp_kth_of_128(m0[q], m1[q], m2[q], m3[q], K, lane) : warp_kth_of_64(m0[q], m1[q], K, lane), 3.0e38f);
    // ---- pass 2: every lane appends its survivors (d <= T) to a private list: no ballots in the scan
    // (the write cursor of each list is a byte offset that advances by one 256-byte row per survivor; predicated, no branch)
    uint32_t cur[KS_QPW];
    const uint32_t wsm_s = (uint32_t)__cvta_generic_to_shared(wsm);
#pragma unroll
    for (int q = 0; q < KS_QPW; ++q) cur[q] = wsm_s + (uint32_t)(q * qstride + lane * 8);
    for (int t0 = 0; t0 < N; t0 += tile) {
        const int nt = min(tile, N - t0), ntp = (nt + 63) & ~63;
        if (!single) {
            __syncthreads();
            stage_xyz_tile(pts, cloud, t0, nt, ntp);
            __syncthreads();
        }
#pragma unroll 2
        for (int j0 = 0; j0 < ntp; j0 += 32) {
            const float4 p = pts[j0 + lane];
            const float jf = __int_as_float(t0 + j0 + lane);
#pragma unroll
            for (int q = 0; q < KS_QPW; ++q) {
                const float d = xyz_dist(a0[q], a1[q], a2[q], qq[q], p);
                const bool pass = d <= T[q];
                const uint32_t lim = wsm_s + (uint32_t)((q + 1) * qstride);
                const uint32_t st = (pass && cur[q] < lim) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.v2.f32 [%1], {%2, %3};\n\t}"
                             :: "r"(st), "r"(cur[q]), "f"(d), "f"(jf) : "memory");
                cur[q] += pass ? 256u : 0u;
            }
        }
    }
    int cnt[KS_QPW];
#pragma unroll
    for (int q = 0; q < KS_QPW; ++q) cnt[q] = (int)((cur[q] - wsm_s - (uint32_t)(q * qstride + lane * 8)) >> 8);
    // ---- compaction + ranks, one query at a time
    bool slow = false;
    bool redo[KS_QPW];
#pragma unroll
    for (int q = 0; q < KS_QPW; ++q) {
        const int qi = qbase + q;
        int incl = cnt[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        const bool over = __any_sync(FULL, cnt[q] > slots);
        redo[q] = qi < N && (over || total < K || total > KS_CAP);
        slow |= redo[q];
        if (qi >= N || redo[q]) continue;
        const int off = incl - cnt[q];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < KS_SLOTS_MAX; ++s)
            if (s < cnt[q]) {
                const float2 e = *reinterpret_cast<const float2*>(wsm + q * qstride + s * 256 + lane * 8);
                compact[off + s] = make_uint2(ordered_key(e.x), (uint32_t)__float_as_int(e.y));
            }
        __syncwarp();
        const size_t o = ((size_t)b * N + qi) * k;
        warp_rank_store<KS_CAP / 32>(compact, total, k, lane, idx64 ? idx64 + o : nullptr, idx32 ? idx32 + o : nullptr);
    }
    // ---- slow path (massive distance ties, e.g. padded / duplicated points, or non-finite input): insertion list
    if (!single) slow = __syncthreads_or(slow);            // tiles must be restaged by the whole CTA
    if (slow) {
        for (int t0 = 0; t0 < N; t0 += tile) {
            const int nt = min(tile, N - t0), ntp = (nt + 63) & ~63;
            if (!single) {
                __syncthreads();
                stage_xyz_tile(pts, cloud, t0, nt, ntp);
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < KS_QPW; ++q) {
                const int qi = qbase + q;
                if (!redo[q]) continue;
                constexpr int TS = G4 ? 2 : 1;
                WarpTopList<TS> top;
                float* sd = reinterpret_cast<float*>(wsm + q * qstride);            // list parked between tiles
                int* si = reinterpret_cast<int*>(sd + 32 * TS);
                if (t0 == 0) top.init();
                else {
#pragma unroll
                    for (int s_ = 0; s_ < TS; ++s_) { top.d[s_] = sd[s_ * 32 + lane]; top.i[s_] = si[s_ * 32 + lane]; }
                }
                // nothing above T can be among the K nearest: admission is capped just above it from the start
                const float cap = T[q] < 3.0e38f ? nextafterf(T[q], CUDART_INF_F) : CUDART_INF_F;
                float th = fminf(top.thresh(K), cap);
                for (int j0 = 0; j0 < ntp; j0 += 32)
                    top.admit(xyz_dist(a0[q], a1[q], a2[q], qq[q], pts[j0 + lane]), t0 + j0, lane, th, K, cap);
                if (t0 + nt >= N) {
#pragma unroll
                    for (int s_ = 0; s_ < TS; ++s_) {
                        const int rank = s_ * 32 + lane;
                        if (rank >= 1 && rank <= k) {
                            const size_t o = ((size_t)b * N + qi) * k + rank - 1;
                            const int v = top.i[s_] < 0 ? 0 : top.i[s_];      // unfilled rank (NaN input): a valid index
                            if (idx64) idx64[o] = v;
                            if (idx32) idx32[o] = v;
                        }
                    }
                } else {
#pragma unroll
                    for (int s_ = 0; s_ < TS; ++s_) { sd[s_ * 32 + lane] = top.d[s_]; si[s_ * 32 + lane] = top.i[s_]; }
                }
                __syncwarp();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// nearest source point: d = (s_j + t_i) - 2*inner (gcn3d.py:33, a different op order from kNN),
// strict '<' so the lowest index wins.  One thread per target, sources staged as (x,y,z,|s|^2).
constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 2048;

__global__ void __launch_bounds__(NN_THREADS)
nearest_kernel(const float* __restrict__ tgt, const float* __restrict__ src, int N, int M,
               int64_t* __restrict__ idx64, int32_t* __restrict__ idx32) {
    __shared__ float4 sp[NN_TILE];
    const int b = blockIdx.y;
    const int i = blockIdx.x * NN_THREADS + threadIdx.x;
    const float* sb = src + (size_t)b * M * 3;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    if (i < N) {
        const float* t = tgt + ((size_t)b * N + i) * 3;
        t0 = t[0]; t1 = t[1]; t2 = t[2];
    }
    const float tn = __fadd_rn(__fadd_rn(__fmul_rn(t0, t0), __fmul_rn(t1, t1)), __fmul_rn(t2, t2));
    float best = CUDART_INF_F;
    int bj = 0;
    for (int m0 = 0; m0 < M; m0 += NN_TILE) {
        const int nt = min(NN_TILE, M - m0);
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += NN_THREADS) {
            const float x0 = sb[(m0 + j) * 3], x1 = sb[(m0 + j) * 3 + 1], x2 = sb[(m0 + j) * 3 + 2];
            sp[j] = make_float4(x0, x1, x2,
                                __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nt; ++j) {
            const float4 p = sp[j];
            const float inner = __fmaf_rn(t2, p.z, __fmaf_rn(t1, p.y, __fmul_rn(t0, p.x)));
            const float d = __fsub_rn(__fadd_rn(p.w, tn), __fmul_rn(2.0f, inner));
            if (d < best) { best = d; bj = m0 + j; }
        }
    }
    if (i < N) {
        if (idx64) idx64[(size_t)b * N + i] = bj;
        if (idx32) idx32[(size_t)b * N + i] = bj;
    }
}

// ------------------------------------------------------------------------------------------
// feature-space kNN (RF-F, gcn3d.py:201-206): same expanded formula with D = 128/256.
// A CTA owns 64 queries; candidate tiles of 128 points are contracted against them with a
// register-tiled fp32 FMA micro-kernel (4x8 per thread), the 64x128 distance tile is parked
// in shared memory and the 8 warps run the same register top-list selection over its rows.
constexpr int KF_BM = 64, KF_BN = 128, KF_BK = 16, KF_THREADS = 256;
constexpr int KF_LDA = KF_BM + 4, KF_LDB = KF_BN + 4, KF_LDD = KF_BN + 4;

__global__ void rownorm_kernel(const float* __restrict__ x, long rows, int D, float* __restrict__ q) {
    const long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* p = x + r * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(p[c], p[c], s);
    s = warp_sum(s);
    if (lane == 0) q[r] = s;
}

template <int SLOTS>
__global__ void __launch_bounds__(KF_THREADS)
knn_feat_kernel(const float* __restrict__ x, const float* __restrict__ qn, int N, int D, int k,
                int64_t* __restrict__ idx64, int32_t* __restrict__ idx32,
                const int* __restrict__ unit_list, int q_tiles, int q_step) {
    // unit_list != nullptr: fix-up launch behind the tensor-core kernel for k > 31 -- only query tiles that overlap a
    // listed unit ([0] = count; unit = cloud * q_tiles + tile of q_step queries) are redone (normally none)
    if (unit_list) {
        const int cntu = unit_list[0];
        const int qa = blockIdx.x * KF_BM, qe = qa + KF_BM;
        bool hit = false;
        for (int i = 0; i < cntu; ++i) {
            const int u = unit_list[1 + i];
            const int u0 = (u % q_tiles) * q_step;
            hit |= (u / q_tiles == (int)blockIdx.y) && u0 < qe && u0 + q_step > qa;
        }
        if (!hit) return;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);              // [KF_BK][KF_LDA]
    float* Bs = As + KF_BK * KF_LDA;                             // [KF_BK][KF_LDB]
    float* Ds = Bs + KF_BK * KF_LDB;                             // [KF_BM][KF_LDD]
    float* st_d = Ds + KF_BM * KF_LDD;                           // [KF_BM][32*SLOTS]
    int* st_i = reinterpret_cast<int*>(st_d + KF_BM * 32 * SLOTS);

    const int b = blockIdx.y;
    const int q0 = blockIdx.x * KF_BM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const float* xb = x + (size_t)b * N * D;
    const float* qb = qn + (size_t)b * N;
    const int K = k + 1;
    const bool vec = (D & 3) == 0;

    for (int c0 = 0; c0 < N; c0 += KF_BN) {
        float acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

        for (int k0 = 0; k0 < D; k0 += KF_BK) {
            __syncthreads();
            // A chunk: 64 rows x 16 k  (one float4 per thread); B chunk: 128 rows x 16 k (two)
            {
                const int row = tid >> 2, kq = (tid & 3) * 4;
                const int gr = q0 + row;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < N) {
                    const float* p = xb + (size_t)gr * D + k0 + kq;
                    if (vec && k0 + kq + 3 < D) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (k0 + kq + 0 < D) v.x = p[0];
                        if (k0 + kq + 1 < D) v.y = p[1];
                        if (k0 + kq + 2 < D) v.z = p[2];
                        if (k0 + kq + 3 < D) v.w = p[3];
                    }
                }
                As[(kq + 0) * KF_LDA + row] = v.x; As[(kq + 1) * KF_LDA + row] = v.y;
                As[(kq + 2) * KF_LDA + row] = v.z; As[(kq + 3) * KF_LDA + row] = v.w;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
                const int gr = c0 + row;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < N) {
                    const float* p = xb + (size_t)gr * D + k0 + kq;
                    if (vec && k0 + kq + 3 < D) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (k0 + kq + 0 < D) v.x = p[0];
                        if (k0 + kq + 1 < D) v.y = p[1];
                        if (k0 + kq + 2 < D) v.z = p[2];
                        if (k0 + kq + 3 < D) v.w = p[3];
                    }
                }
                Bs[(kq + 0) * KF_LDB + row] = v.x; Bs[(kq + 1) * KF_LDB + row] = v.y;
                Bs[(kq + 2) * KF_LDB + row] = v.z; Bs[(kq + 3) * KF_LDB + row] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KF_BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(As + kk * KF_LDA + ty * 4);
                const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * KF_LDB + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(Bs + kk * KF_LDB + 64 + tx * 4);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
            }
        }
        // distance tile: d = ((inner*-2) + q_j) + q_i   (gcn3d.py:20)
        {
            float qj[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int col = c0 + (c < 4 ? tx * 4 + c : 64 + tx * 4 + (c - 4));
                qj[c] = col < N ? __ldg(qb + col) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int row = q0 + ty * 4 + r;
                const float qi = row < N ? __ldg(qb + row) : 0.f;
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = __fadd_rn(__fadd_rn(__fmul_rn(acc[r][c], -2.0f), qj[c]), qi);
                *reinterpret_cast<float4*>(Ds + (ty * 4 + r) * KF_LDD + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(Ds + (ty * 4 + r) * KF_LDD + 64 + tx * 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
        __syncthreads();
        const int nt = min(KF_BN, N - c0);
        const bool last = c0 + KF_BN >= N;
        for (int ql = warp; ql < KF_BM; ql += KF_THREADS / 32) {
            const int qi = q0 + ql;
            if (qi >= N) break;
            WarpTopList<SLOTS> top;
            if (c0 == 0) top.init();
            else {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    top.d[s] = st_d[(ql * SLOTS + s) * 32 + lane];
                    top.i[s] = st_i[(ql * SLOTS + s) * 32 + lane];
                }
            }
            float th = top.thresh(K);
            const float* drow = Ds + ql * KF_LDD;
#pragma unroll
            for (int j0 = 0; j0 < KF_BN; j0 += 32) {
                const int j = j0 + lane;
                const float dd = j < nt ? drow[j] : CUDART_INF_F;
                top.admit(dd, c0 + j0, lane, th, K);
            }
            if (last) {
                if (idx64) top.store_ranks(idx64 + ((size_t)b * N + qi) * k, k, lane);
                if (idx32) top.store_ranks(idx32 + ((size_t)b * N + qi) * k, k, lane);
            } else {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    st_d[(ql * SLOTS + s) * 32 + lane] = top.d[s];
                    st_i[(ql * SLOTS + s) * 32 + lane] = top.i[s];
                }
            }
        }
    }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_knn_xyz(const float* xyz, int B, int N, int k, int64_t* idx64, int32_t* idx32,
                           tgp_stream_t stream) {
    if (!xyz || (!idx64 && !idx32)) return fail(TGP_EINVAL, "tgp_knn_xyz: null pointer");
    if (B <= 0 || N <= 0 || k <= 0) return fail(TGP_EINVAL, "tgp_knn_xyz: B, N, k must be positive");
    if (k + 1 > N) return fail(TGP_EINVAL, "tgp_knn_xyz: k+1 > N (torch.topk would raise, gcn3d.py:21)");
    if (k + 1 > 64) return fail(TGP_EINVAL, "tgp_knn_xyz: k > 63 unsupported");
    if (B > 65535) return fail(TGP_EINVAL, "tgp_knn_xyz: B > 65535");
    cudaStream_t st = as_stream(stream);
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("TGP_KNN_LEGACY"); legacy = (e && e[0] == '1') ? 1 : 0; }
    if (k + 1 <= 64 && !legacy) {
        // threshold selection; the whole cloud stays resident when it fits, otherwise both passes walk 8192-point tiles
        const int tile = N <= KS_TILE_MAX ? ((N + 63) & ~63) : KS_TILE_MAX;
        const size_t smem = (size_t)tile * sizeof(float4) + (size_t)(KS_THREADS / 32) * ks_warp_smem(k);
        static std::atomic<unsigned long long> attr_set{0};   // one bit per device: function attributes are per device
        if (first_on_device(attr_set)) {
            cudaFuncSetAttribute(knn_xyz_sel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(KS_TILE_MAX * sizeof(float4) + (KS_THREADS / 32) * ks_warp_smem(63)));
            cudaFuncSetAttribute(knn_xyz_sel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(KS_TILE_MAX * sizeof(float4) + (KS_THREADS / 32) * ks_warp_smem(63)));
        }
        dim3 grid((N + KS_QPC - 1) / KS_QPC, B);
        if (k + 1 <= 32) knn_xyz_sel_kernel<false><<<grid, KS_THREADS, smem, st>>>(xyz, N, k, tile, ks_slots(k), idx64, idx32);
        else knn_xyz_sel_kernel<true><<<grid, KS_THREADS, smem, st>>>(xyz, N, k, tile, ks_slots(k), idx64, idx32);
        return check_launch("knn_xyz_sel_kernel");
    }
    const int tile = N < KNN_TILE_MAX ? N : KNN_TILE_MAX;
    const int slots = (k + 1 > 32) ? 2 : 1;
    size_t smem = (size_t)tile * sizeof(float4);
    if (N > tile) smem += (size_t)KNN_QPC * 32 * slots * 8;
    dim3 grid((N + KNN_QPC - 1) / KNN_QPC, B);
    if (slots == 1) {
        cudaFuncSetAttribute(knn_xyz_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        knn_xyz_kernel<1><<<grid, KNN_THREADS, smem, st>>>(xyz, N, k, tile, idx64, idx32);
    } else {
        cudaFuncSetAttribute(knn_xyz_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        knn_xyz_kernel<2><<<grid, KNN_THREADS, smem, st>>>(xyz, N, k, tile, idx64, idx32);
    }
    return check_launch("knn_xyz_kernel");
}

bool tgp_knn_tc_eligible(int B, int N, int D, int k);
size_t tgp_knn_tc_fix_bytes(int B, int N);
int tgp_knn_tc(const float* x_split, const float* qn, int B, int N, int D, int k, int64_t* idx64, int32_t* idx32,
               int* fix_list, cudaStream_t st);

static size_t knn_qn_bytes(int B, int N) { return (((size_t)B * N * sizeof(float)) + 255) & ~(size_t)255; }

// workspace layout: [row norms | fix-up unit list | split operand (unless the caller passes one)]
extern "C" size_t tgp_knn_feat_workspace(int B, int N, int D, int have_split) {
    size_t s = knn_qn_bytes(B, N) + tgp_knn_tc_fix_bytes(B, N);
    if (!have_split) s += (size_t)B * N * 2 * tgp_split_kpad(D) * sizeof(float);
    return s;
}

extern "C" int tgp_knn_feat(const float* x, const float* x_split, int B, int N, int D, int k, int64_t* idx64,
                            int32_t* idx32, void* workspace, size_t workspace_bytes, tgp_stream_t stream) {
    if (!x || (!idx64 && !idx32) || !workspace) return fail(TGP_EINVAL, "tgp_knn_feat: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || D <= 0) return fail(TGP_EINVAL, "tgp_knn_feat: sizes must be positive");
    if (k + 1 > N) return fail(TGP_EINVAL, "tgp_knn_feat: k+1 > N (torch.topk would raise, gcn3d.py:21)");
    if (k + 1 > 64) return fail(TGP_EINVAL, "tgp_knn_feat: k > 63 unsupported");
    if (B > 65535) return fail(TGP_EINVAL, "tgp_knn_feat: B > 65535");
    if (workspace_bytes < tgp_knn_feat_workspace(B, N, D, x_split != nullptr)) return fail(TGP_ENOSPACE, "tgp_knn_feat: workspace too small");
    cudaStream_t st = as_stream(stream);
    float* qn = static_cast<float*>(workspace);
    const long rows = (long)B * N;
    rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, D, qn);
    int rc = check_launch("rownorm_kernel");
    if (rc) return rc;
    static int force_fp32 = -1;
    if (force_fp32 < 0) { const char* e = getenv("TGP_KNN_FP32"); force_fp32 = (e && e[0] == '1') ? 1 : 0; }
    if (!force_fp32 && tgp_knn_tc_eligible(B, N, D, k)) {
        // inner products on the tensor cores (knn_tc.cu); the split operand is built here unless the caller has it
        unsigned char* ws = static_cast<unsigned char*>(workspace);
        if (!x_split) {
            float* spl = reinterpret_cast<float*>(ws + knn_qn_bytes(B, N) + tgp_knn_tc_fix_bytes(B, N));
            rc = tgp_split_tf32(x, rows, D, D, 0, spl, stream);
            if (rc) return rc;
            x_split = spl;
        }
        static int legacy = -1;
        if (legacy < 0) { const char* e = getenv("TGP_KNN_LEGACY"); legacy = (e && e[0] == '1') ? 1 : 0; }
        int* fix = reinterpret_cast<int*>(ws + knn_qn_bytes(B, N));
        if (k + 1 <= 32) return tgp_knn_tc(x_split, qn, B, N, D, k, idx64, idx32, legacy ? nullptr : fix, st);
        // k = 32..63: threshold selection on 128 group minima; units it could not finish (massive ties, non-finite input)
        // are redone by the fp32 tile kernel below, which returns at once when the list is empty
        rc = tgp_knn_tc(x_split, qn, B, N, D, k, idx64, idx32, fix, st);
        if (rc) return rc;
        const int q_tiles = (N + 127) / 128, q_step = (N + q_tiles - 1) / q_tiles;      // unit geometry of tgp_knn_tc
        const size_t smem = sizeof(float) * (KF_BK * KF_LDA + KF_BK * KF_LDB + KF_BM * KF_LDD) + (size_t)KF_BM * 32 * 2 * 8;
        cudaFuncSetAttribute(knn_feat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid((N + KF_BM - 1) / KF_BM, B);
        knn_feat_kernel<2><<<grid, KF_THREADS, smem, st>>>(x, qn, N, D, k, idx64, idx32, fix, q_tiles, q_step);
        return check_launch("knn_feat_kernel (fix-up)");
    }
    const int slots = (k + 1 > 32) ? 2 : 1;
    const size_t smem = sizeof(float) * (KF_BK * KF_LDA + KF_BK * KF_LDB + KF_BM * KF_LDD) + (size_t)KF_BM * 32 * slots * 8;
    dim3 grid((N + KF_BM - 1) / KF_BM, B);
    if (slots == 1) {
        cudaFuncSetAttribute(knn_feat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        knn_feat_kernel<1><<<grid, KF_THREADS, smem, st>>>(x, qn, N, D, k, idx64, idx32, nullptr, 0, 0);
    } else {
        cudaFuncSetAttribute(knn_feat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        knn_feat_kernel<2><<<grid, KF_THREADS, smem, st>>>(x, qn, N, D, k, idx64, idx32, nullptr, 0, 0);
    }
    return check_launch("knn_feat_kernel");
}

extern "C" int tgp_nearest(const float* target, const float* source, int B, int N, int M,
                           int64_t* idx64, int32_t* idx32, tgp_stream_t stream) {
    if (!target || !source || (!idx64 && !idx32)) return fail(TGP_EINVAL, "tgp_nearest: null pointer");
    if (B <= 0 || N <= 0 || M <= 0) return fail(TGP_EINVAL, "tgp_nearest: sizes must be positive");
    if (B > 65535) return fail(TGP_EINVAL, "tgp_nearest: B > 65535");
    dim3 grid((N + NN_THREADS - 1) / NN_THREADS, B);
    nearest_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(target, source, N, M, idx64, idx32);
    return check_launch("nearest_kernel");
}
