// gemm_simt.cu -- fp32 FMA-pipe contraction with the fused epilogue of the 3D-GCN path.
//
// Replaces `feature_map @ self.weights + self.bias` (gcn3d.py:170) and the 1x1 Conv1d layers
// STE_layer / conv2 (gcn3d.py:70-71,130,132,148,185).  Exact-fp32 path: used for small or
// oddly shaped problems (K = 3, tiny M) and as the numerical reference of the tcgen05 path
// in gemm_tc.cu, which takes the large projections.
//
// Epilogue (all optional): + bias[col] + group_bias[row / rows_per_group, col] (the ORL
// cloud-global term) + res1 + res2, per-column affine (eval BatchNorm), ReLU;
// each column range goes to its own destination, either row-major or the channel-group
// "slab" layout [cgroup][row][S*4] that layer_conv_kernel bulk-copies into shared memory.
#include "common.cuh"

namespace tgp {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_THREADS = 256;
constexpr int GM_LD = GM_BM + 4;

struct GemmDev {
    tgp_gemm_args a;
};

__device__ __forceinline__ float4 ld4_guard(const float* base, long row, long nrows, long ld, int k0, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows) {
        const float* p = base + row * ld + k0;
        if (vec && k0 + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
            if (k0 + 0 < K) v.x = __ldg(p);
            if (k0 + 1 < K) v.y = __ldg(p + 1);
            if (k0 + 2 < K) v.z = __ldg(p + 2);
            if (k0 + 3 < K) v.w = __ldg(p + 3);
        }
    }
    return v;
}

__device__ __forceinline__ void epilogue_store(const tgp_gemm_args& g, long row, int col, float v) {
    if (g.bias) v += __ldg(g.bias + col);
    if (g.group_bias) v += __ldg(g.group_bias + (row / g.rows_per_group) * g.Ncols + col);
    if (g.res1) v += __ldg(g.res1 + (g.res1_idx ? (long)__ldg(g.res1_idx + row) : row) * g.ld_res1 + col);
    if (g.res2) v += __ldg(g.res2 + (g.res2_idx ? (long)__ldg(g.res2_idx + row) : row) * g.ld_res2 + col);
    if (g.scale) v = fmaf(v, __ldg(g.scale + col), __ldg(g.shift + col));
    if (g.neg_slope) v = v > 0.f ? v : v * __ldg(g.neg_slope + col);
    else if (g.relu) v = fmaxf(v, 0.f);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (s < g.nseg && col >= g.seg[s].col_begin && col < g.seg[s].col_end) {
            const int rel = col - g.seg[s].col_begin;
            if (g.seg[s].mode == 3) {
                const int i = __float_as_int(v);
                atomicMax(reinterpret_cast<int*>(g.seg[s].ptr) + (row / g.rows_per_group) * (g.seg[s].col_end - g.seg[s].col_begin) + rel,
                          i >= 0 ? i : i ^ 0x7fffffff);
            } else if (g.seg[s].mode == 0) g.seg[s].ptr[row * g.seg[s].ld + rel] = v;
            else if (g.seg[s].mode == 2) {
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
                const float hi = __uint_as_float(hb);
                g.seg[s].ptr[row * g.seg[s].ld + rel] = hi;
                g.seg[s].ptr[row * g.seg[s].ld + g.seg[s].slab_width + rel] = v - hi;
            } else {
                const int w = g.seg[s].slab_width;
                const int cg = rel / w, r = rel - cg * w;
                g.seg[s].ptr[((long)cg * g.M + row) * w + r] = v;
            }
        }
    }
}

template <bool B_NK>
__global__ void __launch_bounds__(GM_THREADS)
gemm_simt_kernel(const __grid_constant__ GemmDev P) {
    __shared__ __align__(16) float As[GM_BK * GM_LD];
    __shared__ __align__(16) float Bs[GM_BK * GM_LD];
    const tgp_gemm_args& g = P.a;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long m0 = (long)blockIdx.x * GM_BM;
    const int n0 = blockIdx.y * GM_BN;
    const bool vecA = (g.lda % 4 == 0) && ((uintptr_t)g.A % 16 == 0);
    const bool vecB = (g.ldb % 4 == 0) && ((uintptr_t)g.Bmat % 16 == 0);

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += GM_BK) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
            const float4 v = ld4_guard(g.A, m0 + row, g.M, g.lda, k0 + kq, g.K, vecA);
            As[(kq + 0) * GM_LD + row] = v.x; As[(kq + 1) * GM_LD + row] = v.y;
            As[(kq + 2) * GM_LD + row] = v.z; As[(kq + 3) * GM_LD + row] = v.w;
        }
        if (B_NK) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
                const float4 v = ld4_guard(g.Bmat, n0 + row, g.Ncols, g.ldb, k0 + kq, g.K, vecB);
                Bs[(kq + 0) * GM_LD + row] = v.x; Bs[(kq + 1) * GM_LD + row] = v.y;
                Bs[(kq + 2) * GM_LD + row] = v.z; Bs[(kq + 3) * GM_LD + row] = v.w;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kr = (tid >> 5) + h * 8, c4 = (tid & 31) * 4;
                // (K, Ncols) N-contiguous: row index is k, "K extent" is Ncols
                const float4 v = ld4_guard(g.Bmat, k0 + kr, g.K, g.ldb, n0 + c4, g.Ncols, vecB);
                *reinterpret_cast<float4*>(Bs + kr * GM_LD + c4) = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GM_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + kk * GM_LD + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(As + kk * GM_LD + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * GM_LD + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(Bs + kk * GM_LD + 64 + tx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long row = m0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        if (row >= g.M) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int col = n0 + (c < 4 ? tx * 4 + c : 64 + tx * 4 + (c - 4));
            if (col < g.Ncols) epilogue_store(g, row, col, acc[r][c]);
        }
    }
}

// skinny contraction (M <= 64, one row per cloud: the ORL cloud-global term g @ W2b^T, SURVEY 8a a8, and the per-cloud
// head tails).  HBM/L2-bound on the weights: every weight row is streamed exactly once, 512 B per warp load.
//   CTA  = 32 rows of A x 16 output columns; warp = 16 rows x 4 columns (64 accumulators), lanes split K (4 consecutive k
//          each per 128-wide chunk); the A chunk (32 x 128) is staged in shared memory once per CTA and read back as
//          conflict-free LDS.128; weights and the next A chunk are prefetched one chunk ahead in registers;
//   end  = butterfly transpose-reduction of the 64 lane-partials (62 shuffles), fused epilogue.
// A row's result never depends on the other rows or on M (fixed k partition), so a cloud's output is batch-independent.
constexpr int SK_KC = 128;

__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const __grid_constant__ GemmDev P) {
    __shared__ __align__(16) float As[2][32 * SK_KC];
    const tgp_gemm_args& g = P.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rh = warp & 1, cgp = warp >> 1;
    const long m0 = (long)blockIdx.y * 32;
    const int n0 = blockIdx.x * 16 + cgp * 4;
    const bool vecA = (g.lda % 4 == 0) && ((uintptr_t)g.A % 16 == 0);
    const bool vecB = (g.ldb % 4 == 0) && ((uintptr_t)g.Bmat % 16 == 0);
    const int arow = tid >> 3, akq = (tid & 7) * 16;       // staging role: 8 threads per A row, 16 floats each

    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    float4 wn[4], an[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) wn[c] = ld4_guard(g.Bmat, n0 + c, g.Ncols, g.ldb, lane * 4, g.K, vecB);
#pragma unroll
    for (int j = 0; j < 4; ++j) an[j] = ld4_guard(g.A, m0 + arow, g.M, g.lda, akq + j * 4, g.K, vecA);
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += SK_KC, buf ^= 1) {
        float* as = As[buf];
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(as + arow * SK_KC + akq + j * 4) = an[j];
        float4 w[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = wn[c];
        __syncthreads();                                    // chunk visible; the other buffer is free again
        if (k0 + SK_KC < g.K) {
#pragma unroll
            for (int c = 0; c < 4; ++c) wn[c] = ld4_guard(g.Bmat, n0 + c, g.Ncols, g.ldb, k0 + SK_KC + lane * 4, g.K, vecB);
#pragma unroll
            for (int j = 0; j < 4; ++j) an[j] = ld4_guard(g.A, m0 + arow, g.M, g.lda, k0 + SK_KC + akq + j * 4, g.K, vecA);
        }
        const float* ar = as + rh * 16 * SK_KC + lane * 4;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(ar + r * SK_KC);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = acc[r * 4 + c];
                v = fmaf(a.x, w[c].x, v); v = fmaf(a.y, w[c].y, v); v = fmaf(a.z, w[c].z, v); v = fmaf(a.w, w[c].w, v);
                acc[r * 4 + c] = v;
            }
        }
    }
    // transpose-reduce: after the step with offset `off` a lane keeps the upper half of its values iff (lane & off)
#pragma unroll
    for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
            const float send = up ? acc[i] : acc[i + cnt];
            const float keep = up ? acc[i + cnt] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    const int base = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = base + i;
        const long row = m0 + rh * 16 + (idx >> 2);
        const int col = n0 + (idx & 3);
        if (row < g.M && col < g.Ncols) epilogue_store(g, row, col, acc[i]);
    }
}

// Same tiling with a 4-deep cp.async ring for both operands (16-byte aligned rows only): one chunk of register prefetch
// (~0.3 us of work) does not cover the DRAM latency of the weight stream, three chunks in flight do.
constexpr int SKP_STAGES = 4;
constexpr int SKP_STAGE_FLOATS = (32 + 16) * SK_KC;      // A chunk 32 x 128, W chunk 16 x 128

__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(256)
gemm_skinny_pipe_kernel(const __grid_constant__ GemmDev P) {
    extern __shared__ __align__(16) float sk_smem[];
    const tgp_gemm_args& g = P.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rh = warp & 1, cgp = warp >> 1;
    const long m0 = (long)blockIdx.y * 32;
    const int nb = blockIdx.x * 16;
    const int nchunks = (g.K + SK_KC - 1) / SK_KC;

    // staging roles: A piece j of thread t -> row t/8, floats (t%8)*16 + 4j; W piece j -> column (t + 256 j)/32, floats ((t + 256 j)%32)*4
    auto issue = [&](int chunk) {
        float* st = sk_smem + (chunk % SKP_STAGES) * SKP_STAGE_FLOATS;
        const int k0 = chunk * SK_KC;
        const int arow = tid >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kq = (tid & 7) * 16 + j * 4, k = k0 + kq;
            const bool ok = m0 + arow < g.M && k < g.K;
            const int bytes = ok ? min(16, (g.K - k) * 4) : 0;
            cp_async16(st + arow * SK_KC + kq, ok ? g.A + (m0 + arow) * g.lda + k : g.A, bytes);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int p = tid + 256 * j, col = p >> 5, kq = (p & 31) * 4, k = k0 + kq;
            const bool ok = nb + col < g.Ncols && k < g.K;
            const int bytes = ok ? min(16, (g.K - k) * 4) : 0;
            cp_async16(st + 32 * SK_KC + col * SK_KC + kq, ok ? g.Bmat + (long)(nb + col) * g.ldb + k : g.Bmat, bytes);
        }
    };

    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < SKP_STAGES - 1; ++c) {
        if (c < nchunks) issue(c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int c = 0; c < nchunks; ++c) {
        asm volatile("cp.async.wait_group %0;" ::"n"(SKP_STAGES - 2) : "memory");
        __syncthreads();                                    // chunk c has landed for everyone; the stage of chunk c-1 is free
        if (c + SKP_STAGES - 1 < nchunks) issue(c + SKP_STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float* st = sk_smem + (c % SKP_STAGES) * SKP_STAGE_FLOATS;
        float4 w[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) w[cc] = *reinterpret_cast<const float4*>(st + 32 * SK_KC + (cgp * 4 + cc) * SK_KC + lane * 4);
        const float* ar = st + rh * 16 * SK_KC + lane * 4;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(ar + r * SK_KC);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float v = acc[r * 4 + cc];
                v = fmaf(a.x, w[cc].x, v); v = fmaf(a.y, w[cc].y, v); v = fmaf(a.z, w[cc].z, v); v = fmaf(a.w, w[cc].w, v);
                acc[r * 4 + cc] = v;
            }
        }
    }
#pragma unroll
    for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
            const float send = up ? acc[i] : acc[i + cnt];
            const float keep = up ? acc[i + cnt] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    const int base = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = base + i;
        const long row = m0 + rh * 16 + (idx >> 2);
        const int col = nb + cgp * 4 + (idx & 3);
        if (row < g.M && col < g.Ncols) epilogue_store(g, row, col, acc[i]);
    }
}

// tiny-K / tiny-N contraction (K = 3: STE of the surface layer, gcn3d.py:70,84; N = 3: recon head): one thread
// per output element.
template <bool B_NK>
__global__ void gemm_naive_kernel(const __grid_constant__ GemmDev P, long total) {
    const tgp_gemm_args& g = P.a;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long m = e / g.Ncols;
    const int n = (int)(e - m * g.Ncols);
    const float* a = g.A + m * g.lda;
    float acc = 0.f;
    if (B_NK) {
        const float* w = g.Bmat + (long)n * g.ldb;
        for (int k = 0; k < g.K; ++k) acc = fmaf(__ldg(a + k), __ldg(w + k), acc);
    } else {
        for (int k = 0; k < g.K; ++k) acc = fmaf(__ldg(a + k), __ldg(g.Bmat + (long)k * g.ldb + n), acc);
    }
    epilogue_store(g, m, n, acc);
}

// tiny-N contraction with K-contiguous weight rows (N <= 8, e.g. the recon head's 128 -> 3): a warp per row, lanes
// split K with 128-bit loads of the row (read once, coalesced) and of the <= 8 weight rows (L1-resident), then one
// warp reduction per output.
__global__ void __launch_bounds__(256)
gemm_smalln_kernel(const __grid_constant__ GemmDev P) {
    const tgp_gemm_args& g = P.a;
    const int lane = threadIdx.x & 31;
    const long m = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= g.M) return;
    const bool vecA = (g.lda % 4 == 0) && ((uintptr_t)g.A % 16 == 0);
    const bool vecB = (g.ldb % 4 == 0) && ((uintptr_t)g.Bmat % 16 == 0);
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    for (int k0 = lane * 4; k0 < g.K; k0 += 128) {
        const float4 a = ld4_guard(g.A, m, g.M, g.lda, k0, g.K, vecA);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            if (n < g.Ncols) {
                const float4 w = ld4_guard(g.Bmat, n, g.Ncols, g.ldb, k0, g.K, vecB);
                acc[n] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[n]))));
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (n < g.Ncols) {
            const float v = warp_sum(acc[n]);
            if (lane == n) epilogue_store(g, m, n, v);
        }
    }
}

__global__ void decode_max_kernel(const int* __restrict__ enc, long n, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int e = enc[i];
        out[i] = __int_as_float(e >= 0 ? e : e ^ 0x7fffffff);
    }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_decode_max(const int32_t* enc, long n, float* out, tgp_stream_t stream) {
    if (!enc || !out || n <= 0) return fail(TGP_EINVAL, "tgp_decode_max: bad arguments");
    decode_max_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(enc, n, out);
    return check_launch("decode_max_kernel");
}

int tgp_gemm_validate(const tgp_gemm_args* a) {
    if (!a) return fail(TGP_EINVAL, "tgp_gemm: null args");
    if ((!a->A || !a->Bmat) && (!a->A_split || !a->B_split)) return fail(TGP_EINVAL, "tgp_gemm: null operand");
    if (a->M <= 0 || a->K <= 0 || a->Ncols <= 0) return fail(TGP_EINVAL, "tgp_gemm: sizes must be positive");
    if (a->a_group_cols != 0 && (!a->mixed || !a->A_split || !a->B_split)) return fail(TGP_EINVAL, "tgp_gemm: a grouped contraction needs mixed operands");
    if (a->nseg < 1 || a->nseg > 4) return fail(TGP_EINVAL, "tgp_gemm: nseg must be 1..4");
    if (a->rows_per_group < 0) return fail(TGP_EINVAL, "tgp_gemm: rows_per_group must be >= 0");
    if (a->group_bias && a->rows_per_group <= 0) return fail(TGP_EINVAL, "tgp_gemm: rows_per_group must be positive");
    if ((a->scale == nullptr) != (a->shift == nullptr)) return fail(TGP_EINVAL, "tgp_gemm: scale and shift go together");
    if ((a->res1_idx && !a->res1) || (a->res2_idx && !a->res2)) return fail(TGP_EINVAL, "tgp_gemm: res*_idx without res*");
    for (int s = 0; s < a->nseg; ++s) {
        const tgp_out_seg& sg = a->seg[s];
        if (!sg.ptr || sg.col_begin < 0 || sg.col_end > a->Ncols || sg.col_begin >= sg.col_end)
            return fail(TGP_EINVAL, "tgp_gemm: bad output segment");
        if (sg.mode == 1 && (sg.slab_width <= 0 || (sg.col_end - sg.col_begin) % sg.slab_width))
            return fail(TGP_EINVAL, "tgp_gemm: slab segment must be a multiple of slab_width");
        if (sg.mode < 0 || sg.mode > 5) return fail(TGP_EINVAL, "tgp_gemm: bad segment mode");
        if (sg.mode >= 4 && (sg.slab_width % 64 || sg.slab_width < sg.col_end - sg.col_begin || sg.ld != 2L * sg.slab_width))
            return fail(TGP_EINVAL, "tgp_gemm: mixed segment needs slab_width = tgp_mixed_kpad(width) and ld = 2*slab_width");
        if (sg.mode >= 4 && !(a->A_split && a->B_split)) return fail(TGP_EINVAL, "tgp_gemm: mixed output is written by the tensor-core path only");
        if (sg.mode == 3 && a->rows_per_group < 32) return fail(TGP_EINVAL, "tgp_gemm: column-max segment needs rows_per_group >= 32");
        if (sg.mode == 2 && sg.slab_width < sg.col_end - sg.col_begin) return fail(TGP_EINVAL, "tgp_gemm: split segment wider than Kp");
    }
    return TGP_OK;
}

int tgp_gemm_simt(const tgp_gemm_args* a, cudaStream_t st) {
    GemmDev P;
    P.a = *a;
    if (a->M <= 64 && a->b_is_nk && a->K >= 32) {
        const dim3 grid((a->Ncols + 15) / 16, (unsigned)((a->M + 31) / 32));
        // (the summation order over k is the same in both kernels: the choice only depends on operand alignment)
        if (a->lda % 4 == 0 && a->ldb % 4 == 0 && (uintptr_t)a->A % 16 == 0 && (uintptr_t)a->Bmat % 16 == 0) {
            const size_t smem = (size_t)SKP_STAGES * SKP_STAGE_FLOATS * sizeof(float);
            static std::atomic<unsigned long long> attr_set{0};   // one bit per device: function attributes are per device
            if (first_on_device(attr_set)) {
                cudaFuncSetAttribute(gemm_skinny_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            }
            gemm_skinny_pipe_kernel<<<grid, 256, smem, st>>>(P);
            return check_launch("gemm_skinny_pipe_kernel");
        }
        gemm_skinny_kernel<<<grid, 256, 0, st>>>(P);
        return check_launch("gemm_skinny_kernel");
    }
    if (a->Ncols <= 8 && a->b_is_nk && a->K >= 32) {
        gemm_smalln_kernel<<<(unsigned)((a->M + 7) / 8), 256, 0, st>>>(P);
        return check_launch("gemm_smalln_kernel");
    }
    if (a->K <= 8 || a->Ncols <= 8) {
        const long total = a->M * a->Ncols;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        if (a->b_is_nk) gemm_naive_kernel<true><<<blocks, 256, 0, st>>>(P, total);
        else gemm_naive_kernel<false><<<blocks, 256, 0, st>>>(P, total);
        return check_launch("gemm_naive_kernel");
    }
    dim3 grid((unsigned)((a->M + GM_BM - 1) / GM_BM), (a->Ncols + GM_BN - 1) / GM_BN);
    if (a->b_is_nk) gemm_simt_kernel<true><<<grid, GM_THREADS, 0, st>>>(P);
    else gemm_simt_kernel<false><<<grid, GM_THREADS, 0, st>>>(P);
    return check_launch("gemm_simt_kernel");
}
