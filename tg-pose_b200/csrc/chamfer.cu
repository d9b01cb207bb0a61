// chamfer.cu -- chamfer3D nearest-neighbour distance, forward and backward, for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel (reference losses/chamfer3D/chamfer3D.cu:12-195).
// The reference launches dim3(32,16)x512 twice, re-stages the target cloud in every y-block
// (13 of 16 idle at n ~ 1k), tracks (min, argmin) with a compare + two selects per pair and keeps
// the running minimum in global memory.  Here both directions are one launch; a thread owns FOUR
// queries (every candidate fetched from shared memory feeds 4 distance evaluations); candidates are
// staged as PAIRS (x0,x1,y0,y1 | z0,z1) so that the 6 arithmetic operations of a distance run as
// packed FADD2 / FMUL2 / FFMA2 on two candidates at once; and the inner loop tracks only the running
// MINIMUM (one 3-input FMNMX per query and candidate pair, on the ALU pipe, beside the FMA pipe)
// plus, per 32-candidate chunk, which chunk last lowered it.  The arg-min is recovered afterwards
// by re-evaluating that one chunk (bit-identical arithmetic) and taking the first candidate whose
// distance equals the minimum -- the lowest index on ties, like the reference's strict '<'.
// The per-cloud sums calc_cd needs (TDA_loss_sym_recon.py:495-509) are reduced in the same kernel.
//
// Arithmetic: d = fma(dz,dz, fma(dx,dx, dy*dy)) with dx = x2 - x1 -- what nvcc 12.9's default
// -fmad=true makes of chamfer3D.cu:32-35 for sm_100a (FMUL on y, FFMA on x, FFMA on z in the SASS of
// oracle/_ref/chamfer3D); tests/test_gpu_ref_chamfer.py holds this kernel bit-equal to that build.
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace tgp {

constexpr int CH_Q = 4;            // queries per thread
constexpr int CH_CHUNK = 8;        // candidates per arg-min chunk (4 pairs)
constexpr int CH_PPC = CH_CHUNK / 2;
constexpr int CH_XY_STRIDE = CH_PPC + 1;   // float4 slots per chunk: one pad slot rotates the banks from chunk to chunk, so the
constexpr int CH_Z_STRIDE = CH_PPC + 2;    // arg-min pass (every lane in a different chunk) spreads over the banks; z: 16-byte aligned
constexpr int CH_TILE = 1152;      // candidates staged per pass (18 KB: the 1024 / 1028-point clouds of the path in one pass)
constexpr int CH_NCHUNK = CH_TILE / CH_CHUNK;
constexpr int CH_MAX_THREADS = 128;

__device__ __forceinline__ unsigned long long ch_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ch_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ch_add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long ch_mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long ch_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ch_min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__global__ void __launch_bounds__(CH_MAX_THREADS)
chamfer_fwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, int m,
                   float* __restrict__ dist1, float* __restrict__ dist2, int32_t* __restrict__ idx1,
                   int32_t* __restrict__ idx2, float* __restrict__ sums) {
    __shared__ __align__(16) float4 cxy[CH_NCHUNK * CH_XY_STRIDE];     // (x0, x1, y0, y1) of candidate pair p
    __shared__ __align__(16) float2 cz[CH_NCHUNK * CH_Z_STRIDE];       // (z0, z1)
    const int dir = blockIdx.z;
    const long b = blockIdx.y;
    const int nq = dir == 0 ? n : m, nc = dir == 0 ? m : n;
    const int qbase = blockIdx.x * (blockDim.x * CH_Q);
    if (qbase >= nq) return;  // uniform per CTA
    const float* q = (dir == 0 ? xyz1 : xyz2) + b * nq * 3;
    const float* c = (dir == 0 ? xyz2 : xyz1) + b * nc * 3;
    float* dist = (dir == 0 ? dist1 : dist2) + b * nq;
    int32_t* idx = (dir == 0 ? idx1 : idx2) + b * nq;

    // query u of this thread: qbase + u * blockDim.x + threadIdx.x (coalesced result stores); x2 - x1 is one packed add
    // of the broadcast (-x1)
    float qx[CH_Q], qy[CH_Q], qz[CH_Q];
    float best[CH_Q], cm[CH_Q];
    int bchunk[CH_Q], bi[CH_Q];
#pragma unroll
    for (int u = 0; u < CH_Q; ++u) {
        const int i = min(qbase + u * (int)blockDim.x + (int)threadIdx.x, nq - 1);
        qx[u] = __ldg(q + i * 3); qy[u] = __ldg(q + i * 3 + 1); qz[u] = __ldg(q + i * 3 + 2);
        best[u] = CUDART_INF_F; cm[u] = CUDART_INF_F; bchunk[u] = -1; bi[u] = 0;
    }
    for (int t0 = 0; t0 < nc; t0 += CH_TILE) {
        const int nt = min(CH_TILE, nc - t0);
        const int nchunk = (nt + CH_CHUNK - 1) / CH_CHUNK;
        __syncthreads();
        // candidates past the end are +inf: their distance is +inf and never lowers a minimum
        for (int p = threadIdx.x; p < nchunk * CH_PPC; p += blockDim.x) {
            float x0 = CUDART_INF_F, y0 = CUDART_INF_F, z0 = CUDART_INF_F, x1 = CUDART_INF_F, y1 = CUDART_INF_F, z1 = CUDART_INF_F;
            if (2 * p < nt) {
                const float* s = c + (long)(t0 + 2 * p) * 3;
                x0 = __ldg(s); y0 = __ldg(s + 1); z0 = __ldg(s + 2);
                if (2 * p + 1 < nt) { x1 = __ldg(s + 3); y1 = __ldg(s + 4); z1 = __ldg(s + 5); }
            }
            const int ch = p / CH_PPC, pp = p - ch * CH_PPC;
            cxy[ch * CH_XY_STRIDE + pp] = make_float4(x0, x1, y0, y1);
            cz[ch * CH_Z_STRIDE + pp] = make_float2(z0, z1);
        }
        __syncthreads();
        unsigned long long nqx[CH_Q], nqy[CH_Q], nqz[CH_Q];
#pragma unroll
        for (int u = 0; u < CH_Q; ++u) {
            nqx[u] = ch_pack(-qx[u], -qx[u]); nqy[u] = ch_pack(-qy[u], -qy[u]); nqz[u] = ch_pack(-qz[u], -qz[u]);
        }
#pragma unroll 2
        for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll
            for (int pp = 0; pp < CH_PPC; ++pp) {
                const float4 a = cxy[ch * CH_XY_STRIDE + pp];
                const float2 zz = cz[ch * CH_Z_STRIDE + pp];
                const unsigned long long ax = ch_pack(a.x, a.y), ay = ch_pack(a.z, a.w), az = ch_pack(zz.x, zz.y);
#pragma unroll
                for (int u = 0; u < CH_Q; ++u) {
                    const unsigned long long dx = ch_add2(ax, nqx[u]), dy = ch_add2(ay, nqy[u]), dz = ch_add2(az, nqz[u]);
                    float d0, d1;
                    ch_unpack(ch_fma2(dz, dz, ch_fma2(dx, dx, ch_mul2(dy, dy))), d0, d1);
                    cm[u] = ch_min3(cm[u], d0, d1);
                }
            }
#pragma unroll
            for (int u = 0; u < CH_Q; ++u) {
                if (cm[u] < best[u]) { best[u] = cm[u]; bchunk[u] = ch; }     // strict: the earliest chunk keeps a tie
                cm[u] = CUDART_INF_F;
            }
        }
        // arg-min of the queries whose minimum fell inside this tile: first candidate of the winning chunk whose
        // (bit-identically recomputed) distance equals the minimum; the chunk is still in shared memory
#pragma unroll
        for (int u = 0; u < CH_Q; ++u) {
            if (bchunk[u] < 0) continue;
            const int ch = bchunk[u];
            int w = CH_CHUNK - 1;
#pragma unroll
            for (int pp = CH_PPC - 1; pp >= 0; --pp) {
                const float4 a = cxy[ch * CH_XY_STRIDE + pp];
                const float2 zz = cz[ch * CH_Z_STRIDE + pp];
                const float dx1 = __fadd_rn(a.y, -qx[u]), dy1 = __fadd_rn(a.w, -qy[u]), dz1 = __fadd_rn(zz.y, -qz[u]);
                const float dx0 = __fadd_rn(a.x, -qx[u]), dy0 = __fadd_rn(a.z, -qy[u]), dz0 = __fadd_rn(zz.x, -qz[u]);
                if (__fmaf_rn(dz1, dz1, __fmaf_rn(dx1, dx1, __fmul_rn(dy1, dy1))) == best[u]) w = 2 * pp + 1;
                if (__fmaf_rn(dz0, dz0, __fmaf_rn(dx0, dx0, __fmul_rn(dy0, dy0))) == best[u]) w = 2 * pp;
            }
            bi[u] = t0 + ch * CH_CHUNK + w;
            bchunk[u] = -1;
        }
    }
    float s_d = 0.f, s_r = 0.f;
#pragma unroll
    for (int u = 0; u < CH_Q; ++u) {
        const int i = qbase + u * (int)blockDim.x + (int)threadIdx.x;
        if (i < nq) {
            // every distance +inf / NaN (no chunk ever lowered the minimum): the reference takes candidate 0 (chamfer3D.cu:36)
            dist[i] = best[u];
            idx[i] = bi[u];
            s_d += best[u];
            s_r += sqrtf(best[u]);
        }
    }
    if (sums) {
        s_d = warp_sum(s_d);
        s_r = warp_sum(s_r);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(sums + b * 4 + dir, s_d);
            atomicAdd(sums + b * 4 + 2 + dir, s_r);
        }
    }
}

// direct terms: gradxyz1[i] = 2*gd1[i]*(x1_i - x2[idx1[i]]), same for side 2 (chamfer3D.cu:158-168)
__global__ void chamfer_bwd_direct_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                          const float* __restrict__ gd1, const float* __restrict__ gd2,
                                          const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                                          long B, int n, int m, float* __restrict__ g1, float* __restrict__ g2) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long tot1 = B * n, tot2 = B * m;
    if (e < tot1) {
        const long b = e / n;
        const float* p = xyz1 + e * 3;
        const float* q = xyz2 + (b * m + __ldg(idx1 + e)) * 3;
        const float g = __ldg(gd1 + e) * 2.f;
        g1[e * 3] = g * (p[0] - q[0]); g1[e * 3 + 1] = g * (p[1] - q[1]); g1[e * 3 + 2] = g * (p[2] - q[2]);
    } else if (e < tot1 + tot2) {
        const long f = e - tot1;
        const long b = f / m;
        const float* p = xyz2 + f * 3;
        const float* q = xyz1 + (b * n + __ldg(idx2 + f)) * 3;
        const float g = __ldg(gd2 + f) * 2.f;
        g2[f * 3] = g * (p[0] - q[0]); g2[f * 3 + 1] = g * (p[1] - q[1]); g2[f * 3 + 2] = g * (p[2] - q[2]);
    }
}

// scattered terms: gradxyz2[idx1[i]] -= term1_i ; gradxyz1[idx2[j]] -= term2_j (chamfer3D.cu:169-171)
__global__ void chamfer_bwd_scatter_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                           const float* __restrict__ gd1, const float* __restrict__ gd2,
                                           const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                                           long B, int n, int m, float* __restrict__ g1, float* __restrict__ g2) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long tot1 = B * n, tot2 = B * m;
    if (e < tot1) {
        const long b = e / n;
        const long j2 = b * m + __ldg(idx1 + e);
        const float* p = xyz1 + e * 3;
        const float* q = xyz2 + j2 * 3;
        const float g = __ldg(gd1 + e) * 2.f;
        atomicAdd(g2 + j2 * 3, -(g * (p[0] - q[0])));
        atomicAdd(g2 + j2 * 3 + 1, -(g * (p[1] - q[1])));
        atomicAdd(g2 + j2 * 3 + 2, -(g * (p[2] - q[2])));
    } else if (e < tot1 + tot2) {
        const long f = e - tot1;
        const long b = f / m;
        const long j2 = b * n + __ldg(idx2 + f);
        const float* p = xyz2 + f * 3;
        const float* q = xyz1 + j2 * 3;
        const float g = __ldg(gd2 + f) * 2.f;
        atomicAdd(g1 + j2 * 3, -(g * (p[0] - q[0])));
        atomicAdd(g1 + j2 * 3 + 1, -(g * (p[1] - q[1])));
        atomicAdd(g1 + j2 * 3 + 2, -(g * (p[2] - q[2])));
    }
}


// ------------------------------------------------------------------------------------------
// density-aware chamfer tail (calc_dcd, losses/TDA_loss_sym_recon.py:411-450, non_reg=False).
// The reference loops over the batch in Python and calls torch.bincount per cloud (B host syncs);
// here one CTA per cloud builds both histograms in shared memory, evaluates
//   loss[b] = mean_i(1 - exp(-a d1_i) w1_i) + 0.5 mean_j(1 - exp(-a d2_j) w2_j),
//   w1_i = (count1[idx1_i]^lambda + 1e-6)^-1 * (m/n),  w2_j = (count2[idx2_j]^lambda + 1e-6)^-1 * (n/m)
// and the coefficients d loss / d dist (the weights are detached in the reference, :433,438).
constexpr int DCD_THREADS = 256;
__global__ void __launch_bounds__(DCD_THREADS)
dcd_kernel(const float* __restrict__ dist1, const float* __restrict__ dist2, const int32_t* __restrict__ idx1,
           const int32_t* __restrict__ idx2, int n, int m, float alpha, float lambda, int non_reg, float* __restrict__ loss,
           float* __restrict__ coef1, float* __restrict__ coef2) {
    extern __shared__ int hist[];            // [m] counts of idx1 values, then [n] counts of idx2 values
    __shared__ float red[2][DCD_THREADS / 32];
    int* c1 = hist;
    int* c2 = hist + m;
    const long b = blockIdx.x;
    for (int i = threadIdx.x; i < n + m; i += DCD_THREADS) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += DCD_THREADS) atomicAdd(c1 + __ldg(idx1 + b * n + i), 1);
    for (int j = threadIdx.x; j < m; j += DCD_THREADS) atomicAdd(c2 + __ldg(idx2 + b * m + j), 1);
    __syncthreads();
    float frac_12 = (float)n / (float)m, frac_21 = (float)m / (float)n;
    if (non_reg) { frac_12 = fmaxf(1.f, frac_12); frac_21 = fmaxf(1.f, frac_21); }   // TDA_loss_sym_recon.py:418-420
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n; i += DCD_THREADS) {
        const float w = frac_21 / (powf((float)c1[__ldg(idx1 + b * n + i)], lambda) + 1e-6f);
        const float e = expf(-__ldg(dist1 + b * n + i) * alpha);
        s1 += 1.f - e * w;
        if (coef1) coef1[b * n + i] = alpha * e * w / (float)n;
    }
    for (int j = threadIdx.x; j < m; j += DCD_THREADS) {
        const float w = frac_12 / (powf((float)c2[__ldg(idx2 + b * m + j)], lambda) + 1e-6f);
        const float e = expf(-__ldg(dist2 + b * m + j) * alpha);
        s2 += 1.f - e * w;
        if (coef2) coef2[b * m + j] = 0.5f * alpha * e * w / (float)m;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int w = 0; w < DCD_THREADS / 32; ++w) { a += red[0][w]; c += red[1][w]; }
        loss[b] = a / (float)n + 0.5f * (c / (float)m);
    }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1, float* dist2,
                               int32_t* idx1, int32_t* idx2, float* sums, tgp_stream_t stream) {
    if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return fail(TGP_EINVAL, "tgp_chamfer_fwd: null pointer");
    if (B <= 0 || n <= 0 || m <= 0) return fail(TGP_EINVAL, "tgp_chamfer_fwd: sizes must be positive");
    if (B > 65535) return fail(TGP_EINVAL, "tgp_chamfer_fwd: B > 65535");
    // threads per CTA: 4 queries each; chosen so that the CTAs of a cloud come out evenly filled (1028 queries ->
    // 3 CTAs of 96 threads rather than 2 full + 1 nearly empty CTA of 128)
    const int nmax = n > m ? n : m;
    int threads = CH_MAX_THREADS;
    {
        double best = -1.0;
        for (int t = CH_MAX_THREADS; t >= 32; t -= 32) {
            const int g = (nmax + t * CH_Q - 1) / (t * CH_Q);
            double eff = (double)nmax / ((double)g * t * CH_Q);
            if ((long)g * B * 2 < 2 * TGP_NUM_SMS) eff *= 0.8;      // too few CTAs to fill the machine: prefer smaller ones
            if (eff > best + 0.03) { best = eff; threads = t; }
        }
    }
    if (const char* e = getenv("TGP_CH_THREADS")) { const int v = atoi(e); if (v >= 32 && v <= CH_MAX_THREADS && v % 32 == 0) threads = v; }   // tuning hook
    dim3 grid((nmax + threads * CH_Q - 1) / (threads * CH_Q), B, 2);
    chamfer_fwd_kernel<<<grid, threads, 0, as_stream(stream)>>>(xyz1, xyz2, n, m, dist1, dist2, idx1, idx2, sums);
    return check_launch("chamfer_fwd_kernel");
}

extern "C" int tgp_chamfer_bwd(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                               const int32_t* idx1, const int32_t* idx2, int B, int n, int m, float* gradxyz1,
                               float* gradxyz2, tgp_stream_t stream) {
    if (!xyz1 || !xyz2 || !graddist1 || !graddist2 || !idx1 || !idx2 || !gradxyz1 || !gradxyz2)
        return fail(TGP_EINVAL, "tgp_chamfer_bwd: null pointer");
    if (B <= 0 || n <= 0 || m <= 0) return fail(TGP_EINVAL, "tgp_chamfer_bwd: sizes must be positive");
    const long total = (long)B * (n + m);
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    cudaStream_t st = as_stream(stream);
    chamfer_bwd_direct_kernel<<<blocks, threads, 0, st>>>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, B, n, m, gradxyz1, gradxyz2);
    int rc = check_launch("chamfer_bwd_direct_kernel");
    if (rc) return rc;
    chamfer_bwd_scatter_kernel<<<blocks, threads, 0, st>>>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, B, n, m, gradxyz1, gradxyz2);
    return check_launch("chamfer_bwd_scatter_kernel");
}

extern "C" int tgp_dcd(const float* dist1, const float* dist2, const int32_t* idx1, const int32_t* idx2, int B, int n,
                       int m, float alpha, float n_lambda, int non_reg, float* loss, float* coef1,
                       float* coef2, tgp_stream_t stream) {
    if (!dist1 || !dist2 || !idx1 || !idx2 || !loss) return fail(TGP_EINVAL, "tgp_dcd: null pointer");
    if (B <= 0 || n <= 0 || m <= 0) return fail(TGP_EINVAL, "tgp_dcd: sizes must be positive");
    const size_t smem = (size_t)(n + m) * sizeof(int);
    if (smem > 200 * 1024) return fail(TGP_EINVAL, "tgp_dcd: n + m too large for the shared-memory histograms");
    static std::atomic<unsigned long long> attr{0};   // one bit per device: function attributes are per device
    if (first_on_device(attr)) {
        cudaFuncSetAttribute(dcd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    dcd_kernel<<<B, DCD_THREADS, smem, as_stream(stream)>>>(dist1, dist2, idx1, idx2, n, m, alpha, n_lambda, non_reg, loss, coef1, coef2);
    return check_launch("dcd_kernel");
}
