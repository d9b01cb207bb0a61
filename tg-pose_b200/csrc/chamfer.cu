// chamfer.cu -- chamfer3D nearest-neighbour distance, forward and backward, for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel (reference losses/chamfer3D/chamfer3D.cu:12-195).
// The reference launches dim3(32,16)x512 twice, re-stages the target cloud in every y-block
// (13 of 16 idle at n ~ 1k) and keeps the running minimum in global memory; here both directions
// are one launch, a CTA owns 128 queries of one (cloud, direction), the running (min, argmin)
// stays in registers and the per-cloud sums calc_cd needs (TDA_loss_sym_recon.py:495-509) are
// reduced in the same kernel.
//
// Arithmetic: d = fma(dz,dz, fma(dy,dy, dx*dx)) with dx = x2 - x1 -- what nvcc's default
// -fmad=true makes of chamfer3D.cu:32-35 -- and strict '<', so the lowest index wins ties.
#include "common.cuh"
#include <math_constants.h>

namespace tgp {

constexpr int CH_THREADS = 128;
constexpr int CH_TILE = 2048;

__global__ void __launch_bounds__(CH_THREADS)
chamfer_fwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, int m,
                   float* __restrict__ dist1, float* __restrict__ dist2, int32_t* __restrict__ idx1,
                   int32_t* __restrict__ idx2, float* __restrict__ sums) {
    __shared__ float4 cand[CH_TILE];
    const int dir = blockIdx.z;
    const long b = blockIdx.y;
    const int nq = dir == 0 ? n : m, nc = dir == 0 ? m : n;
    if ((int)blockIdx.x * CH_THREADS >= nq) return;  // uniform per CTA
    const float* q = (dir == 0 ? xyz1 : xyz2) + b * nq * 3;
    const float* c = (dir == 0 ? xyz2 : xyz1) + b * nc * 3;
    float* dist = (dir == 0 ? dist1 : dist2) + b * nq;
    int32_t* idx = (dir == 0 ? idx1 : idx2) + b * nq;

    const int i = blockIdx.x * CH_THREADS + threadIdx.x;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (i < nq) { x1 = __ldg(q + i * 3); y1 = __ldg(q + i * 3 + 1); z1 = __ldg(q + i * 3 + 2); }
    float best = CUDART_INF_F;
    int bi = 0;
    for (int t0 = 0; t0 < nc; t0 += CH_TILE) {
        const int nt = min(CH_TILE, nc - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += CH_THREADS)
            cand[j] = make_float4(__ldg(c + (t0 + j) * 3), __ldg(c + (t0 + j) * 3 + 1), __ldg(c + (t0 + j) * 3 + 2), 0.f);
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < nt; ++j) {
            const float4 p = cand[j];
            const float dx = p.x - x1, dy = p.y - y1, dz = p.z - z1;
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            if (d < best) { best = d; bi = t0 + j; }
        }
    }
    float s_d = 0.f, s_r = 0.f;
    if (i < nq) {
        dist[i] = best;
        idx[i] = bi;
        s_d = best;
        s_r = sqrtf(best);
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
    const int nmax = n > m ? n : m;
    dim3 grid((nmax + CH_THREADS - 1) / CH_THREADS, B, 2);
    chamfer_fwd_kernel<<<grid, CH_THREADS, 0, as_stream(stream)>>>(xyz1, xyz2, n, m, dist1, dist2, idx1, idx2, sums);
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
