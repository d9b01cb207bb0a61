// backward.cu -- backward kernels of the 3D-GCN path (north_star item 5, SURVEY 8a').
//
// The reference has no hand-written backward for gcn3d: autograd differentiates the
// materialised graph (theta (B,N,k,S*C), gathered support, product, max, mean;
// gcn3d.py:91-106,157-180,210-217,225-245).  Here the forward saved one uint8 arg-max slot
// per reduced element, so every backward is a single pass over the OUTPUT-sized tensors:
//   * indices carry no gradient, vertices never require grad (trainer/RL_TDA.py:111), so the
//     unit direction vectors are constants;
//   * arg-max ties only occur at value 0 and carry zero gradient whichever slot was saved.
// Scatter-adds use fp32 atomics (like ATen's index_put_/scatter backward that the reference
// runs), so the last bits depend on arrival order; every cross-cloud reduction (directions,
// column sums) goes through fixed-order partial sums and is deterministic.
#include "common.cuh"
#include <cuda_bf16.h>
#include <float.h>

namespace tgp {

// ------------------------------------------------------------------------------------------
// activation backward of a fused epilogue: gz = g * [y > 0 | slope] * scale
__global__ void act_bwd_kernel(const float* __restrict__ g, long ld_g, const float* __restrict__ y, long ld_y,
                               const float* __restrict__ scale, int relu, long M, int C, float* __restrict__ gz,
                               long ld_gz) {
    const long total = M * C;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / C;
        const int c = (int)(e - r * C);
        float v = __ldg(g + r * ld_g + c);
        if (relu && !(__ldg(y + r * ld_y + c) > 0.f)) v = 0.f;
        if (scale) v *= __ldg(scale + c);
        gz[r * ld_gz + c] = v;
    }
}

// ------------------------------------------------------------------------------------------
// per-group column sums: out[g, c] = sum_{r in group g} x[r, c]   (rows_per_group rows each)
// CTA = (32-column chunk, group, split); partials reduced in a fixed order by the finalize kernel.
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS)
colsum_partial_kernel(const float* __restrict__ x, long ld, long rows_per_group, int C, float* __restrict__ partial) {
    __shared__ float part[CS_THREADS / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const long grp = blockIdx.y;
    const int nsplit = gridDim.z, sp = blockIdx.z;
    const long per = (rows_per_group + nsplit - 1) / nsplit;
    const long r_beg = grp * rows_per_group + sp * per;
    const long r_end = min((grp + 1) * rows_per_group, r_beg + per);
    float acc0 = 0.f, acc1 = 0.f;
    if (c < C) {
        long r = r_beg + warp;
        for (; r + CS_THREADS / 32 < r_end; r += 2 * (CS_THREADS / 32)) {
            acc0 += __ldg(x + r * ld + c);
            acc1 += __ldg(x + (r + CS_THREADS / 32) * ld + c);
        }
        if (r < r_end) acc0 += __ldg(x + r * ld + c);
    }
    part[warp][lane] = acc0 + acc1;
    __syncthreads();
    if (warp == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CS_THREADS / 32; ++w) s += part[w][lane];
        partial[(grp * nsplit + sp) * C + c] = s;
    }
}

__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nsplit, int C, long total,
                                       float* __restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long g = e / C;
    const int c = (int)(e - g * C);
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[(g * nsplit + sp) * C + c];
    out[e] = s;
}

// ------------------------------------------------------------------------------------------
// gather-max backward: df[b, idx[b, rows[m], arg[b,m,c]], c] += scale * g[b, m, c]
// (g_bcast: g is (B,C) and is broadcast over m -- the ORL mean, gcn3d.py:216, scale = 1/N)
template <typename IdxT>
__global__ void gather_max_bwd_kernel(const float* __restrict__ g, int g_bcast, float scale,
                                      const IdxT* __restrict__ idx, const int64_t* __restrict__ rows,
                                      const uint8_t* __restrict__ arg, int N, int M, int k, int C, long total,
                                      float* __restrict__ df) {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / C;          // b*M + m
        const int c = (int)(e - r * C);
        const long b = r / M;
        const int m = (int)(r - b * M);
        const long n = rows ? (long)__ldg(rows + m) : m;
        const float v = scale * (g_bcast ? __ldg(g + b * C + c) : __ldg(g + e));
        const int j = arg[e];
        const int nb = ld_idx(idx, (b * N + n) * k + j);
        atomicAdd(df + (b * N + nb) * (long)C + c, v);
    }
}

// indexing_neighbor_new backward: dt[b, index[r], :] += g[r, :]
template <typename IdxT>
__global__ void scatter_add_rows_kernel(const float* __restrict__ g, const IdxT* __restrict__ index, long rows,
                                        long rows_per_cloud, int N, int C, float* __restrict__ dt) {
    const long total = rows * C;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / C;
        const int c = (int)(e - r * C);
        const long b = r / rows_per_cloud;
        atomicAdd(dt + (b * N + ld_idx(index, r)) * (long)C + c, __ldg(g + e));
    }
}

// ------------------------------------------------------------------------------------------
// column-normalised support direction and its norm (F.normalize(directions, dim=0), gcn3d.py:99,165)
__device__ __forceinline__ void load_sd_bwd(const float* __restrict__ directions, int SC, int col, float& x, float& y,
                                            float& z) {
    x = __ldg(directions + col);
    y = __ldg(directions + SC + col);
    z = __ldg(directions + 2 * SC + col);
    normalize3(x, y, z);
}

// layer conv backward.  CTA = (4-channel group cg, cloud b), lane = s*4 + c4 like the forward.
//   d_support[b, nb*, s, c] += (G[b,n,c]/S) * theta*            -> shared-memory table [N][W], then ONE coalesced
//                                                                   store into the gradient operand dP (row-major)
//   du[:, s*C+c]           += (G[b,n,c]/S) * sup* * [theta*>0] * d*   -> per-CTA partial [3][W]
// with * = the saved arg-max neighbour.  The support value is read from the slab in global memory (one gather
// per output element, not k as in the forward, so no table staging).
template <int DUMMY>
__global__ void __launch_bounds__(1024)
layer_conv_bwd_kernel(const float4* __restrict__ rec, const float* __restrict__ directions,
                      const float* __restrict__ slab, const uint8_t* __restrict__ arg_slab,
                      const float* __restrict__ G, long ld_g, long M, int N, int k, int S, int C,
                      float* __restrict__ d_support, long ld_ds, float* __restrict__ du_part) {
    extern __shared__ __align__(16) float dtab[];            // [N][W]
    __shared__ float red[32][3][32];
    const int W = S * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cg = blockIdx.x;
    const long b = blockIdx.y;
    const int SC = S * C;
    for (int i = threadIdx.x; i < N * W; i += blockDim.x) dtab[i] = 0.f;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    const int s_l = lane >> 2, c4 = lane & 3;
    if (lane < W) load_sd_bwd(directions, SC, s_l * C + cg * 4 + c4, sx, sy, sz);
    __syncthreads();
    const float inv_s = 1.0f / (float)S;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (lane < W) {
        for (int n = warp; n < N; n += nwarps) {
            const long pt = b * N + n;
            const float gsc = __ldg(G + pt * ld_g + cg * 4 + c4) * inv_s;
            const int j = arg_slab[((long)cg * M + pt) * W + lane];
            const float4 d = __ldg(rec + (b * k + j) * N + n);      // neighbour-major records (tgp_edge_records)
            const int nb = __float_as_int(d.w);
            const float sup = __ldg(slab + ((long)cg * M + b * N + nb) * W + lane);
            const float th = fmaxf(fmaf(d.z, sz, fmaf(d.y, sy, d.x * sx)), 0.f);
            atomicAdd(dtab + nb * W + lane, gsc * th);
            const float dth = th > 0.f ? gsc * sup : 0.f;
            ux = fmaf(dth, d.x, ux);
            uy = fmaf(dth, d.y, uy);
            uz = fmaf(dth, d.z, uz);
        }
    }
    red[warp][0][lane] = ux; red[warp][1][lane] = uy; red[warp][2][lane] = uz;
    __syncthreads();
    // gradient table -> dP columns [cg*W, (cg+1)*W) of every row of this cloud
    for (int i = threadIdx.x; i < N * W; i += blockDim.x) {
        const int n = i / W, l = i - n * W;
        d_support[(b * N + n) * ld_ds + cg * W + l] = dtab[i];
    }
    if (warp < 3 && lane < W) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += red[w][warp][lane];
        du_part[((b * gridDim.x + cg) * 3 + warp) * W + lane] = s;
    }
}

// surface conv backward: only the support directions receive a gradient (gcn3d.py:91-106).
// CTA = SB_PTS consecutive points of one cloud; thread = support-direction columns col, col+256, ...
constexpr int SB_THREADS = 256;
constexpr int SB_PTS = 64;
constexpr int SB_MAXCOLS = 4;     // S*C <= 1024
template <typename IdxT>
__global__ void __launch_bounds__(SB_THREADS)
surface_conv_bwd_kernel(const float* __restrict__ xyz, const IdxT* __restrict__ idx,
                        const float* __restrict__ directions, const uint8_t* __restrict__ arg,
                        const float* __restrict__ G, long ld_g, int N, int k, int S, int C,
                        float* __restrict__ du_part) {
    extern __shared__ __align__(16) float4 sdirs[];          // [SB_PTS][k]
    const int SC = S * C;
    const long b = blockIdx.y;
    const int n0 = blockIdx.x * SB_PTS;
    const int npts = min(SB_PTS, N - n0);
    for (int e = threadIdx.x; e < npts * k; e += SB_THREADS) {
        const int pl = e / k;
        const long pt = b * N + n0 + pl;
        const float* p = xyz + (b * N + ld_idx(idx, pt * k + (e - pl * k))) * 3;
        float x = __ldg(p) - __ldg(xyz + pt * 3), y = __ldg(p + 1) - __ldg(xyz + pt * 3 + 1),
              z = __ldg(p + 2) - __ldg(xyz + pt * 3 + 2);
        normalize3(x, y, z);
        sdirs[e] = make_float4(x, y, z, 0.f);
    }
    __syncthreads();
    const float inv_s = 1.0f / (float)S;
    const long chunk = b * gridDim.x + blockIdx.x;
#pragma unroll 1
    for (int q = 0; q < SB_MAXCOLS; ++q) {
        const int col = threadIdx.x + q * SB_THREADS;
        if (col >= SC) break;
        float sx, sy, sz;
        load_sd_bwd(directions, SC, col, sx, sy, sz);
        const int c = col % C;
        float ux = 0.f, uy = 0.f, uz = 0.f;
        for (int pl = 0; pl < npts; ++pl) {
            const long pt = b * N + n0 + pl;
            const int j = arg[pt * SC + col];
            const float4 d = sdirs[pl * k + j];
            const float th = fmaf(d.z, sz, fmaf(d.y, sy, d.x * sx));
            const float dth = th > 0.f ? __ldg(G + pt * ld_g + c) * inv_s : 0.f;
            ux = fmaf(dth, d.x, ux);
            uy = fmaf(dth, d.y, uy);
            uz = fmaf(dth, d.z, uz);
        }
        float* o = du_part + chunk * 3 * SC;
        o[col] = ux; o[SC + col] = uy; o[2 * SC + col] = uz;
    }
}

// d directions from the partial du sums: du = sum_p part[p] (fixed order), then the backward of
// F.normalize(directions, dim=0): dv = (du - u (u.du)) / max(||v||, eps).
// layout 0: part[p][3][SC] (surface);  layout 1: part[(b*CG + cg)][3][W], column s*C+c <-> (cg=c/4, lane=s*4+c%4)
__global__ void directions_bwd_kernel(const float* __restrict__ part, int layout, long nparts, int S, int C,
                                      const float* __restrict__ directions, float* __restrict__ d_dir) {
    const int SC = S * C;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= SC) return;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (layout == 0) {
        for (long p = 0; p < nparts; ++p) {
            const float* q = part + p * 3 * SC;
            ux += q[col]; uy += q[SC + col]; uz += q[2 * SC + col];
        }
    } else {
        const int W = S * 4, CG = C / 4;
        const int s = col / C, c = col - s * C;
        const int cg = c >> 2, l = s * 4 + (c & 3);
        for (long bb = 0; bb < nparts; ++bb) {
            const float* q = part + ((bb * CG + cg) * 3) * W + l;
            ux += q[0]; uy += q[W]; uz += q[2 * W];
        }
    }
    const float vx = __ldg(directions + col), vy = __ldg(directions + SC + col), vz = __ldg(directions + 2 * SC + col);
    const float nrm = fmaxf(sqrtf(fmaf(vz, vz, fmaf(vy, vy, vx * vx))), 1e-12f);
    const float inv = 1.0f / nrm;
    const float x = vx * inv, y = vy * inv, z = vz * inv;
    const float dot = x * ux + y * uy + z * uz;
    d_dir[col] = (ux - x * dot) * inv;
    d_dir[SC + col] = (uy - y * dot) * inv;
    d_dir[2 * SC + col] = (uz - z * dot) * inv;
}

// ------------------------------------------------------------------------------------------
// out (K1,K2) = A^T B, A (M,K1), B (M,K2) row-major: the weight-gradient shape (contraction over rows).
// exact fp32; CTA = 32x32 output tile x one split of M; fixed-order partial reduction.
constexpr int TN_T = 32;
__global__ void __launch_bounds__(256)
gemm_tn_partial_kernel(const float* __restrict__ A, long lda, const float* __restrict__ Bm, long ldb, long M, int K1,
                       int K2, float* __restrict__ partial) {
    __shared__ float As[TN_T][TN_T + 1];
    __shared__ float Bs[TN_T][TN_T + 1];
    const int i0 = blockIdx.x * TN_T, j0 = blockIdx.y * TN_T;
    const int nsplit = gridDim.z, sp = blockIdx.z;
    const long per = ((M + nsplit - 1) / nsplit + TN_T - 1) / TN_T * TN_T;
    const long m_beg = sp * per, m_end = min(M, m_beg + per);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // ty 0..7: output rows ty*4..ty*4+3, column tx
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long m0 = m_beg; m0 < m_end; m0 += TN_T) {
#pragma unroll
        for (int r = ty; r < TN_T; r += 8) {
            const long m = m0 + r;
            As[r][tx] = (m < m_end && i0 + tx < K1) ? __ldg(A + m * lda + i0 + tx) : 0.f;
            Bs[r][tx] = (m < m_end && j0 + tx < K2) ? __ldg(Bm + m * ldb + j0 + tx) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < TN_T; ++r) {
            const float bv = Bs[r][tx];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmaf(As[r][ty * 4 + u], bv, acc[u]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + ty * 4 + u, j = j0 + tx;
        if (i < K1 && j < K2) partial[((long)sp * K1 + i) * K2 + j] = acc[u];
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int nsplit, long rows, int cols,
                                     float* __restrict__ out, long ldo) {
    const long total = rows * cols;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[(long)sp * total + e];
        const long r = e / cols;
        out[r * ldo + (e - r * cols)] = s;
    }
}


// ------------------------------------------------------------------------------------------
// train-mode BatchNorm1d + activation around a 1x1 convolution (the heads' Conv1d-BN-ReLU stacks, PoseR.py:26-33,
// PoseTs.py:31-38, FaceRecon.py:95-117,139-141), channel-last rows: statistics over the M = B*N rows per channel.
//   forward : z = x W^T + b (tgp_gemm) ; mean, var (tgp_colsum, tgp_colsumsq_dev) ; y = act(z * scale + shift)
//   backward: g = dy * act'(y) ; dbeta = sum g ; dgamma = sum g * zhat ; dz = gamma*invstd * (g - dbeta/M - zhat*dgamma/M)
// sum over rows of (x - mu[c])^2 (two-pass variance: no cancellation)
__global__ void __launch_bounds__(CS_THREADS)
colsumsq_partial_kernel(const float* __restrict__ x, long ld, long M, int C, const float* __restrict__ mu,
                        float* __restrict__ partial) {
    __shared__ float part[CS_THREADS / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int nsplit = gridDim.z, sp = blockIdx.z;
    const long per = (M + nsplit - 1) / nsplit;
    const long r_beg = sp * per, r_end = min(M, r_beg + per);
    float acc0 = 0.f, acc1 = 0.f;
    if (c < C) {
        const float m = __ldg(mu + c);
        long r = r_beg + warp;
        for (; r + CS_THREADS / 32 < r_end; r += 2 * (CS_THREADS / 32)) {
            const float a = __ldg(x + r * ld + c) - m, b2 = __ldg(x + (r + CS_THREADS / 32) * ld + c) - m;
            acc0 = fmaf(a, a, acc0);
            acc1 = fmaf(b2, b2, acc1);
        }
        if (r < r_end) { const float a = __ldg(x + r * ld + c) - m; acc0 = fmaf(a, a, acc0); }
    }
    part[warp][lane] = acc0 + acc1;
    __syncthreads();
    if (warp == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CS_THREADS / 32; ++w) s += part[w][lane];
        partial[(long)sp * C + c] = s;
    }
}

// y = act(z * scale[c] + shift[c]), act = leaky with `slope` (0: ReLU, 1: identity); raw and/or split destination
__global__ void affine_act_kernel(const float* __restrict__ z, long ld_z, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float slope, long M, int C, float* __restrict__ out,
                                  long ld_out, float* __restrict__ out_split, int kp, int mixed) {
    const long total = M * C;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / C;
        const int c = (int)(e - r * C);
        float v = fmaf(__ldg(z + r * ld_z + c), __ldg(scale + c), __ldg(shift + c));
        v = v > 0.f ? v : v * slope;
        if (out) out[r * ld_out + c] = v;
        if (out_split && mixed) {
            mixed_store1(reinterpret_cast<uint16_t*>(out_split + r * 2 * kp), kp, c, v);      // tgp_gemm_args.mixed
        } else if (out_split) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float hi = __uint_as_float(hb);
            out_split[r * 2 * kp + c] = hi;
            out_split[r * 2 * kp + kp + c] = v - hi;
        }
    }
}

// 128-bit column reductions (C % 4 == 0, 16-byte aligned rows): a lane owns 4 consecutive columns, a CTA a 128-column
// chunk; 4 independent 512-byte row segments in flight per warp.  MODE 0: sum x; 1: sum (x - mu)^2; 2: the two sums of the
// BatchNorm backward (g, g * zhat).  Same partial layout / fixed-order finalize as the scalar kernels.
template <int MODE>
__global__ void __launch_bounds__(CS_THREADS)
colred_vec_kernel(const float* __restrict__ x, long ld, long rows_per_group, int C, const float* __restrict__ mu,
                  const float* __restrict__ y, long ld_y, const float* __restrict__ z, long ld_z,
                  const float* __restrict__ invstd, float slope, float* __restrict__ partial,
                  const float* __restrict__ scale = nullptr, const float* __restrict__ shift = nullptr) {
    __shared__ __align__(16) float part[MODE == 2 ? 2 : 1][CS_THREADS / 32][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const long grp = blockIdx.y;
    const int nsplit = gridDim.z, sp = blockIdx.z;
    const long per = (rows_per_group + nsplit - 1) / nsplit;
    const long r_beg = grp * rows_per_group + sp * per;
    const long r_end = min((grp + 1) * rows_per_group, r_beg + per);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    constexpr int NW = CS_THREADS / 32;
    if (c < C) {
        float4 m = a0, is = a0;
        if (MODE >= 1) m = __ldg(reinterpret_cast<const float4*>(mu + c));
        if (MODE == 2) is = __ldg(reinterpret_cast<const float4*>(invstd + c));
        float4 sc = a0, sh = a0;
        if (MODE == 2 && scale) { sc = __ldg(reinterpret_cast<const float4*>(scale + c)); sh = __ldg(reinterpret_cast<const float4*>(shift + c)); }
        auto body = [&](long r) {
            float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
            if (MODE == 0) { a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w; }
            else if (MODE == 1) {
                v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
                a0.x = fmaf(v.x, v.x, a0.x); a0.y = fmaf(v.y, v.y, a0.y); a0.z = fmaf(v.z, v.z, a0.z); a0.w = fmaf(v.w, v.w, a0.w);
            } else {
                const float4 zz = __ldg(reinterpret_cast<const float4*>(z + r * ld_z + c));
                // activation mask: from y, or recomputed bit-identically from z (y = act(fma(z, scale, shift))) -- one read less
                float4 yy;
                if (scale) yy = make_float4(fmaf(zz.x, sc.x, sh.x), fmaf(zz.y, sc.y, sh.y), fmaf(zz.z, sc.z, sh.z), fmaf(zz.w, sc.w, sh.w));
                else yy = __ldg(reinterpret_cast<const float4*>(y + r * ld_y + c));
                if (!(yy.x > 0.f)) v.x *= slope;
                if (!(yy.y > 0.f)) v.y *= slope;
                if (!(yy.z > 0.f)) v.z *= slope;
                if (!(yy.w > 0.f)) v.w *= slope;
                a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w;
                a1.x = fmaf(v.x, (zz.x - m.x) * is.x, a1.x); a1.y = fmaf(v.y, (zz.y - m.y) * is.y, a1.y);
                a1.z = fmaf(v.z, (zz.z - m.z) * is.z, a1.z); a1.w = fmaf(v.w, (zz.w - m.w) * is.w, a1.w);
            }
        };
        long r = r_beg + warp;
        for (; r + 3 * NW < r_end; r += 4 * NW) { body(r); body(r + NW); body(r + 2 * NW); body(r + 3 * NW); }
        for (; r < r_end; r += NW) body(r);
    }
    *reinterpret_cast<float4*>(&part[0][warp][lane * 4]) = a0;
    if (MODE == 2) *reinterpret_cast<float4*>(&part[MODE == 2 ? 1 : 0][warp][lane * 4]) = a1;
    __syncthreads();
    for (int i = threadIdx.x; i < (MODE == 2 ? 256 : 128); i += CS_THREADS) {
        const int which = i >> 7, cc = blockIdx.x * 128 + (i & 127);
        if (cc < C) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += part[which][w][i & 127];
            if (MODE == 2) partial[((long)sp * 2 + which) * C + cc] = s;
            else partial[(grp * nsplit + sp) * C + cc] = s;
        }
    }
}

// 128-bit variants of the two element-wise passes above / below (C % 4 == 0, 16-byte aligned rows): CTA-strided rows,
// a thread owns 4 consecutive channels, no 64-bit division per element, packed operand stores.
__global__ void __launch_bounds__(256)
affine_act_vec_kernel(const float* __restrict__ z, long ld_z, const float* __restrict__ scale, const float* __restrict__ shift,
                      float slope, long M, int C, float* __restrict__ out, long ld_out, float* __restrict__ out_split, int kp,
                      int mixed) {
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c)), sh = __ldg(reinterpret_cast<const float4*>(shift + c));
        for (long r = blockIdx.x; r < M; r += gridDim.x) {
            float4 v = __ldg(reinterpret_cast<const float4*>(z + r * ld_z + c));
            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
            v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
            if (out) *reinterpret_cast<float4*>(out + r * ld_out + c) = v;
            if (out_split && mixed) {
                mixed_store4(reinterpret_cast<uint16_t*>(out_split + r * 2 * kp), kp, c, v);
            } else if (out_split) {
                float4 hi;
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.x)); hi.x = __uint_as_float(hb);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.y)); hi.y = __uint_as_float(hb);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.z)); hi.z = __uint_as_float(hb);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.w)); hi.w = __uint_as_float(hb);
                float* q = out_split + r * 2 * kp + c;
                *reinterpret_cast<float4*>(q) = hi;
                *reinterpret_cast<float4*>(q + kp) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(const float* __restrict__ dy, long ld_dy, const float* __restrict__ y, long ld_y,
                        const float* __restrict__ z, long ld_z, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ dbeta,
                        const float* __restrict__ dgamma, float slope, long M, int C, float* __restrict__ dz, long ld_dz,
                        float* __restrict__ dz_mixed, int kp, const float* __restrict__ scale, const float* __restrict__ shift) {
    const float inv_m = 1.0f / (float)M;
    if (dz_mixed)      // zero padding of the operand columns past C (Kp = ceil64(C))
        for (long r = blockIdx.x; r < M; r += gridDim.x)
            for (int c = C + threadIdx.x * 4; c < kp; c += blockDim.x * 4)
                mixed_store4(reinterpret_cast<uint16_t*>(dz_mixed + r * 2 * kp), kp, c, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
        const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c)), is = __ldg(reinterpret_cast<const float4*>(invstd + c));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 db = __ldg(reinterpret_cast<const float4*>(dbeta + c)), dg = __ldg(reinterpret_cast<const float4*>(dgamma + c));
        float4 sc = mu, sh = mu;
        if (scale) { sc = __ldg(reinterpret_cast<const float4*>(scale + c)); sh = __ldg(reinterpret_cast<const float4*>(shift + c)); }
        for (long r = blockIdx.x; r < M; r += gridDim.x) {
            float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * ld_dy + c));
            const float4 zz = __ldg(reinterpret_cast<const float4*>(z + r * ld_z + c));
            float4 yy;
            if (scale) yy = make_float4(fmaf(zz.x, sc.x, sh.x), fmaf(zz.y, sc.y, sh.y), fmaf(zz.z, sc.z, sh.z), fmaf(zz.w, sc.w, sh.w));
            else yy = __ldg(reinterpret_cast<const float4*>(y + r * ld_y + c));
            if (!(yy.x > 0.f)) g.x *= slope;
            if (!(yy.y > 0.f)) g.y *= slope;
            if (!(yy.z > 0.f)) g.z *= slope;
            if (!(yy.w > 0.f)) g.w *= slope;
            float4 o;
            o.x = ga.x * is.x * (g.x - db.x * inv_m - (zz.x - mu.x) * is.x * dg.x * inv_m);
            o.y = ga.y * is.y * (g.y - db.y * inv_m - (zz.y - mu.y) * is.y * dg.y * inv_m);
            o.z = ga.z * is.z * (g.z - db.z * inv_m - (zz.z - mu.z) * is.z * dg.z * inv_m);
            o.w = ga.w * is.w * (g.w - db.w * inv_m - (zz.w - mu.w) * is.w * dg.w * inv_m);
            *reinterpret_cast<float4*>(dz + r * ld_dz + c) = o;
            if (dz_mixed) mixed_store4(reinterpret_cast<uint16_t*>(dz_mixed + r * 2 * kp), kp, c, o);
        }
    }
}

// partial[sp][0][c] = sum g, partial[sp][1][c] = sum g * zhat over this split's rows
__global__ void __launch_bounds__(CS_THREADS)
bn_bwd_reduce_kernel(const float* __restrict__ dy, long ld_dy, const float* __restrict__ y, long ld_y,
                     const float* __restrict__ z, long ld_z, const float* __restrict__ mean,
                     const float* __restrict__ invstd, float slope, long M, int C, float* __restrict__ partial) {
    __shared__ float part[2][CS_THREADS / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int nsplit = gridDim.z, sp = blockIdx.z;
    const long per = (M + nsplit - 1) / nsplit;
    const long r_beg = sp * per, r_end = min(M, r_beg + per);
    float s0 = 0.f, s1 = 0.f;
    if (c < C) {
        const float m = __ldg(mean + c), is = __ldg(invstd + c);
        for (long r = r_beg + warp; r < r_end; r += CS_THREADS / 32) {
            float g = __ldg(dy + r * ld_dy + c);
            if (!(__ldg(y + r * ld_y + c) > 0.f)) g *= slope;
            s0 += g;
            s1 = fmaf(g, (__ldg(z + r * ld_z + c) - m) * is, s1);
        }
    }
    part[0][warp][lane] = s0;
    part[1][warp][lane] = s1;
    __syncthreads();
    if (warp < 2 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CS_THREADS / 32; ++w) s += part[warp][w][lane];
        partial[((long)sp * 2 + warp) * C + c] = s;
    }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nsplit, int C, float* __restrict__ dbeta,
                                       float* __restrict__ dgamma) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b2 = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) { a += partial[((long)sp * 2) * C + c]; b2 += partial[((long)sp * 2 + 1) * C + c]; }
    dbeta[c] = a;
    dgamma[c] = b2;
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, long ld_dy, const float* __restrict__ y, long ld_y,
                                    const float* __restrict__ z, long ld_z, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ dbeta, const float* __restrict__ dgamma, float slope,
                                    long M, int C, float* __restrict__ dz, long ld_dz) {
    const long total = M * C;
    const float inv_m = 1.0f / (float)M;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / C;
        const int c = (int)(e - r * C);
        float g = __ldg(dy + r * ld_dy + c);
        if (!(__ldg(y + r * ld_y + c) > 0.f)) g *= slope;
        const float is = __ldg(invstd + c);
        const float zh = (__ldg(z + r * ld_z + c) - __ldg(mean + c)) * is;
        dz[r * ld_dz + c] = __ldg(gamma + c) * is * (g - __ldg(dbeta + c) * inv_m - zh * __ldg(dgamma + c) * inv_m);
    }
}

}  // namespace tgp

using namespace tgp;

static unsigned grid_for(long total, int threads, int per_sm = 16) {
    long nb = (total + threads - 1) / threads;
    const long cap = (long)TGP_NUM_SMS * per_sm;
    if (nb > cap) nb = cap;
    return (unsigned)(nb < 1 ? 1 : nb);
}

extern "C" int tgp_act_bwd(const float* grad, long ld_grad, const float* y, long ld_y, const float* scale, int relu,
                           long M, int C, float* gz, long ld_gz, tgp_stream_t stream) {
    if (!grad || !gz || (relu && !y)) return fail(TGP_EINVAL, "tgp_act_bwd: null pointer");
    if (M <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_act_bwd: sizes must be positive");
    act_bwd_kernel<<<grid_for(M * C, 256), 256, 0, as_stream(stream)>>>(grad, ld_grad, y, ld_y, scale, relu, M, C, gz, ld_gz);
    return check_launch("act_bwd_kernel");
}

static int colsum_nsplit(long rows_per_group, long groups, int C) {
    // enough CTAs to fill the machine, at least 64 rows each; depends on the shape only (deterministic)
    const long chunks = C % 4 == 0 ? (C + 127) / 128 : (C + 31) / 32;     // CTAs per split (128-bit kernels own 128 columns)
    long want = ((long)TGP_NUM_SMS * 4 + chunks * groups - 1) / (chunks * groups);
    long cap = (rows_per_group + 63) / 64;
    long s = want < cap ? want : cap;
    return (int)(s < 1 ? 1 : (s > 256 ? 256 : s));
}

extern "C" size_t tgp_colsum_workspace(long M, int C, long rows_per_group) {
    if (rows_per_group <= 0 || M <= 0 || C <= 0) return 0;
    const long groups = M / rows_per_group;
    return (size_t)groups * colsum_nsplit(rows_per_group, groups, C) * C * sizeof(float);
}

extern "C" int tgp_colsum(const float* x, long ld, long M, int C, long rows_per_group, float* out, void* workspace,
                          size_t workspace_bytes, tgp_stream_t stream) {
    if (!x || !out || !workspace) return fail(TGP_EINVAL, "tgp_colsum: null pointer");
    if (M <= 0 || C <= 0 || rows_per_group <= 0 || M % rows_per_group) return fail(TGP_EINVAL, "tgp_colsum: M must be a positive multiple of rows_per_group");
    if (workspace_bytes < tgp_colsum_workspace(M, C, rows_per_group)) return fail(TGP_ENOSPACE, "tgp_colsum: workspace too small");
    const long groups = M / rows_per_group;
    if (groups > 65535) return fail(TGP_EINVAL, "tgp_colsum: more than 65535 groups");
    const int nsplit = colsum_nsplit(rows_per_group, groups, C);
    cudaStream_t st = as_stream(stream);
    dim3 grid((C + 31) / 32, (unsigned)groups, nsplit);
    int rc;
    if (C % 4 == 0 && ld % 4 == 0 && ((uintptr_t)x & 15) == 0) {
        grid.x = (C + 127) / 128;
        colred_vec_kernel<0><<<grid, CS_THREADS, 0, st>>>(x, ld, rows_per_group, C, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f,
                                                         static_cast<float*>(workspace));
        rc = check_launch("colred_vec_kernel<0>");
    } else {
        colsum_partial_kernel<<<grid, CS_THREADS, 0, st>>>(x, ld, rows_per_group, C, static_cast<float*>(workspace));
        rc = check_launch("colsum_partial_kernel");
    }
    if (rc) return rc;
    const long total = groups * C;
    colsum_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(static_cast<const float*>(workspace), nsplit, C, total, out);
    return check_launch("colsum_finalize_kernel");
}

extern "C" int tgp_gather_max_bwd(const float* grad, int grad_is_per_cloud, float scale, const void* idx, int idx_bits,
                                  const int64_t* rows, const uint8_t* arg, int B, int N, int M, int k, int C,
                                  float* dfeat, tgp_stream_t stream) {
    if (!grad || !idx || !arg || !dfeat) return fail(TGP_EINVAL, "tgp_gather_max_bwd: null pointer");
    if (B <= 0 || N <= 0 || M <= 0 || k <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_gather_max_bwd: sizes must be positive");
    if (!rows && M != N) return fail(TGP_EINVAL, "tgp_gather_max_bwd: rows == NULL requires M == N");
    const long total = (long)B * M * C;
    cudaStream_t st = as_stream(stream);
    TGP_DISPATCH_IDX(idx_bits, {
        gather_max_bwd_kernel<IdxT><<<grid_for(total, 256, 32), 256, 0, st>>>(grad, grad_is_per_cloud, scale, (const IdxT*)idx, rows,
                                                                             arg, N, M, k, C, total, dfeat);
    });
    return check_launch("gather_max_bwd_kernel");
}

extern "C" int tgp_scatter_add_rows(const float* grad, const void* index, int idx_bits, int B, int N, int M, int k,
                                    int C, float* dtensor, tgp_stream_t stream) {
    if (!grad || !index || !dtensor) return fail(TGP_EINVAL, "tgp_scatter_add_rows: null pointer");
    if (B <= 0 || N <= 0 || M <= 0 || k <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_scatter_add_rows: sizes must be positive");
    const long rows = (long)B * M * k;
    cudaStream_t st = as_stream(stream);
    TGP_DISPATCH_IDX(idx_bits, {
        scatter_add_rows_kernel<IdxT><<<grid_for(rows * C, 256, 32), 256, 0, st>>>(grad, (const IdxT*)index, rows, (long)M * k, N, C, dtensor);
    });
    return check_launch("scatter_add_rows_kernel");
}

extern "C" size_t tgp_layer_conv_bwd_workspace(int B, int S, int C) { return (size_t)B * (C / 4) * 3 * S * 4 * sizeof(float); }

extern "C" int tgp_layer_conv_bwd(const float* edge_rec, const float* directions, const float* support_slab,
                                  const uint8_t* arg_slab, const float* grad, long ld_grad, int B, int N, int k, int S,
                                  int C, float* d_support, long ld_ds, float* d_directions, void* workspace,
                                  size_t workspace_bytes, tgp_stream_t stream) {
    if (!edge_rec || !directions || !support_slab || !arg_slab || !grad || !d_support || !d_directions || !workspace)
        return fail(TGP_EINVAL, "tgp_layer_conv_bwd: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || S <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_layer_conv_bwd: sizes must be positive");
    if (C % 4 || S * 4 > 32) return fail(TGP_EINVAL, "tgp_layer_conv_bwd: C % 4 != 0 or S > 8");
    if (B > 65535) return fail(TGP_EINVAL, "tgp_layer_conv_bwd: B > 65535");
    if (workspace_bytes < tgp_layer_conv_bwd_workspace(B, S, C)) return fail(TGP_ENOSPACE, "tgp_layer_conv_bwd: workspace too small");
    const int W = S * 4;
    const size_t smem = (size_t)N * W * sizeof(float);
    if (smem + 13 * 1024 > 227 * 1024) return fail(TGP_EINVAL, "tgp_layer_conv_bwd: N*S too large for the shared-memory gradient table");
    const int threads = N >= 512 ? 1024 : (N >= 128 ? 256 : 128);
    cudaStream_t st = as_stream(stream);
    static std::atomic<unsigned long long> attr{0};   // one bit per device: function attributes are per device
    if (first_on_device(attr)) {
        cudaFuncSetAttribute(layer_conv_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 13 * 1024);
    }
    dim3 grid(C / 4, B);
    float* part = static_cast<float*>(workspace);
    layer_conv_bwd_kernel<0><<<grid, threads, smem, st>>>(reinterpret_cast<const float4*>(edge_rec), directions, support_slab,
                                                        arg_slab, grad, ld_grad, (long)B * N, N, k, S, C, d_support, ld_ds, part);
    int rc = check_launch("layer_conv_bwd_kernel");
    if (rc) return rc;
    const int SC = S * C;
    directions_bwd_kernel<<<(SC + 127) / 128, 128, 0, st>>>(part, 1, B, S, C, directions, d_directions);
    return check_launch("directions_bwd_kernel");
}

extern "C" size_t tgp_surface_conv_bwd_workspace(int B, int N, int S, int C) {
    return (size_t)B * ((N + SB_PTS - 1) / SB_PTS) * 3 * S * C * sizeof(float);
}

extern "C" int tgp_surface_conv_bwd(const float* xyz, const void* idx, int idx_bits, const float* directions,
                                    const uint8_t* arg, const float* grad, long ld_grad, int B, int N, int k, int S,
                                    int C, float* d_directions, void* workspace, size_t workspace_bytes,
                                    tgp_stream_t stream) {
    if (!xyz || !idx || !directions || !arg || !grad || !d_directions || !workspace)
        return fail(TGP_EINVAL, "tgp_surface_conv_bwd: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || S <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_surface_conv_bwd: sizes must be positive");
    if (S * C > SB_THREADS * SB_MAXCOLS) return fail(TGP_EINVAL, "tgp_surface_conv_bwd: S*C > 1024 unsupported");
    if (B > 65535) return fail(TGP_EINVAL, "tgp_surface_conv_bwd: B > 65535");
    if (workspace_bytes < tgp_surface_conv_bwd_workspace(B, N, S, C)) return fail(TGP_ENOSPACE, "tgp_surface_conv_bwd: workspace too small");
    const size_t smem = sizeof(float4) * SB_PTS * k;
    if (smem > 200 * 1024) return fail(TGP_EINVAL, "tgp_surface_conv_bwd: k too large");
    cudaStream_t st = as_stream(stream);
    const int chunks = (N + SB_PTS - 1) / SB_PTS;
    dim3 grid(chunks, B);
    float* part = static_cast<float*>(workspace);
    TGP_DISPATCH_IDX(idx_bits, {
        if (smem > 48 * 1024) cudaFuncSetAttribute(surface_conv_bwd_kernel<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        surface_conv_bwd_kernel<IdxT><<<grid, SB_THREADS, smem, st>>>(xyz, (const IdxT*)idx, directions, arg, grad, ld_grad, N, k, S, C, part);
    });
    int rc = check_launch("surface_conv_bwd_kernel");
    if (rc) return rc;
    const int SC = S * C;
    directions_bwd_kernel<<<(SC + 127) / 128, 128, 0, st>>>(part, 0, (long)B * chunks, S, C, directions, d_directions);
    return check_launch("directions_bwd_kernel");
}

static int tn_nsplit(long M, int K1, int K2) {
    const long tiles = (long)((K1 + TN_T - 1) / TN_T) * ((K2 + TN_T - 1) / TN_T);
    long want = ((long)TGP_NUM_SMS * 4 + tiles - 1) / tiles;
    long cap = (M + 127) / 128;
    long s = want < cap ? want : cap;
    return (int)(s < 1 ? 1 : (s > 1024 ? 1024 : s));
}

extern "C" size_t tgp_gemm_tn_workspace(long M, int K1, int K2) {
    if (M <= 0 || K1 <= 0 || K2 <= 0) return 0;
    return (size_t)tn_nsplit(M, K1, K2) * K1 * K2 * sizeof(float);
}

extern "C" int tgp_gemm_tn(const float* A, long lda, const float* Bm, long ldb, long M, int K1, int K2, float* out,
                           long ldo, void* workspace, size_t workspace_bytes, tgp_stream_t stream) {
    if (!A || !Bm || !out || !workspace) return fail(TGP_EINVAL, "tgp_gemm_tn: null pointer");
    if (M <= 0 || K1 <= 0 || K2 <= 0) return fail(TGP_EINVAL, "tgp_gemm_tn: sizes must be positive");
    if (workspace_bytes < tgp_gemm_tn_workspace(M, K1, K2)) return fail(TGP_ENOSPACE, "tgp_gemm_tn: workspace too small");
    const int nsplit = tn_nsplit(M, K1, K2);
    cudaStream_t st = as_stream(stream);
    dim3 grid((K1 + TN_T - 1) / TN_T, (K2 + TN_T - 1) / TN_T, nsplit);
    float* part = static_cast<float*>(workspace);
    gemm_tn_partial_kernel<<<grid, 256, 0, st>>>(A, lda, Bm, ldb, M, K1, K2, part);
    int rc = check_launch("gemm_tn_partial_kernel");
    if (rc) return rc;
    splitk_reduce_kernel<<<grid_for((long)K1 * K2, 256), 256, 0, st>>>(part, nsplit, K1, K2, out, ldo);
    return check_launch("splitk_reduce_kernel");
}

static int rowsplit(long M, int C, int per_row_min) {
    const long chunks = C % 4 == 0 ? (C + 127) / 128 : (C + 31) / 32;
    long want = ((long)TGP_NUM_SMS * 4 + chunks - 1) / chunks;
    long cap = (M + per_row_min - 1) / per_row_min;
    long s = want < cap ? want : cap;
    return (int)(s < 1 ? 1 : (s > 256 ? 256 : s));
}

extern "C" size_t tgp_bn_workspace(long M, int C) { return (size_t)rowsplit(M, C, 64) * 2 * C * sizeof(float); }

extern "C" int tgp_colsumsq_dev(const float* x, long ld, long M, int C, const float* mu, float* out, void* workspace,
                                size_t workspace_bytes, tgp_stream_t stream) {
    if (!x || !mu || !out || !workspace) return fail(TGP_EINVAL, "tgp_colsumsq_dev: null pointer");
    if (M <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_colsumsq_dev: sizes must be positive");
    if (workspace_bytes < tgp_bn_workspace(M, C)) return fail(TGP_ENOSPACE, "tgp_colsumsq_dev: workspace too small");
    const int nsplit = rowsplit(M, C, 64);
    cudaStream_t st = as_stream(stream);
    dim3 grid((C + 31) / 32, 1, nsplit);
    int rc;
    if (C % 4 == 0 && ld % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)mu & 15) == 0) {
        grid.x = (C + 127) / 128;
        colred_vec_kernel<1><<<grid, CS_THREADS, 0, st>>>(x, ld, M, C, mu, nullptr, 0, nullptr, 0, nullptr, 0.f, static_cast<float*>(workspace));
        rc = check_launch("colred_vec_kernel<1>");
    } else {
        colsumsq_partial_kernel<<<grid, CS_THREADS, 0, st>>>(x, ld, M, C, mu, static_cast<float*>(workspace));
        rc = check_launch("colsumsq_partial_kernel");
    }
    if (rc) return rc;
    colsum_finalize_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(static_cast<const float*>(workspace), nsplit, C, C, out);
    return check_launch("colsum_finalize_kernel");
}

extern "C" int tgp_affine_act(const float* z, long ld_z, const float* scale, const float* shift, float slope, long M,
                              int C, float* out, long ld_out, float* out_split, int Kp, int mixed, tgp_stream_t stream) {
    if (!z || !scale || !shift || (!out && !out_split)) return fail(TGP_EINVAL, "tgp_affine_act: null pointer");
    if (M <= 0 || C <= 0 || (out_split && Kp < C) || (out_split && mixed && Kp % 64)) return fail(TGP_EINVAL, "tgp_affine_act: bad sizes");
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (C % 4 == 0 && ld_z % 4 == 0 && al16(z) && al16(scale) && al16(shift) && (!out || (ld_out % 4 == 0 && al16(out))) &&
        (!out_split || (al16(out_split) && Kp % 4 == 0))) {
        const long nb = M < (long)TGP_NUM_SMS * 16 ? M : (long)TGP_NUM_SMS * 16;
        affine_act_vec_kernel<<<(unsigned)nb, C >= 1024 ? 256 : (C >= 256 ? 64 : 32), 0, as_stream(stream)>>>(
            z, ld_z, scale, shift, slope, M, C, out, ld_out, out_split, Kp, mixed);
        return check_launch("affine_act_vec_kernel");
    }
    affine_act_kernel<<<grid_for(M * C, 256), 256, 0, as_stream(stream)>>>(z, ld_z, scale, shift, slope, M, C, out, ld_out, out_split, Kp, mixed);
    return check_launch("affine_act_kernel");
}

extern "C" int tgp_bn_bwd(const float* dy, long ld_dy, const float* y, long ld_y, const float* z, long ld_z,
                          const float* mean, const float* invstd, const float* gamma, const float* scale,
                          const float* shift, float slope, long M, int C,
                          float* dz, long ld_dz, float* dz_mixed, float* dbeta, float* dgamma, void* workspace,
                          size_t workspace_bytes, tgp_stream_t stream) {
    if (!dy || !y || !z || !mean || !invstd || !gamma || !dz || !dbeta || !dgamma || !workspace)
        return fail(TGP_EINVAL, "tgp_bn_bwd: null pointer");
    if (M <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_bn_bwd: sizes must be positive");
    if (workspace_bytes < tgp_bn_workspace(M, C)) return fail(TGP_ENOSPACE, "tgp_bn_bwd: workspace too small");
    const int nsplit = rowsplit(M, C, 64);
    cudaStream_t st = as_stream(stream);
    dim3 grid((C + 31) / 32, 1, nsplit);
    float* part = static_cast<float*>(workspace);
    int rc;
    auto a16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (C % 4 == 0 && ld_dy % 4 == 0 && ld_y % 4 == 0 && ld_z % 4 == 0 && a16(dy) && a16(y) && a16(z) && a16(mean) && a16(invstd)) {
        grid.x = (C + 127) / 128;
        const bool mask_from_z = scale && shift && a16(scale) && a16(shift);
        colred_vec_kernel<2><<<grid, CS_THREADS, 0, st>>>(dy, ld_dy, M, C, mean, y, ld_y, z, ld_z, invstd, slope, part,
                                                         mask_from_z ? scale : nullptr, mask_from_z ? shift : nullptr);
        rc = check_launch("colred_vec_kernel<2>");
    } else {
        bn_bwd_reduce_kernel<<<grid, CS_THREADS, 0, st>>>(dy, ld_dy, y, ld_y, z, ld_z, mean, invstd, slope, M, C, part);
        rc = check_launch("bn_bwd_reduce_kernel");
    }
    if (rc) return rc;
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(part, nsplit, C, dbeta, dgamma);
    rc = check_launch("bn_bwd_finalize_kernel");
    if (rc) return rc;
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (C % 4 == 0 && ld_dy % 4 == 0 && ld_y % 4 == 0 && ld_z % 4 == 0 && ld_dz % 4 == 0 && al16(dy) && al16(y) && al16(z) &&
        al16(dz) && al16(mean) && al16(invstd) && al16(gamma) && al16(dbeta) && al16(dgamma)) {
        const long nb = M < (long)TGP_NUM_SMS * 16 ? M : (long)TGP_NUM_SMS * 16;
        if (dz_mixed && !al16(dz_mixed)) return fail(TGP_EINVAL, "tgp_bn_bwd: dz_mixed must be 16-byte aligned");
        bn_bwd_apply_vec_kernel<<<(unsigned)nb, C >= 1024 ? 256 : (C >= 256 ? 64 : 32), 0, st>>>(
            dy, ld_dy, y, ld_y, z, ld_z, mean, invstd, gamma, dbeta, dgamma, slope, M, C, dz, ld_dz, dz_mixed,
            tgp_mixed_kpad(C), (scale && shift && al16(scale) && al16(shift)) ? scale : nullptr, shift);
        return check_launch("bn_bwd_apply_vec_kernel");
    }
    if (dz_mixed) return fail(TGP_EINVAL, "tgp_bn_bwd: dz_mixed needs C % 4 == 0 and 16-byte aligned operands");
    bn_bwd_apply_kernel<<<grid_for(M * C, 256), 256, 0, st>>>(dy, ld_dy, y, ld_y, z, ld_z, mean, invstd, gamma, dbeta, dgamma, slope, M, C, dz, ld_dz);
    return check_launch("bn_bwd_apply_kernel");
}
