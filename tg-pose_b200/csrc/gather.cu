// gather.cu -- index-driven data movement of the 3D-GCN path (HBM/L2-bound kernels).
//
// Replaces indexing_neighbor_new (gcn3d.py:38-46), get_neighbor_direction_norm (:48-58),
// the gather+max of get_ORL_global (:210-217) and Pool_layer (:225-245).  The reference
// writes the k-fold expanded (B,M,k,C) tensor to memory and reduces it in a second pass;
// the *_max kernels below reduce while gathering, so only (B,M,C) is ever written.
#include "common.cuh"
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

namespace tgp {

// out[r, :] = tensor[b*N + index[r], :]  for r over B*M*k rows; VEC = floats per thread access
template <typename IdxT, int VEC>
__global__ void gather_rows_kernel(const float* __restrict__ t, const IdxT* __restrict__ index, long rows,
                                   long rows_per_cloud, int N, int C, float* __restrict__ out) {
    const int cv = C / VEC;
    const long total = rows * cv;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / cv;
        const int c = (int)(e - r * cv) * VEC;
        const long b = r / rows_per_cloud;
        const float* src = t + (b * N + ld_idx(index, r)) * (long)C + c;
        float* dst = out + r * (long)C + c;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
        else *dst = __ldg(src);
    }
}

// out[b,m,:] = t[b, rows[m], :]
__global__ void select_rows_kernel(const float* __restrict__ t, const int64_t* __restrict__ rows, long total, int N,
                                   int M, int C, float* __restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long r = e / C;
    const int c = (int)(e - r * C);
    const long b = r / M;
    out[e] = __ldg(t + (b * N + __ldg(rows + (r - b * M))) * C + c);
}

// (B,N,k) -> (B,N,k,3) unit direction from the centre point to each neighbour
template <typename IdxT>
__global__ void direction_norm_kernel(const float* __restrict__ xyz, const IdxT* __restrict__ idx, long total,
                                      int N, int k, float* __restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long pt = e / k;            // b*N + n
    const long b = pt / N;
    const long nb = b * N + ld_idx(idx, e);
    float x = __ldg(xyz + nb * 3) - __ldg(xyz + pt * 3);
    float y = __ldg(xyz + nb * 3 + 1) - __ldg(xyz + pt * 3 + 1);
    float z = __ldg(xyz + nb * 3 + 2) - __ldg(xyz + pt * 3 + 2);
    normalize3(x, y, z);
    out[e * 3] = x; out[e * 3 + 1] = y; out[e * 3 + 2] = z;
}

// out[b,m,c] = max_j f[b, idx[b, rows[m], j], c]; one thread per VEC channels of one output row
template <typename IdxT, int VEC, bool ARG>
__global__ void gather_max_kernel(const float* __restrict__ f, const IdxT* __restrict__ idx,
                                  const int64_t* __restrict__ rows, int N, int M, int k, int C, long total,
                                  float* __restrict__ out, uint8_t* __restrict__ arg) {
    const int cv = C / VEC;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long r = e / cv;            // b*M + m
    const int c = (int)(e - r * cv) * VEC;
    const long b = r / M;
    const int m = (int)(r - b * M);
    const long n = rows ? (long)__ldg(rows + m) : m;
    const IdxT* id = idx + (b * N + n) * k;
    float best[VEC];
    int bj[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { best[v] = -FLT_MAX; bj[v] = 0; }
    for (int j = 0; j < k; ++j) {
        const float* src = f + (b * N + ld_idx(id, j)) * (long)C + c;
        float val[VEC];
        if (VEC == 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(src));
            val[0] = t4.x; val[1 % VEC] = t4.y; val[2 % VEC] = t4.z; val[3 % VEC] = t4.w;
        } else val[0] = __ldg(src);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (val[v] > best[v]) { best[v] = val[v]; bj[v] = j; }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        out[r * (long)C + c + v] = best[v];
        if (ARG) arg[r * (long)C + c + v] = (uint8_t)bj[v];
    }
}

// ORL global feature: part[b,split,c] = sum over this CTA's points of max_j f[b, idx[b,n,j], c];
// orl_finalize_kernel adds the splits in a fixed order and divides by N (deterministic).
// CTA = (32-channel chunk, cloud, point split); lane = channel, warps stride over points.
constexpr int ORL_THREADS = 256;
template <typename IdxT, bool ARG>
__global__ void __launch_bounds__(ORL_THREADS)
orl_global_kernel(const float* __restrict__ f, const IdxT* __restrict__ idx, int N, int k, int C,
                  float* __restrict__ partial, uint8_t* __restrict__ arg) {
    __shared__ float part[ORL_THREADS / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const long b = blockIdx.y;
    const int nsplit = gridDim.z, sp = blockIdx.z;
    const int per = (N + nsplit - 1) / nsplit;
    const int n_beg = sp * per, n_end = min(N, n_beg + per);
    float acc = 0.f;
    for (int n = n_beg + warp; n < n_end; n += ORL_THREADS / 32) {
        const IdxT* id = idx + (b * N + n) * k;
        float best = -FLT_MAX;
        int bj = 0;
        if (c < C) {
#pragma unroll 4
            for (int j = 0; j < k; ++j) {
                const float v = __ldg(f + (b * N + ld_idx(id, j)) * (long)C + c);
                if (v > best) { best = v; bj = j; }
            }
            acc += best;
            if (ARG) arg[(b * N + n) * (long)C + c] = (uint8_t)bj;
        }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < ORL_THREADS / 32; ++w) s += part[w][lane];
        if (c < C) partial[(b * nsplit + sp) * C + c] = s;
    }
}


// ORL global feature with the cloud's 32-channel slice staged in shared memory: CTA = (32-channel chunk, cloud).
// The k-fold gather (20 x the feature map through L2 in the kernel above) becomes shared-memory reads; HBM/L2 traffic
// drops to the compulsory one read of the feature map.
//   stage  : the whole slice (N x 128 B) goes out as 16-byte cp.async pieces at once (no per-row load latency chain);
//   gather : 8 lanes x 4 channels per point, 4 points per warp step; the point's neighbour indices sit 8 per register
//            across its 8 lanes (the next 8 are prefetched while these are consumed) and reach the lanes by width-8
//            shuffles; one LDS.128 per neighbour and lane: 1.75 instructions per (neighbour, channel) instead of 4.
// The per-point maxima are summed per lane, then over the 4 point groups and the warps in a fixed order: deterministic,
// and independent of the batch.
constexpr int ORLS_THREADS = 1024;                          // upper bound; the launch picks 512 or 1024 (tgp_orl_global)
template <typename IdxT, bool ARG>
__global__ void __launch_bounds__(ORLS_THREADS)
orl_smem_kernel(const float* __restrict__ f, const IdxT* __restrict__ idx, int N, int k, int C,
                float* __restrict__ g, uint8_t* __restrict__ arg) {
    extern __shared__ __align__(16) float orl_tab[];       // [N][32]
    __shared__ __align__(16) float part[ORLS_THREADS / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32;
    const long b = blockIdx.y;
    const float* fb = f + b * N * (long)C;
    if (C % 4 == 0 && c0 + 32 <= C && ((uintptr_t)f & 15) == 0) {
        for (int p = threadIdx.x; p < N * 8; p += blockDim.x) {
            const int n = p >> 3, q = p & 7;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(orl_tab + n * 32 + q * 4)),
                         "l"(fb + (long)n * C + c0 + q * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    } else {
        for (int n = warp; n < N; n += blockDim.x / 32) orl_tab[n * 32 + lane] = c0 + lane < C ? __ldg(fb + (long)n * C + c0 + lane) : 0.f;
    }
    __syncthreads();
    const int grp = lane >> 3, l8 = lane & 7;
    const float4* tab4 = reinterpret_cast<const float4*>(orl_tab) + l8;     // row r -> tab4[r * 8]
    const int PSTEP = (blockDim.x / 32) * 4;                                  // points per CTA step
    auto load = [&](int n, int j0) -> int { return (n < N && j0 + l8 < k) ? ld_idx(idx + (b * N + n) * k, j0 + l8) : 0; };
    int n = warp * 4 + grp, j0 = 0;
    int cur = load(n, 0);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    while (n - grp < N) {                                                     // warp-uniform: the 4 groups advance together
        int nj0 = j0 + 8, nn = n;
        if (nj0 >= k) { nj0 = 0; nn = n + PSTEP; }
        const int nxt = load(nn, nj0);
        const int cnt = min(8, k - j0);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            if (jj < cnt) {
                const int nb = __shfl_sync(0xffffffffu, cur, jj, 8);
                const float4 v = tab4[nb * 8];
                if (ARG) {
                    if (v.x > best.x) { best.x = v.x; b0 = j0 + jj; }
                    if (v.y > best.y) { best.y = v.y; b1 = j0 + jj; }
                    if (v.z > best.z) { best.z = v.z; b2 = j0 + jj; }
                    if (v.w > best.w) { best.w = v.w; b3 = j0 + jj; }
                } else {
                    best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y);
                    best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
                }
            }
        }
        if (nj0 == 0) {                                                       // point finished
            if (n < N) {
                acc.x += best.x; acc.y += best.y; acc.z += best.z; acc.w += best.w;
                if (ARG) {
                    uint8_t* ap = arg + (b * N + n) * (long)C + c0 + l8 * 4;
                    if (c0 + l8 * 4 + 3 < C && (C & 3) == 0) *reinterpret_cast<uchar4*>(ap) = make_uchar4((uint8_t)b0, (uint8_t)b1, (uint8_t)b2, (uint8_t)b3);
                    else {
                        if (c0 + l8 * 4 + 0 < C) ap[0] = (uint8_t)b0;
                        if (c0 + l8 * 4 + 1 < C) ap[1] = (uint8_t)b1;
                        if (c0 + l8 * 4 + 2 < C) ap[2] = (uint8_t)b2;
                        if (c0 + l8 * 4 + 3 < C) ap[3] = (uint8_t)b3;
                    }
                }
            }
            best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            b0 = b1 = b2 = b3 = 0;
        }
        cur = nxt; j0 = nj0; n = nn;
    }
    // fixed-order sums: the 4 point groups of a warp, then the warps
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (lane < 8) *reinterpret_cast<float4*>(&part[warp][lane * 4]) = acc;
    __syncthreads();
    if (warp == 0 && c0 + lane < C) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x / 32); ++w) s += part[w][lane];
        g[b * C + c0 + lane] = s / (float)N;
    }
}

__global__ void orl_finalize_kernel(const float* __restrict__ partial, int nsplit, int C, long total, int N,
                                    float* __restrict__ g) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long b = e / C;
    const int c = (int)(e - b * C);
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[(b * nsplit + sp) * C + c];
    g[e] = s / (float)N;
}


// ------------------------------------------------------------------------------------------
// gather-concatenate: out[r, off_s + c] = src_s[row_s(r), c] for up to 8 sources side by side.
// Replaces the chain indexing_neighbor_new(...).squeeze(2) x3 + one_hot expand + torch.cat (FaceRecon.py:69-81)
// and the second cat with the centred points (PoseNet9D.py:63): one pass that reads every source row once and
// writes the concatenated row as raw fp32 and/or directly as the [tf32 | residual] operand of the head GEMMs.
struct ConcatDev {
    tgp_concat_src s[8];
    int nsrc;
};

__global__ void __launch_bounds__(256)
concat_rows_kernel(const __grid_constant__ ConcatDev P, int N, float* __restrict__ out_raw, long ld_raw,
                   float* __restrict__ out_split, int Kp, int mixed, long rows) {
    // a WARP per output row (8 rows per CTA): the per-(row, source) set-up -- row gather, pointer and alignment
    // arithmetic -- is paid by one warp instead of eight; with 8 sources per row that set-up, not the copy, was the cost
    const int lane = threadIdx.x & 31;
    const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const long b = r / N;
    int off = 0;
    for (int si = 0; si < P.nsrc; ++si) {
        const tgp_concat_src& s = P.s[si];
        long row;
        if (s.n_src == 0) row = b;                                         // one row per cloud, broadcast
        else if (s.idx) row = b * s.n_src + __ldg(s.idx + r);              // nearest-upsampling gather
        else row = r;
        const float* src = s.ptr + row * s.ld;
        // 128-bit path: source block and its destination offset are multiples of 4 floats and 16-byte aligned
        const bool v4 = (s.C % 4 == 0) && (off % 4 == 0) && (s.ld % 4 == 0) && (((uintptr_t)s.ptr & 15) == 0) &&
                        (!out_raw || (ld_raw % 4 == 0 && ((uintptr_t)out_raw & 15) == 0));
        if (v4) {
            for (int c = lane * 4; c < s.C; c += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
                if (out_raw) *reinterpret_cast<float4*>(out_raw + r * ld_raw + off + c) = v;
                if (out_split && mixed) {
                    mixed_store4(reinterpret_cast<uint16_t*>(out_split + r * 2 * Kp), Kp, off + c, v);
                } else if (out_split) {
                    float4 hi;
                    uint32_t hb;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.x)); hi.x = __uint_as_float(hb);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.y)); hi.y = __uint_as_float(hb);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.z)); hi.z = __uint_as_float(hb);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v.w)); hi.w = __uint_as_float(hb);
                    float* q = out_split + r * 2 * Kp + off + c;
                    *reinterpret_cast<float4*>(q) = hi;
                    *reinterpret_cast<float4*>(q + Kp) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                }
            }
            off += s.C;
            continue;
        }
        for (int c = lane; c < s.C; c += 32) {
            const float v = __ldg(src + c);
            if (out_raw) out_raw[r * ld_raw + off + c] = v;
            if (out_split && mixed) {
                mixed_store1(reinterpret_cast<uint16_t*>(out_split + r * 2 * Kp), Kp, off + c, v);
            } else if (out_split) {
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
                const float hi = __uint_as_float(hb);
                out_split[r * 2 * Kp + off + c] = hi;
                out_split[r * 2 * Kp + Kp + off + c] = v - hi;
            }
        }
        off += s.C;
    }
    if (out_split)
        for (int c = off + lane; c < Kp; c += 32) {
            if (mixed) mixed_store1(reinterpret_cast<uint16_t*>(out_split + r * 2 * Kp), Kp, c, 0.f);
            else {
                out_split[r * 2 * Kp + c] = 0.f;
                out_split[r * 2 * Kp + Kp + c] = 0.f;
            }
        }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_gather_rows(const float* tensor, const void* index, int idx_bits, int B, int N, int M, int k,
                               int C, float* out, tgp_stream_t stream) {
    if (!tensor || !index || !out) return fail(TGP_EINVAL, "tgp_gather_rows: null pointer");
    if (B <= 0 || N <= 0 || M <= 0 || k <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_gather_rows: sizes must be positive");
    const long rows = (long)B * M * k;
    const bool vec = (C % 4 == 0) && ((uintptr_t)tensor % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const long total = rows * (vec ? C / 4 : C);
    const int threads = 256;
    long nb = (total + threads - 1) / threads;
    if (nb > (long)TGP_NUM_SMS * 32) nb = (long)TGP_NUM_SMS * 32;
    const unsigned blocks = (unsigned)nb;
    cudaStream_t st = as_stream(stream);
    TGP_DISPATCH_IDX(idx_bits, {
        if (vec) gather_rows_kernel<IdxT, 4><<<blocks, threads, 0, st>>>(tensor, (const IdxT*)index, rows, (long)M * k, N, C, out);
        else gather_rows_kernel<IdxT, 1><<<blocks, threads, 0, st>>>(tensor, (const IdxT*)index, rows, (long)M * k, N, C, out);
    });
    return check_launch("gather_rows_kernel");
}

extern "C" int tgp_select_rows(const float* t, const int64_t* rows, int B, int N, int M, int C, float* out,
                               tgp_stream_t stream) {
    if (!t || !rows || !out) return fail(TGP_EINVAL, "tgp_select_rows: null pointer");
    if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_select_rows: sizes must be positive");
    const long total = (long)B * M * C;
    select_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(t, rows, total, N, M, C, out);
    return check_launch("select_rows_kernel");
}

extern "C" int tgp_direction_norm(const float* xyz, const void* idx, int idx_bits, int B, int N, int k, float* out,
                                  tgp_stream_t stream) {
    if (!xyz || !idx || !out) return fail(TGP_EINVAL, "tgp_direction_norm: null pointer");
    if (B <= 0 || N <= 0 || k <= 0) return fail(TGP_EINVAL, "tgp_direction_norm: sizes must be positive");
    const long total = (long)B * N * k;
    const int threads = 256;
    TGP_DISPATCH_IDX(idx_bits, {
        direction_norm_kernel<IdxT><<<(unsigned)((total + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
            xyz, (const IdxT*)idx, total, N, k, out);
    });
    return check_launch("direction_norm_kernel");
}

extern "C" int tgp_gather_max(const float* f, const void* idx, int idx_bits, const int64_t* rows, int B, int N, int M,
                              int k, int C, float* out, uint8_t* arg, tgp_stream_t stream) {
    if (!f || !idx || !out) return fail(TGP_EINVAL, "tgp_gather_max: null pointer");
    if (B <= 0 || N <= 0 || M <= 0 || k <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_gather_max: sizes must be positive");
    if (!rows && M != N) return fail(TGP_EINVAL, "tgp_gather_max: rows == NULL requires M == N");
    if (k > 255) return fail(TGP_EINVAL, "tgp_gather_max: k > 255");
    const bool vec = (C % 4 == 0) && ((uintptr_t)f % 16 == 0);
    const long total = (long)B * M * (vec ? C / 4 : C);
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    cudaStream_t st = as_stream(stream);
    TGP_DISPATCH_IDX(idx_bits, {
        const IdxT* ip = (const IdxT*)idx;
        if (vec && arg) gather_max_kernel<IdxT, 4, true><<<blocks, threads, 0, st>>>(f, ip, rows, N, M, k, C, total, out, arg);
        else if (vec) gather_max_kernel<IdxT, 4, false><<<blocks, threads, 0, st>>>(f, ip, rows, N, M, k, C, total, out, arg);
        else if (arg) gather_max_kernel<IdxT, 1, true><<<blocks, threads, 0, st>>>(f, ip, rows, N, M, k, C, total, out, arg);
        else gather_max_kernel<IdxT, 1, false><<<blocks, threads, 0, st>>>(f, ip, rows, N, M, k, C, total, out, arg);
    });
    return check_launch("gather_max_kernel");
}

static int orl_nsplit(int N) {   // depends on N only, so a cloud's result is independent of the batch it is in
    int s = (N + 63) / 64;
    return s < 1 ? 1 : (s > 16 ? 16 : s);
}

extern "C" size_t tgp_orl_workspace(int B, int N, int C) { return (size_t)B * orl_nsplit(N) * C * sizeof(float); }

extern "C" int tgp_orl_global(const float* f, const void* idx, int idx_bits, int B, int N, int k, int C, float* g,
                              uint8_t* arg, void* workspace, size_t workspace_bytes, tgp_stream_t stream) {
    if (!f || !idx || !g || !workspace) return fail(TGP_EINVAL, "tgp_orl_global: null pointer");
    if (workspace_bytes < tgp_orl_workspace(B, N, C)) return fail(TGP_ENOSPACE, "tgp_orl_global: workspace too small");
    if (B <= 0 || N <= 0 || k <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_orl_global: sizes must be positive");
    if (k > 255 || B > 65535) return fail(TGP_EINVAL, "tgp_orl_global: k > 255 or B > 65535");
    cudaStream_t st = as_stream(stream);
    const int chunks = (C + 31) / 32;
    const size_t tab_bytes = (size_t)N * 32 * sizeof(float);
    if (tab_bytes <= 200 * 1024) {
        // the cloud's channel slice fits in shared memory (N <= 1600): one pass, no partial sums
        dim3 grid(chunks, B);
        // warps: the gather is a chain of dependent shared-memory reads -- a large table (one CTA per SM) gets all 32 warps;
        // depends on N only, so a cloud's result does not change with the batch it sits in
        int threads = N >= 512 ? 1024 : 512;
        if (const char* e = getenv("TGP_ORL_THREADS")) { const int v = atoi(e); if (v >= 128 && v <= 1024 && v % 32 == 0) threads = v; }   // tuning hook
        TGP_DISPATCH_IDX(idx_bits, {
            if (arg) {
                cudaFuncSetAttribute(orl_smem_kernel<IdxT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                orl_smem_kernel<IdxT, true><<<grid, threads, tab_bytes, st>>>(f, (const IdxT*)idx, N, k, C, g, arg);
            } else {
                cudaFuncSetAttribute(orl_smem_kernel<IdxT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                orl_smem_kernel<IdxT, false><<<grid, threads, tab_bytes, st>>>(f, (const IdxT*)idx, N, k, C, g, arg);
            }
        });
        return check_launch("orl_smem_kernel");
    }
    const int nsplit = orl_nsplit(N);
    float* partial = static_cast<float*>(workspace);
    dim3 grid(chunks, B, nsplit);
    TGP_DISPATCH_IDX(idx_bits, {
        if (arg) orl_global_kernel<IdxT, true><<<grid, ORL_THREADS, 0, st>>>(f, (const IdxT*)idx, N, k, C, partial, arg);
        else orl_global_kernel<IdxT, false><<<grid, ORL_THREADS, 0, st>>>(f, (const IdxT*)idx, N, k, C, partial, arg);
    });
    int rc = check_launch("orl_global_kernel");
    if (rc) return rc;
    const long total = (long)B * C;
    orl_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, nsplit, C, total, N, g);
    return check_launch("orl_finalize_kernel");
}

extern "C" int tgp_concat_rows(const tgp_concat_src* srcs_host, int nsrc, int B, int N, float* out_raw, long ld_raw,
                               float* out_split, int Kp, int mixed, tgp_stream_t stream) {
    if (!srcs_host || (!out_raw && !out_split)) return fail(TGP_EINVAL, "tgp_concat_rows: null pointer");
    if (out_split && mixed && Kp % 64) return fail(TGP_EINVAL, "tgp_concat_rows: a mixed operand needs Kp % 64 == 0");
    if (nsrc < 1 || nsrc > 8 || B <= 0 || N <= 0) return fail(TGP_EINVAL, "tgp_concat_rows: bad sizes");
    ConcatDev P;
    P.nsrc = nsrc;
    int total = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (!srcs_host[i].ptr || srcs_host[i].C <= 0 || srcs_host[i].n_src < 0) return fail(TGP_EINVAL, "tgp_concat_rows: bad source");
        P.s[i] = srcs_host[i];
        total += srcs_host[i].C;
    }
    if (out_split && (Kp < total || Kp % 4)) return fail(TGP_EINVAL, "tgp_concat_rows: Kp smaller than the concatenated width");
    if (out_raw && ld_raw < total) return fail(TGP_EINVAL, "tgp_concat_rows: ld_raw smaller than the concatenated width");
    const long rows = (long)B * N;
    concat_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(P, N, out_raw, ld_raw, out_split, Kp, mixed, rows);
    return check_launch("concat_rows_kernel");
}
