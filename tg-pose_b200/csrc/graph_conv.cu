// graph_conv.cu -- fused support-direction graph convolutions for sm_100a.
//
// Replaces HSlayer_surface.graph_conv (gcn3d.py:91-106) and HS_layer.graph_conv after the
// projection (gcn3d.py:157-180).  The reference materialises theta (B,N,k,S*C), the gathered
// support (B,N,k,S*C) and their product in memory (~0.5 GB per cloud per layer of traffic);
// here gather, direction cosine, ReLU, multiply, max-over-neighbours and mean-over-supports
// happen in registers and the only HBM traffic is the compulsory input/output.
//
// Layer conv data layout: the projection GEMM writes the support features channel-group-major
// ("slab": [C/4][B*N][S][4]).  A CTA owns (cloud b, 4-channel group): its whole neighbour
// table -- N rows of S*4 floats, 115 KB at N=1028,S=7 -- is ONE contiguous block that a TMA
// bulk copy (cp.async.bulk, mbarrier-tracked) drops into shared memory; every gather after that
// is a conflict-free 112-byte shared-memory row read by lanes 0..27 of a warp.
#include "common.cuh"
#include <float.h>

namespace tgp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();  // a lost TMA must fault, not hang the box
    }
}

// column-normalised support direction (F.normalize(directions, dim=0), gcn3d.py:99,165)
__device__ __forceinline__ void load_sd(const float* __restrict__ directions, int SC, int col, float& x, float& y, float& z) {
    x = __ldg(directions + col);
    y = __ldg(directions + SC + col);
    z = __ldg(directions + 2 * SC + col);
    normalize3(x, y, z);
}

// result store: raw row-major and, optionally, the [tf32 | residual] operand of the next contraction
__device__ __forceinline__ void store_out(float* __restrict__ out, float* __restrict__ out_split, int kp, long pt,
                                          int C, int c, float v) {
    out[pt * C + c] = v;
    if (out_split) {
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        out_split[pt * 2 * kp + c] = hi;
        out_split[pt * 2 * kp + kp + c] = v - hi;
    }
}

// ------------------------------------------------------------------------------------------
// edge records: (dx,dy,dz, idx) per (b,n,j)
template <typename IdxT>
__global__ void edge_record_kernel(const float* __restrict__ xyz, const IdxT* __restrict__ idx, long total, int N,
                                   int k, float4* __restrict__ rec) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long pt = e / k;
    const long b = pt / N;
    const int nb = ld_idx(idx, e);
    const float* p = xyz + (b * N + nb) * 3;
    const float* c = xyz + pt * 3;
    float x = __ldg(p) - __ldg(c), y = __ldg(p + 1) - __ldg(c + 1), z = __ldg(p + 2) - __ldg(c + 2);
    normalize3(x, y, z);
    rec[e] = make_float4(x, y, z, __int_as_float(nb));
}

// ------------------------------------------------------------------------------------------
// surface conv: out[b,n,c] = mean_s max_j relu(<d_j, sd[:,s,c]>).
// CTA = SURF_PTS points of one cloud; normalised support directions staged once per CTA;
// a warp owns a point: lanes 0..k-1 build the k unit directions, then lane = channel (mod 32)
// runs all S supports of its channels against the broadcast directions.
constexpr int SURF_THREADS = 256;
constexpr int SURF_PTS = 32;

template <typename IdxT, int S_T, bool ARG>
__global__ void __launch_bounds__(SURF_THREADS)
surface_conv_kernel(const float* __restrict__ xyz, const IdxT* __restrict__ idx, const float* __restrict__ directions,
                    int N, int k, int S_rt, int C, float* __restrict__ out, uint8_t* __restrict__ arg,
                    float* __restrict__ out_split, int kp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = S_T > 0 ? S_T : S_rt;
    const int SC = S * C;
    float* sd = reinterpret_cast<float*>(smem_raw);                 // [3][SC]
    float4* dirs = reinterpret_cast<float4*>(sd + 3 * SC + ((4 - (3 * SC) % 4) % 4));  // [warps][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long b = blockIdx.y;
    for (int col = threadIdx.x; col < SC; col += SURF_THREADS) {
        float x, y, z;
        load_sd(directions, SC, col, x, y, z);
        sd[col] = x; sd[SC + col] = y; sd[2 * SC + col] = z;
    }
    __syncthreads();
    float4* my = dirs + warp * k;
    const float inv_s = 1.0f / (float)S;
    for (int pl = warp; pl < SURF_PTS; pl += SURF_THREADS / 32) {
        const int n = blockIdx.x * SURF_PTS + pl;
        if (n >= N) break;
        const long pt = b * N + n;
        __syncwarp();
        const float cx = __ldg(xyz + pt * 3), cy = __ldg(xyz + pt * 3 + 1), cz = __ldg(xyz + pt * 3 + 2);
        for (int j = lane; j < k; j += 32) {
            const float* p = xyz + (b * N + ld_idx(idx, pt * k + j)) * 3;
            float x = __ldg(p) - cx, y = __ldg(p + 1) - cy, z = __ldg(p + 2) - cz;
            normalize3(x, y, z);
            my[j] = make_float4(x, y, z, 0.f);
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
            if (S_T > 0) {
                float sx[S_T > 0 ? S_T : 1], sy[S_T > 0 ? S_T : 1], sz[S_T > 0 ? S_T : 1], m[S_T > 0 ? S_T : 1];
                int a[S_T > 0 ? S_T : 1];
#pragma unroll
                for (int s = 0; s < S_T; ++s) {
                    sx[s] = sd[s * C + c]; sy[s] = sd[SC + s * C + c]; sz[s] = sd[2 * SC + s * C + c];
                    m[s] = 0.f; a[s] = 0;
                }
#pragma unroll 2
                for (int j = 0; j < k; ++j) {
                    const float4 d = my[j];
#pragma unroll
                    for (int s = 0; s < S_T; ++s) {
                        const float th = fmaf(d.z, sz[s], fmaf(d.y, sy[s], d.x * sx[s]));
                        if (ARG) { if (th > m[s]) { m[s] = th; a[s] = j; } }
                        else m[s] = fmaxf(m[s], th);
                    }
                }
                float acc = 0.f;
#pragma unroll
                for (int s = 0; s < S_T; ++s) {
                    acc += m[s];
                    if (ARG) arg[pt * SC + s * C + c] = (uint8_t)a[s];
                }
                store_out(out, out_split, kp, pt, C, c, acc * inv_s);
            } else {
                float acc = 0.f;
                for (int s = 0; s < S; ++s) {
                    const float sx = sd[s * C + c], sy = sd[SC + s * C + c], sz = sd[2 * SC + s * C + c];
                    float m = 0.f;
                    int a = 0;
                    for (int j = 0; j < k; ++j) {
                        const float4 d = my[j];
                        const float th = fmaf(d.z, sz, fmaf(d.y, sy, d.x * sx));
                        if (th > m) { m = th; a = j; }
                    }
                    acc += m;
                    if (ARG) arg[pt * SC + s * C + c] = (uint8_t)a;
                }
                store_out(out, out_split, kp, pt, C, c, acc * inv_s);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// layer conv.  CTA = (4-channel group cg, cloud b).  lane = s*4 + c4 (< S*4 <= 32).
// TAB: the (cloud, channel-group) support table is staged in shared memory (N*S*16 B <= ~215 KB, i.e. N <= ~1960 at
// S = 7); otherwise (the N = 2048..16384 microbenchmark clouds) the 112-byte rows are gathered through L2.
template <bool ARG, int KT, bool TAB>
__global__ void __launch_bounds__(1024)
layer_conv_kernel(const float4* __restrict__ rec, const float* __restrict__ directions,
                  const float* __restrict__ centre, long ld_centre, const float* __restrict__ slab,
                  long M, int N, int k_rt, int S, int C, float* __restrict__ out, uint8_t* __restrict__ arg_slab,
                  float* __restrict__ out_split, int kp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int k = KT > 0 ? KT : k_rt;                      // neighbour count: compile-time for the encoder's 20 / 8
    const int W = S * 4;                                   // slab row width in floats
    float* tab = reinterpret_cast<float*>(smem_raw);       // [N][W]
    const size_t tab_bytes = TAB ? (size_t)N * W * sizeof(float) : 0;
    float4* recs = reinterpret_cast<float4*>(smem_raw + ((tab_bytes + 15) & ~(size_t)15));  // [warps][k]
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cg = blockIdx.x;
    const long b = blockIdx.y;
    const int SC = S * C;

    // one elected thread arms the barrier and issues the bulk copies (<= 32 KB pieces)
    const float* src = slab + ((long)cg * M + b * N) * W;
    if (TAB && threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (TAB && threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)tab_bytes);
        const uint32_t piece = 32768;
        for (uint32_t off = 0; off < tab_bytes; off += piece) {
            const uint32_t nbytes = (uint32_t)min((size_t)piece, tab_bytes - off);
            bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char*>(src) + off, nbytes, &bar);
        }
    }
    // this lane's support direction while the table is in flight
    float sx = 0.f, sy = 0.f, sz = 0.f;
    const int s_l = lane >> 2, c4 = lane & 3;
    if (lane < W) load_sd(directions, SC, s_l * C + cg * 4 + c4, sx, sy, sz);
    if (TAB) mbar_wait(&bar, 0);

    float4* my = recs + warp * k;
    const float inv_s = 1.0f / (float)S;
    // byte address of this lane's column in table row 0; a neighbour's row is one integer add away
    const uint32_t tab_lane = smem_u32(tab) + (uint32_t)(lane < W ? lane : 0) * 4u;
    const uint32_t row_bytes = (uint32_t)W * 4u;
    // the k edge records of the warp's next point are fetched while the current point is being reduced
    // gridDim.z splits the points of the cloud when (clouds x channel groups) alone cannot fill the machine
    const int per = (N + gridDim.z - 1) / gridDim.z;
    const int n_beg = blockIdx.z * per, n_end = min(N, n_beg + per);
    const float4* rp = rec + (b * N + n_beg + warp) * (long)k;
    const long rstep = (long)nwarps * k;
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_beg + warp < n_end && lane < k) nxt = __ldg(rp + lane);
    for (int n = n_beg + warp; n < n_end; n += nwarps) {
        const long pt = b * N + n;
        __syncwarp();
        if (lane < k) my[lane] = nxt;
        for (int j = lane + 32; j < k; j += 32) my[j] = __ldg(rp + j);
        __syncwarp();
        rp += rstep;
        if (n + nwarps < n_end && lane < k) nxt = __ldg(rp + lane);
        // the centre term is only needed after the neighbour loop: fetch it now so its latency hides behind the loop
        float cen = 0.f;
        if (lane < 4) cen = __ldg(centre + pt * ld_centre + cg * 4 + lane);
        float m = -FLT_MAX;
        int a = 0;
        if (lane < W) {
#pragma unroll (KT > 0 ? KT : 4)
            for (int j = 0; j < k; ++j) {
                const float4 d = my[j];
                float sup;
                if (TAB) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sup) : "r"(tab_lane + (uint32_t)__float_as_int(d.w) * row_bytes));
                else sup = __ldg(src + (long)__float_as_int(d.w) * W + (lane < W ? lane : 0));
                const float th = fmaxf(fmaf(d.z, sz, fmaf(d.y, sy, d.x * sx)), 0.f);
                const float v = th * sup;
                if (ARG) { if (v > m) { m = v; a = j; } }
                else m = fmaxf(m, v);
            }
            if (ARG) arg_slab[((long)cg * M + pt) * W + lane] = (uint8_t)a;
        } else m = 0.f;
        // sum over supports: lanes with equal c4 (stride 4)
        m += __shfl_down_sync(0xffffffffu, m, 16);
        m += __shfl_down_sync(0xffffffffu, m, 8);
        m += __shfl_down_sync(0xffffffffu, m, 4);
        if (lane < 4) store_out(out, out_split, kp, pt, C, cg * 4 + lane, cen + m * inv_s);
    }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_edge_records(const float* xyz, const void* idx, int idx_bits, int B, int N, int k, float* rec,
                                tgp_stream_t stream) {
    if (!xyz || !idx || !rec) return fail(TGP_EINVAL, "tgp_edge_records: null pointer");
    if (B <= 0 || N <= 0 || k <= 0) return fail(TGP_EINVAL, "tgp_edge_records: sizes must be positive");
    if ((uintptr_t)rec % 16) return fail(TGP_EINVAL, "tgp_edge_records: rec must be 16-byte aligned");
    const long total = (long)B * N * k;
    const int threads = 256;
    TGP_DISPATCH_IDX(idx_bits, {
        edge_record_kernel<IdxT><<<(unsigned)((total + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
            xyz, (const IdxT*)idx, total, N, k, reinterpret_cast<float4*>(rec));
    });
    return check_launch("edge_record_kernel");
}

template <typename IdxT, int S_T>
static int launch_surface(const float* xyz, const IdxT* idx, const float* directions, int B, int N, int k, int S, int C,
                          float* out, uint8_t* arg, float* out_split, cudaStream_t st) {
    const int kp = (C + 31) / 32 * 32;
    const int SC = S * C;
    const size_t smem = sizeof(float) * (3 * SC + 4) + sizeof(float4) * (SURF_THREADS / 32) * k;
    if (smem > 227 * 1024) return fail(TGP_EINVAL, "tgp_surface_conv_fwd: S*C too large for shared memory");
    dim3 grid((N + SURF_PTS - 1) / SURF_PTS, B);
    if (arg) {
        cudaFuncSetAttribute(surface_conv_kernel<IdxT, S_T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        surface_conv_kernel<IdxT, S_T, true><<<grid, SURF_THREADS, smem, st>>>(xyz, idx, directions, N, k, S, C, out, arg, out_split, kp);
    } else {
        cudaFuncSetAttribute(surface_conv_kernel<IdxT, S_T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        surface_conv_kernel<IdxT, S_T, false><<<grid, SURF_THREADS, smem, st>>>(xyz, idx, directions, N, k, S, C, out, arg, out_split, kp);
    }
    return check_launch("surface_conv_kernel");
}

extern "C" int tgp_surface_conv_fwd(const float* xyz, const void* idx, int idx_bits, const float* directions, int B,
                                    int N, int k, int S, int C, float* out, uint8_t* arg, float* out_split,
                                    tgp_stream_t stream) {
    if (!xyz || !idx || !directions || !out) return fail(TGP_EINVAL, "tgp_surface_conv_fwd: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || S <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_surface_conv_fwd: sizes must be positive");
    if (k > 255 || B > 65535) return fail(TGP_EINVAL, "tgp_surface_conv_fwd: k > 255 or B > 65535");
    cudaStream_t st = as_stream(stream);
    TGP_DISPATCH_IDX(idx_bits, {
        if (S == 7) return launch_surface<IdxT, 7>(xyz, (const IdxT*)idx, directions, B, N, k, S, C, out, arg, out_split, st);
        return launch_surface<IdxT, 0>(xyz, (const IdxT*)idx, directions, B, N, k, S, C, out, arg, out_split, st);
    });
    return TGP_OK;
}

extern "C" int tgp_layer_conv_fwd(const float* edge_rec, const float* directions, const float* centre, long ld_centre,
                                  const float* support_slab, int B, int N, int k, int S, int C, float* out,
                                  uint8_t* arg_slab, float* out_split, tgp_stream_t stream) {
    if (!edge_rec || !directions || !centre || !support_slab || !out) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || S <= 0 || C <= 0) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: sizes must be positive");
    if (C % 4) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: C must be a multiple of 4 (slab layout)");
    if (S * 4 > 32) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: S > 8 unsupported");
    if (k > 255 || B > 65535) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: k > 255 or B > 65535");
    if ((uintptr_t)support_slab % 16 || (uintptr_t)edge_rec % 16) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: slab / edge_rec must be 16-byte aligned");
    const int W = S * 4;
    const size_t tab_bytes = (size_t)N * W * sizeof(float);
    // enough warps to hide the gather latency, few enough that several CTAs share an SM when the table is small
    int threads = N >= 512 ? 1024 : (N >= 128 ? 256 : 128);
    size_t smem = ((tab_bytes + 15) & ~(size_t)15) + sizeof(float4) * (threads / 32) * k;
    const bool tab = smem <= 227 * 1024;        // else: gather the support rows through L2 instead of a shared-memory table
    if (!tab) smem = sizeof(float4) * (threads / 32) * k;
    if (smem > 227 * 1024) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: k too large");
    // point splits: aim at >= 2 CTAs per SM; depends on the shapes only.  (Each split re-stages the table when TAB.)
    int zs = (2 * TGP_NUM_SMS + (C / 4) * B - 1) / ((C / 4) * B);
    const int zcap = (N + 255) / 256;
    if (zs > zcap) zs = zcap;
    if (zs < 1) zs = 1;
    if (zs > 64) zs = 64;
    dim3 grid(C / 4, B, zs);
    cudaStream_t st = as_stream(stream);
    const long M = (long)B * N;
    const int kp = (C + 31) / 32 * 32;
#define TGP_LAUNCH_LC(ARGV, KTV, TABV)                                                                                   \
    do {                                                                                                                \
        cudaFuncSetAttribute(layer_conv_kernel<ARGV, KTV, TABV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        layer_conv_kernel<ARGV, KTV, TABV><<<grid, threads, smem, st>>>(reinterpret_cast<const float4*>(edge_rec),       \
                                                                        directions, centre, ld_centre, support_slab, M,  \
                                                                        N, k, S, C, out, arg_slab, out_split, kp);       \
    } while (0)
    if (!tab) {
        if (arg_slab) TGP_LAUNCH_LC(true, 0, false); else TGP_LAUNCH_LC(false, 0, false);
    } else if (arg_slab) {
        if (k == 20) TGP_LAUNCH_LC(true, 20, true); else if (k == 8) TGP_LAUNCH_LC(true, 8, true); else TGP_LAUNCH_LC(true, 0, true);
    } else {
        if (k == 20) TGP_LAUNCH_LC(false, 20, true); else if (k == 8) TGP_LAUNCH_LC(false, 8, true); else TGP_LAUNCH_LC(false, 0, true);
    }
#undef TGP_LAUNCH_LC
    return check_launch("layer_conv_kernel");
}
