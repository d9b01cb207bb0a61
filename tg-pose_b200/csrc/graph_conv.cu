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
#include <stdlib.h>

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
// packed fp32 pairs (FFMA2 / FMUL2 on sm_100a): two channels per instruction -- same .rn arithmetic as the scalar
// forms, half the issue slots.  A pair built from one scalar twice is encoded by ptxas as a broadcast operand.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ------------------------------------------------------------------------------------------
// edge records: (dx,dy,dz, idx) per (b,j,n) -- NEIGHBOUR-major, so the points of a warp's quad read contiguous bytes
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
    const long n = pt - b * N;
    rec[(b * k + (e - pt * k)) * N + n] = make_float4(x, y, z, __int_as_float(nb));
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
        // two channels (c, c + 32) per lane and iteration as one packed pair: the cosine of a neighbour against a support is
        // FMUL2 + 2 FFMA2 for both channels -- 5 instead of 8 issue slots per 2 elements (the scalar loop ran at IPC 2.8 with the
        // FMA pipe 47 % busy: issue-bound)
        if (S_T > 0 && !ARG && (C & 63) == 0) {
            for (int c = lane; c < C; c += 64) {
                unsigned long long sx[S_T > 0 ? S_T : 1], sy[S_T > 0 ? S_T : 1], sz[S_T > 0 ? S_T : 1];
                float ma[S_T > 0 ? S_T : 1], mb[S_T > 0 ? S_T : 1];
#pragma unroll
                for (int s = 0; s < S_T; ++s) {
                    sx[s] = f2_pack(sd[s * C + c], sd[s * C + c + 32]);
                    sy[s] = f2_pack(sd[SC + s * C + c], sd[SC + s * C + c + 32]);
                    sz[s] = f2_pack(sd[2 * SC + s * C + c], sd[2 * SC + s * C + c + 32]);
                    ma[s] = 0.f; mb[s] = 0.f;
                }
#pragma unroll 2
                for (int j = 0; j < k; ++j) {
                    const float4 d = my[j];
                    const unsigned long long dx = f2_pack(d.x, d.x), dy = f2_pack(d.y, d.y), dz = f2_pack(d.z, d.z);
#pragma unroll
                    for (int s = 0; s < S_T; ++s) {
                        float ta, tb;
                        f2_unpack(f2_fma(dz, sz[s], f2_fma(dy, sy[s], f2_mul(dx, sx[s]))), ta, tb);
                        ma[s] = fmaxf(ma[s], ta); mb[s] = fmaxf(mb[s], tb);
                    }
                }
                float acca = 0.f, accb = 0.f;
#pragma unroll
                for (int s = 0; s < S_T; ++s) { acca += ma[s]; accb += mb[s]; }
                store_out(out, out_split, kp, pt, C, c, acca * inv_s);
                store_out(out, out_split, kp, pt, C, c + 32, accb * inv_s);
            }
            continue;
        }
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

// 16-byte asynchronous global -> shared copy (LDGSTS, L1 bypassed) and its group bookkeeping
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// ------------------------------------------------------------------------------------------
// layer conv.  CTA = (4-channel group cg, cloud b[, point split z]).
// A warp works on FOUR points at a time: lane = p*8 + s (point p of the quad, support s < S); a lane owns the 4 channels
// of its (point, support) -- one 16-byte chunk of a slab row.  Per neighbour a lane issues ONE 16-byte record load
// (direction + neighbour index, from a per-warp shared-memory ring that cp.async fills one quad ahead; the records are
// neighbour-major in HBM, so a quad's k x 64 bytes are k contiguous pieces), ONE 16-byte gather of its chunk of the
// neighbour's slab row (the 8 lanes of a quarter-warp read one contiguous 112-byte row: conflict-free) and 16 arithmetic
// instructions for its 4 (neighbour, support, channel) elements (cosine as packed FMUL2/FFMA2, ReLU, packed multiply,
// max).  What bounded the round-1 mapping (lane = (support, channel), one point per warp) was the shared-memory pipe:
// one broadcast record load (2 cycles) + one 4-byte gather (1 cycle) per 28 elements = 0.107 cycles per element; this
// one spends 4 (gather) + ~3.2 (record) cycles per 112 elements = 0.064 (profiles/r02_layer_conv.md; ncu: LSU data pipe
// 81 % busy, i.e. the kernel now sits on that roof).  Loading the records straight from global memory into registers
// (prefetch distance 4) was 1.9x slower: ptxas sinks the loads next to their uses under the 64-register cap.
// Also tried and dropped (round 2): a lane owning HALF a 128-byte-padded row (4 rotated chunks, 16 points per warp, one
// record load per 448 elements: 0.042 pipe cycles per element on paper).  Its 48 direction constants + 16 running maxima
// need ~128 registers, i.e. 13 warps per SM: 322 us against this kernel's 207 us at 32 x 1028 x 128 (ncu: 18 % of the warp
// slots occupied, every warp waiting on its own fixed-latency chain).  A support PAIR per lane (8 points per warp, 72
// registers, 26 warps) ran correctly at 262 us: it executes MORE instructions per element than this kernel (0.29 vs 0.23:
// its per-group ring bookkeeping and register-pair moves outweigh the shared record load) at 61 % pipe utilisation.
// TAB: the (cloud, channel-group) support table is staged in shared memory by TMA bulk copies (N*S*16 B <= ~227 KB,
// i.e. N <= ~2070 at S = 7); otherwise (the N = 2048..16384 microbenchmark clouds) the rows are gathered through L2.
template <bool ARG, int KT, bool TAB>
__global__ void __launch_bounds__(ARG ? 512 : 1024, 1)    // the arg-max slots of the training variant need the registers
layer_conv_kernel(const float4* __restrict__ rec, const float* __restrict__ directions,
                  const float* __restrict__ centre, long ld_centre, const float* __restrict__ slab,
                  long M, int N, int k_rt, int S, int C, float* __restrict__ out, uint8_t* __restrict__ arg_slab,
                  float* __restrict__ out_split, int kp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int k = KT > 0 ? KT : k_rt;                      // neighbour count: compile-time for the encoder's 20 / 8
    const int W = S * 4;                                   // slab row width in floats
    float* tab = reinterpret_cast<float*>(smem_raw);       // [N][W]
    const size_t tab_bytes = TAB ? (size_t)N * W * sizeof(float) : 0;
    // per-warp record ring: [2 quads][k neighbours][4 points] float4, filled by cp.async one quad ahead
    const uint32_t ring_base = smem_u32(smem_raw) + (uint32_t)((tab_bytes + 127) & ~(size_t)127);
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
    // this lane's four support directions (channels cg*4 .. cg*4+3 of support s) while the table is in flight
    const int p = lane >> 3, s = lane & 7;
    const bool act = s < S;
    unsigned long long sx01 = 0, sx23 = 0, sy01 = 0, sy23 = 0, sz01 = 0, sz23 = 0;
    if (act) {
        float x[4], y[4], z[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) load_sd(directions, SC, s * C + cg * 4 + c, x[c], y[c], z[c]);
        sx01 = f2_pack(x[0], x[1]); sx23 = f2_pack(x[2], x[3]);
        sy01 = f2_pack(y[0], y[1]); sy23 = f2_pack(y[2], y[3]);
        sz01 = f2_pack(z[0], z[1]); sz23 = f2_pack(z[2], z[3]);
    }
    const int s_c = act ? s : 0;                           // idle lanes (s >= S) shadow support 0; nothing of theirs is stored
    const uint32_t tab_lane = smem_u32(tab) + (uint32_t)s_c * 16u;
    const uint32_t row_bytes = (uint32_t)W * 4u;
    const float4* src4 = reinterpret_cast<const float4*>(src) + s_c;
    const float inv_s = 1.0f / (float)S;
    // gridDim.z splits the quads of the cloud when (clouds x channel groups) alone cannot fill the machine
    const int quads = (N + 3) >> 2;
    const int per = (quads + gridDim.z - 1) / gridDim.z;
    const int q_beg = blockIdx.z * per, q_end = min(quads, q_beg + per);
    const float4* recb = rec + b * (long)k * N;
    const uint32_t quad_bytes = (uint32_t)k * 64u;
    const uint32_t ring = ring_base + (uint32_t)warp * 2u * quad_bytes;
    // the records of one quad = k rows of 64 contiguous bytes: chunk c = j*4 + p  ->  rec[b][j][4q + p]
    auto fetch_quad = [&](int q, uint32_t dst) {
        for (int c = lane; c < k * 4; c += 32) {
            const int j = c >> 2, n = min(q * 4 + (c & 3), N - 1);
            cp_async16(dst + (uint32_t)c * 16u, recb + (long)j * N + n);
        }
        cp_async_commit();
    };
    int q = q_beg + warp;
    if (q < q_end) fetch_quad(q, ring);
    if (TAB) mbar_wait(&bar, 0);

    for (uint32_t buf = 0; q < q_end; q += nwarps, buf ^= 1u) {
        const int n = q * 4 + p;
        const bool valid = n < N;
        const int n_ld = valid ? n : N - 1;
        const long pt = b * N + n_ld;
        if (q + nwarps < q_end) { fetch_quad(q + nwarps, ring + (buf ^ 1u) * quad_bytes); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const uint32_t rq = ring + buf * quad_bytes + (uint32_t)p * 16u;
        float m0 = -FLT_MAX, m1 = -FLT_MAX, m2 = -FLT_MAX, m3 = -FLT_MAX;
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        // one neighbour: record (broadcast within the point's 8 lanes), gather, 16 arithmetic instructions for the
        // lane's 4 (neighbour, support, channel) elements
        auto step = [&](const int j) {
            float4 d;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "r"(rq + (uint32_t)j * 64u));
            float4 sup;
            if (TAB) {
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(sup.x), "=f"(sup.y), "=f"(sup.z), "=f"(sup.w)
                             : "r"(tab_lane + (uint32_t)__float_as_int(d.w) * row_bytes));
            } else {
                sup = __ldg(src4 + (long)__float_as_int(d.w) * S);
            }
            const unsigned long long dx = f2_pack(d.x, d.x), dy = f2_pack(d.y, d.y), dz = f2_pack(d.z, d.z);
            float t0, t1, t2, t3;
            f2_unpack(f2_fma(dz, sz01, f2_fma(dy, sy01, f2_mul(dx, sx01))), t0, t1);
            f2_unpack(f2_fma(dz, sz23, f2_fma(dy, sy23, f2_mul(dx, sx23))), t2, t3);
            float v0, v1, v2, v3;
            f2_unpack(f2_mul(f2_pack(fmaxf(t0, 0.f), fmaxf(t1, 0.f)), f2_pack(sup.x, sup.y)), v0, v1);
            f2_unpack(f2_mul(f2_pack(fmaxf(t2, 0.f), fmaxf(t3, 0.f)), f2_pack(sup.z, sup.w)), v2, v3);
            if (ARG) {
                if (v0 > m0) { m0 = v0; a0 = j; }
                if (v1 > m1) { m1 = v1; a1 = j; }
                if (v2 > m2) { m2 = v2; a2 = j; }
                if (v3 > m3) { m3 = v3; a3 = j; }
            } else {
                m0 = fmaxf(m0, v0); m1 = fmaxf(m1, v1); m2 = fmaxf(m2, v2); m3 = fmaxf(m3, v3);
            }
        };
        if (KT > 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j) step(j);
        } else {
#pragma unroll 4
            for (int j = 0; j < k; ++j) step(j);
        }
        __syncwarp();                                      // every lane is done with this buffer before it is refilled
        float4 cen = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s == 0) {
            const float* cp = centre + pt * ld_centre + cg * 4;
            cen = make_float4(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2), __ldg(cp + 3));
        }
        if (ARG && act && valid) {
            const uint32_t packed = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16) | ((uint32_t)a3 << 24);
            *reinterpret_cast<uint32_t*>(arg_slab + ((long)cg * M + pt) * W + s * 4) = packed;
        }
        if (!act) { m0 = 0.f; m1 = 0.f; m2 = 0.f; m3 = 0.f; }
        // sum over the supports of a point: butterfly over the 8 lanes of its group (fixed order)
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            m0 += __shfl_xor_sync(0xffffffffu, m0, o);
            m1 += __shfl_xor_sync(0xffffffffu, m1, o);
            m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            m3 += __shfl_xor_sync(0xffffffffu, m3, o);
        }
        if (s == 0 && valid) {
            const float4 r = make_float4(cen.x + m0 * inv_s, cen.y + m1 * inv_s, cen.z + m2 * inv_s, cen.w + m3 * inv_s);
            *reinterpret_cast<float4*>(out + pt * C + cg * 4) = r;
            if (out_split) {
                float4 hi, lo;
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(r.x)); hi.x = __uint_as_float(hb); lo.x = r.x - hi.x;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(r.y)); hi.y = __uint_as_float(hb); lo.y = r.y - hi.y;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(r.z)); hi.z = __uint_as_float(hb); lo.z = r.z - hi.z;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(r.w)); hi.w = __uint_as_float(hb); lo.w = r.w - hi.w;
                *reinterpret_cast<float4*>(out_split + pt * 2 * kp + cg * 4) = hi;
                *reinterpret_cast<float4*>(out_split + pt * 2 * kp + kp + cg * 4) = lo;
            }
        }
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
    cudaStream_t st = as_stream(stream);
    const long M = (long)B * N;
    const int kp = (C + 31) / 32 * 32;
    const int W = S * 4;
    const size_t tab_bytes = (size_t)N * W * sizeof(float);
    const size_t ring_per_warp = (size_t)2 * k * 64;          // two quads of k records x 4 points x 16 B
    const size_t tab_al = (tab_bytes + 127) & ~(size_t)127;
    const bool tab = tab_al + 4 * ring_per_warp <= 227 * 1024;   // else: gather the support rows through L2
    if (4 * ring_per_warp > 227 * 1024) return fail(TGP_EINVAL, "tgp_layer_conv_fwd: k too large");
    // point splits (each split re-stages the table when TAB): only when (channel groups x clouds) is under ~4 waves.
    const int quads = (N + 3) / 4;
    int wmax = arg_slab ? 16 : 32;
    {
        const size_t room = 227 * 1024 - (tab ? tab_al : 0);
        if ((size_t)wmax * ring_per_warp > room) wmax = (int)(room / ring_per_warp);
    }
    int zs = (4 * TGP_NUM_SMS + (C / 4) * B - 1) / ((C / 4) * B);
    const int zcap = (quads + wmax - 1) / wmax;
    if (zs > zcap) zs = zcap;
    if (zs < 1) zs = 1;
    if (zs > 64) zs = 64;
    if (const char* e = getenv("TGP_LC_ZS")) { const int v = atoi(e); if (v > 0 && v <= 64) zs = v; }   // tuning hook
    // warps per CTA (a warp works on one quad of points at a time)
    const int qcta = (quads + zs - 1) / zs;
    // ~8 quads per warp (measured best at N = 1028 / 257 / 64: 32 / 8 / 4 warps): small clouds run as several small
    // resident CTAs per SM so that table loads, set-up and tails of one overlap the neighbour loops of the others
    int nw = (qcta + 7) / 8;
    if (nw > wmax) nw = wmax;
    if (nw < 2) nw = 2;
    if (const char* e = getenv("TGP_LC_WARPS")) { const int v = atoi(e); if (v > 0 && v <= wmax) nw = v; }   // tuning hook
    const int threads = nw * 32;
    const size_t smem = (tab ? tab_al : 0) + (size_t)nw * ring_per_warp;
    dim3 grid(C / 4, B, zs);
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
