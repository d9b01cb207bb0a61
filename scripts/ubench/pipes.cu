// pipes.cu -- issue-rate microbenchmarks for the instruction mixes the layer-conv redesign depends on (sm_100a).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run: ./pipes
// Output: warp-instructions per clock per SM for each mix (1 CTA of `warps` warps per SM, all 148 SMs).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define ITERS 2048

__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c)); return a;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b)); return a;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b)); return a;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c)); return a;
}

// mode 0: scalar FFMA x8 independent chains; 1: FFMA2 x8; 2: FMNMX x8; 3: FMNMX3 x8; 4: 4 FFMA2 + 4 FMNMX; 5: FMUL2 x8
// 6: 4 FFMA + 4 FMNMX ; 7: 6 FFMA2 + 4 FMNMX + 2 FMUL2 + 2 FMNMX3 (the planned inner mix without loads)
template <int MODE>
__global__ void __launch_bounds__(1024) alu_kernel(float* out, long long* cyc, float seed) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; p[i] = f2_pack(a[i], a[i] + 0.5f); }
    const float b = seed * 0.999f, c = seed * 0.001f;
    const unsigned long long pb = f2_pack(b, b), pc = f2_pack(c, c);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
      for (int rep = 0; rep < 4; ++rep) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pb, pc);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmax3(a[i], b, c);
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { p[i] = ffma2(p[i], pb, pc); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fmul2(p[i], pb);
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i + 4]) : "f"(b), "f"(c)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); }
        } else if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], pb);
        } else if (MODE == 9) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        } else if (MODE == 10) {   // chamfer mix per 2 candidates x 1 query: 3 FADD2 + FMUL2 + 2 FFMA2 + FMNMX3
            p[0] = fadd2(p[0], pb); p[1] = fadd2(p[1], pb); p[2] = fadd2(p[2], pb);
            p[3] = fmul2(p[3], pb); p[4] = ffma2(p[4], pb, pc); p[5] = ffma2(p[5], pb, pc);
            a[0] = fmax3(a[0], b, c);
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 6; ++i) p[i] = ffma2(p[i], pb, pc);
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            p[6] = fmul2(p[6], pb); p[7] = fmul2(p[7], pb);
            a[4] = fmax3(a[4], b, c); a[5] = fmax3(a[5], b, c);
        }
      }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) { s += a[i]; s += __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory loads.  mode 0: LDS.32 conflict-free (lane-consecutive); 1: LDS.128 conflict-free (lane-consecutive 16 B);
// 2: LDS.128 broadcast (all lanes one address); 3: LDS.128, two half-warps two addresses; 4: LDS.64 consecutive;
// 5: LDS.128 rows of 128 B at random row, lane-rotated chunk (the planned gather); 6: LDS.128 random rows, SAME chunk (8-way conflict)
// 7: LDS.128 random 112-byte rows, chunk i (unrotated; the naive thread-per-point gather)
template <int MODE>
__global__ void __launch_bounds__(1024) lds_kernel(float* out, long long* cyc, const int* rows, int nrows) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* f = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < (nrows + 40) * 32; i += blockDim.x) f[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t addr[8];
    for (int i = 0; i < 8; ++i) {
        const int r = rows[(threadIdx.x * 8 + i) % 4096] % nrows;
        if (MODE == 0) addr[i] = base + (i * 32 + lane) * 4;
        else if (MODE == 1) addr[i] = base + (i * 32 + lane) * 16;
        else if (MODE == 2) addr[i] = base + i * 16 + (r & ~0xffff);
        else if (MODE == 3) addr[i] = base + (i * 2 + (lane >> 4)) * 16 + (r & ~0xffff);
        else if (MODE == 4) addr[i] = base + (i * 32 + lane) * 8;
        else if (MODE == 5) addr[i] = base + r * 128 + (((i + lane) & 7) << 4);
        else if (MODE == 6) addr[i] = base + r * 128 + (i << 4);
        else addr[i] = base + r * 112 + ((i % 7) << 4);
    }
    float acc = 0.f;
    uint32_t iacc = 0;
    uint32_t tog = 0;     // toggles between two copies of the address pattern (0 / +off) so the loads cannot be hoisted
    const uint32_t off = (MODE == 0) ? 1024u : (MODE == 4 ? 2048u : 4096u);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr[i] + tog) : "memory"); acc += v; }
            else if (MODE == 4) { float v, w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(addr[i] + tog) : "memory"); acc += v + w; }
            else { uint32_t x, y, z, w; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr[i] + tog) : "memory"); iacc ^= x ^ y; iacc ^= z ^ w; }
        }
        tog ^= off;
        asm volatile("" : "+r"(tog));
      }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(iacc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static double run(F launch, int blocks, int threads, int instr_per_iter, long long* d_cyc) {
    launch();
    cudaDeviceSynchronize();
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 0; }
    std::vector<long long> h(blocks);
    cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double cyc = (double)h[blocks / 2];
    return (double)(threads / 32) * ITERS * instr_per_iter / cyc;   // warp-instructions per clock per SM
}

int main() {
    float* d_out; long long* d_cyc; int* d_rows;
    const int blocks = 148;
    cudaMalloc(&d_out, blocks * 1024 * sizeof(float));
    cudaMalloc(&d_cyc, blocks * sizeof(long long));
    std::vector<int> rows(4096);
    uint32_t s = 12345;
    for (auto& r : rows) { s = s * 1664525u + 1013904223u; r = (s >> 8) % 1028; }
    cudaMalloc(&d_rows, rows.size() * sizeof(int));
    cudaMemcpy(d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice);
    const char* alu_names[] = {"FFMA x8", "FFMA2 x8", "FMNMX x8", "FMNMX3 x8", "4 FFMA2 + 4 FMNMX", "FMUL2 x8", "4 FFMA + 4 FMNMX", "6 FFMA2+4 FMNMX+2 FMUL2+2 FMNMX3", "FADD2 x8", "FADD x8", "3 FADD2+FMUL2+2 FFMA2+FMNMX3"};
    const int alu_ipi[] = {8, 8, 8, 8, 8, 8, 8, 14, 8, 8, 7};
    for (int threads : {256, 512, 1024}) {
        printf("== ALU mixes, %d threads/SM (warp-instr / clk / SM; 4.0 = one per SMSP per clock)\n", threads);
#define ALU(M) printf("  %-36s %.3f\n", alu_names[M], run([&] { alu_kernel<M><<<blocks, threads>>>(d_out, d_cyc, 1.0001f); }, blocks, threads, alu_ipi[M], d_cyc));
        ALU(0) ALU(1) ALU(2) ALU(3) ALU(4) ALU(5) ALU(6) ALU(7) ALU(8) ALU(9) ALU(10)
    }
    const char* lds_names[] = {"LDS.32 consecutive", "LDS.128 consecutive", "LDS.128 broadcast", "LDS.128 2 addresses (half-warps)", "LDS.64 consecutive",
                               "LDS.128 random 128B rows, rotated chunk", "LDS.128 random 128B rows, same chunk", "LDS.128 random 112B rows, chunk i"};
    const int nrows = 1028;
    const size_t smem = (nrows + 40) * 128;
#define LDSK(M) cudaFuncSetAttribute(lds_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LDSK(0) LDSK(1) LDSK(2) LDSK(3) LDSK(4) LDSK(5) LDSK(6) LDSK(7)
    for (int threads : {256, 512, 1024}) {
        printf("== LDS, %d threads/SM (warp-instr / clk / SM)\n", threads);
#define LDSR(M) printf("  %-44s %.3f\n", lds_names[M], run([&] { lds_kernel<M><<<blocks, threads, smem>>>(d_out, d_cyc, d_rows, nrows); }, blocks, threads, 8, d_cyc));
        LDSR(0) LDSR(1) LDSR(2) LDSR(3) LDSR(4) LDSR(5) LDSR(6) LDSR(7)
    }
    return 0;
}
