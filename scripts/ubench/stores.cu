// store-throughput microbenchmark: how fast can ONE CTA per SM (W warps) push 16-byte stores to global memory?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/stores scripts/ubench/stores.cu && scripts/ubench/stores
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// pattern 0: a warp instruction writes 512 contiguous bytes; pattern 1: 4 rows x 128 B, row pitch `pitch` bytes;
// pattern 2: 32 pieces of 16 B at 128 B pitch (worst case).  Every warp owns a private region; total bytes fixed.
template <int CS>
__global__ void __launch_bounds__(1024, 1) store_kernel(float4* out, long bytes_per_warp, int pattern, long pitch) {
    extern __shared__ char pad[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    char* base = reinterpret_cast<char*>(out) + ((long)blockIdx.x * W + warp) * bytes_per_warp;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
    const long n = bytes_per_warp / 512;
    for (long i = 0; i < n; ++i) {
        char* p;
        if (pattern == 0) p = base + i * 512 + lane * 16;
        else if (pattern == 1) p = base + (i >> 3) * 4 * pitch + (i & 7) * 128 + (lane >> 3) * pitch + (lane & 7) * 16;   // 8 instr fill 4 rows x 1 KB
        else p = base + (i >> 3) * 4096 + lane * 128 + (i & 7) * 16;
        if (CS) asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

int main() {
    const long total = 152L << 20;
    float4* out;
    cudaMalloc(&out, total * 2 + (64 << 20));
    float* flush;
    cudaMalloc(&flush, 256 << 20);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(store_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int cs = 0; cs < 2; ++cs)
        for (int pattern = 0; pattern < 3; ++pattern)
            for (int W = 4; W <= 32; W *= 2) {
                long bpw = total / (148L * W) / 4096 * 4096;
                float best = 1e9f;
                for (int it = 0; it < 5; ++it) {
                    cudaMemsetAsync(flush, 0, 256 << 20);
                    cudaEventRecord(a);
                    if (cs) store_kernel<1><<<148, W * 32, 200 * 1024>>>(out, bpw, pattern, 4608);
                    else store_kernel<0><<<148, W * 32, 200 * 1024>>>(out, bpw, pattern, 4608);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    if (ms < best) best = ms;
                }
                const double bytes = (double)bpw * 148 * W;
                printf("cs=%d pattern=%d warps/SM=%2d : %7.1f us  %6.2f TB/s  %5.1f B/clk/SM (1.95 GHz)\n", cs, pattern, W, best * 1e3,
                       bytes / best / 1e9, bytes / 148 / (best * 1e-3 * 1.95e9));
            }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
