// store-throughput microbenchmark for the MIXED-operand epilogue (gemm_tc.cu, destination kind 4): 16-bit slots, three parts per
// row.  pattern 0: what the 32-column chunk does -- a warp instruction = 4 rows x 64 contiguous bytes (8 lanes x st.v2);
// pattern 1: a 64-column chunk -- 4 rows x 128 contiguous bytes (8 lanes x st.v4).  Same bytes, one CTA of W warps per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/stores_mixed scripts/ubench/stores_mixed.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k(char* out, long rows_per_warp, int pattern, long pitch, long part) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    // a warp owns `rows_per_warp` rows of 64 columns (128 B per part) in a (rows, pitch) matrix with three parts `part` bytes apart
    const long row0 = ((long)blockIdx.x * W + warp) * rows_per_warp;
    const int rsub = lane >> 3, c = lane & 7;
    for (long r = 0; r < rows_per_warp; r += 4) {
        char* q = out + (row0 + r + rsub) * pitch;
        if (pattern == 0) {
            for (int half = 0; half < 2; ++half)          // two 32-column chunks (in the kernel they come from different warps / times)
                for (int p = 0; p < 3; ++p)
                    asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(q + p * part + half * 64 + c * 8), "r"(lane), "r"(c) : "memory");
        } else {
            for (int p = 0; p < 3; ++p)
                asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(q + p * part + c * 16), "r"(lane), "r"(c), "r"(lane), "r"(c) : "memory");
        }
    }
}

int main() {
    const long pitch = 3072L * 8, part = 3072L * 2;      // the heads' shared hidden operand: Kp = 3072 slots, 8*Kp bytes per row
    const long rows = 32896L * 48;                        // 48 column blocks of 64 -> as many "rows of 64 columns"
    char* out;
    cudaMalloc(&out, rows * 128 * 4 + (1 << 20));
    float* flush;
    cudaMalloc(&flush, 256 << 20);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int pattern = 0; pattern < 2; ++pattern)
        for (int W = 8; W <= 32; W *= 2) {
            // rows laid out so that consecutive "rows" of a warp are consecutive matrix rows of one 64-column block
            const long rpw = 32896L * 48 / (148L * W) / 4 * 4;
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaMemsetAsync(flush, 0, 256 << 20);
                cudaEventRecord(a);
                k<<<148, W * 32>>>(out, rpw, pattern, 512 /* dense rows of 4 parts x 128 B */, 128);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            const double bytes = (double)rpw * 148 * W * 384;
            printf("pattern=%d warps/SM=%2d : %7.1f us  %6.2f TB/s\n", pattern, W, best * 1e3, bytes / best / 1e9);
        }
    (void)pitch; (void)part;
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
