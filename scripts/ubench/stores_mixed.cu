// store-throughput microbenchmark for the MIXED-operand epilogue (gemm_tc.cu, destination kind 4): 16-bit slots, three parts per
// row, written into the heads' shared hidden operand (32896 rows x 3072 columns: 24576 bytes per row, parts 6144 bytes apart).
// A warp owns (a block of rows, one 64-column block = 128 bytes per part and row); one CTA of W warps per SM.
//   pattern 0: a warp instruction = 4 rows x 64 contiguous bytes (8 lanes x st.v2), both halves of a line from the SAME warp,
//              back to back
//   pattern 1: 4 rows x 128 contiguous bytes (8 lanes x st.v4): what a 64-column chunk would do
//   pattern 2: what the kernel does -- the two 64-byte halves of a line come from two DIFFERENT warps running side by side
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/stores_mixed scripts/ubench/stores_mixed.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr long ROWS = 32896, CB = 48, PITCH = 24576, PART = 6144;

__global__ void __launch_bounds__(1024, 1) k(char* out, int pattern, int rows_per_item) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int rsub = lane >> 3, c = lane & 7;
    const long nwarps = (long)gridDim.x * W;
    const long items = (ROWS / rows_per_item) * CB;            // (row block, column block)
    // pattern 2: warps pair up on an item, each writing one half of every line
    const long gw = (long)blockIdx.x * W + warp;
    const long first = pattern == 2 ? gw / 2 : gw, stride = pattern == 2 ? nwarps / 2 : nwarps;
    for (long it = first; it < items; it += stride) {
        const long rb = it / CB, cb = it % CB;
        for (int r = 0; r < rows_per_item; r += 4) {
            char* q = out + (rb * rows_per_item + r + rsub) * PITCH + cb * 128;
            if (pattern == 0) {
                for (int half = 0; half < 2; ++half)
                    for (int p = 0; p < 3; ++p)
                        asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(q + p * PART + half * 64 + c * 8), "r"(lane), "r"(c) : "memory");
            } else if (pattern == 1) {
                for (int p = 0; p < 3; ++p)
                    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(q + p * PART + c * 16), "r"(lane), "r"(c), "r"(lane), "r"(c) : "memory");
            } else {
                for (int p = 0; p < 3; ++p)
                    asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(q + p * PART + (warp & 1) * 64 + c * 8), "r"(lane), "r"(c) : "memory");
            }
        }
    }
}

int main() {
    char* out;
    cudaMalloc(&out, ROWS * PITCH);
    float* flush;
    cudaMalloc(&flush, 256 << 20);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int pattern = 0; pattern < 3; ++pattern)
        for (int W = 8; W <= 32; W *= 2) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaMemsetAsync(flush, 0, 256 << 20);
                cudaEventRecord(a);
                k<<<148, W * 32>>>(out, pattern, 32);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            const double bytes = (double)ROWS * CB * 384;
            printf("pattern=%d warps/SM=%2d : %7.1f us  %6.2f TB/s\n", pattern, W, best * 1e3, bytes / best / 1e9);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
