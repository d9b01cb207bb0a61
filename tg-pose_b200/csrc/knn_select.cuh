// knn_select.cuh -- warp-level top-k selection shared by the kNN kernels (knn.cu, knn_tc.cu).
//
// Ordering contract: ascending by (distance, index); candidates are offered in increasing index
// order, so equal distances queue behind earlier ones (lowest index first); rank 0 is dropped positionally
// by the caller (gcn3d.py:22).
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace tgp {

constexpr unsigned FULL = 0xffffffffu;

// Sorted list of 32*SLOTS (distance, index) pairs spread over a warp: rank = lane + 32*s.
template <int SLOTS>
struct WarpTopList {
    float d[SLOTS];
    int i[SLOTS];

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) { d[s] = CUDART_INF_F; i[s] = -1; }
    }
    // distance currently at rank K-1 (the admission threshold)
    __device__ __forceinline__ float thresh(int K) const {
        const int r = K - 1;
        float v = 0.f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
            if ((r >> 5) == s) v = __shfl_sync(FULL, d[s], r & 31);
        return v;
    }
    // insert (cd, cj); candidates arrive in increasing cj, so equal distances go behind
    __device__ __forceinline__ void insert(float cd, int cj, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) pos += __popc(__ballot_sync(FULL, d[s] <= cd));
#pragma unroll
        for (int s = SLOTS - 1; s >= 0; --s) {
            float up = __shfl_up_sync(FULL, d[s], 1);
            int upi = __shfl_up_sync(FULL, i[s], 1);
            if (s > 0) {
                float cr = __shfl_sync(FULL, d[s - 1], 31);
                int cri = __shfl_sync(FULL, i[s - 1], 31);
                if (lane == 0) { up = cr; upi = cri; }
            }
            const int rank = lane + 32 * s;
            if (rank > pos) { d[s] = up; i[s] = upi; }
            else if (rank == pos) { d[s] = cd; i[s] = cj; }
        }
    }
    // admit every lane's candidate (dd, base+lane) that beats the threshold, lowest lane first
    __device__ __forceinline__ void admit(float dd, int base, int lane, float& th, int K) {
        unsigned m = __ballot_sync(FULL, dd < th);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(FULL, dd, src);
            if (cd < th) {
                insert(cd, base + src, lane);
                th = thresh(K);
            }
        }
    }
    template <typename T>
    __device__ __forceinline__ void store_ranks(T* out, int k, int lane) const {
        // ranks 1..k -> out[0..k)
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int rank = lane + 32 * s;
            if (rank >= 1 && rank <= k) out[rank - 1] = (T)i[s];
        }
    }
};


// Two independent 32-entry lists advanced in lock step: the (shuffle-latency-bound) insertion chains of the two
// queries are interleaved in one basic block, which doubles the instruction-level parallelism of the selection.
// All control flow is warp-uniform; an exhausted side keeps executing with pos = 33 (no lane changes).
struct WarpTopPair {
    float dA, dB;
    int iA, iB;
    float thA, thB;

    __device__ __forceinline__ void admit2(float ddA, float ddB, int base, int lane, int K) {
        unsigned mA = __ballot_sync(FULL, ddA < thA), mB = __ballot_sync(FULL, ddB < thB);
        while (mA | mB) {
            const int sA = mA ? __ffs(mA) - 1 : 0, sB = mB ? __ffs(mB) - 1 : 0;
            const float cdA = __shfl_sync(FULL, ddA, sA), cdB = __shfl_sync(FULL, ddB, sB);
            const bool vA = mA != 0 && cdA < thA, vB = mB != 0 && cdB < thB;
            mA &= mA - 1;
            mB &= mB - 1;
            const unsigned leA = __ballot_sync(FULL, dA <= cdA), leB = __ballot_sync(FULL, dB <= cdB);
            const int posA = vA ? __popc(leA) : 33, posB = vB ? __popc(leB) : 33;
            const float upA = __shfl_up_sync(FULL, dA, 1), upB = __shfl_up_sync(FULL, dB, 1);
            const int upiA = __shfl_up_sync(FULL, iA, 1), upiB = __shfl_up_sync(FULL, iB, 1);
            if (lane > posA) { dA = upA; iA = upiA; } else if (lane == posA) { dA = cdA; iA = base + sA; }
            if (lane > posB) { dB = upB; iB = upiB; } else if (lane == posB) { dB = cdB; iB = base + sB; }
            thA = __shfl_sync(FULL, dA, K - 1);
            thB = __shfl_sync(FULL, dB, K - 1);
        }
    }
};

}  // namespace tgp
