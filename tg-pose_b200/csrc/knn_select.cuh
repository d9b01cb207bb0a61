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
    // (cap: an upper bound the caller knows for the admission threshold, e.g. from a previous pass)
    __device__ __forceinline__ void admit(float dd, int base, int lane, float& th, int K, float cap = 3.402823466e+38f) {
        unsigned m = __ballot_sync(FULL, dd < th);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(FULL, dd, src);
            if (cd < th) {
                insert(cd, base + src, lane);
                th = fminf(thresh(K), cap);
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

// ------------------------------------------------------------------------------------------------------------
// Threshold selection (used by knn_xyz_sel_kernel and knn_tc2_kernel): instead of inserting candidates one by one,
//   1. one pass over the distances keeps 64 GROUP MINIMA per query (the candidates are partitioned into 64 groups);
//      the K-th smallest of them, T, bounds the K-th smallest distance from above, and only ~K + 4 candidates
//      (25 +- 2 at K = 21, 41 +- 4 at K = 31, measured on uniform clouds) lie at or below it;
//   2. a second pass appends every candidate with d <= T to a small shared-memory buffer;
//   3. each buffered candidate computes its own rank by counting the buffered (distance, index) keys below it.
// Exact whatever T is, provided K <= count <= capacity -- the caller checks that and otherwise takes a slow path.

// ascending bitonic sort of one value per lane
__device__ __forceinline__ float warp_sort32(float v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(FULL, v, j);
            const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
            v = keep_min ? fminf(v, o) : fmaxf(v, o);
        }
    return v;
}

// K-th smallest (1-based, K <= 32) of the 64 values {v0, v1} held two per lane
__device__ __forceinline__ float warp_kth_of_64(float v0, float v1, int K, int lane) {
    const float a = warp_sort32(v0, lane), b = warp_sort32(v1, lane);
    // take i values from a and K - i from b: the K-th smallest is min_i max(a[i-1], b[K-i-1])
    float pa = __shfl_up_sync(FULL, a, 1);
    if (lane == 0) pa = -CUDART_INF_F;
    float pb = __shfl_sync(FULL, b, (K - 1 - lane) & 31);
    if (lane == K) pb = -CUDART_INF_F;
    float v = lane <= K ? fmaxf(pa, pb) : CUDART_INF_F;
    const float a31 = __shfl_sync(FULL, a, 31);
    if (K == 32 && lane == 0) v = fminf(v, a31);          // i = 32: all of a
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// K-th smallest (1-based, K <= 128) of the 128 values held four per lane (element e = 32 r + lane): a full bitonic sort
// (28 compare-exchange steps, across lanes for strides < 32 and across the four registers above) -- ~250 instructions,
// once per query
__device__ __forceinline__ float warp_kth_of_128(float v0, float v1, float v2, float v3, int K, int lane) {
    float v[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dr = j >> 5;                      // partner register: r ^ dr, same lane
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if ((r & dr) == 0) {
                        const int e = 32 * r + lane;
                        const bool asc = (e & k) == 0;      // (k = 128: always ascending)
                        const float lo = fminf(v[r], v[r | dr]), hi = fmaxf(v[r], v[r | dr]);
                        v[r] = asc ? lo : hi;
                        v[r | dr] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int e = 32 * r + lane;
                    const float o = __shfl_xor_sync(FULL, v[r], j);
                    const bool keep_min = ((e & k) == 0) == ((e & j) == 0);
                    v[r] = keep_min ? fminf(v[r], o) : fmaxf(v[r], o);
                }
            }
        }
    }
    const int e = K - 1;
    float out = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float c = __shfl_sync(FULL, v[r], e & 31);
        if ((e >> 5) == r) out = c;
    }
    return out;
}

// monotone map float -> uint32 (a < b <=> key(a) < key(b) for non-NaN values)
__device__ __forceinline__ uint32_t ordered_key(float d) {
    const uint32_t u = __float_as_uint(d);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}

// buf[0..cnt) holds (ordered_key(d), index) pairs in any order; writes the indices of ranks 1..k (ascending by
// (distance, index), rank 0 dropped, gcn3d.py:22) to out64 / out32.  cnt <= 32 * SLOTS_MAX.
template <int SLOTS_MAX>
__device__ __forceinline__ void warp_rank_store(const uint2* buf, int cnt, int k, int lane, int64_t* out64, int32_t* out32) {
#pragma unroll
    for (int s = 0; s < SLOTS_MAX; ++s) {
        if (s * 32 >= cnt) break;
        const int a = s * 32 + lane;
        const bool valid = a < cnt;
        const uint2 me = buf[valid ? a : 0];
        const unsigned long long ka = ((unsigned long long)me.x << 32) | me.y;
        int rank = 0;
#pragma unroll 4
        for (int b = 0; b < cnt; ++b) {
            const uint2 o = buf[b];
            const unsigned long long kb = ((unsigned long long)o.x << 32) | o.y;
            rank += kb < ka;
        }
        if (valid && rank >= 1 && rank <= k) {
            if (out64) out64[rank - 1] = (int64_t)me.y;
            if (out32) out32[rank - 1] = (int32_t)me.y;
        }
    }
}

}  // namespace tgp
