// optim.cu -- the two calls that close a training step of the reference (trainer/RL_TDA.py:223-224):
//   torch.nn.utils.clip_grad_norm_(net1.parameters(), 5); optimizer.step()   with optimizer = Ranger
//   (tools/torch_utils/solver/ranger2020.py:44-235: RAdam + Lookahead + gradient centralisation).
// The reference walks the ~100 parameter tensors in Python, ~15 ATen launches each.  Here parameters, gradients, both
// moments and the Lookahead copy live in five flat fp32 arenas with identical offsets, described by one ROW TABLE
// (a row = dim-0 slice of a centralised tensor, or a <= 4096-element piece of a 1-D tensor), and a step is two
// HBM-bound passes:
//   tgp_ranger_reduce  reads the gradients once: per-row sums (the centralisation means) + the global sum of squares
//                      (the clip norm)                                                      4 B / element
//   tgp_ranger_update  g' = clip * (g - mean_row); moments; RAdam step; Lookahead          28 B / element (+8 on Lookahead steps)
// A warp owns a row, so the mean is a register; lanes stride the row with 128-bit accesses between a scalar head (up to
// the first 16-byte boundary -- the arenas share offsets and are all 256-byte aligned, so one split serves all five) and
// a scalar tail: rows of 1286 = 1024 + 256 + 6 columns, the heads' usual width, are never 16-byte aligned as a whole.
#include "common.cuh"

namespace tgp {

constexpr int OPT_THREADS = 256;
constexpr int OPT_WARPS = OPT_THREADS / 32;

// row = [head scalars | n4 float4 | tail scalars]
struct RowSplit { int head, n4, tail; };
__device__ __forceinline__ RowSplit split_row(const tgp_ranger_row& r) {
    RowSplit s;
    s.head = min(r.len, (int)((4 - (r.off & 3)) & 3));
    s.n4 = (r.len - s.head) >> 2;
    s.tail = r.len - s.head - 4 * s.n4;
    return s;
}

__global__ void __launch_bounds__(OPT_THREADS)
ranger_reduce_kernel(const float* __restrict__ grads, const tgp_ranger_row* __restrict__ rows, int n_rows,
                     const int* __restrict__ active, float* __restrict__ row_sum, double* __restrict__ sumsq) {
    __shared__ float red[OPT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * OPT_WARPS + warp;
    float s = 0.f, q = 0.f;
    if (r < n_rows) {
        const tgp_ranger_row row = rows[r];
        if (!active || active[row.tensor]) {
            const float* g = grads + row.off;
            const RowSplit sp = split_row(row);
            const float4* g4 = reinterpret_cast<const float4*>(g + sp.head);
#pragma unroll 4
            for (int i = lane; i < sp.n4; i += 32) {
                const float4 v = __ldg(g4 + i);
                s += (v.x + v.y) + (v.z + v.w);
                q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
            }
            // head (< 4) and tail (< 4) elements: lanes 0..2 and 4..6
            const int e = lane < sp.head ? lane : (lane >= 4 && lane - 4 < sp.tail ? sp.head + 4 * sp.n4 + lane - 4 : -1);
            if (e >= 0) {
                const float v = __ldg(g + e);
                s += v;
                q = fmaf(v, v, q);
            }
        }
        s = warp_sum(s);
        if (lane == 0) row_sum[r] = s;
    }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < OPT_WARPS; ++w) t += (double)red[w];
        if (t != 0.0) atomicAdd(sumsq, t);     // a few thousand fp64 adds on one address; order-dependence ~1e-16 relative
    }
}

struct ElemCoef {
    float clip, beta1, beta2, omb1, omb2, eps, wd, neg_step, la_alpha;
    int rectified, lookahead;
};

__device__ __forceinline__ void ranger_elem(const ElemCoef& c, float mean, float g, float& p, float& m, float& v, float& slow) {
    const float gc = (g - mean) * c.clip;                     // clip_grad_norm_ then centralized_gradient (ranger2020.py:31-41)
    v = fmaf(c.omb2 * gc, gc, v * c.beta2);                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)   :174
    m = fmaf(c.omb1, gc, m * c.beta1);                        // exp_avg.mul_(beta1).add_(grad, alpha = 1 - beta1)          :177
    float G = c.rectified ? m / (sqrtf(v) + c.eps) : m;       // :208-212
    if (c.wd != 0.f) {                                        // G_grad.add_(p, alpha = weight_decay)                       :214-215
        G = fmaf(c.wd, p, G);
        if (!c.rectified) m = G;                              // `G_grad = exp_avg` (:212) is an alias: the reference's in-place add
    }                                                         // also lands in exp_avg on un-rectified steps; kept for parity
    p = fmaf(c.neg_step, G, p);                               // p.add_(G_grad, alpha = -step_size * lr)                    :220
    if (c.lookahead) {                                        // slow += alpha (p - slow); p = slow                         :225-231
        slow = fmaf(c.la_alpha, p - slow, slow);
        p = slow;
    }
}

__global__ void __launch_bounds__(OPT_THREADS, 4)
ranger_update_kernel(float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ exp_avg,
                     float* __restrict__ exp_avg_sq, float* __restrict__ slow, const tgp_ranger_row* __restrict__ rows,
                     int row_begin, int row_end, const int* __restrict__ active, const float* __restrict__ row_sum,
                     const double* __restrict__ sumsq, tgp_ranger_hyper h, float* __restrict__ total_norm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = row_begin + blockIdx.x * OPT_WARPS + warp;
    const float norm = sumsq ? (float)sqrt(*sumsq) : 0.f;
    if (total_norm && blockIdx.x == 0 && threadIdx.x == 0) *total_norm = norm;
    if (r >= row_end) return;
    const tgp_ranger_row row = rows[r];
    if (active && !active[row.tensor]) return;                 // p.grad is None: the reference skips the tensor (:146-147)
    ElemCoef c;
    c.clip = (h.max_norm > 0.f && sumsq) ? fminf(1.f, h.max_norm / (norm + 1e-6f)) : 1.f;   // clip_grad_norm_: coef clamped to 1
    c.beta1 = h.beta1; c.beta2 = h.beta2; c.omb1 = h.one_minus_beta1; c.omb2 = h.one_minus_beta2;
    c.eps = h.eps; c.wd = h.weight_decay; c.neg_step = h.neg_step; c.la_alpha = h.la_alpha;
    c.rectified = h.rectified; c.lookahead = h.lookahead;
    const float mean = row.gc ? row_sum[r] / (float)row.len : 0.f;
    float* p = params + row.off;
    const float* g = grads + row.off;
    float* m = exp_avg + row.off;
    float* v = exp_avg_sq + row.off;
    float* s = slow + row.off;
    const RowSplit sp = split_row(row);
    {
        float4* pv = reinterpret_cast<float4*>(p + sp.head);
        float4* mv = reinterpret_cast<float4*>(m + sp.head);
        float4* vv = reinterpret_cast<float4*>(v + sp.head);
        float4* sv = reinterpret_cast<float4*>(s + sp.head);
        const float4* gv = reinterpret_cast<const float4*>(g + sp.head);
#pragma unroll 2
        for (int i = lane; i < sp.n4; i += 32) {
            const float4 g4 = __ldg(gv + i);
            float4 p4 = pv[i], m4 = mv[i], v4 = vv[i];
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c.lookahead) s4 = sv[i];
            ranger_elem(c, mean, g4.x, p4.x, m4.x, v4.x, s4.x);
            ranger_elem(c, mean, g4.y, p4.y, m4.y, v4.y, s4.y);
            ranger_elem(c, mean, g4.z, p4.z, m4.z, v4.z, s4.z);
            ranger_elem(c, mean, g4.w, p4.w, m4.w, v4.w, s4.w);
            pv[i] = p4;
            mv[i] = m4;
            vv[i] = v4;
            if (c.lookahead) sv[i] = s4;
        }
    }
    // head (< 4) and tail (< 4) elements: lanes 0..2 and 4..6
    const int i = lane < sp.head ? lane : (lane >= 4 && lane - 4 < sp.tail ? sp.head + 4 * sp.n4 + lane - 4 : -1);
    if (i >= 0) {
        float pi = p[i], mi = m[i], vi = v[i], si = c.lookahead ? s[i] : 0.f;
        ranger_elem(c, mean, __ldg(g + i), pi, mi, vi, si);
        p[i] = pi; m[i] = mi; v[i] = vi;
        if (c.lookahead) s[i] = si;
    }
}

// gc_loc = False (ranger2020.py:217-218): the UPDATE G_grad is centralised instead of the gradient.  The mean of a row of
// G only exists once the whole row's moments are updated, so the warp walks its row twice: pass 1 updates the moments,
// forms G (Adam quotient or exp_avg, plus weight decay) and sums it; pass 2 re-forms G from the moments just written
// (L1/L2-resident: a row is <= a few thousand elements), removes the mean and steps.  On un-rectified steps G_grad IS
// exp_avg in the reference (an alias, :212), so weight decay and the centralisation land in exp_avg as well; kept.
// Not the reference's default and not on the benchmark path: scalar accesses.
__global__ void __launch_bounds__(OPT_THREADS)
ranger_update_gcu_kernel(float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ exp_avg,
                         float* __restrict__ exp_avg_sq, float* __restrict__ slow, const tgp_ranger_row* __restrict__ rows,
                         int row_begin, int row_end, const int* __restrict__ active, const double* __restrict__ sumsq,
                         tgp_ranger_hyper h, float* __restrict__ total_norm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = row_begin + blockIdx.x * OPT_WARPS + warp;
    const float norm = sumsq ? (float)sqrt(*sumsq) : 0.f;
    if (total_norm && blockIdx.x == 0 && threadIdx.x == 0) *total_norm = norm;
    if (r >= row_end) return;
    const tgp_ranger_row row = rows[r];
    if (active && !active[row.tensor]) return;
    const float clip = (h.max_norm > 0.f && sumsq) ? fminf(1.f, h.max_norm / (norm + 1e-6f)) : 1.f;
    float* p = params + row.off;
    const float* g = grads + row.off;
    float* m = exp_avg + row.off;
    float* v = exp_avg_sq + row.off;
    float* s = slow + row.off;
    float sum = 0.f;
    for (int i = lane; i < row.len; i += 32) {
        const float gc = __ldg(g + i) * clip;
        const float vi = fmaf(h.one_minus_beta2 * gc, gc, v[i] * h.beta2);
        float mi = fmaf(h.one_minus_beta1, gc, m[i] * h.beta1);
        float G = h.rectified ? mi / (sqrtf(vi) + h.eps) : mi;
        if (h.weight_decay != 0.f) {
            G = fmaf(h.weight_decay, p[i], G);
            if (!h.rectified) mi = G;
        }
        v[i] = vi;
        m[i] = mi;
        sum += G;
    }
    sum = warp_sum(sum);
    const float mean = row.gc ? sum / (float)row.len : 0.f;
    __syncwarp();
    for (int i = lane; i < row.len; i += 32) {      // each lane re-reads only what it wrote itself
        const float mi = m[i];
        float pi = p[i];
        float G;
        if (h.rectified) {
            G = mi / (sqrtf(v[i]) + h.eps);
            if (h.weight_decay != 0.f) G = fmaf(h.weight_decay, pi, G);
            G -= mean;
        } else {
            G = mi - mean;                            // exp_avg already carries the weight-decay term (alias)
            m[i] = G;
        }
        pi = fmaf(h.neg_step, G, pi);
        if (h.lookahead) {
            const float si = fmaf(h.la_alpha, pi - s[i], s[i]);
            s[i] = si;
            pi = si;
        }
        p[i] = pi;
    }
}

}  // namespace tgp

using namespace tgp;

extern "C" int tgp_ranger_reduce(const float* grads, const tgp_ranger_row* rows_dev, int n_rows, const int* active_dev,
                                 float* row_sum, double* sumsq, tgp_stream_t stream) {
    if (!grads || !rows_dev || !row_sum || !sumsq) return fail(TGP_EINVAL, "tgp_ranger_reduce: null pointer");
    if (n_rows <= 0) return fail(TGP_EINVAL, "tgp_ranger_reduce: n_rows must be positive");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sumsq, 0, sizeof(double), st);
    if (e != cudaSuccess) return fail((int)e, "tgp_ranger_reduce: cudaMemsetAsync failed");
    const int blocks = (n_rows + OPT_WARPS - 1) / OPT_WARPS;
    ranger_reduce_kernel<<<blocks, OPT_THREADS, 0, st>>>(grads, rows_dev, n_rows, active_dev, row_sum, sumsq);
    return check_launch("ranger_reduce_kernel");
}

extern "C" int tgp_ranger_update(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* slow,
                                 const tgp_ranger_row* rows_dev, int row_begin, int row_end, const int* active_dev,
                                 const float* row_sum, const double* sumsq, const tgp_ranger_hyper* hyper_host,
                                 float* total_norm, tgp_stream_t stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !slow || !rows_dev || !row_sum || !hyper_host)
        return fail(TGP_EINVAL, "tgp_ranger_update: null pointer");
    if (row_begin < 0 || row_end <= row_begin) return fail(TGP_EINVAL, "tgp_ranger_update: empty row range");
    const tgp_ranger_hyper h = *hyper_host;
    if (!(h.beta1 >= 0.f && h.beta1 < 1.f && h.beta2 >= 0.f && h.beta2 < 1.f) || !(h.eps > 0.f))
        return fail(TGP_EINVAL, "tgp_ranger_update: betas must be in [0, 1) and eps > 0");
    if (h.lookahead && !(h.la_alpha >= 0.f && h.la_alpha <= 1.f))
        return fail(TGP_EINVAL, "tgp_ranger_update: Lookahead alpha must be in [0, 1]");   // ranger2020.py:81-82
    if (h.max_norm > 0.f && !sumsq) return fail(TGP_EINVAL, "tgp_ranger_update: clipping needs the sumsq of tgp_ranger_reduce");
    const int blocks = (row_end - row_begin + OPT_WARPS - 1) / OPT_WARPS;
    if (h.gc_on_update) {
        ranger_update_gcu_kernel<<<blocks, OPT_THREADS, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, slow,
                                                                               rows_dev, row_begin, row_end, active_dev,
                                                                               sumsq, h, total_norm);
        return check_launch("ranger_update_gcu_kernel");
    }
    ranger_update_kernel<<<blocks, OPT_THREADS, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, slow, rows_dev,
                                                                       row_begin, row_end, active_dev, row_sum, sumsq, h,
                                                                       total_norm);
    return check_launch("ranger_update_kernel");
}
