/*
 * oracle.c -- CPU restatement of TG-Pose's 3D-GCN + chamfer3D hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tg-pose_b200/ may import, link or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg can check (and time) the
 * algorithm on host cores.  Plain C, fp32 arithmetic in the reference's
 * operation order, OpenMP over (cloud, point).
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against
 * the .npz files under tests/golden/, which tests/golden/make_golden.py produced by importing
 * the unmodified reference (network/fs_net_repo/gcn3d.py, FaceRecon.py,
 * losses/metrics/CD/chamfer_python.py) on CPU.
 *
 * All "ref:" citations are relative to the reference checkout.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).  Contraction
 * is OFF so that every fmaf() below is deliberate and every a*b+c is two
 * roundings, which is what the bit-exact xyz kNN recipe needs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- ordered selection of the kk smallest (d, j) pairs, ties -> lower j ---- */
static inline void select_insert(float* bd, int* bi, int* cnt, int kk, float d, int j) {
    /* candidates arrive in increasing j, so on equal d the earlier one stays in front */
    int c = *cnt;
    if (c == kk && !(d < bd[kk - 1])) return;
    int p = (c < kk) ? c : kk - 1;
    while (p > 0 && d < bd[p - 1]) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
    }
    bd[p] = d;
    bi[p] = j;
    if (c < kk) *cnt = c + 1;
}

/*
 * xyz-space kNN.  ref: gcn3d.py:14-23 get_neighbor_index (D == 3).
 *   inner = bmm(x, x^T); q = sum(x**2, 2); d = inner*(-2) + q[None,:] + q[:,None]
 *   topk(k+1, smallest, sorted) then drop column 0 (positional, never "j == i").
 * Bit-exact recipe for CPU torch:
 *   q = (x0*x0 + x1*x1) + x2*x2, every op rounded;
 *   inner = fma(a2, b2, fma(a1, b1, a0*b0));
 *   d = ((inner * -2) + q_j) + q_i.
 * dist_out (optional, B*N*N) receives the full fp32 distance matrix.
 */
void orc_knn_xyz(const float* x, int B, int N, int k, int64_t* idx, float* dist_out) {
    const int kk = k + 1;
#pragma omp parallel
    {
        float* bd = (float*)malloc(sizeof(float) * kk);
        int* bi = (int*)malloc(sizeof(int) * kk);
        float* q = (float*)malloc(sizeof(float) * N);
        int cur_b = -1;
#pragma omp for schedule(static)
        for (long r = 0; r < (long)B * N; ++r) {
            int b = (int)(r / N), i = (int)(r % N);
            const float* xb = x + (size_t)b * N * 3;
            if (b != cur_b) {
                for (int j = 0; j < N; ++j) {
                    float a0 = xb[j * 3], a1 = xb[j * 3 + 1], a2 = xb[j * 3 + 2];
                    float s = a0 * a0;
                    s = s + a1 * a1;
                    s = s + a2 * a2;
                    q[j] = s;
                }
                cur_b = b;
            }
            float a0 = xb[i * 3], a1 = xb[i * 3 + 1], a2 = xb[i * 3 + 2];
            int cnt = 0;
            for (int j = 0; j < N; ++j) {
                float inner = fmaf(a2, xb[j * 3 + 2], fmaf(a1, xb[j * 3 + 1], a0 * xb[j * 3]));
                float d = inner * -2.0f;
                d = d + q[j];
                d = d + q[i];
                if (dist_out) dist_out[(size_t)r * N + j] = d;
                select_insert(bd, bi, &cnt, kk, d, j);
            }
            for (int t = 0; t < k; ++t) idx[(size_t)r * k + t] = bi[t + 1];
        }
        free(bd); free(bi); free(q);
    }
}

/*
 * feature-space kNN.  ref: gcn3d.py:14-23 with D = 128/256 (RF-F, gcn3d.py:201-206).
 * Same expanded fp32 formula; the channel sum of the inner product is accumulated
 * in double and rounded once (the centre of what any fp32 GEMM backend produces --
 * no backend order is reproducible, SURVEY 8c').  q is a plain fp32 running sum.
 * dist_out (optional, B*N*N) lets the comparator bound near-tie index swaps.
 */
void orc_knn_feat(const float* x, int B, int N, int D, int k, int64_t* idx, float* dist_out) {
    const int kk = k + 1;
#pragma omp parallel
    {
        float* bd = (float*)malloc(sizeof(float) * kk);
        int* bi = (int*)malloc(sizeof(int) * kk);
        float* q = (float*)malloc(sizeof(float) * N);
        int cur_b = -1;
#pragma omp for schedule(static)
        for (long r = 0; r < (long)B * N; ++r) {
            int b = (int)(r / N), i = (int)(r % N);
            const float* xb = x + (size_t)b * N * D;
            if (b != cur_b) {
                for (int j = 0; j < N; ++j) {
                    double s = 0.0;
                    for (int c = 0; c < D; ++c) s += (double)xb[(size_t)j * D + c] * xb[(size_t)j * D + c];
                    q[j] = (float)s;
                }
                cur_b = b;
            }
            const float* xi = xb + (size_t)i * D;
            int cnt = 0;
            for (int j = 0; j < N; ++j) {
                const float* xj = xb + (size_t)j * D;
                double acc = 0.0;
                for (int c = 0; c < D; ++c) acc += (double)xi[c] * (double)xj[c];
                float d = (float)acc * -2.0f;
                d = d + q[j];
                d = d + q[i];
                if (dist_out) dist_out[(size_t)r * N + j] = d;
                select_insert(bd, bi, &cnt, kk, d, j);
            }
            for (int t = 0; t < k; ++t) idx[(size_t)r * k + t] = bi[t + 1];
        }
        free(bd); free(bi); free(q);
    }
}

/*
 * nearest source point for every target point.  ref: gcn3d.py:26-35 get_nearest_index.
 *   d = (s_norm[j] + t_norm[i]) - 2*inner   (note: different op order from kNN)
 *   topk(k=1, smallest) -> lowest index on exact ties here.
 */
void orc_nearest(const float* tgt, const float* src, int B, int N, int M, int64_t* idx) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * N; ++r) {
        int b = (int)(r / N);
        const float* t = tgt + (size_t)r * 3;
        const float* sb = src + (size_t)b * M * 3;
        float tn = t[0] * t[0];
        tn = tn + t[1] * t[1];
        tn = tn + t[2] * t[2];
        float best = 0.f; int bj = 0;
        for (int j = 0; j < M; ++j) {
            const float* s = sb + j * 3;
            float sn = s[0] * s[0];
            sn = sn + s[1] * s[1];
            sn = sn + s[2] * s[2];
            float inner = fmaf(t[2], s[2], fmaf(t[1], s[1], t[0] * s[0]));
            float d = (sn + tn) - 2.0f * inner;
            if (j == 0 || d < best) { best = d; bj = j; }
        }
        idx[r] = bj;
    }
}

/* ref: gcn3d.py:38-46 indexing_neighbor_new -- out[b,m,j,:] = t[b, index[b,m,j], :] */
void orc_gather(const float* t, const int64_t* index, int B, int N, int M, int k, int C, float* out) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * M * k; ++r) {
        int b = (int)(r / ((long)M * k));
        memcpy(out + (size_t)r * C, t + ((size_t)b * N + index[r]) * C, sizeof(float) * C);
    }
}

/*
 * ref: gcn3d.py:48-58 get_neighbor_direction_norm
 *   v = xyz[idx] - xyz[n];  v / max(||v||_2, 1e-12)   (F.normalize, eps 1e-12)
 */
static inline void dir_norm(const float* c, const float* p, float* o) {
    float vx = p[0] - c[0], vy = p[1] - c[1], vz = p[2] - c[2];
    float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
    float den = nrm > 1e-12f ? nrm : 1e-12f;
    o[0] = vx / den; o[1] = vy / den; o[2] = vz / den;
}

void orc_dirnorm(const float* xyz, const int64_t* idx, int B, int N, int k, float* out) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * N; ++r) {
        int b = (int)(r / N);
        for (int j = 0; j < k; ++j)
            dir_norm(xyz + (size_t)r * 3, xyz + ((size_t)b * N + idx[(size_t)r * k + j]) * 3,
                     out + ((size_t)r * k + j) * 3);
    }
}

/* column-normalise directions (3, SC): ref gcn3d.py:99,165 F.normalize(directions, dim=0) */
static void normalize_dirs(const float* dirs, int SC, float* sd) {
    for (int c = 0; c < SC; ++c) {
        float a = dirs[c], b = dirs[SC + c], d = dirs[2 * SC + c];
        float nrm = sqrtf(a * a + b * b + d * d);
        float den = nrm > 1e-12f ? nrm : 1e-12f;
        sd[c] = a / den; sd[SC + c] = b / den; sd[2 * SC + c] = d / den;
    }
}

/*
 * surface graph conv.  ref: gcn3d.py:91-106 HSlayer_surface.graph_conv
 *   theta = relu(dirnorm @ normalize(directions, dim=0)) -> (B,N,k,S,C) support-major
 *   out = mean_s max_j theta
 */
void orc_surface_conv(const float* xyz, const int64_t* idx, const float* directions,
                      int B, int N, int k, int S, int C, float* out) {
    const int SC = S * C;
    float* sd = (float*)malloc(sizeof(float) * 3 * SC);
    normalize_dirs(directions, SC, sd);
#pragma omp parallel
    {
        float* mx = (float*)malloc(sizeof(float) * SC);
        float* dn = (float*)malloc(sizeof(float) * 3 * k);
#pragma omp for schedule(static)
        for (long r = 0; r < (long)B * N; ++r) {
            int b = (int)(r / N);
            for (int j = 0; j < k; ++j)
                dir_norm(xyz + (size_t)r * 3, xyz + ((size_t)b * N + idx[(size_t)r * k + j]) * 3, dn + j * 3);
            for (int c = 0; c < SC; ++c) {
                float m = -FLT_MAX;
                for (int j = 0; j < k; ++j) {
                    float th = dn[j * 3] * sd[c] + dn[j * 3 + 1] * sd[SC + c] + dn[j * 3 + 2] * sd[2 * SC + c];
                    th = th > 0.f ? th : 0.f;
                    if (th > m) m = th;
                }
                mx[c] = m;
            }
            for (int c = 0; c < C; ++c) {
                float s = 0.f;
                for (int t = 0; t < S; ++t) s += mx[t * C + c];
                out[(size_t)r * C + c] = s / (float)S;
            }
        }
        free(mx); free(dn);
    }
    free(sd);
}

/*
 * layer graph conv given the projected features P = fm @ weights + bias.
 * ref: gcn3d.py:157-180 HS_layer.graph_conv
 *   centre = P[..., :C]; support = P[..., C:] (support-major: col = s*C + c)
 *   out = centre + mean_s max_j relu(theta)[j,s,c] * support[idx_j, s, c]
 * dirs come from xyz, idx from feature space (gcn3d.py:201-207).
 */
void orc_layer_conv(const float* xyz, const int64_t* idx, const float* directions, const float* P,
                    int B, int N, int k, int S, int C, float* out) {
    const int SC = S * C, PC = (S + 1) * C;
    float* sd = (float*)malloc(sizeof(float) * 3 * SC);
    normalize_dirs(directions, SC, sd);
#pragma omp parallel
    {
        float* mx = (float*)malloc(sizeof(float) * SC);
        float* dn = (float*)malloc(sizeof(float) * 3 * k);
#pragma omp for schedule(static)
        for (long r = 0; r < (long)B * N; ++r) {
            int b = (int)(r / N);
            for (int j = 0; j < k; ++j)
                dir_norm(xyz + (size_t)r * 3, xyz + ((size_t)b * N + idx[(size_t)r * k + j]) * 3, dn + j * 3);
            for (int c = 0; c < SC; ++c) mx[c] = -FLT_MAX;
            for (int j = 0; j < k; ++j) {
                const float* sup = P + ((size_t)b * N + idx[(size_t)r * k + j]) * PC + C;
                for (int c = 0; c < SC; ++c) {
                    float th = dn[j * 3] * sd[c] + dn[j * 3 + 1] * sd[SC + c] + dn[j * 3 + 2] * sd[2 * SC + c];
                    th = th > 0.f ? th : 0.f;
                    float v = th * sup[c];
                    if (v > mx[c]) mx[c] = v;
                }
            }
            for (int c = 0; c < C; ++c) {
                float s = 0.f;
                for (int t = 0; t < S; ++t) s += mx[t * C + c];
                out[(size_t)r * C + c] = P[(size_t)r * PC + c] + s / (float)S;
            }
        }
        free(mx); free(dn);
    }
    free(sd);
}

/*
 * out = A (M,K) @ W (K,Nout) + bias (Nout, may be NULL).  ref: gcn3d.py:170
 * `feature_map @ self.weights + self.bias`, and every 1x1 Conv1d on the path
 * (gcn3d.py:70-71,130,132).  Accumulated in double, rounded once.
 */
void orc_gemm_bias(const float* A, const float* W, const float* bias, long M, int K, int Nout, float* out) {
#pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * Nout);
#pragma omp for schedule(static)
        for (long m = 0; m < M; ++m) {
            for (int n = 0; n < Nout; ++n) acc[n] = bias ? (double)bias[n] : 0.0;
            for (int kq = 0; kq < K; ++kq) {
                double a = A[(size_t)m * K + kq];
                const float* w = W + (size_t)kq * Nout;
                for (int n = 0; n < Nout; ++n) acc[n] += a * (double)w[n];
            }
            for (int n = 0; n < Nout; ++n) out[(size_t)m * Nout + n] = (float)acc[n];
        }
        free(acc);
    }
}

/*
 * out[b,m,c] = max_j f[b, idx[b,rows[m],j], c] for the selected rows (rows == NULL -> all N).
 * ref: gcn3d.py:214-215 (ORL gather + max), :236-239,:244 (Pool: max then row subset).
 * arg (optional) receives the winning neighbour slot j (first maximum).
 */
void orc_gather_max(const float* f, const int64_t* idx, const int64_t* rows, int B, int N, int M, int k, int C,
                    float* out, uint8_t* arg) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * M; ++r) {
        int b = (int)(r / M), m = (int)(r % M);
        long n = rows ? rows[m] : m;
        const int64_t* id = idx + ((size_t)b * N + n) * k;
        for (int c = 0; c < C; ++c) {
            float best = -FLT_MAX; int bj = 0;
            for (int j = 0; j < k; ++j) {
                float v = f[((size_t)b * N + id[j]) * C + c];
                if (v > best) { best = v; bj = j; }
            }
            out[(size_t)r * C + c] = best;
            if (arg) arg[(size_t)r * C + c] = (uint8_t)bj;
        }
    }
}

/* ref: gcn3d.py:210-217 get_ORL_global -- g[b,c] = mean_n max_j f[b, idx[b,n,j], c] (before .repeat) */
void orc_orl_global(const float* f, const int64_t* idx, int B, int N, int k, int C, float* g) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        double* acc = (double*)calloc(C, sizeof(double));
        for (int n = 0; n < N; ++n) {
            const int64_t* id = idx + ((size_t)b * N + n) * k;
            for (int c = 0; c < C; ++c) {
                float best = -FLT_MAX;
                for (int j = 0; j < k; ++j) {
                    float v = f[((size_t)b * N + id[j]) * C + c];
                    if (v > best) best = v;
                }
                acc[c] += best;
            }
        }
        for (int c = 0; c < C; ++c) g[(size_t)b * C + c] = (float)(acc[c] / N);
        free(acc);
    }
}

/*
 * chamfer forward, one direction.  ref: losses/chamfer3D/chamfer3D.cu:12-134 NmDistanceKernel
 * and tools/pyTorchChamferDistance/chamfer_distance.cpp:59-87 nnsearch:
 *   d = dx*dx + dy*dy + dz*dz with dx = x2 - x1 (exact difference form), strict '<'
 *   so the lowest index wins ties.
 * contract != 0 reproduces what nvcc 12.9 (default -fmad=true) makes of the CUDA source's
 * `x2*x2+y2*y2+z2*z2` for sm_100a -- fma(dz,dz, fma(dx,dx, dy*dy)): FMUL on y, FFMA on x, FFMA on z, read off the
 * SASS of oracle/_ref/chamfer3D (built from the reference's own chamfer3D.cu by oracle/build_ref.py) and checked
 * bit for bit against that kernel on the GPU (tests/test_gpu_ref_chamfer.py);
 * contract == 0 is the plain C++ rounding of nnsearch.
 */
void orc_chamfer_nn(const float* a, const float* b2, int B, int n, int m, int contract, float* dist, int32_t* idx) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * n; ++r) {
        int b = (int)(r / n);
        const float* p = a + (size_t)r * 3;
        const float* qb = b2 + (size_t)b * m * 3;
        float best = 0.f; int bj = 0;
        for (int j = 0; j < m; ++j) {
            float dx = qb[j * 3] - p[0], dy = qb[j * 3 + 1] - p[1], dz = qb[j * 3 + 2] - p[2];
            float d;
            if (contract) d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            else { d = dx * dx; d = d + dy * dy; d = d + dz * dz; }
            if (j == 0 || d < best) { best = d; bj = j; }
        }
        dist[r] = best;
        idx[r] = bj;
    }
}

/*
 * chamfer backward.  ref: chamfer3D.cu:155-174 NmDistanceGradKernel (x2 launches, :184-185)
 * and chamfer_distance.cpp:114-177.  g1/g2 must be zeroed by the caller
 * (dist_chamfer_3D.py:56-60).  Serial so the fp32 accumulation order is fixed.
 */
void orc_chamfer_bwd(const float* x1, const float* x2, const float* gd1, const float* gd2,
                     const int32_t* i1, const int32_t* i2, int B, int n, int m, float* g1, float* g2) {
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < n; ++j) {
            const float* p = x1 + ((size_t)b * n + j) * 3;
            int j2 = i1[(size_t)b * n + j];
            const float* q = x2 + ((size_t)b * m + j2) * 3;
            float g = gd1[(size_t)b * n + j] * 2.f;
            for (int a = 0; a < 3; ++a) {
                float v = g * (p[a] - q[a]);
                g1[((size_t)b * n + j) * 3 + a] += v;
                g2[((size_t)b * m + j2) * 3 + a] -= v;
            }
        }
        for (int j = 0; j < m; ++j) {
            const float* p = x2 + ((size_t)b * m + j) * 3;
            int j2 = i2[(size_t)b * m + j];
            const float* q = x1 + ((size_t)b * n + j2) * 3;
            float g = gd2[(size_t)b * m + j] * 2.f;
            for (int a = 0; a < 3; ++a) {
                float v = g * (p[a] - q[a]);
                g2[((size_t)b * m + j) * 3 + a] += v;
                g1[((size_t)b * n + j2) * 3 + a] -= v;
            }
        }
    }
}
