/*
 * tgpose_b200.h -- C-ABI of libtgpose_b200.so: the sm_100a implementation of TG-Pose's
 * 3D-GCN + chamfer3D hot path.
 *
 * Conventions (SURVEY 8b "What a C-ABI replacement must export"):
 *   - plain pointers + int sizes, no torch / ATen types; every pointer is DEVICE memory
 *     unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (PyTorch's current stream);
 *   - the callee never allocates, never synchronises, keeps no references, is re-entrant;
 *   - returns 0 on success, a negative TGP_E* for an argument error, or a positive
 *     cudaError_t from the launch; tgp_last_error() gives a thread-local message.
 *     (The reference's chamfer extension returns 1/0 and printf()s -- chamfer_cuda.cpp:17-33,
 *     chamfer3D.cu:145-151; the Python shim chamfer_3D.forward() maps 0 -> 1 to keep that contract.)
 *   - all float tensors are fp32, C-contiguous; shapes are given in the comments;
 *   - index tensors: `idx_bits` selects int64 (what torch.topk hands the reference's callers,
 *     gcn3d.py:21) or int32 (what the fused encoder keeps internally).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout, network/fs_net_repo/ unless noted).
 */
#ifndef TGPOSE_B200_H
#define TGPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGP_OK 0
#define TGP_EINVAL (-1)   /* bad size / null pointer / unsupported combination */
#define TGP_ENOSPACE (-2) /* workspace too small */

typedef void* tgp_stream_t;

/* library / build info */
int tgp_version(void);
const char* tgp_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py "gpu_launches") */
unsigned long long tgp_launch_count(void);

/* ------------------------------------------------------------------ kNN (gcn3d.py:14-35) */

/* get_neighbor_index(vertices (B,N,3), k), gcn3d.py:14-23, xyz space.
 * Bit-exact distance recipe, top-(k+1) by (distance, index) ascending, rank 0 dropped
 * positionally.  idx64 (B,N,k) int64 and/or idx32 (B,N,k) int32; either may be NULL. */
int tgp_knn_xyz(const float* xyz, int B, int N, int k, int64_t* idx64, int32_t* idx32, tgp_stream_t stream);

/* get_neighbor_index(feature_map (B,N,D), k), gcn3d.py:14-23 as used by 'RF-F' (:201-206).
 * The pairwise inner products (the reference's torch.bmm, gcn3d.py:18) run on the tensor cores (tcgen05, 3xTF32)
 * fused with the warp-shuffle top-(k+1) selection; shapes the tensor-core kernel does not cover (k > 31,
 * N > 4096, tiny problems) run on an fp32 FMA tile kernel with the same selection.
 * x_split, optional: x already split as a tensor-core operand (B*N, 2*Kp) by tgp_split_tf32 / a mode-2 epilogue;
 * workspace: tgp_knn_feat_workspace(B,N,D, x_split != NULL) bytes (row norms + the split when not supplied). */
size_t tgp_knn_feat_workspace(int B, int N, int D, int have_split);
int tgp_knn_feat(const float* x, const float* x_split, int B, int N, int D, int k, int64_t* idx64, int32_t* idx32,
                 void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* get_nearest_index(target (B,N,3), source (B,M,3)), gcn3d.py:26-35 -> (B,N,1). */
int tgp_nearest(const float* target, const float* source, int B, int N, int M,
                int64_t* idx64, int32_t* idx32, tgp_stream_t stream);

/* ------------------------------------------------------------------ gathers (gcn3d.py:38-58) */

/* indexing_neighbor_new(tensor (B,N,C), index (B,M,k)), gcn3d.py:38-46 -> out (B,M,k,C). */
int tgp_gather_rows(const float* tensor, const void* index, int idx_bits, int B, int N, int M, int k, int C,
                    float* out, tgp_stream_t stream);

/* t[:, rows, :] with one row list shared by the whole batch (Pool_layer's vertices[:, sample_idx, :],
 * gcn3d.py:243): t (B,N,C), rows (M) int64 -> out (B,M,C). */
int tgp_select_rows(const float* t, const int64_t* rows, int B, int N, int M, int C, float* out, tgp_stream_t stream);

/* get_neighbor_direction_norm(vertices (B,N,3), idx (B,N,k)), gcn3d.py:48-58 -> out (B,N,k,3). */
int tgp_direction_norm(const float* xyz, const void* idx, int idx_bits, int B, int N, int k,
                       float* out, tgp_stream_t stream);

/* max over the k gathered rows: out[b,m,c] = max_j f[b, idx[b,rows[m],j], c].
 * rows (M) int64 device pointer selects points (Pool_layer, gcn3d.py:236-244); NULL -> all N (M == N).
 * arg (B,M,C) uint8, optional: winning neighbour slot (saved for backward). */
int tgp_gather_max(const float* f, const void* idx, int idx_bits, const int64_t* rows, int B, int N, int M, int k,
                   int C, float* out, uint8_t* arg, tgp_stream_t stream);

/* get_ORL_global(feature (B,N,C), vertices, k), gcn3d.py:210-217, given the xyz kNN:
 * g[b,c] = mean_n max_j f[b, idx[b,n,j], c]   (the (B,1,C) tensor before .repeat).
 * arg (B,N,C) uint8 optional.  workspace: tgp_orl_workspace(B,N,C) bytes of scratch for the per-chunk
 * partial sums; they are reduced in a fixed order, so the result is deterministic and does not depend
 * on which batch a cloud sits in. */
size_t tgp_orl_workspace(int B, int N, int C);
int tgp_orl_global(const float* f, const void* idx, int idx_bits, int B, int N, int k, int C,
                   float* g, uint8_t* arg, void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* One source of tgp_concat_rows: (rows, C) fp32 with row stride ld.
 * idx == NULL, n_src > 0 : row r of the output takes row r of the source (n_src is ignored);
 * idx != NULL            : int32 (B,N): output row (b,n) takes source row b*n_src + idx[b,n]
 *                          (the nearest upsampling indexing_neighbor_new(fm, nearest).squeeze(2), FaceRecon.py:71-73);
 * n_src == 0             : one source row per cloud, broadcast to its N points (the one-hot category, FaceRecon.py:75-79). */
typedef struct {
    const float* ptr;
    int C;
    long ld;
    const int32_t* idx;
    int n_src;
} tgp_concat_src;

/* torch.cat([...], dim=2) of FaceRecon.py:81 (+ the cat with the points, PoseNet9D.py:63) fused with the gathers that
 * feed it: out[(b,n), :] = [src_0 row | src_1 row | ...].  out_raw (B*N, ld_raw) and/or out_split (B*N, 2*Kp): the
 * same row as a tensor-core operand [tf32 | residual] (mixed = 0) or as a MIXED operand (mixed = 1, see
 * tgp_gemm_args.mixed), zero padded to Kp.  Either may be NULL. */
int tgp_concat_rows(const tgp_concat_src* srcs_host, int nsrc, int B, int N, float* out_raw, long ld_raw,
                    float* out_split, int Kp, int mixed, tgp_stream_t stream);

/* ------------------------------------------------------------------ graph convolutions */

/* HSlayer_surface.graph_conv, gcn3d.py:91-106:
 * out[b,n,c] = mean_s max_j relu(<dirnorm(b,n,j), normalize(directions,dim=0)[:, s*C+c]>).
 * xyz (B,N,3), idx (B,N,k), directions (3,S*C) raw parameter, out (B,N,C),
 * arg (B,N,S*C) uint8 optional (arg-max neighbour slot per (n,s,c)).
 * out_split (B*N, 2*Kp) optional, Kp = tgp_split_kpad(C): the same result as a tensor-core operand
 * [tf32(v) | v - tf32(v)] for the contraction that follows (padding columns must be pre-zeroed). */
int tgp_surface_conv_fwd(const float* xyz, const void* idx, int idx_bits, const float* directions,
                         int B, int N, int k, int S, int C, float* out, uint8_t* arg, float* out_split,
                         tgp_stream_t stream);

/* Edge records for the layer convolution: rec[b,j,n] = (dx,dy,dz, bitcast<float>(int idx[b,n,j]))
 * with (dx,dy,dz) = get_neighbor_direction_norm (gcn3d.py:48-58).  rec: (B,k,N,4) fp32 -- NEIGHBOUR-major, so that
 * consecutive points read consecutive records -- 16-byte aligned. */
int tgp_edge_records(const float* xyz, const void* idx, int idx_bits, int B, int N, int k,
                     float* rec, tgp_stream_t stream);

/* HS_layer.graph_conv after the projection, gcn3d.py:157-180:
 * out[b,n,c] = centre[b,n,c] + mean_s max_j relu(theta[b,n,j,s,c]) * support[b, idx[b,n,j], s, c].
 * edge_rec from tgp_edge_records (directions from xyz, indices from feature space, gcn3d.py:201-207);
 * centre (B*N, C) row-major with leading dimension ld_centre;
 * support in SLAB layout [C/4][B*N][S][4] as written by tgp_gemm (mode 1), C % 4 == 0, S*4 <= 32;
 * arg_slab [C/4][B*N][S*4] uint8 optional: arg-max neighbour slot per (n,s,c), same layout;
 * out_split (B*N, 2*Kp) optional: as in tgp_surface_conv_fwd. */
int tgp_layer_conv_fwd(const float* edge_rec, const float* directions,
                       const float* centre, long ld_centre, const float* support_slab,
                       int B, int N, int k, int S, int C, float* out, uint8_t* arg_slab, float* out_split,
                       tgp_stream_t stream);

/* ------------------------------------------------------------------ dense contraction */

/* One output column range of tgp_gemm (ranges may overlap: every matching segment is written).
 * mode 0: row-major, out[m*ld + (col - col_begin)].
 * mode 1: SLAB, columns are ordered (cgroup, s, c4) and go to [cgroup][m][S*4] (slab_width = S*4).
 * mode 2: SPLIT, the value is written as a tensor-core operand for the next contraction:
 *         tf32(v) at out[m*ld + rel] and v - tf32(v) at out[m*ld + slab_width + rel] (slab_width = Kp).
 * mode 4: MIXED, the value is written as a mixed tensor-core operand (see tgp_gemm_args.mixed): ptr = operand base,
 *         ld = 2*Kp (fp32 slots per row), slab_width = Kp = tgp_mixed_kpad(width); col_begin maps to operand column 0.
 * mode 5: as mode 4 without the bf16(x) slot (left unwritten): the operand of a mixed = 2 contraction, which never reads it --
 *         one conversion and one store per four values less.  Modes 4 and 5 cannot be combined in one launch.
 * mode 3: COLUMN MAX per group of rows_per_group rows (torch.max over the points of a cloud, PoseR.py:33,
 *         FaceRecon.py:146): ptr is an int32 (M / rows_per_group, col_end - col_begin) buffer pre-filled with
 *         INT_MIN that receives atomicMax of the order-preserving encoding e(v) = bits(v) >= 0 ? bits(v) :
 *         bits(v) ^ 0x7fffffff (decode with the same map); the (M, cols) tensor itself is never written. */
typedef struct {
    int col_begin, col_end;
    int mode;
    int slab_width; /* S*4 for mode 1 */
    long ld;        /* mode 0 leading dimension */
    float* ptr;
} tgp_out_seg;

typedef struct {
    /* C = A (M,K; lda) * Bmat + epilogue.  b_is_nk: Bmat given as (Ncols,K) K-contiguous
     * (a Conv1d weight, gcn3d.py:70-71,130,132) instead of (K,Ncols) (HS_layer.weights, gcn3d.py:125). */
    const float* A; long lda;
    const float* Bmat; long ldb; int b_is_nk;
    long M; int K; int Ncols;
    const float* bias;            /* (Ncols) or NULL */
    const float* group_bias;      /* (M / rows_per_group, Ncols) or NULL: per-cloud bias (ORL split) */
    int rows_per_group;
    const float* res1; long ld_res1;  /* optional residuals, (M, Ncols) */
    const float* res2; long ld_res2;
    const float* scale; const float* shift; /* optional per-column affine (eval BatchNorm) applied last */
    int relu;                     /* 1: max(v, 0) after the affine */
    const float* neg_slope;       /* optional per-column leaky slope: v > 0 ? v : v*neg_slope[col] (overrides relu) */
    int nseg; tgp_out_seg seg[4];
    /* tensor-core path (tcgen05, 3xTF32): both operands pre-split by tgp_split_tf32 into
     * [tf32(x) | x - tf32(x)], (rows, 2*Kp) with Kp = tgp_split_kpad(K).  A_split: (M, 2Kp),
     * B_split: (Ncols, 2Kp).  When both are non-NULL they are used instead of A / Bmat;
     * when NULL the exact-fp32 FMA path runs on A / Bmat. */
    const float* A_split;
    const float* B_split;
    /* mixed = 1: A_split / B_split are MIXED operands (tgp_split_mixed, output mode 4): rows of 8*Kp bytes,
     * Kp = tgp_mixed_kpad(K), holding 16-bit slots [fp16(x) x Kp | bf16(x) x Kp | bf16(x - fp16(x)) x Kp | unused x Kp]
     * (fp16 saturating at +-65504; the residual then carries the remainder).  The product runs as
     * fp16(a).fp16(b) + lo(a).bf16(b) + bf16(a).lo(b), three 16-bit tensor-core passes at twice the TF32 rate into one
     * fp32 accumulator = 1.5 TF32-pass equivalents instead of the three of 3xTF32, relative error ~2^-19 per product
     * (fp32-summation-noise level at the heads' K ~ 1e3).  Used for the heads' 1x1 convolutions (PoseR.py, PoseTs.py,
     * FaceRecon.py:89-167), whose outputs feed no neighbour search; the encoder and the kNN keep the 3xTF32 operands.
     * mixed = 2: B_split is a tgp_split_mixed_w16 operand (residual slot in fp16) and the third pass is fp16(a).lo16(b): the
     * bf16(x) slot of A_split is never read, so producers may leave it out (output mode 5).  Same ~2^-19 per product. */
    int mixed;
    /* block-diagonal ("grouped") contraction on mixed operands: when a_group_cols > 0, output columns
     * [g*a_group_cols, (g+1)*a_group_cols) contract the A columns [g*Kp, (g+1)*Kp) of a WIDER operand of padded width a_kp
     * (several equally shaped layers that read different column blocks of one activation buffer and would each fill the
     * machine badly -- the three pose-head conv2 layers: 257 row tiles on 148 SMs -- run as one launch).  a_group_cols must be
     * a multiple of 256, K a multiple of 64; B_split stacks the layers' (a_group_cols, K) weights.  0: plain contraction. */
    int a_kp;
    int a_group_cols;
    /* GATHERED residuals: when res1_idx / res2_idx is non-NULL (int32, M entries), output row m adds row
     * res1_idx[m] / res2_idx[m] of res1 / res2 instead of row m (both the fp32 FMA and the tensor-core path; the latter needs
     * every width / leading dimension a multiple of 4 floats and 16-byte aligned pointers, else TGP_EINVAL).
     * A 1x1 convolution over a concatenation whose column blocks are nearest-neighbour upsamplings of coarser levels (FaceRecon.py:69-81 feeding PoseR.py:26, PoseTs.py:24,
     * FaceRecon.py:100,133) is W.[x | up(y)] = W_x.x + up(W_y.y): the W_y.y products are computed once per COARSE point
     * and gathered here, instead of once per fine point. */
    const int32_t* res1_idx;
    const int32_t* res2_idx;
} tgp_gemm_args;

/* feature_map @ weights + bias (gcn3d.py:170) and every 1x1 Conv1d on the path. */
int tgp_gemm(const tgp_gemm_args* args_host, tgp_stream_t stream);

/* decode the int32 cells of a mode-3 (column max) segment back to fp32: out[i] = float(e^-1(enc[i])). */
int tgp_decode_max(const int32_t* enc, long n, float* out, tgp_stream_t stream);

/* K rounded up to the tensor-core K block (32). */
int tgp_split_kpad(int K);
/* dst (rows, 2*Kp) = [tf32(src) | src - tf32(src)], zero padded.  src is (rows, K) with row stride ld,
 * or, if src_is_kn, (K, rows) with row stride ld (transposed on the fly: HS_layer.weights, gcn3d.py:125). */
int tgp_split_tf32(const float* src, long rows, int K, long ld, int src_is_kn, float* dst, tgp_stream_t stream);

/* mixed operand (tgp_gemm_args.mixed): K rounded up to 64, and the split of a row-major (rows, K) matrix (zero padded). */
int tgp_mixed_kpad(int K);
int tgp_split_mixed(const float* src, long rows, int K, long ld, float* dst, tgp_stream_t stream);
/* the WEIGHT operand of a tgp_gemm_args.mixed = 2 contraction: same layout, the residual slot in fp16:
 * [fp16(w) | bf16(w) | fp16(w - fp16(w)) | unused] (the weights of the heads' 1x1 convolutions, PoseR.py / PoseTs.py /
 * FaceRecon.py:89-167, folded and packed once). */
int tgp_split_mixed_w16(const float* src, long rows, int K, long ld, float* dst, tgp_stream_t stream);
/* transposed mixed operand, K-blocked: dst = tgp_split_mixed_t_bytes(rows, K) bytes (128-byte aligned), 16-bit slots
 * [part 0..2][ceil(rows/64)][ceil256(K)][64]: per block of 64 source rows a dense (ceil256(K) x 128 B) matrix for each of
 * fp16(x), bf16(x), bf16(x - fp16(x)); source rows past `rows` are zero.  Operand of tgp_gemm_tn_tc(mixed = 1). */
size_t tgp_split_mixed_t_bytes(long rows, int K);
/* the same for the 3xTF32 operands of the encoder's weight gradients: src (Kdim, rows) row-major (Kdim = contraction
 * length), dst = tgp_split_tf32_t_bytes(Kdim, rows) bytes, floats [part hi|lo][ceil(Kdim/32)][ceil256(rows)][32].
 * Operand of tgp_gemm_tn_tc(mixed = 0). */
size_t tgp_split_tf32_t_bytes(long Kdim, int rows);
int tgp_split_tf32_t(const float* src, long Kdim, int rows, long ld, float* dst, tgp_stream_t stream);
int tgp_split_mixed_t(const float* src, long rows, int K, long ld, float* dst, tgp_stream_t stream);

/* ------------------------------------------------------------------ chamfer (losses/chamfer3D) */

/* chamfer_3D.forward, chamfer_cuda.cpp:17-19 -> chamfer3D.cu:136-154 (NmDistanceKernel x2).
 * xyz1 (B,n,3), xyz2 (B,m,3) -> dist1 (B,n), dist2 (B,m) fp32, idx1 (B,n), idx2 (B,m) int32.
 * sums (B,4) optional: [sum dist1, sum dist2, sum sqrt(dist1), sum sqrt(dist2)] (calc_cd,
 * losses/TDA_loss_sym_recon.py:495-509); must be zeroed by the caller. */
int tgp_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int n, int m,
                    float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* sums, tgp_stream_t stream);

/* chamfer_3D.backward, chamfer_cuda.cpp:22-27 -> chamfer3D.cu:155-195.  gradxyz1/2 are
 * OVERWRITTEN (no pre-zeroing needed): the direct terms are stored, the scattered terms are
 * added with fp32 atomics like the reference, so the last bits depend on arrival order. */
int tgp_chamfer_bwd(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                    const int32_t* idx1, const int32_t* idx2, int B, int n, int m,
                    float* gradxyz1, float* gradxyz2, tgp_stream_t stream);

/* calc_dcd, losses/TDA_loss_sym_recon.py:411-450, on the outputs of tgp_chamfer_fwd (frac = n/m resp. m/n; non_reg != 0
 * clamps both to >= 1, :418-420):
 * loss (B) = mean_i(1 - exp(-alpha d1_i) w1_i) + 0.5 mean_j(1 - exp(-alpha d2_j) w2_j) with the bincount weights
 * w = (count[idx]^n_lambda + 1e-6)^-1 * frac built in shared memory (replaces the reference's Python loop over
 * the batch with torch.bincount).  coef1 (B,n) / coef2 (B,m), optional: d loss[b] / d dist (weights detached). */
int tgp_dcd(const float* dist1, const float* dist2, const int32_t* idx1, const int32_t* idx2, int B, int n, int m,
            float alpha, float n_lambda, int non_reg, float* loss, float* coef1, float* coef2, tgp_stream_t stream);

/* ------------------------------------------------------------------ backward (SURVEY 8a', north_star item 5)
 * The reference has no hand-written backward for gcn3d: torch autograd differentiates the graph built by
 * gcn3d.py:78-112 (HSlayer_surface), :142-186 (HS_layer), :210-217 (ORL), :225-245 (Pool_layer) and
 * FaceRecon.py:69-73 (nearest upsampling).  These entry points are what the torch.autograd.Function
 * wrappers of the drop-in module call instead.  Indices carry no gradient and vertices never require one
 * (trainer/RL_TDA.py:111).  Scatter-adds use fp32 atomics (as ATen's index backward does); reductions across
 * clouds go through fixed-order partial sums in the caller-provided workspace. */

/* backward of a fused epilogue activation: gz = grad * [y > 0 (if relu)] * scale[col] (scale may be NULL).
 * grad, y, gz: (M, C) with row strides ld_*. */
int tgp_act_bwd(const float* grad, long ld_grad, const float* y, long ld_y, const float* scale, int relu,
                long M, int C, float* gz, long ld_gz, tgp_stream_t stream);

/* per-group column sums: out[g, c] = sum of x[r, c] over the rows_per_group rows of group g (bias gradients,
 * the per-cloud ORL bias gradient).  x (M, C) row stride ld, out (M / rows_per_group, C). */
size_t tgp_colsum_workspace(long M, int C, long rows_per_group);
int tgp_colsum(const float* x, long ld, long M, int C, long rows_per_group, float* out,
               void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* backward of tgp_gather_max / tgp_orl_global (torch.max over the gathered neighbours, gcn3d.py:215,239):
 * dfeat[b, idx[b, rows[m], arg[b,m,c]], c] += scale * grad[b,m,c]; dfeat (B,N,C) is ACCUMULATED into.
 * grad_is_per_cloud: grad is (B,C), broadcast over m (the mean over N of the ORL, scale = 1/N). */
int tgp_gather_max_bwd(const float* grad, int grad_is_per_cloud, float scale, const void* idx, int idx_bits,
                       const int64_t* rows, const uint8_t* arg, int B, int N, int M, int k, int C,
                       float* dfeat, tgp_stream_t stream);

/* backward of tgp_gather_rows (indexing_neighbor_new, gcn3d.py:38-46; the nearest upsampling FaceRecon.py:71-73):
 * dtensor[b, index[b,m,j], :] += grad[b,m,j,:]; dtensor (B,N,C) is ACCUMULATED into. */
int tgp_scatter_add_rows(const float* grad, const void* index, int idx_bits, int B, int N, int M, int k, int C,
                         float* dtensor, tgp_stream_t stream);

/* backward of tgp_layer_conv_fwd w.r.t. the support features and the support directions
 * (the centre term's gradient is grad itself).  grad (B*N, C) row stride ld_grad;
 * d_support: row-major (B*N, S*C) block with row stride ld_ds whose columns are in SLAB order (cgroup, s, c4),
 *            i.e. the column order of the packed projection weight -- fully overwritten;
 * d_directions (3, S*C): gradient of the raw `directions` parameter (through F.normalize(dim=0)) -- overwritten. */
size_t tgp_layer_conv_bwd_workspace(int B, int S, int C);
int tgp_layer_conv_bwd(const float* edge_rec, const float* directions, const float* support_slab,
                       const uint8_t* arg_slab, const float* grad, long ld_grad, int B, int N, int k, int S, int C,
                       float* d_support, long ld_ds, float* d_directions,
                       void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* backward of tgp_surface_conv_fwd: only `directions` receives a gradient. arg (B,N,S*C) from the forward. */
size_t tgp_surface_conv_bwd_workspace(int B, int N, int S, int C);
int tgp_surface_conv_bwd(const float* xyz, const void* idx, int idx_bits, const float* directions,
                         const uint8_t* arg, const float* grad, long ld_grad, int B, int N, int k, int S, int C,
                         float* d_directions, void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* weight-gradient contraction out (K1,K2; row stride ldo) = A^T B with A (M,K1), B (M,K2) row-major
 * (dW = x^T dY of `feature_map @ weights`, gcn3d.py:170, and of every 1x1 Conv1d on the path).
 * tgp_gemm_tn: exact fp32 FMA path on the raw operands (small / odd shapes).
 * tgp_gemm_tn_tc: tcgen05 path; operands are the K-blocked TRANSPOSED splits: mixed = 0 (3xTF32) from tgp_split_tf32_t;
 *                 mixed = 1 (fp16 + bf16 cross terms,
 *                 see tgp_gemm_args.mixed) from tgp_split_mixed_t (K-blocked layout).
 *                 Split-K over M, partial sums added in a fixed order (deterministic). */
size_t tgp_gemm_tn_workspace(long M, int K1, int K2);
int tgp_gemm_tn(const float* A, long lda, const float* Bm, long ldb, long M, int K1, int K2, float* out, long ldo,
                void* workspace, size_t workspace_bytes, tgp_stream_t stream);
size_t tgp_gemm_tn_tc_workspace(long M, int K1, int K2);
int tgp_gemm_tn_tc(const float* At_split, const float* Bt_split, long M, int K1, int K2, float* out, long ldo,
                   int mixed, void* workspace, size_t workspace_bytes, tgp_stream_t stream);
/* tgp_gemm_tn_tc_rm: the same contraction on ROW-MAJOR mixed operands read in place (no transposing split): A_mixed (M rows,
 *                 4*kp_a 16-bit slots per row) and B_mixed (M rows, 4*kp_b slots) as written by tgp_split_mixed, tgp_gemm
 *                 mode 4 or tgp_affine_act(mixed); they enter the MMA as MN-major operands.  kp_a >= K1, kp_b >= K2,
 *                 multiples of 64.  Workspace: tgp_gemm_tn_tc_workspace(M, K1, K2). */
int tgp_gemm_tn_tc_rm(const float* A_mixed, int kp_a, const float* B_mixed, int kp_b, long M, int K1, int K2,
                      float* out, long ldo, void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* ------------------------------------------------------------------ heads in training
 * Train-mode Conv1d(k=1) + BatchNorm1d + ReLU/LeakyReLU stacks of the heads (PoseR.py:26-33, PoseTs.py:31-38,
 * FaceRecon.py:95-117,139-141) on channel-last rows; the contraction itself is tgp_gemm / tgp_gemm_tn_tc.
 * All matrices (M, C) fp32 with the given row strides; per-channel vectors (C). */

/* out[c] = sum_r (x[r,c] - mu[c])^2   (second pass of the batch variance).  workspace: tgp_bn_workspace(M, C). */
size_t tgp_bn_workspace(long M, int C);
int tgp_colsumsq_dev(const float* x, long ld, long M, int C, const float* mu, float* out,
                     void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* y = act(z * scale[c] + shift[c]), act(v) = v > 0 ? v : v * slope.  out (M,C) and/or out_split (M, 2*Kp) as the
 * next contraction's tensor-core operand ([tf32 | residual], or the MIXED layout when mixed = 1; padding columns must
 * be pre-zeroed). */
int tgp_affine_act(const float* z, long ld_z, const float* scale, const float* shift, float slope, long M, int C,
                   float* out, long ld_out, float* out_split, int Kp, int mixed, tgp_stream_t stream);

/* backward of y = act(BN_train(z)): dbeta[c] = sum g, dgamma[c] = sum g * zhat, g = dy * act'(y),
 * dz = gamma * invstd * (g - dbeta / M - zhat * dgamma / M).  workspace: tgp_bn_workspace(M, C).
 * dz_mixed (optional): dz also written as the MIXED operand (M, 8*tgp_mixed_kpad(C) bytes) of the dx contraction.
 * scale / shift (optional, the forward's tgp_affine_act arguments): the activation mask is recomputed from z
 * (y = act(fma(z, scale, shift)), bit-identical) instead of being read from y. */
int tgp_bn_bwd(const float* dy, long ld_dy, const float* y, long ld_y, const float* z, long ld_z,
               const float* mean, const float* invstd, const float* gamma, const float* scale, const float* shift,
               float slope, long M, int C,
               float* dz, long ld_dz, float* dz_mixed, float* dbeta, float* dgamma,
               void* workspace, size_t workspace_bytes, tgp_stream_t stream);

/* ------------------------------------------------------------------ optimiser step
 * trainer/RL_TDA.py:223-224: torch.nn.utils.clip_grad_norm_(net1.parameters(), 5); optimizer.step() with
 * optimizer = Ranger (tools/torch_utils/solver/ranger2020.py:44-235: RAdam + Lookahead + gradient centralisation,
 * gc_loc = True).  Parameters, gradients, exp_avg, exp_avg_sq and the Lookahead slow copy are five flat fp32 arenas
 * with identical element offsets; the row table says how they are cut. */

/* `len` consecutive elements from element `off` of every arena, belonging to parameter tensor `tensor`.
 * gc != 0: the row is a dim-0 slice of a centralised tensor and its mean is removed from the gradient
 * (centralized_gradient, ranger2020.py:31-41); gc == 0: a piece of a tensor that is not centralised. */
typedef struct {
    long long off;
    int len;
    int gc;
    int tensor;
    int pad_;
} tgp_ranger_row;

/* per-step scalars, computed by the host as ranger2020.py:179-201 does */
typedef struct {
    float beta1, beta2, eps, weight_decay;
    float one_minus_beta1, one_minus_beta2;   /* rounded from the host's doubles like the `alpha=` / `value=` of :174,177
                                                 (1.f - beta2 in fp32 would be off by 5e-5 relative) */
    float neg_step;     /* -(step_size * lr), rounded from double like the `alpha=` of p.add_ (:220) */
    int rectified;      /* N_sma > N_sma_threshhold (:208): divide by sqrt(exp_avg_sq) + eps */
    int lookahead;      /* step % k == 0 (:225) */
    float la_alpha;     /* Lookahead interpolation factor */
    float max_norm;     /* > 0: gradients are scaled by min(1, max_norm / (total_norm + 1e-6)) as clip_grad_norm_ does */
    int gc_on_update;   /* 0: gc_loc = True, rows with gc != 0 lose the mean of the GRADIENT (:170-171, the default);
                           1: gc_loc = False, they lose the mean of the update G_grad (:217-218) */
} tgp_ranger_hyper;

/* pass 1: row_sum[r] = sum of row r's gradient (all n_rows rows), *sumsq = sum of squares of every active element
 * (zeroed by the call).  active_dev (optional, int per tensor): 0 = the tensor has no gradient this step. */
int tgp_ranger_reduce(const float* grads, const tgp_ranger_row* rows_dev, int n_rows, const int* active_dev,
                      float* row_sum, double* sumsq, tgp_stream_t stream);

/* pass 2 over rows [row_begin, row_end) (one call per parameter group): g' = clip * (g - mean_row); moments; RAdam
 * step; Lookahead.  sumsq may be NULL when max_norm <= 0; total_norm (optional, device float) = sqrt(*sumsq). */
int tgp_ranger_update(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* slow,
                      const tgp_ranger_row* rows_dev, int row_begin, int row_end, const int* active_dev,
                      const float* row_sum, const double* sumsq, const tgp_ranger_hyper* hyper_host,
                      float* total_norm, tgp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TGPOSE_B200_H */
