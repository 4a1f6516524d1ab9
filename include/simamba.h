/* simamba.h - C ABI of libsimamba_b200.so (sm_100a): the B200-native hot path of SI-Mamba's
 * spectrally-ordered token encoder.
 *
 * The reference (denix56/SI-Mamba) is 100% Python; its native boundary sits inside third-party
 * wheels (mamba-ssm, causal-conv1d, pytorch3d) and ATen/cuSOLVER.  Each entry point below names
 * the reference call site it replaces (file:line relative to the upstream repo).  The host side
 * that binds these (si_mamba_b200/_lib.py, ctypes) mirrors the reference's nn.Module API.
 *
 * Conventions
 *  - every function returns 0 on success and a negative sim_status on failure; the message of the
 *    last failure on the calling thread is available from sim_last_error_string(); nothing throws.
 *  - all buffers are caller-allocated DEVICE memory passed as raw pointers with explicit sizes
 *    and row strides (in elements); no hidden allocation, no host synchronisation; work is
 *    enqueued on `stream` (a cudaStream_t) of the current device.
 *  - dtype codes: SIM_F32 = 0, SIM_BF16 = 1.  Index buffers are int32.
 *  - activation tensors of the Mamba mixer are TOKEN-major: (batch*L, D) row-major with a row
 *    stride, so the x / z halves of the in_proj output and the B / C column blocks of the x_proj
 *    output are consumed in place.  mamba-ssm's channel-major (B, D, L) view of the same data is
 *    a transpose; the Python wrapper accepts both.
 *  - the library is re-entrant: no global mutable state besides the thread-local error string.
 */
#ifndef SIMAMBA_H_
#define SIMAMBA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sim_stream_t; /* cudaStream_t */

enum sim_status {
  SIM_STATUS_OK = 0,
  SIM_STATUS_INVALID = -1,   /* bad argument / unsupported shape */
  SIM_STATUS_ALIGN = -2,     /* pointer or stride not aligned for vector / TMA access */
  SIM_STATUS_CUDA = -3,      /* CUDA runtime error (message holds cudaGetErrorString) */
  SIM_STATUS_WORKSPACE = -4  /* workspace missing or too small */
};

enum sim_dtype { SIM_F32 = 0, SIM_BF16 = 1 };

/* flags of sim_spectral_eig */
enum sim_spectral_flags {
  SIM_GRAPH_SYMMETRIC = 1 << 0,  /* config key `symmetric` */
  SIM_GRAPH_SELF_LOOP = 1 << 1,  /* config key `self_loop` */
  SIM_GRAPH_BINARY = 1 << 2,     /* config key `binary` */
  SIM_EIG_SMALLEST = 1 << 3,     /* config key `smallest` */
  SIM_LAP_SYMMETRIC = 1 << 4,    /* config key `matrix` != "laplacian": I - D^-1/2 A D^-1/2, first pair dropped */
  SIM_LAP_EPS_CLAMP = 1 << 5,    /* deg.clamp(min=1e-12) (batched MAE variant) instead of deg + 1e-6 */
  SIM_EIG_CANONICAL_SIGN = 1 << 6 /* work_order.py:360-365 sign rule (first entry >= 0) */
};

int sim_version(void);
const char* sim_last_error_string(void);

/* a-1  Farthest point sampling.  Replaces pytorch3d sample_farthest_points at
 * models/point_mamba.py:93 (seg: part_segmentation/models/pt_mamba.py:175).
 * xyz (B,N,3) f32 -> idx (B,G) i32, center (B,G,3) f32.  Start index 0, ties -> lowest index. */
int sim_fps(const float* xyz, int B, int N, int G, int32_t* idx, float* center, sim_stream_t stream);

/* a-1  kNN grouping + gather + centre subtraction.  Replaces pytorch3d knn_points and the index
 * gather at models/point_mamba.py:96-110.  idx (B,G,M) i32 ascending point index; nbr (centred)
 * and nbr_org (B,G,M,3) f32; either of nbr / nbr_org may be NULL. */
int sim_knn_group(const float* xyz, const float* center, int B, int N, int G, int M, int32_t* idx, float* nbr,
                  float* nbr_org, sim_stream_t stream);

/* sim_fps / sim_knn_group with a choice of distance arithmetic.  flags = 0 is the contract above: ((dx*dx)+(dy*dy))+(dz*dz),
 * every product and sum rounded to fp32 on its own.  SIM_DIST_FMA selects fma(dz,dz, fma(dy,dy, dx*dx)) - what nvcc's default
 * -fmad=true makes of the `dist2 += diff * diff` loops in pytorch3d's sample_farthest_points / knn_points device code.
 * Distances differ in the last bit, so near-tied candidates can swap: a maintainer pinning bit-exact indices against a
 * real pytorch3d build picks the flag that build agrees with (both are tested against their own oracle; the wheel is not
 * available to this repo). */
enum sim_distance_flags { SIM_DIST_FMA = 1 };
int sim_fps_ex(const float* xyz, int B, int N, int G, int32_t* idx, float* center, int flags, sim_stream_t stream);
int sim_knn_group_ex(const float* xyz, const float* center, int B, int N, int G, int M, int32_t* idx, float* nbr,
                     float* nbr_org, int flags, sim_stream_t stream);

/* a-3 + a-4 + a-5  centres -> kNN graph -> Laplacian -> k extremal eigenpairs -> argsort.
 * Replaces create_graph_from_* (models/point_mamba.py:620-715), calc_top_k_eigenvalues_eigenvectors
 * (:717-761, :3001-3050, symmetric :764-814) and the torch.sort of sort_points_by_fiedler (:817-826).
 * eigvals (B,k) f32, eigvecs (B,G,k) f32, perm (B,k,G) i32 with perm[b,s,r] = index of the r-th
 * smallest entry of eigenvector s (ties -> lower index), inv_perm (B,k,G) i32 or NULL,
 * adjacency (B,G,G) f32 or NULL (the scattered adjacency before symmetrisation). */
size_t sim_spectral_eig_workspace_bytes(int B, int G, int k);
int sim_spectral_eig(const float* center, int B, int G, int k_nn, float alpha, int flags, int k, float* eigvals,
                     float* eigvecs, int32_t* perm, int32_t* inv_perm, float* adjacency, void* workspace,
                     size_t workspace_bytes, sim_stream_t stream);

/* sim_spectral_eig with the remaining reference entry points folded in:
 *   adjacency_in (B,G,G) f32 or NULL: skip the graph construction and decompose the Laplacian of a caller-supplied
 *     adjacency - PointMamba.calc_top_k_eigenvalues_eigenvectors(adj_matrices, k, smallest) and its _symmetric twin
 *     (models/point_mamba.py:717-814); center may then be NULL;
 *   sigma (device scalar) or NULL: the `alpha == 0` weights exp(-d^2 / (2 sigma^2)) of create_graph_from_centers
 *     (:628, :647), sigma = mean of ALL pairwise centre distances of the batch (sim_pairwise_dist_mean);
 *   first: index (in the requested order, after the dropped pair of SIM_LAP_SYMMETRIC) of the first of the k wanted
 *     eigenpairs, so ceil(G / 8) calls return the full decomposition the reference's 4-tuple carries. */
int sim_spectral_eig_ex(const float* center, const float* adjacency_in, const float* sigma, int B, int G, int k_nn,
                        float alpha, int flags, int first, int k, float* eigvals, float* eigvecs, int32_t* perm,
                        int32_t* inv_perm, float* adjacency, void* workspace, size_t workspace_bytes,
                        sim_stream_t stream);
/* sigma = torch.mean(dist_matrix) of create_graph_from_centers (:626-628): centres (B,G,3) -> *sigma (device f32);
 * partial (B) f64 scratch.  Deterministic (per-cloud fp64 sums added in index order). */
int sim_pairwise_dist_mean(const float* center, int B, int G, double* partial, float* sigma, sim_stream_t stream);

/* a-5  stable ascending argsort of fp32 keys along rows (the torch.sort inside
 * sort_points_by_fiedler, models/point_mamba.py:820).  keys (rows, n) with row stride ld and element
 * stride es (so a column of (B,G,k) eigenvectors can be sorted in place): perm (rows, n) i32. */
int sim_argsort_rows(const float* keys, long ld, long es, int rows, int n, int32_t* perm, int32_t* inv_perm,
                     sim_stream_t stream);

/* a-6  SAST order assembly: out[b, s*G+r] = x[b, perm[b,s,r]] and, when reverse, the mirrored copy
 * (cat + flip + cat at models/point_mamba.py:889-898, 982-989).  x, x2 (B,G,C); o1, o2 (B,T,C) with
 * T = (reverse ? 2 : 1) * k * G.  x2/o2 may be NULL; x2 != NULL with o2 == NULL writes o1 = x[..] + x2[..]
 * (tokens + pos of MixerModel.forward, models/point_mamba.py:250, folded into the gather). */
int sim_order_gather_fwd(const void* x, const void* x2, const int32_t* perm, void* o1, void* o2, int B, int G,
                         int k, int C, int reverse, int dtype, sim_stream_t stream);
/* backward of sim_order_gather_fwd w.r.t. x: dx[b,g] = sum over the (reverse ? 2 : 1)*k rows that read g. */
int sim_order_gather_bwd(const void* dout, const int32_t* inv_perm, void* dx, int B, int G, int k, int C,
                         int reverse, int dtype, sim_stream_t stream);

/* a-8 / a-16 / a-17  general row gather: out[b,t] = src_idx[b,t] >= 0 ? x[b, src_idx[b,t]] : fill
 * (fill == NULL -> zeros).  HLT layout (part_segmentation/models/pt_mamba.py:670-723), MAE visible-token
 * compaction (models/point_mamba.py:2734-2772) and token restore (:3147-3197). */
int sim_gather_rows(const void* x, const int32_t* src_idx, const void* fill, void* out, int B, int R_in, int R_out,
                    int C, int dtype, sim_stream_t stream);

/* a-9 / a-13  res_out = x (+ x2) (+ res_in); y = LayerNorm(res_out) * gamma + beta
 * (models/block.py:56-58; models/point_mamba.py:250, 256-258).  res_* are f32; x2, res_in, res_out may be NULL. */
int sim_add_layernorm(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                      float* res_out, void* y, long rows, int C, float eps, int dtype_x, int dtype_y,
                      sim_stream_t stream);

/* a-12  causal depthwise conv1d (+ SiLU).  Replaces causal_conv1d_fn inside Mamba.forward
 * (models/block.py:72).  x, y (batch*L, D) token-major with row strides; w (D, width) f32; bias (D) f32 or NULL. */
int sim_causal_conv1d_fwd(const void* x, long ld_x, const float* w, const float* bias, void* y, long ld_y,
                          int batch, int L, int D, int width, int silu, int dtype, sim_stream_t stream);

/* a-11  selective scan forward.  Replaces mamba-ssm selective_scan_fn inside Mamba.forward
 * (models/block.py:72); semantics of selective_scan_ref.  u, delta, z, out (batch*L, D) token-major;
 * Bm, Cm (batch*L, N) token-major; A (D,N) f32; Dvec, delta_bias (D) f32 or NULL; z may be NULL.
 * N must be 16.  variant: 0 = auto, else states per thread (2, 4, 8, 16).
 * delta_softplus: bit 0 = apply softplus to delta + delta_bias; bit 1 (inference, also in sim_selective_scan_fwd_split3) =
 * `z` already holds the gate silu(z) (written by sim_gemm_planes act_mode 1), the kernel multiplies by it as is. */
int sim_selective_scan_fwd(const void* u, long ld_u, const void* delta, long ld_delta, const float* A,
                           const void* Bm, long ld_B, const void* Cm, long ld_C, const float* Dvec, const void* z,
                           long ld_z, const float* delta_bias, void* out, long ld_out, float* checkpoints, int batch,
                           int L, int D, int N, int delta_softplus, int dtype, int variant, sim_stream_t stream);

/* Training forward only: `checkpoints` (NULL for inference) receives the SSM state at the start of every
 * 8th step, fp32 (batch, ceil(L/8), D, 16); it is the only tensor saved for the backward pass (mamba-ssm
 * recomputes from chunk states in the same spirit).  Size in bytes: */
size_t sim_selective_scan_checkpoint_bytes(int batch, int L, int D);

/* a-11  selective scan backward (recompute from checkpoints).  du, ddelta, dz (batch*L, D) in the input dtype;
 * dB, dC (batch*L, N) f32, dA (D,N), dD, ddelta_bias (D) f32 are ACCUMULATED into (the caller zeroes them).
 * z / dz / Dvec / delta_bias / dD / ddelta_bias may be NULL. */
int sim_selective_scan_bwd(const void* u, long ld_u, const void* delta, long ld_delta, const float* A,
                           const void* Bm, long ld_B, const void* Cm, long ld_C, const float* Dvec, const void* z,
                           long ld_z, const float* delta_bias, const void* dout, long ld_dout,
                           const float* checkpoints, void* du, long ld_du, void* ddelta, long ld_ddelta, void* dz,
                           long ld_dz, float* dB, float* dC, float* dA, float* dD, float* ddelta_bias, int batch,
                           int L, int D, int N, int delta_softplus, int dtype, sim_stream_t stream);

/* a-12  causal conv1d backward (recomputes the pre-activation from x).  dx (batch*L, D) in the input dtype;
 * dw (D, width) and dbias (D) f32 are ACCUMULATED into (the caller zeroes them); dbias may be NULL. */
int sim_causal_conv1d_bwd(const void* x, long ld_x, const float* w, const float* bias, const void* dy, long ld_dy,
                          void* dx, long ld_dx, float* dw, float* dbias, int batch, int L, int D, int width,
                          int silu, int dtype, sim_stream_t stream);

/* a-5 / a-7 / a-8  spectral permutation from sort keys: stable ascending argsort of an eigenvector column
 * (sort_points_by_fiedler, models/point_mamba.py:817-826) or of the HLT bucket keys id + u with the caller's
 * tie-break noise (part_segmentation/models/pt_mamba.py:670-680).  Same contract as sim_argsort_rows. */
int sim_spectral_perm(const float* keys, long ld, long es, int rows, int n, int32_t* perm, int32_t* inv_perm,
                      sim_stream_t stream);

/* a-16 / a-17  MAE masked sort + token restore (models/point_mamba.py:2734-2796, 3147-3197).
 * sim_mae_index_maps: perm (B,k,G) i32 + mask (B,G) u8 (1 = masked, G - n_vis masked patches per cloud) -> every map
 *   of the layout, T = 2kG decoder positions, R_vis = 2k n_vis encoder rows, R_msk = T - R_vis:
 *     perm_full (B,T) patch behind each position; mask_full (B,T) u8; restore_src (B,T) encoder row of each position
 *     or -1 (mask token); src_vis (B,R_vis) patch of each encoder row; vis_pos (B,R_vis) position of each encoder row;
 *     rec_src (B,R_msk) masked positions, ascending; inv_vis (B,G,2k) encoder rows that show patch g (-1 = masked).
 *   *err_flag (device int, caller zeroes) receives b+1 if cloud b does not have exactly n_vis visible patches.
 * sim_mae_compact_fwd: x_vis[b,r] = tokens[b, src_vis[b,r]];   _bwd: dtokens[b,g] = sum_j dx_vis[b, inv_vis[b,g,j]].
 * sim_mae_restore_fwd: x_full[b,t] = restore_src >= 0 ? x_vis[b, restore_src] : mask_token;
 *   _bwd: dx_vis[b,r] = dx_full[b, vis_pos[b,r]], dmask_token[c] += sum of the masked rows (fp32, caller zeroes). */
int sim_mae_index_maps(const int32_t* perm, const unsigned char* mask, int B, int k, int G, int n_vis,
                       int32_t* perm_full, unsigned char* mask_full, int32_t* restore_src, int32_t* src_vis,
                       int32_t* vis_pos, int32_t* rec_src, int32_t* inv_vis, int32_t* err_flag, sim_stream_t stream);
int sim_mae_compact_fwd(const void* tokens, const int32_t* src_vis, void* x_vis, int B, int G, int R_vis, int C,
                        int dtype, sim_stream_t stream);
int sim_mae_compact_bwd(const void* dx_vis, const int32_t* inv_vis, void* dtokens, int B, int G, int R_vis, int J,
                        int C, int dtype, sim_stream_t stream);
int sim_mae_restore_fwd(const void* x_vis, const int32_t* restore_src, const void* mask_token, void* x_full, int B,
                        int R_vis, int T, int C, int dtype, sim_stream_t stream);
int sim_mae_restore_bwd(const void* dx_full, const int32_t* vis_pos, const int32_t* restore_src, void* dx_vis,
                        float* dmask_token, int B, int R_vis, int T, int C, int dtype, sim_stream_t stream);
/* out[b,r] = sum_j x[b, idx[b,r,j]] over idx >= 0: deterministic backward of any row gather given its inverse map */
int sim_gather_sum_rows(const void* x, const int32_t* idx, void* out, int B, int R_in, int R_out, int J, int C,
                        int dtype, sim_stream_t stream);
/* inverse of a row-gather map: inv[b,r,0..J) = ascending output rows t with src_idx[b,t] == r, -1 padded, so that the
 * backward of sim_gather_rows (HLT layout pt_mamba.py:670-723; MAE pos / reconstruction gathers point_mamba.py:3192-3197)
 * is sim_gather_sum_rows instead of an atomic scatter-add.  *err_flag = b + 1 if a row of cloud b has more than J
 * readers.  src_idx (B,R_out), inv (B,R_in,J). */
int sim_invert_row_map(const int32_t* src_idx, int B, int R_in, int R_out, int J, int32_t* inv, int32_t* err_flag,
                       sim_stream_t stream);

/* f-1  data-prep farthest-point sampling with pointnet2_ops semantics, as the runners call it right before the model
 * (utils/misc.py:14-21 fps(data, number); tools/runner_finetune.py:177-194): furthest_point_sample + gather_operation
 * in one kernel.  Start index 0, running minima initialised to 1e10, points with |p|^2 <= 1e-3 never visited,
 * FMA-contracted distance, upstream tie rule (lowest upstream thread, then lowest index).  xyz (B,N,3) ->
 * idx (B,npoint) i32, sampled (B,npoint,3). */
int sim_fps_pointnet2(const float* xyz, int B, int N, int npoint, int32_t* idx, float* sampled, sim_stream_t stream);

/* a-2 / a-14  row passes between the Encoder's library GEMMs and in the classification tail (inference):
 * sim_group_max       out[g,c] = max over the M rows of group g of x (rows = groups*M, C)   (torch.max(feature, dim=2),
 *                     models/point_mamba.py:66, 71);
 * sim_group_bias_relu x[p,c] = relu(x[p,c] + gvec[p / M, c]) in place: the conv over cat([global, local]) (:67-69) split
 *                     into a per-point and a per-patch GEMM, recombined here;
 * sim_layernorm_mean  out[b,c] += mean over the L tokens of LayerNorm(x[b,t,:])[c] (self.norm(x).mean(1), :1122-1123;
 *                     fp32, out zeroed by the caller);
 * sim_mlp3_relu_rows  y = W3 relu(W2 relu(W1 x + b1) + b2) + b3 per row: cls_head_finetune in eval mode (:1124-1130) with
 *                     BatchNorm folded into W / b by the caller; weights TRANSPOSED (in, out) fp32, widths d1, d2, d3 <= 256,
 *                     biases may be NULL. */
int sim_group_max(const void* x, void* out, long groups, int M, int C, int dtype, sim_stream_t stream);
/* a-2  y[p,c] = act(b[c] + <w[c,:], x[p,:]>) for 3-D points: x (rows,3), w (C,3), b (C) or NULL, y (rows,C) fp32, C % 4 == 0;
 * act 0 = none, 1 = ReLU (Encoder.first_conv[0..2] with eval BatchNorm folded, models/point_mamba.py:47-49), 2 = GELU (erf;
 * pos_embed[0..1], :470-474). */
int sim_point_linear3(const float* x, const float* w, const float* b, float* y, long rows, int C, int act, sim_stream_t stream);
int sim_group_bias_relu(void* x, const void* gvec, long rows, int M, int C, int dtype, sim_stream_t stream);
int sim_layernorm_mean(const float* x, const float* gamma, const float* beta, float* out, int B, int L, int C, float eps,
                       sim_stream_t stream);
int sim_mlp3_relu_rows(const float* x, long ldx, long rows, int d0, const float* w1t, const float* b1, int d1,
                       const float* w2t, const float* b2, int d2, const float* w3t, const float* b3, int d3, float* y,
                       long ldy, sim_stream_t stream);

/* a-19: 3-nearest-centre inverse-squared-distance interpolation (PointNetFeaturePropagation.forward,
 * part_segmentation/models/pointnet2_utils.py:273-311; replaces square_distance + full sort + index_points gathers).
 * xyz1 (B,N,3) query points, xyz2 (B,S,3) centres (3 <= S <= 2048), points2 (B,S,C) feature rows, all fp32 contiguous.
 * idx (B,N,3) = the three nearest centres in ascending (distance, index) order, weight (B,N,3) = normalised
 * 1/(dist + 1e-8); out (B,N,C) = sum_k weight_k * points2[idx_k] (points2 and out may both be NULL: indices only).
 * sim_three_interp_bwd: dpoints2 (B,S,C) = scatter-add of weight_k * dout (zeroed by the call; fp32 atomics). */
int sim_three_nn_interp_fwd(const float* xyz1, const float* xyz2, const float* points2, int B, int N, int S, int C,
                            float* out, int32_t* idx, float* weight, sim_stream_t stream);
int sim_three_interp_bwd(const float* dout, const int32_t* idx, const float* weight, int B, int N, int S, int C,
                         float* dpoints2, sim_stream_t stream);

/* a-18  Chamfer-L2 of R pairs of small point sets, x (R,P,3), y (R,Q,3) fp32, P, Q <= 256:
 * loss[r] = mean_i min_j |x_i - y_j|^2 + mean_j min_i |x_i - y_j|^2  (pytorch3d chamfer_distance(x, y,
 * batch_reduction=None)[0], models/point_mamba.py:2950, 3199-3213).  idx_x (R,P) / idx_y (R,Q) receive the arg-mins
 * (lowest index on ties) that sim_chamfer_l2_bwd differentiates through; dx or dy may be NULL. */
int sim_chamfer_l2_fwd(const float* x, const float* y, long R, int P, int Q, float* loss, int32_t* idx_x,
                       int32_t* idx_y, sim_stream_t stream);
int sim_chamfer_l2_bwd(const float* x, const float* y, const int32_t* idx_x, const int32_t* idx_y, const float* gloss,
                       long R, int P, int Q, float* dx, float* dy, sim_stream_t stream);

/* a-10  bf16 projection GEMM (hand-written TMA + tcgen05 + TMEM kernel, csrc/gemm_bf16.cu): the in_proj / x_proj /
 * dt_proj / out_proj products of Mamba.forward (models/block.py:72) under bf16 autocast (tools/runner_pretrain.py:243)
 * and the dgrad / wgrad GEMMs of their backward, read in place from row-major tensors:
 *   Y[M,N] = op(A) . op(B)^T, bf16 operands, fp32 accumulation in tensor memory;
 *   a_mn = 0: A is (M,K) with row stride lda;  a_mn = 1: A is (K,M) with row stride lda (contraction index = row);
 *   b_mn = 0: B is (N,K) with row stride ldb;  b_mn = 1: B is (K,N) with row stride ldb;
 *   forward  Y = X W^T : (X,0, W,0);   dgrad dX = dY W : (dY,0, W,1);   wgrad dW = dY^T X : (dY,1, X,1).
 *   out_bf16 = 1: Y bf16, else fp32 (row stride ldy, N and ldy multiples of 4).  splits > 1: split-K, every CTA ADDS its
 *   partial tile to a caller-zeroed fp32 Y; splits = 0 chooses a split that fills the SMs (weight gradients).
 * Operand bases 16-byte aligned, lda / ldb multiples of 8 elements; M, N, K arbitrary (tiles are zero-filled). */
int sim_gemm_bf16(const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* Y, long ldy, int out_bf16,
                  int M, int N, int K, int splits, sim_stream_t stream);

/* a-2  the same kernel on fp32 operands consumed as TF32 (tcgen05 kind::tf32), with an optional + bias[N] and ReLU
 * epilogue: the 1x1 convolutions of Encoder.forward (models/point_mamba.py:59-73) written as row-major GEMMs on the
 * (B*G*M, C) point matrix, at the precision the reference's Conv1d layers run at (torch default cudnn.allow_tf32 = True);
 * K-major operands only (a_mn = b_mn = 0: 32-bit MN-major tiles need another swizzle atom and are not built).
 * lda / ldb multiples of 4 elements. */
int sim_gemm_tf32(const float* A, long lda, int a_mn, const float* B, long ldb, int b_mn, void* Y, long ldy, int out_bf16,
                  int M, int N, int K, int splits, const float* bias, int relu, sim_stream_t stream);

/* a-2  sim_gemm_tf32 (K-major operands) with the per-patch glue of Encoder.forward in its epilogue; rows = points, every 32
 * consecutive rows = one patch (group_size 32 of every shipped config; M % 32 == 0):
 *   Y[M,N] (f32, may be NULL) = relu?(A . B^T + bias[N] + gbias[row / 32][N])   -- `cat([feature_global.expand, feature])`
 *                               followed by the second_conv's first Conv1d + BN + ReLU (models/point_mamba.py:67-69), the
 *                               global half of that conv computed once per patch and passed in as gbias (row stride ld_gbias);
 *   gmax[M / 32, N] (f32, may be NULL) = max of Y over each patch's rows                -- `torch.max(feature, dim=2)` (:66, :72).
 * bias / gbias may be NULL; at least one of gbias / gmax must be given. */
int sim_gemm_tf32_group(const float* A, long lda, const float* B, long ldb, float* Y, long ldy, int M, int N, int K,
                        const float* bias, const float* gbias, long ld_gbias, int relu, float* gmax, long ld_gmax,
                        sim_stream_t stream);

/* a-10  sim_gemm_bf16 (K-major operands, no split-K) whose output columns >= silu_col0 leave as silu(v), computed on the fp32
 * accumulator before the rounding to bf16: in_proj of the bf16 (autocast) inference mixer hands the scan the gate silu(z)
 * (pair with bit 1 of sim_selective_scan_fwd's delta_softplus), as sim_gemm_planes act_mode 1 does for fp32. */
int sim_gemm_bf16_silu(const void* A, long lda, const void* B, long ldb, void* Y, long ldy, int out_bf16, int M, int N, int K,
                       int silu_col0, sim_stream_t stream);

/* a-10  the same fp32-accurate projection from PRE-SPLIT operands (hand-written TMA + tcgen05 + TMEM kernel,
 * csrc/gemm_split3.cu).  sim_split3_bf16 writes x = x0 + x1 + x2 as three bf16 planes (plane q at out + q * plane
 * elements, row stride ldo); the weights are split once per model, activations by their producer.
 * sim_gemm_bf16x3: Y[M,N] = sum of the six leading plane products X_i . W_j^T, fp32 accumulation in tensor memory.
 * ldx / ldw / plane strides multiples of 8 elements, ldd and N multiples of 4. */
/* a-9 backward (training configs): res (rows,C) = the fp32 residual stream sim_add_layernorm wrote, dy = gradient of
 * its normalised output (dtype_y), dres_out = gradient arriving in the residual stream from later layers (fp32, may be
 * NULL).  dres (fp32) = gradient of both inputs of the add; dgamma / dbeta (fp32, C) are accumulated into. */
int sim_add_layernorm_bwd(const float* res, const void* dy, const float* dres_out, const float* gamma, float* dres,
                          float* dgamma, float* dbeta, long rows, int C, float eps, int dtype_y, sim_stream_t stream);

/* Producers that emit the split operand directly (fp32 activations), so no separate split pass is needed:
 * LayerNorm output -> in_proj, conv output (fp32 u for the scan AND planes for x_proj), scan output -> out_proj. */
/* 8f-4  torch.optim.AdamW (tools/builder.py:74, part_segmentation/main.py:201) over flat buffers of n floats (n % 4 == 0):
 * p, g, m = exp_avg, v = exp_avg_sq, wd = per-element weight decay (< 0: leave the element alone).  lr, step (the number
 * of steps taken so far, incremented by the call) and grad_scale (NULL = 1; the clip_grad_norm_ coefficient) are DEVICE
 * scalars, so the update can be captured in a CUDA graph:  g' = g * grad_scale;  p *= 1 - lr wd;  m = b1 m + (1 - b1) g';
 * v = b2 v + (1 - b2) g'^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),  t = step + 1. */
int sim_adamw_flat(float* p, const float* g, float* m, float* v, const float* wd, long n, const float* lr, float* step,
                   const float* grad_scale, float beta1, float beta2, float eps, sim_stream_t stream);

/* a-9 with DropPath folded in (training; models/block.py:59 `residual = self.drop_path(hidden_states) + residual`, timm
 * drop_path = x / keep * mask per sample): res_out = row_scale[row / rows_per_sample] * x + res_in, y = LN(res_out).
 * row_scale (samples) f32 = mask_b / keep_prob, or NULL (then identical to sim_add_layernorm without x2).
 * sim_add_layernorm_bwd_dx: sim_add_layernorm_bwd that also writes dx = row_scale * dres in x's dtype (dtype_dx; the
 * gradient of the x operand), so the backward needs neither a cast nor the DropPath multiply; row_scale may be NULL. */
int sim_add_layernorm_droppath(const void* x, const float* row_scale, int rows_per_sample, const float* res_in,
                               const float* gamma, const float* beta, float* res_out, void* y, long rows, int C, float eps,
                               int dtype_x, int dtype_y, sim_stream_t stream);
int sim_add_layernorm_bwd_dx(const float* res, const void* dy, const float* dres_out, const float* gamma,
                             const float* row_scale, int rows_per_sample, float* dres, void* dx, int dtype_dx, float* dgamma,
                             float* dbeta, long rows, int C, float eps, int dtype_y, sim_stream_t stream);
int sim_add_layernorm_split3(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                             float* res_out, void* planes, long plane, long rows, int C, float eps, int dtype_x,
                             sim_stream_t stream);
int sim_causal_conv1d_fwd_split3(const float* x, long ld_x, const float* w, const float* bias, float* y, long ld_y,
                                 void* planes, long ld_p, long plane, int batch, int L, int D, int width, int silu,
                                 sim_stream_t stream);
int sim_selective_scan_fwd_split3(const void* u, long ld_u, const void* delta, long ld_delta, const float* A,
                                  const void* Bm, long ld_B, const void* Cm, long ld_C, const float* Dvec,
                                  const void* z, long ld_z, const float* delta_bias, void* out_planes, long ld_planes,
                                  long plane, int batch, int L, int D, int N, int delta_softplus, sim_stream_t stream);
/* sim_gemm_bf16x3 for N <= 64 (x_proj) that ALSO writes the first planes_cols output columns as split bf16 planes
 * (row stride ld_p, plane stride `plane`): the operand of the dt_proj GEMM, without a separate split pass. */
int sim_gemm_bf16x3_split_out(const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y,
                              long ldd, int M, int N, int K, void* planes_out, int planes_cols, long ld_p, long plane,
                              sim_stream_t stream);

/* Narrow projection (N <= 64: x_proj) straight from the fp32 activation X (M,K) (row stride ldx): the three bf16 planes of X
 * are formed inside the kernel by transform warps, the weight comes pre-split; bit-identical to sim_split3_bf16 +
 * sim_gemm_bf16x3.  planes_out (optional): the first planes_cols output columns as split planes (dt_proj operand). */
int sim_gemm_f32a_bf16x3(const float* X, long ldx, const void* Ws, long ldw, long wplane, float* Y, long ldd, int M, int N,
                         int K, void* planes_out, int planes_cols, long ld_p, long plane, sim_stream_t stream);

/* a-12 + a-10 fused (fp32 inference): u = silu(causal_conv1d(x)) (width 4) AND x_dbl = u . W_x^T in one kernel.  x (batch*L, D)
 * token-major fp32 with row stride ld_x (the x half of the in_proj output, in place), conv_w (D,4), conv_b (D) or NULL;
 * u (batch*L, D) is written for the scan; Ws = split planes of x_proj.weight (N <= 64 rows); planes_out as above. */
int sim_conv_xproj_f32(const float* x, long ld_x, const float* conv_w, const float* conv_b, float* u, long ld_u,
                       const void* Ws, long ldw, long wplane, float* x_dbl, long ldd, int batch, int L, int D, int N,
                       void* planes_out, int planes_cols, long ld_p, long plane, sim_stream_t stream);

/* a-10 + a-11 fused: the selective scan with dt_proj computed in-kernel.  x_dbl = the x_proj output rows
 * (dt_low[dt_rank = 24] | B[16] | C[16], row stride ld_x); wdt_planes = dt_proj.weight as bf16 planes, K zero-padded to 32:
 * (3, D, 32) from sim_split3_bf16 for fp32 activations, (1, D, 32) for bf16.  delta = dt_low . W_dt^T never touches HBM
 * (mma.sync in the kernel's elementwise warps; 3 x bf16 split with fp32 accumulation for fp32 activations).  Exactly one of
 * out / out_planes is written (out_planes: fp32 activations only, see sim_selective_scan_fwd_split3).  Inference only. */
int sim_selective_scan_fwd_fused_dt(const void* u, long ld_u, const void* x_dbl, long ld_x, int dt_rank,
                                    const void* wdt_planes, const float* A, const float* Dvec, const void* z, long ld_z,
                                    const float* delta_bias, void* out, long ld_out, void* out_planes, long ld_planes,
                                    long plane, int batch, int L, int D, int N, int delta_softplus, int dtype,
                                    sim_stream_t stream);
int sim_split3_bf16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, sim_stream_t stream);
/* the three planes of the TRANSPOSE: x (rows,K) f32 (row stride ld) -> out[q][k][r] (row stride ldo >= rows, plane stride
 * `plane`): the K-major operands of the dgrad / wgrad GEMMs of an fp32 Linear (W^T, dY^T, X^T) in one pass */
int sim_split3_bf16_t(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, sim_stream_t stream);
int sim_gemm_bf16x3(const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd,
                    int M, int N, int K, sim_stream_t stream);

/* a-10  the plane GEMM with a choice of operand format and an optional activation fused into its epilogue (inference).
 * np = 3: three bf16 planes per operand (sim_split3_bf16), six products - any fp32 operand;
 * np = 2: two fp16 planes x0 = fp16(x), x1' = fp16(2^11 (x - x0)) (sim_split2_f16 / sim_add_layernorm_split2h), three
 *         products: half the tensor-core work at fp32-GEMM accuracy, valid for |x|, |w| < 65504 - used for in_proj, whose
 *         operand is a LayerNorm output (models/block.py:56-58) bounded by sqrt(C) max|gamma| + max|beta|.
 * act_mode 0: none.  1: columns >= act_col0 leave as silu(v) - in_proj hands the scan the gate silu(z) instead of z
 * (Mamba.forward -> selective_scan_fn(..., z=z), models/block.py:72; pair with bit 1 of sim_selective_scan_fwd's
 * delta_softplus).  2: every column leaves as softplus(v + act_bias[col]) - dt_proj hands the scan dt itself (then call the
 * scan with delta_bias = NULL, delta_softplus = 0).  Same device arithmetic as the scan's own pre-pass: bit-identical. */
int sim_gemm_planes(int np, const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd,
                    int M, int N, int K, int act_mode, int act_col0, const float* act_bias, sim_stream_t stream);
int sim_split2_f16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, sim_stream_t stream);
int sim_add_layernorm_split2h(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                              float* res_out, void* planes, long plane, long rows, int C, float eps, int dtype_x,
                              sim_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMAMBA_H_ */
