// Internal (C++) interface between abi.cu and the kernel translation units.
#pragma once

#include "common.cuh"

namespace sim {

// implemented in the kernel translation units
int fps(const float*, int, int, int, int*, float*, cudaStream_t, int pointnet2 = 0, int fma = 0);
int knn_group(const float*, const float*, int, int, int, int, int*, float*, float*, cudaStream_t, int fma = 0);
int add_layernorm(const void*, const void*, const float*, const float*, const float*, float*, void*, long, int, float,
                  int, int, cudaStream_t, void* planes = nullptr, long plane = 0, const float* row_scale = nullptr,
                  int rows_per_sample = 0, int plane_fmt = 0);  // plane_fmt: 0 = three bf16 planes, 1 = two fp16 planes
int order_gather_fwd(const void*, const void*, const int*, void*, void*, int, int, int, int, int, int, cudaStream_t);
int order_gather_bwd(const void*, const int*, void*, int, int, int, int, int, int, cudaStream_t);
int gather_rows(const void*, const int*, const void*, void*, int, int, int, int, int, cudaStream_t);
int argsort_rows(const float*, long, long, int, int, int*, int*, cudaStream_t);
int causal_conv1d_fwd(const void*, long, const float*, const float*, void*, long, int, int, int, int, int, int,
                      cudaStream_t, void* planes = nullptr, long ld_p = 0, long plane = 0);

struct ScanParams {
  const void* u;
  const void* delta;
  const void* z;
  const void* Bm;
  const void* Cm;
  void* out;
  const float* A;
  const float* Dv;
  const float* dbias;
  long ld_u, ld_delta, ld_z, ld_B, ld_C, ld_out;
  int batch, L, D;
  int softplus;
  int z_gate = 0;  // 1: `z` already holds the gate silu(z) (applied by the in_proj epilogue, gemm_split3.cu EpiAct mode 1)
  float* ckpt;  // optional (batch, ceil(L/kScanCkpt), D, 16) fp32: state BEFORE every kScanCkpt-th step (for backward)
  // optional (fp32 activations, warp-specialised kernel): write the result as three bf16 planes instead of `out`
  // (operand format of gemm_split3.cu: out_proj consumes it directly); plane q at out_planes + q * plane elements
  void* out_planes = nullptr;
  long ld_planes = 0, plane = 0;
  // optional (warp-specialised kernel): fused dt_proj.  `delta` then points at the x_proj output rows
  // (dt_low[24] | B[16] | C[16], row stride ld_delta; Bm / Cm unused) and wdt at the bf16 planes of dt_proj.weight,
  // (3, D, 32) for fp32 activations / (1, D, 32) for bf16, K zero-padded from 24 to 32
  const void* wdt = nullptr;
};
constexpr int kScanTile = 16;  // time steps per tile of the forward scan kernels
constexpr int kScanCkpt = 8;   // checkpoint interval of the training forward == tile of the backward kernel (kScanTile % kScanCkpt == 0)
int selective_scan_fwd(const ScanParams&, int, int, cudaStream_t);

struct ScanBwdParams {
  const void* u;
  const void* delta;
  const void* z;
  const void* Bm;
  const void* Cm;
  const void* dout;
  const float* A;
  const float* Dv;
  const float* dbias;
  const float* ckpt;
  void* du;
  void* ddelta;
  void* dz;
  float* dB;      // (batch*L, 16) fp32, accumulated into (caller zeroes)
  float* dC;
  float* dA;      // (D, 16) fp32, accumulated into
  float* dD;      // (D)
  float* ddbias;  // (D)
  long ld_u, ld_delta, ld_z, ld_B, ld_C, ld_dout, ld_du, ld_ddelta, ld_dz;
  int batch, L, D;
  int softplus;
};
int selective_scan_bwd(const ScanBwdParams&, int, cudaStream_t);
int gemm_bf16(const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* Y, long ldd, int out_bf16, int M,
              int N, int K, int splits, cudaStream_t stream);
int gemm_tf32(const float* A, long lda, int a_mn, const float* B, long ldb, int b_mn, void* Y, long ldd, int out_bf16, int M,
              int N, int K, int splits, const float* bias, int relu, cudaStream_t stream);
int gemm_tf32_group(const float* A, long lda, const float* B, long ldb, float* Y, long ldd, int M, int N, int K, const float* bias,
                    const float* gbias, long ld_gbias, int relu, float* gmax, long ld_gmax, cudaStream_t stream);
int gemm_bf16_silu(const void* A, long lda, const void* B, long ldb, void* Y, long ldd, int out_bf16, int M, int N, int K,
                   int silu_col0, cudaStream_t stream);
int causal_conv1d_bwd(const void*, long, const float*, const float*, const void*, long, void*, long, float*, float*,
                      int, int, int, int, int, int, cudaStream_t);

struct SpectralParams {
  const float* center;
  float* eigvals;
  float* eigvecs;
  int* perm;
  int* inv_perm;
  float* adjacency;
  double* ws_mat;
  float* ws_adj;
  int B, G, k_nn, k;
  float alpha;
  int symmetric, self_loop, binary;
  int matrix_sym;
  int eps_clamp;
  int smallest;
  int sign_rule;
  int lu_alias;
  // extensions (sim_spectral_eig_ex): sigma != NULL selects the alpha == 0 weights exp(-d^2 / (2 sigma^2)) of
  // create_graph_from_centers (:647), *sigma = mean pairwise distance over the whole batch (device scalar);
  // adj_in != NULL replaces the graph construction by a caller-supplied (B, G, G) adjacency
  // (calc_top_k_eigenvalues_eigenvectors(adj, k, smallest), :717); first = index of the first wanted eigenpair.
  const float* sigma;
  const float* adj_in;
  int first;
};
int pairwise_dist_mean(const float* center, int B, int G, double* partial, float* sigma, cudaStream_t stream);
size_t spectral_workspace_bytes(int, int, int);
int spectral_eig(SpectralParams, void*, size_t, cudaStream_t);

int mae_index_maps(const int* perm, const unsigned char* mask, int B, int k, int G, int n_vis, int* perm_full,
                   unsigned char* mask_full, int* restore_src, int* src_vis, int* vis_pos, int* rec_src, int* inv_vis,
                   int* err_flag, cudaStream_t stream);
int gather_sum_rows(const void* x, const int* idx, void* out, int B, int R_in, int R_out, int J, int C, int dtype,
                    cudaStream_t stream);
int masked_colsum(const void* x, const int* sel, long rows, int C, float* dfill, int dtype, cudaStream_t stream);
int invert_row_map(const int* src_idx, int B, int R_in, int R_out, int J, int* inv, int* err_flag, cudaStream_t stream);
int three_nn_interp_fwd(const float* xyz1, const float* xyz2, const float* points2, int B, int N, int S, int C, float* out,
                        int* idx, float* weight, cudaStream_t stream);
int three_interp_bwd(const float* dout, const int* idx, const float* weight, int B, int N, int S, int C, float* dpoints2,
                     cudaStream_t stream);
int chamfer_l2_fwd(const float* x, const float* y, long R, int P, int Q, float* loss, int* idx_x, int* idx_y,
                   cudaStream_t stream);
int chamfer_l2_bwd(const float* x, const float* y, const int* idx_x, const int* idx_y, const float* gloss, long R, int P,
                   int Q, float* dx, float* dy, cudaStream_t stream);
int add_layernorm_bwd(const float* res, const void* dy, const float* dres_out, const float* gamma, float* dres,
                      float* dgamma, float* dbeta, long rows, int C, float eps, int dtype_y, cudaStream_t stream,
                      void* dx = nullptr, int dtype_dx = 0, const float* row_scale = nullptr, int rows_per_sample = 0);
int group_max(const void* x, void* out, long groups, int M, int C, int dtype, cudaStream_t stream);
int group_bias_relu(void* x, const void* gvec, long rows, int M, int C, int dtype, cudaStream_t stream);
int adamw_flat(float* p, const float* g, float* m, float* v, const float* wd, long n, const float* lr, float* step,
               const float* grad_scale, float beta1, float beta2, float eps, cudaStream_t stream);
int point_linear3(const float* x, const float* w, const float* b, float* y, long rows, int C, int act, cudaStream_t stream);
int mlp3_relu_rows(const float* x, long ldx, long rows, int d0, const float* w1t, const float* b1, int d1, const float* w2t,
                   const float* b2, int d2, const float* w3t, const float* b3, int d3, float* y, long ldy,
                   cudaStream_t stream);
int layernorm_mean(const float* x, const float* gamma, const float* beta, float* out, int B, int L, int C, float eps,
                   cudaStream_t stream);
int gemm_f32a_bf16x3(const float* X, long ldx, const void* Ws, long ldw, long wplane, float* Y, long ldd, int M, int N,
                     int K, cudaStream_t stream, void* po = nullptr, int po_cols = 0, long po_ld = 0, long po_plane = 0,
                     const float* conv_w = nullptr, const float* conv_b = nullptr, float* U = nullptr, long ldu = 0,
                     int batch = 0, int L = 0);
int split3_bf16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, cudaStream_t stream);
int split2_f16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, cudaStream_t stream);
int gemm_planes(int np, const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd,
                int M, int N, int K, cudaStream_t stream, void* po, int po_cols, long po_ld, long po_plane, int act_mode,
                int act_col0, const float* act_bias);
int split3_bf16_t(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, cudaStream_t stream);
int gemm_bf16x3(const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd, int M,
                int N, int K, cudaStream_t stream, void* po = nullptr, int po_cols = 0, long po_ld = 0, long po_plane = 0);

}  // namespace sim
