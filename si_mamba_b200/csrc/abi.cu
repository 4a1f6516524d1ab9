// extern "C" surface of libsimamba_b200.so - see include/simamba.h for the contract and the
// reference call site each entry point replaces.  No torch types, no exceptions across the ABI.

#include <stdarg.h>

#include "../../include/simamba.h"
#include "kernels.cuh"

namespace sim {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky-less error so the next call starts clean
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SIM_ERR_CUDA;
  }
  return SIM_OK;
}

}  // namespace sim

using sim::SIM_ERR_INVALID;

extern "C" {

int sim_version(void) { return 100; }

const char* sim_last_error_string(void) { return sim::g_err; }

int sim_fps(const float* xyz, int B, int N, int G, int32_t* idx, float* center, sim_stream_t stream) {
  return sim::fps(xyz, B, N, G, idx, center, static_cast<cudaStream_t>(stream));
}

int sim_knn_group(const float* xyz, const float* center, int B, int N, int G, int M, int32_t* idx, float* nbr,
                  float* nbr_org, sim_stream_t stream) {
  return sim::knn_group(xyz, center, B, N, G, M, idx, nbr, nbr_org, static_cast<cudaStream_t>(stream));
}

int sim_fps_ex(const float* xyz, int B, int N, int G, int32_t* idx, float* center, int flags, sim_stream_t stream) {
  return sim::fps(xyz, B, N, G, idx, center, static_cast<cudaStream_t>(stream), 0, (flags & SIM_DIST_FMA) != 0);
}

int sim_knn_group_ex(const float* xyz, const float* center, int B, int N, int G, int M, int32_t* idx, float* nbr,
                     float* nbr_org, int flags, sim_stream_t stream) {
  return sim::knn_group(xyz, center, B, N, G, M, idx, nbr, nbr_org, static_cast<cudaStream_t>(stream),
                        (flags & SIM_DIST_FMA) != 0);
}

size_t sim_spectral_eig_workspace_bytes(int B, int G, int k) { return sim::spectral_workspace_bytes(B, G, k); }

int sim_spectral_eig(const float* center, int B, int G, int k_nn, float alpha, int flags, int k, float* eigvals,
                     float* eigvecs, int32_t* perm, int32_t* inv_perm, float* adjacency, void* workspace,
                     size_t workspace_bytes, sim_stream_t stream) {
  sim::SpectralParams P;
  memset(&P, 0, sizeof(P));
  P.center = center;
  P.eigvals = eigvals;
  P.eigvecs = eigvecs;
  P.perm = perm;
  P.inv_perm = inv_perm;
  P.adjacency = adjacency;
  P.B = B, P.G = G, P.k_nn = k_nn, P.k = k;
  P.alpha = alpha;
  P.symmetric = (flags & SIM_GRAPH_SYMMETRIC) != 0;
  P.self_loop = (flags & SIM_GRAPH_SELF_LOOP) != 0;
  P.binary = (flags & SIM_GRAPH_BINARY) != 0;
  P.smallest = (flags & SIM_EIG_SMALLEST) != 0;
  P.matrix_sym = (flags & SIM_LAP_SYMMETRIC) != 0;
  P.eps_clamp = (flags & SIM_LAP_EPS_CLAMP) != 0;
  P.sign_rule = (flags & SIM_EIG_CANONICAL_SIGN) != 0;
  return sim::spectral_eig(P, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int sim_spectral_eig_ex(const float* center, const float* adjacency_in, const float* sigma, int B, int G, int k_nn,
                        float alpha, int flags, int first, int k, float* eigvals, float* eigvecs, int32_t* perm,
                        int32_t* inv_perm, float* adjacency, void* workspace, size_t workspace_bytes, sim_stream_t stream) {
  sim::SpectralParams P;
  memset(&P, 0, sizeof(P));
  P.center = center;
  P.adj_in = adjacency_in;
  P.sigma = sigma;
  P.first = first;
  P.eigvals = eigvals;
  P.eigvecs = eigvecs;
  P.perm = perm;
  P.inv_perm = inv_perm;
  P.adjacency = adjacency;
  P.B = B, P.G = G, P.k_nn = k_nn, P.k = k;
  P.alpha = alpha;
  P.symmetric = (flags & SIM_GRAPH_SYMMETRIC) != 0;
  P.self_loop = (flags & SIM_GRAPH_SELF_LOOP) != 0;
  P.binary = (flags & SIM_GRAPH_BINARY) != 0;
  P.smallest = (flags & SIM_EIG_SMALLEST) != 0;
  P.matrix_sym = (flags & SIM_LAP_SYMMETRIC) != 0;
  P.eps_clamp = (flags & SIM_LAP_EPS_CLAMP) != 0;
  P.sign_rule = (flags & SIM_EIG_CANONICAL_SIGN) != 0;
  return sim::spectral_eig(P, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int sim_gemm_bf16(const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* Y, long ldy, int out_bf16,
                  int M, int N, int K, int splits, sim_stream_t stream) {
  return sim::gemm_bf16(A, lda, a_mn, B, ldb, b_mn, Y, ldy, out_bf16, M, N, K, splits, static_cast<cudaStream_t>(stream));
}

int sim_gemm_tf32(const float* A, long lda, int a_mn, const float* B, long ldb, int b_mn, void* Y, long ldy, int out_bf16,
                  int M, int N, int K, int splits, const float* bias, int relu, sim_stream_t stream) {
  return sim::gemm_tf32(A, lda, a_mn, B, ldb, b_mn, Y, ldy, out_bf16, M, N, K, splits, bias, relu,
                        static_cast<cudaStream_t>(stream));
}

int sim_gemm_tf32_group(const float* A, long lda, const float* B, long ldb, float* Y, long ldy, int M, int N, int K,
                        const float* bias, const float* gbias, long ld_gbias, int relu, float* gmax, long ld_gmax,
                        sim_stream_t stream) {
  return sim::gemm_tf32_group(A, lda, B, ldb, Y, ldy, M, N, K, bias, gbias, ld_gbias, relu, gmax, ld_gmax,
                              static_cast<cudaStream_t>(stream));
}

int sim_gemm_bf16_silu(const void* A, long lda, const void* B, long ldb, void* Y, long ldy, int out_bf16, int M, int N, int K,
                       int silu_col0, sim_stream_t stream) {
  return sim::gemm_bf16_silu(A, lda, B, ldb, Y, ldy, out_bf16, M, N, K, silu_col0, static_cast<cudaStream_t>(stream));
}

int sim_pairwise_dist_mean(const float* center, int B, int G, double* partial, float* sigma, sim_stream_t stream) {
  return sim::pairwise_dist_mean(center, B, G, partial, sigma, static_cast<cudaStream_t>(stream));
}

int sim_argsort_rows(const float* keys, long ld, long es, int rows, int n, int32_t* perm, int32_t* inv_perm,
                     sim_stream_t stream) {
  return sim::argsort_rows(keys, ld, es, rows, n, perm, inv_perm, static_cast<cudaStream_t>(stream));
}

int sim_order_gather_fwd(const void* x, const void* x2, const int32_t* perm, void* o1, void* o2, int B, int G,
                         int k, int C, int reverse, int dtype, sim_stream_t stream) {
  return sim::order_gather_fwd(x, x2, perm, o1, o2, B, G, k, C, reverse, dtype, static_cast<cudaStream_t>(stream));
}

int sim_order_gather_bwd(const void* dout, const int32_t* inv_perm, void* dx, int B, int G, int k, int C,
                         int reverse, int dtype, sim_stream_t stream) {
  return sim::order_gather_bwd(dout, inv_perm, dx, B, G, k, C, reverse, dtype, static_cast<cudaStream_t>(stream));
}

int sim_gather_rows(const void* x, const int32_t* src_idx, const void* fill, void* out, int B, int R_in, int R_out,
                    int C, int dtype, sim_stream_t stream) {
  return sim::gather_rows(x, src_idx, fill, out, B, R_in, R_out, C, dtype, static_cast<cudaStream_t>(stream));
}

int sim_add_layernorm(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                      float* res_out, void* y, long rows, int C, float eps, int dtype_x, int dtype_y,
                      sim_stream_t stream) {
  return sim::add_layernorm(x, x2, res_in, gamma, beta, res_out, y, rows, C, eps, dtype_x, dtype_y,
                            static_cast<cudaStream_t>(stream));
}

int sim_causal_conv1d_fwd(const void* x, long ld_x, const float* w, const float* bias, void* y, long ld_y,
                          int batch, int L, int D, int width, int silu, int dtype, sim_stream_t stream) {
  return sim::causal_conv1d_fwd(x, ld_x, w, bias, y, ld_y, batch, L, D, width, silu, dtype,
                                static_cast<cudaStream_t>(stream));
}

int sim_selective_scan_fwd(const void* u, long ld_u, const void* delta, long ld_delta, const float* A,
                           const void* Bm, long ld_B, const void* Cm, long ld_C, const float* Dvec, const void* z,
                           long ld_z, const float* delta_bias, void* out, long ld_out, float* checkpoints, int batch,
                           int L, int D, int N, int delta_softplus, int dtype, int variant, sim_stream_t stream) {
  if (N != 16) {
    sim::set_error("sim_selective_scan_fwd: d_state must be 16 (got %d)", N);
    return SIM_ERR_INVALID;
  }
  sim::ScanParams p;
  p.u = u, p.delta = delta, p.z = z, p.Bm = Bm, p.Cm = Cm, p.out = out;
  p.A = A, p.Dv = Dvec, p.dbias = delta_bias;
  p.ld_u = ld_u, p.ld_delta = ld_delta, p.ld_z = ld_z, p.ld_B = ld_B, p.ld_C = ld_C, p.ld_out = ld_out;
  p.batch = batch, p.L = L, p.D = D, p.softplus = delta_softplus & 1, p.z_gate = (delta_softplus >> 1) & 1;
  p.ckpt = checkpoints;
  return sim::selective_scan_fwd(p, dtype, variant, static_cast<cudaStream_t>(stream));
}

int sim_selective_scan_fwd_split3(const void* u, long ld_u, const void* delta, long ld_delta, const float* A,
                                  const void* Bm, long ld_B, const void* Cm, long ld_C, const float* Dvec,
                                  const void* z, long ld_z, const float* delta_bias, void* out_planes, long ld_planes,
                                  long plane, int batch, int L, int D, int N, int delta_softplus, sim_stream_t stream) {
  if (N != 16) {
    sim::set_error("sim_selective_scan_fwd_split3: d_state must be 16 (got %d)", N);
    return SIM_ERR_INVALID;
  }
  sim::ScanParams p;
  p.u = u, p.delta = delta, p.z = z, p.Bm = Bm, p.Cm = Cm, p.out = nullptr;
  p.A = A, p.Dv = Dvec, p.dbias = delta_bias;
  p.ld_u = ld_u, p.ld_delta = ld_delta, p.ld_z = ld_z, p.ld_B = ld_B, p.ld_C = ld_C, p.ld_out = 0;
  p.batch = batch, p.L = L, p.D = D, p.softplus = delta_softplus & 1, p.z_gate = (delta_softplus >> 1) & 1;
  p.ckpt = nullptr;
  p.out_planes = out_planes, p.ld_planes = ld_planes, p.plane = plane;
  return sim::selective_scan_fwd(p, 0, 0, static_cast<cudaStream_t>(stream));
}

int sim_selective_scan_fwd_fused_dt(const void* u, long ld_u, const void* x_dbl, long ld_x, int dt_rank,
                                    const void* wdt_planes, const float* A, const float* Dvec, const void* z, long ld_z,
                                    const float* delta_bias, void* out, long ld_out, void* out_planes, long ld_planes,
                                    long plane, int batch, int L, int D, int N, int delta_softplus, int dtype,
                                    sim_stream_t stream) {
  if (N != 16 || dt_rank != 24) {
    sim::set_error("sim_selective_scan_fwd_fused_dt: built for d_state 16 and dt_rank 24 (got %d, %d)", N, dt_rank);
    return SIM_ERR_INVALID;
  }
  sim::ScanParams p;
  p.u = u, p.delta = x_dbl, p.z = z, p.Bm = nullptr, p.Cm = nullptr, p.out = out;
  p.A = A, p.Dv = Dvec, p.dbias = delta_bias;
  p.ld_u = ld_u, p.ld_delta = ld_x, p.ld_z = ld_z, p.ld_B = 0, p.ld_C = 0, p.ld_out = ld_out;
  p.batch = batch, p.L = L, p.D = D, p.softplus = delta_softplus;
  p.ckpt = nullptr;
  p.out_planes = out_planes, p.ld_planes = ld_planes, p.plane = plane;
  p.wdt = wdt_planes;
  return sim::selective_scan_fwd(p, dtype, 0, static_cast<cudaStream_t>(stream));
}

int sim_add_layernorm_split3(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                             float* res_out, void* planes, long plane, long rows, int C, float eps, int dtype_x,
                             sim_stream_t stream) {
  return sim::add_layernorm(x, x2, res_in, gamma, beta, res_out, nullptr, rows, C, eps, dtype_x, 0,
                            static_cast<cudaStream_t>(stream), planes, plane);
}

int sim_add_layernorm_split2h(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                              float* res_out, void* planes, long plane, long rows, int C, float eps, int dtype_x,
                              sim_stream_t stream) {
  return sim::add_layernorm(x, x2, res_in, gamma, beta, res_out, nullptr, rows, C, eps, dtype_x, 0,
                            static_cast<cudaStream_t>(stream), planes, plane, nullptr, 0, 1);
}

int sim_causal_conv1d_fwd_split3(const float* x, long ld_x, const float* w, const float* bias, float* y, long ld_y,
                                 void* planes, long ld_p, long plane, int batch, int L, int D, int width, int silu,
                                 sim_stream_t stream) {
  return sim::causal_conv1d_fwd(x, ld_x, w, bias, y, ld_y, batch, L, D, width, silu, 0,
                                static_cast<cudaStream_t>(stream), planes, ld_p, plane);
}

size_t sim_selective_scan_checkpoint_bytes(int batch, int L, int D) {
  if (batch <= 0 || L <= 0 || D <= 0) return 0;
  return (size_t)batch * ((L + sim::kScanCkpt - 1) / sim::kScanCkpt) * D * 16 * sizeof(float);
}

int sim_selective_scan_bwd(const void* u, long ld_u, const void* delta, long ld_delta, const float* A,
                           const void* Bm, long ld_B, const void* Cm, long ld_C, const float* Dvec, const void* z,
                           long ld_z, const float* delta_bias, const void* dout, long ld_dout,
                           const float* checkpoints, void* du, long ld_du, void* ddelta, long ld_ddelta, void* dz,
                           long ld_dz, float* dB, float* dC, float* dA, float* dD, float* ddelta_bias, int batch,
                           int L, int D, int N, int delta_softplus, int dtype, sim_stream_t stream) {
  if (N != 16) {
    sim::set_error("sim_selective_scan_bwd: d_state must be 16 (got %d)", N);
    return SIM_ERR_INVALID;
  }
  sim::ScanBwdParams p;
  p.u = u, p.delta = delta, p.z = z, p.Bm = Bm, p.Cm = Cm, p.dout = dout;
  p.A = A, p.Dv = Dvec, p.dbias = delta_bias, p.ckpt = checkpoints;
  p.du = du, p.ddelta = ddelta, p.dz = dz, p.dB = dB, p.dC = dC, p.dA = dA, p.dD = dD, p.ddbias = ddelta_bias;
  p.ld_u = ld_u, p.ld_delta = ld_delta, p.ld_z = ld_z, p.ld_B = ld_B, p.ld_C = ld_C, p.ld_dout = ld_dout;
  p.ld_du = ld_du, p.ld_ddelta = ld_ddelta, p.ld_dz = ld_dz;
  p.batch = batch, p.L = L, p.D = D, p.softplus = delta_softplus;
  return sim::selective_scan_bwd(p, dtype, static_cast<cudaStream_t>(stream));
}

int sim_causal_conv1d_bwd(const void* x, long ld_x, const float* w, const float* bias, const void* dy, long ld_dy,
                          void* dx, long ld_dx, float* dw, float* dbias, int batch, int L, int D, int width,
                          int silu, int dtype, sim_stream_t stream) {
  return sim::causal_conv1d_bwd(x, ld_x, w, bias, dy, ld_dy, dx, ld_dx, dw, dbias, batch, L, D, width, silu, dtype,
                                static_cast<cudaStream_t>(stream));
}

int sim_mae_index_maps(const int32_t* perm, const unsigned char* mask, int B, int k, int G, int n_vis,
                       int32_t* perm_full, unsigned char* mask_full, int32_t* restore_src, int32_t* src_vis,
                       int32_t* vis_pos, int32_t* rec_src, int32_t* inv_vis, int32_t* err_flag, sim_stream_t stream) {
  return sim::mae_index_maps(perm, mask, B, k, G, n_vis, perm_full, mask_full, restore_src, src_vis, vis_pos, rec_src,
                             inv_vis, err_flag, static_cast<cudaStream_t>(stream));
}

int sim_mae_compact_fwd(const void* tokens, const int32_t* src_vis, void* x_vis, int B, int G, int R_vis, int C,
                        int dtype, sim_stream_t stream) {
  return sim::gather_rows(tokens, src_vis, nullptr, x_vis, B, G, R_vis, C, dtype, static_cast<cudaStream_t>(stream));
}

int sim_mae_compact_bwd(const void* dx_vis, const int32_t* inv_vis, void* dtokens, int B, int G, int R_vis, int J,
                        int C, int dtype, sim_stream_t stream) {
  return sim::gather_sum_rows(dx_vis, inv_vis, dtokens, B, R_vis, G, J, C, dtype, static_cast<cudaStream_t>(stream));
}

int sim_mae_restore_fwd(const void* x_vis, const int32_t* restore_src, const void* mask_token, void* x_full, int B,
                        int R_vis, int T, int C, int dtype, sim_stream_t stream) {
  return sim::gather_rows(x_vis, restore_src, mask_token, x_full, B, R_vis, T, C, dtype,
                          static_cast<cudaStream_t>(stream));
}

int sim_mae_restore_bwd(const void* dx_full, const int32_t* vis_pos, const int32_t* restore_src, void* dx_vis,
                        float* dmask_token, int B, int R_vis, int T, int C, int dtype, sim_stream_t stream) {
  int rc = sim::gather_rows(dx_full, vis_pos, nullptr, dx_vis, B, T, R_vis, C, dtype, static_cast<cudaStream_t>(stream));
  if (rc || !dmask_token) return rc;
  return sim::masked_colsum(dx_full, restore_src, (long)B * T, C, dmask_token, dtype, static_cast<cudaStream_t>(stream));
}

int sim_gather_sum_rows(const void* x, const int32_t* idx, void* out, int B, int R_in, int R_out, int J, int C,
                        int dtype, sim_stream_t stream) {
  return sim::gather_sum_rows(x, idx, out, B, R_in, R_out, J, C, dtype, static_cast<cudaStream_t>(stream));
}

int sim_invert_row_map(const int32_t* src_idx, int B, int R_in, int R_out, int J, int32_t* inv, int32_t* err_flag,
                       sim_stream_t stream) {
  return sim::invert_row_map(src_idx, B, R_in, R_out, J, inv, err_flag, static_cast<cudaStream_t>(stream));
}

int sim_spectral_perm(const float* keys, long ld, long es, int rows, int n, int32_t* perm, int32_t* inv_perm,
                      sim_stream_t stream) {
  return sim::argsort_rows(keys, ld, es, rows, n, perm, inv_perm, static_cast<cudaStream_t>(stream));
}

int sim_three_nn_interp_fwd(const float* xyz1, const float* xyz2, const float* points2, int B, int N, int S, int C,
                            float* out, int32_t* idx, float* weight, sim_stream_t stream) {
  return sim::three_nn_interp_fwd(xyz1, xyz2, points2, B, N, S, C, out, idx, weight, static_cast<cudaStream_t>(stream));
}

int sim_three_interp_bwd(const float* dout, const int32_t* idx, const float* weight, int B, int N, int S, int C,
                         float* dpoints2, sim_stream_t stream) {
  return sim::three_interp_bwd(dout, idx, weight, B, N, S, C, dpoints2, static_cast<cudaStream_t>(stream));
}

int sim_chamfer_l2_fwd(const float* x, const float* y, long R, int P, int Q, float* loss, int32_t* idx_x,
                       int32_t* idx_y, sim_stream_t stream) {
  return sim::chamfer_l2_fwd(x, y, R, P, Q, loss, idx_x, idx_y, static_cast<cudaStream_t>(stream));
}

int sim_chamfer_l2_bwd(const float* x, const float* y, const int32_t* idx_x, const int32_t* idx_y, const float* gloss,
                       long R, int P, int Q, float* dx, float* dy, sim_stream_t stream) {
  return sim::chamfer_l2_bwd(x, y, idx_x, idx_y, gloss, R, P, Q, dx, dy, static_cast<cudaStream_t>(stream));
}

int sim_fps_pointnet2(const float* xyz, int B, int N, int npoint, int32_t* idx, float* sampled, sim_stream_t stream) {
  return sim::fps(xyz, B, N, npoint, idx, sampled, static_cast<cudaStream_t>(stream), 1);
}

int sim_add_layernorm_bwd(const float* res, const void* dy, const float* dres_out, const float* gamma, float* dres,
                          float* dgamma, float* dbeta, long rows, int C, float eps, int dtype_y, sim_stream_t stream) {
  return sim::add_layernorm_bwd(res, dy, dres_out, gamma, dres, dgamma, dbeta, rows, C, eps, dtype_y,
                                static_cast<cudaStream_t>(stream));
}

int sim_gemm_bf16x3_split_out(const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y,
                              long ldd, int M, int N, int K, void* planes_out, int planes_cols, long ld_p, long plane,
                              sim_stream_t stream) {
  return sim::gemm_bf16x3(Xs, ldx, xplane, Ws, ldw, wplane, Y, ldd, M, N, K, static_cast<cudaStream_t>(stream), planes_out,
                          planes_cols, ld_p, plane);
}

int sim_gemm_f32a_bf16x3(const float* X, long ldx, const void* Ws, long ldw, long wplane, float* Y, long ldd, int M, int N,
                         int K, void* planes_out, int planes_cols, long ld_p, long plane, sim_stream_t stream) {
  return sim::gemm_f32a_bf16x3(X, ldx, Ws, ldw, wplane, Y, ldd, M, N, K, static_cast<cudaStream_t>(stream), planes_out,
                               planes_cols, ld_p, plane);
}

int sim_conv_xproj_f32(const float* x, long ld_x, const float* conv_w, const float* conv_b, float* u, long ld_u,
                       const void* Ws, long ldw, long wplane, float* x_dbl, long ldd, int batch, int L, int D, int N,
                       void* planes_out, int planes_cols, long ld_p, long plane, sim_stream_t stream) {
  return sim::gemm_f32a_bf16x3(x, ld_x, Ws, ldw, wplane, x_dbl, ldd, batch * L, N, D, static_cast<cudaStream_t>(stream),
                               planes_out, planes_cols, ld_p, plane, conv_w, conv_b, u, ld_u, batch, L);
}

int sim_group_max(const void* x, void* out, long groups, int M, int C, int dtype, sim_stream_t stream) {
  return sim::group_max(x, out, groups, M, C, dtype, static_cast<cudaStream_t>(stream));
}

int sim_group_bias_relu(void* x, const void* gvec, long rows, int M, int C, int dtype, sim_stream_t stream) {
  return sim::group_bias_relu(x, gvec, rows, M, C, dtype, static_cast<cudaStream_t>(stream));
}

int sim_add_layernorm_droppath(const void* x, const float* row_scale, int rows_per_sample, const float* res_in,
                               const float* gamma, const float* beta, float* res_out, void* y, long rows, int C, float eps,
                               int dtype_x, int dtype_y, sim_stream_t stream) {
  return sim::add_layernorm(x, nullptr, res_in, gamma, beta, res_out, y, rows, C, eps, dtype_x, dtype_y,
                            static_cast<cudaStream_t>(stream), nullptr, 0, row_scale, rows_per_sample);
}

int sim_add_layernorm_bwd_dx(const float* res, const void* dy, const float* dres_out, const float* gamma,
                             const float* row_scale, int rows_per_sample, float* dres, void* dx, int dtype_dx, float* dgamma,
                             float* dbeta, long rows, int C, float eps, int dtype_y, sim_stream_t stream) {
  return sim::add_layernorm_bwd(res, dy, dres_out, gamma, dres, dgamma, dbeta, rows, C, eps, dtype_y,
                                static_cast<cudaStream_t>(stream), dx, dtype_dx, row_scale, rows_per_sample);
}

int sim_adamw_flat(float* p, const float* g, float* m, float* v, const float* wd, long n, const float* lr, float* step,
                   const float* grad_scale, float beta1, float beta2, float eps, sim_stream_t stream) {
  return sim::adamw_flat(p, g, m, v, wd, n, lr, step, grad_scale, beta1, beta2, eps, static_cast<cudaStream_t>(stream));
}

int sim_point_linear3(const float* x, const float* w, const float* b, float* y, long rows, int C, int act, sim_stream_t stream) {
  return sim::point_linear3(x, w, b, y, rows, C, act, static_cast<cudaStream_t>(stream));
}

int sim_mlp3_relu_rows(const float* x, long ldx, long rows, int d0, const float* w1t, const float* b1, int d1,
                       const float* w2t, const float* b2, int d2, const float* w3t, const float* b3, int d3, float* y,
                       long ldy, sim_stream_t stream) {
  return sim::mlp3_relu_rows(x, ldx, rows, d0, w1t, b1, d1, w2t, b2, d2, w3t, b3, d3, y, ldy,
                             static_cast<cudaStream_t>(stream));
}

int sim_layernorm_mean(const float* x, const float* gamma, const float* beta, float* out, int B, int L, int C, float eps,
                       sim_stream_t stream) {
  return sim::layernorm_mean(x, gamma, beta, out, B, L, C, eps, static_cast<cudaStream_t>(stream));
}

int sim_split3_bf16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, sim_stream_t stream) {
  return sim::split3_bf16(x, ld, rows, K, out, ldo, plane, static_cast<cudaStream_t>(stream));
}

int sim_split3_bf16_t(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, sim_stream_t stream) {
  return sim::split3_bf16_t(x, ld, rows, K, out, ldo, plane, static_cast<cudaStream_t>(stream));
}

int sim_split2_f16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, sim_stream_t stream) {
  return sim::split2_f16(x, ld, rows, K, out, ldo, plane, static_cast<cudaStream_t>(stream));
}

int sim_gemm_planes(int np, const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd,
                    int M, int N, int K, int act_mode, int act_col0, const float* act_bias, sim_stream_t stream) {
  return sim::gemm_planes(np, Xs, ldx, xplane, Ws, ldw, wplane, Y, ldd, M, N, K, static_cast<cudaStream_t>(stream), nullptr, 0,
                          0, 0, act_mode, act_col0, act_bias);
}

int sim_gemm_bf16x3(const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd,
                    int M, int N, int K, sim_stream_t stream) {
  return sim::gemm_bf16x3(Xs, ldx, xplane, Ws, ldw, wplane, Y, ldd, M, N, K, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
