// Chamfer-L2 between small point sets ("next" row f-3): the MAE reconstruction loss
//   chamfer_distance(rebuild, gt, batch_reduction=None)[0]      models/point_mamba.py:2950, 3199-3213
// (pytorch3d: squared L2, point_reduction "mean":  mean_i min_j |x_i - y_j|^2 + mean_j min_i |x_i - y_j|^2 per pair).
// The reference runs two pytorch3d kNN(K=1) kernels plus gathers over B*2k*m = 4864 pairs of 32-point patches; here one
// warp owns a pair, both sets staged in shared memory, and the forward also records the arg-mins so the backward is a
// second single pass (d/dx of a min is the gradient through its arg-min; ties -> lowest index, like torch.min).

#include "kernels.cuh"

namespace sim {

namespace {

constexpr int kMaxPts = 256;  // points per set a warp stages in shared memory

__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) chamfer_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          long R, int P, int Q, float* __restrict__ loss,
                                                          int* __restrict__ idx_x, int* __restrict__ idx_y) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * 8 + warp;
  if (r >= R) return;
  float* sx = sm + (size_t)warp * 3 * (P + Q);
  float* sy = sx + 3 * P;
  for (int i = lane; i < 3 * P; i += 32) sx[i] = x[r * 3 * P + i];
  for (int i = lane; i < 3 * Q; i += 32) sy[i] = y[r * 3 * Q + i];
  __syncwarp();
  float sum_x = 0.f, sum_y = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float ax = sx[3 * i], ay = sx[3 * i + 1], az = sx[3 * i + 2];
    float best = INFINITY;
    int bj = 0;
    for (int j = 0; j < Q; ++j) {
      const float d = sqdist(ax, ay, az, sy[3 * j], sy[3 * j + 1], sy[3 * j + 2]);
      if (d < best) best = d, bj = j;
    }
    sum_x += best;
    if (idx_x) idx_x[r * P + i] = bj;
  }
  for (int j = lane; j < Q; j += 32) {
    const float bx = sy[3 * j], by = sy[3 * j + 1], bz = sy[3 * j + 2];
    float best = INFINITY;
    int bi = 0;
    for (int i = 0; i < P; ++i) {
      const float d = sqdist(sx[3 * i], sx[3 * i + 1], sx[3 * i + 2], bx, by, bz);
      if (d < best) best = d, bi = i;
    }
    sum_y += best;
    if (idx_y) idx_y[r * Q + j] = bi;
  }
  sum_x = warp_sum_f(sum_x);
  sum_y = warp_sum_f(sum_y);
  if (lane == 0) loss[r] = sum_x / (float)P + sum_y / (float)Q;
}

// dx_i = g (2/P) (x_i - y[idx_x[i]]) + g (2/Q) sum_{j : idx_y[j] == i} (x_i - y_j);   dy symmetric (optional)
__global__ void __launch_bounds__(256) chamfer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          const int* __restrict__ idx_x, const int* __restrict__ idx_y,
                                                          const float* __restrict__ gloss, long R, int P, int Q,
                                                          float* __restrict__ dx, float* __restrict__ dy) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * 8 + warp;
  if (r >= R) return;
  float* sx = sm + (size_t)warp * 4 * (P + Q);
  float* sy = sx + 3 * P;
  int* ix = reinterpret_cast<int*>(sy + 3 * Q);
  int* iy = ix + P;
  for (int i = lane; i < 3 * P; i += 32) sx[i] = x[r * 3 * P + i];
  for (int i = lane; i < 3 * Q; i += 32) sy[i] = y[r * 3 * Q + i];
  for (int i = lane; i < P; i += 32) ix[i] = idx_x[r * P + i];
  for (int j = lane; j < Q; j += 32) iy[j] = idx_y[r * Q + j];
  __syncwarp();
  const float g = gloss[r];
  const float wp = 2.f * g / (float)P, wq = 2.f * g / (float)Q;
  if (dx) {
    for (int i = lane; i < P; i += 32) {
      const float ax = sx[3 * i], ay = sx[3 * i + 1], az = sx[3 * i + 2];
      const int j0 = ix[i];
      float gx = wp * (ax - sy[3 * j0]), gy = wp * (ay - sy[3 * j0 + 1]), gz = wp * (az - sy[3 * j0 + 2]);
      for (int j = 0; j < Q; ++j)
        if (iy[j] == i) gx += wq * (ax - sy[3 * j]), gy += wq * (ay - sy[3 * j + 1]), gz += wq * (az - sy[3 * j + 2]);
      dx[(r * P + i) * 3] = gx, dx[(r * P + i) * 3 + 1] = gy, dx[(r * P + i) * 3 + 2] = gz;
    }
  }
  if (dy) {
    for (int j = lane; j < Q; j += 32) {
      const float bx = sy[3 * j], by = sy[3 * j + 1], bz = sy[3 * j + 2];
      const int i0 = iy[j];
      float gx = wq * (bx - sx[3 * i0]), gy = wq * (by - sx[3 * i0 + 1]), gz = wq * (bz - sx[3 * i0 + 2]);
      for (int i = 0; i < P; ++i)
        if (ix[i] == j) gx += wp * (bx - sx[3 * i]), gy += wp * (by - sx[3 * i + 1]), gz += wp * (bz - sx[3 * i + 2]);
      dy[(r * Q + j) * 3] = gx, dy[(r * Q + j) * 3 + 1] = gy, dy[(r * Q + j) * 3 + 2] = gz;
    }
  }
}

}  // namespace

int chamfer_l2_fwd(const float* x, const float* y, long R, int P, int Q, float* loss, int* idx_x, int* idx_y,
                   cudaStream_t stream) {
  SIM_REQUIRE(x && y && loss && R > 0 && P > 0 && Q > 0, SIM_ERR_INVALID, "chamfer_l2_fwd: null tensor / empty problem");
  SIM_REQUIRE(P <= kMaxPts && Q <= kMaxPts, SIM_ERR_INVALID, "chamfer_l2_fwd: at most %d points per set (got %d, %d)",
              kMaxPts, P, Q);
  const size_t smem = (size_t)8 * 3 * (P + Q) * sizeof(float);
  chamfer_fwd_kernel<<<(int)((R + 7) / 8), 256, smem, stream>>>(x, y, R, P, Q, loss, idx_x, idx_y);
  return check_launch("chamfer_l2_fwd");
}

int chamfer_l2_bwd(const float* x, const float* y, const int* idx_x, const int* idx_y, const float* gloss, long R, int P,
                   int Q, float* dx, float* dy, cudaStream_t stream) {
  SIM_REQUIRE(x && y && idx_x && idx_y && gloss && (dx || dy) && R > 0 && P > 0 && Q > 0, SIM_ERR_INVALID,
              "chamfer_l2_bwd: null tensor / empty problem");
  SIM_REQUIRE(P <= kMaxPts && Q <= kMaxPts, SIM_ERR_INVALID, "chamfer_l2_bwd: at most %d points per set", kMaxPts);
  const size_t smem = (size_t)8 * 4 * (P + Q) * sizeof(float);
  chamfer_bwd_kernel<<<(int)((R + 7) / 8), 256, smem, stream>>>(x, y, idx_x, idx_y, gloss, R, P, Q, dx, dy);
  return check_launch("chamfer_l2_bwd");
}

}  // namespace sim
