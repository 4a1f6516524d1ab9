// 3-nearest-centre inverse-squared-distance interpolation:
//   PointNetFeaturePropagation.forward        part_segmentation/models/pointnet2_utils.py:273-311
//   square_distance                           part_segmentation/models/pointnet2_utils.py:20-38
// The reference builds the full (B, N, S) distance matrix with a matmul, SORTS every row (S = 256 centres for each of
// the 2048 points), keeps three columns, and gathers the three feature rows through an expanded (B, N, S, C) view.
// Here one warp owns a query point: the centres of its cloud are staged once per CTA in shared memory, every lane keeps
// the three best of its S/32 candidates in registers, three rounds of a warp arg-min merge them (ties -> lowest index,
// i.e. a stable sort), and the same warp writes the weighted sum of the three feature rows with 16-byte accesses.
// idx / weight are kept for the backward, which is one vector atomic per (query, neighbour, 4 channels).

#include "kernels.cuh"

namespace sim {

namespace {

constexpr int kQueriesPerCta = 64;  // 8 warps x 8 queries share one staging of the centres
constexpr int kMaxCentres = 2048;   // 32 KB of shared memory

struct Top3 {
  float d0, d1, d2;
  int i0, i1, i2;
};

__device__ __forceinline__ void top3_insert(Top3& t, float d, int i) {
  // strict < keeps the earlier (lower) index on ties: candidates arrive in ascending index order within a lane
  if (d < t.d2) {
    if (d < t.d1) {
      t.d2 = t.d1, t.i2 = t.i1;
      if (d < t.d0) {
        t.d1 = t.d0, t.i1 = t.i0;
        t.d0 = d, t.i0 = i;
      } else {
        t.d1 = d, t.i1 = i;
      }
    } else {
      t.d2 = d, t.i2 = i;
    }
  }
}

// dist = -2 a.b + |a|^2 + |b|^2 in the reference's order of operations (two in-place adds onto the scaled matmul)
__device__ __forceinline__ float sqdist_expanded(float ax, float ay, float az, float sa, float bx, float by, float bz,
                                                 float sb) {
  const float dot = fmaf(az, bz, fmaf(ay, by, __fmul_rn(ax, bx)));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), sa), sb);
}

__global__ void __launch_bounds__(256) three_nn_interp_fwd_kernel(const float* __restrict__ xyz1,
                                                                  const float* __restrict__ xyz2,
                                                                  const float* __restrict__ points2, int N, int S, int C,
                                                                  float* __restrict__ out, int* __restrict__ idx,
                                                                  float* __restrict__ weight) {
  extern __shared__ float sm[];  // [S][4]: x, y, z, |b|^2
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* c2 = xyz2 + (long)b * S * 3;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const float x = c2[3 * j], y = c2[3 * j + 1], z = c2[3 * j + 2];
    sm[4 * j] = x, sm[4 * j + 1] = y, sm[4 * j + 2] = z;
    sm[4 * j + 3] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  }
  __syncthreads();
  const int q_end = min(N, (blockIdx.x + 1) * kQueriesPerCta);
  for (int q = blockIdx.x * kQueriesPerCta + warp; q < q_end; q += 8) {
    const float* a = xyz1 + ((long)b * N + q) * 3;
    const float ax = a[0], ay = a[1], az = a[2];
    const float sa = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
    Top3 t{INFINITY, INFINITY, INFINITY, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int j = lane; j < S; j += 32) {
      const float4 c = *reinterpret_cast<const float4*>(sm + 4 * j);
      top3_insert(t, sqdist_expanded(ax, ay, az, sa, c.x, c.y, c.z, c.w), j);
    }
    float dk[3];
    int ik[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // pop the warp-wide minimum (distance, index) three times
      float d = t.d0;
      int i = t.i0;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (od < d || (od == d && oi < i)) d = od, i = oi;
      }
      dk[k] = d, ik[k] = i;
      if (t.i0 == i) {  // this lane held the winner: advance its list
        t.d0 = t.d1, t.i0 = t.i1;
        t.d1 = t.d2, t.i1 = t.i2;
        t.d2 = INFINITY, t.i2 = 0x7fffffff;
      }
    }
    // weight = (1 / (d + 1e-8)) / sum_k (1 / (d_k + 1e-8))        pointnet2_utils.py:297-299
    const float r0 = __fdiv_rn(1.f, __fadd_rn(dk[0], 1e-8f)), r1 = __fdiv_rn(1.f, __fadd_rn(dk[1], 1e-8f)),
                r2 = __fdiv_rn(1.f, __fadd_rn(dk[2], 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
    const long qo = ((long)b * N + q) * 3;
    if (lane < 3) {
      idx[qo + lane] = lane == 0 ? ik[0] : (lane == 1 ? ik[1] : ik[2]);
      weight[qo + lane] = lane == 0 ? w0 : (lane == 1 ? w1 : w2);
    }
    if (out) {
      const float* p0 = points2 + ((long)b * S + ik[0]) * C;
      const float* p1 = points2 + ((long)b * S + ik[1]) * C;
      const float* p2 = points2 + ((long)b * S + ik[2]) * C;
      float* o = out + ((long)b * N + q) * C;
      if ((C & 3) == 0) {
        for (int c = lane * 4; c < C; c += 128) {
          const float4 v0 = *reinterpret_cast<const float4*>(p0 + c), v1 = *reinterpret_cast<const float4*>(p1 + c),
                       v2 = *reinterpret_cast<const float4*>(p2 + c);
          float4 r;  // (g0 w0 + g1 w1) + g2 w2, products rounded first (the reference materialises gathered * weight)
          r.x = __fadd_rn(__fadd_rn(__fmul_rn(v0.x, w0), __fmul_rn(v1.x, w1)), __fmul_rn(v2.x, w2));
          r.y = __fadd_rn(__fadd_rn(__fmul_rn(v0.y, w0), __fmul_rn(v1.y, w1)), __fmul_rn(v2.y, w2));
          r.z = __fadd_rn(__fadd_rn(__fmul_rn(v0.z, w0), __fmul_rn(v1.z, w1)), __fmul_rn(v2.z, w2));
          r.w = __fadd_rn(__fadd_rn(__fmul_rn(v0.w, w0), __fmul_rn(v1.w, w1)), __fmul_rn(v2.w, w2));
          __stcs(reinterpret_cast<float4*>(o + c), r);  // streamed: 150 MB of output must not evict the feature rows
        }
      } else {
        for (int c = lane; c < C; c += 32)
          o[c] = __fadd_rn(__fadd_rn(__fmul_rn(p0[c], w0), __fmul_rn(p1[c], w1)), __fmul_rn(p2[c], w2));
      }
    }
  }
}

// d points2[b, idx[q,k], :] += weight[q,k] * dout[b, q, :]   (the gradient of the gather-and-weight; xyz carries none)
__global__ void __launch_bounds__(256) three_interp_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx,
                                                               const float* __restrict__ weight, long Q, int N, int S,
                                                               int C, float* __restrict__ dp2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long q = (long)blockIdx.x * 8 + warp;
  if (q >= Q) return;
  const long b = q / N;
  const int i0 = idx[q * 3], i1 = idx[q * 3 + 1], i2 = idx[q * 3 + 2];
  const float w0 = weight[q * 3], w1 = weight[q * 3 + 1], w2 = weight[q * 3 + 2];
  const float* g = dout + q * C;
  float* d0 = dp2 + (b * S + i0) * C;
  float* d1 = dp2 + (b * S + i1) * C;
  float* d2 = dp2 + (b * S + i2) * C;
  if ((C & 3) == 0) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(g + c));
      atomicAdd(reinterpret_cast<float4*>(d0 + c), make_float4(v.x * w0, v.y * w0, v.z * w0, v.w * w0));
      atomicAdd(reinterpret_cast<float4*>(d1 + c), make_float4(v.x * w1, v.y * w1, v.z * w1, v.w * w1));
      atomicAdd(reinterpret_cast<float4*>(d2 + c), make_float4(v.x * w2, v.y * w2, v.z * w2, v.w * w2));
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float v = g[c];
      atomicAdd(d0 + c, v * w0);
      atomicAdd(d1 + c, v * w1);
      atomicAdd(d2 + c, v * w2);
    }
  }
}

}  // namespace

int three_nn_interp_fwd(const float* xyz1, const float* xyz2, const float* points2, int B, int N, int S, int C, float* out,
                        int* idx, float* weight, cudaStream_t stream) {
  SIM_REQUIRE(xyz1 && xyz2 && idx && weight && B > 0 && N > 0, SIM_ERR_INVALID, "three_nn_interp_fwd: null tensor / empty problem");
  SIM_REQUIRE(S >= 3 && S <= kMaxCentres, SIM_ERR_INVALID, "three_nn_interp_fwd: 3 <= S <= %d centres (got %d)", kMaxCentres, S);
  SIM_REQUIRE((out == nullptr) == (points2 == nullptr) && (!out || C > 0), SIM_ERR_INVALID,
              "three_nn_interp_fwd: points2 and out go together");
  SIM_REQUIRE(!out || (C & 3) || (aligned16(points2) && aligned16(out)), SIM_ERR_ALIGN,
              "three_nn_interp_fwd: feature rows must be 16-byte aligned when C %% 4 == 0");
  const dim3 grid((N + kQueriesPerCta - 1) / kQueriesPerCta, B);
  three_nn_interp_fwd_kernel<<<grid, 256, (size_t)S * 16, stream>>>(xyz1, xyz2, points2, N, S, C, out, idx, weight);
  return check_launch("three_nn_interp_fwd");
}

int three_interp_bwd(const float* dout, const int* idx, const float* weight, int B, int N, int S, int C, float* dpoints2,
                     cudaStream_t stream) {
  SIM_REQUIRE(dout && idx && weight && dpoints2 && B > 0 && N > 0 && S > 0 && C > 0, SIM_ERR_INVALID,
              "three_interp_bwd: null tensor / empty problem");
  SIM_REQUIRE((C & 3) || (aligned16(dout) && aligned16(dpoints2)), SIM_ERR_ALIGN,
              "three_interp_bwd: rows must be 16-byte aligned when C %% 4 == 0");
  if (cudaMemsetAsync(dpoints2, 0, (size_t)B * S * C * sizeof(float), stream) != cudaSuccess)
    return check_launch("three_interp_bwd memset");
  const long Q = (long)B * N;
  three_interp_bwd_kernel<<<(int)((Q + 7) / 8), 256, 0, stream>>>(dout, idx, weight, Q, N, S, C, dpoints2);
  return check_launch("three_interp_bwd");
}

}  // namespace sim
