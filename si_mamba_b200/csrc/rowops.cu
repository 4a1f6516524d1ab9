// Row-granular HBM-bound kernels: residual add + LayerNorm, spectral order
// gather (+ reverse), general row gather (HLT / MAE restore), token mean.
//
// Reference rows:
//   a-9   Block.forward add + LayerNorm           models/block.py:56-58
//   a-13  MixerModel.forward tokens + pos, norm_f models/point_mamba.py:250,256-258
//   a-5/6 sort_points_by_fiedler gather + cat + flip  models/point_mamba.py:817-826, 889-898, 982-989
//   a-8   HLT layout                              part_segmentation/models/pt_mamba.py:670-723
//   a-17  MAE token restore                       models/point_mamba.py:3147-3197
// One warp moves one row with 16-byte vector accesses; rows are C <= 1024*... floats.

#include <algorithm>

#include "kernels.cuh"

namespace sim {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ----------------------------------------------------------------------------- add + LayerNorm
// res_out = x (+ res_in) [(+ x2)];  y = LN(res_out) * gamma + beta.  Statistics in fp32, two-pass in registers.
// MAXV = float4 vectors per lane (C <= 128 * MAXV).
template <typename TX, typename TY, int MAXV>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const TX* __restrict__ x, const TX* __restrict__ x2,
                                                            const float* __restrict__ res_in,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            float* __restrict__ res_out, TY* __restrict__ y,
                                                            long rows, int C, float eps,
                                                            __nv_bfloat16* __restrict__ planes, long plane,
                                                            const float* __restrict__ row_scale, int rows_per_sample,
                                                            int plane_fmt) {
  const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = C / 4;
  // DropPath folded in (models/block.py:59, timm drop_path): x is scaled by mask_b / keep of its sample before the add
  const float sc = row_scale ? row_scale[row / rows_per_sample] : 1.f;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < nv) {
      float4 a = ld4<TX>(x + row * C + 4 * q);
      if (row_scale) a.x *= sc, a.y *= sc, a.z *= sc, a.w *= sc;
      if (x2) {
        const float4 c = ld4<TX>(x2 + row * C + 4 * q);
        a.x += c.x, a.y += c.y, a.z += c.z, a.w += c.w;
      }
      if (res_in) {
        const float4 r = *reinterpret_cast<const float4*>(res_in + row * C + 4 * q);
        a.x += r.x, a.y += r.y, a.z += r.z, a.w += r.w;
      }
      v[i] = a;
      s += (a.x + a.y) + (a.z + a.w);
      if (res_out) *reinterpret_cast<float4*>(res_out + row * C + 4 * q) = a;
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < nv) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < nv) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * q);
      const float4 bb = *reinterpret_cast<const float4*>(beta + 4 * q);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bb.x;
      o.y = (v[i].y - mean) * rstd * g.y + bb.y;
      o.z = (v[i].z - mean) * rstd * g.z + bb.z;
      o.w = (v[i].w - mean) * rstd * g.w + bb.w;
      if (y) st4<TY>(y + row * C + 4 * q, o);
      if (planes) {
        if (plane_fmt) split2h_store4(reinterpret_cast<__half*>(planes) + row * C + 4 * q, plane, o);
        else split3_store4(planes + row * C + 4 * q, plane, o);
      }
    }
  }
}

int add_layernorm(const void* x, const void* x2, const float* res_in, const float* gamma, const float* beta,
                  float* res_out, void* y, long rows, int C, float eps, int dtype_x, int dtype_y,
                  cudaStream_t stream, void* planes, long plane, const float* row_scale, int rows_per_sample,
                  int plane_fmt) {
  SIM_REQUIRE(!row_scale || rows_per_sample > 0, SIM_ERR_INVALID, "add_layernorm: row_scale needs rows_per_sample > 0");
  SIM_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 1024, SIM_ERR_INVALID,
              "add_layernorm: C must be a multiple of 4 and <= 1024 (got %d)", C);
  SIM_REQUIRE(x && gamma && beta && (y || planes), SIM_ERR_INVALID, "add_layernorm: null tensor");
  SIM_REQUIRE((!planes || ((reinterpret_cast<uintptr_t>(planes) & 7u) == 0 && plane % 4 == 0)), SIM_ERR_ALIGN,
              "add_layernorm: split planes need 8-byte alignment");
  SIM_REQUIRE(aligned16(x) && (!y || aligned16(y)) && aligned16(gamma) && aligned16(beta) && (!x2 || aligned16(x2)) &&
                  (!res_in || aligned16(res_in)) && (!res_out || aligned16(res_out)),
              SIM_ERR_ALIGN, "add_layernorm: tensors must be 16-byte aligned");
  const int grid = (int)((rows + 7) / 8);
#define SIM_LN_LAUNCH(TX, TY, MAXV)                                                                          \
  add_layernorm_kernel<TX, TY, MAXV><<<grid, 256, 0, stream>>>(static_cast<const TX*>(x),                    \
                                                               static_cast<const TX*>(x2), res_in, gamma,    \
                                                               beta, res_out, static_cast<TY*>(y), rows, C, eps,  \
                                                               static_cast<__nv_bfloat16*>(planes), plane, row_scale, \
                                                               rows_per_sample, plane_fmt)
#define SIM_LN_MAXV(TX, TY)                     \
  if (C <= 384) {                               \
    SIM_LN_LAUNCH(TX, TY, 3);                   \
  } else if (C <= 512) {                        \
    SIM_LN_LAUNCH(TX, TY, 4);                   \
  } else {                                      \
    SIM_LN_LAUNCH(TX, TY, 8);                   \
  }
  if (dtype_x == 0 && dtype_y == 0) {
    SIM_LN_MAXV(float, float)
  } else if (dtype_x == 0 && dtype_y == 1) {
    SIM_LN_MAXV(float, __nv_bfloat16)
  } else if (dtype_x == 1 && dtype_y == 1) {
    SIM_LN_MAXV(__nv_bfloat16, __nv_bfloat16)
  } else if (dtype_x == 1 && dtype_y == 0) {
    SIM_LN_MAXV(__nv_bfloat16, float)
  } else {
    set_error("add_layernorm: bad dtype codes %d/%d", dtype_x, dtype_y);
    return SIM_ERR_INVALID;
  }
#undef SIM_LN_MAXV
#undef SIM_LN_LAUNCH
  return check_launch("add_layernorm");
}

// ----------------------------------------------------------------------------- add + LayerNorm backward
// res = x + residual; y = LN(res) gamma + beta  (Block.forward, models/block.py:56-58, training configs C2-C4).
// Given dy and the gradient dres_out that later layers send into the residual stream:
//   dres[r,:] = dres_out[r,:] + rstd (g - mean(g) - xhat mean(g xhat)),  g = dy gamma, xhat = (res - mean) rstd
// (the gradient of BOTH x and residual), dgamma += sum_r dy xhat, dbeta += sum_r dy.  Statistics are recomputed from
// the saved fp32 residual stream (one warp per row, row in registers), so the forward saves nothing extra.  dgamma /
// dbeta: per-thread partials over the warp's rows -> shared-memory reduce over the CTA -> one fp32 atomic per column.
// (torch's native_layer_norm_backward spends 290 us per call in its gamma/beta kernel at the C2 shape: 27 % of the step.)
template <typename TY, int MAXV>
__global__ void __launch_bounds__(256) add_layernorm_bwd_kernel(const float* __restrict__ res, const TY* __restrict__ dy,
                                                                const float* __restrict__ dres_out,
                                                                const float* __restrict__ gamma, float* __restrict__ dres,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                long rows, int C, float eps, void* __restrict__ dx,
                                                                int dtype_dx, const float* __restrict__ row_scale,
                                                                int rows_per_sample) {
  extern __shared__ float s_red[];  // 2 * C
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = C / 4;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float4 ag[MAXV], ab[MAXV], gm[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int q = lane + 32 * i;
    gm[i] = q < nv ? *reinterpret_cast<const float4*>(gamma + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long row = (long)blockIdx.x * 8 + warp; row < rows; row += (long)gridDim.x * 8) {
    float4 v[MAXV], d[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q < nv) {
        v[i] = *reinterpret_cast<const float4*>(res + row * C + 4 * q);
        d[i] = ld4<TY>(dy + row * C + 4 * q);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      } else {
        v[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q < nv) {
        v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
        ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      v[i].x *= rstd, v[i].y *= rstd, v[i].z *= rstd, v[i].w *= rstd;  // xhat
      ab[i].x += d[i].x, ab[i].y += d[i].y, ab[i].z += d[i].z, ab[i].w += d[i].w;
      ag[i].x += d[i].x * v[i].x, ag[i].y += d[i].y * v[i].y, ag[i].z += d[i].z * v[i].z, ag[i].w += d[i].w * v[i].w;
      d[i].x *= gm[i].x, d[i].y *= gm[i].y, d[i].z *= gm[i].z, d[i].w *= gm[i].w;  // g
      m1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      m2 += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    m1 = warp_sum(m1) / (float)C;
    m2 = warp_sum(m2) / (float)C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q < nv) {
        float4 o;
        o.x = rstd * (d[i].x - m1 - v[i].x * m2), o.y = rstd * (d[i].y - m1 - v[i].y * m2);
        o.z = rstd * (d[i].z - m1 - v[i].z * m2), o.w = rstd * (d[i].w - m1 - v[i].w * m2);
        if (dres_out) {
          const float4 r = *reinterpret_cast<const float4*>(dres_out + row * C + 4 * q);
          o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
        }
        *reinterpret_cast<float4*>(dres + row * C + 4 * q) = o;
        if (dx) {  // gradient of the x operand in its own dtype, with the DropPath scale of its sample
          const float sc = row_scale ? row_scale[row / rows_per_sample] : 1.f;
          o.x *= sc, o.y *= sc, o.z *= sc, o.w *= sc;
          if (dtype_dx == 0) st4<float>(static_cast<float*>(dx) + row * C + 4 * q, o);
          else st4<__nv_bfloat16>(static_cast<__nv_bfloat16*>(dx) + row * C + 4 * q, o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < nv) {
      atomicAdd(s_red + 4 * q + 0, ag[i].x), atomicAdd(s_red + 4 * q + 1, ag[i].y);
      atomicAdd(s_red + 4 * q + 2, ag[i].z), atomicAdd(s_red + 4 * q + 3, ag[i].w);
      atomicAdd(s_red + C + 4 * q + 0, ab[i].x), atomicAdd(s_red + C + 4 * q + 1, ab[i].y);
      atomicAdd(s_red + C + 4 * q + 2, ab[i].z), atomicAdd(s_red + C + 4 * q + 3, ab[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, s_red[i]);
    atomicAdd(dbeta + i, s_red[C + i]);
  }
}

int add_layernorm_bwd(const float* res, const void* dy, const float* dres_out, const float* gamma, float* dres,
                      float* dgamma, float* dbeta, long rows, int C, float eps, int dtype_y, cudaStream_t stream, void* dx,
                      int dtype_dx, const float* row_scale, int rows_per_sample) {
  SIM_REQUIRE((!row_scale || (dx && rows_per_sample > 0)) && (!dx || (aligned16(dx) && (dtype_dx == 0 || dtype_dx == 1))),
              SIM_ERR_INVALID, "add_layernorm_bwd: row_scale needs dx and rows_per_sample; dx must be aligned fp32 / bf16");
  SIM_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 1024, SIM_ERR_INVALID,
              "add_layernorm_bwd: C must be a multiple of 4 and <= 1024 (got %d)", C);
  SIM_REQUIRE(res && dy && gamma && dres && dgamma && dbeta, SIM_ERR_INVALID, "add_layernorm_bwd: null tensor");
  SIM_REQUIRE(aligned16(res) && aligned16(dy) && aligned16(gamma) && aligned16(dres) && (!dres_out || aligned16(dres_out)),
              SIM_ERR_ALIGN, "add_layernorm_bwd: tensors must be 16-byte aligned");
  const long want = (rows + 7) / 8;
  const int grid = (int)(want < 148L * 8 ? want : 148L * 8);
  const size_t smem = (size_t)2 * C * sizeof(float);
#define SIM_LNB_LAUNCH(TY, MAXV)                                                                                  \
  add_layernorm_bwd_kernel<TY, MAXV><<<grid, 256, smem, stream>>>(res, static_cast<const TY*>(dy), dres_out, gamma, \
                                                                  dres, dgamma, dbeta, rows, C, eps, dx, dtype_dx,   \
                                                                  row_scale, rows_per_sample)
#define SIM_LNB_MAXV(TY)      \
  if (C <= 384) {             \
    SIM_LNB_LAUNCH(TY, 3);    \
  } else if (C <= 512) {      \
    SIM_LNB_LAUNCH(TY, 4);    \
  } else {                    \
    SIM_LNB_LAUNCH(TY, 8);    \
  }
  if (dtype_y == 0) {
    SIM_LNB_MAXV(float)
  } else if (dtype_y == 1) {
    SIM_LNB_MAXV(__nv_bfloat16)
  } else {
    set_error("add_layernorm_bwd: bad dtype code %d", dtype_y);
    return SIM_ERR_INVALID;
  }
#undef SIM_LNB_MAXV
#undef SIM_LNB_LAUNCH
  return check_launch("add_layernorm_bwd");
}

// ----------------------------------------------------------------------------- Encoder / head glue (a-2, a-14)
// The per-patch PointNet (Encoder.forward, models/point_mamba.py:59-73) stays on library GEMMs; these are the
// row-granular passes between them, each one read + one write instead of ATen's add -> relu and generic reductions.
//   group_max:        out[g, c] = max_{m < M} x[g*M + m, c]                     (torch.max(feature, dim=2))
//   group_bias_relu:  x[p, c] = relu(x[p, c] + gvec[p / M, c])   in place         (cat([global, local]) conv, split)
//   layernorm_mean:   out[b, c] = mean_t LN(x[b, t, :])[c]                       (self.norm(x).mean(1), :1122-1123)
template <typename T>
__global__ void __launch_bounds__(256) group_max_kernel(const T* __restrict__ x, T* __restrict__ out, long groups, int M,
                                                        int C) {
  const int nv = C / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * nv) return;
  const long g = i / nv;
  const int q = (int)(i % nv);
  const T* p = x + (g * M) * C + 4 * q;
  float4 m = ld4<T>(p);
  for (int r = 1; r < M; ++r) {
    const float4 v = ld4<T>(p + (long)r * C);
    m.x = fmaxf(m.x, v.x), m.y = fmaxf(m.y, v.y), m.z = fmaxf(m.z, v.z), m.w = fmaxf(m.w, v.w);
  }
  st4<T>(out + g * C + 4 * q, m);
}

template <typename T>
__global__ void __launch_bounds__(256) group_bias_relu_kernel(T* __restrict__ x, const T* __restrict__ gvec, long rows,
                                                              int M, int C) {
  const int nv = C / 4;
  const long n = rows * nv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nv;
    const int q = (int)(i % nv);
    float4 v = ld4<T>(x + r * C + 4 * q);
    const float4 b = ld4<T>(gvec + (r / M) * C + 4 * q);
    v.x = fmaxf(v.x + b.x, 0.f), v.y = fmaxf(v.y + b.y, 0.f), v.z = fmaxf(v.z + b.z, 0.f), v.w = fmaxf(v.w + b.w, 0.f);
    st4<T>(x + r * C + 4 * q, v);
  }
}

// one warp per row as in add_layernorm; a CTA (8 warps) owns 8-row strides of ONE batch element's tokens and adds its
// rows' normalised values into that element's mean with one shared-memory reduce + one atomic per column
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_mean_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ out,
                                                             int L, int C, float eps, int ctas_per_batch) {
  extern __shared__ float s_acc[];  // C
  const int b = blockIdx.x / ctas_per_batch, part = blockIdx.x % ctas_per_batch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = C / 4;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float4 acc[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = part * 8 + warp; t < L; t += ctas_per_batch * 8) {
    const float* row = x + ((long)b * L + t) * C;
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      v[i] = q < nv ? *reinterpret_cast<const float4*>(row + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q < nv) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      acc[i].x += (v[i].x - mean) * rstd, acc[i].y += (v[i].y - mean) * rstd;
      acc[i].z += (v[i].z - mean) * rstd, acc[i].w += (v[i].w - mean) * rstd;
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < nv) {
      atomicAdd(s_acc + 4 * q, acc[i].x), atomicAdd(s_acc + 4 * q + 1, acc[i].y);
      atomicAdd(s_acc + 4 * q + 2, acc[i].z), atomicAdd(s_acc + 4 * q + 3, acc[i].w);
    }
  }
  __syncthreads();
  // mean_t (xhat gamma + beta) = gamma * mean_t(xhat) + beta: the affine part is applied once per column
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    atomicAdd(out + (long)b * C + i, s_acc[i] * gamma[i] / (float)L + (part == 0 ? beta[i] : 0.f));
}

int group_max(const void* x, void* out, long groups, int M, int C, int dtype, cudaStream_t stream) {
  SIM_REQUIRE(x && out && groups > 0 && M > 0 && C > 0 && C % 4 == 0, SIM_ERR_INVALID, "group_max: bad arguments");
  SIM_REQUIRE(aligned16(x) && aligned16(out), SIM_ERR_ALIGN, "group_max: tensors must be 16-byte aligned");
  const long n = groups * (C / 4);
  const int grid = (int)((n + 255) / 256);
  if (dtype == 0)
    group_max_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(out), groups, M, C);
  else
    group_max_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                              static_cast<__nv_bfloat16*>(out), groups, M, C);
  return check_launch("group_max");
}

// y[p, c] = act(b[c] + w[c,0] x[p,0] + w[c,1] x[p,1] + w[c,2] x[p,2]): the 3 -> C first layer of the Encoder (Conv1d(3, 128, 1)
// with eval-mode BatchNorm folded in, + ReLU; models/point_mamba.py:47-49) and of pos_embed (Linear(3, 128) + GELU, :470-474).
// K = 3 is no GEMM: a thread computes four adjacent channels of one point with three FMAs each; the output write is the cost.
template <int ACT>
__global__ void __launch_bounds__(256) point_linear3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, float* __restrict__ y, long rows,
                                                            int C) {
  extern __shared__ float s_w[];  // [C][4]: w0, w1, w2, bias
  for (int i = threadIdx.x; i < C; i += 256) {
    s_w[4 * i] = w[3 * i], s_w[4 * i + 1] = w[3 * i + 1], s_w[4 * i + 2] = w[3 * i + 2];
    s_w[4 * i + 3] = b ? b[i] : 0.f;
  }
  __syncthreads();
  const int cq = C / 4;
  for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < rows * cq; e += (long)gridDim.x * 256) {
    const long p = e / cq;
    const int c = (int)(e % cq) * 4;
    const float px = x[3 * p], py = x[3 * p + 1], pz = x[3 * p + 2];
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 wv = *reinterpret_cast<const float4*>(s_w + 4 * (c + i));
      float v = fmaf(wv.z, pz, fmaf(wv.y, py, fmaf(wv.x, px, wv.w)));
      if (ACT == 1) v = fmaxf(v, 0.f);
      if (ACT == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // nn.GELU() (erf form)
      o[i] = v;
    }
    *reinterpret_cast<float4*>(y + p * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int point_linear3(const float* x, const float* w, const float* b, float* y, long rows, int C, int act, cudaStream_t stream) {
  SIM_REQUIRE(x && w && y && rows > 0 && C > 0 && C % 4 == 0 && C <= 2048 && act >= 0 && act <= 2, SIM_ERR_INVALID,
              "point_linear3: bad arguments");
  SIM_REQUIRE(aligned16(y), SIM_ERR_ALIGN, "point_linear3: the output must be 16-byte aligned");
  const long work = rows * (C / 4);
  const int grid = (int)std::min<long>((work + 255) / 256, 148L * 8);
  const size_t smem = (size_t)C * 16;
  if (act == 0) point_linear3_kernel<0><<<grid, 256, smem, stream>>>(x, w, b, y, rows, C);
  else if (act == 1) point_linear3_kernel<1><<<grid, 256, smem, stream>>>(x, w, b, y, rows, C);
  else point_linear3_kernel<2><<<grid, 256, smem, stream>>>(x, w, b, y, rows, C);
  return check_launch("point_linear3");
}

// ----------------------------------------------------------------------------- flat AdamW (trainer plumbing)
// torch.optim.AdamW (tools/builder.py:74, part_segmentation/main.py:201) over ONE flat parameter / gradient / moment buffer:
//   g = grad * grad_scale (the clip_grad_norm_ coefficient, optional);  p *= 1 - lr wd;  m = b1 m + (1 - b1) g;
//   v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),  t = *step + 1.
// wd[i] < 0 marks slots the optimizer must not touch (padding, parameters without a gradient).  lr, step and grad_scale
// are device scalars, so the launch can live in a CUDA graph; torch's capturable multi-tensor AdamW issues ~700 scalar
// kernels per step for the same update (1.5 - 2.5 ms of a 9 - 22 ms step, profiles/r02_launches_c4.md).
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         const float* __restrict__ wd, long n, const float* __restrict__ lr_p,
                                                         const float* __restrict__ step_p, const float* __restrict__ gs_p,
                                                         float b1, float b2, float eps) {
  const float lr = *lr_p, t = *step_p + 1.f;
  const float gs = gs_p ? *gs_p : 1.f;
  const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (long i = ((long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long)gridDim.x * 256 * 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(wd + i);
    float4 p4 = *reinterpret_cast<float4*>(p + i);
    const float4 g4 = *reinterpret_cast<const float4*>(g + i);
    float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
    float* pp = &p4.x;
    float* mm = &m4.x;
    float* vv = &v4.x;
    const float* gg = &g4.x;
    const float* ww = &w4.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (ww[k] < 0.f) continue;
      const float gk = gg[k] * gs;
      float pk = pp[k] * (1.f - lr * ww[k]);
      const float mk = b1 * mm[k] + (1.f - b1) * gk;
      const float vk = b2 * vv[k] + (1.f - b2) * gk * gk;
      pk -= step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
      pp[k] = pk, mm[k] = mk, vv[k] = vk;
    }
    *reinterpret_cast<float4*>(p + i) = p4;
    *reinterpret_cast<float4*>(m + i) = m4;
    *reinterpret_cast<float4*>(v + i) = v4;
  }
}
__global__ void adamw_step_incr_kernel(float* step) { *step += 1.f; }

int adamw_flat(float* p, const float* g, float* m, float* v, const float* wd, long n, const float* lr, float* step,
               const float* grad_scale, float beta1, float beta2, float eps, cudaStream_t stream) {
  SIM_REQUIRE(p && g && m && v && wd && lr && step && n > 0 && n % 4 == 0, SIM_ERR_INVALID, "adamw_flat: bad arguments (n %% 4)");
  SIM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && aligned16(wd), SIM_ERR_ALIGN,
              "adamw_flat: buffers must be 16-byte aligned");
  const int grid = (int)std::min<long>((n / 4 + 255) / 256, 148L * 8);
  adamw_flat_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, wd, n, lr, step, grad_scale, beta1, beta2, eps);
  adamw_step_incr_kernel<<<1, 1, 0, stream>>>(step);
  return check_launch("adamw_flat");
}

int group_bias_relu(void* x, const void* gvec, long rows, int M, int C, int dtype, cudaStream_t stream) {
  SIM_REQUIRE(x && gvec && rows > 0 && M > 0 && C > 0 && C % 4 == 0 && rows % M == 0, SIM_ERR_INVALID,
              "group_bias_relu: bad arguments");
  SIM_REQUIRE(aligned16(x) && aligned16(gvec), SIM_ERR_ALIGN, "group_bias_relu: tensors must be 16-byte aligned");
  const long n = rows * (C / 4);
  const int grid = (int)((n + 255) / 256 < 148L * 32 ? (n + 255) / 256 : 148L * 32);
  if (dtype == 0)
    group_bias_relu_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(x), static_cast<const float*>(gvec), rows, M, C);
  else
    group_bias_relu_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x),
                                                                    static_cast<const __nv_bfloat16*>(gvec), rows, M, C);
  return check_launch("group_bias_relu");
}

int layernorm_mean(const float* x, const float* gamma, const float* beta, float* out, int B, int L, int C, float eps,
                   cudaStream_t stream) {
  SIM_REQUIRE(x && gamma && beta && out && B > 0 && L > 0 && C > 0 && C % 4 == 0 && C <= 1024, SIM_ERR_INVALID,
              "layernorm_mean: bad arguments (C must be a multiple of 4, <= 1024)");
  SIM_REQUIRE(aligned16(x), SIM_ERR_ALIGN, "layernorm_mean: x must be 16-byte aligned");
  int per = (L + 7) / 8;
  const int want = (148 * 4 + B - 1) / B;  // enough CTAs to fill the machine at small batch
  if (per > want) per = want;
  if (per < 1) per = 1;
  const size_t smem = (size_t)C * sizeof(float);
  if (C <= 384)
    layernorm_mean_kernel<3><<<B * per, 256, smem, stream>>>(x, gamma, beta, out, L, C, eps, per);
  else if (C <= 512)
    layernorm_mean_kernel<4><<<B * per, 256, smem, stream>>>(x, gamma, beta, out, L, C, eps, per);
  else
    layernorm_mean_kernel<8><<<B * per, 256, smem, stream>>>(x, gamma, beta, out, L, C, eps, per);
  return check_launch("layernorm_mean");
}

// ----------------------------------------------------------------------------- classifier head (a-14 tail)
// y = W3 relu(W2 relu(W1 x + b1) + b2) + b3 for a handful of rows: cls_head_finetune in eval mode
// (models/point_mamba.py:1124-1130: Linear-BN-ReLU-Dropout x2 + Linear; BatchNorm folded into W / b by the caller,
// Dropout is the identity).  The reference's 32-row GEMMs are pure launch + latency (11 small kernels, ~80 us cold); here
// one CTA per row keeps the activations in shared memory, weights are pre-transposed (in, out) so a warp reads 128
// contiguous bytes per k, and the contraction is split over kMlpParts thread groups to shorten the dependent FMA chain.
constexpr int kMlpWidth = 256, kMlpParts = 4;

__device__ __forceinline__ void mlp_layer(const float* __restrict__ in, int din, const float* __restrict__ wt,
                                          const float* __restrict__ b, int dout, float* __restrict__ part,
                                          float* __restrict__ out_s, float* __restrict__ out_g, bool relu) {
  const int j = threadIdx.x % kMlpWidth, pz = threadIdx.x / kMlpWidth;
  const int k0 = (int)((long)din * pz / kMlpParts), k1 = (int)((long)din * (pz + 1) / kMlpParts);
  float acc = 0.f;
  if (j < dout) {
#pragma unroll 8
    for (int k = k0; k < k1; ++k) acc = fmaf(in[k], __ldg(wt + (long)k * dout + j), acc);
  }
  part[pz * kMlpWidth + j] = acc;
  __syncthreads();
  if (pz == 0 && j < dout) {
    float v = b ? b[j] : 0.f;
#pragma unroll
    for (int q = 0; q < kMlpParts; ++q) v += part[q * kMlpWidth + j];
    if (relu) v = fmaxf(v, 0.f);
    if (out_g) out_g[j] = v;
    else out_s[j] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kMlpWidth * kMlpParts) mlp3_relu_rows_kernel(
    const float* __restrict__ x, long ldx, int d0, const float* __restrict__ w1t, const float* __restrict__ b1, int d1,
    const float* __restrict__ w2t, const float* __restrict__ b2, int d2, const float* __restrict__ w3t,
    const float* __restrict__ b3, int d3, float* __restrict__ y, long ldy) {
  extern __shared__ float sm[];
  float* s0 = sm;                        // [d0]
  float* s1 = s0 + d0;                   // [kMlpWidth]
  float* s2 = s1 + kMlpWidth;            // [kMlpWidth]
  float* part = s2 + kMlpWidth;          // [kMlpParts][kMlpWidth]
  const long row = blockIdx.x;
  for (int k = threadIdx.x; k < d0; k += blockDim.x) s0[k] = x[row * ldx + k];
  __syncthreads();
  mlp_layer(s0, d0, w1t, b1, d1, part, s1, nullptr, true);
  mlp_layer(s1, d1, w2t, b2, d2, part, s2, nullptr, true);
  mlp_layer(s2, d2, w3t, b3, d3, part, nullptr, y + row * ldy, false);
}

int mlp3_relu_rows(const float* x, long ldx, long rows, int d0, const float* w1t, const float* b1, int d1, const float* w2t,
                   const float* b2, int d2, const float* w3t, const float* b3, int d3, float* y, long ldy,
                   cudaStream_t stream) {
  SIM_REQUIRE(x && w1t && w2t && w3t && y && rows > 0 && d0 > 0 && d1 > 0 && d2 > 0 && d3 > 0, SIM_ERR_INVALID,
              "mlp3_relu_rows: null tensor / empty problem");
  SIM_REQUIRE(d1 <= kMlpWidth && d2 <= kMlpWidth && d3 <= kMlpWidth && d0 <= 8192 && rows <= 65535, SIM_ERR_INVALID,
              "mlp3_relu_rows: built for a classifier head (layer widths <= %d, <= 65535 rows)", kMlpWidth);
  const size_t smem = ((size_t)d0 + (2 + kMlpParts) * kMlpWidth) * sizeof(float);
  mlp3_relu_rows_kernel<<<(int)rows, kMlpWidth * kMlpParts, smem, stream>>>(x, ldx, d0, w1t, b1, d1, w2t, b2, d2, w3t, b3, d3,
                                                                           y, ldy);
  return check_launch("mlp3_relu_rows");
}

// ----------------------------------------------------------------------------- SAST order gather
// out[b, s*G + r, :] = x[b, perm[b,s,r], :] (+ x2[...]) and, if reverse, the mirrored row
// out[b, 2kG-1-(s*G+r), :] gets the same data: each source row is read once and written twice.
template <typename T>
__global__ void __launch_bounds__(256) order_gather_kernel(const T* __restrict__ x, const T* __restrict__ x2,
                                                           const int* __restrict__ perm, T* __restrict__ o1,
                                                           T* __restrict__ o2, int B, int G, int k, int C,
                                                           int reverse) {
  const long w = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const long total = (long)B * k * G;
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const int b = w / ((long)k * G);
  const int t = w % ((long)k * G);
  const int src = perm[w];
  const long T_out = (long)(reverse ? 2 : 1) * k * G;
  const T* xr = x + ((long)b * G + src) * C;
  const T* x2r = x2 ? x2 + ((long)b * G + src) * C : nullptr;
  T* d1 = o1 + ((long)b * T_out + t) * C;
  T* d1m = o1 + ((long)b * T_out + (T_out - 1 - t)) * C;
  T* d2 = o2 ? o2 + ((long)b * T_out + t) * C : nullptr;
  T* d2m = o2 ? o2 + ((long)b * T_out + (T_out - 1 - t)) * C : nullptr;
  for (int q = lane; q < C / 4; q += 32) {
    float4 a = ld4<T>(xr + 4 * q);
    if (x2r) {
      const float4 c = ld4<T>(x2r + 4 * q);
      if (d2) {
        st4<T>(d2 + 4 * q, c);
        if (reverse) st4<T>(d2m + 4 * q, c);
      } else {
        a.x += c.x, a.y += c.y, a.z += c.z, a.w += c.w;
      }
    }
    st4<T>(d1 + 4 * q, a);
    if (reverse) st4<T>(d1m + 4 * q, a);
  }
}

int order_gather_fwd(const void* x, const void* x2, const int* perm, void* o1, void* o2, int B, int G, int k, int C,
                     int reverse, int dtype, cudaStream_t stream) {
  SIM_REQUIRE(B > 0 && G > 0 && k > 0 && C > 0 && C % 4 == 0, SIM_ERR_INVALID,
              "order_gather_fwd: C must be a multiple of 4");
  SIM_REQUIRE(x && perm && o1 && (!o2 || x2), SIM_ERR_INVALID, "order_gather_fwd: null tensor / o2 without x2");
  SIM_REQUIRE(aligned16(x) && aligned16(o1) && (!x2 || aligned16(x2)) && (!o2 || aligned16(o2)), SIM_ERR_ALIGN,
              "order_gather_fwd: tensors must be 16-byte aligned");
  const long rows = (long)B * k * G;
  const int grid = (int)((rows + 7) / 8);
  if (dtype == 0)
    order_gather_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<const float*>(x2),
                                                         perm, static_cast<float*>(o1), static_cast<float*>(o2), B,
                                                         G, k, C, reverse);
  else if (dtype == 1)
    order_gather_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(x2), perm,
        static_cast<__nv_bfloat16*>(o1), static_cast<__nv_bfloat16*>(o2), B, G, k, C, reverse);
  else {
    set_error("order_gather_fwd: bad dtype %d", dtype);
    return SIM_ERR_INVALID;
  }
  return check_launch("order_gather_fwd");
}

// ----------------------------------------------------------------------------- general row gather
// out[b, t, :] = src_idx[b,t] >= 0 ? x[b, src_idx[b,t], :] : fill (zeros when fill == nullptr).
// Serves the HLT layout (zero tokens), MAE restore (fill = mask_token) and MAE compaction.
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ x, const int* __restrict__ src_idx,
                                                          const T* __restrict__ fill, T* __restrict__ out, int B,
                                                          int R_in, int R_out, int C) {
  const long w = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (w >= (long)B * R_out) return;
  const int lane = threadIdx.x & 31;
  const int b = w / R_out;
  const int src = src_idx[w];
  T* d = out + w * C;
  if (src >= 0) {
    const T* xr = x + ((long)b * R_in + src) * C;
    for (int q = lane; q < C / 4; q += 32) st4<T>(d + 4 * q, ld4<T>(xr + 4 * q));
  } else if (fill) {
    for (int q = lane; q < C / 4; q += 32) st4<T>(d + 4 * q, ld4<T>(fill + 4 * q));
  } else {
    for (int q = lane; q < C / 4; q += 32) st4<T>(d + 4 * q, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

int gather_rows(const void* x, const int* src_idx, const void* fill, void* out, int B, int R_in, int R_out, int C,
                int dtype, cudaStream_t stream) {
  SIM_REQUIRE(B > 0 && R_in > 0 && R_out > 0 && C > 0 && C % 4 == 0, SIM_ERR_INVALID,
              "gather_rows: C must be a multiple of 4");
  SIM_REQUIRE(x && src_idx && out, SIM_ERR_INVALID, "gather_rows: null tensor");
  SIM_REQUIRE(aligned16(x) && aligned16(out) && (!fill || aligned16(fill)), SIM_ERR_ALIGN,
              "gather_rows: tensors must be 16-byte aligned");
  const long rows = (long)B * R_out;
  const int grid = (int)((rows + 7) / 8);
  if (dtype == 0)
    gather_rows_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), src_idx,
                                                        static_cast<const float*>(fill), static_cast<float*>(out), B,
                                                        R_in, R_out, C);
  else if (dtype == 1)
    gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), src_idx, static_cast<const __nv_bfloat16*>(fill),
        static_cast<__nv_bfloat16*>(out), B, R_in, R_out, C);
  else {
    set_error("gather_rows: bad dtype %d", dtype);
    return SIM_ERR_INVALID;
  }
  return check_launch("gather_rows");
}

// ----------------------------------------------------------------------------- order gather backward
// dx[b,g,:] = sum_s dout[b, s*G + inv[b,s,g], :] (+ the mirrored row when reverse): a gather over the
// inverse permutations - deterministic, no atomics (every source row is read by exactly (1|2)*k output rows).
template <typename T>
__global__ void __launch_bounds__(256) order_gather_bwd_kernel(const T* __restrict__ dout,
                                                               const int* __restrict__ inv_perm, T* __restrict__ dx,
                                                               int B, int G, int k, int C, int reverse) {
  const long w = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (w >= (long)B * G) return;
  const int lane = threadIdx.x & 31;
  const int b = w / G, g = w % G;
  const long T_out = (long)(reverse ? 2 : 1) * k * G;
  for (int q = lane; q < C / 4; q += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < k; ++s) {
      const long t = (long)s * G + inv_perm[((long)b * k + s) * G + g];
      float4 a = ld4<T>(dout + ((long)b * T_out + t) * C + 4 * q);
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      if (reverse) {
        a = ld4<T>(dout + ((long)b * T_out + (T_out - 1 - t)) * C + 4 * q);
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      }
    }
    st4<T>(dx + w * C + 4 * q, acc);
  }
}

int order_gather_bwd(const void* dout, const int* inv_perm, void* dx, int B, int G, int k, int C, int reverse,
                     int dtype, cudaStream_t stream) {
  SIM_REQUIRE(B > 0 && G > 0 && k > 0 && C > 0 && C % 4 == 0, SIM_ERR_INVALID,
              "order_gather_bwd: C must be a multiple of 4");
  SIM_REQUIRE(dout && inv_perm && dx, SIM_ERR_INVALID, "order_gather_bwd: null tensor");
  SIM_REQUIRE(aligned16(dout) && aligned16(dx), SIM_ERR_ALIGN, "order_gather_bwd: tensors must be 16-byte aligned");
  const long rows = (long)B * G;
  const int grid = (int)((rows + 7) / 8);
  if (dtype == 0)
    order_gather_bwd_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(dout), inv_perm,
                                                             static_cast<float*>(dx), B, G, k, C, reverse);
  else if (dtype == 1)
    order_gather_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dout),
                                                                     inv_perm, static_cast<__nv_bfloat16*>(dx), B, G,
                                                                     k, C, reverse);
  else {
    set_error("order_gather_bwd: bad dtype %d", dtype);
    return SIM_ERR_INVALID;
  }
  return check_launch("order_gather_bwd");
}

// ----------------------------------------------------------------------------- stable argsort of rows
// One warp per row, keys staged in shared memory, rank counting on (key, index).
__global__ void __launch_bounds__(256) argsort_rows_kernel(const float* __restrict__ keys, long ld, long es, int rows,
                                                           int n, int* __restrict__ perm,
                                                           int* __restrict__ inv_perm) {
  extern __shared__ float s_keys[];  // 8 * n
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  float* kbuf = s_keys + (size_t)warp * n;
  for (int i = lane; i < n; i += 32) kbuf[i] = keys[row * ld + (long)i * es];
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const float ki = kbuf[i];
    int rank = 0;
    for (int q = 0; q < n; ++q) {
      const float kq = kbuf[q];
      rank += (kq < ki) || (kq == ki && q < i);
    }
    perm[row * n + rank] = i;
    if (inv_perm) inv_perm[row * n + i] = rank;
  }
}

int argsort_rows(const float* keys, long ld, long es, int rows, int n, int* perm, int* inv_perm,
                 cudaStream_t stream) {
  SIM_REQUIRE(rows > 0 && n > 0 && n <= 4096, SIM_ERR_INVALID, "argsort_rows: n must be in [1, 4096] (got %d)", n);
  SIM_REQUIRE(keys && perm, SIM_ERR_INVALID, "argsort_rows: null tensor");
  const size_t smem = (size_t)8 * n * sizeof(float);
  static SmemAttrCache attr;
  if (smem > 48 * 1024) ensure_dyn_smem(argsort_rows_kernel, smem, attr);
  argsort_rows_kernel<<<(rows + 7) / 8, 256, smem, stream>>>(keys, ld, es, rows, n, perm, inv_perm);
  return check_launch("argsort_rows");
}

}  // namespace sim
