// Causal depthwise conv1d (+ SiLU), forward and backward.
//
// Replaces causal-conv1d's causal_conv1d_fn as SI-Mamba reaches it through
// Mamba.forward (models/block.py:72).  Semantics (oracle/mamba.py causal_conv1d_ref):
//   y[b,t,d] = act(bias[d] + sum_{j<W} w[d,j] * x[b, t-(W-1)+j, d]),  x = 0 for t < 0.
//
// Layout: token-major (batch*L, D) with explicit row strides, so x is consumed
// in place as the first D columns of the in_proj output.  Each thread owns 4
// adjacent channels (one 16-byte / 8-byte vector) and slides a W-tap register
// window down a chunk of TC time steps: every load is a fully coalesced row
// segment, the only re-read is the (W-1)-row halo per chunk.
// Roofline: HBM, algorithmic bytes 2*B*L*D*s forward, 3*B*L*D*s backward.

#include <stdlib.h>

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
  __device__ static float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
  __device__ static float4 load(const __nv_bfloat16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  __device__ static void store(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

constexpr int kConvW = 4;

template <typename T, int TC>
__global__ void __launch_bounds__(256) causal_conv1d_fwd_kernel(const T* __restrict__ x, long ld_x,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias, T* __restrict__ y,
                                                                long ld_y, int batch, int L, int D, int silu,
                                                                __nv_bfloat16* __restrict__ planes, long ld_p,
                                                                long plane) {
  const int nv = D / 4;
  const int nchunk = (L + TC - 1) / TC;
  const long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long)batch * nchunk * nv) return;
  const int v = item % nv;
  const int ch = (item / nv) % nchunk;
  const int b = item / ((long)nv * nchunk);
  const int d0 = v * 4;
  const int t0 = ch * TC;
  const int t1 = min(L, t0 + TC);

  float wr[4][kConvW];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (long)(d0 + c) * kConvW);
    wr[c][0] = wv.x, wr[c][1] = wv.y, wr[c][2] = wv.z, wr[c][3] = wv.w;
  }
  float4 bv = bias ? *reinterpret_cast<const float4*>(bias + d0) : make_float4(0.f, 0.f, 0.f, 0.f);

  const T* xb = x + ((long)b * L) * ld_x + d0;
  T* yb = y + ((long)b * L) * ld_y + d0;
  float4 win[kConvW];  // win[j] = x[t - (W-1) + j]
#pragma unroll
  for (int j = 0; j < kConvW - 1; ++j) {
    const int t = t0 - (kConvW - 1) + j;
    win[j + 1] = t >= 0 ? Vec4<T>::load(xb + (long)t * ld_x) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll 4
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int j = 0; j < kConvW - 1; ++j) win[j] = win[j + 1];
    win[kConvW - 1] = Vec4<T>::load(xb + (long)t * ld_x);
    float4 acc = bv;
#pragma unroll
    for (int j = 0; j < kConvW; ++j) {
      acc.x = fmaf(wr[0][j], win[j].x, acc.x);
      acc.y = fmaf(wr[1][j], win[j].y, acc.y);
      acc.z = fmaf(wr[2][j], win[j].z, acc.z);
      acc.w = fmaf(wr[3][j], win[j].w, acc.w);
    }
    if (silu) {
      acc.x = silu_f(acc.x), acc.y = silu_f(acc.y), acc.z = silu_f(acc.z), acc.w = silu_f(acc.w);
    }
    Vec4<T>::store(yb + (long)t * ld_y, acc);
    if (planes) split3_store4(planes + ((long)b * L + t) * ld_p + d0, plane, acc);  // x_proj operand (gemm_split3.cu)
  }
}

// bf16 forward: 8 adjacent channels per thread so every access is a full 16-byte vector (the 4-channel kernel above
// moves only 8 bytes per bf16 access and measured 0.31 of the HBM roofline against 0.67 for fp32).
template <int TC>
__global__ void __launch_bounds__(256) causal_conv1d_fwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, long ld_x,
                                                                       const float* __restrict__ w,
                                                                       const float* __restrict__ bias,
                                                                       __nv_bfloat16* __restrict__ y, long ld_y,
                                                                       int batch, int L, int D, int silu) {
  const int nv = D / 8;
  const int nchunk = (L + TC - 1) / TC;
  const long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long)batch * nchunk * nv) return;
  const int v = item % nv;
  const int ch = (item / nv) % nchunk;
  const int b = item / ((long)nv * nchunk);
  const int d0 = v * 8;
  const int t0 = ch * TC;
  const int t1 = min(L, t0 + TC);
  float wr[8][kConvW], bv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (long)(d0 + c) * kConvW);
    wr[c][0] = wv.x, wr[c][1] = wv.y, wr[c][2] = wv.z, wr[c][3] = wv.w;
    bv[c] = bias ? bias[d0 + c] : 0.f;
  }
  const __nv_bfloat16* xb = x + ((long)b * L) * ld_x + d0;
  __nv_bfloat16* yb = y + ((long)b * L) * ld_y + d0;
  auto ld8 = [&](int t, float (&o)[8]) {
    if (t < 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = 0.f;
      return;
    }
    const uint4 r = *reinterpret_cast<const uint4*>(xb + (long)t * ld_x);
    const unsigned wds[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = __uint_as_float(wds[k] << 16);
      o[2 * k + 1] = __uint_as_float(wds[k] & 0xffff0000u);
    }
  };
  float win[kConvW][8];
#pragma unroll
  for (int j = 0; j < kConvW - 1; ++j) ld8(t0 - (kConvW - 1) + j, win[j + 1]);
#pragma unroll 4
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int j = 0; j < kConvW - 1; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) win[j][c] = win[j + 1][c];
    ld8(t, win[kConvW - 1]);
    unsigned out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a0 = bv[2 * k], a1 = bv[2 * k + 1];
#pragma unroll
      for (int j = 0; j < kConvW; ++j) {
        a0 = fmaf(wr[2 * k][j], win[j][2 * k], a0);
        a1 = fmaf(wr[2 * k + 1][j], win[j][2 * k + 1], a1);
      }
      if (silu) a0 = silu_f(a0), a1 = silu_f(a1);
      const __nv_bfloat162 pk = __floats2bfloat162_rn(a0, a1);
      out[k] = *reinterpret_cast<const unsigned*>(&pk);
    }
    *reinterpret_cast<uint4*>(yb + (long)t * ld_y) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// ----------------------------------------------------------------------------- forward through TMA tiles
// The register-window kernels above are latency-bound at the C1 / C2 batch sizes (ncu, bf16: 2.6 warps per scheduler,
// long-scoreboard stall 7.5 per issue, 0.26 of the HBM roofline): too few loads in flight.  Here a CTA owns 64 channels
// of one cloud and walks L in 32-step tiles that arrive - WITH their 3-row causal halo, rows before the sequence
// zero-filled by the TMA unit - through a 3-deep ring of tensor copies; four threads per channel convolve 8 rows each
// out of shared memory, results leave through double-buffered output tiles and TMA tensor stores (fp32 results
// optionally also as the three split bf16 planes of the x_proj operand).
constexpr int kCvTT = 32, kCvNS = 3;

struct ConvTmaps {
  CUtensorMap x, y, p[3];
};

template <typename T>
struct ConvTmaCfg {
  static constexpr int CPT = 4 / (int)sizeof(T);        // channels per thread (one 32-bit word)
  static constexpr int COLS = sizeof(T) == 4 ? 64 : 32; // four-byte columns per CTA: 64 fp32 / 64 bf16 channels (measured)
  static constexpr int SEGS = 256 / COLS;               // row segments per tile (threads = COLS x SEGS = 256)
  static constexpr int CH = COLS * CPT;                 // channels per CTA
  static constexpr int IN_TILE = (kCvTT + kConvW - 1) * COLS * 4;
  static constexpr int IN_STAGE = (IN_TILE + 127) / 128 * 128;
  static constexpr int OUT_TILE = kCvTT * COLS * 4;
  static constexpr int PL_TILE = kCvTT * COLS * 2;   // fp32 only: one bf16 plane of the 64 channels
  static constexpr int OUT_BUF = OUT_TILE + (sizeof(T) == 4 ? 3 * PL_TILE : 0);
  static constexpr int SMEM = kCvNS * IN_STAGE + 2 * OUT_BUF + kCvNS * 8 + 16;
};

template <typename T>
__global__ void __launch_bounds__(256) causal_conv1d_fwd_tma_kernel(const __grid_constant__ ConvTmaps tm,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias, int L, int D, int silu,
                                                                     int has_planes) {
  using Cfg = ConvTmaCfg<T>;
  constexpr int COLS = Cfg::COLS, CPT = Cfg::CPT, CH = Cfg::CH, TT = kCvTT, NS = kCvNS, HALO = kConvW - 1, RPS = kCvTT / Cfg::SEGS;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* obuf = smem + NS * Cfg::IN_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(obuf + 2 * Cfg::OUT_BUF);
  const int tid = threadIdx.x;
  const int nchunk = D / CH;
  const int b = blockIdx.x / nchunk, c0 = (blockIdx.x % nchunk) * CH;
  const int col = tid % COLS, seg = tid / COLS;  // SEGS row segments of RPS steps per column
  const int ntiles = (L + TT - 1) / TT;

  if (tid == 0) {
    tma_prefetch_desc(&tm.x);
    tma_prefetch_desc(&tm.y);
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int tile) {
    const int s = tile % NS;
    mbar_arrive_expect_tx(&full[s], (TT + HALO) * COLS * 4);
    tma_load_3d(smem + s * Cfg::IN_STAGE, &tm.x, c0, tile * TT - HALO, b, &full[s]);  // rows < 0: zero-filled
  };
  if (tid == 0)
    for (int k = 0; k < NS && k < ntiles; ++k) issue(k);

  float wt[CPT][kConvW], bv[CPT];
#pragma unroll
  for (int q = 0; q < CPT; ++q) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (long)(c0 + col * CPT + q) * kConvW);
    wt[q][0] = wv.x, wt[q][1] = wv.y, wt[q][2] = wv.z, wt[q][3] = wv.w;
    bv[q] = bias ? bias[c0 + col * CPT + q] : 0.f;
  }
  // one 32-bit word of a tile row = this thread's CPT channels
  auto unpack = [](uint32_t word, float (&v)[CPT]) {
    if constexpr (CPT == 1) {
      v[0] = __uint_as_float(word);
    } else {
      v[0] = __uint_as_float(word << 16), v[1] = __uint_as_float(word & 0xffff0000u);
    }
  };

  for (int tile = 0; tile < ntiles; ++tile) {
    const int s = tile % NS, ob = tile & 1;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(smem + s * Cfg::IN_STAGE);  // [TT + 3][COLS], row i <-> t0 - 3 + i
    uint32_t* out = reinterpret_cast<uint32_t*>(obuf + ob * Cfg::OUT_BUF);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(obuf + ob * Cfg::OUT_BUF + Cfg::OUT_TILE);
    mbar_wait(&full[s], (tile / NS) & 1);
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // output buffer `ob` (tile - 2) was read
    __syncthreads();
    const int r0 = seg * RPS;
    float x0[CPT], x1[CPT], x2[CPT], x3[CPT];
    unpack(in[(r0 + 0) * COLS + col], x0);
    unpack(in[(r0 + 1) * COLS + col], x1);
    unpack(in[(r0 + 2) * COLS + col], x2);
#pragma unroll
    for (int j = 0; j < RPS; ++j) {
      unpack(in[(r0 + HALO + j) * COLS + col], x3);
      float acc[CPT];
#pragma unroll
      for (int q = 0; q < CPT; ++q) {
        acc[q] = fmaf(wt[q][3], x3[q], fmaf(wt[q][2], x2[q], fmaf(wt[q][1], x1[q], fmaf(wt[q][0], x0[q], bv[q]))));
        if (silu) acc[q] = silu_f(acc[q]);
        x0[q] = x1[q], x1[q] = x2[q], x2[q] = x3[q];
      }
      if constexpr (CPT == 1) {
        out[(r0 + j) * COLS + col] = __float_as_uint(acc[0]);
        if (has_planes) {
          const __nv_bfloat16 p0 = __float2bfloat16_rn(acc[0]);
          const float r1 = acc[0] - __bfloat162float(p0);
          const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
          pl[(r0 + j) * COLS + col] = p0;
          pl[TT * COLS + (r0 + j) * COLS + col] = p1;
          pl[2 * TT * COLS + (r0 + j) * COLS + col] = __float2bfloat16_rn(r1 - __bfloat162float(p1));
        }
      } else {
        const __nv_bfloat162 pk = __floats2bfloat162_rn(acc[0], acc[1]);
        out[(r0 + j) * COLS + col] = *reinterpret_cast<const uint32_t*>(&pk);
      }
    }
    fence_proxy_async();
    __syncthreads();  // the tile is complete; input stage s is free again
    if (tid == 0) {
      tma_store_3d(&tm.y, c0, tile * TT, b, out);  // rows past L are clipped
      if (CPT == 1 && has_planes) {
#pragma unroll
        for (int q = 0; q < 3; ++q) tma_store_3d(&tm.p[q], c0, tile * TT, b, pl + q * TT * COLS);
      }
      bulk_commit();
      if (tile + NS < ntiles) issue(tile + NS);
    }
  }
  if (tid == 0) bulk_wait0();
}

template <typename T>
int launch_conv_tma(const void* x, long ld_x, const float* w, const float* bias, void* y, long ld_y, int batch, int L, int D,
                    int silu, int dtype, void* planes, long ld_p, long plane, cudaStream_t stream) {
  using Cfg = ConvTmaCfg<T>;
  auto kern = causal_conv1d_fwd_tma_kernel<T>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, Cfg::SMEM, attr) != cudaSuccess) return check_launch("causal_conv1d_fwd_tma attr");
  ConvTmaps tm;
  int rc;
  if ((rc = make_tmap_tokens(&tm.x, x, dtype, D, L, batch, ld_x, Cfg::CH, kCvTT + kConvW - 1))) return rc;
  if ((rc = make_tmap_tokens(&tm.y, y, dtype, D, L, batch, ld_y, Cfg::CH, kCvTT))) return rc;
  for (int q = 0; q < 3; ++q) {
    if (planes) {
      if ((rc = make_tmap_tokens(&tm.p[q], static_cast<__nv_bfloat16*>(planes) + q * plane, 1, D, L, batch, ld_p, Cfg::CH, kCvTT)))
        return rc;
    } else {
      tm.p[q] = tm.y;
    }
  }
  kern<<<batch * (D / Cfg::CH), 256, Cfg::SMEM, stream>>>(tm, w, bias, L, D, silu, planes != nullptr);
  return check_launch("causal_conv1d_fwd_tma");
}

constexpr int kTcBf16 = 32, kTcF32 = 32, kTcF32Planes = 32;  // time steps per thread (see causal_conv1d_fwd)

int causal_conv1d_fwd(const void* x, long ld_x, const float* w, const float* bias, void* y, long ld_y, int batch,
                      int L, int D, int width, int silu, int dtype, cudaStream_t stream, void* planes, long ld_p,
                      long plane) {
  SIM_REQUIRE(!planes || (dtype == 0 && (reinterpret_cast<uintptr_t>(planes) & 7u) == 0 && ld_p % 4 == 0 && plane % 4 == 0),
              SIM_ERR_INVALID, "causal_conv1d_fwd: split planes need fp32 activations and 8-byte alignment");
  SIM_REQUIRE(width == kConvW, SIM_ERR_INVALID, "causal_conv1d_fwd: only width 4 is built (got %d)", width);
  SIM_REQUIRE(D % 4 == 0 && batch > 0 && L > 0, SIM_ERR_INVALID, "causal_conv1d_fwd: D must be a multiple of 4");
  SIM_REQUIRE(dtype == 0 || dtype == 1, SIM_ERR_INVALID, "causal_conv1d_fwd: dtype must be 0 (fp32) or 1 (bf16)");
  const int es = dtype == 0 ? 4 : 2;
  const uintptr_t vmask = 4 * es - 1;
  SIM_REQUIRE(((uintptr_t)x & vmask) == 0 && ((uintptr_t)y & vmask) == 0 && aligned16(w) && ld_x % 4 == 0 &&
                  ld_y % 4 == 0 && (!bias || aligned16(bias)),
              SIM_ERR_ALIGN, "causal_conv1d_fwd: x/y/w/bias need 16-byte bases and vector-aligned row strides");
  // Time chunk per thread.  Shorter chunks (more threads, 3-row halo from L1/L2) measured no faster on B200:
  // fp32 22.2 us at TC=16 vs 22.5 at TC=32; bf16 27.3 us at TC=8 vs 23.1 at TC=32.
  // SIM_CONV_TC (bench-only): time steps per thread.  Default chosen from the measurements in the comment above /
  // profiles/r01_kernel_bench_all.jsonl.
  static const int tc_env = [] { const char* e = getenv("SIM_CONV_TC"); return e ? atoi(e) : 0; }();
  const int es_ = dtype == 0 ? 4 : 2;
  if (tc_env == 0 && D % 64 == 0 && aligned16(x) && aligned16(y) && (ld_x * es_) % 16 == 0 && (ld_y * es_) % 16 == 0 &&
      (!planes || (aligned16(planes) && ld_p % 8 == 0 && plane % 8 == 0))) {
    return dtype == 0 ? launch_conv_tma<float>(x, ld_x, w, bias, y, ld_y, batch, L, D, silu, dtype, planes, ld_p, plane, stream)
                      : launch_conv_tma<__nv_bfloat16>(x, ld_x, w, bias, y, ld_y, batch, L, D, silu, dtype, nullptr, 0, 0, stream);
  }
  if (dtype == 1 && D % 8 == 0 && aligned16(x) && aligned16(y) && ld_x % 8 == 0 && ld_y % 8 == 0) {
    const int tc = tc_env ? tc_env : kTcBf16;
    const long items = (long)batch * ((L + tc - 1) / tc) * (D / 8);
    const int grid8 = (int)((items + 255) / 256);
    const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
    if (tc == 8) causal_conv1d_fwd_bf16x8_kernel<8><<<grid8, 256, 0, stream>>>(xb, ld_x, w, bias, yb, ld_y, batch, L, D, silu);
    else if (tc == 16) causal_conv1d_fwd_bf16x8_kernel<16><<<grid8, 256, 0, stream>>>(xb, ld_x, w, bias, yb, ld_y, batch, L, D, silu);
    else causal_conv1d_fwd_bf16x8_kernel<32><<<grid8, 256, 0, stream>>>(xb, ld_x, w, bias, yb, ld_y, batch, L, D, silu);
    return check_launch("causal_conv1d_fwd");
  }
  const int tc = tc_env ? tc_env : (planes ? kTcF32Planes : kTcF32);
  const long items = (long)batch * ((L + tc - 1) / tc) * (D / 4);
  const int grid = (int)((items + 255) / 256);
  if (dtype == 0) {
    const float* xf = static_cast<const float*>(x);
    float* yf = static_cast<float*>(y);
    __nv_bfloat16* pp = static_cast<__nv_bfloat16*>(planes);
    if (tc == 8) causal_conv1d_fwd_kernel<float, 8><<<grid, 256, 0, stream>>>(xf, ld_x, w, bias, yf, ld_y, batch, L, D, silu, pp, ld_p, plane);
    else if (tc == 16) causal_conv1d_fwd_kernel<float, 16><<<grid, 256, 0, stream>>>(xf, ld_x, w, bias, yf, ld_y, batch, L, D, silu, pp, ld_p, plane);
    else causal_conv1d_fwd_kernel<float, 32><<<grid, 256, 0, stream>>>(xf, ld_x, w, bias, yf, ld_y, batch, L, D, silu, pp, ld_p, plane);
  } else {
    causal_conv1d_fwd_kernel<__nv_bfloat16, 32><<<(int)(((long)batch * ((L + 31) / 32) * (D / 4) + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), ld_x, w, bias, static_cast<__nv_bfloat16*>(y), ld_y, batch, L, D, silu, nullptr, 0, 0);
  }
  return check_launch("causal_conv1d_fwd");
}

// ----------------------------------------------------------------------------- backward
// p = bias + sum_j w_j x[t-3+j];  y = silu(p).  dp = dy * silu'(p);  dx[t] = sum_j w_j dp[t+3-j];
// dw_j += sum_t dp[t] x[t-3+j];  db += sum_t dp[t].  Each thread owns 4 adjacent channels and a chunk of TC steps;
// it recomputes p over [t0, t0+TC+3) from x (halo 3 rows each side), keeps dp in a 4-deep register window and
// accumulates its dw / db partials in registers -> one fp32 atomic per (thread, channel, tap).
template <typename T, int TC>
__global__ void __launch_bounds__(128) causal_conv1d_bwd_kernel(const T* __restrict__ x, long ld_x,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                const T* __restrict__ dy, long ld_dy,
                                                                T* __restrict__ dx, long ld_dx, float* __restrict__ dw,
                                                                float* __restrict__ db, int batch, int L, int D,
                                                                int silu) {
  const int nv = D / 4;
  const int nchunk = (L + TC - 1) / TC;
  const long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long)batch * nchunk * nv) return;
  const int v = item % nv;
  const int ch = (item / nv) % nchunk;
  const int b = item / ((long)nv * nchunk);
  const int d0 = v * 4;
  const int t0 = ch * TC;
  const int t1 = min(L, t0 + TC);

  float wr[4][kConvW];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (long)(d0 + c) * kConvW);
    wr[c][0] = wv.x, wr[c][1] = wv.y, wr[c][2] = wv.z, wr[c][3] = wv.w;
  }
  const float4 bv = bias ? *reinterpret_cast<const float4*>(bias + d0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const T* xb = x + ((long)b * L) * ld_x + d0;
  const T* gb = dy + ((long)b * L) * ld_dy + d0;
  T* ob = dx + ((long)b * L) * ld_dx + d0;

  auto ldx = [&](int t) { return (t >= 0 && t < L) ? Vec4<T>::load(xb + (long)t * ld_x) : make_float4(0.f, 0.f, 0.f, 0.f); };
  // xw[j] = x[s-3+j] for the step s whose dp is being produced
  float4 xw[kConvW];
#pragma unroll
  for (int j = 0; j < kConvW - 1; ++j) xw[j + 1] = ldx(t0 - (kConvW - 1) + j);
  float4 dpw[kConvW];  // dpw[j] = dp[s-3+j] after producing step s
#pragma unroll
  for (int j = 0; j < kConvW; ++j) dpw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dwa[4][kConvW] = {};
  float dba[4] = {};

  // produce dp for s = t0 .. t1+2 (clipped to L); once dp[s] is known, dx[s-3] is complete
  for (int s = t0; s < t1 + kConvW - 1; ++s) {
#pragma unroll
    for (int j = 0; j < kConvW - 1; ++j) xw[j] = xw[j + 1];
    xw[kConvW - 1] = ldx(s);
    float4 dps = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < L) {
      float4 pre = bv;
#pragma unroll
      for (int j = 0; j < kConvW; ++j) {
        pre.x = fmaf(wr[0][j], xw[j].x, pre.x);
        pre.y = fmaf(wr[1][j], xw[j].y, pre.y);
        pre.z = fmaf(wr[2][j], xw[j].z, pre.z);
        pre.w = fmaf(wr[3][j], xw[j].w, pre.w);
      }
      const float4 g = Vec4<T>::load(gb + (long)s * ld_dy);
      if (silu) {
        const float sx = sigmoid_f(pre.x), sy = sigmoid_f(pre.y), sz = sigmoid_f(pre.z), sw = sigmoid_f(pre.w);
        dps.x = g.x * sx * (1.f + pre.x * (1.f - sx));
        dps.y = g.y * sy * (1.f + pre.y * (1.f - sy));
        dps.z = g.z * sz * (1.f + pre.z * (1.f - sz));
        dps.w = g.w * sw * (1.f + pre.w * (1.f - sw));
      } else {
        dps = g;
      }
      if (s < t1) {  // parameter gradients: each time step is owned by exactly one chunk
#pragma unroll
        for (int j = 0; j < kConvW; ++j) {
          dwa[0][j] = fmaf(dps.x, xw[j].x, dwa[0][j]);
          dwa[1][j] = fmaf(dps.y, xw[j].y, dwa[1][j]);
          dwa[2][j] = fmaf(dps.z, xw[j].z, dwa[2][j]);
          dwa[3][j] = fmaf(dps.w, xw[j].w, dwa[3][j]);
        }
        dba[0] += dps.x, dba[1] += dps.y, dba[2] += dps.z, dba[3] += dps.w;
      }
    }
#pragma unroll
    for (int j = 0; j < kConvW - 1; ++j) dpw[j] = dpw[j + 1];
    dpw[kConvW - 1] = dps;
    const int t = s - (kConvW - 1);  // dx[t] = sum_j w_j dp[t+3-j] = sum_j w_j dpw[3-j]
    if (t >= t0 && t < t1) {
      float4 o;
      o.x = wr[0][0] * dpw[3].x + wr[0][1] * dpw[2].x + wr[0][2] * dpw[1].x + wr[0][3] * dpw[0].x;
      o.y = wr[1][0] * dpw[3].y + wr[1][1] * dpw[2].y + wr[1][2] * dpw[1].y + wr[1][3] * dpw[0].y;
      o.z = wr[2][0] * dpw[3].z + wr[2][1] * dpw[2].z + wr[2][2] * dpw[1].z + wr[2][3] * dpw[0].z;
      o.w = wr[3][0] * dpw[3].w + wr[3][1] * dpw[2].w + wr[3][2] * dpw[1].w + wr[3][3] * dpw[0].w;
      Vec4<T>::store(ob + (long)t * ld_dx, o);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int j = 0; j < kConvW; ++j) atomicAdd(dw + (long)(d0 + c) * kConvW + j, dwa[c][j]);
    if (db) atomicAdd(db + d0 + c, dba[c]);
  }
}

// Tiled backward (default when D % 32 == 0).  The register-window kernel above issues one dependent global load per
// step per thread and 20 atomics per thread (195 us at the C2 layer shape, 0.12 of the HBM roofline).  Here a CTA owns
// 32 channels of one cloud and walks L in 64-step tiles: x (3-row halo on both sides) and dy (3 rows past the end) are
// staged in shared memory with 16-byte loads, dp = dy * silu'(pre) is formed in place, dx comes out of shared memory, and
// dw / db accumulate in registers over the whole sequence -> one warp-shuffle + shared-memory reduction and 160 atomics
// per CTA at the end.
constexpr int kBwdCH = 32, kBwdTC = 64, kBwdLC = 256;

template <typename T>
__global__ void __launch_bounds__(256) causal_conv1d_bwd_tiled_kernel(const T* __restrict__ x, long ld_x,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ bias,
                                                                      const T* __restrict__ dy, long ld_dy,
                                                                      T* __restrict__ dx, long ld_dx,
                                                                      float* __restrict__ dw, float* __restrict__ db, int L,
                                                                      int D, int silu) {
  constexpr int CH = kBwdCH, TC = kBwdTC, NV = CH / 4, NRG = 256 / NV;  // 8 channel quads x 32 row groups
  __shared__ __align__(16) float xs[(TC + 2 * (kConvW - 1)) * CH];  // x rows t0-3 .. t0+TC+2
  __shared__ __align__(16) float dp[(TC + kConvW - 1) * CH];        // dy, then dp, rows t0 .. t0+TC+2
  __shared__ float red[8][NV][20];
  const int nchunk = D / CH;
  const int b = blockIdx.x / nchunk, c0 = (blockIdx.x % nchunk) * CH;
  const int tid = threadIdx.x, c4 = tid % NV, rg = tid / NV;
  const int d0 = c0 + 4 * c4;
  float wr[4][kConvW];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (long)(d0 + c) * kConvW);
    wr[c][0] = wv.x, wr[c][1] = wv.y, wr[c][2] = wv.z, wr[c][3] = wv.w;
  }
  const float4 bv = bias ? *reinterpret_cast<const float4*>(bias + d0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const T* xb = x + ((long)b * L) * ld_x + d0;
  const T* gb = dy + ((long)b * L) * ld_dy + d0;
  T* ob = dx + ((long)b * L) * ld_dx + d0;
  float dwa[4][kConvW] = {};
  float dba[4] = {};
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // tiles of this CTA: steps [t_begin, t_end) (gridDim.y chunks of kBwdLC steps keep ~7 waves of CTAs in flight); the
  // next tile's rows are fetched into registers while the current tile is processed (software pipeline, no smem cost)
  constexpr int NLX = (TC + 2 * (kConvW - 1) + NRG - 1) / NRG, NLG = (TC + kConvW - 1 + NRG - 1) / NRG;
  const int t_begin = blockIdx.y * kBwdLC, t_end = min(L, t_begin + kBwdLC);
  float4 px[NLX], pg[NLG];
  auto fetch = [&](int t0) {
#pragma unroll
    for (int i = 0; i < NLX; ++i) {
      const int r = rg + i * NRG, t = t0 - (kConvW - 1) + r;
      px[i] = (r < TC + 2 * (kConvW - 1) && t >= 0 && t < L) ? Vec4<T>::load(xb + (long)t * ld_x) : zero4;
    }
#pragma unroll
    for (int i = 0; i < NLG; ++i) {
      const int r = rg + i * NRG, t = t0 + r;
      pg[i] = (r < TC + kConvW - 1 && t < L) ? Vec4<T>::load(gb + (long)t * ld_dy) : zero4;
    }
  };
  fetch(t_begin);
  for (int t0 = t_begin; t0 < t_end; t0 += TC) {
#pragma unroll
    for (int i = 0; i < NLX; ++i) {
      const int r = rg + i * NRG;
      if (r < TC + 2 * (kConvW - 1)) *reinterpret_cast<float4*>(xs + r * CH + 4 * c4) = px[i];
    }
#pragma unroll
    for (int i = 0; i < NLG; ++i) {
      const int r = rg + i * NRG;
      if (r < TC + kConvW - 1) *reinterpret_cast<float4*>(dp + r * CH + 4 * c4) = pg[i];
    }
    __syncthreads();
    if (t0 + TC < t_end) fetch(t0 + TC);
    // dp[s] = dy[s] * silu'(pre[s]) in place; parameter gradients from the rows this tile owns (r < TC)
    for (int r = rg; r < TC + kConvW - 1; r += NRG) {
      float4 xv[kConvW];
#pragma unroll
      for (int j = 0; j < kConvW; ++j) xv[j] = *reinterpret_cast<const float4*>(xs + (r + j) * CH + 4 * c4);  // x[s-3+j]
      float4 d = *reinterpret_cast<const float4*>(dp + r * CH + 4 * c4);
      if (silu) {
        float4 pre = bv;
#pragma unroll
        for (int j = 0; j < kConvW; ++j) {
          pre.x = fmaf(wr[0][j], xv[j].x, pre.x);
          pre.y = fmaf(wr[1][j], xv[j].y, pre.y);
          pre.z = fmaf(wr[2][j], xv[j].z, pre.z);
          pre.w = fmaf(wr[3][j], xv[j].w, pre.w);
        }
        const float sx = sigmoid_f(pre.x), sy = sigmoid_f(pre.y), sz = sigmoid_f(pre.z), sw = sigmoid_f(pre.w);
        d.x = d.x * sx * (1.f + pre.x * (1.f - sx));
        d.y = d.y * sy * (1.f + pre.y * (1.f - sy));
        d.z = d.z * sz * (1.f + pre.z * (1.f - sz));
        d.w = d.w * sw * (1.f + pre.w * (1.f - sw));
        *reinterpret_cast<float4*>(dp + r * CH + 4 * c4) = d;
      }
      if (r < TC && t0 + r < t_end) {  // each step's parameter gradient belongs to exactly one chunk
#pragma unroll
        for (int j = 0; j < kConvW; ++j) {
          dwa[0][j] = fmaf(d.x, xv[j].x, dwa[0][j]);
          dwa[1][j] = fmaf(d.y, xv[j].y, dwa[1][j]);
          dwa[2][j] = fmaf(d.z, xv[j].z, dwa[2][j]);
          dwa[3][j] = fmaf(d.w, xv[j].w, dwa[3][j]);
        }
        dba[0] += d.x, dba[1] += d.y, dba[2] += d.z, dba[3] += d.w;
      }
    }
    __syncthreads();
    // dx[t] = sum_j w_j dp[t+3-j]
    for (int r = rg; r < TC && t0 + r < t_end; r += NRG) {
      float4 o = zero4;
#pragma unroll
      for (int j = 0; j < kConvW; ++j) {
        const float4 q = *reinterpret_cast<const float4*>(dp + (r + kConvW - 1 - j) * CH + 4 * c4);
        o.x = fmaf(wr[0][j], q.x, o.x);
        o.y = fmaf(wr[1][j], q.y, o.y);
        o.z = fmaf(wr[2][j], q.z, o.z);
        o.w = fmaf(wr[3][j], q.w, o.w);
      }
      Vec4<T>::store(ob + (long)(t0 + r) * ld_dx, o);
    }
    __syncthreads();  // the next tile overwrites xs / dp
  }
  // reduce dw / db over the 32 row groups: lanes with equal (lane % 8) inside a warp, then the 8 warps
  float vals[20];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int j = 0; j < kConvW; ++j) vals[c * kConvW + j] = dwa[c][j];
    vals[16 + c] = dba[c];
  }
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], 8);
    vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], 16);
  }
  const int lane = tid & 31, warp = tid >> 5;
  if (lane < NV) {
#pragma unroll
    for (int i = 0; i < 20; ++i) red[warp][lane][i] = vals[i];
  }
  __syncthreads();
  if (tid < NV * 20) {
    const int q = tid / 20, i = tid % 20;
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][q][i];
    if (i < 16) atomicAdd(dw + (long)(c0 + 4 * q + i / kConvW) * kConvW + (i % kConvW), v);
    else if (db) atomicAdd(db + c0 + 4 * q + (i - 16), v);
  }
}

int causal_conv1d_bwd(const void* x, long ld_x, const float* w, const float* bias, const void* dy, long ld_dy,
                      void* dx, long ld_dx, float* dw, float* db, int batch, int L, int D, int width, int silu,
                      int dtype, cudaStream_t stream) {
  SIM_REQUIRE(width == kConvW, SIM_ERR_INVALID, "causal_conv1d_bwd: only width 4 is built (got %d)", width);
  SIM_REQUIRE(D % 4 == 0 && batch > 0 && L > 0, SIM_ERR_INVALID, "causal_conv1d_bwd: D must be a multiple of 4");
  SIM_REQUIRE(dtype == 0 || dtype == 1, SIM_ERR_INVALID, "causal_conv1d_bwd: dtype must be 0 (fp32) or 1 (bf16)");
  SIM_REQUIRE(x && w && dy && dx && dw, SIM_ERR_INVALID, "causal_conv1d_bwd: null tensor");
  const int es = dtype == 0 ? 4 : 2;
  const uintptr_t vmask = 4 * es - 1;
  SIM_REQUIRE(((uintptr_t)x & vmask) == 0 && ((uintptr_t)dy & vmask) == 0 && ((uintptr_t)dx & vmask) == 0 &&
                  aligned16(w) && ld_x % 4 == 0 && ld_dy % 4 == 0 && ld_dx % 4 == 0 && (!bias || aligned16(bias)),
              SIM_ERR_ALIGN, "causal_conv1d_bwd: tensors need vector-aligned bases and row strides");
  static const int force_window = [] { const char* e = getenv("SIM_CONV_BWD_WINDOW"); return e ? atoi(e) : 0; }();
  if (D % kBwdCH == 0 && !force_window) {
    const dim3 grid(batch * (D / kBwdCH), (L + kBwdLC - 1) / kBwdLC);
    if (dtype == 0)
      causal_conv1d_bwd_tiled_kernel<float><<<grid, 256, 0, stream>>>(
          static_cast<const float*>(x), ld_x, w, bias, static_cast<const float*>(dy), ld_dy, static_cast<float*>(dx), ld_dx,
          dw, db, L, D, silu);
    else
      causal_conv1d_bwd_tiled_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(x), ld_x, w, bias, static_cast<const __nv_bfloat16*>(dy), ld_dy,
          static_cast<__nv_bfloat16*>(dx), ld_dx, dw, db, L, D, silu);
    return check_launch("causal_conv1d_bwd_tiled");
  }
  constexpr int TC = 64;
  const long items = (long)batch * ((L + TC - 1) / TC) * (D / 4);
  const int grid = (int)((items + 127) / 128);
  if (dtype == 0)
    causal_conv1d_bwd_kernel<float, TC><<<grid, 128, 0, stream>>>(
        static_cast<const float*>(x), ld_x, w, bias, static_cast<const float*>(dy), ld_dy, static_cast<float*>(dx),
        ld_dx, dw, db, batch, L, D, silu);
  else
    causal_conv1d_bwd_kernel<__nv_bfloat16, TC><<<grid, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), ld_x, w, bias, static_cast<const __nv_bfloat16*>(dy), ld_dy,
        static_cast<__nv_bfloat16*>(dx), ld_dx, dw, db, batch, L, D, silu);
  return check_launch("causal_conv1d_bwd");
}

}  // namespace sim
