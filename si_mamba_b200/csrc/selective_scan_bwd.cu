// Selective scan backward (training configs C2-C4), sm_100a.
//
// Replaces mamba-ssm's selective_scan_bwd as reached through Mamba.forward's autograd node
// (models/block.py:72).  Gradients of
//   dt = softplus(delta + bias); h_t = exp(dt_t A) h_{t-1} + dt_t B_t u_t; y_t = <h_t, C_t> + D u_t; out = y silu(z)
// w.r.t. u, delta, z (token-major, input dtype), B, C (fp32, summed over channels), A, D, bias (fp32).
//
// No (B, D, L, N) state tensor is stored: the training forward keeps only the state at the start of every
// kScanTile-step tile (ScanParams::ckpt, one fp32 tensor the size of an activation).  Per tile, walking the
// sequence backwards, this kernel (1) recomputes h_t forward from the checkpoint into shared memory and y_t on the
// fly, (2) runs the adjoint recurrence dh_{t-1} = a_t dh_t in reverse.  Thread = one channel x all 16 states, so
// every reduction over states is thread-local; dB / dC (reductions over channels) go through a transposed warp
// butterfly (32 values over 32 lanes) and one shared + one global fp32 atomic per (t, n) per CTA.
// Operand tiles arrive by TMA tensor copies exactly as in the forward kernel.
//
// Roofline class: SM issue (about 25 instructions per state update); HBM algorithmic bytes are
// (5 reads + 3 writes) * E * s + checkpoint E * 4 + O(S).  First version - correctness first.

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

namespace {

constexpr int kN = 16;
constexpr int TT = kScanTile;

struct BwdTmaps {
  CUtensorMap u, delta, z, B, C, dout, du, ddelta, dz;
};

template <typename T>
__device__ __forceinline__ float4 ldsv4(const T* p);
template <>
__device__ __forceinline__ float4 ldsv4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 ldsv4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}

template <typename T, int CH_>
struct BwdCfg {
  static constexpr int CH = CH_;
  static constexpr int NT = CH_;  // one thread per channel
  static constexpr int NS = 2;
  static constexpr int RAW_MAIN = TT * CH_ * (int)sizeof(T);
  static constexpr int RAW_BC = TT * kN * (int)sizeof(T);
  static constexpr int RAW_STAGE = 4 * RAW_MAIN + 2 * RAW_BC;  // u, delta, z, dout, B, C
  static constexpr int OUT = 3 * RAW_MAIN;                      // du, ddelta, dz
  static constexpr int WORK = 7 * TT * CH_ * 4                  // dt, u, sg, dy, dzc, y, (spare)
                              + 2 * TT * kN * 4                 // B, C fp32
                              + 2 * TT * kN * 4;                // dB, dC tile accumulators
  static constexpr int HBUF = TT * CH_ * kN * 4;
  static constexpr int SMEM = NS * RAW_STAGE + OUT + WORK + HBUF + NS * 8 + 64;
  static_assert(RAW_MAIN % 128 == 0 && RAW_BC % 128 == 0, "TMA tiles must stay 128-B aligned");
};

template <typename Cfg, typename T>
__global__ void __launch_bounds__(Cfg::NT) selective_scan_bwd_kernel(const __grid_constant__ BwdTmaps tm,
                                                                     const ScanBwdParams p) {
  constexpr int CH = Cfg::CH, NT = Cfg::NT, NS = Cfg::NS;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;
  T* __restrict__ o_du = reinterpret_cast<T*>(smem + NS * Cfg::RAW_STAGE);
  T* __restrict__ o_dd = o_du + TT * CH;
  T* __restrict__ o_dz = o_dd + TT * CH;
  float* __restrict__ w_dt = reinterpret_cast<float*>(smem + NS * Cfg::RAW_STAGE + Cfg::OUT);
  float* __restrict__ w_u = w_dt + TT * CH;
  float* __restrict__ w_sg = w_u + TT * CH;
  float* __restrict__ w_dy = w_sg + TT * CH;
  float* __restrict__ w_dzc = w_dy + TT * CH;
  float* __restrict__ w_y = w_dzc + TT * CH;
  float* __restrict__ w_sp = w_y + TT * CH;
  float* __restrict__ w_B = w_sp + TT * CH;
  float* __restrict__ w_C = w_B + TT * kN;
  float* __restrict__ a_dB = w_C + TT * kN;
  float* __restrict__ a_dC = a_dB + TT * kN;
  float* __restrict__ hbuf = a_dC + TT * kN;  // [t][c][n]
  uint64_t* full = reinterpret_cast<uint64_t*>(hbuf + TT * CH * kN);

  const int tid = threadIdx.x, lane = tid & 31;
  const int nchunk = p.D / CH;
  const int b = blockIdx.x / nchunk;
  const int c0 = (blockIdx.x % nchunk) * CH;
  const int c = tid;
  const int ntiles = (p.L + TT - 1) / TT;
  const bool has_z = p.z != nullptr;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue_tile = [&](int k) {  // k-th tile in processing order = tile index ntiles-1-k
    const int s = k % NS;
    const int t0 = (ntiles - 1 - k) * TT;
    unsigned char* st = raw + s * Cfg::RAW_STAGE;
    mbar_arrive_expect_tx(&full[s], (has_z ? 4u : 3u) * Cfg::RAW_MAIN + 2u * Cfg::RAW_BC);
    tma_load_3d(st, &tm.u, c0, t0, b, &full[s]);
    tma_load_3d(st + Cfg::RAW_MAIN, &tm.delta, c0, t0, b, &full[s]);
    if (has_z) tma_load_3d(st + 2 * Cfg::RAW_MAIN, &tm.z, c0, t0, b, &full[s]);
    tma_load_3d(st + 3 * Cfg::RAW_MAIN, &tm.dout, c0, t0, b, &full[s]);
    tma_load_3d(st + 4 * Cfg::RAW_MAIN, &tm.B, 0, t0, b, &full[s]);
    tma_load_3d(st + 4 * Cfg::RAW_MAIN + Cfg::RAW_BC, &tm.C, 0, t0, b, &full[s]);
  };
  if (tid == 0) {
    for (int k = 0; k < NS && k < ntiles; ++k) issue_tile(k);
  }

  float A[kN], A2[kN], dA[kN], dh[kN];
#pragma unroll
  for (int n = 0; n < kN; ++n) {
    A[n] = p.A[(long)(c0 + c) * kN + n];
    A2[n] = A[n] * kLog2e;
    dA[n] = 0.f;
    dh[n] = 0.f;
  }
  const float Dc = p.Dv ? p.Dv[c0 + c] : 0.f;
  const float bias_c = p.dbias ? p.dbias[c0 + c] : 0.f;
  float dD = 0.f, dbias = 0.f;

  for (int k = 0; k < ntiles; ++k) {
    const int s = k % NS;
    const int tile = ntiles - 1 - k;
    const int t0 = tile * TT;
    const int rows = min(TT, p.L - t0);
    unsigned char* st = raw + s * Cfg::RAW_STAGE;
    const T* su = reinterpret_cast<const T*>(st);
    const T* sd = reinterpret_cast<const T*>(st + Cfg::RAW_MAIN);
    const T* sz = reinterpret_cast<const T*>(st + 2 * Cfg::RAW_MAIN);
    const T* so = reinterpret_cast<const T*>(st + 3 * Cfg::RAW_MAIN);
    const T* sB = reinterpret_cast<const T*>(st + 4 * Cfg::RAW_MAIN);
    const T* sC = reinterpret_cast<const T*>(st + 4 * Cfg::RAW_MAIN + Cfg::RAW_BC);

    // state at the start of this tile (from the training forward)
    float h[kN];
    {
      const float4* cp = reinterpret_cast<const float4*>(p.ckpt + (((long)b * ntiles + tile) * p.D + c0 + c) * kN);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = cp[q];
        h[4 * q] = v.x, h[4 * q + 1] = v.y, h[4 * q + 2] = v.z, h[4 * q + 3] = v.w;
      }
    }

    mbar_wait(&full[s], (k / NS) & 1);

    // ---- pre-pass (each thread its own channel column): activations and their derivatives
    for (int r = 0; r < TT; ++r) {
      const int e = r * CH + c;
      const float x = to_f32<T>(sd[e]) + bias_c;
      w_dt[e] = p.softplus ? softplus_f(x) : x;
      w_sg[e] = p.softplus ? ((x > 20.f) ? 1.f : sigmoid_f(x)) : 1.f;  // d softplus / dx
      w_u[e] = to_f32<T>(su[e]);
      const float go = to_f32<T>(so[e]);
      if (has_z) {
        const float zv = to_f32<T>(sz[e]);
        const float sg = sigmoid_f(zv);
        w_dy[e] = go * zv * sg;                               // dL/dy = dout * silu(z)
        w_dzc[e] = go * sg * (1.f + zv * (1.f - sg));         // dout * silu'(z); dz = this * y
      } else {
        w_dy[e] = go;
        w_dzc[e] = 0.f;
      }
    }
    for (int e = tid; e < TT * kN; e += NT) {
      w_B[e] = to_f32<T>(sB[e]);
      w_C[e] = to_f32<T>(sC[e]);
      a_dB[e] = 0.f;
      a_dC[e] = 0.f;
    }
    if (tid == 0) bulk_wait_read0();  // previous tile's stores have finished reading the output tiles
    __syncthreads();
    if (tid == 0 && k + NS < ntiles) issue_tile(k + NS);  // raw stage s is free again

    // ---- (1) forward recompute inside the tile: h_t -> hbuf, y_t -> w_y
    for (int r = 0; r < rows; ++r) {
      const float dtv = w_dt[r * CH + c], uv = w_u[r * CH + c];
      const float dtu = dtv * uv;
      float y = Dc * uv;
      float* hb = hbuf + ((long)r * CH + c) * kN;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 Bq = *reinterpret_cast<const float4*>(w_B + r * kN + 4 * q);
        const float4 Cq = *reinterpret_cast<const float4*>(w_C + r * kN + 4 * q);
        const float Bv[4] = {Bq.x, Bq.y, Bq.z, Bq.w}, Cv[4] = {Cq.x, Cq.y, Cq.z, Cq.w};
        float hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = 4 * q + i;
          const float a = ex2_approx(dtv * A2[n]);
          h[n] = fmaf(a, h[n], dtu * Bv[i]);
          y = fmaf(h[n], Cv[i], y);
          hv[i] = h[n];
        }
        *reinterpret_cast<float4*>(hb + 4 * q) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
      w_y[r * CH + c] = y;
    }
    // h[] now holds the state at the END of the tile; reload the start state for the t == first step below
    float h0[kN];
    {
      const float4* cp = reinterpret_cast<const float4*>(p.ckpt + (((long)b * ntiles + tile) * p.D + c0 + c) * kN);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = cp[q];
        h0[4 * q] = v.x, h0[4 * q + 1] = v.y, h0[4 * q + 2] = v.z, h0[4 * q + 3] = v.w;
      }
    }

    // ---- (2) adjoint recurrence, backwards in time
    for (int r = rows - 1; r >= 0; --r) {
      const int e = r * CH + c;
      const float dtv = w_dt[e], uv = w_u[e], dy = w_dy[e];
      const float dtu = dtv * uv;
      float ddt = 0.f, du = dy * Dc;
      float red[2 * kN];  // [0,16): dB partials, [16,32): dC partials of this channel
      const float* hb = hbuf + ((long)r * CH + c) * kN;
      const float* hp = hbuf + ((long)(r - 1) * CH + c) * kN;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 Bq = *reinterpret_cast<const float4*>(w_B + r * kN + 4 * q);
        const float4 Cq = *reinterpret_cast<const float4*>(w_C + r * kN + 4 * q);
        const float4 hq = *reinterpret_cast<const float4*>(hb + 4 * q);
        float4 pq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r > 0) pq = *reinterpret_cast<const float4*>(hp + 4 * q);
        const float Bv[4] = {Bq.x, Bq.y, Bq.z, Bq.w}, Cv[4] = {Cq.x, Cq.y, Cq.z, Cq.w};
        const float hv[4] = {hq.x, hq.y, hq.z, hq.w};
        const float pv[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = 4 * q + i;
          const float hprev = (r > 0) ? pv[i] : h0[n];
          const float dhn = fmaf(dy, Cv[i], dh[n]);  // dL/dh_t
          const float a = ex2_approx(dtv * A2[n]);
          red[kN + n] = dy * hv[i];
          red[n] = dhn * dtu;
          const float tmp = dhn * hprev * a;          // dL/da * a
          ddt = fmaf(dhn * Bv[i], uv, fmaf(tmp, A[n], ddt));
          du = fmaf(dhn * dtv, Bv[i], du);
          dA[n] = fmaf(tmp, dtv, dA[n]);
          dh[n] = a * dhn;
        }
      }
      const float dd = ddt * w_sg[e];
      dbias += dd;
      dD = fmaf(dy, uv, dD);
      o_du[e] = from_f32<T>(du);
      o_dd[e] = from_f32<T>(dd);
      o_dz[e] = from_f32<T>(w_dzc[e] * w_y[e]);
      // reduce the 32 per-channel partials over the warp's 32 channels: lane l ends with the sum of red[l]
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
          const float send = up ? red[i] : red[i + o];
          const float keep = up ? red[i + o] : red[i];
          red[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      float* acc = (lane < kN) ? (a_dB + r * kN + lane) : (a_dC + r * kN + lane - kN);
      if (NT == 32)
        *acc += red[0];  // single warp per CTA: plain read-modify-write
      else
        atomicAdd(acc, red[0]);
    }

    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm.du, c0, t0, b, o_du);
      tma_store_3d(&tm.ddelta, c0, t0, b, o_dd);
      if (has_z) tma_store_3d(&tm.dz, c0, t0, b, o_dz);
      bulk_commit();
    }
    for (int e = tid; e < rows * kN; e += NT) {
      atomicAdd(p.dB + ((long)b * p.L + t0) * kN + e, a_dB[e]);
      atomicAdd(p.dC + ((long)b * p.L + t0) * kN + e, a_dC[e]);
    }
    __syncthreads();  // a_dB / a_dC / work arrays are rewritten by the next tile's pre-pass
  }
  if (tid == 0) bulk_wait0();
#pragma unroll
  for (int n = 0; n < kN; ++n) atomicAdd(p.dA + (long)(c0 + c) * kN + n, dA[n]);
  if (p.dD) atomicAdd(p.dD + c0 + c, dD);
  if (p.ddbias) atomicAdd(p.ddbias + c0 + c, dbias);
}

template <typename T, int CH>
int launch_bwd(const ScanBwdParams& p, int dtype, cudaStream_t stream) {
  using Cfg = BwdCfg<T, CH>;
  auto kern = selective_scan_bwd_kernel<Cfg, T>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, Cfg::SMEM, attr) != cudaSuccess) return check_launch("selective_scan_bwd attr");
  BwdTmaps tm;
  int rc;
  if ((rc = make_tmap_tokens(&tm.u, p.u, dtype, p.D, p.L, p.batch, p.ld_u, CH, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.delta, p.delta, dtype, p.D, p.L, p.batch, p.ld_delta, CH, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.dout, p.dout, dtype, p.D, p.L, p.batch, p.ld_dout, CH, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.B, p.Bm, dtype, kN, p.L, p.batch, p.ld_B, kN, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.C, p.Cm, dtype, kN, p.L, p.batch, p.ld_C, kN, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.du, p.du, dtype, p.D, p.L, p.batch, p.ld_du, CH, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.ddelta, p.ddelta, dtype, p.D, p.L, p.batch, p.ld_ddelta, CH, TT))) return rc;
  if (p.z) {
    if ((rc = make_tmap_tokens(&tm.z, p.z, dtype, p.D, p.L, p.batch, p.ld_z, CH, TT))) return rc;
    if ((rc = make_tmap_tokens(&tm.dz, p.dz, dtype, p.D, p.L, p.batch, p.ld_dz, CH, TT))) return rc;
  } else {
    tm.z = tm.u;
    tm.dz = tm.du;
  }
  kern<<<p.batch * (p.D / CH), Cfg::NT, Cfg::SMEM, stream>>>(tm, p);
  return check_launch("selective_scan_bwd");
}

}  // namespace

int selective_scan_bwd(const ScanBwdParams& p, int dtype, cudaStream_t stream) {
  const int es = dtype == 0 ? 4 : 2;
  SIM_REQUIRE(dtype == 0 || dtype == 1, SIM_ERR_INVALID, "selective_scan_bwd: dtype must be 0 (fp32) or 1 (bf16)");
  SIM_REQUIRE(p.batch > 0 && p.L > 0 && p.D > 0 && p.D % 32 == 0, SIM_ERR_INVALID,
              "selective_scan_bwd: D must be a positive multiple of 32 (got %d)", p.D);
  SIM_REQUIRE(p.u && p.delta && p.Bm && p.Cm && p.dout && p.A && p.ckpt && p.du && p.ddelta && p.dB && p.dC && p.dA,
              SIM_ERR_INVALID, "selective_scan_bwd: null tensor (checkpoints come from the training forward)");
  SIM_REQUIRE(!p.z || p.dz, SIM_ERR_INVALID, "selective_scan_bwd: z given without dz");
  const void* ptrs[] = {p.u, p.delta, p.z, p.Bm, p.Cm, p.dout, p.du, p.ddelta, p.dz};
  const long lds[] = {p.ld_u, p.ld_delta, p.ld_z, p.ld_B, p.ld_C, p.ld_dout, p.ld_du, p.ld_ddelta, p.ld_dz};
  for (int i = 0; i < 9; ++i) {
    if (!ptrs[i]) continue;
    SIM_REQUIRE(aligned16(ptrs[i]) && (lds[i] * es) % 16 == 0, SIM_ERR_ALIGN,
                "selective_scan_bwd: tensor %d needs a 16-byte aligned base and row stride (TMA tensor maps)", i);
  }
  return dtype == 0 ? launch_bwd<float, 32>(p, dtype, stream) : launch_bwd<__nv_bfloat16, 32>(p, dtype, stream);
}

}  // namespace sim
