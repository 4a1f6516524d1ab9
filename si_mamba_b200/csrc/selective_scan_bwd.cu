// Selective scan backward (training configs C2-C4), sm_100a.
//
// Replaces mamba-ssm's selective_scan_bwd as reached through Mamba.forward's autograd node
// (models/block.py:72).  Gradients of
//   dt = softplus(delta + bias); h_t = exp(dt_t A) h_{t-1} + dt_t B_t u_t; y_t = <h_t, C_t> + D u_t; out = y silu(z)
// w.r.t. u, delta, z (token-major, input dtype), B, C (fp32, summed over channels), A, D, bias (fp32).
//
// No (B, D, L, N) state tensor is stored: the training forward keeps only the state before every kScanCkpt-th (8th) step
// (ScanParams::ckpt, one fp32 tensor twice the size of an fp32 activation), which arrives with the operand tiles (TMA
// tensor copies + one bulk copy).  Per 8-step tile, walking the sequence backwards, the kernel recomputes the states
// sub-tile by sub-tile into registers and runs the adjoint recurrence dh_{t-1} = a_t dh_t straight from them (see BwdCfg).
// Thread = one channel x 16 / LPC states (default LPC = 2 lanes per channel, 8 states per thread): the partial sums over a
// thread's states (<h, C>, s1, s2) go to shared memory per lane and are added by the elementwise epilogue; dB / dC
// (reductions over channels) go through a transposed warp butterfly over the warp's 32 / LPC channels, per-warp
// shared-memory accumulators and one global fp32 atomic per (t, n) per CTA.
// History: r01 one thread per channel x 16 states with a state-history buffer 1742 us at the C2 bf16 layer shape -> register
// sub-tiles, 4 lanes per channel 878 us; r02 (profiles/r02_scan_bwd.md): ncu showed the l1tex data pipe (LDS + SHFL
// wavefronts) as the busiest unit, so 8 states per thread (half the B / C / (dt, u, dy) loads and butterfly shuffles per
// state update; needs 8-step tiles to keep 3 CTAs per SM) -> 675 us, state sums through shared memory instead of
// shuffles -> 629 us (C1 fp32 layer shape: 440 -> 312 us).
//
// Roofline class: l1tex data pipe / issue latency; HBM algorithmic bytes are (4 reads + 3 writes) * E * s + checkpoint
// 2 E * 4 + O(S).

#include <stdlib.h>

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

namespace {

constexpr int kN = 16;
constexpr int TT = kScanCkpt;  // steps per tile == checkpoint interval of the training forward

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

template <int N>
__device__ __forceinline__ void lds_vec_n(const float* p, float (&v)[N]) {
  if constexpr (N == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x, v[1] = t.y;
  } else {
    static_assert(N == 4, "lanes per channel");
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
}

template <typename T>
__device__ __forceinline__ void sts4(T* dst, float4 v);
template <>
__device__ __forceinline__ void sts4<float>(float* dst, float4 v) {
  *reinterpret_cast<float4*>(dst) = v;
}
template <>
__device__ __forceinline__ void sts4<__nv_bfloat16>(__nv_bfloat16* dst, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<const unsigned*>(&lo);
  r.y = *reinterpret_cast<const unsigned*>(&hi);
  *reinterpret_cast<uint2*>(dst) = r;
}

// State history without a (step x state) buffer.  The first cut kept every h_t of a tile in shared memory (32 KB per two
// warps): 6 warps per SM, ncu: 1.3 warps per scheduler, issue 30 %.  Now a tile is walked as TT / SUB sub-tiles of 4 steps,
// back to front: a forward sweep over the leading sub-tiles leaves their end states in shared memory, then each sub-tile is
// recomputed into REGISTERS (4 steps x the thread's states, plus the decay factors) and its adjoint steps run straight from
// them.  With the 8-step tiles: 1.5 exps per state-step, no history traffic.
constexpr int SUB = 4;  // steps per sub-tile

// CPT = channels per thread.  With CPT = 2 a thread owns two channels (CH / 2 apart) x 16 / LPC states (LPC = 4: the same eight state
// chains per thread as LPC = 2, CPT = 1): the B / C row slices it loads serve both channels (2 instead of 4 LDS.128 per step and
// phase), and the dB / dC partials of its two channels are added in registers BEFORE the warp butterfly, which then reduces
// 2 S = 8 values over 8 lane groups (7 shuffles) instead of 16 values over 16 (15 shuffles).
template <typename T, int CH_, int LPC_, bool SAVE_A_ = false, int NS_ = 2, int MINB_ = 1, int CPT_ = 1>
struct BwdCfg {
  static constexpr int CPT = CPT_;
  static constexpr bool SAVE_A = SAVE_A_;  // keep exp(dt A) of a recomputed sub-tile in registers for its adjoint steps
  static constexpr int MINB = MINB_;       // resident CTAs per SM the register allocation aims at
  static constexpr int CH = CH_;
  static constexpr int LPC = LPC_;       // lanes per channel
  static constexpr int S = kN / LPC_;    // states per thread
  static constexpr int NT = LPC_ * CH_ / CPT_;  // threads per CTA
  static constexpr int NW = NT / 32;
  // raw TMA stages.  The pre-pass consumes a stage completely, so ONE stage already lets the next tile's copies fly during
  // the whole recurrence phase; the second stage only costs shared memory (fp32: 3 -> 4 CTAs per SM, bf16: 4 -> 5)
  static constexpr int NS = NS_;
  static constexpr int RAW_MAIN = TT * CH_ * (int)sizeof(T);
  static constexpr int RAW_BC = TT * kN * (int)sizeof(T);
  static constexpr int RAW_CK = CH_ * kN * 4;                   // tile-start states of the CTA's channels (fp32)
  static constexpr int RAW_STAGE = 4 * RAW_MAIN + 2 * RAW_BC + RAW_CK;  // u, delta, z, dout, B, C, checkpoint
  static constexpr int OUT = 3 * RAW_MAIN;                      // du, ddelta, dz
  static constexpr int WORK = (6 + 3 * LPC_) * TT * CH_ * 4     // (dt, u, dy, dt*u) packed, sg, dzc; y, s1, s2 per lane of a channel
                              + 2 * TT * kN * 4                 // B, C fp32
                              + NW * TT * 2 * kN * 4;           // per-warp dB | dC tile sums
  static constexpr int SCK = (TT / SUB - 1) * (CPT_ * S / 4) * NT * 16;  // sub-tile start states, float4 planes
  static constexpr int SMEM = NS * RAW_STAGE + OUT + WORK + SCK + NS * 8 + 64;
  static_assert(RAW_MAIN % 128 == 0 && RAW_BC % 128 == 0 && RAW_CK % 128 == 0, "TMA tiles must stay 128-B aligned");
  static_assert(TT % SUB == 0 && NT % CH_ == 0, "whole sub-tiles; a thread's elementwise channel is fixed");
  static_assert(S % 4 == 0 && CH_ % CPT_ == 0, "float4 state groups; whole channel groups");
};

template <typename Cfg, typename T>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) selective_scan_bwd_kernel(const __grid_constant__ BwdTmaps tm,
                                                                     const ScanBwdParams p) {
  constexpr int CH = Cfg::CH, NT = Cfg::NT, NS = Cfg::NS, S = Cfg::S, NW = Cfg::NW, LPC = Cfg::LPC, SQ = Cfg::S / 4,
                CPT = Cfg::CPT;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;
  T* __restrict__ o_du = reinterpret_cast<T*>(smem + NS * Cfg::RAW_STAGE);
  T* __restrict__ o_dd = o_du + TT * CH;
  T* __restrict__ o_dz = o_dd + TT * CH;
  float4* __restrict__ w4 = reinterpret_cast<float4*>(smem + NS * Cfg::RAW_STAGE + Cfg::OUT);  // (dt, u, dy, dt*u)
  float* __restrict__ w_sg = reinterpret_cast<float*>(w4 + TT * CH);
  float* __restrict__ w_dzc = w_sg + TT * CH;
  float* __restrict__ w_y = w_dzc + TT * CH;
  float* __restrict__ w_s1 = w_y + LPC * TT * CH;   // y, s1, s2: [t][channel][lane of the channel] partial sums over the
  float* __restrict__ w_s2 = w_s1 + LPC * TT * CH;  // lane's states, added up by the epilogue (no shuffle in the recurrence)
  float* __restrict__ w_B = w_s2 + LPC * TT * CH;
  float* __restrict__ w_C = w_B + TT * kN;
  float* __restrict__ a_dBC = w_C + TT * kN;                                  // [warp][t][dB(16) | dC(16)]
  float4* __restrict__ sck = reinterpret_cast<float4*>(a_dBC + NW * TT * 2 * kN);  // [sub-1][half][thread]
  uint64_t* full = reinterpret_cast<uint64_t*>(sck + (TT / SUB - 1) * CPT * SQ * NT);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunk = p.D / CH;
  const int b = blockIdx.x / nchunk;
  const int c0 = (blockIdx.x % nchunk) * CH;
  const int sub = tid % LPC;  // which slice of the 16 states
  const int c = tid / LPC;  // first of this thread's CPT channels within the CTA: c, c + CH / CPT, ... (the stride keeps
  constexpr int CSTR = CH / CPT;  // every per-(step, channel) shared-memory access of a warp on distinct banks)
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
    mbar_arrive_expect_tx(&full[s], (has_z ? 4u : 3u) * Cfg::RAW_MAIN + 2u * Cfg::RAW_BC + Cfg::RAW_CK);
    bulk_g2s(st + 4 * Cfg::RAW_MAIN + 2 * Cfg::RAW_BC, p.ckpt + (((long)b * ntiles + (ntiles - 1 - k)) * p.D + c0) * kN,
             Cfg::RAW_CK, &full[s]);
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

  float A[CPT][S], A2[CPT][S], dA[CPT][S], dh[CPT][S];
#pragma unroll
  for (int j = 0; j < CPT; ++j)
#pragma unroll
    for (int n = 0; n < S; ++n) {
      A[j][n] = p.A[(long)(c0 + c + j * CSTR) * kN + sub * S + n];
      A2[j][n] = A[j][n] * kLog2e;
      dA[j][n] = 0.f;
      dh[j][n] = 0.f;
    }
  const int ce = tid % CH;  // channel of this thread's elements in the elementwise passes (NT % CH == 0)
  const float De = p.Dv ? p.Dv[c0 + ce] : 0.f;
  const float bias_e = p.dbias ? p.dbias[c0 + ce] : 0.f;
  float dD = 0.f, dbias = 0.f;
  // lane (sub, channel-in-warp cw) ends the dB / dC butterfly with the total of value v = cw (the warp has 32 / LPC
  // channels and every thread 2 S = 32 / LPC values): v < S -> dB of state sub*S + v, else dC of state sub*S + v - S
  const int vfin = lane / LPC;
  float* my_acc = a_dBC + warp * TT * 2 * kN + (vfin < S ? sub * S + vfin : kN + sub * S + vfin - S);

  // one forward step of this thread's 8 states: hn = a * hp + dt*u*B; returns the thread's share of <h, C>
  auto fwd_step = [&](int j_, int r, const float (&hp)[S], float (&hn)[S], float* aout = nullptr) -> float {
    const float4 w = w4[r * CH + c + j_ * CSTR];
    const float dtv = w.x, dtu = w.w;
    float2 y2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < S / 4; ++q) {
      const float4 Bq = *reinterpret_cast<const float4*>(w_B + r * kN + sub * S + 4 * q);
      const float4 Cq = *reinterpret_cast<const float4*>(w_C + r * kN + sub * S + 4 * q);
      const float Bv[4] = {Bq.x, Bq.y, Bq.z, Bq.w}, Cv[4] = {Cq.x, Cq.y, Cq.z, Cq.w};
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        const int n = 4 * q + i;
        const float2 x = __fmul2_rn(make_float2(dtv, dtv), make_float2(A2[j_][n], A2[j_][n + 1]));
        const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
        const float2 bu = __fmul2_rn(make_float2(dtu, dtu), make_float2(Bv[i], Bv[i + 1]));
        const float2 h2 = __ffma2_rn(a, make_float2(hp[n], hp[n + 1]), bu);
        if (aout) aout[n] = a.x, aout[n + 1] = a.y;  // kept for the adjoint step (SAVE_A)
        hn[n] = h2.x, hn[n + 1] = h2.y;
        y2 = __ffma2_rn(h2, make_float2(Cv[i], Cv[i + 1]), y2);
      }
    }
    return y2.x + y2.y;
  };

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

    const float* sck0 = reinterpret_cast<const float*>(st + 4 * Cfg::RAW_MAIN + 2 * Cfg::RAW_BC);  // [channel][16]

    mbar_wait(&full[s], (k / NS) & 1);

    // ---- pre-pass (elementwise over the tile): activations and their derivatives.  Rows past the end of the sequence
    // were zero-filled by the TMA unit: their dout is 0, so every gradient they produce is 0 and no step needs a guard.
    // The softplus / sigmoid chains are serial MUFU chains (ex2 -> lg2, ex2 -> rcp): with one element per loop iteration
    // the pass was latency-bound (ncu: 18 % of the kernel's samples).  The trip count is a compile-time constant, so the
    // iterations are unrolled and their chains interleave; element e = tid + i NT keeps the accesses conflict-free.
    static_assert((TT * CH) % NT == 0, "whole elementwise passes");
#pragma unroll
    for (int i = 0; i < TT * CH / NT; ++i) {
      const int e = tid + i * NT;
      const float x = to_f32<T>(sd[e]) + bias_e;
      const float dtv = p.softplus ? softplus_f(x) : x;
      w_sg[e] = p.softplus ? ((x > 20.f) ? 1.f : sigmoid_f(x)) : 1.f;  // d softplus / dx
      const float uv = to_f32<T>(su[e]);
      const float go = to_f32<T>(so[e]);
      float dy = go;
      if (has_z) {
        const float zv = to_f32<T>(sz[e]);
        const float sg = sigmoid_f(zv);
        dy = go * zv * sg;                                    // dL/dy = dout * silu(z)
        w_dzc[e] = go * sg * (1.f + zv * (1.f - sg));         // dout * silu'(z); dz = this * y
      } else {
        w_dzc[e] = 0.f;
      }
      w4[e] = make_float4(dtv, uv, dy, dtv * uv);
    }
    for (int e = tid; e < TT * kN; e += NT) {
      w_B[e] = to_f32<T>(sB[e]);
      w_C[e] = to_f32<T>(sC[e]);
    }
    // state at the start of this tile (from the training forward; arrived with the operand tiles)
    float h0[CPT][S];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
      for (int q = 0; q < SQ; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(sck0 + (c + j * CSTR) * kN + sub * S + 4 * q);
        h0[j][4 * q] = v.x, h0[j][4 * q + 1] = v.y, h0[j][4 * q + 2] = v.z, h0[j][4 * q + 3] = v.w;
      }
    if (tid == 0) bulk_wait_read0();  // previous tile's stores have finished reading the output tiles
    __syncthreads();
    if (tid == 0 && k + NS < ntiles) issue_tile(k + NS);  // raw stage s is free again

    // ---- (1) forward sweep over steps 0 .. TT-SUB-1: leaves the start state of sub-tiles 1 .. 3 in shared memory
    {
      float h[CPT][S];
#pragma unroll
      for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int n = 0; n < S; ++n) h[j][n] = h0[j][n];
#pragma unroll 1
      for (int q = 0; q < TT / SUB - 1; ++q) {
#pragma unroll
        for (int i = 0; i < SUB; ++i)
#pragma unroll
          for (int j = 0; j < CPT; ++j) fwd_step(j, SUB * q + i, h[j], h[j]);
#pragma unroll
        for (int j = 0; j < CPT; ++j)
#pragma unroll
          for (int g = 0; g < SQ; ++g)
            sck[((q * CPT + j) * SQ + g) * NT + tid] = make_float4(h[j][4 * g], h[j][4 * g + 1], h[j][4 * g + 2], h[j][4 * g + 3]);
      }
    }

    // ---- (2) sub-tiles back to front: recompute 4 states into registers, then their adjoint steps
#pragma unroll 1
    for (int sb = TT / SUB - 1; sb >= 0; --sb) {
      const int rb = SUB * sb;
      float hs[CPT][S], hq[CPT][SUB][S];
      [[maybe_unused]] float aq[CPT][Cfg::SAVE_A ? SUB : 1][S];  // exp(dt A) of the sub-tile's steps
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        if (sb == 0) {
#pragma unroll
          for (int n = 0; n < S; ++n) hs[j][n] = h0[j][n];
        } else {
#pragma unroll
          for (int g = 0; g < SQ; ++g) {
            const float4 pj = sck[(((sb - 1) * CPT + j) * SQ + g) * NT + tid];
            hs[j][4 * g] = pj.x, hs[j][4 * g + 1] = pj.y, hs[j][4 * g + 2] = pj.z, hs[j][4 * g + 3] = pj.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < SUB; ++i) {
        const int r = rb + i;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          float* ao = Cfg::SAVE_A ? aq[j][Cfg::SAVE_A ? i : 0] : nullptr;
          float y = (i == 0) ? fwd_step(j, r, hs[j], hq[j][0], ao) : fwd_step(j, r, hq[j][i > 0 ? i - 1 : 0], hq[j][i], ao);
          w_y[(r * CH + c + j * CSTR) * LPC + sub] = y;  // partial <h, C>; D u is added in the epilogue
        }
      }
#pragma unroll
      for (int i = SUB - 1; i >= 0; --i) {
        const int r = rb + i;
        float red[2 * S];  // [0,S): dB partials, [S,2S): dC partials of this state slice, summed over the thread's channels
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float4 w = w4[r * CH + c + j * CSTR];
          const float dtv = w.x, dy = w.z, dtu = w.w;
          // packed f32x2 over state pairs.  s1 = sum_n dh_n B_n feeds both ddt (x u) and du (x dt); s2 = sum_n tmp_n A_n
          float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
          const float2 dy2 = make_float2(dy, dy), dt2 = make_float2(dtv, dtv), dtu2 = make_float2(dtu, dtu);
#pragma unroll
          for (int q = 0; q < S / 4; ++q) {
            const float4 Bq = *reinterpret_cast<const float4*>(w_B + r * kN + sub * S + 4 * q);
            const float4 Cq = *reinterpret_cast<const float4*>(w_C + r * kN + sub * S + 4 * q);
            const float Bv[4] = {Bq.x, Bq.y, Bq.z, Bq.w}, Cv[4] = {Cq.x, Cq.y, Cq.z, Cq.w};
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {
              const int n = 4 * q + jj;
              const float2 hprev = (i == 0) ? make_float2(hs[j][n], hs[j][n + 1])
                                            : make_float2(hq[j][i > 0 ? i - 1 : 0][n], hq[j][i > 0 ? i - 1 : 0][n + 1]);
              const float2 dhn = __ffma2_rn(dy2, make_float2(Cv[jj], Cv[jj + 1]), make_float2(dh[j][n], dh[j][n + 1]));  // dL/dh_t
              float2 a;
              if constexpr (Cfg::SAVE_A) {
                a = make_float2(aq[j][Cfg::SAVE_A ? i : 0][n], aq[j][Cfg::SAVE_A ? i : 0][n + 1]);
              } else {
                const float2 x = __fmul2_rn(dt2, make_float2(A2[j][n], A2[j][n + 1]));
                a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
              }
              const float2 hq2 = make_float2(hq[j][i][n], hq[j][i][n + 1]);
              if (j == 0) {
                const float2 rc = __fmul2_rn(dy2, hq2);
                const float2 rb2 = __fmul2_rn(dhn, dtu2);
                red[S + n] = rc.x, red[S + n + 1] = rc.y;
                red[n] = rb2.x, red[n + 1] = rb2.y;
              } else {  // the second channel's partials join the first's before the butterfly
                const float2 rc = __ffma2_rn(dy2, hq2, make_float2(red[S + n], red[S + n + 1]));
                const float2 rb2 = __ffma2_rn(dhn, dtu2, make_float2(red[n], red[n + 1]));
                red[S + n] = rc.x, red[S + n + 1] = rc.y;
                red[n] = rb2.x, red[n + 1] = rb2.y;
              }
              const float2 dhp = __fmul2_rn(a, dhn);                    // dL/dh_{t-1} through the decay
              const float2 tmp = __fmul2_rn(dhp, hprev);                // dL/da * a
              s1 = __ffma2_rn(dhn, make_float2(Bv[jj], Bv[jj + 1]), s1);
              s2 = __ffma2_rn(tmp, make_float2(A[j][n], A[j][n + 1]), s2);
              const float2 dAn = __ffma2_rn(tmp, dt2, make_float2(dA[j][n], dA[j][n + 1]));
              dA[j][n] = dAn.x, dA[j][n + 1] = dAn.y;
              dh[j][n] = dhp.x, dh[j][n + 1] = dhp.y;
            }
          }
          w_s1[(r * CH + c + j * CSTR) * LPC + sub] = s1.x + s1.y;
          w_s2[(r * CH + c + j * CSTR) * LPC + sub] = s2.x + s2.y;
        }
        // reduce the 2 S partials over the warp's 32 / LPC lane groups: transposed butterfly over the group lane bits
#pragma unroll
        for (int ov = S, ol = 16; ov >= 1; ov >>= 1, ol >>= 1) {
          const bool up = (lane & ol) != 0;
#pragma unroll
          for (int v = 0; v < ov; ++v) {
            const float send = up ? red[v] : red[v + ov];
            const float keep = up ? red[v + ov] : red[v];
            red[v] = keep + __shfl_xor_sync(0xffffffffu, send, ol);
          }
        }
        my_acc[r * 2 * kN] = red[0];  // this warp's own accumulator row: no atomics
      }
    }
    __syncthreads();

    // ---- epilogue (elementwise, every lane busy): ddelta = (u s1 + s2) softplus', du = dy D + dt s1, dz = dzc (y + D u)
#pragma unroll
    for (int i = 0; i < TT * CH / NT; ++i) {
      const int e = tid + i * NT;
      const float4 w = w4[e];
      float lp[3][LPC];
      lds_vec_n<LPC>(w_y + e * LPC, lp[0]);
      lds_vec_n<LPC>(w_s1 + e * LPC, lp[1]);
      lds_vec_n<LPC>(w_s2 + e * LPC, lp[2]);
      float yv = lp[0][0], s1v = lp[1][0], s2v = lp[2][0];
#pragma unroll
      for (int l = 1; l < LPC; ++l) yv += lp[0][l], s1v += lp[1][l], s2v += lp[2][l];
      const float dd = fmaf(w.y, s1v, s2v) * w_sg[e];
      dbias += dd;
      dD = fmaf(w.z, w.y, dD);
      o_du[e] = from_f32<T>(fmaf(w.z, De, w.x * s1v));
      o_dd[e] = from_f32<T>(dd);
      o_dz[e] = from_f32<T>(w_dzc[e] * fmaf(De, w.y, yv));
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm.du, c0, t0, b, o_du);
      tma_store_3d(&tm.ddelta, c0, t0, b, o_dd);
      if (has_z) tma_store_3d(&tm.dz, c0, t0, b, o_dz);
      bulk_commit();
    }
    for (int e = tid; e < rows * 2 * kN; e += NT) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) v += a_dBC[w * TT * 2 * kN + e];
      const int r = e / (2 * kN), n = e % (2 * kN);
      float* dst = (n < kN ? p.dB : p.dC) + ((long)b * p.L + t0 + r) * kN + (n & (kN - 1));
      atomicAdd(dst, v);
    }
    __syncthreads();  // accumulators / work arrays are rewritten by the next tile's pre-pass
  }
  if (tid == 0) bulk_wait0();
#pragma unroll
  for (int j = 0; j < CPT; ++j)
#pragma unroll
    for (int n = 0; n < S; ++n) atomicAdd(p.dA + (long)(c0 + c + j * CSTR) * kN + sub * S + n, dA[j][n]);
  if (p.dD) atomicAdd(p.dD + c0 + ce, dD);
  if (p.ddbias) atomicAdd(p.ddbias + c0 + ce, dbias);
}

template <typename T, int CH, int LPC, bool SAVE_A = false, int NS = 2, int MINB = 1, int CPT = 1>
int launch_bwd(const ScanBwdParams& p, int dtype, cudaStream_t stream) {
  using Cfg = BwdCfg<T, CH, LPC, SAVE_A, NS, MINB, CPT>;
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
  // Configurations (SIM_SCAN_BWD_CFG overrides the default for tools/kernel_bench.py; r02 numbers at the C1 fp32 / C2 bf16
  // layer shapes, profiles/r02_scan_bwd.md):
  //   22 (default): 64-channel CTAs, 2 lanes per channel (8 states per thread), 128 threads, 3 CTAs per SM: 312 / 629 us
  //   42: 32-channel CTAs, 4 lanes per channel (the r01 mapping, twice the operand loads and shuffles per state update): 408 / 792 us
  //   2200: as 22 with exp(dt A) recomputed in the adjoint (120 registers, 4 CTAs per SM): 332 / 666 us
  static const int cfg = [] { const char* e = getenv("SIM_SCAN_BWD_CFG"); return e ? atoi(e) : 22; }();
  if (cfg == 42 || p.D % 64 != 0)
    return dtype == 0 ? launch_bwd<float, 32, 4, true, 1>(p, dtype, stream) : launch_bwd<__nv_bfloat16, 32, 4, true, 2>(p, dtype, stream);
  if (cfg == 24)  // 64-channel CTAs, 4 lanes per channel pair (2 channels x 4 states per thread), 128 threads
    return dtype == 0 ? launch_bwd<float, 64, 4, true, 1, 3, 2>(p, dtype, stream) : launch_bwd<__nv_bfloat16, 64, 4, true, 1, 3, 2>(p, dtype, stream);
  if (cfg == 2200)
    return dtype == 0 ? launch_bwd<float, 64, 2, false, 1, 4>(p, dtype, stream) : launch_bwd<__nv_bfloat16, 64, 2, false, 1, 4>(p, dtype, stream);
  return dtype == 0 ? launch_bwd<float, 64, 2, true, 1, 3>(p, dtype, stream) : launch_bwd<__nv_bfloat16, 64, 2, true, 1, 3>(p, dtype, stream);
}

}  // namespace sim
