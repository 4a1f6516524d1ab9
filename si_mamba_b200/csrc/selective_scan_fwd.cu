// Selective scan forward, hand-written for sm_100a.
//
// Replaces mamba-ssm's selective_scan_fwd CUDA kernel as SI-Mamba reaches it through
// Mamba.forward (models/block.py:72 -> mamba_ssm selective_scan_fn).  Semantics
// are those of mamba-ssm's ``selective_scan_ref`` (restated in oracle/mamba.py):
//   dt = softplus(delta + delta_bias)
//   h_t[d,n] = exp(dt_t[d] * A[d,n]) * h_{t-1}[d,n] + dt_t[d] * B_t[n] * u_t[d]
//   y_t[d]   = sum_n h_t[d,n] * C_t[n] + D[d] * u_t[d];   out = y * silu(z)
//
// HBM layout (B200-first, differs from mamba-ssm's channel-major (B,D,L)):
// TOKEN-major.  u / delta / z / out are (batch*L, D) row-major with an explicit
// row stride, so x and z are read in place as column slices of the in_proj
// output and B / C as column slices of the x_proj output ((batch*L, 56)):
// no transposes, no chunk copies.
//
// Work decomposition: one CTA owns CH consecutive channels of one cloud and
// walks all L steps.  The recurrence is evaluated sequentially in time by the
// thread that owns (channel, S of the 16 states) - 4 FMA-pipe ops + 1 MUFU.EX2
// per state update, the minimum (B200 measures 16 MUFU/clk/SM, which with HBM
// at ~6.5 TB/s makes the fp32 scan MUFU-bound at ~0.7 of the HBM roofline; see
// DESIGN.md "scan cost model").  The 16/S partial sums of a channel are combined
// with a transposed butterfly so each lane of the group finishes (and writes) a
// different time step.  Time tiles of TT steps arrive as ONE TMA tensor copy per
// operand (cp.async.bulk.tensor + mbarrier, NS-deep ring); softplus / silu are
// applied by all threads in a balanced elementwise pre-pass on the staged tile;
// results leave through a shared-memory tile and one TMA tensor store.
//
// Roofline: HBM.  Algorithmic bytes per launch = (3 reads + 1 write) * B*L*D*s
// + 2 * B*L*16*s (s = bytes per element); see DESIGN.md.

#include "scan_common.cuh"

namespace sim {

// S states x CPT channels per thread: every step a thread needs 2*S words of B/C and 2*CPT words of dt/u from
// shared memory for S*CPT state updates, i.e. 2/CPT + 2/S shared-memory wavefronts per warp-update.  ncu on the
// first version (S=4, CPT=1: 2.5 + epilogue) showed the LSU shared pipe at 73% with MUFU at 46%, so B/C reuse
// across channels (CPT > 1) is what moves the kernel towards its MUFU bound.
template <typename T, int S_, int CPT_, int CH_, int TT_, int NS_, int DBG_ = 0>
struct ScanCfg {
  static constexpr int DBG = DBG_;  // bench-only ablation switches (results are WRONG when non-zero): 1 no MUFU in the
                                    // recurrence, 2 no activation math in the pre-pass, 4 no butterfly, 8 no B/C loads
  static constexpr int S = S_;                  // states per thread
  static constexpr int LPC = kNState / S_;      // lanes per channel group
  static constexpr int CPT = CPT_;              // adjacent channels per thread
  static constexpr int CH = CH_;                // channels per CTA
  static constexpr int TT = TT_;                // time steps per tile
  static_assert(TT_ == kScanTile, "checkpoint interval is fixed by kScanTile");
  static constexpr int NS = NS_;                // raw stages in flight
  static constexpr int NT = (CH_ / CPT_) * LPC; // threads per CTA
  static constexpr int CHP = CH_ + 8;           // padded row stride (floats) of the work arrays
  static constexpr int RAW_MAIN = TT_ * CH_ * (int)sizeof(T);
  static constexpr int RAW_BC = TT_ * kNState * (int)sizeof(T);
  static constexpr int RAW_STAGE = 3 * RAW_MAIN + 2 * RAW_BC;
  static constexpr int OBUF = TT_ * CH_ * (int)sizeof(T);
  static constexpr int WORK = 3 * TT_ * CHP * 4 + 2 * TT_ * kNState * 4 + CH_ * 4;
  static constexpr int SMEM = NS_ * RAW_STAGE + OBUF + WORK + NS_ * 8 + 16;
  static_assert(RAW_STAGE % 128 == 0 && RAW_MAIN % 128 == 0 && RAW_BC % 128 == 0 && OBUF % 128 == 0,
                "TMA tiles must stay 128-B aligned");
  static_assert(NT % 32 == 0 && (CH_ & (CH_ - 1)) == 0, "whole warps, power-of-two channel tile");
};

template <typename Cfg, typename T>
__global__ void __launch_bounds__(Cfg::NT) selective_scan_fwd_kernel(const __grid_constant__ ScanTmaps tm,
                                                                     const ScanParams p) {
  constexpr int S = Cfg::S, LPC = Cfg::LPC, CPT = Cfg::CPT, CH = Cfg::CH, CHP = Cfg::CHP, TT = Cfg::TT,
                NS = Cfg::NS, NT = Cfg::NT;
  // NOTE: derive every pointer from the extern array itself (no integer round-trips), otherwise the
  // compiler loses the shared address space and emits generic LD/ST instead of LDS/STS.
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;
  T* __restrict__ obuf = reinterpret_cast<T*>(smem + NS * Cfg::RAW_STAGE);
  float* __restrict__ w_dt = reinterpret_cast<float*>(smem + NS * Cfg::RAW_STAGE + Cfg::OBUF);
  float* __restrict__ w_u = w_dt + TT * CHP;
  float* __restrict__ w_g = w_u + TT * CHP;
  float* __restrict__ w_B = w_g + TT * CHP;
  float* __restrict__ w_C = w_B + TT * kNState;
  float* __restrict__ s_bias = w_C + TT * kNState;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_bias + CH);

  const int tid = threadIdx.x;
  const int nchunk = p.D / CH;
  const int b = blockIdx.x / nchunk;
  const int c0 = (blockIdx.x % nchunk) * CH;
  const int sub = tid % LPC;          // which S-state slice
  const int ch0 = (tid / LPC) * CPT;  // first of this thread's CPT channels (within the CTA)
  const int ntiles = (p.L + TT - 1) / TT;
  const int nck = (p.L + kScanCkpt - 1) / kScanCkpt;
  const bool has_z = p.z != nullptr;

  if (tid == 0) {
    tma_prefetch_desc(&tm.u);
    tma_prefetch_desc(&tm.delta);
    tma_prefetch_desc(&tm.B);
    tma_prefetch_desc(&tm.C);
    tma_prefetch_desc(&tm.out);
    if (has_z) tma_prefetch_desc(&tm.z);
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < CH; i += NT) s_bias[i] = p.dbias ? p.dbias[c0 + i] : 0.f;
  __syncthreads();

  // producer (one thread): one tensor copy per operand per tile; rows past L are zero-filled by the TMA unit
  auto issue_tile = [&](int tile) {
    const int s = tile % NS;
    const int t0 = tile * TT;
    unsigned char* st = raw + s * Cfg::RAW_STAGE;
    mbar_arrive_expect_tx(&full[s], (has_z ? 3u : 2u) * Cfg::RAW_MAIN + 2u * Cfg::RAW_BC);
    tma_load_3d(st, &tm.u, c0, t0, b, &full[s]);
    tma_load_3d(st + Cfg::RAW_MAIN, &tm.delta, c0, t0, b, &full[s]);
    if (has_z) tma_load_3d(st + 2 * Cfg::RAW_MAIN, &tm.z, c0, t0, b, &full[s]);
    tma_load_3d(st + 3 * Cfg::RAW_MAIN, &tm.B, 0, t0, b, &full[s]);
    tma_load_3d(st + 3 * Cfg::RAW_MAIN + Cfg::RAW_BC, &tm.C, 0, t0, b, &full[s]);
  };
  if (tid == 0) {
    for (int k = 0; k < NS && k < ntiles; ++k) issue_tile(k);
  }

  // per-thread constants: A pre-scaled by log2(e) so exp(dt*A) = ex2(dt*A2)
  float2 A2[CPT][S / 2];
  float2 h[CPT][S / 2];
  float Dc[CPT];
#pragma unroll
  for (int cp = 0; cp < CPT; ++cp) {
#pragma unroll
    for (int j = 0; j < S / 2; ++j) {
      const float* Ap = p.A + (long)(c0 + ch0 + cp) * kNState + sub * S + 2 * j;
      A2[cp][j] = make_float2(Ap[0] * kLog2e, Ap[1] * kLog2e);
      h[cp][j] = make_float2(0.f, 0.f);
    }
    Dc[cp] = p.Dv ? p.Dv[c0 + ch0 + cp] : 0.f;
  }

  for (int tile = 0; tile < ntiles; ++tile) {
    const int s = tile % NS;
    const int t0 = tile * TT;
    unsigned char* st = raw + s * Cfg::RAW_STAGE;
    const T* su = reinterpret_cast<const T*>(st);
    const T* sd = reinterpret_cast<const T*>(st + Cfg::RAW_MAIN);
    const T* sz = reinterpret_cast<const T*>(st + 2 * Cfg::RAW_MAIN);
    const T* sB = reinterpret_cast<const T*>(st + 3 * Cfg::RAW_MAIN);
    const T* sC = reinterpret_cast<const T*>(st + 3 * Cfg::RAW_MAIN + Cfg::RAW_BC);

    // training forward: state before every kScanCkpt-th step, layout (batch, ceil(L / kScanCkpt), D, 16)
    auto save_ckpt = [&](int idx) {
#pragma unroll
      for (int cp = 0; cp < CPT; ++cp) {
        float2* dst = reinterpret_cast<float2*>(p.ckpt + (((long)b * nck + idx) * p.D + c0 + ch0 + cp) * kNState + sub * S);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) dst[j] = h[cp][j];
      }
    };
    if (p.ckpt) save_ckpt(tile * (TT / kScanCkpt));
    mbar_wait(&full[s], (tile / NS) & 1);

    // ---- pre-pass: softplus(delta + bias), silu(z), widen to fp32.  Always the whole tile: rows past L were
    // zero-filled by the TMA unit, their results are clipped by the TMA store and nothing after them is used.
    // Four adjacent channels per thread and iteration (128-bit LDS / STS); ncu showed this pass at ~35% of all
    // issued instructions when done element-wise.
#pragma unroll 2
    for (int g = tid; g < TT * CH / 4; g += NT) {
      const int r = g / (CH / 4), cc = (g % (CH / 4)) * 4;
      const float4 bs = *reinterpret_cast<const float4*>(s_bias + cc);
      float4 dv = lds4<T>(sd + r * CH + cc);
      dv.x += bs.x, dv.y += bs.y, dv.z += bs.z, dv.w += bs.w;
      if (p.softplus && !(Cfg::DBG & 2)) dv = make_float4(softplus_f(dv.x), softplus_f(dv.y), softplus_f(dv.z), softplus_f(dv.w));
      *reinterpret_cast<float4*>(w_dt + r * CHP + cc) = dv;
      *reinterpret_cast<float4*>(w_u + r * CHP + cc) = lds4<T>(su + r * CH + cc);
      float4 gv = make_float4(1.f, 1.f, 1.f, 1.f);
      if (has_z && !(Cfg::DBG & 2)) {
        const float4 zv = lds4<T>(sz + r * CH + cc);
        gv = p.z_gate ? zv : make_float4(silu_f(zv.x), silu_f(zv.y), silu_f(zv.z), silu_f(zv.w));
      }
      *reinterpret_cast<float4*>(w_g + r * CHP + cc) = gv;
    }
    for (int g = tid; g < TT * kNState / 4; g += NT) {
      *reinterpret_cast<float4*>(w_B + 4 * g) = lds4<T>(sB + 4 * g);
      *reinterpret_cast<float4*>(w_C + 4 * g) = lds4<T>(sC + 4 * g);
    }
    // the previous tile's TMA store must have finished reading obuf before it is rewritten
    if (tid == 0) bulk_wait_read0();
    __syncthreads();

    // raw stage s is free again: refill it with tile + NS
    if (tid == 0 && tile + NS < ntiles) issue_tile(tile + NS);

    // ---- the recurrence over this tile, fully unrolled and software-pipelined by hand.
    // Warps issue in order, so a consumer stalled on a MUFU result also blocks the warp's NEXT MUFUs; with only
    // 2-3 warps per scheduler at B=32 the XU queue then drains (ncu: XU 62 % busy, per-step critical path ~370
    // cycles).  Stage A of step t+PD (LDS dt/u, FMUL2, MUFU.EX2) is therefore issued ahead of stage B of step t
    // (h update, <h,C> on two accumulators), and B/C operands are fetched one step ahead of their use.
    if constexpr ((Cfg::DBG & 16) != 0) {  // ablation: data movement only
      for (int e = tid; e < TT * CH; e += NT) obuf[e] = from_f32<T>(w_u[(e / CH) * CHP + e % CH] + w_dt[(e / CH) * CHP + e % CH] + w_g[(e / CH) * CHP + e % CH]);
    } else {
    constexpr int PD = 2;
    float2 ar[PD + 1][CPT][S / 2];
    float dtu_r[PD + 1][CPT];
    auto stage_a = [&](int t) {
      const int slot = t % (PD + 1);
      float dtv[CPT], uv[CPT];
      lds_vec<CPT>(w_dt + t * CHP + ch0, dtv);
      lds_vec<CPT>(w_u + t * CHP + ch0, uv);
#pragma unroll
      for (int cp = 0; cp < CPT; ++cp) {
        dtu_r[slot][cp] = dtv[cp] * uv[cp];
        const float2 dt2 = make_float2(dtv[cp], dtv[cp]);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
          const float2 x = __fmul2_rn(dt2, A2[cp][j]);
          ar[slot][cp][j] = (Cfg::DBG & 1) ? make_float2(x.x + 1.f, x.y + 1.f) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
        }
      }
    };
#pragma unroll
    for (int t = 0; t < PD; ++t) stage_a(t);
    float Bc[S], Cc[S], Bn[S], Cn[S];
    lds_vec<S>(w_B + sub * S, Bc);
    lds_vec<S>(w_C + sub * S, Cc);
    float part[LPC][CPT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      if (t > 0 && t % kScanCkpt == 0) {
        if (p.ckpt && tile * (TT / kScanCkpt) + t / kScanCkpt < nck) save_ckpt(tile * (TT / kScanCkpt) + t / kScanCkpt);
      }
      if (t + PD < TT) stage_a(t + PD);
      if (t + 1 < TT && !(Cfg::DBG & 8)) {
        lds_vec<S>(w_B + (t + 1) * kNState + sub * S, Bn);
        lds_vec<S>(w_C + (t + 1) * kNState + sub * S, Cn);
      }
      const int slot = t % (PD + 1);
#pragma unroll
      for (int cp = 0; cp < CPT; ++cp) {
        const float2 dtu2 = make_float2(dtu_r[slot][cp], dtu_r[slot][cp]);
        float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
          const float2 bu = __fmul2_rn(dtu2, make_float2(Bc[2 * j], Bc[2 * j + 1]));
          h[cp][j] = __ffma2_rn(ar[slot][cp][j], h[cp][j], bu);
          acc[j & 1] = __ffma2_rn(h[cp][j], make_float2(Cc[2 * j], Cc[2 * j + 1]), acc[j & 1]);
        }
        part[t % LPC][cp] = (acc[0].x + acc[1].x) + (acc[0].y + acc[1].y);
      }
      if ((t + 1) % LPC == 0) {
        // transposed butterfly: lane `sub` ends with the full sums of step t + 1 - LPC + sub, and writes it
#pragma unroll
        for (int o = (Cfg::DBG & 4) ? 0 : LPC / 2; o >= 1; o >>= 1) {
          const bool up = (sub & o) != 0;
#pragma unroll
          for (int i = 0; i < o; ++i) {
#pragma unroll
            for (int cp = 0; cp < CPT; ++cp) {
              const float send = up ? part[i][cp] : part[i + o][cp];
              const float keep = up ? part[i + o][cp] : part[i][cp];
              part[i][cp] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
        }
        const int r = t + 1 - LPC + sub;
        float uw[CPT], gw[CPT];
        lds_vec<CPT>(w_u + r * CHP + ch0, uw);
        lds_vec<CPT>(w_g + r * CHP + ch0, gw);
#pragma unroll
        for (int cp = 0; cp < CPT; ++cp)
          obuf[r * CH + ch0 + cp] = from_f32<T>((part[0][cp] + Dc[cp] * uw[cp]) * gw[cp]);
      }
      if (t + 1 < TT && !(Cfg::DBG & 8)) {
#pragma unroll
        for (int i = 0; i < S; ++i) Bc[i] = Bn[i], Cc[i] = Cn[i];
      }
    }
    }  // DBG & 16


    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm.out, c0, t0, b, obuf);  // rows past L are clipped by the TMA unit
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait0();
}

template <typename T, int S, int CPT, int CH, int TT, int NS, int DBG = 0>
static int launch_scan(const ScanParams& p, int dtype, cudaStream_t stream) {
  using Cfg = ScanCfg<T, S, CPT, CH, TT, NS, DBG>;
  auto kern = selective_scan_fwd_kernel<Cfg, T>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, Cfg::SMEM, attr) != cudaSuccess) return check_launch("selective_scan_fwd attr");
  ScanTmaps tm;
  int rc;
  if ((rc = make_tmap_tokens(&tm.u, p.u, dtype, p.D, p.L, p.batch, p.ld_u, CH, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.delta, p.delta, dtype, p.D, p.L, p.batch, p.ld_delta, CH, TT))) return rc;
  if (p.z) {
    if ((rc = make_tmap_tokens(&tm.z, p.z, dtype, p.D, p.L, p.batch, p.ld_z, CH, TT))) return rc;
  } else {
    tm.z = tm.u;
  }
  if ((rc = make_tmap_tokens(&tm.B, p.Bm, dtype, kNState, p.L, p.batch, p.ld_B, kNState, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.C, p.Cm, dtype, kNState, p.L, p.batch, p.ld_C, kNState, TT))) return rc;
  if ((rc = make_tmap_tokens(&tm.out, p.out, dtype, p.D, p.L, p.batch, p.ld_out, CH, TT))) return rc;
  const int grid = p.batch * (p.D / CH);
  kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(tm, p);
  return check_launch("selective_scan_fwd");
}

template <typename T>
static int dispatch_scan(const ScanParams& p, int dtype, int variant, cudaStream_t stream) {
  // variant = 100 * (channels per thread) + (states per thread); 0 = heuristic on the independent work available.
  // Measured on B200 (profiles/r01_kernel_bench_scan_variants.jsonl, r01_kernel_bench_scan_ws.jsonl): the
  // warp-specialised kernel (selective_scan_fwd_ws.cu, S = 8, 64-channel CTAs) is the fastest at every batch size
  // and both dtypes (84 vs 93 us at B=32, 532 vs 536-609 us at B=256); the tile-synchronous kernel below remains
  // for channel counts that are not a multiple of 64.
  if (variant == 0) variant = (p.D % 64 == 0) ? 5008 : 108;
  if (variant >= 5000 && variant < 9000) return selective_scan_fwd_ws(p, dtype, variant, stream);
  if (variant == 108) {  // the one ablation kept: the tile-synchronous kernel on 64-channel CTAs (93 us vs 84 us at C1)
    if (p.D % 64 == 0) return launch_scan<T, 8, 1, 64, 16, 3>(p, dtype, stream);
    if (p.D % 16 == 0) return launch_scan<T, 2, 1, 16, 16, 3>(p, dtype, stream);
  }
  set_error("selective_scan_fwd: unsupported D=%d / variant=%d (D must be a multiple of 16)", p.D, variant);
  return SIM_ERR_INVALID;
}

int selective_scan_fwd(const ScanParams& p, int dtype, int variant, cudaStream_t stream) {
  const int es = dtype == 0 ? 4 : 2;
  SIM_REQUIRE(dtype == 0 || dtype == 1, SIM_ERR_INVALID, "selective_scan_fwd: dtype must be 0 (fp32) or 1 (bf16)");
  SIM_REQUIRE(p.batch > 0 && p.L > 0 && p.D > 0, SIM_ERR_INVALID, "selective_scan_fwd: empty problem");
  SIM_REQUIRE(p.u && p.delta && (p.wdt || (p.Bm && p.Cm)) && (p.out || p.out_planes) && p.A, SIM_ERR_INVALID,
              "selective_scan_fwd: null tensor");
  SIM_REQUIRE(!p.z_gate || (p.z && !p.ckpt), SIM_ERR_INVALID,
              "selective_scan_fwd: the pre-gated z flag is an inference feature (needs z, no checkpoints: the backward differentiates silu)");
  SIM_REQUIRE(!p.wdt || (p.D % 64 == 0 && aligned16(p.wdt) && !p.ckpt), SIM_ERR_INVALID,
              "selective_scan_fwd: fused dt_proj needs D %% 64 == 0, 16-byte aligned weight planes and no checkpoints");
  if (p.wdt) variant = 5900;
  SIM_REQUIRE(!p.out_planes || (dtype == 0 && p.D % 64 == 0 && (variant == 0 || variant >= 5000) && aligned16(p.out_planes) &&
                                p.ld_planes % 8 == 0 && p.plane % 8 == 0),
              SIM_ERR_INVALID, "selective_scan_fwd: split-plane output needs fp32 activations, D %% 64 == 0 and 16-byte aligned planes");
  const void* ptrs[] = {p.u, p.delta, p.z, p.Bm, p.Cm, p.out};
  const long lds[] = {p.ld_u, p.ld_delta, p.ld_z, p.ld_B, p.ld_C, p.ld_out};
  for (int i = 0; i < 6; ++i) {
    if (!ptrs[i]) continue;
    SIM_REQUIRE(aligned16(ptrs[i]) && (lds[i] * es) % 16 == 0, SIM_ERR_ALIGN,
                "selective_scan_fwd: tensor %d needs a 16-byte aligned base and row stride (TMA tensor maps)", i);
  }
  return dtype == 0 ? dispatch_scan<float>(p, dtype, variant, stream)
                    : dispatch_scan<__nv_bfloat16>(p, dtype, variant, stream);
}

}  // namespace sim
