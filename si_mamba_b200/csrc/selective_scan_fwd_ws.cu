// Selective scan forward, warp-specialised kernel for sm_100a.
//
// Same contract as selective_scan_fwd.cu (mamba-ssm selective_scan_fn as reached from Mamba.forward,
// models/block.py:72; semantics of selective_scan_ref restated in oracle/mamba.py).  What changes is the schedule.
// ncu on the tile-synchronous kernel showed its phases adding up instead of overlapping (profiles/
// r01_scan_ablation.md): every 16-step tile paid a pre-pass, two block barriers and an epilogue in series with the
// recurrence, with only 2-3 warps per scheduler to hide any of it.  Here a CTA (CH channels of one cloud) is split
// into roles that only meet through mbarriers:
//
//   * elementwise warps (NE threads; thread 0 is also the TMA producer): for tile k+1 they turn the raw TMA stage
//     into fp32 work arrays - (dt, dt*u) interleaved per channel with dt = softplus(delta + bias), B and C widened -
//     and keep silu(z) and D*u for their elements IN REGISTERS; for tile k-1 they read the <h, C> sums the
//     recurrence warps left in shared memory, apply y = (sum + D*u) * silu(z) in place and hand the tile to one
//     TMA tensor store.
//   * recurrence warps (CH * 16/S threads): nothing but the recurrence - one LDS.64 (dt, dt*u) and the B / C
//     broadcast rows per step, FMUL2 / MUFU.EX2 / FFMA2, the transposed butterfly, one STS per finished step.
//     No block barrier, no activation math, no global memory access in steady state.
//
// Optional exp offload (POLY pairs of the S states of a thread): B200 issues 16 MUFU.EX2 /clk/SM, which bounds the
// fp32 scan below the HBM roofline; part of the exps can be evaluated on the FMA pipe instead
// (round-to-nearest range reduction + degree-5 polynomial in packed f32x2, exponent inserted with integer adds).

#include <type_traits>

#include "scan_common.cuh"

namespace sim {

namespace {

constexpr int kDtK = 24;  // dt_rank of the x_proj output row the fused kernel consumes (d_model 384 -> dt_rank 24)

// D (16x8 fp32) += A (16x16 bf16, row) . B (16x8 bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <typename T, int S_, int CH_, int NS_, int POLY_, int EP_ = 0, int NOBF_ = 0, int FUSE_ = 0, int NE_ = 128>
struct ScanWsCfg {
  static constexpr int FUSE = FUSE_;       // 1: delta = W_dt . x_dbl[:, :dt_rank] is computed in the elementwise warps (mma.sync)
  static constexpr int NOBF = NOBF_;       // 1: recurrence lanes store their partial <h, C> sums, the elementwise warps add them
  static constexpr int MINB = (CH_ * (kNState / S_) + NE_) * 3 <= 1152 ? 3 : 1;  // aim at 3 resident CTAs per SM
  static constexpr int EP = EP_;           // 1: the exps of softplus / silu in the elementwise warps run on the FMA pipe too
  static constexpr int S = S_;
  static constexpr int LPC = kNState / S_;
  static constexpr int CH = CH_;
  static constexpr int TT = kScanTile;
  static constexpr int NS = NS_;
  static constexpr int POLY = POLY_;       // packed state pairs per thread whose exp runs on the FMA pipe
  static constexpr int NR = CH_ * LPC;     // recurrence threads
  static constexpr int NE = NE_;           // elementwise threads
  static constexpr int NT = NR + NE;
  static constexpr int CQ = CH_ / 4;       // float4 groups per row
  static constexpr int RPP = NE_ / CQ;     // rows per elementwise pass
  static constexpr int GPT = TT / RPP;     // float4 groups per elementwise thread and tile
  static constexpr int RAW_MAIN = TT * CH_ * (int)sizeof(T);
  static constexpr int RAW_BC = TT * kNState * (int)sizeof(T);
  static constexpr int XW = kDtK + 2 * kNState;                    // columns of the x_proj output row: dt_low | B | C
  static constexpr int RAW_X = TT * XW * (int)sizeof(T);            // fused: one tile of x_dbl rows instead of delta, B, C
  static constexpr int RAW_STAGE = FUSE_ ? 2 * RAW_MAIN + RAW_X : 3 * RAW_MAIN + 2 * RAW_BC;
  static constexpr int NPL = sizeof(T) == 4 ? 3 : 1;                // bf16 planes of an operand (fp32 -> 3, bf16 -> 1)
  static constexpr int XS_LD = 40;                                  // padded row (bf16 elements) of the A-operand planes
  static constexpr int FUSE_SMEM = FUSE_ ? NPL * TT * XS_LD * 2 + TT * CH_ * 4 : 0;  // A planes + fp32 delta tile
  static constexpr int WORK_DT = TT * CH_ * 8;
  static constexpr int WORK_BC = TT * kNState * 4;
  static constexpr int WORK = WORK_DT + 2 * WORK_BC;
  // fp32 <h, C> sums (4 B, or 4 B per lane of a channel without the butterfly) / results in place, or three bf16 planes
  static constexpr int YBUF = NOBF_ ? (TT * CH_ * 4 * (kNState / S_) > TT * CH_ * 6 ? TT * CH_ * 4 * (kNState / S_) : TT * CH_ * 6) : TT * CH_ * 6;
  static constexpr int SMEM = NS_ * RAW_STAGE + 2 * WORK + 2 * YBUF + FUSE_SMEM + (NS_ + 4) * 8 + 16;
  static_assert(RAW_STAGE % 128 == 0 && RAW_MAIN % 128 == 0 && RAW_BC % 128 == 0 && WORK % 128 == 0 && YBUF % 128 == 0 &&
                    RAW_X % 128 == 0 && FUSE_SMEM % 128 == 0,
                "TMA tiles must stay 128-B aligned");
  static_assert(!FUSE_ || (CH_ == 64 && NE_ == 128 && TT == 16), "fused dt_proj: 4 elementwise warps x 2 n-blocks of 8 channels");
  static_assert(NR % 32 == 0 && NE_ % 32 == 0 && NE_ % CQ == 0 && TT % RPP == 0 && GPT >= 1, "role split");
  static_assert(2 * TT * kNState / 4 <= NE_ * 4, "B / C widening loop");
  static_assert(POLY_ <= S_ / 2, "at most S/2 packed pairs");
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_le1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// 2^x for a packed pair on the FMA pipe.  t = x + 1.5*2^23 rounds x to the nearest integer n (kept in the low
// mantissa bits of t), f = x - n lies in [-0.5, 0.5], p(f) is the degree-5 least-maximum polynomial of 2^f
// (max rel. error 7.5e-8, < 2.4e-7 incl. fp32 Horner rounding: tools/exp2_poly.py), and n is added to the exponent field of p.
// x is clamped at -125 (result 2^-125 ~ 2e-38 instead of 0: below fp32 resolution of any product it enters).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  constexpr float kMagic = 12582912.f;
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
  const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  constexpr float c0 = 1.0000001192092896f, c1 = 0.6931469440460205f, c2 = 0.24022120237350464f,
                  c3 = 0.05550713464617729f, c4 = 0.009675540961325169f, c5 = 0.0013276457320898771f;
  float2 p = __ffma2_rn(make_float2(c5, c5), f, make_float2(c4, c4));
  p = __ffma2_rn(p, f, make_float2(c3, c3));
  p = __ffma2_rn(p, f, make_float2(c2, c2));
  p = __ffma2_rn(p, f, make_float2(c1, c1));
  p = __ffma2_rn(p, f, make_float2(c0, c0));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// softplus / silu for four adjacent channels with the exponentials on the FMA pipe (lg2 / rcp stay on the MUFU):
// halves the XU work of the elementwise warps.  Same thresholds as softplus_f / silu_f (common.cuh).
__device__ __forceinline__ float softplus_from_e(float x, float e) {
  const float series = e * (1.f - e * (0.5f - e * (0.33333334f - e * (0.25f - e * 0.2f))));
  const float big = lg2_approx(1.f + e) * kLn2;
  const float r = (e < 0.0625f) ? series : big;
  return (x > 20.f) ? x : r;
}
__device__ __forceinline__ float4 softplus4_poly(float4 x) {
  const float2 a = ex2_poly2(make_float2(fminf(x.x * kLog2e, 126.f), fminf(x.y * kLog2e, 126.f)));
  const float2 b = ex2_poly2(make_float2(fminf(x.z * kLog2e, 126.f), fminf(x.w * kLog2e, 126.f)));
  return make_float4(softplus_from_e(x.x, a.x), softplus_from_e(x.y, a.y), softplus_from_e(x.z, b.x),
                     softplus_from_e(x.w, b.y));
}
__device__ __forceinline__ float4 silu4_poly(float4 z) {
  const float2 a = ex2_poly2(make_float2(fminf(-z.x * kLog2e, 126.f), fminf(-z.y * kLog2e, 126.f)));
  const float2 b = ex2_poly2(make_float2(fminf(-z.z * kLog2e, 126.f), fminf(-z.w * kLog2e, 126.f)));
  return make_float4(z.x * rcp_approx(1.f + a.x), z.y * rcp_approx(1.f + a.y), z.z * rcp_approx(1.f + b.x),
                     z.w * rcp_approx(1.f + b.y));
}

template <typename T>
__device__ __forceinline__ void sts_out4(T* dst, float4 v);
template <>
__device__ __forceinline__ void sts_out4<float>(float* dst, float4 v) {
  *reinterpret_cast<float4*>(dst) = v;
}
template <>
__device__ __forceinline__ void sts_out4<__nv_bfloat16>(__nv_bfloat16* dst, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<const unsigned*>(&lo);
  r.y = *reinterpret_cast<const unsigned*>(&hi);
  *reinterpret_cast<uint2*>(dst) = r;
}

struct PlaneTmaps {
  CUtensorMap p[3];
};

template <typename Cfg, typename T>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) selective_scan_fwd_ws_kernel(const __grid_constant__ ScanTmaps tm,
                                                                        const __grid_constant__ PlaneTmaps ptm,
                                                                        const ScanParams p) {
  constexpr int S = Cfg::S, LPC = Cfg::LPC, CH = Cfg::CH, TT = Cfg::TT, NS = Cfg::NS, NR = Cfg::NR, NE = Cfg::NE,
                CQ = Cfg::CQ, RPP = Cfg::RPP, GPT = Cfg::GPT;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;
  unsigned char* work = smem + NS * Cfg::RAW_STAGE;
  unsigned char* ybuf = work + 2 * Cfg::WORK;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ybuf + 2 * Cfg::YBUF);               // [plane][TT][XS_LD]
  float* sdt = reinterpret_cast<float*>(ybuf + 2 * Cfg::YBUF + Cfg::NPL * TT * Cfg::XS_LD * 2);  // [TT][CH] fp32 delta
  uint64_t* full = reinterpret_cast<uint64_t*>(ybuf + 2 * Cfg::YBUF + Cfg::FUSE_SMEM);
  uint64_t* ready = full + NS;  // [2] work arrays of tile k are complete (NE arrivals)
  uint64_t* done = ready + 2;   // [2] <h, C> sums of tile k are in ybuf (one arrival per recurrence warp)

  const int tid = threadIdx.x;
  const int nchunk = p.D / CH;
  const int b = blockIdx.x / nchunk;
  const int c0 = (blockIdx.x % nchunk) * CH;
  const int ntiles = (p.L + TT - 1) / TT;
  const int nck = (p.L + kScanCkpt - 1) / kScanCkpt;
  const bool has_z = p.z != nullptr;

  if (tid == NR) {
    tma_prefetch_desc(&tm.u);
    tma_prefetch_desc(&tm.delta);  // fused: the map of the x_dbl rows
    if constexpr (!Cfg::FUSE) {
      tma_prefetch_desc(&tm.B);
      tma_prefetch_desc(&tm.C);
    }
    if (p.out_planes) {
      for (int q = 0; q < 3; ++q) tma_prefetch_desc(&ptm.p[q]);
    } else {
      tma_prefetch_desc(&tm.out);
    }
    if (has_z) tma_prefetch_desc(&tm.z);
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ready[s], NE);
      mbar_init(&done[s], NR / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= NR) {
    // ================================================================ elementwise warps (+ TMA producer)
    const int te = tid - NR;
    auto issue_tile = [&](int tile) {
      const int s = tile % NS;
      const int t0 = tile * TT;
      unsigned char* st = raw + s * Cfg::RAW_STAGE;
      if constexpr (Cfg::FUSE) {
        mbar_arrive_expect_tx(&full[s], (has_z ? 2u : 1u) * Cfg::RAW_MAIN + Cfg::RAW_X);
        tma_load_3d(st, &tm.u, c0, t0, b, &full[s]);
        if (has_z) tma_load_3d(st + Cfg::RAW_MAIN, &tm.z, c0, t0, b, &full[s]);
        tma_load_3d(st + 2 * Cfg::RAW_MAIN, &tm.delta, 0, t0, b, &full[s]);  // (dt_low | B | C) rows of x_dbl
      } else {
        mbar_arrive_expect_tx(&full[s], (has_z ? 3u : 2u) * Cfg::RAW_MAIN + 2u * Cfg::RAW_BC);
        tma_load_3d(st, &tm.u, c0, t0, b, &full[s]);
        tma_load_3d(st + Cfg::RAW_MAIN, &tm.delta, c0, t0, b, &full[s]);
        if (has_z) tma_load_3d(st + 2 * Cfg::RAW_MAIN, &tm.z, c0, t0, b, &full[s]);
        tma_load_3d(st + 3 * Cfg::RAW_MAIN, &tm.B, 0, t0, b, &full[s]);
        tma_load_3d(st + 3 * Cfg::RAW_MAIN + Cfg::RAW_BC, &tm.C, 0, t0, b, &full[s]);
      }
    };
    if (te == 0) {
      for (int k = 0; k < NS && k < ntiles; ++k) issue_tile(k);
    }
    const int cc = (te % CQ) * 4;  // my four adjacent channels (same for every tile)
    const int r0 = te / CQ;        // my rows: r0 + i * RPP
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), D4 = bias4;
    if (p.dbias) bias4 = *reinterpret_cast<const float4*>(p.dbias + c0 + cc);
    if (p.Dv) D4 = *reinterpret_cast<const float4*>(p.Dv + c0 + cc);
    float4 gate[2][GPT], du[2][GPT];
    // fused dt_proj: warp `we` owns n-blocks 2 we, 2 we + 1 (8 channels each); its B fragments (W_dt rows of those
    // channels, K padded to 32, NPL bf16 planes) stay in registers for the whole kernel
    const int we = te >> 5, lg = (te & 31) >> 2, lt = te & 3;
    uint32_t bw[2][Cfg::NPL][2][2];
    if constexpr (Cfg::FUSE) {
      const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(p.wdt);
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int pl = 0; pl < Cfg::NPL; ++pl)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const __nv_bfloat16* row = wp + ((long)pl * p.D + c0 + (2 * we + nb) * 8 + lg) * 32 + ks * 16 + 2 * lt;
            bw[nb][pl][ks][0] = *reinterpret_cast<const uint32_t*>(row);
            bw[nb][pl][ks][1] = *reinterpret_cast<const uint32_t*>(row + 8);
          }
    }

    // delta tile of `tile` = dt_low . W_dt^T -> sdt (fused dt_proj).  Runs one tile AHEAD of the pre-pass that consumes it
    // (after `ready` of the current tile has been signalled), so it never delays the recurrence warps.
    auto compute_delta = [&](int tile) {
      if constexpr (Cfg::FUSE) {
          const int s_ = tile % NS;
          mbar_wait(&full[s_], (tile / NS) & 1);
          const T* sx = reinterpret_cast<const T*>(raw + s_ * Cfg::RAW_STAGE + 2 * Cfg::RAW_MAIN);
          constexpr int XW = Cfg::XW, XL = Cfg::XS_LD;
          // (a) A operand: the dt_low columns of the 16 rows as NPL bf16 planes [16][32] (K padded with zeros)
          {
            const int row = te >> 3, k0 = (te & 7) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 < kDtK) v = lds4<T>(sx + row * XW + k0);
            if constexpr (sizeof(T) == 4) {
              split3_store4(xs + row * XL + k0, TT * XL, v);
            } else {
              sts_out4<__nv_bfloat16>(xs + row * XL + k0, v);
            }
          }
          bar_sync(2, NE);
          // (b) delta tile = x_low . W_dt^T on the tensor cores (mma.sync, fp32 accumulate); the six partial products
          // of the 3 x bf16 split, smallest first (same scheme as gemm_split3.cu)
          float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          constexpr int NPROD = Cfg::NPL == 3 ? 6 : 1;
          constexpr int PA[6] = {0, 2, 1, 0, 1, 0}, PB[6] = {2, 0, 1, 1, 0, 0};
#pragma unroll
          for (int q = 0; q < NPROD; ++q) {
            const int pa = Cfg::NPL == 3 ? PA[q] : 0, pb = Cfg::NPL == 3 ? PB[q] : 0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const __nv_bfloat16* ap = xs + pa * TT * XL + ks * 16 + 2 * lt;
              uint32_t a[4];
              a[0] = *reinterpret_cast<const uint32_t*>(ap + lg * XL);
              a[1] = *reinterpret_cast<const uint32_t*>(ap + (lg + 8) * XL);
              a[2] = *reinterpret_cast<const uint32_t*>(ap + lg * XL + 8);
              a[3] = *reinterpret_cast<const uint32_t*>(ap + (lg + 8) * XL + 8);
              mma_bf16_16816(acc[0], a, bw[0][pb][ks]);
              mma_bf16_16816(acc[1], a, bw[1][pb][ks]);
            }
          }
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const int col = (2 * we + nb) * 8 + 2 * lt;
            *reinterpret_cast<float2*>(sdt + lg * CH + col) = make_float2(acc[nb][0], acc[nb][1]);
            *reinterpret_cast<float2*>(sdt + (lg + 8) * CH + col) = make_float2(acc[nb][2], acc[nb][3]);
          }
          bar_sync(2, NE);
      }
    };
    if constexpr (Cfg::FUSE) compute_delta(0);

    auto body = [&](auto PAR, int k) {
      constexpr int par = decltype(PAR)::value;
      if (k < ntiles) {
        const int s = k % NS;
        unsigned char* st = raw + s * Cfg::RAW_STAGE;
        const T* su = reinterpret_cast<const T*>(st);
        const T* sd = reinterpret_cast<const T*>(st + Cfg::RAW_MAIN);
        const T* sz = reinterpret_cast<const T*>(st + (Cfg::FUSE ? 1 : 2) * Cfg::RAW_MAIN);
        const T* sB = reinterpret_cast<const T*>(st + 3 * Cfg::RAW_MAIN);
        const T* sx = reinterpret_cast<const T*>(st + 2 * Cfg::RAW_MAIN);  // fused: [TT][XW] rows of x_dbl
        unsigned char* wk = work + par * Cfg::WORK;
        float* w_dt = reinterpret_cast<float*>(wk);
        float* w_BC = reinterpret_cast<float*>(wk + Cfg::WORK_DT);
        mbar_wait(&full[s], (k / NS) & 1);
        // work arrays `par` were last read by the recurrence of tile k-2, whose `done` this thread observed in
        // the previous iteration
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
          const int r = r0 + i * RPP;
          float4 dv = Cfg::FUSE ? *reinterpret_cast<const float4*>(sdt + r * CH + cc) : lds4<T>(sd + r * CH + cc);
          dv.x += bias4.x, dv.y += bias4.y, dv.z += bias4.z, dv.w += bias4.w;
          if (p.softplus) {
            if constexpr (Cfg::EP) dv = softplus4_poly(dv);
            else dv = make_float4(softplus_f(dv.x), softplus_f(dv.y), softplus_f(dv.z), softplus_f(dv.w));
          }
          const float4 uv = lds4<T>(su + r * CH + cc);
          float4* wd = reinterpret_cast<float4*>(w_dt + (r * CH + cc) * 2);
          wd[0] = make_float4(dv.x, dv.x * uv.x, dv.y, dv.y * uv.y);
          wd[1] = make_float4(dv.z, dv.z * uv.z, dv.w, dv.w * uv.w);
          du[par][i] = make_float4(D4.x * uv.x, D4.y * uv.y, D4.z * uv.z, D4.w * uv.w);
          float4 gv = make_float4(1.f, 1.f, 1.f, 1.f);
          if (has_z) {
            const float4 zv = lds4<T>(sz + r * CH + cc);
            if (p.z_gate) gv = zv;  // the in_proj epilogue already applied silu (gemm_split3.cu, EpiAct mode 1)
            else if constexpr (Cfg::EP) gv = silu4_poly(zv);
            else gv = make_float4(silu_f(zv.x), silu_f(zv.y), silu_f(zv.z), silu_f(zv.w));
          }
          gate[par][i] = gv;
        }
        // B and C are adjacent in the raw stage and in the work arrays: widen both with one loop
        if constexpr (Cfg::FUSE) {
          for (int g = te; g < 2 * TT * kNState / 4; g += NE) {  // B then C, [TT][16] each, from the x_dbl rows
            const int which = g / (TT * kNState / 4), rr = (g % (TT * kNState / 4)) / (kNState / 4), c4 = (g % (kNState / 4)) * 4;
            *reinterpret_cast<float4*>(w_BC + 4 * g) = lds4<T>(sx + rr * Cfg::XW + kDtK + which * kNState + c4);
          }
        } else {
          for (int g = te; g < 2 * TT * kNState / 4; g += NE)
            *reinterpret_cast<float4*>(w_BC + 4 * g) = lds4<T>(sB + 4 * g);
        }
        // ybuf[par] is rewritten by the recurrence of tile k: the TMA store of tile k-2 must have read it
        if (te == 0) bulk_wait_read0();
        mbar_arrive(&ready[par]);
        bar_sync(1, NE);  // every elementwise thread has left raw stage s
        if (te == 0 && k + NS < ntiles) issue_tile(k + NS);
        if constexpr (Cfg::FUSE) {
          if (k + 1 < ntiles) compute_delta(k + 1);
        }
      }
      if (k >= 1) {
        constexpr int q = par ^ 1;  // parity of tile j = k - 1
        const int j = k - 1;
        float* yb = reinterpret_cast<float*>(ybuf + q * Cfg::YBUF);
        mbar_wait(&done[q], (j / 2) & 1);
        float4 o[GPT];
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
          const int r = r0 + i * RPP;
          float4 y;
          if constexpr (Cfg::NOBF) {
            const float* yp = yb + (r * CH + cc) * LPC;  // LPC partial sums per channel, channel-major
            float acc4[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              acc4[c4] = 0.f;
#pragma unroll
              for (int l = 0; l < LPC; ++l) acc4[c4] += yp[c4 * LPC + l];
            }
            y = make_float4(acc4[0], acc4[1], acc4[2], acc4[3]);
          } else {
            y = *reinterpret_cast<const float4*>(yb + r * CH + cc);
          }
          o[i] = make_float4((y.x + du[q][i].x) * gate[q][i].x, (y.y + du[q][i].y) * gate[q][i].y,
                             (y.z + du[q][i].z) * gate[q][i].z, (y.w + du[q][i].w) * gate[q][i].w);
        }
        const bool split = p.out_planes != nullptr;
        if (sizeof(T) != 4 || split || Cfg::NOBF) bar_sync(2, NE);  // the outputs overlap other threads' fp32 sums
        if (split) {
          // out_proj operand: three bf16 planes, each a dense (TT, CH) tile
#pragma unroll
          for (int i = 0; i < GPT; ++i)
            split3_store4(reinterpret_cast<__nv_bfloat16*>(yb) + (r0 + i * RPP) * CH + cc, TT * CH, o[i]);
        } else {
#pragma unroll
          for (int i = 0; i < GPT; ++i) sts_out4<T>(reinterpret_cast<T*>(yb) + (r0 + i * RPP) * CH + cc, o[i]);
        }
        fence_proxy_async();
        bar_sync(1, NE);
        if (te == 0) {
          if (split) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
              tma_store_3d(&ptm.p[q], c0, j * TT, b, reinterpret_cast<__nv_bfloat16*>(yb) + q * TT * CH);
          } else {
            tma_store_3d(&tm.out, c0, j * TT, b, yb);  // rows past L are clipped by the TMA unit
          }
          bulk_commit();
        }
      }
    };
    for (int k = 0; k <= ntiles; k += 2) {
      body(std::integral_constant<int, 0>{}, k);
      if (k + 1 <= ntiles) body(std::integral_constant<int, 1>{}, k + 1);
    }
    if (te == 0) bulk_wait0();
  } else {
    // ================================================================ recurrence warps
    const int sub = tid % LPC;  // which S-state slice
    const int ch = tid / LPC;   // channel within the CTA
    float2 A2[S / 2], h[S / 2];
#pragma unroll
    for (int j = 0; j < S / 2; ++j) {
      const float* Ap = p.A + (long)(c0 + ch) * kNState + sub * S + 2 * j;
      A2[j] = make_float2(Ap[0] * kLog2e, Ap[1] * kLog2e);
      h[j] = make_float2(0.f, 0.f);
    }
    for (int k = 0; k < ntiles; ++k) {
      const int par = k & 1;
      // training forward: state before every kScanCkpt-th step, layout (batch, ceil(L / kScanCkpt), D, 16)
      auto save_ckpt = [&](int idx) {
        float2* dst = reinterpret_cast<float2*>(p.ckpt + (((long)b * nck + idx) * p.D + c0 + ch) * kNState + sub * S);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) dst[j] = h[j];
      };
      if (p.ckpt) save_ckpt(k * (TT / kScanCkpt));
      const unsigned char* wk = work + par * Cfg::WORK;
      const float2* w_dt = reinterpret_cast<const float2*>(wk) + ch;
      const float* w_B = reinterpret_cast<const float*>(wk + Cfg::WORK_DT) + sub * S;
      const float* w_C = w_B + TT * kNState;
      float* yb = reinterpret_cast<float*>(ybuf + par * Cfg::YBUF) + (Cfg::NOBF ? ch * LPC + sub : ch);
      mbar_wait(&ready[par], (k / 2) & 1);
      float part[LPC];
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        if (t > 0 && t % kScanCkpt == 0) {
          if (p.ckpt && k * (TT / kScanCkpt) + t / kScanCkpt < nck) save_ckpt(k * (TT / kScanCkpt) + t / kScanCkpt);
        }
        const float2 d = w_dt[t * CH];  // (dt, dt * u)
        float Bv[S], Cv[S];
        lds_vec<S>(w_B + t * kNState, Bv);
        lds_vec<S>(w_C + t * kNState, Cv);
        const float2 dt2 = make_float2(d.x, d.x), dtu2 = make_float2(d.y, d.y);
        float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
          const float2 x = __fmul2_rn(dt2, A2[j]);
          const float2 a = (j < Cfg::POLY) ? ex2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
          const float2 bu = __fmul2_rn(dtu2, make_float2(Bv[2 * j], Bv[2 * j + 1]));
          h[j] = __ffma2_rn(a, h[j], bu);
          acc[j & 1] = __ffma2_rn(h[j], make_float2(Cv[2 * j], Cv[2 * j + 1]), acc[j & 1]);
        }
        const float2 acc01 = __fadd2_rn(acc[0], acc[1]);
        if constexpr (Cfg::NOBF) {
          yb[t * CH * LPC] = acc01.x + acc01.y;
          continue;
        }
        part[t % LPC] = acc01.x + acc01.y;
        if ((t + 1) % LPC == 0) {
          // transposed butterfly: lane `sub` ends with the full sum of step t + 1 - LPC + sub, and stores it
#pragma unroll
          for (int o = LPC / 2; o >= 1; o >>= 1) {
            const bool up = (sub & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
              const float send = up ? part[i] : part[i + o];
              const float keep = up ? part[i + o] : part[i];
              part[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          yb[(t + 1 - LPC + sub) * CH] = part[0];
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&done[par]);
    }
  }
}

template <typename T, int S, int CH, int NS, int POLY, int EP = 0, int NOBF = 0, int FUSE = 0>
int launch_scan_ws(const ScanParams& p, int dtype, cudaStream_t stream) {
  using Cfg = ScanWsCfg<T, S, CH, NS, POLY, EP, NOBF, FUSE>;
  constexpr int TT = Cfg::TT;
  auto kern = selective_scan_fwd_ws_kernel<Cfg, T>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, Cfg::SMEM, attr) != cudaSuccess) return check_launch("selective_scan_fwd_ws attr");
  ScanTmaps tm;
  int rc;
  if ((rc = make_tmap_tokens(&tm.u, p.u, dtype, p.D, p.L, p.batch, p.ld_u, CH, TT))) return rc;
  if (p.z) {
    if ((rc = make_tmap_tokens(&tm.z, p.z, dtype, p.D, p.L, p.batch, p.ld_z, CH, TT))) return rc;
  } else {
    tm.z = tm.u;
  }
  if constexpr (FUSE) {
    // `delta` is the x_proj output (rows of dt_low | B | C): one box of all its columns per tile
    if ((rc = make_tmap_tokens(&tm.delta, p.delta, dtype, Cfg::XW, p.L, p.batch, p.ld_delta, Cfg::XW, TT))) return rc;
    tm.B = tm.C = tm.delta;
  } else {
    if ((rc = make_tmap_tokens(&tm.delta, p.delta, dtype, p.D, p.L, p.batch, p.ld_delta, CH, TT))) return rc;
    if ((rc = make_tmap_tokens(&tm.B, p.Bm, dtype, kNState, p.L, p.batch, p.ld_B, kNState, TT))) return rc;
    if ((rc = make_tmap_tokens(&tm.C, p.Cm, dtype, kNState, p.L, p.batch, p.ld_C, kNState, TT))) return rc;
  }
  PlaneTmaps ptm;
  if (p.out_planes) {
    if (dtype != 0) {
      set_error("selective_scan_fwd_ws: split bf16 planes are an fp32-activation output format");
      return SIM_ERR_INVALID;
    }
    for (int q = 0; q < 3; ++q)
      if ((rc = make_tmap_tokens(&ptm.p[q], static_cast<const __nv_bfloat16*>(p.out_planes) + q * p.plane, 1, p.D, p.L,
                                 p.batch, p.ld_planes, CH, TT)))
        return rc;
    tm.out = tm.u;
  } else {
    if ((rc = make_tmap_tokens(&tm.out, p.out, dtype, p.D, p.L, p.batch, p.ld_out, CH, TT))) return rc;
    ptm.p[0] = ptm.p[1] = ptm.p[2] = tm.out;
  }
  kern<<<p.batch * (p.D / CH), Cfg::NT, Cfg::SMEM, stream>>>(tm, ptm, p);
  return check_launch("selective_scan_fwd_ws");
}

template <typename T>
int dispatch_ws(const ScanParams& p, int dtype, int variant, cudaStream_t stream) {
  // variant = 5000 + 100 * (polynomial pairs) + states per thread (6000: 32-channel CTAs, 7000: polynomial
  // exps in the elementwise warps too)
  switch (variant) {
    case 5004: return launch_scan_ws<T, 4, 64, 3, 0>(p, dtype, stream);
    case 5008: return launch_scan_ws<T, 8, 64, 3, 0>(p, dtype, stream);
    case 5016: return launch_scan_ws<T, 16, 64, 3, 0>(p, dtype, stream);
    case 5108: return launch_scan_ws<T, 8, 64, 3, 1>(p, dtype, stream);
    case 5208: return launch_scan_ws<T, 8, 64, 3, 2>(p, dtype, stream);
    case 5104: return launch_scan_ws<T, 4, 64, 3, 1>(p, dtype, stream);
    case 5116: return launch_scan_ws<T, 16, 64, 3, 1>(p, dtype, stream);
    case 5216: return launch_scan_ws<T, 16, 64, 3, 2>(p, dtype, stream);
    case 5900: return launch_scan_ws<T, 8, 64, 3, 0, 0, 0, 1>(p, dtype, stream);  // fused dt_proj
    case 5508: return launch_scan_ws<T, 8, 64, 2, 0>(p, dtype, stream);
    case 8008: return launch_scan_ws<T, 8, 64, 2, 0, 0, 1>(p, dtype, stream);
    case 8004: return launch_scan_ws<T, 4, 64, 2, 0, 0, 1>(p, dtype, stream);
    case 7008: return launch_scan_ws<T, 8, 64, 3, 0, 1>(p, dtype, stream);
    case 7108: return launch_scan_ws<T, 8, 64, 3, 1, 1>(p, dtype, stream);
    case 7208: return launch_scan_ws<T, 8, 64, 3, 2, 1>(p, dtype, stream);
    case 6008: return launch_scan_ws<T, 8, 32, 3, 0>(p, dtype, stream);
    case 6108: return launch_scan_ws<T, 8, 32, 3, 1>(p, dtype, stream);
  }
  set_error("selective_scan_fwd_ws: unknown variant %d", variant);
  return SIM_ERR_INVALID;
}

}  // namespace

int selective_scan_fwd_ws(const ScanParams& p, int dtype, int variant, cudaStream_t stream) {
  if (p.D % 64 != 0) {
    set_error("selective_scan_fwd_ws: D=%d must be a multiple of 64", p.D);
    return SIM_ERR_INVALID;
  }
  return dtype == 0 ? dispatch_ws<float>(p, dtype, variant, stream)
                    : dispatch_ws<__nv_bfloat16>(p, dtype, variant, stream);
}

}  // namespace sim
