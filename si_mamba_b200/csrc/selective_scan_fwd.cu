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

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

constexpr int kNState = 16;

struct ScanTmaps {
  CUtensorMap u, delta, z, B, C, out;
};

template <typename T, int S_, int CH_, int TT_, int NS_>
struct ScanCfg {
  static constexpr int S = S_;              // states per thread
  static constexpr int LPC = kNState / S_;  // lanes per channel
  static constexpr int CH = CH_;            // channels per CTA
  static constexpr int TT = TT_;            // time steps per tile
  static constexpr int NS = NS_;            // raw stages in flight
  static constexpr int NT = CH_ * LPC;      // threads per CTA
  static constexpr int RAW_MAIN = TT_ * CH_ * (int)sizeof(T);
  static constexpr int RAW_BC = TT_ * kNState * (int)sizeof(T);
  static constexpr int RAW_STAGE = 3 * RAW_MAIN + 2 * RAW_BC;  // multiple of 128 for every built config
  static constexpr int WORK = 3 * TT_ * CH_ * 4 + 2 * TT_ * kNState * 4;
  static constexpr int OBUF = TT_ * CH_ * (int)sizeof(T);
  static constexpr int SMEM = NS_ * RAW_STAGE + WORK + OBUF + NS_ * 8 + 128;
  static_assert(RAW_STAGE % 128 == 0 && RAW_MAIN % 128 == 0 && RAW_BC % 128 == 0, "TMA tiles must stay 128-B aligned");
};

template <typename Cfg, typename T>
__global__ void __launch_bounds__(Cfg::NT) selective_scan_fwd_kernel(const __grid_constant__ ScanTmaps tm,
                                                                     const ScanParams p) {
  constexpr int S = Cfg::S, LPC = Cfg::LPC, CH = Cfg::CH, TT = Cfg::TT, NS = Cfg::NS, NT = Cfg::NT;
  // NOTE: derive every pointer from the extern array itself (no integer round-trips), otherwise the
  // compiler loses the shared address space and emits generic LD/ST instead of LDS/STS.
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;
  T* obuf = reinterpret_cast<T*>(smem + NS * Cfg::RAW_STAGE);
  float* w_dt = reinterpret_cast<float*>(smem + NS * Cfg::RAW_STAGE + Cfg::OBUF);
  float* w_u = w_dt + TT * CH;
  float* w_g = w_u + TT * CH;
  float* w_B = w_g + TT * CH;
  float* w_C = w_B + TT * kNState;
  uint64_t* full = reinterpret_cast<uint64_t*>(w_C + TT * kNState);

  const int tid = threadIdx.x;
  const int nchunk = p.D / CH;
  const int b = blockIdx.x / nchunk;
  const int c0 = (blockIdx.x % nchunk) * CH;
  const int c = tid / LPC;    // channel within the CTA
  const int sub = tid % LPC;  // which S-state slice of the channel
  const int ntiles = (p.L + TT - 1) / TT;
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
  float2 A2[S / 2];
  float2 h[S / 2];
#pragma unroll
  for (int j = 0; j < S / 2; ++j) {
    const float* Ap = p.A + (long)(c0 + c) * kNState + sub * S + 2 * j;
    A2[j] = make_float2(Ap[0] * kLog2e, Ap[1] * kLog2e);
    h[j] = make_float2(0.f, 0.f);
  }
  const float Dc = p.Dv ? p.Dv[c0 + c] : 0.f;
  // pre-pass mapping: element e = r*CH + cc; NT is a multiple of CH so cc is fixed per thread
  const int cc = tid % CH;
  const float bias_cc = p.dbias ? p.dbias[c0 + cc] : 0.f;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int s = tile % NS;
    const int t0 = tile * TT;
    unsigned char* st = raw + s * Cfg::RAW_STAGE;
    const T* su = reinterpret_cast<const T*>(st);
    const T* sd = reinterpret_cast<const T*>(st + Cfg::RAW_MAIN);
    const T* sz = reinterpret_cast<const T*>(st + 2 * Cfg::RAW_MAIN);
    const T* sB = reinterpret_cast<const T*>(st + 3 * Cfg::RAW_MAIN);
    const T* sC = reinterpret_cast<const T*>(st + 3 * Cfg::RAW_MAIN + Cfg::RAW_BC);

    mbar_wait(&full[s], (tile / NS) & 1);

    // ---- pre-pass: softplus(delta + bias), silu(z), widen to fp32.  Always the whole tile: rows past L were
    // zero-filled by the TMA unit, their results are clipped by the TMA store and nothing after them is used.
#pragma unroll
    for (int r = tid / CH; r < TT; r += NT / CH) {
      const int e = r * CH + cc;
      const float dv = to_f32<T>(sd[e]) + bias_cc;
      w_dt[e] = p.softplus ? softplus_f(dv) : dv;
      w_u[e] = to_f32<T>(su[e]);
      w_g[e] = has_z ? silu_f(to_f32<T>(sz[e])) : 1.f;
    }
    for (int e = tid; e < TT * kNState; e += NT) {
      w_B[e] = to_f32<T>(sB[e]);
      w_C[e] = to_f32<T>(sC[e]);
    }
    // the previous tile's TMA store must have finished reading obuf before it is rewritten
    if (tid == 0) bulk_wait_read0();
    __syncthreads();

    // raw stage s is free again: refill it with tile + NS
    if (tid == 0 && tile + NS < ntiles) issue_tile(tile + NS);

    // ---- the recurrence over this tile, GRP steps at a time: all shared-memory operands of a group are
    // fetched up front (independent LDS in flight), then the dependent FMUL2 / MUFU / FFMA2 chain runs.
    constexpr int GRP = LPC >= 2 ? LPC : 2;
#pragma unroll 1
    for (int r0 = 0; r0 < TT; r0 += GRP) {
      float dtv[GRP], uv[GRP];
      float2 Bv[GRP][S / 2], Cv[GRP][S / 2];
#pragma unroll
      for (int q = 0; q < GRP; ++q) {
        const int r = r0 + q;
        dtv[q] = w_dt[r * CH + c];
        uv[q] = w_u[r * CH + c];
        const float2* Bp = reinterpret_cast<const float2*>(w_B + r * kNState + sub * S);
        const float2* Cp = reinterpret_cast<const float2*>(w_C + r * kNState + sub * S);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
          Bv[q][j] = Bp[j];
          Cv[q][j] = Cp[j];
        }
      }
      float part[GRP];
#pragma unroll
      for (int q = 0; q < GRP; ++q) {
        const float dtu = dtv[q] * uv[q];
        const float2 dt2 = make_float2(dtv[q], dtv[q]);
        const float2 dtu2 = make_float2(dtu, dtu);
        float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
          const float2 x = __fmul2_rn(dt2, A2[j]);
          const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          const float2 bu = __fmul2_rn(dtu2, Bv[q][j]);
          h[j] = __ffma2_rn(a, h[j], bu);
          acc2 = __ffma2_rn(h[j], Cv[q][j], acc2);
        }
        part[q] = acc2.x + acc2.y;
      }
#pragma unroll
      for (int g2 = 0; g2 < GRP; g2 += LPC) {
        // transposed butterfly: lane `sub` ends with the full sum of step r0 + g2 + sub
#pragma unroll
        for (int o = LPC / 2; o >= 1; o >>= 1) {
          const bool up = (sub & o) != 0;
#pragma unroll
          for (int i = 0; i < o; ++i) {
            const float send = up ? part[g2 + i] : part[g2 + i + o];
            const float keep = up ? part[g2 + i + o] : part[g2 + i];
            part[g2 + i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        const int e = (r0 + g2 + sub) * CH + c;
        obuf[e] = from_f32<T>((part[g2] + Dc * w_u[e]) * w_g[e]);
      }
    }

    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm.out, c0, t0, b, obuf);  // rows past L are clipped by the TMA unit
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait0();
}

template <typename T, int S, int CH, int TT, int NS>
static int launch_scan(const ScanParams& p, int dtype, cudaStream_t stream) {
  using Cfg = ScanCfg<T, S, CH, TT, NS>;
  auto kern = selective_scan_fwd_kernel<Cfg, T>;
  static bool attr_done = false;  // idempotent; a benign race only repeats the call
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
      return check_launch("selective_scan_fwd attr");
    attr_done = true;
  }
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
  // variant: states per thread.  0 = heuristic on the amount of independent work.
  const long rows = (long)p.batch * p.D;
  if (variant == 0) variant = rows >= 148L * 4 * 32 * 16 ? 16 : (rows >= 148L * 4 * 32 * 4 ? 8 : 4);
  if (p.D % 64 == 0) {
    switch (variant) {
      case 16: return launch_scan<T, 16, 64, 16, 3>(p, dtype, stream);
      case 8: return launch_scan<T, 8, 64, 16, 3>(p, dtype, stream);
      case 4: return launch_scan<T, 4, 64, 16, 3>(p, dtype, stream);
      case 2: return launch_scan<T, 2, 32, 16, 3>(p, dtype, stream);
      // tuning alternatives (100*k + S): narrower CTAs / longer tiles
      case 104: return launch_scan<T, 4, 32, 16, 3>(p, dtype, stream);
      case 108: return launch_scan<T, 8, 32, 16, 3>(p, dtype, stream);
      case 204: return launch_scan<T, 4, 64, 32, 2>(p, dtype, stream);
      case 208: return launch_scan<T, 8, 64, 32, 2>(p, dtype, stream);
      case 304: return launch_scan<T, 4, 32, 32, 2>(p, dtype, stream);
      case 102: return launch_scan<T, 2, 64, 16, 3>(p, dtype, stream);
    }
  } else if (p.D % 16 == 0) {
    return launch_scan<T, 4, 16, 32, 2>(p, dtype, stream);
  }
  set_error("selective_scan_fwd: unsupported D=%d / variant=%d (D must be a multiple of 16)", p.D, variant);
  return SIM_ERR_INVALID;
}

int selective_scan_fwd(const ScanParams& p, int dtype, int variant, cudaStream_t stream) {
  const int es = dtype == 0 ? 4 : 2;
  SIM_REQUIRE(dtype == 0 || dtype == 1, SIM_ERR_INVALID, "selective_scan_fwd: dtype must be 0 (fp32) or 1 (bf16)");
  SIM_REQUIRE(p.batch > 0 && p.L > 0 && p.D > 0, SIM_ERR_INVALID, "selective_scan_fwd: empty problem");
  SIM_REQUIRE(p.u && p.delta && p.Bm && p.Cm && p.out && p.A, SIM_ERR_INVALID, "selective_scan_fwd: null tensor");
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
