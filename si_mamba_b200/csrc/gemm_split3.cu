// fp32-accurate projection GEMM on the 5th-gen tensor cores from PRE-SPLIT bf16 operands, hand-written for sm_100a
// (TMA tensor loads -> shared memory -> tcgen05.mma with the accumulator in TMEM -> tcgen05.ld epilogue).
//
//   Y[M,N] (fp32) = X[M,K] . W[N,K]^T      (in_proj / x_proj / dt_proj / out_proj of Mamba.forward, models/block.py:72,
//                                            on the reference's no-autocast fp32 finetune / test path)
//
// Every fp32 operand is carried as three bf16 planes x = x0 + x1 + x2 (8 + 8 + 8 significand bits, residuals formed
// exactly in fp32: sim_split3_bf16 below, or the producing kernel's epilogue).  The product keeps the six partial
// products whose weight is >= 2^-16 of the leading one,
//       x0.w2 + x2.w0 + x1.w1 + x0.w1 + x1.w0 + x0.w0       (dropped: x1.w2, x2.w1, x2.w2 <= 2^-24 relative)
// accumulated in fp32 in tensor memory - the same "3 x bf16" emulation as CUTLASS' FastF32 kernels (gemm_fastf32.cu),
// but with the operand split hoisted out of the main loop: the weights are split once per model, the activations by
// the kernel that produces them, so the GEMM main loop is pure TMA + MMA (no transform warps, 6 instead of 9 MMAs).
// The tensor core's fp32 accumulation truncates (measured: one shared accumulator drifts to 5e-6 at K = 768, 5x a
// true fp32 GEMM), so the leading product x0.w0 and the five correction products (2^-8 .. 2^-16 of it) accumulate in
// TWO TMEM accumulators that are only added in the epilogue: the truncation of the small one is 2^-8 smaller, the
// large one sees 6x fewer accumulations.
//
// Roofline: tensor pipe, 6 x (2 M N K) bf16 flops.  One CTA per 128 x BN output tile, 6 warps: warp 0 = TMA producer
// (one thread), warp 1 = TMEM allocator + MMA issuer (one thread), warps 2-5 = epilogue (TMEM -> registers -> padded
// shared-memory staging -> coalesced 16-byte global stores).

#include <stdlib.h>

#include <algorithm>

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

namespace {

constexpr int kBM = 128;  // rows of X per CTA == TMEM lanes
constexpr int kBK = 32;   // K elements per pipeline stage = one 64-byte swizzle row of bf16
constexpr int kUK = 16;   // K of one tcgen05.mma.kind::f16

// NP = operand planes: 3 = bf16 planes, six products (any fp32 operand); 2 = fp16 planes hi + 2^11 lo, three products
// (operands whose magnitude is bounded below the fp16 range, see the header comment of gemm_planes below)
template <int BN, int NSTAGE_, int NP = 3>
struct GemmCfg {
  static constexpr int A_TILE = kBM * kBK * 2;
  static constexpr int B_TILE = BN * kBK * 2;
  static constexpr int STAGE = NP * A_TILE + NP * B_TILE;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int MINB = (NSTAGE_ * STAGE + 2304) * 2 <= 232448 && 4 * BN <= 512 && BN != 192 ? 2 : 1;  // CTAs per SM
  static constexpr int STG_LD = 36;  // padded row stride (floats) of the epilogue staging tile: 16-B aligned, conflict-free
  static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
  static_assert(4 * 32 * STG_LD * 4 <= STAGE, "epilogue staging reuses pipeline stage 0");
  static_assert(A_TILE % 1024 == 0 && B_TILE % 512 == 0, "swizzle-64B tiles need 512-byte aligned bases");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- PTX wrappers (tcgen05 / TMEM)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]^T, both K-major, bf16 in, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- CTA pairs (thread-block cluster of two): TMA multicast of the shared W tile, cross-CTA stage release
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one tensor copy whose box lands at the SAME shared-memory offset, and signals the mbarrier at the same offset, in every
// CTA of cta_mask
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, "
      "%4}], [%5], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// arrive on the mbarrier at this offset in every CTA of cta_mask once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- cta_group::2: one MMA spans the two SMs of a pair (M = 256; each CTA supplies its 128 rows of A, HALF of the W tile
// and holds its 128 accumulator lanes).  PTX forms as in CUTLASS' sm100 2-SM atoms (cute/arch/{mma_sm100_umma,
// copy_sm100_tma,tmem_allocator_sm100}.hpp, cutlass/arch/barrier.h).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: -> rank 0's copy
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// tensor copy into MY shared memory whose completion bytes are counted on rank 0's mbarrier (the MMA issuer's)
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], "
      "[%5];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar) & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor of a K-major bf16 tile whose rows are one 64-byte swizzle span (kBK = 32):
// start address >> 4 in bits [0,14), stride between 8-row groups (8 x 64 B) >> 4 in bits [32,46), descriptor version 1
// in bits [46,48), layout SWIZZLE_64B (= 4) in bits [61,64).  (Field layout: cute/arch/mma_sm100_desc.hpp.)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// Instruction descriptor: D fp32 (bits [4,6) = 1), A and B bf16 (bits [7,10) and [10,13) = 1), both K-major,
// N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A / B element format field of the instruction descriptor: 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t umma_idesc_planes(int M, int N, int np) {
  return np == 3 ? umma_idesc_bf16(M, N) : ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24));
}
// (X plane, W plane) pairs, smallest contributions first; the LAST pair is the leading product (accumulator 0)
template <int NP>
struct PlaneProducts;
template <>
struct PlaneProducts<3> {
  static constexpr int N = 6;
  __host__ __device__ static constexpr int pa(int q) { return q == 0 ? 0 : q == 1 ? 2 : q == 2 ? 1 : q == 3 ? 0 : q == 4 ? 1 : 0; }
  __host__ __device__ static constexpr int pb(int q) { return q == 0 ? 2 : q == 1 ? 0 : q == 2 ? 1 : q == 3 ? 1 : q == 4 ? 0 : 0; }
  static constexpr float LOW = 1.f;  // weight of the correction accumulator in the epilogue
};
template <>
struct PlaneProducts<2> {  // x = x0 + 2^-11 x1', w = w0 + 2^-11 w1': x0.w1' + x1'.w0 carry 2^11, x1'.w1' (2^-22) is dropped
  static constexpr int N = 3;
  __host__ __device__ static constexpr int pa(int q) { return q == 1 ? 1 : 0; }
  __host__ __device__ static constexpr int pb(int q) { return q == 0 ? 1 : 0; }
  static constexpr float LOW = 1.f / 2048.f;
};

// Optional activation fused into the epilogue (inference): the producer of an activation applies what its only
// consumer, the scan, would otherwise evaluate on the XU pipe that bounds it:
//   mode 1: columns >= col0 -> silu(v)           (in_proj: the z half leaves as the gate silu(z))
//   mode 2: every column   -> softplus(v + bias)  (dt_proj: delta leaves as dt = softplus(delta + dt_proj.bias))
// Same device functions as the scan's own pre-pass (common.cuh), so both paths agree bit for bit.
struct EpiAct {
  int mode;
  int col0;
  const float* bias;
};
__device__ __forceinline__ float epi_act(float v, int col, int N, const EpiAct& a) {
  if (a.mode == 1) return col >= a.col0 ? silu_f(v) : v;
  if (a.mode == 2) return softplus_f(v + (col < N ? a.bias[col] : 0.f));
  return v;
}

struct GemmTmaps {
  CUtensorMap x, w;
};

template <int BN, int NSTAGE_, int NP>
__global__ void __launch_bounds__(192, GemmCfg<BN, NSTAGE_, NP>::MINB) gemm_split3_kernel(const __grid_constant__ GemmTmaps tm, float* __restrict__ Y,
                                                             long ldd, int M, int N, int K, int n_tiles,
                                                             __nv_bfloat16* __restrict__ po, int po_cols, long po_ld,
                                                             long po_plane, int kb_per_split, const EpiAct act) {
  using Cfg = GemmCfg<BN, NSTAGE_, NP>;
  using PP = PlaneProducts<NP>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  extern __shared__ unsigned char smem_raw[];
  // swizzled tiles: align the carve-up to 1024 bytes of the shared ADDRESS space
  const uint32_t base_u32 = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (base_u32 & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* accum_full = empty + NSTAGE;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x / n_tiles) * kBM;
  const int n0 = (blockIdx.x % n_tiles) * BN;
  // split-K (gridDim.y > 1; long contractions with few output tiles, e.g. a weight gradient): CTA y contracts k-blocks
  // [kb0, kb0 + nk) and ADDS its partial tile to the zero-initialised output with 16-byte vector atomics.  Besides
  // filling the SMs this bounds the length of one tensor-core accumulation chain (its fp32 adds truncate).
  const int nk_all = (K + kBK - 1) / kBK;
  const int kb0 = blockIdx.y * kb_per_split;
  const int nk = gridDim.y > 1 ? min(kb_per_split, nk_all - kb0) : nk_all;
  const bool accumulate_out = gridDim.y > 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.x);
    tma_prefetch_desc(&tm.w);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_mbar_init();
  }
  constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);  // power of two >= 2 BN
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);  // columns [0, BN): x0.w0, [BN, 2 BN): the five corrections
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer: one tensor copy per operand and stage brings all three planes of the tile
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        if (kb >= NSTAGE) mbar_wait(&empty[s], ((kb / NSTAGE) - 1) & 1);
        unsigned char* st = smem + s * Cfg::STAGE;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
        tma_load_3d(st, &tm.x, (kb0 + kb) * kBK, m0, 0, &full[s]);
        tma_load_3d(st + NP * Cfg::A_TILE, &tm.w, (kb0 + kb) * kBK, n0, 0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_planes(kBM, BN, NP);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        mbar_wait(&full[s], (kb / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * Cfg::STAGE);
        const uint32_t b0 = a0 + NP * Cfg::A_TILE;
#pragma unroll
        for (int q = 0; q < PP::N; ++q) {
#pragma unroll
          for (int k = 0; k < kBK / kUK; ++k) {
            const uint64_t da = umma_desc_sw64(a0 + PP::pa(q) * Cfg::A_TILE + k * kUK * 2);
            const uint64_t db = umma_desc_sw64(b0 + PP::pb(q) * Cfg::B_TILE + k * kUK * 2);
            umma_bf16(tmem_d + (q == PP::N - 1 ? 0 : BN), da, db, idesc, q == PP::N - 1 ? (kb | k) != 0 : (kb | q | k) != 0);
          }
        }
        umma_commit(&empty[s]);  // stage s may be refilled once these MMAs have read it
      }
      umma_commit(accum_full);
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = rows m0 + 32 (w % 4) + lane
    const int quad = warp & 3;
    mbar_wait(accum_full, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + quad * 32 * Cfg::STG_LD;  // stage 0 is idle now
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= N) break;
      float v[32], sm[32];
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + c * 32, v);
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + BN + c * 32, sm);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = NP == 3 ? v[i] + sm[i] : fmaf(sm[i], 1.f / 2048.f, v[i]);
      if (act.mode) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = epi_act(v[i], n0 + c * 32 + i, N, act);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stg + lane * Cfg::STG_LD + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int row = rr * 4 + (lane >> 3), col = (lane & 7) * 4;
        const int gm = m0 + quad * 32 + row, gn = n0 + c * 32 + col;
        if (gm < M && gn < N) {
          const float4 o = *reinterpret_cast<const float4*>(stg + row * Cfg::STG_LD + col);
          if (accumulate_out) atomicAdd(reinterpret_cast<float4*>(Y + (long)gm * ldd + gn), o);
          else *reinterpret_cast<float4*>(Y + (long)gm * ldd + gn) = o;
          // optionally also emit the first po_cols output columns as split bf16 planes: the operand of a GEMM that
          // consumes them (x_proj -> dt_proj), saving a separate split pass
          if (po && gn < po_cols) split3_store4(po + (long)gm * po_ld + gn, po_plane, o);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (large problems).  ncu on the one-tile-per-CTA kernel above: tensor pipe busy 65 % of the active
// time and SMs active 82 % of the kernel - every tile pays barrier init, TMEM allocation, the first TMA round trip and
// a TMEM -> registers -> staging -> global epilogue that nothing overlaps, because a 200 KB CTA has no co-resident CTA.
// Here one CTA per SM walks tiles j, j + gridDim.x, ...: the smem pipeline keeps rolling across tiles, and EIGHT
// epilogue warps (two per TMEM lane quadrant, half of the tile's columns each) first drain both accumulators into
// registers (BN / 2 fp32 values per thread), release the accumulators to the MMA warp, and only then do the slow part
// (staging + global stores) while the next tile's MMAs are already running.
template <int BN, int NSTAGE_, int NP = 3, int MC = 0>
struct GemmPCfg {
  static constexpr int A_TILE = kBM * kBK * 2;
  static constexpr int B_TILE = (MC == 2 ? BN / 2 : BN) * kBK * 2;  // MC == 2: this CTA holds half of the pair's W tile
  static constexpr int STAGE = NP * A_TILE + NP * B_TILE;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int NEPI = 8;                      // epilogue warps
  static constexpr int NT = 64 + 32 * NEPI;
  static constexpr int STG_LD = 36;
  static constexpr int STG = NEPI * 32 * STG_LD * 4;  // one 32 x 32 (padded) staging tile per epilogue warp
  static constexpr int SMEM = NSTAGE * STAGE + STG + 1024 + 256;
  static constexpr int CW = BN / 2;                   // columns per epilogue warp
  // accumulator sets in TMEM (each = leading + correction accumulator, 2 BN columns): two when they fit (BN <= 128), so the
  // MMAs of tile i + 1 can start while tile i is still being drained.  Measured with 128-wide tiles (profiles/r02_gemm_ncu.md):
  // slower than 192-wide tiles with one set, by the extra operand bytes per flop - the kernel is ingest-bound, not
  // turn-around-bound - so no 128-wide persistent instance is dispatched and NACC is 1 in every built kernel.
  static constexpr int NACC = 4 * BN <= 512 ? 2 : 1;
  static_assert(CW % 32 == 0 && 2 * BN <= 512, "two column halves of whole 32-column chunks; both accumulators in TMEM");
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(A_TILE % 512 == 0 && B_TILE % 512 == 0 && STAGE % 512 == 0, "swizzle-64B tiles need 512-byte aligned bases");
};

// MC = 1: the CTAs run as PAIRS (cluster of two, launch attribute).  A pair owns two vertically adjacent 128-row tiles of
// the same BN columns, so both need the same W tile: each CTA fetches half of its rows and the TMA unit multicasts that
// half into both CTAs' shared memory - one L2 read of W per pair instead of two.  (ncu, profiles/r02_gemm_ncu.md: the kernel
// is bound by operand delivery from L2 at ~9 TB/s, not by the tensor pipe, in both operand formats.)  A stage may be
// refilled once BOTH CTAs' MMAs have read it (the peer's half lands in my buffer), so the MMA warp's commit arrives on
// the `empty` barrier of both CTAs (count 2); the CTAs walk the same number of tiles and k-blocks.
// MC = 2: the pair additionally shares ONE tensor-core instruction stream: rank 0 issues tcgen05.mma.cta_group::2 with
// M = 256, each CTA loads only ITS half of the W rows (no multicast: the MMA reads both halves where they lie), so the
// bytes delivered into each SM per k-block drop from A + W to A + W / 2.  (Measured: multicast alone, MC = 1, changes
// nothing - the bound is bytes delivered to the SMs, not L2 reads.)  Barriers: both CTAs' tensor copies count on rank 0's
// `full`; rank 0's commits arrive on both CTAs' `empty` / `acc_full`; both CTAs' epilogue warps arrive on rank 0's `acc_empty`.
template <int BN, int NSTAGE_, int NP, int MC>
__global__ void __launch_bounds__(GemmPCfg<BN, NSTAGE_, NP, MC>::NT, 1)
    gemm_split3_persistent_kernel(const __grid_constant__ GemmTmaps tm, float* __restrict__ Y, long ldd, int M, int N, int K,
                                  int n_tiles, int total_tiles, const EpiAct act) {
  using Cfg = GemmPCfg<BN, NSTAGE_, NP, MC>;
  static_assert(!MC || (BN / 2) % 8 == 0, "half W tiles must be whole 8-row swizzle groups");
  const int crank = MC ? (int)cluster_ctarank() : 0;
  const int tile_first = MC ? blockIdx.x >> 1 : blockIdx.x;  // MC: `tile` counts pair tiles (two 128-row tiles)
  const int tile_step = MC ? gridDim.x >> 1 : gridDim.x;
  using PP = PlaneProducts<NP>;
  constexpr int NSTAGE = Cfg::NSTAGE, CW = Cfg::CW;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base_u32 = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (base_u32 & 1023u)) & 1023u);
  float* stg_base = reinterpret_cast<float*>(smem + NSTAGE * Cfg::STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE + Cfg::STG);
  uint64_t* empty = full + NSTAGE;
  constexpr int NACC = Cfg::NACC;
  uint64_t* acc_full = empty + NSTAGE;
  uint64_t* acc_empty = acc_full + NACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (K + kBK - 1) / kBK;
  constexpr uint32_t kTmemCols = 2 * BN * NACC <= 128 ? 128 : (2 * BN * NACC <= 256 ? 256 : 512);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.x);
    tma_prefetch_desc(&tm.w);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MC == 1 ? 2 : 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], MC == 2 ? 2 * Cfg::NEPI : Cfg::NEPI);
    }
    fence_mbar_init();
  }
  if constexpr (MC == 2) {
    cluster_sync_all();  // both CTAs are resident before the paired allocation
    if (warp == 1) tmem_alloc2(tmem_ptr, kTmemCols);
  } else {
    if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  auto tile_m0 = [&](int tile) { return (MC ? 2 * (tile / n_tiles) + crank : tile / n_tiles) * kBM; };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int m0 = tile_m0(tile), n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NSTAGE;
          if (it >= NSTAGE) mbar_wait(&empty[s], ((it / NSTAGE) - 1) & 1);
          unsigned char* st = smem + s * Cfg::STAGE;
          if constexpr (MC == 2) {
            if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE);  // both CTAs' copies count on rank 0's barrier
            tma_load_3d_pair(st, &tm.x, kb * kBK, m0, 0, &full[s]);
            tma_load_3d_pair(st + NP * Cfg::A_TILE, &tm.w, kb * kBK, n0 + crank * (BN / 2), 0, &full[s]);
            continue;
          }
          mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
          tma_load_3d(st, &tm.x, kb * kBK, m0, 0, &full[s]);
          if constexpr (MC == 1) {
            // my half of the W rows, plane by plane (the box is (kBK, BN / 2, 1)), into both CTAs
#pragma unroll
            for (int q = 0; q < NP; ++q)
              tma_load_3d_mc(st + NP * Cfg::A_TILE + q * Cfg::B_TILE + crank * (BN / 2) * kBK * 2, &tm.w, kb * kBK,
                             n0 + crank * (BN / 2), q, &full[s], (uint16_t)3);
          } else {
            tma_load_3d(st + NP * Cfg::A_TILE, &tm.w, kb * kBK, n0, 0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (MC != 2 || crank == 0)) {
      constexpr uint32_t idesc = umma_idesc_planes(MC == 2 ? 2 * kBM : kBM, BN, NP);
      int it = 0, i = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++i) {
        const int ab = i % NACC;
        const uint32_t tmem_d = tmem_base + ab * 2 * BN;
        if (i >= NACC) {  // the epilogue warps have drained the tile that used this accumulator set last
          mbar_wait(&acc_empty[ab], ((i / NACC) - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(&full[s], (it / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * Cfg::STAGE);
          const uint32_t b0 = a0 + NP * Cfg::A_TILE;
#pragma unroll
          for (int q = 0; q < PP::N; ++q) {
#pragma unroll
            for (int k = 0; k < kBK / kUK; ++k) {
              const uint64_t da = umma_desc_sw64(a0 + PP::pa(q) * Cfg::A_TILE + k * kUK * 2);
              const uint64_t db = umma_desc_sw64(b0 + PP::pb(q) * Cfg::B_TILE + k * kUK * 2);
              if constexpr (MC == 2)
                umma2_f16(tmem_d + (q == PP::N - 1 ? 0 : BN), da, db, idesc, q == PP::N - 1 ? (kb | k) != 0 : (kb | q | k) != 0);
              else
                umma_bf16(tmem_d + (q == PP::N - 1 ? 0 : BN), da, db, idesc, q == PP::N - 1 ? (kb | k) != 0 : (kb | q | k) != 0);
            }
          }
          if constexpr (MC == 2) umma2_commit_mc(&empty[s], (uint16_t)3);
          else if constexpr (MC == 1) umma_commit_mc(&empty[s], (uint16_t)3);
          else umma_commit(&empty[s]);
        }
        if constexpr (MC == 2) umma2_commit_mc(&acc_full[ab], (uint16_t)3);
        else umma_commit(&acc_full[ab]);
      }
    }
  } else {
    // epilogue warp ew: TMEM lanes [32 (warp % 4), +32), columns [half * CW, +CW)
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    float* stg = stg_base + ew * 32 * Cfg::STG_LD;
    int i = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++i) {
      const int m0 = tile_m0(tile), n0 = (tile % n_tiles) * BN;
      const int ab = i % NACC;
      const uint32_t tmem_d = tmem_base + ab * 2 * BN;
      mbar_wait(&acc_full[ab], (i / NACC) & 1);
      tc_fence_after();
      float v[CW / 32][32];
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) {
        float sm[32];
        const uint32_t col = half * CW + c * 32;
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + col, v[c]);
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + BN + col, sm);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[c][e] = NP == 3 ? v[c][e] + sm[e] : fmaf(sm[e], 1.f / 2048.f, v[c][e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the MMA warp may overwrite this set: the rest overlaps its work
        if constexpr (MC == 2) mbar_arrive_rank0(&acc_empty[ab]);
        else mbar_arrive(&acc_empty[ab]);
      }
#pragma unroll
      for (int c = 0; c < CW / 32; ++c) {
        const int nc = n0 + half * CW + c * 32;
        if (nc >= N) break;
        if (act.mode) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[c][e] = epi_act(v[c][e], nc + e, N, act);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          *reinterpret_cast<float4*>(stg + lane * Cfg::STG_LD + 4 * e) =
              make_float4(v[c][4 * e], v[c][4 * e + 1], v[c][4 * e + 2], v[c][4 * e + 3]);
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int row = rr * 4 + (lane >> 3), col = (lane & 7) * 4;
          const int gm = m0 + quad * 32 + row, gn = nc + col;
          if (gm < M && gn < N)
            *reinterpret_cast<float4*>(Y + (long)gm * ldd + gn) = *reinterpret_cast<const float4*>(stg + row * Cfg::STG_LD + col);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers / read its tiles
  if (warp == 1) {
    tc_fence_after();
    if constexpr (MC == 2) tmem_dealloc2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32-A variant for narrow outputs (x_proj: N <= 64).  The activation is read as fp32 and split into its three bf16
// planes INSIDE the kernel by four transform warps (one thread per row of the 128-row tile), so its producer (the conv)
// does not have to write 6 bytes of planes per element next to the fp32 result and this kernel reads 4 instead of 6.
// Per 32-deep k-block: TMA brings the fp32 tile (128-byte rows, SWIZZLE_128B so the row-per-thread reads are
// conflict-free) and the weight planes; the transform warps write the three A-operand tiles in the K-major SWIZZLE_64B
// layout the MMA descriptors expect (16-byte chunk c of row r lands at chunk c ^ ((r >> 1) & 3)); one thread issues
// the same twelve MMAs as gemm_split3_kernel, so the results are bit-identical to the pre-split path.
constexpr int kCW = 4;  // taps of the fused causal conv (d_conv of the reference's Mamba blocks)

// CONV = true fuses the causal depthwise conv1d + SiLU that precedes x_proj (Mamba.forward: x -> conv1d -> silu -> x_proj):
// the fp32 tile then holds the PRE-conv activation with a 3-row halo (tiles are per cloud, rows before the sequence are
// zero-filled by the TMA unit), the transform threads convolve their (row, 8 channels), write u = silu(conv(x)) to global
// memory for the scan, and split it into the operand planes - the conv kernel and one 50 MB read of u disappear.
template <int BN, bool CONV>
struct GemmF32ACfg {
  static constexpr int RAW_ROWS = kBM + (CONV ? kCW - 1 : 0);
  static constexpr int RAW = (RAW_ROWS * kBK * 4 + 1023) / 1024 * 1024;  // fp32 tile, 128-byte rows
  static constexpr int A_TILE = kBM * kBK * 2;
  static constexpr int B_TILE = BN * kBK * 2;
  static constexpr int STAGE = RAW + 3 * A_TILE + 3 * B_TILE;
  static constexpr int NSTAGE = CONV ? 3 : 4;  // 53 KB stages + 15 KB of conv weights: three fit in 227 KB
  static constexpr int STG_LD = 36;
  static constexpr int CONVW = CONV ? 1 : 0;  // conv weights + bias of all K channels staged once per CTA: K * 5 floats
  static_assert(STAGE % 1024 == 0, "stage bases stay 1024-byte aligned (SWIZZLE_128B)");
};

struct GemmF32ATmaps {
  CUtensorMap x, w;
};

template <int BN, bool CONV>
__global__ void __launch_bounds__(576, 1) gemm_f32a_split3_kernel(const __grid_constant__ GemmF32ATmaps tm,
                                                                  float* __restrict__ Y, long ldd, int M, int N, int K,
                                                                  __nv_bfloat16* __restrict__ po, int po_cols, long po_ld,
                                                                  long po_plane, const float* __restrict__ cw,
                                                                  const float* __restrict__ cb, float* __restrict__ U,
                                                                  long ldu, int L, int tiles_per_cloud) {
  using Cfg = GemmF32ACfg<BN, CONV>;
  constexpr int NSTAGE = Cfg::NSTAGE, HALO = CONV ? kCW - 1 : 0;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base_u32 = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (base_u32 & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE);  // TMA landed (raw + weight planes)
  uint64_t* tfull = full + NSTAGE;                                           // A planes written (512 arrivals)
  uint64_t* empty = tfull + NSTAGE;                                          // MMAs of the stage done
  uint64_t* accum_full = empty + NSTAGE;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_full + 1);
  float* s_cw = reinterpret_cast<float*>(smem + NSTAGE * Cfg::STAGE + 256);  // CONV: [K][4] taps then [K] bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CONV: tiles are per cloud (the halo must not reach into the previous cloud); otherwise rows are one flat M
  const int cloud = CONV ? blockIdx.x / tiles_per_cloud : 0;
  const int t0 = CONV ? (blockIdx.x % tiles_per_cloud) * kBM : 0;
  const int m0 = CONV ? cloud * L + t0 : blockIdx.x * kBM;   // first output row of this tile
  const int m_end = CONV ? cloud * L + L : M;                // rows of this tile are valid below m_end
  const int nk = (K + kBK - 1) / kBK;
  constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.x);
    tma_prefetch_desc(&tm.w);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&tfull[s], 512);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  if constexpr (CONV) {
    for (int i = threadIdx.x; i < K * kCW; i += blockDim.x) s_cw[i] = cw[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_cw[K * kCW + i] = cb ? cb[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        if (kb >= NSTAGE) mbar_wait(&empty[s], ((kb / NSTAGE) - 1) & 1);
        unsigned char* st = smem + s * Cfg::STAGE;
        mbar_arrive_expect_tx(&full[s], Cfg::RAW_ROWS * kBK * 4 + 3 * Cfg::B_TILE);
        if constexpr (CONV) tma_load_3d(st, &tm.x, kb * kBK, t0 - HALO, cloud, &full[s]);  // rows < 0 / >= L: zero-filled
        else tma_load_3d(st, &tm.x, kb * kBK, m0, 0, &full[s]);
        tma_load_3d(st + Cfg::RAW + 3 * Cfg::A_TILE, &tm.w, kb * kBK, 0, 0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      constexpr int PA[6] = {0, 2, 1, 0, 1, 0};
      constexpr int PB[6] = {2, 0, 1, 1, 0, 0};
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        mbar_wait(&tfull[s], (kb / NSTAGE) & 1);  // implies full[s]: the transform warps waited on it
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * Cfg::STAGE + Cfg::RAW);
        const uint32_t b0 = a0 + 3 * Cfg::A_TILE;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
#pragma unroll
          for (int k = 0; k < kBK / kUK; ++k) {
            const uint64_t da = umma_desc_sw64(a0 + PA[q] * Cfg::A_TILE + k * kUK * 2);
            const uint64_t db = umma_desc_sw64(b0 + PB[q] * Cfg::B_TILE + k * kUK * 2);
            umma_bf16(tmem_d + (q == 5 ? 0 : BN), da, db, idesc, q == 5 ? (kb | k) != 0 : (kb | q | k) != 0);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum_full);
    }
  } else {
    // ===== sixteen transform warps (then epilogue): four threads per row of the tile, 8 of the 32 k each
    const int r = (threadIdx.x - 64) & 127, kh = (threadIdx.x - 64) >> 7;
    const int quad = warp & 3;  // TMEM lane quadrant of this warp for the epilogue
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % NSTAGE;
      unsigned char* st = smem + s * Cfg::STAGE;
      mbar_wait(&full[s], (kb / NSTAGE) & 1);
      __nv_bfloat16* arow = reinterpret_cast<__nv_bfloat16*>(st + Cfg::RAW) + r * kBK;
      {  // eight consecutive k = two fp32 chunks in, one 16-byte chunk per plane out
        const int c8 = kh;
        float f[8];
        if constexpr (CONV) {
          // raw row i <-> token t0 - 3 + i: output row r reads raw rows r .. r + 3 (SWIZZLE_128B: chunk ^ (row % 8))
          const int ch = kb * kBK + c8 * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = s_cw[K * kCW + (ch + j < K ? ch + j : 0)];
#pragma unroll
          for (int tap = 0; tap < kCW; ++tap) {
            const int i = r + tap;
            const float4* rawrow = reinterpret_cast<const float4*>(st + i * 128);
            const float4 v0 = rawrow[(2 * c8) ^ (i & 7)], v1 = rawrow[(2 * c8 + 1) ^ (i & 7)];
            const float xv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(s_cw[(ch + j < K ? ch + j : 0) * kCW + tap], xv[j], f[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
          if (m0 + r < m_end && ch < K) {  // u for the scan (token-major, row stride ldu)
            float* up = U + (long)(m0 + r) * ldu + ch;
            *reinterpret_cast<float4*>(up) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(up + 4) = make_float4(f[4], f[5], f[6], f[7]);
          }
        } else {
          const float4* rawrow = reinterpret_cast<const float4*>(st + r * 128);
          const float4 v0 = rawrow[(2 * c8) ^ (r & 7)], v1 = rawrow[(2 * c8 + 1) ^ (r & 7)];  // SWIZZLE_128B: chunk ^ (row % 8)
          f[0] = v0.x, f[1] = v0.y, f[2] = v0.z, f[3] = v0.w, f[4] = v1.x, f[5] = v1.y, f[6] = v1.z, f[7] = v1.w;
        }
        __nv_bfloat16 p[3][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          p[0][j] = __float2bfloat16_rn(f[j]);
          const float r1 = f[j] - __bfloat162float(p[0][j]);
          p[1][j] = __float2bfloat16_rn(r1);
          p[2][j] = __float2bfloat16_rn(r1 - __bfloat162float(p[1][j]));
        }
        const int pc = c8 ^ ((r >> 1) & 3);  // SWIZZLE_64B position of chunk c8 in row r
#pragma unroll
        for (int q = 0; q < 3; ++q)
          *reinterpret_cast<uint4*>(arow + q * (Cfg::A_TILE / 2) + pc * 8) = *reinterpret_cast<const uint4*>(p[q]);
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(&tfull[s]);
    }
    // ===== epilogue (as gemm_split3_kernel)
    mbar_wait(accum_full, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + (quad + 4 * kh) * 32 * Cfg::STG_LD;  // stages 0.. are idle now
#pragma unroll 1
    for (int c = kh; c < BN / 32; c += 4) {  // the four warps of a lane quadrant take alternate 32-column chunks
      if (c * 32 >= N) break;
      float v[32], sm[32];
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + c * 32, v);
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + BN + c * 32, sm);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += sm[i];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stg + lane * Cfg::STG_LD + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int row = rr * 4 + (lane >> 3), col = (lane & 7) * 4;
        const int gm = m0 + quad * 32 + row, gn = c * 32 + col;
        if (gm < m_end && gn < N) {
          const float4 o = *reinterpret_cast<const float4*>(stg + row * Cfg::STG_LD + col);
          *reinterpret_cast<float4*>(Y + (long)gm * ldd + gn) = o;
          if (po && gn < po_cols) split3_store4(po + (long)gm * po_ld + gn, po_plane, o);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, kTmemCols);
  }
}

// 3-D map over the three bf16 planes of a row-major (rows, K) operand: dims (K, rows, 3), 64-byte swizzle.
int make_tmap_planes(CUtensorMap* m, const void* base, int K, int rows, long ld, long plane, int box_rows, int np = 3,
                     int box_planes = 0) {
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SIM_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)np};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, (cuuint32_t)(box_planes ? box_planes : np)};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(m, np == 3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (planes) failed with CUresult %d (K=%d rows=%d ld=%ld plane=%ld)", (int)r, K, rows, ld,
              plane);
    return SIM_ERR_CUDA;
  }
  return SIM_OK;
}

template <int BN, int NSTAGE, int NP = 3>
int launch_gemm(const GemmTmaps& tm, float* Y, long ldd, int M, int N, int K, cudaStream_t stream, void* po = nullptr,
                int po_cols = 0, long po_ld = 0, long po_plane = 0, EpiAct act = EpiAct{0, 0, nullptr}) {
  using Cfg = GemmCfg<BN, NSTAGE, NP>;
  auto kern = gemm_split3_kernel<BN, NSTAGE, NP>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, Cfg::SMEM, attr) != cudaSuccess) return check_launch("gemm_split3 attr");
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + kBM - 1) / kBM;
  // split-K when the output tiles cannot fill the SMs and the contraction is long (>= 32 k-blocks = 1024): chunks of at
  // least 8 k-blocks, enough of them to give every SM a CTA
  const int nk = (K + kBK - 1) / kBK, tiles = m_tiles * n_tiles;
  int splits = 1, kps = nk;
  static const int splitk = [] { const char* e = getenv("SIM_GEMM_SPLITK"); return e ? atoi(e) : 1; }();
  if (splitk && !po && !act.mode && tiles * 2 <= 148 && nk >= 32) {
    splits = std::min((148 + tiles - 1) / tiles, nk / 8);
    kps = (nk + splits - 1) / splits;
    splits = (nk + kps - 1) / kps;
  }
  if (splits > 1 && cudaMemset2DAsync(Y, (size_t)ldd * 4, 0, (size_t)N * 4, M, stream) != cudaSuccess)
    return check_launch("gemm_split3 memset");
  kern<<<dim3(tiles, splits), 192, Cfg::SMEM, stream>>>(tm, Y, ldd, M, N, K, n_tiles, static_cast<__nv_bfloat16*>(po),
                                                        po_cols, po_ld, po_plane, kps, act);
  return check_launch("gemm_split3");
}

template <int BN, int NSTAGE, int NP = 3, int MC = 0>
int launch_gemm_persistent(const GemmTmaps& tm, float* Y, long ldd, int M, int N, int K, cudaStream_t stream,
                           EpiAct act = EpiAct{0, 0, nullptr}) {
  using Cfg = GemmPCfg<BN, NSTAGE, NP, MC>;
  auto kern = gemm_split3_persistent_kernel<BN, NSTAGE, NP, MC>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, Cfg::SMEM, attr) != cudaSuccess) return check_launch("gemm_split3 persistent attr");
  static int sm_count[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!sm_count[dev & 63]) cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + kBM - 1) / kBM;
  if constexpr (MC) {
    const int total = ((m_tiles + 1) / 2) * n_tiles;  // pair tiles
    int grid = 2 * total < sm_count[dev & 63] ? 2 * total : (sm_count[dev & 63] & ~1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(Cfg::NT), cfg.dynamicSmemBytes = Cfg::SMEM, cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, tm, Y, ldd, M, N, K, n_tiles, total, act) != cudaSuccess)
      return check_launch("gemm_split3 persistent (CTA pairs)");
    return check_launch("gemm_split3 persistent (CTA pairs)");
  } else {
    const int total = m_tiles * n_tiles;
    const int grid = total < sm_count[dev & 63] ? total : sm_count[dev & 63];
    kern<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(tm, Y, ldd, M, N, K, n_tiles, total, act);
    return check_launch("gemm_split3 persistent");
  }
}

// x = x0 + x1 + x2 with every residual formed exactly in fp32 (the differences are representable)
__global__ void split3_kernel(const float* __restrict__ x, long ld, int rows, int K, __nv_bfloat16* __restrict__ out,
                              long ldo, long plane) {
  const int kq = K / 4;
  const long n = (long)rows * kq;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / kq;
    const int c = (int)(i % kq) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 p[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p[0][j] = __float2bfloat16_rn(f[j]);
      const float r1 = f[j] - __bfloat162float(p[0][j]);
      p[1][j] = __float2bfloat16_rn(r1);
      p[2][j] = __float2bfloat16_rn(r1 - __bfloat162float(p[1][j]));
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) *reinterpret_cast<uint2*>(out + q * plane + r * ldo + c) = *reinterpret_cast<const uint2*>(p[q]);
  }
}

// x = x0 + 2^-11 x1' as two fp16 planes (operand format of the NP = 2 kernels; split2h_store4, common.cuh)
__global__ void split2h_kernel(const float* __restrict__ x, long ld, int rows, int K, __half* __restrict__ out, long ldo,
                               long plane) {
  const int kq = K / 4;
  const long n = (long)rows * kq;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / kq;
    const int c = (int)(i % kq) * 4;
    split2h_store4(out + r * ldo + c, plane, *reinterpret_cast<const float4*>(x + r * ld + c));
  }
}

}  // namespace

int split2_f16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, cudaStream_t stream) {
  SIM_REQUIRE(x && out && rows > 0 && K > 0, SIM_ERR_INVALID, "split2_f16: empty problem / null tensor");
  SIM_REQUIRE(K % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && plane % 4 == 0 && aligned16(x) &&
                  (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
              SIM_ERR_ALIGN, "split2_f16: K, strides must be multiples of 4 and bases 16-byte aligned");
  const long n = (long)rows * (K / 4);
  const int grid = (int)((n + 255) / 256 < 148L * 16 ? (n + 255) / 256 : 148L * 16);
  split2h_kernel<<<grid, 256, 0, stream>>>(x, ld, rows, K, static_cast<__half*>(out), ldo, plane);
  return check_launch("split2_f16");
}

int split3_bf16(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, cudaStream_t stream) {
  SIM_REQUIRE(x && out && rows > 0 && K > 0, SIM_ERR_INVALID, "split3_bf16: empty problem / null tensor");
  SIM_REQUIRE(K % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && plane % 4 == 0 && aligned16(x) &&
                  (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
              SIM_ERR_ALIGN, "split3_bf16: K, strides must be multiples of 4 and bases 16-byte aligned");
  const long n = (long)rows * (K / 4);
  const int grid = (int)((n + 255) / 256 < 148L * 16 ? (n + 255) / 256 : 148L * 16);
  split3_kernel<<<grid, 256, 0, stream>>>(x, ld, rows, K, static_cast<__nv_bfloat16*>(out), ldo, plane);
  return check_launch("split3_bf16");
}

// The three planes of the TRANSPOSE of x (rows, K) -> out[q][k][r]: the operands of the dgrad / wgrad GEMMs of an fp32
// Linear (dX = dY W needs W^T, dW = dY^T X needs dY^T and X^T as K-major operands).  One pass through a 32 x 32 shared-memory
// tile - reads coalesced along K, writes coalesced along rows - instead of a transposing copy followed by sim_split3_bf16.
__global__ void __launch_bounds__(256) split3_t_kernel(const float* __restrict__ x, long ld, int rows, int K,
                                                       __nv_bfloat16* __restrict__ out, long ldo, long plane) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, k = k0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && k < K) ? x[(long)r * ld + k] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty + 8 * i, r = r0 + tx;
    if (k < K && r < rows) {
      const float f = tile[tx][ty + 8 * i];
      const __nv_bfloat16 p0 = __float2bfloat16_rn(f);
      const float r1 = f - __bfloat162float(p0);
      const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
      __nv_bfloat16* o = out + (long)k * ldo + r;
      o[0] = p0, o[plane] = p1, o[2 * plane] = p2;
    }
  }
}

int split3_bf16_t(const float* x, long ld, int rows, int K, void* out, long ldo, long plane, cudaStream_t stream) {
  SIM_REQUIRE(x && out && rows > 0 && K > 0 && ldo >= rows, SIM_ERR_INVALID, "split3_bf16_t: bad arguments");
  const dim3 grid((K + 31) / 32, (rows + 31) / 32);
  SIM_REQUIRE(grid.y <= 65535, SIM_ERR_INVALID, "split3_bf16_t: more than 2 M rows");
  split3_t_kernel<<<grid, 256, 0, stream>>>(x, ld, rows, K, static_cast<__nv_bfloat16*>(out), ldo, plane);
  return check_launch("split3_bf16_t");
}

// Two operand formats (np):
//   3: three bf16 planes, six products - exact for any fp32 operand (bf16 keeps the fp32 exponent range);
//   2: two fp16 planes x0 = fp16(x), x1' = fp16(2^11 (x - x0)), three products x0.w0 + 2^-11 (x0.w1' + x1'.w0) with the
//      scaled pair in the second TMEM accumulator: representation error 2^-22 per operand (measured 7e-8 of the result
//      norm at K = 384, below the 3.5e-7 of an fp32 SGEMM's own accumulation) for HALF the tensor-core work, valid while
//      |x|, |w| < 65504 - the host side selects it only where that bound is provable (in_proj: a LayerNorm output is
//      bounded by sqrt(C) max|gamma| + max|beta|; the weights are checked when they are split).
// act_mode / act_col0 / act_bias: optional epilogue activation (EpiAct), not combined with split-K or planes_out.
int gemm_planes(int np, const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd,
                int M, int N, int K, cudaStream_t stream, void* po, int po_cols, long po_ld, long po_plane, int act_mode,
                int act_col0, const float* act_bias) {
  SIM_REQUIRE(np == 3 || np == 2, SIM_ERR_INVALID, "gemm_planes: np must be 3 (bf16 planes) or 2 (fp16 planes)");
  SIM_REQUIRE(act_mode >= 0 && act_mode <= 2 && (act_mode != 2 || act_bias) && (!act_mode || !po), SIM_ERR_INVALID,
              "gemm_planes: epilogue activation 0 / 1 (silu from act_col0) / 2 (softplus(v + bias), bias required), no planes_out");
  SIM_REQUIRE(np == 3 || !po, SIM_ERR_INVALID, "gemm_planes: planes_out is a bf16 x 3 feature");
  const EpiAct act{act_mode, act_col0, act_bias};
  SIM_REQUIRE(!po || (N <= 64 && po_cols % 4 == 0 && po_cols <= N + 3 && po_ld % 4 == 0 && po_plane % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(po) & 7u) == 0),
              SIM_ERR_INVALID, "gemm_bf16x3: split-plane output is built for N <= 64 (x_proj) and 8-byte aligned planes");
  SIM_REQUIRE(Xs && Ws && Y && M > 0 && N > 0 && K > 0, SIM_ERR_INVALID, "gemm_bf16x3: empty problem / null tensor");
  SIM_REQUIRE(aligned16(Xs) && aligned16(Ws) && aligned16(Y) && ldx % 8 == 0 && ldw % 8 == 0 && xplane % 8 == 0 &&
                  wplane % 8 == 0 && ldd % 4 == 0 && N % 4 == 0,
              SIM_ERR_ALIGN, "gemm_bf16x3: TMA needs 16-byte aligned bases / strides, the epilogue N and ldd multiples of 4");
  // tile width: the kernel time is (rounds of tiles over the SMs) x (tile cost), tile cost ~ bn + a fixed prologue /
  // epilogue share (measured ~40 columns' worth); e.g. in_proj (128 x 1536 outputs per row block): bn 256 -> 6 rounds,
  // 192 -> 7 rounds of 3/4 the cost, 128 -> 11 rounds
  const int m_tiles = (M + kBM - 1) / kBM;
  int bn = 64;
  if (N > 64) {
    long best = -1;
    for (int cand : {256, 192, 128}) {
      const long tiles = (long)m_tiles * ((N + cand - 1) / cand);
      const long cost = ((tiles + 147) / 148) * (cand + 40);
      if (best < 0 || cost < best) best = cost, bn = cand;
    }
  }
  static const int force = [] { const char* e = getenv("SIM_GEMM_CFG"); return e ? atoi(e) : 0; }();  // bench-only override
  if (force == 128 || force == 1282) bn = N <= 64 ? 64 : 128;
  if (force == 256 || force == 192 || force == 96) bn = N <= 64 ? 64 : force;
  // one or two k-blocks (dt_proj, K = 24): the tile is all prologue + epilogue, so prefer two co-resident CTAs per SM
  // that overlap each other's phases (measured 21.3 vs 27.3 us)
  static const int shallow_mode = [] { const char* e = getenv("SIM_GEMM_SHALLOW"); return e ? atoi(e) : 1; }();  // 1: persistent 192-wide tiles (13.6 vs 16.3 us on dt_proj)
  const bool shallow = K <= 2 * kBK && N > 64 && shallow_mode == 0;
  if (shallow) bn = 128;
  if (K <= 2 * kBK && N > 64 && shallow_mode == 1) bn = 192;
  GemmTmaps tm;
  int rc;
  static const int persist = [] { const char* e = getenv("SIM_GEMM_PERSIST"); return e ? atoi(e) : 1; }();
  // CTA pairs with the W tile multicast (persistent 192-wide kernel, enough pair tiles to fill the SMs); 0 = ablation
  // CTA pairs (opt-in, SIM_GEMM_PAIRS): 1 = W tile multicast into both CTAs, 2 = one MMA stream per pair (cta_group::2, M = 256,
  // each CTA loads half of W).  Both are bit-identical to the default and both measured NEUTRAL at the in_proj / out_proj
  // shapes (profiles/r02_gemm_ncu.md: 55.2 / 56.7 vs 54.7 us, 83.1 / 86.7 vs 84.6 us): the main loop is bound by the bytes
  // that enter each SM (~42 B/clk/SM, the TMA / L2 -> SM ingest rate), which multicast does not change and which the pair's
  // operand exchange re-spends; so the default stays the plain one-CTA-per-SM kernel.
  static const int pairs_on = [] { const char* e = getenv("SIM_GEMM_PAIRS"); return e ? atoi(e) : 0; }();
  if (np == 2 && bn == 96) bn = 128;
  const bool persist192 = bn == 192 && persist && (long)m_tiles * ((N + 191) / 192) >= 148;
  const int pairs = persist192 && K > 2 * kBK ? pairs_on : 0;
  if ((rc = make_tmap_planes(&tm.x, Xs, K, M, ldx, xplane, kBM, np))) return rc;
  if ((rc = make_tmap_planes(&tm.w, Ws, K, N, ldw, wplane, pairs ? bn / 2 : bn, np, pairs == 1 ? 1 : 0))) return rc;
  if (np == 2) {
    switch (bn) {
      case 64: return launch_gemm<64, 6, 2>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
      case 128: return launch_gemm<128, 5, 2>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
      case 192:
        if (pairs == 2) return launch_gemm_persistent<192, 6, 2, 2>(tm, Y, ldd, M, N, K, stream, act);
        if (pairs == 1) return launch_gemm_persistent<192, 4, 2, 1>(tm, Y, ldd, M, N, K, stream, act);
        if (persist192) return launch_gemm_persistent<192, 4, 2>(tm, Y, ldd, M, N, K, stream, act);
        return launch_gemm<192, 4, 2>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
      default: return launch_gemm<256, 4, 2>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
    }
  }
  switch (bn) {
    case 64: return launch_gemm<64, 5>(tm, Y, ldd, M, N, K, stream, po, po_cols, po_ld, po_plane, act);
    case 128:
      return (force == 1282 || shallow) ? launch_gemm<128, 2>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act)
                                        : launch_gemm<128, 4>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
    case 96: return launch_gemm<96, 2>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
    case 192:
      if (pairs == 2) return launch_gemm_persistent<192, 4, 3, 2>(tm, Y, ldd, M, N, K, stream, act);
      if (pairs == 1) return launch_gemm_persistent<192, 3, 3, 1>(tm, Y, ldd, M, N, K, stream, act);
      if (persist192) return launch_gemm_persistent<192, 3>(tm, Y, ldd, M, N, K, stream, act);
      return launch_gemm<192, 3>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
    default: return launch_gemm<256, 3>(tm, Y, ldd, M, N, K, stream, nullptr, 0, 0, 0, act);
  }
}

int gemm_bf16x3(const void* Xs, long ldx, long xplane, const void* Ws, long ldw, long wplane, float* Y, long ldd, int M,
                int N, int K, cudaStream_t stream, void* po, int po_cols, long po_ld, long po_plane) {
  return gemm_planes(3, Xs, ldx, xplane, Ws, ldw, wplane, Y, ldd, M, N, K, stream, po, po_cols, po_ld, po_plane, 0, 0, nullptr);
}

int gemm_f32a_bf16x3(const float* X, long ldx, const void* Ws, long ldw, long wplane, float* Y, long ldd, int M, int N,
                     int K, cudaStream_t stream, void* po, int po_cols, long po_ld, long po_plane, const float* conv_w,
                     const float* conv_b, float* U, long ldu, int batch, int L) {
  SIM_REQUIRE(X && Ws && Y && M > 0 && N > 0 && N <= 64 && K > 0, SIM_ERR_INVALID,
              "gemm_f32a_bf16x3: built for N <= 64 (x_proj); empty problem / null tensor");
  SIM_REQUIRE(aligned16(X) && aligned16(Ws) && aligned16(Y) && ldx % 4 == 0 && ldw % 8 == 0 && wplane % 8 == 0 && ldd % 4 == 0 &&
                  N % 4 == 0,
              SIM_ERR_ALIGN, "gemm_f32a_bf16x3: TMA needs 16-byte aligned bases / strides, the epilogue N and ldd multiples of 4");
  SIM_REQUIRE(!po || (po_cols % 4 == 0 && po_ld % 4 == 0 && po_plane % 4 == 0 && (reinterpret_cast<uintptr_t>(po) & 7u) == 0),
              SIM_ERR_ALIGN, "gemm_f32a_bf16x3: split-plane output needs 8-byte aligned planes");
  const bool conv = conv_w != nullptr;
  SIM_REQUIRE(!conv || (U && batch > 0 && L > 0 && (long)batch * L == M && K % 8 == 0 && ldu % 4 == 0 && aligned16(U) &&
                        (size_t)K * 5 * 4 <= 20 * 1024),
              SIM_ERR_INVALID, "gemm_f32a_bf16x3: fused conv needs u, batch * L == M, K %% 8 == 0 and K <= 1024");
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SIM_ERR_CUDA;
  }
  GemmF32ATmaps tm;
  {  // fp32 activation, 128-byte swizzle: flat (K, M, 1) rows, or per cloud (K, L, batch) with the conv's 3-row halo
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)(conv ? L : M), (cuuint64_t)(conv ? batch : 1)};
    cuuint64_t gstr[2] = {(cuuint64_t)ldx * 4, (cuuint64_t)ldx * 4 * (cuuint64_t)(conv ? L : M)};
    cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)(kBM + (conv ? kCW - 1 : 0)), 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(&tm.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (fp32 A) failed with CUresult %d (K=%d M=%d ld=%ld)", (int)r, K, M, ldx);
      return SIM_ERR_CUDA;
    }
  }
  int rc;
  if ((rc = make_tmap_planes(&tm.w, Ws, K, N, ldw, wplane, 64))) return rc;
  static SmemAttrCache attr0, attr1;
  if (conv) {
    using Cfg = GemmF32ACfg<64, true>;
    auto kern = gemm_f32a_split3_kernel<64, true>;
    const size_t smem = Cfg::NSTAGE * Cfg::STAGE + 1024 + 256 + (size_t)K * 5 * 4;
    if (ensure_dyn_smem(kern, smem, attr1) != cudaSuccess) return check_launch("gemm_f32a_split3 attr");
    const int tpc = (L + kBM - 1) / kBM;
    kern<<<batch * tpc, 576, smem, stream>>>(tm, Y, ldd, M, N, K, static_cast<__nv_bfloat16*>(po), po_cols, po_ld, po_plane,
                                             conv_w, conv_b, U, ldu, L, tpc);
  } else {
    using Cfg = GemmF32ACfg<64, false>;
    auto kern = gemm_f32a_split3_kernel<64, false>;
    const size_t smem = Cfg::NSTAGE * Cfg::STAGE + 1024 + 256;
    if (ensure_dyn_smem(kern, smem, attr0) != cudaSuccess) return check_launch("gemm_f32a_split3 attr");
    kern<<<(M + kBM - 1) / kBM, 576, smem, stream>>>(tm, Y, ldd, M, N, K, static_cast<__nv_bfloat16*>(po), po_cols, po_ld,
                                                     po_plane, nullptr, nullptr, nullptr, 0, 0, 1);
  }
  return check_launch("gemm_f32a_split3");
}

}  // namespace sim
