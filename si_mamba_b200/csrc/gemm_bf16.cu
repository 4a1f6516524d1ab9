// bf16 projection GEMM on the 5th-gen tensor cores, hand-written for sm_100a (TMA -> shared memory -> tcgen05.mma with the
// accumulator in TMEM -> tcgen05.ld epilogue): the projections of Mamba.forward (models/block.py:72) under the reference's
// bf16 autocast (tools/runner_pretrain.py:243) and all three GEMMs of their backward, straight from the tensors autograd
// holds - no transposed copies:
//
//   forward   Y[M,N]  = X[M,K] . W[N,K]^T        A = X  K-major,   B = W   K-major
//   dgrad     dX[M,K] = dY[M,N] . W[N,K]         A = dY K-major,   B = W   MN-major (its rows are the contraction index)
//   wgrad     dW[N,K] = dY[M,N]^T . X[M,K]       A = dY MN-major,  B = X   MN-major (contraction over the M tokens), split-K
//
// "K-major" = the contraction index is the contiguous one (a row-major operand whose rows are output rows / columns);
// "MN-major" = the output index is the contiguous one.  tcgen05.mma reads either straight from shared memory: a K-major
// tile is rows x 64 bf16 (one 128-byte swizzle span per row), an MN-major tile is 64-column chunks of 64 contraction rows
// x 128 bytes; the instruction descriptor's a_major / b_major bits say which.  Both arrive by 2-D tensor copies with
// SWIZZLE_128B (rows / columns past the tensor read as zero), so M, N and K need no padding.
//
// One CTA per 128 x BN output tile, 6 warps: warp 0 = TMA producer (one thread), warp 1 = TMEM allocator + MMA issuer (one
// thread, 4 MMAs of K = 16 per 64-deep stage), warps 2-5 = epilogue (TMEM -> registers -> padded staging -> 16-byte /
// 8-byte global stores, fp32 or bf16; split-K partial tiles are added with 16-byte vector atomics).
//
// Roofline: tensor pipe, 2 M N K bf16 flops (MEASURED_PEAKS.json bf16_tflops).

#include <stdlib.h>

#include <algorithm>

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

namespace {

constexpr int kBM = 128;  // output rows per CTA == TMEM lanes
constexpr int kBK = 64;   // contraction elements per pipeline stage (one 128-byte swizzle span of bf16)
constexpr int kUK = 16;   // K of one tcgen05.mma.kind::f16

template <int BN, int NSTAGE_>
struct Cfg {
  static constexpr int A_TILE = kBM * kBK * 2;
  static constexpr int B_TILE = BN * kBK * 2;
  static constexpr int STAGE = A_TILE + B_TILE;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int STG_LD = 36;  // padded row stride (floats) of the epilogue staging tile
  static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
  static_assert(4 * 32 * STG_LD * 4 <= STAGE, "epilogue staging reuses pipeline stage 0");
  static_assert(A_TILE % 1024 == 0 && B_TILE % 1024 == 0, "swizzle-128B tiles need 1024-byte aligned bases");
  static_assert(SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
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

// Shared-memory matrix descriptors, SWIZZLE_128B (layout type 2 in bits [61,64)), descriptor version 1 in bits [46,48).
// Fields in 16-byte units: start address bits [0,14), leading byte offset bits [16,30), stride byte offset bits [32,46)
// (cute/arch/mma_sm100_desc.hpp; canonical layouts in cute/atom/mma_traits_sm100.hpp).
//   K-major : rows of 128 B, groups of 8 rows 1024 B apart (stride offset); the leading offset is not used (1).
//   MN-major: 64-column chunks, each 64 contraction rows x 128 B; groups of 8 contraction rows are 1024 B apart (stride
//             offset), the next 64-column chunk follows `chunk_bytes` later (leading offset).
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t smem_addr, uint32_t chunk_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((chunk_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D fp32 (bits [4,6) = 1), A and B bf16 (bits [7,10), [10,13) = 1), a_major bit 15, b_major bit 16
// (0 = K-major, 1 = MN-major), N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Tmaps {
  CUtensorMap a, b;
};

template <int BN, int NSTAGE_, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
    gemm_bf16_kernel(const __grid_constant__ Tmaps tm, void* __restrict__ Yv, long ldd, int M, int N, int K, int n_tiles,
                     int out_bf16, int kb_per_split) {
  using C = Cfg<BN, NSTAGE_>;
  constexpr int NSTAGE = C::NSTAGE;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base_u32 = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (base_u32 & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * C::STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* accum_full = empty + NSTAGE;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x / n_tiles) * kBM;
  const int n0 = (blockIdx.x % n_tiles) * BN;
  const int nk_all = (K + kBK - 1) / kBK;
  const int kb0 = blockIdx.y * kb_per_split;
  const int nk = gridDim.y > 1 ? min(kb_per_split, nk_all - kb0) : nk_all;
  const bool accumulate_out = gridDim.y > 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_mbar_init();
  }
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        if (kb >= NSTAGE) mbar_wait(&empty[s], ((kb / NSTAGE) - 1) & 1);
        unsigned char* st = smem + s * C::STAGE;
        const int k0 = (kb0 + kb) * kBK;
        mbar_arrive_expect_tx(&full[s], C::STAGE);
        if constexpr (A_MN) {  // map dims (output rows [contiguous], contraction rows): one copy per 64-column chunk
#pragma unroll
          for (int c = 0; c < kBM / 64; ++c) tma_load_2d(st + c * 64 * 128, &tm.a, m0 + c * 64, k0, &full[s]);
        } else {               // map dims (contraction [contiguous], output rows)
          tma_load_2d(st, &tm.a, k0, m0, &full[s]);
        }
        if constexpr (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(st + C::A_TILE + c * 64 * 128, &tm.b, n0 + c * 64, k0, &full[s]);
        } else {
          tma_load_2d(st + C::A_TILE, &tm.b, k0, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(kBM, BN, A_MN, B_MN);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        mbar_wait(&full[s], (kb / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * C::STAGE);
        const uint32_t b0 = a0 + C::A_TILE;
#pragma unroll
        for (int k = 0; k < kBK / kUK; ++k) {
          // K-major: 16 contraction elements = 32 bytes further along the swizzled row; MN-major: 16 contraction rows
          const uint64_t da = A_MN ? desc_mnmajor_sw128(a0 + k * kUK * 128, 64 * 128) : desc_kmajor_sw128(a0 + k * kUK * 2);
          const uint64_t db = B_MN ? desc_mnmajor_sw128(b0 + k * kUK * 128, 64 * 128) : desc_kmajor_sw128(b0 + k * kUK * 2);
          umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum_full);
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = rows m0 + 32 (w % 4) + lane
    const int quad = warp & 3;
    mbar_wait(accum_full, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + quad * 32 * C::STG_LD;  // stage 0 is idle now
    float* Yf = static_cast<float*>(Yv);
    __nv_bfloat16* Yh = static_cast<__nv_bfloat16*>(Yv);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= N) break;
      float v[32];
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + c * 32, v);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stg + lane * C::STG_LD + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int row = rr * 4 + (lane >> 3), col = (lane & 7) * 4;
        const int gm = m0 + quad * 32 + row, gn = n0 + c * 32 + col;
        if (gm < M && gn < N) {
          const float4 o = *reinterpret_cast<const float4*>(stg + row * C::STG_LD + col);
          if (out_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned*>(&lo);
            pk.y = *reinterpret_cast<const unsigned*>(&hi);
            *reinterpret_cast<uint2*>(Yh + (long)gm * ldd + gn) = pk;
          } else if (accumulate_out) {
            atomicAdd(reinterpret_cast<float4*>(Yf + (long)gm * ldd + gn), o);
          } else {
            *reinterpret_cast<float4*>(Yf + (long)gm * ldd + gn) = o;
          }
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
// Persistent variant (everything that is not split-K).  One CTA per SM walks tiles j, j + gridDim.x, ...: the shared-memory
// pipeline keeps rolling across tiles, the accumulator is DOUBLE-BUFFERED in tensor memory (2 x BN columns), and eight
// epilogue warps (two per TMEM lane quadrant, half of the tile's columns each) drain tile i while the MMA warp is already
// filling the other buffer with tile i + 1.  The one-tile-per-CTA kernel above paid barrier init, TMEM allocation, the
// first TMA round trip and an un-overlapped epilogue per tile: 3.05 ms against cuBLAS' 2.39 ms on the bf16 C1 forward.
template <int BN, int NSTAGE_>
struct PCfg {
  static constexpr int A_TILE = kBM * kBK * 2;
  static constexpr int B_TILE = BN * kBK * 2;
  static constexpr int STAGE = A_TILE + B_TILE;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int NEPI = 8;
  static constexpr int NT = 64 + 32 * NEPI;
  static constexpr int STG_LD = 36;
  static constexpr int STG = NEPI * 32 * STG_LD * 4;
  static constexpr int SMEM = NSTAGE * STAGE + STG + 1024 + 256;
  static constexpr int CW = BN / 2;  // columns per epilogue warp
  static_assert(CW % 32 == 0 && 2 * BN <= 512, "two column halves of whole 32-column chunks; two accumulators in TMEM");
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int BN, int NSTAGE_, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(PCfg<BN, NSTAGE_>::NT, 1)
    gemm_bf16_persistent_kernel(const __grid_constant__ Tmaps tm, void* __restrict__ Yv, long ldd, int M, int N, int K,
                                int n_tiles, int total_tiles, int out_bf16) {
  using C = PCfg<BN, NSTAGE_>;
  constexpr int NSTAGE = C::NSTAGE, CW = C::CW;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base_u32 = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (base_u32 & 1023u)) & 1023u);
  float* stg_base = reinterpret_cast<float*>(smem + NSTAGE * C::STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * C::STAGE + C::STG);
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;   // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (K + kBK - 1) / kBK;
  constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&acc_full[q], 1);
      mbar_init(&acc_empty[q], C::NEPI);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NSTAGE;
          if (it >= NSTAGE) mbar_wait(&empty[s], ((it / NSTAGE) - 1) & 1);
          unsigned char* st = smem + s * C::STAGE;
          const int k0 = kb * kBK;
          mbar_arrive_expect_tx(&full[s], C::STAGE);
          if constexpr (A_MN) {
#pragma unroll
            for (int c = 0; c < kBM / 64; ++c) tma_load_2d(st + c * 64 * 128, &tm.a, m0 + c * 64, k0, &full[s]);
          } else {
            tma_load_2d(st, &tm.a, k0, m0, &full[s]);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(st + C::A_TILE + c * 64 * 128, &tm.b, n0 + c * 64, k0, &full[s]);
          } else {
            tma_load_2d(st + C::A_TILE, &tm.b, k0, n0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(kBM, BN, A_MN, B_MN);
      int it = 0, i = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
        const int buf = i & 1;
        if (i >= 2) {  // the epilogue warps have drained this accumulator (tile i - 2)
          mbar_wait(&acc_empty[buf], ((i >> 1) - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(&full[s], (it / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * C::STAGE);
          const uint32_t b0 = a0 + C::A_TILE;
#pragma unroll
          for (int k = 0; k < kBK / kUK; ++k) {
            const uint64_t da = A_MN ? desc_mnmajor_sw128(a0 + k * kUK * 128, 64 * 128) : desc_kmajor_sw128(a0 + k * kUK * 2);
            const uint64_t db = B_MN ? desc_mnmajor_sw128(b0 + k * kUK * 128, 64 * 128) : desc_kmajor_sw128(b0 + k * kUK * 2);
            umma_bf16(tmem_d + buf * BN, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===== eight epilogue warps: TMEM lane quadrant (warp & 3), column half (ew >> 2)
    const int ew = warp - 2;
    const int quad = warp & 3, half = ew >> 2;
    float* stg = stg_base + ew * 32 * C::STG_LD;
    float* Yf = static_cast<float*>(Yv);
    __nv_bfloat16* Yh = static_cast<__nv_bfloat16*>(Yv);
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++i) {
      const int buf = i & 1;
      const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * BN;
      mbar_wait(&acc_full[buf], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < CW / 32; ++c) {
        const int cb = half * CW + c * 32;
        if (n0 + cb >= N) break;
        float v[32];
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + buf * BN + cb, v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(stg + lane * C::STG_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int row = rr * 4 + (lane >> 3), col = (lane & 7) * 4;
          const int gm = m0 + quad * 32 + row, gn = n0 + cb + col;
          if (gm < M && gn < N) {
            const float4 o = *reinterpret_cast<const float4*>(stg + row * C::STG_LD + col);
            if (out_bf16) {
              const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
              uint2 pk;
              pk.x = *reinterpret_cast<const unsigned*>(&lo);
              pk.y = *reinterpret_cast<const unsigned*>(&hi);
              *reinterpret_cast<uint2*>(Yh + (long)gm * ldd + gn) = pk;
            } else {
              *reinterpret_cast<float4*>(Yf + (long)gm * ldd + gn) = o;
            }
          }
        }
        __syncwarp();
      }
      // this warp's TMEM reads of the buffer are complete (tcgen05.wait::ld inside tmem_ld32): hand it back
      tc_fence_before();
      if (lane == 0) mbar_arrive1(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, kTmemCols);
  }
}

// 2-D bf16 tensor map with SWIZZLE_128B: dims (inner [contiguous], outer), outer stride ld elements, box (64, box_outer)
int make_tmap_2d(CUtensorMap* m, const void* base, long inner, long outer, long ld, int box_outer) {
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SIM_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 2-D) failed with CUresult %d (inner=%ld outer=%ld ld=%ld box=%d)", (int)r, inner,
              outer, ld, box_outer);
    return SIM_ERR_CUDA;
  }
  return SIM_OK;
}

template <int BN, int NSTAGE, bool A_MN, bool B_MN>
int launch(const Tmaps& tm, void* Y, long ldd, int M, int N, int K, int out_bf16, int splits, cudaStream_t stream) {
  using C = Cfg<BN, NSTAGE>;
  auto kern = gemm_bf16_kernel<BN, NSTAGE, A_MN, B_MN>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, C::SMEM, attr) != cudaSuccess) return check_launch("gemm_bf16 attr");
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + kBM - 1) / kBM;
  const int nk_all = (K + kBK - 1) / kBK;
  const int kb_per_split = (nk_all + splits - 1) / splits;
  const int ny = (nk_all + kb_per_split - 1) / kb_per_split;
  kern<<<dim3(m_tiles * n_tiles, ny), 192, C::SMEM, stream>>>(tm, Y, ldd, M, N, K, n_tiles, out_bf16, kb_per_split);
  return check_launch("gemm_bf16");
}

template <int BN, int NSTAGE, bool A_MN, bool B_MN>
int launch_persistent(const Tmaps& tm, void* Y, long ldd, int M, int N, int K, int out_bf16, cudaStream_t stream) {
  using C = PCfg<BN, NSTAGE>;
  auto kern = gemm_bf16_persistent_kernel<BN, NSTAGE, A_MN, B_MN>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, C::SMEM, attr) != cudaSuccess) return check_launch("gemm_bf16_persistent attr");
  const int n_tiles = (N + BN - 1) / BN;
  const int total = ((M + kBM - 1) / kBM) * n_tiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static int sm_count[64] = {};
  if (!sm_count[dev & 63]) cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  if (sm_count[dev & 63] > 0) sms = sm_count[dev & 63];
  kern<<<std::min(total, sms), C::NT, C::SMEM, stream>>>(tm, Y, ldd, M, N, K, n_tiles, total, out_bf16);
  return check_launch("gemm_bf16_persistent");
}

template <bool A_MN, bool B_MN>
int dispatch(const Tmaps& tm, void* Y, long ldd, int M, int N, int K, int out_bf16, int splits, int bn, cudaStream_t stream) {
  static const int persist = [] { const char* e = getenv("SIM_GEMM_BF16_PERSIST"); return e ? atoi(e) : 1; }();
  if (splits == 1 && persist) {
    switch (bn) {
      case 64: return launch_persistent<64, 6, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, stream);
      case 128: return launch_persistent<128, 5, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, stream);
      case 192: return launch_persistent<192, 4, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, stream);
      default: return launch_persistent<256, 3, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, stream);
    }
  }
  switch (bn) {
    case 64: return launch<64, 6, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, splits, stream);
    case 128: return launch<128, 5, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, splits, stream);
    default: return launch<256, 4, A_MN, B_MN>(tm, Y, ldd, M, N, K, out_bf16, splits, stream);
  }
}

}  // namespace

// Y[M,N] = op(A) . op(B)^T with bf16 operands and fp32 accumulation.
//   a_mn = 0: A is (M, K) row-major with row stride lda (K-major);  a_mn = 1: A is (K, M) row-major (MN-major)
//   b_mn = 0: B is (N, K) row-major with row stride ldb (K-major);  b_mn = 1: B is (K, N) row-major (MN-major)
//   out_bf16: Y bf16 (row stride ldd), else fp32.  splits > 1: split-K, partial tiles are ADDED to a zeroed fp32 Y.
int gemm_bf16(const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* Y, long ldd, int out_bf16, int M,
              int N, int K, int splits, cudaStream_t stream) {
  SIM_REQUIRE(A && B && Y && M > 0 && N > 0 && K > 0, SIM_ERR_INVALID, "gemm_bf16: empty problem / null tensor");
  SIM_REQUIRE(aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0, SIM_ERR_ALIGN,
              "gemm_bf16: TMA needs 16-byte aligned operand bases and row strides (lda=%ld ldb=%ld)", lda, ldb);
  SIM_REQUIRE(N % 4 == 0 && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & (out_bf16 ? 7u : 15u)) == 0, SIM_ERR_ALIGN,
              "gemm_bf16: the epilogue stores 4 columns at a time (N=%d ldd=%ld)", N, ldd);
  const int m_tiles = (M + kBM - 1) / kBM;
  if (splits <= 0) {
    // automatic split-K (weight gradients: few output tiles, a contraction over every token): about two waves of CTAs, at
    // least four 64-deep k-blocks per CTA.  The caller zeroes the fp32 output.
    const long tiles = (long)m_tiles * ((N + 127) / 128);
    const int nk_all = (K + kBK - 1) / kBK;
    splits = (int)std::max<long>(1, std::min<long>(296 / tiles, nk_all / 4));  // at most two full waves of CTAs
  }
  SIM_REQUIRE(splits == 1 || !out_bf16, SIM_ERR_INVALID, "gemm_bf16: split-K accumulates into an fp32 output");
  int bn = 64;
  if (N > 64) {
    long best = -1;
    static const int force = [] { const char* e = getenv("SIM_GEMM_BF16_BN"); return e ? atoi(e) : 0; }();  // bench override
    for (int cand : {256, 192, 128}) {
      if (cand == 192 && splits != 1) continue;  // the split-K kernel is built for 64 / 128 / 256
      if (force && cand != force) continue;
      const long tiles = (long)m_tiles * ((N + cand - 1) / cand) * splits;
      const long cost = ((tiles + 147) / 148) * (cand + 40);
      if (best < 0 || cost < best) best = cost, bn = cand;
    }
  }
  Tmaps tm;
  int rc;
  if (a_mn) rc = make_tmap_2d(&tm.a, A, M, K, lda, kBK); else rc = make_tmap_2d(&tm.a, A, K, M, lda, kBM);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_2d(&tm.b, B, N, K, ldb, kBK); else rc = make_tmap_2d(&tm.b, B, K, N, ldb, bn);
  if (rc) return rc;
  if (a_mn) return b_mn ? dispatch<true, true>(tm, Y, ldd, M, N, K, out_bf16, splits, bn, stream)
                        : dispatch<true, false>(tm, Y, ldd, M, N, K, out_bf16, splits, bn, stream);
  return b_mn ? dispatch<false, true>(tm, Y, ldd, M, N, K, out_bf16, splits, bn, stream)
              : dispatch<false, false>(tm, Y, ldd, M, N, K, out_bf16, splits, bn, stream);
}

}  // namespace sim
