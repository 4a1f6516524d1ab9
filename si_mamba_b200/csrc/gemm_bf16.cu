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
// A pipeline stage holds one 128-byte swizzle span of contraction per row: 64 bf16 or 32 fp32 (TF32) elements, consumed by
// four tcgen05.mma of 32 bytes of K each (K = 16 for kind::f16, K = 8 for kind::tf32).  In bytes the two element types share
// every tile size and descriptor; TF32 = false / true selects the element size, the MMA kind and the operand formats.
template <bool TF32>
struct El {
  static constexpr int ES = TF32 ? 4 : 2;     // bytes per element
  static constexpr int BK = 128 / ES;         // contraction elements per stage == columns of one MN-major chunk
  static constexpr int UK = 32 / ES;          // K of one MMA
};

template <int BN, int NSTAGE_>
struct Cfg {
  static constexpr int A_TILE = kBM * 128;
  static constexpr int B_TILE = BN * 128;
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
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
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
// Instruction descriptor: D fp32 (bits [4,6) = 1), A and B format in bits [7,10), [10,13) (1 = bf16, 2 = tf32), a_major bit
// 15, b_major bit 16 (0 = K-major, 1 = MN-major), N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t idesc_tc(int M, int N, bool a_mn, bool b_mn, bool tf32) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// optional epilogue: + bias[column], ReLU; and, for the per-patch PointNet of Encoder.forward (models/point_mamba.py:59-73,
// rows = points, 32 consecutive rows = one patch = the 32 TMEM lanes of one epilogue warp): + gbias[row / 32][column] (the
// patch's pooled global feature pushed through its half of the next conv: the `cat([global.expand, local])` of :67-68 without
// the cat) and gmax[row / 32][column] = max over the patch's rows (the `torch.max(feature, dim=2)` of :66 / :72), written
// instead of or next to Y.  Persistent kernel only.
struct Epi {
  const float* bias;
  int relu;
  const float* gbias = nullptr;
  long ld_gbias = 0;
  float* gmax = nullptr;
  long ld_gmax = 0;
  int silu_col0 = -1;  // >= 0: columns from here on leave as silu(v) (in_proj hands the scan the gate, see gemm_split3.cu EpiAct)
};
constexpr int kGroupRows = 32;
__device__ __forceinline__ float4 apply_epi(float4 o, const Epi& e, int gn) {
  if (e.bias) {
    const float4 b = *reinterpret_cast<const float4*>(e.bias + gn);
    o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
  }
  if (e.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
  if (e.silu_col0 >= 0 && gn >= e.silu_col0) o = make_float4(silu_f(o.x), silu_f(o.y), silu_f(o.z), silu_f(o.w));
  return o;
}

struct Tmaps {
  CUtensorMap a, b;
};

template <int BN, int NSTAGE_, bool A_MN, bool B_MN, bool TF32>
__global__ void __launch_bounds__(192, 1)
    gemm_bf16_kernel(const __grid_constant__ Tmaps tm, void* __restrict__ Yv, long ldd, int M, int N, int K, int n_tiles,
                     int out_bf16, int kb_per_split, const Epi epi) {
  using C = Cfg<BN, NSTAGE_>;
  constexpr int NSTAGE = C::NSTAGE;
  constexpr int kBK = El<TF32>::BK, kUK = El<TF32>::UK;
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
          for (int c = 0; c < kBM / kBK; ++c) tma_load_2d(st + c * kBK * 128, &tm.a, m0 + c * kBK, k0, &full[s]);
        } else {               // map dims (contraction [contiguous], output rows)
          tma_load_2d(st, &tm.a, k0, m0, &full[s]);
        }
        if constexpr (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / kBK; ++c) tma_load_2d(st + C::A_TILE + c * kBK * 128, &tm.b, n0 + c * kBK, k0, &full[s]);
        } else {
          tma_load_2d(st + C::A_TILE, &tm.b, k0, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tc(kBM, BN, A_MN, B_MN, TF32);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NSTAGE;
        mbar_wait(&full[s], (kb / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * C::STAGE);
        const uint32_t b0 = a0 + C::A_TILE;
#pragma unroll
        for (int k = 0; k < kBK / kUK; ++k) {
          // K-major: 16 contraction elements = 32 bytes further along the swizzled row; MN-major: 16 contraction rows
          const uint64_t da = A_MN ? desc_mnmajor_sw128(a0 + k * kUK * 128, kBK * 128) : desc_kmajor_sw128(a0 + k * 32);
          const uint64_t db = B_MN ? desc_mnmajor_sw128(b0 + k * kUK * 128, kBK * 128) : desc_kmajor_sw128(b0 + k * 32);
          umma<TF32>(tmem_d, da, db, idesc, (kb | k) != 0);
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
          const float4 o = apply_epi(*reinterpret_cast<const float4*>(stg + row * C::STG_LD + col), epi, gn);
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
  static constexpr int A_TILE = kBM * 128;
  static constexpr int B_TILE = BN * 128;
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

template <int BN, int NSTAGE_, bool A_MN, bool B_MN, bool TF32>
__global__ void __launch_bounds__(PCfg<BN, NSTAGE_>::NT, 1)
    gemm_bf16_persistent_kernel(const __grid_constant__ Tmaps tm, void* __restrict__ Yv, long ldd, int M, int N, int K,
                                int n_tiles, int total_tiles, int out_bf16, const Epi epi) {
  using C = PCfg<BN, NSTAGE_>;
  constexpr int NSTAGE = C::NSTAGE, CW = C::CW;
  constexpr int kBK = El<TF32>::BK, kUK = El<TF32>::UK;
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
            for (int c = 0; c < kBM / kBK; ++c) tma_load_2d(st + c * kBK * 128, &tm.a, m0 + c * kBK, k0, &full[s]);
          } else {
            tma_load_2d(st, &tm.a, k0, m0, &full[s]);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / kBK; ++c) tma_load_2d(st + C::A_TILE + c * kBK * 128, &tm.b, n0 + c * kBK, k0, &full[s]);
          } else {
            tma_load_2d(st + C::A_TILE, &tm.b, k0, n0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tc(kBM, BN, A_MN, B_MN, TF32);
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
            const uint64_t da = A_MN ? desc_mnmajor_sw128(a0 + k * kUK * 128, kBK * 128) : desc_kmajor_sw128(a0 + k * 32);
            const uint64_t db = B_MN ? desc_mnmajor_sw128(b0 + k * kUK * 128, kBK * 128) : desc_kmajor_sw128(b0 + k * 32);
            umma<TF32>(tmem_d + buf * BN, da, db, idesc, (kb | k) != 0);
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
        const bool grouped = epi.gbias != nullptr || epi.gmax != nullptr;
        if (grouped) {
          // this warp's 32 rows are one patch: bias / patch bias / ReLU go on the registers (one broadcast row of each),
          // so the staged tile already holds the final values for both the row stores and the column maxima
          const long grow = (m0 + quad * 32) / kGroupRows;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int gn = n0 + cb + 4 * q;
            float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gn < N) {
              if (epi.bias) add = *reinterpret_cast<const float4*>(epi.bias + gn);
              if (epi.gbias && m0 + quad * 32 < M) {
                const float4 g = *reinterpret_cast<const float4*>(epi.gbias + grow * epi.ld_gbias + gn);
                add.x += g.x, add.y += g.y, add.z += g.z, add.w += g.w;
              }
            }
            v[4 * q] += add.x, v[4 * q + 1] += add.y, v[4 * q + 2] += add.z, v[4 * q + 3] += add.w;
            if (epi.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[4 * q + e] = fmaxf(v[4 * q + e], 0.f);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(stg + lane * C::STG_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        if (epi.gmax) {
          // lane l takes column l of the staged 32 x 32 tile (consecutive lanes -> consecutive banks): max over the patch
          float mx = stg[lane];
#pragma unroll
          for (int r = 1; r < 32; ++r) mx = fmaxf(mx, stg[r * C::STG_LD + lane]);
          const int gn = n0 + cb + lane;
          if (m0 + quad * 32 < M && gn < N) epi.gmax[((long)(m0 + quad * 32) / kGroupRows) * epi.ld_gmax + gn] = mx;
        }
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          if (Yv == nullptr) break;
          const int row = rr * 4 + (lane >> 3), col = (lane & 7) * 4;
          const int gm = m0 + quad * 32 + row, gn = n0 + cb + col;
          if (gm < M && gn < N) {
            const float4 staged = *reinterpret_cast<const float4*>(stg + row * C::STG_LD + col);
            const float4 o = grouped ? staged : apply_epi(staged, epi, gn);
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

// 2-D tensor map with SWIZZLE_128B: dims (inner [contiguous], outer), outer stride ld elements, box (128 bytes, box_outer)
int make_tmap_2d(CUtensorMap* m, const void* base, long inner, long outer, long ld, int box_outer, bool tf32) {
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SIM_ERR_CUDA;
  }
  const int es = tf32 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (2-D operand) failed with CUresult %d (inner=%ld outer=%ld ld=%ld box=%d es=%d)", (int)r,
              inner, outer, ld, box_outer, es);
    return SIM_ERR_CUDA;
  }
  return SIM_OK;
}

struct Problem {
  Tmaps tm;
  void* Y;
  long ldd;
  int M, N, K, out_bf16, splits;
  Epi epi;
};

template <int BN, int NSTAGE, bool A_MN, bool B_MN, bool TF32>
int launch(const Problem& q, cudaStream_t stream) {
  using C = Cfg<BN, NSTAGE>;
  auto kern = gemm_bf16_kernel<BN, NSTAGE, A_MN, B_MN, TF32>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, C::SMEM, attr) != cudaSuccess) return check_launch("gemm_tc attr");
  const int n_tiles = (q.N + BN - 1) / BN;
  const int m_tiles = (q.M + kBM - 1) / kBM;
  const int nk_all = (q.K + El<TF32>::BK - 1) / El<TF32>::BK;
  const int kb_per_split = (nk_all + q.splits - 1) / q.splits;
  const int ny = (nk_all + kb_per_split - 1) / kb_per_split;
  kern<<<dim3(m_tiles * n_tiles, ny), 192, C::SMEM, stream>>>(q.tm, q.Y, q.ldd, q.M, q.N, q.K, n_tiles, q.out_bf16, kb_per_split,
                                                               q.epi);
  return check_launch("gemm_tc");
}

template <int BN, int NSTAGE, bool A_MN, bool B_MN, bool TF32>
int launch_persistent(const Problem& q, cudaStream_t stream) {
  using C = PCfg<BN, NSTAGE>;
  auto kern = gemm_bf16_persistent_kernel<BN, NSTAGE, A_MN, B_MN, TF32>;
  static SmemAttrCache attr;
  if (ensure_dyn_smem(kern, C::SMEM, attr) != cudaSuccess) return check_launch("gemm_tc_persistent attr");
  const int n_tiles = (q.N + BN - 1) / BN;
  const int total = ((q.M + kBM - 1) / kBM) * n_tiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static int sm_count[64] = {};
  if (!sm_count[dev & 63]) cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  if (sm_count[dev & 63] > 0) sms = sm_count[dev & 63];
  kern<<<std::min(total, sms), C::NT, C::SMEM, stream>>>(q.tm, q.Y, q.ldd, q.M, q.N, q.K, n_tiles, total, q.out_bf16, q.epi);
  return check_launch("gemm_tc_persistent");
}

template <bool A_MN, bool B_MN, bool TF32>
int dispatch(const Problem& q, int bn, cudaStream_t stream) {
  static const int persist = [] { const char* e = getenv("SIM_GEMM_BF16_PERSIST"); return e ? atoi(e) : 1; }();
  if (q.splits == 1 && (persist || q.epi.gbias || q.epi.gmax)) {  // the per-patch epilogue lives in the persistent kernel
    switch (bn) {
      case 64: return launch_persistent<64, 6, A_MN, B_MN, TF32>(q, stream);
      case 128: return launch_persistent<128, 5, A_MN, B_MN, TF32>(q, stream);
      case 192: return launch_persistent<192, 4, A_MN, B_MN, TF32>(q, stream);
      default: return launch_persistent<256, 3, A_MN, B_MN, TF32>(q, stream);
    }
  }
  switch (bn) {
    case 64: return launch<64, 6, A_MN, B_MN, TF32>(q, stream);
    case 128: return launch<128, 5, A_MN, B_MN, TF32>(q, stream);
    default: return launch<256, 4, A_MN, B_MN, TF32>(q, stream);
  }
}

template <bool TF32>
int gemm_any(const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* Y, long ldd, int out_bf16, int M,
             int N, int K, int splits, const float* bias, int relu, cudaStream_t stream, const float* gbias = nullptr,
             long ld_gbias = 0, float* gmax = nullptr, long ld_gmax = 0, int silu_col0 = -1) {
  constexpr int ES = El<TF32>::ES, BK = El<TF32>::BK;
  const bool grouped = gbias != nullptr || gmax != nullptr;
  SIM_REQUIRE(A && B && (Y || gmax) && M > 0 && N > 0 && K > 0, SIM_ERR_INVALID, "gemm_tc: empty problem / null tensor");
  SIM_REQUIRE(!grouped || (M % kGroupRows == 0 && splits == 1 && !out_bf16 && (!gbias || (aligned16(gbias) && ld_gbias % 4 == 0)) &&
                           (!gmax || ld_gmax >= N)),
              SIM_ERR_INVALID, "gemm_tc: the per-patch epilogue needs M %% 32 == 0, an fp32 result, no split-K, 16-byte aligned patch bias rows");
  SIM_REQUIRE(!TF32 || (!a_mn && !b_mn), SIM_ERR_INVALID, "gemm_tf32: only K-major operands ((M,K) and (N,K) row-major) are built");
  SIM_REQUIRE(aligned16(A) && aligned16(B) && (lda * ES) % 16 == 0 && (ldb * ES) % 16 == 0, SIM_ERR_ALIGN,
              "gemm_tc: TMA needs 16-byte aligned operand bases and row strides (lda=%ld ldb=%ld)", lda, ldb);
  SIM_REQUIRE(N % 4 == 0 && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & (out_bf16 ? 7u : 15u)) == 0 &&
                  (!bias || aligned16(bias)),
              SIM_ERR_ALIGN, "gemm_tc: the epilogue stores 4 columns at a time (N=%d ldd=%ld)", N, ldd);
  const int m_tiles = (M + kBM - 1) / kBM;
  if (splits <= 0) {
    // automatic split-K (weight gradients: few output tiles, a contraction over every token): at most two full waves of
    // CTAs, at least four k-blocks per CTA.  The caller zeroes the fp32 output.
    const long tiles = (long)m_tiles * ((N + 127) / 128);
    const int nk_all = (K + BK - 1) / BK;
    splits = (int)std::max<long>(1, std::min<long>(296 / tiles, nk_all / 4));
  }
  SIM_REQUIRE(splits == 1 || (!out_bf16 && !bias && !relu), SIM_ERR_INVALID,
              "gemm_tc: split-K accumulates into an fp32 output and has no bias / ReLU epilogue");
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
  Problem q;
  q.Y = Y, q.ldd = ldd, q.M = M, q.N = N, q.K = K, q.out_bf16 = out_bf16, q.splits = splits;
  q.epi.bias = bias, q.epi.relu = relu;
  q.epi.gbias = gbias, q.epi.ld_gbias = ld_gbias, q.epi.gmax = gmax, q.epi.ld_gmax = ld_gmax;
  SIM_REQUIRE(silu_col0 < 0 || (silu_col0 % 4 == 0 && splits == 1 && !grouped), SIM_ERR_INVALID,
              "gemm_tc: the silu epilogue starts at a multiple of 4 columns and excludes split-K / the per-patch epilogue");
  q.epi.silu_col0 = silu_col0;
  int rc;
  if (a_mn) rc = make_tmap_2d(&q.tm.a, A, M, K, lda, BK, TF32); else rc = make_tmap_2d(&q.tm.a, A, K, M, lda, kBM, TF32);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_2d(&q.tm.b, B, N, K, ldb, BK, TF32); else rc = make_tmap_2d(&q.tm.b, B, K, N, ldb, bn, TF32);
  if (rc) return rc;
  if constexpr (TF32) {
    // 32-bit MN-major operands need the SWIZZLE_128B_BASE32B atom (a different tile layout and TMA swizzle mode); the TF32
    // uses of this kernel (inference convolutions of the Encoder) are K-major on both sides, so only that form is built
    return dispatch<false, false, true>(q, bn, stream);
  } else {
    if (a_mn) return b_mn ? dispatch<true, true, false>(q, bn, stream) : dispatch<true, false, false>(q, bn, stream);
    return b_mn ? dispatch<false, true, false>(q, bn, stream) : dispatch<false, false, false>(q, bn, stream);
  }
}

}  // namespace

// Y[M,N] = op(A) . op(B)^T (+ bias[N], ReLU) on the tensor cores, fp32 accumulation.
//   a_mn = 0: A is (M, K) row-major with row stride lda (K-major);  a_mn = 1: A is (K, M) row-major (MN-major)
//   b_mn = 0: B is (N, K) row-major with row stride ldb (K-major);  b_mn = 1: B is (K, N) row-major (MN-major)
//   out_bf16: Y bf16 (row stride ldd), else fp32.  splits > 1: split-K, partial tiles are ADDED to a zeroed fp32 Y.
// gemm_bf16: bf16 operands (kind::f16).  gemm_tf32: fp32 operands consumed as TF32 (kind::tf32 reads the upper 19 bits) -
// the precision the reference's Conv1d layers run at under torch's default cudnn.allow_tf32 = True.
int gemm_bf16(const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* Y, long ldd, int out_bf16, int M,
              int N, int K, int splits, cudaStream_t stream) {
  return gemm_any<false>(A, lda, a_mn, B, ldb, b_mn, Y, ldd, out_bf16, M, N, K, splits, nullptr, 0, stream);
}

// gemm_bf16 (K-major operands) whose columns >= silu_col0 leave as silu(v): in_proj of the bf16 inference mixer
int gemm_bf16_silu(const void* A, long lda, const void* B, long ldb, void* Y, long ldd, int out_bf16, int M, int N, int K,
                   int silu_col0, cudaStream_t stream) {
  return gemm_any<false>(A, lda, 0, B, ldb, 0, Y, ldd, out_bf16, M, N, K, 1, nullptr, 0, stream, nullptr, 0, nullptr, 0, silu_col0);
}

int gemm_tf32(const float* A, long lda, int a_mn, const float* B, long ldb, int b_mn, void* Y, long ldd, int out_bf16, int M,
              int N, int K, int splits, const float* bias, int relu, cudaStream_t stream) {
  return gemm_any<true>(A, lda, a_mn, B, ldb, b_mn, Y, ldd, out_bf16, M, N, K, splits, bias, relu, stream);
}

// Y (fp32, may be NULL) = relu?(A . B^T + bias[N] + gbias[row / 32][N]);  gmax[row / 32][N] = max over each 32-row patch
int gemm_tf32_group(const float* A, long lda, const float* B, long ldb, float* Y, long ldd, int M, int N, int K, const float* bias,
                    const float* gbias, long ld_gbias, int relu, float* gmax, long ld_gmax, cudaStream_t stream) {
  SIM_REQUIRE(gbias || gmax, SIM_ERR_INVALID, "gemm_tf32_group: neither a patch bias nor a patch max was asked for");
  return gemm_any<true>(A, lda, 0, B, ldb, 0, Y, Y ? ldd : 4, 0, M, N, K, 1, bias, relu, stream, gbias, ld_gbias, gmax, ld_gmax);
}

}  // namespace sim
