// TMA tensor-map helpers: host-side cuTensorMapEncodeTiled (resolved through the runtime's driver
// entry-point query, so the library does not link libcuda) and the device-side
// cp.async.bulk.tensor wrappers (SASS: UTMALDG / UTMASTG).
//
// Measured on B200 (profiles/r01_notes.md): the TMA unit services roughly one copy instruction per
// ~40-46 cycles per SM regardless of its size, so a tile must arrive as ONE tensor copy per operand;
// row-by-row 256-byte cp.async.bulk copies made the scan 5x slower than its MUFU bound.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace sim {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static inline PFN_tmapEncodeTiled tmap_encode_fn() {
  // cuTensorMapEncodeTiled is a driver call and needs a context bound to the CALLING thread (CUDA_ERROR_INVALID_CONTEXT
  // otherwise).  A thread that has only been told its device through cudaGetDevice - PyTorch's autograd worker when one of
  // these kernels is the first thing its backward runs - has none yet: cudaSetDevice binds the primary context (no stream
  // operation, so it is legal during graph capture).
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
    bound = true;
  }
  static PFN_tmapEncodeTiled fn = nullptr;  // benign race: every thread resolves the same pointer
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 3-D map over a token-major activation: dims (cols, L, batch), row stride ld elements, box (box_c, box_t, 1).
// Out-of-range rows (t >= L) read as zero and are clipped on store, so tail tiles need no special casing.
static inline int make_tmap_tokens(CUtensorMap* m, const void* base, int dtype, int cols, int L, int batch, long ld,
                                   int box_c, int box_t) {
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SIM_ERR_CUDA;
  }
  const cuuint64_t es = dtype == 0 ? 4 : 2;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * es, (cuuint64_t)ld * es * (cuuint64_t)L};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_t, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%d L=%d batch=%d ld=%ld box=%dx%d)", (int)r, cols,
              L, batch, ld, box_c, box_t);
    return SIM_ERR_CUDA;
  }
  return SIM_OK;
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// global -> shared, 3-D tile, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// global -> shared, 2-D tile, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// shared -> global, 3-D tile, bulk-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, int c0, int c1, int c2, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
               : "memory");
}

}  // namespace sim
