// Shared device/host helpers for the si-mamba B200 (sm_100a) kernels.
// No torch types anywhere below the C-ABI (include/simamba.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace sim {

// ----------------------------------------------------------------------------- errors
enum : int {
  SIM_OK = 0,
  SIM_ERR_INVALID = -1,   // bad argument / unsupported shape
  SIM_ERR_ALIGN = -2,     // pointer or stride not 16-byte aligned where TMA needs it
  SIM_ERR_CUDA = -3,      // a CUDA runtime call failed
  SIM_ERR_WORKSPACE = -4  // workspace missing or too small
};

void set_error(const char* fmt, ...);  // thread-local message, see abi.cu
int check_launch(const char* what);    // cudaPeekAtLastError -> SIM_ERR_CUDA

#define SIM_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::sim::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

// Per-device, grow-only cache of cudaFuncAttributeMaxDynamicSharedMemorySize.  Function attributes are per device and
// nn.DataParallel drives several devices from one process, so a process-wide flag is not enough.  The
// first (warm-up) call on a device sets the attribute, outside any stream capture; later calls are free.
struct SmemAttrCache {
  size_t set[64] = {};
};
template <typename K>
static inline cudaError_t ensure_dyn_smem(K kern, size_t bytes, SmemAttrCache& c) {
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& cur = c.set[dev & 63];
  if (bytes <= cur) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ----------------------------------------------------------------------------- PTX: mbarrier + bulk async copy (TMA unit)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global bulk copy (bulk-group completion)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------- math
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// softplus with PyTorch's threshold (x > 20 -> x).  For small e = exp(x) the
// lg2(1+e) form loses relative accuracy, so use the alternating series there.
__device__ __forceinline__ float softplus_f(float x) {
  float e = ex2_approx(x * kLog2e);
  float series = e * (1.f - e * (0.5f - e * (0.33333334f - e * (0.25f - e * 0.2f))));
  float big = lg2_approx(1.f + e) * kLn2;
  float r = (e < 0.0625f) ? series : big;
  return (x > 20.f) ? x : r;
}

__device__ __forceinline__ float silu_f(float z) {
  float e = ex2_approx(-z * kLog2e);
  return z * rcp_approx(1.f + e);
}
__device__ __forceinline__ float sigmoid_f(float z) {
  float e = ex2_approx(-z * kLog2e);
  return rcp_approx(1.f + e);
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// x = p0 + p1 + p2 as three bf16 values with exactly formed fp32 residuals (operand format of gemm_split3.cu);
// four adjacent elements -> one 8-byte store per plane.
__device__ __forceinline__ void split3_store4(__nv_bfloat16* dst, long plane, float4 v) {
  const float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 q[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q[0][j] = __float2bfloat16_rn(f[j]);
    const float r1 = f[j] - __bfloat162float(q[0][j]);
    q[1][j] = __float2bfloat16_rn(r1);
    q[2][j] = __float2bfloat16_rn(r1 - __bfloat162float(q[1][j]));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) *reinterpret_cast<uint2*>(dst + k * plane) = *reinterpret_cast<const uint2*>(q[k]);
}

// x = x0 + 2^-11 x1' as two fp16 values: x0 = fp16(x) (11 significant bits), x1' = fp16(2^11 (x - x0)) - the residual is
// exact in fp32, the scaling keeps it out of the fp16 subnormal range for |x| down to 2^-14.  Operand format of the NP = 2
// kernels of gemm_split3.cu; valid for |x| < 65504.  Four adjacent elements -> one 8-byte store per plane.
__device__ __forceinline__ void split2h_store4(__half* dst, long plane, float4 v) {
  const float f[4] = {v.x, v.y, v.z, v.w};
  __half q[2][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q[0][j] = __float2half_rn(f[j]);
    q[1][j] = __float2half_rn((f[j] - __half2float(q[0][j])) * 2048.f);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) *reinterpret_cast<uint2*>(dst + k * plane) = *reinterpret_cast<const uint2*>(q[k]);
}

}  // namespace sim
