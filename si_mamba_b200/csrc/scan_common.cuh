// Pieces shared by the selective-scan forward kernels (selective_scan_fwd.cu, selective_scan_fwd_ws.cu).
#pragma once

#include "kernels.cuh"
#include "tma.cuh"

namespace sim {

constexpr int kNState = 16;

struct ScanTmaps {
  CUtensorMap u, delta, z, B, C, out;
};

// four consecutive elements of T from shared memory as fp32
template <typename T>
__device__ __forceinline__ float4 lds4(const T* p);
template <>
__device__ __forceinline__ float4 lds4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 lds4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  // bf16 -> fp32 is a 16-bit shift
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}

template <int N>
__device__ __forceinline__ void lds_vec(const float* p, float (&v)[N]) {
  if constexpr (N == 1) {
    v[0] = p[0];
  } else if constexpr (N == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x, v[1] = t.y;
  } else {
    static_assert(N % 4 == 0, "vector width");
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// S consecutive elements of T (a broadcast row slice of B or C) from shared memory as fp32
template <typename T, int N>
__device__ __forceinline__ void lds_row(const T* p, float (&v)[N]) {
  if constexpr (sizeof(T) == 4) {
    lds_vec<N>(reinterpret_cast<const float*>(p), v);
  } else {
    static_assert(N % 2 == 0, "pairs");
    if constexpr (N % 8 == 0) {
#pragma unroll
      for (int i = 0; i < N / 8; ++i) {
        const uint4 r = reinterpret_cast<const uint4*>(p)[i];
        const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
          v[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
      }
    } else if constexpr (N % 4 == 0) {
#pragma unroll
      for (int i = 0; i < N / 4; ++i) {
        const uint2 r = reinterpret_cast<const uint2*>(p)[i];
        v[4 * i] = __uint_as_float(r.x << 16), v[4 * i + 1] = __uint_as_float(r.x & 0xffff0000u);
        v[4 * i + 2] = __uint_as_float(r.y << 16), v[4 * i + 3] = __uint_as_float(r.y & 0xffff0000u);
      }
    } else {
      const unsigned r = *reinterpret_cast<const unsigned*>(p);
      v[0] = __uint_as_float(r << 16), v[1] = __uint_as_float(r & 0xffff0000u);
    }
  }
}

// implemented in selective_scan_fwd_ws.cu: warp-specialised kernel (variants 5000+)
int selective_scan_fwd_ws(const ScanParams& p, int dtype, int variant, cudaStream_t stream);

}  // namespace sim
