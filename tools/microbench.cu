// Pipe-throughput microbenchmark for B200 (sm_100a): how many FFMA / FFMA2 / MUFU.EX2 / mixed
// instructions per clock per SM?  The selective scan is balanced between HBM, MUFU and FMA issue,
// so these numbers decide its inner-loop design ("scan cost model").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096
#define UNR 8

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  float a[UNR], b[UNR];
  float2 p[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) {
    a[i] = seed + i + threadIdx.x;
    b[i] = seed * 0.5f + i;
    p[i] = make_float2(a[i], b[i]);
  }
  const float c1 = seed * 1.0001f, c2 = seed * 0.37f;
  const float2 q1 = make_float2(c1, c2), q2 = make_float2(c2, c1);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      if (MODE == 0) {  // FFMA
        a[i] = fmaf(a[i], c1, c2);
        b[i] = fmaf(b[i], c2, c1);
      } else if (MODE == 1) {  // FFMA2
        p[i] = __ffma2_rn(p[i], q1, q2);
      } else if (MODE == 2) {  // MUFU.EX2
        a[i] = ex2(a[i]);
        b[i] = ex2(b[i]);
      } else if (MODE == 3) {  // scan-like mix, scalar: FMUL, EX2, FMUL, FFMA, FFMA per update
        float x = a[i] * c1;
        float e = ex2(x);
        float bu = c2 * b[i];
        a[i] = fmaf(e, a[i], bu);
        b[i] = fmaf(a[i], c1, b[i]);
      } else if (MODE == 4) {  // scan-like mix, packed: FMUL2, 2xEX2, FMUL2, FFMA2, FFMA2 per two updates
        float2 x = __fmul2_rn(p[i], q1);
        float2 e = make_float2(ex2(x.x), ex2(x.y));
        float2 bu = __fmul2_rn(q2, p[i]);
        p[i] = __ffma2_rn(e, p[i], bu);
        p[i] = __ffma2_rn(p[i], q1, q2);
      } else if (MODE == 5) {  // FMUL2 only
        p[i] = __fmul2_rn(p[i], q1);
      } else if (MODE == 6) {  // FFMA + EX2 1:1 (do the pipes overlap?)
        a[i] = fmaf(a[i], c1, c2);
        b[i] = ex2(b[i]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < UNR; ++i) s += a[i] + b[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread, float* out) {
  const int grid = 148 * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<grid, block>>>(out, 0.001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, block>>>(out, 0.001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double total = (double)grid * block * ITERS * UNR * ops_per_iter_per_thread;
  printf("%-28s %8.3f ms  %9.1f Gop/s  (%6.1f ops/clk/SM at max clock %d MHz)\n", name, ms, total / ms * 1e-6,
         total / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
}

// ---- shared-memory broadcast patterns: how many LSU wavefronts does a warp-wide LDS cost when lanes share addresses?
template <int MODE>
__global__ void __launch_bounds__(256) ks(float* out, int stride_sel) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int base;
  if (MODE == 0) base = 0;                       // LDS.128, all lanes same address
  else if (MODE == 1) base = (lane & 3) * 4;     // LDS.128, 4 distinct addresses (8 lanes each, interleaved)
  else if (MODE == 2) base = lane * 4;           // LDS.128, all distinct (512 B)
  else if (MODE == 3) base = 0;                  // LDS.32, all lanes same word
  else if (MODE == 4) base = lane;               // LDS.32, all distinct
  else if (MODE == 5) base = (lane >> 3) * 4;    // LDS.128, 4 distinct addresses, one per quarter-warp
  else base = (lane & 1) * 4;                    // LDS.128, 2 distinct addresses interleaved
  base += stride_sel;  // 0 at run time; keeps the address opaque
  float acc = 0.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      const int off = (base + i * 128 + (it & 7) * 16) & 4095;
      if (MODE == 3 || MODE == 4) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + off)));
        acc += v;
      } else {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x), "=f"(y), "=f"(z), "=f"(w)
                     : "r"((unsigned)__cvta_generic_to_shared(sm + (off & ~3))));
        acc += x + w;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) kshfl(float* out, float seed) {
  float a[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) a[i] = seed + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < UNR; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MUFU.EX2 next to shared-memory loads / shuffles: do the XU and LSU pipes overlap, or serialise in the MIO queue?
// per inner step: 2 MUFU.EX2 + NL broadcast LDS.128 + NS SHFL
template <int NL, int NSH>
__global__ void __launch_bounds__(256) kmix(float* out, float seed, int zero) {
  __shared__ __align__(16) float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = i * 1e-3f;
  __syncthreads();
  float a[UNR], b[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) a[i] = seed + i, b[i] = seed - i;
  float acc = 0.f;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x & 3) * 16 + zero);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      a[i] = ex2(a[i]);
      b[i] = ex2(b[i]);
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x), "=f"(y), "=f"(z), "=f"(w)
                     : "r"(base + ((i * NL + l) * 64 + (it & 3) * 1024) % 8192));
        acc += x + w;
      }
#pragma unroll
      for (int l = 0; l < NSH; ++l) acc += __shfl_xor_sync(0xffffffffu, acc, 1 + l);
    }
  }
#pragma unroll
  for (int i = 0; i < UNR; ++i) acc += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
void run_generic(const char* name, F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  // warp-instructions per SM: grid*block/32*ITERS*UNR / 148
  const double winst = (double)148 * 8 * 256 / 32 * ITERS * UNR / 148.0;
  printf("%-44s %8.3f ms  %6.2f cycles per warp-instruction per SM (at %d MHz)\n", name, ms,
         ms * 1e-3 * clk_khz * 1e3 / winst, clk_khz / 1000);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0>("FFMA (scalar)", 2, out);
  run<1>("FFMA2 (packed, FMAs)", 2, out);
  run<5>("FMUL2 (packed, muls)", 2, out);
  run<2>("MUFU.EX2", 2, out);
  run<6>("FFMA+EX2 1:1 (pairs)", 1, out);
  run<3>("scan mix scalar (updates)", 1, out);
  run<4>("scan mix packed (updates)", 2, out);
  // occupancy sensitivity of the scan-like mix: k CTAs of 128 threads per SM = k warps per scheduler
  for (int kk : {1, 2, 3, 4, 6, 8, 16}) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<4><<<148 * kk, 128>>>(out, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<4><<<148 * kk, 128>>>(out, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double upd = (double)148 * kk * 128 * ITERS * UNR * 2;
    printf("scan mix packed, %2d warps/scheduler: %7.3f ms  %5.2f updates/clk/SM\n", kk, ms,
           upd / (ms * 1e-3) / 148.0 / 1.965e9);
  }
  {
    auto mix = [&](const char* name, auto launch) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      launch();
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double mufu = (double)148 * 8 * 256 * ITERS * UNR * 2;
      printf("%-40s %7.3f ms  %5.2f MUFU/clk/SM\n", name, ms, mufu / (ms * 1e-3) / 148.0 / 1.965e9);
    };
    mix("2 MUFU + 0 LDS.128 + 0 SHFL", [&] { kmix<0, 0><<<148 * 8, 256>>>(out, 0.5f, 0); });
    mix("2 MUFU + 1 LDS.128 (broadcast)", [&] { kmix<1, 0><<<148 * 8, 256>>>(out, 0.5f, 0); });
    mix("2 MUFU + 2 LDS.128 (broadcast)", [&] { kmix<2, 0><<<148 * 8, 256>>>(out, 0.5f, 0); });
    mix("2 MUFU + 4 LDS.128 (broadcast)", [&] { kmix<4, 0><<<148 * 8, 256>>>(out, 0.5f, 0); });
    mix("2 MUFU + 1 SHFL", [&] { kmix<0, 1><<<148 * 8, 256>>>(out, 0.5f, 0); });
    mix("2 MUFU + 1 LDS.128 + 1 SHFL", [&] { kmix<1, 1><<<148 * 8, 256>>>(out, 0.5f, 0); });
  }
  run_generic("LDS.128 all lanes same address", [&] { ks<0><<<148 * 8, 256>>>(out, 0); });
  run_generic("LDS.128 4 addresses interleaved (lane&3)", [&] { ks<1><<<148 * 8, 256>>>(out, 0); });
  run_generic("LDS.128 4 addresses, one per quarter-warp", [&] { ks<5><<<148 * 8, 256>>>(out, 0); });
  run_generic("LDS.128 2 addresses interleaved (lane&1)", [&] { ks<6><<<148 * 8, 256>>>(out, 0); });
  run_generic("LDS.128 all distinct", [&] { ks<2><<<148 * 8, 256>>>(out, 0); });
  run_generic("LDS.32 all lanes same word", [&] { ks<3><<<148 * 8, 256>>>(out, 0); });
  run_generic("LDS.32 all distinct", [&] { ks<4><<<148 * 8, 256>>>(out, 0); });
  run_generic("SHFL.BFLY", [&] { kshfl<<<148 * 8, 256>>>(out, 0.5f); });
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
