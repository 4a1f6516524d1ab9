// Farthest-point sampling and kNN patch grouping.
//
// Replaces pytorch3d sample_farthest_points / knn_points as called by
// Group.forward (models/point_mamba.py:93-110; seg twin pt_mamba.py:175-191).
// Contract (oracle/tokenizer.py): squared distance ((dx*dx)+(dy*dy))+(dz*dz)
// with separately rounded fp32 products and sums (no FMA contraction), ties
// towards the lower point index; FPS starts at index 0; the kNN result is the
// SET of the M smallest (distance, index) pairs, emitted in ascending index order.
//
// Both kernels are latency / SM-issue bound (HBM traffic is a few hundred KB):
// FPS is a chain of G serial arg-max steps, one CTA per cloud, points and
// running minima in registers, REDUX-based warp arg-max on the float bit
// pattern and one __syncthreads per step; kNN runs one warp per centre with
// the cloud staged in shared memory and selects the M-th smallest distance by
// a 31-step MSB-first radix descent on the bit pattern.

#include "kernels.cuh"

namespace sim {

// fma = 0 (default contract): every product and sum rounded on its own.  fma = 1: the contraction a CUDA compiler makes
// of `dist2 = 0; for d: dist2 += diff * diff` (pytorch3d's sample_farthest_points / knn_points device loops built with
// the default -fmad=true): fma(dz, dz, fma(dy, dy, dx * dx)).  Which one the reference's wheel used cannot be checked
// here (the wheel is absent), so both conventions are built and tested against their own oracle (include/simamba.h).
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz, int fma = 0) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  if (fma) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ----------------------------------------------------------------------------- FPS
// PN2 = false: pytorch3d semantics (Group.forward).  PN2 = true: pointnet2_ops furthest_point_sample as the runners
// call it right before the model (utils/misc.py:14-21, tools/runner_finetune.py:177-194): running
// minima start at 1e10, points with |p|^2 <= 1e-3 are never visited, the distance is the FMA-contracted
// dx*dx + dy*dy + dz*dz of the upstream .cu, and ties go to the candidate of the lowest upstream thread
// (index mod bs_up, bs_up = upstream block size) and then the lowest index - `key` below orders exactly that way.
template <int NT, int PPT, bool PN2>
__global__ void __launch_bounds__(NT) fps_kernel(const float* __restrict__ xyz, int N, int G, int* __restrict__ idx,
                                                 float* __restrict__ center, int bs_up, int fma) {
  extern __shared__ float s_xyz[];  // N*3
  constexpr int NW = NT / 32;
  __shared__ unsigned s_bits[2][NW];
  __shared__ unsigned s_idx[2][NW];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* src = xyz + (long)b * N * 3;
  for (int i = tid; i < N * 3; i += NT) s_xyz[i] = src[i];
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = tid + j * NT;
    if (i < N) {
      px[j] = s_xyz[3 * i], py[j] = s_xyz[3 * i + 1], pz[j] = s_xyz[3 * i + 2];
      md[j] = PN2 ? 1e10f : __int_as_float(0x7f800000);  // +inf
    } else {
      px[j] = py[j] = pz[j] = 0.f;
      md[j] = 0.f;  // padding can never beat a real point (index tie-break)
    }
  }
  unsigned last = 0;
  for (int g = 0; g < G; ++g) {
    if (tid == 0) {
      idx[(long)b * G + g] = (int)last;
      center[((long)b * G + g) * 3 + 0] = s_xyz[3 * last];
      center[((long)b * G + g) * 3 + 1] = s_xyz[3 * last + 1];
      center[((long)b * G + g) * 3 + 2] = s_xyz[3 * last + 2];
    }
    if (g == G - 1) break;
    const float wx = s_xyz[3 * last], wy = s_xyz[3 * last + 1], wz = s_xyz[3 * last + 2];
    unsigned best_bits = 0, best_idx = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int i = tid + j * NT;
      if constexpr (PN2) {
        const float mag = fmaf(pz[j], pz[j], fmaf(py[j], py[j], px[j] * px[j]));
        if (i >= N || mag <= 1e-3f) continue;
        const float dx = px[j] - wx, dy = py[j] - wy, dz = pz[j] - wz;
        md[j] = fminf(md[j], fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        const unsigned bits = __float_as_uint(md[j]);
        const unsigned key = (unsigned)(i % bs_up) * 64u + (unsigned)(i / bs_up);
        if (best_idx == 0xffffffffu || bits > best_bits || (bits == best_bits && key < best_idx)) {
          best_bits = bits;
          best_idx = key;
        }
      } else {
        const float d = sqdist3(px[j], py[j], pz[j], wx, wy, wz, fma);
        md[j] = fminf(md[j], d);
        const unsigned bits = __float_as_uint(md[j]);
        if (i < N && (best_idx == 0xffffffffu || bits > best_bits)) {  // strict >: keeps the lowest own index
          best_bits = bits;
          best_idx = (unsigned)i;
        }
      }
    }
    // warp arg-max on (bits desc, idx asc)
    unsigned m = __reduce_max_sync(0xffffffffu, best_bits);
    unsigned cand = (best_bits == m) ? best_idx : 0xffffffffu;
    unsigned mi = __reduce_min_sync(0xffffffffu, cand);
    const int buf = g & 1;
    if (lane == 0) {
      s_bits[buf][warp] = m;
      s_idx[buf][warp] = mi;
    }
    __syncthreads();
    unsigned wb = lane < NW ? s_bits[buf][lane] : 0u;
    unsigned wi = lane < NW ? s_idx[buf][lane] : 0xffffffffu;
    m = __reduce_max_sync(0xffffffffu, wb);
    cand = (wb == m) ? wi : 0xffffffffu;
    last = __reduce_min_sync(0xffffffffu, cand);
    if constexpr (PN2) last = last == 0xffffffffu ? 0u : (last % 64u) * (unsigned)bs_up + last / 64u;  // key -> index
  }
}

int fps(const float* xyz, int B, int N, int G, int* idx, float* center, cudaStream_t stream, int pointnet2, int fma) {
  SIM_REQUIRE(B > 0 && N > 0 && G > 0 && G <= N, SIM_ERR_INVALID, "fps: need 0 < G <= N (G=%d N=%d)", G, N);
  SIM_REQUIRE(xyz && idx && center, SIM_ERR_INVALID, "fps: null tensor");
  const size_t smem = (size_t)N * 3 * sizeof(float);
  int bs_up = 1;  // pointnet2_ops opt_n_threads: largest power of two <= N, at most 512
  while (bs_up * 2 <= N && bs_up < 512) bs_up *= 2;
#define SIM_FPS_LAUNCH(NT, PPT)                                                                         \
  do {                                                                                                  \
    if (pointnet2) {                                                                                    \
      auto kern = fps_kernel<NT, PPT, true>;                                                            \
      static SmemAttrCache attr; /* per device, grow-only, set by the first (warm-up) call */             \
      if (smem + 2048 > 48 * 1024) ensure_dyn_smem(kern, smem, attr);                                   \
      kern<<<B, NT, smem, stream>>>(xyz, N, G, idx, center, bs_up, fma);                                \
    } else {                                                                                            \
      auto kern = fps_kernel<NT, PPT, false>;                                                           \
      static SmemAttrCache attr;                                                                        \
      if (smem + 2048 > 48 * 1024) ensure_dyn_smem(kern, smem, attr);                                   \
      kern<<<B, NT, smem, stream>>>(xyz, N, G, idx, center, bs_up, fma);                                \
    }                                                                                                   \
  } while (0)
  if (N <= 512)
    SIM_FPS_LAUNCH(128, 4);
  else if (N <= 1024)
    SIM_FPS_LAUNCH(256, 4);
  else if (N <= 2048)
    SIM_FPS_LAUNCH(256, 8);
  else if (N <= 4096)
    SIM_FPS_LAUNCH(512, 8);
  else if (N <= 16384)
    SIM_FPS_LAUNCH(1024, 16);
  else {
    set_error("fps: N=%d exceeds the built maximum of 16384", N);
    return SIM_ERR_INVALID;
  }
#undef SIM_FPS_LAUNCH
  return check_launch("fps");
}

// ----------------------------------------------------------------------------- kNN group
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) knn_group_kernel(const float* __restrict__ xyz,
                                                               const float* __restrict__ center, int N, int G,
                                                               int M, int* __restrict__ idx_out,
                                                               float* __restrict__ nbr, float* __restrict__ nbr_org,
                                                               int fma) {
  extern __shared__ float sm[];
  float* s_xyz = sm;                                              // N*3
  unsigned* s_d = reinterpret_cast<unsigned*>(sm + (size_t)N * 3);  // WARPS * N distance bit patterns
  int* s_sel = reinterpret_cast<int*>(s_d + (size_t)WARPS * N);   // WARPS * M
  const int groups_per_cloud = (G + WARPS - 1) / WARPS;
  const int b = blockIdx.x / groups_per_cloud;
  const int g = (blockIdx.x % groups_per_cloud) * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = xyz + (long)b * N * 3;
  for (int i = threadIdx.x; i < N * 3; i += WARPS * 32) s_xyz[i] = src[i];
  __syncthreads();
  if (g >= G) return;
  const float cx = center[((long)b * G + g) * 3], cy = center[((long)b * G + g) * 3 + 1],
              cz = center[((long)b * G + g) * 3 + 2];
  unsigned* d = s_d + (size_t)warp * N;
  for (int i = lane; i < N; i += 32)
    d[i] = __float_as_uint(sqdist3(cx, cy, cz, s_xyz[3 * i], s_xyz[3 * i + 1], s_xyz[3 * i + 2], fma));
  __syncwarp();
  // T = M-th smallest bit pattern: the largest v with #{d < v} < M (distances are >= 0, so bit 31 is clear)
  unsigned T = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const unsigned cand = T | (1u << bit);
    unsigned cnt = 0;
    for (int i = lane; i < N; i += 32) cnt += d[i] < cand;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((int)cnt < M) T = cand;
  }
  unsigned cnt_lt = 0;
  for (int i = lane; i < N; i += 32) cnt_lt += d[i] < T;
  cnt_lt = __reduce_add_sync(0xffffffffu, cnt_lt);
  const int need_eq = M - (int)cnt_lt;  // ties at T taken in ascending index order
  int* sel = s_sel + warp * M;
  int out_base = 0, eq_seen = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int i0 = 0; i0 < N; i0 += 32) {
    const int i = i0 + lane;
    const unsigned v = i < N ? d[i] : 0xffffffffu;
    const bool is_eq = v == T;
    const unsigned bal_eq = __ballot_sync(0xffffffffu, is_eq);
    const bool take = (v < T) || (is_eq && (eq_seen + __popc(bal_eq & lt_mask)) < need_eq);
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (take) sel[out_base + __popc(bal & lt_mask)] = i;
    out_base += __popc(bal);
    eq_seen += __popc(bal_eq);
  }
  __syncwarp();
  const long row = (long)b * G + g;
  for (int e = lane; e < M; e += 32) idx_out[row * M + e] = sel[e];
  const float cc[3] = {cx, cy, cz};
  for (int e = lane; e < M * 3; e += 32) {
    const int comp = e % 3;
    const float v = s_xyz[3 * sel[e / 3] + comp];
    if (nbr_org) nbr_org[row * M * 3 + e] = v;
    if (nbr) nbr[row * M * 3 + e] = __fsub_rn(v, cc[comp]);
  }
}

int knn_group(const float* xyz, const float* center, int B, int N, int G, int M, int* idx, float* nbr,
              float* nbr_org, cudaStream_t stream, int fma) {
  SIM_REQUIRE(B > 0 && N > 0 && G > 0 && M > 0 && M <= N, SIM_ERR_INVALID, "knn_group: need 0 < M <= N");
  SIM_REQUIRE(xyz && center && idx, SIM_ERR_INVALID, "knn_group: null tensor");
  constexpr int WARPS = 8;
  const size_t smem = ((size_t)N * 3 + (size_t)WARPS * N + (size_t)WARPS * M) * 4;
  SIM_REQUIRE(smem <= 227 * 1024, SIM_ERR_INVALID, "knn_group: N=%d does not fit the shared-memory staging", N);
  auto kern = knn_group_kernel<WARPS>;
  static SmemAttrCache attr;  // per device, grow-only, set by the first (warm-up) call
  if (smem > 48 * 1024) ensure_dyn_smem(kern, smem, attr);
  const int grid = B * ((G + WARPS - 1) / WARPS);
  kern<<<grid, WARPS * 32, smem, stream>>>(xyz, center, N, G, M, idx, nbr, nbr_org, fma);
  return check_launch("knn_group");
}

}  // namespace sim
