// MAE token layout on the device: index maps of the masked spectral sort and of
// the token restore, and the deterministic backward of both row movements.
//
// Reference: MaskMamba_3.forward (models/point_mamba.py:2734-2796) sorts tokens / mask / pos per spectral order, keeps
// the visible rows with boolean-mask indexing (a host sync per order) and concatenates k orders + their flip;
// Point_MAE_Mamba.forward (:3147-3197) rebuilds the full sequence with ten torch.where / index_put rounds.  Here one
// warp per cloud walks the 2kG decoder positions once (ballot prefix count of the visible ones) and writes every
// map the forward and backward row kernels need; the rows themselves move through gather_rows (rowops.cu) and the
// two kernels below.  No atomics on the data path except the mask-token gradient (one fp32 atomic per CTA and column).

#include "kernels.cuh"

namespace sim {

namespace {

template <typename T>
__device__ __forceinline__ float4 ldg4(const T* p);
template <>
__device__ __forceinline__ float4 ldg4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 ldg4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void stg4(T* p, float4 v);
template <>
__device__ __forceinline__ void stg4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void stg4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<const unsigned*>(&lo);
  r.y = *reinterpret_cast<const unsigned*>(&hi);
  *reinterpret_cast<uint2*>(p) = r;
}

// decoder position t of cloud b shows patch perm_full[t]: t < kG -> perm[b, t / G, t % G]; the second half mirrors it
__global__ void __launch_bounds__(256) mae_index_maps_kernel(const int* __restrict__ perm,
                                                             const unsigned char* __restrict__ mask, int B, int k, int G,
                                                             int n_vis, int* __restrict__ perm_full,
                                                             unsigned char* __restrict__ mask_full,
                                                             int* __restrict__ restore_src, int* __restrict__ src_vis,
                                                             int* __restrict__ vis_pos, int* __restrict__ rec_src,
                                                             int* __restrict__ inv_vis, int* __restrict__ err) {
  const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const int kG = k * G, T = 2 * kG, J = 2 * k;
  const int RV = J * n_vis, RM = T - RV;
  const int* pb = perm + (long)b * kG;
  const unsigned char* mb = mask + (long)b * G;
  int run = 0;  // visible positions before this chunk
  for (int c = 0; c < T; c += 32) {
    const int t = c + lane;
    int p = 0, vis = 0;
    if (t < T) {
      p = pb[t < kG ? t : T - 1 - t];
      vis = mb[p] == 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, vis);
    if (t < T) {
      const int rank = run + __popc(bal & ((1u << lane) - 1u));
      const int seg = t / G;  // every patch appears exactly once in each of the 2k segments
      perm_full[(long)b * T + t] = p;
      mask_full[(long)b * T + t] = (unsigned char)!vis;
      if (vis) {
        restore_src[(long)b * T + t] = rank;
        if (rank < RV) {
          src_vis[(long)b * RV + rank] = p;
          vis_pos[(long)b * RV + rank] = t;
        }
        inv_vis[((long)b * G + p) * J + seg] = rank < RV ? rank : -1;
      } else {
        restore_src[(long)b * T + t] = -1;
        const int mr = t - rank;
        if (mr < RM) rec_src[(long)b * RM + mr] = t;
        inv_vis[((long)b * G + p) * J + seg] = -1;
      }
    }
    run += __popc(bal);
  }
  if (lane == 0 && run != RV) atomicExch(err, b + 1);  // this cloud does not have n_vis visible patches
}

// out[b, r, :] = sum_j x[b, idx[b, r, j], :] over idx >= 0  (deterministic scatter-add written as a gather)
template <typename T>
__global__ void __launch_bounds__(256) gather_sum_rows_kernel(const T* __restrict__ x, const int* __restrict__ idx,
                                                              T* __restrict__ out, int B, int R_in, int R_out, int J,
                                                              int C) {
  const long w = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (w >= (long)B * R_out) return;
  const int lane = threadIdx.x & 31;
  const int b = w / R_out;
  const int* ip = idx + w * J;
  for (int q = lane; q < C / 4; q += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < J; ++j) {
      const int s = ip[j];
      if (s >= 0) {
        const float4 a = ldg4<T>(x + ((long)b * R_in + s) * C + 4 * q);
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      }
    }
    stg4<T>(out + w * C + 4 * q, acc);
  }
}

// inv[b, r, 0..J) = ascending list of the output rows t with src_idx[b, t] == r (-1 padded): the inverse of any row
// gather map, so that its backward runs as gather_sum_rows (deterministic, no atomics).  One CTA per cloud; thread r
// walks the map (staged in shared memory) for source row r.  A row read by more than J outputs raises err = b + 1.
__global__ void __launch_bounds__(256) invert_row_map_kernel(const int* __restrict__ src_idx, int R_in, int R_out, int J,
                                                             int* __restrict__ inv, int* __restrict__ err) {
  extern __shared__ int s_map[];
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < R_out; t += blockDim.x) s_map[t] = src_idx[(long)b * R_out + t];
  __syncthreads();
  for (int r = threadIdx.x; r < R_in; r += blockDim.x) {
    int* o = inv + ((long)b * R_in + r) * J;
    int n = 0;
    for (int t = 0; t < R_out; ++t) {
      if (s_map[t] == r) {
        if (n < J) o[n] = t;
        ++n;
      }
    }
    if (n > J) atomicExch(err, b + 1);
    for (; n < J; ++n) o[n] = -1;
  }
}

// dfill[c] += sum over rows with sel[row] < 0 of x[row, c]   (gradient of the mask token)
template <typename T>
__global__ void __launch_bounds__(256) masked_colsum_kernel(const T* __restrict__ x, const int* __restrict__ sel, long rows,
                                                            int C, int rows_per_cta, float* __restrict__ dfill) {
  const long r0 = (long)blockIdx.x * rows_per_cta;
  const long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (long r = r0; r < r1; ++r)
      if (sel[r] < 0) acc += to_f32<T>(x[r * C + c]);
    if (acc != 0.f) atomicAdd(dfill + c, acc);
  }
}

}  // namespace

int mae_index_maps(const int* perm, const unsigned char* mask, int B, int k, int G, int n_vis, int* perm_full,
                   unsigned char* mask_full, int* restore_src, int* src_vis, int* vis_pos, int* rec_src, int* inv_vis,
                   int* err_flag, cudaStream_t stream) {
  SIM_REQUIRE(B > 0 && k > 0 && G > 0 && n_vis >= 0 && n_vis <= G, SIM_ERR_INVALID, "mae_index_maps: bad sizes");
  SIM_REQUIRE(perm && mask && perm_full && mask_full && restore_src && src_vis && vis_pos && rec_src && inv_vis && err_flag,
              SIM_ERR_INVALID, "mae_index_maps: null tensor");
  mae_index_maps_kernel<<<(B + 7) / 8, 256, 0, stream>>>(perm, mask, B, k, G, n_vis, perm_full, mask_full, restore_src,
                                                        src_vis, vis_pos, rec_src, inv_vis, err_flag);
  return check_launch("mae_index_maps");
}

int gather_sum_rows(const void* x, const int* idx, void* out, int B, int R_in, int R_out, int J, int C, int dtype,
                    cudaStream_t stream) {
  SIM_REQUIRE(B > 0 && R_in > 0 && R_out > 0 && J > 0 && C > 0 && C % 4 == 0, SIM_ERR_INVALID,
              "gather_sum_rows: C must be a multiple of 4");
  SIM_REQUIRE(x && idx && out, SIM_ERR_INVALID, "gather_sum_rows: null tensor");
  SIM_REQUIRE(aligned16(x) && aligned16(out), SIM_ERR_ALIGN, "gather_sum_rows: tensors must be 16-byte aligned");
  const int grid = (int)(((long)B * R_out + 7) / 8);
  if (dtype == 0)
    gather_sum_rows_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), idx, static_cast<float*>(out), B,
                                                            R_in, R_out, J, C);
  else if (dtype == 1)
    gather_sum_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), idx,
                                                                    static_cast<__nv_bfloat16*>(out), B, R_in, R_out, J, C);
  else {
    set_error("gather_sum_rows: bad dtype %d", dtype);
    return SIM_ERR_INVALID;
  }
  return check_launch("gather_sum_rows");
}

int invert_row_map(const int* src_idx, int B, int R_in, int R_out, int J, int* inv, int* err_flag, cudaStream_t stream) {
  SIM_REQUIRE(src_idx && inv && err_flag, SIM_ERR_INVALID, "invert_row_map: null tensor");
  SIM_REQUIRE(B > 0 && R_in > 0 && R_out > 0 && J > 0 && R_out <= 12288, SIM_ERR_INVALID,
              "invert_row_map: bad sizes (R_out <= 12288)");
  invert_row_map_kernel<<<B, 256, R_out * sizeof(int), stream>>>(src_idx, R_in, R_out, J, inv, err_flag);
  return check_launch("invert_row_map");
}

int masked_colsum(const void* x, const int* sel, long rows, int C, float* dfill, int dtype, cudaStream_t stream) {
  SIM_REQUIRE(x && sel && dfill && rows > 0 && C > 0, SIM_ERR_INVALID, "masked_colsum: null tensor / empty problem");
  const int rows_per_cta = 64;
  const int grid = (int)((rows + rows_per_cta - 1) / rows_per_cta);
  if (dtype == 0)
    masked_colsum_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), sel, rows, C, rows_per_cta, dfill);
  else if (dtype == 1)
    masked_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), sel, rows, C,
                                                                  rows_per_cta, dfill);
  else {
    set_error("masked_colsum: bad dtype %d", dtype);
    return SIM_ERR_INVALID;
  }
  return check_launch("masked_colsum");
}

}  // namespace sim
