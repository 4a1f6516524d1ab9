// Patch-graph Laplacian + batched small symmetric eigensolver + spectral argsort,
// one CTA per cloud.
//
// Replaces, in models/point_mamba.py:
//   create_graph_from_feature_space_gpu_weighted_adjacency :664-715 / create_graph_from_centers :620-661
//   calc_top_k_eigenvalues_eigenvectors (B serial cuSOLVER eigh calls) :717-761, batched twin :3001-3050,
//   calc_top_k_eigenvalues_eigenvectors_symmetric :764-814
//   the torch.sort of sort_points_by_fiedler :817-826 (sign rule: work_order.py:360-365)
//
// Pipeline inside the CTA (everything stays in shared memory for G <= 128; larger G
// uses a caller-provided global workspace with the same code):
//   1. pairwise sqrt-distances of the G centres (fp32, fixed op order, no FMA)
//   2. per-row (k_nn+1)-nearest selection by (distance, index) rank counting, adjacency scatter
//   3. A = (A + A^T)/2, degrees, Laplacian in fp32 exactly as the reference writes it; the
//      operator eigh sees is the LOWER triangle mirrored (UPLO='L')
//   4. fp64 Householder tridiagonalisation (reflectors kept in the eliminated rows)
//   5. fp64 multisection on Sturm counts for the k wanted eigenvalues (all k concurrently)
//   6. fp64 inverse iteration on the tridiagonal (pivoted LU), one thread per eigenvector
//   7. back-transformation by the stored reflectors, one warp per eigenvector
//   8. normalise, sign-canonicalise, rank-count argsort -> perm / inverse perm
// fp64 end to end after step 3 is what makes the ordering reproducible against the
// fp64 LAPACK oracle; the kernel is latency / issue bound, HBM
// traffic is ~G*(3+4k)*4 bytes per cloud.

#include "kernels.cuh"

namespace sim {


__device__ __forceinline__ float sqdist3s(const float* a, const float* b) {
  const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; scratch has >= 33 doubles; two barriers, result broadcast to all threads
template <int NT>
__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) r += scratch[w];
  __syncthreads();
  return r;
}

// number of eigenvalues of the tridiagonal (d, e) that are < x : sign changes of the scaled
// characteristic-polynomial recurrence p_{i+1} = (d_i - x) p_i - e_{i-1}^2 p_{i-1}
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int G, double x) {
  double pm1 = 1.0, p = d[0] - x;
  int neg = p < 0.0;
  int cnt = neg;
  for (int i = 1; i < G; ++i) {
    double pn = (d[i] - x) * p - e2[i - 1] * pm1;
    pm1 = p;
    p = pn;
    const double a = fabs(p);
    if (a > 1e100) {
      p *= 1e-100;
      pm1 *= 1e-100;
    } else if (a < 1e-100) {
      p *= 1e100;
      pm1 *= 1e100;
    }
    const int s = (p == 0.0) ? !neg : (p < 0.0);
    cnt += (s != neg);
    neg = s;
  }
  return cnt;
}

template <int NT, bool IN_SMEM>
__global__ void __launch_bounds__(NT) spectral_kernel(const SpectralParams P) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int G = P.G, k = P.k;
  const int GP = (G + 31) & ~31;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;

  // ---- carve shared memory
  unsigned char* ptr = sm_raw;
  double* Mat;
  float* Adj;
  if (IN_SMEM) {
    Mat = reinterpret_cast<double*>(ptr);
    ptr += (size_t)G * G * sizeof(double);
    Adj = reinterpret_cast<float*>(ptr);
    ptr += (size_t)G * G * sizeof(float);
  } else {
    Mat = P.ws_mat + (size_t)b * G * G;
    Adj = P.ws_adj + (size_t)b * G * G;
  }
  ptr = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ptr) + 15) & ~(uintptr_t)15);
  float* dist = reinterpret_cast<float*>(Mat);  // aliases Mat until the Laplacian is formed
  double* s_d = reinterpret_cast<double*>(ptr);  ptr += GP * sizeof(double);
  double* s_e = reinterpret_cast<double*>(ptr);  ptr += GP * sizeof(double);
  double* s_e2 = reinterpret_cast<double*>(ptr); ptr += GP * sizeof(double);
  double* s_tau = reinterpret_cast<double*>(ptr); ptr += GP * sizeof(double);
  double* s_v = reinterpret_cast<double*>(ptr);  ptr += GP * sizeof(double);
  double* s_w = reinterpret_cast<double*>(ptr);  ptr += GP * sizeof(double);
  double* s_part = reinterpret_cast<double*>(ptr); ptr += (size_t)NT * sizeof(double);  // symv partials [jgroup][col], NT/GP groups
  double* s_Z = reinterpret_cast<double*>(ptr);  ptr += (size_t)k * GP * sizeof(double);
  double* s_lu;  // u0,u1,u2,l,y per vector
  if (P.lu_alias) {
    s_lu = reinterpret_cast<double*>(Adj);  // Adj is dead once the Laplacian is formed
  } else {
    s_lu = reinterpret_cast<double*>(ptr);
    ptr += (size_t)k * 5 * GP * sizeof(double);
  }
  double* s_scr = reinterpret_cast<double*>(ptr); ptr += 40 * sizeof(double);
  double* s_lo = reinterpret_cast<double*>(ptr); ptr += 8 * sizeof(double);
  double* s_hi = reinterpret_cast<double*>(ptr); ptr += 8 * sizeof(double);
  double* s_lam = reinterpret_cast<double*>(ptr); ptr += 8 * sizeof(double);
  int* s_cnt = reinterpret_cast<int*>(ptr);      ptr += NT * sizeof(int);
  float* s_c = reinterpret_cast<float*>(ptr);    ptr += GP * 3 * sizeof(float);
  float* s_deg = reinterpret_cast<float*>(ptr);  ptr += GP * sizeof(float);
  unsigned char* s_piv = ptr;                    ptr += (size_t)k * GP;

  if (P.adj_in) {
    // caller-supplied adjacency (calc_top_k_eigenvalues_eigenvectors(adj_matrices, k, smallest), :717)
    const float* ai = P.adj_in + (size_t)b * G * G;
    for (int e = tid; e < G * G; e += NT) Adj[e] = ai[e];
    __syncthreads();
  } else {
  // ================================================================ 1. distances
  const float* cen = P.center + (size_t)b * G * 3;
  for (int i = tid; i < G * 3; i += NT) s_c[i] = cen[i];
  __syncthreads();
  for (int e = tid; e < G * G; e += NT) {
    const int i = e / G, j = e % G;
    dist[e] = __fsqrt_rn(sqdist3s(s_c + 3 * i, s_c + 3 * j));
    Adj[e] = 0.f;
  }
  __syncthreads();

  // ================================================================ 2. kNN selection + adjacency scatter
  const int kk = P.k_nn + 1;
  // alpha == 0 branch of create_graph_from_centers (:628, :647): exp(-d^2 / (2 sigma^2)), sigma = batch-wide mean distance
  float two_sigma2 = 1.f;
  if (P.sigma) {
    const float sg = *P.sigma;
    two_sigma2 = __fmul_rn(2.f, __fmul_rn(sg, sg));
  }
  for (int i = warp; i < G; i += NW) {
    const float* di = dist + (size_t)i * G;
    for (int j = lane; j < G; j += 32) {
      const float dij = di[j];
      int rank = 0;
      for (int q = 0; q < G; ++q) {
        const float dq = di[q];
        rank += (dq < dij) || (dq == dij && q < j);
      }
      if (rank < kk && (P.self_loop || rank >= 1)) {
        const float wgt = P.binary ? 1.f
                          : P.sigma ? expf(__fdiv_rn(-__fmul_rn(dij, dij), two_sigma2))
                                    : expf(__fmul_rn(-P.alpha, __fmul_rn(dij, dij)));
        Adj[(size_t)i * G + j] = wgt;
        if (P.symmetric) Adj[(size_t)j * G + i] = wgt;
      }
    }
  }
  __syncthreads();
  }  // !adj_in
  if (P.adjacency) {
    float* ao = P.adjacency + (size_t)b * G * G;
    for (int e = tid; e < G * G; e += NT) ao[e] = Adj[e];
  }

  // ================================================================ 3. symmetrise, degree, Laplacian (fp32 as the reference)
  for (int i = warp; i < G; i += NW) {
    double s = 0.0;
    for (int j = lane; j < G; j += 32)
      s += (double)__fmul_rn(__fadd_rn(Adj[(size_t)i * G + j], Adj[(size_t)j * G + i]), 0.5f);
    s = warp_sum_d(s);
    if (lane == 0) s_deg[i] = (float)s;
  }
  __syncthreads();
  for (int e = tid; e < G * G; e += NT) {
    const int i = e / G, j = e % G;
    if (i < j) continue;
    const float a = __fmul_rn(__fadd_rn(Adj[(size_t)i * G + j], Adj[(size_t)j * G + i]), 0.5f);
    const float delta = (i == j) ? 1.f : 0.f;
    float l;
    if (P.matrix_sym) {
      const float di = (float)(1.0 / sqrt((double)s_deg[i])), dj = (float)(1.0 / sqrt((double)s_deg[j]));
      l = __fsub_rn(delta, __fmul_rn(__fmul_rn(di, a), dj));
    } else if (P.eps_clamp) {
      l = __fsub_rn(delta, __fdiv_rn(a, fmaxf(s_deg[i], 1e-12f)));
    } else {
      const float r = __fdiv_rn(1.0f, __fadd_rn(s_deg[i], 1e-6f));
      l = __fsub_rn(delta, __fmul_rn(r, a));
    }
    // NOTE: dist aliases Mat, but dist is dead after step 2 and every (i,j) slot is written here
    // only after all Adj reads of this thread; Mat writes never touch Adj.
    Mat[(size_t)i * G + j] = (double)l;
    Mat[(size_t)j * G + i] = (double)l;
  }
  __syncthreads();

  // ================================================================ 4. Householder tridiagonalisation (fp64)
  // At step kk the trailing square [kk:, kk:] is full symmetric.  Row kk (right of the diagonal) is the
  // vector to eliminate; its reflector v (v[0] = 1) overwrites that row for the back-transformation.
  for (int kc = 0; kc < G - 2; ++kc) {
    const int o = kc + 1, n = G - o;
    double* rowk = Mat + (size_t)kc * G + o;
    double part = 0.0;
    for (int i = 1 + tid; i < n; i += NT) part += rowk[i] * rowk[i];
    const double xnorm2 = block_sum_d<NT>(part, s_scr);
    const double alpha = rowk[0];
    if (xnorm2 == 0.0) {  // already tridiagonal here: H = I
      if (tid == 0) {
        s_d[kc] = Mat[(size_t)kc * G + kc];
        s_e[kc] = alpha;
        s_tau[kc] = 0.0;
      }
      __syncthreads();
      continue;
    }
    const double beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
    const double tau = (beta - alpha) / beta;
    const double scale = 1.0 / (alpha - beta);
    for (int i = tid; i < n; i += NT) {
      const double vi = (i == 0) ? 1.0 : rowk[i] * scale;
      s_v[i] = vi;
    }
    if (tid == 0) {
      s_d[kc] = Mat[(size_t)kc * G + kc];
      s_e[kc] = beta;
      s_tau[kc] = tau;
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) rowk[i] = s_v[i];  // keep the reflector
    // p = tau * A v  (column access through symmetry: lanes run along a row of A)
    {
      const int col = tid % GP, jg = tid / GP, njg = NT / GP;
      if (col < n) {
        // four independent partial sums: the loads of a column walk (L2-resident workspace for G > 128) overlap instead of
        // waiting on one another through a single fp64 dependency chain
        constexpr int U = IN_SMEM ? 4 : 8;  // loads in flight per thread
        double acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = 0.0;
        const double* Ac = Mat + (size_t)o * G + o + col;
        int j = jg;
        for (; j + (U - 1) * njg < n; j += U * njg) {
          double x[U];
#pragma unroll
          for (int u = 0; u < U; ++u) x[u] = Ac[(size_t)(j + u * njg) * G];
#pragma unroll
          for (int u = 0; u < U; ++u) acc[u] = fma(x[u], s_v[j + u * njg], acc[u]);
        }
        for (; j < n; j += njg) acc[0] = fma(Ac[(size_t)j * G], s_v[j], acc[0]);
        double tot = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) tot += acc[u];
        s_part[jg * GP + col] = tot;
      }
    }
    __syncthreads();
    double pv_part = 0.0;
    for (int i = tid; i < n; i += NT) {
      double acc = 0.0;
      for (int jg = 0; jg < NT / GP; ++jg) acc += s_part[jg * GP + i];
      acc *= tau;
      s_w[i] = acc;  // p for now
      pv_part += acc * s_v[i];
    }
    const double pv = block_sum_d<NT>(pv_part, s_scr);
    const double hh = 0.5 * tau * pv;
    for (int i = tid; i < n; i += NT) s_w[i] = s_w[i] - hh * s_v[i];
    __syncthreads();
    // rank-2 update of the full trailing square
    {
      const int col = tid % GP, ig = tid / GP, nig = NT / GP;
      if (col < n) {
        const double vc = s_v[col], wc = s_w[col];
        constexpr int U = IN_SMEM ? 4 : 8;  // rows in flight per thread
        int i = ig;
        for (; i + (U - 1) * nig < n; i += U * nig) {
          double* a = Mat + (size_t)(o + i) * G + o + col;
          const size_t st = (size_t)nig * G;
          double x[U];
#pragma unroll
          for (int u = 0; u < U; ++u) x[u] = a[u * st];
#pragma unroll
          for (int u = 0; u < U; ++u) a[u * st] = x[u] - (s_v[i + u * nig] * wc + s_w[i + u * nig] * vc);
        }
        for (; i < n; i += nig) {
          double* a = Mat + (size_t)(o + i) * G + o + col;
          *a = *a - (s_v[i] * wc + s_w[i] * vc);
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (G >= 2) {
      s_d[G - 2] = Mat[(size_t)(G - 2) * G + (G - 2)];
      s_e[G - 2] = Mat[(size_t)(G - 2) * G + (G - 1)];
    }
    s_d[G - 1] = Mat[(size_t)(G - 1) * G + (G - 1)];
    s_e[G - 1] = 0.0;
  }
  __syncthreads();
  for (int i = tid; i < G; i += NT) s_e2[i] = s_e[i] * s_e[i];

  // ================================================================ 5. multisection for the k wanted eigenvalues
  {
    double lo_p = 1e300, hi_p = -1e300, nrm_p = 0.0;
    for (int i = tid; i < G; i += NT) {
      const double r = (i > 0 ? fabs(s_e[i - 1]) : 0.0) + (i < G - 1 ? fabs(s_e[i]) : 0.0);
      lo_p = fmin(lo_p, s_d[i] - r);
      hi_p = fmax(hi_p, s_d[i] + r);
    }
    // block min / max through the scratch (NT <= 1024)
    for (int o2 = 16; o2 >= 1; o2 >>= 1) {
      lo_p = fmin(lo_p, __shfl_xor_sync(0xffffffffu, lo_p, o2));
      hi_p = fmax(hi_p, __shfl_xor_sync(0xffffffffu, hi_p, o2));
    }
    __syncthreads();
    if (lane == 0) {
      s_part[warp] = lo_p;
      s_part[NW + warp] = hi_p;
    }
    __syncthreads();
    if (tid == 0) {
      double lo = s_part[0], hi = s_part[NW];
      for (int w = 1; w < NW; ++w) {
        lo = fmin(lo, s_part[w]);
        hi = fmax(hi, s_part[NW + w]);
      }
      const double span = hi - lo;
      lo -= 1e-3 * span + 1e-30;
      hi += 1e-3 * span + 1e-30;
      for (int s = 0; s < k; ++s) {
        s_lo[s] = lo;
        s_hi[s] = hi;
      }
      s_scr[34] = fmax(fabs(lo), fabs(hi));  // ||T|| bound
    }
    (void)nrm_p;
    __syncthreads();
  }
  const int off = P.matrix_sym ? 1 : 0;
  const int PTS = NT / k;  // points per eigenvalue per round
  {
    const int s = tid / PTS, j = tid % PTS;
    const bool active = s < k;
    // 0-based ascending index this group is after
    const int want = active ? (P.smallest ? (off + P.first + s) : (G - 1 - off - P.first - s)) : 0;
    for (int round = 0; round < 12; ++round) {
      double x = 0.0;
      int cnt = 0;
      if (active) {
        const double lo = s_lo[s], hi = s_hi[s];
        x = lo + (hi - lo) * ((double)(j + 1) / (double)(PTS + 1));
        cnt = sturm_count(s_d, s_e2, G, x);
        s_cnt[tid] = cnt;
      }
      __syncthreads();
      if (active) {
        // counts are monotone in j: the bracket is between the last point with cnt <= want and the next one
        if (cnt <= want && (j == PTS - 1 || s_cnt[tid + 1] > want)) s_lo[s] = x;
        if (cnt > want && (j == 0 || s_cnt[tid - 1] <= want)) s_hi[s] = x;
      }
      __syncthreads();
      bool done = true;
      for (int q = 0; q < k; ++q) done = done && (s_hi[q] - s_lo[q] <= 8e-16 * s_scr[34]);
      if (done) break;
    }
    if (tid < k) s_lam[tid] = 0.5 * (s_lo[tid] + s_hi[tid]);
    __syncthreads();
  }

  // ================================================================ 6. inverse iteration on the tridiagonal
  // thread 0 of warp s handles eigenvector s (spread over warps so the serial chains run on different schedulers)
  const double tnorm = s_scr[34];
  auto inverse_iterate = [&](int s, int n_orth) {
    double* u0 = s_lu + (size_t)s * 5 * GP;
    double* u1 = u0 + GP;
    double* u2 = u1 + GP;
    double* lm = u2 + GP;
    double* y = lm + GP;
    unsigned char* piv = s_piv + (size_t)s * GP;
    double* z = s_Z + (size_t)s * GP;
    const double lam = s_lam[s];
    const double tiny = 2.3e-16 * tnorm + 1e-300;
    // pivoted LU of T - lam I
    double r0 = s_d[0] - lam, r1 = (G > 1) ? s_e[0] : 0.0, r2 = 0.0;
    for (int i = 0; i < G - 1; ++i) {
      double q0 = s_e[i], q1 = s_d[i + 1] - lam, q2 = (i + 2 < G) ? s_e[i + 1] : 0.0;
      unsigned char pv = 0;
      if (fabs(q0) > fabs(r0)) {
        pv = 1;
        double t;
        t = r0, r0 = q0, q0 = t;
        t = r1, r1 = q1, q1 = t;
        t = r2, r2 = q2, q2 = t;
      }
      if (r0 == 0.0) r0 = tiny;
      const double m = q0 / r0;
      u0[i] = 1.0 / r0;  // store reciprocals: the solves then only multiply
      u1[i] = r1;
      u2[i] = r2;
      lm[i] = m;
      piv[i] = pv;
      r0 = q1 - m * r1;
      r1 = q2 - m * r2;
      r2 = 0.0;
    }
    if (fabs(r0) < tiny) r0 = (r0 < 0.0) ? -tiny : tiny;
    u0[G - 1] = 1.0 / r0;
    u1[G - 1] = 0.0;
    u2[G - 1] = 0.0;
    for (int i = 0; i < G; ++i) z[i] = (double)((i * 7919) % 13 - 6) * (1.0 / 6.0) + 0.37;
    for (int it = 0; it < 3; ++it) {
      for (int i = 0; i < G; ++i) y[i] = z[i];
      for (int i = 0; i < G - 1; ++i) {
        if (piv[i]) {
          const double t = y[i];
          y[i] = y[i + 1];
          y[i + 1] = t;
        }
        y[i + 1] -= lm[i] * y[i];
      }
      double x1 = 0.0, x2 = 0.0, nn = 0.0;
      for (int i = G - 1; i >= 0; --i) {
        const double xi = (y[i] - u1[i] * x1 - u2[i] * x2) * u0[i];
        z[i] = xi;
        x2 = x1;
        x1 = xi;
      }
      // scale early to dodge overflow, orthogonalise inside a cluster, normalise
      double mx = 0.0;
      for (int i = 0; i < G; ++i) mx = fmax(mx, fabs(z[i]));
      const double im = 1.0 / mx;
      for (int i = 0; i < G; ++i) z[i] *= im;
      for (int p = s - n_orth; p < s; ++p) {
        const double* zp = s_Z + (size_t)p * GP;
        double dt = 0.0;
        for (int i = 0; i < G; ++i) dt += zp[i] * z[i];
        for (int i = 0; i < G; ++i) z[i] -= dt * zp[i];
      }
      for (int i = 0; i < G; ++i) nn += z[i] * z[i];
      const double inrm = 1.0 / sqrt(nn);
      for (int i = 0; i < G; ++i) z[i] *= inrm;
    }
  };
  if (lane == 0 && warp < k) inverse_iterate(warp, 0);
  __syncthreads();
  if (tid == 0) {
    // clusters (eigenvalues closer than 1e-7 ||T||): redo members with in-loop Gram-Schmidt against predecessors
    int run = 0;
    for (int s = 1; s < k; ++s) {
      run = (fabs(s_lam[s] - s_lam[s - 1]) < 1e-7 * tnorm) ? run + 1 : 0;
      if (run > 0) inverse_iterate(s, run);
    }
  }
  __syncthreads();

  // ================================================================ 7. back-transformation, one warp per eigenvector
  for (int s = warp; s < k; s += NW) {
    double* z = s_Z + (size_t)s * GP;
    for (int kc = G - 3; kc >= 0; --kc) {
      const int o = kc + 1, n = G - o;
      const double tau = s_tau[kc];
      if (tau == 0.0) continue;
      const double* v = Mat + (size_t)kc * G + o;
      double dt = 0.0;
      for (int i = lane; i < n; i += 32) dt = fma(v[i], z[o + i], dt);
      dt = warp_sum_d(dt) * tau;
      for (int i = lane; i < n; i += 32) z[o + i] = fma(-dt, v[i], z[o + i]);
      __syncwarp();
    }
    // ================================================================ 8a. normalise + sign
    double nn = 0.0;
    for (int i = lane; i < G; i += 32) nn += z[i] * z[i];
    nn = warp_sum_d(nn);
    double sc = 1.0 / sqrt(nn);
    if (P.sign_rule) {
      const double z0 = z[0] * sc;
      double ref = z0;
      if (fabs(z0) < 1e-6) {  // fallback: entry of largest magnitude, lowest index on ties
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int i = lane; i < G; i += 32) {
          const double a = fabs(z[i]);
          if (a > best) best = a, bi = i;
        }
        for (int o2 = 16; o2 >= 1; o2 >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o2);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
          if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        ref = z[bi];
      }
      if (ref < 0.0) sc = -sc;
    }
    __syncwarp();
    for (int i = lane; i < G; i += 32) z[i] *= sc;
  }
  __syncthreads();

  // ================================================================ 8b. outputs + rank-count argsort
  for (int e = tid; e < G * k; e += NT) {
    const int i = e / k, s = e % k;
    P.eigvecs[((size_t)b * G + i) * k + s] = (float)s_Z[(size_t)s * GP + i];
  }
  if (tid < k) P.eigvals[(size_t)b * k + tid] = (float)s_lam[tid];
  for (int e = tid; e < G * k; e += NT) {
    const int s = e / G, i = e % G;
    const double* z = s_Z + (size_t)s * GP;
    const double zi = z[i];
    int rank = 0;
    for (int q = 0; q < G; ++q) {
      const double zq = z[q];
      rank += (zq < zi) || (zq == zi && q < i);
    }
    P.perm[((size_t)b * k + s) * G + rank] = i;
    if (P.inv_perm) P.inv_perm[((size_t)b * k + s) * G + i] = rank;
  }
}

static bool spectral_lu_alias(int G, int k, bool in_smem) {
  const size_t GP = (G + 31) & ~31;
  return in_smem && (size_t)G * G * 4 >= (size_t)k * 5 * GP * 8;
}

static size_t spectral_smem_bytes(int G, int k, int NT, bool in_smem) {
  const size_t GP = (G + 31) & ~31;
  size_t s = 16;
  if (in_smem) s += (size_t)G * G * 12;
  s += 6 * GP * 8 + (size_t)NT * 8 + (size_t)k * GP * 8 + 40 * 8 + 3 * 8 * 8;
  if (!spectral_lu_alias(G, k, in_smem)) s += (size_t)k * 5 * GP * 8;
  s += (size_t)NT * 4 + GP * 3 * 4 + GP * 4 + (size_t)k * GP;
  return (s + 15) & ~(size_t)15;
}

size_t spectral_workspace_bytes(int B, int G, int k) {
  if (G <= 0 || B <= 0) return 0;
  const int NT = G <= 64 ? 256 : 512;
  if (spectral_smem_bytes(G, k, NT, true) <= 227 * 1024) return 0;  // (the workspace path itself runs 1024 threads)
  return (size_t)B * G * G * 12 + 256;
}

int spectral_eig(SpectralParams P, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SIM_REQUIRE(P.B > 0 && P.G >= 4 && P.G <= 512, SIM_ERR_INVALID, "spectral_eig: G must be in [4, 512] (got %d)", P.G);
  // first may be -1 with SIM_LAP_SYMMETRIC: the pair that variant drops is then returned too (full decomposition)
  SIM_REQUIRE(P.k >= 1 && P.k <= 8 && P.first + (P.matrix_sym ? 1 : 0) >= 0 && P.first + P.k + (P.matrix_sym ? 1 : 0) <= P.G,
              SIM_ERR_INVALID, "spectral_eig: k must be in [1, 8] and first + k <= G (got k=%d first=%d)", P.k, P.first);
  SIM_REQUIRE(P.adj_in || (P.k_nn >= 1 && P.k_nn + 1 <= P.G), SIM_ERR_INVALID, "spectral_eig: k_nn+1 must be <= G");
  SIM_REQUIRE((P.center || P.adj_in) && P.eigvals && P.eigvecs && P.perm, SIM_ERR_INVALID, "spectral_eig: null tensor");
  int NT = P.G <= 64 ? 256 : 512;
  const bool in_smem = spectral_smem_bytes(P.G, P.k, NT, true) <= 227 * 1024;
  // matrix in the L2-resident workspace (G > 128): one CTA per cloud is latency-bound on its column / row walks, so it runs
  // 1024 threads (twice the loads in flight; 9.1 -> see profiles/r02_sweep_c5.jsonl)
  if (!in_smem) NT = 1024;
  const size_t smem = spectral_smem_bytes(P.G, P.k, NT, in_smem);
  P.lu_alias = spectral_lu_alias(P.G, P.k, in_smem) ? 1 : 0;
  if (!in_smem) {
    const size_t need = spectral_workspace_bytes(P.B, P.G, P.k);
    SIM_REQUIRE(workspace && workspace_bytes >= need, SIM_ERR_WORKSPACE,
                "spectral_eig: G=%d needs a %zu-byte workspace (see sim_spectral_eig_workspace_bytes)", P.G, need);
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
    P.ws_mat = reinterpret_cast<double*>(base);
    P.ws_adj = reinterpret_cast<float*>(base + (size_t)P.B * P.G * P.G * 8);
  }
#define SIM_SPEC_LAUNCH(NT_, INS)                                                                   \
  do {                                                                                              \
    auto kern = spectral_kernel<NT_, INS>;                                                          \
    static SmemAttrCache attr; /* per device, grow-only, set by the first (warm-up) call */             \
    if (smem > 48 * 1024 && ensure_dyn_smem(kern, smem, attr) != cudaSuccess)                        \
      return check_launch("spectral_eig attr");                                                     \
    kern<<<P.B, NT_, smem, stream>>>(P);                                                            \
  } while (0)
  if (NT == 1024) {
    SIM_SPEC_LAUNCH(1024, false);
  } else if (NT == 256) {
    SIM_SPEC_LAUNCH(256, true);
  } else {
    SIM_SPEC_LAUNCH(512, true);
  }
#undef SIM_SPEC_LAUNCH
  return check_launch("spectral_eig");
}

// sigma = mean over the WHOLE (B, G, G) tensor of pairwise centre distances (torch.mean(dist_matrix), :628): one CTA per
// cloud leaves its fp64 sum in `partial`, the last kernel adds the B partials in index order (deterministic).
__global__ void __launch_bounds__(256) dist_sum_kernel(const float* __restrict__ center, int G, double* __restrict__ partial) {
  extern __shared__ float sc[];
  __shared__ double red[8];
  const float* cen = center + (size_t)blockIdx.x * G * 3;
  for (int i = threadIdx.x; i < G * 3; i += 256) sc[i] = cen[i];
  __syncthreads();
  double s = 0.0;
  for (int e = threadIdx.x; e < G * G; e += 256) s += (double)__fsqrt_rn(sqdist3s(sc + 3 * (e / G), sc + 3 * (e % G)));
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void dist_mean_kernel(const double* __restrict__ partial, int B, double count, float* __restrict__ sigma) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < B; ++b) t += partial[b];
    *sigma = (float)(t / count);
  }
}

int pairwise_dist_mean(const float* center, int B, int G, double* partial, float* sigma, cudaStream_t stream) {
  SIM_REQUIRE(center && partial && sigma && B > 0 && G >= 1 && G <= 4096, SIM_ERR_INVALID, "pairwise_dist_mean: bad arguments");
  dist_sum_kernel<<<B, 256, (size_t)G * 3 * sizeof(float), stream>>>(center, G, partial);
  dist_mean_kernel<<<1, 32, 0, stream>>>(partial, B, (double)B * G * G, sigma);
  return check_launch("pairwise_dist_mean");
}

}  // namespace sim
