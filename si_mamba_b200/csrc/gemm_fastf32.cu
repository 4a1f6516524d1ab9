// fp32-accurate projection GEMMs on the 5th-gen tensor cores (tcgen05 / TMEM), for the fp32 inference path.
//
// SI-Mamba's finetune / test runs are fp32 without autocast (tools/runner_finetune.py has no autocast), so the
// in_proj / x_proj / dt_proj / out_proj GEMMs inside Mamba.forward (models/block.py:72) are cuBLAS SIMT SGEMMs that
// never touch a tensor core: ncu puts them at 72 % of the fp32 forward on B200 (profiles/r01_launches_bench_fp32.md).
// This kernel keeps fp32 inputs, outputs and accumulation but feeds the tensor cores: every fp32 operand tile is
// split on the fly into three bf16 terms (hi + mid + lo = 24 mantissa bits) and the product is rebuilt from the
// significant bf16 x bf16 partial products with fp32 accumulation in TMEM ("9xBF16" FastFP32 emulation).  It is
// written with the CUTLASS / CuTe sm100 collective templates vendored in this image (TMA loads, tcgen05.mma,
// TMEM accumulators, warp-specialised transform / MMA / epilogue warps) instantiated inside this translation unit.
//
//   Y[M,N] = X[M,K] . W[N,K]^T      X row-major (lda), W row-major = nn.Linear.weight (ldb), Y row-major (ldd)

#include <stdlib.h>

#include "kernels.cuh"

#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/gemm/kernel/gemm_universal.hpp"

namespace sim {

namespace {

using namespace cute;

template <class TileShape_, class ClusterShape_, class Schedule_>
struct FastF32Gemm {
  using ElementA = float;
  using ElementB = float;
  using ElementC = float;
  using ElementAcc = float;
  using LayoutA = cutlass::layout::RowMajor;     // X (M, K)
  using LayoutB = cutlass::layout::ColumnMajor;  // W (N, K) row-major == (K, N) column-major
  using LayoutC = cutlass::layout::RowMajor;     // Y (M, N)
  static constexpr int kAlign = 4;               // 16-byte vectors of fp32

  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, TileShape_, ClusterShape_,
      cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, kAlign, ElementC,
      LayoutC, kAlign, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;

  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlign, ElementB, LayoutB, kAlign,
      ElementAcc, TileShape_, ClusterShape_,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(
          sizeof(typename CollectiveEpilogue::SharedStorage))>,
      Schedule_>::CollectiveOp;

  using GemmKernel =
      cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
  using StrideA = typename GemmKernel::StrideA;
  using StrideB = typename GemmKernel::StrideB;
  using StrideC = typename GemmKernel::StrideC;
  using StrideD = typename GemmKernel::StrideD;

  static typename Gemm::Arguments make_args(const float* X, long lda, const float* W, long ldb, float* Y, long ldd,
                                            int M, int N, int K) {
    StrideA sa = make_stride(int64_t(lda), Int<1>{}, int64_t(0));
    StrideB sb = make_stride(int64_t(ldb), Int<1>{}, int64_t(0));
    StrideC sc = make_stride(int64_t(ldd), Int<1>{}, int64_t(0));
    StrideD sd = make_stride(int64_t(ldd), Int<1>{}, int64_t(0));
    typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                  {M, N, K, 1},
                                  {X, sa, W, sb},
                                  {{1.0f, 0.0f}, Y, sc, Y, sd}};
    return args;
  }

  static size_t workspace(int M, int N, int K) {
    float* np = nullptr;
    return Gemm::get_workspace_size(make_args(np, K, np, K, np, N, M, N, K));
  }

  static int run(const float* X, long lda, const float* W, long ldb, float* Y, long ldd, int M, int N, int K,
                 void* ws, cudaStream_t stream) {
    Gemm gemm;
    auto args = make_args(X, lda, W, ldb, Y, ldd, M, N, K);
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) {
      set_error("gemm_f32_tc: shape M=%d N=%d K=%d (lda=%ld ldb=%ld ldd=%ld) is not supported by the tcgen05 tile", M, N,
                K, lda, ldb, ldd);
      return SIM_ERR_INVALID;
    }
    if (gemm.initialize(args, ws, stream) != cutlass::Status::kSuccess) {
      set_error("gemm_f32_tc: initialize failed");
      return SIM_ERR_CUDA;
    }
    if (gemm.run(stream) != cutlass::Status::kSuccess) {
      set_error("gemm_f32_tc: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      return SIM_ERR_CUDA;
    }
    return SIM_OK;
  }
};

// 128 x 128 x 16 MMA tile, one SM per tile (M = batch*L is large, N is 56 .. 1536)
using GemmWide = FastF32Gemm<Shape<_128, _128, _16>, Shape<_1, _1, _1>, cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>;
using GemmNarrow = FastF32Gemm<Shape<_128, _64, _16>, Shape<_1, _1, _1>, cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>;
// CTA pair (cta_group::2): 256 x 128 MMA tile over two SMs, B operand shared through the pair
using GemmPair = FastF32Gemm<Shape<_256, _128, _16>, Shape<_2, _1, _1>, cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32Sm100>;

}  // namespace

size_t gemm_f32_tc_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const size_t a = GemmWide::workspace(M, N, K), b = GemmNarrow::workspace(M, N, K), c = GemmPair::workspace(M, N, K);
  return (a > b ? (a > c ? a : c) : (b > c ? b : c)) + 256;
}

int gemm_f32_tc(const float* X, long lda, const float* W, long ldb, float* Y, long ldd, int M, int N, int K,
                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SIM_REQUIRE(M > 0 && N > 0 && K > 0 && X && W && Y, SIM_ERR_INVALID, "gemm_f32_tc: empty problem / null tensor");
  SIM_REQUIRE(aligned16(X) && aligned16(W) && aligned16(Y) && lda % 4 == 0 && ldb % 4 == 0 && ldd % 4 == 0 && K % 4 == 0 &&
                  N % 4 == 0,
              SIM_ERR_ALIGN, "gemm_f32_tc: TMA needs 16-byte aligned bases, leading dimensions, K and N");
  const size_t need = gemm_f32_tc_workspace_bytes(M, N, K);
  SIM_REQUIRE(need <= 256 || (workspace && workspace_bytes >= need), SIM_ERR_WORKSPACE,
              "gemm_f32_tc: needs a %zu-byte workspace", need);
  void* ws = workspace ? reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255) : nullptr;
  if (N <= 64) return GemmNarrow::run(X, lda, W, ldb, Y, ldd, M, N, K, ws, stream);
  static const int use_pair = [] { const char* e = getenv("SIM_GEMM_PAIR"); return e ? atoi(e) : 1; }();  // measured: 205 vs 213 us on in_proj
  if (use_pair && M >= 512 && N >= 128) return GemmPair::run(X, lda, W, ldb, Y, ldd, M, N, K, ws, stream);
  return GemmWide::run(X, lda, W, ldb, Y, ldd, M, N, K, ws, stream);
}

}  // namespace sim
