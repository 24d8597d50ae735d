// gemm_host.cuh — host side of the tcgen05 3xTF32 GEMM: TMA descriptor creation (driver entry point
// resolved at run time, no link-time libcuda dependency) and the templated launcher.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gemm_tf32x3.cuh"

namespace rama {

typedef CUresult (*TmapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmapEncodeTiledFn tmap_encode_fn() {
  static TmapEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmapEncodeTiledFn>(p);
  });
  return fn;
}

// f32 matrix [rows][K] with row pitch `ld` floats; box = box_rows × BK, swizzle = BK·4 bytes; out-of-bounds → 0
inline bool make_tmap_2d(CUtensorMap* m, const float* base, size_t rows, size_t K, size_t ld, int box_rows, int BK) {
  TmapEncodeTiledFn f = tmap_encode_fn();
  if (!f) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 4) % 16 || box_rows > 256 || box_rows < 1) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int g_gemm_pf_rows = -1;   // RAMA_GEMM_PFROWS (read once): k-blocks per L2 row burst of the streamed A operand (0: box prefetch, RAMA_GEMM_PF)
inline int g_gemm_pf_ahead = -1;  // RAMA_GEMM_PF (read once): k-blocks of L2 prefetch ahead of the ring when the A operand streams from HBM
inline long long* g_gemm_trace = nullptr;  // debug: set by rama_debug_gemm_trace (device buffer [128][8]) or null

struct GemmOperand {  // a K-major f32 matrix [rows][K], row pitch ld
  const float* p;
  size_t rows, ld;
};

// C (per group g) = A[ga] · B[gb]ᵀ.  n_a > 1: groups differ in A (weights as the 128-row operand, batched decode);
// n_b > 1: groups differ in B (prefill QKV); a dual epilogue stacks B[0] on B[1] inside one tile.
template <int BN, int STAGES, int CH, int NX, int AS = 0, int PS = 0, int NACC = 2, int PAIR = 0, class Epi>
cudaError_t launch_gemm_tf32x3(cudaStream_t st, const GemmOperand* A, int n_a, const GemmOperand* B, int n_b, int M,
                               int N, int K, int hi_round, int ksplit, const Epi& epi, bool pdl = false) {
  using SM = GemmSmem<BN, STAGES, NX, AS, PS, NACC, PAIR>;
  constexpr int BK = kGemmBK;
  static_assert(SM::kTotal <= 227 * 1024, "tile does not fit shared memory");
  auto kern = gemm_tf32x3_kernel<BN, STAGES, CH, NX, Epi, AS, PS, NACC, PAIR>;
  static std::atomic<unsigned long long> attr_done{0};
  const cudaError_t attr_err = ensure_dyn_smem((const void*)kern, SM::kTotal, attr_done);
  if (attr_err != cudaSuccess) return attr_err;
  if (M <= 0 || N <= 0 || K <= 0 || K % 4 || n_a < 1 || n_a > 3 || n_b < 1 || n_b > 3 || ksplit < 1) return cudaErrorInvalidValue;
  if (n_a > 1 && (n_b > 1 || Epi::kDual)) return cudaErrorInvalidValue;
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  // PS: the B operand is the pre-split [2·BN][K] matrix; PAIR: each CTA of the pair loads half of the B tile
  constexpr int box_n = (Epi::kDual || PAIR) ? BN / 2 : (PS ? 2 * BN : BN);
  for (int i = 0; i < 3; ++i) {
    const GemmOperand& a = A[i < n_a ? i : 0];
    const GemmOperand& b = B[i < n_b ? i : 0];
    if (!make_tmap_2d(&maps.a[i], a.p, a.rows, (size_t)K, a.ld, kGemmBM, BK)) return cudaErrorInvalidValue;
    if (!make_tmap_2d(&maps.b[i], b.p, b.rows, (size_t)K, b.ld, box_n, BK)) return cudaErrorInvalidValue;
  }
  if (g_gemm_pf_ahead < 0) {
    const char* v = getenv("RAMA_GEMM_PF");
    g_gemm_pf_ahead = v && *v ? atoi(v) : 2;  // measured on B200 (batched-decode GEMMs, µs): 0 → 70.4/126.5/63.1, 2 → 66.6/119.1/59.8, 8 → 75.9/133.0/68.6
  }
  // only where A is the streamed operand (weights as the 128-row operand: BN = 64, batched decode)
  if (g_gemm_pf_rows < 0) {
    const char* v = getenv("RAMA_GEMM_PFROWS");
    // measured on B200 (7B decode shapes, round-1 tile): 0 (box prefetch) 9.35 ms per 64-sequence step, 8 → 12.2, 16 → 11.2:
    // the bursts lose — kept as a run-time option only
    g_gemm_pf_rows = v && *v ? atoi(v) : 0;
  }
  const int pf_rows = BN == 64 ? g_gemm_pf_rows : 0;
  GemmShape shp{M, N, K, hi_round, ksplit, n_a > 1 ? 1 : 0, (BN == 64 && pf_rows == 0) ? g_gemm_pf_ahead : 0, pf_rows,
                {A[0].p, A[n_a > 1 ? 1 : 0].p, A[n_a > 2 ? 2 : 0].p}, (long long)A[0].ld, g_gemm_trace};
  const int groups = Epi::kDual ? 1 : std::max(n_a, n_b);
  constexpr int tile_n = Epi::kDual ? BN / 2 : BN;
  dim3 grid((M + kGemmBM - 1) / kGemmBM, (N + tile_n - 1) / tile_n, groups * ksplit);
  if (PAIR) grid.x = (grid.x + 1) / 2 * 2;  // whole pairs: a CTA beyond M loads zeros (TMA out-of-bounds fill) and stores nothing
  if (grid.y > 65535) return cudaErrorInvalidValue;
  if (pdl || PAIR) {  // programmatic dependent launch: barrier init / TMEM allocation overlap the previous kernel's tail
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = SM::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (PAIR) {  // the two CTAs of a pair sit on the two SMs of one TPC
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, maps, shp, epi);
  }
  kern<<<grid, kGemmThreads, SM::kTotal, st>>>(maps, shp, epi);
  return cudaGetLastError();
}

}  // namespace rama
