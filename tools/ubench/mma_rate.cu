// mma_rate.cu — issue-rate microbenchmark for tcgen05.mma.kind::tf32 on sm_100a (a measurement tool, not product code).
// One CTA per SM; one thread issues ITER MMAs back to back (M = 128, K = 8, N = 64/128/256; A from TMEM or from shared
// memory; one accumulator or two alternating), commits, waits, and reports clocks per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rama_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_tf32x3.cuh"

using namespace rama;

template <int BN, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters, int per_commit) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 16;
  uint8_t* gen = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(gen)[i] = 1.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 96 * 1024 + 16);
  // zero the A columns in TMEM (columns 384..511)
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0x3f800000u;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 384; c < 512; c += 16) tmem_st_32x16(trow + c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32<BN>();
    const uint64_t db = umma_smem_desc<32>(base);                   // B tile: BN rows x 32 floats
    const uint64_t da = umma_smem_desc<32>(base + 48 * 1024);       // A tile (SS mode): 128 rows x 32 floats
    const uint32_t hi32 = (uint32_t)(db >> 32), blo = (uint32_t)db, alo = (uint32_t)da;
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += per_commit) {
#pragma unroll 4
      for (int j = 0; j < per_commit; ++j) {
        const int ks = j & 3;
        const uint32_t acc = tmem + ((NACC > 1) ? (uint32_t)((j % NACC) * BN) : 0u);
        const uint64_t bdesc = ((uint64_t)hi32 << 32) | (blo + 2 * ks);
        if (TS) umma_tf32_ts(acc, tmem + 384 + 8 * ks, bdesc, idesc, 1);
        else umma_tf32(acc, ((uint64_t)hi32 << 32) | (alo + 2 * ks), bdesc, idesc, 1);
      }
      umma_commit(bar);
      mbar_wait(bar, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// The issuer loop of gemm_tf32x3_kernel in isolation: per "k-block" 12 MMAs, then optional commit(s) to barriers nobody waits on,
// optional wait on an already-completed barrier + tcgen05.fence::after_thread_sync.  FLAGS: 1 = commit, 2 = second commit,
// 4 = mbar_wait(completed), 8 = fence after, 16 = stamp clock64 to global twice per k-block
template <int BN, int FLAGS>
__global__ void __launch_bounds__(128, 1) issuer_loop_kernel(long long* out, long long* scratch, int kblocks) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, bar2 = bar + 8, bar3 = bar + 16, barw = bar + 24, slot = bar + 32;
  uint8_t* gen = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(gen)[i] = 1.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); mbar_init(barw, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 96 * 1024 + 32);
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0x3f800000u;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 384; c < 512; c += 16) tmem_st_32x16(trow + c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32<BN>();
    const uint64_t db = umma_smem_desc<32>(base);
    const uint32_t hi32 = (uint32_t)(db >> 32), blo = (uint32_t)db;
    const long long t0 = clock64();
    for (int kb = 0; kb < kblocks; ++kb) {
      if (FLAGS & 4) mbar_wait(barw, 1);  // fresh barrier: the "previous" phase counts as complete
      if (FLAGS & 8) tc_fence_after();
      if (FLAGS & 16) scratch[(kb & 63) * 2] = clock64();
      const uint32_t acc = tmem + (uint32_t)(((kb >> 2) & 1) * BN);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bdesc = ((uint64_t)hi32 << 32) | (blo + 2 * ks);
        umma_tf32_ts(acc, tmem + 384 + 32 + 8 * ks, bdesc, idesc, 1);
        umma_tf32_ts(acc, tmem + 384 + 8 * ks, bdesc + 1024, idesc, 1);
        umma_tf32_ts(acc, tmem + 384 + 8 * ks, bdesc, idesc, 1);
      }
      if (FLAGS & 1) umma_commit(bar);
      if (FLAGS & 2) umma_commit((kb & 1) ? bar2 : bar3);
      if (FLAGS & 16) scratch[(kb & 63) * 2 + 1] = clock64();
    }
    umma_commit(barw);
    mbar_wait(barw, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int BN, int FLAGS>
static void run_loop(int kblocks) {
  auto k = issuer_loop_kernel<BN, FLAGS>;
  const int smem = 96 * 1024 + 64 + 1024, grid = 148;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *d, *sc;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaMalloc(&sc, 1024 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, smem>>>(d, sc, kblocks);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  printf("issuer loop N=%3d flags %2d (commit %d, commit2 %d, wait %d, fence %d, stamps %d): %7.1f clk/k-block = %5.1f clk/MMA %s\n", BN,
         FLAGS, FLAGS & 1, (FLAGS >> 1) & 1, (FLAGS >> 2) & 1, (FLAGS >> 3) & 1, (FLAGS >> 4) & 1, (double)mx / kblocks,
         (double)mx / kblocks / 12, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sc);
}

template <int BN, bool TS, int NACC>
static void run(const char* name, int iters, int per_commit, int grid) {
  auto k = mma_rate_kernel<BN, TS, NACC>;
  const int smem = 96 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, smem>>>(d, iters, per_commit);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (auto v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
  printf("%-34s per_commit %4d grid %3d: %7.1f clk/MMA (min CTA %7.1f) %s\n", name, per_commit, grid, (double)mx / iters,
         (double)mn / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run_loop<128, 0>(512);  run_loop<128, 1>(512);  run_loop<128, 3>(512);  run_loop<128, 4>(512);  run_loop<128, 8>(512);
  run_loop<128, 12>(512); run_loop<128, 13>(512); run_loop<128, 15>(512); run_loop<128, 31>(512);
  run_loop<64, 0>(512);   run_loop<64, 1>(512);   run_loop<64, 13>(512);  run_loop<64, 15>(512);
  if (getenv("MMA_RATE_LOOP_ONLY")) return 0;
  const int iters = 4096;
  for (int grid : {1, 148}) {
    for (int pc : {4096, 12}) {
      run<64, true, 1>("N=64  A=TMEM 1 acc", iters / pc * pc, pc, grid);
      run<128, true, 1>("N=128 A=TMEM 1 acc", iters / pc * pc, pc, grid);
      run<256, true, 1>("N=256 A=TMEM 1 acc", iters / pc * pc, pc, grid);
      run<64, false, 1>("N=64  A=smem 1 acc", iters / pc * pc, pc, grid);
      run<128, false, 1>("N=128 A=smem 1 acc", iters / pc * pc, pc, grid);
      run<256, false, 1>("N=256 A=smem 1 acc", iters / pc * pc, pc, grid);
      run<64, true, 2>("N=64  A=TMEM 2 acc", iters / pc * pc, pc, grid);
      run<128, true, 2>("N=128 A=TMEM 2 acc", iters / pc * pc, pc, grid);
    }
  }
  return 0;
}
