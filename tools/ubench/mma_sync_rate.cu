// mma_sync_rate.cu — throughput of the LEGACY warp-level tensor path (mma.sync m16n8k8 tf32 → SASS HMMA.1688.F32.TF32) on sm_100a
// (a measurement tool, not product code).  One CTA per SM, W warps per scheduler, every warp keeps CH independent accumulator
// chains in flight; prints scheduler cycles per HMMA.  The prefill attention kernel (prefill.cuh) is bounded by this number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_sync_rate mma_sync_rate.cu
#include <cstdint>
#include <cstdio>
#include <vector>

template <int CH>
__global__ void rate_kernel(long long* out, float* sink, int iters) {
  float d[CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
  const uint32_t a0 = __float_as_uint(1.0f + threadIdx.x), b0 = __float_as_uint(0.5f);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a0), "r"(a0), "r"(a0), "r"(a0), "r"(b0), "r"(b0));
  }
  const long long t1 = clock64();
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int CH>
__global__ void rate_bf16_kernel(long long* out, float* sink, int iters) {
  float d[CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
  const uint32_t a0 = 0x3f803f80u, b0 = 0x3f003f00u;   // bf16 pairs (1.0, 1.0) / (0.5, 0.5)
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a0), "r"(a0), "r"(a0), "r"(a0), "r"(b0), "r"(b0));
  }
  const long long t1 = clock64();
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int CH>
static void run_bf16(int warps_per_sched, int iters) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* out; float* sink;
  cudaMalloc(&out, sms * sizeof(long long)); cudaMalloc(&sink, 4);
  const int threads = 32 * 4 * warps_per_sched;
  rate_bf16_kernel<CH><<<sms, threads>>>(out, sink, 16);
  rate_bf16_kernel<CH><<<sms, threads>>>(out, sink, iters);
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  const double per_sched = (double)iters * CH * warps_per_sched;
  printf("bf16 m16n8k16: chains %d warps/scheduler %d: %.2f cycles per HMMA.16816 per scheduler  (%.0f bf16 MAC/clk/SM)  err=%s\n", CH,
         warps_per_sched, mx / per_sched, 4.0 * 2048.0 * per_sched / mx, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(sink);
}

template <int CH>
static void run(int warps_per_sched, int iters) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* out; float* sink;
  cudaMalloc(&out, sms * sizeof(long long)); cudaMalloc(&sink, 4);
  const int threads = 32 * 4 * warps_per_sched;
  rate_kernel<CH><<<sms, threads>>>(out, sink, 16);   // warm-up
  rate_kernel<CH><<<sms, threads>>>(out, sink, iters);
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  const double per_sched = (double)iters * CH * warps_per_sched;   // HMMAs one scheduler issued
  printf("chains %d warps/scheduler %d: %.2f cycles per HMMA.1688.TF32 per scheduler  (%.0f tf32 MAC/clk/SM)  err=%s\n", CH,
         warps_per_sched, mx / per_sched, 4.0 * 1024.0 * per_sched / mx, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int w : {1, 2, 4}) { run<1>(w, 4096); run<4>(w, 4096); run<8>(w, 4096); }
  for (int w : {1, 4}) { run_bf16<4>(w, 4096); run_bf16<8>(w, 4096); }
  return 0;
}
