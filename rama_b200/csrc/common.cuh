// common.cuh — device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDACC_RTC__
#include <atomic>
#endif

namespace rama {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: a process that drives several devices
// (rama_ctx_create_multi) must apply it on each of them, not once per process.  `mask` = one bit per device done.
inline cudaError_t ensure_dyn_smem(const void* kern, int bytes, std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
  return e;
}

constexpr int kWarp = 32;
constexpr int kAttnMaxHs = 128;  // largest head_size the attention kernels hold per head (checked at load)

// Device-resident step control block (one per session).  Kernels read pos/token from here so
// that one captured CUDA graph can be replayed for every position.
struct StepCtrl {
  int32_t pos;          // position of the step being computed
  int32_t token;        // input token of the step
  int32_t chained;      // 1: device-resident generate loop (token feedback on device)
  int32_t n_prompt;     // chained mode: prompt length
  float temperature;    // chained mode sampling parameters
  float topp;
  int32_t next;         // output of the sampler
  int32_t error;        // sticky device-side error flag (e.g. empty top-p candidate list)
};

// 128-bit streaming load: read-only path, do not allocate in L1 (weights are read once per token).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Exact remainder of an f32 after the tensor core's tf32 truncation (it ignores the 13 low mantissa bits): the lo plane of a
// pre-split B operand (gemm_tf32x3.cuh, PS mode); the raw value itself serves as the hi operand.
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float4 tf32_lo4(const float4& v) { return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)); }

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum with a fixed association (deterministic run to run). red: >= 32 floats smem.
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red) {
  constexpr int NW = NT / kWarp;
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect red from a previous use
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < NW) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the total
}

// Programmatic dependent launch (PDL): let the next kernel in the stream start its
// weight-prefetch prologue while this one drains; wait before touching activations.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 prefetch of a contiguous byte range (bytes % 16 == 0, 16-byte aligned), issued by one thread.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- thread-block cluster helpers (distributed shared memory) -------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem, unsigned rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_st_f4(uint32_t addr, const float4& v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void dsmem_st_f1(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- peer-memory exchange (tensor parallelism over NVLink; see gemv.cuh PeerOut/PeerIn) -----------
// Low-latency "LL" protocol: every 4-byte payload travels in ONE 8-byte store together with a 32-bit
// epoch, so arrival of the epoch proves arrival of the payload — no fences, no atomics, no flags;
// the reader spins per element on the epoch (normally already there).
constexpr int kMaxPeers = 8;

__device__ __forceinline__ void st_ll(void* p, unsigned payload, unsigned epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(epoch) : "memory");
}
// two adjacent LL elements in ONE 16-byte store (each half still carries its own epoch, so a split delivery is harmless)
__device__ __forceinline__ void st_ll2(void* p, unsigned payload0, unsigned payload1, unsigned epoch) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(p), "r"(payload0), "r"(epoch), "r"(payload1) : "memory");
}
__device__ __forceinline__ uint4 ld_ll2(const void* p) {  // two adjacent {payload, epoch} elements
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Reads two LL elements, spinning until both carry `epoch`.  Bounded: after ~4 s the sticky error
// flag is raised instead of hanging the GPU (a peer process died).
__device__ __forceinline__ uint4 ld_ll2_wait(const void* p, unsigned epoch, int32_t* error) {
  uint4 v = ld_ll2(p);
  if (v.y == epoch && v.w == epoch) return v;
  if (*reinterpret_cast<volatile int32_t*>(error) == 3) return v;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  for (;;) {
    v = ld_ll2(p);
    if (v.y == epoch && v.w == epoch) break;
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ull) {
      *error = 3;
      break;
    }
  }
  return v;
}

// (value, index) argmax merge with the reference's tie rule: later index wins on ties
// (cpu.rs:165-167: `if v1 > v2 {a} else {b}` in a left fold).
__device__ __forceinline__ void argmax_merge(float& bv, int& bi, float v, int i) {
  // (bv,bi) and (v,i) are partial results over disjoint index sets.
  if (v > bv || (v == bv && i > bi)) { bv = v; bi = i; }
}

}  // namespace rama
