// tp_exchange.cuh — bulk tensor-parallel exchange for the row-batched paths (prompt prefill, batched decode).
//
// Row-parallel wo / w2 leave a partial [rows][D] on every rank (the reference is single-device).  Round 1
// all-reduced those partials with ncclAllReduce — 8 MB per call at 512 prompt rows, 64 calls per pass, unoverlapped:
// 5.5 of 16.1 ms at 8 GPUs.  Here the exchange is part of the kernels on either side of it, over peer-mapped HBM (NVLink):
//
//   reduce-scatter   the producing GEMM's epilogue stores row m of its partial straight into the inbox of the rank that
//                    OWNS row m (rows are dealt in contiguous blocks of rpr = ceil(rows / P)): inbox[src rank][row][D].
//                    The stores leave while the GEMM is still computing other tiles — the transfer overlaps the math.
//   add + rmsnorm    tp_addnorm_kernel on the owner: x[m] += Σ_src inbox[src][m] in rank order (bit-identical on every
//                    rank), rmsnorm (cpu.rs:99-117) — the residual stream is row-parallel, only the owner keeps x[m].
//   all-gather       the same kernel stores the normalised row into EVERY rank's xn buffer (the next GEMM's A operand).
//
// Synchronisation is three epoch flags per (rank, peer) in peer-mapped memory, written with system-scope fences:
//   A  "my partial rows are in your inbox"   sent by the first CTA of tp_addnorm (the producing GEMM is the previous kernel
//      of the stream: its remote stores are complete when this kernel starts; fence.sc.sys + flag store publish them)
//   B  "my normalised rows are in your xn"   sent by the last CTA of tp_addnorm to finish; that CTA then waits for every
//      rank's B before it exits, so the kernel boundary orders the next GEMM's TMA loads after all arrivals.
//   C  plain barrier (logits all-gather of the batched step).
// Buffer reuse needs no double buffering: a rank pushes into inbox again only after its own next tp_addnorm... of the
// FOLLOWING exchange, which it enters only after receiving B of this exchange from everybody (all readers done).
#pragma once
#include "common.cuh"

namespace rama {

struct TpPeers {
  unsigned* flags[kMaxPeers];  // rank q's flag array [3][P] (peer-mapped; [me] is local)
  int P, me;
  __device__ __forceinline__ unsigned* flags_of(int r) const {
    unsigned* f = flags[0];
#pragma unroll
    for (int i = 1; i < kMaxPeers; ++i) f = (r == i) ? flags[i] : f;
    return f;
  }
};

struct TpEpoch {  // host value, or (captured graphs) derived from a device counter bumped once per step
  const unsigned* dev;
  unsigned mul, add;
  __device__ __forceinline__ unsigned get() const { return dev ? __ldcg(dev) * mul + add : add; }
};

// lanes 0..P-1 of the calling warp: publish everything this GPU wrote before (fence) and raise flag `which` at every rank
__device__ __forceinline__ void tp_signal(const TpPeers& tp, int which, unsigned epoch) {
  const int lane = threadIdx.x & 31;
  __threadfence_system();
  if (lane < tp.P) {
    unsigned* f = tp.flags_of(lane) + which * tp.P + tp.me;
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
}
// lanes 0..P-1 of the calling warp spin until every rank's flag `which` has reached `epoch` (bounded: ~4 s → error 3)
__device__ __forceinline__ void tp_wait(const TpPeers& tp, int which, unsigned epoch, int32_t* error) {
  const int lane = threadIdx.x & 31;
  if (lane < tp.P) {
    const unsigned* f = tp.flags_of(tp.me) + which * tp.P + lane;  // the LOCAL array: peers write into it
    unsigned v;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if ((++spins & 255u) == 0) {
        if (!t0) t0 = globaltimer_ns();
        else if (globaltimer_ns() - t0 > 4000000000ull) { *error = 3; break; }
        if (*reinterpret_cast<volatile int32_t*>(error) == 3) break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

struct TpAddNormParams {
  float* x;               // local residual rows [rows][D]; only this rank's rows are kept current after the first exchange
  const float* inbox;     // local [n_slab][rpr][D]; slab = source rank (× split-K partial for the batched step)
  int n_slab;
  size_t slab_stride;     // floats between slabs (rpr · D)
  const float* w;         // rmsnorm weight, or null: residual add only (last exchange of a prefill pass)
  float* xn[kMaxPeers];   // every rank's normalised-activation buffer [rows][D] (peer-mapped)
  size_t lo_off;          // > 0: also store the tf32 remainder of the row at xn + lo_off (pre-split B operand, batched decode)
  float* xlast[kMaxPeers];  // optional: every rank's x0[D] receives the updated residual of global row last_row
  int last_row;
  int row0, n_rows, rpr;  // this rank owns global rows [row0, row0 + n_rows)
  int D;
  TpPeers tp;
  TpEpoch epoch;
  unsigned* done;         // local counter of finished CTAs (returns to 0)
  int32_t* error;
};

constexpr int kTpNormThreads = 512;

// grid = max(1, n_rows) CTAs; CTA i handles local row i.
static __global__ void __launch_bounds__(kTpNormThreads) tp_addnorm_kernel(const TpAddNormParams p) {
  pdl_launch_dependents(); pdl_wait();  // the producing GEMM (previous kernel) has completed, its remote stores included
  __shared__ float red[2 * kWarp];
  __shared__ int s_last;
  const unsigned ep = p.epoch.get();
  const int warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && warp == 0) tp_signal(p.tp, 0, ep);  // A: my partials are in your inboxes
  if (warp == 0) tp_wait(p.tp, 0, ep, p.error);              //    everybody's partials are in mine
  __syncthreads();
  const int i = blockIdx.x;
  if (i < p.n_rows) {
    const int m = p.row0 + i;
    const int D4 = p.D >> 2;
    float4* xr = reinterpret_cast<float4*>(p.x + (size_t)m * p.D);
    const float4* in = reinterpret_cast<const float4*>(p.inbox + (size_t)i * p.D);
    const size_t ss4 = p.slab_stride >> 2;
    float ss = 0.f;
    for (int c = threadIdx.x; c < D4; c += kTpNormThreads) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      for (; s + 4 <= p.n_slab; s += 4) {  // four independent loads in flight, summed in slab (= rank, split) order
        const float4 t0 = __ldcg(in + (size_t)s * ss4 + c), t1 = __ldcg(in + (size_t)(s + 1) * ss4 + c),
                     t2 = __ldcg(in + (size_t)(s + 2) * ss4 + c), t3 = __ldcg(in + (size_t)(s + 3) * ss4 + c);
        a.x += t0.x; a.y += t0.y; a.z += t0.z; a.w += t0.w;
        a.x += t1.x; a.y += t1.y; a.z += t1.z; a.w += t1.w;
        a.x += t2.x; a.y += t2.y; a.z += t2.z; a.w += t2.w;
        a.x += t3.x; a.y += t3.y; a.z += t3.z; a.w += t3.w;
      }
      for (; s < p.n_slab; ++s) {
        const float4 t = __ldcg(in + (size_t)s * ss4 + c);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      float4 v = xr[c];
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      xr[c] = v;
      ss = dot4(v, v, ss);
      if (m == p.last_row) {
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
          if (q < p.tp.P && p.xlast[q]) reinterpret_cast<float4*>(p.xlast[q])[c] = v;
      }
    }
    if (p.w) {
      ss = block_sum<kTpNormThreads>(ss, red);
      const float scale = 1.0f / sqrtf(ss / (float)p.D + 1e-5f);
      const float4* w4 = reinterpret_cast<const float4*>(p.w);
      for (int c = threadIdx.x; c < D4; c += kTpNormThreads) {  // same c as above: each thread re-reads its own writes
        const float4 v = xr[c], g = __ldg(w4 + c);
        const float4 o = make_float4(g.x * (scale * v.x), g.y * (scale * v.y), g.z * (scale * v.z), g.w * (scale * v.w));
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)  // all-gather by remote stores: the row lands in every rank's xn
          if (q < p.tp.P) reinterpret_cast<float4*>(p.xn[q] + (size_t)m * p.D)[c] = o;
        if (p.lo_off) {
          const float4 ol = tf32_lo4(o);
#pragma unroll
          for (int q = 0; q < kMaxPeers; ++q)
            if (q < p.tp.P) reinterpret_cast<float4*>(p.xn[q] + p.lo_off + (size_t)m * p.D)[c] = ol;
        }
      }
    }
  }
  // B: the last CTA of this rank to finish publishes the rows, then holds the kernel open until every rank has done so
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(p.done, 1u);
    s_last = old == gridDim.x - 1;
    if (s_last) *p.done = 0u;
  }
  __syncthreads();
  if (s_last && warp == 0) {
    tp_signal(p.tp, 1, ep);
    tp_wait(p.tp, 1, ep, p.error);
  }
}

// plain barrier over the TP group (flag C): everything this rank's earlier kernels stored into peer memory is visible
// to every rank's later kernels once this kernel has completed everywhere
static __global__ void __launch_bounds__(32) tp_barrier_kernel(const TpPeers tp, const TpEpoch epoch, int32_t* error) {
  pdl_launch_dependents(); pdl_wait();
  const unsigned ep = epoch.get();
  tp_signal(tp, 2, ep);
  tp_wait(tp, 2, ep, error);
}

// select with static indices (a runtime index would spill the pointer table to local memory)
__device__ __forceinline__ float* tp_sel(float* const (&tbl)[kMaxPeers], int r) {
  float* f = tbl[0];
#pragma unroll
  for (int i = 1; i < kMaxPeers; ++i) f = (r == i) ? tbl[i] : f;
  return f;
}

// ---- pushing the partial rows to their owners ------------------------------------------------------------------------------

// prefill: C row m (prompt row), columns n..n+31 → inbox of rank m / rpr, slab `me`.
// A thread of the GEMM epilogue holds 32 columns of ONE row, so a warp-wide store would scatter 32 × 16 bytes over 32 rows —
// harmless into local L2, but over NVLink every 16-byte piece travels as its own write (measured: the pushing wo / w2 GEMMs
// of a TP = 4 prefill ran 22 % slower than the storing ones).  The 32×32 block therefore goes through 4 KB of shared memory
// private to the warp (`scratch`, set by the kernel — the pipeline stages are idle by then) and leaves as whole 128-byte row
// segments: 8 lanes per row, 4 rows per store instruction.
struct EpiPushNT {
  static constexpr bool kDual = false;
  float* inbox[kMaxPeers];
  int rpr, me, ldc, N;
  float* scratch;
  __device__ __forceinline__ void operator()(int m, int n, const float (&v)[32], int, int, bool valid) const {
    const int lane = threadIdx.x & 31;
    float4* s4 = reinterpret_cast<float4*>(scratch);
#pragma unroll
    for (int c = 0; c < 8; ++c)  // row `lane`, 16-byte chunk c, XOR-swizzled: conflict-free both ways
      s4[lane * 8 + (c ^ (lane & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int m_base = m - lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3), c = lane & 7;
      if ((vmask >> r) & 1u) {
        const float4 x = s4[r * 8 + (c ^ (r & 7))];
        const int mr = m_base + r, q = mr / rpr, col = n + 4 * c;
        float* dst = tp_sel(inbox, q) + ((size_t)me * rpr + (mr - q * rpr)) * ldc + col;
        if (col + 3 < N) {
          *reinterpret_cast<float4*>(dst) = x;
        } else {
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < N) dst[e] = xv[e];
        }
      }
    }
    __syncwarp();
  }
};

// batched decode: Σ over the split-K partials of this rank, then ONE store of every reduced row into its owner's inbox
// (slab `me`) — pushing the S partials themselves would put S times the bytes on NVLink.  Grid (ceil(D/4/256), n).
struct TpInboxes { float* p[kMaxPeers]; };
static __global__ void __launch_bounds__(256) tp_sum_push_kernel(const float* __restrict__ part, int S, size_t slab, const TpInboxes inbox,
                                                          int bpr, int me, int D) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= (D >> 2)) return;
  const float4* src = reinterpret_cast<const float4*>(part + (size_t)b * D) + i;
  const size_t slab4 = slab >> 2;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < S; ++s) {  // split order: deterministic
    const float4 t = src[(size_t)s * slab4];
    a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
  }
  const int q = b / bpr;
  reinterpret_cast<float4*>(tp_sel(inbox.p, q) + ((size_t)me * bpr + (b - q * bpr)) * D)[i] = a;
}

}  // namespace rama
