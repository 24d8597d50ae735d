// attention.cuh — split-K flash-decode attention over the device-resident KV cache.
//
// Replaces calculate_attention + update_xb (math.cu:72-114; one thread per head, and a grid
// bug at 7B — SURVEY App. C) ≙ Device::multi_head_attention (cpu.rs:23-52):
//     score_t = (q·k_t) / sqrt(hs);  p = softmax(score_0..pos);  xb = Σ_t p_t · v_t
//
// Grid (heads, splits): CTA (h, c) owns timesteps [c·TC, (c+1)·TC) of head h.  The kernel is
// latency-bound at short contexts (a few KB per head), so it is organised for memory-level
// parallelism: 8 warps × 4 timesteps each; every lane issues its 4 K-row and 4 V-row 128-bit loads
// back to back before the first dot product.  Each warp folds its 4 timesteps into a running
// (max, sum, acc[hs]) (online softmax), the 8 warps merge through shared memory and the CTA writes
// one partial; the last CTA of a head to finish (atomic ticket) merges the partials in split order
// (deterministic) and writes xb — no second launch.
// KV cache layout is the reference's [T][Dq] per layer (head h at column h·hs): each timestep's
// head row is hs·4 contiguous bytes (512 B at 7B) ⇒ one coalesced 128-bit load per lane.
// While it runs, HBM is otherwise idle: every CTA (also the ones beyond `pos` that exit at once)
// prefetches its slice of the NEXT kernel's weights (wo of this layer) into L2.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace rama {

constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / kWarp;
constexpr int kAttnPerWarp = 4;                          // timesteps per warp
constexpr int kAttnChunk = kAttnWarps * kAttnPerWarp;    // 32 timesteps per CTA

struct AttnParams {
  const float* q;            // [H_loc*hs]
  const float* key_cache;    // this layer: [T][Dq]
  const float* value_cache;
  float* out;                // xb [H_loc*hs]
  float* att;                // optional [H_loc][T] (reference RunState.att) or nullptr
  float* ws;                 // workspace [H_loc][n_split][hs + 2]
  unsigned int* tickets;     // [H_loc], zero between launches
  const StepCtrl* ctrl;      // pos read from here unless pos_override >= 0
  int pos_override;
  int T, Dq, hs, n_split;
  const float* prefetch;     // next kernel's weights (or nullptr)
  size_t prefetch_bytes;
  float* out_lo = nullptr;   // optional: tf32 remainder of `out` (the pre-split B operand of the batched wo GEMM)
};

// One (head h, chunk) work item of the flash-decode pass for the sequence described by p, executed by a
// CTA of NW warps (NW·4 timesteps per item).  q is read with ld.global.cg: inside the persistent step kernel
// it was written by other CTAs earlier in the same launch.
template <int NW>
__device__ __forceinline__ void attn_item(const AttnParams& p, int pos, int h, int chunk) {
  constexpr int kChunk = NW * kAttnPerWarp;
  constexpr int kThreads = NW * kWarp;
  __shared__ float s_m[NW], s_l[NW];
  __shared__ __align__(16) float s_acc[NW][kAttnMaxHs];
  __shared__ unsigned int s_ticket;
  const int n = pos + 1;
  const int n_chunks = (n + kChunk - 1) / kChunk;
  if (chunk >= n_chunks) return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hs = p.hs, hs4 = hs >> 2;
  const bool active = lane < hs4;
  const float div = sqrtf((float)hs);
  const size_t col = (size_t)h * hs;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // this warp's timesteps: t0 + j, j < kAttnPerWarp
  const int t0 = chunk * kChunk + warp * kAttnPerWarp;
  float4 kk[kAttnPerWarp], vv[kAttnPerWarp];
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) {
    const bool ok = active && (t0 + j) < n;
    kk[j] = ok ? reinterpret_cast<const float4*>(p.key_cache + (size_t)(t0 + j) * p.Dq + col)[lane] : zero4;
  }
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) {
    const bool ok = active && (t0 + j) < n;
    vv[j] = ok ? reinterpret_cast<const float4*>(p.value_cache + (size_t)(t0 + j) * p.Dq + col)[lane] : zero4;
  }
  const float4 q4 = active ? __ldcg(reinterpret_cast<const float4*>(p.q + col) + lane) : zero4;

  float sc[kAttnPerWarp];
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) sc[j] = dot4(q4, kk[j], 0.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int j = 0; j < kAttnPerWarp; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);

  float m = -INFINITY, l = 0.f;
  float4 acc = zero4;
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) {
    if (t0 + j < n) {                                  // warp-uniform
      const float s = sc[j] / div;                     // divide, as cpu.rs:41
      if (p.att && lane == 0) p.att[(size_t)h * p.T + t0 + j] = s;  // raw score; normalised by the merger
      const float mn = fmaxf(m, s);
      const float f = expf(m - mn), pe = expf(s - mn);
      l = l * f + pe;
      acc.x = acc.x * f + pe * vv[j].x; acc.y = acc.y * f + pe * vv[j].y;
      acc.z = acc.z * f + pe * vv[j].z; acc.w = acc.w * f + pe * vv[j].w;
      m = mn;
    }
  }

  // merge the warps of this CTA (fixed order)
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (active) reinterpret_cast<float4*>(s_acc[warp])[lane] = acc;
  __syncthreads();
  float* wsp = p.ws + ((size_t)h * p.n_split + chunk) * (hs + 2);
  {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) M = fmaxf(M, s_m[w]);
    // a warp that saw no timestep has m = -inf, l = 0, acc = 0: exp(-inf - M) = 0 contributes nothing
    // (warp 0 always has one, so M is finite)
    for (int i = threadIdx.x; i < hs; i += kThreads) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) a += s_acc[w][i] * expf(s_m[w] - M);
      wsp[2 + i] = a;
    }
    if (threadIdx.x == 0) {
      float L = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) L += s_l[w] * expf(s_m[w] - M);
      wsp[0] = M;
      wsp[1] = L;
    }
  }

  // ticket: last CTA of this head merges all chunks
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.tickets[h], 1u);
  __syncthreads();
  if (s_ticket != (unsigned)(n_chunks - 1)) return;
  __threadfence();

  const float* wh = p.ws + (size_t)h * p.n_split * (hs + 2);
  float M = -INFINITY;
  for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, __ldcg(wh + (size_t)c * (hs + 2)));
  float L = 0.f;
  for (int c = 0; c < n_chunks; ++c)
    L += __ldcg(wh + (size_t)c * (hs + 2) + 1) * expf(__ldcg(wh + (size_t)c * (hs + 2)) - M);
  for (int i = threadIdx.x; i < hs; i += kThreads) {
    float a = 0.f;
#pragma unroll 4
    for (int c = 0; c < n_chunks; ++c)
      a += __ldcg(wh + (size_t)c * (hs + 2) + 2 + i) * expf(__ldcg(wh + (size_t)c * (hs + 2)) - M);
    const float o = a / L;
    p.out[(size_t)h * hs + i] = o;
    if (p.out_lo) p.out_lo[(size_t)h * hs + i] = tf32_lo(o);
  }
  if (p.att) {
    for (int t = threadIdx.x; t < n; t += kThreads) {
      const float s = __ldcg(p.att + (size_t)h * p.T + t);
      p.att[(size_t)h * p.T + t] = expf(s - M) / L;
    }
  }
  if (threadIdx.x == 0) p.tickets[h] = 0u;  // ready for the next launch
}

// CTA (head = blockIdx.x, y = blockIdx.y) of a flash-decode launch: chunks y, y + gridDim.y, … of the head.
// The launch does not need one CTA per possible chunk of the whole context window (2048 CTAs at 7B, nearly all of
// which would exit at once): the host picks gridDim.y for the current position range.
__device__ __forceinline__ void attn_decode_body(const AttnParams& p, int pos) {
  const int n_chunks = (pos + kAttnChunk) / kAttnChunk;
  for (int chunk = blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
    attn_item<kAttnWarps>(p, pos, blockIdx.x, chunk);
    __syncthreads();
  }
}

static __global__ void __launch_bounds__(kAttnThreads) attn_decode_kernel(const AttnParams p, int use_pdl, unsigned long long* trace = nullptr) {
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;  // rama_step_timeline
  if (tr) trace[0] = globaltimer_ns();
  if (use_pdl) pdl_launch_dependents();
  if (p.prefetch && threadIdx.x == 0) {  // weights do not depend on the previous kernel
    const size_t n_cta = (size_t)gridDim.x * gridDim.y, me = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    const size_t per = ((p.prefetch_bytes + n_cta - 1) / n_cta + 15) & ~(size_t)15;
    const size_t off = me * per;
    if (off < p.prefetch_bytes) {
      const size_t n = min(per, p.prefetch_bytes - off) & ~(size_t)15;
      if (n) l2_prefetch_bulk(reinterpret_cast<const char*>(p.prefetch) + off, (uint32_t)n);
    }
  }
  if (use_pdl) pdl_wait();
  if (tr) trace[1] = trace[2] = globaltimer_ns();
  attn_decode_body(p, p.pos_override >= 0 ? p.pos_override : p.ctrl->pos);
  if (tr) trace[3] = globaltimer_ns();
}

// ---- flash-decode with the splits of a head in ONE thread-block cluster ---------------------------------------
// attn_decode_kernel merges the splits of a head through global memory: partial → __threadfence → atomic ticket →
// the last CTA re-reads the partials in three dependent L2 round trips.  At decode contexts of a few hundred
// positions the launch is pure latency (8 µs for ~4 MB of K/V at 7B), and that chain is half of it.  Here the
// CS CTAs that share a head form a cluster: each folds its chunks (rank, rank+CS, …) into a running (M, L, acc[hs]),
// publishes it in its own shared memory, and after ONE cluster barrier every CTA merges hs/CS output elements
// straight out of its siblings' shared memory (DSMEM) in split order — no workspace, no fence, no atomics.
//
// Second change: the K/V rows of EARLIER positions do not depend on the previous kernel (the fused QKV GEMV, which
// only appends row `pos`).  That kernel releases this launch after its own griddepcontrol.wait (gemv.cuh use_pdl
// bit 1), so when a CTA of this kernel starts, every kernel older than the QKV GEMV has completed: the position
// and the rows t < pos are read BEFORE griddepcontrol.wait and their latency overlaps the QKV kernel's tail.  After the
// wait only q and row `pos` remain to be fetched.
constexpr int kAttnClusterMax = 8;  // portable cluster size

static __global__ void __launch_bounds__(kAttnThreads, 2) attn_cluster_kernel(const AttnParams p, int use_pdl, unsigned long long* trace = nullptr) {
  const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;  // rama_step_timeline
  if (tr) trace[0] = globaltimer_ns();
  namespace cg = cooperative_groups;
  constexpr int NW = kAttnWarps;
  __shared__ float s_m[NW], s_l[NW];
  __shared__ __align__(16) float s_acc[NW][kAttnMaxHs];
  __shared__ __align__(16) float s_part[kAttnMaxHs + 4];  // this CTA's acc[hs] | M | L, read by the whole cluster
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int h = blockIdx.x / CS;

  if ((use_pdl & 3) == 1) pdl_launch_dependents();
  if (p.prefetch && threadIdx.x == 0) {  // HBM idles during attention: pull the next kernel's weights (wo) into L2
    const size_t n_cta = gridDim.x, me = blockIdx.x;
    const size_t per = ((p.prefetch_bytes + n_cta - 1) / n_cta + 15) & ~(size_t)15;
    const size_t off = me * per;
    if (off < p.prefetch_bytes) {
      const size_t nb = min(per, p.prefetch_bytes - off) & ~(size_t)15;
      if (nb) l2_prefetch_bulk(reinterpret_cast<const char*>(p.prefetch) + off, (uint32_t)nb);
    }
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hs = p.hs, hs4 = hs >> 2;
  const bool active = lane < hs4;
  const float div = sqrtf((float)hs);
  const size_t col = (size_t)h * hs;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* kbase = reinterpret_cast<const float4*>(p.key_cache + col) + lane;
  const float4* vbase = reinterpret_cast<const float4*>(p.value_cache + col) + lane;
  const size_t row4 = (size_t)p.Dq >> 2;

  // (older than the previous kernel ⇒ complete: see the header comment)
  const int pos = p.pos_override >= 0 ? p.pos_override : *reinterpret_cast<const volatile int32_t*>(&p.ctrl->pos);
  const int n = pos + 1;
  const int n_chunks = pos / kAttnChunk + 1;

  int c = rank;
  int t0 = c * kAttnChunk + warp * kAttnPerWarp;
  float4 kk[kAttnPerWarp], vv[kAttnPerWarp];
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) kk[j] = (active && t0 + j < pos) ? __ldcg(kbase + (size_t)(t0 + j) * row4) : zero4;
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) vv[j] = (active && t0 + j < pos) ? __ldcg(vbase + (size_t)(t0 + j) * row4) : zero4;

  if (use_pdl & 1) pdl_wait();
  if ((use_pdl & 3) == 3) pdl_launch_dependents();
  if (tr) trace[1] = trace[2] = globaltimer_ns();

  const float4 q4 = active ? __ldcg(reinterpret_cast<const float4*>(p.q + col) + lane) : zero4;
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) {  // this step's row, appended by the previous kernel
    if (active && t0 + j == pos) {
      kk[j] = __ldcg(kbase + (size_t)pos * row4);
      vv[j] = __ldcg(vbase + (size_t)pos * row4);
    }
  }

  float m = -INFINITY, l = 0.f;
  float4 acc = zero4;
  while (c < n_chunks) {
    // next pass of this CTA (contexts beyond CS·32 positions): loads in flight while this pass is folded
    const int c2 = c + CS, t2 = t0 + CS * kAttnChunk;
    float4 kn[kAttnPerWarp], vn[kAttnPerWarp];
    if (c2 < n_chunks) {
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) kn[j] = (active && t2 + j < n) ? __ldcg(kbase + (size_t)(t2 + j) * row4) : zero4;
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) vn[j] = (active && t2 + j < n) ? __ldcg(vbase + (size_t)(t2 + j) * row4) : zero4;
    }
    float sc[kAttnPerWarp];
#pragma unroll
    for (int j = 0; j < kAttnPerWarp; ++j) sc[j] = dot4(q4, kk[j], 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
#pragma unroll
    for (int j = 0; j < kAttnPerWarp; ++j) {
      if (t0 + j < n) {                                  // warp-uniform
        const float s = sc[j] / div;                     // divide, as cpu.rs:41
        const float mn = fmaxf(m, s);
        const float f = expf(m - mn), pe = expf(s - mn);
        l = l * f + pe;
        acc.x = acc.x * f + pe * vv[j].x; acc.y = acc.y * f + pe * vv[j].y;
        acc.z = acc.z * f + pe * vv[j].z; acc.w = acc.w * f + pe * vv[j].w;
        m = mn;
      }
    }
    if (c2 < n_chunks) {
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) { kk[j] = kn[j]; vv[j] = vn[j]; }
    }
    c = c2; t0 = t2;
  }

  // merge the warps of this CTA (fixed order) into s_part
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (active) reinterpret_cast<float4*>(s_acc[warp])[lane] = acc;
  __syncthreads();
  {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) M = fmaxf(M, s_m[w]);
    const bool live = M > -INFINITY;  // a CTA (or warp) without timesteps contributes (−inf, 0, 0)
    for (int i = threadIdx.x; i < hs; i += kAttnThreads) {
      float a = 0.f;
      if (live) {
#pragma unroll
        for (int w = 0; w < NW; ++w) a += s_acc[w][i] * expf(s_m[w] - M);
      }
      s_part[i] = a;
    }
    if (threadIdx.x == 0) {
      float L = 0.f;
      if (live) {
#pragma unroll
        for (int w = 0; w < NW; ++w) L += s_l[w] * expf(s_m[w] - M);
      }
      s_part[hs] = M;
      s_part[hs + 1] = L;
    }
  }
  cluster.sync();  // every CTA's partial is published (release/acquire at cluster scope)

  // CTA `rank` produces output elements [rank·per, rank·per + per) of the head from all CS partials, in split order
  {
    const int per = (hs + CS - 1) / CS;
    const int i = rank * per + (int)threadIdx.x;
    if ((int)threadIdx.x < per && i < hs) {
      float Mc[kAttnClusterMax], Lc[kAttnClusterMax], ac[kAttnClusterMax];
#pragma unroll
      for (int r = 0; r < kAttnClusterMax; ++r) {
        if (r < CS) {
          const float* rp = cluster.map_shared_rank(s_part, r);
          Mc[r] = rp[hs]; Lc[r] = rp[hs + 1]; ac[r] = rp[i];
        } else {
          Mc[r] = -INFINITY; Lc[r] = 0.f; ac[r] = 0.f;
        }
      }
      float M = -INFINITY;
#pragma unroll
      for (int r = 0; r < kAttnClusterMax; ++r) M = fmaxf(M, Mc[r]);  // chunk 0 always holds a timestep: M is finite
      float L = 0.f, a = 0.f;
#pragma unroll
      for (int r = 0; r < kAttnClusterMax; ++r) {
        const float e = expf(Mc[r] - M);
        L += Lc[r] * e;
        a += ac[r] * e;
      }
      p.out[col + i] = a / L;
    }
  }
  cluster.sync();  // no CTA may exit while a sibling still reads its shared memory
  if (tr) trace[3] = globaltimer_ns();
}

// ---- attention + wo in one launch (small models, short contexts) ---------------------------------------------
// The small models are bound by the length of the kernel chain, not by bytes.  Here CTA (h, j) computes the whole
// attention of head h over the cache (cpu.rs:23-52; redundantly for its J row-splits — a few KB of K/V at these
// sizes) and then its rows of the head's slice of wo:  part[h][r] = Σ_i wo[r][h·hs+i] · xb_h[i]  (infer.rs:35 split by
// head).  The next kernel's prologue sums the H partials in head order (deterministic) into the residual.  One launch
// fewer per layer, no tickets, no workspace.  Grid = H·J CTAs of kAttnWoThreads.
constexpr int kAttnWoThreads = 512;
constexpr int kAttnWoWarps = kAttnWoThreads / kWarp;

struct AttnWoParams {
  const float* q;            // [H*hs]
  const float* key_cache;    // this layer: [T][Dq]
  const float* value_cache;
  const float* wo;           // this layer: [D][Dq]
  float* xb;                 // [Dq] attention output (RunState.xb), written by the j = 0 CTAs
  float* part;               // [H][D] per-head partial outputs of wo
  const StepCtrl* ctrl;
  int Dq, hs, D, J;
};

static __global__ void __launch_bounds__(kAttnWoThreads) attn_wo_kernel(const AttnWoParams p, int use_pdl) {
  __shared__ float s_m[kAttnWoWarps], s_l[kAttnWoWarps];
  __shared__ __align__(16) float s_acc[kAttnWoWarps][kAttnMaxHs];
  __shared__ __align__(16) float s_xb[kAttnMaxHs];
  constexpr int kB = 8;        // timesteps per warp and batch
  constexpr int kRowsMax = 8;  // wo rows per warp held in registers (host guarantees nr ≤ kRowsMax · warps)
  if (use_pdl) { pdl_launch_dependents(); pdl_wait(); }
  const int h = blockIdx.x / p.J, j = blockIdx.x - h * p.J;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hs = p.hs, hs4 = hs >> 2;
  const bool active = lane < hs4;
  const float div = sqrtf((float)hs);
  const size_t col = (size_t)h * hs;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // this warp's rows of wo[:, h·hs .. +hs): issued first, needed last — their latency hides behind the attention
  const int per = p.D / p.J, rem = p.D % p.J;
  const int r0 = j * per + min(j, rem), nr = per + (j < rem ? 1 : 0);
  float4 wrow[kRowsMax];
#pragma unroll
  for (int i = 0; i < kRowsMax; ++i) {
    const int r = warp + i * kAttnWoWarps;
    wrow[i] = (active && r < nr) ? ldg_stream(reinterpret_cast<const float4*>(p.wo + (size_t)(r0 + r) * p.Dq + col) + lane) : zero4;
  }
  const int n = p.ctrl->pos + 1;
  const float4 q4 = active ? __ldcg(reinterpret_cast<const float4*>(p.q + col) + lane) : zero4;
  float m = -INFINITY, l = 0.f;
  float4 acc = zero4;
  for (int tb = warp; tb < n; tb += kB * kAttnWoWarps) {
    float4 kk[kB], vv[kB];
#pragma unroll
    for (int i = 0; i < kB; ++i) {
      const int t = tb + i * kAttnWoWarps;
      kk[i] = (active && t < n) ? reinterpret_cast<const float4*>(p.key_cache + (size_t)t * p.Dq + col)[lane] : zero4;
    }
#pragma unroll
    for (int i = 0; i < kB; ++i) {
      const int t = tb + i * kAttnWoWarps;
      vv[i] = (active && t < n) ? reinterpret_cast<const float4*>(p.value_cache + (size_t)t * p.Dq + col)[lane] : zero4;
    }
    float sc[kB];
#pragma unroll
    for (int i = 0; i < kB; ++i) sc[i] = dot4(q4, kk[i], 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < kB; ++i) sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], o);
#pragma unroll
    for (int i = 0; i < kB; ++i) {
      if (tb + i * kAttnWoWarps < n) {                 // warp-uniform
        const float s = sc[i] / div;                   // divide, as cpu.rs:41
        const float mn = fmaxf(m, s);
        const float f = expf(m - mn), pe = expf(s - mn);
        l = l * f + pe;
        acc.x = acc.x * f + pe * vv[i].x; acc.y = acc.y * f + pe * vv[i].y;
        acc.z = acc.z * f + pe * vv[i].z; acc.w = acc.w * f + pe * vv[i].w;
        m = mn;
      }
    }
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (active) reinterpret_cast<float4*>(s_acc[warp])[lane] = acc;
  __syncthreads();
  {  // merge the warps in fixed order (a warp without timesteps has m = -inf, l = 0: contributes nothing)
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWoWarps; ++w) M = fmaxf(M, s_m[w]);
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWoWarps; ++w) L += s_l[w] * expf(s_m[w] - M);
    for (int i = threadIdx.x; i < hs; i += kAttnWoThreads) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWoWarps; ++w) a += s_acc[w][i] * expf(s_m[w] - M);
      const float o = a / L;
      s_xb[i] = o;
      if (j == 0) p.xb[col + i] = o;
    }
  }
  __syncthreads();
  const float4 x4 = active ? reinterpret_cast<const float4*>(s_xb)[lane] : zero4;
#pragma unroll
  for (int i = 0; i < kRowsMax; ++i) {
    const int r = warp + i * kAttnWoWarps;
    const float v = warp_sum(dot4(wrow[i], x4, 0.f));
    if (lane == 0 && r < nr) p.part[(size_t)h * p.D + r0 + r] = v;
  }
}

// ---- attention + wo in one CLUSTER launch (small models) -----------------------------------------------------
// Same idea as attn_wo_kernel below (one launch fewer per layer; per-head partial outputs of wo summed by the next
// prologue), but the head's attention is computed ONCE, split over the CS CTAs of a cluster exactly like
// attn_cluster_kernel, instead of redundantly by every CTA that owns rows of wo.  After the cluster barrier each CTA
// rebuilds the head's output vector from its siblings' shared memory and multiplies it into its D/CS rows of the head's
// column block of wo — which it loaded into registers BEFORE griddepcontrol.wait (weights depend on nothing).
// Rows of hs ≤ 64 floats take half a warp: two rows per 128-bit warp load.
constexpr int kAwcRowLoads = 8;  // 128-bit wo loads per lane held in registers across the attention

static __global__ void __launch_bounds__(kAttnThreads, 2) attn_wo_cluster_kernel(const AttnWoParams p, int use_pdl) {
  namespace cg = cooperative_groups;
  constexpr int NW = kAttnWarps;
  __shared__ float s_m[NW], s_l[NW];
  __shared__ __align__(16) float s_acc[NW][kAttnMaxHs];
  __shared__ __align__(16) float s_part[kAttnMaxHs + 4];  // this CTA's acc[hs] | M | L, read by the whole cluster
  __shared__ __align__(16) float s_xb[kAttnMaxHs];
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int h = blockIdx.x / CS;
  if ((use_pdl & 3) == 1) pdl_launch_dependents();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hs = p.hs, hs4 = hs >> 2;
  const bool active = lane < hs4;
  const float div = sqrtf((float)hs);
  const size_t col = (size_t)h * hs;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // this CTA's rows [r0, r0+nr) of wo[:, h·hs .. +hs): row (k·NW + warp)·RPL + sub sits in lanes [sub·seg, sub·seg + hs4)
  const int seg = hs4 > 16 ? 32 : 16, RPL = 32 / seg;
  const int sub = lane / seg, sl = lane - sub * seg;
  const bool wactive = sl < hs4;
  const int per = p.D / CS, rem = p.D % CS;
  const int r0 = rank * per + min(rank, rem), nr = per + (rank < rem ? 1 : 0);
  float4 wrow[kAwcRowLoads];
#pragma unroll
  for (int k = 0; k < kAwcRowLoads; ++k) {
    const int r = (k * NW + warp) * RPL + sub;
    wrow[k] = (wactive && r < nr) ? ldg_stream(reinterpret_cast<const float4*>(p.wo + (size_t)(r0 + r) * p.Dq + col) + sl) : zero4;
  }

  const float4* kbase = reinterpret_cast<const float4*>(p.key_cache + col) + lane;
  const float4* vbase = reinterpret_cast<const float4*>(p.value_cache + col) + lane;
  const size_t row4 = (size_t)p.Dq >> 2;
  // older than the previous kernel (which released this launch after its own wait) ⇒ complete: see attn_cluster_kernel
  const int pos = *reinterpret_cast<const volatile int32_t*>(&p.ctrl->pos);
  const int n = pos + 1;
  const int n_chunks = pos / kAttnChunk + 1;
  int c = rank;
  int t0 = c * kAttnChunk + warp * kAttnPerWarp;
  float4 kk[kAttnPerWarp], vv[kAttnPerWarp];
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) kk[j] = (active && t0 + j < pos) ? __ldcg(kbase + (size_t)(t0 + j) * row4) : zero4;
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) vv[j] = (active && t0 + j < pos) ? __ldcg(vbase + (size_t)(t0 + j) * row4) : zero4;

  if (use_pdl & 1) pdl_wait();
  if ((use_pdl & 3) == 3) pdl_launch_dependents();

  const float4 q4 = active ? __ldcg(reinterpret_cast<const float4*>(p.q + col) + lane) : zero4;
#pragma unroll
  for (int j = 0; j < kAttnPerWarp; ++j) {  // this step's row, appended by the previous kernel
    if (active && t0 + j == pos) {
      kk[j] = __ldcg(kbase + (size_t)pos * row4);
      vv[j] = __ldcg(vbase + (size_t)pos * row4);
    }
  }
  float m = -INFINITY, l = 0.f;
  float4 acc = zero4;
  while (c < n_chunks) {
    const int c2 = c + CS, t2 = t0 + CS * kAttnChunk;
    float4 kn[kAttnPerWarp], vn[kAttnPerWarp];
    if (c2 < n_chunks) {
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) kn[j] = (active && t2 + j < n) ? __ldcg(kbase + (size_t)(t2 + j) * row4) : zero4;
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) vn[j] = (active && t2 + j < n) ? __ldcg(vbase + (size_t)(t2 + j) * row4) : zero4;
    }
    float sc[kAttnPerWarp];
#pragma unroll
    for (int j = 0; j < kAttnPerWarp; ++j) sc[j] = dot4(q4, kk[j], 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
#pragma unroll
    for (int j = 0; j < kAttnPerWarp; ++j) {
      if (t0 + j < n) {                                  // warp-uniform
        const float s = sc[j] / div;                     // divide, as cpu.rs:41
        const float mn = fmaxf(m, s);
        const float f = expf(m - mn), pe = expf(s - mn);
        l = l * f + pe;
        acc.x = acc.x * f + pe * vv[j].x; acc.y = acc.y * f + pe * vv[j].y;
        acc.z = acc.z * f + pe * vv[j].z; acc.w = acc.w * f + pe * vv[j].w;
        m = mn;
      }
    }
    if (c2 < n_chunks) {
#pragma unroll
      for (int j = 0; j < kAttnPerWarp; ++j) { kk[j] = kn[j]; vv[j] = vn[j]; }
    }
    c = c2; t0 = t2;
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (active) reinterpret_cast<float4*>(s_acc[warp])[lane] = acc;
  __syncthreads();
  {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) M = fmaxf(M, s_m[w]);
    const bool live = M > -INFINITY;
    for (int i = threadIdx.x; i < hs; i += kAttnThreads) {
      float a = 0.f;
      if (live) {
#pragma unroll
        for (int w = 0; w < NW; ++w) a += s_acc[w][i] * expf(s_m[w] - M);
      }
      s_part[i] = a;
    }
    if (threadIdx.x == 0) {
      float L = 0.f;
      if (live) {
#pragma unroll
        for (int w = 0; w < NW; ++w) L += s_l[w] * expf(s_m[w] - M);
      }
      s_part[hs] = M;
      s_part[hs + 1] = L;
    }
  }
  cluster.sync();
  // every CTA rebuilds the whole head output (it multiplies all of it into its rows of wo)
  if ((int)threadIdx.x < hs) {
    const int i = threadIdx.x;
    float Mc[kAttnClusterMax], Lc[kAttnClusterMax], ac[kAttnClusterMax];
#pragma unroll
    for (int r = 0; r < kAttnClusterMax; ++r) {
      if (r < CS) {
        const float* rp = cluster.map_shared_rank(s_part, r);
        Mc[r] = rp[hs]; Lc[r] = rp[hs + 1]; ac[r] = rp[i];
      } else {
        Mc[r] = -INFINITY; Lc[r] = 0.f; ac[r] = 0.f;
      }
    }
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < kAttnClusterMax; ++r) M = fmaxf(M, Mc[r]);
    float L = 0.f, a = 0.f;
#pragma unroll
    for (int r = 0; r < kAttnClusterMax; ++r) {
      const float e = expf(Mc[r] - M);
      L += Lc[r] * e;
      a += ac[r] * e;
    }
    const float o = a / L;
    s_xb[i] = o;
    if (rank == 0) p.xb[col + i] = o;
  }
  cluster.barrier_arrive();  // done reading the siblings' shared memory
  __syncthreads();
  const float4 x4 = wactive ? reinterpret_cast<const float4*>(s_xb)[sl] : zero4;
#pragma unroll
  for (int k = 0; k < kAwcRowLoads; ++k) {
    const int r = (k * NW + warp) * RPL + sub;
    float v = dot4(wrow[k], x4, 0.f);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    if (seg == 32) v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (sl == 0 && r < nr) p.part[(size_t)h * p.D + r0 + r] = v;
  }
  cluster.barrier_wait();  // no CTA may exit while a sibling still reads its shared memory
}

}  // namespace rama
