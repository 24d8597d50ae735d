// step_kernel.cuh — the whole decode step (infer.rs:8-53 + greedy Device::sample, cpu.rs:155-168) as ONE
// persistent cooperative launch.
//
// The multi-kernel step (session.cu enqueue_step) costs ~3-4 µs per kernel boundary even inside a CUDA graph
// with programmatic dependent launch; with 1 + 5·L + 1 kernels per token that is what bounds the small
// models (stories110M: 62 kernels × 4.7 µs against 67 µs of weight streaming) and tensor parallelism at
// 8 GPUs.  Here the same prologue / row / epilogue functors (gemv.cuh) and the flash-decode work item
// (attention.cuh) run as PHASES of one kernel with one 512-thread CTA per SM; a grid-wide barrier
// (one release-add + acquire-spin on an L2 counter, ≈1 µs) replaces each kernel boundary, and before
// arriving at a barrier every CTA prefetches the head of its weight slab of the next phase into L2, so the
// HBM pipe keeps streaming while the barrier settles.
//
// Phases per layer: [x += pending; rmsnorm; wq|wk|wv; RoPE; KV write] | attention | wo | [x += ·; rmsnorm;
// w1|w3; SwiGLU] | w2, then [x += ·; final rmsnorm; wcls; per-CTA argmax] | greedy finish (CTA 0).
// Activations written by other CTAs are read with ld.global.cg (L1 is not coherent across SMs).
// Tensor parallelism: the peer-memory exchange is already part of the wo/w2 epilogues and the following
// prologues (PeerOut/PeerIn), so the kernel is the same under TP.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"

namespace rama {

struct StepParams {
  // model (this rank's shard)
  int D, Dq, Fl, L, V, Vl, v0, T, hs, Hl;
  const float *emb, *rms_att, *wq, *wk, *wv, *wo, *rms_ffn, *w1, *w2, *w3, *rms_final, *freq_real, *freq_imag, *wcls;
  // session
  float *x0, *x1, *xfinal, *xb, *xb2, *w2out, *hb, *hb2, *q, *k, *v, *att, *logits, *key_cache, *value_cache, *attn_ws;
  unsigned int* tickets;
  ArgPart* part;              // [world][grid] greedy partials (this rank's row written by the classifier phase)
  unsigned* seq;              // TP epoch counter (bumped at the end of the step)
  StepCtrl* ctrl;
  unsigned long long* bar;    // [0] barrier arrivals (monotonic), [1] persistent steps completed
  const int32_t* prompt;      // chained mode
  int32_t* out_tokens;
  int n_split;                // attention workspace stride (chunks per head)
  int wk_d, wk_wo, wk_w2;     // warps splitting K for K = D, Dq, Fl
  int mode;                   // 0: forward (logits + argmax partials); 1: + greedy sample / chained feedback
  // tensor-parallel peer exchange (world > 1 with the fused exchange; else world = 1)
  int rank, world;
  char* peer_base[kMaxPeers];
  size_t off_inbox, off_parts;
  long long* trace;           // optional: CTA 0 stamps clock64() at kernel entry, before and after every barrier
  // Cluster mode (tiny models, RAMA_STEP=cluster): the grid is ONE thread-block cluster and the phases of the LAYERS are separated
  // by barrier.cluster (hardware, ≈0.2–0.4 µs) instead of the L2-counter grid barrier (≈2 µs) or a kernel boundary (≈3 µs); the
  // kernel ends after the last w2 and the classifier runs as the usual full-grid GEMV (16 SMs cannot stream a 37 MB classifier).
  int cluster;
};

constexpr int kStepAttnWarps = kGemvWarps;  // 16 warps × 4 timesteps = 64 per attention work item

__device__ __forceinline__ void grid_barrier(unsigned long long* cnt, unsigned long long target, int32_t* error) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(cnt) : "memory");
    unsigned long long v, t0 = 0;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(cnt) : "memory");
      if (v >= target) break;
      if ((++spins & 0x3fffu) == 0) {  // bounded: a bug must not hang the GPU; once tripped, every later barrier falls through
        if (*reinterpret_cast<volatile int32_t*>(error) == 4) break;
        if (!t0) t0 = globaltimer_ns();
        else if (globaltimer_ns() - t0 > 2000000000ull) { *error = 4; break; }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ PeerIn step_peer_in(const StepParams& p, int stage, int layer) {
  PeerIn pi{};
  if (p.world <= 1) return pi;
  pi.inbox = reinterpret_cast<const uint2*>(p.peer_base[p.rank] + p.off_inbox) + (size_t)stage * p.world * p.D;
  pi.seq = p.seq; pi.error = &p.ctrl->error; pi.P = p.world; pi.n = p.D; pi.L = p.L + 1; pi.layer = layer; pi.seq_add = 1;
  return pi;
}

// Epilogues of the persistent kernel: same arithmetic as EpiStore / EpiQKV / EpiSwiGLU / EpiCls (gemv.cuh), but they
// hold only a reference to the kernel's parameter block (constant bank) and a layer index, so nothing of theirs
// occupies registers while the streaming loop runs with 16 128-bit loads in flight per thread.
struct StepEpiStore {  // wo (stage 0 → xb2) / w2 (stage 1 → w2out); under TP the outputs go to every rank's inbox
  const StepParams& p;
  int stage, layer;
  __device__ __forceinline__ void operator()(int pr, float v0, float v1) const {
    if (p.world > 1) {
      const unsigned ep = (*p.seq + 1u) * (unsigned)(p.L + 1) + (unsigned)layer + 1u;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {
        if (r < p.world) {
          uint2* inbox = reinterpret_cast<uint2*>(p.peer_base[r] + p.off_inbox) + ((size_t)stage * p.world + p.rank) * p.D;
          st_ll2(inbox + 2 * pr, __float_as_uint(v0), __float_as_uint(v1), ep);
        }
      }
      return;
    }
    float* o = stage == 0 ? p.xb2 : p.w2out;
    *reinterpret_cast<float2*>(o + 2 * pr) = make_float2(v0, v1);  // D is even
  }
  __device__ __forceinline__ void finish(float*) const {}
  __device__ __forceinline__ void prepare() const {}
  __device__ __forceinline__ void prefetch(int) const {}
};

struct StepEpiQKV {  // RoPE (cpu.rs:74-97) on q,k + KV-cache row write (infer.rs:31-33)
  const StepParams& p;
  int layer, pos;
  __device__ __forceinline__ void operator()(int pr, float v0, float v1) const {
    const int pairs_per = p.Dq >> 1, hs2 = p.hs >> 1;
    const int sec = pr / pairs_per, i = pr - sec * pairs_per;
    const size_t row = ((size_t)layer * p.T + pos) * p.Dq;
    if (sec < 2) {
      const int f = pos * hs2 + (i % hs2);
      const float c = p.freq_real[f], s = p.freq_imag[f];
      const float2 o = make_float2(__fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, s)), __fadd_rn(__fmul_rn(v0, s), __fmul_rn(v1, c)));
      if (sec == 0) {
        reinterpret_cast<float2*>(p.q)[i] = o;
      } else {
        reinterpret_cast<float2*>(p.k)[i] = o;
        reinterpret_cast<float2*>(p.key_cache + row)[i] = o;
      }
    } else {
      reinterpret_cast<float2*>(p.v)[i] = make_float2(v0, v1);
      reinterpret_cast<float2*>(p.value_cache + row)[i] = make_float2(v0, v1);
    }
  }
  __device__ __forceinline__ void finish(float*) const {}
  __device__ __forceinline__ void prepare() const {}
  __device__ __forceinline__ void prefetch(int) const {}
};

struct StepEpiSwiGLU {  // cpu.rs:54-64
  const StepParams& p;
  __device__ __forceinline__ void operator()(int pr, float h1, float h3) const {
    p.hb[pr] = (h1 * (1.0f / (1.0f + expf(-h1)))) * h3;
    p.hb2[pr] = h3;
  }
  __device__ __forceinline__ void finish(float*) const {}
  __device__ __forceinline__ void prepare() const {}
  __device__ __forceinline__ void prefetch(int) const {}
};

struct StepEpiCls {  // logits + per-CTA greedy partial (ties → higher index, cpu.rs:165-167)
  const StepParams& p;
  float bv;
  int bi;
  __device__ __forceinline__ void operator()(int pr, float v0, float v1) {
    float* lg = p.logits + p.v0;
    lg[2 * pr] = v0;
    argmax_merge(bv, bi, v0, p.v0 + 2 * pr);
    if (2 * pr + 1 < p.Vl) {
      lg[2 * pr + 1] = v1;
      argmax_merge(bv, bi, v1, p.v0 + 2 * pr + 1);
    }
  }
  __device__ __forceinline__ void prepare() const {}
  __device__ __forceinline__ void prefetch(int) const {}
  __device__ __forceinline__ void finish(float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_merge(bv, bi, ov, oi);
    }
    int* redi = reinterpret_cast<int*>(red + kGemvWarps);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { red[w] = bv; redi[w] = bi; }
    __syncthreads();
    if (w == 0) {
      float tv = l < kGemvWarps ? red[l] : -INFINITY;
      int ti = l < kGemvWarps ? redi[l] : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, tv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ti, o);
        argmax_merge(tv, ti, ov, oi);
      }
      if (p.world > 1) {
        if (l < p.world) {  // {value, epoch}, {index, epoch} into rank l's array
          char* base = p.peer_base[0];
#pragma unroll
          for (int r = 1; r < kMaxPeers; ++r) base = (l == r) ? p.peer_base[r] : base;
          uint2* dst = reinterpret_cast<uint2*>(base + p.off_parts) + ((size_t)p.rank * gridDim.x + blockIdx.x) * 2;
          const unsigned ep = (*p.seq + 1u) * (unsigned)(p.L + 1) + (unsigned)p.L + 1u;
          st_ll(dst, __float_as_uint(tv), ep);
          st_ll(dst + 1, (unsigned)ti, ep);
        }
      } else if (l == 0) {
        p.part[blockIdx.x].v = tv;
        p.part[blockIdx.x].i = ti;
      }
    }
  }
};

static __global__ void __launch_bounds__(kGemvThreads, 1) decode_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ float4 gemv_smem[];
  __shared__ float red[2 * kWarp];
  float4* xs = gemv_smem;

  StepCtrl* ctrl = p.ctrl;
  const int pos = ctrl->pos;
  int token = ctrl->token;
  if (token < 0 || token >= p.V) {  // the reference would panic on the slice (infer.rs:13)
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->error = 1;
    token = 0;
  }
  const int D = p.D, Dq = p.Dq, Fl = p.Fl, L = p.L;
  const unsigned long long step_id = p.bar[1];
  unsigned long long target = step_id * (unsigned long long)(5 * L + 1) * gridDim.x;  // 5L+1 barriers per step
  int n_stamp = 0;
  auto stamp = [&]() {
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[n_stamp++] = clock64();
  };
  stamp();
  auto barrier = [&]() {
    target += gridDim.x;
    stamp();
    if (p.cluster) cluster_sync_all();  // release / acquire at cluster scope: covers the global-memory activations too
    else grid_barrier(&p.bar[0], target, &p.ctrl->error);
    stamp();
  };
  constexpr size_t kPrefetch = 192 * 1024;

  for (int l = 0; l < L; ++l) {
    // ---- rmsnorm → [wq|wk|wv] → RoPE → KV write (infer.rs:13,19-33); layer 0 reads the embedding row itself
    {
      ProNorm pro{l == 0 ? p.emb + (size_t)token * D : p.x0, l == 0 ? nullptr : p.w2out, p.x1,
                  p.rms_att + (size_t)l * D, nullptr, step_peer_in(p, 1, l - 1)};
      RowsQKV rows{p.wq + (size_t)l * Dq * D, p.wk + (size_t)l * Dq * D, p.wv + (size_t)l * Dq * D, D, Dq / 2};
      gemv_phase(pro, rows, StepEpiQKV{p, l, pos}, D / 4, 3 * Dq / 2, p.wk_d, xs, red);
    }
    {  // HBM idles during attention: pull (most of) this layer's wo into L2 meanwhile
      RowsPlain rows_wo{p.wo + (size_t)l * D * Dq, Dq, D};
      gemv_prefetch_slab(rows_wo, D / 2, 4 * kPrefetch);
    }
    barrier();
    // ---- attention (infer.rs:34): work items (head, 64-timestep chunk)
    {
      AttnParams ap{p.q, p.key_cache + (size_t)l * p.T * Dq, p.value_cache + (size_t)l * p.T * Dq, p.xb, p.att, p.attn_ws,
                    p.tickets, nullptr, pos, p.T, Dq, p.hs, p.n_split, nullptr, 0};
      const int n_chunks = (pos + 1 + kStepAttnWarps * kAttnPerWarp - 1) / (kStepAttnWarps * kAttnPerWarp);
      const int items = p.Hl * n_chunks;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        attn_item<kStepAttnWarps>(ap, pos, it % p.Hl, it / p.Hl);
        __syncthreads();
      }
    }
    barrier();
    // ---- wo (infer.rs:35); the residual add (:37) is folded into the next prologue
    {
      RowsPlain rows_wo{p.wo + (size_t)l * D * Dq, Dq, D};
      gemv_phase(ProPlain{p.xb}, rows_wo, StepEpiStore{p, 0, l}, Dq / 4, D / 2, p.wk_wo, xs, red);
    }
    {
      RowsW13 rows_w13{p.w1 + (size_t)l * Fl * D, p.w3 + (size_t)l * Fl * D, D};
      gemv_prefetch_slab(rows_w13, Fl, kPrefetch);
      barrier();
      // ---- x += xb2; rmsnorm → [w1|w3] → SwiGLU (infer.rs:37-45)
      ProNorm pro{p.x1, p.xb2, p.x0, p.rms_ffn + (size_t)l * D, nullptr, step_peer_in(p, 0, l)};
      gemv_phase(pro, rows_w13, StepEpiSwiGLU{p}, D / 4, Fl, p.wk_d, xs, red);
    }
    {
      RowsPlain rows_w2{p.w2 + (size_t)l * D * Fl, Fl, D};
      gemv_prefetch_slab(rows_w2, D / 2, kPrefetch);
      barrier();
      // ---- w2 (infer.rs:46); residual add (:47) folded into the next prologue
      gemv_phase(ProPlain{p.hb}, rows_w2, StepEpiStore{p, 1, l}, Fl / 4, D / 2, p.wk_w2, xs, red);
    }
    if (l + 1 < L) {
      RowsQKV nxt{p.wq + (size_t)(l + 1) * Dq * D, p.wk + (size_t)(l + 1) * Dq * D, p.wv + (size_t)(l + 1) * Dq * D, D, Dq / 2};
      gemv_prefetch_slab(nxt, 3 * Dq / 2, kPrefetch);
    } else if (p.cluster) {
      return;  // the classifier GEMV and the sampler follow as kernels of their own (stream order: this kernel's w2 is complete)
    } else {
      RowsPlain nxt{p.wcls, D, p.Vl};
      gemv_prefetch_slab(nxt, (p.Vl + 1) / 2, kPrefetch);
    }
    barrier();
  }
  // ---- x += w2out; final rmsnorm → wcls → logits + per-CTA argmax partial (infer.rs:49-51)
  {
    ProNorm pro{p.x0, p.w2out, p.x1, p.rms_final, p.xfinal, step_peer_in(p, 1, L - 1)};
    RowsPlain rows{p.wcls, D, p.Vl};
    gemv_phase(pro, rows, StepEpiCls{p, -INFINITY, -1}, D / 4, (p.Vl + 1) / 2, p.wk_d, xs, red);
  }
  barrier();
  // ---- greedy Device::sample (cpu.rs:163-168: ties → the later index) + generate()'s bookkeeping (mod.rs:187-203)
  if (blockIdx.x == 0) {
    if (p.mode == 1) {
      float bv = -INFINITY;
      int bi = -1;
      if (p.world > 1) {
        const uint2* inbox = reinterpret_cast<const uint2*>(p.peer_base[p.rank] + p.off_parts);
        const unsigned ep = (*p.seq + 1u) * (unsigned)(L + 1) + (unsigned)L + 1u;
        for (int i = threadIdx.x; i < p.world * (int)gridDim.x; i += kGemvThreads) {
          const uint4 e = ld_ll2_wait(inbox + 2 * (size_t)i, ep, &ctrl->error);
          if ((int)e.z >= 0) argmax_merge(bv, bi, __uint_as_float(e.x), (int)e.z);
        }
      } else {
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kGemvThreads) {
          const float pv = __ldcg(&p.part[i].v);
          const int pi = __ldcg(&p.part[i].i);
          if (pi >= 0) argmax_merge(bv, bi, pv, pi);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_merge(bv, bi, ov, oi);
      }
      int* redi = reinterpret_cast<int*>(red + kWarp);
      const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
      if (ln == 0) { red[w] = bv; redi[w] = bi; }
      __syncthreads();
      if (w == 0) {
        bv = ln < kGemvWarps ? red[ln] : -INFINITY;
        bi = ln < kGemvWarps ? redi[ln] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          argmax_merge(bv, bi, ov, oi);
        }
        if (ln == 0) {
          int next = bi < 0 ? 0 : bi;
          if (ctrl->chained) {
            if (pos < ctrl->n_prompt) next = p.prompt[pos];  // prompt forcing (mod.rs:190-191)
            p.out_tokens[pos] = next;
            ctrl->token = next;                              // token feedback stays on the device
            ctrl->pos = pos + 1;
          }
          ctrl->next = next;
        }
      }
    }
    if (threadIdx.x == 0) {
      p.bar[1] = step_id + 1;
      *p.seq += 1u;
    }
  }
}

}  // namespace rama
