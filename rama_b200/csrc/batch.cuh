// batch.cuh — multi-sequence batched decode: one step for B concurrent sequences (the server path:
// the reference runs one RunState and one forward()/sample() loop per request task, lib.rs:127-160, so B
// requests stream the weights B times).  Here the B token vectors go through every weight matrix together
// as a tensor-core contraction (gemm_tf32x3.cuh, weights as the 128-row operand, split-K so that all SMs
// stream), which reads the weights once per step; each sequence keeps its own position and its own
// session KV cache, attention is the per-sequence flash-decode kernel with the batch as a third grid axis.
// The small kernels here finish the split-K partials: they sum the splits in a fixed order (deterministic)
// and apply what the batch-1 GEMV epilogues apply (RoPE + KV write, SwiGLU, residual + rmsnorm, logits).
#pragma once
#include "attention.cuh"
#include "common.cuh"

namespace rama {

struct BatchSeq {  // one sequence of a batched step (device array, rewritten by the host every step)
  float* key_cache;    // the session's caches [L][T][Dq]
  float* value_cache;
  float* logits;       // the session's logits [V]
  StepCtrl* ctrl;      // the session's control block (pos/token mirrored there; sampler output)
  int32_t pos, token;
  float* logits_peer[kMaxPeers];  // tensor parallelism over peer memory: the same session's logits on every rank
};

// out[i] = Σ_s part[s][i], s ascending
static __global__ void sum_partials_kernel(float* __restrict__ out, const float* __restrict__ part, size_t n, int S) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * n + i];
    out[i] = a;
  }
}

// x[b] = token_embedding_table[token_b] (infer.rs:13); mirrors (token,pos) into the session's control block
// chained (rama_generate_batch: the loop stays on the device): the step's (token, pos) come from the sequence's control block,
// where the sampler of the previous step left them, and are copied into the table the other kernels of the step read.
static __global__ void __launch_bounds__(256) batch_embed_kernel(BatchSeq* __restrict__ seqs, const float* __restrict__ emb,
                                                          float* __restrict__ x, int D, int vocab, unsigned* step_counter,
                                                          int chained) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  if (step_counter && blockIdx.x == 0 && threadIdx.x == 0) *step_counter += 1u;  // epoch source of the step's TP exchanges
  const BatchSeq sq = seqs[blockIdx.x];
  int token = chained ? sq.ctrl->token : sq.token;
  if (threadIdx.x == 0) {
    if (chained) {
      seqs[blockIdx.x].pos = sq.ctrl->pos;
      seqs[blockIdx.x].token = token;
      if (token < 0 || token >= vocab) sq.ctrl->error = 1;
    } else {
      // the error flag describes THIS step (rama_sample_batch reads it per step): a session handed back to a pool and reused
      // must not carry an earlier request's error
      sq.ctrl->pos = sq.pos; sq.ctrl->token = token; sq.ctrl->chained = 0;
      sq.ctrl->error = (token < 0 || token >= vocab) ? 1 : 0;
    }
  }
  if (token < 0 || token >= vocab) token = 0;
  const float4* src = reinterpret_cast<const float4*>(emb + (size_t)token * D);
  float4* dst = reinterpret_cast<float4*>(x + (size_t)blockIdx.x * D);
  for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) dst[i] = src[i];
}

// x[b] += Σ_s y[s][b] (pending residual as split-K partials, may be null); xn[b] = w·(scale·x[b])
// ≙ array_add (cpu.rs:16-21) + rmsnorm (cpu.rs:99-117); one 1024-thread CTA per sequence, the S partial loads of an
// element issued back to back (the kernel is pure latency: 64 CTAs, a few KB each)
constexpr int kBatchNormThreads = 1024;
// xn_lo (optional): the tf32 remainder plane of xn — xn is the pre-split B operand of the next GEMM (gemm_tf32x3.cuh, PS)
static __global__ void __launch_bounds__(kBatchNormThreads) batch_addnorm_kernel(float* __restrict__ x, const float* __restrict__ y,
                                                                          int S, size_t slab, const float* __restrict__ w,
                                                                          float* __restrict__ xn, int D, float* __restrict__ xn_lo) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  __shared__ float red[2 * kWarp];
  const int b = blockIdx.x;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)b * D);
  float ss = 0.f;
  for (int i = threadIdx.x; i < (D >> 2); i += kBatchNormThreads) {
    float4 v = xr[i];
    if (y) {
      const float4* yp = reinterpret_cast<const float4*>(y + (size_t)b * D) + i;
      const size_t slab4 = slab >> 2;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      for (; s + 4 <= S; s += 4) {  // four independent loads in flight, summed in split order
        const float4 t0 = yp[(size_t)s * slab4], t1 = yp[(size_t)(s + 1) * slab4], t2 = yp[(size_t)(s + 2) * slab4],
                     t3 = yp[(size_t)(s + 3) * slab4];
        a.x += t0.x; a.y += t0.y; a.z += t0.z; a.w += t0.w;
        a.x += t1.x; a.y += t1.y; a.z += t1.z; a.w += t1.w;
        a.x += t2.x; a.y += t2.y; a.z += t2.z; a.w += t2.w;
        a.x += t3.x; a.y += t3.y; a.z += t3.z; a.w += t3.w;
      }
      for (; s < S; ++s) {
        const float4 t = yp[(size_t)s * slab4];
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      xr[i] = v;
    }
    ss = dot4(v, v, ss);
  }
  ss = block_sum<kBatchNormThreads>(ss, red);
  const float scale = 1.0f / sqrtf(ss / (float)D + 1e-5f);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* o = reinterpret_cast<float4*>(xn + (size_t)b * D);
  float4* ol = xn_lo ? reinterpret_cast<float4*>(xn_lo + (size_t)b * D) : nullptr;
  for (int i = threadIdx.x; i < (D >> 2); i += kBatchNormThreads) {
    const float4 v = xr[i], g = w4[i];
    const float4 r = make_float4(g.x * (scale * v.x), g.y * (scale * v.y), g.z * (scale * v.z), g.w * (scale * v.w));
    o[i] = r;
    if (ol) ol[i] = tf32_lo4(r);
  }
}

// [wq;wk;wv] partials [3][S][B][Dq] → sum, RoPE on q,k at the sequence's own position (cpu.rs:74-97),
// q → Q[b], k/v → the session's cache row pos_b of this layer (infer.rs:31-33).  Grid (B, ceil(Dq/2/256)).
static __global__ void __launch_bounds__(256) batch_qkv_finish_kernel(const float* __restrict__ part, int S, size_t slab,
                                                               const BatchSeq* __restrict__ seqs, size_t layer_off,
                                                               float* __restrict__ q, const float* __restrict__ freq_real,
                                                               const float* __restrict__ freq_imag, int Dq, int hs2) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.x;
  const int i = blockIdx.y * 256 + threadIdx.x;  // pair index
  if (2 * i >= Dq) return;
  const BatchSeq sq = seqs[b];
  float2 v[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    float2 a = make_float2(0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float2 t = reinterpret_cast<const float2*>(part + (size_t)(g * S + s) * slab + (size_t)b * Dq)[i];
      a.x += t.x; a.y += t.y;
    }
    v[g] = a;
  }
  const int f = sq.pos * hs2 + i % hs2;
  const float c = freq_real[f], sn = freq_imag[f];
  const float2 qo = make_float2(__fsub_rn(__fmul_rn(v[0].x, c), __fmul_rn(v[0].y, sn)),
                                __fadd_rn(__fmul_rn(v[0].x, sn), __fmul_rn(v[0].y, c)));
  const float2 ko = make_float2(__fsub_rn(__fmul_rn(v[1].x, c), __fmul_rn(v[1].y, sn)),
                                __fadd_rn(__fmul_rn(v[1].x, sn), __fmul_rn(v[1].y, c)));
  reinterpret_cast<float2*>(q + (size_t)b * Dq)[i] = qo;
  reinterpret_cast<float2*>(sq.key_cache + layer_off + (size_t)sq.pos * Dq)[i] = ko;
  reinterpret_cast<float2*>(sq.value_cache + layer_off + (size_t)sq.pos * Dq)[i] = v[2];
}

// attention: grid (heads, splits, B) — the batch-1 flash-decode CTA with per-sequence cache and position
struct AttnBatchParams {
  const BatchSeq* seqs;
  const float* q;        // [B][Dq]
  float* out;            // [B][Dq]
  float* ws;             // [B][H][n_split][hs+2]
  unsigned int* tickets; // [B][H]
  size_t layer_off;
  int T, Dq, hs, n_split, H;
  float* out_lo;         // optional [B][Dq]: tf32 remainder plane of `out` (pre-split B operand of the wo GEMM)
};
static __global__ void __launch_bounds__(kAttnThreads) attn_decode_batch_kernel(const AttnBatchParams bp) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.z;
  const BatchSeq sq = bp.seqs[b];
  AttnParams p;
  p.q = bp.q + (size_t)b * bp.Dq;
  p.key_cache = sq.key_cache + bp.layer_off;
  p.value_cache = sq.value_cache + bp.layer_off;
  p.out = bp.out + (size_t)b * bp.Dq;
  p.att = nullptr;
  p.ws = bp.ws + (size_t)b * bp.H * bp.n_split * (bp.hs + 2);
  p.tickets = bp.tickets + (size_t)b * bp.H;
  p.ctrl = nullptr; p.pos_override = sq.pos;
  p.T = bp.T; p.Dq = bp.Dq; p.hs = bp.hs; p.n_split = bp.n_split;
  p.prefetch = nullptr; p.prefetch_bytes = 0;
  p.out_lo = bp.out_lo ? bp.out_lo + (size_t)b * bp.Dq : nullptr;
  attn_decode_body(p, sq.pos);
}

// [w1;w3] partials [2][S][B][F] → hb[b][j] = (h1·(1/(1+exp(−h1))))·h3   (cpu.rs:54-64)
static __global__ void __launch_bounds__(256) batch_swiglu_finish_kernel(const float* __restrict__ part, int S, size_t slab,
                                                                  float* __restrict__ hb, int F, int B, float* __restrict__ hb_lo) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const size_t n = (size_t)B * F;
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    float h1 = 0.f, h3 = 0.f;
    for (int s = 0; s < S; ++s) {
      h1 += part[(size_t)s * slab + i];
      h3 += part[(size_t)(S + s) * slab + i];
    }
    const float r = (h1 * (1.0f / (1.0f + expf(-h1)))) * h3;
    hb[i] = r;
    if (hb_lo) hb_lo[i] = tf32_lo(r);  // pre-split B operand of the w2 GEMM
  }
}

// tensor parallelism: staging [P][B][Vl] (all-gathered) → every session's full logits [V]
static __global__ void __launch_bounds__(256) batch_logits_scatter_kernel(const float* __restrict__ staging,
                                                                   const BatchSeq* __restrict__ seqs, int Vl, int P, int B) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.y;
  float* dst = seqs[b].logits;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < P * Vl; i += gridDim.x * 256) {
    const int r = i / Vl, j = i - r * Vl;
    dst[i] = staging[((size_t)r * B + b) * Vl + j];
  }
}

// classifier partials [S][B][Vl] → staging[b][Vl] (tensor parallelism: this rank's block of the all-gather buffer)
static __global__ void __launch_bounds__(256) batch_cls_stage_kernel(const float* __restrict__ part, int S, size_t slab,
                                                              float* __restrict__ out, int Vl) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.y;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < Vl; i += gridDim.x * 256) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * slab + (size_t)b * Vl + i];
    out[(size_t)b * Vl + i] = a;
  }
}

// classifier partials [S][B][Vl] → the session's logits[v0 .. v0+Vl)
static __global__ void __launch_bounds__(256) batch_cls_finish_kernel(const float* __restrict__ part, int S, size_t slab,
                                                               const BatchSeq* __restrict__ seqs, int Vl, int v0) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.y;
  float* dst = seqs[b].logits + v0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < Vl; i += gridDim.x * 256) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * slab + (size_t)b * Vl + i];
    dst[i] = a;
  }
}

// the same under the peer exchange: the slice goes into the session's logits on EVERY rank (vocabulary all-gather by
// remote stores; a tp_barrier_kernel afterwards makes all slices visible before the samplers run)
static __global__ void __launch_bounds__(256) batch_cls_push_kernel(const float* __restrict__ part, int S, size_t slab,
                                                             const BatchSeq* __restrict__ seqs, int Vl, int v0, int P) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int b = blockIdx.y;
  const BatchSeq sq = seqs[b];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < Vl; i += gridDim.x * 256) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * slab + (size_t)b * Vl + i];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < P) sq.logits_peer[r][v0 + i] = a;
  }
}

// after sample_batch_kernel: next[b] = ctrl_b->next, err[b] = ctrl_b->error
static __global__ void batch_collect_kernel(const BatchSeq* __restrict__ seqs, int32_t* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    out[2 * b] = seqs[b].ctrl->next;
    out[2 * b + 1] = seqs[b].ctrl->error;
  }
}

}  // namespace rama
