// prefill.cuh — prompt prefill: the M prompt tokens at positions [pos0, pos0+M) go through the layers
// together, so every weight matrix is read once per prompt instead of once per token and the
// contractions run on the tensor cores (gemm_tf32x3.cuh).
//
// The reference has no such function: generate() feeds the prompt one token at a time through
// forward() and discards the logits (mod.rs:187-192).  What prefill must leave behind is therefore
// exactly what those M forward() calls leave behind: the KV-cache rows pos0..pos0+M-1 of every layer
// (infer.rs:31-33) and the logits of the last position (infer.rs:51).  Everything else here restates
// infer.rs:19-47 row-wise: rmsnorm (cpu.rs:99-117), RoPE (cpu.rs:74-97), causal attention
// (cpu.rs:23-52 for each query position), SwiGLU (cpu.rs:54-64), residual adds (cpu.rs:16-21).
#pragma once
#include "common.cuh"

namespace rama {

// ---- x[m] = token_embedding_table[token[m]] (infer.rs:13, for every prompt position) -----------------
static __global__ void __launch_bounds__(256) prefill_embed_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ emb,
                                                            float* __restrict__ x, int D, int vocab, int32_t* error,
                                                            unsigned* seq) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  const int m = blockIdx.x;
  if (m == 0 && threadIdx.x == 0 && seq) *seq += 1u;  // one "step" for the TP exchange epochs (see step_begin_kernel)
  int token = tokens[m];
  if (token < 0 || token >= vocab) {  // the reference would panic on the slice (infer.rs:13)
    if (threadIdx.x == 0) *error = 1;
    token = 0;
  }
  const float4* src = reinterpret_cast<const float4*>(emb + (size_t)token * D);
  float4* dst = reinterpret_cast<float4*>(x + (size_t)m * D);
  for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) dst[i] = src[i];
}

// ---- x[m] += y[m] (pending residual, may be null);  xn[m] = w · (rsqrt-scale · x[m]) -------------------
// one CTA per row; ≙ array_add (cpu.rs:16-21) folded in front of rmsnorm (cpu.rs:99-117)
static __global__ void __launch_bounds__(256) prefill_addnorm_kernel(float* __restrict__ x, const float* __restrict__ y,
                                                              const float* __restrict__ w, float* __restrict__ xn, int D) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  __shared__ float red[2 * kWarp];
  const int m = blockIdx.x;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)m * D);
  const float4* yr = y ? reinterpret_cast<const float4*>(y + (size_t)m * D) : nullptr;
  float ss = 0.f;
  for (int i = threadIdx.x; i < (D >> 2); i += 256) {
    float4 v = xr[i];
    if (yr) {
      const float4 a = yr[i];
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      xr[i] = v;
    }
    ss = dot4(v, v, ss);
  }
  ss = block_sum<256>(ss, red);
  if (!xn) return;
  const float scale = 1.0f / sqrtf(ss / (float)D + 1e-5f);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* o = reinterpret_cast<float4*>(xn + (size_t)m * D);
  for (int i = threadIdx.x; i < (D >> 2); i += 256) {  // same i as above: each thread re-reads its own writes
    const float4 v = xr[i], g = w4[i];
    o[i] = make_float4(g.x * (scale * v.x), g.y * (scale * v.y), g.z * (scale * v.z), g.w * (scale * v.w));
  }
}

// ---- GEMM epilogues -----------------------------------------------------------------------------------------
// [wq;wk;wv] launch (group 0/1/2): RoPE on q and k with the row's own position (cpu.rs:74-97: simultaneous
// pair update, unfused mul/sub), q → Q[m][Dq], k/v → the KV-cache rows pos0+m of this layer (infer.rs:31-33)
struct EpiQKVPrefill {
  static constexpr bool kDual = false;
  float* q;                 // [M][Dq]
  float* key_cache;         // this layer's [T][Dq]
  float* value_cache;
  const float* freq_real;   // [T][hs/2]
  const float* freq_imag;
  int pos0, Dq, hs2;
  __device__ __forceinline__ void operator()(int m, int n, const float (&v)[32], int group, int, bool valid) const {
    if (!valid) return;
    const int pos = pos0 + m;
    float* dst = (group == 0 ? q + (size_t)m * Dq : (group == 1 ? key_cache : value_cache) + (size_t)pos * Dq) + n;
    float o[32];
    if (group < 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int pair = (n >> 1) + j;                 // pair index inside the row
        const int f = pos * hs2 + pair % hs2;
        const bool ok = n + 2 * j < Dq;
        const float c = ok ? freq_real[f] : 0.f, s = ok ? freq_imag[f] : 0.f;
        o[2 * j] = __fsub_rn(__fmul_rn(v[2 * j], c), __fmul_rn(v[2 * j + 1], s));
        o[2 * j + 1] = __fadd_rn(__fmul_rn(v[2 * j], s), __fmul_rn(v[2 * j + 1], c));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = v[j];
    }
    if (n + 32 <= Dq) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < Dq) dst[j] = o[j];
    }
  }
};

// [w1|w3] dual tile: hb = (h1 · (1/(1+exp(−h1)))) · h3   (cpu.rs:54-64)
struct EpiSwiGLUPrefill {
  static constexpr bool kDual = true;
  float* hb;   // [M][F]
  int F;
  __device__ __forceinline__ void operator()(int m, int n, const float (&h1)[32], const float (&h3)[32], bool valid) const {
    if (!valid) return;
    float* dst = hb + (size_t)m * F + n;
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = (h1[j] * (1.0f / (1.0f + expf(-h1[j])))) * h3[j];
    if (n + 32 <= F) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < F) dst[j] = o[j];
    }
  }
};

// ---- causal attention over the KV cache for a block of query positions -----------------------------------
// Query row m (position pos0+m) attends to cache rows 0..pos0+m of its head: exactly what
// multi_head_attention computes for that position (cpu.rs:23-52), with an online softmax.
// Grid (ceil(ceil(M/64)/2), heads), each CTA two query blocks (see below); 256 threads as 16×16: thread (ty,tx) owns score rows 4ty..4ty+3 × key columns
// tx+16j, and output rows 4ty..4ty+3 × head columns 4tx+64j.  f32 CUDA-core math: attention is ~1 % of the
// prefill flops, the GEMMs own the tensor cores.
constexpr int kPfBK = 64, kPfThreads = 256, kPfMaxHs = 128;
// query rows per thread RQ ∈ {4, 1} → 64 or 16 queries per block: the small block keeps the SMs busy when a rank
// holds few heads (tensor parallelism) or the prompt is short
__host__ __device__ constexpr int pf_bq(int rq) { return 16 * rq; }

struct PrefillAttnParams {
  const float* q;          // [M][Dq]
  const float* key_cache;  // this layer [T][Dq]
  const float* value_cache;
  float* out;              // [M][Dq]
  int M, pos0, Dq, hs;
};

__host__ __device__ inline size_t prefill_attn_smem_bytes(int hs, int rq) {
  return (size_t)(pf_bq(rq) * (hs + 4) + kPfBK * (hs + 4) + kPfBK * hs + pf_bq(rq) * (kPfBK + 4)) * sizeof(float);
}

template <int RQ>
__global__ void __launch_bounds__(kPfThreads) prefill_attn_kernel(const PrefillAttnParams p) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  constexpr int kPfBQ = pf_bq(RQ);
  extern __shared__ __align__(16) float pf_smem[];
  const int hs = p.hs, ldq = hs + 4;
  float* Qs = pf_smem;                 // [64][hs+4]
  float* Ks = Qs + kPfBQ * ldq;        // [64][hs+4]
  float* Vs = Ks + kPfBK * ldq;        // [64][hs]
  float* Ps = Vs + kPfBK * hs;         // [64][68]
  constexpr int ldp = kPfBK + 4;

  const int h = blockIdx.y;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const size_t col = (size_t)h * hs;
  const int hs4 = hs >> 2;
  const float div = sqrtf((float)hs);
  // Causal work grows with the query block index, so a CTA takes blocks x and nq−1−x: every CTA does the same
  // number of key-block iterations (nq+1) and the whole grid is one balanced wave.
  const int nq = (p.M + kPfBQ - 1) / kPfBQ;
  for (int pass = 0; pass < 2; ++pass) {
  const int qb = pass == 0 ? (int)blockIdx.x : nq - 1 - (int)blockIdx.x;
  if (pass == 1 && qb <= (int)blockIdx.x) break;
  const int q0 = qb * kPfBQ;
  __syncthreads();  // the previous pass is done with Qs

  for (int i = tid; i < kPfBQ * hs4; i += kPfThreads) {
    const int r = i / hs4, c = i - r * hs4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < p.M) v = reinterpret_cast<const float4*>(p.q + (size_t)(q0 + r) * p.Dq + col)[c];
    reinterpret_cast<float4*>(Qs + r * ldq)[c] = v;
  }

  float mrow[RQ], lrow[RQ], acc[RQ][8];
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  const int last_q = min(q0 + kPfBQ, p.M) - 1;
  const int n_keys = p.pos0 + last_q + 1;  // keys 0 .. position of the block's last query

  for (int k0 = 0; k0 < n_keys; k0 += kPfBK) {
    __syncthreads();  // previous iteration's readers of Ks/Vs/Ps are done (also orders the Qs fill)
    for (int i = tid; i < kPfBK * hs4; i += kPfThreads) {
      const int r = i / hs4, c = i - r * hs4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < n_keys) {
        kv = reinterpret_cast<const float4*>(p.key_cache + (size_t)(k0 + r) * p.Dq + col)[c];
        vv = reinterpret_cast<const float4*>(p.value_cache + (size_t)(k0 + r) * p.Dq + col)[c];
      }
      reinterpret_cast<float4*>(Ks + r * ldq)[c] = kv;
      reinterpret_cast<float4*>(Vs + r * hs)[c] = vv;
    }
    __syncthreads();

    // scores: 4 rows × 4 key columns per thread
    float s[RQ][4];
#pragma unroll
    for (int i = 0; i < RQ; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int d = 0; d < hs4; ++d) {
      float4 qv[RQ], kv[4];
#pragma unroll
      for (int i = 0; i < RQ; ++i) qv[i] = reinterpret_cast<const float4*>(Qs + (RQ * ty + i) * ldq)[d];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * ldq)[d];
#pragma unroll
      for (int i = 0; i < RQ; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = dot4(qv[i], kv[j], s[i][j]);
    }
    // scale, causal mask, online softmax (row statistics shared by the 16 tx lanes of a row group)
#pragma unroll
    for (int i = 0; i < RQ; ++i) {
      const int qpos = p.pos0 + q0 + RQ * ty + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = k0 + tx + 16 * j;
        s[i][j] = (t <= qpos) ? s[i][j] / div : -INFINITY;   // divide, as cpu.rs:41
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float mn = fmaxf(mrow[i], mx);            // finite from the first block on (key 0 ≤ every query)
      const float f = expf(mrow[i] - mn);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = expf(s[i][j] - mn);           // exp(-inf) = 0 for masked keys
        Ps[(RQ * ty + i) * ldp + tx + 16 * j] = e;
        ps += e;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      lrow[i] = lrow[i] * f + ps;
      mrow[i] = mn;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] *= f;
    }
    __syncthreads();
    // O += P · V : rows 4ty+i, head columns 4tx + 64·jj + (0..3)
    const int kmax = min(kPfBK, n_keys - k0);
    for (int t = 0; t < kmax; ++t) {
      float pv[RQ];
#pragma unroll
      for (int i = 0; i < RQ; ++i) pv[i] = Ps[(RQ * ty + i) * ldp + t];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int c = 4 * tx + 64 * jj;
        if (c < hs) {
          const float4 vv = *reinterpret_cast<const float4*>(Vs + t * hs + c);
#pragma unroll
          for (int i = 0; i < RQ; ++i) {
            acc[i][4 * jj + 0] = fmaf(pv[i], vv.x, acc[i][4 * jj + 0]);
            acc[i][4 * jj + 1] = fmaf(pv[i], vv.y, acc[i][4 * jj + 1]);
            acc[i][4 * jj + 2] = fmaf(pv[i], vv.z, acc[i][4 * jj + 2]);
            acc[i][4 * jj + 3] = fmaf(pv[i], vv.w, acc[i][4 * jj + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    const int m = q0 + RQ * ty + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int c = 4 * tx + 64 * jj;
      if (c < hs)
        *reinterpret_cast<float4*>(p.out + (size_t)m * p.Dq + col + c) =
            make_float4(acc[i][4 * jj] / lrow[i], acc[i][4 * jj + 1] / lrow[i], acc[i][4 * jj + 2] / lrow[i],
                        acc[i][4 * jj + 3] / lrow[i]);
    }
  }
  }  // pass
}

// last prompt row → the decode path's residual buffer: x0 = x[M-1] + y[M-1]
static __global__ void prefill_last_row_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ x0,
                                        int D) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) x0[i] = x[i] + y[i];
}

}  // namespace rama
