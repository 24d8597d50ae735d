// prefill.cuh — prompt prefill: the M prompt tokens at positions [pos0, pos0+M) go through the layers
// together, so every weight matrix is read once per prompt instead of once per token and the
// contractions run on the tensor cores (gemm_tf32x3.cuh for the weight products, prefill_attn_mma_kernel
// below for Q·Kᵀ and P·V).
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
// tx+16j, and output rows 4ty..4ty+3 × head columns 4tx+64j.  f32 CUDA-core math: since round 2 the fallback of the
// tensor-core kernel below (head sizes outside its table, grids too small for it, RAMA_PREFILL_ATTN=cuda).
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

// ---- the same attention on the tensor cores ----------------------------------------------------------------
// mma.sync m16n8k8 tf32 with the 3xTF32 split of both operands (x = hi + lo, hi = the 19 bits the tensor core reads;
// lo·hi + hi·lo + hi·hi, small terms first), f32 accumulation: products are exact to ≈2^-21, the same arithmetic the
// prefill GEMMs use.  A CTA of NW × KS warps owns a block of BQ = 16·NW queries and walks the causal range in 64-key tiles;
// warp (wq, ks) owns queries 16wq..16wq+15 against keys [ks·64/KS, (ks+1)·64/KS) of every tile:
//   S = Q·Kᵀ   A = Q rows (shared memory, split on load), B = K rows (K[key][d] IS the "col" operand), 8 keys per n-tile
//   online softmax on the accumulator fragments (a thread holds 2 rows × 2 keys per n-tile; row statistics over the quad)
//   O += P·V   the S accumulators ARE the A operand: with the k index of the MMA mapped to keys (t → 8j+2t, t+4 → 8j+2t+1) the
//              accumulator fragment of n-tile j is exactly the A fragment of k-step j — no shuffle, no shared-memory round trip;
//              B = V[key][col] read with the same key mapping.
// KS > 1 (keys of a tile split over warps) is what puts two or more warps on every scheduler when the grid has only one CTA
// per SM (512 rows × 32 heads = 128 paired blocks): mma.sync chains are latency-bound with one warp per scheduler.  The KS
// partial (m, l, O) triples of a query row meet in shared memory after the last tile and are merged by the ks = 0 warp, share by share.
// K and V tiles arrive by cp.async (zero-filled beyond the causal range).  Single-buffered (NW < 4): K of tile i+1 lands under the
// softmax and P·V of tile i, V of tile i+1 under the next Q·Kᵀ; double-buffered (NW = 4): tile i+1 lands under the whole of tile i.  n-tiles / k-steps wholly above a warp's last query are skipped (the diagonal
// tile costs about half).  Row pitch HS+4 words: every fragment load of Q, K and V is bank-conflict free.
// Query blocks are paired (x, nq−1−x) like the CUDA-core kernel: one balanced wave.
__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xFFFFE000u;   // (x − hi is exact; its own 13 low bits are dropped)
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// G independent accumulator tiles against one A fragment, term by term (lo·hi of every tile, then hi·lo, then hi·hi): consecutive
// MMAs never depend on each other; a tile's three terms sit G instructions apart.  b(i, half) returns the f32 B element of tile i.
template <int G, typename BF>
__device__ __forceinline__ void mma_tf32x3_group(float (*d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], BF b) {
  uint32_t bh[G][2], bl[G][2];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    tf32_split(b(i, 0), bh[i][0], bl[i][0]);
    tf32_split(b(i, 1), bh[i][1], bl[i][1]);
  }
#pragma unroll
  for (int i = 0; i < G; ++i) mma_tf32(d[i], alo, bh[i][0], bh[i][1]);
#pragma unroll
  for (int i = 0; i < G; ++i) mma_tf32(d[i], ahi, bl[i][0], bl[i][1]);
#pragma unroll
  for (int i = 0; i < G; ++i) mma_tf32(d[i], ahi, bh[i][0], bh[i][1]);
}

// 2^x on the special-function unit (ex2.approx: relative error ≤ 2^-22; 2^-inf = +0)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kPfMmaBK = 64;
// 64-query blocks (NW = 4: at most one CTA per SM fits anyway) double-buffer the K / V tiles: the copy of tile i+1 runs under the
// whole of tile i and ONE barrier per tile is left, so the warps of a CTA drift apart and one warp's softmax overlaps another's MMAs.
// The smaller blocks keep single tiles (two CTAs per SM is what overlaps their phases).
__host__ __device__ constexpr int prefill_attn_mma_bufs(int nw) { return nw >= 4 ? 2 : 1; }
__host__ __device__ constexpr size_t prefill_attn_mma_smem_bytes(int hs, int nw) {
  return (size_t)(16 * nw + 2 * prefill_attn_mma_bufs(nw) * kPfMmaBK) * (hs + 4) * sizeof(float);
}

template <int HS, int NW, int KS>
__global__ void __launch_bounds__(32 * NW * KS) prefill_attn_mma_kernel(const PrefillAttnParams p) {
  pdl_launch_dependents(); pdl_wait();  // (no-ops unless launched with programmatic stream serialization)
  static_assert(HS % 16 == 0 && HS <= kPfMaxHs, "head size: a multiple of 16 (pairs of 8-wide MMA tiles)");
  constexpr int BQ = 16 * NW, BK = kPfMmaBK, LD = HS + 4, KT = HS / 8, HS4 = HS / 4, NTHR = 32 * NW * KS;
  constexpr int KW = BK / KS, NT = KW / 8;   // keys, key n-tiles of a tile per warp
  static_assert(NT >= 2 && BQ <= BK, "at least one pair of key tiles per warp; the merge buffers live in the K / V tiles");
  constexpr int GS = NT >= 8 ? 8 : NT, GO = KT % 8 == 0 ? 8 : (KT % 4 == 0 ? 4 : 2);   // independent MMA chains per group
  extern __shared__ __align__(16) float pf_smem[];
  float* Qs = pf_smem;           // [BQ][LD]
  constexpr int NBUF = prefill_attn_mma_bufs(NW);
  constexpr bool DB = NBUF == 2;
  float* Ks = Qs + BQ * LD;          // [NBUF][BK][LD]   (after the last tile: a ks > 0 share's O partial, [BQ][LD])
  float* Vs = Ks + NBUF * BK * LD;   // [NBUF][BK][LD]   (after the last tile: its (m, l) pairs, [BQ][2])

  const int h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wq = warp % NW, ksp = warp / NW, kof = ksp * KW;
  const size_t col = (size_t)h * HS;
  // softmax in base 2: exp(s/√hs − m) = 2^(s·c − m'), c = log2(e)/√hs — one multiply per score and one MUFU per exponential
  const float sc = 1.4426950408889634f / sqrtf((float)HS);
  const int nq = (p.M + BQ - 1) / BQ;

  for (int pass = 0; pass < 2; ++pass) {
    const int qb = pass == 0 ? (int)blockIdx.x : nq - 1 - (int)blockIdx.x;
    if (pass == 1 && qb <= (int)blockIdx.x) break;
    const int q0 = qb * BQ;
    const int n_keys = p.pos0 + min(q0 + BQ, p.M);   // keys 0 .. position of the block's last query
    const int nkb = (n_keys + BK - 1) / BK;
    auto issue_tile = [&](float* tile, const float* cache, int kb, bool commit = true) {
      if (kb < nkb) {
        const int k0 = kb * BK;
        for (int i = tid; i < BK * HS4; i += NTHR) {
          const int r = i / HS4, c = i - r * HS4;
          const bool valid = k0 + r < n_keys;
          cp_async16_zfill(tile + r * LD + 4 * c, cache + (size_t)(valid ? k0 + r : 0) * p.Dq + col + 4 * c, valid);
        }
      }
      if (commit) cp_async_commit();   // (also for an empty tile: the wait counts below assume one group per call)
    };
    auto issue_pair = [&](int kb) {   // double-buffered: K and V of tile kb as ONE group into buffer kb & 1
      issue_tile(Ks + (kb & 1) * BK * LD, p.key_cache, kb, false);
      issue_tile(Vs + (kb & 1) * BK * LD, p.value_cache, kb, true);
    };
    __syncthreads();  // the previous pass is done with Qs / Ks / Vs
    for (int i = tid; i < BQ * HS4; i += NTHR) {   // the query block rides in the first K tile's copy group (rows ≥ M: zeros)
      const int r = i / HS4, c = i - r * HS4;
      const bool valid = q0 + r < p.M;
      cp_async16_zfill(Qs + r * LD + 4 * c, p.q + (size_t)(valid ? q0 + r : 0) * p.Dq + col + 4 * c, valid);
    }
    if (DB) {
      issue_pair(0);
    } else {
      issue_tile(Ks, p.key_cache, 0);
      issue_tile(Vs, p.value_cache, 0);
    }

    float o[KT][4];
#pragma unroll
    for (int n = 0; n < KT; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int qpos0 = p.pos0 + q0 + 16 * wq + g, qpos1 = qpos0 + 8;   // this thread's two query positions
    const int wlast = p.pos0 + q0 + 16 * wq + 15;                     // the warp's last query position
    const float* qrow0 = Qs + (16 * wq + g) * LD + t;
    const float* qrow1 = qrow0 + 8 * LD;

    for (int kb = 0; kb < nkb; ++kb) {
      const int k0 = kb * BK + kof;   // first key of this warp's share of the tile
      const int nlive = wlast >= k0 ? min(NT, (wlast - k0) / 8 + 1) : 0;   // key n-tiles with any unmasked entry for this warp
      if (DB) {
        cp_async_wait<0>();   // this tile's K and V (the only group in flight)
        __syncthreads();      // … from every thread; and every warp has left tile kb−1, whose buffers the next copy overwrites
        issue_pair(kb + 1);
      } else {
        cp_async_wait<1>();   // K of this tile (V may still be in flight)
        __syncthreads();      // … from every thread (first tile: the query block too)
      }
      const float* Kt = Ks + (DB ? (kb & 1) * BK * LD : 0);
      const float* Vt = Vs + (DB ? (kb & 1) * BK * LD : 0);
      float s[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      if (nlive > 0) {
#pragma unroll 2
        for (int ks = 0; ks < KT; ++ks) {
          uint32_t ahi[4], alo[4];
          tf32_split(qrow0[8 * ks], ahi[0], alo[0]);
          tf32_split(qrow1[8 * ks], ahi[1], alo[1]);
          tf32_split(qrow0[8 * ks + 4], ahi[2], alo[2]);
          tf32_split(qrow1[8 * ks + 4], ahi[3], alo[3]);
          const float* kp = Kt + (kof + g) * LD + 8 * ks + t;
#pragma unroll
          for (int n0 = 0; n0 < NT; n0 += GS)   // (a group with any live tile is computed whole)
            if (n0 < nlive)
              mma_tf32x3_group<GS>(&s[n0], ahi, alo, [&](int i, int half) { return kp[8 * (n0 + i) * LD + 4 * half]; });
        }
      }
      if (!DB) {
        __syncthreads();   // every warp is done with Ks
        issue_tile(Ks, p.key_cache, kb + 1);
      }

      if (nlive > 0) {
        // scale (cpu.rs:41 divides by √hs; here folded into the base-2 constant), causal mask, online softmax on the fragments
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          if (n < nlive) {
            const int key = k0 + 8 * n + 2 * t;
            s[n][0] = key <= qpos0 ? s[n][0] * sc : -INFINITY;
            s[n][1] = key + 1 <= qpos0 ? s[n][1] * sc : -INFINITY;
            s[n][2] = key <= qpos1 ? s[n][2] * sc : -INFINITY;
            s[n][3] = key + 1 <= qpos1 ? s[n][3] * sc : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
            mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
          }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        // a row may have seen no key yet (KS > 1: the upper key shares of the first tile): keep its statistics at (−inf, 0)
        // and subtract 0 instead of −inf, so that every power below is 2^(−inf) = 0 rather than 2^NaN
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float sub0 = mn0 == -INFINITY ? 0.f : mn0, sub1 = mn1 == -INFINITY ? 0.f : mn1;
        const float f0 = fast_exp2(m0 - sub0), f1 = fast_exp2(m1 - sub1);
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          if (n < nlive) {
            s[n][0] = fast_exp2(s[n][0] - sub0); s[n][1] = fast_exp2(s[n][1] - sub0);   // 2^(−inf) = 0 for masked keys
            s[n][2] = fast_exp2(s[n][2] - sub1); s[n][3] = fast_exp2(s[n][3] - sub1);
            ps0 += s[n][0] + s[n][1];
            ps1 += s[n][2] + s[n][3];
          }
        }
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
        l0 = l0 * f0 + ps0; l1 = l1 * f1 + ps1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int n = 0; n < KT; ++n) { o[n][0] *= f0; o[n][1] *= f0; o[n][2] *= f1; o[n][3] *= f1; }
      }

      if (!DB) {
        cp_async_wait<1>();   // V of this tile (the next K may still be in flight)
        __syncthreads();
      }
      if (nlive > 0) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (j < nlive) {
            // A fragment of k-step j = accumulator fragment of n-tile j under the key mapping above: (row g, key 8j+2t), (row g+8, same),
            // (row g, key 8j+2t+1), (row g+8, same)
            uint32_t ahi[4], alo[4];
            tf32_split(s[j][0], ahi[0], alo[0]);
            tf32_split(s[j][2], ahi[1], alo[1]);
            tf32_split(s[j][1], ahi[2], alo[2]);
            tf32_split(s[j][3], ahi[3], alo[3]);
            const float* vp = Vt + (kof + 8 * j + 2 * t) * LD + g;
#pragma unroll
            for (int n0 = 0; n0 < KT; n0 += GO)
              mma_tf32x3_group<GO>(&o[n0], ahi, alo, [&](int i, int half) { return vp[half * LD + 8 * (n0 + i)]; });
          }
        }
      }
      if (!DB) {
        __syncthreads();   // every warp is done with Vs
        issue_tile(Vs, p.value_cache, kb + 1);
      }
    }
    cp_async_wait<0>();
    if (DB && KS > 1) __syncthreads();   // (single tiles: the loop's last barrier) every warp is done with the tiles the merge re-uses

    const int r0 = q0 + 16 * wq + g, r1 = r0 + 8;
    if (KS > 1) {
      // merge the key shares, one share per round: warp ks = k publishes (m, l, O) in the K / V tile space, the ks = 0 warp (finite m:
      // it holds key 0) rescales and adds — fixed order.  The barrier above / the loop's last barrier retired every reader of the tiles, and no
      // copy is in flight.
      float* Os = Ks;   // [BQ][LD]
      float* Ms = Vs;   // [BQ][2]
      const int row = 16 * wq + g;
      for (int k = 1; k < KS; ++k) {
        if (ksp == k) {
#pragma unroll
          for (int n = 0; n < KT; ++n) {
            *reinterpret_cast<float2*>(Os + row * LD + 8 * n + 2 * t) = make_float2(o[n][0], o[n][1]);
            *reinterpret_cast<float2*>(Os + (row + 8) * LD + 8 * n + 2 * t) = make_float2(o[n][2], o[n][3]);
          }
          if (t == 0) {
            Ms[2 * row] = m0; Ms[2 * row + 1] = l0;
            Ms[2 * (row + 8)] = m1; Ms[2 * (row + 8) + 1] = l1;
          }
        }
        __syncthreads();
        if (ksp == 0) {
          const float pm0 = Ms[2 * row], pl0 = Ms[2 * row + 1], pm1 = Ms[2 * (row + 8)], pl1 = Ms[2 * (row + 8) + 1];
          const float mn0 = fmaxf(m0, pm0), mn1 = fmaxf(m1, pm1);
          const float a0 = fast_exp2(m0 - mn0), b0 = fast_exp2(pm0 - mn0), a1 = fast_exp2(m1 - mn1), b1 = fast_exp2(pm1 - mn1);   // (pm = −inf → b = 0)
          l0 = l0 * a0 + pl0 * b0; l1 = l1 * a1 + pl1 * b1;
          m0 = mn0; m1 = mn1;
#pragma unroll
          for (int n = 0; n < KT; ++n) {
            const float2 u0 = *reinterpret_cast<const float2*>(Os + row * LD + 8 * n + 2 * t);
            const float2 u1 = *reinterpret_cast<const float2*>(Os + (row + 8) * LD + 8 * n + 2 * t);
            o[n][0] = o[n][0] * a0 + u0.x * b0; o[n][1] = o[n][1] * a0 + u0.y * b0;
            o[n][2] = o[n][2] * a1 + u1.x * b1; o[n][3] = o[n][3] * a1 + u1.y * b1;
          }
        }
        if (k + 1 < KS) __syncthreads();   // the next share overwrites the buffers
      }
    }
    if (ksp == 0) {
#pragma unroll
      for (int n = 0; n < KT; ++n) {
        if (r0 < p.M) *reinterpret_cast<float2*>(p.out + (size_t)r0 * p.Dq + col + 8 * n + 2 * t) = make_float2(o[n][0] / l0, o[n][1] / l0);
        if (r1 < p.M) *reinterpret_cast<float2*>(p.out + (size_t)r1 * p.Dq + col + 8 * n + 2 * t) = make_float2(o[n][2] / l1, o[n][3] / l1);
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
