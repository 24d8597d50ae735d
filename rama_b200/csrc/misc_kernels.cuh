// misc_kernels.cuh — step begin/finish (embedding gather, sampler), the element-wise Device ops
// kept 1:1 with the reference trait for parity tests, and the synthetic-weight generator.
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace rama {

// ---- step begin: x ← token_embedding_table[token] (infer.rs:13) ---------------------------------
static __global__ void __launch_bounds__(256) step_begin_kernel(StepCtrl* ctrl, unsigned* seq,
                                                         const float* __restrict__ emb,
                                                         float* __restrict__ x, int D, int vocab,
                                                         int use_pdl, unsigned long long* trace = nullptr) {
  const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;  // rama_step_timeline
  if (tr) trace[0] = globaltimer_ns();
  if (use_pdl) { pdl_launch_dependents(); pdl_wait(); }
  if (tr) trace[1] = trace[2] = trace[3] = globaltimer_ns();
  if (blockIdx.x == 0 && threadIdx.x == 0) *seq += 1u;  // step counter: epoch source of the TP exchange
  int token = ctrl->token;
  if (token < 0 || token >= vocab) {  // the reference would panic on the slice (infer.rs:13)
    if (threadIdx.x == 0 && blockIdx.x == 0) ctrl->error = 1;
    token = 0;
  }
  const float4* src = reinterpret_cast<const float4*>(emb + (size_t)token * D);
  float4* dst = reinterpret_cast<float4*>(x);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (D >> 2); i += gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ---- sampler --------------------------------------------------------------------------------------
// Device::sample (cpu.rs:155-179) + sample_top_q (infer.rs:55-85), one CTA of 1024 threads.
//
// temperature == 0: argmax, ties → highest index.  Otherwise: logits /= T (only if T < 1),
// softmax in place, candidates p > (1-topp)/(V-1), order by (p desc, index asc) — a total order
// equal to the reference's stable descending sort —, walk the cumulative sum sequentially in f32
// exactly like the reference until it exceeds topp, draw r = U·cum with the reference's constant
// U (ChaCha20 seeded with 100 on every call, cpu.rs:161-162 ⇒ SURVEY App. B), walk the cdf.
constexpr int kSampleThreads = 1024;
constexpr float kRefU = 0.2721174359321594f;  // f32 bits 0x3e8b52fa

struct SampleParams {
  float* logits;            // [V] (full vocabulary on this rank)
  const ArgPart* part;      // optional greedy partials from the classifier kernel(s)
  int n_part;               // number of partial slots (P * SMs under TP)
  int n_live;               // TP: classifier CTAs per rank that actually write a slot
  int V;
  StepCtrl* ctrl;
  const int32_t* prompt;    // chained mode
  int32_t* out_tokens;      // chained mode: out_tokens[pos] = next
  unsigned long long* keys; // scratch [2][V] composite sort keys (top-p path)
  float temperature, topp;  // used when !chained (host-driven rama_sample)
  int chained;              // 1 (only from rama_generate and the graphs it captures): device-resident loop — temperature/topp
                            // from ctrl, prompt forcing, token feedback and pos advance on the device.  A launch parameter, not
                            // device state: a later rama_sample / rama_prefill / rama_sample_batch on the same session must not
                            // inherit it (stale ctrl->pos would index out_tokens out of bounds)
  PeerIn pin;               // pin.P > 0: part is an LL array [P*n_per][2] written by every rank's classifier
};

__device__ __forceinline__ unsigned long long sample_key(float p, int idx) {
  // ascending order of this key == (p descending, idx ascending); p > 0 so the f32 bit
  // pattern is monotonic in p.
  return ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(p)) << 32) | (unsigned)idx;
}

// Bitonic sort of 64-bit composite keys by one CTA.  n_pow2 ≤ kSortSmem keys are sorted entirely in
// shared memory (the common case: a trained model leaves few candidates above the cutoff); larger
// candidate sets (flat distributions, e.g. random weights) run the wide strides through L2.
constexpr int kSortSmem = 4096;
static __device__ void bitonic_sort_cta(unsigned long long* a, int n_pow2, unsigned long long* sm) {
  if (n_pow2 <= kSortSmem) {
    for (int i = threadIdx.x; i < n_pow2; i += kSampleThreads) sm[i] = a[i];
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < n_pow2; i += kSampleThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long x = sm[i], y = sm[ixj];
            if ((x > y) == ((i & k) == 0)) { sm[i] = y; sm[ixj] = x; }
          }
        }
        __syncthreads();
      }
    for (int i = threadIdx.x; i < n_pow2; i += kSampleThreads) a[i] = sm[i];
    __syncthreads();
    return;
  }
  for (int k = 2; k <= n_pow2; k <<= 1) {
    int j = k >> 1;
    for (; j >= kSortSmem; j >>= 1) {  // wide strides: through global memory (L2 resident)
      for (int i = threadIdx.x; i < n_pow2; i += kSampleThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          if ((x > y) == ((i & k) == 0)) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
    // remaining strides stay inside aligned tiles of kSortSmem keys: finish each tile in smem
    for (int base = 0; base < n_pow2; base += kSortSmem) {
      for (int i = threadIdx.x; i < kSortSmem; i += kSampleThreads) sm[i] = a[base + i];
      __syncthreads();
      for (int jj = j; jj > 0; jj >>= 1) {
        for (int i = threadIdx.x; i < kSortSmem; i += kSampleThreads) {
          const int ixj = i ^ jj;
          if (ixj > i) {
            const unsigned long long x = sm[i], y = sm[ixj];
            if ((x > y) == (((base + i) & k) == 0)) { sm[i] = y; sm[ixj] = x; }
          }
        }
        __syncthreads();
      }
      for (int i = threadIdx.x; i < kSortSmem; i += kSampleThreads) a[base + i] = sm[i];
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void sample_body(const SampleParams& p) {
  __shared__ float red[2 * kWarp];
  __shared__ int s_next, s_count, s_done;
  __shared__ float s_cum;
  __shared__ unsigned long long s_sort[kSortSmem];  // 32 KB: sort tile / walk staging
  StepCtrl* ctrl = p.ctrl;
  const float temperature = p.chained ? ctrl->temperature : p.temperature;
  const float topp = p.chained ? ctrl->topp : p.topp;
  const int V = p.V;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  int* redi = reinterpret_cast<int*>(red + kWarp);

  if (temperature != 0.0f && p.pin.P > 0) {
    // the vocabulary slices of the other ranks are stored into this rank's logits by their classifier epilogues: every
    // CTA's LL partial (written after a system-scope fence) proves that CTA's slice has arrived
    const unsigned ep = p.pin.epoch();
    for (int i = threadIdx.x; i < p.n_part; i += kSampleThreads) {
      if ((i % p.pin.n) >= p.n_live) continue;
      ld_ll2_wait(p.pin.inbox + 2 * (size_t)i, ep, p.pin.error);
    }
    __threadfence_system();
    __syncthreads();
  }
  if (temperature == 0.0f) {
    float bv = -INFINITY;
    int bi = -1;
    if (p.pin.P > 0) {
      // peer-written {value,epoch},{index,epoch} pairs; slots of CTAs that do not exist keep epoch 0
      const unsigned ep = p.pin.epoch();
      for (int i = threadIdx.x; i < p.n_part; i += kSampleThreads) {
        if ((i % p.pin.n) >= p.n_live) continue;
        const uint4 e = ld_ll2_wait(p.pin.inbox + 2 * (size_t)i, ep, p.pin.error);
        if ((int)e.z >= 0) argmax_merge(bv, bi, __uint_as_float(e.x), (int)e.z);
      }
    } else if (p.part) {
      for (int i = threadIdx.x; i < p.n_part; i += kSampleThreads)
      {
        const float pv = __ldcg(&p.part[i].v);
        const int pi = __ldcg(&p.part[i].i);
        if (pi >= 0) argmax_merge(bv, bi, pv, pi);
      }
    } else {
      for (int i = threadIdx.x; i < V; i += kSampleThreads) argmax_merge(bv, bi, p.logits[i], i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_merge(bv, bi, ov, oi);
    }
    if (l == 0) { red[w] = bv; redi[w] = bi; }
    __syncthreads();
    if (w == 0) {
      bv = red[l]; bi = redi[l];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_merge(bv, bi, ov, oi);
      }
      if (l == 0) s_next = bi < 0 ? 0 : bi;
    }
    __syncthreads();
  } else {
    float* x = p.logits;
    // temperature scaling only below 1 (cpu.rs:170-172), then softmax_num (cpu.rs:187-192)
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += kSampleThreads) {
      float z = x[i];
      if (temperature < 1.0f) { z = z / temperature; x[i] = z; }
      mx = fmaxf(mx, z);
    }
    mx = warp_max(mx);
    if (l == 0) red[w] = mx;
    __syncthreads();
    mx = warp_max(red[l]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < V; i += kSampleThreads) {
      const float e = expf(x[i] - mx);
      x[i] = e;
      sum += e;
    }
    sum = block_sum<kSampleThreads>(sum, red);
    const float cutoff = (1.0f - topp) / (float)(V - 1);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // candidates (order irrelevant: the composite key is a total order)
    for (int i = threadIdx.x; i < V; i += kSampleThreads) {
      const float pr = x[i] / sum;
      x[i] = pr;
      if (pr > cutoff) {
        const int slot = atomicAdd(&s_count, 1);
        p.keys[slot] = sample_key(pr, i);
      }
    }
    __syncthreads();
    const int C = s_count;
    if (C == 0) {  // the reference underflows and panics here (infer.rs:66)
      if (threadIdx.x == 0) { ctrl->error = 2; s_next = 0; }
      __syncthreads();
    } else {
      int n2 = 1;
      while (n2 < C) n2 <<= 1;
      for (int i = C + threadIdx.x; i < n2; i += kSampleThreads) p.keys[i] = ~0ull;
      __syncthreads();
      bitonic_sort_cta(p.keys, n2, s_sort);
      // Sequential f32 walks exactly as infer.rs:63-84; all threads stage kSortSmem keys at a time
      // into shared memory, thread 0 walks them.
      // walk 1: cum += p_i until cum > topp  → last
      if (threadIdx.x == 0) { s_cum = 0.f; s_done = 0; s_count = C - 1; }  // s_count reused as `last`
      __syncthreads();
      for (int base = 0; base < C; base += kSortSmem) {
        const int m = min(kSortSmem, C - base);
        for (int i = threadIdx.x; i < m; i += kSampleThreads) s_sort[i] = p.keys[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
          float cum = s_cum;
          for (int i = 0; i < m; ++i) {
            cum = cum + __uint_as_float(0xFFFFFFFFu - (unsigned)(s_sort[i] >> 32));
            if (cum > topp) { s_count = base + i; s_done = 1; break; }
          }
          s_cum = cum;
        }
        __syncthreads();
        if (s_done) break;
      }
      const int last = s_count;
      const float r = kRefU * s_cum;
      // walk 2: first i < last with r < cdf, else last
      __syncthreads();
      if (threadIdx.x == 0) { s_cum = 0.f; s_done = 0; s_next = (int)(p.keys[last] & 0xFFFFFFFFu); }
      __syncthreads();
      for (int base = 0; base < last; base += kSortSmem) {
        const int m = min(kSortSmem, last - base);
        for (int i = threadIdx.x; i < m; i += kSampleThreads) s_sort[i] = p.keys[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
          float cdf = s_cum;
          for (int i = 0; i < m; ++i) {
            cdf = cdf + __uint_as_float(0xFFFFFFFFu - (unsigned)(s_sort[i] >> 32));
            if (r < cdf) { s_next = (int)(s_sort[i] & 0xFFFFFFFFu); s_done = 1; break; }
          }
          s_cum = cdf;
        }
        __syncthreads();
        if (s_done) break;
      }
      __syncthreads();
    }
  }

  if (threadIdx.x == 0) {
    int next = s_next;
    if (p.chained) {
      const int pos = ctrl->pos;
      if (pos < ctrl->n_prompt) next = p.prompt[pos];  // prompt forcing (mod.rs:190-191)
      p.out_tokens[pos] = next;
      ctrl->token = next;                              // token feedback stays on the device
      ctrl->pos = pos + 1;
    }
    ctrl->next = next;
  }
}

static __global__ void __launch_bounds__(kSampleThreads) sample_kernel(const SampleParams p, int use_pdl,
                                                                       unsigned long long* trace = nullptr) {
  const bool tr = trace != nullptr && threadIdx.x == 0;  // rama_step_timeline
  if (tr) trace[0] = globaltimer_ns();
  if (use_pdl) { pdl_launch_dependents(); pdl_wait(); }
  if (tr) trace[1] = trace[2] = globaltimer_ns();
  sample_body(p);
  if (tr) trace[3] = globaltimer_ns();
}

// Tensor parallelism, host reads of the logits (rama_logits_to_host): wait until every classifier CTA of every rank has
// delivered its slice (its LL partial carries this step's epoch), then the stream-ordered copy may read the array.
static __global__ void __launch_bounds__(kSampleThreads) peer_parts_wait_kernel(const PeerIn pin, int n_part, int n_live) {
  const unsigned ep = pin.epoch();
  for (int i = threadIdx.x; i < n_part; i += kSampleThreads) {
    if ((i % pin.n) >= n_live) continue;
    ld_ll2_wait(pin.inbox + 2 * (size_t)i, ep, pin.error);
  }
  __threadfence_system();
}

// one CTA per sequence of a batch (server path: concurrent requests, lib.rs:127-160)
static __global__ void __launch_bounds__(kSampleThreads) sample_batch_kernel(const SampleParams* __restrict__ ps) {
  const SampleParams p = ps[blockIdx.x];
  sample_body(p);
}

// ---- element-wise Device ops (trait parity; the fused step does not launch these) -----------------
static __global__ void op_array_add_kernel(float* t, const float* s, size_t n) {  // cpu.rs:16-21
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    t[i] = t[i] + s[i];
}
static __global__ void op_array_mult_kernel(float* t, const float* s, size_t n) {  // cpu.rs:59-64
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    t[i] = t[i] * s[i];
}
static __global__ void op_sinu_kernel(float* o, size_t n) {  // cpu.rs:54-57
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a = o[i];
    o[i] = a * (1.0f / (1.0f + expf(-a)));
  }
}
static __global__ void op_copy_kernel(float* t, const float* s, size_t n) {  // cpu.rs:66-72
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    t[i] = s[i];
}
// cpu.rs:99-117, one CTA.  o may alias x.
static __global__ void __launch_bounds__(1024) op_rmsnorm_kernel(float* o, const float* x, const float* w, int n) {
  __shared__ float red[kWarp];
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) ss = fmaf(x[i], x[i], ss);
  ss = block_sum<1024>(ss, red);
  const float v = 1.0f / sqrtf(ss / (float)n + 1e-5f);
  for (int i = threadIdx.x; i < n; i += 1024) o[i] = w[i] * (v * x[i]);
}
// cpu.rs:74-97: one head of q and k
static __global__ void op_apply_position_kernel(float* q, float* k, const float* pr, const float* pi, int hs2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hs2) return;
  const float c = pr[i], s = pi[i];
  const float q0 = q[2 * i], q1 = q[2 * i + 1];
  q[2 * i] = __fsub_rn(__fmul_rn(q0, c), __fmul_rn(q1, s));
  q[2 * i + 1] = __fadd_rn(__fmul_rn(q0, s), __fmul_rn(q1, c));
  const float k0 = k[2 * i], k1 = k[2 * i + 1];
  k[2 * i] = __fsub_rn(__fmul_rn(k0, c), __fmul_rn(k1, s));
  k[2 * i + 1] = __fadd_rn(__fmul_rn(k0, s), __fmul_rn(k1, c));
}
// cpu.rs:119-125, one CTA
static __global__ void __launch_bounds__(1024) op_softmax_kernel(float* x, int n) {
  __shared__ float red[kWarp];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 1024) mx = fmaxf(mx, x[i]);
  mx = warp_max(mx);
  if (l == 0) red[w] = mx;
  __syncthreads();
  mx = warp_max(red[l]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float e = expf(x[i] - mx);
    x[i] = e;
    sum += e;
  }
  sum = block_sum<1024>(sum, red);
  for (int i = threadIdx.x; i < n; i += 1024) x[i] = x[i] / sum;
}
// general o_cols > 1 matmul (device.rs:13) — not on the decode path (which only uses o_cols = 1);
// kept so the trait is complete.  One thread per output, serial k.
static __global__ void op_matmul_general_kernel(float* o, const float* a, const float* b, int width, int o_rows,
                                         int o_cols) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)o_rows * o_cols) return;
  const int r = (int)(idx / o_cols), c = (int)(idx % o_cols);
  float acc = 0.f;
  for (int k = 0; k < width; ++k) acc = fmaf(a[(size_t)r * width + k], b[(size_t)k * o_cols + c], acc);
  o[idx] = acc;
}

// ---- synthetic weights (bench/test data; recipe of rama_b200/checkpoint.py) -------------------------
struct ShardMap {  // local [L][Rl][Cl] window of a global [L][R][C] tensor at (r0, c0)
  unsigned long long R, C, r0, Rl, c0, Cl;
};
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned sum16(unsigned long long h) {
  return (unsigned)(h & 0xFFFF) + (unsigned)((h >> 16) & 0xFFFF) + (unsigned)((h >> 32) & 0xFFFF) +
         (unsigned)(h >> 48);
}
static __global__ void synth_fill_kernel(float* dst, unsigned long long n, unsigned long long key, ShardMap m,
                                  unsigned long long start, float scale, float offset) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long per = m.Rl * m.Cl;
    const unsigned long long l = i / per, rem = i - l * per;
    const unsigned long long r = rem / m.Cl, c = rem - r * m.Cl;
    const unsigned long long g = start + (l * m.R + m.r0 + r) * m.C + m.c0 + c;
    const unsigned S = sum16(splitmix64(key + 2 * g)) + sum16(splitmix64(key + 2 * g + 1));
    const float v = __fmul_rn((float)(2 * (int)S - 8 * 65535), scale);
    dst[i] = __fadd_rn(v, offset);
  }
}

}  // namespace rama
