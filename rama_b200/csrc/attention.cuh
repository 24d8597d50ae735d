// attention.cuh — split-K flash-decode attention over the device-resident KV cache.
//
// Replaces calculate_attention + update_xb (math.cu:72-114; one thread per head, and a grid
// bug at 7B — SURVEY App. C) ≙ Device::multi_head_attention (cpu.rs:23-52):
//     score_t = (q·k_t) / sqrt(hs);  p = softmax(score_0..pos);  xb = Σ_t p_t · v_t
//
// Grid (heads, splits): CTA (h, c) owns timesteps [c·TC, (c+1)·TC) of head h, keeps a running
// (max, sum, acc[hs]) per warp (online softmax), merges its 4 warps in shared memory and writes
// one partial to the workspace.  The last CTA of a head to finish (atomic ticket) merges the
// partials in split order (deterministic) and writes xb — no second launch.
// KV cache layout is the reference's [T][Dq] per layer (head h at column h·hs): each timestep's
// head row is hs·4 contiguous bytes (512 B at 7B) ⇒ one coalesced 128-bit load per lane.
#pragma once
#include "common.cuh"

namespace rama {

constexpr int kAttnThreads = 128;
constexpr int kAttnWarps = kAttnThreads / kWarp;
constexpr int kAttnChunk = 64;       // timesteps per CTA
constexpr int kAttnMaxHs = 128;

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
};

__global__ void __launch_bounds__(kAttnThreads) attn_decode_kernel(const AttnParams p, int use_pdl) {
  __shared__ float s_m[kAttnWarps], s_l[kAttnWarps];
  __shared__ __align__(16) float s_acc[kAttnWarps][kAttnMaxHs];
  __shared__ unsigned int s_ticket;

  if (use_pdl) { pdl_launch_dependents(); pdl_wait(); }

  const int pos = p.pos_override >= 0 ? p.pos_override : p.ctrl->pos;
  const int n = pos + 1;
  const int n_chunks = (n + kAttnChunk - 1) / kAttnChunk;
  const int h = blockIdx.x, chunk = blockIdx.y;
  if (chunk >= n_chunks) return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hs = p.hs, hs4 = hs >> 2;
  const bool active = lane < hs4;
  const float inv_div = sqrtf((float)hs);

  const float4 q4 = active ? reinterpret_cast<const float4*>(p.q + (size_t)h * hs)[lane]
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  const int t0 = chunk * kAttnChunk, t1 = min(n, t0 + kAttnChunk);
  const size_t col = (size_t)h * hs;

  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  // two timesteps per iteration per warp for memory-level parallelism
  for (int t = t0 + warp; t < t1; t += 2 * kAttnWarps) {
    const int tb = t + kAttnWarps;
    const bool hb = tb < t1;
    float4 ka = make_float4(0.f, 0.f, 0.f, 0.f), kb = ka, va = ka, vb = ka;
    if (active) {
      ka = reinterpret_cast<const float4*>(p.key_cache + (size_t)t * p.Dq + col)[lane];
      va = reinterpret_cast<const float4*>(p.value_cache + (size_t)t * p.Dq + col)[lane];
      if (hb) {
        kb = reinterpret_cast<const float4*>(p.key_cache + (size_t)tb * p.Dq + col)[lane];
        vb = reinterpret_cast<const float4*>(p.value_cache + (size_t)tb * p.Dq + col)[lane];
      }
    }
    float sa = warp_sum(dot4(q4, ka, 0.f)) / inv_div;   // divide, as cpu.rs:41
    float sb = warp_sum(dot4(q4, kb, 0.f)) / inv_div;
    if (p.att && lane == 0) {
      p.att[(size_t)h * p.T + t] = sa;                  // raw score; normalised by the merger
      if (hb) p.att[(size_t)h * p.T + tb] = sb;
    }
    {
      const float mn = fmaxf(m, sa);
      const float sc = expf(m - mn), pe = expf(sa - mn);
      l = l * sc + pe;
      acc.x = acc.x * sc + pe * va.x; acc.y = acc.y * sc + pe * va.y;
      acc.z = acc.z * sc + pe * va.z; acc.w = acc.w * sc + pe * va.w;
      m = mn;
    }
    if (hb) {
      const float mn = fmaxf(m, sb);
      const float sc = expf(m - mn), pe = expf(sb - mn);
      l = l * sc + pe;
      acc.x = acc.x * sc + pe * vb.x; acc.y = acc.y * sc + pe * vb.y;
      acc.z = acc.z * sc + pe * vb.z; acc.w = acc.w * sc + pe * vb.w;
      m = mn;
    }
  }

  // merge the 4 warps of this CTA (fixed order)
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (active) reinterpret_cast<float4*>(s_acc[warp])[lane] = acc;
  __syncthreads();
  float* wsp = p.ws + ((size_t)h * p.n_split + chunk) * (hs + 2);
  {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, s_m[w]);
    // a warp that saw no timestep has m = -inf, l = 0: exp(-inf - M) = 0 contributes nothing
    for (int i = threadIdx.x; i < hs; i += kAttnThreads) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) a += s_acc[w][i] * expf(s_m[w] - M);
      wsp[2 + i] = a;
    }
    if (threadIdx.x == 0) {
      float L = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) L += s_l[w] * expf(s_m[w] - M);
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
  for (int i = threadIdx.x; i < hs; i += kAttnThreads) {
    float a = 0.f;
    for (int c = 0; c < n_chunks; ++c)
      a += __ldcg(wh + (size_t)c * (hs + 2) + 2 + i) * expf(__ldcg(wh + (size_t)c * (hs + 2)) - M);
    p.out[(size_t)h * hs + i] = a / L;
  }
  if (p.att) {
    for (int t = threadIdx.x; t < n; t += kAttnThreads) {
      const float s = __ldcg(p.att + (size_t)h * p.T + t);
      p.att[(size_t)h * p.T + t] = expf(s - M) / L;
    }
  }
  if (threadIdx.x == 0) p.tickets[h] = 0u;  // ready for the next launch
}

}  // namespace rama
