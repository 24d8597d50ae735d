// gemv.cuh — the batch-1 f32 GEMV family (HBM-bound) with fused prologues / epilogues.
//
// Replaces, for the decode step, the reference's cuBLAS sgemm m=1 calls (gpu.rs:175-189; 8·L+1
// per token) plus the rmsnorm / apply_position / copy_from_slice / sinu / array_mult /
// array_add launches around them (math.cu:17-49,117-136): SURVEY.md §2.3.
//
// Shape of one launch: W is row-major [rows][K] (llama2.c layout), x has K floats.
//   * Work unit = a PAIR of rows (RoPE rotates output pairs (2i,2i+1); SwiGLU pairs w1 row j
//     with w3 row j), so every epilogue sees both members of its pair in one thread.
//   * One persistent CTA per SM (512 threads); CTA b owns a contiguous, balanced range of
//     pairs ⇒ each SM streams one contiguous slab of W (coalesced 128-bit ld.global.nc with
//     L1::no_allocate), ≤1 pair of imbalance across the grid.
//   * Inside the CTA the 16 warps form WR×WK: WK warps split K of a row, WR warps take different
//     pair groups.  Each warp keeps 2·RP·U independent 128-bit loads in flight.
//   * x (after the fused prologue: residual add + rmsnorm, or plain copy) lives in shared memory,
//     read as conflict-free LDS.128; partial sums of the WK k-slices meet in shared memory and are
//     combined in a fixed order (deterministic), then the fused epilogue runs.
#pragma once
#include "common.cuh"

namespace rama {

constexpr int kGemvThreads = 512;
constexpr int kGemvWarps = kGemvThreads / kWarp;
constexpr int kPdlPrefetchBytes = 192 * 1024;  // per CTA: ≈ what HBM delivers to one SM in ~4 µs
constexpr int kMaxClusterShare = 8;            // cluster size of the shared peer reduction (ProNorm)

struct ArgPart {  // greedy partial: best value and its global vocabulary index
  float v;
  int i;
};

// ---- tensor-parallel exchange fused into the GEMVs (one-shot all-reduce over peer memory) -------
// Row-parallel wo / w2 leave a D-float PARTIAL on every rank.  Instead of a separate collective,
// the producing GEMV's epilogue stores each of its outputs straight into EVERY rank's inbox slot
// [stage][src rank][D] (peer-mapped HBM over NVLink) as {value, epoch} 8-byte elements; the
// consuming GEMV's prologue (which already adds the pending residual) reads the P partials, spinning
// per element until the epoch of this (step, layer) shows, and sums them in rank order —
// bit-identical on every rank.  epoch = step_seq·L + layer + 1, where step_seq is a per-session
// device counter bumped by step_begin_kernel (all ranks run the same launch sequence).
// Slot reuse is safe without double buffering because the two stages (wo, w2) alternate and each
// consumer waits for all ranks before the next producer of that slot can run.
struct PeerOut {
  uint2* inbox[kMaxPeers];   // rank r's slot for (stage, src = me), peer-mapped
  const unsigned* seq;       // step counter (device)
  int P;                     // 0 ⇒ no exchange (single GPU / NCCL mode)
  int L, layer;
  unsigned seq_add;          // 1 inside the persistent step kernel (the counter is bumped at its end), else 0
  __device__ __forceinline__ unsigned epoch() const { return (*seq + seq_add) * (unsigned)L + (unsigned)layer + 1u; }
  // select with static indices (a runtime index would spill the table to local memory)
  __device__ __forceinline__ uint2* inbox_of(int r) const {
    uint2* f = inbox[0];
#pragma unroll
    for (int i = 1; i < kMaxPeers; ++i) f = (r == i) ? inbox[i] : f;
    return f;
  }
};
struct PeerIn {
  const uint2* inbox;        // local [P][n] LL elements of this stage
  const unsigned* seq;
  int32_t* error;            // StepCtrl.error
  int P, n;                  // n = elements per partial (row stride of inbox)
  int L, layer;
  unsigned seq_add;
  __device__ __forceinline__ unsigned epoch() const { return (*seq + seq_add) * (unsigned)L + (unsigned)layer + 1u; }
};

// ---- prologues: fill xs[0..K4) (float4) ---------------------------------------------------

// xs = x                                      (wo: attention output; w2: SwiGLU output)
struct ProPlain {
  const float* x;
  template <int N>
  struct Pre {};
  template <int N>
  __device__ __forceinline__ Pre<N> preload(int) const { return Pre<N>{}; }
  template <int N>
  __device__ __forceinline__ void operator()(float4* xs, int K4, float* red, const Pre<N>&) const {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int i = threadIdx.x; i < K4; i += kGemvThreads) xs[i] = __ldcg(x4 + i);  // may be another CTA's output of this launch
    (void)red;
  }
};

// v = xin (+ add);  [CTA 0: xout = v];  xs = w * (rsqrt_scale * v)   [CTA 0: xnorm = xs]
// ≙ array_add (cpu.rs:16-21) folded in front of rmsnorm (cpu.rs:99-117):
//   scale = 1/sqrt(Σv²/n + 1e-5);  o[i] = w[i] * (scale * v[i])
struct ProNorm {
  const float* xin;    // residual stream (D)
  const float* add;    // pending residual contribution (wo / w2 output) or nullptr
  float* xout;         // updated residual stream (ping-pong buffer ≠ xin)
  const float* w;      // norm weight (D)
  float* xnorm;        // optional: normalised vector out (final rmsnorm → RunState.x) or nullptr
  PeerIn pin;          // pin.P > 0: the pending contribution is the sum of P peer partials instead of `add`
  int n_add = 1;       // > 1: `add` holds n_add partial vectors (per-head wo partials of attn_wo_kernel), stride K floats,
  float* add_out = nullptr;  //      summed in index order; CTA 0 stores the sum here (RunState.xb2)
  // First kernel of a step (layer 0): the residual stream starts as the embedding row of ctrl->token (infer.rs:13) — the gather
  // is this prologue's load instead of a kernel of its own; CTA 0 also validates the token and bumps the step counter.
  const float* emb = nullptr;  // token_embedding_table, or null
  StepCtrl* ctrl = nullptr;
  unsigned* seq = nullptr;     // step counter: epoch source of the TP exchange
  int vocab = 0;
  int cluster_share = 0;       // launched as clusters: the CTAs of a cluster split the peer reduction (see operator())
  // Two-phase peer reduction (P ≥ 4): the first red_n CTAs each reduce ONE slice of the vector (P LL partials + the residual)
  // and publish it — again as LL elements, so no fence — in this local buffer; every CTA then reads the reduced vector:
  // 32 KB per CTA at D = 4096 instead of P · 32 KB (measured on one rank's share of a TP = 8 step, RAMA_TP_SIM: the plain
  // prologue spends 7.2 µs pulling 148 × 256 KB through L2, cluster pairs 5.1 µs).
  uint2* red_ll = nullptr;     // local [K] LL elements of this stage, or null
  int red_n = 0;
  // The kernel runs WITHOUT griddepcontrol.wait (use_pdl bit 3): the arrival of the peers' partials — this rank's own
  // included — is the dependency.  The residual stream is then read only AFTER an element of the partial has arrived: a CTA of
  // this rank's producing GEMV has stored it, so that kernel had passed ITS wait and everything older (the kernel that wrote
  // the residual stream) is complete and visible.
  int x_after_peers = 0;
  // The norm weights never depend on the previous kernel: the first N float4 per thread are loaded BEFORE
  // griddepcontrol.wait (one L2 round trip off the critical path of every norm-prologue kernel).
  template <int N>
  struct Pre { float4 g[N]; };
  template <int N>
  __device__ __forceinline__ Pre<N> preload(int K4) const {
    Pre<N> r;
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int i = threadIdx.x + j * kGemvThreads;
      r.g[j] = i < K4 ? __ldg(w4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return r;
  }
  template <int N>
  __device__ __forceinline__ void operator()(float4* xs, int K4, float* red, const Pre<N>& pre) const {
    const float* xsrc = xin;
    if (emb) {
      int token = ctrl->token;
      if (token < 0 || token >= vocab) {  // the reference would panic on the slice (infer.rs:13)
        if (threadIdx.x == 0 && blockIdx.x == 0) ctrl->error = 1;
        token = 0;
      }
      if (threadIdx.x == 0 && blockIdx.x == 0) *seq += 1u;
      xsrc = emb + (size_t)token * (size_t)(K4 * 4);
    }
    const float4* x4 = reinterpret_cast<const float4*>(xsrc);
    const float4* a4 = reinterpret_cast<const float4*>(add);
    const bool peers = pin.P > 0 && add != nullptr;
    const unsigned ep = peers ? pin.epoch() : 0u;
    // Tensor parallelism, launched as thread-block clusters (cluster_share): the CTAs of a cluster split the vector — CTA r
    // reads the P LL partials of ITS slice only (1/CS of the L2 reads and of the spinning; at P = 8 every CTA of a plain launch
    // pulls 256 KB through L2, 148 CTAs at once), adds them in rank order and stores the slice into every sibling's xs over
    // DSMEM; one cluster barrier later every CTA holds the whole vector and the CS partial sums of squares.
    const unsigned CS = (peers && cluster_share) ? cluster_nctarank() : 1u;
    const unsigned cr = CS > 1 ? cluster_ctarank() : 0u;
    const int lo = (int)((long long)K4 * cr / CS), hi = (int)((long long)K4 * (cr + 1) / CS);
    const bool first_cluster = blockIdx.x < CS;  // the cluster (or CTA) that also writes the RunState copies
    const bool cta0 = blockIdx.x == 0;
    uint32_t sib[kMaxClusterShare];
#pragma unroll
    for (int j = 0; j < kMaxClusterShare; ++j) sib[j] = (CS > 1 && j < (int)CS) ? dsmem_addr(xs, j) : 0u;
    float ss = 0.f;
    if (peers && red_ll) {
      const int NR = min((int)gridDim.x, red_n);
      if ((int)blockIdx.x < NR) {  // phase 1: reduce slice blockIdx.x
        const int s_lo = (int)((long long)K4 * blockIdx.x / NR), s_hi = (int)((long long)K4 * (blockIdx.x + 1) / NR);
        for (int i = s_lo + threadIdx.x; i < s_hi; i += kGemvThreads) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!x_after_peers) v = __ldcg(x4 + i);
          uint4 lo_[kMaxPeers], hi_[kMaxPeers];
#pragma unroll
          for (int r = 0; r < kMaxPeers; ++r) {
            if (r < pin.P) {
              const uint2* e = pin.inbox + (size_t)r * pin.n + 4 * i;
              lo_[r] = ld_ll2(e);
              hi_[r] = ld_ll2(e + 2);
            }
          }
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < kMaxPeers; ++r) {  // rank order: same association on every rank
            if (r < pin.P) {
              const uint2* e = pin.inbox + (size_t)r * pin.n + 4 * i;
              if ((lo_[r].y != ep || lo_[r].w != ep)) lo_[r] = ld_ll2_wait(e, ep, pin.error);
              if ((hi_[r].y != ep || hi_[r].w != ep)) hi_[r] = ld_ll2_wait(e + 2, ep, pin.error);
              a.x += __uint_as_float(lo_[r].x); a.y += __uint_as_float(lo_[r].z);
              a.z += __uint_as_float(hi_[r].x); a.w += __uint_as_float(hi_[r].z);
            }
          }
          if (x_after_peers) v = __ldcg(x4 + i);
          v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
          st_ll2(red_ll + 4 * i, __float_as_uint(v.x), __float_as_uint(v.y), ep);
          st_ll2(red_ll + 4 * i + 2, __float_as_uint(v.z), __float_as_uint(v.w), ep);
          reinterpret_cast<float4*>(const_cast<float*>(add))[i] = a;  // RunState.xb2 / xb
        }
      }
      for (int i = threadIdx.x; i < K4; i += kGemvThreads) {  // phase 2: the reduced vector (residual included)
        uint4 l = ld_ll2(red_ll + 4 * i), h = ld_ll2(red_ll + 4 * i + 2);
        if ((l.y != ep || l.w != ep)) l = ld_ll2_wait(red_ll + 4 * i, ep, pin.error);
        if ((h.y != ep || h.w != ep)) h = ld_ll2_wait(red_ll + 4 * i + 2, ep, pin.error);
        const float4 v = make_float4(__uint_as_float(l.x), __uint_as_float(l.z), __uint_as_float(h.x), __uint_as_float(h.z));
        xs[i] = v;
        ss = dot4(v, v, ss);
        if (cta0) reinterpret_cast<float4*>(xout)[i] = v;
      }
    } else
    for (int i = lo + threadIdx.x; i < hi; i += kGemvThreads) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!(peers && x_after_peers)) v = __ldcg(x4 + i);
      if (peers) {
        // issue every rank's two 16-byte LL loads first (independent), then validate the epochs;
        // only an element that has not arrived yet falls into the spinning reload
        uint4 lo_[kMaxPeers], hi_[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
          if (r < pin.P) {
            const uint2* e = pin.inbox + (size_t)r * pin.n + 4 * i;
            lo_[r] = ld_ll2(e);
            hi_[r] = ld_ll2(e + 2);
          }
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {  // rank order: same association on every rank
          if (r < pin.P) {
            const uint2* e = pin.inbox + (size_t)r * pin.n + 4 * i;
            if ((lo_[r].y != ep || lo_[r].w != ep)) lo_[r] = ld_ll2_wait(e, ep, pin.error);
            if ((hi_[r].y != ep || hi_[r].w != ep)) hi_[r] = ld_ll2_wait(e + 2, ep, pin.error);
            a.x += __uint_as_float(lo_[r].x); a.y += __uint_as_float(lo_[r].z);
            a.z += __uint_as_float(hi_[r].x); a.w += __uint_as_float(hi_[r].z);
          }
        }
        if (x_after_peers) v = __ldcg(x4 + i);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        if (first_cluster) reinterpret_cast<float4*>(const_cast<float*>(add))[i] = a;  // RunState.xb2 / xb
      } else if (add) {
        float4 a = __ldcg(a4 + i);
        for (int h = 1; h < n_add; ++h) {
          const float4 t = __ldcg(a4 + (size_t)h * K4 + i);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        if (add_out && cta0) reinterpret_cast<float4*>(add_out)[i] = a;
      }
      if (CS > 1) {
#pragma unroll
        for (int j = 0; j < kMaxClusterShare; ++j)
          if (j < (int)CS) dsmem_st_f4(sib[j] + (uint32_t)i * 16u, v);  // own copy included
      } else {
        xs[i] = v;
      }
      ss = dot4(v, v, ss);
      if (first_cluster) reinterpret_cast<float4*>(xout)[i] = v;
    }
    ss = block_sum<kGemvThreads>(ss, red);
    if (CS > 1) {
      // red[40 .. 40+CS): the slices' sums of squares, written by each owner into every CTA of the cluster
      if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < kMaxClusterShare; ++j)
          if (j < (int)CS) dsmem_st_f1(dsmem_addr(red + 40 + cr, j), ss);
      }
      cluster_sync_all();  // slices and partial sums of every sibling have landed (release / acquire at cluster scope)
      ss = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxClusterShare; ++j)
        if (j < (int)CS) ss += red[40 + j];  // slice order: identical on every CTA and every rank
    }
    const float scale = 1.0f / sqrtf(ss / (float)(K4 * 4) + 1e-5f);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    auto apply = [&](int i, const float4 g) {  // CS == 1: same i as above, no sync needed; CS > 1: after the cluster barrier
      float4 v = xs[i];
      v.x = g.x * (scale * v.x); v.y = g.y * (scale * v.y);
      v.z = g.z * (scale * v.z); v.w = g.w * (scale * v.w);
      xs[i] = v;
      if (xnorm && cta0) reinterpret_cast<float4*>(xnorm)[i] = v;
    };
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int i = threadIdx.x + j * kGemvThreads;
      if (i < K4) apply(i, pre.g[j]);
    }
    for (int i = threadIdx.x + N * kGemvThreads; i < K4; i += kGemvThreads) apply(i, w4[i]);
  }
};

// ---- row addressing: pair p → two row pointers ----------------------------------------------

struct RowsPlain {  // rows 2p, 2p+1 of one matrix
  const float* w;
  int K, n_rows;
  static constexpr int kStreams = 1;
  // contiguous byte range that pairs [p0, p0+np) read from stream s (for the PDL L2 prefetch)
  __device__ __forceinline__ void slab(int p0, int np, int, const float*& ptr, size_t& bytes) const {
    ptr = w + (size_t)(2 * p0) * K;
    bytes = (size_t)min(2 * np, n_rows - 2 * p0) * K * 4;
  }
  __device__ __forceinline__ void operator()(int p, const float4*& r0, const float4*& r1) const {
    const size_t a = (size_t)(2 * p) * K;
    r0 = reinterpret_cast<const float4*>(w + a);
    r1 = (2 * p + 1 < n_rows) ? reinterpret_cast<const float4*>(w + a + K) : r0;
  }
  // shared-memory staging (gemv_smem_kernel): contiguous global ranges of pairs [p0, p0+np) and where they land
  static constexpr int kMaxRanges = 1;
  __device__ __forceinline__ int ranges(int p0, int np, const float* (&src)[3], uint32_t (&dst_off)[3], uint32_t (&bytes)[3]) const {
    src[0] = w + (size_t)(2 * p0) * K; dst_off[0] = 0;
    bytes[0] = (uint32_t)min(2 * np, n_rows - 2 * p0) * K * 4;
    return 1;
  }
  __device__ __forceinline__ void staged(int i, int np, const float* base, const float4*& r0, const float4*& r1, int p0) const {
    r0 = reinterpret_cast<const float4*>(base + (size_t)(2 * i) * K);
    r1 = (2 * (p0 + i) + 1 < n_rows) ? r0 + (K >> 2) : r0;
  }
};

struct RowsQKV {  // virtual matrix [wq; wk; wv] of this layer, each `rows_per` rows (even)
  const float* wq;
  const float* wk;
  const float* wv;
  int K, pairs_per;  // pairs per section = rows_per / 2
  static constexpr int kStreams = 1;
  __device__ __forceinline__ void slab(int p0, int np, int, const float*& ptr, size_t& bytes) const {
    const int sec = p0 / pairs_per, i = p0 - sec * pairs_per;
    const float* b = sec == 0 ? wq : (sec == 1 ? wk : wv);
    ptr = b + (size_t)(2 * i) * K;
    bytes = (size_t)(2 * min(np, pairs_per - i)) * K * 4;  // up to the end of this section
  }
  __device__ __forceinline__ void operator()(int p, const float4*& r0, const float4*& r1) const {
    const int sec = p / pairs_per, i = p - sec * pairs_per;
    const float* b = sec == 0 ? wq : (sec == 1 ? wk : wv);
    r0 = reinterpret_cast<const float4*>(b + (size_t)(2 * i) * K);
    r1 = r0 + (K >> 2);
  }
  static constexpr int kMaxRanges = 3;
  __device__ __forceinline__ int ranges(int p0, int np, const float* (&src)[3], uint32_t (&dst_off)[3], uint32_t (&bytes)[3]) const {
    int n = 0;
    for (int sec = 0; sec < 3; ++sec) {  // intersect [p0, p0+np) with the section's pairs
      const int a = max(p0, sec * pairs_per), b = min(p0 + np, (sec + 1) * pairs_per);
      if (a < b) {
        const float* base = sec == 0 ? wq : (sec == 1 ? wk : wv);
        src[n] = base + (size_t)(2 * (a - sec * pairs_per)) * K;
        dst_off[n] = (uint32_t)(2 * (a - p0)) * K * 4;
        bytes[n] = (uint32_t)(2 * (b - a)) * K * 4;
        ++n;
      }
    }
    return n;
  }
  __device__ __forceinline__ void staged(int i, int, const float* base, const float4*& r0, const float4*& r1, int) const {
    r0 = reinterpret_cast<const float4*>(base + (size_t)(2 * i) * K);
    r1 = r0 + (K >> 2);
  }
};

struct RowsW13 {  // pair p = (w1 row p, w3 row p)
  const float* w1;
  const float* w3;
  int K;
  static constexpr int kStreams = 2;
  __device__ __forceinline__ void slab(int p0, int np, int s, const float*& ptr, size_t& bytes) const {
    ptr = (s == 0 ? w1 : w3) + (size_t)p0 * K;
    bytes = (size_t)np * K * 4;
  }
  __device__ __forceinline__ void operator()(int p, const float4*& r0, const float4*& r1) const {
    r0 = reinterpret_cast<const float4*>(w1 + (size_t)p * K);
    r1 = reinterpret_cast<const float4*>(w3 + (size_t)p * K);
  }
  static constexpr int kMaxRanges = 2;
  __device__ __forceinline__ int ranges(int p0, int np, const float* (&src)[3], uint32_t (&dst_off)[3], uint32_t (&bytes)[3]) const {
    src[0] = w1 + (size_t)p0 * K; dst_off[0] = 0; bytes[0] = (uint32_t)np * K * 4;
    src[1] = w3 + (size_t)p0 * K; dst_off[1] = bytes[0]; bytes[1] = bytes[0];
    return 2;
  }
  __device__ __forceinline__ void staged(int i, int np, const float* base, const float4*& r0, const float4*& r1, int) const {
    r0 = reinterpret_cast<const float4*>(base + (size_t)i * K);
    r1 = reinterpret_cast<const float4*>(base + (size_t)(np + i) * K);
  }
};

// ---- epilogues: (pair index, dot0, dot1) ------------------------------------------------------

struct EpiStore {  // o[2p], o[2p+1]    (wo → xb2, w2 → residual contribution, op-level matmul)
  float* o;
  int n_rows;
  PeerOut po;      // po.P > 0: the outputs are a TP partial → stored into every rank's inbox instead
  __device__ __forceinline__ void prepare() {}
  __device__ __forceinline__ void prefetch(int) {}
  __device__ __forceinline__ void operator()(int p, float v0, float v1) {
    if (po.P > 0) {
      const unsigned ep = po.epoch();
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {  // static indexing keeps the pointer table in param space
        if (r < po.P) {  // the pair's two elements are adjacent and 16-byte aligned: one remote store instead of two
          if (2 * p + 1 < n_rows) st_ll2(po.inbox[r] + 2 * p, __float_as_uint(v0), __float_as_uint(v1), ep);
          else st_ll(po.inbox[r] + 2 * p, __float_as_uint(v0), ep);
        }
      }
      return;
    }
    o[2 * p] = v0;
    if (2 * p + 1 < n_rows) o[2 * p + 1] = v1;
  }
  __device__ __forceinline__ void finish(float*) {}
};

// RoPE (cpu.rs:74-97, simultaneous pair update, unfused mul/sub as in the reference) on q and k,
// then the KV-cache row write (infer.rs:31-33) — replaces H·L apply_position launches and
// 2·L copy_from_slice launches per token.
struct EpiQKV {
  float* q;
  float* k;
  float* v;
  float* key_cache;    // this layer's [T][Dq] block
  float* value_cache;
  const float* freq_real;  // [T][hs/2]
  const float* freq_imag;
  const StepCtrl* ctrl;
  int pairs_per, hs2, Dq;  // pairs per section, head_size/2, row stride of the cache
  int pos;                 // set by prepare()
  int pf_p = -1;           // pair whose RoPE factors this thread fetched ahead (prefetch())
  float pf_c = 0.f, pf_s = 0.f;
  // prepare(): every thread, right after griddepcontrol.wait — only ISSUES the load of the step's position (its first use is
  // in prefetch(), after the prologue has issued the activation loads, so it delays nothing).
  // prefetch(p): the thread that will run the epilogue of pair p fetches that pair's cos/sin before the dot products, so the
  // epilogue — the tail of the kernel — touches no global memory before its stores.
  __device__ __forceinline__ void prepare() { pos = ctrl->pos; }
  __device__ __forceinline__ void prefetch(int p) {
    const int sec = p / pairs_per;
    if (sec < 2) {
      const int j = (p - sec * pairs_per) % hs2;
      pf_c = freq_real[(size_t)pos * hs2 + j];
      pf_s = freq_imag[(size_t)pos * hs2 + j];
      pf_p = p;
    }
  }
  __device__ __forceinline__ void operator()(int p, float v0, float v1) {
    const int sec = p / pairs_per, i = p - sec * pairs_per;
    if (sec < 2) {
      const int j = i % hs2;
      const float c = p == pf_p ? pf_c : freq_real[(size_t)pos * hs2 + j];
      const float s = p == pf_p ? pf_s : freq_imag[(size_t)pos * hs2 + j];
      const float o0 = __fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, s));
      const float o1 = __fadd_rn(__fmul_rn(v0, s), __fmul_rn(v1, c));
      if (sec == 0) {
        reinterpret_cast<float2*>(q)[i] = make_float2(o0, o1);
      } else {
        reinterpret_cast<float2*>(k)[i] = make_float2(o0, o1);
        reinterpret_cast<float2*>(key_cache + (size_t)pos * Dq)[i] = make_float2(o0, o1);
      }
    } else {
      reinterpret_cast<float2*>(v)[i] = make_float2(v0, v1);
      reinterpret_cast<float2*>(value_cache + (size_t)pos * Dq)[i] = make_float2(v0, v1);
    }
  }
  __device__ __forceinline__ void finish(float*) {}
};

// SwiGLU (cpu.rs:54-64): hb = (h1 * (1/(1+exp(-h1)))) * h3 — replaces sinu + array_mult.
struct EpiSwiGLU {
  float* hb;
  float* hb2;
  __device__ __forceinline__ void prepare() {}
  __device__ __forceinline__ void prefetch(int) {}
  __device__ __forceinline__ void operator()(int p, float h1, float h3) {
    const float a = h1 * (1.0f / (1.0f + expf(-h1)));
    hb[p] = a * h3;
    hb2[p] = h3;
  }
  __device__ __forceinline__ void finish(float*) {}
};

// Classifier: logits + per-CTA greedy argmax partial (ties → higher index, cpu.rs:165-167).
struct EpiCls {
  float* logits;       // local logits [n_rows]
  ArgPart* part;       // [gridDim.x]: this CTA's best (value, global vocab index)
  int n_rows, row_offset;
  float bv;
  int bi;
  PeerOut po;          // po.P > 0: the partial goes to every rank's part array (po.inbox[r] = its slot base)
  // po.P > 0 and lpeer[0] != null: the logits slice itself is stored into EVERY rank's full logits array [V] (peer-mapped), so
  // the vocabulary all-gather is part of this epilogue; the LL partial written by finish() — after a system-scope fence — tells
  // a reader that this CTA's slice has arrived (sample_body / peer_parts_wait_kernel wait for all of them).
  float* lpeer[kMaxPeers];
  __device__ __forceinline__ void prepare() {}
  __device__ __forceinline__ void prefetch(int) {}
  __device__ __forceinline__ void operator()(int p, float v0, float v1) {
    const bool two = 2 * p + 1 < n_rows;
    if (po.P > 0 && lpeer[0]) {
      const int g = row_offset + 2 * p;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {
        if (r < po.P) {
          if (two && !(g & 1)) *reinterpret_cast<float2*>(lpeer[r] + g) = make_float2(v0, v1);
          else { lpeer[r][g] = v0; if (two) lpeer[r][g + 1] = v1; }
        }
      }
    } else {
      logits[2 * p] = v0;
      if (two) logits[2 * p + 1] = v1;
    }
    argmax_merge(bv, bi, v0, row_offset + 2 * p);
    if (two) argmax_merge(bv, bi, v1, row_offset + 2 * p + 1);
  }
  __device__ __forceinline__ void finish(float* red) {
    if (po.P > 0) __threadfence_system();  // this thread's remote logits stores before the CTA's partial (the arrival flag)
    // block argmax: warp shuffle then 16 warps through shared memory
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
      if (po.P > 0) {
        if (l < po.P) {  // {value, epoch}, {index, epoch} into rank l's array
          uint2* dst = po.inbox_of(l) + 2 * blockIdx.x;
          const unsigned ep = po.epoch();
          st_ll(dst, __float_as_uint(tv), ep);
          st_ll(dst + 1, (unsigned)ti, ep);
        }
      } else if (l == 0) {
        part[blockIdx.x].v = tv;
        part[blockIdx.x].i = ti;
      }
    }
  }
};

// ---- the streaming core ---------------------------------------------------------------------

// WK (warps that split K) is a run-time argument of the core so that the persistent step kernel can use ONE instantiation per
// phase type (gemv_pairs_rt); the per-variant kernels pass a constant through the force-inlined wrapper and get the same code
// as a template parameter would give.
template <int RP, int U, class Rows>
__device__ __forceinline__ void gemv_pairs_core(const Rows& rows, int K4, int p0, int np, const float4* __restrict__ xs,
                                                float* __restrict__ part, const int WK) {
  const int WR = kGemvWarps / WK;
  constexpr int R = 2 * RP;
  const int stride = kWarp * WK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = warp / WK, wk = warp - wr * WK;

  for (int base = wr * RP; base < np; base += WR * RP) {
    const float4* rp[R];
#pragma unroll
    for (int j = 0; j < RP; ++j) {
      const int p = min(base + j, np - 1);  // clamp: a duplicate is computed but never stored
      rows(p0 + p, rp[2 * j], rp[2 * j + 1]);
    }
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;

    int c = wk * kWarp + lane;
    for (; c + (U - 1) * stride < K4; c += U * stride) {
      float4 w[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) w[u][r] = ldg_stream(rp[r] + c + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 xv = xs[c + u * stride];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = dot4(w[u][r], xv, acc[r]);
      }
    }
    for (; c < K4; c += stride) {
      float4 w[R];
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = ldg_stream(rp[r] + c);
      const float4 xv = xs[c];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = dot4(w[r], xv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < RP; ++j) {
        if (base + j < np) {
          part[((base + j) * 2 + 0) * WK + wk] = acc[2 * j];
          part[((base + j) * 2 + 1) * WK + wk] = acc[2 * j + 1];
        }
      }
    }
  }
}

template <int WK, int RP, int U, class Rows>
__device__ __forceinline__ void gemv_pairs(const Rows& rows, int K4, int p0, int np,
                                           const float4* __restrict__ xs, float* __restrict__ part) {
  gemv_pairs_core<RP, U>(rows, K4, p0, np, xs, part, WK);
}

// Shared memory: float4 xs[K4] | float part[max_pairs_per_cta * 2 * WK]
__host__ __device__ inline size_t gemv_smem_bytes(int K4, int n_pairs, int grid, int WK) {
  const int maxp = (n_pairs + grid - 1) / grid;
  return (size_t)K4 * 16 + (size_t)maxp * 2 * WK * 4;
}

template <int WK, int RP, int U, class Pro, class Rows, class Epi>
__global__ void __launch_bounds__(kGemvThreads, 1)
gemv_fused_kernel(const Pro pro, const Rows rows, const Epi epi_in, int K4, int n_pairs, int use_pdl,
                  unsigned long long* trace) {
  // trace (rama_step_timeline, else null): CTA 0 stamps %globaltimer at entry, after griddepcontrol.wait (= the previous kernel
  // of the chain has completed), after the prologue (activations in shared memory; under TP: all peer partials arrived) and at its end
  const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tr) trace[0] = globaltimer_ns();
  extern __shared__ float4 gemv_smem[];
  __shared__ float red[2 * kWarp];
  float4* xs = gemv_smem;
  float* part = reinterpret_cast<float*>(xs + K4);

  const int per = n_pairs / gridDim.x, rem = n_pairs % gridDim.x;
  const int p0 = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int np = per + ((int)blockIdx.x < rem ? 1 : 0);

  // Weights never depend on the previous kernel: pull the head of this CTA's slab into L2 while
  // the previous kernel drains (PDL) and while the prologue below computes the norm, then wait
  // for the previous kernel's activations.
  // use_pdl: bit 0 = launched with programmatic stream serialization; bit 1 = release the dependent launch only
  // AFTER this kernel's own wait (then everything older than this kernel is complete when the dependent starts,
  // which lets the dependent read older results — the KV cache, the step control block — ahead of ITS wait).
  if ((use_pdl & 3) == 1) pdl_launch_dependents();
  const auto pre = pro.template preload<2>(K4);
  {
    if (np > 0 && threadIdx.x < Rows::kStreams) {
      const float* ptr;
      size_t bytes;
      rows.slab(p0, np, threadIdx.x, ptr, bytes);
      bytes = min(bytes, (size_t)kPdlPrefetchBytes / Rows::kStreams);
      if (bytes >= 16) l2_prefetch_bulk(ptr, (uint32_t)(bytes & ~(size_t)15));
    }
  }
  // bit 3 (tensor parallelism, prologues fed by the peer exchange): no griddepcontrol.wait — the epochs of the peers' partials
  // carry the dependency (ProNorm::x_after_peers), so the CTA reduces the partials as they arrive, under the producer's tail
  const bool no_wait = (use_pdl & 8) != 0;
  if ((use_pdl & 1) && !no_wait) pdl_wait();
  if ((use_pdl & 3) == 3 && !no_wait) pdl_launch_dependents();
  if (tr) trace[1] = globaltimer_ns();

  Epi epi = epi_in;
  epi.prepare();
  pro(xs, K4, red, pre);
  if ((use_pdl & 3) == 3 && no_wait) pdl_launch_dependents();  // everything older than this kernel is complete now
  if ((int)threadIdx.x < np) epi.prefetch(p0 + threadIdx.x);  // this thread's (first) epilogue pair
  __syncthreads();
  if (tr) trace[2] = globaltimer_ns();
  gemv_pairs<WK, RP, U>(rows, K4, p0, np, xs, part);
  __syncthreads();

  for (int i = threadIdx.x; i < np; i += kGemvThreads) {
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      v0 += part[(i * 2 + 0) * WK + k];
      v1 += part[(i * 2 + 1) * WK + k];
    }
    epi(p0 + i, v0, v1);
  }
  epi.finish(red);
  if (tr) trace[3] = globaltimer_ns();
}

// this CTA's balanced contiguous range of pairs
__device__ __forceinline__ void gemv_cta_range(int n_pairs, int& p0, int& np) {
  const int per = n_pairs / gridDim.x, rem = n_pairs % gridDim.x;
  p0 = blockIdx.x * per + min((int)blockIdx.x, rem);
  np = per + ((int)blockIdx.x < rem ? 1 : 0);
}

// ---- small slabs: stage the CTA's whole weight slab in shared memory BEFORE waiting for the previous kernel ---
// For the small models a CTA's share of a matrix is 16-90 KB and a launch is pure latency (dependency → x →
// norm → first weight byte → reduce → store).  Weights never depend on the previous kernel, so this variant
// issues cp.async.bulk copies of its whole slab into shared memory at entry, only then executes
// griddepcontrol.wait, and computes out of shared memory.  It is sized (≤ 64 registers, ≤ 110 KB of shared
// memory) so that TWO such CTAs fit an SM: with programmatic dependent launch the next kernel of the graph is
// resident and has its weights on chip while the current one is still finishing.
constexpr size_t kGemvSmemStageMax = 110 * 1024;       // two CTAs per SM below this
constexpr size_t kGemvSmemStageMaxSolo = 208 * 1024;   // one CTA per SM: the slab still lands ahead of the dependency

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <class Pro, class Rows, class Epi>
__global__ void __launch_bounds__(kGemvThreads, 2)
gemv_smem_kernel(const Pro pro, const Rows rows, const Epi epi_in, int K4, int n_pairs, int use_pdl,
                 unsigned long long* trace) {
  const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;  // (stamps as in gemv_fused_kernel)
  if (tr) trace[0] = globaltimer_ns();
  extern __shared__ float4 gemv_smem[];
  __shared__ float red[2 * kWarp];
  __shared__ __align__(8) unsigned long long bar_storage;
  float4* xs = gemv_smem;
  float* slab = reinterpret_cast<float*>(xs + K4);
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_storage);

  int p0, np;
  gemv_cta_range(n_pairs, p0, np);
  if ((use_pdl & 3) == 1) pdl_launch_dependents();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const float* src[3];
    uint32_t off[3], bytes[3], total = 0;
    const int nr = np > 0 ? rows.ranges(p0, np, src, off, bytes) : 0;
    for (int i = 0; i < nr; ++i) total += bytes[i];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(slab);
    for (int i = 0; i < nr; ++i) bulk_g2s(sbase + off[i], src[i], bytes[i], bar);
  }
  const auto pre = pro.template preload<1>(K4);
  if (use_pdl & 1) pdl_wait();
  if ((use_pdl & 3) == 3) pdl_launch_dependents();
  if (tr) trace[1] = globaltimer_ns();

  Epi epi = epi_in;
  epi.prepare();
  pro(xs, K4, red, pre);
  __syncthreads();  // xs complete; also makes the mbarrier init visible to every waiter
  if (tr) trace[2] = globaltimer_ns();
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < np; i += kGemvWarps) {
    const float4 *r0, *r1;
    rows.staged(i, np, slab, r0, r1, p0);
    if (lane == 0) epi.prefetch(p0 + i);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 2
    for (int c = lane; c < K4; c += kWarp) {
      const float4 xv = xs[c];
      a0 = dot4(r0[c], xv, a0);
      a1 = dot4(r1[c], xv, a1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) epi(p0 + i, a0, a1);
  }
  epi.finish(red);
  if (tr) trace[3] = globaltimer_ns();
}

// ---- the same streaming core with a run-time k-split (persistent step kernel: one instantiation per
// phase type instead of one per variant); RP = 2, U = 4 as in every default variant ------------------------
template <class Rows>
__device__ __forceinline__ void gemv_pairs_rt(const Rows& rows, int K4, int p0, int np, const float4* __restrict__ xs,
                                              float* __restrict__ part, int WK) {
  gemv_pairs_core<2, 4>(rows, K4, p0, np, xs, part, WK);
}

// L2 prefetch of the head of this CTA's slab of a coming phase (weights never depend on activations)
template <class Rows>
__device__ __forceinline__ void gemv_prefetch_slab(const Rows& rows, int n_pairs, size_t cap_bytes) {
  if (threadIdx.x < Rows::kStreams) {
    int p0, np;
    gemv_cta_range(n_pairs, p0, np);
    if (np > 0) {
      const float* ptr;
      size_t bytes;
      rows.slab(p0, np, threadIdx.x, ptr, bytes);
      bytes = min(bytes, cap_bytes / Rows::kStreams) & ~(size_t)15;
      const char* c = reinterpret_cast<const char*>(ptr);
      for (size_t off = 0; off < bytes; off += 65536)  // a bulk prefetch carries a 32-bit size; keep requests modest
        l2_prefetch_bulk(c + off, (uint32_t)min((size_t)65536, bytes - off));
    }
  }
}

// one GEMV phase of the persistent step kernel: prologue → stream → fused epilogue (all CTAs of the grid)
template <class Pro, class Rows, class Epi>
__device__ __forceinline__ void gemv_phase(const Pro& pro, const Rows& rows, const Epi& epi_in, int K4, int n_pairs,
                                           int WK, float4* xs, float* red) {
  float* part = reinterpret_cast<float*>(xs + K4);
  int p0, np;
  gemv_cta_range(n_pairs, p0, np);
  Epi epi = epi_in;
  epi.prepare();
  pro(xs, K4, red, pro.template preload<1>(K4));
  if ((int)threadIdx.x < np) epi.prefetch(p0 + threadIdx.x);
  __syncthreads();
  gemv_pairs_rt(rows, K4, p0, np, xs, part, WK);
  __syncthreads();
  for (int i = threadIdx.x; i < np; i += kGemvThreads) {
    float v0 = 0.f, v1 = 0.f;
    for (int k = 0; k < WK; ++k) {
      v0 += part[(i * 2 + 0) * WK + k];
      v1 += part[(i * 2 + 1) * WK + k];
    }
    epi(p0 + i, v0, v1);
  }
  epi.finish(red);
  __syncthreads();  // xs / part / red are reused by the next phase
}

}  // namespace rama
