// internal.cuh — declarations shared by the translation units of the library (host side of the C ABI,
// include/rama_b200.h): error plumbing, the run-time-resolved NCCL table, context / session structures, the GEMV
// dispatch templates and the few helpers that cross translation units.  Nothing here is part of the ABI.
#pragma once
#include "../../include/rama_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemv.cuh"
#include "misc_kernels.cuh"
#include "tp_exchange.cuh"

using namespace rama;

// ---- errors (ctx.cu) ----
int fail(int code, const char* fmt, ...);
int env_int(const char* name, int dflt);
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(RAMA_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define RK(call)                     \
  do {                               \
    int r_ = (call);                 \
    if (r_ != RAMA_OK) return r_;    \
  } while (0)
// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time (only when tp->world > 1) so that the single-GPU path has no
// dependency on it and the process shares whatever libnccl.so.2 is already loaded.
// ------------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
extern NcclApi g_nccl;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;
int nccl_load();
#define NK(call)                                                                                   \
  do {                                                                                             \
    int e_ = (call);                                                                               \
    if (e_ != 0) return fail(RAMA_E_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(e_)); \
  } while (0)
// ------------------------------------------------------------------------------------------------
// structures
// ------------------------------------------------------------------------------------------------
struct TensorPlan {  // local window [Lc][Rl][Cl] of the global tensor [Lc][R][C] at (r0, c0)
  size_t Lc = 1, R = 0, C = 0, r0 = 0, Rl = 0, c0 = 0, Cl = 0;
  size_t local_elems() const { return Lc * Rl * Cl; }
  size_t global_elems() const { return Lc * R * C; }
};

struct rama_ctx {
  int device = 0, sm_count = 148;
  int rank = 0, world = 1;
  NcclComm comm = nullptr;
  bool loaded = false;
  rama_config cfg{};
  int D = 0, F = 0, L = 0, H = 0, V = 0, T = 0, hs = 0;  // global
  int Dq = 0, Fl = 0, Hl = 0, Vl = 0, v0 = 0;            // this rank's shard
  TensorPlan plan[RAMA_T_COUNT];
  float* w[RAMA_T_COUNT] = {nullptr};
  const float* wcls = nullptr;  // this rank's classifier rows (own tensor, or a window of the embedding)
  cudaStream_t op_stream = nullptr;
  int use_pdl = 0;
  int tp_nowait = 0;  // RAMA_TP_NOWAIT=1: the GEMVs fed by the peer exchange skip griddepcontrol.wait (gemv.cuh use_pdl bit 3) — measured on
                      // one rank's share of a TP step (RAMA_TP_SIM): +1 % at P = 2/4, −4 % at P = 8 (the spinning consumers compete with the
                      // producer's tail for L2): off by default
  int variant_override = -1;
  int p2p = 0;  // TP exchange: 1 = fused peer-memory all-reduce (default), 0 = NCCL collectives
  int staged = 1;      // RAMA_GEMV_STAGED=0 disables the shared-memory-staged GEMV for small slabs
  int embed_kernel = 1;  // RAMA_EMBED_KERNEL=0: fold the embedding gather into the layer-0 QKV prologue (ProNorm::emb)
  int stage_max_kb = 110;  // RAMA_GEMV_STAGE_KB: largest x + slab the staged GEMV takes (≤ 110: two CTAs per SM; ≤ 208: one)
  int attn_cluster = 1;  // RAMA_ATTN=split selects the global-memory split merge (attn_decode_kernel) at every context length
  // RAMA_TP_SIM=P (measurement tool, single GPU, no `tp`): this device behaves like rank 0 of P — 1/P shards, the peer
  // exchange with every "peer" slot delivered locally (P remote stores become P local ones, the consumer still reduces P
  // LL partials).  Results are meaningless (the same partial P times); kernel shapes, bytes and the dependent chain are
  // those of one rank of a P-GPU run minus NVLink latency and rank skew: what ncu and the in-graph timeline can see.
  int tp_sim = 0;
  int tp_reduce = -1;  // RAMA_TP_REDUCE: how a norm prologue reduces the P peer partials — 0 every CTA reads all of them, 1 the CTAs
                       // of a cluster share the reads (tp_cluster), 2 two-phase through a local LL buffer; -1 (default): 1
  int tp_cluster = -1;  // RAMA_TP_CLUSTER: largest cluster size of the shared peer reduction in the norm prologues (0/1: off; -1: 8 at P = 8, else 4)
  int cluster_step = -1;  // RAMA_STEP=cluster / kernels: the layers of a tiny model as ONE cluster-scope kernel (step_kernel.cuh); -1: by size
  int cluster_step_ctas = 16;  // RAMA_STEP_CLUSTER: CTAs of that cluster (16 needs the non-portable cluster size attribute)
  int persistent = 0;  // RAMA_STEP=persistent: the decode step as one persistent cooperative kernel (step_kernel.cuh);
                       // default: one fused kernel per op group in a CUDA graph (measured faster)
  std::mutex mu;
  // A stream capture is invalidated by a device-wide synchronisation (cudaDeviceSynchronize, default-stream work,
  // cudaFree) issued by ANOTHER host thread of the same context — the server creates and drops sessions while other
  // request threads capture their step graphs.  Captures and those device-wide operations take this lock.
  std::mutex cap_mu;
  std::atomic<int> n_objects{0};  // live sessions + batches: their captured graphs hold the weight pointers, so no reload
  // Single-process tensor parallelism (rama_ctx_create_multi): a GROUP context owns one rank context per device
  // (`ranks`, each with `group` pointing back); every entry point that takes the group handle drives all ranks from
  // the calling thread.  Rank contexts of a group use the peer-memory exchange with directly addressable peer
  // pointers (cudaDeviceEnablePeerAccess) — no NCCL, no IPC.
  std::vector<rama_ctx*> ranks;
  rama_ctx* group = nullptr;
  struct GroupPool* pool = nullptr;  // group only: one host thread per rank (group.cu)
};

// A device allocation that every rank of the TP group can address (NVLink peer memory): CUDA IPC between the processes
// of a torchrun launch, plain peer access inside a single-process group.
struct PeerBlock {
  char* local = nullptr;
  char* base[kMaxPeers] = {nullptr};  // every rank's block as seen from this rank (base[rank] == local)
  size_t bytes = 0;
  bool ipc = false;                   // peers were opened with cudaIpcOpenMemHandle (must be closed)
};
int peer_block_alloc(rama_ctx* c, size_t bytes, PeerBlock* b);                 // local allocation, zero-filled (session.cu)
int peer_block_connect(rama_ctx* c, PeerBlock* b, cudaStream_t st);            // collective over the processes; group ranks: no-op
void peer_block_free(rama_ctx* c, PeerBlock* b, cudaStream_t st, float* scratch);  // collective (barrier before the owner frees)
void peer_blocks_connect_group(PeerBlock* const* blocks, int P);               // single process: cross-wire P blocks


// A batched step runs on the batch's stream, the per-session entry points on the session's own stream.  The batch
// re-records ONE event after every step it launches; each session of that step keeps a reference and makes its own
// stream wait on it before its next operation (waiting on a later record of the same event is merely conservative).
struct BatchFence {
  cudaEvent_t ev = nullptr;
  ~BatchFence() { if (ev) cudaEventDestroy(ev); }
};

struct rama_session {
  rama_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  float *x0 = nullptr, *x1 = nullptr, *xfinal = nullptr, *xb = nullptr, *xb2 = nullptr, *w2out = nullptr;
  float *hb = nullptr, *hb2 = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr;
  float *logits = nullptr;  // [V]; this rank's rows live at logits + v0
  float *key_cache = nullptr, *value_cache = nullptr;
  float* attn_ws = nullptr;
  unsigned int* tickets = nullptr;
  ArgPart* part = nullptr;      // [world * sm_count] greedy partials (single GPU / NCCL mode)
  unsigned* seq = nullptr;      // device step counter (epoch source of the fused TP exchange)
  unsigned long long* bar = nullptr;  // persistent step kernel: [0] grid-barrier arrivals, [1] steps completed
  bool persistent = false;
  bool cluster_step = false;    // the layers run as one cluster-scope kernel (tiny models on one GPU)
  int cls_grid = 0;             // CTAs of the classifier launch (slots that get written)
  // fused TP exchange: one peer-addressable block per session
  //   parts[P][SMs][2] (LL) | inbox[2][P][D] (LL) | logits[V] | x0[D] | bulk flags[3][P] | bulk done counter
  // (logits and x0 live here so that peers can store into them: classifier slices, last prefill row)
  PeerBlock blk;
  size_t off_parts = 0, off_inbox = 0, off_logits = 0, off_x0 = 0, off_flags = 0, off_done = 0;
  bool p2p = false;
  unsigned bulk_epoch = 0;      // host counter of bulk exchanges on this session (prefill); identical on every rank
  uint2* red_ll = nullptr;      // two-phase peer reduction: local [2 stages][D] LL elements (gemv.cuh ProNorm::red_ll)
  PeerBlock pf_blk;             // prefill exchange buffers (allocated with the prefill workspace): inbox[P][rpr][D] | xn[cap][D]
  size_t pf_off_xn = 0;
  int pf_rpr_max = 0;
  // single-process group session: one session per rank, driven together
  std::vector<rama_session*> ranks;
  rama_session* parent = nullptr;
  unsigned long long* sort_keys = nullptr;
  StepCtrl* ctrl = nullptr;     // device
  int32_t *d_prompt = nullptr, *d_out = nullptr;
  StepCtrl* h_ring = nullptr;   // pinned ring for host-driven (token,pos)
  int ring_i = 0;
  int32_t* h_ret = nullptr;     // pinned {next, error}
  // step graphs per attention grid bucket (chunks-per-head CTAs for the position range): [mode][bucket],
  // mode 0 = forward, 1 = chained greedy, 2 = chained sampled
  cudaGraphExec_t g[3][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  int attn_gy = 1;              // gridDim.y of the attention launch being enqueued
  int attn_bk = 0;              // its bucket (0: positions < 256)
  float* wo_part = nullptr;     // [H][D] per-head wo partials of the fused attention+wo kernel (small models), or null
  int attn_wo_mode = 0;         // 1: attn_wo_cluster_kernel, 2: attn_wo_kernel, 0: separate launches
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int keep_att = 0;
  int n_split = 1;
  int host_mode_set = 0;
  int launches = 0;
  bool logits_gathered = false;
  bool parts_valid = true;       // the classifier GEMV's per-CTA argmax partials describe the current logits
  // prefill workspace (allocated on first use): activations of one chunk of prompt rows
  int pf_cap = 0;                // rows per chunk
  int pf_min = 16;               // rama_generate: prompts of at least this many rows (BOS included) are prefilled
  float *pf_x = nullptr, *pf_xn = nullptr, *pf_q = nullptr, *pf_att = nullptr, *pf_y = nullptr, *pf_h = nullptr;
  int32_t* pf_tokens = nullptr;
  int32_t* h_tokens = nullptr;   // pinned staging of a prompt (a pageable source could make the async copy wait for the stream)
  std::shared_ptr<BatchFence> fence;  // set by rama_forward_batch: work of another stream this session must wait for
  bool async_pending = false;         // work enqueued on s->stream since it was last synchronised (a batch must wait for it)
};

// called first by every per-session entry point: order this session's stream after the batched step that touched it
inline cudaError_t session_enter(rama_session* s) {
  if (!s->fence) return cudaSuccess;
  cudaError_t e = cudaStreamWaitEvent(s->stream, s->fence->ev, 0);
  s->fence.reset();
  return e;
}

// ------------------------------------------------------------------------------------------------
// GEMV dispatch
// ------------------------------------------------------------------------------------------------
constexpr int kNumVariants = 8;
constexpr int kVariantStaged = 100;  // gemv_smem_kernel
struct Variant { int WK, RP, U; };
inline const Variant kVariants[kNumVariants] = {{16, 2, 2}, {8, 2, 4}, {4, 2, 4}, {1, 2, 4},
                                                {16, 4, 2}, {8, 4, 2}, {2, 2, 4}, {16, 1, 4}};
constexpr size_t kMaxDynSmem = 200 * 1024;

inline void set_cluster_share(ProNorm& p, int on) { p.cluster_share = on; }
inline void set_cluster_share(ProPlain&, int) {}

// want_cluster > 1 (tensor parallelism, prologues that reduce peer partials): launch as thread-block clusters so that the
// CTAs of a cluster share the reduction (ProNorm).  The cluster size is the largest power of two ≤ want_cluster whose
// clusters still cover (almost) every SM in ONE wave — a CTA of this kernel owns an SM, and clusters cannot span GPCs.
template <int WK, int RP, int U, class Pro, class Rows, class Epi>
inline cudaError_t launch_gemv_t(int grid, cudaStream_t st, int pdl, const Pro& pro_in, const Rows& rows,
                                 const Epi& epi, int K4, int n_pairs, int want_cluster = 0,
                                 unsigned long long* trace = nullptr) {
  auto kern = gemv_fused_kernel<WK, RP, U, Pro, Rows, Epi>;
  static std::atomic<unsigned long long> attr_done{0};
  const cudaError_t attr_err = ensure_dyn_smem((const void*)kern, (int)kMaxDynSmem, attr_done);
  if (attr_err != cudaSuccess) return attr_err;
  Pro pro = pro_in;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kGemvThreads);
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0, cs = 1;
  for (int c = std::min(want_cluster, kMaxClusterShare); c >= 2; c >>= 1) {
    const int g = grid / c * c;
    if (g < c) continue;
    cfg.gridDim = dim3(g);
    cfg.dynamicSmemBytes = gemv_smem_bytes(K4, n_pairs, g, WK);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n_active = 0;
    if (cudaOccupancyMaxActiveClusters(&n_active, kern, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
    // share of the SM-filling grid that must fit ONE wave of clusters (clusters cannot span GPCs, and a CTA of this kernel owns an
    // SM: 8-CTA clusters leave 128 of 148 CTAs).  Measured on one rank's share of a TP = 8 step (RAMA_TP_SIM, tok/s): clusters of
    // 2 (all 148 CTAs) 863, of 4 (136) 879–908, of 8 (128) 921 — the shared reduction saves more than the lost SMs cost
    static const double min_frac = env_int("RAMA_TP_CLUSTER_MINPCT", 80) / 100.0;
    if ((double)n_active * c >= min_frac * g) {
      cs = c;
      grid = std::min(g, n_active * c);
      na = 1;
      break;
    }
  }
  set_cluster_share(pro, cs > 1);
  const size_t smem = gemv_smem_bytes(K4, n_pairs, grid, WK);
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  cfg.gridDim = dim3(grid);
  cfg.dynamicSmemBytes = smem;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, pro, rows, epi, K4, n_pairs, pdl, trace);
}

// bytes of shared memory the staged variant needs for this launch (x + the largest CTA slab)
inline size_t gemv_stage_bytes(int K4, int n_pairs, int grid, int rows_per_pair) {
  const int maxp = (n_pairs + grid - 1) / grid;
  return (size_t)K4 * 16 + (size_t)maxp * rows_per_pair * K4 * 16;
}

template <class Pro, class Rows, class Epi>
inline cudaError_t launch_gemv_staged(int grid, cudaStream_t st, int pdl, const Pro& pro, const Rows& rows, const Epi& epi,
                                      int K4, int n_pairs, unsigned long long* trace = nullptr) {
  auto kern = gemv_smem_kernel<Pro, Rows, Epi>;
  static std::atomic<unsigned long long> attr_done{0};
  const cudaError_t attr_err = ensure_dyn_smem((const void*)kern, (int)kGemvSmemStageMaxSolo, attr_done);
  if (attr_err != cudaSuccess) return attr_err;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemvThreads);
  cfg.dynamicSmemBytes = gemv_stage_bytes(K4, n_pairs, grid, 2);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (pdl) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, pro, rows, epi, K4, n_pairs, pdl, trace);
}

template <class Pro, class Rows, class Epi>
inline cudaError_t launch_gemv(int variant, int grid, cudaStream_t st, int pdl, const Pro& pro,
                               const Rows& rows, const Epi& epi, int K4, int n_pairs, int want_cluster = 0,
                               unsigned long long* trace = nullptr) {
  // small slabs (the small models): whole slab staged in shared memory ahead of the dependency (gemv_smem_kernel)
  if (variant == kVariantStaged) return launch_gemv_staged(grid, st, pdl, pro, rows, epi, K4, n_pairs, trace);
  switch (variant) {
    case 0: return launch_gemv_t<16, 2, 2>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 1: return launch_gemv_t<8, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 2: return launch_gemv_t<4, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 3: return launch_gemv_t<1, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 4: return launch_gemv_t<16, 4, 2>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 5: return launch_gemv_t<8, 4, 2>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 6: return launch_gemv_t<2, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    case 7: return launch_gemv_t<16, 1, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs, want_cluster, trace);
    default: return cudaErrorInvalidValue;
  }
}

inline int pick_variant(const rama_ctx* c, int K4, int n_pairs = 0) {
  if (c->variant_override >= 0 && c->variant_override < kNumVariants) return c->variant_override;
  if (c->staged && n_pairs > 0 && gemv_stage_bytes(K4, n_pairs, c->sm_count, 2) <= (size_t)c->stage_max_kb * 1024 &&
      n_pairs >= c->sm_count)
    return kVariantStaged;
  if (K4 >= 1024) return 1;
  if (K4 >= 512) return 2;
  if (K4 >= 128) return 6;
  return 3;
}
inline int pick_grid(const rama_ctx* c, int variant, int n_pairs) {
  if (variant == kVariantStaged) return c->sm_count;
  const int rp = kVariants[variant].RP;
  return std::max(1, std::min(c->sm_count, (n_pairs + rp - 1) / rp));
}

template <class Tp>
inline cudaError_t dalloc(Tp** p, size_t n) {
  cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(Tp));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(Tp));
  return e;
}

constexpr int kRing = 64;

inline PeerOut peer_out(const rama_session* s, int stage, int layer) {  // stage 0 = wo, 1 = w2
  PeerOut po{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return po;
  po.P = c->world; po.seq = s->seq; po.L = c->L + 1; po.layer = layer;
  for (int r = 0; r < c->world; ++r)
    po.inbox[r] = reinterpret_cast<uint2*>(s->blk.base[r] + s->off_inbox) + ((size_t)stage * c->world + (c->tp_sim ? r : c->rank)) * c->D;
  return po;
}
inline PeerIn peer_in(const rama_session* s, int stage, int layer) {
  PeerIn pi{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return pi;
  pi.inbox = reinterpret_cast<const uint2*>(s->blk.local + s->off_inbox) + (size_t)stage * c->world * c->D;
  pi.seq = s->seq; pi.error = &s->ctrl->error;
  pi.P = c->world; pi.n = c->D; pi.L = c->L + 1; pi.layer = layer;
  return pi;
}
inline PeerOut peer_out_parts(const rama_session* s) {  // classifier partials, "layer" L
  PeerOut po{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return po;
  po.P = c->world; po.seq = s->seq; po.L = c->L + 1; po.layer = c->L;
  for (int r = 0; r < c->world; ++r)
    po.inbox[r] = reinterpret_cast<uint2*>(s->blk.base[r] + s->off_parts) + (size_t)(c->tp_sim ? r : c->rank) * c->sm_count * 2;
  return po;
}
inline PeerIn peer_in_parts(const rama_session* s) {
  PeerIn pi{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return pi;
  pi.inbox = reinterpret_cast<const uint2*>(s->blk.local + s->off_parts);
  pi.seq = s->seq; pi.error = &s->ctrl->error;
  pi.P = c->world; pi.n = c->sm_count; pi.L = c->L + 1; pi.layer = c->L;
  return pi;
}

// ------------------------------------------------------------------------------------------------
// prompt prefill (tensor cores): ≙ the prompt part of generate()'s loop, mod.rs:187-192
// ------------------------------------------------------------------------------------------------
// kernel launch with the programmatic-stream-serialization attribute (PDL): the next kernel's CTAs start while this one
// drains; every kernel of the batched step executes griddepcontrol.wait before it touches memory (batch.cuh)
template <class... P, class... A>
inline cudaError_t launch_k(bool pdl, void (*kern)(P...), dim3 grid, dim3 block, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (pdl) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}


inline int tp_cluster_size(const rama_ctx* c) { return c->tp_cluster >= 0 ? c->tp_cluster : (c->world >= 8 ? 8 : 4); }

// how this session's norm prologues reduce the peer partials (rama_ctx::tp_reduce)
inline int tp_reduce_mode(const rama_session* s) {
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return 0;
  if (c->tp_reduce >= 0) return c->tp_reduce;
  return 1;  // measured (RAMA_TP_SIM, tok/s at P = 8 / 4 / 2): mode 0 779 / 617 / 400, mode 1 866 / 642 / 406, mode 2 869 / 627 / 398
}
// returns the extra use_pdl bits of the launch (bit 3: no griddepcontrol.wait, the peer epochs carry the dependency)
inline int set_peer_reduce(const rama_session* s, ProNorm& pro, int stage, int pdl) {
  if (!s->p2p) return 0;
  if (tp_reduce_mode(s) == 2 && s->red_ll) { pro.red_ll = s->red_ll + (size_t)stage * s->ctx->D; pro.red_n = 64; }
  if (pdl && s->ctx->tp_nowait) { pro.x_after_peers = 1; return 8; }
  return 0;
}

// the classifier epilogue stores every logits slice into every rank's array (no all-gather afterwards)
inline bool logits_pushed(const rama_session* s) { return s->p2p && !s->persistent; }

// classifier epilogue of the decode path (and of prefill's last row): local slice + greedy partial; under the peer
// exchange the slice goes into every rank's logits array (EpiCls::lpeer)
inline EpiCls make_epi_cls(const rama_session* s) {
  const rama_ctx* c = s->ctx;
  EpiCls e{s->logits + c->v0, s->part + (size_t)c->rank * c->sm_count, c->Vl, c->v0, -INFINITY, -1, peer_out_parts(s), {nullptr}};
  if (logits_pushed(s))
    for (int r = 0; r < c->world; ++r) e.lpeer[r] = reinterpret_cast<float*>(s->blk.base[r] + s->off_logits);
  return e;
}

// bulk exchange (tp_exchange.cuh) over a session's block: flags / done counter live in the session block
inline TpPeers tp_peers(const rama_session* s) {
  TpPeers t{};
  const rama_ctx* c = s->ctx;
  t.P = c->world; t.me = c->rank;
  for (int r = 0; r < c->world; ++r) t.flags[r] = reinterpret_cast<unsigned*>(s->blk.base[r] + s->off_flags);
  return t;
}

// ---- single-process group (group.cu) ----
int group_run(rama_ctx* g, const std::function<int(int)>& fn);  // fn(rank index) on every rank's thread; first failure
void group_destroy(rama_ctx* g);
void group_adopt_config(rama_ctx* g);
inline bool is_group(const rama_ctx* c) { return !c->ranks.empty(); }
inline rama_ctx* rank0(rama_ctx* c) { return c->ranks.empty() ? c : c->ranks[0]; }

// ---- helpers that cross translation units ----
int init_parts(rama_session* s);                        // session.cu
int gather_logits(rama_session* s);                     // session.cu
int read_ret(rama_session* s, int32_t* next);           // session.cu
int prefill_run(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float ms_kind[RAMA_PK_COUNT],
                int32_t* n_launch);                     // prefill_api.cu
int prefill_alloc_ws(rama_session* s);                  // prefill_api.cu
int group_prefill_prepare(rama_session* g);             // prefill_api.cu
