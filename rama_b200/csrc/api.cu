// api.cu — host side of the C ABI (include/rama_b200.h): context (device + sharded weights),
// session (RunState in HBM, per-session stream, captured step graph), loaders, NCCL tensor
// parallelism and the op-level Device entry points.  No cuBLAS, no NVRTC, no CPU fallback.
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
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "attention.cuh"
#include "common.cuh"
#include "gemm_host.cuh"
#include "gemv.cuh"
#include "misc_kernels.cuh"
#include "prefill.cuh"
#include "step_kernel.cuh"
#include "batch.cuh"

using namespace rama;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
// used by the other translation units of the library (tokenizer.cpp)
extern "C" int rama_set_error(int code, const char* msg) { return fail(code, "%s", msg); }

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

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

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
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

static int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.h) return RAMA_OK;
  const char* names[] = {getenv("RAMA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(RAMA_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                               \
  *(void**)(&g_nccl.field) = dlsym(h, name);                           \
  if (!g_nccl.field) return fail(RAMA_E_NCCL, "libnccl lacks %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.h = h;
  return RAMA_OK;
}
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
  int variant_override = -1;
  int p2p = 0;  // TP exchange: 1 = fused peer-memory all-reduce (default), 0 = NCCL collectives
  int staged = 1;      // RAMA_GEMV_STAGED=0 disables the shared-memory-staged GEMV for small slabs
  int embed_kernel = 1;  // RAMA_EMBED_KERNEL=0: fold the embedding gather into the layer-0 QKV prologue (ProNorm::emb)
  int stage_max_kb = 110;  // RAMA_GEMV_STAGE_KB: largest x + slab the staged GEMV takes (≤ 110: two CTAs per SM; ≤ 208: one)
  int attn_cluster = 1;  // RAMA_ATTN=split selects the global-memory split merge (attn_decode_kernel) at every context length
  int persistent = 0;  // RAMA_STEP=persistent: the decode step as one persistent cooperative kernel (step_kernel.cuh);
                       // default: one fused kernel per op group in a CUDA graph (measured faster)
  std::mutex mu;
  // A stream capture is invalidated by a device-wide synchronisation (cudaDeviceSynchronize, default-stream work,
  // cudaFree) issued by ANOTHER host thread of the same context — the server creates and drops sessions while other
  // request threads capture their step graphs.  Captures and those device-wide operations take this lock.
  std::mutex cap_mu;
  std::atomic<int> n_objects{0};  // live sessions + batches: their captured graphs hold the weight pointers, so no reload
};

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
  int cls_grid = 0;             // CTAs of the classifier launch (slots that get written)
  // fused TP exchange: one IPC-exported block per session {flags[3][P] | parts[P][SMs] | inbox[2][P][D]}
  char* peer_mem = nullptr;
  char* peer_base[kMaxPeers] = {nullptr};  // every rank's block, peer-mapped (own block at [rank])
  size_t off_parts = 0, off_inbox = 0, peer_bytes = 0;
  bool p2p = false;
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
  std::shared_ptr<BatchFence> fence;  // set by rama_forward_batch: work of another stream this session must wait for
  bool async_pending = false;         // work enqueued on s->stream since it was last synchronised (a batch must wait for it)
};

// called first by every per-session entry point: order this session's stream after the batched step that touched it
static cudaError_t session_enter(rama_session* s) {
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
static const Variant kVariants[kNumVariants] = {{16, 2, 2}, {8, 2, 4}, {4, 2, 4}, {1, 2, 4},
                                                {16, 4, 2}, {8, 4, 2}, {2, 2, 4}, {16, 1, 4}};
constexpr size_t kMaxDynSmem = 200 * 1024;

template <int WK, int RP, int U, class Pro, class Rows, class Epi>
static cudaError_t launch_gemv_t(int grid, cudaStream_t st, int pdl, const Pro& pro, const Rows& rows,
                                 const Epi& epi, int K4, int n_pairs) {
  auto kern = gemv_fused_kernel<WK, RP, U, Pro, Rows, Epi>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const size_t smem = gemv_smem_bytes(K4, n_pairs, grid, WK);
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (pdl) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, pro, rows, epi, K4, n_pairs, pdl);
}

// bytes of shared memory the staged variant needs for this launch (x + the largest CTA slab)
static size_t gemv_stage_bytes(int K4, int n_pairs, int grid, int rows_per_pair) {
  const int maxp = (n_pairs + grid - 1) / grid;
  return (size_t)K4 * 16 + (size_t)maxp * rows_per_pair * K4 * 16;
}

template <class Pro, class Rows, class Epi>
static cudaError_t launch_gemv_staged(int grid, cudaStream_t st, int pdl, const Pro& pro, const Rows& rows, const Epi& epi,
                                      int K4, int n_pairs) {
  auto kern = gemv_smem_kernel<Pro, Rows, Epi>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemvSmemStageMaxSolo);
  });
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
  return cudaLaunchKernelEx(&cfg, kern, pro, rows, epi, K4, n_pairs, pdl);
}

template <class Pro, class Rows, class Epi>
static cudaError_t launch_gemv(int variant, int grid, cudaStream_t st, int pdl, const Pro& pro,
                               const Rows& rows, const Epi& epi, int K4, int n_pairs) {
  // small slabs (the small models): whole slab staged in shared memory ahead of the dependency (gemv_smem_kernel)
  if (variant == kVariantStaged) return launch_gemv_staged(grid, st, pdl, pro, rows, epi, K4, n_pairs);
  switch (variant) {
    case 0: return launch_gemv_t<16, 2, 2>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 1: return launch_gemv_t<8, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 2: return launch_gemv_t<4, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 3: return launch_gemv_t<1, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 4: return launch_gemv_t<16, 4, 2>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 5: return launch_gemv_t<8, 4, 2>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 6: return launch_gemv_t<2, 2, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    case 7: return launch_gemv_t<16, 1, 4>(grid, st, pdl, pro, rows, epi, K4, n_pairs);
    default: return cudaErrorInvalidValue;
  }
}

static int pick_variant(const rama_ctx* c, int K4, int n_pairs = 0) {
  if (c->variant_override >= 0 && c->variant_override < kNumVariants) return c->variant_override;
  if (c->staged && n_pairs > 0 && gemv_stage_bytes(K4, n_pairs, c->sm_count, 2) <= (size_t)c->stage_max_kb * 1024 &&
      n_pairs >= c->sm_count)
    return kVariantStaged;
  if (K4 >= 1024) return 1;
  if (K4 >= 512) return 2;
  if (K4 >= 128) return 6;
  return 3;
}
static int pick_grid(const rama_ctx* c, int variant, int n_pairs) {
  if (variant == kVariantStaged) return c->sm_count;
  const int rp = kVariants[variant].RP;
  return std::max(1, std::min(c->sm_count, (n_pairs + rp - 1) / rp));
}

// ------------------------------------------------------------------------------------------------
// misc API
// ------------------------------------------------------------------------------------------------
extern "C" int rama_abi_version(void) { return RAMA_ABI_VERSION; }
extern "C" const char* rama_last_error(void) { return g_err; }
extern "C" int rama_device_count(int* n) {
  if (!n) return fail(RAMA_E_INVALID, "n is NULL");
  CK(cudaGetDeviceCount(n));
  return RAMA_OK;
}
extern "C" int rama_tp_unique_id(uint8_t out[128]) {
  if (!out) return fail(RAMA_E_INVALID, "out is NULL");
  RK(nccl_load());
  NcclId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(out, &id, 128);
  return RAMA_OK;
}

extern "C" int rama_ctx_create(int device, const rama_tp* tp, rama_ctx** out) {
  if (!out) return fail(RAMA_E_INVALID, "out is NULL");
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(RAMA_E_CUDA, "device %d not present (%d CUDA devices)", device, n);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(RAMA_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  rama_ctx* c = new rama_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->use_pdl = env_int("RAMA_PDL", 1);
  c->variant_override = env_int("RAMA_GEMV_VARIANT", -1);
  c->staged = env_int("RAMA_GEMV_STAGED", 1);
  c->embed_kernel = env_int("RAMA_EMBED_KERNEL", 1);  // the fold measured +0.3 % (stories15M 10173 → 10201 tok/s): a tiny kernel in a PDL chain is almost free
  c->stage_max_kb = std::max(0, std::min((int)(kGemvSmemStageMaxSolo / 1024), env_int("RAMA_GEMV_STAGE_KB", 110)));
  {
    const char* m = getenv("RAMA_ATTN");
    c->attn_cluster = !(m && strcmp(m, "split") == 0);
  }
  {
    const char* m = getenv("RAMA_STEP");
    c->persistent = m && strcmp(m, "persistent") == 0;
  }
  {
    const char* m = getenv("RAMA_TP_COMM");
    c->p2p = !(m && strcmp(m, "nccl") == 0);
  }
  if (tp && tp->world > 1) {
    if (tp->rank < 0 || tp->rank >= tp->world) { delete c; return fail(RAMA_E_INVALID, "bad tp rank"); }
    int r = nccl_load();
    if (r != RAMA_OK) { delete c; return r; }
    NcclId id;
    memcpy(&id, tp->nccl_id, 128);
    int e = g_nccl.CommInitRank(&c->comm, tp->world, id, tp->rank);
    if (e != 0) { delete c; return fail(RAMA_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(e)); }
    c->rank = tp->rank;
    c->world = tp->world;
  }
  cudaError_t e = cudaStreamCreateWithFlags(&c->op_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return fail(RAMA_E_CUDA, "stream: %s", cudaGetErrorString(e)); }
  *out = c;
  return RAMA_OK;
}

static void free_weights(rama_ctx* c) {
  for (int i = 0; i < RAMA_T_COUNT; ++i) {
    if (c->w[i]) cudaFree(c->w[i]);
    c->w[i] = nullptr;
  }
  c->wcls = nullptr;
  c->loaded = false;
}

extern "C" int rama_ctx_destroy(rama_ctx* c) {
  if (!c) return RAMA_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_weights(c);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->op_stream) cudaStreamDestroy(c->op_stream);
  delete c;
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// weights: validation, shard plan, loaders
// ------------------------------------------------------------------------------------------------
static int set_config(rama_ctx* c, const rama_config* cfg) {
  const int P = c->world;
  if (cfg->dim <= 0 || cfg->hidden_dim <= 0 || cfg->n_layers <= 0 || cfg->n_heads <= 0 ||
      cfg->vocab_size <= 1 || cfg->seq_len <= 0)
    return fail(RAMA_E_INVALID, "non-positive dimension in config");
  if (cfg->dim % cfg->n_heads) return fail(RAMA_E_INVALID, "dim %% n_heads != 0");
  const int hs = cfg->dim / cfg->n_heads;
  // The reference's forward ignores n_kv_heads (K/V are dim wide, infer.rs:22-23,31-33) while its
  // cache is sized by it (ram.rs:8,20-21): only n_kv_heads == n_heads is consistent there.
  if (cfg->n_kv_heads != cfg->n_heads)
    return fail(RAMA_E_INVALID, "n_kv_heads (%d) != n_heads (%d): not supported by the reference forward",
                cfg->n_kv_heads, cfg->n_heads);
  if (hs % 4 || hs > kAttnMaxHs) return fail(RAMA_E_INVALID, "head_size %d must be a multiple of 4 and <= %d", hs, kAttnMaxHs);
  if (cfg->dim % 4 || cfg->hidden_dim % 4)
    return fail(RAMA_E_INVALID, "dim and hidden_dim must be multiples of 4 (reference matmul, cpu.rs:142)");
  if (cfg->n_heads % P || cfg->hidden_dim % P || cfg->vocab_size % P || (cfg->hidden_dim / P) % 4)
    return fail(RAMA_E_INVALID, "n_heads/hidden_dim/vocab_size not divisible by tp world %d", P);
  c->cfg = *cfg;
  c->D = cfg->dim; c->F = cfg->hidden_dim; c->L = cfg->n_layers; c->H = cfg->n_heads;
  c->V = cfg->vocab_size; c->T = cfg->seq_len; c->hs = hs;
  c->Hl = c->H / P; c->Dq = c->Hl * hs; c->Fl = c->F / P; c->Vl = c->V / P; c->v0 = c->rank * c->Vl;
  const size_t D = c->D, F = c->F, L = c->L, V = c->V, T = c->T;
  const size_t Dq = c->Dq, Fl = c->Fl, r = c->rank;
  auto full = [](size_t Lc, size_t R, size_t C) { TensorPlan p; p.Lc = Lc; p.R = R; p.C = C; p.Rl = R; p.Cl = C; return p; };
  auto rows = [](size_t Lc, size_t R, size_t C, size_t r0, size_t Rl) { TensorPlan p; p.Lc = Lc; p.R = R; p.C = C; p.r0 = r0; p.Rl = Rl; p.Cl = C; return p; };
  auto cols = [](size_t Lc, size_t R, size_t C, size_t c0, size_t Cl) { TensorPlan p; p.Lc = Lc; p.R = R; p.C = C; p.Rl = R; p.c0 = c0; p.Cl = Cl; return p; };
  c->plan[RAMA_T_TOKEN_EMBEDDING] = full(1, V, D);
  c->plan[RAMA_T_RMS_ATT] = full(1, L, D);
  c->plan[RAMA_T_WQ] = rows(L, D, D, r * Dq, Dq);   // column-parallel: this rank's heads
  c->plan[RAMA_T_WK] = rows(L, D, D, r * Dq, Dq);
  c->plan[RAMA_T_WV] = rows(L, D, D, r * Dq, Dq);
  c->plan[RAMA_T_WO] = cols(L, D, D, r * Dq, Dq);   // row-parallel: repacked to [D][Dq]
  c->plan[RAMA_T_RMS_FFN] = full(1, L, D);
  c->plan[RAMA_T_W1] = rows(L, F, D, r * Fl, Fl);
  c->plan[RAMA_T_W2] = cols(L, D, F, r * Fl, Fl);   // row-parallel: repacked to [D][Fl]
  c->plan[RAMA_T_W3] = rows(L, F, D, r * Fl, Fl);
  c->plan[RAMA_T_RMS_FINAL] = full(1, 1, D);
  c->plan[RAMA_T_FREQ_REAL] = full(1, T, hs / 2);
  c->plan[RAMA_T_FREQ_IMAG] = full(1, T, hs / 2);
  if (cfg->shared_weight) c->plan[RAMA_T_WCLS] = TensorPlan();
  else c->plan[RAMA_T_WCLS] = rows(1, V, D, (size_t)c->v0, (size_t)c->Vl);
  return RAMA_OK;
}

static int alloc_weights(rama_ctx* c) {
  free_weights(c);
  for (int i = 0; i < RAMA_T_COUNT; ++i) {
    const size_t n = c->plan[i].local_elems();
    if (!n) continue;
    cudaError_t e = cudaMalloc(&c->w[i], n * sizeof(float));
    if (e != cudaSuccess) {
      free_weights(c);
      return fail(RAMA_E_CUDA, "cudaMalloc of tensor %d (%zu floats): %s", i, n, cudaGetErrorString(e));
    }
  }
  c->wcls = c->cfg.shared_weight ? c->w[RAMA_T_TOKEN_EMBEDDING] + (size_t)c->v0 * c->D : c->w[RAMA_T_WCLS];
  return RAMA_OK;
}

// Sessions and batches hold captured graphs with the weight pointers baked in: loading again under them would leave
// dangling pointers in every replay.
static int reload_allowed(const rama_ctx* c) {
  const int n = c->n_objects.load();
  if (n > 0) return fail(RAMA_E_STATE, "%d session(s)/batch(es) of this context are alive: destroy them before loading weights again", n);
  return RAMA_OK;
}

// One pass host → HBM of this rank's window of tensor i (src = full global tensor on the host).
static int upload_tensor(rama_ctx* c, int i, const float* src, cudaStream_t st) {
  const TensorPlan& p = c->plan[i];
  if (!p.local_elems()) return RAMA_OK;
  if (!src) return fail(RAMA_E_INVALID, "tensor %d is NULL", i);
  for (size_t l = 0; l < p.Lc; ++l) {
    float* dst = c->w[i] + l * p.Rl * p.Cl;
    const float* s = src + (l * p.R + p.r0) * p.C + p.c0;
    if (p.Cl == p.C) CK(cudaMemcpyAsync(dst, s, p.Rl * p.C * sizeof(float), cudaMemcpyHostToDevice, st));
    else CK(cudaMemcpy2DAsync(dst, p.Cl * sizeof(float), s, p.C * sizeof(float), p.Cl * sizeof(float), p.Rl,
                              cudaMemcpyHostToDevice, st));
  }
  return RAMA_OK;
}

extern "C" int rama_ctx_load_host(rama_ctx* c, const rama_config* cfg, const float* const tensors[RAMA_T_COUNT]) {
  if (!c || !cfg || !tensors) return fail(RAMA_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  RK(reload_allowed(c));
  CK(cudaSetDevice(c->device));
  RK(set_config(c, cfg));
  RK(alloc_weights(c));
  for (int i = 0; i < RAMA_T_COUNT; ++i) {
    int r = upload_tensor(c, i, tensors[i], c->op_stream);
    if (r != RAMA_OK) { free_weights(c); return r; }
  }
  CK(cudaStreamSynchronize(c->op_stream));
  c->loaded = true;
  return RAMA_OK;
}

// ---- file → HBM pipeline --------------------------------------------------------------------------------------
// The reference reads the checkpoint one f32 at a time (read.rs:25-33: minutes at 7B) into Vecs and then uploads
// them.  Here reader threads pread() row blocks of this rank's window straight into a ring of pinned buffers while the
// calling thread issues the DMA of the blocks that are ready (1-D, or 2-D for the column windows of row-parallel
// wo / w2): disk/page-cache reads, and PCIe transfers overlap, nothing is staged in pageable memory, and under TP
// a rank only reads the rows it keeps.
struct LoadPiece {
  size_t file_off;     // first byte of the block in the file (full rows)
  size_t rows, row_bytes;         // rows in the block, bytes of a full file row
  size_t col_off, col_bytes;      // window inside a row
  char* dst;                      // device destination (pitch col_bytes)
};

static int load_file_pipelined(rama_ctx* c, int fd, const std::vector<LoadPiece>& pieces, double* gbps) {
  constexpr int kBuf = 8, kReaders = 4;
  constexpr size_t kBufBytes = (size_t)32 << 20;
  char* ring[kBuf] = {nullptr};
  cudaEvent_t done[kBuf];
  for (int i = 0; i < kBuf; ++i) {
    if (cudaHostAlloc((void**)&ring[i], kBufBytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
      for (int j = 0; j <= i; ++j) if (ring[j]) cudaFreeHost(ring[j]);
      return fail(RAMA_E_CUDA, "pinned staging ring: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  // Slot (i % kBuf) is used by pieces i, i+kBuf, i+2·kBuf, …  Two counters per slot enforce that order whatever the
  // scheduling of the reader threads (a reader that claimed piece i and was descheduled must not find its slot taken by
  // the reader of piece i+kBuf): issued[slot] = pieces of this slot whose DMA has been issued, filled[slot] = pieces of
  // this slot read from the file.  Piece i (the k-th use of its slot, k = i / kBuf) may be read only when issued == k
  // (and after that DMA has drained the buffer); the main loop issues it only when filled == k + 1.
  std::mutex mu;
  std::condition_variable cv;
  std::vector<size_t> issued(kBuf, 0), filled(kBuf, 0);
  std::atomic<size_t> next{0};
  std::atomic<int> io_error{0};
  const size_t n = pieces.size();
  const int reader_delay_us = env_int("RAMA_LOAD_TEST_DELAY_US", 0);  // test hook: widen the claim → fill window
  auto reader = [&](int t) {
    cudaSetDevice(c->device);
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= n || io_error.load()) return;
      const int slot = (int)(i % kBuf);
      const size_t k = i / kBuf;
      if (reader_delay_us > 0 && (i + t) % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(reader_delay_us));
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return issued[slot] == k || io_error.load(); });
        if (io_error.load()) return;
      }
      if (k > 0) cudaEventSynchronize(done[slot]);  // the DMA of piece i - kBuf has left the buffer
      const LoadPiece& p = pieces[i];
      size_t got = 0;
      const size_t want = p.rows * p.row_bytes;
      while (got < want) {
        const ssize_t r = pread(fd, ring[slot] + got, want - got, (off_t)(p.file_off + got));
        if (r <= 0) { io_error.store(1); break; }
        got += (size_t)r;
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        filled[slot] = k + 1;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < kReaders; ++t) th.emplace_back(reader, t);
  const auto t0 = std::chrono::steady_clock::now();
  size_t bytes = 0;
  int rc = RAMA_OK;
  for (size_t i = 0; i < n && rc == RAMA_OK; ++i) {
    const int slot = (int)(i % kBuf);
    const size_t k = i / kBuf;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return filled[slot] == k + 1 || io_error.load(); });
    }
    if (io_error.load()) { rc = fail(RAMA_E_IO, "short read from the checkpoint file"); break; }
    const LoadPiece& p = pieces[i];
    cudaError_t e;
    if (p.col_bytes == p.row_bytes)
      e = cudaMemcpyAsync(p.dst, ring[slot], p.rows * p.row_bytes, cudaMemcpyHostToDevice, c->op_stream);
    else
      e = cudaMemcpy2DAsync(p.dst, p.col_bytes, ring[slot] + p.col_off, p.row_bytes, p.col_bytes, p.rows,
                            cudaMemcpyHostToDevice, c->op_stream);
    if (e == cudaSuccess) e = cudaEventRecord(done[slot], c->op_stream);
    if (e != cudaSuccess) { rc = fail(RAMA_E_CUDA, "upload: %s", cudaGetErrorString(e)); io_error.store(1); }
    bytes += p.rows * p.col_bytes;
    {
      std::lock_guard<std::mutex> lk(mu);
      issued[slot] = k + 1;
    }
    cv.notify_all();
  }
  if (rc != RAMA_OK) { io_error.store(1); cv.notify_all(); }
  for (auto& t : th) t.join();
  if (rc == RAMA_OK && cudaStreamSynchronize(c->op_stream) != cudaSuccess) rc = fail(RAMA_E_CUDA, "sync after upload");
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (gbps) *gbps = sec > 0 ? bytes / sec / 1e9 : 0.0;
  for (int i = 0; i < kBuf; ++i) { cudaFreeHost(ring[i]); cudaEventDestroy(done[i]); }
  return rc;
}

static double g_last_load_gbps = 0.0;

extern "C" int rama_ctx_load_file(rama_ctx* c, const char* path) {
  if (!c || !path) return fail(RAMA_E_INVALID, "NULL argument");
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(RAMA_E_IO, "cannot open %s", path);
  struct stat st;
  int32_t h[7];
  if (fstat(fd, &st) != 0 || st.st_size < 28 || pread(fd, h, 28, 0) != 28) {
    close(fd);
    return fail(RAMA_E_IO, "%s: too short for a v0 header", path);
  }
  posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
  // header: 7 LE i32; vocab > 0 ⇒ shared classifier (mod.rs:140-166)
  if (h[5] == INT32_MIN) {  // |vocab| does not fit an i32 (untrusted file; -INT_MIN is undefined behaviour)
    close(fd);
    return fail(RAMA_E_INVALID, "%s: vocabulary size out of range", path);
  }
  rama_config cfg{h[0], h[1], h[2], h[3], h[4], h[5] > 0 ? h[5] : -h[5], h[6], h[5] > 0 ? 1 : 0};
  int rc;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    rc = reload_allowed(c);
    if (rc == RAMA_OK) rc = set_config(c, &cfg);
    if (rc == RAMA_OK) {
      size_t need = 28;
      for (int i = 0; i < RAMA_T_COUNT; ++i) need += c->plan[i].global_elems() * 4;
      if ((size_t)st.st_size < need) rc = fail(RAMA_E_IO, "%s: %lld bytes, config needs %zu", path, (long long)st.st_size, need);
    }
    if (rc == RAMA_OK) rc = alloc_weights(c);
    if (rc == RAMA_OK) {
      // row blocks of ≤ 32 MB (the staging buffer) of every (tensor, layer) window of this rank
      std::vector<LoadPiece> pieces;
      size_t base = 28;
      for (int i = 0; i < RAMA_T_COUNT; ++i) {
        const TensorPlan& p = c->plan[i];
        if (p.local_elems()) {
          const size_t piece_bytes = std::min<size_t>((size_t)32 << 20, (size_t)std::max(1, env_int("RAMA_LOAD_PIECE_KB", 32 << 10)) << 10);
          const size_t row_bytes = p.C * 4, max_rows = std::max<size_t>(1, piece_bytes / row_bytes);
          for (size_t l = 0; l < p.Lc; ++l) {
            for (size_t r = 0; r < p.Rl; r += max_rows) {
              const size_t nr = std::min(max_rows, p.Rl - r);
              pieces.push_back(LoadPiece{base + ((l * p.R + p.r0 + r) * p.C) * 4, nr, row_bytes, p.c0 * 4, p.Cl * 4,
                                         (char*)(c->w[i] + (l * p.Rl + r) * p.Cl)});
            }
          }
        }
        base += p.global_elems() * 4;
      }
      rc = load_file_pipelined(c, fd, pieces, &g_last_load_gbps);
      if (rc != RAMA_OK) free_weights(c); else c->loaded = true;
    }
  }
  close(fd);
  return rc;
}

// GB/s of this rank's window through the last rama_ctx_load_file (file → pinned ring → HBM), for tools/load_bench.py
extern "C" int rama_last_load_gbps(double* out) {
  if (!out) return fail(RAMA_E_INVALID, "NULL argument");
  *out = g_last_load_gbps;
  return RAMA_OK;
}

extern "C" int rama_ctx_load_synthetic(rama_ctx* c, const rama_config* cfg, uint64_t seed,
                                       const float scale[RAMA_T_COUNT], const float offset[RAMA_T_COUNT],
                                       const float* freq_real, const float* freq_imag) {
  if (!c || !cfg || !scale || !offset || !freq_real || !freq_imag) return fail(RAMA_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  RK(reload_allowed(c));
  CK(cudaSetDevice(c->device));
  RK(set_config(c, cfg));
  RK(alloc_weights(c));
  for (int i = 0; i < RAMA_T_COUNT; ++i) {
    const TensorPlan& p = c->plan[i];
    const size_t n = p.local_elems();
    if (!n) continue;
    if (i == RAMA_T_FREQ_REAL || i == RAMA_T_FREQ_IMAG) {
      CK(cudaMemcpyAsync(c->w[i], i == RAMA_T_FREQ_REAL ? freq_real : freq_imag, n * sizeof(float),
                         cudaMemcpyHostToDevice, c->op_stream));
      continue;
    }
    // key = splitmix64(seed ^ tensor_id * 0xD1B54A32D192ED03)  (rama_b200/checkpoint.py)
    unsigned long long z = seed ^ ((unsigned long long)i * 0xD1B54A32D192ED03ull);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    ShardMap m{p.R, p.C, p.r0, p.Rl, p.c0, p.Cl};
    synth_fill_kernel<<<c->sm_count * 8, 256, 0, c->op_stream>>>(c->w[i], n, z, m, 0ull, scale[i], offset[i]);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(c->op_stream));
  c->loaded = true;
  return RAMA_OK;
}

extern "C" int rama_ctx_config(const rama_ctx* c, rama_config* out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!c->loaded) return fail(RAMA_E_STATE, "no weights loaded");
  *out = c->cfg;
  return RAMA_OK;
}

extern "C" int rama_ctx_weight_to_host(rama_ctx* c, int tensor, float* dst, size_t n, size_t* n_out) {
  if (!c || tensor < 0 || tensor >= RAMA_T_COUNT) return fail(RAMA_E_INVALID, "bad argument");
  if (!c->loaded) return fail(RAMA_E_STATE, "no weights loaded");
  const size_t have = c->plan[tensor].local_elems();
  if (n_out) *n_out = have;
  if (!dst) return RAMA_OK;
  if (n < have) return fail(RAMA_E_INVALID, "buffer too small: %zu < %zu", n, have);
  CK(cudaSetDevice(c->device));
  if (have) CK(cudaMemcpy(dst, c->w[tensor], have * sizeof(float), cudaMemcpyDeviceToHost));
  return RAMA_OK;
}

extern "C" int rama_ctx_mem_info(rama_ctx* c, size_t* free_bytes, size_t* total_bytes) {
  if (!c || !free_bytes || !total_bytes) return fail(RAMA_E_INVALID, "NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemGetInfo(free_bytes, total_bytes));
  return RAMA_OK;
}

extern "C" int rama_ctx_weight_bytes(const rama_ctx* c, size_t* bytes) {
  if (!c || !bytes) return fail(RAMA_E_INVALID, "NULL argument");
  size_t t = 0;
  for (int i = 0; i < RAMA_T_COUNT; ++i) t += c->plan[i].local_elems() * 4;
  *bytes = t;
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
template <class Tp>
static cudaError_t dalloc(Tp** p, size_t n) {
  cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(Tp));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(Tp));
  return e;
}

static void session_free(rama_session* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (auto& gm : s->g) for (auto& g : gm) if (g) cudaGraphExecDestroy(g);
  if (s->p2p) {
    rama_ctx* c = s->ctx;
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && s->peer_base[r]) cudaIpcCloseMemHandle(s->peer_base[r]);
    // every rank must have unmapped this block before its owner frees it: barrier through NCCL
    if (c->comm && s->xb2) {
      g_nccl.AllReduce(s->xb2, s->xb2, 1, kNcclFloat32, kNcclSum, c->comm, s->stream);
      cudaStreamSynchronize(s->stream);
    }
    if (s->peer_mem) cudaFree(s->peer_mem);
  }
  void* bufs[] = {s->x0, s->x1, s->xfinal, s->xb, s->xb2, s->w2out, s->hb, s->hb2, s->q, s->k, s->v, s->att,
                  s->logits, s->key_cache, s->value_cache, s->attn_ws, s->tickets, s->part, s->sort_keys,
                  s->ctrl, s->d_prompt, s->d_out, s->seq, s->bar, s->wo_part, s->pf_x, s->pf_xn, s->pf_q, s->pf_att, s->pf_y, s->pf_h,
                  s->pf_tokens};
  for (void* b : bufs) if (b) cudaFree(b);
  if (s->h_ring) cudaFreeHost(s->h_ring);
  if (s->h_ret) cudaFreeHost(s->h_ret);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

constexpr int kRing = 64;

// Collective over the TP group: allocate this session's exchange block, swap CUDA IPC handles through
// NCCL and map every peer's block (NVLink P2P).  Layout (LL elements = {payload, epoch} uint2):
//   uint2 parts[P][SMs][2] | uint2 inbox[2 stages][P][D]
static int setup_peer_exchange(rama_session* s) {
  rama_ctx* c = s->ctx;
  const int P = c->world;
  if (P > kMaxPeers) return fail(RAMA_E_INVALID, "tp world %d > %d", P, kMaxPeers);
  s->off_parts = 0;
  s->off_inbox = ((size_t)P * c->sm_count * 2 * sizeof(uint2) + 255) / 256 * 256;
  s->peer_bytes = s->off_inbox + (size_t)2 * P * c->D * sizeof(uint2);
  CK(cudaMalloc((void**)&s->peer_mem, s->peer_bytes));
  CK(cudaMemset(s->peer_mem, 0, s->peer_bytes));  // epoch 0 is never used
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mine;
  CK(cudaIpcGetMemHandle(&mine, s->peer_mem));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  char* d_h = nullptr;
  CK(cudaMalloc((void**)&d_h, 64 * (size_t)P));
  CK(cudaMemcpy(d_h + 64 * (size_t)c->rank, &mine, 64, cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());  // s->stream is non-blocking: make the staged copy land first
  int e = g_nccl.AllGather(d_h + 64 * (size_t)c->rank, d_h, 16, kNcclFloat32, c->comm, s->stream);
  if (e) { cudaFree(d_h); return fail(RAMA_E_NCCL, "handle all-gather: %s", g_nccl.GetErrorString(e)); }
  CK(cudaStreamSynchronize(s->stream));
  std::vector<cudaIpcMemHandle_t> all(P);
  CK(cudaMemcpy(all.data(), d_h, 64 * (size_t)P, cudaMemcpyDeviceToHost));
  CK(cudaFree(d_h));
  for (int r = 0; r < P; ++r) {
    if (r == c->rank) { s->peer_base[r] = s->peer_mem; continue; }
    void* p = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess)
      return fail(RAMA_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (set RAMA_TP_COMM=nccl to fall back)", r,
                  cudaGetErrorString(ce));
    s->peer_base[r] = (char*)p;
  }
  // nobody may write into a peer block before its owner has zeroed it: barrier
  e = g_nccl.AllReduce(s->xb2, s->xb2, 1, kNcclFloat32, kNcclSum, c->comm, s->stream);
  if (e) return fail(RAMA_E_NCCL, "barrier: %s", g_nccl.GetErrorString(e));
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaMemsetAsync(s->xb2, 0, sizeof(float), s->stream));
  return RAMA_OK;
}

static PeerOut peer_out(const rama_session* s, int stage, int layer) {  // stage 0 = wo, 1 = w2
  PeerOut po{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return po;
  po.P = c->world; po.seq = s->seq; po.L = c->L + 1; po.layer = layer;
  for (int r = 0; r < c->world; ++r)
    po.inbox[r] = reinterpret_cast<uint2*>(s->peer_base[r] + s->off_inbox) + ((size_t)stage * c->world + c->rank) * c->D;
  return po;
}
static PeerIn peer_in(const rama_session* s, int stage, int layer) {
  PeerIn pi{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return pi;
  pi.inbox = reinterpret_cast<const uint2*>(s->peer_mem + s->off_inbox) + (size_t)stage * c->world * c->D;
  pi.seq = s->seq; pi.error = &s->ctrl->error;
  pi.P = c->world; pi.n = c->D; pi.L = c->L + 1; pi.layer = layer;
  return pi;
}
static PeerOut peer_out_parts(const rama_session* s) {  // classifier partials, "layer" L
  PeerOut po{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return po;
  po.P = c->world; po.seq = s->seq; po.L = c->L + 1; po.layer = c->L;
  for (int r = 0; r < c->world; ++r)
    po.inbox[r] = reinterpret_cast<uint2*>(s->peer_base[r] + s->off_parts) + (size_t)c->rank * c->sm_count * 2;
  return po;
}
static PeerIn peer_in_parts(const rama_session* s) {
  PeerIn pi{};
  const rama_ctx* c = s->ctx;
  if (!s->p2p) return pi;
  pi.inbox = reinterpret_cast<const uint2*>(s->peer_mem + s->off_parts);
  pi.seq = s->seq; pi.error = &s->ctrl->error;
  pi.P = c->world; pi.n = c->sm_count; pi.L = c->L + 1; pi.layer = c->L;
  return pi;
}

extern "C" int rama_session_create(rama_ctx* c, rama_session** out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!c->loaded) return fail(RAMA_E_STATE, "no weights loaded");
  CK(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> cap_lk(c->cap_mu);
  rama_session* s = new rama_session();
  s->ctx = c;
  s->n_split = (c->T + kAttnChunk - 1) / kAttnChunk;
  {
    const int var = pick_variant(c, c->D / 4);
    s->cls_grid = pick_grid(c, var, (c->Vl + 1) / 2);
  }
  const size_t D = c->D, Dq = c->Dq, Fl = c->Fl, V = c->V, T = c->T, L = c->L;
  size_t vp2 = 1;
  while (vp2 < V) vp2 <<= 1;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
#define A(call) if (e == cudaSuccess) e = (call)
  A(dalloc(&s->x0, D)); A(dalloc(&s->x1, D)); A(dalloc(&s->xfinal, D));
  A(dalloc(&s->xb, Dq)); A(dalloc(&s->xb2, D)); A(dalloc(&s->w2out, D));
  A(dalloc(&s->hb, Fl)); A(dalloc(&s->hb2, Fl));
  A(dalloc(&s->q, Dq)); A(dalloc(&s->k, Dq)); A(dalloc(&s->v, Dq));
  A(dalloc(&s->att, (size_t)c->Hl * T));
  A(dalloc(&s->logits, V));
  A(dalloc(&s->key_cache, L * T * Dq)); A(dalloc(&s->value_cache, L * T * Dq));
  A(dalloc(&s->attn_ws, (size_t)c->Hl * s->n_split * (c->hs + 2)));
  A(dalloc(&s->tickets, (size_t)c->Hl));
  s->p2p = c->world > 1 && c->p2p;
  s->persistent = c->persistent && (c->world == 1 || s->p2p);
  if (s->persistent) s->cls_grid = c->sm_count;  // every CTA of the persistent kernel writes a classifier partial
  A(dalloc(&s->part, (size_t)c->world * c->sm_count));
  A(dalloc(&s->seq, 1));
  A(dalloc(&s->bar, 2));
  // small models: attention + wo as one kernel with per-head partial outputs (attention.cuh attn_wo_kernel)
  // (measured: a win up to stories15M's size — 8425 → 9340 tok/s; at stories110M the redundant per-CTA attention and the
  // 12-way partial sum cost more than the saved launch — 3757 → 3536 — so dim ≤ 512 only)
  // Attention + wo as ONE launch (per-head partial outputs of wo, summed by the next prologue) — opt-in since the cluster
  // attention kernel: RAMA_ATTN_WO = 0 (default) separate launches, 1 = attn_wo_cluster_kernel, 2 = attn_wo_kernel (per-CTA
  // redundant attention); RAMA_ATTN_WO_MAXDIM moves the size limit of mode 1.  Measured on B200 (tok/s):
  //   stories15M   separate split-merge attention 8425 | per-CTA fused 9487 | cluster fused 10286 | cluster attention + wo 10342
  //   stories110M  3745 | 3536 | 4212 | 4577
  // — with the K/V fetch ahead of the dependency and the DSMEM merge the separate attention launch costs less than the
  // H-way partial sum in the next prologue, so the fused variants are kept only as measured alternatives.
  s->attn_wo_mode = env_int("RAMA_ATTN_WO", 0);
  if (s->attn_wo_mode == 1) {
    const int rpl = c->hs <= 64 ? 2 : 1;  // wo rows per 128-bit warp load
    if (!(c->world == 1 && c->D <= std::min(1024, env_int("RAMA_ATTN_WO_MAXDIM", 512)) && (size_t)c->H * c->D * sizeof(float) <= (size_t)64 * 1024 &&
          (c->D + kAttnClusterMax - 1) / kAttnClusterMax <= kAwcRowLoads * kAttnWarps * rpl))
      s->attn_wo_mode = 0;
  } else if (s->attn_wo_mode == 2) {
    if (!(c->world == 1 && c->D <= 512 && (size_t)c->H * c->D * sizeof(float) <= (size_t)64 * 1024 && c->H <= c->sm_count &&
          (c->D + c->sm_count / c->H - 1) / (c->sm_count / c->H) <= 8 * kAttnWoWarps))  // wo rows per CTA held in registers
      s->attn_wo_mode = 0;
  }
  if (s->attn_wo_mode) A(dalloc(&s->wo_part, (size_t)c->H * c->D));
  A(dalloc(&s->sort_keys, vp2));
  A(dalloc(&s->ctrl, 1));
  A(dalloc(&s->d_prompt, T)); A(dalloc(&s->d_out, T));
  A(cudaHostAlloc((void**)&s->h_ring, kRing * sizeof(StepCtrl), cudaHostAllocDefault));
  A(cudaHostAlloc((void**)&s->h_ret, 4 * sizeof(int32_t), cudaHostAllocDefault));
  A(cudaEventCreate(&s->ev0)); A(cudaEventCreate(&s->ev1));
  A(cudaDeviceSynchronize());  // the zero-fills above ran on the default stream; session streams are non-blocking
#undef A
  if (e != cudaSuccess) {
    session_free(s);
    return fail(RAMA_E_CUDA, "session allocation: %s", cudaGetErrorString(e));
  }
  if (s->p2p) {
    int rc = setup_peer_exchange(s);
    if (rc != RAMA_OK) {
      s->p2p = false;
      session_free(s);
      return rc;
    }
  }
  c->n_objects.fetch_add(1);
  *out = s;
  return RAMA_OK;
}

extern "C" int rama_session_destroy(rama_session* s) {
  if (!s) return RAMA_OK;
  rama_ctx* c = s->ctx;
  std::lock_guard<std::mutex> cap_lk(c->cap_mu);
  if (s->fence) { cudaSetDevice(c->device); cudaEventSynchronize(s->fence->ev); s->fence.reset(); }
  session_free(s);
  c->n_objects.fetch_sub(1);
  return RAMA_OK;
}

extern "C" int rama_session_sync(rama_session* s) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  CK(cudaSetDevice(s->ctx->device));
  CK(session_enter(s));
  CK(cudaStreamSynchronize(s->stream));
  s->async_pending = false;
  return RAMA_OK;
}

extern "C" int rama_session_reset(rama_session* s) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  rama_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  const size_t kv = (size_t)c->L * c->T * c->Dq * sizeof(float);
  CK(cudaMemsetAsync(s->key_cache, 0, kv, s->stream));
  CK(cudaMemsetAsync(s->value_cache, 0, kv, s->stream));
  CK(cudaMemsetAsync(s->ctrl, 0, sizeof(StepCtrl), s->stream));
  CK(cudaMemsetAsync(s->tickets, 0, c->Hl * sizeof(unsigned int), s->stream));
  s->host_mode_set = 0;
  CK(cudaStreamSynchronize(s->stream));
  return RAMA_OK;
}

extern "C" int rama_session_set_debug(rama_session* s, int keep_att) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  CK(cudaSetDevice(s->ctx->device));
  CK(session_enter(s));
  CK(cudaStreamSynchronize(s->stream));
  if (s->keep_att != keep_att) {  // the captured graphs bake the att pointer in
    for (auto& gm : s->g) for (auto& g : gm) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  }
  s->keep_att = keep_att;
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// the decode step
// ------------------------------------------------------------------------------------------------
struct StepTrace {  // optional per-kernel CUDA-event timing (rama_profile_step)
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
};

struct StepEnq {
  rama_session* s;
  cudaStream_t st;
  StepTrace* tr;
  int launches = 0;
  int pdl;
  cudaError_t err = cudaSuccess;
  int nccl_err = 0;
  void pre(int kind) {
    if (tr) {
      cudaEvent_t a;
      cudaEventCreate(&a);
      cudaEventRecord(a, st);
      tr->ev.push_back(a);
      tr->kind.push_back(kind);
    }
  }
  void post(cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
    if (err == cudaSuccess) { cudaError_t l = cudaGetLastError(); if (l != cudaSuccess) err = l; }
    ++launches;
    if (tr) {
      cudaEvent_t b;
      cudaEventCreate(&b);
      cudaEventRecord(b, st);
      tr->ev.push_back(b);
    }
  }
};

template <class F>
static void launch_plain(StepEnq& q, int kind, F&& f) {
  q.pre(kind);
  f();
  q.post(cudaSuccess);
}

static int wk_for(int K4) { return K4 >= 1024 ? 8 : (K4 >= 512 ? 4 : (K4 >= 128 ? 2 : 1)); }

// The step as ONE persistent cooperative launch (step_kernel.cuh); mode as enqueue_step.
static int enqueue_step_persistent(rama_session* s, cudaStream_t st, int mode, int* n_launch, long long* trace = nullptr) {
  rama_ctx* c = s->ctx;
  StepParams p{};
  p.trace = trace;
  p.D = c->D; p.Dq = c->Dq; p.Fl = c->Fl; p.L = c->L; p.V = c->V; p.Vl = c->Vl; p.v0 = c->v0; p.T = c->T; p.hs = c->hs; p.Hl = c->Hl;
  p.emb = c->w[RAMA_T_TOKEN_EMBEDDING]; p.rms_att = c->w[RAMA_T_RMS_ATT]; p.wq = c->w[RAMA_T_WQ]; p.wk = c->w[RAMA_T_WK];
  p.wv = c->w[RAMA_T_WV]; p.wo = c->w[RAMA_T_WO]; p.rms_ffn = c->w[RAMA_T_RMS_FFN]; p.w1 = c->w[RAMA_T_W1]; p.w2 = c->w[RAMA_T_W2];
  p.w3 = c->w[RAMA_T_W3]; p.rms_final = c->w[RAMA_T_RMS_FINAL]; p.freq_real = c->w[RAMA_T_FREQ_REAL];
  p.freq_imag = c->w[RAMA_T_FREQ_IMAG]; p.wcls = c->wcls;
  p.x0 = s->x0; p.x1 = s->x1; p.xfinal = s->xfinal; p.xb = s->xb; p.xb2 = s->xb2; p.w2out = s->w2out; p.hb = s->hb; p.hb2 = s->hb2;
  p.q = s->q; p.k = s->k; p.v = s->v; p.att = s->keep_att ? s->att : nullptr; p.logits = s->logits;
  p.key_cache = s->key_cache; p.value_cache = s->value_cache; p.attn_ws = s->attn_ws; p.tickets = s->tickets; p.part = s->part;
  p.seq = s->seq; p.ctrl = s->ctrl; p.bar = s->bar; p.prompt = s->d_prompt; p.out_tokens = s->d_out; p.n_split = s->n_split;
  p.wk_d = wk_for(c->D / 4); p.wk_wo = wk_for(c->Dq / 4); p.wk_w2 = wk_for(c->Fl / 4);
  p.mode = mode == 1 ? 1 : 0;
  p.rank = c->rank; p.world = s->p2p ? c->world : 1;
  for (int r = 0; r < kMaxPeers; ++r) p.peer_base[r] = r < c->world ? s->peer_base[r] : nullptr;
  p.off_inbox = s->off_inbox; p.off_parts = s->off_parts;
  const int grid = c->sm_count;
  size_t smem = 0;
  auto need = [&](int K4, int n_pairs, int wk) { smem = std::max(smem, gemv_smem_bytes(K4, n_pairs, grid, wk)); };
  need(c->D / 4, 3 * c->Dq / 2, p.wk_d); need(c->Dq / 4, c->D / 2, p.wk_wo); need(c->D / 4, c->Fl, p.wk_d);
  need(c->Fl / 4, c->D / 2, p.wk_w2); need(c->D / 4, (c->Vl + 1) / 2, p.wk_d);
  if (smem > kMaxDynSmem) return fail(RAMA_E_INVALID, "persistent step: %zu bytes of shared memory needed", smem);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
  });
  if (attr_err != cudaSuccess) return fail(RAMA_E_CUDA, "persistent step attribute: %s", cudaGetErrorString(attr_err));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barriers cannot deadlock
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_step_kernel, p);
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "persistent step launch: %s", cudaGetErrorString(e));
  int launches = 1;
  if (mode == 2) {
    if (c->world > 1) {
      NK(g_nccl.AllGather(s->logits + c->v0, s->logits, (size_t)c->Vl, kNcclFloat32, c->comm, st));
      ++launches;
    }
    SampleParams sp{s->logits, s->part, (c->world > 1 ? c->world : 1) * c->sm_count, c->sm_count, c->V, s->ctrl, s->d_prompt,
                    s->d_out, s->sort_keys, 0.f, 0.f, 1, peer_in_parts(s)};
    sample_kernel<<<1, kSampleThreads, 0, st>>>(sp, 0);
    CK(cudaGetLastError());
    ++launches;
  }
  if (n_launch) *n_launch = launches;
  return RAMA_OK;
}

// mode: 0 = forward only (logits + greedy partials), 1 = + chained greedy sampler, 2 = + chained top-p
static int enqueue_step(rama_session* s, cudaStream_t st, int mode, StepTrace* tr, int* n_launch) {
  rama_ctx* c = s->ctx;
  if (s->persistent && !tr) return enqueue_step_persistent(s, st, mode, n_launch);
  StepEnq q{s, st, tr};
  // PDL edges are only used inside captured graphs / plain streams without event timing
  q.pdl = tr ? 0 : c->use_pdl;
  const int D = c->D, Dq = c->Dq, Fl = c->Fl, T = c->T, hs = c->hs, L = c->L;
  const float* W[RAMA_T_COUNT];
  for (int i = 0; i < RAMA_T_COUNT; ++i) W[i] = c->w[i];
  const bool fuse_attn_wo = s->wo_part && !s->keep_att && (s->attn_wo_mode == 1 ? s->attn_bk < 2 : s->attn_bk == 0);
  const bool fuse_cluster = fuse_attn_wo && s->attn_wo_mode == 1;
  // contexts below 1024 positions: the splits of a head merge inside a thread-block cluster (attention.cuh)
  const bool attn_cluster = c->attn_cluster && !fuse_attn_wo && !s->keep_att && s->attn_bk < 2;

  // x ← embedding row of ctrl->token (infer.rs:13): a gather kernel of its own, or (RAMA_EMBED_KERNEL=0) folded into the
  // layer-0 QKV prologue (ProNorm::emb)
  const bool embed_kernel = c->embed_kernel != 0;
  if (embed_kernel) {
    q.pre(RAMA_K_EMBED);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::max(1, std::min(8, D / 4 / 256)));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (q.pdl) {
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
    }
    q.post(cudaLaunchKernelEx(&cfg, step_begin_kernel, s->ctrl, s->seq, W[RAMA_T_TOKEN_EMBEDDING], s->x0, D, c->V, q.pdl));
  }

  for (int l = 0; l < L; ++l) {
    // ---- rmsnorm → [wq|wk|wv] → RoPE → KV write (infer.rs:19-33) ----
    {
      ProNorm pro{s->x0, l == 0 ? nullptr : s->w2out, s->x1, W[RAMA_T_RMS_ATT] + (size_t)l * D, nullptr, peer_in(s, 1, l - 1)};
      if (l == 0 && !embed_kernel) { pro.emb = W[RAMA_T_TOKEN_EMBEDDING]; pro.ctrl = s->ctrl; pro.seq = s->seq; pro.vocab = c->V; }
      RowsQKV rows{W[RAMA_T_WQ] + (size_t)l * Dq * D, W[RAMA_T_WK] + (size_t)l * Dq * D,
                   W[RAMA_T_WV] + (size_t)l * Dq * D, D, Dq / 2};
      EpiQKV epi{s->q, s->k, s->v, s->key_cache + (size_t)l * T * Dq, s->value_cache + (size_t)l * T * Dq,
                 W[RAMA_T_FREQ_REAL], W[RAMA_T_FREQ_IMAG], s->ctrl, Dq / 2, hs / 2, Dq};
      const int np = 3 * Dq / 2, var = pick_variant(c, D / 4, np);
      q.pre(RAMA_K_QKV);
      // the cluster attention kernel reads older K/V rows ahead of its wait: release it after this kernel's own wait
      const int pdl_flags = q.pdl ? ((attn_cluster || fuse_cluster) ? 3 : 1) : 0;
      q.post(launch_gemv(var, pick_grid(c, var, np), st, pdl_flags, pro, rows, epi, D / 4, np));
    }
    if (fuse_attn_wo) {
      // ---- attention + wo in one launch, per-head partial outputs (infer.rs:34-35) ----
      const int J = c->sm_count / c->H;
      AttnWoParams ap{s->q, s->key_cache + (size_t)l * T * Dq, s->value_cache + (size_t)l * T * Dq,
                      W[RAMA_T_WO] + (size_t)l * D * Dq, s->xb, s->wo_part, s->ctrl, Dq, hs, D, J};
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = fuse_cluster ? dim3(c->H * kAttnClusterMax) : dim3(c->H * J);
      cfg.blockDim = fuse_cluster ? dim3(kAttnThreads) : dim3(kAttnWoThreads);
      cfg.stream = st;
      cudaLaunchAttribute at[2];
      int na = 0;
      if (q.pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
      if (fuse_cluster) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = kAttnClusterMax;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
      }
      cfg.attrs = at; cfg.numAttrs = na;
      q.pre(RAMA_K_ATTN);
      if (fuse_cluster) q.post(cudaLaunchKernelEx(&cfg, attn_wo_cluster_kernel, ap, q.pdl));
      else q.post(cudaLaunchKernelEx(&cfg, attn_wo_kernel, ap, q.pdl));
    } else {
    // ---- attention (infer.rs:34) ----
    {
      AttnParams ap{s->q, s->key_cache + (size_t)l * T * Dq, s->value_cache + (size_t)l * T * Dq, s->xb,
                    s->keep_att ? s->att : nullptr, s->attn_ws, s->tickets, s->ctrl, -1, T, Dq, hs, s->n_split,
                    // HBM idles during attention: pull this layer's wo (≤ 64 MB, fits L2) in meanwhile
                    W[RAMA_T_WO] + (size_t)l * D * Dq, std::min((size_t)D * Dq * sizeof(float), (size_t)96 << 20)};
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = attn_cluster ? dim3(c->Hl * kAttnClusterMax) : dim3(c->Hl, s->attn_gy);
      cfg.blockDim = dim3(kAttnThreads);
      cfg.stream = st;
      cudaLaunchAttribute at[2];
      int na = 0;
      if (q.pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
      if (attn_cluster) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = kAttnClusterMax;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
      }
      cfg.attrs = at; cfg.numAttrs = na;
      q.pre(RAMA_K_ATTN);
      if (attn_cluster) q.post(cudaLaunchKernelEx(&cfg, attn_cluster_kernel, ap, q.pdl));
      else q.post(cudaLaunchKernelEx(&cfg, attn_decode_kernel, ap, q.pdl));
    }
    // ---- wo (infer.rs:35); the residual add (:37) is folded into the next prologue ----
    {
      ProPlain pro{s->xb};
      RowsPlain rows{W[RAMA_T_WO] + (size_t)l * D * Dq, Dq, D};
      EpiStore epi{s->xb2, D, peer_out(s, 0, l)};
      const int np = D / 2, var = pick_variant(c, Dq / 4, np);
      q.pre(RAMA_K_WO);
      q.post(launch_gemv(var, pick_grid(c, var, np), st, q.pdl, pro, rows, epi, Dq / 4, np));
    }
    if (c->world > 1 && !s->p2p) {
      q.pre(RAMA_K_COMM);
      int e = g_nccl.AllReduce(s->xb2, s->xb2, D, kNcclFloat32, kNcclSum, c->comm, st);
      if (e && !q.nccl_err) q.nccl_err = e;
      q.post(cudaSuccess);
    }
    }  // !fuse_attn_wo
    // ---- x += xb2; rmsnorm → [w1|w3] → SwiGLU (infer.rs:37-45) ----
    {
      ProNorm pro{s->x1, s->xb2, s->x0, W[RAMA_T_RMS_FFN] + (size_t)l * D, nullptr, peer_in(s, 0, l)};
      if (fuse_attn_wo) { pro.add = s->wo_part; pro.n_add = c->H; pro.add_out = s->xb2; }
      RowsW13 rows{W[RAMA_T_W1] + (size_t)l * Fl * D, W[RAMA_T_W3] + (size_t)l * Fl * D, D};
      EpiSwiGLU epi{s->hb, s->hb2};
      const int np = Fl, var = pick_variant(c, D / 4, np);
      q.pre(RAMA_K_W13);
      q.post(launch_gemv(var, pick_grid(c, var, np), st, q.pdl, pro, rows, epi, D / 4, np));
    }
    // ---- w2 (infer.rs:46); residual add (:47) folded into the next prologue ----
    {
      ProPlain pro{s->hb};
      RowsPlain rows{W[RAMA_T_W2] + (size_t)l * D * Fl, Fl, D};
      EpiStore epi{s->w2out, D, peer_out(s, 1, l)};
      const int np = D / 2, var = pick_variant(c, Fl / 4, np);
      q.pre(RAMA_K_W2);
      q.post(launch_gemv(var, pick_grid(c, var, np), st, q.pdl, pro, rows, epi, Fl / 4, np));
    }
    if (c->world > 1 && !s->p2p) {
      q.pre(RAMA_K_COMM);
      int e = g_nccl.AllReduce(s->w2out, s->w2out, D, kNcclFloat32, kNcclSum, c->comm, st);
      if (e && !q.nccl_err) q.nccl_err = e;
      q.post(cudaSuccess);
    }
  }
  // ---- x += w2out; final rmsnorm → wcls → logits (+ greedy partials) (infer.rs:49-51) ----
  int cls_grid;
  {
    ProNorm pro{s->x0, s->w2out, s->x1, W[RAMA_T_RMS_FINAL], s->xfinal, peer_in(s, 1, L - 1)};
    RowsPlain rows{c->wcls, D, c->Vl};
    EpiCls epi{s->logits + c->v0, s->part + (size_t)c->rank * c->sm_count, c->Vl, c->v0, -INFINITY, -1,
               peer_out_parts(s)};
    const int np = (c->Vl + 1) / 2, var = pick_variant(c, D / 4);
    cls_grid = pick_grid(c, var, np);
    q.pre(RAMA_K_CLS);
    q.post(launch_gemv(var, cls_grid, st, q.pdl, pro, rows, epi, D / 4, np));
  }
  int n_part = cls_grid;
  if (c->world > 1) {
    // every rank learns every rank's per-CTA (value, index) partials: 8 B × SMs per rank
    // (p2p mode: the classifier epilogue already wrote them into every rank's array)
    int e = 0;
    if (!s->p2p) {
      q.pre(RAMA_K_COMM);
      e = g_nccl.AllGather(s->part + (size_t)c->rank * c->sm_count, s->part, (size_t)c->sm_count * 2,
                           kNcclFloat32, c->comm, st);
      if (e && !q.nccl_err) q.nccl_err = e;
      q.post(cudaSuccess);
    }
    n_part = c->world * c->sm_count;
    if (mode == 2) {
      q.pre(RAMA_K_COMM);
      e = g_nccl.AllGather(s->logits + c->v0, s->logits, (size_t)c->Vl, kNcclFloat32, c->comm, st);
      if (e && !q.nccl_err) q.nccl_err = e;
      q.post(cudaSuccess);
    }
  }
  if (mode >= 1) {
    SampleParams sp{s->logits, s->part, n_part, s->cls_grid, c->V, s->ctrl, s->d_prompt, s->d_out, s->sort_keys,
                    0.f, 0.f, 1, peer_in_parts(s)};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kSampleThreads);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (q.pdl) {
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
    }
    q.pre(RAMA_K_SAMPLE);
    q.post(cudaLaunchKernelEx(&cfg, sample_kernel, sp, q.pdl));
  }
  if (n_launch) *n_launch = q.launches;
  if (q.nccl_err) return fail(RAMA_E_NCCL, "nccl collective in step: %s", g_nccl.GetErrorString(q.nccl_err));
  if (q.err != cudaSuccess) return fail(RAMA_E_CUDA, "kernel launch in step: %s", cudaGetErrorString(q.err));
  return RAMA_OK;
}

// unused partial slots (a CTA-less tail when the classifier grid < sm_count) must read as "empty"
static int init_parts(rama_session* s) {
  rama_ctx* c = s->ctx;
  std::vector<ArgPart> h((size_t)c->world * c->sm_count, ArgPart{-INFINITY, -1});
  CK(cudaMemcpyAsync(s->part, h.data(), h.size() * sizeof(ArgPart), cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return RAMA_OK;
}

// attention grid bucket for a position: one CTA per 32-timestep chunk up to 8 / 32 / all chunks of the window
static int attn_bucket(const rama_session* s, int pos, int* gy) {
  const int need = pos / kAttnChunk + 1;
  const int b = need <= 8 ? 0 : (need <= 32 ? 1 : 2);
  *gy = std::min(s->n_split, b == 0 ? 8 : (b == 1 ? 32 : s->n_split));
  return b;
}
static int set_attn_bucket(rama_session* s, int pos) {
  s->attn_bk = attn_bucket(s, pos, &s->attn_gy);
  return s->attn_bk;
}

static int capture(rama_session* s, int mode, cudaGraphExec_t* out) {
  std::lock_guard<std::mutex> cap_lk(s->ctx->cap_mu);
  RK(init_parts(s));
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
  int n = 0;
  int rc = enqueue_step(s, s->stream, mode, nullptr, &n);
  cudaError_t e = cudaStreamEndCapture(s->stream, &g);
  if (rc != RAMA_OK) { if (g) cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  s->launches = n;
  return RAMA_OK;
}

extern "C" int rama_session_launches_per_step(const rama_session* s, int* n) {
  if (!s || !n) return fail(RAMA_E_INVALID, "NULL argument");
  const rama_ctx* c = s->ctx;
  if (s->persistent) { *n = 1; return RAMA_OK; }  // the whole greedy step is one persistent cooperative launch
  // embed + L·(qkv, attn, wo, w13, w2) + cls + sample (+ collectives under TP); attention + wo are one launch
  // for the small models at positions < 256
  *n = (c->embed_kernel ? 1 : 0) + (s->wo_part ? 4 : 5) * c->L + 1 + 1 + (c->world > 1 && !s->p2p ? 2 * c->L + 1 : 0);
  return RAMA_OK;
}

extern "C" int rama_forward(rama_session* s, int32_t token, int32_t pos) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  rama_ctx* c = s->ctx;
  if (pos < 0 || pos >= c->T)  // the reference panics on the cache slice (infer.rs:32)
    return fail(RAMA_E_STATE, "pos %d outside [0, seq_len=%d)", pos, c->T);
  if (token < 0 || token >= c->V) return fail(RAMA_E_INVALID, "token %d outside the vocabulary", token);
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  const int bk = set_attn_bucket(s, pos);
  if (!s->g[0][bk]) RK(capture(s, 0, &s->g[0][bk]));
  StepCtrl* h = &s->h_ring[s->ring_i];
  if (++s->ring_i == kRing) {  // never overwrite a slot a pending copy may still read
    s->ring_i = 0;
    CK(cudaStreamSynchronize(s->stream));
  }
  memset(h, 0, sizeof(*h));
  h->pos = pos;
  h->token = token;
  h->chained = 0;
  CK(cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, s->stream));
  CK(cudaGraphLaunch(s->g[0][bk], s->stream));
  s->logits_gathered = false;
  s->parts_valid = true;
  s->async_pending = true;
  return RAMA_OK;
}

static int gather_logits(rama_session* s) {
  rama_ctx* c = s->ctx;
  if (c->world > 1 && !s->logits_gathered) {
    NK(g_nccl.AllGather(s->logits + c->v0, s->logits, (size_t)c->Vl, kNcclFloat32, c->comm, s->stream));
    s->logits_gathered = true;
  }
  return RAMA_OK;
}

static int read_ret(rama_session* s, int32_t* next) {
  CK(cudaMemcpyAsync(s->h_ret, &s->ctrl->next, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  s->async_pending = false;
  if (s->h_ret[1] == 1) return fail(RAMA_E_STATE, "token id outside the vocabulary reached the device step");
  if (s->h_ret[1] == 2) return fail(RAMA_E_STATE, "top-p candidate list is empty (the reference panics here, infer.rs:66)");
  if (s->h_ret[1] == 3) return fail(RAMA_E_NCCL, "timed out waiting for a tensor-parallel peer's partial results");
  if (s->h_ret[1] == 4) return fail(RAMA_E_CUDA, "grid barrier of the persistent step kernel timed out");
  if (next) *next = s->h_ret[0];
  return RAMA_OK;
}

extern "C" int rama_sample(rama_session* s, float temperature, float topp, int32_t* next) {
  if (!s || !next) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  const bool greedy = temperature == 0.0f;
  if (!greedy) RK(gather_logits(s));
  const int n_part = c->world > 1 ? c->world * c->sm_count : c->sm_count;
  SampleParams sp{s->logits, s->part, n_part, s->cls_grid, c->V, s->ctrl, s->d_prompt, s->d_out, s->sort_keys,
                  temperature, topp, 0, peer_in_parts(s)};
  if (!s->parts_valid) {  // logits written by a batched step: no per-CTA partials, scan the logits
    sp.part = nullptr; sp.n_part = 0; sp.pin = PeerIn{};
  }
  sample_kernel<<<1, kSampleThreads, 0, s->stream>>>(sp, 0);
  CK(cudaGetLastError());
  return read_ret(s, next);
}

static int prefill_run(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float ms_kind[RAMA_PK_COUNT],
                       int32_t* n_launch);

extern "C" int rama_generate(rama_session* s, const int32_t* prompt, int32_t n_prompt, int32_t steps,
                             float temperature, float topp, int32_t* out_tokens, float* elapsed_ms) {
  if (!s || (n_prompt > 0 && !prompt) || !out_tokens) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = s->ctx;
  if (steps < 0 || steps > c->T)  // mod.rs has no guard: the reference panics past seq_len
    return fail(RAMA_E_STATE, "steps %d exceeds seq_len %d", steps, c->T);
  if (n_prompt < 0) return fail(RAMA_E_INVALID, "n_prompt < 0");
  for (int i = 0; i < n_prompt; ++i)
    if (prompt[i] < 0 || prompt[i] >= c->V) return fail(RAMA_E_INVALID, "prompt token %d outside the vocabulary", prompt[i]);
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  const int gi = temperature == 0.0f ? 0 : 1;
  for (int i = 0; i < steps; i += kAttnChunk) {  // make sure every bucket this run touches is captured before timing
    const int bk = set_attn_bucket(s, i);
    if (!s->g[gi + 1][bk]) RK(capture(s, gi + 1, &s->g[gi + 1][bk]));
  }
  const int np = std::min<int>(n_prompt, c->T);
  if (np) CK(cudaMemcpyAsync(s->d_prompt, prompt, np * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
  StepCtrl* h = &s->h_ring[s->ring_i];
  if (++s->ring_i == kRing) { s->ring_i = 0; CK(cudaStreamSynchronize(s->stream)); }
  memset(h, 0, sizeof(*h));
  h->pos = 0;
  h->token = 1;  // BOS (mod.rs:182)
  h->chained = 1;
  h->n_prompt = n_prompt;
  h->temperature = temperature;
  h->topp = topp;
  // Long prompts: one tensor-core prefill pass over [BOS, prompt...] instead of n_prompt+1 per-token steps
  // (the reference loop feeds them one by one and throws the logits away, mod.rs:187-192).
  const bool use_prefill = s->pf_min > 0 && n_prompt + 1 >= s->pf_min && steps > n_prompt && n_prompt < c->T;
  int first_step = 0;
  if (use_prefill) {
    std::vector<int32_t> rows((size_t)n_prompt + 1);
    rows[0] = 1;  // BOS (mod.rs:182)
    for (int i = 0; i < n_prompt; ++i) rows[i + 1] = prompt[i];
    h->pos = n_prompt;  // the step whose logits prefill leaves behind
    CK(cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->d_out, prompt, (size_t)n_prompt * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    CK(cudaEventRecord(s->ev0, s->stream));
    RK(prefill_run(s, rows.data(), n_prompt + 1, 0, nullptr, nullptr));
    if (gi == 1) RK(gather_logits(s));
    const int n_part = c->world > 1 ? c->world * c->sm_count : c->sm_count;
    SampleParams sp{s->logits, s->part, n_part, s->cls_grid, c->V, s->ctrl, s->d_prompt, s->d_out, s->sort_keys,
                    0.f, 0.f, 1, peer_in_parts(s)};
    sample_kernel<<<1, kSampleThreads, 0, s->stream>>>(sp, 0);  // samples step n_prompt, feeds the token back
    CK(cudaGetLastError());
    first_step = n_prompt + 1;
  } else {
    CK(cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, s->stream));
    CK(cudaEventRecord(s->ev0, s->stream));
  }
  for (int i = first_step; i < steps; ++i) {
    int gy;
    CK(cudaGraphLaunch(s->g[gi + 1][attn_bucket(s, i, &gy)], s->stream));
  }
  CK(cudaEventRecord(s->ev1, s->stream));
  if (steps) CK(cudaMemcpyAsync(out_tokens, s->d_out, steps * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  s->logits_gathered = gi == 1;
  RK(read_ret(s, nullptr));
  if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
  return RAMA_OK;
}

extern "C" int rama_profile_step(rama_session* s, int32_t token, int32_t pos, float ms[RAMA_K_COUNT],
                                 int32_t launches[RAMA_K_COUNT]) {
  if (!s || !ms || !launches) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = s->ctx;
  if (pos < 0 || pos >= c->T) return fail(RAMA_E_STATE, "pos %d outside [0, seq_len=%d)", pos, c->T);
  if (token < 0 || token >= c->V) return fail(RAMA_E_INVALID, "token outside the vocabulary");
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  RK(init_parts(s));
  set_attn_bucket(s, pos);
  StepCtrl* h = &s->h_ring[s->ring_i];
  if (++s->ring_i == kRing) { s->ring_i = 0; CK(cudaStreamSynchronize(s->stream)); }
  memset(h, 0, sizeof(*h));
  h->pos = pos; h->token = token; h->chained = 0;
  CK(cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, s->stream));
  StepTrace tr;
  int n = 0;
  int rc = enqueue_step(s, s->stream, 0, &tr, &n);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  for (int i = 0; i < RAMA_K_COUNT; ++i) { ms[i] = 0.f; launches[i] = 0; }
  for (size_t i = 0; i < tr.kind.size(); ++i) {
    float t = 0.f;
    if (rc == RAMA_OK && e == cudaSuccess) cudaEventElapsedTime(&t, tr.ev[2 * i], tr.ev[2 * i + 1]);
    ms[tr.kind[i]] += t;
    launches[tr.kind[i]] += 1;
  }
  for (cudaEvent_t ev : tr.ev) cudaEventDestroy(ev);
  s->logits_gathered = false;
  if (rc != RAMA_OK) return rc;
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "profile step: %s", cudaGetErrorString(e));
  return RAMA_OK;
}

// Phase timeline of one persistent step: clock64() of CTA 0 at kernel entry and before/after each grid barrier
// (2·(5L+1)+1 stamps).  Tool for tools/step_trace.py.
extern "C" int rama_step_trace(rama_session* s, int32_t token, int32_t pos, long long* stamps, int32_t cap, int32_t* n_out) {
  if (!s || !stamps || !n_out) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = s->ctx;
  if (!s->persistent) return fail(RAMA_E_STATE, "session does not use the persistent step kernel");
  if (pos < 0 || pos >= c->T || token < 0 || token >= c->V) return fail(RAMA_E_INVALID, "token/pos out of range");
  const int n = 2 * (5 * c->L + 1) + 1;
  if (cap < n) return fail(RAMA_E_INVALID, "need room for %d stamps", n);
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  long long* d = nullptr;
  CK(cudaMalloc((void**)&d, n * sizeof(long long)));
  StepCtrl* h = &s->h_ring[s->ring_i];
  if (++s->ring_i == kRing) { s->ring_i = 0; CK(cudaStreamSynchronize(s->stream)); }
  memset(h, 0, sizeof(*h));
  h->pos = pos; h->token = token;
  CK(cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, s->stream));
  int rc = enqueue_step_persistent(s, s->stream, 0, nullptr, d);
  if (rc == RAMA_OK) {
    CK(cudaMemcpyAsync(stamps, d, n * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
  }
  cudaFree(d);
  *n_out = n;
  s->logits_gathered = false;
  s->parts_valid = true;
  return rc;
}

extern "C" int rama_logits_to_host(rama_session* s, float* dst, size_t n) {
  if (!s || !dst) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = s->ctx;
  if (n < (size_t)c->V) return fail(RAMA_E_INVALID, "buffer too small");
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  RK(gather_logits(s));
  CK(cudaMemcpyAsync(dst, s->logits, (size_t)c->V * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return RAMA_OK;
}

extern "C" int rama_state_to_host(rama_session* s, int buf, float* dst, size_t n, size_t* n_out) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  rama_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  const float* src = nullptr;
  size_t have = 0;
  switch (buf) {
    case RAMA_S_X: src = s->xfinal; have = c->D; break;     // after forward: final rmsnorm output (infer.rs:50)
    case RAMA_S_XB: src = s->x1; have = c->D; break;        // pre-norm residual copy (infer.rs:49)
    case RAMA_S_XB2: src = s->xb2; have = c->D; break;
    case RAMA_S_HB: src = s->hb; have = c->Fl; break;
    case RAMA_S_HB2: src = s->hb2; have = c->Fl; break;
    case RAMA_S_Q: src = s->q; have = c->Dq; break;
    case RAMA_S_K: src = s->k; have = c->Dq; break;
    case RAMA_S_V: src = s->v; have = c->Dq; break;
    case RAMA_S_ATT: src = s->att; have = (size_t)c->Hl * c->T; break;
    case RAMA_S_LOGITS: RK(gather_logits(s)); src = s->logits; have = c->V; break;
    case RAMA_S_KEY_CACHE: src = s->key_cache; have = (size_t)c->L * c->T * c->Dq; break;
    case RAMA_S_VALUE_CACHE: src = s->value_cache; have = (size_t)c->L * c->T * c->Dq; break;
    default: return fail(RAMA_E_INVALID, "unknown state buffer %d", buf);
  }
  if (n_out) *n_out = have;
  if (!dst) return RAMA_OK;
  if (n < have) return fail(RAMA_E_INVALID, "buffer too small: %zu < %zu", n, have);
  CK(cudaMemcpyAsync(dst, src, have * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// prompt prefill (tensor cores): ≙ the prompt part of generate()'s loop, mod.rs:187-192
// ------------------------------------------------------------------------------------------------
// kernel launch with the programmatic-stream-serialization attribute (PDL): the next kernel's CTAs start while this one
// drains; every kernel of the batched step executes griddepcontrol.wait before it touches memory (batch.cuh)
template <class... P, class... A>
static cudaError_t launch_k(bool pdl, void (*kern)(P...), dim3 grid, dim3 block, cudaStream_t st, A&&... args) {
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

constexpr int kPrefillChunk = 512;

static int ensure_prefill_ws(rama_session* s) {
  if (s->pf_cap) return RAMA_OK;
  rama_ctx* c = s->ctx;
  std::lock_guard<std::mutex> cap_lk(c->cap_mu);  // allocations vs another thread's stream capture (rama_ctx::cap_mu)
  const size_t cap = std::min(c->T, kPrefillChunk);
  cudaError_t e = cudaSuccess;
#define A(call) if (e == cudaSuccess) e = (call)
  A(cudaMalloc((void**)&s->pf_x, cap * c->D * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_xn, cap * c->D * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_q, cap * c->Dq * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_att, cap * c->Dq * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_y, cap * c->D * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_h, cap * c->Fl * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_tokens, cap * sizeof(int32_t)));
#undef A
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "prefill workspace: %s", cudaGetErrorString(e));
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(prefill_attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)prefill_attn_smem_bytes(kPfMaxHs, 4));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(prefill_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)prefill_attn_smem_bytes(kPfMaxHs, 1));
  });
  if (attr_err != cudaSuccess) return fail(RAMA_E_CUDA, "prefill attention smem: %s", cudaGetErrorString(attr_err));
  s->pf_cap = (int)cap;
  return RAMA_OK;
}

struct PfTrace {  // optional per-launch CUDA-event timing by kind (rama_prefill's ms_kind)
  cudaStream_t st;
  bool on;
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  void pre(int k) {
    if (!on) return;
    cudaEvent_t a; cudaEventCreate(&a); cudaEventRecord(a, st); ev.push_back(a); kind.push_back(k);
  }
  void post() {
    if (!on) return;
    cudaEvent_t b; cudaEventCreate(&b); cudaEventRecord(b, st); ev.push_back(b);
  }
};

// one chunk of M ≤ pf_cap rows at positions [pos0, pos0+M); `last`: also produce the logits of the final row
static int prefill_chunk(rama_session* s, const int32_t* tokens, int M, int pos0, bool last, PfTrace& tr, int* n_launch) {
  rama_ctx* c = s->ctx;
  cudaStream_t st = s->stream;
  const int D = c->D, Dq = c->Dq, Fl = c->Fl, T = c->T, hs = c->hs, L = c->L;
  const float* const* W = c->w;
  int launches = 0;
#define GK(kind, call)                                                                                         \
  do {                                                                                                         \
    tr.pre(kind);                                                                                              \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ == cudaSuccess) e_ = cudaGetLastError();                                                            \
    tr.post();                                                                                                 \
    ++launches;                                                                                                \
    if (e_ != cudaSuccess) return fail(RAMA_E_CUDA, "prefill launch %s: %s", #call, cudaGetErrorString(e_));   \
  } while (0)
  CK(cudaMemcpyAsync(s->pf_tokens, tokens, (size_t)M * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  tr.pre(RAMA_PK_OTHER);
  // programmatic dependent launch along the whole chain (not while per-launch events are being recorded)
  const bool pdl = c->use_pdl && !tr.on && env_int("RAMA_PREFILL_PDL", 1);
  CK(launch_k(pdl, prefill_embed_kernel, dim3(M), dim3(256), st, (const int32_t*)s->pf_tokens, W[RAMA_T_TOKEN_EMBEDDING], s->pf_x, D, c->V,
              &s->ctrl->error, s->seq));
  tr.post(); ++launches;
  CK(cudaGetLastError());
  for (int l = 0; l < L; ++l) {
    float* kc = s->key_cache + (size_t)l * T * Dq;
    float* vc = s->value_cache + (size_t)l * T * Dq;
    // x += pending w2 output; xn = rmsnorm(x)·w_att   (infer.rs:19)
    tr.pre(RAMA_PK_NORM);
    CK(launch_k(pdl, prefill_addnorm_kernel, dim3(M), dim3(256), st, s->pf_x, (const float*)(l == 0 ? nullptr : s->pf_y),
                W[RAMA_T_RMS_ATT] + (size_t)l * D, s->pf_xn, D));
    tr.post(); ++launches;
    // [wq;wk;wv] → RoPE → Q, KV-cache rows   (infer.rs:20-33)
    {
      GemmOperand A{s->pf_xn, (size_t)M, (size_t)D};
      GemmOperand B[3] = {{W[RAMA_T_WQ] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WK] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WV] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D}};
      EpiQKVPrefill epi{s->pf_q, kc, vc, W[RAMA_T_FREQ_REAL], W[RAMA_T_FREQ_IMAG], pos0, Dq, hs / 2};
      GK(RAMA_PK_GEMM, (launch_gemm_tf32x3<128, 4, 4, 0>(st, &A, 1, B, 3, M, Dq, D, 0, 1, epi, pdl)));
    }
    // causal attention of every prompt row over the cache   (infer.rs:34)
    {
      PrefillAttnParams ap{s->pf_q, kc, vc, s->pf_att, M, pos0, Dq, hs};
      tr.pre(RAMA_PK_ATTN);
      // 64-query blocks when that still fills the machine, else 16-query blocks (few heads per rank, short prompts)
      const int nq64 = (M + 63) / 64, nq16 = (M + 15) / 16;
      cudaLaunchConfig_t cfg{};
      cfg.blockDim = dim3(kPfThreads);
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      if (pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
      }
      if (((nq64 + 1) / 2) * c->Hl >= 96) {
        cfg.gridDim = dim3((nq64 + 1) / 2, c->Hl);
        cfg.dynamicSmemBytes = prefill_attn_smem_bytes(hs, 4);
        CK(cudaLaunchKernelEx(&cfg, prefill_attn_kernel<4>, ap));
      } else {
        cfg.gridDim = dim3((nq16 + 1) / 2, c->Hl);
        cfg.dynamicSmemBytes = prefill_attn_smem_bytes(hs, 1);
        CK(cudaLaunchKernelEx(&cfg, prefill_attn_kernel<1>, ap));
      }
      tr.post(); ++launches;
    }
    // wo   (infer.rs:35); the residual add is the next addnorm
    {
      GemmOperand A{s->pf_att, (size_t)M, (size_t)Dq};
      GemmOperand B{W[RAMA_T_WO] + (size_t)l * D * Dq, (size_t)D, (size_t)Dq};
      EpiStoreNT epi{s->pf_y, D, D, 0};
      GK(RAMA_PK_GEMM, (launch_gemm_tf32x3<128, 4, 4, 0>(st, &A, 1, &B, 1, M, D, Dq, 0, 1, epi, pdl)));
    }
    if (c->world > 1) {
      tr.pre(RAMA_PK_COMM);
      NK(g_nccl.AllReduce(s->pf_y, s->pf_y, (size_t)M * D, kNcclFloat32, kNcclSum, c->comm, st));
      tr.post(); ++launches;
    }
    tr.pre(RAMA_PK_NORM);
    CK(launch_k(pdl, prefill_addnorm_kernel, dim3(M), dim3(256), st, s->pf_x, (const float*)s->pf_y, W[RAMA_T_RMS_FFN] + (size_t)l * D,
                s->pf_xn, D));
    tr.post(); ++launches;
    // [w1|w3] → SwiGLU   (infer.rs:39-45)
    {
      GemmOperand A{s->pf_xn, (size_t)M, (size_t)D};
      GemmOperand B[2] = {{W[RAMA_T_W1] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D},
                          {W[RAMA_T_W3] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D}};
      EpiSwiGLUPrefill epi{s->pf_h, Fl};
      GK(RAMA_PK_GEMM, (launch_gemm_tf32x3<128, 4, 4, 0>(st, &A, 1, B, 2, M, Fl, D, 0, 1, epi, pdl)));
    }
    // w2   (infer.rs:46)
    {
      GemmOperand A{s->pf_h, (size_t)M, (size_t)Fl};
      GemmOperand B{W[RAMA_T_W2] + (size_t)l * D * Fl, (size_t)D, (size_t)Fl};
      EpiStoreNT epi{s->pf_y, D, D, 0};
      GK(RAMA_PK_GEMM, (launch_gemm_tf32x3<128, 4, 4, 0>(st, &A, 1, &B, 1, M, D, Fl, 0, 1, epi, pdl)));
    }
    if (c->world > 1) {
      tr.pre(RAMA_PK_COMM);
      NK(g_nccl.AllReduce(s->pf_y, s->pf_y, (size_t)M * D, kNcclFloat32, kNcclSum, c->comm, st));
      tr.post(); ++launches;
    }
  }
  CK(cudaGetLastError());
  if (last) {
    // only the last row's logits exist after the reference's prompt loop: x0 = x + y of that row, then the
    // decode path's fused final-rmsnorm → classifier GEMV (infer.rs:49-51)
    tr.pre(RAMA_PK_OTHER);
    CK(launch_k(pdl, prefill_last_row_kernel, dim3(std::max(1, D / 256)), dim3(256), st, (const float*)(s->pf_x + (size_t)(M - 1) * D),
                (const float*)(s->pf_y + (size_t)(M - 1) * D), s->x0, D));
    tr.post(); ++launches;
    RK(init_parts(s));
    ProNorm pro{s->x0, nullptr, s->x1, W[RAMA_T_RMS_FINAL], s->xfinal, PeerIn{}};
    RowsPlain rows{c->wcls, D, c->Vl};
    EpiCls epi{s->logits + c->v0, s->part + (size_t)c->rank * c->sm_count, c->Vl, c->v0, -INFINITY, -1, peer_out_parts(s)};
    const int np = (c->Vl + 1) / 2, var = pick_variant(c, D / 4);
    GK(RAMA_PK_OTHER, launch_gemv(var, pick_grid(c, var, np), st, 0, pro, rows, epi, D / 4, np));
    if (c->world > 1 && !s->p2p) {
      NK(g_nccl.AllGather(s->part + (size_t)c->rank * c->sm_count, s->part, (size_t)c->sm_count * 2, kNcclFloat32, c->comm, st));
      ++launches;
    }
  }
#undef GK
  if (n_launch) *n_launch += launches;
  return RAMA_OK;
}

static int prefill_run(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float ms_kind[RAMA_PK_COUNT],
                       int32_t* n_launch) {
  rama_ctx* c = s->ctx;
  if (n <= 0 || pos0 < 0 || (long long)pos0 + n > c->T)  // the reference panics past seq_len (infer.rs:32)
    return fail(RAMA_E_STATE, "prefill rows [%d, %d) outside [0, seq_len=%d)", pos0, pos0 + n, c->T);
  for (int i = 0; i < n; ++i)
    if (tokens[i] < 0 || tokens[i] >= c->V) return fail(RAMA_E_INVALID, "prompt token %d outside the vocabulary", tokens[i]);
  CK(cudaSetDevice(c->device));
  RK(ensure_prefill_ws(s));
  PfTrace tr{s->stream, ms_kind != nullptr, {}, {}};
  int launches = 0, rc = RAMA_OK;
  for (int c0 = 0; c0 < n && rc == RAMA_OK; c0 += s->pf_cap) {
    const int M = std::min(s->pf_cap, n - c0);
    rc = prefill_chunk(s, tokens + c0, M, pos0 + c0, c0 + M == n, tr, &launches);
  }
  cudaError_t e = cudaSuccess;
  if (tr.on) {
    e = cudaStreamSynchronize(s->stream);
    for (int i = 0; i < RAMA_PK_COUNT; ++i) ms_kind[i] = 0.f;
    for (size_t i = 0; i < tr.kind.size() && 2 * i + 1 < tr.ev.size(); ++i) {
      float t = 0.f;
      if (rc == RAMA_OK && e == cudaSuccess) cudaEventElapsedTime(&t, tr.ev[2 * i], tr.ev[2 * i + 1]);
      ms_kind[tr.kind[i]] += t;
    }
    for (cudaEvent_t ev : tr.ev) cudaEventDestroy(ev);
  }
  if (n_launch) *n_launch = launches;
  s->logits_gathered = false;
  s->parts_valid = true;
  if (rc != RAMA_OK) return rc;
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "prefill: %s", cudaGetErrorString(e));
  return RAMA_OK;
}

extern "C" int rama_prefill(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float* elapsed_ms,
                            float ms_kind[RAMA_PK_COUNT], int32_t* n_launch) {
  if (!s || !tokens) return fail(RAMA_E_INVALID, "NULL argument");
  CK(cudaSetDevice(s->ctx->device));
  CK(session_enter(s));
  CK(cudaEventRecord(s->ev0, s->stream));
  RK(prefill_run(s, tokens, n, pos0, ms_kind, n_launch));
  CK(cudaEventRecord(s->ev1, s->stream));
  RK(read_ret(s, nullptr));  // synchronises; surfaces device-side errors
  if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
  return RAMA_OK;
}

extern "C" int rama_session_set_prefill(rama_session* s, int32_t min_rows) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  s->pf_min = min_rows;
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// batched multi-sequence decode (server path, lib.rs:127-160): one step for n sessions
// ------------------------------------------------------------------------------------------------
extern "C" int rama_batch_destroy(rama_batch* b);
constexpr int kBatchMax = 64;    // one 64-column MMA tile of sequences
constexpr int kBatchRing = 8;
// batched-decode GEMM tile: 128 weight rows × 64 sequences, 2 stages, chunks of 2 k-blocks; two CTAs fit an SM
// (64 KB of shared memory and 256 TMEM columns each) and interleave their pipelines
#define BATCH_GEMM launch_gemm_tf32x3<64, 2, 4, 0>  // 128-k chunks as in prefill (CH = 2: 9.18 ms per 64-sequence step)
constexpr int kBatchCtasPerSm = GemmSmem<64, 2, 0>::kCtasPerSm;

struct rama_batch {
  rama_ctx* ctx = nullptr;
  int cap = 0, n_split = 1;
  cudaStream_t stream = nullptr;
  float *x = nullptr, *xn = nullptr, *q = nullptr, *att = nullptr, *h = nullptr, *part = nullptr, *attn_ws = nullptr;
  float *red = nullptr, *lstage = nullptr;  // tensor parallelism: all-reduce buffer [B][D], logits all-gather staging [P][B][Vl]
  unsigned int* tickets = nullptr;
  size_t part_floats = 0;
  BatchSeq* d_seqs = nullptr;
  BatchSeq* h_seqs = nullptr;        // pinned ring [kBatchRing][cap]
  SampleParams* d_sp = nullptr;
  SampleParams* h_sp = nullptr;      // pinned [cap]
  int32_t* d_next = nullptr;
  int32_t* h_next = nullptr;         // pinned [2·cap]
  int ring_i = 0;
  std::vector<cudaGraphExec_t> graphs;  // by batch size
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int launches = 0;
  std::shared_ptr<BatchFence> fence;    // re-recorded after every step; the sessions of the step hold a reference
};

// split-K factor: the smallest one that fills ≥ 92 % of the CTA slots of its last wave (every extra split writes and
// re-reads another [n][rows] partial), else the best-filling one; ≥ 8 k-blocks per split
static int pick_ksplit(const rama_ctx* c, int tiles, int K) {
  const int total_kb = (K + kGemmBK - 1) / kGemmBK;
  const int slots = c->sm_count * kBatchCtasPerSm;  // CTAs resident at once
  int best = 1;
  double best_eff = 0.0;
  for (int S = 1; S <= 16; ++S) {
    if (S > 1 && total_kb / S < 8) break;
    const int units = tiles * S, waves = (units + slots - 1) / slots;
    const double eff = (double)units / ((double)waves * slots);
    if (eff >= 0.92) return S;
    if (eff > best_eff + 0.02) { best_eff = eff; best = S; }
  }
  return best;
}

extern "C" int rama_batch_create(rama_ctx* c, int32_t max_seqs, rama_batch** out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!c->loaded) return fail(RAMA_E_STATE, "no weights loaded");
  if (max_seqs < 1 || max_seqs > kBatchMax) return fail(RAMA_E_INVALID, "max_seqs must be in [1, %d]", kBatchMax);
  CK(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> cap_lk(c->cap_mu);
  rama_batch* b = new rama_batch();
  b->ctx = c;
  b->cap = max_seqs;
  b->n_split = (c->T + kAttnChunk - 1) / kAttnChunk;
  const size_t B = max_seqs, D = c->D, Dq = c->Dq, Fl = c->Fl;
  // partial buffer: the largest of [3][S][B][Dq], [S][B][D], [2][S][B][Fl], [S][B][Vl] over the chosen split factors
  auto tiles = [](int rows) { return (rows + kGemmBM - 1) / kGemmBM; };
  size_t pf = 0;
  pf = std::max(pf, (size_t)3 * pick_ksplit(c, 3 * tiles(c->Dq), c->D) * B * Dq);
  pf = std::max(pf, (size_t)pick_ksplit(c, tiles(c->D), c->Dq) * B * D);
  pf = std::max(pf, (size_t)2 * pick_ksplit(c, 2 * tiles(c->Fl), c->D) * B * Fl);
  pf = std::max(pf, (size_t)pick_ksplit(c, tiles(c->D), c->Fl) * B * D);
  pf = std::max(pf, (size_t)pick_ksplit(c, tiles(c->Vl), c->D) * B * c->Vl);
  b->part_floats = pf;
  cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
#define A(call) if (e == cudaSuccess) e = (call)
  A(dalloc(&b->x, B * D)); A(dalloc(&b->xn, B * D)); A(dalloc(&b->q, B * Dq)); A(dalloc(&b->att, B * Dq));
  A(dalloc(&b->h, B * Fl)); A(dalloc(&b->part, pf));
  if (c->world > 1) { A(dalloc(&b->red, B * D)); A(dalloc(&b->lstage, (size_t)c->world * B * c->Vl)); }
  A(dalloc(&b->attn_ws, B * c->Hl * b->n_split * (c->hs + 2)));
  A(dalloc(&b->tickets, B * c->Hl));
  A(dalloc(&b->d_seqs, B)); A(dalloc(&b->d_sp, B)); A(dalloc(&b->d_next, 2 * B));
  A(cudaHostAlloc((void**)&b->h_seqs, kBatchRing * B * sizeof(BatchSeq), cudaHostAllocDefault));
  A(cudaHostAlloc((void**)&b->h_sp, B * sizeof(SampleParams), cudaHostAllocDefault));
  A(cudaHostAlloc((void**)&b->h_next, 2 * B * sizeof(int32_t), cudaHostAllocDefault));
  A(cudaEventCreate(&b->ev0)); A(cudaEventCreate(&b->ev1));
  A(cudaDeviceSynchronize());
#undef A
  if (e != cudaSuccess) {
    rama_batch_destroy(b);
    return fail(RAMA_E_CUDA, "batch allocation: %s", cudaGetErrorString(e));
  }
  b->graphs.assign(max_seqs + 1, nullptr);
  b->fence = std::make_shared<BatchFence>();
  if (cudaEventCreateWithFlags(&b->fence->ev, cudaEventDisableTiming) != cudaSuccess) {
    rama_batch_destroy(b);
    return fail(RAMA_E_CUDA, "batch fence event");
  }
  c->n_objects.fetch_add(1);
  *out = b;
  return RAMA_OK;
}

extern "C" int rama_batch_destroy(rama_batch* b) {
  if (!b) return RAMA_OK;
  cudaSetDevice(b->ctx->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  std::lock_guard<std::mutex> cap_lk(b->ctx->cap_mu);
  for (auto g : b->graphs) if (g) cudaGraphExecDestroy(g);
  void* bufs[] = {b->x, b->xn, b->q, b->att, b->h, b->part, b->attn_ws, b->tickets, b->d_seqs, b->d_sp, b->d_next, b->red, b->lstage};
  for (void* p : bufs) if (p) cudaFree(p);
  if (b->h_seqs) cudaFreeHost(b->h_seqs);
  if (b->h_sp) cudaFreeHost(b->h_sp);
  if (b->h_next) cudaFreeHost(b->h_next);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  if (b->fence) b->ctx->n_objects.fetch_sub(1);  // counted only once fully created
  delete b;
  return RAMA_OK;
}

// enqueue one batched step for n sequences (everything per-sequence is read from b->d_seqs on the device,
// so the captured graph of a batch size serves every step)
static int enqueue_batch_step(rama_batch* b, int n, int* n_launch) {
  rama_ctx* c = b->ctx;
  cudaStream_t st = b->stream;
  const int D = c->D, Dq = c->Dq, Fl = c->Fl, T = c->T, hs = c->hs, L = c->L, Vl = c->Vl;
  const float* const* W = c->w;
  int launches = 0;
  const bool pdl = c->use_pdl && env_int("RAMA_BATCH_PDL", 1);
  auto tiles = [](int rows) { return (rows + kGemmBM - 1) / kGemmBM; };
#define LK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ == cudaSuccess) e_ = cudaGetLastError();                                                           \
    ++launches;                                                                                               \
    if (e_ != cudaSuccess) return fail(RAMA_E_CUDA, "batched step launch %s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
  LK(launch_k(pdl, batch_embed_kernel, dim3(n), dim3(256), st, b->d_seqs, W[RAMA_T_TOKEN_EMBEDDING], b->x, D, c->V));
  GemmOperand X{b->xn, (size_t)n, (size_t)D};
  int S_prev = 0;  // split factor of the pending residual partials in b->part (0: none)
  const float* pending = b->part;
  // tensor parallelism: row-parallel wo / w2 leave a partial [n][D] on every rank — sum the split-K partials, all-reduce
  // over NVLink (NCCL, 1 MB at 64 sequences), and hand the reduced buffer to the next addnorm as a single "split"
  auto reduce_ranks = [&](int& S) -> int {
    if (c->world <= 1) return RAMA_OK;
    LK(launch_k(pdl, sum_partials_kernel, dim3(c->sm_count * 2), dim3(256), st, b->red, (const float*)b->part, (size_t)n * D, S));
    NK(g_nccl.AllReduce(b->red, b->red, (size_t)n * D, kNcclFloat32, kNcclSum, c->comm, st));
    ++launches;
    S = 1;
    pending = b->red;
    return RAMA_OK;
  };
  for (int l = 0; l < L; ++l) {
    const size_t layer_off = (size_t)l * T * Dq;
    // x += pending w2 output; xn = rmsnorm(x)   (infer.rs:19, :47 of the previous layer)
    LK(launch_k(pdl, batch_addnorm_kernel, dim3(n), dim3(kBatchNormThreads), st, b->x, S_prev ? pending : nullptr, S_prev,
                (size_t)n * D, W[RAMA_T_RMS_ATT] + (size_t)l * D, b->xn, D));
    {  // [wq;wk;wv] (weights = the 128-row operand, the batch = the 64-column operand)   (infer.rs:20-23)
      GemmOperand A[3] = {{W[RAMA_T_WQ] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WK] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WV] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D}};
      const int S = pick_ksplit(c, 3 * tiles(Dq), D);
      EpiStoreT epi{b->part, Dq, n, S, (size_t)n * Dq};
      LK((BATCH_GEMM(st, A, 3, &X, 1, Dq, n, D, 0, S, epi, pdl)));
      LK(launch_k(pdl, batch_qkv_finish_kernel, dim3(n, (Dq / 2 + 255) / 256), dim3(256), st, b->part, S, (size_t)n * Dq, b->d_seqs, layer_off, b->q, W[RAMA_T_FREQ_REAL], W[RAMA_T_FREQ_IMAG], Dq, hs / 2));
    }
    {  // attention per sequence   (infer.rs:34)
      AttnBatchParams ap{b->d_seqs, b->q, b->att, b->attn_ws, b->tickets, layer_off, T, Dq, hs, b->n_split, c->Hl};
      LK(launch_k(pdl, attn_decode_batch_kernel, dim3(c->Hl, std::min(b->n_split, 2), n), dim3(kAttnThreads), st, ap));  // CTAs stride over the chunks
    }
    int S_wo;
    {  // wo   (infer.rs:35)
      GemmOperand A{W[RAMA_T_WO] + (size_t)l * D * Dq, (size_t)D, (size_t)Dq};
      GemmOperand Bm{b->att, (size_t)n, (size_t)Dq};
      S_wo = pick_ksplit(c, tiles(D), Dq);
      EpiStoreT epi{b->part, D, n, S_wo, (size_t)n * D};
      LK((BATCH_GEMM(st, &A, 1, &Bm, 1, D, n, Dq, 0, S_wo, epi, pdl)));
      RK(reduce_ranks(S_wo));
    }
    // x += wo output; xn = rmsnorm(x)   (infer.rs:37-38)
    LK(launch_k(pdl, batch_addnorm_kernel, dim3(n), dim3(kBatchNormThreads), st, b->x, pending, S_wo, (size_t)n * D, W[RAMA_T_RMS_FFN] + (size_t)l * D, b->xn, D));
    {  // [w1;w3] → SwiGLU   (infer.rs:39-45)
      GemmOperand A[2] = {{W[RAMA_T_W1] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D},
                          {W[RAMA_T_W3] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D}};
      const int S = pick_ksplit(c, 2 * tiles(Fl), D);
      EpiStoreT epi{b->part, Fl, n, S, (size_t)n * Fl};
      LK((BATCH_GEMM(st, A, 2, &X, 1, Fl, n, D, 0, S, epi, pdl)));
      LK(launch_k(pdl, batch_swiglu_finish_kernel, dim3(std::min(c->sm_count * 4, (n * Fl + 255) / 256)), dim3(256), st, b->part, S, (size_t)n * Fl, b->h, Fl, n));
    }
    {  // w2   (infer.rs:46)
      GemmOperand A{W[RAMA_T_W2] + (size_t)l * D * Fl, (size_t)D, (size_t)Fl};
      GemmOperand Bm{b->h, (size_t)n, (size_t)Fl};
      S_prev = pick_ksplit(c, tiles(D), Fl);
      EpiStoreT epi{b->part, D, n, S_prev, (size_t)n * D};
      LK((BATCH_GEMM(st, &A, 1, &Bm, 1, D, n, Fl, 0, S_prev, epi, pdl)));
      RK(reduce_ranks(S_prev));
    }
  }
  // x += w2 output; final rmsnorm; classifier → each session's logits   (infer.rs:49-51)
  LK(launch_k(pdl, batch_addnorm_kernel, dim3(n), dim3(kBatchNormThreads), st, b->x, pending, S_prev, (size_t)n * D, W[RAMA_T_RMS_FINAL], b->xn, D));
  {
    GemmOperand A{c->wcls, (size_t)Vl, (size_t)D};
    const int S = pick_ksplit(c, tiles(Vl), D);
    EpiStoreT epi{b->part, Vl, n, S, (size_t)n * Vl};
    LK((BATCH_GEMM(st, &A, 1, &X, 1, Vl, n, D, 0, S, epi, pdl)));
    if (c->world > 1) {  // vocabulary rows are split: gather every rank's block, then scatter into the sessions' logits
      float* mine = b->lstage + (size_t)c->rank * n * Vl;
      LK(launch_k(pdl, batch_cls_stage_kernel, dim3(std::min(64, (Vl + 255) / 256), n), dim3(256), st, b->part, S, (size_t)n * Vl, mine, Vl));
      NK(g_nccl.AllGather(mine, b->lstage, (size_t)n * Vl, kNcclFloat32, c->comm, st));
      ++launches;
      LK(launch_k(pdl, batch_logits_scatter_kernel, dim3(std::min(64, (c->V + 255) / 256), n), dim3(256), st, b->lstage, b->d_seqs, Vl, c->world, n));
    } else {
      LK(launch_k(pdl, batch_cls_finish_kernel, dim3(std::min(64, (Vl + 255) / 256), n), dim3(256), st, b->part, S, (size_t)n * Vl, b->d_seqs, Vl, c->v0));
    }
  }
#undef LK
  if (n_launch) *n_launch = launches;
  return RAMA_OK;
}

static int batch_check_sessions(rama_batch* b, rama_session* const* sessions, int32_t n) {
  if (n < 1 || n > b->cap) return fail(RAMA_E_INVALID, "batch of %d sequences outside [1, %d]", n, b->cap);
  for (int i = 0; i < n; ++i) {
    if (!sessions[i] || sessions[i]->ctx != b->ctx) return fail(RAMA_E_INVALID, "session %d is NULL or belongs to another context", i);
    for (int j = 0; j < i; ++j)
      if (sessions[j] == sessions[i]) return fail(RAMA_E_INVALID, "session %d appears twice in the batch", i);
  }
  return RAMA_OK;
}

extern "C" int rama_forward_batch(rama_batch* b, rama_session* const* sessions, const int32_t* tokens,
                                  const int32_t* pos, int32_t n) {
  if (!b || !sessions || !tokens || !pos) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = b->ctx;
  RK(batch_check_sessions(b, sessions, n));
  for (int i = 0; i < n; ++i) {
    if (pos[i] < 0 || pos[i] >= c->T) return fail(RAMA_E_STATE, "pos %d of sequence %d outside [0, seq_len=%d)", pos[i], i, c->T);
    if (tokens[i] < 0 || tokens[i] >= c->V) return fail(RAMA_E_INVALID, "token %d of sequence %d outside the vocabulary", tokens[i], i);
  }
  CK(cudaSetDevice(c->device));
  BatchSeq* hs = b->h_seqs + (size_t)b->ring_i * b->cap;
  if (++b->ring_i == kBatchRing) { b->ring_i = 0; CK(cudaStreamSynchronize(b->stream)); }
  for (int i = 0; i < n; ++i) {
    rama_session* s = sessions[i];
    hs[i] = BatchSeq{s->key_cache, s->value_cache, s->logits, s->ctrl, pos[i], tokens[i]};
    s->logits_gathered = true;   // under TP the batched step leaves the full vocabulary in every session
    s->parts_valid = false;
    // stream ordering, both ways: the step waits for work the session still has in flight on its own stream (an async
    // rama_forward), and the session's next own operation waits for this step (session_enter)
    if (s->fence && s->fence != b->fence) CK(cudaStreamWaitEvent(b->stream, s->fence->ev, 0));  // last touched by another batch
    if (s->async_pending) {
      CK(cudaEventRecord(s->ev1, s->stream));
      CK(cudaStreamWaitEvent(b->stream, s->ev1, 0));
      s->async_pending = false;
    }
    s->fence = b->fence;
  }
  CK(cudaMemcpyAsync(b->d_seqs, hs, (size_t)n * sizeof(BatchSeq), cudaMemcpyHostToDevice, b->stream));
  if (!b->graphs[n]) {
    std::lock_guard<std::mutex> cap_lk(c->cap_mu);
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeRelaxed));
    int nl = 0;
    int rc = enqueue_batch_step(b, n, &nl);
    cudaError_t e = cudaStreamEndCapture(b->stream, &g);
    if (rc != RAMA_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&b->graphs[n], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    b->launches = nl;
  }
  CK(cudaGraphLaunch(b->graphs[n], b->stream));
  CK(cudaEventRecord(b->fence->ev, b->stream));
  return RAMA_OK;
}

extern "C" int rama_sample_batch(rama_batch* b, rama_session* const* sessions, int32_t n, float temperature,
                                 float topp, int32_t* next) {
  if (!b || !sessions || !next) return fail(RAMA_E_INVALID, "NULL argument");
  rama_ctx* c = b->ctx;
  RK(batch_check_sessions(b, sessions, n));
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(b->stream));  // h_sp / h_next are single-buffered
  BatchSeq* hs = b->h_seqs + (size_t)b->ring_i * b->cap;
  if (++b->ring_i == kBatchRing) b->ring_i = 0;
  for (int i = 0; i < n; ++i) {
    rama_session* s = sessions[i];
    b->h_sp[i] = SampleParams{s->logits, nullptr, 0, 0, c->V, s->ctrl, nullptr, nullptr, s->sort_keys, temperature, topp, 0, PeerIn{}};
    hs[i] = BatchSeq{s->key_cache, s->value_cache, s->logits, s->ctrl, 0, 0};
    if (s->fence && s->fence != b->fence) CK(cudaStreamWaitEvent(b->stream, s->fence->ev, 0));
    if (s->async_pending) {  // logits written by an asynchronous rama_forward on the session's own stream
      CK(cudaEventRecord(s->ev1, s->stream));
      CK(cudaStreamWaitEvent(b->stream, s->ev1, 0));
      s->async_pending = false;
    }
  }
  CK(cudaMemcpyAsync(b->d_sp, b->h_sp, (size_t)n * sizeof(SampleParams), cudaMemcpyHostToDevice, b->stream));
  // d_seqs still describes this batch when sample follows forward; rewrite only if the caller passes other sessions
  CK(cudaMemcpyAsync(b->d_seqs, hs, (size_t)n * sizeof(BatchSeq), cudaMemcpyHostToDevice, b->stream));
  sample_batch_kernel<<<n, kSampleThreads, 0, b->stream>>>(b->d_sp);
  CK(cudaGetLastError());
  batch_collect_kernel<<<1, 64, 0, b->stream>>>(b->d_seqs, b->d_next, n);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(b->h_next, b->d_next, (size_t)2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  for (int i = 0; i < n; ++i) {
    if (b->h_next[2 * i + 1] == 1) return fail(RAMA_E_STATE, "token id outside the vocabulary reached the device step (sequence %d)", i);
    if (b->h_next[2 * i + 1] == 2) return fail(RAMA_E_STATE, "top-p candidate list is empty (the reference panics here, infer.rs:66) (sequence %d)", i);
    next[i] = b->h_next[2 * i];
  }
  return RAMA_OK;
}

extern "C" int rama_batch_sync(rama_batch* b) {
  if (!b) return fail(RAMA_E_INVALID, "NULL batch");
  CK(cudaSetDevice(b->ctx->device));
  CK(cudaStreamSynchronize(b->stream));
  return RAMA_OK;
}

extern "C" int rama_batch_launches_per_step(const rama_batch* b, int32_t* n) {
  if (!b || !n) return fail(RAMA_E_INVALID, "NULL argument");
  *n = b->launches;
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// op level (Device trait)
// ------------------------------------------------------------------------------------------------
extern "C" int rama_dev_alloc(rama_ctx* c, size_t n, float** out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMalloc((void**)out, std::max<size_t>(n, 1) * sizeof(float)));
  // RunState::from_config zero-fills (ram.rs:7-23).  On the op stream (non-blocking: it does not
  // synchronise with the legacy default stream, so a default-stream memset could land after later copies).
  CK(cudaMemsetAsync(*out, 0, std::max<size_t>(n, 1) * sizeof(float), c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}
extern "C" int rama_dev_free(rama_ctx* c, float* p) {
  if (!c) return fail(RAMA_E_INVALID, "NULL ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaFree(p));
  return RAMA_OK;
}
extern "C" int rama_dev_h2d(rama_ctx* c, float* dst, const float* src, size_t n) {
  if (!c || (!dst && n) || (!src && n)) return fail(RAMA_E_INVALID, "NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}
extern "C" int rama_dev_d2h(rama_ctx* c, float* dst, const float* src, size_t n) {
  if (!c || (!dst && n) || (!src && n)) return fail(RAMA_E_INVALID, "NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}
extern "C" int rama_ctx_sync(rama_ctx* c) {
  if (!c) return fail(RAMA_E_INVALID, "NULL ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}

static int ew_grid(const rama_ctx* c, size_t n) {
  return (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8));
}
#define OP_PRE(c)                                         \
  if (!(c)) return fail(RAMA_E_INVALID, "NULL ctx");      \
  CK(cudaSetDevice((c)->device));

extern "C" int rama_op_array_add(rama_ctx* c, float* t, const float* s, size_t n) {
  OP_PRE(c);
  if (n) op_array_add_kernel<<<ew_grid(c, n), 256, 0, c->op_stream>>>(t, s, n);
  CK(cudaGetLastError());
  return RAMA_OK;
}
extern "C" int rama_op_array_mult(rama_ctx* c, float* t, const float* s, size_t n) {
  OP_PRE(c);
  if (n) op_array_mult_kernel<<<ew_grid(c, n), 256, 0, c->op_stream>>>(t, s, n);
  CK(cudaGetLastError());
  return RAMA_OK;
}
extern "C" int rama_op_sinu(rama_ctx* c, float* o, size_t n) {
  OP_PRE(c);
  if (n) op_sinu_kernel<<<ew_grid(c, n), 256, 0, c->op_stream>>>(o, n);
  CK(cudaGetLastError());
  return RAMA_OK;
}
extern "C" int rama_op_copy_from_slice(rama_ctx* c, float* t, const float* s, size_t n) {
  OP_PRE(c);
  if (n) op_copy_kernel<<<ew_grid(c, n), 256, 0, c->op_stream>>>(t, s, n);
  CK(cudaGetLastError());
  return RAMA_OK;
}
extern "C" int rama_op_rmsnorm(rama_ctx* c, float* o, const float* x, const float* w, size_t n) {
  OP_PRE(c);
  if (!n) return fail(RAMA_E_INVALID, "rmsnorm of an empty vector");
  op_rmsnorm_kernel<<<1, 1024, 0, c->op_stream>>>(o, x, w, (int)n);
  CK(cudaGetLastError());
  return RAMA_OK;
}
extern "C" int rama_op_apply_position(rama_ctx* c, float* q, float* k, const float* pr, const float* pi,
                                      size_t head_size) {
  OP_PRE(c);
  const int hs2 = (int)(head_size / 2);
  if (hs2) op_apply_position_kernel<<<(hs2 + 127) / 128, 128, 0, c->op_stream>>>(q, k, pr, pi, hs2);
  CK(cudaGetLastError());
  return RAMA_OK;
}
extern "C" int rama_op_softmax(rama_ctx* c, float* x, size_t n) {
  OP_PRE(c);
  if (!n) return fail(RAMA_E_INVALID, "softmax of an empty vector");
  op_softmax_kernel<<<1, 1024, 0, c->op_stream>>>(x, (int)n);
  CK(cudaGetLastError());
  return RAMA_OK;
}

static int matvec(rama_ctx* c, float* o, const float* a, const float* b, size_t width, size_t o_rows,
                  int variant, cudaStream_t st) {
  if (width % 4) return fail(RAMA_E_INVALID, "width %% 4 != 0 (the reference steps k by 4, cpu.rs:142)");
  if (((uintptr_t)a | (uintptr_t)b) & 15) return fail(RAMA_E_INVALID, "matmul operands must be 16-byte aligned");
  ProPlain pro{b};
  RowsPlain rows{a, (int)width, (int)o_rows};
  EpiStore epi{o, (int)o_rows, PeerOut{}};
  const int np = (int)((o_rows + 1) / 2), K4 = (int)(width / 4);
  const int var = variant >= 0 ? variant : pick_variant(c, K4);
  cudaError_t e = launch_gemv(var, pick_grid(c, var, np), st, 0, pro, rows, epi, K4, np);
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "gemv launch: %s", cudaGetErrorString(e));
  CK(cudaGetLastError());
  return RAMA_OK;
}

extern "C" int rama_op_matmul(rama_ctx* c, float* o, const float* a, const float* b, size_t width,
                              size_t o_rows, size_t o_cols) {
  OP_PRE(c);
  if (!width || !o_rows || !o_cols) return fail(RAMA_E_INVALID, "empty matmul");
  if (o_cols == 1) return matvec(c, o, a, b, width, o_rows, -1, c->op_stream);
  const size_t n = o_rows * o_cols;
  op_matmul_general_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->op_stream>>>(o, a, b, (int)width, (int)o_rows, (int)o_cols);
  CK(cudaGetLastError());
  return RAMA_OK;
}

extern "C" int rama_op_multi_head_attention(rama_ctx* c, float* xb, float* att, const float* q,
                                            const float* key_cache, const float* value_cache,
                                            const rama_config* cfg, int32_t layer, int32_t pos) {
  OP_PRE(c);
  if (!cfg) return fail(RAMA_E_INVALID, "NULL cfg");
  const int D = cfg->dim, H = cfg->n_heads, T = cfg->seq_len;
  if (H <= 0 || D % H) return fail(RAMA_E_INVALID, "dim %% n_heads != 0");
  const int hs = D / H;
  if (hs % 4 || hs > kAttnMaxHs) return fail(RAMA_E_INVALID, "head_size %d unsupported", hs);
  if (pos < 0 || pos >= T || layer < 0 || layer >= cfg->n_layers) return fail(RAMA_E_STATE, "layer/pos out of range");
  const int n_split = (T + kAttnChunk - 1) / kAttnChunk;
  float* ws = nullptr;
  unsigned int* tickets = nullptr;
  CK(cudaMallocAsync((void**)&ws, (size_t)H * n_split * (hs + 2) * sizeof(float), c->op_stream));
  CK(cudaMallocAsync((void**)&tickets, H * sizeof(unsigned int), c->op_stream));
  CK(cudaMemsetAsync(tickets, 0, H * sizeof(unsigned int), c->op_stream));
  const size_t lo = (size_t)layer * T * D;
  AttnParams ap{q, key_cache + lo, value_cache + lo, xb, att, ws, tickets, nullptr, pos, T, D, hs, n_split, nullptr, 0};
  attn_decode_kernel<<<dim3(H, n_split), kAttnThreads, 0, c->op_stream>>>(ap, 0);
  CK(cudaGetLastError());
  CK(cudaFreeAsync(ws, c->op_stream));
  CK(cudaFreeAsync(tickets, c->op_stream));
  return RAMA_OK;
}

extern "C" int rama_op_sample(rama_ctx* c, float* logits, size_t vocab_size, float temperature, float topp,
                              int32_t* next) {
  OP_PRE(c);
  if (!logits || !next || vocab_size < 2) return fail(RAMA_E_INVALID, "bad argument");
  size_t vp2 = 1;
  while (vp2 < vocab_size) vp2 <<= 1;
  StepCtrl* ctrl = nullptr;
  unsigned long long* keys = nullptr;
  CK(cudaMallocAsync((void**)&ctrl, sizeof(StepCtrl), c->op_stream));
  CK(cudaMemsetAsync(ctrl, 0, sizeof(StepCtrl), c->op_stream));
  CK(cudaMallocAsync((void**)&keys, vp2 * sizeof(unsigned long long), c->op_stream));
  SampleParams sp{logits, nullptr, 0, 0, (int)vocab_size, ctrl, nullptr, nullptr, keys, temperature, topp, 0, PeerIn{}};
  sample_kernel<<<1, kSampleThreads, 0, c->op_stream>>>(sp, 0);
  CK(cudaGetLastError());
  int32_t ret[2] = {0, 0};
  CK(cudaMemcpyAsync(ret, &ctrl->next, sizeof(ret), cudaMemcpyDeviceToHost, c->op_stream));
  CK(cudaFreeAsync(ctrl, c->op_stream));
  CK(cudaFreeAsync(keys, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  if (ret[1] == 2) return fail(RAMA_E_STATE, "top-p candidate list is empty (the reference panics here, infer.rs:66)");
  *next = ret[0];
  return RAMA_OK;
}

extern "C" int rama_synth_fill(rama_ctx* c, float* dst, size_t n, uint64_t seed, uint64_t tensor_id,
                               uint64_t start, float scale, float offset) {
  OP_PRE(c);
  unsigned long long z = seed ^ ((unsigned long long)tensor_id * 0xD1B54A32D192ED03ull);
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  ShardMap m{1, n, 0, 1, 0, n};
  if (n) synth_fill_kernel<<<c->sm_count * 8, 256, 0, c->op_stream>>>(dst, n, z, m, start, scale, offset);
  CK(cudaGetLastError());
  return RAMA_OK;
}

extern "C" int rama_bench_gemv(rama_ctx* c, float* o, const float* w, const float* x, size_t rows, size_t width,
                               size_t n_mats, int variant, int iters, float* avg_ms) {
  OP_PRE(c);
  if (!avg_ms || iters <= 0 || n_mats == 0) return fail(RAMA_E_INVALID, "bad argument");
  if (variant >= kNumVariants) return fail(RAMA_E_INVALID, "variant %d out of range", variant);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) RK(matvec(c, o, w + (i % n_mats) * rows * width, x, width, rows, variant, c->op_stream));
  CK(cudaEventRecord(a, c->op_stream));
  for (int i = 0; i < iters; ++i)
    RK(matvec(c, o, w + ((i + 3) % n_mats) * rows * width, x, width, rows, variant, c->op_stream));
  CK(cudaEventRecord(b, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  *avg_ms = ms / iters;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return RAMA_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor-core contraction (tcgen05, 3xTF32): C[M][N] = A[M][K] · B[N][K]^T
// ------------------------------------------------------------------------------------------------
extern "C" int rama_op_matmul_nt(rama_ctx* c, float* out, const float* a, const float* b, size_t M, size_t N,
                                 size_t K, int variant, int flags) {
  OP_PRE(c);
  if (!out || !a || !b || !M || !N || !K) return fail(RAMA_E_INVALID, "empty matmul_nt");
  if (K % 4) return fail(RAMA_E_INVALID, "K %% 4 != 0 (TMA needs 16-byte row pitch; the reference steps k by 4, cpu.rs:142)");
  const int hi_round = flags & 1;
  const bool transposed = (flags & 2) != 0;
  const int ksplit = std::max(1, (flags >> 8) & 0xff);
  GemmOperand A{a, M, K}, B{b, N, K};
  cudaError_t e;
  const int m = (int)M, n = (int)N, k = (int)K;
  if (transposed) {
    // split-K partials land in a scratch [ksplit][N][M]; a reduction kernel sums them into out
    float* dst = out;
    if (ksplit > 1) CK(cudaMallocAsync((void**)&dst, (size_t)ksplit * N * M * sizeof(float), c->op_stream));
    EpiStoreT epi{dst, m, n, ksplit, (size_t)N * M};
    switch (variant) {
      case 0: e = launch_gemm_tf32x3<64, 4, 2, 2>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, ksplit, epi); break;
      case 1: e = launch_gemm_tf32x3<64, 4, 4, 2>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, ksplit, epi); break;
      case 2: e = launch_gemm_tf32x3<64, 4, 2, 0>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, ksplit, epi); break;
      case 3: e = launch_gemm_tf32x3<64, 2, 2, 0>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, ksplit, epi); break;  // 2 CTAs/SM
      case 4: e = launch_gemm_tf32x3<64, 2, 4, 0>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, ksplit, epi); break;  // + 128-k chunks
      case 5: e = launch_gemm_tf32x3<64, 2, 2, 0, 4>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, ksplit, epi); break;  // decoupled A ring
      default: return fail(RAMA_E_INVALID, "matmul_nt: unknown transposed variant %d", variant);
    }
    if (ksplit > 1) {
      if (e == cudaSuccess) {
        sum_partials_kernel<<<c->sm_count * 4, 256, 0, c->op_stream>>>(out, dst, (size_t)N * M, ksplit);
        e = cudaGetLastError();
      }
      CK(cudaFreeAsync(dst, c->op_stream));
    }
  } else {
    if (ksplit > 1) return fail(RAMA_E_INVALID, "matmul_nt: split-K only in the transposed (batched decode) orientation");
    EpiStoreNT epi{out, n, n, 0};
    switch (variant) {
      case 0: e = launch_gemm_tf32x3<128, 2, 4, 1>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, 1, epi); break;
      case 1: e = launch_gemm_tf32x3<128, 4, 4, 0>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, 1, epi); break;
      case 2: e = launch_gemm_tf32x3<128, 2, 2, 1>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, 1, epi); break;
      case 3: e = launch_gemm_tf32x3<64, 4, 2, 2>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, 1, epi); break;
      default: return fail(RAMA_E_INVALID, "matmul_nt: unknown variant %d", variant);
    }
  }
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "gemm_tf32x3 launch: %s", cudaGetErrorString(e));
  return RAMA_OK;
}

// Debug hook (tools/gemm_trace.py): per-role clock64() timeline of CTA (0,0,0) of one GEMM launch, [128 k-blocks][8 events].
extern "C" int rama_debug_gemm_trace(rama_ctx* c, float* out, const float* a, const float* b, size_t M, size_t N, size_t K,
                                     int variant, int flags, long long* host_trace) {
  OP_PRE(c);
  if (!host_trace) return fail(RAMA_E_INVALID, "NULL trace");
  long long* d = nullptr;
  CK(cudaMalloc((void**)&d, 128 * 8 * sizeof(long long)));
  CK(cudaMemset(d, 0, 128 * 8 * sizeof(long long)));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 3; ++i) RK(rama_op_matmul_nt(c, out, a, b, M, N, K, variant, flags));  // warm
  g_gemm_trace = d;
  int rc = rama_op_matmul_nt(c, out, a, b, M, N, K, variant, flags);
  g_gemm_trace = nullptr;
  cudaStreamSynchronize(c->op_stream);
  if (rc == RAMA_OK) cudaMemcpy(host_trace, d, 128 * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}

// Micro-benchmark hook (tools/gemm_sweep.py): average milliseconds of rama_op_matmul_nt over `iters` launches.
extern "C" int rama_bench_matmul_nt(rama_ctx* c, float* out, const float* a, const float* b, size_t M, size_t N,
                                    size_t K, int variant, int flags, int iters, float* avg_ms) {
  OP_PRE(c);
  if (!avg_ms || iters <= 0) return fail(RAMA_E_INVALID, "bad argument");
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) RK(rama_op_matmul_nt(c, out, a, b, M, N, K, variant, flags));
  CK(cudaEventRecord(e0, c->op_stream));
  for (int i = 0; i < iters; ++i) RK(rama_op_matmul_nt(c, out, a, b, M, N, K, variant, flags));
  CK(cudaEventRecord(e1, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  *avg_ms = ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return RAMA_OK;
}
