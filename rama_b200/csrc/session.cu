// session.cu — the session (RunState in HBM, per-session stream, captured step graphs), the fused decode step and the
// per-session entry points: forward, sample, generate, profile, to_host.
#include "internal.cuh"

#include "attention.cuh"
#include "step_kernel.cuh"

#include <array>

// ------------------------------------------------------------------------------------------------
// peer-addressable blocks (NVLink peer memory)
// ------------------------------------------------------------------------------------------------
int peer_block_alloc(rama_ctx* c, size_t bytes, PeerBlock* b) {
  if (c->world > kMaxPeers) return fail(RAMA_E_INVALID, "tp world %d > %d", c->world, kMaxPeers);
  b->bytes = (bytes + 255) / 256 * 256;
  CK(cudaMalloc((void**)&b->local, b->bytes));
  CK(cudaMemset(b->local, 0, b->bytes));  // epoch 0 is never used
  CK(cudaDeviceSynchronize());
  for (int r = 0; r < kMaxPeers; ++r) b->base[r] = nullptr;
  b->base[c->rank] = b->local;
  return RAMA_OK;
}

// Processes of a torchrun launch: swap CUDA IPC handles through NCCL and map every peer's block.  Collective.
int peer_block_connect(rama_ctx* c, PeerBlock* b, cudaStream_t st) {
  if (c->group) return RAMA_OK;  // single-process group: the group cross-wires its ranks' blocks (peer_blocks_connect_group)
  if (c->tp_sim) {               // measurement mode: every "peer" is this device
    for (int r = 0; r < c->world; ++r) b->base[r] = b->local;
    return RAMA_OK;
  }
  const int P = c->world;
  cudaIpcMemHandle_t mine;
  CK(cudaIpcGetMemHandle(&mine, b->local));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  char* d_h = nullptr;
  CK(cudaMalloc((void**)&d_h, 64 * (size_t)P + 64));
  CK(cudaMemcpy(d_h + 64 * (size_t)c->rank, &mine, 64, cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());  // `st` is non-blocking: make the staged copy land first
  int e = g_nccl.AllGather(d_h + 64 * (size_t)c->rank, d_h, 16, kNcclFloat32, c->comm, st);
  if (e) { cudaFree(d_h); return fail(RAMA_E_NCCL, "handle all-gather: %s", g_nccl.GetErrorString(e)); }
  CK(cudaStreamSynchronize(st));
  std::vector<cudaIpcMemHandle_t> all(P);
  CK(cudaMemcpy(all.data(), d_h, 64 * (size_t)P, cudaMemcpyDeviceToHost));
  for (int r = 0; r < P; ++r) {
    if (r == c->rank) continue;
    void* p = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess) {
      cudaFree(d_h);
      return fail(RAMA_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (set RAMA_TP_COMM=nccl to fall back)", r, cudaGetErrorString(ce));
    }
    b->base[r] = (char*)p;
  }
  b->ipc = true;
  // nobody may write into a peer block before its owner has zeroed it: barrier
  e = g_nccl.AllReduce(d_h + 64 * (size_t)P, d_h + 64 * (size_t)P, 1, kNcclFloat32, kNcclSum, c->comm, st);
  if (e) { cudaFree(d_h); return fail(RAMA_E_NCCL, "barrier: %s", g_nccl.GetErrorString(e)); }
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(d_h));
  return RAMA_OK;
}

void peer_blocks_connect_group(PeerBlock* const* blocks, int P) {
  for (int r = 0; r < P; ++r)
    for (int q = 0; q < P; ++q) blocks[r]->base[q] = blocks[q]->local;
}

// Collective between processes: every rank must have unmapped the block before its owner frees it.  (A group frees its
// ranks' blocks only after every device has been synchronised.)
void peer_block_free(rama_ctx* c, PeerBlock* b, cudaStream_t st, float* scratch) {
  if (!b->local) return;
  if (b->ipc) {
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && b->base[r]) cudaIpcCloseMemHandle(b->base[r]);
    if (c->comm && scratch) {
      g_nccl.AllReduce(scratch, scratch, 1, kNcclFloat32, kNcclSum, c->comm, st);
      cudaStreamSynchronize(st);
    }
  }
  cudaFree(b->local);
  *b = PeerBlock{};
}

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
static void session_free(rama_session* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (auto& gm : s->g) for (auto& g : gm) if (g) cudaGraphExecDestroy(g);
  if (s->p2p) {
    rama_ctx* c = s->ctx;
    peer_block_free(c, &s->pf_blk, s->stream, s->xb2);
    peer_block_free(c, &s->blk, s->stream, s->xb2);
    s->logits = nullptr; s->x0 = nullptr;  // lived inside the block
    s->pf_xn = nullptr;
  }
  void* bufs[] = {s->x0, s->x1, s->xfinal, s->xb, s->xb2, s->w2out, s->hb, s->hb2, s->q, s->k, s->v, s->att,
                  s->logits, s->key_cache, s->value_cache, s->attn_ws, s->tickets, s->part, s->sort_keys,
                  s->ctrl, s->d_prompt, s->d_out, s->seq, s->bar, s->wo_part, s->red_ll, s->pf_x, s->pf_xn, s->pf_q, s->pf_att, s->pf_y, s->pf_h,
                  s->pf_tokens};
  for (void* b : bufs) if (b) cudaFree(b);
  if (s->h_ring) cudaFreeHost(s->h_ring);
  if (s->h_ret) cudaFreeHost(s->h_ret);
  if (s->h_tokens) cudaFreeHost(s->h_tokens);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

// The session's exchange block.  Layout (LL elements = {payload, epoch} uint2):
//   uint2 parts[P][SMs][2] | uint2 inbox[2 stages][P][D] | float logits[V] | float x0[D] | unsigned flags[3][P] | unsigned done
static int setup_peer_exchange(rama_session* s, bool connect) {
  rama_ctx* c = s->ctx;
  const int P = c->world;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  s->off_parts = 0;
  s->off_inbox = al((size_t)P * c->sm_count * 2 * sizeof(uint2));
  s->off_logits = al(s->off_inbox + (size_t)2 * P * c->D * sizeof(uint2));
  s->off_x0 = al(s->off_logits + (size_t)c->V * sizeof(float));
  s->off_flags = al(s->off_x0 + (size_t)c->D * sizeof(float));
  s->off_done = al(s->off_flags + (size_t)3 * P * sizeof(unsigned));
  RK(peer_block_alloc(c, s->off_done + 256, &s->blk));
  s->logits = reinterpret_cast<float*>(s->blk.local + s->off_logits);
  s->x0 = reinterpret_cast<float*>(s->blk.local + s->off_x0);
  if (connect) RK(peer_block_connect(c, &s->blk, s->stream));
  return RAMA_OK;
}

static int session_create_rank(rama_ctx* c, rama_session** out, bool connect) {
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
  s->p2p = c->world > 1 && c->p2p;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
#define A(call) if (e == cudaSuccess) e = (call)
  if (!s->p2p) { A(dalloc(&s->x0, D)); A(dalloc(&s->logits, V)); }  // p2p: inside the exchange block (peers store into them)
  A(dalloc(&s->x1, D)); A(dalloc(&s->xfinal, D));
  A(dalloc(&s->xb, Dq)); A(dalloc(&s->xb2, D)); A(dalloc(&s->w2out, D));
  A(dalloc(&s->hb, Fl)); A(dalloc(&s->hb2, Fl));
  A(dalloc(&s->q, Dq)); A(dalloc(&s->k, Dq)); A(dalloc(&s->v, Dq));
  A(dalloc(&s->att, (size_t)c->Hl * T));
  A(dalloc(&s->key_cache, L * T * Dq)); A(dalloc(&s->value_cache, L * T * Dq));
  A(dalloc(&s->attn_ws, (size_t)c->Hl * s->n_split * (c->hs + 2)));
  A(dalloc(&s->tickets, (size_t)c->Hl));
  s->persistent = c->persistent && !c->group && (c->world == 1 || s->p2p);
  if (s->persistent) s->cls_grid = c->sm_count;  // every CTA of the persistent kernel writes a classifier partial
  // Tiny models on one GPU: the whole stack of layers fits L2 several times over and a token is ~30 dependent kernels of ~3 µs
  // each for ~1 µs of streaming.  The layers then run as ONE kernel whose CTAs form a single cluster (step_kernel.cuh, cluster
  // mode).  Worth it only while a handful of SMs can stream the layers from L2 faster than the kernel chain runs.
  {
    const size_t layer_bytes = (size_t)c->L * (4 * D * D + 3 * D * c->F) * sizeof(float);
    const bool fits = c->world == 1 && !s->persistent && layer_bytes <= ((size_t)env_int("RAMA_STEP_CLUSTER_MAX_MB", 32) << 20);
    s->cluster_step = fits && c->cluster_step != 0 && (c->cluster_step == 1 || env_int("RAMA_STEP_CLUSTER_AUTO", 0));
  }
  A(dalloc(&s->part, (size_t)c->world * c->sm_count));
  if (s->p2p) A(dalloc(&s->red_ll, (size_t)2 * D));
  A(dalloc(&s->seq, 1));
  A(dalloc(&s->bar, 2));
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
    int rc = setup_peer_exchange(s, connect);
    if (rc != RAMA_OK) {
      session_free(s);
      return rc;
    }
  }
  c->n_objects.fetch_add(1);
  *out = s;
  return RAMA_OK;
}

static void session_destroy_rank(rama_session* s) {
  rama_ctx* c = s->ctx;
  std::lock_guard<std::mutex> cap_lk(c->cap_mu);
  if (s->fence) { cudaSetDevice(c->device); cudaEventSynchronize(s->fence->ev); s->fence.reset(); }
  session_free(s);
  c->n_objects.fetch_sub(1);
}

// every stream of every rank idle: nothing is in flight towards a peer block any more
static void group_quiesce(rama_session* g) {
  for (rama_session* r : g->ranks) {
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->stream);
  }
}

extern "C" int rama_session_create(rama_ctx* c, rama_session** out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!c->loaded) return fail(RAMA_E_STATE, "no weights loaded");
  if (c->ranks.empty()) return session_create_rank(c, out, true);
  // single-process group: one session per rank, exchange blocks cross-wired directly
  rama_session* g = new rama_session();
  g->ctx = c;
  for (rama_ctx* rc : c->ranks) {
    rama_session* rs = nullptr;
    int rc_ = session_create_rank(rc, &rs, false);
    if (rc_ != RAMA_OK) {
      for (rama_session* x : g->ranks) session_destroy_rank(x);
      delete g;
      return rc_;
    }
    rs->parent = g;
    g->ranks.push_back(rs);
  }
  std::vector<PeerBlock*> blocks;
  for (rama_session* r : g->ranks) blocks.push_back(&r->blk);
  peer_blocks_connect_group(blocks.data(), (int)blocks.size());
  c->n_objects.fetch_add(1);
  *out = g;
  return RAMA_OK;
}

extern "C" int rama_session_destroy(rama_session* s) {
  if (!s) return RAMA_OK;
  if (!s->ranks.empty()) {
    group_quiesce(s);
    for (rama_session* r : s->ranks) session_destroy_rank(r);
    s->ctx->n_objects.fetch_sub(1);
    delete s;
    return RAMA_OK;
  }
  session_destroy_rank(s);
  return RAMA_OK;
}

extern "C" int rama_session_sync(rama_session* s) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  if (!s->ranks.empty()) return group_run(s->ctx, [&](int r) { return rama_session_sync(s->ranks[r]); });
  CK(cudaSetDevice(s->ctx->device));
  CK(session_enter(s));
  CK(cudaStreamSynchronize(s->stream));
  s->async_pending = false;
  return RAMA_OK;
}

extern "C" int rama_session_reset(rama_session* s) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  if (!s->ranks.empty()) return group_run(s->ctx, [&](int r) { return rama_session_reset(s->ranks[r]); });
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
  if (!s->ranks.empty()) return group_run(s->ctx, [&](int r) { return rama_session_set_debug(s->ranks[r], keep_att); });
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
  // in-graph timeline (rama_step_timeline): every kernel of the step gets a 4-stamp slot of this device buffer
  unsigned long long* tl = nullptr;
  std::vector<int>* tl_kinds = nullptr;
  unsigned long long* slot(int kind) {
    if (!tl) return nullptr;
    tl_kinds->push_back(kind);
    return tl + 4 * (tl_kinds->size() - 1);
  }
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
static StepParams step_params(rama_session* s, int mode, long long* trace) {
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
  for (int r = 0; r < kMaxPeers; ++r) p.peer_base[r] = r < c->world ? s->blk.base[r] : nullptr;
  p.off_inbox = s->off_inbox; p.off_parts = s->off_parts;
  return p;
}

// The LAYERS of the step as one cluster-scope launch (tiny models); the classifier and the sampler stay separate kernels.
static cudaError_t launch_cluster_layers(rama_session* s, cudaStream_t st, int pdl) {
  rama_ctx* c = s->ctx;
  StepParams p = step_params(s, 0, nullptr);
  p.cluster = 1;
  int cs = c->cluster_step_ctas;
  static std::atomic<unsigned long long> attr_done{0}, np_done{0};
  cudaError_t e = ensure_dyn_smem((const void*)decode_step_kernel, (int)kMaxDynSmem, attr_done);
  if (e != cudaSuccess) return e;
  if (cs > 8) {  // non-portable cluster size: opt in once per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(np_done.load() & (1ull << dev))) {
      if (cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) np_done.fetch_or(1ull << dev);
      else { cudaGetLastError(); cs = 8; }
    }
  }
  for (;; cs /= 2) {  // largest cluster the device can co-schedule
    size_t smem = 0;
    auto need = [&](int K4, int n_pairs, int wk) { smem = std::max(smem, gemv_smem_bytes(K4, n_pairs, cs, wk)); };
    need(c->D / 4, 3 * c->Dq / 2, p.wk_d); need(c->Dq / 4, c->D / 2, p.wk_wo); need(c->D / 4, c->Fl, p.wk_d);
    need(c->Fl / 4, c->D / 2, p.wk_w2);
    if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(kGemvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    int n_active = 0;
    cfg.attrs = at; cfg.numAttrs = na;
    if (cudaOccupancyMaxActiveClusters(&n_active, decode_step_kernel, &cfg) != cudaSuccess || n_active < 1) {
      cudaGetLastError();
      if (cs > 2) continue;
      return cudaErrorInvalidConfiguration;
    }
    (void)pdl;  // (the kernel reads ctrl at entry: it keeps a full stream dependency on the previous step's sampler)
    return cudaLaunchKernelEx(&cfg, decode_step_kernel, p);
  }
}

static int enqueue_step_persistent(rama_session* s, cudaStream_t st, int mode, int* n_launch, long long* trace = nullptr) {
  rama_ctx* c = s->ctx;
  StepParams p = step_params(s, mode, trace);
  const int grid = c->sm_count;
  size_t smem = 0;
  auto need = [&](int K4, int n_pairs, int wk) { smem = std::max(smem, gemv_smem_bytes(K4, n_pairs, grid, wk)); };
  need(c->D / 4, 3 * c->Dq / 2, p.wk_d); need(c->Dq / 4, c->D / 2, p.wk_wo); need(c->D / 4, c->Fl, p.wk_d);
  need(c->Fl / 4, c->D / 2, p.wk_w2); need(c->D / 4, (c->Vl + 1) / 2, p.wk_d);
  if (smem > kMaxDynSmem) return fail(RAMA_E_INVALID, "persistent step: %zu bytes of shared memory needed", smem);
  static std::atomic<unsigned long long> attr_done{0};
  const cudaError_t attr_err = ensure_dyn_smem((const void*)decode_step_kernel, (int)kMaxDynSmem, attr_done);
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
static int enqueue_step(rama_session* s, cudaStream_t st, int mode, StepTrace* tr, int* n_launch,
                        unsigned long long* tl = nullptr, std::vector<int>* tl_kinds = nullptr) {
  rama_ctx* c = s->ctx;
  if (s->persistent && !tr) return enqueue_step_persistent(s, st, mode, n_launch);
  StepEnq q{s, st, tr};
  q.tl = tl; q.tl_kinds = tl_kinds;
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
  // tiny models: embedding gather + every layer in ONE cluster-scope kernel (launch_cluster_layers); not for the event-timed /
  // stamped variants of the step, which describe the multi-kernel chain
  const bool cluster_layers = s->cluster_step && !tr && !tl && !s->keep_att;
  if (cluster_layers) {
    q.pre(RAMA_K_QKV);
    q.post(launch_cluster_layers(s, st, q.pdl));
  }
  if (embed_kernel && !cluster_layers) {
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
    q.post(cudaLaunchKernelEx(&cfg, step_begin_kernel, s->ctrl, s->seq, W[RAMA_T_TOKEN_EMBEDDING], s->x0, D, c->V, q.pdl, q.slot(RAMA_K_EMBED)));
  }

  for (int l = 0; l < (cluster_layers ? 0 : L); ++l) {
    // ---- rmsnorm → [wq|wk|wv] → RoPE → KV write (infer.rs:19-33) ----
    {
      ProNorm pro{s->x0, l == 0 ? nullptr : s->w2out, s->x1, W[RAMA_T_RMS_ATT] + (size_t)l * D, nullptr, peer_in(s, 1, l - 1)};
      if (l == 0 && !embed_kernel) { pro.emb = W[RAMA_T_TOKEN_EMBEDDING]; pro.ctrl = s->ctrl; pro.seq = s->seq; pro.vocab = c->V; }
      const int nowait = l > 0 ? set_peer_reduce(s, pro, 1, q.pdl) : 0;
      const int want_cluster = (s->p2p && l > 0 && tp_reduce_mode(s) == 1) ? tp_cluster_size(c) : 0;
      RowsQKV rows{W[RAMA_T_WQ] + (size_t)l * Dq * D, W[RAMA_T_WK] + (size_t)l * Dq * D,
                   W[RAMA_T_WV] + (size_t)l * Dq * D, D, Dq / 2};
      EpiQKV epi{s->q, s->k, s->v, s->key_cache + (size_t)l * T * Dq, s->value_cache + (size_t)l * T * Dq,
                 W[RAMA_T_FREQ_REAL], W[RAMA_T_FREQ_IMAG], s->ctrl, Dq / 2, hs / 2, Dq};
      const int np = 3 * Dq / 2, var = pick_variant(c, D / 4, np);
      q.pre(RAMA_K_QKV);
      // the cluster attention kernel reads older K/V rows ahead of its wait: release it after this kernel's own wait
      const int pdl_flags = (q.pdl ? ((attn_cluster || fuse_cluster) ? 3 : 1) : 0) | nowait;
      q.post(launch_gemv(var, pick_grid(c, var, np), st, pdl_flags, pro, rows, epi, D / 4, np, want_cluster, q.slot(RAMA_K_QKV)));
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
      unsigned long long* tslot = q.slot(RAMA_K_ATTN);
      if (attn_cluster) q.post(cudaLaunchKernelEx(&cfg, attn_cluster_kernel, ap, q.pdl, tslot));
      else q.post(cudaLaunchKernelEx(&cfg, attn_decode_kernel, ap, q.pdl, tslot));
    }
    // ---- wo (infer.rs:35); the residual add (:37) is folded into the next prologue ----
    {
      ProPlain pro{s->xb};
      RowsPlain rows{W[RAMA_T_WO] + (size_t)l * D * Dq, Dq, D};
      EpiStore epi{s->xb2, D, peer_out(s, 0, l)};
      const int np = D / 2, var = pick_variant(c, Dq / 4, np);
      q.pre(RAMA_K_WO);
      q.post(launch_gemv(var, pick_grid(c, var, np), st, q.pdl, pro, rows, epi, Dq / 4, np, 0, q.slot(RAMA_K_WO)));
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
      const int nowait = set_peer_reduce(s, pro, 0, q.pdl);
      RowsW13 rows{W[RAMA_T_W1] + (size_t)l * Fl * D, W[RAMA_T_W3] + (size_t)l * Fl * D, D};
      EpiSwiGLU epi{s->hb, s->hb2};
      const int np = Fl, var = pick_variant(c, D / 4, np);
      q.pre(RAMA_K_W13);
      q.post(launch_gemv(var, pick_grid(c, var, np), st, q.pdl | nowait, pro, rows, epi, D / 4, np, (s->p2p && tp_reduce_mode(s) == 1) ? tp_cluster_size(c) : 0, q.slot(RAMA_K_W13)));
    }
    // ---- w2 (infer.rs:46); residual add (:47) folded into the next prologue ----
    {
      ProPlain pro{s->hb};
      RowsPlain rows{W[RAMA_T_W2] + (size_t)l * D * Fl, Fl, D};
      EpiStore epi{s->w2out, D, peer_out(s, 1, l)};
      const int np = D / 2, var = pick_variant(c, Fl / 4, np);
      q.pre(RAMA_K_W2);
      q.post(launch_gemv(var, pick_grid(c, var, np), st, q.pdl, pro, rows, epi, Fl / 4, np, 0, q.slot(RAMA_K_W2)));
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
    const int nowait = set_peer_reduce(s, pro, 1, q.pdl);
    RowsPlain rows{c->wcls, D, c->Vl};
    EpiCls epi = make_epi_cls(s);
    const int np = (c->Vl + 1) / 2, var = pick_variant(c, D / 4);
    cls_grid = pick_grid(c, var, np);
    q.pre(RAMA_K_CLS);
    q.post(launch_gemv(var, cls_grid, st, q.pdl | nowait, pro, rows, epi, D / 4, np, 0, q.slot(RAMA_K_CLS)));  // (no cluster launch: every one of the cls_grid partial slots must be written)
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
    if (mode == 2 && !logits_pushed(s)) {
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
    q.post(cudaLaunchKernelEx(&cfg, sample_kernel, sp, q.pdl, q.slot(RAMA_K_SAMPLE)));
  }
  if (n_launch) *n_launch = q.launches;
  if (q.nccl_err) return fail(RAMA_E_NCCL, "nccl collective in step: %s", g_nccl.GetErrorString(q.nccl_err));
  if (q.err != cudaSuccess) return fail(RAMA_E_CUDA, "kernel launch in step: %s", cudaGetErrorString(q.err));
  return RAMA_OK;
}

// unused partial slots (a CTA-less tail when the classifier grid < sm_count) must read as "empty"
static __global__ void fill_parts_kernel(ArgPart* part, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) part[i] = ArgPart{-INFINITY, -1};
}
int init_parts(rama_session* s) {
  rama_ctx* c = s->ctx;
  const int n = c->world * c->sm_count;
  fill_parts_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->part, n);
  CK(cudaGetLastError());
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
  if (!s->ranks.empty()) return rama_session_launches_per_step(s->ranks[0], n);
  const rama_ctx* c = s->ctx;
  if (s->persistent) { *n = 1; return RAMA_OK; }  // the whole greedy step is one persistent cooperative launch
  if (s->cluster_step && !s->keep_att) { *n = 3; return RAMA_OK; }  // layers (one cluster-scope kernel) + classifier + sampler
  // embed + L·(qkv, attn, wo, w13, w2) + cls + sample (+ collectives under TP); attention + wo are one launch
  // for the small models at positions < 256
  *n = (c->embed_kernel ? 1 : 0) + (s->wo_part ? 4 : 5) * c->L + 1 + 1 + (c->world > 1 && !s->p2p ? 2 * c->L + 1 : 0);
  return RAMA_OK;
}

extern "C" int rama_forward(rama_session* s, int32_t token, int32_t pos) {
  if (!s) return fail(RAMA_E_INVALID, "NULL session");
  if (!s->ranks.empty()) return group_run(s->ctx, [&](int r) { return rama_forward(s->ranks[r], token, pos); });
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

// Full vocabulary in s->logits, stream-ordered.  Peer exchange: the classifier epilogues of all ranks store their slices
// into every rank's array, so this only waits for their arrival; NCCL mode: all-gather of the slices.
int gather_logits(rama_session* s) {
  rama_ctx* c = s->ctx;
  if (c->world > 1 && !s->logits_gathered) {
    if (logits_pushed(s)) {
      if (s->parts_valid) {
        peer_parts_wait_kernel<<<1, kSampleThreads, 0, s->stream>>>(peer_in_parts(s), c->world * c->sm_count, s->cls_grid);
        CK(cudaGetLastError());
      }
    } else {
      if (!c->comm) return fail(RAMA_E_STATE, "logits all-gather needs NCCL (persistent step kernel inside a single-process group)");
      NK(g_nccl.AllGather(s->logits + c->v0, s->logits, (size_t)c->Vl, kNcclFloat32, c->comm, s->stream));
    }
    s->logits_gathered = true;
  }
  return RAMA_OK;
}

int read_ret(rama_session* s, int32_t* next) {
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
  // group: every rank holds every rank's greedy partials and the full logits — rank 0 samples (host-driven loop: the
  // token comes back through the next rama_forward on every rank)
  if (!s->ranks.empty()) return rama_sample(s->ranks[0], temperature, topp, next);
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


extern "C" int rama_generate(rama_session* s, const int32_t* prompt, int32_t n_prompt, int32_t steps,
                             float temperature, float topp, int32_t* out_tokens, float* elapsed_ms) {
  if (!s || (n_prompt > 0 && !prompt) || !out_tokens) return fail(RAMA_E_INVALID, "NULL argument");
  if (!s->ranks.empty()) {
    // every rank runs the device-resident loop (each samples the same token from the same exchanged partials); rank 0's
    // stream of tokens and its device time are returned
    if (s->ranks[0]->pf_min > 0 && n_prompt + 1 >= s->ranks[0]->pf_min) RK(group_prefill_prepare(s));
    std::vector<std::vector<int32_t>> scratch(s->ranks.size());
    return group_run(s->ctx, [&](int r) {
      if (r == 0) return rama_generate(s->ranks[0], prompt, n_prompt, steps, temperature, topp, out_tokens, elapsed_ms);
      scratch[r].resize((size_t)std::max(steps, 1));
      return rama_generate(s->ranks[r], prompt, n_prompt, steps, temperature, topp, scratch[r].data(), nullptr);
    });
  }
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
  if (!s->ranks.empty()) {  // rank 0's timings; the other ranks run the same un-graphed step (the exchange needs them)
    std::vector<std::array<float, RAMA_K_COUNT>> m(s->ranks.size());
    std::vector<std::array<int32_t, RAMA_K_COUNT>> l(s->ranks.size());
    return group_run(s->ctx, [&](int r) {
      return r == 0 ? rama_profile_step(s->ranks[0], token, pos, ms, launches)
                    : rama_profile_step(s->ranks[r], token, pos, m[r].data(), l[r].data());
    });
  }
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

// In-graph per-kernel timeline of one decode step.  rama_profile_step times every launch with an event pair on an un-graphed
// stream — launch latency included, no overlap between kernels — and so overstates each kernel by ~20 %; here the step is captured
// into a CUDA graph exactly like the production one (same launch attributes, programmatic dependent launch included) with a
// %globaltimer slot per kernel: CTA 0 of each kernel stamps entry, "dependency resolved" (after griddepcontrol.wait), "prologue
// done" (activations in shared memory; under TP: every peer partial has arrived) and its own end.  The graph is replayed `reps`
// times at the same (token, pos); the stamps of the last replay are returned in nanoseconds relative to the first stamp.
// Under tensor parallelism every rank calls this collectively (a group context fans it out; rank 0's stamps are returned).
extern "C" int rama_step_timeline(rama_session* s, int32_t token, int32_t pos, int32_t mode, int32_t reps, double* stamps_ns,
                                  int32_t* kinds, int32_t cap, int32_t* n_out) {
  if (!s || !stamps_ns || !kinds || !n_out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!s->ranks.empty()) {
    std::vector<std::vector<double>> st(s->ranks.size(), std::vector<double>((size_t)4 * std::max(cap, 1)));
    std::vector<std::vector<int32_t>> kd(s->ranks.size(), std::vector<int32_t>((size_t)std::max(cap, 1)));
    std::vector<int32_t> nn(s->ranks.size(), 0);
    return group_run(s->ctx, [&](int r) {
      return r == 0 ? rama_step_timeline(s->ranks[0], token, pos, mode, reps, stamps_ns, kinds, cap, n_out)
                    : rama_step_timeline(s->ranks[r], token, pos, mode, reps, st[r].data(), kd[r].data(), cap, &nn[r]);
    });
  }
  rama_ctx* c = s->ctx;
  if (s->persistent) return fail(RAMA_E_STATE, "the persistent step kernel has its own phase trace (rama_step_trace)");
  if (pos < 0 || pos >= c->T || token < 0 || token >= c->V) return fail(RAMA_E_INVALID, "token/pos out of range");
  if (mode < 0 || mode > 1 || reps < 1) return fail(RAMA_E_INVALID, "mode must be 0 (forward) or 1 (+ greedy sampler), reps ≥ 1");
  CK(cudaSetDevice(c->device));
  CK(session_enter(s));
  const int max_k = 1 + 5 * c->L + 2;
  if (cap < max_k) return fail(RAMA_E_INVALID, "need room for %d kernels", max_k);
  unsigned long long* d = nullptr;
  CK(cudaMalloc((void**)&d, (size_t)max_k * 4 * sizeof(unsigned long long)));
  CK(cudaMemset(d, 0, (size_t)max_k * 4 * sizeof(unsigned long long)));
  set_attn_bucket(s, pos);
  std::vector<int> kk;
  cudaGraphExec_t ge = nullptr;
  int rc = RAMA_OK;
  {
    std::lock_guard<std::mutex> cap_lk(c->cap_mu);
    rc = init_parts(s);
    cudaGraph_t g = nullptr;
    if (rc == RAMA_OK && cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) rc = fail(RAMA_E_CUDA, "begin capture");
    if (rc == RAMA_OK) {
      int n = 0;
      rc = enqueue_step(s, s->stream, mode, nullptr, &n, d, &kk);
      cudaError_t e = cudaStreamEndCapture(s->stream, &g);
      if (rc == RAMA_OK && e != cudaSuccess) rc = fail(RAMA_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
      if (rc == RAMA_OK && cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) rc = fail(RAMA_E_CUDA, "cudaGraphInstantiate");
      if (g) cudaGraphDestroy(g);
    }
  }
  if (rc == RAMA_OK) {
    for (int i = 0; i < reps && rc == RAMA_OK; ++i) {
      StepCtrl* h = &s->h_ring[s->ring_i];
      if (++s->ring_i == kRing) { s->ring_i = 0; cudaStreamSynchronize(s->stream); }
      memset(h, 0, sizeof(*h));
      h->pos = pos; h->token = token; h->chained = 0;
      if (cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
          cudaGraphLaunch(ge, s->stream) != cudaSuccess)
        rc = fail(RAMA_E_CUDA, "timeline replay: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) rc = fail(RAMA_E_CUDA, "timeline sync: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc == RAMA_OK) {
    std::vector<unsigned long long> hst((size_t)kk.size() * 4);
    if (cudaMemcpy(hst.data(), d, hst.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = fail(RAMA_E_CUDA, "timeline copy");
    unsigned long long t0 = ~0ull;
    for (unsigned long long v : hst) if (v && v < t0) t0 = v;
    for (size_t i = 0; i < hst.size(); ++i) stamps_ns[i] = hst[i] ? (double)(hst[i] - t0) : -1.0;
    for (size_t i = 0; i < kk.size(); ++i) kinds[i] = kk[i];
    *n_out = (int32_t)kk.size();
  }
  if (ge) cudaGraphExecDestroy(ge);
  cudaFree(d);
  s->logits_gathered = false;
  s->parts_valid = true;
  return rc;
}

// Phase timeline of one persistent step: clock64() of CTA 0 at kernel entry and before/after each grid barrier
// (2·(5L+1)+1 stamps).  Tool for tools/step_trace.py.
extern "C" int rama_step_trace(rama_session* s, int32_t token, int32_t pos, long long* stamps, int32_t cap, int32_t* n_out) {
  if (!s || !stamps || !n_out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!s->ranks.empty()) return fail(RAMA_E_STATE, "the persistent step kernel is not used by a single-process group");
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
  if (!s->ranks.empty()) return rama_logits_to_host(s->ranks[0], dst, n);  // every rank holds the full vocabulary
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
  if (!s->ranks.empty()) return rama_state_to_host(s->ranks[0], buf, dst, n, n_out);  // rank 0's shard (logits, x: complete)
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

