// batch_api.cu — batched multi-sequence decode (host side): one step for n sessions (server path, lib.rs:127-160).
#include "internal.cuh"

#include "gemm_host.cuh"
#include "batch.cuh"

// ------------------------------------------------------------------------------------------------
// batched multi-sequence decode (server path, lib.rs:127-160): one step for n sessions
// ------------------------------------------------------------------------------------------------
extern "C" int rama_batch_destroy(rama_batch* b);
constexpr int kBatchMax = 64;    // one 64-column MMA tile of sequences
constexpr int kBatchRing = 8;
// batched-decode GEMM tile: 128 weight rows × 64 sequences, 2 stages, chunks of 2 k-blocks; two CTAs fit an SM
// (64 KB of shared memory and 256 TMEM columns each) and interleave their pipelines
// Round 2 experiment (RAMA_BATCH_PS=1, off by default): pre-split activations (gemm_tf32x3.cuh PS mode) — the producing kernels
// write the tf32 remainder plane next to the activations, the GEMM's workers only split the weight k-blocks, two MMAs per k-step
// instead of three, one CTA per SM with a deep weight ring (kBatchAS × 16 KB in flight per SM).  Parity-green, and measured
// SLOWER than the round-1 tile (10.5 vs 9.2 ms per 64-sequence step): the role timeline shows the single issuer thread as the
// bound (907 clk per k-block, 600 of them issuing 8 MMAs), where two round-1 CTAs per SM reach 762 clk per k-block and SM —
// and the second operand plane doubles the L2 → SM traffic of the activations.  DESIGN.md §4.8.
constexpr int kBatchAS = 8;
#define BATCH_GEMM_R1 launch_gemm_tf32x3<64, 2, 4, 0>
#define BATCH_GEMM_PS launch_gemm_tf32x3<64, 4, 4, 0, kBatchAS, 1>
constexpr int kBatchCtasPerSmR1 = GemmSmem<64, 2, 0>::kCtasPerSm;
constexpr int kBatchCtasPerSmPS = GemmSmem<64, 4, 0, kBatchAS, 1>::kCtasPerSm;
constexpr int kBatchPlane = kBatchMax;  // rows between an activation matrix and its remainder plane

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
  int32_t* d_err = nullptr;          // device error word of the TP exchange (3 = a peer timed out)
  int32_t* h_next = nullptr;         // pinned [2·cap]
  int ring_i = 0;
  std::vector<cudaGraphExec_t> graphs;  // by batch size
  std::vector<cudaGraphExec_t> graphs_chained;  // by batch size: step + samplers of the device-resident loop (rama_generate_batch)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int launches = 0;
  std::shared_ptr<BatchFence> fence;    // re-recorded after every step; the sessions of the step hold a reference
  // tensor parallelism over peer memory (tp_exchange.cuh): inbox[P][S_max][bpr_max][D] | xn[cap][D] | flags[3][P] | done | step counter
  bool p2p = false;
  PeerBlock blk;
  size_t off_xn = 0, off_flags = 0, off_done = 0, off_step = 0;
  int s_max = 1;
  // single-process group batch: one batch per rank
  std::vector<rama_batch*> ranks;
};

// split-K factor: the smallest one that fills ≥ 92 % of the CTA slots of its last wave (every extra split writes and
// re-reads another [n][rows] partial), else the best-filling one; ≥ 8 k-blocks per split
static int batch_ps(const rama_ctx*) {
  static const int v = env_int("RAMA_BATCH_PS", 0);
  return v;
}
static int pick_ksplit(const rama_ctx* c, int tiles, int K) {
  const int total_kb = (K + kGemmBK - 1) / kGemmBK;
  const int slots = c->sm_count * (batch_ps(c) ? kBatchCtasPerSmPS : kBatchCtasPerSmR1);  // CTAs resident at once
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

static int batch_create_rank(rama_ctx* c, int32_t max_seqs, rama_batch** out, bool connect) {
  if (c->tp_sim) return fail(RAMA_E_STATE, "RAMA_TP_SIM measures the decode step only");
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
  b->p2p = c->world > 1 && c->p2p;
  // xn / att / h are B operands of the GEMMs: [2·64][K] — rows 0..63 the activations, rows 64..127 their tf32 remainders
  constexpr size_t R2 = 2 * kBatchPlane;
  A(dalloc(&b->x, B * D)); A(dalloc(&b->q, B * Dq)); A(dalloc(&b->att, R2 * Dq));
  if (!b->p2p) A(dalloc(&b->xn, R2 * D));
  A(dalloc(&b->h, R2 * Fl)); A(dalloc(&b->part, pf));
  if (c->world > 1 && !b->p2p) { A(dalloc(&b->red, B * D)); A(dalloc(&b->lstage, (size_t)c->world * B * c->Vl)); }
  A(dalloc(&b->attn_ws, B * c->Hl * b->n_split * (c->hs + 2)));
  A(dalloc(&b->tickets, B * c->Hl));
  A(dalloc(&b->d_seqs, B)); A(dalloc(&b->d_sp, B)); A(dalloc(&b->d_next, 2 * B)); A(dalloc(&b->d_err, 1));
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
  if (b->p2p) {
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    b->s_max = std::max(pick_ksplit(c, tiles(c->D), c->Dq), pick_ksplit(c, tiles(c->D), c->Fl));
    const size_t bpr_max = (B + c->world - 1) / c->world;
    b->off_xn = al((size_t)c->world * b->s_max * bpr_max * D * sizeof(float));
    b->off_flags = al(b->off_xn + 2 * kBatchPlane * D * sizeof(float));
    b->off_done = al(b->off_flags + (size_t)3 * c->world * sizeof(unsigned));
    b->off_step = b->off_done + 256;
    int rc = peer_block_alloc(c, b->off_step + 256, &b->blk);
    if (rc == RAMA_OK && connect) rc = peer_block_connect(c, &b->blk, b->stream);
    if (rc != RAMA_OK) { rama_batch_destroy(b); return rc; }
    b->xn = reinterpret_cast<float*>(b->blk.local + b->off_xn);
  }
  b->graphs.assign(max_seqs + 1, nullptr);
  b->graphs_chained.assign(max_seqs + 1, nullptr);
  b->fence = std::make_shared<BatchFence>();
  if (cudaEventCreateWithFlags(&b->fence->ev, cudaEventDisableTiming) != cudaSuccess) {
    rama_batch_destroy(b);
    return fail(RAMA_E_CUDA, "batch fence event");
  }
  c->n_objects.fetch_add(1);
  *out = b;
  return RAMA_OK;
}

extern "C" int rama_batch_create(rama_ctx* c, int32_t max_seqs, rama_batch** out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  if (!c->loaded) return fail(RAMA_E_STATE, "no weights loaded");
  if (max_seqs < 1 || max_seqs > kBatchMax) return fail(RAMA_E_INVALID, "max_seqs must be in [1, %d]", kBatchMax);
  if (c->ranks.empty()) return batch_create_rank(c, max_seqs, out, true);
  rama_batch* g = new rama_batch();  // single-process group: one batch per rank, exchange blocks cross-wired
  g->ctx = c;
  g->cap = max_seqs;
  for (rama_ctx* rc : c->ranks) {
    rama_batch* rb = nullptr;
    int r = batch_create_rank(rc, max_seqs, &rb, false);
    if (r != RAMA_OK) {
      for (rama_batch* x : g->ranks) rama_batch_destroy(x);
      delete g;
      return r;
    }
    g->ranks.push_back(rb);
  }
  std::vector<PeerBlock*> blocks;
  for (rama_batch* r : g->ranks) blocks.push_back(&r->blk);
  peer_blocks_connect_group(blocks.data(), (int)blocks.size());
  c->n_objects.fetch_add(1);
  *out = g;
  return RAMA_OK;
}

extern "C" int rama_batch_destroy(rama_batch* b) {
  if (!b) return RAMA_OK;
  if (!b->ranks.empty()) {
    for (rama_batch* r : b->ranks) { cudaSetDevice(r->ctx->device); cudaStreamSynchronize(r->stream); }
    for (rama_batch* r : b->ranks) rama_batch_destroy(r);
    b->ctx->n_objects.fetch_sub(1);
    delete b;
    return RAMA_OK;
  }
  cudaSetDevice(b->ctx->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  std::lock_guard<std::mutex> cap_lk(b->ctx->cap_mu);
  for (auto g : b->graphs) if (g) cudaGraphExecDestroy(g);
  for (auto g : b->graphs_chained) if (g) cudaGraphExecDestroy(g);
  if (b->p2p) {
    peer_block_free(b->ctx, &b->blk, b->stream, b->x);
    b->xn = nullptr;  // lived inside the block
  }
  void* bufs[] = {b->x, b->xn, b->q, b->att, b->h, b->part, b->attn_ws, b->tickets, b->d_seqs, b->d_sp, b->d_next, b->red, b->lstage, b->d_err};
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
static int enqueue_batch_step(rama_batch* b, int n, int* n_launch, bool chained = false) {
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
  const bool xchg = b->p2p;
  unsigned* step_counter = xchg ? reinterpret_cast<unsigned*>(b->blk.local + b->off_step) : nullptr;
  LK(launch_k(pdl, batch_embed_kernel, dim3(n), dim3(256), st, b->d_seqs, W[RAMA_T_TOKEN_EMBEDDING], b->x, D, c->V, step_counter, chained ? 1 : 0));
  const bool ps = batch_ps(c) != 0;
  // the batch's activation operands: n rows, or (pre-split) the [128][K] matrix of activations + remainder plane
  const size_t brows = ps ? 2 * kBatchPlane : (size_t)n;
  GemmOperand X{b->xn, brows, (size_t)D};
  float* const xn_lo = ps ? b->xn + (size_t)kBatchPlane * D : nullptr;
  float* const att_lo = ps ? b->att + (size_t)kBatchPlane * Dq : nullptr;
  float* const h_lo = ps ? b->h + (size_t)kBatchPlane * Fl : nullptr;
  auto gemm = [&](const GemmOperand* A, int n_a, const GemmOperand* Bop, int rows, int K, int S, const auto& epi) -> cudaError_t {
    return ps ? BATCH_GEMM_PS(st, A, n_a, Bop, 1, rows, n, K, 0, S, epi, pdl) : BATCH_GEMM_R1(st, A, n_a, Bop, 1, rows, n, K, 0, S, epi, pdl);
  };
  // Tensor parallelism over peer memory (tp_exchange.cuh): sequences dealt in blocks of bpr; the wo / w2 GEMMs push their
  // split-K partials to the owners, tp_addnorm sums ranks × splits in fixed order, normalises and stores into every rank's xn.
  const int P = c->world;
  const int bpr = xchg ? (n + P - 1) / P : n;
  const int row0 = std::min(n, c->rank * bpr), n_rows = std::max(0, std::min(n - row0, bpr));
  const unsigned n_epochs = 2u * (unsigned)L + 2u;  // exchanges of one step (wo, w2 per layer; + the logits barrier)
  unsigned xi = 0;                                   // index of the next exchange inside the step
  TpPeers tpp{};
  if (xchg) {
    tpp.P = P; tpp.me = c->rank;
    for (int r = 0; r < P; ++r) tpp.flags[r] = reinterpret_cast<unsigned*>(b->blk.base[r] + b->off_flags);
  }
  // this rank's split-K partials [S][n][D] → summed, every reduced row stored once into its owner's inbox (slab = this rank)
  auto sum_push = [&](int S) -> int {
    TpInboxes ib{};
    for (int r = 0; r < P; ++r) ib.p[r] = reinterpret_cast<float*>(b->blk.base[r]);
    LK(launch_k(pdl, tp_sum_push_kernel, dim3((D / 4 + 255) / 256, n), dim3(256), st, (const float*)b->part, S, (size_t)n * D, ib, bpr, c->rank, D));
    return RAMA_OK;
  };
  auto exchange = [&](int S, const float* norm_w) -> int {
    TpAddNormParams ap{};
    ap.x = b->x; ap.inbox = reinterpret_cast<const float*>(b->blk.local);
    (void)S;
    ap.n_slab = P; ap.slab_stride = (size_t)bpr * D; ap.w = norm_w;
    for (int r = 0; r < P; ++r) { ap.xn[r] = reinterpret_cast<float*>(b->blk.base[r] + b->off_xn); ap.xlast[r] = nullptr; }
    ap.last_row = -1; ap.row0 = row0; ap.n_rows = n_rows; ap.rpr = bpr; ap.D = D;
    ap.lo_off = ps ? (size_t)kBatchPlane * D : 0;
    ap.tp = tpp;
    ap.epoch = TpEpoch{step_counter, n_epochs, ++xi};
    ap.done = reinterpret_cast<unsigned*>(b->blk.local + b->off_done);
    ap.error = &b->d_err[0];
    LK(launch_k(pdl, tp_addnorm_kernel, dim3(std::max(1, n_rows)), dim3(kTpNormThreads), st, ap));
    return RAMA_OK;
  };
  int S_prev = 0;  // split factor of the pending residual partials in b->part (0: none)
  const float* pending = b->part;
  // tensor parallelism: row-parallel wo / w2 leave a partial [n][D] on every rank — sum the split-K partials, all-reduce
  // over NVLink (NCCL, 1 MB at 64 sequences), and hand the reduced buffer to the next addnorm as a single "split"
  auto reduce_ranks = [&](int& S) -> int {
    if (c->world <= 1 || xchg) return RAMA_OK;
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
    if (xchg && l > 0) RK(exchange(S_prev, W[RAMA_T_RMS_ATT] + (size_t)l * D));
    else
    LK(launch_k(pdl, batch_addnorm_kernel, dim3(n), dim3(kBatchNormThreads), st, b->x, S_prev ? pending : nullptr, S_prev,
                (size_t)n * D, W[RAMA_T_RMS_ATT] + (size_t)l * D, b->xn, D, xn_lo));
    {  // [wq;wk;wv] (weights = the 128-row operand, the batch = the 64-column operand)   (infer.rs:20-23)
      GemmOperand A[3] = {{W[RAMA_T_WQ] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WK] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WV] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D}};
      const int S = pick_ksplit(c, 3 * tiles(Dq), D);
      EpiStoreT epi{b->part, Dq, n, S, (size_t)n * Dq};
      LK(gemm(A, 3, &X, Dq, D, S, epi));
      LK(launch_k(pdl, batch_qkv_finish_kernel, dim3(n, (Dq / 2 + 255) / 256), dim3(256), st, b->part, S, (size_t)n * Dq, b->d_seqs, layer_off, b->q, W[RAMA_T_FREQ_REAL], W[RAMA_T_FREQ_IMAG], Dq, hs / 2));
    }
    {  // attention per sequence   (infer.rs:34)
      AttnBatchParams ap{b->d_seqs, b->q, b->att, b->attn_ws, b->tickets, layer_off, T, Dq, hs, b->n_split, c->Hl, att_lo};
      LK(launch_k(pdl, attn_decode_batch_kernel, dim3(c->Hl, std::min(b->n_split, 2), n), dim3(kAttnThreads), st, ap));  // CTAs stride over the chunks
    }
    int S_wo;
    {  // wo   (infer.rs:35)
      GemmOperand A{W[RAMA_T_WO] + (size_t)l * D * Dq, (size_t)D, (size_t)Dq};
      GemmOperand Bm{b->att, brows, (size_t)Dq};
      S_wo = pick_ksplit(c, tiles(D), Dq);
      EpiStoreT epi{b->part, D, n, S_wo, (size_t)n * D};
      LK(gemm(&A, 1, &Bm, D, Dq, S_wo, epi));
      if (xchg) RK(sum_push(S_wo));
      RK(reduce_ranks(S_wo));
    }
    // x += wo output; xn = rmsnorm(x)   (infer.rs:37-38)
    if (xchg) RK(exchange(S_wo, W[RAMA_T_RMS_FFN] + (size_t)l * D));
    else
    LK(launch_k(pdl, batch_addnorm_kernel, dim3(n), dim3(kBatchNormThreads), st, b->x, pending, S_wo, (size_t)n * D, W[RAMA_T_RMS_FFN] + (size_t)l * D, b->xn, D, xn_lo));
    {  // [w1;w3] → SwiGLU   (infer.rs:39-45)
      GemmOperand A[2] = {{W[RAMA_T_W1] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D},
                          {W[RAMA_T_W3] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D}};
      const int S = pick_ksplit(c, 2 * tiles(Fl), D);
      EpiStoreT epi{b->part, Fl, n, S, (size_t)n * Fl};
      LK(gemm(A, 2, &X, Fl, D, S, epi));
      LK(launch_k(pdl, batch_swiglu_finish_kernel, dim3(std::min(c->sm_count * 4, (n * Fl + 255) / 256)), dim3(256), st, b->part, S, (size_t)n * Fl, b->h, Fl, n, h_lo));
    }
    {  // w2   (infer.rs:46)
      GemmOperand A{W[RAMA_T_W2] + (size_t)l * D * Fl, (size_t)D, (size_t)Fl};
      GemmOperand Bm{b->h, brows, (size_t)Fl};
      S_prev = pick_ksplit(c, tiles(D), Fl);
      EpiStoreT epi{b->part, D, n, S_prev, (size_t)n * D};
      LK(gemm(&A, 1, &Bm, D, Fl, S_prev, epi));
      if (xchg) RK(sum_push(S_prev));
      RK(reduce_ranks(S_prev));
    }
  }
  // x += w2 output; final rmsnorm; classifier → each session's logits   (infer.rs:49-51)
  if (xchg) RK(exchange(S_prev, W[RAMA_T_RMS_FINAL]));
  else
  LK(launch_k(pdl, batch_addnorm_kernel, dim3(n), dim3(kBatchNormThreads), st, b->x, pending, S_prev, (size_t)n * D, W[RAMA_T_RMS_FINAL], b->xn, D, xn_lo));
  {
    GemmOperand A{c->wcls, (size_t)Vl, (size_t)D};
    const int S = pick_ksplit(c, tiles(Vl), D);
    EpiStoreT epi{b->part, Vl, n, S, (size_t)n * Vl};
    LK(gemm(&A, 1, &X, Vl, D, S, epi));
    if (xchg) {  // vocabulary rows are split: every rank stores its slice into the sessions' logits on every rank, then a barrier
      LK(launch_k(pdl, batch_cls_push_kernel, dim3(std::min(64, (Vl + 255) / 256), n), dim3(256), st, b->part, S, (size_t)n * Vl, b->d_seqs, Vl, c->v0, P));
      LK(launch_k(pdl, tp_barrier_kernel, dim3(1), dim3(32), st, tpp, TpEpoch{step_counter, n_epochs, n_epochs}, &b->d_err[0]));
    } else if (c->world > 1) {  // NCCL mode: gather every rank's block, then scatter into the sessions' logits
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

// group batch: the rank-r batch steps the rank-r sessions
static int group_sessions(rama_batch* b, rama_session* const* sessions, int32_t n, std::vector<std::vector<rama_session*>>& per_rank) {
  if (n < 1 || n > b->cap) return fail(RAMA_E_INVALID, "batch of %d sequences outside [1, %d]", n, b->cap);
  per_rank.assign(b->ranks.size(), std::vector<rama_session*>((size_t)n));
  for (int i = 0; i < n; ++i) {
    if (!sessions[i] || sessions[i]->ctx != b->ctx || sessions[i]->ranks.size() != b->ranks.size())
      return fail(RAMA_E_INVALID, "session %d is NULL or belongs to another context", i);
    for (size_t r = 0; r < b->ranks.size(); ++r) per_rank[r][i] = sessions[i]->ranks[r];
  }
  return RAMA_OK;
}

extern "C" int rama_forward_batch(rama_batch* b, rama_session* const* sessions, const int32_t* tokens,
                                  const int32_t* pos, int32_t n) {
  if (!b || !sessions || !tokens || !pos) return fail(RAMA_E_INVALID, "NULL argument");
  if (!b->ranks.empty()) {
    std::vector<std::vector<rama_session*>> pr;
    RK(group_sessions(b, sessions, n, pr));
    return group_run(b->ctx, [&](int r) { return rama_forward_batch(b->ranks[r], pr[r].data(), tokens, pos, n); });
  }
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
    hs[i] = BatchSeq{s->key_cache, s->value_cache, s->logits, s->ctrl, pos[i], tokens[i], {nullptr}};
    if (b->p2p) {
      if (!s->p2p) return fail(RAMA_E_STATE, "session %d was not created for the peer exchange", i);
      for (int r = 0; r < c->world; ++r) hs[i].logits_peer[r] = reinterpret_cast<float*>(s->blk.base[r] + s->off_logits);
    }
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
  if (!b->ranks.empty()) {  // every rank holds every session's full logits: rank 0 samples
    std::vector<std::vector<rama_session*>> pr;
    RK(group_sessions(b, sessions, n, pr));
    RK(group_run(b->ctx, [&](int r) { return r == 0 ? RAMA_OK : rama_batch_sync(b->ranks[r]); }));
    return rama_sample_batch(b->ranks[0], pr[0].data(), n, temperature, topp, next);
  }
  rama_ctx* c = b->ctx;
  RK(batch_check_sessions(b, sessions, n));
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(b->stream));  // h_sp / h_next are single-buffered
  BatchSeq* hs = b->h_seqs + (size_t)b->ring_i * b->cap;
  if (++b->ring_i == kBatchRing) b->ring_i = 0;
  for (int i = 0; i < n; ++i) {
    rama_session* s = sessions[i];
    b->h_sp[i] = SampleParams{s->logits, nullptr, 0, 0, c->V, s->ctrl, nullptr, nullptr, s->sort_keys, temperature, topp, 0, PeerIn{}};
    hs[i] = BatchSeq{s->key_cache, s->value_cache, s->logits, s->ctrl, 0, 0, {nullptr}};
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
  if (b->p2p) {
    int32_t xerr = 0;
    CK(cudaMemcpy(&xerr, b->d_err, sizeof(xerr), cudaMemcpyDeviceToHost));
    if (xerr == 3) return fail(RAMA_E_NCCL, "timed out waiting for a tensor-parallel peer in the batched step");
  }
  for (int i = 0; i < n; ++i) {
    if (b->h_next[2 * i + 1] == 1) return fail(RAMA_E_STATE, "token id outside the vocabulary reached the device step (sequence %d)", i);
    if (b->h_next[2 * i + 1] == 2) return fail(RAMA_E_STATE, "top-p candidate list is empty (the reference panics here, infer.rs:66) (sequence %d)", i);
    next[i] = b->h_next[2 * i];
  }
  return RAMA_OK;
}

// ≙ n generate() loops (mod.rs:169-206) advanced together with the token feedback on the device: every sequence starts at
// BOS / position 0, its prompt is forced, `steps` tokens each (out_tokens[i·steps + t]).  One captured graph per batch size
// holds the batched step and the n samplers; nothing crosses PCIe between steps.  Collective under tensor parallelism.
extern "C" int rama_generate_batch(rama_batch* b, rama_session* const* sessions, int32_t n, const int32_t* const* prompts,
                                   const int32_t* n_prompt, int32_t steps, float temperature, float topp, int32_t* out_tokens,
                                   float* elapsed_ms) {
  if (!b || !sessions || !n_prompt || !out_tokens) return fail(RAMA_E_INVALID, "NULL argument");
  if (!b->ranks.empty()) {
    std::vector<std::vector<rama_session*>> pr;
    RK(group_sessions(b, sessions, n, pr));
    std::vector<std::vector<int32_t>> scratch(b->ranks.size());
    return group_run(b->ctx, [&](int r) {
      if (r == 0) return rama_generate_batch(b->ranks[0], pr[0].data(), n, prompts, n_prompt, steps, temperature, topp, out_tokens, elapsed_ms);
      scratch[r].resize((size_t)std::max(1, n * steps));
      return rama_generate_batch(b->ranks[r], pr[r].data(), n, prompts, n_prompt, steps, temperature, topp, scratch[r].data(), nullptr);
    });
  }
  rama_ctx* c = b->ctx;
  RK(batch_check_sessions(b, sessions, n));
  if (steps < 0 || steps > c->T) return fail(RAMA_E_STATE, "steps %d exceeds seq_len %d", steps, c->T);  // the reference panics past seq_len
  for (int i = 0; i < n; ++i) {
    if (n_prompt[i] < 0 || (n_prompt[i] > 0 && (!prompts || !prompts[i]))) return fail(RAMA_E_INVALID, "prompt %d is NULL", i);
    for (int j = 0; j < n_prompt[i]; ++j)
      if (prompts[i][j] < 0 || prompts[i][j] >= c->V) return fail(RAMA_E_INVALID, "prompt token %d of sequence %d outside the vocabulary", prompts[i][j], i);
  }
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(b->stream));  // h_sp and the pinned rings are reused below
  BatchSeq* hs = b->h_seqs + (size_t)b->ring_i * b->cap;
  if (++b->ring_i == kBatchRing) b->ring_i = 0;
  for (int i = 0; i < n; ++i) {
    rama_session* s = sessions[i];
    CK(session_enter(s));
    CK(cudaStreamSynchronize(s->stream));
    s->async_pending = false;
    hs[i] = BatchSeq{s->key_cache, s->value_cache, s->logits, s->ctrl, 0, 1, {nullptr}};
    if (b->p2p) {
      if (!s->p2p) return fail(RAMA_E_STATE, "session %d was not created for the peer exchange", i);
      for (int r = 0; r < c->world; ++r) hs[i].logits_peer[r] = reinterpret_cast<float*>(s->blk.base[r] + s->off_logits);
    }
    b->h_sp[i] = SampleParams{s->logits, nullptr, 0, 0, c->V, s->ctrl, s->d_prompt, s->d_out, s->sort_keys, 0.f, 0.f, 1, PeerIn{}};
    const int np = std::min<int>(n_prompt[i], c->T);
    if (np) CK(cudaMemcpyAsync(s->d_prompt, prompts[i], (size_t)np * sizeof(int32_t), cudaMemcpyHostToDevice, b->stream));
    StepCtrl* h = &s->h_ring[s->ring_i];
    if (++s->ring_i == kRing) s->ring_i = 0;
    memset(h, 0, sizeof(*h));
    h->pos = 0; h->token = 1;  // BOS (mod.rs:182)
    h->chained = 1; h->n_prompt = n_prompt[i]; h->temperature = temperature; h->topp = topp;
    CK(cudaMemcpyAsync(s->ctrl, h, sizeof(StepCtrl), cudaMemcpyHostToDevice, b->stream));
    s->logits_gathered = true;
    s->parts_valid = false;
    s->fence = b->fence;
  }
  CK(cudaMemcpyAsync(b->d_seqs, hs, (size_t)n * sizeof(BatchSeq), cudaMemcpyHostToDevice, b->stream));
  CK(cudaMemcpyAsync(b->d_sp, b->h_sp, (size_t)n * sizeof(SampleParams), cudaMemcpyHostToDevice, b->stream));
  if (!b->graphs_chained[n]) {
    std::lock_guard<std::mutex> cap_lk(c->cap_mu);
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeRelaxed));
    int nl = 0;
    int rc = enqueue_batch_step(b, n, &nl, true);
    if (rc == RAMA_OK) {
      sample_batch_kernel<<<n, kSampleThreads, 0, b->stream>>>(b->d_sp);
      if (cudaGetLastError() != cudaSuccess) rc = fail(RAMA_E_CUDA, "sampler launch in the batched loop");
    }
    cudaError_t e = cudaStreamEndCapture(b->stream, &g);
    if (rc != RAMA_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&b->graphs_chained[n], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    b->launches = nl;
  }
  CK(cudaEventRecord(b->ev0, b->stream));
  for (int t = 0; t < steps; ++t) CK(cudaGraphLaunch(b->graphs_chained[n], b->stream));
  CK(cudaEventRecord(b->ev1, b->stream));
  batch_collect_kernel<<<1, 64, 0, b->stream>>>(b->d_seqs, b->d_next, n);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(b->h_next, b->d_next, (size_t)2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
  for (int i = 0; i < n && steps > 0; ++i)
    CK(cudaMemcpyAsync(out_tokens + (size_t)i * steps, sessions[i]->d_out, (size_t)steps * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaEventRecord(b->fence->ev, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, b->ev0, b->ev1));
  if (b->p2p) {
    int32_t xerr = 0;
    CK(cudaMemcpy(&xerr, b->d_err, sizeof(xerr), cudaMemcpyDeviceToHost));
    if (xerr == 3) return fail(RAMA_E_NCCL, "timed out waiting for a tensor-parallel peer in the batched loop");
  }
  for (int i = 0; i < n; ++i) {
    if (b->h_next[2 * i + 1] == 1) return fail(RAMA_E_STATE, "token id outside the vocabulary reached the device step (sequence %d)", i);
    if (b->h_next[2 * i + 1] == 2) return fail(RAMA_E_STATE, "top-p candidate list is empty (the reference panics here, infer.rs:66) (sequence %d)", i);
  }
  return RAMA_OK;
}

extern "C" int rama_batch_sync(rama_batch* b) {
  if (!b) return fail(RAMA_E_INVALID, "NULL batch");
  if (!b->ranks.empty()) return group_run(b->ctx, [&](int r) { return rama_batch_sync(b->ranks[r]); });
  CK(cudaSetDevice(b->ctx->device));
  CK(cudaStreamSynchronize(b->stream));
  return RAMA_OK;
}

extern "C" int rama_batch_launches_per_step(const rama_batch* b, int32_t* n) {
  if (!b || !n) return fail(RAMA_E_INVALID, "NULL argument");
  if (!b->ranks.empty()) return rama_batch_launches_per_step(b->ranks[0], n);
  *n = b->launches;
  return RAMA_OK;
}

