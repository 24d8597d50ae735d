// prefill_api.cu — prompt prefill on the tensor cores (host side): ≙ the prompt part of generate()'s loop, mod.rs:187-192.
#include "internal.cuh"

#include "gemm_host.cuh"
#include "prefill.cuh"

#include <array>

constexpr int kPrefillChunk = 512;

// The prefill GEMM tile: 128×128 per CTA, or (RAMA_PREFILL_PAIR=1, chunks of more than 128 rows) a CTA pair on the two SMs of a
// TPC computing 256×128 with tcgen05.mma.cta_group::2 — each CTA holds half of the weight k-block (gemm_tf32x3.cuh PAIR).
// The pair is parity-green and measured SLOWER (420 vs 552 TFLOP/s tf32 issued at 512×4096×4096): with K = 8 per tf32
// instruction each SM must pull 2 KB of the peer's B half per MMA through the SM-to-SM path (≈21 B/clk), 97 clk against the
// 64 clk of the math — the operand sharing that pays for 16-bit kinds does not for 3xTF32.  Off by default.
static int prefill_pair() {
  static const int v = env_int("RAMA_PREFILL_PAIR", 0);
  return v;
}
#define PREFILL_GEMM(st, A, nA, B, nB, M, N, K, epi, pdl)                                                       \
  ((prefill_pair() && (M) > kGemmBM) ? launch_gemm_tf32x3<128, 4, 4, 0, 0, 0, 2, 1>(st, A, nA, B, nB, M, N, K, 0, 1, epi, pdl) \
                                     : launch_gemm_tf32x3<128, 4, 4, 0>(st, A, nA, B, nB, M, N, K, 0, 1, epi, pdl))

// Allocations of the prefill workspace.  Under the peer exchange the normalised activations (every rank stores its rows
// into every rank's copy) and the inbox of partial rows live in a peer-addressable block: collective between processes;
// a single-process group allocates on all ranks first and cross-wires the blocks (prefill_prepare_group).
int prefill_alloc_ws(rama_session* s) {
  if (s->pf_cap) return RAMA_OK;
  rama_ctx* c = s->ctx;
  CK(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> cap_lk(c->cap_mu);  // allocations vs another thread's stream capture (rama_ctx::cap_mu)
  const size_t cap = std::min(c->T, kPrefillChunk);
  cudaError_t e = cudaSuccess;
#define A(call) if (e == cudaSuccess) e = (call)
  A(cudaMalloc((void**)&s->pf_x, cap * c->D * sizeof(float)));
  if (!s->p2p) A(cudaMalloc((void**)&s->pf_xn, cap * c->D * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_q, cap * c->Dq * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_att, cap * c->Dq * sizeof(float)));
  if (!s->p2p) A(cudaMalloc((void**)&s->pf_y, cap * c->D * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_h, cap * c->Fl * sizeof(float)));
  A(cudaMalloc((void**)&s->pf_tokens, cap * sizeof(int32_t)));
  A(cudaHostAlloc((void**)&s->h_tokens, (size_t)c->T * sizeof(int32_t), cudaHostAllocDefault));
#undef A
  if (e != cudaSuccess) return fail(RAMA_E_CUDA, "prefill workspace: %s", cudaGetErrorString(e));
  if (s->p2p) {
    // inbox[P][rpr_max][D] | xn[cap][D]
    s->pf_rpr_max = (int)((cap + c->world - 1) / c->world);
    s->pf_off_xn = ((size_t)c->world * s->pf_rpr_max * c->D * sizeof(float) + 255) / 256 * 256;
    RK(peer_block_alloc(c, s->pf_off_xn + cap * c->D * sizeof(float), &s->pf_blk));
    s->pf_xn = reinterpret_cast<float*>(s->pf_blk.local + s->pf_off_xn);
  }
  static std::atomic<unsigned long long> done4{0}, done1{0};
  cudaError_t attr_err = ensure_dyn_smem((const void*)prefill_attn_kernel<4>, (int)prefill_attn_smem_bytes(kPfMaxHs, 4), done4);
  if (attr_err == cudaSuccess)
    attr_err = ensure_dyn_smem((const void*)prefill_attn_kernel<1>, (int)prefill_attn_smem_bytes(kPfMaxHs, 1), done1);
  if (attr_err != cudaSuccess) return fail(RAMA_E_CUDA, "prefill attention smem: %s", cudaGetErrorString(attr_err));
  s->pf_cap = (int)cap;
  return RAMA_OK;
}

static int ensure_prefill_ws(rama_session* s) {
  if (s->pf_cap) return RAMA_OK;
  RK(prefill_alloc_ws(s));
  if (s->p2p) RK(peer_block_connect(s->ctx, &s->pf_blk, s->stream));  // (no-op for the ranks of a group)
  return RAMA_OK;
}

// Prefill attention on the tensor cores (prefill.cuh, prefill_attn_mma_kernel): one instantiation per head size × (query warps NW,
// key shares KS).  RAMA_PREFILL_ATTN=cuda keeps the f32 CUDA-core kernel (the only path for head sizes outside the table below);
// mma41 / mma42 / mma44 / mma22 / mma14 force one shape (tests, A/B runs).
static int prefill_attn_mode() {   // (read per call: a getenv next to a multi-millisecond prefill, and the tests switch it)
  const char* e = getenv("RAMA_PREFILL_ATTN");
  if (!e || !*e || !strcmp(e, "mma")) return 1;
  if (!strcmp(e, "cuda")) return 0;
  if (!strcmp(e, "mma41")) return 41;
  if (!strcmp(e, "mma42")) return 42;
  if (!strcmp(e, "mma44")) return 44;
  if (!strcmp(e, "mma22")) return 22;
  if (!strcmp(e, "mma14")) return 14;
  return 1;
}

template <int HS, int NW, int KS>
static cudaError_t launch_attn_mma(cudaLaunchConfig_t cfg, int M, int heads, const PrefillAttnParams& ap) {
  static std::atomic<unsigned long long> done{0};
  cudaError_t e = ensure_dyn_smem((const void*)prefill_attn_mma_kernel<HS, NW, KS>, (int)prefill_attn_mma_smem_bytes(HS, NW), done);
  if (e != cudaSuccess) return e;
  const int nq = (M + 16 * NW - 1) / (16 * NW);
  cfg.gridDim = dim3((nq + 1) / 2, heads);
  cfg.blockDim = dim3(32 * NW * KS);
  cfg.dynamicSmemBytes = prefill_attn_mma_smem_bytes(HS, NW);
  return cudaLaunchKernelEx(&cfg, prefill_attn_mma_kernel<HS, NW, KS>, ap);
}

template <int NW, int KS>
static cudaError_t launch_attn_mma_hs(const cudaLaunchConfig_t& cfg, int M, int heads, const PrefillAttnParams& ap) {
  switch (ap.hs) {
    case 16: return launch_attn_mma<16, NW, KS>(cfg, M, heads, ap);
    case 32: return launch_attn_mma<32, NW, KS>(cfg, M, heads, ap);
    case 48: return launch_attn_mma<48, NW, KS>(cfg, M, heads, ap);
    case 64: return launch_attn_mma<64, NW, KS>(cfg, M, heads, ap);
    case 96: return launch_attn_mma<96, NW, KS>(cfg, M, heads, ap);
    case 128: return launch_attn_mma<128, NW, KS>(cfg, M, heads, ap);
    default: return cudaErrorInvalidValue;
  }
}
static bool attn_mma_has(int hs) { return hs == 16 || hs == 32 || hs == 48 || hs == 64 || hs == 96 || hs == 128; }

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

// one chunk of M ≤ pf_cap rows at positions [pos0, pos0+M); `last`: also produce the logits of the final row.
// `tokens` must be pinned host memory (s->h_tokens).
static int prefill_chunk(rama_session* s, const int32_t* tokens, int M, int pos0, bool last, PfTrace& tr, int* n_launch) {
  rama_ctx* c = s->ctx;
  cudaStream_t st = s->stream;
  const int D = c->D, Dq = c->Dq, Fl = c->Fl, T = c->T, hs = c->hs, L = c->L, P = c->world;
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
  // programmatic dependent launch along the whole chain (not while per-launch events are being recorded)
  const bool pdl = c->use_pdl && !tr.on && env_int("RAMA_PREFILL_PDL", 1);
  GK(RAMA_PK_OTHER, launch_k(pdl, prefill_embed_kernel, dim3(M), dim3(256), st, (const int32_t*)s->pf_tokens, W[RAMA_T_TOKEN_EMBEDDING],
                             s->pf_x, D, c->V, &s->ctrl->error, s->seq));

  // Tensor parallelism over peer memory (tp_exchange.cuh): rows dealt in blocks of rpr; the wo / w2 GEMMs push their partial
  // rows to the owners, tp_addnorm sums them, normalises and stores the rows into every rank's xn.
  const bool xchg = s->p2p;
  const int rpr = xchg ? (M + P - 1) / P : M;
  const int row0 = std::min(M, c->rank * rpr), n_rows = std::max(0, std::min(M - row0, rpr));
  EpiPushNT push{};
  if (xchg) {
    for (int r = 0; r < P; ++r) push.inbox[r] = reinterpret_cast<float*>(s->pf_blk.base[r]);
    push.rpr = rpr; push.me = c->rank; push.ldc = D; push.N = D;
  }
  // norm_w == null: residual add only; last_row ≥ 0: that row's updated residual goes to every rank's x0
  auto exchange = [&](const float* norm_w, int last_row) -> int {
    TpAddNormParams ap{};
    ap.x = s->pf_x; ap.inbox = reinterpret_cast<const float*>(s->pf_blk.local);
    ap.n_slab = P; ap.slab_stride = (size_t)rpr * D; ap.w = norm_w;
    for (int r = 0; r < P; ++r) {
      ap.xn[r] = reinterpret_cast<float*>(s->pf_blk.base[r] + s->pf_off_xn);
      ap.xlast[r] = last_row >= 0 ? reinterpret_cast<float*>(s->blk.base[r] + s->off_x0) : nullptr;
    }
    ap.last_row = last_row; ap.row0 = row0; ap.n_rows = n_rows; ap.rpr = rpr; ap.D = D;
    ap.tp = tp_peers(s);
    ap.epoch = TpEpoch{nullptr, 0u, ++s->bulk_epoch};
    ap.done = reinterpret_cast<unsigned*>(s->blk.local + s->off_done);
    ap.error = &s->ctrl->error;
    GK(RAMA_PK_COMM, launch_k(pdl, tp_addnorm_kernel, dim3(std::max(1, n_rows)), dim3(kTpNormThreads), st, ap));
    return RAMA_OK;
  };

  for (int l = 0; l < L; ++l) {
    float* kc = s->key_cache + (size_t)l * T * Dq;
    float* vc = s->value_cache + (size_t)l * T * Dq;
    // x += pending w2 output; xn = rmsnorm(x)·w_att   (infer.rs:19)
    if (xchg && l > 0) {
      RK(exchange(W[RAMA_T_RMS_ATT] + (size_t)l * D, -1));
    } else {  // layer 0: x is the embedding on every rank, nothing pending
      GK(RAMA_PK_NORM, launch_k(pdl, prefill_addnorm_kernel, dim3(M), dim3(256), st, s->pf_x, (const float*)(l == 0 ? nullptr : s->pf_y),
                                W[RAMA_T_RMS_ATT] + (size_t)l * D, s->pf_xn, D));
    }
    // [wq;wk;wv] → RoPE → Q, KV-cache rows   (infer.rs:20-33)
    {
      GemmOperand A{s->pf_xn, (size_t)M, (size_t)D};
      GemmOperand B[3] = {{W[RAMA_T_WQ] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WK] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D},
                          {W[RAMA_T_WV] + (size_t)l * Dq * D, (size_t)Dq, (size_t)D}};
      EpiQKVPrefill epi{s->pf_q, kc, vc, W[RAMA_T_FREQ_REAL], W[RAMA_T_FREQ_IMAG], pos0, Dq, hs / 2};
      GK(RAMA_PK_GEMM, PREFILL_GEMM(st, &A, 1, B, 3, M, Dq, D, epi, pdl));
    }
    // causal attention of every prompt row over the cache   (infer.rs:34)
    {
      PrefillAttnParams ap{s->pf_q, kc, vc, s->pf_att, M, pos0, Dq, hs};
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
      // tensor-core kernel: the largest query block whose paired grid still covers most of the SMs; eight warps per CTA either way
      int amode = attn_mma_has(hs) ? prefill_attn_mode() : 0;
      const int nq32 = (M + 31) / 32;
      if (amode == 1) amode = ((nq64 + 1) / 2) * c->Hl >= 96 ? 42 : ((nq32 + 1) / 2) * c->Hl >= 96 ? 22 : ((nq16 + 1) / 2) * c->Hl >= 48 ? 14 : 0;
      if (amode == 44) {
        GK(RAMA_PK_ATTN, (launch_attn_mma_hs<4, 4>(cfg, M, c->Hl, ap)));
      } else if (amode == 42) {
        GK(RAMA_PK_ATTN, (launch_attn_mma_hs<4, 2>(cfg, M, c->Hl, ap)));
      } else if (amode == 41) {
        GK(RAMA_PK_ATTN, (launch_attn_mma_hs<4, 1>(cfg, M, c->Hl, ap)));
      } else if (amode == 22) {
        GK(RAMA_PK_ATTN, (launch_attn_mma_hs<2, 2>(cfg, M, c->Hl, ap)));
      } else if (amode == 14) {
        GK(RAMA_PK_ATTN, (launch_attn_mma_hs<1, 4>(cfg, M, c->Hl, ap)));
      } else if (((nq64 + 1) / 2) * c->Hl >= 96) {
        cfg.gridDim = dim3((nq64 + 1) / 2, c->Hl);
        cfg.dynamicSmemBytes = prefill_attn_smem_bytes(hs, 4);
        GK(RAMA_PK_ATTN, cudaLaunchKernelEx(&cfg, prefill_attn_kernel<4>, ap));
      } else {
        cfg.gridDim = dim3((nq16 + 1) / 2, c->Hl);
        cfg.dynamicSmemBytes = prefill_attn_smem_bytes(hs, 1);
        GK(RAMA_PK_ATTN, cudaLaunchKernelEx(&cfg, prefill_attn_kernel<1>, ap));
      }
    }
    // wo   (infer.rs:35); the residual add is the next addnorm
    {
      GemmOperand A{s->pf_att, (size_t)M, (size_t)Dq};
      GemmOperand B{W[RAMA_T_WO] + (size_t)l * D * Dq, (size_t)D, (size_t)Dq};
      if (xchg) {
        GK(RAMA_PK_GEMM, PREFILL_GEMM(st, &A, 1, &B, 1, M, D, Dq, push, pdl));
      } else {
        EpiStoreNT epi{s->pf_y, D, D, 0};
        GK(RAMA_PK_GEMM, PREFILL_GEMM(st, &A, 1, &B, 1, M, D, Dq, epi, pdl));
      }
    }
    if (xchg) {
      RK(exchange(W[RAMA_T_RMS_FFN] + (size_t)l * D, -1));
    } else {
      if (P > 1) {
        tr.pre(RAMA_PK_COMM);
        NK(g_nccl.AllReduce(s->pf_y, s->pf_y, (size_t)M * D, kNcclFloat32, kNcclSum, c->comm, st));
        tr.post(); ++launches;
      }
      GK(RAMA_PK_NORM, launch_k(pdl, prefill_addnorm_kernel, dim3(M), dim3(256), st, s->pf_x, (const float*)s->pf_y,
                                W[RAMA_T_RMS_FFN] + (size_t)l * D, s->pf_xn, D));
    }
    // [w1|w3] → SwiGLU   (infer.rs:39-45)
    {
      GemmOperand A{s->pf_xn, (size_t)M, (size_t)D};
      GemmOperand B[2] = {{W[RAMA_T_W1] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D},
                          {W[RAMA_T_W3] + (size_t)l * Fl * D, (size_t)Fl, (size_t)D}};
      EpiSwiGLUPrefill epi{s->pf_h, Fl};
      GK(RAMA_PK_GEMM, PREFILL_GEMM(st, &A, 1, B, 2, M, Fl, D, epi, pdl));
    }
    // w2   (infer.rs:46)
    {
      GemmOperand A{s->pf_h, (size_t)M, (size_t)Fl};
      GemmOperand B{W[RAMA_T_W2] + (size_t)l * D * Fl, (size_t)D, (size_t)Fl};
      if (xchg) {
        GK(RAMA_PK_GEMM, PREFILL_GEMM(st, &A, 1, &B, 1, M, D, Fl, push, pdl));
      } else {
        EpiStoreNT epi{s->pf_y, D, D, 0};
        GK(RAMA_PK_GEMM, PREFILL_GEMM(st, &A, 1, &B, 1, M, D, Fl, epi, pdl));
      }
    }
    if (P > 1 && !xchg) {
      tr.pre(RAMA_PK_COMM);
      NK(g_nccl.AllReduce(s->pf_y, s->pf_y, (size_t)M * D, kNcclFloat32, kNcclSum, c->comm, st));
      tr.post(); ++launches;
    }
  }
  CK(cudaGetLastError());
  // the last w2 output: under the exchange its owner adds it and hands the last row's residual to every rank's x0
  if (xchg) RK(exchange(nullptr, last ? M - 1 : -1));
  if (last) {
    // only the last row's logits exist after the reference's prompt loop: x0 = x + y of that row, then the
    // decode path's fused final-rmsnorm → classifier GEMV (infer.rs:49-51)
    if (!xchg)
      GK(RAMA_PK_OTHER, launch_k(pdl, prefill_last_row_kernel, dim3(std::max(1, D / 256)), dim3(256), st,
                                 (const float*)(s->pf_x + (size_t)(M - 1) * D), (const float*)(s->pf_y + (size_t)(M - 1) * D), s->x0, D));
    RK(init_parts(s));
    ProNorm pro{s->x0, nullptr, s->x1, W[RAMA_T_RMS_FINAL], s->xfinal, PeerIn{}};
    RowsPlain rows{c->wcls, D, c->Vl};
    EpiCls epi = make_epi_cls(s);
    const int np = (c->Vl + 1) / 2, var = pick_variant(c, D / 4);
    GK(RAMA_PK_OTHER, launch_gemv(var, pick_grid(c, var, np), st, 0, pro, rows, epi, D / 4, np));
    if (P > 1 && !s->p2p) {
      NK(g_nccl.AllGather(s->part + (size_t)c->rank * c->sm_count, s->part, (size_t)c->sm_count * 2, kNcclFloat32, c->comm, st));
      ++launches;
    }
  }
#undef GK
  if (n_launch) *n_launch += launches;
  return RAMA_OK;
}

int prefill_run(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float ms_kind[RAMA_PK_COUNT],
                int32_t* n_launch) {
  rama_ctx* c = s->ctx;
  if (n <= 0 || pos0 < 0 || (long long)pos0 + n > c->T)  // the reference panics past seq_len (infer.rs:32)
    return fail(RAMA_E_STATE, "prefill rows [%d, %d) outside [0, seq_len=%d)", pos0, pos0 + n, c->T);
  for (int i = 0; i < n; ++i)
    if (tokens[i] < 0 || tokens[i] >= c->V) return fail(RAMA_E_INVALID, "prompt token %d outside the vocabulary", tokens[i]);
  if (c->tp_sim) return fail(RAMA_E_STATE, "RAMA_TP_SIM measures the decode step only");
  CK(cudaSetDevice(c->device));
  RK(ensure_prefill_ws(s));
  PfTrace tr{s->stream, ms_kind != nullptr, {}, {}};
  int launches = 0, rc = RAMA_OK;
  CK(cudaStreamSynchronize(s->stream));  // h_tokens may still feed a copy of the previous call
  memcpy(s->h_tokens, tokens, (size_t)n * sizeof(int32_t));
  for (int c0 = 0; c0 < n && rc == RAMA_OK; c0 += s->pf_cap) {
    const int M = std::min(s->pf_cap, n - c0);
    rc = prefill_chunk(s, s->h_tokens + c0, M, pos0 + c0, c0 + M == n, tr, &launches);
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

// single-process group: the prefill workspaces of all ranks are allocated and their exchange blocks cross-wired before
// any rank starts (a rank-level call cannot see its peers' allocations)
int group_prefill_prepare(rama_session* g) {
  if (g->ranks.empty() || g->ranks[0]->pf_cap) return RAMA_OK;
  for (rama_session* r : g->ranks) RK(prefill_alloc_ws(r));
  std::vector<PeerBlock*> blocks;
  for (rama_session* r : g->ranks) blocks.push_back(&r->pf_blk);
  peer_blocks_connect_group(blocks.data(), (int)blocks.size());
  return RAMA_OK;
}

extern "C" int rama_prefill(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float* elapsed_ms,
                            float ms_kind[RAMA_PK_COUNT], int32_t* n_launch) {
  if (!s || !tokens) return fail(RAMA_E_INVALID, "NULL argument");
  if (!s->ranks.empty()) {
    RK(group_prefill_prepare(s));
    std::vector<std::array<float, RAMA_PK_COUNT>> k(s->ranks.size());
    return group_run(s->ctx, [&](int r) {
      return r == 0 ? rama_prefill(s->ranks[0], tokens, n, pos0, elapsed_ms, ms_kind, n_launch)
                    : rama_prefill(s->ranks[r], tokens, n, pos0, nullptr, ms_kind ? k[r].data() : nullptr, nullptr);
    });
  }
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
  for (rama_session* r : s->ranks) r->pf_min = min_rows;
  s->pf_min = min_rows;
  return RAMA_OK;
}

