// prefill_api.cu — prompt prefill on the tensor cores (host side): ≙ the prompt part of generate()'s loop, mod.rs:187-192.
#include "internal.cuh"

#include "gemm_host.cuh"
#include "prefill.cuh"

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

int prefill_run(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float ms_kind[RAMA_PK_COUNT],
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

