// ctx.cu — errors, NCCL (resolved at run time), the context (device + sharded weights) and the loaders.
// No cuBLAS, no NVRTC, no CPU fallback.
#include "internal.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
// used by the other translation units of the library (tokenizer.cpp)
extern "C" int rama_set_error(int code, const char* msg) { return fail(code, "%s", msg); }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time (only when tp->world > 1) so that the single-GPU path has no
// dependency on it and the process shares whatever libnccl.so.2 is already loaded.
// ------------------------------------------------------------------------------------------------
NcclApi g_nccl;
static std::mutex g_nccl_mu;

int nccl_load() {
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
  c->tp_nowait = env_int("RAMA_TP_NOWAIT", 0);
  c->variant_override = env_int("RAMA_GEMV_VARIANT", -1);
  c->staged = env_int("RAMA_GEMV_STAGED", 1);
  c->tp_cluster = env_int("RAMA_TP_CLUSTER", -1);
  c->tp_reduce = env_int("RAMA_TP_REDUCE", -1);
  c->embed_kernel = env_int("RAMA_EMBED_KERNEL", 1);  // the fold measured +0.3 % (stories15M 10173 → 10201 tok/s): a tiny kernel in a PDL chain is almost free
  c->stage_max_kb = std::max(0, std::min((int)(kGemvSmemStageMaxSolo / 1024), env_int("RAMA_GEMV_STAGE_KB", 110)));
  {
    const char* m = getenv("RAMA_ATTN");
    c->attn_cluster = !(m && strcmp(m, "split") == 0);
  }
  {
    const char* m = getenv("RAMA_STEP");
    c->persistent = m && strcmp(m, "persistent") == 0;
    c->cluster_step = m && strcmp(m, "cluster") == 0 ? 1 : (m && strcmp(m, "kernels") == 0 ? 0 : -1);
    c->cluster_step_ctas = std::max(2, std::min(16, env_int("RAMA_STEP_CLUSTER", 16)));
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
  if (c->world == 1) {
    const int sim = env_int("RAMA_TP_SIM", 0);
    if (sim > 1 && sim <= kMaxPeers) { c->world = sim; c->rank = 0; c->p2p = 1; c->tp_sim = 1; c->persistent = 0; }
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
  if (is_group(c)) {
    group_destroy(c);
    delete c;
    return RAMA_OK;
  }
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

// a group loads every rank's shard concurrently (each rank thread reads / generates only the rows it keeps)
template <class F>
static int group_load(rama_ctx* g, F&& load_rank) {
  if (g->n_objects.load() > 0) return fail(RAMA_E_STATE, "sessions or batches of this context are alive: destroy them before loading weights again");
  g->loaded = false;
  RK(group_run(g, [&](int r) { return load_rank(g->ranks[r]); }));
  group_adopt_config(g);
  return RAMA_OK;
}

extern "C" int rama_ctx_load_host(rama_ctx* c, const rama_config* cfg, const float* const tensors[RAMA_T_COUNT]) {
  if (!c || !cfg || !tensors) return fail(RAMA_E_INVALID, "NULL argument");
  if (is_group(c)) return group_load(c, [&](rama_ctx* rc) { return rama_ctx_load_host(rc, cfg, tensors); });
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
  if (is_group(c)) return group_load(c, [&](rama_ctx* rc) { return rama_ctx_load_file(rc, path); });
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
  if (is_group(c))
    return group_load(c, [&](rama_ctx* rc) { return rama_ctx_load_synthetic(rc, cfg, seed, scale, offset, freq_real, freq_imag); });
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
  if (is_group(c)) c = c->ranks[0];  // rank 0's shard
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
  if (is_group(c)) {  // what every rank can still allocate
    size_t f = ~(size_t)0, t = 0;
    for (rama_ctx* rc : c->ranks) {
      size_t fr, tr;
      RK(rama_ctx_mem_info(rc, &fr, &tr));
      f = std::min(f, fr);
      t = tr;
    }
    *free_bytes = f; *total_bytes = t;
    return RAMA_OK;
  }
  CK(cudaSetDevice(c->device));
  CK(cudaMemGetInfo(free_bytes, total_bytes));
  return RAMA_OK;
}

extern "C" int rama_ctx_weight_bytes(const rama_ctx* c, size_t* bytes) {
  if (!c || !bytes) return fail(RAMA_E_INVALID, "NULL argument");
  if (is_group(c)) {
    size_t t = 0;
    for (const rama_ctx* rc : c->ranks) {
      size_t b = 0;
      RK(rama_ctx_weight_bytes(rc, &b));
      t += b;
    }
    *bytes = t;
    return RAMA_OK;
  }
  size_t t = 0;
  for (int i = 0; i < RAMA_T_COUNT; ++i) t += c->plan[i].local_elems() * 4;
  *bytes = t;
  return RAMA_OK;
}

