// ops.cu — the op-level Device entry points (1:1 with the reference trait, device.rs:3-24), raw device buffers, and the
// measurement hooks of the tools.
#include "internal.cuh"

#include "attention.cuh"
#include "gemm_host.cuh"
#include "batch.cuh"

// ------------------------------------------------------------------------------------------------
// op level (Device trait)
// ------------------------------------------------------------------------------------------------
extern "C" int rama_dev_alloc(rama_ctx* c, size_t n, float** out) {
  if (!c || !out) return fail(RAMA_E_INVALID, "NULL argument");
  c = rank0(c);
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
  c = rank0(c);
  CK(cudaSetDevice(c->device));
  CK(cudaFree(p));
  return RAMA_OK;
}
extern "C" int rama_dev_h2d(rama_ctx* c, float* dst, const float* src, size_t n) {
  if (!c || (!dst && n) || (!src && n)) return fail(RAMA_E_INVALID, "NULL argument");
  c = rank0(c);
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}
extern "C" int rama_dev_d2h(rama_ctx* c, float* dst, const float* src, size_t n) {
  if (!c || (!dst && n) || (!src && n)) return fail(RAMA_E_INVALID, "NULL argument");
  c = rank0(c);
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, c->op_stream));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}
extern "C" int rama_ctx_sync(rama_ctx* c) {
  if (!c) return fail(RAMA_E_INVALID, "NULL ctx");
  c = rank0(c);
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->op_stream));
  return RAMA_OK;
}

static int ew_grid(const rama_ctx* c, size_t n) {
  return (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8));
}
// (a group context runs the single-device Device ops on its first rank's device)
#define OP_PRE(c)                                         \
  if (!(c)) return fail(RAMA_E_INVALID, "NULL ctx");      \
  (c) = rank0(c);                                         \
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

// rows 0..N-1 of dst = src, rows plane.. = the tf32 remainders (pre-split B operand, gemm_tf32x3.cuh PS mode)
static __global__ void split_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t N, size_t K, size_t plane) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N * K; i += (size_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    dst[i] = v;
    dst[plane + i] = tf32_lo(v);
  }
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
      case 6: case 7: case 8: case 9: {  // pre-split B (the batched-decode tile of round 2): build the [128][K] activations + remainder matrix first
        if (N > 64) return fail(RAMA_E_INVALID, "matmul_nt: the pre-split variant takes N ≤ 64 columns");
        float* bs = nullptr;
        CK(cudaMallocAsync((void**)&bs, (size_t)128 * K * sizeof(float), c->op_stream));
        CK(cudaMemsetAsync(bs, 0, (size_t)128 * K * sizeof(float), c->op_stream));
        split_rows_kernel<<<c->sm_count * 2, 256, 0, c->op_stream>>>(bs, b, N, K, (size_t)64 * K);
        GemmOperand Bs{bs, 128, K};
        if (variant == 6) e = launch_gemm_tf32x3<64, 4, 4, 0, 8, 1>(c->op_stream, &A, 1, &Bs, 1, m, n, k, hi_round, ksplit, epi);
        else if (variant == 7) e = launch_gemm_tf32x3<64, 4, 4, 0, 8, 2>(c->op_stream, &A, 1, &Bs, 1, m, n, k, hi_round, ksplit, epi);  // all products wide
        // two CTAs per SM: one accumulator, two B / TMEM-A stages and a 4-slot weight ring each (96 KB, 256 TMEM columns)
        else if (variant == 8) e = launch_gemm_tf32x3<64, 2, 4, 0, 4, 1, 1>(c->op_stream, &A, 1, &Bs, 1, m, n, k, hi_round, ksplit, epi);
        else e = launch_gemm_tf32x3<64, 2, 4, 0, 4, 2, 1>(c->op_stream, &A, 1, &Bs, 1, m, n, k, hi_round, ksplit, epi);
        CK(cudaFreeAsync(bs, c->op_stream));
        break;
      }
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
      // CTA pair (tcgen05.mma.cta_group::2): 256×128 tile over two SMs, each holding half of the B k-block
      case 4: e = launch_gemm_tf32x3<128, 4, 4, 0, 0, 0, 2, 1>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, 1, epi); break;
      case 5: e = launch_gemm_tf32x3<128, 4, 2, 0, 0, 0, 2, 1>(c->op_stream, &A, 1, &B, 1, m, n, k, hi_round, 1, epi); break;
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

