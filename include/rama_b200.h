/* rama_b200.h — C ABI of the B200-native (sm_100a) llama2 f32 decode path.
 *
 * This is the drop-in boundary for ONE hot path of oliverhu/rama: the per-token decode loop
 * (engine/src/transformer/infer.rs:8-53 `forward`, engine/src/device/{device,cpu,gpu}.rs `Device`,
 * engine/src/transformer/mod.rs:169-248 `generate`).  Every entry point below names the
 * reference interface it replaces (paths relative to the reference repo root).  The Rust
 * binding a maintainer adds (gpu.rs / hbm.rs / build.rs) is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; device memory is passed as `float*` device pointers (rama_dev_alloc).
 *  - every function returns 0 (RAMA_OK) or a negative rama_status; the message is available
 *    from rama_last_error() (thread-local).  Nothing throws or aborts across the boundary; the
 *    Rust shim `.unwrap()`s to keep the reference's panic-on-error behaviour.
 *  - a rama_ctx is shared and thread-safe for concurrent sessions (reference: `GPU` lives in a
 *    static OnceLock, engine/src/lib.rs:56); a rama_session is used by one thread at a time
 *    (reference: one RunState per request task, engine/src/lib.rs:133-153).
 *  - there is no CPU fallback: without a CUDA device every call fails with RAMA_E_CUDA.
 */
#ifndef RAMA_B200_H
#define RAMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAMA_ABI_VERSION 1

typedef enum rama_status {
  RAMA_OK = 0,
  RAMA_E_INVALID = -1, /* bad argument / unsupported shape */
  RAMA_E_CUDA = -2,    /* CUDA runtime error (no device, launch failure, ...) */
  RAMA_E_IO = -3,      /* checkpoint file problem */
  RAMA_E_NCCL = -4,    /* NCCL unavailable or failed */
  RAMA_E_STATE = -5    /* call out of order (no weights loaded, pos >= seq_len, ...) */
} rama_status;

/* engine/src/transformer/mod.rs:128-138 `Config` (shared_weight: vocab>0 in the file header) */
typedef struct rama_config {
  int32_t dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len, shared_weight;
} rama_config;

/* The 14 tensors in llama2.c v0 file order (engine/export/export.py:75-127 ⇔
 * engine/src/transformer/ram.rs:30-49) ≙ fields of TransformerWeights (state.rs:54-74). */
enum rama_tensor {
  RAMA_T_TOKEN_EMBEDDING = 0, RAMA_T_RMS_ATT, RAMA_T_WQ, RAMA_T_WK, RAMA_T_WV, RAMA_T_WO,
  RAMA_T_RMS_FFN, RAMA_T_W1, RAMA_T_W2, RAMA_T_W3, RAMA_T_RMS_FINAL, RAMA_T_FREQ_REAL,
  RAMA_T_FREQ_IMAG, RAMA_T_WCLS, RAMA_T_COUNT
};

/* The 12 RunState buffers (engine/src/transformer/state.rs:4-17). */
enum rama_state_buf {
  RAMA_S_X = 0, RAMA_S_XB, RAMA_S_XB2, RAMA_S_HB, RAMA_S_HB2, RAMA_S_Q, RAMA_S_K, RAMA_S_V,
  RAMA_S_ATT, RAMA_S_LOGITS, RAMA_S_KEY_CACHE, RAMA_S_VALUE_CACHE, RAMA_S_COUNT
};

/* Tensor-parallel placement of this process (one process per GPU).  NULL ⇒ single GPU.
 * nccl_id is an ncclUniqueId made by rama_tp_unique_id() on rank 0 and distributed by the
 * caller (torch.distributed / files / sockets).  New relative to the reference, which has no
 * multi-GPU path (gpu.rs:215 `CudaDevice::new(0)`). */
typedef struct rama_tp {
  int32_t rank, world;
  uint8_t nccl_id[128];
} rama_tp;

typedef struct rama_ctx rama_ctx;         /* ≙ `GPU` + `TransformerWeights<Dev>` */
typedef struct rama_session rama_session; /* ≙ `RunState<Dev>` (+ stream, step graph) */

int rama_abi_version(void);
const char* rama_last_error(void);
int rama_device_count(int* n);

/* ---- context: device + weights -------------------------------------------------------- */

/* ≙ GPU::new() (engine/src/device/gpu.rs:213-234).  No NVRTC, no cuBLAS. */
int rama_ctx_create(int device, const rama_tp* tp, rama_ctx** out);
int rama_ctx_destroy(rama_ctx* ctx);
int rama_tp_unique_id(uint8_t out[128]);
/* ≙ GPU::new() for a caller that is ONE process with ONE device handle — the reference's engine binary and server
 * (engine/src/main.rs:70-98, engine/src/lib.rs:99-119) cannot launch a rank per GPU.  The returned context owns
 * n_gpus devices (devices[] or, when NULL, 0..n_gpus-1; all pairs must be peer-addressable) and is accepted by every
 * entry point below exactly like a single-device context: loads shard the weights over the devices (column-parallel
 * wq/wk/wv/w1/w3, row-parallel wo/w2, vocabulary-split classifier), a session holds one KV-cache shard per device, and
 * rama_forward / rama_generate / rama_prefill / rama_forward_batch drive every device's captured step from the calling
 * thread (one library-owned host thread per device issues its launches).  The devices exchange through NVLink peer
 * memory inside the kernels; NCCL is not involved.  Op-level calls (rama_op_*, rama_dev_*) run on the first device;
 * rama_state_to_host returns the first device's shard of sharded buffers (logits and x are complete).
 * n_gpus == 1 returns a plain single-device context. */
int rama_ctx_create_multi(int32_t n_gpus, const int32_t* devices, rama_ctx** out);

/* ≙ Config::from_file + TransformerWeights::from_file + ::from_weight
 * (mod.rs:140-166, ram.rs:27-51, hbm.rs:55-90): mmap the v0 .bin, stream it through pinned
 * staging into HBM once, keeping only this rank's shard under TP. */
int rama_ctx_load_file(rama_ctx* ctx, const char* path);
/* throughput of this rank's window through the last rama_ctx_load_file (tools/load_bench.py) */
int rama_last_load_gbps(double* out);
/* ≙ TransformerWeights::from_weight(&mut TransformerWeights<Vec<f32>>, &GPU) (hbm.rs:55-90).
 * tensors[] in rama_tensor order, full (unsharded) host arrays; tensors[WCLS] may be NULL when
 * cfg->shared_weight (state.rs:111-117: wcls then aliases the embedding). */
int rama_ctx_load_host(rama_ctx* ctx, const rama_config* cfg, const float* const tensors[RAMA_T_COUNT]);
/* Bench/test helper: fill the weights in HBM with the counter-based synthetic recipe of
 * rama_b200/checkpoint.py (bit-identical to the numpy and oracle generators).  scale[i] == 0
 * with offset[i] == 0 leaves tensor i to the host pointer given (RoPE tables). */
int rama_ctx_load_synthetic(rama_ctx* ctx, const rama_config* cfg, uint64_t seed,
                            const float scale[RAMA_T_COUNT], const float offset[RAMA_T_COUNT],
                            const float* freq_real, const float* freq_imag);
int rama_ctx_config(const rama_ctx* ctx, rama_config* out);
/* Copies this rank's shard of a weight tensor back (tests). n = capacity in floats; *n_out = shard size. */
int rama_ctx_weight_to_host(rama_ctx* ctx, int tensor, float* dst, size_t n, size_t* n_out);
int rama_ctx_weight_bytes(const rama_ctx* ctx, size_t* bytes);
/* free / total HBM of the context's device (sizing the number of concurrent sessions: one session of
 * llama2-7B holds a 2 GiB KV cache, ram.rs:20-21) */
int rama_ctx_mem_info(rama_ctx* ctx, size_t* free_bytes, size_t* total_bytes);

/* ---- session: RunState on the device --------------------------------------------------- */

/* ≙ RunState::from_config + RunState::from_state (ram.rs:6-23, hbm.rs:19-34), without the
 * host round trip: buffers (incl. the KV cache) are allocated and zeroed in HBM. */
int rama_session_create(rama_ctx* ctx, rama_session** out);
int rama_session_reset(rama_session* s);
int rama_session_destroy(rama_session* s);

/* ≙ forward(cfg, wv, rsv, token, pos, device) (infer.rs:8-53): one fused, CUDA-graph-replayed
 * decode step; asynchronous (returns after enqueue on the session stream). */
int rama_forward(rama_session* s, int32_t token, int32_t pos);
/* ≙ Device::sample(cfg, rsv, temperature, topp) -> usize (device.rs:16, cpu.rs:155-179,
 * infer.rs:55-85) run ON THE DEVICE; only the 4-byte token id crosses PCIe.  Synchronous.
 * As in the reference the logits buffer is overwritten with probabilities when temperature != 0. */
int rama_sample(rama_session* s, float temperature, float topp, int32_t* next);
/* ≙ generate(...) token loop (mod.rs:169-206) without tokenizer/printing: BOS at pos 0, prompt
 * forcing, then sampling; the token fed back never leaves the device.  out_tokens[pos] = `next`
 * of step pos.  Synchronous; *elapsed_ms (optional) = CUDA-event time of the step loop. */
int rama_generate(rama_session* s, const int32_t* prompt, int32_t n_prompt, int32_t steps,
                  float temperature, float topp, int32_t* out_tokens, float* elapsed_ms);
int rama_session_sync(rama_session* s);

/* Prompt prefill on the tensor cores: processes tokens[0..n) at positions [pos0, pos0+n) in one pass —
 * every weight matrix is read once per prompt, the contractions run as tcgen05 3xTF32 GEMMs, attention is
 * causal over the KV cache (a 3xTF32 mma.sync flash kernel; RAMA_PREFILL_ATTN=cuda selects the f32 CUDA-core one).  The reference has no such entry point: generate() feeds the prompt through
 * forward() one token at a time and discards the logits (mod.rs:187-192).  Afterwards the session is in the
 * state those n forward() calls would have left: KV-cache rows pos0..pos0+n-1 of every layer (infer.rs:31-33)
 * and the logits of the last position (infer.rs:51), so rama_sample / rama_forward(token, pos0+n) continue
 * from it.  Synchronous.  elapsed_ms (optional): CUDA-event time; ms_kind (optional): per-kind event times
 * (slower: events around every launch); n_launch (optional): kernels launched. */
enum rama_prefill_kind { RAMA_PK_GEMM = 0, RAMA_PK_ATTN, RAMA_PK_NORM, RAMA_PK_COMM, RAMA_PK_OTHER, RAMA_PK_COUNT };
int rama_prefill(rama_session* s, const int32_t* tokens, int32_t n, int32_t pos0, float* elapsed_ms,
                 float ms_kind[RAMA_PK_COUNT], int32_t* n_launch);
/* rama_generate prefills prompts of at least min_rows rows (BOS included; default 16; 0 = never: every
 * prompt token goes through the per-token step like the reference loop). */
int rama_session_set_prefill(rama_session* s, int32_t min_rows);

/* ---- batched multi-sequence decode (server path) ------------------------------------------------
 * The reference server runs one RunState and one forward()/sample() loop per request task
 * (engine/src/lib.rs:127-160; the batcher in server/src/batcher.rs:8-38 is dead code), so n concurrent
 * requests stream the weights n times per step.  rama_forward_batch is n concurrent forward() calls in one
 * pass: the n token vectors go through every weight matrix together on the tensor cores (weights read once
 * per step), each sequence at its own position with its own session's KV cache.  Afterwards every
 * session's logits buffer holds its logits, exactly as after rama_forward(session, token, pos), so
 * rama_sample / rama_logits_to_host work per session; rama_sample_batch samples all of them in one launch.
 * A batch object owns the workspace and the stream; sessions in a batch must not be used concurrently
 * through their own entry points (one after the other is fine: the library orders the batch's stream and the
 * sessions' streams with events, no explicit rama_batch_sync / rama_session_sync is needed in between).
 * Works on one GPU, under tensor parallelism with one process per GPU (every rank calls it with its own sessions,
 * collectively) and on a multi-device context of rama_ctx_create_multi. */
typedef struct rama_batch rama_batch;
int rama_batch_create(rama_ctx* ctx, int32_t max_seqs /* ≤ 64 */, rama_batch** out);
int rama_batch_destroy(rama_batch* b);
/* asynchronous on the batch's stream */
int rama_forward_batch(rama_batch* b, rama_session* const* sessions, const int32_t* tokens, const int32_t* pos,
                       int32_t n);
/* ≙ Device::sample for each of the n sessions (same temperature/topp), one launch; synchronous */
int rama_sample_batch(rama_batch* b, rama_session* const* sessions, int32_t n, float temperature, float topp,
                      int32_t* next);
/* ≙ n generate() loops (mod.rs:169-206) advanced together with the token feedback on the device: every sequence starts at
 * BOS / position 0, prompts[i][0..n_prompt[i]) is forced, `steps` tokens per sequence are written to out_tokens[i*steps + t]
 * (same temperature / topp for all, as rama_sample_batch).  One captured graph per batch size holds the batched step and the n
 * samplers; nothing crosses PCIe between steps.  elapsed_ms = device time of the step loop.  Synchronous. */
int rama_generate_batch(rama_batch* b, rama_session* const* sessions, int32_t n, const int32_t* const* prompts,
                        const int32_t* n_prompt, int32_t steps, float temperature, float topp, int32_t* out_tokens,
                        float* elapsed_ms);
int rama_batch_sync(rama_batch* b);
int rama_batch_launches_per_step(const rama_batch* b, int32_t* n);

/* ≙ Device::to_cpu / RunState::into_state (device.rs:21, gpu.rs:196-209, hbm.rs:38-51).
 * buf in rama_state_buf; layouts as the reference's RunState (KV cache [L][T][D]).  Under TP the
 * sharded buffers (q,k,v,hb,hb2,att,caches) hold this rank's slice.  RAMA_S_ATT is only kept
 * when rama_session_set_debug(s, 1). */
int rama_state_to_host(rama_session* s, int buf, float* dst, size_t n, size_t* n_out);
int rama_logits_to_host(rama_session* s, float* dst, size_t n);
int rama_session_set_debug(rama_session* s, int keep_att);
/* number of kernels one step enqueues (bench `gpu_launches`). */
int rama_session_launches_per_step(const rama_session* s, int* n);

/* Tracing: runs `forward(token,pos)` un-graphed with a CUDA-event pair around every kernel and
 * returns the summed milliseconds and launch counts per kernel kind. */
enum rama_kernel_kind {
  RAMA_K_EMBED = 0, RAMA_K_QKV, RAMA_K_ATTN, RAMA_K_WO, RAMA_K_W13, RAMA_K_W2, RAMA_K_CLS,
  RAMA_K_SAMPLE, RAMA_K_COMM, RAMA_K_COUNT
};
int rama_profile_step(rama_session* s, int32_t token, int32_t pos, float ms[RAMA_K_COUNT],
                      int32_t launches[RAMA_K_COUNT]);
/* In-graph timeline of one decode step: the step is captured into a CUDA graph exactly like the production one (programmatic
 * dependent launch included) with a %globaltimer slot per kernel; CTA 0 of kernel i stamps stamps_ns[4i + 0] entry,
 * [4i + 1] dependency resolved (the previous kernel of the chain has completed), [4i + 2] prologue done (activations in shared
 * memory; under tensor parallelism: every peer's partial has arrived), [4i + 3] its own end — nanoseconds relative to the first
 * stamp of the last of `reps` replays; kinds[i] is the rama_kernel_kind.  mode 0 = forward, 1 = + greedy sampler.
 * cap ≥ 5·n_layers + 3.  Collective under tensor parallelism.  (rama_profile_step's event pairs include launch latency and
 * remove all overlap; this is the attribution to trust.) */
int rama_step_timeline(rama_session* s, int32_t token, int32_t pos, int32_t mode, int32_t reps, double* stamps_ns,
                       int32_t* kinds, int32_t cap, int32_t* n_out);

/* Phase timeline of one persistent step (tools/step_trace.py): SM-clock stamps of CTA 0 at kernel entry and
 * before/after each of the 5L+1 grid barriers. */
int rama_step_trace(rama_session* s, int32_t token, int32_t pos, long long* stamps, int32_t cap, int32_t* n_out);

/* ---- tokenizer: the text boundary of generate() (engine/src/tokenizer/bpe.rs) ------------------------
 * Host code (the reference tokenizes on the CPU as well).  tokenizer.bin keeps the llama2.c layout byte for
 * byte: u32 max_token_length, then per token f32 score, i32 length, bytes (bpe.rs:27-43). */
typedef struct rama_tokenizer rama_tokenizer;
int rama_tokenizer_load(const char* path, int32_t vocab_size, rama_tokenizer** out);   /* Tokenizer::new, bpe.rs:19-45 */
int rama_tokenizer_free(rama_tokenizer* t);
int rama_tokenizer_info(const rama_tokenizer* t, int32_t* vocab_size, int32_t* max_token_length);
/* Tokenizer::encode (bpe.rs:50-97): trim, drop '\n', per-char lookup, greedy best-score merges.  out may be NULL
 * to query *n_out.  Inputs the reference panics on (unknown char, nothing left to encode) are RAMA_E_INVALID. */
int rama_tokenizer_encode(const rama_tokenizer* t, const char* text, int32_t* out, int32_t cap, int32_t* n_out);
/* decode(vocab[token]) (bpe.rs:102-116): "<s>" → "", "<0xAB>" → char::from(0xAB) as UTF-8, else the piece. */
int rama_tokenizer_decode(const rama_tokenizer* t, int32_t token, char* out, int32_t cap, int32_t* n_out);

/* ---- op level: 1:1 with `trait Device<T>` (engine/src/device/device.rs:3-24) -------------
 * Pointers are device pointers already offset by the view's range.start (a View is
 * (storage, absolute range), mod.rs:16-51).  Asynchronous on the ctx's op stream;
 * rama_dev_d2h / rama_ctx_sync synchronise. */
int rama_dev_alloc(rama_ctx* ctx, size_t n_floats, float** out);          /* hbm.rs:14-16 allocate */
int rama_dev_free(rama_ctx* ctx, float* p);
int rama_dev_h2d(rama_ctx* ctx, float* dst, const float* src, size_t n);  /* htod_sync_copy */
int rama_dev_d2h(rama_ctx* ctx, float* dst, const float* src, size_t n);  /* dtoh_sync_copy_into */
int rama_ctx_sync(rama_ctx* ctx);

int rama_op_array_add(rama_ctx* ctx, float* target, const float* source, size_t n);   /* device.rs:4 */
int rama_op_array_mult(rama_ctx* ctx, float* target, const float* source, size_t n);  /* device.rs:5 */
int rama_op_sinu(rama_ctx* ctx, float* o, size_t n);                                  /* device.rs:6 */
/* device.rs:7-8; buffers as RunStateView fields; cfg gives dim/n_heads/seq_len. att may be NULL. */
int rama_op_multi_head_attention(rama_ctx* ctx, float* xb, float* att, const float* q,
                                 const float* key_cache, const float* value_cache,
                                 const rama_config* cfg, int32_t layer, int32_t pos);
int rama_op_copy_from_slice(rama_ctx* ctx, float* target, const float* source, size_t n); /* device.rs:9 */
int rama_op_rmsnorm(rama_ctx* ctx, float* o, const float* x, const float* weight, size_t n); /* device.rs:10-11 */
int rama_op_apply_position(rama_ctx* ctx, float* q, float* k, const float* pos_real,
                           const float* pos_img, size_t head_size);                  /* device.rs:12 */
/* device.rs:13: o[r*o_cols+c] = Σ_i a[r*width+i]·b[i*o_cols+c]; the hot path only uses o_cols=1. */
int rama_op_matmul(rama_ctx* ctx, float* o, const float* a, const float* b, size_t width,
                   size_t o_rows, size_t o_cols);
int rama_op_softmax(rama_ctx* ctx, float* x, size_t n);                               /* device.rs:14 */
/* device.rs:16 on raw logits (in place, like the reference). */
int rama_op_sample(rama_ctx* ctx, float* logits, size_t vocab_size, float temperature, float topp,
                   int32_t* next);

/* Dense contraction on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split for f32 accuracy):
 *   out[M][N] = a[M][K] · b[N][K]^T          (both operands row-major with K contiguous)
 * The shape the reference's Device::matmul (device.rs:13, cpu.rs:127-153) takes when a whole prompt or a
 * batch of sequences goes through one weight matrix instead of o_cols = 1 per token (mod.rs:187-192,
 * lib.rs:127-160).  variant selects the tile configuration (tests sweep them); flags bit 0: keep the raw
 * f32 tile as the hi operand (hardware truncation) instead of rounding it; bit 1: store out^T, i.e.
 * out[N][M] (batched decode orientation: a = weights, b = activations). */
int rama_op_matmul_nt(rama_ctx* ctx, float* out, const float* a, const float* b, size_t M, size_t N, size_t K,
                      int variant, int flags);

/* Micro-benchmark hook used by tools/gemm_sweep.py: average milliseconds per rama_op_matmul_nt launch. */
int rama_bench_matmul_nt(rama_ctx* ctx, float* out, const float* a, const float* b, size_t M, size_t N, size_t K,
                         int variant, int flags, int iters, float* avg_ms);

/* Debug hook used by tools/gemm_trace.py: one rama_op_matmul_nt launch whose CTA (0,0,0) stamps clock64() per k-block and
 * role into host_trace[128][8] (0 TMA issue, 1 tile landed, 2 split done, 3 MMA sees ready, 4 MMAs issued,
 * 5/6 drain wait begin/end per (chunk, worker group), 7 MMA got its accumulator back). */
int rama_debug_gemm_trace(rama_ctx* ctx, float* out, const float* a, const float* b, size_t M, size_t N, size_t K,
                          int variant, int flags, long long* host_trace);

/* Synthetic fill of a device buffer (bench/test data): elements [start, start+n) of tensor_id. */
int rama_synth_fill(rama_ctx* ctx, float* dst, size_t n, uint64_t seed, uint64_t tensor_id,
                    uint64_t start, float scale, float offset);

/* Micro-benchmark hook used by tools/gemv_sweep.py: runs the plain GEMV kernel `iters` times and
 * returns the average milliseconds per launch (CUDA events on its stream).  `w` holds n_mats
 * matrices back to back; launch i reads matrix i % n_mats, so a set larger than L2 stays cold. */
int rama_bench_gemv(rama_ctx* ctx, float* o, const float* w, const float* x, size_t rows, size_t width,
                    size_t n_mats, int variant, int iters, float* avg_ms);

#ifdef __cplusplus
}
#endif
#endif /* RAMA_B200_H */
