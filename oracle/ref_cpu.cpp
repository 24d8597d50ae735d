// oracle/ref_cpu.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's (oliverhu/rama) per-token decode path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (rama_b200/csrc) never links or calls it.
//
// The reference is Rust and cannot be compiled in this image (no cargo/rustc), so
// this file restates its algorithm operation by operation.  Each function cites the
// reference file:line it follows (paths relative to /root/reference).
//
// PINNING STATUS: the reference holds no golden vector / KAT for this path
// (SURVEY.md §4, §8c).  The restatement is pinned against the one part of the
// reference that CAN run here: engine/export/model.py + engine/export/export.py
// (legacy_export) — see tests/golden/make_golden.py and tests/test_oracle_golden.py.
// Against the Rust CPU device itself parity is UNPINNED (nothing to run it with);
// the three unverifiable details are exposed as switches below:
//   * wide::f32x4::reduce_add lane order      (ref_set_reduce_mode)
//   * rayon par_iter().sum() association      (sequential here)
//   * rand/rand_chacha versions ("*")         (ref_chacha_first_f32 derives the constant)
//
// Build: see oracle/Makefile (-O3 -march=x86-64-v3 -ffp-contract=off -fopenmp).
// -ffp-contract=off matters: Rust never fuses a*b+c, GCC on x86+FMA would.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <immintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

int g_reduce_mode = 0;  // 0: ((l0+l1)+l2)+l3   1: (l0+l1)+(l2+l3)   2: (l0+l2)+(l1+l3)

inline float reduce_add4(const float l[4]) {
  switch (g_reduce_mode) {
    case 1: return (l[0] + l[1]) + (l[2] + l[3]);
    case 2: return (l[0] + l[2]) + (l[1] + l[3]);
    default: return ((l[0] + l[1]) + l[2]) + l[3];
  }
}

// engine/src/transformer/mod.rs:128-167
struct Config {
  int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len, shared_weight;
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------
// Device ops — engine/src/device/cpu.rs
// ---------------------------------------------------------------------------------

void ref_set_reduce_mode(int mode) { g_reduce_mode = mode; }

void ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int ref_get_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// cpu.rs:16-21   *a += *b
void ref_array_add(float* target, const float* source, int n) {
  for (int i = 0; i < n; ++i) target[i] = target[i] + source[i];
}

// cpu.rs:59-64   *a *= *b
void ref_array_mult(float* target, const float* source, int n) {
  for (int i = 0; i < n; ++i) target[i] = target[i] * source[i];
}

// cpu.rs:54-57   a = a * (1.0 / (1.0 + exp(-a)))
void ref_sinu(float* o, int n) {
  for (int i = 0; i < n; ++i) {
    float a = o[i];
    o[i] = a * (1.0f / (1.0f + expf(-a)));
  }
}

// cpu.rs:66-72
void ref_copy_from_slice(float* target, const float* source, int n) {
  memmove(target, source, sizeof(float) * (size_t)n);
}

// cpu.rs:99-117   v = 1/sqrt(sum(x*x)/n + 1e-5);  o[i] = w[i] * (v * x[i])
void ref_rmsnorm(float* o, const float* x, const float* weight, int n) {
  float ss = 0.0f;  // Iterator::sum::<f32>() is a sequential left fold
  for (int i = 0; i < n; ++i) ss = ss + x[i] * x[i];
  float v = 1.0f / sqrtf(ss / (float)n + 1e-5f);
  for (int i = 0; i < n; ++i) o[i] = weight[i] * (v * x[i]);
}

// cpu.rs:74-97   RoPE on adjacent pairs of ONE head of q and k, simultaneous pair update
void ref_apply_position(float* q, float* k, const float* pos_real, const float* pos_img,
                        int head_size) {
  for (int i = 0; i < head_size / 2; ++i) {
    float fcr = pos_real[i], fci = pos_img[i];
    float q0 = q[2 * i], q1 = q[2 * i + 1];
    q[2 * i] = q0 * fcr - q1 * fci;
    q[2 * i + 1] = q0 * fci + q1 * fcr;
    float k0 = k[2 * i], k1 = k[2 * i + 1];
    k[2 * i] = k0 * fcr - k1 * fci;
    k[2 * i + 1] = k0 * fci + k1 * fcr;
  }
}

// cpu.rs:127-153   o[idx] = reduce_add( Σ_k f32x4(a[r,k..k+4]) * f32x4(b[(k+j)*o_cols+c]) )
// n_out = length of the output view (the reference ignores o_rows).  width % 4 == 0.
// rayon over outputs ⇒ each output is an independent sequential chain: threading here
// (OpenMP over outputs) cannot change any result bit.
void ref_matmul(float* o, const float* a, const float* b, int width, int n_out, int o_cols) {
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < n_out; ++idx) {
    int r = idx / o_cols, c = idx % o_cols;
    const float* ar = a + (size_t)r * width;
    __m128 v = _mm_setzero_ps();
    if (o_cols == 1) {
      for (int k = 0; k < width; k += 4) {
        __m128 aw = _mm_loadu_ps(ar + k);
        __m128 bw = _mm_loadu_ps(b + k);
        v = _mm_add_ps(v, _mm_mul_ps(aw, bw));  // separate IEEE mul then add, no FMA
      }
    } else {
      for (int k = 0; k < width; k += 4) {
        __m128 aw = _mm_loadu_ps(ar + k);
        __m128 bw = _mm_set_ps(b[(size_t)(k + 3) * o_cols + c], b[(size_t)(k + 2) * o_cols + c],
                               b[(size_t)(k + 1) * o_cols + c], b[(size_t)k * o_cols + c]);
        v = _mm_add_ps(v, _mm_mul_ps(aw, bw));
      }
    }
    float l[4];
    _mm_storeu_ps(l, v);
    o[idx] = reduce_add4(l);
  }
}

// cpu.rs:187-192 (softmax_num) and cpu.rs:119-125 (softmax)
// max → exp(a-max) → sum → a/sum.  The reference's sum is rayon's (unspecified association);
// sequential here.
void ref_softmax(float* x, int n) {
  float mx = x[0];
  for (int i = 0; i < n; ++i) mx = fmaxf(mx, x[i]);
  for (int i = 0; i < n; ++i) x[i] = expf(x[i] - mx);
  float sum = 0.0f;
  for (int i = 0; i < n; ++i) sum = sum + x[i];
  for (int i = 0; i < n; ++i) x[i] = x[i] / sum;
}

// cpu.rs:23-52.  q [H*hs]; key/value cache [L][T][D]; att [H][T]; xb [D].
void ref_multi_head_attention(float* xb, float* att, const float* q, const float* key_cache,
                              const float* value_cache, const Config* cfg, int layer, int pos) {
  const int D = cfg->dim, H = cfg->n_heads, T = cfg->seq_len;
  const int hs = D / H;
  const size_t lo = (size_t)layer * T * D;
  const float sq = sqrtf((float)hs);
#pragma omp parallel for schedule(static)
  for (int h = 0; h < H; ++h) {
    float* a = att + (size_t)h * T;
    const float* qh = q + (size_t)h * hs;
    for (int t = 0; t <= pos; ++t) {
      const float* k = key_cache + lo + (size_t)t * D + (size_t)h * hs;
      float s = 0.0f;
      for (int i = 0; i < hs; ++i) s = s + qh[i] * k[i];
      a[t] = s / sq;
    }
    ref_softmax(a, pos + 1);
    float* o = xb + (size_t)h * hs;
    for (int i = 0; i < hs; ++i) o[i] = 0.0f;
    for (int t = 0; t <= pos; ++t) {
      const float* v = value_cache + lo + (size_t)t * D + (size_t)h * hs;
      float w = a[t];
      for (int i = 0; i < hs; ++i) o[i] = o[i] + w * v[i];
    }
  }
}

// ---------------------------------------------------------------------------------
// Sampler — cpu.rs:155-179, infer.rs:55-85, Appendix B of SURVEY.md
// ---------------------------------------------------------------------------------

// rand_core::SeedableRng::seed_from_u64 (PCG32 expansion) + ChaCha20 block 0 word 0,
// then Standard f32 = (w >> 8) * 2^-24.  Versions of rand/rand_chacha are unpinned ("*",
// engine/Cargo.toml:15,17); this follows rand 0.8 / rand_core 0.6.
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
uint32_t ref_chacha_first_u32(uint64_t seed) {
  uint32_t key[8];
  uint64_t state = seed;
  for (int i = 0; i < 8; ++i) {
    state = state * 6364136223846793005ULL + 11634580027462260723ULL;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
  uint32_t in[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
  for (int i = 0; i < 8; ++i) in[4 + i] = key[i];
  in[12] = in[13] = in[14] = in[15] = 0;
  uint32_t x[16];
  memcpy(x, in, sizeof(x));
#define QR(a, b, c, d)                                   \
  x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 16);   \
  x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 12);   \
  x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 8);    \
  x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 7);
  for (int r = 0; r < 10; ++r) {
    QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
    QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
  }
#undef QR
  return x[0] + in[0];
}
float ref_chacha_first_f32(uint64_t seed) {
  return (float)(ref_chacha_first_u32(seed) >> 8) * (1.0f / 16777216.0f);
}

// infer.rs:55-85 with the rng draw passed in (it is a constant, cpu.rs:161-162).
// Returns -1 where the reference would panic (empty candidate list, infer.rs:66; NaN in sort).
int ref_sample_top_q(const float* probabilities, int num, float topp, float u) {
  float cutoff = (1.0f - topp) / (float)(num - 1);
  std::vector<std::pair<int, float>> pi;
  for (int i = 0; i < num; ++i) {
    if (probabilities[i] != probabilities[i]) return -1;
    if (probabilities[i] > cutoff) pi.emplace_back(i, probabilities[i]);
  }
  if (pi.empty()) return -1;
  // slice::sort_by is stable; descending by p
  std::stable_sort(pi.begin(), pi.end(),
                   [](const std::pair<int, float>& a, const std::pair<int, float>& b) {
                     return a.second > b.second;
                   });
  float cum = 0.0f;
  size_t last = pi.size() - 1;
  for (size_t i = 0; i < pi.size(); ++i) {
    cum = cum + pi[i].second;
    if (cum > topp) { last = i; break; }
  }
  float r = u * cum;
  float cdf = 0.0f;
  for (size_t i = 0; i < last; ++i) {
    cdf = cdf + pi[i].second;
    if (r < cdf) return pi[i].first;
  }
  return pi[last].first;
}

// cpu.rs:155-179.  Mutates logits in place exactly as the reference does.
int ref_sample(float* logits, int vocab_size, float temperature, float topp) {
  if (temperature == 0.0f) {
    // reduce(|(i1,v1),(i2,v2)| if v1 > v2 {a} else {b}): ties and NaN → later index
    int bi = 0;
    float bv = logits[0];
    for (int i = 1; i < vocab_size; ++i) {
      if (bv > logits[i]) { /* keep */ } else { bi = i; bv = logits[i]; }
    }
    return bi;
  }
  if (temperature < 1.0f)
    for (int i = 0; i < vocab_size; ++i) logits[i] = logits[i] / temperature;
  ref_softmax(logits, vocab_size);
  return ref_sample_top_q(logits, vocab_size, topp, ref_chacha_first_f32(100));
}

// ---------------------------------------------------------------------------------
// Weights / RunState — transformer/state.rs, ram.rs, utils/read.rs
// ---------------------------------------------------------------------------------

enum { T_EMB = 0, T_RMS_ATT, T_WQ, T_WK, T_WV, T_WO, T_RMS_FFN, T_W1, T_W2, T_W3, T_RMS_FINAL,
       T_FREQ_REAL, T_FREQ_IMAG, T_WCLS, T_COUNT };

struct RefModel {
  Config cfg;
  const float* t[T_COUNT];
  void* map = nullptr;
  size_t map_len = 0;
  std::vector<float> owned;  // optional owned storage
};

// element counts in llama2.c v0 file order (export.py:75-127 ⇔ ram.rs:30-49)
void ref_tensor_sizes(const Config* c, int64_t out[T_COUNT]) {
  int64_t D = c->dim, F = c->hidden_dim, L = c->n_layers, V = c->vocab_size, T = c->seq_len;
  int64_t hs = D / c->n_heads;
  out[T_EMB] = V * D; out[T_RMS_ATT] = L * D;
  out[T_WQ] = out[T_WK] = out[T_WV] = out[T_WO] = L * D * D;
  out[T_RMS_FFN] = L * D;
  out[T_W1] = out[T_W2] = out[T_W3] = L * D * F;
  out[T_RMS_FINAL] = D;
  out[T_FREQ_REAL] = out[T_FREQ_IMAG] = T * (hs / 2);
  out[T_WCLS] = c->shared_weight ? 0 : V * D;
}

// mod.rs:140-166: 7 LE i32; vocab > 0 ⇒ shared classifier
int ref_parse_header(const int32_t h[7], Config* c) {
  c->dim = h[0]; c->hidden_dim = h[1]; c->n_layers = h[2]; c->n_heads = h[3];
  c->n_kv_heads = h[4];
  c->shared_weight = h[5] > 0 ? 1 : 0;
  c->vocab_size = h[5] > 0 ? h[5] : -h[5];
  c->seq_len = h[6];
  return 0;
}

RefModel* ref_model_from_host(const Config* cfg, const float* const tensors[T_COUNT]) {
  RefModel* m = new RefModel();
  m->cfg = *cfg;
  for (int i = 0; i < T_COUNT; ++i) m->t[i] = tensors[i];
  // state.rs:111-117: wcls aliases the embedding when shared
  if (cfg->shared_weight || !m->t[T_WCLS]) m->t[T_WCLS] = m->t[T_EMB];
  return m;
}

RefModel* ref_model_from_file(const char* path) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return nullptr;
  struct stat st;
  fstat(fd, &st);
  void* p = mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return nullptr;
  RefModel* m = new RefModel();
  m->map = p; m->map_len = st.st_size;
  ref_parse_header((const int32_t*)p, &m->cfg);
  int64_t sz[T_COUNT];
  ref_tensor_sizes(&m->cfg, sz);
  const float* f = (const float*)((const char*)p + 28);
  for (int i = 0; i < T_COUNT; ++i) { m->t[i] = f; f += sz[i]; }
  if (m->cfg.shared_weight) m->t[T_WCLS] = m->t[T_EMB];
  if ((const char*)f > (const char*)p + st.st_size) { munmap(p, st.st_size); delete m; return nullptr; }
  return m;
}

void ref_model_config(const RefModel* m, Config* out) { *out = m->cfg; }
const float* ref_model_tensor(const RefModel* m, int id) { return m->t[id]; }
void ref_model_free(RefModel* m) {
  if (!m) return;
  if (m->map) munmap(m->map, m->map_len);
  delete m;
}

// state.rs:4-17, ram.rs:6-23
struct RefState {
  std::vector<float> x, xb, xb2, hb, hb2, q, k, v, att, logits, key_cache, value_cache;
};
enum { S_X = 0, S_XB, S_XB2, S_HB, S_HB2, S_Q, S_K, S_V, S_ATT, S_LOGITS, S_KEY, S_VALUE };

RefState* ref_state_create(const RefModel* m) {
  const Config& c = m->cfg;
  RefState* s = new RefState();
  size_t kv_dim = (size_t)c.dim * c.n_kv_heads / c.n_heads;
  s->x.assign(c.dim, 0.f); s->xb.assign(c.dim, 0.f); s->xb2.assign(c.dim, 0.f);
  s->hb.assign(c.hidden_dim, 0.f); s->hb2.assign(c.hidden_dim, 0.f);
  s->q.assign(c.dim, 0.f); s->k.assign(c.dim, 0.f); s->v.assign(c.dim, 0.f);
  s->att.assign((size_t)c.n_heads * c.seq_len, 0.f);
  s->logits.assign(c.vocab_size, 0.f);
  s->key_cache.assign((size_t)c.n_layers * c.seq_len * kv_dim, 0.f);
  s->value_cache.assign((size_t)c.n_layers * c.seq_len * kv_dim, 0.f);
  return s;
}
void ref_state_free(RefState* s) { delete s; }
float* ref_state_ptr(RefState* s, int which, int64_t* n) {
  std::vector<float>* v[] = {&s->x, &s->xb, &s->xb2, &s->hb, &s->hb2, &s->q, &s->k, &s->v,
                             &s->att, &s->logits, &s->key_cache, &s->value_cache};
  if (n) *n = (int64_t)v[which]->size();
  return v[which]->data();
}

// ---------------------------------------------------------------------------------
// forward — engine/src/transformer/infer.rs:8-53
// ---------------------------------------------------------------------------------
void ref_forward(const RefModel* m, RefState* s, int token, int pos) {
  const Config& c = m->cfg;
  const int D = c.dim, F = c.hidden_dim, H = c.n_heads, hs = D / H;
  ref_copy_from_slice(s->x.data(), m->t[T_EMB] + (size_t)token * D, D);            // :13
  const float* pos_real = m->t[T_FREQ_REAL] + (size_t)pos * (hs / 2);              // :15
  const float* pos_img = m->t[T_FREQ_IMAG] + (size_t)pos * (hs / 2);               // :16
  for (int l = 0; l < c.n_layers; ++l) {
    ref_rmsnorm(s->xb.data(), s->x.data(), m->t[T_RMS_ATT] + (size_t)l * D, D);    // :19
    // :20-21 issues wq twice; the second call recomputes the same values — issued once here.
    ref_matmul(s->q.data(), m->t[T_WQ] + (size_t)l * D * D, s->xb.data(), D, D, 1);
    ref_matmul(s->k.data(), m->t[T_WK] + (size_t)l * D * D, s->xb.data(), D, D, 1);  // :22
    ref_matmul(s->v.data(), m->t[T_WV] + (size_t)l * D * D, s->xb.data(), D, D, 1);  // :23
    for (int h = 0; h < H; ++h)                                                      // :25-29
      ref_apply_position(s->q.data() + h * hs, s->k.data() + h * hs, pos_real, pos_img, hs);
    size_t lo = (size_t)l * c.seq_len * D;                                           // :31
    ref_copy_from_slice(s->key_cache.data() + lo + (size_t)pos * D, s->k.data(), D);   // :32
    ref_copy_from_slice(s->value_cache.data() + lo + (size_t)pos * D, s->v.data(), D); // :33
    ref_multi_head_attention(s->xb.data(), s->att.data(), s->q.data(), s->key_cache.data(),
                             s->value_cache.data(), &c, l, pos);                     // :34
    ref_matmul(s->xb2.data(), m->t[T_WO] + (size_t)l * D * D, s->xb.data(), D, D, 1);  // :35
    ref_array_add(s->x.data(), s->xb2.data(), D);                                    // :37
    ref_rmsnorm(s->xb.data(), s->x.data(), m->t[T_RMS_FFN] + (size_t)l * D, D);      // :39
    ref_matmul(s->hb.data(), m->t[T_W1] + (size_t)l * F * D, s->xb.data(), D, F, 1);   // :41
    ref_matmul(s->hb2.data(), m->t[T_W3] + (size_t)l * F * D, s->xb.data(), D, F, 1);  // :42
    ref_sinu(s->hb.data(), F);                                                       // :44
    ref_array_mult(s->hb.data(), s->hb2.data(), F);                                  // :45
    ref_matmul(s->xb.data(), m->t[T_W2] + (size_t)l * D * F, s->hb.data(), F, D, 1);   // :46
    ref_array_add(s->x.data(), s->xb.data(), D);                                     // :47
  }
  ref_copy_from_slice(s->xb.data(), s->x.data(), D);                                 // :49
  ref_rmsnorm(s->x.data(), s->xb.data(), m->t[T_RMS_FINAL], D);                      // :50
  ref_matmul(s->logits.data(), m->t[T_WCLS], s->x.data(), D, c.vocab_size, 1);       // :51
}

// ---------------------------------------------------------------------------------
// generate — engine/src/transformer/mod.rs:169-206 (token ids only; text via ref_tok_*)
// out_tokens[pos] = `next` of that step (forced prompt token or sample).
// out_logits (optional) receives steps*vocab logits (pre-sampling) for teacher-forced checks.
// min_gap (optional) = min over sampled steps of (top1 - top2) logit.
// Returns elapsed seconds of the step loop.
// ---------------------------------------------------------------------------------
double ref_generate(const RefModel* m, RefState* s, const int32_t* prompt, int n_prompt,
                    int steps, float temperature, float topp, int32_t* out_tokens,
                    float* out_logits, float* min_gap) {
  int token = 1, pos = 0;  // mod.rs:182-183 (BOS)
  float gap = INFINITY;
  const int V = m->cfg.vocab_size;
  auto t0 = std::chrono::steady_clock::now();
  while (pos < steps) {
    ref_forward(m, s, token, pos);
    if (out_logits) memcpy(out_logits + (size_t)pos * V, s->logits.data(), sizeof(float) * V);
    if (min_gap) {
      float a = -INFINITY, b = -INFINITY;
      for (int i = 0; i < V; ++i) {
        float z = s->logits[i];
        if (z > a) { b = a; a = z; } else if (z > b) { b = z; }
      }
      if (pos >= n_prompt && a - b < gap) gap = a - b;
    }
    int next;
    if (pos < n_prompt) next = prompt[pos];
    else next = ref_sample(s->logits.data(), V, temperature, topp);
    if (out_tokens) out_tokens[pos] = next;
    token = next;
    pos += 1;
  }
  double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (min_gap) *min_gap = gap;
  return el;
}

// ---------------------------------------------------------------------------------
// Tokenizer — engine/src/tokenizer/bpe.rs
// ---------------------------------------------------------------------------------
struct RefTok {
  std::vector<std::string> vocab;
  std::vector<float> scores;
  std::unordered_map<std::string, int> map;
  uint32_t max_token_length = 0;
};

// bpe.rs:19-45
RefTok* ref_tok_load(const char* path, int vocab_size) {
  FILE* f = fopen(path, "rb");
  if (!f) return nullptr;
  RefTok* t = new RefTok();
  if (fread(&t->max_token_length, 4, 1, f) != 1) { fclose(f); delete t; return nullptr; }
  for (int i = 0; i < vocab_size; ++i) {
    float sc; int32_t len;
    if (fread(&sc, 4, 1, f) != 1 || fread(&len, 4, 1, f) != 1) { fclose(f); delete t; return nullptr; }
    std::string s(len, '\0');
    if (len && fread(&s[0], 1, len, f) != (size_t)len) { fclose(f); delete t; return nullptr; }
    t->scores.push_back(sc);
    t->vocab.push_back(s);
    t->map[s] = i;  // HashMap::insert: later duplicates overwrite
  }
  fclose(f);
  return t;
}
void ref_tok_free(RefTok* t) { delete t; }
int ref_tok_max_len(const RefTok* t) { return (int)t->max_token_length; }

static size_t utf8_len(unsigned char c) {
  if (c < 0x80) return 1;
  if ((c >> 5) == 6) return 2;
  if ((c >> 4) == 14) return 3;
  if ((c >> 3) == 30) return 4;
  return 1;
}

// bpe.rs:50-96.  Returns token count, -1 where the reference panics (char not in vocab,
// bpe.rs:55) and -2 for the whitespace-only underflow (bpe.rs:66).
int ref_tok_encode(const RefTok* t, const char* s, int32_t* out, int cap) {
  std::string str(s);
  // str::trim(): strip leading/trailing chars with the Unicode White_Space property (char::is_whitespace)
  size_t b = 0, e = str.size();
  auto cp_at = [&](size_t i, size_t n) -> unsigned {
    const unsigned char* q = (const unsigned char*)str.data() + i;
    if (n == 1) return q[0];
    if (n == 2) return ((q[0] & 0x1Fu) << 6) | (q[1] & 0x3Fu);
    if (n == 3) return ((q[0] & 0x0Fu) << 12) | ((q[1] & 0x3Fu) << 6) | (q[2] & 0x3Fu);
    return ((q[0] & 0x07u) << 18) | ((q[1] & 0x3Fu) << 12) | ((q[2] & 0x3Fu) << 6) | (q[3] & 0x3Fu);
  };
  auto ws = [](unsigned c) {
    return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
           c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
  };
  for (;;) {  // leading
    if (b >= e) break;
    size_t n = utf8_len((unsigned char)str[b]);
    if (b + n > e || !ws(cp_at(b, n))) break;
    b += n;
  }
  for (;;) {  // trailing: step back to the start byte of the last char
    if (e <= b) break;
    size_t i = e - 1;
    while (i > b && ((unsigned char)str[i] & 0xC0) == 0x80) --i;
    if (!ws(cp_at(i, e - i))) break;
    e = i;
  }
  std::vector<int> tokens;
  for (size_t i = b; i < e;) {
    size_t n = utf8_len((unsigned char)str[i]);
    std::string ch = str.substr(i, n);
    i += n;
    if (ch == "\n") continue;
    auto it = t->map.find(ch);
    if (it == t->map.end()) return -1;
    tokens.push_back(it->second);
  }
  if (tokens.empty()) return -2;
  std::string buf;
  for (;;) {
    float best_score = -1e10f;
    int best_id = -1, best_idx = -1;
    for (size_t idx = 0; idx + 1 < tokens.size(); ++idx) {
      buf = t->vocab[tokens[idx]];
      buf += t->vocab[tokens[idx + 1]];
      auto it = t->map.find(buf);
      if (it != t->map.end() && t->scores[it->second] > best_score) {
        best_score = t->scores[it->second];
        best_id = it->second;
        best_idx = (int)idx;
      }
    }
    if (best_idx < 0) break;
    tokens[best_idx] = best_id;
    tokens.erase(tokens.begin() + best_idx + 1);
  }
  int n = (int)tokens.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = tokens[i];
  return n;
}

// bpe.rs:99-116 applied to vocab[id]; writes UTF-8 (char::from(byte) is Latin-1 → 1-2 bytes)
int ref_tok_decode(const RefTok* t, int id, char* out, int cap) {
  const std::string& s = t->vocab[id];
  std::string r;
  if (s.find("<s>") != std::string::npos) {
    r = "";
  } else if (!s.empty() && s.front() == '<' && s.back() == '>') {
    if (s.size() < 5) return -2;  // &str[3..5] is out of bounds: the reference panics (e.g. "</s>")
    char* endp = nullptr;
    const std::string hex = s.substr(3, 2);
    unsigned c = (unsigned)strtoul(hex.c_str(), &endp, 16);
    if (endp != hex.c_str() + 2) return -2;  // u8::from_str_radix(..).unwrap() panics (e.g. "<unk>")
    if (c < 0x80) r.push_back((char)c);
    else { r.push_back((char)(0xC0 | (c >> 6))); r.push_back((char)(0x80 | (c & 0x3F))); }
  } else {
    r = s;
  }
  int n = (int)r.size();
  if (n + 1 > cap) return -1;
  memcpy(out, r.data(), n);
  out[n] = 0;
  return n;
}

// ---------------------------------------------------------------------------------
// Synthetic-checkpoint generator (bench/test data; NOT part of the reference).
// Same integer recipe as rama_b200/checkpoint.py and rama_b200/csrc/synth.cu so the three
// produce bit-identical tensors: value = float(2*S - 8*65535) * scale + offset (unfused), S = sum of the eight
// 16-bit fields of two splitmix64 hashes of (key + 2*idx), (key + 2*idx + 1).
// ---------------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline uint32_t sum16(uint64_t h) {
  return (uint32_t)(h & 0xFFFF) + (uint32_t)((h >> 16) & 0xFFFF) + (uint32_t)((h >> 32) & 0xFFFF) +
         (uint32_t)(h >> 48);
}
void ref_synth_fill(float* dst, int64_t n, uint64_t seed, uint64_t tensor_id, float scale,
                    float offset) {
  const uint64_t key = splitmix64(seed ^ (tensor_id * 0xD1B54A32D192ED03ULL));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    uint32_t S = sum16(splitmix64(key + 2 * (uint64_t)i)) + sum16(splitmix64(key + 2 * (uint64_t)i + 1));
    float v = (float)(2 * (int32_t)S - 8 * 65535) * scale;
    dst[i] = v + offset;  // separate mul and add (-ffp-contract=off), as numpy does
  }
}

}  // extern "C"
