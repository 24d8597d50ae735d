// rama_b200.hpp — C++ host-side mirror of the reference engine's interface for the decode path, over the C ABI.
//
// The reference is Rust; there is no cargo/rustc in this image, so the host side above `rama_b200.h` is written in C++
// with the reference's own names, argument meaning and error behaviour (it panics — here: throws `rama::Panic`):
//
//   reference (paths relative to the reference repo root)                      here
//   engine/src/transformer/mod.rs:16-126   Storage / View / MutView / range    rama::Storage, View<T>, MutView<T>
//   engine/src/transformer/mod.rs:128-167  Config::from_file                   rama::Config
//   engine/src/transformer/state.rs:4-52   RunState / RunStateView             rama::RunState<T>, RunStateView<T>
//   engine/src/transformer/state.rs:54-122 TransformerWeights(+View), wcls alias  rama::TransformerWeights<T>, ...View<T>
//   engine/src/transformer/ram.rs          from_config / from_file             RunState<HostVec>::from_config, TransformerWeights<HostVec>::from_file
//   engine/src/transformer/hbm.rs          from_state / from_weight / from_gpu_ws  same names on DevBuf
//   engine/src/device/device.rs:3-24       trait Device<T> (11 methods)        rama::Device<T>
//   engine/src/device/gpu.rs:213-234       GPU::new()                          rama::GPU
//   engine/src/transformer/infer.rs:8-53   forward()                           rama::forward()
//   engine/src/tokenizer/bpe.rs            Tokenizer::new / encode, decode     rama::Tokenizer, rama::decode
//   engine/src/transformer/mod.rs:169-206  generate()                          rama::generate()
//
// `forward()` keeps its signature and its per-op body (infer.rs line by line, including the duplicated wq matmul of
// infer.rs:20-21); it first offers the step to `Device::forward_fused` — a defaulted trait method that only `GPU` overrides
// with the CUDA-graph-replayed `rama_forward` — so engine/src/main.rs reads the same with either device.
// The same shim in Rust is in integration/rust/ (INTEGRATION.md).  Header-only; link with librama_b200.so / .a.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "rama_b200.h"

namespace rama {

// The reference `.unwrap()`s every fallible call: an error is a panic.  Here it is an exception carrying the library's message.
struct Panic : std::runtime_error {
  int code;
  Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void ck(int rc) {
  if (rc != RAMA_OK) throw Panic(rc, std::string("rama_b200: ") + rama_last_error());
}

using usize = std::size_t;
constexpr usize kEnd = (usize)-1;  // an unbounded range end (`a..`)

// ---- mod.rs:128-167 ------------------------------------------------------------------------------------------------
struct Config {
  usize dim = 0, hidden_dim = 0, n_layers = 0, n_heads = 0, n_kv_heads = 0, vocab_size = 0, seq_len = 0;
  bool shared_weight = false;
  // 7 little-endian i32; vocab > 0 ⇒ the classifier shares the embedding (mod.rs:147-155)
  static Config from_file(std::istream& f) {
    int32_t h[7];
    f.read(reinterpret_cast<char*>(h), sizeof(h));
    if (!f) throw Panic(RAMA_E_IO, "Config::from_file: short read");
    Config c;
    c.dim = (usize)h[0]; c.hidden_dim = (usize)h[1]; c.n_layers = (usize)h[2]; c.n_heads = (usize)h[3];
    c.n_kv_heads = (usize)h[4];
    c.shared_weight = h[5] > 0;
    c.vocab_size = (usize)(h[5] > 0 ? h[5] : -h[5]);
    c.seq_len = (usize)h[6];
    return c;
  }
  rama_config c() const {
    return rama_config{(int32_t)dim, (int32_t)hidden_dim, (int32_t)n_layers, (int32_t)n_heads, (int32_t)n_kv_heads,
                       (int32_t)vocab_size, (int32_t)seq_len, shared_weight ? 1 : 0};
  }
};

// ---- mod.rs:16-126: a view is (storage, ABSOLUTE element range); slice() is relative to the storage, not to the view ----
struct Range { usize start, end; };

struct HostVec {  // ≙ Vec<f32>
  std::vector<float> v;
  usize length() const { return v.size(); }
};

struct DevBuf {  // ≙ CudaSlice<f32> (hbm.rs:6-10); `session` only on RunState.x: the fused path's handle
  float* ptr = nullptr;
  usize len = 0;
  rama_ctx* ctx = nullptr;
  rama_session* session = nullptr;
  usize length() const { return len; }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      ptr = o.ptr; len = o.len; ctx = o.ctx; session = o.session;
      o.ptr = nullptr; o.session = nullptr; o.len = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }

 private:
  void release() {
    if (session) rama_session_destroy(session);
    if (ptr) rama_dev_free(ctx, ptr);
    session = nullptr; ptr = nullptr;
  }
};

inline Range range_from(usize start, usize end, usize max_len) { return Range{start, end == kEnd ? max_len : end}; }

template <class T>
struct View {
  const T* data;
  Range range;
  explicit View(const T& storage) : data(&storage), range{0, storage.length()} {}
  View(const T* d, Range r) : data(d), range(r) {}
  View slice(usize start, usize end = kEnd) const { return View(data, range_from(start, end, data->length())); }
};

template <class T>
struct MutView {
  T* data;
  Range range;
  explicit MutView(T& storage) : data(&storage), range{0, storage.length()} {}
  MutView(T* d, Range r) : data(d), range(r) {}
  View<T> as_view() const { return View<T>(data, range); }
  View<T> slice(usize start, usize end = kEnd) const { return View<T>(data, range_from(start, end, data->length())); }
  MutView mut_slice(usize start, usize end = kEnd) { return MutView(data, range_from(start, end, data->length())); }
};

// ---- state.rs:4-52 ---------------------------------------------------------------------------------------------------
template <class T>
struct RunState {
  T x, xb, xb2, hb, hb2, q, k, v, att, logits, key_cache, value_cache;
};

template <class T>
struct RunStateView {
  MutView<T> x, xb, xb2, hb, hb2, q, k, v, att, logits, key_cache, value_cache;
  static RunStateView from_rs(RunState<T>& rs) {
    return RunStateView{MutView<T>(rs.x), MutView<T>(rs.xb), MutView<T>(rs.xb2), MutView<T>(rs.hb), MutView<T>(rs.hb2),
                        MutView<T>(rs.q), MutView<T>(rs.k), MutView<T>(rs.v), MutView<T>(rs.att), MutView<T>(rs.logits),
                        MutView<T>(rs.key_cache), MutView<T>(rs.value_cache)};
  }
};

// ram.rs:6-23
inline RunState<HostVec> run_state_from_config(const Config& cfg) {
  const usize kv_dim = cfg.dim * cfg.n_kv_heads / cfg.n_heads;
  auto z = [](usize n) { return HostVec{std::vector<float>(n, 0.0f)}; };
  return RunState<HostVec>{z(cfg.dim), z(cfg.dim), z(cfg.dim), z(cfg.hidden_dim), z(cfg.hidden_dim), z(cfg.dim), z(cfg.dim),
                           z(cfg.dim), z(cfg.n_heads * cfg.seq_len), z(cfg.vocab_size), z(cfg.n_layers * cfg.seq_len * kv_dim),
                           z(cfg.n_layers * cfg.seq_len * kv_dim)};
}

// ---- state.rs:54-122 -------------------------------------------------------------------------------------------------
template <class T>
struct TransformerWeights {
  T token_embedding_table, rms_att_weight, rms_ffn_weight, wq, wk, wv, wo, w1, w2, w3, rms_final_weight, freq_cis_real,
      freq_cis_imag;
  bool wcls_exists = false;
  T wcls;
};

template <class T>
struct TransformerWeightsView {
  View<T> token_embedding_table, rms_att_weight, rms_ffn_weight, wq, wk, wv, wo, w1, w2, w3, rms_final_weight, freq_cis_real,
      freq_cis_imag;
  bool wcls_exists;
  View<T> wcls;
  // from_ws (state.rs:94-122) / from_gpu_ws (hbm.rs:93-121): wcls aliases the embedding when the file has no classifier
  static TransformerWeightsView from_ws(const TransformerWeights<T>& ws) {
    return TransformerWeightsView{View<T>(ws.token_embedding_table), View<T>(ws.rms_att_weight), View<T>(ws.rms_ffn_weight),
                                  View<T>(ws.wq), View<T>(ws.wk), View<T>(ws.wv), View<T>(ws.wo), View<T>(ws.w1), View<T>(ws.w2),
                                  View<T>(ws.w3), View<T>(ws.rms_final_weight), View<T>(ws.freq_cis_real),
                                  View<T>(ws.freq_cis_imag), ws.wcls_exists,
                                  ws.wcls_exists ? View<T>(ws.wcls) : View<T>(ws.token_embedding_table)};
  }
  static TransformerWeightsView from_gpu_ws(const TransformerWeights<T>& ws) { return from_ws(ws); }
};

// utils/read.rs + ram.rs:27-51: the llama2.c v0 tensor order
inline HostVec read_vec(std::istream& f, usize n) {
  HostVec h{std::vector<float>(n)};
  f.read(reinterpret_cast<char*>(h.v.data()), (std::streamsize)(n * sizeof(float)));
  if (!f) throw Panic(RAMA_E_IO, "read_vec: short read");
  return h;
}
inline TransformerWeights<HostVec> weights_from_file(std::istream& f, const Config& c) {
  const usize head_size = c.dim / c.n_heads;
  TransformerWeights<HostVec> w;
  w.token_embedding_table = read_vec(f, c.vocab_size * c.dim);
  w.rms_att_weight = read_vec(f, c.n_layers * c.dim);
  w.wq = read_vec(f, c.n_layers * c.dim * c.dim);
  w.wk = read_vec(f, c.n_layers * c.dim * c.dim);
  w.wv = read_vec(f, c.n_layers * c.dim * c.dim);
  w.wo = read_vec(f, c.n_layers * c.dim * c.dim);
  w.rms_ffn_weight = read_vec(f, c.n_layers * c.dim);
  w.w1 = read_vec(f, c.n_layers * c.dim * c.hidden_dim);
  w.w2 = read_vec(f, c.n_layers * c.dim * c.hidden_dim);
  w.w3 = read_vec(f, c.n_layers * c.dim * c.hidden_dim);
  w.rms_final_weight = read_vec(f, c.dim);
  w.freq_cis_real = read_vec(f, c.seq_len * head_size / 2);
  w.freq_cis_imag = read_vec(f, c.seq_len * head_size / 2);
  w.wcls_exists = !c.shared_weight;
  w.wcls = c.shared_weight ? HostVec{std::vector<float>{1.0f}} : read_vec(f, c.vocab_size * c.dim);
  return w;
}

// ---- device.rs:3-24 --------------------------------------------------------------------------------------------------
template <class T>
struct Device {
  virtual ~Device() = default;
  virtual void array_add(MutView<T>& target, const View<T>& source, usize n) const = 0;
  virtual void array_mult(MutView<T>& target, const View<T>& source, usize n) const = 0;
  virtual void sinu(MutView<T>& o, usize n) const = 0;
  virtual void multi_head_attention(RunStateView<T>& rsv, const Config& cfg, usize layer, usize pos) const = 0;
  virtual void copy_from_slice(MutView<T>& target, const View<T>& source, usize n) const = 0;
  virtual void rmsnorm(MutView<T>& o, const View<T>& x, const View<T>& weight, usize n) const = 0;
  virtual void apply_position(MutView<T>& q, MutView<T>& k, const View<T>& pos_real, const View<T>& pos_img,
                              usize head_size) const = 0;
  virtual void matmul(MutView<T>& o, const View<T>& a, const View<T>& b, usize width, usize o_rows, usize o_cols) const = 0;
  virtual void softmax(MutView<T>& x, usize n) const = 0;
  virtual usize sample(const Config& cfg, RunStateView<T>& rsv, float temperature, float topp) const = 0;
  virtual void to_cpu(const RunStateView<T>& state, RunState<HostVec>& cpu_state) const = 0;
  // new, defaulted: the whole step in one call when the device has a fused path for this RunState
  virtual bool forward_fused(RunStateView<T>&, usize /*token*/, usize /*pos*/) const { return false; }
};

// ---- gpu.rs:213-234: the device.  Send + Sync in the reference (lib.rs:56): op-level calls share the context's op stream
// and take a lock; sessions are independent. ----
class GPU final : public Device<DevBuf> {
 public:
  rama_ctx* ctx = nullptr;
  // ≙ GPU::new(): no NVRTC / cuBLAS.  RAMA_GPUS=N (or GPU(N)) makes the one handle a tensor-parallel context over devices
  // 0..N-1 of this process (rama_ctx_create_multi) — callers stay unchanged, exactly as with one device.
  GPU() : GPU(gpus_from_env()) {}
  explicit GPU(int n_gpus) {
    if (n_gpus > 1) ck(rama_ctx_create_multi(n_gpus, nullptr, &ctx));
    else ck(rama_ctx_create(0, nullptr, &ctx));
  }
  static int gpus_from_env() {
    const char* v = std::getenv("RAMA_GPUS");
    const int n = v && *v ? std::atoi(v) : 1;
    return n < 1 ? 1 : n;
  }
  ~GPU() override { if (ctx) rama_ctx_destroy(ctx); }
  GPU(const GPU&) = delete;
  GPU& operator=(const GPU&) = delete;

  DevBuf alloc(const std::vector<float>& host) const {  // hbm.rs:14-16 `allocate`
    DevBuf b;
    b.ctx = ctx; b.len = host.size();
    ck(rama_dev_alloc(ctx, host.size(), &b.ptr));
    ck(rama_dev_h2d(ctx, b.ptr, host.data(), host.size()));
    return b;
  }

  void array_add(MutView<DevBuf>& t, const View<DevBuf>& s, usize n) const override { L g(mu_); ck(rama_op_array_add(ctx, pm(t, n), p(s, n), n)); }
  void array_mult(MutView<DevBuf>& t, const View<DevBuf>& s, usize n) const override { L g(mu_); ck(rama_op_array_mult(ctx, pm(t, n), p(s, n), n)); }
  void sinu(MutView<DevBuf>& o, usize n) const override { L g(mu_); ck(rama_op_sinu(ctx, pm(o, n), n)); }
  void multi_head_attention(RunStateView<DevBuf>& rsv, const Config& cfg, usize layer, usize pos) const override {
    L g(mu_);
    const rama_config c = cfg.c();
    const usize row_end = (layer * cfg.seq_len + pos + 1) * cfg.dim;  // last cache row the op reads (cpu.rs:32-36)
    ck(rama_op_multi_head_attention(ctx, pm(rsv.xb, cfg.dim), pm(rsv.att, cfg.n_heads * cfg.seq_len), pm(rsv.q, cfg.dim),
                                    pm(rsv.key_cache, row_end), pm(rsv.value_cache, row_end), &c,
                                    (int32_t)layer, (int32_t)pos));
  }
  void copy_from_slice(MutView<DevBuf>& t, const View<DevBuf>& s, usize n) const override { L g(mu_); ck(rama_op_copy_from_slice(ctx, pm(t, n), p(s, n), n)); }
  void rmsnorm(MutView<DevBuf>& o, const View<DevBuf>& x, const View<DevBuf>& w, usize n) const override { L g(mu_); ck(rama_op_rmsnorm(ctx, pm(o, n), p(x, n), p(w, n), n)); }
  void apply_position(MutView<DevBuf>& q, MutView<DevBuf>& k, const View<DevBuf>& pr, const View<DevBuf>& pi,
                      usize head_size) const override { L g(mu_); ck(rama_op_apply_position(ctx, pm(q, head_size), pm(k, head_size), p(pr, head_size / 2), p(pi, head_size / 2), head_size)); }
  void matmul(MutView<DevBuf>& o, const View<DevBuf>& a, const View<DevBuf>& b, usize width, usize o_rows,
              usize o_cols) const override { L g(mu_); ck(rama_op_matmul(ctx, pm(o, o_rows * o_cols), p(a, o_rows * width), p(b, width * o_cols), width, o_rows, o_cols)); }
  void softmax(MutView<DevBuf>& x, usize n) const override { L g(mu_); ck(rama_op_softmax(ctx, pm(x, n), n)); }

  // cpu.rs:155-179 + infer.rs:55-85 on the device; with a session only the token id crosses PCIe
  usize sample(const Config& cfg, RunStateView<DevBuf>& rsv, float temperature, float topp) const override {
    int32_t next = 0;
    if (rsv.x.data->session) {
      ck(rama_sample(rsv.x.data->session, temperature, topp, &next));
    } else {
      L g(mu_);
      ck(rama_op_sample(ctx, pm(rsv.logits), cfg.vocab_size, temperature, topp, &next));
    }
    return (usize)next;
  }
  // gpu.rs:196-209
  void to_cpu(const RunStateView<DevBuf>& st, RunState<HostVec>& cpu) const override {
    const MutView<DevBuf>* d[RAMA_S_COUNT] = {&st.x, &st.xb, &st.xb2, &st.hb, &st.hb2, &st.q, &st.k, &st.v, &st.att, &st.logits,
                                              &st.key_cache, &st.value_cache};
    HostVec* h[RAMA_S_COUNT] = {&cpu.x, &cpu.xb, &cpu.xb2, &cpu.hb, &cpu.hb2, &cpu.q, &cpu.k, &cpu.v, &cpu.att, &cpu.logits,
                                &cpu.key_cache, &cpu.value_cache};
    for (int i = 0; i < RAMA_S_COUNT; ++i) {
      if (st.x.data->session) {
        usize n = 0;
        ck(rama_state_to_host(st.x.data->session, i, h[i]->v.data(), h[i]->v.size(), &n));
      } else {
        L g(mu_);
        ck(rama_dev_d2h(ctx, h[i]->v.data(), d[i]->data->ptr + d[i]->range.start, h[i]->v.size()));
      }
    }
  }
  bool forward_fused(RunStateView<DevBuf>& rsv, usize token, usize pos) const override {
    if (!rsv.x.data->session) return false;
    ck(rama_forward(rsv.x.data->session, (int32_t)token, (int32_t)pos));
    return true;
  }

 private:
  using L = std::lock_guard<std::mutex>;
  mutable std::mutex mu_;
  // a view that does not hold n elements from its start is the reference's slice panic (cpu.rs indexes `&data[range]`) — and
  // here it also keeps a bad position or token from touching device memory outside the buffer
  static const float* p(const View<DevBuf>& v, usize n = 0) {
    if (v.range.start + n > v.data->len) throw Panic(RAMA_E_STATE, "range end index out of range for slice (View)");
    return v.data->ptr + v.range.start;
  }
  static float* pm(const MutView<DevBuf>& v, usize n = 0) {
    if (v.range.start + n > v.data->len) throw Panic(RAMA_E_STATE, "range end index out of range for slice (MutView)");
    return v.data->ptr + v.range.start;
  }
};

// ---- hbm.rs:19-34 / 55-90 ----------------------------------------------------------------------------------------------
// from_state: device buffers for the op-level path; with `fused` (default) RunState.x also owns a session — the KV cache
// is allocated and zeroed in HBM by the library, so the op-level caches shrink to one element (the reference uploads 2 GiB
// of zeros per session at 7B).
inline RunState<DevBuf> run_state_from_state(RunState<HostVec>& state, const GPU& device, bool fused = true) {
  RunState<DevBuf> r;
  r.x = device.alloc(state.x.v);
  if (fused) ck(rama_session_create(device.ctx, &r.x.session));
  r.xb = device.alloc(state.xb.v); r.xb2 = device.alloc(state.xb2.v); r.hb = device.alloc(state.hb.v);
  r.hb2 = device.alloc(state.hb2.v); r.q = device.alloc(state.q.v); r.k = device.alloc(state.k.v);
  r.v = device.alloc(state.v.v); r.att = device.alloc(state.att.v); r.logits = device.alloc(state.logits.v);
  const std::vector<float> one(1, 0.0f);
  r.key_cache = device.alloc(fused ? one : state.key_cache.v);
  r.value_cache = device.alloc(fused ? one : state.value_cache.v);
  return r;
}
// from_weight: one pass host → HBM inside the library (what the fused step reads); `per_op_views` additionally uploads the
// 14 tensors as separate device buffers so the reference-shaped forward() over the 11 ops can run on them (small models).
inline TransformerWeights<DevBuf> weights_from_weight(TransformerWeights<HostVec>& tw, const Config& cfg, const GPU& device,
                                                      bool per_op_views = false) {
  const rama_config c = cfg.c();
  const float* t[RAMA_T_COUNT] = {tw.token_embedding_table.v.data(), tw.rms_att_weight.v.data(), tw.wq.v.data(), tw.wk.v.data(),
                                  tw.wv.v.data(), tw.wo.v.data(), tw.rms_ffn_weight.v.data(), tw.w1.v.data(), tw.w2.v.data(),
                                  tw.w3.v.data(), tw.rms_final_weight.v.data(), tw.freq_cis_real.v.data(),
                                  tw.freq_cis_imag.v.data(), tw.wcls_exists ? tw.wcls.v.data() : nullptr};
  ck(rama_ctx_load_host(device.ctx, &c, t));
  const std::vector<float> one(1, 0.0f);
  auto up = [&](const HostVec& h) { return device.alloc(per_op_views ? h.v : one); };
  TransformerWeights<DevBuf> w;
  w.token_embedding_table = up(tw.token_embedding_table); w.rms_att_weight = up(tw.rms_att_weight);
  w.rms_ffn_weight = up(tw.rms_ffn_weight); w.wq = up(tw.wq); w.wk = up(tw.wk); w.wv = up(tw.wv); w.wo = up(tw.wo);
  w.w1 = up(tw.w1); w.w2 = up(tw.w2); w.w3 = up(tw.w3); w.rms_final_weight = up(tw.rms_final_weight);
  w.freq_cis_real = up(tw.freq_cis_real); w.freq_cis_imag = up(tw.freq_cis_imag);
  w.wcls_exists = tw.wcls_exists; w.wcls = up(tw.wcls);
  return w;
}

// The fast loader (BASELINE north star (2)): the library reads the v0 file itself — pread → pinned ring → DMA, sharded at
// upload — instead of per-f32 `read_exact` into host vectors (utils/read.rs:25-33).  The 14 DevBufs are one-element
// placeholders (the fused step reads the library's copy); wcls_exists comes from the header.
inline TransformerWeights<DevBuf> weights_from_path(const std::string& path, const Config& cfg, const GPU& device) {
  ck(rama_ctx_load_file(device.ctx, path.c_str()));
  const std::vector<float> one(1, 0.0f);
  TransformerWeights<DevBuf> w;
  w.token_embedding_table = device.alloc(one); w.rms_att_weight = device.alloc(one); w.rms_ffn_weight = device.alloc(one);
  w.wq = device.alloc(one); w.wk = device.alloc(one); w.wv = device.alloc(one); w.wo = device.alloc(one);
  w.w1 = device.alloc(one); w.w2 = device.alloc(one); w.w3 = device.alloc(one); w.rms_final_weight = device.alloc(one);
  w.freq_cis_real = device.alloc(one); w.freq_cis_imag = device.alloc(one);
  w.wcls_exists = !cfg.shared_weight; w.wcls = device.alloc(one);
  return w;
}

// ---- infer.rs:8-53 ---------------------------------------------------------------------------------------------------
template <class T, class D>
void forward(const Config& cfg, const TransformerWeightsView<T>& wv, RunStateView<T>& rsv, usize token, usize pos,
             const D& device) {
  if (device.forward_fused(rsv, token, pos)) return;
  const usize dim = cfg.dim, hidden_dim = cfg.hidden_dim, head_size = dim / cfg.n_heads;
  device.copy_from_slice(rsv.x, wv.token_embedding_table.slice(token * dim, (token + 1) * cfg.dim), dim);

  const View<T> pos_real = wv.freq_cis_real.slice(pos * (head_size / 2));
  const View<T> pos_img = wv.freq_cis_imag.slice(pos * (head_size / 2));

  for (usize layer = 0; layer < cfg.n_layers; ++layer) {
    device.rmsnorm(rsv.xb, rsv.x.as_view(), wv.rms_att_weight.slice(layer * dim), dim);
    device.matmul(rsv.q, wv.wq.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1);
    device.matmul(rsv.q, wv.wq.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1);  // (sic) infer.rs:21
    device.matmul(rsv.k, wv.wk.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1);
    device.matmul(rsv.v, wv.wv.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1);

    for (usize h = 0; h < cfg.n_heads; ++h) {
      MutView<T> q = rsv.q.mut_slice(h * head_size);
      MutView<T> k = rsv.k.mut_slice(h * head_size);
      device.apply_position(q, k, pos_real, pos_img, head_size);
    }

    const usize lo = layer * cfg.seq_len * dim;
    MutView<T> kc = rsv.key_cache.mut_slice(lo + pos * dim, lo + (pos + 1) * dim);
    MutView<T> vc = rsv.value_cache.mut_slice(lo + pos * dim, lo + (pos + 1) * dim);
    device.copy_from_slice(kc, rsv.k.as_view(), dim);
    device.copy_from_slice(vc, rsv.v.as_view(), dim);
    device.multi_head_attention(rsv, cfg, layer, pos);
    device.matmul(rsv.xb2, wv.wo.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1);

    device.array_add(rsv.x, rsv.xb2.as_view(), dim);

    device.rmsnorm(rsv.xb, rsv.x.as_view(), wv.rms_ffn_weight.slice(layer * dim), dim);

    device.matmul(rsv.hb, wv.w1.slice(layer * hidden_dim * dim), rsv.xb.as_view(), dim, hidden_dim, 1);
    device.matmul(rsv.hb2, wv.w3.slice(layer * hidden_dim * dim), rsv.xb.as_view(), dim, hidden_dim, 1);

    device.sinu(rsv.hb, hidden_dim);
    device.array_mult(rsv.hb, rsv.hb2.as_view(), hidden_dim);
    device.matmul(rsv.xb, wv.w2.slice(layer * dim * hidden_dim), rsv.hb.as_view(), hidden_dim, dim, 1);
    device.array_add(rsv.x, rsv.xb.as_view(), dim);
  }
  device.copy_from_slice(rsv.xb, rsv.x.as_view(), dim);
  device.rmsnorm(rsv.x, rsv.xb.as_view(), wv.rms_final_weight, dim);
  device.matmul(rsv.logits, wv.wcls, rsv.x.as_view(), dim, cfg.vocab_size, 1);
}

// ---- tokenizer/bpe.rs --------------------------------------------------------------------------------------------------
class Tokenizer {
 public:
  Tokenizer(const std::string& path, usize vocab_size) { ck(rama_tokenizer_load(path.c_str(), (int32_t)vocab_size, &t_)); }
  ~Tokenizer() { if (t_) rama_tokenizer_free(t_); }
  Tokenizer(const Tokenizer&) = delete;
  Tokenizer& operator=(const Tokenizer&) = delete;
  std::vector<usize> encode(const std::string& text) const {  // bpe.rs:50-97 (panics on what the reference panics on)
    int32_t n = 0;
    ck(rama_tokenizer_encode(t_, text.c_str(), nullptr, 0, &n));
    std::vector<int32_t> ids((usize)n);
    ck(rama_tokenizer_encode(t_, text.c_str(), ids.data(), n, &n));
    return std::vector<usize>(ids.begin(), ids.end());
  }
  std::string decode(usize token) const {  // decode(vocab[token]) bpe.rs:102-116
    char buf[512];
    int32_t n = 0;
    ck(rama_tokenizer_decode(t_, (int32_t)token, buf, (int32_t)sizeof(buf), &n));
    return std::string(buf, (usize)n);
  }

 private:
  rama_tokenizer* t_ = nullptr;
};

// ---- mod.rs:169-206 ----------------------------------------------------------------------------------------------------
template <class T, class D>
std::string generate(const Config& cfg, const Tokenizer& tokenizer, const std::string& prompt, float temperature, usize steps,
                     float topp, const TransformerWeightsView<T>& wv, RunStateView<T>& rsv, const D& device,
                     std::ostream* out = &std::cout, std::vector<usize>* tokens_out = nullptr) {
  const std::vector<usize> prompt_tokens = prompt.size() > 0 ? tokenizer.encode(prompt) : std::vector<usize>();
  usize token = 1, pos = 0, next;
  std::string response;
  while (pos < steps) {
    forward(cfg, wv, rsv, token, pos, device);
    if (pos < prompt_tokens.size()) next = prompt_tokens[pos];
    else next = device.sample(cfg, rsv, temperature, topp);
    const std::string token_str = tokenizer.decode(next);
    response += token_str;
    if (out) { (*out) << token_str; out->flush(); }
    if (tokens_out) tokens_out->push_back(next);
    token = next;
    pos += 1;
  }
  return response;
}

}  // namespace rama
