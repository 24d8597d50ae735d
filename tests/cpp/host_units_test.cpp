// host_units_test.cpp — the C++ host mirror's pure host logic (no CUDA call): Config::from_file, the absolute-range View /
// MutView semantics of mod.rs:16-126, TransformerWeights::from_file sizes and the wcls alias (state.rs:111-117), the bounded
// channel and get_batch (batcher.rs:8-38).  Run by tests/test_host_cpp.py on the CPU.
//
//   host_units_test <shared.bin> <untied.bin>
#include <chrono>
#include <cstdio>
#include <thread>

#include "rama_b200_service.hpp"

using namespace rama;

static int fails = 0;
#define EXPECT(cond)                                                            \
  do {                                                                          \
    if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } \
  } while (0)

static void views() {
  HostVec st{std::vector<float>(100, 0.f)};
  for (int i = 0; i < 100; ++i) st.v[(usize)i] = (float)i;
  View<HostVec> v(st);
  EXPECT(v.range.start == 0 && v.range.end == 100);
  const View<HostVec> a = v.slice(10, 20);
  EXPECT(a.range.start == 10 && a.range.end == 20);
  // slice() of a view is relative to the STORAGE, not to the view (mod.rs:44-51): a.slice(5..) starts at 5, not at 15
  const View<HostVec> b = a.slice(5);
  EXPECT(b.range.start == 5 && b.range.end == 100);
  MutView<HostVec> m(st);
  MutView<HostVec> m2 = m.mut_slice(30, 40);
  EXPECT(m2.range.start == 30 && m2.range.end == 40);
  EXPECT(m2.as_view().range.start == 30 && m2.as_view().range.end == 40);
  EXPECT(m2.slice(7).range.start == 7 && m2.slice(7).range.end == 100);
  EXPECT(m2.mut_slice(1, 2).range.start == 1);
  const Range r = range_from(3, kEnd, 9);
  EXPECT(r.start == 3 && r.end == 9);
}

static void files(const char* shared, const char* untied) {
  for (int which = 0; which < 2; ++which) {
    std::ifstream rd(which == 0 ? shared : untied, std::ios::binary);
    EXPECT((bool)rd);
    const Config c = Config::from_file(rd);
    EXPECT(c.shared_weight == (which == 0));  // vocab > 0 in the header ⇔ shared classifier (mod.rs:147-155)
    EXPECT(c.dim > 0 && c.n_heads > 0 && c.dim % c.n_heads == 0 && c.n_kv_heads == c.n_heads);
    const TransformerWeights<HostVec> w = weights_from_file(rd, c);
    EXPECT(rd.peek() == std::ifstream::traits_type::eof());  // the tensor list consumes the file exactly
    EXPECT(w.token_embedding_table.length() == c.vocab_size * c.dim);
    EXPECT(w.wq.length() == c.n_layers * c.dim * c.dim && w.w1.length() == c.n_layers * c.dim * c.hidden_dim);
    EXPECT(w.freq_cis_real.length() == c.seq_len * (c.dim / c.n_heads) / 2);
    EXPECT(w.wcls_exists == !c.shared_weight);
    EXPECT(w.wcls.length() == (c.shared_weight ? 1 : c.vocab_size * c.dim));  // ram.rs:44-48: vec![1.0] placeholder
    const TransformerWeightsView<HostVec> wv = TransformerWeightsView<HostVec>::from_ws(w);
    EXPECT(wv.wcls.data == (c.shared_weight ? &w.token_embedding_table : &w.wcls));  // state.rs:111-117
    const RunState<HostVec> rs = run_state_from_config(c);
    EXPECT(rs.att.length() == c.n_heads * c.seq_len && rs.key_cache.length() == c.n_layers * c.seq_len * c.dim);
    const rama_config cc = c.c();
    EXPECT(cc.dim == (int32_t)c.dim && cc.shared_weight == (c.shared_weight ? 1 : 0));
  }
  bool panicked = false;
  try { std::ifstream none("/nonexistent"); Config::from_file(none); } catch (const Panic&) { panicked = true; }
  EXPECT(panicked);
}

static void channel_and_batcher() {
  using namespace std::chrono;
  Channel<int> ch(3);
  for (int i = 0; i < 3; ++i) ch.send(i);
  std::thread late([&] { std::this_thread::sleep_for(milliseconds(30)); ch.send(3); ch.send(4); });  // first send blocks until a recv
  std::vector<int> got;
  get_batch(ch, got, 2, duration<double>(1.0));        // batcher.rs:33: stops at batch_size
  EXPECT(got.size() == 2 && got[0] == 0 && got[1] == 1);
  const auto t0 = steady_clock::now();
  get_batch(ch, got, 8, duration<double>(0.25));       // batcher.rs:14: the timeout ends the wait with what arrived
  const double waited = duration<double>(steady_clock::now() - t0).count();
  late.join();
  EXPECT(got.size() == 5 && got[4] == 4);
  EXPECT(waited >= 0.2 && waited < 1.0);
  get_batch(ch, got, 8, duration<double>(0.0));        // nothing queued, zero wait: returns at once
  EXPECT(got.size() == 5);
  ch.close();
  int x = -1;
  EXPECT(!ch.recv(x, steady_clock::now() + milliseconds(50)));
  const EngineConfig ec = EngineConfig::from_model_tokenizer("m.bin", "t.bin");  // lib.rs:36-45 defaults
  EXPECT(ec.step == 255 && ec.temperature == 1.0f && ec.topp == 0.9f && ec.mode == "generate");
  bool panicked = false;
  try { EngineService::global(); } catch (const Panic&) { panicked = true; }  // "Engine not initialized" (lib.rs:84-86)
  EXPECT(panicked);
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: host_units_test shared.bin untied.bin\n"); return 2; }
  views();
  files(argv[1], argv[2]);
  channel_and_batcher();
  std::printf(fails ? "host units: %d FAILED\n" : "host units ok\n", fails);
  return fails ? 1 : 0;
}
