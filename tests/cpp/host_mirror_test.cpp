// host_mirror_test.cpp — drives the C++ mirror of the reference interface (include/rama_b200.hpp) the way the reference's own
// main.rs / lib.rs drive the engine, and dumps what a parity test needs (tests/test_host_cpp.py compares it with the oracle
// and the golden logits of the reference's torch model).
//
//   host_mirror_test <model.bin> <tokens.i32> <out.f32> <mode: fused|per-op>
//
// tokens.i32: n little-endian int32 token ids (teacher forcing).  out.f32: n × vocab logits after forward(token_i, i), then
// n int32-as-f32 greedy samples (Device::sample at temperature 0), then the key cache of layer 0 after the last step
// (seq_len × dim) read back through Device::to_cpu.
#include <cstdio>
#include <fstream>
#include <vector>

#include "rama_b200.hpp"

using namespace rama;

int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: host_mirror_test model.bin tokens.i32 out.f32 fused|per-op\n"); return 2; }
  const std::string model = argv[1], tok_path = argv[2], out_path = argv[3], mode = argv[4];
  const bool per_op = mode == "per-op";
  try {
    std::ifstream rd(model, std::ios::binary);
    if (!rd) throw Panic(RAMA_E_IO, "cannot open " + model);
    const Config config = Config::from_file(rd);
    const GPU device;
    TransformerWeights<HostVec> host_weights = weights_from_file(rd, config);
    TransformerWeights<DevBuf> weights = weights_from_weight(host_weights, config, device, per_op);
    RunState<HostVec> cpu_state = run_state_from_config(config);
    RunState<DevBuf> state = run_state_from_state(cpu_state, device, !per_op);
    const TransformerWeightsView<DevBuf> wv = TransformerWeightsView<DevBuf>::from_gpu_ws(weights);
    RunStateView<DevBuf> rsv = RunStateView<DevBuf>::from_rs(state);

    std::ifstream tf(tok_path, std::ios::binary);
    std::vector<int32_t> tokens;
    for (int32_t t; tf.read(reinterpret_cast<char*>(&t), 4);) tokens.push_back(t);

    std::ofstream out(out_path, std::ios::binary);
    std::vector<float> samples;
    for (usize pos = 0; pos < tokens.size(); ++pos) {
      forward(config, wv, rsv, (usize)tokens[pos], pos, device);
      device.to_cpu(rsv, cpu_state);  // ≙ Device::to_cpu: all 12 RunState buffers
      out.write(reinterpret_cast<const char*>(cpu_state.logits.v.data()), (std::streamsize)(config.vocab_size * sizeof(float)));
      samples.push_back((float)device.sample(config, rsv, 0.0f, 0.9f));
    }
    out.write(reinterpret_cast<const char*>(samples.data()), (std::streamsize)(samples.size() * sizeof(float)));
    out.write(reinterpret_cast<const char*>(cpu_state.key_cache.v.data()), (std::streamsize)(config.seq_len * config.dim * sizeof(float)));
    // the reference panics on a position past the window (infer.rs:32 slice): so must the mirror
    bool panicked = false;
    try { forward(config, wv, rsv, 1, config.seq_len, device); } catch (const Panic&) { panicked = true; }
    if (!panicked) { std::fprintf(stderr, "forward past seq_len did not panic\n"); return 3; }
    panicked = false;  // ... and a token outside the vocabulary (infer.rs:13 slice)
    try { forward(config, wv, rsv, config.vocab_size, 0, device); } catch (const Panic&) { panicked = true; }
    if (!panicked) { std::fprintf(stderr, "forward with token == vocab_size did not panic\n"); return 3; }
    std::printf("host-mirror %s ok: %zu steps\n", mode.c_str(), tokens.size());
  } catch (const Panic& e) {
    std::fprintf(stderr, "panicked: %s\n", e.what());
    return 101;
  }
  return 0;
}
