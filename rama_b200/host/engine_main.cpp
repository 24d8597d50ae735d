// engine_main.cpp — the engine CLI (≙ engine/src/main.rs with `--features gpu`) on the C++ mirror of the reference
// interface (include/rama_b200.hpp): same flags, same call sequence, same output.
//
//   engine -m <model.bin> -t <tokenizer.bin> [-p prompt] [-s steps=255] [-r temperature=1.0] [-l topp=0.9] [--per-op] [--gpus N]
//
// main.rs:61-105: Config::from_file → GPU::new() → TransformerWeights::from_file/from_weight → RunState::from_config/
// from_state → views → Tokenizer::new → generate(...) → "elapsed: S.mmm s, avg tok/s: (step-1)/elapsed".
// Default: the library's file loader + the fused CUDA-graph step; `--per-op` runs the reference-shaped forward() over the
// 11 Device ops on per-tensor device buffers (small models: every tensor is uploaded a second time).
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "rama_b200.hpp"

using namespace rama;

struct Args {  // main.rs:20-50
  std::string model, tokenizer, prompt, mode = "generate";
  unsigned step = 255;
  float temperature = 1.0f, topp = 0.9f;
  bool per_op = false;
  int gpus = 0;  // 0: RAMA_GPUS or 1; N > 1: one tensor-parallel handle over devices 0..N-1 of this process
};

static void usage() {
  std::fprintf(stderr,
               "Usage: engine -m <MODEL> -t <TOKENIZER> [-p <PROMPT>] [-s <STEP>] [-r <TEMPERATURE>] [-l <TOPP>] [-o <MODE>] [--per-op] [--gpus <N>]\n");
}

static bool parse(int argc, char** argv, Args& a) {
  for (int i = 1; i < argc; ++i) {
    const std::string k = argv[i];
    auto val = [&]() -> const char* { return i + 1 < argc ? argv[++i] : nullptr; };
    const char* v = nullptr;
    if (k == "--per-op") { a.per_op = true; continue; }
    if (k == "-h" || k == "--help") return false;
    if (!(v = val())) return false;
    if (k == "-m" || k == "--model") a.model = v;
    else if (k == "-t" || k == "--tokenizer") a.tokenizer = v;
    else if (k == "-p" || k == "--prompt") a.prompt = v;
    else if (k == "-s" || k == "--step") a.step = (unsigned)std::strtoul(v, nullptr, 10) & 0xFFFFu;  // u16
    else if (k == "-r" || k == "--temperature") a.temperature = std::strtof(v, nullptr);
    else if (k == "-l" || k == "--topp") a.topp = std::strtof(v, nullptr);
    else if (k == "-o" || k == "--mode") a.mode = v;
    else if (k == "--gpus") a.gpus = std::atoi(v);
    else return false;
  }
  return !a.model.empty() && !a.tokenizer.empty();
}

int main(int argc, char** argv) {
  Args args;
  if (!parse(argc, argv, args)) { usage(); return 2; }
  try {
    std::ifstream rd(args.model, std::ios::binary);
    if (!rd) throw Panic(RAMA_E_IO, "cannot open " + args.model);  // File::open(path).unwrap()
    const Config config = Config::from_file(rd);

    // GPU::new(); --gpus N (or RAMA_GPUS=N): the same single handle, tensor-parallel over N devices of this process
    if (args.gpus > 1 && args.per_op) throw Panic(RAMA_E_INVALID, "--per-op runs the single-device Device ops; drop --gpus");
    const GPU device(args.gpus > 0 ? args.gpus : GPU::gpus_from_env());

    TransformerWeights<HostVec> host_weights;
    TransformerWeights<DevBuf> weights;
    if (args.per_op) {
      host_weights = weights_from_file(rd, config);
      weights = weights_from_weight(host_weights, config, device, /*per_op_views=*/true);
    } else {
      weights = weights_from_path(args.model, config, device);
    }
    RunState<HostVec> host_state = run_state_from_config(config);
    RunState<DevBuf> state = run_state_from_state(host_state, device, /*fused=*/!args.per_op);

    const TransformerWeightsView<DevBuf> wv = TransformerWeightsView<DevBuf>::from_gpu_ws(weights);
    RunStateView<DevBuf> rsv = RunStateView<DevBuf>::from_rs(state);

    const Tokenizer tokenizer(args.tokenizer, config.vocab_size);

    const auto start = std::chrono::steady_clock::now();
    generate(config, tokenizer, args.prompt, args.temperature, args.step, args.topp, wv, rsv, device);
    const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::printf("\n--------------------------------\n");
    std::printf("elapsed: %.3f s, avg tok/s: %g\n", elapsed, (double)(args.step - 1) / elapsed);
  } catch (const Panic& e) {
    std::fprintf(stderr, "thread 'main' panicked: %s\n", e.what());  // the reference unwraps
    return 101;
  }
  return 0;
}
