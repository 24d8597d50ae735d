// service_test.cpp — the engine service (include/rama_b200_service.hpp ≙ engine/src/lib.rs + server/src/batcher.rs) under
// concurrent clients: every request's event stream must equal the text the single-request generate() produces for its prompt.
//
//   service_test <model.bin> <tokenizer.bin> <n_clients> <max_batch> <steps> <temperature>
#include <chrono>
#include <cstdio>
#include <future>
#include <thread>

#include "rama_b200_service.hpp"

using namespace rama;

int main(int argc, char** argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: service_test model.bin tokenizer.bin n_clients max_batch steps temperature\n"); return 2; }
  const int n_clients = std::atoi(argv[3]);
  const usize max_batch = (usize)std::atoi(argv[4]);
  try {
    EngineConfig ec = EngineConfig::from_model_tokenizer(argv[1], argv[2]);
    ec.step = (uint16_t)std::atoi(argv[5]);
    ec.temperature = std::strtof(argv[6], nullptr);
    auto channel = std::make_shared<Channel<ClientRequest>>(30);  // server/src/main.rs:68
    EngineService es(ec, channel, max_batch);
    EngineService::set_global(&es);
    EngineService::global().init();

    const char* prompts[] = {"", "once upon a time", "a", "the cat sat", "zebra", "hello world", "", "abc abc abc", "to be or not"};
    const int n_prompts = (int)(sizeof(prompts) / sizeof(prompts[0]));
    std::vector<std::string> got((usize)n_clients);
    std::vector<int> events((usize)n_clients, 0);
    std::vector<std::promise<void>> done((usize)n_clients);
    std::vector<std::thread> clients;
    for (int i = 0; i < n_clients; ++i) {
      clients.emplace_back([&, i] {
        std::this_thread::sleep_for(std::chrono::microseconds(300 * (i % 5)));  // arrivals spread over a few steps
        ClientRequest cr;
        cr.prompt = prompts[i % n_prompts];
        cr.sender.send = [&got, &events, i](const std::string& s) { got[(usize)i] += s; events[(usize)i] += 1; };
        cr.sender.close = [&done, i] { done[(usize)i].set_value(); };
        channel->send(std::move(cr));
      });
    }
    for (auto& t : clients) t.join();
    for (auto& d : done) d.get_future().wait();
    const usize steps_run = es.steps_run(), max_live = es.max_live();
    es.shutdown();

    // the single-request path on the same device and weights (≙ the reference's one task per request)
    const Config& cfg = es.model_config();
    const usize steps = ec.step > cfg.seq_len ? cfg.seq_len : ec.step;
    const TransformerWeightsView<DevBuf> wv = TransformerWeightsView<DevBuf>::from_gpu_ws(es.weights());
    int bad = 0;
    for (int i = 0; i < n_clients; ++i) {
      RunState<HostVec> hs = run_state_from_config(cfg);
      RunState<DevBuf> st = run_state_from_state(hs, es.device());
      RunStateView<DevBuf> rsv = RunStateView<DevBuf>::from_rs(st);
      const std::string want = generate(cfg, es.tokenizer(), prompts[i % n_prompts], ec.temperature, steps, ec.topp, wv, rsv,
                                        es.device(), nullptr);
      if (want != got[(usize)i] || events[(usize)i] != (int)steps) {
        std::fprintf(stderr, "request %d (prompt '%s'): %d events\n  got  %s\n  want %s\n", i, prompts[i % n_prompts],
                     events[(usize)i], got[(usize)i].c_str(), want.c_str());
        ++bad;
      }
    }
    std::printf("service %s: %d requests, %zu batched steps, up to %zu sequences per step\n", bad ? "MISMATCH" : "ok", n_clients,
                steps_run, max_live);
    if (n_clients > 1 && max_batch > 1 && max_live < 2) { std::fprintf(stderr, "no step ever held two sequences\n"); return 4; }
    return bad ? 1 : 0;
  } catch (const Panic& e) {
    std::fprintf(stderr, "panicked: %s\n", e.what());
    return 101;
  }
}
