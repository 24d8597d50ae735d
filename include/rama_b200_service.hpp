// rama_b200_service.hpp — C++ mirror of the reference's engine service (the caller of the decode path on the server side),
// wired to the batched decode entry points.
//
//   reference                                                            here
//   engine/src/lib.rs:14-46     EngineConfig (+ from_model_tokenizer)    rama::EngineConfig
//   engine/src/lib.rs:48-54     ClientRequest { prompt, sender }          rama::ClientRequest
//   server/src/main.rs:68       async_channel::bounded(30)                rama::Channel<T> (bounded, blocking)
//   server/src/batcher.rs:8-38  get_batch(receiver, prompts, n, wait)     rama::get_batch
//   engine/src/lib.rs:56-125    ENGINE_SERVICE / EngineService::{new,init,global}   rama::EngineService
//   engine/src/lib.rs:127-160   handler(): one task + RunState per request, generate_stream per request
//   engine/src/transformer/mod.rs:208-248  generate_stream: forward, sample, decode, send event ("\n" → "\\n"), sleep 1 ms
//
// The reference handler gives every request its own forward()/sample() loop, so n live requests stream the weights n times per
// token, and its batcher is dead code.  Here the handler thread IS the batcher: it admits requests between steps
// (get_batch semantics: up to the free slots, bounded wait when idle), prefills each prompt (rama_prefill), and advances all
// live requests with ONE rama_forward_batch + rama_sample_batch per token (continuous batching).  Per request the event stream is
// the one generate_stream produces: the decoded forced prompt tokens, then the decoded samples, `step` events in total; the
// 1 ms sleep per token (mod.rs:245) is not on this path.  Header-only, over include/rama_b200.hpp.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>

#include "rama_b200.hpp"

namespace rama {

// ---- lib.rs:14-46 ----
struct EngineConfig {
  std::string model, tokenizer;
  uint16_t step = 255;
  float temperature = 1.0f, topp = 0.9f;
  std::string mode = "generate";
  static EngineConfig from_model_tokenizer(std::string model, std::string tokenizer) {
    EngineConfig c;
    c.model = std::move(model);
    c.tokenizer = std::move(tokenizer);
    return c;
  }
};

// the SSE side of a request: `send` once per token (the event's data), `close` when the stream ends (the reference drops the Sender)
struct EventSender {
  std::function<void(const std::string&)> send;
  std::function<void()> close;
};

// ---- lib.rs:48-54 ----
struct ClientRequest {
  std::string prompt;
  EventSender sender;
};

// ---- async_channel::bounded(cap): blocking send when full, recv with an optional deadline ----
template <class T>
class Channel {
 public:
  explicit Channel(usize cap) : cap_(cap) {}
  void send(T v) {
    std::unique_lock<std::mutex> lk(mu_);
    not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
    if (closed_) return;
    q_.push_back(std::move(v));
    not_empty_.notify_one();
  }
  // true + value, or false on timeout / closed-and-empty
  bool recv(T& out, std::chrono::steady_clock::time_point deadline) {
    std::unique_lock<std::mutex> lk(mu_);
    if (!not_empty_.wait_until(lk, deadline, [&] { return !q_.empty() || closed_; })) return false;
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    not_full_.notify_one();
    return true;
  }
  void close() {
    std::lock_guard<std::mutex> lk(mu_);
    closed_ = true;
    not_empty_.notify_all();
    not_full_.notify_all();
  }
  bool closed() const { std::lock_guard<std::mutex> lk(mu_); return closed_; }

 private:
  usize cap_;
  mutable std::mutex mu_;
  std::condition_variable not_empty_, not_full_;
  std::deque<T> q_;
  bool closed_ = false;
};

// ---- batcher.rs:8-38: append received items until `prompts` holds batch_size of them or wait_time has passed ----
template <class T>
void get_batch(Channel<T>& receiver, std::vector<T>& prompts, usize batch_size, std::chrono::duration<double> wait_time) {
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration_cast<std::chrono::steady_clock::duration>(wait_time);
  while (prompts.size() < batch_size) {
    T item;
    if (!receiver.recv(item, deadline)) return;
    prompts.push_back(std::move(item));
  }
}

// ---- continuous batching over rama_forward_batch (the C++ twin of rama_b200/serving.py) ----
class BatchedEngine {
 public:
  struct Request {
    usize rid = 0;
    std::vector<int32_t> prompt_tokens;
    usize steps = 0;
    std::function<bool(usize, int32_t)> on_token;  // (request id, token): generate_stream's sender; false = the request's task died
    std::function<void(usize)> on_done;
    std::vector<int32_t> tokens;                   // `next` of every step so far (mod.rs:187-203)
    rama_session* session = nullptr;
    int32_t pos = 0, cur = 1;
  };

  BatchedEngine(const GPU& gpu, const Config& cfg, usize max_batch = 64, float temperature = 0.0f, float topp = 0.9f,
                usize prefill_min_rows = 2)
      : gpu_(gpu), cfg_(cfg), max_batch_(max_batch), temperature_(temperature), topp_(topp), prefill_min_rows_(prefill_min_rows) {
    ck(rama_batch_create(gpu.ctx, (int32_t)max_batch, &batch_));
  }
  ~BatchedEngine() {
    if (batch_) rama_batch_destroy(batch_);
    for (auto& r : live_) if (r->session) rama_session_destroy(r->session);
    for (auto* s : pool_) rama_session_destroy(s);
  }
  BatchedEngine(const BatchedEngine&) = delete;
  BatchedEngine& operator=(const BatchedEngine&) = delete;

  // steps as in generate(): total number of positions (prompt included); steps > seq_len is the reference's slice panic
  usize submit(std::vector<int32_t> prompt_tokens, usize steps, std::function<bool(usize, int32_t)> on_token = nullptr,
               std::function<void(usize)> on_done = nullptr) {
    if (steps > cfg_.seq_len) throw Panic(RAMA_E_STATE, "steps exceeds seq_len (infer.rs:32 slice)");
    auto r = std::make_unique<Request>();
    r->rid = next_id_++;
    r->prompt_tokens = std::move(prompt_tokens);
    r->steps = steps;
    r->on_token = std::move(on_token);
    r->on_done = std::move(on_done);
    const usize rid = r->rid;
    waiting_.push_back(std::move(r));
    return rid;
  }
  usize free_slots() const { return max_batch_ - live_.size() > waiting_.size() ? max_batch_ - live_.size() - waiting_.size() : 0; }
  bool idle() const { return live_.empty() && waiting_.empty(); }
  usize live() const { return live_.size(); }

  // one token for every live request; returns how many sequences the step advanced
  usize step() {
    admit();
    if (live_.empty()) return 0;
    const usize n = live_.size();
    sess_.resize(n); tok_.resize(n); pos_.resize(n); next_.resize(n);
    for (usize i = 0; i < n; ++i) { sess_[i] = live_[i]->session; tok_[i] = live_[i]->cur; pos_[i] = live_[i]->pos; }
    ck(rama_forward_batch(batch_, sess_.data(), tok_.data(), pos_.data(), (int32_t)n));
    ck(rama_sample_batch(batch_, sess_.data(), (int32_t)n, temperature_, topp_, next_.data()));
    std::vector<std::unique_ptr<Request>> still;
    for (usize i = 0; i < n; ++i) {
      Request& r = *live_[i];
      const bool alive = emit(r, next_[i]);
      r.pos += 1;
      r.cur = next_[i];
      if (!alive || r.tokens.size() >= r.steps) finish(std::move(live_[i]));
      else still.push_back(std::move(live_[i]));
    }
    live_ = std::move(still);
    return n;
  }
  void run_until_idle() { while (!idle()) step(); }
  // A failed step (e.g. the empty top-p candidate list the reference panics on, infer.rs:66) ends the requests that were
  // in it — in the reference only that request's task dies — and the engine keeps serving: their streams are closed, the
  // sessions go back to the pool (the next batched step rewrites each session's error flag).
  void abort_live() {
    for (auto& r : live_) finish(std::move(r));
    live_.clear();
  }

 private:
  bool emit(Request& r, int32_t token) {
    r.tokens.push_back(token);
    return r.on_token ? r.on_token(r.rid, token) : true;
  }
  void finish(std::unique_ptr<Request> r) {
    if (r->session) { pool_.push_back(r->session); r->session = nullptr; }  // KV rows are rewritten before they are read again
    if (r->on_done) r->on_done(r->rid);
  }
  void admit() {
    while (!waiting_.empty() && live_.size() < max_batch_) {
      std::unique_ptr<Request> r = std::move(waiting_.front());
      waiting_.pop_front();
      if (r->steps == 0) { finish(std::move(r)); continue; }
      if (pool_.empty()) { ck(rama_session_create(gpu_.ctx, &r->session)); }
      else { r->session = pool_.back(); pool_.pop_back(); }
      std::vector<int32_t> rows;                     // BOS first (mod.rs:182), then the forced prompt
      rows.push_back(1);
      rows.insert(rows.end(), r->prompt_tokens.begin(), r->prompt_tokens.end());
      const usize n_forced = std::min(r->prompt_tokens.size(), r->steps);
      bool alive = true;
      for (usize i = 0; i < n_forced && alive; ++i) alive = emit(*r, r->prompt_tokens[i]);  // positions 0..n_forced-1 emit the prompt (mod.rs:189-191)
      if (!alive || n_forced >= r->steps) { finish(std::move(r)); continue; }  // the prompt alone exhausts the step budget
      const usize n_feed = n_forced + 1;             // tokens fed at positions 0..n_forced; the last one joins the batched step
      if (n_feed - 1 >= prefill_min_rows_) {
        ck(rama_prefill(r->session, rows.data(), (int32_t)(n_feed - 1), 0, nullptr, nullptr, nullptr));
      } else {
        for (usize p = 0; p + 1 < n_feed; ++p) ck(rama_forward(r->session, rows[p], (int32_t)p));
        ck(rama_session_sync(r->session));
      }
      r->pos = (int32_t)(n_feed - 1);
      r->cur = rows[n_feed - 1];
      live_.push_back(std::move(r));
    }
  }

  const GPU& gpu_;
  Config cfg_;
  usize max_batch_;
  float temperature_, topp_;
  usize prefill_min_rows_;
  rama_batch* batch_ = nullptr;
  std::vector<rama_session*> pool_;
  std::vector<std::unique_ptr<Request>> live_;
  std::deque<std::unique_ptr<Request>> waiting_;
  usize next_id_ = 0;
  std::vector<rama_session*> sess_;
  std::vector<int32_t> tok_, pos_, next_;
};

// ---- lib.rs:56-160 ----
class EngineService {
 public:
  // ≙ EngineService::new(eng_config, receiver): model + tokenizer loaded once, weights shipped to HBM by the library's loader
  EngineService(EngineConfig eng_config, std::shared_ptr<Channel<ClientRequest>> receiver, usize max_batch = 64,
                std::chrono::duration<double> idle_wait = std::chrono::milliseconds(2))
      : eng_config_(std::move(eng_config)), receiver_(std::move(receiver)), max_batch_(max_batch), idle_wait_(idle_wait) {
    std::ifstream rd(eng_config_.model, std::ios::binary);
    if (!rd) throw Panic(RAMA_E_IO, "cannot open " + eng_config_.model);  // File::open(path).unwrap()
    model_config_ = Config::from_file(rd);
    weights_ = weights_from_path(eng_config_.model, model_config_, device_);
    tokenizer_ = std::make_unique<Tokenizer>(eng_config_.tokenizer, model_config_.vocab_size);
  }
  ~EngineService() { shutdown(); }

  // ≙ ENGINE_SERVICE.set(...) / EngineService::global()
  static void set_global(EngineService* es) { global_slot() = es; }
  static EngineService& global() {
    if (!global_slot()) throw Panic(RAMA_E_STATE, "Engine not initialized");
    return *global_slot();
  }

  // ≙ init(): tokio::spawn(handler())
  void init() {
    if (handler_.joinable()) return;
    handler_ = std::thread([this] { handler(); });
  }
  void shutdown() {
    stop_ = true;
    if (receiver_) receiver_->close();
    if (handler_.joinable()) handler_.join();
  }
  const Config& model_config() const { return model_config_; }
  const Tokenizer& tokenizer() const { return *tokenizer_; }
  const GPU& device() const { return device_; }
  const TransformerWeights<DevBuf>& weights() const { return weights_; }
  const EngineConfig& eng_config() const { return eng_config_; }
  usize steps_run() const { return steps_run_; }
  usize max_live() const { return max_live_; }

 private:
  static EngineService*& global_slot() { static EngineService* p = nullptr; return p; }

  // ≙ handler() (lib.rs:127-160), batching instead of one task per request
  void handler() {
    BatchedEngine eng(device_, model_config_, max_batch_, eng_config_.temperature, eng_config_.topp);
    std::vector<ClientRequest> fresh;
    while (!stop_ || !eng.idle()) {
      fresh.clear();
      const usize room = eng.free_slots();
      if (room > 0 && !stop_)  // between steps: take what is queued; when nothing is running, wait a little for company
        get_batch(*receiver_, fresh, room, eng.idle() ? idle_wait_ : std::chrono::duration<double>(0));
      for (auto& cr : fresh) admit(eng, std::move(cr));
      if (eng.idle()) continue;
      usize n = 0;
      try {
        n = eng.step();
      } catch (const Panic& e) {  // must not escape the handler thread (std::terminate would take the server down)
        std::fprintf(stderr, "batched step panicked: %s\n", e.what());
        eng.abort_live();
        continue;
      }
      steps_run_ += 1;
      if (n > max_live_) max_live_ = n;
    }
  }
  void admit(BatchedEngine& eng, ClientRequest cr) {
    auto sender = std::make_shared<EventSender>(std::move(cr.sender));
    try {
      std::vector<int32_t> prompt;
      if (!cr.prompt.empty())
        for (usize t : tokenizer_->encode(cr.prompt)) prompt.push_back((int32_t)t);
      eng.submit(std::move(prompt), eng_config_.step > model_config_.seq_len ? model_config_.seq_len : eng_config_.step,
                 [this, sender](usize, int32_t token) -> bool {
                   try {
                     std::string s = tokenizer_->decode((usize)token);
                     for (usize p = 0; (p = s.find('\n', p)) != std::string::npos; p += 2) s.replace(p, 1, "\\n");  // mod.rs:244
                     if (sender->send) sender->send(s);
                     return true;
                   } catch (const Panic& e) {
                     // decode() of a piece like "<unk>" panics in the reference (bpe.rs:110): that request's task dies there and its
                     // stream ends; the other requests go on
                     std::fprintf(stderr, "request panicked: %s\n", e.what());
                     return false;
                   }
                 },
                 [sender](usize) { if (sender->close) sender->close(); });
    } catch (const Panic& e) {
      // the reference's request task panics (encode of an unknown character, bpe.rs:55): this stream ends, the service lives on
      std::fprintf(stderr, "request panicked: %s\n", e.what());
      if (sender->close) sender->close();
    }
  }

  EngineConfig eng_config_;
  std::shared_ptr<Channel<ClientRequest>> receiver_;
  usize max_batch_;
  std::chrono::duration<double> idle_wait_;
  GPU device_;
  Config model_config_;
  TransformerWeights<DevBuf> weights_;
  std::unique_ptr<Tokenizer> tokenizer_;
  std::thread handler_;
  std::atomic<bool> stop_{false};
  std::atomic<usize> steps_run_{0}, max_live_{0};
};

}  // namespace rama
