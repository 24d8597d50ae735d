// group.cu — single-process tensor parallelism: one GROUP context that owns a rank context per device
// (rama_ctx_create_multi).  The reference is one binary with one `GPU` handle (engine/src/main.rs:70-98,
// engine/src/lib.rs:99-119): nothing in it can launch a process per GPU, so `--features gpu` callers get tensor
// parallelism through this handle — every entry point that receives it fans the call out to the ranks.
//
// Each rank has a dedicated host thread (its device stays current there), so the ranks' launch sequences are issued
// concurrently — a rank-level call may synchronise its own stream without holding up the enqueue of the peers it
// exchanges with.  On the device nothing differs from the multi-process layout: the ranks run the same kernels and
// exchange through peer memory (cudaDeviceEnablePeerAccess instead of CUDA IPC); NCCL is not used at all.
#include "internal.cuh"

#include <functional>

struct GroupPool {
  std::mutex run_mu;  // one fan-out at a time (callers on several host threads take turns)
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::vector<std::thread> th;
  const std::function<int(int)>* fn = nullptr;
  unsigned long long gen = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> msg;
};

static void pool_worker(GroupPool* p, int r, int device) {
  cudaSetDevice(device);
  unsigned long long seen = 0;
  for (;;) {
    const std::function<int(int)>* fn;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv_go.wait(lk, [&] { return p->stop || p->gen != seen; });
      if (p->stop) return;
      seen = p->gen;
      fn = p->fn;
    }
    const int rc = (*fn)(r);
    {
      std::lock_guard<std::mutex> lk(p->mu);
      p->rc[r] = rc;
      if (rc != RAMA_OK) p->msg[r] = rama_last_error();
      if (--p->pending == 0) p->cv_done.notify_all();
    }
  }
}

// runs fn(rank index) on every rank's thread and waits; returns the first failure (message prefixed with the rank)
int group_run(rama_ctx* g, const std::function<int(int)>& fn) {
  GroupPool* p = g->pool;
  std::lock_guard<std::mutex> run_lk(p->run_mu);
  const int n = (int)g->ranks.size();
  {
    std::unique_lock<std::mutex> lk(p->mu);
    p->fn = &fn;
    p->pending = n;
    for (int r = 0; r < n; ++r) { p->rc[r] = RAMA_OK; p->msg[r].clear(); }
    ++p->gen;
    p->cv_go.notify_all();
    p->cv_done.wait(lk, [&] { return p->pending == 0; });
  }
  // a rank that failed on its own makes its peers time out waiting for it: report the cause, not the symptom
  for (int pass = 0; pass < 2; ++pass)
    for (int r = 0; r < n; ++r)
      if (p->rc[r] != RAMA_OK && (pass == 1 || p->rc[r] != RAMA_E_NCCL)) return fail(p->rc[r], "rank %d: %s", r, p->msg[r].c_str());
  return RAMA_OK;
}

static void pool_stop(GroupPool* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->stop = true;
  }
  p->cv_go.notify_all();
  for (auto& t : p->th) t.join();
  delete p;
}

void group_destroy(rama_ctx* g) {
  pool_stop(g->pool);
  g->pool = nullptr;
  for (rama_ctx* rc : g->ranks) {
    rc->group = nullptr;
    rama_ctx_destroy(rc);
  }
  g->ranks.clear();
}

extern "C" int rama_ctx_create_multi(int32_t n_gpus, const int32_t* devices, rama_ctx** out) {
  if (!out) return fail(RAMA_E_INVALID, "out is NULL");
  int have = 0;
  CK(cudaGetDeviceCount(&have));
  if (n_gpus < 1 || n_gpus > kMaxPeers) return fail(RAMA_E_INVALID, "n_gpus %d outside [1, %d]", n_gpus, kMaxPeers);
  std::vector<int> dev(n_gpus);
  for (int i = 0; i < n_gpus; ++i) {
    dev[i] = devices ? devices[i] : i;
    if (dev[i] < 0 || dev[i] >= have) return fail(RAMA_E_CUDA, "device %d not present (%d CUDA devices)", dev[i], have);
    for (int j = 0; j < i; ++j)
      if (dev[j] == dev[i]) return fail(RAMA_E_INVALID, "device %d listed twice", dev[i]);
  }
  if (n_gpus == 1) return rama_ctx_create(dev[0], nullptr, out);  // a plain context: nothing to exchange
  // every pair of devices must be peer-addressable (NVLink / NVSwitch on an HGX board)
  for (int i = 0; i < n_gpus; ++i) {
    CK(cudaSetDevice(dev[i]));
    for (int j = 0; j < n_gpus; ++j) {
      if (i == j) continue;
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, dev[i], dev[j]));
      if (!can) return fail(RAMA_E_CUDA, "device %d cannot address device %d's memory (no peer access)", dev[i], dev[j]);
      cudaError_t e = cudaDeviceEnablePeerAccess(dev[j], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(RAMA_E_CUDA, "cudaDeviceEnablePeerAccess(%d → %d): %s", dev[i], dev[j], cudaGetErrorString(e));
    }
  }
  rama_ctx* g = new rama_ctx();
  g->device = dev[0];
  g->world = n_gpus;
  for (int r = 0; r < n_gpus; ++r) {
    rama_ctx* rc = nullptr;
    int e = rama_ctx_create(dev[r], nullptr, &rc);
    if (e != RAMA_OK) { group_destroy(g); delete g; return e; }
    rc->rank = r;
    rc->world = n_gpus;
    rc->p2p = 1;          // the only exchange a group has
    rc->persistent = 0;
    rc->group = g;
    g->ranks.push_back(rc);
  }
  g->sm_count = g->ranks[0]->sm_count;
  GroupPool* p = new GroupPool();
  p->rc.assign(n_gpus, RAMA_OK);
  p->msg.assign(n_gpus, std::string());
  g->pool = p;
  for (int r = 0; r < n_gpus; ++r) p->th.emplace_back(pool_worker, p, r, dev[r]);
  *out = g;
  return RAMA_OK;
}

// after a load: the group mirrors the (global) configuration of its ranks
void group_adopt_config(rama_ctx* g) {
  const rama_ctx* r0 = g->ranks[0];
  g->cfg = r0->cfg;
  g->D = r0->D; g->F = r0->F; g->L = r0->L; g->H = r0->H; g->V = r0->V; g->T = r0->T; g->hs = r0->hs;
  g->loaded = true;
  for (const rama_ctx* rc : g->ranks) g->loaded = g->loaded && rc->loaded;
}
