// plan.cu -- launch plans: a small cache of instantiated CUDA graphs inside the library.
//
// The entry points of this C-ABI only ENQUEUE: a call with a given (descriptor, pointer arguments, workspace) always
// enqueues the same kernel sequence with the same kernel arguments (tensor maps included) -- tens to thousands of short
// dependent launches for the narrow / deep configurations (BASELINE configs 1-3: 24 .. 150 launches per train step at
// 5-9 us of host time each, more than the GPU needs to run them).  The reference-shaped functional API re-issues that
// sequence every step, and because its output buffers come from a caching allocator the pointer sets REPEAT with a short
// period (ping-pong).  So: the `min_calls`-th time an identical call is seen, the sequence is captured once on a
// library-owned stream (the caller's may be the legacy default stream, which cannot be captured) and instantiated;
// from then on the call is ONE cudaGraphLaunch on the caller's stream.  Replaying also removes the per-launch gaps on
// the device (config 3: 1.43 -> 1.07 ms per step measured with an explicit graph in round 1).
//
// Keys are exact (all words compared, the hash only indexes); a call that is already being captured by the caller
// (PCTrainer, torch.cuda.graph) bypasses the cache; a sequence that fails to capture is remembered and never retried.
// Evicted graphs are destroyed only after their last launch has completed (event query; graveyard otherwise).
#include <cuda_runtime.h>

#include <cstdlib>
#include <list>
#include <mutex>
#include <unordered_map>

#include "host.cuh"

namespace jpcb {

namespace {

struct PlanEntry {
  std::vector<uint64_t> key;
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t last = nullptr;  // recorded after every launch: the graph may be destroyed once this has completed
  unsigned long long nodes = 0;
  int seen = 0;
  bool bad = false;            // capture failed once: enqueue directly from now on
  std::list<uint64_t>::iterator lru;
  bool in_lru = false;
};

struct PlanCache {
  std::mutex mu;
  std::unordered_map<uint64_t, PlanEntry> map;
  std::list<uint64_t> lru;  // front = most recently launched; only entries that own a graph
  std::vector<std::pair<cudaGraphExec_t, cudaEvent_t>> graveyard;
  cudaStream_t cap[64] = {};
  int max_plans = 16, min_calls = 2;
  unsigned long long hits = 0, captures = 0, bypass = 0;
  unsigned long long b_hits = 0, b_caps = 0;  // the capture budget's own counters (reset by clear / configure)
  PlanCache() {
    if (const char* v = getenv("JPCB_PLAN")) max_plans = atoi(v) > 0 ? (atoi(v) == 1 ? 16 : atoi(v)) : 0;
    if (const char* v = getenv("JPCB_PLAN_MIN")) min_calls = atoi(v) > 1 ? atoi(v) : 1;
  }
  void sweep() {
    size_t k = 0;
    for (auto& g : graveyard) {
      if (cudaEventQuery(g.second) == cudaSuccess) { cudaGraphExecDestroy(g.first); cudaEventDestroy(g.second); }
      else { cudaGetLastError(); graveyard[k++] = g; }
    }
    graveyard.resize(k);
  }
  void drop_graph(PlanEntry& e) {
    if (e.exec) graveyard.emplace_back(e.exec, e.last);
    e.exec = nullptr; e.last = nullptr;
    if (e.in_lru) { lru.erase(e.lru); e.in_lru = false; }
  }
};

PlanCache& cache() { static PlanCache* c = new PlanCache(); return *c; }  // (leaked on purpose: no teardown-order hazards at exit)

uint64_t hash_words(const std::vector<uint64_t>& w) {
  uint64_t h = 0x9e3779b97f4a7c15ull ^ (w.size() * 0xff51afd7ed558ccdull);
  for (uint64_t v : w) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 29; }
  return h;
}

// experiment / test switches that are read from the environment on every call and change the kernel sequence: part of the key
uint64_t env_signature() {
  uint64_t h = 0;
  for (const char* name : {"JPCB_STAGED", "JPCB_TS_STG", "JPCB_SOLVE_MIN_ROUNDS", "JPCB_TS_PROMO", "JPCB_CHAIN_FFWD", "JPCB_TS_HINTS"}) {
    const char* v = getenv(name);
    h = h * 1000003ull + (v ? (uint64_t)atoll(v) + 1ull : 0ull);
  }
  return h;
}

}  // namespace

void PlanKey::bytes(const void* p, size_t n) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  w.push_back(n);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v; memcpy(&v, c + i, 8); w.push_back(v); }
  if (i < n) { uint64_t v = 0; memcpy(&v, c + i, n - i); w.push_back(v); }
}

int plan_run(PlanKey& key, void* stream, const std::function<int(void*)>& body) {
  PlanCache& C = cache();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C.max_plans <= 0) return body(stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();  // (the caller is capturing: its graph will hold the kernels themselves)
    ++C.bypass;
    return body(stream);
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return body(stream);
  key.ptr(stream);
  key.word((uint64_t)dev);
  key.word(env_signature());
  key.word(reproducible() ? 1u : 0u);
  const uint64_t h = hash_words(key.w);

  std::unique_lock<std::mutex> lock(C.mu);
  if (!C.graveyard.empty()) C.sweep();
  auto it = C.map.find(h);
  if (it == C.map.end()) {
    if (C.map.size() >= (size_t)C.max_plans * 8) {  // forget keys that never came back
      for (auto m = C.map.begin(); m != C.map.end();) m = m->second.exec ? std::next(m) : C.map.erase(m);
    }
    PlanEntry& e = C.map[h];
    e.key = key.w;
    e.seen = 1;
    lock.unlock();
    return body(stream);
  }
  PlanEntry& e = it->second;
  if (e.key != key.w || e.bad) { lock.unlock(); return body(stream); }  // (hash collision: leave the first owner alone)
  if (e.exec) {
    cudaError_t le = cudaGraphLaunch(e.exec, st);
    if (le != cudaSuccess) return fail(JPCB_ECUDA, "cudaGraphLaunch (launch plan) -> %s", cudaGetErrorString(le));
    cudaEventRecord(e.last, st);
    count_launch((int)e.nodes);
    ++C.hits; ++C.b_hits;
    C.lru.splice(C.lru.begin(), C.lru, e.lru);
    return JPCB_OK;
  }
  if (++e.seen < C.min_calls) { lock.unlock(); return body(stream); }
  // capture budget: a caller whose pointer sets repeat once or twice and then move on would pay for captures that are
  // never replayed -- new plans are only built while replays keep up with captures (existing plans keep replaying)
  if (C.b_caps >= 4 && C.b_hits < 2 * C.b_caps) { lock.unlock(); return body(stream); }
  // ---- capture (under the lock: one capture stream per device) ----
  if (!C.cap[dev] && cudaStreamCreateWithFlags(&C.cap[dev], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); e.bad = true; lock.unlock(); return body(stream); }
  if (cudaStreamBeginCapture(C.cap[dev], cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); e.bad = true; lock.unlock(); return body(stream); }
  const unsigned long long before = jpcb_launch_count();
  const int rc = body(C.cap[dev]);
  const unsigned long long nodes = jpcb_launch_count() - before;
  cudaGraph_t g = nullptr;
  const cudaError_t ee = cudaStreamEndCapture(C.cap[dev], &g);
  count_launch(-(int)nodes);  // (captured, not launched yet)
  if (rc != JPCB_OK || ee != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    e.bad = true;
    lock.unlock();
    return rc != JPCB_OK ? rc : body(stream);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, g, 0);
  cudaGraphDestroy(g);
  cudaEvent_t ev = nullptr;
  if (ie != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
    if (exec) cudaGraphExecDestroy(exec);
    cudaGetLastError();
    e.bad = true;
    lock.unlock();
    return body(stream);
  }
  while ((int)C.lru.size() >= C.max_plans) {  // evict the least recently launched plan
    auto victim = C.map.find(C.lru.back());
    if (victim == C.map.end()) { C.lru.pop_back(); continue; }
    C.drop_graph(victim->second);
    victim->second.seen = 0;
  }
  e.exec = exec; e.last = ev; e.nodes = nodes;
  C.lru.push_front(h); e.lru = C.lru.begin(); e.in_lru = true;
  ++C.captures; ++C.b_caps;
  const cudaError_t le = cudaGraphLaunch(exec, st);
  if (le != cudaSuccess) return fail(JPCB_ECUDA, "cudaGraphLaunch (launch plan) -> %s", cudaGetErrorString(le));
  cudaEventRecord(ev, st);
  count_launch((int)nodes);
  return JPCB_OK;
}

}  // namespace jpcb

extern "C" {

int jpcb_plan_cache_configure(int max_plans, int min_calls) {
  using namespace jpcb;
  PlanCache& C = cache();
  std::lock_guard<std::mutex> lock(C.mu);
  C.max_plans = max_plans < 0 ? 0 : max_plans;
  C.min_calls = min_calls < 1 ? 1 : min_calls;
  C.b_hits = C.b_caps = 0;
  if (C.max_plans == 0) {
    for (auto& m : C.map) C.drop_graph(m.second);
    C.map.clear();
    C.sweep();
  }
  return JPCB_OK;
}

int jpcb_plan_cache_clear(void) {
  using namespace jpcb;
  PlanCache& C = cache();
  std::lock_guard<std::mutex> lock(C.mu);
  for (auto& m : C.map) C.drop_graph(m.second);
  C.map.clear();
  C.sweep();
  C.b_hits = C.b_caps = 0;
  return JPCB_OK;
}

void jpcb_plan_cache_stats(unsigned long long* hits, unsigned long long* captures, unsigned long long* bypassed) {
  using namespace jpcb;
  PlanCache& C = cache();
  std::lock_guard<std::mutex> lock(C.mu);
  if (hits) *hits = C.hits;
  if (captures) *captures = C.captures;
  if (bypassed) *bypassed = C.bypass;
}

}  // extern "C"
