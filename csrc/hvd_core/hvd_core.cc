// hvdcore: background negotiation / fusion / execution engine (see hvd_core.h for the design).
//
// Threads: callers (any thread) push requests under `mu`; ONE engine thread per process owns the rendezvous, the
// replicated coordinator state (`table`, `cache`, `joined`), the fusion buffers and the timeline. Handles are completed
// under `mu` and waiters are woken through `done_cv`.
#include "hvd_core.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../runtime/rendezvous.h"

using b200mpi::kRvMailbox;
using b200mpi::now_ns;
using b200mpi::Rendezvous;

namespace {

thread_local std::string t_err;
int fail(int code, const std::string& msg) { t_err = msg; return code; }

constexpr size_t kInline = 248;          // negotiation bytes that ride in the first (small) exchange of a cycle
constexpr size_t kTail = 1024;           // end of every mailbox: two alternating slots for that exchange
constexpr size_t kBody = kRvMailbox - kTail;   // what long messages / mailbox data phases may use (a multiple of 8)
constexpr size_t kMaxBlob = 1024;        // HVD_EXCHANGE payload limit
constexpr uint32_t kFlagShutdown = 1u, kFlagStallShutdown = 2u, kFlagAutotune = 4u;   // autotune: rank 0's tunables are adopted by all
const char* const kOpName[] = {"ALLREDUCE", "ALLGATHER", "BROADCAST", "ALLTOALL", "BARRIER", "JOIN", "EXCHANGE"};

size_t esize(int dt) {
  switch (dt) {
    case HVD_U8: case HVD_I8: case HVD_BOOL: return 1;
    case HVD_I16: case HVD_F16: case HVD_BF16: return 2;
    case HVD_I32: case HVD_F32: return 4;
    case HVD_I64: case HVD_F64: return 8;
    default: return 0;
  }
}

// ---------------------------------------------------------------------------------------------- half / bf16 --
inline float bf16_to_f(uint16_t v) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); return f; }
inline uint16_t f_to_bf16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                           // round to nearest even
  return (uint16_t)(u >> 16);
}
inline float f16_to_f(uint16_t h) {
  const uint32_t s = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
  if (e == 0) {
    if (m == 0) u = s;
    else { e = 127 - 15 + 1; while (!(m & 0x400u)) { m <<= 1; e--; } u = s | (e << 23) | ((m & 0x3ffu) << 13); }
  } else if (e == 31) u = s | 0x7f800000u | (m << 13);
  else u = s | ((e + 127 - 15) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}
inline uint16_t f_to_f16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  const uint32_t s = (u >> 16) & 0x8000u;
  const int32_t e = (int32_t)((u >> 23) & 0xffu) - 127 + 15;
  uint32_t m = u & 0x7fffffu;
  if (((u >> 23) & 0xffu) == 0xffu) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0));
  if (e >= 31) return (uint16_t)(s | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)s;
    m |= 0x800000u;
    const int shift = 14 - e;
    uint32_t r = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (r & 1u))) r++;
    return (uint16_t)(s | r);
  }
  uint32_t r = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
  return (uint16_t)(s | r);
}

template <typename T> inline T comb(T a, T b, int op) {
  switch (op) {
    case HVD_SUM: return (T)(a + b);
    case HVD_MIN: return b < a ? b : a;
    case HVD_MAX: return a < b ? b : a;
    default: return (T)(a * b);
  }
}
template <typename T> void fold_t(void* acc, const void* x, size_t n, int op) {
  T* a = (T*)acc; const T* b = (const T*)x;
  for (size_t i = 0; i < n; i++) a[i] = comb<T>(a[i], b[i], op);
}
void fold_bool(void* acc, const void* x, size_t n, int op) {  // SUM/MAX = or, MIN/PROD = and
  uint8_t* a = (uint8_t*)acc; const uint8_t* b = (const uint8_t*)x;
  for (size_t i = 0; i < n; i++) a[i] = (op == HVD_SUM || op == HVD_MAX) ? (uint8_t)((a[i] | b[i]) != 0) : (uint8_t)((a[i] && b[i]) ? 1 : 0);
}
// acc (fp32 scratch) op= x (16-bit floats): the running value stays in fp32 until every rank is folded in
template <float (*TO)(uint16_t)> void fold16(float* acc, const void* x, size_t n, int op) {
  const uint16_t* b = (const uint16_t*)x;
  for (size_t i = 0; i < n; i++) acc[i] = comb<float>(acc[i], TO(b[i]), op);
}

void fold(void* acc, const void* x, size_t n, int dt, int op) {
  switch (dt) {
    case HVD_U8: fold_t<uint8_t>(acc, x, n, op); break;
    case HVD_I8: fold_t<int8_t>(acc, x, n, op); break;
    case HVD_I16: fold_t<int16_t>(acc, x, n, op); break;
    case HVD_I32: fold_t<int32_t>(acc, x, n, op); break;
    case HVD_I64: fold_t<int64_t>(acc, x, n, op); break;
    case HVD_F32: fold_t<float>(acc, x, n, op); break;
    case HVD_F64: fold_t<double>(acc, x, n, op); break;
    case HVD_BOOL: fold_bool(acc, x, n, op); break;
    default: break;
  }
}

void scale_buf(void* buf, size_t n, int dt, double f) {
  if (f == 1.0) return;
  switch (dt) {
    case HVD_F32: { float* p = (float*)buf; const float g = (float)f; for (size_t i = 0; i < n; i++) p[i] *= g; break; }
    case HVD_F64: { double* p = (double*)buf; for (size_t i = 0; i < n; i++) p[i] *= f; break; }
    case HVD_F16: { uint16_t* p = (uint16_t*)buf; const float g = (float)f; for (size_t i = 0; i < n; i++) p[i] = f_to_f16(f16_to_f(p[i]) * g); break; }
    case HVD_BF16: { uint16_t* p = (uint16_t*)buf; const float g = (float)f; for (size_t i = 0; i < n; i++) p[i] = f_to_bf16(bf16_to_f(p[i]) * g); break; }
    case HVD_U8: { uint8_t* p = (uint8_t*)buf; for (size_t i = 0; i < n; i++) p[i] = (uint8_t)llround(p[i] * f); break; }
    case HVD_I8: { int8_t* p = (int8_t*)buf; for (size_t i = 0; i < n; i++) p[i] = (int8_t)llround(p[i] * f); break; }
    case HVD_I16: { int16_t* p = (int16_t*)buf; for (size_t i = 0; i < n; i++) p[i] = (int16_t)llround(p[i] * f); break; }
    case HVD_I32: { int32_t* p = (int32_t*)buf; for (size_t i = 0; i < n; i++) p[i] = (int32_t)llround(p[i] * f); break; }
    case HVD_I64: { int64_t* p = (int64_t*)buf; for (size_t i = 0; i < n; i++) p[i] = (int64_t)llround((double)p[i] * f); break; }
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------- wire format --
struct Writer {
  std::string s;
  void raw(const void* p, size_t n) { s.append((const char*)p, n); }
  template <typename T> void put(T v) { raw(&v, sizeof(T)); }
  void str(const std::string& v) { put<uint16_t>((uint16_t)v.size()); raw(v.data(), v.size()); }
};
struct Reader {
  const char* p; const char* end; bool ok = true;
  Reader(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
  bool raw(void* out, size_t n) { if ((size_t)(end - p) < n) { ok = false; return false; } memcpy(out, p, n); p += n; return true; }
  template <typename T> T get() { T v{}; raw(&v, sizeof(T)); return v; }
  std::string str() { uint16_t n = get<uint16_t>(); if (!ok || (size_t)(end - p) < n) { ok = false; return ""; } std::string v(p, n); p += n; return v; }
  bool done() const { return p >= end; }
};

struct Request {
  std::string name;
  uint8_t op = 0, dtype = 0, redop = 0, devkind = 0;   // devkind: 0 host, 1 device
  int32_t root = 0;
  double pre = 1.0, post = 1.0;
  int64_t count = 0;
  std::vector<int64_t> extra;
  std::string blob;
};

// ops whose request is identical on every rank (an allgather carries its own element count, an alltoall its own splits)
bool cacheable(const Request& r) { return r.op == HVD_ALLREDUCE || r.op == HVD_BROADCAST || r.op == HVD_BARRIER; }
bool same_sig(const Request& a, const Request& b) {
  return a.op == b.op && a.dtype == b.dtype && a.redop == b.redop && a.devkind == b.devkind && a.root == b.root &&
         memcmp(&a.pre, &b.pre, 8) == 0 && memcmp(&a.post, &b.post, 8) == 0 && a.count == b.count && a.extra == b.extra;
}
void encode(Writer& w, const Request& r) {
  w.put<uint8_t>(0);
  w.put<uint8_t>(r.op); w.put<uint8_t>(r.dtype); w.put<uint8_t>(r.redop); w.put<uint8_t>(r.devkind);
  w.put<int32_t>(r.root); w.put<double>(r.pre); w.put<double>(r.post); w.put<int64_t>(r.count);
  w.str(r.name);
  w.put<uint16_t>((uint16_t)r.extra.size());
  for (int64_t v : r.extra) w.put<int64_t>(v);
  w.str(r.blob);
}
bool decode_full(Reader& rd, Request* r) {
  r->op = rd.get<uint8_t>(); r->dtype = rd.get<uint8_t>(); r->redop = rd.get<uint8_t>(); r->devkind = rd.get<uint8_t>();
  r->root = rd.get<int32_t>(); r->pre = rd.get<double>(); r->post = rd.get<double>(); r->count = rd.get<int64_t>();
  r->name = rd.str();
  const uint16_t ne = rd.get<uint16_t>();
  r->extra.resize(rd.ok ? ne : 0);
  for (auto& v : r->extra) v = rd.get<int64_t>();
  r->blob = rd.str();
  return rd.ok && r->op <= HVD_EXCHANGE;
}

// ------------------------------------------------------------------------------------------------ engine state --
struct LocalOp {
  Request req;
  const void* in = nullptr;
  void* out = nullptr;
  int handle = 0;
  int device = -1;
  void* ready_event = nullptr;
};
struct Handle { bool done = false; int status = 0; std::string err; };

struct Entry {             // one tensor name being negotiated (replicated on every rank)
  Request first;
  uint64_t seq = 0;
  uint64_t bits = 0;       // ranks that submitted it
  uint64_t t_first = 0, t_warned = 0;
  std::string error;
  std::vector<Request> per_rank;   // ALLTOALL / EXCHANGE: every rank's own extras / blob
  int tl_pid = -1;
};
struct Response {
  Request req;
  std::string error;
  int code = 0;
  uint64_t bits = 0;
  std::vector<Request> per_rank;
  int last_joined = -1;
  int tl_pid = -1;
};
struct CacheItem { uint32_t id; Request sig; uint64_t last_used; };

struct Cuda {  // resolved at run time from the cudart already in the process (torch's); never linked
  int (*SetDevice)(int) = nullptr;
  int (*StreamCreateWithPriority)(void**, unsigned, int) = nullptr;
  int (*StreamWaitEvent)(void*, void*, unsigned) = nullptr;
  int (*EventCreateWithFlags)(void**, unsigned) = nullptr;
  int (*EventRecord)(void*, void*) = nullptr;
  int (*EventQuery)(void*) = nullptr;
  int (*EventDestroy)(void*) = nullptr;
  int (*MemcpyAsync)(void*, const void*, size_t, int, void*) = nullptr;
  int (*MemsetAsync)(void*, int, size_t, void*) = nullptr;
  int (*Malloc)(void**, size_t) = nullptr;
  int (*Free)(void*) = nullptr;
  int (*StreamSynchronize)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string* err) {
    void* h = nullptr;
    for (const char* n : {"libcudart.so.12", "libcudart.so.13", "libcudart.so"}) {
      h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
      if (h) break;
    }
    if (!h) for (const char* n : {"libcudart.so.12", "libcudart.so"}) { h = dlopen(n, RTLD_NOW); if (h) break; }
    if (!h) { *err = "hvdcore: no CUDA runtime library in the process (import torch first)"; return false; }
#define HVD_SYM(field, sym) field = (decltype(field))dlsym(h, sym); if (!field) { *err = std::string("hvdcore: missing ") + sym; return false; }
    HVD_SYM(SetDevice, "cudaSetDevice") HVD_SYM(StreamCreateWithPriority, "cudaStreamCreateWithPriority")
    HVD_SYM(StreamWaitEvent, "cudaStreamWaitEvent") HVD_SYM(EventCreateWithFlags, "cudaEventCreateWithFlags")
    HVD_SYM(EventRecord, "cudaEventRecord") HVD_SYM(EventQuery, "cudaEventQuery") HVD_SYM(EventDestroy, "cudaEventDestroy")
    HVD_SYM(MemcpyAsync, "cudaMemcpyAsync") HVD_SYM(MemsetAsync, "cudaMemsetAsync") HVD_SYM(Malloc, "cudaMalloc")
    HVD_SYM(Free, "cudaFree") HVD_SYM(StreamSynchronize, "cudaStreamSynchronize") HVD_SYM(GetErrorString, "cudaGetErrorString")
#undef HVD_SYM
    return true;
  }
};
struct GpuInflight { void* ev; std::vector<int> handles; std::vector<int> tl_pids; std::string what; };

struct Timeline {
  FILE* f = nullptr;
  bool first = true;
  uint64_t t0 = 0;
  std::mutex mu;
  int next_pid = 1;
  std::unordered_map<std::string, int> pids;
  bool on() const { return f != nullptr; }
  void open(const char* path) {
    std::lock_guard<std::mutex> g(mu);
    if (f) return;
    f = fopen(path, "w");
    if (!f) return;
    fputs("[\n", f);
    first = true; t0 = now_ns(); pids.clear(); next_pid = 1;
  }
  void close() {
    std::lock_guard<std::mutex> g(mu);
    if (!f) return;
    fputs("\n]\n", f);
    fclose(f);
    f = nullptr;
  }
  static std::string esc(const std::string& s) { std::string o; for (char c : s) { if (c == '"' || c == '\\') o.push_back('\\'); o.push_back(c < 0x20 ? ' ' : c); } return o; }
  void emit(const std::string& body) {  // caller holds mu
    if (!first) fputs(",\n", f);
    first = false;
    fputs(body.c_str(), f);
  }
  int pid_for(const std::string& name) {
    std::lock_guard<std::mutex> g(mu);
    if (!f) return -1;
    auto it = pids.find(name);
    if (it != pids.end()) return it->second;
    const int pid = next_pid++;
    pids[name] = pid;
    emit("{\"name\": \"process_name\", \"ph\": \"M\", \"pid\": " + std::to_string(pid) + ", \"args\": {\"name\": \"" + esc(name) + "\"}}");
    emit("{\"name\": \"process_sort_index\", \"ph\": \"M\", \"pid\": " + std::to_string(pid) + ", \"args\": {\"sort_index\": " + std::to_string(pid) + "}}");
    return pid;
  }
  void ev(int pid, char ph, const std::string& name, const std::string& args = "") {
    if (pid < 0) return;
    std::lock_guard<std::mutex> g(mu);
    if (!f) return;
    const double ts = (double)(now_ns() - t0) / 1000.0;
    char head[160];
    snprintf(head, sizeof(head), "{\"ph\": \"%c\", \"pid\": %d, \"tid\": 0, \"ts\": %.3f", ph, pid, ts);
    std::string b = head;
    if (!name.empty()) b += ", \"name\": \"" + esc(name) + "\"";
    if (ph == 'i') b += ", \"s\": \"p\"";
    if (!args.empty()) b += ", \"args\": {" + args + "}";
    b += "}";
    emit(b);
  }
};

struct Stats {
  std::atomic<uint64_t> cycles{0}, tensors{0}, groups{0}, fused_tensors{0}, bytes{0}, cache_hits{0}, cache_misses{0},
      stall_warnings{0}, errors{0}, negotiation_bytes{0};
};

struct Engine {
  int rank = 0, world = 1;
  Rendezvous rv;
  std::thread thread;
  std::mutex mu;                       // queue, handles, inflight_names
  std::condition_variable queue_cv, done_cv;
  std::deque<LocalOp> queue;
  std::unordered_map<int, Handle> handles;
  std::unordered_set<std::string> inflight_names;
  int next_handle = 1;
  uint64_t noname[8] = {};
  std::atomic<bool> shutdown_requested{false}, stopped{false};
  std::string stop_reason;
  int stop_code = HVD_ERR_SHUTDOWN;

  // tunables
  std::atomic<double> cycle_ms{1.0}, stall_check_s{60.0}, stall_shutdown_s{0.0};
  std::atomic<int64_t> fusion_threshold{64ll << 20};
  int cache_capacity = 1024;
  bool stall_check = true, mark_cycles = false;
  int timeout_ms = 600000;

  // engine-thread state
  std::map<std::string, LocalOp> pending;            // submitted locally, waiting for a response
  std::map<std::string, Entry> table;
  uint64_t next_seq = 0, cycle_no = 0;
  uint64_t joined = 0;                               // bit r: rank r called join()
  int last_joined = -1;
  std::unordered_map<std::string, CacheItem> cache;  // name -> item
  std::unordered_map<uint32_t, std::string> cache_by_id;
  uint32_t next_cache_id = 1;
  std::vector<unsigned char> fusion_host;
  std::vector<float> acc32;
  // host data plane: per rank a {data, result} pair of `box` bytes in a segment of its own (unlinked as soon as every rank has
  // mapped it); when it cannot be created the 64 KiB mailboxes of the rendezvous segment carry the data instead
  unsigned char* seg = nullptr;
  size_t seg_bytes = 0, box = 0;
  uint64_t last_stall_scan = 0;
  // HOROVOD_AUTOTUNE (rank 0 decides, see autotune_step): grid search over {cycle time} x {fusion threshold}, scored by
  // reduced bytes per second over windows of busy cycles
  bool autotune = false, autotune_done = false;
  int at_warmup = 3, at_steps = 10, at_idx = -1, at_busy = 0, at_best = -1, at_tail = 3;   // at_tail: cycles that still announce the final choice
  bool at_seen = false;                                                                    // (other ranks) rank 0 has been announcing
  uint64_t at_bytes0 = 0, at_t0 = 0, at_samples = 0;
  double at_best_score = 0;
  std::vector<std::pair<double, int64_t>> at_grid;
  FILE* at_log = nullptr;

  // gpu
  bool has_gpu = false;
  unsigned char* ada_seg = nullptr;   // Adasum scratch: world slots of ada_slot bytes in a shared segment of its own
  size_t ada_slot = 0, ada_total = 0;
  hvdcore_gpu_t gpu{};
  Cuda cu;
  void* stream = nullptr;
  void* fusion_dev = nullptr;
  size_t fusion_dev_bytes = 0;
  std::vector<GpuInflight> gpu_inflight;

  Timeline tl;
  Stats st;
};

Engine* g = nullptr;
std::mutex g_mu;           // guards g itself (init / shutdown)
int g_generation = 0;

double env_d(const char* n, double d) { const char* v = getenv(n); return (v && *v) ? atof(v) : d; }

void complete(Engine* e, int handle, int status, const std::string& err) {
  std::lock_guard<std::mutex> lk(e->mu);
  auto it = e->handles.find(handle);
  if (it == e->handles.end()) return;
  it->second.done = true;
  it->second.status = status;
  it->second.err = err;
  e->done_cv.notify_all();
}
void finish_local(Engine* e, const std::string& name, int status, const std::string& err) {
  auto it = e->pending.find(name);
  if (it == e->pending.end()) return;
  const int h = it->second.handle;
  e->pending.erase(it);
  { std::lock_guard<std::mutex> lk(e->mu); e->inflight_names.erase(name); }
  if (status < 0) e->st.errors++;
  complete(e, h, status, err);
}

// -------------------------------------------------------------------------------------------------- transport --
// One negotiation exchange: a small fixed-size allgather carries {length, flags, first bytes}; longer messages take
// extra rounds whose chunk size every rank derives from the same gathered lengths.
int exchange(Engine* e, const std::string& mine, uint32_t flags, std::vector<std::string>* all, uint32_t* all_flags, int64_t* fusion) {
  std::string err;
  const int W = e->world;
  // fusion threshold: the smallest one wins, so groups agree — unless rank 0 is autotuning, then its threshold and cycle time are adopted
  struct Hdr { uint32_t len, flags; int64_t fusion; double cycle_ms; };
  unsigned char small[sizeof(Hdr) + kInline];
  memset(small, 0, sizeof(small));
  Hdr h{(uint32_t)mine.size(), flags, e->fusion_threshold.load(), e->cycle_ms.load()};
  *fusion = h.fusion;
  memcpy(small, &h, sizeof(h));
  memcpy(small + sizeof(h), mine.data(), std::min(mine.size(), kInline));
  // The small exchange costs ONE barrier: it lives in the last kTail bytes of the mailbox, in two slots used alternately. A rank
  // may write slot p^1 (next cycle) while a slow peer still reads slot p; it cannot get two cycles ahead because the next cycle's
  // barrier needs that peer. Everything else that goes through the mailbox (long messages, data without the box segment) stays
  // below the tail (kBody).
  std::vector<unsigned char> got(sizeof(small) * W);
  {
    static_assert(sizeof(small) <= kTail / 2, "exchange slot too small");
    b200mpi::RvHeader* H = e->rv.header();
    const size_t slot = kBody + (size_t)(e->cycle_no & 1u) * (kTail / 2);
    memcpy(H->slot[e->rank].mailbox + slot, small, sizeof(small));
    if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    for (int r = 0; r < W; r++) memcpy(got.data() + (size_t)r * sizeof(small), H->slot[r].mailbox + slot, sizeof(small));
  }
  all->assign(W, std::string());
  *all_flags = 0;
  int64_t tuned_fusion = -1;
  double tuned_cycle = 0;
  size_t max_len = 0;
  std::vector<size_t> lens(W);
  for (int r = 0; r < W; r++) {
    Hdr hr;
    memcpy(&hr, got.data() + (size_t)r * sizeof(small), sizeof(hr));
    *all_flags |= hr.flags;
    *fusion = std::min(*fusion, hr.fusion);
    if (r == 0 && (hr.flags & kFlagAutotune)) { tuned_fusion = hr.fusion; tuned_cycle = hr.cycle_ms; }
    lens[r] = hr.len;
    max_len = std::max(max_len, (size_t)hr.len);
    (*all)[r].assign((const char*)got.data() + (size_t)r * sizeof(small) + sizeof(hr), std::min((size_t)hr.len, kInline));
  }
  if (tuned_fusion >= 0) {   // every rank runs with the values rank 0 is currently trying (or has settled on)
    *fusion = tuned_fusion;
    if (e->rank != 0) { e->fusion_threshold = tuned_fusion; e->cycle_ms = tuned_cycle; }
  }
  size_t off = kInline;
  std::vector<unsigned char> big, mine_chunk;
  while (off < max_len) {
    const size_t chunk = std::min(kBody, max_len - off);
    mine_chunk.assign(chunk, 0);
    if (mine.size() > off) memcpy(mine_chunk.data(), mine.data() + off, std::min(chunk, mine.size() - off));
    big.resize(chunk * W);
    if (e->rv.allgather(mine_chunk.data(), big.data(), chunk, e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    for (int r = 0; r < W; r++)
      if (lens[r] > off) (*all)[r].append((const char*)big.data() + (size_t)r * chunk, std::min(chunk, lens[r] - off));
    off += chunk;
  }
  size_t total = 0;
  for (auto& s : *all) total += s.size();
  e->st.negotiation_bytes += total;
  return 0;
}

inline unsigned char* box_data(Engine* e, int r) { return e->seg ? e->seg + (size_t)r * 2 * e->box : e->rv.header()->slot[r].mailbox; }
inline unsigned char* box_result(Engine* e, int r) { return e->seg + (size_t)r * 2 * e->box + e->box; }
inline size_t box_bytes(Engine* e) { return e->seg ? e->box : kBody; }

// acc = fold over ranks (in rank order) of n elements found at offset `off` of every rank's data box; result to `out`
void fold_boxes(Engine* e, size_t off, size_t n, int dt, int op, void* out) {
  const int W = e->world;
  if (dt == HVD_F16 || dt == HVD_BF16) {
    if (e->acc32.size() < n) e->acc32.resize(n);
    float* a = e->acc32.data();
    const uint16_t* s0 = (const uint16_t*)(box_data(e, 0) + off);
    if (dt == HVD_F16) for (size_t i = 0; i < n; i++) a[i] = f16_to_f(s0[i]); else for (size_t i = 0; i < n; i++) a[i] = bf16_to_f(s0[i]);
    for (int r = 1; r < W; r++) {
      if (dt == HVD_F16) fold16<f16_to_f>(a, box_data(e, r) + off, n, op); else fold16<bf16_to_f>(a, box_data(e, r) + off, n, op);
    }
    uint16_t* o = (uint16_t*)out;
    if (dt == HVD_F16) for (size_t i = 0; i < n; i++) o[i] = f_to_f16(a[i]); else for (size_t i = 0; i < n; i++) o[i] = f_to_bf16(a[i]);
  } else {
    memcpy(out, box_data(e, 0) + off, n * esize(dt));
    for (int r = 1; r < W; r++) fold(out, box_data(e, r) + off, n, dt, op);
  }
}

// In-place allreduce of `count` elements. Every chunk: publish it; barrier; with the data segment each rank folds ONE slice
// of the chunk over all ranks (rank order: bit-identical everywhere) into its result box, barrier, everyone gathers the
// slices — 1/W of the folding per rank and two barriers per `box` bytes; the next chunk's publish needs no extra barrier
// because nobody reads data boxes while results are gathered. Without the segment (64 KiB mailboxes) every rank folds the
// whole chunk itself. 16-bit floats accumulate in fp32.
int host_allreduce(Engine* e, void* buf, int64_t count, int dt, int op) {
  std::string err;
  const int W = e->world, me = e->rank;
  const size_t es = esize(dt);
  const size_t per = box_bytes(e) / 8 * 8 / es;   // elements per chunk
  unsigned char* p = (unsigned char*)buf;
  for (int64_t done = 0; done < count; done += (int64_t)per) {
    const size_t n = (size_t)std::min<int64_t>((int64_t)per, count - done);
    unsigned char* o = p + (size_t)done * es;
    memcpy(box_data(e, me), o, n * es);
    if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    if (e->seg) {
      const size_t lo = n * (size_t)me / (size_t)W, hi = n * (size_t)(me + 1) / (size_t)W;
      if (hi > lo) fold_boxes(e, lo * es, hi - lo, dt, op, box_result(e, me) + lo * es);
      if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
      for (int r = 0; r < W; r++) {
        const size_t a = n * (size_t)r / (size_t)W, b = n * (size_t)(r + 1) / (size_t)W;
        if (b > a) memcpy(o + a * es, box_result(e, r) + a * es, (b - a) * es);
      }
    } else {
      fold_boxes(e, 0, n, dt, op, o);
      if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    }
  }
  return 0;
}

// ---- Adasum (Horovod's op=hvd.Adasum; the --use-adasum flag of the reference's MNIST example, tensorflow_mnist.py:31-32,133) ----
// adasum(a, b) = (1 - a.b / (2 |a|^2)) a + (1 - a.b / (2 |b|^2)) b, folded as a binary tree over the ranks: orthogonal gradients
// add, parallel gradients average. Horovod does vector-halving / distance-doubling over MPI; on one box every rank's vector
// goes into a slot of a shared scratch segment and ALL ranks work on every pair of a tree level, each on its own 1/W slice:
// partial dot products (double) are exchanged through the rendezvous allgather and summed in rank order - identical
// coefficients everywhere - then each rank combines its slice in place. O(n) work per rank, one allgather per level,
// the result (slot 0) is bit-identical on all ranks.
template <typename T> struct AdaPair { T* a; const T* b; };

template <typename T>
int adasum_levels(Engine* e, size_t n) {
  std::string err;
  const int W = e->world, me = e->rank;
  const size_t lo = n * (size_t)me / (size_t)W, hi = n * (size_t)(me + 1) / (size_t)W;
  std::vector<int> active((size_t)W);
  for (int r = 0; r < W; r++) active[(size_t)r] = r;
  constexpr size_t kBatch = 32;   // pairs per allgather round (768 bytes of the rendezvous mailbox)
  std::vector<double> all((size_t)W * 3 * kBatch);
  while (active.size() > 1) {
    const size_t np = active.size() / 2;
    for (size_t p0 = 0; p0 < np; p0 += kBatch) {
      const size_t nb = std::min(kBatch, np - p0);
      double mine[3 * kBatch];
      for (size_t k = 0; k < nb; k++) {
        const T* a = (const T*)(e->ada_seg + (size_t)active[2 * (p0 + k)] * e->ada_slot);
        const T* b = (const T*)(e->ada_seg + (size_t)active[2 * (p0 + k) + 1] * e->ada_slot);
        double dot = 0, na = 0, nbn = 0;
        for (size_t i = lo; i < hi; i++) { const double x = (double)a[i], y = (double)b[i]; dot += x * y; na += x * x; nbn += y * y; }
        mine[3 * k] = dot; mine[3 * k + 1] = na; mine[3 * k + 2] = nbn;
      }
      if (e->rv.allgather(mine, all.data(), nb * 3 * sizeof(double), e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
      for (size_t k = 0; k < nb; k++) {
        double dot = 0, na = 0, nbn = 0;
        for (int r = 0; r < W; r++) { const double* q = all.data() + ((size_t)r * nb + k) * 3; dot += q[0]; na += q[1]; nbn += q[2]; }
        const double ca = na > 0 ? 1.0 - dot / (2.0 * na) : 1.0, cb = nbn > 0 ? 1.0 - dot / (2.0 * nbn) : 1.0;
        T* a = (T*)(e->ada_seg + (size_t)active[2 * (p0 + k)] * e->ada_slot);
        const T* b = (const T*)(e->ada_seg + (size_t)active[2 * (p0 + k) + 1] * e->ada_slot);
        for (size_t i = lo; i < hi; i++) a[i] = (T)(ca * (double)a[i] + cb * (double)b[i]);
      }
    }
    std::vector<int> next;   // (no barrier between levels: a rank only ever touches its own slice of every slot)
    for (size_t k = 0; k < np; k++) next.push_back(active[2 * k]);
    if (active.size() % 2) next.push_back(active.back());
    active.swap(next);
  }
  if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);   // every slice of the root slot is final
  return 0;
}

// In-place Adasum of `count` elements of dtype f32 / f64 / f16 / bf16 (16-bit floats are combined in fp32). `in` may be null
// (a joined rank contributes zeros, which Adasum treats as neutral).
int host_adasum(Engine* e, const void* in, void* out, int64_t count, int dt) {
  if (dt != HVD_F32 && dt != HVD_F64 && dt != HVD_F16 && dt != HVD_BF16) return fail(HVD_ERR_UNSUPPORTED, "Adasum needs a floating-point tensor");
  if (count == 0) return 0;
  std::string err;
  const size_t n = (size_t)count, ws = dt == HVD_F64 ? 8 : 4;
  const size_t need = (n * ws + 63) / 64 * 64;
  if (e->ada_slot < need) {   // collective decision: every rank sees the same count
    if (e->ada_seg) { Rendezvous::close_boxes(e->ada_seg, e->ada_total); e->ada_seg = nullptr; e->ada_slot = 0; }
    const size_t box = std::max(need, (size_t)1 << 20) / 2;
    e->ada_seg = e->rv.open_boxes(box, e->timeout_ms, &e->ada_total);
    if (!e->ada_seg) return fail(HVD_ERR_TRANSPORT, "Adasum: could not create the shared scratch segment (" + std::to_string((size_t)e->world * 2 * box) + " bytes in /dev/shm)");
    e->ada_slot = 2 * box;
  }
  unsigned char* mine = e->ada_seg + (size_t)e->rank * e->ada_slot;
  if (!in) memset(mine, 0, n * ws);
  else if (dt == HVD_F32 || dt == HVD_F64) memcpy(mine, in, n * ws);
  else {
    float* f = (float*)mine; const uint16_t* h = (const uint16_t*)in;
    if (dt == HVD_F16) for (size_t i = 0; i < n; i++) f[i] = f16_to_f(h[i]); else for (size_t i = 0; i < n; i++) f[i] = bf16_to_f(h[i]);
  }
  if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
  const int rc = dt == HVD_F64 ? adasum_levels<double>(e, n) : adasum_levels<float>(e, n);
  if (rc) return rc;
  if (out) {
    const unsigned char* res = e->ada_seg;   // slot 0 holds the root of the tree
    if (dt == HVD_F32 || dt == HVD_F64) memcpy(out, res, n * ws);
    else {
      const float* f = (const float*)res; uint16_t* h = (uint16_t*)out;
      if (dt == HVD_F16) for (size_t i = 0; i < n; i++) h[i] = f_to_f16(f[i]); else for (size_t i = 0; i < n; i++) h[i] = f_to_bf16(f[i]);
    }
  }
  if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);     // everyone has read slot 0 before it is reused
  return 0;
}

int host_bcast(Engine* e, void* buf, size_t bytes, int root) {
  std::string err;
  const size_t B = box_bytes(e);
  for (size_t off = 0; off < bytes; off += B) {
    const size_t n = std::min(B, bytes - off);
    if (e->rank == root) memcpy(box_data(e, root), (char*)buf + off, n);
    if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    if (e->rank != root) memcpy((char*)buf + off, box_data(e, root), n);
    if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
  }
  return 0;
}

int host_allgatherv(Engine* e, const void* in, void* out, const std::vector<int64_t>& counts) {
  std::string err;
  const int W = e->world;
  const int64_t B = (int64_t)box_bytes(e);
  int64_t mx = 0;
  std::vector<int64_t> displ(W, 0);
  for (int r = 0; r < W; r++) { mx = std::max(mx, counts[r]); if (r) displ[r] = displ[r - 1] + counts[r - 1]; }
  for (int64_t off = 0; off < mx; off += B) {
    const int64_t mine = std::min<int64_t>(B, counts[e->rank] - off);
    if (mine > 0) memcpy(box_data(e, e->rank), (const char*)in + off, (size_t)mine);
    if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    for (int r = 0; r < W; r++) {
      const int64_t n = std::min<int64_t>(B, counts[r] - off);
      if (n > 0) memcpy((char*)out + displ[r] + off, box_data(e, r), (size_t)n);
    }
    if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
  }
  return 0;
}

// send[s][d] = bytes rank s sends to rank d (known to every rank from the negotiation)
int host_alltoallv(Engine* e, const void* in, void* out, const std::vector<std::vector<int64_t>>& send) {
  std::string err;
  const int W = e->world, me = e->rank;
  const int64_t B = (int64_t)box_bytes(e);
  std::vector<int64_t> sd(W, 0), rdp(W, 0);
  for (int r = 1; r < W; r++) { sd[r] = sd[r - 1] + send[me][r - 1]; rdp[r] = rdp[r - 1] + send[r - 1][me]; }
  if (send[me][me] > 0) memcpy((char*)out + rdp[me], (const char*)in + sd[me], (size_t)send[me][me]);
  for (int step = 1; step < W; step++) {
    const int dst = (me + step) % W, src = (me - step + W) % W;
    int64_t mx = 0;
    for (int s = 0; s < W; s++) mx = std::max(mx, send[s][(s + step) % W]);
    for (int64_t off = 0; off < mx; off += B) {
      const int64_t ns = std::min<int64_t>(B, send[me][dst] - off);
      if (ns > 0) memcpy(box_data(e, me), (const char*)in + sd[dst] + off, (size_t)ns);
      if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
      const int64_t nr = std::min<int64_t>(B, send[src][me] - off);
      if (nr > 0) memcpy((char*)out + rdp[src] + off, box_data(e, src), (size_t)nr);
      if (e->rv.barrier(e->timeout_ms, &err)) return fail(HVD_ERR_TRANSPORT, err);
    }
  }
  return 0;
}

// B200MPI_HVD_MAILBOX_KB (default 256, 0 = mailboxes only): box size of the engine's data segment (Rendezvous::open_boxes)
void open_data_segment(Engine* e) {
  const double kb = env_d("B200MPI_HVD_MAILBOX_KB", 256.0);
  if (kb <= 0) return;
  const size_t box = (size_t)std::max(64.0, kb) * 1024;
  e->seg = e->rv.open_boxes(box, e->timeout_ms, &e->seg_bytes);
  e->box = e->seg ? box : 0;
}

// ------------------------------------------------------------------------------------------------ coordinator --
uint64_t kAllBits(int world) { return world >= 64 ? ~0ull : ((1ull << world) - 1); }

std::string describe(const Request& r) {
  char b[256];
  snprintf(b, sizeof(b), "%s dtype=%d count=%lld op=%d root=%d device=%s", kOpName[r.op], r.dtype, (long long)r.count, r.redop, r.root,
           r.devkind ? "gpu" : "cpu");
  return b;
}

void cache_touch(Engine* e, const Request& r) {
  if (e->cache_capacity <= 0 || !cacheable(r)) return;
  auto it = e->cache.find(r.name);
  if (it != e->cache.end()) {
    if (same_sig(it->second.sig, r)) { it->second.last_used = e->cycle_no; return; }
    e->cache_by_id.erase(it->second.id);
    e->cache.erase(it);
  }
  if ((int)e->cache.size() >= e->cache_capacity) {  // deterministic LRU: oldest use, then smallest id
    auto victim = e->cache.end();
    for (auto i = e->cache.begin(); i != e->cache.end(); ++i)
      if (victim == e->cache.end() || i->second.last_used < victim->second.last_used ||
          (i->second.last_used == victim->second.last_used && i->second.id < victim->second.id)) victim = i;
    e->cache_by_id.erase(victim->second.id);
    e->cache.erase(victim);
  }
  const uint32_t id = e->next_cache_id++;
  e->cache[r.name] = CacheItem{id, r, e->cycle_no};
  e->cache_by_id[id] = r.name;
}

// Folds every rank's message of this cycle into the table, then moves complete entries (in first-seen order) to `out`.
void coordinate(Engine* e, const std::vector<std::string>& msgs, std::vector<Response>* out) {
  const int W = e->world;
  const uint64_t now = now_ns();
  std::vector<std::string> invalidate;
  for (int r = 0; r < W; r++) {
    Reader rd(msgs[r]);
    while (rd.ok && !rd.done()) {
      Request q;
      const uint8_t kind = rd.get<uint8_t>();
      if (!rd.ok) break;
      if (kind == 1) {
        const uint32_t id = rd.get<uint32_t>();
        auto ci = e->cache_by_id.find(id);
        if (!rd.ok || ci == e->cache_by_id.end()) { fprintf(stderr, "[hvdcore rank %d] unknown cache id %u from rank %d\n", e->rank, id, r); break; }
        q = e->cache[ci->second].sig;
        if (r == e->rank) e->st.cache_hits++;
      } else {
        if (!decode_full(rd, &q)) { fprintf(stderr, "[hvdcore rank %d] undecodable request from rank %d\n", e->rank, r); break; }
        if (r == e->rank && cacheable(q)) e->st.cache_misses++;
        if (e->cache.count(q.name) && !same_sig(e->cache[q.name].sig, q)) invalidate.push_back(q.name);
      }
      if (q.op == HVD_JOIN) {
        e->joined |= 1ull << r;
        e->last_joined = r;
      }
      auto it = e->table.find(q.name);
      if (it == e->table.end()) {
        Entry en;
        en.first = q;
        en.seq = e->next_seq++;
        en.t_first = now;
        if (q.op == HVD_ALLTOALL || q.op == HVD_EXCHANGE || q.op == HVD_ALLGATHER) en.per_rank.resize(W);
        if (e->tl.on()) {
          en.tl_pid = e->tl.pid_for(q.name);
          e->tl.ev(en.tl_pid, 'B', std::string("NEGOTIATE_") + kOpName[q.op]);
        }
        it = e->table.emplace(q.name, std::move(en)).first;
      }
      Entry& en = it->second;
      if (en.bits & (1ull << r)) {
        if (en.error.empty()) en.error = "rank " + std::to_string(r) + " submitted '" + q.name + "' twice before it completed";
        continue;
      }
      en.bits |= 1ull << r;
      if (e->tl.on()) e->tl.ev(en.tl_pid, 'i', std::to_string(r));
      const Request& f = en.first;
      if (en.error.empty()) {
        bool ok = q.op == f.op && q.dtype == f.dtype && q.devkind == f.devkind;
        if (ok && (q.op == HVD_ALLREDUCE)) ok = q.count == f.count && q.redop == f.redop && memcmp(&q.pre, &f.pre, 8) == 0 && memcmp(&q.post, &f.post, 8) == 0;
        if (ok && q.op == HVD_BROADCAST) ok = q.count == f.count && q.root == f.root;
        if (ok && q.op == HVD_ALLGATHER) ok = q.extra == f.extra;
        if (ok && q.op == HVD_EXCHANGE) ok = q.count == f.count;
        if (!ok) en.error = "mismatched submissions for '" + q.name + "': one rank has [" + describe(f) + "], rank " + std::to_string(r) + " has [" + describe(q) + "]";
      }
      if (!en.per_rank.empty()) en.per_rank[r] = q;
    }
  }
  for (auto& n : invalidate) {
    auto it = e->cache.find(n);
    if (it != e->cache.end()) { e->cache_by_id.erase(it->second.id); e->cache.erase(it); }
  }
  // completion: every rank either submitted the tensor or has joined
  std::vector<Entry*> ready;
  const uint64_t all = kAllBits(W);
  for (auto& kv : e->table) {
    Entry& en = kv.second;
    if (en.first.op == HVD_JOIN) { if (e->joined == all) ready.push_back(&en); continue; }
    if ((en.bits | e->joined) == all) ready.push_back(&en);
  }
  std::sort(ready.begin(), ready.end(), [](Entry* a, Entry* b) { return a->seq < b->seq; });
  bool join_done = false;
  for (Entry* en : ready) {
    Response rs;
    rs.req = en->first;
    rs.error = en->error;
    rs.bits = en->bits;
    rs.per_rank = std::move(en->per_rank);
    rs.tl_pid = en->tl_pid;
    if (rs.error.empty() && en->bits != all && en->first.op != HVD_ALLREDUCE && en->first.op != HVD_JOIN)
      rs.error = std::string(kOpName[en->first.op]) + " '" + en->first.name + "' cannot complete: some ranks have already joined (only allreduce is supported with join)";
    if (rs.error.empty() && en->first.op == HVD_ALLGATHER) {
      for (int s = 0; s < W; s++)
        if ((int)rs.req.extra.size() != W || rs.req.extra[s] != rs.per_rank[s].count * (int64_t)esize(rs.req.dtype)) {
          rs.error = "allgather '" + en->first.name + "': the byte counts do not describe rank " + std::to_string(s) + "'s input";
          break;
        }
    }
    if (rs.error.empty() && en->first.op == HVD_ALLTOALL) {
      for (int s = 0; s < W && rs.error.empty(); s++) {
        if ((int)rs.per_rank[s].extra.size() != 2 * W) { rs.error = "alltoall '" + en->first.name + "': rank " + std::to_string(s) + " passed a malformed split vector"; break; }
        for (int d = 0; d < W; d++)
          if ((int)rs.per_rank[d].extra.size() == 2 * W && rs.per_rank[s].extra[d] != rs.per_rank[d].extra[W + s]) {
            rs.error = "alltoall '" + en->first.name + "': rank " + std::to_string(s) + " sends " + std::to_string(rs.per_rank[s].extra[d]) + " bytes to rank " +
                       std::to_string(d) + ", which expects " + std::to_string(rs.per_rank[d].extra[W + s]);
            break;
          }
      }
    }
    if (!rs.error.empty()) rs.code = HVD_ERR_MISMATCH;
    if (en->first.op == HVD_JOIN) { rs.last_joined = e->last_joined; join_done = true; }
    if (rs.error.empty()) cache_touch(e, rs.req);
    if (e->tl.on()) e->tl.ev(rs.tl_pid, 'E', "");
    e->st.tensors++;
    out->push_back(std::move(rs));
  }
  for (auto& rs : *out) e->table.erase(rs.req.name);
  if (join_done) { e->joined = 0; e->last_joined = -1; }
}

// --------------------------------------------------------------------------------------------------- execution --
std::string gpu_err(Engine* e, const char* what, int rc) {
  typedef const char* (*le_t)(void);
  std::string s = std::string(what) + " failed (" + std::to_string(rc) + ")";
  if (e->gpu.last_error) s += std::string(": ") + ((le_t)e->gpu.last_error)();
  return s;
}

int b200_dtype(int dt) { return dt == HVD_F32 ? 0 : dt == HVD_BF16 ? 1 : dt == HVD_F16 ? 2 : -1; }
int b200_op(int op) { return op == HVD_SUM ? 0 : op == HVD_MAX ? 1 : op == HVD_MIN ? 2 : -1; }

void gpu_poll(Engine* e, bool block) {
  for (size_t i = 0; i < e->gpu_inflight.size();) {
    GpuInflight& f = e->gpu_inflight[i];
    int q = e->cu.EventQuery(f.ev);
    if (q != 0 && block) { e->cu.StreamSynchronize(e->stream); q = e->cu.EventQuery(f.ev); }
    if (q == 600 /* cudaErrorNotReady */) { i++; continue; }
    const int status = q == 0 ? 0 : HVD_ERR_TRANSPORT;
    const std::string err = q == 0 ? "" : std::string("CUDA error after ") + f.what + ": " + e->cu.GetErrorString(q);
    for (int pid : f.tl_pids) { e->tl.ev(pid, 'E', ""); e->tl.ev(pid, 'E', ""); }
    for (int h : f.handles) complete(e, h, status, err);
    e->cu.EventDestroy(f.ev);
    e->gpu_inflight.erase(e->gpu_inflight.begin() + (long)i);
  }
}

void gpu_track(Engine* e, std::vector<int> handles, std::vector<int> pids, const std::string& what) {
  void* ev = nullptr;
  e->cu.EventCreateWithFlags(&ev, 2 /* cudaEventDisableTiming */);
  e->cu.EventRecord(ev, e->stream);
  e->gpu_inflight.push_back(GpuInflight{ev, std::move(handles), std::move(pids), what});
}

// One fused allreduce: `grp` are responses of identical (dtype, op, scales, device kind).
void run_allreduce_group(Engine* e, std::vector<Response*>& grp) {
  const Request& q0 = grp[0]->req;
  const size_t es = esize(q0.dtype);
  int64_t total = 0;
  for (auto* r : grp) total += r->req.count;
  e->st.groups++;
  e->st.fused_tensors += grp.size();
  e->st.bytes += (uint64_t)total * es;
  std::vector<LocalOp*> ops(grp.size(), nullptr);
  for (size_t i = 0; i < grp.size(); i++) {
    auto it = e->pending.find(grp[i]->req.name);
    if (it != e->pending.end()) ops[i] = &it->second;
  }
  const bool tl = e->tl.on();
  for (auto* r : grp) if (tl) e->tl.ev(r->tl_pid, 'B', "ALLREDUCE", "\"fused_with\": " + std::to_string(grp.size() - 1));
  if (q0.devkind == 0) {
    // ---- host path ----
    int rc = 0;
    if (q0.redop == HVD_ADASUM) {   // never fused (execute()): the coefficients are per tensor
      LocalOp* o = ops[0];
      std::vector<unsigned char> tmp;
      const void* src = o ? o->in : nullptr;
      if (o && q0.pre != 1.0) { tmp.assign((const unsigned char*)o->in, (const unsigned char*)o->in + (size_t)total * es); scale_buf(tmp.data(), (size_t)total, q0.dtype, q0.pre); src = tmp.data(); }
      if (tl) e->tl.ev(grp[0]->tl_pid, 'B', "SHM_ADASUM");
      rc = host_adasum(e, src, o ? o->out : nullptr, total, q0.dtype);
      if (tl) e->tl.ev(grp[0]->tl_pid, 'E', "");
      if (!rc && o) scale_buf(o->out, (size_t)total, q0.dtype, q0.post);
    } else if (grp.size() == 1 && ops[0]) {
      LocalOp* o = ops[0];
      if (o->out != o->in) memcpy(o->out, o->in, (size_t)total * es);
      scale_buf(o->out, (size_t)total, q0.dtype, q0.pre);
      if (tl) e->tl.ev(grp[0]->tl_pid, 'B', "SHM_ALLREDUCE");
      rc = host_allreduce(e, o->out, total, q0.dtype, q0.redop);
      if (tl) e->tl.ev(grp[0]->tl_pid, 'E', "");
      if (!rc) scale_buf(o->out, (size_t)total, q0.dtype, q0.post);
    } else {
      if (e->fusion_host.size() < (size_t)total * es) e->fusion_host.resize((size_t)total * es);
      unsigned char* fb = e->fusion_host.data();
      for (auto* r : grp) if (tl) e->tl.ev(r->tl_pid, 'B', "MEMCPY_IN_FUSION_BUFFER");
      size_t off = 0;
      for (size_t i = 0; i < grp.size(); i++) {
        const size_t nb = (size_t)grp[i]->req.count * es;
        if (ops[i]) memcpy(fb + off, ops[i]->in, nb); else memset(fb + off, 0, nb);   // a joined rank contributes zeros
        off += nb;
      }
      scale_buf(fb, (size_t)total, q0.dtype, q0.pre);
      for (auto* r : grp) if (tl) { e->tl.ev(r->tl_pid, 'E', ""); e->tl.ev(r->tl_pid, 'B', "SHM_ALLREDUCE"); }
      rc = host_allreduce(e, fb, total, q0.dtype, q0.redop);
      for (auto* r : grp) if (tl) { e->tl.ev(r->tl_pid, 'E', ""); e->tl.ev(r->tl_pid, 'B', "MEMCPY_OUT_FUSION_BUFFER"); }
      if (!rc) {
        scale_buf(fb, (size_t)total, q0.dtype, q0.post);
        off = 0;
        for (size_t i = 0; i < grp.size(); i++) {
          const size_t nb = (size_t)grp[i]->req.count * es;
          if (ops[i]) memcpy(ops[i]->out, fb + off, nb);
          off += nb;
        }
      }
      for (auto* r : grp) if (tl) e->tl.ev(r->tl_pid, 'E', "");
    }
    const std::string err = rc ? t_err : "";
    for (auto* r : grp) { if (tl) e->tl.ev(r->tl_pid, 'E', ""); finish_local(e, r->req.name, rc, err); }
    return;
  }
  // ---- device path: b200mpi kernels on the engine's stream ----
  typedef int (*ar_t)(void*, const void*, void*, size_t, int, int, float, int, void*);
  const int dt = b200_dtype(q0.dtype), op = b200_op(q0.redop);
  std::string err;
  int rc = 0;
  if (!e->has_gpu) { rc = HVD_ERR_UNSUPPORTED; err = "device tensor submitted but the engine was started without a GPU executor"; }
  else if (q0.redop == HVD_ADASUM) { rc = HVD_ERR_UNSUPPORTED; err = "Adasum of device tensors goes through the communicator directly (hvd/adasum.py)"; }
  else if (dt < 0 || op < 0) { rc = HVD_ERR_UNSUPPORTED; err = "device allreduce supports float32/bfloat16/float16 with sum/min/max (convert first)"; }
  std::vector<int> handles, pids;
  if (!rc) {
    const float scale = (float)(q0.pre * q0.post);
    for (auto* o : ops) if (o && o->ready_event) e->cu.StreamWaitEvent(e->stream, o->ready_event, 0);
    if (grp.size() == 1 && ops[0]) {
      if (tl) e->tl.ev(grp[0]->tl_pid, 'B', "B200MPI_ALLREDUCE");
      const int r2 = ((ar_t)e->gpu.allreduce)(e->gpu.comm, ops[0]->in, ops[0]->out, (size_t)total, dt, op, scale, 0, e->stream);
      if (r2) { rc = HVD_ERR_TRANSPORT; err = gpu_err(e, "b200mpi_allreduce", r2); }
    } else {
      const size_t need = (size_t)total * es;
      if (e->fusion_dev_bytes < need) {
        if (e->fusion_dev) { e->cu.StreamSynchronize(e->stream); e->cu.Free(e->fusion_dev); }
        e->fusion_dev_bytes = std::max(need, (size_t)e->fusion_threshold.load());
        if (e->cu.Malloc(&e->fusion_dev, e->fusion_dev_bytes)) { rc = HVD_ERR_TRANSPORT; err = "cudaMalloc of the fusion buffer failed"; e->fusion_dev = nullptr; e->fusion_dev_bytes = 0; }
      }
      if (!rc) {
        size_t off = 0;
        for (size_t i = 0; i < grp.size(); i++) {
          const size_t nb = (size_t)grp[i]->req.count * es;
          if (ops[i]) e->cu.MemcpyAsync((char*)e->fusion_dev + off, ops[i]->in, nb, 3 /* D2D */, e->stream);
          else e->cu.MemsetAsync((char*)e->fusion_dev + off, 0, nb, e->stream);
          off += nb;
        }
        for (auto* r : grp) if (tl) e->tl.ev(r->tl_pid, 'B', "B200MPI_ALLREDUCE");
        const int r2 = ((ar_t)e->gpu.allreduce)(e->gpu.comm, e->fusion_dev, e->fusion_dev, (size_t)total, dt, op, scale, 0, e->stream);
        if (r2) { rc = HVD_ERR_TRANSPORT; err = gpu_err(e, "b200mpi_allreduce", r2); }
        off = 0;
        for (size_t i = 0; i < grp.size() && !rc; i++) {
          const size_t nb = (size_t)grp[i]->req.count * es;
          if (ops[i]) e->cu.MemcpyAsync(ops[i]->out, (char*)e->fusion_dev + off, nb, 3, e->stream);
          off += nb;
        }
      }
    }
  }
  for (size_t i = 0; i < grp.size(); i++) {
    auto it = e->pending.find(grp[i]->req.name);
    if (it == e->pending.end()) { if (tl) { if (!rc) e->tl.ev(grp[i]->tl_pid, 'E', ""); e->tl.ev(grp[i]->tl_pid, 'E', ""); } continue; }
    if (rc) { if (tl) e->tl.ev(grp[i]->tl_pid, 'E', ""); finish_local(e, grp[i]->req.name, rc, err); continue; }
    handles.push_back(it->second.handle);
    pids.push_back(grp[i]->tl_pid);
    const std::string name = grp[i]->req.name;
    e->pending.erase(it);
    std::lock_guard<std::mutex> lk(e->mu);
    e->inflight_names.erase(name);   // the name may be reused once the work is queued on the stream (stream order)
  }
  if (!handles.empty()) gpu_track(e, std::move(handles), std::move(pids), "allreduce");
}

void run_single(Engine* e, Response& rs) {
  const Request& q = rs.req;
  auto it = e->pending.find(q.name);
  LocalOp* o = it == e->pending.end() ? nullptr : &it->second;
  const bool tl = e->tl.on();
  int rc = 0;
  std::string err;
  if (tl) e->tl.ev(rs.tl_pid, 'B', kOpName[q.op]);
  switch (q.op) {
    case HVD_BARRIER: {
      std::string er;
      if (e->rv.barrier(e->timeout_ms, &er)) { rc = HVD_ERR_TRANSPORT; err = er; }
      break;
    }
    case HVD_JOIN: rc = rs.last_joined; break;
    case HVD_EXCHANGE:
      if (o) for (int r = 0; r < e->world; r++) memcpy((char*)o->out + (size_t)r * (size_t)q.count, rs.per_rank[r].blob.data(), (size_t)q.count);
      break;
    case HVD_BROADCAST: {
      const size_t nb = (size_t)q.count * esize(q.dtype);
      if (q.devkind == 0) {
        rc = host_bcast(e, o->out, nb, q.root);
        if (rc) err = t_err;
      } else if (!e->has_gpu) { rc = HVD_ERR_UNSUPPORTED; err = "device tensor submitted but the engine was started without a GPU executor"; }
      else {
        typedef int (*bc_t)(void*, void*, size_t, int, void*);
        if (o && o->ready_event) e->cu.StreamWaitEvent(e->stream, o->ready_event, 0);
        const int r2 = ((bc_t)e->gpu.broadcast_bytes)(e->gpu.comm, o->out, nb, q.root, e->stream);
        if (r2) { rc = HVD_ERR_TRANSPORT; err = gpu_err(e, "b200mpi_broadcast_bytes", r2); }
        else {
          const int h = o->handle;
          const std::string name = q.name;
          e->pending.erase(it);
          { std::lock_guard<std::mutex> lk(e->mu); e->inflight_names.erase(name); }
          if (tl) e->tl.ev(rs.tl_pid, 'B', "B200MPI_BCAST");
          gpu_track(e, {h}, {rs.tl_pid}, "broadcast");
          return;
        }
      }
      break;
    }
    case HVD_ALLGATHER: {
      if (q.devkind) { rc = HVD_ERR_UNSUPPORTED; err = "allgather of device tensors goes through the communicator directly"; break; }
      rc = host_allgatherv(e, o->in, o->out, q.extra);
      if (rc) err = t_err;
      break;
    }
    case HVD_ALLTOALL: {
      if (q.devkind) { rc = HVD_ERR_UNSUPPORTED; err = "alltoall of device tensors goes through the communicator directly"; break; }
      std::vector<std::vector<int64_t>> send(e->world);
      for (int s = 0; s < e->world; s++) send[s].assign(rs.per_rank[s].extra.begin(), rs.per_rank[s].extra.begin() + e->world);
      rc = host_alltoallv(e, o->in, o->out, send);
      if (rc) err = t_err;
      break;
    }
    default: rc = HVD_ERR_INVALID; err = "unknown operation"; break;
  }
  if (tl) e->tl.ev(rs.tl_pid, 'E', "");
  e->st.bytes += (uint64_t)q.count * esize(q.dtype);
  finish_local(e, q.name, rc, err);
}

void execute(Engine* e, std::vector<Response>& rsp, int64_t threshold) {
  std::vector<char> used(rsp.size(), 0);
  for (size_t i = 0; i < rsp.size(); i++) {
    if (used[i]) continue;
    used[i] = 1;
    Response& r = rsp[i];
    if (!r.error.empty()) {
      if (e->tl.on()) e->tl.ev(r.tl_pid, 'i', "ERROR");
      finish_local(e, r.req.name, r.code ? r.code : HVD_ERR_MISMATCH, r.error);
      continue;
    }
    if (r.req.op != HVD_ALLREDUCE) { run_single(e, r); continue; }
    // greedy fusion with look-ahead: same dtype / op / scales / device kind, total size under the threshold
    std::vector<Response*> grp{&r};
    int64_t bytes = r.req.count * (int64_t)esize(r.req.dtype);
    for (size_t j = i + 1; j < rsp.size() && bytes < threshold && r.req.redop != HVD_ADASUM; j++) {
      if (used[j] || !rsp[j].error.empty() || rsp[j].req.op != HVD_ALLREDUCE) continue;
      const Request& a = r.req; const Request& b = rsp[j].req;
      if (a.dtype != b.dtype || a.redop != b.redop || a.devkind != b.devkind || memcmp(&a.pre, &b.pre, 8) || memcmp(&a.post, &b.post, 8)) continue;
      const int64_t nb = b.count * (int64_t)esize(b.dtype);
      if (bytes + nb > threshold) continue;
      grp.push_back(&rsp[j]);
      used[j] = 1;
      bytes += nb;
    }
    run_allreduce_group(e, grp);
  }
}

// ------------------------------------------------------------------------------------------------ stall check --
void stall_scan(Engine* e, uint32_t* flags) {
  if (!e->stall_check) return;
  const uint64_t now = now_ns();
  if (now - e->last_stall_scan < 250000000ull) return;
  e->last_stall_scan = now;
  const double warn_s = e->stall_check_s.load(), kill_s = e->stall_shutdown_s.load();
  std::map<int, std::vector<std::string>> missing;
  bool kill = false;
  for (auto& kv : e->table) {
    Entry& en = kv.second;
    if (en.first.op == HVD_JOIN) continue;
    const double age = (double)(now - en.t_first) / 1e9;
    if (kill_s > 0 && age > kill_s) kill = true;
    if (age < warn_s || (en.t_warned && (double)(now - en.t_warned) / 1e9 < warn_s)) continue;
    en.t_warned = now;
    for (int r = 0; r < e->world; r++) if (!((en.bits | e->joined) & (1ull << r))) missing[r].push_back(kv.first);
  }
  if (!missing.empty()) {
    e->st.stall_warnings++;
    if (e->rank == 0) {
      std::string m = "[hvdcore] WARNING: some ranks have been waiting more than " + std::to_string((int)warn_s) +
                      " s for the other ranks to submit the same tensors; if ranks submit different tensors this never completes.\n  ranks that have not submitted:";
      for (auto& kv : missing) {
        m += "\n    rank " + std::to_string(kv.first) + ": ";
        for (size_t i = 0; i < kv.second.size() && i < 8; i++) m += (i ? ", " : "") + kv.second[i];
        if (kv.second.size() > 8) m += ", ... (" + std::to_string(kv.second.size()) + " tensors)";
      }
      fprintf(stderr, "%s\n", m.c_str());
      fflush(stderr);
    }
  }
  if (kill && e->rank == 0) *flags |= kFlagStallShutdown;   // the decision travels with the next exchange so every rank stops in the same cycle
}

// ---------------------------------------------------------------------------------------------------- autotune --
// Rank 0 only. A "step" is a cycle that executed at least one response; a sample is `at_steps` steps under one candidate
// setting, scored by allreduce bytes per second; the first `at_warmup` samples are discarded. Every candidate of the grid is
// tried once, then the best one is kept. The other ranks adopt whatever rank 0 announces (kFlagAutotune in the exchange).
void autotune_step(Engine* e, bool busy) {
  if (!e->autotune || e->autotune_done || e->rank != 0 || !busy) return;
  const uint64_t now = now_ns();
  if (e->at_busy == 0) { e->at_t0 = now; e->at_bytes0 = e->st.bytes.load(); }
  if (++e->at_busy < e->at_steps) return;
  const double secs = (double)(now - e->at_t0) / 1e9;
  const double score = secs > 0 ? (double)(e->st.bytes.load() - e->at_bytes0) / secs : 0;
  e->at_busy = 0;
  e->at_samples++;
  if (e->at_warmup > 0) { e->at_warmup--; return; }
  if (e->at_idx >= 0) {
    if (e->at_log) { fprintf(e->at_log, "%.3f,%.3f,%.3f\n", e->at_grid[e->at_idx].first, (double)e->at_grid[e->at_idx].second / 1048576.0, score / 1e6); fflush(e->at_log); }
    if (score > e->at_best_score) { e->at_best_score = score; e->at_best = e->at_idx; }
  }
  if (++e->at_idx >= (int)e->at_grid.size()) {
    e->autotune_done = true;
    e->at_idx = e->at_best >= 0 ? e->at_best : 0;
    if (e->at_log) { fprintf(e->at_log, "# best: cycle %.3f ms, fusion %.3f MiB, %.3f MB/s\n", e->at_grid[e->at_idx].first, (double)e->at_grid[e->at_idx].second / 1048576.0, e->at_best_score / 1e6); fclose(e->at_log); e->at_log = nullptr; }
  }
  e->cycle_ms = e->at_grid[e->at_idx].first;
  e->fusion_threshold = e->at_grid[e->at_idx].second;
}

// ----------------------------------------------------------------------------------------------------- thread --
void fail_everything(Engine* e, int code, const std::string& why) {
  std::vector<std::string> names;
  for (auto& kv : e->pending) names.push_back(kv.first);
  for (auto& n : names) finish_local(e, n, code, why);
  std::lock_guard<std::mutex> lk(e->mu);
  while (!e->queue.empty()) {
    LocalOp op = std::move(e->queue.front());
    e->queue.pop_front();
    e->inflight_names.erase(op.req.name);
    auto it = e->handles.find(op.handle);
    if (it != e->handles.end()) { it->second.done = true; it->second.status = code; it->second.err = why; }
  }
  e->stop_code = code;
  e->stop_reason = why;
  e->stopped.store(true);
  e->done_cv.notify_all();
}

void engine_main(Engine* e) {
  if (e->has_gpu) {
    e->cu.SetDevice(e->gpu.device);
    if (e->cu.StreamCreateWithPriority(&e->stream, 1 /* non-blocking */, -1)) {
      fail_everything(e, HVD_ERR_TRANSPORT, "hvdcore: cannot create the engine's CUDA stream");
      return;
    }
  }
  uint32_t carry_flags = 0;
  int idle_cycles = 0;
  for (;;) {
    const auto t_start = std::chrono::steady_clock::now();
    e->cycle_no++;
    e->st.cycles++;
    // 1. new local requests -> negotiation message
    Writer w;
    {
      std::lock_guard<std::mutex> lk(e->mu);
      while (!e->queue.empty()) {
        LocalOp op = std::move(e->queue.front());
        e->queue.pop_front();
        const Request& q = op.req;
        auto ci = e->cache.find(q.name);
        if (e->cache_capacity > 0 && ci != e->cache.end() && same_sig(ci->second.sig, q)) { w.put<uint8_t>(1); w.put<uint32_t>(ci->second.id); }
        else encode(w, q);
        e->pending.emplace(q.name, std::move(op));
      }
    }
    bool announce = e->autotune && e->rank == 0 && (!e->autotune_done || e->at_tail > 0);
    if (announce && e->autotune_done) e->at_tail--;
    uint32_t flags = carry_flags | (e->shutdown_requested.load() ? kFlagShutdown : 0) | (announce ? kFlagAutotune : 0);
    carry_flags = 0;
    // 2. exchange + replicated coordination
    std::vector<std::string> msgs;
    uint32_t all_flags = 0;
    int64_t fusion = 0;
    if (exchange(e, w.s, flags, &msgs, &all_flags, &fusion)) { fail_everything(e, HVD_ERR_TRANSPORT, "hvdcore: " + t_err); break; }
    if (e->autotune && e->rank != 0) {     // followers: the search is over once rank 0 stops announcing candidates
      if (all_flags & kFlagAutotune) e->at_seen = true;
      else if (e->at_seen) e->autotune_done = true;
    }
    std::vector<Response> rsp;
    coordinate(e, msgs, &rsp);
    if (e->mark_cycles && !rsp.empty() && e->tl.on())   // HOROVOD_TIMELINE_MARK_CYCLES: a tick on row 0 for every cycle with work
      e->tl.ev(0, 'i', "CYCLE_START", "\"cycle\": " + std::to_string(e->cycle_no) + ", \"responses\": " + std::to_string(rsp.size()));
    // 3. fused execution
    execute(e, rsp, fusion);
    if (!e->gpu_inflight.empty()) gpu_poll(e, false);
    if (all_flags & kFlagStallShutdown) {
      if (e->has_gpu) gpu_poll(e, true);
      fail_everything(e, HVD_ERR_STALL, "hvdcore: stall shutdown time exceeded: some ranks never submitted tensors the others are waiting for");
      break;
    }
    if (all_flags & kFlagShutdown) {
      if (e->has_gpu) gpu_poll(e, true);
      fail_everything(e, HVD_ERR_SHUTDOWN, "Horovod has been shut down");
      break;
    }
    stall_scan(e, &carry_flags);
    autotune_step(e, !rsp.empty());
    // 4. sleep out the rest of the cycle unless there is work waiting
    const bool busy = !rsp.empty() || !e->gpu_inflight.empty();
    // nothing submitted, pending or half-negotiated anywhere for 100 cycles: stretch the cycle five-fold (a local submission
    // still wakes this thread at once; a peer that is backing off adds at most that stretch to the first tensor after a pause)
    idle_cycles = (busy || !e->pending.empty() || !e->table.empty()) ? 0 : idle_cycles + 1;
    const double stretch = idle_cycles > 100 ? 5.0 : 1.0;
    if (!busy) {
      std::unique_lock<std::mutex> lk(e->mu);
      const auto deadline = t_start + std::chrono::microseconds((int64_t)(e->cycle_ms.load() * 1000.0 * stretch));
      e->queue_cv.wait_until(lk, deadline, [&] { return !e->queue.empty() || e->shutdown_requested.load(); });
    } else if (!e->gpu_inflight.empty() && rsp.empty()) {
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
  if (e->stream && e->fusion_dev) { e->cu.StreamSynchronize(e->stream); e->cu.Free(e->fusion_dev); e->fusion_dev = nullptr; }
  e->tl.close();
  if (e->seg) { Rendezvous::close_boxes(e->seg, e->seg_bytes); e->seg = nullptr; }
  if (e->ada_seg) { Rendezvous::close_boxes(e->ada_seg, e->ada_total); e->ada_seg = nullptr; e->ada_slot = 0; }
}

}  // namespace

// ======================================================================================================= C ABI ==
extern "C" {

const char* hvdcore_last_error(void) { return t_err.c_str(); }

int hvdcore_init(const char* job_id, int rank, int world, const hvdcore_gpu_t* gpu) {
  std::lock_guard<std::mutex> gl(g_mu);
  if (g) return fail(HVD_ERR_INVALID, "hvdcore: already initialised");
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(HVD_ERR_INVALID, "hvdcore: invalid rank / world (1..64 ranks)");
  auto e = std::make_unique<Engine>();
  e->rank = rank;
  e->world = world;
  e->cycle_ms = env_d("HOROVOD_CYCLE_TIME", 1.0);
  e->fusion_threshold = (int64_t)env_d("HOROVOD_FUSION_THRESHOLD", (double)(64ll << 20));
  e->cache_capacity = (int)env_d("HOROVOD_CACHE_CAPACITY", 1024);
  e->stall_check = env_d("HOROVOD_STALL_CHECK_DISABLE", 0) == 0;
  e->mark_cycles = env_d("HOROVOD_TIMELINE_MARK_CYCLES", 0) != 0;
  e->autotune = env_d("HOROVOD_AUTOTUNE", 0) != 0;
  if (e->autotune) {
    e->at_warmup = (int)env_d("HOROVOD_AUTOTUNE_WARMUP_SAMPLES", 3);
    e->at_steps = std::max(1, (int)env_d("HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE", 10));
    for (double c : {0.5, 1.0, 2.5, 5.0})
      for (int64_t f : {1ll << 20, 4ll << 20, 16ll << 20, 64ll << 20, 128ll << 20}) e->at_grid.emplace_back(c, f);
    const char* lg = getenv("HOROVOD_AUTOTUNE_LOG");
    if (lg && *lg && rank == 0) { e->at_log = fopen(lg, "w"); if (e->at_log) fputs("cycle_time_ms,fusion_threshold_mb,score_mb_per_s\n", e->at_log); }
  }
  e->stall_check_s = env_d("HOROVOD_STALL_CHECK_TIME_SECONDS", 60);
  e->stall_shutdown_s = env_d("HOROVOD_STALL_SHUTDOWN_TIME_SECONDS", 0);
  e->timeout_ms = (int)env_d("B200MPI_HVD_TIMEOUT_MS", 600000);
  if (gpu && gpu->comm) {
    std::string err;
    if (!e->cu.load(&err)) return fail(HVD_ERR_UNSUPPORTED, err);
    if (!gpu->allreduce || !gpu->broadcast_bytes) return fail(HVD_ERR_INVALID, "hvdcore: GPU executor needs allreduce and broadcast_bytes");
    e->gpu = *gpu;
    e->has_gpu = true;
  }
  // A process that re-initialises under the SAME job id (hvd.shutdown(); hvd.init()) needs a fresh segment name, and every
  // rank does that the same number of times. An elastic world re-formed in place carries its generation in the job id
  // itself (survivors re-initialise, newcomers initialise for the first time): a new id restarts the private counter.
  static std::string last_job;
  const std::string jid = job_id && *job_id ? job_id : "default";
  if (jid != last_job) { g_generation = 0; last_job = jid; }
  std::string name = jid + "-hvd";
  if (g_generation) name += "-g" + std::to_string(g_generation);
  std::string err;
  const int init_timeout = (int)env_d("B200MPI_INIT_TIMEOUT_MS", env_d("B200MPI_TIMEOUT_MS", 60000));
  if (e->rv.attach(name, rank, world, gpu ? gpu->device : -1, init_timeout, &err)) return fail(HVD_ERR_TRANSPORT, err);
  g_generation++;
  open_data_segment(e.get());
  const char* tlp = getenv("HOROVOD_TIMELINE");
  if (tlp && *tlp && rank == 0) e->tl.open(tlp);
  Engine* raw = e.release();
  raw->thread = std::thread(engine_main, raw);
  g = raw;
  return 0;
}

int hvdcore_shutdown(void) {
  std::lock_guard<std::mutex> gl(g_mu);
  if (!g) return 0;
  g->shutdown_requested.store(true);
  { std::lock_guard<std::mutex> lk(g->mu); g->queue_cv.notify_all(); }
  if (g->thread.joinable()) g->thread.join();
  g->rv.detach(g->rank == 0);
  delete g;
  g = nullptr;
  return 0;
}

int hvdcore_initialized(void) { return g && !g->stopped.load() ? 1 : 0; }
int hvdcore_rank(void) { return g ? g->rank : -1; }
int hvdcore_size(void) { return g ? g->world : -1; }

int hvdcore_enqueue(hvd_op_t op, const char* name, const void* in, void* out, int64_t count, hvd_dtype_t dtype, hvd_redop_t redop,
                    int root, double prescale, double postscale, int device, void* ready_event, const int64_t* extra, int n_extra) {
  Engine* e = g;
  if (!e) return fail(HVD_ERR_NOT_INIT, "hvdcore: not initialised");
  if ((int)op < 0 || op > HVD_EXCHANGE) return fail(HVD_ERR_INVALID, "hvdcore: unknown operation");
  if (count < 0 || !esize(dtype)) return fail(HVD_ERR_INVALID, "hvdcore: bad count / dtype");
  const bool needs_buf = op != HVD_BARRIER && op != HVD_JOIN;
  if (needs_buf && count > 0 && (!out || (op != HVD_BROADCAST && !in))) return fail(HVD_ERR_INVALID, "hvdcore: null buffer");
  if (op == HVD_BROADCAST && (root < 0 || root >= e->world)) return fail(HVD_ERR_INVALID, "hvdcore: broadcast root out of range");
  if (op == HVD_EXCHANGE && (dtype != HVD_U8 || count > (int64_t)kMaxBlob || device >= 0)) return fail(HVD_ERR_INVALID, "hvdcore: exchange takes <= 1024 host bytes");
  if (op == HVD_ALLGATHER && n_extra != e->world) return fail(HVD_ERR_INVALID, "hvdcore: allgather needs one byte count per rank");
  if (op == HVD_ALLTOALL && n_extra != 2 * e->world) return fail(HVD_ERR_INVALID, "hvdcore: alltoall needs send and receive byte counts per rank");
  if ((op == HVD_ALLREDUCE) && (redop < HVD_SUM || redop > HVD_ADASUM)) return fail(HVD_ERR_INVALID, "hvdcore: unknown reduction");
  LocalOp lo;
  Request& q = lo.req;
  q.op = (uint8_t)op; q.dtype = (uint8_t)dtype; q.redop = (uint8_t)redop; q.devkind = device >= 0 ? 1 : 0;
  q.root = op == HVD_BROADCAST ? root : 0;
  q.pre = op == HVD_ALLREDUCE ? prescale : 1.0;
  q.post = op == HVD_ALLREDUCE ? postscale : 1.0;
  q.count = count;
  if (extra && n_extra > 0) q.extra.assign(extra, extra + n_extra);
  if (op == HVD_EXCHANGE && count) q.blob.assign((const char*)in, (size_t)count);
  lo.in = in; lo.out = out; lo.device = device; lo.ready_event = ready_event;
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->stopped.load()) return fail(e->stop_code, e->stop_reason.empty() ? "Horovod has been shut down" : e->stop_reason);
  if (name && *name) q.name = name;
  else q.name = std::string(kOpName[op]) + ".noname." + std::to_string(e->noname[op]++);
  if (q.name.size() > 4096) return fail(HVD_ERR_INVALID, "hvdcore: tensor name too long");
  if (!e->inflight_names.insert(q.name).second)
    return fail(HVD_ERR_DUPLICATE, "hvdcore: a collective named '" + q.name + "' is already in flight on this rank; names must be unique until the handle completes");
  lo.handle = e->next_handle++;
  e->handles[lo.handle] = Handle{};
  const int h = lo.handle;
  e->queue.push_back(std::move(lo));
  e->queue_cv.notify_one();
  return h;
}

int hvdcore_poll(int handle) {
  Engine* e = g;
  if (!e) return fail(HVD_ERR_NOT_INIT, "hvdcore: not initialised");
  std::lock_guard<std::mutex> lk(e->mu);
  auto it = e->handles.find(handle);
  if (it == e->handles.end()) return fail(HVD_ERR_INVALID, "hvdcore: unknown handle");
  return it->second.done ? 1 : 0;
}

int hvdcore_wait(int handle) {
  Engine* e = g;
  if (!e) return fail(HVD_ERR_NOT_INIT, "hvdcore: not initialised");
  std::unique_lock<std::mutex> lk(e->mu);
  auto it = e->handles.find(handle);
  if (it == e->handles.end()) return fail(HVD_ERR_INVALID, "hvdcore: unknown handle");
  e->done_cv.wait(lk, [&] { return e->handles[handle].done; });
  Handle hd = e->handles[handle];
  e->handles.erase(handle);
  lk.unlock();
  if (hd.status < 0) t_err = hd.err;
  return hd.status;
}

int hvdcore_start_timeline(const char* path) {
  if (!g) return fail(HVD_ERR_NOT_INIT, "hvdcore: not initialised");
  if (g->rank == 0 && path && *path) g->tl.open(path);
  return 0;
}
int hvdcore_stop_timeline(void) {
  if (!g) return fail(HVD_ERR_NOT_INIT, "hvdcore: not initialised");
  g->tl.close();
  return 0;
}

int hvdcore_stats_json(char* buf, size_t cap) {
  Engine* e = g;
  char tmp[768];
  if (!e) { tmp[0] = '{'; tmp[1] = '}'; tmp[2] = 0; }
  else {
    snprintf(tmp, sizeof(tmp),
             "{\"rank\": %d, \"world\": %d, \"cycles\": %llu, \"tensors\": %llu, \"fused_groups\": %llu, \"fused_tensors\": %llu, "
             "\"bytes\": %llu, \"cache_hits\": %llu, \"cache_misses\": %llu, \"negotiation_bytes\": %llu, \"stall_warnings\": %llu, "
             "\"errors\": %llu, \"cycle_time_ms\": %.3f, \"fusion_threshold\": %lld, \"cache_capacity\": %d, \"gpu\": %s, \"mailbox_bytes\": %zu, \"autotune\": %s, \"autotune_samples\": %llu}",
             e->rank, e->world, (unsigned long long)e->st.cycles.load(), (unsigned long long)e->st.tensors.load(),
             (unsigned long long)e->st.groups.load(), (unsigned long long)e->st.fused_tensors.load(), (unsigned long long)e->st.bytes.load(),
             (unsigned long long)e->st.cache_hits.load(), (unsigned long long)e->st.cache_misses.load(),
             (unsigned long long)e->st.negotiation_bytes.load(), (unsigned long long)e->st.stall_warnings.load(),
             (unsigned long long)e->st.errors.load(), e->cycle_ms.load(), (long long)e->fusion_threshold.load(), e->cache_capacity,
             e->has_gpu ? "true" : "false", e->seg ? e->box : (size_t)kRvMailbox,
             !e->autotune ? "\"off\"" : (e->autotune_done ? "\"done\"" : "\"searching\""), (unsigned long long)e->at_samples);
  }
  const size_t n = strlen(tmp);
  if (buf && cap) { const size_t k = n < cap - 1 ? n : cap - 1; memcpy(buf, tmp, k); buf[k] = 0; }
  return (int)n;
}

int hvdcore_set_param(const char* key, double value) {
  Engine* e = g;
  if (!e) return fail(HVD_ERR_NOT_INIT, "hvdcore: not initialised");
  const std::string k = key ? key : "";
  if (k == "cycle_time_ms") e->cycle_ms = value;
  else if (k == "fusion_threshold") e->fusion_threshold = (int64_t)value;
  else if (k == "stall_check_s") e->stall_check_s = value;
  else if (k == "stall_shutdown_s") e->stall_shutdown_s = value;
  else return fail(HVD_ERR_INVALID, "hvdcore: unknown parameter " + k);
  return 0;
}

}  // extern "C"
