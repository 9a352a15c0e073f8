// The point-to-point protocol of csrc/kernels/p2p.cu executed on the host: the same p2p_run_op<> template the kernel runs,
// instantiated with a platform made of std::thread / atomics / memcpy. "Ranks" are mailbox windows in this process, every
// operation of a batch runs in its own thread (= its own CTA). Checks, without a GPU: chunk sequence numbers and slot
// reuse across batches, the counters handed from batch to batch, eager completion of sends up to two chunks, back-pressure
// on the third chunk, several messages per peer inside one batch, ragged sizes. Batches are planned by the real p2p_plan()
// of comm.cc. Built and run by `make test_p2p_protocol` / tests/test_native_cpu.py.
#include "../runtime/comm.cc"
#include "../kernels/p2p.cu"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

using namespace b200mpi;

static int g_failed = 0;
#define EXPECT(cond)                                                                       \
  do {                                                                                     \
    if (!(cond)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); g_failed++; } \
  } while (0)

struct P2PHost {
  std::atomic<int>* timeouts;
  bool wait_ge(const uint32_t* flag, uint32_t want, int) {
    const auto t0 = std::chrono::steady_clock::now();
    while ((int32_t)(__atomic_load_n(flag, __ATOMIC_ACQUIRE) - want) < 0) {
      std::this_thread::yield();
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { (*timeouts)++; return false; }
    }
    return true;
  }
  void sync() {}
  void release(uint32_t* flag, uint32_t v) { __atomic_store_n(flag, v, __ATOMIC_RELEASE); }
  void push(char* dst, const char* src, size_t n) { memcpy(dst, src, n); }
  void pull(char* dst, const char* src, size_t n) { memcpy(dst, src, n); }
};

struct World {
  int n;
  std::vector<char*> box;                       // mailbox window of every rank
  std::vector<std::vector<uint32_t>> cnt;       // per-rank chunk counters (device memory in the real thing)
  std::atomic<int> timeouts{0};
  explicit World(int n_) : n(n_), box(kMaxRanks, nullptr), cnt(n_, std::vector<uint32_t>(2 * kMaxRanks, 0)) {
    for (int r = 0; r < n; r++) box[r] = static_cast<char*>(calloc(1, kP2PWindowBytes));
  }
  ~World() { for (char* b : box) free(b); }

  struct Running { std::vector<std::thread> th; P2PCommit add; int rank; };
  // starts one thread per operation (like one CTA per operation); the caller joins and commits
  Running start(int rank, const std::vector<b200mpi_p2p_op_t>& ops) {
    Running run;
    run.rank = rank;
    auto* a = new P2PArgs;
    EXPECT(p2p_plan(rank, n, ops.data(), (int)ops.size(), a, &run.add) == 0);
    for (int i = 0; i < a->nops; i++)
      run.th.emplace_back([this, a, i, rank] {
        P2PHost pf{&timeouts};
        p2p_run_op(a->ops[i], rank, box.data(), cnt[rank].data(), pf);
      });
    return run;
  }
  void finish(Running& run) {
    for (auto& t : run.th) t.join();
    for (int i = 0; i < 2 * kMaxRanks; i++) cnt[run.rank][i] += run.add.n[i];   // k_p2p_commit
  }
  void run_all(const std::vector<std::vector<b200mpi_p2p_op_t>>& per_rank) {
    std::vector<Running> runs;
    for (int r = 0; r < n; r++) runs.push_back(start(r, per_rank[r]));
    for (auto& run : runs) finish(run);
  }
};

static std::vector<char> pattern(size_t n, int seed) {
  std::vector<char> v(n);
  uint32_t x = 2654435761u * (uint32_t)(seed + 1);
  for (size_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; v[i] = (char)(x >> 24); }
  return v;
}

int main() {
  const int N = 4;
  World w(N);

  // 1. ring shift, sizes around the chunk boundaries; counters carry over from batch to batch
  for (size_t bytes : {(size_t)1, (size_t)1000, kP2PChunk, kP2PChunk + 13, 5 * kP2PChunk, 9 * kP2PChunk + 7, (size_t)17}) {
    std::vector<std::vector<char>> src(N), dst(N);
    std::vector<std::vector<b200mpi_p2p_op_t>> ops(N);
    for (int r = 0; r < N; r++) {
      src[r] = pattern(bytes, 100 * r + (int)(bytes % 97));
      dst[r].assign(bytes + 8, (char)0x5a);   // 8 guard bytes: a receive must not write past its size
      ops[r] = {{src[r].data(), nullptr, bytes, (r + 1) % N, 1}, {nullptr, dst[r].data(), bytes, (r + N - 1) % N, 0}};
    }
    w.run_all(ops);
    for (int r = 0; r < N; r++) {
      EXPECT(memcmp(dst[r].data(), src[(r + N - 1) % N].data(), bytes) == 0);
      for (int g = 0; g < 8; g++) EXPECT(dst[r][bytes + g] == (char)0x5a);
    }
  }
  EXPECT(w.cnt[0][1] == w.cnt[1][kMaxRanks + 0] && w.cnt[0][1] > 10);   // sender's and receiver's view of stream 0 -> 1 agree

  // 2. eager: sends of <= 2 chunks complete although nobody receives yet; the receives then find the data
  {
    std::vector<std::vector<char>> src(N), dst(N);
    const size_t bytes = 2 * kP2PChunk;
    for (int r = 0; r < N; r++) {
      src[r] = pattern(bytes, 7000 + r);
      dst[r].assign(bytes, 0);
      auto run = w.start(r, {{src[r].data(), nullptr, bytes, r ^ 1, 1}});
      w.finish(run);                                  // returns: no receiver involved
    }
    for (int r = 0; r < N; r++) {
      auto run = w.start(r, {{nullptr, dst[r].data(), bytes, r ^ 1, 0}});
      w.finish(run);
      EXPECT(memcmp(dst[r].data(), src[r ^ 1].data(), bytes) == 0);
    }
  }

  // 3. back-pressure: the third chunk of an unmatched send waits for the receiver to free a slot
  {
    const size_t bytes = 3 * kP2PChunk + 5;
    auto src = pattern(bytes, 4242);
    std::vector<char> dst(bytes, 0);
    std::atomic<bool> done{false};
    auto send = w.start(0, {{src.data(), nullptr, bytes, 2, 1}});
    std::thread watcher([&] { for (auto& t : send.th) t.join(); done = true; });
    std::this_thread::sleep_for(std::chrono::milliseconds(150));
    EXPECT(!done.load());                              // still parked on the acknowledgement of chunk 0
    auto recv = w.start(2, {{nullptr, dst.data(), bytes, 0, 0}});
    w.finish(recv);
    watcher.join();
    send.th.clear();
    w.finish(send);
    EXPECT(done.load() && memcmp(dst.data(), src.data(), bytes) == 0);
  }

  // 4. all-to-all by point-to-point, two messages per peer in one batch (per-stream chunk offsets), three rounds
  for (int round = 0; round < 3; round++) {
    std::vector<std::vector<b200mpi_p2p_op_t>> ops(N);
    std::vector<std::vector<std::vector<char>>> big_s(N, std::vector<std::vector<char>>(N)), sm_s = big_s, big_d = big_s, sm_d = big_s;
    for (int r = 0; r < N; r++)
      for (int p = 0; p < N; p++) {
        if (p == r) continue;
        big_s[r][p] = pattern(kP2PChunk + 4096 * (r + 1), round * 1000 + r * 10 + p);
        sm_s[r][p] = pattern(12, round * 1000 + 500 + r * 10 + p);
        big_d[r][p].assign(kP2PChunk + 4096 * (p + 1), 0);
        sm_d[r][p].assign(12, 0);
        ops[r].push_back({big_s[r][p].data(), nullptr, big_s[r][p].size(), p, 1});
        ops[r].push_back({sm_s[r][p].data(), nullptr, sm_s[r][p].size(), p, 1});
        ops[r].push_back({nullptr, big_d[r][p].data(), big_d[r][p].size(), p, 0});
        ops[r].push_back({nullptr, sm_d[r][p].data(), sm_d[r][p].size(), p, 0});
      }
    w.run_all(ops);
    for (int r = 0; r < N; r++)
      for (int p = 0; p < N; p++) {
        if (p == r) continue;
        EXPECT(big_d[r][p] == big_s[p][r]);
        EXPECT(sm_d[r][p] == sm_s[p][r]);
      }
  }
  EXPECT(w.timeouts.load() == 0);
  for (int r = 0; r < N; r++)
    for (int p = 0; p < N; p++) EXPECT(w.cnt[r][p] == w.cnt[p][kMaxRanks + r]);   // every stream: chunks sent == chunks received

  printf(g_failed ? "p2p_protocol_test: %d check(s) FAILED\n" : "p2p_protocol_test: all checks passed\n", g_failed);
  return g_failed ? 1 : 0;
}
