// Multi-process test of the hvdcore engine (csrc/hvd_core): run as `mpirun -n 4 hvd_core_test` (also under ASAN and
// TSAN, see the Makefile). Every check that fails prints a line and makes the process exit non-zero.
//   - name-based negotiation with a different submission order on every rank, fusion, response cache
//   - every dtype / reduction of the host path against closed forms, messages larger than one mailbox
//   - allgatherv / broadcast / alltoallv / exchange / barrier
//   - mismatch and duplicate-name errors leave the engine usable
//   - join() with an uneven number of steps per rank
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "../hvd_core/hvd_core.h"

static int g_rank = 0, g_world = 1, g_bad = 0;
#define CHECK(cond, ...) do { if (!(cond)) { g_bad++; fprintf(stderr, "[rank %d] %s:%d: ", g_rank, __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static int env_int(std::initializer_list<const char*> names, int d) {
  for (const char* n : names) { const char* v = getenv(n); if (v && *v) return atoi(v); }
  return d;
}
// HVD_TEST_JITTER_US=<n>: every submission is preceded by a rank-dependent pause of up to n microseconds, so requests for one
// tensor reach the coordinator in different cycles (entries then live across cycles, the usual case in real training)
static unsigned g_jitter = 0, g_seed = 1;
static void jitter() {
  if (!g_jitter) return;
  g_seed = g_seed * 1103515245u + 12345u;
  usleep((g_seed >> 8) % g_jitter);
}
static int ar(const char* name, const void* in, void* out, int64_t n, hvd_dtype_t dt, hvd_redop_t op = HVD_SUM, double pre = 1, double post = 1) {
  jitter();
  return hvdcore_enqueue(HVD_ALLREDUCE, name, in, out, n, dt, op, 0, pre, post, -1, nullptr, nullptr, 0);
}
static long long stat(const char* key) {
  char buf[1024];
  hvdcore_stats_json(buf, sizeof(buf));
  const char* p = strstr(buf, key);
  return p ? atoll(p + strlen(key) + 3) : -1;   // "key": value
}
static uint16_t to_bf16(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
static float from_bf16(uint16_t v) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  g_rank = env_int({"B200MPI_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK"}, 0);
  g_world = env_int({"B200MPI_WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE"}, 1);
  g_jitter = (unsigned)env_int({"HVD_TEST_JITTER_US"}, 0);
  g_seed = 77u + 13u * (unsigned)g_rank;
  const char* job = getenv("B200MPI_JOB_ID");
  const int W = g_world, R = g_rank;
  int rc = hvdcore_init(job ? job : "hvdcore-test", R, W, nullptr);
  if (rc) { fprintf(stderr, "[rank %d] init failed: %s\n", R, hvdcore_last_error()); return 2; }
  CHECK(hvdcore_rank() == R && hvdcore_size() == W && hvdcore_initialized() == 1, "identity");
  const double tri = W * (W - 1) / 2.0;

  // ---- 1. out-of-order named allreduces: fused, correct, then cached on later iterations --------------------------
  const int NT = 40;
  std::vector<std::vector<float>> grads(NT);
  for (int it = 0; it < 4; it++) {
    std::vector<int> order(NT);
    for (int i = 0; i < NT; i++) order[i] = i;
    std::mt19937 rng(1000 * it + R);        // a different order on every rank
    std::shuffle(order.begin(), order.end(), rng);
    std::vector<int> handles(NT);
    for (int i : order) {
      grads[i].assign(10 + 37 * i, 0.f);
      for (size_t k = 0; k < grads[i].size(); k++) grads[i][k] = (float)(R + 1) * (float)(i + 1) + (float)k;
      const std::string name = "grad." + std::to_string(i);
      handles[i] = ar(name.c_str(), grads[i].data(), grads[i].data(), (int64_t)grads[i].size(), HVD_F32, HVD_SUM, 1.0, 1.0 / W);
      CHECK(handles[i] > 0, "enqueue %s: %s", name.c_str(), hvdcore_last_error());
    }
    for (int i = 0; i < NT; i++) {
      const int s = hvdcore_wait(handles[i]);
      CHECK(s == 0, "wait grad.%d -> %d (%s)", i, s, hvdcore_last_error());
      for (size_t k = 0; k < grads[i].size(); k++) {
        const float want = (float)((tri + W) * (i + 1) / W) + (float)k;   // mean over ranks of (r+1)(i+1) + k
        if (fabsf(grads[i][k] - want) > 1e-3f * fabsf(want)) { CHECK(false, "grad.%d[%zu] = %f, want %f", i, k, grads[i][k], want); break; }
      }
    }
  }
  CHECK(stat("tensors") == 4 * NT, "tensors = %lld", stat("tensors"));
  CHECK(stat("fused_groups") < 4 * NT, "no fusion happened: %lld groups", stat("fused_groups"));
  CHECK(stat("cache_hits") >= 2 * NT, "response cache unused: %lld hits", stat("cache_hits"));
  CHECK(stat("cache_misses") >= NT, "cache misses = %lld", stat("cache_misses"));

  // ---- 2. dtypes and reductions ------------------------------------------------------------------------------------
  {
    std::vector<int32_t> a(1000); for (size_t i = 0; i < a.size(); i++) a[i] = (int32_t)i * (R + 1);
    std::vector<double> b(777); for (size_t i = 0; i < b.size(); i++) b[i] = (double)i - 100.0 * R;
    std::vector<uint16_t> c(513); for (size_t i = 0; i < c.size(); i++) c[i] = to_bf16((float)(R + 1));
    std::vector<uint8_t> d(64); for (size_t i = 0; i < d.size(); i++) d[i] = (uint8_t)(10 + R + i);
    std::vector<int64_t> e64(5, 2 + R);
    std::vector<uint8_t> flags(8); for (size_t i = 0; i < flags.size(); i++) flags[i] = (uint8_t)((int)i % W == R);
    std::vector<int64_t> big(3 * 8192 + 5); for (size_t i = 0; i < big.size(); i++) big[i] = (int64_t)i + R;   // > 64 KiB
    int h[7];
    h[0] = ar("t.i32", a.data(), a.data(), (int64_t)a.size(), HVD_I32);
    h[1] = ar("t.f64max", b.data(), b.data(), (int64_t)b.size(), HVD_F64, HVD_MAX);
    h[2] = ar("t.bf16", c.data(), c.data(), (int64_t)c.size(), HVD_BF16);
    h[3] = ar("t.u8min", d.data(), d.data(), (int64_t)d.size(), HVD_U8, HVD_MIN);
    h[4] = ar("t.i64prod", e64.data(), e64.data(), (int64_t)e64.size(), HVD_I64, HVD_PROD);
    h[5] = ar("t.bool", flags.data(), flags.data(), (int64_t)flags.size(), HVD_BOOL, HVD_SUM);
    h[6] = ar("t.big", big.data(), big.data(), (int64_t)big.size(), HVD_I64);
    for (int i = 0; i < 7; i++) CHECK(hvdcore_wait(h[i]) == 0, "dtype case %d: %s", i, hvdcore_last_error());
    for (size_t i = 0; i < a.size(); i++) if (a[i] != (int32_t)(i * (tri + W))) { CHECK(false, "i32[%zu]=%d", i, a[i]); break; }
    for (size_t i = 0; i < b.size(); i++) if (b[i] != (double)i) { CHECK(false, "f64max[%zu]=%f", i, b[i]); break; }
    for (size_t i = 0; i < c.size(); i++) if (from_bf16(c[i]) != (float)(tri + W)) { CHECK(false, "bf16[%zu]=%f", i, from_bf16(c[i])); break; }
    for (size_t i = 0; i < d.size(); i++) if (d[i] != (uint8_t)(10 + i)) { CHECK(false, "u8min[%zu]=%d", i, d[i]); break; }
    long long prod = 1; for (int r = 0; r < W; r++) prod *= 2 + r;
    CHECK(e64[0] == prod, "i64 prod = %lld want %lld", (long long)e64[0], prod);
    for (size_t i = 0; i < flags.size(); i++) if (flags[i] != 1) { CHECK(false, "bool[%zu]=%d", i, flags[i]); break; }
    for (size_t i = 0; i < big.size(); i++) if (big[i] != (int64_t)(W * i + tri)) { CHECK(false, "big[%zu]=%lld", i, (long long)big[i]); break; }
  }

  // ---- 3. small fusion threshold: several groups, out != in ----------------------------------------------------------
  {
    hvdcore_set_param("fusion_threshold", 4096);
    const long long g0 = stat("fused_groups");
    std::vector<std::vector<float>> in(12), out(12);
    std::vector<int> h(12);
    for (int i = 0; i < 12; i++) {
      in[i].assign(300, (float)(R + i)); out[i].assign(300, -1.f);
      h[i] = ar(("thr." + std::to_string(i)).c_str(), in[i].data(), out[i].data(), 300, HVD_F32);
    }
    for (int i = 0; i < 12; i++) {
      CHECK(hvdcore_wait(h[i]) == 0, "threshold case");
      CHECK(out[i][299] == (float)(tri + W * i) && in[i][0] == (float)(R + i), "thr.%d = %f", i, out[i][299]);
    }
    CHECK(stat("fused_groups") - g0 >= 4, "threshold ignored: %lld groups for 12 x 1200 B with a 4 KiB threshold", stat("fused_groups") - g0);
    hvdcore_set_param("fusion_threshold", 64 << 20);
  }

  // ---- 4. exchange -> allgatherv, broadcast, alltoallv, barrier ------------------------------------------------------
  {
    int64_t mine = 8 * (int64_t)(3 + 2 * R);   // bytes
    std::vector<int64_t> counts(W);
    int h = hvdcore_enqueue(HVD_EXCHANGE, "sizes", &mine, counts.data(), 8, HVD_U8, HVD_SUM, 0, 1, 1, -1, nullptr, nullptr, 0);
    CHECK(hvdcore_wait(h) == 0, "exchange: %s", hvdcore_last_error());
    int64_t total = 0;
    for (int r = 0; r < W; r++) { CHECK(counts[r] == 8 * (3 + 2 * r), "exchange[%d]=%lld", r, (long long)counts[r]); total += counts[r]; }
    std::vector<int64_t> src((size_t)mine / 8, 100 + R), dst((size_t)total / 8, -1);
    h = hvdcore_enqueue(HVD_ALLGATHER, "gather", src.data(), dst.data(), (int64_t)src.size(), HVD_I64, HVD_SUM, 0, 1, 1, -1, nullptr, counts.data(), W);
    CHECK(hvdcore_wait(h) == 0, "allgather: %s", hvdcore_last_error());
    size_t k = 0;
    for (int r = 0; r < W; r++) for (int i = 0; i < 3 + 2 * r; i++, k++) CHECK(dst[k] == 100 + r, "allgather[%zu]=%lld", k, (long long)dst[k]);

    std::vector<float> bc(70000, R == W - 1 ? 3.5f : 0.f);   // larger than a mailbox
    h = hvdcore_enqueue(HVD_BROADCAST, "bcast", nullptr, bc.data(), (int64_t)bc.size(), HVD_F32, HVD_SUM, W - 1, 1, 1, -1, nullptr, nullptr, 0);
    CHECK(hvdcore_wait(h) == 0, "broadcast: %s", hvdcore_last_error());
    CHECK(bc[0] == 3.5f && bc[69999] == 3.5f, "broadcast value %f", bc[69999]);

    // rank s sends (s + d + 1) int32 to rank d, each equal to s * 100 + d
    std::vector<int64_t> splits(2 * W);
    std::vector<int32_t> sbuf, rbuf;
    for (int d = 0; d < W; d++) { splits[d] = 4 * (R + d + 1); for (int i = 0; i < R + d + 1; i++) sbuf.push_back(R * 100 + d); }
    for (int s = 0; s < W; s++) splits[W + s] = 4 * (s + R + 1);
    int64_t rtot = 0; for (int s = 0; s < W; s++) rtot += splits[W + s];
    rbuf.assign((size_t)rtot / 4, -1);
    h = hvdcore_enqueue(HVD_ALLTOALL, "a2a", sbuf.data(), rbuf.data(), (int64_t)sbuf.size(), HVD_I32, HVD_SUM, 0, 1, 1, -1, nullptr, splits.data(), 2 * W);
    CHECK(hvdcore_wait(h) == 0, "alltoall: %s", hvdcore_last_error());
    k = 0;
    for (int s = 0; s < W; s++) for (int i = 0; i < s + R + 1; i++, k++) CHECK(rbuf[k] == s * 100 + R, "alltoall[%zu]=%d", k, rbuf[k]);

    h = hvdcore_enqueue(HVD_BARRIER, nullptr, nullptr, nullptr, 0, HVD_U8, HVD_SUM, 0, 1, 1, -1, nullptr, nullptr, 0);
    CHECK(hvdcore_wait(h) == 0, "barrier");
  }

  // ---- 5. errors: mismatched sizes reach every rank, duplicate names are refused, the engine keeps going --------------
  if (W > 1) {
    std::vector<float> x(16 + (R == 1 ? 4 : 0), 1.f);
    int h = ar("bad.size", x.data(), x.data(), (int64_t)x.size(), HVD_F32);
    const int s = hvdcore_wait(h);
    CHECK(s == HVD_ERR_MISMATCH && strstr(hvdcore_last_error(), "bad.size"), "mismatch -> %d (%s)", s, hvdcore_last_error());
    std::vector<float> y(8, 1.f), z(8, 1.f);
    const hvd_op_t B = HVD_BARRIER;
    if (R == 0) {   // the peers are held back by the barrier below, so the first "dup" cannot have completed yet
      h = ar("dup", y.data(), y.data(), 8, HVD_F32);
      const int h2 = ar("dup", z.data(), z.data(), 8, HVD_F32);
      CHECK(h2 == HVD_ERR_DUPLICATE, "duplicate -> %d", h2);
    }
    CHECK(hvdcore_wait(hvdcore_enqueue(B, "dup.gate", nullptr, nullptr, 0, HVD_U8, HVD_SUM, 0, 1, 1, -1, nullptr, nullptr, 0)) == 0, "gate");
    if (R != 0) h = ar("dup", y.data(), y.data(), 8, HVD_F32);
    CHECK(hvdcore_wait(h) == 0 && y[0] == (float)W, "allreduce after an error: %f", y[0]);
  }

  // ---- 6. join: rank r runs r + 1 steps, then joins; missing ranks contribute zeros ----------------------------------
  {
    for (int step = 0; step <= R; step++) {
      std::vector<float> v(33, 1.f);
      int h = ar(("join.step." + std::to_string(step)).c_str(), v.data(), v.data(), 33, HVD_F32);
      CHECK(hvdcore_wait(h) == 0, "allreduce during join: %s", hvdcore_last_error());
      CHECK(v[32] == (float)(W - step), "step %d: sum %f, want %d participants", step, v[32], W - step);
    }
    int h = hvdcore_enqueue(HVD_JOIN, nullptr, nullptr, nullptr, 0, HVD_U8, HVD_SUM, 0, 1, 1, -1, nullptr, nullptr, 0);
    const int last = hvdcore_wait(h);
    CHECK(last == W - 1, "join returned %d", last);
    std::vector<float> v(4, 2.f);   // everybody is back after the join
    h = ar("after.join", v.data(), v.data(), 4, HVD_F32);
    CHECK(hvdcore_wait(h) == 0 && v[0] == 2.f * W, "allreduce after join: %f", v[0]);
  }

  // ---- 6b. Adasum (HVD_ADASUM): orthogonal vectors add, parallel vectors average, zero is neutral; vs a double tree ---------
  {
    const int n = 1000 + 7;
    std::vector<float> e(W * 3, 0.f);
    e[R * 3] = 1.f; e[R * 3 + 1] = 2.f;                       // pairwise orthogonal across ranks
    int h = ar("adasum.orth", e.data(), e.data(), (int64_t)e.size(), HVD_F32, HVD_ADASUM);
    CHECK(hvdcore_wait(h) == 0, "adasum orth: %s", hvdcore_last_error());
    for (int r = 0; r < W; r++) CHECK(e[r * 3] == 1.f && e[r * 3 + 1] == 2.f && e[r * 3 + 2] == 0.f, "orthogonal vectors must add (slot %d: %f %f)", r, e[r * 3], e[r * 3 + 1]);
    std::vector<double> p(n);
    for (int i = 0; i < n; i++) p[i] = (double)(R + 1) * (0.25 + i % 13);     // parallel across ranks: rank r holds (r+1) * base
    h = ar("adasum.par", p.data(), p.data(), n, HVD_F64, HVD_ADASUM);
    CHECK(hvdcore_wait(h) == 0, "adasum parallel: %s", hvdcore_last_error());
    // tree of averages: level by level (a + b) / 2 over coefficients r + 1
    std::vector<double> c(W);
    for (int r = 0; r < W; r++) c[r] = r + 1;
    while (c.size() > 1) {
      std::vector<double> nx;
      for (size_t i = 0; i + 1 < c.size(); i += 2) nx.push_back((c[i] + c[i + 1]) / 2);
      if (c.size() % 2) nx.push_back(c.back());
      c.swap(nx);
    }
    for (int i = 0; i < n; i += 97) CHECK(fabs(p[i] - c[0] * (0.25 + i % 13)) < 1e-9 * (1 + fabs(p[i])), "parallel vectors must average: %g vs %g", p[i], c[0] * (0.25 + i % 13));
    // random vectors, several in flight at once with other traffic (never fused with each other), bf16 in / out
    std::vector<std::vector<float>> all(W, std::vector<float>(n));
    for (int r = 0; r < W; r++) { std::mt19937 g(4242 + r); std::normal_distribution<float> d; for (auto& x : all[r]) x = d(g); }
    std::vector<float> mine = all[R], other(64, 1.f);
    std::vector<uint16_t> half(n);
    for (int i = 0; i < n; i++) half[i] = to_bf16(all[R][i]);
    const int h1 = ar("adasum.rand", mine.data(), mine.data(), n, HVD_F32, HVD_ADASUM, 2.0, 0.5);
    const int h2 = ar("adasum.side", other.data(), other.data(), 64, HVD_F32);
    const int h3 = ar("adasum.bf16", half.data(), half.data(), n, HVD_BF16, HVD_ADASUM);
    CHECK(hvdcore_wait(h1) == 0 && hvdcore_wait(h2) == 0 && hvdcore_wait(h3) == 0, "adasum batch: %s", hvdcore_last_error());
    CHECK(other[0] == (float)W, "side allreduce %f", other[0]);
    auto tree = [&](std::vector<std::vector<double>> v) {
      while (v.size() > 1) {
        std::vector<std::vector<double>> nx;
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
          double d = 0, na = 0, nb = 0;
          for (int k = 0; k < n; k++) { d += v[i][k] * v[i + 1][k]; na += v[i][k] * v[i][k]; nb += v[i + 1][k] * v[i + 1][k]; }
          const double ca = na > 0 ? 1 - d / (2 * na) : 1, cb = nb > 0 ? 1 - d / (2 * nb) : 1;
          std::vector<double> o(n);
          for (int k = 0; k < n; k++) o[k] = ca * v[i][k] + cb * v[i + 1][k];
          nx.push_back(o);
        }
        if (v.size() % 2) nx.push_back(v.back());
        v.swap(nx);
      }
      return v[0];
    };
    std::vector<std::vector<double>> d32(W, std::vector<double>(n)), d16(W, std::vector<double>(n));
    for (int r = 0; r < W; r++) for (int k = 0; k < n; k++) { d32[r][k] = 2.0 * all[r][k]; d16[r][k] = from_bf16(to_bf16(all[r][k])); }
    const std::vector<double> w32 = tree(d32), w16 = tree(d16);
    double e32 = 0, e16 = 0;
    for (int k = 0; k < n; k++) { e32 = std::max(e32, fabs(mine[k] - 0.5 * w32[k])); e16 = std::max(e16, fabs(from_bf16(half[k]) - w16[k])); }
    CHECK(e32 < 1e-4, "adasum f32 vs double tree: max err %g", e32);
    CHECK(e16 < 5e-2, "adasum bf16 vs double tree: max err %g", e16);
    int64_t bad_i[4] = {1, 2, 3, 4};
    h = hvdcore_enqueue(HVD_ALLREDUCE, "adasum.int", bad_i, bad_i, 4, HVD_I64, HVD_ADASUM, 0, 1, 1, -1, nullptr, nullptr, 0);
    CHECK(hvdcore_wait(h) == HVD_ERR_UNSUPPORTED, "integer Adasum must be refused on every rank");
  }

  // ---- 7. agreement on the verdict, shutdown ---------------------------------------------------------------------------
  int32_t bad = g_bad, any = 0;
  int h = ar("verdict", &bad, &any, 1, HVD_I32, HVD_MAX);
  hvdcore_wait(h);
  hvdcore_shutdown();
  CHECK(hvdcore_initialized() == 0, "still initialised after shutdown");
  if (R == 0) printf(any || g_bad ? "hvd_core_test: FAILED\n" : "hvd_core_test: all checks passed on %d ranks\n", W);
  return (any || g_bad) ? 1 : 0;
}
