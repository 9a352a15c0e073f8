// Stress program for the host-side runtime under ASAN/UBSAN/TSAN (`make asan`, `make tsan`; SURVEY.md §5.2 [NEW]:
// the reference runs `go test` without -race, Makefile:71; here the shared-memory rendezvous, the mailbox
// collectives of the libmpi shim and the launcher's stdio mux are exercised under the sanitizers).
// Every rank runs `iters` rounds of randomly sized Allreduce / Bcast / Allgather / Alltoall / Barrier and checks the
// results against closed forms; every fourth round the same happens inside a freshly split sub-communicator (random colours
// and keys, collectives over the point-to-point layer, interleaved with world traffic on the same ranks) that is freed again;
// exits non-zero on the first mismatch.
#include <mpi.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank = 0, world = 1;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &world);
  const int iters = argc > 1 ? atoi(argv[1]) : 200;
  std::minstd_rand rng(12345);  // same sequence on every rank: sizes agree without communication
  int bad = 0;
  for (int it = 0; it < iters && !bad; it++) {
    // every eighth round is larger than one 64 KiB rendezvous mailbox so the chunked reduce / bcast paths run too
    const int n = it % 8 == 7 ? 8192 + (int)(rng() % 20000) : 1 + (int)(rng() % 3000);
    const int root = (int)(rng() % world);
    // allreduce: sum_r (r + i) = world*i + world*(world-1)/2
    std::vector<long long> a(n), b(n);
    for (int i = 0; i < n; i++) a[i] = rank + i;
    MPI_Allreduce(a.data(), b.data(), n, MPI_LONG_LONG, MPI_SUM, MPI_COMM_WORLD);
    for (int i = 0; i < n; i++) if (b[i] != (long long)world * i + (long long)world * (world - 1) / 2) bad = 1;
    // bcast from a moving root
    std::vector<int> c(n, rank == root ? it * 7 + 1 : -1);
    MPI_Bcast(c.data(), n, MPI_INT, root, MPI_COMM_WORLD);
    for (int i = 0; i < n; i++) if (c[i] != it * 7 + 1) bad = 2;
    // allgather
    const int m = 1 + n % 64;
    std::vector<double> s(m, rank + 0.5), g((size_t)m * world);
    MPI_Allgather(s.data(), m, MPI_DOUBLE, g.data(), m, MPI_DOUBLE, MPI_COMM_WORLD);
    for (int r = 0; r < world; r++) for (int i = 0; i < m; i++) if (g[(size_t)r * m + i] != r + 0.5) bad = 3;
    // alltoall: rank r sends (r*1000 + dst) to dst
    std::vector<int> src((size_t)m * world), dst((size_t)m * world);
    for (int d = 0; d < world; d++) for (int i = 0; i < m; i++) src[(size_t)d * m + i] = rank * 1000 + d;
    MPI_Alltoall(src.data(), m, MPI_INT, dst.data(), m, MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < world; r++) for (int i = 0; i < m; i++) if (dst[(size_t)r * m + i] != r * 1000 + rank) bad = 4;
    // reduce (max) to root
    float mine = (float)(rank * 3 + it), top = -1.f;
    MPI_Reduce(&mine, &top, 1, MPI_FLOAT, MPI_MAX, root, MPI_COMM_WORLD);
    if (rank == root && top != (float)((world - 1) * 3 + it)) bad = 5;
    if (it % 4 == 3) {
      // sub-communicators: colours / keys come from the shared generator, so every rank can compute every group's membership
      const int ncol = 1 + (int)(rng() % 3);
      std::vector<int> col(world), key(world);
      for (int r = 0; r < world; r++) { col[r] = (int)(rng() % ncol); key[r] = (int)(rng() % 5); }
      MPI_Comm sub;
      if (MPI_Comm_split(MPI_COMM_WORLD, col[rank], key[rank], &sub) != MPI_SUCCESS) bad = 10;
      std::vector<int> members;   // world ranks of my group in ITS order: by key, ties by world rank
      for (int k = 0; k < 5; k++) for (int r = 0; r < world; r++) if (col[r] == col[rank] && key[r] == k) members.push_back(r);
      int sr = -1, ss = 0;
      MPI_Comm_rank(sub, &sr);
      MPI_Comm_size(sub, &ss);
      if (ss != (int)members.size() || members[sr] != rank) bad = 11;
      const int k = 1 + n % 97;
      std::vector<long long> x(k), y(k);
      long long want = 0;
      for (int r : members) want += r + 1;
      for (int i = 0; i < k; i++) x[i] = (long long)(rank + 1) * (i + 1);
      MPI_Allreduce(x.data(), y.data(), k, MPI_LONG_LONG, MPI_SUM, sub);
      for (int i = 0; i < k; i++) if (y[i] != want * (i + 1)) bad = 12;
      const int sroot = (int)(rng() % 64) % ss;
      std::vector<int> token(k, sr == sroot ? 1000 + it : -1);
      MPI_Bcast(token.data(), k, MPI_INT, sroot, sub);
      for (int i = 0; i < k; i++) if (token[i] != 1000 + it) bad = 13;
      std::vector<int> who(ss, -1);
      MPI_Allgather(&rank, 1, MPI_INT, who.data(), 1, MPI_INT, sub);
      if (who != members) bad = 14;
      int wsum = 0, one = 1;                       // world traffic between the sub-communicator's collectives
      MPI_Allreduce(&one, &wsum, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
      if (wsum != world) bad = 15;
      std::vector<int> a2a_in(ss), a2a_out(ss, -1);
      for (int d = 0; d < ss; d++) a2a_in[d] = rank * 100 + d;
      MPI_Alltoall(a2a_in.data(), 1, MPI_INT, a2a_out.data(), 1, MPI_INT, sub);
      for (int r = 0; r < ss; r++) if (a2a_out[r] != members[r] * 100 + sr) bad = 16;
      MPI_Barrier(sub);
      MPI_Comm_free(&sub);
    }
    if (it % 16 == 0) MPI_Barrier(MPI_COMM_WORLD);
    if (it % 50 == 0) { printf("rank %d iteration %d ok\n", rank, it); fflush(stdout); }
  }
  int any = 0;
  MPI_Allreduce(&bad, &any, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  if (rank == 0) printf(any ? "mpi_stress: FAILED (check %d)\n" : "mpi_stress: all %d iterations verified on %d ranks\n", any ? any : iters, world);
  MPI_Finalize();
  return any ? 1 : 0;
}
