// Point-to-point and v-collective semantics of the libmpi shim, run under the native mpirun on 4 ranks
// (`make test_mpi_p2p`, also under ASAN via `make asan`): head-to-head large sends (no deadlock), eager send before the
// receiver's first call, non-overtaking order, wildcards, nonblocking requests, probe + get_count, truncation, PROC_NULL,
// self-send, Sendrecv ring, Gatherv / Scatterv / Allgatherv with ragged counts, Scan / Exscan, Reduce_scatter_block.
#include <mpi.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static int g_bad = 0, g_rank = 0;
#define EXPECT(c)                                                                                   \
  do {                                                                                              \
    if (!(c)) { printf("rank %d FAILED %s:%d: %s\n", g_rank, __FILE__, __LINE__, #c); g_bad++; }    \
  } while (0)

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int r, n;
  MPI_Comm_rank(MPI_COMM_WORLD, &r);
  MPI_Comm_size(MPI_COMM_WORLD, &n);
  g_rank = r;
  const int nxt = (r + 1) % n, prv = (r + n - 1) % n;

  // 1. eager send before the receiver's first point-to-point call: A sends then enters the barrier, B leaves the barrier then receives
  if (n >= 2) {
    int v = 4711;
    if (r == 0) MPI_Send(&v, 1, MPI_INT, 1, 5, MPI_COMM_WORLD);
    MPI_Barrier(MPI_COMM_WORLD);
    if (r == 1) { int got = 0; MPI_Status st; MPI_Recv(&got, 1, MPI_INT, 0, 5, MPI_COMM_WORLD, &st); EXPECT(got == 4711 && st.MPI_SOURCE == 0 && st.MPI_TAG == 5); }
  }

  // 2. head-to-head 8 MiB sends in a ring: everybody sends first, then receives (needs the progress-while-sending path)
  {
    const size_t N = 2u << 20;  // ints -> 8 MiB
    std::vector<int> out(N), in(N, -1);
    for (size_t i = 0; i < N; i++) out[i] = (int)(i * 31 + r);
    MPI_Send(out.data(), (int)N, MPI_INT, nxt, 1, MPI_COMM_WORLD);
    MPI_Status st;
    MPI_Recv(in.data(), (int)N, MPI_INT, prv, 1, MPI_COMM_WORLD, &st);
    int cnt = 0;
    MPI_Get_count(&st, MPI_INT, &cnt);
    EXPECT(cnt == (int)N && in[0] == prv && in[N - 1] == (int)((N - 1) * 31 + prv) && in[N / 2] == (int)((N / 2) * 31 + prv));
  }

  // 3. non-overtaking: three messages with the same tag arrive in send order; a different tag can be picked out of order
  {
    for (int k = 0; k < 3; k++) { int v = 100 * r + k; MPI_Send(&v, 1, MPI_INT, nxt, 7, MPI_COMM_WORLD); }
    int other = 900 + r;
    MPI_Send(&other, 1, MPI_INT, nxt, 8, MPI_COMM_WORLD);
    int got = 0;
    MPI_Recv(&got, 1, MPI_INT, prv, 8, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    EXPECT(got == 900 + prv);
    for (int k = 0; k < 3; k++) { MPI_Recv(&got, 1, MPI_INT, prv, 7, MPI_COMM_WORLD, MPI_STATUS_IGNORE); EXPECT(got == 100 * prv + k); }
  }

  // 4. wildcards: rank 0 collects one message from everybody with ANY_SOURCE / ANY_TAG
  {
    if (r != 0) { double v = r * 1.5; MPI_Send(&v, 1, MPI_DOUBLE, 0, 20 + r, MPI_COMM_WORLD); }
    else {
      std::vector<int> seen(n, 0);
      for (int k = 1; k < n; k++) {
        double v = 0; MPI_Status st;
        MPI_Recv(&v, 1, MPI_DOUBLE, MPI_ANY_SOURCE, MPI_ANY_TAG, MPI_COMM_WORLD, &st);
        EXPECT(st.MPI_TAG == 20 + st.MPI_SOURCE && v == st.MPI_SOURCE * 1.5);
        seen[st.MPI_SOURCE]++;
      }
      for (int k = 1; k < n; k++) EXPECT(seen[k] == 1);
    }
    MPI_Barrier(MPI_COMM_WORLD);   // keep the wildcard receives from seeing the next phase's messages
  }

  // 5. nonblocking ring + Test / Waitall, probe before receive, truncation, PROC_NULL, self-send
  {
    long long sendv[2] = {r * 10LL, r * 10LL + 1}, recvv[2] = {-1, -1};
    MPI_Request q[2];
    MPI_Irecv(recvv, 2, MPI_LONG_LONG, prv, 30, MPI_COMM_WORLD, &q[0]);
    MPI_Isend(sendv, 2, MPI_LONG_LONG, nxt, 30, MPI_COMM_WORLD, &q[1]);
    MPI_Status sts[2];
    EXPECT(MPI_Waitall(2, q, sts) == MPI_SUCCESS && q[0] == MPI_REQUEST_NULL);
    EXPECT(recvv[0] == prv * 10LL && recvv[1] == prv * 10LL + 1 && sts[0].MPI_SOURCE == prv);

    char text[32];
    snprintf(text, sizeof(text), "hello from %d", r);
    MPI_Send(text, (int)strlen(text) + 1, MPI_CHAR, nxt, 31, MPI_COMM_WORLD);
    MPI_Status st;
    MPI_Probe(prv, 31, MPI_COMM_WORLD, &st);
    int len = 0;
    MPI_Get_count(&st, MPI_CHAR, &len);
    std::vector<char> buf(len);
    MPI_Recv(buf.data(), len, MPI_CHAR, prv, 31, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    char want[32];
    snprintf(want, sizeof(want), "hello from %d", prv);
    EXPECT(len == (int)strlen(want) + 1 && strcmp(buf.data(), want) == 0);

    int flag = 1;
    MPI_Iprobe(prv, 99, MPI_COMM_WORLD, &flag, &st);
    EXPECT(flag == 0);                                           // nothing with that tag
    int big[4] = {1, 2, 3, 4}, small[2] = {0, 0};
    MPI_Send(big, 4, MPI_INT, nxt, 32, MPI_COMM_WORLD);
    EXPECT(MPI_Recv(small, 2, MPI_INT, prv, 32, MPI_COMM_WORLD, &st) == MPI_ERR_TRUNCATE && small[0] == 1 && small[1] == 2);
    EXPECT(MPI_Send(big, 4, MPI_INT, MPI_PROC_NULL, 0, MPI_COMM_WORLD) == MPI_SUCCESS);
    EXPECT(MPI_Recv(big, 4, MPI_INT, MPI_PROC_NULL, 0, MPI_COMM_WORLD, &st) == MPI_SUCCESS && st.MPI_SOURCE == MPI_PROC_NULL);
    EXPECT(MPI_Send(big, 1, MPI_INT, n, 0, MPI_COMM_WORLD) == MPI_ERR_RANK);
    int me = 77 + r, back = 0;
    MPI_Send(&me, 1, MPI_INT, r, 33, MPI_COMM_WORLD);            // to self: buffered locally
    MPI_Recv(&back, 1, MPI_INT, r, 33, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    EXPECT(back == 77 + r);
    float a = (float)r, b = -1.f;
    MPI_Sendrecv(&a, 1, MPI_FLOAT, nxt, 34, &b, 1, MPI_FLOAT, prv, 34, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    EXPECT(b == (float)prv);
  }

  // 6. ragged collectives: rank k contributes k + 1 elements
  {
    std::vector<int> counts(n), displs(n);
    int total = 0;
    for (int k = 0; k < n; k++) { counts[k] = k + 1; displs[k] = total; total += k + 1; }
    std::vector<int> mine(r + 1, r * 100), all(total, -1);
    MPI_Allgatherv(mine.data(), r + 1, MPI_INT, all.data(), counts.data(), displs.data(), MPI_INT, MPI_COMM_WORLD);
    for (int k = 0; k < n; k++) for (int i = 0; i < counts[k]; i++) EXPECT(all[displs[k] + i] == k * 100);
    std::vector<int> root_buf(total, -1);
    const int root = n - 1;
    MPI_Gatherv(mine.data(), r + 1, MPI_INT, root_buf.data(), counts.data(), displs.data(), MPI_INT, root, MPI_COMM_WORLD);
    if (r == root) for (int k = 0; k < n; k++) for (int i = 0; i < counts[k]; i++) EXPECT(root_buf[displs[k] + i] == k * 100);
    std::vector<int> src(total), part(r + 1, -1);
    for (int i = 0; i < total; i++) src[i] = i * 3;
    MPI_Scatterv(src.data(), counts.data(), displs.data(), MPI_INT, part.data(), r + 1, MPI_INT, 0, MPI_COMM_WORLD);
    for (int i = 0; i <= r; i++) EXPECT(part[i] == (displs[r] + i) * 3);
  }

  // 7. scans and reduce-scatter
  {
    int v[2] = {r + 1, 2 * (r + 1)}, inc[2] = {0, 0}, exc[2] = {-1, -1};
    MPI_Scan(v, inc, 2, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    EXPECT(inc[0] == (r + 1) * (r + 2) / 2 && inc[1] == (r + 1) * (r + 2));
    MPI_Exscan(v, exc, 2, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    if (r > 0) EXPECT(exc[0] == r * (r + 1) / 2 && exc[1] == r * (r + 1));
    int mx = 0;
    int mine = (r * 7) % 5;
    MPI_Scan(&mine, &mx, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    int want = 0;
    for (int k = 0; k <= r; k++) want = (k * 7) % 5 > want ? (k * 7) % 5 : want;
    EXPECT(mx == want);
    std::vector<double> contrib(2 * n), mineout(2, -1);
    for (int i = 0; i < 2 * n; i++) contrib[i] = i + r * 0.5;
    MPI_Reduce_scatter_block(contrib.data(), mineout.data(), 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    for (int i = 0; i < 2; i++) EXPECT(mineout[i] == n * (2.0 * r + i) + 0.5 * n * (n - 1) / 2.0);
  }

  // 8. communicator splits that one box can express
  {
    MPI_Comm shared, self, none;
    EXPECT(MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, r, MPI_INFO_NULL, &shared) == MPI_SUCCESS);
    int lr = -1, ls = -1;
    MPI_Comm_rank(shared, &lr);
    MPI_Comm_size(shared, &ls);
    EXPECT(lr == r && ls == n);                                  // "local rank / local size" the way Horovod derives them
    EXPECT(MPI_Comm_split(MPI_COMM_WORLD, r, 0, &self) == MPI_SUCCESS);
    MPI_Comm_size(self, &ls);
    EXPECT(ls == 1);
    EXPECT(MPI_Comm_split(MPI_COMM_WORLD, MPI_UNDEFINED, 0, &none) == MPI_SUCCESS && none == MPI_COMM_NULL);
  }

  int any = 0;
  MPI_Allreduce(&g_bad, &any, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  if (r == 0) printf(any ? "mpi_p2p_test: FAILED\n" : "mpi_p2p_test: all checks passed on %d ranks\n", n);
  MPI_Finalize();
  return any ? 1 : 0;
}
