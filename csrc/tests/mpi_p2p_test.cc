// Point-to-point and v-collective semantics of the libmpi shim, run under the native mpirun on 4 ranks
// (`make test_mpi_p2p`, also under ASAN via `make asan`): head-to-head large sends (no deadlock), eager send before the
// receiver's first call, non-overtaking order, wildcards, nonblocking requests, probe + get_count, truncation, PROC_NULL,
// self-send, Sendrecv ring, Gatherv / Scatterv / Allgatherv with ragged counts, Scan / Exscan, Reduce_scatter_block.
#include <mpi.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

  // 9. real sub-communicators: parity split with reversed order, every collective inside it, traffic stays inside its context
  {
    MPI_Comm half;
    EXPECT(MPI_Comm_split(MPI_COMM_WORLD, r % 2, -r, &half) == MPI_SUCCESS && half != MPI_COMM_NULL);
    int hr = -1, hs = -1;
    MPI_Comm_rank(half, &hr);
    MPI_Comm_size(half, &hs);
    const int members = (n + 1 - (r % 2)) / 2;                 // ranks of my parity
    std::vector<int> mine_world;                                // world ranks of my communicator, in ITS order (descending: key = -r)
    for (int k = n - 1; k >= 0; k--) if (k % 2 == r % 2) mine_world.push_back(k);
    EXPECT(hs == members && hr >= 0 && hr < hs && mine_world[hr] == r);
    int sum = 0, want = 0;
    for (int k : mine_world) want += k + 1;
    int v = r + 1;
    EXPECT(MPI_Allreduce(&v, &sum, 1, MPI_INT, MPI_SUM, half) == MPI_SUCCESS && sum == want);
    double d[3] = {r * 1.5, 2.0, -r * 1.0}, dmax[3] = {0, 0, 0};
    EXPECT(MPI_Reduce(d, dmax, 3, MPI_DOUBLE, MPI_MAX, hs - 1, half) == MPI_SUCCESS);
    if (hr == hs - 1) EXPECT(dmax[0] == mine_world[0] * 1.5 && dmax[1] == 2.0 && dmax[2] == -mine_world[hs - 1] * 1.0);
    int token = hr == 0 ? 4242 + r : -1;
    EXPECT(MPI_Bcast(&token, 1, MPI_INT, 0, half) == MPI_SUCCESS && token == 4242 + mine_world[0]);
    std::vector<int> gathered(hs, -1);
    EXPECT(MPI_Allgather(&r, 1, MPI_INT, gathered.data(), 1, MPI_INT, half) == MPI_SUCCESS && gathered == mine_world);
    std::vector<int> a2a_in(hs), a2a_out(hs, -1);
    for (int k = 0; k < hs; k++) a2a_in[k] = 1000 * r + k;
    EXPECT(MPI_Alltoall(a2a_in.data(), 1, MPI_INT, a2a_out.data(), 1, MPI_INT, half) == MPI_SUCCESS);
    for (int k = 0; k < hs; k++) EXPECT(a2a_out[k] == 1000 * mine_world[k] + hr);
    EXPECT(MPI_Barrier(half) == MPI_SUCCESS);
    std::vector<int> root_all(hs, -1);
    MPI_Gather(&v, 1, MPI_INT, root_all.data(), 1, MPI_INT, 0, half);
    if (hr == 0) for (int k = 0; k < hs; k++) EXPECT(root_all[k] == mine_world[k] + 1);
    int piece = -1;
    std::vector<int> pieces(hs);
    for (int k = 0; k < hs; k++) pieces[k] = 7 * k + r;
    MPI_Scatter(pieces.data(), 1, MPI_INT, &piece, 1, MPI_INT, hs - 1, half);
    EXPECT(piece == 7 * hr + mine_world[hs - 1]);
    int inc = 0;
    MPI_Scan(&v, &inc, 1, MPI_INT, MPI_SUM, half);
    int pref = 0;
    for (int k = 0; k <= hr; k++) pref += mine_world[k] + 1;
    EXPECT(inc == pref);
    std::vector<int> contrib(hs, r), got1(1, -1);
    MPI_Reduce_scatter_block(contrib.data(), got1.data(), 1, MPI_INT, MPI_SUM, half);
    int wsum = 0;
    for (int k : mine_world) wsum += k;
    EXPECT(got1[0] == wsum);
    // ragged exchange inside the sub-communicator: rank i sends i + 1 copies of its world rank to everybody
    std::vector<int> sc(hs, hr + 1), sd(hs, 0), rc2(hs), rd(hs);
    int tot = 0;
    for (int k = 0; k < hs; k++) { rc2[k] = k + 1; rd[k] = tot; tot += k + 1; }
    std::vector<int> sbuf(hr + 1, r), rbuf(tot, -1);
    EXPECT(MPI_Alltoallv(sbuf.data(), sc.data(), sd.data(), MPI_INT, rbuf.data(), rc2.data(), rd.data(), MPI_INT, half) == MPI_SUCCESS);
    for (int k = 0; k < hs; k++) for (int i = 0; i <= k; i++) EXPECT(rbuf[rd[k] + i] == mine_world[k]);
    // point-to-point inside the communicator: comm ranks address it, status reports comm ranks, the same (source, tag) on
    // MPI_COMM_WORLD is a different message
    if (hs > 1) {
      const int nxt = (hr + 1) % hs, prv = (hr + hs - 1) % hs;
      int out_h = 500 + r, out_w = 900 + r, in_h = -1, in_w = -1;
      MPI_Send(&out_w, 1, MPI_INT, mine_world[nxt], 77, MPI_COMM_WORLD);   // same peer, same tag, other communicator - sent FIRST
      MPI_Send(&out_h, 1, MPI_INT, nxt, 77, half);
      MPI_Status st;
      EXPECT(MPI_Recv(&in_h, 1, MPI_INT, MPI_ANY_SOURCE, 77, half, &st) == MPI_SUCCESS && in_h == 500 + mine_world[prv] && st.MPI_SOURCE == prv);
      EXPECT(MPI_Recv(&in_w, 1, MPI_INT, mine_world[prv], 77, MPI_COMM_WORLD, &st) == MPI_SUCCESS && in_w == 900 + mine_world[prv] && st.MPI_SOURCE == mine_world[prv]);
    }
    // nested split, dup, compare, free
    MPI_Comm solo, twin;
    EXPECT(MPI_Comm_split(half, hr, 0, &solo) == MPI_SUCCESS);
    int ss = 0;
    MPI_Comm_size(solo, &ss);
    EXPECT(ss == 1);
    int alone = 5, alone_out = 0;
    EXPECT(MPI_Allreduce(&alone, &alone_out, 1, MPI_INT, MPI_SUM, solo) == MPI_SUCCESS && alone_out == 5);   // a singleton reduces to itself
    EXPECT(MPI_Allreduce(&alone, &alone_out, 1, MPI_INT, MPI_SUM, MPI_COMM_SELF) == MPI_SUCCESS && alone_out == 5);
    EXPECT(MPI_Comm_dup(half, &twin) == MPI_SUCCESS && twin != half);
    int cmp = -1;
    MPI_Comm_compare(half, twin, &cmp);
    EXPECT(cmp == MPI_CONGRUENT);
    MPI_Comm_compare(half, half, &cmp);
    EXPECT(cmp == MPI_IDENT);
    MPI_Comm_compare(half, MPI_COMM_WORLD, &cmp);
    EXPECT(cmp == (n == 1 ? MPI_CONGRUENT : MPI_UNEQUAL));
    int s2 = 0;
    EXPECT(MPI_Allreduce(&v, &s2, 1, MPI_INT, MPI_SUM, twin) == MPI_SUCCESS && s2 == want);
    MPI_Comm wdup;
    EXPECT(MPI_Comm_dup(MPI_COMM_WORLD, &wdup) == MPI_SUCCESS);
    long long big_sum = 0, mine_ll = r + 1;
    EXPECT(MPI_Allreduce(&mine_ll, &big_sum, 1, MPI_LONG_LONG, MPI_SUM, wdup) == MPI_SUCCESS && big_sum == (long long)n * (n + 1) / 2);
    MPI_Comm_free(&solo); MPI_Comm_free(&twin); MPI_Comm_free(&wdup); MPI_Comm_free(&half);
    EXPECT(half == MPI_COMM_NULL && MPI_Barrier(half) == MPI_ERR_COMM);
  }

  // 10. groups -> MPI_Comm_create; contiguous derived types; nonblocking collectives and the request family
  {
    MPI_Group world_g, ends_g;
    MPI_Comm_group(MPI_COMM_WORLD, &world_g);
    int gs = 0, gr = -1;
    MPI_Group_size(world_g, &gs);
    MPI_Group_rank(world_g, &gr);
    EXPECT(gs == n && gr == r);
    const int pick[2] = {n - 1, 0};
    MPI_Group_incl(world_g, n > 1 ? 2 : 1, pick, &ends_g);
    MPI_Comm ends;
    EXPECT(MPI_Comm_create(MPI_COMM_WORLD, ends_g, &ends) == MPI_SUCCESS);
    if (r == 0 || r == n - 1) {
      int er = -1, es = 0;
      MPI_Comm_rank(ends, &er);
      MPI_Comm_size(ends, &es);
      EXPECT(es == (n > 1 ? 2 : 1) && er == (r == n - 1 ? 0 : 1 % es));
      int x = r, mx = -1;
      EXPECT(MPI_Allreduce(&x, &mx, 1, MPI_INT, MPI_MAX, ends) == MPI_SUCCESS && mx == n - 1);
      MPI_Comm_free(&ends);
    } else {
      EXPECT(ends == MPI_COMM_NULL);
    }
    int tr[2] = {0, n - 1}, tb[2] = {-9, -9};
    MPI_Group_translate_ranks(world_g, 2, tr, ends_g, tb);
    EXPECT(tb[0] == (n > 1 ? 1 : 0) && tb[1] == 0);
    // MPI_Comm_create_group: only the members call (the others do something else meanwhile)
    if (n >= 3) {
      MPI_Group odd_g;
      std::vector<int> odd;
      for (int k = 1; k < n; k += 2) odd.push_back(k);
      MPI_Group_incl(world_g, (int)odd.size(), odd.data(), &odd_g);
      if (r % 2 == 1) {
        MPI_Comm oddc;
        EXPECT(MPI_Comm_create_group(MPI_COMM_WORLD, odd_g, 7, &oddc) == MPI_SUCCESS && oddc != MPI_COMM_NULL);
        int osz = 0, orank = -1, s = 0, one = r;
        MPI_Comm_size(oddc, &osz);
        MPI_Comm_rank(oddc, &orank);
        EXPECT(osz == (int)odd.size() && odd[orank] == r);
        EXPECT(MPI_Allreduce(&one, &s, 1, MPI_INT, MPI_SUM, oddc) == MPI_SUCCESS);
        int want = 0;
        for (int k : odd) want += k;
        EXPECT(s == want);
        MPI_Comm_free(&oddc);
      }
      MPI_Group_free(&odd_g);
      MPI_Barrier(MPI_COMM_WORLD);
    }
    MPI_Group_free(&world_g); MPI_Group_free(&ends_g);

    MPI_Datatype triple;
    EXPECT(MPI_Type_contiguous(3, MPI_DOUBLE, &triple) == MPI_SUCCESS && MPI_Type_commit(&triple) == MPI_SUCCESS);
    int tsz = 0;
    MPI_Type_size(triple, &tsz);
    EXPECT(tsz == 24);
    double pts[6] = {1.0 * r, 2.0, 3.0, 4.0, 5.0, 6.0 + r}, tot[6] = {0, 0, 0, 0, 0, 0};
    EXPECT(MPI_Allreduce(pts, tot, 2, triple, MPI_SUM, MPI_COMM_WORLD) == MPI_SUCCESS);
    EXPECT(tot[0] == n * (n - 1) / 2.0 && tot[1] == 2.0 * n && tot[5] == 6.0 * n + n * (n - 1) / 2.0);
    double moved[3] = {-1, -1, -1};
    MPI_Sendrecv(pts, 1, triple, (r + 1) % n, 90, moved, 1, triple, (r + n - 1) % n, 90, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    EXPECT(moved[0] == 1.0 * ((r + n - 1) % n) && moved[2] == 3.0);
    MPI_Type_free(&triple);

    MPI_Request rq[3];
    int bsum = 0, one = 1, bval = r == 0 ? 31337 : 0;
    EXPECT(MPI_Iallreduce(&one, &bsum, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD, &rq[0]) == MPI_SUCCESS);
    EXPECT(MPI_Ibcast(&bval, 1, MPI_INT, 0, MPI_COMM_WORLD, &rq[1]) == MPI_SUCCESS);
    EXPECT(MPI_Ibarrier(MPI_COMM_WORLD, &rq[2]) == MPI_SUCCESS);
    int all_done = 0;
    EXPECT(MPI_Testall(3, rq, &all_done, MPI_STATUSES_IGNORE) == MPI_SUCCESS && all_done == 1 && bsum == n && bval == 31337);
    // Waitany / Waitsome / Testany over receives that complete in a known order
    int in3[3] = {-1, -1, -1};
    MPI_Request rr[3];
    for (int k = 0; k < 3; k++) MPI_Irecv(&in3[k], 1, MPI_INT, r, 200 + k, MPI_COMM_WORLD, &rr[k]);
    int idx = -1, flag = -1;
    EXPECT(MPI_Testany(3, rr, &idx, &flag, MPI_STATUS_IGNORE) == MPI_SUCCESS && flag == 0);   // nothing sent yet
    int payload = 61;
    MPI_Send(&payload, 1, MPI_INT, r, 201, MPI_COMM_WORLD);
    EXPECT(MPI_Waitany(3, rr, &idx, MPI_STATUS_IGNORE) == MPI_SUCCESS && idx == 1 && in3[1] == 61 && rr[1] == MPI_REQUEST_NULL);
    payload = 62;
    MPI_Send(&payload, 1, MPI_INT, r, 200, MPI_COMM_WORLD);
    payload = 63;
    MPI_Send(&payload, 1, MPI_INT, r, 202, MPI_COMM_WORLD);
    int outc = 0, which[3] = {-1, -1, -1}, seen = 0;
    while (seen < 2) { EXPECT(MPI_Waitsome(3, rr, &outc, which, MPI_STATUSES_IGNORE) == MPI_SUCCESS && outc >= 1); seen += outc; }
    EXPECT(in3[0] == 62 && in3[2] == 63);
    EXPECT(MPI_Waitsome(3, rr, &outc, which, MPI_STATUSES_IGNORE) == MPI_SUCCESS && outc == MPI_UNDEFINED);   // every request is null now
    int vec[4] = {r, r, r, r}, part1 = -1;
    std::vector<int> cnts(n, 0);
    cnts[r] = 0;
    for (int k = 0; k < n; k++) cnts[k] = k == 0 ? 1 : 0;    // everything reduced lands on rank 0
    int onev = r + 1;
    MPI_Reduce_scatter(&onev, &part1, cnts.data(), MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    if (r == 0) EXPECT(part1 == n * (n + 1) / 2);
    (void)vec;
    MPI_Comm_set_errhandler(MPI_COMM_WORLD, MPI_ERRORS_RETURN);
    char nm[MPI_MAX_OBJECT_NAME];
    int nl = 0;
    MPI_Comm_get_name(MPI_COMM_WORLD, nm, &nl);
    EXPECT(std::string(nm) == "MPI_COMM_WORLD");
  }

  // 11. MPI_MAXLOC / MPI_MINLOC on value-index pairs, user-defined (non-commutative) reductions in rank order
  {
    struct DI { double v; int i; } mine_di[2] = {{(double)((r * 7) % 5), r}, {-1.0 * r, r}}, best[2];
    EXPECT(MPI_Allreduce(mine_di, best, 2, MPI_DOUBLE_INT, MPI_MAXLOC, MPI_COMM_WORLD) == MPI_SUCCESS);
    double top = -1; int who = -1;
    for (int k = 0; k < n; k++) if ((k * 7) % 5 > top) { top = (k * 7) % 5; who = k; }       // ties keep the lower rank
    EXPECT(best[0].v == top && best[0].i == who && best[1].v == 0.0 && best[1].i == 0);
    struct II { int v; int i; } mi = {100 - r, r}, lo;
    EXPECT(MPI_Reduce(&mi, &lo, 1, MPI_2INT, MPI_MINLOC, 0, MPI_COMM_WORLD) == MPI_SUCCESS);
    if (r == 0) EXPECT(lo.v == 100 - (n - 1) && lo.i == n - 1);
    int sz = 0;
    MPI_Type_size(MPI_DOUBLE_INT, &sz);
    EXPECT(sz == (int)sizeof(DI));
    MPI_Op horner;
    EXPECT(MPI_Op_create([](void* in, void* inout, int* len, MPI_Datatype*) {
      long long* a = (long long*)in; long long* b = (long long*)inout;
      for (int k = 0; k < *len; k++) b[k] = a[k] * 3 + b[k];                                    // NOT commutative: order matters
    }, 0, &horner) == MPI_SUCCESS);
    std::vector<long long> x(1000), y(1000, -1);
    for (int k = 0; k < 1000; k++) x[k] = r + 1 + k;
    EXPECT(MPI_Allreduce(x.data(), y.data(), 1000, MPI_LONG_LONG, horner, MPI_COMM_WORLD) == MPI_SUCCESS);
    for (int k = 0; k < 1000; k += 111) {
      long long want = 1 + k;
      for (int q = 1; q < n; q++) want = want * 3 + (q + 1 + k);
      EXPECT(y[k] == want);
    }
    MPI_Comm odd_even;
    MPI_Comm_split(MPI_COMM_WORLD, r % 2, r, &odd_even);
    long long one = r + 1, folded = -1, want = -1;
    EXPECT(MPI_Allreduce(&one, &folded, 1, MPI_LONG_LONG, horner, odd_even) == MPI_SUCCESS);    // point-to-point based path
    for (int q = r % 2; q < n; q += 2) want = want < 0 ? q + 1 : want * 3 + (q + 1);
    EXPECT(folded == want);
    long long pre = -1;
    EXPECT(MPI_Scan(&one, &pre, 1, MPI_LONG_LONG, horner, MPI_COMM_WORLD) == MPI_SUCCESS);
    want = 1;
    for (int q = 1; q <= r; q++) want = want * 3 + (q + 1);
    EXPECT(pre == want);
    MPI_Comm_free(&odd_even);
    MPI_Op_free(&horner);
    EXPECT(horner == MPI_OP_NULL && MPI_Allreduce(&one, &folded, 1, MPI_LONG_LONG, 100, MPI_COMM_WORLD) != MPI_SUCCESS);
  }

  int any = 0;
  MPI_Allreduce(&g_bad, &any, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  if (r == 0) printf(any ? "mpi_p2p_test: FAILED\n" : "mpi_p2p_test: all checks passed on %d ranks\n", n);
  MPI_Finalize();
  return any ? 1 : 0;
}
