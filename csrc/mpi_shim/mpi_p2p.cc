// libmpi (b200mpi shim): point-to-point messaging and the collectives built on it.
//
// The reference's MPI images carry a full Open MPI / MPICH / Intel MPI (build/base/*.Dockerfile); its own example only needs
// the collectives of mpi_shim.cc, but MPI programs people launch with an MPIJob routinely use Send/Recv, nonblocking
// requests and the v-collectives. Transport on one box: every rank binds an abstract UNIX datagram socket
// ("b200mpi-<nonce>-msg-<rank>", nonce from the job's rendezvous segment); a message is a sequence of <= 60 KiB datagrams
// {source, tag, communicator, message id, total bytes, offset}. UNIX datagrams are reliable and ordered per sender, which gives
// MPI's non-overtaking rule for free. Sends are eager: data is buffered at the receiver until a matching receive is posted
// (the unexpected queue), and a sender that finds the peer's socket queue full drains its OWN socket while it waits, so two
// ranks sending to each other cannot deadlock. Everything is polled from the calling thread (MPI_THREAD_SERIALIZED).
#include <errno.h>
#include <poll.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <deque>
#include <map>
#include <string>
#include <vector>

#include "mpi_internal.h"

using namespace b200mpi_mpi;

namespace {

constexpr uint32_t kMsgMagic = 0xB2005E9Du;
constexpr size_t kChunk = 60 * 1024;

struct Header {
  uint32_t magic;
  int32_t src, tag, comm;
  uint64_t msg_id, total, offset;
  uint32_t bytes;
  uint32_t pad;
};

struct Message {            // an incoming message, complete or still arriving
  int src = 0, tag = 0, comm = 0;
  uint64_t id = 0, total = 0, have = 0;
  std::vector<unsigned char> data;
  bool complete() const { return have == total; }
};

struct Request {
  bool active = false, is_recv = false, done = false;
  void* buf = nullptr;
  size_t cap = 0;
  int src = 0, tag = 0, comm = 0;
  MPI_Status st{};
  int rc = MPI_SUCCESS;
};

int g_sock = -1;
uint64_t g_next_id = 1;
std::deque<Message> g_inbox;          // arrival order (by first chunk) == matching order per source
std::vector<Request> g_reqs;

std::string sock_name(int rank) {
  char buf[96];
  snprintf(buf, sizeof(buf), "b200mpi-%016llx-msg-%d", (unsigned long long)(g_rv ? g_rv->header()->nonce : 0ull), rank);
  return buf;
}
socklen_t fill_addr(sockaddr_un* a, int rank) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  const std::string n = sock_name(rank);
  memcpy(a->sun_path + 1, n.data(), n.size());   // abstract namespace: leading NUL, nothing on the filesystem
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n.size());
}

int ensure_socket() {
  if (g_sock >= 0 || g_size == 1) return MPI_SUCCESS;   // a single rank only ever sends to itself (inbox, no socket)
  if (!g_rv) return fail("point-to-point before MPI_Init");
  g_sock = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC | SOCK_NONBLOCK, 0);
  if (g_sock < 0) return fail(std::string("socket: ") + strerror(errno));
  int sz = 4 << 20;
  setsockopt(g_sock, SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
  setsockopt(g_sock, SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
  sockaddr_un a;
  const socklen_t len = fill_addr(&a, g_rank);
  if (bind(g_sock, (sockaddr*)&a, len) != 0) {
    const std::string e = std::string("bind message socket: ") + strerror(errno);
    close(g_sock);
    g_sock = -1;
    return fail(e);
  }
  return MPI_SUCCESS;
}

// Pull every datagram that is waiting on our socket into the inbox. Returns the number of datagrams consumed.
int drain() {
  static std::vector<unsigned char> pkt(sizeof(Header) + kChunk);
  int n = 0;
  if (g_sock < 0) return 0;
  for (;;) {
    const ssize_t got = recv(g_sock, pkt.data(), pkt.size(), 0);
    if (got < 0) {
      if (errno == EINTR) continue;
      break;  // EAGAIN: nothing more
    }
    if ((size_t)got < sizeof(Header)) continue;
    Header h;
    memcpy(&h, pkt.data(), sizeof(h));
    if (h.magic != kMsgMagic || (size_t)got != sizeof(Header) + h.bytes) continue;
    Message* m = nullptr;
    for (auto& x : g_inbox)
      if (x.src == h.src && x.id == h.msg_id) { m = &x; break; }
    if (!m) {
      g_inbox.emplace_back();
      m = &g_inbox.back();
      m->src = h.src; m->tag = h.tag; m->comm = h.comm; m->id = h.msg_id; m->total = h.total;
      m->data.resize(h.total);
    }
    if (h.offset + h.bytes <= m->total) {
      memcpy(m->data.data() + h.offset, pkt.data() + sizeof(Header), h.bytes);
      m->have += h.bytes;
    }
    n++;
  }
  return n;
}

void wait_readable(int ms) {
  if (g_sock < 0) { usleep(1000); return; }
  pollfd p{g_sock, POLLIN, 0};
  poll(&p, 1, ms);
}

bool timed_out(const timespec& t0) {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (t.tv_sec - t0.tv_sec) * 1000ll + (t.tv_nsec - t0.tv_nsec) / 1000000ll > g_timeout_ms;
}

int send_bytes(const void* buf, size_t bytes, int dest, int tag, int comm) {
  if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
  if (dest < 0 || dest >= g_size) return MPI_ERR_RANK;
  if (tag < 0) return MPI_ERR_TAG;
  int rc = ensure_socket();
  if (rc) return rc;
  const uint64_t id = g_next_id++;
  if (dest == g_rank) {  // self-send: straight into the inbox
    g_inbox.emplace_back();
    Message& m = g_inbox.back();
    m.src = g_rank; m.tag = tag; m.comm = comm; m.id = id; m.total = m.have = bytes;
    m.data.assign((const unsigned char*)buf, (const unsigned char*)buf + bytes);
    return MPI_SUCCESS;
  }
  sockaddr_un a;
  const socklen_t alen = fill_addr(&a, dest);
  static std::vector<unsigned char> pkt(sizeof(Header) + kChunk);
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  size_t off = 0;
  do {
    const size_t n = std::min(kChunk, bytes - off);
    Header h{kMsgMagic, g_rank, tag, comm, id, bytes, off, (uint32_t)n, 0};
    memcpy(pkt.data(), &h, sizeof(h));
    if (n) memcpy(pkt.data() + sizeof(h), (const unsigned char*)buf + off, n);
    for (;;) {
      if (sendto(g_sock, pkt.data(), sizeof(h) + n, 0, (sockaddr*)&a, alen) >= 0) break;
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK || errno == ENOBUFS || errno == ECONNREFUSED || errno == ENOENT) {
        // peer queue full (or peer not bound yet): make progress on our own inbox so that two ranks sending to each
        // other always drain, then retry
        if (drain() == 0) usleep(200);
        if (g_rv && g_rv->aborted()) return fail("job aborted while sending");
        if (timed_out(t0)) return fail("MPI_Send: timed out delivering to rank " + std::to_string(dest));
        continue;
      }
      return fail(std::string("sendto: ") + strerror(errno));
    }
    off += n;
  } while (off < bytes);
  return MPI_SUCCESS;
}

// First message in arrival order that matches (source, tag, comm); nullptr if none has started arriving.
Message* find_match(int src, int tag, int comm) {
  for (auto& m : g_inbox)
    if (m.comm == comm && (src == MPI_ANY_SOURCE || m.src == src) &&
        (tag == MPI_ANY_TAG ? m.tag <= MPI_TAG_UB : m.tag == tag))   // wildcards never see the collectives' internal tags
      return &m;
  return nullptr;
}
void erase_message(Message* m) {
  for (auto it = g_inbox.begin(); it != g_inbox.end(); ++it)
    if (&*it == m) { g_inbox.erase(it); return; }
}

// Blocks until a matching message is complete; copies it out. `probe_only` leaves it queued.
int recv_bytes(void* buf, size_t cap, int src, int tag, int comm, MPI_Status* st, bool probe_only, bool blocking, int* flag) {
  if (flag) *flag = 0;
  if (src == MPI_PROC_NULL) {
    if (st) { st->MPI_SOURCE = MPI_PROC_NULL; st->MPI_TAG = MPI_ANY_TAG; st->MPI_ERROR = MPI_SUCCESS; st->count_ = 0; }
    if (flag) *flag = 1;
    return MPI_SUCCESS;
  }
  if (src != MPI_ANY_SOURCE && (src < 0 || src >= g_size)) return MPI_ERR_RANK;
  int rc = ensure_socket();
  if (rc) return rc;
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (;;) {
    drain();
    Message* m = find_match(src, tag, comm);
    if (m && m->complete()) {
      if (st) { st->MPI_SOURCE = m->src; st->MPI_TAG = m->tag; st->MPI_ERROR = MPI_SUCCESS; st->count_ = (int)m->total; }
      if (flag) *flag = 1;
      if (probe_only) return MPI_SUCCESS;
      int out = MPI_SUCCESS;
      size_t n = m->total;
      if (n > cap) { n = cap; out = MPI_ERR_TRUNCATE; if (st) st->MPI_ERROR = out; }
      if (n) memcpy(buf, m->data.data(), n);
      erase_message(m);
      return out;
    }
    if (!blocking) return MPI_SUCCESS;
    if (g_rv && g_rv->aborted()) return fail("job aborted while receiving");
    if (timed_out(t0)) return fail("MPI_Recv: timed out waiting for rank " + std::to_string(src));
    wait_readable(50);
  }
}

int new_request() {
  for (size_t i = 0; i < g_reqs.size(); i++)
    if (!g_reqs[i].active) { g_reqs[i] = Request{}; g_reqs[i].active = true; return (int)i; }
  g_reqs.emplace_back();
  g_reqs.back().active = true;
  return (int)g_reqs.size() - 1;
}

int complete_request(int id, bool blocking, int* flag, MPI_Status* st) {
  if (flag) *flag = 1;
  if (id == MPI_REQUEST_NULL) return MPI_SUCCESS;
  if (id < 0 || id >= (int)g_reqs.size() || !g_reqs[id].active) return MPI_ERR_REQUEST;
  Request& r = g_reqs[id];
  if (!r.done && r.is_recv) {
    int got = 0;
    r.rc = recv_bytes(r.buf, r.cap, r.src, r.tag, r.comm, &r.st, false, blocking, &got);
    if (r.rc != MPI_SUCCESS || got) r.done = true;
  }
  if (!r.done) { if (flag) *flag = 0; return MPI_SUCCESS; }
  if (st) *st = r.st;
  const int rc = r.rc;
  r.active = false;
  return rc;
}

}  // namespace

namespace b200mpi_mpi {
// Bound in MPI_Init on every rank: an eager send must be deliverable before the receiver has made its first
// point-to-point call (rank A: Send, Barrier; rank B: Barrier, Recv must not deadlock).
int p2p_init() { return ensure_socket(); }
void p2p_shutdown() {
  if (g_sock >= 0) { close(g_sock); g_sock = -1; }
  g_inbox.clear();
  g_reqs.clear();
}
}  // namespace b200mpi_mpi

extern "C" {

int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es || count < 0) return MPI_ERR_TYPE;
  return send_bytes(buf, es * (size_t)count, c == MPI_COMM_SELF && dest == 0 ? g_rank : dest, tag, c);
}
int MPI_Ssend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) { return MPI_Send(buf, count, t, dest, tag, c); }

int MPI_Recv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm c, MPI_Status* st) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es || count < 0) return MPI_ERR_TYPE;
  return recv_bytes(buf, es * (size_t)count, c == MPI_COMM_SELF && source == 0 ? g_rank : source, tag, c, st, false, true, nullptr);
}

int MPI_Sendrecv(const void* sb, int sc, MPI_Datatype stype, int dest, int stag, void* rb, int rc_, MPI_Datatype rtype, int source, int rtag,
                 MPI_Comm c, MPI_Status* st) {
  int e = MPI_Send(sb, sc, stype, dest, stag, c);   // eager: never waits for the matching receive
  if (e) return e;
  return MPI_Recv(rb, rc_, rtype, source, rtag, c, st);
}

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req) {
  const int id = new_request();
  g_reqs[id].rc = MPI_Send(buf, count, t, dest, tag, c);  // eager send completes here; the request only carries the result
  g_reqs[id].done = true;
  *req = id;
  return g_reqs[id].rc;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm c, MPI_Request* req) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es || count < 0) return MPI_ERR_TYPE;
  const int id = new_request();
  Request& r = g_reqs[id];
  r.is_recv = true; r.buf = buf; r.cap = es * (size_t)count; r.src = source; r.tag = tag; r.comm = c;
  *req = id;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* req, MPI_Status* st) {
  const int rc = complete_request(*req, true, nullptr, st);
  *req = MPI_REQUEST_NULL;
  return rc;
}
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* sts) {
  int out = MPI_SUCCESS;
  for (int i = 0; i < n; i++) {
    const int rc = MPI_Wait(&reqs[i], sts ? &sts[i] : nullptr);
    if (rc && !out) out = rc;
  }
  return out;
}
int MPI_Test(MPI_Request* req, int* flag, MPI_Status* st) {
  const int rc = complete_request(*req, false, flag, st);
  if (*flag) *req = MPI_REQUEST_NULL;
  return rc;
}
int MPI_Probe(int source, int tag, MPI_Comm c, MPI_Status* st) {
  int e = check(c); if (e) return e;
  return recv_bytes(nullptr, 0, source, tag, c, st, true, true, nullptr);
}
int MPI_Iprobe(int source, int tag, MPI_Comm c, int* flag, MPI_Status* st) {
  int e = check(c); if (e) return e;
  return recv_bytes(nullptr, 0, source, tag, c, st, true, false, flag);
}
int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count) {
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  *count = st->count_ % (int)es ? MPI_UNDEFINED : st->count_ / (int)es;
  return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* out) {
  int e = check(c); if (e) return e;
  if (c == MPI_COMM_SELF || g_size == 1) { *out = color == MPI_UNDEFINED ? MPI_COMM_NULL : c; return MPI_SUCCESS; }
  struct CK { int color, key; } mine{color, key};
  std::vector<CK> all(g_size);
  e = allgather_bytes(&mine, all.data(), sizeof(CK));
  if (e) return e;
  if (color == MPI_UNDEFINED) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
  int same = 0;
  bool ordered = true;
  for (int r = 0; r < g_size; r++) {
    if (all[r].color == color) same++;
    if (r > 0 && all[r].key < all[r - 1].key) ordered = false;
  }
  if (same == g_size && ordered) { *out = MPI_COMM_WORLD; return MPI_SUCCESS; }   // everybody together, rank order kept
  if (same == 1) { *out = MPI_COMM_SELF; return MPI_SUCCESS; }                      // a group of one
  return fail("MPI_Comm_split: only the trivial partitions (all ranks together in rank order, or singletons) are provided");
}
int MPI_Comm_split_type(MPI_Comm c, int split_type, int key, MPI_Info, MPI_Comm* out) {
  // one box: every rank shares the node, so MPI_COMM_TYPE_SHARED groups everybody
  return MPI_Comm_split(c, split_type == MPI_UNDEFINED ? MPI_UNDEFINED : 0, key, out);
}

// ---- vector collectives: rooted ones move data point-to-point (reserved tags above MPI_TAG_UB), Allgatherv gathers the padded
// blocks through the mailbox allgather and unpacks
static const int kTagGatherv = MPI_TAG_UB + 1, kTagScatterv = MPI_TAG_UB + 2, kTagScan = MPI_TAG_UB + 3;

int MPI_Allgatherv(const void* sb, int sc, MPI_Datatype st, void* rb, const int* counts, const int* displs, MPI_Datatype rt, MPI_Comm c) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(rt);
  if (!es || type_size(st) != es) return MPI_ERR_TYPE;
  if (c == MPI_COMM_SELF || g_size == 1) { if (sb != MPI_IN_PLACE) memmove((char*)rb + (size_t)displs[0] * es, sb, (size_t)sc * es); return MPI_SUCCESS; }
  size_t width = 0;
  for (int r = 0; r < g_size; r++) width = std::max(width, (size_t)counts[r] * es);
  std::vector<unsigned char> mine(width ? width : 1, 0), all((width ? width : 1) * g_size);
  const void* src = sb == MPI_IN_PLACE ? (const char*)rb + (size_t)displs[g_rank] * es : sb;
  memcpy(mine.data(), src, (size_t)counts[g_rank] * es);
  e = allgather_bytes(mine.data(), all.data(), mine.size());
  if (e) return e;
  for (int r = 0; r < g_size; r++) memcpy((char*)rb + (size_t)displs[r] * es, all.data() + (size_t)r * mine.size(), (size_t)counts[r] * es);
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void* sb, int sc, MPI_Datatype st, void* rb, const int* counts, const int* displs, MPI_Datatype rt, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(st);
  if (!es) return MPI_ERR_TYPE;
  if (g_rank != root) return send_bytes(sb, (size_t)sc * es, root, kTagGatherv, c);
  const size_t rs = type_size(rt);
  for (int r = 0; r < g_size; r++) {
    char* dst = (char*)rb + (size_t)displs[r] * rs;
    if (r == root) { if (sb != MPI_IN_PLACE) memmove(dst, sb, (size_t)sc * es); continue; }
    e = recv_bytes(dst, (size_t)counts[r] * rs, r, kTagGatherv, c, nullptr, false, true, nullptr);
    if (e) return e;
  }
  return MPI_SUCCESS;
}

int MPI_Scatterv(const void* sb, const int* counts, const int* displs, MPI_Datatype st, void* rb, int rc_, MPI_Datatype rt, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  const size_t rs = type_size(rt);
  if (!rs) return MPI_ERR_TYPE;
  if (g_rank != root) return recv_bytes(rb, (size_t)rc_ * rs, root, kTagScatterv, c, nullptr, false, true, nullptr);
  const size_t es = type_size(st);
  for (int r = 0; r < g_size; r++) {
    const char* src = (const char*)sb + (size_t)displs[r] * es;
    if (r == root) { if (rb != MPI_IN_PLACE) memmove(rb, src, (size_t)counts[r] * es); continue; }
    e = send_bytes(src, (size_t)counts[r] * es, r, kTagScatterv, c);
    if (e) return e;
  }
  return MPI_SUCCESS;
}

int MPI_Reduce_scatter_block(const void* sb, void* rb, int rc_, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  const int n = c == MPI_COMM_SELF ? 1 : g_size;
  std::vector<unsigned char> full((size_t)rc_ * es * n);
  e = MPI_Allreduce(sb == MPI_IN_PLACE ? rb : sb, full.data(), rc_ * n, t, op, c);
  if (e) return e;
  memcpy(rb, full.data() + (size_t)(c == MPI_COMM_SELF ? 0 : g_rank) * rc_ * es, (size_t)rc_ * es);
  return MPI_SUCCESS;
}

// inclusive / exclusive prefix reduction along the rank order: a chain of point-to-point messages
static int scan_impl(const void* sb, void* rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c, bool exclusive) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  const size_t bytes = es * (size_t)count;
  std::vector<unsigned char> mine((const unsigned char*)(sb == MPI_IN_PLACE ? rb : sb), (const unsigned char*)(sb == MPI_IN_PLACE ? rb : sb) + bytes);
  if (c == MPI_COMM_SELF || g_size == 1) { if (!exclusive && sb != MPI_IN_PLACE) memmove(rb, sb, bytes); return MPI_SUCCESS; }
  std::vector<unsigned char> prefix(bytes);   // reduction over ranks 0 .. rank-1
  if (g_rank > 0) {
    e = recv_bytes(prefix.data(), bytes, g_rank - 1, kTagScan, c, nullptr, false, true, nullptr);
    if (e) return e;
  }
  std::vector<unsigned char> incl = g_rank > 0 ? prefix : mine;
  if (g_rank > 0 && !reduce_into(incl.data(), mine.data(), count, t, op)) return MPI_ERR_OP;   // prefix (op) mine, rank order preserved
  if (g_rank + 1 < g_size) {
    e = send_bytes(incl.data(), bytes, g_rank + 1, kTagScan, c);
    if (e) return e;
  }
  if (!exclusive) memcpy(rb, incl.data(), bytes);
  else if (g_rank > 0) memcpy(rb, prefix.data(), bytes);   // rank 0's result of MPI_Exscan is undefined
  return MPI_SUCCESS;
}
int MPI_Scan(const void* sb, void* rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return scan_impl(sb, rb, count, t, op, c, false); }
int MPI_Exscan(const void* sb, void* rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return scan_impl(sb, rb, count, t, op, c, true); }

}  // extern "C"
