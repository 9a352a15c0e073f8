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
  int src = 0, tag = 0, comm = 0;   // src: WORLD rank (or MPI_ANY_SOURCE), comm: context id
  MPI_Comm handle = MPI_COMM_WORLD;
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
    if (h.bytes && h.offset + h.bytes <= m->total) {   // zero-byte messages (barrier tokens) carry no payload
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

}  // namespace

namespace b200mpi_mpi {
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

}  // namespace b200mpi_mpi

namespace {
// First message in arrival order that matches (source, tag, context); nullptr if none has started arriving.
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

}  // namespace

namespace b200mpi_mpi {
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

}  // namespace b200mpi_mpi

namespace {
// status->MPI_SOURCE arrives as a world rank: report the rank in the communicator the call was made on
void to_comm_rank(MPI_Comm c, MPI_Status* st) {
  if (!st || st->MPI_SOURCE < 0) return;
  Comm* C = comm_of(c);
  if (!C || C->world_like) return;
  for (int i = 0; i < C->size(); i++)
    if (C->ranks[i] == st->MPI_SOURCE) { st->MPI_SOURCE = i; return; }
}
// comm rank -> world rank (wildcards and MPI_PROC_NULL pass through); -3: out of range
int to_world(Comm* C, int r) {
  if (r == MPI_ANY_SOURCE || r == MPI_PROC_NULL) return r;
  if (r < 0 || r >= C->size()) return -3;
  return C->ranks[r];
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
  if (r.is_recv) to_comm_rank(r.handle, &r.st);
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
  Comm* C = comm_of(c);
  const int w = to_world(C, dest);
  if (w == -3) return MPI_ERR_RANK;
  return send_bytes(buf, es * (size_t)count, w, tag, C->ctx);
}
int MPI_Ssend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) { return MPI_Send(buf, count, t, dest, tag, c); }
int MPI_Bsend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) { return MPI_Send(buf, count, t, dest, tag, c); }   // every send is buffered at the receiver
int MPI_Rsend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) { return MPI_Send(buf, count, t, dest, tag, c); }

int MPI_Recv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm c, MPI_Status* st) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es || count < 0) return MPI_ERR_TYPE;
  Comm* C = comm_of(c);
  const int w = to_world(C, source);
  if (w == -3) return MPI_ERR_RANK;
  MPI_Status local;
  e = recv_bytes(buf, es * (size_t)count, w, tag, C->ctx, st ? st : &local, false, true, nullptr);
  to_comm_rank(c, st);
  return e;
}

int MPI_Sendrecv(const void* sb, int sc, MPI_Datatype stype, int dest, int stag, void* rb, int rc_, MPI_Datatype rtype, int source, int rtag,
                 MPI_Comm c, MPI_Status* st) {
  int e = MPI_Send(sb, sc, stype, dest, stag, c);   // eager: never waits for the matching receive
  if (e) return e;
  return MPI_Recv(rb, rc_, rtype, source, rtag, c, st);
}
int MPI_Sendrecv_replace(void* buf, int count, MPI_Datatype t, int dest, int stag, int source, int rtag, MPI_Comm c, MPI_Status* st) {
  int e = MPI_Send(buf, count, t, dest, stag, c);   // the outgoing copy has left the buffer when the eager send returns
  if (e) return e;
  return MPI_Recv(buf, count, t, source, rtag, c, st);
}

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req) {
  const int rc = MPI_Send(buf, count, t, dest, tag, c);  // eager send completes here; the request only carries the result
  const int id = new_request();
  g_reqs[id].rc = rc;
  g_reqs[id].done = true;
  *req = id;
  return rc;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm c, MPI_Request* req) {
  int e = check(c); if (e) return e;
  const size_t es = type_size(t);
  if (!es || count < 0) return MPI_ERR_TYPE;
  Comm* C = comm_of(c);
  const int w = to_world(C, source);
  if (w == -3) return MPI_ERR_RANK;
  const int id = new_request();
  Request& r = g_reqs[id];
  r.is_recv = true; r.buf = buf; r.cap = es * (size_t)count; r.src = w; r.tag = tag; r.comm = C->ctx; r.handle = c;
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
int MPI_Testall(int n, MPI_Request* reqs, int* flag, MPI_Status* sts) {
  // all-or-nothing: peek first (a request that is not ready stays untouched), complete only when every one is ready
  *flag = 1;
  for (int i = 0; i < n && *flag; i++) {
    const int id = reqs[i];
    if (id == MPI_REQUEST_NULL) continue;
    if (id < 0 || id >= (int)g_reqs.size() || !g_reqs[id].active) return MPI_ERR_REQUEST;
    Request& r = g_reqs[id];
    if (r.done) continue;
    int got = 0;
    r.rc = recv_bytes(r.buf, r.cap, r.src, r.tag, r.comm, &r.st, false, false, &got);
    if (r.rc != MPI_SUCCESS || got) r.done = true; else *flag = 0;
  }
  if (!*flag) return MPI_SUCCESS;
  return MPI_Waitall(n, reqs, sts);
}
int MPI_Testany(int n, MPI_Request* reqs, int* index, int* flag, MPI_Status* st) {
  *index = MPI_UNDEFINED;
  *flag = 1;
  bool any_active = false;
  for (int i = 0; i < n; i++) {
    if (reqs[i] == MPI_REQUEST_NULL) continue;
    any_active = true;
    int f = 0;
    const int rc = MPI_Test(&reqs[i], &f, st);
    if (f) { *index = i; return rc; }
  }
  if (any_active) *flag = 0;
  return MPI_SUCCESS;
}
int MPI_Waitany(int n, MPI_Request* reqs, int* index, MPI_Status* st) {
  for (;;) {
    int flag = 0;
    const int rc = MPI_Testany(n, reqs, index, &flag, st);
    if (rc || flag) return rc;
    if (g_rv && g_rv->aborted()) return fail("job aborted while waiting");
    wait_readable(20);
  }
}
int MPI_Waitsome(int n, MPI_Request* reqs, int* outcount, int* indices, MPI_Status* sts) {
  *outcount = 0;
  bool any_active = false;
  for (int i = 0; i < n; i++) any_active = any_active || reqs[i] != MPI_REQUEST_NULL;
  if (!any_active) { *outcount = MPI_UNDEFINED; return MPI_SUCCESS; }
  int out = MPI_SUCCESS;
  while (*outcount == 0) {
    for (int i = 0; i < n; i++) {
      if (reqs[i] == MPI_REQUEST_NULL) continue;
      int f = 0;
      const int rc = MPI_Test(&reqs[i], &f, sts ? &sts[*outcount] : nullptr);
      if (f) { indices[(*outcount)++] = i; if (rc && !out) out = rc; }
    }
    if (*outcount == 0) {
      if (g_rv && g_rv->aborted()) return fail("job aborted while waiting");
      wait_readable(20);
    }
  }
  return out;
}
int MPI_Request_free(MPI_Request* req) {
  if (*req != MPI_REQUEST_NULL && *req >= 0 && *req < (int)g_reqs.size()) g_reqs[*req].active = false;
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Cancel(MPI_Request*) { return MPI_SUCCESS; }   // eager sends have completed; a posted receive simply stays unmatched
int MPI_Probe(int source, int tag, MPI_Comm c, MPI_Status* st) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const int w = to_world(C, source);
  if (w == -3) return MPI_ERR_RANK;
  e = recv_bytes(nullptr, 0, w, tag, C->ctx, st, true, true, nullptr);
  to_comm_rank(c, st);
  return e;
}
int MPI_Iprobe(int source, int tag, MPI_Comm c, int* flag, MPI_Status* st) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const int w = to_world(C, source);
  if (w == -3) return MPI_ERR_RANK;
  e = recv_bytes(nullptr, 0, w, tag, C->ctx, st, true, false, flag);
  if (*flag) to_comm_rank(c, st);
  return e;
}
int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count) {
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  *count = st->count_ % (int)es ? MPI_UNDEFINED : st->count_ / (int)es;
  return MPI_SUCCESS;
}

}  // extern "C"
