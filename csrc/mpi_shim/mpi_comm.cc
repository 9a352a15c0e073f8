// libmpi (b200mpi shim): communicators, groups, derived datatypes, and the collectives of sub-communicators.
//
// Programs launched through an MPIJob's `mpirun` routinely split MPI_COMM_WORLD (row / column communicators, a "local" and a
// "cross" communicator the way Horovod's MPI bootstrap derives them, a subset that does I/O). The reference gets all of that from
// the Open MPI / MPICH / Intel MPI of its images (build/base/*.Dockerfile); here it is part of the shim:
//
//   * a communicator = ordered list of world ranks + a context id. Context ids are agreed with one MAX-allreduce over the parent
//     (every member proposes its next free id), so two communicators that share a process never share a context;
//   * communicators that contain every rank in world order ("world-like": MPI_COMM_WORLD, its dups, MPI_COMM_TYPE_SHARED splits)
//     keep the shared-memory collectives of mpi_shim.cc; singletons are local copies; everything else runs the collectives
//     below, built on the eager point-to-point layer (mpi_p2p.cc) with reserved tags above MPI_TAG_UB: binomial-tree broadcast,
//     gather-to-root reductions folded in rank order (same result bits as the world path), direct exchanges for allgather /
//     alltoall. Sends are eager, so "everybody sends, then everybody receives" cannot deadlock;
//   * groups (MPI_Comm_group / Group_incl / excl / translate_ranks / Comm_create), contiguous derived datatypes, error handlers
//     and the nonblocking collectives (which complete before the call returns: legal, no overlap).
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "mpi_internal.h"

namespace b200mpi_mpi {

namespace {
std::vector<Comm> g_comms;
int g_next_ctx = 2;
std::vector<std::vector<int>> g_groups;        // MPI_Group handle - 1 -> world ranks (handle 0 = MPI_GROUP_EMPTY)
std::vector<bool> g_group_live;
struct DType { bool live; size_t bytes; MPI_Datatype base; size_t count; };
std::vector<DType> g_types;                     // derived datatype handle - kFirstDerived
constexpr int kFirstDerived = 1000;

const int kTagBarrier = MPI_TAG_UB + 10, kTagBcast = MPI_TAG_UB + 11, kTagReduce = MPI_TAG_UB + 12, kTagGather = MPI_TAG_UB + 13,
          kTagAlltoall = MPI_TAG_UB + 14, kTagGatherv = MPI_TAG_UB + 1, kTagScatterv = MPI_TAG_UB + 2, kTagScan = MPI_TAG_UB + 3,
          kTagAlltoallv = MPI_TAG_UB + 15;

int csend(Comm* C, const void* buf, size_t bytes, int dest, int tag) { return send_bytes(buf, bytes, C->ranks[dest], tag, C->ctx); }
int crecv(Comm* C, void* buf, size_t bytes, int src, int tag) { return recv_bytes(buf, bytes, C->ranks[src], tag, C->ctx, nullptr, false, true, nullptr); }

int new_comm(const std::vector<int>& ranks, int ctx) {
  Comm c;
  c.live = true;
  c.ranks = ranks;
  c.ctx = ctx;
  c.my = -1;
  for (int i = 0; i < (int)ranks.size(); i++) if (ranks[i] == g_rank) c.my = i;
  c.world_like = (int)ranks.size() == g_size;
  for (int i = 0; i < (int)ranks.size() && c.world_like; i++) c.world_like = ranks[i] == i;
  for (size_t i = 2; i < g_comms.size(); i++)
    if (!g_comms[i].live) { g_comms[i] = c; return (int)i; }
  g_comms.push_back(c);
  return (int)g_comms.size() - 1;
}

// One context id for a new communicator (or a family of disjoint ones) created collectively over `P`
int agree_ctx(Comm* P, int* ctx) {
  int mine = g_next_ctx, top = mine;
  if (P->size() > 1) {
    const int e = MPI_Allreduce(&mine, &top, 1, MPI_INT, MPI_MAX, (MPI_Comm)(P - g_comms.data()));
    if (e) return e;
  }
  *ctx = top;
  g_next_ctx = top + 1;
  return MPI_SUCCESS;
}

int new_group(const std::vector<int>& ranks) {
  for (size_t i = 0; i < g_groups.size(); i++)
    if (!g_group_live[i]) { g_groups[i] = ranks; g_group_live[i] = true; return (int)i + 1; }
  g_groups.push_back(ranks);
  g_group_live.push_back(true);
  return (int)g_groups.size();
}
const std::vector<int>* group_of(MPI_Group g) {
  static const std::vector<int> empty;
  if (g == MPI_GROUP_EMPTY) return &empty;
  if (g < 1 || g > (int)g_groups.size() || !g_group_live[g - 1]) return nullptr;
  return &g_groups[g - 1];
}
}  // namespace

Comm* comm_of(MPI_Comm c) {
  if (c < 0 || c >= (int)g_comms.size() || !g_comms[c].live) return nullptr;
  return &g_comms[c];
}

void comms_reset(bool build) {
  g_comms.clear();
  g_groups.clear();
  g_group_live.clear();
  g_types.clear();
  g_next_ctx = 2;
  if (!build) return;
  std::vector<int> all((size_t)g_size);
  for (int i = 0; i < g_size; i++) all[(size_t)i] = i;
  g_comms.resize(2);
  g_comms[0].live = true; g_comms[0].ranks = all; g_comms[0].my = g_rank; g_comms[0].ctx = 0; g_comms[0].world_like = true; g_comms[0].name = "MPI_COMM_WORLD";
  g_comms[1].live = true; g_comms[1].ranks = {g_rank}; g_comms[1].my = 0; g_comms[1].ctx = 1; g_comms[1].world_like = g_size == 1; g_comms[1].name = "MPI_COMM_SELF";
}

size_t derived_type_size(MPI_Datatype t) {
  const int i = t - kFirstDerived;
  return (i >= 0 && i < (int)g_types.size() && g_types[(size_t)i].live) ? g_types[(size_t)i].bytes : 0;
}
// element type and element count behind `count` items of `t` (reductions work on the base type of contiguous types)
bool flatten_type(MPI_Datatype t, size_t count, MPI_Datatype* base, size_t* n) {
  const int i = t - kFirstDerived;
  if (i >= 0 && i < (int)g_types.size() && g_types[(size_t)i].live) { *base = g_types[(size_t)i].base; *n = count * g_types[(size_t)i].count; return true; }
  *base = t; *n = count;
  return type_size(t) != 0;
}

// ------------------------------------------------------------------------------------------ generic collectives --
int gen_barrier(Comm* C) {
  const int n = C->size(), m = C->my;
  char z = 0;
  if (m != 0) {
    int e = csend(C, &z, 0, 0, kTagBarrier); if (e) return e;
    return crecv(C, &z, 0, 0, kTagBarrier);
  }
  for (int r = 1; r < n; r++) { int e = crecv(C, &z, 0, r, kTagBarrier); if (e) return e; }
  for (int r = 1; r < n; r++) { int e = csend(C, &z, 0, r, kTagBarrier); if (e) return e; }
  return MPI_SUCCESS;
}

// binomial tree rooted at `root`
int gen_bcast(Comm* C, void* buf, size_t bytes, int root) {
  const int n = C->size(), vr = (C->my - root + n) % n;
  int mask = 1;
  while (mask < n) {
    if (vr & mask) { int e = crecv(C, buf, bytes, (vr - mask + root) % n, kTagBcast); if (e) return e; break; }
    mask <<= 1;
  }
  mask >>= 1;
  while (mask > 0) {
    if (vr + mask < n) { int e = csend(C, buf, bytes, (vr + mask + root) % n, kTagBcast); if (e) return e; }
    mask >>= 1;
  }
  return MPI_SUCCESS;
}

// gather to the root, fold in rank order (deterministic: the same bits the shared-memory path produces); `all`: then broadcast
int gen_reduce(Comm* C, const void* send, void* recv, size_t count, MPI_Datatype t, MPI_Op op, int root, bool all) {
  MPI_Datatype base; size_t n;
  if (!flatten_type(t, count, &base, &n)) return MPI_ERR_TYPE;
  const size_t bytes = n * type_size(base);
  const int W = C->size(), m = C->my;
  const void* mine = send == MPI_IN_PLACE ? recv : send;
  const int r0 = all ? 0 : root;
  int rc = MPI_SUCCESS;
  if (m != r0) {
    int e = csend(C, mine, bytes, r0, kTagReduce); if (e) return e;
  } else {
    std::vector<unsigned char> acc(bytes), tmp(bytes), own((const unsigned char*)mine, (const unsigned char*)mine + bytes);
    for (int r = 0; r < W; r++) {
      const unsigned char* x = own.data();
      if (r != m) { int e = crecv(C, tmp.data(), bytes, r, kTagReduce); if (e) return e; x = tmp.data(); }
      if (r == 0) memcpy(acc.data(), x, bytes);
      else if (!reduce_into(acc.data(), x, n, base, op)) rc = MPI_ERR_OP;
    }
    memcpy(recv, acc.data(), bytes);
  }
  if (all) { int e = gen_bcast(C, recv, bytes, 0); if (e) return e; }
  return rc;
}

int gen_allgather(Comm* C, const void* in, void* out, size_t bytes) {
  const int W = C->size(), m = C->my;
  std::vector<unsigned char> mine((const unsigned char*)in, (const unsigned char*)in + bytes);   // `in` may alias its own slot of `out`
  for (int k = 1; k < W; k++) { int e = csend(C, mine.data(), bytes, (m + k) % W, kTagGather); if (e) return e; }
  memcpy((char*)out + (size_t)m * bytes, mine.data(), bytes);
  for (int k = 1; k < W; k++) { const int r = (m - k + W) % W; int e = crecv(C, (char*)out + (size_t)r * bytes, bytes, r, kTagGather); if (e) return e; }
  return MPI_SUCCESS;
}

int gen_alltoall(Comm* C, const void* in, void* out, size_t bytes) {
  const int W = C->size(), m = C->my;
  for (int k = 1; k < W; k++) { const int d = (m + k) % W; int e = csend(C, (const char*)in + (size_t)d * bytes, bytes, d, kTagAlltoall); if (e) return e; }
  memmove((char*)out + (size_t)m * bytes, (const char*)in + (size_t)m * bytes, bytes);
  for (int k = 1; k < W; k++) { const int r = (m - k + W) % W; int e = crecv(C, (char*)out + (size_t)r * bytes, bytes, r, kTagAlltoall); if (e) return e; }
  return MPI_SUCCESS;
}

}  // namespace b200mpi_mpi
using namespace b200mpi_mpi;

extern "C" {

// ------------------------------------------------------------------------------------------------ communicators --
int MPI_Comm_rank(MPI_Comm c, int* r) { int e = check(c); if (e) return e; *r = comm_of(c)->my; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* s) { int e = check(c); if (e) return e; *s = comm_of(c)->size(); return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) {
  int e = check(c); if (e) return e;
  Comm* P = comm_of(c);
  int ctx = 0;
  e = agree_ctx(P, &ctx); if (e) return e;
  const std::vector<int> ranks = P->ranks;
  *n = new_comm(ranks, ctx);
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* c) {
  if (*c >= 2) { Comm* C = comm_of(*c); if (C) C->live = false; }
  *c = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* out) {
  int e = check(c); if (e) return e;
  Comm* P = comm_of(c);
  const int W = P->size();
  struct CK { int color, key; } mine{color, key};
  std::vector<CK> all((size_t)W);
  e = MPI_Allgather(&mine, (int)sizeof(CK), MPI_BYTE, all.data(), (int)sizeof(CK), MPI_BYTE, c);
  if (e) return e;
  int ctx = 0;
  e = agree_ctx(comm_of(c), &ctx); if (e) return e;   // every member takes part, MPI_UNDEFINED colours included
  P = comm_of(c);
  if (color == MPI_UNDEFINED) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
  std::vector<std::pair<std::pair<int, int>, int>> members;   // ((key, parent rank), world rank): ties keep the parent order
  for (int r = 0; r < W; r++)
    if (all[(size_t)r].color == color) members.push_back({{all[(size_t)r].key, r}, P->ranks[(size_t)r]});
  std::sort(members.begin(), members.end());
  std::vector<int> ranks;
  for (auto& mbr : members) ranks.push_back(mbr.second);
  *out = new_comm(ranks, ctx);
  return MPI_SUCCESS;
}
int MPI_Comm_split_type(MPI_Comm c, int split_type, int key, MPI_Info, MPI_Comm* out) {
  // one box: every rank shares the node, so MPI_COMM_TYPE_SHARED groups everybody
  return MPI_Comm_split(c, split_type == MPI_UNDEFINED ? MPI_UNDEFINED : 0, key, out);
}
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* result) {
  Comm* A = comm_of(a); Comm* B = comm_of(b);
  if (!A || !B) return MPI_ERR_COMM;
  if (a == b) { *result = MPI_IDENT; return MPI_SUCCESS; }
  if (A->ranks == B->ranks) { *result = MPI_CONGRUENT; return MPI_SUCCESS; }
  std::vector<int> x = A->ranks, y = B->ranks;
  std::sort(x.begin(), x.end()); std::sort(y.begin(), y.end());
  *result = x == y ? MPI_SIMILAR : MPI_UNEQUAL;
  return MPI_SUCCESS;
}
int MPI_Comm_set_name(MPI_Comm c, const char* name) { Comm* C = comm_of(c); if (!C) return MPI_ERR_COMM; C->name = name ? name : ""; return MPI_SUCCESS; }
int MPI_Comm_get_name(MPI_Comm c, char* name, int* len) {
  Comm* C = comm_of(c); if (!C) return MPI_ERR_COMM;
  *len = snprintf(name, MPI_MAX_OBJECT_NAME, "%s", C->name.c_str());
  return MPI_SUCCESS;
}
int MPI_Comm_test_inter(MPI_Comm c, int* flag) { if (!comm_of(c)) return MPI_ERR_COMM; *flag = 0; return MPI_SUCCESS; }
// Errors are always returned to the caller (MPI_ERRORS_RETURN behaviour); MPI_ERRORS_ARE_FATAL is accepted and remembered only
int MPI_Comm_set_errhandler(MPI_Comm c, MPI_Errhandler) { return comm_of(c) ? MPI_SUCCESS : MPI_ERR_COMM; }
int MPI_Comm_get_errhandler(MPI_Comm c, MPI_Errhandler* h) { if (!comm_of(c)) return MPI_ERR_COMM; *h = MPI_ERRORS_RETURN; return MPI_SUCCESS; }
int MPI_Errhandler_set(MPI_Comm c, MPI_Errhandler h) { return MPI_Comm_set_errhandler(c, h); }
int MPI_Errhandler_free(MPI_Errhandler* h) { *h = MPI_ERRHANDLER_NULL; return MPI_SUCCESS; }
int MPI_Error_class(int code, int* cls) { *cls = code; return MPI_SUCCESS; }

// ------------------------------------------------------------------------------------------------------- groups --
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { Comm* C = comm_of(c); if (!C) return MPI_ERR_COMM; *g = new_group(C->ranks); return MPI_SUCCESS; }
int MPI_Group_size(MPI_Group g, int* n) { auto* v = group_of(g); if (!v) return MPI_ERR_GROUP; *n = (int)v->size(); return MPI_SUCCESS; }
int MPI_Group_rank(MPI_Group g, int* r) {
  auto* v = group_of(g); if (!v) return MPI_ERR_GROUP;
  *r = MPI_UNDEFINED;
  for (int i = 0; i < (int)v->size(); i++) if ((*v)[(size_t)i] == g_rank) *r = i;
  return MPI_SUCCESS;
}
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out) {
  auto* v = group_of(g); if (!v) return MPI_ERR_GROUP;
  std::vector<int> sel;
  for (int i = 0; i < n; i++) { if (ranks[i] < 0 || ranks[i] >= (int)v->size()) return MPI_ERR_RANK; sel.push_back((*v)[(size_t)ranks[i]]); }
  *out = n == 0 ? MPI_GROUP_EMPTY : new_group(sel);
  return MPI_SUCCESS;
}
int MPI_Group_excl(MPI_Group g, int n, const int* ranks, MPI_Group* out) {
  auto* v = group_of(g); if (!v) return MPI_ERR_GROUP;
  std::vector<int> sel;
  for (int i = 0; i < (int)v->size(); i++) if (std::find(ranks, ranks + n, i) == ranks + n) sel.push_back((*v)[(size_t)i]);
  *out = sel.empty() ? MPI_GROUP_EMPTY : new_group(sel);
  return MPI_SUCCESS;
}
int MPI_Group_translate_ranks(MPI_Group a, int n, const int* ra, MPI_Group b, int* rb) {
  auto* A = group_of(a); auto* B = group_of(b);
  if (!A || !B) return MPI_ERR_GROUP;
  for (int i = 0; i < n; i++) {
    rb[i] = MPI_UNDEFINED;
    if (ra[i] == MPI_PROC_NULL) { rb[i] = MPI_PROC_NULL; continue; }
    if (ra[i] < 0 || ra[i] >= (int)A->size()) return MPI_ERR_RANK;
    for (int k = 0; k < (int)B->size(); k++) if ((*B)[(size_t)k] == (*A)[(size_t)ra[i]]) rb[i] = k;
  }
  return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group* g) { if (*g >= 1 && *g <= (int)g_groups.size()) g_group_live[(size_t)*g - 1] = false; *g = MPI_GROUP_NULL; return MPI_SUCCESS; }
// Collective over `c`; ranks outside `g` get MPI_COMM_NULL
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* out) {
  int e = check(c); if (e) return e;
  auto* v = group_of(g); if (!v) return MPI_ERR_GROUP;
  const std::vector<int> ranks = *v;
  int ctx = 0;
  e = agree_ctx(comm_of(c), &ctx); if (e) return e;
  if (std::find(ranks.begin(), ranks.end(), g_rank) == ranks.end()) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
  *out = new_comm(ranks, ctx);
  return MPI_SUCCESS;
}

// MPI-3: collective over the MEMBERS of `g` only (MPI_Comm_create involves every rank of `c`). The context id is agreed through
// the group's first member: everybody sends its next free id (point-to-point on the parent's context, reserved tag + `tag`), the
// leader answers with the maximum. What torch.distributed's new_group needs: only members call into the backend.
int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* out) {
  int e = check(c); if (e) return e;
  auto* v = group_of(g); if (!v) return MPI_ERR_GROUP;
  const std::vector<int> ranks = *v;
  Comm* P = comm_of(c);
  int me = -1;
  for (int i = 0; i < (int)ranks.size(); i++) if (ranks[(size_t)i] == g_rank) me = i;
  if (me < 0) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
  const int t = MPI_TAG_UB + 100 + (tag & 0xffff);
  int mine = g_next_ctx, top = mine;
  if (ranks.size() > 1) {
    if (me != 0) {
      e = send_bytes(&mine, sizeof(int), ranks[0], t, P->ctx); if (e) return e;
      e = recv_bytes(&top, sizeof(int), ranks[0], t, P->ctx, nullptr, false, true, nullptr); if (e) return e;
    } else {
      for (size_t k = 1; k < ranks.size(); k++) {
        int theirs = 0;
        e = recv_bytes(&theirs, sizeof(int), ranks[k], t, P->ctx, nullptr, false, true, nullptr); if (e) return e;
        top = std::max(top, theirs);
      }
      for (size_t k = 1; k < ranks.size(); k++) { e = send_bytes(&top, sizeof(int), ranks[k], t, P->ctx); if (e) return e; }
    }
  }
  g_next_ctx = top + 1;
  *out = new_comm(ranks, top);
  return MPI_SUCCESS;
}

// -------------------------------------------------------------------------------------------- derived datatypes --
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype* out) {
  MPI_Datatype base; size_t n;
  if (count < 0 || !flatten_type(old, (size_t)count, &base, &n)) return MPI_ERR_TYPE;
  g_types.push_back(DType{true, n * type_size(base), base, n});
  *out = kFirstDerived + (int)g_types.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* t) {
  const int i = *t - kFirstDerived;
  if (i >= 0 && i < (int)g_types.size()) g_types[(size_t)i].live = false;
  *t = MPI_DATATYPE_NULL;
  return MPI_SUCCESS;
}
int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint* lb, MPI_Aint* extent) {
  const size_t s = type_size(t);
  if (!s) return MPI_ERR_TYPE;
  *lb = 0; *extent = (MPI_Aint)s;
  return MPI_SUCCESS;
}

// -------------------------------------------------------------------------- vector collectives, scans (any comm) --
int MPI_Allgatherv(const void* sb, int sc, MPI_Datatype st, void* rb, const int* counts, const int* displs, MPI_Datatype rt, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t es = type_size(rt);
  if (!es || (sb != MPI_IN_PLACE && type_size(st) != es)) return MPI_ERR_TYPE;
  const int W = C->size(), m = C->my;
  if (W == 1) { if (sb != MPI_IN_PLACE) memmove((char*)rb + (size_t)displs[0] * es, sb, (size_t)sc * es); return MPI_SUCCESS; }
  size_t width = 0;
  for (int r = 0; r < W; r++) width = std::max(width, (size_t)counts[r] * es);
  std::vector<unsigned char> mine(width ? width : 1, 0), all((width ? width : 1) * (size_t)W);
  const void* src = sb == MPI_IN_PLACE ? (const char*)rb + (size_t)displs[m] * es : sb;
  memcpy(mine.data(), src, (size_t)counts[m] * es);
  e = C->world_like ? allgather_bytes(mine.data(), all.data(), mine.size()) : gen_allgather(C, mine.data(), all.data(), mine.size());
  if (e) return e;
  for (int r = 0; r < W; r++) memcpy((char*)rb + (size_t)displs[r] * es, all.data() + (size_t)r * mine.size(), (size_t)counts[r] * es);
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void* sb, int sc, MPI_Datatype st, void* rb, const int* counts, const int* displs, MPI_Datatype rt, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  if (root < 0 || root >= C->size()) return MPI_ERR_ROOT;
  if (C->my != root) {
    const size_t es = type_size(st);
    if (!es) return MPI_ERR_TYPE;
    return csend(C, sb, (size_t)sc * es, root, kTagGatherv);
  }
  const size_t rs = type_size(rt);
  if (!rs) return MPI_ERR_TYPE;
  for (int r = 0; r < C->size(); r++) {
    char* dst = (char*)rb + (size_t)displs[r] * rs;
    if (r == root) { if (sb != MPI_IN_PLACE) memmove(dst, sb, (size_t)sc * type_size(st)); continue; }
    e = crecv(C, dst, (size_t)counts[r] * rs, r, kTagGatherv);
    if (e) return e;
  }
  return MPI_SUCCESS;
}

int MPI_Scatterv(const void* sb, const int* counts, const int* displs, MPI_Datatype st, void* rb, int rc_, MPI_Datatype rt, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  if (root < 0 || root >= C->size()) return MPI_ERR_ROOT;
  if (C->my != root) {
    const size_t rs = type_size(rt);
    if (!rs) return MPI_ERR_TYPE;
    return crecv(C, rb, (size_t)rc_ * rs, root, kTagScatterv);
  }
  const size_t es = type_size(st);
  if (!es) return MPI_ERR_TYPE;
  for (int r = 0; r < C->size(); r++) {
    const char* src = (const char*)sb + (size_t)displs[r] * es;
    if (r == root) { if (rb != MPI_IN_PLACE) memmove(rb, src, (size_t)counts[r] * es); continue; }
    e = csend(C, src, (size_t)counts[r] * es, r, kTagScatterv);
    if (e) return e;
  }
  return MPI_SUCCESS;
}

int MPI_Alltoallv(const void* sb, const int* scounts, const int* sdispls, MPI_Datatype st, void* rb, const int* rcounts, const int* rdispls,
                  MPI_Datatype rt, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t ss = type_size(st), rs = type_size(rt);
  if (!ss || !rs) return MPI_ERR_TYPE;
  const int W = C->size(), m = C->my;
  for (int k = 1; k < W; k++) {
    const int d = (m + k) % W;
    e = csend(C, (const char*)sb + (size_t)sdispls[d] * ss, (size_t)scounts[d] * ss, d, kTagAlltoallv);
    if (e) return e;
  }
  memmove((char*)rb + (size_t)rdispls[m] * rs, (const char*)sb + (size_t)sdispls[m] * ss, (size_t)scounts[m] * ss);
  for (int k = 1; k < W; k++) {
    const int r = (m - k + W) % W;
    e = crecv(C, (char*)rb + (size_t)rdispls[r] * rs, (size_t)rcounts[r] * rs, r, kTagAlltoallv);
    if (e) return e;
  }
  return MPI_SUCCESS;
}

int MPI_Reduce_scatter_block(const void* sb, void* rb, int rc_, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  const int n = C->size();
  std::vector<unsigned char> full((size_t)rc_ * es * (size_t)n);
  if (sb == MPI_IN_PLACE) memcpy(full.data(), rb, full.size());
  e = MPI_Allreduce(sb == MPI_IN_PLACE ? MPI_IN_PLACE : sb, full.data(), rc_ * n, t, op, c);
  if (e) return e;
  memcpy(rb, full.data() + (size_t)C->my * (size_t)rc_ * es, (size_t)rc_ * es);
  return MPI_SUCCESS;
}
int MPI_Reduce_scatter(const void* sb, void* rb, const int* counts, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  size_t total = 0, off = 0;
  for (int r = 0; r < C->size(); r++) { if (r == C->my) off = total; total += (size_t)counts[r]; }
  std::vector<unsigned char> full(total * es);
  if (sb == MPI_IN_PLACE) memcpy(full.data(), rb, full.size());
  e = MPI_Allreduce(sb == MPI_IN_PLACE ? MPI_IN_PLACE : sb, full.data(), (int)total, t, op, c);
  if (e) return e;
  memcpy(rb, full.data() + off * es, (size_t)counts[C->my] * es);
  return MPI_SUCCESS;
}

// inclusive / exclusive prefix reduction along the rank order: a chain of point-to-point messages
static int scan_impl(const void* sb, void* rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c, bool exclusive) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  MPI_Datatype base; size_t n;
  if (!flatten_type(t, (size_t)count, &base, &n)) return MPI_ERR_TYPE;
  if (!op_supported(base, op)) return MPI_ERR_OP;
  const size_t bytes = n * type_size(base);
  const int W = C->size(), m = C->my;
  const unsigned char* in = (const unsigned char*)(sb == MPI_IN_PLACE ? rb : sb);
  std::vector<unsigned char> mine(in, in + bytes);
  if (W == 1) { if (!exclusive && sb != MPI_IN_PLACE) memmove(rb, sb, bytes); return MPI_SUCCESS; }
  std::vector<unsigned char> prefix(bytes);   // reduction over ranks 0 .. m-1
  if (m > 0) { e = crecv(C, prefix.data(), bytes, m - 1, kTagScan); if (e) return e; }
  std::vector<unsigned char> incl = m > 0 ? prefix : mine;
  if (m > 0 && !reduce_into(incl.data(), mine.data(), n, base, op)) return MPI_ERR_OP;   // prefix (op) mine, rank order preserved
  if (m + 1 < W) { e = csend(C, incl.data(), bytes, m + 1, kTagScan); if (e) return e; }
  if (!exclusive) memcpy(rb, incl.data(), bytes);
  else if (m > 0) memcpy(rb, prefix.data(), bytes);   // rank 0's result of MPI_Exscan is undefined
  return MPI_SUCCESS;
}
int MPI_Scan(const void* sb, void* rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return scan_impl(sb, rb, count, t, op, c, false); }
int MPI_Exscan(const void* sb, void* rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return scan_impl(sb, rb, count, t, op, c, true); }

// ------------------------------------------------------------------------------------ nonblocking collectives --
// The operation runs to completion inside the call (allowed: a nonblocking call MAY complete early) and the request is born
// complete, so MPI_Wait / MPI_Test / MPI_Waitall accept it like any other. No overlap, but programs written against
// MPI-3 link and run.
static int done_request(int rc, MPI_Request* req) {
  char dummy = 0;
  // a completed zero-byte self-send carries the result code through the ordinary request table
  const int e = MPI_Isend(&dummy, 0, MPI_BYTE, MPI_PROC_NULL, 0, MPI_COMM_SELF, req);
  return rc ? rc : e;
}
int MPI_Ibarrier(MPI_Comm c, MPI_Request* req) { return done_request(MPI_Barrier(c), req); }
int MPI_Ibcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Bcast(buf, count, t, root, c), req); }
int MPI_Iallreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Allreduce(s, r, n, t, op, c), req); }
int MPI_Ireduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Reduce(s, r, n, t, op, root, c), req); }
int MPI_Iallgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Allgather(s, sn, st, r, rn, rt, c), req); }
int MPI_Ialltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Alltoall(s, sn, st, r, rn, rt, c), req); }
int MPI_Igather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Gather(s, sn, st, r, rn, rt, root, c), req); }
int MPI_Iscatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c, MPI_Request* req) { return done_request(MPI_Scatter(s, sn, st, r, rn, rt, root, c), req); }

}  // extern "C"
