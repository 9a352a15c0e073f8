// libmpi (b200mpi shim): the MPI subset of mpi.h over the shm rendezvous.
// Every collective is built from one primitive — a chunked allgather through
// the per-rank mailboxes of the segment — which is plenty for CPU control
// traffic (pi's 8-byte MPI_Reduce, Horovod-style bootstrap exchanges).
#include "mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../runtime/rendezvous.h"

#include "mpi_internal.h"

using b200mpi::Rendezvous;

namespace b200mpi_mpi {
Rendezvous* g_rv = nullptr;
int g_rank = 0, g_size = 1;
bool g_init = false, g_final = false;
int g_timeout_ms = 60000;
// bulk-data boxes (Rendezvous::open_boxes): large collectives move `g_box` bytes per rank per step instead of 64 KiB, and a
// reduction is folded slice-parallel (every rank reduces 1/world of each chunk); nullptr -> the mailboxes carry everything
unsigned char* g_boxes = nullptr;
size_t g_boxes_bytes = 0, g_box = 0;
inline unsigned char* box_data(int r) { return g_boxes ? g_boxes + (size_t)r * 2 * g_box : g_rv->header()->slot[r].mailbox; }
inline unsigned char* box_result(int r) { return g_boxes + (size_t)r * 2 * g_box + g_box; }
inline size_t box_bytes() { return g_boxes ? g_box : b200mpi::kRvMailbox; }

int env_first(std::initializer_list<const char*> names, int dflt) {
  for (const char* n : names) {
    const char* v = getenv(n);
    if (v && *v) return atoi(v);
  }
  return dflt;
}

int fail(const std::string& what) {
  fprintf(stderr, "[libmpi b200mpi rank %d] %s\n", g_rank, what.c_str());
  return MPI_ERR_OTHER;
}

size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_SIGNED_CHAR: case MPI_UNSIGNED_CHAR: case MPI_BYTE: case MPI_C_BOOL: return 1;
    case MPI_SHORT: case MPI_UNSIGNED_SHORT: return 2;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: case MPI_INT32_T: case MPI_UINT32_T: return 4;
    case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_LONG_LONG: case MPI_UNSIGNED_LONG_LONG: case MPI_DOUBLE:
    case MPI_INT64_T: case MPI_UINT64_T: return 8;
    case MPI_FLOAT_INT: case MPI_2INT: return 8;
    case MPI_DOUBLE_INT: case MPI_LONG_INT: return 16;    // the C struct's size (stride of an array of pairs)
    default: return derived_type_size(t);
  }
}

template <typename T>
void combine(T* acc, const T* x, size_t n, MPI_Op op) {
  for (size_t i = 0; i < n; i++) {
    switch (op) {
      case MPI_SUM: acc[i] = acc[i] + x[i]; break;
      case MPI_PROD: acc[i] = acc[i] * x[i]; break;
      case MPI_MAX: acc[i] = std::max(acc[i], x[i]); break;
      case MPI_MIN: acc[i] = std::min(acc[i], x[i]); break;
      case MPI_LAND: acc[i] = (T)(acc[i] && x[i]); break;
      case MPI_LOR: acc[i] = (T)(acc[i] || x[i]); break;
      default: break;
    }
  }
}
template <typename T>
void combine_bits(T* acc, const T* x, size_t n, MPI_Op op) {
  for (size_t i = 0; i < n; i++) acc[i] = op == MPI_BAND ? (T)(acc[i] & x[i]) : (T)(acc[i] | x[i]);
}

// MPI_MAXLOC / MPI_MINLOC on {value, index} pairs: the extreme value wins, equal values keep the lower index
template <typename V>
void combine_loc(void* acc_, const void* x_, size_t n, bool want_max) {
  struct P { V v; int i; };
  P* a = static_cast<P*>(acc_);
  const P* x = static_cast<const P*>(x_);
  for (size_t k = 0; k < n; k++) {
    const bool better = want_max ? x[k].v > a[k].v : x[k].v < a[k].v;
    if (better || (x[k].v == a[k].v && x[k].i < a[k].i)) a[k] = x[k];
  }
}

std::vector<MPI_User_function*> g_user_ops;   // MPI_Op handle - kFirstUserOp
constexpr int kFirstUserOp = 100;
constexpr int kFirstDerivedType = 1000;       // mpi_comm.cc: handles of MPI_Type_contiguous types (reduced through their base type)

// Is (type, op) something reduce_into can do? Checked by every rank BEFORE a reduction starts to move data, so an invalid
// combination fails on all ranks alike (not only on those whose slice happens to be non-empty).
bool op_supported(MPI_Datatype t, MPI_Op op) {
  if (op >= kFirstUserOp) return (size_t)(op - kFirstUserOp) < g_user_ops.size() && g_user_ops[(size_t)(op - kFirstUserOp)] != nullptr;
  const bool pair = t == MPI_FLOAT_INT || t == MPI_DOUBLE_INT || t == MPI_LONG_INT || t == MPI_2INT;
  if (op == MPI_MAXLOC || op == MPI_MINLOC) return pair;
  if (pair || op < MPI_SUM || op > MPI_BOR || !type_size(t) || t >= kFirstDerivedType) return false;
  if ((op == MPI_BAND || op == MPI_BOR) && (t == MPI_FLOAT || t == MPI_DOUBLE)) return false;
  return true;
}

bool reduce_into(void* acc, const void* x, size_t n, MPI_Datatype t, MPI_Op op) {
  if (op >= kFirstUserOp) {
    const size_t k = (size_t)(op - kFirstUserOp);
    if (k >= g_user_ops.size() || !g_user_ops[k]) return false;
    // acc holds the lower ranks, x the next rank: result = acc (op) x = fn(in = acc, inout = copy of x)
    const size_t bytes = n * type_size(t);
    std::vector<unsigned char> tmp((const unsigned char*)x, (const unsigned char*)x + bytes);
    for (size_t done = 0; done < n;) {   // `len` is an int
      const size_t part = std::min(n - done, (size_t)1 << 30);
      int len = (int)part;
      MPI_Datatype dt = t;
      g_user_ops[k]((unsigned char*)acc + done * type_size(t), tmp.data() + done * type_size(t), &len, &dt);
      done += part;
    }
    memcpy(acc, tmp.data(), bytes);
    return true;
  }
  if (op == MPI_MAXLOC || op == MPI_MINLOC) {
    const bool mx = op == MPI_MAXLOC;
    switch (t) {
      case MPI_FLOAT_INT: combine_loc<float>(acc, x, n, mx); return true;
      case MPI_DOUBLE_INT: combine_loc<double>(acc, x, n, mx); return true;
      case MPI_LONG_INT: combine_loc<long>(acc, x, n, mx); return true;
      case MPI_2INT: combine_loc<int>(acc, x, n, mx); return true;
      default: return false;
    }
  }
#define CASE(MT, CT) case MT: if (op == MPI_BAND || op == MPI_BOR) combine_bits((CT*)acc, (const CT*)x, n, op); else combine((CT*)acc, (const CT*)x, n, op); return true;
#define CASEF(MT, CT) case MT: if (op == MPI_BAND || op == MPI_BOR) return false; combine((CT*)acc, (const CT*)x, n, op); return true;
  switch (t) {
    CASE(MPI_CHAR, char) CASE(MPI_SIGNED_CHAR, signed char) CASE(MPI_UNSIGNED_CHAR, unsigned char) CASE(MPI_BYTE, unsigned char)
    CASE(MPI_C_BOOL, unsigned char) CASE(MPI_SHORT, short) CASE(MPI_UNSIGNED_SHORT, unsigned short) CASE(MPI_INT, int)
    CASE(MPI_INT32_T, int) CASE(MPI_UNSIGNED, unsigned) CASE(MPI_UINT32_T, unsigned) CASE(MPI_LONG, long)
    CASE(MPI_UNSIGNED_LONG, unsigned long) CASE(MPI_LONG_LONG, long long) CASE(MPI_INT64_T, long long)
    CASE(MPI_UNSIGNED_LONG_LONG, unsigned long long) CASE(MPI_UINT64_T, unsigned long long)
    CASEF(MPI_FLOAT, float) CASEF(MPI_DOUBLE, double)
    default: return false;
  }
#undef CASE
#undef CASEF
}

// out[r*bytes .. ] = rank r's `in` (bytes each), any size, one box (or mailbox) per rank per step
int allgather_bytes(const void* in, void* out, size_t bytes) {
  if (g_size == 1) { if (out != in) memmove(out, in, bytes); return MPI_SUCCESS; }
  std::string err;
  const size_t chunk = box_bytes();
  for (size_t done = 0; done < bytes; done += chunk) {
    const size_t n = std::min(chunk, bytes - done);
    memcpy(box_data(g_rank), (const char*)in + done, n);
    if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
    for (int r = 0; r < g_size; r++) memcpy((char*)out + (size_t)r * bytes + done, box_data(r), n);
    if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
  }
  return MPI_SUCCESS;
}

int bcast_bytes(void* buf, size_t bytes, int root) {
  std::string err;
  const size_t chunk = box_bytes();
  for (size_t done = 0; done < bytes; done += chunk) {
    const size_t n = std::min(chunk, bytes - done);
    if (g_rank == root) memcpy(box_data(root), (char*)buf + done, n);
    if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
    if (g_rank != root) memcpy((char*)buf + done, box_data(root), n);
    if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
  }
  return MPI_SUCCESS;
}

int check(MPI_Comm c) {
  if (!g_init || g_final) return fail("MPI call outside MPI_Init/MPI_Finalize");
  if (!comm_of(c)) return MPI_ERR_COMM;
  return MPI_SUCCESS;
}
}  // namespace b200mpi_mpi
using namespace b200mpi_mpi;

extern "C" {

int MPI_Init(int*, char***) {
  if (g_init) return MPI_SUCCESS;
  g_rank = env_first({"B200MPI_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "RANK"}, 0);
  g_size = env_first({"B200MPI_WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "WORLD_SIZE"}, 1);
  g_timeout_ms = env_first({"B200MPI_TIMEOUT_MS"}, 60000);
  if (g_size > 1) {
    const char* job = getenv("B200MPI_JOB_ID");
    std::string id = std::string("mpi-") + (job && *job ? job : "default");
    g_rv = new Rendezvous;
    std::string err;
    if (g_rv->attach(id, g_rank, g_size, -1, g_timeout_ms, &err)) { fail("MPI_Init: " + err); return MPI_ERR_OTHER; }
    if (p2p_init()) return MPI_ERR_OTHER;
    if (g_rv->barrier(g_timeout_ms, &err)) { fail("MPI_Init: " + err); return MPI_ERR_OTHER; }   // every message socket is bound
    const int kb = env_first({"B200MPI_MPI_MAILBOX_KB"}, 256);
    if (kb > 0) {
      g_box = (size_t)std::max(64, kb) * 1024;
      g_boxes = g_rv->open_boxes(g_box, g_timeout_ms, &g_boxes_bytes);
    }
  }
  comms_reset(true);
  g_init = true;
  return MPI_SUCCESS;
}
int MPI_Init_thread(int* a, char*** b, int required, int* provided) {
  if (provided) *provided = required < MPI_THREAD_SERIALIZED ? required : MPI_THREAD_SERIALIZED;
  return MPI_Init(a, b);
}
int MPI_Initialized(int* f) { *f = g_init; return MPI_SUCCESS; }
int MPI_Finalized(int* f) { *f = g_final; return MPI_SUCCESS; }
int MPI_Finalize(void) {
  if (!g_init || g_final) return MPI_SUCCESS;
  if (g_rv) {
    std::string err;
    g_rv->barrier(g_timeout_ms, &err);
    p2p_shutdown();
    Rendezvous::close_boxes(g_boxes, g_boxes_bytes);
    g_boxes = nullptr;
    g_rv->detach(g_rank == 0);
    delete g_rv;
    g_rv = nullptr;
  }
  comms_reset(false);
  g_final = true;
  return MPI_SUCCESS;
}
// Elastic rescale in place (hvd.elastic with B200MPI_ELASTIC_DIR): leave the current world (every rank of it calls this or
// MPI_Finalize) and join the one the environment now describes (B200MPI_RANK / B200MPI_WORLD_SIZE / B200MPI_JOB_ID).
// Not MPI: the standard forbids a second MPI_Init, the Horovod front-end needs exactly that.
extern "C" int b200mpi_mpi_reinit(void) {
  if (g_init && !g_final) MPI_Finalize();
  g_init = false;
  g_final = false;
  return MPI_Init(nullptr, nullptr);
}
int MPI_Abort(MPI_Comm, int code) {
  if (g_rv) g_rv->set_abort();
  fprintf(stderr, "[libmpi b200mpi rank %d] MPI_Abort(%d)\n", g_rank, code);
  _exit(code ? code : 1);
}
int MPI_Get_processor_name(char* name, int* len) {
  const char* h = getenv("B200MPI_HOSTNAME");
  char buf[MPI_MAX_PROCESSOR_NAME];
  if (!h || !*h) { gethostname(buf, sizeof(buf)); buf[sizeof(buf) - 1] = 0; h = buf; }
  strncpy(name, h, MPI_MAX_PROCESSOR_NAME - 1);
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  *len = (int)strlen(name);
  return MPI_SUCCESS;
}
int MPI_Get_version(int* v, int* s) { *v = 3; *s = 1; return MPI_SUCCESS; }
int MPI_Comm_get_attr(MPI_Comm c, int keyval, void* attribute_val, int* flag) {
  int e = check(c); if (e) return e;
  static int tag_ub = MPI_TAG_UB, appnum = 0, universe = 1, wtime_global = 1, host = MPI_PROC_NULL, io = 0;
  appnum = env_first({"B200MPI_APPNUM", "OMPI_MCA_orte_app_num", "PMI_APPNUM"}, 0);
  universe = env_first({"OMPI_UNIVERSE_SIZE"}, g_size);
  io = g_rank;
  int* v = nullptr;
  switch (keyval) {
    case MPI_TAG_UB_KEY: case MPI_TAG_UB: v = &tag_ub; break;   // this header defines MPI_TAG_UB as the bound itself; accept it as the key too
    case MPI_APPNUM: v = &appnum; break;
    case MPI_UNIVERSE_SIZE: v = &universe; break;
    case MPI_WTIME_IS_GLOBAL: v = &wtime_global; break;   // one box, one CLOCK_MONOTONIC
    case MPI_HOST: v = &host; break;
    case MPI_IO: v = &io; break;
    default: break;
  }
  *flag = v != nullptr;
  if (v) *static_cast<int**>(attribute_val) = v;
  return MPI_SUCCESS;
}
int MPI_Attr_get(MPI_Comm c, int keyval, void* attribute_val, int* flag) { return MPI_Comm_get_attr(c, keyval, attribute_val, flag); }
int MPI_Get_library_version(char* v, int* len) {
  *len = snprintf(v, 256, "b200mpi libmpi shim 0.1.0 (shm rendezvous transport)");
  return MPI_SUCCESS;
}
int MPI_Op_create(MPI_User_function* fn, int, MPI_Op* op) {
  if (!fn) return MPI_ERR_ARG;
  for (size_t k = 0; k < g_user_ops.size(); k++)
    if (!g_user_ops[k]) { g_user_ops[k] = fn; *op = kFirstUserOp + (int)k; return MPI_SUCCESS; }
  g_user_ops.push_back(fn);
  *op = kFirstUserOp + (int)g_user_ops.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Op_free(MPI_Op* op) {
  const int k = *op - kFirstUserOp;
  if (k >= 0 && k < (int)g_user_ops.size()) g_user_ops[(size_t)k] = nullptr;
  *op = MPI_OP_NULL;
  return MPI_SUCCESS;
}
int MPI_Type_size(MPI_Datatype t, int* s) { *s = (int)type_size(t); return *s ? MPI_SUCCESS : MPI_ERR_TYPE; }
int MPI_Error_string(int code, char* s, int* len) { *len = snprintf(s, MPI_MAX_ERROR_STRING, "MPI error %d", code); return MPI_SUCCESS; }
double MPI_Wtime(void) { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
double MPI_Wtick(void) { return 1e-9; }

// Every collective: singleton communicators are local copies, world-like ones (all ranks in world order) take the
// shared-memory paths below, any other communicator the point-to-point based algorithms of mpi_comm.cc.
int MPI_Barrier(MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  if (C->size() == 1) return MPI_SUCCESS;
  if (!C->world_like) return gen_barrier(C);
  std::string err;
  if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
  return MPI_SUCCESS;
}
int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  if (root < 0 || root >= C->size()) return MPI_ERR_ROOT;
  if (C->size() == 1) return MPI_SUCCESS;
  if (!C->world_like) return gen_bcast(C, buf, es * (size_t)count, root);
  return bcast_bytes(buf, es * (size_t)count, root);
}
// Chunk by chunk: every rank publishes its chunk; with the boxes each rank folds ONE slice of it over all ranks (rank order:
// the result is bit-identical everywhere) into its result box and the ranks that keep the result gather the slices; with the
// mailboxes only, every keeper folds the whole chunk itself. Two barriers per chunk either way.
static int reduce_impl(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op, int root, bool all) {
  const size_t es = type_size(t);
  if (!es) return MPI_ERR_TYPE;
  const size_t bytes = es * (size_t)count;
  const unsigned char* mine = static_cast<const unsigned char*>(send == MPI_IN_PLACE ? recv : send);
  if (g_size == 1) { if (send != MPI_IN_PLACE && (all || g_rank == root)) memmove(recv, send, bytes); return MPI_SUCCESS; }
  const size_t chunk = box_bytes() / 8 * 8;   // a whole number of elements of every supported type
  std::string err;
  const bool keep = all || g_rank == root;
  const size_t W = (size_t)g_size, me = (size_t)g_rank;
  int rc = MPI_SUCCESS;
  for (size_t done = 0; done < bytes; done += chunk) {
    const size_t n = std::min(chunk, bytes - done), ne = n / es;
    unsigned char* out = static_cast<unsigned char*>(recv) + done;
    memcpy(box_data(g_rank), mine + done, n);
    if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
    if (g_boxes) {
      const size_t lo = ne * me / W, hi = ne * (me + 1) / W;
      if (hi > lo) {
        unsigned char* acc = box_result(g_rank) + lo * es;
        memcpy(acc, box_data(0) + lo * es, (hi - lo) * es);
        for (int r = 1; r < g_size; r++)
          if (!reduce_into(acc, box_data(r) + lo * es, hi - lo, t, op)) rc = MPI_ERR_OP;
      }
      if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
      if (keep)
        for (size_t r = 0; r < W; r++) {
          const size_t a = ne * r / W, b = ne * (r + 1) / W;
          if (b > a) memcpy(out + a * es, box_result((int)r) + a * es, (b - a) * es);
        }
    } else {
      if (keep) {
        std::vector<unsigned char> acc(box_data(0), box_data(0) + n);
        for (int r = 1; r < g_size; r++)
          if (!reduce_into(acc.data(), box_data(r), ne, t, op)) rc = MPI_ERR_OP;
        memcpy(out, acc.data(), n);
      }
      if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
    }
  }
  return rc;
}
static int reduce_any(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c, bool all) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  if (!all && (root < 0 || root >= C->size())) return MPI_ERR_ROOT;
  if (C->size() == 1) {
    const size_t es = type_size(t);
    if (!es) return MPI_ERR_TYPE;
    MPI_Datatype b1; size_t c1;
    if (flatten_type(t, (size_t)n, &b1, &c1) && !op_supported(b1, op)) return MPI_ERR_OP;
    if (s != MPI_IN_PLACE) memmove(r, s, es * (size_t)n);
    return MPI_SUCCESS;
  }
  MPI_Datatype base; size_t cnt;
  if (!flatten_type(t, (size_t)n, &base, &cnt)) return MPI_ERR_TYPE;
  if (!op_supported(base, op)) return MPI_ERR_OP;
  if (!C->world_like) return gen_reduce(C, s, r, (size_t)n, t, op, root, all);
  return reduce_impl(s, r, (int)cnt, base, op, root, all);
}
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { return reduce_any(s, r, n, t, op, root, c, false); }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return reduce_any(s, r, n, t, op, 0, c, true); }
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t bytes = s == MPI_IN_PLACE ? type_size(rt) * (size_t)rn : type_size(st) * (size_t)sn;
  if (s == MPI_IN_PLACE) {
    std::vector<unsigned char> mine((unsigned char*)r + (size_t)C->my * bytes, (unsigned char*)r + (size_t)(C->my + 1) * bytes);
    if (C->size() == 1) return MPI_SUCCESS;
    return C->world_like ? allgather_bytes(mine.data(), r, bytes) : gen_allgather(C, mine.data(), r, bytes);
  }
  if (C->size() == 1) { memmove(r, s, bytes); return MPI_SUCCESS; }
  return C->world_like ? allgather_bytes(s, r, bytes) : gen_allgather(C, s, r, bytes);
}
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  if (root < 0 || root >= C->size()) return MPI_ERR_ROOT;
  const bool in_place = s == MPI_IN_PLACE;   // only meaningful at the root: its block is already in place
  const size_t bytes = in_place ? type_size(rt) * (size_t)rn : type_size(st) * (size_t)sn;
  if (C->size() == 1) { if (!in_place) memmove(r, s, bytes); return MPI_SUCCESS; }
  std::vector<unsigned char> all(bytes * (size_t)C->size());
  const void* mine = in_place ? (const char*)r + (size_t)C->my * bytes : s;
  e = C->world_like ? allgather_bytes(mine, all.data(), bytes) : gen_allgather(C, mine, all.data(), bytes);
  if (!e && C->my == root) memcpy(r, all.data(), all.size());
  return e;
}
int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  if (root < 0 || root >= C->size()) return MPI_ERR_ROOT;
  const size_t bytes = C->my == root ? type_size(st) * (size_t)sn : type_size(rt) * (size_t)rn;
  if (C->size() == 1) { if (r != MPI_IN_PLACE) memmove(r, s, bytes); return MPI_SUCCESS; }
  std::vector<unsigned char> all(bytes * (size_t)C->size());
  if (C->my == root) memcpy(all.data(), s, all.size());
  e = C->world_like ? bcast_bytes(all.data(), all.size(), root) : gen_bcast(C, all.data(), all.size(), root);
  if (e) return e;
  if (r != MPI_IN_PLACE) memcpy(r, all.data() + (size_t)C->my * bytes, bytes);
  return MPI_SUCCESS;
}
// Pairwise exchange: in step k rank r publishes the block for rank (r + k) % W in its box and reads the block rank (r - k) % W
// published for it — W - 1 steps of `bytes` per rank instead of gathering every rank's whole send buffer everywhere.
int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  int e = check(c); if (e) return e;
  Comm* C = comm_of(c);
  const size_t bytes = s == MPI_IN_PLACE ? type_size(rt) * (size_t)rn : type_size(st) * (size_t)sn;
  if (C->size() == 1) { if (s != MPI_IN_PLACE && r != s) memmove(r, s, bytes); return MPI_SUCCESS; }
  std::vector<unsigned char> tmp;
  const unsigned char* src = static_cast<const unsigned char*>(s);
  if (s == MPI_IN_PLACE) { tmp.assign((unsigned char*)r, (unsigned char*)r + bytes * (size_t)C->size()); src = tmp.data(); }
  if (!C->world_like) return gen_alltoall(C, src, r, bytes);
  memcpy((char*)r + (size_t)g_rank * bytes, src + (size_t)g_rank * bytes, bytes);
  std::string err;
  const size_t chunk = box_bytes();
  for (int step = 1; step < g_size; step++) {
    const int dst = (g_rank + step) % g_size, from = (g_rank - step + g_size) % g_size;
    for (size_t done = 0; done < bytes; done += chunk) {
      const size_t n = std::min(chunk, bytes - done);
      memcpy(box_data(g_rank), src + (size_t)dst * bytes + done, n);
      if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
      memcpy((char*)r + (size_t)from * bytes + done, box_data(from), n);
      if (g_rv->barrier(g_timeout_ms, &err)) return fail(err);
    }
  }
  return MPI_SUCCESS;
}

}  // extern "C"
