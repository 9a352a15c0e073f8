// libb200mpi_nccl.so — NCCL C-ABI front-end of the b200mpi runtime, meant to be
// LD_PRELOADed into unmodified torch.distributed / Horovod processes.
//
// PyTorch's libtorch_cuda.so links libnccl.so.2 dynamically, so preloaded
// definitions of ncclAllReduce & co. win symbol resolution: the gradient
// allreduce / allgather / broadcast a training script issues through
// ProcessGroupNCCL then execute as b200mpi peer-memory / NVLS kernels with the
// scale (ncclAvg, PreMulSum) fused in — "LD-injected collective runtime" of the
// north star (BASELINE.json; SURVEY.md §5.9 front-end (2), §7.3 item 4).
// Anything outside the implemented subset (P2P send/recv, >8 ranks, exotic
// dtypes/ops on a failed fast path) is forwarded to the real NCCL found with
// dlsym(RTLD_NEXT); B200MPI_ALGO=nccl forwards everything (the baseline mode).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../include/b200mpi.h"

extern "C" {
// ---- the slice of nccl.h we implement (ABI-compatible declarations) -------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3,
               ncclInvalidArgument = 4, ncclInvalidUsage = 5, ncclRemoteError = 6, ncclInProgress = 7 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4, ncclNumOps = 5 } ncclRedOp_t;
typedef enum { ncclScalarDevice = 0, ncclScalarHostImmediate = 1 } ncclScalarResidence_t;
struct ncclConfig_v;  // opaque
}

namespace {

struct PendingP2P { b200mpi_p2p_op_t op; cudaStream_t stream; };
constexpr uint64_t kShimMagic = 0xB200C0DEC0117EC7ull;
struct Shim {
  uint64_t magic = kShimMagic;    // first word: tells our handles from real ncclComm pointers (see S())
  std::vector<PendingP2P> p2p;    // sends/receives queued inside ncclGroupStart/End (flushed as ONE kernel)
  b200mpi_comm_t mine = nullptr;  // b200mpi communicator (nullptr => forwarded)
  ncclComm_t real = nullptr;      // real NCCL communicator when forwarding
  int rank = 0, world = 1, device = 0;
  std::string id;
  int splits = 0;
  std::map<void*, int> registered;   // base of an ncclMemAlloc allocation -> adopted b200mpi window
};

std::mutex g_mu;
std::map<int, float> g_premul;  // dynamic ncclRedOp_t -> scalar
int g_next_op = 16;
std::atomic<uint64_t> g_calls{0}, g_forwarded{0}, g_registered{0};
thread_local std::string g_last_error;
thread_local int g_group_depth = 0;
thread_local int g_group_fwd = 0;
thread_local std::vector<Shim*> g_group_p2p;  // communicators with queued point-to-point operations

bool forward_all() {
  static int v = -1;
  if (v < 0) {
    const char* a = getenv("B200MPI_ALGO");
    v = (a && strcmp(a, "nccl") == 0) ? 1 : 0;
  }
  return v == 1;
}
bool debug() { static int v = getenv("B200MPI_DEBUG") ? 1 : 0; return v; }
// B200MPI_DEBUG=2: one stderr line per NCCL entry (with tests/mp_launch.py --log-dir: the last call of a failing rank)
int trace_level() { static int v = getenv("B200MPI_DEBUG") ? atoi(getenv("B200MPI_DEBUG")) : 0; return v; }
#define TRACE_CALL(fmt, ...)                                                                                   \
  do {                                                                                                         \
    if (trace_level() >= 2) fprintf(stderr, "[b200mpi nccl shim %d] %s " fmt "\n", (int)getpid(), __func__, ##__VA_ARGS__); \
  } while (0)

// libnccl calls some of its own public entry points through the PLT (objdump -R libnccl.so.2: ncclBroadcast,
// ncclCommGetAsyncError, ncclCommRegister/Deregister, ncclCommWindowDeregister, ncclDevCommDestroy, ncclMemAlloc/Free,
// ncclGetUniqueId, ncclGetVersion ...). With this library preloaded those internal calls land HERE, carrying real
// ncclComm pointers. Two guards keep that transparent: (1) every forwarded call runs under a depth counter, and while
// it is non-zero the comm-less entry points forward verbatim; (2) handles that do not start with kShimMagic are treated
// as real communicators and forwarded untouched (S() below).
thread_local int g_in_real = 0;
template <typename F>
struct RealCall;
template <typename R, typename... A>
struct RealCall<R (*)(A...)> {
  R (*fn)(A...);
  explicit operator bool() const { return fn != nullptr; }
  R operator()(A... a) const {
    g_in_real++;
    R r = fn(a...);
    g_in_real--;
    return r;
  }
};
template <typename F>
RealCall<F> real_fn(const char* name) {
  void* p = dlsym(RTLD_NEXT, name);
  if (!p) {
    static void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h) p = dlsym(h, name);
  }
  return RealCall<F>{reinterpret_cast<F>(p)};
}
#define REAL(name, ...) real_fn<ncclResult_t (*)(__VA_ARGS__)>(#name)

ncclResult_t err(ncclResult_t code, const std::string& msg) {
  g_last_error = msg;
  if (debug()) fprintf(stderr, "[b200mpi nccl shim] %s\n", msg.c_str());
  return code;
}
ncclResult_t from_rc(int rc, const char* what) {
  if (rc == 0) return ncclSuccess;
  return err(rc == B200MPI_ERR_CUDA ? ncclUnhandledCudaError : (rc == B200MPI_ERR_INVALID ? ncclInvalidArgument : ncclSystemError),
             std::string(what) + ": " + b200mpi_last_error());
}

size_t dt_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: case ncclBfloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    case ncclInt64: case ncclUint64: case ncclFloat64: return 8;
    default: return (int)t >= 10 && (int)t <= 11 ? 1 : 0;  // fp8 variants: byte-wise only
  }
}
bool native_dt(ncclDataType_t t, b200mpi_dtype_t* out) {
  if (t == ncclFloat32) { *out = B200MPI_F32; return true; }
  if (t == ncclBfloat16) { *out = B200MPI_BF16; return true; }
  if (t == ncclFloat16) { *out = B200MPI_F16; return true; }
  return false;
}
// -> (op, scale) for the fused kernels; false when the generic path is needed
bool native_op(int op, int world, b200mpi_op_t* out, float* scale) {
  *scale = 1.0f;
  if (op == ncclSum) { *out = B200MPI_SUM; return true; }
  if (op == ncclAvg) { *out = B200MPI_SUM; *scale = 1.0f / world; return true; }
  if (op == ncclMax) { *out = B200MPI_MAX; return true; }
  if (op == ncclMin) { *out = B200MPI_MIN; return true; }
  if (op >= 16) {
    std::lock_guard<std::mutex> l(g_mu);
    auto it = g_premul.find(op);
    if (it != g_premul.end()) { *out = B200MPI_SUM; *scale = it->second; return true; }
  }
  return false;
}

// generic reduction of `world` gathered blocks for dtypes/ops the fused kernels do not cover
template <typename T>
__global__ void k_reduce_gathered(const T* __restrict__ in, T* __restrict__ out, size_t n, size_t stride, int world, int op, double post) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    T acc = in[i];
    for (int r = 1; r < world; r++) {
      const T x = in[(size_t)r * stride + i];
      acc = op == ncclProd ? (T)(acc * x) : (op == ncclMax ? (x > acc ? x : acc) : (op == ncclMin ? (x < acc ? x : acc) : (T)(acc + x)));
    }
    if (op == ncclAvg) acc = (T)(acc / (T)world);
    if (post != 1.0) acc = (T)((double)acc * post);
    out[i] = acc;
  }
}
template <typename T>
void launch_rg(const void* in, void* out, size_t n, size_t stride, int world, int op, double post, cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  k_reduce_gathered<T><<<blocks ? blocks : 1, 256, 0, s>>>((const T*)in, (T*)out, n, stride, world, op, post);
}

ncclResult_t generic_allreduce(Shim* s, const void* send, void* recv, size_t count, ncclDataType_t dt, int op, cudaStream_t st) {
  const size_t es = dt_size(dt);
  if (!es) return err(ncclInvalidArgument, "unsupported datatype");
  double post = 1.0;
  int eff = op;
  if (op >= 16) { std::lock_guard<std::mutex> l(g_mu); auto it = g_premul.find(op); if (it == g_premul.end()) return err(ncclInvalidArgument, "unknown reduction op"); post = it->second; eff = ncclSum; }
  void* tmp = nullptr;
  size_t bytes = count * es;
  size_t padded = (bytes + 1) / 2 * 2;
  if (cudaMallocAsync(&tmp, padded * s->world, st) != cudaSuccess) return err(ncclUnhandledCudaError, "cudaMallocAsync failed");
  const void* src = send;
  void* pad_src = nullptr;
  if (padded != bytes) {  // odd byte counts: the byte-wise allgather moves 2-byte units
    cudaMallocAsync(&pad_src, padded, st);
    cudaMemsetAsync(pad_src, 0, padded, st);
    cudaMemcpyAsync(pad_src, send, bytes, cudaMemcpyDeviceToDevice, st);
    src = pad_src;
  }
  int rc = b200mpi_allgather(s->mine, src, tmp, padded / 2, B200MPI_BF16, st);
  if (rc) { cudaFreeAsync(tmp, st); return from_rc(rc, "allgather"); }
  if (padded != bytes) {
    // gathered blocks are `padded` apart: compact is unnecessary when we reduce with stride = padded/es only if divisible
    cudaFreeAsync(pad_src, st);
    if (padded % es) { cudaFreeAsync(tmp, st); return err(ncclInvalidArgument, "odd-sized reduction not supported"); }
  }
  const size_t stride_elems = padded / es;
  switch (dt) {
    case ncclInt8: launch_rg<int8_t>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclUint8: launch_rg<uint8_t>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclInt32: launch_rg<int32_t>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclUint32: launch_rg<uint32_t>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclInt64: launch_rg<long long>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclUint64: launch_rg<unsigned long long>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclFloat64: launch_rg<double>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    case ncclFloat32: launch_rg<float>(tmp, recv, count, stride_elems, s->world, eff, post, st); break;
    default: cudaFreeAsync(tmp, st); return err(ncclInvalidArgument, "reduction op not supported for this datatype");
  }
  cudaFreeAsync(tmp, st);
  return cudaGetLastError() == cudaSuccess ? ncclSuccess : err(ncclUnhandledCudaError, "reduce_gathered launch failed");
}

std::string id_to_job(const ncclUniqueId* id) {
  char buf[80];
  const unsigned char* b = reinterpret_cast<const unsigned char*>(id->internal);
  int n = snprintf(buf, sizeof(buf), "nccl-");
  for (int i = 0; i < 16; i++) n += snprintf(buf + n, sizeof(buf) - n, "%02x", b[i]);
  return buf;
}

ncclResult_t init_common(ncclComm_t* out, int nranks, const ncclUniqueId* id, int rank, const void* config,
                         bool have_config) {
  int dev = 0;
  cudaGetDevice(&dev);
  TRACE_CALL("rank=%d/%d device=%d id=%s", rank, nranks, dev, id_to_job(id).c_str());
  auto* s = new Shim;
  s->rank = rank; s->world = nranks; s->device = dev; s->id = id_to_job(id);
  if (!forward_all() && nranks <= B200MPI_MAX_RANKS) {
    // torch ranks reach their first collective at very different times (model build, cuDNN init, page-cache misses on a
    // fresh box): give the rendezvous longer than the collective watchdog unless the user set a timeout
    if (!getenv("B200MPI_TIMEOUT_MS")) setenv("B200MPI_INIT_TIMEOUT_MS", "180000", 0);
    int rc = b200mpi_comm_init(&s->mine, rank, nranks, dev, s->id.c_str(), 0, 0);
    if (rc != 0) {
      // The unique id was minted by this shim (ncclGetUniqueId below), real NCCL cannot bootstrap from it, and a rank that
      // fell back alone would deadlock its peers: fail loudly, on every rank that sees the failure.
      fprintf(stderr, "[b200mpi nccl shim] rank %d/%d: communicator %s could not be created on the b200mpi runtime: %s\n"
                      "[b200mpi nccl shim] set B200MPI_ALGO=nccl to run this job on stock NCCL, B200MPI_DEBUG=1 for details\n",
              rank, nranks, s->id.c_str(), b200mpi_last_error());
      std::string why = std::string("b200mpi communicator init failed: ") + b200mpi_last_error();
      delete s;
      return err(ncclSystemError, why.c_str());
    }
  }
  if (!s->mine) {
    g_forwarded++;
    ncclResult_t r;
    if (have_config) {
      auto f = REAL(ncclCommInitRankConfig, ncclComm_t*, int, ncclUniqueId, int, const void*);
      if (!f) { delete s; return err(ncclSystemError, "real NCCL not found for pass-through"); }
      r = f(&s->real, nranks, *id, rank, config);
    } else {
      auto f = REAL(ncclCommInitRank, ncclComm_t*, int, ncclUniqueId, int);
      if (!f) { delete s; return err(ncclSystemError, "real NCCL not found for pass-through"); }
      r = f(&s->real, nranks, *id, rank);
    }
    if (r != ncclSuccess && r != ncclInProgress) { delete s; return r; }
    *out = reinterpret_cast<ncclComm_t>(s);
    return r;  // non-blocking communicators (ncclConfig_t.blocking = 0) report ncclInProgress: pass it on
  } else if (debug() && rank == 0) {
    fprintf(stderr, "[b200mpi nccl shim] communicator %s: %d ranks on b200mpi kernels (NVLS=%d)\n", s->id.c_str(), nranks,
            b200mpi_comm_has_multicast(s->mine));
  }
  *out = reinterpret_cast<ncclComm_t>(s);
  return ncclSuccess;
}

inline bool is_shim(const void* c) { return c && *reinterpret_cast<const uint64_t*>(c) == kShimMagic; }
// Our handle, or — for a real ncclComm that reached us through libnccl's own PLT — a per-thread stand-in whose `real`
// is that communicator, so every `if (s->real) forward` path below forwards it unchanged.
inline Shim* S(ncclComm_t c) {
  if (is_shim(c)) return reinterpret_cast<Shim*>(c);
  thread_local Shim foreign;
  foreign.magic = 0;
  foreign.real = c;
  foreign.mine = nullptr;
  return &foreign;
}

}  // namespace

extern "C" {

// exported for tests / stats
uint64_t b200mpi_shim_calls(void) { return g_calls.load(); }
uint64_t b200mpi_shim_forwarded(void) { return g_forwarded.load(); }
uint64_t b200mpi_shim_registered(void) { return g_registered.load(); }   // ncclMemAlloc buffers adopted as NVLS windows

ncclResult_t ncclGetVersion(int* v) {
  if (forward_all() || g_in_real > 0 || g_forwarded.load()) {
    if (auto f = REAL(ncclGetVersion, int*)) return f(v);
  }
  *v = 22809;  // the NCCL ABI level this shim implements
  return ncclSuccess;
}
const char* ncclGetErrorString(ncclResult_t r) {
  switch (r) {
    case ncclSuccess: return "no error";
    case ncclUnhandledCudaError: return "unhandled cuda error (run with B200MPI_DEBUG=1 for details)";
    case ncclSystemError: return "unhandled system error (run with B200MPI_DEBUG=1 for details)";
    case ncclInternalError: return "internal error";
    case ncclInvalidArgument: return "invalid argument";
    case ncclInvalidUsage: return "invalid usage";
    case ncclRemoteError: return "remote process exited or there was a network error";
    case ncclInProgress: return "NCCL operation in progress";
    default: return "unknown result code";
  }
}
const char* ncclGetLastError(ncclComm_t) { return g_last_error.c_str(); }

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  if (forward_all() || g_in_real > 0) {
    auto f = REAL(ncclGetUniqueId, ncclUniqueId*);
    if (f) return f(id);
  }
  memset(id, 0, sizeof(*id));
  FILE* f = fopen("/dev/urandom", "rb");
  size_t got = f ? fread(id->internal, 1, 32, f) : 0;
  if (f) fclose(f);
  if (got < 32) { uint64_t t = (uint64_t)time(nullptr) ^ ((uint64_t)getpid() << 32); memcpy(id->internal, &t, 8); }
  memcpy(id->internal + 32, "b200mpi", 8);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) { return init_common(comm, nranks, &id, rank, nullptr, false); }
ncclResult_t ncclCommInitRankConfig(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank, const void* config) { return init_common(comm, nranks, &id, rank, config, true); }
ncclResult_t ncclCommInitRankScalable(ncclComm_t* comm, int nranks, int rank, int nid, ncclUniqueId* ids, const void* config) {
  if (nid < 1) return err(ncclInvalidArgument, "no unique ids");
  return init_common(comm, nranks, &ids[0], rank, config, true);
}
ncclResult_t ncclCommInitAll(ncclComm_t* comms, int ndev, const int* devlist) {
  auto f = REAL(ncclCommInitAll, ncclComm_t*, int, const int*);
  if (!f) return err(ncclInvalidUsage, "ncclCommInitAll (single process, many GPUs) is not provided: b200mpi is one process per GPU");
  std::vector<ncclComm_t> real(ndev);
  ncclResult_t r = f(real.data(), ndev, devlist);
  if (r != ncclSuccess) return r;
  for (int i = 0; i < ndev; i++) { auto* s = new Shim; s->real = real[i]; s->rank = i; s->world = ndev; comms[i] = reinterpret_cast<ncclComm_t>(s); }
  g_forwarded++;
  return ncclSuccess;
}

ncclResult_t ncclCommSplit(ncclComm_t comm, int color, int key, ncclComm_t* newcomm, const void* config) {
  if (!is_shim(comm)) {  // internal call of real NCCL: its caller expects a real communicator back
    auto f = REAL(ncclCommSplit, ncclComm_t, int, int, ncclComm_t*, const void*);
    return f ? f(comm, color, key, newcomm, config) : err(ncclSystemError, "real ncclCommSplit missing");
  }
  Shim* s = S(comm);
  if (s->real) {
    auto f = REAL(ncclCommSplit, ncclComm_t, int, int, ncclComm_t*, const void*);
    if (!f) return err(ncclSystemError, "real ncclCommSplit missing");
    ncclComm_t r = nullptr;
    ncclResult_t rc = f(s->real, color, key, &r, config);
    if (rc != ncclSuccess && rc != ncclInProgress) return rc;
    if (!r) { *newcomm = nullptr; return rc; }
    auto* n = new Shim; n->real = r; *newcomm = reinterpret_cast<ncclComm_t>(n);
    return rc;
  }
  struct CK { int color, key, rank; } mine{color, key, s->rank};
  std::vector<CK> all(s->world);
  if (b200mpi_comm_host_allgather(s->mine, &mine, all.data(), sizeof(CK))) return from_rc(-1, "split allgather");
  const int seq = s->splits++;
  if (color < 0) { *newcomm = nullptr; return ncclSuccess; }
  std::vector<CK> members;
  for (auto& x : all) if (x.color == color) members.push_back(x);
  std::stable_sort(members.begin(), members.end(), [](const CK& a, const CK& b) { return a.key != b.key ? a.key < b.key : a.rank < b.rank; });
  int nrank = -1;
  for (size_t i = 0; i < members.size(); i++) if (members[i].rank == s->rank) nrank = (int)i;
  auto* n = new Shim;
  n->rank = nrank; n->world = (int)members.size(); n->device = s->device;
  n->id = s->id + "-s" + std::to_string(seq) + "c" + std::to_string(color);
  int rc = b200mpi_comm_init(&n->mine, nrank, n->world, n->device, n->id.c_str(), 0, 0);
  if (rc) { delete n; return from_rc(rc, "ncclCommSplit"); }
  *newcomm = reinterpret_cast<ncclComm_t>(n);
  return ncclSuccess;
}
ncclResult_t ncclCommShrink(ncclComm_t, int*, int, ncclComm_t*, const void*, int) { return err(ncclInvalidUsage, "ncclCommShrink is not provided; re-form the communicator (elastic rescale re-spawns ranks)"); }

static ncclResult_t destroy(ncclComm_t comm, const char* realname) {
  TRACE_CALL("%s", realname);
  if (!comm) return ncclSuccess;
  if (!is_shim(comm)) { auto f = real_fn<ncclResult_t (*)(ncclComm_t)>(realname); return f ? f(comm) : ncclSuccess; }
  Shim* s = S(comm);
  ncclResult_t r = ncclSuccess;
  if (s->real) { auto f = real_fn<ncclResult_t (*)(ncclComm_t)>(realname); if (f) r = f(s->real); }
  if (s->mine) {
    if (debug() && s->rank == 0) fprintf(stderr, "[b200mpi nccl shim] %s: %llu b200mpi kernel launches\n", s->id.c_str(), (unsigned long long)b200mpi_comm_launch_count(s->mine));
    b200mpi_comm_destroy(s->mine);
  }
  s->magic = 0;
  delete s;
  return r;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { return destroy(c, "ncclCommDestroy"); }
ncclResult_t ncclCommAbort(ncclComm_t c) { return destroy(c, "ncclCommAbort"); }
ncclResult_t ncclCommFinalize(ncclComm_t c) {
  Shim* s = S(c);
  if (s && s->real) { auto f = REAL(ncclCommFinalize, ncclComm_t); return f ? f(s->real) : ncclSuccess; }
  return ncclSuccess;
}
ncclResult_t ncclCommCount(const ncclComm_t c, int* n) { Shim* s = S(c); if (s->real) { auto f = REAL(ncclCommCount, ncclComm_t, int*); return f(s->real, n); } *n = s->world; return ncclSuccess; }
ncclResult_t ncclCommUserRank(const ncclComm_t c, int* r) { Shim* s = S(c); if (s->real) { auto f = REAL(ncclCommUserRank, ncclComm_t, int*); return f(s->real, r); } *r = s->rank; return ncclSuccess; }
ncclResult_t ncclCommCuDevice(const ncclComm_t c, int* d) { Shim* s = S(c); if (s->real) { auto f = REAL(ncclCommCuDevice, ncclComm_t, int*); return f(s->real, d); } *d = s->device; return ncclSuccess; }
ncclResult_t ncclCommGetAsyncError(ncclComm_t c, ncclResult_t* e) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclCommGetAsyncError, ncclComm_t, ncclResult_t*); return f(s->real, e); }
  *e = b200mpi_comm_check_error(s->mine) ? ncclRemoteError : ncclSuccess;
  return ncclSuccess;
}

ncclResult_t ncclRedOpCreatePreMulSum(ncclRedOp_t* op, void* scalar, ncclDataType_t dt, ncclScalarResidence_t res, ncclComm_t c) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclRedOpCreatePreMulSum, ncclRedOp_t*, void*, ncclDataType_t, ncclScalarResidence_t, ncclComm_t); return f(op, scalar, dt, res, s->real); }
  float v;
  if (res == ncclScalarHostImmediate) {
    if (dt == ncclFloat32) v = *(float*)scalar;
    else if (dt == ncclFloat64) v = (float)*(double*)scalar;
    else if (dt == ncclFloat16) v = __half2float(*(__half*)scalar);
    else if (dt == ncclBfloat16) v = __bfloat162float(*(__nv_bfloat16*)scalar);
    else return err(ncclInvalidArgument, "PreMulSum scalar dtype");
  } else {
    if (dt != ncclFloat32) return err(ncclInvalidArgument, "device-resident PreMulSum scalar must be float32");
    if (cudaMemcpy(&v, scalar, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return err(ncclUnhandledCudaError, "reading PreMulSum scalar");
  }
  std::lock_guard<std::mutex> l(g_mu);
  const int id = g_next_op++;
  g_premul[id] = v;
  *op = (ncclRedOp_t)id;
  return ncclSuccess;
}
ncclResult_t ncclRedOpDestroy(ncclRedOp_t op, ncclComm_t c) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclRedOpDestroy, ncclRedOp_t, ncclComm_t); return f(op, s->real); }
  std::lock_guard<std::mutex> l(g_mu);
  g_premul.erase((int)op);
  return ncclSuccess;
}

// Point-to-point on a b200mpi communicator: everything queued on one communicator becomes one b200mpi_p2p_batch (one
// kernel, one CTA per operation). Self-sends are matched with self-receives and done as device copies. Operations queued
// on different streams are serialised onto the first one with events.
static ncclResult_t flush_p2p(Shim* s) {
  std::vector<PendingP2P> q;
  q.swap(s->p2p);
  if (q.empty()) return ncclSuccess;
  cudaStream_t st = q[0].stream;
  for (auto& p : q) {
    if (p.stream != st) {
      cudaEvent_t ev;
      cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      cudaEventRecord(ev, p.stream);
      cudaStreamWaitEvent(st, ev, 0);
      cudaEventDestroy(ev);
    }
  }
  std::vector<b200mpi_p2p_op_t> ops;
  std::vector<const PendingP2P*> self_send, self_recv;
  for (auto& p : q) {
    if (p.op.peer == s->rank) (p.op.is_send ? self_send : self_recv).push_back(&p);
    else ops.push_back(p.op);
  }
  if (self_send.size() != self_recv.size()) return err(ncclInvalidUsage, "unmatched send/recv to self inside a group");
  for (size_t i = 0; i < self_send.size(); i++) {
    if (self_send[i]->op.bytes != self_recv[i]->op.bytes) return err(ncclInvalidArgument, "self send/recv size mismatch");
    if (self_send[i]->op.bytes) cudaMemcpyAsync(self_recv[i]->op.recv, self_send[i]->op.send, self_send[i]->op.bytes, cudaMemcpyDeviceToDevice, st);
  }
  ncclResult_t r = ncclSuccess;
  if (!ops.empty()) r = from_rc(b200mpi_p2p_batch(s->mine, ops.data(), (int)ops.size(), st), "ncclSend/ncclRecv");
  if (r == ncclSuccess) {
    for (auto& p : q) {  // later work on the other streams must see the exchange
      if (p.stream != st) {
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        cudaEventRecord(ev, st);
        cudaStreamWaitEvent(p.stream, ev, 0);
        cudaEventDestroy(ev);
      }
    }
  }
  return r;
}
static ncclResult_t queue_p2p(Shim* s, const b200mpi_p2p_op_t& op, cudaStream_t st) {
  g_calls++;
  if (!b200mpi_comm_has_p2p(s->mine))
    return err(ncclInvalidUsage, "point-to-point on a b200mpi communicator is experimental: set B200MPI_P2P=1 for the job "
                                 "(or B200MPI_ALGO=nccl to run this process group on stock NCCL)");
  s->p2p.push_back(PendingP2P{op, st});
  if (g_group_depth > 0) {
    if (std::find(g_group_p2p.begin(), g_group_p2p.end(), s) == g_group_p2p.end()) g_group_p2p.push_back(s);
    return ncclSuccess;
  }
  return flush_p2p(s);
}

ncclResult_t ncclGroupStart(void) {
  TRACE_CALL("depth=%d", g_group_depth + 1);
  g_group_depth++;
  // Forward only when a real communicator exists (or everything is forwarded); remember it so the
  // matching End is forwarded too and the real library never sees an unbalanced pair.
  if (forward_all() || g_forwarded.load() || g_in_real > 0) {
    if (auto f = real_fn<ncclResult_t (*)()>("ncclGroupStart")) { g_group_fwd++; return f(); }
  }
  return ncclSuccess;
}
ncclResult_t ncclGroupEnd(void) {
  if (g_group_depth > 0) g_group_depth--;
  if (g_group_depth == 0 && !g_group_p2p.empty()) {
    std::vector<Shim*> todo;
    todo.swap(g_group_p2p);
    for (Shim* s : todo) {
      ncclResult_t r = flush_p2p(s);
      if (r != ncclSuccess) return r;
    }
  }
  if (g_group_fwd > 0) {
    g_group_fwd--;
    if (auto f = real_fn<ncclResult_t (*)()>("ncclGroupEnd")) return f();
  }
  return ncclSuccess;  // b200mpi collectives were enqueued eagerly, in order, on their streams
}
ncclResult_t ncclGroupSimulateEnd(void*) { return ncclSuccess; }

ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  TRACE_CALL("count=%zu dt=%d op=%d rank=%d/%d %s", count, (int)dt, (int)op, s->rank, s->world, s->real ? "fwd" : "b200mpi");
  g_calls++;
  if (s->real) { auto f = REAL(ncclAllReduce, const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t); return f(send, recv, count, dt, op, s->real, st); }
  b200mpi_dtype_t bdt; b200mpi_op_t bop; float scale;
  if (native_dt(dt, &bdt) && native_op((int)op, s->world, &bop, &scale))
    return from_rc(b200mpi_allreduce(s->mine, send, recv, count, bdt, bop, scale, B200MPI_ALGO_AUTO, st), "ncclAllReduce");
  return generic_allreduce(s, send, recv, count, dt, (int)op, st);
}

ncclResult_t ncclBroadcast(const void* send, void* recv, size_t count, ncclDataType_t dt, int root, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  TRACE_CALL("count=%zu dt=%d root=%d rank=%d/%d", count, (int)dt, root, s->rank, s->world);
  g_calls++;
  if (s->real) { auto f = REAL(ncclBroadcast, const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t); return f(send, recv, count, dt, root, s->real, st); }
  const size_t bytes = count * dt_size(dt);
  if (s->rank == root && send != recv) cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st);
  return from_rc(b200mpi_broadcast_bytes(s->mine, recv, bytes, root, st), "ncclBroadcast");
}
ncclResult_t ncclBcast(void* buf, size_t count, ncclDataType_t dt, int root, ncclComm_t c, cudaStream_t st) { return ncclBroadcast(buf, buf, count, dt, root, c, st); }

ncclResult_t ncclReduce(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, int root, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  TRACE_CALL("count=%zu dt=%d op=%d root=%d rank=%d/%d", count, (int)dt, (int)op, root, s->rank, s->world);
  g_calls++;
  if (s->real) { auto f = REAL(ncclReduce, const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t); return f(send, recv, count, dt, op, root, s->real, st); }
  b200mpi_dtype_t bdt; b200mpi_op_t bop; float scale;
  if (native_dt(dt, &bdt) && native_op((int)op, s->world, &bop, &scale))
    return from_rc(b200mpi_reduce(s->mine, send, recv, count, bdt, bop, scale, root, st), "ncclReduce");
  // generic: allreduce into a scratch buffer, keep it on the root only
  void* tmp = recv;
  if (s->rank != root) { if (cudaMallocAsync(&tmp, count * dt_size(dt), st) != cudaSuccess) return err(ncclUnhandledCudaError, "scratch"); }
  ncclResult_t r = generic_allreduce(s, send, tmp, count, dt, (int)op, st);
  if (s->rank != root) cudaFreeAsync(tmp, st);
  return r;
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t sendcount, ncclDataType_t dt, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  TRACE_CALL("rank=%d/%d", s->rank, s->world);
  g_calls++;
  if (s->real) { auto f = REAL(ncclAllGather, const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t); return f(send, recv, sendcount, dt, s->real, st); }
  const size_t bytes = sendcount * dt_size(dt);
  if (bytes % 2 == 0) return from_rc(b200mpi_allgather(s->mine, send, recv, bytes / 2, B200MPI_BF16, st), "ncclAllGather");
  // odd byte count: go through padded scratch
  void *in = nullptr, *out = nullptr;
  cudaMallocAsync(&in, bytes + 1, st);
  cudaMallocAsync(&out, (bytes + 1) * s->world, st);
  cudaMemsetAsync(in, 0, bytes + 1, st);
  cudaMemcpyAsync(in, send, bytes, cudaMemcpyDeviceToDevice, st);
  int rc = b200mpi_allgather(s->mine, in, out, (bytes + 1) / 2, B200MPI_BF16, st);
  if (!rc) cudaMemcpy2DAsync(recv, bytes, out, bytes + 1, bytes, s->world, cudaMemcpyDeviceToDevice, st);
  cudaFreeAsync(in, st);
  cudaFreeAsync(out, st);
  return from_rc(rc, "ncclAllGather");
}

ncclResult_t ncclReduceScatter(const void* send, void* recv, size_t recvcount, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  TRACE_CALL("rank=%d/%d", s->rank, s->world);
  g_calls++;
  if (s->real) { auto f = REAL(ncclReduceScatter, const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t); return f(send, recv, recvcount, dt, op, s->real, st); }
  b200mpi_dtype_t bdt; b200mpi_op_t bop; float scale;
  if (native_dt(dt, &bdt) && native_op((int)op, s->world, &bop, &scale))
    return from_rc(b200mpi_reduce_scatter(s->mine, send, recv, recvcount, bdt, bop, scale, st), "ncclReduceScatter");
  const size_t es = dt_size(dt);
  void* tmp = nullptr;
  if (cudaMallocAsync(&tmp, recvcount * s->world * es, st) != cudaSuccess) return err(ncclUnhandledCudaError, "scratch");
  ncclResult_t r = generic_allreduce(s, send, tmp, recvcount * s->world, dt, (int)op, st);
  if (r == ncclSuccess) cudaMemcpyAsync(recv, (char*)tmp + (size_t)s->rank * recvcount * es, recvcount * es, cudaMemcpyDeviceToDevice, st);
  cudaFreeAsync(tmp, st);
  return r;
}

ncclResult_t ncclAlltoAll(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  g_calls++;
  if (s->real) { auto f = REAL(ncclAlltoAll, const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t); return f ? f(send, recv, count, dt, s->real, st) : err(ncclInvalidUsage, "ncclAlltoAll missing in real NCCL"); }
  const size_t bytes = count * dt_size(dt);
  if (bytes % 2) return err(ncclInvalidArgument, "alltoall payload must be an even number of bytes");
  return from_rc(b200mpi_alltoall(s->mine, send, recv, bytes / 2, B200MPI_BF16, st), "ncclAlltoAll");
}

// NCCL 2.28 rooted collectives. Built from the staged all-gather / broadcast kernels: every rank moves the full payload, which
// is fine for the metadata-sized tensors frameworks use them for.
ncclResult_t ncclGather(const void* send, void* recv, size_t count, ncclDataType_t dt, int root, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  g_calls++;
  if (s->real) { auto f = REAL(ncclGather, const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t); return f ? f(send, recv, count, dt, root, s->real, st) : err(ncclInvalidUsage, "ncclGather missing in real NCCL"); }
  const size_t bytes = count * dt_size(dt);
  if (bytes % 2) return err(ncclInvalidArgument, "gather payload must be an even number of bytes");
  if (s->rank == root) return from_rc(b200mpi_allgather(s->mine, send, recv, bytes / 2, B200MPI_BF16, st), "ncclGather");
  void* tmp = nullptr;
  if (cudaMallocAsync(&tmp, bytes * s->world, st) != cudaSuccess) return err(ncclUnhandledCudaError, "cudaMallocAsync failed");
  int rc = b200mpi_allgather(s->mine, send, tmp, bytes / 2, B200MPI_BF16, st);
  cudaFreeAsync(tmp, st);
  return from_rc(rc, "ncclGather");
}
ncclResult_t ncclScatter(const void* send, void* recv, size_t count, ncclDataType_t dt, int root, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  g_calls++;
  if (s->real) { auto f = REAL(ncclScatter, const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t); return f ? f(send, recv, count, dt, root, s->real, st) : err(ncclInvalidUsage, "ncclScatter missing in real NCCL"); }
  const size_t bytes = count * dt_size(dt);
  if (bytes % 2) return err(ncclInvalidArgument, "scatter payload must be an even number of bytes");
  void* tmp = nullptr;
  if (cudaMallocAsync(&tmp, bytes * s->world, st) != cudaSuccess) return err(ncclUnhandledCudaError, "cudaMallocAsync failed");
  if (s->rank == root) cudaMemcpyAsync(tmp, send, bytes * s->world, cudaMemcpyDeviceToDevice, st);
  int rc = b200mpi_broadcast_bytes(s->mine, tmp, bytes * s->world, root, st);
  if (!rc) cudaMemcpyAsync(recv, (const char*)tmp + (size_t)s->rank * bytes, bytes, cudaMemcpyDeviceToDevice, st);
  cudaFreeAsync(tmp, st);
  return from_rc(rc, "ncclScatter");
}
ncclResult_t ncclCommRevoke(ncclComm_t c, int flags) {
  Shim* s = S(c);
  if (s && s->real) { auto f = REAL(ncclCommRevoke, ncclComm_t, int); return f ? f(s->real, flags) : err(ncclInvalidUsage, "ncclCommRevoke missing in real NCCL"); }
  return ncclSuccess;  // nothing in flight on the host side: kernels are stream-ordered and bounded by the device watchdog
}

ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclSend, const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t); return f(buf, count, dt, peer, s->real, st); }
  return queue_p2p(s, b200mpi_p2p_op_t{buf, nullptr, count * dt_size(dt), peer, 1}, st);
}
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t st) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclRecv, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t); return f(buf, count, dt, peer, s->real, st); }
  return queue_p2p(s, b200mpi_p2p_op_t{nullptr, buf, count * dt_size(dt), peer, 0}, st);
}

// Real NCCL's allocations are VMM handles it later retains / registers: when a real communicator can be involved the
// allocation must come from the real library (a cudaMalloc pointer makes its registration path fail with "invalid argument").
static bool real_nccl_in_play() { return forward_all() || g_in_real > 0 || g_forwarded.load() > 0; }
// ncclMemAlloc hands out exportable VMM memory (b200mpi_mem_alloc); ncclCommRegister on such a buffer is COLLECTIVE here
// (every rank registers its allocation of the same size, in the same order - what torch's MemPool registration and NCCL's
// own symmetric registration require as well): the runtime maps the peers' allocations and binds them to an NVLS multicast
// object, and from then on collectives on tensors inside it (>= B200MPI_REG_MIN_BYTES, agreed per call through the shm
// rendezvous) run zero-copy on the window: multimem.ld_reduce / multimem.st straight on the user's memory.
// Buffers that did not come from ncclMemAlloc need no registration: cudaIpc registration is lazy (see reg_exchange).
ncclResult_t ncclMemAlloc(void** ptr, size_t size) {
  if (real_nccl_in_play()) { if (auto f = REAL(ncclMemAlloc, void**, size_t)) return f(ptr, size); }
  if (!ptr) return err(ncclInvalidArgument, "ncclMemAlloc: null pointer");
  if (b200mpi_mem_alloc(size, ptr) == 0) return ncclSuccess;
  return cudaMalloc(ptr, size) == cudaSuccess ? ncclSuccess : err(ncclUnhandledCudaError, "ncclMemAlloc");
}
ncclResult_t ncclMemFree(void* ptr) {
  if (real_nccl_in_play()) { if (auto f = REAL(ncclMemFree, void*)) return f(ptr); }
  if (!ptr) return ncclSuccess;
  void* base = nullptr;
  if (b200mpi_mem_lookup(ptr, &base, nullptr) && base == ptr) return from_rc(b200mpi_mem_free(ptr), "ncclMemFree");
  return cudaFree(ptr) == cudaSuccess ? ncclSuccess : err(ncclUnhandledCudaError, "ncclMemFree");
}
ncclResult_t ncclCommRegister(const ncclComm_t c, void* buff, size_t size, void** handle) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclCommRegister, ncclComm_t, void*, size_t, void**); return f(s->real, buff, size, handle); }
  if (handle) *handle = nullptr;
  void* base = nullptr;
  if (!s->mine || !b200mpi_mem_lookup(buff, &base, nullptr)) return ncclSuccess;   // lazily registered on first large use
  auto it = s->registered.find(base);
  int win = -1;
  if (it == s->registered.end()) {
    if (b200mpi_window_adopt(s->mine, base, &win) != 0) {
      if (debug()) fprintf(stderr, "[b200mpi nccl shim] ncclCommRegister: %s (buffer stays usable through the staged paths)\n", b200mpi_last_error());
      return ncclSuccess;
    }
    s->registered[base] = win;
    g_registered++;
  } else {
    win = it->second;
  }
  if (handle) *handle = reinterpret_cast<void*>((intptr_t)(win + 1));
  return ncclSuccess;
}
ncclResult_t ncclCommDeregister(const ncclComm_t c, void* handle) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclCommDeregister, ncclComm_t, void*); return f(s->real, handle); }
  return ncclSuccess;   // windows are released with the communicator (freeing one is a collective, deregistration is not)
}
ncclResult_t ncclCommWindowRegister(ncclComm_t c, void* buff, size_t size, void** win, int flags) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclCommWindowRegister, ncclComm_t, void*, size_t, void**, int); return f ? f(s->real, buff, size, win, flags) : ncclSuccess; }
  if (win) *win = nullptr;
  return ncclSuccess;
}
ncclResult_t ncclCommWindowDeregister(ncclComm_t c, void* win) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclCommWindowDeregister, ncclComm_t, void*); return f ? f(s->real, win) : ncclSuccess; }
  return ncclSuccess;
}
ncclResult_t ncclDevCommCreate(ncclComm_t c, const void* reqs, void* out) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclDevCommCreate, ncclComm_t, const void*, void*); return f ? f(s->real, reqs, out) : err(ncclInvalidUsage, "ncclDevCommCreate missing in real NCCL"); }
  return err(ncclInvalidUsage, "device-side communicators are not provided by the b200mpi shim");
}
ncclResult_t ncclDevCommDestroy(ncclComm_t c, const void* dev) {
  Shim* s = S(c);
  if (s->real) { auto f = REAL(ncclDevCommDestroy, ncclComm_t, const void*); return f ? f(s->real, dev) : ncclSuccess; }
  return ncclSuccess;
}

}  // extern "C"
