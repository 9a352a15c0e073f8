/*
 * hvdcore — native coordination engine behind the Horovod-compatible front-end (mpi_operator_b200.hvd).
 *
 * What the reference's workloads get from Horovod's C++ core (examples/v2beta1/horovod/tensorflow_mnist.py:133
 * hvd.DistributedOptimizer, examples/v2beta1/tensorflow-benchmarks/tensorflow-benchmarks.yaml:42
 * --variable_update=horovod; SURVEY.md §2.2 "Horovod core", §3.3 "negotiate -> fuse -> allreduce -> unfuse"):
 * a background thread per rank that (1) negotiates, by tensor NAME, which collectives every rank has submitted — ranks
 * may submit in different orders —, (2) fuses ready allreduces of one dtype into a fusion buffer, (3) runs them and
 * completes per-tensor handles, (4) remembers negotiated tensors in a response cache so steady-state steps exchange
 * 4-byte ids, (5) writes a Chrome-trace timeline, (6) reports stalled ranks, (7) implements join() for uneven data.
 *
 * Single-box redesign: there is no coordinator rank and no MPI. Every cycle each rank publishes its new requests in
 * the job's shared-memory rendezvous segment (csrc/runtime/rendezvous.h), every rank reads all of them and runs the
 * SAME deterministic state machine, so all ranks derive the identical fused response list without a second message.
 * Host tensors are reduced through a bulk-data segment (Rendezvous::open_boxes; slice-parallel fold), device tensors by the
 * b200mpi kernels on the engine's own stream and communicator.
 *
 * All functions return 0 or a negative code unless stated; hvdcore_last_error() is thread-local.
 */
#ifndef HVDCORE_H_
#define HVDCORE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HVD_ALLREDUCE = 0,
  HVD_ALLGATHER = 1, /* allgatherv: per-rank byte counts supplied by the caller (see hvdcore_enqueue) */
  HVD_BROADCAST = 2,
  HVD_ALLTOALL = 3,  /* alltoallv: send byte counts per peer supplied by the caller */
  HVD_BARRIER = 4,
  HVD_JOIN = 5,
  HVD_EXCHANGE = 6,  /* <= 1024 bytes per rank carried inside the negotiation itself; result = world blobs */
} hvd_op_t;

typedef enum {
  HVD_U8 = 0, HVD_I8 = 1, HVD_I16 = 2, HVD_I32 = 3, HVD_I64 = 4, HVD_F16 = 5, HVD_BF16 = 6, HVD_F32 = 7, HVD_F64 = 8,
  HVD_BOOL = 9,
} hvd_dtype_t;

typedef enum { HVD_SUM = 0, HVD_MIN = 1, HVD_MAX = 2, HVD_PROD = 3, HVD_ADASUM = 4 /* host float tensors; never fused */ } hvd_redop_t;

enum {
  HVD_OK = 0,
  HVD_ERR_INVALID = -1,        /* bad argument */
  HVD_ERR_NOT_INIT = -2,
  HVD_ERR_SHUTDOWN = -3,       /* "Horovod has been shut down" */
  HVD_ERR_MISMATCH = -4,       /* ranks disagree about a tensor (dtype, size, op, root ...) */
  HVD_ERR_TRANSPORT = -5,      /* rendezvous failure: a peer died or timed out */
  HVD_ERR_DUPLICATE = -6,      /* name already in flight on this rank */
  HVD_ERR_UNSUPPORTED = -7,
  HVD_ERR_STALL = -8,          /* stall shutdown time exceeded */
};

/* GPU executor: entry points of libb200mpi.so handed over as plain addresses (the engine does not link it). */
typedef struct {
  void* comm;                 /* b200mpi_comm_t dedicated to the engine (collectives on a communicator are ordered) */
  int device;
  void* allreduce;            /* int (*)(comm, in, out, count, dtype, op, scale, algo, stream) */
  void* broadcast_bytes;      /* int (*)(comm, buf, bytes, root, stream) */
  void* barrier;              /* int (*)(comm, stream) */
  void* last_error;           /* const char* (*)(void) */
} hvdcore_gpu_t;

/* Attaches to the rendezvous segment "<job_id>-hvd" and starts the background thread. `gpu` may be NULL (host only).
 * Tunables are read from the environment like Horovod's: HOROVOD_CYCLE_TIME (ms, default 1), HOROVOD_FUSION_THRESHOLD
 * (bytes, default 64 MiB), HOROVOD_CACHE_CAPACITY (default 1024, 0 disables), HOROVOD_TIMELINE (path, rank 0 writes),
 * HOROVOD_TIMELINE_MARK_CYCLES, HOROVOD_STALL_CHECK_DISABLE, HOROVOD_STALL_CHECK_TIME_SECONDS (60), HOROVOD_STALL_SHUTDOWN_TIME_SECONDS (0 = never). */
int hvdcore_init(const char* job_id, int rank, int world, const hvdcore_gpu_t* gpu);
/* Tells every rank's engine to stop after the current cycle; outstanding handles fail with HVD_ERR_SHUTDOWN. */
int hvdcore_shutdown(void);
int hvdcore_initialized(void);
int hvdcore_rank(void);
int hvdcore_size(void);

/*
 * Submit one collective; returns a handle (> 0) or a negative error. The buffers must stay valid until the handle
 * completes. `count` is in elements of `dtype`.
 *   ALLREDUCE : in -> out (may alias), out = postscale * reduce(prescale * in)
 *   ALLGATHER : in (count elements) -> out; `extra` = world int64 byte counts (every rank passes the same vector,
 *               e.g. obtained with HVD_EXCHANGE), out holds their sum
 *   BROADCAST : in place on `out` from `root`
 *   ALLTOALL  : `extra` = world int64 send byte counts followed by world int64 receive byte counts
 *   EXCHANGE  : in = blob of `count` bytes (dtype U8, <= 1024), out = world * count bytes
 *   BARRIER, JOIN : no buffers; hvdcore_wait on a JOIN handle returns the last rank that joined (>= 0)
 * `device` < 0: host memory. >= 0: device memory; `ready_event` (cudaEvent_t or NULL) is waited for on the engine's
 * stream before the data is touched.
 */
int hvdcore_enqueue(hvd_op_t op, const char* name, const void* in, void* out, int64_t count, hvd_dtype_t dtype,
                    hvd_redop_t redop, int root, double prescale, double postscale, int device, void* ready_event,
                    const int64_t* extra, int n_extra);
/* 1 = finished (successfully or not), 0 = still running, < 0 = unknown handle */
int hvdcore_poll(int handle);
/* Blocks until the handle finishes, releases it, returns its status (JOIN: last joined rank). */
int hvdcore_wait(int handle);
const char* hvdcore_last_error(void);

/* Timeline control at run time (hvd.start_timeline / hvd.stop_timeline); only rank 0 writes. */
int hvdcore_start_timeline(const char* path);
int hvdcore_stop_timeline(void);
/* snprintf-style JSON: cycles, tensors, fused groups, bytes, cache hits/misses, stalls. */
int hvdcore_stats_json(char* buf, size_t cap);
/* Override a tunable at run time: "cycle_time_ms", "fusion_threshold", "stall_check_s", "stall_shutdown_s". */
int hvdcore_set_param(const char* key, double value);

#ifdef __cplusplus
}
#endif
#endif  /* HVDCORE_H_ */
