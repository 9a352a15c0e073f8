// Point-to-point send/receive over peer memory (ncclSend / ncclRecv of the NCCL-ABI shim).
//
// One launch executes a whole batch (everything between ncclGroupStart and ncclGroupEnd): one CTA per operation,
// all co-resident, so a rank that both sends to and receives from its peers cannot deadlock on stream order the way
// two back-to-back kernels would. Transport: every rank owns a mailbox window with two 1 MiB slots per source rank;
// a message is cut into chunks, chunk q of the (src -> dst) stream goes to slot q & 1 of dst's mailbox for src:
//
//   sender  : wait ack[slot] >= q-1 (slot consumed, only from the third chunk on) -> push chunk -> ready[slot] = q+1
//   receiver: wait ready[slot] >= q+1 -> copy out of its own mailbox -> ack[slot] (in the sender's window) = q+1
//
// q counts chunks of the (src, dst) stream since the communicator was created; the counters live in device memory and
// are advanced by a one-thread commit kernel after the batch, so a captured CUDA graph replays correctly. Up to two
// chunks are "eager": a send of <= 2 MiB completes without the matching receive having started, larger unmatched
// sends wait (NCCL semantics: pair them inside a group). Flags use release/acquire at .sys scope exactly like the
// collectives' rank barrier; every wait is bounded by the communicator's watchdog.
//
// Status: EXPERIMENTAL (B200MPI_P2P=1 allocates the mailbox window); written after the round's GPU budget was spent,
// not yet run on hardware. The reference has no device code; NCCL provides this under Horovod.
#include "kernels.h"

namespace b200mpi {

// The protocol itself is platform-neutral: P supplies the thread geometry, the CTA barrier, flag wait/release and the two
// copy loops. The device platform below is what the kernel runs; csrc/tests/p2p_protocol_test.cc instantiates the same
// function with host threads and atomics (one thread per operation, ranks = buffers in one process) to check sequence
// numbers, slot reuse, eager sends and multi-batch counter hand-over without a GPU.
template <typename P>
__host__ __device__ inline void p2p_run_op(const P2POp& op, int rank, char* const* box, const uint32_t* cnt, P& pf) {
  const int peer = op.peer;
  if (op.bytes == 0) return;
  const size_t nchunks = (op.bytes + kP2PChunk - 1) / kP2PChunk;
  if (op.is_send) {
    const uint32_t base = cnt[peer] + op.seq_off;                         // chunks already sent to `peer`
    char* slots = box[peer] + (size_t)rank * 2 * kP2PChunk;               // my two slots in the peer's mailbox
    uint32_t* ready = p2p_flag(box[peer], P2P_READY, rank);               // in the peer's window
    uint32_t* ack = p2p_flag(box[rank], P2P_ACK, peer);                   // in my window, written by the peer
    for (size_t k = 0; k < nchunks; k++) {
      const uint32_t q = base + (uint32_t)k;
      const int slot = (int)(q & 1u);
      // the slot still holds chunk q-2 until the receiver acknowledges it with value q-1
      if (q >= 2 && !pf.wait_ge(ack + slot * kP2PFlagStride, q - 1, peer)) return;
      const size_t off = k * kP2PChunk;
      const size_t n = op.bytes - off < kP2PChunk ? op.bytes - off : kP2PChunk;
      pf.push(slots + (size_t)slot * kP2PChunk, op.user + off, n);
      pf.sync();  // every thread's pushes are ordered before the release (cumulativity through bar.sync)
      pf.release(ready + slot * kP2PFlagStride, q + 1);
    }
  } else {
    const uint32_t base = cnt[kMaxRanks + peer] + op.seq_off;             // chunks already received from `peer`
    const char* slots = box[rank] + (size_t)peer * 2 * kP2PChunk;         // the peer's two slots in my mailbox
    uint32_t* ready = p2p_flag(box[rank], P2P_READY, peer);
    uint32_t* ack = p2p_flag(box[peer], P2P_ACK, rank);                   // in the sender's window
    char* user = const_cast<char*>(op.user);
    for (size_t k = 0; k < nchunks; k++) {
      const uint32_t q = base + (uint32_t)k;
      const int slot = (int)(q & 1u);
      if (!pf.wait_ge(ready + slot * kP2PFlagStride, q + 1, peer)) return;
      const size_t off = k * kP2PChunk;
      const size_t n = op.bytes - off < kP2PChunk ? op.bytes - off : kP2PChunk;
      pf.pull(user + off, slots + (size_t)slot * kP2PChunk, n);
      pf.sync();  // all reads of the slot are done before it is handed back
      pf.release(ack + slot * kP2PFlagStride, q + 1);
    }
  }
}

#ifdef __CUDACC__
struct P2PDevice {
  const DevComm& c;
  int* s_ok;
  // thread 0 spins (acquire, .sys scope, bounded by the communicator watchdog); the CTA learns the outcome through smem
  __device__ bool wait_ge(const uint32_t* flag, uint32_t want, int peer) {
    if (threadIdx.x == 0) {
      int ok = 1;
      unsigned long long t0 = 0;
      uint32_t spins = 0;
      while ((int32_t)(ld_acquire_sys(flag) - want) < 0) {
        if ((++spins & 0x3ffu) == 0) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > c.timeout_ns) { *c.err = 1 + peer; ok = 0; break; }
        }
      }
      *s_ok = ok;
    }
    __syncthreads();
    const bool ok = *s_ok != 0;
    __syncthreads();  // s_ok may be rewritten by the next wait
    return ok;
  }
  __device__ void sync() { __syncthreads(); }
  __device__ void release(uint32_t* flag, uint32_t v) { if (threadIdx.x == 0) st_release_sys(flag, v); }
  __device__ void push(char* dst, const char* src, size_t n) {      // local user buffer -> peer mailbox
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const size_t nvec = (n + 15) / 16;
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) st_peer_v4(dst + i * 16, user_load(src, i, n, aligned));
  }
  __device__ void pull(char* dst, const char* src, size_t n) {      // own mailbox -> local user buffer
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    const size_t nvec = (n + 15) / 16;
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) user_store(dst, i, n, aligned, ld_sys_v4(src + i * 16));
  }
};

__global__ void __launch_bounds__(kThreads)
k_p2p_batch(const __grid_constant__ P2PArgs a) {
  __shared__ int s_ok;
  P2PDevice pf{a.c, &s_ok};
  p2p_run_op(a.ops[blockIdx.x], a.c.rank, a.box.p, a.cnt, pf);
}

// Advances the per-peer chunk counters by what the batch moved (stream-ordered after k_p2p_batch).
__global__ void k_p2p_commit(uint32_t* cnt, P2PCommit add) {
  const int i = threadIdx.x;
  if (i < 2 * kMaxRanks && add.n[i]) cnt[i] += add.n[i];
}

cudaError_t launch_p2p_batch(cudaStream_t s, const P2PArgs& a, const P2PCommit& add) {
  k_p2p_batch<<<a.nops, kThreads, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_p2p_commit<<<1, 32, 0, s>>>(a.cnt, add);
  return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace b200mpi
