// Host-visible launch interface of the b200mpi collective kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "device.cuh"

namespace b200mpi {

struct KArgs {
  DevComm c;
  Win buf;          // staging region or symmetric user region (per-rank bases)
  const char* in;   // user input  (nullptr: data already in buf)
  char* out;        // user output (nullptr: result stays in buf)
  size_t nbytes;    // payload bytes of one logical unit (op specific)
  size_t nvec;      // 16-byte vectors in the logical buffer
  size_t per;       // vectors per rank slice / slot
  size_t ustride;   // byte stride between per-rank blocks in the user buffer
  float scale;
  int op;
  int root;
  int in_aligned;
  int out_aligned;
  // fused allreduce+SGD
  Win param;        // fp32 parameter window region
  Win lowp;         // optional bf16 parameter shadow (p[0]==nullptr: absent)
  float* mom;       // this rank's momentum slice
  float lr, mu, wd;
  const float* hyper;  // optional device {lr, mu, wd}: lets a captured CUDA graph follow LR schedules
  int nesterov;
  int first_step;
  // pipelined staged allreduce: lanes x {copy-in, reduce, copy-out} CTAs, `depth` staging slots of `per` vectors per lane
  int lanes;
  int depth;
  unsigned long long* dbg;   // optional timeline: [role 3][lane kPipeLanes][chunk kPipeDbgChunks][3] globaltimer ns
};
constexpr int kPipeDbgChunks = 32;

struct Launch {
  cudaStream_t stream;
  int blocks;             // gridDim.x
  int emu_world;          // 0: real multi-process rank; >0: gridDim.y virtual ranks
  const KArgs* emu_args;  // device array of emu_world KArgs (emulated mode)
};

// ---- point-to-point (p2p.cu) -------------------------------------------------------------------------------------
constexpr size_t kP2PChunk = (size_t)1 << 20;                  // mailbox slot size
constexpr size_t kP2PDataBytes = (size_t)kMaxRanks * 2 * kP2PChunk;   // [src rank][2 slots]
constexpr int kP2PFlagStride = 32;                             // u32 words between flags: one 128-byte line each
constexpr size_t kP2PFlagBytes = (size_t)2 * kMaxRanks * 2 * kP2PFlagStride * sizeof(uint32_t);  // READY + ACK
constexpr size_t kP2PWindowBytes = kP2PDataBytes + kP2PFlagBytes;
constexpr int kMaxP2POps = 64;
enum : int { P2P_READY = 0, P2P_ACK = 1 };
// flag of slot 0 for stream (kind, r) inside the mailbox window whose base is `win`; slot 1 is kP2PFlagStride words on
__host__ __device__ inline uint32_t* p2p_flag(char* win, int kind, int r) {
  return reinterpret_cast<uint32_t*>(win + kP2PDataBytes) + ((size_t)kind * kMaxRanks + r) * 2 * kP2PFlagStride;
}
struct P2POp {
  const char* user;   // send: source buffer; receive: destination buffer
  size_t bytes;
  int peer;
  int is_send;
  uint32_t seq_off;   // chunks of earlier operations of this batch on the same (direction, peer) stream
};
struct P2PArgs {
  DevComm c;
  Win box;            // mailbox window of every rank
  uint32_t* cnt;      // device counters: [0,kMaxRanks) chunks sent to peer, [kMaxRanks, 2*kMaxRanks) chunks received from peer
  int nops;
  P2POp ops[kMaxP2POps];
};
struct P2PCommit { uint32_t n[2 * kMaxRanks]; };
cudaError_t launch_p2p_batch(cudaStream_t s, const P2PArgs& a, const P2PCommit& add);

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
enum : int { MODE_P2P = 0, MODE_NVLS = 1 };

// a = args of this rank (ignored in emulated mode where l.emu_args is used)
cudaError_t launch_allreduce_twoshot(const Launch& l, const KArgs& a, int dtype, int mode, bool staged);
enum : int { PIPE_ALLREDUCE = 0, PIPE_ALLGATHER = 1, PIPE_REDUCE_SCATTER = 2, PIPE_BROADCAST = 3 };
cudaError_t launch_pipe(const Launch& l, const KArgs& a, int kind, int dtype, int mode);
cudaError_t launch_allreduce_oneshot(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_allreduce_sgd(const Launch& l, const KArgs& a, int dtype, int mode);
cudaError_t launch_allgather(const Launch& l, const KArgs& a);
cudaError_t launch_allgather_sym(const Launch& l, const KArgs& a, int mode);
cudaError_t launch_broadcast_sym(const Launch& l, const KArgs& a, int mode);
cudaError_t launch_reduce_scatter_sym(const Launch& l, const KArgs& a, int dtype, int mode);
cudaError_t launch_broadcast(const Launch& l, const KArgs& a, int mode);
cudaError_t launch_reduce_scatter(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_reduce(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_alltoall(const Launch& l, const KArgs& a);
cudaError_t launch_barrier(const Launch& l, const KArgs& a);
// Adasum (adasum.cu). Staging layout behind a.buf: [dot board kAdaDotBytes][A: world*per vectors of T][W: world*per vectors as fp32]
constexpr int kAdaMaxBlocks = 64;                                   // CTAs per rank (all co-resident: spin barriers)
constexpr size_t kAdaDotBytes = (size_t)96 << 10;                   // >= (kMaxRanks-1) pairs x kMaxRanks x kAdaMaxBlocks x 3 doubles
static_assert((size_t)(kMaxRanks - 1) * kMaxRanks * kAdaMaxBlocks * 3 * sizeof(double) <= kAdaDotBytes, "dot board too small");
cudaError_t launch_adasum(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_scale_cast(cudaStream_t s, const void* in, int in_dt, void* out, int out_dt,
                              size_t count, float scale);
cudaError_t launch_fill_u32(cudaStream_t s, uint32_t* p, uint32_t v, size_t n);

}  // namespace b200mpi
