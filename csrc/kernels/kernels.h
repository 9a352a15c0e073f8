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
};

struct Launch {
  cudaStream_t stream;
  int blocks;             // gridDim.x
  int emu_world;          // 0: real multi-process rank; >0: gridDim.y virtual ranks
  const KArgs* emu_args;  // device array of emu_world KArgs (emulated mode)
};

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
enum : int { MODE_P2P = 0, MODE_NVLS = 1 };

// a = args of this rank (ignored in emulated mode where l.emu_args is used)
cudaError_t launch_allreduce_twoshot(const Launch& l, const KArgs& a, int dtype, int mode, bool staged);
cudaError_t launch_allreduce_oneshot(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_allreduce_sgd(const Launch& l, const KArgs& a, int dtype, int mode);
cudaError_t launch_allgather(const Launch& l, const KArgs& a);
cudaError_t launch_broadcast(const Launch& l, const KArgs& a, int mode);
cudaError_t launch_reduce_scatter(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_reduce(const Launch& l, const KArgs& a, int dtype);
cudaError_t launch_alltoall(const Launch& l, const KArgs& a);
cudaError_t launch_barrier(const Launch& l, const KArgs& a);
cudaError_t launch_scale_cast(cudaStream_t s, const void* in, int in_dt, void* out, int out_dt,
                              size_t count, float scale);
cudaError_t launch_fill_u32(cudaStream_t s, uint32_t* p, uint32_t v, size_t n);

}  // namespace b200mpi
