// Hardware capability probe for the b200mpi runtime (run under gpurun --gpus 2).
// Checks: device attrs, P2P, VMM posix-fd export, multicast (NVLS) create/bind/map,
// multimem.ld_reduce / multimem.st correctness, cudaIpc across fork().
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>
#include <sys/mman.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while(0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e,&s); printf("CU error %s (%d) at %s:%d\n", s?s:"?", (int)e, __FILE__, __LINE__); return 1;} } while(0)

__global__ void fill(float* p, float v, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void mc_allreduce(float* mc, float* out, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc + 4 * i) : "memory");
    reinterpret_cast<float4*>(out)[i] = v;
  }
}
__global__ void mc_store(float* mc, float v, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n4) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc + 4 * i), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
  }
}

static int probe_multicast(int ndev) {
  CU(cuInit(0));
  int mc = 0;
  for (int d = 0; d < ndev; d++) {
    int v = 0, fd = 0, fab = 0;
    CU(cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d));
    CU(cuDeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d));
    cuDeviceGetAttribute(&fab, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, d);
    printf("dev %d: multicast_supported=%d posix_fd_handle=%d fabric_handle=%d\n", d, v, fd, fab);
    mc += v;
  }
  if (mc < ndev || ndev < 2) { printf("MULTICAST: not available on all devices (or <2 devs)\n"); return 0; }
  CUmulticastObjectProp prop; memset(&prop, 0, sizeof(prop));
  prop.numDevices = ndev; prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR; prop.flags = 0;
  size_t gran = 0, size = 64u << 20;
  prop.size = size;
  CU(cuMulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  printf("multicast granularity recommended=%zu\n", gran);
  size = (size + gran - 1) / gran * gran; prop.size = size;
  CUmemGenericAllocationHandle mch;
  CUresult r = cuMulticastCreate(&mch, &prop);
  if (r != CUDA_SUCCESS) { const char* s; cuGetErrorString(r,&s); printf("MULTICAST: cuMulticastCreate failed: %s\n", s); return 0; }
  int fd = -1;
  r = cuMemExportToShareableHandle(&fd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  printf("multicast export fd: res=%d fd=%d\n", (int)r, fd);
  std::vector<CUmemGenericAllocationHandle> mem(ndev);
  std::vector<CUdeviceptr> uc(ndev), mcva(ndev);
  for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); CK(cudaFree(0)); CU(cuMulticastAddDevice(mch, d)); }
  for (int d = 0; d < ndev; d++) {
    CK(cudaSetDevice(d));
    CUmemAllocationProp ap; memset(&ap, 0, sizeof(ap));
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED; ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ap.location.id = d;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g2 = 0; CU(cuMemGetAllocationGranularity(&g2, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (d == 0) printf("mem granularity recommended=%zu\n", g2);
    CU(cuMemCreate(&mem[d], size, &ap, 0));
    CU(cuMulticastBindMem(mch, 0, mem[d], 0, size, 0));
    CUmemAccessDesc ad[8]; for (int k = 0; k < ndev; k++) { ad[k].location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad[k].location.id = k; ad[k].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; }
    CU(cuMemAddressReserve(&uc[d], size, gran, 0, 0));
    CU(cuMemMap(uc[d], size, 0, mem[d], 0));
    CU(cuMemSetAccess(uc[d], size, ad, ndev));
    CU(cuMemAddressReserve(&mcva[d], size, gran, 0, 0));
    CU(cuMemMap(mcva[d], size, 0, mch, 0));
    CU(cuMemSetAccess(mcva[d], size, &ad[d], 1));
  }
  size_t n = size / 4;
  for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); fill<<<(n + 255) / 256, 256>>>((float*)uc[d], (float)(d + 1), n); CK(cudaDeviceSynchronize()); }
  float expect = ndev * (ndev + 1) / 2.0f;
  for (int d = 0; d < ndev; d++) {
    CK(cudaSetDevice(d));
    float* out; CK(cudaMalloc(&out, size));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mc_allreduce<<<(n / 4 + 255) / 256, 256>>>((float*)mcva[d], out, n / 4);
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 5; it++) mc_allreduce<<<(n / 4 + 255) / 256, 256>>>((float*)mcva[d], out, n / 4);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    float h[4]; CK(cudaMemcpy(h, out + n - 4, 16, cudaMemcpyDeviceToHost));
    printf("dev %d multimem.ld_reduce: got %.1f expect %.1f  (%.3f ms for %zu MiB -> %.1f GB/s out)\n", d, h[0], expect, ms, size >> 20, size / ms / 1e6);
    CK(cudaFree(out));
  }
  CK(cudaSetDevice(0));
  mc_store<<<(n / 4 + 255) / 256, 256>>>((float*)mcva[0], 42.f, n / 4); CK(cudaDeviceSynchronize());
  for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); float h; CK(cudaMemcpy(&h, (float*)uc[d] + 12345, 4, cudaMemcpyDeviceToHost)); printf("dev %d after multimem.st: %.1f (expect 42)\n", d, h); }
  printf("MULTICAST: OK\n");
  return 0;
}

__global__ void peer_write(float* remote, float v, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) remote[i] = v; }

static int probe_ipc(int ndev) {
  // fork before any CUDA use in the child path: parent has used CUDA, so use a fresh exec-less
  // fork from a process that has NOT initialised CUDA (this function is called first in main()).
  struct Sh { cudaIpcMemHandle_t h[2]; volatile int ready[2]; volatile int done[2]; };
  Sh* sh = (Sh*)mmap(nullptr, sizeof(Sh), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset(sh, 0, sizeof(Sh));
  int nproc = 2;
  pid_t pids[2];
  for (int r = 0; r < nproc; r++) {
    pids[r] = fork();
    if (pids[r] == 0) {
      int dev = ndev >= 2 ? r : 0;
      CK(cudaSetDevice(dev));
      float* buf; size_t n = 1 << 20; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
      CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)&sh->h[r], buf));
      __sync_synchronize(); sh->ready[r] = 1;
      while (!sh->ready[1 - r]) usleep(100);
      float* peer = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle((void**)&peer, sh->h[1 - r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { printf("rank %d: cudaIpcOpenMemHandle failed: %s\n", r, cudaGetErrorString(e)); sh->done[r] = 1; _exit(3); }
      peer_write<<<(n + 255) / 256, 256>>>(peer, 100.f + r, n); CK(cudaDeviceSynchronize());
      __sync_synchronize(); sh->done[r] = 1;
      while (!sh->done[1 - r]) usleep(100);
      float h; CK(cudaMemcpy(&h, buf + 777, 4, cudaMemcpyDeviceToHost));
      printf("IPC rank %d (dev %d): my buffer now holds %.1f (expect %.1f)\n", r, dev, h, 100.f + (1 - r));
      _exit(h == 100.f + (1 - r) ? 0 : 4);
    }
  }
  int bad = 0;
  for (int r = 0; r < nproc; r++) { int st; waitpid(pids[r], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st)) bad++; }
  printf("IPC: %s\n", bad ? "FAILED" : "OK");
  return bad;
}

int main() {
  // device count without creating a context in the parent: cuInit+cuDeviceGetCount is context-free
  cuInit(0); int ndev = 0; cuDeviceGetCount(&ndev);
  printf("ndev=%d\n", ndev);
  probe_ipc(ndev);
  for (int d = 0; d < ndev; d++) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, d));
    printf("dev %d: %s cc %d.%d SMs %d mem %.1f GB L2 %d MB\n", d, p.name, p.major, p.minor, p.multiProcessorCount, p.totalGlobalMem / 1e9, p.l2CacheSize >> 20);
    for (int q = 0; q < ndev; q++) if (q != d) { int a = 0; CK(cudaDeviceCanAccessPeer(&a, d, q)); int na = 0; cudaDeviceGetP2PAttribute(&na, cudaDevP2PAttrNativeAtomicSupported, d, q); printf("  p2p %d->%d access=%d native_atomics=%d\n", d, q, a, na); }
  }
  return probe_multicast(ndev);
}
