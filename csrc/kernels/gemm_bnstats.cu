// 1x1-convolution GEMM with BatchNorm statistics fused into the epilogue (sm_100a: TMA + tcgen05 + TMEM).
//
//   Y[M,N] = X[M,K] . W[N,K]^T      bf16 in, fp32 accumulate in TMEM, bf16 out   (NHWC activations: M = N*H*W, K = Cin,
//                                                                                  N = Cout, W is the conv weight as stored)
//   partials[p][2n+0] += sum_m Y[m,n],  partials[p][2n+1] += sum_m Y[m,n]^2      (one row per CTA m-group, the layout the
//                                                                                  BN finalize kernel of bn_act.cu merges)
//
// Why: two thirds of ResNet's BatchNorm layers follow a 1x1 convolution. Their statistics pass re-reads the whole
// convolution output from HBM (profiles/launches_resnet101_step_fusedbn.md: k_bn_fwd_stats = 6.7 % of the step); here the
// tile is still in shared memory when its column sums are taken, so the pass disappears. The reference has no kernels at
// all — this is part of the data plane the new framework adds.
//
// Structure (the canonical Blackwell GEMM: one warp per role, three pipelines):
//   warp 0      TMA producer   : cp.async.bulk.tensor loads of the X and W tiles (128B swizzle) into a 4-stage ring
//   warp 1      MMA issuer     : one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) into one of two TMEM
//                                accumulators; tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..5  epilogue       : tcgen05.ld (32 lanes x 32 columns per warp) -> bf16 -> swizzled smem staging ->
//                                column sums from smem + TMA store of the tile; overlaps the next tile's main loop
// Persistent: CTA c owns column block c % num_n and walks m-blocks with stride grid/num_n, so every epilogue thread keeps one
// column's running sums in registers for the whole kernel and writes a single partial row at the end (no atomics).
//
// Status: first executed on a B200 in round 2 - numerics 10 / 10 against fp32 references (tests/test_zz_gemm_bnstats_gpu.py),
// ncu capture + 12-shape micro-benchmark in profiles/ncu_gemm_bnstats.md, ResNet-101 step 14.15 -> 13.75 ms; on by default
// for eligible 1x1 convolutions (B200MPI_FUSED_CONV1X1=0 restores cuDNN). Every wait is bounded and traps instead of hanging.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../include/b200mpi.h"
#include "gemm_bnstats_logic.h"

namespace b200mpi {
namespace gemm {

constexpr int kThreads = 192;    // 6 warps: TMA, MMA, 4 x epilogue
constexpr int kEpiThreads = 128;
constexpr uint32_t kABytes = BM * BK * 2;   // 16 KiB
constexpr long long kTimeoutCycles = 6000000000LL;  // ~3 s: a broken pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---------------------------------------------------------------- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kTimeoutCycles) asm volatile("trap;");
  }
}

// --------------------------------------------------------------------- TMA ----
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ----------------------------------------------------------------- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp: SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (=1: unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (=64: 8 rows x 128 B between core-matrix groups) | [46,48) version = 1 (sm_100) |
//   [61,64) layout = 2 (SWIZZLE_128B)
__host__ __device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at [7,10)/[10,13),
// a/b K-major (0) at 15/16, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(uint32_t lo_f32_bits, uint32_t hi_f32_bits) {
  __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(lo_f32_bits), __uint_as_float(hi_f32_bits));
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN>
struct Smem {
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kOutBoxes = BN / 64;                  // TMA store boxes of [128 rows x 64 cols] (128 B inner)
  static constexpr uint32_t kOutBytes = kOutBoxes * kBoxBytes;
  static constexpr uint32_t kA = 0;
  static constexpr uint32_t kB = kStages * kABytes;
  static constexpr uint32_t kOut = kB + kStages * kBBytes;
  static constexpr uint32_t kBars = kOut + kOutBytes;             // full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr uint32_t kTmemSlot = kBars + (2 * kStages + 4) * 8;
  static constexpr uint32_t kTotal = kTmemSlot + 16;
  static constexpr uint32_t kDynamic = kTotal + 1024;             // slack to align the base to 1024 B (swizzle atoms)
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_bnstats(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY,
               float* __restrict__ partials, int M, int N, int K, int num_n) {
  using L = Smem<BN>;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulators; 128 or 256: a power of two >= 32
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base + L::kA, sB = base + L::kB, sOut = base + L::kOut, sBar = base + L::kBars;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return sBar + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return sBar + 8u * (2 * kStages + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::kTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TileWalk walk((int)blockIdx.x, (int)gridDim.x, num_n, M);
  const int m_first = walk.m_first, m_step = walk.m_step, num_m = walk.num_m;
  const int num_k = K / BK;
  const int n0 = walk.n_blk * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1) {  // TMEM allocation: one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::kTmemSlot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2 && lane == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(full_bar(s), 1);    // producer's arrive.expect_tx (+ the TMA transaction bytes)
      mbar_init(empty_bar(s), 1);   // tcgen05.commit of the MMAs that read the stage
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);   // tcgen05.commit after the tile's last MMA
      mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer ====
    if (lane == 0) {
      Ring st;
      for (int m_blk = m_first; m_blk < num_m; m_blk += m_step) {
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(empty_bar(st.s), st.ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(st.s), L::kStageBytes);
          tma_load_2d(sA + st.s * kABytes, &tmX, full_bar(st.s), kb * BK, m_blk * BM);
          tma_load_2d(sB + st.s * L::kBBytes, &tmW, full_bar(st.s), kb * BK, n0);
          st.advance(kStages);
        }
      }
    }
    __syncwarp();   // lanes 1..31 skipped the loop: reconverge before the block-wide barrier below
  } else if (warp == 1) {
    // ======================================================= MMA issuer ====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      Ring st, acc;
      for (int m_blk = m_first; m_blk < num_m; m_blk += m_step) {
        mbar_wait(tempty_bar(acc.s), acc.ph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc.s * BN);
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(full_bar(st.s), st.ph);
          tc_fence_after();
          const uint64_t adesc = make_desc_kmajor_sw128(sA + st.s * kABytes);
          const uint64_t bdesc = make_desc_kmajor_sw128(sB + st.s * L::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++)  // +32 bytes along K inside the swizzle atom = +2 in the address field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(st.s));        // stage reusable once these MMAs have read it
          st.advance(kStages);
        }
        umma_commit(tfull_bar(acc.s));         // accumulator complete
        acc.advance(2);
      }
    }
    __syncwarp();
  } else {
    // ========================================================= epilogue ====
    const int et = threadIdx.x - 64;             // 0..127
    const int q = warp & 3;                      // TMEM lane quarter this warp may access: lanes [32q, 32q+32)
    const int row = q * 32 + lane;               // tile row held by this thread
    const bool col_owner = et < BN;              // thread `et` owns column n0 + et for the statistics
    float s1 = 0.f, s2 = 0.f;
    Ring acc;
    for (int m_blk = m_first; m_blk < num_m; m_blk += m_step) {
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous tile's store has read the staging
      epi_bar_sync();
      mbar_wait(tfull_bar(acc.s), acc.ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc.s * BN);
#pragma unroll
      for (int c = 0; c < BN / 32; c++) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        // 32 fp32 -> 32 bf16 = 4 groups of 16 B in the swizzled staging layout (gemm_bnstats_logic.h)
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const uint32_t dst = sOut + stage_group_byte(row, c, g);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[8 * g + 0], v[8 * g + 1])),
                       "r"(pack_bf16(v[8 * g + 2], v[8 * g + 3])), "r"(pack_bf16(v[8 * g + 4], v[8 * g + 5])),
                       "r"(pack_bf16(v[8 * g + 6], v[8 * g + 7]))
                       : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc.s));  // accumulator free for the MMA warp (tile after next)
      acc.advance(2);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
      epi_bar_sync();
      if (et == 0) {
#pragma unroll
        for (int b = 0; b < (int)L::kOutBoxes; b++) tma_store_2d(&tmY, sOut + (uint32_t)b * kBoxBytes, n0 + b * 64, m_blk * BM);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (col_owner) {
        // column sums of the bf16-rounded tile (what the BN apply pass will read back). Rows past M are zero-filled by TMA.
#pragma unroll 8
        for (int r = 0; r < BM; r++) {
          uint16_t h;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(sOut + stage_elem_byte(et, r)));
          const float f = __uint_as_float((uint32_t)h << 16);
          s1 += f;
          s2 = fmaf(f, f, s2);
        }
      }
    }
    if (col_owner && m_first < num_m) {
      float* row_out = partials + partial_index(m_first, N, n0 + et);
      row_out[0] = s1;
      row_out[1] = s2;
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores landed before the CTA exits
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, cols] bf16 matrix, box = [box_rows, 64 cols] with 128-byte swizzle
static bool make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};              // innermost first
  const cuuint64_t strides[1] = {cols * 2};             // bytes, dims 1..rank-1
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN>
static int launch(const void* X, const void* W, void* Y, float* partials, int* parts_out, long long M, int N, int K, cudaStream_t s) {
  CUtensorMap tmX, tmW, tmY;
  if (!make_map(&tmX, X, (uint64_t)M, (uint64_t)K, BM) || !make_map(&tmW, W, (uint64_t)N, (uint64_t)K, BN) ||
      !make_map(&tmY, Y, (uint64_t)M, (uint64_t)N, BM))
    return B200MPI_ERR_CUDA;
  const int num_n = N / BN, num_m = (int)((M + BM - 1) / BM);
  int groups = sm_count() / num_n;
  if (groups < 1) return B200MPI_ERR_UNSUPPORTED;
  if (groups > num_m) groups = num_m;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_gemm_bnstats<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<BN>::kDynamic) != cudaSuccess)
      return B200MPI_ERR_CUDA;
    attr_set = true;
  }
  k_gemm_bnstats<BN><<<groups * num_n, kThreads, Smem<BN>::kDynamic, s>>>(tmX, tmW, tmY, partials, (int)M, N, K, num_n);
  if (parts_out) *parts_out = groups;
  return cudaGetLastError() == cudaSuccess ? 0 : B200MPI_ERR_CUDA;
}

}  // namespace gemm
}  // namespace b200mpi

extern "C" {

// M >= 128: every shape that ran on hardware had at least one full row tile (ragged tails - M = 196, 1000 - are covered); smaller
// problems (a 2x2 feature map at batch 4) are latency-bound anyway and stay with the library.
int b200mpi_gemm_bnstats_supported(long long M, int N, int K) {
  return (M >= 128 && M < (1LL << 31) - 128 && N >= 64 && N % 64 == 0 && K >= 64 && K % 64 == 0 &&
          N / (N % 128 == 0 ? 128 : 64) <= 148) ? 1 : 0;
}

// rows of per-column {sum, sum of squares} the kernel may write: one per m-group, at most one per SM
size_t b200mpi_gemm_bnstats_partial_floats(int N) { return (size_t)148 * 2 * (size_t)N; }

// Y = X . W^T (bf16, row-major, X [M,K], W [N,K], Y [M,N], all 16-byte aligned) and partials[parts][2N]; *parts = rows written.
int b200mpi_gemm_bnstats(const void* X, const void* W, void* Y, float* partials, int* parts, long long M, int N, int K, void* stream) {
  using namespace b200mpi::gemm;
  if (!b200mpi_gemm_bnstats_supported(M, N, K)) return B200MPI_ERR_UNSUPPORTED;
  if (((uintptr_t)X | (uintptr_t)W | (uintptr_t)Y) & 15) return B200MPI_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  return N % 128 == 0 ? launch<128>(X, W, Y, partials, parts, M, N, K, s) : launch<64>(X, W, Y, partials, parts, M, N, K, s);
}

}  // extern "C"
