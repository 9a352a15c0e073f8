#!/bin/bash
# First 1-GPU call of the next round: validate everything that was written after the GPU budget ran out.
# Usage: gpurun --timeout 900 -- 'tools/gpu_session_r2_1gpu.sh'
export B200MPI_NO_AUTOBUILD=1
mkdir -p gpurun_out
echo "=== 1. GPU tier as the driver runs it ==="
timeout 600 python -m pytest tests -x -q -m gpu --timeout=300 2>&1 | tail -5
echo "=== 2. bf16 parameter shadow: numerics (xfail marker removed by --runxfail) + bench ==="
timeout 300 python -m pytest tests/test_trainer_gpu.py -q --runxfail -k "bf16_params or async_h2d or counters" --timeout=200 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_default.json
B200MPI_BF16_PARAMS=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_bf16params.json
B200MPI_ASYNC_H2D=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_asynch2d.json
B200MPI_ASYNC_H2D=1 B200MPI_BF16_PARAMS=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_asynch2d_bf16params.json
echo "=== 3. tcgen05 GEMM + BN statistics (each case in its own process, bounded waits) ==="
B200MPI_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_gemm_bnstats_gpu.py -q --timeout=200 2>&1 | tail -15
echo "=== 3b. GEMM micro-benchmark vs cuBLAS + statistics pass, and one full ncu capture of the kernel (only if step 3 passed) ==="
B200MPI_EXPERIMENTAL=1 timeout 300 python benchmarks/gemm_bnstats_bench.py --out gpurun_out/gemm_bnstats_bench.json 2>&1 | tail -14
B200MPI_EXPERIMENTAL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_bnstats -c 1 -o gpurun_out/prof_gemm_bnstats \
  python benchmarks/gemm_bnstats_bench.py --iters 1 > gpurun_out/ncu_gemm.log 2>&1
echo "=== 4. bench with the tensor-core 1x1 path (only meaningful if step 3 passed) ==="
B200MPI_FUSED_CONV1X1=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_conv1x1.json
B200MPI_FUSED_CONV1X1=1 B200MPI_BF16_PARAMS=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_conv1x1_bf16params.json
echo "=== 4b. everything that passed, together ==="
B200MPI_ASYNC_H2D=1 B200MPI_BF16_PARAMS=1 B200MPI_DEFER_NBT=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/bench_all_flags.json
echo "=== 5. launch list of the best configuration ==="
B200MPI_BF16_PARAMS=1 B200MPI_DEFER_NBT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 2400 --csv --log-file gpurun_out/launches_bf16params.csv \
  python bench.py --steps 3 --warmup 3 --no-graph > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bf16params.csv 2>/dev/null | head -30
