#!/bin/bash
# One GPU-box session: correctness + YAML-driven run + bench + sweep. Usage: tools/gpu_session.sh N
N=${1:-2}
export B200MPI_NO_AUTOBUILD=1
mkdir -p gpurun_out
echo "=== nvidia-smi ==="; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv | head -12
echo "=== pytest gpu (1-GPU emu + trainer) ==="; timeout 400 python -m pytest tests -x -q -m gpu --timeout=300 2>&1 | tail -5
echo "=== mp_worker N=$N ==="; timeout 200 python tests/mp_launch.py -n $N --timeout 180 tests/mp_worker.py 2>&1 | tail -12
echo "=== MPIJob YAML -> operator -> mpirun -> tf_cnn_benchmarks (2 GPUs) ==="
timeout 300 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml --timeout 280 2>&1 | tail -32
echo "=== hvd MNIST MPIJob (2 GPUs) ==="
timeout 200 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/horovod/tensorflow-mnist.yaml --timeout 180 2>&1 | tail -8
echo "=== bench ours N=$N ==="; timeout 400 python bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E '^\{' | tee gpurun_out/bench_ours_n$N.json
echo "=== bench nccl N=$N ==="; B200MPI_FAULTHANDLER=150 timeout 200 python bench.py --gpus $N --steps 20 --warmup 5 --impl nccl > gpurun_out/bench_nccl_n$N.log 2>&1; grep -E '^\{' gpurun_out/bench_nccl_n$N.log | tee gpurun_out/bench_nccl_n$N.json; tail -25 gpurun_out/bench_nccl_n$N.log | grep -v "^\{" | tail -25
echo "=== bench torchddp N=$N ==="; timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --impl torchddp 2>&1 | grep -E '^\{' | tee gpurun_out/bench_torchddp_n$N.json
echo "=== sweep N=$N ==="; timeout 400 python tests/mp_launch.py -n $N --timeout 380 benchmarks/allreduce_sweep.py --max 1073741824 --dtype float32 --tune-blocks 32,128 --out gpurun_out/sweep_n$N.json 2>&1 | tail -90
