#!/bin/bash
# First GPU call of the next round: diagnose the 8-GPU LD_PRELOAD shim failure with per-rank logs.
# Usage: gpurun --gpus 8 --timeout 600 -- 'tools/gpu_session_next.sh 8'
N=${1:-8}
export B200MPI_NO_AUTOBUILD=1 B200MPI_DEBUG=2
SHIM=$PWD/mpi_operator_b200/lib/libb200mpi_nccl.so
mkdir -p gpurun_out/shim_n$N
echo "=== DDP worker under LD_PRELOAD, N=$N (per-rank logs in gpurun_out/shim_n$N) ==="
LD_PRELOAD=$SHIM timeout 150 python tests/mp_launch.py -n $N --timeout 120 --log-dir gpurun_out/shim_n$N tests/ddp_shim_worker.py
for f in gpurun_out/shim_n$N/*.log; do echo "--- $f"; grep -v "^frame\|^$" $f | tail -25; done
echo "=== ResNet-50 DDP script under LD_PRELOAD, N=$N ==="
mkdir -p gpurun_out/ddp_n$N
LD_PRELOAD=$SHIM timeout 200 python tests/mp_launch.py -n $N --timeout 180 --log-dir gpurun_out/ddp_n$N examples/torch-ddp/torch_ddp_resnet50.py --steps 20 --warmup 5
for f in gpurun_out/ddp_n$N/*.log; do echo "--- $f"; grep -v "^frame\|^$" $f | tail -12; done
echo "=== experimental point-to-point (B200MPI_P2P=1), N=$N ==="
mkdir -p gpurun_out/p2p_n$N
B200MPI_P2P=1 timeout 150 python tests/mp_launch.py -n $N --timeout 120 --log-dir gpurun_out/p2p_n$N tests/p2p_worker.py
for f in gpurun_out/p2p_n$N/*.log; do echo "--- $f"; tail -6 $f; done
echo "=== hvdcore engine with CUDA tensors (B200MPI_HVD_ENGINE=1), N=$N ==="
mkdir -p gpurun_out/hvd_n$N
B200MPI_HVD_ENGINE=1 timeout 150 python tests/mp_launch.py -n $N --timeout 120 --log-dir gpurun_out/hvd_n$N tests/hvd_engine_gpu_worker.py
for f in gpurun_out/hvd_n$N/*.log; do echo "--- $f"; tail -6 $f; done
echo "=== bench at N=$N: default vs async input pipeline (+ NUMA binding) ==="
PORT=29611
for cfg in "" "B200MPI_ASYNC_H2D=1" "B200MPI_ASYNC_H2D=1 B200MPI_BIND_NUMA=1"; do
  PORT=$((PORT+1))
  echo "--- cfg: [$cfg]"
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 30 --warmup 5 2>gpurun_out/bench_n${N}_$PORT.err | tail -1 | tee gpurun_out/bench_n${N}_$PORT.json
done
if [ "$N" = "8" ]; then
  echo "=== elastic 4 -> 8 -> 4 on GPUs, timed ==="
  timeout 300 python benchmarks/elastic_demo.py --total-steps 1500 --step-sleep 0.005 --out gpurun_out/elastic_demo_gpu.json 2>&1 | tail -2
fi
