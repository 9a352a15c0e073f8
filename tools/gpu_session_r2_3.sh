#!/bin/bash
# Round-2 multi-GPU session B: real-peer tests (pipe / reg / window / stress), shim under DDP, sweeps with the new kernels.
# Usage: gpurun --gpus N --timeout 900 -- 'tools/gpu_session_r2_3.sh N'
N=${1:-4}
export B200MPI_NO_AUTOBUILD=1
SHIM=$PWD/mpi_operator_b200/lib/libb200mpi_nccl.so
O=gpurun_out/s3_n$N
mkdir -p $O
echo "=== 0. emulated pipe / window kernels (1 GPU) ==="
timeout 300 python -m pytest tests/test_collectives_gpu.py -q -x -k "pipelined or interleave or window_allgather" --timeout=250 2>&1 | tail -4
echo "=== 1. multi-GPU tests, N=$N (mp_worker with stress, shim under DDP at all GPUs, pass-through, p2p, hvd engine) ==="
B200MPI_DEBUG=1 MP_LAUNCH_LOG_DIR=$O/mg timeout 600 python -m pytest tests/test_multigpu.py -q --timeout=500 2>&1 | tail -12
for f in $O/mg/mp_worker*.rank0.log; do echo "--- $f"; grep -v "^$" $f | tail -8 | cut -c1-400; done
for f in $O/mg/ddp_shim_worker*.log; do echo "--- $f"; grep -v "^frame\|^$" $f | tail -4 | cut -c1-300; done
echo "=== 2. native allreduce sweep fp32: window algos, staged, pipe, reg, NCCL, N=$N ==="
timeout 400 python tests/mp_launch.py -n $N --timeout 380 benchmarks/allreduce_sweep.py --dtype float32 --iters 8 --min 262144 --out $O/allreduce_sweep_n${N}_f32.json 2>&1 | grep -v "^$" | grep "staged\|pipe\|reg \|nccl\|nvls \|exit" | tail -80
echo "=== 3. torch.distributed sweep: stock NCCL, then the injected shim ==="
timeout 400 python tests/mp_launch.py -n $N --timeout 380 benchmarks/collective_sweep.py --tag nccl --iters 8 --out $O/sweep_nccl_n$N.json 2>&1 | tail -2
LD_PRELOAD=$SHIM timeout 400 python tests/mp_launch.py -n $N --timeout 380 --log-dir $O/sweep_shim_logs benchmarks/collective_sweep.py --tag shim --iters 8 --out $O/sweep_shim_n$N.json 2>&1 | tail -2
tail -3 $O/sweep_shim_logs/*rank0.log
python benchmarks/roofline_tables.py $O/sweep_nccl_n$N.json $O/sweep_shim_n$N.json --out $O/roofline_shim_vs_nccl_n$N.md 2>&1 | grep -v "^| [0-9]* | [0-9.]* | [0-9.]* | 0.0" | tail -120
