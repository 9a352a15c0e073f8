#!/bin/bash
# Round-2 multi-GPU session: pipelined kernel numerics (emulated), NCCL-ABI shim with per-rank traces, sweeps.
# Usage: gpurun --gpus N --timeout 900 -- 'tools/gpu_session_r2_2.sh N'
N=${1:-2}
export B200MPI_NO_AUTOBUILD=1
SHIM=$PWD/mpi_operator_b200/lib/libb200mpi_nccl.so
O=gpurun_out/s2_n$N
mkdir -p $O
echo "=== 0. pipelined kernel, emulated ranks on one GPU ==="
timeout 300 python -m pytest tests/test_collectives_gpu.py -q -x -k "pipelined or interleave" --timeout=250 2>&1 | tail -4
echo "=== 1. DDP worker under LD_PRELOAD, direct launch, N=$N (traces in $O/shim_direct) ==="
B200MPI_DEBUG=2 LD_PRELOAD=$SHIM timeout 150 python tests/mp_launch.py -n $N --timeout 120 --log-dir $O/shim_direct tests/ddp_shim_worker.py
for f in $O/shim_direct/*.log; do echo "--- $f"; grep -v "^frame\|^$" $f | grep -v "nccl shim.*ncclGroup\|ncclCommGetAsyncError" | tail -8; done
echo "=== 2. the same through pytest (the configuration that failed in round 1) ==="
B200MPI_DEBUG=2 MP_LAUNCH_LOG_DIR=$O/shim_pytest timeout 400 python -m pytest tests/test_multigpu.py -q --timeout=300 2>&1 | tail -12
for f in $O/shim_pytest/ddp_shim_worker*.log; do echo "--- $f"; grep -v "^frame\|^$" $f | grep -v "ncclGroup\|ncclCommGetAsyncError" | tail -6; done
echo "=== 3. native sweep fp32 (window algos, staged, pipe, NCCL) N=$N ==="
timeout 400 python tests/mp_launch.py -n $N --timeout 380 benchmarks/allreduce_sweep.py --dtype float32 --iters 10 --min 65536 --out $O/allreduce_sweep_n${N}_f32.json 2>&1 | grep -v "^$" | tail -75
echo "=== 4. torch.distributed sweep: stock NCCL, then the injected shim ==="
timeout 400 python tests/mp_launch.py -n $N --timeout 380 benchmarks/collective_sweep.py --tag nccl --out $O/sweep_nccl_n$N.json 2>&1 | tail -3
LD_PRELOAD=$SHIM timeout 400 python tests/mp_launch.py -n $N --timeout 380 --log-dir $O/sweep_shim_logs benchmarks/collective_sweep.py --tag shim --out $O/sweep_shim_n$N.json 2>&1 | tail -3
tail -5 $O/sweep_shim_logs/*rank0.log
python benchmarks/roofline_tables.py $O/sweep_nccl_n$N.json $O/sweep_shim_n$N.json 2>&1 | tail -80
