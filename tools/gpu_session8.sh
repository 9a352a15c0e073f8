#!/bin/bash
# 8-GPU validation session (charged 8x: keep it short)
N=${1:-8}
export B200MPI_NO_AUTOBUILD=1
mkdir -p gpurun_out
echo "=== mp_worker N=$N ==="; timeout 200 python tests/mp_launch.py -n $N --timeout 180 tests/mp_worker.py 2>&1 | tail -14
echo "=== MPIJob YAML x$N GPUs ==="
timeout 300 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml --replicas $N --np $N --timeout 280 2>&1 | tail -20 | tee gpurun_out/yaml_n$N.log
for n in $N 4; do
echo "=== bench ours N=$n ==="; timeout 300 python bench.py --gpus $n --steps 20 --warmup 5 2>&1 | grep -E '^\{' | tee gpurun_out/bench_ours_n$n.json
done
echo "=== bench torchddp N=$N ==="; timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --impl torchddp 2>&1 | grep -E '^\{' | tee gpurun_out/bench_torchddp_n$N.json
echo "=== bench nccl N=$N ==="; B200MPI_FAULTHANDLER=150 timeout 200 python bench.py --gpus $N --steps 20 --warmup 5 --impl nccl 2>&1 | grep -E '^\{' | tee gpurun_out/bench_nccl_n$N.json
echo "=== sweep N=$N ==="; timeout 400 python tests/mp_launch.py -n $N --timeout 380 benchmarks/allreduce_sweep.py --max 1073741824 --dtype float32 --tune-blocks 16,32,128 --iters 12 --out gpurun_out/sweep_n$N.json 2>&1 | tail -120
