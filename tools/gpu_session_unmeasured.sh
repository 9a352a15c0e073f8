#!/bin/bash
# Everything that was written or changed after the round's GPU minutes were spent (right-hand column), as one
# bounded session. Usage on a box with N GPUs:   tools/gpu_session_unmeasured.sh [N]      (results under gpurun_out/final/)
N=${1:-8}
export B200MPI_NO_AUTOBUILD=1
O=gpurun_out/final
mkdir -p $O
t0=$(date +%s)
stamp() { echo "=== [$(( $(date +%s) - t0 )) s] $1 ==="; }
stamp "0. smoke + single-GPU tier with the final defaults (tail bucket, deferred BN counters, M >= 128 GEMM gate)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m "gpu and not multigpu" -x -q 2>&1 | tail -3
stamp "1. bench.py 1 GPU (supervisor, self-check, same-box arms)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1 | tee $O/bench_n1.json | cut -c1-600
if [ "$N" -ge 2 ]; then
  stamp "2. multi-GPU tier on $N GPUs (stock / pass-through teardown fix, c10d backend on CUDA tensors)"
  MP_LAUNCH_LOG_DIR=$O/mg timeout 900 python -m pytest tests/test_multigpu.py tests/test_elastic_gpu.py -q --timeout=800 2>&1 | tail -4
  for n in 2 4 8; do
    [ "$n" -le "$N" ] || continue
    stamp "3. bench.py $n GPUs: default and bf16-shadow + tail-bucket candidates, same-box arms"
    timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py \
      --gpus $n --steps 20 --warmup 5 2>$O/bench_n$n.err | tail -1 | tee $O/bench_n$n.json | cut -c1-900
  done
  stamp "3b. bench.py $N GPUs with every rank bound next to its GPU (B200MPI_BIND_TO=numa; off by default until this says it pays)"
  B200MPI_BIND_TO=numa B200MPI_BENCH_SAME_BOX=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29644 bench.py \
    --gpus $N --steps 20 --warmup 5 2>$O/bench_n${N}_numa.err | tail -1 | tee $O/bench_n${N}_numa.json | cut -c1-400
  stamp "4. step timeline at $N GPUs (exposed comm / copy time per step)"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 benchmarks/step_timeline.py \
    --out $O/step_timeline_n$N.md 2>&1 | tail -3
  stamp "5. the reference's YAML through the operator with injection on (the two-runtime-copies fix), fused and Horovod-API engines"
  timeout 200 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml --replicas $N --np $N --timeout 180 > $O/yaml_fused.log 2>&1
  grep "total images/sec\|Succeeded\|Failed" $O/yaml_fused.log | tail -2
  B200MPI_ENGINE=hvd timeout 200 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml --replicas $N --np $N --timeout 180 > $O/yaml_hvd.log 2>&1
  grep "total images/sec\|Succeeded\|Failed" $O/yaml_hvd.log | tail -2
  stamp "6. torch DDP ResNet-50: c10d backend vs injected shim vs stock NCCL"
  for be in "--backend b200mpi" "--backend nccl"; do
    LD_PRELOAD= timeout 200 python tests/mp_launch.py -n $N --timeout 180 examples/torch-ddp/torch_ddp_resnet50.py $be --steps 40 2>&1 | grep "images/sec" | tail -1
  done
  LD_PRELOAD=$PWD/mpi_operator_b200/lib/libb200mpi_nccl.so timeout 200 python tests/mp_launch.py -n $N --timeout 180 examples/torch-ddp/torch_ddp_resnet50.py --steps 40 2>&1 | grep "images/sec" | tail -1
  stamp "7. broadcast / reduce-scatter through the shim after the selection changes"
  SHIM=$PWD/mpi_operator_b200/lib/libb200mpi_nccl.so
  timeout 300 python tests/mp_launch.py -n $N --timeout 280 benchmarks/collective_sweep.py --tag nccl --iters 8 --out $O/sweep_nccl.json 2>&1 | tail -1
  LD_PRELOAD=$SHIM timeout 300 python tests/mp_launch.py -n $N --timeout 280 benchmarks/collective_sweep.py --tag shim --iters 8 --out $O/sweep_shim.json 2>&1 | tail -1
  python benchmarks/roofline_tables.py $O/sweep_nccl.json $O/sweep_shim.json --out $O/roofline_shim_vs_nccl.md 2>&1 | grep "worst\|^## " | head -20
  stamp "8. elastic 4 -> 8 -> 4 in place"
  B200MPI_ELASTIC_INPLACE=1 timeout 300 python benchmarks/elastic_demo.py --total-steps 1500 --step-sleep 0.005 --out $O/elastic_inplace_gpu.json 2>&1 | tail -2
fi
stamp "9. compute-sanitizer (memcheck / racecheck / synccheck / initcheck)"
timeout 900 tools/sanitize.sh 2>&1 | tail -6
stamp "done"
