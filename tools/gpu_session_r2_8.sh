#!/bin/bash
# Round-2 8-GPU session (charged 8x): everything that needs the full box, most important first, every step bounded.
# Usage: gpurun --gpus 8 --timeout 1000 -- 'tools/gpu_session_r2_8.sh'
N=${1:-8}
export B200MPI_NO_AUTOBUILD=1
SHIM=$PWD/mpi_operator_b200/lib/libb200mpi_nccl.so
O=gpurun_out/s8
mkdir -p $O
t0=$(date +%s)
stamp() { echo "=== [$(( $(date +%s) - t0 )) s] $1 ==="; }
stamp "1. multi-GPU tests N=$N (stress 10^4, shim under DDP incl. registered buffers, pass-through, p2p, hvd engine)"
B200MPI_DEBUG=1 MP_LAUNCH_LOG_DIR=$O/mg timeout 420 python -m pytest tests/test_multigpu.py -q --timeout=400 2>&1 | tail -6
grep -h "mp_worker\]\|mp_worker done\|FAILED" $O/mg/mp_worker*.rank0.log | tail -5 | cut -c1-600
grep -h "ddp_shim_worker failures" $O/mg/ddp_shim_worker*.log | sort | uniq -c | head -12 | cut -c1-300
stamp "2. DDP ResNet-50 MPIJob (config #3): injected (default) vs stock NCCL"
timeout 200 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/torch-ddp/resnet50-ddp.yaml --replicas $N --np $N --timeout 180 > $O/ddp_yaml_injected.log 2>&1; grep "images/sec\|Succeeded\|Failed\|Error" $O/ddp_yaml_injected.log | tail -3
B200MPI_INJECT=0 timeout 200 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/torch-ddp/resnet50-ddp.yaml --replicas $N --np $N --timeout 180 > $O/ddp_yaml_nccl.log 2>&1; grep "images/sec\|Succeeded\|Failed\|Error" $O/ddp_yaml_nccl.log | tail -3
stamp "3. bench.py N=$N (ours + same-box arms)"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 30 --warmup 5 2>$O/bench_n$N.err | tail -1 | tee $O/bench_ours_n$N.json | cut -c1-1800
stamp "4. torch.distributed sweep: stock NCCL, then the injected shim"
timeout 300 python tests/mp_launch.py -n $N --timeout 280 benchmarks/collective_sweep.py --tag nccl --iters 8 --out $O/sweep_nccl_n$N.json 2>&1 | tail -1
LD_PRELOAD=$SHIM timeout 300 python tests/mp_launch.py -n $N --timeout 280 --log-dir $O/sweep_shim_logs benchmarks/collective_sweep.py --tag shim --iters 8 --out $O/sweep_shim_n$N.json 2>&1 | tail -1
python benchmarks/roofline_tables.py $O/sweep_nccl_n$N.json $O/sweep_shim_n$N.json --out $O/roofline_shim_vs_nccl_n$N.md 2>&1 | grep "worst\|^## \|^| 1073741824\|^| 16777216 \|^| 1024 " | head -60
stamp "5. native allreduce sweep fp32 (window algos, staged, pipe, reg, NCCL)"
timeout 300 python tests/mp_launch.py -n $N --timeout 280 benchmarks/allreduce_sweep.py --dtype float32 --iters 8 --min 65536 --out $O/allreduce_sweep_n${N}_f32.json 2>&1 | grep " 1073741824 \| 134217728 \| 16777216 \| 1048576 \| 65536 \|exit" | tail -40
stamp "6. tensorflow-benchmarks YAML (config #2): fused engine, then the Horovod API engine"
timeout 150 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml --replicas $N --np $N --timeout 140 > $O/yaml_n${N}_fused.log 2>&1; grep "total images/sec\|Succeeded\|Failed" $O/yaml_n${N}_fused.log | tail -2
B200MPI_ENGINE=hvd timeout 150 python -m mpi_operator_b200.cmd.mpijobctl run -f examples/tensorflow-benchmarks/tensorflow-benchmarks.yaml --replicas $N --np $N --timeout 140 > $O/yaml_n${N}_hvd.log 2>&1; grep "total images/sec\|Succeeded\|Failed" $O/yaml_n${N}_hvd.log | tail -2
stamp "7. elastic 4 -> 8 -> 4 on GPUs, timed"
timeout 300 python benchmarks/elastic_demo.py --total-steps 1500 --step-sleep 0.005 --out $O/elastic_demo_gpu.json 2>&1 | tail -3 | cut -c1-600
stamp "8. elastic GPU test"
timeout 300 python -m pytest tests/test_elastic_gpu.py -q --timeout=280 2>&1 | tail -3
stamp "done"
