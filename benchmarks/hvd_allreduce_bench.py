"""Host allreduce latency of the hvd front-end: background engine (data segment, slice-parallel fold) vs the direct libmpi-shim
path (B200MPI_HVD_ENGINE=0).  mpirun -n 4 python benchmarks/hvd_allreduce_bench.py   (N=<elements> to change the size)"""
import sys, time, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch, horovod.torch as hvd
hvd.init()
n = int(os.environ.get("N", 13098536 // 4))
t = torch.ones(n)
for _ in range(3): hvd.allreduce_(t, op=hvd.Sum, name="w")
hvd.barrier()
t0 = time.perf_counter()
K = 10
for i in range(K): hvd.allreduce_(t, op=hvd.Sum, name="x")
dt = (time.perf_counter() - t0) / K
if hvd.rank() == 0: print(f"engine={os.environ.get('B200MPI_HVD_ENGINE','1')} box={hvd.engine_stats().get('mailbox_bytes')} ranks={hvd.size()} {n*4/1e6:.1f} MB: {dt*1e3:.2f} ms  {n*4/dt/1e9:.2f} GB/s")
hvd.shutdown()
