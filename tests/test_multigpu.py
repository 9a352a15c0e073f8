import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def test_multiprocess_collectives():
    """One rank per GPU (tests/mp_worker.py): every algorithm against fp32 references on REAL peers (NVLS multimem kernels,
    VMM fd export, .sys flag ordering), the pipelined and cudaIpc-registered user-pointer paths, the zero-copy window forms, and
    a stress of 10^4 back-to-back allreduces (random size 4 B - 8 MiB, random algorithm incl. NVLS, pipe and fused allreduce+SGD,
    integer-valued data so every result must be bit-exact, watchdog checked every 100 calls; B200MPI_STRESS_ITERS overrides)."""
    sys.path.insert(0, HERE)
    from mp_launch import launch
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "mp_worker.py")], timeout=240)
    assert rcs == [0] * n


def test_ld_preload_nccl_shim_under_torch_ddp():
    """Unmodified torch.distributed + DDP with LD_PRELOAD=libb200mpi_nccl.so (north star: LD-injection)."""
    sys.path.insert(0, HERE)
    from mp_launch import launch
    shim = os.path.join(os.path.dirname(HERE), "mpi_operator_b200", "lib", "libb200mpi_nccl.so")
    assert os.path.exists(shim), "libb200mpi_nccl.so not built"
    n = 2
    rcs = launch(n, [os.path.join(HERE, "ddp_shim_worker.py")], timeout=240, extra_env={"LD_PRELOAD": shim})
    assert rcs == [0] * n


def test_ld_preload_nccl_shim_under_torch_ddp_all_gpus():
    sys.path.insert(0, HERE)
    from mp_launch import launch
    shim = os.path.join(os.path.dirname(HERE), "mpi_operator_b200", "lib", "libb200mpi_nccl.so")
    n = min(torch.cuda.device_count(), 8)
    if n <= 2:
        pytest.skip("covered by test_ld_preload_nccl_shim_under_torch_ddp")
    rcs = launch(n, [os.path.join(HERE, "ddp_shim_worker.py")], timeout=240, extra_env={"LD_PRELOAD": shim})
    assert rcs == [0] * n


def test_same_script_without_injection_is_the_nccl_baseline():
    """Baseline mode = same launcher, same script, no LD_PRELOAD: stock NCCL (B200MPI_ALGO=nccl semantics)."""
    sys.path.insert(0, HERE)
    from mp_launch import launch
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "ddp_shim_worker.py")], timeout=240)
    assert rcs == [0] * n


def test_ld_preload_shim_passthrough_mode():
    sys.path.insert(0, HERE)
    from mp_launch import launch
    shim = os.path.join(os.path.dirname(HERE), "mpi_operator_b200", "lib", "libb200mpi_nccl.so")
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "ddp_shim_worker.py")], timeout=240, extra_env={"LD_PRELOAD": shim, "B200MPI_ALGO": "nccl"})
    assert rcs == [0] * n


def test_point_to_point_mailboxes():
    """ncclSend/ncclRecv substrate: ring shift, eager send, all-to-all by batches, CUDA-graph replay (tests/p2p_worker.py)."""
    sys.path.insert(0, HERE)
    from mp_launch import launch
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "p2p_worker.py")], timeout=240, extra_env={"B200MPI_P2P": "1"})
    assert rcs == [0] * n


def test_horovod_engine_with_cuda_tensors():
    """Named async allreduces of CUDA tensors, fused on the engine's stream into b200mpi kernels (tests/hvd_engine_gpu_worker.py)."""
    sys.path.insert(0, HERE)
    from mp_launch import launch
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "hvd_engine_gpu_worker.py")], timeout=240, extra_env={"B200MPI_HVD_ENGINE": "1"})
    assert rcs == [0] * n


def test_c10d_backend_on_gpus():
    """torch.distributed backend "b200mpi" (parallel/c10d_backend.py) with CUDA tensors: collectives and DDP on the runtime's
    kernels without LD_PRELOAD (tests/c10d_worker.py; the CPU form runs in tests/test_c10d_backend.py)."""
    sys.path.insert(0, HERE)
    from mp_launch import launch
    n = min(torch.cuda.device_count(), 8)
    rcs = launch(n, [os.path.join(HERE, "c10d_worker.py")], timeout=240)
    assert rcs == [0] * n
