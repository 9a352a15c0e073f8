#!/usr/bin/env python
"""Elastic Horovod-style training (BASELINE.json config #4; reference mechanism:
proposals/elastic-horovod.md:13-31). The world size may change while the job runs:
`mpijobctl scale <job> --replicas M` -> the controller adds/removes workers and rewrites
discover_hosts.sh -> `state.commit()` notices, rank 0's checkpoint is on disk, ranks exit with the
rescale code -> the launcher (restartPolicy OnFailure) re-runs mpirun over the new hostfile ->
`TorchState.restore()/sync()` resumes from the committed step on the new world."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))

import torch
import torch.nn.functional as F

import horovod.torch as hvd
from mpi_operator_b200.models import MnistConvNet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-steps", type=int, default=600)
    ap.add_argument("--commit-every", type=int, default=10)
    ap.add_argument("--checkpoint", default=os.environ.get("B200MPI_ELASTIC_CHECKPOINT", "/tmp/elastic_mnist.pt"))
    ap.add_argument("--step-sleep", type=float, default=0.0)
    a = ap.parse_args()
    hvd.init()
    torch.manual_seed(1)
    dev = torch.device("cuda") if torch.cuda.is_available() and os.environ.get("B200MPI_HVD_DEVICE") != "cpu" else torch.device("cpu")
    model = MnistConvNet().to(dev)
    base_opt = torch.optim.SGD(model.parameters(), lr=0.01 * hvd.size(), momentum=0.9)
    opt = hvd.DistributedOptimizer(base_opt, named_parameters=model.named_parameters(), op=hvd.Average)
    state = hvd.elastic.TorchState(model=model, optimizer=base_opt, checkpoint_path=a.checkpoint, step=0, worlds=[])
    g = torch.Generator().manual_seed(7)
    protos = torch.randn(10, 784, generator=g).to(dev)

    @hvd.elastic.run
    def train(state):
        if not state.worlds or state.worlds[-1] != hvd.size():
            state.worlds = list(state.worlds) + [hvd.size()]
        if hvd.rank() == 0:
            print(f"[elastic] (re)started at step {state.step} with world size {hvd.size()}; worlds so far {state.worlds}", flush=True)
        while state.step < a.total_steps:
            y = torch.randint(0, 10, (64,), device=dev)
            x = protos[y] + 0.5 * torch.randn(64, 784, device=dev)
            opt.zero_grad()
            loss = F.cross_entropy(model(x), y)
            loss.backward()
            opt.step()
            state.step += 1
            if a.step_sleep:
                time.sleep(a.step_sleep)
            if state.step % a.commit_every == 0:
                state.commit()  # rank-0 checkpoint, then host-set check (may end this incarnation)
        return float(loss)

    final = train(state)
    if hvd.rank() == 0:
        print(f"[elastic] finished {state.step} steps, final loss {final:.4f}, world sizes seen: {state.worlds}", flush=True)
    hvd.shutdown()


if __name__ == "__main__":
    main()
