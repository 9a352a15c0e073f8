"""hvd.SyncBatchNorm: batch normalisation whose statistics span every rank's batch (Horovod API of the same name).

Forward: one allreduce of [sum, sum of squares, count] per layer; backward: one allreduce of [sum(dy), sum(dy * xhat)].
Both go through ``hvd.allreduce`` (named, so with the background engine the per-layer tensors of a step are negotiated by
name and fused). Numerically this is nn.BatchNorm over the concatenated global batch."""
from __future__ import annotations

import torch
import torch.nn as nn


class _SyncBNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, training, name):
        from . import Sum, allreduce
        red = [0] + list(range(2, x.dim()))
        c = x.shape[1]
        acc = torch.float64 if x.dtype == torch.float64 else torch.float32   # statistics in fp32 (fp64 stays fp64)
        xf = x.to(acc)
        if training:
            local = torch.cat([xf.sum(red), (xf * xf).sum(red), torch.full((1,), float(xf.numel() // c), device=x.device)])
            tot = allreduce(local, op=Sum, name=f"{name}.fwd")
            count = tot[-1]
            mean = tot[:c] / count
            var = (tot[c:2 * c] / count - mean * mean).clamp_min_(0)
            if running_mean is not None:
                with torch.no_grad():
                    running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                    unbiased = var * (count / (count - 1).clamp_min(1))
                    running_var.mul_(1 - momentum).add_(unbiased.to(running_var.dtype), alpha=momentum)
        else:
            mean, var, count = running_mean.to(acc), running_var.to(acc), None
        invstd = torch.rsqrt(var + eps)
        shape = [1, c] + [1] * (x.dim() - 2)
        xhat = (xf - mean.view(shape)) * invstd.view(shape)
        out = xhat
        if weight is not None:
            out = out * weight.to(acc).view(shape) + bias.to(acc).view(shape)
        ctx.save_for_backward(xhat, invstd, weight)
        ctx.training, ctx.name, ctx.count = training, name, count
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        from . import Sum, allreduce
        xhat, invstd, weight = ctx.saved_tensors
        red = [0] + list(range(2, dy.dim()))
        c = dy.shape[1]
        shape = [1, c] + [1] * (dy.dim() - 2)
        dyf = dy.to(xhat.dtype)
        dweight = (dyf * xhat).sum(red) if weight is not None else None
        dbias = dyf.sum(red) if weight is not None else None
        g = dyf * weight.to(xhat.dtype).view(shape) if weight is not None else dyf
        if ctx.training:
            local = torch.cat([g.sum(red), (g * xhat).sum(red)])
            tot = allreduce(local, op=Sum, name=f"{ctx.name}.bwd")
            mean_g, mean_gx = tot[:c] / ctx.count, tot[c:] / ctx.count
            dx = (g - mean_g.view(shape) - xhat * mean_gx.view(shape)) * invstd.view(shape)
        else:
            dx = g * invstd.view(shape)
        return dx.to(dy.dtype), dweight, dbias, None, None, None, None, None, None


class SyncBatchNorm(nn.modules.batchnorm._BatchNorm):
    """Drop-in for ``nn.BatchNorm{1,2,3}d``; statistics and their gradients are reduced over all ranks."""

    _instances = 0

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        SyncBatchNorm._instances += 1
        self._hvd_name = f"sync_batch_norm.{SyncBatchNorm._instances}"

    def _check_input_dim(self, x):
        if x.dim() < 2:
            raise ValueError(f"expected at least 2D input (got {x.dim()}D input)")

    def forward(self, x):
        from . import size
        self._check_input_dim(x)
        training = self.training or not self.track_running_stats
        if not training or size() == 1:
            return super().forward(x)
        if self.momentum is None:
            raise ValueError("SyncBatchNorm needs a numeric momentum (cumulative averaging is not supported)")
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        return _SyncBNFunction.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum, True,
                                     self._hvd_name)
