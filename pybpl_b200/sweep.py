"""Particle-sweep reduction across ranks: each rank reduces its shard of scores to
(max, sum exp(ll - max)); one all-gather of those two floats per rank gives every rank the global
log-sum-exp.  The per-GPU partial is a CUDA kernel (ops.logsumexp_partial); the combine is 2*world floats."""
import math

import torch


def partial_from_scores(ll):
    """(max, sum exp(ll - max)) as a 2-element float32 tensor on ll's device; CUDA kernel on GPU tensors.
    CPU tensors (gloo tests of the combine logic) use two torch reductions."""
    if ll.is_cuda:
        from . import ops
        return ops.logsumexp_partial(ll)
    if ll.numel() == 0:
        return torch.tensor([-math.inf, 0.0])
    m = ll.max()
    return torch.stack([m, (ll.double() - m.double()).exp().sum().float()])


def combine_partials(part, group=None):
    """All-gather the per-rank partials and return the global log-sum-exp as a Python float."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        out = torch.empty(world * 2, dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(out, part.contiguous().view(2), group=group)
        out = out.view(world, 2)
    else:
        out = part.view(1, 2)
    return logsumexp_from_partials(out)


def logsumexp_from_partials(parts):
    p = parts.detach().double().cpu()
    m = p[:, 0].max()
    if not torch.isfinite(m):
        return float("-inf")
    s = (p[:, 1] * (p[:, 0] - m).exp()).sum()
    return float(m + s.log())
