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


class Sweep:
    """BASELINE.json configs[3] on one rank: the rank's share of the particles resident on its GPU, one target image.
    step() = ONE fused launch (render + score + the rank's log-sum-exp partial, bpl_score_tokens_lse), one all-gather
    of 2 floats per rank when a process group is up, one tiny combine kernel; nothing synchronises with the host.
    The global log-sum-exp is left in `self.total` (1-element CUDA tensor), the scores in `self.ll`."""

    def __init__(self, batch, images, group=None, want_ll=True, ps=None):
        import torch.distributed as dist
        from . import ops
        self.batch, self.images, self.group, self.want_ll = batch, images, group, want_ll
        self.ps = ops.make_params(ps)
        self.state = ops.LseState(batch.device)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.gathered = torch.empty(2 * self.world, dtype=torch.float32, device=batch.device)
        self.total = None
        self.ll = None

    def step(self, batch=None):
        import torch.distributed as dist
        from . import ops
        b = self.batch if batch is None else batch
        self.ll, _ = ops.score_tokens_lse(b, self.images, self.state, accumulate=False, want_ll=self.want_ll, ps=self.ps)
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.state.partial, group=self.group)
            parts = self.gathered
        else:
            parts = self.state.partial
        self.total = ops.logsumexp_combine(parts)
        return self.total


# ---------------------------------------------------------------------------------------------
# Cross-scoring across ranks (BASELINE.json configs[4]): every token against every target image.
# ---------------------------------------------------------------------------------------------
def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n items for `rank` of `world` (the first n % world ranks hold one item more)."""
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gather_rows(block, n_total, group=None):
    """All-gather row blocks of a [n_total, T] matrix whose rows are sharded contiguously (shard_bounds) over the ranks:
    every rank passes its [n_r, T] block and receives the whole matrix.  Blocks are padded to the largest shard so that
    ONE all_gather_into_tensor moves everything (NCCL on GPU tensors, gloo on CPU tensors in the tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return block
    world = dist.get_world_size(group)
    T = block.shape[1]
    rows = -(-n_total // world)
    pad = torch.zeros((rows, T), dtype=block.dtype, device=block.device)
    pad[:block.shape[0]] = block
    out = torch.empty((world * rows, T), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r * rows:r * rows + (hi - lo)])
    return torch.cat(parts)


class CrossScore:
    """One rank of a cross-scoring job: the rank's contiguous shard of the tokens is resident on its GPU, the packed target
    images are replicated (1 380 B each).  step() renders every local token ONCE and scores it against all images
    (bpl_forward1s_kernel<true>: per-pixel score differences gathered at each image's set bits -- the per-pair value
    is CharacterImageDist.score_image, pybpl/model/image_dist.py:60-80), then all-gathers the [tokens, images] blocks:
    the only exchange, 4 bytes per pair."""

    def __init__(self, batch, images, n_tokens_total, group=None, ps=None):
        from . import ops
        self.batch, self.images, self.n_total, self.group = batch, images, int(n_tokens_total), group
        self.ps = ops.make_params(ps)
        self.ll = None

    def step(self):
        from . import ops
        with torch.no_grad():
            block, self.off_page = ops.score_tokens_all_images(self.batch, self.images, ps=self.ps)
        self.ll = gather_rows(block, self.n_total, self.group)
        return self.ll
