#!/usr/bin/env python
"""BASELINE.json configs[4]: cross-scoring.  T types x K tokens are scored against M target images (all pairs): each token
is rendered once and its per-pixel score differences are gathered at every image's set bits.  Under torchrun the tokens
are sharded contiguously over the ranks, the packed images are replicated and the [tokens, images] blocks are
all-gathered (pybpl_b200/sweep.py: CrossScore):

    python examples/cross_score.py [--types 20 --tokens-per-type 20 --images 20]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/cross_score.py --tokens-per-type 2000
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, sweep, workload  # noqa: E402


def make_problem(n_tokens, n_images, seed, dev):
    """Synthetic Omniglot-shaped tokens (strokes ~ the prior's stroke-count distribution, at most 10) and n_images binary
    targets, each a Bernoulli sample of the probability map of one of the tokens (evenly spaced through the batch)."""
    w = workload.synth_tokens(n_tokens, seed=seed, max_strokes=10, kappa="prior", sigma_mode="fixed")
    own = np.arange(0, n_tokens, max(1, n_tokens // n_images))[:n_images]
    with torch.no_grad():
        pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, own), dev))
        g = torch.Generator().manual_seed(seed)
        images = (torch.rand(pimg.shape, generator=g).to(dev) < pimg).to(torch.uint8)
    return w, images, own


def cross_score(n_tokens=400, n_images=20, seed=5, device="cuda:0", reps=1):
    """Single-process entry (tests): returns (ll [n_tokens, n_images], seconds per pass, w, images)."""
    dev = torch.device(device)
    w, images, _ = make_problem(n_tokens, n_images, seed, dev)
    cs = sweep.CrossScore(workload.to_device(w, dev), ops.pack_images(images), n_tokens)
    cs.step()                                           # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ll = cs.step()                                  # every token rendered once, scored against all images
    torch.cuda.synchronize()
    return ll, (time.perf_counter() - t0) / reps, w, images


def main():
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--types", type=int, default=20)
    ap.add_argument("--tokens-per-type", type=int, default=20)
    ap.add_argument("--images", type=int, default=20)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_tokens = a.types * a.tokens_per_type
    w, images, own = make_problem(n_tokens, a.images, 5, dev)      # seeded: identical on every rank
    lo, hi = sweep.shard_bounds(n_tokens, rank, world)
    cs = sweep.CrossScore(workload.to_device(workload.permute(w, np.arange(lo, hi)), dev), ops.pack_images(images), n_tokens)
    for _ in range(3):
        cs.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(a.reps):
        ll = cs.step()
    a1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a0.elapsed_time(a1) / a.reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        acc = (ll[torch.as_tensor(own, device=dev)].argmax(1).cpu() == torch.arange(len(own))).float().mean()
        print(f"{n_tokens} tokens x {a.images} images = {ll.numel()} pairs on {world} GPU(s): {ms.item():.3f} ms per pass = "
              f"{ll.numel() / ms.item() / 1e3:.2f} M pairs/s; checksum {ll.double().sum().item():.6e}; "
              f"each image's own token scores it best in {acc:.0%} of cases")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
