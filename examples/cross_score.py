#!/usr/bin/env python
"""BASELINE.json configs[4]: cross-scoring.  T types x K tokens are scored against M target images (all pairs):
each token is rendered once and its per-pixel score differences are gathered at every image's set bits."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, workload  # noqa: E402


def cross_score(n_tokens=400, n_images=20, seed=5, device="cuda:0"):
    dev = torch.device(device)
    w = workload.synth_tokens(n_tokens, seed=seed, max_strokes=10, kappa="prior", sigma_mode="fixed")
    with torch.no_grad():
        pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(0, n_tokens, n_tokens // n_images)[:n_images]), dev))
        g = torch.Generator().manual_seed(seed)
        images = (torch.rand(pimg.shape, generator=g).to(dev) < pimg).to(torch.uint8)
    imgs = ops.pack_images(images)
    batch = workload.to_device(w, dev)
    ops.score_tokens_all_images(batch, imgs)            # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ll, _ = ops.score_tokens_all_images(batch, imgs)    # every token rendered once, scored against all images
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return ll, dt, w, images


if __name__ == "__main__":
    ll, dt, _, _ = cross_score()
    step = ll.shape[0] // ll.shape[1]
    own = torch.arange(0, ll.shape[0], step)[:ll.shape[1]]        # the tokens the images were sampled from
    acc = (ll[own].argmax(1).cpu() == torch.arange(ll.shape[1])).float().mean()
    print(f"{ll.shape[0]} tokens x {ll.shape[1]} images = {ll.numel()} pairs in {dt * 1e3:.2f} ms; "
          f"each image's own token scores it best in {acc:.0%} of cases")
