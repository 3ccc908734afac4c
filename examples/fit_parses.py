#!/usr/bin/env python
"""BASELINE.json configs[2]: differentiable fit.  P parses (default 1024) are fitted to one binary 105x105 image by
Adam on control points, inverse scales, positions, blur_sigma and epsilon, differentiating through the fused
render+score kernel (the fit_image loop of pybpl/model/model.py:42-65 / examples/fit_image.ipynb, batched).
Parameters are clamped to the reference's bounds after each step (parameters.py:54-63)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, workload  # noqa: E402


def fit(P=1024, steps=100, lr=4e-3, seed=2024, device="cuda:0", log_every=0):
    dev = torch.device(device)
    w = workload.synth_tokens(P, seed=seed, sigma_mode=10.0)
    batch = workload.to_device(w, dev, requires_grad=True)
    # target: a sample of parse 0's own probability map
    with torch.no_grad():
        pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.array([0])), dev))
        g = torch.Generator(device="cpu").manual_seed(seed)
        target = (torch.rand(105, 105, generator=g).to(dev) < pimg[0]).to(torch.uint8)
    imgs = ops.pack_images(target)
    params = [batch.ctrl, batch.invscale, batch.position, batch.sigma, batch.eps]
    opt = torch.optim.Adam(params, lr=lr, fused=True)      # one multi-tensor kernel per step instead of ~25 small ones
    ops.score_tokens(batch, imgs)[0].sum().backward()        # warm-up (module load, allocator), not timed
    opt.zero_grad(set_to_none=True)
    losses = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        opt.zero_grad(set_to_none=True)
        ll, _ = ops.score_tokens(batch, imgs)
        loss = -ll.sum()
        loss.backward()
        opt.step()
        with torch.no_grad():       # bounds of pybpl/parameters.py:54-63
            batch.sigma.clamp_(0.5, 16.0)
            batch.eps.clamp_(1e-4, 0.5)
            batch.invscale.clamp_(min=1e-3)
        if log_every and it % log_every == 0:
            losses.append(float(loss.detach()))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        ll, _ = ops.score_tokens(batch, imgs)
    return dict(seconds=dt, particles_per_s=P * steps / dt, final_mean_ll=float(ll.mean()), losses=losses)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--parses", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    r = fit(a.parses, a.steps, log_every=10)
    print(f"{a.parses} parses x {a.steps} Adam steps: {r['seconds']:.3f} s = {r['particles_per_s']:.0f} fwd+bwd particles/s; "
          f"mean ll {r['losses'][0] / -a.parses:.1f} -> {r['final_mean_ll']:.1f}")
