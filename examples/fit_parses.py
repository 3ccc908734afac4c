#!/usr/bin/env python
"""BASELINE.json configs[2]: differentiable fit.  P parses (default 1024) are fitted to one binary 105x105 image by
Adam on control points, inverse scales, positions, blur_sigma and epsilon, differentiating through the fused
render+score kernel (the fit_image loop of pybpl/model/model.py:42-65 / examples/fit_image.ipynb, batched).
Parameters are clamped to the reference's bounds after each step (parameters.py:54-63).  With --token-prior the
objective also carries log P(token | type) (model.score_token, pybpl/model/model.py:59-62) for synthetic types, and
the relation tokens' evaluation spots join the optimised leaves."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, workload  # noqa: E402


def fit(P=1024, steps=100, lr=4e-3, seed=2024, device="cuda:0", log_every=0, token_prior=False, graph=True, target=None):
    dev = torch.device(device)
    w = workload.synth_tokens(P, seed=seed, sigma_mode=10.0)
    batch = workload.to_device(w, dev, requires_grad=True)
    if target is None:      # a sample of parse 0's own probability map
        with torch.no_grad():
            pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.array([0])), dev))
            g = torch.Generator(device="cpu").manual_seed(seed)
            target = (torch.rand(105, 105, generator=g).to(dev) < pimg[0]).to(torch.uint8)
    else:                   # given (the parity test passes the image the reference's goldens were made with)
        target = torch.as_tensor(np.asarray(target), dtype=torch.uint8, device=dev)
    imgs = ops.pack_images(target)
    params = [batch.ctrl, batch.invscale, batch.position, batch.sigma, batch.eps]
    types = prior_c = None
    if token_prior:
        types = workload.types_to_device(workload.synth_types(w, seed + 1), dev, requires_grad=True)
        prior_c = ops.token_prior_params()
        params.append(types.eval_spot)

    def objective():
        ll, _ = ops.score_tokens(batch, imgs)
        if token_prior:
            ll = ll + ops.score_token_prior(batch, types, c=prior_c)
        return ll
    # One optimisation step = the fused forward+backward kernel, one launch that scales its gradients by the cotangent,
    # and ONE launch for Adam + the parameter bounds of pybpl/parameters.py:54-63 (ops.FusedAdam: torch's update rule;
    # torch's own fused multi-tensor Adam takes 66 us on these five small tensors, plus three clamp kernels), with the
    # reduction of the objective around it; captured once as a CUDA graph it replays as a single submission (graph=False
    # keeps the eager loop for comparison).  The objective sum(ll) is ASCENDED (maximize): no negation kernels.
    bounds = {3: (0.5, 16.0), 4: (1e-4, 0.5), 1: (1e-3, None)}            # blur_sigma, epsilon, invscale
    if token_prior:
        bounds[5] = (2.0 + 1e-4, 6.0 - 1e-4)                               # RelationToken.lbs/ubs (objects/relation.py:96-131)
    opt = ops.FusedAdam(params, lr=lr, bounds=bounds, maximize=True)

    def step():
        opt.zero_grad()
        obj = objective().sum()
        obj.backward()
        opt.step()
        return obj.detach()            # (the loss is its negative; negated on the host when logged)

    # warm-up (module load, allocator, lazily built index maps), then everything it touched is put back
    init = [p.detach().clone() for p in params]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    with torch.no_grad():
        for p, p0 in zip(params, init):
            p.copy_(p0)
        opt.reset()
    g = None
    if graph:
        g = torch.cuda.CUDAGraph()
        opt.zero_grad()
        with torch.cuda.graph(g):
            loss_buf = step()
        with torch.no_grad():      # the capture itself does not run the step, but be explicit about the start state
            for p, p0 in zip(params, init):
                p.copy_(p0)
    losses = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        if g is not None:
            g.replay()
            loss = loss_buf
        else:
            loss = step()
        if log_every and it % log_every == 0:
            losses.append(-float(loss.detach()))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        ll = objective()
    return dict(seconds=dt, particles_per_s=P * steps / dt, final_mean_ll=float(ll.mean()), losses=losses)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--parses", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--token-prior", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    r = fit(a.parses, a.steps, log_every=0, token_prior=a.token_prior, graph=not a.no_graph)
    r["losses"] = fit(a.parses, a.steps, log_every=10, token_prior=a.token_prior, graph=not a.no_graph)["losses"]
    print(f"{a.parses} parses x {a.steps} Adam steps: {r['seconds']:.3f} s = {r['particles_per_s']:.0f} fwd+bwd particles/s; "
          f"mean ll {r['losses'][0] / -a.parses:.1f} -> {r['final_mean_ll']:.1f}")
