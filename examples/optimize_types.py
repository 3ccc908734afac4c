#!/usr/bin/env python
"""The reference's examples/optimize_type.py, batched: N character types are drawn from the prior on the device and their
stroke shapes and scales are optimised by Adam to maximise log P(type) (CharacterTypeDist.score_type,
pybpl/model/type_dist.py:98-125), projecting the scales onto their bound after every step as the reference does
(objects/part.py StrokeType.lbs: invscales >= eps).  Relation parameters are left alone: the histogram position score
is piecewise constant and a uniform spot density is flat, so they receive no gradient (in the reference, score_type with
Library(use_hist=True) raises as soon as gpos requires grad)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops  # noqa: E402


def optimise(n_types=4096, steps=100, lr=1e-3, eps=1e-4, seed=0, device="cuda:0"):
    lib = ops.TypeLibrary(device)
    tso, sso, types, subid = ops.sample_types(lib, n_types, seed=seed)
    types.ctrl_type.requires_grad_(True); types.invscale_type.requires_grad_(True)
    opt = torch.optim.Adam([types.ctrl_type, types.invscale_type], lr=lr, fused=True)
    scores = []
    ops.score_type_prior(lib, tso, sso, types, subid).sum().backward()       # warm-up, not timed
    opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        opt.zero_grad(set_to_none=True)
        ll = ops.score_type_prior(lib, tso, sso, types, subid)
        (-ll.sum()).backward()
        opt.step()
        with torch.no_grad():
            types.invscale_type.clamp_(min=eps)
        if it % max(steps // 10, 1) == 0 or it == steps - 1:
            scores.append(ll.detach().mean())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dict(seconds=dt, types_per_s=n_types * steps / dt, scores=[float(s) for s in scores])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--types", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    r = optimise(a.types, a.steps)
    print(f"{a.types} types x {a.steps} Adam steps on log P(type): {r['seconds']:.3f} s = {r['types_per_s'] / 1e6:.2f} M type-steps/s; "
          f"mean log P {r['scores'][0]:.1f} -> {r['scores'][-1]:.1f}")
