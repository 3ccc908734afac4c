#!/usr/bin/env python
"""BASELINE.json configs[3]: particle-sweep inference.  N token particles (default 1M) are scored against one target
image, sharded across the ranks of a torchrun job (or one GPU), and the ranks' (max, sum-exp) partials are combined
with one NCCL all-gather into the global log-sum-exp.  The particles are tokens drawn on the device from
P(token | type) for a pool of candidate types (bpl_sample_tokens, counter-based streams: every sharding draws the
same particles).

    python examples/particle_sweep.py --particles 1000000
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/particle_sweep.py
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, sweep, workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=131072, help="particles generated and scored per launch")
    ap.add_argument("--types", type=int, default=256, help="candidate types the particles are drawn from")
    ap.add_argument("--token-prior", action="store_true", help="add log P(token | type) to every particle's score")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per = (a.particles + world - 1) // world
    lo, hi = min(a.particles, rank * per), min(a.particles, (rank + 1) * per)
    g = torch.Generator().manual_seed(7)
    target = (torch.rand(105, 105, generator=g) < 0.04).to(torch.uint8).to(dev)
    imgs = ops.pack_images(target)
    # Particles = tokens drawn ON THE DEVICE from P(token | type) for a pool of candidate types (bpl_sample_tokens:
    # token p uses the Philox stream (seed, p), so any sharding sees the same particles); only the type pool
    # (a.types candidate parses) is made on the host.
    pool = workload.synth_tokens(a.types, seed=7, sigma_mode="uniform")
    pool_ty = workload.synth_types(pool, 8)
    K = np.diff(pool["tok_stroke_off"]); nsub = np.diff(pool["stroke_sub_off"])
    sub_tok = np.repeat(np.repeat(np.arange(a.types), K), nsub)
    stroke_tok = np.repeat(np.arange(a.types), K)
    pool_dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
                dict(ctrl=pool["ctrl"], invscale=pool["invscale"], sigma=pool["sigma"], **pool_ty).items()}
    sub_first = np.concatenate([[0], np.cumsum(np.bincount(sub_tok, minlength=a.types))])
    stroke_first = np.concatenate([[0], np.cumsum(K)])
    t_gen = time.perf_counter()
    chunks = []
    for c0 in range(lo, hi, a.chunk):
        n = min(a.chunk, hi - c0)
        tid = (np.arange(c0, c0 + n) % a.types)                       # particle -> candidate type
        ks, ss = K[tid], np.diff(sub_first)[tid]
        tso = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
        ramp = lambda first, cnt: np.repeat(first - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt) + np.arange(int(cnt.sum()))
        stroke_src = ramp(stroke_first[tid], ks)               # the pool's strokes / sub-strokes each particle copies
        sub_src = ramp(sub_first[tid], ss)
        sso = np.concatenate([[0], np.cumsum(nsub[stroke_src])]).astype(np.int32)
        si, ki = torch.from_numpy(sub_src).to(dev), torch.from_numpy(stroke_src).to(dev)
        types = ops.TokenTypes(pool_dev["ctrl"][si], pool_dev["invscale"][si], pool_dev["rel_cat"][ki], pool_dev["attach_ix"][ki],
                               pool_dev["attach_subix"][ki], pool_dev["gpos"][ki], pool_dev["eval_spot_type"][ki],
                               pool_dev["eval_spot"][ki])
        sig = pool_dev["sigma"][torch.from_numpy(tid).to(dev)]
        batch, types = ops.sample_tokens(tso, sso, types, seed=7, first_token=c0, sigma=sig)
        chunks.append((batch, types))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    def run():
        parts = []
        for b, ty in chunks:
            ll, _ = ops.score_tokens(b, imgs)
            if a.token_prior:                                  # log P(image | token) + log P(token | type)
                ll = ll + ops.score_token_prior(b, ty)
            parts.append(ops.logsumexp_partial(ll))
        if parts:
            p = torch.stack(parts)
            m = p[:, 0].max()
            part = torch.stack([m, (p[:, 1] * (p[:, 0] - m).exp()).sum()])
        else:
            part = torch.tensor([-float("inf"), 0.0], device=dev)
        return sweep.combine_partials(part)

    run()                       # warm-up, not timed: CUDA module loads (ours and torch's), clock ramp after the host-side generation
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    total = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"{a.particles} particles on {world} GPU(s): log-sum-exp {total:.4f}; scoring {dt * 1e3:.1f} ms "
              f"= {a.particles / dt / 1e6:.2f} M particles/s; drawing the particles on the device (incl. host index "
              f"arithmetic for the expansion) took {t_gen * 1e3:.0f} ms")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
