"""Seeded synthetic token batches for the bench configs and the parity tests.

Tokens are drawn the way the reference's prior shapes them, without running the reference's
(28 ms/type) sampler: K strokes, nsub sub-strokes per stroke from pmat_nsub, primitive ids uniform over
the 1212 library primitives, control points ~ N(mu[id], sd[id]) + token noise, scales ~ Gamma(theta[id]) +
token noise, first stroke position uniform over the canvas, later strokes attached to the start or end of
an earlier stroke.  The hyper-parameter tables (pybpl_b200/data/workload_lib.npz) are data exported from
the reference's lib_data by tests/golden/make_golden.py:make_assets.  Everything is numpy on the host:
this is input generation, not the hot path.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = dict(np.load(os.path.join(_HERE, "data", "workload_lib.npz")))
    return _LIB


def synth_tokens(P, seed=1234, max_strokes=5, kappa="uniform", sigma_mode="uniform", eps=1e-4, neval=200):
    """Packed ragged batch of P tokens as numpy arrays (the layout of include/pybpl_b200.h).

    kappa: "uniform" (K ~ U{1..max_strokes}, config 2) or "prior" (K ~ pkappa truncated at max_strokes, config 5).
    sigma_mode: "uniform" (blur_sigma ~ U(0.5,16)), "fixed" (0.5) or a float.
    """
    L = _lib()
    rng = np.random.default_rng(seed)
    nprim = L["mu"].shape[0]
    if kappa == "prior":
        pk = L["pkappa"][:max_strokes].astype(np.float64)
        K = rng.choice(np.arange(1, max_strokes + 1), size=P, p=pk / pk.sum())
    else:
        K = rng.integers(1, max_strokes + 1, size=P)
    Ktot = int(K.sum())
    tok_of_stroke = np.repeat(np.arange(P), K)
    # sub-strokes per stroke ~ pmat_nsub[K-1]
    pm = L["pmat_nsub"].astype(np.float64)
    pm = pm / pm.sum(1, keepdims=True)
    u = rng.random(Ktot)
    cdf = np.cumsum(pm[K[tok_of_stroke] - 1], axis=1)
    nsub = 1 + (u[:, None] > cdf).sum(1)
    nsub = np.minimum(nsub, pm.shape[1]).astype(np.int64)
    Stot = int(nsub.sum())
    ids = rng.integers(0, nprim, size=Stot)
    ctrl = L["mu"][ids] + L["sd"][ids] * rng.standard_normal((Stot, 10)).astype(np.float32)
    ctrl = ctrl + float(L["sigma_shape"]) * rng.standard_normal((Stot, 10)).astype(np.float32)
    th = L["theta"][ids].astype(np.float64)
    inv = rng.gamma(th[:, 0], th[:, 1])
    for _ in range(50):                                   # token noise, resampled until positive
        noisy = inv + float(L["sigma_invscale"]) * rng.standard_normal(Stot)
        bad = noisy <= 0
        if not bad.any():
            break
        noisy[bad] = inv[bad]
    inv = np.maximum(noisy, 1e-3)
    ctrl = ctrl.reshape(Stot, 5, 2).astype(np.float32)
    inv = inv.astype(np.float32)
    tso = np.zeros(P + 1, np.int64); tso[1:] = np.cumsum(K)
    sso = np.zeros(Ktot + 1, np.int64); sso[1:] = np.cumsum(nsub)
    # positions: stroke 0 uniform on the canvas; later strokes at the start/end of a random earlier stroke
    # (+ N(0, rel_sigma)); end points are estimated from the control polygons, which is all the attachment
    # needs to look Omniglot-shaped.
    pos = np.zeros((Ktot, 2), np.float32)
    span = (ctrl[:, -1, :] - ctrl[:, 0, :]) * inv[:, None]      # approx. start->end displacement per sub-stroke
    stroke_span = np.add.reduceat(span, sso[:-1], axis=0) if Stot else np.zeros((0, 2), np.float32)
    first = np.zeros(Ktot, bool); first[tso[:-1][K > 0]] = True
    pos[first, 0] = rng.random(first.sum()) * 104.0
    pos[first, 1] = -rng.random(first.sum()) * 104.0
    rel = L["rel_sigma"]
    kk = np.arange(Ktot) - tso[tok_of_stroke]                  # stroke rank inside its token
    for rank in range(1, max_strokes):
        sel = np.nonzero(kk == rank)[0]
        if sel.size == 0:
            continue
        prev = tso[tok_of_stroke[sel]] + (rng.random(sel.size) * rank).astype(np.int64)
        at_end = rng.random(sel.size) < 0.5
        base = pos[prev] + np.where(at_end[:, None], stroke_span[prev], 0.0)
        pos[sel] = base + rel * rng.standard_normal((sel.size, 2)).astype(np.float32)
    if sigma_mode == "uniform":
        sigma = rng.uniform(0.5, 16.0, size=P)
    elif sigma_mode == "fixed":
        sigma = np.full(P, 0.5)
    else:
        sigma = np.full(P, float(sigma_mode))
    return dict(P=P, tok_stroke_off=tso.astype(np.int32), stroke_sub_off=sso.astype(np.int32), ctrl=ctrl, invscale=inv,
                position=pos.astype(np.float32), eps=np.full(P, eps, np.float32), sigma=sigma.astype(np.float32),
                neval=neval)


def synth_types(w, seed=4321):
    """Type-level parameters and relations for the tokens of `w` (numpy arrays in the layout of ops.TokenTypes /
    include/pybpl_b200.h:bpl_token_type), for the token-prior kernel: each token is read as a noisy copy of its type
    (shapes +- sigma_shape, scales +- sigma_invscale), a token's first stroke is independent ('unihist', gpos near its
    position), later strokes attach to the start, the end or a point along a random earlier stroke."""
    L = _lib()
    rng = np.random.default_rng(seed)
    tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
    K = np.diff(tso); Ktot, Stot = int(tso[-1]), int(sso[-1])
    nsub = np.diff(sso)
    ctrl_type = (w["ctrl"] + float(L["sigma_shape"]) * rng.standard_normal(w["ctrl"].shape)).astype(np.float32)
    inv_type = np.abs(w["invscale"] + float(L["sigma_invscale"]) * rng.standard_normal(Stot)).astype(np.float32) + 1e-3
    rank = np.arange(Ktot) - np.repeat(tso[:-1], K)
    cat = np.where(rank == 0, 0, rng.integers(1, 4, size=Ktot)).astype(np.int32)
    attach_ix = np.where(rank == 0, 0, (rng.random(Ktot) * np.maximum(rank, 1)).astype(np.int64)).astype(np.int32)
    att_stroke = np.repeat(tso[:-1], K) + attach_ix
    attach_subix = (rng.random(Ktot) * nsub[att_stroke]).astype(np.int32)
    gpos = (w["position"] + L["rel_sigma"] * rng.standard_normal((Ktot, 2))).astype(np.float32)
    es_type = rng.uniform(2.5, 5.5, size=Ktot).astype(np.float32)
    es = np.clip(es_type + 0.1499 * rng.standard_normal(Ktot), 2.05, 5.95).astype(np.float32)
    return dict(ctrl_type=ctrl_type, invscale_type=inv_type, rel_cat=cat, attach_ix=attach_ix, attach_subix=attach_subix,
                gpos=gpos, eval_spot_type=es_type, eval_spot=es)


def types_to_device(ty, device, requires_grad=False):
    import torch
    from . import ops
    dev = torch.device(device)
    f = lambda k, g: torch.from_numpy(ty[k]).to(dev).requires_grad_(g)
    return ops.TokenTypes(f("ctrl_type", False), f("invscale_type", False), ty["rel_cat"], ty["attach_ix"], ty["attach_subix"],
                          f("gpos", False), f("eval_spot_type", False), f("eval_spot", requires_grad))


def permute(w, perm):
    """The same batch with tokens reordered / subset by `perm` (token indices)."""
    tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
    strokes = np.concatenate([np.arange(tso[p], tso[p + 1]) for p in perm]) if len(perm) else np.zeros(0, np.int64)
    subs = np.concatenate([np.arange(sso[k], sso[k + 1]) for k in strokes]) if len(strokes) else np.zeros(0, np.int64)
    K = (tso[1:] - tso[:-1])[perm]
    ns = (sso[1:] - sso[:-1])[strokes]
    tso2 = np.zeros(len(perm) + 1, np.int64); tso2[1:] = np.cumsum(K)
    sso2 = np.zeros(len(strokes) + 1, np.int64); sso2[1:] = np.cumsum(ns)
    return dict(P=len(perm), tok_stroke_off=tso2.astype(np.int32), stroke_sub_off=sso2.astype(np.int32),
                ctrl=w["ctrl"][subs], invscale=w["invscale"][subs], position=w["position"][strokes],
                eps=w["eps"][perm], sigma=w["sigma"][perm], neval=w["neval"])


def shard(w, rank, world):
    """Contiguous block of ceil(P/world) tokens for `rank`."""
    P = w["P"]
    per = (P + world - 1) // world
    lo, hi = min(P, rank * per), min(P, (rank + 1) * per)
    return permute(w, np.arange(lo, hi))


def to_device(w, device, requires_grad=False, pin=False):
    """ops.TokenBatch on `device` from the numpy batch."""
    import torch
    from . import ops

    def f(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        if pin:
            t = t.pin_memory()
        t = t.to(device, non_blocking=pin)
        return t.requires_grad_(True) if requires_grad and t.is_floating_point() else t
    return ops.TokenBatch(f(w["tok_stroke_off"]), f(w["stroke_sub_off"]), f(w["ctrl"]), f(w["invscale"]),
                          f(w["position"]), f(w["eps"]), f(w["sigma"]), neval=w["neval"], check_limits=False)



def concat(ws):
    """Concatenate numpy batches (CSR offsets rebased)."""
    ws = list(ws)
    if len(ws) == 1:
        return ws[0]
    tso, sso = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    k = s = 0
    for w in ws:
        tso.append(w["tok_stroke_off"][1:].astype(np.int64) + k)
        sso.append(w["stroke_sub_off"][1:].astype(np.int64) + s)
        k += int(w["tok_stroke_off"][-1]); s += int(w["stroke_sub_off"][-1])
    cat = lambda key: np.concatenate([w[key] for w in ws])
    return dict(P=sum(w["P"] for w in ws), tok_stroke_off=np.concatenate(tso).astype(np.int32),
                stroke_sub_off=np.concatenate(sso).astype(np.int32), ctrl=cat("ctrl"), invscale=cat("invscale"),
                position=cat("position"), eps=cat("eps"), sigma=cat("sigma"), neval=ws[0]["neval"])


# BASELINE.json configs[3]: the particle sweep.  The particle set is defined block-wise (block b = synth_tokens with seed
# SWEEP_SEED + b) so that every sharding over 1/2/4/8 ranks -- and the CPU arms, which take the first blocks -- sees the
# very same particles without anyone generating more than its share.
SWEEP_PARTICLES = 1_000_000
SWEEP_BLOCKS = 64
SWEEP_SEED = 7000


def sweep_block(b, particles=SWEEP_PARTICLES, blocks=SWEEP_BLOCKS):
    per = particles // blocks
    n = per + (particles - per * blocks if b == blocks - 1 else 0)
    return synth_tokens(n, seed=SWEEP_SEED + b, sigma_mode="uniform")


def sweep_shard(rank, world, particles=SWEEP_PARTICLES, blocks=SWEEP_BLOCKS):
    """Rank `rank`'s contiguous share of the sweep's blocks as one numpy batch (strong scaling: blocks / world each)."""
    per = (blocks + world - 1) // world
    lo, hi = min(blocks, rank * per), min(blocks, (rank + 1) * per)
    if lo >= hi:
        return permute(sweep_block(0, particles, blocks), np.zeros(0, np.int64))
    return concat([sweep_block(b, particles, blocks) for b in range(lo, hi)])
