#!/usr/bin/env python
"""Throughput of bpl_score_token_prior (log P(token | type)) against its HBM roofline.
Workload: the golden batch (160 prior-sampled type/token pairs from the reference) tiled to ~1 M tokens, resident in
HBM; forward only and forward + all gradients.  Algorithmic bytes per token: the token's and the type's control points
and scales (2 * 44 B per sub-stroke), position / gpos / spots / relation codes (40 B per stroke), offsets and sigma
(12 B), ll (4 B); with gradients the same arrays are written once more.  (The CPU restatement of this score,
oracle/token_prior.py, is test infrastructure and is timed by tests only: ~13 k tokens/s on one core.)"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybpl_b200 import ops
from tests.util_golden import load
from tests.test_gpu_token_prior import make_inputs

d = dict(load("token_prior"))
reps = 6554
big = {k: (np.concatenate([v] * reps) if k != "consts" else v) for k, v in d.items()}
P, K, S = len(big["K"]), int(big["K"].sum()), int(big["nsub"].sum())
bytes_fwd = 2 * 44 * S + 40 * K + 12 * P + 4 * P
bytes_bwd = bytes_fwd + 2 * 44 * S + 28 * K
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
out = {}
for grad in (False, True):
    batch, types, c, leaves = make_inputs(ops, big, grad=grad)
    def step():
        return ops.score_token_prior(batch, types, c=c)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            step()
        e.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e) / 10)
    nb = bytes_bwd if grad else bytes_fwd
    out["fwd+grad" if grad else "fwd"] = {"ms": best, "tokens_per_s": P / best * 1e3, "GBps": nb / best / 1e6,
                                          "hbm_frac": nb / best / 1e6 / peaks["hbm_gbs"]}
# ---- the sampler on the same structure: writes 44 B per sub-stroke + 12 B per stroke, reads as much of type data ----
batch, types, c, _ = make_inputs(ops, big)
def draw():
    return ops.sample_tokens(batch.tok_stroke_off, batch.stroke_sub_off, types, seed=1, c=c)
for _ in range(3):
    draw()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        draw()
    e.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(e) / 5)
nb = 2 * 44 * S + 44 * K + 8 * P
out["sample"] = {"ms": best, "tokens_per_s": P / best * 1e3, "GBps": nb / best / 1e6, "hbm_frac": nb / best / 1e6 / peaks["hbm_gbs"],
                 "note": "includes torch.empty of the outputs and the eps/sigma fills of ops.sample_tokens"}
# ---- the type sampler: two launches (count, write) + two scans; the prior -> token -> score chain for 1 M particles ----
tl = ops.TypeLibrary("cuda:0")
def draw_types():
    return ops.sample_types(tl, P, seed=3)
for _ in range(2):
    draw_types()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    tso, sso, ty, _ = draw_types()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
out["sample_types"] = {"ms": dt * 1e3, "types_per_s": P / dt, "strokes": int(sso.numel() - 1), "substrokes": int(ty.ctrl_type.shape[0]),
                       "note": "wall clock incl. the size read-back between the two passes"}
out["workload"] = {"tokens": P, "strokes": K, "substrokes": S, "bytes_fwd": bytes_fwd, "bytes_fwd_grad": bytes_bwd}
print(json.dumps(out))
