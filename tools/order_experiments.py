#!/usr/bin/env python
"""Which processing order suits the work queue at the bench size?  (expensive-first balances the end of the launch but
puts tokens of the same size -- hence the same phase mix -- on an SM at the same time.)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybpl_b200 import ops, workload
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = workload.synth_tokens(P, seed=1234, sigma_mode="uniform")
b = workload.to_device(w, dev)
imgs = ops.pack_images((torch.rand(105, 105, device=dev) < 0.04).to(torch.uint8))
S = np.diff(w["stroke_sub_off"][w["tok_stroke_off"]])
rng = np.random.default_rng(0)
def classes(n):
    q = np.quantile(S, np.linspace(0, 1, n + 1)[1:-1])
    c = np.searchsorted(q, S, side="left")          # 0 = lightest
    key = -c * 1e6 + rng.random(P)                  # heavy classes first, random inside a class
    return np.argsort(key)
def tail_sorted(frac):
    idx = np.argsort(-S, kind="stable")
    head = idx[: int(P * (1 - frac))].copy(); rng.shuffle(head)
    return np.concatenate([head, idx[int(P * (1 - frac)):]])
cands = {"index order": np.arange(P), "expensive first (LPT)": np.argsort(-S, kind="stable"), "2 classes": classes(2),
         "4 classes": classes(4), "8 classes": classes(8), "random head, sorted 25% tail": tail_sorted(0.25),
         "random head, sorted 50% tail": tail_sorted(0.5), "LPT interleaved x3": None}
lpt = np.argsort(-S, kind="stable")
cands["LPT interleaved x3"] = np.concatenate([lpt[i::3] for i in range(3)]) if False else lpt.reshape(-1)[np.argsort(np.arange(P) % 3, kind="stable")] if False else None
del cands["LPT interleaved x3"]
K = np.diff(w["tok_stroke_off"])
sig = w["sigma"]
cands = {"expensive first (LPT by S)": np.argsort(-S, kind="stable"), "by S then K": np.lexsort((-K, -S)), "by K then S": np.lexsort((-S, -K)),
         "by S + K": np.argsort(-(S + K), kind="stable"), "by S + 2K": np.argsort(-(S + 2 * K), kind="stable"),
         "by 2S + K": np.argsort(-(2 * S + K), kind="stable")}
# extent proxy: spread of the stroke positions (a token whose strokes start far apart has a large box)
pos = w["position"]; tso = w["tok_stroke_off"]
ext = np.array([np.ptp(pos[tso[i]:tso[i + 1], 0]) + np.ptp(pos[tso[i]:tso[i + 1], 1]) if tso[i + 1] > tso[i] else 0 for i in range(P)])
cands["by S + extent/20"] = np.argsort(-(S + ext / 20.0), kind="stable")
cands["by S + extent/10"] = np.argsort(-(S + ext / 10.0), kind="stable")
for name, order in cands.items():
    b._order = b._order_cost = torch.from_numpy(order.astype(np.int32)).to(dev)
    for _ in range(5):
        ops.score_tokens(b, imgs)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            ops.score_tokens(b, imgs)
        e.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e) / 20)
    print(f"{name:32s} {P / best / 1e3:7.3f} M particles/s")
