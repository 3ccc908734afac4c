#!/usr/bin/env python
"""Fits the cost model behind bpl_order_tokens_by_cost: builds the library with -DBPL_TOKEN_TIMING (cycles per token written
to a debug buffer), runs the forward scoring kernel and the fused backward on a synthetic batch, and regresses the cycles
on (1, area of the control-polygon box, sub-strokes, strokes).  GPU box only (needs nvcc):  python tools/token_cost_fit.py [P]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
so = "/tmp/libbpl_tt.so"
src = [os.path.join(ROOT, "pybpl_b200", "csrc", f) for f in ("bpl_render.cu", "bpl_splines.cu", "bpl_token_prior.cu", "bpl_token_sample.cu", "bpl_type_sample.cu", "bpl_type_prior.cu")]
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-Xcompiler", "-fPIC,-ffp-contract=off",
                       "-shared", "-DBPL_TOKEN_TIMING"] + src + ["-o", so])
import torch
from pybpl_b200 import _lib
_lib.SO_PATH = so
from pybpl_b200 import ops, workload
dev = torch.device("cuda", 0)
w = workload.synth_tokens(P, seed=2024, sigma_mode="uniform")
p0, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), dev))
img = (torch.from_numpy(np.random.default_rng(99).random((105, 105))).to(dev) < p0[0]).to(torch.uint8)
imgs = ops.pack_images(img)
buf = torch.zeros(P + 148 * 64, device=dev)
os.environ["BPL_PT_BUF"] = str(buf.data_ptr())
os.environ["BPL_BWD_TEAMS"] = "1"
# features (host numpy): box of the chained control polygons grown by the reach, as token_cost_kernel computes it
C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy")).astype(np.float64)
tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
S = np.diff(sso[tso]).astype(np.float64); K = np.diff(tso).astype(np.float64)
area = np.zeros(P)
for p in range(P):
    lo = np.array([1e9, 1e9]); hi = -lo
    for k in range(tso[p], tso[p + 1]):
        prev = w["position"][k].astype(np.float64)
        for s in range(sso[k], sso[k + 1]):
            Y = w["invscale"][s] * w["ctrl"][s].astype(np.float64)
            off = C[0] @ Y - prev
            prev = C[-1] @ Y - off
            lo = np.minimum(lo, (Y - off).min(0)); hi = np.maximum(hi, (Y - off).max(0))
    c0, c1 = max(0, lo[0] - 12), min(104, hi[0] + 12)
    r0, r1 = max(0, -hi[1] - 12), min(104, -lo[1] + 12)
    area[p] = max(0, c1 - c0 + 1) * max(0, r1 - r0 + 1)
X = np.stack([np.ones(P), area / 4060.0, S / 5.2, K / 3.0], 1)
for name, grad in (("forward (bpl_forward1s_kernel)", False), ("backward (bpl_backward_teams_kernel)", True)):
    b = workload.to_device(w, dev, requires_grad=grad)
    for _ in range(3):
        buf.zero_()
        ll, _ = ops.score_tokens(b, imgs)
        torch.cuda.synchronize()
    t = buf[:P].cpu().numpy().astype(np.float64)
    for cols, label in (((0, 1, 2), "1, area, S"), ((0, 1, 2, 3), "1, area, S, K")):
        coef, *_ = np.linalg.lstsq(X[:, cols], t, rcond=None)
        pred = X[:, cols] @ coef
        r2 = 1 - ((t - pred) ** 2).sum() / ((t - t.mean()) ** 2).sum()
        print(f"{name}: mean {t.mean():.0f} cycles/token; model [{label}] weights (fraction of the mean) "
              + ", ".join(f"{c / t.mean():.3f}" for c in coef) + f"; R^2 {r2:.3f}; corr(t, S) {np.corrcoef(t, S)[0, 1]:.2f}")
