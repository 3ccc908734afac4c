#!/usr/bin/env python
"""How much can a processing order buy the scoring kernel on the 4 096-token step?  Orders tried: index, by sub-stroke count
(the shipped one), by the cost key of bpl_order_tokens_by_cost, and by MEASURED cycles per token (a -DBPL_TOKEN_TIMING build
writes them to a debug buffer) -- the last is the upper bound for any key.  GPU box only (needs nvcc)."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
BWD = len(sys.argv) > 2 and sys.argv[2] == "bwd"      # the fused backward (teams kernel) instead of the scoring kernel
so = "/tmp/libbpl_tt.so"
src = [os.path.join(ROOT, "pybpl_b200", "csrc", f) for f in ("bpl_render.cu", "bpl_splines.cu", "bpl_token_prior.cu", "bpl_token_sample.cu", "bpl_type_sample.cu", "bpl_type_prior.cu")]
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-Xcompiler", "-fPIC,-ffp-contract=off",
                       "-shared", "-DBPL_TOKEN_TIMING"] + src + ["-o", so])
import torch
from pybpl_b200 import _lib
_lib.SO_PATH = so
from pybpl_b200 import ops, workload
dev = torch.device("cuda", 0)
w = workload.synth_tokens(P, seed=2024 if BWD else 1234, sigma_mode=10.0 if BWD else "uniform")
if BWD:
    os.environ["BPL_BWD_TEAMS"] = "1"
p0, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), dev))
imgs = ops.pack_images((torch.from_numpy(np.random.default_rng(99).random((105, 105))).to(dev) < p0[0]).to(torch.uint8))
buf = torch.zeros(P + 148 * 64, device=dev)
os.environ["BPL_PT_BUF"] = str(buf.data_ptr())
b = workload.to_device(w, dev, requires_grad=BWD)


def rate(order):
    b._order = b._order_cost = None if order is None else torch.as_tensor(order, dtype=torch.int32, device=dev)
    b.use_order = order is not None
    with torch.set_grad_enabled(BWD):
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
    return P / best / 1e3


tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
S = np.diff(sso[tso])
print(f"index order            {rate(None):7.3f} M particles/s")
print(f"by sub-stroke count    {rate(np.argsort(-S, kind='stable')):7.3f}")
bb = workload.to_device(w, dev, requires_grad=True)
bb.c_struct(bb.ctrl.detach(), bb.invscale.detach(), bb.position.detach(), None, bb.eps.detach(), bb.sigma.detach(), cost_order=True)
print(f"by the cost key        {rate(bb._order_cost.cpu().numpy()):7.3f}")
buf.zero_(); rate(np.argsort(-S, kind='stable'))
t = buf[:P].cpu().numpy().astype(np.float64)
print(f"by measured cycles     {rate(np.argsort(-t, kind='stable')):7.3f}   (upper bound for any key; corr(cycles, S) {np.corrcoef(t, S)[0, 1]:.2f})")
