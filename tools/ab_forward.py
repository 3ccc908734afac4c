#!/usr/bin/env python
"""A/B of the forward scoring kernels on the bench workload (4096 tokens, device-resident, back to back):
    python tools/ab_forward.py TAG          # kernel choice comes from the environment (BPL_FORCE_SINGLE, BPL_FWD_SHAPE=small|large)
writes gpurun_out/ab_TAG.npy (the scores) and prints particles/s; with two or more earlier tags it also prints the
largest relative difference between their scores."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("BPL_LIB"):          # a variant build (tools/variant_bench.py)
    from pybpl_b200 import _lib
    _lib.SO_PATH = os.environ["BPL_LIB"]
from pybpl_b200 import ops, workload
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
w = workload.synth_tokens(P, seed=1234, sigma_mode="uniform")
b = workload.to_device(w, dev)
pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), dev))
torch.manual_seed(0)
img = (torch.rand_like(pimg[0]) < pimg[0]).to(torch.uint8)
imgs = ops.pack_images(img)
for _ in range(5):
    ll, off = ops.score_tokens(b, imgs)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ll, off = ops.score_tokens(b, imgs)
    e.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(e) / 20)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"ab_{tag}.npy"), ll.cpu().numpy())
print(f"{tag}: {P / best / 1e3:.3f} M particles/s ({best * 1e3:.1f} us per launch)")
others = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "ab_*.npy")))
ref = None
for f in others:
    v = np.load(f)
    if v.shape != (P,):
        continue
    if ref is None:
        ref, rname = v, os.path.basename(f)
        continue
    rel = np.abs(v - ref) / np.abs(ref)
    print(f"  {os.path.basename(f)} vs {rname}: max rel {rel.max():.2e}, median {np.median(rel):.2e}")

# ---- fused forward+backward on 1024 parses (config 3 step) ----
wb = workload.permute(w, np.arange(min(P, 1024)))
bb = workload.to_device(wb, dev, requires_grad=True)
def step():
    llb, _ = ops.score_tokens(bb, imgs)
    return llb
for _ in range(3):
    step()
torch.cuda.synchronize()
bestb = 1e9
for rep in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        step()
    e.record(); torch.cuda.synchronize()
    bestb = min(bestb, a.elapsed_time(e) / 20)
print(f"{tag}: fwd+bwd kernel {bb.P / bestb / 1e3:.3f} M particles/s ({bestb * 1e3:.1f} us per launch)")
# gradients of this build (compared across tags like the scores)
leaves = [bb.ctrl, bb.invscale, bb.position, bb.eps, bb.sigma]
gr = torch.autograd.grad(step().sum(), leaves)
np.savez(os.path.join(ROOT, "gpurun_out", f"abg_{tag}.npz"), **{k: g.cpu().numpy() for k, g in zip(("ctrl", "invscale", "position", "eps", "sigma"), gr)})
refg = None
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "abg_*.npz"))):
    v = dict(np.load(f))
    if v["eps"].shape != (bb.P,):
        continue
    if refg is None:
        refg, rname = v, os.path.basename(f)
        continue
    print(f"  {os.path.basename(f)} vs {rname}: " + ", ".join(f"{k} {np.abs(v[k] - refg[k]).max() / np.abs(refg[k]).max():.1e}" for k in v))
