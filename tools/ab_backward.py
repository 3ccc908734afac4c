#!/usr/bin/env python
"""A/B of the fused forward+backward kernels on the bench's backward workload (seed 2024, blur_sigma 10; device-resident,
back to back):   python tools/ab_backward.py TAG [P ...]     (kernel choice from the environment: BPL_BWD_CANVAS=1)
writes gpurun_out/abb_TAG.npz (scores + gradients of the first size) and compares with earlier tags."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("BPL_LIB"):          # a variant build (tools/variant_bench.py)
    from pybpl_b200 import _lib
    _lib.SO_PATH = os.environ["BPL_LIB"]
from pybpl_b200 import ops, workload
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
sizes = [int(x) for x in sys.argv[2:]] or [1024]
reps = int(os.environ.get("ABB_REPS", "5"))
dev = torch.device("cuda", 0)
out = []
for n, P in enumerate(sizes):
    w = workload.synth_tokens(P, seed=2024, sigma_mode=10.0)
    if n == 0:
        p0, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), dev))
        img = (torch.from_numpy(np.random.default_rng(99).random((105, 105))).to(dev) < p0[0]).to(torch.uint8)
        imgs = ops.pack_images(img)
    bb = workload.to_device(w, dev, requires_grad=True)
    step = lambda: ops.score_tokens(bb, imgs)[0]
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            step()
        e.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e) / 10)
    out.append(f"{P}: {P / best / 1e3:.3f} M/s ({best * 1e3:.0f} us)")
    if n == 0:
        leaves = [bb.ctrl, bb.invscale, bb.position, bb.eps, bb.sigma]
        ll = step()
        gr = torch.autograd.grad(ll.sum(), leaves)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez(os.path.join(ROOT, "gpurun_out", f"abb_{tag}.npz"), ll=ll.detach().cpu().numpy(),
                 **{k: g.cpu().numpy() for k, g in zip(("ctrl", "invscale", "position", "eps", "sigma"), gr)})
print(f"{tag}: fwd+bwd " + "; ".join(out))
ref = None
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "abb_*.npz"))):
    v = dict(np.load(f))
    if v["eps"].shape != (sizes[0],):
        continue
    if ref is None:
        ref, rname = v, os.path.basename(f)
        continue
    print(f"  {os.path.basename(f)} vs {rname}: " + ", ".join(f"{k} {np.abs(v[k] - ref[k]).max() / np.abs(ref[k]).max():.1e}" for k in v))
