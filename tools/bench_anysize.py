"""Throughput of the run-time-geometry kernels (bpl_render_strokes_any[_grad]) next to the fused 105 x 105 path on the same
raw trajectories: 4 096 characters of the bench workload as motor-space strokes.  GPU box only; prints M characters/s."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, workload

DEV = torch.device("cuda", 0)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def main(P=4096):
    w = workload.synth_tokens(P, seed=1234, sigma_mode="uniform")
    tb = workload.to_device(w, DEV)
    with torch.no_grad():
        sb = ops.tokens_as_strokes(tb)
    img = (torch.rand(105, 105, device=DEV) < 0.1).float()
    packed = ops.pack_images(img)
    std = ops.make_params(None)
    gen = ops.make_params(None); gen.imsize = 0; gen.hw = (105, 105)
    imgs = img[None].contiguous()
    with torch.no_grad():
        t_f = timed(lambda: ops.render_strokes(sb, packed, want_pimg=False, ps=std))
        t_a = timed(lambda: ops._RenderStrokesAny.apply(sb.pts, sb.eps, sb.sigma, sb, imgs, None, False, gen))
    print(f"forward, 105 x 105, {P} characters as raw strokes: fused dense kernel {P / t_f / 1e6:.2f} M/s, "
          f"run-time-geometry kernel {P / t_a / 1e6:.2f} M/s")

    def bwd(any_size):
        pts = sb.pts.detach().requires_grad_(True)
        b = ops.StrokeBatch(sb.tok_sub_off, sb.sub_pt_off, pts, sb.eps, sb.sigma)
        if any_size:
            _, ll, _ = ops._RenderStrokesAny.apply(b.pts, b.eps, b.sigma, b, imgs, None, False, gen)
        else:
            _, ll, _ = ops.render_strokes(b, packed, want_pimg=False, ps=std)
        ll.sum().backward()
    t_f = timed(lambda: bwd(False), 10); t_a = timed(lambda: bwd(True), 10)
    print(f"forward + backward (two launches each): fused {P / t_f / 1e6:.2f} M/s, run-time geometry {P / t_a / 1e6:.2f} M/s")
    for (H, W) in ((64, 64), (210, 210)):
        p = ops.make_params(None); p.imsize = 0; p.hw = (H, W)
        im = (torch.rand(1, H, W, device=DEV) < 0.1).float()
        sc = torch.tensor([W / 105.0, H / 105.0], device=DEV)
        b = ops.StrokeBatch(sb.tok_sub_off, sb.sub_pt_off, (sb.pts * sc).contiguous(), sb.eps, sb.sigma)
        with torch.no_grad():
            t = timed(lambda: ops._RenderStrokesAny.apply(b.pts, b.eps, b.sigma, b, im, None, False, p))
        print(f"forward, {H} x {W}: {P / t / 1e6:.2f} M characters/s ({P * H * W / t / 1e9:.1f} G pixels/s)")


if __name__ == "__main__":
    main()
