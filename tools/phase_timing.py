#!/usr/bin/env python
"""Per-phase cycle breakdown of bpl_forward2_kernel.  Builds a -DBPL_PHASE_TIMING copy of the library into /tmp,
runs the bench workload once and prints the share of each phase (thread 0's clock64 between barriers).
Caveat: warp 0 has the lowest issue priority, so right after a barrier it resumes late and a few thousand cycles of
the following phase are booked on the preceding mark; use the split for proportions, not for exact boundaries."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = "/tmp/libbpl_pt.so"
src = [os.path.join(ROOT, "pybpl_b200", "csrc", f) for f in ("bpl_render.cu", "bpl_splines.cu", "bpl_token_prior.cu", "bpl_token_sample.cu", "bpl_type_sample.cu", "bpl_type_prior.cu")]
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DBPL_PHASE_TIMING",
                       "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"] + src + ["-o", so])
from pybpl_b200 import _lib
_lib.SO_PATH = so
from pybpl_b200 import ops, workload
dev = torch.device("cuda", 0)
w = workload.synth_tokens(4096, seed=1234, sigma_mode="uniform")
batch = workload.to_device(w, dev)
imgs = ops.pack_images((torch.rand(105, 105, device=dev) < 0.05).to(torch.uint8))
ps = ops.make_params(None)
P = batch.P
ll = torch.empty(P, device=dev); off = torch.empty(P, dtype=torch.uint8, device=dev)
acc = torch.zeros(296 * 16, device=dev)
b = batch.c_struct(batch.ctrl, batch.invscale, batch.position, None, batch.eps, batch.sigma)
t = _lib.Targets(imgs.bits.data_ptr(), None, 1)
out = _lib.FwdOut(ll.data_ptr(), off.data_ptr(), None)
L = _lib.lib()
# g_pts smuggles the counter buffer in: use the grad struct through score_tokens? the plain entry has no grads,
# so call the kernel twice through bpl_score_tokens_grad's argument filler is not possible; instead the timing
# build reads KArgs.g_pts, which bpl_score_tokens leaves NULL -> expose via env var pointer
os.environ["BPL_PT_BUF"] = str(acc.data_ptr())
for _ in range(3):
    _lib.check(L.bpl_score_tokens(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
a = acc.view(296, 16)[:, :11].cpu().numpy()
names = ["setup(zero,consts,geom)", "chain offsets", "point: fix check/pass", "broaden", "blur", "tail+reduce", "point: warp 0 own work", "point: wait for slowest warp", "  warp_points A (eval)", "  warp_points B (scan+carry)", "  warp_points C (walk+splat)"]
tot_main = a[:, :8].sum()
tot = a[:, :8].sum()
for n, v in zip(names, a.sum(0)):
    print(f"{n:26s} {100 * v / tot:5.1f} %   {v / 296 / (4096 / 2 / 296):9.0f} cycles per pair")
print("total cycles per pair per CTA", tot / 296 / (4096 / 2 / 296))

# ---- backward kernel (1024 parses, 1 CTA/SM) ----
acc.zero_()
wb = workload.permute(w, np.arange(1024))
bb = workload.to_device(wb, dev, requires_grad=True)
for _ in range(3):
    acc.zero_()
    llb, _ = ops.score_tokens(bb, imgs)
torch.cuda.synchronize()
a = acc.view(296, 16)[:148, :8].cpu().numpy()
names = ["setup + chain offsets", "forward splat (+fix)", "forward broaden", "blur fwd + tail + adjoint blur", "adjoint broaden", "scalar reductions", "add_stroke backward", "chain backward + outputs"]
tot = a.sum()
print("backward kernel:")
for n, v in zip(names, a.sum(0)):
    print(f"{n:34s} {100 * v / tot:5.1f} %   {v / 148 / (1024 / 148):9.0f} cycles per token")
print("total cycles per token per CTA", tot / 148 / (1024 / 148))
