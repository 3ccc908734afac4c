#!/usr/bin/env python
"""Per-phase cycle breakdown of bpl_forward1s_kernel : thread 0 of every CTA accumulates clock64() deltas per phase (-DBPL_PHASE_TIMING)."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = "/tmp/libbpl_pt.so"
src = [os.path.join(ROOT, "pybpl_b200", "csrc", f) for f in ("bpl_render.cu", "bpl_splines.cu", "bpl_token_prior.cu", "bpl_token_sample.cu", "bpl_type_sample.cu", "bpl_type_prior.cu")]
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DBPL_PHASE_TIMING",
                       "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"] + sys.argv[1:] + src + ["-o", so])
from pybpl_b200 import _lib
_lib.SO_PATH = so
from pybpl_b200 import ops, workload
dev = torch.device("cuda", 0)
w = workload.synth_tokens(4096, seed=1234, sigma_mode="uniform")
batch = workload.to_device(w, dev)
imgs = ops.pack_images((torch.rand(105, 105, device=dev) < 0.05).to(torch.uint8))
NCTA = 148 * 3
acc = torch.zeros(NCTA * 16, device=dev)
os.environ["BPL_PT_BUF"] = str(acc.data_ptr())
for _ in range(3):
    ops.score_tokens(batch, imgs)
torch.cuda.synchronize()
a = acc.view(NCTA, 16)[:, :8].cpu().numpy()
names = ["setup(geom,bits)", "chain offsets+consts", "point: fix check/pass", "broaden", "blur", "tail+reduce",
         "point: warp 0 own work", "point: wait for slowest warp"]
tot = a.sum()
per = 4096 / NCTA
for n, v in zip(names, a.sum(0)):
    print(f"{n:30s} {100 * v / tot:5.1f} %   {v / NCTA / per:9.0f} cycles per token")
print("total cycles per token per CTA", tot / NCTA / per)
