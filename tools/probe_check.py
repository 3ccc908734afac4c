"""Diagnostic: time the FP32 FFMA probe several times (and with different launch shapes)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pybpl_b200 import _lib
dev = torch.device("cuda:0")
scratch = torch.zeros(1, device=dev)
L = _lib.lib()
def run(iters, blocks):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(L.bpl_fp32_probe(ctypes.c_void_p(scratch.data_ptr()), iters, blocks, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    return blocks * 1024.0 * iters * 16 / (ms * 1e-3) / 1e12, ms
for iters, blocks in [(4096, 148*8)]*4 + [(16384, 148*8)]*3 + [(16384, 148*2)]*2 + [(65536, 148*2)]*2:
    print(iters, blocks, "%.2f TFLOP/s  %.3f ms" % run(iters, blocks))
