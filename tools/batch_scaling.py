#!/usr/bin/env python
"""Forward throughput vs batch size (device-resident inputs), with SM clock samples: is the 4096-token bench step a
burst figure?"""
import os, subprocess, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import ops, workload
dev = torch.device("cuda", 0)
imgs = ops.pack_images((torch.rand(105, 105, device=dev) < 0.04).to(torch.uint8))
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
for P in (4096, 32768, 262144):
    b = workload.to_device(workload.synth_tokens(P, seed=1234, sigma_mode="uniform"), dev)
    for _ in range(3):
        ops.score_tokens(b, imgs)
    torch.cuda.synchronize()
    reps = max(3, 1048576 // P)
    t0 = time.perf_counter()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.score_tokens(b, imgs)
    e.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = a.elapsed_time(e)
    clk = [float(r.split(",")[0]) for t, r in rows if t0 <= t <= t1 and r]
    print(f"P={P:7d} x{reps:4d} back to back: {P * reps / ms / 1e3:6.2f} M particles/s; SM clock during: "
          f"{(np.median(clk) if clk else float('nan')):.0f} MHz ({len(clk)} samples)")
proc.terminate()
