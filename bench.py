#!/usr/bin/env python
"""bench.py -- rendered+scored 105x105 particles/s on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (spline eval -> rasterise -> broaden/blur -> Bernoulli
log-likelihood, one fused kernel) over one batch of 4,096 synthetic tokens scored against one 105x105
binary image (BASELINE.json configs[1]); under torchrun each rank steps its own batch (weak scaling) and the
ranks' (max, sum-exp) partials are combined with one NCCL all-gather per step (the log-sum-exp of the
particle sweep).  One JSON line on stdout (rank 0).

`--impl reference` times the CPU restatement of the reference (oracle/, C + OpenMP on all host cores;
the reference itself is Python and does not travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rendered+scored 105x105 particles/sec (fwd)"
UNIT = "particles/s"
BATCH = 4096
SEED = 1234
# dram__bytes_read.sum + dram__bytes_write.sum of one bpl_forward1s_kernel launch on this workload, from the
# ncu --set full capture of this command summarised in profiles/r01_fwd1s_final.raw.txt (1.831168 MB + 3.84 KB)
NCU_DRAM_BYTES_PER_LAUNCH = 1835008


def flops_fwd(S):      # algorithmic work per particle, separable minimal algorithm
    return 1.51e6 + 1.0e4 * S


def flops_fwd_bwd(S):
    return 4.7e6 + 2.5e4 * S


def hbm_bytes(S, K):   # control points in, score out
    return 44.0 * S + 8.0 * K + 12.0 + 5.0


def workload_config(n_gpus, extra=None):
    cfg = {"workload": "configs[1]: 4096 synthetic tokens (1-5 strokes, ragged sub-strokes, blur_sigma~U(0.5,16), "
                       "eps=1e-4) rendered+scored against one 105x105 binary image per GPU",
           "batch_per_gpu": BATCH, "seed": SEED, "parallelism": f"particle-sharded x{n_gpus}",
           "l2": "flushed between timed steps (256 MiB write)"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ---------------------------------------------------------------------------------------------
def cpu_target_image(w):
    from oracle import Oracle
    from pybpl_b200 import workload
    C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
    import oracle
    return oracle.target_image_for(w, Oracle(np.float32), C), C


def cpu_baseline(w, img, C, n_sample, repeats=1):
    from oracle import Oracle
    from pybpl_b200 import workload
    o = Oracle(np.float32)
    sub = workload.permute(w, np.arange(min(n_sample, w["P"])))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.token_batch(C, sub["tok_stroke_off"], sub["stroke_sub_off"], sub["ctrl"], sub["invscale"], sub["position"],
                      sub["eps"], sub["sigma"], img[None])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sub["P"] / best, sub["P"], best


def cpu_baseline_bounded(w, img, C, budget_s=12.0):
    """The whole step repeated for about `budget_s` seconds of wall time on all host cores; mean rate over the passes
    after the first (which warms the OpenMP pool and the caches)."""
    cpu_baseline(w, img, C, w["P"])
    t_tot, n_tot, passes = 0.0, 0, 0
    while t_tot < budget_s and passes < 64:
        _, n, dt = cpu_baseline(w, img, C, w["P"])
        t_tot += dt; n_tot += n; passes += 1
    return n_tot / t_tot, n_tot, t_tot, passes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pybpl_b200 import workload
    w = workload.synth_tokens(BATCH, seed=SEED, sigma_mode="uniform")
    img, C = cpu_target_image(w)
    import oracle
    cores = oracle.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1 to its workers
    n_sample = 1024
    for _ in range(max(args.warmup, 1)):
        cpu_baseline(w, img, C, 64)
    t_tot = 0.0
    n_tot = 0
    for _ in range(args.steps):
        v, n, dt = cpu_baseline(w, img, C, n_sample)
        t_tot += dt; n_tot += n
    value = n_tot / t_tot
    sample = f"{n_sample} of the 4096 tokens per step, C oracle port with OpenMP over tokens"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pybpl_b200 import _lib, ops, workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    # per-rank batch (weak scaling): same generator, rank-offset seed; one shared target image
    w0 = workload.synth_tokens(BATCH, seed=SEED, sigma_mode="uniform")
    w = w0 if rank == 0 else workload.synth_tokens(BATCH, seed=SEED + rank, sigma_mode="uniform")
    tso = w["tok_stroke_off"].astype(np.int64)
    Sbar = float(np.diff(w["stroke_sub_off"].astype(np.int64)[tso]).mean())
    Kbar = float(np.diff(tso).mean())
    # the target: a Bernoulli sample of token 0's probability map, rendered by the CUDA path itself (the oracle is
    # only used by the cpu_baseline leg below); the uniform noise comes from a seeded numpy generator as in
    # oracle.target_image_for, so the CPU arm sees the same image up to pixels within 1e-6 of the noise
    if rank == 0:
        with torch.no_grad():
            p0, _ = ops.render_tokens(workload.to_device(workload.permute(w0, np.array([0])), dev))
        u = np.random.default_rng(99).random((105, 105))
        img_np = (u < p0[0].cpu().numpy()).astype(np.uint8)
    else:
        img_np = np.zeros((105, 105), np.uint8)
    img_t = torch.from_numpy(img_np).to(dev)
    if world > 1:
        dist.broadcast(img_t, 0)
    ps = ops.make_params(None)

    batch = workload.to_device(w, dev)
    imgs = ops.pack_images(img_t)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # The per-step collective (2 floats per rank) is issued asynchronously on NCCL's stream and overlaps the NEXT step's
    # kernel; a step waits (stream-side) for the previous step's gather, and finish_collectives() -- timed as part of the
    # run -- waits for the last one, so every step's exchange completes inside the timed region.
    gathered = [torch.empty((world * 2,), dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None
    coll = {"work": None, "i": 0}

    def exchange(part):
        if world > 1:
            if coll["work"] is not None:
                coll["work"].wait()
            coll["work"] = dist.all_gather_into_tensor(gathered[coll["i"] & 1], part, async_op=True)
            coll["i"] += 1

    def finish_collectives():
        if coll["work"] is not None:
            coll["work"].wait()
            coll["work"] = None

    def step_device():
        ll, off = ops.score_tokens(batch, imgs, ps=ps)
        part = ops.logsumexp_partial(ll)
        exchange(part)
        return ll, part

    # host-resident inputs for the end-to-end arm: one pinned staging buffer per step's inputs (token arrays + the
    # target image), one H2D copy, pack + render+score + log-sum-exp partial, one D2H copy of the scores; all
    # asynchronous on the current stream, the timed region ends with a synchronize
    staged = ops.HostStagedBatch(w, img_np, dev)
    host_ll = torch.empty(BATCH, dtype=torch.float32).pin_memory()
    h2d = staged.nbytes
    d2h = host_ll.numel() * 4

    # two device slots: the H2D copy of step i+1 is issued on a copy stream before step i's kernels, so copies and
    # kernels overlap; every step's copy still lies inside the timed region (the first step of a phase uploads its
    # own input, the last one prefetches nothing)
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "n": 0}

    def upload_async(slot):
        copy_stream.wait_event(done[slot])              # the kernels that last read this slot have finished
        with torch.cuda.stream(copy_stream):
            staged.upload(slot)
            ready[slot].record(copy_stream)

    def begin_e2e(nsteps):
        torch.cuda.synchronize()
        state["i"], state["n"] = 0, nsteps
        for ev in done:
            ev.record()

    def step_e2e():
        i = state["i"]; slot = i & 1
        if i == 0:
            upload_async(0)                             # the first step of a phase copies its own input
        if i + 1 < state["n"]:
            upload_async(slot ^ 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[slot])
        b, im = staged.views(slot)
        packed = ops.pack_images(im, validate=False)
        ll, _ = ops.score_tokens(b, packed, ps=ps)
        part = ops.logsumexp_partial(ll)
        exchange(part)
        host_ll.copy_(ll, non_blocking=True)
        done[slot].record(cur)
        state["i"] = i + 1
        return host_ll

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, flush_l2=True, begin=None):
        if begin:
            begin(warmup)
        for _ in range(warmup):
            fn()
        finish_collectives()
        barrier()
        if begin:
            begin(steps)
        evs = []
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            if flush_l2:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        finish_collectives()                 # the last step's exchange
        b.record()
        evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, _lib.launch_count() - l0

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, wall, launches = timed(step_device, args.steps, max(args.warmup, 3))
    ms_e2e, wall_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3), begin=begin_e2e)
    clocks = sampler.stop() if sampler else None

    # kernel-only duration of the dominant kernel (events directly around the C-ABI call on this stream)
    ll_buf = torch.empty(BATCH, dtype=torch.float32, device=dev)
    off_buf = torch.empty(BATCH, dtype=torch.uint8, device=dev)

    def kernel_only():
        ll, _ = ops.score_tokens(batch, imgs, ps=ps)
        return ll
    kms = []
    for i in range(13):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernel_only(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            kms.append(a.elapsed_time(b))
    k_ms = float(np.mean(kms))

    # fwd+bwd (config 3 shape: 1024 parses, gradients to ctrl/invscale/position/eps/sigma)
    wb = workload.permute(w, np.arange(1024))
    bb = workload.to_device(wb, dev, requires_grad=True)
    Sb = float(np.diff(wb["stroke_sub_off"].astype(np.int64)[wb["tok_stroke_off"].astype(np.int64)]).mean())

    def step_bwd():
        ll, _ = ops.score_tokens(bb, imgs, ps=ps)
        return ll
    ms_bwd, _, _ = timed(step_bwd, max(args.steps // 4, 5), 3)
    n_bwd = max(args.steps // 4, 5)

    # live FP32 FMA peak (roofline denominator)
    scratch = torch.zeros(1, device=dev)
    import ctypes
    iters, blocks = 4096, 148 * 8
    for _ in range(2):
        _lib.check(_lib.lib().bpl_fp32_probe(ctypes.c_void_p(scratch.data_ptr()), iters, blocks,
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(_lib.lib().bpl_fp32_probe(ctypes.c_void_p(scratch.data_ptr()), iters, blocks,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    b.record(); torch.cuda.synchronize()
    fp32_peak = (blocks * 1024.0 * iters * 8 * 2) / (a.elapsed_time(b) * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    value = world * BATCH * args.steps / (ms_total * 1e-3)
    e2e = world * BATCH * args.steps / (ms_e2e * 1e-3)
    k_rate = BATCH / (k_ms * 1e-3)
    ach_tf = k_rate * flops_fwd(Sbar) / 1e12
    roofline = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "kernel": "bpl_forward1s_kernel<false>", "kernel_ms": k_ms,
                "flop_per_particle": flops_fwd(Sbar), "mean_substrokes": Sbar,
                "peak_source": "live FFMA probe (bpl_fp32_probe) in this run; MEASURED_PEAKS.json holds no CUDA-core FP32 "
                               "figure, and SURVEY.md 8(d) shows the path is FP32-issue/shared-memory bound, not HBM",
                "hbm": {"achieved": k_rate * hbm_bytes(Sbar, Kbar) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": k_rate * hbm_bytes(Sbar, Kbar) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1:
        import oracle
        cores = oracle.set_threads(cores)
        C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
        cv, cn, cdt, cpasses = cpu_baseline_bounded(w0, img_np, C)
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the step's {BATCH} tokens x {cpasses} passes = {cn} particles, C oracle port (fp32), OpenMP over tokens, "
                         f"{cdt:.1f} s wall on {cores} cores"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, {"mean_strokes": Kbar, "mean_substrokes": Sbar}),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks, "wall_s": wall,
            "fwd_bwd": {"value": world * 1024 * n_bwd / (ms_bwd * 1e-3), "unit": UNIT, "batch_per_gpu": 1024,
                        "ms_per_step": ms_bwd / n_bwd,
                        "achieved_tflops": (1024 * n_bwd / (ms_bwd * 1e-3)) * flops_fwd_bwd(Sb) / 1e12}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """Everything any library writes to file descriptor 1 while the bench runs (NCCL prints its version line there) goes
    to stderr instead; emit() writes the one JSON line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20         # bounded: each step is ~1 s of all-core CPU work
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
