#!/usr/bin/env python
"""bench.py -- rendered+scored 105x105 particles/s on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one particle sweep, BASELINE.json configs[3]: 1,000,000 synthetic token particles scored against one
105x105 binary image, the particles sharded over the N ranks of a torchrun job (STRONG scaling: 1M / N each), every rank
running ONE fused launch of the hot path over its share (spline eval -> rasterise -> broaden/blur -> Bernoulli
log-likelihood -> the rank's log-sum-exp partial), then one NCCL all-gather of 2 floats per rank and a one-warp combine
into the global log-sum-exp.  K steps run back to back inside ONE timed region (CUDA events on the launching stream
and wall clock; nothing is untimed between steps); the collective is on the critical path of every step.

`value`   : particles/s with the particles resident in HBM when the timed region starts
`e2e`     : the same sweep from HOST numpy arrays through the public API: host-side ragged packing of every chunk into
            pinned memory, H2D, fused kernel, D2H of the scores, wall clock
`config1` : configs[1], the 4,096-token step on one GPU (kernel time, roofline)       \\
`fwd_bwd` : configs[2] shape, 1,024 parses/step through the fused forward+backward     > secondary objects
`config2` : configs[2], 1,024 parses x 100 Adam steps, total seconds                   /
`config4` : configs[4], cross-scoring of 40,000 tokens x 20 images, tokens sharded over the ranks, blocks all-gathered
`cpu_baseline` = the reference's own torch code (unmodified pybpl from baseline/_ref, one single-thread process per
core; `cpu_baseline_torch` repeats it) and `cpu_baseline_port` (C oracle port, OpenMP over all cores) are timed on rank 0
at N=1 on a bounded sample of the same particles.

`--impl reference` times the CPU arm alone (same config; rank 0 only under torchrun): its `value` is the reference's
own torch code when baseline/_ref is present (kind "reference"), else the port.
"""
import argparse
import json
import math
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
L2_BYTES = 126 * 1024 * 1024
E2E_CHUNK = 65536
CFG1_BATCH = 4096
CFG1_SEED = 1234
BWD_BATCH = 1024
CPU_SAMPLE = 8192            # particles per step of the CPU arm (first tokens of sweep block 0)
# dram__bytes_read.sum + dram__bytes_write.sum of one scoring launch, from the ncu --set full capture summarised under
# profiles/ (filled in from the round's capture; None until one exists for the current kernel)
NCU_SMEM_WAVEFRONTS_PER_PARTICLE = 5848.0   # profiles/r02_fwd1s_sweep_final4.raw.txt: l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
                                            # / 1,000,000 particles; one wavefront moves up to 128 bytes
NCU_DRAM_BYTES_PER_PARTICLE = 291.24        # profiles/r02_fwd1s_sweep_final4.raw.txt: (280.92 MB read + 10.32 MB written) / 1,000,000 particles


def flops_fwd(S):      # algorithmic work per particle, separable minimal algorithm
    return 1.51e6 + 1.0e4 * S


def flops_fwd_bwd(S):
    return 4.7e6 + 2.5e4 * S


def hbm_bytes(S, K):   # control points in, score out
    return 44.0 * S + 8.0 * K + 12.0 + 5.0


def mean_sk(w):
    tso = w["tok_stroke_off"].astype(np.int64)
    return float(np.diff(w["stroke_sub_off"].astype(np.int64)[tso]).mean()), float(np.diff(tso).mean())


def sweep_config(n_gpus):
    """The `config` object: identical for the GPU arm and `--impl reference`."""
    from pybpl_b200 import workload
    per_rank_mb = 277.0 / n_gpus
    copies = max(1, math.ceil(2.0 * L2_BYTES / 1e6 / per_rank_mb))
    return {"workload": "configs[3]: particle sweep, 1,000,000 synthetic token particles (1-5 strokes, ragged sub-strokes, "
                        "blur_sigma~U(0.5,16), eps=1e-4) x one 105x105 binary target, sharded across the ranks; per step and "
                        "rank one fused render+score+log-sum-exp launch, then one all-gather of 2 floats per rank",
            "particles": workload.SWEEP_PARTICLES, "blocks": workload.SWEEP_BLOCKS, "seed": workload.SWEEP_SEED,
            "parallelism": f"particle-sharded x{n_gpus} (strong scaling)",
            "l2": f"inputs larger than L2: every rank cycles through {copies} resident cop{'y' if copies == 1 else 'ies'} of its "
                  f"{per_rank_mb:.0f} MB of particles (> 2 x 126 MB in total), a different one each step"}


# ---------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------
def basis():
    return np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))


def oracle_rate(w, img, C, grad=False, budget_s=12.0, max_passes=64):
    """The C oracle port (OpenMP over tokens, all host cores) repeated over `w` for about budget_s seconds."""
    from oracle import Oracle
    o = Oracle(np.float32)
    args = (C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"], w["eps"], w["sigma"], img[None])
    o.token_batch(*args, grad=grad)                     # warms the OpenMP pool and the caches
    t_tot, n_tot, passes = 0.0, 0, 0
    while t_tot < budget_s and passes < max_passes:
        t0 = time.perf_counter()
        o.token_batch(*args, grad=grad)
        t_tot += time.perf_counter() - t0; n_tot += w["P"]; passes += 1
    return n_tot / t_tot, n_tot, t_tot, passes


def torch_reference_rate(w, img, grad=False, budget_s=10.0):
    """The reference's own torch code (baseline/_ref), one single-thread process per core; None when it is not installed."""
    from baseline import ref_loader, torch_ref
    if not ref_loader.available():
        return None
    cores = os.cpu_count() or 1
    try:
        rate, n, procs, _ = torch_ref.score_rate(w, img, n_procs=cores, per_proc=max(8, w["P"] // cores), grad=grad,
                                                 budget_s=budget_s)
    except Exception as e:                               # a reported baseline must not take the bench down
        return {"unavailable": str(e)[:200]}
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"{n} particles of sweep block 0 scored by the unmodified pybpl CharacterImageDist.score_image"
                      f"{' + .backward() (rendering.py:97 patched in memory)' if grad else ''} from baseline/_ref, "
                      f"{procs} single-thread processes (torch.set_num_threads(1)) side by side, <= {budget_s:.0f} s each"}


def cpu_sample():
    from pybpl_b200 import workload
    import oracle
    from oracle import Oracle
    w0 = workload.sweep_block(0)
    C = basis()
    img = oracle.target_image_for(w0, Oracle(np.float32), C)
    return workload.permute(w0, np.arange(CPU_SAMPLE)), img, C


def run_reference(args):
    """The CPU arm: the reference's OWN torch code (unmodified pybpl from baseline/_ref: CharacterImageDist.score_image,
    one single-thread process per host core) on a bounded sample of the sweep's particles; the C
    oracle port (OpenMP, all cores) is timed beside it and stands in when baseline/_ref is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = oracle.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1 to its workers
    w, img, C = cpu_sample()
    from oracle import Oracle
    o = Oracle(np.float32)
    call = lambda: o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"],
                                 w["eps"], w["sigma"], img[None])
    for _ in range(max(args.warmup, 1)):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    t_tot = time.perf_counter() - t0
    port_value = w["P"] * args.steps / t_tot
    port = {"value": port_value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{CPU_SAMPLE} particles of sweep block 0 per step (of the 1,000,000) x {args.steps} steps, C oracle port "
                      f"(fp32) with OpenMP over tokens on {cores} cores"}
    # each step of the torch arm = CPU_SAMPLE particles shared by the worker processes; all steps in one set of workers
    # (a worker's start-up -- importing torch and the reference -- is outside its own clock)
    torch_ref = torch_reference_rate(w, img, budget_s=max(4.0, min(args.cpu_budget, 1.5 * args.steps)))
    if torch_ref and "value" in torch_ref:
        value, baseline = torch_ref["value"], torch_ref
    else:
        value, baseline = port_value, port
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * CPU_SAMPLE / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": sweep_config(args.gpus),
            "cpu_baseline": baseline, "cpu_baseline_port": port,
            "cpu_baseline_torch": torch_ref,
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
        self.marks = []
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
            self.rows.append((time.perf_counter(), ln.strip()))

    def mark(self):
        self.marks.append(time.perf_counter())

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
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (-1e30, 1e30)
        for ts, r in self.rows:
            if not (lo <= ts <= hi + 0.05):
                continue                                  # only samples taken during the timed regions
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
    import ctypes
    import torch
    import torch.distributed as dist
    from pybpl_b200 import _lib, ops, sweep, workload

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
    steps, warmup = args.steps, max(args.warmup, 3)
    ps = ops.make_params(None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the rank's share of the 1M particles (host numpy; then resident on the device) ----
    w = workload.sweep_shard(rank, world)
    P_rank = int(w["P"])
    P_total = workload.SWEEP_PARTICLES
    Sbar, Kbar = mean_sk(w)
    # target = Bernoulli sample of particle 0's probability map, rendered by the CUDA path (rank 0) and broadcast; the
    # CPU arms render the same particle with the oracle and draw from the same seeded uniform field
    if rank == 0:
        with torch.no_grad():
            p0, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.array([0])), dev))
        img_np = (np.random.default_rng(99).random((105, 105)) < p0[0].cpu().numpy()).astype(np.uint8)
    else:
        img_np = np.zeros((105, 105), np.uint8)
    img_t = torch.from_numpy(img_np).to(dev)
    if world > 1:
        dist.broadcast(img_t, 0)
        img_np = img_t.cpu().numpy()
    imgs = ops.pack_images(img_t)
    in_bytes = sum(v.nbytes for v in w.values() if isinstance(v, np.ndarray))
    n_copies = max(1, math.ceil(2.0 * L2_BYTES / max(in_bytes, 1)))
    copies = [workload.to_device(w, dev) for _ in range(n_copies)]
    sw = sweep.Sweep(copies[0], imgs, want_ll=True, ps=ps)
    counter = {"i": 0}

    def step_device():
        counter["i"] += 1
        return sw.step(copies[counter["i"] % n_copies])

    def timed(fn, nsteps, nwarm, finish=None):
        """nwarm untimed steps, then nsteps back to back inside one region: CUDA events on the launching stream around the
        whole loop (max over ranks) and wall clock around the same loop; a barrier + synchronize on both sides."""
        for _ in range(nwarm):
            fn()
        if finish:
            finish()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        if sampler:
            sampler.mark()
        t0 = time.perf_counter()
        a.record()
        for _ in range(nsteps):
            fn()
        if finish:
            finish()
        b.record()
        torch.cuda.synchronize()
        wall_local = time.perf_counter() - t0
        barrier()
        if sampler:
            sampler.mark()
        return max_over_ranks(a.elapsed_time(b)), max_over_ranks(wall_local * 1e3), _lib.launch_count() - l0

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, wall_dev_ms, launches = timed(step_device, steps, warmup)
    lse_dev = float(sw.total.item())

    # ---- end to end: the same sweep from host numpy arrays ----
    # at least four chunks per step and rank, so that packing / copying chunk c+1 overlaps the kernel of chunk c
    e2e_chunk = int(min(E2E_CHUNK, max(16384, -(-P_rank // 4))))
    stager = ops.HostSweepStager(w, dev, chunk=e2e_chunk)
    nchunks = len(stager.chunks)
    host_ll = torch.empty(P_rank, dtype=torch.float32).pin_memory()
    host_total = torch.empty(1, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    state = ops.LseState(dev)
    gathered = torch.empty(2 * world, dtype=torch.float32, device=dev)
    e2e_n = {"c": 0}

    def step_e2e():
        cur = torch.cuda.current_stream()
        for c in range(nchunks):
            slot = e2e_n["c"] & 1
            e2e_n["c"] += 1
            copied[slot].synchronize()                   # the H2D copy that last read this pinned slot has completed
            stager.pack(c, slot)                         # host-side ragged packing into pinned memory
            copy_stream.wait_event(done[slot])           # ... and the kernel that last read the device slot
            with torch.cuda.stream(copy_stream):
                stager.upload(slot)
                copied[slot].record(copy_stream)
            cur.wait_event(copied[slot])
            b = stager.batch(c, slot)
            ll_c, _ = ops.score_tokens_lse(b, imgs, state, accumulate=c > 0, want_ll=True, ps=ps)
            c0, c1 = stager.chunks[c][0], stager.chunks[c][1]
            host_ll[c0:c1].copy_(ll_c, non_blocking=True)
            done[slot].record(cur)
        if world > 1:
            dist.all_gather_into_tensor(gathered, state.partial)
            total = ops.logsumexp_combine(gathered)
        else:
            total = ops.logsumexp_combine(state.partial)
        host_total.copy_(total, non_blocking=True)

    for ev in copied + done:
        ev.record()
    _, wall_e2e_ms, _ = timed(step_e2e, steps, warmup)
    lse_e2e = float(host_total.item())
    h2d = nchunks * stager.nbytes
    d2h = 4 * P_rank + 4

    # ---- kernel-only duration of the dominant kernel on this workload (events directly around the launch) ----
    kms = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bt = copies[i % n_copies]
        a.record(); ops.score_tokens_lse(bt, imgs, sw.state, want_ll=True, ps=ps); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            kms.append(a.elapsed_time(b))
    k_ms = float(np.mean(kms))

    # ---- fwd+bwd (configs[2] shape: 1024 parses per GPU and step; gradients to ctrl/invscale/position/eps/sigma) ----
    wb = workload.synth_tokens(BWD_BATCH, seed=2024 + rank, sigma_mode=10.0)
    Sb, _ = mean_sk(wb)
    bb = workload.to_device(wb, dev, requires_grad=True)
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)

    def step_bwd():
        ll, _ = ops.score_tokens(bb, imgs, ps=ps)
        return ll
    n_bwd = max(steps * 5, 50)
    ms_bwd, _, _ = timed(step_bwd, n_bwd, 5)
    kb = []
    for i in range(8):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step_bwd(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            kb.append(a.elapsed_time(b))
    kb_ms = float(np.mean(kb))
    # end to end: host numpy -> pinned -> H2D -> fused kernel -> D2H of the scores and all gradients
    hb = ops.HostSweepStager(wb, dev, chunk=BWD_BATCH, slots=1)
    g_host = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory() for k, v in
              dict(ll=wb["eps"], ctrl=wb["ctrl"], invscale=wb["invscale"], position=wb["position"], eps=wb["eps"],
                   sigma=wb["sigma"]).items()}

    def step_bwd_e2e():
        torch.cuda.current_stream().synchronize()          # the pinned slot and the gradient buffers are free again
        hb.pack(0, 0)
        hb.upload(0)
        b = hb.batch(0, 0)
        leaves = [b.ctrl, b.invscale, b.position, b.eps, b.sigma]
        for t in leaves:
            t.requires_grad_(True)
        ll, _ = ops.score_tokens(b, imgs, ps=ps)
        grads = torch.autograd.grad(ll.sum(), leaves)
        g_host["ll"].copy_(ll.detach(), non_blocking=True)
        for k, g in zip(("ctrl", "invscale", "position", "eps", "sigma"), grads):
            g_host[k].copy_(g, non_blocking=True)
    _, wall_bwd_e2e_ms, _ = timed(step_bwd_e2e, n_bwd, 5)
    bwd_h2d = hb.nbytes
    bwd_d2h = sum(t.numel() * 4 for t in g_host.values())

    # ---- configs[4]: cross-scoring, tokens sharded over the ranks, packed images replicated, blocks all-gathered ----
    X_TYPES, X_PER_TYPE, X_IMAGES = 20, 2000, 20
    xw = workload.synth_tokens(X_TYPES * X_PER_TYPE, seed=5, max_strokes=10, kappa="prior", sigma_mode="fixed")
    xlo, xhi = sweep.shard_bounds(xw["P"], rank, world)
    xown = np.arange(0, xw["P"], xw["P"] // X_IMAGES)[:X_IMAGES]
    with torch.no_grad():
        xp, _ = ops.render_tokens(workload.to_device(workload.permute(xw, xown), dev))      # seeded: identical on every rank
    ximgs = ops.pack_images((torch.from_numpy(np.random.default_rng(7).random((X_IMAGES, 105, 105))).to(dev) < xp).to(torch.uint8))
    xs = sweep.CrossScore(workload.to_device(workload.permute(xw, np.arange(xlo, xhi)), dev), ximgs, xw["P"])
    n_x = max(steps, 10)
    ms_x, wall_x_ms, _ = timed(xs.step, n_x, 3)
    x_sum = float(xs.ll.double().sum().item())
    config4 = {"workload": f"configs[4]: cross-scoring, {X_TYPES} types x {X_PER_TYPE} tokens x {X_IMAGES} target images (all pairs, "
                           f"max 10 strokes), tokens sharded over the ranks, [tokens, images] blocks all-gathered every pass",
               "value": xw["P"] * X_IMAGES * n_x / (wall_x_ms * 1e-3), "unit": "scored (token, image) pairs/s", "timing": "wall clock",
               "ms_per_pass": wall_x_ms / n_x, "tokens_per_s": xw["P"] * n_x / (wall_x_ms * 1e-3), "checksum": x_sum,
               "scaling": "strong"}

    # ---- live FP32 FMA peak (roofline denominator): best of 5 timed launches of the FFMA probe ----
    scratch = torch.zeros(1, device=dev)
    iters, blocks = 16384, 148 * 8
    probe = lambda: _lib.check(_lib.lib().bpl_fp32_probe(ctypes.c_void_p(scratch.data_ptr()), iters, blocks,
                                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for _ in range(2):
        probe()
    torch.cuda.synchronize()
    fp32_peak = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); probe(); b.record(); torch.cuda.synchronize()
        fp32_peak = max(fp32_peak, (blocks * 1024.0 * iters * 8 * 2) / (a.elapsed_time(b) * 1e-3) / 1e12)
    # ---- live shared-memory bandwidth (second roofline of SURVEY.md 8(d)): best of 5 launches of the LDS.128 probe ----
    s_iters, s_blocks = 8192, 148 * 2
    sprobe = lambda: _lib.check(_lib.lib().bpl_smem_probe(ctypes.c_void_p(scratch.data_ptr()), s_iters, s_blocks,
                                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for _ in range(2):
        sprobe()
    torch.cuda.synchronize()
    smem_peak = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sprobe(); b.record(); torch.cuda.synchronize()
        smem_peak = max(smem_peak, (s_blocks * 1024.0 * s_iters * 64) / (a.elapsed_time(b) * 1e-3) / 1e9)
    clocks = sampler.stop() if sampler else None

    # ---- secondary single-GPU objects (N=1 only) ----
    config1 = config2 = None
    if world == 1:
        w1 = workload.synth_tokens(CFG1_BATCH, seed=CFG1_SEED, sigma_mode="uniform")
        S1, _ = mean_sk(w1)
        b1 = workload.to_device(w1, dev)
        st1 = lambda: ops.score_tokens(b1, imgs, ps=ps)
        ms1, _, _ = timed(st1, 200, 10)
        k1 = []
        for i in range(13):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); st1(); b.record()
            torch.cuda.synchronize()
            if i >= 3:
                k1.append(a.elapsed_time(b))
        k1_ms = float(np.mean(k1))
        config1 = {"workload": "configs[1]: 4096 synthetic tokens (seed 1234) scored against one image on 1 GPU",
                   "value": CFG1_BATCH * 200 / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / 200,
                   "kernel_ms_l2_flushed": k1_ms,
                   "roofline_frac": (CFG1_BATCH / (k1_ms * 1e-3)) * flops_fwd(S1) / 1e12 / fp32_peak}
        sys.path.insert(0, os.path.join(ROOT, "examples"))
        import fit_parses
        r = fit_parses.fit(P=1024, steps=100, device=str(dev))
        config2 = {"workload": "configs[2]: 1024 parses x 100 Adam steps (fused fwd+bwd, fused Adam, clamps; one CUDA graph per step)",
                   "seconds": r["seconds"], "particles_per_s": r["particles_per_s"], "final_mean_ll": r["final_mean_ll"]}

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
    value = P_total * steps / (ms_dev * 1e-3)
    e2e = P_total * steps / (wall_e2e_ms * 1e-3)
    k_rate = P_rank / (k_ms * 1e-3)
    ach_tf = k_rate * flops_fwd(Sbar) / 1e12
    peak_note = ("live FFMA probe (bpl_fp32_probe) in this run; MEASURED_PEAKS.json holds no CUDA-core FP32 figure, and "
                 "SURVEY.md 8(d) shows the path is FP32-issue/shared-memory bound, not HBM")
    roofline = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                "traffic": None if NCU_DRAM_BYTES_PER_PARTICLE is None else NCU_DRAM_BYTES_PER_PARTICLE * P_rank,
                "peak_nominal": 74.45, "frac_of_nominal": ach_tf / 74.45,      # 148 SMs x 128 FMA/clk x 2 x 1.965 GHz
                "kernel": "bpl_forward1s_kernel<false, 256, 4>", "kernel_ms": k_ms, "particles_per_launch": P_rank,
                "flop_per_particle": flops_fwd(Sbar), "mean_substrokes": Sbar, "mean_strokes": Kbar, "peak_source": peak_note,
                "smem": {"achieved": k_rate * NCU_SMEM_WAVEFRONTS_PER_PARTICLE * 128.0 / 1e9, "peak": smem_peak, "unit": "GB/s",
                         "frac": k_rate * NCU_SMEM_WAVEFRONTS_PER_PARTICLE * 128.0 / 1e9 / smem_peak,
                         "note": "achieved = shared-memory wavefronts per particle (ncu, profiles/) x 128 B x particles/s; peak = live "
                                 "LDS.128 probe (bpl_smem_probe); nominal 148 SMs x 128 B/clk x 1.965 GHz = 37.2 TB/s"},
                "hbm": {"achieved": k_rate * hbm_bytes(Sbar, Kbar) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": k_rate * hbm_bytes(Sbar, Kbar) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    bwd_rate = BWD_BATCH / (kb_ms * 1e-3)
    fwd_bwd = {"value": world * BWD_BATCH * n_bwd / (ms_bwd * 1e-3), "unit": UNIT, "batch_per_gpu": BWD_BATCH, "steps": n_bwd,
               "scaling": "weak", "ms_per_step": ms_bwd / n_bwd,
               "roofline": {"bound": "fp32", "achieved": bwd_rate * flops_fwd_bwd(Sb) / 1e12, "peak": fp32_peak,
                            "unit": "TFLOP/s", "frac": bwd_rate * flops_fwd_bwd(Sb) / 1e12 / fp32_peak,
                            # ncu (profiles/r02_backward_teams_final2_1024.raw.txt, one launch of 1 024 parses): 591.9 KB read +
                            # 4.9 KB written (the gradient stores stay in L2 for the length of the launch; 8 192 parses: 327 B each)
                            "traffic": 596.7e3 if BWD_BATCH == 1024 else 327.0 * BWD_BATCH,
                            "kernel": "bpl_backward_teams_kernel" if BWD_BATCH >= 1024 else "bpl_backward_kernel<false>", "kernel_ms": kb_ms, "flop_per_particle": flops_fwd_bwd(Sb),
                            "mean_substrokes": Sb},
               "e2e": {"value": world * BWD_BATCH * n_bwd / (wall_bwd_e2e_ms * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": int(bwd_h2d), "d2h_bytes_per_step": int(bwd_d2h)}}
    cores = os.cpu_count() or 1
    cpu = cpu_torch = None
    if world == 1:
        import oracle
        cores = oracle.set_threads(cores)
        ws, img_cpu, C = cpu_sample()
        cv, cn, cdt, cpasses = oracle_rate(ws, img_cpu, C, budget_s=args.cpu_budget * 1.2)
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{CPU_SAMPLE} particles of sweep block 0 x {cpasses} passes = {cn} particles, C oracle port (fp32), OpenMP "
                         f"over tokens, {cdt:.1f} s wall on {cores} cores"}
        cpu_torch = torch_reference_rate(ws, img_cpu, budget_s=args.cpu_budget)
        wbs = workload.permute(wb, np.arange(256))
        gv, gn, gdt, gp = oracle_rate(wbs, img_cpu, C, grad=True, budget_s=6.0)
        fwd_bwd["cpu_baseline"] = {"value": gv, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"256 of the step's parses x {gp} passes, C oracle port fwd+bwd, {gdt:.1f} s wall"}
        fwd_bwd["cpu_baseline_torch"] = torch_reference_rate(wbs, img_cpu, grad=True, budget_s=8.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": sweep_config(world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "timing": "wall clock around all steps", "chunk": e2e_chunk, "pack_threads": stager.threads},
            "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_torch if (cpu_torch and "value" in cpu_torch) else cpu, "cpu_baseline_port": cpu,
            "cpu_baseline_torch": cpu_torch, "config4": config4,
            "clocks": clocks, "wall_s": wall_dev_ms * 1e-3, "logsumexp": {"device": lse_dev, "e2e": lse_e2e},
            "particles_per_rank": P_rank, "fwd_bwd": fwd_bwd, "config1": config1, "config2": config2}
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds per CPU-baseline leg (tests shorten it)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20         # bounded: each step is ~1 s of all-core CPU work
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
