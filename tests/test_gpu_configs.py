"""The BASELINE.json configs beyond the bench line, at sizes the oracle can check in seconds."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def _fit64():
    from pybpl_b200 import workload
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "fit64.npz")))
    return d, workload.synth_tokens(int(d["P"]), seed=int(d["seed"]), sigma_mode=10.0)


def test_config3_step0_gradients_match_the_reference():
    """configs[2] at 64 parses, step 0: scores and gradients for all five optimised leaves against the live reference's
    autograd (tests/golden/fit64.npz, made by tests/golden/make_fit_golden.py with the rendering.py:97 patch).  Bars as
    for the other reference-autograd goldens: the reference's own fp32 noise."""
    from pybpl_b200 import ops, workload
    d, w = _fit64()
    dev = torch.device("cuda:0")
    b = workload.to_device(w, dev, requires_grad=True)
    imgs = ops.pack_images(torch.as_tensor(d["target"], device=dev))
    ll, _ = ops.score_tokens(b, imgs)
    ll.sum().backward()
    assert (np.abs(ll.detach().cpu().numpy() - d["ll"]) / np.abs(d["ll"])).max() < 1e-5
    rel = lambda g, r: np.abs(g.cpu().numpy() - r).max() / np.abs(r).max()
    errs = dict(ctrl=rel(b.ctrl.grad, d["g_ctrl"]), invscale=rel(b.invscale.grad, d["g_invscale"]),
                position=rel(b.position.grad, d["g_position"]), sigma=rel(b.sigma.grad, d["g_sigma"]), eps=rel(b.eps.grad, d["g_eps"]))
    print("config 3 step-0 gradients vs reference autograd (inf-norm rel)", errs)
    assert errs["ctrl"] < 5e-4 and errs["invscale"] < 5e-4 and errs["position"] < 5e-4
    assert errs["sigma"] < 1e-2 and errs["eps"] < 2e-3


def test_config3_fit_follows_an_oracle_driven_adam_run():
    """30 Adam steps on 64 parses through the fused kernel (CUDA graph, fused Adam, clamps: examples/fit_parses.py)
    against the same optimisation driven by the CPU oracle's gradients (numpy Adam with torch's update rule and the
    same clamps): the loss curves agree step by step and the loss goes down."""
    import fit_parses
    from oracle import Oracle
    d, w = _fit64()
    steps, lr = 30, 4e-3
    r = fit_parses.fit(P=64, steps=steps, lr=lr, log_every=1, target=d["target"])
    gpu = np.array(r["losses"])
    assert np.isfinite(gpu).all() and gpu[-1] < gpu[0]
    assert r["final_mean_ll"] > -gpu[0] / 64
    o = Oracle(np.float32)
    C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
    x = {k: w[k].astype(np.float32).copy() for k in ("ctrl", "invscale", "position", "sigma", "eps")}
    m = {k: np.zeros_like(v, dtype=np.float64) for k, v in x.items()}
    v2 = {k: np.zeros_like(v, dtype=np.float64) for k, v in x.items()}
    cpu = []
    for t in range(1, steps + 1):
        ref = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], x["ctrl"], x["invscale"], x["position"], x["eps"],
                            x["sigma"], d["target"][None], grad=True)
        cpu.append(-float(ref["ll"].astype(np.float64).sum()))
        for k, gk in (("ctrl", "g_ctrl"), ("invscale", "g_invscale"), ("position", "g_position"), ("sigma", "g_sigma"), ("eps", "g_eps")):
            g = -ref[gk].astype(np.float64)                      # loss = -sum ll
            m[k] = 0.9 * m[k] + 0.1 * g
            v2[k] = 0.999 * v2[k] + 0.001 * g * g
            step = lr / (1 - 0.9 ** t) * m[k] / (np.sqrt(v2[k] / (1 - 0.999 ** t)) + 1e-8)      # torch.optim.Adam
            x[k] = (x[k].astype(np.float64) - step).astype(np.float32)
        x["sigma"] = np.clip(x["sigma"], 0.5, 16.0); x["eps"] = np.clip(x["eps"], 1e-4, 0.5)       # parameters.py:54-63
        x["invscale"] = np.maximum(x["invscale"], 1e-3)
    cpu = np.array(cpu)
    rel = np.abs(gpu - cpu) / np.abs(cpu)
    print("config 3: loss %.1f -> %.1f (GPU), %.1f -> %.1f (oracle Adam); max rel difference over the curve %.2e, final %.2e"
          % (gpu[0], gpu[-1], cpu[0], cpu[-1], rel.max(), rel[-1]))
    assert rel[0] < 1e-5
    assert rel.max() < 2e-3 and rel[-1] < 2e-3


def test_fused_adam_matches_torch_adam():
    """ops.FusedAdam (one launch: Adam + bounds) against torch.optim.Adam + clamp_ on the same gradients, 25 steps, eager
    and replayed from a CUDA graph (the step counter lives on the device)."""
    from pybpl_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    shapes = [(700, 5, 2), (700,), (300, 2), (128,), (128,)]
    init = [torch.randn(s, generator=g).to(dev) for s in shapes]
    grads = [[torch.randn(s, generator=g).to(dev) * (10.0 ** (i % 3 - 1)) for s in shapes] for i in range(25)]
    bounds = {3: (0.5, 16.0), 4: (1e-4, 0.5), 1: (1e-3, None)}
    ref = [p.clone().requires_grad_(True) for p in init]
    topt = torch.optim.Adam(ref, lr=4e-3)
    for gs in grads:
        for p, gr in zip(ref, gs):
            p.grad = -gr                                          # FusedAdam below ascends
        topt.step()
        with torch.no_grad():
            ref[3].clamp_(0.5, 16.0); ref[4].clamp_(1e-4, 0.5); ref[1].clamp_(min=1e-3)
    for use_graph in (False, True):
        mine = [p.clone().requires_grad_(True) for p in init]
        opt = ops.FusedAdam(mine, lr=4e-3, bounds=bounds, maximize=True)
        gbuf = [torch.zeros_like(p) for p in mine]
        for p, gb in zip(mine, gbuf):
            p.grad = gb
        graph = None
        if use_graph:
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                opt.step()                                        # warm-up, then put everything back
            torch.cuda.current_stream().wait_stream(side)
            with torch.no_grad():
                for p, p0 in zip(mine, init):
                    p.copy_(p0)
            opt.reset()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                opt.step()
            with torch.no_grad():
                for p, p0 in zip(mine, init):
                    p.copy_(p0)
            opt.reset()
        for gs in grads:
            for gb, gr in zip(gbuf, gs):
                gb.copy_(gr)
            graph.replay() if graph is not None else opt.step()
        torch.cuda.synchronize()
        assert abs(float(opt.state[0]) - 25.0) < 1e-6
        for a, b in zip(mine, ref):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (use_graph, (a - b).abs().max())


def test_config3_fit_with_token_prior_term():
    """The fit objective with log P(token | type) added (model.py:59-62): finite, decreasing, and the prior term's
    value on the synthetic types agrees with the numpy oracle."""
    import fit_parses
    from oracle import token_prior as tp
    from pybpl_b200 import ops, workload
    r = fit_parses.fit(P=64, steps=30, log_every=1, token_prior=True)
    assert np.isfinite(r["losses"]).all() and r["losses"][-1] < r["losses"][0]
    w = workload.synth_tokens(200, seed=9, sigma_mode="uniform")
    ty = workload.synth_types(w, 10)
    dev = torch.device("cuda:0")
    got = ops.score_token_prior(workload.to_device(w, dev), workload.types_to_device(ty, dev)).cpu().numpy()
    c = ops.token_prior_params()
    d = dict(K=np.diff(w["tok_stroke_off"]), nsub=np.diff(w["stroke_sub_off"]), ctrl=w["ctrl"], invscale=w["invscale"],
             position=w["position"], sigma=w["sigma"], consts=np.array([c.sigma_shape, c.sigma_invscale, c.sigma_attach,
                                                                        c.sigma_x, c.sigma_y, c.min_blur_sigma, c.max_blur_sigma]), **ty)
    C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
    ref = tp.score(d, np.float64, C0=C[0], C1=C[-1])
    assert np.all(np.isfinite(ref))
    assert (np.abs(got - ref) / np.abs(ref)).max() < 1e-5


def test_config4_sweep_logsumexp_matches_torch():
    from pybpl_b200 import ops, sweep, workload
    dev = torch.device("cuda:0")
    w = workload.synth_tokens(3000, seed=7, sigma_mode="uniform")
    g = torch.Generator().manual_seed(7)
    imgs = ops.pack_images((torch.rand(105, 105, generator=g) < 0.04).to(torch.uint8).to(dev))
    ll, _ = ops.score_tokens(workload.to_device(w, dev), imgs)
    total = sweep.combine_partials(ops.logsumexp_partial(ll))
    assert abs(total - torch.logsumexp(ll.double(), 0).item()) < 1e-4
    # sharding invariance: three shards combine to the same value
    parts = []
    for r in range(3):
        lls, _ = ops.score_tokens(workload.to_device(workload.shard(w, r, 3), dev), imgs)
        parts.append(ops.logsumexp_partial(lls))
    assert abs(sweep.logsumexp_from_partials(torch.stack(parts)) - total) < 1e-3


def test_config5_cross_scoring_against_oracle():
    import cross_score
    from oracle import Oracle
    from pybpl_b200 import workload
    ll, _, w, images = cross_score.cross_score(n_tokens=40, n_images=5, seed=5)
    ll = ll.cpu().numpy()
    o = Oracle(np.float32)
    C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
    imgs = images.cpu().numpy()
    rep = np.repeat(np.arange(40), 5)
    wb = workload.permute(w, rep)
    idx = np.tile(np.arange(5, dtype=np.int32), 40)
    args = (wb["tok_stroke_off"], wb["stroke_sub_off"], wb["ctrl"], wb["invscale"], wb["position"], wb["eps"], wb["sigma"], imgs)
    ref32 = o.token_batch(C, *args, img_idx=idx)["ll"].astype(np.float64)
    ref64 = Oracle(np.float64).token_batch(C.astype(np.float64), *args, img_idx=idx)["ll"]
    # A token scored against somebody else's image has x = 0 pixels where p ~ 0.9996: lp = log(1 - p) turns an
    # 8-ulp difference in p into 1e-3 per pixel, so two fp32 evaluations of the same renderer differ by ~1e-4
    # relative (the reference's direct 121-tap convolution is the noisier one).  The bar is therefore the
    # reference's own distance from fp64, not 1e-5 between two fp32 results.
    ek = np.abs(ll.reshape(-1) - ref64) / np.abs(ref64)
    eo = np.abs(ref32 - ref64) / np.abs(ref64)
    print("cross-scoring vs fp64: kernel max %.2e median %.2e | fp32 restatement max %.2e median %.2e" %
          (ek.max(), np.median(ek), eo.max(), np.median(eo)))
    assert ek.max() <= max(eo.max(), 5e-5)
    assert np.median(ek) <= max(np.median(eo), 1e-5)
    own = np.arange(0, 40, 8)                  # tokens scored against their own image: well conditioned
    rel_own = np.abs(ll[own, np.arange(5)] - ref32.reshape(40, 5)[own, np.arange(5)]) / np.abs(ref32.reshape(40, 5)[own, np.arange(5)])
    assert rel_own.max() < 1e-5, rel_own.max()


def test_all_images_scoring_equals_pairwise():
    from pybpl_b200 import ops, workload
    dev = torch.device("cuda:0")
    w = workload.synth_tokens(37, seed=9, max_strokes=10, kappa="prior", sigma_mode="uniform")     # odd count: last pair is half empty
    batch = workload.to_device(w, dev)
    g = torch.Generator().manual_seed(3)
    images = (torch.rand((7, 105, 105), generator=g) < 0.05).to(torch.uint8).to(dev)
    packed = ops.pack_images(images)
    ll_all, off_all = ops.score_tokens_all_images(batch, packed)
    assert ll_all.shape == (37, 7)
    for m in range(7):
        ll_m, off_m = ops.score_tokens(batch, ops.pack_images(images[m]))
        assert torch.equal(off_m, off_all)
        assert ((ll_all[:, m] - ll_m).abs() <= 2e-5 * ll_m.abs()).all()
    # a second call right after (buffer re-zeroing between pairs and launches)
    ll2, _ = ops.score_tokens_all_images(batch, packed)
    assert ((ll2 - ll_all).abs() <= 2e-5 * ll_all.abs()).all()


def test_all_images_scoring_large_batch_shape():
    """From 2 048 tokens on the scoring kernel runs as 256 threads x 4 CTAs per SM (below: 320 x 3), in all-pairs mode too
    (the cross-scoring bench, configs[4], runs there): all-images scores of 2 300 tokens equal their pairwise scores, and
    the first 40 tokens scored alone (the small shape) give the same numbers."""
    from pybpl_b200 import ops, workload
    dev = torch.device("cuda:0")
    w = workload.synth_tokens(2300, seed=19, max_strokes=10, kappa="prior", sigma_mode="fixed")
    batch = workload.to_device(w, dev)
    g = torch.Generator().manual_seed(4)
    images = (torch.rand((3, 105, 105), generator=g) < 0.05).to(torch.uint8).to(dev)
    packed = ops.pack_images(images)
    with torch.no_grad():
        ll_all, off_all = ops.score_tokens_all_images(batch, packed)
        for m in range(3):
            ll_m, off_m = ops.score_tokens(batch, ops.pack_images(images[m]))
            assert torch.equal(off_m, off_all)
            assert ((ll_all[:, m] - ll_m).abs() <= 2e-5 * ll_m.abs()).all()
        small, _ = ops.score_tokens_all_images(workload.to_device(workload.permute(w, np.arange(40)), dev), packed)
        assert ((small - ll_all[:40]).abs() <= 2e-5 * small.abs()).all()


def test_parse_file_upload_scores_like_the_original(tmp_path):
    """a batch + types + image set written to a parse file, loaded with ONE host->device copy, gives the
    scores of the original tensors; the host bit packing equals bpl_pack_images_u8."""
    import time
    from pybpl_b200 import ops, packed, workload
    dev = torch.device("cuda:0")
    w = workload.synth_tokens(2000, seed=21, sigma_mode="uniform")
    ty = workload.synth_types(w, 22)
    rng = np.random.default_rng(1)
    imgs = (rng.random((5, 105, 105)) < 0.08).astype(np.uint8)
    idx = rng.integers(0, 5, size=2000).astype(np.int32)
    path = str(tmp_path / "p.bpl")
    nbytes = packed.save(path, w, ty, imgs)
    pf = packed.ParseFile.load(path)
    t0 = time.perf_counter()
    batch, types, images = pf.upload(dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"parse file: {nbytes / 1e6:.2f} MB, upload {dt * 1e3:.2f} ms")
    ref_imgs = ops.pack_images(torch.from_numpy(imgs).to(dev))
    assert torch.equal(images.bits, ref_imgs.bits)
    ti = torch.from_numpy(idx).to(dev)
    ll_a, off_a = ops.score_tokens(batch, images, ti)
    ll_b, off_b = ops.score_tokens(workload.to_device(w, dev), ref_imgs, ti)
    assert torch.allclose(ll_a, ll_b, rtol=2e-6, atol=0) and torch.equal(off_a, off_b)
    pr_a = ops.score_token_prior(batch, types)
    pr_b = ops.score_token_prior(workload.to_device(w, dev), workload.types_to_device(ty, dev))
    assert torch.equal(pr_a, pr_b)
    allp, _ = ops.score_tokens_all_images(batch, images)
    assert allp.shape == (2000, 5)
    assert torch.allclose(allp[torch.arange(2000, device=dev), ti.long()], ll_a, rtol=2e-5, atol=0)
