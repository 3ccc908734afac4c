"""GPU parity of the run-time-geometry kernels (bpl_render_strokes_any[_grad], csrc/bpl_anysize.cuh): render_image,
the Bernoulli score and their gradients for Parameters.imsize other than 105 x 105 (pybpl/parameters.py:21).

Bars as for the fused kernels: pimg within 1e-5 absolute, ll within 1e-5 relative, off-page flags equal; gradients per
tensor in inf-norm-relative terms (points 5e-4, d/d epsilon 2e-3, d/d sigma 1e-2 -- the reference's own fp32 noise
against fp64).  Goldens: tests/golden/imsize.npz, made from the live reference by
tests/golden/make_fit_golden.py --imsize (64 x 80, 150 x 130, 33 x 47 with the Hinton broaden)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from tests.util_golden import relinf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))


def _ps(H, W, hinton=False):
    return NS(imsize=(H, W), fsize=11, ink_ncon=2, ink_pp=2.0, ink_max_dist=2.0, ink_a=0.5, ink_b=6.0,
              broaden_mode="Hinton" if hinton else "Lake")


def _close(got, want, tol):
    return abs(got - want) <= tol * max(1.0, abs(want))


def test_render_image_at_other_sizes_matches_reference():
    from pybpl_b200 import rendering, ops
    d = np.load(os.path.join(HERE, "golden", "imsize.npz"))
    for key in d["cases"]:
        H, W = int(d[key + "_H"]), int(d[key + "_W"])
        ps = _ps(H, W, bool(d[key + "_hinton"]))
        n = int(d[key + "_n"])
        eps_v, sig_v = float(d[key + "_eps"]), float(d[key + "_sigma"])

        def leaves():
            st = [torch.from_numpy(d[f"{key}_stroke{j}"].copy()).requires_grad_(True) for j in range(n)]
            return st, torch.tensor(eps_v, requires_grad=True), torch.tensor(sig_v, requires_grad=True)

        # probability map and a random cotangent on it (the drop-in call: CPU tensors in, CPU tensors out)
        st, eps, sig = leaves()
        pimg, off = rendering.render_image(st, eps, sig, ps)
        assert tuple(pimg.shape) == (H, W) and bool(off) == bool(d[key + "_off"]), key
        assert np.abs(pimg.detach().numpy() - d[key + "_pimg"]).max() < 1e-5, key
        (pimg * torch.from_numpy(d[key + "_ct"])).sum().backward()
        for j, s in enumerate(st):
            assert relinf(s.grad.numpy(), d[f"{key}_g_stroke{j}"]) < 5e-4, (key, j)
        assert _close(sig.grad.item(), float(d[key + "_g_sigma"]), 1e-2), key
        assert _close(eps.grad.item(), float(d[key + "_g_eps"]), 2e-3), key

        # score against the golden's target image, gradients through the score
        st, eps, sig = leaves()
        pts = torch.cat(st).to(DEV)
        off_ = np.concatenate([[0], np.cumsum([len(s) for s in st])])
        batch = ops.StrokeBatch([0, n], off_, pts, eps.to(DEV).reshape(1), sig.to(DEV).reshape(1))
        p = ops.make_params(ps)
        imgs = ops.binary_images(torch.from_numpy(d[key + "_image"]), p, DEV)
        _, ll, offp = ops.render_strokes(batch, imgs, want_pimg=False, ps=p)
        assert _close(ll.item(), float(d[key + "_ll"]), 1e-5) and bool(offp[0]) == bool(d[key + "_off"]), key
        ll.sum().backward()
        for j, s in enumerate(st):
            assert relinf(s.grad.numpy(), d[f"{key}_gll_stroke{j}"]) < 5e-4, (key, j)
        assert _close(sig.grad.item(), float(d[key + "_gll_sigma"]), 1e-2), key
        assert _close(eps.grad.item(), float(d[key + "_gll_eps"]), 2e-3), key


def test_non_binary_target_is_rejected_at_other_sizes():
    from pybpl_b200 import ops
    p = ops.make_params(_ps(20, 30))
    with pytest.raises(ValueError):
        ops.binary_images(torch.full((20, 30), 0.5), p, DEV)
    with pytest.raises(ValueError):
        ops.binary_images(torch.zeros(30, 20), p, DEV)
    from pybpl_b200 import _lib
    with pytest.raises(_lib.BplError):      # the fused entry points refuse a Params of another size
        from pybpl_b200 import workload
        ops.score_tokens_lse(workload.to_device(workload.synth_tokens(4, seed=1), DEV),
                             ops.pack_images(torch.zeros(105, 105, device=DEV)), ops.LseState(torch.device(DEV)), ps=p)


def test_token_batches_at_another_size_against_the_oracle():
    """Control-point tokens (CharacterImageDist.get_pimg / score_image with image_dist.ps.imsize changed) at 48 x 72, more
    characters than resident CTAs (the kernels' grid-stride loop), with and without an affine warp: vanilla_to_motor
    kernel -> apply_warp -> run-time-geometry render, against the oracle's whole-token path."""
    import oracle as O
    from oracle import Oracle
    from pybpl_b200 import ops, workload
    H, W = 48, 72
    P = 700
    w = workload.synth_tokens(P, seed=77, sigma_mode="uniform")
    # bring the synthetic 105 x 105 coordinates into the smaller page (positions and control points scale together)
    sc = np.array([W / 105.0, H / 105.0], np.float32)
    w["ctrl"] = (w["ctrl"] * sc).astype(np.float32); w["position"] = (w["position"] * sc).astype(np.float32)
    rng = np.random.default_rng(5)
    aff = np.tile(np.array([1, 1, 0, 0], np.float32), (P, 1))
    aff[::3] = np.stack([rng.uniform(0.8, 1.2, len(aff[::3])), rng.uniform(0.8, 1.2, len(aff[::3])),
                         rng.uniform(-4, 4, len(aff[::3])), rng.uniform(-4, 4, len(aff[::3]))], 1).astype(np.float32)
    ops_ps = O.default_params(); ops_ps.H = H; ops_ps.W = W
    o = Oracle(np.float32, ops_ps)
    C = np.load(os.path.join(os.path.dirname(HERE), "pybpl_b200", "data", "basis_5_200.npy"))
    k1, s1 = int(w["tok_stroke_off"][1]), int(w["stroke_sub_off"][int(w["tok_stroke_off"][1])])
    r0 = o.token(C, np.diff(w["stroke_sub_off"][:k1 + 1]), w["ctrl"][:s1], w["invscale"][:s1], w["position"][:k1],
                 float(w["eps"][0]), float(w["sigma"][0]))
    img = (rng.random((H, W)) < r0["pimg"]).astype(np.uint8)      # a sample of token 0's own map
    ref = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"], w["eps"],
                        w["sigma"], img[None], affine=aff, grad=True, want_pimg=True)
    batch = workload.to_device(w, DEV, requires_grad=True)
    batch.affine = torch.from_numpy(aff).to(DEV).requires_grad_(True)
    p = ops.make_params(_ps(H, W))
    pimg, off = ops.render_tokens(batch, ps=p)
    assert tuple(pimg.shape) == (P, H, W)
    assert np.array_equal(off.cpu().numpy(), ref["off_page"])
    assert np.abs(pimg.detach().cpu().numpy() - ref["pimg"]).max() < 1e-5
    ll, off2 = ops.score_tokens(batch, ops.binary_images(torch.from_numpy(img), p, DEV), ps=p)
    # Token 0 is scored against its own sample (1e-5); every other token against SOMEBODY ELSE'S image, where pixels with
    # x = 0 and p close to 1 turn a few ulp of p into 1e-4 of log(1 - p): there two valid fp32 evaluations differ by more
    # than 1e-5, and the bar is the fp32 restatement's own distance from the fp64 oracle (as in
    # test_config5_cross_scoring_against_oracle).
    got = ll.detach().cpu().numpy()
    assert abs(got[0] - ref["ll"][0]) <= 1e-5 * abs(ref["ll"][0])
    o64 = Oracle(np.float64, ops_ps)
    r64 = o64.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"], w["eps"],
                          w["sigma"], img[None], affine=aff, grad=True)
    ek = np.abs(got - r64["ll"]) / np.abs(r64["ll"]); eo = np.abs(ref["ll"] - r64["ll"]) / np.abs(r64["ll"])
    print("48 x 72 tokens vs fp64: kernel max %.2e median %.2e | fp32 restatement max %.2e median %.2e" %
          (ek.max(), np.median(ek), eo.max(), np.median(eo)))
    assert ek.max() <= max(eo.max(), 5e-5) and np.median(ek) <= max(np.median(eo), 1e-5)
    ll.sum().backward()
    # gradients against the fp64 oracle, the fp32 restatement's own distance from it printed beside the kernel's
    for name, t, tol in (("g_ctrl", batch.ctrl, 1e-3), ("g_invscale", batch.invscale, 1e-3), ("g_position", batch.position, 1e-3),
                         ("g_affine", batch.affine, 1e-3), ("g_eps", batch.eps, 2e-3), ("g_sigma", batch.sigma, 1e-2)):
        ek, eo = relinf(t.grad.cpu().numpy(), r64[name]), relinf(ref[name], r64[name])
        print(f"  {name}: kernel {ek:.2e}, fp32 restatement {eo:.2e} (inf-norm relative to fp64)")
        assert ek < tol, (name, ek, eo)


def test_general_kernels_agree_with_the_fused_ones_at_105():
    """The run-time-geometry kernels at H = W = 105 against the fused 105 x 105 path on the same raw trajectories."""
    from pybpl_b200 import ops
    rng = np.random.default_rng(3)
    P = 40
    lens = rng.integers(1, 120, size=3 * P)
    pts = np.concatenate([np.cumsum(rng.normal(0, 1.2, (n, 2)), 0) + [rng.uniform(10, 95), -rng.uniform(10, 95)] for n in lens])
    tok = np.arange(0, 3 * P + 1, 3); off = np.concatenate([[0], np.cumsum(lens)])
    eps = torch.full((P,), 1e-3, device=DEV); sig = torch.from_numpy(rng.uniform(0.5, 8, P).astype(np.float32)).to(DEV)
    mk = lambda: ops.StrokeBatch(tok, off, torch.from_numpy(pts.astype(np.float32)).to(DEV).requires_grad_(True),
                                 eps.clone().requires_grad_(True), sig.clone().requires_grad_(True))
    img = (rng.random((105, 105)) < 0.1).astype(np.float32)
    b1 = mk()
    pimg1, ll1, off1 = ops.render_strokes(b1, ops.pack_images(torch.from_numpy(img).to(DEV)), ps=ops.make_params(None))
    p = ops.make_params(None); p.imsize = 0; p.hw = (105, 105 + 0)
    b2 = mk()
    pimg2, ll2, off2 = ops._RenderStrokesAny.apply(b2.pts, b2.eps, b2.sigma, b2, torch.from_numpy(img).to(DEV)[None].contiguous(),
                                                   None, True, p)
    assert torch.equal(off1, off2.bool())
    assert (pimg1 - pimg2).abs().max().item() < 1e-5
    assert ((ll1 - ll2).abs() / ll1.abs()).max().item() < 1e-5
    ll1.sum().backward(); ll2.sum().backward()
    assert relinf(b2.pts.grad.cpu().numpy(), b1.pts.grad.cpu().numpy()) < 2e-4
    assert relinf(b2.sigma.grad.cpu().numpy(), b1.sigma.grad.cpu().numpy()) < 5e-3
    assert relinf(b2.eps.grad.cpu().numpy(), b1.eps.grad.cpu().numpy()) < 1e-3
