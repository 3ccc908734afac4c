#!/usr/bin/env python
"""tests/golden/fit64.npz: BASELINE.json configs[2] at 64 parses, step 0, from the LIVE reference.

The 64 parses of workload.synth_tokens(64, seed=2024, sigma_mode=10.0) (= examples/fit_parses.py's start state) are turned
into real pybpl CharacterTokens and scored by the unmodified CharacterImageDist.score_image against the target image
(a Bernoulli sample of parse 0's own reference-rendered probability map); autograd gives d ll / d {shapes, invscales,
position, blur_sigma, epsilon}.  torch >= 2 refuses to differentiate through the in-place clamp_ of rendering.py:97, so
that one token is made out-of-place in memory (baseline/ref_loader.py, forward bit-identical).  Build container only:

    python tests/golden/make_fit_golden.py                 # fit64.npz
    python tests/golden/make_fit_golden.py --fit-smooth    # fit_smooth.npz (fit_smooth_stk, see make_fit_smooth)
    python tests/golden/make_fit_golden.py --fsize         # fsize.npz (render_image with Parameters.fsize = 7, 5, 3)
    python tests/golden/make_fit_golden.py --imsize        # imsize.npz (render_image with Parameters.imsize = 64x80, 150x130, 33x47)
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def main(P=64, seed=2024):
    from baseline import ref_loader, torch_ref
    from pybpl_b200 import workload
    ref_loader.load(grad_patch=True)
    from pybpl.model.image_dist import CharacterImageDist
    torch.manual_seed(0)
    w = workload.synth_tokens(P, seed=seed, sigma_mode=10.0)
    toks = torch_ref._tokens_from(w, 0, P)
    dist = CharacterImageDist(None)
    with torch.no_grad():
        pimg0 = dist.get_pimg(toks[0])
    g = torch.Generator(device="cpu").manual_seed(seed)
    target = (torch.rand(105, 105, generator=g) < pimg0).to(torch.float32)
    out = dict(target=target.numpy().astype(np.uint8), ll=np.zeros(P, np.float32), g_ctrl=np.zeros_like(w["ctrl"]),
               g_invscale=np.zeros_like(w["invscale"]), g_position=np.zeros_like(w["position"]),
               g_sigma=np.zeros(P, np.float32), g_eps=np.zeros(P, np.float32))
    tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
    for p, t in enumerate(toks):
        leaves = [x for pt in t.part_tokens for x in (pt.shapes, pt.invscales, pt.position)] + [t.blur_sigma, t.epsilon]
        for x in leaves:
            x.requires_grad_(True)
        ll = dist.score_image(t, target)
        ll.backward()
        out["ll"][p] = float(ll)
        for i, pt in enumerate(t.part_tokens):
            k = tso[p] + i
            out["g_ctrl"][sso[k]:sso[k + 1]] = pt.shapes.grad.permute(2, 0, 1).numpy()
            out["g_invscale"][sso[k]:sso[k + 1]] = pt.invscales.grad.numpy()
            out["g_position"][k] = pt.position.grad.numpy()
        out["g_sigma"][p] = float(t.blur_sigma.grad)
        out["g_eps"][p] = float(t.epsilon.grad)
    path = os.path.join(ROOT, "tests", "golden", "fit64.npz")
    np.savez_compressed(path, P=P, seed=seed, **out)
    print("wrote", path, "mean ll", out["ll"].mean(), {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("g_")})


if __name__ == "__main__" and "--fit-smooth" not in sys.argv and "--fsize" not in sys.argv and "--imsize" not in sys.argv:
    main()


def make_fit_smooth():
    """tests/golden/fit_smooth.npz: the reference's fit_smooth_stk (pybpl/bottomup/initialize/util.py:13-31) on a few
    trajectories.  pybpl.bottomup needs scikit-image at import time, which this image lacks, so the function's own source
    text is executed against the reference's unif_space / fit_bspline_to_traj / get_stk_from_bspline (nothing is edited)."""
    import ast
    from baseline import ref_loader
    ref_loader.load()
    import pybpl.splines as S
    import importlib.util      # pybpl/data/__init__.py does not import under numpy 2 (np.infty): load the one file
    spec = importlib.util.spec_from_file_location("_ref_unif_space", os.path.join(ref_loader.REF, "pybpl", "data", "unif_space.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    unif_space = mod.unif_space
    src = open(os.path.join(ref_loader.REF, "pybpl", "bottomup", "initialize", "util.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "fit_smooth_stk")
    ns = dict(torch=torch, np=np, unif_space=unif_space, fit_bspline_to_traj=S.fit_bspline_to_traj,
              get_stk_from_bspline=S.get_stk_from_bspline)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "util.py", "exec"), ns)
    rng = np.random.default_rng(3)
    t = np.linspace(0, 1, 60)
    strokes = [np.stack([10 + 60 * t, -20 - 40 * t], 1),                                        # a line
               np.stack([50 + 25 * np.cos(5 * t), -50 + 25 * np.sin(5 * t)], 1),                # most of a circle
               np.stack([10 + 80 * t, -50 + 20 * np.sin(9 * t)], 1),                            # a wave
               np.stack([30 + 40 * t * np.cos(8 * t), -50 + 40 * t * np.sin(8 * t)], 1),        # a spiral
               np.stack([20 + 50 * t + rng.normal(0, 0.6, 60), -30 - 30 * t ** 2 + rng.normal(0, 0.6, 60)], 1),   # noisy arc
               np.array([[40.0, -40.0]])]                                                       # a single point
    out = {}
    for i, s in enumerate(strokes):
        s = s.astype(np.float32)
        out[f"in_{i}"] = s
        out[f"unif_{i}"] = np.asarray(unif_space(s), np.float32)
        out[f"out_{i}"] = np.asarray(ns["fit_smooth_stk"](s), np.float32)
    path = os.path.join(ROOT, "tests", "golden", "fit_smooth.npz")
    np.savez_compressed(path, n=len(strokes), **out)
    print("wrote", path, [out[f"out_{i}"].shape for i in range(len(strokes))])


if __name__ == "__main__" and "--fit-smooth" in sys.argv:
    make_fit_smooth()


def make_fsize():
    """tests/golden/fsize.npz: render_image with a smaller Gaussian (Parameters.fsize = 7, 5, 3; pybpl/parameters.py:41)
    from the live reference: probability maps and autograd gradients for a cotangent, three stroke sets per size."""
    from baseline import ref_loader
    ref_loader.load(grad_patch=True)
    from pybpl.parameters import Parameters
    from pybpl.rendering import render_image
    rng = np.random.default_rng(17)
    t = np.linspace(0, 1, 40)
    sets = [[np.stack([20 + 60 * t, -30 - 30 * t], 1)],
            [np.stack([50 + 30 * np.cos(6 * t), -50 + 30 * np.sin(6 * t)], 1), np.stack([10 + 85 * t, -90 + 5 * np.sin(9 * t)], 1)],
            [np.stack([2 + 100 * t, -3 - 99 * t], 1), np.stack([100 - 40 * t, -10 - 20 * t ** 2], 1), np.stack([55 + 0 * t, -5 - 95 * t], 1)]]
    out = {"n_sets": len(sets)}
    for fsize in (7, 5, 3):
        ps = Parameters(); ps.fsize = fsize
        for i, strokes in enumerate(sets):
            st = [torch.tensor(s, dtype=torch.float32, requires_grad=True) for s in strokes]
            eps = torch.tensor(1e-3, requires_grad=True); sig = torch.tensor(0.5 + 1.3 * i, requires_grad=True)
            pimg, off = render_image(st, eps, sig, ps)
            ct = torch.from_numpy(rng.standard_normal((105, 105)).astype(np.float32))
            (pimg * ct).sum().backward()
            key = f"f{fsize}_s{i}"
            out[key + "_pimg"] = pimg.detach().numpy(); out[key + "_ct"] = ct.numpy()
            out[key + "_off"] = bool(off); out[key + "_eps"] = float(eps); out[key + "_sigma"] = float(sig)
            out[key + "_g_eps"] = float(eps.grad); out[key + "_g_sigma"] = float(sig.grad)
            for j, s in enumerate(st):
                out[f"{key}_stroke{j}"] = s.detach().numpy(); out[f"{key}_g_stroke{j}"] = s.grad.numpy()
            out[key + "_n"] = len(st)
    path = os.path.join(ROOT, "tests", "golden", "fsize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__" and "--fsize" in sys.argv:
    make_fsize()


def make_imsize():
    """tests/golden/imsize.npz: render_image with Parameters.imsize other than 105 x 105 (pybpl/parameters.py:21; the image
    is H x W = imsize[0] x imsize[1], rows bounded by imsize[0], rendering.py:27-31,214) from the live reference:
    probability maps, the off-page flag, autograd gradients for a random cotangent, and -- against a Bernoulli sample of
    the map -- the log-likelihood with its gradients (model/image_dist.py:75-78)."""
    from baseline import ref_loader
    ref_loader.load(grad_patch=True)
    from pybpl.parameters import Parameters
    from pybpl.rendering import render_image
    rng = np.random.default_rng(23)
    t = np.linspace(0, 1, 60)

    def sets(H, W):
        # x runs over columns [0, W-1], -y over rows [0, H-1]
        a = [np.stack([0.2 * W + 0.6 * W * t, -(0.25 * H + 0.5 * H * t)], 1)]
        b = [np.stack([0.5 * W + 0.3 * W * np.cos(6 * t), -(0.5 * H + 0.3 * H * np.sin(6 * t))], 1),
             np.stack([0.05 * W + 1.05 * W * t, -(0.8 * H + 0.1 * H * np.sin(9 * t))], 1)]           # runs off the right edge
        c = [np.stack([(W - 1) * t, -(H - 1) * t], 1),                                                # corner to corner, ends exactly on the border
             np.stack([(W - 1) + 0 * t[:20], -(H - 1) * t[:20]], 1),                                   # along the last column
             np.array([[W - 1.0, -(H - 1.0)], [np.nextafter(np.float32(W - 1), np.float32(1e9)), -1.0], [3.0, -2.0]], np.float64),
             np.stack([0.3 * W + 0.6 * t[:12] / t[11], -(0.3 * H + 0.35 * t[:12] / t[11])], 1)]       # 0.7 px long: rescale branch
        return [a, b, c]

    out = {}
    cases = []
    for (H, W, hinton) in ((64, 80, False), (150, 130, False), (33, 47, True)):
        ps = Parameters(); ps.imsize = torch.Size([H, W])
        if hinton:
            ps.broaden_mode = 'Hinton'
        for i, strokes in enumerate(sets(H, W)):
            key = f"h{H}w{W}_s{i}"
            cases.append(key)
            eps_v = (1e-3, 1e-4, 0.0)[i]; sig_v = (0.5, 2.5, 6.0)[i] if not (hinton and i == 0) else 0.0
            mk = lambda: ([torch.tensor(s, dtype=torch.float32, requires_grad=True) for s in strokes],
                          torch.tensor(eps_v, requires_grad=eps_v > 0), torch.tensor(sig_v, requires_grad=sig_v > 0))
            st, eps, sig = mk()
            pimg, off = render_image(st, eps, sig, ps)
            assert tuple(pimg.shape) == (H, W)
            ct = torch.from_numpy(rng.standard_normal((H, W)).astype(np.float32))
            (pimg * ct).sum().backward()
            out[key + "_H"] = H; out[key + "_W"] = W; out[key + "_hinton"] = int(hinton)
            out[key + "_pimg"] = pimg.detach().numpy(); out[key + "_ct"] = ct.numpy()
            out[key + "_off"] = bool(off); out[key + "_eps"] = np.float32(eps_v); out[key + "_sigma"] = np.float32(sig_v)
            out[key + "_g_eps"] = float(eps.grad) if eps.grad is not None else 0.0
            out[key + "_g_sigma"] = float(sig.grad) if sig.grad is not None else 0.0
            for j, s in enumerate(st):
                out[f"{key}_stroke{j}"] = s.detach().numpy(); out[f"{key}_g_stroke{j}"] = s.grad.numpy()
            out[key + "_n"] = len(st)
            # score against a sample of the map
            g = torch.Generator().manual_seed(100 + len(cases))
            image = (torch.rand(H, W, generator=g) < pimg.detach()).float()
            st2, eps2, sig2 = mk()
            pimg2, _ = render_image(st2, eps2, sig2, ps)
            ll = torch.distributions.Bernoulli(pimg2).log_prob(image).sum()
            ll.backward()
            out[key + "_image"] = image.numpy().astype(np.uint8); out[key + "_ll"] = np.float32(ll.item())
            out[key + "_gll_eps"] = float(eps2.grad) if eps2.grad is not None else 0.0
            out[key + "_gll_sigma"] = float(sig2.grad) if sig2.grad is not None else 0.0
            for j, s in enumerate(st2):
                out[f"{key}_gll_stroke{j}"] = s.grad.numpy()
    out["cases"] = np.array(cases)
    path = os.path.join(ROOT, "tests", "golden", "imsize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", cases)


if __name__ == "__main__" and "--imsize" in sys.argv:
    make_imsize()
