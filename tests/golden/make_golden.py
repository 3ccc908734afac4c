#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference (rfeinman/pyBPL at /root/reference) is imported unmodified except
for (a) a stub `matplotlib` (not installed here; only `pybpl.library` wants it)
and (b) the in-memory, forward-bit-identical rewrite of the in-place clamp at
pybpl/rendering.py:97, without which torch>=2 autograd refuses to differentiate
through `add_stroke`.  Everything written here is an OUTPUT of
the reference's own functions:

  splines.npz  <- pybpl/splines.py:57-93,108-138 (tables, evals, adaptive neval)
  render.npz   <- pybpl/rendering.py:185-229 on hand-built edge-case trajectories
                  (+ Bernoulli log_prob as in pybpl/model/image_dist.py:75-78,
                  + autograd gradients)
  tokens.npz   <- CharacterModel.sample_type/sample_token/get_pimg/score_image
                  (pybpl/model/model.py:20-39) on prior-sampled tokens, the
                  pybpl/tests/test_vanilla_to_motor.py:15-42 fixture and the
                  examples/fit_image.ipynb "H" type.

No reference source text is copied; only numbers.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
REF = os.environ.get("PYBPL_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import pybpl.rendering as R
    src = open(R.__file__).read()
    bad = "dist.clamp_(None, ps.ink_max_dist)"
    assert bad in src
    src = src.replace(bad, "dist = dist.clamp(None, ps.ink_max_dist)")
    exec(compile(src, R.__file__, "exec"), R.__dict__)
    import pybpl.model.image_dist as ID
    ID.render_image = R.render_image
    import pybpl.splines as S
    from pybpl.library import Library
    from pybpl.model import CharacterModel
    from pybpl.parameters import Parameters
    return R, S, Library, CharacterModel, Parameters


R, S, Library, CharacterModel, Parameters = load_reference()
from pybpl.objects.part import vanilla_to_motor  # noqa: E402
from pybpl.util.affine import apply_warp  # noqa: E402
import torch.distributions as D  # noqa: E402

f32 = np.float32


def pack_image(img):
    """(105,105) {0,1} float tensor -> packed uint8 rows (np.packbits, bit 7 first)."""
    a = img.detach().numpy().astype(np.uint8)
    assert set(np.unique(a)) <= {0, 1}
    return np.packbits(a, axis=1)


# --------------------------------------------------------------------------
# splines.npz
# --------------------------------------------------------------------------
def make_splines():
    out = {}
    # basis tables: the (nland, neval) pairs the tests exercise
    pairs = [(5, 200), (5, 10), (5, 57), (5, 133), (3, 200), (4, 50), (7, 200), (10, 200), (12, 64)]
    out["table_pairs"] = np.array(pairs, np.int32)
    for nland, neval in pairs:
        S.coefficient_mat.cache_clear()
        out[f"table_{nland}_{neval}"] = S.coefficient_mat(nland, neval).numpy().copy()
        out[f"s_{nland}_{neval}"] = S.bspline_gen_s(nland, neval)[0].numpy().copy()
    # single/explicit-s evaluation
    g = torch.Generator().manual_seed(11)
    svals = torch.cat([torch.tensor([2.0, 2.5, 3.0, 4.5, 5.999, 6.0]),
                       2 + 4 * torch.rand(58, generator=g)])
    S.coefficient_mat.cache_clear()
    out["s_explicit"] = svals.numpy().copy()
    out["table_s_explicit"] = S.coefficient_mat(5, None, s=svals).numpy().copy()
    Y = torch.randn(64, 5, 2, generator=g) * 20
    out["Y_explicit"] = Y.numpy().copy()
    out["X_explicit"] = torch.stack([S.get_stk_from_bspline(y, s=svals) for y in Y]).numpy()
    out["X_200"] = torch.stack([S.get_stk_from_bspline(y, neval=200) for y in Y]).numpy()
    S.coefficient_mat.cache_clear()
    out["X_scalar_s"] = torch.stack(
        [S.get_stk_from_bspline(y, s=torch.tensor(4.5)) for y in Y]).numpy()
    # gradient of bspline eval w.r.t. Y for a random cotangent
    Yg = Y.clone().requires_grad_(True)
    ct = torch.randn(64, 200, 2, generator=g)
    X = torch.stack([S.get_stk_from_bspline(y, neval=200) for y in Yg])
    (X * ct).sum().backward()
    out["ct_200"] = ct.numpy().copy()
    out["gY_200"] = Yg.grad.numpy().copy()
    # other nland
    Y7 = torch.randn(16, 7, 2, generator=g) * 20
    out["Y7"] = Y7.numpy().copy()
    out["X7_200"] = torch.stack([S.get_stk_from_bspline(y, neval=200) for y in Y7]).numpy()
    # adaptive neval (integer work): random splines over a wide scale range
    n = 4000
    scale = torch.exp(torch.rand(n, 1, 1, generator=g) * 5.0 - 1.0)  # e^-1 .. e^4
    Ya = torch.randn(n, 5, 2, generator=g) * scale
    nev = np.zeros(n, np.int32)
    dist = np.zeros(n, f32)
    from pybpl.util.stroke import dist_along_traj
    for i in range(n):
        X = S.get_stk_from_bspline(Ya[i])
        nev[i] = X.shape[0]
        dist[i] = float(dist_along_traj(S.get_stk_from_bspline(Ya[i], neval=200)))
    out["Y_adaptive"] = Ya.numpy().copy()
    out["neval_adaptive"] = nev
    out["dist_adaptive"] = dist
    # a few adaptive trajectories in full
    out["X_adaptive_first8"] = np.concatenate(
        [S.get_stk_from_bspline(Ya[i]).numpy() for i in range(8)])
    np.savez_compressed(os.path.join(OUT, "splines.npz"), **out)
    print("splines.npz:", {k: v.shape for k, v in out.items() if k.startswith("X")},
          "neval hist", np.bincount(nev)[[10, 200]], (nev > 10).sum())


# --------------------------------------------------------------------------
# render.npz : raw-trajectory edge cases through rendering.render_image
# --------------------------------------------------------------------------
def line(p0, p1, n):
    t = torch.linspace(0, 1, n).unsqueeze(1)
    return torch.tensor(p0, dtype=torch.float) * (1 - t) + torch.tensor(p1, dtype=torch.float) * t


def render_cases():
    """Each case: name, list of (n_i,2) motor-space strokes, eps, sigma, broaden_mode."""
    g = torch.Generator().manual_seed(5)
    nxt = float(np.nextafter(f32(104.0), f32(200.0)))
    prv = float(np.nextafter(f32(0.0), f32(-1.0)))
    cases = []

    def add(name, strokes, eps=1e-4, sigma=0.5, mode="Lake", grad=True):
        cases.append(dict(name=name, strokes=[s.clone().float() for s in strokes],
                          eps=eps, sigma=sigma, mode=mode, grad=grad))

    diag = line([20., -20.], [80., -85.], 200)
    curve = torch.stack([30 + 40 * torch.linspace(0, 1, 200),
                         -50 - 25 * torch.sin(torch.linspace(0, 6.0, 200))], 1)
    # (1) all out of bounds
    add("all_out", [line([-50., 10.], [-10., 60.], 200)])
    # (2) exactly one in-bounds point
    s = line([-30., -30.], [-1., -30.], 50)
    s[-1] = torch.tensor([3.25, -30.5])
    add("one_in", [s])
    # (3) all points identical (fill branch)
    add("identical", [torch.tensor([[40.3, -60.7]]).repeat(37, 1)])
    # (4) short stroke, total length < 2 px (rescale branch)
    add("short", [line([50.2, -40.1], [50.9, -40.6], 25)])
    # (5) integer coordinates
    add("integer", [torch.stack([torch.arange(10., 60.), -torch.arange(20., 70.)], 1)])
    # (6) boundary values: exactly 104 / 0 are in, one ulp beyond is out
    add("boundary", [torch.tensor([[0., 0.], [104., 0.], [104., -104.], [0., -104.], [nxt, -50.],
                                   [50., -nxt], [prv, -50.], [50., -prv], [104., -103.5], [52., -52.]])])
    # (7) leave and re-enter the canvas
    add("reenter", [torch.cat([line([90., -50.], [120., -55.], 40), line([120., -55.], [95., -70.], 40)])])
    # (8) segment lengths exactly 2.0, >2 and <2
    add("seglen", [torch.tensor([[10., -10.], [12., -10.], [12., -12.], [15., -12.], [15.5, -12.5],
                                 [25., -30.], [25.5, -30.25]])])
    # (9) heavy overlap: many points on one pixel + dense scribble
    add("overlap", [torch.tensor([[60.5, -60.5]]).repeat(150, 1) + 0.3 * torch.rand(150, 2, generator=g),
                    curve, curve.flip(0)])
    # (10) blur sigma sweep (0 skips the blur branch)
    for sg in (0.0, 0.5, 1.0, 3.0, 10.0, 16.0):
        add(f"sigma_{sg:g}", [diag, curve], sigma=sg)
    # (11) epsilon sweep (0 skips the mix; exercises clamp_probs)
    for ep in (0.0, 1e-4, 0.01, 0.5):
        add(f"eps_{ep:g}", [diag, curve], eps=ep, sigma=1.5)
    # (12) ragged list input
    rag = [torch.tensor([[33.3, -44.4]]), line([10., -90.], [30., -70.], 7), line([70., -10.], [90., -30.], 10),
           curve, torch.stack([20 + 60 * torch.rand(333, generator=g), -20 - 60 * torch.rand(333, generator=g)], 1)]
    add("ragged", rag, sigma=2.0)
    # (13) Hinton broaden mode
    add("hinton", [diag, curve], mode="Hinton", sigma=1.0)
    # (14) corner-hugging strokes (stencil zero padding at the image border)
    add("borders", [line([0., 0.], [104., 0.], 200), line([104., 0.], [104., -104.], 200),
                    line([0., -104.], [104., -104.], 120), line([0., 0.], [0., -104.], 90)], sigma=4.0)
    # (15) NaN-free extreme magnitudes far outside
    add("far_out", [line([-3000., 4000.], [5000., -6000.], 200), diag])
    return cases


def make_render():
    out = {}
    names = []
    for ci, c in enumerate(render_cases()):
        ps = Parameters()
        ps.broaden_mode = c["mode"]
        strokes = [s.clone().requires_grad_(c["grad"]) for s in c["strokes"]]
        eps = torch.tensor(c["eps"], dtype=torch.float, requires_grad=c["grad"] and c["eps"] > 0)
        sig = torch.tensor(c["sigma"], dtype=torch.float, requires_grad=c["grad"] and c["sigma"] > 0)
        pimg, off = R.render_image(strokes, eps, sig, ps)
        torch.manual_seed(1000 + ci)
        image = D.Bernoulli(pimg.detach()).sample()
        ll = D.Bernoulli(pimg).log_prob(image).sum()
        k = f"c{ci:02d}"
        names.append(c["name"])
        out[k + "_pts"] = torch.cat(c["strokes"]).numpy()
        out[k + "_len"] = np.array([len(s) for s in c["strokes"]], np.int32)
        out[k + "_eps"] = f32(c["eps"])
        out[k + "_sigma"] = f32(c["sigma"])
        out[k + "_hinton"] = np.int32(c["mode"] == "Hinton")
        out[k + "_pimg"] = pimg.detach().numpy()
        out[k + "_off"] = np.bool_(bool(off))
        out[k + "_image"] = pack_image(image)
        out[k + "_ll"] = f32(ll.item())
        # integer work per point (rendering.py:27-31,52-53,111-116)
        allp = torch.cat(c["strokes"])
        ip = R.space_motor_to_img(allp)
        outm = R.check_bounds(ip, ps.imsize)
        out[k + "_mask_out"] = outm.numpy()
        out[k + "_floor"] = torch.floor(ip).to(torch.int32).numpy()
        out[k + "_ceil"] = torch.ceil(ip).to(torch.int32).numpy()
        if c["grad"]:
            ll.backward()
            out[k + "_g_pts"] = torch.cat([s.grad if s.grad is not None else torch.zeros_like(s)
                                           for s in strokes]).numpy()
            out[k + "_g_eps"] = f32(eps.grad.item() if eps.grad is not None else 0.0)
            out[k + "_g_sigma"] = f32(sig.grad.item() if sig.grad is not None else 0.0)
            # also a generic cotangent on pimg (render_image drop-in backward)
            strokes2 = [s.clone().requires_grad_(True) for s in c["strokes"]]
            eps2 = eps.detach().clone().requires_grad_(eps.requires_grad)
            sig2 = sig.detach().clone().requires_grad_(sig.requires_grad)
            pimg2, _ = R.render_image(strokes2, eps2, sig2, ps)
            gen = torch.Generator().manual_seed(2000 + ci)
            ct = torch.randn(105, 105, generator=gen)
            if pimg2.requires_grad:
                (pimg2 * ct).sum().backward()
            out[k + "_ct_seed"] = np.int32(2000 + ci)
            out[k + "_gct_pts"] = torch.cat([s.grad if s.grad is not None else torch.zeros_like(s)
                                             for s in strokes2]).numpy()
            out[k + "_gct_eps"] = f32(eps2.grad.item() if eps2.grad is not None else 0.0)
            out[k + "_gct_sigma"] = f32(sig2.grad.item() if sig2.grad is not None else 0.0)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "render.npz"), **out)
    print("render.npz:", len(names), "cases")


# --------------------------------------------------------------------------
# tokens.npz : whole-path cases through CharacterModel
# --------------------------------------------------------------------------
def token_arrays(ctoken):
    """Flatten a reference CharacterToken into the packed arrays the new build takes."""
    ctrl, inv, pos, nsub = [], [], [], []
    for p in ctoken.part_tokens:
        ctrl.append(p.shapes.detach().permute(2, 0, 1))  # (nsub,5,2)
        inv.append(p.invscales.detach())
        pos.append(p.position.detach())
        nsub.append(p.shapes.shape[2])
    return (torch.cat(ctrl).numpy().copy(), torch.cat(inv).numpy().copy(),
            torch.stack(pos).numpy().copy(), np.array(nsub, np.int32))


def score_with_grads(model, ctoken, image):
    leaves = []
    for p in ctoken.part_tokens:
        p.shapes = p.shapes.detach().clone().requires_grad_(True)
        p.invscales = p.invscales.detach().clone().requires_grad_(True)
        p.position = p.position.detach().clone().requires_grad_(True)
        leaves.append(p)
    ctoken.blur_sigma = ctoken.blur_sigma.detach().clone().float().requires_grad_(True)
    ctoken.epsilon = ctoken.epsilon.detach().clone().float().requires_grad_(True)
    if ctoken.affine is not None:
        ctoken.affine = ctoken.affine.detach().clone().requires_grad_(True)
    ll = model.score_image(ctoken, image)
    ll.backward()
    g_ctrl = torch.cat([p.shapes.grad.permute(2, 0, 1) for p in leaves]).numpy().copy()
    g_inv = torch.cat([p.invscales.grad for p in leaves]).numpy().copy()
    g_pos = torch.stack([p.position.grad for p in leaves]).numpy().copy()
    g_sig = f32(ctoken.blur_sigma.grad.item())
    g_eps = f32(ctoken.epsilon.grad.item())
    g_aff = ctoken.affine.grad.numpy().copy() if ctoken.affine is not None else np.zeros(4, f32)
    return f32(ll.item()), g_ctrl, g_inv, g_pos, g_sig, g_eps, g_aff


def make_tokens(n_prior=96, n_full=24):
    torch.manual_seed(0)
    lib = Library(use_hist=True)
    model = CharacterModel(lib)
    ps = Parameters()
    toks = []

    # (15) the reference's own test fixture, through the whole path
    from pybpl.objects.part import StrokeToken
    from pybpl.objects.concept import CharacterToken
    from pybpl.objects.relation import RelationIndependent, RelationToken

    def indep():
        rt = RelationIndependent(category='unihist', gpos=torch.zeros(2),
                                 xlim=lib.Spatial.xlim, ylim=lib.Spatial.ylim)
        return RelationToken(rt)
    shapes = torch.tensor(
        [[[-57.40, 24.25, 39.48], [-24.32, 16.24, -19.35]],
         [[2.67, 4.46, -36.31], [29.17, -22.95, -2.88]],
         [[52.54, -22.68, 40.90], [-21.49, 53.78, 26.89]],
         [[-17.66, 18.58, 50.14], [-2.22, 1.17, -16.29]],
         [[7.58, -30.76, -34.80], [25.86, -61.31, -81.7337]]])
    st = StrokeToken(shapes, torch.tensor([0.51, 0.18, 0.14]), lib.Spatial.xlim, lib.Spatial.ylim)
    st.position = torch.tensor([74.25, -13.28])
    toks.append(("v2m_fixture", CharacterToken([st], [indep()], None, torch.tensor(1e-4), torch.tensor(0.5))))

    # (16) the fit_image notebook "H": primitives 0, 9, 0; invscales .5/.4/.5; sigma 10
    def sub(idx, inv):
        t = StrokeToken(lib.shape['mu'][idx].view(5, 2, 1).clone(), torch.tensor([inv]),
                        lib.Spatial.xlim, lib.Spatial.ylim)
        return t
    h1, h2, h3 = sub(0, 0.5), sub(9, 0.4), sub(0, 0.5)
    h1.position = torch.tensor([32., -20.])
    h3.position = torch.tensor([68., -20.])
    # 'mid' attachment at eval_spot 4.5 along stroke 1 (objects/relation.py:62-65)
    _, ms = vanilla_to_motor(h1.shapes, h1.invscales, h1.position)
    S.coefficient_mat.cache_clear()
    h2.position = S.get_stk_from_bspline(ms[:, :, 0], s=torch.tensor(4.5))[0]
    toks.append(("notebook_H", CharacterToken([h1, h2, h3], [indep() for _ in range(3)], None,
                                              torch.tensor(1e-4), torch.tensor(10.0))))

    # (17) prior-sampled tokens
    g = torch.Generator().manual_seed(77)
    for i in range(n_prior):
        ctype = model.sample_type()
        ctoken = model.sample_token(ctype)
        r = i % 4
        if r == 1:
            ctoken.blur_sigma = 0.5 + 15.5 * torch.rand((), generator=g)
        elif r == 2:
            ctoken.blur_sigma = 0.5 + 3.0 * torch.rand((), generator=g)
            ctoken.epsilon = torch.tensor(10.0) ** (-4 + 3.5 * torch.rand((), generator=g))
        elif r == 3:
            ctoken.affine = torch.tensor([1.0, 1.0, 0.0, 0.0]) + torch.tensor([0.15, 0.15, 4.0, 4.0]) * \
                torch.randn(4, generator=g)
        toks.append((f"prior_{i:03d}", ctoken))

    out = {}
    names, K, S_tot = [], [], []
    ctrl_all, inv_all, pos_all, nsub_all = [], [], [], []
    eps_all, sig_all, aff_all, has_aff = [], [], [], []
    ll_all, off_all, img_all = [], [], []
    gctrl_all, ginv_all, gpos_all, gsig_all, geps_all, gaff_all = [], [], [], [], [], []
    mask_all, nin_all, fsum_all = [], [], []
    for ti, (name, ctoken) in enumerate(toks):
        ctrl, inv, pos, nsub = token_arrays(ctoken)
        with torch.no_grad():
            motor = [p.motor for p in ctoken.part_tokens]
            if ctoken.affine is not None:
                motor = apply_warp(motor, ctoken.affine)
            motor_flat = torch.cat(motor)
            pimg, off = R.render_image(motor_flat, ctoken.epsilon, ctoken.blur_sigma, ps)
            assert torch.equal(pimg, model.get_pimg(ctoken))
            torch.manual_seed(5000 + ti)
            image = D.Bernoulli(pimg).sample()
            ip = R.space_motor_to_img(motor_flat.view(-1, 2))
            outm = R.check_bounds(ip, ps.imsize).view(-1, 200)
            fl = torch.floor(ip).long().view(-1, 200, 2)
        ll, g_ctrl, g_inv, g_pos, g_sig, g_eps, g_aff = score_with_grads(model, ctoken, image)
        names.append(name); K.append(len(nsub)); S_tot.append(int(nsub.sum()))
        ctrl_all.append(ctrl); inv_all.append(inv); pos_all.append(pos); nsub_all.append(nsub)
        eps_all.append(float(ctoken.epsilon)); sig_all.append(float(ctoken.blur_sigma))
        aff_all.append(ctoken.affine.detach().numpy() if ctoken.affine is not None else np.array([1, 1, 0, 0], f32))
        has_aff.append(ctoken.affine is not None)
        ll_all.append(ll); off_all.append(bool(off)); img_all.append(pack_image(image))
        gctrl_all.append(g_ctrl); ginv_all.append(g_inv); gpos_all.append(g_pos)
        gsig_all.append(g_sig); geps_all.append(g_eps); gaff_all.append(g_aff)
        mask_all.append(np.packbits(outm.numpy(), axis=1))
        nin_all.append((~outm).sum(1).numpy().astype(np.int32))
        inb = (~outm).unsqueeze(-1)
        fsum_all.append((fl * inb).sum(1).numpy().astype(np.int64))
        if ti < n_full:
            out[f"motor_{ti:03d}"] = motor_flat.numpy().copy()
            out[f"pimg_{ti:03d}"] = pimg.numpy().copy()
    out["names"] = np.array(names)
    out["n_full"] = np.int32(n_full)
    out["K"] = np.array(K, np.int32)
    out["S"] = np.array(S_tot, np.int32)
    out["nsub"] = np.concatenate(nsub_all)
    out["ctrl"] = np.concatenate(ctrl_all).astype(f32)
    out["invscale"] = np.concatenate(inv_all).astype(f32)
    out["position"] = np.concatenate(pos_all).astype(f32)
    out["eps"] = np.array(eps_all, f32)
    out["sigma"] = np.array(sig_all, f32)
    out["affine"] = np.stack(aff_all).astype(f32)
    out["has_affine"] = np.array(has_aff)
    out["ll"] = np.array(ll_all, f32)
    out["off_page"] = np.array(off_all)
    out["image"] = np.stack(img_all)
    out["g_ctrl"] = np.concatenate(gctrl_all).astype(f32)
    out["g_invscale"] = np.concatenate(ginv_all).astype(f32)
    out["g_position"] = np.concatenate(gpos_all).astype(f32)
    out["g_sigma"] = np.array(gsig_all, f32)
    out["g_eps"] = np.array(geps_all, f32)
    out["g_affine"] = np.stack(gaff_all).astype(f32)
    out["mask_out"] = np.concatenate(mask_all)
    out["n_inbounds"] = np.concatenate(nin_all)
    out["floor_sum"] = np.concatenate(fsum_all)
    S.coefficient_mat.cache_clear()
    out["basis_5_200"] = S.coefficient_mat(5, 200).numpy().copy()
    np.savez_compressed(os.path.join(OUT, "tokens.npz"), **out)
    print("tokens.npz:", len(names), "tokens; S hist", np.bincount(out["S"]),
          "off_page frac", np.mean(off_all), "ll range", min(ll_all), max(ll_all))


# --------------------------------------------------------------------------
# fit.npz : fit_bspline_to_traj (pybpl/splines.py:141-171, torch.linalg.lstsq gels)
# --------------------------------------------------------------------------
def make_fit():
    out = {}
    g = torch.Generator().manual_seed(31)
    cases = [(5, 200), (5, 57), (8, 200), (3, 40), (12, 100), (20, 150), (1, 30), (30, 31)]
    out["cases"] = np.array(cases, np.int32)
    for ci, (nland, neval) in enumerate(cases):
        S.coefficient_mat.cache_clear()
        # smooth-ish trajectories: a random spline of 6 control points evaluated at neval points + noise
        Yt = torch.randn(8, 6, 2, generator=g) * 25
        X = torch.stack([S.get_stk_from_bspline(y, neval=neval) for y in Yt]) + 0.3 * torch.randn(8, neval, 2, generator=g)
        Ys, Rs = [], []
        for x in X:
            Yf, res = S.fit_bspline_to_traj(x, nland, include_resid=True)
            Ys.append(Yf.numpy().copy())
            Rs.append(res.numpy().copy() if res.numel() else np.zeros(2, f32))
        out[f"X_{ci}"] = X.numpy().copy()
        out[f"Y_{ci}"] = np.stack(Ys)
        out[f"res_{ci}"] = np.stack(Rs)
        S.coefficient_mat.cache_clear()
        out[f"C_{ci}"] = S.coefficient_mat(nland, neval).numpy().copy()
    # explicit time points
    S.coefficient_mat.cache_clear()
    sv = torch.sort(2 + 4 * torch.rand(90, generator=g))[0]
    Xs = torch.randn(4, 90, 2, generator=g).cumsum(1)
    out["s_explicit"] = sv.numpy().copy()
    out["X_s"] = Xs.numpy().copy()
    ys, rs = [], []
    for x in Xs:
        S.coefficient_mat.cache_clear()
        Yf, res = S.fit_bspline_to_traj(x, 5, s=sv, include_resid=True)
        ys.append(Yf.numpy().copy()); rs.append(res.numpy().copy())
    out["Y_s"] = np.stack(ys); out["res_s"] = np.stack(rs)
    # gradient of the fit w.r.t. the trajectory
    S.coefficient_mat.cache_clear()
    xg = X[0].clone().requires_grad_(True)
    ct = torch.randn(cases[-1][0], 2, generator=g)
    Yg = S.fit_bspline_to_traj(xg, cases[-1][0])
    (Yg * ct).sum().backward()
    out["ct_fit"] = ct.numpy().copy(); out["gX_fit"] = xg.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "fit.npz"), **out)
    print("fit.npz:", len(cases), "cases")


# --------------------------------------------------------------------------
# token_prior.npz : CharacterModel.score_token = log P(token | type)
# (pybpl/model/token_dist.py:57-80,131-153,205-224,264-287,334-455,489-515,542-572;
#  attach points objects/relation.py:34-67) with autograd gradients w.r.t. every
#  token- and type-level parameter (objects/concept.py:93-107,243-258)
# --------------------------------------------------------------------------
REL_CODE = {"unihist": 0, "start": 1, "end": 2, "mid": 3}


def make_token_prior(n=160):
    torch.manual_seed(11)
    lib = Library(use_hist=True)
    model = CharacterModel(lib)
    g = torch.Generator().manual_seed(5)
    K, nsub_all = [], []
    ctrl_ty, inv_ty, ctrl_tk, inv_tk, pos = [], [], [], [], []
    cat, aix, asub, gpos, es_ty, es_tk = [], [], [], [], [], []
    sig, ll_all = [], []
    g_ctrl_ty, g_inv_ty, g_ctrl_tk, g_inv_tk, g_pos, g_gpos, g_es_ty, g_es_tk = ([] for _ in range(8))
    ncat = np.zeros(4, int)
    for i in range(n):
        ctype = model.sample_type()
        ctoken = model.sample_token(ctype)
        if i % 3 == 1:
            ctoken.blur_sigma = (0.5 + 15.0 * torch.rand((), generator=g))
        ctoken.blur_sigma = torch.as_tensor(ctoken.blur_sigma, dtype=torch.float32).clone()
        if i % 17 == 5:      # a non-positive scale: -inf (token_dist.py:404-406); gradients are not taken
            with torch.no_grad():
                ctoken.part_tokens[0].invscales[0] = -0.01
        if i % 19 == 7:      # an evaluation spot outside [2, ncpt+1]: -inf (token_dist.py:563-564)
            for rt in ctoken.relation_tokens:
                if rt.rtype.category == "mid":
                    rt.eval_spot_token = torch.tensor(6.25)
        ctype.train(); ctoken.train()
        # coefficient_mat is lru_cached on the IDENTITY of the tensor s (splines.py:72): sample_token has already
        # evaluated the attach point with this eval_spot_token before it required grad, and a cache hit would hand
        # score_location that graph-less matrix, silently dropping d ll / d eval_spot_token through the attach
        # point.  Cleared, autograd differentiates the function the reference computes.
        S.coefficient_mat.cache_clear()
        ll = model.score_token(ctype, ctoken)
        finite = bool(torch.isfinite(ll))
        if finite:
            ll.backward()
        K.append(int(ctype.k))
        sig.append(float(ctoken.blur_sigma)); ll_all.append(float(ll))
        for pt, rt, ptk, rtk in zip(ctype.part_types, ctype.relation_types, ctoken.part_tokens, ctoken.relation_tokens):
            ns = pt.shapes.shape[2]
            nsub_all.append(ns)
            z = lambda t: torch.zeros_like(t) if (t.grad is None or not finite) else t.grad
            ctrl_ty.append(pt.shapes.detach().permute(2, 0, 1)); g_ctrl_ty.append(z(pt.shapes).permute(2, 0, 1))
            inv_ty.append(pt.invscales.detach()); g_inv_ty.append(z(pt.invscales))
            ctrl_tk.append(ptk.shapes.detach().permute(2, 0, 1)); g_ctrl_tk.append(z(ptk.shapes).permute(2, 0, 1))
            inv_tk.append(ptk.invscales.detach()); g_inv_tk.append(z(ptk.invscales))
            pos.append(ptk.position.detach()); g_pos.append(z(ptk.position))
            c = REL_CODE[rt.category]
            cat.append(c); ncat[c] += 1
            if c == 0:
                gpos.append(rt.gpos.detach()); g_gpos.append(z(rt.gpos))
                aix.append(-1); asub.append(-1)
            else:
                gpos.append(torch.zeros(2)); g_gpos.append(torch.zeros(2))
                aix.append(int(rt.attach_ix)); asub.append(int(rt.attach_subix) if c == 3 else -1)
            if c == 3:
                es_ty.append(float(rt.eval_spot.detach())); g_es_ty.append(float(z(rt.eval_spot)))
                es_tk.append(float(rtk.eval_spot_token.detach())); g_es_tk.append(float(z(rtk.eval_spot_token)))
            else:
                es_ty.append(0.0); g_es_ty.append(0.0); es_tk.append(0.0); g_es_tk.append(0.0)
    cat3 = lambda L: torch.cat(L).numpy().astype(f32)
    st2 = lambda L: torch.stack(L).numpy().astype(f32)
    out = dict(
        K=np.array(K, np.int32), nsub=np.array(nsub_all, np.int32),
        ctrl_type=cat3(ctrl_ty), invscale_type=cat3(inv_ty), ctrl=cat3(ctrl_tk), invscale=cat3(inv_tk), position=st2(pos),
        rel_cat=np.array(cat, np.int32), attach_ix=np.array(aix, np.int32), attach_subix=np.array(asub, np.int32),
        gpos=st2(gpos), eval_spot_type=np.array(es_ty, f32), eval_spot=np.array(es_tk, f32),
        sigma=np.array(sig, f32), ll=np.array(ll_all, f32),
        g_ctrl_type=cat3(g_ctrl_ty), g_invscale_type=cat3(g_inv_ty), g_ctrl=cat3(g_ctrl_tk), g_invscale=cat3(g_inv_tk),
        g_position=st2(g_pos), g_gpos=st2(g_gpos), g_eval_spot_type=np.array(g_es_ty, f32), g_eval_spot=np.array(g_es_tk, f32),
        consts=np.array([float(lib.tokenvar['sigma_shape']), float(lib.tokenvar['sigma_invscale']),
                         float(lib.tokenvar['sigma_attach']), float(lib.rel['sigma_x']), float(lib.rel['sigma_y']),
                         float(model.token_dist.ps.min_blur_sigma), float(model.token_dist.ps.max_blur_sigma)], f32))
    np.savez_compressed(os.path.join(OUT, "token_prior.npz"), **out)
    print("token_prior.npz:", n, "tokens; relation counts (unihist,start,end,mid)", ncat, "; -inf scores",
          int(np.sum(~np.isfinite(out["ll"]))), "; ll range", np.nanmin(out["ll"][np.isfinite(out["ll"])]), out["ll"].max())


# --------------------------------------------------------------------------
# token_samples.npz : CharacterModel.sample_token draws for a few fixed types
# (pybpl/model/token_dist.py:226-262), the distributional fixture for the
# on-device sampler (RNG streams cannot match; distributions must)
# --------------------------------------------------------------------------
def make_token_samples(n_types=4, n_draws=600):
    torch.manual_seed(23)
    lib = Library(use_hist=True)
    model = CharacterModel(lib)
    out = {}
    chosen = []
    want = {3: 2, 2: 1, 1: 1}           # strokes -> how many types (so that attach relations are exercised)
    while sum(want.values()) > 0:
        ct = model.sample_type()
        k = int(ct.k)
        cats = [r.category for r in ct.relation_types]
        if want.get(k, 0) > 0 and (k == 1 or any(c != "unihist" for c in cats)):
            want[k] -= 1; chosen.append(ct)
    for ti, ct in enumerate(chosen):
        K = int(ct.k)
        nsub = np.array([p.shapes.shape[2] for p in ct.part_types], np.int32)
        out[f"t{ti}_nsub"] = nsub
        out[f"t{ti}_ctrl_type"] = torch.cat([p.shapes.permute(2, 0, 1) for p in ct.part_types]).numpy().astype(f32)
        out[f"t{ti}_invscale_type"] = torch.cat([p.invscales for p in ct.part_types]).numpy().astype(f32)
        cat = [REL_CODE[r.category] for r in ct.relation_types]
        out[f"t{ti}_rel_cat"] = np.array(cat, np.int32)
        out[f"t{ti}_attach_ix"] = np.array([int(r.attach_ix) if c else 0 for r, c in zip(ct.relation_types, cat)], np.int32)
        out[f"t{ti}_attach_subix"] = np.array([int(r.attach_subix) if c == 3 else 0 for r, c in zip(ct.relation_types, cat)], np.int32)
        out[f"t{ti}_gpos"] = np.stack([r.gpos.numpy() if c == 0 else np.zeros(2, f32) for r, c in zip(ct.relation_types, cat)]).astype(f32)
        out[f"t{ti}_eval_spot_type"] = np.array([float(r.eval_spot) if c == 3 else 0.0 for r, c in zip(ct.relation_types, cat)], f32)
        ctrl, inv, pos, es = [], [], [], []
        for _ in range(n_draws):
            tk = model.sample_token(ct)
            ctrl.append(torch.cat([p.shapes.permute(2, 0, 1) for p in tk.part_tokens]).numpy())
            inv.append(torch.cat([p.invscales for p in tk.part_tokens]).numpy())
            pos.append(torch.stack([p.position for p in tk.part_tokens]).numpy())
            es.append(np.array([float(r.eval_spot_token) if c == 3 else 0.0 for r, c in zip(tk.relation_tokens, cat)], f32))
        out[f"t{ti}_ctrl"] = np.stack(ctrl).astype(f32)            # [draws, S, 5, 2]
        out[f"t{ti}_invscale"] = np.stack(inv).astype(f32)
        out[f"t{ti}_position"] = np.stack(pos).astype(f32)
        out[f"t{ti}_eval_spot"] = np.stack(es).astype(f32)
    out["n_types"] = np.int32(len(chosen))
    out["consts"] = np.array([float(lib.tokenvar['sigma_shape']), float(lib.tokenvar['sigma_invscale']),
                              float(lib.tokenvar['sigma_attach']), float(lib.rel['sigma_x']), float(lib.rel['sigma_y']),
                              float(model.token_dist.ps.min_blur_sigma), float(model.token_dist.ps.max_blur_sigma)], f32)
    np.savez_compressed(os.path.join(OUT, "token_samples.npz"), **out)
    print("token_samples.npz:", len(chosen), "types x", n_draws, "draws; relations",
          [list(out[f"t{i}_rel_cat"]) for i in range(len(chosen))])


# --------------------------------------------------------------------------
# type_samples.npz : CharacterModel.sample_type draws (pybpl/model/type_dist.py:56-96,
# 148-161, 245-267, 296-326, 355-386, 425-448, 480-505, 541-597), flattened: the
# distributional fixture for the on-device type sampler
# --------------------------------------------------------------------------
def make_type_samples(n_draws=4000):
    torch.manual_seed(31)
    lib = Library(use_hist=True)
    model = CharacterModel(lib)
    K, nsub, rank, kof = [], [], [], []
    subid, inv, ctrl0, sub_first = [], [], [], []
    cat, aix, asub, gpos, es = [], [], [], [], []
    for _ in range(n_draws):
        ct = model.sample_type()
        k = int(ct.k)
        K.append(k)
        for i, (pt, rt) in enumerate(zip(ct.part_types, ct.relation_types)):
            nsub.append(int(pt.nsub)); rank.append(i); kof.append(k)
            subid.append(pt.ids.numpy().astype(np.int32))
            inv.append(pt.invscales.numpy())
            ctrl0.append(pt.shapes.permute(2, 0, 1).reshape(-1, 10).numpy())
            sub_first.append(np.arange(int(pt.nsub)) == 0)
            c = REL_CODE[rt.category]
            cat.append(c)
            aix.append(int(rt.attach_ix) if c else -1)
            asub.append(int(rt.attach_subix) if c == 3 else -1)
            gpos.append(rt.gpos.numpy() if c == 0 else np.full(2, np.nan, f32))
            es.append(float(rt.eval_spot) if c == 3 else np.nan)
    out = dict(K=np.array(K, np.int32), nsub=np.array(nsub, np.int32), rank=np.array(rank, np.int32), k_of_stroke=np.array(kof, np.int32),
               subid=np.concatenate(subid), invscale=np.concatenate(inv).astype(f32), ctrl=np.concatenate(ctrl0).astype(f32),
               sub_first=np.concatenate(sub_first), rel_cat=np.array(cat, np.int32), attach_ix=np.array(aix, np.int32),
               attach_subix=np.array(asub, np.int32), gpos=np.stack(gpos).astype(f32), eval_spot=np.array(es, f32))
    np.savez_compressed(os.path.join(OUT, "type_samples.npz"), **out)
    print("type_samples.npz:", n_draws, "types;", len(nsub), "strokes;", len(out["subid"]), "sub-strokes; K hist",
          np.bincount(out["K"])[1:], "; relation counts", np.bincount(out["rel_cat"]))


# --------------------------------------------------------------------------
# type_prior.npz : CharacterModel.score_type = log P(type)
# (pybpl/model/type_dist.py:98-125, 163-185, 269-294, 328-353, 388-423, 450-478,
#  507-531, 599-650; position histograms library/spatial_OLD/spatial_model.py:86-112,
#  spatial_hist.py:145-167,235-256) with autograd gradients w.r.t. shapes and invscales
#  (gpos is detached: the histogram score calls .numpy() on it and raises when it requires grad)
# --------------------------------------------------------------------------
def make_type_prior(n=300):
    torch.manual_seed(41)
    lib = Library(use_hist=True)
    model = CharacterModel(lib)
    K, nsub, subid, ctrl, inv, cat, aix, asub, gpos, es, ll_all, g_ctrl, g_inv = ([] for _ in range(13))
    for i in range(n):
        ct = model.sample_type()
        if i % 23 == 7:          # a position outside the histogram's range: -inf (spatial_hist.py:252-253)
            for r in ct.relation_types:
                if r.category == "unihist":
                    r.gpos = torch.tensor([120.0, -10.0])
        for p in ct.part_types:
            p.shapes.requires_grad_(True); p.invscales.requires_grad_(True)
        ll = model.score_type(ct)
        finite = bool(torch.isfinite(ll))
        if finite:
            ll.backward()
        K.append(int(ct.k)); ll_all.append(float(ll))
        for p, r in zip(ct.part_types, ct.relation_types):
            nsub.append(int(p.nsub)); subid.append(p.ids.numpy().astype(np.int32))
            ctrl.append(p.shapes.detach().permute(2, 0, 1)); inv.append(p.invscales.detach())
            z = lambda t: torch.zeros_like(t) if (t.grad is None or not finite) else t.grad
            g_ctrl.append(z(p.shapes).permute(2, 0, 1)); g_inv.append(z(p.invscales))
            c = REL_CODE[r.category]; cat.append(c)
            aix.append(int(r.attach_ix) if c else 0); asub.append(int(r.attach_subix) if c == 3 else 0)
            gpos.append(r.gpos.detach() if c == 0 else torch.zeros(2)); es.append(float(r.eval_spot) if c == 3 else 0.0)
    out = dict(K=np.array(K, np.int32), nsub=np.array(nsub, np.int32), subid=np.concatenate(subid),
               ctrl_type=torch.cat(ctrl).numpy().astype(f32), invscale_type=torch.cat(inv).numpy().astype(f32),
               rel_cat=np.array(cat, np.int32), attach_ix=np.array(aix, np.int32), attach_subix=np.array(asub, np.int32),
               gpos=torch.stack(gpos).numpy().astype(f32), eval_spot_type=np.array(es, f32), ll=np.array(ll_all, f32),
               g_ctrl_type=torch.cat(g_ctrl).numpy().astype(f32), g_invscale_type=torch.cat(g_inv).numpy().astype(f32))
    np.savez_compressed(os.path.join(OUT, "type_prior.npz"), **out)
    print("type_prior.npz:", n, "types; -inf scores", int(np.sum(~np.isfinite(out["ll"]))), "; ll range",
          out["ll"][np.isfinite(out["ll"])].min(), out["ll"].max())


def make_assets():
    """pybpl_b200/data/basis_5_200.npy: coefficient_mat(5, 200) exactly as the reference's torch builds it
    (splines.py:72-93).  torch assembles it with Sleef powf and an MKL dot, which portable C cannot
    reproduce to the bit, and bit-exact integer work (floor/ceil/in-bounds) needs bit-exact trajectories,
    so the product ships this one table as data; every other (nland, neval) is computed by the library."""
    S.coefficient_mat.cache_clear()
    C = S.coefficient_mat(5, 200).numpy().astype(f32)
    dst = os.path.join(os.path.dirname(os.path.dirname(OUT)), "pybpl_b200", "data")
    os.makedirs(dst, exist_ok=True)
    np.save(os.path.join(dst, "basis_5_200.npy"), C)
    print("asset basis_5_200.npy", C.shape)
    # Hyper-parameter tables the synthetic workload generator draws from: data, not code.
    lib = Library(use_hist=True)
    Sig = lib.shape['Sigma'].numpy()
    np.savez_compressed(os.path.join(dst, "workload_lib.npz"),
                        mu=lib.shape['mu'].numpy().astype(f32),
                        sd=np.sqrt(np.stack([np.diag(Sig[i]) for i in range(Sig.shape[0])])).astype(f32),
                        theta=lib.scale['theta'].numpy().astype(f32),
                        pmat_nsub=lib.pmat_nsub.numpy().astype(f32), pkappa=lib.pkappa.numpy().astype(f32),
                        sigma_shape=f32(lib.tokenvar['sigma_shape']), sigma_invscale=f32(lib.tokenvar['sigma_invscale']),
                        rel_sigma=np.array([float(lib.rel['sigma_x']), float(lib.rel['sigma_y'])], f32))
    print("asset workload_lib.npz")
    # Tables of the type prior (pybpl/library/library.py:30-72, model/type_dist.py:137-146, 224-243, 535-539) in the
    # form the on-device type sampler reads them: categorical distributions as cumulative sums, shape covariances
    # as Cholesky factors, the position histograms in the linear bin order SpatialHist.sample draws from
    # (library/spatial_OLD/spatial_hist.py:107-143).  Data, not code.
    def cdf(p):
        p = np.asarray(p, np.float64)
        c = np.cumsum(p / p.sum(-1, keepdims=True), -1)
        c[..., -1] = 1.0
        return c.astype(f32)
    logT = lib.logT.numpy().astype(np.float64)
    SM = lib.Spatial
    sp = []
    for h in SM.list_SH:
        sp.append(np.exp(h.logpYX.numpy().astype(np.float64)).T.reshape(-1))       # lin = xi * ny + yi
    np.savez_compressed(os.path.join(dst, "type_lib.npz"),
                        pkappa=cdf(lib.pkappa.numpy()), pmat_nsub=cdf(lib.pmat_nsub.numpy()),
                        cdf_start=cdf(np.exp(lib.logStart.numpy().astype(np.float64))), cdf_T=cdf(np.exp(logT)),
                        shape_mu=lib.shape['mu'].numpy().astype(f32),
                        shape_L=np.linalg.cholesky(lib.shape['Sigma'].numpy().astype(np.float64)).astype(f32),
                        gamma_con=lib.scale['theta'][:, 0].numpy().astype(f32),
                        gamma_rate=(1.0 / lib.scale['theta'][:, 1].numpy().astype(np.float64)).astype(f32),
                        rel_mixprob=cdf(lib.rel['mixprob'].numpy()), sp_cdf=cdf(np.stack(sp)),
                        sp_xlab=SM.list_SH[0].xlab.numpy().astype(f32), sp_ylab=SM.list_SH[0].ylab.numpy().astype(f32),
                        # log-probability tables for the type score (type_dist.py:163-185, 269-294, 328-353, 599-650)
                        logp_kappa=torch.distributions.Categorical(probs=lib.pkappa).logits.numpy().astype(f32),
                        logp_nsub=torch.distributions.Categorical(probs=lib.pmat_nsub).logits.numpy().astype(f32),
                        logp_start=torch.distributions.Categorical(probs=torch.exp(lib.logStart)).logits.numpy().astype(f32),
                        logp_T=(logT - np.log(np.exp(logT).sum(1, keepdims=True))).astype(f32),
                        logp_mix=torch.distributions.Categorical(probs=lib.rel['mixprob']).logits.numpy().astype(f32),
                        sp_logp=np.stack([h.logpYX.numpy().T.reshape(-1) for h in SM.list_SH]).astype(f32),
                        sp_log_binarea=f32(float(torch.log(SM.list_SH[0].rg_bin[0]) + torch.log(SM.list_SH[0].rg_bin[1]))),
                        # the same distributions as plain probabilities, for tests and inspection
                        p_kappa=(lib.pkappa / lib.pkappa.sum()).numpy().astype(f32),
                        p_nsub=(lib.pmat_nsub / lib.pmat_nsub.sum(1, keepdim=True)).numpy().astype(f32),
                        p_start=np.exp(lib.logStart.numpy().astype(np.float64)).astype(f32))
    print("asset type_lib.npz", os.path.getsize(os.path.join(dst, "type_lib.npz")) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["splines", "render", "tokens", "fit", "token_prior", "token_samples", "type_samples", "type_prior", "assets"]
    if "splines" in which:
        make_splines()
    if "render" in which:
        make_render()
    if "tokens" in which:
        make_tokens()
    if "fit" in which:
        make_fit()
    if "token_prior" in which:
        make_token_prior()
    if "token_samples" in which:
        make_token_samples()
    if "type_samples" in which:
        make_type_samples()
    if "type_prior" in which:
        make_type_prior()
    if "assets" in which:
        make_assets()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
