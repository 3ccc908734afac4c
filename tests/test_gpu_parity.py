"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors generated from the
live reference and against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer work bit-exact; fp32 pimg / ll within 1e-5 relative;
gradients per tensor in inf-norm-relative terms, stated below next to the reference's own fp32 noise
(the reference's autograd is 7e-5 (points) .. 5e-3 (d/d sigma) from fp64)."""
import os

import numpy as np
import pytest
import torch

from tests.util_golden import load, render_cases, token_cases, pack_tokens, relinf

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from pybpl_b200 import ops as _ops
    return _ops


@pytest.fixture(params=["canvas", "teams"])
def bwd_kernel(request):
    """Runs a gradient test once per fused backward kernel: the full-canvas kernel (small batches by default) and the
    team kernel with box-sized arenas (bpl_backward_teams.cuh; batches of 2048 tokens and more by default)."""
    import os
    old = {k: os.environ.pop(k, None) for k in ("BPL_BWD_CANVAS", "BPL_BWD_TEAMS")}
    os.environ["BPL_BWD_CANVAS" if request.param == "canvas" else "BPL_BWD_TEAMS"] = "1"
    yield request.param
    for k, v in old.items():
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = v


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def make_token_batch(ops, toks, grad=False, with_affine=None):
    b = pack_tokens(toks)
    use_aff = bool(b["has_affine"].any()) if with_affine is None else with_affine
    aff = b["affine"].copy()
    aff[~b["has_affine"]] = [1, 1, 0, 0]
    ctrl, inv, pos = T(b["ctrl"]), T(b["invscale"]), T(b["position"])
    eps, sig = T(b["eps"]), T(b["sigma"])
    affine = T(aff) if use_aff else None
    if grad:
        for t in (ctrl, inv, pos, eps, sig):
            t.requires_grad_(True)
        if affine is not None:
            affine.requires_grad_(True)
    batch = ops.TokenBatch(b["tok_stroke_off"], b["stroke_sub_off"], ctrl, inv, pos, eps, sig, affine=affine)
    imgs = ops.pack_images(T(b["images"], torch.uint8))
    return batch, imgs, T(b["img_idx"], torch.int32), b


# ---------------------------------------------------------------------------------------------
# splines
# ---------------------------------------------------------------------------------------------
def test_spline_eval_bit_exact(ops):
    d = load("splines")
    X = ops.spline_eval(T(d["table_5_200"]), T(d["Y_explicit"]))
    assert np.array_equal(X.cpu().numpy(), d["X_200"])
    X7 = ops.spline_eval(T(d["table_7_200"]), T(d["Y7"]))
    assert np.array_equal(X7.cpu().numpy(), d["X7_200"])
    Xe = ops.spline_eval(T(d["table_s_explicit"]), T(d["Y_explicit"]))
    assert np.array_equal(Xe.cpu().numpy(), d["X_explicit"])


def test_shipped_table_is_the_references(ops):
    d = load("splines")
    assert np.array_equal(ops.basis_table(5, 200, DEV).cpu().numpy(), d["table_5_200"])
    for nland, neval in d["table_pairs"]:
        C = ops.basis_table(int(nland), int(neval), DEV).cpu().numpy()
        assert np.abs(C - d[f"table_{nland}_{neval}"]).max() < 1e-6


def test_spline_eval_backward(ops):
    d = load("splines")
    Y = T(d["Y_explicit"]).requires_grad_(True)
    X = ops.spline_eval(T(d["table_5_200"]), Y)
    (X * T(d["ct_200"])).sum().backward()
    assert relinf(Y.grad.cpu().numpy(), d["gY_200"]) < 2e-6


def test_adaptive_neval_integer_exact(ops):
    d = load("splines")
    nev, dist = ops.adaptive_neval(T(d["Y_adaptive"]))
    assert np.array_equal(nev.cpu().numpy(), d["neval_adaptive"])
    assert relinf(dist.cpu().numpy(), d["dist_adaptive"]) < 2e-6


def test_get_stk_from_bspline_dropin():
    from pybpl_b200 import splines
    d = load("splines")
    Y = torch.from_numpy(d["Y_adaptive"][:8])
    outs = [splines.get_stk_from_bspline(y) for y in Y]              # adaptive neval, CPU tensors in/out
    assert [o.shape[0] for o in outs] == list(d["neval_adaptive"][:8])
    assert all(o.device.type == "cpu" for o in outs)
    got = torch.cat(outs).numpy()
    assert np.abs(got - d["X_adaptive_first8"]).max() < 2e-4          # non-(5,200) tables: ~4e-7 coefficient diffs
    Ye = torch.from_numpy(d["Y_explicit"][0])
    X = splines.bspline_eval(torch.tensor(4.5), Ye)
    assert X.shape == (1, 2) and np.abs(X.numpy() - d["X_scalar_s"][0]).max() < 1e-4
    X200 = splines.get_stk_from_bspline(Ye, neval=200)
    assert np.array_equal(X200.numpy(), d["X_200"][0])


def test_bspline_eval_gradient_in_s_and_Y():
    """BSplineEval's autograd: d X / d s through the analytic derivative of the normalised basis and d X / d Y,
    against central differences of the float64 oracle row (oracle/token_prior.py:basis_row)."""
    from oracle.token_prior import basis_row
    from pybpl_b200 import splines
    rng = np.random.default_rng(0)
    Y = rng.normal(0, 10, (5, 2))
    w = rng.normal(0, 1, (3, 2))
    sv = np.array([2.4, 3.77, 5.6])
    st = torch.tensor(sv, dtype=torch.float32, requires_grad=True)
    Yt = torch.tensor(Y, dtype=torch.float32, requires_grad=True)
    X = splines.bspline_eval(st, Yt)
    (X * torch.tensor(w, dtype=torch.float32)).sum().backward()
    C = np.stack([basis_row(u, 5, np.float64)[0] for u in sv]); dC = np.stack([basis_row(u, 5, np.float64)[1] for u in sv])
    assert np.abs(X.detach().numpy() - C @ Y).max() < 1e-4
    assert np.abs(st.grad.numpy() - ((dC @ Y) * w).sum(1)).max() < 1e-4 * np.abs((dC @ Y) * w).max()
    assert np.abs(Yt.grad.numpy() - C.T @ w).max() < 1e-5


# ---------------------------------------------------------------------------------------------
# vanilla_to_motor + integer work
# ---------------------------------------------------------------------------------------------
def test_vanilla_to_motor_bit_exact(ops):
    d, toks = token_cases()
    from pybpl_b200 import objects
    for t in toks[: int(d["n_full"])]:
        if t["has_affine"]:
            continue
        s0 = 0
        for k, n in enumerate(t["nsub"]):
            shapes = torch.from_numpy(t["ctrl"][s0:s0 + n]).permute(1, 2, 0).contiguous()   # (ncpt,2,nsub)
            motor, ms = objects.vanilla_to_motor(shapes, torch.from_numpy(t["invscale"][s0:s0 + n]),
                                                 torch.from_numpy(t["position"][k]))
            assert motor.shape == (n, 200, 2) and ms.shape == (5, 2, n)    # the reference test's own assertion
            assert np.array_equal(motor.numpy(), t["motor"][s0:s0 + n]), t["name"]
            s0 += n


def test_integer_work_bit_exact(ops):
    """In-bounds masks, surviving-point counts and pixel indices, from the fused kernels' own device code."""
    d, toks = token_cases()
    sel = [t for t in toks if not t["has_affine"]]
    batch, _, _, b = make_token_batch(ops, sel, with_affine=False)
    r = ops.token_points(batch)
    mask = r["mask_out"].cpu().numpy()
    fl = r["floor"].cpu().numpy().astype(np.int64)
    n_in = r["n_in"].cpu().numpy()
    s0 = 0
    for t in sel:
        S = len(t["invscale"])
        m = mask[s0:s0 + S]
        assert np.array_equal(m, t["mask_out"]), t["name"]
        assert np.array_equal(n_in[s0:s0 + S], t["n_inbounds"])
        fsum = (fl[s0:s0 + S] * (~m)[..., None]).sum(1)
        assert np.array_equal(fsum, t["floor_sum"])
        if "motor" in t:
            assert np.array_equal(r["motor"][s0:s0 + S].cpu().numpy(), t["motor"])
        s0 += S


# ---------------------------------------------------------------------------------------------
# render_image on raw trajectories (edge cases)
# ---------------------------------------------------------------------------------------------
class _PS:
    def __init__(self, hinton):
        self.broaden_mode = "Hinton" if hinton else "Lake"


@pytest.mark.parametrize("case", render_cases(), ids=lambda c: c["name"])
def test_render_image_cases(ops, case):
    from pybpl_b200 import rendering
    strokes = [torch.from_numpy(s.copy()).requires_grad_(True) for s in case["strokes"]]
    eps = torch.tensor(case["eps"], requires_grad=case["eps"] > 0)
    sig = torch.tensor(case["sigma"], requires_grad=case["sigma"] > 0)
    pimg, off = rendering.render_image(strokes, eps, sig, _PS(case["hinton"]))
    assert pimg.shape == (105, 105) and pimg.dtype == torch.float32 and pimg.device.type == "cpu"
    assert bool(off) == case["off"]
    ref = case["pimg"]
    err = np.abs(pimg.detach().numpy() - ref).max()
    assert err <= 1e-5 * max(1.0, np.abs(ref).max()), err
    # generic cotangent on pimg (render_image's autograd)
    (pimg * torch.from_numpy(case["ct"])).sum().backward()
    g = torch.cat([s.grad for s in strokes]).numpy()
    assert relinf(g, case["gct_pts"]) < 2e-4
    if case["eps"] > 0:
        assert abs(eps.grad.item() - case["gct_eps"]) <= 2e-4 * max(1.0, abs(case["gct_eps"]))
    if case["sigma"] > 0:
        assert abs(sig.grad.item() - case["gct_sigma"]) <= 5e-3 * max(1.0, abs(case["gct_sigma"]))


@pytest.mark.parametrize("case", render_cases(), ids=lambda c: c["name"])
def test_render_and_score_cases(ops, case):
    """Raw trajectories scored in-kernel: ll and d ll / d (points, eps, sigma)."""
    lens = case["lens"]
    off = np.zeros(len(lens) + 1, np.int32); off[1:] = np.cumsum(lens)
    pts = T(case["pts"]).requires_grad_(True)
    eps = T([case["eps"]]).requires_grad_(True)
    sig = T([case["sigma"]]).requires_grad_(True)
    batch = ops.StrokeBatch([0, len(lens)], off, pts, eps, sig)
    imgs = ops.pack_images(T(case["image"], torch.uint8))
    _, ll, offp = ops.render_strokes(batch, images=imgs, want_pimg=False, ps=_PS(case["hinton"]))
    assert bool(offp[0]) == case["off"]
    assert abs(ll.item() - case["ll"]) <= 1e-5 * abs(case["ll"]), (ll.item(), case["ll"])
    ll.sum().backward()
    assert relinf(pts.grad.cpu().numpy(), case["g_pts"]) < 2e-4
    assert abs(eps.grad.item() - case["g_eps"]) <= 5e-4 * max(1.0, abs(case["g_eps"]))
    assert abs(sig.grad.item() - case["g_sigma"]) <= 5e-3 * max(1.0, abs(case["g_sigma"]))


# ---------------------------------------------------------------------------------------------
# whole path on tokens (CharacterImageDist.get_pimg / score_image)
# ---------------------------------------------------------------------------------------------
def test_tokens_forward(ops):
    d, toks = token_cases()
    batch, imgs, idx, b = make_token_batch(ops, toks)
    ll, off = ops.score_tokens(batch, imgs, idx)
    ll = ll.cpu().numpy(); off = off.cpu().numpy()
    ref = np.array([t["ll"] for t in toks])
    assert np.array_equal(off, np.array([t["off_page"] for t in toks]))
    rel = np.abs(ll - ref) / np.abs(ref)
    print("tokens ll max rel", rel.max())
    assert rel.max() < 1e-5
    pimg, off2 = ops.render_tokens(batch)
    pimg = pimg.cpu().numpy()
    for i, t in enumerate(toks):
        if "pimg" in t:
            assert np.abs(pimg[i] - t["pimg"]).max() < 1e-5, t["name"]


def test_tokens_gradients(ops, bwd_kernel):
    d, toks = token_cases()
    batch, imgs, idx, b = make_token_batch(ops, toks, grad=True)
    ll, _ = ops.score_tokens(batch, imgs, idx)
    ll.sum().backward()
    gc = batch.ctrl.grad.cpu().numpy(); gi = batch.invscale.grad.cpu().numpy(); gp = batch.position.grad.cpu().numpy()
    ge = batch.eps.grad.cpu().numpy(); gs = batch.sigma.grad.cpu().numpy(); ga = batch.affine.grad.cpu().numpy()
    ref_ll = np.array([t["ll"] for t in toks])
    assert (np.abs(ll.detach().cpu().numpy() - ref_ll) / np.abs(ref_ll)).max() < 1e-5
    worst = dict(gc=0, gi=0, gp=0, gs=0, ge=0, ga=0)
    s0 = k0 = 0
    for i, t in enumerate(toks):
        S, K = len(t["invscale"]), len(t["nsub"])
        worst["gc"] = max(worst["gc"], relinf(gc[s0:s0 + S], t["g_ctrl"]))
        worst["gi"] = max(worst["gi"], relinf(gi[s0:s0 + S], t["g_invscale"]))
        worst["gp"] = max(worst["gp"], relinf(gp[k0:k0 + K], t["g_position"]))
        worst["gs"] = max(worst["gs"], abs(gs[i] - t["g_sigma"]) / max(1.0, abs(t["g_sigma"])))
        worst["ge"] = max(worst["ge"], abs(ge[i] - t["g_eps"]) / max(1.0, abs(t["g_eps"])))
        if t["has_affine"]:
            worst["ga"] = max(worst["ga"], relinf(ga[i], t["g_affine"]))
        s0 += S; k0 += K
    print("worst grad errors vs reference fp32 autograd", worst)
    assert worst["gc"] < 5e-4 and worst["gi"] < 5e-4 and worst["gp"] < 5e-4
    assert worst["gs"] < 1e-2 and worst["ge"] < 2e-3 and worst["ga"] < 2e-3


def test_gradients_no_worse_than_reference_vs_fp64(ops, bwd_kernel):
    """Per gradient tensor, the distance to the fp64 oracle: max over tokens of |g_cuda - g_fp64| must not exceed the same
    figure for the reference's own fp32 autograd, max |g_ref - g_fp64| (factor 1.0, plus an absolute slack of 1e-6 of
    the tensor's scale; 1.5 for epsilon, see below), and the median token is within 5 % of the reference's -- the build
    is as accurate as the reference for every leaf: control points, scales, positions, blur_sigma, epsilon."""
    from oracle import Oracle
    d, toks = token_cases()
    sel = [t for t in toks if not t["has_affine"]][:24]
    batch, imgs, idx, b = make_token_batch(ops, sel, grad=True, with_affine=False)
    ll, _ = ops.score_tokens(batch, imgs, idx)
    ll.sum().backward()
    got = dict(ctrl=batch.ctrl.grad.cpu().numpy(), invscale=batch.invscale.grad.cpu().numpy(),
               position=batch.position.grad.cpu().numpy(), sigma=batch.sigma.grad.cpu().numpy(), eps=batch.eps.grad.cpu().numpy())
    o64 = Oracle(np.float64)
    C64 = d["basis_5_200"].astype(np.float64)
    s0 = k0 = 0
    ours = {k: [] for k in got}; refs = {k: [] for k in got}
    for i, t in enumerate(sel):
        S, K = len(t["invscale"]), len(t["nsub"])
        r = o64.token(C64, t["nsub"], t["ctrl"], t["invscale"], t["position"], t["eps"], t["sigma"],
                      image=t["image"], grad=True)
        for name, mine, ref, tru in (("ctrl", got["ctrl"][s0:s0 + S], t["g_ctrl"], r["g_ctrl"]),
                                     ("invscale", got["invscale"][s0:s0 + S], t["g_invscale"], r["g_invscale"]),
                                     ("position", got["position"][k0:k0 + K], t["g_position"], r["g_position"]),
                                     ("sigma", got["sigma"][i:i + 1], np.atleast_1d(t["g_sigma"]), np.atleast_1d(r["g_sigma"])),
                                     ("eps", got["eps"][i:i + 1], np.atleast_1d(t["g_eps"]), np.atleast_1d(r["g_eps"]))):
            scale = max(np.abs(tru).max(), 1.0 if name in ("sigma", "eps") else 1e-30)
            ours[name].append(np.abs(mine - tru).max() / scale); refs[name].append(np.abs(np.asarray(ref) - tru).max() / scale)
        s0 += S; k0 += K
    for name in got:
        print("g_%s vs fp64 (%s kernel): ours max %.2e median %.2e | reference max %.2e median %.2e" %
              (name, bwd_kernel, max(ours[name]), np.median(ours[name]), max(refs[name]), np.median(refs[name])))
    # factor 1.0 on the worst token for every leaf but epsilon: its worst token is one whose target has zeros under
    # p ~ 0.9996 ink, where d lp / d p = -1 / (1 - p) turns the 6e-8 rounding of the fp32 probability itself into
    # 1.5e-4 per term -- conditioning both implementations share; the kernel sums those terms in double and is the
    # more accurate of the two on the median token (1.7e-7 vs 3.9e-7), on the worst one it may land either side.
    for name in got:
        factor = 1.5 if name == "eps" else 1.0
        assert max(ours[name]) <= factor * max(refs[name]) + 1e-6, (name, max(ours[name]), max(refs[name]))
        assert np.median(ours[name]) <= 1.05 * np.median(refs[name]) + 1e-7, (name, np.median(ours[name]), np.median(refs[name]))


def test_score_matches_oracle_on_random_batch(ops, bwd_kernel):
    """Seeded synthetic batch (the bench workload generator) against the CPU oracle, forward and backward."""
    from oracle import Oracle
    from pybpl_b200 import workload
    w = workload.synth_tokens(96, seed=1234, sigma_mode="uniform")
    o = Oracle(np.float32)
    C = load("tokens")["basis_5_200"]
    import oracle as _oracle
    img = _oracle.target_image_for(w, o, C)
    ref = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"],
                        w["eps"], w["sigma"], img[None], grad=True)
    ctrl, inv, pos = T(w["ctrl"]).requires_grad_(True), T(w["invscale"]).requires_grad_(True), T(w["position"]).requires_grad_(True)
    eps, sig = T(w["eps"]).requires_grad_(True), T(w["sigma"]).requires_grad_(True)
    batch = ops.TokenBatch(w["tok_stroke_off"], w["stroke_sub_off"], ctrl, inv, pos, eps, sig)
    imgs = ops.pack_images(T(img, torch.uint8))
    ll, off = ops.score_tokens(batch, imgs)
    ll.sum().backward()
    assert np.array_equal(off.cpu().numpy(), ref["off_page"])
    rel = np.abs(ll.detach().cpu().numpy() - ref["ll"]) / np.abs(ref["ll"])
    assert rel.max() < 1e-5, rel.max()
    assert relinf(ctrl.grad.cpu().numpy(), ref["g_ctrl"]) < 2e-4
    assert relinf(inv.grad.cpu().numpy(), ref["g_invscale"]) < 2e-4
    assert relinf(pos.grad.cpu().numpy(), ref["g_position"]) < 2e-4
    assert relinf(eps.grad.cpu().numpy(), ref["g_eps"]) < 5e-4
    assert relinf(sig.grad.cpu().numpy(), ref["g_sigma"]) < 5e-3


def test_other_neval_against_oracle(ops, bwd_kernel):
    """The fused kernels take any number of evaluation points per sub-stroke (neval * 5 <= 1024 table entries), not only
    the reference's default 200 (objects/part.py:290): 64 points, scoring kernel and both backward kernels against the
    oracle with the basis table coefficient_mat(5, 64)."""
    from oracle import Oracle
    from pybpl_b200 import workload
    import oracle as _oracle
    w = workload.synth_tokens(80, seed=31, sigma_mode="uniform", neval=64)
    o = Oracle(np.float32)
    C = ops.basis_table(5, 64, DEV).cpu().numpy()
    img = _oracle.target_image_for(w, o, C)
    ref = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"],
                        w["eps"], w["sigma"], img[None], grad=True)
    imgs = ops.pack_images(T(img, torch.uint8))
    leaves = [T(w["ctrl"]), T(w["invscale"]), T(w["position"]), T(w["eps"]), T(w["sigma"])]
    with torch.no_grad():
        ll0, off0 = ops.score_tokens(ops.TokenBatch(w["tok_stroke_off"], w["stroke_sub_off"], *leaves, neval=64), imgs)
    for x in leaves:
        x.requires_grad_(True)
    ll, off = ops.score_tokens(ops.TokenBatch(w["tok_stroke_off"], w["stroke_sub_off"], *leaves, neval=64), imgs)
    g = torch.autograd.grad(ll.sum(), leaves)
    for got in (ll0, ll.detach()):
        assert (np.abs(got.cpu().numpy() - ref["ll"]) / np.abs(ref["ll"])).max() < 1e-5
    assert np.array_equal(off0.cpu().numpy(), ref["off_page"]) and np.array_equal(off.cpu().numpy(), ref["off_page"])
    # Gradients against the fp64 oracle: with 64 points the segments are three times longer, fewer and larger terms carry
    # the gradient, and the fp32 ORACLE's own noise grows to 4e-4 ... 7e-4 of fp64 (measured), the kernels stay below 1e-4.
    r64 = Oracle(np.float64).token_batch(C.astype(np.float64), w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"],
                                         w["position"], w["eps"], w["sigma"], img[None], grad=True)
    errs = [relinf(x.cpu().numpy(), r64[k]) for x, k in zip(g, ("g_ctrl", "g_invscale", "g_position", "g_eps", "g_sigma"))]
    print("neval 64,", bwd_kernel, "kernel: gradient errors vs fp64 oracle (ctrl, invscale, position, eps, sigma)", errs)
    assert errs[0] < 2e-4 and errs[1] < 2e-4 and errs[2] < 2e-4 and errs[3] < 5e-4 and errs[4] < 5e-3, errs


def test_smaller_gaussian_filter(ops, bwd_kernel):
    """Parameters.fsize (pybpl/parameters.py:41) below its default 11: the 11-tap passes run with zero outer taps.
    render_image against the live reference's goldens (fsize 7, 5, 3: probability maps and autograd gradients,
    tests/golden/fsize.npz); the token kernels (scoring, both backward kernels) against the oracle at fsize = 7."""
    from types import SimpleNamespace as NS
    from oracle import Oracle
    import oracle as _oracle
    from pybpl_b200 import rendering, workload
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fsize.npz"))
    for fsize in (7, 5, 3):
        ps = NS(imsize=(105, 105), fsize=fsize, ink_ncon=2, ink_pp=2.0, ink_max_dist=2.0, ink_a=0.5, ink_b=6.0, broaden_mode="Lake")
        for i in range(int(d["n_sets"])):
            key = f"f{fsize}_s{i}"
            st = [torch.from_numpy(d[f"{key}_stroke{j}"].copy()).requires_grad_(True) for j in range(int(d[key + "_n"]))]
            eps = torch.tensor(float(d[key + "_eps"]), requires_grad=True); sig = torch.tensor(float(d[key + "_sigma"]), requires_grad=True)
            pimg, off = rendering.render_image(st, eps, sig, ps)
            assert bool(off) == bool(d[key + "_off"])
            assert (pimg.detach().numpy() - d[key + "_pimg"]).__abs__().max() < 1e-5, (fsize, i)
            (pimg * torch.from_numpy(d[key + "_ct"])).sum().backward()
            for j, s in enumerate(st):
                assert relinf(s.grad.numpy(), d[f"{key}_g_stroke{j}"]) < 5e-4, (fsize, i, j)
            assert abs(sig.grad.item() - float(d[key + "_g_sigma"])) <= 1e-2 * max(1.0, abs(float(d[key + "_g_sigma"])))
            assert abs(eps.grad.item() - float(d[key + "_g_eps"])) <= 2e-3 * max(1.0, abs(float(d[key + "_g_eps"])))
    # tokens at fsize = 7 against the oracle
    w = workload.synth_tokens(64, seed=41, sigma_mode="uniform")
    ops_ps = ops.make_params(NS(imsize=(105, 105), fsize=7))
    ops7 = _oracle.default_params(); ops7.fsize = 7
    C = ops.basis_table(5, 200, DEV).cpu().numpy()
    o = Oracle(np.float32, ops7)
    img = _oracle.target_image_for(w, o, C)
    # (the fp32 oracle: it takes the same branch decisions as the kernels -- one token of this batch sits on a
    #  discontinuity where fp64 arithmetic decides differently, at any filter size)
    ref = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"], w["eps"], w["sigma"],
                        img[None], grad=True)
    imgs = ops.pack_images(T(img, torch.uint8))
    leaves = [T(w["ctrl"]), T(w["invscale"]), T(w["position"]), T(w["eps"]), T(w["sigma"])]
    with torch.no_grad():
        ll0, _ = ops.score_tokens(ops.TokenBatch(w["tok_stroke_off"], w["stroke_sub_off"], *leaves), imgs, ps=ops_ps)
    for x in leaves:
        x.requires_grad_(True)
    ll, _ = ops.score_tokens(ops.TokenBatch(w["tok_stroke_off"], w["stroke_sub_off"], *leaves), imgs, ps=ops_ps)
    g = torch.autograd.grad(ll.sum(), leaves)
    for got in (ll0, ll.detach()):
        assert (np.abs(got.cpu().numpy() - ref["ll"]) / np.abs(ref["ll"])).max() < 1e-5
    errs = [relinf(x.cpu().numpy(), ref[k]) for x, k in zip(g, ("g_ctrl", "g_invscale", "g_position", "g_eps", "g_sigma"))]
    print("fsize 7,", bwd_kernel, "kernel: gradient errors vs the fp32 oracle", errs)
    assert errs[0] < 2e-4 and errs[1] < 2e-4 and errs[2] < 2e-4 and errs[3] < 5e-4 and errs[4] < 5e-3, errs
    # and a filter the kernels are not built for raises
    with pytest.raises(NotImplementedError):
        ops.make_params(NS(imsize=(105, 105), fsize=13))


def test_backward_kernels_agree_on_every_mode(ops):
    """The two fused backward kernels (full canvas / teams with box-sized arenas) on one mixed batch: affine warps,
    blur_sigma <= 0, epsilon = 0, a token without strokes, a token that is entirely off the page, weighted scores
    (g_ll) and the probability-map cotangent (get_pimg's backward).  Same algorithm, different summation orders."""
    import os
    from pybpl_b200 import workload
    w = workload.synth_tokens(160, seed=77, sigma_mode="uniform")
    rng = np.random.default_rng(5)
    w["sigma"][::7] = 0.0
    w["sigma"][3::11] = -1.0
    w["eps"][::5] = 0.0
    w["position"][w["tok_stroke_off"][10]:w["tok_stroke_off"][11]] += 400.0      # token 10: off the page
    aff = np.tile(np.array([1, 1, 0, 0], np.float32), (160, 1))
    aff[::3] = np.stack([rng.uniform(0.8, 1.25, 54), rng.uniform(0.8, 1.25, 54), rng.uniform(-6, 6, 54), rng.uniform(-6, 6, 54)], 1)
    # a token without strokes in front
    tso = np.concatenate([[0], w["tok_stroke_off"]]).astype(np.int64)
    eps = np.concatenate([[1e-4], w["eps"]]).astype(np.float32); sig = np.concatenate([[2.0], w["sigma"]]).astype(np.float32)
    aff = np.concatenate([[[1, 1, 0, 0]], aff]).astype(np.float32)
    P = 161
    img = (rng.random((3, 105, 105)) < 0.04).astype(np.uint8)
    imgs = ops.pack_images(T(img, torch.uint8))
    idx = T(rng.integers(0, 3, P), torch.int32)
    g_ll = T(rng.standard_normal(P))
    g_pimg = T(rng.standard_normal((P, 105, 105)) * (rng.random((P, 105, 105)) < 0.3))
    res = {}
    for kern in ("BPL_BWD_CANVAS", "BPL_BWD_TEAMS"):
        os.environ.pop("BPL_BWD_CANVAS", None); os.environ.pop("BPL_BWD_TEAMS", None)
        os.environ[kern] = "1"
        try:
            leaves = [T(w["ctrl"]), T(w["invscale"]), T(w["position"]), T(aff), T(eps), T(sig)]
            for t in leaves:
                t.requires_grad_(True)
            batch = ops.TokenBatch(tso, w["stroke_sub_off"], leaves[0], leaves[1], leaves[2], leaves[4], leaves[5], affine=leaves[3])
            ll, off = ops.score_tokens(batch, imgs, idx)
            g1 = torch.autograd.grad((ll * g_ll).sum(), leaves)
            pimg, off2 = ops.render_tokens(batch)
            g2 = torch.autograd.grad((pimg * g_pimg).sum(), leaves)
            res[kern] = (ll.detach().cpu().numpy(), off.cpu().numpy(), [g.cpu().numpy() for g in g1], [g.cpu().numpy() for g in g2])
        finally:
            os.environ.pop(kern, None)
    a, b = res["BPL_BWD_CANVAS"], res["BPL_BWD_TEAMS"]
    assert np.array_equal(a[1], b[1]) and bool(a[1][11]) and not bool(a[1][0])
    assert np.allclose(a[0], b[0], rtol=2e-6, atol=0)
    names = ("ctrl", "invscale", "position", "affine", "eps", "sigma")
    sso = w["stroke_sub_off"]
    tok_of = {"ctrl": np.repeat(np.arange(P), np.diff(sso[tso])), "invscale": np.repeat(np.arange(P), np.diff(sso[tso])),
              "position": np.repeat(np.arange(P), np.diff(tso)), "affine": np.arange(P), "eps": np.arange(P), "sigma": np.arange(P)}
    for which, ga, gb in (("score", a[2], b[2]), ("pimg", a[3], b[3])):
        for n, x, y in zip(names, ga, gb):
            assert np.isfinite(x).all() and np.isfinite(y).all(), (which, n)
            scale = max(np.abs(x).max(), 1e-30)
            per_tok = np.zeros(P)
            np.maximum.at(per_tok, tok_of[n], np.abs(x - y).reshape(x.shape[0], -1).max(1) / scale)
            print(which, n, "canvas vs teams inf-norm rel: worst token", per_tok.max(), "tokens above 1e-4:", int((per_tok > 1e-4).sum()))
            # Same algorithm, different summation orders: half the bar of the oracle tests.  The splat accumulates with
            # floating-point atomics, so a pixel's ink can differ by an ulp between two runs of even the SAME kernel; a
            # pixel that sits exactly on the clamp_(max=1) of rendering.py:171 then flips its mask and moves the
            # gradient of its token discontinuously (measured: one blur-less token of this batch, 4e-4, in about one run
            # in ten; tools/flaky_probe.py).  Such tokens may exceed the bar, by a bounded amount, and must stay rare.
            assert (per_tok > 1e-4).sum() <= 2 and per_tok.max() < 5e-3, (which, n, per_tok.max())
    # the token without strokes and the off-page token get exactly zero stroke gradients from both
    assert not a[2][0][w["stroke_sub_off"][w["tok_stroke_off"][10]]:w["stroke_sub_off"][w["tok_stroke_off"][11]]].any()


# ---------------------------------------------------------------------------------------------
# mirror of the reference API
# ---------------------------------------------------------------------------------------------
def test_character_model_api(ops):
    from pybpl_b200.model import CharacterModel
    from pybpl_b200.objects import CharacterToken, StrokeToken
    d, toks = token_cases()
    model = CharacterModel()
    for t in toks[:6]:
        parts, s0 = [], 0
        for k, n in enumerate(t["nsub"]):
            shapes = torch.from_numpy(t["ctrl"][s0:s0 + n]).permute(1, 2, 0).contiguous().requires_grad_(True)
            parts.append(StrokeToken(shapes, torch.from_numpy(t["invscale"][s0:s0 + n]).requires_grad_(True),
                                     torch.from_numpy(t["position"][k]).requires_grad_(True)))
            s0 += n
        ctok = CharacterToken(parts, torch.from_numpy(t["affine"]) if t["has_affine"] else None,
                              torch.tensor(t["eps"]), torch.tensor(t["sigma"], requires_grad=True))
        image = torch.from_numpy(t["image"]).float()
        ll = model.score_image(ctok, image)
        assert ll.dim() == 0 and ll.device.type == "cpu"
        assert abs(ll.item() - t["ll"]) <= 1e-5 * abs(t["ll"])
        ll.backward()
        g = torch.cat([p.shapes.grad.permute(2, 0, 1) for p in parts]).numpy()
        assert relinf(g, t["g_ctrl"]) < 5e-4
        assert abs(ctok.blur_sigma.grad.item() - t["g_sigma"]) <= 1e-2 * max(1.0, abs(t["g_sigma"]))
        if "pimg" in t:
            pimg = model.get_pimg(ctok)
            assert np.abs(pimg.detach().numpy() - t["pimg"]).max() < 1e-5
        torch.manual_seed(0)
        img = model.sample_image(ctok)
        assert img.shape == (105, 105) and set(np.unique(img.numpy())) <= {0.0, 1.0}
    with pytest.raises(ValueError):
        model.score_image(ctok, image * 0.5)        # torch's Bernoulli support check


def test_sample_image_statistics(ops):
    from pybpl_b200.model import CharacterModel
    from pybpl_b200.objects import CharacterToken, StrokeToken
    d, toks = token_cases()
    t = toks[1]
    parts, s0 = [], 0
    for k, n in enumerate(t["nsub"]):
        parts.append(StrokeToken(torch.from_numpy(t["ctrl"][s0:s0 + n]).permute(1, 2, 0).contiguous(),
                                 torch.from_numpy(t["invscale"][s0:s0 + n]), torch.from_numpy(t["position"][k])))
        s0 += n
    ctok = CharacterToken(parts, None, t["eps"], t["sigma"])
    model = CharacterModel()
    torch.manual_seed(3)
    mean = torch.stack([model.sample_image(ctok) for _ in range(64)]).mean(0).numpy()
    assert abs(mean.mean() - t["pimg"].mean()) < 0.01


# ---------------------------------------------------------------------------------------------
# size-independent properties at the bench size
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(ops):
    from pybpl_b200 import workload
    P = 4096
    w = workload.synth_tokens(P, seed=1234, sigma_mode="uniform")
    batch = workload.to_device(w, DEV)
    img = torch.zeros((105, 105), dtype=torch.uint8, device=DEV)
    imgs0 = ops.pack_images(img)
    imgs1 = ops.pack_images(1 - img)
    ll0, off0 = ops.score_tokens(batch, imgs0)
    ll1, _ = ops.score_tokens(batch, imgs1)
    # determinism of everything but the atomics' summation order
    ll0b, off0b = ops.score_tokens(batch, imgs0)
    assert torch.equal(off0, off0b)
    assert ((ll0 - ll0b).abs() <= 2e-5 * ll0.abs()).all()
    # permutation equivariance: scoring a permuted batch permutes the scores
    perm = torch.randperm(P, generator=torch.Generator().manual_seed(0))
    wp = workload.permute(w, perm.numpy())
    llp, offp = ops.score_tokens(workload.to_device(wp, DEV), imgs0)
    assert torch.equal(offp.cpu(), off0.cpu()[perm])
    assert ((llp.cpu() - ll0.cpu()[perm]).abs() <= 2e-5 * ll0.cpu()[perm].abs()).all()
    # ll(all-zero image) + ll(all-one image) = sum over pixels of [log p + log(1-p)]: check against pimg
    pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(64)), DEV))
    p = pimg.double().clamp(1.1920929e-7, 1 - 1.1920929e-7)
    want = (p.log() + (-p).log1p()).sum((1, 2))
    got = (ll0[:64] + ll1[:64]).double()
    assert ((got - want).abs() <= 2e-5 * want.abs()).all()
    assert torch.isfinite(ll0).all() and (ll0 < 0).all()


def test_multi_image_indexing(ops):
    from pybpl_b200 import workload
    w = workload.synth_tokens(64, seed=5, sigma_mode="fixed")
    batch = workload.to_device(w, DEV)
    g = torch.Generator().manual_seed(1)
    imgs = (torch.rand((5, 105, 105), generator=g) < 0.1).to(torch.uint8).to(DEV)
    packed = ops.pack_images(imgs)
    idx = torch.arange(64, dtype=torch.int32, device=DEV) % 5
    ll, _ = ops.score_tokens(batch, packed, idx)
    for t in range(5):
        llt, _ = ops.score_tokens(batch, ops.pack_images(imgs[t]))
        sel = (idx == t)
        assert ((ll[sel] - llt[sel]).abs() <= 2e-5 * llt[sel].abs()).all()


def test_empty_and_ragged(ops):
    # empty batch
    z = torch.zeros
    batch = ops.TokenBatch([0], [0], z((0, 5, 2), device=DEV), z(0, device=DEV), z((0, 2), device=DEV),
                           z(0, device=DEV), z(0, device=DEV))
    imgs = ops.pack_images(z((105, 105), dtype=torch.uint8, device=DEV))
    ll, off = ops.score_tokens(batch, imgs)
    assert ll.numel() == 0 and off.numel() == 0
    # a token with zero strokes renders the empty image: ll = 11025 * lp(eps, 0)
    batch = ops.TokenBatch([0, 0], [0], z((0, 5, 2), device=DEV), z(0, device=DEV), z((0, 2), device=DEV),
                           torch.tensor([1e-4], device=DEV), torch.tensor([0.5], device=DEV))
    ll, off = ops.score_tokens(batch, imgs)
    want = torch.distributions.Bernoulli(probs=torch.full((105, 105), 1e-4)).log_prob(torch.zeros(105, 105)).sum()
    assert abs(ll.item() - want.item()) <= 1e-5 * abs(want.item())
    assert not bool(off[0])


def test_logsumexp_partial(ops):
    g = torch.Generator().manual_seed(0)
    ll = (-2000 * torch.rand(100000, generator=g) - 50).to(DEV)
    out = ops.logsumexp_partial(ll)
    got = out[0].item() + np.log(out[1].item())
    assert abs(got - torch.logsumexp(ll.double(), 0).item()) < 1e-4


def test_library_fails_loudly_without_cuda_tensors(ops):
    from pybpl_b200 import _lib
    with pytest.raises(_lib.BplError):
        ops.TokenBatch([0, 0], [0], torch.zeros((0, 5, 2)), torch.zeros(0), torch.zeros((0, 2)), torch.zeros(1),
                       torch.zeros(1))


def test_fit_bspline_to_traj(ops):
    """fit_bspline_to_traj (pybpl/splines.py:141-171) against the reference's lstsq outputs and its autograd."""
    from pybpl_b200 import splines
    d = load("fit")
    for ci, (nl, ne) in enumerate(d["cases"]):
        C = T(d[f"C_{ci}"])
        Y, res = ops.spline_fit(C, T(d[f"X_{ci}"]))
        assert relinf(Y.cpu().numpy(), d[f"Y_{ci}"]) < 2e-5, (nl, ne)
        if ne > nl + 1:
            assert relinf(res.cpu().numpy(), d[f"res_{ci}"]) < 2e-4, (nl, ne)
    # drop-in signature: CPU tensor in, CPU tensors out, explicit time points, residuals
    X = torch.from_numpy(d["X_s"][0])
    Y, res = splines.fit_bspline_to_traj(X, 5, s=torch.from_numpy(d["s_explicit"]), include_resid=True)
    assert Y.shape == (5, 2) and res.shape == (2,) and Y.device.type == "cpu"
    assert relinf(Y.numpy(), d["Y_s"][0]) < 5e-5 and relinf(res.numpy(), d["res_s"][0]) < 5e-4
    Y0 = splines.fit_bspline_to_traj(torch.from_numpy(d["X_0"][0]), 5)
    assert relinf(Y0.numpy(), d["Y_0"][0]) < 2e-5
    _, r_empty = splines.fit_bspline_to_traj(torch.randn(5, 2), 5, include_resid=True)
    assert r_empty.numel() == 0                      # torch.linalg.lstsq returns no residuals when m <= n
    # gradient w.r.t. the trajectory
    last = len(d["cases"]) - 1
    xg = torch.from_numpy(d[f"X_{last}"][0]).requires_grad_(True)
    Yg = splines.fit_bspline_to_traj(xg, int(d["cases"][last][0]))
    (Yg * torch.from_numpy(d["ct_fit"])).sum().backward()
    assert relinf(xg.grad.numpy(), d["gX_fit"]) < 1e-4


def test_fit_smooth_stk_matches_reference(ops):
    """fit_smooth_stk (pybpl/bottomup/initialize/util.py:13-31), the adaptive refit loop over fit_bspline_to_traj: same
    number of control points chosen (an integer decision: same output length through the adaptive neval) and the same
    smooth trajectory as the reference on the golden strokes (tests/golden/fit_smooth.npz; inputs already resampled by
    the reference's unif_space, which is host-side data preparation outside this package)."""
    from pybpl_b200 import splines
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_smooth.npz"))
    for i in range(int(d["n"])):
        ref = d[f"out_{i}"]
        got = splines.fit_smooth_stk(d[f"in_{i}"], unif_space=lambda s, u=d[f"unif_{i}"]: u)
        got = np.asarray(got)
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        err = np.abs(got - ref).max()
        print("fit_smooth_stk stroke", i, "points", ref.shape[0], "max abs difference", err)
        assert err < 2e-4          # pixels (measured 3e-5: the fp32 least-squares solutions of gels and the Householder QR)


def test_token_order_is_a_permutation_by_decreasing_cost(ops):
    """bpl_order_tokens: the processing order handed to the persistent kernels (results do not depend on it)."""
    from pybpl_b200 import workload
    for P in (1, 2, 777, 4096, 16384):
        w = workload.synth_tokens(P, seed=5, sigma_mode="uniform")
        b = workload.to_device(w, DEV)
        if b.order_ptr() is None:
            assert P < 2
            continue
        order = b._order.cpu().numpy()
        assert np.array_equal(np.sort(order), np.arange(P))
        S = np.diff(w["stroke_sub_off"][w["tok_stroke_off"]])
        assert np.all(np.diff(S[order]) <= 0)
    # scores with and without the order agree (to the summation order of the splat's shared-memory atomics)
    w = workload.synth_tokens(512, seed=6, sigma_mode="uniform")
    b = workload.to_device(w, DEV)
    pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), DEV))
    imgs = ops.pack_images((pimg[0] > 0.5).to(torch.uint8))
    ll1, _ = ops.score_tokens(b, imgs)
    b._order = b._order_cost = torch.arange(512, dtype=torch.int32, device=DEV)
    ll2, _ = ops.score_tokens(b, imgs)
    assert torch.allclose(ll1, ll2, rtol=2e-6, atol=0)


def test_team_backward_with_arenas_that_do_not_fit_side_by_side(ops):
    """Tokens whose boxes cover the whole canvas and that carry many sub-strokes need 3 x 105 x 105 floats + their
    per-sub-stroke sums each: two such arenas exceed the shared-memory pool, so one team of every CTA has to wait for the
    other to finish (the `waiting` hand-over of the arena allocator).  Scores and gradients must equal the full-canvas
    kernel's; a token beyond the per-CTA limits (129 sub-strokes) is flagged NaN / 255 by both."""
    import os
    rng = np.random.default_rng(11)
    P, K, NS = 40, 10, 10                                   # 100 sub-strokes per token, spread over the canvas
    S = P * K * NS
    ctrl = np.zeros((S, 5, 2), np.float32)
    t = np.linspace(0, 1, 5)
    for s in range(S):
        a, b = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        ctrl[s] = 12.0 * (np.outer(t, a) + np.outer(t * t, b))
    inv = rng.uniform(0.5, 1.5, S).astype(np.float32)
    pos = np.stack([rng.uniform(5, 100, P * K), -rng.uniform(5, 100, P * K)], 1).astype(np.float32)
    tso = np.arange(0, P * K + 1, K); sso = np.arange(0, S + 1, NS)
    # one token more with 129 sub-strokes in a single stroke (beyond BPL_MAX_SUBSTROKES)
    ctrl = np.concatenate([ctrl, ctrl[:129]]); inv = np.concatenate([inv, inv[:129]])
    pos = np.concatenate([pos, pos[:1]]); tso = np.concatenate([tso, [P * K + 1]]); sso = np.concatenate([sso, [S + 129]])
    eps = np.full(P + 1, 1e-4, np.float32); sig = rng.uniform(0.5, 8.0, P + 1).astype(np.float32)
    img = (rng.random((105, 105)) < 0.2).astype(np.uint8)
    imgs = ops.pack_images(T(img, torch.uint8))
    res = {}
    for kern in ("BPL_BWD_CANVAS", "BPL_BWD_TEAMS"):
        os.environ.pop("BPL_BWD_CANVAS", None); os.environ.pop("BPL_BWD_TEAMS", None)
        os.environ[kern] = "1"
        try:
            leaves = [T(ctrl), T(inv), T(pos), T(eps), T(sig)]
            for x in leaves:
                x.requires_grad_(True)
            batch = ops.TokenBatch(tso, sso, *leaves, check_limits=False)
            ll, off = ops.score_tokens(batch, imgs)
            g = torch.autograd.grad(ll[:P].sum(), leaves)
            res[kern] = (ll.detach().cpu().numpy(), off.cpu().numpy(), [x.cpu().numpy() for x in g])
        finally:
            os.environ.pop(kern, None)
    a, b = res["BPL_BWD_CANVAS"], res["BPL_BWD_TEAMS"]
    for r in (a, b):
        assert np.isnan(r[0][P]) and r[1].view(np.uint8)[P] != 0 and np.isfinite(r[0][:P]).all()
    assert np.allclose(a[0][:P], b[0][:P], rtol=2e-6, atol=0) and np.array_equal(a[1][:P], b[1][:P])
    for n, x, y in zip(("ctrl", "invscale", "position", "eps", "sigma"), a[2], b[2]):
        err = np.abs(x - y).max() / np.abs(x).max()
        print("full-canvas tokens:", n, "canvas vs teams", err)
        assert np.isfinite(y).all() and err < 1e-4, (n, err)


def test_work_queue_under_concurrent_streams_and_graph_replay(ops):
    """Every launch of a persistent kernel owns one slot of the work-queue ring (g_sched) until its last CTA resets it.
    Two streams launching back to back at the same time, and a CUDA graph replayed while live launches run on another
    stream, must give exactly the scores of serial launches (a shared or stale slot would skip or repeat tokens)."""
    from pybpl_b200 import workload
    wa = workload.synth_tokens(3000, seed=21, sigma_mode="uniform")
    wb = workload.synth_tokens(2500, seed=22, sigma_mode="uniform")
    ba, bb = workload.to_device(wa, DEV), workload.to_device(wb, DEV)
    pimg, _ = ops.render_tokens(workload.to_device(workload.permute(wa, np.arange(1)), DEV))
    imgs = ops.pack_images((pimg[0] > 0.5).to(torch.uint8))
    with torch.no_grad():
        ra, rb = ops.score_tokens(ba, imgs)[0].clone(), ops.score_tokens(bb, imgs)[0].clone()
        torch.cuda.synchronize()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        outs_a, outs_b = [], []
        for _ in range(12):
            with torch.cuda.stream(s1):
                outs_a.append(ops.score_tokens(ba, imgs)[0])
            with torch.cuda.stream(s2):
                outs_b.append(ops.score_tokens(bb, imgs)[0])
        torch.cuda.synchronize()
        for o in outs_a:
            assert torch.allclose(o, ra, rtol=2e-6, atol=0)
        for o in outs_b:
            assert torch.allclose(o, rb, rtol=2e-6, atol=0)
        # a captured launch replayed on s1 while live launches run on s2
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s1):
            ops.score_tokens(ba, imgs)                       # warm-up on the capture stream
            with torch.cuda.graph(g, stream=s1):
                ga = ops.score_tokens(ba, imgs)[0]
        torch.cuda.synchronize()
        for _ in range(8):
            with torch.cuda.stream(s1):
                g.replay()
            with torch.cuda.stream(s2):
                ob = ops.score_tokens(bb, imgs)[0]
            torch.cuda.synchronize()
            assert torch.allclose(ga, ra, rtol=2e-6, atol=0) and torch.allclose(ob, rb, rtol=2e-6, atol=0)
            ga.zero_()


def test_cost_order_for_the_fused_backward(ops):
    """bpl_order_tokens_by_cost: a permutation, non-increasing in its own cost classes, and the classes track the true ink
    box (the box of the control polygons bounds it); gradients do not depend on the order."""
    from pybpl_b200 import workload
    P = 1500
    w = workload.synth_tokens(P, seed=9, sigma_mode="uniform")
    b = workload.to_device(w, DEV, requires_grad=True)
    cs = b.c_struct(b.ctrl.detach(), b.invscale.detach(), b.position.detach(), None, b.eps.detach(), b.sigma.detach(), cost_order=True)
    assert cs.order is not None and b._order_cost is not None
    order = b._order_cost.cpu().numpy(); keys = b._keys.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(P))
    assert np.all(np.diff(keys[order]) >= 0)                     # class 0 = most expensive, first
    pts = ops.token_points(workload.to_device(w, DEV))
    mask = pts["mask_out"].cpu().numpy().astype(bool); fl = pts["floor"].cpu().numpy(); ce = pts["ceil"].cpu().numpy()
    tso, sso = w["tok_stroke_off"], w["stroke_sub_off"]
    area = np.zeros(P)
    for p in range(P):
        s0, s1 = sso[tso[p]], sso[tso[p + 1]]
        m = ~mask[s0:s1].reshape(-1)
        if m.any():
            f = fl[s0:s1].reshape(-1, 2)[m]; c = ce[s0:s1].reshape(-1, 2)[m]
            h = min(104, c[:, 0].max() + 12) - max(0, f[:, 0].min() - 12) + 1
            wd = min(104, c[:, 1].max() + 12) - max(0, f[:, 1].min() - 12) + 1
            area[p] = h * wd
    S = np.diff(sso[tso])
    cost = 0.34 * area / 4060.0 + 0.17 * S / 5.2
    r = np.corrcoef(cost, -keys.astype(np.float64))[0, 1]
    print("corr(true-box cost, cost class)", r)
    assert r > 0.9
    pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), DEV))
    imgs = ops.pack_images((pimg[0] > 0.5).to(torch.uint8))
    leaves = [b.ctrl, b.invscale, b.position, b.eps, b.sigma]
    g1 = torch.autograd.grad(ops.score_tokens(b, imgs)[0].sum(), leaves)
    b._order_cost = torch.arange(P, dtype=torch.int32, device=DEV)
    g2 = torch.autograd.grad(ops.score_tokens(b, imgs)[0].sum(), leaves)
    for x, y in zip(g1, g2):
        assert (x - y).abs().max() <= 2e-5 * x.abs().max()


def test_cycle_feedback_orders_the_next_backward(ops):
    """A batch that has been through the fused backward (team kernel) carries every token's measured cycles; the next
    evaluation is ordered by them (bpl_order_tokens_by_cycles): a permutation, most expensive class first, and the
    gradients are those of the first evaluation."""
    import os
    from pybpl_b200 import workload
    os.environ.pop("BPL_BWD_CANVAS", None); os.environ.pop("BPL_BWD_TEAMS", None)
    P = 1200                                               # >= 1024: the team kernel, and the feedback buffer is kept
    w = workload.synth_tokens(P, seed=13, sigma_mode=10.0)
    b = workload.to_device(w, DEV, requires_grad=True)
    pimg, _ = ops.render_tokens(workload.to_device(workload.permute(w, np.arange(1)), DEV))
    imgs = ops.pack_images((pimg[0] > 0.5).to(torch.uint8))
    leaves = [b.ctrl, b.invscale, b.position, b.eps, b.sigma]
    g1 = torch.autograd.grad(ops.score_tokens(b, imgs)[0].sum(), leaves)
    assert b._cycles_valid
    first_order = b._order_cost.clone()
    cyc = b._cycles.cpu().numpy().astype(np.int64)
    assert (cyc > 0).all()                                 # every token reported
    g2 = torch.autograd.grad(ops.score_tokens(b, imgs)[0].sum(), leaves)
    order = b._order_cost.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(P))
    assert np.all(np.diff(cyc[order] >> 11) <= 0)          # 2 048-cycle classes, non-increasing
    assert not torch.equal(first_order, b._order_cost)     # (the estimate's order has been replaced)
    for x, y in zip(g1, g2):
        assert (x - y).abs().max() <= 2e-5 * x.abs().max()


# ---------------------------------------------------------------------------------------------
# the benchmarked kernel (bpl_forward1s_kernel<false>, forward only) on the bench workloads, against the oracle
# ---------------------------------------------------------------------------------------------
def _forward_vs_oracle(ops, w, img, sel=None, tag=""):
    """Scores all of `w` WITHOUT gradients (the box-confined scoring kernel) and compares the tokens `sel` with the fp32
    and fp64 oracles.  Bars: ink_off_page bit-exact; ll within 1e-5 relative of the fp32 oracle for every token scored
    against an image it could have produced, and -- for all tokens, including those scored against somebody else's
    image, where log(1 - p) at p ~ 0.9996 turns ulps of p into 1e-4 of ll -- no further from the fp64 oracle than the
    reference's own fp32 arithmetic is (tests/test_gpu_configs.py::test_config5 explains the conditioning)."""
    from oracle import Oracle
    from pybpl_b200 import workload
    batch = workload.to_device(w, DEV)
    imgs = ops.pack_images(T(img, torch.uint8))
    with torch.no_grad():
        ll, off = ops.score_tokens(batch, imgs)
    assert not ll.requires_grad
    ll = ll.cpu().numpy(); off = off.cpu().numpy()
    sub = w if sel is None else workload.permute(w, sel)
    if sel is not None:
        ll, off = ll[sel], off[sel]
    C = load("tokens")["basis_5_200"]
    args = (sub["tok_stroke_off"], sub["stroke_sub_off"], sub["ctrl"], sub["invscale"], sub["position"], sub["eps"], sub["sigma"], img[None])
    r32 = Oracle(np.float32).token_batch(C, *args)
    r64 = Oracle(np.float64).token_batch(C.astype(np.float64), *args)
    assert np.array_equal(off, r32["off_page"])
    l32, l64 = r32["ll"].astype(np.float64), r64["ll"]
    e_gpu32 = np.abs(ll - l32) / np.abs(l32)
    e_gpu64 = np.abs(ll - l64) / np.abs(l64)
    e_ref64 = np.abs(l32 - l64) / np.abs(l64)
    print(f"{tag}: {len(ll)} tokens; vs fp32 oracle max {e_gpu32.max():.2e} (within 1e-5: {(e_gpu32 <= 1e-5).mean():.4f}); "
          f"vs fp64 max {e_gpu64.max():.2e} median {np.median(e_gpu64):.2e} | fp32 oracle vs fp64 max {e_ref64.max():.2e} "
          f"median {np.median(e_ref64):.2e}")
    assert e_gpu64.max() <= max(e_ref64.max(), 1e-5)
    assert np.median(e_gpu64) <= max(np.median(e_ref64), 2e-6)
    assert np.percentile(e_gpu64, 99) <= max(np.percentile(e_ref64, 99), 1e-5)
    return ll, e_gpu32


def test_forward_kernel_on_the_full_config1_batch(ops):
    """configs[1] exactly as bench.py's config1 object runs it: synth_tokens(4096, 1234, 'uniform'), the target sampled
    from token 0, forward only."""
    from oracle import Oracle
    import oracle as _oracle
    from pybpl_b200 import workload
    w = workload.synth_tokens(4096, seed=1234, sigma_mode="uniform")
    C = load("tokens")["basis_5_200"]
    img = _oracle.target_image_for(w, Oracle(np.float32), C)
    ll, e32 = _forward_vs_oracle(ops, w, img, tag="config1 batch")
    assert e32[0] <= 1e-5                         # token 0 against its own sample


def test_forward_kernel_on_a_sweep_shard(ops):
    """configs[3]: the 8-rank shard of the particle sweep (125,000 particles = sweep blocks 0..7) in one launch, 8,192
    of them checked against the oracle; the fused log-sum-exp partial against torch.logsumexp of the scores."""
    from oracle import Oracle
    import oracle as _oracle
    from pybpl_b200 import sweep, workload
    w = workload.sweep_shard(0, 8)
    assert w["P"] == 125000
    C = load("tokens")["basis_5_200"]
    img = _oracle.target_image_for(w, Oracle(np.float32), C)
    sel = np.sort(np.random.default_rng(3).choice(w["P"], 8192, replace=False))
    ll_sel, _ = _forward_vs_oracle(ops, w, img, sel=sel, tag="sweep shard")
    sw = sweep.Sweep(workload.to_device(w, DEV), ops.pack_images(T(img, torch.uint8)))
    total = sw.step()
    assert ((sw.ll.cpu().numpy()[sel] - ll_sel) == 0).all() or np.allclose(sw.ll.cpu().numpy()[sel], ll_sel, rtol=2e-5)
    want = torch.logsumexp(sw.ll.double(), 0).item()
    assert abs(total.item() - want) <= 1e-4 * max(1.0, abs(want)), (total.item(), want)
    # chunked accumulation gives the same partial; a second step (queue slot reuse) the same total
    st = ops.LseState(DEV)
    for c, lo in enumerate(range(0, w["P"], 40000)):
        wc = workload.permute(w, np.arange(lo, min(w["P"], lo + 40000)))
        ops.score_tokens_lse(workload.to_device(wc, DEV), sw.images, st, accumulate=c > 0, want_ll=False)
    assert abs(ops.logsumexp_combine(st.partial).item() - want) <= 1e-4 * max(1.0, abs(want))
    assert abs(sw.step().item() - want) <= 1e-4 * max(1.0, abs(want))


def test_logsumexp_partial_and_combine(ops):
    g = torch.Generator().manual_seed(5)
    for n in (1, 31, 4097, 300001):
        x = (torch.randn(n, generator=g) * 50 - 3000).to(DEV)
        if n > 31:
            x[7] = float("-inf"); x[11] = float("nan")
        part = ops.logsumexp_partial(x)
        ok = torch.isfinite(x)
        want = torch.logsumexp(x[ok].double(), 0).item()
        got = ops.logsumexp_combine(part).item()
        assert abs(got - want) <= 1e-5 * abs(want), (n, got, want)
        assert part[0].item() == x[ok].max().item()
    empty = ops.logsumexp_combine(torch.tensor([float("-inf"), 0.0], device=DEV))
    assert empty.item() == float("-inf")


def test_one_column_box_through_the_scoring_kernel(ops):
    """A token whose ink falls on ONE pixel column (control points with x = 0 exactly, so every trajectory point has the
    stroke position's integer x) with ink_ncon = 0 and blur_sigma = 0: nothing widens the box, and the scoring kernel's
    flattened tail runs on a box one pixel wide (it takes a background column along).  Scores against the oracle; a second
    token hugs the canvas's first column."""
    from types import SimpleNamespace as NS
    from oracle import Oracle
    import oracle as _oracle
    rng = np.random.default_rng(11)
    P = 3
    ctrl = np.zeros((P, 5, 2), np.float32)
    ctrl[:, :, 1] = -np.sort(rng.uniform(0, 60, (P, 5)), axis=1)            # straight down, x stays 0
    w = dict(P=P, tok_stroke_off=np.arange(P + 1, dtype=np.int32), stroke_sub_off=np.arange(P + 1, dtype=np.int32),
             ctrl=ctrl, invscale=np.ones(P, np.float32), position=np.array([[40.0, -10.0], [0.0, -20.0], [104.0, -5.0]], np.float32),
             eps=np.full(P, 1e-3, np.float32), sigma=np.zeros(P, np.float32))
    ops_ps = _oracle.default_params(); ops_ps.ink_ncon = 0
    o = Oracle(np.float32, ops_ps)
    C = load("tokens")["basis_5_200"]
    img = (rng.random((105, 105)) < 0.05).astype(np.uint8)
    img[10:60, 40] = 1
    ref = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"], w["eps"],
                        w["sigma"], img[None], want_pimg=True)
    assert [(ref["pimg"][p] > 2e-3).any(0).sum() for p in range(P)] == [1, 1, 1]      # one ink column each
    ps = ops.make_params(NS(imsize=(105, 105), fsize=11, ink_ncon=0))
    batch = ops.TokenBatch(w["tok_stroke_off"], w["stroke_sub_off"], T(w["ctrl"]), T(w["invscale"]), T(w["position"]),
                           T(w["eps"]), T(w["sigma"]))
    imgs = ops.pack_images(T(img, torch.uint8))
    with torch.no_grad():
        ll, off = ops.score_tokens(batch, imgs, ps=ps)
        ll2, _ = ops.score_tokens(batch, imgs, ps=ps)                           # the buffer is clean again after the first pass
        ll_all = ops.score_tokens_all_images(batch, imgs, ps=ps)[0][:, 0]
    assert np.array_equal(off.cpu().numpy(), ref["off_page"])
    for got in (ll, ll2, ll_all):
        rel = np.abs(got.cpu().numpy() - ref["ll"]) / np.abs(ref["ll"])
        assert rel.max() < 1e-5, rel
