"""BASELINE.json configs[0] under the REAL reference objects: `pybpl_b200.dropin.install()` over the unmodified
rfeinman/pyBPL installed in baseline/_ref (baseline/install_ref.sh), compared in-process with the un-patched reference.

    Library(use_hist=True), torch.manual_seed(0), 1 sampled type, 4 tokens, sample_image + score_image  (CPU reference,
    examples/generate_character.py:21-30, pybpl/model/model.py:32-39)

The reference computes on the CPU with its own torch code; the drop-in computes on cuda:0 through libpybpl_b200.so and
hands CPU tensors back (the reference objects live on the CPU).  Tolerances are the north-star's: pimg <= 1e-5
(absolute, values in [0,1]), ll <= 1e-5 relative; gradients to the per-tensor bars of DESIGN.md section 2 (the
reference's own fp32 autograd noise against fp64)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.fail("baseline/_ref is missing (run baseline/install_ref.sh in the build container; it ships with gpurun)")
    pybpl = ref_loader.load(grad_patch=True)
    from pybpl.library import Library
    from pybpl.model import CharacterModel
    torch.manual_seed(0)
    lib = Library(use_hist=True)
    model = CharacterModel(lib)
    ctype = model.sample_type()
    tokens = [model.sample_token(ctype) for _ in range(4)]
    images = [model.sample_image(t) for t in tokens]
    return dict(pybpl=pybpl, lib=lib, model=model, ctype=ctype, tokens=tokens, images=images)


def _score_with_grads(model, tok, img):
    """score_image + autograd gradients for the leaves the fit_image loop optimises (concept.py:243-258); sample_token
    leaves blur_sigma a Python float (token_dist.py:253-254), so it is made a tensor for the duration."""
    sigma0 = tok.blur_sigma
    tok.blur_sigma = torch.tensor(float(sigma0))
    leaves = []
    for p in tok.part_tokens:
        leaves += [p.shapes, p.invscales, p.position]
    leaves.append(tok.blur_sigma)
    try:
        for x in leaves:
            x.requires_grad_(True)
            x.grad = None
        ll = model.score_image(tok, img)
        ll.backward()
        g = [x.grad.detach().clone() for x in leaves]
    finally:
        for x in leaves:
            x.requires_grad_(False)
            x.grad = None
        tok.blur_sigma = sigma0
    return ll.detach(), g


def _relinf(a, b):
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def test_config0_generate_character_through_the_installed_dropin(ref):
    from pybpl_b200 import _lib, dropin
    model, ctype, tokens, images = ref["model"], ref["ctype"], ref["tokens"], ref["images"]
    import pybpl.splines as S
    # ---- the un-patched reference ----
    want = []
    for tok, img in zip(tokens, images):
        pimg = model.get_pimg(tok).clone()
        ll, grads = _score_with_grads(model, tok, img)
        S.coefficient_mat.cache_clear()
        want.append(dict(pimg=pimg, ll=ll, grads=grads, lt=model.score_token(ctype, tok).detach(),
                         motor=[p.motor.clone() for p in tok.part_tokens]))
    want_type = model.score_type(ctype).detach()
    # ---- the same objects through the drop-in ----
    n0 = _lib.launch_count()
    dropin.install()
    try:
        import pybpl.model.image_dist as ID
        from pybpl_b200 import model as mirror
        assert ID.CharacterImageDist.score_image is mirror.CharacterImageDist.score_image
        for tok, img, w in zip(tokens, images, want):
            pimg = model.get_pimg(tok)
            assert pimg.device.type == "cpu" and pimg.dtype == torch.float32 and pimg.shape == (105, 105)
            assert float((pimg - w["pimg"]).abs().max()) <= 1e-5
            ll, grads = _score_with_grads(model, tok, img)
            assert ll.dim() == 0 and ll.device.type == "cpu"
            assert abs(float(ll) - float(w["ll"])) <= 1e-5 * abs(float(w["ll"]))
            bars = []
            for _ in tok.part_tokens:
                bars += [5e-4, 5e-4, 5e-4]
            bars += [1e-2]
            for g, gw, bar in zip(grads, w["grads"], bars):
                assert g.shape == gw.shape
                assert _relinf(g, gw) <= bar, (_relinf(g, gw), bar)
            lt = model.score_token(ctype, tok)
            assert abs(float(lt) - float(w["lt"])) <= 1e-5 * abs(float(w["lt"]))
            for p, mw in zip(tok.part_tokens, w["motor"]):          # StrokeToken.motor -> patched vanilla_to_motor
                assert torch.equal(p.motor, mw)
            s = model.sample_image(tok)
            assert s.shape == (105, 105) and set(np.unique(s.numpy())) <= {0.0, 1.0}
            assert abs(float(s.sum()) - float(w["pimg"].sum())) < 6 * float((w["pimg"] * (1 - w["pimg"])).sum()) ** 0.5 + 1
        st = model.score_type(ctype)
        assert abs(float(st) - float(want_type)) <= 2e-5 * abs(float(want_type))
        # a non-binary image is rejected like torch's Bernoulli does (image_dist.py:76-77)
        with pytest.raises(ValueError):
            model.score_image(tokens[0], images[0] * 0.5)
    finally:
        dropin.uninstall()
    assert _lib.launch_count() > n0, "the drop-in did not launch any kernel of libpybpl_b200.so"
    # uninstalled: the reference's own code again, bit-identical to before
    assert torch.equal(model.get_pimg(tokens[0]), want[0]["pimg"])


def test_render_image_and_spline_call_sites_on_reference_tokens(ref):
    """pybpl.rendering.render_image / pybpl.splines.get_stk_from_bspline called the way the reference's callers do
    (image_dist.py:34-40, part.py:320, relation.py:63), reference vs drop-in."""
    from pybpl_b200 import dropin
    import pybpl.rendering as R
    import pybpl.splines as S
    tokens = ref["tokens"]
    tok = tokens[1]
    motor = torch.cat([p.motor for p in tok.part_tokens])
    want_p, want_off = R.render_image(motor, tok.epsilon, tok.blur_sigma)
    want_list_p, _ = R.render_image([m for m in motor], 0.01, 4.0)
    Y = tok.part_tokens[0].shapes[:, :, 0] * tok.part_tokens[0].invscales[0]
    want_X = S.get_stk_from_bspline(Y, 200)
    want_Xa = S.get_stk_from_bspline(Y)
    S.coefficient_mat.cache_clear()
    want_Xs = S.get_stk_from_bspline(Y, s=torch.tensor(4.5))
    dropin.install()
    try:
        p, off = R.render_image(motor, tok.epsilon, tok.blur_sigma)
        assert float((p - want_p).abs().max()) <= 1e-5 and bool(off) == bool(want_off)
        p2, _ = R.render_image([m for m in motor], 0.01, 4.0)
        assert float((p2 - want_list_p).abs().max()) <= 1e-5
        assert torch.equal(S.get_stk_from_bspline(Y, 200), want_X)
        Xa = S.get_stk_from_bspline(Y)
        assert Xa.shape == want_Xa.shape and float((Xa - want_Xa).abs().max()) <= 1e-5 * float(want_Xa.abs().max())
        Xs = S.bspline_eval(torch.tensor(4.5), Y)
        assert float((Xs - want_Xs).abs().max()) <= 1e-5 * float(want_Xs.abs().max())
    finally:
        dropin.uninstall()


def test_changed_imsize_under_the_real_reference(ref):
    """Parameters.imsize other than 105 x 105 with the reference's own objects: render_image(strokes, eps, sigma, ps) with
    ps.imsize = (72, 90) (rendering.py:214), and a CharacterImageDist whose ps.imsize was changed (image_dist.py:30), the
    un-patched reference against the installed drop-in (which routes these sizes to bpl_render_strokes_any)."""
    from pybpl_b200 import dropin
    import pybpl.rendering as R
    from pybpl.parameters import Parameters
    model, tokens = ref["model"], ref["tokens"]
    tok = tokens[2]
    ps = Parameters(); ps.imsize = torch.Size([72, 90])
    motor = torch.cat([p.motor for p in tok.part_tokens]) * 0.8           # keep most of the ink on the smaller page
    want_p, want_off = R.render_image(motor, 1e-3, 1.5, ps)
    old_ps = model.image_dist.ps
    model.image_dist.ps = ps
    try:
        want_pimg = model.get_pimg(tok).clone()
        torch.manual_seed(3)
        img = torch.bernoulli(want_pimg)
        want_ll, want_g = _score_with_grads(model, tok, img)
        dropin.install()
        try:
            p, off = R.render_image(motor, 1e-3, 1.5, ps)
            assert p.shape == (72, 90) and float((p - want_p).abs().max()) <= 1e-5 and bool(off) == bool(want_off)
            pimg = model.get_pimg(tok)
            assert pimg.shape == (72, 90) and float((pimg - want_pimg).abs().max()) <= 1e-5
            ll, g = _score_with_grads(model, tok, img)
            assert abs(float(ll) - float(want_ll)) <= 1e-5 * abs(float(want_ll))
            bars = [5e-4] * (3 * len(tok.part_tokens)) + [1e-2]
            for a, b, bar in zip(g, want_g, bars):
                assert _relinf(a, b) <= bar, (_relinf(a, b), bar)
            with pytest.raises(ValueError):
                model.score_image(tok, img * 0.5)
        finally:
            dropin.uninstall()
    finally:
        model.image_dist.ps = old_ps
