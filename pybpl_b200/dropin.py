"""Rebinds the reference's hot-path call sites to the CUDA implementation.

    import pybpl, pybpl_b200.dropin
    pybpl_b200.dropin.install()          # pybpl.* now renders and scores on the GPU
    pybpl_b200.dropin.uninstall()

Patched names (each bound by-name in the reference, so both the defining module and its importers):
  pybpl.rendering.render_image, pybpl.model.image_dist.render_image          (rendering.py:185)
  pybpl.splines.get_stk_from_bspline (+ new alias pybpl.splines.bspline_eval),
  pybpl.objects.relation.get_stk_from_bspline                                 (splines.py:108, relation.py:9)
  pybpl.splines.fit_bspline_to_traj                                           (splines.py:141)
  pybpl.bottomup.initialize.util.{fit_bspline_to_traj, get_stk_from_bspline, fit_smooth_stk}
                                                                              (util.py:5,13; when that module imports)
  pybpl.objects.part.vanilla_to_motor                                         (part.py:290)
  CharacterImageDist.get_pimg / sample_image / score_image                    (image_dist.py:32-80)
  CharacterTokenDist.score_token                                              (token_dist.py:264-287)
  CharacterTypeDist.score_type                                                (type_dist.py:98-125; only when the
                                                                               instance's library is the shipped one,
                                                                               Library(use_hist=True))
"""
import functools
import importlib
import os
import weakref

from . import model, objects, rendering, splines

_saved = []


def _patch(mod, name, new):
    _saved.append((mod, name, getattr(mod, name, None)))
    setattr(mod, name, new)


def install(pybpl_module=None):
    if _saved:
        return
    if pybpl_module is None:
        pybpl_module = importlib.import_module("pybpl")
    R = importlib.import_module("pybpl.rendering")
    S = importlib.import_module("pybpl.splines")
    P = importlib.import_module("pybpl.objects.part")
    REL = importlib.import_module("pybpl.objects.relation")
    ID = importlib.import_module("pybpl.model.image_dist")
    _patch(R, "render_image", rendering.render_image)
    _patch(ID, "render_image", rendering.render_image)
    _patch(S, "get_stk_from_bspline", splines.get_stk_from_bspline)
    _patch(S, "bspline_eval", splines.bspline_eval)
    _patch(S, "fit_bspline_to_traj", splines.fit_bspline_to_traj)
    _patch(REL, "get_stk_from_bspline", splines.get_stk_from_bspline)
    try:      # the bottom-up parser binds both by name (bottomup/initialize/util.py:5); it needs networkx/skimage to import
        BU = importlib.import_module("pybpl.bottomup.initialize.util")
        _patch(BU, "fit_bspline_to_traj", splines.fit_bspline_to_traj)
        _patch(BU, "get_stk_from_bspline", splines.get_stk_from_bspline)
        _patch(BU, "fit_smooth_stk", splines.fit_smooth_stk)
    except Exception:
        pass
    _patch(P, "vanilla_to_motor", objects.vanilla_to_motor)
    mirror = model.CharacterImageDist
    for meth in ("get_pimg", "sample_image", "score_image"):
        _patch(ID.CharacterImageDist, meth, getattr(mirror, meth))
    TD = importlib.import_module("pybpl.model.token_dist")
    scorer = weakref.WeakKeyDictionary()      # per reference instance (an id() can be reused after garbage collection)

    def score_token(self, ctype, ctoken):      # the reference's instance keeps its own lib constants
        sc = scorer.get(self)
        if sc is None:
            sc = scorer[self] = model.CharacterTokenDist(_LibView(self))
        return sc.score_token(ctype, ctoken)
    _patch(TD.CharacterTokenDist, "score_token", score_token)
    YD = importlib.import_module("pybpl.model.type_dist")
    ref_score_type = YD.CharacterTypeDist.score_type
    type_scorer = model.CharacterTypeDist()

    shipped = weakref.WeakKeyDictionary()     # reference instance -> does it hold the shipped tables?

    def score_type(self, ctype):
        # the shipped tables are Library(use_hist=True)'s: any other (re-fitted, modified, use_hist=False) library keeps
        # the reference's own code.  Decided once per instance by comparing table values, not just shapes.
        ok = shipped.get(self)
        if ok is None:
            ok = shipped[self] = _holds_shipped_tables(self, type_scorer)
        if not ok:
            return ref_score_type(self, ctype)
        return type_scorer.score_type(ctype)
    _patch(YD.CharacterTypeDist, "score_type", score_type)


@functools.lru_cache(maxsize=1)
def _shipped_tables():
    import numpy as np
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "type_lib.npz")))


def _holds_shipped_tables(td, type_scorer):
    """True when a reference CharacterTypeDist instance carries the very tables pybpl_b200/data/type_lib.npz was exported
    from: histogram spatial model, 1212 primitives, and equal values in samples of the shape means, the Gamma
    parameters, the start / transition / stroke-count tables."""
    import numpy as np
    try:
        if not type(td.rdist.Spatial).__module__.endswith("spatial_OLD.spatial_model"):
            return False
        tab = _shipped_tables()
        mu = td.pdist.shapes_mu.detach().cpu().numpy().reshape(-1, 10)
        if mu.shape != tab["shape_mu"].reshape(-1, 10).shape:
            return False
        same = np.allclose(mu, tab["shape_mu"].reshape(-1, 10), rtol=1e-6, atol=0)
        same = same and np.allclose(td.kappa.probs.detach().cpu().numpy()[:tab["logp_kappa"].shape[0]], np.exp(tab["logp_kappa"]), rtol=1e-5, atol=1e-12)
        same = same and np.allclose(np.exp(td.pdist.logStart.detach().cpu().numpy()), np.exp(tab["logp_start"]), rtol=1e-5, atol=1e-6)
        same = same and np.allclose(td.pdist.scales_con.detach().cpu().numpy(), tab["gamma_con"], rtol=1e-6, atol=0)
        same = same and np.allclose(td.pdist.scales_rate.detach().cpu().numpy(), tab["gamma_rate"], rtol=1e-6, atol=0)
        return bool(same)
    except Exception:
        return False      # an unfamiliar instance: the reference's own code scores it


class _LibView:
    """The constants a reference CharacterTokenDist instance holds (token_dist.py:90-98, 320-325, 437-441), presented
    as the .tokenvar / .rel mappings ops.token_prior_params reads."""

    def __init__(self, td):
        self.tokenvar = {"sigma_shape": td.pdist.sigma_shape, "sigma_invscale": td.pdist.sigma_invscale,
                         "sigma_attach": td.rdist.sigma_attach}
        scale = td.loc_dist.base_dist.scale
        self.rel = {"sigma_x": scale[0], "sigma_y": scale[1]}


def uninstall():
    while _saved:
        mod, name, old = _saved.pop()
        if old is None:
            delattr(mod, name)
        else:
            setattr(mod, name, old)
