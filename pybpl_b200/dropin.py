"""Rebinds the reference's hot-path call sites to the CUDA implementation.

    import pybpl, pybpl_b200.dropin
    pybpl_b200.dropin.install()          # pybpl.* now renders and scores on the GPU
    pybpl_b200.dropin.uninstall()

Patched names (each bound by-name in the reference, so both the defining module and its importers):
  pybpl.rendering.render_image, pybpl.model.image_dist.render_image          (rendering.py:185)
  pybpl.splines.get_stk_from_bspline (+ new alias pybpl.splines.bspline_eval),
  pybpl.objects.relation.get_stk_from_bspline                                 (splines.py:108, relation.py:9)
  pybpl.splines.fit_bspline_to_traj                                           (splines.py:141)
  pybpl.objects.part.vanilla_to_motor                                         (part.py:290)
  CharacterImageDist.get_pimg / sample_image / score_image                    (image_dist.py:32-80)
  CharacterTokenDist.score_token                                              (token_dist.py:264-287)
  CharacterTypeDist.score_type                                                (type_dist.py:98-125; only when the
                                                                               instance's library is the shipped one,
                                                                               Library(use_hist=True))
"""
import importlib

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
    _patch(P, "vanilla_to_motor", objects.vanilla_to_motor)
    mirror = model.CharacterImageDist
    for meth in ("get_pimg", "sample_image", "score_image"):
        _patch(ID.CharacterImageDist, meth, getattr(mirror, meth))
    TD = importlib.import_module("pybpl.model.token_dist")
    scorer = {}

    def score_token(self, ctype, ctoken):      # the reference's instance keeps its own lib constants
        key = id(self)
        if key not in scorer:
            scorer[key] = model.CharacterTokenDist(_LibView(self))
        return scorer[key].score_token(ctype, ctoken)
    _patch(TD.CharacterTokenDist, "score_token", score_token)
    YD = importlib.import_module("pybpl.model.type_dist")
    ref_score_type = YD.CharacterTypeDist.score_type
    type_scorer = model.CharacterTypeDist()

    def score_type(self, ctype):
        # the shipped tables are Library(use_hist=True)'s: any other library keeps the reference's own code
        hist = type(self.rdist.Spatial).__module__.endswith("spatial_OLD.spatial_model")
        if not hist or self.pdist.shapes_mu.shape[0] != 1212:
            return ref_score_type(self, ctype)
        return type_scorer.score_type(ctype)
    _patch(YD.CharacterTypeDist, "score_type", score_type)


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
