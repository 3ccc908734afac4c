"""Host-side mirror of the reference's image distribution (pybpl/model/image_dist.py:24-80) and the
three CharacterModel methods on the hot path (pybpl/model/model.py:32-39).

The tokens are duck-typed: anything with `.part_tokens[i].{shapes (ncpt,2,nsub), invscales (nsub,),
position (2,)}`, `.affine` (None or (4,)), `.epsilon`, `.blur_sigma` works -- the reference's
CharacterToken/StrokeToken (pybpl/objects/concept.py:217-241, part.py:188-249) or the light containers
in pybpl_b200.objects."""
import torch

from . import ops
from .parameters import Parameters


def _compute_device(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def token_batch_from(ctokens, device=None, neval=200):
    """Pack a list of character tokens into one ops.TokenBatch (keeps the autograd graph to the leaves)."""
    ctrl, inv, pos, tso, sso, eps, sig, aff = [], [], [], [0], [0], [], [], []
    any_aff = any(getattr(t, "affine", None) is not None for t in ctokens)
    if device is None:
        device = _compute_device(ctokens[0].part_tokens[0].shapes)
    for t in ctokens:
        for p in t.part_tokens:
            sh = p.shapes
            assert sh.dim() == 3 and sh.shape[1] == 2
            ctrl.append(sh.permute(2, 0, 1))
            inv.append(p.invscales.reshape(-1))
            pos.append(p.position.reshape(1, 2))
            sso.append(sso[-1] + sh.shape[2])
        tso.append(tso[-1] + len(t.part_tokens))
        eps.append(torch.as_tensor(t.epsilon, dtype=torch.float32).reshape(1))
        sig.append(torch.as_tensor(t.blur_sigma, dtype=torch.float32).reshape(1))
        if any_aff:
            a = getattr(t, "affine", None)
            aff.append((torch.tensor([1., 1., 0., 0.]) if a is None else a).reshape(1, 4))
    to = lambda xs: torch.cat([x.to(device, torch.float32) for x in xs])
    return ops.TokenBatch(tso, sso, to(ctrl).contiguous(), to(inv), to(pos), to(eps), to(sig),
                          affine=to(aff) if any_aff else None, neval=neval)


class CharacterImageDist:
    """P(Image | Token) (pybpl/model/image_dist.py:24-80)."""

    def __init__(self, lib=None):
        self.lib = lib
        self.ps = Parameters()

    def get_pimg(self, ctoken):
        src = ctoken.part_tokens[0].shapes.device
        batch = token_batch_from([ctoken])
        pimg, _ = ops.render_tokens(batch, ps=self.ps)
        return pimg[0].to(src)

    def sample_image(self, ctoken):
        """Binary image sample (image_dist.py:44-58); torch's global RNG, no gradient."""
        with torch.no_grad():
            pimg = self.get_pimg(ctoken)
            return torch.bernoulli(pimg)

    def score_image(self, ctoken, image):
        """Log-likelihood of a binary image (image_dist.py:60-80): 0-d float32 tensor."""
        src = ctoken.part_tokens[0].shapes.device
        batch = token_batch_from([ctoken])
        ps = ops.make_params(self.ps)
        if ops.image_size(ps) != (ops.IMSIZE, ops.IMSIZE):       # self.ps.imsize was changed: run-time-geometry kernels
            imgs = ops.binary_images(image.detach(), ps, batch.device)
        else:
            imgs = ops.pack_images(image.detach().to(batch.device), validate=True)
        ll, _ = ops.score_tokens(batch, imgs, ps=ps)
        return ll[0].to(src)


def token_types_from(ctypes_, ctokens, device):
    """Pack the type-level parameters and the relation tokens of (type, token) pairs into one ops.TokenTypes, aligned
    with token_batch_from(ctokens).  Duck-typed on the reference's objects: ctype.part_types[i].{shapes, invscales}
    (pybpl/objects/part.py:100-160), ctype.relation_types[i].{category, gpos | attach_ix [, attach_subix, eval_spot]}
    (objects/relation.py:238-420), ctoken.relation_tokens[i].eval_spot_token (relation.py:14-32)."""
    ctrl, inv, cat, aix, asub, gpos, es_ty, es = [], [], [], [], [], [], [], []
    z1 = torch.zeros(1)
    for ct, tk in zip(ctypes_, ctokens):
        for pt, rt, rtk in zip(ct.part_types, ct.relation_types, tk.relation_tokens):
            ctrl.append(pt.shapes.permute(2, 0, 1))
            inv.append(pt.invscales.reshape(-1))
            c = ops.REL_CODES[rt.category]
            cat.append(c)
            aix.append(int(rt.attach_ix) if c != 0 else 0)
            asub.append(int(rt.attach_subix) if c == 3 else 0)
            gpos.append(rt.gpos.reshape(1, 2) if c == 0 else torch.zeros(1, 2))
            es_ty.append(torch.as_tensor(rt.eval_spot, dtype=torch.float32).reshape(1) if c == 3 else z1)
            es.append(torch.as_tensor(rtk.eval_spot_token, dtype=torch.float32).reshape(1) if c == 3 else z1)
    to = lambda xs: torch.cat([x.to(device, torch.float32) for x in xs])
    return ops.TokenTypes(to(ctrl).contiguous(), to(inv), cat, aix, asub, to(gpos), to(es_ty), to(es))


class CharacterTokenDist:
    """The scoring half of P(Token | Type) (pybpl/model/token_dist.py:85-287): score_token on the GPU.
    Sampling (sample_token, :226-262) stays with the reference's prior objects."""

    def __init__(self, lib=None):
        self.lib = lib
        self.ps = Parameters()
        self._c = ops.token_prior_params(lib, self.ps)

    def score_tokens(self, ctypes_, ctokens):
        """[P] scores for paired lists of types and tokens (one launch)."""
        batch = token_batch_from(ctokens)
        types = token_types_from(ctypes_, ctokens, batch.device)
        return ops.score_token_prior(batch, types, c=self._c)

    def score_token(self, ctype, ctoken):
        """log P(Token = ctoken | Type = ctype) (token_dist.py:264-287): 0-d float32 tensor on the token's device."""
        src = ctoken.part_tokens[0].shapes.device
        return self.score_tokens([ctype], [ctoken])[0].to(src)


class CharacterTypeDist:
    """The scoring half of P(Type) (pybpl/model/type_dist.py:98-125, 127-207): score_type on the GPU, differentiable in the
    stroke types' shapes and scales (what examples/optimize_type.py optimises).  `library`: an ops.TypeLibrary."""

    def __init__(self, library=None, device=None):
        self.library = library
        self._device = device

    def _lib(self, dev):
        if self.library is None:
            self.library = ops.TypeLibrary(dev)
        return self.library

    def score_types(self, ctypes_):
        """[P] scores for a list of types (one launch).  Duck-typed on ctype.part_types[i].{shapes, invscales, ids} and
        ctype.relation_types[i] (pybpl/objects/part.py:100-160, relation.py:238-420)."""
        dev = self._device or _compute_device(ctypes_[0].part_types[0].shapes)
        lib = self._lib(dev)
        ctrl, inv, ids, tso, sso, cat, aix, asub, gpos, es = [], [], [], [0], [0], [], [], [], [], []
        z1 = torch.zeros(1)
        for ct in ctypes_:
            for pt, rt in zip(ct.part_types, ct.relation_types):
                ctrl.append(pt.shapes.permute(2, 0, 1)); inv.append(pt.invscales.reshape(-1))
                ids.append(torch.as_tensor(pt.ids).reshape(-1).to(torch.int32))
                sso.append(sso[-1] + pt.shapes.shape[2])
                c = ops.REL_CODES[rt.category]
                cat.append(c); aix.append(int(rt.attach_ix) if c else 0); asub.append(int(rt.attach_subix) if c == 3 else 0)
                gpos.append(rt.gpos.detach().reshape(1, 2) if c == 0 else torch.zeros(1, 2))
                es.append(torch.as_tensor(rt.eval_spot, dtype=torch.float32).detach().reshape(1) if c == 3 else z1)
            tso.append(tso[-1] + len(ct.part_types))
        to = lambda xs: torch.cat([x.to(dev, torch.float32) for x in xs])
        types = ops.TokenTypes(to(ctrl).contiguous(), to(inv), cat, aix, asub, to(gpos), to(es),
                               torch.zeros(len(cat), dtype=torch.float32, device=dev))
        return ops.score_type_prior(lib, tso, sso, types, torch.cat(ids).to(dev))

    def score_type(self, ctype):
        """log P(Type = ctype) (type_dist.py:98-125): 0-d float32 tensor on the type's device."""
        src = ctype.part_types[0].shapes.device
        return self.score_types([ctype])[0].to(src)


class CharacterModel:
    """The hot-path part of pybpl/model/model.py:8-39: image methods and score_token run on the GPU.  The type prior and
    token sampling are out of scope: pass `type_dist`/`token_dist` objects (e.g. the reference's) for them."""

    def __init__(self, lib=None, type_dist=None, token_dist=None):
        self.type_dist = type_dist
        self.token_dist = token_dist          # sample_token (and score_token unless overridden below)
        self.token_score = CharacterTokenDist(lib)
        self.type_score = CharacterTypeDist()          # shipped prior tables (pybpl_b200/data/type_lib.npz)
        self.image_dist = CharacterImageDist(lib)

    def score_token(self, ctype, ctoken):
        return self.token_score.score_token(ctype, ctoken)

    def score_type(self, ctype):
        return self.type_score.score_type(ctype)

    def sample_image(self, ctoken):
        return self.image_dist.sample_image(ctoken)

    def score_image(self, ctoken, image):
        return self.image_dist.score_image(ctoken, image)

    def get_pimg(self, ctoken):
        return self.image_dist.get_pimg(ctoken)

    def __getattr__(self, name):
        # sample_type/score_type/sample_token/score_token delegate to the supplied prior objects
        for prefix, d in (("sample_type", "type_dist"), ("score_type", "type_dist"),
                          ("sample_token", "token_dist"), ("score_token", "token_dist")):
            if name == prefix:
                dist = self.__dict__.get(d)
                if dist is None:
                    raise AttributeError(f"{name}: no {d} supplied (priors are outside pybpl_b200's scope)")
                return getattr(dist, name)
        raise AttributeError(name)
