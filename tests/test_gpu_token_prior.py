"""GPU parity for log P(token | type): bpl_score_token_prior through the C ABI against the golden
vectors produced by the live reference's CharacterModel.score_token + autograd (tests/golden/token_prior.npz) and
against the numpy oracle (oracle/token_prior.py) on perturbed copies.

Bars: ll within 1e-5 relative (fp32); -inf exactly where the reference gives -inf; gradients within 2e-4 of the
tensor's inf-norm (the reference's fp32 autograd is itself ~1e-5 from the float64 oracle)."""
import numpy as np
import pytest
import torch

from tests.util_golden import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def make_inputs(ops, d, grad=False):
    K, nsub = d["K"], d["nsub"]
    tso = np.concatenate([[0], np.cumsum(K)]).astype(np.int32)
    sso = np.concatenate([[0], np.cumsum(nsub)]).astype(np.int32)
    leaves = {k: T(d[k]) for k in ("ctrl", "invscale", "position", "eval_spot", "ctrl_type", "invscale_type", "gpos",
                                   "eval_spot_type")}
    if grad:
        for v in leaves.values():
            v.requires_grad_(True)
    P = len(K)
    batch = ops.TokenBatch(tso, sso, leaves["ctrl"], leaves["invscale"], leaves["position"],
                           T(np.full(P, 1e-4, np.float32)), T(d["sigma"]))
    types = ops.TokenTypes(leaves["ctrl_type"], leaves["invscale_type"], d["rel_cat"], np.maximum(d["attach_ix"], 0),
                           np.maximum(d["attach_subix"], 0), leaves["gpos"], leaves["eval_spot_type"], leaves["eval_spot"])
    c = ops.token_prior_params()
    (c.sigma_shape, c.sigma_invscale, c.sigma_attach, c.sigma_x, c.sigma_y, c.min_blur_sigma,
     c.max_blur_sigma) = [float(v) for v in d["consts"]]
    return batch, types, c, leaves


def index_maps(d):
    tk = np.repeat(np.arange(len(d["K"])), d["K"])
    return tk, np.repeat(tk, d["nsub"])


def test_forward_matches_reference():
    from pybpl_b200 import ops
    d = dict(load("token_prior"))
    batch, types, c, _ = make_inputs(ops, d)
    ll = ops.score_token_prior(batch, types, c=c).cpu().numpy()
    ref = d["ll"]; fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(ll), fin)
    assert np.all(ll[~fin] == -np.inf)
    rel = np.abs(ll[fin] - ref[fin]) / np.abs(ref[fin])
    print("token prior ll max rel", rel.max())
    assert rel.max() < 1e-5


def test_gradients_match_reference_autograd():
    from pybpl_b200 import ops
    d = dict(load("token_prior"))
    batch, types, c, leaves = make_inputs(ops, d, grad=True)
    ll = ops.score_token_prior(batch, types, c=c)
    fin = np.isfinite(d["ll"])
    ll[T(fin, torch.bool)].sum().backward()           # the reference never back-propagates a -inf score
    tk, ts = index_maps(d)
    worst = {}
    for name, idx in (("ctrl", ts), ("ctrl_type", ts), ("invscale", ts), ("invscale_type", ts), ("position", tk),
                      ("gpos", tk), ("eval_spot", tk), ("eval_spot_type", tk)):
        m = fin[idx]
        a, b = leaves[name].grad.cpu().numpy()[m], d["g_" + name][m]
        worst[name] = np.abs(a - b).max() / max(1.0, np.abs(b).max())
    print("token prior grad errors (inf-norm relative)", worst)
    assert max(worst.values()) < 2e-4, worst


def test_matches_oracle_on_perturbed_batch():
    """Seeded perturbation of the golden batch (new values, same structure) against the fp64 oracle, forward and backward,
    with a non-uniform cotangent."""
    from oracle import token_prior as tp
    from pybpl_b200 import ops
    d = dict(load("token_prior"))
    rng = np.random.default_rng(3)
    d["ctrl"] = (d["ctrl"] + rng.normal(0, 2.0, d["ctrl"].shape)).astype(np.float32)
    d["invscale"] = np.abs(d["invscale"] + rng.normal(0, 0.02, d["invscale"].shape)).astype(np.float32) + 1e-3
    d["position"] = (d["position"] + rng.normal(0, 1.0, d["position"].shape)).astype(np.float32)
    d["eval_spot"] = np.where(d["rel_cat"] == 3, np.clip(d["eval_spot"] + rng.normal(0, 0.2, d["eval_spot"].shape), 2.05, 5.95),
                              0).astype(np.float32)
    d["sigma"] = rng.uniform(0.5, 15.9, d["sigma"].shape).astype(np.float32)
    C = load("tokens")["basis_5_200"]
    ref, g = tp.score(d, np.float64, grad=True, C0=C[0], C1=C[-1])
    assert np.all(np.isfinite(ref))
    batch, types, c, leaves = make_inputs(ops, d, grad=True)
    ll = ops.score_token_prior(batch, types, c=c)
    w = rng.uniform(0.5, 2.0, len(ref)).astype(np.float32)
    (ll * T(w)).sum().backward()
    rel = np.abs(ll.detach().cpu().numpy() - ref) / np.abs(ref)
    assert rel.max() < 1e-5, rel.max()
    tk, ts = index_maps(d)
    for name, idx in (("ctrl", ts), ("ctrl_type", ts), ("invscale", ts), ("invscale_type", ts), ("position", tk),
                      ("gpos", tk), ("eval_spot", tk), ("eval_spot_type", tk)):
        wv = w[idx].reshape((-1,) + (1,) * (g[name].ndim - 1))
        a, b = leaves[name].grad.cpu().numpy(), g[name] * wv
        assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(b).max()), name


def test_model_mirror_and_full_size_properties():
    """CharacterTokenDist.score_token on duck-typed objects; and at 64k tokens: a batch scored whole equals the same batch
    scored in two halves (warp-per-token independence), repeated runs agree bit for bit in the forward."""
    from pybpl_b200 import ops
    d = dict(load("token_prior"))
    reps = 410                                  # 160 * 410 = 65,600 tokens
    big = {k: (np.concatenate([v] * reps) if k not in ("consts",) else v) for k, v in d.items()}
    batch, types, c, _ = make_inputs(ops, big)
    ll1 = ops.score_token_prior(batch, types, c=c)
    ll2 = ops.score_token_prior(batch, types, c=c)
    assert torch.equal(ll1, ll2)
    small = ops.score_token_prior(*make_inputs(ops, d)[:2], c=c)
    assert torch.equal(ll1.view(reps, -1), small.view(1, -1).expand(reps, -1))
    # the reference-shaped entry point
    from types import SimpleNamespace as NS
    from pybpl_b200.model import CharacterTokenDist
    k, s = 0, 0
    inv_codes = {v: n for n, v in ops.REL_CODES.items()}
    for p in range(6):
        parts_ty, rels_ty, parts_tk, rels_tk = [], [], [], []
        for i in range(d["K"][p]):
            n = d["nsub"][k]
            sl = slice(s, s + n)
            parts_ty.append(NS(shapes=torch.from_numpy(d["ctrl_type"][sl]).permute(1, 2, 0), invscales=torch.from_numpy(d["invscale_type"][sl])))
            parts_tk.append(NS(shapes=torch.from_numpy(d["ctrl"][sl]).permute(1, 2, 0), invscales=torch.from_numpy(d["invscale"][sl]),
                               position=torch.from_numpy(d["position"][k])))
            cat = inv_codes[int(d["rel_cat"][k])]
            rels_ty.append(NS(category=cat, gpos=torch.from_numpy(d["gpos"][k]), attach_ix=int(d["attach_ix"][k]),
                              attach_subix=int(d["attach_subix"][k]), eval_spot=torch.tensor(d["eval_spot_type"][k])))
            rels_tk.append(NS(eval_spot_token=torch.tensor(d["eval_spot"][k])))
            k += 1; s += n
        ctype = NS(part_types=parts_ty, relation_types=rels_ty)
        ctoken = NS(part_tokens=parts_tk, relation_tokens=rels_tk, epsilon=1e-4, blur_sigma=float(d["sigma"][p]), affine=None)
        td = CharacterTokenDist()
        (td._c.sigma_shape, td._c.sigma_invscale, td._c.sigma_attach, td._c.sigma_x, td._c.sigma_y, td._c.min_blur_sigma,
         td._c.max_blur_sigma) = [float(v) for v in d["consts"]]
        got = td.score_token(ctype, ctoken)
        assert got.dim() == 0 and got.device.type == "cpu"
        ref = d["ll"][p]
        assert (np.isinf(ref) and float(got) == ref) or abs(float(got) - ref) <= 1e-5 * abs(ref)
