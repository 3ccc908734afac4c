"""On-device type sampler (type half): bpl_sample_types against (i) the prior's own tables
(pybpl_b200/data/type_lib.npz, produced from the reference's Library) -- exact frequencies and probability-integral
transforms conditional on the drawn primitive ids -- and (ii) 4000 draws of the live reference's
CharacterModel.sample_type (tests/golden/type_samples.npz), two-sample tests per marginal.  The reference draws from
torch's global generator, so parity is distributional; counter-based streams make shards bit-equal."""
import os

import numpy as np
import pytest
import torch

from tests.util_golden import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 200_000


@pytest.fixture(scope="module")
def draw():
    from pybpl_b200 import ops
    lib = ops.TypeLibrary(DEV)
    tso, sso, types, subid = ops.sample_types(lib, N, seed=2025)
    tab = dict(np.load(os.path.join(ROOT, "pybpl_b200", "data", "type_lib.npz")))
    g = dict(tso=tso.cpu().numpy(), sso=sso.cpu().numpy(), subid=subid.cpu().numpy(), ctrl=types.ctrl_type.cpu().numpy(),
             inv=types.invscale_type.cpu().numpy(), cat=types.rel_cat.cpu().numpy(), aix=types.attach_ix.cpu().numpy(),
             asub=types.attach_subix.cpu().numpy(), gpos=types.gpos.cpu().numpy(), es=types.eval_spot_type.cpu().numpy())
    g["K"] = np.diff(g["tso"]); g["nsub"] = np.diff(g["sso"])
    g["rank"] = np.arange(len(g["nsub"])) - np.repeat(g["tso"][:-1], g["K"])
    g["k_of_stroke"] = np.repeat(g["K"], g["K"])
    return ops, lib, tab, g


def freq_ok(counts, p, nsig=5.0):
    n = counts.sum()
    sd = np.sqrt(n * p * (1 - p)) + 1.0
    return np.all(np.abs(counts - n * p) <= nsig * sd)


def test_structure_frequencies_match_the_tables(draw):
    ops, lib, tab, g = draw
    assert g["K"].min() >= 1 and g["K"].max() <= 10 and g["nsub"].min() >= 1 and g["nsub"].max() <= 10
    assert freq_ok(np.bincount(g["K"], minlength=11)[1:], tab["p_kappa"].astype(np.float64))
    for k in range(1, 6):
        sel = g["k_of_stroke"] == k
        assert freq_ok(np.bincount(g["nsub"][sel], minlength=11)[1:], tab["p_nsub"][k - 1].astype(np.float64)), k
    # first primitive of every stroke ~ exp(logStart); transitions out of the most common primitive ~ its row of pT
    first = g["subid"][g["sso"][:-1]]
    assert freq_ok(np.bincount(first, minlength=1212), tab["p_start"].astype(np.float64))
    inner = np.ones(len(g["subid"]), bool); inner[g["sso"][:-1]] = False          # sub-strokes that have a predecessor
    prev, nxt = g["subid"][:-1][inner[1:]], g["subid"][1:][inner[1:]]
    top = np.bincount(prev).argmax()
    row = np.diff(np.concatenate([[0.0], tab["cdf_T"][top].astype(np.float64)]))
    assert (prev == top).sum() > 500
    assert freq_ok(np.bincount(nxt[prev == top], minlength=1212), row)
    # relations: the first stroke is independent, later ones follow mixprob; attachments are uniform
    assert np.all(g["cat"][g["rank"] == 0] == 0)
    mix = np.diff(np.concatenate([[0.0], tab["rel_mixprob"].astype(np.float64)]))
    assert freq_ok(np.bincount(g["cat"][g["rank"] > 0], minlength=4), mix)
    att = g["cat"] > 0
    assert np.all(g["aix"][att] < g["rank"][att]) and np.all(g["aix"][att] >= 0)
    r3 = att & (g["rank"] == 3)
    assert freq_ok(np.bincount(g["aix"][r3], minlength=3), np.full(3, 1 / 3))
    mid = g["cat"] == 3
    tgt = np.repeat(g["tso"][:-1], g["K"])[mid] + g["aix"][mid]
    assert np.all(g["asub"][mid] < g["nsub"][tgt]) and np.all(g["asub"][mid] >= 0)
    assert np.all(g["asub"][~mid] == 0) and np.all(g["es"][~mid] == 0)
    assert np.all((g["es"][mid] >= 2.0) & (g["es"][mid] <= 6.0))


def test_continuous_parts_by_probability_integral_transform(draw):
    from scipy import stats
    ops, lib, tab, g = draw
    ids = g["subid"]
    sel = np.random.default_rng(0).choice(len(ids), 60000, replace=False)
    # shapes | id ~ N(mu[id], L L^T): z = L^-1 (x - mu) is standard normal in every coordinate
    x = g["ctrl"].reshape(-1, 10)[sel].astype(np.float64) - tab["shape_mu"][ids[sel]]
    Ls = tab["shape_L"][ids[sel]].astype(np.float64)
    z = np.stack([np.linalg.solve(Ls[i], x[i]) for i in range(0, len(sel), 6)])       # 10 000 solves
    for c in range(10):
        assert stats.kstest(z[:, c], "norm").pvalue > 1e-4, c
    assert abs(np.corrcoef(z.T)[np.triu_indices(10, 1)]).max() < 0.05
    # invscale | id ~ Gamma(con[id], rate[id])
    u = stats.gamma.cdf(g["inv"][sel].astype(np.float64) * tab["gamma_rate"][ids[sel]], a=tab["gamma_con"][ids[sel]].astype(np.float64))
    assert stats.kstest(u, "uniform").pvalue > 1e-4
    # evaluation spots ~ U(2, 6)
    mid = g["cat"] == 3
    assert stats.kstest((g["es"][mid] - 2.0) / 4.0, "uniform").pvalue > 1e-4
    # positions of independent strokes: bin frequencies of the stroke index's histogram, uniform inside the bin
    for h in range(3):
        s = (g["cat"] == 0) & ((g["rank"] == h) if h < 2 else (g["rank"] >= 2))
        gx, gy = g["gpos"][s, 0], g["gpos"][s, 1]
        xi = np.clip(np.searchsorted(tab["sp_xlab"], gx, side="right") - 1, 0, 24)
        yi = np.clip(np.searchsorted(tab["sp_ylab"], gy, side="right") - 1, 0, 24)
        p = np.diff(np.concatenate([[0.0], tab["sp_cdf"][h].astype(np.float64)]))
        assert freq_ok(np.bincount(xi * 25 + yi, minlength=625), p), h
        fx = (gx - tab["sp_xlab"][xi]) / 4.2
        assert stats.kstest(fx[:20000], "uniform").pvalue > 1e-4


def test_marginals_match_reference_draws(draw):
    """Two-sample tests against CharacterModel.sample_type itself (4000 draws)."""
    from scipy import stats
    ops, lib, tab, g = draw
    d = load("type_samples")
    def chi2(a, b, nb):
        ca, cb = np.bincount(a, minlength=nb)[:nb].astype(np.float64), np.bincount(b, minlength=nb)[:nb].astype(np.float64)
        keep = (ca + cb) > 20
        return stats.chi2_contingency(np.stack([ca[keep], cb[keep]])).pvalue
    ps = {}
    ps["K"] = chi2(g["K"], d["K"], 11)
    ps["nsub"] = chi2(g["nsub"], d["nsub"], 11)
    ps["rel_cat"] = chi2(g["cat"][g["rank"] > 0], d["rel_cat"][d["rank"] > 0], 4)
    ps["invscale"] = stats.ks_2samp(g["inv"][:100000], d["invscale"]).pvalue
    for c in (0, 3, 9):
        ps[f"ctrl{c}"] = stats.ks_2samp(g["ctrl"].reshape(-1, 10)[:100000, c], d["ctrl"][:, c]).pvalue
    r0g, r0d = (g["cat"] == 0) & (g["rank"] == 0), (d["rel_cat"] == 0) & (d["rank"] == 0)
    ps["gpos_x"] = stats.ks_2samp(g["gpos"][r0g, 0][:50000], d["gpos"][r0d, 0]).pvalue
    ps["gpos_y"] = stats.ks_2samp(g["gpos"][r0g, 1][:50000], d["gpos"][r0d, 1]).pvalue
    ps["eval_spot"] = stats.ks_2samp(g["es"][g["cat"] == 3][:50000], d["eval_spot"][d["rel_cat"] == 3]).pvalue
    # coarse primitive usage: the 30 most used primitives of the reference's draws
    topids = np.argsort(-np.bincount(d["subid"], minlength=1212))[:30]
    remap = np.full(1212, 30); remap[topids] = np.arange(30)
    ps["subid_top30"] = chi2(remap[g["subid"]], remap[d["subid"]], 31)
    print("two-sample p-values vs the reference's draws:", {k: round(float(v), 4) for k, v in ps.items()})
    assert min(ps.values()) > 1e-4, ps


def test_streams_shard_and_tokens_follow(draw):
    ops, lib, tab, g = draw
    n = 3000
    tso, sso, ty, sid = ops.sample_types(lib, n, seed=5)
    a = ops.sample_types(lib, n // 2, seed=5, first_type=0)
    b = ops.sample_types(lib, n - n // 2, seed=5, first_type=n // 2)
    assert torch.equal(ty.ctrl_type, torch.cat([a[2].ctrl_type, b[2].ctrl_type]))
    assert torch.equal(ty.gpos, torch.cat([a[2].gpos, b[2].gpos])) and torch.equal(sid, torch.cat([a[3], b[3]]))
    assert torch.equal(tso[: n // 2 + 1], a[0])
    other = ops.sample_types(lib, n, seed=6)
    assert not torch.equal(other[0], tso) or not torch.equal(other[2].gpos, ty.gpos)
    # prior -> tokens -> render + score, all on the device
    batch, ty2 = ops.sample_tokens(tso, sso, ty, seed=9)
    ll_prior = ops.score_token_prior(batch, ty2)
    assert torch.isfinite(ll_prior).all()
    img = ops.pack_images(torch.zeros(105, 105, dtype=torch.uint8, device=DEV))
    ll, off = ops.score_tokens(batch, img)
    assert torch.isfinite(ll).all()


# ---------------------------------------------------------------------------------------------
# the type score, log P(type)  (CharacterTypeDist.score_type, pybpl/model/type_dist.py:98-125)
# ---------------------------------------------------------------------------------------------
def _type_inputs(ops, d, grad=False):
    T = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=DEV)
    tso = np.concatenate([[0], np.cumsum(d["K"])]).astype(np.int32)
    sso = np.concatenate([[0], np.cumsum(d["nsub"])]).astype(np.int32)
    ctrl, inv = T(d["ctrl_type"]).requires_grad_(grad), T(d["invscale_type"]).requires_grad_(grad)
    K = len(d["nsub"])
    types = ops.TokenTypes(ctrl, inv, d["rel_cat"], d["attach_ix"], d["attach_subix"], T(d["gpos"]), T(d["eval_spot_type"]),
                           T(np.zeros(K, np.float32)))
    return tso, sso, types, T(d["subid"], torch.int32)


def test_type_score_matches_reference():
    from oracle import type_prior as tp
    from pybpl_b200 import ops
    d = dict(load("type_prior"))
    lib = ops.TypeLibrary(DEV)
    tso, sso, types, subid = _type_inputs(ops, d, grad=True)
    ll = ops.score_type_prior(lib, tso, sso, types, subid)
    ref = d["ll"]; fin = np.isfinite(ref)
    got = ll.detach().cpu().numpy()
    assert np.array_equal(np.isfinite(got), fin) and np.all(got[~fin] == -np.inf)
    rel = np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])
    print("type score max rel vs reference", rel.max())
    assert rel.max() < 1e-5
    ll[torch.as_tensor(fin, device=DEV)].sum().backward()
    ts = np.repeat(np.repeat(np.arange(len(d["K"])), d["K"]), d["nsub"]); m = fin[ts]
    gc, gi = types.ctrl_type.grad.cpu().numpy(), types.invscale_type.grad.cpu().numpy()
    assert np.abs(gc[m] - d["g_ctrl_type"][m]).max() <= 2e-5 * np.abs(d["g_ctrl_type"]).max()
    assert np.abs(gi[m] - d["g_invscale_type"][m]).max() <= 2e-5 * np.abs(d["g_invscale_type"]).max()
    # float64 oracle on the same inputs
    tab = dict(np.load(os.path.join(ROOT, "pybpl_b200", "data", "type_lib.npz")))
    o = tp.score(d, tab, np.float64)
    assert np.max(np.abs(got[fin] - o[fin]) / np.abs(o[fin])) < 2e-6


def test_type_score_outside_the_support_is_minus_inf():
    """A 'mid' relation whose spline coordinate lies outside Uniform(2, ncpt + 1), or whose attachment indices are out
    of range, has no finite score: the reference's Uniform / Categorical log_prob returns -inf or raises there
    (type_dist.py:636-647).  The kernel gives -inf (and never follows the bad index); the numpy oracle states the rule."""
    from oracle import type_prior as tp
    from pybpl_b200 import ops
    d = dict(load("type_prior"))
    mids = np.flatnonzero(d["rel_cat"] == ops.REL_CODES["mid"])
    later = np.flatnonzero((d["rel_cat"] != ops.REL_CODES["unihist"]))
    assert len(mids) >= 3 and len(later) >= 2
    d["eval_spot_type"] = d["eval_spot_type"].copy(); d["attach_ix"] = d["attach_ix"].copy(); d["attach_subix"] = d["attach_subix"].copy()
    d["eval_spot_type"][mids[0]] = 1.5            # below lb = 2
    d["eval_spot_type"][mids[1]] = 6.25           # above ub = ncpt + 1
    d["attach_subix"][mids[2]] = 99               # no such sub-stroke
    d["attach_ix"][later[-1]] = 31                # no such stroke
    stroke_type = np.repeat(np.arange(len(d["K"])), d["K"])
    hit = np.zeros(len(d["K"]), bool); hit[stroke_type[[mids[0], mids[1], mids[2], later[-1]]]] = True
    lib = ops.TypeLibrary(DEV)
    tso, sso, types, subid = _type_inputs(ops, d)
    got = ops.score_type_prior(lib, tso, sso, types, subid).cpu().numpy()
    tab = dict(np.load(os.path.join(ROOT, "pybpl_b200", "data", "type_lib.npz")))
    o = tp.score(d, tab, np.float64)
    assert np.all(got[hit] == -np.inf) and np.all(o[hit] == -np.inf)
    fin = np.isfinite(o)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max(np.abs(got[fin] - o[fin]) / np.abs(o[fin])) < 2e-6


def test_type_score_of_sampled_types_is_the_negative_entropy_rate(draw):
    """Consistency of sampler and scorer on 200 000 draws: every drawn type has a finite score, and the mean score of
    types with exactly one stroke and one sub-stroke matches the mean score of the reference's own such draws."""
    ops, lib, tab, g = draw
    tso, sso, types, subid = ops.sample_types(lib, 50_000, seed=77)
    ll = ops.score_type_prior(lib, tso, sso, types, subid).cpu().numpy()
    assert np.isfinite(ll).all()
    d = dict(load("type_prior"))
    from oracle import type_prior as tp
    sel = np.flatnonzero((d["K"] == 1) & np.isfinite(d["ll"]))
    K = np.diff(tso.cpu().numpy())
    mine = ll[K == 1]
    ref = d["ll"][sel]
    z = (mine.mean() - ref.mean()) / np.sqrt(mine.var() / len(mine) + ref.var() / len(ref))
    print(f"one-stroke types: mean log P GPU draws {mine.mean():.2f} (n={len(mine)}), reference draws {ref.mean():.2f} (n={len(ref)}), z={z:.2f}")
    assert abs(z) < 4.5


def test_model_mirror_score_type():
    """CharacterTypeDist.score_type on duck-typed type objects equals the golden reference scores."""
    from types import SimpleNamespace as NS
    from pybpl_b200 import ops
    from pybpl_b200.model import CharacterModel
    d = dict(load("type_prior"))
    inv_codes = {v: n for n, v in ops.REL_CODES.items()}
    model = CharacterModel()
    k = s = 0
    for p in range(12):
        parts, rels = [], []
        for i in range(d["K"][p]):
            n = d["nsub"][k]; sl = slice(s, s + n)
            parts.append(NS(shapes=torch.from_numpy(d["ctrl_type"][sl]).permute(1, 2, 0), invscales=torch.from_numpy(d["invscale_type"][sl]),
                            ids=torch.from_numpy(d["subid"][sl])))
            rels.append(NS(category=inv_codes[int(d["rel_cat"][k])], gpos=torch.from_numpy(d["gpos"][k]), attach_ix=int(d["attach_ix"][k]),
                           attach_subix=int(d["attach_subix"][k]), eval_spot=torch.tensor(d["eval_spot_type"][k])))
            k += 1; s += n
        got = model.score_type(NS(part_types=parts, relation_types=rels))
        ref = d["ll"][p]
        assert got.dim() == 0 and got.device.type == "cpu"
        assert (np.isinf(ref) and float(got) == ref) or abs(float(got) - ref) <= 1e-5 * abs(ref)


def test_optimize_types_example_raises_the_prior_score():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import optimize_types
    r = optimize_types.optimise(n_types=256, steps=40, lr=1e-2)
    assert np.isfinite(r["scores"]).all() and r["scores"][-1] > r["scores"][0]
