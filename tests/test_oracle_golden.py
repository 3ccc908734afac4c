"""Pins the CPU oracle (oracle/bpl_oracle.c) against the golden vectors that
tests/golden/make_golden.py produced from the live reference (rfeinman/pyBPL).

The reference's own tests hold no values for this path
(pybpl/tests/test_vanilla_to_motor.py:47-48 asserts shapes only), so these
fixtures are the pin.  Tolerances: integer work bit-exact; trajectories
bit-exact given the reference's basis table; fp32 images/ll/grads to the
reference's own fp32 noise.
"""
import os

import numpy as np
import pytest

import oracle as O
from oracle import Oracle
from tests.util_golden import load, render_cases, token_cases, relinf


@pytest.fixture(scope="module")
def o32():
    return Oracle(np.float32)


def test_basis_tables_close_to_reference(o32):
    # splines.py:72-93.  torch builds the table with Sleef powf + an MKL dot, which no
    # portable C sequence reproduces to the bit; the restatement is within 1 ulp-ish.
    d = load("splines")
    for nland, neval in d["table_pairs"]:
        C = o32.basis_table(int(nland), int(neval))
        ref = d[f"table_{nland}_{neval}"]
        assert C.shape == ref.shape
        assert np.abs(C - ref).max() < 1e-6
        assert np.array_equal(C != 0, ref != 0)
    C = o32.basis_rows(5, d["s_explicit"])
    assert np.abs(C - d["table_s_explicit"]).max() < 1e-6


def test_spline_eval_bit_exact_given_table(o32):
    # splines.py:136: torch.matmul(C, Y) == fma chain
    d = load("splines")
    assert np.array_equal(o32.spline_eval(d["table_5_200"], d["Y_explicit"]), d["X_200"])
    assert np.array_equal(o32.spline_eval(d["table_7_200"], d["Y7"]), d["X7_200"])
    assert np.array_equal(o32.spline_eval(d["table_s_explicit"], d["Y_explicit"]), d["X_explicit"])


def test_adaptive_neval_integer_exact(o32):
    # splines.py:129-133 (integer work)
    d = load("splines")
    nev, dist = o32.adaptive_neval(d["table_5_200"], d["Y_adaptive"])
    assert np.array_equal(nev, d["neval_adaptive"])
    assert relinf(dist, d["dist_adaptive"]) < 2e-6


@pytest.mark.parametrize("case", render_cases(), ids=lambda c: c["name"])
def test_render_cases(case):
    # rendering.py:185-229 + image_dist.py:75-78 on raw trajectories
    o = Oracle(np.float32, O.default_params(hinton=case["hinton"]))
    m, fl, ce = o.point_indices(case["pts"])
    assert np.array_equal(m, case["mask_out"])
    inb = ~m
    assert np.array_equal(fl[inb], case["floor"][inb]) and np.array_equal(ce[inb], case["ceil"][inb])
    res = o.render_grad(case["strokes"], case["eps"], case["sigma"], image=case["image"])
    assert res["off_page"] == case["off"]
    assert np.abs(res["pimg"] - case["pimg"]).max() <= 3e-6 * max(1.0, np.abs(case["pimg"]).max())
    assert abs(res["ll"] - case["ll"]) <= 2e-6 * abs(case["ll"])
    assert relinf(res["g_pts"], case["g_pts"]) < 1e-4
    assert abs(res["g_eps"] - case["g_eps"]) <= 2e-4 * max(1.0, abs(case["g_eps"]))
    assert abs(res["g_sigma"] - case["g_sigma"]) <= 2e-3 * max(1.0, abs(case["g_sigma"]))
    # generic cotangent (render_image drop-in backward)
    ct = case["ct"]
    r2 = o.render_grad(case["strokes"], case["eps"], case["sigma"], gP=ct)
    assert relinf(r2["g_pts"], case["gct_pts"]) < 1e-4
    assert abs(r2["g_eps"] - case["gct_eps"]) <= 1e-4 * max(1.0, abs(case["gct_eps"]))
    assert abs(r2["g_sigma"] - case["gct_sigma"]) <= 2e-3 * max(1.0, abs(case["gct_sigma"]))


def test_tokens_whole_path(o32):
    # CharacterImageDist.get_pimg/score_image (image_dist.py:32-42,75-78) incl. vanilla_to_motor + affine
    d, toks = token_cases()
    C = d["basis_5_200"]
    worst = dict(ll=0, gc=0, gi=0, gp=0, gs=0, ge=0, ga=0)
    for t in toks:
        res = o32.token(C, t["nsub"], t["ctrl"], t["invscale"], t["position"], t["eps"], t["sigma"],
                        affine=t["affine"] if t["has_affine"] else None, image=t["image"], grad=True)
        assert res["off_page"] == t["off_page"], t["name"]
        if "motor" in t and not t["has_affine"]:
            assert np.array_equal(res["motor"], t["motor"]), t["name"]     # bit-exact trajectories
            assert np.abs(res["pimg"] - t["pimg"]).max() < 3e-6
        if not t["has_affine"]:
            # integer work: in-bounds counts and floor sums per sub-stroke (bit-exact)
            m, fl, _ = o32.point_indices(res["motor"].reshape(-1, 2))
            inb = (~m).reshape(-1, 200)
            assert np.array_equal(inb.sum(1).astype(np.int32), t["n_inbounds"])
            fsum = (fl.reshape(-1, 200, 2).astype(np.int64) * inb[..., None]).sum(1)
            assert np.array_equal(fsum, t["floor_sum"])
        worst["ll"] = max(worst["ll"], abs(res["ll"] - t["ll"]) / abs(t["ll"]))
        worst["gc"] = max(worst["gc"], relinf(res["g_ctrl"], t["g_ctrl"]))
        worst["gi"] = max(worst["gi"], relinf(res["g_invscale"], t["g_invscale"]))
        worst["gp"] = max(worst["gp"], relinf(res["g_position"], t["g_position"]))
        worst["gs"] = max(worst["gs"], abs(res["g_sigma"] - t["g_sigma"]) / max(1.0, abs(t["g_sigma"])))
        worst["ge"] = max(worst["ge"], abs(res["g_eps"] - t["g_eps"]) / max(1.0, abs(t["g_eps"])))
        if t["has_affine"]:
            worst["ga"] = max(worst["ga"], relinf(res["g_affine"], t["g_affine"]))
    print("worst", worst)
    assert worst["ll"] < 1e-5
    assert worst["gc"] < 5e-4 and worst["gi"] < 5e-4 and worst["gp"] < 5e-4  # reference fp32 autograd noise (App. B.4)
    assert worst["gs"] < 1e-2 and worst["ge"] < 2e-3 and worst["ga"] < 2e-3


def test_batch_equals_single(o32):
    d, toks = token_cases()
    from tests.util_golden import pack_tokens
    sel = [t for t in toks if not t["has_affine"]][:12]
    b = pack_tokens(sel)
    res = o32.token_batch(d["basis_5_200"], b["tok_stroke_off"], b["stroke_sub_off"], b["ctrl"], b["invscale"],
                          b["position"], b["eps"], b["sigma"], b["images"], img_idx=b["img_idx"])
    for i, t in enumerate(sel):
        assert abs(res["ll"][i] - t["ll"]) <= 1e-5 * abs(t["ll"])
        assert bool(res["off_page"][i]) == t["off_page"]


def test_fp64_noise_floor():
    # the fp64 instantiation only measures how far the fp32 reference is from exact arithmetic
    d, toks = token_cases()
    o64 = Oracle(np.float64)
    C64 = d["basis_5_200"].astype(np.float64)
    errs = []
    for t in toks[:16]:
        if t["has_affine"]:
            continue
        r = o64.token(C64, t["nsub"], t["ctrl"], t["invscale"], t["position"], t["eps"], t["sigma"], image=t["image"])
        errs.append(abs(r["ll"] - t["ll"]) / abs(t["ll"]))
    assert max(errs) < 2e-4  # reference fp32 ll is itself only ~5e-5 from fp64


def test_fit_bspline_to_traj():
    # splines.py:141-171 (torch.linalg.lstsq gels): control points and residuals
    d = load("fit")
    for dt, tol in ((np.float32, 5e-6), (np.float64, 2e-6)):
        o = Oracle(dt)
        for ci, (nl, ne) in enumerate(d["cases"]):
            Y, res = o.lstsq(d[f"C_{ci}"], d[f"X_{ci}"])
            assert relinf(Y, d[f"Y_{ci}"]) < tol
            assert relinf(res, d[f"res_{ci}"]) < 1e-4


# ---------------------------------------------------------------------------------------------
# token prior, log P(token | type): oracle/token_prior.py against the live reference's outputs
# ---------------------------------------------------------------------------------------------
def _token_prior_golden():
    d = dict(load("token_prior"))
    C = load("tokens")["basis_5_200"]
    K, nsub = d["K"], d["nsub"]
    tok_of_stroke = np.repeat(np.arange(len(K)), K)
    return d, C, tok_of_stroke, np.repeat(tok_of_stroke, nsub)


def test_token_prior_oracle_matches_reference():
    """score_token and its autograd gradients w.r.t. all token- and type-level leaves (160 prior-sampled pairs)."""
    from oracle import token_prior as tp
    d, C, tk, ts = _token_prior_golden()
    ref = d["ll"]; fin = np.isfinite(ref)
    assert (~fin).sum() >= 8 and fin.sum() >= 100            # the fixture holds -inf cases (invscale <= 0, spot out of range)
    assert set(np.unique(d["rel_cat"])) == {0, 1, 2, 3}      # ... and every relation category
    for dt, tol in ((np.float64, 1e-6), (np.float32, 1e-5)):
        ll = tp.score(d, dt, C0=C[0], C1=C[-1])
        assert np.array_equal(np.isfinite(ll), fin)
        assert np.all(ll[~fin] == -np.inf)
        assert np.max(np.abs(ll[fin] - ref[fin]) / np.abs(ref[fin])) < tol
    _, g = tp.score(d, np.float64, grad=True, C0=C[0], C1=C[-1])
    for name, idx in (("ctrl", ts), ("ctrl_type", ts), ("invscale", ts), ("invscale_type", ts), ("position", tk),
                      ("gpos", tk), ("eval_spot", tk), ("eval_spot_type", tk)):
        m = fin[idx]
        a, b = g[name][m], d["g_" + name][m]
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), name


def test_token_prior_oracle_gradient_is_the_derivative():
    """Central differences of the float64 forward, for the leaves the attach point couples (the reference's own
    autograd drops d/d eval_spot_token through a stale lru_cache unless the cache is cleared)."""
    from oracle import token_prior as tp
    d, C, tk, ts = _token_prior_golden()
    C = C.astype(np.float64)
    fin = np.isfinite(d["ll"])
    _, g = tp.score(d, np.float64, grad=True, C0=C[0], C1=C[-1])
    rng = np.random.default_rng(0)
    mids = [k for k in range(len(tk)) if d["rel_cat"][k] == 3 and fin[tk[k]]]
    checks = [("eval_spot", k, tk[k]) for k in mids[:10]]
    for k in mids[:6]:                  # the stroke attached to: its position and first sub-stroke feed the attach point
        j = int(np.flatnonzero(tk == tk[k])[0] + d["attach_ix"][k])
        s = int(np.concatenate([[0], np.cumsum(d["nsub"])])[j])
        checks += [("position", (j, 0), tk[k]), ("invscale", s, tk[k]), ("ctrl", (s, int(rng.integers(5)), 1), tk[k])]
    h = 1e-6
    for name, idx, p in checks:
        hi = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in d.items()}
        lo = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in hi.items()}
        hi = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in hi.items()}
        hi[name][idx] += h; lo[name][idx] -= h
        fd = (tp.score(hi, np.float64, C0=C[0], C1=C[-1])[p] - tp.score(lo, np.float64, C0=C[0], C1=C[-1])[p]) / (2 * h)
        assert abs(fd - g[name][idx]) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, g[name][idx])


def test_type_prior_oracle_matches_reference():
    """oracle/type_prior.py against the live reference's score_type + autograd (300 prior-sampled types, 13 with a
    position outside the histogram -> -inf)."""
    from oracle import type_prior as tp
    d = dict(load("type_prior"))
    tab = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pybpl_b200", "data",
                                    "type_lib.npz")))
    ref = d["ll"]; fin = np.isfinite(ref)
    assert (~fin).sum() >= 5 and set(np.unique(d["rel_cat"])) == {0, 1, 2, 3}
    ll, g = tp.score(d, tab, np.float64, grad=True)
    assert np.array_equal(np.isfinite(ll), fin) and np.all(ll[~fin] == -np.inf)
    assert np.max(np.abs(ll[fin] - ref[fin]) / np.abs(ref[fin])) < 1e-5          # the reference is float32
    ts = np.repeat(np.repeat(np.arange(len(d["K"])), d["K"]), d["nsub"])
    m = fin[ts]
    assert np.abs(g["ctrl_type"][m] - d["g_ctrl_type"][m]).max() <= 2e-5 * np.abs(d["g_ctrl_type"]).max()
    assert np.abs(g["invscale_type"][m] - d["g_invscale_type"][m]).max() <= 2e-5 * np.abs(d["g_invscale_type"]).max()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_reference_on_the_config3_start_state():
    """The C oracle (fp32, forward + backward) on the 64 parses of configs[2] against the live reference's autograd
    (tests/golden/fit64.npz)."""
    from oracle import Oracle
    from pybpl_b200 import workload
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "fit64.npz")))
    w = workload.synth_tokens(int(d["P"]), seed=int(d["seed"]), sigma_mode=10.0)
    C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
    ref = Oracle(np.float32).token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"],
                                         w["eps"], w["sigma"], d["target"][None], grad=True)
    assert (np.abs(ref["ll"] - d["ll"]) / np.abs(d["ll"])).max() < 1e-5
    rel = lambda g, r: np.abs(g - r).max() / np.abs(r).max()
    assert rel(ref["g_ctrl"], d["g_ctrl"]) < 5e-4 and rel(ref["g_invscale"], d["g_invscale"]) < 5e-4
    assert rel(ref["g_position"], d["g_position"]) < 5e-4
    assert rel(ref["g_sigma"], d["g_sigma"]) < 1e-2 and rel(ref["g_eps"], d["g_eps"]) < 2e-3


def test_oracle_smaller_gaussian_matches_reference():
    """Parameters.fsize = 7, 5, 3 (pybpl/parameters.py:41): the oracle's probability maps against the live reference's
    (tests/golden/fsize.npz, made by tests/golden/make_fit_golden.py --fsize)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "fsize.npz"))
    for fsize in (7, 5, 3):
        ps = O.default_params(); ps.fsize = fsize
        o = Oracle(np.float32, ps)
        for i in range(int(d["n_sets"])):
            key = f"f{fsize}_s{i}"
            strokes = [d[f"{key}_stroke{j}"] for j in range(int(d[key + "_n"]))]
            pimg, off = o.render(strokes, float(d[key + "_eps"]), float(d[key + "_sigma"]))
            assert off == bool(d[key + "_off"])
            assert np.abs(pimg - d[key + "_pimg"]).max() < 1e-5, (fsize, i)


def test_oracle_other_image_sizes_match_reference():
    """Parameters.imsize other than 105 x 105 (pybpl/parameters.py:21; H x W = imsize[0] x imsize[1]): the oracle's
    probability maps, off-page flags, log-likelihoods and gradients against the live reference's (tests/golden/imsize.npz,
    made by tests/golden/make_fit_golden.py --imsize: 64 x 80, 150 x 130 and 33 x 47 with the Hinton broaden)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "imsize.npz"))
    rel = lambda g, r: np.abs(g - r).max() / max(np.abs(r).max(), 1e-12)
    for key in d["cases"]:
        ps = O.default_params(hinton=bool(d[key + "_hinton"])); ps.H = int(d[key + "_H"]); ps.W = int(d[key + "_W"])
        o = Oracle(np.float32, ps)
        strokes = [d[f"{key}_stroke{j}"] for j in range(int(d[key + "_n"]))]
        eps, sigma = float(d[key + "_eps"]), float(d[key + "_sigma"])
        pimg, off = o.render(strokes, eps, sigma)
        assert pimg.shape == (ps.H, ps.W) and off == bool(d[key + "_off"])
        assert np.abs(pimg - d[key + "_pimg"]).max() < 1e-5, key
        r = o.render_grad(strokes, eps, sigma, gP=d[key + "_ct"])
        g_ref = np.concatenate([d[f"{key}_g_stroke{j}"] for j in range(len(strokes))])
        assert rel(r["g_pts"], g_ref) < 5e-4, key
        assert abs(r["g_sigma"] - float(d[key + "_g_sigma"])) <= 1e-2 * max(1.0, abs(float(d[key + "_g_sigma"])))
        assert abs(r["g_eps"] - float(d[key + "_g_eps"])) <= 2e-3 * max(1.0, abs(float(d[key + "_g_eps"])))
        r = o.render_grad(strokes, eps, sigma, image=d[key + "_image"])
        assert abs(r["ll"] - float(d[key + "_ll"])) <= 1e-5 * abs(float(d[key + "_ll"])), key
        g_ref = np.concatenate([d[f"{key}_gll_stroke{j}"] for j in range(len(strokes))])
        assert rel(r["g_pts"], g_ref) < 5e-4, key
        assert abs(r["g_sigma"] - float(d[key + "_gll_sigma"])) <= 1e-2 * max(1.0, abs(float(d[key + "_gll_sigma"])))
        assert abs(r["g_eps"] - float(d[key + "_gll_eps"])) <= 2e-3 * max(1.0, abs(float(d[key + "_gll_eps"])))
