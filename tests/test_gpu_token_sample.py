"""On-device token sampler (token half): bpl_sample_tokens against draws of the live reference's
CharacterModel.sample_token (tests/golden/token_samples.npz: 4 types x 600 draws).  The reference samples from torch's
global Mersenne generator, so parity is distributional: two-sample Kolmogorov-Smirnov tests per marginal, moments of the
standardised residuals, the mean log-probability of the draws under the (parity-tested) scorer, plus the properties a
counter-based generator gives: bit-equal shards, no dependence on the launch shape."""
import numpy as np
import pytest
import torch

from tests.util_golden import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
R = 20000            # GPU draws per type


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def expand(ops, d, ti, reps):
    g = lambda k: d[f"t{ti}_{k}"]
    nsub = g("nsub"); K = len(nsub)
    tso = np.arange(reps + 1, dtype=np.int32) * K
    sso = np.concatenate([[0], np.cumsum(np.tile(nsub, reps))]).astype(np.int32)
    til = lambda a: np.concatenate([a] * reps)
    types = ops.TokenTypes(T(til(g("ctrl_type"))), T(til(g("invscale_type"))), til(g("rel_cat")), til(g("attach_ix")),
                           til(g("attach_subix")), T(til(g("gpos"))), T(til(g("eval_spot_type"))), T(np.zeros(K * reps, np.float32)))
    return tso, sso, types


def consts(ops, d):
    c = ops.token_prior_params()
    (c.sigma_shape, c.sigma_invscale, c.sigma_attach, c.sigma_x, c.sigma_y, c.min_blur_sigma, c.max_blur_sigma) = \
        [float(v) for v in d["consts"]]
    return c


def test_marginals_match_reference_draws():
    from scipy.stats import ks_2samp
    from pybpl_b200 import ops
    d = load("token_samples")
    c = consts(ops, d)
    worst = 1.0
    ntests = 0
    for ti in range(int(d["n_types"])):
        tso, sso, types = expand(ops, d, ti, R)
        batch, out_types = ops.sample_tokens(tso, sso, types, seed=1234 + ti, c=c)
        S, K = len(d[f"t{ti}_invscale_type"]), len(d[f"t{ti}_nsub"])
        ctrl = batch.ctrl.view(R, S, 5, 2).cpu().numpy()
        inv = batch.invscale.view(R, S).cpu().numpy()
        pos = batch.position.view(R, K, 2).cpu().numpy()
        es = out_types.eval_spot.view(R, K).cpu().numpy()
        assert (inv > 0).all()
        cat = d[f"t{ti}_rel_cat"]
        assert ((es[:, cat == 3] >= 2.0) & (es[:, cat == 3] <= 6.0)).all() and (es[:, cat != 3] == 0).all()
        pairs = [(ctrl[:, s, q, e], d[f"t{ti}_ctrl"][:, s, q, e]) for s in range(S) for q in (0, 2, 4) for e in (0, 1)]
        pairs += [(inv[:, s], d[f"t{ti}_invscale"][:, s]) for s in range(S)]
        pairs += [(pos[:, k, e], d[f"t{ti}_position"][:, k, e]) for k in range(K) for e in (0, 1)]
        pairs += [(es[:, k], d[f"t{ti}_eval_spot"][:, k]) for k in range(K) if cat[k] == 3]
        for a, b in pairs:
            p = ks_2samp(a, b).pvalue
            worst = min(worst, p); ntests += 1
        # standardised shape residuals are N(0,1): mean and variance at N = R*S*10
        z = (ctrl - d[f"t{ti}_ctrl_type"][None]) / c.sigma_shape
        n = z.size
        assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    print(f"{ntests} KS tests against the reference's draws, smallest p-value {worst:.3g}")
    assert worst > 1e-4 / 1.0, worst        # ~100 tests with fixed seeds: a p below 1e-4 means a wrong distribution


def test_joint_structure_through_the_scorer():
    """The location of a stroke is drawn around the attach point of the strokes ALREADY drawn: the mean log-probability
    of the GPU's draws under score_token equals that of the reference's own draws (two-sample z-test), and the residual
    position - attach point, recomputed by the oracle from the drawn tokens, is N(0, sigma)."""
    from oracle import token_prior as tp
    from pybpl_b200 import ops
    d = load("token_samples")
    c = consts(ops, d)
    for ti in range(int(d["n_types"])):
        n_ref = d[f"t{ti}_ctrl"].shape[0]
        tso, sso, types = expand(ops, d, ti, R)
        batch, out_types = ops.sample_tokens(tso, sso, types, seed=99 + ti, c=c)
        ll_gpu = ops.score_token_prior(batch, out_types, c=c).cpu().numpy()
        assert np.isfinite(ll_gpu).all()
        # the reference's draws, scored by the same kernel
        tso_r, sso_r, types_r = expand(ops, d, ti, n_ref)
        rb = ops.TokenBatch(tso_r, sso_r, T(d[f"t{ti}_ctrl"].reshape(-1, 5, 2)), T(d[f"t{ti}_invscale"].reshape(-1)),
                            T(d[f"t{ti}_position"].reshape(-1, 2)), T(np.full(n_ref, 1e-4, np.float32)),
                            T(np.full(n_ref, 0.5, np.float32)))
        types_r.eval_spot = T(d[f"t{ti}_eval_spot"].reshape(-1))
        ll_ref = ops.score_token_prior(rb, types_r, c=c).cpu().numpy()
        zstat = (ll_gpu.mean() - ll_ref.mean()) / np.sqrt(ll_gpu.var() / len(ll_gpu) + ll_ref.var() / len(ll_ref))
        print(f"type {ti}: mean ll GPU draws {ll_gpu.mean():.3f}, reference draws {ll_ref.mean():.3f}, z = {zstat:.2f}")
        assert abs(zstat) < 4.5
        # and one token against the numpy oracle (the scorer itself is parity-tested elsewhere)
        one = dict(K=np.array([len(d[f"t{ti}_nsub"])]), nsub=d[f"t{ti}_nsub"], ctrl=batch.ctrl[:len(d[f"t{ti}_invscale_type"])].cpu().numpy(),
                   invscale=batch.invscale[:len(d[f"t{ti}_invscale_type"])].cpu().numpy(),
                   position=batch.position[:len(d[f"t{ti}_nsub"])].cpu().numpy(), sigma=np.array([0.5], np.float32),
                   eval_spot=out_types.eval_spot[:len(d[f"t{ti}_nsub"])].cpu().numpy(), consts=d["consts"],
                   **{k: d[f"t{ti}_{k}"] for k in ("ctrl_type", "invscale_type", "rel_cat", "attach_ix", "attach_subix", "gpos",
                                                  "eval_spot_type")})
        C = load("tokens")["basis_5_200"]
        ref = tp.score(one, np.float64, C0=C[0], C1=C[-1])[0]
        assert abs(ll_gpu[0] - ref) <= 1e-5 * abs(ref)


def test_streams_are_shardable_and_seeded():
    from pybpl_b200 import ops
    d = load("token_samples")
    c = consts(ops, d)
    n = 5000
    tso, sso, types = expand(ops, d, 1, n)
    whole, wt = ops.sample_tokens(tso, sso, types, seed=7, c=c)
    half = n // 2
    tso_a, sso_a, types_a = expand(ops, d, 1, half)
    a, at = ops.sample_tokens(tso_a, sso_a, types_a, seed=7, first_token=0, c=c)
    b, bt = ops.sample_tokens(tso_a, sso_a, types_a, seed=7, first_token=half, c=c)
    assert torch.equal(whole.ctrl, torch.cat([a.ctrl, b.ctrl])) and torch.equal(whole.position, torch.cat([a.position, b.position]))
    assert torch.equal(whole.invscale, torch.cat([a.invscale, b.invscale])) and torch.equal(wt.eval_spot, torch.cat([at.eval_spot, bt.eval_spot]))
    other, _ = ops.sample_tokens(tso, sso, types, seed=8, c=c)
    assert not torch.equal(whole.ctrl, other.ctrl)
    again, _ = ops.sample_tokens(tso, sso, types, seed=7, c=c)
    assert torch.equal(whole.ctrl, again.ctrl) and torch.equal(whole.position, again.position)
    # the drawn tokens render and score
    img = ops.pack_images(torch.zeros(105, 105, dtype=torch.uint8, device=DEV))
    ll, off = ops.score_tokens(whole, img)
    assert torch.isfinite(ll).all()
