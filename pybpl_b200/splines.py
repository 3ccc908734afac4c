"""Drop-in for the evaluation half of pybpl/splines.py (bspline_gen_s, coefficient_mat,
get_stk_from_bspline) plus the `bspline_eval` alias the north-star names.  Same signatures, argument
meaning and assertions; the arithmetic runs in libpybpl_b200.so on the GPU."""
import torch

from . import ops
from .parameters import Parameters

__all__ = ['bspline_gen_s', 'coefficient_mat', 'get_stk_from_bspline', 'bspline_eval', 'fit_bspline_to_traj']

PM = Parameters()


def _compute_device(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def bspline_gen_s(nland, neval=200, device=None):
    """Time points for evaluating a spline (pybpl/splines.py:57-69): (s, lb, ub)."""
    lb = float(2)
    ub = float(nland + 1)
    s = torch.linspace(lb, ub, neval, device=device)
    return s, lb, ub


def coefficient_mat(nland, neval=None, s=None, device=None):
    """B-spline coefficient matrix (pybpl/splines.py:72-93).  Not cached on tensor identity: the
    reference's lru_cache returns a stale matrix after an in-place update of `s`."""
    if s is None:
        assert neval is not None, 'neval must be provided when s not provided.'
        dev = torch.device("cpu") if device is None else torch.device(device)
        if dev.type == "cuda":
            return ops.basis_table(nland, neval, dev)
        return torch.from_numpy(ops._basis_host(nland, neval)).clone()
    if s.dim() == 0:
        s = s.view(1)
    assert s.dim() == 1
    return ops.basis_rows(nland, s)


def _check_input(x):
    assert torch.is_tensor(x)
    assert x.dim() == 2
    assert x.size(1) == 2


def get_stk_from_bspline(Y, neval=None, s=None):
    """Evaluate a B-spline: Y [nland,2] -> X [neval,2] (pybpl/splines.py:108-138).

    With neither `neval` nor `s`, neval is chosen from the trajectory length exactly as the reference
    does (integer work: clamp(ceil(dist/1.5), 10, 200))."""
    _check_input(Y)
    nland = Y.size(0)
    dev = _compute_device(Y)
    Yd = Y.to(dev)
    if neval is None and s is None:
        nev, _ = ops.adaptive_neval(Yd.detach().unsqueeze(0), PM.spline_grain, PM.spline_min_neval,
                                    PM.spline_max_neval)
        neval = int(nev.item())
    if s is not None:
        sv = s.view(1) if s.dim() == 0 else s
        C = ops.basis_rows(nland, sv).to(dev)
        if s.requires_grad:
            # d C / d s: analytic derivative of the normalised cubic basis
            C = _BasisRowsWithGrad.apply(sv.to(dev), nland, C)
    else:
        C = ops.basis_table(nland, int(neval), dev)
    X = ops.spline_eval(C, Yd.unsqueeze(0))[0]
    return X.to(Y.device)


def fit_bspline_to_traj(X, nland, s=None, include_resid=False):
    """Least-squares B-spline of a trajectory: X [neval,2] -> Y [nland,2] (pybpl/splines.py:141-171).
    The residuals are the squared 2-norms of C Y - X per coordinate ([2]); like torch.linalg.lstsq they are
    empty when neval <= nland."""
    _check_input(X)
    neval = X.size(0)
    dev = _compute_device(X)
    if s is not None:
        sv = s.view(1) if s.dim() == 0 else s
        C = ops.basis_rows(nland, sv).to(dev)
    else:
        C = ops.basis_table(nland, neval, dev)
    Y, res = ops.spline_fit(C, X.to(dev).unsqueeze(0))
    Y = Y[0].to(X.device)
    if include_resid:
        return Y, (res[0].to(X.device) if neval > nland else torch.empty(0, device=X.device))
    return Y


def fit_smooth_stk(stroke, loss_threshold=0.3, max_nland=100, unif_space=None):
    """The "smooth" version of a stroke: the spline with the fewest control points whose mean squared residual stays
    below `loss_threshold`, evaluated with the adaptive number of points (pybpl/bottomup/initialize/util.py:13-31: the
    refit loop over fit_bspline_to_traj, then get_stk_from_bspline).  The least-squares fits and the evaluation run on
    the GPU; the loop itself is the reference's (first nland under the threshold, in increasing order).

    stroke: [n,2] numpy array or tensor.  unif_space: the resampling function the reference applies first
    (pybpl/data/unif_space.py, host-side data preparation, outside this package); None uses the reference's when it is
    importable and otherwise expects `stroke` to be uniformly spaced already.  Returns a numpy array like the reference."""
    import numpy as np
    ntraj = len(stroke)
    if ntraj == 1:
        return stroke
    if unif_space is None:
        try:
            from pybpl.data.unif_space import unif_space
        except Exception:
            unif_space = None
    if unif_space is not None:
        stroke = unif_space(stroke)
    stroke = torch.as_tensor(np.asarray(stroke), dtype=torch.float)
    Xd = stroke.to(_compute_device(stroke))
    for nland in range(1, min(max_nland, ntraj + 1)):
        spline, residuals = fit_bspline_to_traj(Xd, nland, include_resid=True)
        loss = torch.sum(residuals) / ntraj
        if loss.item() < loss_threshold:
            break
    return get_stk_from_bspline(spline).cpu().numpy()


def bspline_eval(sval, cpts):
    """Evaluate the spline with control points `cpts` [nland,2] at times `sval` (scalar or [n])."""
    return get_stk_from_bspline(cpts, s=sval)


def _basis_rows_derivative(nland, s):
    """d C / d s of coefficient_mat's rows (pybpl/splines.py:19-54, 72-93): the piecewise-cubic basis B_k(s) / 6,
    row-normalised, differentiated analytically: c = B / sum B, c' = (B' - c sum B') / sum B.  s: [n] -> [n, nland]."""
    s = s.detach().to(torch.float64).view(-1, 1)
    d = s - torch.arange(nland, dtype=torch.float64, device=s.device).view(1, -1)
    B = torch.zeros_like(d); dB = torch.zeros_like(d)
    m = (d >= 0) & (d < 1); e = d
    B = torch.where(m, e ** 3, B); dB = torch.where(m, 3 * e ** 2, dB)
    m = (d >= 1) & (d < 2); e = d - 1
    B = torch.where(m, -3 * e ** 3 + 3 * e ** 2 + 3 * e + 1, B); dB = torch.where(m, -9 * e ** 2 + 6 * e + 3, dB)
    m = (d >= 2) & (d < 3); e = d - 2
    B = torch.where(m, 3 * e ** 3 - 6 * e ** 2 + 4, B); dB = torch.where(m, 9 * e ** 2 - 12 * e, dB)
    m = (d >= 3) & (d < 4); e = d - 3
    B = torch.where(m, (1 - e) ** 3, B); dB = torch.where(m, -3 * (1 - e) ** 2, dB)
    tot = B.sum(1, keepdim=True)
    c = B / tot
    return ((dB - c * dB.sum(1, keepdim=True)) / tot).to(torch.float32)


class _BasisRowsWithGrad(torch.autograd.Function):
    """Attaches d C / d s (analytic) so that eval_spot-style parameters receive a gradient, which the reference intends
    but loses to its cache."""

    @staticmethod
    def forward(ctx, s, nland, C):
        ctx.nland = nland
        ctx.save_for_backward(s)
        return C.clone()

    @staticmethod
    def backward(ctx, gC):
        (s,) = ctx.saved_tensors
        dC = _basis_rows_derivative(ctx.nland, s).to(gC.device)
        return (gC * dC).sum(1), None, None
