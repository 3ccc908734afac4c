"""CPU oracle for the pyBPL render+score hot path -- TEST INFRASTRUCTURE ONLY.

`oracle/bpl_oracle.c` (plain C, fp32 and fp64 instantiations) restates the
reference algorithm (each function cites the /root/reference file:line it
follows); this module is its ctypes/numpy binding.  `oracle/token_prior.py` and
`oracle/type_prior.py` are numpy restatements of the two prior scores.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package.  The product (pybpl_b200/)
never does; it fails loudly when its CUDA library is missing.

Parity pinning: the reference's own tests hold no values for this path
(pybpl/tests/test_vanilla_to_motor.py:47-48 asserts shapes only), so the oracle
is pinned against tests/golden/*.npz, produced by tests/golden/make_golden.py
from the live reference (tests/test_oracle_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbpl_oracle.so")
_lib = None


class Params(ctypes.Structure):
    _fields_ = [("H", ctypes.c_int), ("W", ctypes.c_int), ("ink_pp", ctypes.c_float),
                ("ink_max_dist", ctypes.c_float), ("ink_ncon", ctypes.c_int), ("ink_a", ctypes.c_float),
                ("ink_b", ctypes.c_float), ("hinton", ctypes.c_int), ("fsize", ctypes.c_int)]


def build(force=False):
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in
                ("bpl_oracle.c", "bpl_oracle_impl.h", "bpl_oracle.h", "build.sh"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.bplo_loglik_f32.restype = ctypes.c_float
        _lib.bplo_loglik_f64.restype = ctypes.c_double
    return _lib


def set_threads(n):
    """OpenMP threads for the batch entry points (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
    return int(lib().bplo_set_threads(int(n)))


def default_params(hinton=False):
    ps = Params()
    lib().bplo_default_params(ctypes.byref(ps))
    ps.hinton = int(hinton)
    return ps


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def unpack_image(packed):
    """np.packbits rows (105,14) uint8 -> (105,105) uint8 {0,1}."""
    return np.unpackbits(packed, axis=-1)[..., :105].astype(np.uint8)


class Oracle:
    """numpy front-end; dtype np.float32 (the restatement) or np.float64 (noise floor)."""

    def __init__(self, dtype=np.float32, ps=None):
        self.dt = np.dtype(dtype)
        self.suf = "_f32" if self.dt == np.float32 else "_f64"
        self.c_real = ctypes.c_float if self.dt == np.float32 else ctypes.c_double
        self.ps = ps if ps is not None else default_params()
        self.L = lib()

    def _f(self, name):
        return getattr(self.L, name + self.suf)

    def _a(self, x):
        return None if x is None else np.ascontiguousarray(x, dtype=self.dt)

    # -- splines ----------------------------------------------------------
    def basis_table(self, nland, neval):
        C = np.empty((neval, nland), self.dt)
        self._f("bplo_basis_table")(nland, neval, _p(C))
        return C

    def basis_rows(self, nland, s):
        s = self._a(np.atleast_1d(s))
        C = np.empty((len(s), nland), self.dt)
        self._f("bplo_basis_rows")(nland, len(s), _p(s), _p(C))
        return C

    def spline_eval(self, C, Y):
        C = self._a(C); Y = self._a(Y)
        squeeze = Y.ndim == 2
        Y3 = Y[None] if squeeze else Y
        X = np.empty((Y3.shape[0], C.shape[0], 2), self.dt)
        self._f("bplo_spline_eval")(_p(C), C.shape[0], C.shape[1], Y3.shape[0], _p(Y3), _p(X))
        return X[0] if squeeze else X

    def adaptive_neval(self, C200, Y):
        C200 = self._a(C200); Y = self._a(Y)
        n = Y.shape[0]
        nev = np.empty(n, np.int32); dist = np.empty(n, self.dt)
        self._f("bplo_adaptive_neval")(_p(C200), C200.shape[1], n, _p(Y), _p(nev), _p(dist))
        return nev, dist

    def lstsq(self, C, X):
        """fit_bspline_to_traj for a batch: C [m,n], X [b,m,2] -> Y [b,n,2], residuals [b,2]."""
        C = self._a(C); X = self._a(X)
        b, m, n = X.shape[0], C.shape[0], C.shape[1]
        Y = np.empty((b, n, 2), self.dt); res = np.empty((b, 2), self.dt)
        self._f("bplo_lstsq")(_p(C), m, n, b, _p(X), _p(Y), _p(res))
        return Y, res

    def vanilla_to_motor(self, C, ctrl, invscale, pos):
        C = self._a(C); ctrl = self._a(ctrl); invscale = self._a(invscale); pos = self._a(pos)
        nsub, ncpt = ctrl.shape[0], ctrl.shape[1]
        motor = np.empty((nsub, C.shape[0], 2), self.dt); ms = np.empty((nsub, ncpt, 2), self.dt)
        self._f("bplo_vanilla_to_motor")(_p(C), C.shape[0], ncpt, nsub, _p(ctrl), _p(invscale), _p(pos),
                                         _p(motor), _p(ms))
        return motor, ms

    # -- rendering from raw motor-space trajectories ------------------------
    @staticmethod
    def _ragged(strokes):
        lens = np.array([len(s) for s in strokes], np.int32)
        off = np.zeros(len(lens) + 1, np.int32); off[1:] = np.cumsum(lens)
        return off

    def render(self, strokes, eps, sigma):
        off = self._ragged(strokes)
        pts = self._a(np.concatenate([np.asarray(s).reshape(-1, 2) for s in strokes]))
        pimg = np.empty((self.ps.H, self.ps.W), self.dt)
        f = self._f("bplo_render"); f.restype = ctypes.c_int
        o = f(ctypes.byref(self.ps), len(strokes), _p(off), _p(pts), self.c_real(eps), self.c_real(sigma), _p(pimg))
        return pimg, bool(o)

    def loglik(self, pimg, image):
        pimg = self._a(pimg); image = np.ascontiguousarray(image, np.uint8)
        return self.dt.type(self._f("bplo_loglik")(_p(pimg), _p(image), pimg.size))

    def point_indices(self, pts):
        pts = self._a(np.asarray(pts).reshape(-1, 2)); n = len(pts)
        m = np.empty(n, np.uint8); fl = np.empty((n, 2), np.int32); ce = np.empty((n, 2), np.int32)
        self._f("bplo_point_indices")(ctypes.byref(self.ps), n, _p(pts), _p(m), _p(fl), _p(ce))
        return m.astype(bool), fl, ce

    def render_grad(self, strokes, eps, sigma, image=None, gP=None):
        off = self._ragged(strokes)
        pts = self._a(np.concatenate([np.asarray(s).reshape(-1, 2) for s in strokes]))
        H, W = self.ps.H, self.ps.W
        pimg = np.empty((H, W), self.dt); g_pts = np.empty_like(pts)
        ll = self.c_real(0); ge = self.c_real(0); gs = self.c_real(0)
        image = None if image is None else np.ascontiguousarray(image, np.uint8)
        gP = self._a(gP)
        f = self._f("bplo_render_grad"); f.restype = ctypes.c_int
        o = f(ctypes.byref(self.ps), len(strokes), _p(off), _p(pts), self.c_real(eps), self.c_real(sigma),
              _p(image), _p(gP), _p(pimg), ctypes.byref(ll), _p(g_pts), ctypes.byref(ge), ctypes.byref(gs))
        return dict(pimg=pimg, ll=ll.value, off_page=bool(o), g_pts=g_pts, g_eps=ge.value, g_sigma=gs.value)

    # -- whole path, one token ---------------------------------------------
    def token(self, C, nsub, ctrl, invscale, position, eps, sigma, affine=None, image=None, gP=None, grad=False):
        C = self._a(C); ctrl = self._a(ctrl); invscale = self._a(invscale); position = self._a(position)
        nsub = np.ascontiguousarray(nsub, np.int32)
        S, ncpt, neval = ctrl.shape[0], ctrl.shape[1], C.shape[0]
        H, W = self.ps.H, self.ps.W
        pimg = np.empty((H, W), self.dt); motor = np.empty((S, neval, 2), self.dt)
        ll = self.c_real(0); off = ctypes.c_int(0)
        affine = self._a(affine); gP = self._a(gP)
        image = None if image is None else np.ascontiguousarray(image, np.uint8)
        res = {}
        if grad:
            g_ctrl = np.zeros_like(ctrl); g_inv = np.zeros_like(invscale); g_pos = np.zeros_like(position)
            g_aff = np.zeros(4, self.dt); gs = self.c_real(0); ge = self.c_real(0)
            gargs = (_p(g_ctrl), _p(g_inv), _p(g_pos), _p(g_aff), ctypes.byref(gs), ctypes.byref(ge))
        else:
            gargs = (None,) * 6
        self._f("bplo_token")(ctypes.byref(self.ps), ncpt, neval, _p(C), len(nsub), _p(nsub), _p(ctrl), _p(invscale),
                              _p(position), _p(affine), self.c_real(eps), self.c_real(sigma), _p(image), _p(gP),
                              _p(pimg), ctypes.byref(ll), ctypes.byref(off), _p(motor), *gargs)
        res.update(pimg=pimg, ll=ll.value, off_page=bool(off.value), motor=motor)
        if grad:
            res.update(g_ctrl=g_ctrl, g_invscale=g_inv, g_position=g_pos, g_affine=g_aff, g_sigma=gs.value,
                       g_eps=ge.value)
        return res

    # -- packed batch (threads over tokens) ----------------------------------
    def token_batch(self, C, tok_stroke_off, stroke_sub_off, ctrl, invscale, position, eps, sigma, images,
                    img_idx=None, affine=None, grad=False, want_pimg=False):
        C = self._a(C); ctrl = self._a(ctrl); invscale = self._a(invscale); position = self._a(position)
        eps = self._a(eps); sigma = self._a(sigma); affine = self._a(affine)
        tso = np.ascontiguousarray(tok_stroke_off, np.int32); sso = np.ascontiguousarray(stroke_sub_off, np.int32)
        images = np.ascontiguousarray(images, np.uint8)
        img_idx = None if img_idx is None else np.ascontiguousarray(img_idx, np.int32)
        P = len(tso) - 1; H, W = self.ps.H, self.ps.W
        ll = np.empty(P, self.dt); off = np.empty(P, np.uint8)
        pimg = np.empty((P, H, W), self.dt) if want_pimg else None
        res = {}
        if grad:
            g = dict(g_ctrl=np.zeros_like(ctrl), g_invscale=np.zeros_like(invscale),
                     g_position=np.zeros_like(position), g_affine=np.zeros((P, 4), self.dt),
                     g_sigma=np.zeros(P, self.dt), g_eps=np.zeros(P, self.dt))
            gargs = tuple(_p(g[k]) for k in ("g_ctrl", "g_invscale", "g_position", "g_affine", "g_sigma", "g_eps"))
            res.update(g)
        else:
            gargs = (None,) * 6
        self._f("bplo_token_batch")(ctypes.byref(self.ps), ctrl.shape[1], C.shape[0], _p(C), P, _p(tso), _p(sso),
                                    _p(ctrl), _p(invscale), _p(position), _p(affine), _p(eps), _p(sigma),
                                    _p(images), _p(img_idx), _p(ll), _p(off), _p(pimg), *gargs)
        res.update(ll=ll, off_page=off.astype(bool))
        if want_pimg:
            res["pimg"] = pimg
        return res


def target_image_for(w, oracle, C, token=0, seed=99):
    """Binary target for a packed batch `w` (pybpl_b200.workload layout) = Bernoulli sample of one token's probability
    map, rendered by this CPU oracle (tests, smoke() and the bench's CPU arm)."""
    sub_k0, sub_k1 = int(w["tok_stroke_off"][token]), int(w["tok_stroke_off"][token + 1])
    s0, s1 = int(w["stroke_sub_off"][sub_k0]), int(w["stroke_sub_off"][sub_k1])
    nsub = np.diff(w["stroke_sub_off"][sub_k0:sub_k1 + 1])
    r = oracle.token(C, nsub, w["ctrl"][s0:s1], w["invscale"][s0:s1], w["position"][sub_k0:sub_k1], float(w["eps"][token]),
                     float(w["sigma"][token]))
    rng = np.random.default_rng(seed)
    return (rng.random((105, 105)) < r["pimg"]).astype(np.uint8)
