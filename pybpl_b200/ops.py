"""Batched operators over the C-ABI library: packed ragged batches, autograd.Functions.

torch is plumbing here (device memory, streams, autograd graph); every number is produced by
the CUDA kernels in pybpl_b200/csrc through libpybpl_b200.so.
"""
import ctypes
import math
import functools
import os

import numpy as np
import torch

from . import _lib
from ._lib import IMSIZE, IMG_WORDS, NCPT, check, lib

_HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(t, name):
    if not t.is_cuda:
        raise _lib.BplError(f"{name} must be a CUDA tensor: pybpl_b200 has no CPU path")


def _f32(t, device):
    return torch.as_tensor(t, dtype=torch.float32, device=device).contiguous()


def _i32(t, device):
    return torch.as_tensor(t, dtype=torch.int32, device=device).contiguous()


def make_params(ps=None):
    """bpl_params from None or a reference-style Parameters object (pybpl/parameters.py:19-41)."""
    if isinstance(ps, _lib.Params):
        return ps
    p = _lib.default_params()
    if ps is not None:
        imsize = tuple(int(v) for v in getattr(ps, "imsize", (IMSIZE, IMSIZE)))
        fsize = int(getattr(ps, "fsize", 11))
        if not (1 <= fsize <= 11 and fsize % 2 == 1):
            raise NotImplementedError("pybpl_b200 kernels are built for odd fsize <= 11")
        if imsize != (IMSIZE, IMSIZE):
            # Any other size runs through the run-time-geometry kernels (render_strokes / rendering.render_image /
            # CharacterImageDist); the fused 105 x 105 entry points refuse a Params with imsize != 105.
            if len(imsize) != 2 or min(imsize) < 1 or imsize[0] * imsize[1] > (1 << 24):
                raise ValueError(f"imsize must be (H, W) with H, W >= 1 and H * W <= 2^24; got {imsize}")
            p.imsize = 0
            p.hw = imsize
        p.fsize = fsize
        p.ink_ncon = int(getattr(ps, "ink_ncon", p.ink_ncon))
        p.ink_pp = float(getattr(ps, "ink_pp", p.ink_pp))
        p.ink_max_dist = float(getattr(ps, "ink_max_dist", p.ink_max_dist))
        p.ink_a = float(getattr(ps, "ink_a", p.ink_a))
        p.ink_b = float(getattr(ps, "ink_b", p.ink_b))
        mode = getattr(ps, "broaden_mode", "Lake")
        if mode not in ("Lake", "Hinton"):
            raise Exception("'broaden_mode' must be either 'Lake' or 'Hinton'")     # rendering.py:157-158
        p.hinton = int(mode == "Hinton")
    return p


@functools.lru_cache(maxsize=None)
def _basis_host(nland, neval):
    """coefficient_mat(nland, neval) (pybpl/splines.py:72-93) as a numpy array.

    (5, 200) -- the only table on the render path -- is the reference's own output shipped as data
    (pybpl_b200/data/basis_5_200.npy, see tests/golden/make_golden.py:make_assets); all others are
    computed by the library and agree with torch's to ~4e-7."""
    asset = os.path.join(_HERE, "data", f"basis_{nland}_{neval}.npy")
    if os.path.exists(asset):
        return np.load(asset).astype(np.float32)
    C = np.empty((neval, nland), np.float32)
    check(lib().bpl_basis_table_host(nland, neval, C.ctypes.data_as(ctypes.c_void_p)))
    return C


_basis_dev = {}


def basis_table(nland, neval, device):
    key = (nland, neval, str(device))
    t = _basis_dev.get(key)
    if t is None:
        t = torch.from_numpy(_basis_host(nland, neval)).to(device)
        _basis_dev[key] = t
    return t


def basis_rows(nland, s):
    """Rows of the coefficient matrix at explicit times s (1-D tensor); host computation, result on s.device."""
    sh = s.detach().to("cpu", torch.float32).contiguous().numpy()
    C = np.empty((len(sh), nland), np.float32)
    check(lib().bpl_basis_rows_host(nland, len(sh), sh.ctypes.data_as(ctypes.c_void_p),
                                    C.ctypes.data_as(ctypes.c_void_p)))
    return torch.from_numpy(C).to(s.device)


# ---------------------------------------------------------------------------------------------
# target images
# ---------------------------------------------------------------------------------------------
class PackedImages:
    """T binary 105x105 images bit-packed on the device ([T,345] int32 words; pixel i = bit i&31 of word i>>5)."""

    def __init__(self, bits, bad):
        self.bits = bits
        self._bad = bad
        self.n = bits.shape[0]
        self.device = bits.device

    def validate(self):
        # torch's Bernoulli.log_prob raises ValueError for values outside {0,1} (image_dist.py:76-77)
        if int(self._bad.item()) != 0:
            raise ValueError("Expected value argument to be within the support (Boolean()) of the "
                             "distribution Bernoulli(), but found invalid values")
        return self


def pack_images(images, validate=True):
    """images: [T,105,105] or [105,105], uint8/bool/float on a CUDA device."""
    _need_cuda(images, "images")
    if images.dim() == 2:
        images = images.unsqueeze(0)
    if images.shape[-2:] != (IMSIZE, IMSIZE):
        raise ValueError("images must be 105x105")
    T = images.shape[0]
    dev = images.device
    bits = torch.empty((T, IMG_WORDS), dtype=torch.int32, device=dev)
    bad = torch.zeros((), dtype=torch.int32, device=dev)
    if images.dtype in (torch.uint8, torch.bool):
        src = images.contiguous().view(torch.uint8)
        check(lib().bpl_pack_images_u8(_ptr(src), T, _ptr(bits), _ptr(bad), _stream(dev)))
    else:
        src = images.to(torch.float32).contiguous()
        check(lib().bpl_pack_images_f32(_ptr(src), T, _ptr(bits), _ptr(bad), _stream(dev)))
    out = PackedImages(bits, bad)
    return out.validate() if validate else out


# ---------------------------------------------------------------------------------------------
# packed ragged batches
# ---------------------------------------------------------------------------------------------
class TokenBatch:
    """P character tokens in control-point form (what CharacterImageDist.get_pimg reads).

    tok_stroke_off [P+1], stroke_sub_off [K+1] int32 CSR offsets; ctrl [S,5,2]; invscale [S];
    position [K,2]; affine [P,4] or None; eps [P]; sigma [P].  All on one CUDA device.
    """

    def __init__(self, tok_stroke_off, stroke_sub_off, ctrl, invscale, position, eps, sigma, affine=None, neval=200,
                 check_limits=True, use_order=True):
        dev = ctrl.device
        _need_cuda(ctrl, "ctrl")
        if check_limits:
            tso = torch.as_tensor(tok_stroke_off).cpu().numpy().astype(np.int64)
            sso = torch.as_tensor(stroke_sub_off).cpu().numpy().astype(np.int64)
            if len(tso) > 1:
                if np.diff(tso).max() > _lib.MAX_STROKES or np.diff(sso[tso]).max() > _lib.MAX_SUBSTROKES:
                    raise _lib.BplError("a token exceeds 32 strokes / 128 sub-strokes (BPL_MAX_*)")
        self.tok_stroke_off = _i32(tok_stroke_off, dev)
        self.stroke_sub_off = _i32(stroke_sub_off, dev)
        self.ctrl = ctrl if ctrl.dtype == torch.float32 and ctrl.is_contiguous() else _f32(ctrl, dev)
        self.invscale = _f32(invscale, dev) if not isinstance(invscale, torch.Tensor) else invscale
        self.position = position
        self.eps = eps
        self.sigma = sigma
        self.affine = affine
        self.neval = int(neval)
        self.P = self.tok_stroke_off.numel() - 1
        self.K = self.stroke_sub_off.numel() - 1
        self.S = self.ctrl.shape[0]
        if self.ctrl.shape[1:] != (NCPT, 2):
            raise _lib.BplError("fused path takes ncpt=5 control points per sub-stroke")
        self.device = dev
        self._sub_token = None
        self._stroke_token = None
        self._order = None
        self._order_cost = None
        self._cycles = None             # cycles per token of the last fused backward (team kernel), see order_ptr
        self._cycles_valid = False
        self.use_order = use_order      # False: index order (streamed chunks: the order kernel would cost more than it saves)

    # Processing order for the persistent kernels: expensive tokens first, so that a launch of a few thousand tokens
    # (a few dozen per SM) does not end with one SM still busy on a large one.  Two keys, each computed once per batch
    # and kept (an order is only a heuristic, results do not depend on it):
    #   estimated cost   : area of the box of the control polygons + sub-stroke count (bpl_order_tokens_by_cost, from the
    #                      geometry at that time) -- the scoring calls and the fused backward: 1 024 parses 3.67 -> 4.04 M/s
    #                      fwd+bwd; 4 096 tokens forward 16.5 M/s in index order, 17.6 by sub-stroke count, 17.9 by this key
    #                      (18.2 when ordered by MEASURED cycles per token, tools/order_upper_bound.py)
    #   sub-stroke count : offsets only, one small kernel -- callers without the geometry at hand (token_points, staging)
    # Large batches keep index order.
    ORDER_MAX_TOKENS = 65536      # measured gain of an order at all: +6.7 % at 4 096 tokens, +2.5 % at 8 192, +1.2 % at 16 384

    def order_ptr(self, b=None):
        """b: the batch's C struct (geometry filled in) for the cost order of the fused backward; None = by sub-stroke count."""
        if not self.use_order or self.P > self.ORDER_MAX_TOKENS or self.P < 2 or os.environ.get("BPL_NO_ORDER"):      # (env: A/B aid)
            return None
        if b is not None and self.S > 0 and not os.environ.get("BPL_ORDER_BY_S"):      # (env: A/B aid)
            if self._cycles_valid and not os.environ.get("BPL_NO_CYCLE_FEEDBACK"):
                # the batch has been through the fused backward before: its tokens' MEASURED cycles order this launch
                # (a fit evaluates the same parses a hundred times; +5 % over the estimate at 1 024 parses)
                if self._order_cost is None:
                    self._order_cost = torch.empty(self.P, dtype=torch.int32, device=self.device)
                    self._keys = torch.empty(self.P, dtype=torch.int32, device=self.device)
                check(lib().bpl_order_tokens_by_cycles(self._cycles.data_ptr(), self.P, self._keys.data_ptr(),
                                                       self._order_cost.data_ptr(), _stream(self.device)))
                return self._order_cost.data_ptr()
            if self._order_cost is None:
                order = torch.empty(self.P, dtype=torch.int32, device=self.device)
                self._keys = torch.empty(self.P, dtype=torch.int32, device=self.device)
                check(lib().bpl_order_tokens_by_cost(ctypes.byref(make_params(None)), ctypes.byref(b), self._keys.data_ptr(),
                                                     order.data_ptr(), _stream(self.device)))
                self._order_cost = order
            return self._order_cost.data_ptr()
        if self._order is None:
            order = torch.empty(self.P, dtype=torch.int32, device=self.device)
            o = _lib.TokenBatchC()
            o.n_tokens = self.P
            o.tok_stroke_off = self.tok_stroke_off.data_ptr(); o.stroke_sub_off = self.stroke_sub_off.data_ptr()
            check(lib().bpl_order_tokens(ctypes.byref(o), order.data_ptr(), _stream(self.device)))
            self._order = order
        return self._order.data_ptr()

    # token index of every sub-stroke / stroke (for scaling gradients by the per-token cotangent)
    @property
    def sub_token(self):
        if self._sub_token is None:
            tso = self.tok_stroke_off.long()
            sub_off = self.stroke_sub_off.long()[tso]          # [P+1] token -> sub-strokes
            self._sub_token = torch.repeat_interleave(torch.arange(self.P, device=self.device),
                                                      sub_off[1:] - sub_off[:-1], output_size=self.S)
        return self._sub_token

    @property
    def stroke_token(self):
        if self._stroke_token is None:
            tso = self.tok_stroke_off.long()
            self._stroke_token = torch.repeat_interleave(torch.arange(self.P, device=self.device),
                                                         tso[1:] - tso[:-1], output_size=self.K)
        return self._stroke_token

    def c_struct(self, ctrl, invscale, position, affine, eps, sigma, cost_order=False):
        b = _lib.TokenBatchC()
        b.n_tokens = self.P; b.ncpt = NCPT; b.neval = self.neval
        b.tok_stroke_off = self.tok_stroke_off.data_ptr(); b.stroke_sub_off = self.stroke_sub_off.data_ptr()
        b.ctrl = ctrl.data_ptr(); b.invscale = invscale.data_ptr(); b.position = position.data_ptr()
        b.affine = None if affine is None else affine.data_ptr()
        b.eps = eps.data_ptr(); b.sigma = sigma.data_ptr()
        b.basis = basis_table(NCPT, self.neval, self.device).data_ptr()
        b.order = None
        b.cycles_out = None
        b.order = self.order_ptr(b if cost_order else None)
        return b

    def cycles_ptr(self):
        """Device buffer the team backward writes every token's cycles to (small batches only: the feedback pays where a
        launch is a few tokens per team deep)."""
        if not self.use_order or self.P > self.ORDER_MAX_TOKENS or self.P < 1024:
            return None
        if self._cycles is None:
            self._cycles = torch.zeros(self.P, dtype=torch.int32, device=self.device)
        return self._cycles.data_ptr()


class HostStagedBatch:
    """A token batch (+ one target image) staged in ONE pinned host buffer, so that a step needs a single
    host->device copy: [tok_stroke_off | stroke_sub_off | ctrl | invscale | position | eps | sigma | image u8],
    every section 16-byte aligned.  `upload()` copies it to a device buffer of the same layout on the current stream
    (asynchronously) and returns (TokenBatch, image tensor) whose tensors are views of that buffer."""

    _ORDER = (("tok_stroke_off", torch.int32), ("stroke_sub_off", torch.int32), ("ctrl", torch.float32),
              ("invscale", torch.float32), ("position", torch.float32), ("eps", torch.float32), ("sigma", torch.float32),
              ("order", torch.int32))

    def __init__(self, w, image_u8, device):
        self.device = torch.device(device)
        self.neval = int(w.get("neval", 200))
        self.layout = []
        off = 0
        arrs = []
        # the processing order (expensive tokens first, TokenBatch.order_ptr) depends only on the offsets: made here on
        # the host once and shipped with the inputs, instead of one more kernel per step
        tso = np.asarray(w["tok_stroke_off"], np.int64)
        S_tok = np.diff(np.asarray(w["stroke_sub_off"], np.int64)[tso])
        w = dict(w, order=np.argsort(-S_tok, kind="stable").astype(np.int32))
        for name, dt in self._ORDER:
            a = torch.from_numpy(np.ascontiguousarray(w[name])).to(dt).contiguous()
            arrs.append(a)
            nbytes = a.numel() * a.element_size()
            self.layout.append((name, dt, tuple(a.shape), off, nbytes))
            off = (off + nbytes + 15) // 16 * 16
        img = torch.as_tensor(image_u8, dtype=torch.uint8).reshape(-1).contiguous()
        self.img_off, self.img_bytes = off, img.numel()
        self.nbytes = (off + self.img_bytes + 15) // 16 * 16
        self.host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
        for a, (_, _, _, o, nb) in zip(arrs, self.layout):
            self.host[o:o + nb] = a.view(torch.uint8).reshape(-1)
        self.host[self.img_off:self.img_off + self.img_bytes] = img
        self.devs = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]

    def upload(self, slot=0):
        """Asynchronous H2D into device buffer `slot` (two slots allow the next step's copy to overlap this step's
        kernels when issued on a second stream)."""
        dev = self.devs[slot]
        dev.copy_(self.host, non_blocking=True)
        return self.views(slot)

    def views(self, slot=0):
        dev = self.devs[slot]
        v = {}
        for name, dt, shape, o, nb in self.layout:
            v[name] = dev[o:o + nb].view(dt).view(shape)
        batch = TokenBatch(v["tok_stroke_off"], v["stroke_sub_off"], v["ctrl"], v["invscale"], v["position"], v["eps"],
                           v["sigma"], neval=self.neval, check_limits=False)
        if 2 <= batch.P <= TokenBatch.ORDER_MAX_TOKENS:
            batch._order = v["order"]
        image = dev[self.img_off:self.img_off + self.img_bytes].view(IMSIZE, IMSIZE)
        return batch, image


class HostSweepStager:
    """Streams a HOST-resident numpy token batch (pybpl_b200.workload layout) to the GPU in chunks of `chunk` tokens.

    pack(c, slot) is the host-side ragged packing of chunk c: the chunk's slices of the CSR offsets (rebased to the
    chunk) and of ctrl / invscale / position / eps / sigma are copied into ONE pinned staging buffer (16-byte aligned
    sections sized for the largest chunk); upload(slot) is one asynchronous H2D copy of that buffer; batch(c, slot)
    views the device copy as a TokenBatch.  Two slots let the packing / copy of chunk c+1 overlap the kernels of c."""

    _SECTIONS = (("tok_stroke_off", np.int32, 1), ("stroke_sub_off", np.int32, 1), ("ctrl", np.float32, 2 * NCPT),
                 ("invscale", np.float32, 1), ("position", np.float32, 2), ("eps", np.float32, 1), ("sigma", np.float32, 1))

    def __init__(self, w, device, chunk=65536, slots=2, pin=True, threads=None):
        # packing threads: the host cores this rank can count on (numpy's copies release the GIL); one rank of eight on a
        # 16-core box packs with two threads, a single rank with up to eight
        if threads is None:
            threads = max(1, min(8, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))))
        self.threads = int(threads)
        self._pool = None
        if self.threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.threads)
        self.device = torch.device(device)
        self.w = w
        self.neval = int(w.get("neval", 200))
        P = int(w["P"])
        tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
        self.chunks = []
        for c0 in range(0, P, chunk):
            c1 = min(P, c0 + chunk)
            k0, k1 = int(tso[c0]), int(tso[c1])
            self.chunks.append((c0, c1, k0, k1, int(sso[k0]), int(sso[k1])))
        n = max((c[1] - c[0] for c in self.chunks), default=0)
        K = max((c[3] - c[2] for c in self.chunks), default=0)
        S = max((c[5] - c[4] for c in self.chunks), default=0)
        rows = dict(tok_stroke_off=n + 1, stroke_sub_off=K + 1, ctrl=S, invscale=S, position=K, eps=n, sigma=n)
        self.layout, off = {}, 0
        for name, dt, width in self._SECTIONS:
            nb = rows[name] * width * 4
            self.layout[name] = (off, dt, width)
            off = (off + nb + 15) // 16 * 16
        self.nbytes = max(off, 16)
        self.host = [torch.empty(self.nbytes, dtype=torch.uint8) for _ in range(slots)]
        if pin:
            self.host = [h.pin_memory() for h in self.host]
        self.host_np = [h.numpy() for h in self.host]
        self.dev = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(slots)]

    def _host_view(self, slot, name, rows):
        off, dt, width = self.layout[name]
        v = self.host_np[slot][off:off + rows * width * 4].view(dt)
        return v.reshape(rows, width) if width > 1 else v

    def pack(self, c, slot):
        c0, c1, k0, k1, s0, s1 = self.chunks[c]
        w = self.w
        np.subtract(w["tok_stroke_off"][c0:c1 + 1], k0, out=self._host_view(slot, "tok_stroke_off", c1 - c0 + 1))
        np.subtract(w["stroke_sub_off"][k0:k1 + 1], s0, out=self._host_view(slot, "stroke_sub_off", k1 - k0 + 1))
        dst, src = self._host_view(slot, "ctrl", s1 - s0), w["ctrl"][s0:s1].reshape(s1 - s0, 2 * NCPT)
        futs = []
        if self._pool is not None and s1 - s0 >= 4096:      # the control points are 87 % of the bytes: split over the pool
            n = s1 - s0
            cuts = [n * i // self.threads for i in range(self.threads + 1)]
            futs = [self._pool.submit(np.copyto, dst[a:b], src[a:b]) for a, b in zip(cuts[1:-1], cuts[2:])]
            np.copyto(dst[:cuts[1]], src[:cuts[1]])
        else:
            np.copyto(dst, src)
        np.copyto(self._host_view(slot, "invscale", s1 - s0), w["invscale"][s0:s1])
        np.copyto(self._host_view(slot, "position", k1 - k0), w["position"][k0:k1])
        np.copyto(self._host_view(slot, "eps", c1 - c0), w["eps"][c0:c1])
        np.copyto(self._host_view(slot, "sigma", c1 - c0), w["sigma"][c0:c1])
        for f in futs:
            f.result()

    def upload(self, slot):
        self.dev[slot].copy_(self.host[slot], non_blocking=True)

    def batch(self, c, slot):
        c0, c1, k0, k1, s0, s1 = self.chunks[c]
        d = self.dev[slot]

        def view(name, rows):
            off, dt, width = self.layout[name]
            t = d[off:off + rows * width * 4].view(torch.int32 if dt == np.int32 else torch.float32)
            return t.view(rows, width) if width > 1 else t
        return TokenBatch(view("tok_stroke_off", c1 - c0 + 1), view("stroke_sub_off", k1 - k0 + 1),
                          view("ctrl", s1 - s0).view(s1 - s0, NCPT, 2), view("invscale", s1 - s0), view("position", k1 - k0),
                          view("eps", c1 - c0), view("sigma", c1 - c0), neval=self.neval, check_limits=False, use_order=False)


def _targets(images, img_idx):
    t = _lib.Targets()
    if images is not None:
        t.image_bits = images.bits.data_ptr()
        t.img_idx = None if img_idx is None else img_idx.data_ptr()
        t.n_images = images.n
    return t


def _c32(t):
    return t.detach().to(torch.float32).contiguous()


class _ScoreTokens(torch.autograd.Function):
    """ll[p] = score_image(token p, image[img_idx[p]]) with the fused backward.

    When any input needs a gradient the forward runs the fused forward+backward kernel once and keeps
    d ll / d inputs; backward() only scales them by the incoming cotangent (one scalar per token)."""

    @staticmethod
    def forward(ctx, ctrl, invscale, position, affine, eps, sigma, batch, images, img_idx, ps, grad_on=True):
        dev = batch.device
        P = batch.P
        ctrl_c, inv_c, pos_c = _c32(ctrl), _c32(invscale), _c32(position)
        aff_c = None if affine is None else _c32(affine)
        eps_c, sig_c = _c32(eps), _c32(sigma)
        ll = torch.empty(P, dtype=torch.float32, device=dev)
        off = torch.empty(P, dtype=torch.uint8, device=dev)
        out = _lib.FwdOut(ll.data_ptr(), off.data_ptr(), None)
        t = _targets(images, img_idx)
        # grad_on = torch.is_grad_enabled() of the CALLER (inside forward() it is always off): leaves evaluated under
        # no_grad take the cheaper forward-only kernel
        need = grad_on and any(ctx.needs_input_grad[:6])
        b = batch.c_struct(ctrl_c, inv_c, pos_c, aff_c, eps_c, sig_c, cost_order=True)
        if need:
            g_ctrl = torch.empty_like(ctrl_c); g_inv = torch.empty_like(inv_c); g_pos = torch.empty_like(pos_c)
            g_aff = None if aff_c is None else torch.empty_like(aff_c)
            g_eps = torch.empty(P, dtype=torch.float32, device=dev); g_sig = torch.empty_like(g_eps)
            g = _lib.Grads(None, None, g_ctrl.data_ptr(), g_inv.data_ptr(), g_pos.data_ptr(),
                           None if g_aff is None else g_aff.data_ptr(), None, g_eps.data_ptr(), g_sig.data_ptr())
            b.cycles_out = batch.cycles_ptr()
            check(lib().bpl_score_tokens_grad(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                              ctypes.byref(g), _stream(dev)))
            batch._cycles_valid = b.cycles_out is not None and not os.environ.get("BPL_BWD_CANVAS")      # (the full-canvas kernel does not report cycles)
            ctx.batch = batch
            ctx.has_aff = g_aff is not None
            ctx.save_for_backward(g_ctrl, g_inv, g_pos, g_eps, g_sig, *([g_aff] if g_aff is not None else []))
        else:
            check(lib().bpl_score_tokens(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                         _stream(dev)))
        ctx.mark_non_differentiable(off)
        return ll, off

    @staticmethod
    def backward(ctx, g_ll, _g_off):
        saved = ctx.saved_tensors
        g_ctrl, g_inv, g_pos, g_eps, g_sig = saved[:5]
        g_aff = saved[5] if ctx.has_aff else None
        b = ctx.batch
        gs = _c32(g_ll)
        # every gradient of token p times g_ll[p]: one kernel over the CSR structure (was three gathers + six multiplies)
        o_ctrl = torch.empty_like(g_ctrl); o_inv = torch.empty_like(g_inv); o_pos = torch.empty_like(g_pos)
        o_aff = None if g_aff is None else torch.empty_like(g_aff)
        o_eps = torch.empty_like(g_eps); o_sig = torch.empty_like(g_sig)
        gin = _lib.Grads(None, None, g_ctrl.data_ptr(), g_inv.data_ptr(), g_pos.data_ptr(),
                         None if g_aff is None else g_aff.data_ptr(), None, g_eps.data_ptr(), g_sig.data_ptr())
        gout = _lib.Grads(None, None, o_ctrl.data_ptr(), o_inv.data_ptr(), o_pos.data_ptr(),
                          None if o_aff is None else o_aff.data_ptr(), None, o_eps.data_ptr(), o_sig.data_ptr())
        check(lib().bpl_scale_token_grads(b.P, b.K, b.S, b.tok_stroke_off.data_ptr(), b.stroke_sub_off.data_ptr(), gs.data_ptr(),
                                          ctypes.byref(gin), ctypes.byref(gout), _stream(b.device)))
        return o_ctrl, o_inv, o_pos, o_aff, o_eps, o_sig, None, None, None, None, None


class _RenderTokens(torch.autograd.Function):
    """pimg[p] = get_pimg(token p); backward takes a full cotangent on pimg and re-runs the fused kernel."""

    @staticmethod
    def forward(ctx, ctrl, invscale, position, affine, eps, sigma, batch, ps):
        dev = batch.device
        P = batch.P
        tens = [_c32(ctrl), _c32(invscale), _c32(position), None if affine is None else _c32(affine), _c32(eps),
                _c32(sigma)]
        pimg = torch.empty((P, IMSIZE, IMSIZE), dtype=torch.float32, device=dev)
        off = torch.empty(P, dtype=torch.uint8, device=dev)
        out = _lib.FwdOut(None, off.data_ptr(), pimg.data_ptr())
        b = batch.c_struct(*tens)
        t = _lib.Targets()
        check(lib().bpl_score_tokens(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                     _stream(dev)))
        ctx.batch = batch; ctx.ps = ps; ctx.has_aff = affine is not None
        ctx.save_for_backward(*[x for x in tens if x is not None])
        ctx.mark_non_differentiable(off)
        return pimg, off

    @staticmethod
    def backward(ctx, g_pimg, _g_off):
        saved = list(ctx.saved_tensors)
        ctrl, inv, pos = saved[:3]
        aff = saved[3] if ctx.has_aff else None
        eps, sig = saved[-2], saved[-1]
        batch = ctx.batch
        dev = batch.device
        P = batch.P
        gp = g_pimg.to(torch.float32).contiguous()
        g_ctrl = torch.empty_like(ctrl); g_inv = torch.empty_like(inv); g_pos = torch.empty_like(pos)
        g_aff = None if aff is None else torch.empty_like(aff)
        g_eps = torch.empty(P, dtype=torch.float32, device=dev); g_sig = torch.empty_like(g_eps)
        g = _lib.Grads(None, gp.data_ptr(), g_ctrl.data_ptr(), g_inv.data_ptr(), g_pos.data_ptr(),
                       None if g_aff is None else g_aff.data_ptr(), None, g_eps.data_ptr(), g_sig.data_ptr())
        b = batch.c_struct(ctrl, inv, pos, aff, eps, sig, cost_order=True)
        t = _lib.Targets()
        out = _lib.FwdOut(None, None, None)
        check(lib().bpl_score_tokens_grad(ctypes.byref(ctx.ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                          ctypes.byref(g), _stream(dev)))
        return g_ctrl, g_inv, g_pos, g_aff, g_eps, g_sig, None, None


def tokens_as_strokes(batch):
    """A TokenBatch as raw motor-space trajectories, what CharacterImageDist.get_pimg hands to render_image
    (model/image_dist.py:34-40): vanilla_to_motor (objects/part.py:290-328) by its kernel, then apply_warp
    (util/affine.py:29-55: com_char = mean of all the character's points, util/stroke.py:119-136) as a few elementwise
    tensor operations on the device.  Only the non-default-imsize path goes this way (the fused 105 x 105 kernels evaluate
    the splines themselves); differentiable in every leaf of the batch."""
    motor, _ = vanilla_to_motor_batch(batch.ctrl, batch.invscale, batch.position, batch.stroke_sub_off, batch.neval)
    S, n = motor.shape[0], motor.shape[1]
    sub_tok = batch.sub_token.long()
    if batch.affine is not None:
        A = batch.affine.to(torch.float32)
        cnt = torch.zeros(batch.P, dtype=torch.float32, device=batch.device).index_add_(
            0, sub_tok, torch.full((S,), float(n), device=batch.device))
        com = torch.zeros(batch.P, 2, dtype=torch.float32, device=batch.device).index_add(0, sub_tok, motor.sum(1))
        com = com / cnt.clamp(min=1.0).unsqueeze(1)
        B = torch.stack([A[:, 0], A[:, 1], A[:, 2] - (A[:, 0] - 1) * com[:, 0], A[:, 3] - (A[:, 1] - 1) * com[:, 1]], 1)
        Bs = B[sub_tok]
        motor = motor * Bs[:, None, :2] + Bs[:, None, 2:]
    tso = batch.stroke_sub_off.long()[batch.tok_stroke_off.long()]              # token -> sub-strokes
    spo = torch.arange(S + 1, device=batch.device, dtype=torch.int32) * n
    return StrokeBatch(tso.to(torch.int32), spo, motor.reshape(-1, 2), batch.eps, batch.sigma)


def score_tokens(batch, images, img_idx=None, ps=None):
    """Render and score a TokenBatch.  images: PackedImages; img_idx: int32 [P] or None (image 0).
    Returns (ll [P] float32, ink_off_page [P] bool).  Differentiable w.r.t. batch.ctrl / invscale /
    position / affine / eps / sigma.  (A non-default Parameters.imsize: images from binary_images.)"""
    if not isinstance(ps, _lib.Params):
        ps = make_params(ps)
    if image_size(ps) != (IMSIZE, IMSIZE):
        _, ll, off = render_strokes(tokens_as_strokes(batch), images, img_idx, want_pimg=False, ps=ps)
        return ll, off
    if img_idx is not None:
        img_idx = _i32(img_idx, batch.device)
        _check_img_idx(img_idx, images)
    ll, off = _ScoreTokens.apply(batch.ctrl, batch.invscale, batch.position, batch.affine, batch.eps, batch.sigma,
                                 batch, images, img_idx, ps, torch.is_grad_enabled())
    return ll, off.bool()


def _check_img_idx(img_idx, images):
    """An index outside [0, T) would read beyond the packed images: one synchronising reduction, once per index tensor."""
    if not img_idx.numel() or images is None or getattr(img_idx, "_bpl_checked_n", None) == images.n:
        return
    if torch.cuda.is_current_stream_capturing():      # a graph capture must not synchronise; checked on the eager warm-up
        return
    lo, hi = int(img_idx.min()), int(img_idx.max())
    if lo < 0 or hi >= images.n:
        raise IndexError(f"img_idx must lie in [0, {images.n}); got [{lo}, {hi}]")
    img_idx._bpl_checked_n = images.n                 # the same index tensor is not reduced again (fit loops)


class LseState:
    """Device state of a particle sweep's log-sum-exp partial on one GPU: `partial` = (max, sum exp(ll - max)) and
    the scratch the fused kernel needs (bpl_score_tokens_lse)."""

    def __init__(self, device):
        self.partial = torch.tensor([-float("inf"), 0.0], dtype=torch.float32, device=device)
        self.scratch = torch.empty(2 * _lib.LSE_SCRATCH_CTAS, dtype=torch.float32, device=device)


def score_tokens_lse(batch, images, state, accumulate=False, want_ll=True, ps=None):
    """score_tokens against image 0 plus the launch's log-sum-exp partial, in ONE kernel (forward only, no autograd):
    returns (ll [P] or None, ink_off_page [P] uint8); state.partial holds (max, sum exp(ll - max)) afterwards, merged
    with its previous value when `accumulate`."""
    if not isinstance(ps, _lib.Params):
        ps = make_params(ps)
    dev = batch.device
    P = batch.P
    tens = [_c32(batch.ctrl), _c32(batch.invscale), _c32(batch.position),
            None if batch.affine is None else _c32(batch.affine), _c32(batch.eps), _c32(batch.sigma)]
    ll = torch.empty(P, dtype=torch.float32, device=dev) if want_ll else None
    off = torch.empty(P, dtype=torch.uint8, device=dev)
    b = batch.c_struct(*tens, cost_order=True)
    t = _targets(images, None)
    out = _lib.FwdOut(None if ll is None else ll.data_ptr(), off.data_ptr(), None)
    check(lib().bpl_score_tokens_lse(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                     _ptr(state.partial), _ptr(state.scratch), int(bool(accumulate)), _stream(dev)))
    return ll, off


def score_tokens_all_images(batch, images, ps=None):
    """Every token of the batch against every image of `images` (PackedImages): ll [P, T], ink_off_page [P].
    Each token is rendered once; forward only (no autograd)."""
    if not isinstance(ps, _lib.Params):
        ps = make_params(ps)
    dev = batch.device
    P, Tn = batch.P, images.n
    ll = torch.empty((P, Tn), dtype=torch.float32, device=dev)
    off = torch.empty(P, dtype=torch.uint8, device=dev)
    tens = [_c32(batch.ctrl), _c32(batch.invscale), _c32(batch.position),
            None if batch.affine is None else _c32(batch.affine), _c32(batch.eps), _c32(batch.sigma)]
    b = batch.c_struct(*tens)
    t = _lib.Targets(images.bits.data_ptr(), None, Tn, 1)
    out = _lib.FwdOut(ll.data_ptr(), off.data_ptr(), None)
    check(lib().bpl_score_tokens(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out), _stream(dev)))
    return ll, off.bool()


def render_tokens(batch, ps=None):
    """Probability maps of a TokenBatch: (pimg [P,H,W], ink_off_page [P] bool); H x W = Parameters.imsize."""
    if not isinstance(ps, _lib.Params):
        ps = make_params(ps)
    if image_size(ps) != (IMSIZE, IMSIZE):
        pimg, _, off = render_strokes(tokens_as_strokes(batch), want_pimg=True, ps=ps)
        return pimg, off
    pimg, off = _RenderTokens.apply(batch.ctrl, batch.invscale, batch.position, batch.affine, batch.eps, batch.sigma,
                                    batch, ps)
    return pimg, off.bool()


def token_points(batch, ps=None):
    """Per-point integer work of a TokenBatch via the fused kernels' own device code: dict with
    motor [S,neval,2], mask_out [S,neval] bool, floor/ceil [S,neval,2] int32 (row, col), n_in [S] int32."""
    if not isinstance(ps, _lib.Params):
        ps = make_params(ps)
    dev = batch.device
    S, n = batch.S, batch.neval
    motor = torch.empty((S, n, 2), dtype=torch.float32, device=dev)
    mask = torch.empty((S, n), dtype=torch.uint8, device=dev)
    fl = torch.empty((S, n, 2), dtype=torch.int32, device=dev)
    ce = torch.empty((S, n, 2), dtype=torch.int32, device=dev)
    n_in = torch.empty(S, dtype=torch.int32, device=dev)
    tens = [_c32(batch.ctrl), _c32(batch.invscale), _c32(batch.position),
            None if batch.affine is None else _c32(batch.affine), _c32(batch.eps), _c32(batch.sigma)]
    b = batch.c_struct(*tens)
    check(lib().bpl_token_points(ctypes.byref(ps), ctypes.byref(b), _ptr(motor), _ptr(mask), _ptr(fl), _ptr(ce),
                                 _ptr(n_in), _stream(dev)))
    return dict(motor=motor, mask_out=mask.bool(), floor=fl, ceil=ce, n_in=n_in)


# ---------------------------------------------------------------------------------------------
# raw trajectories (rendering.render_image)
# ---------------------------------------------------------------------------------------------
class StrokeBatch:
    """P characters given as ragged motor-space trajectories: tok_sub_off [P+1], sub_pt_off [S+1], pts [N,2]."""

    def __init__(self, tok_sub_off, sub_pt_off, pts, eps, sigma):
        _need_cuda(pts, "pts")
        dev = pts.device
        self.tok_sub_off = _i32(tok_sub_off, dev)
        self.sub_pt_off = _i32(sub_pt_off, dev)
        self.pts = pts
        self.eps = eps
        self.sigma = sigma
        self.P = self.tok_sub_off.numel() - 1
        self.device = dev

    def c_struct(self, pts, eps, sigma):
        b = _lib.StrokeBatchC()
        b.n_tokens = self.P
        b.tok_sub_off = self.tok_sub_off.data_ptr(); b.sub_pt_off = self.sub_pt_off.data_ptr()
        b.pts = pts.data_ptr(); b.eps = eps.data_ptr(); b.sigma = sigma.data_ptr()
        return b


class _RenderStrokes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, eps, sigma, batch, images, img_idx, want_pimg, ps):
        dev = batch.device
        P = batch.P
        pts_c, eps_c, sig_c = _c32(pts), _c32(eps), _c32(sigma)
        pimg = torch.empty((P, IMSIZE, IMSIZE), dtype=torch.float32, device=dev) if want_pimg else None
        ll = torch.empty(P, dtype=torch.float32, device=dev) if images is not None else None
        off = torch.empty(P, dtype=torch.uint8, device=dev)
        out = _lib.FwdOut(None if ll is None else ll.data_ptr(), off.data_ptr(),
                          None if pimg is None else pimg.data_ptr())
        b = batch.c_struct(pts_c, eps_c, sig_c)
        t = _targets(images, img_idx)
        check(lib().bpl_render_strokes(ctypes.byref(ps), ctypes.byref(b), ctypes.byref(t), ctypes.byref(out),
                                       _stream(dev)))
        ctx.batch = batch; ctx.ps = ps; ctx.images = images; ctx.img_idx = img_idx
        ctx.save_for_backward(pts_c, eps_c, sig_c)
        ctx.mark_non_differentiable(off)
        if pimg is None:
            pimg = torch.empty(0, device=dev)
        if ll is None:
            ll = torch.zeros(P, dtype=torch.float32, device=dev)
        return pimg, ll, off

    @staticmethod
    def backward(ctx, g_pimg, g_ll, _g_off):
        pts, eps, sig = ctx.saved_tensors
        batch = ctx.batch
        dev = batch.device
        P = batch.P

        def run(gp, gl, images):
            g_pts = torch.zeros_like(pts)
            g_eps = torch.empty(P, dtype=torch.float32, device=dev); g_sig = torch.empty_like(g_eps)
            g = _lib.Grads(None if gl is None else gl.data_ptr(), None if gp is None else gp.data_ptr(), None, None,
                           None, None, g_pts.data_ptr(), g_eps.data_ptr(), g_sig.data_ptr())
            b = batch.c_struct(pts, eps, sig)
            t = _targets(images, ctx.img_idx)
            out = _lib.FwdOut(None, None, None)
            check(lib().bpl_render_strokes_grad(ctypes.byref(ctx.ps), ctypes.byref(b), ctypes.byref(t),
                                                ctypes.byref(out), ctypes.byref(g), _stream(dev)))
            return g_pts, g_eps, g_sig

        tot = None
        if g_pimg is not None and g_pimg.numel() > 0:
            tot = run(g_pimg.to(torch.float32).contiguous(), None, None)
        if ctx.images is not None and g_ll is not None:
            r = run(None, g_ll.to(torch.float32).contiguous(), ctx.images)
            tot = r if tot is None else tuple(a + b for a, b in zip(tot, r))
        if tot is None:
            tot = (torch.zeros_like(pts), torch.zeros(P, device=dev), torch.zeros(P, device=dev))
        return tot[0], tot[1], tot[2], None, None, None, None, None


def image_size(ps):
    """(H, W) a Params stands for: (105, 105) unless make_params saw another Parameters.imsize."""
    return getattr(ps, "hw", (IMSIZE, IMSIZE))


class _RenderStrokesAny(torch.autograd.Function):
    """render_image (+ Bernoulli score) for any Parameters.imsize: bpl_render_strokes_any[_grad].  images: float32
    [T,H,W] of exact 0 / 1 values or None."""

    @staticmethod
    def _geom(ps, images, img_idx, P, grad, dev):
        H, W = image_size(ps)
        n = int(lib().bpl_render_any_workspace(H, W, P, int(grad)))
        if n < 0:
            raise _lib.BplError("bpl_render_any_workspace: bad argument")
        ws = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        g = _lib.AnyGeomC(H, W, None if images is None else images.data_ptr(),
                          None if img_idx is None else img_idx.data_ptr(), ws.data_ptr(), n)
        return g, ws

    @staticmethod
    def forward(ctx, pts, eps, sigma, batch, images, img_idx, want_pimg, ps):
        dev = batch.device
        P = batch.P
        H, W = image_size(ps)
        pts_c, eps_c, sig_c = _c32(pts), _c32(eps), _c32(sigma)
        pimg = torch.empty((P, H, W), dtype=torch.float32, device=dev) if want_pimg else None
        ll = torch.empty(P, dtype=torch.float32, device=dev) if images is not None else None
        off = torch.empty(P, dtype=torch.uint8, device=dev)
        out = _lib.FwdOut(None if ll is None else ll.data_ptr(), off.data_ptr(), None if pimg is None else pimg.data_ptr())
        b = batch.c_struct(pts_c, eps_c, sig_c)
        g, ws = _RenderStrokesAny._geom(ps, images, img_idx, P, False, dev)
        check(lib().bpl_render_strokes_any(ctypes.byref(ps), ctypes.byref(g), ctypes.byref(b), ctypes.byref(out),
                                           _stream(dev)))
        ctx.batch = batch; ctx.ps = ps; ctx.images = images; ctx.img_idx = img_idx
        ctx.save_for_backward(pts_c, eps_c, sig_c)
        ctx.mark_non_differentiable(off)
        if pimg is None:
            pimg = torch.empty(0, device=dev)
        if ll is None:
            ll = torch.zeros(P, dtype=torch.float32, device=dev)
        return pimg, ll, off

    @staticmethod
    def backward(ctx, g_pimg, g_ll, _g_off):
        pts, eps, sig = ctx.saved_tensors
        batch = ctx.batch
        dev = batch.device
        P = batch.P

        def run(gp, gl, images):
            g_pts = torch.zeros_like(pts)
            g_eps = torch.empty(P, dtype=torch.float32, device=dev); g_sig = torch.empty_like(g_eps)
            gr = _lib.Grads(None if gl is None else gl.data_ptr(), None if gp is None else gp.data_ptr(), None, None,
                            None, None, g_pts.data_ptr(), g_eps.data_ptr(), g_sig.data_ptr())
            b = batch.c_struct(pts, eps, sig)
            geo, ws = _RenderStrokesAny._geom(ctx.ps, images, ctx.img_idx, P, True, dev)
            out = _lib.FwdOut(None, None, None)
            check(lib().bpl_render_strokes_any_grad(ctypes.byref(ctx.ps), ctypes.byref(geo), ctypes.byref(b),
                                                    ctypes.byref(out), ctypes.byref(gr), _stream(dev)))
            return g_pts, g_eps, g_sig

        tot = None
        if g_pimg is not None and g_pimg.numel() > 0:
            tot = run(g_pimg.to(torch.float32).contiguous(), None, None)
        if ctx.images is not None and g_ll is not None:
            r = run(None, g_ll.to(torch.float32).contiguous(), ctx.images)
            tot = r if tot is None else tuple(a + b for a, b in zip(tot, r))
        if tot is None:
            tot = (torch.zeros_like(pts), torch.zeros(P, device=dev), torch.zeros(P, device=dev))
        return tot[0], tot[1], tot[2], None, None, None, None, None


def binary_images(images, ps, device):
    """Target images for a non-default imsize: float32 [T,H,W] on `device`, checked to be exactly 0 / 1 the way torch's
    Bernoulli.log_prob checks its argument (model/image_dist.py:76-77).  (105 x 105 targets are bit-packed: pack_images.)"""
    H, W = image_size(ps)
    t = torch.as_tensor(images).to(device, torch.float32)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if tuple(t.shape[1:]) != (H, W):
        raise ValueError(f"images must be [T,{H},{W}]; got {tuple(t.shape)}")
    if not bool(((t == 0) | (t == 1)).all()):
        raise ValueError("Expected value argument to be within the support (Boolean()) of the "
                         "distribution Bernoulli(), but found invalid values")
    return t.contiguous()


def render_strokes(batch, images=None, img_idx=None, want_pimg=True, ps=None):
    """render_image over a StrokeBatch: returns (pimg [P,H,W] or None, ll [P] or None, off_page [P] bool).  H x W is
    Parameters.imsize: 105 x 105 runs the fused kernels (images: PackedImages), any other size the run-time-geometry
    kernels (images: float32 [T,H,W] from binary_images)."""
    if not isinstance(ps, _lib.Params):
        ps = make_params(ps)
    if img_idx is not None:
        img_idx = _i32(img_idx, batch.device)
    if image_size(ps) != (IMSIZE, IMSIZE):
        if images is not None and not torch.is_tensor(images):
            raise TypeError("a non-default imsize takes its target images from ops.binary_images, not pack_images")
        pimg, ll, off = _RenderStrokesAny.apply(batch.pts, batch.eps, batch.sigma, batch, images, img_idx, want_pimg, ps)
        return (pimg if want_pimg else None), (ll if images is not None else None), off.bool()
    pimg, ll, off = _RenderStrokes.apply(batch.pts, batch.eps, batch.sigma, batch, images, img_idx, want_pimg, ps)
    return (pimg if want_pimg else None), (ll if images is not None else None), off.bool()


# ---------------------------------------------------------------------------------------------
# splines
# ---------------------------------------------------------------------------------------------
class _SplineEval(torch.autograd.Function):
    """X[b] = C @ Y[b] (pybpl/splines.py:136); gradients to Y and to C (C carries d/ds when s needs grad)."""

    @staticmethod
    def forward(ctx, C, Y):
        dev = Y.device
        Cc, Yc = _c32(C), _c32(Y)
        nspl, nland = Yc.shape[0], Yc.shape[1]
        neval = Cc.shape[0]
        X = torch.empty((nspl, neval, 2), dtype=torch.float32, device=dev)
        check(lib().bpl_spline_eval(_ptr(Cc), neval, nland, nspl, _ptr(Yc), _ptr(X), _stream(dev)))
        ctx.save_for_backward(Cc, Yc)
        return X

    @staticmethod
    def backward(ctx, gX):
        Cc, Yc = ctx.saved_tensors
        dev = Yc.device
        nspl, nland = Yc.shape[0], Yc.shape[1]
        neval = Cc.shape[0]
        gXc = gX.to(torch.float32).contiguous()
        gY = None
        if ctx.needs_input_grad[1]:
            gY = torch.empty_like(Yc)
            check(lib().bpl_spline_eval_bwd(_ptr(Cc), neval, nland, nspl, _ptr(gXc), _ptr(gY), _stream(dev)))
        gC = None
        if ctx.needs_input_grad[0]:
            gC = torch.einsum("bjc,bkc->jk", gXc, Yc)     # only when the time points s carry a gradient
        return gC, gY


def spline_eval(C, Y):
    """Y: [nspl, nland, 2] CUDA; C: [neval, nland] CUDA -> X [nspl, neval, 2]."""
    _need_cuda(Y, "Y")
    return _SplineEval.apply(C, Y)


def lstsq_operator(C):
    """P = R^-1 Q^T of a coefficient matrix C [neval, nland] (host, double inside); returned on C's device."""
    Ch = C.detach().to("cpu", torch.float32).contiguous().numpy()
    m, n = Ch.shape
    P = np.empty((n, m), np.float32)
    check(lib().bpl_lstsq_operator_host(m, n, Ch.ctypes.data_as(ctypes.c_void_p), P.ctypes.data_as(ctypes.c_void_p)))
    return torch.from_numpy(P).to(C.device)


def spline_fit(C, X, P=None):
    """Least-squares control points of trajectories X [nspl, neval, 2] for the coefficient matrix C [neval, nland]
    (fit_bspline_to_traj, pybpl/splines.py:141-171): returns (Y [nspl, nland, 2], residuals [nspl, 2]).
    Differentiable w.r.t. X."""
    _need_cuda(X, "X")
    if P is None:
        P = lstsq_operator(C)
    Y = _SplineEval.apply(P, X)
    Xc, Yc, Cc = _c32(X), _c32(Y), _c32(C)
    res = torch.empty((Xc.shape[0], 2), dtype=torch.float32, device=X.device)
    check(lib().bpl_spline_residual(_ptr(Cc), Cc.shape[0], Cc.shape[1], Xc.shape[0], _ptr(Xc), _ptr(Yc), _ptr(res),
                                    _stream(X.device)))
    return Y, res


def adaptive_neval(Y, grain=1.5, min_neval=10, max_neval=200):
    """Integer neval per spline (pybpl/splines.py:129-133) and the trajectory lengths."""
    _need_cuda(Y, "Y")
    Yc = _c32(Y)
    nspl, nland = Yc.shape[0], Yc.shape[1]
    C = basis_table(nland, max_neval, Y.device)
    nev = torch.empty(nspl, dtype=torch.int32, device=Y.device)
    dist = torch.empty(nspl, dtype=torch.float32, device=Y.device)
    check(lib().bpl_spline_adaptive_neval(_ptr(C), max_neval, nland, nspl, _ptr(Yc), float(grain), int(min_neval),
                                          _ptr(nev), _ptr(dist), _stream(Y.device)))
    return nev, dist


class _VanillaToMotor(torch.autograd.Function):
    """Batched vanilla_to_motor (pybpl/objects/part.py:290-328): ctrl [S,ncpt,2], invscale [S], position [K,2],
    stroke_sub_off [K+1] -> motor [S,neval,2], motor_spline [S,ncpt,2]."""

    @staticmethod
    def forward(ctx, ctrl, invscale, position, stroke_sub_off, neval):
        dev = ctrl.device
        c, i, p = _c32(ctrl), _c32(invscale), _c32(position)
        S, ncpt = c.shape[0], c.shape[1]
        K = p.shape[0]
        C = basis_table(ncpt, neval, dev)
        motor = torch.empty((S, neval, 2), dtype=torch.float32, device=dev)
        ms = torch.empty((S, ncpt, 2), dtype=torch.float32, device=dev)
        check(lib().bpl_vanilla_to_motor(_ptr(C), neval, ncpt, K, _ptr(stroke_sub_off), _ptr(c), _ptr(i), _ptr(p),
                                         _ptr(motor), _ptr(ms), _stream(dev)))
        ctx.save_for_backward(c, i, C, stroke_sub_off)
        ctx.neval = neval
        return motor, ms

    @staticmethod
    def backward(ctx, g_motor, g_ms):
        # motor_i = C @ Y_i - off_i, ms_i = Y_i - off_i, off_i = (C[0] @ Y_i) - prev_i, prev_{i+1} = motor_i[-1];
        # a short chain of tiny tensor ops on the device -- this standalone operator is not on the hot path
        # (the fused kernels differentiate the chain themselves).
        c, inv, C, sso = ctx.saved_tensors
        dev = c.device
        S, ncpt = c.shape[0], c.shape[1]
        neval = ctx.neval
        gm = torch.zeros((S, neval, 2), device=dev) if g_motor is None else g_motor.to(torch.float32).clone()
        gs = torch.zeros((S, ncpt, 2), device=dev) if g_ms is None else g_ms.to(torch.float32)
        off = sso.tolist()
        K = len(off) - 1
        g_off_tot = -(gm.sum(1) + gs.sum(1))                     # d/d off_i from the direct uses
        g_prev = torch.zeros((S, 2), device=dev)
        g_pos = torch.zeros((K, 2), device=dev)
        for k in range(K):
            carry = torch.zeros(2, device=dev)
            for s in range(off[k + 1] - 1, off[k] - 1, -1):
                gm[s, -1] += carry
                g_off = g_off_tot[s] - carry                      # the carry also flows through -off_s
                g_prev[s] = -g_off
                gm[s, 0] += g_off                                # off_s = traj_s[0] - prev_s
                carry = g_prev[s]
            g_pos[k] = carry
        gY = torch.empty((S, ncpt, 2), dtype=torch.float32, device=dev)
        check(lib().bpl_spline_eval_bwd(_ptr(C), neval, ncpt, S, _ptr(gm.contiguous()), _ptr(gY), _stream(dev)))
        gY = gY + gs
        return gY * inv.view(-1, 1, 1), (gY * c).sum((1, 2)), g_pos, None, None


def vanilla_to_motor_batch(ctrl, invscale, position, stroke_sub_off, neval=200):
    _need_cuda(ctrl, "ctrl")
    return _VanillaToMotor.apply(ctrl, invscale, position, _i32(stroke_sub_off, ctrl.device), int(neval))


# ---------------------------------------------------------------------------------------------
# particle-sweep reduction
# ---------------------------------------------------------------------------------------------

# ---------------------------------------------------------------------------------------------
# token prior: log P(token | type)  (CharacterTokenDist.score_token, pybpl/model/token_dist.py:264-287)
# ---------------------------------------------------------------------------------------------
REL_CODES = {"unihist": 0, "start": 1, "end": 2, "mid": 3}      # pybpl/objects/relation.py:11


def token_prior_params(library=None, ps=None):
    """The constants CharacterTokenDist reads from a Library and Parameters (token_dist.py:90-107, 320-325, 437-441).
    `lib` is anything with .tokenvar / .rel mappings (the reference's Library); None gives the shipped values."""
    c = _lib.TokenPriorParams()
    lib = library
    if lib is None:
        c.sigma_shape, c.sigma_invscale, c.sigma_attach = 3.3593, 0.0277, 0.1499
        c.sigma_x, c.sigma_y = 1.3536, 1.2056
    else:
        c.sigma_shape = float(lib.tokenvar["sigma_shape"]); c.sigma_invscale = float(lib.tokenvar["sigma_invscale"])
        c.sigma_attach = float(lib.tokenvar["sigma_attach"])
        c.sigma_x = float(lib.rel["sigma_x"]); c.sigma_y = float(lib.rel["sigma_y"])
    c.min_blur_sigma = float(getattr(ps, "min_blur_sigma", 0.5))
    c.max_blur_sigma = float(getattr(ps, "max_blur_sigma", 16.0))
    return c


class TokenTypes:
    """Type-level parameters and relation tokens aligned with a TokenBatch's sub-strokes [S] and strokes [K]:
    ctrl_type [S,5,2], invscale_type [S], rel_cat [K] (REL_CODES), attach_ix [K], attach_subix [K], gpos [K,2],
    eval_spot_type [K], eval_spot [K] (the relation TOKEN's spot).  Entries a category does not use are ignored."""

    def __init__(self, ctrl_type, invscale_type, rel_cat, attach_ix, attach_subix, gpos, eval_spot_type, eval_spot):
        dev = ctrl_type.device
        _need_cuda(ctrl_type, "ctrl_type")
        self.ctrl_type, self.invscale_type = ctrl_type, invscale_type
        self.rel_cat, self.attach_ix, self.attach_subix = _i32(rel_cat, dev), _i32(attach_ix, dev), _i32(attach_subix, dev)
        self.gpos, self.eval_spot_type, self.eval_spot = gpos, eval_spot_type, eval_spot
        self.device = dev


class _ScoreTokenPrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ctrl, invscale, position, eval_spot, ctrl_type, invscale_type, gpos, eval_spot_type, batch, types, c,
                grad_on=True):
        dev = batch.device
        args = [_c32(t) for t in (ctrl, invscale, position, eval_spot, ctrl_type, invscale_type, gpos, eval_spot_type)]
        ctrl_c, inv_c, pos_c, es_c, ctrl_ty, inv_ty, gpos_c, es_ty = args
        b = _lib.TokenBatchC()
        b.n_tokens = batch.P; b.ncpt = NCPT; b.neval = batch.neval
        b.tok_stroke_off = batch.tok_stroke_off.data_ptr(); b.stroke_sub_off = batch.stroke_sub_off.data_ptr()
        b.ctrl = ctrl_c.data_ptr(); b.invscale = inv_c.data_ptr(); b.position = pos_c.data_ptr()
        sig, eps_c = _c32(batch.sigma), _c32(batch.eps)          # bound to locals: the pointers must outlive the launch
        b.sigma = sig.data_ptr(); b.eps = eps_c.data_ptr()
        b.basis = basis_table(NCPT, batch.neval, dev).data_ptr()
        ty = _lib.TokenTypeC(ctrl_ty.data_ptr(), inv_ty.data_ptr(), types.rel_cat.data_ptr(), types.attach_ix.data_ptr(),
                             types.attach_subix.data_ptr(), gpos_c.data_ptr(), es_ty.data_ptr(), es_c.data_ptr())
        ll = torch.empty(batch.P, dtype=torch.float32, device=dev)
        need = grad_on and any(ctx.needs_input_grad[:8])      # grad_on: the caller's torch.is_grad_enabled()
        if need:
            gs = [torch.empty_like(t) for t in args]
            g = _lib.TokenPriorGrads(*[t.data_ptr() for t in gs])
            check(lib().bpl_score_token_prior(ctypes.byref(c), ctypes.byref(b), ctypes.byref(ty), ll.data_ptr(),
                                              ctypes.byref(g), _stream(dev)))
            ctx.batch = batch
            ctx.save_for_backward(*gs)
        else:
            check(lib().bpl_score_token_prior(ctypes.byref(c), ctypes.byref(b), ctypes.byref(ty), ll.data_ptr(), None,
                                              _stream(dev)))
        return ll

    @staticmethod
    def backward(ctx, g_ll):
        g_ctrl, g_inv, g_pos, g_es, g_ctrl_ty, g_inv_ty, g_gpos, g_es_ty = ctx.saved_tensors
        b = ctx.batch
        gs = g_ll.to(torch.float32)
        st, kt = gs[b.sub_token], gs[b.stroke_token]
        return (g_ctrl * st.view(-1, 1, 1), g_inv * st, g_pos * kt.view(-1, 1), g_es * kt, g_ctrl_ty * st.view(-1, 1, 1),
                g_inv_ty * st, g_gpos * kt.view(-1, 1), g_es_ty * kt, None, None, None, None)


def score_token_prior(batch, types, library=None, ps=None, c=None):
    """ll[p] = log P(token p | its type) = CharacterTokenDist.score_token (pybpl/model/token_dist.py:264-287), batched,
    differentiable in the token's ctrl / invscale / position / eval_spot and the type's ctrl_type / invscale_type /
    gpos / eval_spot_type.  A token whose score is -inf gets gradients of the finite terms only."""
    if c is None:
        c = token_prior_params(library, ps)
    return _ScoreTokenPrior.apply(batch.ctrl, batch.invscale, batch.position, types.eval_spot, types.ctrl_type,
                                  types.invscale_type, types.gpos, types.eval_spot_type, batch, types, c, torch.is_grad_enabled())


def sample_tokens(tok_stroke_off, stroke_sub_off, types, seed, first_token=0, eps=1e-4, sigma=0.5, library=None, ps=None,
                  c=None, neval=200):
    """Draw one token per entry of the (expanded) type batch from P(token | type): the batched, on-device
    CharacterTokenDist.sample_token (pybpl/model/token_dist.py:226-262).  `types` holds the type-level arrays aligned with
    the CSR structure; its eval_spot is replaced by the sampled relation tokens' spots.  eps / sigma: the reference fixes
    them to Parameters.min_epsilon / min_blur_sigma (token_dist.py:253-260); scalars or [P] tensors.  Token p uses the
    Philox stream (seed, first_token + p).  Returns (TokenBatch, TokenTypes)."""
    dev = types.device
    if c is None:
        c = token_prior_params(library, ps)
    tso, sso = _i32(tok_stroke_off, dev), _i32(stroke_sub_off, dev)
    P, K, S = tso.numel() - 1, sso.numel() - 1, types.ctrl_type.shape[0]
    ctrl = torch.empty((S, NCPT, 2), dtype=torch.float32, device=dev)
    inv = torch.empty(S, dtype=torch.float32, device=dev)
    pos = torch.empty((K, 2), dtype=torch.float32, device=dev)
    es = torch.empty(K, dtype=torch.float32, device=dev)
    b = _lib.TokenBatchC()
    b.n_tokens = P; b.ncpt = NCPT; b.neval = int(neval)
    b.tok_stroke_off = tso.data_ptr(); b.stroke_sub_off = sso.data_ptr()
    b.basis = basis_table(NCPT, int(neval), dev).data_ptr()
    # converted copies are bound to locals so that they outlive the launch (a temporary's block could be handed to the
    # next allocation before the kernel reads it)
    ct_c, it_c, gp_c, est_c = _c32(types.ctrl_type), _c32(types.invscale_type), _c32(types.gpos), _c32(types.eval_spot_type)
    ty = _lib.TokenTypeC(ct_c.data_ptr(), it_c.data_ptr(), types.rel_cat.data_ptr(),
                         types.attach_ix.data_ptr(), types.attach_subix.data_ptr(), gp_c.data_ptr(),
                         est_c.data_ptr(), None)
    check(lib().bpl_sample_tokens(ctypes.byref(c), ctypes.byref(b), ctypes.byref(ty), int(seed), int(first_token),
                                   ctrl.data_ptr(), inv.data_ptr(), pos.data_ptr(), es.data_ptr(), _stream(dev)))
    full = lambda v: v.to(dev, torch.float32) if isinstance(v, torch.Tensor) else torch.full((P,), float(v), device=dev)
    batch = TokenBatch(tso, sso, ctrl, inv, pos, full(eps), full(sigma), neval=neval, check_limits=False)
    out_types = TokenTypes(types.ctrl_type, types.invscale_type, types.rel_cat, types.attach_ix, types.attach_subix,
                           types.gpos, types.eval_spot_type, es)
    return batch, out_types


class TypeLibrary:
    """The type prior's tables on one device, in the form bpl_sample_types reads them (include/pybpl_b200.h:bpl_type_lib).
    Default: pybpl_b200/data/type_lib.npz, generated from the reference's Library(use_hist=True) by
    tests/golden/make_golden.py:make_assets (pybpl/library/library.py:30-72)."""
    _KEYS = ("pkappa", "pmat_nsub", "cdf_start", "cdf_T", "shape_mu", "shape_L", "gamma_con", "gamma_rate", "rel_mixprob",
             "sp_cdf", "sp_xlab", "sp_ylab")

    def __init__(self, device, path=None):
        path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "type_lib.npz")
        d = np.load(path)
        self.device = torch.device(device)
        self.t = {k: torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float32)).to(self.device) for k in self._KEYS}
        c = _lib.TypeLibC()
        c.n_prim = d["cdf_start"].shape[0]; c.kmax = d["pkappa"].shape[0]; c.nsub_max = d["pmat_nsub"].shape[1]
        c.n_hist = d["sp_cdf"].shape[0]; c.hist_bins = d["sp_xlab"].shape[0] - 1
        for k in self._KEYS:
            setattr(c, k, self.t[k].data_ptr())
        self.c = c
        lt = _lib.TypeLogTablesC()
        for k in ("logp_kappa", "logp_nsub", "logp_start", "logp_T", "logp_mix", "sp_logp"):
            self.t[k] = torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float32)).to(self.device)
            setattr(lt, k, self.t[k].data_ptr())
        lt.sp_log_binarea = float(d["sp_log_binarea"])
        self.logtables = lt


def sample_types(library, n_types, seed, first_type=0):
    """Draw n_types character types from the prior on the device: the batched CharacterTypeDist.sample_type
    (pybpl/model/type_dist.py:56-96,187-207).  Two launches over the same counter-based streams (count, scan, write).
    Returns (tok_stroke_off [P+1], stroke_sub_off [K+1], TokenTypes with eval_spot = 0, subid [S])."""
    dev = library.device
    P = int(n_types)
    nk = torch.empty(P, dtype=torch.int32, device=dev)
    ns = torch.empty(P, dtype=torch.int32, device=dev)
    L = lib()
    check(L.bpl_sample_types(ctypes.byref(library.c), P, int(seed), int(first_type), None, None, nk.data_ptr(), ns.data_ptr(),
                             None, _stream(dev)))
    tso = torch.zeros(P + 1, dtype=torch.int32, device=dev); tso[1:] = torch.cumsum(nk, 0)
    tsub = torch.zeros(P + 1, dtype=torch.int32, device=dev); tsub[1:] = torch.cumsum(ns, 0)
    K, S = int(tso[-1]), int(tsub[-1])               # (one synchronisation: the output sizes)
    i32 = lambda n: torch.empty(n, dtype=torch.int32, device=dev)
    f32 = lambda *n: torch.empty(n, dtype=torch.float32, device=dev)
    stroke_nsub, subid, cat, aix, asub = i32(K), i32(S), i32(K), i32(K), i32(K)
    ctrl, inv, gpos, es = f32(S, NCPT, 2), f32(S), f32(K, 2), f32(K)
    out = _lib.TypeOutC(stroke_nsub.data_ptr(), subid.data_ptr(), ctrl.data_ptr(), inv.data_ptr(), cat.data_ptr(), aix.data_ptr(),
                        asub.data_ptr(), gpos.data_ptr(), es.data_ptr())
    check(L.bpl_sample_types(ctypes.byref(library.c), P, int(seed), int(first_type), tso.data_ptr(), tsub.data_ptr(), None, None,
                             ctypes.byref(out), _stream(dev)))
    sso = torch.zeros(K + 1, dtype=torch.int32, device=dev); sso[1:] = torch.cumsum(stroke_nsub, 0)
    types = TokenTypes(ctrl, inv, cat, aix, asub, gpos, es, torch.zeros(K, dtype=torch.float32, device=dev))
    return tso, sso, types, subid


class _ScoreTypePrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ctrl_type, invscale_type, library, tok_stroke_off, stroke_sub_off, subid, types, grad_on=True):
        dev = library.device
        ct, it = _c32(ctrl_type), _c32(invscale_type)
        P = tok_stroke_off.numel() - 1
        gp_c, est_c = _c32(types.gpos), _c32(types.eval_spot_type)      # locals: must outlive the launch
        ty = _lib.TokenTypeC(ct.data_ptr(), it.data_ptr(), types.rel_cat.data_ptr(), types.attach_ix.data_ptr(),
                             types.attach_subix.data_ptr(), gp_c.data_ptr(), est_c.data_ptr(), None)
        ll = torch.empty(P, dtype=torch.float32, device=dev)
        need = grad_on and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])      # grad_on: the caller's grad mode
        g_c = torch.empty_like(ct) if need else None
        g_i = torch.empty_like(it) if need else None
        check(lib().bpl_score_type_prior(ctypes.byref(library.c), ctypes.byref(library.logtables), P, tok_stroke_off.data_ptr(),
                                         stroke_sub_off.data_ptr(), subid.data_ptr(), ctypes.byref(ty), ll.data_ptr(),
                                         None if g_c is None else g_c.data_ptr(), None if g_i is None else g_i.data_ptr(),
                                         _stream(dev)))
        if need:
            tso = tok_stroke_off.long()
            sub_off = stroke_sub_off.long()[tso]
            ctx.sub_token = torch.repeat_interleave(torch.arange(P, device=dev), sub_off[1:] - sub_off[:-1], output_size=ct.shape[0])
            ctx.save_for_backward(g_c, g_i)
        return ll

    @staticmethod
    def backward(ctx, g_ll):
        g_c, g_i = ctx.saved_tensors
        st = g_ll.to(torch.float32)[ctx.sub_token]
        return g_c * st.view(-1, 1, 1), g_i * st, None, None, None, None, None, None


def score_type_prior(library, tok_stroke_off, stroke_sub_off, types, subid):
    """ll[p] = log P(type p) = CharacterTypeDist.score_type (pybpl/model/type_dist.py:98-125), batched; differentiable in
    types.ctrl_type and types.invscale_type (the leaves the reference's autograd reaches).  `subid` [S]: primitive ids."""
    dev = library.device
    return _ScoreTypePrior.apply(types.ctrl_type, types.invscale_type, library, _i32(tok_stroke_off, dev),
                                 _i32(stroke_sub_off, dev), _i32(subid, dev), types, torch.is_grad_enabled())


class FusedAdam:
    """torch.optim.Adam's update (no weight decay, no amsgrad) over a handful of small leaf tensors in ONE launch, each
    clamped to its bounds afterwards: the optimiser step of the differentiable fit (fit_image, pybpl/model/model.py:42-65;
    bounds as pybpl/parameters.py:54-63).  The step counter lives on the device, so a step captured in a CUDA graph
    advances on replay.  params: float32 contiguous CUDA leaves (at most 8); bounds: {index: (lo, hi)}, None = unbounded;
    maximize: ascend along the gradients (use with `objective.sum().backward()`)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, bounds=None, maximize=False):
        self.params = list(params)
        if not 0 < len(self.params) <= _lib.ADAM_MAX_TENSORS:
            raise ValueError("FusedAdam takes 1..8 tensors")
        for p in self.params:
            _need_cuda(p, "param")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("FusedAdam parameters must be contiguous float32")
        dev = self.params[0].device
        self.lr, self.betas, self.eps, self.maximize = float(lr), betas, float(eps), bool(maximize)
        self.bounds = dict(bounds or {})
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.state = torch.zeros(2, dtype=torch.float32, device=dev)

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def reset(self):
        self.state.zero_()
        for t in self.exp_avg + self.exp_avg_sq:
            t.zero_()

    @torch.no_grad()
    def step(self):
        a = _lib.AdamArgs()
        a.n_tensors = len(self.params); a.maximize = int(self.maximize)
        a.lr, a.beta1, a.beta2, a.eps = self.lr, float(self.betas[0]), float(self.betas[1]), self.eps
        a.state = self.state.data_ptr()
        grads = []
        for k, p in enumerate(self.params):
            g = p.grad
            if g is None:
                raise RuntimeError("FusedAdam.step: a parameter has no gradient")
            g = _c32(g); grads.append(g)
            lo, hi = self.bounds.get(k, (None, None))
            a.param[k] = p.data_ptr(); a.grad[k] = g.data_ptr()
            a.exp_avg[k] = self.exp_avg[k].data_ptr(); a.exp_avg_sq[k] = self.exp_avg_sq[k].data_ptr()
            a.n[k] = p.numel()
            a.lo[k] = -math.inf if lo is None else float(lo); a.hi[k] = math.inf if hi is None else float(hi)
        check(lib().bpl_adam_step(ctypes.byref(a), _stream(self.params[0].device)))


def logsumexp_partial(ll):
    """(max, sum exp(ll - max)) of a CUDA float32 vector as a 2-element tensor (one kernel)."""
    _need_cuda(ll, "ll")
    llc = _c32(ll)
    out = torch.empty(2, dtype=torch.float32, device=ll.device)
    check(lib().bpl_logsumexp_partial(_ptr(llc), llc.numel(), _ptr(out), _stream(ll.device)))
    return out


def logsumexp_combine(parts):
    """log sum_i s_i exp(m_i) of (max, sum-exp) pairs ([n,2] or flat [2n] CUDA float32) as a 1-element tensor (one kernel)."""
    _need_cuda(parts, "parts")
    pc = _c32(parts).reshape(-1)
    out = torch.empty(1, dtype=torch.float32, device=parts.device)
    check(lib().bpl_logsumexp_combine(_ptr(pc), pc.numel() // 2, _ptr(out), _stream(parts.device)))
    return out
