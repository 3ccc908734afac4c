"""Packed on-disk / wire format for parses and bit-packed image sets.

The reference keeps parses as Python object graphs (CharacterToken -> StrokeToken / RelationToken,
pybpl/objects/concept.py:217-241, part.py:188-249, relation.py:14-32) and images as float tensors; scoring a
million parses against an image set means traversing a million objects.  A parse FILE here is one flat buffer whose
body is, byte for byte, what the GPU wants: the CSR offsets and parameter arrays of `bpl_token_batch`, optionally
the `bpl_token_type` arrays and a set of binary 105x105 images already bit-packed (345 little-endian u32 words per
image, pixel i = bit i&31 of word i>>5, the layout `bpl_pack_images_*` produces).  Loading it is one read into
pinned memory and ONE host->device copy; the tensors handed to the kernels are views of that device buffer.

Layout (all little-endian; every section starts on a 16-byte boundary):
    0   8   magic  b"BPLPARS1"
    8   4   u32 header bytes (offset of the first section)
    12  4   u32 number of sections
    16  4x  i32 P (tokens), K (strokes), S (sub-strokes), T (images), ncpt, neval, reserved x2
    48      section table: n x { char name[16]; u32 dtype (0 f32, 1 i32, 2 u32); u32 ndim; u32 shape[3]; u64 offset; u64 nbytes }
    ...     sections
`include/pybpl_parsefile.h` states the same layout for non-Python readers.
"""
import struct

import numpy as np
import torch

from . import ops
from ._lib import IMG_WORDS, IMSIZE, MAX_STROKES, MAX_SUBSTROKES, NCPT

MAGIC = b"BPLPARS1"
_DT = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint32): 2}
_DT_INV = {0: np.float32, 1: np.int32, 2: np.uint32}
_ENTRY = struct.Struct("<16sII3IQQ")
_HEAD = struct.Struct("<8sII8i")
TOKEN_KEYS = ("tok_stroke_off", "stroke_sub_off", "ctrl", "invscale", "position", "eps", "sigma")
OPTIONAL_KEYS = ("affine",)
TYPE_KEYS = ("ctrl_type", "invscale_type", "rel_cat", "attach_ix", "attach_subix", "gpos", "eval_spot_type", "eval_spot")


def pack_bits(images):
    """[T,105,105] {0,1} array -> [T,345] u32 words (host-side twin of bpl_pack_images_u8)."""
    a = np.asarray(images)
    if a.ndim == 2:
        a = a[None]
    if a.shape[1:] != (IMSIZE, IMSIZE):
        raise ValueError("images must be 105x105")
    if not np.isin(a, (0, 1)).all():
        raise ValueError("Expected value argument to be within the support (Boolean()) of the distribution Bernoulli(), "
                         "but found invalid values")
    flat = np.zeros((a.shape[0], IMG_WORDS * 32), np.uint8)
    flat[:, :IMSIZE * IMSIZE] = a.reshape(a.shape[0], -1)
    return np.packbits(flat, axis=1, bitorder="little").view("<u4").reshape(a.shape[0], IMG_WORDS).copy()


def unpack_bits(bits):
    b = np.ascontiguousarray(np.asarray(bits, dtype="<u4").reshape(-1, IMG_WORDS))
    flat = np.unpackbits(b.view(np.uint8), axis=1, bitorder="little")
    return flat[:, :IMSIZE * IMSIZE].reshape(-1, IMSIZE, IMSIZE)


def to_bytes(w, types=None, images=None, image_bits=None):
    """Serialise a packed batch `w` (dict in the layout of workload.synth_tokens), optional type arrays (layout of
    workload.synth_types) and optional images ([T,105,105] {0,1}, or already packed [T,345] words)."""
    sections = []
    for k in TOKEN_KEYS:
        sections.append((k, np.ascontiguousarray(w[k])))
    for k in OPTIONAL_KEYS:
        if w.get(k) is not None:
            sections.append((k, np.ascontiguousarray(w[k], dtype=np.float32)))
    if types is not None:
        for k in TYPE_KEYS:
            sections.append((k, np.ascontiguousarray(types[k])))
    if image_bits is None and images is not None:
        image_bits = pack_bits(images)
    if image_bits is not None:
        sections.append(("image_bits", np.ascontiguousarray(image_bits, dtype=np.uint32)))
    P = len(w["tok_stroke_off"]) - 1
    K = len(w["stroke_sub_off"]) - 1
    S = int(np.asarray(w["ctrl"]).shape[0])
    T = 0 if image_bits is None else int(np.asarray(image_bits).shape[0])
    if np.asarray(w["ctrl"]).shape[1:] != (NCPT, 2):
        raise ValueError("ctrl must be [S,5,2]")
    head_bytes = (_HEAD.size + _ENTRY.size * len(sections) + 15) // 16 * 16
    off = head_bytes
    table, blobs = [], []
    for name, a in sections:
        if a.dtype not in _DT:
            a = a.astype(np.int32 if np.issubdtype(a.dtype, np.integer) else np.float32)
        shape = list(a.shape) + [0] * (3 - a.ndim)
        table.append(_ENTRY.pack(name.encode().ljust(16, b"\0"), _DT[a.dtype], a.ndim, *shape, off, a.nbytes))
        blobs.append((off, a.tobytes()))
        off = (off + a.nbytes + 15) // 16 * 16
    buf = bytearray(off)
    buf[:_HEAD.size] = _HEAD.pack(MAGIC, head_bytes, len(sections), P, K, S, T, NCPT, int(w.get("neval", 200)), 0, 0)
    buf[_HEAD.size:_HEAD.size + _ENTRY.size * len(table)] = b"".join(table)
    for o, b in blobs:
        buf[o:o + len(b)] = b
    return bytes(buf)


def save(path, w, types=None, images=None, image_bits=None):
    data = to_bytes(w, types, images, image_bits)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


class ParseFile:
    """A parse file in host memory (pinned when a CUDA device is present).  `.arrays()` gives numpy views,
    `.upload(device)` copies the whole buffer once and returns (TokenBatch, TokenTypes or None, PackedImages or None)
    whose tensors are views of the device copy."""

    def __init__(self, data, pin=None):
        n = len(data)
        if pin is None:
            pin = torch.cuda.is_available()
        host = torch.empty(n, dtype=torch.uint8)
        if pin:
            host = host.pin_memory()
        host.numpy()[:] = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
        self.host = host
        raw = host.numpy()
        magic, head_bytes, nsec, P, K, S, T, ncpt, neval, _, _ = _HEAD.unpack_from(raw, 0)
        if magic != MAGIC:
            raise ValueError("not a pybpl_b200 parse file")
        if ncpt != NCPT:
            raise ValueError("parse file has ncpt != 5")
        self.P, self.K, self.S, self.T, self.neval = P, K, S, T, neval
        self.table = {}
        for i in range(nsec):
            name, dt, ndim, s0, s1, s2, off, nb = _ENTRY.unpack_from(raw, _HEAD.size + i * _ENTRY.size)
            shape = (s0, s1, s2)[:ndim]
            if off % 16 or off + nb > n or int(np.prod(shape, dtype=np.int64)) * 4 != nb:
                raise ValueError("corrupt parse file section table")
            self.table[name.rstrip(b"\0").decode()] = (_DT_INV[dt], shape, off, nb)
        for k in TOKEN_KEYS:
            if k not in self.table:
                raise ValueError(f"parse file lacks section {k}")
        a = self.arrays()
        tso, sso = a["tok_stroke_off"].astype(np.int64), a["stroke_sub_off"].astype(np.int64)
        if len(tso) != P + 1 or len(sso) != K + 1 or (P and (tso[0] != 0 or tso[-1] != K)) or (K and (sso[0] != 0 or sso[-1] != S)) \
                or np.any(np.diff(tso) < 0) or np.any(np.diff(sso) < 0):
            raise ValueError("parse file offsets are inconsistent")
        if P and (np.diff(tso).max() > MAX_STROKES or np.diff(sso[tso]).max() > MAX_SUBSTROKES):
            raise ValueError("a token exceeds 32 strokes / 128 sub-strokes (BPL_MAX_*)")
        # every section against the header's counts: the kernels index these arrays by the offsets without bounds checks
        want = {"ctrl": (S, NCPT, 2), "invscale": (S,), "position": (K, 2), "eps": (P,), "sigma": (P,), "affine": (P, 4),
                "ctrl_type": (S, NCPT, 2), "invscale_type": (S,), "rel_cat": (K,), "attach_ix": (K,), "attach_subix": (K,),
                "gpos": (K, 2), "eval_spot_type": (K,), "eval_spot": (K,), "image_bits": (T, IMG_WORDS)}
        for k, shp in want.items():
            if k in self.table and tuple(self.table[k][1]) != shp:
                raise ValueError(f"parse file section {k} has shape {tuple(self.table[k][1])}, the header says {shp}")
        if "rel_cat" in self.table and K:
            rank = np.arange(K) - np.repeat(tso[:-1], np.diff(tso))          # stroke index inside its token
            cat, aix, asub = a["rel_cat"], a["attach_ix"], a["attach_subix"]
            if cat.min() < 0 or cat.max() > 3:
                raise ValueError("parse file: rel_cat outside 0..3")
            att = cat != 0
            if np.any(att & ((aix < 0) | (aix >= rank))):
                raise ValueError("parse file: attach_ix does not name an earlier stroke of the same token")
            mid = cat == 3
            if mid.any():
                tgt = np.repeat(tso[:-1], np.diff(tso))[mid] + aix[mid]
                if np.any((asub[mid] < 0) | (asub[mid] >= np.diff(sso)[tgt])):
                    raise ValueError("parse file: attach_subix outside the attached stroke's sub-strokes")
        if "image_bits" in self.table and T:
            used = IMSIZE * IMSIZE - 32 * (IMG_WORDS - 1)                      # 17 pixels in the last word, 15 padding bits
            if np.any(a["image_bits"].view(np.uint32)[:, -1] >> np.uint32(used)):
                raise ValueError("parse file: padding bits of an image are set (the ones count would be wrong)")

    @classmethod
    def load(cls, path, pin=None):
        return cls(np.fromfile(path, dtype=np.uint8), pin=pin)

    def arrays(self):
        raw = self.host.numpy()
        return {k: raw[off:off + nb].view(dt).reshape(shape) for k, (dt, shape, off, nb) in self.table.items()}

    def upload(self, device, non_blocking=True):
        dev = torch.device(device)
        d = torch.empty(self.host.numel(), dtype=torch.uint8, device=dev)
        d.copy_(self.host, non_blocking=non_blocking)
        tdt = {np.float32: torch.float32, np.int32: torch.int32, np.uint32: torch.int32}

        def view(k):
            dt, shape, off, nb = self.table[k]
            return d[off:off + nb].view(tdt[dt]).view(shape)
        batch = ops.TokenBatch(view("tok_stroke_off"), view("stroke_sub_off"), view("ctrl"), view("invscale"),
                               view("position"), view("eps"), view("sigma"),
                               affine=view("affine") if "affine" in self.table else None, neval=self.neval,
                               check_limits=False)
        types = None
        if "ctrl_type" in self.table:
            types = ops.TokenTypes(*[view(k) for k in TYPE_KEYS])
        images = None
        if "image_bits" in self.table:
            images = ops.PackedImages(view("image_bits"), torch.zeros((), dtype=torch.int32, device=dev))
        self.device_buffer = d
        return batch, types, images
