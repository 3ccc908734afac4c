"""Loaders for the golden fixtures under tests/golden (made by make_golden.py from the live reference)."""
import functools
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def unpack_image(packed):
    """np.packbits rows (105,14) uint8 -> (105,105) uint8 {0,1}."""
    return np.unpackbits(packed, axis=-1)[..., :105].astype(np.uint8)


def relinf(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


@functools.lru_cache(maxsize=None)
def _render_cases():
    import torch
    r = load("render")
    out = []
    for i, n in enumerate(r["names"]):
        k = f"c{i:02d}"
        lens = r[k + "_len"]
        pts = r[k + "_pts"]
        gen = torch.Generator().manual_seed(int(r[k + "_ct_seed"]))
        ct = torch.randn(105, 105, generator=gen).numpy()
        out.append(dict(name=str(n), pts=pts, lens=lens, strokes=np.split(pts, np.cumsum(lens)[:-1]),
                        eps=float(r[k + "_eps"]), sigma=float(r[k + "_sigma"]), hinton=bool(r[k + "_hinton"]),
                        pimg=r[k + "_pimg"], off=bool(r[k + "_off"]), image=unpack_image(r[k + "_image"]),
                        ll=float(r[k + "_ll"]), mask_out=r[k + "_mask_out"], floor=r[k + "_floor"],
                        ceil=r[k + "_ceil"], g_pts=r[k + "_g_pts"], g_eps=float(r[k + "_g_eps"]),
                        g_sigma=float(r[k + "_g_sigma"]), ct=ct, gct_pts=r[k + "_gct_pts"],
                        gct_eps=float(r[k + "_gct_eps"]), gct_sigma=float(r[k + "_gct_sigma"])))
    return out


def render_cases():
    return _render_cases()


@functools.lru_cache(maxsize=None)
def _token_cases():
    d = load("tokens")
    K = d["K"]; S = d["S"]
    toks = []
    k0 = 0; s0 = 0
    for i, name in enumerate(d["names"]):
        k1 = k0 + int(K[i]); s1 = s0 + int(S[i])
        t = dict(name=str(name), nsub=d["nsub"][k0:k1], ctrl=d["ctrl"][s0:s1], invscale=d["invscale"][s0:s1],
                 position=d["position"][k0:k1], eps=float(d["eps"][i]), sigma=float(d["sigma"][i]),
                 affine=d["affine"][i], has_affine=bool(d["has_affine"][i]), ll=float(d["ll"][i]),
                 off_page=bool(d["off_page"][i]), image=unpack_image(d["image"][i]),
                 g_ctrl=d["g_ctrl"][s0:s1], g_invscale=d["g_invscale"][s0:s1], g_position=d["g_position"][k0:k1],
                 g_sigma=float(d["g_sigma"][i]), g_eps=float(d["g_eps"][i]), g_affine=d["g_affine"][i],
                 n_inbounds=d["n_inbounds"][s0:s1], floor_sum=d["floor_sum"][s0:s1],
                 mask_out=np.unpackbits(d["mask_out"][s0:s1], axis=1)[:, :200].astype(bool))
        if i < int(d["n_full"]):
            t["motor"] = d[f"motor_{i:03d}"]
            t["pimg"] = d[f"pimg_{i:03d}"]
        toks.append(t)
        k0, s0 = k1, s1
    return d, toks


def token_cases():
    return _token_cases()


def pack_tokens(toks):
    """Packed ragged batch (the layout of include/pybpl_b200.h) from a list of token dicts."""
    tso = [0]; sso = [0]
    for t in toks:
        tso.append(tso[-1] + len(t["nsub"]))
        for n in t["nsub"]:
            sso.append(sso[-1] + int(n))
    return dict(tok_stroke_off=np.array(tso, np.int32), stroke_sub_off=np.array(sso, np.int32),
                ctrl=np.concatenate([t["ctrl"] for t in toks]).astype(np.float32),
                invscale=np.concatenate([t["invscale"] for t in toks]).astype(np.float32),
                position=np.concatenate([t["position"] for t in toks]).astype(np.float32),
                eps=np.array([t["eps"] for t in toks], np.float32),
                sigma=np.array([t["sigma"] for t in toks], np.float32),
                affine=np.stack([t["affine"] for t in toks]).astype(np.float32),
                has_affine=np.array([t["has_affine"] for t in toks]),
                images=np.stack([t["image"] for t in toks]), img_idx=np.arange(len(toks), dtype=np.int32))
