import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import cross_score
from oracle import Oracle
from pybpl_b200 import workload, ops
ll, _, w, images = cross_score.cross_score(n_tokens=40, n_images=5, seed=5)
ll = ll.cpu().numpy()
o = Oracle(np.float32)
C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
imgs = images.cpu().numpy()
rep = np.repeat(np.arange(40), 5)
wb = workload.permute(w, rep)
ref = o.token_batch(C, wb["tok_stroke_off"], wb["stroke_sub_off"], wb["ctrl"], wb["invscale"], wb["position"], wb["eps"], wb["sigma"], imgs, img_idx=np.tile(np.arange(5, dtype=np.int32), 40), want_pimg=True)
rel = (np.abs(ll.reshape(-1) - ref["ll"]) / np.abs(ref["ll"])).reshape(40, 5)
tso = w["tok_stroke_off"].astype(np.int64); sso = w["stroke_sub_off"].astype(np.int64)
S = np.diff(sso[tso])
bad = np.nonzero(rel.max(1) > 1e-5)[0]
print("bad tokens", bad, "S", S[bad], "rel", rel[bad].max(1))
print("ll", ll[bad[:3]], ref["ll"].reshape(40,5)[bad[:3]])
dev = torch.device("cuda:0")
pimg, _ = ops.render_tokens(workload.to_device(w, dev))
pimg = pimg.cpu().numpy()
pref = ref["pimg"].reshape(40, 5, 105, 105)[:, 0]
for t in bad[:5]:
    d = np.abs(pimg[t] - pref[t]); print("token", t, "S", S[t], "pimg maxdiff (single-token kernel)", d.max(), "at", np.unravel_index(d.argmax(), d.shape))
# pixel-class analysis for token 0 vs image 1
t, im = 0, 1
p = torch.from_numpy(pimg[t]); x = torch.from_numpy(imgs[im]).float()
lp_t = torch.distributions.Bernoulli(probs=p).log_prob(x)
print("torch ll on kernel pimg", lp_t.sum().item(), "kernel ll", ll[t, im], "oracle", ref["ll"].reshape(40,5)[t, im])
p64 = p.double().clamp(1.1920929e-7, 1-1.1920929e-7)
exact = torch.where(x > 0, p64.log(), (-p64).log1p())
bg = (p == 1e-4)
for name, m in [("bg x=0", bg & (x == 0)), ("bg x=1", bg & (x == 1)), ("nonbg x=0", ~bg & (x == 0)), ("nonbg x=1 p<.5", ~bg & (x == 1) & (p < .5)), ("nonbg x=1 p>=.5", ~bg & (x == 1) & (p >= .5))]:
    print(f"{name:18s} n={int(m.sum()):5d} torch-exact sum diff {float((lp_t.double()-exact)[m].sum()):+.6f}")
po = torch.from_numpy(pref[t].copy())
lp_o = torch.distributions.Bernoulli(probs=po).log_prob(x)
print("torch ll on ORACLE pimg", lp_o.sum().item(), " (oracle's own ll", ref["ll"].reshape(40,5)[t, im], ")")
d = (po - p).abs()
print("pimg diff at x=1 pixels max", float(d[x > 0].max()), "relative max", float((d / p)[x > 0].max()))
# oracle per-pixel loglik in python (mirror of pixel_lp_f32)
import math
def plp(pv, xv):
    f = np.float32; e = f(1.1920928955078125e-07)
    pt = min(max(f(pv), e), f(1) - e)
    l1 = f(math.log(float(pt))); l2 = f(math.log1p(-float(pt))); l = f(l1 - l2)
    ex = f(math.exp(-abs(float(l)))); lg = f(math.log1p(float(ex))); ls = f(min(l, f(0)) - lg)
    tt = f((f(1) - f(xv)) * l)
    return float(-(tt - ls))
tot = sum(plp(float(a), float(b)) for a, b in zip(po.reshape(-1), x.reshape(-1)))
print("python mirror of oracle pixel_lp on oracle pimg:", tot)
one = [plp(1e-4, 1.0), float(lp_t[(p == 1e-4) & (x == 1)][0])]
print("lp(eps, x=1): oracle-mirror", one[0], "torch", one[1], " lp(eps,x=0):", plp(1e-4, 0.0), float(lp_t[(p == 1e-4) & (x == 0)][0]))
dl = (lp_t - lp_o)
idx = torch.argsort(dl.abs().reshape(-1), descending=True)[:8]
for i in idx.tolist():
    r, c = divmod(i, 105)
    print(f"px ({r},{c}) x={int(x[r,c])} p_kernel={float(p[r,c]):.9e} p_oracle={float(po[r,c]):.9e} lp_k={float(lp_t[r,c]):.6f} lp_o={float(lp_o[r,c]):.6f}")
print("sum |dl|>1e-6 count", int((dl.abs() > 1e-6).sum()), "sum dl", float(dl.sum()))
print("n pixels where p differs", int((p != po).sum()), "n where kernel==eps but oracle!=eps", int(((p == 1e-4) & (po != 1e-4)).sum()), "reverse", int(((p != 1e-4) & (po == 1e-4)).sum()))
o64 = Oracle(np.float64)
ref64 = o64.token_batch(C.astype(np.float64), wb["tok_stroke_off"], wb["stroke_sub_off"], wb["ctrl"], wb["invscale"], wb["position"], wb["eps"], wb["sigma"], imgs, img_idx=np.tile(np.arange(5, dtype=np.int32), 40))
l64 = ref64["ll"].reshape(40, 5); l32 = ref["ll"].reshape(40, 5).astype(np.float64)
ek = np.abs(ll - l64) / np.abs(l64); eo = np.abs(l32 - l64) / np.abs(l64)
print("vs fp64: kernel max rel %.2e median %.2e | fp32 oracle (direct conv) max rel %.2e median %.2e" % (ek.max(), np.median(ek), eo.max(), np.median(eo)))
print("kernel closer than oracle32 in %d of %d pairs" % ((ek <= eo).sum(), ek.size))
