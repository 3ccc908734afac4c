"""Repeats the two fused backward kernels on the mixed batch of test_backward_kernels_agree_on_every_mode and reports which
kernel's gradients change from run to run, and on which token."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybpl_b200 import ops, workload
DEV = torch.device("cuda:0")
T = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=DEV)
w = workload.synth_tokens(160, seed=77, sigma_mode="uniform")
rng = np.random.default_rng(5)
w["sigma"][::7] = 0.0; w["sigma"][3::11] = -1.0; w["eps"][::5] = 0.0
w["position"][w["tok_stroke_off"][10]:w["tok_stroke_off"][11]] += 400.0
aff = np.tile(np.array([1, 1, 0, 0], np.float32), (160, 1))
aff[::3] = np.stack([rng.uniform(0.8, 1.25, 54), rng.uniform(0.8, 1.25, 54), rng.uniform(-6, 6, 54), rng.uniform(-6, 6, 54)], 1)
tso = np.concatenate([[0], w["tok_stroke_off"]]).astype(np.int64)
eps = np.concatenate([[1e-4], w["eps"]]).astype(np.float32); sig = np.concatenate([[2.0], w["sigma"]]).astype(np.float32)
aff = np.concatenate([[[1, 1, 0, 0]], aff]).astype(np.float32)
P = 161
img = (rng.random((3, 105, 105)) < 0.04).astype(np.uint8)
imgs = ops.pack_images(T(img, torch.uint8))
idx = T(rng.integers(0, 3, P), torch.int32)
g_ll = T(rng.standard_normal(P))
sso = w["stroke_sub_off"]
sub_tok = np.repeat(np.arange(P), np.diff(sso[tso]))
def run(kern):
    os.environ.pop("BPL_BWD_CANVAS", None); os.environ.pop("BPL_BWD_TEAMS", None); os.environ[kern] = "1"
    leaves = [T(w["ctrl"]), T(w["invscale"]), T(w["position"]), T(aff), T(eps), T(sig)]
    for t in leaves: t.requires_grad_(True)
    batch = ops.TokenBatch(tso, sso, leaves[0], leaves[1], leaves[2], leaves[4], leaves[5], affine=leaves[3])
    ll, off = ops.score_tokens(batch, imgs, idx)
    g = torch.autograd.grad((ll * g_ll).sum(), leaves)
    return ll.detach().cpu().numpy(), g[0].cpu().numpy()
ref = {}
for rep in range(40):
    for kern in ("BPL_BWD_CANVAS", "BPL_BWD_TEAMS"):
        ll, gc = run(kern)
        if kern not in ref:
            ref[kern] = (ll, gc); continue
        d = np.abs(gc - ref[kern][1]).reshape(gc.shape[0], -1).max(1)
        bad = np.flatnonzero(d > 1e-4 * np.abs(ref[kern][1]).max())
        if bad.size:
            print(f"rep {rep} {kern}: {bad.size} sub-strokes differ from the first run; tokens {sorted(set(sub_tok[bad]))}; max abs diff {d.max():.3e}; ll diff there "
                  f"{np.abs(ll - ref[kern][0])[sorted(set(sub_tok[bad]))]}; sigma {sig[sorted(set(sub_tok[bad]))]} eps {eps[sorted(set(sub_tok[bad]))]} aff {aff[sorted(set(sub_tok[bad]))]}")
a, b = ref["BPL_BWD_CANVAS"][1], ref["BPL_BWD_TEAMS"][1]
d = np.abs(a - b).reshape(a.shape[0], -1).max(1)
print("canvas vs teams (first runs): max", d.max() / np.abs(a).max(), "worst sub-stroke", d.argmax(), "token", sub_tok[d.argmax()])
