"""The reference's OWN torch implementation of the hot path, timed on the host cores: N single-thread
worker processes, each calling the unmodified `pybpl.model.image_dist.CharacterImageDist.score_image`
(pybpl/model/image_dist.py:60-80) from baseline/_ref on its share of a packed numpy token batch.

Used only by bench.py (`cpu_baseline_torch`) and the tests.  Each particle becomes a real reference CharacterToken
(pybpl/objects/concept.py:217-241) with StrokeTokens (objects/part.py:188-249) and independent RelationTokens, i.e.
exactly what `CharacterModel.score_image` would be handed.  fwd+bwd uses the one-token in-memory patch of
rendering.py:97 (baseline/ref_loader.py) and differentiates w.r.t. shapes, invscales, position and blur_sigma.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tokens_from(w, lo, hi):
    """Reference CharacterToken objects for tokens [lo, hi) of the numpy batch `w`."""
    import torch
    from pybpl.objects import CharacterToken, StrokeToken
    from pybpl.objects.relation import RelationIndependent, RelationToken
    tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
    lim = torch.tensor([0.0, 105.0]), torch.tensor([-105.0, 0.0])
    out = []
    for p in range(lo, hi):
        P, R = [], []
        for k in range(tso[p], tso[p + 1]):
            s0, s1 = sso[k], sso[k + 1]
            shapes = torch.from_numpy(np.ascontiguousarray(w["ctrl"][s0:s1])).permute(1, 2, 0).contiguous()   # (5,2,nsub)
            st = StrokeToken(shapes, torch.from_numpy(w["invscale"][s0:s1].copy()), lim[0], lim[1])
            st.position = torch.from_numpy(w["position"][k].copy())
            P.append(st)
            R.append(RelationToken(RelationIndependent("unihist", st.position.clone(), lim[0], lim[1])))
        out.append(CharacterToken(P, R, None, torch.tensor(float(w["eps"][p])), torch.tensor(float(w["sigma"][p]))))
    return out


def _worker(w, lo, hi, image, grad, budget_s):
    import warnings
    warnings.filterwarnings("ignore")
    import torch
    torch.set_num_threads(1)
    from baseline import ref_loader
    ref_loader.load(grad_patch=grad)
    from pybpl.model.image_dist import CharacterImageDist
    dist = CharacterImageDist(None)
    toks = _tokens_from(w, lo, hi)
    img = torch.from_numpy(image.astype(np.float32))
    lls = []

    def one(t):
        if not grad:
            with torch.no_grad():
                return float(dist.score_image(t, img))
        leaves = [x for p in t.part_tokens for x in (p.shapes, p.invscales, p.position)] + [t.blur_sigma]
        for x in leaves:
            x.requires_grad_(True)
        ll = dist.score_image(t, img)
        ll.backward()
        return float(ll)
    one(toks[0])                                   # warm-up (lru caches of the basis table, torch dispatch)
    n = 0
    t0 = time.perf_counter()
    for t in toks:
        lls.append(one(t))
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return n, time.perf_counter() - t0, lls


def score_rate(w, image, n_procs=None, per_proc=64, grad=False, budget_s=20.0):
    """Aggregate particles/s of n_procs single-thread reference processes (plain subprocesses of this module, started
    together), each scoring (fwd, or fwd+bwd) up to `per_proc` tokens of `w` against `image` within `budget_s` seconds.
    Returns (rate, n_particles, procs, ll list of process 0)."""
    import json
    import subprocess
    import tempfile
    n_procs = n_procs or (os.cpu_count() or 1)
    per_proc = max(1, min(per_proc, w["P"] // n_procs))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "batch.npz")
        np.savez(path, image=image, **{k: v for k, v in w.items() if isinstance(v, np.ndarray)})
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
        procs = [subprocess.Popen([sys.executable, "-W", "ignore", "-m", "baseline.torch_ref", path, str(i * per_proc),
                                   str((i + 1) * per_proc), str(int(grad)), str(budget_s)], cwd=ROOT, env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for i in range(n_procs)]
        res = []
        for pr in procs:
            try:
                out, err = pr.communicate(timeout=budget_s + 240)
            except subprocess.TimeoutExpired:
                pr.kill()
                raise RuntimeError("reference worker timed out")
            if pr.returncode != 0:
                raise RuntimeError("reference worker failed: " + err[-800:])
            res.append(json.loads(out.strip().splitlines()[-1]))
    # every process runs concurrently on its own core: the aggregate rate is the sum of the per-process rates
    rate = sum(r["n"] / r["dt"] for r in res)
    return rate, sum(r["n"] for r in res), n_procs, res[0]["ll"]


if __name__ == "__main__":
    import json
    _path, _lo, _hi, _grad, _budget = sys.argv[1:6]
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    _d = dict(np.load(_path))
    _n, _dt, _lls = _worker(_d, int(_lo), int(_hi), _d["image"], bool(int(_grad)), float(_budget))
    print(json.dumps({"n": _n, "dt": _dt, "ll": _lls}))
