"""CPU oracle for log P(type) -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A numpy restatement (pure-Python loops: small cases only) of CharacterTypeDist.score_type
(pybpl/model/type_dist.py:98-125):

    score_k                                                 :163-185   log pkappa[k-1]
    per stroke  StrokeTypeDist.score_part_type              :507-531
                    score_nsub                              :269-294   log pmat_nsub[k-1][nsub-1]
                    score_subIDs                            :328-353   log start / transition probabilities along the id chain
                    score_shapes_type                       :388-423   MultivariateNormal(mu[id], Sigma[id]).log_prob
                    score_invscales_type                    :450-478   Gamma(con[id], rate[id]).log_prob
                RelationTypeDist.score_relation_type        :599-650   category (later strokes), then 'unihist': position
                    histogram of the stroke index (library/spatial_OLD/spatial_model.py:86-112, spatial_hist.py:145-167,
                    235-256: log p of the bin - log bin area, -inf outside); else -log(#earlier strokes)
                    [- log(#sub-strokes of the stroke attached to) - log(ub - lb) for 'mid']

and of its gradient with respect to the shapes and scales (the leaves autograd can reach: the histogram score
detaches through numpy, a uniform density is flat).  `tab` is pybpl_b200/data/type_lib.npz (tables produced from the
reference's Library).  Pinned against tests/golden/type_prior.npz by tests/test_oracle_golden.py.
"""
import math

import numpy as np


def score(d, tab, dtype=np.float64, grad=False):
    f = dtype
    K, nsub = d["K"], d["nsub"]
    ctrl = d["ctrl_type"].reshape(-1, 10).astype(f); inv = d["invscale_type"].astype(f)
    ids = d["subid"]
    P = len(K)
    ll = np.zeros(P, f)
    g = dict(ctrl_type=np.zeros_like(ctrl), invscale_type=np.zeros_like(inv)) if grad else None
    sso = np.concatenate([[0], np.cumsum(nsub)])
    xlab, ylab = tab["sp_xlab"].astype(np.float64), tab["sp_ylab"].astype(np.float64)
    nb = len(xlab) - 1
    k0 = 0
    for p in range(P):
        k = int(K[p])
        tot = f(tab["logp_kappa"][k - 1])
        for i in range(k):
            ki = k0 + i
            n = int(nsub[ki])
            tot += f(tab["logp_nsub"][k - 1][n - 1])
            prev = -1
            for s in range(sso[ki], sso[ki + 1]):
                sid = int(ids[s])
                tot += f(tab["logp_start"][sid] if prev < 0 else tab["logp_T"][prev][sid])
                prev = sid
                L = tab["shape_L"][sid].astype(f)
                y = np.linalg.solve(L, ctrl[s] - tab["shape_mu"][sid].astype(f))            # L^-1 (x - mu)
                tot += -0.5 * (y @ y) - np.sum(np.log(np.diag(L))) - f(5 * math.log(2 * math.pi))
                a, b = f(tab["gamma_con"][sid]), f(tab["gamma_rate"][sid])
                x = inv[s]
                tot += a * np.log(b) + (a - 1) * np.log(x) - b * x - f(math.lgamma(a))
                if grad:
                    g["ctrl_type"][s] = -np.linalg.solve(L.T, y)
                    g["invscale_type"][s] = (a - 1) / x - b
            cat = int(d["rel_cat"][ki])
            if i > 0:
                tot += f(tab["logp_mix"][cat])
            if cat == 0:
                h = min(i, tab["sp_logp"].shape[0] - 1)
                gx, gy = float(d["gpos"][ki][0]), float(d["gpos"][ki][1])
                # numpy.histogram2d bins: [e_j, e_j+1), the last one closed
                def bin_of(v, e):
                    if v < e[0] or v > e[-1]:
                        return -1
                    return min(int(np.searchsorted(e, v, side="right")) - 1, nb - 1)
                xi, yi = bin_of(gx, xlab.astype(np.float32)), bin_of(gy, ylab.astype(np.float32))
                if xi < 0 or yi < 0:
                    tot += f(-np.inf)
                else:
                    tot += f(tab["sp_logp"][h][xi * nb + yi]) - f(tab["sp_log_binarea"])
            else:
                ai = int(d["attach_ix"][ki])
                if not 0 <= ai < i:                       # Categorical(ones(nprev)).log_prob outside its support (:636-638)
                    tot += f(-np.inf)
                    continue
                tot += -f(math.log(i))
                if cat == 3:
                    na = int(nsub[k0 + ai])
                    es = float(d["eval_spot_type"][ki]) if "eval_spot_type" in d else 2.0
                    if 0 <= int(d["attach_subix"][ki]) < na and 2.0 <= es <= 6.0:
                        tot += -f(math.log(na)) - f(math.log(4.0))                             # Uniform(2, ncpt + 1) (:639-647)
                    else:
                        tot += f(-np.inf)
        ll[p] = tot
        k0 += k
    if grad:
        g["ctrl_type"] = g["ctrl_type"].reshape(-1, 5, 2)
        return ll, g
    return ll
