"""CPU oracle for log P(token | type) -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A numpy restatement (pure-Python loops over strokes: small cases only) of the reference's token score,
CharacterTokenDist.score_token (pybpl/model/token_dist.py:264-287):

    sum over strokes i of
        StrokeTokenDist.score_part_token        token_dist.py:432-455  (shapes :334-354, invscales :379-408)
      + RelationTokenDist.score_relation_token  token_dist.py:489-515  (score_eval_spot_token :542-572)
      + CharacterTokenDist.score_location       token_dist.py:131-153  (attach point objects/relation.py:34-67,
                                                                        trajectories objects/part.py:290-328,
                                                                        spline row splines.py:19-54,72-93,108-138)
    + score_affine (0, :169-170) + score_image_noise (0, :186-187) + score_image_blur (:205-224)

and of its gradient with respect to every token- and type-level parameter (what autograd returns for the
leaves listed by objects/concept.py:93-107 and :243-258), written out analytically.  The CUDA kernel
(pybpl_b200/csrc/bpl_token_prior.cu) follows the same formulas; tests/test_oracle_golden.py pins this file
against tests/golden/token_prior.npz (outputs of the live reference) and against finite differences of its
own forward in float64.
"""
import math

import numpy as np

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)
REL_UNIHIST, REL_START, REL_END, REL_MID = 0, 1, 2, 3


def _erf(x):
    return math.erf(float(x))


def _Phi(z):
    return 0.5 * (1.0 + _erf(z / math.sqrt(2.0)))


def _phi(z):
    return math.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)


def basis_raw(u, nland, dtype):
    """vectorized_bspline_coeff for one time u (splines.py:19-54): B[k] and dB[k]/du, before normalisation."""
    B = np.zeros(nland, dtype)
    dB = np.zeros(nland, dtype)
    for k in range(nland):
        d = dtype(u) - dtype(k)
        if 0 <= d < 1:
            B[k] = d ** 3; dB[k] = 3 * d * d
        elif 1 <= d < 2:
            e = d - 1
            B[k] = -3 * e ** 3 + 3 * e * e + 3 * e + 1; dB[k] = -9 * e * e + 6 * e + 3
        elif 2 <= d < 3:
            e = d - 2
            B[k] = 3 * e ** 3 - 6 * e * e + 4; dB[k] = 9 * e * e - 12 * e
        elif 3 <= d < 4:
            e = d - 3
            B[k] = (1 - e) ** 3; dB[k] = -3 * (1 - e) ** 2
    return B / dtype(6), dB / dtype(6)


def basis_row(u, nland, dtype):
    """coefficient_mat row (splines.py:72-93): c = B / sum B, and dc/du."""
    B, dB = basis_raw(u, nland, dtype)
    tot = B.sum()
    c = B / tot
    dc = (dB - c * dB.sum()) / tot
    return c, dc


def score(d, dtype=np.float64, grad=False, C0=None, C1=None):
    """d: dict of packed arrays (the layout of tests/golden/token_prior.npz).  Returns ll[P] and, with grad=True,
    a dict of gradients named like the inputs.  C0 / C1: first / last row of coefficient_mat(ncpt, neval)."""
    f = dtype
    K = d["K"]; nsub = d["nsub"]
    sig_s, sig_i, sig_a, sig_x, sig_y, blur_lo, blur_hi = [f(v) for v in d["consts"]]
    ncpt = d["ctrl"].shape[1]
    lb, ub = f(2), f(ncpt + 1)                                       # bspline_gen_s (splines.py:57-69)
    if C0 is None:
        C0 = basis_row(lb, ncpt, f)[0]; C1 = basis_row(ub, ncpt, f)[0]
    C0 = np.asarray(C0, f); C1 = np.asarray(C1, f)
    ctrl = d["ctrl"].astype(f); ctrl_ty = d["ctrl_type"].astype(f)
    inv = d["invscale"].astype(f); inv_ty = d["invscale_type"].astype(f)
    pos = d["position"].astype(f); gpos = d["gpos"].astype(f)
    es = d["eval_spot"].astype(f); es_ty = d["eval_spot_type"].astype(f)
    P = len(K)
    ll = np.zeros(P, f)
    g = None
    if grad:
        g = {k: np.zeros_like(v) for k, v in dict(ctrl=ctrl, ctrl_type=ctrl_ty, invscale=inv, invscale_type=inv_ty,
                                                  position=pos, gpos=gpos, eval_spot=es, eval_spot_type=es_ty).items()}
    stroke_sub = np.concatenate([[0], np.cumsum(nsub)])
    k0 = 0
    for p in range(P):
        tot = f(0)
        for i in range(K[p]):
            ki = k0 + i
            # ---- score_part_token (token_dist.py:432-455) ----
            for s in range(stroke_sub[ki], stroke_sub[ki + 1]):
                diff = ctrl[s] - ctrl_ty[s]                                                    # Normal(shapes_type, sigma_shape) :334-354
                tot += np.sum(-(diff * diff) / (2 * sig_s * sig_s) - f(math.log(sig_s)) - f(LOG_SQRT_2PI))
                di = inv[s] - inv_ty[s]                                                        # :379-408
                l = -(di * di) / (2 * sig_i * sig_i) - f(math.log(sig_i)) - f(LOG_SQRT_2PI)
                p_above = 1 - f(_Phi((0 - inv_ty[s]) / sig_i))                                 # :399-401
                l = l - f(math.log(p_above)) if p_above > 0 else f(np.inf)
                if inv[s] <= 0:                                                                # :404-406
                    l = f(-np.inf)
                tot += l
                if grad:
                    g["ctrl"][s] += -diff / (sig_s * sig_s); g["ctrl_type"][s] += diff / (sig_s * sig_s)
                    g["invscale"][s] += -di / (sig_i * sig_i)
                    z = inv_ty[s] / sig_i
                    g["invscale_type"][s] += di / (sig_i * sig_i) - f(_phi(z)) / (sig_i * f(_Phi(z)))
            # ---- score_relation_token (token_dist.py:489-515, 542-572) ----
            cat = int(d["rel_cat"][ki])
            if cat == REL_MID:
                u = es[ki]
                if u < lb or u > ub:
                    tot += f(-np.inf)
                else:
                    de = u - es_ty[ki]
                    zu, zl = (ub - es_ty[ki]) / sig_a, (lb - es_ty[ki]) / sig_a
                    p_within = f(_Phi(zu)) - f(_Phi(zl))
                    tot += -(de * de) / (2 * sig_a * sig_a) - f(math.log(sig_a)) - f(LOG_SQRT_2PI) - f(math.log(p_within))
                    if grad:
                        g["eval_spot"][ki] += -de / (sig_a * sig_a)
                        g["eval_spot_type"][ki] += de / (sig_a * sig_a) - (f(_phi(zl)) - f(_phi(zu))) / (sig_a * p_within)
            # ---- score_location (token_dist.py:131-153) with get_attach_point (relation.py:34-67) ----
            if cat == REL_UNIHIST:
                base = gpos[ki]
            else:
                j = k0 + int(d["attach_ix"][ki])
                sj0, sj1 = stroke_sub[j], stroke_sub[j + 1]
                # vanilla_to_motor (part.py:305-328): first/last trajectory point of every sub-stroke, chained
                end = pos[j].copy()
                ends = []                       # end point of the previous sub-stroke, per sub-stroke
                for s in range(sj0, sj1):
                    Y = inv[s] * ctrl[s]
                    a, b = C0 @ Y, C1 @ Y
                    ends.append(end.copy())
                    off = a - end
                    end = b - off
                if cat == REL_START:
                    Y = inv[sj0] * ctrl[sj0]
                    a = C0 @ Y
                    base = a - (a - pos[j])                                      # motor[0][0]
                elif cat == REL_END:
                    base = end                                                   # motor[-1][-1]
                else:
                    sub = sj0 + int(d["attach_subix"][ki])
                    Y = inv[sub] * ctrl[sub]
                    off = C0 @ Y - ends[sub - sj0]
                    c, dc = basis_row(es[ki], ncpt, f)                           # get_stk_from_bspline(bspline, s=eval_spot_token)
                    base = c @ (Y - off)                                         # motor_spline[:, :, subix] = Y - offset (part.py:323)
            r = pos[ki] - base
            sg = np.array([sig_x, sig_y], f)
            tot += np.sum(-(r * r) / (2 * sg * sg) - np.log(sg) - f(LOG_SQRT_2PI))
            if grad:
                gb = r / (sg * sg)                                               # d ll / d base
                g["position"][ki] += -gb
                if cat == REL_UNIHIST:
                    g["gpos"][ki] += gb
                elif cat == REL_START:
                    g["position"][j] += gb                                       # d/d a cancels: a - (a - pos)
                else:
                    last = sj1 if cat == REL_END else sub
                    w_end = f(1) if cat == REL_END else c.sum()                 # weight of the chained end point
                    for s in range(sj0, last):                                   # end_s = end_{s-1} + (C1 - C0) @ Y_s
                        gY = np.outer(C1 - C0, gb) * w_end
                        g["ctrl"][s] += inv[s] * gY
                        g["invscale"][s] += np.sum(ctrl[s] * gY)
                    g["position"][j] += gb * w_end
                    if cat == REL_MID:
                        gY = np.outer(c - c.sum() * C0, gb)
                        g["ctrl"][sub] += inv[sub] * gY
                        g["invscale"][sub] += np.sum(ctrl[sub] * gY)
                        g["eval_spot"][ki] += np.dot(dc @ (Y - off), gb)
        # ---- image blur (token_dist.py:205-224): Uniform(lb, ub).log_prob; affine and noise score 0 ----
        sg_b = f(d["sigma"][p])
        tot += -f(math.log(blur_hi - blur_lo)) if (blur_lo <= sg_b < blur_hi) else f(-np.inf)
        ll[p] = tot
        k0 += K[p]
    return (ll, g) if grad else ll
