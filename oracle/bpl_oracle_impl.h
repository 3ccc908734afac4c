/*
 * bpl_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Generic-precision body of the CPU oracle.  Included twice by bpl_oracle.c:
 * once with REAL=float (the fp32 restatement of what the reference computes with
 * torch-CPU fp32 tensors) and once with REAL=double (the same algorithm in fp64,
 * used only to measure the reference's own fp32 noise floor).
 *
 * Every function restates one piece of rfeinman/pyBPL (paths relative to
 * /root/reference) and cites the lines it follows.  Nothing under pybpl_b200/
 * may include, link or call this file: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * Pinning: tests/test_oracle_golden.py checks every function here against
 * the tests/golden npz fixtures, which tests/golden/make_golden.py produced by running the
 * live reference (the reference's own tests hold no values for this path:
 * pybpl/tests/test_vanilla_to_motor.py:47-48 asserts shapes only).
 */

#ifndef REAL
#error "define REAL, SUF and the math macros before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ---------------------------------------------------------------------------
 * B-spline basis.  pybpl/splines.py:19-54 (vectorized_bspline_coeff),
 * :57-69 (bspline_gen_s), :72-93 (coefficient_mat: row normalisation).
 * ------------------------------------------------------------------------- */
static void FN(basis_row)(int nland, REAL s, REAL *row) {
    REAL sum = 0;
    for (int i = 0; i < nland; ++i) {
        REAL vi = (REAL)i, c = 0;
        if (s >= vi && s < vi + 1) { /* splines.py:37-39 */
            REAL d = s - vi;
            c = d * d * d;
        } else if (s >= vi + 1 && s < vi + 2) { /* :41-43  -3d^3+3d^2+3d+1 */
            REAL d = s - vi - 1;
            c = ((REAL)-3 * (d * d * d) + (REAL)3 * (d * d)) + (REAL)3 * d + (REAL)1;
        } else if (s >= vi + 2 && s < vi + 3) { /* :45-47  3d^3-6d^2+4 */
            REAL d = s - vi - 2;
            c = ((REAL)3 * (d * d * d) - (REAL)6 * (d * d)) + (REAL)4;
        } else if (s >= vi + 3 && s < vi + 4) { /* :49-51  (1-d)^3 */
            REAL d = s - vi - 3;
            REAL e = (REAL)1 - d;
            c = e * e * e;
        }
        c = c / (REAL)6; /* :53 */
        row[i] = c;
        sum += c;
    }
    for (int i = 0; i < nland; ++i) row[i] = row[i] / sum; /* splines.py:91 */
}

/* linspace(2, nland+1, neval): splines.py:66-67.  torch fills the first half
 * from the start and the second half back from the end. */
static void FN(gen_s)(int nland, int neval, REAL *s) {
    REAL lo = 2, hi = (REAL)(nland + 1);
    if (neval == 1) { s[0] = lo; return; }
    REAL step = (hi - lo) / (REAL)(neval - 1);
    for (int j = 0; j < neval; ++j)
        s[j] = (j < neval / 2) ? lo + step * (REAL)j : hi - step * (REAL)(neval - 1 - j);
}

static void FN(basis_table)(int nland, int neval, REAL *C) {
    REAL *s = (REAL *)malloc(sizeof(REAL) * (size_t)neval);
    FN(gen_s)(nland, neval, s);
    for (int j = 0; j < neval; ++j) FN(basis_row)(nland, s[j], C + (size_t)j * nland);
    free(s);
}

/* X = C @ Y, splines.py:136.  torch's CPU matmul for these shapes is bit-equal
 * to a sequential fma chain over k starting from 0. */
static void FN(spline_eval)(const REAL *C, int neval, int nland, const REAL *Y, REAL *X) {
    for (int j = 0; j < neval; ++j)
        for (int c = 0; c < 2; ++c) {
            REAL acc = 0;
            for (int k = 0; k < nland; ++k) acc = R_FMA(C[(size_t)j * nland + k], Y[2 * k + c], acc);
            X[2 * j + c] = acc;
        }
}

/* Sum |X[j+1]-X[j]|: pybpl/util/stroke.py:22-28 (dist_along_traj). */
static REAL FN(dist_along)(const REAL *X, int n) {
    REAL tot = 0;
    for (int j = 0; j + 1 < n; ++j) {
        REAL dx = X[2 * (j + 1)] - X[2 * j], dy = X[2 * (j + 1) + 1] - X[2 * j + 1];
        tot += R_SQRT(dx * dx + dy * dy);
    }
    return tot;
}

/* Adaptive neval, splines.py:129-133: clamp(ceil(dist/1.5), 10, 200). */
static int FN(adaptive_neval)(const REAL *C200, int nland, const REAL *Y, REAL *dist_out) {
    REAL X[400];
    FN(spline_eval)(C200, 200, nland, Y, X);
    REAL d = FN(dist_along)(X, 200);
    if (dist_out) *dist_out = d;
    REAL q = R_CEIL(d / (REAL)1.5);
    if (q < 10) q = 10;
    if (q > 200) q = 200;
    return (int)q;
}

/* ---------------------------------------------------------------------------
 * fit_bspline_to_traj, pybpl/splines.py:141-171: Y = argmin |C Y - X|, residuals.
 * The arithmetic lives in torch.linalg.lstsq(driver='gels') = LAPACK ?gels
 * (un-vendored; torch 2.11 links MKL 2024.2): Householder QR of C (geqrf), Q^T
 * applied to X (ormqr), back-substitution with R (trtrs), residual = squared norm
 * of the rows of Q^T X below n.  Restated here unblocked.
 *   C [m][n] row-major, X [m][2], Y [n][2], resid[2] (0 when m <= n).
 * ------------------------------------------------------------------------- */
static void FN(lstsq)(const REAL *C, int m, int n, const REAL *X, REAL *Y, REAL *resid) {
    REAL *A = (REAL *)malloc(sizeof(REAL) * (size_t)m * n);
    REAL *B = (REAL *)malloc(sizeof(REAL) * (size_t)m * 2);
    memcpy(A, C, sizeof(REAL) * (size_t)m * n);
    memcpy(B, X, sizeof(REAL) * (size_t)m * 2);
    for (int k = 0; k < n && k < m; ++k) {
        /* Householder vector for column k, rows k..m-1 */
        REAL norm = 0;
        for (int i = k; i < m; ++i) norm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        norm = R_SQRT(norm);
        if (norm == 0) continue;
        REAL alpha = A[(size_t)k * n + k] > 0 ? -norm : norm;
        REAL v0 = A[(size_t)k * n + k] - alpha;
        /* v = (v0, A[k+1..m-1][k]); beta = 2 / (v^T v) */
        REAL vtv = v0 * v0;
        for (int i = k + 1; i < m; ++i) vtv += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        if (vtv == 0) continue;
        REAL beta = (REAL)2 / vtv;
        for (int j = k + 1; j < n; ++j) {
            REAL dot = v0 * A[(size_t)k * n + j];
            for (int i = k + 1; i < m; ++i) dot += A[(size_t)i * n + k] * A[(size_t)i * n + j];
            dot *= beta;
            A[(size_t)k * n + j] -= dot * v0;
            for (int i = k + 1; i < m; ++i) A[(size_t)i * n + j] -= dot * A[(size_t)i * n + k];
        }
        for (int c = 0; c < 2; ++c) {
            REAL dot = v0 * B[2 * k + c];
            for (int i = k + 1; i < m; ++i) dot += A[(size_t)i * n + k] * B[2 * i + c];
            dot *= beta;
            B[2 * k + c] -= dot * v0;
            for (int i = k + 1; i < m; ++i) B[2 * i + c] -= dot * A[(size_t)i * n + k];
        }
        A[(size_t)k * n + k] = alpha;
    }
    for (int c = 0; c < 2; ++c) {
        for (int k = n - 1; k >= 0; --k) {
            REAL acc = B[2 * k + c];
            for (int j = k + 1; j < n; ++j) acc -= A[(size_t)k * n + j] * Y[2 * j + c];
            Y[2 * k + c] = acc / A[(size_t)k * n + k];
        }
        REAL r = 0;
        for (int i = n; i < m; ++i) r += B[2 * i + c] * B[2 * i + c];
        resid[c] = r;
    }
    free(A); free(B);
}

/* ---------------------------------------------------------------------------
 * vanilla_to_motor, pybpl/objects/part.py:305-328: scale, evaluate, chain.
 *   ctrl     [nsub][ncpt][2]  (the reference's shapes[:, :, i] per sub-stroke)
 *   motor    [nsub][neval][2]
 *   mspline  [nsub][ncpt][2]  or NULL (motor_spline, part.py:324)
 * ------------------------------------------------------------------------- */
static void FN(vanilla_to_motor)(const REAL *C, int neval, int ncpt, int nsub, const REAL *ctrl,
                                 const REAL *invscale, const REAL *pos, REAL *motor, REAL *mspline) {
    REAL prev[2] = {pos[0], pos[1]};
    REAL Ys[2 * 64];
    for (int i = 0; i < nsub; ++i) {
        for (int k = 0; k < 2 * ncpt; ++k) Ys[k] = invscale[i] * ctrl[(size_t)i * 2 * ncpt + k]; /* :318 */
        REAL *m = motor + (size_t)i * neval * 2;
        FN(spline_eval)(C, neval, ncpt, Ys, m); /* :320 */
        REAL off[2] = {m[0] - prev[0], m[1] - prev[1]}; /* :322 */
        for (int j = 0; j < neval; ++j) { m[2 * j] -= off[0]; m[2 * j + 1] -= off[1]; } /* :323 */
        if (mspline)
            for (int k = 0; k < ncpt; ++k) { /* :324 */
                mspline[(size_t)i * 2 * ncpt + 2 * k] = Ys[2 * k] - off[0];
                mspline[(size_t)i * 2 * ncpt + 2 * k + 1] = Ys[2 * k + 1] - off[1];
            }
        prev[0] = m[2 * (neval - 1)]; prev[1] = m[2 * (neval - 1) + 1]; /* :326 */
    }
}

/* apply_warp, pybpl/util/affine.py:48-53 with com_char, util/stroke.py:134-135.
 * In place over npts points; returns com and B for the backward. */
static void FN(apply_warp)(REAL *pts, size_t npts, const REAL *A, REAL *com, REAL *B) {
    double sx = 0, sy = 0;
    for (size_t i = 0; i < npts; ++i) { sx += pts[2 * i]; sy += pts[2 * i + 1]; }
    com[0] = (REAL)(sx / (double)npts); com[1] = (REAL)(sy / (double)npts);
    B[0] = A[0]; B[1] = A[1];
    B[2] = A[2] - (A[0] - (REAL)1) * com[0];
    B[3] = A[3] - (A[1] - (REAL)1) * com[1];
    for (size_t i = 0; i < npts; ++i) {
        pts[2 * i] = pts[2 * i] * B[0] + B[2];
        pts[2 * i + 1] = pts[2 * i + 1] * B[1] + B[3];
    }
}

/* ---------------------------------------------------------------------------
 * add_stroke, pybpl/rendering.py:78-127 (+ check_bounds :27-31, flip :52-53).
 * The per-point record is kept so the backward can re-use it.
 * ------------------------------------------------------------------------- */
typedef struct {
    int n_in;      /* surviving points */
    int branch;    /* 0 none drawn, 1 single point, 2 fill, 3 rescale, 4 normal */
    REAL sumink;   /* sum of ink before the min-ink fix-up */
} FN(substroke_info);

/* pts: [n][2] motor space.  idx/rc/ink sized n (compacted, first n_in valid).
 * Returns 1 if any point was out of bounds. */
static int FN(substroke_ink)(const bplo_params *ps, const REAL *pts, int n, int *idx, REAL *rc, REAL *ink,
                             FN(substroke_info) *info) {
    int any_out = 0, m = 0;
    for (int j = 0; j < n; ++j) {
        REAL r = pts[2 * j + 1] * (REAL)-1; /* rendering.py:52-53: flip, times [-1, 1] */
        REAL c = pts[2 * j] * (REAL)1;
        int out = (R_FLOOR(r) < 0) | (R_CEIL(r) >= (REAL)ps->H) | (R_FLOOR(c) < 0) | (R_CEIL(c) >= (REAL)ps->W);
        if (out) { any_out = 1; continue; } /* :85-90 */
        idx[m] = j; rc[2 * m] = r; rc[2 * m + 1] = c; ++m;
    }
    info->n_in = m; info->branch = 0; info->sumink = 0;
    if (m == 0) return any_out;
    if (m == 1) { /* :93-94 */
        ink[0] = (REAL)ps->ink_pp;
        info->branch = 1;
    } else { /* :95-99 */
        for (int k = 0; k + 1 < m; ++k) {
            REAL dr = rc[2 * (k + 1)] - rc[2 * k], dc = rc[2 * (k + 1) + 1] - rc[2 * k + 1];
            REAL d = R_SQRT(dr * dr + dc * dc);
            if (d > (REAL)ps->ink_max_dist) d = (REAL)ps->ink_max_dist;
            ink[k + 1] = ((REAL)ps->ink_pp * d) / (REAL)ps->ink_max_dist;
        }
        ink[0] = ink[1];
        info->branch = 4;
    }
    REAL sum = 0; /* :103 */
    for (int k = 0; k < m; ++k) sum += ink[k];
    info->sumink = sum;
    if (sum < (REAL)2.22e-6) { /* :104-105 */
        REAL v = (REAL)((double)ps->ink_pp / (double)m);
        for (int k = 0; k < m; ++k) ink[k] = v;
        info->branch = 2;
    } else if (sum < (REAL)ps->ink_pp) { /* :106-107 */
        REAL sc = (REAL)ps->ink_pp / sum;
        for (int k = 0; k < m; ++k) ink[k] *= sc;
        info->branch = 3;
    }
    return any_out;
}

static void FN(splat)(const bplo_params *ps, REAL *img, const REAL *rc, const REAL *ink, int m) {
    const int W = ps->W;
    /* rendering.py:111-125: four accumulate passes, each in point order */
    for (int pass = 0; pass < 4; ++pass)
        for (int k = 0; k < m; ++k) {
            REAL r = rc[2 * k], c = rc[2 * k + 1];
            REAL rf = R_FLOOR(r), cf = R_FLOOR(c), rce = R_CEIL(r), cce = R_CEIL(c);
            REAL fr = r - rf, fc = c - cf;
            REAL gr = (REAL)1 - fr, gc = (REAL)1 - fc;
            int ri, ci; REAL v;
            switch (pass) {
            case 0: ri = (int)rf; ci = (int)cf; v = ink[k] * gr * gc; break;
            case 1: ri = (int)rce; ci = (int)cf; v = ink[k] * fr * gc; break;
            case 2: ri = (int)rf; ci = (int)cce; v = ink[k] * gr * fc; break;
            default: ri = (int)rce; ci = (int)cce; v = ink[k] * fr * fc; break;
            }
            img[ri * W + ci] += v;
        }
}

/* zero-padded cross-correlation, pybpl/util/general.py:156-164 (imfilter -> F.conv2d) */
static void FN(corr2d)(const REAL *in, REAL *out, int H, int W, const REAL *ker, int kh, int kw) {
    int ph = kh / 2, pw = kw / 2;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            REAL acc = 0;
            for (int u = 0; u < kh; ++u) {
                int ii = i + u - ph;
                if (ii < 0 || ii >= H) continue;
                for (int v = 0; v < kw; ++v) {
                    int jj = j + v - pw;
                    if (jj < 0 || jj >= W) continue;
                    acc += in[ii * W + jj] * ker[u * kw + v];
                }
            }
            out[i * W + j] = acc;
        }
}

/* broadening kernel, rendering.py:151-168 */
static void FN(broaden_kernel)(const bplo_params *ps, REAL *Hk) {
    double a = ps->ink_a, b = ps->ink_b;
    double hs = ps->hinton ? b * (1 + a) : b; /* :153-156 */
    REAL base[9] = {(REAL)(a / 12), (REAL)(a / 6), (REAL)(a / 12), (REAL)(a / 6), (REAL)(1 - a),
                    (REAL)(a / 6),  (REAL)(a / 12), (REAL)(a / 6), (REAL)(a / 12)};
    for (int i = 0; i < 9; ++i) Hk[i] = (REAL)hs * base[i];
}

/* fspecial('gaussian'), pybpl/util/general.py:196-207 */
static void FN(gauss_kernel)(int fsize, REAL sigma, REAL *K) {
    int mu = fsize / 2;
    REAL sum = 0;
    for (int i = 0; i < fsize; ++i)
        for (int j = 0; j < fsize; ++j) {
            REAL dx = ((REAL)j - (REAL)mu) / sigma, dy = ((REAL)i - (REAL)mu) / sigma;
            REAL e = R_EXP((REAL)-0.5 * (dx * dx + dy * dy));
            K[i * fsize + j] = e;
            sum += e;
        }
    for (int i = 0; i < fsize * fsize; ++i) K[i] = K[i] / sum;
}

/* Saved intermediates of one render (everything the backward needs). */
typedef struct {
    REAL *I0, *I2, *I3, *I4, *I5, *I6; /* each H*W; see render_core */
    REAL Hk[9];
    REAL *K; /* fsize*fsize */
} FN(render_tape);

static void FN(tape_alloc)(FN(render_tape) * t, int HW, int fsize) {
    t->I0 = (REAL *)calloc((size_t)HW * 6, sizeof(REAL));
    t->I2 = t->I0 + HW; t->I3 = t->I2 + HW; t->I4 = t->I3 + HW; t->I5 = t->I4 + HW; t->I6 = t->I5 + HW;
    t->K = (REAL *)calloc((size_t)fsize * fsize, sizeof(REAL));
}
static void FN(tape_free)(FN(render_tape) * t) { free(t->I0); free(t->K); }

/* render_image, rendering.py:209-229 + broaden_and_blur :150-182, on motor-space
 * points given as nsub ragged sub-strokes (pt_off[nsub+1] into pts[.][2]).
 * Writes pimg[H*W]; returns ink_off_page. */
static int FN(render_core)(const bplo_params *ps, int nsub, const int *pt_off, const REAL *pts, REAL eps,
                           REAL sigma, REAL *pimg, FN(render_tape) * tape) {
    const int H = ps->H, W = ps->W, HW = H * W;
    int maxn = 1;
    for (int s = 0; s < nsub; ++s) if (pt_off[s + 1] - pt_off[s] > maxn) maxn = pt_off[s + 1] - pt_off[s];
    int *idx = (int *)malloc(sizeof(int) * (size_t)maxn);
    REAL *rc = (REAL *)malloc(sizeof(REAL) * 2 * (size_t)maxn);
    REAL *ink = (REAL *)malloc(sizeof(REAL) * (size_t)maxn);
    REAL *A = (REAL *)calloc((size_t)HW, sizeof(REAL)), *B = (REAL *)calloc((size_t)HW, sizeof(REAL));
    int off_page = 0;
    for (int s = 0; s < nsub; ++s) { /* rendering.py:218-220 */
        FN(substroke_info) info;
        int n = pt_off[s + 1] - pt_off[s];
        if (n <= 0) continue;
        off_page |= FN(substroke_ink)(ps, pts + 2 * (size_t)pt_off[s], n, idx, rc, ink, &info);
        if (info.n_in > 0) FN(splat)(ps, A, rc, ink, info.n_in);
    }
    if (tape) memcpy(tape->I0, A, sizeof(REAL) * HW);
    REAL Hk[9];
    FN(broaden_kernel)(ps, Hk);
    if (tape) memcpy(tape->Hk, Hk, sizeof(Hk));
    for (int it = 0; it < ps->ink_ncon; ++it) { /* :167-168 */
        FN(corr2d)(A, B, H, W, Hk, 3, 3);
        REAL *t = A; A = B; B = t;
    }
    if (tape) memcpy(tape->I2, A, sizeof(REAL) * HW);
    for (int i = 0; i < HW; ++i) if (A[i] > 1) A[i] = 1; /* :171 */
    if (tape) memcpy(tape->I3, A, sizeof(REAL) * HW);
    if (sigma > 0) { /* :174-177 */
        REAL *K = (REAL *)malloc(sizeof(REAL) * (size_t)ps->fsize * ps->fsize);
        FN(gauss_kernel)(ps->fsize, sigma, K);
        if (tape) memcpy(tape->K, K, sizeof(REAL) * (size_t)ps->fsize * ps->fsize);
        FN(corr2d)(A, B, H, W, K, ps->fsize, ps->fsize);
        if (tape) memcpy(tape->I4, B, sizeof(REAL) * HW);
        FN(corr2d)(B, A, H, W, K, ps->fsize, ps->fsize);
        free(K);
    }
    if (tape) memcpy(tape->I5, A, sizeof(REAL) * HW);
    for (int i = 0; i < HW; ++i) { if (A[i] < 0) A[i] = 0; if (A[i] > 1) A[i] = 1; } /* :180 */
    if (tape) memcpy(tape->I6, A, sizeof(REAL) * HW);
    if (eps > 0) { /* :226-227 */
        REAL a = (REAL)1 - eps;
        for (int i = 0; i < HW; ++i) pimg[i] = a * A[i] + eps * ((REAL)1 - A[i]);
    } else {
        memcpy(pimg, A, sizeof(REAL) * HW);
    }
    free(idx); free(rc); free(ink); free(A); free(B);
    return off_page;
}

/* ---------------------------------------------------------------------------
 * Bernoulli(pimg).log_prob(image).sum(), pybpl/model/image_dist.py:75-78, in
 * torch's formulation (torch/distributions/utils.py clamp_probs/probs_to_logits,
 * bernoulli.py log_prob -> -binary_cross_entropy_with_logits):
 *   pt = clamp(p, eps32, 1-eps32); l = log(pt) - log1p(-pt)
 *   lp = -((1-x)*l - (min(l,0) - log1p(exp(-|l|))))
 * Each libm call is evaluated in double and rounded to REAL (i.e. correctly
 * rounded for REAL=float), which reproduces torch-CPU on >99% of inputs and
 * on all background constants tried (numerics).
 * ------------------------------------------------------------------------- */
static REAL FN(pixel_lp)(REAL p, int x) {
    const REAL e = (REAL)1.1920928955078125e-07; /* torch.finfo(float32).eps */
    REAL pt = p < e ? e : (p > (REAL)1 - e ? (REAL)1 - e : p);
    REAL l1 = (REAL)log((double)pt);
    REAL l2 = (REAL)log1p(-(double)pt);
    REAL l = l1 - l2;
    REAL al = l < 0 ? -l : l;
    REAL ex = (REAL)exp(-(double)al);
    REAL lg = (REAL)log1p((double)ex);
    REAL ls = (l < 0 ? l : (REAL)0) - lg;
    REAL t = ((REAL)1 - (REAL)x) * l;
    return -(t - ls);
}

static REAL FN(loglik)(const REAL *pimg, const unsigned char *image, int HW) {
    double acc = 0;
    for (int i = 0; i < HW; ++i) acc += (double)FN(pixel_lp)(pimg[i], image[i] != 0);
    return (REAL)acc;
}

/* d lp / d p as autograd sees it: (x - sigmoid(l)) * (1/pt + 1/(1-pt)) inside the
 * clamp, 0 outside (clamp passes gradient at the bounds). */
static REAL FN(pixel_dlp)(REAL p, int x) {
    const REAL e = (REAL)1.1920928955078125e-07;
    if (p < e || p > (REAL)1 - e) return 0;
    REAL l = (REAL)log((double)p) - (REAL)log1p(-(double)p);
    REAL sg = (REAL)(1.0 / (1.0 + exp(-(double)l)));
    return ((REAL)x - sg) * ((REAL)1 / p + (REAL)1 / ((REAL)1 - p));
}

/* ---------------------------------------------------------------------------
 * Backward of render_core (autograd-equivalent; SURVEY.md App. A "Backward").
 *   gP        [H*W] cotangent on pimg
 *   g_pts     [npts][2] out: gradient w.r.t. the motor-space points (zeroed here)
 *   g_eps, g_sigma out (double)
 * ------------------------------------------------------------------------- */
static void FN(render_backward)(const bplo_params *ps, int nsub, const int *pt_off, const REAL *pts, REAL eps,
                                REAL sigma, const FN(render_tape) * tape, const REAL *gP, REAL *g_pts,
                                double *g_eps, double *g_sigma) {
    const int H = ps->H, W = ps->W, HW = H * W, fs = ps->fsize;
    REAL *G = (REAL *)malloc(sizeof(REAL) * HW), *T = (REAL *)malloc(sizeof(REAL) * HW);
    *g_eps = 0; *g_sigma = 0;
    /* rendering.py:227 */
    if (eps > 0) {
        double ge = 0;
        REAL a = (REAL)1 - eps;
        for (int i = 0; i < HW; ++i) {
            ge += (double)gP[i] * ((double)1 - 2 * (double)tape->I6[i]);
            G[i] = a * gP[i] - eps * gP[i];
        }
        *g_eps = ge;
    } else {
        memcpy(G, gP, sizeof(REAL) * HW);
    }
    /* :180 clamp(0,1): gradient passes where 0 <= I5 <= 1 */
    for (int i = 0; i < HW; ++i) if (!(tape->I5[i] >= 0 && tape->I5[i] <= 1)) G[i] = 0;
    if (sigma > 0) { /* :174-177 and general.py:196-207 */
        double *gK = (double *)calloc((size_t)fs * fs, sizeof(double));
        const REAL *ins[2] = {tape->I4, tape->I3};
        int p = fs / 2;
        for (int pass = 0; pass < 2; ++pass) {
            const REAL *In = ins[pass];
            for (int u = 0; u < fs; ++u)
                for (int v = 0; v < fs; ++v) {
                    double acc = 0;
                    for (int i = 0; i < H; ++i) {
                        int ii = i + u - p;
                        if (ii < 0 || ii >= H) continue;
                        for (int j = 0; j < W; ++j) {
                            int jj = j + v - p;
                            if (jj < 0 || jj >= W) continue;
                            acc += (double)G[i * W + j] * (double)In[ii * W + jj];
                        }
                    }
                    gK[u * fs + v] += acc;
                }
            /* adjoint of a zero-padded correlation with a symmetric kernel is itself */
            FN(corr2d)(G, T, H, W, tape->K, fs, fs);
            REAL *t = G; G = T; T = t;
        }
        /* K = E/sum(E), E = exp(-r2/(2 s^2)):  dK/ds = K (r2 - m)/s^3, m = sum K r2 */
        double m = 0, s3 = (double)sigma * sigma * sigma, gs = 0;
        for (int u = 0; u < fs; ++u)
            for (int v = 0; v < fs; ++v)
                m += (double)tape->K[u * fs + v] * (double)((u - p) * (u - p) + (v - p) * (v - p));
        for (int u = 0; u < fs; ++u)
            for (int v = 0; v < fs; ++v)
                gs += gK[u * fs + v] * (double)tape->K[u * fs + v] *
                      ((double)((u - p) * (u - p) + (v - p) * (v - p)) - m) / s3;
        *g_sigma = gs;
        free(gK);
    }
    /* :171 clamp(max=1) */
    for (int i = 0; i < HW; ++i) if (!(tape->I2[i] <= 1)) G[i] = 0;
    for (int it = 0; it < ps->ink_ncon; ++it) { /* :167-168 */
        FN(corr2d)(G, T, H, W, tape->Hk, 3, 3);
        REAL *t = G; G = T; T = t;
    }
    /* add_stroke backward per sub-stroke (rendering.py:93-125) */
    int maxn = 1, npts = pt_off[nsub];
    for (int s = 0; s < nsub; ++s) if (pt_off[s + 1] - pt_off[s] > maxn) maxn = pt_off[s + 1] - pt_off[s];
    int *idx = (int *)malloc(sizeof(int) * (size_t)maxn);
    REAL *rc = (REAL *)malloc(sizeof(REAL) * 2 * (size_t)maxn);
    REAL *ink = (REAL *)malloc(sizeof(REAL) * (size_t)maxn);
    REAL *gink = (REAL *)malloc(sizeof(REAL) * (size_t)maxn);
    REAL *grc = (REAL *)malloc(sizeof(REAL) * 2 * (size_t)maxn);
    REAL *ink0 = (REAL *)malloc(sizeof(REAL) * (size_t)maxn);
    for (int i = 0; i < 2 * npts; ++i) g_pts[i] = 0;
    for (int s = 0; s < nsub; ++s) {
        int n = pt_off[s + 1] - pt_off[s];
        if (n <= 0) continue;
        FN(substroke_info) info;
        FN(substroke_ink)(ps, pts + 2 * (size_t)pt_off[s], n, idx, rc, ink, &info);
        int m = info.n_in;
        if (m == 0) continue;
        for (int k = 0; k < m; ++k) { /* :111-125 */
            REAL r = rc[2 * k], c = rc[2 * k + 1];
            int rf = (int)R_FLOOR(r), cf = (int)R_FLOOR(c), rce = (int)R_CEIL(r), cce = (int)R_CEIL(c);
            REAL fr = r - (REAL)rf, fc = c - (REAL)cf, gr = (REAL)1 - fr, gc = (REAL)1 - fc;
            REAL G00 = G[rf * W + cf], G10 = G[rce * W + cf], G01 = G[rf * W + cce], G11 = G[rce * W + cce];
            gink[k] = gr * gc * G00 + fr * gc * G10 + gr * fc * G01 + fr * fc * G11;
            grc[2 * k] = ink[k] * (gc * (G10 - G00) + fc * (G11 - G01));
            grc[2 * k + 1] = ink[k] * (gr * (G01 - G00) + fr * (G11 - G10));
        }
        if (info.branch == 3 || info.branch == 4) {
            if (info.branch == 3) { /* ink' = ink * (ink_pp/sum): :106-107 */
                REAL sc = (REAL)ps->ink_pp / info.sumink;
                double dot = 0;
                for (int k = 0; k < m; ++k) { ink0[k] = ink[k] / sc; }
                /* recompute pre-scale ink exactly rather than dividing */
                for (int k = 0; k + 1 < m; ++k) {
                    REAL dr = rc[2 * (k + 1)] - rc[2 * k], dc = rc[2 * (k + 1) + 1] - rc[2 * k + 1];
                    REAL d = R_SQRT(dr * dr + dc * dc);
                    if (d > (REAL)ps->ink_max_dist) d = (REAL)ps->ink_max_dist;
                    ink0[k + 1] = ((REAL)ps->ink_pp * d) / (REAL)ps->ink_max_dist;
                }
                ink0[0] = ink0[1];
                for (int k = 0; k < m; ++k) dot += (double)gink[k] * (double)ink0[k];
                REAL gsum = (REAL)(-dot * (double)ps->ink_pp / ((double)info.sumink * (double)info.sumink));
                for (int k = 0; k < m; ++k) gink[k] = gink[k] * sc + gsum;
            }
            /* ink_ext = ink_pp*cat(d0, d)/max_dist ; d = min(|dq|, max_dist) : :96-99 */
            REAL f = (REAL)ps->ink_pp / (REAL)ps->ink_max_dist;
            for (int k = 0; k + 1 < m; ++k) {
                REAL gd = gink[k + 1] * f;
                if (k == 0) gd += gink[0] * f;
                REAL dr = rc[2 * (k + 1)] - rc[2 * k], dc = rc[2 * (k + 1) + 1] - rc[2 * k + 1];
                REAL d = R_SQRT(dr * dr + dc * dc);
                if (!(d <= (REAL)ps->ink_max_dist) || d == 0) continue; /* clamp mask; norm grad 0 at 0 */
                REAL ur = gd * (dr / d), uc = gd * (dc / d);
                grc[2 * (k + 1)] += ur; grc[2 * (k + 1) + 1] += uc;
                grc[2 * k] -= ur; grc[2 * k + 1] -= uc;
            }
        }
        /* un-compact and un-flip: r = -y, c = x */
        for (int k = 0; k < m; ++k) {
            size_t j = (size_t)pt_off[s] + idx[k];
            g_pts[2 * j] += grc[2 * k + 1];
            g_pts[2 * j + 1] += -grc[2 * k];
        }
    }
    free(idx); free(rc); free(ink); free(gink); free(grc); free(ink0); free(G); free(T);
}

/* ---------------------------------------------------------------------------
 * Whole path for one token = CharacterImageDist.get_pimg / score_image,
 * pybpl/model/image_dist.py:32-42,75-78.
 *   nstroke, nsub[nstroke]; ctrl [S][ncpt][2]; invscale [S]; position [nstroke][2]
 *   affine NULL or [4]; C = basis [neval][ncpt]
 *   image NULL or [H*W] {0,1}
 * outputs (any may be NULL): pimg[H*W], ll, off_page, motor [S][neval][2]
 * gradient outputs (all-or-none; g_ctrl non-NULL enables the backward; cotangent is
 *   d ll unless gP_in != NULL): g_ctrl [S][ncpt][2], g_invscale [S], g_position [nstroke][2],
 *   g_affine[4], g_sigma, g_eps
 * ------------------------------------------------------------------------- */
static void FN(token)(const bplo_params *ps, int ncpt, int neval, const REAL *C, int nstroke, const int *nsub,
                      const REAL *ctrl, const REAL *invscale, const REAL *position, const REAL *affine, REAL eps,
                      REAL sigma, const unsigned char *image, const REAL *gP_in, REAL *pimg_out, REAL *ll_out,
                      int *off_page_out, REAL *motor_out, REAL *g_ctrl, REAL *g_invscale, REAL *g_position,
                      REAL *g_affine, REAL *g_sigma, REAL *g_eps) {
    const int HW = ps->H * ps->W;
    int S = 0;
    for (int k = 0; k < nstroke; ++k) S += nsub[k];
    size_t npts = (size_t)S * neval;
    REAL *motor = (REAL *)malloc(sizeof(REAL) * 2 * npts);
    REAL *pre = NULL;
    int *pt_off = (int *)malloc(sizeof(int) * (size_t)(S + 1));
    for (int s = 0; s <= S; ++s) pt_off[s] = s * neval;
    int s0 = 0;
    for (int k = 0; k < nstroke; ++k) { /* image_dist.py:34 */
        FN(vanilla_to_motor)(C, neval, ncpt, nsub[k], ctrl + (size_t)s0 * 2 * ncpt, invscale + s0, position + 2 * k,
                             motor + (size_t)s0 * neval * 2, NULL);
        s0 += nsub[k];
    }
    REAL com[2] = {0, 0}, B[4] = {1, 1, 0, 0};
    if (affine) { /* image_dist.py:36-37 */
        if (g_ctrl) { pre = (REAL *)malloc(sizeof(REAL) * 2 * npts); memcpy(pre, motor, sizeof(REAL) * 2 * npts); }
        FN(apply_warp)(motor, npts, affine, com, B);
    }
    if (motor_out) memcpy(motor_out, motor, sizeof(REAL) * 2 * npts);
    REAL *pimg = (REAL *)malloc(sizeof(REAL) * HW);
    FN(render_tape) tape;
    FN(tape_alloc)(&tape, HW, ps->fsize);
    int off = FN(render_core)(ps, S, pt_off, motor, eps, sigma, pimg, &tape);
    if (off_page_out) *off_page_out = off;
    if (pimg_out) memcpy(pimg_out, pimg, sizeof(REAL) * HW);
    if (ll_out && image) *ll_out = FN(loglik)(pimg, image, HW);
    if (g_ctrl) {
        REAL *gP = (REAL *)malloc(sizeof(REAL) * HW);
        if (gP_in) memcpy(gP, gP_in, sizeof(REAL) * HW);
        else for (int i = 0; i < HW; ++i) gP[i] = FN(pixel_dlp)(pimg[i], image[i] != 0);
        REAL *gm = (REAL *)malloc(sizeof(REAL) * 2 * npts);
        double ge, gs;
        FN(render_backward)(ps, S, pt_off, motor, eps, sigma, &tape, gP, gm, &ge, &gs);
        if (g_eps) *g_eps = (REAL)ge;
        if (g_sigma) *g_sigma = (REAL)gs;
        if (affine) { /* affine.py:48-53 */
            double gB[4] = {0, 0, 0, 0};
            for (size_t i = 0; i < npts; ++i) {
                gB[0] += (double)gm[2 * i] * pre[2 * i]; gB[1] += (double)gm[2 * i + 1] * pre[2 * i + 1];
                gB[2] += gm[2 * i]; gB[3] += gm[2 * i + 1];
            }
            double gcx = -((double)affine[0] - 1) * gB[2] / (double)npts;
            double gcy = -((double)affine[1] - 1) * gB[3] / (double)npts;
            for (size_t i = 0; i < npts; ++i) {
                gm[2 * i] = (REAL)((double)gm[2 * i] * B[0] + gcx);
                gm[2 * i + 1] = (REAL)((double)gm[2 * i + 1] * B[1] + gcy);
            }
            if (g_affine) {
                g_affine[0] = (REAL)(gB[0] - (double)com[0] * gB[2]);
                g_affine[1] = (REAL)(gB[1] - (double)com[1] * gB[3]);
                g_affine[2] = (REAL)gB[2]; g_affine[3] = (REAL)gB[3];
            }
        } else if (g_affine) {
            g_affine[0] = g_affine[1] = g_affine[2] = g_affine[3] = 0;
        }
        /* part.py:316-326 backward: motor_i = traj_i - traj_i[0] + prev_i */
        int send = S;
        for (int k = nstroke - 1; k >= 0; --k) {
            int sbeg = send - nsub[k];
            double carry[2] = {0, 0}; /* gradient w.r.t. prev of the following sub-stroke */
            for (int i = send - 1; i >= sbeg; --i) {
                REAL *g = gm + (size_t)i * neval * 2;
                g[2 * (neval - 1)] += (REAL)carry[0]; g[2 * (neval - 1) + 1] += (REAL)carry[1];
                double tot[2] = {0, 0};
                for (int j = 0; j < neval; ++j) { tot[0] += g[2 * j]; tot[1] += g[2 * j + 1]; }
                g[0] -= (REAL)tot[0]; g[1] -= (REAL)tot[1]; /* through offset = traj[0] - prev */
                carry[0] = tot[0]; carry[1] = tot[1];
                /* traj = C @ (invscale * shapes): splines.py:136, part.py:318 */
                double gi = 0;
                for (int q = 0; q < ncpt; ++q)
                    for (int c = 0; c < 2; ++c) {
                        double acc = 0;
                        for (int j = 0; j < neval; ++j) acc += (double)C[(size_t)j * ncpt + q] * (double)g[2 * j + c];
                        size_t o = (size_t)i * 2 * ncpt + 2 * q + c;
                        g_ctrl[o] = (REAL)(acc * (double)invscale[i]);
                        gi += acc * (double)ctrl[o];
                    }
                g_invscale[i] = (REAL)gi;
            }
            g_position[2 * k] = (REAL)carry[0]; g_position[2 * k + 1] = (REAL)carry[1];
            send = sbeg;
        }
        free(gP); free(gm);
    }
    FN(tape_free)(&tape);
    free(pimg); free(motor); free(pt_off); free(pre);
}

#undef FN
#undef CAT
#undef CAT_
