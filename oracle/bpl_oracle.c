/*
 * bpl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of pyBPL's render+score hot path, used as the
 * parity checker for the CUDA kernels in pybpl_b200/csrc and as the CPU
 * baseline leg of bench.py.  See bpl_oracle_impl.h for the algorithm and the
 * reference file:line each function follows.  The product (pybpl_b200/) never
 * links or calls this library.
 *
 * Build: oracle/build.sh  ->  oracle/_build/libbpl_oracle.so
 *   gcc -O2 -ffp-contract=off (no fused contraction: the fp32 build must
 *   round after every operation exactly where the reference's torch ops do).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "bpl_oracle.h"

/* ---- float instantiation ---- */
#define REAL float
#define SUF _f32
#define R_FMA(a, b, c) fmaf((a), (b), (c))
#define R_SQRT(x) sqrtf(x)
#define R_FLOOR(x) floorf(x)
#define R_CEIL(x) ceilf(x)
#define R_EXP(x) ((float)exp((double)(x)))
#include "bpl_oracle_impl.h"
#undef REAL
#undef SUF
#undef R_FMA
#undef R_SQRT
#undef R_FLOOR
#undef R_CEIL
#undef R_EXP

/* ---- double instantiation ---- */
#define REAL double
#define SUF _f64
#define R_FMA(a, b, c) fma((a), (b), (c))
#define R_SQRT(x) sqrt(x)
#define R_FLOOR(x) floor(x)
#define R_CEIL(x) ceil(x)
#define R_EXP(x) exp(x)
#include "bpl_oracle_impl.h"
#undef REAL
#undef SUF
#undef R_FMA
#undef R_SQRT
#undef R_FLOOR
#undef R_CEIL
#undef R_EXP

int bplo_version(void) { return 1; }

void bplo_default_params(bplo_params *ps) {
    /* pybpl/parameters.py:19-41 */
    ps->H = 105; ps->W = 105;
    ps->ink_pp = 2.0f; ps->ink_max_dist = 2.0f;
    ps->ink_ncon = 2; ps->ink_a = 0.5f; ps->ink_b = 6.0f;
    ps->hinton = 0; ps->fsize = 11;
}

#define EXPORTS(REAL, SUF)                                                                                         \
    void bplo_basis_table##SUF(int nland, int neval, REAL *C) { basis_table##SUF(nland, neval, C); }              \
    void bplo_basis_rows##SUF(int nland, int n, const REAL *s, REAL *C) {                                          \
        for (int j = 0; j < n; ++j) basis_row##SUF(nland, s[j], C + (size_t)j * nland);                            \
    }                                                                                                              \
    void bplo_spline_eval##SUF(const REAL *C, int neval, int nland, int nspl, const REAL *Y, REAL *X) {            \
        for (int b = 0; b < nspl; ++b)                                                                             \
            spline_eval##SUF(C, neval, nland, Y + (size_t)b * 2 * nland, X + (size_t)b * 2 * neval);               \
    }                                                                                                              \
    void bplo_adaptive_neval##SUF(const REAL *C200, int nland, int nspl, const REAL *Y, int *neval, REAL *dist) {  \
        for (int b = 0; b < nspl; ++b)                                                                             \
            neval[b] = adaptive_neval##SUF(C200, nland, Y + (size_t)b * 2 * nland, dist ? dist + b : NULL);        \
    }                                                                                                              \
    void bplo_lstsq##SUF(const REAL *C, int m, int n, int nprob, const REAL *X, REAL *Y, REAL *resid) {            \
        for (int b = 0; b < nprob; ++b)                                                                            \
            lstsq##SUF(C, m, n, X + (size_t)b * m * 2, Y + (size_t)b * n * 2, resid + 2 * b);                      \
    }                                                                                                              \
    void bplo_vanilla_to_motor##SUF(const REAL *C, int neval, int ncpt, int nsub, const REAL *ctrl,                \
                                    const REAL *invscale, const REAL *pos, REAL *motor, REAL *mspline) {           \
        vanilla_to_motor##SUF(C, neval, ncpt, nsub, ctrl, invscale, pos, motor, mspline);                          \
    }                                                                                                              \
    int bplo_render##SUF(const bplo_params *ps, int nsub, const int *pt_off, const REAL *pts, REAL eps,            \
                         REAL sigma, REAL *pimg) {                                                                 \
        return render_core##SUF(ps, nsub, pt_off, pts, eps, sigma, pimg, NULL);                                    \
    }                                                                                                              \
    REAL bplo_loglik##SUF(const REAL *pimg, const unsigned char *image, int HW) {                                  \
        return loglik##SUF(pimg, image, HW);                                                                       \
    }                                                                                                              \
    /* per-point integer work of one sub-stroke: mask_out[n], floor/ceil [n][2] (row, col) */                      \
    void bplo_point_indices##SUF(const bplo_params *ps, int n, const REAL *pts, unsigned char *mask_out,           \
                                 int *fl, int *ce) {                                                               \
        for (int j = 0; j < n; ++j) {                                                                              \
            REAL r = pts[2 * j + 1] * (REAL)-1, c = pts[2 * j];                                                    \
            mask_out[j] = (floor((double)r) < 0) | (ceil((double)r) >= ps->H) | (floor((double)c) < 0) |           \
                          (ceil((double)c) >= ps->W);                                                              \
            fl[2 * j] = (int)floor((double)r); fl[2 * j + 1] = (int)floor((double)c);                              \
            ce[2 * j] = (int)ceil((double)r); ce[2 * j + 1] = (int)ceil((double)c);                                \
        }                                                                                                          \
    }                                                                                                              \
    /* render + (ll | generic cotangent) + gradients w.r.t. raw points, eps, sigma */                              \
    int bplo_render_grad##SUF(const bplo_params *ps, int nsub, const int *pt_off, const REAL *pts, REAL eps,       \
                              REAL sigma, const unsigned char *image, const REAL *gP_in, REAL *pimg, REAL *ll,     \
                              REAL *g_pts, REAL *g_eps, REAL *g_sigma) {                                           \
        int HW = ps->H * ps->W;                                                                                    \
        render_tape##SUF tape;                                                                                     \
        tape_alloc##SUF(&tape, HW, ps->fsize);                                                                     \
        REAL *P = (REAL *)malloc(sizeof(REAL) * HW), *gP = (REAL *)malloc(sizeof(REAL) * HW);                      \
        int off = render_core##SUF(ps, nsub, pt_off, pts, eps, sigma, P, &tape);                                   \
        if (pimg) memcpy(pimg, P, sizeof(REAL) * HW);                                                              \
        if (ll && image) *ll = loglik##SUF(P, image, HW);                                                          \
        if (gP_in) memcpy(gP, gP_in, sizeof(REAL) * HW);                                                           \
        else for (int i = 0; i < HW; ++i) gP[i] = pixel_dlp##SUF(P[i], image[i] != 0);                             \
        double ge, gs;                                                                                             \
        render_backward##SUF(ps, nsub, pt_off, pts, eps, sigma, &tape, gP, g_pts, &ge, &gs);                       \
        *g_eps = (REAL)ge; *g_sigma = (REAL)gs;                                                                    \
        tape_free##SUF(&tape); free(P); free(gP);                                                                  \
        return off;                                                                                                \
    }                                                                                                              \
    void bplo_token##SUF(const bplo_params *ps, int ncpt, int neval, const REAL *C, int nstroke, const int *nsub,  \
                         const REAL *ctrl, const REAL *invscale, const REAL *position, const REAL *affine,         \
                         REAL eps, REAL sigma, const unsigned char *image, const REAL *gP_in, REAL *pimg,          \
                         REAL *ll, int *off_page, REAL *motor, REAL *g_ctrl, REAL *g_invscale, REAL *g_position,   \
                         REAL *g_affine, REAL *g_sigma, REAL *g_eps) {                                             \
        token##SUF(ps, ncpt, neval, C, nstroke, nsub, ctrl, invscale, position, affine, eps, sigma, image, gP_in,  \
                   pimg, ll, off_page, motor, g_ctrl, g_invscale, g_position, g_affine, g_sigma, g_eps);           \
    }                                                                                                              \
    /* Packed ragged batch (same layout as include/pybpl_b200.h): P tokens, CSR offsets                           \
     * tok_stroke_off[P+1] -> strokes, stroke_sub_off[K_tot+1] -> sub-strokes.                                     \
     * images [T][HW] u8, img_idx[P] (NULL => image 0).  grads optional (g_ctrl NULL => forward only).             \
     * Threads: OpenMP over tokens when built with -fopenmp. */                                                    \
    void bplo_token_batch##SUF(const bplo_params *ps, int ncpt, int neval, const REAL *C, int P,                   \
                               const int *tok_stroke_off, const int *stroke_sub_off, const REAL *ctrl,             \
                               const REAL *invscale, const REAL *position, const REAL *affine, const REAL *eps,    \
                               const REAL *sigma, const unsigned char *images, const int *img_idx, REAL *ll,       \
                               unsigned char *off_page, REAL *pimg, REAL *g_ctrl, REAL *g_invscale,                \
                               REAL *g_position, REAL *g_affine, REAL *g_sigma, REAL *g_eps) {                     \
        const int HW = ps->H * ps->W;                                                                              \
        _Pragma("omp parallel for schedule(dynamic, 4)")                                                           \
        for (int p = 0; p < P; ++p) {                                                                              \
            int k0 = tok_stroke_off[p], k1 = tok_stroke_off[p + 1], K = k1 - k0;                                   \
            int s0 = stroke_sub_off[k0];                                                                           \
            int nsub[64];                                                                                          \
            for (int k = 0; k < K; ++k) nsub[k] = stroke_sub_off[k0 + k + 1] - stroke_sub_off[k0 + k];             \
            const unsigned char *im = images ? images + (size_t)(img_idx ? img_idx[p] : 0) * HW : NULL;            \
            int off = 0;                                                                                           \
            REAL l = 0, gs = 0, ge = 0, ga[4];                                                                     \
            token##SUF(ps, ncpt, neval, C, K, nsub, ctrl + (size_t)s0 * 2 * ncpt, invscale + s0,                   \
                       position + 2 * (size_t)k0, affine ? affine + 4 * (size_t)p : NULL, eps[p], sigma[p], im,    \
                       NULL, pimg ? pimg + (size_t)p * HW : NULL, &l, &off, NULL,                                  \
                       g_ctrl ? g_ctrl + (size_t)s0 * 2 * ncpt : NULL, g_ctrl ? g_invscale + s0 : NULL,            \
                       g_ctrl ? g_position + 2 * (size_t)k0 : NULL, ga, &gs, &ge);                                 \
            if (ll) ll[p] = l;                                                                                     \
            if (off_page) off_page[p] = (unsigned char)off;                                                        \
            if (g_ctrl) {                                                                                          \
                if (g_sigma) g_sigma[p] = gs;                                                                      \
                if (g_eps) g_eps[p] = ge;                                                                          \
                if (g_affine) for (int q = 0; q < 4; ++q) g_affine[4 * (size_t)p + q] = ga[q];                     \
            }                                                                                                      \
        }                                                                                                          \
    }

EXPORTS(float, _f32)
EXPORTS(double, _f64)

/* Number of OpenMP threads the batch entry points use (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU
 * arm of the bench sets the host's core count back explicitly).  Returns the count in effect. */
#ifdef _OPENMP
#include <omp.h>
int bplo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int bplo_set_threads(int n) { (void)n; return 1; }
#endif
