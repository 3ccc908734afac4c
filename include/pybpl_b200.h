/*
 * pybpl_b200.h -- C ABI of libpybpl_b200.so, the B200 (sm_100a) implementation of
 * pyBPL's render+score hot path.
 *
 * The reference (rfeinman/pyBPL) is pure Python with no FFI of its own: the boundary is
 * a set of Python call sites.  Each entry point below names the reference
 * function(s) it replaces (paths relative to the reference root).  Every pointer marked
 * "dev" is a borrowed CUDA device pointer valid for the duration of the call on `stream`
 * (a cudaStream_t passed as void*); outputs are caller-allocated; the library never
 * allocates, frees or synchronises.  Every entry returns 0 on success, a negative
 * BPL_E* code for a bad argument, or a positive cudaError_t; bpl_last_error() gives text.
 * Host code binds it with ctypes (pybpl_b200/_lib.py); see INTEGRATION.md.
 */
#ifndef PYBPL_B200_H
#define PYBPL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPL_IMSIZE 105          /* pybpl/parameters.py:21  imsize = (105,105) */
#define BPL_FSIZE 11            /* pybpl/parameters.py:41  fsize */
#define BPL_IMG_WORDS 345       /* 11025 pixels, row-major, packed LSB-first: pixel i is bit (i & 31) of word (i >> 5) */
#define BPL_MAX_SUBSTROKES 128  /* sub-strokes per token handled by one CTA */
#define BPL_MAX_STROKES 32      /* strokes per token */
#define BPL_NCPT 5              /* control points per sub-stroke on the fused path (lib_data shape/mu: 1212x10) */
#define BPL_MAX_NEVAL 204       /* fused path: neval * BPL_NCPT <= 1024 basis entries held in shared memory */

#define BPL_EINVAL (-1)
#define BPL_ELIMIT (-2)
#define BPL_ENODEV (-3)

/* Rendering constants: pybpl/parameters.py:19-41 (class Parameters). */
typedef struct {
    int imsize;          /* must be 105 (any H x W: bpl_render_strokes_any, which ignores this field) */
    int fsize;           /* odd, 1..11 (the reference default is 11; smaller filters run as 11-tap passes with zero outer taps) */
    int ink_ncon;        /* :31 */
    float ink_pp;        /* :25 */
    float ink_max_dist;  /* :27 */
    float ink_a, ink_b;  /* :33, :35 */
    int hinton;          /* broaden_mode == 'Hinton' (:37) */
} bpl_params;

/* A ragged batch of character tokens in control-point form = what
 * CharacterImageDist.get_pimg reads from a CharacterToken (pybpl/model/image_dist.py:32-42,
 * pybpl/objects/concept.py:235-241, pybpl/objects/part.py:214-224). */
typedef struct {
    int n_tokens;               /* P */
    int ncpt;                   /* must be BPL_NCPT */
    int neval;                  /* points per sub-stroke (vanilla_to_motor neval, default 200) */
    const int *tok_stroke_off;  /* dev [P+1]   CSR: token -> strokes */
    const int *stroke_sub_off;  /* dev [K+1]   CSR: stroke -> sub-strokes */
    const float *ctrl;          /* dev [S][ncpt][2]  StrokeToken.shapes[:, :, i] per sub-stroke */
    const float *invscale;      /* dev [S]           StrokeToken.invscales */
    const float *position;      /* dev [K][2]        StrokeToken.position */
    const float *affine;        /* dev [P][4] or NULL  CharacterToken.affine */
    const float *eps;           /* dev [P]  CharacterToken.epsilon */
    const float *sigma;         /* dev [P]  CharacterToken.blur_sigma */
    const float *basis;         /* dev [neval][ncpt]  coefficient_mat(ncpt, neval) (pybpl/splines.py:72-93) */
    const int *order;           /* dev [P] or NULL: a permutation of 0..P-1, the order in which the persistent kernels
                                 * take the tokens (results do not depend on it; bpl_order_tokens gives the order that
                                 * ends a small launch on its cheapest tokens) */
    unsigned *cycles_out;       /* dev [P] or NULL: bpl_score_tokens_grad (team kernel) stores the SM cycles every token took.
                                 * A fit evaluates the same parses a hundred times: the cycles of one evaluation order the
                                 * next (bpl_order_tokens_by_cycles) -- the measured cost, not an estimate of it. */
} bpl_token_batch;

/* A ragged batch of characters given as raw motor-space trajectories = the `strokes`
 * argument of rendering.render_image (pybpl/rendering.py:185-229). */
typedef struct {
    int n_tokens;               /* P */
    const int *tok_sub_off;     /* dev [P+1]  CSR: token -> strokes */
    const int *sub_pt_off;      /* dev [S+1]  CSR: stroke -> points */
    const float *pts;           /* dev [N][2] motor space (x right, y negative down) */
    const float *eps;           /* dev [P] */
    const float *sigma;         /* dev [P] */
} bpl_stroke_batch;

/* Target images and where each token's score goes. */
typedef struct {
    const uint32_t *image_bits; /* dev [T][345] from bpl_pack_images_*; NULL => no scoring */
    const int *img_idx;         /* dev [P] image per token, or NULL => image 0 */
    int n_images;               /* T */
    int all_pairs;              /* 1: score every token against all T images (img_idx ignored); ll is [P][T] */
} bpl_targets;

/* Outputs of the forward (any may be NULL). */
typedef struct {
    float *ll;                  /* dev [P] (or [P][T] with all_pairs)  sum Bernoulli(pimg).log_prob(image) (image_dist.py:75-78) */
    uint8_t *off_page;          /* dev [P]  ink_off_page (rendering.py:85-88,219-220) */
    float *pimg;                /* dev [P][105][105] probability maps (render_image / get_pimg) */
} bpl_fwd_out;

/* Gradient outputs of the fused forward+backward.  The cotangent is g_ll[p] on ll[p]
 * (NULL => 1) or, when g_pimg != NULL, a full cotangent on pimg ([P][105][105]). */
typedef struct {
    const float *g_ll;          /* dev [P] or NULL */
    const float *g_pimg;        /* dev [P][105][105] or NULL */
    float *g_ctrl;              /* dev [S][ncpt][2]   (token batches) */
    float *g_invscale;          /* dev [S] */
    float *g_position;          /* dev [K][2] */
    float *g_affine;            /* dev [P][4] or NULL */
    float *g_pts;               /* dev [N][2]  (stroke batches; must be zero-filled by the caller) */
    float *g_eps;               /* dev [P] */
    float *g_sigma;             /* dev [P] */
} bpl_grads;

int bpl_version(void);
const char *bpl_last_error(void);
void bpl_default_params(bpl_params *ps);                    /* pybpl/parameters.py:19-41 */
/* device queries used to size launches (SM count, opt-in shared memory) */
int bpl_device_info(int *sm_count, int *smem_optin);

/* ---- splines (pybpl/splines.py) -------------------------------------------------- */
/* bspline_gen_s + coefficient_mat (splines.py:57-93), host arrays. */
int bpl_basis_table_host(int nland, int neval, float *C);
int bpl_basis_rows_host(int nland, int n, const float *s, float *C);
/* get_stk_from_bspline: X[b] = C @ Y[b] (splines.py:136); nspl splines share one table. */
int bpl_spline_eval(const float *C, int neval, int nland, int nspl, const float *Y, float *X, void *stream);
/* its backward: gY[b] = C^T @ gX[b] */
int bpl_spline_eval_bwd(const float *C, int neval, int nland, int nspl, const float *gX, float *gY, void *stream);
/* adaptive neval (splines.py:129-133): dist along the 200-point trajectory and
 * clamp(ceil(dist/grain), min_neval, max_neval); C200 is the [max_neval][nland] table. */
int bpl_spline_adaptive_neval(const float *C200, int max_neval, int nland, int nspl, const float *Y, float grain,
                              int min_neval, int *neval, float *dist, void *stream);
/* fit_bspline_to_traj (splines.py:141-171; torch.linalg.lstsq gels = Householder QR): the least-squares operator
 * P [nland][neval] of a coefficient matrix C [neval][nland] (host, double precision inside); then on the device
 * Y[b] = P @ X[b] is bpl_spline_eval(P, nland, neval, nspl, X, Y) and the squared residuals are: */
int bpl_lstsq_operator_host(int neval, int nland, const float *C, float *P);
int bpl_spline_residual(const float *C, int neval, int nland, int nspl, const float *X, const float *Y, float *resid,
                        void *stream);
/* vanilla_to_motor (pybpl/objects/part.py:290-328) for a batch of strokes:
 * stroke_sub_off [K+1]; motor [S][neval][2]; motor_spline [S][ncpt][2] (or NULL). */
int bpl_vanilla_to_motor(const float *C, int neval, int ncpt, int n_strokes, const int *stroke_sub_off,
                         const float *ctrl, const float *invscale, const float *position, float *motor,
                         float *motor_spline, void *stream);

/* ---- token prior: log P(token | type) -------------------------------------------------
 * CharacterTokenDist.score_token (pybpl/model/token_dist.py:264-287): per stroke the shapes and
 * invscales of StrokeTokenDist.score_part_token (:432-455), RelationTokenDist.score_relation_token
 * (:489-515) and score_location (:131-153) with RelationToken.get_attach_point
 * (pybpl/objects/relation.py:34-67), plus score_image_blur (:205-224).  SURVEY.md 8(f)1. */
#define BPL_REL_UNIHIST 0   /* relation categories (pybpl/objects/relation.py:11) */
#define BPL_REL_START 1
#define BPL_REL_END 2
#define BPL_REL_MID 3

/* Library constants the token distribution reads (token_dist.py:90-107,320-325,437-441). */
typedef struct {
    float sigma_shape;      /* lib.tokenvar['sigma_shape'] */
    float sigma_invscale;   /* lib.tokenvar['sigma_invscale'] */
    float sigma_attach;     /* lib.tokenvar['sigma_attach'] */
    float sigma_x, sigma_y; /* lib.rel['sigma_x'], lib.rel['sigma_y'] */
    float min_blur_sigma, max_blur_sigma;   /* Parameters (parameters.py:56-57) */
} bpl_token_prior_params;

/* Type-level parameters and relation tokens of the batch's tokens, aligned with the token batch's
 * sub-strokes ([S]) and strokes ([K]) (CharacterType: objects/concept.py:66-107, objects/part.py:100-160,
 * objects/relation.py:238-420). */
typedef struct {
    const float *ctrl_type;      /* dev [S][5][2]  StrokeType.shapes[:, :, i] */
    const float *invscale_type;  /* dev [S]        StrokeType.invscales */
    const int *rel_cat;          /* dev [K]  BPL_REL_* */
    const int *attach_ix;        /* dev [K]  stroke (within its token) attached to; ignored for unihist */
    const int *attach_subix;     /* dev [K]  sub-stroke of that stroke ('mid' only) */
    const float *gpos;           /* dev [K][2]  RelationIndependent.gpos ('unihist' only) */
    const float *eval_spot_type; /* dev [K]  RelationAttachAlong.eval_spot ('mid' only) */
    const float *eval_spot;      /* dev [K]  RelationToken.eval_spot_token ('mid' only) */
} bpl_token_type;

/* d ll / d (every leaf of CharacterToken.parameters() and CharacterType.parameters()); the first three are
 * required, the others may be NULL.  All are fully written by the call. */
typedef struct {
    float *g_ctrl, *g_invscale, *g_position;              /* dev [S][5][2], [S], [K][2] */
    float *g_eval_spot;                                   /* dev [K] */
    float *g_ctrl_type, *g_invscale_type;                 /* dev [S][5][2], [S] */
    float *g_gpos, *g_eval_spot_type;                     /* dev [K][2], [K] */
} bpl_token_prior_grads;

/* ll[P] <- log P(token | type); -inf for a non-positive invscale, an eval_spot outside [2, ncpt+1] or a blur
 * outside [min, max) (the reference raises on the last).  Reads b's CSR offsets, ctrl, invscale, position,
 * sigma and basis (rows 0 and neval-1).  g == NULL: forward only. */
int bpl_score_token_prior(const bpl_token_prior_params *c, const bpl_token_batch *b, const bpl_token_type *ty,
                          float *ll, const bpl_token_prior_grads *g, void *stream);

/* Sample tokens from P(token | type): CharacterTokenDist.sample_token (pybpl/model/token_dist.py:226-262; strokes in
 * order: sample_part_token :410-430, sample_relation_token :465-487, sample_location :109-129).  Reads b's CSR
 * offsets, neval and basis, and ty's type-level arrays (ty->eval_spot is ignored); writes ctrl [S][5][2],
 * invscale [S], position [K][2] and, unless NULL, eval_spot [K].  Token p draws from the Philox-4x32-10 stream
 * (seed, subsequence first_token + p), so a particle set can be produced in shards.  SURVEY.md 8(f)3 (token half). */
int bpl_sample_tokens(const bpl_token_prior_params *c, const bpl_token_batch *b, const bpl_token_type *ty,
                      unsigned long long seed, unsigned long long first_token, float *ctrl, float *invscale,
                      float *position, float *eval_spot, void *stream);

/* Sample character types from the prior: CharacterTypeDist.sample_type (pybpl/model/type_dist.py:56-96,187-207; per
 * stroke sample_part_type :480-505 and sample_relation_type :541-597).  SURVEY.md 8(f)3 (type half).
 * The library tables are device arrays prepared from the reference's Library (pybpl/library/library.py:30-72;
 * pybpl_b200/data/type_lib.npz): every categorical distribution as a cumulative sum whose last entry is 1. */
typedef struct {
    int n_prim;               /* primitives (1212) */
    int kmax;                 /* entries of pkappa = largest stroke count (10) */
    int nsub_max;             /* columns of pmat_nsub = largest sub-stroke count per stroke (10) */
    int n_hist, hist_bins;    /* position histograms (3: stroke 0, 1, >= 2), bins per side (25) */
    const float *pkappa;      /* dev [kmax]             cdf of the stroke count (lib.pkappa) */
    const float *pmat_nsub;   /* dev [kmax][nsub_max]   cdf of the sub-stroke count given the stroke count */
    const float *cdf_start;   /* dev [n_prim]           cdf of the first primitive (exp lib.logStart) */
    const float *cdf_T;       /* dev [n_prim][n_prim]   cdf of the next primitive given the previous one (lib.pT) */
    const float *shape_mu;    /* dev [n_prim][10]       lib.shape['mu'] */
    const float *shape_L;     /* dev [n_prim][10][10]   lower Cholesky factor of lib.shape['Sigma'] */
    const float *gamma_con;   /* dev [n_prim]           lib.scale['theta'][:, 0] */
    const float *gamma_rate;  /* dev [n_prim]           1 / lib.scale['theta'][:, 1] */
    const float *rel_mixprob; /* dev [4]                cdf of the relation category (lib.rel['mixprob']) */
    const float *sp_cdf;      /* dev [n_hist][bins*bins] cdf over bins, linear index = x bin * bins + y bin */
    const float *sp_xlab, *sp_ylab;   /* dev [bins+1]   bin edges */
} bpl_type_lib;

/* Outputs of the writing pass, at the offsets the caller scanned from the counting pass. */
typedef struct {
    int *stroke_nsub;         /* dev [K]   sub-strokes per stroke (scan it for stroke_sub_off) */
    int *subid;               /* dev [S]   StrokeType.ids */
    float *ctrl_type;         /* dev [S][5][2] */
    float *invscale_type;     /* dev [S] */
    int *rel_cat, *attach_ix, *attach_subix;   /* dev [K] */
    float *gpos;              /* dev [K][2] */
    float *eval_spot_type;    /* dev [K] */
} bpl_type_out;

/* Counting pass (tok_stroke_off == NULL): n_strokes[P], n_subs[P] <- strokes / sub-strokes of every type.
 * Writing pass: tok_stroke_off[P+1], tok_sub_off[P+1] = exclusive scans of those; draws the same types again (type p
 * uses the Philox stream (seed, first_type + p)) and fills *out. */
int bpl_sample_types(const bpl_type_lib *lib, int n_types, unsigned long long seed, unsigned long long first_type,
                     const int *tok_stroke_off, const int *tok_sub_off, int *n_strokes, int *n_subs,
                     const bpl_type_out *out, void *stream);

/* Log-probability tables of the type prior, for the type score (device arrays; pybpl_b200/data/type_lib.npz). */
typedef struct {
    const float *logp_kappa;   /* dev [kmax]             log lib.pkappa (normalised) */
    const float *logp_nsub;    /* dev [kmax][nsub_max]   log lib.pmat_nsub rows */
    const float *logp_start;   /* dev [n_prim]           lib.logStart (normalised) */
    const float *logp_T;       /* dev [n_prim][n_prim]   log lib.pT(prev)[next] */
    const float *logp_mix;     /* dev [4]                log lib.rel['mixprob'] */
    const float *sp_logp;      /* dev [n_hist][bins*bins] SpatialHist.logpYX, linear index = x bin * bins + y bin */
    float sp_log_binarea;      /* log rg_bin[0] + log rg_bin[1] (spatial_hist.py:163-164) */
} bpl_type_logtables;

/* ll[P] <- log P(type): CharacterTypeDist.score_type (pybpl/model/type_dist.py:98-125; score_k :163-185,
 * score_part_type :507-531, score_relation_type :599-650).  The types are given as for bpl_score_token_prior (CSR
 * offsets + bpl_token_type; eval_spot_type only enters through its support [2, ncpt + 1]) plus the primitive id of
 * every sub-stroke.  g_ctrl_type [S][5][2] / g_invscale_type [S]: gradients, or NULL.  -inf for a position outside
 * the histogram's range, an impossible count, an attachment index outside [0, i) / [0, nsub of the attached stroke)
 * or a type-level spline coordinate outside [2, ncpt + 1] (the reference's Categorical / Uniform log_prob raises or
 * returns -inf there, type_dist.py:636-647). */
int bpl_score_type_prior(const bpl_type_lib *lib, const bpl_type_logtables *tabs, int n_types,
                         const int *tok_stroke_off, const int *stroke_sub_off, const int *subid,
                         const bpl_token_type *ty, float *ll, float *g_ctrl_type, float *g_invscale_type, void *stream);

/* ---- images ------------------------------------------------------------------------ */
/* Pack T 105x105 images into bit rows.  `bad` (dev int, zero-filled by the caller) is set
 * to 1 if any pixel is not exactly 0 or 1 (torch's Bernoulli.log_prob support check). */
int bpl_pack_images_u8(const uint8_t *img, int T, uint32_t *bits, int *bad, void *stream);
int bpl_pack_images_f32(const float *img, int T, uint32_t *bits, int *bad, void *stream);

/* ---- render + score ------------------------------------------------------------------
 * Forward: CharacterImageDist.get_pimg + score_image (pybpl/model/image_dist.py:32-42,60-80)
 * = vanilla_to_motor (objects/part.py:290-328) -> apply_warp (util/affine.py:29-55) ->
 * render_image (rendering.py:185-229: add_stroke :58-127, broaden_and_blur :130-182) ->
 * Bernoulli log_prob.  One CTA per token; the image never leaves shared memory. */
int bpl_score_tokens(const bpl_params *ps, const bpl_token_batch *b, const bpl_targets *t, const bpl_fwd_out *out,
                     void *stream);
/* Particle sweep (BASELINE.json configs[3]): bpl_score_tokens with one target per token, plus the log-sum-exp partial
 * of the launch's scores produced by the same kernel (each CTA keeps a running (max, sum exp) of its tokens, the CTA
 * that finishes last combines them): lse[0] = max_p ll[p], lse[1] = sum_p exp(ll[p] - max) (dev float[2]; NaN scores
 * are skipped).  accumulate != 0 merges with the value already in lse, so a sweep scored in several launches still ends
 * with one pair.  lse_scratch: dev float[2 * BPL_LSE_SCRATCH_CTAS], private to the call while it runs.  out->ll may be
 * NULL when only the partial is wanted.  Semantics of the reduction: torch.logsumexp over
 * CharacterImageDist.score_image values (pybpl/model/image_dist.py:60-80); the reference has no batched counterpart. */
#define BPL_LSE_SCRATCH_CTAS 2048
int bpl_score_tokens_lse(const bpl_params *ps, const bpl_token_batch *b, const bpl_targets *t, const bpl_fwd_out *out,
                         float *lse, float *lse_scratch, int accumulate, void *stream);
/* order[P] <- the tokens by decreasing sub-stroke count (only the CSR offsets of b are read).  No reference
 * counterpart: the reference renders one token per Python call. */
int bpl_order_tokens(const bpl_token_batch *b, int *order, void *stream);
/* order[P] <- the tokens by decreasing ESTIMATED COST: the area of the box of the chained control polygons (a spline lies
 * in the hull of its control points, pybpl/splines.py:72-93; chaining as objects/part.py:316-326) grown by the stencils'
 * reach, and the sub-stroke count, in 256 cost classes.  keys[P]: scratch (receives the classes).  Reads the batch's
 * geometry as it is now; an order made from earlier parameter values stays a valid (merely less sharp) order. */
int bpl_order_tokens_by_cost(const bpl_params *ps, const bpl_token_batch *b, int *keys, int *order, void *stream);
/* order[P] <- the tokens by decreasing measured cost: cycles[P] as written through bpl_token_batch.cycles_out by an earlier
 * evaluation of the same batch (2 048-cycle classes).  keys[P]: scratch. */
int bpl_order_tokens_by_cycles(const unsigned *cycles, int n, int *keys, int *order, void *stream);
/* Same, plus the fused backward (autograd-equivalent gradients). */
int bpl_score_tokens_grad(const bpl_params *ps, const bpl_token_batch *b, const bpl_targets *t,
                          const bpl_fwd_out *out, const bpl_grads *g, void *stream);
/* out.g_* <- in.g_* scaled, token by token, by the cotangent g_ll[P] (the chain rule step torch's autograd performs on
 * the outputs of bpl_score_tokens_grad, which runs before the cotangent exists): g_ctrl, g_invscale, g_position, g_eps,
 * g_sigma and, when in->g_affine is set, g_affine.  One launch.  No reference counterpart (autograd does this there). */
int bpl_scale_token_grads(int P, int K, int S, const int *tok_stroke_off, const int *stroke_sub_off, const float *g_ll,
                          const bpl_grads *in, const bpl_grads *out, void *stream);
/* One Adam step (torch.optim.Adam's rule, no weight decay / amsgrad) over up to BPL_ADAM_MAX_TENSORS parameter tensors
 * in ONE launch, each then clamped to [lo, hi]: the optimiser step of the differentiable fit (fit_image,
 * pybpl/model/model.py:42-65, with the bounds of pybpl/parameters.py:54-63).  state: dev [2] floats, zero before the first
 * step -- state[0] counts the steps on the device (a captured CUDA graph advances it), state[1] is scratch. */
#define BPL_ADAM_MAX_TENSORS 8
typedef struct {
    int n_tensors, maximize;                       /* maximize: ascend (the gradient of sum ll is used as it is) */
    float lr, beta1, beta2, eps;
    float *state;
    float *param[BPL_ADAM_MAX_TENSORS];
    const float *grad[BPL_ADAM_MAX_TENSORS];
    float *exp_avg[BPL_ADAM_MAX_TENSORS], *exp_avg_sq[BPL_ADAM_MAX_TENSORS];
    long long n[BPL_ADAM_MAX_TENSORS];
    float lo[BPL_ADAM_MAX_TENSORS], hi[BPL_ADAM_MAX_TENSORS];   /* -inf / +inf: unbounded */
} bpl_adam_args;
int bpl_adam_step(const bpl_adam_args *a, void *stream);
/* rendering.render_image on raw trajectories (+ optional scoring). */
int bpl_render_strokes(const bpl_params *ps, const bpl_stroke_batch *b, const bpl_targets *t,
                       const bpl_fwd_out *out, void *stream);
int bpl_render_strokes_grad(const bpl_params *ps, const bpl_stroke_batch *b, const bpl_targets *t,
                            const bpl_fwd_out *out, const bpl_grads *g, void *stream);

/* rendering.render_image for ANY Parameters.imsize (pybpl/parameters.py:21; rendering.py:185-229 with ps.imsize = (H, W):
 * the map is H x W, rows bounded by imsize[0], columns by imsize[1], rendering.py:27-31,214), with the Bernoulli score
 * of model/image_dist.py:75-78 and the fused backward.  ps->imsize is ignored here.  The fused kernels above are built
 * for the reference's default 105 x 105; these run the same pipeline with run-time geometry, one CTA per character,
 * on dense canvases in a caller-provided workspace.  Outputs / cotangents as bpl_render_strokes[_grad] with
 * [P][H][W] maps; target images are plain float maps (the caller checks they are exactly 0 / 1). */
typedef struct {
    int H, W;                   /* imsize[0] (rows), imsize[1] (columns); H * W <= 2^24 */
    const float *images;        /* dev [T][H][W], every value exactly 0 or 1; NULL => no scoring */
    const int *img_idx;         /* dev [P] image per character, or NULL => image 0 */
    float *workspace;           /* dev, at least bpl_render_any_workspace(...) floats, private to the call while it runs */
    long long workspace_floats;
} bpl_any_geom;
long long bpl_render_any_workspace(int H, int W, int n_tokens, int with_grad);      /* floats; < 0: bad argument */
int bpl_render_strokes_any(const bpl_params *ps, const bpl_any_geom *geo, const bpl_stroke_batch *b, const bpl_fwd_out *out,
                           void *stream);
int bpl_render_strokes_any_grad(const bpl_params *ps, const bpl_any_geom *geo, const bpl_stroke_batch *b,
                                const bpl_fwd_out *out, const bpl_grads *g, void *stream);

/* Inspection: the path's integer work per point, computed by the same device code as the fused
 * kernels: motor [S][neval][2] (post-warp motor space = torch.cat(motor) in image_dist.py:34-38),
 * mask_out [S][neval] (check_bounds, rendering.py:27-31), floor/ceil [S][neval][2] (row, col;
 * rendering.py:113-116), n_in [S] surviving points per sub-stroke.  Any output may be NULL. */
int bpl_token_points(const bpl_params *ps, const bpl_token_batch *b, float *motor, uint8_t *mask_out, int *fl, int *ce,
                     int *n_in, void *stream);

/* ---- particle sweep reduction (per-GPU partial of the final log-sum-exp) ------------- */
/* out[0] = max_p ll[p], out[1] = sum_p exp(ll[p] - max)   (dev float[2]); one launch. */
int bpl_logsumexp_partial(const float *ll, int n, float *out2, void *stream);
/* out[0] = log sum_i parts[2i+1] * exp(parts[2i]) over n (max, sum-exp) pairs (dev), e.g. the ranks' partials after
 * the all-gather: the sweep's final log-sum-exp without leaving the device. */
int bpl_logsumexp_combine(const float *parts, int n, float *out, void *stream);

/* Roofline probe: `blocks` CTAs x 1024 threads x `iters` x 8 FFMA (2 flop each); scratch = 1 dev float. */
int bpl_fp32_probe(float *scratch, int iters, int blocks, void *stream);
/* Shared-memory bandwidth probe: blocks x 1024 threads, each streaming iters x 4 conflict-free 128-bit loads
 * (= blocks * 1024 * iters * 64 bytes) -- the measured peak the kernels' shared-memory traffic is put against. */
int bpl_smem_probe(float *scratch, int iters, int blocks, void *stream);

/* number of kernel launches issued by this library in this process (bench.py gpu_launches) */
long long bpl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
