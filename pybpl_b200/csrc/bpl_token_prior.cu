// bpl_token_prior.cu -- log P(token | type) for a ragged batch of character tokens, with its gradient.
//
// Replaces CharacterTokenDist.score_token (pybpl/model/token_dist.py:264-287), i.e. per stroke
//   StrokeTokenDist.score_part_token        token_dist.py:432-455  (shapes :334-354, invscales :379-408)
//   RelationTokenDist.score_relation_token  token_dist.py:489-515  (score_eval_spot_token :542-572)
//   CharacterTokenDist.score_location       token_dist.py:131-153  (get_attach_point, objects/relation.py:34-67;
//                                                                   trajectories objects/part.py:290-328)
// plus score_image_blur (:205-224); score_affine and score_image_noise are 0 in the reference (:169-170,186-187).
// It is the other per-particle term of the fit_image objective (model/model.py:59-62), SURVEY.md 8(f)1.
//
// One warp per token: lanes over the control-point coordinates (coalesced), then over the sub-strokes, then
// over the strokes.  Every gradient element is first written by its owner (direct terms), then the attach-point
// terms -- a stroke's location depends on an EARLIER stroke's trajectory end points or on a point along one of
// its sub-strokes -- are added with atomics.  A few hundred flops per token: the kernel is bound by reading the
// token (44 S + 20 K bytes of token and as much of type parameters), i.e. by HBM.
#include <cuda_runtime.h>
#include <math.h>

#include "bpl_common.cuh"
#include "bpl_token_common.cuh"

namespace bpl {
void set_error(const char *fmt, const char *detail);
void count_launch(int n);
int sm_count();

namespace {

constexpr float LOG_SQRT_2PI = 0.91893853320467274178f;
constexpr float INV_SQRT_2 = 0.70710678118654752440f;
constexpr float INV_SQRT_2PI = 0.39894228040143267794f;

struct TPArgs {
    int n_tokens, neval;
    const int *tok_stroke_off, *stroke_sub_off;
    const float *ctrl, *invscale, *position, *sigma, *basis;
    bpl_token_type ty;
    bpl_token_prior_params c;
    float *ll;
    bpl_token_prior_grads g;
    int want_grad;
};

// torch.distributions.Normal: log_prob and cdf in float32 (torch/distributions/normal.py)
__device__ __forceinline__ float normal_logpdf(float diff, float sigma) {
    return -(diff * diff) / (2.0f * sigma * sigma) - logf(sigma) - LOG_SQRT_2PI;
}
__device__ __forceinline__ float normal_cdf(float x, float mu, float sigma) {
    return 0.5f * (1.0f + erff((x - mu) * (1.0f / sigma) * INV_SQRT_2));
}
__device__ __forceinline__ float std_pdf(float z) { return INV_SQRT_2PI * expf(-0.5f * z * z); }

#ifndef BPL_TP_BLOCKS
#define BPL_TP_BLOCKS 8      // 32 registers, 64 warps per SM: the kernel is latency-bound (dependent offset -> data loads); measured 1 / 4 / 6 / 8 blocks: 412 / 616 / 656 / 768 M tokens/s
#endif
__global__ void __launch_bounds__(256, BPL_TP_BLOCKS) token_prior_kernel(const TPArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float ss = a.c.sigma_shape, si = a.c.sigma_invscale, sa = a.c.sigma_attach;
    const float lb = 2.0f, ub = (float)(NCPT + 1);                    // bspline_gen_s (splines.py:57-69)
    float C0[NCPT], C1[NCPT];                                           // first / last row of coefficient_mat(ncpt, neval)
#pragma unroll
    for (int k = 0; k < NCPT; ++k) { C0[k] = a.basis[k]; C1[k] = a.basis[(a.neval - 1) * NCPT + k]; }
    const bool wg = a.want_grad != 0;

    for (int p = warp; p < a.n_tokens; p += nwarps) {
        const int k0 = a.tok_stroke_off[p], k1 = a.tok_stroke_off[p + 1];
        const int s0 = a.stroke_sub_off[k0], s1 = a.stroke_sub_off[k1];
        float acc = 0.0f;
        // ---- shapes: Normal(shapes_type, sigma_shape).log_prob, every coordinate (token_dist.py:334-354) ----
        {
            const float c_const = -logf(ss) - LOG_SQRT_2PI, inv2 = 1.0f / (ss * ss);
            const int e0 = s0 * NCPT * 2, e1 = s1 * NCPT * 2;
            for (int e = e0 + lane; e < e1; e += 32) {
                const float diff = a.ctrl[e] - a.ty.ctrl_type[e];
                acc += -(diff * diff) / (2.0f * ss * ss) + c_const;
                if (wg) { a.g.g_ctrl[e] = -diff * inv2; if (a.g.g_ctrl_type) a.g.g_ctrl_type[e] = diff * inv2; }
            }
        }
        // ---- invscales: Normal restricted to positive values (token_dist.py:379-408) ----
        for (int s = s0 + lane; s < s1; s += 32) {
            const float tk = a.invscale[s], tyv = a.ty.invscale_type[s];
            const float diff = tk - tyv;
            const float p_above = 1.0f - normal_cdf(0.0f, tyv, si);
            float l = normal_logpdf(diff, si) - logf(p_above);
            if (tk <= 0.0f) l = -INFINITY;
            acc += l;
            if (wg) {
                a.g.g_invscale[s] = -diff / (si * si);
                if (a.g.g_invscale_type) {
                    const float z = tyv / si;
                    const float Phi = 0.5f * (1.0f + erff(z * INV_SQRT_2));
                    a.g.g_invscale_type[s] = diff / (si * si) - std_pdf(z) / (si * Phi);
                }
            }
        }
        // ---- owners initialise the per-stroke gradients before any attach-point term is added ----
        if (wg) {
            for (int k = k0 + lane; k < k1; k += 32) {
                a.g.g_position[2 * k] = 0.0f; a.g.g_position[2 * k + 1] = 0.0f;
                if (a.g.g_eval_spot) a.g.g_eval_spot[k] = 0.0f;
                if (a.g.g_eval_spot_type) a.g.g_eval_spot_type[k] = 0.0f;
                if (a.g.g_gpos) { a.g.g_gpos[2 * k] = 0.0f; a.g.g_gpos[2 * k + 1] = 0.0f; }
            }
            __threadfence_block();
        }
        __syncwarp();
        // ---- per stroke: relation token + location (token_dist.py:489-515, 131-153) ----
        for (int k = k0 + lane; k < k1; k += 32) {
            const int cat = a.ty.rel_cat[k];
            float cu[NCPT], dcu[NCPT], csum = 1.0f;
            const float u = a.ty.eval_spot ? a.ty.eval_spot[k] : 0.0f;
            if (cat == BPL_REL_MID) {
                const float uty = a.ty.eval_spot_type[k];
                if (u < lb || u > ub) {
                    acc += -INFINITY;                                                      // :563-564
                } else {
                    const float de = u - uty;
                    const float p_within = normal_cdf(ub, uty, sa) - normal_cdf(lb, uty, sa);   // :568-570
                    acc += normal_logpdf(de, sa) - logf(p_within);
                    if (wg) {
                        if (a.g.g_eval_spot) atomicAdd(a.g.g_eval_spot + k, -de / (sa * sa));
                        if (a.g.g_eval_spot_type) {
                            const float zu = (ub - uty) / sa, zl = (lb - uty) / sa;
                            a.g.g_eval_spot_type[k] = de / (sa * sa) - (std_pdf(zl) - std_pdf(zu)) / (sa * p_within);
                        }
                    }
                }
                basis_row_dev(u, cu, dcu, csum);
            }
            // attach point (relation.py:49-65)
            float bx, by;
            int j = 0, sj0 = 0, sub = 0;
            float mx = 0.0f, my = 0.0f;          // (mid) sum_k dc_k (Y_k - off): d base / d u
            if (cat == BPL_REL_UNIHIST) {
                bx = a.ty.gpos[2 * k]; by = a.ty.gpos[2 * k + 1];
            } else {
                j = k0 + a.ty.attach_ix[k];
                sj0 = a.stroke_sub_off[j];
                const int sj1 = a.stroke_sub_off[j + 1];
                const float pjx = a.position[2 * j], pjy = a.position[2 * j + 1];
                if (cat == BPL_REL_START) {                       // motor[0][0] = traj[0] - (traj[0] - position) (part.py:321-322)
                    const float inv = a.invscale[sj0];
                    const float *cp = a.ctrl + (size_t)sj0 * NCPT * 2;
                    float ax = 0.0f, ay = 0.0f;
#pragma unroll
                    for (int q = 0; q < NCPT; ++q) { ax = fmaf(C0[q], inv * cp[2 * q], ax); ay = fmaf(C0[q], inv * cp[2 * q + 1], ay); }
                    bx = ax - (ax - pjx); by = ay - (ay - pjy);
                } else {
                    sub = (cat == BPL_REL_MID) ? sj0 + a.ty.attach_subix[k] : sj1;      // chain up to (not including) sub
                    float ex = pjx, ey = pjy;                      // end point of the previous sub-stroke (part.py:316-326)
                    for (int s = sj0; s < sub; ++s) {
                        const float inv = a.invscale[s];
                        const float *cp = a.ctrl + (size_t)s * NCPT * 2;
                        float ax = 0.0f, ay = 0.0f, zx = 0.0f, zy = 0.0f;
#pragma unroll
                        for (int q = 0; q < NCPT; ++q) {
                            const float yx = inv * cp[2 * q], yy = inv * cp[2 * q + 1];
                            ax = fmaf(C0[q], yx, ax); ay = fmaf(C0[q], yy, ay);
                            zx = fmaf(C1[q], yx, zx); zy = fmaf(C1[q], yy, zy);
                        }
                        ex = zx - (ax - ex); ey = zy - (ay - ey);
                    }
                    if (cat == BPL_REL_END) {
                        bx = ex; by = ey;                          // motor[-1][-1]
                    } else {                                      // C(u) @ motor_spline[:, :, subix], motor_spline = Y - offset (part.py:323)
                        const float inv = a.invscale[sub];
                        const float *cp = a.ctrl + (size_t)sub * NCPT * 2;
                        float ax = 0.0f, ay = 0.0f;
#pragma unroll
                        for (int q = 0; q < NCPT; ++q) { ax = fmaf(C0[q], inv * cp[2 * q], ax); ay = fmaf(C0[q], inv * cp[2 * q + 1], ay); }
                        const float ox = ax - ex, oy = ay - ey;
                        bx = 0.0f; by = 0.0f;
#pragma unroll
                        for (int q = 0; q < NCPT; ++q) {
                            const float yx = inv * cp[2 * q] - ox, yy = inv * cp[2 * q + 1] - oy;
                            bx = fmaf(cu[q], yx, bx); by = fmaf(cu[q], yy, by);
                            mx = fmaf(dcu[q], yx, mx); my = fmaf(dcu[q], yy, my);
                        }
                    }
                }
            }
            const float rx = a.position[2 * k] - bx, ry = a.position[2 * k + 1] - by;
            acc += normal_logpdf(rx, a.c.sigma_x) + normal_logpdf(ry, a.c.sigma_y);     // Independent(Normal(0, (sigma_x, sigma_y)))
            if (wg) {
                const float gbx = rx / (a.c.sigma_x * a.c.sigma_x), gby = ry / (a.c.sigma_y * a.c.sigma_y);   // d ll / d base
                atomicAdd(a.g.g_position + 2 * k, -gbx); atomicAdd(a.g.g_position + 2 * k + 1, -gby);
                if (cat == BPL_REL_UNIHIST) {
                    if (a.g.g_gpos) { a.g.g_gpos[2 * k] = gbx; a.g.g_gpos[2 * k + 1] = gby; }
                } else if (cat == BPL_REL_START) {                 // d/d traj[0] cancels
                    atomicAdd(a.g.g_position + 2 * j, gbx); atomicAdd(a.g.g_position + 2 * j + 1, gby);
                } else {
                    const float w = (cat == BPL_REL_END) ? 1.0f : csum;      // weight of the chained end point
                    for (int s = sj0; s < sub; ++s) {              // end_s = end_{s-1} + (C1 - C0) @ Y_s
                        const float inv = a.invscale[s];
                        const float *cp = a.ctrl + (size_t)s * NCPT * 2;
                        float gi = 0.0f;
#pragma unroll
                        for (int q = 0; q < NCPT; ++q) {
                            const float wq = (C1[q] - C0[q]) * w;
                            atomicAdd(a.g.g_ctrl + ((size_t)s * NCPT + q) * 2, inv * wq * gbx);
                            atomicAdd(a.g.g_ctrl + ((size_t)s * NCPT + q) * 2 + 1, inv * wq * gby);
                            gi = fmaf(cp[2 * q], wq * gbx, fmaf(cp[2 * q + 1], wq * gby, gi));
                        }
                        atomicAdd(a.g.g_invscale + s, gi);
                    }
                    atomicAdd(a.g.g_position + 2 * j, w * gbx); atomicAdd(a.g.g_position + 2 * j + 1, w * gby);
                    if (cat == BPL_REL_MID) {
                        const float inv = a.invscale[sub];
                        const float *cp = a.ctrl + (size_t)sub * NCPT * 2;
                        float gi = 0.0f;
#pragma unroll
                        for (int q = 0; q < NCPT; ++q) {
                            const float wq = cu[q] - csum * C0[q];
                            atomicAdd(a.g.g_ctrl + ((size_t)sub * NCPT + q) * 2, inv * wq * gbx);
                            atomicAdd(a.g.g_ctrl + ((size_t)sub * NCPT + q) * 2 + 1, inv * wq * gby);
                            gi = fmaf(cp[2 * q], wq * gbx, fmaf(cp[2 * q + 1], wq * gby, gi));
                        }
                        atomicAdd(a.g.g_invscale + sub, gi);
                        if (a.g.g_eval_spot) atomicAdd(a.g.g_eval_spot + k, mx * gbx + my * gby);
                    }
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            // score_image_blur: Uniform(lb, ub).log_prob (token_dist.py:205-224); outside the support the reference
            // raises (validate_args), here the score is -inf.  score_affine and score_image_noise are 0.
            const float sg = a.sigma[p];
            const float blur = (sg >= a.c.min_blur_sigma && sg < a.c.max_blur_sigma) ? -logf(a.c.max_blur_sigma - a.c.min_blur_sigma)
                                                                                      : -INFINITY;
            a.ll[p] = acc + blur;
        }
    }
}

}  // namespace
}  // namespace bpl

using namespace bpl;

extern "C" int bpl_score_token_prior(const bpl_token_prior_params *c, const bpl_token_batch *b, const bpl_token_type *ty,
                                     float *ll, const bpl_token_prior_grads *g, void *stream) {
    if (!c || !b || !ty || !ll) { set_error("%s", "bpl_score_token_prior: NULL argument"); return BPL_EINVAL; }
    if (b->n_tokens <= 0) return 0;
    if (b->ncpt != NCPT) { set_error("%s", "bpl_score_token_prior is built for ncpt=5"); return BPL_ELIMIT; }
    if (!b->tok_stroke_off || !b->stroke_sub_off || !b->sigma || !b->basis || b->neval < 1 ||
        !ty->rel_cat || !ty->attach_ix || !ty->attach_subix || !ty->gpos || !ty->eval_spot_type || !ty->eval_spot) {
        set_error("%s", "bpl_score_token_prior: a required array is NULL"); return BPL_EINVAL;
    }
    if (g && (!g->g_ctrl || !g->g_invscale || !g->g_position)) {
        set_error("%s", "bpl_score_token_prior: g_ctrl/g_invscale/g_position are required when gradients are requested"); return BPL_EINVAL;
    }
    TPArgs a{};
    a.n_tokens = b->n_tokens; a.neval = b->neval;
    a.tok_stroke_off = b->tok_stroke_off; a.stroke_sub_off = b->stroke_sub_off;
    a.ctrl = b->ctrl; a.invscale = b->invscale; a.position = b->position; a.sigma = b->sigma; a.basis = b->basis;
    a.ty = *ty; a.c = *c; a.ll = ll;
    a.want_grad = g != nullptr;
    if (g) a.g = *g;
    const int wpb = 8;
    long long blocks = ((long long)b->n_tokens + wpb - 1) / wpb;
    const int cap = sm_count() * 32;
    if (cap > 0 && blocks > cap) blocks = cap;
    token_prior_kernel<<<(int)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(a);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
