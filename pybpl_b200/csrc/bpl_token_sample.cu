// bpl_token_sample.cu -- sample character tokens from P(token | type) on the device (token half).
//
// Replaces, for a whole batch, CharacterTokenDist.sample_token (pybpl/model/token_dist.py:226-262) =
// ConceptTokenDist.sample_token (:31-57): per stroke, in order,
//   StrokeTokenDist.sample_part_token        :410-430  shapes ~ Normal(shapes_type, sigma_shape)        (:314-332)
//                                                       invscales ~ Normal(type, sigma_invscale), drawn again
//                                                       until every scale is positive                     (:356-377)
//   RelationTokenDist.sample_relation_token  :465-487  'mid': eval_spot ~ Normal(type, sigma_attach), drawn again
//                                                       until inside [2, ncpt+1]                           (:517-540)
//   CharacterTokenDist.sample_location       :109-129  position = attach point of the ALREADY SAMPLED earlier
//                                                       strokes (objects/relation.py:34-67) + Normal(0, (sigma_x, sigma_y))
// (the reference fixes affine = None, epsilon = min_epsilon, blur_sigma = min_blur_sigma, :249-260; the caller fills
// eps / sigma).  Drawing the scale vector again until all entries are positive is the product of independent
// positive-truncated normals, so each scale is redrawn on its own.
//
// The reference draws from torch's global Mersenne generator, which cannot be reproduced on a GPU; parity for this
// kernel is distributional (tests/test_gpu_token_sample.py against samples of the live reference).  Here every token
// owns the Philox-4x32-10 stream (seed, subsequence = first_token + p): results do not depend on the launch shape or
// on how a particle set is sharded over GPUs.  One thread per token (strokes are sequential by construction).
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>

#include "bpl_common.cuh"
#include "bpl_token_common.cuh"

namespace bpl {
void set_error(const char *fmt, const char *detail);
void count_launch(int n);
int sm_count();

namespace {

struct TSArgs {
    int n_tokens, neval;
    const int *tok_stroke_off, *stroke_sub_off;
    const float *basis;
    bpl_token_type ty;              // eval_spot is ignored (it is an output here)
    bpl_token_prior_params c;
    unsigned long long seed, first_token;
    float *ctrl, *invscale, *position, *eval_spot;
};

__global__ void __launch_bounds__(128) token_sample_kernel(const TSArgs a) {
    float C0[NCPT], C1[NCPT];
#pragma unroll
    for (int k = 0; k < NCPT; ++k) { C0[k] = a.basis[k]; C1[k] = a.basis[(a.neval - 1) * NCPT + k]; }
    const float lb = 2.0f, ub = (float)(NCPT + 1);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.n_tokens; p += gridDim.x * blockDim.x) {
        curandStatePhilox4_32_10_t st;
        curand_init(a.seed, a.first_token + (unsigned long long)p, 0ULL, &st);
        const int k0 = a.tok_stroke_off[p], k1 = a.tok_stroke_off[p + 1];
        for (int k = k0; k < k1; ++k) {
            const int s0 = a.stroke_sub_off[k], s1 = a.stroke_sub_off[k + 1];
            // ---- stroke token: shapes, then scales (token_dist.py:410-430) ----
            for (int s = s0; s < s1; ++s) {
                const float *ty = a.ty.ctrl_type + (size_t)s * NCPT * 2;
                float *out = a.ctrl + (size_t)s * NCPT * 2;
#pragma unroll
                for (int e = 0; e < NCPT * 2; e += 2) {
                    const float2 z = curand_normal2(&st);
                    out[e] = fmaf(a.c.sigma_shape, z.x, ty[e]);
                    out[e + 1] = fmaf(a.c.sigma_shape, z.y, ty[e + 1]);
                }
            }
            for (int s = s0; s < s1; ++s) {
                const float ty = a.ty.invscale_type[s];
                float v = 0.0f;
                int tries = 0;
                do { v = fmaf(a.c.sigma_invscale, curand_normal(&st), ty); } while (v <= 0.0f && ++tries < 4096);
                a.invscale[s] = v > 0.0f ? v : fmaxf(ty, 1e-6f);      // (a type scale 100 sigma below zero has no positive draw)
            }
            // ---- relation token (token_dist.py:465-487) ----
            const int cat = a.ty.rel_cat[k];
            float u = 0.0f;
            if (cat == BPL_REL_MID) {
                const float ty = a.ty.eval_spot_type[k];
                int tries = 0;
                do { u = fmaf(a.c.sigma_attach, curand_normal(&st), ty); } while ((u < lb || u > ub) && ++tries < 4096);
                u = fminf(fmaxf(u, lb), ub);
            }
            if (a.eval_spot) a.eval_spot[k] = u;
            // ---- location (token_dist.py:109-129): attach point of the strokes sampled so far + noise ----
            float bx, by;
            if (cat == BPL_REL_UNIHIST) {
                bx = a.ty.gpos[2 * k]; by = a.ty.gpos[2 * k + 1];
            } else {
                const int j = k0 + a.ty.attach_ix[k];               // j < k: written above by this very thread
                const int sj0 = a.stroke_sub_off[j], sj1 = a.stroke_sub_off[j + 1];
                attach_point(cat, sj0, sj1, sj0 + a.ty.attach_subix[k], u, a.ctrl, a.invscale, a.position[2 * j],
                             a.position[2 * j + 1], C0, C1, bx, by);
            }
            const float2 z = curand_normal2(&st);
            a.position[2 * k] = fmaf(a.c.sigma_x, z.x, bx);
            a.position[2 * k + 1] = fmaf(a.c.sigma_y, z.y, by);
        }
    }
}

}  // namespace
}  // namespace bpl

using namespace bpl;

extern "C" int bpl_sample_tokens(const bpl_token_prior_params *c, const bpl_token_batch *b, const bpl_token_type *ty,
                                 unsigned long long seed, unsigned long long first_token, float *ctrl, float *invscale,
                                 float *position, float *eval_spot, void *stream) {
    if (!c || !b || !ty || !ctrl || !invscale || !position) { set_error("%s", "bpl_sample_tokens: NULL argument"); return BPL_EINVAL; }
    if (b->n_tokens <= 0) return 0;
    if (b->ncpt != NCPT) { set_error("%s", "bpl_sample_tokens is built for ncpt=5"); return BPL_ELIMIT; }
    if (!b->tok_stroke_off || !b->stroke_sub_off || !b->basis || b->neval < 1 || !ty->ctrl_type || !ty->invscale_type ||
        !ty->rel_cat || !ty->attach_ix || !ty->attach_subix || !ty->gpos || !ty->eval_spot_type) {
        set_error("%s", "bpl_sample_tokens: a required array is NULL"); return BPL_EINVAL;
    }
    TSArgs a{};
    a.n_tokens = b->n_tokens; a.neval = b->neval;
    a.tok_stroke_off = b->tok_stroke_off; a.stroke_sub_off = b->stroke_sub_off; a.basis = b->basis;
    a.ty = *ty; a.c = *c; a.seed = seed; a.first_token = first_token;
    a.ctrl = ctrl; a.invscale = invscale; a.position = position; a.eval_spot = eval_spot;
    long long blocks = ((long long)b->n_tokens + 127) / 128;
    const int cap = sm_count() * 64;
    if (cap > 0 && blocks > cap) blocks = cap;
    token_sample_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(a);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
