// bpl_type_prior.cu -- log P(type) for a ragged batch of character types, with its gradient.
//
// Replaces CharacterTypeDist.score_type (pybpl/model/type_dist.py:98-125), the objective of examples/optimize_type.py
// and the third term of the fit_image objective (model/model.py:59-62):
//   score_k :163-185; per stroke score_part_type :507-531 = score_nsub :269-294 + score_subIDs :328-353 +
//   score_shapes_type :388-423 (MultivariateNormal.log_prob) + score_invscales_type :450-478 (Gamma.log_prob);
//   score_relation_type :599-650 (category, then the position histogram of the stroke index --
//   library/spatial_OLD/spatial_model.py:86-112, spatial_hist.py:145-167,235-256 -- or the uniform attachment terms).
// Gradients go to the shapes and the scales: the only leaves the reference's autograd reaches (its histogram score
// leaves the graph through numpy and raises if gpos requires grad; a uniform density is flat).
//
// One warp per type: lanes over the type's sub-strokes (id chain, a 10x10 triangular solve in double -- the Cholesky
// factor of the primitive's covariance is read from the table -- and the Gamma term), then over its strokes.
#include <cuda_runtime.h>
#include <math.h>

#include "bpl_common.cuh"

namespace bpl {
void set_error(const char *fmt, const char *detail);
void count_launch(int n);
int sm_count();

namespace {

struct TQArgs {
    bpl_type_lib L;
    bpl_type_logtables T;
    int n_types;
    const int *tok_stroke_off, *stroke_sub_off, *subid;
    bpl_token_type ty;
    float *ll, *g_ctrl_type, *g_invscale_type;
};

__global__ void __launch_bounds__(256) type_prior_kernel(const TQArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const bpl_type_lib &L = a.L;
    for (int p = warp; p < a.n_types; p += nwarps) {
        const int k0 = a.tok_stroke_off[p], k1 = a.tok_stroke_off[p + 1], k = k1 - k0;
        const int s0 = a.stroke_sub_off[k0], s1 = a.stroke_sub_off[k1];
        double acc = 0.0;
        const bool bad = k < 1 || k > L.kmax;
        // ---- sub-strokes: id chain, shapes, scales ----
        for (int s = s0 + lane; s < s1 && !bad; s += 32) {
            const int id = a.subid[s];
            // first sub-stroke of its stroke?  (binary search over the token's strokes)
            int lo = k0, hi = k1 - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.stroke_sub_off[mid] <= s) lo = mid; else hi = mid - 1; }
            const bool first = a.stroke_sub_off[lo] == s;
            acc += (double)(first ? a.T.logp_start[id] : a.T.logp_T[(size_t)a.subid[s - 1] * L.n_prim + id]);      // :328-353
            // MultivariateNormal(mu, L L^T).log_prob(x) = -1/2 |L^-1 (x - mu)|^2 - sum log L_ii - 5 log 2 pi       // :388-423
            const float *mu = L.shape_mu + (size_t)id * 10, *Lc = L.shape_L + (size_t)id * 100;
            const float *x = a.ty.ctrl_type + (size_t)s * 10;
            double y[10], q = 0.0, logdet = 0.0;
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                double v = (double)x[r] - (double)mu[r];
#pragma unroll
                for (int c = 0; c < r; ++c) v -= (double)Lc[r * 10 + c] * y[c];
                y[r] = v / (double)Lc[r * 10 + r];
                q += y[r] * y[r];
                logdet += log((double)Lc[r * 10 + r]);
            }
            acc += -0.5 * q - logdet - 5.0 * 1.8378770664093454836;
            if (a.g_ctrl_type) {                   // d/dx = -Sigma^-1 (x - mu) = -L^-T y
                double z[10];
#pragma unroll
                for (int r = 9; r >= 0; --r) {
                    double v = y[r];
#pragma unroll
                    for (int c = r + 1; c < 10; ++c) v -= (double)Lc[c * 10 + r] * z[c];
                    z[r] = v / (double)Lc[r * 10 + r];
                }
#pragma unroll
                for (int r = 0; r < 10; ++r) a.g_ctrl_type[(size_t)s * 10 + r] = (float)(-z[r]);
            }
            // Gamma(con, rate).log_prob(x) = con log rate + (con - 1) log x - rate x - lgamma(con)                   // :450-478
            const double al = (double)L.gamma_con[id], be = (double)L.gamma_rate[id], xv = (double)a.ty.invscale_type[s];
            acc += al * log(be) + (al - 1.0) * log(xv) - be * xv - lgamma(al);
            if (a.g_invscale_type) a.g_invscale_type[s] = (float)((al - 1.0) / xv - be);
        }
        // ---- strokes: sub-stroke count and relation ----
        for (int kk = k0 + lane; kk < k1 && !bad; kk += 32) {
            const int i = kk - k0;
            const int nsub = a.stroke_sub_off[kk + 1] - a.stroke_sub_off[kk];
            acc += (nsub >= 1 && nsub <= L.nsub_max) ? (double)a.T.logp_nsub[(size_t)(k - 1) * L.nsub_max + nsub - 1] : -INFINITY;   // :269-294
            const int cat = a.ty.rel_cat[kk];
            if (i > 0) acc += (double)a.T.logp_mix[cat];                                                            // :623-626
            if (cat == BPL_REL_UNIHIST) {                                                                           // :629-634
                const int h = min(i, L.n_hist - 1), nb = L.hist_bins;
                const float gx = a.ty.gpos[2 * kk], gy = a.ty.gpos[2 * kk + 1];
                // numpy.histogram2d: bins [e_j, e_j+1), the last one closed; outside -> -inf (spatial_hist.py:252-253)
                int xi = -1, yi = -1;
                if (gx >= L.sp_xlab[0] && gx <= L.sp_xlab[nb]) { xi = 0; while (xi < nb - 1 && gx >= L.sp_xlab[xi + 1]) ++xi; }
                if (gy >= L.sp_ylab[0] && gy <= L.sp_ylab[nb]) { yi = 0; while (yi < nb - 1 && gy >= L.sp_ylab[yi + 1]) ++yi; }
                acc += (xi >= 0 && yi >= 0) ? (double)a.T.sp_logp[(size_t)h * nb * nb + xi * nb + yi] - (double)a.T.sp_log_binarea : -INFINITY;
            } else {
                // Outside the support the reference's Categorical / Uniform log_prob raises (validate_args) or returns
                // -inf; here such a type scores -inf -- and an out-of-range index is never followed.
                const int ai = a.ty.attach_ix[kk];
                const bool ai_ok = ai >= 0 && ai < i;
                acc += ai_ok ? -log((double)i) : -INFINITY;                                                          // :636-638
                if (cat == BPL_REL_MID && ai_ok) {                                                                  // :639-647
                    const int j = k0 + ai;
                    const int na = a.stroke_sub_off[j + 1] - a.stroke_sub_off[j];
                    const int asub = a.ty.attach_subix ? a.ty.attach_subix[kk] : 0;
                    const float es = a.ty.eval_spot_type ? a.ty.eval_spot_type[kk] : 2.0f;                          // (NULL: not checked)
                    const bool in_support = asub >= 0 && asub < na && es >= 2.0f && es <= (float)(BPL_NCPT + 1);     // Uniform(lb = 2, ub = ncpt + 1)
                    acc += in_support ? -log((double)na) - log((double)(BPL_NCPT + 1) - 2.0) : -INFINITY;
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) a.ll[p] = bad ? -INFINITY : (float)(acc + (double)a.T.logp_kappa[k - 1]);                    // :163-185
    }
}

}  // namespace
}  // namespace bpl

using namespace bpl;

extern "C" int bpl_score_type_prior(const bpl_type_lib *lib, const bpl_type_logtables *tabs, int n_types,
                                    const int *tok_stroke_off, const int *stroke_sub_off, const int *subid,
                                    const bpl_token_type *ty, float *ll, float *g_ctrl_type, float *g_invscale_type,
                                    void *stream) {
    if (!lib || !tabs || !ty || !ll || !tok_stroke_off || !stroke_sub_off || !subid) {
        set_error("%s", "bpl_score_type_prior: NULL argument"); return BPL_EINVAL;
    }
    if (n_types <= 0) return 0;
    if (!lib->shape_mu || !lib->shape_L || !lib->gamma_con || !lib->gamma_rate || !lib->sp_xlab || !lib->sp_ylab || !tabs->logp_kappa ||
        !tabs->logp_nsub || !tabs->logp_start || !tabs->logp_T || !tabs->logp_mix || !tabs->sp_logp || !ty->ctrl_type ||
        !ty->invscale_type || !ty->rel_cat || !ty->attach_ix || !ty->gpos) {
        set_error("%s", "bpl_score_type_prior: incomplete tables or type arrays"); return BPL_EINVAL;
    }
    TQArgs a{};
    a.L = *lib; a.T = *tabs; a.n_types = n_types;
    a.tok_stroke_off = tok_stroke_off; a.stroke_sub_off = stroke_sub_off; a.subid = subid;
    a.ty = *ty; a.ll = ll; a.g_ctrl_type = g_ctrl_type; a.g_invscale_type = g_invscale_type;
    const int wpb = 8;
    long long blocks = ((long long)n_types + wpb - 1) / wpb;
    const int cap = sm_count() * 32;
    if (cap > 0 && blocks > cap) blocks = cap;
    type_prior_kernel<<<(int)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(a);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
