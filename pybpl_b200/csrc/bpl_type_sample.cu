// bpl_type_sample.cu -- sample character types from the prior on the device (type half).
//
// Replaces, for a whole batch, CharacterTypeDist.sample_type (pybpl/model/type_dist.py:187-207) =
// ConceptTypeDist.sample_type (:56-96):
//   k ~ Categorical(pkappa) + 1                                                   sample_k            :148-161
//   per stroke, in order:
//     StrokeTypeDist.sample_part_type                                             :480-505
//       nsub ~ Categorical(pmat_nsub[k-1]) + 1                                    sample_nsub         :245-267
//       sub-stroke ids: a Markov chain over the primitives (logStart, logT rows)  sample_subIDs       :296-326
//       shapes ~ MultivariateNormal(mu[id], Sigma[id])  (10-d, Cholesky factor)   sample_shapes_type  :355-386
//       invscales ~ Gamma(concentration[id], rate[id])                            sample_invscales_type :425-448
//     RelationTypeDist.sample_relation_type                                       :541-597
//       first stroke 'unihist', later ones ~ Categorical(rel.mixprob);
//       'unihist': gpos ~ position histogram of the stroke index (last histogram for index >= 2):
//                  a bin ~ Categorical(exp logpYX), then uniform inside it
//                  (library/spatial_OLD/spatial_model.py:140-163, spatial_hist.py:107-143)
//       else: attach_ix ~ uniform over the earlier strokes; 'mid': attach_subix ~ uniform over that stroke's
//             sub-strokes, eval_spot ~ Uniform(2, ncpt+1)
//
// Types are ragged, so the kernel runs twice over the same counter-based streams (Philox-4x32-10, subsequence =
// first_type + p): a counting pass that only reports strokes / sub-strokes per type, and -- after an exclusive scan by
// the caller -- a writing pass that draws the very same numbers again and stores them at the scanned offsets.
// Categorical draws invert cumulative tables (binary search; the 1212 x 1212 transition table is 5.9 MB, L2-resident).
// The reference's torch generator cannot be reproduced on a GPU: parity is distributional
// (tests/test_gpu_type_sample.py, against draws of the live reference and against the tables themselves).
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>

#include "bpl_common.cuh"

namespace bpl {
void set_error(const char *fmt, const char *detail);
void count_launch(int n);
int sm_count();

namespace {

typedef curandStatePhilox4_32_10_t Rng;

// smallest i with cdf[i] >= u  (u in (0, 1]; the last entry of every table is exactly 1)
__device__ __forceinline__ int invert_cdf(const float *cdf, int n, float u) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] >= u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Gamma(alpha, 1) by Marsaglia & Tsang (2000); alpha < 1 through Gamma(alpha + 1) * U^(1/alpha)
__device__ float gamma_draw(Rng *st, float alpha) {
    const float a = alpha < 1.0f ? alpha + 1.0f : alpha;
    const float d = a - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
    float v, x;
    for (int tries = 0; tries < 256; ++tries) {
        x = curand_normal(st);
        v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        const float u = curand_uniform(st);
        if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) break;
    }
    float g = d * v;
    if (alpha < 1.0f) g *= powf(curand_uniform(st), 1.0f / alpha);
    return g;
}

struct TYArgs {
    bpl_type_lib L;
    int n_types;
    unsigned long long seed, first_type;
    const int *tok_stroke_off, *tok_sub_off;      // NULL in the counting pass
    int *n_strokes, *n_subs;                      // counting pass outputs
    bpl_type_out o;
};

constexpr int MAX_NSUB = 16;        // sub-strokes per stroke the tables can produce (pmat_nsub has 10 columns)

__global__ void __launch_bounds__(128) type_sample_kernel(const TYArgs a) {
    const bpl_type_lib &L = a.L;
    const bool write = a.tok_stroke_off != nullptr;
    const float lb = 2.0f, ub = (float)(BPL_NCPT + 1);                                           // bspline_gen_s
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.n_types; p += gridDim.x * blockDim.x) {
        Rng st;
        curand_init(a.seed, a.first_type + (unsigned long long)p, 0ULL, &st);
        const int k = invert_cdf(L.pkappa, L.kmax, curand_uniform(&st)) + 1;                    // sample_k (:148-161)
        const int kbase = write ? a.tok_stroke_off[p] : 0;
        int sbase = write ? a.tok_sub_off[p] : 0;
        int subs_total = 0;
        int nsub_of[BPL_MAX_STROKES];
        for (int i = 0; i < k; ++i) {
            // ---- stroke type (:480-505) ----
            const int nsub = invert_cdf(L.pmat_nsub + (size_t)(k - 1) * L.nsub_max, L.nsub_max, curand_uniform(&st)) + 1;
            nsub_of[i] = nsub;
            int ids[MAX_NSUB];
            int prev = -1;
            for (int s = 0; s < nsub; ++s) {                                                    // sample_subIDs (:296-326)
                const float *cdf = prev < 0 ? L.cdf_start : L.cdf_T + (size_t)prev * L.n_prim;
                prev = invert_cdf(cdf, L.n_prim, curand_uniform(&st));
                ids[s] = prev;
                if (write) a.o.subid[sbase + s] = prev;
            }
            for (int s = 0; s < nsub; ++s) {                                                    // sample_shapes_type (:355-386)
                float z[BPL_NCPT * 2];
#pragma unroll
                for (int e = 0; e < BPL_NCPT * 2; e += 2) { const float2 n2 = curand_normal2(&st); z[e] = n2.x; z[e + 1] = n2.y; }
                if (write) {
                    const float *mu = L.shape_mu + (size_t)ids[s] * 10, *Lc = L.shape_L + (size_t)ids[s] * 100;
                    float *out = a.o.ctrl_type + (size_t)(sbase + s) * 10;
#pragma unroll
                    for (int r = 0; r < 10; ++r) {
                        float acc = mu[r];
#pragma unroll
                        for (int c = 0; c <= r; ++c) acc = fmaf(Lc[r * 10 + c], z[c], acc);
                        out[r] = acc;
                    }
                }
            }
            for (int s = 0; s < nsub; ++s) {                                                    // sample_invscales_type (:425-448)
                const float g = gamma_draw(&st, L.gamma_con[ids[s]]) / L.gamma_rate[ids[s]];
                if (write) a.o.invscale_type[sbase + s] = g;
            }
            // ---- relation type (:541-597) ----
            int cat = BPL_REL_UNIHIST, aix = 0, asub = 0;
            float gx = 0.0f, gy = 0.0f, spot = 0.0f;
            if (i > 0) cat = invert_cdf(L.rel_mixprob, 4, curand_uniform(&st));
            if (cat == BPL_REL_UNIHIST) {
                const int h = min(i, L.n_hist - 1);                                             // spatial_model.py:179-189
                const int nb = L.hist_bins;
                const int lin = invert_cdf(L.sp_cdf + (size_t)h * nb * nb, nb * nb, curand_uniform(&st));
                const int xi = lin / nb, yi = lin - xi * nb;                                    // ind2sub, spatial_hist.py:124
                const float ux = curand_uniform(&st), uy = curand_uniform(&st);
                gx = L.sp_xlab[xi] + ux * (L.sp_xlab[xi + 1] - L.sp_xlab[xi]);
                gy = L.sp_ylab[yi] + uy * (L.sp_ylab[yi + 1] - L.sp_ylab[yi]);
            } else {
                aix = min((int)(curand_uniform(&st) * (float)i), i - 1);                        // uniform over the earlier strokes
                if (cat == BPL_REL_MID) {
                    const int na = nsub_of[aix];
                    asub = min((int)(curand_uniform(&st) * (float)na), na - 1);
                    spot = lb + curand_uniform(&st) * (ub - lb);
                }
            }
            if (write) {
                const int kk = kbase + i;
                a.o.stroke_nsub[kk] = nsub;
                a.o.rel_cat[kk] = cat; a.o.attach_ix[kk] = aix; a.o.attach_subix[kk] = asub;
                a.o.gpos[2 * kk] = gx; a.o.gpos[2 * kk + 1] = gy;
                a.o.eval_spot_type[kk] = spot;
            }
            sbase += nsub;
            subs_total += nsub;
        }
        if (!write) { a.n_strokes[p] = k; a.n_subs[p] = subs_total; }
    }
}

}  // namespace
}  // namespace bpl

using namespace bpl;

extern "C" int bpl_sample_types(const bpl_type_lib *lib, int n_types, unsigned long long seed, unsigned long long first_type,
                                const int *tok_stroke_off, const int *tok_sub_off, int *n_strokes, int *n_subs,
                                const bpl_type_out *out, void *stream) {
    if (!lib) { set_error("%s", "bpl_sample_types: lib is NULL"); return BPL_EINVAL; }
    if (n_types <= 0) return 0;
    if (!lib->pkappa || !lib->pmat_nsub || !lib->cdf_start || !lib->cdf_T || !lib->shape_mu || !lib->shape_L || !lib->gamma_con ||
        !lib->gamma_rate || !lib->rel_mixprob || !lib->sp_cdf || !lib->sp_xlab || !lib->sp_ylab || lib->n_prim < 1 || lib->kmax < 1 ||
        lib->kmax > BPL_MAX_STROKES || lib->nsub_max < 1 || lib->nsub_max > MAX_NSUB || lib->n_hist < 1 || lib->hist_bins < 1) {
        set_error("%s", "bpl_sample_types: incomplete or oversized type library"); return BPL_EINVAL;
    }
    const bool write = tok_stroke_off != nullptr;
    if (write) {
        if (!tok_sub_off || !out || !out->stroke_nsub || !out->subid || !out->ctrl_type || !out->invscale_type || !out->rel_cat ||
            !out->attach_ix || !out->attach_subix || !out->gpos || !out->eval_spot_type) {
            set_error("%s", "bpl_sample_types: writing pass needs both offset arrays and every output"); return BPL_EINVAL;
        }
    } else if (!n_strokes || !n_subs) {
        set_error("%s", "bpl_sample_types: counting pass needs n_strokes and n_subs"); return BPL_EINVAL;
    }
    TYArgs a{};
    a.L = *lib; a.n_types = n_types; a.seed = seed; a.first_type = first_type;
    a.tok_stroke_off = tok_stroke_off; a.tok_sub_off = tok_sub_off; a.n_strokes = n_strokes; a.n_subs = n_subs;
    if (out) a.o = *out;
    long long blocks = ((long long)n_types + 127) / 128;
    const int cap = sm_count() * 64;
    if (cap > 0 && blocks > cap) blocks = cap;
    type_sample_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(a);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
