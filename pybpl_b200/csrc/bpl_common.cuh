// bpl_common.cuh -- shared device helpers for the pyBPL render+score kernels (sm_100a).
//
// Geometry of one image buffer in shared memory: the 105x105 image sits inside a 115x115
// frame of zeros (5-pixel border = the half-width of the 11-tap Gaussian, parameters.py:41).
// The row stride 115 is odd, so a warp whose lanes walk down a column (stride 115) and a warp
// whose lanes walk along a row (stride 1) are both bank-conflict free, and the zero border
// implements the zero padding of F.conv2d (pybpl/util/general.py:161-163) with no predicates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pybpl_b200.h"

// Optional per-phase cycle counters (build with -DBPL_PHASE_TIMING; thread 0 of every CTA accumulates clock64()
// deltas, written to a.pt_buf[blockIdx.x*16 + phase]; used only by tools/phase_timing_sparse.py).
#ifdef BPL_PHASE_TIMING
#define BPL_PT(n) do { if (threadIdx.x == 0) { const long long t1_ = clock64(); pt_acc[n] += t1_ - pt_t0; pt_t0 = t1_; } } while (0)
#else
#define BPL_PT(n) do { } while (0)
#endif

namespace bpl {

constexpr int IMG = BPL_IMSIZE;            // 105
constexpr int PAD = BPL_FSIZE / 2;         // 5
constexpr int STRIDE = IMG + 2 * PAD;      // 115 (odd)
constexpr int BUFN = STRIDE * STRIDE;      // 13225 floats
constexpr int BUFP = (BUFN + 3) / 4 * 4;   // allocation per buffer (16-byte multiple): 13228 floats = 52,912 B
constexpr int ORIGIN = PAD * STRIDE + PAD; // buffer index of pixel (0,0)
constexpr int NPIX = IMG * IMG;
constexpr int MAXS = BPL_MAX_SUBSTROKES;
constexpr int MAXK = BPL_MAX_STROKES;
constexpr int NCPT = BPL_NCPT;
constexpr int MAX_TABLE = 1024;            // floats of basis table kept in shared memory
constexpr int IMG_WORDS = BPL_IMG_WORDS;
constexpr float TORCH_EPS = 1.1920928955078125e-07f;  // torch.finfo(float32).eps (clamp_probs)

// Rendering constants pre-digested on the host (pybpl/parameters.py:19-41, rendering.py:151-166).
struct RenderConsts {
    int ink_ncon;
    float ink_pp, ink_max_dist;
    float ink_f;     // ink_pp / ink_max_dist (d ink / d dist)
    float inv_max_dist;   // 1 / ink_max_dist when that is exact (a power of two), else 0: x / 2 == x * 0.5 bit for bit
    float c1, c2;    // H_broaden = c1 * (1,2,1)x(1,2,1) + c2 * delta
    int fhalf;       // half-width of the Gaussian, fsize / 2 <= PAD: taps beyond it are zero (parameters.py:41, general.py:196-207)
    float rmax, cmax; // largest in-bounds row / column coordinate, imsize - 1 (rendering.py:27-31); read by the raw-trajectory
                      // walk only (the any-size kernels set them to H - 1, W - 1), the control-point stages keep IMG - 1
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV doubles over the block; result valid in every thread.  scratch: NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = (lane < nw) ? scratch[i * 32 + lane] : 0.0;
        v[i] = warp_sum(x);
    }
}

// ---------------------------------------------------------------------------------------
// One separable stencil pass over the whole image (a "sweep").
//   VERT : slide down rows (vertical filter), lanes across columns; else slide along a row,
//          lanes across rows.  Either way lane addresses are conflict free (odd stride).
//   TAPS : 3 (broaden) or 11 (Gaussian); symmetric taps w[f][d], d = distance from centre.
//   R    : outputs per work item (IMG % R == 0); an item loads R+TAPS-1 inputs into registers
//          once and produces R outputs, so shared-memory reads per output are (R+TAPS-1)/R.
//   NF   : filters evaluated on the same loaded window (the backward needs g and h~ together).
//   ld   : applied to every loaded input (identity, or min(x,1) = rendering.py:171 on the fly)
//   epi  : epi(r, c, acc[NF]) consumes each output (store / combine / likelihood).
// ---------------------------------------------------------------------------------------
template <int TAPS, int R, bool VERT, int NF, class Load, class Epi>
__device__ __forceinline__ void sweep(const float *__restrict__ in, const float (&w)[NF][TAPS / 2 + 1], Load ld,
                                      Epi epi) {
    static_assert(IMG % R == 0, "bands must tile the image");
    constexpr int HALF = TAPS / 2;
    constexpr int NB = IMG / R;
    constexpr int SA = VERT ? STRIDE : 1;   // stride along the sliding axis
    constexpr int SL = VERT ? 1 : STRIDE;   // stride across lanes
    constexpr int U = (R % 7 == 0) ? 7 : 5;  // outputs per unrolled block (keeps the loop body inside the I-cache)
    static_assert(R % U == 0, "block structure");
    for (int item = threadIdx.x; item < NB * IMG; item += blockDim.x) {
        const int band = item / IMG;
        const int line = item - band * IMG;
        const int a0 = band * R;
        const float *p = in + ORIGIN + (a0 - HALF) * SA + line * SL;
        float x[U + TAPS - 1];              // x[0 .. TAPS-2]: inputs before the block's first new one
#pragma unroll
        for (int i = 0; i < TAPS - 1; ++i) x[i] = ld(p[i * SA]);
#pragma unroll 1
        for (int ob = 0; ob < R; ob += U) {
            const float *q = p + (ob + TAPS - 1) * SA;
#pragma unroll
            for (int u = 0; u < U; ++u) x[TAPS - 1 + u] = ld(q[u * SA]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float acc[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) acc[f] = w[f][0] * x[u + HALF];
#pragma unroll
                for (int d = 1; d <= HALF; ++d) {
                    const float s = x[u + HALF - d] + x[u + HALF + d];
#pragma unroll
                    for (int f = 0; f < NF; ++f) acc[f] = fmaf(w[f][d], s, acc[f]);
                }
                if (VERT) epi(a0 + ob + u, line, acc); else epi(line, a0 + ob + u, acc);
            }
#pragma unroll
            for (int k = 0; k < TAPS - 1; ++k) x[k] = x[k + U];
        }
    }
}

// The same pass confined to a box of the image (inclusive bounds; the extent along the sliding axis is a whole
// number of bands of R, so that no output needs a predicate and every pass of a chain touches exactly the box).
// Measured alternatives: a run length chosen at run time so that a small box still occupies every thread, and
// predicated partial bands, were both ~10 % slower (with 24 warps per SM the passes are issue bound; shorter runs
// re-load more halo per output).
struct Box {
    int r0, r1, c0, c1;
    __device__ __forceinline__ bool empty() const { return r0 > r1 || c0 > c1; }
};
template <int TAPS, int R, bool VERT, int NF, class Load, class Epi>
__device__ __forceinline__ void sweep_box(const float *__restrict__ in, const float (&w)[NF][TAPS / 2 + 1], Load ld,
                                          Epi epi, const Box &bx) {
    static_assert(IMG % R == 0, "bands must tile the image");
    constexpr int HALF = TAPS / 2;
    constexpr int SA = VERT ? STRIDE : 1;   // stride along the sliding axis
    constexpr int SL = VERT ? 1 : STRIDE;   // stride across lanes
    constexpr int U = (R % 7 == 0) ? 7 : 5;  // outputs per unrolled block (keeps the loop body inside the I-cache)
    static_assert(R % U == 0, "block structure");
    const int l0 = VERT ? bx.c0 : bx.r0, nl = (VERT ? bx.c1 : bx.r1) - l0 + 1;
    const int s0 = VERT ? bx.r0 : bx.c0, nb = ((VERT ? bx.r1 : bx.c1) - s0 + 1) / R;
    if (nl <= 0 || nb <= 0) return;
    for (int item = threadIdx.x; item < nb * nl; item += blockDim.x) {
        int bi = 0;                          // item / nl without a division: a line has at most IMG / R bands
#pragma unroll
        for (int k = 1; k < IMG / R; ++k) bi += (item >= k * nl);
        const int line = l0 + item - bi * nl;
        const int a0 = s0 + bi * R;
        const float *p = in + ORIGIN + (a0 - HALF) * SA + line * SL;
        float x[U + TAPS - 1];              // x[0 .. TAPS-2]: inputs before the block's first new one
#pragma unroll
        for (int i = 0; i < TAPS - 1; ++i) x[i] = ld(p[i * SA]);
#pragma unroll 1
        for (int ob = 0; ob < R; ob += U) {
            const float *q = p + (ob + TAPS - 1) * SA;
#pragma unroll
            for (int u = 0; u < U; ++u) x[TAPS - 1 + u] = ld(q[u * SA]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float acc[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) acc[f] = w[f][0] * x[u + HALF];
#pragma unroll
                for (int d = 1; d <= HALF; ++d) {
                    const float s = x[u + HALF - d] + x[u + HALF + d];
#pragma unroll
                    for (int f = 0; f < NF; ++f) acc[f] = fmaf(w[f][d], s, acc[f]);
                }
                if (VERT) epi(a0 + ob + u, line, acc); else epi(line, a0 + ob + u, acc);
            }
#pragma unroll
            for (int k = 0; k < TAPS - 1; ++k) x[k] = x[k + U];
        }
    }
}

struct LoadId { __device__ __forceinline__ float operator()(float v) const { return v; } };
struct LoadMin1 { __device__ __forceinline__ float operator()(float v) const { return fminf(v, 1.0f); } };

// ---------------------------------------------------------------------------------------
// Per-pixel Bernoulli log-likelihood in torch's formulation (pybpl/model/image_dist.py:75-78;
// torch/distributions/utils.py clamp_probs + probs_to_logits, bernoulli.py log_prob =
// -binary_cross_entropy_with_logits):
//   pt = clamp(p, eps32, 1-eps32);  l = log(pt) - log1p(-pt)
//   lp = -((1-x) * l - (min(l,0) - log1p(exp(-|l|))))
// The operation order is kept (the catastrophic cancellation in it is part of the reference's
// result).  dlp is what autograd returns for d lp / d p.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pixel_lp(float p, bool x, bool want_grad, float &lp, float &dlp) {
    const float pt = fminf(fmaxf(p, TORCH_EPS), 1.0f - TORCH_EPS);
    // The log that is large in magnitude may be the 3-ulp hardware one (MUFU.LG2); the one that is
    // close to zero keeps libm accuracy through log1p.  (1 - pt is exact for pt >= 0.5.)
    float L1, L2;
    if (pt < 0.5f) { L1 = __logf(pt); L2 = log1pf(-pt); }
    else { const float q = 1.0f - pt; L2 = __logf(q); L1 = log1pf(-q); }
    const float l = L1 - L2;
    // log1p(exp(-|l|)) equals -log(1-p) for l < 0 and -log(p) for l >= 0 (exp(l) = p/(1-p)); using the
    // two logs already at hand instead of exp+log1p moves individual pixels by <= 1 ulp(l), with no
    // common sign (only values repeated across thousands of pixels can bias the sum, and the one such
    // value -- the background -- is handled exactly by pixel_lp_exact).
    const float lg = (l < 0.0f) ? -L2 : -L1;
    const float ls = fminf(l, 0.0f) - lg;
    const float t = x ? 0.0f : l;
    lp = -(t - ls);
    if (want_grad) {
        // autograd: (x - sigmoid(l)) (1/p + 1/(1-p)) = 1/p for x = 1, -1/(1-p) for x = 0; zero outside clamp_probs
        if (p < TORCH_EPS || p > 1.0f - TORCH_EPS) dlp = 0.0f;
        else dlp = x ? __frcp_rn(p) : -__frcp_rn(1.0f - p);
    }
}

// Forward-only score of one non-background pixel with a single logarithm: lp = log p (x = 1) or
// log(1 - p) (x = 0).  torch's formulation (above) equals this up to a rounding term of at most ulp(l)/2
// per pixel with no common sign; the one value that is repeated over thousands of pixels (the background)
// never comes here.  Every lp is <= 0, so a per-pixel RELATIVE error bound carries over to the sum:
//   a < 1/2        : hardware log (MUFU.LG2), 3 ulp of a value >= 0.69             -> 4e-7
//   1 - a < 1/8    : -b (1 + b/2 + ... + b^7/8), truncation b^8/9                  -> 7e-9 (+ rounding 2e-7)
//   otherwise      : hardware log, absolute error 2^-21.41 on |log a| >= 0.1335    -> 2.7e-6
// i.e. ll is within 3e-6 relative of the exact sum; a and its complement b = 1 - a are both exact inputs
// (b = p or 1 - p with p >= 1/2).  The fused backward kernels score with this function too; the dense probability-map
// kernel keeps the two-logarithm form (pixel_lp).
__device__ __forceinline__ float pixel_lp_fwd(float p, bool x) {
    const float pt = fminf(fmaxf(p, TORCH_EPS), 1.0f - TORCH_EPS);
    const float q = 1.0f - pt;
    const float a = x ? pt : q, b = x ? q : pt;
    float s = fmaf(b, 0.125f, 1.0f / 7.0f);
    s = fmaf(b, s, 1.0f / 6.0f);
    s = fmaf(b, s, 0.2f);
    s = fmaf(b, s, 0.25f);
    s = fmaf(b, s, 1.0f / 3.0f);
    s = fmaf(b, s, 0.5f);
    s = fmaf(b, s, 1.0f);
    return (b < 0.125f) ? -b * s : __logf(a);
}

// The same in double, each libm result rounded to float (= correctly rounded fp32 libm), used
// once per token for the background value shared by ~90% of the pixels: a 1-ulp wobble there
// would be multiplied by ~10^4 pixels.
__device__ inline void pixel_lp_exact(float p, bool x, float &lp, float &dlp) {
    const float pt = fminf(fmaxf(p, TORCH_EPS), 1.0f - TORCH_EPS);
    const float l1 = (float)log((double)pt);
    const float l2 = (float)log1p(-(double)pt);
    const float l = l1 - l2;
    const float ex = (float)exp(-(double)fabsf(l));
    const float lg = (float)log1p((double)ex);
    const float ls = fminf(l, 0.0f) - lg;
    const float t = x ? 0.0f * l : l;
    lp = -(t - ls);
    if (p < TORCH_EPS || p > 1.0f - TORCH_EPS) {
        dlp = 0.0f;
    } else {
        const float sg = (float)(1.0 / (1.0 + exp(-(double)l)));
        dlp = ((x ? 1.0f : 0.0f) - sg) * (1.0f / p + 1.0f / (1.0f - p));
    }
}

}  // namespace bpl
