// bpl_anysize.cuh -- rendering.render_image for ANY Parameters.imsize (pybpl/parameters.py:21; rendering.py:185-229 with
// ps.imsize = (H, W): rows are bounded by imsize[0], columns by imsize[1], rendering.py:27-31,214).
//
// The fused kernels are built around the reference's one default, 105 x 105 (band structure 105 = 3*35 = 7*15 = 21*5,
// image buffers that fill the shared memory).  This file is the general case: the same pipeline -- add_stroke
// (:58-127), broaden_and_blur (:130-182), the epsilon mix (:226-227), Bernoulli log_prob (model/image_dist.py:75-78) and
// the adjoint of all of it -- with run-time geometry, one CTA per character and dense H x W canvases in a workspace the
// caller provides (global memory: for the sizes this serves the canvases of the resident CTAs live in the 126 MB L2).
// It shares the raw-trajectory walk (walk_round / substroke_stats / substroke_backward), the filter constants and the
// per-pixel likelihood with the 105 x 105 kernels, so the two paths agree to rounding where they overlap.
// Bound: L2 bandwidth -- every stencil pass reads and writes whole canvases; this is the completeness path, the hot
// path is bpl_forward1s / bpl_backward_teams.
#pragma once

namespace bpl {

struct AnyGeom {
    int H, W;
    float *ws;                // [gridDim.x][NBUF][H*W] canvases of the resident CTAs
    const float *images;      // [T][H][W] float, exactly 0 or 1 (validated by the caller), or NULL
};

constexpr int ANY_THREADS = 256;
constexpr int ANY_FWD_BUFS = 2;      // ping-pong
constexpr int ANY_BWD_BUFS = 7;      // I_n, V1, X1, V2, X2, G, T

// one stroke: bilinear splat with global atomics (rendering.py:111-125), run-time row stride
template <class Src>
__device__ __forceinline__ void any_splat(const Src &src, int n, int lane, const RenderConsts &rc, const Stats &s,
                                          float *img, int W) {
    Walk w{0, 0.0f, 0.0f, 0};
    for (int base = 0; base < n; base += 32) {
        const Pt p = walk_round(src, n, base, lane, rc, w);
        if (!p.inb) continue;
        const float ink = final_ink(p, s);
        const float rf = floorf(p.r), cf = floorf(p.c), rce = ceilf(p.r), cce = ceilf(p.c);
        const float fr = __fsub_rn(p.r, rf), fc = __fsub_rn(p.c, cf);
        const float gr = __fsub_rn(1.0f, fr), gc = __fsub_rn(1.0f, fc);
        const int r0 = (int)rf, c0 = (int)cf, r1 = (int)rce, c1 = (int)cce;
        atomicAdd(img + r0 * W + c0, __fmul_rn(__fmul_rn(ink, gr), gc));
        atomicAdd(img + r1 * W + c0, __fmul_rn(__fmul_rn(ink, fr), gc));
        atomicAdd(img + r0 * W + c1, __fmul_rn(__fmul_rn(ink, gr), fc));
        atomicAdd(img + r1 * W + c1, __fmul_rn(__fmul_rn(ink, fr), fc));
    }
}

// dst = c1 * (1,2,1)x(1,2,1) * src + c2 * src with zero padding (rendering.py:160-169; symmetric, so also its adjoint)
__device__ __forceinline__ void any_broaden(float *dst, const float *src, int H, int W, float c1, float c2, bool clamp1) {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int r = i / W, c = i - r * W;
        float h[3];
#pragma unroll
        for (int dr = -1; dr <= 1; ++dr) {
            const int rr = r + dr;
            float v = 0.0f;
            if (rr >= 0 && rr < H) {
                const float *row = src + rr * W;
                const float xl = c > 0 ? row[c - 1] : 0.0f, xr = c < W - 1 ? row[c + 1] : 0.0f;
                v = xl + 2.0f * row[c] + xr;
            }
            h[dr + 1] = v;
        }
        float v = fmaf(c1, h[0] + 2.0f * h[1] + h[2], c2 * src[i]);
        if (clamp1) v = fminf(v, 1.0f);                     // clamp_(max=1)  (:171)
        dst[i] = v;
    }
}

// one 11-tap pass with zero padding (util/general.py:161-163) along rows (VERT) or columns; MIN1: min(., 1) on load.
// Returns acc at pixel i for the filter k[0..PAD].
template <bool VERT, bool MIN1>
__device__ __forceinline__ float any_tap(const float *src, int r, int c, int H, int W, const float *k) {
    const int i = r * W + c;
    const int step = VERT ? W : 1, pos = VERT ? r : c, lim = VERT ? H : W;
    float x0 = src[i];
    if (MIN1) x0 = fminf(x0, 1.0f);
    float acc = k[0] * x0;
#pragma unroll
    for (int d = 1; d <= PAD; ++d) {
        float lo = (pos - d >= 0) ? src[i - d * step] : 0.0f;
        float hi = (pos + d < lim) ? src[i + d * step] : 0.0f;
        if (MIN1) { lo = fminf(lo, 1.0f); hi = fminf(hi, 1.0f); }
        acc = fmaf(k[d], lo + hi, acc);
    }
    return acc;
}
template <bool VERT, bool MIN1>
__device__ __forceinline__ void any_gauss(float *dst, const float *src, int H, int W, const float *g) {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int r = i / W, c = i - r * W;
        dst[i] = any_tap<VERT, MIN1>(src, r, c, H, W, g);
    }
}

struct AnyShared {
    Misc ms;
};

// the points of character p into the zeroed canvas A; sets ms->off_page
__device__ __forceinline__ void any_points(const KArgs &a, const TokGeom &t, float *A, int W, Misc *ms) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    bool any_out = false;
    for (int s = warp; s < t.S; s += nw) {
        const int o = a.sub_pt_off[t.s0 + s], n = a.sub_pt_off[t.s0 + s + 1] - o;
        const RawSrc src{reinterpret_cast<const float2 *>(a.pts) + o};
        const Stats st = substroke_stats(src, n, lane, a.rc);
        any_out |= st.any_out;
        if (st.cnt > 0) any_splat(src, n, lane, a.rc, st, A, W);
    }
    any_out = __any_sync(0xffffffffu, any_out);
    if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
}

__device__ __forceinline__ float any_mix(float v, float eps) {      // clamp_(0,1) (:180) and the mix (:226-227)
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    if (eps > 0.0f) v = __fadd_rn(__fmul_rn(1.0f - eps, v), __fmul_rn(eps, __fsub_rn(1.0f, v)));
    return v;
}

// ---- forward: pimg [P][H][W], off_page [P], ll [P] against images ----
__global__ void __launch_bounds__(ANY_THREADS) bpl_render_any_kernel(const KArgs a, const AnyGeom gm) {
    __shared__ AnyShared sh;
    Misc *ms = &sh.ms;
    const int H = gm.H, W = gm.W, N = H * W;
    float *buf0 = gm.ws + (size_t)blockIdx.x * ANY_FWD_BUFS * N, *buf1 = buf0 + N;
    if (threadIdx.x == 0) { ms->c_eps = ms->c_sigma = nanf(""); }
    __syncthreads();
    for (int p = blockIdx.x; p < a.n_tokens; p += gridDim.x) {
        const TokGeom t = token_geom<true>(a, p);
        const float eps = a.eps[p], sigma = a.sigma[p];
        if (threadIdx.x == 0) ms->off_page = 0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) buf0[i] = 0.0f;
        token_consts(eps, sigma, false, ms, a.rc.fhalf);
        __syncthreads();
        any_points(a, t, buf0, W, ms);
        __syncthreads();
        float *src = buf0, *dst = buf1;
        for (int it = 0; it < a.rc.ink_ncon; ++it) {
            any_broaden(dst, src, H, W, a.rc.c1, a.rc.c2, it == a.rc.ink_ncon - 1);
            __syncthreads();
            float *tmp = src; src = dst; dst = tmp;
        }
        const bool min1 = a.rc.ink_ncon == 0;      // no broaden pass carried the clamp
        if (sigma > 0.0f) {
            float g[PAD + 1];
#pragma unroll
            for (int d = 0; d <= PAD; ++d) g[d] = ms->g[d];
            for (int rep = 0; rep < 2; ++rep) {
                if (rep == 0 && min1) any_gauss<true, true>(dst, src, H, W, g);
                else any_gauss<true, false>(dst, src, H, W, g);
                __syncthreads();
                any_gauss<false, false>(src, dst, H, W, g);
                __syncthreads();
            }
        }
        const float bgv = eps > 0.0f ? eps : 0.0f;
        const float lp_bg0 = ms->lp_bg[0], lp_bg1 = ms->lp_bg[1];
        const float *img = gm.images ? gm.images + (size_t)(a.img_idx ? a.img_idx[p] : 0) * N : nullptr;
        float *out = a.pimg ? a.pimg + (size_t)p * N : nullptr;
        double acc = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            float v = src[i];
            if (min1 && !(sigma > 0.0f)) v = fminf(v, 1.0f);
            v = any_mix(v, eps);
            if (out) out[i] = v;
            if (img) {
                const bool x = img[i] != 0.0f;
                float lp = x ? lp_bg1 : lp_bg0, d;
                if (v != bgv) pixel_lp(v, x, false, lp, d);
                acc += (double)lp;
            }
        }
        double v[1] = {acc};
        block_sum<1>(v, ms->red);
        if (threadIdx.x == 0) {
            if (a.ll) a.ll[p] = img ? (float)v[0] : 0.0f;
            if (a.off_page) a.off_page[p] = (uint8_t)(ms->off_page != 0);
        }
        __syncthreads();
    }
}

// ---- forward + adjoint: g_pts (scattered with atomics; zero-filled by the caller), g_eps [P], g_sigma [P] ----
// Cotangent: g_pimg [P][H][W], or g_ll [P] (NULL = 1) on the score against `images`.
__global__ void __launch_bounds__(ANY_THREADS) bpl_render_any_grad_kernel(const KArgs a, const AnyGeom gm) {
    __shared__ AnyShared sh;
    Misc *ms = &sh.ms;
    const int H = gm.H, W = gm.W, N = H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float *base = gm.ws + (size_t)blockIdx.x * ANY_BWD_BUFS * N;
    if (threadIdx.x == 0) { ms->c_eps = ms->c_sigma = nanf(""); }
    __syncthreads();
    for (int p = blockIdx.x; p < a.n_tokens; p += gridDim.x) {
        const TokGeom t = token_geom<true>(a, p);
        const float eps = a.eps[p], sigma = a.sigma[p];
        const bool blur = sigma > 0.0f;
        if (threadIdx.x == 0) ms->off_page = 0;
        float *b0 = base, *b1 = base + N;
        float *V1 = base + 2 * N, *X1 = base + 3 * N, *V2 = base + 4 * N, *X2 = base + 5 * N, *G = base + 6 * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) b0[i] = 0.0f;
        token_consts(eps, sigma, true, ms, a.rc.fhalf);
        __syncthreads();
        any_points(a, t, b0, W, ms);
        __syncthreads();
        // broaden WITHOUT the clamp: In = pre-clamp image (its mask is needed), consumers apply min(., 1) on load
        float *In = b0, *T = b1;
        for (int it = 0; it < a.rc.ink_ncon; ++it) {
            any_broaden(T, In, H, W, a.rc.c1, a.rc.c2, false);
            __syncthreads();
            float *tmp = In; In = T; T = tmp;
        }
        float g[PAD + 1], ht[PAD + 1];
#pragma unroll
        for (int d = 0; d <= PAD; ++d) { g[d] = blur ? ms->g[d] : 0.0f; ht[d] = blur ? ms->ht[d] : 0.0f; }
        if (blur) {
            any_gauss<true, true>(V1, In, H, W, g);
            __syncthreads();
            any_gauss<false, false>(X1, V1, H, W, g);
            __syncthreads();
            any_gauss<true, false>(V2, X1, H, W, g);
            __syncthreads();
            any_gauss<false, false>(X2, V2, H, W, g);
            __syncthreads();
        }
        // ---- pointwise tail and the head of the backward: G = d L / d X2 ----
        const float bgv = eps > 0.0f ? eps : 0.0f;
        const float mixd = eps > 0.0f ? (1.0f - 2.0f * eps) : 1.0f;
        const float lp_bg0 = ms->lp_bg[0], lp_bg1 = ms->lp_bg[1];
        const float dlp_bg0 = ms->dlp_bg[0], dlp_bg1 = ms->dlp_bg[1];
        const float *img = gm.images ? gm.images + (size_t)(a.img_idx ? a.img_idx[p] : 0) * N : nullptr;
        const float *gP_ext = a.g_pimg ? a.g_pimg + (size_t)p * N : nullptr;
        const float gscale = (!gP_ext && a.g_ll) ? a.g_ll[p] : 1.0f;
        float *out = a.pimg ? a.pimg + (size_t)p * N : nullptr;
        double acc_ll = 0.0, acc_eps = 0.0, acc_sig = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const float x2 = blur ? X2[i] : fminf(In[i], 1.0f);
            const bool pass = (x2 >= 0.0f) && (x2 <= 1.0f);                  // clamp_(0,1) passes gradient on [0,1]
            const float i6 = fminf(fmaxf(x2, 0.0f), 1.0f);
            const float pm = any_mix(x2, eps);
            if (out) out[i] = pm;
            float gp = 0.0f;
            if (gP_ext) {
                gp = gP_ext[i];
            } else if (img) {
                const bool x = img[i] != 0.0f;
                float lp = x ? lp_bg1 : lp_bg0, d = x ? dlp_bg1 : dlp_bg0;
                if (pm != bgv) pixel_lp(pm, x, true, lp, d);
                acc_ll += (double)lp;
                gp = d * gscale;
            }
            if (eps > 0.0f) acc_eps += (double)gp * (double)(1.0f - 2.0f * i6);      // d/d eps of (1-eps) v + eps (1-v)
            G[i] = pass ? gp * mixd : 0.0f;
        }
        __syncthreads();
        if (blur) {
            // sigma^3 dK/dsigma = h~ (x) g + g (x) h~ (token_consts_sigma); a zero-padded symmetric pass is its own adjoint
            const float *Xin[2] = {X1, In};      // input of blur 2, blur 1 (the latter through min(., 1))
            const float *Vin[2] = {V2, V1};
            for (int k = 0; k < 2; ++k) {
                // T = g_h * G;  sum G . (h~_h * V)
                for (int i = threadIdx.x; i < N; i += blockDim.x) {
                    const int r = i / W, c = i - r * W;
                    T[i] = any_tap<false, false>(G, r, c, H, W, g);
                    acc_sig += (double)G[i] * (double)any_tap<false, false>(Vin[k], r, c, H, W, ht);
                }
                __syncthreads();
                // G' = g_v * T;  sum T . (h~_v * X)
                for (int i = threadIdx.x; i < N; i += blockDim.x) {
                    const int r = i / W, c = i - r * W;
                    const float hv = (k == 0) ? any_tap<true, false>(Xin[0], r, c, H, W, ht)
                                              : any_tap<true, true>(Xin[1], r, c, H, W, ht);
                    acc_sig += (double)T[i] * (double)hv;
                    G[i] = any_tap<true, false>(T, r, c, H, W, g);
                }
                __syncthreads();
            }
        }
        // clamp_(max=1) (:171) passes gradient where In <= 1
        for (int i = threadIdx.x; i < N; i += blockDim.x) if (!(In[i] <= 1.0f)) G[i] = 0.0f;
        __syncthreads();
        float *Gs = G, *Gd = T;
        for (int it = 0; it < a.rc.ink_ncon; ++it) {
            any_broaden(Gd, Gs, H, W, a.rc.c1, a.rc.c2, false);
            __syncthreads();
            float *tmp = Gs; Gs = Gd; Gd = tmp;
        }
        // ---- add_stroke backward ----
        if (a.g_pts) {
            const LayD lay{W, 0};
            for (int s = warp; s < t.S; s += nw) {
                const int o = a.sub_pt_off[t.s0 + s], n = a.sub_pt_off[t.s0 + s + 1] - o;
                const RawSrc src{reinterpret_cast<const float2 *>(a.pts) + o};
                const Stats st = substroke_stats(src, n, lane, a.rc);
                float *gpt = a.g_pts + 2 * (size_t)o;
                substroke_backward(src, n, lane, a.rc, st, Gs, [&](int j, float gx, float gy) {
                    atomicAdd(gpt + 2 * j, gx);
                    atomicAdd(gpt + 2 * j + 1, gy);
                }, lay);
            }
        }
        double v[3] = {acc_ll, acc_eps, acc_sig};
        block_sum<3>(v, ms->red);
        if (threadIdx.x == 0) {
            if (a.ll) a.ll[p] = (img && !gP_ext) ? (float)v[0] : 0.0f;
            if (a.off_page) a.off_page[p] = (uint8_t)(ms->off_page != 0);
            if (a.g_eps) a.g_eps[p] = (float)v[1];
            if (a.g_sigma) {
                const double s3 = (double)sigma * (double)sigma * (double)sigma;
                a.g_sigma[p] = blur ? (float)(v[2] / s3) : 0.0f;
            }
        }
        __syncthreads();
    }
}

}  // namespace bpl
