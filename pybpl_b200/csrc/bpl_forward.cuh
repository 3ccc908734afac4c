// bpl_forward.cuh -- the forward render+score kernel (included by bpl_render.cu).
//
// One CTA per token, ONE 105x105 fp32 image buffer in shared memory (44 KB), every stencil pass in
// place, so four CTAs' worth of shared memory fit on an SM and the register file decides occupancy
// (3 CTAs x 320 threads).  Row stride 105 is odd: a warp whose lanes walk down a column (stride 105)
// and a warp whose lanes walk along a row (stride 1) are both bank-conflict free.
//
// In-place separable pass: a work item is (band, line) -- a run of R=35 outputs along the filter
// axis.  Every item first loads the 5+5 halo values it needs from the neighbouring bands into
// registers, the CTA synchronises, and then each item slides an 11-register window down its own run,
// overwriting inputs it has already consumed.  Zero padding (F.conv2d, util/general.py:161-163) is the
// halo of the first/last band being zero.
#pragma once

#include "bpl_points.cuh"

namespace bpl {

constexpr int FWD_THREADS = 320;            // 3 bands x 105 lines = 315 sweep items, one per thread
constexpr int FWD_CTAS_PER_SM = 3;
constexpr int GR = 35;                      // outputs per Gaussian sweep item
constexpr int GNB = IMG / GR;               // 3 bands
constexpr int FBUF = (NPIX + 3) / 4 * 4;    // 11028 floats

struct FwdSmem {
    float A[FBUF];
    float side[2 * IMG + 2];                // copies of columns 63 and 64 for the in-place broaden
    float tab[MAX_TABLE];
    Misc ms;
};

// ---- in-place 11-tap pass ---------------------------------------------------------------------------
// Partially unrolled (blocks of 7 outputs) so that the loop body stays inside the instruction cache:
// three CTAs in different phases share one SM's instruction fetch, and a fully unrolled 35-output
// body per pass thrashed it (ncu: stall_no_instruction was the top stall reason).
constexpr int GU = 7;                       // outputs per unrolled block, GR = 5 * GU

template <bool VERT>
__device__ __forceinline__ void gauss_inplace(float *buf, const float (&g)[PAD + 1]) {
    static_assert(GR % GU == 0 && GU > PAD, "block structure");
    constexpr int SA = VERT ? IMG : 1;
    constexpr int SL = VERT ? 1 : IMG;
    const int item = threadIdx.x;
    const bool active = item < GNB * IMG;
    const int band = item / IMG;
    const int line = item - band * IMG;
    float *p = buf + band * GR * SA + line * SL;
    float W[GU + 2 * PAD];      // W[0..9]: the ten inputs before the block's first new one; W[10..16]: new inputs
    float trail[PAD];
    if (active) {
#pragma unroll
        for (int i = 0; i < PAD; ++i) {
            W[i] = (band > 0) ? p[(i - PAD) * SA] : 0.0f;
            trail[i] = (band < GNB - 1) ? p[(GR + i) * SA] : 0.0f;
        }
    }
    __syncthreads();        // every halo is in registers: runs may now be overwritten
    if (!active) return;
#pragma unroll
    for (int i = 0; i < PAD; ++i) W[PAD + i] = p[i * SA];
    auto block = [&](float *q) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float acc = g[0] * W[u + PAD];
#pragma unroll
            for (int d = 1; d <= PAD; ++d) acc = fmaf(g[d], W[u + PAD - d] + W[u + PAD + d], acc);
            q[u * SA] = acc;
        }
    };
#pragma unroll 1
    for (int ob = 0; ob < GR - GU; ob += GU) {
        float *q = p + ob * SA;
#pragma unroll
        for (int u = 0; u < GU; ++u) W[2 * PAD + u] = q[(u + PAD) * SA];
        block(q);
#pragma unroll
        for (int k = 0; k < 2 * PAD; ++k) W[k] = W[k + GU];
    }
    float *q = p + (GR - GU) * SA;
#pragma unroll
    for (int u = 0; u < GU; ++u) W[2 * PAD + u] = (u + PAD < GU) ? q[(u + PAD) * SA] : trail[u + PAD - GU];
    block(q);
}

// ---- in-place 3x3 broaden (rendering.py:160-168) --------------------------------------------------
// out = c1 * (1,2,1)x(1,2,1) * in + c2 * in.  A thread owns TWO adjacent columns of a 21-row band
// (5 bands x 2 warps = all 10 warps); the horizontal neighbours of its column pair come from the
// adjacent lanes by shuffle, so nothing inside a warp is read from shared memory after a neighbour may
// have overwritten it.  The single warp boundary (columns 63|64) is served from `side`, a copy taken
// before the pass; band halo rows are preloaded before the barrier.
constexpr int BR = 21;                      // rows per broaden band
constexpr int NSIDE = 2;                    // copies of columns 63 and 64

template <bool CLAMP>
__device__ __forceinline__ void broaden_inplace(float *A, float *side, const RenderConsts &rc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int band = warp >> 1, half = warp & 1;
    const int i2 = half * 32 + lane;            // column pair index
    const int c0 = 2 * i2, c1 = c0 + 1;
    const bool ok0 = c0 < IMG, ok1 = c1 < IMG;
    const int a0 = band * BR;
    const float *side63 = side, *side64 = side + IMG;
    const float c1w = rc.c1, c2w = rc.c2;
    for (int it = 0; it < rc.ink_ncon; ++it) {
        for (int i = threadIdx.x; i < NSIDE * IMG; i += blockDim.x) {
            const int sidx = i / IMG, r = i - sidx * IMG;
            side[i] = A[r * IMG + 63 + sidx];
        }
        // halo rows (a0-1 and a0+BR): horizontal sums straight from the image (nothing is being written yet)
        float hp0 = 0.0f, hp1 = 0.0f, ht0 = 0.0f, ht1 = 0.0f;
        auto hdirect = [&](int r, float &h0, float &h1) {
            const float *row = A + r * IMG;
            const float xl = (ok0 && c0 > 0) ? row[c0 - 1] : 0.0f;
            const float x0 = ok0 ? row[c0] : 0.0f, x1 = ok1 ? row[c1] : 0.0f;
            const float xr = (c1 + 1 < IMG) ? row[c1 + 1] : 0.0f;
            h0 = xl + 2.0f * x0 + x1;
            h1 = x0 + 2.0f * x1 + xr;
        };
        if (band > 0) hdirect(a0 - 1, hp0, hp1);
        if (band < IMG / BR - 1) hdirect(a0 + BR, ht0, ht1);
        __syncthreads();
        const bool last = CLAMP && (it == rc.ink_ncon - 1);
        auto hrow = [&](int r, float &x0, float &x1, float &h0, float &h1) {
            const float *row = A + r * IMG;
            x0 = ok0 ? row[c0] : 0.0f;
            x1 = ok1 ? row[c1] : 0.0f;
            float xl = __shfl_up_sync(0xffffffffu, x1, 1);
            float xr = __shfl_down_sync(0xffffffffu, x0, 1);
            const float s63 = side63[r], s64 = side64[r];       // broadcast loads
            if (lane == 0) xl = half ? s63 : 0.0f;
            if (lane == 31) xr = half ? 0.0f : s64;
            h0 = xl + 2.0f * x0 + x1;
            h1 = x0 + 2.0f * x1 + xr;
        };
        float xc0, xc1, hc0, hc1;
        hrow(a0, xc0, xc1, hc0, hc1);
#pragma unroll
        for (int o = 0; o < BR; ++o) {
            float xn0 = 0.0f, xn1 = 0.0f, hn0 = ht0, hn1 = ht1;
            if (o + 1 < BR) hrow(a0 + o + 1, xn0, xn1, hn0, hn1);
            float v0 = fmaf(c1w, hp0 + 2.0f * hc0 + hn0, c2w * xc0);
            float v1 = fmaf(c1w, hp1 + 2.0f * hc1 + hn1, c2w * xc1);
            if (last) { v0 = fminf(v0, 1.0f); v1 = fminf(v1, 1.0f); }      // clamp_(max=1)  (:171)
            float *row = A + (a0 + o) * IMG;
            if (ok0) row[c0] = v0;
            if (ok1) row[c1] = v1;
            hp0 = hc0; hp1 = hc1; hc0 = hn0; hc1 = hn1; xc0 = xn0; xc1 = xn1;
        }
        __syncthreads();
    }
    if (CLAMP && rc.ink_ncon == 0) {
        for (int i = threadIdx.x; i < NPIX; i += blockDim.x) A[i] = fminf(A[i], 1.0f);
        __syncthreads();
    }
}

template <bool RAW>
__global__ void __launch_bounds__(FWD_THREADS, FWD_CTAS_PER_SM) bpl_forward_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem &sm = *reinterpret_cast<FwdSmem *>(smem_raw);
    float *A = sm.A;
    Misc *ms = &sm.ms;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;

    if (!RAW)
        for (int i = threadIdx.x; i < a.neval * NCPT; i += blockDim.x) sm.tab[i] = a.basis[i];
    if (threadIdx.x == 0) { ms->cur_img = -1; ms->c_eps = ms->c_sigma = nanf(""); }
    __syncthreads();

    int item = blockIdx.x;                                      // position in the work queue (advanced by thread 0 only)
    int p = (!RAW && a.order) ? a.order[item] : item;           // the token it stands for
    while (p < a.n_tokens) {
        int p_next = 0;
        if (threadIdx.x == 0) p_next = sched_next_token(a, item);      // consumed at the end of the iteration
        const TokGeom t = token_geom<RAW>(a, p);
        if (!RAW && (t.S > MAXS || t.K > MAXK)) {      // beyond the per-CTA tables: flag, do not guess
            if (threadIdx.x == 0) {
                if (a.ll) a.ll[p] = nanf("");
                if (a.off_page) a.off_page[p] = 255;
                ms->next_item = p_next;
            }
            __syncthreads();
            p = ms->next_item;
            __syncthreads();
            continue;
        }
        const float eps = a.eps[p], sigma = a.sigma[p];
        const bool aff = !RAW && a.affine != nullptr;
        zero_floats(A, FBUF);
        load_image_bits(a, p, ms);
        token_consts(eps, sigma, false, ms, a.rc.fhalf);
        if (threadIdx.x == 0) ms->off_page = 0;
        if (!RAW) {
            chain_offsets(a, t, sm.tab, ms);
            if (aff) affine_setup(a, t, p, sm.tab, ms);
        } else {
            __syncthreads();
        }
        if (threadIdx.x == 0 && a.image_bits) ms->cur_img = a.img_idx ? a.img_idx[p] : 0;

        // ---- add_stroke ----
        bool any_out = false;
        if (RAW) {          // ragged raw trajectories: one warp per stroke, two passes (not the hot path)
            for (int s = warp; s < t.S; s += nw) {
                const int o = a.sub_pt_off[t.s0 + s], n = a.sub_pt_off[t.s0 + s + 1] - o;
                const RawSrc src{reinterpret_cast<const float2 *>(a.pts) + o};
                const Stats st = substroke_stats(src, n, lane, a.rc);
                any_out |= st.any_out;
                if (st.cnt > 0) substroke_splat<0, IMG>(src, n, lane, a.rc, st, A);
            }
            any_out = __any_sync(0xffffffffu, any_out);
            if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
            __syncthreads();
        } else {
            const int QN = (a.neval + PPL - 1) / PPL, items = t.S * QN;
            for (int base = 0; base < items; base += blockDim.x) {
                const int item = base + threadIdx.x;
                const bool active = item < items;
                const int s = active ? item / QN : 0x7fffffff;
                CtrlSrc src{};
                if (active) src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
                const LaneAcc acc = warp_points<false, 1>(active, src, sm.tab, s, a.neval, active ? item - s * QN : 0, lane,
                                                          a.rc, A, 0.0f, 0);
                any_out |= acc.any_out;
                reduce_by_substroke(active, s, acc.sum, acc.cnt, ms->sub, ms->ink64, lane);
            }
            any_out = __any_sync(0xffffffffu, any_out);
            if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
            __syncthreads();
            // min-ink branches (rendering.py:104-107): every thread reaches the same verdict from the totals
            bool need_fix = false;
            for (int s = 0; s < t.S; ++s) {
                float4 v = ms->sub[s];
                v.z = ink_sum(ms, s);
                need_fix |= (__float_as_int(v.w) >= 2 && v.z < a.rc.ink_pp);
            }
            if (need_fix) {
                for (int base = 0; base < items; base += blockDim.x) {
                    const int item = base + threadIdx.x;
                    bool active = item < items;
                    const int s = active ? item / QN : 0x7fffffff;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (active) v = ms->sub[s];
                    active = active && (__float_as_int(v.w) >= 2 && v.z < a.rc.ink_pp);
                    CtrlSrc src{};
                    if (active) src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
                    // inactive lanes keep their sub-stroke id so that the segmented scan still sees run order
                    warp_points<true, 1>(active, src, sm.tab, s, a.neval, item < items ? item - s * QN : 0, lane, a.rc, A, v.z,
                                         __float_as_int(v.w));
                }
                __syncthreads();
            }
        }

        // ---- broaden + clamp ----
        broaden_inplace<true>(A, sm.side, a.rc);

        // ---- blur, clamp, mix, likelihood ----
        const float bgv = eps > 0.0f ? eps : 0.0f;
        const float om = 1.0f - eps;
        const float lp_bg0 = ms->lp_bg[0], lp_bg1 = ms->lp_bg[1];
        const bool score = a.image_bits != nullptr;
        const bool want_pimg = a.pimg != nullptr;
        float acc_ll = 0.0f;
        auto finish = [&](float v, bool x, float *dst) {
            v = fminf(fmaxf(v, 0.0f), 1.0f);                               // clamp_(0,1)  (:180)
            if (eps > 0.0f) v = __fadd_rn(__fmul_rn(om, v), __fmul_rn(eps, __fsub_rn(1.0f, v)));   // (:227)
            if (want_pimg) *dst = v;
            if (score) {
                if (v == bgv) {
                    acc_ll += x ? lp_bg1 : lp_bg0;
                } else {
                    float lp, d;
                    pixel_lp(v, x, false, lp, d);
                    acc_ll += lp;
                }
            }
        };
        if (sigma > 0.0f) {
            float g[PAD + 1];
#pragma unroll
            for (int d = 0; d <= PAD; ++d) g[d] = ms->g[d];
#pragma unroll 1
            for (int rep = 0; rep < 2; ++rep) {          // imfilter twice (rendering.py:176-177); one code copy
                gauss_inplace<true>(A, g);
                __syncthreads();
                gauss_inplace<false>(A, g);
                __syncthreads();
            }
        }
        // pointwise tail: lanes on consecutive pixels, the 32 target bits of a warp are one broadcast word
        if (score && !want_pimg && eps > 0.0f) {        // the scoring hot path, conditions hoisted
            for (int i = threadIdx.x; i < NPIX; i += blockDim.x) {
                const bool x = (ms->bits[i >> 5] >> (i & 31)) & 1u;
                float v = fminf(fmaxf(A[i], 0.0f), 1.0f);
                v = __fadd_rn(__fmul_rn(om, v), __fmul_rn(eps, __fsub_rn(1.0f, v)));
                float lp = x ? lp_bg1 : lp_bg0, d;
                if (v != bgv) pixel_lp(v, x, false, lp, d);
                acc_ll += lp;
            }
        } else {
            for (int i = threadIdx.x; i < NPIX; i += blockDim.x) {
                const bool x = score && ((ms->bits[i >> 5] >> (i & 31)) & 1u);
                finish(A[i], x, A + i);
            }
        }
        double v[1] = {(double)acc_ll};
        block_sum<1>(v, ms->red);      // contains the barriers that order the pimg stores below
        if (threadIdx.x == 0) {
            if (a.ll) a.ll[p] = score ? (float)v[0] : 0.0f;
            if (a.off_page) a.off_page[p] = (uint8_t)(ms->off_page != 0);
        }
        if (want_pimg) {
            float *dst = a.pimg + (size_t)p * NPIX;
            for (int i = threadIdx.x; i < NPIX; i += blockDim.x) dst[i] = A[i];
        }
        if (threadIdx.x == 0) ms->next_item = p_next;
        __syncthreads();
        p = ms->next_item;      // rewritten only at the end of the next iteration, many barriers from here
    }
    if (threadIdx.x == 0) sched_finish(a);
}

}  // namespace bpl
