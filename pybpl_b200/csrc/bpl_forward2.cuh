// bpl_forward2.cuh -- the scoring hot path: TWO tokens per CTA, their images interleaved as float2.
//
// Every stencil pass is identical for the two tokens except for the weights, so the image buffer holds
// (token0, token1) pairs per pixel and the passes run on Blackwell's packed FP32 instructions
// (fma.rn.f32x2 / add.f32x2 -> SASS FFMA2 / FADD2, sm_100+) with 64-bit shared-memory accesses:
// half the issue slots, loads, stores, address arithmetic and barriers per token.  The per-token work
// that is not data-parallel across the pair (point phase, likelihood tail) stays scalar and simply
// addresses the .x / .y halves.  A lane walking down a column strides 105 float2 = 210 words = 18 mod 32
// banks: sixteen lanes of a 64-bit access hit sixteen distinct even banks, so both orientations of the
// separable passes remain conflict free.
//
// Shared memory: 88 KB image pairs + 1.7 KB side columns + 4 KB basis + 4 KB sub-stroke records + 2 x 5.1 KB bookkeeping = 108 KB,
// two CTAs (= four tokens) per SM, 96 registers per thread.
#pragma once

#include <type_traits>

namespace bpl {

constexpr int F2_THREADS = 320;
constexpr int F2_CTAS_PER_SM = 2;

struct Fwd2Smem {
    float2 A[NPIX + 1];
    float2 side[2 * IMG];
    float tab[MAX_TABLE];
    float4 sub[2 * MAXS];                   // per sub-stroke (offx, offy, sum ink, survivor count); token 1 at +MAXS
    Misc ms[2];
};

__device__ __forceinline__ float2 shfl_up2(float2 v) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float2 shfl_down2(float2 v) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}

// ---- packed in-place 11-tap pass (see gauss_inplace in bpl_forward.cuh for the scheme) ----------------
template <bool VERT>
__device__ __forceinline__ void gauss_inplace2(float2 *buf, const float2 (&g)[PAD + 1]) {
    constexpr int SA = VERT ? IMG : 1;
    constexpr int SL = VERT ? 1 : IMG;
    const int item = threadIdx.x;
    const bool active = item < GNB * IMG;
    const int band = item / IMG;
    const int line = item - band * IMG;
    float2 *p = buf + band * GR * SA + line * SL;
    float2 W[GU + 2 * PAD];
    float2 trail[PAD];
    const float2 z = make_float2(0.0f, 0.0f);
    if (active) {
#pragma unroll
        for (int i = 0; i < PAD; ++i) {
            W[i] = (band > 0) ? p[(i - PAD) * SA] : z;
            trail[i] = (band < GNB - 1) ? p[(GR + i) * SA] : z;
        }
    }
    __syncthreads();
    if (!active) return;
#pragma unroll
    for (int i = 0; i < PAD; ++i) W[PAD + i] = p[i * SA];
    auto block = [&](float2 *q) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float2 acc = __fmul2_rn(g[0], W[u + PAD]);
#pragma unroll
            for (int d = 1; d <= PAD; ++d) acc = __ffma2_rn(g[d], __fadd2_rn(W[u + PAD - d], W[u + PAD + d]), acc);
            q[u * SA] = acc;
        }
    };
#pragma unroll 1
    for (int ob = 0; ob < GR - GU; ob += GU) {
        float2 *q = p + ob * SA;
#pragma unroll
        for (int u = 0; u < GU; ++u) W[2 * PAD + u] = q[(u + PAD) * SA];
        block(q);
#pragma unroll
        for (int k = 0; k < 2 * PAD; ++k) W[k] = W[k + GU];
    }
    float2 *q = p + (GR - GU) * SA;
#pragma unroll
    for (int u = 0; u < GU; ++u) W[2 * PAD + u] = (u + PAD < GU) ? q[(u + PAD) * SA] : trail[u + PAD - GU];
    block(q);
}

// ---- packed in-place broaden (see broaden_inplace) -------------------------------------------------------
__device__ __forceinline__ void broaden_inplace2(float2 *A, float2 *side, const RenderConsts &rc) {
    static_assert(F2_THREADS == 2 * (IMG / BR) * 32, "one warp per (band, half)");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int band = warp >> 1, half = warp & 1;
    const int i2 = half * 32 + lane;
    const int c0 = 2 * i2, c1 = c0 + 1;
    const bool ok0 = c0 < IMG, ok1 = c1 < IMG;
    const int a0 = band * BR;
    const float2 *side63 = side, *side64 = side + IMG;
    const float2 z = make_float2(0.0f, 0.0f), two = make_float2(2.0f, 2.0f);
    const float2 c1w = make_float2(rc.c1, rc.c1), c2w = make_float2(rc.c2, rc.c2);
    for (int it = 0; it < rc.ink_ncon; ++it) {
        for (int i = threadIdx.x; i < 2 * IMG; i += blockDim.x) {
            const int sidx = i / IMG, r = i - sidx * IMG;
            side[i] = A[r * IMG + 63 + sidx];
        }
        float2 hp0 = z, hp1 = z, ht0 = z, ht1 = z;
        auto hdirect = [&](int r, float2 &h0, float2 &h1) {
            const float2 *row = A + r * IMG;
            const float2 xl = (ok0 && c0 > 0) ? row[c0 - 1] : z;
            const float2 x0 = ok0 ? row[c0] : z, x1 = ok1 ? row[c1] : z;
            const float2 xr = (c1 + 1 < IMG) ? row[c1 + 1] : z;
            h0 = __fadd2_rn(__ffma2_rn(two, x0, xl), x1);
            h1 = __fadd2_rn(__ffma2_rn(two, x1, x0), xr);
        };
        if (band > 0) hdirect(a0 - 1, hp0, hp1);
        if (band < IMG / BR - 1) hdirect(a0 + BR, ht0, ht1);
        __syncthreads();
        const bool last = (it == rc.ink_ncon - 1);
        auto hrow = [&](int r, float2 &x0, float2 &x1, float2 &h0, float2 &h1) {
            const float2 *row = A + r * IMG;
            x0 = ok0 ? row[c0] : z;
            x1 = ok1 ? row[c1] : z;
            float2 xl = shfl_up2(x1);
            float2 xr = shfl_down2(x0);
            const float2 s63 = side63[r], s64 = side64[r];
            if (lane == 0) xl = half ? s63 : z;
            if (lane == 31) xr = half ? z : s64;
            h0 = __fadd2_rn(__ffma2_rn(two, x0, xl), x1);
            h1 = __fadd2_rn(__ffma2_rn(two, x1, x0), xr);
        };
        float2 xc0, xc1, hc0, hc1;
        hrow(a0, xc0, xc1, hc0, hc1);
#pragma unroll
        for (int o = 0; o < BR; ++o) {
            float2 xn0 = z, xn1 = z, hn0 = ht0, hn1 = ht1;
            if (o + 1 < BR) hrow(a0 + o + 1, xn0, xn1, hn0, hn1);
            float2 v0 = __ffma2_rn(c1w, __fadd2_rn(__ffma2_rn(two, hc0, hp0), hn0), __fmul2_rn(c2w, xc0));
            float2 v1 = __ffma2_rn(c1w, __fadd2_rn(__ffma2_rn(two, hc1, hp1), hn1), __fmul2_rn(c2w, xc1));
            if (last) {                                              // clamp_(max=1)  (:171)
                v0.x = fminf(v0.x, 1.0f); v0.y = fminf(v0.y, 1.0f);
                v1.x = fminf(v1.x, 1.0f); v1.y = fminf(v1.y, 1.0f);
            }
            float2 *row = A + (a0 + o) * IMG;
            if (ok0) row[c0] = v0;
            if (ok1) row[c1] = v1;
            hp0 = hc0; hp1 = hc1; hc0 = hn0; hc1 = hn1; xc0 = xn0; xc1 = xn1;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(F2_THREADS, F2_CTAS_PER_SM) bpl_forward2_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fwd2Smem &sm = *reinterpret_cast<Fwd2Smem *>(smem_raw);
    float2 *A = sm.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool aff = a.affine != nullptr;

    for (int i = threadIdx.x; i < a.neval * NCPT; i += blockDim.x) sm.tab[i] = a.basis[i];
    if (threadIdx.x < 2) { sm.ms[threadIdx.x].cur_img = -1; sm.ms[threadIdx.x].c_eps = sm.ms[threadIdx.x].c_sigma = nanf(""); }
    int q = blockIdx.x;                 // the pair being processed: the first one is fixed, later ones come from the queue
    if (threadIdx.x == 0) sm.ms[0].next_item = sched_claim(a, q);
    zero_floats(reinterpret_cast<float *>(A), 2 * NPIX);      // afterwards the tail of every pair re-zeroes what it reads
    __syncthreads();
    int q1 = sm.ms[0].next_item;        // the pair after it (claimed one ahead so that its geometry can be prefetched)

#ifdef BPL_PHASE_TIMING
    long long pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pt_acc2[3] = {0, 0, 0};
    long long pt_t0 = clock64();
#endif
    const int npairs = (a.n_tokens + 1) >> 1;
    // The geometry (two dependent global loads), eps and sigma of the NEXT pair are fetched while the current
    // pair is processed, so their latency is off the critical path of the prologue.
    struct PairIn { TokGeom t[2]; float eps[2], sigma[2]; int pt[2]; bool have1; };
    auto fetch = [&](int q) {
        PairIn in;
        const int p0 = 2 * q;
        in.have1 = p0 + 1 < a.n_tokens;
        in.pt[0] = p0; in.pt[1] = in.have1 ? p0 + 1 : p0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            in.t[k] = token_geom<false>(a, in.pt[k]);
            in.eps[k] = a.eps[in.pt[k]];
            in.sigma[k] = a.sigma[in.pt[k]];
        }
        if (!in.have1) { in.t[1].S = 0; in.t[1].K = 0; }
        return in;
    };
    PairIn nxt{};
    if (q < npairs) nxt = fetch(q);
    while (q < npairs) {
        const PairIn cur = nxt;
        int q2 = 0;
        if (threadIdx.x == 0) q2 = sched_claim(a, q1);      // consumed at the end of the iteration
        if (q1 < npairs) nxt = fetch(q1);
        const bool have1 = cur.have1;
        const int pt[2] = {cur.pt[0], cur.pt[1]};
        TokGeom t[2] = {cur.t[0], cur.t[1]};
        bool bad[2];
        float eps[2] = {cur.eps[0], cur.eps[1]}, sigma[2] = {cur.sigma[0], cur.sigma[1]};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            bad[k] = t[k].S > MAXS || t[k].K > MAXK;      // beyond the per-CTA tables: flagged below, rendered empty
            if (bad[k]) { t[k].S = 0; t[k].K = 0; }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!a.all_pairs) load_image_bits(a, pt[k], &sm.ms[k]);
            if (threadIdx.x == 0) sm.ms[k].off_page = 0;
        }
        BPL_PT(0);      // geometry loads, zeroing, constants (no barrier yet)
        // ---- chain offsets for both tokens (one pair of barriers) ----
        for (int s = threadIdx.x; s < t[0].S + t[1].S; s += blockDim.x) {
            const int k = s >= t[0].S, sl = k ? s - t[0].S : s;
            const int gs = t[k].s0 + sl;
            const float inv = a.invscale[gs];
            const float *cp = a.ctrl + (size_t)gs * NCPT * 2;
            float ax = 0.0f, ay = 0.0f, bx = 0.0f, by = 0.0f;
            const float *r0 = sm.tab, *r1 = sm.tab + (a.neval - 1) * NCPT;
#pragma unroll
            for (int c = 0; c < NCPT; ++c) {
                const float yx = __fmul_rn(inv, cp[2 * c]), yy = __fmul_rn(inv, cp[2 * c + 1]);
                ax = __fmaf_rn(r0[c], yx, ax); ay = __fmaf_rn(r0[c], yy, ay);
                bx = __fmaf_rn(r1[c], yx, bx); by = __fmaf_rn(r1[c], yy, by);
            }
            sm.sub[k * MAXS + sl] = make_float4(ax, ay, bx, by);
        }
        __syncthreads();
        for (int s = threadIdx.x; s < t[0].K + t[1].K; s += blockDim.x) {
            const int k = s >= t[0].K, kl = k ? s - t[0].K : s;
            const int gk = t[k].k0 + kl;
            float px = a.position[2 * gk], py = a.position[2 * gk + 1];
            const int b = a.stroke_sub_off[gk] - t[k].s0, e = a.stroke_sub_off[gk + 1] - t[k].s0;
            for (int ss = b; ss < e; ++ss) {
                const float4 v = sm.sub[k * MAXS + ss];
                const float ox = __fsub_rn(v.x, px), oy = __fsub_rn(v.y, py);
                px = __fsub_rn(v.z, ox); py = __fsub_rn(v.w, oy);
                sm.sub[k * MAXS + ss] = make_float4(ox, oy, 0.0f, 0.0f);
            }
        }
        __syncthreads();
        if (aff) {      // rare: affine_setup reads the offsets from Misc::sub
            for (int s = threadIdx.x; s < t[0].S + t[1].S; s += blockDim.x) {
                const int k = s >= t[0].S, sl = k ? s - t[0].S : s;
                sm.ms[k].sub[sl] = sm.sub[k * MAXS + sl];
            }
            __syncthreads();
            affine_setup(a, t[0], pt[0], sm.tab, &sm.ms[0]);
            affine_setup(a, t[1], pt[1], sm.tab, &sm.ms[1]);
        }
        // double-precision constants by the last two warps, hidden behind the point phase of the others
        token_consts(eps[0], sigma[0], false, &sm.ms[0], 0);
        token_consts(eps[1], sigma[1], false, &sm.ms[1], 1);
        if (threadIdx.x < 2 && a.image_bits) sm.ms[threadIdx.x].cur_img = a.img_idx ? a.img_idx[pt[threadIdx.x]] : 0;

        BPL_PT(1);      // chain offsets
        // ---- add_stroke: a thread owns 7 consecutive points of a sub-stroke of either token ----
        const int QN = (a.neval + PPL - 1) / PPL;
        const int it0 = t[0].S * QN, items = it0 + t[1].S * QN;
        bool any0 = false, any1 = false;
        for (int base = 0; base < items; base += blockDim.x) {
            const int item = base + threadIdx.x;
            const bool active = item < items;
            int sid = 0x7fffffff, k = 0, qq = 0;
            CtrlSrc src{};
            if (active) {
                k = item >= it0;
                const int il = k ? item - it0 : item;
                const int s = il / QN;
                qq = il - s * QN;
                sid = k * MAXS + s;
                src = make_ctrl_src_p(a, t[k], s, sm.tab, sm.sub[sid], sm.ms[k].affB, aff);
            }
            const LaneAcc acc = warp_points<false, 2>(active, src, sm.tab, sid, a.neval, qq, lane, a.rc,
                                                      reinterpret_cast<float *>(A) + k, 0.0f, 0);
            if (k) any1 |= acc.any_out; else any0 |= acc.any_out;
#ifdef BPL_PHASE_TIMING
            if (threadIdx.x == 0) { pt_acc2[0] += acc.tA; pt_acc2[1] += acc.tB; pt_acc2[2] += acc.tC; }
#endif
            reduce_by_substroke(active, sid, acc.sum, acc.cnt, sm.sub, lane);
        }
        BPL_PT(6);      // (timing) thread 0's own share of the point phase
        any0 = __any_sync(0xffffffffu, any0);
        any1 = __any_sync(0xffffffffu, any1);
        if (lane == 0) {
            if (any0) atomicOr(&sm.ms[0].off_page, 1);
            if (any1) atomicOr(&sm.ms[1].off_page, 1);
        }
        __syncthreads();
        BPL_PT(7);      // (timing) waiting for the slowest warp of the point phase
        // min-ink branches (rendering.py:104-107): every thread reaches the same verdict from the totals
        bool need_fix = false;
        for (int s = 0; s < t[0].S + t[1].S; ++s) {
            const float4 v = sm.sub[s >= t[0].S ? MAXS + s - t[0].S : s];
            need_fix |= (__float_as_int(v.w) >= 2 && v.z < a.rc.ink_pp);
        }
        if (need_fix) {
            for (int base = 0; base < items; base += blockDim.x) {
                const int item = base + threadIdx.x;
                bool active = item < items;
                int sid = 0x7fffffff, k = 0, qq = 0, s = 0;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) {
                    k = item >= it0;
                    const int il = k ? item - it0 : item;
                    s = il / QN;
                    qq = il - s * QN;
                    sid = k * MAXS + s;
                    v = sm.sub[sid];
                }
                active = active && (__float_as_int(v.w) >= 2 && v.z < a.rc.ink_pp);
                CtrlSrc src{};
                if (active) src = make_ctrl_src_p(a, t[k], s, sm.tab, v, sm.ms[k].affB, aff);
                warp_points<true, 2>(active, src, sm.tab, sid, a.neval, qq, lane, a.rc, reinterpret_cast<float *>(A) + k, v.z,
                                     __float_as_int(v.w));
            }
            __syncthreads();
        }

        BPL_PT(2);      // point phase
        // ---- broaden + clamp ----
        broaden_inplace2(A, sm.side, a.rc);
        if (a.rc.ink_ncon == 0) {
            for (int i = threadIdx.x; i < NPIX; i += blockDim.x) {
                float2 v = A[i];
                A[i] = make_float2(fminf(v.x, 1.0f), fminf(v.y, 1.0f));
            }
            __syncthreads();
        }

        BPL_PT(3);      // broaden
        // ---- blur (a token without blur gets the identity kernel: 1*x + 0*(...) = x exactly) ----
        if (sigma[0] > 0.0f || sigma[1] > 0.0f) {
            float2 g[PAD + 1];
#pragma unroll
            for (int d = 0; d <= PAD; ++d) {
                const float id = (d == 0) ? 1.0f : 0.0f;
                g[d] = make_float2(sigma[0] > 0.0f ? sm.ms[0].g[d] : id, sigma[1] > 0.0f ? sm.ms[1].g[d] : id);
            }
#pragma unroll 1
            for (int rep = 0; rep < 2; ++rep) {
                gauss_inplace2<true>(A, g);
                __syncthreads();
                gauss_inplace2<false>(A, g);
                __syncthreads();
            }
        }

        BPL_PT(4);      // blur
        // ---- clamp, mix, likelihood ----
        float acc[2] = {0.0f, 0.0f};
        {
            const float bg0 = eps[0] > 0.0f ? eps[0] : 0.0f, bg1 = eps[1] > 0.0f ? eps[1] : 0.0f;
            const float om0 = 1.0f - eps[0], om1 = 1.0f - eps[1];
            const float l00 = sm.ms[0].lp_bg[0], l01 = sm.ms[0].lp_bg[1];
            const float l10 = sm.ms[1].lp_bg[0], l11 = sm.ms[1].lp_bg[1];
            const float2 zz = make_float2(0.0f, 0.0f);
            if (!a.all_pairs) {
                for (int i = threadIdx.x; i < NPIX; i += blockDim.x) {
                    const float2 v = A[i];
                    A[i] = zz;                                               // leaves the buffer zeroed for the next pair
                    const bool x0 = (sm.ms[0].bits[i >> 5] >> (i & 31)) & 1u;
                    const bool x1 = (sm.ms[1].bits[i >> 5] >> (i & 31)) & 1u;
                    float u0 = fminf(fmaxf(v.x, 0.0f), 1.0f), u1 = fminf(fmaxf(v.y, 0.0f), 1.0f);      // clamp_(0,1) (:180)
                    if (eps[0] > 0.0f) u0 = __fadd_rn(__fmul_rn(om0, u0), __fmul_rn(eps[0], __fsub_rn(1.0f, u0)));
                    if (eps[1] > 0.0f) u1 = __fadd_rn(__fmul_rn(om1, u1), __fmul_rn(eps[1], __fsub_rn(1.0f, u1)));
                    float lp0 = x0 ? l01 : l00, lp1 = x1 ? l11 : l10;
                    if (u0 != bg0) lp0 = pixel_lp_fwd(u0, x0);
                    if (u1 != bg1) lp1 = pixel_lp_fwd(u1, x1);
                    acc[0] += lp0; acc[1] += lp1;
                }
            } else {
                // every token against every image: the per-pixel scores for x = 0 are summed once (base), the
                // difference delta = lp(x=1) - lp(x=0) replaces the pixel, and each image then only visits its set bits
                for (int i = threadIdx.x; i < NPIX; i += blockDim.x) {
                    const float2 v = A[i];
                    float u0 = fminf(fmaxf(v.x, 0.0f), 1.0f), u1 = fminf(fmaxf(v.y, 0.0f), 1.0f);
                    if (eps[0] > 0.0f) u0 = __fadd_rn(__fmul_rn(om0, u0), __fmul_rn(eps[0], __fsub_rn(1.0f, u0)));
                    if (eps[1] > 0.0f) u1 = __fadd_rn(__fmul_rn(om1, u1), __fmul_rn(eps[1], __fsub_rn(1.0f, u1)));
                    float a0 = l00, b0 = l01, a1 = l10, b1 = l11;
                    if (u0 != bg0) { a0 = pixel_lp_fwd(u0, false); b0 = pixel_lp_fwd(u0, true); }
                    if (u1 != bg1) { a1 = pixel_lp_fwd(u1, false); b1 = pixel_lp_fwd(u1, true); }
                    acc[0] += a0; acc[1] += a1;
                    A[i] = make_float2(b0 - a0, b1 - a1);
                }
            }
        }
        double v[2] = {(double)acc[0], (double)acc[1]};
        block_sum<2>(v, sm.ms[0].red);
        if (!a.all_pairs) {
            if (threadIdx.x < 2) {
                const int k = threadIdx.x;
                if (k == 0 || have1) {
                    if (a.ll) a.ll[pt[k]] = bad[k] ? nanf("") : (float)v[k];
                    if (a.off_page) a.off_page[pt[k]] = bad[k] ? (uint8_t)255 : (uint8_t)(sm.ms[k].off_page != 0);
                }
            }
        } else {
            // block_sum's barriers have made the deltas visible; one warp per image walks the image's set bits
            for (int m = warp; m < a.n_images; m += nw) {
                const uint32_t *bw = a.image_bits + (size_t)m * IMG_WORDS;
                float s0 = 0.0f, s1 = 0.0f;
                for (int wi = lane; wi < IMG_WORDS; wi += 32) {
                    for (unsigned bits = bw[wi]; bits; bits &= bits - 1u) {
                        const int i = wi * 32 + __ffs(bits) - 1;
                        if (i < NPIX) { const float2 d = A[i]; s0 += d.x; s1 += d.y; }
                    }
                }
                s0 = warp_sum(s0); s1 = warp_sum(s1);
                if (lane == 0 && a.ll) {
                    a.ll[(size_t)pt[0] * a.n_images + m] = bad[0] ? nanf("") : (float)(v[0] + (double)s0);
                    if (have1) a.ll[(size_t)pt[1] * a.n_images + m] = bad[1] ? nanf("") : (float)(v[1] + (double)s1);
                }
            }
            if (threadIdx.x < 2 && a.off_page && (threadIdx.x == 0 || have1))
                a.off_page[pt[threadIdx.x]] = bad[threadIdx.x] ? (uint8_t)255 : (uint8_t)(sm.ms[threadIdx.x].off_page != 0);
            __syncthreads();
            zero_floats(reinterpret_cast<float *>(A), 2 * NPIX);      // the deltas must not leak into the next pair
        }
        if (threadIdx.x == 0) sm.ms[0].next_item = q2;
        __syncthreads();
        q = q1;
        q1 = sm.ms[0].next_item;      // rewritten only at the end of the next iteration, many barriers from here
        BPL_PT(5);      // tail + reduction
    }
    if (threadIdx.x == 0) sched_finish(a);
#ifdef BPL_PHASE_TIMING
    if (threadIdx.x == 0 && a.pt_buf)
        for (int i = 0; i < 8; ++i) a.pt_buf[blockIdx.x * 16 + i] = (float)pt_acc[i];
    if (threadIdx.x == 0 && a.pt_buf)
        for (int i = 0; i < 3; ++i) a.pt_buf[blockIdx.x * 16 + 8 + i] = (float)pt_acc2[i];
#endif
}

}  // namespace bpl
