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

namespace bpl {

constexpr int FWD_THREADS = 320;            // 3 bands x 105 lines = 315 sweep items, one per thread
constexpr int FWD_CTAS_PER_SM = 3;
constexpr int GR = 35;                      // outputs per Gaussian sweep item
constexpr int GNB = IMG / GR;               // 3 bands
constexpr int FBUF = (NPIX + 3) / 4 * 4;    // 11028 floats
constexpr int BR_SPLIT = 53;                // broaden: band 0 = rows [0,53), band 1 = rows [53,105)
constexpr int NSIDE = 6;                    // warp-boundary columns 31|32, 63|64, 95|96

struct FwdSmem {
    float A[FBUF];
    float side[NSIDE * IMG];                // copies of the warp-boundary columns for the in-place broaden
    float tab[MAX_TABLE];
    Misc ms;
};

// ---- in-place 11-tap pass ---------------------------------------------------------------------------
// epi(r, c, value, ptr): consumes the output for pixel (r,c); ptr is its location in the buffer.
template <bool VERT, class Epi>
__device__ __forceinline__ void gauss_inplace(float *buf, const float (&g)[PAD + 1], Epi epi) {
    constexpr int SA = VERT ? IMG : 1;
    constexpr int SL = VERT ? 1 : IMG;
    const int item = threadIdx.x;
    const bool active = item < GNB * IMG;
    const int band = item / IMG;
    const int line = item - band * IMG;
    const int a0 = band * GR;
    float *p = buf + a0 * SA + line * SL;
    float w[BPL_FSIZE];     // sliding window: w[k] = in[a0 + o - 5 + k]
    float trail[PAD];
    if (active) {
#pragma unroll
        for (int i = 0; i < PAD; ++i) {
            w[i] = (band > 0) ? p[(i - PAD) * SA] : 0.0f;
            trail[i] = (band < GNB - 1) ? p[(GR + i) * SA] : 0.0f;
        }
    }
    __syncthreads();        // every halo is in registers: runs may now be overwritten
    if (active) {
#pragma unroll
        for (int i = 0; i < PAD; ++i) w[PAD + i] = p[i * SA];
#pragma unroll
        for (int o = 0; o < GR; ++o) {
            w[2 * PAD] = (o + PAD < GR) ? p[(o + PAD) * SA] : trail[o + PAD - GR];
            float acc = g[0] * w[PAD];
#pragma unroll
            for (int d = 1; d <= PAD; ++d) acc = fmaf(g[d], w[PAD - d] + w[PAD + d], acc);
            if (VERT) epi(a0 + o, line, acc, p + o * SA); else epi(line, a0 + o, acc, p + o * SA);
#pragma unroll
            for (int k = 0; k < 2 * PAD; ++k) w[k] = w[k + 1];
        }
    }
}

// ---- in-place 3x3 broaden (rendering.py:160-168) --------------------------------------------------
// out = c1 * (1,2,1)x(1,2,1) * in + c2 * in.  Thread = one column of one row band; horizontal
// neighbours come from the adjacent lanes by shuffle (no shared-memory hazard inside a warp) and, at
// warp boundaries, from `side` (copied before the pass).  Two bands x four warps.
template <bool CLAMP>
__device__ __forceinline__ void broaden_inplace(float *A, float *side, const RenderConsts &rc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_pass = warp < 8;
    const int band = warp >> 2, wq = warp & 3;
    const int c = wq * 32 + lane;
    const bool col_ok = in_pass && c < IMG;
    const int r0 = band ? BR_SPLIT : 0, r1 = band ? IMG : BR_SPLIT;
    // side column slots: for warp wq, left neighbour column (32wq-1) is slot 2wq-2, right (32wq+32) is slot 2wq+1
    const float *sl = side + (wq > 0 ? 2 * wq - 2 : 0) * IMG, *sr = side + (wq < 3 ? 2 * wq + 1 : 0) * IMG;
    for (int it = 0; it < rc.ink_ncon; ++it) {
        // (a) copy warp-boundary columns; preload the halo rows' horizontal sums straight from the image
        for (int i = threadIdx.x; i < NSIDE * IMG; i += blockDim.x) {
            const int s = i / IMG, r = i - s * IMG;
            side[i] = A[r * IMG + ((s >> 1) * 32 + 31 + (s & 1))];      // slots 0..5 -> columns 31,32,63,64,95,96
        }
        float h_prev = 0.0f, h_cur = 0.0f, x_cur = 0.0f, h_trail = 0.0f;
        auto hsum_direct = [&](int r) -> float {
            const float *row = A + r * IMG;
            const float xl = (c > 0) ? row[c - 1] : 0.0f, xr = (c < IMG - 1) ? row[c + 1] : 0.0f;
            return xl + 2.0f * row[c] + xr;
        };
        if (col_ok) {
            if (r0 > 0) h_prev = hsum_direct(r0 - 1);
            if (r1 < IMG) h_trail = hsum_direct(r1);
        }
        __syncthreads();
        if (in_pass) {
            auto hsum = [&](int r, float &x) -> float {
                x = col_ok ? A[r * IMG + c] : 0.0f;
                float xl = __shfl_up_sync(0xffffffffu, x, 1);
                float xr = __shfl_down_sync(0xffffffffu, x, 1);
                if (lane == 0) xl = (wq > 0) ? sl[r] : 0.0f;
                if (lane == 31) xr = (wq < 3) ? sr[r] : 0.0f;      // c = 104 is lane 8 of warp 3: its right lane holds 0
                return xl + 2.0f * x + xr;
            };
            const bool last = CLAMP && (it == rc.ink_ncon - 1);
            h_cur = hsum(r0, x_cur);
            for (int r = r0; r < r1; ++r) {
                float x_next = 0.0f;
                const float h_next = (r + 1 < r1) ? hsum(r + 1, x_next) : h_trail;
                float v = fmaf(rc.c1, h_prev + 2.0f * h_cur + h_next, rc.c2 * x_cur);
                if (last) v = fminf(v, 1.0f);                          // clamp_(max=1)  (:171)
                if (col_ok) A[r * IMG + c] = v;
                h_prev = h_cur; h_cur = h_next; x_cur = x_next;
            }
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
    if (threadIdx.x == 0) ms->cur_img = -1;
    __syncthreads();

    for (int p = blockIdx.x; p < a.n_tokens; p += gridDim.x) {
        const TokGeom t = token_geom<RAW>(a, p);
        if (!RAW && (t.S > MAXS || t.K > MAXK)) {      // beyond the per-CTA tables: flag, do not guess
            if (threadIdx.x == 0) {
                if (a.ll) a.ll[p] = nanf("");
                if (a.off_page) a.off_page[p] = 255;
            }
            continue;
        }
        const float eps = a.eps[p], sigma = a.sigma[p];
        const bool aff = !RAW && a.affine != nullptr;
        zero_floats(A, FBUF);
        load_image_bits(a, p, ms);
        token_consts(eps, sigma, false, ms);
        if (threadIdx.x == 0) ms->off_page = 0;
        if (!RAW) {
            chain_offsets(a, t, sm.tab, ms);
            if (aff) affine_setup(a, t, p, sm.tab, ms);
        } else {
            __syncthreads();
        }
        if (threadIdx.x == 0 && a.image_bits) ms->cur_img = a.img_idx ? a.img_idx[p] : 0;

        // ---- add_stroke over sub-strokes: one warp each ----
        bool any_out = false;
        for (int s = warp; s < t.S; s += nw) {
            if (RAW) {
                const int o = a.sub_pt_off[t.s0 + s], n = a.sub_pt_off[t.s0 + s + 1] - o;
                const RawSrc src{reinterpret_cast<const float2 *>(a.pts) + o};
                const Stats st = substroke_stats(src, n, lane, a.rc);
                any_out |= st.any_out;
                if (st.cnt > 0) substroke_splat<0, IMG>(src, n, lane, a.rc, st, A);
            } else {
                const CtrlSrc src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
                const Stats st = substroke_stats(src, a.neval, lane, a.rc);
                any_out |= st.any_out;
                if (st.cnt > 0) substroke_splat<0, IMG>(src, a.neval, lane, a.rc, st, A);
            }
        }
        if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
        __syncthreads();

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
            auto store = [&](int, int, float v, float *q) { *q = v; };
            gauss_inplace<true>(A, g, store);
            __syncthreads();
            gauss_inplace<false>(A, g, store);
            __syncthreads();
            gauss_inplace<true>(A, g, store);
            __syncthreads();
            // last pass: thread = (row, 35-column band); its target bits sit in two 32-bit words
            const int item = threadIdx.x, band = item / IMG, row = item - band * IMG, c0 = band * GR;
            unsigned long long xbits = 0ull;
            if (score && item < GNB * IMG) {
                const uint32_t *bw = ms->bits + row * IMG_WORDS + (c0 >> 5);
                const unsigned long long lo = bw[0], hi = ((c0 >> 5) + 1 < IMG_WORDS) ? bw[1] : 0u;
                xbits = (lo | (hi << 32)) >> (c0 & 31);
            }
            gauss_inplace<false>(A, g, [&](int, int c, float v, float *q) {
                finish(v, (xbits >> (c - c0)) & 1ull, q); });
        } else {
            for (int i = threadIdx.x; i < NPIX; i += blockDim.x) {
                const int r = i / IMG, c = i - r * IMG;
                const bool x = score && ((ms->bits[r * IMG_WORDS + (c >> 5)] >> (c & 31)) & 1u);
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
        __syncthreads();
    }
}

}  // namespace bpl
