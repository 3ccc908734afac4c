// bpl_forward1s.cuh -- scoring kernel, one token per CTA, work confined to the token's ink bounding box.
//
// A token's ink touches about a third of the 105x105 canvas (34 % including the 12-pixel reach of two
// broaden and two blur passes, measured on the bench workload); everything outside is exactly zero
// before the epsilon mix and exactly epsilon after it.  The point phase records the bounding box of the
// splatted pixels; every later phase maps its threads onto the lines and rows that can be non-zero:
//   broaden    the warps split the box's rows, the lanes hold column pairs of the box (broaden_box)
//   Gaussian   ONE run per line of the box over the 7-pixel blocks that hold its rows: no halo exchange, no barrier
//              inside the pass (gauss_run; round 1's banded variant with halo exchange is in the history: 631856b)
//   tail       ll = [score of an all-background render: lp_bg[1] * ones(target) + lp_bg[0] * zeros(target)]
//              + sum over the box's pixels of (lp - background lp), the box's pixels flattened over the CTA
// The box grows by 1 per broaden pass and by 5 along the filtered axis per Gaussian pass.
// Scalar FP32, 64 registers, 55 KB shared memory; 256 threads x 4 CTAs per SM from 2 048 tokens on, 320 x 3 below.
#pragma once

#ifndef BPL_S_THREADS
#define BPL_S_THREADS 320
#endif
#ifndef BPL_S_CTAS
#define BPL_S_CTAS 3
#endif
#ifndef BPL_S_THREADS_LARGE
#define BPL_S_THREADS_LARGE 256
#endif
#ifndef BPL_S_CTAS_LARGE
#define BPL_S_CTAS_LARGE 4
#endif
#ifndef BPL_S_LARGE_MIN
#define BPL_S_LARGE_MIN 2048
#endif

#ifndef BPL_BROADEN_UNROLL
#define BPL_BROADEN_UNROLL 2      // rows per trip of the broaden loop (A/B: 3 and 6 remove the register rotation)
#endif
#ifndef BPL_GAUSS_UNROLL
#define BPL_GAUSS_UNROLL 1        // 7-blocks per trip of the Gaussian run loop
#endif

namespace bpl {

constexpr int BROADEN_UNROLL = BPL_BROADEN_UNROLL, GAUSS_UNROLL = BPL_GAUSS_UNROLL;
#ifdef BPL_S_UNROLL_A
constexpr bool S_ROLL_A = false;      // A/B: phase A of the point stage unrolled, as in the backward kernels
#else
constexpr bool S_ROLL_A = true;
#endif
constexpr int S_THREADS = BPL_S_THREADS;
constexpr int S_CTAS_PER_SM = BPL_S_CTAS;
constexpr int S_THREADS_LARGE = BPL_S_THREADS_LARGE;      // launch shape for batches of S_LARGE_MIN tokens and more
constexpr int S_CTAS_LARGE = BPL_S_CTAS_LARGE;
constexpr int S_LARGE_MIN = BPL_S_LARGE_MIN;

struct BBox {
    int r0, r1, c0, c1;     // inclusive; empty when r0 > r1
    __device__ __forceinline__ bool empty() const { return r0 > r1; }
    __device__ __forceinline__ void grow_rows(int d) { if (!empty()) { r0 = max(0, r0 - d); r1 = min(IMG - 1, r1 + d); } }
    __device__ __forceinline__ void grow_cols(int d) { if (!empty()) { c0 = max(0, c0 - d); c1 = min(IMG - 1, c1 + d); } }
};

struct G6 { float v[PAD + 1]; };      // the half kernel by value: the pass is a real (non-inlined) function so that
                                       // its instantiations do not add to the register pressure of the kernel body
#ifdef BPL_GAUSS_INLINE
#define BPL_GAUSS_ATTR __forceinline__
#else
#define BPL_GAUSS_ATTR __noinline__
#endif

// One run per line, exactly the 7-blocks that hold rows [a0, a1]: nobody else touches the line, so the pass needs no
// halo exchange and no barrier of its own (the in-place window reads ahead of what it writes).  Only `nl` threads work
// (a third of the CTA), but with four CTAs on the SM the others' warps fill the issue slots, and the saved barrier and
// halo loads are worth +4 % at 4 096 tokens, +6 % on 131 072 (19.1 -> 20.3 M particles/s).  Measured alternatives under
// the same launch shape: round 1's banded variants (35-row runs, halos taken before a barrier), the
// box in two groups of whole lines (-6 %), and the same idea for the broaden -- one warp for all rows of a box up to 64
// columns wide, no barriers between its passes -- which was 11 % slower (too serial: 2 x 45 rows by one warp).
template <bool VERT>
__device__ BPL_GAUSS_ATTR void gauss_run(float *buf, const G6 gg, int l0, int nl, int a0, int a1) {
    const float (&g)[PAD + 1] = gg.v;
    constexpr int SA = VERT ? IMG : 1;
    constexpr int SL = VERT ? 1 : IMG;
    constexpr int GU = 7;
    if ((int)threadIdx.x >= nl) return;
    const int s0 = a0 / GU * GU, nblk = a1 / GU - a0 / GU + 1;
    float *p = buf + s0 * SA + (l0 + (int)threadIdx.x) * SL;
    float W[GU + 2 * PAD];
#pragma unroll
    for (int i = 0; i < PAD; ++i) {
        W[i] = (s0 > 0) ? p[(i - PAD) * SA] : 0.0f;
        W[PAD + i] = p[i * SA];
    }
#ifndef BPL_GAUSS_PREDICATED
    // Only the canvas's LAST 7-block reads beyond the canvas (its new inputs are positions 103..109, of which two
    // exist), so it is peeled off: the loop's loads need no bounds predicate (7 compares and 7 zero-moves per block, a
    // ninth of the loop's instructions), and the peeled block knows at compile time which inputs are zero.
    static_assert(IMG % GU == 0, "block structure");
    constexpr int BLAST = IMG / GU - 1, NV = IMG - (BLAST * GU + PAD);      // 14; 2
    auto block = [&](float *q) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float acc = g[0] * W[u + PAD];
#pragma unroll
            for (int d = 1; d <= PAD; ++d) acc = fmaf(g[d], W[u + PAD - d] + W[u + PAD + d], acc);
            q[u * SA] = acc;
        }
#pragma unroll
        for (int k = 0; k < 2 * PAD; ++k) W[k] = W[k + GU];
    };
    const bool to_end = (a1 / GU == BLAST);
    const int nfull = nblk - (to_end ? 1 : 0);
#pragma unroll GAUSS_UNROLL
    for (int blk = 0; blk < nfull; ++blk) {
        float *q = p + blk * GU * SA;
#pragma unroll
        for (int u = 0; u < GU; ++u) W[2 * PAD + u] = q[(u + PAD) * SA];
        block(q);
    }
    if (to_end) {
        float *q = p + nfull * GU * SA;
#pragma unroll
        for (int u = 0; u < GU; ++u) W[2 * PAD + u] = (u < NV) ? q[(u + PAD) * SA] : 0.0f;
        block(q);
    }
#else
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        float *q = p + blk * GU * SA;
        const int pos = s0 + blk * GU + PAD;          // canvas position of the first new input
#pragma unroll
        for (int u = 0; u < GU; ++u) W[2 * PAD + u] = (pos + u < IMG) ? q[(u + PAD) * SA] : 0.0f;
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float acc = g[0] * W[u + PAD];
#pragma unroll
            for (int d = 1; d <= PAD; ++d) acc = fmaf(g[d], W[u + PAD - d] + W[u + PAD + d], acc);
            q[u * SA] = acc;
        }
#pragma unroll
        for (int k = 0; k < 2 * PAD; ++k) W[k] = W[k + GU];
    }
#endif
}

template <bool VERT, int S_THREADS>
__device__ __forceinline__ void gauss_box(float *buf, const G6 &g, int l0, int l1, int a0, int a1) {
    gauss_run<VERT>(buf, g, l0, l1 - l0 + 1, a0, a1);
}

// In-place 3x3 broaden confined to the box.  The lanes of a warp hold column pairs of the box, the warps split
// the box's rows evenly: a box up to 64 columns wide takes one warp per row band (the columns left and right
// of the warp's span are zero), a wider one two (the columns either side of the boundary between them are
// served from `side`, a copy taken before anything is overwritten).  Halo rows are summed before the
// barrier; horizontal neighbours travel by shuffle, so nothing is read after a neighbour may have written it.
// bb is the box of non-zero INPUT pixels; the caller grows it by ink_ncon afterwards.
#ifdef BPL_BROADEN_INLINE
#define BPL_BROADEN_ATTR __forceinline__
#else
#define BPL_BROADEN_ATTR __noinline__
#endif
template <int S_THREADS>
__device__ BPL_BROADEN_ATTR void broaden_box(float *A, float *side, const RenderConsts rc, BBox bb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nw = S_THREADS / 32;             // (compile time: the divisions below become multiplications)
    const float c1w = rc.c1, c2w = rc.c2;
    float *sideL = side, *sideR = side + IMG;      // columns cb-1 and cb, cb = first column of the second half
    for (int it = 0; it < rc.ink_ncon; ++it) {
        bb.grow_rows(1); bb.grow_cols(1);                  // now the OUTPUT box of this pass
        const int cs = bb.c0 & ~1;
        const int npair = (bb.c1 - cs) / 2 + 1;
        const int nhalf = npair > 32 ? 2 : 1;
        const int nrows = bb.r1 - bb.r0 + 1;
        const int nbands = nhalf == 2 ? nw / 2 : nw;
        const int h = nhalf == 2 ? (nrows + nw / 2 - 1) / (nw / 2) : (nrows + nw - 1) / nw;
        const int band = nhalf == 2 ? warp >> 1 : warp, half = nhalf == 2 ? (warp & 1) : 0;
        const int a0 = bb.r0 + band * h, a1 = min(a0 + h - 1, bb.r1);
        const bool live = band < nbands && a0 <= a1;
        const int cb = cs + 64;
        const int pi = half * 32 + lane;
        const int c0 = cs + 2 * pi, c1 = c0 + 1;
        const bool ok0 = pi < npair && c0 < IMG, ok1 = pi < npair && c1 < IMG;
        const bool last = (it == rc.ink_ncon - 1);
        if (nhalf == 2)
            for (int i = threadIdx.x; i < 2 * nrows; i += blockDim.x) {
                const int sidx = i >= nrows, r = bb.r0 + i - sidx * nrows;
                side[sidx * IMG + r] = (cb - 1 + sidx < IMG) ? A[r * IMG + cb - 1 + sidx] : 0.0f;
            }
        // horizontal (1,2,1) sums of a row.  DIRECT: before the barrier the boundary neighbours are still intact in A
#ifndef BPL_BROADEN_BRANCHY
        // What lies left of lane 0 / right of lane 31 comes through a pointer chosen once per pass -- the side copies for
        // the inner edge of a two-warp row, a column of zeros otherwise -- so the row loop has two predicated loads
        // instead of a nest of lane- and shape-dependent branches (the nest was a quarter of the loop's instructions).
        const bool inL = nhalf == 2 && half == 1, inR = nhalf == 2 && half == 0 && cb < IMG;
        const float *pl = inL ? sideL : side + 2 * IMG + 2, *pr = inR ? sideR : side + 2 * IMG + 2;
        auto hrow = [&](int r, bool direct, float &x0, float &x1, float &h0, float &h1) {
            const float *row = A + r * IMG;
            x0 = ok0 ? row[c0] : 0.0f;
            x1 = ok1 ? row[c1] : 0.0f;
            float xl = __shfl_up_sync(0xffffffffu, x1, 1);
            float xr = __shfl_down_sync(0xffffffffu, x0, 1);
            if (direct) {
                if (lane == 0) xl = inL ? row[cb - 1] : 0.0f;
                if (lane == 31) xr = inR ? row[cb] : 0.0f;
            } else {
                if (lane == 0) xl = pl[r];
                if (lane == 31) xr = pr[r];
            }
            h0 = xl + 2.0f * x0 + x1;
            h1 = x0 + 2.0f * x1 + xr;
        };
#else
        auto hrow = [&](int r, bool direct, float &x0, float &x1, float &h0, float &h1) {
            const float *row = A + r * IMG;
            x0 = ok0 ? row[c0] : 0.0f;
            x1 = ok1 ? row[c1] : 0.0f;
            float xl = __shfl_up_sync(0xffffffffu, x1, 1);
            float xr = __shfl_down_sync(0xffffffffu, x0, 1);
            if (nhalf == 1) {                    // the usual case: nothing lies left or right of the warp's span
                if (lane == 0) xl = 0.0f;
                if (lane == 31) xr = 0.0f;
            } else {
                if (lane == 0) xl = (half == 0) ? 0.0f : (direct ? row[cb - 1] : sideL[r]);
                if (lane == 31) xr = (half == 1) ? 0.0f : (cb < IMG ? (direct ? row[cb] : sideR[r]) : 0.0f);
            }
            h0 = xl + 2.0f * x0 + x1;
            h1 = x0 + 2.0f * x1 + xr;
        };
#endif
        float hp0 = 0.0f, hp1 = 0.0f, ht0 = 0.0f, ht1 = 0.0f, d0, d1;
        if (live) {
            if (a0 > 0) hrow(a0 - 1, true, d0, d1, hp0, hp1);
            if (a1 < IMG - 1) hrow(a1 + 1, true, d0, d1, ht0, ht1);
        }
        __syncthreads();
        if (live) {
            float xc0, xc1, hc0, hc1;
            hrow(a0, false, xc0, xc1, hc0, hc1);
#ifndef BPL_BROADEN_BRANCHY
            const float vmax = last ? 1.0f : 3.0e38f;                          // clamp_(max=1) in the last pass only (:171)
            auto put = [&](int r, float hn0, float hn1) {
                const float v0 = fminf(fmaf(c1w, hp0 + 2.0f * hc0 + hn0, c2w * xc0), vmax);
                const float v1 = fminf(fmaf(c1w, hp1 + 2.0f * hc1 + hn1, c2w * xc1), vmax);
                float *row = A + r * IMG;
                if (ok0) row[c0] = v0;
                if (ok1) row[c1] = v1;
            };
#pragma unroll BROADEN_UNROLL
            for (int r = a0; r < a1; ++r) {                                     // every row but the band's last: the next row is the band's own
                float xn0, xn1, hn0, hn1;
                hrow(r + 1, false, xn0, xn1, hn0, hn1);
                put(r, hn0, hn1);
                hp0 = hc0; hp1 = hc1; hc0 = hn0; hc1 = hn1; xc0 = xn0; xc1 = xn1;
            }
            put(a1, ht0, ht1);                                                  // below the last row: the halo taken before the barrier
#else
#pragma unroll 2
            for (int r = a0; r <= a1; ++r) {
                float xn0 = 0.0f, xn1 = 0.0f, hn0 = ht0, hn1 = ht1;
                if (r < a1) hrow(r + 1, false, xn0, xn1, hn0, hn1);
                float v0 = fmaf(c1w, hp0 + 2.0f * hc0 + hn0, c2w * xc0);
                float v1 = fmaf(c1w, hp1 + 2.0f * hc1 + hn1, c2w * xc1);
                if (last) { v0 = fminf(v0, 1.0f); v1 = fminf(v1, 1.0f); }      // clamp_(max=1)  (:171)
                float *row = A + r * IMG;
                if (ok0) row[c0] = v0;
                if (ok1) row[c1] = v1;
                hp0 = hc0; hp1 = hc1; hc0 = hn0; hc1 = hn1; xc0 = xn0; xc1 = xn1;
            }
#endif
        }
        __syncthreads();
    }
}

struct Fwd1sSmem {
    float A[FBUF];
    float side[2 * IMG + 2 + IMG];      // two boundary columns (broaden_box), then a column of zeros
    float tab[MAX_TABLE];
    int bbox[4];
    Misc ms;
};

// four CTAs of the large-batch shape must fit an SM (228 KB, 1 KB reserved per CTA)
static_assert((sizeof(Fwd1sSmem) + 1024) * S_CTAS_LARGE <= 233472, "Fwd1sSmem too large for the large-batch launch shape");

// AP: all-pairs mode (bpl_targets.all_pairs): the token is rendered once and scored against every target image.
// Launch shape: 320 threads x 3 CTAs per SM below 2 048 tokens (fewer, faster CTAs: a shorter tail of the launch), 256 x 4
// from there on (measured M particles/s, 320x3 / 256x4: 1 024 tokens 10.9 / 10.6, 2 048: 14.2 / 14.8, 4 096: 16.2 / 17.0,
// 32 768: 18.0 / 19.2, 131 072: 18.3 / 19.2).
template <bool AP, int S_THREADS, int S_CTAS>
__global__ void __launch_bounds__(S_THREADS, S_CTAS) bpl_forward1s_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fwd1sSmem &sm = *reinterpret_cast<Fwd1sSmem *>(smem_raw);
    float *A = sm.A;
    Misc *ms = &sm.ms;
    const int lane = threadIdx.x & 31;
    const bool aff = a.affine != nullptr;

    for (int i = threadIdx.x; i < a.neval * NCPT; i += blockDim.x) sm.tab[i] = a.basis[i];
    if (threadIdx.x == 0) { ms->cur_img = -1; ms->c_eps = ms->c_sigma = nanf(""); ms->n_ones = 0; ms->ones_acc = 0; }
    zero_floats(A, FBUF);              // afterwards the tail re-zeroes exactly the rows it touched
    for (int i = threadIdx.x; i < IMG; i += blockDim.x) sm.side[2 * IMG + 2 + i] = 0.0f;
    __syncthreads();
#ifdef BPL_PHASE_TIMING
    long long pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pt_t0 = clock64();
#endif

    float run_m = -INFINITY, run_s = 0.0f;                      // thread 0: log-sum-exp of this CTA's scores (a.lse)
    int item = blockIdx.x;                                      // position in the work queue (advanced by thread 0 only)
    int p = a.order ? a.order[item] : item;                     // the token it stands for
    while (p < a.n_tokens) {
        int p_next = 0;
#ifdef BPL_TOKEN_TIMING
        const long long tt0 = clock64();      // cycles per token -> a.pt_buf[p] (tools/token_cost_fit.py)
#endif
        if (threadIdx.x == 0) p_next = sched_next_token(a, item);      // consumed at the end of the iteration
        TokGeom t = token_geom<false>(a, p);
        const bool bad = t.S > MAXS || t.K > MAXK;
        if (bad) { t.S = 0; t.K = 0; }
        const float eps = a.eps[p], sigma = a.sigma[p];
        if (!AP) load_image_bits(a, p, ms);
        if (threadIdx.x == 0) { ms->off_page = 0; sm.bbox[0] = IMG; sm.bbox[1] = -1; sm.bbox[2] = IMG; sm.bbox[3] = -1; }
        BPL_PT(0);
        chain_offsets(a, t, sm.tab, ms);
        if (aff) affine_setup(a, t, p, sm.tab, ms);
        token_consts(eps, sigma, false, ms, a.rc.fhalf);
        if (threadIdx.x == 0) {
            const int img = a.img_idx ? a.img_idx[p] : 0;
            if (img != ms->cur_img) { ms->n_ones = ms->ones_acc; ms->cur_img = img; }      // a new image was loaded above
        }

        BPL_PT(1);
        // ---- add_stroke ----
        const int QN = (a.neval + PPL - 1) / PPL, items = t.S * QN;
        bool any_out = false;
        for (int base = 0; base < items; base += blockDim.x) {
            const int item = base + threadIdx.x;
            const bool active = item < items;
            const int s = active ? item / QN : 0x7fffffff;
            CtrlSrc src{};
            if (active) src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
            const LaneAcc acc = warp_points<false, 1, IMG, 0, PPL, LayS<1, IMG, 0>, S_ROLL_A>(active, src, sm.tab, s, a.neval, active ? item - s * QN : 0, lane,
                                                      a.rc, A, 0.0f, 0, sm.bbox);
            any_out |= acc.any_out;
            reduce_by_substroke(active, s, acc.sum, acc.cnt, ms->sub, ms->ink64, lane);
        }
        BPL_PT(6);
        any_out = __any_sync(0xffffffffu, any_out);
        if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
        __syncthreads();
        BPL_PT(7);
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
                warp_points<true, 1, IMG, 0, PPL, LayS<1, IMG, 0>, S_ROLL_A>(active, src, sm.tab, s, a.neval, item < items ? item - s * QN : 0, lane, a.rc, A, v.z,
                                     __float_as_int(v.w));
            }
            __syncthreads();
        }
        BBox bb{sm.bbox[0], sm.bbox[1], sm.bbox[2], sm.bbox[3]};
        BPL_PT(2);

        // ---- broaden + clamp, blur: only where the image can be non-zero ----
        if (!bb.empty()) {
            broaden_box<S_THREADS>(A, sm.side, a.rc, bb);
            bb.grow_rows(a.rc.ink_ncon); bb.grow_cols(a.rc.ink_ncon);
            if (a.rc.ink_ncon == 0) {
                for (int i = threadIdx.x; i < NPIX; i += blockDim.x) A[i] = fminf(A[i], 1.0f);
                __syncthreads();
            }
            BPL_PT(3);
            if (sigma > 0.0f) {
                G6 g;
#pragma unroll
                for (int d = 0; d <= PAD; ++d) g.v[d] = ms->g[d];
#pragma unroll 1
                for (int rep = 0; rep < 2; ++rep) {
                    // vertical: lines = the box's columns, bands = those holding output rows [r0-5, r1+5]
                    bb.grow_rows(PAD);
                    gauss_box<true, S_THREADS>(A, g, bb.c0, bb.c1, bb.r0, bb.r1);
                    __syncthreads();
                    bb.grow_cols(PAD);
                    gauss_box<false, S_THREADS>(A, g, bb.r0, bb.r1, bb.c0, bb.c1);
                    __syncthreads();
                }
            }
        }

        BPL_PT(4);
        // ---- clamp, mix, likelihood ----
        // ll = [score of an all-background render] + sum over the box of (lp - background lp): the first term is
        // lp_bg[1] * (ones of the target) + lp_bg[0] * (zeros), the ones counted once per image; background pixels
        // inside the box contribute exactly 0 to the second.
        const float bgv = eps > 0.0f ? eps : 0.0f;
        const float om = 1.0f - eps;
        const float l0 = ms->lp_bg[0], l1 = ms->lp_bg[1];
        float acc_ll = 0.0f;
        if (!AP) {
            if (!bb.empty()) {
                int wbox = bb.c1 - bb.c0 + 1, c0t = bb.c0;
                if (wbox == 1) { wbox = 2; c0t = max(c0t - 1, 0); }            // (a one-column box takes a background column along:
                                                                                //  the multiply-high division needs wbox >= 2)
#ifndef BPL_TAIL_ROWS
                // the box's pixels flattened over the CTA: no lanes idle at the ends of rows (a 65-pixel row costs a warp
                // three rounds, the third for one pixel), at the price of a reciprocal multiplication per pixel: +1.0 % at 4 096
                // tokens, +1.4 % at 131 072 (20.24 -> 20.52 M particles/s); -DBPL_TAIL_ROWS keeps one warp per row
                const int nbox = wbox * (bb.r1 - bb.r0 + 1);
                // row of box pixel k = k / wbox = umulhi(k, ceil(2^32 / wbox)), exact while k * wbox < 2^32; one integer
                // multiply-high instead of two conversions through the special-function pipe
                const unsigned mw = 0xffffffffu / (unsigned)wbox + 1u;         // (wbox >= 2: see below)
                const int skip = IMG - wbox, base0 = bb.r0 * IMG + c0t;
                for (int k = threadIdx.x; k < nbox; k += blockDim.x) {
                    const int rr = (int)__umulhi((unsigned)k, mw);
                    {
                        const int i = rr * skip + (k + base0);                  // (r0 + rr) * IMG + c0 + (k - rr * wbox)
#else
                const int nw_ = blockDim.x >> 5, warp_ = threadIdx.x >> 5;
                for (int r = bb.r0 + warp_; r <= bb.r1; r += nw_) {            // one warp per row of the box
                    for (int cc = lane; cc < wbox; cc += 32) {
                        const int i = r * IMG + bb.c0 + cc;
#endif
                        float v = fminf(fmaxf(A[i], 0.0f), 1.0f);                                          // clamp_(0,1) (:180)
                        A[i] = 0.0f;                                            // the buffer is all zero again afterwards
                        if (eps > 0.0f) v = __fadd_rn(__fmul_rn(om, v), __fmul_rn(eps, __fsub_rn(1.0f, v)));   // (:227)
                        if (v != bgv) {
                            const bool x = (ms->bits[i >> 5] >> (i & 31)) & 1u;
                            acc_ll += pixel_lp_fwd(v, x) - (x ? l1 : l0);
                        }
                    }
                }
            }
            double v[1] = {(double)acc_ll};
            if (threadIdx.x == 0) v[0] += (double)ms->n_ones * (double)l1 + (double)(NPIX - ms->n_ones) * (double)l0;
            block_sum<1>(v, ms->red);
            if (threadIdx.x == 0) {
                if (a.ll) a.ll[p] = bad ? nanf("") : (float)v[0];
                if (a.off_page) a.off_page[p] = bad ? (uint8_t)255 : (uint8_t)(ms->off_page != 0);
                if (a.lse && !bad) lse_push(run_m, run_s, (float)v[0]);
                ms->ones_acc = 0;
                ms->next_item = p_next;
            }
        } else {
            // Every image m: ll_m = [all pixels scored as x = 0] + sum over the set bits of m of delta, delta = lp(x=1) - lp(x=0).
            // With delta_bg the background's delta: ll_m = base + ones(m) * delta_bg + sum over set bits of (delta - delta_bg),
            // and (delta - delta_bg) is zero outside the box and at background pixels, i.e. wherever the buffer holds zero.
            const float dbg = l1 - l0;
            if (!bb.empty()) {
                int wbox = bb.c1 - bb.c0 + 1, c0t = bb.c0;
                if (wbox == 1) { wbox = 2; c0t = max(c0t - 1, 0); }            // (a one-column box takes a background column along:
                                                                                //  the multiply-high division needs wbox >= 2)
#ifndef BPL_TAIL_ROWS
                const int nbox = wbox * (bb.r1 - bb.r0 + 1);
                const unsigned mw = 0xffffffffu / (unsigned)wbox + 1u;
                const int skip = IMG - wbox, base0 = bb.r0 * IMG + c0t;
                for (int k = threadIdx.x; k < nbox; k += blockDim.x) {
                    const int rr = (int)__umulhi((unsigned)k, mw);
                    {
                        const int i = rr * skip + (k + base0);
#else
                const int nw_ = blockDim.x >> 5, warp_ = threadIdx.x >> 5;
                for (int r = bb.r0 + warp_; r <= bb.r1; r += nw_) {
                    for (int cc = lane; cc < wbox; cc += 32) {
                        const int i = r * IMG + bb.c0 + cc;
#endif
                        float v = fminf(fmaxf(A[i], 0.0f), 1.0f);
                        if (eps > 0.0f) v = __fadd_rn(__fmul_rn(om, v), __fmul_rn(eps, __fsub_rn(1.0f, v)));
                        float dl = 0.0f;
                        if (v != bgv) {
                            const float s0 = pixel_lp_fwd(v, false), s1 = pixel_lp_fwd(v, true);
                            acc_ll += s0 - l0;
                            dl = (s1 - s0) - dbg;
                        }
                        A[i] = dl;
                    }
                }
            }
            double v[1] = {(double)acc_ll};
            block_sum<1>(v, ms->red);      // its barriers also publish the deltas
            const double base = v[0] + (double)NPIX * (double)l0;
            const int nw_ = blockDim.x >> 5, warp_ = threadIdx.x >> 5;
            for (int m = warp_; m < a.n_images; m += nw_) {      // one warp per image walks the image's set bits
                const uint32_t *bw = a.image_bits + (size_t)m * IMG_WORDS;
                float s = 0.0f;
                int ones = 0;
                for (int wi = lane; wi < IMG_WORDS; wi += 32) {
                    unsigned bits = bw[wi];
                    if (wi == IMG_WORDS - 1) bits &= (1u << (NPIX - 32 * (IMG_WORDS - 1))) - 1u;      // 11025 = 344 * 32 + 17
                    ones += __popc(bits);
                    for (; bits; bits &= bits - 1u) s += A[wi * 32 + __ffs(bits) - 1];
                }
                s = warp_sum(s);
                ones = __reduce_add_sync(0xffffffffu, ones);
                if (lane == 0 && a.ll)
                    a.ll[(size_t)p * a.n_images + m] = bad ? nanf("") : (float)(base + (double)ones * (double)dbg + (double)s);
            }
            if (threadIdx.x == 0) {
                if (a.off_page) a.off_page[p] = bad ? (uint8_t)255 : (uint8_t)(ms->off_page != 0);
                ms->next_item = p_next;
            }
            __syncthreads();
            if (!bb.empty()) {                                      // the deltas must not leak into the next token
                const int wbox = bb.c1 - bb.c0 + 1, nbox = wbox * (bb.r1 - bb.r0 + 1);
                for (int i = threadIdx.x; i < nbox; i += blockDim.x) {
                    const int r = i / wbox;
                    A[(bb.r0 + r) * IMG + bb.c0 + (i - r * wbox)] = 0.0f;
                }
            }
        }
        __syncthreads();
#ifdef BPL_TOKEN_TIMING
        if (threadIdx.x == 0 && a.pt_buf) a.pt_buf[p] = (float)(clock64() - tt0);
#endif
        p = ms->next_item;      // rewritten only at the end of the next iteration, many barriers from here
        BPL_PT(5);
    }
    if (!AP && a.lse) lse_finish(a, run_m, run_s, &ms->next_item);
    else if (threadIdx.x == 0) sched_finish(a);
#ifdef BPL_PHASE_TIMING
    if (threadIdx.x == 0 && a.pt_buf)
        for (int i = 0; i < 8; ++i) a.pt_buf[blockIdx.x * 16 + i] = (float)pt_acc[i];
#endif
}

}  // namespace bpl
