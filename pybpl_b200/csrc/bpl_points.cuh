// bpl_points.cuh -- add_stroke (rendering.py:58-127) for control-point tokens, lane-serial layout.
//
// A thread owns PPL = 7 CONSECUTIVE points of one sub-stroke and walks them serially, so a point's
// previous surviving neighbour (out-of-bounds points are dropped before distances are taken,
// rendering.py:85-98) is simply the thread's own previous iteration; only the neighbour before the
// run is re-evaluated (one extra spline evaluation, a short backward scan if that one is off the page).
// The 32 lanes of one shared-memory atomic instruction therefore hold points 7 apart along the stroke:
// with the earlier one-point-per-lane layout six or more lanes hit the same pixel and the CAS loop that
// implements a shared-memory float atomicAdd (there is no native one) replayed that many times -- ncu
// showed it as the top stall and 23 % of all shared-memory wavefronts.
//
// Ink is splatted at once assuming the common branch while (sum of ink, survivor count) accumulate per
// sub-stroke; the rare min-ink branches (rendering.py:104-107) are patched by a second pass with
// FIX = true that adds (final - common) ink.
#pragma once

// Phase A of the point stages (a lane evaluates its PPL points once to see which are on the page): unrolled by default;
// -DBPL_ROLL_PHASE_A keeps one copy of the point evaluation per stage (instruction-cache footprint A/B).
#ifdef BPL_ROLL_PHASE_A
#define BPL_UNROLL_A _Pragma("unroll 1")
#else
#define BPL_UNROLL_A _Pragma("unroll")
#endif
// The gather stage (warp_points_bwd) and the team kernel's bounding-box pass (warp_bbox) take phase A ROLLED by default:
// together with the rolled splat of the team kernel (BT_ROLL_SPLAT) it is +4.5 % at 8 192 parses (5.09 -> 5.33 M/s) and
// within the noise at 1 024 -- the team kernel is bound by instruction fetch (253 KB of SASS), see DESIGN.md.
// -DBPL_UNROLL_PHASE_A_BWD restores the unrolled form.
#if !defined(BPL_UNROLL_PHASE_A_BWD)
#define BPL_UNROLL_A_BWD _Pragma("unroll 1")
#define BPL_UNROLL_A_BBOX _Pragma("unroll 1")
#else
#define BPL_UNROLL_A_BWD _Pragma("unroll")
#define BPL_UNROLL_A_BBOX _Pragma("unroll")
#endif

namespace bpl {

#ifndef BPL_PPL
#define BPL_PPL 4
#endif
constexpr int PPL_DEFAULT = BPL_PPL;      // points per lane (forward kernels)
constexpr int PPL = PPL_DEFAULT;

struct PEval { float r, c; bool inb; };

// Where pixel (r, c) of the canvas lives in an image buffer: a compile-time layout (the full-canvas kernels) or a
// run-time one (the team kernels' box-sized buffers: org = -(r0 * stride + c0) of the box).
template <int PIXSTRIDE, int ROWSTRIDE, int ORG>
struct LayS { __device__ __forceinline__ int operator()(int r, int c) const { return ORG + PIXSTRIDE * (r * ROWSTRIDE + c); } };
struct LayD {
    int stride, org;
    __device__ __forceinline__ int operator()(int r, int c) const { return org + r * stride + c; }
};

// one trajectory point in image space: spline + chain offset (+ affine), flip (rendering.py:52-53), check_bounds (:27-31)
__device__ __forceinline__ PEval eval_pt(const CtrlSrc &src, int j, bool valid) {
    PEval e;
    float x = 0.0f, y = 0.0f;
    if (valid) src.point(j, x, y);
    e.r = -y; e.c = x;
    e.inb = valid && (e.r >= 0.0f) && (e.r <= (float)(IMG - 1)) && (e.c >= 0.0f) && (e.c <= (float)(IMG - 1));
    return e;
}

// ink of a point from the clamped distance to its neighbour (rendering.py:96-99)
__device__ __forceinline__ float ink_from(float r, float c, float qr, float qc, const RenderConsts &rc) {
    const float dr = __fsub_rn(r, qr), dc = __fsub_rn(c, qc);
    const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(dc, dc)));
    const float num = __fmul_rn(rc.ink_pp, fminf(dist, rc.ink_max_dist));
    return rc.inv_max_dist != 0.0f ? __fmul_rn(num, rc.inv_max_dist) : __fdiv_rn(num, rc.ink_max_dist);
}

template <int PIXSTRIDE>
__device__ __forceinline__ void splat_px(float *img, float r, float c, float ink) {     // rendering.py:111-125
    const float rf = floorf(r), cf = floorf(c), rce = ceilf(r), cce = ceilf(c);
    const float fr = __fsub_rn(r, rf), fc = __fsub_rn(c, cf);
    const float gr = __fsub_rn(1.0f, fr), gc = __fsub_rn(1.0f, fc);
    const int r0 = (int)rf, c0 = (int)cf, r1 = (int)rce, c1 = (int)cce;
    atomicAdd(img + PIXSTRIDE * (r0 * IMG + c0), __fmul_rn(__fmul_rn(ink, gr), gc));
    atomicAdd(img + PIXSTRIDE * (r1 * IMG + c0), __fmul_rn(__fmul_rn(ink, fr), gc));
    atomicAdd(img + PIXSTRIDE * (r0 * IMG + c1), __fmul_rn(__fmul_rn(ink, gr), fc));
    atomicAdd(img + PIXSTRIDE * (r1 * IMG + c1), __fmul_rn(__fmul_rn(ink, fr), fc));
}

struct LaneAcc {
    float sum;      // sum of common-branch ink of this thread's survivors
    int cnt;        // survivors
    bool any_out;
#ifdef BPL_PHASE_TIMING
    long long tA, tB, tC;
#endif
};

__device__ __forceinline__ CtrlSrc shfl_src(const CtrlSrc &s, int from, const float *tab) {
    CtrlSrc o;
    o.tab = tab;        // not s.tab: inactive lanes hold no source at all
#pragma unroll
    for (int k = 0; k < NCPT; ++k) {
        o.yx[k] = __shfl_sync(0xffffffffu, s.yx[k], from);
        o.yy[k] = __shfl_sync(0xffffffffu, s.yy[k], from);
    }
    o.offx = __shfl_sync(0xffffffffu, s.offx, from); o.offy = __shfl_sync(0xffffffffu, s.offy, from);
    o.aff = __shfl_sync(0xffffffffu, (int)s.aff, from) != 0;
    o.b0 = __shfl_sync(0xffffffffu, s.b0, from); o.b1 = __shfl_sync(0xffffffffu, s.b1, from);
    o.b2 = __shfl_sync(0xffffffffu, s.b2, from); o.b3 = __shfl_sync(0xffffffffu, s.b3, from);
    return o;
}

// Warp-collective (all 32 lanes call; `active` lanes own run q of sub-stroke `sid`; lanes of a warp hold
// consecutive runs, so sid is non-decreasing across lanes).  Three steps:
//  A. every lane evaluates its PPL points and notes its last survivor;
//  B. a segmented shuffle scan hands every lane the last survivor of the earlier lanes of its sub-stroke;
//     the part of the warp's first sub-stroke that lies before the warp is searched cooperatively,
//     32 points per step (a sub-stroke that is entirely off the page costs 7 such steps, not 200);
//  C. every lane walks its points: ink from the distance to the previous survivor, splat.
// FIX: fix_sum / fix_cnt are the sub-stroke totals; ink becomes fill or scale*ink and only the
// difference to the common-branch ink is splatted.
// ROLL_A: phase A as a rolled loop (one copy of the point evaluation instead of PPL).  The box-confined scoring kernel
// takes it: with its loop bodies trimmed instruction fetch weighs more than the loop overhead,
// +1.4 % on 131 072 tokens, +0.9 % on 4 096; the backward kernels keep the unrolled form (-2 % at 1 024 parses rolled).
template <bool FIX, int PIXSTRIDE, int ROWSTRIDE = IMG, int ORG = 0, int PPL = PPL_DEFAULT,
          class Lay = LayS<PIXSTRIDE, ROWSTRIDE, ORG>, bool ROLL_A = false>
__device__ __forceinline__ LaneAcc warp_points(bool active, const CtrlSrc &src, const float *tab, int sid, int n, int q,
                                               int lane, const RenderConsts &rc, float *img, float fix_sum, int fix_cnt,
                                               int *bbox = nullptr, Lay lay = Lay()) {
    LaneAcc acc{};
    if (!__any_sync(0xffffffffu, active)) return acc;      // e.g. the fix pass: most warps have nothing to patch
    const int j0 = q * PPL;
#ifdef BPL_PHASE_TIMING
    long long tt0 = clock64();
#endif
    // ---- A ----
    unsigned inbm = 0u;
    bool hv = false;
    float lr = 0.0f, lc = 0.0f;
#ifdef BPL_ROLL_PHASE_A
#pragma unroll 1
#else
#pragma unroll (ROLL_A ? 1 : PPL)
#endif
    for (int k = 0; k < PPL; ++k) {
        const int j = j0 + k;
        const PEval e = eval_pt(src, j, active && j < n);
        acc.any_out |= active && (j < n) && !e.inb;
        if (e.inb) { inbm |= 1u << k; hv = true; lr = e.r; lc = e.c; }
    }
#ifdef BPL_PHASE_TIMING
    { const long long t1 = clock64(); acc.tA = t1 - tt0; tt0 = t1; }
#endif
    // ---- B ----
    // Fast path: the predecessor of a run is simply point j0-1 when that point is on the page (nearly always);
    // only a warp in which some run's predecessor is off the page takes the scan below.
    bool has_prev = false;
    float pr = 0.0f, pc = 0.0f;
    bool need_scan = false;
    if (active && hv && j0 > 0) {
        const PEval e = eval_pt(src, j0 - 1, true);
        if (e.inb) { has_prev = true; pr = e.r; pc = e.c; } else need_scan = true;
    }
    if (__any_sync(0xffffffffu, need_scan)) {
    bool iv = hv;
    float ir = lr, ic = lc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool ov = __shfl_up_sync(0xffffffffu, (int)iv, d) != 0;
        const float orr = __shfl_up_sync(0xffffffffu, ir, d), oc = __shfl_up_sync(0xffffffffu, ic, d);
        const int os = __shfl_up_sync(0xffffffffu, sid, d);
        if (lane >= d && os == sid && !iv && ov) { iv = true; ir = orr; ic = oc; }
    }
    has_prev = __shfl_up_sync(0xffffffffu, (int)iv, 1) != 0;
    pr = __shfl_up_sync(0xffffffffu, ir, 1); pc = __shfl_up_sync(0xffffffffu, ic, 1);
    {
        const int ps = __shfl_up_sync(0xffffffffu, sid, 1);
        if (lane == 0 || ps != sid) has_prev = false;
    }
    // carry from before the warp, for lanes of the warp's first sub-stroke that still have no predecessor
    const int sid0 = __shfl_sync(0xffffffffu, sid, 0);
    const int q0 = __shfl_sync(0xffffffffu, q, 0);
    const bool act0 = __shfl_sync(0xffffffffu, (int)active, 0) != 0;
    const bool want = active && hv && !has_prev && sid == sid0;
    if (act0 && q0 > 0 && __any_sync(0xffffffffu, want)) {
        const CtrlSrc s0 = shfl_src(src, 0, tab);
        bool found = false;
        float cr = 0.0f, cc = 0.0f;
        for (int hi = q0 * PPL; hi > 0 && !found; hi -= 32) {
            const int j = hi - 32 + lane;
            const PEval e = eval_pt(s0, j, j >= 0);
            const unsigned m = __ballot_sync(0xffffffffu, e.inb);
            if (m) {
                const int l = 31 - __clz(m);
                cr = __shfl_sync(0xffffffffu, e.r, l); cc = __shfl_sync(0xffffffffu, e.c, l);
                found = true;
            }
        }
        if (found && sid == sid0 && !has_prev) { has_prev = true; pr = cr; pc = cc; }
    }
    }
#ifdef BPL_PHASE_TIMING
    { const long long t1 = clock64(); acc.tB = t1 - tt0; tt0 = t1; }
#endif
    // ---- C ----
    float fill = 0.0f, scale = 1.0f;
    if (FIX) {
        if (fix_sum < 2.22e-6f) fill = (float)((double)rc.ink_pp / (double)fix_cnt);      // (:104-105)
        else scale = __fdiv_rn(rc.ink_pp, fix_sum);                                     // (:106-107)
    }
    // Consecutive points of a run mostly fall into the same pixel cell: their four corner contributions are
    // summed in registers and flushed with one set of atomics when the cell changes (2-4x fewer atomics).
    int cell = -1;
    int brmin = IMG, brmax = -1, bcmin = IMG, bcmax = -1;      // pixels this lane touches (for the sparse kernel)
    float a00 = 0.0f, a10 = 0.0f, a01 = 0.0f, a11 = 0.0f;
    auto flush = [&]() {
        if (cell >= 0) {
            const int r0 = cell & 127, c0 = (cell >> 7) & 127, r1 = r0 + ((cell >> 14) & 1), c1 = c0 + ((cell >> 15) & 1);
            atomicAdd(img + lay(r0, c0), a00);
            atomicAdd(img + lay(r1, c0), a10);
            atomicAdd(img + lay(r0, c1), a01);
            atomicAdd(img + lay(r1, c1), a11);
        }
    };
    auto emit = [&](float r, float c, float ink0) {
        float ink = ink0;
        if (!FIX) acc.sum += ink0;
        else ink = ((fix_sum < 2.22e-6f) ? fill : __fmul_rn(scale, ink0)) - ink0;
        const float rf = floorf(r), cf = floorf(c), rce = ceilf(r), cce = ceilf(c);      // rendering.py:111-125
        const float fr_ = __fsub_rn(r, rf), fc_ = __fsub_rn(c, cf);
        const float gr = __fsub_rn(1.0f, fr_), gc = __fsub_rn(1.0f, fc_);
        const int key = (int)rf | ((int)cf << 7) | ((int)(rce - rf) << 14) | ((int)(cce - cf) << 15);
        if (key != cell) {
            flush();
            cell = key; a00 = a10 = a01 = a11 = 0.0f;
            brmin = min(brmin, (int)rf); brmax = max(brmax, (int)rce);
            bcmin = min(bcmin, (int)cf); bcmax = max(bcmax, (int)cce);
        }
        a00 += __fmul_rn(__fmul_rn(ink, gr), gc);
        a10 += __fmul_rn(__fmul_rn(ink, fr_), gc);
        a01 += __fmul_rn(__fmul_rn(ink, gr), fc_);
        a11 += __fmul_rn(__fmul_rn(ink, fr_), fc_);
    };
    // The walk is a rolled loop that re-evaluates each surviving point (14 instructions) instead of keeping
    // seven positions in registers: one copy of the splat code, fewer registers, no spills.
    bool pend = false;          // the sub-stroke's first survivor waits for its successor (dist[:1], :98)
    bool have_first = false;    // ... and is splatted last, with that successor's ink
    float fr = 0.0f, fc = 0.0f, first_ink = 0.0f;
#ifdef BPL_WALK_UNROLL
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
        if (!((inbm >> k) & 1u)) continue;
#else
#pragma unroll 1
    for (unsigned rest = inbm; rest; rest &= rest - 1u) {
        const int k = __ffs(rest) - 1;
#endif
        const PEval e = eval_pt(src, j0 + k, true);
        ++acc.cnt;
        if (has_prev) {
            const float ink0 = ink_from(e.r, e.c, pr, pc, rc);
            if (pend) { first_ink = ink0; have_first = true; pend = false; }
            emit(e.r, e.c, ink0);
        } else {
            pend = true; fr = e.r; fc = e.c;
        }
        has_prev = true; pr = e.r; pc = e.c;
    }
    if (pend) {                 // successor beyond the run, or none at all: the only survivor gets ink_pp (:93-94)
        first_ink = rc.ink_pp;
        for (int j = j0 + PPL; j < n; ++j) {
            const PEval e = eval_pt(src, j, true);
            if (e.inb) { first_ink = ink_from(e.r, e.c, fr, fc, rc); break; }
        }
        have_first = true;
    }
    if (have_first) emit(fr, fc, first_ink);
    flush();
    if (bbox) {         // warp-uniform pointer: one set of shared-memory atomics per warp
        brmin = __reduce_min_sync(0xffffffffu, brmin); brmax = __reduce_max_sync(0xffffffffu, brmax);
        bcmin = __reduce_min_sync(0xffffffffu, bcmin); bcmax = __reduce_max_sync(0xffffffffu, bcmax);
        if (lane == 0 && brmax >= 0) {
            atomicMin(bbox + 0, brmin); atomicMax(bbox + 1, brmax);
            atomicMin(bbox + 2, bcmin); atomicMax(bbox + 3, bcmax);
        }
    }
#ifdef BPL_PHASE_TIMING
    { const long long t1 = clock64(); acc.tC = t1 - tt0; tt0 = t1; }
#endif
    return acc;
}

// ------------------------------------------------------------------------------------------------
// Backward of add_stroke in the same lane-serial layout (control-point tokens).
//   G        image gradient d L / d I0 in the padded layout of the backward kernel (ORG/ROWSTRIDE)
//   branch   decided from the sub-stroke totals (sum, cnt) the forward pass accumulated
//   DOT      pre-pass for the rescale branch: only dot = sum_k gink[k] * ink0[k] is accumulated
// emit(j, gx, gy) receives motor-space (post-warp) point gradients; a point may be emitted several
// times (own splat, own segment, the next segment).  The sub-stroke's first survivor has the ink of the
// second (dist[:1], rendering.py:98): its lane looks up its successor and differentiates that copy itself.
// ------------------------------------------------------------------------------------------------
struct Corner2 { float gink, gr, gc; };

template <class Lay>
__device__ __forceinline__ Corner2 gather_px(const float *G, float r, float c, const Lay &lay) {
    const float rf = floorf(r), cf = floorf(c), rce = ceilf(r), cce = ceilf(c);
    const float fr = r - rf, fc = c - cf, gr = 1.0f - fr, gc = 1.0f - fc;
    const int r0 = (int)rf, c0 = (int)cf, r1 = (int)rce, c1 = (int)cce;
    const float G00 = G[lay(r0, c0)], G10 = G[lay(r1, c0)];
    const float G01 = G[lay(r0, c1)], G11 = G[lay(r1, c1)];
    Corner2 k;
    k.gink = gr * gc * G00 + fr * gc * G10 + gr * fc * G01 + fr * fc * G11;
    k.gr = gc * (G10 - G00) + fc * (G11 - G01);
    k.gc = gr * (G01 - G00) + fr * (G11 - G10);
    return k;
}

template <bool DOT, int ROWSTRIDE, int ORG, int PPL, class Emit, class Lay = LayS<1, ROWSTRIDE, ORG>>
__device__ __forceinline__ float warp_points_bwd(bool active, const CtrlSrc &src, const float *tab, int sid, int n, int q,
                                                 int lane, const RenderConsts &rc, const float *G, float sum, int cnt,
                                                 float dot_in, Emit emit, Lay lay = Lay()) {
    if (!__any_sync(0xffffffffu, active)) return 0.0f;
    const int j0 = q * PPL;
    // ---- A: own points ----
    unsigned inbm = 0u;
    bool hv = false;
    float lr = 0.0f, lc = 0.0f;
    int lj = 0;
    BPL_UNROLL_A_BWD
    for (int k = 0; k < PPL; ++k) {
        const int j = j0 + k;
        const PEval e = eval_pt(src, j, active && j < n);
        if (e.inb) { inbm |= 1u << k; hv = true; lr = e.r; lc = e.c; lj = j; }
    }
    // ---- B: last survivor before the run (position and point index) ----
    // Fast path as in warp_points: point j0-1 when it is on the page; the scan only when some run's is not.
    bool has_prev = false;
    float pr = 0.0f, pc = 0.0f;
    int pj = 0;
    bool need_scan = false;
    if (active && hv && j0 > 0) {
        const PEval e = eval_pt(src, j0 - 1, true);
        if (e.inb) { has_prev = true; pr = e.r; pc = e.c; pj = j0 - 1; } else need_scan = true;
    }
    if (__any_sync(0xffffffffu, need_scan)) {
    bool iv = hv;
    float ir = lr, ic = lc;
    int ij = lj;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool ov = __shfl_up_sync(0xffffffffu, (int)iv, d) != 0;
        const float orr = __shfl_up_sync(0xffffffffu, ir, d), oc = __shfl_up_sync(0xffffffffu, ic, d);
        const int oj = __shfl_up_sync(0xffffffffu, ij, d), os = __shfl_up_sync(0xffffffffu, sid, d);
        if (lane >= d && os == sid && !iv && ov) { iv = true; ir = orr; ic = oc; ij = oj; }
    }
    has_prev = __shfl_up_sync(0xffffffffu, (int)iv, 1) != 0;
    pr = __shfl_up_sync(0xffffffffu, ir, 1); pc = __shfl_up_sync(0xffffffffu, ic, 1);
    pj = __shfl_up_sync(0xffffffffu, ij, 1);
    {
        const int ps = __shfl_up_sync(0xffffffffu, sid, 1);
        if (lane == 0 || ps != sid) has_prev = false;
    }
    const int sid0 = __shfl_sync(0xffffffffu, sid, 0);
    const int q0 = __shfl_sync(0xffffffffu, q, 0);
    const bool act0 = __shfl_sync(0xffffffffu, (int)active, 0) != 0;
    const bool want = active && hv && !has_prev && sid == sid0;
    if (act0 && q0 > 0 && __any_sync(0xffffffffu, want)) {
        const CtrlSrc s0 = shfl_src(src, 0, tab);
        bool found = false;
        float cr = 0.0f, cc = 0.0f;
        int cj = 0;
        for (int hi = q0 * PPL; hi > 0 && !found; hi -= 32) {
            const int j = hi - 32 + lane;
            const PEval e = eval_pt(s0, j, j >= 0);
            const unsigned m = __ballot_sync(0xffffffffu, e.inb);
            if (m) {
                const int l = 31 - __clz(m);
                cr = __shfl_sync(0xffffffffu, e.r, l); cc = __shfl_sync(0xffffffffu, e.c, l);
                cj = hi - 32 + l;
                found = true;
            }
        }
        if (found && sid == sid0 && !has_prev) { has_prev = true; pr = cr; pc = cc; pj = cj; }
    }
    }
    // ---- C: walk ----
    const bool fill = sum < 2.22e-6f;                               // (:104-105): constant ink, no distance gradient
    const bool rescale = !fill && sum < rc.ink_pp;                  // (:106-107)
    const float sc = rescale ? __fdiv_rn(rc.ink_pp, sum) : 1.0f;
    const float gsum = rescale ? -dot_in * rc.ink_pp / (sum * sum) : 0.0f;
    const float fillv = fill ? (float)((double)rc.ink_pp / (double)(cnt > 0 ? cnt : 1)) : 0.0f;
    float dot = 0.0f;
    bool pend = false;
    float fr = 0.0f, fc = 0.0f;
    int fj = 0;
    // One segment (earlier point e, later point l) defines ink0 of the later point -- and of the sub-stroke's first
    // survivor when e is that survivor (dist[:1], rendering.py:98) -- with d L / d ink0 = gi in total.  Returns the
    // gradient (ur, uc) the segment puts on the later point (the earlier one gets its negative); (0, 0) where the
    // clamp or the norm's zero gradient cuts it.
    // Every point receives ONE emit per walk step with all its contributions summed (own splat + own segment; the
    // previous survivor: its end of the segment + its splat when it is the first survivor, whose ink this segment also
    // defines): four inlined copies of the caller's emit instead of ten, and half as many emits executed.
    auto segment = [&](float er, float ec, float lr_, float lc_, float gi, float &ur, float &uc) {
        const float dr = __fsub_rn(lr_, er), dc = __fsub_rn(lc_, ec);
        const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(dc, dc)));
        ur = 0.0f; uc = 0.0f;
        if (!fill && dist <= rc.ink_max_dist && dist != 0.0f) {         // clamp passes at the bound; norm' = 0 at 0
            const float gd = rc.ink_f * gi;
            ur = gd * (dr / dist); uc = gd * (dc / dist);
        }
    };
#pragma unroll 1
    for (unsigned rest = active ? inbm : 0u; rest; rest &= rest - 1u) {
        const int j = j0 + __ffs(rest) - 1;
        const PEval e = eval_pt(src, j, true);
        const Corner2 K = gather_px(G, e.r, e.c, lay);
        if (has_prev) {
            const float ink0 = ink_from(e.r, e.c, pr, pc, rc);
            if (DOT) {
                dot = fmaf(K.gink, ink0, dot);
                if (pend) dot = fmaf(gather_px(G, fr, fc, lay).gink, ink0, dot);
            } else {
                const float ink = fill ? fillv : __fmul_rn(sc, ink0);
                float gi = fmaf(K.gink, sc, gsum);
                float gpc = 0.0f, gpr = 0.0f;                  // splat gradient of the previous survivor (col, row)
                if (pend) {             // the previous survivor is the first one (pj == fj): it carries a copy of this ink
                    const Corner2 Kf = gather_px(G, fr, fc, lay);
                    gpc = ink * Kf.gc; gpr = ink * Kf.gr;
                    gi += fmaf(Kf.gink, sc, gsum);
                }
                float ur, uc;
                segment(pr, pc, e.r, e.c, gi, ur, uc);
                emit(j, fmaf(ink, K.gc, uc), -fmaf(ink, K.gr, ur));        // un-flip: row = -y, col = x
                gpc -= uc; gpr -= ur;
                if (gpc != 0.0f || gpr != 0.0f) emit(pj, gpc, -gpr);
            }
            pend = false;
        } else {
            pend = true; fr = e.r; fc = e.c; fj = j;
        }
        has_prev = true; pr = e.r; pc = e.c; pj = j;
    }
    if (pend) {                 // successor beyond the run, or the only survivor (ink_pp, :93-94)
        bool found = false;
        float nr = 0.0f, nc = 0.0f;
        int nj = 0;
        for (int j = j0 + PPL; j < n; ++j) {
            const PEval e = eval_pt(src, j, true);
            if (e.inb) { found = true; nr = e.r; nc = e.c; nj = j; break; }
        }
        const Corner2 Kf = gather_px(G, fr, fc, lay);
        if (DOT) {
            if (found) dot = fmaf(Kf.gink, ink_from(nr, nc, fr, fc, rc), dot);
        } else {
            float ink = rc.ink_pp, ur = 0.0f, uc = 0.0f;
            if (found) {
                const float ink0 = ink_from(nr, nc, fr, fc, rc);
                ink = fill ? fillv : __fmul_rn(sc, ink0);
                segment(fr, fc, nr, nc, fmaf(Kf.gink, sc, gsum), ur, uc);
                if (ur != 0.0f || uc != 0.0f) emit(nj, uc, -ur);
            }
            emit(fj, fmaf(ink, Kf.gc, -uc), -fmaf(ink, Kf.gr, -ur));
        }
    }
    return dot;
}

// Sixteen sums over the warp for the price of sixteen shuffles (sixteen butterfly reductions take eighty): at every
// stage a lane hands half of the values it still holds to its partner and keeps the other half, so the number of
// values per lane halves as the partner distance does.  Afterwards lanes 2i and 2i+1 both hold the total of value i.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float send = up ? v[k] : v[k + half], keep = up ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Add each lane's (sum, cnt) into ink64[sid] / rec[sid].w, combining the lanes of a warp that share a sub-stroke first (a
// warp spans at most a few sub-strokes).  All 32 lanes must call this.  The sum is kept as 16.40 fixed point in two
// 32-bit words (whole multiples of 2^-16, and the remainder in units of 2^-40): integer atomics are native on shared
// memory and add in any order to the same bits, and a warp's partial is a float reduced in a fixed order, so the total
// is reproducible bit for bit (a float atomicAdd is a compare-and-swap loop whose result depends on arrival order).
__device__ __forceinline__ void reduce_by_substroke(bool active, int sid, float sum, int cnt, float4 *rec,
                                                    unsigned long long *ink64, int lane) {
    unsigned todo = __ballot_sync(0xffffffffu, active);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int cur = __shfl_sync(0xffffffffu, sid, leader);
        const bool mine = active && sid == cur;
        const float v = warp_sum(mine ? sum : 0.0f);
        const int c = __reduce_add_sync(0xffffffffu, mine ? cnt : 0);
        if (lane == leader) {
            const float t = v * 65536.0f, whole = floorf(t);                  // exact scaling; t < 2^25 (ink of a sub-stroke < 400)
            unsigned *w = reinterpret_cast<unsigned *>(ink64 + cur);
            atomicAdd(w + 1, (unsigned)whole);
            atomicAdd(w, (unsigned)((t - whole) * 16777216.0f));
            atomicAdd(reinterpret_cast<int *>(&rec[cur].w), c);
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}

}  // namespace bpl
