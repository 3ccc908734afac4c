// bpl_render.cu -- fused render+score kernels for pyBPL's P(Image | Token) on B200 (sm_100a).
//
// One CTA renders one token: spline evaluation + sub-stroke chaining (objects/part.py:290-328),
// optional affine warp (util/affine.py:29-55), bounds check / compaction / ink / bilinear splat
// with shared-memory atomics (rendering.py:58-127), 2x broaden + clamp + 2x Gaussian + clamp
// (rendering.py:130-182), epsilon mix (:226-227) and the Bernoulli log-likelihood
// (model/image_dist.py:75-78).  The image lives in shared memory from the first splat to the
// final warp-shuffle reduction; HBM sees ~0.2 KB of control points in and 5 bytes out.
// The *_grad kernel recomputes the forward and runs the adjoint in the same CTA.
//
// No tensor cores: nothing here is a dense contraction.  FP32 issue and
// shared-memory bandwidth bound it; see DESIGN.md for the roofline.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "bpl_common.cuh"

namespace bpl {

// ------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------
// Work queue of one launch of a persistent kernel.  Tokens differ in cost by 3x (1-15 sub-strokes) and a launch of
// the bench batch gives an SM only ~28 of them, so a fixed token -> CTA assignment leaves the slowest SM ~15 %
// behind the mean.  CTAs take their first item by blockIdx.x and every later one from `next`; the CTA that
// finishes last resets the slot, so a slot is all-zero again when its launch completes.
struct Sched { unsigned next, done; };
constexpr int SCHED_RING = 4096;      // lower half: live launches in flight (never near 2 048); upper half: launches captured in CUDA graphs
__device__ Sched g_sched[SCHED_RING];

struct KArgs {
    RenderConsts rc;
    int n_tokens;
    // token (control point) mode
    int neval;
    const int *tok_stroke_off, *stroke_sub_off;
    const float *ctrl, *invscale, *position, *affine, *basis;
    const int *order;        // optional processing order of the tokens (bpl_order_tokens); NULL = index order
    unsigned *cycles_out;    // optional: cycles every token took (team backward), fed back into the next launch's order
    // raw stroke mode
    const int *tok_sub_off, *sub_pt_off;
    const float *pts;
    // common
    const float *eps, *sigma;
    const uint32_t *image_bits;
    const int *img_idx;
    int n_images;
    int all_pairs;           // score every token against all images (forward scoring kernel only)
    float *ll;
    uint8_t *off_page;
    float *pimg;
    // gradients
    const float *g_ll, *g_pimg;
    float *g_ctrl, *g_invscale, *g_position, *g_affine, *g_pts, *g_eps, *g_sigma;
    float *pt_buf;           // phase-timing counters (only read by -DBPL_PHASE_TIMING builds)
    Sched *sched;            // work queue of this launch; NULL = fixed stride (BPL_STATIC_SCHED=1, A/B only)
    // particle sweep: log-sum-exp partial of this launch's scores, produced by the scoring kernel itself
    float *lse;              // [2] (max, sum exp(ll - max)); NULL = not wanted
    float *lse_scratch;      // [2 * gridDim.x] per-CTA partials
    int lse_accumulate;      // merge with the value already in lse (a sweep scored in several launches)
};

// thread 0 only: the item after `cur`
__device__ __forceinline__ int sched_claim(const KArgs &a, int cur) {
    return a.sched ? (int)(atomicAdd(&a.sched->next, 1u) + gridDim.x) : cur + (int)gridDim.x;
}
// thread 0 only: the token to process after work item `item` (which is advanced); n_tokens when none is left
__device__ __forceinline__ int sched_next_token(const KArgs &a, int &item) {
    item = sched_claim(a, item);
    if (item >= a.n_tokens) return a.n_tokens;
    return a.order ? a.order[item] : item;
}
// thread 0 only, once per CTA, after its last claim
__device__ __forceinline__ void sched_finish(const KArgs &a) {
    if (!a.sched) return;
    __threadfence();
    if (atomicAdd(&a.sched->done, 1u) == gridDim.x - 1u) { a.sched->next = 0u; a.sched->done = 0u; }
}

// Online log-sum-exp: (m, s) stands for m + log s.  -inf scores add nothing; NaN (a token beyond the per-CTA limits)
// is skipped.
__device__ __forceinline__ void lse_push(float &m, float &s, float x) {
    if (!(x > -INFINITY)) return;                 // -inf or NaN
    if (x > m) { s = fmaf(s, __expf(m - x), 1.0f); m = x; }
    else s += __expf(x - m);
}
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
    if (!(m2 > -INFINITY) || !(s2 > 0.0f)) return;
    if (m2 > m) { s = fmaf(s, __expf(m - m2), s2); m = m2; }
    else s = fmaf(s2, __expf(m2 - m), s);
}
// Whole CTA, once, after its last token (replaces sched_finish when a.lse is set): every CTA publishes its partial,
// the CTA that finishes last combines them (one warp) and resets the queue slot.
__device__ __forceinline__ void lse_finish(const KArgs &a, float m, float s, int *flag) {
    if (threadIdx.x == 0) {
        a.lse_scratch[2 * blockIdx.x] = m;
        a.lse_scratch[2 * blockIdx.x + 1] = s;
        __threadfence();
        *flag = (atomicAdd(&a.sched->done, 1u) == gridDim.x - 1u);
    }
    __syncthreads();
    if (!*flag || threadIdx.x >= 32) return;
    __threadfence();
    const int lane = threadIdx.x;
    float M = -INFINITY, S = 0.0f;
    for (int i = lane; i < (int)gridDim.x; i += 32) lse_merge(M, S, __ldcg(a.lse_scratch + 2 * i), __ldcg(a.lse_scratch + 2 * i + 1));
    if (a.lse_accumulate && lane == 0) lse_merge(M, S, a.lse[0], a.lse[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, M, o), s2 = __shfl_xor_sync(0xffffffffu, S, o);
        lse_merge(M, S, m2, s2);
    }
    if (lane == 0) { a.lse[0] = M; a.lse[1] = S; a.sched->next = 0u; a.sched->done = 0u; }
}

// Per-CTA bookkeeping in shared memory.
struct Misc {
    double red[4 * 32];      // block_sum scratch
    float4 sub[MAXS];        // per sub-stroke: (t0x,t0y,tlastx,tlasty) then (offx, offy, sum ink, survivors)
    unsigned long long ink64[MAXS];   // the sum of ink accumulated as fixed point (high word: multiples of 2^-16, low word:
                             // units of 2^-40): integer atomics add in any order to the same bits, so the min-ink
                             // branch taken for a borderline sub-stroke (rendering.py:104-107) does not depend on
                             // which warp's atomic lands first
    float dot[MAXS];         // rescale branch: sum_k gink[k] ink0[k] (backward)
    float g[8], ht[8];       // Gaussian half kernel g[d], and h~[d] = (d^2 - mu2) g[d]
    float lp_bg[2], dlp_bg[2];
    float c_eps, c_sigma;    // the values lp_bg / g were last computed for (token_consts)
    float affB[4], com[2];
    float colsum[8];         // column sums of the basis table (affine backward)
    double aff_acc[4];       // sum gx*prex, sum gy*prey, sum gx, sum gy
    double dlp_bg_d[2];      // d lp / d p of a background pixel, analytic in double (for the closed-form d/d epsilon)
    int off_page;
    int cur_img;
    int next_item;           // work queue hand-over (thread 0 -> block)
    int bbox[4];             // rows / columns touched by the splat (min r, max r, min c, max c)
    int n_ones;              // set bits of the current target image (maintained by the sparse scoring kernel)
    int ones_acc;            // accumulator for n_ones while a new image is loaded
    uint32_t bits[IMG_WORDS + 1];
};

// The sub-stroke's total ink after the point stage's barrier: every thread converts the fixed-point sum and stores it
// in the record (all threads store the same bits, so later readers of sub[s].z need no further barrier).
__device__ __forceinline__ float ink_sum(Misc *ms, int s) {
    const unsigned long long w = ms->ink64[s];
    const float v = fmaf((float)(unsigned)(w >> 32), 1.0f / 65536.0f, (float)(unsigned)w * (1.0f / 1099511627776.0f));
    ms->sub[s].z = v;
    return v;
}

enum : int { BR_NONE = 0, BR_SINGLE = 1, BR_FILL = 2, BR_RESCALE = 3, BR_NORMAL = 4 };

// ------------------------------------------------------------------------------------------
// point sources
// ------------------------------------------------------------------------------------------
struct CtrlSrc {   // X = C @ (invscale * shapes) - offset   (splines.py:136, part.py:318-323)
    const float *tab;
    float yx[NCPT], yy[NCPT];
    float offx, offy;
    bool aff;
    float b0, b1, b2, b3;
    __device__ __forceinline__ void pre(int j, float &x, float &y) const {
        float ax = 0.0f, ay = 0.0f;
        const float *row = tab + j * NCPT;
#pragma unroll
        for (int k = 0; k < NCPT; ++k) {          // sequential fma chain == torch.matmul here (App. B.1)
            const float c = row[k];
            ax = __fmaf_rn(c, yx[k], ax);
            ay = __fmaf_rn(c, yy[k], ay);
        }
        x = __fsub_rn(ax, offx);
        y = __fsub_rn(ay, offy);
    }
    __device__ __forceinline__ void point(int j, float &x, float &y) const {
        pre(j, x, y);
        if (aff) {                                  // stk * B[:2] + B[2:]  (affine.py:26)
            x = __fadd_rn(__fmul_rn(x, b0), b2);
            y = __fadd_rn(__fmul_rn(y, b1), b3);
        }
    }
};

struct RawSrc {
    const float2 *p;
    __device__ __forceinline__ void point(int j, float &x, float &y) const {
        const float2 v = p[j];
        x = v.x; y = v.y;
    }
    __device__ __forceinline__ void pre(int j, float &x, float &y) const { point(j, x, y); }
};

// ------------------------------------------------------------------------------------------
// Sub-stroke walk: 32 points per round, one per lane.  Out-of-bounds points are dropped
// BEFORE distances are taken (rendering.py:85-98), so each surviving point needs its previous
// surviving neighbour: inside a round that is the nearest lower lane whose ballot bit is set,
// across rounds it is carried in warp-uniform registers.
// ------------------------------------------------------------------------------------------
struct Walk {
    int cnt;          // survivors so far
    float cr, cc;     // last survivor (row, col)
    int cj;           // its point index
};

struct Pt {
    bool valid, inb;
    int j, idx, pj;           // point index, survivor rank, previous survivor's point index
    float r, c, pr, pc;       // image coords (row = -y, col = x) and previous survivor's
    float dr, dc, dist, ink0; // segment to the previous survivor; ink0 = ink_pp*min(dist,max)/max
};

template <class Src>
__device__ __forceinline__ Pt walk_round(const Src &src, int n, int base, int lane, const RenderConsts &rc, Walk &w) {
    Pt p;
    p.j = base + lane;
    p.valid = p.j < n;
    float x = 0.0f, y = 0.0f;
    if (p.valid) src.point(p.j, x, y);
    p.r = -y;   // space_motor_to_img (rendering.py:52-53): flip, times (-1, 1)
    p.c = x;
    // check_bounds (rendering.py:27-31): floor<0 | ceil>=105  <=>  not (0 <= v <= 104)
    p.inb = p.valid && (p.r >= 0.0f) && (p.r <= rc.rmax) && (p.c >= 0.0f) && (p.c <= rc.cmax);
    const unsigned m = __ballot_sync(0xffffffffu, p.inb);
    const unsigned lower = m & ((1u << lane) - 1u);
    const int srcl = lower ? (31 - __clz(lower)) : 0;
    p.pr = __shfl_sync(0xffffffffu, p.r, srcl);
    p.pc = __shfl_sync(0xffffffffu, p.c, srcl);
    p.pj = base + srcl;
    if (!lower) { p.pr = w.cr; p.pc = w.cc; p.pj = w.cj; }
    p.idx = w.cnt + __popc(lower);
    p.dr = 0.0f; p.dc = 0.0f; p.dist = 0.0f; p.ink0 = 0.0f;
    if (p.inb && p.idx >= 1) {
        p.dr = __fsub_rn(p.r, p.pr);
        p.dc = __fsub_rn(p.c, p.pc);
        p.dist = __fsqrt_rn(__fadd_rn(__fmul_rn(p.dr, p.dr), __fmul_rn(p.dc, p.dc)));   // torch.norm (:96)
        const float d = fminf(p.dist, rc.ink_max_dist);                                   // clamp (:97)
        p.ink0 = __fdiv_rn(__fmul_rn(rc.ink_pp, d), rc.ink_max_dist);                      // (:99)
    }
    if (m) {
        const int last = 31 - __clz(m);
        w.cr = __shfl_sync(0xffffffffu, p.r, last);
        w.cc = __shfl_sync(0xffffffffu, p.c, last);
        w.cj = base + last;
        w.cnt += __popc(m);
    }
    return p;
}

struct Stats {
    int cnt;        // surviving points
    int branch;
    float sum;      // sum of ink before the min-ink fix-up (rendering.py:103)
    float first;    // ink of survivor 0 (= ink of survivor 1, rendering.py:98)
    float scale;    // rescale factor ink_pp / sum
    float fill;     // ink_pp / n
    bool any_out;
};

template <class Src>
__device__ __forceinline__ Stats substroke_stats(const Src &src, int n, int lane, const RenderConsts &rc) {
    Walk w{0, 0.0f, 0.0f, 0};
    float psum = 0.0f, first = 0.0f;
    bool any_out = false;
    for (int base = 0; base < n; base += 32) {
        const Pt p = walk_round(src, n, base, lane, rc, w);
        any_out |= (p.valid && !p.inb);
        if (p.inb && p.idx >= 1) psum += p.ink0;
        const unsigned b1 = __ballot_sync(0xffffffffu, p.inb && p.idx == 1);
        if (b1) first = __shfl_sync(0xffffffffu, p.ink0, __ffs(b1) - 1);
    }
    Stats s;
    s.cnt = w.cnt;
    s.any_out = __any_sync(0xffffffffu, any_out);
    s.first = first;
    s.sum = warp_sum(psum) + first;
    s.scale = 1.0f;
    s.fill = 0.0f;
    if (s.cnt == 0) {
        s.branch = BR_NONE;
    } else if (s.cnt == 1) {                       // myink = [ink_pp]  (:93-94)
        s.branch = BR_SINGLE;
        s.first = rc.ink_pp;
        s.sum = rc.ink_pp;
    } else if (s.sum < 2.22e-6f) {                 // (:104-105)
        s.branch = BR_FILL;
        s.fill = (float)((double)rc.ink_pp / (double)s.cnt);
    } else if (s.sum < rc.ink_pp) {                // (:106-107)
        s.branch = BR_RESCALE;
        s.scale = __fdiv_rn(rc.ink_pp, s.sum);
    } else {
        s.branch = BR_NORMAL;
    }
    return s;
}

__device__ __forceinline__ float final_ink(const Pt &p, const Stats &s) {
    float ink = (p.idx == 0) ? s.first : p.ink0;
    if (s.branch == BR_FILL) ink = s.fill;
    else if (s.branch == BR_RESCALE) ink = __fmul_rn(s.scale, ink);
    return ink;
}

// bilinear splat with shared-memory atomics (rendering.py:111-125)
template <int ORG, int STR, class Src>
__device__ __forceinline__ void substroke_splat(const Src &src, int n, int lane, const RenderConsts &rc,
                                                const Stats &s, float *img) {
    Walk w{0, 0.0f, 0.0f, 0};
    for (int base = 0; base < n; base += 32) {
        const Pt p = walk_round(src, n, base, lane, rc, w);
        if (!p.inb) continue;
        const float ink = final_ink(p, s);
        const float rf = floorf(p.r), cf = floorf(p.c), rce = ceilf(p.r), cce = ceilf(p.c);
        const float fr = __fsub_rn(p.r, rf), fc = __fsub_rn(p.c, cf);
        const float gr = __fsub_rn(1.0f, fr), gc = __fsub_rn(1.0f, fc);
        const int r0 = (int)rf, c0 = (int)cf, r1 = (int)rce, c1 = (int)cce;
        atomicAdd(img + ORG + r0 * STR + c0, __fmul_rn(__fmul_rn(ink, gr), gc));
        atomicAdd(img + ORG + r1 * STR + c0, __fmul_rn(__fmul_rn(ink, fr), gc));
        atomicAdd(img + ORG + r0 * STR + c1, __fmul_rn(__fmul_rn(ink, gr), fc));
        atomicAdd(img + ORG + r1 * STR + c1, __fmul_rn(__fmul_rn(ink, fr), fc));
    }
}

// ------------------------------------------------------------------------------------------
// token geometry shared by forward and backward
// ------------------------------------------------------------------------------------------
struct TokGeom {
    int k0, K;      // strokes [k0, k0+K)
    int s0, S;      // sub-strokes [s0, s0+S)   (raw mode: strokes)
};

template <bool RAW>
__device__ __forceinline__ TokGeom token_geom(const KArgs &a, int p) {
    TokGeom t;
    if (RAW) {
        t.k0 = 0; t.K = 0;
        t.s0 = a.tok_sub_off[p];
        t.S = a.tok_sub_off[p + 1] - t.s0;
    } else {
        t.k0 = a.tok_stroke_off[p];
        t.K = a.tok_stroke_off[p + 1] - t.k0;
        t.s0 = a.stroke_sub_off[t.k0];
        t.S = a.stroke_sub_off[t.k0 + t.K] - t.s0;
    }
    return t;
}

// Sub-stroke chaining (part.py:316-326): offset_i = traj_i[0] - prev_i, prev_{i+1} = traj_i[-1] - offset_i.
// Phase 1 (one thread per sub-stroke): first and last trajectory points.  Phase 2 (one thread
// per stroke): the serial recurrence, two subtractions per sub-stroke and coordinate.
__device__ __forceinline__ void chain_offsets(const KArgs &a, const TokGeom &t, const float *tab, Misc *ms) {
    for (int s = threadIdx.x; s < t.S; s += blockDim.x) {
        const float inv = a.invscale[t.s0 + s];
        const float *cp = a.ctrl + (size_t)(t.s0 + s) * NCPT * 2;
        float ax = 0.0f, ay = 0.0f, bx = 0.0f, by = 0.0f;
        const float *r0 = tab, *r1 = tab + (a.neval - 1) * NCPT;
#pragma unroll
        for (int k = 0; k < NCPT; ++k) {
            const float yx = __fmul_rn(inv, cp[2 * k]), yy = __fmul_rn(inv, cp[2 * k + 1]);
            ax = __fmaf_rn(r0[k], yx, ax); ay = __fmaf_rn(r0[k], yy, ay);
            bx = __fmaf_rn(r1[k], yx, bx); by = __fmaf_rn(r1[k], yy, by);
        }
        ms->sub[s] = make_float4(ax, ay, bx, by);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < t.K; k += blockDim.x) {
        float px = a.position[2 * (t.k0 + k)], py = a.position[2 * (t.k0 + k) + 1];
        const int b = a.stroke_sub_off[t.k0 + k] - t.s0, e = a.stroke_sub_off[t.k0 + k + 1] - t.s0;
        for (int s = b; s < e; ++s) {
            const float4 v = ms->sub[s];
            const float ox = __fsub_rn(v.x, px), oy = __fsub_rn(v.y, py);
            px = __fsub_rn(v.z, ox); py = __fsub_rn(v.w, oy);
            ms->sub[s] = make_float4(ox, oy, 0.0f, 0.0f);
            ms->ink64[s] = 0ull;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ CtrlSrc make_ctrl_src(const KArgs &a, const TokGeom &t, int s, const float *tab,
                                                 const Misc *ms, bool aff) {
    CtrlSrc src;
    src.tab = tab;
    const float inv = a.invscale[t.s0 + s];
    const float *cp = a.ctrl + (size_t)(t.s0 + s) * NCPT * 2;
#pragma unroll
    for (int k = 0; k < NCPT; ++k) {
        src.yx[k] = __fmul_rn(inv, cp[2 * k]);        // invscales[i] * shapes[:,:,i]  (part.py:318)
        src.yy[k] = __fmul_rn(inv, cp[2 * k + 1]);
    }
    const float4 o = ms->sub[s];
    src.offx = o.x; src.offy = o.y;
    src.aff = aff;
    src.b0 = ms->affB[0]; src.b1 = ms->affB[1]; src.b2 = ms->affB[2]; src.b3 = ms->affB[3];
    return src;
}

__device__ __forceinline__ CtrlSrc make_ctrl_src_p(const KArgs &a, const TokGeom &t, int s, const float *tab,
                                                   const float4 off, const float *affB, bool aff) {
    CtrlSrc src;
    src.tab = tab;
    const float inv = a.invscale[t.s0 + s];
    const float *cp = a.ctrl + (size_t)(t.s0 + s) * NCPT * 2;
#pragma unroll
    for (int k = 0; k < NCPT; ++k) {
        src.yx[k] = __fmul_rn(inv, cp[2 * k]);
        src.yy[k] = __fmul_rn(inv, cp[2 * k + 1]);
    }
    src.offx = off.x; src.offy = off.y;
    src.aff = aff;
    src.b0 = affB[0]; src.b1 = affB[1]; src.b2 = affB[2]; src.b3 = affB[3];
    return src;
}

// com_char (util/stroke.py:134-135) and B (affine.py:48-53).  All S*neval points, in-bounds or not.
__device__ __forceinline__ void affine_setup(const KArgs &a, const TokGeom &t, int p, const float *tab, Misc *ms) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double sx = 0.0, sy = 0.0;
    for (int s = warp; s < t.S; s += nw) {
        const CtrlSrc src = make_ctrl_src(a, t, s, tab, ms, false);
        for (int j = lane; j < a.neval; j += 32) {
            float x, y;
            src.pre(j, x, y);
            sx += (double)x; sy += (double)y;
        }
    }
    double v[2] = {sx, sy};
    block_sum<2>(v, ms->red);
    if (threadIdx.x == 0) {
        const double npts = (double)t.S * (double)a.neval;
        const float cx = npts > 0.0 ? (float)(v[0] / npts) : 0.0f, cy = npts > 0.0 ? (float)(v[1] / npts) : 0.0f;      // (no points: nothing depends on the warp)
        const float *A = a.affine + 4 * (size_t)p;
        ms->com[0] = cx; ms->com[1] = cy;
        ms->affB[0] = A[0]; ms->affB[1] = A[1];
        ms->affB[2] = __fsub_rn(A[2], __fmul_rn(__fsub_rn(A[0], 1.0f), cx));
        ms->affB[3] = __fsub_rn(A[3], __fmul_rn(__fsub_rn(A[1], 1.0f), cy));
    }
    __syncthreads();
}

// Per-token constants computed by a few lanes of the last warp while the others zero buffers:
// the 1-D Gaussian g (K = g (x) g exactly: fspecial normalises by sum(E) = (sum e)^2,
// util/general.py:196-207), h~ for d/dsigma, and the background log-prob constants.
__device__ __noinline__ void token_consts_eps(float eps, bool x, Misc *ms);
__device__ __noinline__ void token_consts_sigma(float sigma, int lane, int fhalf, Misc *ms);
__device__ __forceinline__ void token_consts(float eps, float sigma, bool grad, Misc *ms, int fhalf, int wsel = 0) {
    const int lane = threadIdx.x & 31;
    if ((int)(threadIdx.x >> 5) != (int)(blockDim.x >> 5) - 1 - wsel) return;     // wsel-th warp from the end
    // Both sets of constants are kept from one token to the next (only this warp reads or writes c_eps / c_sigma):
    // epsilon is usually one value for the whole batch, and the double-precision libm code below is several
    // thousand instructions that would otherwise pass through the instruction cache once per token.
    const bool same_eps = (eps == ms->c_eps), same_sigma = (sigma == ms->c_sigma);
    __syncwarp();
    if (lane == 0) { ms->c_eps = eps; ms->c_sigma = sigma; }
    if (!same_sigma) token_consts_sigma(sigma, lane, fhalf, ms);
    if (!same_eps && (lane == 8 || lane == 9)) token_consts_eps(eps, lane == 9, ms);
    (void)grad;
}

__device__ __noinline__ void token_consts_eps(float eps, bool x, Misc *ms) {
    // background pixel (pimg == eps after the mix, or 0 when eps <= 0)
    const float bg = eps > 0.0f ? eps : 0.0f;
    float lp, dlp;
    pixel_lp_exact(bg, x, lp, dlp);
    ms->lp_bg[x] = lp;
    ms->dlp_bg[x] = dlp;
    ms->dlp_bg_d[x] = (bg < TORCH_EPS || bg > 1.0f - TORCH_EPS) ? 0.0 : (x ? 1.0 / (double)bg : -1.0 / (1.0 - (double)bg));
}

__device__ __noinline__ void token_consts_sigma(float sigma, int lane, int fhalf, Misc *ms) {
    // lanes 0..fhalf: e[d] = exp(-0.5 (d/sigma)^2); the taps beyond the filter's half-width (fsize < 11) are zero, so the
    // 11-tap passes compute the fsize-tap filter exactly (fspecial normalises over its own fsize x fsize support)
    double e = 0.0;
    if (lane <= fhalf && sigma > 0.0f) {
        const double q = (double)lane / (double)sigma;
        e = exp(-0.5 * q * q);
    }
    double tot = (lane == 0) ? e : 2.0 * e;
    double m2 = (lane == 0) ? 0.0 : 2.0 * e * (double)(lane * lane);
    tot = warp_sum(tot);
    m2 = warp_sum(m2);
    if (lane <= PAD && sigma > 0.0f) {
        const double g = e / tot;
        const double mu2 = m2 / tot;      // sum_u u^2 g_u
        ms->g[lane] = (float)g;
        ms->ht[lane] = (float)(((double)(lane * lane) - mu2) * g);
    }
}

__device__ __forceinline__ void load_image_bits(const KArgs &a, int p, Misc *ms) {
    if (!a.image_bits) return;
    const int t = a.img_idx ? a.img_idx[p] : 0;
    if (t == ms->cur_img) return;      // cur_img is only written after the barrier below, read before it
    int ones = 0;
    for (int i = threadIdx.x; i < IMG_WORDS; i += blockDim.x) {
        const uint32_t wv = a.image_bits[(size_t)t * IMG_WORDS + i];
        ms->bits[i] = wv;
        ones += __popc(wv);
    }
    ones = __reduce_add_sync(0xffffffffu, ones);      // every thread of the block is here (cur_img is block-uniform)
    if ((threadIdx.x & 31) == 0 && ones) atomicAdd(&ms->ones_acc, ones);
}

__device__ __forceinline__ void zero_floats(float *p, int n) {
    float4 *p4 = reinterpret_cast<float4 *>(p);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) p4[i] = z;
    for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) p[i] = 0.0f;
}

// broaden (rendering.py:160-171): ncon x [ c1 * S_h S_v + c2 ], then min(.,1) if CLAMP.
// A holds the image and receives the result; T is scratch.
template <int R3>
__device__ __forceinline__ void broaden(float *A, float *T, const RenderConsts &rc, const Box &bx) {
    const float w3[1][2] = {{2.0f, 1.0f}};
    for (int it = 0; it < rc.ink_ncon; ++it) {
        sweep_box<3, R3, true, 1>(A, w3, LoadId(), [&](int r, int c, const float (&acc)[1]) {
            T[ORIGIN + r * STRIDE + c] = acc[0];
        }, bx);
        __syncthreads();
        sweep_box<3, R3, false, 1>(T, w3, LoadId(), [&](int r, int c, const float (&acc)[1]) {
            float *q = A + ORIGIN + r * STRIDE + c;
            *q = fmaf(rc.c1, acc[0], rc.c2 * (*q));
        }, bx);
        __syncthreads();
    }
}

}  // namespace bpl

#include "bpl_forward.cuh"
#include "bpl_forward1s.cuh"

namespace bpl {

// ------------------------------------------------------------------------------------------
// fused forward + backward kernel (4 image buffers, one CTA per SM)
// ------------------------------------------------------------------------------------------
#ifndef BPL_BWD_THREADS
#define BPL_BWD_THREADS 640
#endif
constexpr int BWD_THREADS = BPL_BWD_THREADS;   // 20 warps (measured 384 / 512 / 640 / 768 threads: 3.57 / 3.70 / 3.81 / 3.65 M particles/s; a box rarely has more than 375 sweep items, a full canvas takes two rounds)
#ifndef BPL_BWD_R
#define BPL_BWD_R 15
#endif
constexpr int BWD_R = BPL_BWD_R;      // outputs per sweep item = granularity of the box (measured 15 / 21 / 35 below)
constexpr int BWD_R3 = BPL_BWD_R;
#ifndef BPL_BWD_PPL
#define BPL_BWD_PPL 3
#endif
constexpr int BWD_PPL = BPL_BWD_PPL;         // points per lane in the backward kernel's point stages: the CTA has the SM to
                                   // itself, so shorter runs (more lanes busy) beat fewer atomic collisions
constexpr int NACC = 2 * NCPT + 2; // per sub-stroke: W[k].x, W[k].y, Gsum.x, Gsum.y

struct BwdSmem {
    float A[BUFP];
    float B[BUFP];
    float C[BUFP];
    float D[BUFP];
    float tab[MAX_TABLE];
    Misc ms;
};

// gather the image gradient at the four splat corners (adjoint of rendering.py:111-125)
struct Corner {
    float gink, gr, gc;     // d/d ink, d/d row, d/d col (for ink weight 1)
};
using LayPadded = LayS<1, STRIDE, ORIGIN>;      // the full-canvas backward's padded buffers
template <class Lay = LayPadded>
__device__ __forceinline__ Corner gather_corners(const float *G, float r, float c, const Lay lay = Lay()) {
    const float rf = floorf(r), cf = floorf(c), rce = ceilf(r), cce = ceilf(c);
    const float fr = r - rf, fc = c - cf, gr = 1.0f - fr, gc = 1.0f - fc;
    const int r0 = (int)rf, c0 = (int)cf, r1 = (int)rce, c1 = (int)cce;
    const float G00 = G[lay(r0, c0)], G10 = G[lay(r1, c0)];
    const float G01 = G[lay(r0, c1)], G11 = G[lay(r1, c1)];
    Corner k;
    k.gink = gr * gc * G00 + fr * gc * G10 + gr * fc * G01 + fr * fc * G11;
    k.gr = gc * (G10 - G00) + fc * (G11 - G01);
    k.gc = gr * (G01 - G00) + fr * (G11 - G10);
    return k;
}

// Backward of add_stroke for one sub-stroke.  `emit(j, gx, gy)` receives motor-space (post-warp)
// point gradients; a point may be emitted more than once (own splat, own segment, next segment).
template <class Src, class Emit, class Lay = LayPadded>
__device__ __forceinline__ void substroke_backward(const Src &src, int n, int lane, const RenderConsts &rc,
                                                   const Stats &st, const float *G, Emit emit, const Lay lay = Lay()) {
    if (st.cnt == 0) return;
    const bool use_dist = (st.branch == BR_NORMAL || st.branch == BR_RESCALE);
    float sc = 1.0f, gsum = 0.0f;
    if (st.branch == BR_RESCALE) {      // ink' = ink * (ink_pp / sum)   (rendering.py:106-107)
        Walk w{0, 0.0f, 0.0f, 0};
        float dot = 0.0f;
        for (int base = 0; base < n; base += 32) {
            const Pt p = walk_round(src, n, base, lane, rc, w);
            if (p.inb) {
                const Corner k = gather_corners(G, p.r, p.c, lay);
                dot += k.gink * ((p.idx == 0) ? st.first : p.ink0);
            }
        }
        dot = warp_sum(dot);
        sc = st.scale;
        gsum = -dot * rc.ink_pp / (st.sum * st.sum);
    }
    Walk w{0, 0.0f, 0.0f, 0};
    float cgink = 0.0f;      // d/d ink0 of the last survivor of earlier rounds
    for (int base = 0; base < n; base += 32) {
        const Pt p = walk_round(src, n, base, lane, rc, w);
        float ginkp = 0.0f, g_r = 0.0f, g_c = 0.0f;
        if (p.inb) {
            const Corner k = gather_corners(G, p.r, p.c, lay);
            const float ink = final_ink(p, st);
            g_r = ink * k.gr;
            g_c = ink * k.gc;
            ginkp = use_dist ? fmaf(k.gink, sc, gsum) : 0.0f;
        }
        // d/d ink0 of the previous survivor (needed only by survivor 1: ink0[0] = ink0[1])
        const unsigned m = __ballot_sync(0xffffffffu, p.inb);
        const unsigned lower = m & ((1u << lane) - 1u);
        float pgink = __shfl_sync(0xffffffffu, ginkp, lower ? (31 - __clz(lower)) : 0);
        if (!lower) pgink = cgink;
        if (m) cgink = __shfl_sync(0xffffffffu, ginkp, 31 - __clz(m));
        if (p.inb) {
            float ur = 0.0f, uc = 0.0f;
            if (use_dist && p.idx >= 1 && p.dist <= rc.ink_max_dist && p.dist != 0.0f) {
                const float gd = rc.ink_f * (ginkp + (p.idx == 1 ? pgink : 0.0f));
                ur = gd * (p.dr / p.dist);
                uc = gd * (p.dc / p.dist);
            }
            // un-flip: row = -y, col = x
            emit(p.j, g_c + uc, -(g_r + ur));
            if (ur != 0.0f || uc != 0.0f) emit(p.pj, -uc, ur);
        }
    }
}

template <bool RAW>
__global__ void __launch_bounds__(BWD_THREADS, 1) bpl_backward_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    float *A = sm.A, *B = sm.B, *C = sm.C, *D = sm.D;
    Misc *ms = &sm.ms;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;

    zero_floats(A, 4 * BUFP);    // A, B, C, D are contiguous; every token leaves them all-zero again (see the end of the loop)
    if (!RAW) {
        for (int i = threadIdx.x; i < a.neval * NCPT; i += blockDim.x) sm.tab[i] = a.basis[i];
    }
    if (threadIdx.x == 0) { ms->cur_img = -1; ms->c_eps = ms->c_sigma = nanf(""); ms->n_ones = 0; ms->ones_acc = 0; }
    __syncthreads();
    if (!RAW && warp == 0) {     // column sums of the table (d com / d control points)
        float cs[NCPT];
#pragma unroll
        for (int k = 0; k < NCPT; ++k) cs[k] = 0.0f;
        for (int j = lane; j < a.neval; j += 32)
#pragma unroll
            for (int k = 0; k < NCPT; ++k) cs[k] += sm.tab[j * NCPT + k];
#pragma unroll
        for (int k = 0; k < NCPT; ++k) {
            const float s = warp_sum(cs[k]);
            if (lane == 0) ms->colsum[k] = s;
        }
    }

#ifdef BPL_PHASE_TIMING
    long long pt_acc[12] = {0};
    long long pt_t0 = clock64();
#endif
    int item = blockIdx.x;                                      // position in the work queue (advanced by thread 0 only)
    int p = a.order ? a.order[item] : item;                     // the token it stands for
    while (p < a.n_tokens) {
        int p_next = 0;
        if (threadIdx.x == 0) p_next = sched_next_token(a, item);      // consumed at the end of the iteration
        const TokGeom t = token_geom<RAW>(a, p);
        const float eps = a.eps[p], sigma = a.sigma[p];
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
        const bool aff = !RAW && a.affine != nullptr;
        load_image_bits(a, p, ms);
        if (threadIdx.x == 0) {
            ms->off_page = 0;
            ms->aff_acc[0] = ms->aff_acc[1] = ms->aff_acc[2] = ms->aff_acc[3] = 0.0;
            ms->bbox[0] = IMG; ms->bbox[1] = -1; ms->bbox[2] = IMG; ms->bbox[3] = -1;
        }
        if (!RAW) {
            chain_offsets(a, t, sm.tab, ms);
            if (aff) affine_setup(a, t, p, sm.tab, ms);
        } else {
            __syncthreads();
        }
        token_consts(eps, sigma, true, ms, a.rc.fhalf);      // double-precision constants by one warp, hidden behind the splat
        if (threadIdx.x == 0 && a.image_bits) {
            const int img = a.img_idx ? a.img_idx[p] : 0;
            if (img != ms->cur_img) { ms->n_ones = ms->ones_acc; ms->cur_img = img; }      // a new image was loaded above
        }

        BPL_PT(0);      // setup + chain offsets
        // ---- forward: splat ----
        bool any_out = false;
        if (RAW) {
            for (int s = warp; s < t.S; s += nw) {
                const int o = a.sub_pt_off[t.s0 + s], n = a.sub_pt_off[t.s0 + s + 1] - o;
                const RawSrc src{reinterpret_cast<const float2 *>(a.pts) + o};
                const Stats st = substroke_stats(src, n, lane, a.rc);
                any_out |= st.any_out;
                if (st.cnt > 0) substroke_splat<ORIGIN, STRIDE>(src, n, lane, a.rc, st, A);
            }
            any_out = __any_sync(0xffffffffu, any_out);
            if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
            if (threadIdx.x == 0) { ms->bbox[0] = 0; ms->bbox[1] = IMG - 1; ms->bbox[2] = 0; ms->bbox[3] = IMG - 1; }      // raw mode: whole canvas
            __syncthreads();
        } else {
            const int QN = (a.neval + BWD_PPL - 1) / BWD_PPL, items = t.S * QN;
            for (int base = 0; base < items; base += blockDim.x) {
                const int item = base + threadIdx.x;
                const bool active = item < items;
                const int s = active ? item / QN : 0x7fffffff;
                CtrlSrc src{};
                if (active) src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
                const LaneAcc acc = warp_points<false, 1, STRIDE, ORIGIN, BWD_PPL>(active, src, sm.tab, s, a.neval,
                                                                          active ? item - s * QN : 0, lane, a.rc, A, 0.0f, 0,
                                                                          ms->bbox);
                any_out |= acc.any_out;
                reduce_by_substroke(active, s, acc.sum, acc.cnt, ms->sub, ms->ink64, lane);
            }
            any_out = __any_sync(0xffffffffu, any_out);
            if (any_out && lane == 0) atomicOr(&ms->off_page, 1);
            __syncthreads();
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
                    warp_points<true, 1, STRIDE, ORIGIN, BWD_PPL>(active, src, sm.tab, s, a.neval, item < items ? item - s * QN : 0,
                                                         lane, a.rc, A, v.z, __float_as_int(v.w));
                }
                __syncthreads();
            }
        }

        BPL_PT(1);      // forward splat (+fix)
        // Everything from here to the gather stage works on one box of the canvas: the pixels the splat touched
        // grown by the reach of all stencil passes (ink_ncon + 2 * 5).  Forward values are exactly zero outside it.
        // The cotangent d L / d pimg is NOT zero outside it, but the gather stage only reads the adjoint at the
        // splat's own pixels, and those are exact when the cotangent is known on the grown box (each adjoint pass
        // gives up its reach at the box's rim: 5 + 5 + 5 + 5 + ncon + ncon along each axis in total).  The scalar
        // outputs (ll, d/d epsilon) get the closed-form contribution of an all-background image and sum
        // differences to it inside the box; d/d sigma has no contribution outside it.
        Box bx{ms->bbox[0], ms->bbox[1], ms->bbox[2], ms->bbox[3]};
        if (bx.r0 <= bx.r1) {
            const int reach = a.rc.ink_ncon + 2 * PAD;
            bx.r0 = max(0, bx.r0 - reach); bx.r1 = min(IMG - 1, bx.r1 + reach);
            bx.c0 = max(0, bx.c0 - reach); bx.c1 = min(IMG - 1, bx.c1 + reach);
            // rounded out to whole bands of the sweeps (105 = 7 * 15, so this stays on the canvas)
            static_assert(BWD_R == BWD_R3, "one band size for all passes");
            bx.r0 = bx.r0 / BWD_R * BWD_R; bx.r1 = bx.r1 / BWD_R * BWD_R + BWD_R - 1;
            bx.c0 = bx.c0 / BWD_R * BWD_R; bx.c1 = bx.c1 / BWD_R * BWD_R + BWD_R - 1;
        }
        // ---- forward: broaden; A keeps I2 (pre-clamp); consumers apply min(.,1) on load ----
        broaden<BWD_R3>(A, B, a.rc, bx);

        BPL_PT(2);      // forward broaden
        const float bgv = eps > 0.0f ? eps : 0.0f;
        const float om = 1.0f - eps;
        const float mixd = eps > 0.0f ? (1.0f - 2.0f * eps) : 1.0f;     // d mix / d I6
        const float lp_bg0 = ms->lp_bg[0], lp_bg1 = ms->lp_bg[1];
        const float dlp_bg0 = ms->dlp_bg[0], dlp_bg1 = ms->dlp_bg[1];
        const bool score = a.image_bits != nullptr;
        const float *gP_ext = a.g_pimg ? a.g_pimg + (size_t)p * NPIX : nullptr;
        const float gscale = (!gP_ext && a.g_ll) ? a.g_ll[p] : 1.0f;
        const double dlp_bgd0 = ms->dlp_bg_d[0], dlp_bgd1 = ms->dlp_bg_d[1];
        float acc_ll = 0.0f, acc_sig = 0.0f;
        double acc_eps = 0.0;      // terms of +-1e4 that cancel to a few tens: summed in double

        // pointwise tail of the forward + head of the backward for pixel (r,c) with blurred value v:
        // returns d L / d I5 (cotangent entering the second blur pass).
        auto tail = [&](int r, int c, float v) -> float {
            const bool pass5 = (v >= 0.0f) && (v <= 1.0f);                  // clamp_(0,1) passes gradient on [0,1]
            const float i6 = fminf(fmaxf(v, 0.0f), 1.0f);
            float pm = i6;
            if (eps > 0.0f) pm = __fadd_rn(__fmul_rn(om, i6), __fmul_rn(eps, __fsub_rn(1.0f, i6)));
            float gp;
            if (gP_ext) {
                gp = gP_ext[r * IMG + c];
                if (eps > 0.0f) acc_eps += (double)(-2.0f * gp * i6);        // d/d eps of (1-eps) p + eps (1-p) is 1 - 2 p;
                                                                            // the sum of gp over the canvas is added below
            } else {
                const bool x = score && ((ms->bits[(r * IMG + c) >> 5] >> ((r * IMG + c) & 31)) & 1u);
                const float lb = x ? lp_bg1 : lp_bg0, db = x ? dlp_bg1 : dlp_bg0;
                float lp = lb, d = db;
                if (pm != bgv) {                                             // a background pixel adds nothing to the sum
                    lp = pixel_lp_fwd(pm, x);                                // the forward kernels' single bounded-error logarithm
                    d = (pm < TORCH_EPS || pm > 1.0f - TORCH_EPS) ? 0.0f : (x ? __frcp_rn(pm) : -__frcp_rn(1.0f - pm));
                    acc_ll += lp - lb;
                }
                // A pixel without ink has the analytic background value (closed form below); pm == bgv alone does not
                // mean "no ink" (eps = 0.5 maps every v to 0.5).
                if (eps > 0.0f && (pm != bgv || i6 != 0.0f))
                    acc_eps += ((double)(d * (1.0f - 2.0f * i6)) - (x ? dlp_bgd1 : dlp_bgd0)) * (double)gscale;
                gp = d * gscale;
            }
            return pass5 ? gp * mixd : 0.0f;
        };

        if (sigma > 0.0f) {
            float wg[1][PAD + 1], wgh[2][PAD + 1];
#pragma unroll
            for (int d = 0; d <= PAD; ++d) { wg[0][d] = ms->g[d]; wgh[0][d] = ms->g[d]; wgh[1][d] = ms->ht[d]; }
            // blur A forward: I3 = min(A,1) -V-> B -H-> C = I4
            sweep_box<BPL_FSIZE, BWD_R, true, 1>(A, wg, LoadMin1(), [&](int r, int c, const float (&acc)[1]) {
                B[ORIGIN + r * STRIDE + c] = acc[0]; }, bx);
            __syncthreads();
            sweep_box<BPL_FSIZE, BWD_R, false, 1>(B, wg, LoadId(), [&](int r, int c, const float (&acc)[1]) {
                C[ORIGIN + r * STRIDE + c] = acc[0]; }, bx);
            __syncthreads();
            // blur B forward, first axis with both filters: U = g_v * I4 -> B, W = h~_v * I4 -> D
            sweep_box<BPL_FSIZE, BWD_R, true, 2>(C, wgh, LoadId(), [&](int r, int c, const float (&acc)[2]) {
                B[ORIGIN + r * STRIDE + c] = acc[0];
                D[ORIGIN + r * STRIDE + c] = acc[1]; }, bx);
            __syncthreads();
            // second axis: I5 = g_h * U; pointwise tail; G_B -> C
            sweep_box<BPL_FSIZE, BWD_R, false, 1>(B, wg, LoadId(), [&](int r, int c, const float (&acc)[1]) {
                C[ORIGIN + r * STRIDE + c] = tail(r, c, acc[0]); }, bx);
            __syncthreads();
            // adjoint of blur B along h, reduced against U, W for d/d sigma; t_g -> D (in place of W)
            sweep_box<BPL_FSIZE, BWD_R, false, 2>(C, wgh, LoadId(), [&](int r, int c, const float (&acc)[2]) {
                const int q = ORIGIN + r * STRIDE + c;
                acc_sig = fmaf(acc[0], D[q], fmaf(acc[1], B[q], acc_sig));
                D[q] = acc[0]; }, bx);
            __syncthreads();
            // adjoint along v: G_A = d L / d I4 -> C
            sweep_box<BPL_FSIZE, BWD_R, true, 1>(D, wg, LoadId(), [&](int r, int c, const float (&acc)[1]) {
                C[ORIGIN + r * STRIDE + c] = acc[0]; }, bx);
            __syncthreads();
            // recompute U_A, W_A from I3 = min(A,1)
            sweep_box<BPL_FSIZE, BWD_R, true, 2>(A, wgh, LoadMin1(), [&](int r, int c, const float (&acc)[2]) {
                B[ORIGIN + r * STRIDE + c] = acc[0];
                D[ORIGIN + r * STRIDE + c] = acc[1]; }, bx);
            __syncthreads();
            sweep_box<BPL_FSIZE, BWD_R, false, 2>(C, wgh, LoadId(), [&](int r, int c, const float (&acc)[2]) {
                const int q = ORIGIN + r * STRIDE + c;
                acc_sig = fmaf(acc[0], D[q], fmaf(acc[1], B[q], acc_sig));
                D[q] = acc[0]; }, bx);
            __syncthreads();
            // G_3 = d L / d I3, masked by clamp_(max=1) (:171) -> A (pointwise in place over I2)
            sweep_box<BPL_FSIZE, BWD_R, true, 1>(D, wg, LoadId(), [&](int r, int c, const float (&acc)[1]) {
                float *q = A + ORIGIN + r * STRIDE + c;
                *q = (*q <= 1.0f) ? acc[0] : 0.0f; }, bx);
            __syncthreads();
        } else {
            const int bw = bx.c1 - bx.c0 + 1, bn = bx.empty() ? 0 : bw * (bx.r1 - bx.r0 + 1);
            for (int i = threadIdx.x; i < bn; i += blockDim.x) {
                const int r = bx.r0 + i / bw, c = bx.c0 + i % bw;
                float *q = A + ORIGIN + r * STRIDE + c;
                const float i2 = *q;
                const float g = tail(r, c, fminf(i2, 1.0f));
                *q = (i2 <= 1.0f) ? g : 0.0f;
            }
            __syncthreads();
        }
        BPL_PT(3);      // blur forward + tail + adjoint blur (9 sweeps)
        // adjoint broaden (symmetric kernel, zero padding): A -> A
        broaden<BWD_R3>(A, B, a.rc, bx);

        BPL_PT(4);      // adjoint broaden
        // ---- scalar outputs ----
        {
            if (gP_ext && eps > 0.0f)
                for (int i = threadIdx.x; i < NPIX; i += blockDim.x) acc_eps += (double)gP_ext[i];
            double v[3] = {(double)acc_ll, acc_eps, (double)acc_sig};
            block_sum<3>(v, ms->red);
            if (threadIdx.x == 0) {
                // the all-background image: lp_bg / dlp_bg at every pixel
                const double n1 = (double)ms->n_ones, n0 = (double)(NPIX - ms->n_ones);
                if (score && !gP_ext) {
                    v[0] += n1 * (double)lp_bg1 + n0 * (double)lp_bg0;
                    if (eps > 0.0f) v[1] += (double)gscale * (n1 * dlp_bgd1 + n0 * dlp_bgd0);
                }
                if (a.ll) a.ll[p] = (score && !gP_ext) ? (float)v[0] : 0.0f;
                if (a.off_page) a.off_page[p] = (uint8_t)(ms->off_page != 0);
                if (a.g_eps) a.g_eps[p] = (float)v[1];
                if (a.g_sigma) {
                    const double s3 = (double)sigma * (double)sigma * (double)sigma;
                    a.g_sigma[p] = sigma > 0.0f ? (float)(v[2] / s3) : 0.0f;
                }
            }
        }

        BPL_PT(5);      // scalar reductions
        // ---- add_stroke backward: gather at the splat points ----
        float *accs = B;       // [S][NACC] sums per sub-stroke (B is free; re-zeroed below)
        if (RAW) {
            for (int s = warp; s < t.S; s += nw) {
                const int o = a.sub_pt_off[t.s0 + s], n = a.sub_pt_off[t.s0 + s + 1] - o;
                const RawSrc src{reinterpret_cast<const float2 *>(a.pts) + o};
                const Stats st = substroke_stats(src, n, lane, a.rc);
                float *gp = a.g_pts + 2 * (size_t)o;
                substroke_backward(src, n, lane, a.rc, st, A, [&](int j, float gx, float gy) {
                    atomicAdd(gp + 2 * j, gx);
                    atomicAdd(gp + 2 * j + 1, gy);
                });
            }
            __syncthreads();
        } else {
            const int QN = (a.neval + BWD_PPL - 1) / BWD_PPL, items = t.S * QN;
            // zero the per-sub-stroke accumulators; does any sub-stroke sit in the rescale branch?
            for (int i = threadIdx.x; i < t.S * NACC; i += blockDim.x) accs[ORIGIN + i] = 0.0f;
            for (int i = threadIdx.x; i < t.S; i += blockDim.x) ms->dot[i] = 0.0f;
            bool need_dot = false;
            for (int s = 0; s < t.S; ++s) {
                const float4 v = ms->sub[s];
                need_dot |= (__float_as_int(v.w) >= 2 && v.z >= 2.22e-6f && v.z < a.rc.ink_pp);
            }
            __syncthreads();
            auto noemit = [](int, float, float) {};
            if (need_dot) {
                for (int base = 0; base < items; base += blockDim.x) {
                    const int item = base + threadIdx.x;
                    bool active = item < items;
                    const int s = active ? item / QN : 0x7fffffff;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (active) v = ms->sub[s];
                    active = active && (__float_as_int(v.w) >= 2 && v.z >= 2.22e-6f && v.z < a.rc.ink_pp);
                    CtrlSrc src{};
                    if (active) src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
                    const float d = warp_points_bwd<true, STRIDE, ORIGIN, BWD_PPL>(active, src, sm.tab, s, a.neval,
                                                                         item < items ? item - s * QN : 0, lane, a.rc, A, v.z,
                                                                         __float_as_int(v.w), 0.0f, noemit);
                    unsigned todo = __ballot_sync(0xffffffffu, active);
                    while (todo) {
                        const int leader = __ffs(todo) - 1;
                        const int cur = __shfl_sync(0xffffffffu, s, leader);
                        const bool mine = active && s == cur;
                        const float tot = warp_sum(mine ? d : 0.0f);
                        if (lane == leader) atomicAdd(&ms->dot[cur], tot);
                        todo &= ~__ballot_sync(0xffffffffu, mine);
                    }
                }
                __syncthreads();
            }
            for (int base = 0; base < items; base += blockDim.x) {
                const int item = base + threadIdx.x;
                bool active = item < items;
                const int s = active ? item / QN : 0x7fffffff;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) v = ms->sub[s];
                active = active && __float_as_int(v.w) >= 1;
                CtrlSrc src{};
                if (active) src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
                float wx[NCPT], wy[NCPT], sx = 0.0f, sy = 0.0f, ax = 0.0f, ay = 0.0f;
#pragma unroll
                for (int k = 0; k < NCPT; ++k) { wx[k] = 0.0f; wy[k] = 0.0f; }
                warp_points_bwd<false, STRIDE, ORIGIN, BWD_PPL>(active, src, sm.tab, s, a.neval, item < items ? item - s * QN : 0, lane,
                                                       a.rc, A, v.z, __float_as_int(v.w), active ? ms->dot[s] : 0.0f,
                                                       [&](int j, float gx, float gy) {
                    const float *row = sm.tab + j * NCPT;
#pragma unroll
                    for (int k = 0; k < NCPT; ++k) { wx[k] = fmaf(row[k], gx, wx[k]); wy[k] = fmaf(row[k], gy, wy[k]); }
                    sx += gx; sy += gy;
                    if (aff) {
                        float px, py;
                        src.pre(j, px, py);
                        ax = fmaf(gx, px, ax); ay = fmaf(gy, py, ay);
                    }
                });
                // combine the lanes of a warp that share a sub-stroke: the 12 (+ 2 affine) sums are reduced together
                // (warp_reduce16: lanes 2i, 2i+1 end up with sum i), then the even lanes add their sum to the record
                unsigned todo = __ballot_sync(0xffffffffu, active);
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    const int cur = __shfl_sync(0xffffffffu, s, leader);
                    const bool mine = active && s == cur;
                    float v[16];
#pragma unroll
                    for (int k = 0; k < NCPT; ++k) { v[2 * k] = mine ? wx[k] : 0.0f; v[2 * k + 1] = mine ? wy[k] : 0.0f; }
                    v[2 * NCPT] = mine ? sx : 0.0f; v[2 * NCPT + 1] = mine ? sy : 0.0f;
                    v[12] = mine ? ax : 0.0f; v[13] = mine ? ay : 0.0f; v[14] = 0.0f; v[15] = 0.0f;
                    const float tot = warp_reduce16(v, lane);
                    const int idx = lane >> 1;
                    if (!(lane & 1)) {
                        if (idx < NACC) atomicAdd(accs + ORIGIN + cur * NACC + idx, tot);
                        if (aff) {
                            if (idx == 12) atomicAdd(&ms->aff_acc[0], (double)tot);
                            else if (idx == 13) atomicAdd(&ms->aff_acc[1], (double)tot);
                            else if (idx == 2 * NCPT) atomicAdd(&ms->aff_acc[2], (double)tot);
                            else if (idx == 2 * NCPT + 1) atomicAdd(&ms->aff_acc[3], (double)tot);
                        }
                    }
                    todo &= ~__ballot_sync(0xffffffffu, mine);
                }
            }
            __syncthreads();
        }

        BPL_PT(6);      // add_stroke backward (incl. rescale dot pre-pass)
        // ---- chain backward (part.py:316-326) and control-point gradients ----
        if (!RAW) {
            float gcx = 0.0f, gcy = 0.0f, b0 = 1.0f, b1 = 1.0f;
            if (aff) {      // p' = p * B[:2] + B[2:], B[2:] = A[2:] - (A[:2]-1) com, com = mean(p)
                const float *Aa = a.affine + 4 * (size_t)p;
                const double npts = (double)t.S * (double)a.neval;
                b0 = ms->affB[0]; b1 = ms->affB[1];
                gcx = npts > 0.0 ? (float)(-((double)Aa[0] - 1.0) * ms->aff_acc[2] / npts) : 0.0f;
                gcy = npts > 0.0 ? (float)(-((double)Aa[1] - 1.0) * ms->aff_acc[3] / npts) : 0.0f;
                if (threadIdx.x == 0 && a.g_affine) {
                    float *ga = a.g_affine + 4 * (size_t)p;
                    ga[0] = (float)(ms->aff_acc[0] - (double)ms->com[0] * ms->aff_acc[2]);
                    ga[1] = (float)(ms->aff_acc[1] - (double)ms->com[1] * ms->aff_acc[3]);
                    ga[2] = (float)ms->aff_acc[2];
                    ga[3] = (float)ms->aff_acc[3];
                }
            } else if (threadIdx.x == 0 && a.g_affine) {
                float *ga = a.g_affine + 4 * (size_t)p;
                ga[0] = ga[1] = ga[2] = ga[3] = 0.0f;
            }
            const float *rfirst = sm.tab, *rlast = sm.tab + (a.neval - 1) * NCPT;
            // (i) one thread per stroke: suffix sums of the sub-strokes' summed gradients, on shared memory only;
            //     the carry entering each sub-stroke and its total overwrite the two Gsum slots
            for (int k = threadIdx.x; k < t.K; k += blockDim.x) {
                const int b = a.stroke_sub_off[t.k0 + k] - t.s0, e = a.stroke_sub_off[t.k0 + k + 1] - t.s0;
                float cx = 0.0f, cy = 0.0f;       // gradient w.r.t. previous_pos of the following sub-stroke
                for (int s = e - 1; s >= b; --s) {
                    float *o = accs + ORIGIN + s * NACC;
                    const float sx = fmaf(b0, o[2 * NCPT], (float)a.neval * gcx);
                    const float sy = fmaf(b1, o[2 * NCPT + 1], (float)a.neval * gcy);
                    ms->sub[s].z = cx; ms->sub[s].w = cy;            // carry (the forward's totals are no longer needed)
                    cx += sx; cy += sy;
                    o[2 * NCPT] = cx; o[2 * NCPT + 1] = cy;          // total
                }
                a.g_position[2 * (t.k0 + k)] = cx;
                a.g_position[2 * (t.k0 + k) + 1] = cy;
            }
            __syncthreads();
            // (ii) one thread per sub-stroke: control-point and scale gradients (global loads in parallel)
            for (int s = threadIdx.x; s < t.S; s += blockDim.x) {
                const float *o = accs + ORIGIN + s * NACC;
                const float cx = ms->sub[s].z, cy = ms->sub[s].w;
                const float tx = o[2 * NCPT], ty = o[2 * NCPT + 1];
                const float inv = a.invscale[t.s0 + s];
                const float *cp = a.ctrl + (size_t)(t.s0 + s) * NCPT * 2;
                float *gc = a.g_ctrl + (size_t)(t.s0 + s) * NCPT * 2;
                float gi = 0.0f;
#pragma unroll
                for (int q = 0; q < NCPT; ++q) {
                    const float wx = fmaf(b0, o[2 * q], gcx * ms->colsum[q]);
                    const float wy = fmaf(b1, o[2 * q + 1], gcy * ms->colsum[q]);
                    const float gyx = wx + rlast[q] * cx - rfirst[q] * tx;
                    const float gyy = wy + rlast[q] * cy - rfirst[q] * ty;
                    gc[2 * q] = inv * gyx;
                    gc[2 * q + 1] = inv * gyy;
                    gi = fmaf(cp[2 * q], gyx, fmaf(cp[2 * q + 1], gyy, gi));
                }
                a.g_invscale[t.s0 + s] = gi;
            }
            __syncthreads();
            // B's interior was used as scratch: restore zeros (its border was never touched)
            for (int i = threadIdx.x; i < t.S * NACC; i += blockDim.x) accs[ORIGIN + i] = 0.0f;
        }
        // leave all four buffers zero: every pass wrote inside the box only
        if (!bx.empty()) {
            for (int r = bx.r0 + warp; r <= bx.r1; r += nw) {      // a warp per row: no index divisions
                const int q0 = ORIGIN + r * STRIDE;
                for (int c = bx.c0 + lane; c <= bx.c1; c += 32) { A[q0 + c] = 0.0f; B[q0 + c] = 0.0f; C[q0 + c] = 0.0f; D[q0 + c] = 0.0f; }
            }
        }
        if (threadIdx.x == 0) { ms->next_item = p_next; ms->ones_acc = 0; }
        __syncthreads();
        p = ms->next_item;      // rewritten only at the end of the next iteration, many barriers from here
        BPL_PT(7);      // chain backward + outputs
    }
    if (threadIdx.x == 0) sched_finish(a);
#ifdef BPL_PHASE_TIMING
    if (threadIdx.x == 0 && a.pt_buf)
        for (int i = 0; i < 8; ++i) a.pt_buf[blockIdx.x * 16 + i] = (float)pt_acc[i];
#endif
}


}  // namespace bpl

#include "bpl_backward_teams.cuh"
#include "bpl_anysize.cuh"

namespace bpl {

// ------------------------------------------------------------------------------------------
// Processing order for small batches: tokens by decreasing sub-stroke count (a counting sort in one CTA), so
// that the work queue hands out the expensive tokens first and the launch ends on cheap ones.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) order_tokens_kernel(const int *__restrict__ tok_stroke_off,
                                                            const int *__restrict__ stroke_sub_off, int n, int *__restrict__ order) {
    __shared__ int hist[MAXS + 1], start[MAXS + 1];
    for (int i = threadIdx.x; i <= MAXS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    auto key = [&](int p) {
        const int S = stroke_sub_off[tok_stroke_off[p + 1]] - stroke_sub_off[tok_stroke_off[p]];
        return MAXS - min(max(S, 0), MAXS);          // 0 = most expensive
    };
    constexpr int KEEP = 16;                         // keys of a thread's first 16 tokens stay in registers
    int keys[KEEP];
#pragma unroll
    for (int i = 0; i < KEEP; ++i) {
        const int p = threadIdx.x + i * (int)blockDim.x;
        keys[i] = p < n ? key(p) : 0;
        if (p < n) atomicAdd(&hist[keys[i]], 1);
    }
    for (int p = threadIdx.x + KEEP * blockDim.x; p < n; p += blockDim.x) atomicAdd(&hist[key(p)], 1);
    __syncthreads();
    if (threadIdx.x < 32) {                          // exclusive scan of the 129 bins by one warp
        int carry = 0;
        for (int base = 0; base <= MAXS; base += 32) {
            const int i = base + (int)threadIdx.x;
            const int v = i <= MAXS ? hist[i] : 0;
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += o; }
            if (i <= MAXS) start[i] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KEEP; ++i) {
        const int p = threadIdx.x + i * (int)blockDim.x;
        if (p < n) order[atomicAdd(&start[keys[i]], 1)] = p;
    }
    for (int p = threadIdx.x + KEEP * blockDim.x; p < n; p += blockDim.x) order[atomicAdd(&start[key(p)], 1)] = p;
}

// ------------------------------------------------------------------------------------------
// Processing order by ESTIMATED COST.  The sub-stroke count explains a third of a token's cost (corr. 0.58 on the bench
// workload); most of it is stencil work, proportional to the area of the ink's bounding box, which varies from 3 % to
// 100 % of the canvas independently of the sub-stroke count.  A spline lies in the convex hull of its control points
// (the rows of coefficient_mat are non-negative and sum to one, pybpl/splines.py:72-93), so the box of the chained
// control polygons bounds the ink box: one thread per token walks its sub-strokes (first / last trajectory point from
// rows 0 and neval-1 of the basis table, part.py:316-326) and turns (area, sub-strokes) into a key of 256 cost classes.
// ------------------------------------------------------------------------------------------
constexpr int COST_BINS = 256;

__global__ void __launch_bounds__(128) token_cost_kernel(const bpl_token_batch b, int reach, float w_area, float w_sub,
                                                         int *__restrict__ keys) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.n_tokens) return;
    const int k0 = b.tok_stroke_off[p], k1 = b.tok_stroke_off[p + 1];
    const float *r0 = b.basis, *r1 = b.basis + (size_t)(b.neval - 1) * NCPT;
    float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
    int S = 0;
    for (int k = k0; k < k1; ++k) {
        float px = b.position[2 * k], py = b.position[2 * k + 1];
        for (int s = b.stroke_sub_off[k]; s < b.stroke_sub_off[k + 1]; ++s, ++S) {
            const float inv = b.invscale[s];
            const float *cp = b.ctrl + (size_t)s * NCPT * 2;
            float yx[NCPT], yy[NCPT], ax = 0.0f, ay = 0.0f, ex = 0.0f, ey = 0.0f;
#pragma unroll
            for (int q = 0; q < NCPT; ++q) {
                yx[q] = inv * cp[2 * q]; yy[q] = inv * cp[2 * q + 1];
                ax = fmaf(r0[q], yx[q], ax); ay = fmaf(r0[q], yy[q], ay);
                ex = fmaf(r1[q], yx[q], ex); ey = fmaf(r1[q], yy[q], ey);
            }
            const float ox = ax - px, oy = ay - py;
#pragma unroll
            for (int q = 0; q < NCPT; ++q) {
                xmin = fminf(xmin, yx[q] - ox); xmax = fmaxf(xmax, yx[q] - ox);
                ymin = fminf(ymin, yy[q] - oy); ymax = fmaxf(ymax, yy[q] - oy);
            }
            px = ex - ox; py = ey - oy;
        }
    }
    float area = 0.0f;
    if (S > 0 && xmin <= xmax) {
        if (b.affine) {      // scale about the box centre (the centre of mass is near it), then shift (affine.py:48-53)
            const float *A = b.affine + 4 * (size_t)p;
            const float cx = 0.5f * (xmin + xmax), cy = 0.5f * (ymin + ymax);
            const float hx = 0.5f * (xmax - xmin) * fabsf(A[0]), hy = 0.5f * (ymax - ymin) * fabsf(A[1]);
            xmin = cx + A[2] - hx; xmax = cx + A[2] + hx; ymin = cy + A[3] - hy; ymax = cy + A[3] + hy;
        }
        // image rows = -y, columns = x (rendering.py:52-53); grown by the stencils' reach, clipped to the canvas
        const float c0 = fmaxf(0.0f, xmin - reach), c1 = fminf((float)(IMG - 1), xmax + reach);
        const float q0 = fmaxf(0.0f, -ymax - reach), q1 = fminf((float)(IMG - 1), -ymin + reach);
        area = fmaxf(0.0f, c1 - c0 + 1.0f) * fmaxf(0.0f, q1 - q0 + 1.0f);
    }
    // cost above the fixed part, in units of an average token's (area 4060 px, 5.2 sub-strokes); 256 classes over 0 .. 2
    const float cost = w_area * area * (1.0f / 4060.0f) + w_sub * (float)S * (1.0f / 5.2f);
    keys[p] = COST_BINS - 1 - min(COST_BINS - 1, max(0, (int)(cost * 128.0f)));                         // 0 = most expensive
}

__global__ void __launch_bounds__(256) cycles_key_kernel(const unsigned *__restrict__ cycles, int n, int *__restrict__ keys) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) keys[p] = COST_BINS - 1 - min(COST_BINS - 1, (int)(cycles[p] >> 11));      // 0 = most expensive
}

// counting sort of the tokens by key (one CTA; keys < COST_BINS)
__global__ void __launch_bounds__(1024) order_by_key_kernel(const int *__restrict__ keys, int n, int *__restrict__ order) {
    __shared__ int hist[COST_BINS], start[COST_BINS];
    for (int i = threadIdx.x; i < COST_BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) atomicAdd(&hist[keys[p]], 1);
    __syncthreads();
    if (threadIdx.x < 32) {                          // exclusive scan of the bins by one warp
        int carry = 0;
        for (int base = 0; base < COST_BINS; base += 32) {
            const int i = base + (int)threadIdx.x;
            const int v = hist[i];
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += o; }
            start[i] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) order[atomicAdd(&start[keys[p]], 1)] = p;
}

// ------------------------------------------------------------------------------------------
// Inspection kernel: the integer work of the path, per point, through the very same device
// functions the fused kernels use (chain_offsets / affine_setup / CtrlSrc / walk_round):
// motor-space points, out-of-bounds mask (rendering.py:27-31), floor/ceil pixel indices
// (rendering.py:113-116) and surviving-point count per sub-stroke.  Used by the parity tests and by
// StrokeToken.motor-style consumers that want all trajectories of a batch.
// ------------------------------------------------------------------------------------------
struct PtsSmem {
    float tab[MAX_TABLE];
    Misc ms;
};

__global__ void __launch_bounds__(256) bpl_points_kernel(const KArgs a, float *__restrict__ motor,
                                                          uint8_t *__restrict__ mask_out, int *__restrict__ fl,
                                                          int *__restrict__ ce, int *__restrict__ n_in) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PtsSmem &sm = *reinterpret_cast<PtsSmem *>(smem_raw);
    Misc *ms = &sm.ms;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < a.neval * NCPT; i += blockDim.x) sm.tab[i] = a.basis[i];
    __syncthreads();
    for (int p = blockIdx.x; p < a.n_tokens; p += gridDim.x) {
        const TokGeom t = token_geom<false>(a, p);
        if (t.S > MAXS || t.K > MAXK) continue;
        const bool aff = a.affine != nullptr;
        chain_offsets(a, t, sm.tab, ms);
        if (aff) affine_setup(a, t, p, sm.tab, ms);
        for (int s = warp; s < t.S; s += nw) {
            const CtrlSrc src = make_ctrl_src(a, t, s, sm.tab, ms, aff);
            Walk w{0, 0.0f, 0.0f, 0};
            const size_t o = (size_t)(t.s0 + s) * a.neval;
            for (int base = 0; base < a.neval; base += 32) {
                const Pt q = walk_round(src, a.neval, base, lane, a.rc, w);
                if (!q.valid) continue;
                if (motor) { motor[2 * (o + q.j)] = q.c; motor[2 * (o + q.j) + 1] = -q.r; }
                if (mask_out) mask_out[o + q.j] = (uint8_t)(!q.inb);
                if (fl) { fl[2 * (o + q.j)] = (int)floorf(q.r); fl[2 * (o + q.j) + 1] = (int)floorf(q.c); }
                if (ce) { ce[2 * (o + q.j)] = (int)ceilf(q.r); ce[2 * (o + q.j) + 1] = (int)ceilf(q.c); }
            }
            if (n_in && lane == 0) n_in[t.s0 + s] = w.cnt;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, const char *detail) { snprintf(g_err, sizeof(g_err), fmt, detail); }
const char *last_error() { return g_err; }
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches.load(); }

static int make_consts(const bpl_params *ps, RenderConsts *rc, bool any_size = false) {
    if (!ps) { set_error("%s", "params is NULL"); return BPL_EINVAL; }
    if ((!any_size && ps->imsize != BPL_IMSIZE) || ps->fsize < 1 || ps->fsize > BPL_FSIZE || !(ps->fsize & 1)) {
        set_error("%s", "imsize must be 105 (other sizes: bpl_render_strokes_any) and fsize an odd number <= 11 (pybpl/parameters.py:21,41)");
        return BPL_EINVAL;
    }
    if (ps->ink_ncon < 0 || ps->ink_ncon > 16) { set_error("%s", "ink_ncon out of range"); return BPL_EINVAL; }
    const double a = ps->ink_a, b = ps->ink_b;
    const double hs = ps->hinton ? b * (1.0 + a) : b;      // rendering.py:153-156
    rc->ink_ncon = ps->ink_ncon;
    rc->ink_pp = ps->ink_pp;
    rc->ink_max_dist = ps->ink_max_dist;
    rc->ink_f = ps->ink_pp / ps->ink_max_dist;
    {
        int ex = 0;
        const float mant = frexpf(ps->ink_max_dist, &ex);
        rc->inv_max_dist = (mant == 0.5f && ex > -100 && ex < 100) ? 1.0f / ps->ink_max_dist : 0.0f;
    }
    rc->fhalf = ps->fsize / 2;
    rc->rmax = rc->cmax = (float)(IMG - 1);
    rc->c1 = (float)(hs * a / 12.0);
    rc->c2 = (float)(hs * (1.0 - 4.0 * a / 3.0));
    return 0;
}

struct DevInfo { int sm_count; int smem_optin; bool ok; };
static DevInfo dev_info() {
    static thread_local int cached_dev = -1;
    static thread_local DevInfo info{0, 0, false};
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return DevInfo{0, 0, false};
    if (dev != cached_dev) {
        info.ok = cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&info.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess;
        cached_dev = dev;
    }
    return info;
}

int sm_count() {      // SMs of the current device (148 on B200); 0 when there is no device
    const DevInfo di = dev_info();
    return di.ok ? di.sm_count : 0;
}

template <class K>
static int launch(K kernel, int threads, size_t smem, int ctas_per_sm, const KArgs &ka, cudaStream_t st,
                  int work_items = -1, bool need_queue = false) {
    const DevInfo di = dev_info();
    if (!di.ok) { set_error("%s", "no CUDA device"); return BPL_ENODEV; }
    if ((size_t)di.smem_optin < smem) { set_error("%s", "device lacks the shared memory this kernel needs (built for sm_100a)"); return BPL_ENODEV; }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int grid = di.sm_count * ctas_per_sm;
    if (work_items < 0) work_items = ka.n_tokens;
    if (grid > work_items) grid = work_items;
    if (grid <= 0) return 0;
    KArgs kb = ka;
    static const bool static_sched = getenv("BPL_STATIC_SCHED") != nullptr;
    if (!static_sched || need_queue) {
        static std::atomic<unsigned> slot{0};
        static thread_local int ring_dev = -1;
        static thread_local Sched *ring = nullptr;      // g_sched has one instance per device
        int dev = -1;
        cudaGetDevice(&dev);
        if (dev != ring_dev) {
            cudaError_t es = cudaGetSymbolAddress(reinterpret_cast<void **>(&ring), g_sched);
            if (es != cudaSuccess) { set_error("cudaGetSymbolAddress: %s", cudaGetErrorString(es)); return (int)es; }
            ring_dev = dev;
        }
        // A launch owns its slot from launch to completion (the last CTA resets it).  Live launches rotate through
        // the lower half of the ring -- 2 048 launches would have to be in flight at once to collide.  A launch that is
        // being CAPTURED into a CUDA graph keeps its slot for every replay, so it takes one from the upper half, which
        // live launches never touch; replays of one graph are serialised by CUDA when they are launched into one
        // stream (examples/fit_parses.py).  Launching the SAME captured graph into two streams at once, or capturing
        // more than 2 048 launches that later run concurrently, is not supported.
        static std::atomic<unsigned> slot_captured{0};
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cap = cudaStreamCaptureStatusNone; (void)cudaGetLastError(); }
        kb.sched = (cap == cudaStreamCaptureStatusActive)
                       ? ring + SCHED_RING / 2 + (slot_captured.fetch_add(1u) % (SCHED_RING / 2))
                       : ring + (slot.fetch_add(1u) % (SCHED_RING / 2));
    }
#if defined(BPL_PHASE_TIMING) || defined(BPL_TOKEN_TIMING)
    if (const char *pb = getenv("BPL_PT_BUF")) kb.pt_buf = reinterpret_cast<float *>(strtoull(pb, nullptr, 10));
#endif
    kernel<<<grid, threads, smem, st>>>(kb);
    count_launch(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

static int fill_common(KArgs &ka, const bpl_params *ps, const bpl_targets *t, const bpl_fwd_out *out,
                       const bpl_grads *g) {
    int rcode = make_consts(ps, &ka.rc);
    if (rcode) return rcode;
    if (t) { ka.image_bits = t->image_bits; ka.img_idx = t->img_idx; ka.n_images = t->n_images; ka.all_pairs = t->all_pairs; }
    if (out) { ka.ll = out->ll; ka.off_page = out->off_page; ka.pimg = out->pimg; }
    if (g) {
        ka.g_ll = g->g_ll; ka.g_pimg = g->g_pimg; ka.g_ctrl = g->g_ctrl; ka.g_invscale = g->g_invscale;
        ka.g_position = g->g_position; ka.g_affine = g->g_affine; ka.g_pts = g->g_pts; ka.g_eps = g->g_eps;
        ka.g_sigma = g->g_sigma;
    }
    return 0;
}

static int fill_tokens(KArgs &ka, const bpl_token_batch *b) {
    if (!b) { set_error("%s", "batch is NULL"); return BPL_EINVAL; }
    if (b->ncpt != NCPT) { set_error("%s", "fused path is built for ncpt=5 (use bpl_vanilla_to_motor + bpl_render_strokes otherwise)"); return BPL_ELIMIT; }
    if (b->neval < 1 || b->neval * NCPT > MAX_TABLE) { set_error("%s", "neval*ncpt exceeds the shared-memory basis table (use bpl_vanilla_to_motor + bpl_render_strokes)"); return BPL_ELIMIT; }
    if (!b->tok_stroke_off || !b->stroke_sub_off || !b->eps || !b->sigma || !b->basis) {   // ctrl etc. may be NULL when no token has a stroke
        set_error("%s", "token batch has a NULL array"); return BPL_EINVAL;
    }
    ka.n_tokens = b->n_tokens; ka.neval = b->neval;
    ka.tok_stroke_off = b->tok_stroke_off; ka.stroke_sub_off = b->stroke_sub_off;
    ka.ctrl = b->ctrl; ka.invscale = b->invscale; ka.position = b->position; ka.affine = b->affine;
    ka.basis = b->basis; ka.eps = b->eps; ka.sigma = b->sigma;
    ka.order = b->order;
    ka.cycles_out = b->cycles_out;
    return 0;
}

static int fill_strokes(KArgs &ka, const bpl_stroke_batch *b) {
    if (!b) { set_error("%s", "batch is NULL"); return BPL_EINVAL; }
    if (!b->tok_sub_off || !b->sub_pt_off || !b->eps || !b->sigma) {
        set_error("%s", "stroke batch has a NULL array"); return BPL_EINVAL;
    }
    ka.n_tokens = b->n_tokens;
    ka.tok_sub_off = b->tok_sub_off; ka.sub_pt_off = b->sub_pt_off; ka.pts = b->pts;
    ka.eps = b->eps; ka.sigma = b->sigma;
    return 0;
}

}  // namespace bpl

using namespace bpl;

extern "C" {

int bpl_order_tokens(const bpl_token_batch *b, int *order, void *stream) {
    if (!b || !order || !b->tok_stroke_off || !b->stroke_sub_off) { set_error("%s", "bpl_order_tokens: NULL argument"); return BPL_EINVAL; }
    if (b->n_tokens <= 0) return 0;
    order_tokens_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(b->tok_stroke_off, b->stroke_sub_off, b->n_tokens, order);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// launch shape of the scoring kernel: see bpl_forward1s.cuh (BPL_FWD_SHAPE=small|large forces one, an A/B aid)
static bool large_shape(int n_tokens) {
    static const char *force = getenv("BPL_FWD_SHAPE");
    if (force) return force[0] == 'l';
    return n_tokens >= S_LARGE_MIN;
}

int bpl_order_tokens_by_cost(const bpl_params *ps, const bpl_token_batch *b, int *keys, int *order, void *stream) {
    if (!ps || !b || !order || !keys || !b->tok_stroke_off || !b->stroke_sub_off || !b->basis) {
        set_error("%s", "bpl_order_tokens_by_cost: NULL argument"); return BPL_EINVAL;
    }
    if (b->ncpt != NCPT || b->neval < 1) { set_error("%s", "bpl_order_tokens_by_cost: fused path is built for ncpt=5"); return BPL_ELIMIT; }
    if (b->n_tokens <= 0) return 0;
    // weights fitted to measured cycles per token (tools/token_cost_fit.py, 4 096 bench tokens): forward 0.39 + 0.36 area
    // + 0.19 sub-strokes, fused backward 0.53 + 0.30 area + 0.12 sub-strokes; one key serves both
    float w_area = 0.34f, w_sub = 0.17f;
    if (const char *wenv = getenv("BPL_COST_W")) sscanf(wenv, "%f,%f", &w_area, &w_sub);      // tuning aid
    token_cost_kernel<<<(b->n_tokens + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*b, ps->ink_ncon + 2 * PAD, w_area, w_sub, keys);
    order_by_key_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(keys, b->n_tokens, order);
    count_launch(2);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int bpl_order_tokens_by_cycles(const unsigned *cycles, int n, int *keys, int *order, void *stream) {
    if (!cycles || !keys || !order) { set_error("%s", "bpl_order_tokens_by_cycles: NULL argument"); return BPL_EINVAL; }
    if (n <= 0) return 0;
    cycles_key_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cycles, n, keys);
    order_by_key_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(keys, n, order);
    count_launch(2);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int bpl_score_tokens(const bpl_params *ps, const bpl_token_batch *b, const bpl_targets *t, const bpl_fwd_out *out,
                     void *stream) {
    if (b && b->n_tokens == 0) return 0;
    KArgs ka{};
    int r = fill_common(ka, ps, t, out, nullptr);
    if (r) return r;
    if ((r = fill_tokens(ka, b))) return r;
    if (ka.all_pairs && (!ka.image_bits || ka.pimg || ka.n_images < 1)) {
        set_error("%s", "all_pairs scoring needs target images and no pimg output"); return BPL_EINVAL;
    }
    // Scoring: the box-confined kernel (one token per CTA).  Probability maps (pimg), and BPL_FORCE_SINGLE=1 (A/B and
    // profiling aid, one target per token only), take the dense single-token kernel.
    static const bool force_single = getenv("BPL_FORCE_SINGLE") != nullptr;
    if (ka.image_bits && !ka.pimg && (ka.all_pairs || !force_single)) {
        if (ka.all_pairs) {      // (no processing order in this mode)
            ka.order = nullptr;
            if (large_shape(ka.n_tokens))
                return launch(bpl_forward1s_kernel<true, S_THREADS_LARGE, S_CTAS_LARGE>, S_THREADS_LARGE, sizeof(Fwd1sSmem), S_CTAS_LARGE, ka, (cudaStream_t)stream);
            return launch(bpl_forward1s_kernel<true, S_THREADS, S_CTAS_PER_SM>, S_THREADS, sizeof(Fwd1sSmem), S_CTAS_PER_SM, ka, (cudaStream_t)stream);
        }
        if (large_shape(ka.n_tokens))
            return launch(bpl_forward1s_kernel<false, S_THREADS_LARGE, S_CTAS_LARGE>, S_THREADS_LARGE, sizeof(Fwd1sSmem), S_CTAS_LARGE, ka, (cudaStream_t)stream);
        return launch(bpl_forward1s_kernel<false, S_THREADS, S_CTAS_PER_SM>, S_THREADS, sizeof(Fwd1sSmem), S_CTAS_PER_SM, ka, (cudaStream_t)stream);
    }
    return launch(bpl_forward_kernel<false>, FWD_THREADS, sizeof(FwdSmem), FWD_CTAS_PER_SM, ka, (cudaStream_t)stream);
}

int bpl_score_tokens_lse(const bpl_params *ps, const bpl_token_batch *b, const bpl_targets *t, const bpl_fwd_out *out,
                         float *lse, float *lse_scratch, int accumulate, void *stream) {
    if (!lse || !lse_scratch) { set_error("%s", "bpl_score_tokens_lse: lse and lse_scratch are required"); return BPL_EINVAL; }
    if (!b) { set_error("%s", "batch is NULL"); return BPL_EINVAL; }
    KArgs ka{};
    int r = fill_common(ka, ps, t, out, nullptr);
    if (r) return r;
    if ((r = fill_tokens(ka, b))) return r;
    if (!ka.image_bits || ka.pimg || ka.all_pairs) { set_error("%s", "bpl_score_tokens_lse scores one target per token (no pimg, no all_pairs)"); return BPL_EINVAL; }
    if (b->n_tokens == 0) {      // an empty shard: (-inf, 0) unless accumulating
        if (!accumulate) {
            const float init[2] = {-INFINITY, 0.0f};
            cudaError_t e = cudaMemcpyAsync(lse, init, sizeof(init), cudaMemcpyHostToDevice, (cudaStream_t)stream);
            if (e != cudaSuccess) { set_error("cudaMemcpyAsync: %s", cudaGetErrorString(e)); return (int)e; }
        }
        return 0;
    }
    ka.lse = lse; ka.lse_scratch = lse_scratch; ka.lse_accumulate = accumulate;
    if (large_shape(ka.n_tokens))
        return launch(bpl_forward1s_kernel<false, S_THREADS_LARGE, S_CTAS_LARGE>, S_THREADS_LARGE, sizeof(Fwd1sSmem), S_CTAS_LARGE, ka, (cudaStream_t)stream, -1, true);
    return launch(bpl_forward1s_kernel<false, S_THREADS, S_CTAS_PER_SM>, S_THREADS, sizeof(Fwd1sSmem), S_CTAS_PER_SM, ka, (cudaStream_t)stream, -1, true);
}

int bpl_render_strokes(const bpl_params *ps, const bpl_stroke_batch *b, const bpl_targets *t, const bpl_fwd_out *out,
                       void *stream) {
    if (b && b->n_tokens == 0) return 0;
    KArgs ka{};
    int r = fill_common(ka, ps, t, out, nullptr);
    if (r) return r;
    if ((r = fill_strokes(ka, b))) return r;
    return launch(bpl_forward_kernel<true>, FWD_THREADS, sizeof(FwdSmem), FWD_CTAS_PER_SM, ka, (cudaStream_t)stream);
}

#ifndef BPL_BWD_TEAMS_MIN
#define BPL_BWD_TEAMS_MIN 1024
#endif
static constexpr int BWD_TEAMS_MIN = BPL_BWD_TEAMS_MIN;

int bpl_score_tokens_grad(const bpl_params *ps, const bpl_token_batch *b, const bpl_targets *t,
                          const bpl_fwd_out *out, const bpl_grads *g, void *stream) {
    if (b && b->n_tokens == 0) return 0;
    KArgs ka{};
    int r = fill_common(ka, ps, t, out, g);
    if (r) return r;
    if ((r = fill_tokens(ka, b))) return r;
    if (!g || !g->g_ctrl || !g->g_invscale || !g->g_position) { set_error("%s", "g_ctrl/g_invscale/g_position required"); return BPL_EINVAL; }
    if (ka.all_pairs) { set_error("%s", "all_pairs scoring is forward only"); return BPL_EINVAL; }
    if (!ka.image_bits && !ka.g_pimg) { set_error("%s", "need target images or a pimg cotangent"); return BPL_EINVAL; }
    // Small batches: the full-canvas kernel, one 640-thread CTA per token and SM (shortest per-token latency, so the
    // shortest tail of the launch).  From BWD_TEAMS_MIN tokens on: teams of one persistent CTA per SM share the SM's
    // whole shared memory as a pool of box-sized arenas, expensive tokens first (bpl_backward_teams.cuh).  Measured
    // M particles/s, canvas / teams: 1 024 tokens 3.80 / 4.12, 2 048: 4.12 / 4.46, 8 192: 4.49 / 4.85, 32 768: 4.60 / 4.95.
    // BPL_BWD_CANVAS=1 / BPL_BWD_TEAMS=1 force one or the other (A/B aids); BPL_BWD_ORDER=1 lets the canvas kernel follow
    // the processing order too (measured 2 % slower).
    // (read per call, not cached: the parity tests switch kernels inside one process)
    const bool bwd_order = getenv("BPL_BWD_ORDER") != nullptr;
    const bool bwd_canvas = getenv("BPL_BWD_CANVAS") != nullptr;
    const bool bwd_teams = getenv("BPL_BWD_TEAMS") != nullptr;
    if (bwd_canvas || (!bwd_teams && ka.n_tokens < BWD_TEAMS_MIN)) {
        if (!bwd_order) ka.order = nullptr;
        return launch(bpl_backward_kernel<false>, BWD_THREADS, sizeof(BwdSmem), 1, ka, (cudaStream_t)stream);
    }
    return launch(bpl_backward_teams_kernel, BT_NT * BT_TT, (size_t)dev_info().smem_optin, 1, ka, (cudaStream_t)stream,
                  (ka.n_tokens + BT_NT - 1) / BT_NT);
}

int bpl_token_points(const bpl_params *ps, const bpl_token_batch *b, float *motor, uint8_t *mask_out, int *fl, int *ce,
                      int *n_in, void *stream) {
    if (b && b->n_tokens == 0) return 0;
    KArgs ka{};
    int r = fill_common(ka, ps, nullptr, nullptr, nullptr);
    if (r) return r;
    if ((r = fill_tokens(ka, b))) return r;
    if (ka.n_tokens <= 0) return 0;
    const int cap = sm_count() * 8;
    int grid = ka.n_tokens < cap ? ka.n_tokens : cap;
    bpl_points_kernel<<<grid, 256, sizeof(PtsSmem), (cudaStream_t)stream>>>(ka, motor, mask_out, fl, ce, n_in);
    count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int bpl_render_strokes_grad(const bpl_params *ps, const bpl_stroke_batch *b, const bpl_targets *t,
                            const bpl_fwd_out *out, const bpl_grads *g, void *stream) {
    if (b && b->n_tokens == 0) return 0;
    KArgs ka{};
    int r = fill_common(ka, ps, t, out, g);
    if (r) return r;
    if ((r = fill_strokes(ka, b))) return r;
    if (!g || !g->g_pts) { set_error("%s", "g_pts required"); return BPL_EINVAL; }
    if (!ka.image_bits && !ka.g_pimg) { set_error("%s", "need target images or a pimg cotangent"); return BPL_EINVAL; }
    return launch(bpl_backward_kernel<true>, BWD_THREADS, sizeof(BwdSmem), 1, ka, (cudaStream_t)stream);
}


// ---- any image size -----------------------------------------------------------------------------------
static int any_grid(int n_tokens) {
    const DevInfo di = dev_info();
    if (!di.ok) return 0;
    const int g = di.sm_count * 4;
    return n_tokens < g ? n_tokens : g;
}

long long bpl_render_any_workspace(int H, int W, int n_tokens, int with_grad) {
    if (H < 1 || W < 1 || n_tokens < 0) return -1;
    const int grid = any_grid(n_tokens);
    return (long long)grid * (with_grad ? ANY_BWD_BUFS : ANY_FWD_BUFS) * H * W;
}

static int any_launch(const bpl_params *ps, const bpl_any_geom *geo, const bpl_stroke_batch *b, const bpl_fwd_out *out,
                      const bpl_grads *g, bool grad, cudaStream_t st) {
    if (b && b->n_tokens == 0) return 0;
    if (!geo) { set_error("%s", "geometry is NULL"); return BPL_EINVAL; }
    if (geo->H < 1 || geo->W < 1 || (long long)geo->H * geo->W > (1ll << 24)) {
        set_error("%s", "imsize out of range (1 <= H, W and H * W <= 2^24)"); return BPL_EINVAL;
    }
    KArgs ka{};
    int r = make_consts(ps, &ka.rc, true);
    if (r) return r;
    ka.rc.rmax = (float)(geo->H - 1);
    ka.rc.cmax = (float)(geo->W - 1);
    if ((r = fill_strokes(ka, b))) return r;
    ka.img_idx = geo->img_idx;
    if (out) { ka.ll = out->ll; ka.off_page = out->off_page; ka.pimg = out->pimg; }
    if (grad) {
        if (!g || !g->g_pts) { set_error("%s", "g_pts required"); return BPL_EINVAL; }
        if (!geo->images && !g->g_pimg) { set_error("%s", "need target images or a pimg cotangent"); return BPL_EINVAL; }
        ka.g_ll = g->g_ll; ka.g_pimg = g->g_pimg; ka.g_pts = g->g_pts; ka.g_eps = g->g_eps; ka.g_sigma = g->g_sigma;
    }
    const int grid = any_grid(ka.n_tokens);
    if (grid <= 0) { set_error("%s", "no CUDA device"); return BPL_ENODEV; }
    const long long need = bpl_render_any_workspace(geo->H, geo->W, ka.n_tokens, grad);
    if (!geo->workspace || geo->workspace_floats < need) {
        set_error("%s", "workspace too small (bpl_render_any_workspace)"); return BPL_EINVAL;
    }
    AnyGeom gm{geo->H, geo->W, geo->workspace, geo->images};
    if (grad) bpl_render_any_grad_kernel<<<grid, ANY_THREADS, 0, st>>>(ka, gm);
    else bpl_render_any_kernel<<<grid, ANY_THREADS, 0, st>>>(ka, gm);
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int bpl_render_strokes_any(const bpl_params *ps, const bpl_any_geom *geo, const bpl_stroke_batch *b, const bpl_fwd_out *out,
                           void *stream) {
    return any_launch(ps, geo, b, out, nullptr, false, (cudaStream_t)stream);
}

int bpl_render_strokes_any_grad(const bpl_params *ps, const bpl_any_geom *geo, const bpl_stroke_batch *b,
                                const bpl_fwd_out *out, const bpl_grads *g, void *stream) {
    return any_launch(ps, geo, b, out, g, true, (cudaStream_t)stream);
}

}  // extern "C"
