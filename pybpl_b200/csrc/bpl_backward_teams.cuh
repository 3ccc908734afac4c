// bpl_backward_teams.cuh -- fused forward + backward for control-point tokens, box-sized buffers.
//
// One persistent CTA per SM owns the SM's whole shared memory and runs BT_NT independent TEAMS of BT_TT threads.
// A team renders and differentiates one token at a time (pybpl/rendering.py:58-229 forward, its adjoint backward,
// model/image_dist.py:60-80 for the score) inside an ARENA cut out of the CTA's shared-memory pool: three image
// buffers of the token's own box -- the pixels its ink touches grown by the reach of all stencil passes
// (ink_ncon + 2 * 5), 37 % of the canvas on average, anything from 3 % to 100 % -- instead of four full 115x115
// canvases.  Teams synchronise with named barriers (bar.sync id, BT_TT), never with __syncthreads(), so they drift
// apart and one team's stencil passes fill the issue slots another leaves idle at its barriers: the full-canvas
// kernel it replaces ran ONE 640-thread CTA per SM (211 KB of buffers), 38 % issue-active with 43 % of its stall
// samples at barriers.
//
// Arena allocation: first fit over the (at most BT_NT) live arenas by the team's leader under a shared-memory spin
// lock; a team that finds no room raises `waiting`, which keeps the other teams from starting new tokens until it
// has been served (they finish and free what they hold, so it always is: the largest arena, 3 x 105 x 105 floats, is
// 62 % of the pool).
//
// Box geometry: H x W pixels, both multiples of BT_GU = 5 (the run granularity of the sweeps), origin shifted inwards
// so that the box stays on the canvas; row stride W | 1 (odd: lanes walking down a column or along a row are both
// bank-conflict free).  Forward values are exactly zero outside the box; the adjoint is not, but the gather stage
// only reads it at the splat's own pixels, which are exact when the cotangent is known on the grown box (see
// bpl_render.cu, bpl_backward_kernel).  Zero padding of F.conv2d (util/general.py:161-163) = zeros outside the box.
//
// Sweeps are in-place capable: an item (a run of whole 5-blocks of one line) loads the halo it needs from the
// neighbouring runs into registers, the team synchronises, and the item then slides a register window along its
// run.  With three buffers A (broadened image, kept for the clamp mask), B, C the blur chain and its adjoint are
//   P1 B = g_v min(A,1)      P2 C = g_h B = I4         P3 B = U = g_v C, C = W = h~_v C        (in place on C)
//   P4 B = G5 = tail(g_h U); dsigma += G5 . (h~_h U)   (in place)      P5 B = t = g_h G5; dsigma += t . W  (in place)
//   P6 C = G4 = g_v t        P7 B = t1 = g_h G4, C = t1' = h~_h G4     (in place on C)
//   P9 dsigma += min(A,1) . (g_v t1')                  P8 dsigma += min(A,1) . (h~_v t1); A = [A <= 1] g_v t1
//   (P9 and P8 run as one sweep with two windows, bt_pass_mask2)
// (sigma^3 dK/dsigma = h~ (x) g + g (x) h~ with h~[u] = (u^2 - mu2) g[u]; all filters symmetric, so every adjoint pass
// is the same zero-padded convolution.)
#pragma once

namespace bpl {

#ifndef BPL_BT_TEAMS
#define BPL_BT_TEAMS 2
#endif
#ifndef BPL_BT_THREADS
#define BPL_BT_THREADS 320
#endif
#ifndef BPL_BT_PPL
#define BPL_BT_PPL 4
#endif
constexpr int BT_NT = BPL_BT_TEAMS;        // teams per CTA
constexpr int BT_TT = BPL_BT_THREADS;      // threads per team
constexpr int BT_PPL = BPL_BT_PPL;         // points per lane in the point stages
#ifdef BPL_UNROLL_PHASE_A_BWD
constexpr bool BT_ROLL_SPLAT = false;       // A/B: phase A of the point stages unrolled
#else
constexpr bool BT_ROLL_SPLAT = true;        // phase A of the splat as a rolled loop (as in the scoring kernel; see bpl_points.cuh)
#endif
#ifndef BPL_BT_GU
#define BPL_BT_GU 5
#endif
constexpr int BT_GU = BPL_BT_GU;           // outputs per unrolled block = granularity of the box (5 or 7: 105 = 21 * 5 = 15 * 7)
static_assert(BT_TT % 32 == 0 && BT_TT >= IMG, "a team needs one thread per line of the widest box");
static_assert(BT_NT >= 1 && BT_NT <= 8 && BT_NT * BT_TT <= 1024, "team shape");
static_assert(IMG % BT_GU == 0 && BT_GU >= PAD, "block structure");

struct TeamCtx {
    int id, tid, lane, warp;      // team, thread in team, lane, warp in team
    static constexpr int NW = BT_TT / 32;
    __device__ __forceinline__ void sync() const {      // named barrier of the team (ids 1..BT_NT; 0 is __syncthreads)
        __syncwarp();
        asm volatile("bar.sync %0, %1;" ::"r"(id + 1), "n"(BT_TT) : "memory");
    }
};

// the box of one token in its arena
struct LBox {
    int r0, c0, H, W, stride;     // canvas position of local (0,0); size; row stride
    __device__ __forceinline__ int area() const { return (H * stride + 3) & ~3; }
};

// Per-team record of the token in flight: written by the team's leader, read by everybody after a team barrier.  The
// phases below are real (non-inlined) functions that take everything from here, so nothing of a token has to stay in
// registers across the stencil passes and every phase gets the whole register budget for itself.
struct BtTok {
    int p;                        // the token
    TokGeom t;
    int bbox[4];                  // ink box on the canvas: rows [0]..[1], columns [2]..[3]; [1] < 0 when nothing is on the page
    LBox bx;
    int arena, need;              // first float / length of the team's arena in the pool
    int next[2];                  // the token after this one (double-buffered by the parity of the team's iteration)
    float eps, sigma, om, mixd, bgv, gscale;  // tail constants: 1 - eps, d mix / d I6, background probability, d L / d ll
};

struct BtShared {
    KArgs ka;                     // the kernel's arguments (non-inlined phases cannot take the parameter space by reference)
    float tab[MAX_TABLE];
    int lock, waiting, teams_done, pool_floats;
    int live[8][2];               // per team: (start, length) of its arena in floats; length 0 = none
    BtTok tok[BT_NT];
    Misc ms[BT_NT];
};

__device__ __forceinline__ TeamCtx team_ctx() {
    TeamCtx tm;
    tm.id = threadIdx.x / BT_TT;
    tm.tid = threadIdx.x - tm.id * BT_TT;
    tm.lane = threadIdx.x & 31;
    tm.warp = tm.tid >> 5;
    return tm;
}
__device__ __forceinline__ float *bt_pool(BtShared *sh) {
    constexpr int HDR = (int)((sizeof(BtShared) + 15) / 16 * 16);
    return reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sh) + HDR);
}

template <int NV>
__device__ __forceinline__ void team_sum(const TeamCtx &tm, double (&v)[NV], double *scratch) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    tm.sync();
    if (tm.lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + tm.warp] = v[i];
    }
    tm.sync();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = (tm.lane < TeamCtx::NW) ? scratch[i * 32 + tm.lane] : 0.0;
        v[i] = warp_sum(x);
    }
}

// leader only: first fit over the live arenas (see the header)
__device__ __noinline__ int arena_alloc(BtShared &sh, int me, int need, int pool) {
    volatile int *live = &sh.live[0][0];
    volatile int *waiting = &sh.waiting;
    for (;;) {
        while (atomicCAS(&sh.lock, 0, 1) != 0) __nanosleep(32);
        __threadfence_block();
        int got = -1;
        const int w = *waiting;
        if (w < 0 || w == me) {
            for (int ci = -1; ci < BT_NT; ++ci) {
                int cand = 0;
                if (ci >= 0) {
                    if (ci == me || live[2 * ci + 1] == 0) continue;
                    cand = live[2 * ci] + live[2 * ci + 1];
                }
                if (cand + need > pool) continue;
                bool ok = true;
                for (int j = 0; j < BT_NT; ++j) {
                    const int s = live[2 * j], l = live[2 * j + 1];
                    if (j != me && l > 0 && cand < s + l && s < cand + need) ok = false;
                }
                if (ok && (got < 0 || cand < got)) got = cand;
            }
            if (got >= 0) {
                live[2 * me] = got; live[2 * me + 1] = need;
                if (w == me) *waiting = -1;
            } else if (w < 0) {
                *waiting = me;
            }
        }
        __threadfence_block();
        atomicExch(&sh.lock, 0);
        if (got >= 0) return got;
        __nanosleep(256);
    }
}

// ---- team versions of the per-token set-up (bpl_render.cu: chain_offsets, affine_setup, token_consts, load_image_bits) ----
__device__ __forceinline__ void chain_offsets_t(const TeamCtx &tm, const KArgs &a, const TokGeom &t, const float *tab, Misc *ms) {
    for (int s = tm.tid; s < t.S; s += BT_TT) {
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
    tm.sync();
    for (int k = tm.tid; k < t.K; k += BT_TT) {      // part.py:316-326: the serial recurrence, one thread per stroke
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
    tm.sync();
}

__device__ __forceinline__ void affine_setup_t(const TeamCtx &tm, const KArgs &a, const TokGeom &t, int p, const float *tab, Misc *ms) {
    double sx = 0.0, sy = 0.0;      // com_char (util/stroke.py:134-135): all S * neval points, in bounds or not
    for (int s = tm.warp; s < t.S; s += TeamCtx::NW) {
        const CtrlSrc src = make_ctrl_src(a, t, s, tab, ms, false);
        for (int j = tm.lane; j < a.neval; j += 32) {
            float x, y;
            src.pre(j, x, y);
            sx += (double)x; sy += (double)y;
        }
    }
    double v[2] = {sx, sy};
    team_sum<2>(tm, v, ms->red);
    if (tm.tid == 0) {
        const double npts = (double)t.S * (double)a.neval;
        const float cx = npts > 0.0 ? (float)(v[0] / npts) : 0.0f, cy = npts > 0.0 ? (float)(v[1] / npts) : 0.0f;      // (no points: nothing depends on the warp)
        const float *A = a.affine + 4 * (size_t)p;
        ms->com[0] = cx; ms->com[1] = cy;
        ms->affB[0] = A[0]; ms->affB[1] = A[1];                                  // affine.py:48-53
        ms->affB[2] = __fsub_rn(A[2], __fmul_rn(__fsub_rn(A[0], 1.0f), cx));
        ms->affB[3] = __fsub_rn(A[3], __fmul_rn(__fsub_rn(A[1], 1.0f), cy));
    }
    tm.sync();
}

__device__ __forceinline__ void token_consts_t(const TeamCtx &tm, float eps, float sigma, int fhalf, Misc *ms) {
    if (tm.warp != TeamCtx::NW - 1) return;      // the team's last warp; only it reads or writes c_eps / c_sigma
    const bool same_eps = (eps == ms->c_eps), same_sigma = (sigma == ms->c_sigma);
    __syncwarp();
    if (tm.lane == 0) { ms->c_eps = eps; ms->c_sigma = sigma; }
    if (!same_sigma) token_consts_sigma(sigma, tm.lane, fhalf, ms);
    if (!same_eps && (tm.lane == 8 || tm.lane == 9)) token_consts_eps(eps, tm.lane == 9, ms);
}

__device__ __forceinline__ void load_image_bits_t(const TeamCtx &tm, const KArgs &a, int p, Misc *ms) {
    if (!a.image_bits) return;
    const int t = a.img_idx ? a.img_idx[p] : 0;
    if (t == ms->cur_img) return;      // team-uniform; cur_img is only written after the next team barrier
    int ones = 0;
    for (int i = tm.tid; i < IMG_WORDS; i += BT_TT) {
        const uint32_t wv = a.image_bits[(size_t)t * IMG_WORDS + i];
        ms->bits[i] = wv;
        ones += __popc(wv);
    }
    ones = __reduce_add_sync(0xffffffffu, ones);
    if (tm.lane == 0 && ones) atomicAdd(&ms->ones_acc, ones);
}

// bounding box of the pixels the splat will touch (floor / ceil of every on-page point), and the off-page flag
__device__ __forceinline__ bool warp_bbox(bool active, const CtrlSrc &src, int n, int q, int lane, int *bbox) {
    int rmin = IMG, rmax = -1, cmin = IMG, cmax = -1;
    bool any_out = false;
    BPL_UNROLL_A_BBOX
    for (int k = 0; k < BT_PPL; ++k) {
        const int j = q * BT_PPL + k;
        const bool valid = active && j < n;
        const PEval e = eval_pt(src, j, valid);
        any_out |= valid && !e.inb;
        if (e.inb) {
            rmin = min(rmin, (int)floorf(e.r)); rmax = max(rmax, (int)ceilf(e.r));
            cmin = min(cmin, (int)floorf(e.c)); cmax = max(cmax, (int)ceilf(e.c));
        }
    }
    rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
    cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
    if (lane == 0 && rmax >= 0) {
        atomicMin(bbox + 0, rmin); atomicMax(bbox + 1, rmax);
        atomicMin(bbox + 2, cmin); atomicMax(bbox + 3, cmax);
    }
    return any_out;
}

// ------------------------------------------------------------------------------------------------
// One separable pass over a REGION of the box.  Every pass has its own region: the ink box grown by the margin up to
// which its output can be non-zero (forward) or is still needed (adjoint), see bt_blur.
//   lines [l0, l0 + nl) across the lanes, positions [s0, s0 + 5 * nblk) along the sliding axis (local coordinates,
//   s0 a multiple of 5); positions outside [0, L) are the zero padding.
// The orientation is data: SA = stride along the sliding axis, SL = stride across the lines (a pass down the columns
// has SA = row stride, SL = 1), so one instantiation serves both axes.
// INPLACE: the epilogue may overwrite `in`: every run takes the halo it needs from the neighbouring runs before a
//   team barrier.  Otherwise the pass has no barrier and its last block reads the trailing halo from memory.
// CLAMP1: inputs pass through min(., 1) (clamp_(max=1) of rendering.py:171 on the fly).
// epi(line, pos, q, acc): q = offset of the output in the buffer.
// ------------------------------------------------------------------------------------------------
struct Reg { int l0, nl, s0, nblk; };
struct Axis { int SA, SL, L; };

// floor(a / b) for a / b < 2048, 1 <= b <= 1024 without an integer division ((a + 1/2) / b is at least 1/(2b) away
// from every integer, the float error is below 2048 * 2^-22)
__device__ __forceinline__ int small_div(int a, float rcp_b) { return __float2int_rd(((float)a + 0.5f) * rcp_b); }

__device__ __forceinline__ Axis make_axis(const LBox &bx, bool vert) {
    Axis x;
    x.SA = vert ? bx.stride : 1; x.SL = vert ? 1 : bx.stride; x.L = vert ? bx.H : bx.W;
    return x;
}
// region of a pass: the ink box grown by mr rows and mc columns, whole 5-blocks along the sliding axis
__device__ __forceinline__ Reg make_reg(const BtTok *tk, bool vert, int mr, int mc) {
    const int lr0 = max(0, tk->bbox[0] - mr) - tk->bx.r0, lr1 = min(IMG - 1, tk->bbox[1] + mr) - tk->bx.r0;
    const int lc0 = max(0, tk->bbox[2] - mc) - tk->bx.c0, lc1 = min(IMG - 1, tk->bbox[3] + mc) - tk->bx.c0;
    Reg g;
    const int a0 = vert ? lr0 : lc0, a1 = vert ? lr1 : lc1;
    g.l0 = vert ? lc0 : lr0;
    g.nl = (vert ? lc1 : lr1) - g.l0 + 1;
    g.s0 = a0 / BT_GU * BT_GU;
    g.nblk = a1 / BT_GU - a0 / BT_GU + 1;
    return g;
}

template <int TAPS, int NF, bool INPLACE, bool CLAMP1, class Epi>
__device__ __forceinline__ void sweep_core(const TeamCtx &tm, const float *__restrict__ in, const Axis ax, const Reg g,
                                           const float *__restrict__ w0, const float *__restrict__ w1, Epi epi) {
    auto ld = [](float v) { return CLAMP1 ? fminf(v, 1.0f) : v; };
    constexpr int HALF = TAPS / 2, GU = BT_GU;
    float w[NF][HALF + 1];
#pragma unroll
    for (int d = 0; d <= HALF; ++d) {
        w[0][d] = w0[d];
        if (NF > 1) w[NF - 1][d] = w1[d];
    }
    const float rnl = __frcp_rn((float)g.nl);
    const int NB = min(g.nblk, small_div(BT_TT, rnl));                    // runs per line (>= 1: nl <= IMG <= BT_TT)
    const int qb = small_div(g.nblk, __frcp_rn((float)NB)), rem = g.nblk - qb * NB;      // the first `rem` runs have qb + 1 blocks
    const int bi = small_div(tm.tid, rnl), line = g.l0 + tm.tid - bi * g.nl;
    const bool active = bi < NB;
    const int nb = qb + (bi < rem ? 1 : 0);
    const int a0 = g.s0 + (bi * qb + min(bi, rem)) * GU, len = nb * GU;
    const int SA = ax.SA;
    const int q0 = a0 * SA + line * ax.SL;
    const float *p = in + q0;
    float x[GU + 2 * HALF];
    float trail[HALF];
    const bool more = a0 + len < ax.L;      // something lies behind the run
    if (active) {
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            x[i] = (a0 > 0) ? ld(p[(i - HALF) * SA]) : 0.0f;
            if (INPLACE) trail[i] = more ? ld(p[(len + i) * SA]) : 0.0f;
        }
    }
    if (INPLACE) tm.sync();
    if (!active) return;
#pragma unroll
    for (int i = 0; i < HALF; ++i) x[HALF + i] = ld(p[i * SA]);
#pragma unroll 1
    for (int blk = 0; blk < nb; ++blk) {
        const float *qp = p + blk * GU * SA;
        const bool lastb = (blk == nb - 1);
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            if (u + HALF < GU) x[2 * HALF + u] = ld(qp[(u + HALF) * SA]);
            else if (INPLACE) x[2 * HALF + u] = lastb ? trail[u + HALF - GU] : ld(qp[(u + HALF) * SA]);
            else x[2 * HALF + u] = (!lastb || more) ? ld(qp[(u + HALF) * SA]) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float acc[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] = w[f][0] * x[u + HALF];
#pragma unroll
            for (int d = 1; d <= HALF; ++d) {
                const float s = x[u + HALF - d] + x[u + HALF + d];
#pragma unroll
                for (int f = 0; f < NF; ++f) acc[f] = fmaf(w[f][d], s, acc[f]);
            }
            epi(line, a0 + blk * GU + u, q0 + (blk * GU + u) * SA, acc);
        }
#pragma unroll
        for (int k = 0; k < 2 * HALF; ++k) x[k] = x[k + GU];
    }
}

// ---- the pass types of the blur chain (one instantiation each, both orientations) ----
// out = f * in
__device__ __noinline__ void bt_pass_store(const float *in, float *out, Axis ax, Reg g, const float *w0) {
    const TeamCtx tm = team_ctx();
    sweep_core<BPL_FSIZE, 1, false, false>(tm, in, ax, g, w0, w0, [&](int, int, int q, const float (&acc)[1]) { out[q] = acc[0]; });
}
// out = f * min(in, 1)
__device__ __noinline__ void bt_pass_store_clamp(const float *in, float *out, Axis ax, Reg g, const float *w0) {
    const TeamCtx tm = team_ctx();
    sweep_core<BPL_FSIZE, 1, false, true>(tm, in, ax, g, w0, w0, [&](int, int, int q, const float (&acc)[1]) { out[q] = acc[0]; });
}
// o0 = f0 * io, io = f1 * io (in place)
__device__ __noinline__ void bt_pass_split(float *io, float *o0, Axis ax, Reg g, const float *w0, const float *w1) {
    const TeamCtx tm = team_ctx();
    sweep_core<BPL_FSIZE, 2, true, false>(tm, io, ax, g, w0, w1, [&](int, int, int q, const float (&acc)[2]) {
        o0[q] = acc[0]; io[q] = acc[1]; });
}
// io = t = f0 * io (in place); returns sum t . W
__device__ __noinline__ float bt_pass_dot(float *io, const float *W, Axis ax, Reg g, const float *w0) {
    const TeamCtx tm = team_ctx();
    float s = 0.0f;
    sweep_core<BPL_FSIZE, 1, true, false>(tm, io, ax, g, w0, w0, [&](int, int, int q, const float (&acc)[1]) {
        s = fmaf(acc[0], W[q], s); io[q] = acc[0]; });
    return s;
}
// returns sum min(I2, 1) . (f0 * in)
__device__ __noinline__ float bt_pass_reduce(const float *in, const float *I2, Axis ax, Reg g, const float *w0) {
    const TeamCtx tm = team_ctx();
    float s = 0.0f;
    sweep_core<BPL_FSIZE, 1, false, false>(tm, in, ax, g, w0, w0, [&](int, int, int q, const float (&acc)[1]) {
        s = fmaf(fminf(I2[q], 1.0f), acc[0], s); });
    return s;
}
// returns sum min(I2, 1) . (f1 * in); I2 = [I2 <= 1] (f0 * in)   (clamp_(max=1) mask, rendering.py:171)
__device__ __noinline__ float bt_pass_mask(const float *in, float *I2, Axis ax, Reg g, const float *w0, const float *w1) {
    const TeamCtx tm = team_ctx();
    float s = 0.0f;
    sweep_core<BPL_FSIZE, 2, false, false>(tm, in, ax, g, w0, w1, [&](int, int, int q, const float (&acc)[2]) {
        const float i2 = I2[q];
        s = fmaf(fminf(i2, 1.0f), acc[1], s);
        I2[q] = (i2 <= 1.0f) ? acc[0] : 0.0f; });
    return s;
}

// P9 and P8 of the blur chain in ONE pass (same orientation, same region): two windows slide together,
//   sig += min(I2,1) . (f1 * inB + f0 * inC);   I2 = [I2 <= 1] (f0 * inB)
// -- one pass set-up, one item mapping and one loop instead of two.
__device__ __noinline__ float bt_pass_mask2(const float *__restrict__ inB, const float *__restrict__ inC, float *I2, Axis ax, Reg g,
                                            const float *w0p, const float *w1p) {
    const TeamCtx tm = team_ctx();
    constexpr int HALF = PAD, GU = BT_GU;
    float w0[HALF + 1], w1[HALF + 1];
#pragma unroll
    for (int d = 0; d <= HALF; ++d) { w0[d] = w0p[d]; w1[d] = w1p[d]; }
    const float rnl = __frcp_rn((float)g.nl);
    const int NB = min(g.nblk, small_div(BT_TT, rnl));
    const int qb = small_div(g.nblk, __frcp_rn((float)NB)), rem = g.nblk - qb * NB;
    const int bi = small_div(tm.tid, rnl), line = g.l0 + tm.tid - bi * g.nl;
    if (bi >= NB) return 0.0f;
    const int nb = qb + (bi < rem ? 1 : 0);
    const int a0 = g.s0 + (bi * qb + min(bi, rem)) * GU, len = nb * GU;
    const int SA = ax.SA;
    const int q0 = a0 * SA + line * ax.SL;
    const float *pb = inB + q0, *pc = inC + q0;
    const bool more = a0 + len < ax.L;
    float xb[GU + 2 * HALF], xc[GU + 2 * HALF];
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        xb[i] = (a0 > 0) ? pb[(i - HALF) * SA] : 0.0f; xc[i] = (a0 > 0) ? pc[(i - HALF) * SA] : 0.0f;
        xb[HALF + i] = pb[i * SA]; xc[HALF + i] = pc[i * SA];
    }
    float s = 0.0f;
#pragma unroll 1
    for (int blk = 0; blk < nb; ++blk) {
        const int o = blk * GU * SA;
        const bool live = blk < nb - 1 || more;      // inputs beyond the last block: the box's zero padding
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            xb[2 * HALF + u] = live ? pb[o + (u + HALF) * SA] : 0.0f;
            xc[2 * HALF + u] = live ? pc[o + (u + HALF) * SA] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float g3 = w0[0] * xb[u + HALF], sB = w1[0] * xb[u + HALF], sC = w0[0] * xc[u + HALF];
#pragma unroll
            for (int d = 1; d <= HALF; ++d) {
                const float tb = xb[u + HALF - d] + xb[u + HALF + d], tc = xc[u + HALF - d] + xc[u + HALF + d];
                g3 = fmaf(w0[d], tb, g3); sB = fmaf(w1[d], tb, sB); sC = fmaf(w0[d], tc, sC);
            }
            const int q = q0 + o + u * SA;
            const float i2 = I2[q];
            s = fmaf(fminf(i2, 1.0f), sB + sC, s);
            I2[q] = (i2 <= 1.0f) ? g3 : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 2 * HALF; ++k) { xb[k] = xb[k + GU]; xc[k] = xc[k + GU]; }
    }
    return s;
}

struct TailAcc { float ll, sig; double eps; };

// pointwise tail of the forward + head of the backward for LOCAL pixel (r,c) with blurred value v: returns d L / d I5
// (the cotangent entering the second blur pass).  Constants come from shared memory.
__device__ __forceinline__ float bt_tail(const BtShared *sh, const BtTok *tk, const Misc *ms, int r, int c, float v, TailAcc &acc) {
    const int gi = (tk->bx.r0 + r) * IMG + tk->bx.c0 + c;
    const float eps = tk->eps, bgv = tk->bgv, gscale = tk->gscale;
    const bool pass5 = (v >= 0.0f) && (v <= 1.0f);                  // clamp_(0,1) passes gradient on [0,1]
    const float i6 = fminf(fmaxf(v, 0.0f), 1.0f);
    float pm = i6;
    if (eps > 0.0f) pm = __fadd_rn(__fmul_rn(tk->om, i6), __fmul_rn(eps, __fsub_rn(1.0f, i6)));      // rendering.py:226-227
    float gp;
    if (sh->ka.g_pimg) {
        gp = sh->ka.g_pimg[(size_t)tk->p * NPIX + gi];
        if (eps > 0.0f) acc.eps += (double)(-2.0f * gp * i6);        // d/d eps of (1-eps) p + eps (1-p) is 1 - 2 p;
                                                                    // the sum of gp over the canvas is added by bt_scalars
    } else {
        const int x = (sh->ka.image_bits != nullptr) ? (int)((ms->bits[gi >> 5] >> (gi & 31)) & 1u) : 0;
        float d = ms->dlp_bg[x];
        if (pm != bgv) {                                             // a background pixel adds nothing to the sum
            const float lp = pixel_lp_fwd(pm, x != 0);
            d = (pm < TORCH_EPS || pm > 1.0f - TORCH_EPS) ? 0.0f : (x ? __frcp_rn(pm) : -__frcp_rn(1.0f - pm));
            acc.ll += lp - ms->lp_bg[x];
        }
        // A pixel without ink has the analytic background value (closed form in bt_scalars); pm == bgv alone does not
        // mean "no ink" (eps = 0.5 maps every v to 0.5).
        if (eps > 0.0f && (pm != bgv || i6 != 0.0f))
            acc.eps += ((double)(d * (1.0f - 2.0f * i6)) - ms->dlp_bg_d[x]) * (double)gscale;
        gp = d * gscale;
    }
    return pass5 ? gp * tk->mixd : 0.0f;
}

// I5 = f0 * io along the rows; io = G5 = tail(I5) (in place); sig += G5 . (f1 * io)
__device__ __noinline__ TailAcc bt_pass_tail(BtShared *sh, float *io, Axis ax, Reg g) {
    const TeamCtx tm = team_ctx();
    const BtTok *tk = &sh->tok[tm.id];
    const Misc *ms = &sh->ms[tm.id];
    TailAcc acc{0.0f, 0.0f, 0.0};
    sweep_core<BPL_FSIZE, 2, true, false>(tm, io, ax, g, ms->g, ms->ht, [&](int line, int pos, int q, const float (&a2)[2]) {
        const float g5 = bt_tail(sh, tk, ms, line, pos, a2[0], acc);
        acc.sig = fmaf(g5, a2[1], acc.sig);
        io[q] = g5; });
    return acc;
}

// The blur chain and its adjoint.  Region of every pass = the ink box grown by (rows, columns):
//   forward : the margin up to which the pass's output can be non-zero
//   adjoint : the margin up to which its output is still needed by what follows
__device__ __noinline__ TailAcc bt_blur(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    BtTok *tk = &sh->tok[tm.id];
    const Misc *ms = &sh->ms[tm.id];
    const LBox bx = tk->bx;
    const int area = bx.area(), nc = sh->ka.rc.ink_ncon;
    float *A = bt_pool(sh) + tk->arena, *B = A + area, *C = B + area;
    const Axis V = make_axis(bx, true), H = make_axis(bx, false);
    const float *wg = ms->g, *wh = ms->ht;
    TailAcc acc{0.0f, 0.0f, 0.0};
    if (tk->sigma > 0.0f) {
        bt_pass_store_clamp(A, B, V, make_reg(tk, true, nc + PAD, nc), wg);                       // P1: B = g_v min(A,1)
        tm.sync();
        bt_pass_store(B, C, H, make_reg(tk, false, nc + PAD, nc + PAD), wg);             // P2: C = I4 = g_h B
        tm.sync();
        bt_pass_split(C, B, V, make_reg(tk, true, nc + 2 * PAD, nc + PAD), wg, wh);                // P3: B = U = g_v I4, C = W = h~_v I4
        tm.sync();
        acc = bt_pass_tail(sh, B, H, make_reg(tk, false, nc + 2 * PAD, nc + 2 * PAD));             // P4: B = G5; sig += G5 . (h~_h U)
        tm.sync();
        acc.sig += bt_pass_dot(B, C, H, make_reg(tk, false, nc + 2 * PAD, nc + PAD), wg);          // P5: B = t = g_h G5; sig += t . W
        tm.sync();
        bt_pass_store(B, C, V, make_reg(tk, true, nc + PAD, nc + PAD), wg);              // P6: C = G4 = g_v t
        tm.sync();
        bt_pass_split(C, B, H, make_reg(tk, false, nc + PAD, nc), wg, wh);                         // P7: B = t1 = g_h G4, C = t1' = h~_h G4
        tm.sync();
#ifdef BPL_BT_SPLIT_P98
        const Reg g8 = make_reg(tk, true, nc, nc);
        acc.sig += bt_pass_reduce(C, A, V, g8, wg);                                                // P9: sig += I3 . (g_v t1')
        acc.sig += bt_pass_mask(B, A, V, g8, wg, wh);                                              // P8: sig += I3 . (h~_v t1); A = G3
#else
        acc.sig += bt_pass_mask2(B, C, A, V, make_reg(tk, true, nc, nc), wg, wh);                  // P9 + P8 in one pass
#endif
        tm.sync();
    } else {
        const Reg g0 = make_reg(tk, false, nc, nc);      // rows l0.., columns s0.. (whole 5-blocks)
        const int wn = g0.nblk * BT_GU, bn = wn * g0.nl;
        const float rw = __frcp_rn((float)wn);
        for (int i = tm.tid; i < bn; i += BT_TT) {
            const int rr = small_div(i, rw), r = g0.l0 + rr, c = g0.s0 + i - rr * wn;
            float *q = A + r * bx.stride + c;
            const float i2 = *q;
            const float gq = bt_tail(sh, tk, ms, r, c, fminf(i2, 1.0f), acc);
            *q = (i2 <= 1.0f) ? gq : 0.0f;
        }
        tm.sync();
    }
    return acc;
}

__constant__ float c_w3[2] = {2.0f, 1.0f};      // the (1,2,1) factor of H_broaden

__device__ __noinline__ void bt_pass3_store(const float *in, float *out, Axis ax, Reg g) {
    const TeamCtx tm = team_ctx();
    sweep_core<3, 1, false, false>(tm, in, ax, g, c_w3, c_w3, [&](int, int, int q, const float (&acc)[1]) { out[q] = acc[0]; });
}
__device__ __noinline__ void bt_pass3_update(const float *in, float *A, Axis ax, Reg g, float c1, float c2) {
    const TeamCtx tm = team_ctx();
    sweep_core<3, 1, false, false>(tm, in, ax, g, c_w3, c_w3, [&](int, int, int q, const float (&acc)[1]) {
        A[q] = fmaf(c1, acc[0], c2 * A[q]); });
}

// One broaden pass as ONE two-dimensional sweep, out of place: out = c1 * (1,2,1)x(1,2,1) * in + c2 * in over the region.
// A thread owns a run of rows of one column and slides down it: per row it loads the three horizontal neighbours (lanes
// hold adjacent columns: conflict free), forms the row's (1,2,1) sum and combines the sums of three consecutive rows.
// Half the passes, barriers and per-pass set-up of the separable form (which is kept below for A/B: -DBPL_BT_BROADEN_SEP).
__device__ __noinline__ void bt_pass_broaden2d(const float *__restrict__ in, float *__restrict__ out, int stride, int H, int W,
                                               Reg g, float c1, float c2) {
    const TeamCtx tm = team_ctx();
    constexpr int GU = BT_GU;
    const float rnl = __frcp_rn((float)g.nl);
    const int NB = min(g.nblk, small_div(BT_TT, rnl));
    const int qb = small_div(g.nblk, __frcp_rn((float)NB)), rem = g.nblk - qb * NB;
    const int bi = small_div(tm.tid, rnl), col = g.l0 + tm.tid - bi * g.nl;
    if (bi >= NB) return;
    const int nb = qb + (bi < rem ? 1 : 0);
    const int a0 = g.s0 + (bi * qb + min(bi, rem)) * GU, a1 = a0 + nb * GU;      // rows [a0, a1)
    const bool okl = col > 0, okr = col < W - 1;
    const float *p = in + a0 * stride + col;
    float *o = out + a0 * stride + col;
    auto hsum = [&](const float *q, float &xc) {      // (1,2,1) sum of one row at this column; zero outside the box
        xc = q[0];
        const float xl = okl ? q[-1] : 0.0f, xr = okr ? q[1] : 0.0f;
        return xl + 2.0f * xc + xr;
    };
    float xc, xn, dummy;
    float hp = (a0 > 0) ? hsum(p - stride, dummy) : 0.0f;
    float hc = hsum(p, xc);
#pragma unroll 1
    for (int r = a0; r < a1; r += GU) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const float hn = (r + u + 1 < H) ? hsum(p + stride, xn) : 0.0f;
            *o = fmaf(c1, hp + 2.0f * hc + hn, c2 * xc);
            hp = hc; hc = hn; xc = xn;
            p += stride; o += stride;
        }
    }
}

// broaden (rendering.py:160-168): ncon x [ c1 * S_h S_v + c2 ]; A holds the image and receives the result, B is scratch.
// m: margin (around the ink box) up to which the input is non-zero (forward, dm = +1: the image grows by one pixel per
// pass) or the adjoint's input is valid (dm = -1: the region in which the result is still needed shrinks).
__device__ __noinline__ void bt_broaden(BtShared *sh, int m, int dm) {
    const TeamCtx tm = team_ctx();
    const BtTok *tk = &sh->tok[tm.id];
    const LBox bx = tk->bx;
    float *A = bt_pool(sh) + tk->arena, *T = A + bx.area();
    const float c1 = sh->ka.rc.c1, c2 = sh->ka.rc.c2;
    const int ncon = sh->ka.rc.ink_ncon;
#ifndef BPL_BT_BROADEN_SEP
    // ping-pong A -> T -> A ...: a pass reads its input one pixel beyond its output region; there the input holds either
    // what the previous pass wrote (adjoint: the regions shrink) or the zeros nothing has written yet (forward: T is
    // first used here, and the true image is zero beyond the previous region)
    for (int it = 0; it < ncon; ++it) {
        const int mo = m + dm;
        bt_pass_broaden2d((it & 1) ? T : A, (it & 1) ? A : T, bx.stride, bx.H, bx.W, make_reg(tk, true, mo, mo), c1, c2);
        tm.sync();
        m = mo;
    }
    if (ncon & 1) {      // an odd number of passes leaves the result in T
        const Reg g = make_reg(tk, false, m, m);
        const int wn = g.nblk * BT_GU, bn = wn * g.nl;
        const float rw = __frcp_rn((float)wn);
        for (int i = tm.tid; i < bn; i += BT_TT) {
            const int rr = small_div(i, rw), q = (g.l0 + rr) * bx.stride + g.s0 + i - rr * wn;
            A[q] = T[q];
        }
        tm.sync();
    }
#else
    const Axis V = make_axis(bx, true), H = make_axis(bx, false);
    for (int it = 0; it < ncon; ++it) {
        const int mo = m + dm;
        bt_pass3_store(A, T, V, make_reg(tk, true, mo, max(m, mo)));      // rows of the result, columns the row pass reads
        tm.sync();
        bt_pass3_update(T, A, H, make_reg(tk, false, mo, mo), c1, c2);
        tm.sync();
        m = mo;
    }
#endif
}

// Per-token set-up: target bits, chain offsets, affine, constants, the ink box and the team's arena (zeroed).
__device__ __noinline__ void bt_setup(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const float *tab = sh->tab;
    const int p = tk->p;
    const TokGeom t = tk->t;
    const bool aff = a.affine != nullptr;
    load_image_bits_t(tm, a, p, ms);
    if (tm.tid == 0) {
        ms->off_page = 0;
        ms->aff_acc[0] = ms->aff_acc[1] = ms->aff_acc[2] = ms->aff_acc[3] = 0.0;
        tk->bbox[0] = IMG; tk->bbox[1] = -1; tk->bbox[2] = IMG; tk->bbox[3] = -1;
    }
    chain_offsets_t(tm, a, t, tab, ms);
    if (aff) affine_setup_t(tm, a, t, p, tab, ms);
    token_consts_t(tm, a.eps[p], a.sigma[p], a.rc.fhalf, ms);      // double-precision constants by one warp, hidden behind the point stage
    if (tm.tid == 0 && a.image_bits) {
        const int img = a.img_idx ? a.img_idx[p] : 0;
        if (img != ms->cur_img) { ms->n_ones = ms->ones_acc; ms->cur_img = img; }      // a new image was loaded above
    }
    // ---- where will the ink fall? ----
    const int QN = (a.neval + BT_PPL - 1) / BT_PPL, items = t.S * QN;
    {
        bool any_out = false;
        for (int base = 0; base < items; base += BT_TT) {
            const int it = base + tm.tid;
            const bool active = it < items;
            const int s = active ? it / QN : 0;
            CtrlSrc src{};
            if (active) src = make_ctrl_src(a, t, s, tab, ms, aff);
            any_out |= warp_bbox(active, src, a.neval, active ? it - s * QN : 0, tm.lane, tk->bbox);
        }
        any_out = __any_sync(0xffffffffu, any_out);
        if (any_out && tm.lane == 0) atomicOr(&ms->off_page, 1);
    }
    tm.sync();
    if (tm.tid == 0) {      // the box, its arena, the tail's constants
        LBox bx{0, 0, 0, 0, 1};
        if (tk->bbox[1] >= 0) {
            const int reach = a.rc.ink_ncon + 2 * PAD;
            const int r0 = max(0, tk->bbox[0] - reach), r1 = min(IMG - 1, tk->bbox[1] + reach);
            const int c0 = max(0, tk->bbox[2] - reach), c1 = min(IMG - 1, tk->bbox[3] + reach);
            bx.H = (r1 - r0 + BT_GU) / BT_GU * BT_GU;      // whole 5-blocks (105 = 21 * 5, so H <= 105) ...
            bx.W = (c1 - c0 + BT_GU) / BT_GU * BT_GU;
            bx.r0 = min(r0, IMG - bx.H);                   // ... shifted inwards to stay on the canvas
            bx.c0 = min(c0, IMG - bx.W);
            bx.stride = bx.W | 1;
        }
        tk->bx = bx;
        const float eps = a.eps[p];
        tk->eps = eps;
        tk->sigma = a.sigma[p];
        tk->om = 1.0f - eps;
        tk->mixd = eps > 0.0f ? (1.0f - 2.0f * eps) : 1.0f;
        tk->bgv = eps > 0.0f ? eps : 0.0f;
        tk->gscale = (!a.g_pimg && a.g_ll) ? a.g_ll[p] : 1.0f;
        tk->need = 3 * bx.area() + ((t.S * NACC + 3) & ~3) + 4;
        tk->arena = arena_alloc(*sh, tm.id, tk->need, sh->pool_floats);
    }
    tm.sync();
    {
        float4 *z4 = reinterpret_cast<float4 *>(bt_pool(sh) + tk->arena);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n4 = tk->need / 4;
        for (int i = tm.tid; i < n4; i += BT_TT) z4[i] = z;
    }
    tm.sync();
}

// the rare min-ink branches (rendering.py:104-107): a second pass adds (final - common) ink.  Its own function: cold
// code stays out of the instruction stream the teams share.
__device__ __noinline__ void bt_splat_fix(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    const BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const float *tab = sh->tab;
    const TokGeom t = tk->t;
    const bool aff = a.affine != nullptr;
    const LBox bx = tk->bx;
    float *A = bt_pool(sh) + tk->arena;
    const LayD lay{bx.stride, -(bx.r0 * bx.stride + bx.c0)};
    const int QN = (a.neval + BT_PPL - 1) / BT_PPL, items = t.S * QN;
    for (int base = 0; base < items; base += BT_TT) {
        const int it = base + tm.tid;
        bool active = it < items;
        const int s = active ? it / QN : 0x7fffffff;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) v = ms->sub[s];
        active = active && (__float_as_int(v.w) >= 2 && v.z < a.rc.ink_pp);
        CtrlSrc src{};
        if (active) src = make_ctrl_src(a, t, s, tab, ms, aff);
        warp_points<true, 1, IMG, 0, BT_PPL, LayD, BT_ROLL_SPLAT>(active, src, tab, s, a.neval, it < items ? it - s * QN : 0, tm.lane,
                                                  a.rc, A, v.z, __float_as_int(v.w), nullptr, lay);
    }
    tm.sync();
}

// forward: splat (rendering.py:58-127) into buffer A of the arena
__device__ __noinline__ void bt_splat(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    const BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const float *tab = sh->tab;
    const TokGeom t = tk->t;
    const bool aff = a.affine != nullptr;
    const LBox bx = tk->bx;
    float *A = bt_pool(sh) + tk->arena;
    const LayD lay{bx.stride, -(bx.r0 * bx.stride + bx.c0)};
    const int QN = (a.neval + BT_PPL - 1) / BT_PPL, items = t.S * QN;
    for (int base = 0; base < items; base += BT_TT) {
        const int it = base + tm.tid;
        const bool active = it < items;
        const int s = active ? it / QN : 0x7fffffff;
        CtrlSrc src{};
        if (active) src = make_ctrl_src(a, t, s, tab, ms, aff);
        const LaneAcc acc = warp_points<false, 1, IMG, 0, BT_PPL, LayD, BT_ROLL_SPLAT>(active, src, tab, s, a.neval, active ? it - s * QN : 0,
                                                                       tm.lane, a.rc, A, 0.0f, 0, nullptr, lay);
        reduce_by_substroke(active, s, acc.sum, acc.cnt, ms->sub, ms->ink64, tm.lane);
    }
    tm.sync();
    bool need_fix = false;
    for (int s = 0; s < t.S; ++s) {
        float4 v = ms->sub[s];
        v.z = ink_sum(ms, s);
        need_fix |= (__float_as_int(v.w) >= 2 && v.z < a.rc.ink_pp);
    }
    if (need_fix) bt_splat_fix(sh);
}

// scalar outputs: ll, off_page, d/d eps, d/d sigma
__device__ __noinline__ void bt_scalars(BtShared *sh, TailAcc acc) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    const BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const int p = tk->p;
    const float eps = tk->eps, sigma = tk->sigma;
    const float *gP_ext = a.g_pimg ? a.g_pimg + (size_t)p * NPIX : nullptr;
    if (gP_ext && eps > 0.0f)
        for (int i = tm.tid; i < NPIX; i += BT_TT) acc.eps += (double)gP_ext[i];
    double v[3] = {(double)acc.ll, acc.eps, (double)acc.sig};
    team_sum<3>(tm, v, ms->red);
    if (tm.tid == 0) {
        // the all-background image: lp_bg / dlp_bg at every pixel
        const bool score = a.image_bits != nullptr;
        const double n1 = (double)ms->n_ones, n0 = (double)(NPIX - ms->n_ones);
        if (score && !gP_ext) {
            v[0] += n1 * (double)ms->lp_bg[1] + n0 * (double)ms->lp_bg[0];
            if (eps > 0.0f) v[1] += (double)tk->gscale * (n1 * ms->dlp_bg_d[1] + n0 * ms->dlp_bg_d[0]);
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

// rescale branch (rendering.py:106-107) only: sum_k gink[k] ink0[k] per sub-stroke, needed before the gather proper
__device__ __noinline__ void bt_gather_dot(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    const BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const float *tab = sh->tab;
    const TokGeom t = tk->t;
    const bool aff = a.affine != nullptr;
    const LBox bx = tk->bx;
    const float *A = bt_pool(sh) + tk->arena;
    const LayD lay{bx.stride, -(bx.r0 * bx.stride + bx.c0)};
    const int QN = (a.neval + BT_PPL - 1) / BT_PPL, items = t.S * QN;
    auto noemit = [](int, float, float) {};
    for (int base = 0; base < items; base += BT_TT) {
        const int it = base + tm.tid;
        bool active = it < items;
        const int s = active ? it / QN : 0x7fffffff;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) v = ms->sub[s];
        active = active && (__float_as_int(v.w) >= 2 && v.z >= 2.22e-6f && v.z < a.rc.ink_pp);
        CtrlSrc src{};
        if (active) src = make_ctrl_src(a, t, s, tab, ms, aff);
        const float d = warp_points_bwd<true, IMG, 0, BT_PPL>(active, src, tab, s, a.neval, it < items ? it - s * QN : 0,
                                                              tm.lane, a.rc, A, v.z, __float_as_int(v.w), 0.0f, noemit, lay);
        unsigned todo = __ballot_sync(0xffffffffu, active);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int cur = __shfl_sync(0xffffffffu, s, leader);
            const bool mine = active && s == cur;
            const float tot = warp_sum(mine ? d : 0.0f);
            if (tm.lane == leader) atomicAdd(&ms->dot[cur], tot);
            todo &= ~__ballot_sync(0xffffffffu, mine);
        }
    }
    tm.sync();
}

// add_stroke backward: gather the image gradient (buffer A) at the splat points; per sub-stroke sums into accs
__device__ __noinline__ void bt_gather(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    const BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const float *tab = sh->tab;
    const TokGeom t = tk->t;
    const bool aff = a.affine != nullptr;
    const LBox bx = tk->bx;
    const float *A = bt_pool(sh) + tk->arena;
    float *accs = bt_pool(sh) + tk->arena + 3 * bx.area();      // [S][NACC], zero since the arena was cleared
    const LayD lay{bx.stride, -(bx.r0 * bx.stride + bx.c0)};
    const int QN = (a.neval + BT_PPL - 1) / BT_PPL, items = t.S * QN;
    for (int i = tm.tid; i < t.S; i += BT_TT) ms->dot[i] = 0.0f;
    bool need_dot = false;
    for (int s = 0; s < t.S; ++s) {
        const float4 v = ms->sub[s];
        need_dot |= (__float_as_int(v.w) >= 2 && v.z >= 2.22e-6f && v.z < a.rc.ink_pp);
    }
    tm.sync();
    if (need_dot) bt_gather_dot(sh);
    for (int base = 0; base < items; base += BT_TT) {
        const int it = base + tm.tid;
        bool active = it < items;
        const int s = active ? it / QN : 0x7fffffff;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) v = ms->sub[s];
        active = active && __float_as_int(v.w) >= 1;
        CtrlSrc src{};
        if (active) src = make_ctrl_src(a, t, s, tab, ms, aff);
        float wx[NCPT], wy[NCPT], sx = 0.0f, sy = 0.0f, ax = 0.0f, ay = 0.0f;
#pragma unroll
        for (int k = 0; k < NCPT; ++k) { wx[k] = 0.0f; wy[k] = 0.0f; }
        warp_points_bwd<false, IMG, 0, BT_PPL>(active, src, tab, s, a.neval, it < items ? it - s * QN : 0, tm.lane, a.rc, A, v.z,
                                               __float_as_int(v.w), active ? ms->dot[s] : 0.0f,
                                               [&](int j, float gx, float gy) {
            const float *row = tab + j * NCPT;
#pragma unroll
            for (int k = 0; k < NCPT; ++k) { wx[k] = fmaf(row[k], gx, wx[k]); wy[k] = fmaf(row[k], gy, wy[k]); }
            sx += gx; sy += gy;
            if (aff) {
                float px, py;
                src.pre(j, px, py);
                ax = fmaf(gx, px, ax); ay = fmaf(gy, py, ay);
            }
        }, lay);
        // combine the lanes of a warp that share a sub-stroke: the 12 (+ 2 affine) sums are reduced together
        // (warp_reduce16: lanes 2i, 2i+1 end up with sum i), then the even lanes add their sum to the sub-stroke's record
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
            const float tot = warp_reduce16(v, tm.lane);
            const int idx = tm.lane >> 1;
            if (!(tm.lane & 1)) {
                if (idx < NACC) atomicAdd(accs + cur * NACC + idx, tot);
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
    tm.sync();
}

// chain backward (part.py:316-326) and the control-point, scale, position and affine gradients
__device__ __noinline__ void bt_chain(BtShared *sh) {
    const TeamCtx tm = team_ctx();
    const KArgs &a = sh->ka;
    const BtTok *tk = &sh->tok[tm.id];
    Misc *ms = &sh->ms[tm.id];
    const float *tab = sh->tab;
    const TokGeom t = tk->t;
    const int p = tk->p;
    const bool aff = a.affine != nullptr;
    float *accs = bt_pool(sh) + tk->arena + 3 * tk->bx.area();
    float gcx = 0.0f, gcy = 0.0f, b0 = 1.0f, b1 = 1.0f;
    if (aff) {      // p' = p * B[:2] + B[2:], B[2:] = A[2:] - (A[:2]-1) com, com = mean(p)
        const float *Aa = a.affine + 4 * (size_t)p;
        const double npts = (double)t.S * (double)a.neval;
        b0 = ms->affB[0]; b1 = ms->affB[1];
        gcx = npts > 0.0 ? (float)(-((double)Aa[0] - 1.0) * ms->aff_acc[2] / npts) : 0.0f;
        gcy = npts > 0.0 ? (float)(-((double)Aa[1] - 1.0) * ms->aff_acc[3] / npts) : 0.0f;
        if (tm.tid == 0 && a.g_affine) {
            float *ga = a.g_affine + 4 * (size_t)p;
            ga[0] = (float)(ms->aff_acc[0] - (double)ms->com[0] * ms->aff_acc[2]);
            ga[1] = (float)(ms->aff_acc[1] - (double)ms->com[1] * ms->aff_acc[3]);
            ga[2] = (float)ms->aff_acc[2];
            ga[3] = (float)ms->aff_acc[3];
        }
    } else if (tm.tid == 0 && a.g_affine) {
        float *ga = a.g_affine + 4 * (size_t)p;
        ga[0] = ga[1] = ga[2] = ga[3] = 0.0f;
    }
    const float *rfirst = tab, *rlast = tab + (a.neval - 1) * NCPT;
    // (i) one thread per stroke: suffix sums of the sub-strokes' summed gradients, on shared memory only;
    //     the carry entering each sub-stroke and its total overwrite the two Gsum slots
    for (int k = tm.tid; k < t.K; k += BT_TT) {
        const int b = a.stroke_sub_off[t.k0 + k] - t.s0, e = a.stroke_sub_off[t.k0 + k + 1] - t.s0;
        float cx = 0.0f, cy = 0.0f;       // gradient w.r.t. previous_pos of the following sub-stroke
        for (int s = e - 1; s >= b; --s) {
            float *o = accs + s * NACC;
            const float sx = fmaf(b0, o[2 * NCPT], (float)a.neval * gcx);
            const float sy = fmaf(b1, o[2 * NCPT + 1], (float)a.neval * gcy);
            ms->sub[s].z = cx; ms->sub[s].w = cy;            // carry (the forward's totals are no longer needed)
            cx += sx; cy += sy;
            o[2 * NCPT] = cx; o[2 * NCPT + 1] = cy;          // total
        }
        a.g_position[2 * (t.k0 + k)] = cx;
        a.g_position[2 * (t.k0 + k) + 1] = cy;
    }
    tm.sync();
    // (ii) one thread per sub-stroke: control-point and scale gradients (global loads in parallel)
    for (int s = tm.tid; s < t.S; s += BT_TT) {
        const float *o = accs + s * NACC;
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
    if (tm.tid == 0) ms->ones_acc = 0;
    tm.sync();      // the arena is no longer read: give it back
    if (tm.tid == 0) {
        __threadfence_block();
        *reinterpret_cast<volatile int *>(&sh->live[tm.id][1]) = 0;
    }
}

__global__ void __launch_bounds__(BT_NT *BT_TT, 1) bpl_backward_teams_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BtShared *sh = reinterpret_cast<BtShared *>(smem_raw);
    const TeamCtx tm = team_ctx();
    Misc *ms = &sh->ms[tm.id];
    BtTok *tk = &sh->tok[tm.id];
    {
        const int *src = reinterpret_cast<const int *>(&a);
        int *dst = reinterpret_cast<int *>(&sh->ka);
        for (int i = threadIdx.x; i < (int)(sizeof(KArgs) / 4); i += blockDim.x) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < a.neval * NCPT; i += blockDim.x) sh->tab[i] = a.basis[i];
    if (threadIdx.x == 0) {
        unsigned dyn_smem;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
        sh->pool_floats = (int)((dyn_smem - (unsigned)((sizeof(BtShared) + 15) / 16 * 16)) / 4) & ~3;
        sh->lock = 0; sh->waiting = -1; sh->teams_done = 0;
        for (int j = 0; j < 8; ++j) { sh->live[j][0] = 0; sh->live[j][1] = 0; }
    }
    if (tm.tid == 0) { ms->cur_img = -1; ms->c_eps = ms->c_sigma = nanf(""); ms->n_ones = 0; ms->ones_acc = 0; }
    __syncthreads();
    if (tm.warp == 0) {      // column sums of the basis table (d com / d control points), one copy per team
        float cs[NCPT];
#pragma unroll
        for (int k = 0; k < NCPT; ++k) cs[k] = 0.0f;
        for (int j = tm.lane; j < a.neval; j += 32)
#pragma unroll
            for (int k = 0; k < NCPT; ++k) cs[k] += sh->tab[j * NCPT + k];
#pragma unroll
        for (int k = 0; k < NCPT; ++k) {
            const float s = warp_sum(cs[k]);
            if (tm.lane == 0) ms->colsum[k] = s;
        }
    }
    tm.sync();

    const int n_teams = (int)gridDim.x * BT_NT;
    int item = (int)blockIdx.x * BT_NT + tm.id;              // position in the work queue (advanced by the leader)
    int p = a.n_tokens;
    if (item < a.n_tokens) p = a.order ? a.order[item] : item;
    int par = 0;
    while (p < a.n_tokens) {
        par ^= 1;
        const long long tc0 = (a.cycles_out && tm.tid == 0) ? clock64() : 0;      // measured cost, fed back into the next order
#ifdef BPL_TOKEN_TIMING
        const long long tt0 = clock64();      // cycles per token -> a.pt_buf[p] (tools/token_cost_fit.py)
#endif
        const TokGeom t = token_geom<false>(a, p);
        if (tm.tid == 0) {
            item = a.sched ? (int)atomicAdd(&a.sched->next, 1u) + n_teams : item + n_teams;
            tk->next[par] = item >= a.n_tokens ? a.n_tokens : (a.order ? a.order[item] : item);      // consumed at the end
            tk->p = p;
            tk->t = t;
        }
        if (t.S > MAXS || t.K > MAXK) {      // beyond the per-team tables: flag, do not guess
            if (tm.tid == 0) {
                if (a.ll) a.ll[p] = nanf("");
                if (a.off_page) a.off_page[p] = 255;
            }
            tm.sync();
            p = tk->next[par];
            continue;
        }
        tm.sync();
        bt_setup(sh);
        bt_splat(sh);
        TailAcc acc{0.0f, 0.0f, 0.0};
        if (tk->bx.H > 0) {
            bt_broaden(sh, 0, +1);                      // A = I2 (pre-clamp; consumers apply min(.,1) on load)
            acc = bt_blur(sh);                          // A = d L / d I2
            bt_broaden(sh, a.rc.ink_ncon, -1);          // adjoint broaden (symmetric kernel, zero padding): A = d L / d I0
        }
        bt_scalars(sh, acc);
        bt_gather(sh);
        bt_chain(sh);
        if (a.cycles_out && tm.tid == 0) a.cycles_out[p] = (unsigned)min(clock64() - tc0, 0xffffffffll);
#ifdef BPL_TOKEN_TIMING
        if (tm.tid == 0 && a.pt_buf) a.pt_buf[p] = (float)(clock64() - tt0);
#endif
        p = tk->next[par];
    }
    if (tm.tid == 0) {
        __threadfence();
        if (atomicAdd(&sh->teams_done, 1) == BT_NT - 1) sched_finish(a);
    }
}

}  // namespace bpl
