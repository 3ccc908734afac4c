// bpl_token_common.cuh -- device helpers shared by the token prior (bpl_token_prior.cu) and the token sampler
// (bpl_token_sample.cu): the spline basis row at an arbitrary coordinate and the attach point of a relation.
#pragma once
#include "bpl_common.cuh"

namespace bpl {

// One normalised row of the cubic B-spline basis at time u, and its derivative (splines.py:19-54, 72-93).
__device__ __forceinline__ void basis_row_dev(float u, float (&c)[NCPT], float (&dc)[NCPT], float &csum) {
    float tot = 0.0f, dtot = 0.0f;
#pragma unroll
    for (int k = 0; k < NCPT; ++k) {
        const float d = u - (float)k;
        float b = 0.0f, db = 0.0f;
        if (d >= 0.0f && d < 1.0f) { b = d * d * d; db = 3.0f * d * d; }
        else if (d >= 1.0f && d < 2.0f) { const float e = d - 1.0f; b = ((-3.0f * e * e * e + 3.0f * e * e) + 3.0f * e) + 1.0f; db = (-9.0f * e * e + 6.0f * e) + 3.0f; }
        else if (d >= 2.0f && d < 3.0f) { const float e = d - 2.0f; b = (3.0f * e * e * e - 6.0f * e * e) + 4.0f; db = 9.0f * e * e - 12.0f * e; }
        else if (d >= 3.0f && d < 4.0f) { const float e = 1.0f - (d - 3.0f); b = e * e * e; db = -3.0f * e * e; }
        c[k] = b / 6.0f; dc[k] = db / 6.0f;
        tot += c[k]; dtot += dc[k];
    }
    csum = 0.0f;
#pragma unroll
    for (int k = 0; k < NCPT; ++k) {
        c[k] = c[k] / tot;
        dc[k] = (dc[k] - c[k] * dtot) / tot;
        csum += c[k];
    }
}


// RelationToken.get_attach_point (pybpl/objects/relation.py:34-67) for a 'start' / 'end' / 'mid' relation to stroke j
// of the same token, from that stroke's token-level parameters: the vanilla_to_motor chain (objects/part.py:305-328)
// over its sub-strokes [sj0, sj1) using rows C0 / C1 (first / last) of coefficient_mat(ncpt, neval).
//   start: motor[0][0]    end: motor[-1][-1]    mid: C(u) @ motor_spline[:, :, sub]  (sub = absolute sub-stroke index)
__device__ __forceinline__ void attach_point(int cat, int sj0, int sj1, int sub, float u, const float *ctrl,
                                             const float *invscale, float pjx, float pjy, const float (&C0)[NCPT],
                                             const float (&C1)[NCPT], float &bx, float &by) {
    if (cat == BPL_REL_START) {
        const float inv = invscale[sj0];
        const float *cp = ctrl + (size_t)sj0 * NCPT * 2;
        float ax = 0.0f, ay = 0.0f;
#pragma unroll
        for (int q = 0; q < NCPT; ++q) { ax = fmaf(C0[q], inv * cp[2 * q], ax); ay = fmaf(C0[q], inv * cp[2 * q + 1], ay); }
        bx = ax - (ax - pjx); by = ay - (ay - pjy);
        return;
    }
    const int last = (cat == BPL_REL_MID) ? sub : sj1;
    float ex = pjx, ey = pjy;
    for (int s = sj0; s < last; ++s) {
        const float inv = invscale[s];
        const float *cp = ctrl + (size_t)s * NCPT * 2;
        float ax = 0.0f, ay = 0.0f, zx = 0.0f, zy = 0.0f;
#pragma unroll
        for (int q = 0; q < NCPT; ++q) {
            const float yx = inv * cp[2 * q], yy = inv * cp[2 * q + 1];
            ax = fmaf(C0[q], yx, ax); ay = fmaf(C0[q], yy, ay);
            zx = fmaf(C1[q], yx, zx); zy = fmaf(C1[q], yy, zy);
        }
        ex = zx - (ax - ex); ey = zy - (ay - ey);
    }
    if (cat == BPL_REL_END) { bx = ex; by = ey; return; }
    float cu[NCPT], dcu[NCPT], csum;
    basis_row_dev(u, cu, dcu, csum);
    const float inv = invscale[sub];
    const float *cp = ctrl + (size_t)sub * NCPT * 2;
    float ax = 0.0f, ay = 0.0f;
#pragma unroll
    for (int q = 0; q < NCPT; ++q) { ax = fmaf(C0[q], inv * cp[2 * q], ax); ay = fmaf(C0[q], inv * cp[2 * q + 1], ay); }
    const float ox = ax - ex, oy = ay - ey;
    bx = 0.0f; by = 0.0f;
#pragma unroll
    for (int q = 0; q < NCPT; ++q) { bx = fmaf(cu[q], inv * cp[2 * q] - ox, bx); by = fmaf(cu[q], inv * cp[2 * q + 1] - oy, by); }
}

}  // namespace bpl
