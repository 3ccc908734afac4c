// bpl_splines.cu -- B-spline basis tables (host), batched spline evaluation, vanilla_to_motor,
// image bit-packing and the per-GPU log-sum-exp partial.  Small kernels around the fused
// render+score kernels in bpl_render.cu; all HBM-bound or latency-bound, no tensor cores.
#include <cuda_runtime.h>

#include <atomic>
#include <math.h>
#include <stdlib.h>

#include "bpl_common.cuh"

namespace bpl {
void set_error(const char *fmt, const char *detail);
const char *last_error();
void count_launch(int n);
int sm_count();
long long launch_count();

// ---- host: basis table (pybpl/splines.py:19-93) ----------------------------------------------
// Uniform cubic B-spline weights of control point i at time s, divided by 6, then each row
// normalised by its sum (splines.py:91; matters at the ends where s is in [2,3) or (nland,nland+1]).
static void basis_row(int nland, float s, float *row) {
    float sum = 0.0f;
    for (int i = 0; i < nland; ++i) {
        const float vi = (float)i;
        float c = 0.0f;
        if (s >= vi && s < vi + 1.0f) {
            const float d = s - vi;
            c = d * d * d;
        } else if (s >= vi + 1.0f && s < vi + 2.0f) {
            const float d = s - vi - 1.0f;
            c = ((-3.0f * (d * d * d) + 3.0f * (d * d)) + 3.0f * d) + 1.0f;
        } else if (s >= vi + 2.0f && s < vi + 3.0f) {
            const float d = s - vi - 2.0f;
            c = (3.0f * (d * d * d) - 6.0f * (d * d)) + 4.0f;
        } else if (s >= vi + 3.0f && s < vi + 4.0f) {
            const float e = 1.0f - (s - vi - 3.0f);
            c = e * e * e;
        }
        c = c / 6.0f;
        row[i] = c;
        sum += c;
    }
    for (int i = 0; i < nland; ++i) row[i] = row[i] / sum;
}

// torch.linspace(2, nland+1, neval) (splines.py:66-67): filled from both ends.
static void gen_s(int nland, int neval, float *s) {
    const float lo = 2.0f, hi = (float)(nland + 1);
    if (neval == 1) { s[0] = lo; return; }
    const float step = (hi - lo) / (float)(neval - 1);
    for (int j = 0; j < neval; ++j)
        s[j] = (j < neval / 2) ? lo + step * (float)j : hi - step * (float)(neval - 1 - j);
}

// ---- host: least-squares operator for fit_bspline_to_traj (pybpl/splines.py:141-171) --------------------
// torch.linalg.lstsq(C, X, driver='gels') solves min |C Y - X| by Householder QR.  C depends only on
// (nland, neval, s), never on the data, so the operator P = R^-1 Q^T (n x m) is formed once here in double
// (same Householder sequence, applied to the identity) and the per-trajectory work -- Y = P X and the
// residual -- runs on the GPU.
static int lstsq_operator(int m, int n, const float *C, float *P) {
    if (n > m) return -1;
    double *A = (double *)malloc(sizeof(double) * (size_t)m * n);
    double *Qt = (double *)calloc((size_t)m * m, sizeof(double));      // accumulates Q^T
    if (!A || !Qt) { free(A); free(Qt); return -2; }
    for (size_t i = 0; i < (size_t)m * n; ++i) A[i] = C[i];
    for (int i = 0; i < m; ++i) Qt[(size_t)i * m + i] = 1.0;
    int rc = 0;
    for (int k = 0; k < n; ++k) {
        double norm = 0.0;
        for (int i = k; i < m; ++i) norm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        norm = sqrt(norm);
        if (norm == 0.0) { rc = -3; break; }          // rank deficient
        const double alpha = A[(size_t)k * n + k] > 0 ? -norm : norm;
        const double v0 = A[(size_t)k * n + k] - alpha;
        double vtv = v0 * v0;
        for (int i = k + 1; i < m; ++i) vtv += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        if (vtv == 0.0) { A[(size_t)k * n + k] = alpha; continue; }
        const double beta = 2.0 / vtv;
        for (int j = k + 1; j < n; ++j) {
            double dot = v0 * A[(size_t)k * n + j];
            for (int i = k + 1; i < m; ++i) dot += A[(size_t)i * n + k] * A[(size_t)i * n + j];
            dot *= beta;
            A[(size_t)k * n + j] -= dot * v0;
            for (int i = k + 1; i < m; ++i) A[(size_t)i * n + j] -= dot * A[(size_t)i * n + k];
        }
        for (int j = 0; j < m; ++j) {                 // same reflection on the columns of Q^T
            double dot = v0 * Qt[(size_t)k * m + j];
            for (int i = k + 1; i < m; ++i) dot += A[(size_t)i * n + k] * Qt[(size_t)i * m + j];
            dot *= beta;
            Qt[(size_t)k * m + j] -= dot * v0;
            for (int i = k + 1; i < m; ++i) Qt[(size_t)i * m + j] -= dot * A[(size_t)i * n + k];
        }
        A[(size_t)k * n + k] = alpha;
    }
    if (rc == 0) {
        for (int j = 0; j < m; ++j)                   // back-substitute every column of Q^T[:n]
            for (int k = n - 1; k >= 0; --k) {
                double acc = Qt[(size_t)k * m + j];
                for (int q = k + 1; q < n; ++q) acc -= A[(size_t)k * n + q] * Qt[(size_t)q * m + j];
                acc /= A[(size_t)k * n + k];
                Qt[(size_t)k * m + j] = acc;
                P[(size_t)k * m + j] = (float)acc;
            }
    }
    free(A); free(Qt);
    return rc;
}

// ---- device kernels ------------------------------------------------------------------------------

// X[b][j][:] = sum_k C[j][k] Y[b][k][:], sequential fma chain from 0 (== torch.matmul, App. B.1)
__global__ void spline_eval_kernel(const float *__restrict__ C, int neval, int nland, int nspl,
                                   const float *__restrict__ Y, float *__restrict__ X) {
    const long long total = (long long)nspl * neval;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / neval), j = (int)(i - (long long)b * neval);
        const float *y = Y + (size_t)b * nland * 2;
        const float *row = C + (size_t)j * nland;
        float ax = 0.0f, ay = 0.0f;
        for (int k = 0; k < nland; ++k) {
            const float c = row[k];
            ax = __fmaf_rn(c, y[2 * k], ax);
            ay = __fmaf_rn(c, y[2 * k + 1], ay);
        }
        reinterpret_cast<float2 *>(X)[i] = make_float2(ax, ay);
    }
}

// gY[b][k][:] = sum_j C[j][k] gX[b][j][:]; one warp per (spline, control point)
__global__ void spline_eval_bwd_kernel(const float *__restrict__ C, int neval, int nland, int nspl,
                                       const float *__restrict__ gX, float *__restrict__ gY) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = (long long)nspl * nland;
    for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += nwarps) {
        const int b = (int)(wi / nland), k = (int)(wi - (long long)b * nland);
        const float2 *g = reinterpret_cast<const float2 *>(gX) + (size_t)b * neval;
        float ax = 0.0f, ay = 0.0f;
        for (int j = lane; j < neval; j += 32) {
            const float c = C[(size_t)j * nland + k];
            const float2 v = g[j];
            ax = fmaf(c, v.x, ax);
            ay = fmaf(c, v.y, ay);
        }
        ax = warp_sum(ax); ay = warp_sum(ay);
        if (lane == 0) reinterpret_cast<float2 *>(gY)[wi] = make_float2(ax, ay);
    }
}

// dist_along_traj (util/stroke.py:22-28) over the max_neval-point trajectory, then
// clamp(ceil(dist / grain), min_neval, max_neval) (splines.py:129-133).  One warp per spline;
// the segment lengths are summed in index order by lane 0 so that the integer result does not
// depend on a reduction tree.
__global__ void adaptive_neval_kernel(const float *__restrict__ C, int max_neval, int nland, int nspl,
                                      const float *__restrict__ Y, float grain, int min_neval,
                                      int *__restrict__ neval, float *__restrict__ dist) {
    extern __shared__ float seg[];      // [warps][max_neval]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float *my = seg + (size_t)warp * max_neval;
    for (int b = blockIdx.x * wpb + warp; b < nspl; b += gridDim.x * wpb) {
        const float *y = Y + (size_t)b * nland * 2;
        for (int j = lane; j < max_neval - 1; j += 32) {
            float ax = 0.0f, ay = 0.0f, bx = 0.0f, by = 0.0f;
            for (int k = 0; k < nland; ++k) {
                const float c0 = C[(size_t)j * nland + k], c1 = C[(size_t)(j + 1) * nland + k];
                ax = __fmaf_rn(c0, y[2 * k], ax); ay = __fmaf_rn(c0, y[2 * k + 1], ay);
                bx = __fmaf_rn(c1, y[2 * k], bx); by = __fmaf_rn(c1, y[2 * k + 1], by);
            }
            const float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay);
            my[j] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        }
        __syncwarp();
        if (lane == 0) {
            float tot = 0.0f;
            for (int j = 0; j < max_neval - 1; ++j) tot = __fadd_rn(tot, my[j]);
            float q = ceilf(__fdiv_rn(tot, grain));
            q = fminf(fmaxf(q, (float)min_neval), (float)max_neval);
            neval[b] = (int)q;
            if (dist) dist[b] = tot;
        }
        __syncwarp();
    }
}

// resid[b][c] = sum_j (X[b][j][c] - sum_k C[j][k] Y[b][k][c])^2 : one warp per trajectory
__global__ void spline_residual_kernel(const float *__restrict__ C, int neval, int nland, int nspl,
                                       const float *__restrict__ X, const float *__restrict__ Y,
                                       float *__restrict__ resid) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < nspl; b += gridDim.x * wpb) {
        const float2 *x = reinterpret_cast<const float2 *>(X) + (size_t)b * neval;
        const float2 *y = reinterpret_cast<const float2 *>(Y) + (size_t)b * nland;
        float rx = 0.0f, ry = 0.0f;
        for (int j = lane; j < neval; j += 32) {
            float ax = 0.0f, ay = 0.0f;
            for (int k = 0; k < nland; ++k) {
                const float c = C[(size_t)j * nland + k];
                const float2 v = y[k];
                ax = fmaf(c, v.x, ax); ay = fmaf(c, v.y, ay);
            }
            const float2 t = x[j];
            rx = fmaf(t.x - ax, t.x - ax, rx); ry = fmaf(t.y - ay, t.y - ay, ry);
        }
        rx = warp_sum(rx); ry = warp_sum(ry);
        if (lane == 0) reinterpret_cast<float2 *>(resid)[b] = make_float2(rx, ry);
    }
}

// vanilla_to_motor (objects/part.py:316-326): one warp per stroke, sub-strokes in sequence.
__global__ void vanilla_to_motor_kernel(const float *__restrict__ C, int neval, int ncpt, int n_strokes,
                                        const int *__restrict__ stroke_sub_off, const float *__restrict__ ctrl,
                                        const float *__restrict__ invscale, const float *__restrict__ position,
                                        float *__restrict__ motor, float *__restrict__ mspline) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < n_strokes; k += gridDim.x * wpb) {
        float px = position[2 * k], py = position[2 * k + 1];
        for (int s = stroke_sub_off[k]; s < stroke_sub_off[k + 1]; ++s) {
            const float inv = invscale[s];
            const float *cp = ctrl + (size_t)s * ncpt * 2;
            // first and last points give the offset and the next start (all lanes, redundantly)
            float ax = 0.0f, ay = 0.0f, bx = 0.0f, by = 0.0f;
            for (int q = 0; q < ncpt; ++q) {
                const float yx = __fmul_rn(inv, cp[2 * q]), yy = __fmul_rn(inv, cp[2 * q + 1]);
                const float c0 = C[q], c1 = C[(size_t)(neval - 1) * ncpt + q];
                ax = __fmaf_rn(c0, yx, ax); ay = __fmaf_rn(c0, yy, ay);
                bx = __fmaf_rn(c1, yx, bx); by = __fmaf_rn(c1, yy, by);
            }
            const float ox = __fsub_rn(ax, px), oy = __fsub_rn(ay, py);
            for (int j = lane; j < neval; j += 32) {
                float tx = 0.0f, ty = 0.0f;
                for (int q = 0; q < ncpt; ++q) {
                    const float c = C[(size_t)j * ncpt + q];
                    tx = __fmaf_rn(c, __fmul_rn(inv, cp[2 * q]), tx);
                    ty = __fmaf_rn(c, __fmul_rn(inv, cp[2 * q + 1]), ty);
                }
                reinterpret_cast<float2 *>(motor)[(size_t)s * neval + j] =
                    make_float2(__fsub_rn(tx, ox), __fsub_rn(ty, oy));
            }
            if (mspline)
                for (int q = lane; q < ncpt; q += 32)
                    reinterpret_cast<float2 *>(mspline)[(size_t)s * ncpt + q] =
                        make_float2(__fsub_rn(__fmul_rn(inv, cp[2 * q]), ox), __fsub_rn(__fmul_rn(inv, cp[2 * q + 1]), oy));
            px = __fsub_rn(bx, ox); py = __fsub_rn(by, oy);
        }
    }
}

// One warp packs 32 consecutive pixels (row-major) into one word, LSB first.
template <class T>
__global__ void pack_images_kernel(const T *__restrict__ img, int n_img, uint32_t *__restrict__ bits,
                                   int *__restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long words = (long long)n_img * IMG_WORDS;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words; w += nwarps) {
        const long long t = w / IMG_WORDS;
        const int k = (int)(w - t * IMG_WORDS);
        const int i = k * 32 + lane;
        bool one = false, any_bad = false;
        if (i < NPIX) {
            const T v = img[t * NPIX + i];
            one = (v == (T)1);
            any_bad = !(v == (T)0 || v == (T)1);
        }
        const unsigned m = __ballot_sync(0xffffffffu, one);
        if (lane == 0) bits[w] = m;
        if (any_bad && bad) atomicOr(bad, 1);
    }
}

// out[0] = max ll, out[1] = sum exp(ll - max).  Grid-stride online log-sum-exp per thread ((m, s) stands for
// m + log s; -inf and NaN entries add nothing), combined per warp, per CTA and -- by the CTA that finishes last --
// over the CTAs' partials parked in one slot of a small ring (a slot is all-zero again when its launch completes,
// so launches on different streams never share state).
constexpr int LSE_MAX_CTAS = 256, LSE_RING = 64, LSE_THREADS = 512;
struct LseSlot { unsigned done; float part[2 * LSE_MAX_CTAS]; };
__device__ LseSlot g_lse_ring[LSE_RING];

__device__ __forceinline__ void lse_merge2(float &m, float &s, float m2, float s2) {
    if (!(m2 > -INFINITY) || !(s2 > 0.0f)) return;
    if (m2 > m) { s = fmaf(s, __expf(m - m2), s2); m = m2; }
    else s = fmaf(s2, __expf(m2 - m), s);
}
__device__ __forceinline__ void lse_warp(float &m, float &s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge2(m, s, m2, s2);
    }
}

__global__ void __launch_bounds__(LSE_THREADS) logsumexp_partial_kernel(const float *__restrict__ ll, int n,
                                                                        float *__restrict__ out2, LseSlot *slot) {
    __shared__ float rm[32], rs[32];
    __shared__ int last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float m = -INFINITY, s = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        lse_merge2(m, s, ll[i], 1.0f);
    lse_warp(m, s);
    if (lane == 0) { rm[warp] = m; rs[warp] = s; }
    __syncthreads();
    if (warp == 0) {
        m = lane < nw ? rm[lane] : -INFINITY; s = lane < nw ? rs[lane] : 0.0f;
        lse_warp(m, s);
        if (lane == 0) {
            slot->part[2 * blockIdx.x] = m; slot->part[2 * blockIdx.x + 1] = s;
            __threadfence();
            last = (atomicAdd(&slot->done, 1u) == gridDim.x - 1u);
        }
    }
    __syncthreads();
    if (!last || warp != 0) return;
    __threadfence();
    m = -INFINITY; s = 0.0f;
    for (int i = lane; i < (int)gridDim.x; i += 32) lse_merge2(m, s, __ldcg(slot->part + 2 * i), __ldcg(slot->part + 2 * i + 1));
    lse_warp(m, s);
    if (lane == 0) { out2[0] = m; out2[1] = s; slot->done = 0u; }
}

// out[0] = log sum_i s_i exp(m_i) over n (max, sum-exp) pairs, e.g. the ranks' partials after the all-gather
__global__ void logsumexp_combine_kernel(const float *__restrict__ parts, int n, float *__restrict__ out) {
    const int lane = threadIdx.x;
    float m = -INFINITY, s = 0.0f;
    for (int i = lane; i < n; i += 32) lse_merge2(m, s, parts[2 * i], parts[2 * i + 1]);
    lse_warp(m, s);
    if (lane == 0) out[0] = (m > -INFINITY && s > 0.0f) ? (float)((double)m + log((double)s)) : -INFINITY;
}

// Live FP32 peak probe for the roofline denominator (MEASURED_PEAKS.json has HBM and bf16-tensor
// figures but no CUDA-core FP32 one): 8 independent FFMA chains per thread, `iters` rounds.
__global__ void fp32_probe_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;      // never true in practice; keeps the chains alive
}

static int check_launch(const char *what) {
    count_launch(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s", cudaGetErrorString(e)); (void)what; return (int)e; }
    return 0;
}

static int grid_for(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    const long long cap = (long long)sm_count() * 16;
    if (cap > 0 && g > cap) g = cap;
    return (int)g;
}

}  // namespace bpl

using namespace bpl;

// Adam over up to BPL_ADAM_MAX_TENSORS small parameter tensors in ONE launch (torch.optim.Adam's update rule: no weight
// decay, no amsgrad; `maximize` flips the gradient's sign), followed by the clamp to each tensor's bounds -- the
// optimiser step of the differentiable fit (fit_image, pybpl/model/model.py:42-65 with the bounds of
// parameters.py:54-63).  torch's fused multi-tensor Adam takes 66 us on these five tensors (one block per tensor
// chunk); this is 3 us.  The step count lives on the device (state[0]) so that a captured CUDA graph advances it: the
// block that finishes last increments it.
__global__ void __launch_bounds__(256) adam_step_kernel(const bpl_adam_args a) {
    const float t = a.state[0] + 1.0f;
    const float bias1 = 1.0f - powf(a.beta1, t), bias2 = 1.0f - powf(a.beta2, t);
    const float step_size = a.lr / bias1, rs2 = rsqrtf(bias2);
    long long base = 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (int k = 0; k < a.n_tensors; ++k) {
        const long long n = a.n[k];
        float *p = a.param[k], *m = a.exp_avg[k], *v = a.exp_avg_sq[k];
        const float *g = a.grad[k];
        const float lo = a.lo[k], hi = a.hi[k];
        // a rotating start so that consecutive small tensors land on different threads
        for (long long i = (tid + nth - base % nth) % nth; i < n; i += nth) {
            const float gi = a.maximize ? -g[i] : g[i];
            const float mi = fmaf(a.beta1, m[i], (1.0f - a.beta1) * gi);
            const float vi = fmaf(a.beta2, v[i], (1.0f - a.beta2) * gi * gi);
            m[i] = mi; v[i] = vi;
            const float denom = fmaf(sqrtf(vi), rs2, a.eps);
            p[i] = fminf(fmaxf(p[i] - step_size * (mi / denom), lo), hi);
        }
        base += n;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *done = reinterpret_cast<unsigned *>(a.state + 1);
        if (atomicAdd(done, 1u) == gridDim.x - 1u) { *done = 0u; a.state[0] = t; }
    }
}

// Shared-memory bandwidth probe (the second roofline SURVEY.md 8(d) names): every thread streams conflict-free 128-bit
// loads from a 32 KB shared array, four independent accumulators; bytes moved = blocks * threads * iters * 4 loads * 16 B.
__global__ void __launch_bounds__(1024) smem_probe_kernel(float *scratch, int iters) {
    __shared__ float4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_float4((float)i, 1.0f, 2.0f, 3.0f);
    __syncthreads();
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int j = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const float4 v0 = buf[j & 2047], v1 = buf[(j + 512) & 2047], v2 = buf[(j + 1024) & 2047], v3 = buf[(j + 1536) & 2047];
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        j += 32;
    }
    const float r = a0.x + a0.y + a0.z + a0.w + a1.x + a1.y + a1.z + a1.w + a2.x + a2.y + a2.z + a2.w + a3.x + a3.y + a3.z + a3.w;
    if (r == -1.0f) scratch[0] = r;      // never true: keeps the loads alive
}

// Autograd scaling of the fused backward's outputs: every gradient of token p times the incoming cotangent g_ll[p]
// (the fused kernel runs in the autograd forward, before the cotangent exists).  One warp per token walks the token's
// strokes and sub-strokes through the CSR offsets: ONE launch instead of three gathers and six multiplies.
__global__ void __launch_bounds__(128) scale_token_grads_kernel(int P, const int *__restrict__ tso, const int *__restrict__ sso,
                                                                const float *__restrict__ g_ll, const bpl_grads in, const bpl_grads out) {
    const int lane = threadIdx.x & 31;
    for (int p = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); p < P; p += (int)((gridDim.x * blockDim.x) >> 5)) {
        const float g = g_ll[p];
        const int k0 = tso[p], k1 = tso[p + 1];
        const int s0 = sso[k0], s1 = sso[k1];
        for (int i = s0 * 10 + lane; i < s1 * 10; i += 32) out.g_ctrl[i] = in.g_ctrl[i] * g;
        for (int i = s0 + lane; i < s1; i += 32) out.g_invscale[i] = in.g_invscale[i] * g;
        for (int i = k0 * 2 + lane; i < k1 * 2; i += 32) out.g_position[i] = in.g_position[i] * g;
        if (lane < 4 && in.g_affine) out.g_affine[4 * (size_t)p + lane] = in.g_affine[4 * (size_t)p + lane] * g;
        if (lane == 4) out.g_eps[p] = in.g_eps[p] * g;
        if (lane == 5) out.g_sigma[p] = in.g_sigma[p] * g;
    }
}

extern "C" {

int bpl_version(void) { return 1; }
const char *bpl_last_error(void) { return last_error(); }
long long bpl_launch_count(void) { return launch_count(); }

void bpl_default_params(bpl_params *ps) {      // pybpl/parameters.py:19-41
    ps->imsize = BPL_IMSIZE; ps->fsize = BPL_FSIZE; ps->ink_ncon = 2;
    ps->ink_pp = 2.0f; ps->ink_max_dist = 2.0f; ps->ink_a = 0.5f; ps->ink_b = 6.0f; ps->hinton = 0;
}

int bpl_device_info(int *sm_count, int *smem_optin) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("%s", cudaGetErrorString(e)); return BPL_ENODEV; }
    if (sm_count) cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (smem_optin) cudaDeviceGetAttribute(smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return 0;
}

int bpl_basis_rows_host(int nland, int n, const float *s, float *C) {
    if (nland < 1 || n < 0 || !s || !C) { set_error("%s", "bad basis_rows arguments"); return BPL_EINVAL; }
    for (int j = 0; j < n; ++j) basis_row(nland, s[j], C + (size_t)j * nland);
    return 0;
}

int bpl_basis_table_host(int nland, int neval, float *C) {
    if (nland < 1 || neval < 1 || !C) { set_error("%s", "bad basis_table arguments"); return BPL_EINVAL; }
    float *s = (float *)malloc(sizeof(float) * (size_t)neval);
    if (!s) { set_error("%s", "out of host memory"); return BPL_EINVAL; }
    gen_s(nland, neval, s);
    for (int j = 0; j < neval; ++j) basis_row(nland, s[j], C + (size_t)j * nland);
    free(s);
    return 0;
}

int bpl_spline_eval(const float *C, int neval, int nland, int nspl, const float *Y, float *X, void *stream) {
    if (!C || !Y || !X || neval < 1 || nland < 1 || nspl < 0) { set_error("%s", "bad spline_eval arguments"); return BPL_EINVAL; }
    if (nspl == 0) return 0;
    spline_eval_kernel<<<grid_for((long long)nspl * neval, 256), 256, 0, (cudaStream_t)stream>>>(C, neval, nland, nspl, Y, X);
    return check_launch("spline_eval");
}

int bpl_spline_eval_bwd(const float *C, int neval, int nland, int nspl, const float *gX, float *gY, void *stream) {
    if (!C || !gX || !gY || neval < 1 || nland < 1 || nspl < 0) { set_error("%s", "bad spline_eval_bwd arguments"); return BPL_EINVAL; }
    if (nspl == 0) return 0;
    spline_eval_bwd_kernel<<<grid_for((long long)nspl * nland * 32, 256), 256, 0, (cudaStream_t)stream>>>(C, neval, nland, nspl, gX, gY);
    return check_launch("spline_eval_bwd");
}

int bpl_spline_adaptive_neval(const float *C200, int max_neval, int nland, int nspl, const float *Y, float grain,
                              int min_neval, int *neval, float *dist, void *stream) {
    if (!C200 || !Y || !neval || max_neval < 2 || max_neval > 2048 || nland < 1 || nspl < 0 || !(grain > 0.0f)) {
        set_error("%s", "bad adaptive_neval arguments"); return BPL_EINVAL;
    }
    if (nspl == 0) return 0;
    const int wpb = 4;
    adaptive_neval_kernel<<<grid_for(nspl, wpb), wpb * 32, sizeof(float) * wpb * max_neval, (cudaStream_t)stream>>>(
        C200, max_neval, nland, nspl, Y, grain, min_neval, neval, dist);
    return check_launch("adaptive_neval");
}

int bpl_lstsq_operator_host(int neval, int nland, const float *C, float *P) {
    if (neval < 1 || nland < 1 || !C || !P) { set_error("%s", "bad lstsq_operator arguments"); return BPL_EINVAL; }
    const int r = lstsq_operator(neval, nland, C, P);
    if (r == -1) { set_error("%s", "fit_bspline_to_traj needs neval >= nland"); return BPL_EINVAL; }
    if (r == -2) { set_error("%s", "out of host memory"); return BPL_EINVAL; }
    if (r == -3) { set_error("%s", "coefficient matrix is rank deficient"); return BPL_EINVAL; }
    return 0;
}

int bpl_spline_residual(const float *C, int neval, int nland, int nspl, const float *X, const float *Y, float *resid,
                        void *stream) {
    if (!C || !X || !Y || !resid || neval < 1 || nland < 1 || nspl < 0) { set_error("%s", "bad spline_residual arguments"); return BPL_EINVAL; }
    if (nspl == 0) return 0;
    spline_residual_kernel<<<grid_for(nspl, 4), 128, 0, (cudaStream_t)stream>>>(C, neval, nland, nspl, X, Y, resid);
    return check_launch("spline_residual");
}

int bpl_vanilla_to_motor(const float *C, int neval, int ncpt, int n_strokes, const int *stroke_sub_off,
                         const float *ctrl, const float *invscale, const float *position, float *motor,
                         float *motor_spline, void *stream) {
    if (!C || !stroke_sub_off || !ctrl || !invscale || !position || !motor || neval < 1 || ncpt < 1 || n_strokes < 0) {
        set_error("%s", "bad vanilla_to_motor arguments"); return BPL_EINVAL;
    }
    if (n_strokes == 0) return 0;
    vanilla_to_motor_kernel<<<grid_for(n_strokes, 4), 128, 0, (cudaStream_t)stream>>>(
        C, neval, ncpt, n_strokes, stroke_sub_off, ctrl, invscale, position, motor, motor_spline);
    return check_launch("vanilla_to_motor");
}

int bpl_pack_images_u8(const uint8_t *img, int T, uint32_t *bits, int *bad, void *stream) {
    if (!img || !bits || T < 0) { set_error("%s", "bad pack_images arguments"); return BPL_EINVAL; }
    if (T == 0) return 0;
    pack_images_kernel<uint8_t><<<grid_for((long long)T * IMG_WORDS * 32, 256), 256, 0, (cudaStream_t)stream>>>(img, T, bits, bad);
    return check_launch("pack_images_u8");
}

int bpl_pack_images_f32(const float *img, int T, uint32_t *bits, int *bad, void *stream) {
    if (!img || !bits || T < 0) { set_error("%s", "bad pack_images arguments"); return BPL_EINVAL; }
    if (T == 0) return 0;
    pack_images_kernel<float><<<grid_for((long long)T * IMG_WORDS * 32, 256), 256, 0, (cudaStream_t)stream>>>(img, T, bits, bad);
    return check_launch("pack_images_f32");
}

int bpl_fp32_probe(float *scratch, int iters, int blocks, void *stream) {
    if (!scratch || iters < 1 || blocks < 1) { set_error("%s", "bad fp32_probe arguments"); return BPL_EINVAL; }
    fp32_probe_kernel<<<blocks, 1024, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 0.001f);
    return check_launch("fp32_probe");
}

int bpl_smem_probe(float *scratch, int iters, int blocks, void *stream) {
    if (!scratch || iters < 1 || blocks < 1) { set_error("%s", "bad smem_probe arguments"); return BPL_EINVAL; }
    smem_probe_kernel<<<blocks, 1024, 0, (cudaStream_t)stream>>>(scratch, iters);
    return check_launch("smem_probe");
}

int bpl_logsumexp_partial(const float *ll, int n, float *out2, void *stream) {
    if (!ll || !out2 || n < 0) { set_error("%s", "bad logsumexp arguments"); return BPL_EINVAL; }
    static std::atomic<unsigned> next_slot{0};
    static thread_local int ring_dev = -1;
    static thread_local LseSlot *ring = nullptr;      // one instance of g_lse_ring per device
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != ring_dev) {
        const cudaError_t es = cudaGetSymbolAddress(reinterpret_cast<void **>(&ring), g_lse_ring);
        if (es != cudaSuccess) { set_error("cudaGetSymbolAddress: %s", cudaGetErrorString(es)); return (int)es; }
        ring_dev = dev;
    }
    int grid = (n + LSE_THREADS * 8 - 1) / (LSE_THREADS * 8);
    const int cap = sm_count() * 1 < LSE_MAX_CTAS ? sm_count() : LSE_MAX_CTAS;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    logsumexp_partial_kernel<<<grid, LSE_THREADS, 0, (cudaStream_t)stream>>>(ll, n, out2, ring + (next_slot.fetch_add(1u) % LSE_RING));
    return check_launch("logsumexp_partial");
}

int bpl_logsumexp_combine(const float *parts, int n, float *out, void *stream) {
    if (!parts || !out || n < 0) { set_error("%s", "bad logsumexp_combine arguments"); return BPL_EINVAL; }
    logsumexp_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(parts, n, out);
    return check_launch("logsumexp_combine");
}

int bpl_adam_step(const bpl_adam_args *a, void *stream) {
    if (!a || !a->state || a->n_tensors < 0 || a->n_tensors > BPL_ADAM_MAX_TENSORS) { set_error("%s", "bad bpl_adam_step arguments"); return BPL_EINVAL; }
    long long total = 0;
    for (int k = 0; k < a->n_tensors; ++k) {
        if (a->n[k] < 0 || (a->n[k] > 0 && (!a->param[k] || !a->grad[k] || !a->exp_avg[k] || !a->exp_avg_sq[k]))) {
            set_error("%s", "bpl_adam_step: NULL tensor"); return BPL_EINVAL;
        }
        total += a->n[k];
    }
    if (total == 0) return 0;
    const long long cap = (long long)sm_count() * 8;
    long long blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    adam_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    return check_launch("adam_step");
}

int bpl_scale_token_grads(int P, int K, int S, const int *tok_stroke_off, const int *stroke_sub_off, const float *g_ll,
                          const bpl_grads *in, const bpl_grads *out, void *stream) {
    if (!tok_stroke_off || !stroke_sub_off || !g_ll || !in || !out) { set_error("%s", "bpl_scale_token_grads: NULL argument"); return BPL_EINVAL; }
    if (P <= 0) return 0;
    const int cap = sm_count() * 8;
    const int blocks = (P + 3) / 4 < cap ? (P + 3) / 4 : cap;       // one warp per token, four per block
    scale_token_grads_kernel<<<blocks > 0 ? blocks : 1, 128, 0, (cudaStream_t)stream>>>(P, tok_stroke_off, stroke_sub_off, g_ll, *in, *out);
    (void)K; (void)S;
    return check_launch("scale_token_grads");
}

}  // extern "C"
