// Device-side building blocks shared by the forward (dsac.cu) and the backward pass (dsac_backward.cu) of the pose
// solver: block / cluster reductions and the restatement of cv::solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess) that
// refineHyp drives (/root/reference/dsacstar/dsacstar_util.h:522-597).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cfloat>

#include "dsac_common.cuh"

namespace cl {
namespace {

constexpr int kRefineThreads = 256;   // per CTA; an image is refined by a cluster of 1..8 CTAs
constexpr int kMaxRefSteps = 100;     // dsacstar.cpp:47

__device__ __forceinline__ void cell_pixel(int x, int y, int S, int& px, int& py)
{
    px = x * S + S / 2;   // dsacstar_util.h:70-72
    py = y * S + S / 2;
}


template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double* smem /* [THREADS/32 * NV + NV] */)
{
    constexpr int kWarps = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; j++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    }
    __syncthreads();   // protects smem reuse across consecutive reductions
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NV; j++) smem[warp * NV + j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < kWarps; w++) s += smem[w * NV + threadIdx.x];
        smem[kWarps * NV + threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = smem[kWarps * NV + j];
}

// Sum over every thread of the cluster working on one image.  Warp shuffles, one shared-memory slot per warp, then
// every CTA publishes its NV partial sums and reads those of its peers through distributed shared memory in rank
// order, so all CTAs (and all threads) end up with bit-identical totals and take the same control flow.  `part` is
// double-buffered by the caller-held parity: one cluster barrier per reduction suffices (a CTA can only overwrite
// buffer p after every peer has passed the barrier of the reduction in between, i.e. finished reading p).
struct ClusterRed {
    double* warp_part;   // [THREADS / 32][28]
    double* part;        // [2][28], read by the peers
    double* total;       // [28]
    unsigned parity;
};

template <int NV>
__device__ __forceinline__ void cluster_reduce_sum(double (&v)[NV], ClusterRed& cr)
{
    namespace cg = cooperative_groups;
    constexpr int kWarps = kRefineThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; j++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NV; j++) cr.warp_part[warp * 28 + j] = v[j];
    }
    __syncthreads();
    double* mine = cr.part + (cr.parity & 1u) * 28;
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < kWarps; w++) s += cr.warp_part[w * 28 + threadIdx.x];
        mine[threadIdx.x] = s;
    }
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nranks = cluster.num_blocks();
    if (nranks > 1) {
        cluster.sync();
        if (threadIdx.x < NV) {
            double s = 0;
            for (unsigned r = 0; r < nranks; r++) s += cluster.map_shared_rank(mine, r)[threadIdx.x];
            cr.total[threadIdx.x] = s;
        }
    } else {
        __syncthreads();
        if (threadIdx.x < NV) cr.total[threadIdx.x] = mine[threadIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = cr.total[j];
    cr.parity++;
}

struct LmSums {
    double JtJ[36];
    double Jte[6];
    double err;
};

// Residuals (and, if want_j, the normal equations) of the pixel reprojection error over the current
// inlier set { i : errs[i] < thr }, reduced over the block.  Every thread returns the same sums.
template <bool WANT_J>
__device__ void lm_accumulate(const double prm[6], const float* X, const float* errs, int n, int Wc, int S, float thr,
                              double f, double cx, double cy, ClusterRed& smem, LmSums& out, int first, int stride)
{
    double R[9], M[9], RM[9];
    rodrigues(prm, R);
    if (WANT_J) {
        rotation_jacobian_factor(prm, R, M);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) RM[3 * i + j] = R[3 * i] * M[j] + R[3 * i + 1] * M[3 + j] + R[3 * i + 2] * M[6 + j];
    }
    constexpr int NV = WANT_J ? 28 : 1;
    double acc[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) acc[j] = 0;
    for (int i = first; i < n; i += stride) {
        if (!(errs[i] < thr)) continue;   // strict <, dsacstar_util.h:550
        const int yy = i / Wc, xx = i - yy * Wc;
        int px, py;
        cell_pixel(xx, yy, S, px, py);
        const double Xw = X[i], Yw = X[n + i], Zw = X[2 * n + i];
        const double qx = R[0] * Xw + R[1] * Yw + R[2] * Zw;
        const double qy = R[3] * Xw + R[4] * Yw + R[5] * Zw;
        const double qz = R[6] * Xw + R[7] * Yw + R[8] * Zw;
        const double x = qx + prm[3], y = qy + prm[4], z = qz + prm[5];
        const double iz = z ? 1. / z : 1;
        const double eu = f * x * iz + cx - (double)px, ev = f * y * iz + cy - (double)py;
        acc[NV - 1] += eu * eu + ev * ev;
        if (WANT_J) {
            const double a0 = f * iz, a2 = -f * x * iz * iz, b2 = -f * y * iz * iz;
            // dp/dr = -[q]x (R M)
            double dpdr[9];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                dpdr[j] = -(-qz * RM[3 + j] + qy * RM[6 + j]);
                dpdr[3 + j] = -(qz * RM[j] - qx * RM[6 + j]);
                dpdr[6 + j] = -(-qy * RM[j] + qx * RM[3 + j]);
            }
            double Ju[6], Jv[6];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Ju[j] = a0 * dpdr[j] + a2 * dpdr[6 + j];
                Jv[j] = a0 * dpdr[3 + j] + b2 * dpdr[6 + j];
            }
            Ju[3] = a0; Ju[4] = 0; Ju[5] = a2;
            Jv[3] = 0; Jv[4] = a0; Jv[5] = b2;
            int k = 0;
#pragma unroll
            for (int r2 = 0; r2 < 6; r2++)
#pragma unroll
                for (int c = r2; c < 6; c++) acc[k++] += Ju[r2] * Ju[c] + Jv[r2] * Jv[c];
#pragma unroll
            for (int r2 = 0; r2 < 6; r2++) acc[21 + r2] += Ju[r2] * eu + Jv[r2] * ev;
        }
    }
    cluster_reduce_sum<NV>(acc, smem);
    out.err = sqrt(acc[NV - 1]);
    if (WANT_J) {
        int k = 0;
#pragma unroll
        for (int r2 = 0; r2 < 6; r2++)
#pragma unroll
            for (int c = r2; c < 6; c++) { out.JtJ[6 * r2 + c] = acc[k]; out.JtJ[6 * c + r2] = acc[k]; k++; }
#pragma unroll
        for (int r2 = 0; r2 < 6; r2++) out.Jte[r2] = acc[21 + r2];
    }
}

__device__ bool lm_step(const LmSums& s, int lambdaLg10, const double prev[6], double prm[6])
{
    double A[36], b[6], x[6];
    const double lambda = exp(lambdaLg10 * log(10.));
    for (int i = 0; i < 36; i++) A[i] = s.JtJ[i];
    for (int i = 0; i < 6; i++) { b[i] = s.Jte[i]; A[7 * i] *= 1. + lambda; }
    if (!solve6(A, b, x)) return false;
    for (int i = 0; i < 6; i++) prm[i] = prev[i] - x[i];
    return true;
}

// cv::solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess = true): CvLevMarq::update as driven by
// cvFindExtrinsicCameraParams2 -- at most 20 iterations, eps = FLT_EPSILON on the relative parameter
// change, lambda = 10^k from k = -3, k+1 on a worse step (at most 16), k-1 on an accepted one.
__device__ bool lm_solve(double prm[6], const float* X, const float* errs, int n, int Wc, int S, float thr, double f,
                         double cx, double cy, ClusterRed& smem, int first, int stride)
{
    LmSums s;
    double prev[6];
    int lambdaLg10 = -3, iters = 0;
    double prevErr = DBL_MAX, errNorm;
    for (;;) {
        lm_accumulate<true>(prm, X, errs, n, Wc, S, thr, f, cx, cy, smem, s, first, stride);
        for (int i = 0; i < 6; i++) prev[i] = prm[i];
        if (!lm_step(s, lambdaLg10, prev, prm)) return false;
        if (iters == 0) prevErr = s.err;
        for (;;) {
            LmSums e;
            lm_accumulate<false>(prm, X, errs, n, Wc, S, thr, f, cx, cy, smem, e, first, stride);
            errNorm = e.err;
            if (errNorm > prevErr && ++lambdaLg10 <= 16) {
                if (!lm_step(s, lambdaLg10, prev, prm)) return false;
                continue;
            }
            break;
        }
        lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
        double dn = 0, pn = 0;
        for (int i = 0; i < 6; i++) { dn += (prm[i] - prev[i]) * (prm[i] - prev[i]); pn += prev[i] * prev[i]; }
        if (++iters >= 20 || sqrt(dn) / (sqrt(pn) + DBL_EPSILON) < FLT_EPSILON) break;
        prevErr = errNorm;
    }
    return true;
}

// Error map of one pose into errs[] + inlier count (block-uniform return value).
__device__ int error_map(const double prm[6], const float* X, float* errs, int n, int Wc, int S, float f, float cx,
                         float cy, float thr, float max_reproj, ClusterRed& smem, int first, int stride)
{
    double R[9];
    rodrigues(prm, R);
    double cnt[1] = {0};
    for (int i = first; i < n; i += stride) {
        const int y = i / Wc, x = i - y * Wc;
        int px, py;
        cell_pixel(x, y, S, px, py);
        const float e = repro_error(R, prm + 3, f, cx, cy, X[i], X[n + i], X[2 * n + i], px, py, max_reproj);
        errs[i] = e;
        if (e < thr) cnt[0] += 1;
    }
    cluster_reduce_sum<1>(cnt, smem);
    return (int)cnt[0];
}

}  // namespace
}  // namespace cl
