// DSAC* RGB forward on the GPU (sm_100a): hypothesis sampling + P3P, soft-inlier scoring,
// selection and iterative refinement, batched over images.
//
// Replaces /root/reference/dsacstar/dsacstar.cpp:63-178 (dsacstar_rgb_forward) and the helpers in
// /root/reference/dsacstar/dsacstar_util.h that it calls:
//   sampleHypotheses :135-221 -> dsac_sample_kernel   one warp per hypothesis, lane = try index
//   getReproErrs     :356-446 \  dsac_score_kernel    one block per (image, hypothesis); the [3,Hc*Wc]
//   getHypScores     :316-343 /                       map is streamed with coalesced fp32 loads
//   softMax/draw     :684-752 \  dsac_refine_kernel   one block per image
//   refineHyp        :522-597 |
//   pose2trans       :759-770 /
// The per-hypothesis error maps the reference materialises (dsacstar.cpp:121) are never written:
// only the winner's map is recomputed for the refinement.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "dsac.h"
#include "dsac_common.cuh"
#include "dsac_refine.cuh"

namespace cl {

namespace {

constexpr int kSampleWarps = 4;
constexpr int kScoreThreads = 256;
// Register cap of the refinement kernel: 256 threads x 136 registers fit on an SM NEXT TO a convolution CTA (256 threads
// x 112 registers, ~200 KB of shared memory), so a solve running on its own stream shares SMs with the tensor-core-bound
// convolutions of the following batch instead of waiting for whole SMs (an uncapped build needs 232 registers).
constexpr int kRefineMaxRegs = 136;
constexpr double kProbEps = 1e-8;   // dsacstar_util.h:45

// ---------------------------------------------------------------------------------------------
// Sampling: lane l of the warp evaluates try (round * 32 + l); the lowest accepted try wins, which is
// what the reference's sequential retry loop (dsacstar_util.h:159-220) returns for the same draws.
__global__ void __launch_bounds__(kSampleWarps * 32) dsac_sample_kernel(DsacArgs a)
{
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (h >= a.hyps) return;
    const int n = a.Hc * a.Wc;
    const float* X = a.coords + (size_t)b * 3 * n;
    const double f = a.focal[b], cx = a.cx, cy = a.cy;
    const double thr = a.thr;
    const uint32_t image = a.image_base + (uint32_t)b;
    const bool forced = a.forced != nullptr;
    const uint32_t max_tries = forced ? 1u : a.max_tries;

    Pose win;
    uint32_t tries_used = max_tries;
    for (uint32_t base = 0; base < max_tries; base += 32) {
        const uint32_t t = base + lane;
        Pose cand;
#pragma unroll
        for (int j = 0; j < 3; j++) { cand.r[j] = 0; cand.t[j] = 0; }
        bool accept = false;
        int cells[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (t < max_tries) {
            if (forced) {
#pragma unroll
                for (int j = 0; j < 8; j++) cells[j] = a.forced[((size_t)b * a.hyps + h) * 8 + j];
            } else {
                sample_cells(a.seed, image, (uint32_t)h, t, a.Wc, a.Hc, cells);
            }
            double obj[12], img[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int x = cells[2 * j], y = cells[2 * j + 1], i = y * a.Wc + x;
                int px, py;
                cell_pixel(x, y, a.S, px, py);
                img[2 * j] = px; img[2 * j + 1] = py;
                obj[3 * j] = X[i]; obj[3 * j + 1] = X[n + i]; obj[3 * j + 2] = X[2 * n + i];
            }
            if (p3p_solve(obj, img, f, cx, cy, cand)) {
                double R[9];
                rodrigues(cand.r, R);
                accept = true;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double u, v;
                    project_point(R, cand.t, f, cx, cy, obj[3 * j], obj[3 * j + 1], obj[3 * j + 2], u, v);
                    const float du = (float)img[2 * j] - (float)u, dv = (float)img[2 * j + 1] - (float)v;
                    if (!(sqrt((double)du * du + (double)dv * dv) < thr)) accept = false;   // strict <, :210
                }
            } else {
#pragma unroll
                for (int j = 0; j < 3; j++) { cand.r[j] = 0; cand.t[j] = 0; }   // safeSolvePnP, :114-116
            }
            if (forced) accept = true;   // replay mode: the injected sample is the hypothesis
        }
        const unsigned m = __ballot_sync(0xffffffffu, accept);
        const bool last_round = base + 32 >= max_tries;
        if (m || last_round) {
            // no accepted try within max_tries: the reference is left with its last try's result
            const int src = m ? __ffs(m) - 1 : (int)(max_tries - 1 - base);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                win.r[j] = __shfl_sync(0xffffffffu, cand.r[j], src);
                win.t[j] = __shfl_sync(0xffffffffu, cand.t[j], src);
            }
            tries_used = base + (uint32_t)src + 1;
            if (a.out_cells) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int c = __shfl_sync(0xffffffffu, cells[j], src);
                    if (lane == 0) a.out_cells[((size_t)b * a.hyps + h) * 8 + j] = c;
                }
            }
            break;
        }
    }
    if (lane == 0) {
        double* o = a.hyp_rt + ((size_t)b * a.hyps + h) * 6;
#pragma unroll
        for (int j = 0; j < 3; j++) { o[j] = win.r[j]; o[3 + j] = win.t[j]; }
        if (a.tries) a.tries[(size_t)b * a.hyps + h] = (int32_t)tries_used;
    }
}

// ---------------------------------------------------------------------------------------------
// Scoring: block (h, b) streams the image's planar X, Y, Z once (12 B per cell, coalesced), projects
// in double like cv::projectPoints, and accumulates 1 - sigmoid(beta (err - thr)) in double.
__global__ void __launch_bounds__(kScoreThreads) dsac_score_kernel(DsacArgs a)
{
    __shared__ double red[(kScoreThreads / 32 + 1) * 1];
    const int h = blockIdx.x, b = blockIdx.y;
    const int n = a.Hc * a.Wc;
    const float* X = a.coords + (size_t)b * 3 * n;
    const double* rt = a.hyp_rt + ((size_t)b * a.hyps + h) * 6;
    const double r[3] = {rt[0], rt[1], rt[2]}, t[3] = {rt[3], rt[4], rt[5]};
    double R[9];
    rodrigues(r, R);
    const float f = a.focal[b];
    const float beta = 5 / a.thr;   // dsacstar_util.h:324
    double acc[1] = {0};
    for (int i = threadIdx.x; i < n; i += kScoreThreads) {
        const int y = i / a.Wc, x = i - y * a.Wc;
        int px, py;
        cell_pixel(x, y, a.S, px, py);
        const float e = repro_error(R, t, f, a.cx, a.cy, X[i], X[n + i], X[2 * n + i], px, py, a.max_reproj);
        double soft = beta * (e - a.thr);   // float arithmetic widened to double, :331
        soft = 1 / (1 + exp(-soft));
        acc[0] += 1 - soft;
    }
    block_reduce_sum<1, kScoreThreads>(acc, red);
    if (threadIdx.x == 0) a.scores[(size_t)b * a.hyps + h] = acc[0] * (double)(a.alpha / a.Wc / a.Hc);   // :339
}

// ---------------------------------------------------------------------------------------------
// One CLUSTER of CTAs per image (cluster size chosen by the launcher from the map size: the fp64 normal equations
// make a pass over the inliers throughput-bound on one SM's fp64 pipe -- 0.78 ms per image at 60 x 90 cells, 58 of
// 81 ms per 32 frames at 480 x 720 with one 512-thread block).  Every thread owns the cells first, first + stride,
// ... for the whole kernel, so the error map it reads back is always the one it wrote itself.
__global__ void __maxnreg__(kRefineMaxRegs) dsac_refine_kernel(DsacArgs a)
{
    namespace cg = cooperative_groups;
    __shared__ double s_warp_part[(kRefineThreads / 32) * 28];
    __shared__ double s_part[2 * 28];
    __shared__ double s_total[28];
    __shared__ int s_best;
    cg::cluster_group cluster = cg::this_cluster();
    const int nranks = (int)cluster.num_blocks();
    const int b = (int)blockIdx.x / nranks;
    const int first = (int)cluster.block_rank() * kRefineThreads + (int)threadIdx.x;
    const int stride = nranks * kRefineThreads;
    ClusterRed red{s_warp_part, s_part, s_total, 0u};
    const int n = a.Hc * a.Wc;
    const float* X = a.coords + (size_t)b * 3 * n;
    float* errs = a.errs + (size_t)b * n;
    const double* scores = a.scores + (size_t)b * a.hyps;

    // softMax + draw(probs, false): first maximal probability among those >= 1e-8 (dsacstar_util.h:684-752)
    if (threadIdx.x == 0) {
        double mx = scores[0], sum = 0;
        for (int h = 1; h < a.hyps; h++) if (scores[h] > mx) mx = scores[h];
        for (int h = 0; h < a.hyps; h++) sum += exp(scores[h] - mx);
        int best = 0;
        double bestp = -1;
        for (int h = 0; h < a.hyps; h++) {
            const double p = exp(scores[h] - mx) / sum;
            if (p < kProbEps) continue;
            if (bestp < 0 || p > bestp) { bestp = p; best = h; }
        }
        s_best = best;
        if (a.out_best && first == 0) a.out_best[b] = best;
    }
    __syncthreads();
    const int best = s_best;

    double prm[6];
#pragma unroll
    for (int j = 0; j < 6; j++) prm[j] = a.hyp_rt[((size_t)b * a.hyps + best) * 6 + j];
    const float f = a.focal[b];

    if (a.refine) {
        // refineHyp (dsacstar_util.h:522-597): refit to all inliers while the inlier count grows
        int inliers = error_map(prm, X, errs, n, a.Wc, a.S, f, a.cx, a.cy, a.thr, a.max_reproj, red, first, stride);
        int best_inliers = 4;
        for (int step = 0; step < kMaxRefSteps; step++) {
            if (a.out_counts && first == 0) a.out_counts[(size_t)b * kMaxRefSteps + step] = inliers;
            if (inliers <= best_inliers) break;
            best_inliers = inliers;
            double upd[6];
#pragma unroll
            for (int j = 0; j < 6; j++) upd[j] = prm[j];
            if (!lm_solve(upd, X, errs, n, a.Wc, a.S, a.thr, (double)f, (double)a.cx, (double)a.cy, red, first, stride)) break;
#pragma unroll
            for (int j = 0; j < 6; j++) prm[j] = upd[j];
            inliers = error_map(prm, X, errs, n, a.Wc, a.S, f, a.cx, a.cy, a.thr, a.max_reproj, red, first, stride);
        }
    }

    if (nranks > 1) cluster.sync();   // no CTA retires while a peer may still read its partial sums
    if (first == 0) {
        // pose2trans (dsacstar_util.h:759-770): inverse of [R t; 0 1], row-major float (dsacstar.cpp:174-177)
        double R[9];
        rodrigues(prm, R);
        float* o = a.out_pose + (size_t)b * 16;
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) o[4 * i + j] = (float)R[3 * j + i];
            o[4 * i + 3] = (float)-(R[i] * prm[3] + R[3 + i] * prm[4] + R[6 + i] * prm[5]);
        }
        o[12] = o[13] = o[14] = 0;
        o[15] = 1;
        if (a.out_rt)
            for (int j = 0; j < 6; j++) a.out_rt[(size_t)b * 6 + j] = prm[j];
    }
}

}  // namespace

cudaError_t dsac_sample_score_launch(const DsacArgs& a, cudaStream_t stream)
{
    if (a.B <= 0 || a.hyps <= 0) return cudaSuccess;
    dim3 gs((a.hyps + kSampleWarps - 1) / kSampleWarps, a.B);
    dsac_sample_kernel<<<gs, kSampleWarps * 32, 0, stream>>>(a);
    dsac_score_kernel<<<dim3(a.hyps, a.B), kScoreThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t dsac_forward_launch(const DsacArgs& a, cudaStream_t stream, cudaEvent_t* ev)
{
    if (a.B <= 0 || a.hyps <= 0) return cudaSuccess;
    if (a.out_counts) {
        cudaError_t e = cudaMemsetAsync(a.out_counts, 0xff, sizeof(int32_t) * (size_t)a.B * kMaxRefSteps, stream);
        if (e != cudaSuccess) return e;
    }
    dim3 gs((a.hyps + kSampleWarps - 1) / kSampleWarps, a.B);
    if (ev) cudaEventRecord(ev[0], stream);
    dsac_sample_kernel<<<gs, kSampleWarps * 32, 0, stream>>>(a);
    if (ev) cudaEventRecord(ev[1], stream);
    dsac_score_kernel<<<dim3(a.hyps, a.B), kScoreThreads, 0, stream>>>(a);
    if (ev) cudaEventRecord(ev[2], stream);
    {
        const int n = a.Hc * a.Wc;
        // measured at 60 x 90 cells (tools/dbg_refine_cluster.py, 32 frames): 1 / 2 / 4 / 8 CTAs per image = 0.80 / 0.74 / 0.58 /
        // 0.76 ms -- beyond four CTAs the cluster barriers of the ~300 reductions cost more than the shorter passes save; in
        // the live step (solve of batch k next to the stem of batch k + 1) four instead of eight: 19.48 -> 18.89 ms per step
        int cs = 1;
        while (cs < 8 && n > cs * kRefineThreads * 6) cs *= 2;
        static const int cs_env = [] { const char* e = getenv("CROSSLOC_B200_REFINE_CLUSTER"); return e ? atoi(e) : 0; }();
        if (cs_env == 1 || cs_env == 2 || cs_env == 4 || cs_env == 8) cs = cs_env;   // tuning knob
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(a.B * cs));
        cfg.blockDim = dim3(kRefineThreads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, dsac_refine_kernel, a);
        if (e != cudaSuccess) return e;
    }
    if (ev) cudaEventRecord(ev[3], stream);
    return cudaGetLastError();
}

}  // namespace cl
