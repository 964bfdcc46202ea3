// DSAC* backward pass for sm_100a: the expected pose loss over all hypotheses and its gradient w.r.t. the scene
// coordinate map.  Reference: /root/reference/dsacstar/dsacstar.cpp:200-483 with
// dsacstar_derivative.h (dPNP, dScore, dSMScore, dProjectdObj) and dsacstar_loss.h (loss, dLoss).
//
// The reference walks the hypotheses with OpenMP and holds a dense 6 x 3n Jacobian per hypothesis; here one CTA owns
// one (image, hypothesis) pair end to end and nothing larger than the per-hypothesis gradient map is materialised:
//   1. dsac_sample_kernel / dsac_score_kernel (dsac.cu)  minimal sets, P3P hypotheses, soft inlier scores
//   2. bwd_probs_kernel    softMax over the scores (sequential sums, as dsacstar_util.h:684-706)
//   3. bwd_refine_kernel   refineHyp for every hypothesis with probability >= 0.001, keeping the inlier map of the
//                          last accepted step; pose loss and dLoss of the refined pose
//   4. bwd_grad_kernel     path I (implicit derivative of the refinement at its optimum, -(J^T J)^+ J^T) and path II
//                          (soft inlier score: direct term per cell + central-difference dPNP for the minimal set)
//   5. bwd_assemble_kernel hypotheses summed in index order into the float gradient map (+=), expected loss
// Every cell of a map is owned by the same thread through a kernel (i = thread, thread + 256, ...), so no atomics are
// needed and the result is deterministic.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "dsac.h"
#include "dsac_backward_math.cuh"
#include "dsac_common.cuh"
#include "dsac_refine.cuh"

namespace cl {

namespace {

constexpr int kBwdThreads = kRefineThreads;

__global__ void bwd_probs_kernel(DsacBwdArgs a)
{
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    const int hyps = a.fwd.hyps;
    const double* scores = a.fwd.scores + (size_t)b * hyps;
    double* probs = a.probs + (size_t)b * hyps;
    double mx = scores[0], sum = 0;
    for (int h = 1; h < hyps; h++) if (scores[h] > mx) mx = scores[h];
    for (int h = 0; h < hyps; h++) { probs[h] = exp(scores[h] - mx); sum += probs[h]; }
    for (int h = 0; h < hyps; h++) probs[h] /= sum;
}

__global__ void __launch_bounds__(kBwdThreads) bwd_refine_kernel(DsacBwdArgs a)
{
    __shared__ double s_warp_part[(kBwdThreads / 32) * 28];
    __shared__ double s_part[2 * 28];
    __shared__ double s_total[28];
    const DsacArgs& fa = a.fwd;
    const int h = blockIdx.x, b = blockIdx.y;
    const size_t idx = (size_t)b * fa.hyps + h;
    const int n = fa.Hc * fa.Wc;
    const int first = threadIdx.x, stride = kBwdThreads;
    ClusterRed red{s_warp_part, s_part, s_total, 0u};
    const float* X = fa.coords + (size_t)b * 3 * n;
    float* errs = a.errs + idx * n;
    uint8_t* inl = a.inlier + idx * n;
    const float fl = fa.focal[b];
    const double p = a.probs[idx];

    double prm[6];
#pragma unroll
    for (int j = 0; j < 6; j++) prm[j] = fa.hyp_rt[idx * 6 + j];
    int accepted = 0;
    if (p >= kBwdProbThresh) {   // dsacstar.cpp:310
        int inliers = error_map(prm, X, errs, n, fa.Wc, fa.S, fl, fa.cx, fa.cy, fa.thr, fa.max_reproj, red, first, stride);
        int best_inliers = 4;
        for (int step = 0; step < kMaxRefSteps; step++) {
            if (inliers <= best_inliers) break;
            best_inliers = inliers;
            double upd[6];
#pragma unroll
            for (int j = 0; j < 6; j++) upd[j] = prm[j];
            if (!lm_solve(upd, X, errs, n, fa.Wc, fa.S, fa.thr, (double)fl, (double)fa.cx, (double)fa.cy, red, first, stride))
                break;
            for (int i = first; i < n; i += stride) inl[i] = errs[i] < fa.thr ? 1 : 0;   // inlierMap = localInlierMap
#pragma unroll
            for (int j = 0; j < 6; j++) prm[j] = upd[j];
            accepted++;
            inliers = error_map(prm, X, errs, n, fa.Wc, fa.S, fl, fa.cx, fa.cy, fa.thr, fa.max_reproj, red, first, stride);
        }
    }
    if (threadIdx.x == 0) {
        double gt[16];
        for (int j = 0; j < 16; j++) gt[j] = a.gt_pose[(size_t)b * 16 + j];
        for (int j = 0; j < 6; j++) a.ref_rt[idx * 6 + j] = prm[j];
        a.accepted[idx] = accepted;
        a.losses[idx] = pose_loss(prm, gt, a.w_rot, a.w_trans, a.soft_clamp);
        double jac[6] = {0, 0, 0, 0, 0, 0};
        if (p >= kBwdProbThresh) {
            double gt_rt[6];
            trans_to_pose(gt, gt_rt);
            pose_loss_jacobian(prm, gt_rt, a.w_rot, a.w_trans, a.soft_clamp, jac);
        }
        for (int j = 0; j < 6; j++) a.dloss[idx * 6 + j] = jac[j];
    }
}

__device__ __forceinline__ double block_max(double v, double* smem /* [kBwdThreads / 32 + 1] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = smem[0];
        for (int w = 1; w < kBwdThreads / 32; w++) m = fmax(m, smem[w]);
        smem[kBwdThreads / 32] = m;
    }
    __syncthreads();
    return smem[kBwdThreads / 32];
}

__global__ void __launch_bounds__(kBwdThreads) bwd_grad_kernel(DsacBwdArgs a)
{
    __shared__ double s_warp_part[(kBwdThreads / 32) * 28];
    __shared__ double s_part[2 * 28];
    __shared__ double s_total[28];
    __shared__ double s_max[kBwdThreads / 32 + 1];
    __shared__ double s_pinv[36];
    __shared__ double s_sg;
    __shared__ double s_pert[18][12];   // dPNP: the 18 perturbed minimal sets
    __shared__ double s_sol[18][6];
    __shared__ int s_ok[18];
    __shared__ double s_dhdo[6 * 12];
    __shared__ int s_dhdo_zero;

    const DsacArgs& fa = a.fwd;
    const int h = blockIdx.x, b = blockIdx.y;
    const size_t idx = (size_t)b * fa.hyps + h;
    const double p = a.probs[idx];
    if (p < kBwdProbThresh) return;   // block-uniform: dsacstar.cpp:352, dsacstar_derivative.h:262
    const int n = fa.Hc * fa.Wc;
    const int first = threadIdx.x, stride = kBwdThreads;
    ClusterRed red{s_warp_part, s_part, s_total, 0u};
    const float* X = fa.coords + (size_t)b * 3 * n;
    const uint8_t* inl = a.inlier + idx * n;
    double* T = a.hyp_grad + idx * n * 3;
    const float fl = fa.focal[b];
    const double f = fl, cx = fa.cx, cy = fa.cy;
    const int32_t* cells = fa.out_cells + idx * 8;

    // ---- per-hypothesis scalars: score gradient (dSMScore) and the perturbed minimal sets of dPNP
    if (threadIdx.x == 0) {
        const double* probs = a.probs + (size_t)b * fa.hyps;
        const double* losses = a.losses + (size_t)b * fa.hyps;
        double sg = p * losses[h];
        for (int j = 0; j < fa.hyps; j++) sg -= p * probs[j] * losses[j];   // dsacstar_derivative.h:373-375
        s_sg = sg;
        // float perturbations applied in sequence, drift included (dsacstar_derivative.h:153-176)
        float obj[12];
        for (int j = 0; j < 4; j++) {
            const int i = cells[2 * j + 1] * fa.Wc + cells[2 * j];
            obj[3 * j] = X[i]; obj[3 * j + 1] = X[n + i]; obj[3 * j + 2] = X[2 * n + i];
        }
        const float eps = 0.001f;
        for (int c = 0; c < 9; c++) {
            obj[c] += eps;
            for (int k = 0; k < 12; k++) s_pert[2 * c][k] = obj[k];
            obj[c] -= 2 * eps;
            for (int k = 0; k < 12; k++) s_pert[2 * c + 1][k] = obj[k];
            obj[c] += eps;
        }
        s_dhdo_zero = 0;
    }
    __syncthreads();
    if (threadIdx.x < 18) {
        double img[8];
        for (int j = 0; j < 4; j++) {
            int px, py;
            cell_pixel(cells[2 * j], cells[2 * j + 1], fa.S, px, py);
            img[2 * j] = px; img[2 * j + 1] = py;
        }
        Pose sol;
        const bool ok = p3p_solve(s_pert[threadIdx.x], img, f, cx, cy, sol);
        s_ok[threadIdx.x] = ok;
        for (int j = 0; j < 3; j++) { s_sol[threadIdx.x][j] = sol.r[j]; s_sol[threadIdx.x][3 + j] = sol.t[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int c = threadIdx.x;
        bool bad = false;
        double mx = 0;
        for (int k = 0; k < 6; k++) {
            double v = 0;
            if (c < 9) {   // the 4th point only disambiguates: zero derivative
                if (!s_ok[2 * c] || !s_ok[2 * c + 1]) bad = true;
                v = (s_sol[2 * c][k] - s_sol[2 * c + 1][k]) / (double)(2 * 0.001f);
                if (v != v) bad = true;
            }
            s_dhdo[12 * k + c] = v;
            mx = fmax(mx, fabs(v));
        }
        if (bad || mx > 10) atomicOr(&s_dhdo_zero, 1);   // failure / NaN -> zeros; getMax > 10 -> zeros (:289)
    }
    __syncthreads();
    const double sg = s_sg;

    // ---- path I: -(J^T J)^+ J^T at the refined pose over the inliers of the last accepted step
    double ref[6], dl[6], a6[6];
#pragma unroll
    for (int j = 0; j < 6; j++) { ref[j] = a.ref_rt[idx * 6 + j]; dl[j] = a.dloss[idx * 6 + j]; a6[j] = 0; }
    ProjJac pjr;
    proj_jac_setup(ref, pjr);
    bool path1 = a.accepted[idx] > 0;
    if (path1) {
        double acc[22];
#pragma unroll
        for (int j = 0; j < 22; j++) acc[j] = 0;
        for (int i = first; i < n; i += stride) {
            if (!inl[i]) continue;
            const int yy = i / fa.Wc, xx = i - yy * fa.Wc;
            int px, py;
            cell_pixel(xx, yy, fa.S, px, py);
            double row[6];
            residual_jacobian_row(pjr, f, cx, cy, X[i], X[n + i], X[2 * n + i], px, py, fa.max_reproj, row);
            int k = 0;
#pragma unroll
            for (int r2 = 0; r2 < 6; r2++)
#pragma unroll
                for (int c = r2; c < 6; c++) acc[k++] += row[r2] * row[c];
            acc[21] += 1;
        }
        cluster_reduce_sum<22>(acc, red);
        if (acc[21] < 4) {
            path1 = false;   // dsacstar.cpp:381
        } else {
            if (threadIdx.x == 0) {
                double A[36];
                int k = 0;
                for (int r2 = 0; r2 < 6; r2++)
                    for (int c = r2; c < 6; c++) { A[6 * r2 + c] = acc[k]; A[6 * c + r2] = acc[k]; k++; }
                double P[36];
                sym6_pinv(A, P);
                for (int i = 0; i < 36; i++) s_pinv[i] = P[i];
            }
            __syncthreads();
            double mx = 0;
            for (int i = first; i < n; i += stride) {
                if (!inl[i]) continue;
                const int yy = i / fa.Wc, xx = i - yy * fa.Wc;
                int px, py;
                cell_pixel(xx, yy, fa.S, px, py);
                double row[6];
                residual_jacobian_row(pjr, f, cx, cy, X[i], X[n + i], X[2 * n + i], px, py, fa.max_reproj, row);
                for (int r2 = 0; r2 < 6; r2++) {
                    double v = 0;
                    for (int c = 0; c < 6; c++) v += s_pinv[6 * r2 + c] * row[c];
                    mx = fmax(mx, fabs(v));
                }
            }
            mx = block_max(mx, s_max);
            if (mx > 10) {
                path1 = false;   // jacobeanR = 0, dsacstar.cpp:411
            } else {
                for (int c = 0; c < 6; c++) {
                    double v = 0;
                    for (int r2 = 0; r2 < 6; r2++) v += dl[r2] * s_pinv[6 * r2 + c];
                    a6[c] = -v;   // dLoss * (-(J^T J)^+)
                }
            }
        }
    }

    // ---- path II direct term + path I scatter, one store per cell
    double ini[6];
#pragma unroll
    for (int j = 0; j < 6; j++) ini[j] = fa.hyp_rt[idx * 6 + j];
    ProjJac pji;
    proj_jac_setup(ini, pji);
    const float beta = 5 / fa.thr;
    const double scale = (double)(fa.alpha / fa.Wc / fa.Hc);
    double sup[6] = {0, 0, 0, 0, 0, 0};
    for (int i = first; i < n; i += stride) {
        const int yy = i / fa.Wc, xx = i - yy * fa.Wc;
        int px, py;
        cell_pixel(xx, yy, fa.S, px, py);
        const float Xw = X[i], Yw = X[n + i], Zw = X[2 * n + i];
        const float e = repro_error(pji.R, pji.t, fl, fa.cx, fa.cy, Xw, Yw, Zw, px, py, fa.max_reproj);
        double soft = beta * (e - fa.thr);
        soft = 1 / (1 + exp(-soft));
        double d = -soft * (1 - soft) * (double)beta * sg;   // dsacstar_derivative.h:269-271
        d *= scale;
        double dpo[3], row[6];
        d_project_d_obj((float)px, (float)py, Xw, Yw, Zw, pji.R, pji.t, f, cx, cy, fa.max_reproj, dpo);
        residual_jacobian_row(pji, f, cx, cy, Xw, Yw, Zw, px, py, fa.max_reproj, row);
#pragma unroll
        for (int k = 0; k < 6; k++) sup[k] += d * row[k];
        double g[3] = {dpo[0] * d, dpo[1] * d, dpo[2] * d};
        if (path1 && inl[i]) {
            residual_jacobian_row(pjr, f, cx, cy, Xw, Yw, Zw, px, py, fa.max_reproj, row);
            double s = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) s += a6[k] * row[k];
            d_project_d_obj((float)px, (float)py, Xw, Yw, Zw, pjr.R, pjr.t, f, cx, cy, fa.max_reproj, dpo);
#pragma unroll
            for (int c = 0; c < 3; c++) g[c] += p * (s * dpo[c]);
        }
        T[(size_t)i * 3] = g[0]; T[(size_t)i * 3 + 1] = g[1]; T[(size_t)i * 3 + 2] = g[2];
    }
    cluster_reduce_sum<6>(sup, red);   // ends with a block barrier: every T store above is visible below
    if (threadIdx.x == 0 && !s_dhdo_zero) {
        for (int j = 0; j < 4; j++) {
            const size_t i = (size_t)cells[2 * j + 1] * fa.Wc + cells[2 * j];
            for (int c = 0; c < 3; c++) {
                double v = 0;
                for (int k = 0; k < 6; k++) v += sup[k] * s_dhdo[12 * k + 3 * j + c];
                T[i * 3 + c] += v;
            }
        }
    }
}

__global__ void __launch_bounds__(256) bwd_assemble_kernel(DsacBwdArgs a)
{
    const DsacArgs& fa = a.fwd;
    const int b = blockIdx.y;
    const int n = fa.Hc * fa.Wc;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const double* probs = a.probs + (size_t)b * fa.hyps;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double* losses = a.losses + (size_t)b * fa.hyps;
        double e = 0;
        for (int h = 0; h < fa.hyps; h++) e += probs[h] * losses[h];   // dsacstar.cpp:331-336
        a.out_loss[b] = e;
    }
    if (i >= n) return;
    float g[3];
#pragma unroll
    for (int c = 0; c < 3; c++) g[c] = a.grad[((size_t)b * 3 + c) * n + i];
    for (int h = 0; h < fa.hyps; h++) {
        if (probs[h] < kBwdProbThresh) continue;
        const double* T = a.hyp_grad + (((size_t)b * fa.hyps + h) * n + i) * 3;
#pragma unroll
        for (int c = 0; c < 3; c++) g[c] = (float)((double)g[c] + T[c]);   // float += double, dsacstar.cpp:471-476
    }
#pragma unroll
    for (int c = 0; c < 3; c++) a.grad[((size_t)b * 3 + c) * n + i] = g[c];
}

}  // namespace

cudaError_t dsac_backward_launch(const DsacBwdArgs& a, cudaStream_t stream)
{
    const DsacArgs& fa = a.fwd;
    if (fa.B <= 0 || fa.hyps <= 0) return cudaSuccess;
    cudaError_t e = dsac_sample_score_launch(fa, stream);
    if (e != cudaSuccess) return e;
    bwd_probs_kernel<<<fa.B, 32, 0, stream>>>(a);
    bwd_refine_kernel<<<dim3(fa.hyps, fa.B), kBwdThreads, 0, stream>>>(a);
    bwd_grad_kernel<<<dim3(fa.hyps, fa.B), kBwdThreads, 0, stream>>>(a);
    const int n = fa.Hc * fa.Wc;
    bwd_assemble_kernel<<<dim3((n + 255) / 256, fa.B), 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace cl
