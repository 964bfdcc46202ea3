// Launch interface of the DSAC* kernels (device pointers only; the C-ABI in cabi.cu stages host buffers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace cl {

struct DsacArgs {
    const float* coords;    // [B, 3, Hc, Wc] planar X, Y, Z (the layout dsacstar.cpp:78-79 reads)
    int B, Hc, Wc, hyps;
    float thr;              // inlier threshold, px
    const float* focal;     // [B]
    float cx, cy, alpha, max_reproj;
    int S;                  // sub-sampling of the map w.r.t. the image
    uint64_t seed;
    uint32_t image_base;    // RNG image index of batch entry 0
    uint32_t max_tries;     // dsacstar.cpp:48 uses 1000000
    const int32_t* forced;  // nullable [B, hyps, 4, 2]: replay these cells, one try per hypothesis
    int refine;
    // workspace / outputs
    double* hyp_rt;         // [B, hyps, 6] rvec, tvec
    int32_t* tries;         // nullable [B, hyps]
    int32_t* out_cells;     // nullable [B, hyps, 4, 2]: the minimal set (x, y cells) each hypothesis was solved from
    double* scores;         // [B, hyps]
    float* errs;            // [B, Hc*Wc] scratch: error map of the pose being refined
    float* out_pose;        // [B, 16] row-major camera-to-world
    int32_t* out_best;      // nullable [B]
    int32_t* out_counts;    // nullable [B, 100] inlier count per refinement step, -1 = not reached
    double* out_rt;         // nullable [B, 6] refined rvec, tvec
};

// ev: nullable [4] events recorded before sample, score, refine and after refine (cl_dsac_timing)
cudaError_t dsac_forward_launch(const DsacArgs& a, cudaStream_t stream, cudaEvent_t* ev = nullptr);

// Sampling + scoring only (hyp_rt, tries, out_cells, scores): the first two stages of the backward pass.
cudaError_t dsac_sample_score_launch(const DsacArgs& a, cudaStream_t stream);

// dsacstar_rgb_backward (/root/reference/dsacstar/dsacstar.cpp:200-483).  Workspaces are per (image, hypothesis).
struct DsacBwdArgs {
    DsacArgs fwd;           // coords, sizes, camera, RNG; hyp_rt / tries / out_cells / scores must be set
    const float* gt_pose;   // [B, 16] ground-truth camera-to-world transforms, row-major
    float w_rot, w_trans, soft_clamp;
    double* probs;          // [B, hyps] softmax of the scores
    double* losses;         // [B, hyps] pose loss of every (refined) hypothesis
    double* ref_rt;         // [B, hyps, 6] refined rvec, tvec (the initial one where the probability is < 0.001)
    double* dloss;          // [B, hyps, 6] dLoss of the refined hypothesis
    int32_t* accepted;      // [B, hyps] refinement steps accepted
    float* errs;            // [B, hyps, Hc*Wc] scratch error maps
    uint8_t* inlier;        // [B, hyps, Hc*Wc] inlier map of the last accepted refinement step
    double* hyp_grad;       // [B, hyps, Hc*Wc, 3] per-hypothesis gradient p * (path I) + (path II)
    float* grad;            // [B, 3, Hc, Wc] accumulated into (+=), like the reference's outSceneCoordinatesGrad
    double* out_loss;       // [B] expected pose loss
};
cudaError_t dsac_backward_launch(const DsacBwdArgs& a, cudaStream_t stream);

}  // namespace cl
