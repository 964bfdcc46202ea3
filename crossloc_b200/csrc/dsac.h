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
    double* scores;         // [B, hyps]
    float* errs;            // [B, Hc*Wc] scratch: error map of the pose being refined
    float* out_pose;        // [B, 16] row-major camera-to-world
    int32_t* out_best;      // nullable [B]
    int32_t* out_counts;    // nullable [B, 100] inlier count per refinement step, -1 = not reached
    double* out_rt;         // nullable [B, 6] refined rvec, tvec
};

// ev: nullable [4] events recorded before sample, score, refine and after refine (cl_dsac_timing)
cudaError_t dsac_forward_launch(const DsacArgs& a, cudaStream_t stream, cudaEvent_t* ev = nullptr);

}  // namespace cl
