/*
 * crossloc_b200 -- C ABI of the B200-native localization hot path.
 *
 * Drop-in boundary for TOPO-EPFL/CrossLoc's per-image localization path.  The reference has no C ABI:
 * its native boundary is the pybind11 module `dsacstar` (/root/reference/dsacstar/dsacstar.cpp:887-892)
 * taking ATen tensors, and its CNN is stock torch.nn.  Every entry point below names the reference
 * interface it replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer may be host or device memory unless stated otherwise
 *     (host buffers are staged through the library's workspace and the call synchronises the stream
 *     before returning when any output is a host buffer);
 *   - the caller owns all buffers; outputs are always written (the reference always writes outPose,
 *     dsacstar.cpp:174-177);
 *   - return value 0 on success, negative on error, message via cl_last_error() (thread-local);
 *   - `cuda_stream` is a cudaStream_t (NULL = legacy default stream); work is stream-ordered.
 */
#ifndef CROSSLOC_B200_H
#define CROSSLOC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Library / build identification: "crossloc_b200 <version> sm_100a". */
const char* cl_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* cl_last_error(void);

/*
 * DSAC* RGB forward for a batch of scene-coordinate maps.
 * Replaces dsacstar.forward_rgb / dsacstar_rgb_forward (/root/reference/dsacstar/dsacstar.cpp:63-178, :888),
 * called from /root/reference/utils/evaluation.py:162-172, generalised from batch size 1 to B.
 *
 *   coords          [B, 3, Hc, Wc] float, planar X/Y/Z                      (sceneCoordinatesSrc)
 *   out_pose        [B, 16] float, row-major 4x4 camera-to-world            (outPoseSrc)
 *   hyps            number of RANSAC hypotheses                             (ransacHypotheses)
 *   thr             inlier threshold in px                                  (inlierThreshold)
 *   focal           [B] float focal length in px                            (focalLength)
 *   cx, cy          principal point                                         (ppointX, ppointY)
 *   alpha           soft-inlier scale                                       (inlierAlpha)
 *   max_reproj      reprojection errors are clamped to this (px)            (maxReproj)
 *   subsample       map sub-sampling w.r.t. the image                       (subSampling)
 *   seed            RNG seed; the reference seeds mt19937 with 1305 once per process
 *                   (thread_rand.cpp:13-30); here every (seed, image, hypothesis, try) names a fixed draw
 *   image_base      RNG image index of batch entry 0 (entry b uses image_base + b)
 *   max_tries       hypothesis re-sampling limit (dsacstar.cpp:48 uses 1000000)
 *   refine          0 skips refineHyp (debug), 1 = reference behaviour
 *   forced_samples  nullable [B, hyps, 4, 2] int32 (x, y) cells: replay mode, one try per hypothesis
 *   out_best        nullable [B] int32 index of the selected hypothesis
 *   out_scores      nullable [B, hyps] double soft-inlier scores
 *   out_hyps        nullable [B, hyps, 6] double rvec, tvec of every hypothesis
 *   out_tries       nullable [B, hyps] int32 tries used
 *   out_counts      nullable [B, 100] int32 inlier count at each refinement step (-1 = not reached)
 *   out_rt          nullable [B, 6] double refined rvec, tvec
 */
int cl_dsac_forward_rgb(const float* coords, int B, int Hc, int Wc, float* out_pose, int hyps, float thr,
                        const float* focal, float cx, float cy, float alpha, float max_reproj, int subsample,
                        uint64_t seed, uint32_t image_base, uint32_t max_tries, int refine,
                        const int32_t* forced_samples, int32_t* out_best, double* out_scores, double* out_hyps,
                        int32_t* out_tries, int32_t* out_counts, double* out_rt, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* CROSSLOC_B200_H */
