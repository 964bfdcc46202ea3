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

#include <stddef.h>
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

/*
 * cl_dsac_backward_rgb -- replaces dsacstar_rgb_backward (/root/reference/dsacstar/dsacstar.cpp:200-215, bound as
 * `dsacstar.backward_rgb` at :889): expected pose loss over all hypotheses and its gradient w.r.t. the scene
 * coordinates (path I through the refined hypotheses, path II through the soft inlier scores; dsacstar_derivative.h,
 * dsacstar_loss.h).  The reference handles one image per call; here B images, each as the reference would.
 *
 *   coords, B, Hc, Wc, hyps, thr, focal, cx, cy, alpha, max_reproj, subsample, seed, image_base, max_tries,
 *   forced_samples       as for cl_dsac_forward_rgb
 *   grad                 [B, 3, Hc, Wc] float32, ACCUMULATED into (+=) like outSceneCoordinatesGradSrc (:469-477)
 *   gt_pose              [B, 16] float32 row-major ground-truth camera-to-world transforms (gtPoseSrc)
 *   w_rot, w_trans       loss weights of the rotation (degrees) and translation error (wLossRot, wLossTrans)
 *   soft_clamp           loss value above which sqrt(soft_clamp * loss) is used (softClamp)
 *   out_loss             [B] double expected loss (the reference's return value)
 *   out_probs            nullable [B, hyps] double selection probabilities
 *   out_losses           nullable [B, hyps] double loss of every (refined) hypothesis
 *   out_hyps             nullable [B, hyps, 6] double initial rvec, tvec
 *   out_ref_rt           nullable [B, hyps, 6] double refined rvec, tvec (initial where the probability is < 0.001)
 *   out_tries            nullable [B, hyps] int32 tries used
 *   out_cells            nullable [B, hyps, 4, 2] int32 minimal sets (x, y)
 * Workspace: B * hyps * Hc * Wc * 29 bytes (per-hypothesis gradient, error and inlier maps).
 */
int cl_dsac_backward_rgb(const float* coords, int B, int Hc, int Wc, float* grad, const float* gt_pose, int hyps,
                         float thr, const float* focal, float cx, float cy, float w_rot, float w_trans,
                         float soft_clamp, float alpha, float max_reproj, int subsample, uint64_t seed,
                         uint32_t image_base, uint32_t max_tries, const int32_t* forced_samples, double* out_loss,
                         double* out_probs, double* out_losses, double* out_hyps, double* out_ref_rt,
                         int32_t* out_tries, int32_t* out_cells, void* cuda_stream);

/*
 * Re-entrancy.  Intermediates of cl_dsac_forward_rgb (hypotheses, scores, error maps, staged host buffers) live in a
 * workspace private to (current device, cuda_stream): calls on different streams may be in flight together, calls
 * on one stream are serialised by the library.  Buffers grow on demand (the stream is synchronised before an old
 * allocation is replaced) and are kept until cl_release_workspaces(), which frees those of the current device.
 *
 * cl_dsac_timing(enable, stream, ms, solves): enable != 0 makes later solves record CUDA events between their kernels;
 * with ms != NULL it first returns the device milliseconds {sample, score, refine} summed over the `solves` timed
 * solves issued on `stream` since the last read, and clears that record.
 */
int cl_dsac_timing(int enable, void* cuda_stream, float* ms, int* solves);
int cl_release_workspaces(void);

/*
 * ---- Scene-coordinate CNN operators ---------------------------------------------------------------
 * These replace the cuDNN / ATen kernels stock PyTorch launches for the reference network
 * (/root/reference/networks/networks.py:175-256 encoder, :276-360 decoder, :43-130 vanilla Network).
 * All tensor arguments of this group must be DEVICE pointers (activations never leave HBM); small
 * parameter tables (`tap_a_row`) are host pointers.
 *
 * Activation layout ("padded-flat", PF): fp16 matrix [rows][C],
 *   row(t, ph, b, y, x) = ((t * P + ph) * B + b) * (H + 2) * (W + 2) + (y + 1) * (W + 2) + (x + 1)
 * t = fp16 split term (0 hi, 1 lo), ph = parity phase (P = 1; P = 4 for the input of a stride-2
 * convolution, stored at the OUTPUT resolution), borders zero.  In this layout every filter tap is a
 * constant row shift, so a convolution is a sum of shifted GEMMs fed by plain 2-D TMA boxes.
 */

/*
 * Implicit-GEMM convolution on the tcgen05 tensor cores (+ bias, + GroupNorm partial sums).
 * Replaces nn.Conv2d 3x3 s1 / 3x3 s2 / 1x1 (networks.py:191-213, 133-146, 297-306).
 *   act           fp16 PF activation matrix (all planes), a_total_rows rows of Cin channels
 *   a_lo_rows     row distance between the hi and the lo plane (ignored when nterms == 1)
 *   weights       fp16 [nterms == 3 ? 2 : 1][num_taps][Cout][Cin], pre-scaled by 1 / out_scale
 *   tap_a_row     HOST int32 [num_taps]: activation row shift of each tap (phase offset included)
 *   nterms        1 = one fp16 pass; 3 = fp16x3 split (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo), fp32-grade;
 *                 2 = fp16 + fp8: a_hi*w_hi in fp16, the two correction products as e4m3 MMAs on scaled operands
 *   act8          nterms == 2 only: e4m3 PF matrix with a8_total_rows rows, plane 0 = fp8(a_hi * 2^2),
 *                 plane 1 (a8_lo_rows further) = fp8((a - a_hi) * 2^14), as cl_gn_apply writes them
 *   weights8      nterms == 2 only: e4m3 [2][num_taps][Cout][Cin] = fp8(w_hi), fp8(w_lo * 2^12)
 *   Mp, Hp, Wp    output rows B * Hp * Wp and padded plane size (Hp = H + 2, Wp = W + 2)
 *   group_ch      channels per GroupNorm group for the statistics (0 = none; 2, 4, 8 or 16)
 *   raw           fp32 [Mp][Cout] output (interior rows written), bias fp32 [Cout]
 *   stats         fp64 [B][Cout / group_ch][2] (sum, sum of squares), accumulated: caller zeroes it
 */
int cl_conv_igemm(const void* act, int64_t a_total_rows, int64_t a_lo_rows, int Cin, const void* weights, int Cout,
                  int num_taps, const int32_t* tap_a_row, int nterms, int Mp, int Hp, int Wp, int group_ch,
                  float out_scale, float* raw, const float* bias, double* stats, const void* act8,
                  int64_t a8_total_rows, int64_t a8_lo_rows, const void* weights8, void* cuda_stream);

/*
 * The same convolution in fp16 + fp4 mode (CTA-pair kernel only): a_hi*w_hi in fp16, the two correction products
 * (each 2^-11 of the result) as block-scaled e2m1 MMAs (tcgen05 kind::mxf4, four times the fp16 rate) -- 1.5 instead
 * of 2 fp16-MMA equivalents per product.  Cin % 256 == 0, Cout % 256 == 0.
 *   act4      e2m1 PF matrix [2][a4_lo_rows][Cin/2], two values per byte (even channel in the low nibble):
 *             plane 0 = fp4(a_hi / 2^s_hi), plane 1 = fp4((a - a_hi) / 2^s_lo), as cl_gn_apply_fp4 writes them
 *   act_sf    uint32 [Cin/256][a4_lo_rows]: ue8m0 exponent bytes (s_lo, s_lo, s_hi, s_hi) of a pixel row's 256-channel block
 *   weights4, w_sf   as cl_pack_conv_fp4 writes them
 */
int cl_conv_igemm_fp4(const void* act, int64_t a_total_rows, int64_t a_lo_rows, int Cin, const void* weights, int Cout,
                      int num_taps, const int32_t* tap_a_row, int Mp, int Hp, int Wp, int group_ch, float out_scale,
                      float* raw, const float* bias, double* stats, const void* act4, int64_t a4_total_rows,
                      int64_t a4_lo_rows, const void* act_sf, const void* weights4, const void* w_sf,
                      void* cuda_stream);
/* OIHW fp32 filter (x scale, a power of two) -> weights4 e2m1 [2][num_taps][Cout][Cin/2] = fp4(w_hi), fp4(w_lo), one
 * ue8m0 scale per (tap, output channel, 256 input channels) and plane; w_sf uint32 [num_taps][Cin/256][Cout/128][128],
 * bytes (s_hi, s_hi, s_lo, s_lo), word l*4 + j of a 128-channel block = output channel 32*j + l (tcgen05.cp order). */
int cl_pack_conv_fp4(const float* w, int Cout, int Cin, int num_taps, float scale, void* weights4, void* w_sf,
                     void* cuda_stream);

/*
 * Convolution weight gradient (training step) on padded-flat operands, the layout
 * of cl_conv_igemm / cl_gn_backward:
 *   dw[tap][co][ci] += out_scale * sum over rows r < Mp of grad[r][co] * act[tap_phase[tap]][r + tap_shift[tap]][ci]
 * Both operands are fp16 hi/lo PF matrices [planes][rows][C] read as MN-major tensor-core operands (64-row x
 * 64-channel TMA boxes, SWIZZLE_128B): no channel-major copies.  grad: plane 0 = hi, 1 = lo, g_plane_rows apart,
 * zero border rows.  act: plane term * phases + phase, x_plane_rows apart; rows outside a plane read as zero.
 * tap_shift / tap_phase are the forward convolution's tap table (phase and in-plane row shift kept apart).
 * Replaces the cuDNN wgrad kernels autograd launches for nn.Conv2d (train_single_task.py:298).
 *   Cout % 64 == 0, Cin = 32 or a multiple of 64, nterms 1 | 3 (fp16x3).  dw fp32, accumulated with atomics (caller
 *   zeroes it): [num_taps][Cout][Cin], or with oihw = 1 [Cout][Cin][num_taps] = torch's OIHW weight layout.
 *   scale_dev: nullable device scalar multiplied into the result (the 2^-k of a device-rescaled gradient).
 */
int cl_conv_wgrad_pf(const void* grad, int64_t g_plane_rows, const void* act, int64_t x_plane_rows, int Mp, int Cout,
                     int Cin, int phases, int num_taps, const int32_t* tap_shift, const int32_t* tap_phase, int nterms,
                     float out_scale, const float* scale_dev, int oihw, float* dw, void* cuda_stream);

/*
 * Backward of one "GroupNorm -> ReLU -> (residual merge -> ReLU)" stage on padded-flat tensors (training step):
 * the autograd kernels behind nn.GroupNorm / F.relu / the residual adds (networks.py:231-254, 332-343;
 * train_single_task.py:298).  pass 0 sums the `num_src` gradient sources (fp32 PF, each x two optional device
 * scalars; `src_phased[i]` = 1 for the 4-phase result of a stride-2 data gradient), masks them with the sign of
 * the stage's merged output (`mask_out`, fp16 hi plane, nullable), optionally stores the result in `g_out`, and
 * accumulates ab[b][c] = (sum dy, sum dy * xhat) in fp64 plus the bits of max |dy * gamma| * rstd.  pass 1 reads
 * src[0] only and writes d_raw = gradient of the raw convolution output as fp16 hi / lo PF planes x 2^k
 * (scale_out = {2^k, 2^-k}, derived on the device), adding its per-channel sums to `dbias` (nullable).
 * d_gamma = sum_b ab[b][c][1], d_beta = sum_b ab[b][c][0].  group_ch = 0: no normalisation (vanilla Network).
 * ab is indexed [b][ab_C][2] (ab_C = 0 means C): the stages of one backward pass can share a buffer, each at its offset.
 * Caller zeroes ab, gmax_bits, dbias; d_raw keeps zero border rows (only interior pixels are written).
 * d_raw_f32 (nullable): the same gradient, unscaled, as fp32 PF [rows][C] (the stem's weight gradient is left to torch).
 */
int cl_gn_backward(int pass, int B, int H, int W, int C, int group_ch, const float* raw, const double* stats,
                   const float* gamma, const float* beta, float eps, int relu_inner, int num_src,
                   const float* const* src, const float* const* src_scale_a, const float* const* src_scale_b,
                   const int32_t* src_stride, const int32_t* src_phased, const void* mask_out, float* g_out, double* ab,
                   int ab_C, void* gmax_bits, void* d_raw, int64_t d_raw_lo_rows, float* scale_out, double* dbias,
                   float* d_raw_f32, void* cuda_stream);
/* The same with, in pass 1, block-scaled e2m1 planes of the scaled gradient next to the fp16 ones (C % 256 == 0), in the
 * layout cl_gn_apply_fp4 writes for the forward: d_raw4 [2][d_raw4_lo_rows][C/2], d_raw_sf uint32 [C/256][d_raw4_lo_rows].
 * They feed cl_conv_igemm_fp4 with the transposed filter = the data gradient at 1.5 fp16-MMA equivalents per product. */
int cl_gn_backward_fp4(int pass, int B, int H, int W, int C, int group_ch, const float* raw, const double* stats,
                       const float* gamma, const float* beta, float eps, int relu_inner, int num_src,
                       const float* const* src, const float* const* src_scale_a, const float* const* src_scale_b,
                       const int32_t* src_stride, const int32_t* src_phased, const void* mask_out, float* g_out, double* ab,
                       int ab_C, void* gmax_bits, void* d_raw, int64_t d_raw_lo_rows, float* scale_out, double* dbias,
                       float* d_raw_f32, void* d_raw4, int64_t d_raw4_lo_rows, void* d_raw_sf, void* cuda_stream);

/*
 * Backward of the 1x1 output head (fc3, networks.py:349) on the padded-flat activation it read (training step):
 *   g_x[row][c] = sum_o g_sc[b][o][y][x] * weight[o][c]      fp32 PF [B*(H+2)*(W+2)][C], interior rows -- the gradient
 *                                                             source of the last stage's cl_gn_backward
 *   g_w[o][c]  += sum over pixels g_sc[b][o][y][x] * act[row][c]   (fp32 atomics, caller zeroes)
 * g_sc is the gradient of the head's pre-activation output (NCHW fp32; the mean offset and exp(clamp) derivatives
 * are applied by the caller on the small output map).  act: fp16 hi / lo planes act_lo_rows apart.  Co <= 8.
 */
int cl_head_backward(const void* act, int64_t act_lo_rows, int B, int H, int W, int C, int Co, const float* weight,
                     const float* g_sc, float* g_x, float* g_w, void* cuda_stream);

/*
 * Layout kernels of the training path (device pointers): NCHW fp32 tensors, as autograd hands them over, to and
 * from the operand layouts of cl_conv_igemm / cl_conv_wgrad_pf, and filter packing.  `scale` arguments are device
 * scalars (power-of-two factors computed on the GPU, no host synchronisation), NULL = 1.
 *   cl_nchw_to_pf   x [B][C][H][W] -> fp16 hi/lo PF [2][phases][B*(H'+2)*(W'+2)][C] (interior; caller zero-fills)
 *   cl_pf_to_nchw   raw fp32 PF [B*(H+2)*(W+2)][Craw] -> out [B][C][Hout][Wout] at (y*step+off_y, x*step+off_x),
 *                   x scale, + bias (the per-phase results of a stride-2 data gradient use step 2)
 *   cl_pack_filter  w OIHW -> fp16 hi/lo [2][num_taps][N][K] for taps (tap_kh, tap_kw), transposed for dgrad
 */
int cl_nchw_to_pf(const float* x, const float* scale, void* out, int B, int C, int H, int W, int phases,
                  void* cuda_stream);
int cl_pf_to_nchw(const float* raw, int B, int H, int W, int Craw, float* out, int C, int Hout, int Wout, int step,
                  int off_y, int off_x, const float* scale, const float* bias, void* cuda_stream);
/* workspace: 4 x 32-bit device words; on completion workspace[0] = 2^k, workspace[1] = 2^-k (floats) with
 * k = floor(log2(target / max|x|)) -- the power-of-two operand scales, computed without a host round trip */
int cl_pow2_scale(const float* x, int64_t n, float target, void* workspace, void* cuda_stream);
int cl_pack_filter(const float* w, const float* scale, void* out, int Cout, int Cin, int ksize, int num_taps,
                   const int32_t* tap_kh, const int32_t* tap_kw, int transpose, int N, int K, void* cuda_stream);

/*
 * GroupNorm apply + ReLU + residual merge, fp32 raw -> fp16 hi/lo PF input of the next convolution.
 * Replaces nn.GroupNorm + F.relu (+ `res + x`) (networks.py:231-254, 332-343).
 *   out = relu_outer( add + relu_inner( gn(raw) ) ),  add = 0 | res_hi + res_lo | gn2(raw2)
 *   res_lo_rows: row distance of the residual's lo plane, 0 when it has none (single-pass mode)
 *   out_phases 1: same geometry;  4: the four parity phases at (ceil(H/2), ceil(W/2)).
 *   out8: nullable e4m3 planes [2][B*(H+2)*(W+2)][C] written alongside (input of an nterms == 2 convolution).
 *   out_C / out_c0: the destination matrices have out_C channels per row (0 = C) and the result goes to channels
 *   out_c0 .. out_c0 + C - 1: the encoders of the MLR model write side by side into one concatenated activation
 *   (torch.cat(..., dim=1), networks.py:488).
 */
int cl_gn_apply(const float* raw, int B, int H, int W, int C, int group_ch, const double* stats, const float* gamma,
                const float* beta, float eps, int relu_inner, int add_kind, const void* res, int64_t res_lo_rows,
                const float* raw2, const double* stats2, const float* gamma2, const float* beta2, int relu_outer,
                void* out, int out_phases, int out_terms, void* out8, int out_C, int out_c0, void* cuda_stream);

/* cl_gn_apply that also writes the block-scaled e2m1 planes of a consumer in fp16 + fp4 mode (nullable out4 / out_sf;
 * out_phases 1, C % 256 == 0, full-width destination): out4 [2][B*(H+2)*(W+2)][C/2], out_sf [C/256][B*(H+2)*(W+2)]. */
int cl_gn_apply_fp4(const float* raw, int B, int H, int W, int C, int group_ch, const double* stats, const float* gamma,
                    const float* beta, float eps, int relu_inner, int add_kind, const void* res, int64_t res_lo_rows,
                    const float* raw2, const double* stats2, const float* gamma2, const float* beta2, int relu_outer,
                    void* out, int out_phases, int out_terms, void* out8, int out_C, int out_c0, void* out4,
                    void* out_sf, void* cuda_stream);

/*
 * GroupNorm (no ReLU) of a padded-flat fp16 hi/lo activation [2][B*(H+2)*(W+2)][C]: mlr_norm of the MLR model
 * normalises the concatenated encoder outputs (networks.py:421-439, 491-494), which are not the raw output of a
 * convolution.  Two launches: fp64 sums per (image, group) into `stats` ([B][C/group_ch][2], zeroed by the caller),
 * then the normalised hi / lo (out_terms 2) and optional e4m3 planes.  Any group size that divides C.
 */
int cl_pf_groupnorm(const void* in, int64_t in_lo_rows, int B, int H, int W, int C, int group_ch, const float* gamma,
                    const float* beta, float eps, double* stats, void* out, int out_terms, void* out8,
                    void* cuda_stream);

/*
 * Stem: conv3x3 s1 (Cin = 1 or 3 -> 32) + per-channel GroupNorm(32, 32) + ReLU, written as the
 * four-phase fp16 PF input of conv2.  Two passes over the image (statistics, then recompute + store):
 * the 32-channel full-resolution fp32 tensor the reference materialises is never written.
 * Replaces encoder.conv1 / norm1 (networks.py:186-190, 231) and Network.conv1 (:59, 96; has_gn = 0).
 *   image NCHW fp32 [B][Cin][H][W]; weight OIHW fp32 [32][Cin][3][3]; stats fp64 [B][32][2] zeroed by caller.
 *   raw_out (nullable, training): the raw conv1 output incl. bias as fp32 PF [B*(H+2)*(W+2)][32], interior rows.
 */
int cl_stem_forward(const float* image, int B, int Cin, int H, int W, const float* weight, const float* bias,
                    int has_gn, double* stats, const float* gamma, const float* beta, float eps, void* out,
                    int out_terms, float* raw_out, void* cuda_stream);

/*
 * Output head: 1x1 convolution C -> Co (Co <= 8) + mean offset on the task channels +
 * exp(clamp(x, clamp_lo, clamp_hi)) on the remaining (uncertainty) channels; NCHW fp32 output.
 * Replaces decoder.fc3 and the output maps (networks.py:349-358; Network.fc3 :124-128).
 */
int cl_head_forward(const void* act, int64_t act_lo_rows, int in_terms, int B, int H, int W, int C, int Co,
                    const float* weight, const float* bias, const float* mean, int num_task, float clamp_lo,
                    float clamp_hi, float* out, void* cuda_stream);

/*
 * Full-size output head: GroupNorm + ReLU of the DUC convolution's raw output, PixelShuffle(rate), bilinear
 * resize (align_corners = False) to (Ho, Wo), 1x1 fc3 (Co -> Co), mean offset on the task channels and
 * exp(clamp()) on the rest; NCHW fp32 [B][Co][Ho][Wo] output.  The shuffled map is never materialised.
 * Replaces DenseUpsamplingConvolution.forward after its convolution (networks.py:269-272), F.interpolate
 * (:347) and fc3 / the output maps (:349-358) of the full_size_output decoder.
 *   raw fp32 PF [B*(Hc+2)*(Wc+2)][C] with C = Co * rate^2 (cl_conv_igemm output), stats fp64 [B][C/group_ch][2].
 */
int cl_duc_head_forward(const float* raw, int B, int Hc, int Wc, int C, int Co, int rate, int group_ch,
                        const double* stats, const float* gamma, const float* beta, float eps, const float* weight,
                        const float* bias, const float* mean, int num_task, float clamp_lo, float clamp_hi,
                        float* out, int Ho, int Wo, void* cuda_stream);

/*
 * Input frames on the device: uint8 HWC [B][H][W][C] (decoder / PIL layout) -> fp32 NCHW [B][C][H][W] as
 * x / 255, then (v - mean[c]) / std[c] when mean / std (device, [C]) are given -- the fp32 operations of
 * torchvision's ToTensor and Normalize (dataloader/dataloader.py:189-212), bit for bit.  The host-to-device copy of
 * a frame shrinks from 4 bytes to 1 byte per sample.  Frames must already have the network resolution (the
 * reference's Resize(480) is the identity for its 480 x 720 datasets).
 */
int cl_frames_to_nchw(const uint8_t* frames, int B, int H, int W, int C, const float* mean, const float* stdv, float* out,
                      void* cuda_stream);

/*
 * ---- Input frames ---------------------------------------------------------------
 * The image path of the reference's data loader, /root/reference/dataloader/dataloader.py:306-316 (io.imread, gray2rgb,
 * RGBA -> RGB) and :189-212 (ToPILImage -> Resize(image_height) -> ToTensor [-> Normalize]), without the DataLoader
 * workers: PNG files are decoded on host threads straight into a (pinned) uint8 batch, Resize runs on the device and
 * cl_frames_to_nchw / cl_net_forward_frames do ToTensor / Normalize.
 *   cl_png_info          size of a PNG file image (and the channel count stored in the file)
 *   cl_decode_png        8-bit gray / gray+alpha / RGB / RGBA / palette, non-interlaced -> uint8 [H][W][3]
 *   cl_decode_png_batch  `count` files of one size -> uint8 [count][H][W][3] on `threads` host threads
 *   cl_resize_frames     Pillow's antialiased bilinear resample (what torchvision Resize does to a PIL image), uint8
 *                        [B][Hin][Win][3] -> [B][Hout][Wout][3] on the device, bit-identical to Image.resize(BILINEAR);
 *                        workspace: device memory of cl_resize_workspace_bytes() bytes
 *   cl_resize_coeffs     the fixed-point coefficient table of one axis (host only; bounds / kk may be NULL to query ksize)
 */
int cl_png_info(const void* file, size_t file_bytes, int* height, int* width, int* file_channels);
int cl_decode_png(const void* file, size_t file_bytes, uint8_t* out_rgb, int height, int width);
int cl_decode_png_batch(const void* const* files, const size_t* file_bytes, int count, uint8_t* out_rgb, int height, int width,
                        int threads);
int cl_resize_frames(const uint8_t* src, int B, int Hin, int Win, uint8_t* dst, int Hout, int Wout, void* workspace,
                     size_t workspace_bytes, void* cuda_stream);
size_t cl_resize_workspace_bytes(int B, int Hin, int Win, int Hout, int Wout);
int cl_resize_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk);

/*
 * ---- Whole-network runtime -------------------------------------------------------------------------------
 * `network(image)` of the reference's evaluation loop (/root/reference/test_single_task.py:347; the module is built
 * by utils/evaluation.py:106-116 as networks.TransPoseNet / Network, /root/reference/networks/networks.py:363-502,
 * :43-130) as ONE call: cl_net_create() takes the module's layer table (state-dict tensors as plain pointers, host or
 * device), packs the filters for the tensor cores and owns every workspace; cl_net_forward() runs stem -> strided
 * ladder -> residual blocks -> head for a batch of frames.  The layer plan of a (B, H, W), its TMA tensor maps and
 * buffers are built on first use, cached (three sizes) and replayed as a CUDA graph.
 *
 * A handle is bound to the device that was current at creation and serves one forward at a time (calls on one handle
 * are serialised by an internal lock; use one handle per concurrent stream).  Different handles are independent.
 */
#define CL_NET_ABI_VERSION 1
#define CL_NET_MAX_BLOCK_CONVS 4
#define CL_BLOCK_RESIDUAL 0       /* res = [relu](res + chain(res))                       networks.py:236-240, 251-254 */
#define CL_BLOCK_RESIDUAL_SKIP 1  /* res = [relu](skip_norm(skip(res)) + chain(res))      networks.py:242-249          */
#define CL_BLOCK_PLAIN 2          /* res = chain(res) (fc1, fc2)                          networks.py:342-343          */

typedef struct cl_net cl_net;

typedef struct cl_net_layer {     /* one nn.Conv2d [+ nn.GroupNorm] [+ ReLU] */
    int32_t cin, cout, ksize, stride;   /* 3x3 (padding 1) or 1x1, stride 1 or 2 */
    const float* weight;          /* OIHW fp32 (the state-dict tensor) */
    const float* bias;            /* [cout] or NULL */
    int32_t gn_groups;            /* 0: no GroupNorm after this convolution */
    const float* gn_weight;       /* [cout] */
    const float* gn_bias;         /* [cout] */
    float gn_eps;
} cl_net_layer;

typedef struct cl_net_block {
    int32_t kind;                 /* CL_BLOCK_* */
    int32_t n_convs;
    int32_t convs[CL_NET_MAX_BLOCK_CONVS];   /* indices into the layer table, in execution order */
    int32_t skip;                 /* CL_BLOCK_RESIDUAL_SKIP: the 1x1 skip convolution (+ its GroupNorm) */
} cl_net_block;

typedef struct cl_net_desc {
    int32_t abi_version;          /* CL_NET_ABI_VERSION */
    int32_t precision;            /* 1 = one fp16 pass (misses the 1e-3 bar), 2 = fp16 + e4m3 corrections, 3 = fp16x3, 4 = fp16 + block-scaled e2m1 corrections */
    int32_t relu_after_add;       /* 1: TransPoseNet (ReLU after every residual add), 0: vanilla Network */
    int32_t n_layers;
    const cl_net_layer* layers;
    int32_t stem[4];              /* conv1 .. conv4 (3x3; strides 1, 2, 2, 2) */
    int32_t n_blocks;
    const cl_net_block* blocks;
    int32_t head_layer;           /* fc3: 1x1, at most 8 output channels */
    const float* head_mean;       /* [num_task] added to the task channels */
    int32_t num_task;             /* output channels >= num_task get exp(clamp(x, clamp_lo, clamp_hi)) */
    float clamp_lo, clamp_hi;
    int32_t duc_layer;            /* full-size variant: the DUC 3x3 convolution (networks.py:259-273), else -1 */
    int32_t duc_rate;             /* PixelShuffle factor (8) */
} cl_net_desc;

int cl_net_create(const cl_net_desc* desc, cl_net** out);
/* Re-reads every parameter tensor named at creation (they must still be alive at the same addresses) and repacks the
 * filters: call after load_state_dict() / an optimizer step.  Plans and graphs stay valid. */
int cl_net_update(cl_net* net);
/* image: fp32 NCHW [B][Cin][H][W], host or device.  out: fp32 NCHW [B][Co][Ho][Wo] (cl_net_output_shape), host or
 * device; NULL leaves the result in the plan's output buffer (cl_net_buffers).  Stream-ordered; synchronises only when
 * `out` is host memory. */
int cl_net_forward(cl_net* net, const float* image, int B, int H, int W, float* out, void* cuda_stream);
/* Same for uint8 HWC frames [B][H][W][Cin] (what the decoders deliver; a quarter of the copy): the device applies
 * torchvision's ToTensor [+ Normalize(mean, std), both [Cin] or both NULL] bit for bit (dataloader.py:189-212). */
int cl_net_forward_frames(cl_net* net, const uint8_t* frames, int B, int H, int W, const float* mean, const float* stdv,
                          float* out, void* cuda_stream);
/* Makes `cuda_stream` wait until the most recently enqueued forward of this handle has finished its stem and strided
 * ladder and enters the tensor-core-bound residual blocks: work queued on that stream afterwards (the previous batch's
 * pose solve: fp64 CUDA-core kernels) then shares the SMs with convolutions that leave the CUDA cores idle. */
int cl_net_wait_fork(cl_net* net, void* cuda_stream);
int cl_net_output_shape(cl_net* net, int B, int H, int W, int* out_c, int* out_h, int* out_w);
/* Plan-owned device buffers of a (B, H, W): a caller may fill `in_f32` / `in_u8` itself and pass that pointer to the
 * forward call (no staging copy), and read the result from `out_buf`.  `launches` = kernels of one forward. */
int cl_net_buffers(cl_net* net, int B, int H, int W, float** in_f32, uint8_t** in_u8, float** out_buf, int* launches);
/* Per-op device times.  enable = 1 switches the handle to eager launches with a CUDA event between ops (no graph);
 * a later call with tables returns, for the plan of (B, H, W), the op list (kinds: 0 memset, 1 stem, 2 convolution,
 * 3 GroupNorm apply, 4 head, 5 full-size head, 6 statistics; labels [n][4]: convolution cin, cout, ksize, stride /
 * apply channels, output phases, merge kind), the algorithmic FLOPs per launch and the milliseconds summed over the
 * `forwards` recorded passes, then clears the record.  Returns the number of ops (negative on error). */
int cl_net_profile(cl_net* net, int enable, int B, int H, int W, int max_ops, int32_t* kinds, int32_t* labels,
                   double* flops, float* ms, int* forwards);
void cl_net_destroy(cl_net* net);

#ifdef __cplusplus
}
#endif
#endif /* CROSSLOC_B200_H */
