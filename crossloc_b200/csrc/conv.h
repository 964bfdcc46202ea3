// Launch interfaces of the CNN kernels (device pointers only).
//
// Activation layout ("padded-flat", PF): fp16 matrix [rows][C], row(t, ph, b, y, x) =
//   ((t * P + ph) * B + b) * (H + 2) * (W + 2) + (y + 1) * (W + 2) + (x + 1)
// with t = 0 (hi) / 1 (lo) fp16 split term, ph = parity phase (P = 1, or 4 when the consumer is a
// stride-2 convolution: ph = (y_in & 1) * 2 + (x_in & 1), (y, x) = (y_in / 2, x_in / 2)), zero borders.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace cl {

// power-of-two scales of the e4m3 operand planes (fp16 + fp8 convolution mode)
constexpr float kAct8HiScale = 4.0f;        // fp8(a_hi * 2^2)
constexpr float kAct8LoScale = 16384.0f;    // fp8((a - a_hi) * 2^14)
constexpr float kW8LoScale = 4096.0f;       // fp8(w_lo * 2^12)   (w_hi is stored unscaled)
constexpr float kCorrScale = 1.0f / 16384.0f;   // both correction products carry 2^14

// ---------------------------------------------------------------- tcgen05 implicit GEMM
struct ConvIgemmDesc {
    const void* act;        // fp16 PF matrix, all planes
    int64_t a_total_rows;   // rows of the whole activation matrix (bounds of the TMA map)
    int64_t a_lo_rows;      // row distance between the hi and the lo plane
    int Cin;
    const void* weights;    // fp16 [term][tap][Cout][Cin]
    int Cout;
    int num_taps;
    int tap_a_row[9];       // activation row shift of every tap (phase offset included)
    int nterms;             // 1 (single fp16 pass), 2 (fp16 + e4m3 corrections), 3 (fp16x3 split) or 4 (fp16 + block-scaled e2m1 corrections)
    const void* act8;       // nterms == 2: e4m3 PF matrix, plane 0 = fp8(a * 2^2), plane 1 = fp8((a - a_hi) * 2^14)
    int64_t a8_total_rows, a8_lo_rows;
    const void* weights8;   // nterms == 2: e4m3 [2][tap][Cout][Cin]: fp8(w_hi), fp8(w_lo * 2^12)
    float corr_scale;       // 2^-14: scale of the correction accumulator
    // nterms == 4 (CTA-pair kernel, Cin % 256 == 0): the two correction products as kind::mxf4 MMAs (4x the fp16 rate).
    const void* act4;       // e2m1 PF matrix [2][a4_lo_rows][Cin / 2] (two values per byte, even channel in the low nibble):
    int64_t a4_total_rows, a4_lo_rows;   // plane 0 = fp4(a_hi / 2^s_hi), plane 1 = fp4((a - a_hi) / 2^s_lo)
    const uint32_t* act_sf; // [Cin / 256][a4_lo_rows] ue8m0 scales of one pixel row and 256-channel block: bytes (lo, lo, hi, hi)
    const void* weights4;   // e2m1 [2][tap][Cout][Cin / 2]: fp4(w_hi), fp4(w_lo), scaled per (tap, output channel, 256-channel block)
    const uint32_t* w_sf;   // [tap][Cin / 256][Cout / 128][128]: bytes (hi, hi, lo, lo); word l * 4 + j belongs to output channel 32 j + l
    int cluster;            // CTAs sharing a weight tile by TMA multicast (0 = default 2; 1, 2 or 4)
    int Mp, Hp, Wp;         // output rows (B * Hp * Wp) and padded plane size
    int group_ch;           // GroupNorm channels per group (0: no statistics)
    float out_scale;        // undoes the power-of-two weight pre-scale
    float* raw;             // fp32 [Mp][Cout] (interior rows only are written); unused when `fuse` is set
    const float* bias;      // fp32 [Cout]
    double* stats;          // fp64 [B][groups][2] sum, sum of squares (accumulated, caller zeroes)
    // ---- optional: dynamic tile scheduling (CTA-pair kernel).  Clusters fetch tiles from a global counter instead of a
    // fixed stride, so a cluster that starts late (SMs shared with another stream's kernels) simply takes fewer tiles.
    int* tile_counter;      // nullable device int, zeroed by the caller before every launch
    // ---- optional: GroupNorm + ReLU (+ residual add + ReLU) fused into the epilogue (CTA-pair kernel, needs tile_counter).
    // The accumulator of a tile stays in tensor memory until the statistics of its image(s) are complete across the
    // grid; the epilogue then normalises straight from TMEM and writes the next layer's operand planes -- the fp32 raw
    // tensor and the separate gn_apply pass (write 4 B + read 4 B per element) disappear.
    int fuse;
    const float* gamma;     // [Cout] (group_ch != 0)
    const float* beta;
    float eps;
    int H, W;               // interior size (statistics count = group_ch * H * W)
    int relu_inner, relu_outer;
    const __half* res;      // nullable residual stream, fp16 PF [Mp][Cout] hi plane, lo plane res_lo_rows further (0: none)
    int64_t res_lo_rows;
    __half* out16;          // fp16 PF [out_terms][Mp][Cout]: hi (and lo) operand planes of the consumer
    int out_terms;
    uint8_t* out8;          // nullable e4m3 planes [2][Mp][Cout]
    int* unit_done;         // [Cout / BN][B] ints, zeroed by the caller: warp slices that have published their statistics
};

struct ConvIgemmParams {
    int num_taps;
    int tap_a_row[9];
    int kblocks_per_tap, nterms, a_lo_rows, a8_lo_rows, w_tap_rows, w_lo_rows;
    float corr_scale;
    int Mp, Cout, BN, tiles_m, tiles_n, Hp, Wp, group_ch, groups;
    int cluster, super_m;
    float out_scale;
    float* raw;
    const float* bias;
    double* stats;
    int num_stages, accum_stages;
    uint32_t a_bytes, w_bytes, stage_bytes;
    int* tile_counter;
    // fp16 + fp4 mode
    int a4_lo_rows, kgroups;
    int kw_share;           // 3x3 stride 1 (+1 / -1 = rows ascend / descend with kw, 0 = off): pass 1 loads one 136-row activation
                            // tile per (kh, k-block) and reads it at three row shifts
    int kw_share0;          // the same for pass 0 (e2m1 planes): a stage holds one plane's 136-row tile and three weight half-tiles
    const uint32_t* act_sf;
    const uint32_t* w_sf;
    // fused GroupNorm epilogue
    int fuse, H, W, B, relu_inner, relu_outer, out_terms, has_out8;
    const float* gamma;
    const float* beta;
    float eps;
    const __half* res;
    long long res_lo_rows;
    int* unit_done;
    __half* out16;
    uint8_t* out8;
};

// returns nullptr on success, else a static error string
const char* conv_igemm_launch(const ConvIgemmDesc& d, cudaStream_t stream);

// The same launch split in two: everything that depends only on shapes and pointers (tile schedule, the five TMA
// tensor maps: ~10 us of driver calls) is prepared once and replayed (csrc/net.cu caches one plan per layer).
struct alignas(64) ConvOutMaps {   // TMA store maps of the fused epilogue: fp16 hi / lo planes, e4m3 hi / lo planes
    CUtensorMap hi, lo, hi8, lo8;
};
struct alignas(64) ConvIgemmPlan {
    CUtensorMap tmA, tmW, tmO, tmA8, tmW8;
    ConvOutMaps out;
    ConvIgemmParams p;
    size_t smem;
    int grid, cluster;
    int variant;            // bit 2: fused GroupNorm epilogue, bit 1: CTA-pair kernel, bit 0: 64-channel k-blocks; 8: CTA pairs with fp4 corrections
};
const char* conv_igemm_prepare(const ConvIgemmDesc& d, ConvIgemmPlan* plan);
const char* conv_igemm_run(const ConvIgemmPlan& plan, cudaStream_t stream);

// ---------------------------------------------------------------- weight gradient on padded-flat operands
// dW[tap][co][ci] += out_scale * sum over PF rows r of dY[r][co] * X[tap_phase][r + tap_shift][ci]; both operands
// are the [pixel rows][channels] matrices of the forward / GroupNorm-backward kernels (MN-major tensor-core operands).
struct ConvWgradPfDesc {
    const void* grad;       // dY fp16 PF [nA][g_plane_rows][Cout]: plane 0 = hi, plane 1 = lo; border rows are zero
    int64_t g_plane_rows;   // row distance between the planes of dY
    const void* act;        // X fp16 PF, plane (term * phases + phase) starts x_plane_rows * plane rows in
    int64_t x_plane_rows;
    int Mp;                 // rows of one plane: B * (H + 2) * (W + 2) at the OUTPUT resolution
    int Cout, Cin, phases;  // phases 1, or 4 for a stride-2 convolution
    int num_taps;
    int tap_shift[9];       // row shift of X inside its plane for each tap
    int tap_phase[9];       // phase plane of X for each tap
    int nterms;             // 1 or 3
    float out_scale;
    const float* scale_dev; // nullable device scalar multiplied into the result (the 2^-k of a rescaled gradient)
    int oihw;               // 0: dw is [tap][Cout][Cin];  1: [Cout][Cin][taps], i.e. torch's OIHW weight layout
    float* dw;              // fp32, accumulated with atomics: caller zeroes it
};
const char* conv_wgrad_pf_launch(const ConvWgradPfDesc& d, cudaStream_t stream);

// ---------------------------------------------------------------- layout kernels of the training path
struct NchwToPfDesc {
    const float* x;         // NCHW fp32 [B][C][H][W]
    const float* scale;     // nullable device scalar multiplied in
    __half* out;            // PF hi/lo [2][phases][B*(H'+2)*(W'+2)][C], zero-initialised by the caller
    int B, C, H, W, phases;
};
const char* nchw_to_pf_launch(const NchwToPfDesc& d, cudaStream_t stream);

struct PfToNchwDesc {
    const float* raw;       // fp32 PF [B*(H+2)*(W+2)][Craw]
    int B, H, W, Craw;
    float* out;             // NCHW fp32 [B][C][Hout][Wout]; pixel (y, x) of raw goes to (y*step + off_y, x*step + off_x)
    int C, Hout, Wout, step, off_y, off_x;
    const float* scale;     // nullable device scalar
    const float* bias;      // nullable [C]
};
const char* pf_to_nchw_launch(const PfToNchwDesc& d, cudaStream_t stream);

// {2^k, 2^-k} with k = floor(log2(target / max|x|)); scratch: two 32-bit words (zeroed by the launch), out: two floats
const char* pow2_scale_launch(const float* x, size_t n, float target, unsigned* scratch, float* out, cudaStream_t stream);

struct PackFilterDesc {
    const float* w;         // OIHW fp32 [Cout][Cin][k][k]
    const float* scale;     // device scalar (power of two)
    __half* out;            // hi/lo [2][num_taps][N][K]
    int Cout, Cin, ksize;
    int num_taps;
    int tap_kh[9], tap_kw[9];
    int transpose;          // 0: N = Cout, K = Cin (forward);  1: N = Cin (zero-padded to N), K = Cout (data gradient)
    int N, K;
};
const char* pack_filter_launch(const PackFilterDesc& d, cudaStream_t stream);

// ---------------------------------------------------------------- GroupNorm apply / residual merge
struct GnApplyDesc {
    const float* raw;       // fp32 PF [B*(H+2)*(W+2)][C]
    int B, H, W, C;
    int group_ch;           // 0: no normalisation
    const double* stats;    // [B][C/group_ch][2]
    const float* gamma;
    const float* beta;
    float eps;
    int relu_inner;         // ReLU right after the normalisation
    int add_kind;           // 0 none | 1 fp16 hi/lo residual (same geometry, P = 1) | 2 second raw tensor with its own GroupNorm
    const __half* res;      // add_kind 1: PF matrix, lo plane res_lo_rows rows further (0: no lo plane)
    int64_t res_lo_rows;
    const float* raw2;      // add_kind 2
    const double* stats2;
    const float* gamma2;
    const float* beta2;
    int relu_outer;         // ReLU after the add
    __half* out;            // PF matrix at (H, W) for out_phases == 1, at (ceil(H/2), ceil(W/2)) x 4 phases otherwise
    int out_phases;
    int out_terms;          // 2: write hi and lo, 1: hi only
    uint8_t* out8;          // nullable: e4m3 planes [2][B*(H+2)*(W+2)][C] for a consumer in fp16 + fp8 mode
    int out_C;              // channels of the destination matrices (>= C: the output may be a channel slice), 0 = C
    int out_c0;             // first destination channel
    // nullable: block-scaled e2m1 planes for a consumer in fp16 + fp4 mode (out_phases == 1, C % 256 == 0, no channel slice):
    uint8_t* out4;          // [2][B*(H+2)*(W+2)][C / 2]: fp4(hi / 2^s_hi), fp4((x - hi) / 2^s_lo), even channel in the low nibble
    uint32_t* out_sf;       // [C / 256][B*(H+2)*(W+2)]: ue8m0 bytes (s_lo, s_lo, s_hi, s_hi) of the pixel's 256-channel block
};
const char* gn_apply_launch(const GnApplyDesc& d, cudaStream_t stream);

// OIHW fp32 filter x scale -> e2m1 planes [2][tap][Cout][Cin / 2] (fp4(w_hi), fp4(w_lo)) and their scale words
// [tap][Cin / 256][Cout / 128][128] as ConvIgemmDesc::weights4 / w_sf expect them
const char* pack_conv_fp4_launch(const float* w, int Cout, int Cin, int taps, float scale, uint8_t* w4, uint32_t* w_sf, cudaStream_t stream);

// ---------------------------------------------------------------- GroupNorm / ReLU / residual-merge backward (training)
struct GnBwdSrc {
    const float* g;         // fp32 PF gradient [rows][stride] (interior rows valid)
    const float* scale_a;   // nullable device scalars multiplied into this source
    const float* scale_b;
    int stride;             // floats per row (>= C)
    int phased;             // 1: stored as 4 parity phases at half resolution (data gradient of a stride-2 convolution)
};
struct GnBwdDesc {
    int B, H, W, C;
    int group_ch;           // 0: no normalisation (vanilla Network)
    const float* raw;       // fp32 PF raw convolution output of this layer
    const double* stats;    // [B][C/group_ch][2]
    const float* gamma;
    const float* beta;
    float eps;
    int relu_inner;
    int num_src;            // pass 1 sums src[0 .. num_src); pass 2 reads src[0] only
    GnBwdSrc src[3];
    const __half* mask_out; // pass 1, nullable: hi plane of the stage's merged output; the gradient passes where it is > 0
    float* g_out;           // pass 1, nullable: summed / masked gradient, fp32 PF [rows][C]
    double* ab;             // [B][ab_C][2]: sum dy, sum dy * xhat (pass 1 accumulates, pass 2 reads); caller zeroes
    int ab_C;               // channels per image in `ab` (0 = C): several stages can share one buffer, each at its own offset
    unsigned* gmax_bits;    // float bits of max |dy * gamma| * rstd (pass 1 atomicMax, pass 2 reads); caller zeroes
    __half* d_raw;          // pass 2: gradient of the raw convolution output, fp16 hi / lo PF planes x 2^k
    int64_t d_raw_lo_rows;
    float* scale_out;       // pass 2: {2^k, 2^-k}
    double* dbias;          // pass 2, nullable: [C] += sum d_raw (gradient of the convolution bias)
    float* d_raw_f32;       // pass 2, nullable: the same gradient unscaled in fp32 PF [rows][C] (stem: weight gradient in torch)
    // pass 2, nullable (C % 256 == 0): block-scaled e2m1 planes of the scaled gradient for a data gradient in fp16 + fp4 mode,
    // in the layout cl_gn_apply_fp4 writes for the forward: [2][d_raw4_lo_rows][C/2] and scale words [C/256][d_raw4_lo_rows]
    uint8_t* d_raw4;
    int64_t d_raw4_lo_rows;
    uint32_t* d_raw_sf;
};
const char* gn_bwd_reduce_launch(const GnBwdDesc& d, cudaStream_t stream);

// backward of the 1x1 output head on the padded-flat activation it read
struct HeadBwdDesc {
    const __half* act;      // PF fp16 hi / lo [2][B*(H+2)*(W+2)][C] input of the head
    int64_t act_lo_rows;
    int B, H, W, C, Co;
    const float* weight;    // fp32 [Co][C]
    const float* g_sc;      // gradient of the head's pre-activation output, NCHW fp32 [B][Co][H][W]
    float* g_x;             // out: fp32 PF [B*(H+2)*(W+2)][C] (interior rows)
    float* g_w;             // out: fp32 [Co][C], accumulated with atomics (caller zeroes)
};
const char* head_bwd_launch(const HeadBwdDesc& d, cudaStream_t stream);
const char* gn_bwd_apply_launch(const GnBwdDesc& d, cudaStream_t stream);

// ---------------------------------------------------------------- stem convolution (3x3, stride 1, Cin in {1, 3}, Cout = 32)
struct StemDesc {
    const float* image;     // NCHW fp32 [B][Cin][H][W]
    int B, Cin, H, W;
    const float* weight;    // OIHW fp32 [32][Cin][3][3]
    const float* bias;      // [32]
    int has_gn;             // per-channel GroupNorm(32, 32)
    double* stats;          // [B][32][2]
    const float* gamma;
    const float* beta;
    float eps;
    __half* out;            // PF, 4 phases at (ceil(H/2), ceil(W/2)), C = 32
    int out_terms;
    float* raw_out;         // nullable (tensor-core version only): raw conv1 output, fp32 PF [B*(H+2)*(W+2)][32], for training
};
// tensor-core version (stem_tc.cu): im2col rows built in shared memory, tcgen05 fp16x3; same two passes
const char* stem_tc_launch(const StemDesc& d, bool stats_pass, cudaStream_t stream);
// CUDA-core version (cnn_pointwise.cu), CROSSLOC_B200_STEM=cuda
const char* stem_stats_launch(const StemDesc& d, cudaStream_t stream);   // pass 1: statistics only
const char* stem_apply_launch(const StemDesc& d, cudaStream_t stream);   // pass 2: recompute, normalise, ReLU, store

// ---------------------------------------------------------------- 1x1 head (C -> Co <= 8) with the decoder's output maps
struct HeadDesc {
    const __half* act;      // PF P = 1 at (H, W), C channels, lo plane act_lo_rows further
    int64_t act_lo_rows;
    int in_terms;
    int B, H, W, C, Co;
    const float* weight;    // fp32 [Co][C]
    const float* bias;      // [Co]
    const float* mean;      // [num_task] added to the task channels
    int num_task;           // channels >= num_task get exp(clamp(x, lo, hi))
    float clamp_lo, clamp_hi;
    float* out;             // NCHW fp32 [B][Co][H][W]
};
const char* head_launch(const HeadDesc& d, cudaStream_t stream);

// ---------------------------------------------------------------- full-size head: GN + ReLU + PixelShuffle + bilinear + fc3
struct DucHeadDesc {
    const float* raw;       // fp32 PF [B*(Hc+2)*(Wc+2)][C] raw output of the DUC convolution, C = Co * rate^2
    int B, Hc, Wc, C, Co, rate;
    int group_ch;           // channels per GroupNorm group of the DUC norm
    const double* stats;    // [B][C/group_ch][2]
    const float* gamma;
    const float* beta;
    float eps;
    const float* weight;    // fc3 fp32 [Co][Co]
    const float* bias;      // [Co]
    const float* mean;      // [num_task]
    int num_task;
    float clamp_lo, clamp_hi;
    float* out;             // NCHW fp32 [B][Co][Ho][Wo]
    int Ho, Wo;             // frame size the shuffled (Hc*rate, Wc*rate) map is resized to
};
const char* duc_head_launch(const DucHeadDesc& d, cudaStream_t stream);

// ---------------------------------------------------------------- input frames: uint8 HWC -> fp32 NCHW (ToTensor [+ Normalize])
const char* frames_to_nchw_launch(const uint8_t* frames, int B, int H, int W, int C, const float* mean, const float* stdv,
                                  float* out, cudaStream_t stream);

// ---------------------------------------------------------------- GroupNorm of a padded-flat activation (MLR merge)
struct PfGroupNormDesc {
    const __half* in;       // fp16 PF [2][B*(H+2)*(W+2)][C] hi / lo planes (P = 1)
    int64_t in_lo_rows;
    int B, H, W, C;
    int group_ch;
    const float* gamma;
    const float* beta;
    float eps;
    double* stats;          // [B][C/group_ch][2], zeroed by the caller
    __half* out;            // same geometry, normalised (no ReLU)
    int out_terms;
    uint8_t* out8;          // nullable e4m3 planes
};
const char* pf_groupnorm_launch(const PfGroupNormDesc& d, cudaStream_t stream);

}  // namespace cl
