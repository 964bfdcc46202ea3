// extern "C" entry points of the CNN operators declared in include/crossloc_b200.h.
#include "../../include/crossloc_b200.h"

#include <cstdlib>

#include "cabi_common.h"
#include "conv.h"

namespace {

int need_device(const char* fn, const char* name, const void* p, bool nullable = false)
{
    if (!p) return nullable ? 0 : cl::fail(-1, "%s: %s must not be NULL", fn, name);
    if (!cl::is_device_ptr(p)) return cl::fail(-1, "%s: %s must be a device pointer", fn, name);
    return 0;
}

int finish(const char* fn, const char* err)
{
    if (err) return cl::fail(-2, "%s: %s", fn, err);
    return 0;
}

}  // namespace

#define NEED_DEV(name, ...)                                                \
    do {                                                                   \
        if (int _rc = need_device(kFn, #name, name, ##__VA_ARGS__)) return _rc; \
    } while (0)

extern "C" int cl_conv_igemm(const void* act, int64_t a_total_rows, int64_t a_lo_rows, int Cin, const void* weights,
                             int Cout, int num_taps, const int32_t* tap_a_row, int nterms, int Mp, int Hp, int Wp,
                             int group_ch, float out_scale, float* raw, const float* bias, double* stats,
                             const void* act8, int64_t a8_total_rows, int64_t a8_lo_rows, const void* weights8,
                             void* cuda_stream)
{
    static const char* kFn = "cl_conv_igemm";
    NEED_DEV(act); NEED_DEV(weights); NEED_DEV(raw); NEED_DEV(bias);
    if (nterms == 2) { NEED_DEV(act8); NEED_DEV(weights8); }
    if (group_ch) NEED_DEV(stats);
    if (!tap_a_row) return cl::fail(-1, "%s: tap_a_row must not be NULL", kFn);
    if (num_taps < 1 || num_taps > 9) return cl::fail(-1, "%s: num_taps=%d out of range", kFn, num_taps);
    if (Mp <= 0 || Hp < 3 || Wp < 3 || Mp % (Hp * Wp) != 0)
        return cl::fail(-1, "%s: Mp=%d is not a whole number of %dx%d planes", kFn, Mp, Hp, Wp);
    cl::ConvIgemmDesc d{};
    d.act = act; d.a_total_rows = a_total_rows; d.a_lo_rows = a_lo_rows; d.Cin = Cin; d.weights = weights;
    d.Cout = Cout; d.num_taps = num_taps;
    for (int i = 0; i < num_taps; i++) d.tap_a_row[i] = tap_a_row[i];
    d.nterms = nterms; d.Mp = Mp; d.Hp = Hp; d.Wp = Wp; d.group_ch = group_ch; d.out_scale = out_scale;
    d.raw = raw; d.bias = bias; d.stats = stats;
    d.act8 = act8; d.a8_total_rows = a8_total_rows; d.a8_lo_rows = a8_lo_rows; d.weights8 = weights8;
    d.corr_scale = cl::kCorrScale;
    {
        const char* env = getenv("CROSSLOC_B200_CONV_CLUSTER");   // tuning knob: 1, 2 or 4 CTAs share a weight tile
        d.cluster = env ? atoi(env) : 0;
    }
    return finish(kFn, cl::conv_igemm_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_conv_igemm_fp4(const void* act, int64_t a_total_rows, int64_t a_lo_rows, int Cin, const void* weights,
                                 int Cout, int num_taps, const int32_t* tap_a_row, int Mp, int Hp, int Wp, int group_ch,
                                 float out_scale, float* raw, const float* bias, double* stats, const void* act4,
                                 int64_t a4_total_rows, int64_t a4_lo_rows, const void* act_sf, const void* weights4,
                                 const void* w_sf, void* cuda_stream)
{
    static const char* kFn = "cl_conv_igemm_fp4";
    NEED_DEV(act); NEED_DEV(weights); NEED_DEV(raw); NEED_DEV(bias);
    NEED_DEV(act4); NEED_DEV(act_sf); NEED_DEV(weights4); NEED_DEV(w_sf);
    if (group_ch) NEED_DEV(stats);
    if (!tap_a_row) return cl::fail(-1, "%s: tap_a_row must not be NULL", kFn);
    if (num_taps < 1 || num_taps > 9) return cl::fail(-1, "%s: num_taps=%d out of range", kFn, num_taps);
    if (Mp <= 0 || Hp < 3 || Wp < 3 || Mp % (Hp * Wp) != 0)
        return cl::fail(-1, "%s: Mp=%d is not a whole number of %dx%d planes", kFn, Mp, Hp, Wp);
    cl::ConvIgemmDesc d{};
    d.act = act; d.a_total_rows = a_total_rows; d.a_lo_rows = a_lo_rows; d.Cin = Cin; d.weights = weights;
    d.Cout = Cout; d.num_taps = num_taps;
    for (int i = 0; i < num_taps; i++) d.tap_a_row[i] = tap_a_row[i];
    d.nterms = 4; d.Mp = Mp; d.Hp = Hp; d.Wp = Wp; d.group_ch = group_ch; d.out_scale = out_scale;
    d.raw = raw; d.bias = bias; d.stats = stats;
    d.act4 = act4; d.a4_total_rows = a4_total_rows; d.a4_lo_rows = a4_lo_rows;
    d.act_sf = static_cast<const uint32_t*>(act_sf);
    d.weights4 = weights4; d.w_sf = static_cast<const uint32_t*>(w_sf);
    d.corr_scale = cl::kCorrScale;
    d.cluster = 2;   // this mode exists in the CTA-pair kernel only
    return finish(kFn, cl::conv_igemm_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_pack_conv_fp4(const float* w, int Cout, int Cin, int num_taps, float scale, void* weights4, void* w_sf,
                                void* cuda_stream)
{
    static const char* kFn = "cl_pack_conv_fp4";
    NEED_DEV(w); NEED_DEV(weights4); NEED_DEV(w_sf);
    return finish(kFn, cl::pack_conv_fp4_launch(w, Cout, Cin, num_taps, scale, static_cast<uint8_t*>(weights4),
                                                static_cast<uint32_t*>(w_sf), static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_nchw_to_pf(const float* x, const float* scale, void* out, int B, int C, int H, int W, int phases,
                             void* cuda_stream)
{
    static const char* kFn = "cl_nchw_to_pf";
    NEED_DEV(x); NEED_DEV(out);
    if (scale) NEED_DEV(scale);
    cl::NchwToPfDesc d{x, scale, static_cast<__half*>(out), B, C, H, W, phases};
    return finish(kFn, cl::nchw_to_pf_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_pf_to_nchw(const float* raw, int B, int H, int W, int Craw, float* out, int C, int Hout, int Wout,
                             int step, int off_y, int off_x, const float* scale, const float* bias, void* cuda_stream)
{
    static const char* kFn = "cl_pf_to_nchw";
    NEED_DEV(raw); NEED_DEV(out);
    if (scale) NEED_DEV(scale);
    if (bias) NEED_DEV(bias);
    if (step != 1 && step != 2) return cl::fail(-1, "%s: step must be 1 or 2", kFn);
    cl::PfToNchwDesc d{raw, B, H, W, Craw, out, C, Hout, Wout, step, off_y, off_x, scale, bias};
    return finish(kFn, cl::pf_to_nchw_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_pow2_scale(const float* x, int64_t n, float target, void* workspace, void* cuda_stream)
{
    static const char* kFn = "cl_pow2_scale";
    NEED_DEV(x); NEED_DEV(workspace);
    if (n <= 0) return cl::fail(-1, "%s: empty tensor", kFn);
    float* out = static_cast<float*>(workspace);
    return finish(kFn, cl::pow2_scale_launch(x, (size_t)n, target, reinterpret_cast<unsigned*>(out + 2), out,
                                             static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_pack_filter(const float* w, const float* scale, void* out, int Cout, int Cin, int ksize, int num_taps,
                              const int32_t* tap_kh, const int32_t* tap_kw, int transpose, int N, int K,
                              void* cuda_stream)
{
    static const char* kFn = "cl_pack_filter";
    NEED_DEV(w); NEED_DEV(scale); NEED_DEV(out);
    if (!tap_kh || !tap_kw || num_taps < 1 || num_taps > 9) return cl::fail(-1, "%s: invalid tap table", kFn);
    cl::PackFilterDesc d{};
    d.w = w; d.scale = scale; d.out = static_cast<__half*>(out); d.Cout = Cout; d.Cin = Cin; d.ksize = ksize;
    d.num_taps = num_taps;
    for (int t = 0; t < num_taps; t++) { d.tap_kh[t] = tap_kh[t]; d.tap_kw[t] = tap_kw[t]; }
    d.transpose = transpose; d.N = N; d.K = K;
    return finish(kFn, cl::pack_filter_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_gn_apply(const float* raw, int B, int H, int W, int C, int group_ch, const double* stats,
                           const float* gamma, const float* beta, float eps, int relu_inner, int add_kind,
                           const void* res, int64_t res_lo_rows, const float* raw2, const double* stats2,
                           const float* gamma2, const float* beta2, int relu_outer, void* out, int out_phases,
                           int out_terms, void* out8, int out_C, int out_c0, void* cuda_stream)
{
    return cl_gn_apply_fp4(raw, B, H, W, C, group_ch, stats, gamma, beta, eps, relu_inner, add_kind, res, res_lo_rows, raw2,
                           stats2, gamma2, beta2, relu_outer, out, out_phases, out_terms, out8, out_C, out_c0, nullptr,
                           nullptr, cuda_stream);
}

extern "C" int cl_gn_apply_fp4(const float* raw, int B, int H, int W, int C, int group_ch, const double* stats,
                               const float* gamma, const float* beta, float eps, int relu_inner, int add_kind,
                               const void* res, int64_t res_lo_rows, const float* raw2, const double* stats2,
                               const float* gamma2, const float* beta2, int relu_outer, void* out, int out_phases,
                               int out_terms, void* out8, int out_C, int out_c0, void* out4, void* out_sf,
                               void* cuda_stream)
{
    static const char* kFn = "cl_gn_apply";
    NEED_DEV(raw); NEED_DEV(out);
    if (group_ch) { NEED_DEV(stats); NEED_DEV(gamma); NEED_DEV(beta); }
    if (add_kind == 1) NEED_DEV(res);
    if (add_kind == 2) { NEED_DEV(raw2); NEED_DEV(stats2); NEED_DEV(gamma2); NEED_DEV(beta2); }
    if (add_kind < 0 || add_kind > 2) return cl::fail(-1, "%s: add_kind=%d", kFn, add_kind);
    if (group_ch && C % group_ch != 0) return cl::fail(-1, "%s: C=%d not divisible by group_ch=%d", kFn, C, group_ch);
    cl::GnApplyDesc d{};
    d.raw = raw; d.B = B; d.H = H; d.W = W; d.C = C; d.group_ch = group_ch; d.stats = stats; d.gamma = gamma;
    d.beta = beta; d.eps = eps; d.relu_inner = relu_inner; d.add_kind = add_kind;
    d.res = static_cast<const __half*>(res); d.res_lo_rows = res_lo_rows; d.raw2 = raw2; d.stats2 = stats2;
    d.gamma2 = gamma2; d.beta2 = beta2; d.relu_outer = relu_outer; d.out = static_cast<__half*>(out);
    d.out_phases = out_phases; d.out_terms = out_terms; d.out8 = static_cast<uint8_t*>(out8);
    d.out_C = out_C; d.out_c0 = out_c0;
    if (out8) NEED_DEV(out8);
    if (out4) { NEED_DEV(out4); NEED_DEV(out_sf); }
    d.out4 = static_cast<uint8_t*>(out4); d.out_sf = static_cast<uint32_t*>(out_sf);
    return finish(kFn, cl::gn_apply_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_stem_forward(const float* image, int B, int Cin, int H, int W, const float* weight,
                               const float* bias, int has_gn, double* stats, const float* gamma, const float* beta,
                               float eps, void* out, int out_terms, float* raw_out, void* cuda_stream)
{
    static const char* kFn = "cl_stem_forward";
    if (raw_out) NEED_DEV(raw_out);
    NEED_DEV(image); NEED_DEV(weight); NEED_DEV(bias); NEED_DEV(out);
    if (has_gn) { NEED_DEV(stats); NEED_DEV(gamma); NEED_DEV(beta); }
    cl::StemDesc d{};
    d.image = image; d.B = B; d.Cin = Cin; d.H = H; d.W = W; d.weight = weight; d.bias = bias; d.has_gn = has_gn;
    d.stats = stats; d.gamma = gamma; d.beta = beta; d.eps = eps; d.out = static_cast<__half*>(out);
    d.out_terms = out_terms; d.raw_out = raw_out;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const char* env = getenv("CROSSLOC_B200_STEM");
    if (env && env[0] == 'c' && !raw_out) {   // "cuda": the CUDA-core kernels, kept for comparison (inference only)
        if (has_gn)
            if (int rc = finish(kFn, cl::stem_stats_launch(d, s))) return rc;
        return finish(kFn, cl::stem_apply_launch(d, s));
    }
    if (has_gn)
        if (int rc = finish(kFn, cl::stem_tc_launch(d, true, s))) return rc;
    return finish(kFn, cl::stem_tc_launch(d, false, s));
}

extern "C" int cl_head_forward(const void* act, int64_t act_lo_rows, int in_terms, int B, int H, int W, int C,
                               int Co, const float* weight, const float* bias, const float* mean, int num_task,
                               float clamp_lo, float clamp_hi, float* out, void* cuda_stream)
{
    static const char* kFn = "cl_head_forward";
    NEED_DEV(act); NEED_DEV(weight); NEED_DEV(bias); NEED_DEV(out);
    if (num_task > 0) NEED_DEV(mean);
    cl::HeadDesc d{};
    d.act = static_cast<const __half*>(act); d.act_lo_rows = act_lo_rows; d.in_terms = in_terms; d.B = B; d.H = H;
    d.W = W; d.C = C; d.Co = Co; d.weight = weight; d.bias = bias; d.mean = mean; d.num_task = num_task;
    d.clamp_lo = clamp_lo; d.clamp_hi = clamp_hi; d.out = out;
    return finish(kFn, cl::head_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_duc_head_forward(const float* raw, int B, int Hc, int Wc, int C, int Co, int rate, int group_ch,
                                   const double* stats, const float* gamma, const float* beta, float eps,
                                   const float* weight, const float* bias, const float* mean, int num_task,
                                   float clamp_lo, float clamp_hi, float* out, int Ho, int Wo, void* cuda_stream)
{
    static const char* kFn = "cl_duc_head_forward";
    NEED_DEV(raw); NEED_DEV(stats); NEED_DEV(gamma); NEED_DEV(beta); NEED_DEV(weight); NEED_DEV(bias); NEED_DEV(out);
    if (num_task > 0) NEED_DEV(mean);
    cl::DucHeadDesc d{};
    d.raw = raw; d.B = B; d.Hc = Hc; d.Wc = Wc; d.C = C; d.Co = Co; d.rate = rate; d.group_ch = group_ch;
    d.stats = stats; d.gamma = gamma; d.beta = beta; d.eps = eps; d.weight = weight; d.bias = bias; d.mean = mean;
    d.num_task = num_task; d.clamp_lo = clamp_lo; d.clamp_hi = clamp_hi; d.out = out; d.Ho = Ho; d.Wo = Wo;
    return finish(kFn, cl::duc_head_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_conv_wgrad_pf(const void* grad, int64_t g_plane_rows, const void* act, int64_t x_plane_rows, int Mp,
                                int Cout, int Cin, int phases, int num_taps, const int32_t* tap_shift,
                                const int32_t* tap_phase, int nterms, float out_scale, const float* scale_dev, int oihw,
                                float* dw, void* cuda_stream)
{
    static const char* kFn = "cl_conv_wgrad_pf";
    NEED_DEV(grad); NEED_DEV(act); NEED_DEV(dw);
    if (scale_dev) NEED_DEV(scale_dev);
    if (!tap_shift || !tap_phase) return cl::fail(-1, "%s: tap tables must not be NULL", kFn);
    if (num_taps < 1 || num_taps > 9) return cl::fail(-1, "%s: num_taps=%d out of range", kFn, num_taps);
    if (Mp <= 0 || phases < 1 || phases > 4) return cl::fail(-1, "%s: invalid sizes", kFn);
    for (int i = 0; i < num_taps; i++)
        if (tap_phase[i] < 0 || tap_phase[i] >= phases) return cl::fail(-1, "%s: tap_phase[%d]=%d out of range", kFn, i, tap_phase[i]);
    cl::ConvWgradPfDesc d{};
    d.grad = grad; d.g_plane_rows = g_plane_rows; d.act = act; d.x_plane_rows = x_plane_rows; d.Mp = Mp; d.Cout = Cout;
    d.Cin = Cin; d.phases = phases; d.num_taps = num_taps;
    for (int i = 0; i < num_taps; i++) { d.tap_shift[i] = tap_shift[i]; d.tap_phase[i] = tap_phase[i]; }
    d.nterms = nterms; d.out_scale = out_scale; d.scale_dev = scale_dev; d.oihw = oihw; d.dw = dw;
    return finish(kFn, cl::conv_wgrad_pf_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_gn_backward_fp4(int pass, int B, int H, int W, int C, int group_ch, const float* raw, const double* stats,
                                  const float* gamma, const float* beta, float eps, int relu_inner, int num_src,
                                  const float* const* src, const float* const* src_scale_a, const float* const* src_scale_b,
                                  const int32_t* src_stride, const int32_t* src_phased, const void* mask_out, float* g_out,
                                  double* ab, int ab_C, void* gmax_bits, void* d_raw, int64_t d_raw_lo_rows,
                                  float* scale_out, double* dbias, float* d_raw_f32, void* d_raw4, int64_t d_raw4_lo_rows,
                                  void* d_raw_sf, void* cuda_stream)
{
    static const char* kFn = "cl_gn_backward";
    NEED_DEV(raw); NEED_DEV(ab); NEED_DEV(gmax_bits);
    if (group_ch) { NEED_DEV(stats); NEED_DEV(gamma); NEED_DEV(beta); }
    if (pass != 0 && pass != 1) return cl::fail(-1, "%s: pass must be 0 (reduce) or 1 (apply)", kFn);
    if (num_src < 1 || num_src > 3 || !src || !src_scale_a || !src_scale_b || !src_stride || !src_phased)
        return cl::fail(-1, "%s: 1..3 gradient sources with their tables", kFn);
    cl::GnBwdDesc d{};
    d.B = B; d.H = H; d.W = W; d.C = C; d.group_ch = group_ch; d.raw = raw; d.stats = stats; d.gamma = gamma; d.beta = beta;
    d.eps = eps; d.relu_inner = relu_inner; d.num_src = num_src;
    for (int i = 0; i < num_src; i++) {
        d.src[i].g = src[i]; d.src[i].scale_a = src_scale_a[i]; d.src[i].scale_b = src_scale_b[i];
        d.src[i].stride = src_stride[i]; d.src[i].phased = src_phased[i];
        if (int rc = need_device(kFn, "src[i]", src[i])) return rc;
    }
    d.mask_out = static_cast<const __half*>(mask_out); d.g_out = g_out; d.ab = ab; d.ab_C = ab_C;
    if (ab_C && ab_C < C) return cl::fail(-1, "%s: ab_C=%d smaller than C=%d", kFn, ab_C, C);
    d.gmax_bits = static_cast<unsigned*>(gmax_bits); d.d_raw = static_cast<__half*>(d_raw); d.d_raw_lo_rows = d_raw_lo_rows;
    d.scale_out = scale_out; d.dbias = dbias; d.d_raw_f32 = d_raw_f32;
    if (d_raw_f32) NEED_DEV(d_raw_f32);
    if (d_raw4) { NEED_DEV(d_raw4); NEED_DEV(d_raw_sf); }
    d.d_raw4 = static_cast<uint8_t*>(d_raw4); d.d_raw4_lo_rows = d_raw4_lo_rows; d.d_raw_sf = static_cast<uint32_t*>(d_raw_sf);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (pass == 0) return finish(kFn, cl::gn_bwd_reduce_launch(d, s));
    NEED_DEV(d_raw); NEED_DEV(scale_out);
    return finish(kFn, cl::gn_bwd_apply_launch(d, s));
}

extern "C" int cl_gn_backward(int pass, int B, int H, int W, int C, int group_ch, const float* raw, const double* stats,
                              const float* gamma, const float* beta, float eps, int relu_inner, int num_src,
                              const float* const* src, const float* const* src_scale_a, const float* const* src_scale_b,
                              const int32_t* src_stride, const int32_t* src_phased, const void* mask_out, float* g_out,
                              double* ab, int ab_C, void* gmax_bits, void* d_raw, int64_t d_raw_lo_rows,
                              float* scale_out, double* dbias, float* d_raw_f32, void* cuda_stream)
{
    return cl_gn_backward_fp4(pass, B, H, W, C, group_ch, raw, stats, gamma, beta, eps, relu_inner, num_src, src, src_scale_a,
                              src_scale_b, src_stride, src_phased, mask_out, g_out, ab, ab_C, gmax_bits, d_raw, d_raw_lo_rows,
                              scale_out, dbias, d_raw_f32, nullptr, 0, nullptr, cuda_stream);
}

extern "C" int cl_frames_to_nchw(const uint8_t* frames, int B, int H, int W, int C, const float* mean, const float* stdv,
                                 float* out, void* cuda_stream)
{
    static const char* kFn = "cl_frames_to_nchw";
    NEED_DEV(frames); NEED_DEV(out);
    if (mean) NEED_DEV(mean);
    if (stdv) NEED_DEV(stdv);
    return finish(kFn, cl::frames_to_nchw_launch(frames, B, H, W, C, mean, stdv, out, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_pf_groupnorm(const void* in, int64_t in_lo_rows, int B, int H, int W, int C, int group_ch,
                               const float* gamma, const float* beta, float eps, double* stats, void* out, int out_terms,
                               void* out8, void* cuda_stream)
{
    static const char* kFn = "cl_pf_groupnorm";
    NEED_DEV(in); NEED_DEV(gamma); NEED_DEV(beta); NEED_DEV(stats); NEED_DEV(out);
    if (out8) NEED_DEV(out8);
    cl::PfGroupNormDesc d{};
    d.in = static_cast<const __half*>(in); d.in_lo_rows = in_lo_rows; d.B = B; d.H = H; d.W = W; d.C = C;
    d.group_ch = group_ch; d.gamma = gamma; d.beta = beta; d.eps = eps; d.stats = stats;
    d.out = static_cast<__half*>(out); d.out_terms = out_terms; d.out8 = static_cast<uint8_t*>(out8);
    return finish(kFn, cl::pf_groupnorm_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

extern "C" int cl_head_backward(const void* act, int64_t act_lo_rows, int B, int H, int W, int C, int Co,
                                const float* weight, const float* g_sc, float* g_x, float* g_w, void* cuda_stream)
{
    static const char* kFn = "cl_head_backward";
    NEED_DEV(act); NEED_DEV(weight); NEED_DEV(g_sc); NEED_DEV(g_x); NEED_DEV(g_w);
    cl::HeadBwdDesc d{};
    d.act = static_cast<const __half*>(act); d.act_lo_rows = act_lo_rows; d.B = B; d.H = H; d.W = W; d.C = C; d.Co = Co;
    d.weight = weight; d.g_sc = g_sc; d.g_x = g_x; d.g_w = g_w;
    return finish(kFn, cl::head_bwd_launch(d, static_cast<cudaStream_t>(cuda_stream)));
}

#ifdef CL_DEBUG_TRAP
namespace cl { void conv_debug_counters(unsigned long long* out16, bool reset); }
// debug builds only (-DCL_DEBUG_TRAP): wait-cycle counters of the fp4 convolution kernel
extern "C" int cl_debug_counters(unsigned long long* out16, int reset)
{
    cl::conv_debug_counters(out16, reset != 0);
    return 0;
}
#endif
