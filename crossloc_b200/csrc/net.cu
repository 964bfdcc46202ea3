// Network runtime: the whole scene-coordinate CNN behind one C-ABI call.
//
// Replaces `network(image)` of the reference's evaluation loop (/root/reference/test_single_task.py:347,
// utils/evaluation.py:106-116: TransPoseNet / Network built from stock nn.Conv2d + nn.GroupNorm -> cuDNN / ATen) for
// a caller bound where the reference binds its own extension (dsacstar.cpp:887-892): cl_net_create() takes the
// layer table of the module (state-dict tensors as plain pointers), cl_net_forward() runs it.
//
// What lives here (and used to be Python, crossloc_b200/cnn.py):
//   * filter packing: OIHW fp32 -> fp16 [term][tap][Cout][Cin] (+ e4m3 planes), power-of-two pre-scale;
//   * the layer plan of one (batch, height, width): padded-flat activation / raw buffers, the tap tables, which
//     operand planes every GroupNorm stage has to write for its consumers, one prepared ConvIgemmPlan (tile
//     schedule + five TMA tensor maps) per convolution;
//   * execution: the plan is captured once into a CUDA graph (one cudaGraphLaunch per forward instead of ~60 kernel
//     launches with a dozen driver calls each); CROSSLOC_B200_NET_GRAPH=0 replays the op list directly.
// Frames enter through a plan-owned staging buffer (host or device source, fp32 NCHW or uint8 HWC) and the result
// leaves through one, so the captured graph never depends on caller pointers.
#include "../../include/crossloc_b200.h"

#include <cuda_fp8.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "cabi_common.h"
#include "conv.h"

namespace cl {

namespace {

// ------------------------------------------------------------------------------------------- small device helpers
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    // keeps the allocation (and its address: prepared tensor maps point at it) when it is already large enough
    cudaError_t ensure(size_t n, bool zero)
    {
        if (p && bytes >= n) return zero ? cudaMemset(p, 0, bytes) : cudaSuccess;
        return alloc(n, zero);
    }
    cudaError_t alloc(size_t n, bool zero)
    {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n ? n : 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { p = nullptr; return e; }
        return zero ? cudaMemset(p, 0, bytes) : cudaSuccess;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// OIHW fp32 filter -> tensor-core planes.  out16: [planes16][tap][Cout][Cin] fp16 (hi, then lo when planes16 == 2);
// out8 (nullable): [2][tap][Cout][Cin] e4m3 = fp8(w_hi), fp8(w_lo * 2^12).  `scale` is the power-of-two pre-scale.
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float scale,
                                                       __half* __restrict__ out16, int planes16, uint8_t* __restrict__ out8)
{
    const size_t per = (size_t)taps * Cout * Cin;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout);
        const int tap = (int)(i / ((size_t)Cin * Cout));
        const float v = w[((size_t)co * Cin + ci) * taps + tap] * scale;
        const __half hi = __float2half_rn(v);
        const float lo = v - __half2float(hi);
        out16[i] = hi;
        if (planes16 == 2) out16[per + i] = __float2half_rn(lo);
        if (out8) {
            out8[i] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(hi), __NV_SATFINITE, __NV_E4M3);
            out8[per + i] = (uint8_t)__nv_cvt_float_to_fp8(lo * kW8LoScale, __NV_SATFINITE, __NV_E4M3);
        }
    }
}

// GroupNorm sums of a raw fp32 padded-flat tensor for group sizes the convolution epilogue does not cover (the DUC
// head of the full-size variant has 2 * Co channels per group: 6, 12, 14 ...).  One block per (image, group slice).
__global__ void __launch_bounds__(256) raw_stats_kernel(const float* __restrict__ raw, int H, int W, int C, int group_ch,
                                                       double* __restrict__ stats)
{
    const int groups = C / group_ch;
    const int b = blockIdx.y, g = blockIdx.x;
    const int Wp = W + 2;
    const size_t plane = (size_t)(H + 2) * Wp;
    double s = 0, ss = 0;
    const int n = H * W * group_ch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = i % group_ch, pix = i / group_ch;
        const int y = pix / W, x = pix - y * W;
        const float v = raw[((size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1)) * C + g * group_ch + j];
        s += v;
        ss += (double)v * v;
    }
    __shared__ double red[2][256];
    red[0][threadIdx.x] = s;
    red[1][threadIdx.x] = ss;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            red[0][threadIdx.x] += red[0][threadIdx.x + off];
            red[1][threadIdx.x] += red[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[((size_t)b * groups + g) * 2] = red[0][0];
        stats[((size_t)b * groups + g) * 2 + 1] = red[1][0];
    }
}

bool igemm_group_ok(int group_ch) { return group_ch == 0 || group_ch == 2 || group_ch == 4 || group_ch == 8 || group_ch == 16; }

// ------------------------------------------------------------------------------------------- parameters
struct Layer {
    int cin = 0, cout = 0, ksize = 1, stride = 1, taps = 1;
    int gn_groups = 0;
    float gn_eps = 1e-5f;
    // caller-owned sources (re-read by cl_net_update)
    const float *src_w = nullptr, *src_b = nullptr, *src_g = nullptr, *src_be = nullptr;
    // engine-owned device copies
    DevBuf w32, bias, gamma, beta, w16, w8, w4, wsf;
    int nterms = 3;
    float out_scale = 1.f;
    bool tensor_core = false;   // runs through conv_igemm (everything except the stem and the small heads)
    int group_ch() const { return gn_groups ? cout / gn_groups : 0; }
};

struct Block {
    int kind = 0;   // CL_BLOCK_*
    std::vector<int> convs;
    int skip = -1;
};

// ------------------------------------------------------------------------------------------- plan
struct Geometry {
    int B = 0, H = 0, W = 0, Hp = 0, Wp = 0, plane = 0, Mp = 0;
    void set(int b, int h, int w) { B = b; H = h; W = w; Hp = h + 2; Wp = w + 2; plane = Hp * Wp; Mp = b * plane; }
};

struct PF {   // one padded-flat activation: fp16 hi / lo planes and (on demand) the e4m3 planes
    const Geometry* geo = nullptr;
    int channels = 0, phases = 1, terms = 2;
    DevBuf h16, f8, f4, sf;   // f4 / sf: block-scaled e2m1 planes and their scale words (fp16 + fp4 consumers)
    int64_t rows16() const { return (int64_t)terms * phases * geo->Mp; }
    int64_t rows8() const { return (int64_t)2 * phases * geo->Mp; }
};

struct Op {
    enum Kind { kMemset, kStem, kConv, kApply, kHead, kDucHead, kRawStats, kFrames, kFork, kConvFused } kind = kMemset;
    cudaEvent_t event = nullptr;  // kFork
    int label[4] = {0, 0, 0, 0};  // profiling: conv (cin, cout, ksize, stride), apply (channels, out phases, add kind, 0)
    double flops = 0;             // algorithmic FLOPs of the launch (convolutions and the stem)
    // kMemset
    void* ptr = nullptr;
    size_t bytes = 0;
    // kStem
    StemDesc stem{};
    bool stem_stats = false;
    // kConv
    std::unique_ptr<ConvIgemmPlan> conv;
    // kApply
    GnApplyDesc apply{};
    // kHead / kDucHead
    HeadDesc head{};
    DucHeadDesc duc{};
    // kRawStats
    const float* rs_raw = nullptr;
    int rs_H = 0, rs_W = 0, rs_C = 0, rs_group_ch = 0, rs_B = 0;
    double* rs_stats = nullptr;
    // kFrames
    const uint8_t* fr_src = nullptr;
    int fr_B = 0, fr_H = 0, fr_W = 0, fr_C = 0;
    const float *fr_mean = nullptr, *fr_std = nullptr;
    float* fr_out = nullptr;
};

struct Plan {
    int B = 0, H = 0, W = 0, Cin = 0;
    Geometry geo[4];
    std::map<std::string, std::unique_ptr<PF>> acts;
    std::map<std::string, std::unique_ptr<DevBuf>> raws;
    DevBuf stats, in_f32, in_u8, out, norm_mean, norm_std;
    int out_C = 0, out_H = 0, out_W = 0;
    std::vector<Op> ops;          // the forward pass proper (graph-captured)
    Op frames_op;                 // uint8 HWC staging -> in_f32 (captured in its own graph variant)
    cudaGraphExec_t graph = nullptr, graph_frames = nullptr;
    bool graph_failed = false;
    bool warm = false, warm_frames = false;   // the op list ran eagerly once (function attributes set, modules loaded)
    int launches = 0;             // kernel launches of one forward
    uint64_t last_use = 0;
    std::vector<std::vector<cudaEvent_t>> profile;   // profiling mode: ops + 1 events per recorded forward
    void drop_profile()
    {
        for (auto& f : profile)
            for (cudaEvent_t e : f) cudaEventDestroy(e);
        profile.clear();
    }
    ~Plan()
    {
        drop_profile();
        if (graph) cudaGraphExecDestroy(graph);
        if (graph_frames) cudaGraphExecDestroy(graph_frames);
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------- engine
struct Net {
    int device = 0;
    int sms = 148;
    int precision = 2;         // 1 fp16x1 | 2 fp16 + fp8 | 3 fp16x3 | 4 fp16 + fp4 (fp16 + fp8 where a layer cannot)
    int terms = 2;             // fp16 planes per activation
    bool relu_after_add = true;
    bool fp8_1x1 = true;
    bool use_graph = true;
    bool profiling = false;    // eager launches with a CUDA event between ops (cl_net_profile)
    bool fuse_gn = false;      // GroupNorm / ReLU / residual merge in the convolution epilogue where the plan allows it
    bool dynamic_tiles = true; // tile indices from a global counter instead of a fixed stride (CTA-pair kernel)
    std::vector<std::unique_ptr<Layer>> layers;
    std::vector<Block> blocks;
    int stem[4] = {-1, -1, -1, -1};
    int head_layer = -1, duc_layer = -1, duc_rate = 0, num_task = 0;
    float clamp_lo = 0.f, clamp_hi = 0.f;
    DevBuf head_mean, head_w, head_b;
    const float* src_mean = nullptr;
    std::map<std::string, std::unique_ptr<Plan>> plans;
    uint64_t tick = 0;
    cudaStream_t capture_stream = nullptr;
    cudaEvent_t fork_event = nullptr;   // recorded by every forward when it enters the residual blocks (cl_net_wait_fork)
    cudaEvent_t done_event = nullptr;   // end of the last forward of THIS handle: its plans share workspaces, so forwards of
    bool done_recorded = false;         // one handle issued on different streams are ordered by it
    std::mutex mu;
    std::string err;

    ~Net()
    {
        plans.clear();
        if (capture_stream) cudaStreamDestroy(capture_stream);
        if (fork_event) cudaEventDestroy(fork_event);
        if (done_event) cudaEventDestroy(done_event);
    }
};

namespace {

int nterms_for(const Net& n, const Layer& L)
{
    if (n.precision == 1) return 1;
    if ((n.precision == 2 || n.precision == 4) && L.stride == 1 && L.cin % 128 == 0 && L.cin >= 256 &&
        (L.ksize == 3 || (n.fp8_1x1 && L.cin >= 512)))
        return n.precision == 4 && L.cin % 256 == 0 && L.cout % 256 == 0 ? 4 : 2;
    return 3;
}

const char* copy_param(DevBuf& dst, const float* src, size_t count)
{
    if (!src || !count) return nullptr;
    if (!dst.p || dst.bytes < count * sizeof(float))
        if (cudaError_t e = dst.alloc(count * sizeof(float), false)) return cudaGetErrorString(e);
    cudaError_t e = cudaMemcpy(dst.p, src, count * sizeof(float), cudaMemcpyDefault);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// (re)reads the caller's parameter tensors and repacks the filters
const char* load_params(Net& n)
{
    std::vector<float> host;
    for (size_t li = 0; li < n.layers.size(); li++) {
        Layer& L = *n.layers[li];
        const size_t wcount = (size_t)L.cout * L.cin * L.taps;
        if (const char* e = copy_param(L.w32, L.src_w, wcount)) return e;
        if (L.src_b) {
            if (const char* e = copy_param(L.bias, L.src_b, L.cout)) return e;
        } else {
            if (cudaError_t e = L.bias.ensure((size_t)L.cout * sizeof(float), true)) return cudaGetErrorString(e);
        }
        if (L.gn_groups) {
            if (const char* e = copy_param(L.gamma, L.src_g, L.cout)) return e;
            if (const char* e = copy_param(L.beta, L.src_be, L.cout)) return e;
        }
        if (!L.tensor_core) continue;
        // power-of-two pre-scale: max |w| lands in [64, 128) so that the low-order term stays out of the fp16 / e4m3
        // subnormals; undone by out_scale in the epilogue
        host.resize(wcount);
        if (cudaError_t e = cudaMemcpy(host.data(), L.w32.p, wcount * sizeof(float), cudaMemcpyDeviceToHost)) return cudaGetErrorString(e);
        float amax = 0.f;
        for (float v : host) amax = fmaxf(amax, fabsf(v));
        int ex = 0;
        if (amax > 0.f && std::isfinite(amax)) ex = (int)floorf(log2f(128.0f / amax));
        ex = ex < -24 ? -24 : (ex > 24 ? 24 : ex);
        L.out_scale = ldexpf(1.0f, -ex);
        L.nterms = nterms_for(n, L);
        const int planes16 = L.nterms == 3 ? 2 : 1;
        if (cudaError_t e = L.w16.ensure((size_t)planes16 * wcount * sizeof(__half), false)) return cudaGetErrorString(e);
        if (L.nterms == 2)
            if (cudaError_t e = L.w8.ensure((size_t)2 * wcount, false)) return cudaGetErrorString(e);
        if (L.nterms == 4) {
            if (cudaError_t e = L.w4.ensure(wcount, false)) return cudaGetErrorString(e);
            if (cudaError_t e = L.wsf.ensure((size_t)L.taps * (L.cin / 256) * L.cout * sizeof(uint32_t), false)) return cudaGetErrorString(e);
            if (const char* e = pack_conv_fp4_launch(L.w32.as<float>(), L.cout, L.cin, L.taps, ldexpf(1.0f, ex), L.w4.as<uint8_t>(),
                                                     L.wsf.as<uint32_t>(), nullptr))
                return e;
        }
        size_t blocks = (wcount + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        pack_conv_kernel<<<(unsigned)blocks, 256>>>(L.w32.as<float>(), L.cout, L.cin, L.taps, ldexpf(1.0f, ex), L.w16.as<__half>(),
                                                   planes16, L.nterms == 2 ? L.w8.as<uint8_t>() : nullptr);
        if (cudaError_t e = cudaGetLastError()) return cudaGetErrorString(e);
    }
    if (n.head_layer >= 0) {
        Layer& H = *n.layers[n.head_layer];
        if (const char* e = copy_param(n.head_w, H.src_w, (size_t)H.cout * H.cin)) return e;
        if (H.src_b) {
            if (const char* e = copy_param(n.head_b, H.src_b, H.cout)) return e;
        } else if (cudaError_t e = n.head_b.ensure((size_t)H.cout * sizeof(float), true)) return cudaGetErrorString(e);
    }
    if (n.num_task > 0)
        if (const char* e = copy_param(n.head_mean, n.src_mean, n.num_task)) return e;
    if (cudaError_t e = cudaDeviceSynchronize()) return cudaGetErrorString(e);
    return nullptr;
}

// ------------------------------------------------------------------------------------------- plan builder
struct Builder {
    Net& n;
    Plan& P;
    std::string err;
    int stat_i = 0;
    int max_groups = 32;
    int counter_i = 0;
    size_t stats_bytes = 0;

    // scheduler / publication counters live behind the statistics in one allocation: a single memset per forward
    int* next_counters(int count)
    {
        int* base = reinterpret_cast<int*>(static_cast<char*>(P.stats.p) + stats_bytes) + counter_i;
        counter_i += count;
        return base;
    }

    Builder(Net& net, Plan& plan) : n(net), P(plan) {}

    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }

    PF* act(const std::string& tag, int level, int channels, int phases)
    {
        const std::string key = tag + ":" + std::to_string(level) + ":" + std::to_string(channels) + ":" + std::to_string(phases);
        auto it = P.acts.find(key);
        if (it != P.acts.end()) return it->second.get();
        std::unique_ptr<PF> pf(new PF);
        pf->geo = &P.geo[level];
        pf->channels = channels;
        pf->phases = phases;
        pf->terms = n.terms;
        // zero-initialised once: kernels only ever write interior pixels, so the borders stay zero
        if (cudaError_t e = pf->h16.alloc((size_t)pf->rows16() * channels * sizeof(__half), true)) { fail(cudaGetErrorString(e)); return nullptr; }
        PF* out = pf.get();
        P.acts[key] = std::move(pf);
        return out;
    }
    uint8_t* f8(PF* a)
    {
        if (!a->f8.p)
            if (cudaError_t e = a->f8.alloc((size_t)a->rows8() * a->channels, true)) { fail(cudaGetErrorString(e)); return nullptr; }
        return a->f8.as<uint8_t>();
    }
    float* raw(const std::string& tag, int level, int channels)
    {
        const std::string key = tag + ":" + std::to_string(level) + ":" + std::to_string(channels);
        auto it = P.raws.find(key);
        if (it != P.raws.end()) return it->second->as<float>();
        std::unique_ptr<DevBuf> b(new DevBuf);
        if (cudaError_t e = b->alloc((size_t)P.geo[level].Mp * channels * sizeof(float), false)) { fail(cudaGetErrorString(e)); return nullptr; }
        float* out = b->as<float>();
        P.raws[key] = std::move(b);
        return out;
    }
    bool f4(PF* a)
    {
        if (a->f4.p) return true;
        if (a->phases != 1 || a->channels % 256 != 0) return fail("e2m1 planes need a same-resolution activation with C % 256 == 0");
        if (cudaError_t e = a->f4.alloc((size_t)a->geo->Mp * a->channels, true)) return fail(cudaGetErrorString(e));
        if (cudaError_t e = a->sf.alloc((size_t)(a->channels / 256) * a->geo->Mp * sizeof(uint32_t), true)) return fail(cudaGetErrorString(e));
        return true;
    }
    double* next_stats() { return P.stats.as<double>() + (size_t)(stat_i++) * P.B * max_groups * 2; }

    // activation row shift of every filter tap in the padded-flat layout of the OUTPUT resolution
    static int taps_of(const Layer& L, const Geometry& g, int* out)
    {
        if (L.ksize == 1) { out[0] = 0; return 1; }
        int k = 0;
        for (int kh = 0; kh < 3; kh++)
            for (int kw = 0; kw < 3; kw++) {
                if (L.stride == 1) {
                    out[k++] = (kh - 1) * g.Wp + (kw - 1);
                } else {
                    const int a = kh == 1 ? 0 : 1, dy = kh == 0 ? -1 : 0;
                    const int b = kw == 1 ? 0 : 1, dx = kw == 0 ? -1 : 0;
                    out[k++] = (a * 2 + b) * g.Mp + dy * g.Wp + dx;
                }
            }
        return 9;
    }

    struct Fuse {             // GroupNorm + ReLU (+ residual + ReLU) in the epilogue: the consumer's operand planes
        PF* out = nullptr;
        PF* res = nullptr;
        bool relu_inner = true, relu_outer = false, want_lo = true, want8 = false;
    };

    bool conv(int li, PF* a, const Geometry& g, float* rawbuf, double* st, const Fuse* fuse = nullptr)
    {
        Layer& L = *n.layers[li];
        if (!a || (!rawbuf && !fuse)) return false;
        ConvIgemmDesc d{};
        d.act = a->h16.p;
        d.a_total_rows = a->rows16();
        d.a_lo_rows = (int64_t)a->phases * g.Mp;
        d.Cin = L.cin;
        d.weights = L.w16.p;
        d.Cout = L.cout;
        d.num_taps = taps_of(L, g, d.tap_a_row);
        d.nterms = L.nterms;
        if (L.nterms == 2) {
            d.act8 = f8(a);
            if (!d.act8) return false;
            d.a8_total_rows = a->rows8();
            d.a8_lo_rows = (int64_t)a->phases * g.Mp;
            d.weights8 = L.w8.p;
        }
        d.corr_scale = kCorrScale;
        d.cluster = 0;
        if (L.nterms == 4) {
            if (!f4(a)) return false;
            d.act4 = a->f4.p;
            d.a4_total_rows = (int64_t)2 * g.Mp;
            d.a4_lo_rows = g.Mp;
            d.act_sf = a->sf.as<uint32_t>();
            d.weights4 = L.w4.p;
            d.w_sf = L.wsf.as<uint32_t>();
            d.cluster = 2;
        }
        d.Mp = g.Mp; d.Hp = g.Hp; d.Wp = g.Wp;
        const int gc = L.group_ch();
        const bool fused_stats = igemm_group_ok(gc);
        d.group_ch = fused_stats ? gc : 0;
        d.out_scale = L.out_scale;
        d.raw = rawbuf;
        d.bias = L.bias.as<float>();
        d.stats = (gc && fused_stats) ? st : nullptr;
        if (n.dynamic_tiles) d.tile_counter = next_counters(1);
        if (fuse) {
            d.fuse = 1;
            d.gamma = gc ? L.gamma.as<float>() : nullptr;
            d.beta = gc ? L.beta.as<float>() : nullptr;
            d.eps = L.gn_eps;
            d.H = g.H; d.W = g.W;
            d.relu_inner = fuse->relu_inner; d.relu_outer = fuse->relu_outer;
            if (fuse->res) {
                d.res = fuse->res->h16.as<__half>();
                d.res_lo_rows = n.terms == 2 ? g.Mp : 0;
            }
            d.out16 = fuse->out->h16.as<__half>();
            d.out_terms = (fuse->want_lo && n.terms == 2) ? 2 : 1;
            d.out8 = fuse->want8 ? f8(fuse->out) : nullptr;
            if (fuse->want8 && !d.out8) return false;
            d.unit_done = next_counters(g.B * 8);
        }
        Op op;
        op.kind = fuse ? Op::kConvFused : Op::kConv;
        op.conv.reset(new ConvIgemmPlan);
        if (const char* e = conv_igemm_prepare(d, op.conv.get())) return fail(std::string(e));
        op.label[0] = L.cin; op.label[1] = L.cout; op.label[2] = L.ksize; op.label[3] = L.stride;
        op.flops = 2.0 * g.B * g.H * g.W * (double)L.cout * L.cin * d.num_taps;   // borders and split terms excluded
        P.ops.push_back(std::move(op));
        P.launches++;
        if (gc && !fused_stats) {
            Op rs;
            rs.kind = Op::kRawStats;
            rs.rs_raw = rawbuf; rs.rs_H = g.H; rs.rs_W = g.W; rs.rs_C = L.cout; rs.rs_group_ch = gc; rs.rs_B = g.B; rs.rs_stats = st;
            P.ops.push_back(std::move(rs));
            P.launches++;
        }
        return true;
    }

    // Can layer li (a convolution at the output resolution followed by GroupNorm / ReLU / an optional fp16 residual merge)
    // run with the fused epilogue?  Mirrors the checks of conv_igemm_prepare so that the plan never has to back out.
    bool can_fuse(int li, const Geometry& g) const
    {
        if (!n.fuse_gn || !n.dynamic_tiles || n.precision == 4) return false;
        const Layer& L = *n.layers[li];
        const int gc = L.group_ch();
        if (!igemm_group_ok(gc)) return false;
        const int BN = L.cout % 256 == 0 ? 256 : (L.cout % 128 == 0 ? 128 : 64);
        if (gc && BN / gc > 32) return false;
        if (2 * BN > 512) return false;
        if (L.nterms == 2 && L.cin % 128 != 0) return false;
        const int tiles_m = (g.Mp + 127) / 128;
        if (tiles_m < 16 || n.sms % 2 != 0) return false;                      // CTA-pair kernel only
        if (g.plane < 32) return false;
        const int tiles_per_image = ((g.plane + 255) / 256 + 1) * (L.cout / BN);
        return tiles_per_image <= n.sms / 2;
    }

    struct Merge { PF* res = nullptr; float* raw2 = nullptr; int layer2 = -1; double* stats2 = nullptr; };

    bool apply(float* rawbuf, const Geometry& g, int li, double* st, PF* out, bool relu_inner, const Merge& m, bool relu_outer,
               bool want_lo, int want8)   // want8: bit 0 = e4m3 planes, bit 1 = block-scaled e2m1 planes
    {
        Layer& L = *n.layers[li];
        if (!out || !rawbuf) return false;
        GnApplyDesc d{};
        d.raw = rawbuf;
        d.B = g.B; d.H = g.H; d.W = g.W; d.C = L.cout;
        d.group_ch = L.group_ch();
        d.stats = d.group_ch ? st : nullptr;
        d.gamma = d.group_ch ? L.gamma.as<float>() : nullptr;
        d.beta = d.group_ch ? L.beta.as<float>() : nullptr;
        d.eps = L.gn_eps;
        d.relu_inner = relu_inner ? 1 : 0;
        d.add_kind = m.res ? 1 : (m.raw2 ? 2 : 0);
        if (m.res) {
            d.res = m.res->h16.as<__half>();
            d.res_lo_rows = n.terms == 2 ? g.Mp : 0;
        }
        if (m.raw2) {
            Layer& S = *n.layers[m.layer2];
            d.raw2 = m.raw2;
            d.stats2 = m.stats2;
            d.gamma2 = S.gamma.as<float>();
            d.beta2 = S.beta.as<float>();
        }
        d.relu_outer = relu_outer ? 1 : 0;
        d.out = out->h16.as<__half>();
        d.out_phases = out->phases;
        d.out_terms = (want_lo && n.terms == 2) ? 2 : 1;
        d.out8 = (want8 & 1) ? f8(out) : nullptr;
        if ((want8 & 1) && !d.out8) return false;
        if (want8 & 2) {
            if (!f4(out)) return false;
            d.out4 = out->f4.as<uint8_t>();
            d.out_sf = out->sf.as<uint32_t>();
        }
        d.out_C = 0; d.out_c0 = 0;
        Op op;
        op.kind = Op::kApply;
        op.apply = d;
        op.label[0] = L.cout; op.label[1] = out->phases; op.label[2] = d.add_kind;
        P.ops.push_back(std::move(op));
        P.launches++;
        return true;
    }

    void planes_for(const std::vector<int>& consumers, bool also_lo, bool& want_lo, int& want8) const
    {
        want8 = 0;
        want_lo = also_lo;
        for (int c : consumers) {
            if (c < 0) continue;
            const Layer& L = *n.layers[c];
            if (L.nterms == 2) want8 |= 1;
            if (L.nterms == 4) want8 |= 2;
            if (L.nterms == 3) want_lo = true;
        }
    }

    std::vector<int> first_conv_of(size_t bi) const
    {
        std::vector<int> out;
        if (bi >= n.blocks.size()) {
            if (n.duc_layer >= 0) out.push_back(n.duc_layer);
            return out;
        }
        const Block& b = n.blocks[bi];
        out.push_back(b.convs[0]);
        if (b.kind == CL_BLOCK_RESIDUAL_SKIP) out.push_back(b.skip);
        return out;
    }

    PF* scratch(int channels, PF* a, PF* b)
    {
        for (int i = 0; i < 4; i++) {
            PF* cand = act("pool" + std::to_string(i), 3, channels, 1);
            if (!cand) return nullptr;
            if (cand != a && cand != b) return cand;
        }
        fail("no free scratch activation");
        return nullptr;
    }

    // conv -> [GN] -> relu per layer; the last layer merges the residual stream
    PF* chain(const std::vector<int>& convs, PF* x, PF* res_in, bool outer_relu, const std::vector<int>& next_readers, const Merge* skip)
    {
        const Geometry& g3 = P.geo[3];
        for (size_t i = 0; i < convs.size(); i++) {
            const int li = convs[i];
            Layer& L = *n.layers[li];
            double* st = L.gn_groups ? next_stats() : nullptr;
            PF* out = scratch(L.cout, x, res_in);
            if (!out) return nullptr;
            const bool last = i + 1 == convs.size();
            bool want_lo;
            int want8;
            if (!last) planes_for({convs[i + 1]}, false, want_lo, want8);
            else planes_for(next_readers, true, want_lo, want8);
            const bool merge_raw2 = last && skip && skip->raw2;
            if (!merge_raw2 && can_fuse(li, g3)) {
                Fuse f;
                f.out = out;
                f.res = last ? (skip ? skip->res : res_in) : nullptr;
                f.relu_inner = true;
                f.relu_outer = last && outer_relu;
                f.want_lo = want_lo;
                f.want8 = want8 != 0;
                if (!conv(li, x, g3, nullptr, st, &f)) return nullptr;
            } else {
                float* r = raw(i % 2 ? "r1" : "r0", 3, L.cout);
                if (!conv(li, x, g3, r, st)) return nullptr;
                if (!last) {
                    if (!apply(r, g3, li, st, out, true, Merge{}, false, want_lo, want8)) return nullptr;
                } else {
                    Merge m;
                    if (skip) m = *skip;
                    else m.res = res_in;
                    if (!apply(r, g3, li, st, out, true, m, outer_relu, want_lo, want8)) return nullptr;
                }
            }
            x = out;
        }
        return x;
    }

    bool build(int B, int Cin, int H, int W)
    {
        P.B = B; P.H = H; P.W = W; P.Cin = Cin;
        int hh = H, ww = W;
        for (int level = 0; level < 4; level++) {
            P.geo[level].set(B, hh, ww);
            if (level < 3) { hh = (hh + 1) / 2; ww = (ww + 1) / 2; }
        }
        for (auto& L : n.layers)
            if (L->gn_groups > max_groups) max_groups = L->gn_groups;
        const size_t n_stat = n.layers.size() + 2;
        stats_bytes = n_stat * B * max_groups * 2 * sizeof(double);
        const size_t counter_bytes = n_stat * ((size_t)B * 8 + 1) * sizeof(int);
        if (cudaError_t e = P.stats.alloc(stats_bytes + counter_bytes, true)) return fail(cudaGetErrorString(e));
        if (cudaError_t e = P.in_f32.alloc((size_t)B * Cin * H * W * sizeof(float), false)) return fail(cudaGetErrorString(e));
        {
            Op op;
            op.kind = Op::kMemset;
            op.ptr = P.stats.p;
            op.bytes = P.stats.bytes;
            P.ops.push_back(std::move(op));
        }
        // ---- stem: conv1 (+ norm1) + relu, written as the 4-phase input of conv2
        Layer& C1 = *n.layers[n.stem[0]];
        if (C1.cout != 32 || (C1.gn_groups && C1.gn_groups != 32)) return fail("the stem kernel is built for 32 channels / 32 groups");
        if (C1.cin != Cin) return fail("image channels do not match conv1");
        PF* a = act("stem", 1, 32, 4);
        if (!a) return false;
        double* st = C1.gn_groups ? next_stats() : nullptr;
        StemDesc sd{};
        sd.image = P.in_f32.as<float>();
        sd.B = B; sd.Cin = Cin; sd.H = H; sd.W = W;
        sd.weight = C1.w32.as<float>();
        sd.bias = C1.bias.as<float>();
        sd.has_gn = C1.gn_groups ? 1 : 0;
        sd.stats = st;
        sd.gamma = C1.gamma.as<float>();
        sd.beta = C1.beta.as<float>();
        sd.eps = C1.gn_eps;
        sd.out = a->h16.as<__half>();
        sd.out_terms = n.terms;
        sd.raw_out = nullptr;
        for (int pass = C1.gn_groups ? 0 : 1; pass < 2; pass++) {
            Op op;
            op.kind = Op::kStem;
            op.stem = sd;
            op.stem_stats = pass == 0;
            op.label[0] = pass;
            op.flops = 2.0 * B * H * W * 32.0 * Cin * 9;
            P.ops.push_back(std::move(op));
            P.launches++;
        }
        // ---- strided ladder conv2..conv4
        for (int level = 1; level <= 3; level++) {
            const int li = n.stem[level];
            Layer& L = *n.layers[li];
            float* r = raw("ladder", level, L.cout);
            double* s2 = L.gn_groups ? next_stats() : nullptr;
            if (!conv(li, a, P.geo[level], r, s2)) return false;
            PF* out;
            bool want_lo = true;
            int want8 = 0;
            if (level < 3) {
                out = act("ladder", level + 1, L.cout, 4);
            } else {
                out = act("res", 3, L.cout, 1);
                planes_for(first_conv_of(0), true, want_lo, want8);
            }
            if (!apply(r, P.geo[level], li, s2, out, true, Merge{}, false, want_lo, want8)) return false;
            a = out;
        }
        PF* res = a;
        {
            // From here on the plan is tensor-core bound (3x3 / 1x1 convolutions at the output resolution): the point where
            // a caller's side-stream work (the previous batch's pose solve) best starts to share the SMs.
            Op op;
            op.kind = Op::kFork;
            op.event = n.fork_event;
            P.ops.push_back(std::move(op));
        }
        const bool outer = n.relu_after_add;
        const Geometry& g3 = P.geo[3];
        for (size_t bi = 0; bi < n.blocks.size(); bi++) {
            const Block& blk = n.blocks[bi];
            const std::vector<int> readers = first_conv_of(bi + 1);
            if (blk.kind == CL_BLOCK_RESIDUAL) {
                res = chain(blk.convs, res, res, outer, readers, nullptr);
            } else if (blk.kind == CL_BLOCK_RESIDUAL_SKIP) {
                Layer& S = *n.layers[blk.skip];
                float* rs = raw("rs", 3, S.cout);
                double* st_s = S.gn_groups ? next_stats() : nullptr;
                if (!conv(blk.skip, res, g3, rs, st_s)) return false;
                if (S.gn_groups) {
                    Merge m;
                    m.raw2 = rs; m.layer2 = blk.skip; m.stats2 = st_s;
                    res = chain(blk.convs, res, res, outer, readers, &m);
                } else {
                    // vanilla Network: res = skip(res) + relu(conv(x)), no normalisation anywhere
                    PF* skip_pf = scratch(S.cout, res, nullptr);
                    if (!skip_pf) return false;
                    if (!apply(rs, g3, blk.skip, nullptr, skip_pf, false, Merge{}, false, true, false)) return false;
                    res = chain(blk.convs, res, skip_pf, outer, readers, nullptr);
                }
            } else if (blk.kind == CL_BLOCK_PLAIN) {
                for (size_t i = 0; i < blk.convs.size(); i++) {
                    const int li = blk.convs[i];
                    Layer& L = *n.layers[li];
                    double* s2 = L.gn_groups ? next_stats() : nullptr;
                    PF* out = scratch(L.cout, res, nullptr);
                    if (!out) return false;
                    bool want_lo;
            int want8;
                    planes_for(i + 1 < blk.convs.size() ? std::vector<int>{blk.convs[i + 1]} : readers, true, want_lo, want8);
                    if (can_fuse(li, g3)) {
                        Fuse f;
                        f.out = out;
                        f.want_lo = want_lo;
                        f.want8 = want8 != 0;
                        if (!conv(li, res, g3, nullptr, s2, &f)) return false;
                    } else {
                        float* r = raw("r0", 3, L.cout);
                        if (!conv(li, res, g3, r, s2)) return false;
                        if (!apply(r, g3, li, s2, out, true, Merge{}, false, want_lo, want8)) return false;
                    }
                    res = out;
                }
            } else {
                return fail("unknown block kind");
            }
            if (!res) return false;
        }
        // ---- head
        Layer& HL = *n.layers[n.head_layer];
        const int co = HL.cout;
        if (n.duc_layer >= 0) {
            Layer& D = *n.layers[n.duc_layer];
            float* r = raw("duc", 3, D.cout);
            double* s2 = next_stats();
            if (!D.gn_groups) return fail("the DUC convolution is followed by a GroupNorm");
            if (!conv(n.duc_layer, res, g3, r, s2)) return false;
            P.out_C = co; P.out_H = H; P.out_W = W;
            if (cudaError_t e = P.out.alloc((size_t)B * co * H * W * sizeof(float), false)) return fail(cudaGetErrorString(e));
            DucHeadDesc d{};
            d.raw = r; d.B = B; d.Hc = g3.H; d.Wc = g3.W; d.C = D.cout; d.Co = co; d.rate = n.duc_rate;
            d.group_ch = D.group_ch();
            d.stats = s2; d.gamma = D.gamma.as<float>(); d.beta = D.beta.as<float>(); d.eps = D.gn_eps;
            d.weight = n.head_w.as<float>(); d.bias = n.head_b.as<float>(); d.mean = n.head_mean.as<float>();
            d.num_task = n.num_task; d.clamp_lo = n.clamp_lo; d.clamp_hi = n.clamp_hi;
            d.out = P.out.as<float>(); d.Ho = H; d.Wo = W;
            Op op;
            op.kind = Op::kDucHead;
            op.duc = d;
            P.ops.push_back(std::move(op));
        } else {
            P.out_C = co; P.out_H = g3.H; P.out_W = g3.W;
            if (cudaError_t e = P.out.alloc((size_t)B * co * g3.H * g3.W * sizeof(float), false)) return fail(cudaGetErrorString(e));
            HeadDesc d{};
            d.act = res->h16.as<__half>();
            d.act_lo_rows = g3.Mp;
            d.in_terms = n.terms;
            d.B = B; d.H = g3.H; d.W = g3.W; d.C = HL.cin; d.Co = co;
            d.weight = n.head_w.as<float>(); d.bias = n.head_b.as<float>(); d.mean = n.head_mean.as<float>();
            d.num_task = n.num_task; d.clamp_lo = n.clamp_lo; d.clamp_hi = n.clamp_hi;
            d.out = P.out.as<float>();
            Op op;
            op.kind = Op::kHead;
            op.head = d;
            P.ops.push_back(std::move(op));
        }
        P.launches++;
        return err.empty();
    }
};

const char* run_op(const Op& op, cudaStream_t s)
{
    switch (op.kind) {
        case Op::kMemset: {
            cudaError_t e = cudaMemsetAsync(op.ptr, 0, op.bytes, s);
            return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
        }
        case Op::kStem: return stem_tc_launch(op.stem, op.stem_stats, s);
        case Op::kConv:
        case Op::kConvFused: return conv_igemm_run(*op.conv, s);
        case Op::kApply: return gn_apply_launch(op.apply, s);
        case Op::kHead: return head_launch(op.head, s);
        case Op::kDucHead: return duc_head_launch(op.duc, s);
        case Op::kRawStats: {
            raw_stats_kernel<<<dim3(op.rs_C / op.rs_group_ch, op.rs_B), 256, 0, s>>>(op.rs_raw, op.rs_H, op.rs_W, op.rs_C, op.rs_group_ch,
                                                                                    op.rs_stats);
            cudaError_t e = cudaGetLastError();
            return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
        }
        case Op::kFork: {
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(s, &st);
            cudaError_t e = cudaEventRecordWithFlags(op.event, s, st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
            return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
        }
        case Op::kFrames:
            return frames_to_nchw_launch(op.fr_src, op.fr_B, op.fr_H, op.fr_W, op.fr_C, op.fr_mean, op.fr_std, op.fr_out, s);
    }
    return "unknown op";
}

const char* run_ops(const Plan& P, bool frames, cudaStream_t s)
{
    if (frames)
        if (const char* e = run_op(P.frames_op, s)) return e;
    for (const Op& op : P.ops)
        if (const char* e = run_op(op, s)) return e;
    return nullptr;
}

// Captures the op list into an executable graph (on the engine's own capture stream, so that the caller's stream is never
// put into capture mode).  Every kernel has been launched eagerly once before: function attributes are set and lazily
// loaded modules are resident, which capture would otherwise trip over.
const char* capture(Net& n, Plan& P, bool frames, cudaGraphExec_t* out)
{
    if (!n.capture_stream)
        if (cudaError_t e = cudaStreamCreateWithFlags(&n.capture_stream, cudaStreamNonBlocking)) return cudaGetErrorString(e);
    if (cudaError_t e = cudaStreamBeginCapture(n.capture_stream, cudaStreamCaptureModeThreadLocal)) return cudaGetErrorString(e);
    const char* err = run_ops(P, frames, n.capture_stream);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(n.capture_stream, &g);
    if (err) { if (g) cudaGraphDestroy(g); return err; }
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

Plan* plan_for(Net& n, int B, int Cin, int H, int W, std::string& err)
{
    const std::string key = std::to_string(B) + "x" + std::to_string(Cin) + "x" + std::to_string(H) + "x" + std::to_string(W);
    auto it = n.plans.find(key);
    if (it != n.plans.end()) { it->second->last_use = ++n.tick; return it->second.get(); }
    while (n.plans.size() >= 3) {   // evaluation frames come in a few sizes at most; training crops go through the Python plan
        auto victim = n.plans.begin();
        for (auto i = n.plans.begin(); i != n.plans.end(); ++i)
            if (i->second->last_use < victim->second->last_use) victim = i;
        cudaDeviceSynchronize();    // the victim's buffers may still be in use by queued work
        n.plans.erase(victim);
    }
    std::unique_ptr<Plan> P(new Plan);
    Builder b(n, *P);
    if (!b.build(B, Cin, H, W)) { err = b.err.empty() ? "plan construction failed" : b.err; return nullptr; }
    if (cudaError_t e = cudaDeviceSynchronize()) { err = cudaGetErrorString(e); return nullptr; }   // buffer memsets done
    P->last_use = ++n.tick;
    Plan* out = P.get();
    n.plans[key] = std::move(P);
    return out;
}

// With the (opt-in) fused GroupNorm epilogue, forwards of ALL handles on one device are chained by an event: a
// fused-epilogue convolution spins on publications of its own grid, so two such launches from different streams, each
// holding part of the SMs, could wait for each other forever.  Without it forwards on different streams may overlap:
// the memory-bound passes of one (GroupNorm apply, stem, head) then share the SMs with the tensor-bound convolutions
// of the other (two-lane mode of crossloc_b200.pipeline).
std::mutex g_chain_mutex;
cudaEvent_t g_chain_event[64] = {};

bool g_chain_enabled = false;   // set once a handle uses the fused GroupNorm epilogue (cl_net_create); never cleared

const char* chain_begin(cudaStream_t stream, int dev)
{
    if (!g_chain_enabled) return nullptr;
    std::lock_guard<std::mutex> lock(g_chain_mutex);
    cudaEvent_t& ev = g_chain_event[dev & 63];
    if (!ev) {
        if (cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) return cudaGetErrorString(e);
        return nullptr;
    }
    cudaError_t e = cudaStreamWaitEvent(stream, ev, 0);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* chain_end(cudaStream_t stream, int dev)
{
    if (!g_chain_enabled) return nullptr;
    std::lock_guard<std::mutex> lock(g_chain_mutex);
    cudaError_t e = cudaEventRecord(g_chain_event[dev & 63], stream);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

int forward_impl(Net& n, const void* image, bool frames_u8, const float* mean, const float* stdv, int B, int H, int W, float* out,
                 cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(n.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != n.device) return fail(-1, "cl_net_forward: the network lives on device %d, the current device is %d", n.device, dev);
    if (B <= 0 || H < 8 || W < 8) return fail(-1, "cl_net_forward: invalid sizes B=%d H=%d W=%d", B, H, W);
    const int Cin = n.layers[n.stem[0]]->cin;
    std::string err;
    Plan* P = plan_for(n, B, Cin, H, W, err);
    if (!P) return fail(-2, "cl_net_forward: %s", err.c_str());
    // the plans of a handle share input / activation / output buffers: order this forward behind the previous one of the
    // handle, whatever stream that was issued on (forwards of different handles may overlap)
    if (n.done_recorded) CL_CUDA(cudaStreamWaitEvent(stream, n.done_event, 0));
    // ---- stage the frames
    if (frames_u8) {
        const size_t bytes = (size_t)B * H * W * Cin;
        if (P->frames_op.kind != Op::kFrames) {
            if (!P->in_u8.p) CL_CUDA(P->in_u8.alloc(bytes, false));
            if (mean) {
                CL_CUDA(P->norm_mean.alloc(Cin * sizeof(float), false));
                CL_CUDA(P->norm_std.alloc(Cin * sizeof(float), false));
            }
            Op& f = P->frames_op;
            f.kind = Op::kFrames;
            f.fr_src = P->in_u8.as<uint8_t>();
            f.fr_B = B; f.fr_H = H; f.fr_W = W; f.fr_C = Cin;
            f.fr_mean = mean ? P->norm_mean.as<float>() : nullptr;
            f.fr_std = mean ? P->norm_std.as<float>() : nullptr;
            f.fr_out = P->in_f32.as<float>();
        }
        if ((mean != nullptr) != (P->frames_op.fr_mean != nullptr))
            return fail(-1, "cl_net_forward_frames: mean / std must be given on every call of a plan or on none");
        if (mean) {
            CL_CUDA(cudaMemcpyAsync(P->norm_mean.p, mean, Cin * sizeof(float), cudaMemcpyDefault, stream));
            CL_CUDA(cudaMemcpyAsync(P->norm_std.p, stdv, Cin * sizeof(float), cudaMemcpyDefault, stream));
        }
        if (image != P->in_u8.p) CL_CUDA(cudaMemcpyAsync(P->in_u8.p, image, bytes, cudaMemcpyDefault, stream));
    } else if (image != P->in_f32.p) {
        CL_CUDA(cudaMemcpyAsync(P->in_f32.p, image, (size_t)B * Cin * H * W * sizeof(float), cudaMemcpyDefault, stream));
    }
    // ---- run
    if (const char* e = chain_begin(stream, dev)) return fail(-2, "cl_net_forward: %s", e);
    cudaGraphExec_t& g = frames_u8 ? P->graph_frames : P->graph;
    bool& warm = frames_u8 ? P->warm_frames : P->warm;
    if (n.use_graph && !n.profiling && !P->graph_failed && warm && !g) {
        if (const char* e = capture(n, *P, frames_u8, &g)) {
            P->graph_failed = true;   // this plan keeps launching directly
            g = nullptr;
            n.err = std::string("CUDA graph capture failed: ") + e;
            cudaGetLastError();
        }
    }
    if (n.profiling) {
        std::vector<cudaEvent_t> ev(P->ops.size() + 1);
        for (auto& e : ev) CL_CUDA(cudaEventCreate(&e));
        if (frames_u8)
            if (const char* e = run_op(P->frames_op, stream)) return fail(-2, "cl_net_forward: %s", e);
        CL_CUDA(cudaEventRecord(ev[0], stream));
        for (size_t i = 0; i < P->ops.size(); i++) {
            if (const char* e = run_op(P->ops[i], stream)) return fail(-2, "cl_net_forward: %s", e);
            CL_CUDA(cudaEventRecord(ev[i + 1], stream));
        }
        P->profile.push_back(std::move(ev));
        warm = true;
    } else if (g) {
        CL_CUDA(cudaGraphLaunch(g, stream));
    } else {
        if (const char* e = run_ops(*P, frames_u8, stream)) return fail(-2, "cl_net_forward: %s", e);
        warm = true;
    }
    if (const char* e = chain_end(stream, dev)) return fail(-2, "cl_net_forward: %s", e);
    // ---- hand the result over
    const size_t out_bytes = (size_t)B * P->out_C * P->out_H * P->out_W * sizeof(float);
    if (out && out != P->out.p) {
        CL_CUDA(cudaMemcpyAsync(out, P->out.p, out_bytes, cudaMemcpyDefault, stream));
        if (!is_device_ptr(out)) CL_CUDA(cudaStreamSynchronize(stream));
    }
    CL_CUDA(cudaEventRecord(n.done_event, stream));
    n.done_recorded = true;
    return 0;
}

}  // namespace
}  // namespace cl

// ------------------------------------------------------------------------------------------- C ABI
extern "C" int cl_net_create(const cl_net_desc* desc, cl_net** out)
{
    using namespace cl;
    if (!desc || !out) return fail(-1, "cl_net_create: desc and out must not be NULL");
    *out = nullptr;
    if (desc->abi_version != CL_NET_ABI_VERSION) return fail(-1, "cl_net_create: abi_version %d, library has %d", desc->abi_version, CL_NET_ABI_VERSION);
    if (desc->n_layers <= 0 || !desc->layers) return fail(-1, "cl_net_create: empty layer table");
    if (desc->precision < 1 || desc->precision > 4) return fail(-1, "cl_net_create: precision must be 1 (fp16x1), 2 (fp16+fp8), 3 (fp16x3) or 4 (fp16+fp4)");
    std::unique_ptr<Net> n(new Net);
    cudaGetDevice(&n->device);
    if (cudaDeviceGetAttribute(&n->sms, cudaDevAttrMultiProcessorCount, n->device) != cudaSuccess) n->sms = 148;
    n->precision = desc->precision;
    n->terms = desc->precision == 1 ? 1 : 2;
    n->relu_after_add = desc->relu_after_add != 0;
    const char* env = getenv("CROSSLOC_B200_FP8_1X1");
    n->fp8_1x1 = !(env && env[0] == '0');
    env = getenv("CROSSLOC_B200_NET_GRAPH");
    n->use_graph = !(env && env[0] == '0');
    env = getenv("CROSSLOC_B200_FUSE_GN");
    n->fuse_gn = env && env[0] == '1';   // opt-in: measured slower than conv + gn_apply (26.5 vs 23.2 ms per 32 frames)
    if (n->fuse_gn) g_chain_enabled = true;
    env = getenv("CROSSLOC_B200_DYNAMIC_TILES");
    n->dynamic_tiles = !(env && env[0] == '0');
    auto idx_ok = [&](int i) { return i >= 0 && i < desc->n_layers; };
    for (int i = 0; i < desc->n_layers; i++) {
        const cl_net_layer& s = desc->layers[i];
        std::unique_ptr<Layer> L(new Layer);
        if ((s.ksize != 1 && s.ksize != 3) || (s.stride != 1 && s.stride != 2) || s.cin <= 0 || s.cout <= 0 || !s.weight)
            return fail(-1, "cl_net_create: layer %d: unsupported convolution (ksize %d stride %d cin %d cout %d)", i, s.ksize, s.stride, s.cin, s.cout);
        if (s.gn_groups < 0 || (s.gn_groups && (s.cout % s.gn_groups != 0 || !s.gn_weight || !s.gn_bias)))
            return fail(-1, "cl_net_create: layer %d: invalid GroupNorm (%d groups over %d channels)", i, s.gn_groups, s.cout);
        L->cin = s.cin; L->cout = s.cout; L->ksize = s.ksize; L->stride = s.stride; L->taps = s.ksize * s.ksize;
        L->gn_groups = s.gn_groups; L->gn_eps = s.gn_eps;
        L->src_w = s.weight; L->src_b = s.bias; L->src_g = s.gn_weight; L->src_be = s.gn_bias;
        n->layers.push_back(std::move(L));
    }
    for (int k = 0; k < 4; k++) {
        if (!idx_ok(desc->stem[k])) return fail(-1, "cl_net_create: stem[%d] = %d is not a layer", k, desc->stem[k]);
        n->stem[k] = desc->stem[k];
        if (k > 0) n->layers[desc->stem[k]]->tensor_core = true;
    }
    for (int b = 0; b < desc->n_blocks; b++) {
        const cl_net_block& s = desc->blocks[b];
        Block blk;
        blk.kind = s.kind;
        if (s.n_convs < 1 || s.n_convs > CL_NET_MAX_BLOCK_CONVS) return fail(-1, "cl_net_create: block %d has %d convolutions", b, s.n_convs);
        for (int k = 0; k < s.n_convs; k++) {
            if (!idx_ok(s.convs[k])) return fail(-1, "cl_net_create: block %d: convolution index %d out of range", b, s.convs[k]);
            blk.convs.push_back(s.convs[k]);
            n->layers[s.convs[k]]->tensor_core = true;
        }
        if (s.kind == CL_BLOCK_RESIDUAL_SKIP) {
            if (!idx_ok(s.skip)) return fail(-1, "cl_net_create: block %d: skip index %d out of range", b, s.skip);
            blk.skip = s.skip;
            n->layers[s.skip]->tensor_core = true;
        } else if (s.kind != CL_BLOCK_RESIDUAL && s.kind != CL_BLOCK_PLAIN) {
            return fail(-1, "cl_net_create: block %d: unknown kind %d", b, s.kind);
        }
        n->blocks.push_back(blk);
    }
    if (!idx_ok(desc->head_layer)) return fail(-1, "cl_net_create: head_layer %d is not a layer", desc->head_layer);
    n->head_layer = desc->head_layer;
    if (n->layers[n->head_layer]->ksize != 1 || n->layers[n->head_layer]->cout > 8)
        return fail(-1, "cl_net_create: the head is a 1x1 convolution with at most 8 output channels");
    n->num_task = desc->num_task;
    if (n->num_task < 0 || n->num_task > n->layers[n->head_layer]->cout || (n->num_task > 0 && !desc->head_mean))
        return fail(-1, "cl_net_create: num_task = %d needs 0 <= num_task <= Co and the mean vector", desc->num_task);
    n->src_mean = desc->head_mean;
    n->clamp_lo = desc->clamp_lo; n->clamp_hi = desc->clamp_hi;
    n->duc_layer = -1;
    if (desc->duc_layer >= 0) {
        if (!idx_ok(desc->duc_layer) || desc->duc_rate < 1) return fail(-1, "cl_net_create: invalid DUC layer / rate");
        n->duc_layer = desc->duc_layer;
        n->duc_rate = desc->duc_rate;
        n->layers[n->duc_layer]->tensor_core = true;
    }
    for (auto& L : n->layers) {
        if (!L->tensor_core) continue;
        if (L->cin % 32 != 0 || L->cout % 64 != 0)
            return fail(-1, "cl_net_create: tensor-core layers need Cin %% 32 == 0 and Cout %% 64 == 0 (got %d -> %d)", L->cin, L->cout);
    }
    if (const char* e = load_params(*n)) return fail(-2, "cl_net_create: %s", e);
    CL_CUDA(cudaEventCreateWithFlags(&n->fork_event, cudaEventDisableTiming));
    CL_CUDA(cudaEventCreateWithFlags(&n->done_event, cudaEventDisableTiming));
    *out = reinterpret_cast<cl_net*>(n.release());
    return 0;
}

extern "C" int cl_net_update(cl_net* net)
{
    using namespace cl;
    if (!net) return fail(-1, "cl_net_update: NULL handle");
    Net& n = *reinterpret_cast<Net*>(net);
    std::lock_guard<std::mutex> lock(n.mu);
    if (cudaError_t e = cudaDeviceSynchronize()) return fail(-2, "cl_net_update: %s", cudaGetErrorString(e));
    // filter planes keep their addresses and sizes (the packing scheme of a layer only depends on its shape), so prepared
    // plans and captured graphs stay valid
    if (const char* e = load_params(n)) return fail(-2, "cl_net_update: %s", e);
    return 0;
}

extern "C" int cl_net_forward(cl_net* net, const float* image, int B, int H, int W, float* out, void* cuda_stream)
{
    using namespace cl;
    if (!net || !image) return fail(-1, "cl_net_forward: NULL handle or image");
    return forward_impl(*reinterpret_cast<Net*>(net), image, false, nullptr, nullptr, B, H, W, out, static_cast<cudaStream_t>(cuda_stream));
}

extern "C" int cl_net_forward_frames(cl_net* net, const uint8_t* frames, int B, int H, int W, const float* mean, const float* stdv,
                                     float* out, void* cuda_stream)
{
    using namespace cl;
    if (!net || !frames) return fail(-1, "cl_net_forward_frames: NULL handle or frames");
    if ((mean == nullptr) != (stdv == nullptr)) return fail(-1, "cl_net_forward_frames: mean and std come together");
    return forward_impl(*reinterpret_cast<Net*>(net), frames, true, mean, stdv, B, H, W, out, static_cast<cudaStream_t>(cuda_stream));
}

extern "C" int cl_net_wait_fork(cl_net* net, void* cuda_stream)
{
    using namespace cl;
    if (!net) return fail(-1, "cl_net_wait_fork: NULL handle");
    Net& n = *reinterpret_cast<Net*>(net);
    CL_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(cuda_stream), n.fork_event, 0));
    return 0;
}

extern "C" int cl_net_output_shape(cl_net* net, int B, int H, int W, int* out_c, int* out_h, int* out_w)
{
    using namespace cl;
    if (!net || !out_c || !out_h || !out_w) return fail(-1, "cl_net_output_shape: NULL argument");
    Net& n = *reinterpret_cast<Net*>(net);
    (void)B;
    *out_c = n.layers[n.head_layer]->cout;
    if (n.duc_layer >= 0) { *out_h = H; *out_w = W; return 0; }
    for (int i = 0; i < 3; i++) { H = (H + 1) / 2; W = (W + 1) / 2; }
    *out_h = H; *out_w = W;
    return 0;
}

extern "C" int cl_net_buffers(cl_net* net, int B, int H, int W, float** in_f32, uint8_t** in_u8, float** out_buf, int* launches)
{
    using namespace cl;
    if (!net) return fail(-1, "cl_net_buffers: NULL handle");
    Net& n = *reinterpret_cast<Net*>(net);
    std::lock_guard<std::mutex> lock(n.mu);
    std::string err;
    Plan* P = plan_for(n, B, n.layers[n.stem[0]]->cin, H, W, err);
    if (!P) return fail(-2, "cl_net_buffers: %s", err.c_str());
    if (in_u8 && !P->in_u8.p) CL_CUDA(P->in_u8.alloc((size_t)B * H * W * P->Cin, false));
    if (in_f32) *in_f32 = P->in_f32.as<float>();
    if (in_u8) *in_u8 = P->in_u8.as<uint8_t>();
    if (out_buf) *out_buf = P->out.as<float>();
    if (launches) *launches = P->launches;
    return 0;
}

extern "C" int cl_net_profile(cl_net* net, int enable, int B, int H, int W, int max_ops, int32_t* kinds, int32_t* labels,
                              double* flops, float* ms, int* forwards)
{
    using namespace cl;
    if (!net) return fail(-1, "cl_net_profile: NULL handle");
    Net& n = *reinterpret_cast<Net*>(net);
    std::lock_guard<std::mutex> lock(n.mu);
    int count = 0;
    if (ms && max_ops > 0) {
        std::string err;
        Plan* P = plan_for(n, B, n.layers[n.stem[0]]->cin, H, W, err);
        if (!P) return fail(-2, "cl_net_profile: %s", err.c_str());
        count = (int)P->ops.size();
        if (count > max_ops) return fail(-1, "cl_net_profile: the plan has %d ops, the caller's tables hold %d", count, max_ops);
        for (int i = 0; i < count; i++) {
            ms[i] = 0.f;
            if (kinds) kinds[i] = (int32_t)P->ops[i].kind;
            if (labels) for (int k = 0; k < 4; k++) labels[4 * i + k] = P->ops[i].label[k];
            if (flops) flops[i] = P->ops[i].flops;
        }
        for (auto& f : P->profile) {
            CL_CUDA(cudaEventSynchronize(f.back()));
            for (int i = 0; i < count; i++) {
                float t = 0.f;
                CL_CUDA(cudaEventElapsedTime(&t, f[i], f[i + 1]));
                ms[i] += t;
            }
        }
        if (forwards) *forwards = (int)P->profile.size();
        P->drop_profile();
    }
    n.profiling = enable != 0;
    return count;
}

extern "C" void cl_net_destroy(cl_net* net)
{
    if (!net) return;
    cudaDeviceSynchronize();
    delete reinterpret_cast<cl::Net*>(net);
}
