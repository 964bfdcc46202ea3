// Backward of the fused "GroupNorm -> ReLU -> (residual merge -> ReLU)" stage of the coordinate network on
// padded-flat (PF) tensors -- the autograd kernels behind nn.GroupNorm / F.relu / the residual adds of
// /root/reference/networks/networks.py:231-254, 332-343 in a train_single_task.py step (:298, loss.backward()).
//
// For one convolution layer with raw output r (fp32 PF), statistics (mean, rstd per image and group), affine
// (gamma, beta) and output gradient g (sum of up to three fp32 PF sources, optionally masked by the sign of the
// layer's merged output):
//     xhat = (r - mean) * rstd,  y = xhat * gamma + beta,  dy = g * [y > 0]
//     d_r  = rstd * (dy * gamma - mean_group(dy * gamma) - xhat * mean_group(dy * gamma * xhat))
//     d_gamma = sum dy * xhat,  d_beta = sum dy,  d_bias(conv) = sum d_r
// Pass 1 (gn_bwd_reduce) accumulates A1 = sum dy and A2 = sum dy * xhat per (image, channel) in fp64, records
// max |dy * gamma| * rstd and, on request, stores the summed / masked gradient (the part that continues down the
// residual stream).  Pass 2 (gn_bwd_apply) forms the group means from A1 / A2, writes d_r as fp16 hi / lo PF
// operand planes for the data- and weight-gradient GEMMs, scaled by a power of two derived on the device from the
// recorded maximum (no host synchronisation), and accumulates the convolution's bias gradient.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "conv.h"
#include "fp4_planes.cuh"

namespace cl {

namespace {

constexpr int kThreads = 256;

// UNI: the 8 channels of a thread share one GroupNorm group (group size a multiple of 8, or no normalisation):
// mean / rstd are then scalars (entry 0) and the kernel needs ~14 registers less.
template <bool UNI>
struct Lane {
    int c;                      // first of the 8 channels of this thread
    float ga[8], be[8], mean[UNI ? 1 : 8], rstd[UNI ? 1 : 8];
};

template <bool UNI>
__device__ __forceinline__ void load_lane(const GnBwdDesc& d, int b, int c, Lane<UNI>& t)
{
    t.c = c;
    const int groups = d.group_ch ? d.C / d.group_ch : 0;
    const double count = (double)d.group_ch * d.H * d.W;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (d.group_ch) {
            t.ga[j] = d.gamma[c + j];
            t.be[j] = d.beta[c + j];
        } else {
            t.ga[j] = 1.f; t.be[j] = 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < (UNI ? 1 : 8); j++) {
        if (d.group_ch) {
            const int g = (c + j) / d.group_ch;
            const double s = d.stats[((size_t)b * groups + g) * 2], ss = d.stats[((size_t)b * groups + g) * 2 + 1];
            const double m = s / count;
            double var = ss / count - m * m;
            var = var > 0 ? var : 0;
            t.mean[j] = (float)m;
            t.rstd[j] = (float)(1.0 / sqrt(var + (double)d.eps));   // same expression as the forward (cnn_pointwise.cu)
        } else {
            t.mean[j] = 0.f; t.rstd[j] = 1.f;
        }
    }
}

__device__ __forceinline__ float src_scale(const GnBwdSrc& s)
{
    float v = 1.f;
    if (s.scale_a) v *= *s.scale_a;
    if (s.scale_b) v *= *s.scale_b;
    return v;
}

// row of pixel (b, y, x) in a source: same geometry, or the 4-phase form at half resolution
__device__ __forceinline__ size_t src_row(const GnBwdDesc& d, const GnBwdSrc& s, int b, int y, int x)
{
    if (!s.phased) return (size_t)b * (d.H + 2) * (d.W + 2) + (size_t)(y + 1) * (d.W + 2) + (x + 1);
    const int Hh = (d.H + 1) / 2, Wh = (d.W + 1) / 2, Wp = Wh + 2;
    const size_t plane = (size_t)(Hh + 2) * Wp;
    const int ph = (y & 1) * 2 + (x & 1);
    return ((size_t)ph * d.B + b) * plane + (size_t)((y >> 1) + 1) * Wp + ((x >> 1) + 1);
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// gradient of the stage output at (b, y, x), channels c .. c + 7: sum of the sources (pass 1) or source 0 (pass 2)
__device__ __forceinline__ void gather_grad(const GnBwdDesc& d, int nsrc, const float (&scale)[3], int b, int y, int x, int c,
                                            float (&g)[8])
{
#pragma unroll
    for (int j = 0; j < 8; j++) g[j] = 0.f;
    for (int s = 0; s < nsrc; s++) {
        float v[8];
        load8(d.src[s].g + src_row(d, d.src[s], b, y, x) * d.src[s].stride + c, v);
#pragma unroll
        for (int j = 0; j < 8; j++) g[j] = fmaf(v[j], scale[s], g[j]);
    }
}

template <bool APPLY, bool UNI>
__global__ void __launch_bounds__(kThreads, APPLY ? 4 : 3) gn_bwd_kernel(GnBwdDesc d)
{
    __shared__ float red[kThreads][17];
    __shared__ float2 group_means[256];     // APPLY: (mean_group(dy*gamma), mean_group(dy*gamma*xhat)) of this image
    __shared__ float scale_s;
    const int b = blockIdx.y;
    const int chunks = d.C / 8;
    const int chunk = threadIdx.x % chunks, pslot = threadIdx.x / chunks, pslots = kThreads / chunks;
    const int c = chunk * 8;
    const int Wp = d.W + 2;
    const size_t plane = (size_t)(d.H + 2) * Wp;
    const int ab_C = d.ab_C ? d.ab_C : d.C;   // channels per image in the ab buffer (it may hold several stages side by side)
    Lane<UNI> t;
    load_lane<UNI>(d, b, c, t);
    const int nsrc = APPLY ? 1 : d.num_src;
    float scale[3] = {1.f, 1.f, 1.f};
    for (int s = 0; s < nsrc; s++) scale[s] = src_scale(d.src[s]);

    float out_scale = 1.f;
    if (APPLY) {
        const int groups = d.group_ch ? d.C / d.group_ch : 0;
        const double inv_n = d.group_ch ? 1.0 / ((double)d.group_ch * d.H * d.W) : 0.0;
        for (int g = threadIdx.x; g < groups; g += kThreads) {
            double s1 = 0, s2 = 0;
            for (int j = 0; j < d.group_ch; j++) {
                const int ch = g * d.group_ch + j;
                const double ga = d.gamma[ch];
                s1 += ga * d.ab[((size_t)b * ab_C + ch) * 2];
                s2 += ga * d.ab[((size_t)b * ab_C + ch) * 2 + 1];
            }
            group_means[g] = make_float2((float)(s1 * inv_n), (float)(s2 * inv_n));
        }
        if (threadIdx.x == 0) {
            // power-of-two scale bringing max |dy * gamma| * rstd to [128, 256): |d_r| stays below ~12x that bound
            const float m = __uint_as_float(*d.gmax_bits);
            int k = 0;
            if (m > 0.f && isfinite(m)) {
                k = (int)floorf(log2f(256.f / m));
                k = k < -60 ? -60 : (k > 60 ? 60 : k);
            }
            scale_s = exp2f((float)k);
            if (blockIdx.x == 0 && b == 0) { d.scale_out[0] = exp2f((float)k); d.scale_out[1] = exp2f((float)-k); }
        }
        __syncthreads();
        out_scale = scale_s;
    }

    float a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a1[j] = 0.f; a2[j] = 0.f; }
    float gmax = 0.f;
    const int total = d.H * d.W;
    for (int pix = blockIdx.x * pslots + pslot; pix < total; pix += gridDim.x * pslots) {
        const int y = pix / d.W, x = pix - y * d.W;
        const size_t row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
        float g[8], r[8];
        gather_grad(d, nsrc, scale, b, y, x, c, g);
        load8(d.raw + row * d.C + c, r);
        if (!APPLY && d.mask_out) {
            const uint4 hq = __ldg(reinterpret_cast<const uint4*>(d.mask_out + row * d.C + c));
            const __half* hh = reinterpret_cast<const __half*>(&hq);
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = __half2float(hh[j]) > 0.f ? g[j] : 0.f;
        }
        if (!APPLY && d.g_out) {
            float4* o = reinterpret_cast<float4*>(d.g_out + row * d.C + c);
            o[0] = make_float4(g[0], g[1], g[2], g[3]);
            o[1] = make_float4(g[4], g[5], g[6], g[7]);
        }
        float dr[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float mean = t.mean[UNI ? 0 : j], rstd = t.rstd[UNI ? 0 : j];
            const float xhat = (r[j] - mean) * rstd;
            const float v = (r[j] - mean) * (rstd * t.ga[j]) + t.be[j];   // the forward's expression
            const float dy = (d.relu_inner && !(v > 0.f)) ? 0.f : g[j];
            if (!APPLY) {
                a1[j] += dy;
                a2[j] += dy * xhat;
                gmax = fmaxf(gmax, fabsf(dy * t.ga[j]) * rstd);
            } else {
                float val = dy;
                if (d.group_ch) {
                    const float2 gm = group_means[(c + (UNI ? 0 : j)) / d.group_ch];
                    val = rstd * (dy * t.ga[j] - gm.x - xhat * gm.y);
                }
                dr[j] = val;
                a1[j] += val;
            }
        }
        if (APPLY && d.d_raw_f32) {
            float4* o = reinterpret_cast<float4*>(d.d_raw_f32 + row * d.C + c);
            o[0] = make_float4(dr[0], dr[1], dr[2], dr[3]);
            o[1] = make_float4(dr[4], dr[5], dr[6], dr[7]);
        }
        if (APPLY) {
            __align__(16) __half h[8];
            __align__(16) __half l[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float sv = dr[j] * out_scale;
                h[j] = __float2half_rn(sv);
                l[j] = __float2half_rn(sv - __half2float(h[j]));
            }
            *reinterpret_cast<uint4*>(d.d_raw + row * d.C + c) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(d.d_raw + (row + (size_t)d.d_raw_lo_rows) * d.C + c) = *reinterpret_cast<const uint4*>(l);
            if (d.d_raw4) {
                // C % 256 == 0: the 32 lanes of a warp hold the 256 channels of one pixel and walk the pixels together
                float sv8[8];
#pragma unroll
                for (int j = 0; j < 8; j++) sv8[j] = dr[j] * out_scale;
                const size_t half_c = (size_t)d.C / 2;
                const uint32_t word = fp4_store8(d.d_raw4 + row * half_c + c / 2,
                                                 d.d_raw4 + (row + (size_t)d.d_raw4_lo_rows) * half_c + c / 2, sv8);
                if ((c & 255) == 0) d.d_raw_sf[(size_t)(c >> 8) * (size_t)d.d_raw4_lo_rows + row] = word;
            }
        }
    }

    // block reduction over the pixel slots of every channel chunk, then one fp64 atomic per (block, channel, moment)
#pragma unroll
    for (int j = 0; j < 8; j++) { red[threadIdx.x][j] = a1[j]; red[threadIdx.x][8 + j] = a2[j]; }
    red[threadIdx.x][16] = gmax;
    __syncthreads();
    if (pslot == 0) {
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { s1[j] = 0.f; s2[j] = 0.f; }
        for (int p = 0; p < pslots; p++) {
#pragma unroll
            for (int j = 0; j < 8; j++) { s1[j] += red[p * chunks + chunk][j]; s2[j] += red[p * chunks + chunk][8 + j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (!APPLY) {
                atomicAdd(d.ab + ((size_t)b * ab_C + c + j) * 2, (double)s1[j]);
                atomicAdd(d.ab + ((size_t)b * ab_C + c + j) * 2 + 1, (double)s2[j]);
            } else if (d.dbias) {
                atomicAdd(d.dbias + c + j, (double)s1[j]);
            }
        }
    }
    if (!APPLY) {
        float m = 0.f;
        if (threadIdx.x < 32) {
            for (int i = threadIdx.x; i < kThreads; i += 32) m = fmaxf(m, red[i][16]);
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
            if (threadIdx.x == 0 && m > 0.f) atomicMax(d.gmax_bits, __float_as_uint(m));
        }
    }
}

// Backward of the 1x1 output head (fc3: C -> Co <= 8; networks.py:349) on the padded-flat activation it read:
//   g_x[row][c] = sum_o g_sc[b][o][y][x] * W[o][c]          (fp32 PF source for the last stage's cl_gn_backward)
//   g_W[o][c]  += sum over pixels of g_sc[b][o][y][x] * (x_hi + x_lo)[row][c]
// One thread owns 8 channels (C / 8 divides the block); g_sc is the gradient of the pre-activation map, NCHW fp32.
constexpr int kHeadBwdMaxCo = 8;

template <int CO>
__global__ void __launch_bounds__(kThreads) head_bwd_kernel(HeadBwdDesc d)
{
    __shared__ float red[kThreads][8 * 4 + 1];   // one (o-slice of 4) x 8 channels per pass
    const int b = blockIdx.y;
    const int chunks = d.C / 8;
    const int chunk = threadIdx.x % chunks, pslot = threadIdx.x / chunks, pslots = kThreads / chunks;
    const int c = chunk * 8;
    const int Wp = d.W + 2;
    const size_t plane = (size_t)(d.H + 2) * Wp;
    const int hw = d.H * d.W;
    float w[CO][8];
    float acc[CO][8];
#pragma unroll
    for (int o = 0; o < CO; o++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            w[o][j] = o < d.Co ? d.weight[o * d.C + c + j] : 0.f;
            acc[o][j] = 0.f;
        }
    for (int pix = blockIdx.x * pslots + pslot; pix < hw; pix += gridDim.x * pslots) {
        const int y = pix / d.W, x = pix - y * d.W;
        const size_t row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
        const uint4 hq = __ldg(reinterpret_cast<const uint4*>(d.act + row * d.C + c));
        const uint4 lq = __ldg(reinterpret_cast<const uint4*>(d.act + (row + (size_t)d.act_lo_rows) * d.C + c));
        const __half* hh = reinterpret_cast<const __half*>(&hq);
        const __half* ll = reinterpret_cast<const __half*>(&lq);
        float xv[8], gx[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { xv[j] = __half2float(hh[j]) + __half2float(ll[j]); gx[j] = 0.f; }
#pragma unroll
        for (int o = 0; o < CO; o++) {
            if (o < d.Co) {
                const float g = __ldg(d.g_sc + ((size_t)b * d.Co + o) * hw + pix);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    gx[j] = fmaf(g, w[o][j], gx[j]);
                    acc[o][j] = fmaf(g, xv[j], acc[o][j]);
                }
            }
        }
        float4* out = reinterpret_cast<float4*>(d.g_x + row * d.C + c);
        out[0] = make_float4(gx[0], gx[1], gx[2], gx[3]);
        out[1] = make_float4(gx[4], gx[5], gx[6], gx[7]);
    }
    // block reduction of the weight-gradient partials, four output channels per pass
#pragma unroll
    for (int o0 = 0; o0 < CO; o0 += 4) {
        if (o0 >= d.Co) break;   // uniform across the block
        __syncthreads();
#pragma unroll
        for (int oo = 0; oo < 4; oo++)
#pragma unroll
            for (int j = 0; j < 8; j++) red[threadIdx.x][oo * 8 + j] = acc[o0 + oo][j];
        __syncthreads();
        if (pslot == 0) {
            for (int oo = 0; oo < 4 && o0 + oo < d.Co; oo++) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float t = 0.f;
                    for (int p = 0; p < pslots; p++) t += red[p * chunks + chunk][oo * 8 + j];
                    atomicAdd(d.g_w + (size_t)(o0 + oo) * d.C + c + j, t);
                }
            }
        }
    }
}

const char* check(const GnBwdDesc& d)
{
    if (d.C % 8 != 0 || d.C > 2048) return "gn_backward: C must be a multiple of 8 and at most 2048";
    if (kThreads % (d.C / 8) != 0) return "gn_backward: C / 8 must divide 256";
    if (d.group_ch && (d.C % d.group_ch != 0 || d.C / d.group_ch > 256)) return "gn_backward: at most 256 groups";
    if (d.B <= 0 || d.B > 65535 || d.H <= 0 || d.W <= 0) return "gn_backward: invalid sizes";
    if (d.num_src < 1 || d.num_src > 3) return "gn_backward: 1..3 gradient sources";
    for (int s = 0; s < d.num_src; s++)
        if (!d.src[s].g || d.src[s].stride < d.C || d.src[s].stride % 4 != 0) return "gn_backward: invalid gradient source";
    return nullptr;
}

// one resident wave of blocks (the launch bounds allow 3 / 4 blocks per SM): every extra block adds 16 fp64 atomics per
// channel chunk to the tail, measured 0.17 ms (one wave) vs 0.22 ms (four waves) for a 512-channel stage of 12 frames
int grid_x(const GnBwdDesc& d, int blocks_per_sm)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int pslots = kThreads / (d.C / 8);
    int bx = (d.H * d.W + pslots * 4 - 1) / (pslots * 4);     // >= 4 pixels per thread: the reduction tail stays small
    const char* env = getenv("CL_GNBWD_CAP");
    const int per_sm = env ? atoi(env) : blocks_per_sm;
    const int cap = (sms * per_sm) / d.B;
    if (bx > cap) bx = cap;
    return bx < 1 ? 1 : bx;
}

}  // namespace

const char* head_bwd_launch(const HeadBwdDesc& d, cudaStream_t stream)
{
    if (d.Co < 1 || d.Co > kHeadBwdMaxCo) return "head_backward: 1..8 output channels";
    if (d.C % 8 != 0 || kThreads % (d.C / 8) != 0) return "head_backward: C / 8 must divide 256";
    if (d.B <= 0 || d.B > 65535 || d.H <= 0 || d.W <= 0) return "head_backward: invalid sizes";
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int pslots = kThreads / (d.C / 8);
    int bx = (d.H * d.W + pslots * 4 - 1) / (pslots * 4);
    const int cap = (sms * 2) / d.B > 0 ? (sms * 2) / d.B : 1;     // one resident wave: every block ends with atomics
    if (bx > cap) bx = cap;
    if (d.Co <= 4) head_bwd_kernel<4><<<dim3(bx, d.B), kThreads, 0, stream>>>(d);
    else head_bwd_kernel<8><<<dim3(bx, d.B), kThreads, 0, stream>>>(d);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* gn_bwd_reduce_launch(const GnBwdDesc& d, cudaStream_t stream)
{
    if (const char* e = check(d)) return e;
    if (!d.ab || !d.gmax_bits) return "gn_backward: the reduce pass needs the ab and gmax buffers";
    if (d.group_ch % 8 == 0) gn_bwd_kernel<false, true><<<dim3(grid_x(d, 3), d.B), kThreads, 0, stream>>>(d);
    else gn_bwd_kernel<false, false><<<dim3(grid_x(d, 3), d.B), kThreads, 0, stream>>>(d);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* gn_bwd_apply_launch(const GnBwdDesc& d, cudaStream_t stream)
{
    if (const char* e = check(d)) return e;
    if (!d.ab || !d.gmax_bits || !d.d_raw || !d.scale_out) return "gn_backward: the apply pass needs ab, gmax, d_raw and scale_out";
    if (d.d_raw4 && (d.C % 256 != 0 || !d.d_raw_sf || d.d_raw4_lo_rows <= 0))
        return "gn_backward: e2m1 gradient planes need C % 256 == 0, the scale table and the plane distance";
    if (d.group_ch % 8 == 0) gn_bwd_kernel<true, true><<<dim3(grid_x(d, 4), d.B), kThreads, 0, stream>>>(d);
    else gn_bwd_kernel<true, false><<<dim3(grid_x(d, 4), d.B), kThreads, 0, stream>>>(d);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cl
