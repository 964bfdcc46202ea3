// HBM-bound kernels of the coordinate network: GroupNorm apply / residual merge, the 3-channel stem
// convolution (recomputed instead of stored) and the 1x1 output head.
//
// Reference ops replaced (all /root/reference/networks/networks.py):
//   F.relu(norm(conv(x)))                 :231-238, 242-244, 336-343     -> gn_apply_kernel (add_kind 0)
//   F.relu(res + x)                        :240, 254, 334, 340           -> gn_apply_kernel (add_kind 1)
//   res2_skip_norm(res2_skip(res)) + x     :246-249                      -> gn_apply_kernel (add_kind 2)
//   conv1 + norm1 + relu                   :189-190, 231                 -> stem_kernel
//   fc3, += mean, exp(hardtanh())          :349-358                      -> head_kernel
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv.h"

namespace cl {

namespace {

__device__ __forceinline__ void mean_rstd(const double* stats, int b, int groups, int g, double count, float eps,
                                          float& mean, float& rstd)
{
    const double s = stats[((size_t)b * groups + g) * 2], ss = stats[((size_t)b * groups + g) * 2 + 1];
    const double m = s / count;
    double var = ss / count - m * m;   // biased variance, as torch.nn.GroupNorm
    var = var > 0 ? var : 0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

__device__ __forceinline__ void split_store8(__half* hi_ptr, __half* lo_ptr, const float (&v)[8], bool write_lo)
{
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        h[j] = __float2half_rn(v[j]);
        l[j] = __float2half_rn(v[j] - __half2float(h[j]));
    }
    *reinterpret_cast<uint4*>(hi_ptr) = *reinterpret_cast<const uint4*>(h);
    if (write_lo) *reinterpret_cast<uint4*>(lo_ptr) = *reinterpret_cast<const uint4*>(l);
}

// One thread = 8 consecutive channels of one interior pixel.
__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyDesc d)
{
    const int c8n = d.C / 8;
    const long long total = (long long)d.B * d.H * d.W * c8n;
    const int Wp = d.W + 2, plane = (d.H + 2) * Wp;
    const int Ho = (d.H + 1) / 2, Wo = (d.W + 1) / 2, Wop = Wo + 2, oplane = (Ho + 2) * Wop;
    const int groups = d.group_ch ? d.C / d.group_ch : 1;
    const double count = (double)d.group_ch * d.H * d.W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % c8n) * 8;
        const long long pix = idx / c8n;
        const int x = (int)(pix % d.W);
        const int y = (int)((pix / d.W) % d.H);
        const int b = (int)(pix / ((long long)d.W * d.H));
        const size_t row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);

        float v[8];
        {
            const float4* r4 = reinterpret_cast<const float4*>(d.raw + row * d.C + c);
            const float4 a = __ldg(r4), bq = __ldg(r4 + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
        }
        if (d.group_ch) {
            int gprev = -1;
            float mean = 0.f, rstd = 1.f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int g = (c + j) / d.group_ch;
                if (g != gprev) { mean_rstd(d.stats, b, groups, g, count, d.eps, mean, rstd); gprev = g; }
                v[j] = (v[j] - mean) * rstd * __ldg(d.gamma + c + j) + __ldg(d.beta + c + j);
            }
        }
        if (d.relu_inner) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
        }
        if (d.add_kind == 1) {
            const uint4 hq = __ldg(reinterpret_cast<const uint4*>(d.res + row * d.C + c));
            const __half* hh = reinterpret_cast<const __half*>(&hq);
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] += __half2float(hh[j]);
            if (d.res_lo_rows > 0) {   // the residual carries a low-order plane
                const uint4 lq = __ldg(reinterpret_cast<const uint4*>(d.res + (row + (size_t)d.res_lo_rows) * d.C + c));
                const __half* ll = reinterpret_cast<const __half*>(&lq);
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] += __half2float(ll[j]);
            }
        } else if (d.add_kind == 2) {
            const float4* r4 = reinterpret_cast<const float4*>(d.raw2 + row * d.C + c);
            const float4 a = __ldg(r4), bq = __ldg(r4 + 1);
            float w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
            int gprev = -1;
            float mean = 0.f, rstd = 1.f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int g = (c + j) / d.group_ch;
                if (g != gprev) { mean_rstd(d.stats2, b, groups, g, count, d.eps, mean, rstd); gprev = g; }
                v[j] += (w[j] - mean) * rstd * __ldg(d.gamma2 + c + j) + __ldg(d.beta2 + c + j);
            }
        }
        if (d.relu_outer) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
        }

        size_t orow, olo;
        if (d.out_phases == 1) {
            orow = row;
            olo = (size_t)d.B * plane;
        } else {
            const int ph = (y & 1) * 2 + (x & 1);
            orow = ((size_t)ph * d.B + b) * oplane + (size_t)(y / 2 + 1) * Wop + (x / 2 + 1);
            olo = (size_t)4 * d.B * oplane;
        }
        split_store8(d.out + orow * d.C + c, d.out + (orow + olo) * d.C + c, v, d.out_terms == 2);
    }
}

// ---------------------------------------------------------------------------------------------
// Stem: direct 3x3 convolution, one output pixel (32 channels) per thread, weights broadcast from smem.
// STATS pass accumulates per-(image, channel) sums; APPLY pass recomputes the convolution, normalises,
// applies ReLU and writes the fp16 hi/lo planes straight into the four-phase input of the next layer.
constexpr int kStemThreads = 256;
constexpr int kStemCo = 32;

template <bool STATS>
__global__ void __launch_bounds__(kStemThreads) stem_kernel(StemDesc d)
{
    __shared__ float w_s[27 * kStemCo];   // [ci*9 + kh*3 + kw][co]
    __shared__ float b_s[kStemCo];
    __shared__ float red_s[kStemThreads / 32][2 * kStemCo];
    const int taps = d.Cin * 9;
    for (int i = threadIdx.x; i < taps * kStemCo; i += kStemThreads) {
        const int co = i % kStemCo, t = i / kStemCo;
        w_s[i] = d.weight[co * taps + t];
    }
    if (threadIdx.x < kStemCo) b_s[threadIdx.x] = d.bias[threadIdx.x];
    __syncthreads();

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int segs_x = (d.W + 31) / 32;
    const int segs = d.H * segs_x;
    const float* img = d.image + (size_t)b * d.Cin * d.H * d.W;
    const int Ho = (d.H + 1) / 2, Wo = (d.W + 1) / 2, Wop = Wo + 2, oplane = (Ho + 2) * Wop;

    float scale[kStemCo], shift[kStemCo];
    if (!STATS) {
        const double count = (double)d.H * d.W;
#pragma unroll
        for (int co = 0; co < kStemCo; co++) {
            if (d.has_gn) {
                float mean, rstd;
                mean_rstd(d.stats, b, kStemCo, co, count, d.eps, mean, rstd);
                scale[co] = rstd * d.gamma[co];
                shift[co] = d.beta[co] - mean * scale[co];
            } else {
                scale[co] = 1.f;
                shift[co] = 0.f;
            }
        }
    }
    float s1[kStemCo], s2[kStemCo];
    if (STATS) {
#pragma unroll
        for (int co = 0; co < kStemCo; co++) { s1[co] = 0.f; s2[co] = 0.f; }
    }

    for (int seg = blockIdx.x * (kStemThreads / 32) + warp; seg < segs; seg += gridDim.x * (kStemThreads / 32)) {
        const int y = seg / segs_x, x = (seg % segs_x) * 32 + lane;
        const bool inside = x < d.W;
        float acc[kStemCo];
#pragma unroll
        for (int co = 0; co < kStemCo; co++) acc[co] = b_s[co];
        for (int ci = 0; ci < d.Cin; ci++) {
#pragma unroll
            for (int kh = 0; kh < 3; kh++) {
                const int yy = y + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; kw++) {
                    const int xx = x + kw - 1;
                    float xv = 0.f;
                    if (inside && yy >= 0 && yy < d.H && xx >= 0 && xx < d.W) xv = __ldg(img + ((size_t)ci * d.H + yy) * d.W + xx);
                    const float4* w4 = reinterpret_cast<const float4*>(w_s + (ci * 9 + kh * 3 + kw) * kStemCo);
#pragma unroll
                    for (int j = 0; j < kStemCo / 4; j++) {
                        const float4 w = w4[j];
                        acc[4 * j + 0] = fmaf(xv, w.x, acc[4 * j + 0]);
                        acc[4 * j + 1] = fmaf(xv, w.y, acc[4 * j + 1]);
                        acc[4 * j + 2] = fmaf(xv, w.z, acc[4 * j + 2]);
                        acc[4 * j + 3] = fmaf(xv, w.w, acc[4 * j + 3]);
                    }
                }
            }
        }
        if (STATS) {
            if (inside) {
#pragma unroll
                for (int co = 0; co < kStemCo; co++) { s1[co] += acc[co]; s2[co] += acc[co] * acc[co]; }
            }
        } else if (inside) {
            const int ph = (y & 1) * 2 + (x & 1);
            const size_t orow = ((size_t)ph * d.B + b) * oplane + (size_t)(y / 2 + 1) * Wop + (x / 2 + 1);
            const size_t olo = (size_t)4 * d.B * oplane;
#pragma unroll
            for (int c0 = 0; c0 < kStemCo; c0 += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = fmaxf(fmaf(acc[c0 + j], scale[c0 + j], shift[c0 + j]), 0.f);
                split_store8(d.out + orow * kStemCo + c0, d.out + (orow + olo) * kStemCo + c0, v, d.out_terms == 2);
            }
        }
    }

    if (STATS) {
        // warp totals -> shared -> one fp64 atomic per (block, channel, moment)
#pragma unroll
        for (int co = 0; co < kStemCo; co++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1[co] += __shfl_xor_sync(0xffffffffu, s1[co], o);
                s2[co] += __shfl_xor_sync(0xffffffffu, s2[co], o);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int co = 0; co < kStemCo; co++) { red_s[warp][2 * co] = s1[co]; red_s[warp][2 * co + 1] = s2[co]; }
        }
        __syncthreads();
        if (threadIdx.x < 2 * kStemCo) {
            double t = 0;
            for (int w = 0; w < kStemThreads / 32; w++) t += (double)red_s[w][threadIdx.x];
            atomicAdd(d.stats + (size_t)b * kStemCo * 2 + threadIdx.x, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Head: one warp per pixel, fp32 weights in shared memory, fp16 hi+lo activations.
constexpr int kHeadThreads = 256;
constexpr int kHeadMaxCo = 8;

__global__ void __launch_bounds__(kHeadThreads) head_kernel(HeadDesc d)
{
    extern __shared__ float hw_s[];   // [Co][C]
    for (int i = threadIdx.x; i < d.Co * d.C; i += kHeadThreads) hw_s[i] = d.weight[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Wp = d.W + 2, plane = (d.H + 2) * Wp;
    const long long total = (long long)d.B * d.H * d.W;
    for (long long pix = blockIdx.x * (long long)(kHeadThreads / 32) + warp; pix < total;
         pix += (long long)gridDim.x * (kHeadThreads / 32)) {
        const int x = (int)(pix % d.W);
        const int y = (int)((pix / d.W) % d.H);
        const int b = (int)(pix / ((long long)d.W * d.H));
        const size_t row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
        float acc[kHeadMaxCo];
#pragma unroll
        for (int o = 0; o < kHeadMaxCo; o++) acc[o] = 0.f;
        for (int c = lane * 8; c < d.C; c += 256) {
            const uint4 hq = __ldg(reinterpret_cast<const uint4*>(d.act + row * d.C + c));
            const __half* hh = reinterpret_cast<const __half*>(&hq);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = __half2float(hh[j]);
            if (d.in_terms == 2) {
                const uint4 lq = __ldg(reinterpret_cast<const uint4*>(d.act + (row + (size_t)d.act_lo_rows) * d.C + c));
                const __half* ll = reinterpret_cast<const __half*>(&lq);
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] += __half2float(ll[j]);
            }
#pragma unroll
            for (int o = 0; o < kHeadMaxCo; o++) {
                if (o < d.Co) {
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[o] = fmaf(v[j], hw_s[o * d.C + c + j], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kHeadMaxCo; o++) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
        }
        if (lane < d.Co) {
            float r = 0.f;
#pragma unroll
            for (int o = 0; o < kHeadMaxCo; o++)
                if (o == lane) r = acc[o];
            r += d.bias[lane];
            if (lane < d.num_task) r += d.mean[lane];
            else r = expf(fminf(fmaxf(r, d.clamp_lo), d.clamp_hi));
            d.out[(((size_t)b * d.Co + lane) * d.H + y) * d.W + x] = r;
        }
    }
}

int sm_count()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

const char* last_error()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* gn_apply_launch(const GnApplyDesc& d, cudaStream_t stream)
{
    if (d.C % 8 != 0) return "gn_apply: C must be a multiple of 8";
    if (d.out_phases != 1 && d.out_phases != 4) return "gn_apply: out_phases must be 1 or 4";
    if (d.add_kind != 0 && d.out_phases != 1 && d.add_kind != 2) return "gn_apply: residual add needs a same-resolution output";
    const long long total = (long long)d.B * d.H * d.W * (d.C / 8);
    if (total == 0) return nullptr;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gn_apply_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d);
    return last_error();
}

static const char* stem_check(const StemDesc& d)
{
    if (d.Cin != 1 && d.Cin != 3) return "stem: Cin must be 1 or 3";
    if (d.B <= 0 || d.B > 65535) return "stem: batch size out of range";
    return nullptr;
}

const char* stem_stats_launch(const StemDesc& d, cudaStream_t stream)
{
    if (const char* e = stem_check(d)) return e;
    int bx = sm_count() * 4 / d.B;
    if (bx < 1) bx = 1;
    stem_kernel<true><<<dim3(bx, d.B), kStemThreads, 0, stream>>>(d);
    return last_error();
}

const char* stem_apply_launch(const StemDesc& d, cudaStream_t stream)
{
    if (const char* e = stem_check(d)) return e;
    const int segs = d.H * ((d.W + 31) / 32);
    int bx = (segs + 7) / 8;
    const int cap = sm_count() * 8 / d.B > 0 ? sm_count() * 8 / d.B : 1;
    if (bx > cap) bx = cap;
    stem_kernel<false><<<dim3(bx, d.B), kStemThreads, 0, stream>>>(d);
    return last_error();
}

const char* head_launch(const HeadDesc& d, cudaStream_t stream)
{
    if (d.Co < 1 || d.Co > kHeadMaxCo) return "head: 1..8 output channels";
    if (d.C % 8 != 0) return "head: C must be a multiple of 8";
    const size_t smem = (size_t)d.Co * d.C * sizeof(float);
    if (smem > 48 * 1024) return "head: weights exceed 48 KB of shared memory";
    const long long total = (long long)d.B * d.H * d.W;
    if (total == 0) return nullptr;
    long long blocks = (total + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    head_kernel<<<(unsigned)blocks, kHeadThreads, smem, stream>>>(d);
    return last_error();
}

}  // namespace cl
