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
#include <cuda_fp4.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "conv.h"
#include "fp4_planes.cuh"

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

// e4m3 operand planes of the fp16 + fp8 convolution mode: plane 0 = fp8(a_hi * 2^2), plane 1 = fp8((a - a_hi) * 2^14)
// (power-of-two scales keep both inside e4m3's normal range for activations up to ~50; saturating conversion)
__device__ __forceinline__ void fp8_store8(uint8_t* hi_ptr, uint8_t* lo_ptr, const float (&v)[8])
{
    __align__(8) __nv_fp8x2_storage_t h[4];
    __align__(8) __nv_fp8x2_storage_t l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float a0 = __half2float(__float2half_rn(v[2 * j])), a1 = __half2float(__float2half_rn(v[2 * j + 1]));
        h[j] = __nv_cvt_float2_to_fp8x2(make_float2(a0 * kAct8HiScale, a1 * kAct8HiScale), __NV_SATFINITE, __NV_E4M3);
        l[j] = __nv_cvt_float2_to_fp8x2(make_float2((v[2 * j] - a0) * kAct8LoScale, (v[2 * j + 1] - a1) * kAct8LoScale),
                                        __NV_SATFINITE, __NV_E4M3);
    }
    *reinterpret_cast<uint2*>(hi_ptr) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo_ptr) = *reinterpret_cast<const uint2*>(l);
}

// One work item = 8 consecutive channels of one interior pixel; every thread handles kGnUnroll items per
// grid-stride step so that several independent 16-byte loads are in flight.  The per-(image, group) mean and
// 1/sigma are derived once per block from the fp64 sums into shared memory.
constexpr int kGnThreads = 256;
constexpr int kGnUnroll = 4;

struct GnLane {
    int c;              // first of the 8 channels this thread owns (constant over the grid-stride loop)
    int gshift;         // log2(channels per group)
    float ga[8], be[8];
};

struct GnItem {
    size_t row;         // PF row of the pixel in the input geometry
    unsigned b;
    int y, x;
    float4 r0, r1;      // raw conv output
    uint4 a0, a1;       // ADD_KIND 1: residual hi / lo halves;  ADD_KIND 2: second raw tensor (as bits)
    bool live;
};

// phase 1: address arithmetic + all global loads of one item (no stores in between items: the loads of the
// kGnUnroll items of a thread are in flight together)
template <int ADD_KIND>
__device__ __forceinline__ void gn_load(const GnApplyDesc& d, int c, unsigned pix, unsigned total_pix, int Wp, int plane,
                                        GnItem& it)
{
    it.live = pix < total_pix;
    if (!it.live) return;
    const unsigned hw = (unsigned)(d.H * d.W);
    it.b = pix / hw;
    const unsigned rem = pix - it.b * hw;
    it.y = (int)(rem / (unsigned)d.W);
    it.x = (int)(rem - (unsigned)it.y * (unsigned)d.W);
    it.row = (size_t)it.b * plane + (size_t)(it.y + 1) * Wp + (it.x + 1);
    const float4* r4 = reinterpret_cast<const float4*>(d.raw + it.row * d.C + c);
    it.r0 = __ldg(r4);
    it.r1 = __ldg(r4 + 1);
    if (ADD_KIND == 1) {
        it.a0 = __ldg(reinterpret_cast<const uint4*>(d.res + it.row * d.C + c));
        it.a1 = d.res_lo_rows > 0 ? __ldg(reinterpret_cast<const uint4*>(d.res + (it.row + (size_t)d.res_lo_rows) * d.C + c))
                                  : make_uint4(0, 0, 0, 0);
    } else if (ADD_KIND == 2) {
        const uint4* q4 = reinterpret_cast<const uint4*>(d.raw2 + it.row * d.C + c);
        it.a0 = __ldg(q4);
        it.a1 = __ldg(q4 + 1);
    }
}

// phase 2: normalise, merge, split into fp16 hi / lo and store
template <int ADD_KIND>
__device__ __forceinline__ void gn_finish(const GnApplyDesc& d, const GnLane& t, const GnItem& it, int plane, int Wop,
                                          int oplane, int groups, const float2* tab1, const float2* tab2)
{
    if (!it.live) return;
    const int c = t.c;
    float v[8] = {it.r0.x, it.r0.y, it.r0.z, it.r0.w, it.r1.x, it.r1.y, it.r1.z, it.r1.w};
    if (d.group_ch) {
        const float2* tb = tab1 + it.b * groups;
        if (t.gshift >= 3) {   // the 8 channels share one group
            const float2 mr = tb[c >> t.gshift];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (v[j] - mr.x) * (mr.y * t.ga[j]) + t.be[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float2 mr = tb[(c + j) >> t.gshift];
                v[j] = (v[j] - mr.x) * (mr.y * t.ga[j]) + t.be[j];
            }
        }
    }
    if (d.relu_inner) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
    }
    if (ADD_KIND == 1) {
        const __half2* hh = reinterpret_cast<const __half2*>(&it.a0);
        const __half2* ll = reinterpret_cast<const __half2*>(&it.a1);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 a = __half22float2(hh[j]), bq = __half22float2(ll[j]);
            v[2 * j] += a.x + bq.x;
            v[2 * j + 1] += a.y + bq.y;
        }
    } else if (ADD_KIND == 2) {
        const float w[8] = {__uint_as_float(it.a0.x), __uint_as_float(it.a0.y), __uint_as_float(it.a0.z), __uint_as_float(it.a0.w),
                            __uint_as_float(it.a1.x), __uint_as_float(it.a1.y), __uint_as_float(it.a1.z), __uint_as_float(it.a1.w)};
        const float2* tb = tab2 + it.b * groups;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 mr = tb[(c + j) >> t.gshift];
            v[j] += (w[j] - mr.x) * (mr.y * __ldg(d.gamma2 + c + j)) + __ldg(d.beta2 + c + j);
        }
    }
    if (d.relu_outer) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
    }
    size_t orow, olo;
    if (d.out_phases == 1) {
        orow = it.row;
        olo = (size_t)d.B * plane;
    } else {
        const int ph = (it.y & 1) * 2 + (it.x & 1);
        orow = ((size_t)ph * d.B + it.b) * oplane + (size_t)(it.y / 2 + 1) * Wop + (it.x / 2 + 1);
        olo = (size_t)4 * d.B * oplane;
    }
    // out_C / out_c0: the destination may be a channel slice of a wider matrix (encoder outputs of the MLR model are
    // written side by side into one concatenated activation)
    const size_t oc = (size_t)d.out_C, o0 = (size_t)d.out_c0 + c;
    split_store8(d.out + orow * oc + o0, d.out + (orow + olo) * oc + o0, v, d.out_terms == 2);
    if (d.out8) fp8_store8(d.out8 + orow * oc + o0, d.out8 + (orow + olo) * oc + o0, v);
    if (d.out4) {   // the warp's 32 lanes share the pixel (C % 256 == 0): warp-collective, `live` is warp-uniform
        const size_t half_c = (size_t)d.C / 2;
        const uint32_t word = fp4_store8(d.out4 + orow * half_c + c / 2, d.out4 + (orow + olo) * half_c + c / 2, v);
        if ((c & 255) == 0) d.out_sf[(size_t)(c >> 8) * olo + orow] = word;
    }
}

template <int ADD_KIND>
__global__ void __launch_bounds__(kGnThreads, 2) gn_apply_kernel(GnApplyDesc d)
{
    extern __shared__ float2 gn_tab[];   // [B * groups] (mean, 1/sigma), twice when a second GroupNorm is merged
    const int groups = d.group_ch ? d.C / d.group_ch : 1;
    const float2* tab1 = gn_tab;
    const float2* tab2 = gn_tab + d.B * groups;
    if (d.group_ch) {
        const double count = (double)d.group_ch * d.H * d.W;
        for (int i = threadIdx.x; i < d.B * groups; i += kGnThreads) {
            float m, r;
            mean_rstd(d.stats, i / groups, groups, i % groups, count, d.eps, m, r);
            gn_tab[i] = make_float2(m, r);
            if (ADD_KIND == 2) {
                mean_rstd(d.stats2, i / groups, groups, i % groups, count, d.eps, m, r);
                gn_tab[d.B * groups + i] = make_float2(m, r);
            }
        }
        __syncthreads();
    }
    // C / 8 divides the block size, so a thread keeps the same 8 channels for every pixel it visits:
    // the affine parameters live in registers and only the pixel index advances.
    const unsigned c8n = (unsigned)(d.C / 8);
    const unsigned pix_per_block = kGnThreads / c8n;
    GnLane t;
    t.c = (int)(threadIdx.x % c8n) * 8;
    t.gshift = d.group_ch ? 31 - __clz(d.group_ch) : 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        t.ga[j] = d.group_ch ? __ldg(d.gamma + t.c + j) : 1.f;
        t.be[j] = d.group_ch ? __ldg(d.beta + t.c + j) : 0.f;
    }
    const unsigned total_pix = (unsigned)d.B * (unsigned)(d.H * d.W);
    const int Wp = d.W + 2, plane = (d.H + 2) * Wp;
    const int Ho = (d.H + 1) / 2, Wo = (d.W + 1) / 2, Wop = Wo + 2, oplane = (Ho + 2) * Wop;
    const unsigned stride = gridDim.x * pix_per_block;
    for (unsigned base = blockIdx.x * pix_per_block + threadIdx.x / c8n; base < total_pix; base += stride * kGnUnroll) {
        GnItem items[kGnUnroll];
#pragma unroll
        for (int u = 0; u < kGnUnroll; u++) gn_load<ADD_KIND>(d, t.c, base + (unsigned)u * stride, total_pix, Wp, plane, items[u]);
#pragma unroll
        for (int u = 0; u < kGnUnroll; u++) gn_finish<ADD_KIND>(d, t, items[u], plane, Wop, oplane, groups, tab1, tab2);
    }
}

// ---- row-wise variant for the wide same-resolution layers (C % 256 == 0, groups of >= 8 channels, full-width output).
// The generic kernel above spends ~500 instructions per 8-channel item (pixel index divisions, per-item table look-ups,
// 64-bit address arithmetic for every variant) and is issue bound at 45-60 % of the HBM rate.  Here a block owns one
// image row: no divisions, the lane's affine parameters AND its group's (mean, 1/sigma) live in registers, a warp walks
// along x with constant pointer strides (one pixel x 256 channels per step: 1 KB coalesced loads), four pixels in flight.
// PLANES: bit 0 = e4m3 planes, bit 1 = block-scaled e2m1 planes.  Same arithmetic, bit for bit, as gn_apply_kernel.
constexpr int kRowsU = 4;   // pixels a warp keeps in flight per batch

template <int ADD_KIND>
struct RowsBatch {
    float4 r0[kRowsU], r1[kRowsU];
    uint4 a0[kRowsU], a1[kRowsU];
};

template <int ADD_KIND>
__device__ __forceinline__ void rows_load(RowsBatch<ADD_KIND>& q, const GnApplyDesc& d, const float* praw, const __half* pres,
                                          const float* praw2, size_t estep, size_t res_lo, int x0, int xstep)
{
#pragma unroll
    for (int u = 0; u < kRowsU; u++) {
        if (x0 + u * xstep < d.W) {
            const float4* r4 = reinterpret_cast<const float4*>(praw + u * estep);
            q.r0[u] = __ldg(r4);
            q.r1[u] = __ldg(r4 + 1);
            if (ADD_KIND == 1) {
                q.a0[u] = __ldg(reinterpret_cast<const uint4*>(pres + u * estep));
                q.a1[u] = d.res_lo_rows > 0 ? __ldg(reinterpret_cast<const uint4*>(pres + u * estep + res_lo)) : make_uint4(0, 0, 0, 0);
            } else if (ADD_KIND == 2) {
                const uint4* q4 = reinterpret_cast<const uint4*>(praw2 + u * estep);
                q.a0[u] = __ldg(q4);
                q.a1[u] = __ldg(q4 + 1);
            }
        }
    }
}

template <int ADD_KIND, int PLANES>
__global__ void __launch_bounds__(256, 2) gn_apply_rows_kernel(GnApplyDesc d)
{
    constexpr int U = kRowsU;
    // register double buffering of the loads; the merging variants hold twice the operands and would spill
    constexpr bool kPrefetch = ADD_KIND == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kgroups = d.C >> 8;
    const int kg = warp % kgroups;
    const int xstep = 8 / kgroups;   // pixels the block's 8 warps cover per step
    const int b = blockIdx.x / d.H, y = blockIdx.x - b * d.H;
    const int c = kg * 256 + lane * 8;
    const int Wp = d.W + 2;
    const size_t rows = (size_t)d.B * (d.H + 2) * Wp;
    // running pointers of the warp's current pixel; one pixel step = xstep rows of the padded-flat matrices
    const int xw = warp / kgroups;
    const size_t row_first = ((size_t)b * (d.H + 2) + (y + 1)) * Wp + 1 + xw;
    const size_t e_first = row_first * d.C + c;
    const size_t estep = (size_t)xstep * d.C;        // elements per pixel step
    const size_t lo_off = rows * d.C;                // element distance of the lo planes
    const float* praw = d.raw + e_first;
    const __half* pres = ADD_KIND == 1 ? d.res + e_first : nullptr;
    const size_t res_lo = (size_t)d.res_lo_rows * d.C;
    const float* praw2 = ADD_KIND == 2 ? d.raw2 + e_first : nullptr;
    // the first batch of loads does not depend on the statistics: issue it before the fp64 mean / 1/sigma arithmetic
    RowsBatch<ADD_KIND> cur, nxt;
    rows_load<ADD_KIND>(cur, d, praw, pres, praw2, estep, res_lo, xw, xstep);
    float ga[8], be[8], mean = 0.f;
    float ga2[8], be2[8], mean2 = 0.f;
    {
        float rstd = 1.f;
        if (d.group_ch) {
            const int groups = d.C / d.group_ch;
            mean_rstd(d.stats, b, groups, c / d.group_ch, (double)d.group_ch * d.H * d.W, d.eps, mean, rstd);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            ga[j] = d.group_ch ? rstd * __ldg(d.gamma + c + j) : 1.f;
            be[j] = d.group_ch ? __ldg(d.beta + c + j) : 0.f;
        }
        if (ADD_KIND == 2) {
            float rstd2 = 1.f;
            const int groups = d.C / d.group_ch;
            mean_rstd(d.stats2, b, groups, c / d.group_ch, (double)d.group_ch * d.H * d.W, d.eps, mean2, rstd2);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                ga2[j] = rstd2 * __ldg(d.gamma2 + c + j);
                be2[j] = __ldg(d.beta2 + c + j);
            }
        }
    }
    __half* pout = d.out + e_first;
    uint8_t* pout8 = (PLANES & 1) ? d.out8 + e_first : nullptr;
    uint8_t* pout4 = (PLANES & 2) ? d.out4 + e_first / 2 : nullptr;
    uint32_t* psf = (PLANES & 2) ? d.out_sf + (size_t)kg * rows + row_first : nullptr;
    const bool write_lo = d.out_terms == 2;
    for (int x0 = xw; x0 < d.W; x0 += xstep * U) {
        // (plain variant) the next batch is in flight while this one is normalised and stored
        praw += U * estep;
        if (ADD_KIND == 1) pres += U * estep;
        if (ADD_KIND == 2) praw2 += U * estep;
        const bool more = x0 + xstep * U < d.W;
        if (kPrefetch && more) rows_load<ADD_KIND>(nxt, d, praw, pres, praw2, estep, res_lo, x0 + xstep * U, xstep);
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (x0 + u * xstep >= d.W) break;   // warp-uniform
            float v[8] = {cur.r0[u].x, cur.r0[u].y, cur.r0[u].z, cur.r0[u].w, cur.r1[u].x, cur.r1[u].y, cur.r1[u].z, cur.r1[u].w};
            if (d.group_ch) {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = (v[j] - mean) * ga[j] + be[j];
            }
            if (d.relu_inner) {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
            }
            if (ADD_KIND == 1) {
                const __half2* hh = reinterpret_cast<const __half2*>(&cur.a0[u]);
                const __half2* ll = reinterpret_cast<const __half2*>(&cur.a1[u]);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 a = __half22float2(hh[j]), bq = __half22float2(ll[j]);
                    v[2 * j] += a.x + bq.x;
                    v[2 * j + 1] += a.y + bq.y;
                }
            } else if (ADD_KIND == 2) {
                const float w[8] = {__uint_as_float(cur.a0[u].x), __uint_as_float(cur.a0[u].y), __uint_as_float(cur.a0[u].z), __uint_as_float(cur.a0[u].w),
                                    __uint_as_float(cur.a1[u].x), __uint_as_float(cur.a1[u].y), __uint_as_float(cur.a1[u].z), __uint_as_float(cur.a1[u].w)};
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] += (w[j] - mean2) * ga2[j] + be2[j];
            }
            if (d.relu_outer) {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
            }
            __half* o = pout + u * estep;
            split_store8(o, o + lo_off, v, write_lo);
            if (PLANES & 1) fp8_store8(pout8 + u * estep, pout8 + u * estep + lo_off, v);
            if (PLANES & 2) {
                uint8_t* o4 = pout4 + u * (estep / 2);
                const uint32_t word = fp4_store8(o4, o4 + lo_off / 2, v);
                if (lane == 0) psf[u * xstep] = word;
            }
        }
        pout += U * estep;
        if (PLANES & 1) pout8 += U * estep;
        if (PLANES & 2) { pout4 += U * (estep / 2); psf += U * xstep; }
        if (kPrefetch) cur = nxt;
        else if (more) rows_load<ADD_KIND>(cur, d, praw, pres, praw2, estep, res_lo, x0 + xstep * U, xstep);
    }
}

template <int ADD_KIND>
void gn_apply_rows_dispatch(const GnApplyDesc& d, int planes, cudaStream_t stream)
{
    const unsigned grid = (unsigned)(d.B * d.H);
    switch (planes) {
        case 0: gn_apply_rows_kernel<ADD_KIND, 0><<<grid, 256, 0, stream>>>(d); break;
        case 1: gn_apply_rows_kernel<ADD_KIND, 1><<<grid, 256, 0, stream>>>(d); break;
        case 2: gn_apply_rows_kernel<ADD_KIND, 2><<<grid, 256, 0, stream>>>(d); break;
        default: gn_apply_rows_kernel<ADD_KIND, 3><<<grid, 256, 0, stream>>>(d); break;
    }
}

// ---------------------------------------------------------------------------------------------
// Stem: direct 3x3 convolution on the CUDA cores (K = 27 is too thin for the tensor pipe), weights broadcast
// from shared memory.  One thread = two vertically adjacent output pixels x 32 channels, so every weight
// vector fetched feeds 64 FMAs and the 4x3 input patch is shared.  STATS pass: per-(image, channel) sum and sum
// of squares (reduce-scatter over the warp after every segment, two running registers per lane, one fp64
// atomic per block and moment).  APPLY pass: recomputes the convolution, normalises, applies ReLU and writes
// the fp16 hi/lo planes straight into the four-phase input of conv2 -- the 32-channel full-resolution fp32
// tensor (1.4 GB at 32 frames) is never materialised.
constexpr int kStemThreads = 256;
constexpr int kStemCo = 32;

// recursive-halving butterfly over 64 interleaved (sum, sum of squares) values: on return lane l holds the
// totals of channel stem_channel_of(l) in v[0], v[1]
__device__ __forceinline__ void stem_reduce_scatter(float (&v)[2 * kStemCo], int lane)
{
    int cur = 2 * kStemCo;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const int half = cur >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < kStemCo; i++) {
            if (i < half) {
                const float send = upper ? v[i] : v[i + half];
                const float keep = upper ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        cur = half;
    }
}
__device__ __forceinline__ int stem_channel_of(int lane)
{
    return ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
}

template <bool STATS>
__global__ void __launch_bounds__(kStemThreads, 2) stem_kernel(StemDesc d)
{
    __shared__ __align__(16) float w_s[27 * kStemCo];   // [ci*9 + kh*3 + kw][co]
    __shared__ float b_s[kStemCo];
    __shared__ float2 affine_s[kStemCo];                // APPLY: y = max(x * scale + shift, 0)
    __shared__ float red_s[kStemThreads / 32][2 * kStemCo];
    const int taps = d.Cin * 9;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < taps * kStemCo; i += kStemThreads) {
        const int co = i % kStemCo, t = i / kStemCo;
        w_s[i] = d.weight[co * taps + t];
    }
    if (threadIdx.x < kStemCo) {
        b_s[threadIdx.x] = d.bias[threadIdx.x];
        if (!STATS) {
            float scale = 1.f, shift = 0.f;
            if (d.has_gn) {
                float mean, rstd;
                mean_rstd(d.stats, b, kStemCo, threadIdx.x, (double)d.H * d.W, d.eps, mean, rstd);
                scale = rstd * d.gamma[threadIdx.x];
                shift = d.beta[threadIdx.x] - mean * scale;
            }
            affine_s[threadIdx.x] = make_float2(scale, shift);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int segs_x = (d.W + 31) / 32;
    const int pairs_y = (d.H + 1) / 2;
    const int segs = pairs_y * segs_x;
    const float* img = d.image + (size_t)b * d.Cin * d.H * d.W;
    const int Ho = (d.H + 1) / 2, Wo = (d.W + 1) / 2, Wop = Wo + 2, oplane = (Ho + 2) * Wop;
    float tot1 = 0.f, tot2 = 0.f;   // STATS: running totals of channel stem_channel_of(lane)

    for (int seg = blockIdx.x * (kStemThreads / 32) + warp; seg < segs; seg += gridDim.x * (kStemThreads / 32)) {
        const int yp = seg / segs_x, y0 = 2 * yp, x = (seg % segs_x) * 32 + lane;
        const bool inside = x < d.W;
        const bool row1 = y0 + 1 < d.H;
        float acc0[kStemCo], acc1[kStemCo];
#pragma unroll
        for (int co = 0; co < kStemCo; co++) { acc0[co] = b_s[co]; acc1[co] = b_s[co]; }
        for (int ci = 0; ci < d.Cin; ci++) {
            float in[4][3];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int yy = y0 + r - 1;
#pragma unroll
                for (int kw = 0; kw < 3; kw++) {
                    const int xx = x + kw - 1;
                    in[r][kw] = (inside && yy >= 0 && yy < d.H && xx >= 0 && xx < d.W)
                                    ? __ldg(img + ((size_t)ci * d.H + yy) * d.W + xx) : 0.f;
                }
            }
#pragma unroll
            for (int kh = 0; kh < 3; kh++) {
#pragma unroll
                for (int kw = 0; kw < 3; kw++) {
                    const float xv0 = in[kh][kw], xv1 = in[kh + 1][kw];
                    const float4* w4 = reinterpret_cast<const float4*>(w_s + (ci * 9 + kh * 3 + kw) * kStemCo);
#pragma unroll
                    for (int j = 0; j < kStemCo / 4; j++) {
                        const float4 w = w4[j];
                        acc0[4 * j + 0] = fmaf(xv0, w.x, acc0[4 * j + 0]);
                        acc0[4 * j + 1] = fmaf(xv0, w.y, acc0[4 * j + 1]);
                        acc0[4 * j + 2] = fmaf(xv0, w.z, acc0[4 * j + 2]);
                        acc0[4 * j + 3] = fmaf(xv0, w.w, acc0[4 * j + 3]);
                        acc1[4 * j + 0] = fmaf(xv1, w.x, acc1[4 * j + 0]);
                        acc1[4 * j + 1] = fmaf(xv1, w.y, acc1[4 * j + 1]);
                        acc1[4 * j + 2] = fmaf(xv1, w.z, acc1[4 * j + 2]);
                        acc1[4 * j + 3] = fmaf(xv1, w.w, acc1[4 * j + 3]);
                    }
                }
            }
        }
        if (STATS) {
            float v[2 * kStemCo];
            const float m0 = inside ? 1.f : 0.f, m1 = (inside && row1) ? 1.f : 0.f;
#pragma unroll
            for (int co = 0; co < kStemCo; co++) {
                const float a0 = acc0[co] * m0, a1 = acc1[co] * m1;
                v[2 * co] = a0 + a1;
                v[2 * co + 1] = a0 * a0 + a1 * a1;
            }
            stem_reduce_scatter(v, lane);
            tot1 += v[0];
            tot2 += v[1];
        } else if (inside) {
            const size_t olo = (size_t)4 * d.B * oplane;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (r == 1 && !row1) break;
                const int ph = r * 2 + (x & 1);
                const size_t orow = ((size_t)ph * d.B + b) * oplane + (size_t)(yp + 1) * Wop + (x / 2 + 1);
#pragma unroll
                for (int c0 = 0; c0 < kStemCo; c0 += 8) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float2 af = affine_s[c0 + j];
                        v[j] = fmaxf(fmaf(r == 0 ? acc0[c0 + j] : acc1[c0 + j], af.x, af.y), 0.f);
                    }
                    split_store8(d.out + orow * kStemCo + c0, d.out + (orow + olo) * kStemCo + c0, v, d.out_terms == 2);
                }
            }
        }
    }

    if (STATS) {
        const int ch = stem_channel_of(lane);
        red_s[warp][2 * ch] = tot1;
        red_s[warp][2 * ch + 1] = tot2;
        __syncthreads();
        if (threadIdx.x < 2 * kStemCo) {
            double t = 0;
            for (int w = 0; w < kStemThreads / 32; w++) t += (double)red_s[w][threadIdx.x];
            atomicAdd(d.stats + (size_t)b * kStemCo * 2 + threadIdx.x, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Head: one warp per pixel; every lane owns 8 (x2 for C > 256) fixed input channels and keeps their fp32 weights
// for all output channels in registers, so the per-pixel work is two 16-byte loads per plane, Co x 8 FMAs per
// slice and a shuffle reduction.
constexpr int kHeadThreads = 256;
constexpr int kHeadMaxCo = 8;
constexpr int kHeadMaxC = 512;
constexpr int kHeadUnroll = 2;

template <int CO>
__global__ void __launch_bounds__(kHeadThreads) head_kernel(HeadDesc d)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Wp = d.W + 2, plane = (d.H + 2) * Wp;
    const long long total = (long long)d.B * d.H * d.W;
    const int slices = (d.C + 255) / 256;   // 1 or 2
    float w[2][CO][8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int c = s * 256 + lane * 8;
#pragma unroll
        for (int o = 0; o < CO; o++)
#pragma unroll
            for (int j = 0; j < 8; j++) w[s][o][j] = (s < slices && c + j < d.C && o < d.Co) ? d.weight[o * d.C + c + j] : 0.f;
    }
    // kHeadUnroll pixels per warp and step, all their loads issued before the first use: the kernel is a pure stream of
    // 354 MB and with one pixel in flight per warp (105 registers, 16 warps per SM) it reached 1.6 TB/s only
    const long long step = (long long)gridDim.x * (kHeadThreads / 32);
    for (long long pix0 = blockIdx.x * (long long)(kHeadThreads / 32) + warp; pix0 < total; pix0 += step * kHeadUnroll) {
        uint4 hq[kHeadUnroll][2], lq[kHeadUnroll][2];
        size_t rows[kHeadUnroll];
#pragma unroll
        for (int u = 0; u < kHeadUnroll; u++) {
            const long long pix = pix0 + u * step;
            const long long pc = pix < total ? pix : pix0;
            const int x = (int)(pc % d.W);
            const int y = (int)((pc / d.W) % d.H);
            const int b = (int)(pc / ((long long)d.W * d.H));
            rows[u] = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int c = s * 256 + lane * 8;
                hq[u][s] = make_uint4(0, 0, 0, 0);
                lq[u][s] = make_uint4(0, 0, 0, 0);
                if (s < slices && c < d.C) {
                    hq[u][s] = __ldg(reinterpret_cast<const uint4*>(d.act + rows[u] * d.C + c));
                    if (d.in_terms == 2)
                        lq[u][s] = __ldg(reinterpret_cast<const uint4*>(d.act + (rows[u] + (size_t)d.act_lo_rows) * d.C + c));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kHeadUnroll; u++) {
            const long long pix = pix0 + u * step;
            if (pix >= total) break;
            const int x = (int)(pix % d.W);
            const int y = (int)((pix / d.W) % d.H);
            const int b = (int)(pix / ((long long)d.W * d.H));
            float acc[CO];
#pragma unroll
            for (int o = 0; o < CO; o++) acc[o] = 0.f;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const __half* hh = reinterpret_cast<const __half*>(&hq[u][s]);
                const __half* ll = reinterpret_cast<const __half*>(&lq[u][s]);
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = __half2float(hh[j]) + __half2float(ll[j]);
#pragma unroll
                for (int o = 0; o < CO; o++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[o] = fmaf(v[j], w[s][o][j], acc[o]);
            }
#pragma unroll
            for (int o = 0; o < CO; o++) {
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], sft);
            }
            if (lane < d.Co) {
                float r = 0.f;
#pragma unroll
                for (int o = 0; o < CO; o++)
                    if (o == lane) r = acc[o];
                r += d.bias[lane];
                if (lane < d.num_task) r += d.mean[lane];
                else r = expf(fminf(fmaxf(r, d.clamp_lo), d.clamp_hi));
                d.out[(((size_t)b * d.Co + lane) * d.H + y) * d.W + x] = r;
            }
        }
    }
}

// Full-size output head (networks.py:259-273, 344-358): GroupNorm + ReLU of the DUC convolution, PixelShuffle(rate),
// bilinear resize (align_corners = False, as F.interpolate) to the frame size, 1x1 fc3, mean offset / exp(clamp).
// One thread per output pixel; the shuffled map is never materialised: a source value at (c, sy, sx) is channel
// c * rate^2 + (sy % rate) * rate + sx % rate of cell (sy / rate, sx / rate) of the raw DUC output.
constexpr int kDucThreads = 256;
constexpr int kDucMaxCo = 8;

__global__ void __launch_bounds__(kDucThreads) duc_head_kernel(DucHeadDesc d)
{
    extern __shared__ float2 duc_tab[];   // [B][groups] (mean, rstd)
    const int groups = d.C / d.group_ch;
    const double count = (double)d.group_ch * d.Hc * d.Wc;
    for (int i = threadIdx.x; i < d.B * groups; i += kDucThreads) {
        float mean, rstd;
        mean_rstd(d.stats, i / groups, groups, i % groups, count, d.eps, mean, rstd);
        duc_tab[i] = make_float2(mean, rstd);
    }
    __syncthreads();
    const int Wp = d.Wc + 2, plane = (d.Hc + 2) * Wp;
    const int Hs = d.Hc * d.rate, Ws = d.Wc * d.rate, r2 = d.rate * d.rate;
    const float rh = (float)Hs / (float)d.Ho, rw = (float)Ws / (float)d.Wo;   // area_pixel_compute_scale
    const long long total = (long long)d.B * d.Ho * d.Wo;
    for (long long pix = blockIdx.x * (long long)kDucThreads + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * kDucThreads) {
        const int X = (int)(pix % d.Wo);
        const int Y = (int)((pix / d.Wo) % d.Ho);
        const int b = (int)(pix / ((long long)d.Wo * d.Ho));
        const float sy = fmaxf(rh * ((float)Y + 0.5f) - 0.5f, 0.f), sx = fmaxf(rw * ((float)X + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
        const float ly = sy - (float)y0, lx = sx - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float2* tab = duc_tab + b * groups;
        auto fetch = [&](int c, int yy, int xx) -> float {
            const int ch = c * r2 + (yy % d.rate) * d.rate + (xx % d.rate);
            const size_t row = (size_t)b * plane + (size_t)(yy / d.rate + 1) * Wp + (xx / d.rate + 1);
            const float2 mr = tab[ch / d.group_ch];
            const float v = (__ldg(d.raw + row * d.C + ch) - mr.x) * (mr.y * __ldg(d.gamma + ch)) + __ldg(d.beta + ch);
            return fmaxf(v, 0.f);
        };
        float val[kDucMaxCo];
#pragma unroll
        for (int c = 0; c < kDucMaxCo; c++) {
            if (c < d.Co) {
                // same expression order as ATen's upsample_bilinear2d kernel
                const float v00 = fetch(c, y0, x0);
                const float v01 = lx != 0.f ? fetch(c, y0, x1) : v00;
                float bottom = 0.f;
                if (ly != 0.f) {
                    const float v10 = fetch(c, y1, x0);
                    const float v11 = lx != 0.f ? fetch(c, y1, x1) : v10;
                    bottom = ly * (hx * v10 + lx * v11);
                }
                val[c] = hy * (hx * v00 + lx * v01) + bottom;
            } else {
                val[c] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < kDucMaxCo; k++) {
            if (k < d.Co) {
                float r = __ldg(d.bias + k);
#pragma unroll
                for (int c = 0; c < kDucMaxCo; c++)
                    if (c < d.Co) r = fmaf(__ldg(d.weight + k * d.Co + c), val[c], r);
                if (k < d.num_task) r += __ldg(d.mean + k);
                else r = expf(fminf(fmaxf(r, d.clamp_lo), d.clamp_hi));
                d.out[(((size_t)b * d.Co + k) * d.Ho + Y) * d.Wo + X] = r;
            }
        }
    }
}

// Input frames: uint8 HWC (what PIL / the decoders deliver) -> fp32 NCHW, x / 255 [then (v - mean[c]) / std[c]] with the
// same fp32 operations as torchvision's ToTensor / Normalize (dataloader/dataloader.py:189-212), so the result is
// bit-identical to the host transform it replaces while the host-to-device copy moves a quarter of the bytes.
__global__ void __launch_bounds__(256) frames_to_nchw_kernel(const uint8_t* __restrict__ frames, int B, int H, int W, int C,
                                                            const float* __restrict__ mean, const float* __restrict__ stdv,
                                                            float* __restrict__ out)
{
    const long long total = (long long)B * H * W;
    const long long hw = (long long)H * W;
    for (long long pix = blockIdx.x * 256ll + threadIdx.x; pix < total; pix += (long long)gridDim.x * 256) {
        const long long b = pix / hw, r = pix - b * hw;
        for (int c = 0; c < C; c++) {
            float v = __fdiv_rn((float)frames[pix * C + c], 255.f);
            if (mean) v = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
            out[(b * C + c) * hw + r] = v;
        }
    }
}

// GroupNorm of a padded-flat fp16 hi/lo activation (no ReLU): mlr_norm of the MLR model normalises the CONCATENATED
// encoder outputs (networks.py:421-439, 491-494), i.e. a tensor that is not the raw output of a convolution.  Pass 1
// accumulates sum / sum of squares per (image, group) in fp64, pass 2 writes the normalised hi / lo (+ e4m3) planes.
// Any group size (the 384-channel tiny model has 12 channels per group): the group of every channel is looked up.
constexpr int kPfGnMaxThreads = 256;

// blockDim = chunks * pslots (chunks = C / 8 <= 256): a thread keeps the same 8 channels for every pixel it visits.
// Statistics: per-channel fp32 partial sums in registers, a fixed-order block reduction per group in fp64, one fp64
// atomic per (block, group, moment) -- the result does not depend on thread scheduling beyond fp64 round-off.
template <bool APPLY>
__global__ void __launch_bounds__(kPfGnMaxThreads) pf_groupnorm_kernel(PfGroupNormDesc d)
{
    __shared__ float red[kPfGnMaxThreads][17];
    __shared__ float2 tab[512];             // APPLY: (mean, rstd) per group of this image
    const int b = blockIdx.y;
    const int groups = d.C / d.group_ch;
    const int chunks = d.C / 8;
    const int pslots = blockDim.x / chunks;
    const int chunk = threadIdx.x % chunks, pslot = threadIdx.x / chunks;
    const int c = chunk * 8;
    const int Wp = d.W + 2;
    const size_t plane = (size_t)(d.H + 2) * Wp;
    float be[8], mu[8], rs[8];
    if (APPLY) {
        const double count = (double)d.group_ch * d.H * d.W;
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            float m, r;
            mean_rstd(d.stats, b, groups, g, count, d.eps, m, r);
            tab[g] = make_float2(m, r);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 mr = tab[(c + j) / d.group_ch];
            mu[j] = mr.x;
            rs[j] = mr.y * __ldg(d.gamma + c + j);
            be[j] = __ldg(d.beta + c + j);
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { s1[j] = 0.f; s2[j] = 0.f; }
    const int total = d.H * d.W;
    const size_t olo = (size_t)d.B * plane;
    for (int pix = blockIdx.x * pslots + pslot; pix < total; pix += gridDim.x * pslots) {
        const int y = pix / d.W, x = pix - y * d.W;
        const size_t row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
        const uint4 hq = __ldg(reinterpret_cast<const uint4*>(d.in + row * d.C + c));
        const uint4 lq = __ldg(reinterpret_cast<const uint4*>(d.in + (row + (size_t)d.in_lo_rows) * d.C + c));
        const __half* hh = reinterpret_cast<const __half*>(&hq);
        const __half* ll = reinterpret_cast<const __half*>(&lq);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __half2float(hh[j]) + __half2float(ll[j]);
        if (!APPLY) {
#pragma unroll
            for (int j = 0; j < 8; j++) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (v[j] - mu[j]) * rs[j] + be[j];
            split_store8(d.out + row * d.C + c, d.out + (row + olo) * d.C + c, v, d.out_terms == 2);
            if (d.out8) fp8_store8(d.out8 + row * d.C + c, d.out8 + (row + olo) * d.C + c, v);
        }
    }
    if (!APPLY) {
#pragma unroll
        for (int j = 0; j < 8; j++) { red[threadIdx.x][j] = s1[j]; red[threadIdx.x][8 + j] = s2[j]; }
        __syncthreads();
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            double t1 = 0, t2 = 0;
            for (int ch = g * d.group_ch; ch < (g + 1) * d.group_ch; ch++) {
                for (int p = 0; p < pslots; p++) {
                    t1 += (double)red[p * chunks + (ch >> 3)][ch & 7];
                    t2 += (double)red[p * chunks + (ch >> 3)][8 + (ch & 7)];
                }
            }
            atomicAdd(d.stats + ((size_t)b * groups + g) * 2, t1);
            atomicAdd(d.stats + ((size_t)b * groups + g) * 2 + 1, t2);
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

// ---------------------------------------------------------------------------------------------
// Filter planes of the fp16 + fp4 convolution mode: e2m1 [2][tap][Cout][Cin / 2] = fp4(w_hi), fp4(w_lo) with one
// ue8m0 scale per (tap, output channel, 256 input channels) and plane; the scale words (hi, hi, lo, lo) are stored in
// the 32-row interleave tcgen05.cp copies into tensor memory.  One warp per (tap, output channel, 256-channel block).
__global__ void __launch_bounds__(256) pack_conv_fp4_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float scale,
                                                            uint8_t* __restrict__ w4, uint32_t* __restrict__ w_sf)
{
    const int lane = threadIdx.x & 31;
    const int kgroups = Cin / 256;
    const long long units = (long long)taps * Cout * kgroups;
    const size_t plane_bytes = (size_t)taps * Cout * (Cin / 2);
    for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int kg = (int)(u % kgroups);
        const int co = (int)((u / kgroups) % Cout);
        const int tap = (int)(u / ((long long)kgroups * Cout));
        const int ci = kg * 256 + lane * 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = w[((size_t)co * Cin + ci + j) * taps + tap] * scale;
        const size_t byte = ((size_t)tap * Cout + co) * (Cin / 2) + ci / 2;
        const uint32_t word = fp4_store8(w4 + byte, w4 + plane_bytes + byte, v, true);   // (lo, lo, hi, hi)
        if (lane == 0)
            w_sf[((size_t)(tap * kgroups + kg) * (Cout / 128) + co / 128) * 128 + (co & 31) * 4 + ((co & 127) >> 5)] = (word >> 16) | (word << 16);
    }
}

const char* last_error()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* gn_apply_launch(const GnApplyDesc& desc, cudaStream_t stream)
{
    GnApplyDesc d = desc;
    if (d.out_C == 0) d.out_C = d.C;
    if (d.out_C < d.C || d.out_c0 < 0 || d.out_c0 + d.C > d.out_C || d.out_C % 8 != 0 || d.out_c0 % 8 != 0)
        return "gn_apply: invalid destination channel slice";
    if (d.out_C != d.C && d.out_phases != 1) return "gn_apply: a channel slice needs a same-resolution output";
    if (d.C % 8 != 0) return "gn_apply: C must be a multiple of 8";
    if (d.out_phases != 1 && d.out_phases != 4) return "gn_apply: out_phases must be 1 or 4";
    if (d.add_kind == 1 && d.out_phases != 1) return "gn_apply: residual add needs a same-resolution output";
    if (d.out8 && d.out_phases != 1) return "gn_apply: e4m3 planes are only produced for same-resolution outputs";
    if (d.out4 && (d.out_phases != 1 || d.C % 256 != 0 || d.out_C != d.C || !d.out_sf))
        return "gn_apply: e2m1 planes need a same-resolution, full-width output with C % 256 == 0 and a scale buffer";
    const long long total = (long long)d.B * d.H * d.W * (d.C / 8);
    if (total == 0) return nullptr;
    if (total >= (1ll << 31)) return "gn_apply: tensor too large for 32-bit item indices";
    if (kGnThreads % (d.C / 8) != 0) return "gn_apply: C / 8 must divide 256 (C in {8, 16, ..., 2048} powers of two)";
    if (d.group_ch & (d.group_ch - 1)) return "gn_apply: channels per group must be a power of two";
    const int groups = d.group_ch ? d.C / d.group_ch : 0;
    const size_t smem = (size_t)d.B * groups * sizeof(float2) * (d.add_kind == 2 ? 2 : 1);
    if (smem > 48 * 1024) return "gn_apply: batch * groups exceeds the 48 KB statistics table";
    const int kgroups = d.C / 256;
    if (d.out_phases == 1 && d.C % 256 == 0 && (kgroups == 1 || kgroups == 2 || kgroups == 4 || kgroups == 8) && d.out_C == d.C &&
        d.group_ch % 8 == 0 && !(d.add_kind == 2 && !d.group_ch)) {
        const int planes = (d.out8 ? 1 : 0) | (d.out4 ? 2 : 0);
        if (d.add_kind == 0) gn_apply_rows_dispatch<0>(d, planes, stream);
        else if (d.add_kind == 1) gn_apply_rows_dispatch<1>(d, planes, stream);
        else gn_apply_rows_dispatch<2>(d, planes, stream);
        return last_error();
    }
    long long blocks = (total + kGnThreads * kGnUnroll - 1) / (kGnThreads * kGnUnroll);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (d.add_kind == 0) gn_apply_kernel<0><<<(unsigned)blocks, kGnThreads, smem, stream>>>(d);
    else if (d.add_kind == 1) gn_apply_kernel<1><<<(unsigned)blocks, kGnThreads, smem, stream>>>(d);
    else gn_apply_kernel<2><<<(unsigned)blocks, kGnThreads, smem, stream>>>(d);
    return last_error();
}

const char* pack_conv_fp4_launch(const float* w, int Cout, int Cin, int taps, float scale, uint8_t* w4, uint32_t* w_sf, cudaStream_t stream)
{
    if (Cin % 256 != 0 || Cout % 128 != 0 || taps < 1) return "pack_conv_fp4: Cin % 256 == 0 and Cout % 128 == 0 required";
    const long long units = (long long)taps * Cout * (Cin / 256);
    long long blocks = (units + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_conv_fp4_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w, Cout, Cin, taps, scale, w4, w_sf);
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
    int bx = sm_count() * 8 / d.B;
    if (bx < 1) bx = 1;
    stem_kernel<true><<<dim3(bx, d.B), kStemThreads, 0, stream>>>(d);
    return last_error();
}

const char* stem_apply_launch(const StemDesc& d, cudaStream_t stream)
{
    if (const char* e = stem_check(d)) return e;
    const int segs = ((d.H + 1) / 2) * ((d.W + 31) / 32);
    int bx = (segs + 7) / 8;
    const int cap = sm_count() * 8 / d.B > 0 ? sm_count() * 8 / d.B : 1;
    if (bx > cap) bx = cap;
    stem_kernel<false><<<dim3(bx, d.B), kStemThreads, 0, stream>>>(d);
    return last_error();
}

const char* head_launch(const HeadDesc& d, cudaStream_t stream)
{
    if (d.Co < 1 || d.Co > kHeadMaxCo) return "head: 1..8 output channels";
    if (d.C % 8 != 0 || d.C > kHeadMaxC) return "head: C must be a multiple of 8 and at most 512";
    const long long total = (long long)d.B * d.H * d.W;
    if (total == 0) return nullptr;
    long long blocks = (total + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (d.Co <= 4) head_kernel<4><<<(unsigned)blocks, kHeadThreads, 0, stream>>>(d);
    else head_kernel<8><<<(unsigned)blocks, kHeadThreads, 0, stream>>>(d);
    return last_error();
}

const char* duc_head_launch(const DucHeadDesc& d, cudaStream_t stream)
{
    if (d.Co < 1 || d.Co > kDucMaxCo) return "duc_head: 1..8 output channels";
    if (d.rate < 1 || d.C != d.Co * d.rate * d.rate) return "duc_head: C must be Co * rate^2";
    if (d.group_ch < 1 || d.C % d.group_ch != 0) return "duc_head: the DUC convolution is followed by a GroupNorm";
    if (d.Ho < 1 || d.Wo < 1) return "duc_head: empty output";
    const long long total = (long long)d.B * d.Ho * d.Wo;
    if (total == 0) return nullptr;
    const size_t smem = (size_t)d.B * (d.C / d.group_ch) * sizeof(float2);
    if (smem > 48 * 1024) return "duc_head: batch * groups exceeds the 48 KB statistics table";
    long long blocks = (total + kDucThreads - 1) / kDucThreads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    duc_head_kernel<<<(unsigned)blocks, kDucThreads, smem, stream>>>(d);
    return last_error();
}

const char* frames_to_nchw_launch(const uint8_t* frames, int B, int H, int W, int C, const float* mean, const float* stdv,
                                  float* out, cudaStream_t stream)
{
    if (C < 1 || C > 4) return "frames_to_nchw: 1..4 channels";
    if ((mean == nullptr) != (stdv == nullptr)) return "frames_to_nchw: mean and std come together";
    const long long total = (long long)B * H * W;
    if (total <= 0) return "frames_to_nchw: empty batch";
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    frames_to_nchw_kernel<<<(unsigned)blocks, 256, 0, stream>>>(frames, B, H, W, C, mean, stdv, out);
    return last_error();
}

const char* pf_groupnorm_launch(const PfGroupNormDesc& d, cudaStream_t stream)
{
    if (d.C % 8 != 0 || d.group_ch < 1 || d.C % d.group_ch != 0) return "pf_groupnorm: C must be a multiple of 8 and of the group size";
    if (d.C / 8 > kPfGnMaxThreads) return "pf_groupnorm: at most 2048 channels";
    if (d.C / d.group_ch > 512) return "pf_groupnorm: at most 512 groups";
    if (d.B <= 0 || d.B > 65535 || d.H <= 0 || d.W <= 0) return "pf_groupnorm: invalid sizes";
    const int chunks = d.C / 8;
    const int pslots = kPfGnMaxThreads / chunks;
    const int threads = chunks * pslots;
    long long bx = ((long long)d.H * d.W + pslots * 8 - 1) / (pslots * 8);
    const long long cap = ((long long)sm_count() * 6) / d.B;     // about one resident wave
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    pf_groupnorm_kernel<false><<<dim3((unsigned)bx, d.B), threads, 0, stream>>>(d);
    if (const char* e = last_error()) return e;
    pf_groupnorm_kernel<true><<<dim3((unsigned)bx, d.B), threads, 0, stream>>>(d);
    return last_error();
}

}  // namespace cl
