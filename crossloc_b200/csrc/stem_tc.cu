// Stem convolution on the tensor cores: conv1 (3x3, stride 1, Cin in {1, 3}, Cout = 32) + GroupNorm(32, 32) + ReLU,
// written as the 4-phase fp16 hi/lo padded-flat input of conv2
// (/root/reference/networks/networks.py:186-190, 231; Network.conv1 :59, 96 without normalisation).
//
// The CUDA-core version of this layer (cnn_pointwise.cu stem_kernel) issues 864 FMAs per pixel and pass; here a pixel
// costs one 32-element im2col row.  Per 128-pixel tile (a segment of one image row):
//   builder warps (4)  gather the 3x3xCin patch of every pixel (K = 27 padded to 32), split it into fp16 hi / lo and
//                      write it straight into shared memory in the K-major SWIZZLE_64B layout the tensor core reads
//                      (generic-proxy stores + fence.proxy.async), double buffered;
//   MMA warp           6 tcgen05.mma (128 x 32 x 16, fp16x3 split: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo) into one of two
//                      32-column TMEM accumulators;
//   epilogue warps (4) tcgen05.ld of the pixel's 32 channels -> pass 1: per-channel sum / sum of squares (butterfly
//                      reduction, fp64 atomics once per block); pass 2: normalise, ReLU, fp16 hi / lo split, two 64-byte
//                      stores into the parity phase of the pixel.
// The filter is scaled by a power of two derived from its maximum on the device (keeps w_lo out of fp16 subnormals).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv.h"
#include "ptx_sm100.cuh"

namespace cl {

namespace {

constexpr int kCo = 32;
constexpr int kTile = 128;                     // pixels per tile
constexpr int kThreads = 288;                  // 4 builder warps, 4 epilogue warps, 1 MMA / TMEM warp
constexpr uint32_t kPlaneBytes = kTile * 64;   // one fp16 plane of the A tile: 128 rows x 32 K
constexpr uint32_t kABytes = 2 * kPlaneBytes;  // hi + lo
constexpr uint32_t kBPlaneBytes = kCo * 64;
constexpr uint32_t kStageBytes = 4096;          // per epilogue warp: [plane hi/lo][x parity][16 rows][64 B]
constexpr uint32_t kSmemBytes = 2 * kABytes + 2 * kBPlaneBytes + 4 * kStageBytes + 1024;
constexpr uint32_t kTmemCols = 64;             // two 32-column accumulators

// byte offset of 16-byte chunk c (8 K-elements) of row r in a K-major tile with 64-byte rows, SWIZZLE_64B
__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo)
{
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        h[j] = __float2half_rn(v[j]);
        l[j] = __float2half_rn(v[j] - __half2float(h[j]));
    }
    hi = *reinterpret_cast<const uint4*>(h);
    lo = *reinterpret_cast<const uint4*>(l);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// recursive-halving butterfly over 64 interleaved (sum, sum of squares) values: on return lane l holds the totals of
// channel channel_of(l) in v[0], v[1]
__device__ __forceinline__ void reduce_scatter(float (&v)[2 * kCo], int lane)
{
    int cur = 2 * kCo;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const int half = cur >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < kCo; i++) {
            if (i < half) {
                const float send = upper ? v[i] : v[i + half];
                const float keep = upper ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        cur = half;
    }
}
__device__ __forceinline__ int channel_of(int lane)
{
    return ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
}

template <bool STATS>
__global__ void __launch_bounds__(kThreads, 2) stem_tc_kernel(StemDesc d)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], d_full[2], d_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float red_s[4][2 * kCo];
    __shared__ float bias_s[kCo];
    __shared__ float2 affine_s[kCo];
    __shared__ float amax_s[kThreads / 32];
    __shared__ float wscale_s[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int taps = d.Cin * 9;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + 2 * kABytes;

    // ---- prologue: barriers, TMEM, filter scale, B operand, per-channel affine
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(ptx::smem_u32(&a_full[s]), 128);
            ptx::mbar_init(ptx::smem_u32(&a_empty[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&d_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&d_empty[s]), 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 8) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_s), kTmemCols);
        ptx::tmem_relinquish();
    }
    {
        float m = 0.f;
        for (int i = threadIdx.x; i < kCo * taps; i += kThreads) m = fmaxf(m, fabsf(d.weight[i]));
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
        if (lane == 0) amax_s[warp] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < kThreads / 32; i++) m = fmaxf(m, amax_s[i]);
        int k = 0;
        if (m > 0.f && isfinite(m)) {
            k = (int)floorf(log2f(128.f / m));
            k = k < -24 ? -24 : (k > 24 ? 24 : k);
        }
        wscale_s[0] = exp2f((float)k);
        wscale_s[1] = exp2f((float)-k);
    }
    if (threadIdx.x < kCo) {
        bias_s[threadIdx.x] = d.bias[threadIdx.x];
        if (!STATS) {
            float scale = 1.f, shift = 0.f;
            if (d.has_gn) {
                const double cnt = (double)d.H * d.W;
                const double s = d.stats[((size_t)b * kCo + threadIdx.x) * 2], ss = d.stats[((size_t)b * kCo + threadIdx.x) * 2 + 1];
                const double mean = s / cnt;
                double var = ss / cnt - mean * mean;
                var = var > 0 ? var : 0;
                const float rstd = (float)(1.0 / sqrt(var + (double)d.eps));
                scale = rstd * d.gamma[threadIdx.x];
                shift = d.beta[threadIdx.x] - (float)mean * scale;
            }
            affine_s[threadIdx.x] = make_float2(scale, shift);
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {   // B operand: row = output channel, 4 chunks of 8 K-elements, hi and lo planes
        const int co = threadIdx.x >> 2, c = threadIdx.x & 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = c * 8 + j;
            v[j] = k < taps ? d.weight[co * taps + k] * wscale_s[0] : 0.f;
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        st_shared_v4(b_base + sw64(co, c), hi);
        st_shared_v4(b_base + kBPlaneBytes + sw64(co, c), lo);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int tiles_x = (d.W + kTile - 1) / kTile;
    const int tiles = d.H * tiles_x;
    const float* img = d.image + (size_t)b * d.Cin * d.H * d.W;

    if (warp < 4) {
        // ------------------------------------------------------------------ builders: im2col rows into shared memory
        const int r = threadIdx.x;   // pixel slot of the tile = row of the A operand
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
            const int s = it & 1;
            const uint32_t par = (uint32_t)(it >> 1) & 1u;
            const int y = t / tiles_x, x = (t - y * tiles_x) * kTile + r;
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = 0.f;
            if (x < d.W) {
#pragma unroll
                for (int ci = 0; ci < 3; ci++) {   // k = ci * 9 + kh * 3 + kw, fully unrolled: v[] stays in registers
                    if (ci < d.Cin) {
#pragma unroll
                        for (int kh = 0; kh < 3; kh++) {
                            const int yy = y + kh - 1;
#pragma unroll
                            for (int kw = 0; kw < 3; kw++) {
                                const int xx = x + kw - 1;
                                v[ci * 9 + kh * 3 + kw] = (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W)
                                                              ? __ldg(img + ((size_t)ci * d.H + yy) * d.W + xx) : 0.f;
                            }
                        }
                    }
                }
            }
            ptx::mbar_wait(ptx::smem_u32(&a_empty[s]), par ^ 1u);
            const uint32_t a_hi = smem_base + (uint32_t)s * kABytes, a_lo = a_hi + kPlaneBytes;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float w8[8];
#pragma unroll
                for (int j = 0; j < 8; j++) w8[j] = v[c * 8 + j];
                uint4 hi, lo;
                split8(w8, hi, lo);
                st_shared_v4(a_hi + sw64(r, c), hi);
                st_shared_v4(a_lo + sw64(r, c), lo);
            }
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(ptx::smem_u32(&a_full[s]));
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = ptx::make_idesc_f16(kTile, kCo);
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
            const int s = it & 1;
            const uint32_t par = (uint32_t)(it >> 1) & 1u;
            ptx::mbar_wait(ptx::smem_u32(&d_empty[s]), par ^ 1u);
            ptx::mbar_wait(ptx::smem_u32(&a_full[s]), par);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {   // elect.sync keeps the MMA operands in uniform registers
                const uint32_t a_hi = smem_base + (uint32_t)s * kABytes, a_lo = a_hi + kPlaneBytes;
                const uint32_t tmem_d = tmem_base + (uint32_t)(s * kCo);
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint32_t a_addr = term == 1 ? a_lo : a_hi;
                    const uint32_t w_addr = b_base + (term == 2 ? kBPlaneBytes : 0u);
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const uint64_t da = ptx::make_kmajor_desc<64>(a_addr + k * 32);
                        const uint64_t db = ptx::make_kmajor_desc<64>(w_addr + k * 32);
                        ptx::mma_f16_ss(tmem_d, da, db, idesc, (term | k) != 0 ? 1u : 0u);
                    }
                }
                ptx::mma_commit(ptx::smem_u32(&a_empty[s]));
                ptx::mma_commit(ptx::smem_u32(&d_full[s]));
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp - 4;
        const float inv_scale = wscale_s[1];
        const uint32_t stage_base = b_base + 2 * kBPlaneBytes;
        const int Ho = (d.H + 1) / 2, Wo = (d.W + 1) / 2, Wop = Wo + 2;
        const size_t oplane = (size_t)(Ho + 2) * Wop;
        const size_t olo = (size_t)4 * d.B * oplane;
        float tot[STATS ? 2 * kCo : 1];   // pass 1: this thread's running (sum, sum of squares) of every channel
#pragma unroll
        for (int i = 0; i < (STATS ? 2 * kCo : 1); i++) tot[i] = 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
            const int s = it & 1;
            const uint32_t par = (uint32_t)(it >> 1) & 1u;
            const int y = t / tiles_x, x = (t - y * tiles_x) * kTile + q * 32 + lane;
            ptx::mbar_wait(ptx::smem_u32(&d_full[s]), par);
            ptx::tc_fence_after();
            uint32_t u[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kCo), u);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&d_empty[s]));
            const bool inside = x < d.W;
            if (STATS) {
                if (inside) {
#pragma unroll
                    for (int co = 0; co < kCo; co++) {
                        const float a = fmaf(__uint_as_float(u[co]), inv_scale, bias_s[co]);
                        tot[(2 * co) % (STATS ? 2 * kCo : 1)] += a;
                        tot[(2 * co + 1) % (STATS ? 2 * kCo : 1)] = fmaf(a, a, tot[(2 * co + 1) % (STATS ? 2 * kCo : 1)]);
                    }
                }
            } else {
                // Stores through a per-warp staging tile: lane = pixel holds 64 B per plane, but neighbouring pixels
                // alternate between two parity phases, so direct stores touch 32 half-used sectors per instruction.
                // Staged, every instruction writes 512 contiguous bytes (eight 64-byte rows of one phase plane).
                if (d.raw_out && inside) {   // training: keep the raw convolution output (fp32 PF at the input resolution)
                    float4* ro = reinterpret_cast<float4*>(
                        d.raw_out + (((size_t)b * (d.H + 2) + (y + 1)) * (d.W + 2) + (x + 1)) * kCo);
#pragma unroll
                    for (int c0 = 0; c0 < kCo; c0 += 4)
                        ro[c0 >> 2] = make_float4(fmaf(__uint_as_float(u[c0]), inv_scale, bias_s[c0]),
                                                  fmaf(__uint_as_float(u[c0 + 1]), inv_scale, bias_s[c0 + 1]),
                                                  fmaf(__uint_as_float(u[c0 + 2]), inv_scale, bias_s[c0 + 2]),
                                                  fmaf(__uint_as_float(u[c0 + 3]), inv_scale, bias_s[c0 + 3]));
                }
                const uint32_t st = stage_base + (uint32_t)q * kStageBytes;
                const int px = lane & 1, rrow = lane >> 1;
                const uint32_t key = (uint32_t)(((rrow >> 1) & 1) | (px << 1));
                if (inside) {
#pragma unroll
                    for (int c0 = 0; c0 < kCo; c0 += 8) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float2 af = affine_s[c0 + j];
                            v[j] = fmaxf(fmaf(fmaf(__uint_as_float(u[c0 + j]), inv_scale, bias_s[c0 + j]), af.x, af.y), 0.f);
                        }
                        uint4 hi, lo;
                        split8(v, hi, lo);
                        const uint32_t off = (uint32_t)(px * 16 + rrow) * 64u + ((((uint32_t)c0 >> 3) ^ key) << 4);
                        st_shared_v4(st + off, hi);
                        st_shared_v4(st + 2048u + off, lo);
                    }
                }
                __syncwarp();
                const int xw = (t - y * tiles_x) * kTile + q * 32;   // first pixel of this warp (even)
#pragma unroll
                for (int plane_i = 0; plane_i < 2; plane_i++) {
                    if (plane_i == 1 && d.out_terms != 2) break;
#pragma unroll
                    for (int ppx = 0; ppx < 2; ppx++) {
                        const int ph = (y & 1) * 2 + ppx;
                        const size_t orow0 = ((size_t)ph * d.B + b) * oplane + (size_t)(y / 2 + 1) * Wop + (xw / 2 + 1) +
                                             (plane_i ? olo : 0);
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const int j = lane + 32 * k, rr = j >> 2, cc = j & 3;
                            if (xw + 2 * rr + ppx < d.W) {
                                const uint32_t kk = (uint32_t)(((rr >> 1) & 1) | (ppx << 1));
                                uint4 val;
                                const uint32_t addr = st + (uint32_t)plane_i * 2048u + (uint32_t)(ppx * 16 + rr) * 64u + ((((uint32_t)cc) ^ kk) << 4);
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                                             : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(addr));
                                *reinterpret_cast<uint4*>(d.out + (orow0 + rr) * kCo + cc * 8) = val;
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (STATS) {
            float v[2 * kCo];
#pragma unroll
            for (int i = 0; i < 2 * kCo; i++) v[i] = tot[i % (STATS ? 2 * kCo : 1)];
            reduce_scatter(v, lane);   // once per kernel: the per-tile cost of pass 1 is 64 FADD / FFMA per pixel
            const int ch = channel_of(lane);
            red_s[q][2 * ch] = v[0];
            red_s[q][2 * ch + 1] = v[1];
            asm volatile("bar.sync 1, 128;\n" ::: "memory");   // the four epilogue warps only
            const int i = threadIdx.x - 128;
            if (i < 2 * kCo) {
                double tsum = 0;
                for (int w = 0; w < 4; w++) tsum += (double)red_s[w][i];
                atomicAdd(d.stats + (size_t)b * kCo * 2 + i, tsum);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

const char* stem_tc_launch(const StemDesc& d, bool stats_pass, cudaStream_t stream)
{
    if (d.Cin != 1 && d.Cin != 3) return "stem: Cin must be 1 or 3";
    if (d.B <= 0 || d.B > 65535) return "stem: batch size out of range";
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = d.H * ((d.W + kTile - 1) / kTile);
    int bx = (sms * 2) / d.B;             // one resident wave: two CTAs per SM
    if (bx > tiles) bx = tiles;
    if (bx < 1) bx = 1;
    cudaError_t e;
    if (stats_pass) {
        e = cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        stem_tc_kernel<true><<<dim3(bx, d.B), kThreads, kSmemBytes, stream>>>(d);
    } else {
        e = cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        stem_tc_kernel<false><<<dim3(bx, d.B), kThreads, kSmemBytes, stream>>>(d);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cl
