// Layout kernels of the training path: NCHW fp32 tensors (what autograd hands over) <-> the fp16 hi/lo operand
// layouts of the tensor-core kernels, and filter packing.  All HBM-bound copies / transposes.
//
//   nchw_to_pf   activations or output gradients -> padded-flat pixel-major [term][phase][B*(H+2)*(W+2)][C]
//                (operand of cl_conv_igemm: forward and data gradient)
//   pf_to_nchw   raw fp32 padded-flat result -> NCHW (optionally into one parity phase of a 2x larger tensor:
//                the per-phase results of a stride-2 data gradient), x scale, + bias
//   pack_filter  OIHW fp32 filter -> fp16 hi/lo [2][tap][N][K] for a tap subset, optionally transposed
//                (data gradient) and zero-padded in N, x a power-of-two scale read from device memory
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv.h"

namespace cl {

namespace {

constexpr int kTile = 32;

__device__ __forceinline__ void split2(float v, __half& hi, __half& lo)
{
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// block (32, 8); grid (ceil(W / 32), H, B * C / 32)
__global__ void __launch_bounds__(256) nchw_to_pf_kernel(NchwToPfDesc d)
{
    __shared__ float tile[kTile][kTile + 1];
    const int ctiles = d.C / kTile;
    const int b = blockIdx.z / ctiles, c0 = (blockIdx.z % ctiles) * kTile;
    const int y = blockIdx.y, x0 = blockIdx.x * kTile;
    const float sc = d.scale ? __ldg(d.scale) : 1.f;
    for (int cc = threadIdx.y; cc < kTile; cc += 8) {
        const int x = x0 + threadIdx.x;
        tile[cc][threadIdx.x] = x < d.W ? __ldg(d.x + (((size_t)b * d.C + c0 + cc) * d.H + y) * d.W + x) * sc : 0.f;
    }
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int p = tid >> 3, q = tid & 7;   // pixel within the tile, 4-channel chunk
    const int x = x0 + p;
    if (x >= d.W) return;
    size_t row, lo_rows;
    if (d.phases == 1) {
        const int Wp = d.W + 2, plane = (d.H + 2) * Wp;
        row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
        lo_rows = (size_t)d.B * plane;
    } else {
        const int Ho = (d.H + 1) / 2, Wo = (d.W + 1) / 2, Wop = Wo + 2, oplane = (Ho + 2) * Wop;
        const int ph = (y & 1) * 2 + (x & 1);
        row = ((size_t)ph * d.B + b) * oplane + (size_t)(y / 2 + 1) * Wop + (x / 2 + 1);
        lo_rows = (size_t)4 * d.B * oplane;
    }
    __align__(8) __half h[4];
    __align__(8) __half l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) split2(tile[q * 4 + j][p], h[j], l[j]);
    *reinterpret_cast<uint2*>(d.out + row * d.C + c0 + q * 4) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(d.out + (row + lo_rows) * d.C + c0 + q * 4) = *reinterpret_cast<const uint2*>(l);
}

// block (32, 8); grid (ceil(W / 32), H, B * ceil(C / 32)) over the PF geometry (H, W) of the raw matrix
__global__ void __launch_bounds__(256) pf_to_nchw_kernel(PfToNchwDesc d)
{
    __shared__ float tile[kTile][kTile + 1];
    const int ctiles = (d.C + kTile - 1) / kTile;
    const int b = blockIdx.z / ctiles, c0 = (blockIdx.z % ctiles) * kTile;
    const int y = blockIdx.y, x0 = blockIdx.x * kTile;
    const int Wp = d.W + 2, plane = (d.H + 2) * Wp;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    {
        const int p = tid >> 3, q = tid & 7;
        const int x = x0 + p;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < d.W && c0 + q * 4 < d.Craw) {
            const size_t row = (size_t)b * plane + (size_t)(y + 1) * Wp + (x + 1);
            v = __ldg(reinterpret_cast<const float4*>(d.raw + row * d.Craw + c0 + q * 4));
        }
        tile[q * 4 + 0][p] = v.x; tile[q * 4 + 1][p] = v.y; tile[q * 4 + 2][p] = v.z; tile[q * 4 + 3][p] = v.w;
    }
    __syncthreads();
    const float sc = d.scale ? __ldg(d.scale) : 1.f;
    const int yo = y * d.step + d.off_y;
    if (yo >= d.Hout) return;
    for (int cc = threadIdx.y; cc < kTile; cc += 8) {
        const int c = c0 + cc, xo = (x0 + threadIdx.x) * d.step + d.off_x;
        if (c < d.C && x0 + threadIdx.x < d.W && xo < d.Wout) {
            float v = tile[cc][threadIdx.x] * sc;
            if (d.bias) v += __ldg(d.bias + c);
            d.out[(((size_t)b * d.C + c) * d.Hout + yo) * d.Wout + xo] = v;
        }
    }
}

// amax -> power-of-two scale, one launch: every block folds its maximum into scratch[0] (float bits of a
// non-negative value order like unsigned integers), the last block to finish writes {2^k, 2^-k} with
// k = floor(log2(target / amax)).  scratch = two zeroed 32-bit words.
__global__ void __launch_bounds__(256) pow2_scale_kernel(const float* x, size_t n, float target, unsigned* scratch, float* out)
{
    float m = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = fmaxf(m, wm[w]);
        atomicMax(scratch, __float_as_uint(m));
        __threadfence();
        if (atomicAdd(scratch + 1, 1u) == gridDim.x - 1) {
            const float amax = fmaxf(__uint_as_float(atomicMax(scratch, 0u)), 1e-30f);
            const float s = exp2f(floorf(log2f(target / amax)));
            out[0] = s;
            out[1] = 1.f / s;
        }
    }
}

// one thread per output element of one term
__global__ void __launch_bounds__(256) pack_filter_kernel(PackFilterDesc d)
{
    const size_t per_term = (size_t)d.num_taps * d.N * d.K;
    const float sc = __ldg(d.scale);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_term; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % d.K);
        const int n = (int)((i / d.K) % d.N);
        const int t = (int)(i / ((size_t)d.K * d.N));
        const int kk = d.ksize * d.ksize;
        const int tap = d.tap_kh[t] * d.ksize + d.tap_kw[t];
        float v = 0.f;
        if (!d.transpose) {
            if (n < d.Cout) v = __ldg(d.w + ((size_t)n * d.Cin + k) * kk + tap);          // [N = co][K = ci]
        } else {
            if (n < d.Cin) v = __ldg(d.w + ((size_t)k * d.Cin + n) * kk + tap);           // [N = ci][K = co]
        }
        __half hi, lo;
        split2(v * sc, hi, lo);
        d.out[i] = hi;
        d.out[i + per_term] = lo;
    }
}

const char* last_error()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

unsigned blocks_for(size_t n)
{
    size_t b = (n + 255) / 256;
    return (unsigned)(b > 148 * 32 ? 148 * 32 : (b ? b : 1));
}

}  // namespace

const char* nchw_to_pf_launch(const NchwToPfDesc& d, cudaStream_t stream)
{
    if (d.C % 32 != 0) return "nchw_to_pf: C must be a multiple of 32";
    if (d.phases != 1 && d.phases != 4) return "nchw_to_pf: phases must be 1 or 4";
    if ((long long)d.B * (d.C / 32) > 65535 || d.H > 65535) return "nchw_to_pf: grid too large";
    if (d.B == 0) return nullptr;
    nchw_to_pf_kernel<<<dim3((d.W + 31) / 32, d.H, d.B * (d.C / 32)), dim3(32, 8), 0, stream>>>(d);
    return last_error();
}

const char* pf_to_nchw_launch(const PfToNchwDesc& d, cudaStream_t stream)
{
    if (d.Craw % 4 != 0) return "pf_to_nchw: raw channel count must be a multiple of 4";
    if ((long long)d.B * ((d.C + 31) / 32) > 65535 || d.H > 65535) return "pf_to_nchw: grid too large";
    if (d.B == 0) return nullptr;
    pf_to_nchw_kernel<<<dim3((d.W + 31) / 32, d.H, d.B * ((d.C + 31) / 32)), dim3(32, 8), 0, stream>>>(d);
    return last_error();
}

const char* pow2_scale_launch(const float* x, size_t n, float target, unsigned* scratch, float* out, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    size_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    pow2_scale_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, target, scratch, out);
    return last_error();
}

const char* pack_filter_launch(const PackFilterDesc& d, cudaStream_t stream)
{
    if (d.num_taps < 1 || d.num_taps > 9) return "pack_filter: 1..9 taps";
    pack_filter_kernel<<<blocks_for((size_t)d.num_taps * d.N * d.K), 256, 0, stream>>>(d);
    return last_error();
}

}  // namespace cl
