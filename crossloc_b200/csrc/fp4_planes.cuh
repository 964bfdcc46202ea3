// Block-scaled e2m1 operand planes of the fp16 + fp4 convolution mode (shared by the GroupNorm apply pass of the forward
// and the GroupNorm backward pass that feeds the data gradient).
#pragma once
#include <cuda_fp4.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace cl {
namespace {

// Block-scaled e2m1 planes of the fp16 + fp4 convolution mode.  The 32 lanes of a warp hold the 256 channels of one
// pixel (8 each): one power-of-two scale per plane and warp, chosen so that the block maximum lands in (3, 6] -- the
// top binade of e2m1 -- and stored as a ue8m0 byte.  Returns the scale word (lo, lo, hi, hi) in every lane.
__device__ __forceinline__ uint32_t ue8m0_for(float amax)
{
    // smallest e with amax <= 6 * 2^e:  amax = m * 2^ex, m in [1, 2)  ->  e = ex - 2 (m <= 1.5) or ex - 1
    const uint32_t bits = __float_as_uint(amax);
    int sf = (int)(bits >> 23) - 2 + ((bits & 0x7FFFFFu) > 0x400000u ? 1 : 0);
    return amax > 0.f ? (uint32_t)(sf < 1 ? 1 : sf) : 127u;
}

__device__ __forceinline__ uint32_t fp4_pack8(const float (&x)[8], float inv)
{
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        out |= (uint32_t)__nv_cvt_float2_to_fp4x2(make_float2(x[2 * j] * inv, x[2 * j + 1] * inv), __NV_E2M1, cudaRoundNearest) << (8 * j);
    return out;
}

__device__ __forceinline__ uint32_t fp4_store8(uint8_t* hi_ptr, uint8_t* lo_ptr, const float (&v)[8], bool exact_lo = false)
{
    float hi[8], lo[8];
    float mh = 0.f, ml = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        hi[j] = __half2float(__float2half_rn(v[j]));
        lo[j] = v[j] - hi[j];
        mh = fmaxf(mh, fabsf(hi[j]));
        ml = fmaxf(ml, fabsf(lo[j]));
    }
    // block maximum over the warp: non-negative floats order like their bit patterns (one redux.sync instead of a butterfly)
    const uint32_t mh_bits = __reduce_max_sync(0xffffffffu, __float_as_uint(mh));
    const uint32_t sh = ue8m0_for(__uint_as_float(mh_bits));
    uint32_t sl;
    if (exact_lo) {
        sl = ue8m0_for(__uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ml))));
    } else {
        // |x - fp16(x)| <= half an fp16 ulp of the block maximum = 2^(ex - 11) (2^-25 below the fp16 normal range):
        // the scale 2^(ex - 13) puts that bound at 4.0, inside e2m1's range, without a second reduction
        const uint32_t e = mh_bits >> 23;
        sl = (e < 113u ? 113u : e) - 13u;
    }
    *reinterpret_cast<uint32_t*>(hi_ptr) = fp4_pack8(hi, __uint_as_float((254u - sh) << 23));
    *reinterpret_cast<uint32_t*>(lo_ptr) = fp4_pack8(lo, __uint_as_float((254u - sl) << 23));
    return sl | (sl << 8) | (sh << 16) | (sh << 24);
}

}  // namespace
}  // namespace cl
