// Implicit-GEMM convolution for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> fused epilogue.
//
// Replaces the cuDNN kernels behind every nn.Conv2d of the coordinate network except the 3-channel stem
// (/root/reference/networks/networks.py:191-213, 133-146, 297-306: conv3x3 s1/s2 and conv1x1).
//
// Formulation.  Activations live in HBM as fp16 "padded-flat" matrices [rows][C]: one row per pixel of a
// zero-bordered (H+2) x (W+2) image, images back to back.  In that layout the input of filter tap
// (kh, kw) is the same matrix shifted by a constant number of rows, so a convolution is a sum over taps
// of plain GEMMs  D[m, n] += A[m + shift(tap), k] * W[tap][n, k]  and every operand tile is one 2-D TMA
// box (out-of-range rows are zero-filled by the TMA unit).  Stride-2 layers read an input that was
// written as four parity phases at the output resolution, which again makes every tap a pure row shift.
// Rows that correspond to border pixels produce garbage that is never stored.
//
// Precision.  fp32 parity with the reference (1e-3 relative on the coordinate map) needs more than one
// fp16/TF32 pass (measured 1.07e-3), so by default each product is evaluated as three fp16
// MMAs  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  with fp32 accumulation in TMEM ("fp16x3", ~2^-22 relative).
// The two correction products are 2^-11 of the result, so they only need a few bits: in "fp16+fp8" mode
// (nterms == 2) they are issued as e4m3 MMAs (kind::f8f6f4, twice the fp16 rate) on power-of-two scaled
// operands into a second TMEM accumulator and folded in by the epilogue: 2/3 of the tensor work of fp16x3
// at 2e-5 relative error on the coordinate map (measured).
//
// L2 traffic.  A 128 x 256 tile needs 96 KB of operands per 64-wide K block, two thirds of it weights; at full
// tensor rate that is more than the L2 can deliver to 148 SMs (measured: 12-13 TB/s).  CTAs are therefore
// launched as clusters of `cluster` (2 or 4): the CTAs of a cluster work on consecutive pixel tiles of the same
// output-channel tile, each loads 1/cluster of the weight tile and TMA-multicasts it to its peers.  A stage is
// released to the producers only when every CTA of the cluster has consumed it (tcgen05.commit multicast).
//
// Warp roles (256 threads, one CTA per SM, persistent over output tiles):
//   warp 0   TMA producer          warp 1   tcgen05.mma issuer       warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld -> scale + bias -> swizzled smem staging -> TMA store (fp32) + GroupNorm
//             partial sums (shuffle reduction, fp64 atomics)
#include <cuda.h>

#include <cstdlib>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "conv.h"
#include "ptx_sm100.cuh"

namespace cl {

namespace {

constexpr int kBlockM = 128;
constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kCorrShift = 14;   // conv.h kCorrScale = 2^-14
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kStageChunkBytes = 32 * 32 * 4;   // one epilogue staging chunk: 32 rows x 32 fp32
constexpr uint32_t kEpilogueStagingBytes = 4 * 2 * kStageChunkBytes;

// wait / busy cycles by site (debug builds only, CL_DEBUG_TRAP=1): read with cl_debug_counters, tools/dbg_conv_waits.py
#ifdef CL_DEBUG_TRAP
__device__ unsigned long long g_conv_dbg[16];   // wait cycles by site (debug builds only): see cl_debug_counters
#define CL_DBG_T0() const long long _t0 = clock64()
#define CL_DBG_ADD(i) do { if (lane == 0) atomicAdd(&g_conv_dbg[i], (unsigned long long)(clock64() - _t0)); } while (0)
#define CL_DBG_MARK(name) const long long name = clock64()
#define CL_DBG_SINCE(i, name) do { if (lane == 0) atomicAdd(&g_conv_dbg[i], (unsigned long long)(clock64() - name)); } while (0)
#else
#define CL_DBG_T0() do {} while (0)
#define CL_DBG_ADD(i) do {} while (0)
#define CL_DBG_MARK(name) do {} while (0)
#define CL_DBG_SINCE(i, name) do {} while (0)
#endif

// Sums NV per-lane values across the warp with a recursive-halving butterfly (NV - 1 + log2(32 / NV)
// shuffles per value set).  On return v[0] of lane l holds the total of value index
// scatter_index<NV>(l); lanes sharing that index hold copies.
template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NV], int lane)
{
    int cur = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (cur > 1) {
            const int half = cur >> 1;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; i++) {
                if (i < half) {
                    const float send = upper ? v[i] : v[i + half];
                    const float keep = upper ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            cur = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
}

template <int NV>
__device__ __forceinline__ int scatter_index(int lane)
{
    int idx = 0, cur = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (cur > 1) {
            cur >>= 1;
            if (lane & off) idx += cur;
        }
    }
    return idx;
}

template <int NV>
__device__ __forceinline__ bool scatter_owner(int lane)
{
    // lanes whose low bits (those reduced by plain butterfly steps) are zero publish the value
    int cur = NV, mask = 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (cur > 1) cur >>= 1;
        else mask |= off;
    }
    return (lane & mask) == 0;
}

// Which of the warp's 32 accumulator rows are interior pixels and whether they all belong to one image: the same for every
// chunk of a tile, so the three warp collectives are paid once per tile.
struct StatsRows {
    unsigned vmask;
    int ref_image;
    bool uniform;
};
__device__ __forceinline__ StatsRows stats_rows(bool valid, int image)
{
    StatsRows r;
    r.vmask = __ballot_sync(0xffffffffu, valid);
    r.ref_image = r.vmask ? __shfl_sync(0xffffffffu, image, __ffs(r.vmask) - 1) : 0;
    r.uniform = __all_sync(0xffffffffu, !valid || image == r.ref_image);
    return r;
}

// GroupNorm partial sums of one 32-column chunk held by the warp (one row per lane).
// GC = channels per group; the chunk covers 32 / GC whole groups.
template <int GC, int NC = 32>
__device__ __forceinline__ void stats_chunk(const float (&f)[NC], bool valid, int image, int lane, double* stats,
                                            int groups, int first_group, const StatsRows* rows = nullptr)
{
    constexpr int NG = NC / GC, NV = 2 * NG;
    float v[NV];
#pragma unroll
    for (int g = 0; g < NG; g++) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int j = 0; j < GC; j++) {
            const float x = valid ? f[g * GC + j] : 0.f;
            s += x;
            ss += x * x;
        }
        v[2 * g] = s;
        v[2 * g + 1] = ss;
    }
    const StatsRows rr = rows ? *rows : stats_rows(valid, image);
    if (rr.vmask == 0) return;
    const int ref_image = rr.ref_image;
    if (rr.uniform) {
        warp_reduce_scatter<NV>(v, lane);
        if (scatter_owner<NV>(lane)) {
            const int idx = scatter_index<NV>(lane);
            atomicAdd(stats + ((size_t)ref_image * groups + first_group + (idx >> 1)) * 2 + (idx & 1), (double)v[0]);
        }
    } else if (valid) {
        // the warp's rows straddle two images (once per image boundary): every lane publishes its own sums
#pragma unroll
        for (int i = 0; i < NV; i++)
            atomicAdd(stats + ((size_t)image * groups + first_group + (i >> 1)) * 2 + (i & 1), (double)v[i]);
    }
}

// Epilogue of one 128 x BN tile for one warp (32 accumulator rows): TMEM -> registers -> (+ corrections) x scale
// + bias -> swizzled smem staging -> TMA store, plus the GroupNorm partial sums of the interior rows.
__device__ __forceinline__ void epilogue_tile(const ConvIgemmParams& p, const CUtensorMap* tmO, uint32_t taddr, int m0,
                                              int n0, int q, int lane, uint32_t stage_base, uint32_t& chunk_no, bool f8c,
                                              bool valid, int image)
{
            for (int c0 = 0; c0 < p.BN; c0 += 32, chunk_no++) {
                uint32_t u[32];
                ptx::tmem_ld_32x32(taddr + (uint32_t)c0, u);
                if (f8c) {
                    uint32_t u2[32];
                    ptx::tmem_ld_32x32(taddr + (uint32_t)(p.BN + c0), u2);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++) u[j] = __float_as_uint(fmaf(__uint_as_float(u2[j]), p.corr_scale, __uint_as_float(u[j])));
                } else {
                    ptx::tmem_ld_wait();
                }
                float f[32];
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 bb = __ldg(b4 + j);
                    f[4 * j + 0] = __uint_as_float(u[4 * j + 0]) * p.out_scale + bb.x;
                    f[4 * j + 1] = __uint_as_float(u[4 * j + 1]) * p.out_scale + bb.y;
                    f[4 * j + 2] = __uint_as_float(u[4 * j + 2]) * p.out_scale + bb.z;
                    f[4 * j + 3] = __uint_as_float(u[4 * j + 3]) * p.out_scale + bb.w;
                }
                // the staging buffer used two chunks ago must have been read by its TMA store
                if (lane == 0) ptx::tma_store_wait_read<1>();
                __syncwarp();
                const uint32_t sbuf = stage_base + (chunk_no & 1u) * kStageChunkBytes;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t dst = sbuf + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                                 "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                                 : "memory");
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    // border rows carry garbage that nothing reads; rows >= Mp are clipped by the TMA unit
                    ptx::tma_store_2d(tmO, sbuf, n0 + c0, m0 + q * 32);
                    ptx::tma_store_commit();
                }
                if (p.group_ch) {
                    const int first_group = (n0 + c0) / p.group_ch;
                    switch (p.group_ch) {
                        case 2: stats_chunk<2>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                        case 4: stats_chunk<4>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                        case 8: stats_chunk<8>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                        case 16: stats_chunk<16>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                        default: break;
                    }
                }
            }
}

template <int BK>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA8,
                  const __grid_constant__ CUtensorMap tmW8, const ConvIgemmParams p)
{
    constexpr int kSwizzle = BK * 2;   // bytes per operand row = swizzle span (128 B or 64 B)
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle atoms need 1 KB alignment
    const int nA = p.nterms == 1 ? 1 : 2;   // 16-bit-equivalent operand tiles per stage (fp16+fp8: hi + two half-size fp8 tiles)
    const bool f8c = p.nterms == 2;
    // cluster geometry: rank r of cluster c handles pixel tile (super_m * cluster + r) of super-tile list entry c, c + n, ...
    const int cs = p.cluster;
    const uint32_t crank = cs > 1 ? ptx::cluster_ctarank() : 0u;
    const int cluster_id = cs > 1 ? (int)ptx::cluster_id_x() : (int)blockIdx.x;
    const int num_clusters = cs > 1 ? (int)ptx::cluster_count_x() : (int)gridDim.x;
    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
    const int num_tiles = p.super_m * p.tiles_n;   // super-tiles: `cluster` consecutive pixel tiles x one channel tile
    const int kblocks = p.num_taps * p.kblocks_per_tap;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
        ptx::prefetch_tensormap(&tmO);
        if (f8c) {
            ptx::prefetch_tensormap(&tmA8);
            ptx::prefetch_tensormap(&tmW8);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; s++) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), (uint32_t)cs);   // every CTA of the cluster releases the stage
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 4);   // one arrival per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_s), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (cs > 1) ptx::cluster_sync();   // peers' barriers are initialised before anything is multicast to them
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)nA * (p.a_bytes + p.w_bytes);
            const int w_slice_rows = p.BN / cs;                      // rows of the weight tile this CTA fetches
            const uint32_t w_slice_off = crank * (p.w_bytes / (uint32_t)cs);
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = ((tile / p.tiles_n) * cs + (int)crank) * kBlockM;
                const int n0 = (tile % p.tiles_n) * p.BN;
                for (int tap = 0; tap < p.num_taps; tap++) {
                    const int a_row = p.tap_a_row[tap] + m0;
                    const int w_row = tap * p.w_tap_rows + n0;
                    for (int kb = 0; kb < p.kblocks_per_tap; kb++) {
                        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                        const uint32_t bar = ptx::smem_u32(&full_bar[stage]);
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                        ptx::mbar_expect_tx(bar, tx_bytes);
                        ptx::tma_load_2d(sa, &tmA, bar, kb * BK, a_row);
                        const int wr = w_row + (int)crank * w_slice_rows;   // this CTA's slice of the weight tile
                        if (cs == 1) {
                            ptx::tma_load_2d(sw, &tmW, bar, kb * BK, wr);
                        } else {
                            ptx::tma_load_2d_multicast(sw + w_slice_off, &tmW, bar, kb * BK, wr, cmask);
                        }
                        if (f8c) {
                            // e4m3 planes: [0] = hi8, [1] = lo8; tiles are half the bytes of the fp16 ones
                            ptx::tma_load_2d(sa + p.a_bytes, &tmA8, bar, kb * BK, p.a8_lo_rows + a_row);                 // a_lo8
                            ptx::tma_load_2d(sa + p.a_bytes + p.a_bytes / 2, &tmA8, bar, kb * BK, a_row);               // a_hi8
                            const uint32_t w8hi = sw + p.w_bytes, w8lo = w8hi + p.w_bytes / 2;
                            if (cs == 1) {
                                ptx::tma_load_2d(w8hi, &tmW8, bar, kb * BK, wr);                 // w_hi8
                                ptx::tma_load_2d(w8lo, &tmW8, bar, kb * BK, p.w_lo_rows + wr);   // w_lo8
                            } else {
                                ptx::tma_load_2d_multicast(w8hi + w_slice_off / 2, &tmW8, bar, kb * BK, wr, cmask);
                                ptx::tma_load_2d_multicast(w8lo + w_slice_off / 2, &tmW8, bar, kb * BK, p.w_lo_rows + wr, cmask);
                            }
                        } else if (nA == 2) {
                            ptx::tma_load_2d(sa + p.a_bytes, &tmA, bar, kb * BK, p.a_lo_rows + a_row);
                            if (cs == 1) {
                                ptx::tma_load_2d(sw + p.w_bytes, &tmW, bar, kb * BK, p.w_lo_rows + wr);
                            } else {
                                ptx::tma_load_2d_multicast(sw + p.w_bytes + w_slice_off, &tmW, bar, kb * BK, p.w_lo_rows + wr, cmask);
                            }
                        }
                        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = ptx::make_idesc_f16(kBlockM, p.BN);
        int stage = 0, local = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, local++) {
            const int as = local % p.accum_stages;
            const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
            ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aphase ^ 1u);   // epilogue has drained this accumulator
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.BN);
            for (int kbi = 0; kbi < kblocks; kbi++) {
                ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {   // elect.sync: the compiler keeps the descriptors in uniform registers
                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                    if (f8c) {
                        // D1 += a_hi * w_hi (fp16);  D2 += a_lo8 * w_hi8 + a_hi8 * w_lo8 (e4m3, K = 32 per MMA)
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) {
                            const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(sa + k * 32);
                            const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(sw + k * 32);
                            ptx::mma_f16_ss(tmem_d, da, db, idesc, (kbi | k) != 0 ? 1u : 0u);
                        }
                        const uint32_t a8lo = sa + p.a_bytes, a8hi = a8lo + p.a_bytes / 2;
                        const uint32_t w8hi = sw + p.w_bytes, w8lo = w8hi + p.w_bytes / 2;
#pragma unroll
                        for (int k = 0; k < BK / 32; k++) {
                            const uint64_t da = ptx::make_kmajor_desc<kSwizzle / 2>(a8lo + k * 32);
                            const uint64_t db = ptx::make_kmajor_desc<kSwizzle / 2>(w8hi + k * 32);
                            ptx::mma_f8_ss(tmem_d + (uint32_t)p.BN, da, db, idesc, (kbi | k) != 0 ? 1u : 0u);
                        }
#pragma unroll
                        for (int k = 0; k < BK / 32; k++) {
                            const uint64_t da = ptx::make_kmajor_desc<kSwizzle / 2>(a8hi + k * 32);
                            const uint64_t db = ptx::make_kmajor_desc<kSwizzle / 2>(w8lo + k * 32);
                            ptx::mma_f8_ss(tmem_d + (uint32_t)p.BN, da, db, idesc, 1u);
                        }
                    } else {
                    // term 0: a_hi * w_hi, term 1: a_lo * w_hi, term 2: a_hi * w_lo
                    for (int term = 0; term < p.nterms; term++) {
                        const uint32_t a_addr = sa + (term == 1 ? p.a_bytes : 0u);
                        const uint32_t w_addr = sw + (term == 2 ? p.w_bytes : 0u);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) {
                            const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a_addr + k * 32);
                            const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w_addr + k * 32);
                            ptx::mma_f16_ss(tmem_d, da, db, idesc, (kbi | term | k) != 0 ? 1u : 0u);
                        }
                    }
                    }
                    // frees the smem stage -- in every CTA of the cluster, whose multicast loads also fill it
                    if (cs == 1) ptx::mma_commit(ptx::smem_u32(&empty_bar[stage]));
                    else ptx::mma_commit_multicast(ptx::smem_u32(&empty_bar[stage]), cmask);
                    if (kbi == kblocks - 1) ptx::mma_commit(ptx::smem_u32(&tfull_bar[as]));   // accumulator ready
                }
                __syncwarp();
                if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;   // TMEM lane quarter this warp may read
        const int plane = p.Hp * p.Wp;
        // per-warp staging: two 32 x 32 fp32 chunks (4 KB each, 128-byte rows, SWIZZLE_128B) feeding TMA stores
        const uint32_t stage_base = smem_base + (uint32_t)p.num_stages * p.stage_bytes + (uint32_t)q * 2u * kStageChunkBytes;
        uint32_t chunk_no = 0;
        int local = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, local++) {
            const int as = local % p.accum_stages;
            const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
            const int m0 = ((tile / p.tiles_n) * cs + (int)crank) * kBlockM;
            const int n0 = (tile % p.tiles_n) * p.BN;
            const int m = m0 + q * 32 + lane;
            int image = 0;
            bool valid = false;
            if (m < p.Mp) {
                image = m / plane;
                const int r = m - image * plane;
                const int y = r / p.Wp, x = r - y * p.Wp;
                valid = y >= 1 && y <= p.Hp - 2 && x >= 1 && x <= p.Wp - 2;
            }
            ptx::mbar_wait(ptx::smem_u32(&tfull_bar[as]), aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
            epilogue_tile(p, &tmO, taddr, m0, n0, q, lane, stage_base, chunk_no, f8c, valid, image);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
        }
        if (lane == 0) ptx::tma_store_wait<0>();   // all output tiles are in global memory before the CTA retires
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (cs > 1) ptx::cluster_sync();   // no CTA retires while a peer may still multicast into it or signal its barriers
    if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---- dynamic tile scheduling -------------------------------------------------------------------------------
// With p.tile_counter set, the leader CTA's warp 2 (idle after the TMEM allocation) fetches tile indices from a global
// counter and publishes them through an eight-slot ring that exists in both CTAs of the pair: it writes the index into
// its own and (st.shared::cluster) the peer's ring, then arrives on `sched_full[slot]` in both.  Every consumer role
// (producer warp(s), MMA warp, the four epilogue warps of each CTA) reads the slot and arrives on the LEADER's
// `sched_empty[slot]`.  -1 terminates.  A cluster whose CTAs start late -- SMs still held by another stream's blocks --
// takes fewer tiles instead of delaying the whole launch by its statically assigned share.
// Ring depth: the consumers of one cluster spread over at most ~6 tiles (the scale loader of the fp4 kernel prefetches up to
// two short tiles ahead of the MMA warp, the epilogue trails it by one); a consumer that blocks on a slot the scheduler
// cannot publish yet would stall the pipeline it is part of.
constexpr int kRing = 8;

struct TileFeed {
    // static mode
    int next_static, stride;
    // dynamic mode
    uint32_t full_bar0, empty_bar0, ring0;   // shared-memory addresses of slot 0
    int slot;
    uint32_t phase;
    bool dynamic, leader;
};

__device__ __forceinline__ int feed_next(TileFeed& f, int num_tiles)
{
    if (!f.dynamic) {
        const int t = f.next_static;
        f.next_static += f.stride;
        return t < num_tiles ? t : -1;
    }
    ptx::mbar_wait_cluster(f.full_bar0 + 8u * (uint32_t)f.slot, f.phase, 10);
    int t;
    asm volatile("ld.volatile.shared.s32 %0, [%1];\n" : "=r"(t) : "r"(f.ring0 + 4u * (uint32_t)f.slot) : "memory");
    if (f.leader) ptx::mbar_arrive(f.empty_bar0 + 8u * (uint32_t)f.slot);
    else ptx::mbar_arrive_remote(f.empty_bar0 + 8u * (uint32_t)f.slot, 0u);
    if (++f.slot == kRing) { f.slot = 0; f.phase ^= 1u; }
    return t;
}

// ---- fused GroupNorm epilogue -------------------------------------------------------------------------------
// 32-row slices [r0, r0 + 32) of the padded-flat matrix that intersect image b
__device__ __forceinline__ int slices_of_image(int b, int plane) { return ((b + 1) * plane - 1) / 32 - (b * plane) / 32 + 1; }

__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Phase 1 of a tile for one epilogue warp: GroupNorm partial sums of the warp's 32 accumulator rows (nothing is stored),
// then the slice is counted as published for every image it touches.
__device__ __forceinline__ void fused_phase1(const ConvIgemmParams& p, uint32_t taddr, int n0, int lane, bool valid, int image,
                                             int r0, int tn)
{
    if (p.group_ch) {
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
            uint32_t u[32];
            ptx::tmem_ld_32x32(taddr + (uint32_t)c0, u);
            ptx::tmem_ld_wait();
            float f[32];
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c0);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 bb = __ldg(b4 + j);
                f[4 * j + 0] = __uint_as_float(u[4 * j + 0]) * p.out_scale + bb.x;
                f[4 * j + 1] = __uint_as_float(u[4 * j + 1]) * p.out_scale + bb.y;
                f[4 * j + 2] = __uint_as_float(u[4 * j + 2]) * p.out_scale + bb.z;
                f[4 * j + 3] = __uint_as_float(u[4 * j + 3]) * p.out_scale + bb.w;
            }
            const int first_group = (n0 + c0) / p.group_ch;
            switch (p.group_ch) {
                case 2: stats_chunk<2>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                case 4: stats_chunk<4>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                case 8: stats_chunk<8>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                case 16: stats_chunk<16>(f, valid, image, lane, p.stats, p.groups, first_group); break;
                default: break;
            }
        }
        __threadfence();   // this lane's statistics atomics are visible before the slice is counted
        __syncwarp();
        if (lane == 0 && r0 < p.Mp) {
            const int plane = p.Hp * p.Wp;
            const int b0 = r0 / plane;
            const int last = (r0 + 31 < p.Mp ? r0 + 31 : p.Mp - 1) / plane;
            for (int b = b0; b <= last; b++) atomicAdd(p.unit_done + tn * p.B + b, 1);
        }
    }
}

// true once every slice touching the image(s) of this warp's rows has published its statistics for channel tile tn
__device__ __forceinline__ bool fused_unit_complete(const ConvIgemmParams& p, int r0, int tn, int lane)
{
    if (!p.group_ch || r0 >= p.Mp) return true;
    int ok = 1;
    if (lane == 0) {
        const int plane = p.Hp * p.Wp;
        const int b0 = r0 / plane;
        const int last = (r0 + 31 < p.Mp ? r0 + 31 : p.Mp - 1) / plane;
        for (int b = b0; b <= last; b++)
            if (ld_acquire(p.unit_done + tn * p.B + b) < slices_of_image(b, plane)) ok = 0;
    }
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

// Phase 2: normalise the warp's 32 rows straight from tensor memory, ReLU, residual merge, split into the consumer's
// operand planes (fp16 hi / lo, e4m3 hi / lo) and store them with TMA.  Border rows are written as zeros: the planes
// are the zero-bordered input of the next convolution.
__device__ __forceinline__ void fused_phase2(const ConvIgemmParams& p, uint32_t taddr, int n0, int lane, bool valid, int image,
                                             int r0, int m, float2* tab)
{
    // ---- (mean, 1/sigma) of the tile's groups for the (at most two) images of this slice
    const int plane = p.Hp * p.Wp;
    const int img0 = r0 < p.Mp ? r0 / plane : 0;
    if (p.group_ch) {
        const int ng = p.BN / p.group_ch;           // groups of this channel tile (<= 32 ... 128)
        const double count = (double)p.group_ch * p.H * p.W;
        for (int which = 0; which < 2; which++) {
            const int b = img0 + which;
            for (int g = lane; g < ng; g += 32) {   // ng <= 32 (checked by the launcher)
                float2 mr = make_float2(0.f, 1.f);
                if (b < p.B) {
                    const double* st = p.stats + ((size_t)b * p.groups + n0 / p.group_ch + g) * 2;
                    const double s = __ldcg(st), ss = __ldcg(st + 1);
                    const double mu = s / count;
                    double var = ss / count - mu * mu;
                    var = var > 0 ? var : 0;
                    mr = make_float2((float)mu, (float)(1.0 / sqrt(var + (double)p.eps)));
                }
                tab[which * 32 + g] = mr;
            }
        }
        __syncwarp();
    }
    const int which = (valid && image != img0) ? 1 : 0;
    const int gshift = p.group_ch ? 31 - __clz(p.group_ch) : 0;
    const bool add_res = p.res != nullptr && valid;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
        const int c = n0 + c0;
        // residual stream first: its L2 latency hides behind the TMEM load and the normalisation
        uint4 rh[4], rl[4];
        if (add_res) {
            const uint4* ph = reinterpret_cast<const uint4*>(p.res + (size_t)m * p.Cout + c);
            const uint4* pl = reinterpret_cast<const uint4*>(p.res + ((size_t)m + (size_t)p.res_lo_rows) * p.Cout + c);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                rh[j] = __ldg(ph + j);
                rl[j] = p.res_lo_rows > 0 ? __ldg(pl + j) : make_uint4(0, 0, 0, 0);
            }
        }
        uint32_t u[32];
        ptx::tmem_ld_32x32(taddr + (uint32_t)c0, u);
        ptx::tmem_ld_wait();
        float v[32];
        {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 bb = __ldg(b4 + j);
                v[4 * j + 0] = __uint_as_float(u[4 * j + 0]) * p.out_scale + bb.x;
                v[4 * j + 1] = __uint_as_float(u[4 * j + 1]) * p.out_scale + bb.y;
                v[4 * j + 2] = __uint_as_float(u[4 * j + 2]) * p.out_scale + bb.z;
                v[4 * j + 3] = __uint_as_float(u[4 * j + 3]) * p.out_scale + bb.w;
            }
        }
        if (p.group_ch) {
            const float4* g4 = reinterpret_cast<const float4*>(p.gamma + c);
            const float4* e4 = reinterpret_cast<const float4*>(p.beta + c);
            const float2* tb = tab + which * 32;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 ga = __ldg(g4 + j), be = __ldg(e4 + j);
                const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 mr = tb[(c0 + 4 * j + k) >> gshift];
                    v[4 * j + k] = (v[4 * j + k] - mr.x) * (mr.y * gg[k]) + bb[k];
                }
            }
        }
        if (p.relu_inner) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        }
        if (add_res) {
            // res_hi + res_lo first, then the sum is added: the rounding order of gn_apply_kernel
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const __half2* hh = reinterpret_cast<const __half2*>(&rh[j]);
                const __half2* ll = reinterpret_cast<const __half2*>(&rl[j]);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 a = __half22float2(hh[k]), bq = __half22float2(ll[k]);
                    v[8 * j + 2 * k] += a.x + bq.x;
                    v[8 * j + 2 * k + 1] += a.y + bq.y;
                }
            }
        }
        if (p.relu_outer) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        }
        if (valid) {   // border rows of the planes stay zero (they are never written by anyone)
        // Straight from registers to the planes: every lane owns 32 consecutive channels of its pixel row (64 contiguous
        // bytes per fp16 plane, 32 per e4m3 plane).  No shared-memory staging, nothing to wait for between chunks.
        __half* o_hi = p.out16 + (size_t)m * p.Cout + c;
        __half* o_lo = o_hi + (size_t)p.Mp * p.Cout;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            __align__(16) __half h[8];
            __align__(16) __half l[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                h[k] = __float2half_rn(v[8 * j + k]);
                l[k] = __float2half_rn(v[8 * j + k] - __half2float(h[k]));
            }
            *reinterpret_cast<uint4*>(o_hi + 8 * j) = *reinterpret_cast<const uint4*>(h);
            if (p.out_terms == 2) *reinterpret_cast<uint4*>(o_lo + 8 * j) = *reinterpret_cast<const uint4*>(l);
        }
        if (p.has_out8) {
            uint8_t* o_h8 = p.out8 + (size_t)m * p.Cout + c;
            uint8_t* o_l8 = o_h8 + (size_t)p.Mp * p.Cout;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                __align__(16) __nv_fp8x2_storage_t h8[8];
                __align__(16) __nv_fp8x2_storage_t l8[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float x0 = v[16 * j + 2 * k], x1 = v[16 * j + 2 * k + 1];
                    const float a0 = __half2float(__float2half_rn(x0)), a1 = __half2float(__float2half_rn(x1));
                    h8[k] = __nv_cvt_float2_to_fp8x2(make_float2(a0 * kAct8HiScale, a1 * kAct8HiScale), __NV_SATFINITE, __NV_E4M3);
                    l8[k] = __nv_cvt_float2_to_fp8x2(make_float2((x0 - a0) * kAct8LoScale, (x1 - a1) * kAct8LoScale), __NV_SATFINITE, __NV_E4M3);
                }
                *reinterpret_cast<uint4*>(o_h8 + 16 * j) = *reinterpret_cast<const uint4*>(h8);
                *reinterpret_cast<uint4*>(o_l8 + 16 * j) = *reinterpret_cast<const uint4*>(l8);
            }
        }
        }
        __syncwarp();   // reconverge before the next warp-collective tcgen05.ld
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2).  In the single-CTA kernel every 128 x 256 x 16 MMA reads 12 KB of
// operands from shared memory in 128 cycles while TMA writes the next stage: more than the 128 B/clk an SM's
// shared memory delivers, so the tensor pipe starves (measured: 1.75 ms where the MMA count needs 1.2 ms).
// Here two CTAs form one 256 x 256 tile: each keeps its own 128 pixel rows and only HALF of the weight tile,
// the MMA unit of the pair reads the other half from the peer.  Per CTA and stage: half the weight bytes from
// L2, 8 KB instead of 12 KB of smem reads per MMA, three pipeline stages instead of two.
// fp16 + fp8 mode: a tile first accumulates BOTH e4m3 correction products over all k-blocks (they carry 2^14),
// then the fp16 main product; the first main MMA folds the corrections with scale-input-d (D = A*B + D * 2^-14).
// One 256-column accumulator per tile instead of two, so the accumulator is double-buffered and the epilogue of a
// tile overlaps the MMAs of the next one.  A stage holds 2 * BK channels of the e4m3 planes (128-byte rows, 4 boxes)
// or of the fp16 planes (4 boxes): fewer, wider TMA rows than interleaving both in every stage.
// Protocol (leader = even CTA of the pair): both CTAs' producers load with the 2-SM TMA form that signals the
// LEADER's full barrier; the leader's MMA warp issues for the pair and releases stages / publishes accumulators
// with tcgen05.commit multicast to both CTAs; both epilogues signal the leader's accumulator-empty barrier.
//
// FUSED (GroupNorm in the epilogue, see ConvIgemmDesc::fuse): the second accumulator doubles as the waiting room of a
// finished tile.  An epilogue warp first sums the statistics of its 32 rows (phase 1) and counts its slice as published;
// the accumulator is drained (phase 2) as soon as every slice of the image(s) it touches -- across the whole grid --
// has been published, which with the dynamic scheduler (consecutive tile indices are handed out within microseconds)
// is typically long before the next tile's MMAs finish.  Deadlock freedom: a warp only blocks on phase 2 of tile i
// AFTER it has published phase 1 of tile i + 1, and the launcher admits fused mode only when all tiles of an image
// fit into the tiles in flight (num_clusters), so the oldest unpublished image always completes.
template <int BK, bool FUSED>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA8,
                       const __grid_constant__ CUtensorMap tmW8, const __grid_constant__ ConvOutMaps outMaps,
                       const ConvIgemmParams p)
{
    constexpr int kSwizzle = BK * 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ __align__(8) uint64_t sched_full[kRing];
    __shared__ __align__(8) uint64_t sched_empty[kRing];
    __shared__ int tile_ring[kRing];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nA = p.nterms == 1 ? 1 : 2;
    const bool f8c = p.nterms == 2;
    const uint32_t crank = ptx::cluster_ctarank();
    const bool leader = crank == 0;
    const int cluster_id = (int)ptx::cluster_id_x();
    const int num_clusters = (int)ptx::cluster_count_x();
    const int num_tiles = p.super_m * p.tiles_n;
    const int kblocks = p.num_taps * p.kblocks_per_tap;
    const uint32_t w_half = p.w_bytes / 2;   // this CTA's half of the weight tile (BN / 2 rows)
    const bool dynamic = p.tile_counter != nullptr;
    const bool split = p.kblocks_per_tap == 1 && !f8c;   // two producer threads per CTA (see below)

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
        if (FUSED) {
            ptx::prefetch_tensormap(&outMaps.hi);
        } else {
            ptx::prefetch_tensormap(&tmO);
        }
        if (f8c) {
            ptx::prefetch_tensormap(&tmA8);
            ptx::prefetch_tensormap(&tmW8);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; s++) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);    // leader's producer arms it; both CTAs' TMA complete_tx
            ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);   // the leader's MMA commit, multicast to both CTAs
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);   // MMA commit, multicast to both CTAs
            ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 8);  // leader only: 4 epilogue warps of each CTA
        }
        const uint32_t consumers = 2u * ((split ? 2u : 1u) + 4u) + 1u;   // producers + epilogue warps of both CTAs, MMA warp
        for (int s = 0; s < kRing; s++) {
            ptx::mbar_init(ptx::smem_u32(&sched_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&sched_empty[s]), consumers);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc_pair(ptx::smem_u32(&tmem_base_s), kTmemCols);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    TileFeed feed;
    feed.next_static = cluster_id;
    feed.stride = num_clusters;
    feed.full_bar0 = ptx::smem_u32(&sched_full[0]);
    feed.empty_bar0 = ptx::smem_u32(&sched_empty[0]);
    feed.ring0 = ptx::smem_u32(&tile_ring[0]);
    feed.slot = 0;
    feed.phase = 0;
    feed.dynamic = dynamic;
    feed.leader = leader;

    if (warp == 2) {
        // ------------------------------------------------------------------ tile scheduler (leader CTA, one thread)
        if (dynamic && leader && lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (int local = 0;; local++) {
                // A tile is only claimed once the pair has an accumulator for it (the tile two claims ago is drained): a
                // cluster never sits on tiles it cannot start, so whatever is claimed anywhere in the grid gets accumulated
                // and published without waiting for anyone else -- the progress argument of the fused epilogue.
                if (FUSED) {
                    const int as = local % p.accum_stages;
                    ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), ((uint32_t)(local / p.accum_stages) & 1u) ^ 1u);
                }
                ptx::mbar_wait(ptx::smem_u32(&sched_empty[slot]), phase ^ 1u);   // every consumer has read this slot's last content
                int t = atomicAdd(p.tile_counter, 1);
                if (t >= num_tiles) t = -1;
                const uint32_t ring = ptx::smem_u32(&tile_ring[slot]);
                asm volatile("st.volatile.shared.s32 [%0], %1;\n" ::"r"(ring), "r"(t) : "memory");
                ptx::st_shared_remote_u32(ring, 1u, (uint32_t)t);
                const uint32_t fb = ptx::smem_u32(&sched_full[slot]);
                ptx::mbar_arrive_release_cluster(fb);
                ptx::mbar_arrive_remote(fb, 1u);
                if (t < 0) break;
                if (++slot == kRing) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 0 || warp == 3) {
        // ------------------------------------------------------------------ TMA producers (both CTAs)
        // Layers with one k-block per tap (conv2, conv3) use two producer threads per CTA: warp 0 loads the activation
        // boxes (and arms the barrier), warp 3 the weight boxes.  One thread needs ~250 instructions per stage for four
        // loads; with short stages that is on the critical path (conv2: producer and MMA issuer are both
        // instruction bound, ncu: tensor pipe 18 %; 1.06 -> 0.98 ms stand-alone).
        // Only for layers with a single k-block per tap: with longer K loops the MMAs hide the producer anyway and the
        // second thread costs ~1 % (measured on the 512-channel layers).
        const bool load_a = warp == 0, load_w = split ? warp == 3 : warp == 0;
        if (lane == 0 && (load_a || load_w)) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_pair = 2u * (uint32_t)nA * (p.a_bytes + w_half);
            const int w_rows = p.BN / 2;
            // fp16 + fp8: a stage covers 2 * BK input channels -- pass 0 streams the e4m3 planes (128-byte rows, one box
            // per plane), pass 1 the fp16 planes (two boxes per operand); the stage size is the same in both passes
            const int kstep = f8c ? 2 : 1;
            for (int tile = feed_next(feed, num_tiles); tile >= 0; tile = feed_next(feed, num_tiles)) {
                const int m0 = ((tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
                const int n0 = (tile % p.tiles_n) * p.BN;
                for (int pass = f8c ? 0 : 1; pass < 2; pass++) {
                    for (int tap = 0; tap < p.num_taps; tap++) {
                        const int a_row = p.tap_a_row[tap] + m0;
                        const int w_row = tap * p.w_tap_rows + n0 + (int)crank * w_rows;
                        for (int kb = 0; kb < p.kblocks_per_tap; kb += kstep) {
                            { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u); if (leader && warp == 0) CL_DBG_ADD(2); }
                            const uint32_t bar = ptx::smem_u32(&full_bar[stage]);   // resolved to the leader's copy by the load
                            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                            const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                            if (leader && load_a) ptx::mbar_expect_tx(bar, tx_pair);
                            if (pass == 0) {
                                ptx::tma_load_2d_pair(sa, &tmA8, bar, kb * BK, p.a8_lo_rows + a_row);           // a_lo8
                                ptx::tma_load_2d_pair(sa + p.a_bytes, &tmA8, bar, kb * BK, a_row);              // a_hi8
                                ptx::tma_load_2d_pair(sw, &tmW8, bar, kb * BK, w_row);                          // w_hi8
                                ptx::tma_load_2d_pair(sw + w_half, &tmW8, bar, kb * BK, p.w_lo_rows + w_row);   // w_lo8
                            } else if (f8c) {
                                ptx::tma_load_2d_pair(sa, &tmA, bar, kb * BK, a_row);
                                ptx::tma_load_2d_pair(sa + p.a_bytes, &tmA, bar, (kb + 1) * BK, a_row);
                                ptx::tma_load_2d_pair(sw, &tmW, bar, kb * BK, w_row);
                                ptx::tma_load_2d_pair(sw + w_half, &tmW, bar, (kb + 1) * BK, w_row);
                            } else {
                                if (load_a) {
                                    ptx::tma_load_2d_pair(sa, &tmA, bar, kb * BK, a_row);
                                    if (nA == 2) ptx::tma_load_2d_pair(sa + p.a_bytes, &tmA, bar, kb * BK, p.a_lo_rows + a_row);
                                }
                                if (load_w) {
                                    ptx::tma_load_2d_pair(sw, &tmW, bar, kb * BK, w_row);
                                    if (nA == 2) ptx::tma_load_2d_pair(sw + w_half, &tmW, bar, kb * BK, p.w_lo_rows + w_row);
                                }
                            }
                            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            const uint32_t idesc = ptx::make_idesc_f16(2 * kBlockM, p.BN);
            int stage = 0, local = 0;
            uint32_t phase = 0;
            CL_DBG_MARK(t_mma_loop);
            for (;; local++) {
                int tile = 0;
                if (lane == 0) tile = feed_next(feed, num_tiles);
                tile = __shfl_sync(0xffffffffu, tile, 0);
                if (tile < 0) {
                    CL_DBG_SINCE(6, t_mma_loop);
#ifdef CL_DEBUG_TRAP
                    if (lane == 0) atomicAdd(&g_conv_dbg[7], (unsigned long long)local);
#endif
                    break;
                }
                const int as = local % p.accum_stages;
                const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
                { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aphase ^ 1u); CL_DBG_ADD(0); }
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.BN);
                const int nstages_tile = f8c ? kblocks / 2 : kblocks;
                if (f8c) {
                    // pass 0: both correction products (scaled by 2^14) of 2 * BK input channels per stage
                    for (int kbi = 0; kbi < nstages_tile; kbi++) {
                        ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {   // elect.sync: the compiler keeps the descriptors in uniform registers
                            const uint32_t a8lo = smem_base + (uint32_t)stage * p.stage_bytes, a8hi = a8lo + p.a_bytes;
                            const uint32_t w8hi = a8lo + 2u * p.a_bytes, w8lo = w8hi + w_half;
#pragma unroll
                            for (int k = 0; k < BK / 16; k++) {
                                const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a8lo + k * 32);
                                const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w8hi + k * 32);
                                ptx::mma_f8_ss_pair(tmem_d, da, db, idesc, (kbi | k) != 0 ? 1u : 0u);
                            }
#pragma unroll
                            for (int k = 0; k < BK / 16; k++) {
                                const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a8hi + k * 32);
                                const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w8lo + k * 32);
                                ptx::mma_f8_ss_pair(tmem_d, da, db, idesc, 1u);
                            }
                            ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 0x3);
                        }
                        __syncwarp();
                        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                    }
                }
                for (int kbi = 0; kbi < nstages_tile; kbi++) {
                    { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase); CL_DBG_ADD(1); }
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {   // elect.sync: the compiler keeps the descriptors in uniform registers
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                        if (f8c) {
                            // pass 1: the fp16 product; its first MMA rescales the corrections: D = A*B + D * 2^-14
#pragma unroll
                            for (int half = 0; half < 2; half++) {
#pragma unroll
                                for (int k = 0; k < BK / 16; k++) {
                                    const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(sa + half * p.a_bytes + k * 32);
                                    const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(sw + half * w_half + k * 32);
                                    if (kbi == 0 && half == 0 && k == 0) ptx::mma_f16_ss_pair_scaled_d<kCorrShift>(tmem_d, da, db, idesc);
                                    else ptx::mma_f16_ss_pair(tmem_d, da, db, idesc, 1u);
                                }
                            }
                        } else {
                            for (int term = 0; term < p.nterms; term++) {
                                const uint32_t a_addr = sa + (term == 1 ? p.a_bytes : 0u);
                                const uint32_t w_addr = sw + (term == 2 ? w_half : 0u);
#pragma unroll
                                for (int k = 0; k < BK / 16; k++) {
                                    const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a_addr + k * 32);
                                    const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w_addr + k * 32);
                                    ptx::mma_f16_ss_pair(tmem_d, da, db, idesc, (kbi | term | k) != 0 ? 1u : 0u);
                                }
                            }
                        }
                        ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 0x3);            // frees the stage in both CTAs
                        if (kbi == nstages_tile - 1) ptx::mma_commit_pair(ptx::smem_u32(&tfull_bar[as]), 0x3);   // accumulators ready
                    }
                    __syncwarp();
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int plane = p.Hp * p.Wp;
        const uint32_t stage_base = smem_base + (uint32_t)p.num_stages * p.stage_bytes + (uint32_t)q * 2u * kStageChunkBytes;
        uint32_t chunk_no = 0;
        // a finished tile whose accumulator still waits for the statistics of its image(s) (FUSED only)
        int pend_tile = -1, pend_as = 0;
        auto tile_rows = [&](int tile, int& m0, int& n0, int& m, bool& valid, int& image) {
            m0 = ((tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
            n0 = (tile % p.tiles_n) * p.BN;
            m = m0 + q * 32 + lane;
            image = 0;
            valid = false;
            if (m < p.Mp) {
                image = m / plane;
                const int r = m - image * plane;
                const int y = r / p.Wp, x = r - y * p.Wp;
                valid = y >= 1 && y <= p.Hp - 2 && x >= 1 && x <= p.Wp - 2;
            }
        };
        auto release = [&](int as) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
                else ptx::mbar_arrive_remote(ptx::smem_u32(&tempty_bar[as]), 0u);
            }
        };
        auto drain_pending = [&]() {   // phase 2 of the waiting tile; its statistics are complete
            int m0, n0, m, image;
            bool valid;
            tile_rows(pend_tile, m0, n0, m, valid, image);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pend_as * p.BN);
            ptx::tc_fence_after();
            // (mean, 1/sigma) table of this warp: the 2 KB of its staging area that phase 2 leaves unused
            float2* tab = reinterpret_cast<float2*>(smem_raw + (stage_base + 6144u - ptx::smem_u32(smem_raw)));
            fused_phase2(p, taddr, n0, lane, valid, image, m0 + q * 32, m, tab);
            release(pend_as);
            pend_tile = -1;
        };
        for (int local = 0;; local++) {
            int tile = 0;
            if (lane == 0) tile = feed_next(feed, num_tiles);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile < 0) break;
            const int as = local % p.accum_stages;
            const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
            int m0, n0, m, image;
            bool valid;
            tile_rows(tile, m0, n0, m, valid, image);
            const uint32_t tfull = ptx::smem_u32(&tfull_bar[as]);
            if (FUSED) {
                // while this tile's MMAs run: drain the waiting tile as soon as its image(s) are complete
                while (pend_tile >= 0) {
                    int ready = lane == 0 ? (int)ptx::mbar_test(tfull, aphase) : 0;
                    ready = __shfl_sync(0xffffffffu, ready, 0);
                    if (ready) break;
                    const int pm0 = ((pend_tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
                    if (fused_unit_complete(p, pm0 + q * 32, pend_tile % p.tiles_n, lane)) drain_pending();
                    else __nanosleep(200);
                }
            }
            { CL_DBG_T0(); ptx::mbar_wait(tfull, aphase); if (leader && warp == 4) CL_DBG_ADD(4); }
            CL_DBG_MARK(t_epi);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
            if (FUSED) {
                fused_phase1(p, taddr, n0, lane, valid, image, m0 + q * 32, tile % p.tiles_n);
                if (pend_tile >= 0) {
                    // the previous tile is still waiting: its images only need slices that were handed out before this
                    // tile, all of which publish without waiting for anyone -- blocking here cannot deadlock
                    const int pm0 = ((pend_tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
                    const long long t0 = clock64();
                    while (!fused_unit_complete(p, pm0 + q * 32, pend_tile % p.tiles_n, lane)) {
                        __nanosleep(200);
                        if (clock64() - t0 > 4000000000ll) __trap();
                    }
                    drain_pending();
                }
                pend_tile = tile;
                pend_as = as;
            } else {
                epilogue_tile(p, &tmO, taddr, m0, n0, q, lane, stage_base, chunk_no, false, valid, image);   // corrections already folded
                release(as);
            }
            if (leader && warp == 4) CL_DBG_SINCE(5, t_epi);
        }
        if (FUSED && pend_tile >= 0) {
            const int pm0 = ((pend_tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
            const long long t0 = clock64();
            while (!fused_unit_complete(p, pm0 + q * 32, pend_tile % p.tiles_n, lane)) {
                __nanosleep(200);
                if (clock64() - t0 > 4000000000ll) __trap();
            }
            drain_pending();
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel with block-scaled FP4 corrections (nterms == 4, "fp16 + fp4").  Same tile, pipeline and protocol as
// conv_igemm_pair_kernel in fp16 + fp8 mode, except that pass 0 issues the two correction products as kind::mxf4 MMAs
// (e2m1 operands, K = 64 per instruction, four times the fp16 rate: 1.5 instead of 2 fp16-MMA equivalents per product;
// tools/emulate_precision.py: 1.5e-4 relative on the coordinate map, against 1.07e-3 of a single fp16 pass).
//   * A stage of pass 0 holds 256 input channels of the four e2m1 planes (128-byte rows, same 64 KB as an fp16 stage).
//   * Scales: one ue8m0 exponent per (row, 256 channels) for each plane -- the emulation shows no difference between
//     32- and 512-channel blocks, the corrections only need ~3 bits -- stored twice so that both 32-element blocks of an
//     instruction read the same value.  The activation word of a pixel row is (lo, lo, hi, hi), the weight word of an
//     output channel (hi, hi, lo, lo): product a_lo * w_hi uses scale bytes 0-1, a_hi * w_lo bytes 2-3 of the same words.
//     Warp 3 of each CTA moves them global -> shared (the tap shift makes the pixel-row order of a tile unknown to any
//     fixed tiling, so the 32-row interleave tcgen05.cp expects is done here); the MMA thread copies them into tensor
//     memory (tcgen05.cp, in order with the MMAs) right before the stage's instructions.
//   * The corrections carry their true scale (no 2^14, no scale-input-d): pass 1 simply keeps accumulating.
//   * Tensor memory: the scale columns do not fit next to two 256-column accumulators, so the second accumulator starts
//     at column 224 and the scales live in columns 480..511.  The epilogue drains the 32 shared columns first and then
//     lets the MMA warp start the next tile; the rest of the drain overlaps that tile's MMAs as before.
constexpr uint32_t kSfStageBytes = 4096;   // shared memory reserved per pipeline stage for the scale ring below
constexpr int kSfSlots = 8;                // the scales have their own ring (only pass 0 uses them): a slot = 512 B activation
constexpr uint32_t kSfSlotBytes = 1536;    // scales + 2 x 512 B weight scales; kSfSlots * kSfSlotBytes <= 3 * kSfStageBytes
constexpr uint32_t kAcc1Col = 224;
constexpr uint32_t kSfCol = 480;
constexpr int kFp4Threads = 384;           // eight epilogue warps: two per tensor-memory lane quarter, alternating 16-column chunks
constexpr uint32_t kFp4ChunkBytes = 32 * 16 * 4;               // one staging chunk: 32 rows x 16 fp32 (64-byte rows, SWIZZLE_64B)
constexpr uint32_t kFp4StagingBytes = 8 * kFp4ChunkBytes;      // one staging chunk per epilogue warp

__global__ void __launch_bounds__(kFp4Threads, 1)
conv_igemm_pair_fp4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                           const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA4,
                           const __grid_constant__ CUtensorMap tmW4, const __grid_constant__ CUtensorMap tmA136,
                           const __grid_constant__ CUtensorMap tmA4x136, const ConvIgemmParams p)
{
    constexpr int BK = 64;
    constexpr int kSwizzle = 128;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t sf_full[kSfSlots];
    __shared__ __align__(8) uint64_t sf_empty[kSfSlots];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t ovl_bar;
    __shared__ __align__(8) uint64_t sched_full[kRing];
    __shared__ __align__(8) uint64_t sched_empty[kRing];
    __shared__ int tile_ring[kRing];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sf_base = smem_base + (uint32_t)p.num_stages * p.stage_bytes;
    const uint32_t crank = ptx::cluster_ctarank();
    const bool leader = crank == 0;
    const int cluster_id = (int)ptx::cluster_id_x();
    const int num_clusters = (int)ptx::cluster_count_x();
    const int num_tiles = p.super_m * p.tiles_n;
    const uint32_t w_half = p.w_bytes / 2;
    const bool dynamic = p.tile_counter != nullptr;
    // stages of pass 0: 256 channels of the four e2m1 planes of one tap each, or (kw_share0) of ONE activation plane shared by
    // the three kw taps of a filter row plus the matching weight plane of the three taps (two stages per filter row and group)
    const int n4 = p.kw_share0 ? 6 * p.kgroups : p.num_taps * p.kgroups;
    // stages of pass 1: 128 channels of the fp16 planes of one tap each, or (kw_share) 64 channels of the three kw taps of a
    // filter row: ONE 136-row activation tile read at row shifts 0, 1, 2 plus the three weight tiles -- a third of the
    // activation bytes the L2 has to deliver for this pass
    constexpr uint32_t kA136Bytes = 136 * 128;
    const int n16 = p.kw_share ? 3 * p.kblocks_per_tap : p.num_taps * p.kblocks_per_tap / 2;
    static_assert(kSfSlots * kSfSlotBytes <= 3 * kSfStageBytes, "scale ring exceeds the shared memory reserved for it");

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
        ptx::prefetch_tensormap(&tmO);
        ptx::prefetch_tensormap(&tmA4);
        ptx::prefetch_tensormap(&tmW4);
        if (p.kw_share) ptx::prefetch_tensormap(&tmA136);
        if (p.kw_share0) ptx::prefetch_tensormap(&tmA4x136);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; s++) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < kSfSlots; s++) {
            ptx::mbar_init(ptx::smem_u32(&sf_full[s]), 2);     // leader only: the scale loader warp of each CTA
            ptx::mbar_init(ptx::smem_u32(&sf_empty[s]), 1);    // the leader's MMA commit, multicast to both CTAs
        }
        for (int s = 0; s < 2; s++) ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
        ptx::mbar_init(ptx::smem_u32(&ovl_bar), 16);           // leader only: 8 epilogue warps of each CTA
        const uint32_t consumers = 2u * (1u + 1u + 8u) + 1u;   // producer, scale loader and epilogue warps of both CTAs, MMA warp
        for (int s = 0; s < kRing; s++) {
            ptx::mbar_init(ptx::smem_u32(&sched_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&sched_empty[s]), consumers);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc_pair(ptx::smem_u32(&tmem_base_s), kTmemCols);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    TileFeed feed;
    feed.next_static = cluster_id;
    feed.stride = num_clusters;
    feed.full_bar0 = ptx::smem_u32(&sched_full[0]);
    feed.empty_bar0 = ptx::smem_u32(&sched_empty[0]);
    feed.ring0 = ptx::smem_u32(&tile_ring[0]);
    feed.slot = 0;
    feed.phase = 0;
    feed.dynamic = dynamic;
    feed.leader = leader;

    if (warp == 2) {
        // ------------------------------------------------------------------ tile scheduler (leader CTA, one thread)
        if (dynamic && leader && lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (;;) {
                ptx::mbar_wait(ptx::smem_u32(&sched_empty[slot]), phase ^ 1u, 1);
                int t = atomicAdd(p.tile_counter, 1);
                if (t >= num_tiles) t = -1;
                const uint32_t ring = ptx::smem_u32(&tile_ring[slot]);
                asm volatile("st.volatile.shared.s32 [%0], %1;\n" ::"r"(ring), "r"(t) : "memory");
                ptx::st_shared_remote_u32(ring, 1u, (uint32_t)t);
                const uint32_t fb = ptx::smem_u32(&sched_full[slot]);
                ptx::mbar_arrive_release_cluster(fb);
                ptx::mbar_arrive_remote(fb, 1u);
                if (t < 0) break;
                if (++slot == kRing) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const bool skip_w = p.corr_scale < 0.f;   // EXPERIMENT: no weight loads (timing only, results are garbage)
            const uint32_t tx_pair = skip_w ? 4u * p.a_bytes : 4u * (p.a_bytes + w_half);
            const int w_rows = p.BN / 2;
            for (int tile = feed_next(feed, num_tiles); tile >= 0; tile = feed_next(feed, num_tiles)) {
                const int m0 = ((tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
                const int n0 = (tile % p.tiles_n) * p.BN;
                for (int pass = 0; pass < 2; pass++) {
                    if (pass == 0 && p.kw_share0) {
                        const uint32_t tx_kw = 2u * (kA136Bytes + 3u * w_half);
                        for (int kh = 0; kh < 3; kh++) {
                            const int a_row = min(p.tap_a_row[3 * kh], p.tap_a_row[3 * kh + 2]) + m0;
                            for (int kg = 0; kg < p.kgroups; kg++) {
                                for (int half = 0; half < 2; half++) {   // a_lo4 with w_hi4, then a_hi4 with w_lo4
                                    ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u, 2);
                                    const uint32_t bar = ptx::smem_u32(&full_bar[stage]);
                                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                                    if (leader) ptx::mbar_expect_tx(bar, tx_kw);
                                    ptx::tma_load_2d_pair(sa, &tmA4x136, bar, kg * 128, (half == 0 ? p.a4_lo_rows : 0) + a_row);
#pragma unroll
                                    for (int kw = 0; kw < 3; kw++)
                                        ptx::tma_load_2d_pair(sa + kA136Bytes + (uint32_t)kw * w_half, &tmW4, bar, kg * 128,
                                                              (half == 0 ? 0 : p.w_lo_rows) + (3 * kh + kw) * p.w_tap_rows + n0 +
                                                                  (int)crank * w_rows);
                                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                                }
                            }
                        }
                        continue;
                    }
                    if (pass == 1 && p.kw_share) {
                        const uint32_t tx_kw = 2u * (kA136Bytes + 3u * w_half);
                        for (int kh = 0; kh < 3; kh++) {
                            // first of the three consecutive rows (kw = 0 for a forward filter, kw = 2 for a transposed one)
                            const int a_row = min(p.tap_a_row[3 * kh], p.tap_a_row[3 * kh + 2]) + m0;
                            for (int kb = 0; kb < p.kblocks_per_tap; kb++) {
                                ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u, 2);
                                const uint32_t bar = ptx::smem_u32(&full_bar[stage]);
                                const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                                if (leader) ptx::mbar_expect_tx(bar, tx_kw);
                                ptx::tma_load_2d_pair(sa, &tmA136, bar, kb * BK, a_row);
#pragma unroll
                                for (int kw = 0; kw < 3; kw++)
                                    ptx::tma_load_2d_pair(sa + kA136Bytes + (uint32_t)kw * w_half, &tmW, bar, kb * BK,
                                                          (3 * kh + kw) * p.w_tap_rows + n0 + (int)crank * w_rows);
                                if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                            }
                        }
                        continue;
                    }
                    for (int tap = 0; tap < p.num_taps; tap++) {
                        const int a_row = p.tap_a_row[tap] + m0;
                        const int w_row = tap * p.w_tap_rows + n0 + (int)crank * w_rows;
                        const int steps = pass == 0 ? p.kgroups : p.kblocks_per_tap / 2;
                        for (int ks = 0; ks < steps; ks++) {
                            ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u, 2);
                            const uint32_t bar = ptx::smem_u32(&full_bar[stage]);
                            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                            const uint32_t sw = sa + 2u * p.a_bytes;
                            if (leader) ptx::mbar_expect_tx(bar, tx_pair);
                            if (pass == 0) {
                                ptx::tma_load_2d_pair(sa, &tmA4, bar, ks * 128, p.a4_lo_rows + a_row);        // a_lo4
                                ptx::tma_load_2d_pair(sa + p.a_bytes, &tmA4, bar, ks * 128, a_row);            // a_hi4
                                if (!skip_w) {
                                ptx::tma_load_2d_pair(sw, &tmW4, bar, ks * 128, w_row);                        // w_hi4
                                ptx::tma_load_2d_pair(sw + w_half, &tmW4, bar, ks * 128, p.w_lo_rows + w_row); // w_lo4
                                }
                            } else {
                                const int kb = 2 * ks;
                                ptx::tma_load_2d_pair(sa, &tmA, bar, kb * BK, a_row);
                                ptx::tma_load_2d_pair(sa + p.a_bytes, &tmA, bar, (kb + 1) * BK, a_row);
                                if (!skip_w) {
                                ptx::tma_load_2d_pair(sw, &tmW, bar, kb * BK, w_row);
                                ptx::tma_load_2d_pair(sw + w_half, &tmW, bar, (kb + 1) * BK, w_row);
                                }
                            }
                            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ scale-factor loader (both CTAs, whole warp)
        // The global loads of a stage's scales run kSfAhead stages ahead of the store into shared memory (a rotating set
        // of registers): their L2 latency (~1 us) is several times the MMA time of a stage.
        constexpr int kSfAhead = 4;
        uint4 qa[kSfAhead], qb0[kSfAhead], qb1[kSfAhead];
        bool qv[kSfAhead];
        const int sf_rows = p.a4_lo_rows;
        const int nblocks = p.Cout / 128;
        int ltap = p.num_taps, lkg = 0, lm0 = 0, ln0 = 0;
        int lkh = 0, lkw = 0;   // kw_share0: slots in the order (kh, group, kw) the shared-tile stages consume them
        bool ldone = false;
        auto load_item = [&](uint4& wa, uint4& b0, uint4& b1) -> bool {
            if (ldone) return false;
            if (ltap == p.num_taps) {
                int tile = 0;
                { CL_DBG_T0(); if (lane == 0) tile = feed_next(feed, num_tiles); if (leader) CL_DBG_ADD(12); }
                tile = __shfl_sync(0xffffffffu, tile, 0);
                if (tile < 0) { ldone = true; return false; }
                lm0 = ((tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
                ln0 = (tile % p.tiles_n) * p.BN;
                ltap = 0;
                lkg = 0;
                lkh = 0;
                lkw = 0;
            }
            // lane l collects the words of rows l, l + 32, l + 64, l + 96 of the tile: one 16-byte row of the
            // 32 x 128-bit block tcgen05.cp broadcasts to the four lane quarters
            const int a_row = p.tap_a_row[ltap] + lm0;
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int r = a_row + 32 * j + lane;
                w[j] = (r >= 0 && r < sf_rows) ? __ldg(p.act_sf + (size_t)lkg * sf_rows + r) : 0u;
            }
            wa = make_uint4(w[0], w[1], w[2], w[3]);
            const uint4* wb = reinterpret_cast<const uint4*>(p.w_sf + ((size_t)(ltap * p.kgroups + lkg) * nblocks + ln0 / 128) * 128);
            b0 = __ldg(wb + lane);
            b1 = __ldg(wb + 32 + lane);
            if (p.kw_share0) {
                if (++lkw == 3) {
                    lkw = 0;
                    if (++lkg == p.kgroups) { lkg = 0; ++lkh; }
                }
                ltap = lkh == 3 ? p.num_taps : 3 * lkh + lkw;
            } else if (++lkg == p.kgroups) { lkg = 0; ++ltap; }
            return true;
        };
#pragma unroll
        for (int u = 0; u < kSfAhead; u++) qv[u] = load_item(qa[u], qb0[u], qb1[u]);
        // The scale ring is independent of the operand stages: a slot is handed over per pass-0 stage and released by the
        // MMA warp's commit, so the loader neither has to follow the stages of pass 1 nor can it be lapped by them.
        int slot = 0;
        uint32_t sphase = 0;
        for (bool more = true; more;) {
#pragma unroll
            for (int u = 0; u < kSfAhead; u++) {
                if (!qv[u]) { more = false; break; }
                { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&sf_empty[slot]), sphase ^ 1u, 3); if (leader) CL_DBG_ADD(11); }
                const uint32_t dst = sf_base + (uint32_t)slot * kSfSlotBytes + (uint32_t)lane * 16u;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(qa[u].x), "r"(qa[u].y), "r"(qa[u].z), "r"(qa[u].w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + 512u), "r"(qb0[u].x), "r"(qb0[u].y), "r"(qb0[u].z), "r"(qb0[u].w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + 1024u), "r"(qb1[u].x), "r"(qb1[u].y), "r"(qb1[u].z), "r"(qb1[u].w) : "memory");
                ptx::fence_proxy_async_smem();   // tcgen05.cp reads through the async proxy
                __syncwarp();
                if (lane == 0) {
                    if (leader) ptx::mbar_arrive_release_cluster(ptx::smem_u32(&sf_full[slot]));
                    else ptx::mbar_arrive_remote(ptx::smem_u32(&sf_full[slot]), 0u);
                }
                if (++slot == kSfSlots) { slot = 0; sphase ^= 1u; }
                qv[u] = load_item(qa[u], qb0[u], qb1[u]);   // the stage kSfAhead further on
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            const uint32_t idesc = ptx::make_idesc_f16(2 * kBlockM, p.BN);
            const uint32_t idesc4a = ptx::make_idesc_mxf4(2 * kBlockM, p.BN, 0u, 0u);   // a_lo4 * w_hi4: scale bytes 0-1
            const uint32_t idesc4b = ptx::make_idesc_mxf4(2 * kBlockM, p.BN, 2u, 2u);   // a_hi4 * w_lo4: scale bytes 2-3
            int stage = 0, local = 0;
            uint32_t phase = 0, sphase = 0;
            int slot = 0;
#ifdef CL_DEBUG_TRAP
            const long long _tk = clock64();
#endif
            for (;; local++) {
                int tile = 0;
                if (lane == 0) tile = feed_next(feed, num_tiles);
                tile = __shfl_sync(0xffffffffu, tile, 0);
                if (tile < 0) break;
                const int as = local & 1;
                if (local > 0) {
                    // the previous tile's epilogue has drained the columns both accumulators share (and, before that,
                    // everything of the tile that used this accumulator last)
                    { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&ovl_bar), (uint32_t)(local - 1) & 1u, 5); CL_DBG_ADD(0); }
                    ptx::tc_fence_after();
                }
                const uint32_t tmem_d = tmem_base + (as ? kAcc1Col : 0u);
                if (p.kw_share0) {
                    // shared-tile stages: (filter row, channel group) x (a_lo4 * w_hi4 | a_hi4 * w_lo4); the three scale slots of
                    // the row's taps are copied into tensor memory for both products and released after the second one
                    for (int j = 0; j < 3 * p.kgroups; j++) {
                        const int slot0 = slot;
                        const uint32_t sph0 = sphase;
                        for (int half = 0; half < 2; half++) {
                            { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase, 6); CL_DBG_ADD(1); }
                            if (half == 0) {
                                CL_DBG_T0();
                                for (int kw = 0; kw < 3; kw++) {
                                    const int sl = slot0 + kw;
                                    ptx::mbar_wait(ptx::smem_u32(&sf_full[sl % kSfSlots]), sl >= kSfSlots ? sph0 ^ 1u : sph0, 7);
                                }
                                CL_DBG_ADD(2);
                            }
                            ptx::tc_fence_after();
                            if (ptx::elect_one()) {
                                const uint32_t a4 = smem_base + (uint32_t)stage * p.stage_bytes, w4 = a4 + kA136Bytes;
#pragma unroll
                                for (int kw = 0; kw < 3; kw++) {
                                    const uint32_t sfs = sf_base + (uint32_t)((slot0 + kw) % kSfSlots) * kSfSlotBytes;
                                    const uint32_t sfa = tmem_base + kSfCol + (uint32_t)((half * 3 + kw) & 1) * 16u, sfb = sfa + 4u;
                                    ptx::tmem_cp_sf_pair(sfa, ptx::make_sf_desc(sfs));
                                    ptx::tmem_cp_sf_pair(sfb, ptx::make_sf_desc(sfs + 512u));
                                    ptx::tmem_cp_sf_pair(sfb + 4u, ptx::make_sf_desc(sfs + 1024u));
                                    const uint32_t shift = (uint32_t)(p.kw_share > 0 ? kw : 2 - kw) * 128u;
#pragma unroll
                                    for (int k = 0; k < 4; k++) {
                                        const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a4 + shift + k * 32);
                                        const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w4 + kw * w_half + k * 32);
                                        ptx::mma_mxf4_ss_pair(tmem_d, da, db, half ? idesc4b : idesc4a, (j | half | kw | k) != 0 ? 1u : 0u, sfa, sfb);
                                    }
                                }
                                ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 0x3);
                                if (half == 1)
                                    for (int kw = 0; kw < 3; kw++)
                                        ptx::mma_commit_pair(ptx::smem_u32(&sf_empty[(slot0 + kw) % kSfSlots]), 0x3);
                            }
                            __syncwarp();
                            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                        }
                        slot = slot0 + 3;
                        if (slot >= kSfSlots) { slot -= kSfSlots; sphase ^= 1u; }
                    }
                }
                for (int i = 0; i < (p.kw_share0 ? 0 : n4); i++) {
                    { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase, 6); CL_DBG_ADD(1); }
                    // (CTA-scope wait, as for the operand barrier: the peer's scales stay in the peer's shared memory and are
                    // read there by its own tensor core; a cluster-scope acquire here costs ~350 cycles per stage)
                    { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&sf_full[slot]), sphase, 7); CL_DBG_ADD(2); }
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {   // elect.sync: the compiler keeps the descriptors in uniform registers
                        const uint32_t a4lo = smem_base + (uint32_t)stage * p.stage_bytes, a4hi = a4lo + p.a_bytes;
                        const uint32_t w4hi = a4lo + 2u * p.a_bytes, w4lo = w4hi + w_half;
                        const uint32_t sfs = sf_base + (uint32_t)slot * kSfSlotBytes;
                        const uint32_t sfa = tmem_base + kSfCol + (uint32_t)(i & 1) * 16u, sfb = sfa + 4u;
                        ptx::tmem_cp_sf_pair(sfa, ptx::make_sf_desc(sfs));
                        ptx::tmem_cp_sf_pair(sfb, ptx::make_sf_desc(sfs + 512u));
                        ptx::tmem_cp_sf_pair(sfb + 4u, ptx::make_sf_desc(sfs + 1024u));
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a4lo + k * 32);
                            const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w4hi + k * 32);
                            ptx::mma_mxf4_ss_pair(tmem_d, da, db, idesc4a, (i | k) != 0 ? 1u : 0u, sfa, sfb);
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(a4hi + k * 32);
                            const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(w4lo + k * 32);
                            ptx::mma_mxf4_ss_pair(tmem_d, da, db, idesc4b, 1u, sfa, sfb);
                        }
                        ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 0x3);
                        ptx::mma_commit_pair(ptx::smem_u32(&sf_empty[slot]), 0x3);   // the scale slot has been copied and used
                    }
                    __syncwarp();
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                    if (++slot == kSfSlots) { slot = 0; sphase ^= 1u; }
                }
                for (int i = 0; i < n16; i++) {
                    { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase, 8); CL_DBG_ADD(3); }
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {   // elect.sync: the compiler keeps the descriptors in uniform registers
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        if (p.kw_share) {
                            const uint32_t sw = sa + kA136Bytes;
                            const int kw_dir = p.kw_share;   // +1: rows ascend with kw, -1: they descend (transposed filter)
#pragma unroll
                            for (int kw = 0; kw < 3; kw++) {
                                const uint32_t shift = (uint32_t)(kw_dir > 0 ? kw : 2 - kw) * 128u;
#pragma unroll
                                for (int k = 0; k < BK / 16; k++) {
                                    // row shift of the shared tile: plain start-address offset (ptx_sm100.cuh, note on swizzles)
                                    const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(sa + shift + k * 32);
                                    const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(sw + kw * w_half + k * 32);
                                    ptx::mma_f16_ss_pair(tmem_d, da, db, idesc, 1u);
                                }
                            }
                        } else {
                        const uint32_t sw = sa + 2u * p.a_bytes;
#pragma unroll
                        for (int half = 0; half < 2; half++) {
#pragma unroll
                            for (int k = 0; k < BK / 16; k++) {
                                const uint64_t da = ptx::make_kmajor_desc<kSwizzle>(sa + half * p.a_bytes + k * 32);
                                const uint64_t db = ptx::make_kmajor_desc<kSwizzle>(sw + half * w_half + k * 32);
                                ptx::mma_f16_ss_pair(tmem_d, da, db, idesc, 1u);
                            }
                        }
                        }
                        ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 0x3);
                        if (i == n16 - 1) ptx::mma_commit_pair(ptx::smem_u32(&tfull_bar[as]), 0x3);
                    }
                    __syncwarp();
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
#ifdef CL_DEBUG_TRAP
            if (lane == 0) { atomicAdd(&g_conv_dbg[6], (unsigned long long)(clock64() - _tk)); atomicAdd(&g_conv_dbg[7], (unsigned long long)local); }
#endif
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        // Two warps per lane quarter drain alternating 16-column chunks: with short K loops (1x1 layers: six stages per
        // tile) the drain of a tile, not its MMAs, is the critical path -- TMEM load, scale + bias, staging, TMA store and
        // the GroupNorm sums of 128 x 256 values take longer than 48 MMAs.
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int plane = p.Hp * p.Wp;
        const uint32_t sbuf = sf_base + (uint32_t)p.num_stages * kSfStageBytes + (uint32_t)(warp - 4) * kFp4ChunkBytes;
        for (int local = 0;; local++) {
            int tile = 0;
            if (lane == 0) tile = feed_next(feed, num_tiles);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile < 0) break;
            const int as = local & 1;
            const uint32_t aphase = (uint32_t)(local >> 1) & 1u;
            const int m0 = ((tile / p.tiles_n) * 2 + (int)crank) * kBlockM;
            const int n0 = (tile % p.tiles_n) * p.BN;
            const int m = m0 + q * 32 + lane;
            int image = 0;
            bool valid = false;
            if (m < p.Mp) {
                image = m / plane;
                const int r = m - image * plane;
                const int y = r / p.Wp, x = r - y * p.Wp;
                valid = y >= 1 && y <= p.Hp - 2 && x >= 1 && x <= p.Wp - 2;
            }
            { CL_DBG_T0(); ptx::mbar_wait(ptx::smem_u32(&tfull_bar[as]), aphase, 9); if (warp == 4) CL_DBG_ADD(4); }
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (as ? kAcc1Col : 0u);
            const int nchunks = p.BN / 16;
            const StatsRows srows = stats_rows(valid, image);
            // drain order: the two chunks in tensor-memory columns 224..255 (the only ones the other accumulator
            // shares) go first, one per warp group
            auto chunk_col = [&](int idx) { return as ? idx * 16 : (idx < 2 ? (nchunks - 2 + idx) * 16 : (idx - 2) * 16); };
            // groups of 16 channels (the 512-channel layers): the (sum, sum of squares) of the warp's up to eight chunks wait in
            // registers for ONE reduce-scatter per tile -- the chain of dependent shuffles per chunk was the longest part of the
            // 1x1 layers' drain (810 of ~1650 cycles per chunk, debug counters)
            const bool defer16 = p.group_ch == 16;
            float gsum[16];
#pragma unroll
            for (int i = 0; i < 16; i++) gsum[i] = 0.f;
#pragma unroll 1
            for (int k = 0; k < 8; k++) {
                const int idx = grp + 2 * k;
                if (idx >= nchunks) break;
                const int c0 = chunk_col(idx);
                uint32_t u[16];
                { CL_DBG_T0(); ptx::tmem_ld_32x16(taddr + (uint32_t)c0, u);
                ptx::tmem_ld_wait(); if (warp == 4) CL_DBG_ADD(8); }
                CL_DBG_MARK(t_store);
                if (idx < 2) {
                    // this warp is done with the previous tile and with the shared columns of this one
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (leader) ptx::mbar_arrive(ptx::smem_u32(&ovl_bar));
                        else ptx::mbar_arrive_remote(ptx::smem_u32(&ovl_bar), 0u);
                    }
                }
                float f[16];
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 bb = __ldg(b4 + j);
                    f[4 * j + 0] = __uint_as_float(u[4 * j + 0]) * p.out_scale + bb.x;
                    f[4 * j + 1] = __uint_as_float(u[4 * j + 1]) * p.out_scale + bb.y;
                    f[4 * j + 2] = __uint_as_float(u[4 * j + 2]) * p.out_scale + bb.z;
                    f[4 * j + 3] = __uint_as_float(u[4 * j + 3]) * p.out_scale + bb.w;
                }
                { CL_DBG_T0(); if (lane == 0) ptx::tma_store_wait_read<0>(); if (warp == 4) CL_DBG_ADD(5); }   // single staging chunk per warp: the previous store has read it
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t dst = sbuf + (uint32_t)lane * 64u + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                                 "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                                 : "memory");
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmO, sbuf, n0 + c0, m0 + q * 32);
                    ptx::tma_store_commit();
                }
                // Measured with the debug counters (1x1 512->512, cycles per 16-column chunk and warp): tensor-memory load
                // 35, scale / bias / staging / TMA store 700, GroupNorm sums 810 -> 300 with one reduce-scatter per tile (below;
                // same-box A/B: 1x1 layers 0.212 -> 0.199 ms).  Tried and dropped: staged coalesced st.global instead of the
                // TMA store (920 cycles).
                if (warp == 4) CL_DBG_SINCE(9, t_store);
                CL_DBG_MARK(t_stats);
                if (defer16) {
                    float sm = 0.f, sq = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float xv = valid ? f[j] : 0.f;
                        sm += xv;
                        sq += xv * xv;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {   // static register indices: select instead of gsum[2 * k]
                        gsum[2 * i] = i == k ? sm : gsum[2 * i];
                        gsum[2 * i + 1] = i == k ? sq : gsum[2 * i + 1];
                    }
                } else if (p.group_ch) {
                    const int first_group = (n0 + c0) / p.group_ch;
                    switch (p.group_ch) {
                        case 2: stats_chunk<2, 16>(f, valid, image, lane, p.stats, p.groups, first_group, &srows); break;
                        case 4: stats_chunk<4, 16>(f, valid, image, lane, p.stats, p.groups, first_group, &srows); break;
                        case 8: stats_chunk<8, 16>(f, valid, image, lane, p.stats, p.groups, first_group, &srows); break;
                        case 16: stats_chunk<16, 16>(f, valid, image, lane, p.stats, p.groups, first_group, &srows); break;
                        default: break;
                    }
                }
                if (warp == 4) CL_DBG_SINCE(10, t_stats);
            }
            if (defer16 && srows.vmask != 0) {
                CL_DBG_MARK(t_flush);
                if (srows.uniform) {
                    warp_reduce_scatter<16>(gsum, lane);
                    if (scatter_owner<16>(lane)) {
                        const int vi = scatter_index<16>(lane);
                        const int idx = grp + 2 * (vi >> 1);
                        if (idx < nchunks)
                            atomicAdd(p.stats + ((size_t)srows.ref_image * p.groups + (n0 + chunk_col(idx)) / 16) * 2 + (vi & 1), (double)gsum[0]);
                    }
                } else if (valid) {
                    // the warp's rows straddle two images (once per image boundary): every lane publishes its own sums
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int idx = grp + 2 * (i >> 1);
                        if (idx < nchunks)
                            atomicAdd(p.stats + ((size_t)image * p.groups + (n0 + chunk_col(idx)) / 16) * 2 + (i & 1), (double)gsum[i]);
                    }
                }
                if (warp == 4) CL_DBG_SINCE(10, t_flush);
            }
            ptx::tc_fence_before();
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// row-major [rows][cols] matrix of fp16 (elem_bytes 2) or fp32 (4), box = box_rows x box_cols elements whose
// byte width is the swizzle span; rows out of range are zero-filled on loads and clipped on stores
bool make_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                     uint32_t box_cols, uint32_t elem_bytes, bool no_swizzle = false)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const uint32_t row_bytes = box_cols * elem_bytes;
    const CUtensorMapSwizzle sw = no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                                  : (row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                     : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B));
    const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                   : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    return fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

const char* conv_igemm_prepare(const ConvIgemmDesc& d, ConvIgemmPlan* plan)
{
    if (d.Cin % 32 != 0) return "conv_igemm: Cin must be a multiple of 32";
    if (d.Cout % 64 != 0) return "conv_igemm: Cout must be a multiple of 64";
    if (d.num_taps < 1 || d.num_taps > 9) return "conv_igemm: 1..9 taps";
    if (d.nterms < 1 || d.nterms > 4) return "conv_igemm: nterms must be 1 (fp16), 2 (fp16 + fp8 corrections), 3 (fp16x3) or 4 (fp16 + fp4 corrections)";
    if (d.nterms == 4 && (d.Cin % 256 != 0 || d.Cout % 256 != 0 || !d.act4 || !d.act_sf || !d.weights4 || !d.w_sf || d.fuse))
        return "conv_igemm: the fp16 + fp4 mode needs Cin % 256 == 0, Cout % 256 == 0, the e2m1 operand planes with their scales, and no fused epilogue";
    if (d.nterms == 2 && (d.Cin % 64 != 0 || !d.act8 || !d.weights8))
        return "conv_igemm: the fp16 + fp8 mode needs Cin % 64 == 0 and the e4m3 operand planes";
    if (d.group_ch != 0 && d.group_ch != 2 && d.group_ch != 4 && d.group_ch != 8 && d.group_ch != 16)
        return "conv_igemm: GroupNorm group size must be 2, 4, 8 or 16 channels";
    const int BK = d.Cin % 64 == 0 ? 64 : 32;
    const int BN = d.Cout % 256 == 0 ? 256 : (d.Cout % 128 == 0 ? 128 : 64);
    const int nA = d.nterms == 1 ? 1 : 2;

    ConvIgemmParams& p = plan->p;
    p = ConvIgemmParams{};
    p.num_taps = d.num_taps;
    for (int i = 0; i < d.num_taps; i++) p.tap_a_row[i] = d.tap_a_row[i];
    p.kblocks_per_tap = d.Cin / BK;
    p.nterms = d.nterms;
    p.a_lo_rows = (int)d.a_lo_rows;
    p.a8_lo_rows = (int)d.a8_lo_rows;
    p.corr_scale = (d.nterms == 4 && getenv("CL_EXPERIMENT_SKIP_W")) ? -1.f : d.corr_scale;
    p.w_tap_rows = d.Cout;
    p.w_lo_rows = d.num_taps * d.Cout;
    p.Mp = d.Mp; p.Cout = d.Cout; p.BN = BN;
    p.tiles_m = (d.Mp + kBlockM - 1) / kBlockM;
    p.tiles_n = d.Cout / BN;
    static int sms_cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& sms = sms_cached[dev & 63];
    if (sms == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    // d.cluster: 0 / 2 = CTA pairs (cta_group::2, default), 1 = single CTAs, 12 / 14 = single-CTA MMAs with the
    // weight tile TMA-multicast across clusters of 2 / 4 (kept for comparison)
    const bool pair = (d.cluster == 2 || (d.cluster == 0 && p.tiles_m >= 16)) && sms % 2 == 0 && BN % 32 == 0;
    int cluster = pair ? 2 : (d.cluster > 10 ? d.cluster - 10 : 1);
    while (!pair && cluster > 1 && (p.tiles_m < cluster * 8 || sms % cluster != 0 || BN % (cluster * 8) != 0)) cluster >>= 1;
    p.cluster = cluster;
    p.super_m = (p.tiles_m + cluster - 1) / cluster;
    p.Hp = d.Hp; p.Wp = d.Wp;
    p.group_ch = d.group_ch;
    p.groups = d.group_ch ? d.Cout / d.group_ch : 0;
    p.out_scale = d.out_scale;
    p.raw = d.raw; p.bias = d.bias; p.stats = d.stats;
    p.a_bytes = (uint32_t)(kBlockM * BK * 2);
    p.w_bytes = (uint32_t)(BN * BK * 2);
    // a CTA pair splits the weight tile (in fp16 + fp8 mode its stages hold the e4m3 planes OR the fp16 planes of 2 * BK channels)
    p.stage_bytes = (uint32_t)nA * (p.a_bytes + (pair ? p.w_bytes / 2 : p.w_bytes));
    const int smem_budget = 227 * 1024 - 2048 - (int)kEpilogueStagingBytes;
    p.num_stages = smem_budget / (int)p.stage_bytes;
    if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
    if (p.num_stages < 2) return "conv_igemm: tile does not fit two pipeline stages";
    // fp16 + fp8 mode keeps two accumulators (main, corrections) per tile in the single-CTA kernel; the pair kernel
    // folds the corrections with scale-input-d and needs one
    if (d.nterms == 2 && d.corr_scale != 1.0f / (float)(1 << kCorrShift)) return "conv_igemm: corr_scale must be 2^-14";
    if (d.nterms == 4 && !pair) return "conv_igemm: the fp16 + fp4 mode exists in the CTA-pair kernel only";
    p.accum_stages = (d.nterms == 2 && !pair ? 4 : 2) * BN <= (int)kTmemCols ? 2 : 1;
    plan->smem = (size_t)p.num_stages * p.stage_bytes + kEpilogueStagingBytes + 1024;
    p.kw_share = 0;
    if (d.nterms == 4) {
        // filter rows whose three taps are consecutive activation rows (3x3, stride 1): pass 1 shares one activation tile
        // between them (stage = 136-row activation tile + three weight half-tiles = 65 KB instead of 64 KB)
        // (CROSSLOC_B200_KW_SHARE=0 restores one tile per tap; same-box A/B: 3x3 512->512 0.881 -> 0.862 ms, step 19.26 -> 18.95 ms.
        //  =2 also shares the e2m1 planes of pass 0 -- identical results, but no measurable gain: 0.905 vs 0.902 ms, opt-in only)
        static const int kw_env = [] { const char* e = getenv("CROSSLOC_B200_KW_SHARE"); return e ? atoi(e) : 1; }();
        int dir = d.num_taps == 9 ? d.tap_a_row[1] - d.tap_a_row[0] : 0;   // +1 forward filter, -1 transposed (data gradient)
        bool rows3 = kw_env != 0 && (dir == 1 || dir == -1) && BK == 64 && BN == 256;
        for (int kh = 0; kh < 3 && rows3; kh++)
            rows3 = d.tap_a_row[3 * kh + 1] == d.tap_a_row[3 * kh] + dir && d.tap_a_row[3 * kh + 2] == d.tap_a_row[3 * kh] + 2 * dir;
        p.kw_share0 = 0;
        if (rows3) {
            p.kw_share0 = kw_env == 2 ? 1 : 0;
            p.kw_share = dir;
            p.stage_bytes = 136u * 128u + 3u * (p.w_bytes / 2);
        }
        const int budget4 = 227 * 1024 - 2048 - (int)kFp4StagingBytes;
        p.num_stages = budget4 / (int)(p.stage_bytes + kSfStageBytes);
        if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
        if (p.num_stages < 3) return "conv_igemm: the fp16 + fp4 tile does not fit three pipeline stages";
        plan->smem = (size_t)p.num_stages * (p.stage_bytes + kSfStageBytes) + kFp4StagingBytes + 1024;
        p.a4_lo_rows = (int)d.a4_lo_rows;
        p.kgroups = d.Cin / 256;
        p.act_sf = d.act_sf;
        p.w_sf = d.w_sf;
    }
    p.tile_counter = pair ? d.tile_counter : nullptr;
    p.fuse = 0;
    if (d.fuse) {
        const int plane = d.Hp * d.Wp;
        if (!pair || !d.tile_counter) return "conv_igemm: the fused GroupNorm epilogue needs the CTA-pair kernel and a tile counter";
        if (!d.out16 || (d.out_terms != 1 && d.out_terms != 2)) return "conv_igemm: fused epilogue without output planes";
        if (d.group_ch && (!d.stats || !d.gamma || !d.beta || !d.unit_done)) return "conv_igemm: fused epilogue without GroupNorm parameters";
        if (d.Mp % plane != 0 || plane < 32) return "conv_igemm: fused epilogue needs whole planes of at least 32 rows";
        if (p.accum_stages != 2) return "conv_igemm: fused epilogue needs two accumulators";
        if (BN / (d.group_ch ? d.group_ch : BN) > 32) return "conv_igemm: too many groups per channel tile for the fused epilogue";
        // deadlock freedom: every tile touching one image must be among the tiles in flight
        const int tiles_per_image = ((plane + 2 * kBlockM - 1) / (2 * kBlockM) + 1) * p.tiles_n;
        if (tiles_per_image > sms / 2) return "conv_igemm: image too large for the fused epilogue (tiles of one image exceed the grid)";
        p.fuse = 1;
        p.H = d.H; p.W = d.W; p.B = d.Mp / plane;
        p.relu_inner = d.relu_inner; p.relu_outer = d.relu_outer;
        p.out_terms = d.out_terms; p.has_out8 = d.out8 ? 1 : 0;
        p.gamma = d.gamma; p.beta = d.beta; p.eps = d.eps;
        p.res = d.res; p.res_lo_rows = d.res_lo_rows;
        p.unit_done = d.unit_done;
        p.out16 = d.out16;
        p.out8 = d.out8;
        const size_t plane16 = (size_t)d.Mp * d.Cout;
        if (!make_tensor_map(&plan->out.hi, d.out16, (uint64_t)d.Mp, (uint64_t)d.Cout, 32, 32, 2))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the fp16 output plane";
        plan->out.lo = plan->out.hi;
        plan->out.hi8 = plan->out.hi;
        plan->out.lo8 = plan->out.hi;
        if (d.out_terms == 2 && !make_tensor_map(&plan->out.lo, d.out16 + plane16, (uint64_t)d.Mp, (uint64_t)d.Cout, 32, 32, 2))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the fp16 lo output plane";
        if (d.out8) {
            if (!make_tensor_map(&plan->out.hi8, d.out8, (uint64_t)d.Mp, (uint64_t)d.Cout, 32, 32, 1, true) ||
                !make_tensor_map(&plan->out.lo8, d.out8 + plane16, (uint64_t)d.Mp, (uint64_t)d.Cout, 32, 32, 1, true))
                return "conv_igemm: cuTensorMapEncodeTiled failed for the e4m3 output planes";
        }
    }

    if (!make_tensor_map(&plan->tmA, d.act, (uint64_t)d.a_total_rows, (uint64_t)d.Cin, kBlockM, BK, 2))
        return "conv_igemm: cuTensorMapEncodeTiled failed for the activation matrix";
    if (!make_tensor_map(&plan->tmW, d.weights, (uint64_t)(d.nterms == 3 ? 2 : 1) * d.num_taps * d.Cout, (uint64_t)d.Cin, BN / cluster, BK, 2))
        return "conv_igemm: cuTensorMapEncodeTiled failed for the weight matrix";
    if (d.fuse) plan->tmO = plan->tmA;
    else if (!make_tensor_map(&plan->tmO, d.raw, (uint64_t)d.Mp, (uint64_t)d.Cout, 32, d.nterms == 4 ? 16 : 32, 4))
        return "conv_igemm: cuTensorMapEncodeTiled failed for the output matrix";
    plan->tmA8 = plan->tmA;
    plan->tmW8 = plan->tmW;
    if (d.nterms == 4) {
        // e2m1 planes as byte matrices: 128-byte rows = 256 channels per box
        if (!make_tensor_map(&plan->tmA8, d.act4, (uint64_t)d.a4_total_rows, (uint64_t)d.Cin / 2, kBlockM, 128, 1))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the e2m1 activation matrix";
        if (!make_tensor_map(&plan->tmW8, d.weights4, (uint64_t)2 * d.num_taps * d.Cout, (uint64_t)d.Cin / 2, BN / 2, 128, 1))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the e2m1 weight matrix";
        plan->out.hi = plan->tmA;
        plan->out.lo = plan->tmA8;
        if (p.kw_share && !make_tensor_map(&plan->out.hi, d.act, (uint64_t)d.a_total_rows, (uint64_t)d.Cin, 136, BK, 2))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the 136-row activation box";
        if (p.kw_share0 && !make_tensor_map(&plan->out.lo, d.act4, (uint64_t)d.a4_total_rows, (uint64_t)d.Cin / 2, 136, 128, 1))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the 136-row e2m1 activation box";
    }
    if (d.nterms == 2) {
        // the pair kernel streams the e4m3 planes in 128-byte rows (2 * BK channels per box)
        const int bk8 = pair ? 2 * BK : BK;
        if (pair && (d.Cin % 128 != 0)) return "conv_igemm: the fp16 + fp8 mode of the CTA-pair kernel needs Cin % 128 == 0";
        if (!make_tensor_map(&plan->tmA8, d.act8, (uint64_t)d.a8_total_rows, (uint64_t)d.Cin, kBlockM, bk8, 1))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the e4m3 activation matrix";
        if (!make_tensor_map(&plan->tmW8, d.weights8, (uint64_t)2 * d.num_taps * d.Cout, (uint64_t)d.Cin, BN / cluster, bk8, 1))
            return "conv_igemm: cuTensorMapEncodeTiled failed for the e4m3 weight matrix";
    }

    const int num_super = p.super_m * p.tiles_n;
    int clusters = sms / cluster;
    if (clusters > num_super) clusters = num_super;
    plan->grid = clusters * cluster;
    plan->cluster = cluster;
    plan->variant = d.nterms == 4 ? 8 : (p.fuse ? 4 : 0) + (pair ? 2 : 0) + (BK == 64 ? 1 : 0);
    return nullptr;
}

const char* conv_igemm_run(const ConvIgemmPlan& plan, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(plan.variant == 8 ? kFp4Threads : kThreads);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;

    // the opt-in shared-memory size is a per-function, per-device attribute: raise it once to the maximum any plan uses
    static bool attr_set[64][16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& done = attr_set[dev & 63][plan.variant & 15];
    const int max_smem = 227 * 1024 - 1024;   // dynamic part only: the kernels also hold ~300 bytes of static shared memory
    cudaError_t e = cudaSuccess;
#define CL_LAUNCH(KERNEL, ...)                                                                              \
    do {                                                                                                    \
        if (!done) e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); \
        if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, KERNEL, __VA_ARGS__);                            \
    } while (0)
    switch (plan.variant) {
        case 8: CL_LAUNCH(conv_igemm_pair_fp4_kernel, plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.out.hi, plan.out.lo, plan.p); break;
        case 7: CL_LAUNCH((conv_igemm_pair_kernel<64, true>), plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.out, plan.p); break;
        case 6: CL_LAUNCH((conv_igemm_pair_kernel<32, true>), plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.out, plan.p); break;
        case 3: CL_LAUNCH((conv_igemm_pair_kernel<64, false>), plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.out, plan.p); break;
        case 2: CL_LAUNCH((conv_igemm_pair_kernel<32, false>), plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.out, plan.p); break;
        case 1: CL_LAUNCH(conv_igemm_kernel<64>, plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.p); break;
        default: CL_LAUNCH(conv_igemm_kernel<32>, plan.tmA, plan.tmW, plan.tmO, plan.tmA8, plan.tmW8, plan.p); break;
    }
#undef CL_LAUNCH
    if (e != cudaSuccess) return cudaGetErrorString(e);
    done = true;
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

#ifdef CL_DEBUG_TRAP
void conv_debug_counters(unsigned long long* out16, bool reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_conv_dbg, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long zero[16] = {};
        cudaMemcpyToSymbol(g_conv_dbg, zero, sizeof(zero));
    }
}
#endif

const char* conv_igemm_launch(const ConvIgemmDesc& d, cudaStream_t stream)
{
    ConvIgemmPlan plan;
    if (const char* err = conv_igemm_prepare(d, &plan)) return err;
    return conv_igemm_run(plan, stream);
}

}  // namespace cl
