// Weight gradient on padded-flat (PF) operands for sm_100a: both GEMM operands are read straight from the
// [pixel rows][channels] matrices the forward pass and the GroupNorm backward already produce, as MN-major
// tcgen05 operands -- no channel-major copies, no column-shifted duplicates.
//
//   dW[tap][co][ci] = sum over PF rows r of  dY[r][co] * X[phase(tap)][r + shift(tap)][ci]
//
// (the autograd weight gradient of every nn.Conv2d on the path, /root/reference/train_single_task.py:298).
// K = pixel row index; a TMA box of 64 rows x 64 channels lands in shared memory as eight-row 128-byte swizzle
// atoms, which is the canonical MN-major SWIZZLE_128B layout of the tensor core (atoms 1024 bytes apart along K,
// 64-channel blocks BK * 128 bytes apart along M/N).  Border rows of dY are zero (PF invariant), rows outside a
// plane are zero-filled by the TMA unit, so shifted reads never contribute garbage.  fp16x3 split as in the forward
// kernel (dY_hi*X_hi + dY_lo*X_hi + dY_hi*X_lo, fp32 accumulation in TMEM); split-K over row ranges, fp32 atomics.
//
// Warp roles (256 threads, persistent): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (tcgen05.ld -> atomicAdd).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "conv.h"
#include "ptx_sm100.cuh"

namespace cl {

namespace {

constexpr int kBlockM = 128;
constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kBK = 64;              // pixel rows per pipeline stage
constexpr uint32_t kTmemCols = 512;

struct WgradPfParams {
    int num_taps;
    int tap_shift[9], tap_phase[9];
    int Cout, Cin, BN, tiles_co, tiles_ci, splits, kb_per_split, total_kb, phases, nterms;
    int inner_g, inner_x;            // channels per swizzle row (64, or 32 for 32-channel tensors)
    float out_scale;
    const float* scale_dev;          // nullable device scalar multiplied into the result
    int oihw;                        // 1: dw is [Cout][Cin][taps] (torch's OIHW weight layout), 0: [tap][Cout][Cin]
    float* dw;
    int num_stages, accum_stages;
    uint32_t a_bytes, w_bytes, stage_bytes;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// MN-major operand tile: 8-row swizzle atoms of `row_bytes` (128 or 64) bytes; LBO = distance between channel
// blocks, SBO = distance between consecutive 8-row atoms along K; descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t row_bytes)
{
    const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | (layout << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_pf_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradPfParams p)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nA = p.nterms == 3 ? 2 : 1;
    const int items = p.num_taps * p.tiles_co * p.tiles_ci * p.splits;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmG);
        ptx::prefetch_tensormap(&tmX);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; s++) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_s), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // item -> (tap, co tile, ci tile, split); split fastest so that CTAs working on one dW tile run together
    auto decode = [&](int item, int& tap, int& co0, int& ci0, int& kb0, int& kb1) {
        const int split = item % p.splits;
        int r = item / p.splits;
        const int tci = r % p.tiles_ci; r /= p.tiles_ci;
        const int tco = r % p.tiles_co;
        tap = r / p.tiles_co;
        co0 = tco * kBlockM;
        ci0 = tci * p.BN;
        kb0 = split * p.kb_per_split;
        kb1 = kb0 + p.kb_per_split < p.total_kb ? kb0 + p.kb_per_split : p.total_kb;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)nA * (p.a_bytes + p.w_bytes);
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int tap, co0, ci0, kb0, kb1;
                decode(item, tap, co0, ci0, kb0, kb1);
                const int shift = p.tap_shift[tap], xplane = p.tap_phase[tap];
                const int gblk = co0 / p.inner_g, xblk = ci0 / p.inner_x;
                for (int kb = kb0; kb < kb1; kb++) {
                    ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t bar = ptx::smem_u32(&full_bar[stage]);
                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                    ptx::mbar_expect_tx(bar, tx_bytes);
                    tma_load_4d(sa, &tmG, bar, 0, kb * kBK, gblk, 0);
                    tma_load_4d(sw, &tmX, bar, 0, kb * kBK + shift, xblk, xplane);
                    if (nA == 2) {
                        tma_load_4d(sa + p.a_bytes, &tmG, bar, 0, kb * kBK, gblk, 1);
                        tma_load_4d(sw + p.w_bytes, &tmX, bar, 0, kb * kBK + shift, xblk, p.phases + xplane);
                    }
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // both operands MN-major: transpose bits 15 (A) and 16 (B) of the instruction descriptor
        const uint32_t idesc = ptx::make_idesc_f16(kBlockM, p.BN) | (1u << 15) | (1u << 16);
        const uint32_t rowb_g = (uint32_t)p.inner_g * 2u, rowb_x = (uint32_t)p.inner_x * 2u;
        const uint32_t atom_g = 8u * rowb_g, atom_x = 8u * rowb_x;
        const uint32_t lbo_g = (uint32_t)kBK * rowb_g, lbo_x = (uint32_t)kBK * rowb_x;
        int stage = 0, local = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, local++) {
            int tap, co0, ci0, kb0, kb1;
            decode(item, tap, co0, ci0, kb0, kb1);
            const int kblocks = kb1 - kb0;
            const int as = local % p.accum_stages;
            const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
            ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aphase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.BN);
            for (int kbi = 0; kbi < kblocks; kbi++) {
                ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {   // elect.sync keeps the MMA operands in uniform registers
                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                    for (int term = 0; term < p.nterms; term++) {
                        const uint32_t a_addr = sa + (term == 1 ? p.a_bytes : 0u);
                        const uint32_t w_addr = sw + (term == 2 ? p.w_bytes : 0u);
#pragma unroll
                        for (int k = 0; k < kBK / 16; k++) {   // one MMA = 16 rows = two 8-row atoms
                            const uint64_t da = make_mnmajor_desc(a_addr + (uint32_t)k * 2u * atom_g, lbo_g, atom_g, rowb_g);
                            const uint64_t db = make_mnmajor_desc(w_addr + (uint32_t)k * 2u * atom_x, lbo_x, atom_x, rowb_x);
                            ptx::mma_f16_ss(tmem_d, da, db, idesc, (kbi | term | k) != 0 ? 1u : 0u);
                        }
                    }
                    ptx::mma_commit(ptx::smem_u32(&empty_bar[stage]));
                    if (kbi == kblocks - 1) ptx::mma_commit(ptx::smem_u32(&tfull_bar[as]));
                }
                __syncwarp();
                if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int local = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, local++) {
            int tap, co0, ci0, kb0, kb1;
            decode(item, tap, co0, ci0, kb0, kb1);
            const int as = local % p.accum_stages;
            const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
            ptx::mbar_wait(ptx::smem_u32(&tfull_bar[as]), aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
            const int co = co0 + q * 32 + lane;
            const float scale = p.out_scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
            const size_t estride = p.oihw ? (size_t)p.num_taps : 1;
            float* dst = p.oihw ? p.dw + ((size_t)co * p.Cin + ci0) * p.num_taps + tap
                                : p.dw + ((size_t)tap * p.Cout + co) * p.Cin + ci0;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                uint32_t u[32];
                ptx::tmem_ld_32x32(taddr + (uint32_t)c0, u);
                ptx::tmem_ld_wait();
                if (co < p.Cout) {
#pragma unroll
                    for (int j = 0; j < 32; j++) atomicAdd(dst + (size_t)(c0 + j) * estride, __uint_as_float(u[j]) * scale);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
}


__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3)
{
    // 2-SM form: completion is signalled on the barrier at this offset in the pair's leader (even-ranked) CTA
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
        ::"r"(dst), "l"(m), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// CTA-pair variant (tcgen05 cta_group::2) for Cout % 256 == 0 and Cin % 256 == 0: the pair owns a 256 (co) x 256 (ci)
// tile of dW; each CTA loads its own 128 dY channels and HALF of the X tile (128 channels), the pair's MMA unit reads the
// other half from the peer.  8 KB instead of 12 KB of shared-memory reads per CTA and MMA and three 64 KB stages instead
// of two 96 KB ones (the single-CTA kernel is bound by shared-memory bandwidth, like the forward kernel was).
// Protocol as in conv_igemm_pair_kernel: 2-SM TMA loads signal the leader's full barrier, the leader's MMA warp issues
// for the pair and releases stages / publishes accumulators with multicast commits, both epilogues drain their own
// 128 TMEM lanes and signal the leader's accumulator-empty barrier.
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_pf_pair_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradPfParams p)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nA = p.nterms == 3 ? 2 : 1;
    const uint32_t crank = ptx::cluster_ctarank();
    const bool leader = crank == 0;
    const int cluster_id = (int)ptx::cluster_id_x();
    const int num_clusters = (int)ptx::cluster_count_x();
    const int items = p.num_taps * p.tiles_co * p.tiles_ci * p.splits;
    const uint32_t w_half = p.w_bytes / 2;   // this CTA's 128 channels of the X tile

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmG);
        ptx::prefetch_tensormap(&tmX);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; s++) {
            ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 8);   // leader only: 4 epilogue warps of each CTA
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

    auto decode = [&](int item, int& tap, int& co0, int& ci0, int& kb0, int& kb1) {
        const int split = item % p.splits;
        int r = item / p.splits;
        const int tci = r % p.tiles_ci; r /= p.tiles_ci;
        const int tco = r % p.tiles_co;
        tap = r / p.tiles_co;
        co0 = tco * 2 * kBlockM;
        ci0 = tci * p.BN;
        kb0 = split * p.kb_per_split;
        kb1 = kb0 + p.kb_per_split < p.total_kb ? kb0 + p.kb_per_split : p.total_kb;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_pair = 2u * (uint32_t)nA * (p.a_bytes + w_half);
            for (int item = cluster_id; item < items; item += num_clusters) {
                int tap, co0, ci0, kb0, kb1;
                decode(item, tap, co0, ci0, kb0, kb1);
                const int shift = p.tap_shift[tap], xplane = p.tap_phase[tap];
                const int gblk = (co0 + (int)crank * kBlockM) / 64, xblk = (ci0 + (int)crank * (p.BN / 2)) / 64;
                for (int kb = kb0; kb < kb1; kb++) {
                    ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t bar = ptx::smem_u32(&full_bar[stage]);
                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                    if (leader) ptx::mbar_expect_tx(bar, tx_pair);
                    tma_load_4d_pair(sa, &tmG, bar, 0, kb * kBK, gblk, 0);
                    tma_load_4d_pair(sw, &tmX, bar, 0, kb * kBK + shift, xblk, xplane);
                    if (nA == 2) {
                        tma_load_4d_pair(sa + p.a_bytes, &tmG, bar, 0, kb * kBK, gblk, 1);
                        tma_load_4d_pair(sw + w_half, &tmX, bar, 0, kb * kBK + shift, xblk, p.phases + xplane);
                    }
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            const uint32_t idesc = ptx::make_idesc_f16(2 * kBlockM, p.BN) | (1u << 15) | (1u << 16);
            constexpr uint32_t kRowB = 128, kAtom = 8 * kRowB, kLbo = kBK * kRowB;
            int stage = 0, local = 0;
            uint32_t phase = 0;
            for (int item = cluster_id; item < items; item += num_clusters, local++) {
                int tap, co0, ci0, kb0, kb1;
                decode(item, tap, co0, ci0, kb0, kb1);
                const int kblocks = kb1 - kb0;
                const int as = local % p.accum_stages;
                const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
                ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.BN);
                for (int kbi = 0; kbi < kblocks; kbi++) {
                    ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {   // elect.sync keeps the MMA operands in uniform registers
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint32_t sw = sa + (uint32_t)nA * p.a_bytes;
                        for (int term = 0; term < p.nterms; term++) {
                            const uint32_t a_addr = sa + (term == 1 ? p.a_bytes : 0u);
                            const uint32_t w_addr = sw + (term == 2 ? w_half : 0u);
#pragma unroll
                            for (int k = 0; k < kBK / 16; k++) {
                                const uint64_t da = make_mnmajor_desc(a_addr + (uint32_t)k * 2u * kAtom, kLbo, kAtom, kRowB);
                                const uint64_t db = make_mnmajor_desc(w_addr + (uint32_t)k * 2u * kAtom, kLbo, kAtom, kRowB);
                                ptx::mma_f16_ss_pair(tmem_d, da, db, idesc, (kbi | term | k) != 0 ? 1u : 0u);
                            }
                        }
                        ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 0x3);
                        if (kbi == kblocks - 1) ptx::mma_commit_pair(ptx::smem_u32(&tfull_bar[as]), 0x3);
                    }
                    __syncwarp();
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int local = 0;
        for (int item = cluster_id; item < items; item += num_clusters, local++) {
            int tap, co0, ci0, kb0, kb1;
            decode(item, tap, co0, ci0, kb0, kb1);
            const int as = local % p.accum_stages;
            const uint32_t aphase = (uint32_t)(local / p.accum_stages) & 1u;
            ptx::mbar_wait(ptx::smem_u32(&tfull_bar[as]), aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
            const int co = co0 + (int)crank * kBlockM + q * 32 + lane;
            const float scale = p.out_scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
            const size_t estride = p.oihw ? (size_t)p.num_taps : 1;
            float* dst = p.oihw ? p.dw + ((size_t)co * p.Cin + ci0) * p.num_taps + tap
                                : p.dw + ((size_t)tap * p.Cout + co) * p.Cin + ci0;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                uint32_t u[32];
                ptx::tmem_ld_32x32(taddr + (uint32_t)c0, u);
                ptx::tmem_ld_wait();
                if (co < p.Cout) {
#pragma unroll
                    for (int j = 0; j < 32; j++) atomicAdd(dst + (size_t)(c0 + j) * estride, __uint_as_float(u[j]) * scale);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
                else ptx::mbar_arrive_remote(ptx::smem_u32(&tempty_bar[as]), 0u);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();   // no CTA retires while its peer may still read its shared memory or signal its barriers
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

// 4-D view of stacked PF planes [plane][row][C] fp16: (channel in block, row, channel block, plane); a box is
// `inner` channels x kBK rows x `blocks` channel blocks of one plane.  Rows / blocks out of range are zero-filled.
bool make_pf_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t channels, uint64_t planes,
                 uint64_t plane_rows, uint32_t inner, uint32_t blocks)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {inner, rows, channels / inner, planes};
    const cuuint64_t strides[3] = {channels * 2, (cuuint64_t)inner * 2, plane_rows * channels * 2};
    const cuuint32_t box[4] = {inner, (cuuint32_t)kBK, blocks, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = inner * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

const char* conv_wgrad_pf_launch(const ConvWgradPfDesc& d, cudaStream_t stream)
{
    if (d.Cout % 64 != 0) return "conv_wgrad_pf: Cout must be a multiple of 64";
    if (d.Cin % 32 != 0 || (d.Cin > 32 && d.Cin % 64 != 0)) return "conv_wgrad_pf: Cin must be 32 or a multiple of 64";
    if (d.nterms != 1 && d.nterms != 3) return "conv_wgrad_pf: nterms must be 1 or 3";
    if (d.num_taps < 1 || d.num_taps > 9) return "conv_wgrad_pf: 1..9 taps";
    if (d.Mp <= 0) return "conv_wgrad_pf: empty operand";
    const int BN = d.Cin % 256 == 0 ? 256 : (d.Cin % 128 == 0 ? 128 : (d.Cin % 64 == 0 ? 64 : 32));
    const int nA = d.nterms == 3 ? 2 : 1;
    const char* env = getenv("CROSSLOC_B200_WGRAD_PAIR");
    const bool pair = d.Cout % 256 == 0 && BN == 256 && !(env && env[0] == '0');
    WgradPfParams p{};
    p.num_taps = d.num_taps;
    for (int i = 0; i < d.num_taps; i++) { p.tap_shift[i] = d.tap_shift[i]; p.tap_phase[i] = d.tap_phase[i]; }
    p.Cout = d.Cout; p.Cin = d.Cin; p.BN = BN;
    p.tiles_co = pair ? d.Cout / (2 * kBlockM) : (d.Cout + kBlockM - 1) / kBlockM;
    p.tiles_ci = d.Cin / BN;
    p.phases = d.phases; p.nterms = d.nterms; p.out_scale = d.out_scale; p.dw = d.dw;
    p.scale_dev = d.scale_dev; p.oihw = d.oihw;
    p.inner_g = 64;
    p.inner_x = d.Cin >= 64 ? 64 : 32;
    p.total_kb = (d.Mp + kBK - 1) / kBK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = d.num_taps * p.tiles_co * p.tiles_ci;
    const int workers = pair ? sms / 2 : sms;
    int splits = (2 * workers + tiles - 1) / tiles;      // about two items per CTA (pair)
    if (pair) {
        // pick the split count whose item count fills whole waves of CTA pairs best; among near-equal choices the
        // smallest one (every item ends with 65,536 atomics per CTA, so fewer, longer items are cheaper)
        double best = 0.0;
        for (int sp = 1; sp <= 16; sp++) {
            const int it = tiles * sp;
            const double eff = (double)it / (double)(((it + workers - 1) / workers) * workers);
            if (eff > best + 0.02) { best = eff; splits = sp; }
        }
    }
    if (splits > p.total_kb) splits = p.total_kb;
    if (splits < 1) splits = 1;
    p.kb_per_split = (p.total_kb + splits - 1) / splits;
    p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;   // every split owns at least one row block
    p.a_bytes = (uint32_t)(kBlockM * kBK * 2);
    p.w_bytes = (uint32_t)(BN * kBK * 2);
    p.stage_bytes = (uint32_t)nA * (p.a_bytes + (pair ? p.w_bytes / 2 : p.w_bytes));
    p.num_stages = (227 * 1024 - 2048) / (int)p.stage_bytes;
    if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
    if (p.num_stages < 2) return "conv_wgrad_pf: tile does not fit two pipeline stages";
    p.accum_stages = 2 * BN <= (int)kTmemCols ? 2 : 1;
    const size_t smem = (size_t)p.num_stages * p.stage_bytes + 1024;

    CUtensorMap tmG, tmX;
    if (!make_pf_map(&tmG, d.grad, (uint64_t)d.Mp, (uint64_t)d.Cout, (uint64_t)nA, (uint64_t)d.g_plane_rows,
                     (uint32_t)p.inner_g, (uint32_t)(kBlockM / p.inner_g)))
        return "conv_wgrad_pf: cuTensorMapEncodeTiled failed for the gradient matrix";
    if (!make_pf_map(&tmX, d.act, (uint64_t)d.Mp, (uint64_t)d.Cin, (uint64_t)nA * d.phases, (uint64_t)d.x_plane_rows,
                     (uint32_t)p.inner_x, (uint32_t)((pair ? BN / 2 : BN) / p.inner_x)))
        return "conv_wgrad_pf: cuTensorMapEncodeTiled failed for the activation matrix";
    const int items = tiles * p.splits;
    cudaError_t e;
    if (pair) {
        int clusters = workers < items ? workers : items;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaFuncSetAttribute(conv_wgrad_pf_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        e = cudaLaunchKernelEx(&cfg, conv_wgrad_pf_pair_kernel, tmG, tmX, p);
        if (e != cudaSuccess) return cudaGetErrorString(e);
    } else {
        const int grid = items < sms ? items : sms;
        e = cudaFuncSetAttribute(conv_wgrad_pf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        conv_wgrad_pf_kernel<<<grid, kThreads, smem, stream>>>(tmG, tmX, p);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cl
