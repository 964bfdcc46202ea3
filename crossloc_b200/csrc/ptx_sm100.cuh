// Thin inline-PTX wrappers for the sm_100a features the convolution kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace cl {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// Waits until the phase with the given parity has completed.  The wait is bounded (~2 s of SM clocks):
// a protocol bug becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int site = 0)
{
    uint32_t done = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (uint32_t spin = 0;; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((spin & 255u) == 255u && clock64() - t0 > 4000000000ll) {
#ifdef CL_DEBUG_TRAP
            printf("mbar_wait timeout: block %d thread %d site %d bar 0x%x parity %u\n", (int)blockIdx.x, (int)threadIdx.x, site, bar, parity);
            while (clock64() - t0 < 6000000000ll) {}   // let every other stuck waiter report too
#endif
            __trap();
        }
    }
}

// Non-blocking probe: true when the phase with the given parity has completed.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Same as mbar_wait with cluster-scope acquire: pairs with data a peer CTA wrote into this CTA's shared memory
// (st.shared::cluster) before its release.cluster arrive.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int site = 0)
{
    uint32_t done = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (uint32_t spin = 0;; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((spin & 255u) == 255u && clock64() - t0 > 4000000000ll) {
#ifdef CL_DEBUG_TRAP
            printf("mbar_wait_cluster timeout: block %d thread %d site %d bar 0x%x parity %u\n", (int)blockIdx.x, (int)threadIdx.x, site, bar, parity);
            while (clock64() - t0 < 6000000000ll) {}
#endif
            __trap();
        }
    }
}
// arrive (release at cluster scope) on this CTA's own barrier
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// 32-bit store into the shared memory of CTA `cta` of the cluster (same offset as `addr` in this CTA)
__device__ __forceinline__ void st_shared_remote_u32(uint32_t addr, uint32_t cta, uint32_t value)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}\n"
        ::"r"(addr), "r"(cta), "r"(value)
        : "memory");
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m)
{
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(m) : "memory");
}
// 2-D tiled load: coordinates (c0 = innermost element index, c1 = row), completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// Same, multicast: the box lands at the same shared-memory offset of every CTA in cta_mask, and each of those
// CTAs' mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                      uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;\n"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

// 2-SM form: issued by either CTA of a pair, completion is signalled on the mbarrier at this offset in the pair's
// leader (even-ranked) CTA -- the peer bit of the shared::cluster address is cleared.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
        ::"r"(dst), "l"(m), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}

// ------------------------------------------------------------------ clusters
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
        ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_count_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// 2-D tiled store shared -> global (bulk async group); rows/columns outside the tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n"
                 ::"l"(m), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// waits until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }
// makes generic-proxy writes to shared memory visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulation; issued by one thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with e4m3 (fp8) inputs, K = 32 per instruction.
__device__ __forceinline__ void mma_f8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

// Same, arriving on the barrier at this offset in every CTA of cta_mask (stage release under multicast loads).
__device__ __forceinline__ void mma_commit_multicast(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(bar), "h"(cta_mask)
                 : "memory");
}

// ---- cta_group::2: one MMA spans the CTA pair (M = 256: 128 accumulator rows in each CTA's TMEM, every CTA
// supplies its own A rows and half of the B rows from its shared memory); issued by the leader CTA only.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D = A * B + D * 2^-SHIFT (scale-input-d, kind::f16 only): folds an accumulator that carries a power-of-two scale
template <int SHIFT>
__device__ __forceinline__ void mma_f16_ss_pair_scaled_d(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p, %4;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "n"(SHIFT)
        : "memory");
}
__device__ __forceinline__ void mma_f8_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- block-scaled FP4 (kind::mxf4: e2m1 operands, one ue8m0 scale per 32 elements along K, K = 64 per instruction).
// Instruction descriptor (block-scaled form): e2m1 x e2m1 -> fp32, both operands K-major, ue8m0 scales; a_sf / b_sf =
// byte offset (0 or 2) of the instruction's two scale bytes inside the 32-bit scale column of tensor memory.
__device__ __host__ __forceinline__ uint32_t make_idesc_mxf4(int M, int N, uint32_t a_sf, uint32_t b_sf)
{
    return (b_sf << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24) | (a_sf << 29);
}
// D[tmem] (+)= (A * 2^sfa)[smem] * (B * 2^sfb)[smem]^T across the CTA pair; sfa / sfb are tensor-memory column addresses
// of the scale factors (row r of the operand: lane r % 32, column r / 32, replicated in the four lane quarters)
__device__ __forceinline__ void mma_mxf4_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                                 uint32_t tmem_sfa, uint32_t tmem_sfb)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// Shared-memory descriptor of one 32-row x 16-byte scale-factor block (512 contiguous bytes, no swizzle).
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
// shared memory -> tensor memory, 32 rows x 128 bits broadcast to the four lane quarters (4 columns); with cta_group::2
// each CTA of the pair copies from its own shared memory into its own tensor memory.  Ordered with the MMAs of the
// issuing thread (same pipe).
__device__ __forceinline__ void tmem_cp_sf_pair(uint32_t tmem_addr, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;\n" ::"r"(tmem_addr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(bar), "h"(cta_mask)
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t addr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}
// 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t addr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile whose rows are SWIZZLE_BYTES long
// (one swizzle atom = 8 rows): start address, SBO = 8 rows, descriptor version 1 (Blackwell).
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr)
{
    static_assert(SWIZZLE_BYTES == 128 || SWIZZLE_BYTES == 64 || SWIZZLE_BYTES == 32, "unsupported swizzle");
    constexpr uint64_t layout = SWIZZLE_BYTES == 128 ? 2 : (SWIZZLE_BYTES == 64 ? 4 : 6);
    constexpr uint64_t sbo = (8 * SWIZZLE_BYTES) >> 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | (sbo << 32) | ((uint64_t)1 << 46) | (layout << 61);
}

// Measured (tools/dbg_kw_share.py): the XOR pattern of a swizzled K-major tile is a function of the shared-memory ADDRESS
// bits, for the TMA unit that writes it and for the tensor core that reads it alike.  A descriptor may therefore start at any
// 128-byte row of a loaded tile with the matrix-base-offset field (bits 49-51) left at 0 -- setting it to (address >> 7) & 7,
// as the field's description suggests for unaligned starts, reads the wrong rows.  The fp4 convolution uses this to read ONE
// activation tile at three row shifts (the kw taps of a 3x3 filter row).

// Instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M x N tile.
__device__ __host__ __forceinline__ uint32_t make_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace ptx
}  // namespace cl
