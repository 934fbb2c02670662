// Thin inline-PTX layer for the Blackwell (sm_100a) tensor path: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05.{alloc,mma,commit,ld}, UMMA shared-memory / instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05" descriptor tables (as mirrored by CUTLASS
// cute/arch/mma_sm100_desc.hpp); nothing here depends on CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace ipdm { namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
        "elect.sync R|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the thread may sleep up to `ns` and is woken by the phase completion, so a waiting warp does
// not burn issue slots (ncu on the width-folded layers: ~45 polls x 9 instructions per wait with the plain form, a third of all
// instructions issued by the 16 transform / epilogue warps)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch error), never as a hung GPU box.
// mbar_wait polls (producers and MMA issuers: single threads on the critical path); mbar_wait_idle sleeps between polls (whole warps
// that wait for data: they must not take issue slots from the working warps).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }   // ~2 s at 1.9 GHz
}
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t it = 0;
    while (!mbar_try_wait_hint(bar, parity, 96u))
        if (++it > 40000000u) { asm volatile("trap;"); }                 // every poll may sleep ~0.1 us: seconds before a stuck pipeline traps
}
// named barrier among `count` threads (count a multiple of 32): warps wait without issuing anything
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
// warpgroup-wide register reallocation (all four warps of an aligned group of four execute the same instruction)
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N)); }

// explicit shared-state-space accesses (32-bit shared addresses): the compiler keeps them out of the generic path and may reorder them freely
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// v if keep, else zeros: two predicated stores instead of four selects and a store
__device__ __forceinline__ void sts128_or_zero(uint32_t saddr, const float4& v, bool keep) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "@p st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n\t"
        "@!p st.shared.v4.f32 [%0], {%6,%6,%6,%6};\n\t}\n"
        :: "r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((uint32_t)keep), "f"(0.f) : "memory");
}
__device__ __forceinline__ void sts64_or_zero(uint32_t saddr, uint32_t a, uint32_t b, bool keep) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
        "@p st.shared.v2.b32 [%0], {%1,%2};\n\t"
        "@!p st.shared.v2.b32 [%0], {%4,%4};\n\t}\n"
        :: "r"(saddr), "r"(a), "r"(b), "r"((uint32_t)keep), "r"(0u) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t saddr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(saddr), "r"(a), "r"(b) : "memory");
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// TMA store of a shared-memory tile (generic-proxy writes must be fenced with fence_proxy_async first) as one bulk async-group;
// tma_store_wait_read: the tile may be overwritten (the engine has read it), not that the bytes are globally visible
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t saddr, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"((uint64_t)m), "r"(saddr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tensor memory -------------------------------------------------------------------------
// one full warp allocates `ncols` (power of two >= 32) columns and publishes the base address in smem
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier receives one arrival when all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of the accumulator: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------
// K-major operand tile stored as rows of 128 bytes with the 128B swizzle (what TMA SWIZZLE_128B writes):
// 8-row groups are 1024 bytes apart (SBO); LBO is unused for swizzled K-major layouts (canonical value 1).
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // leading byte offset (16-byte units)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// same for any swizzle width: layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B; sbo = 8 rows x row bytes
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t saddr, int layout, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// instruction descriptor: D fp32, A/B both `fmt` (1 = bf16, 2 = tf32), both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int m, int n) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr int FMT_BF16 = 1, FMT_TF32 = 2;

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Lean form for issue-bound loops: both descriptors share the upper word (SWIZZLE_128B, SBO 1024, version 1) and differ in the
// 14-bit start-address field, so the issuer keeps 32-bit `lo` words (advance = one integer add) and the instruction is predicated
// on `enable` instead of being branched around.
constexpr uint32_t DESC_SW128_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_tf32_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t enable) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\tmov.b64 db, {%2, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(enable), "r"(DESC_SW128_HI) : "memory");
}

}  // namespace tc

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda dependency) ----
int tmap_encode(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box, CUtensorMapSwizzle swz);

}  // namespace ipdm
