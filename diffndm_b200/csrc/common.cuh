// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX).
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptors" (cross-checked against the
// field tables in CUTLASS cute/arch/mma_sm100_desc.hpp; no CUTLASS code is used).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

namespace dndm {

#define DNDM_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------
DNDM_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DNDM_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

DNDM_DEVICE float silu_f(float v) {
    // x * sigmoid(x) ; __expf -> MUFU.EX2, __fdividef -> MUFU.RCP.  |err| ~ 2 ulp.
    return __fdividef(v, 1.0f + __expf(-v));
}
DNDM_DEVICE float sigmoid_f(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

DNDM_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
DNDM_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DNDM_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DNDM_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
DNDM_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DNDM_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
DNDM_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05.mma operand reads)
DNDM_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) : 2D tiled load, completion on an mbarrier
// ------------------------------------------------------------------------------------------
DNDM_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
DNDM_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2D tiled store smem -> global (bulk async group of the issuing thread)
DNDM_DEVICE void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
DNDM_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DNDM_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// Bulk L2 prefetch of a contiguous global range (address and size multiples of 16 bytes).
DNDM_DEVICE void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
DNDM_DEVICE void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
DNDM_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
template <uint32_t kCols>
DNDM_DEVICE void tmem_alloc(uint32_t* smem_slot) {   // one full warp; writes the TMEM base address to smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
DNDM_DEVICE void tmem_dealloc(uint32_t taddr) {      // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
DNDM_DEVICE void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DNDM_DEVICE void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate), single CTA
DNDM_DEVICE void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete
// (implies tcgen05.fence::before_thread_sync)
DNDM_DEVICE void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i <-> lane base+i)
DNDM_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
DNDM_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
DNDM_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but ties the destination registers of an earlier (still in-flight) tcgen05.ld to the wait so that the
// compiler cannot schedule a read of them above it -- needed when other work is placed between the load and the wait.
DNDM_DEVICE void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// registers -> TMEM (same shape)
DNDM_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
DNDM_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand stored as rows of 128 bytes (64 bf16) with the
// 128-byte swizzle (8-row x 128 B atoms, 1024 B apart).  Fields: [0,14) start>>4, [16,30) LBO>>4 (ignored
// for swizzled K-major, 1), [32,46) SBO>>4 (1024 B between 8-row groups), [46,48) version=1 (sm_100),
// [61,64) layout type (2 = SWIZZLE_128B).
DNDM_DEVICE uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// K-major operand WITHOUT swizzle: 8-row x 16-byte core matrices stored contiguously (128 B); `lbo` = byte distance
// between the two core matrices of a K=16 slice, `sbo` = byte distance between consecutive 8-row groups.
DNDM_DEVICE uint64_t make_kmajor_noswz_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor for kind::f16: c_format [4,6)=1 (F32), a_format [7,10)=1 (BF16), b_format [10,13)=1,
// a_major bit15 = 0 (K), b_major bit16 = 0 (K), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Byte offset of element (row r, 16-byte unit u in [0,8)) inside a SW128 K-major tile whose rows are 128 B.
DNDM_DEVICE uint32_t sw128_offset(uint32_t r, uint32_t u) { return r * 128u + ((u ^ (r & 7u)) << 4); }

// ------------------------------------------------------------------------------------------
// CTA pairs (clusters of two, tcgen05 cta_group::2)
// ------------------------------------------------------------------------------------------
// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of burning issue slots; a failed
// (timed-out) try_wait backs off before polling again
DNDM_DEVICE void mbar_wait_park(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1, %2;\n\t"
        "@P bra DONE_%=;\n\t"
        "nanosleep.u32 %3;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(200000u), "r"(256u)
        : "memory");
}
DNDM_DEVICE uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
DNDM_DEVICE uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// Remote arrive with the default semantics, as CUTLASS's generic-proxy -> 2-SM UMMA pipelines do (fence.proxy.async.shared::cta
// + mbarrier.arrive.shared::cluster on the leader's barrier).  An explicit .release.cluster was measured first: 2 800 cycles
// per arrive (it drains every outstanding memory operation of the warp, global gathers included).
DNDM_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
DNDM_DEVICE void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// wait with cluster-scope acquire: the arrivals come from both CTAs of the pair
DNDM_DEVICE void mbar_wait_park_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P;\n"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1, %2;\n\t"
        "@P bra DONEC_%=;\n\t"
        "nanosleep.u32 %3;\n\t"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(200000u), "r"(128u)
        : "memory");
}
DNDM_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <uint32_t kCols>
DNDM_DEVICE void tmem_alloc_pair(uint32_t* smem_slot) {    // one full warp of EACH CTA of the pair, same smem offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
DNDM_DEVICE void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the pair: M = 256 (128 rows from each CTA's smem), N = 256 (128 B rows from each)
DNDM_DEVICE void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far arrive on the mbarrier at this shared-memory offset in BOTH CTAs when complete
DNDM_DEVICE void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// 2D tiled TMA load issued by either CTA of a pair whose completion bytes are counted on an mbarrier of the LEADER CTA
// (`bar_cluster_addr`: shared::cluster address, see mapa_shared); the data lands in the issuing CTA's own shared memory
DNDM_DEVICE void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}

// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running: everything before pdl_wait() (barrier init, TMEM
// allocation, the TMA load of the resident weights -- constant data only) overlaps the predecessor's tail; pdl_wait()
// returns once the predecessor has completed and its writes are visible.  pdl_trigger() lets the successor be
// scheduled as soon as SM resources free up.  Both are no-ops for a kernel launched without the attribute.
DNDM_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DNDM_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace dndm
