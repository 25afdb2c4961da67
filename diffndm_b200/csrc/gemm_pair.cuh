// Weight-resident node GEMM on CTA pairs:   C[M, BN*G] = epilogue( A[M,K] (bf16) x W[BN*G, K]^T (bf16) ),
// (K, BN) = (256, 256) projections, (512, 128) node-MLP layer 1, (256, 128) node-MLP layer 2 with the fp32 residual stream.
//
// A pair of CTAs (cluster of two, tcgen05 cta_group::2) owns one BN-column group of W and streams 256-row blocks of A
// through it: each CTA TMA-loads ITS 128 rows of A and holds HALF of the group's weight rows (BN/2 x K bf16, <= 64 KiB,
// resident for the CTA's lifetime); the leader issues one tcgen05.mma of M = 256, N = BN per K = 16 step and each CTA's TMEM
// receives its own 128 rows x all BN columns.  Compared with one CTA per column group (round 1: 128 KiB of weights per CTA,
// an A ring of ONE row block, a stage refilled only after its MMA completed -> a TMA round trip exposed per row block, row-
// block period 4 100 cycles against 1 100 of MMA) the halved weights leave room for an 8-stage ring = two row blocks in
// flight, and the B-operand shared-memory reads per row are halved.
// Measured on B200: node GEMMs 0.444 -> 0.432 ms per step.  The clock64 timelines (scripts/wr_timeline.py) show what is left:
// a CTA sees only 2-5 row-block iterations per launch, so the fixed part -- set-up 2 700 cycles, the first TMA round trip
// 2 500 (up to 9 000 next to the residual epilogue's global traffic), the last drain 3 000 -- is as long as the streaming part.
// Whole 256-column groups for the node MLP (its input read once instead of twice; (512, 256) and (256, 256) + residual fit a
// pair) were measured SLOWER (0.447): half as many iterations per pair and a half-row-block ring for K = 512.  So was a residual
// epilogue on 16 x 256-bit accumulator fragments (tcgen05.ld.16x256b: 8 rows x 32 contiguous bytes per warp instruction, no
// transposition tile, two more ring stages): 0.453 -- 128-byte rows through the tile beat 32-byte sectors.
//
// Persistent, warp-specialised (1 CTA / SM):
//   warp 0     TMA producer : half of the W group once; A tiles (128 x 64 bf16, SWIZZLE_128B) into the ring; the completion
//                             bytes of BOTH CTAs are counted on the leader's full barrier (cp.async.bulk.tensor .cta_group::2)
//   warp 1     MMA issuer   : leader only; commits multicast to both CTAs (stage free, accumulator ready)
//   warps 2..17 epilogue    : 16 accumulator-drain warps (TMEM lane quarter x column span): tcgen05.ld -> bias / residual /
//                             SiLU -> fp32 rows to global and/or bf16 through a swizzled smem slab + TMA store; the residual
//                             epilogue goes through a per-warp transpose tile (coalesced read-modify-write of the fp32 stream)
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int WR_BM = 128;                                    // rows per CTA and row block (the pair covers 256)
constexpr int WR_MAX_STAGES = 8;
constexpr int WR_EPI_SPLIT = 4;                               // epilogue warps per TMEM lane quarter
constexpr int WR_EPI_WARPS = 4 * WR_EPI_SPLIT;
constexpr int WR_THREADS = 64 + 32 * WR_EPI_WARPS;
constexpr int WR_STAGE_BYTES = WR_BM * 64 * 2;               //  16384  (128 rows x 64 k)
constexpr int WR_SLAB16_BYTES = 32 * 32 * 2;                 //   2048  [32 rows][32 bf16], SWIZZLE_64B
constexpr int WR_OUT_BYTES = WR_EPI_WARPS * WR_SLAB16_BYTES;
constexpr int WR_RTILE_LD = 33;                               // fp32 [32][33] transpose tile per epilogue warp (padded: conflict-free)
constexpr int WR_RTILE_BYTES = 32 * WR_RTILE_LD * 4;         //   4224
template <int kK, int kBN>
struct WresShape {
    static constexpr int w_bytes = kK * (kBN / 2) * 2;        // this CTA's half of the column group
    // the residual epilogue (node-MLP layer 2) goes through a per-warp transpose tile so that the fp32 residual stream is
    // read and written with full 128-byte rows
    static constexpr bool has_rtile = (kK == 256 && kBN == 128);
    static constexpr int rtile_bytes = has_rtile ? WR_EPI_WARPS * WR_RTILE_BYTES : 0;
    static constexpr int avail = 232448 - 256 - WR_OUT_BYTES - w_bytes - rtile_bytes;
    static constexpr int stages = avail / WR_STAGE_BYTES > WR_MAX_STAGES ? WR_MAX_STAGES : avail / WR_STAGE_BYTES;
    static constexpr int smem_bytes = w_bytes + stages * WR_STAGE_BYTES + WR_OUT_BYTES + rtile_bytes + 256;
    static_assert(stages >= 4, "gemm_pair: A ring too shallow");
};

struct WresEpilogue {
    const float* bias;        // [BN*G] or nullptr
    const float* residual;    // [M, ldr] fp32 or nullptr (may alias out_f32: every element is read by the thread that writes it)
    int ldr;
    float* out_f32;           // [M, ld_f32] fp32 destination or nullptr; column offset col0_f32
    int ld_f32;
    int col0_f32;
    int has_bf16;             // store bf16 at column offset col0_bf16: through tmap_o16 (TMA) or, in the residual epilogue,
    int col0_bf16;            // directly to out_bf16 [M, ld_bf16]
    int act;                  // 1: SiLU
    __nv_bfloat16* out_bf16;  // residual epilogue only
    int ld_bf16;
};

#ifdef DNDM_EK_TRACE
#ifndef DNDM_WR_TRACE_K            // which shape is traced (default: the merged projection)
#define DNDM_WR_TRACE_K 256
#define DNDM_WR_TRACE_BN 256
#endif
__device__ unsigned long long g_wr_trace[64 * 8];
#define WR_STAMP(it, ev)                                                                                      \
    do {                                                                                                      \
        if (kK == DNDM_WR_TRACE_K && kBN == DNDM_WR_TRACE_BN && blockIdx.x == 0 && (it) < 64)                 \
            g_wr_trace[(it) * 8 + (ev)] = clock64();                                                          \
    } while (0)
#else
#define WR_STAMP(it, ev) do {} while (0)
#endif

// Grid: pairs_full pairs for each of the first n_full column groups (all M rows), then pairs_tail pairs for each remaining
// ("tail") group, which covers only the first M_tail rows (the ligand-row-only projections).  blockIdx.x = 2 * pair + rank.
template <int kK, int kBN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WR_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_o16, int M, int M_tail, int a_col0, int g0, int n_full,
                 int pairs_full, int pairs_tail, WresEpilogue ep) {
    static_assert((kBN == 256 || kBN == 128) && kK % 64 == 0, "unsupported shape");
    constexpr int WR_BN = kBN;
    constexpr int WR_STAGES = WresShape<kK, kBN>::stages;
    constexpr int kWBytes = WresShape<kK, kBN>::w_bytes;
    constexpr int KB = kK / 64;                                   // 64-column k-blocks per row block
    constexpr int WR_CHUNKS = kBN / 32 / WR_EPI_SPLIT;            // 32-column chunks per epilogue warp and row block
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sW = smem;                                           // [KB k-chunks][BN/2 rows][128 B]
    uint8_t* sA = smem + kWBytes;                                 // ring of [128 rows][128 B]
    uint8_t* sOut = sA + WR_STAGES * WR_STAGE_BYTES;
    float* sRT = reinterpret_cast<float*>(sOut + WR_OUT_BYTES);   // [epilogue warp][32][33] fp32 (has_rtile shapes only)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + WR_OUT_BYTES + WresShape<kK, kBN>::rtile_bytes);   // leader's are live
    uint64_t* empty_bar = full_bar + WR_STAGES;      // both CTAs (multicast commit)
    uint64_t* acc_full = empty_bar + WR_STAGES;      // [2] both CTAs (multicast commit)
    uint64_t* acc_empty = acc_full + 2;              // [2] leader's are live
    uint64_t* w_bar = acc_empty + 2;                 // leader's is live: both halves of the weight group have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform role index (see edge_pair.cuh)
    const uint32_t rank = cluster_ctarank();
    const int pair_id = blockIdx.x >> 1;
    int grp, pair_rank, pair_stride;
    if (pair_id < n_full * pairs_full) {
        grp = g0 + pair_id / pairs_full;
        pair_rank = pair_id % pairs_full;
        pair_stride = pairs_full;
    } else {
        const int b2 = pair_id - n_full * pairs_full;
        grp = g0 + n_full + b2 / pairs_tail;
        pair_rank = b2 % pairs_tail;
        pair_stride = pairs_tail;
        M = M_tail;
    }
    const int m_tiles = (M + WR_BM - 1) / WR_BM;
    const int m_pairs = (m_tiles + 1) >> 1;          // 256-row blocks; CTA `rank` takes row block 2 * mp + rank
    const bool has_work = pair_rank < m_pairs;
    pdl_trigger();
    if (threadIdx.x == 0) WR_STAMP(0, 7);            // kernel entry

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < WR_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 2 * WR_EPI_WARPS);
        }
        mbar_init(w_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before_sync();
    cluster_sync_all();                              // barriers of both CTAs initialised before any cross-CTA signal
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) WR_STAMP(2, 7);            // set-up done

    if (warp == 0) {
        if (elect_one() && has_work) {
            // constant weights: loaded under the predecessor's tail (before pdl_wait); bytes of both halves on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(w_bar, 2 * kWBytes);
            const uint32_t w_bar_leader = mapa_shared(smem_u32(w_bar), 0);
#pragma unroll
            for (int kc = 0; kc < KB; ++kc)
                tma_load_2d_pair(sW + kc * (WR_BN / 2 * 128), &tmap_w, w_bar_leader, kc * 64, grp * WR_BN + (int)rank * (WR_BN / 2));
            pdl_wait();                              // A belongs to earlier kernels of the stream
            WR_STAMP(3, 7);
            const uint32_t full_leader = mapa_shared(smem_u32(full_bar), 0);
            int kq = 0, itp = 0;
            for (int mp = pair_rank; mp < m_pairs; mp += pair_stride, ++itp) {
                const int m_blk = 2 * mp + (int)rank;
                WR_STAMP(itp, 0);
                if (ep.residual && ep.ldr == 256 && (kBN == 256 || grp == g0) && m_blk < m_tiles) {   // fp32 residual rows: contiguous
                    const int rows = min(WR_BM, M - m_blk * WR_BM);
                    bulk_prefetch_l2(ep.residual + (size_t)m_blk * WR_BM * 256, (uint32_t)rows * 256 * 4);
                }
                for (int kb = 0; kb < KB; ++kb, ++kq) {
                    const int s = kq % WR_STAGES;
                    const uint32_t ph = (kq / WR_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);                         // the MMA that read this stage (both CTAs') is done
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * WR_STAGE_BYTES);
                    // a row block past the end (odd number of blocks) is all out of bounds: zero-filled, never stored
                    tma_load_2d_pair(sA + s * WR_STAGE_BYTES, &tmap_a, full_leader + (uint32_t)s * 8, a_col0 + kb * 64, m_blk * WR_BM);
                }
                WR_STAMP(itp, 1);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one() && has_work && rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(2 * WR_BM, WR_BN);
            mbar_wait_park_cluster(w_bar, 0);
            int kq = 0, it = 0;
            for (int mp = pair_rank; mp < m_pairs; mp += pair_stride, ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait_park_cluster(&acc_empty[buf], ((it - 2) >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * WR_BN;
                for (int kb = 0; kb < KB; ++kb, ++kq) {
                    const int s = kq % WR_STAGES;
                    const uint32_t ph = (kq / WR_STAGES) & 1;
                    mbar_wait_park_cluster(&full_bar[s], ph);
                    if (kb == 0) WR_STAMP(it, 2);
                    if (kb == KB - 1) WR_STAMP(it, 3);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(sA + s * WR_STAGE_BYTES);
                    const uint32_t sb = smem_u32(sW + kb * (WR_BN / 2 * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_pair(d_tmem, make_kmajor_sw128_desc(sa + k * 32), make_kmajor_sw128_desc(sb + k * 32), idesc,
                                       (kb | k) != 0);
                    umma_commit_pair(&empty_bar[s]);
                }
                umma_commit_pair(&acc_full[buf]);
                WR_STAMP(it, 4);
            }
        }
        __syncwarp();
    } else {
        pdl_wait();                  // the residual and every output buffer belong to earlier kernels of the stream
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int part = (warp - 2) >> 2;                         // which WR_CHUNKS-chunk column span of the BN columns
        uint8_t* s16 = sOut + (warp - 2) * WR_SLAB16_BYTES;
        const uint32_t sw64 = (lane >> 1) & 3;
        const uint32_t acc_empty_leader = mapa_shared(smem_u32(acc_empty), 0);
        int it = 0;
        for (int mp = pair_rank; mp < m_pairs; mp += pair_stride, ++it) {
            const int m_blk = 2 * mp + (int)rank;
            const int buf = it & 1;
            const int row0 = m_blk * WR_BM + q * 32;
            const long grow = (long)row0 + lane;
            const bool row_ok = grow < M;
            // bias + residual of a chunk: independent of the accumulator, so chunk 0's loads are issued before the wait
            auto load_addend = [&](int c, float* f) {
                const int col0 = grp * WR_BN + c * 32;
                if (ep.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                        f[j] = b.x; f[j + 1] = b.y; f[j + 2] = b.z; f[j + 3] = b.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 0.f;
                }
                if (ep.residual && row_ok) {
                    const float* rp = ep.residual + grow * ep.ldr + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 r = *reinterpret_cast<const float4*>(rp + j);
                        f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
                    }
                }
            };
            auto release_acc = [&]() {               // this warp's part of accumulator `buf` is drained
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)buf * 8);
            };
            if constexpr (WresShape<kK, kBN>::has_rtile) {
                if (ep.residual) {
                    // ---- residual epilogue with coalesced global access: h_new = h_old + acc + bias, fp32 in place + bf16 copy.
                    //      The accumulator arrives one ROW per thread; a padded per-warp tile turns it into 4 rows x 128
                    //      contiguous bytes per warp instruction for the global read-modify-write. ----
                    float* tile = sRT + (warp - 2) * (32 * WR_RTILE_LD);
                    const int rr = lane >> 3, cq = (lane & 7) * 4;          // this lane's row-in-group and 4-column quad
                    mbar_wait_park(&acc_full[buf], (it >> 1) & 1);
                    tc_fence_after_sync();
#pragma unroll
                    for (int cc = 0; cc < WR_CHUNKS; ++cc) {
                        const int c = part * WR_CHUNKS + cc;
                        const int col0 = grp * WR_BN + c * 32;
                        uint32_t v[32];
                        tmem_ld32(tmem_base + buf * WR_BN + ((uint32_t)(q * 32) << 16) + c * 32, v);
                        // the old h of the 32 x 32 block: 8 coalesced 16-byte loads per lane, in flight during the TMEM read
                        float4 hold[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const long gr = (long)row0 + i * 4 + rr;
                            hold[i] = gr < M ? *reinterpret_cast<const float4*>(ep.residual + gr * ep.ldr + col0 + cq)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        float bq[4] = {0.f, 0.f, 0.f, 0.f};                 // bias of this lane's column quad
                        if (ep.bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + cq));
                            bq[0] = b.x; bq[1] = b.y; bq[2] = b.z; bq[3] = b.w;
                        }
                        tmem_ld_wait();
                        __syncwarp();                                       // the previous chunk's reads of the tile are done
#pragma unroll
                        for (int j = 0; j < 32; ++j) tile[lane * WR_RTILE_LD + j] = __uint_as_float(v[j]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + rr;
                            const long gr = (long)row0 + r;
                            const float* tp = tile + r * WR_RTILE_LD + cq;
                            float4 o = make_float4(hold[i].x + (tp[0] + bq[0]), hold[i].y + (tp[1] + bq[1]),
                                                   hold[i].z + (tp[2] + bq[2]), hold[i].w + (tp[3] + bq[3]));
                            if (ep.act) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
                            if (gr < M) {
                                if (ep.out_f32)
                                    *reinterpret_cast<float4*>(ep.out_f32 + gr * ep.ld_f32 + ep.col0_f32 + col0 + cq) = o;
                                if (ep.has_bf16)
                                    *reinterpret_cast<uint2*>(ep.out_bf16 + gr * ep.ld_bf16 + ep.col0_bf16 + col0 + cq) =
                                        make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
                            }
                        }
                    }
                    release_acc();
                    continue;
                }
            }
            float f[32];
            load_addend(part * WR_CHUNKS, f);
            mbar_wait_park(&acc_full[buf], (it >> 1) & 1);
            if (warp == 2 && lane == 0) WR_STAMP(it, 5);
            tc_fence_after_sync();
#pragma unroll
            for (int cc = 0; cc < WR_CHUNKS; ++cc) {
                const int c = part * WR_CHUNKS + cc;
                const int col0 = grp * WR_BN + c * 32;
                uint32_t v[32];
                tmem_ld32(tmem_base + buf * WR_BN + ((uint32_t)(q * 32) << 16) + c * 32, v);
                if (cc > 0) load_addend(c, f);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                if (ep.act) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
                }
                if (ep.out_f32 && row_ok) {
                    float* op = ep.out_f32 + grow * ep.ld_f32 + ep.col0_f32 + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
                if (ep.has_bf16) {
                    // the TMA store that last used this slab has finished reading it
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                    uint8_t* rowp = s16 + lane * 64;
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        uint4 pk;
                        pk.x = pack_bf16x2(f[8 * uu], f[8 * uu + 1]);     pk.y = pack_bf16x2(f[8 * uu + 2], f[8 * uu + 3]);
                        pk.z = pack_bf16x2(f[8 * uu + 4], f[8 * uu + 5]); pk.w = pack_bf16x2(f[8 * uu + 6], f[8 * uu + 7]);
                        *reinterpret_cast<uint4*>(rowp + ((uu ^ sw64) << 4)) = pk;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && row0 < M) {
                        tma_store_2d(&tmap_o16, s16, ep.col0_bf16 + col0, row0);
                        tma_store_commit();
                    }
                }
            }
            if (warp == 2 && lane == 0) WR_STAMP(it, 6);
            release_acc();
        }
        if (lane == 0) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before_sync();
    cluster_sync_all();              // no CTA leaves while its peer may still signal it or read its tiles
    if (threadIdx.x == 0) WR_STAMP(1, 7);            // kernel exit
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace dndm
