// Node side of one EquivariantBlock in ONE persistent kernel (egnn_new.py:49-66 node_model + the hoisted first-layer
// projections of the next edge MLPs, see edge_mlp.cuh):
//     hid   = SiLU(W3 [h | agg] + b3)                 K = 512, N = 256     (node-MLP layer 1)
//     h    += W4 hid + b4                             K = 256, N = 256     (layer 2 + fp32 residual stream; bf16 copy -> hcat[:, :256])
//     pq[:, 256 g : 256 g + 256] = Wm_g h + bm_g      K = 256, 2-6 groups  (next edge P|Q, coord / cross Q; coord / cross P for ligand rows)
// All three are row-local, so a CTA PAIR (cluster of two, tcgen05 cta_group::2, M = 256) takes a 256-row block through
// the whole chain: `hid` and the bf16 copy of the new `h` never leave shared memory as operands (one 64 KiB SWIZZLE_128B
// K-major activation tile per CTA, written by the epilogue warps, read by the next GEMM), `hcat` is streamed ONCE, and the
// three launches per block (each 2-5 row-block iterations per CTA: set-up, first TMA round trip and last drain as long as
// the streaming part) become one.  Weights are not resident: they stream through the same ring as the
// A k-blocks (16 KiB stages: 128 rows x 64 k of either operand; each CTA loads its half of the 256 weight rows of a group).
//
// Roles (576 threads): warp 0 TMA producer (both CTAs, completion bytes counted on the leader's barriers), warp 1 MMA issuer
// (leader only, multicast commits), warps 2-17 accumulator drain (TMEM lane quarter x 64-column span).  Two TMEM accumulators
// of 256 columns alternate over the GEMMs of the chain ("use" index u): layer 1 -> u, layer 2 -> u + 1, groups -> u + 2 ...
#pragma once
#include "common.cuh"
#include "edge_mlp.cuh"     // tanh_approx

namespace dndm {

constexpr int NF_STAGES = 8;
constexpr int NF_STAGE_BYTES = 128 * 64 * 2;                 //  16384
constexpr int NF_TILE_BYTES = 128 * 256 * 2;                 //  65536  activation tile: [4 k chunks][128 rows][128 B], SW128
constexpr int NF_EPI_WARPS = 16;
constexpr int NF_THREADS = 64 + 32 * NF_EPI_WARPS;
constexpr int NF_SLAB_BYTES = 2048;                          // [32 rows][32 bf16], SWIZZLE_64B, one per drain warp
constexpr int NF_SMEM_BYTES = NF_TILE_BYTES + NF_STAGES * NF_STAGE_BYTES + NF_EPI_WARPS * NF_SLAB_BYTES + 256;
static_assert(NF_SMEM_BYTES <= 232448, "fused node kernel shared memory exceeds 227 KiB");

#ifdef DNDM_EK_TRACE
// development aid: clock64 milestones of CTA 0 -- [0,128) MMA issuer, [128,256) drain warp 2, [256,384) producer (scripts/nf_timeline.py)
__device__ unsigned long long g_nf_trace[384];
#define NF_STAMP(base, idx) do { if (blockIdx.x == 0 && (idx) < 128) g_nf_trace[(base) + (idx)] = clock64(); } while (0)
#else
#define NF_STAMP(base, idx) do {} while (0)
#endif

struct NodeFusedParams {
    const float* b3h;         // [256] HALF of the layer-1 bias (W3 is pre-halved as well: SiLU(x) = h + h tanh(h), h = x/2)
    const float* b4;          // [256]
    const float* bias_m;      // [1536]
    float* h;                 // [M, 256] fp32 residual stream, updated in place
    int M;                    // nodes
    int n_lig;                // rows [0, n_lig) also get projection groups 4 and 5
    int g_begin;              // first projection group (0, or 2 when there is no next block)
};

template <int kDummy = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NF_THREADS, 1)
node_fused_kernel(const __grid_constant__ CUtensorMap tmap_hcat,   // [M, 512] bf16, box 64 x 128, SW128: A of layer 1 and the store of bf16(h)
                  const __grid_constant__ CUtensorMap tmap_w3,     // [256, 512] bf16 (halved), box 64 x 128
                  const __grid_constant__ CUtensorMap tmap_w4,     // [256, 256] bf16, box 64 x 128
                  const __grid_constant__ CUtensorMap tmap_wm,     // [1536, 256] bf16, box 64 x 128
                  const __grid_constant__ CUtensorMap tmap_pq,     // [M, 1536] bf16, box 32 x 32, SW64 (store)
                  NodeFusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sTile = smem;
    uint8_t* sRing = smem + NF_TILE_BYTES;
    uint8_t* sSlab = sRing + NF_STAGES * NF_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sSlab + NF_EPI_WARPS * NF_SLAB_BYTES);   // [stages] leader's are live
    uint64_t* empty_bar = full_bar + NF_STAGES;     // [stages] both CTAs (multicast commit)
    uint64_t* acc_full = empty_bar + NF_STAGES;     // [2] both CTAs (multicast commit)
    uint64_t* acc_empty = acc_full + 2;             // [2] leader's are live: 16 drain warps of each CTA
    uint64_t* tile_ready = acc_empty + 2;           // leader's is live: both CTAs' activation tiles are written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_ready + 1);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 127) / 128;
    const int m_pairs = (m_tiles + 1) >> 1;
    pdl_trigger();

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(&tmap_hcat); tma_prefetch_desc(&tmap_w3); tma_prefetch_desc(&tmap_w4);
        tma_prefetch_desc(&tmap_wm);   tma_prefetch_desc(&tmap_pq);
        for (int s = 0; s < NF_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 2 * NF_EPI_WARPS);
        }
        mbar_init(tile_ready, 2 * NF_EPI_WARPS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before_sync();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // projection groups of a row-block pair: [g_begin, 4) for every row, 4 and 5 when the pair holds ligand rows
    auto n_groups = [&](int mp) { return (4 - p.g_begin) + ((mp * 256 < p.n_lig) ? 2 : 0); };

    if (warp == 0) {
        // =========================== TMA producer: operand tiles in the order the MMA warp consumes them ===========================
        if (elect_one()) {
            pdl_wait();
            const uint32_t full_leader = mapa_shared(smem_u32(full_bar), 0);
            int kq = 0, ts = 0;
            NF_STAMP(256, ts++);
            auto load = [&](const CUtensorMap* tm, int c0, int c1) {
                const int s = kq % NF_STAGES;
                mbar_wait(&empty_bar[s], ((kq / NF_STAGES) & 1) ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * NF_STAGE_BYTES);
                tma_load_2d_pair(sRing + s * NF_STAGE_BYTES, tm, full_leader + (uint32_t)s * 8, c0, c1);
                ++kq;
            };
            for (int mp = pair; mp < m_pairs; mp += num_pairs) {
                const int row0 = (2 * mp + (int)rank) * 128;            // a block past the end is all out of bounds: zero-filled
                if (row0 < p.M) {                                       // the fp32 residual rows of the block: contiguous, wanted by layer 2's drain
                    const int rows = min(128, p.M - row0);
                    bulk_prefetch_l2(p.h + (size_t)row0 * 256, (uint32_t)rows * 256 * 4);
                }
                for (int kb = 0; kb < 8; ++kb) {                        // layer 1: A and W3 k-blocks alternate
                    load(&tmap_hcat, kb * 64, row0);
                    load(&tmap_w3, kb * 64, (int)rank * 128);
                }
                NF_STAMP(256, ts++);
                for (int kb = 0; kb < 4; ++kb) load(&tmap_w4, kb * 64, (int)rank * 128);
                NF_STAMP(256, ts++);
                const int ng = n_groups(mp);
                for (int gi = 0; gi < ng; ++gi)
                    for (int kb = 0; kb < 4; ++kb) load(&tmap_wm, kb * 64, (p.g_begin + gi) * 256 + (int)rank * 128);
                NF_STAMP(256, ts++);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =========================== MMA issuer (leader CTA) ===========================
        if (elect_one() && rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(256, 256);
            const uint32_t tile0 = smem_u32(sTile);
            int kq = 0, u = 0, tq = 0, ts = 0;
            NF_STAMP(0, ts++);
            auto stage_wait = [&]() {                                   // next ring stage has landed in both CTAs
                const int s = kq % NF_STAGES;
                mbar_wait_park_cluster(&full_bar[s], (kq / NF_STAGES) & 1);
                ++kq;
                return s;
            };
            auto acc_begin = [&]() {                                    // accumulator of use u is drained (both CTAs)
                if (u >= 2) mbar_wait_park_cluster(&acc_empty[u & 1], ((u - 2) >> 1) & 1);
                tc_fence_after_sync();
                return tmem_base + (uint32_t)(u & 1) * 256;
            };
            auto acc_end = [&]() { umma_commit_pair(&acc_full[u & 1]); ++u; NF_STAMP(0, ts++); };
            // K = 256 GEMM whose A operand is the activation tile, weights from the ring
            auto gemm_tile = [&](uint32_t d) {
                for (int kb = 0; kb < 4; ++kb) {
                    const int s = stage_wait();
                    tc_fence_after_sync();
                    const uint32_t sb = smem_u32(sRing + s * NF_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_pair(d, make_kmajor_sw128_desc(tile0 + kb * 16384 + k * 32), make_kmajor_sw128_desc(sb + k * 32), idesc,
                                       (kb | k) != 0);
                    umma_commit_pair(&empty_bar[s]);
                }
            };
            for (int mp = pair; mp < m_pairs; mp += num_pairs) {
                {   // ---- layer 1: K = 512, both operands from the ring ----
                    const uint32_t d = acc_begin();
                    for (int kb = 0; kb < 8; ++kb) {
                        const int sa = stage_wait();
                        const int sw = stage_wait();
                        tc_fence_after_sync();
                        const uint32_t a = smem_u32(sRing + sa * NF_STAGE_BYTES), b = smem_u32(sRing + sw * NF_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_pair(d, make_kmajor_sw128_desc(a + k * 32), make_kmajor_sw128_desc(b + k * 32), idesc, (kb | k) != 0);
                        umma_commit_pair(&empty_bar[sa]);
                        umma_commit_pair(&empty_bar[sw]);
                    }
                    acc_end();
                }
                {   // ---- layer 2: A = hid tile ----
                    const uint32_t d = acc_begin();
                    mbar_wait_park_cluster(tile_ready, tq & 1); ++tq;
                    NF_STAMP(0, ts++);
                    tc_fence_after_sync();
                    gemm_tile(d);
                    acc_end();
                }
                // ---- projections: A = bf16(h) tile ----
                mbar_wait_park_cluster(tile_ready, tq & 1); ++tq;
                NF_STAMP(0, ts++);
                const int ng = n_groups(mp);
                for (int gi = 0; gi < ng; ++gi) {
                    const uint32_t d = acc_begin();
                    gemm_tile(d);
                    acc_end();
                }
            }
        }
        __syncwarp();
    } else {
        // =========================== accumulator drain ===========================
        pdl_wait();
        const int ew = warp - 2;
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int part = ew >> 2;                                 // 64-column span of the 256
        const int r = q * 32 + lane;                              // row inside the CTA's 128-row block
        uint8_t* slab = sSlab + ew * NF_SLAB_BYTES;
        const uint32_t sw64 = (lane >> 1) & 3;
        const uint32_t acc_empty_leader = mapa_shared(smem_u32(acc_empty), 0);
        const uint32_t tile_ready_leader = mapa_shared(smem_u32(tile_ready), 0);
        uint8_t* tile_row = sTile + part * 16384 + r * 128;       // this thread's 128-byte row of k chunk `part`
        const bool storer = (warp == 2 && lane == 0);
        int u = 0, ts = 0;
        if (storer) NF_STAMP(128, ts++);
        auto acc_wait = [&]() {
            mbar_wait_park(&acc_full[u & 1], (u >> 1) & 1);
            if (storer) NF_STAMP(128, ts++);
            tc_fence_after_sync();
            return tmem_base + (uint32_t)(u & 1) * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)part * 64;
        };
        auto acc_release = [&]() {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)(u & 1) * 8);
            ++u;
            if (storer) NF_STAMP(128, ts++);
        };
        auto tile_done = [&]() {                                  // this warp's rows / columns of the activation tile are written
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tile_ready_leader);
        };
        // 32 fp32 values of this thread's row -> bf16 -> the four 16-byte units [4 half, 4 half + 4) of its tile row
        auto to_tile = [&](const float* f, int half) {
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                uint4 pk;
                pk.x = pack_bf16x2(f[8 * uu], f[8 * uu + 1]);     pk.y = pack_bf16x2(f[8 * uu + 2], f[8 * uu + 3]);
                pk.z = pack_bf16x2(f[8 * uu + 4], f[8 * uu + 5]); pk.w = pack_bf16x2(f[8 * uu + 6], f[8 * uu + 7]);
                const uint32_t unit = (uint32_t)(half * 4 + uu);
                *reinterpret_cast<uint4*>(tile_row + ((unit ^ (uint32_t)(r & 7)) << 4)) = pk;
            }
        };
        int it = 0;
        for (int mp = pair; mp < m_pairs; mp += num_pairs, ++it) {
            const int row0 = (2 * mp + (int)rank) * 128;
            const long grow = (long)row0 + r;
            const bool row_ok = grow < p.M;
            // ---- layer 1: hid = SiLU(acc + b3) -> activation tile.  The tile was last read by the previous block's projection
            //      MMAs (complete: this use's accumulator was committed after them) and by the TMA store of bf16(h). ----
            {
                if (storer) tma_store_wait_read();
                named_bar_sync(1, 32 * NF_EPI_WARPS);
                const uint32_t d = acc_wait();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(d + half * 32, v);
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.b3h + part * 64 + half * 32 + j));
                        f[j] = b.x; f[j + 1] = b.y; f[j + 2] = b.z; f[j + 3] = b.w;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = f[j] + __uint_as_float(v[j]);          // half the pre-activation
                        f[j] = fmaf(x, tanh_approx(x), x);
                    }
                    to_tile(f, half);
                }
                tile_done();
                acc_release();
            }
            // ---- layer 2: h += acc + b4 (fp32, in place), bf16 copy -> activation tile (the hid tile is no longer read).
            //      The accumulator is read as 16 x 256-bit fragments: a thread holds column PAIRS of two rows, so that a warp
            //      instruction on the fp32 residual stream touches 8 rows x 32 contiguous bytes = full sectors (one row per
            //      thread, the tcgen05.ld.32x32b layout, made this drain 20 000 cycles per row block: 32 half-used sectors
            //      per instruction on 72 MB of read-modify-write). ----
            {
                const int rq = lane >> 2, cq = 2 * (lane & 3);    // fragment coordinates: row within 8, column pair within 8
                // the old h of all 32 rows x 64 columns of this warp (64 registers per thread), in flight during layer 2's MMAs
                float2 ho[2][2][8];
#pragma unroll
                for (int hr = 0; hr < 2; ++hr)
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int rr = 0; rr < 2; ++rr) {
                                const long g = (long)row0 + q * 32 + hr * 16 + rq + 8 * rr;
                                ho[hr][cg][2 * i + rr] = g < p.M ? *reinterpret_cast<const float2*>(p.h + g * 256 + part * 64 + cg * 32 + cq + 8 * i)
                                                                 : make_float2(0.f, 0.f);
                            }
                const uint32_t d = [&] {
                    mbar_wait_park(&acc_full[u & 1], (u >> 1) & 1);
                    if (storer) NF_STAMP(128, ts++);
                    tc_fence_after_sync();
                    return tmem_base + (uint32_t)(u & 1) * 256 + (uint32_t)part * 64;
                }();
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {                  // 16-row halves of the warp's TMEM lane quarter
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg) {              // 32-column groups of the warp's 64 columns
                        const int ra = q * 32 + hr * 16 + rq;     // rows ra and ra + 8 of the CTA's block
                        const int colb = part * 64 + cg * 32 + cq;
                        uint32_t v[16];
                        tmem_ld_16x256b_x4(d + ((uint32_t)(q * 32 + hr * 16) << 16) + (uint32_t)cg * 32, v);
                        float2 bb[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) bb[i] = __ldg(reinterpret_cast<const float2*>(p.b4 + colb + 8 * i));
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
#pragma unroll
                            for (int rr = 0; rr < 2; ++rr) {
                                const int row = ra + 8 * rr;
                                const long g = (long)row0 + row;
                                float2 o;
                                o.x = ho[hr][cg][2 * i + rr].x + (__uint_as_float(v[4 * i + 2 * rr]) + bb[i].x);
                                o.y = ho[hr][cg][2 * i + rr].y + (__uint_as_float(v[4 * i + 2 * rr + 1]) + bb[i].y);
                                if (g < p.M) *reinterpret_cast<float2*>(p.h + g * 256 + colb + 8 * i) = o;
                                const uint32_t unit = (uint32_t)(cg * 4 + i);
                                *reinterpret_cast<uint32_t*>(sTile + part * 16384 + row * 128 + ((unit ^ (uint32_t)(row & 7)) << 4) + 2 * cq) =
                                    pack_bf16x2(o.x, o.y);
                            }
                        }
                    }
                }
                tile_done();
                acc_release();
                // bf16(h) -> hcat[:, 0:256]: four 64-column TMA stores straight from the tile once every drain warp has written it
                named_bar_sync(2, 32 * NF_EPI_WARPS);
                if (storer) {
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) tma_store_2d(&tmap_hcat, sTile + kc * 16384, kc * 64, row0);
                    tma_store_commit();
                }
            }
            // ---- projections: pq[:, 256 g + ...] = acc + bias_m, bf16 through the warp's slab + TMA store ----
            const int ng = n_groups(mp);
            for (int gi = 0; gi < ng; ++gi) {
                const int g = p.g_begin + gi;
                float f[32];
                auto bias32 = [&](int half) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias_m + g * 256 + part * 64 + half * 32 + j));
                        f[j] = b.x; f[j + 1] = b.y; f[j + 2] = b.z; f[j + 3] = b.w;
                    }
                };
                bias32(0);
                const uint32_t d = acc_wait();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(d + half * 32, v);
                    if (half == 1) bias32(1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                    if (lane == 0) tma_store_wait_read();          // the TMA store that last used this slab has read it
                    __syncwarp();
                    uint8_t* rowp = slab + lane * 64;
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        uint4 pk;
                        pk.x = pack_bf16x2(f[8 * uu], f[8 * uu + 1]);     pk.y = pack_bf16x2(f[8 * uu + 2], f[8 * uu + 3]);
                        pk.z = pack_bf16x2(f[8 * uu + 4], f[8 * uu + 5]); pk.w = pack_bf16x2(f[8 * uu + 6], f[8 * uu + 7]);
                        *reinterpret_cast<uint4*>(rowp + (((uint32_t)uu ^ sw64) << 4)) = pk;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && row0 + q * 32 < p.M) {
                        tma_store_2d(&tmap_pq, slab, g * 256 + part * 64 + half * 32, row0 + q * 32);
                        tma_store_commit();
                    }
                }
                acc_release();
            }
        }
        if (lane == 0) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before_sync();
    cluster_sync_all();              // no CTA leaves while its peer may still signal it or read its tiles
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace dndm
