// Fused per-edge MLP on a CTA PAIR (tcgen05 cta_group::2): the same computation as edge_mlp.cuh (egnn_new.py:31-47 and
// :96-110), one 256-edge tile per pair and iteration.
//
// Why a pair.  With one CTA per 128-edge tile the resident second-layer weights W2 (128 KiB) leave room for ONE A tile
// (64 KiB): the producers of tile i+1 store only after the MMA of tile i has finished reading it, and the tensor core
// re-reads all of W2 from shared memory for every 128 edges.  As a pair the two CTAs issue ONE tcgen05.mma of M = 256
// (each CTA contributes its own 128 edges as A rows) x N = 256 x K = 16, with the B operand SPLIT over the pair: CTA r holds
// output channels [128 r, 128 r + 128) of W2 (64 KiB).  Per CTA that is 64 KiB of weights + TWO A tiles (2 x 64 KiB): the
// producers run a whole tile ahead of the tensor core, and the B-operand shared-memory reads per edge are halved.  Each
// CTA's TMEM receives its own 128 edges x all 256 channels, so the epilogue is unchanged (epilogue_row).
//
// Roles per CTA (928 threads, <= 64 registers): warps 0-11 epilogue (4 TMEM lane quarters x 3 column groups), 12-27 producers,
// warp 28 lane 0: MMA issue (leader CTA, rank 0) / weight-arrival forwarding (rank 1).  Cross-CTA signalling goes through mbarriers in the LEADER's shared memory:
//   a_full[2]     : 16 producer warps of each CTA arrive (remote arrive, release.cluster) when their rows of A[buf] are written
//   tmem_empty[2] : 8 epilogue warps of each CTA arrive when accumulator buf is drained
//   mma_done[2]   : in BOTH CTAs, signalled by tcgen05.commit.multicast -- accumulator ready (epilogue) and A[buf] free (producers)
#pragma once
#include "edge_mlp.cuh"

namespace dndm {

constexpr int EP_WH_BYTES = 128 * EK_H * 2;             //  65536  this CTA's half of W2: 128 output channels x 256 inputs
constexpr int EP_BX_BYTES = 128 * 16 * 2;               //   4096  bias step, B slice of this CTA's 128 channels
// 29 warps (928 threads, <= 64 registers): 12 epilogue warps = 4 TMEM lane quarters x 3 column groups, 16 producers, 1 issuer
constexpr int EP_EPI_WARPS = 12;
constexpr int EP_PROD_WARPS = 16;
constexpr int EP_THREADS = (EP_EPI_WARPS + EP_PROD_WARPS + 1) * 32;
constexpr int EP_SLAB_BYTES = EP_EPI_WARPS * 2048;      //  24576  message staging, one [32 rows][64 B] SW64 slab per epilogue warp
constexpr int EP_AX_BYTES = 256;                        //    256  bias step, A slice: ONE 8-row group (every row is [1, 1, 0 ...]), SBO = 0
constexpr int EP_DOT_BYTES = 2 * 2 * EK_TILE * 4;       //   2048  partial dot products of column groups 1 and 2, per accumulator
constexpr int EP_MISC_BYTES = EP_SLAB_BYTES + EP_AX_BYTES + EP_BX_BYTES + EK_META_BYTES + EP_DOT_BYTES + 256;
constexpr int EP_SMEM_BYTES = EP_WH_BYTES + 2 * EK_A_BYTES + EP_MISC_BYTES;
static_assert(EP_SMEM_BYTES <= 232448, "pair edge kernel shared memory exceeds 227 KiB");

// kEdgeTypes: the first layer also sees a learned embedding of the edge type (moad_fullatom_cond); its contribution is one
// constant hidden vector per type (EdgeProblem::etab), added to the pre-activation in the producers.
template <bool kGCL, bool kBf16Radial = true, bool kEdgeTypes = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EP_THREADS, 1)
edge_pair_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_msg, const __grid_constant__ EdgeConsts c0, const __grid_constant__ EdgeConsts c1,
                 EdgeGraph g, EdgeProblem p0, EdgeProblem p1) {
    extern __shared__ __align__(1024) uint8_t smem[];            // SW128 operand tiles need 1024-B alignment
    uint8_t* sW = smem;                                          // [4 k chunks][128 channels][128 B], SW128
    uint8_t* sA = smem + EP_WH_BYTES;                            // [2 buffers][4 k chunks][128 edges][128 B], SW128
    uint8_t* misc = smem + EP_WH_BYTES + 2 * EK_A_BYTES;
    uint8_t* sSlab = misc;                                       // [12 epilogue warps][32 rows][64 B] message staging
    uint8_t* sAx = misc + EP_SLAB_BYTES;                         // bias step A slice: [2 k cores][8 rows][16 B], shared by all row groups
    uint8_t* sBx = sAx + EP_AX_BYTES;                            // bias step B slice: [16 row groups][2 k cores][8 rows][16 B]
    int4* sMeta = reinterpret_cast<int4*>(sBx + EP_BX_BYTES);    // [16 producer warps][2 slots][8 edges]
    float* sDot = reinterpret_cast<float*>(sBx + EP_BX_BYTES + EK_META_BYTES);   // [2 accumulators][2 column groups][128 rows]
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(sBx + EP_BX_BYTES + EK_META_BYTES + EP_DOT_BYTES);
    uint64_t* w_peer = w_bar + 1;                                // leader: the peer's half of W2 has landed
    uint64_t* mma_done = w_bar + 2;                              // [2]
    uint64_t* tmem_empty = mma_done + 2;                         // [2] (leader's copy is the live one)
    uint64_t* a_full = tmem_empty + 2;                           // [2] (leader's copy is the live one)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 2);

    const bool second = (blockIdx.y != 0);
    const CUtensorMap* tmap_w = second ? &tmap_w1 : &tmap_w0;
    const EdgeProblem& pr = second ? p1 : p0;

    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast through a shuffle: the compiler then KNOWS the role index is warp-uniform and keeps descriptors, barrier
    // addresses and loop state of the role branches on the uniform datapath (without it: an R2UR pair in front of every LDG / LDTM)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    pdl_trigger();

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(tmap_w);
        mbar_init(w_bar, 1);
        mbar_init(w_peer, 1);
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        mbar_init(&tmem_empty[0], 2 * EP_EPI_WARPS);
        mbar_init(&tmem_empty[1], 2 * EP_EPI_WARPS);
        mbar_init(&a_full[0], 2 * EP_PROD_WARPS);
        mbar_init(&a_full[1], 2 * EP_PROD_WARPS);
        fence_mbar_init();
        // the resident second-layer weights are constant: their load runs under the predecessor's tail (before pdl_wait)
        mbar_arrive_expect_tx(w_bar, EP_WH_BYTES);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) tma_load_2d(sW + kc * 16384, tmap_w, w_bar, kc * 64, (int)rank * 128);
    }
    if (warp == 0) tmem_alloc_pair<512>(tmem_slot);
    // bias step operands (K-major core matrices, no swizzle): A[r][0] = A[r][1] = 1 ; B[n][0] + B[n][1] = b2[n]
    {
        const EdgeConsts& cc = second ? c1 : c0;
        for (int i = tid; i < (EP_AX_BYTES + EP_BX_BYTES) / 16; i += EP_THREADS) {
            const int core = i >> 3, r8 = i & 7;                 // 16-byte row r8 of core matrix `core`
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if ((core & 1) == 0) {                               // k core 0 holds k = 0..7
                if (i < EP_AX_BYTES / 16) {
                    v.x = 0x3f803f80u;                           // bf16 (1, 1)
                } else {
                    const int n = (int)rank * 128 + ((core - EP_AX_BYTES / 128) >> 1) * 8 + r8;
                    const float b = cc.b2[n];
                    const float hi = __bfloat162float(__float2bfloat16_rn(b));
                    v.x = pack_bf16x2(hi, b - hi);
                }
            }
            *reinterpret_cast<uint4*>(sAx + (size_t)i * 16) = v;
        }
        fence_proxy_async_all();
    }
    tc_fence_before_sync();
    cluster_sync_all();                          // barriers of both CTAs initialised, TMEM allocated, bias operands written
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                  // everything below reads / writes what earlier kernels of the stream produce
    const int E = *g.n_edges;
    const int num_tiles = (E + EK_TILE - 1) / EK_TILE;
    const int num_tp = (num_tiles + 1) >> 1;     // tile pairs: pair p, iteration it works on tiles 2 (p + it num_pairs) + rank

    if (warp == EP_EPI_WARPS + EP_PROD_WARPS) {
        // =========================== MMA issuer (one lane of the leader CTA) ===========================
        if (lane == 0) {
            mbar_wait(w_bar, 0);                                                      // own half of W2 has landed
            if (rank != 0) {
                mbar_arrive_cluster(mapa_shared(smem_u32(w_peer), 0));                // ... tell the leader
            } else if (pair < num_tp) {
                constexpr uint32_t idesc = make_idesc_bf16_f32(2 * EK_TILE, EK_H);
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW);
                // A slice: stride 0 between the 8-row groups -- all 128 rows read the same core-matrix pair
                const uint64_t ax = make_kmajor_noswz_desc(smem_u32(sAx), 128, 0), bx = make_kmajor_noswz_desc(smem_u32(sBx), 128, 256);
                mbar_wait_park_cluster(w_peer, 0);
                int it = 0;
                for (int tp = pair; tp < num_tp; tp += num_pairs, ++it) {
                    const int buf = it & 1;
                    mbar_wait_park_cluster(&a_full[buf], (it >> 1) & 1);                          // both CTAs wrote A[buf]
                    EK_STAMP(it, 11);
                    if (it >= 2) mbar_wait_park_cluster(&tmem_empty[buf], ((it - 2) >> 1) & 1);   // D[buf] drained in both CTAs
                    EK_STAMP(it, 5);
                    tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * EK_H;
                    const uint32_t ab = a0 + (uint32_t)buf * EK_A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_bf16_pair(d_tmem, make_kmajor_sw128_desc(ab + kk * 16384 + k * 32),
                                           make_kmajor_sw128_desc(b0 + kk * 16384 + k * 32), idesc, (kk | k) != 0);
                        }
                    }
                    umma_bf16_pair(d_tmem, ax, bx, idesc, 1u);                                    // + b2
                    umma_commit_pair(&mma_done[buf]);
                    EK_STAMP(it, 6);
                }
            }
        }
        __syncwarp();
    } else if (warp >= EP_EPI_WARPS) {
        // =========================== producers ===========================
        const int pw = warp - EP_EPI_WARPS;
        // lane owns k = 8*lane .. 8*lane+7 of the (halved) first-layer pre-activation: one 16-byte bf16 unit
        float wr[8], w0[8];
        {
            const float4 a = __ldg(reinterpret_cast<const float4*>(pr.w1e + 8 * lane));
            const float4 b = __ldg(reinterpret_cast<const float4*>(pr.w1e + 8 * lane + 4));
            const float4 c = __ldg(reinterpret_cast<const float4*>(pr.w1e + 256 + 8 * lane));
            const float4 d = __ldg(reinterpret_cast<const float4*>(pr.w1e + 256 + 8 * lane + 4));
            wr[0] = a.x; wr[1] = a.y; wr[2] = a.z; wr[3] = a.w; wr[4] = b.x; wr[5] = b.y; wr[6] = b.z; wr[7] = b.w;
            w0[0] = c.x; w0[1] = c.y; w0[2] = c.z; w0[3] = c.w; w0[4] = d.x; w0[5] = d.y; w0[6] = d.z; w0[7] = d.w;
        }
        uint32_t wr2[4], w02[4];                     // the same weights as bf16x2 pairs (kBf16Radial producers)
        uint64_t wrf[4], w0f[4];                     // ... and as fp32 pairs (FFMA2)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            wr2[i] = pack_bf16x2(wr[2 * i], wr[2 * i + 1]);
            w02[i] = pack_bf16x2(w0[2 * i], w0[2 * i + 1]);
            wrf[i] = f2_pack(wr[2 * i], wr[2 * i + 1]);
            w0f[i] = f2_pack(w0[2 * i], w0[2 * i + 1]);
        }
        constexpr bool kPacked = kGCL && kBf16Radial; // all-bf16x2 producer arithmetic
        constexpr bool kMixed = kGCL && !kBf16Radial; // bf16x2 P + Q, fp32 radial terms and activation; the coordinate heads stay
                                                      // fp32 throughout (measured: packed doubles the x error)
        // this lane's 16-byte unit of a P / Q row; row address = base + row * (row bytes) as ONE 32 x 32 -> 64-bit multiply-add
        const char* Pb = reinterpret_cast<const char*>(pr.P) + lane * 16;
        const char* Qb = reinterpret_cast<const char*>(pr.Q) + lane * 16;
        const uint32_t ldb = (uint32_t)g.ldpq * 2;
        const uint32_t u = lane & 7;
        int4* meta = sMeta + pw * 16;                // warp-private slots: [2 tiles][8 edges] x (row, col, radial_now, radial_input)
        const int l8 = lane & 7;
        const uint32_t a_full_leader = mapa_shared(smem_u32(a_full), 0);

        struct Meta { int row, col; float r0; };
        auto meta_l1 = [&](int tile) {                       // level 1: edge -> (row, col, r0); padding edges use node 0
            Meta m{0, 0, 0.f};
            const int e = tile * EK_TILE + pw * 8 + l8;
            if (tile < num_tiles && e < E) { m.row = g.erow[e]; m.col = g.ecol[e]; m.r0 = g.r0[e]; }
            return m;
        };
        // level 2: current squared distance (six dependent loads: issued early, consumed by meta_publish); publish to the warp
        struct Pos { float rx, ry, rz, cx, cy, cz; };
        auto meta_pos = [&](const Meta& m) {
            Pos p;
            p.rx = g.x[3 * m.row]; p.ry = g.x[3 * m.row + 1]; p.rz = g.x[3 * m.row + 2];
            p.cx = g.x[3 * m.col]; p.cy = g.x[3 * m.col + 1]; p.cz = g.x[3 * m.col + 2];
            return p;
        };
        auto meta_publish = [&](const Meta& m, const Pos& p, int slot) {
            const float dx = p.rx - p.cx, dy = p.ry - p.cy, dz = p.rz - p.cz;
            const float rad = dx * dx + dy * dy + dz * dz;
            if (lane < 8)
                meta[slot * 8 + l8] = kPacked ? make_int4(m.row, m.col, (int)pack_bf16x2(rad, rad), (int)pack_bf16x2(m.r0, m.r0))
                                           : make_int4(m.row, m.col, __float_as_int(rad), __float_as_int(m.r0));
            __syncwarp();
        };
        // The warp's 8 edges of a tile are processed as four ROUNDS of two edges.  The gathers of round r + 1 (P[row], Q[col]:
        // two 16-byte bf16 units per edge and lane) are issued BEFORE round r is evaluated, across tile boundaries as well
        // (the next tile's metadata is already in the other slot), so a warp always has one round of loads in flight while it
        // computes -- with the A tile double-buffered nothing else hides a producer warp's own gather latency.  (A ring of four
        // single-edge slots, three edges in flight, measured 2 % slower.)
        auto issue2 = [&](int slot, int r, uint4 (&pv)[2], uint4 (&qv)[2]) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int2 rc = *reinterpret_cast<const int2*>(&meta[slot * 8 + r * 2 + jj]);        // (row, col): warp-uniform LDS
                pv[jj] = __ldg(reinterpret_cast<const uint4*>(Pb + (uint64_t)(uint32_t)rc.x * ldb));
                qv[jj] = __ldg(reinterpret_cast<const uint4*>(Qb + (uint64_t)(uint32_t)rc.y * ldb));
            }
        };
        // first-layer activation of two edges, bf16 pack, store into this lane's 16-byte unit of the A rows
        auto finish2 = [&](int slot, int r, const uint4 (&pv)[2], const uint4 (&qv)[2], uint8_t* sA_lane) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int2 zw = *(reinterpret_cast<const int2*>(&meta[slot * 8 + r * 2 + jj]) + 1);  // (radial_now, radial_input)
                uint32_t pw_[4] = {pv[jj].x, pv[jj].y, pv[jj].z, pv[jj].w};
                const uint32_t qw_[4] = {qv[jj].x, qv[jj].y, qv[jj].z, qv[jj].w};
                int etype = 0;
                if (kEdgeTypes) {
                    const int2 rc = *reinterpret_cast<const int2*>(&meta[slot * 8 + r * 2 + jj]);
                    const bool rl = rc.x < g.n_lig, cl = rc.y < g.n_lig;
                    etype = (rl != cl) ? 0 : (rl ? 1 : 2);
                    if (kGCL) {                      // this lane's 8 channels of the type vector, bf16: folded into P
                        const uint4 ct = __ldg(reinterpret_cast<const uint4*>(pr.etab) + etype * 32 + lane);
                        pw_[0] = bf2_add(pw_[0], ct.x); pw_[1] = bf2_add(pw_[1], ct.y);
                        pw_[2] = bf2_add(pw_[2], ct.z); pw_[3] = bf2_add(pw_[3], ct.w);
                    }
                }
                uint4 o;
                if (kPacked) {
                    const uint32_t rad2 = (uint32_t)zw.x, r02 = (uint32_t)zw.y;
                    uint32_t ow[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        ow[i] = bf2_silu_half(bf2_fma(w02[i], r02, bf2_fma(wr2[i], rad2, bf2_add(pw_[i], qw_[i]))));
                    o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                } else if (kMixed) {
                    const float rad = __int_as_float(zw.x), r0v = __int_as_float(zw.y);
                    const uint64_t rad2 = f2_pack(rad, rad), r02 = f2_pack(r0v, r0v);
                    uint32_t ow[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t s2 = bf2_add(pw_[i], qw_[i]);
                        const uint64_t pre = f2_fma(w0f[i], r02, f2_fma(wrf[i], rad2,
                                                    f2_pack(__uint_as_float(s2 << 16), __uint_as_float(s2 & 0xffff0000u))));
                        float lo, hi;
                        f2_unpack(pre, lo, hi);
                        ow[i] = pack_bf16x2(silu_half(lo), silu_half(hi));
                    }
                    o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                } else {
                    // coordinate heads: fp32 throughout, two channels per instruction (FADD2 / FFMA2 -- same IEEE results as scalar)
                    const float rad = __int_as_float(zw.x), r0v = __int_as_float(zw.y);
                    const uint64_t rad2 = f2_pack(rad, rad), r02 = f2_pack(r0v, r0v);
                    uint32_t ow[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint64_t p2 = f2_pack(__uint_as_float(pw_[i] << 16), __uint_as_float(pw_[i] & 0xffff0000u));
                        const uint64_t q2 = f2_pack(__uint_as_float(qw_[i] << 16), __uint_as_float(qw_[i] & 0xffff0000u));
                        uint64_t pre = f2_fma(w0f[i], r02, f2_fma(wrf[i], rad2, f2_add(p2, q2)));
                        if (kEdgeTypes) {
                            const float2 ct = __ldg(reinterpret_cast<const float2*>(pr.etab) + etype * 128 + lane * 4 + i);
                            pre = f2_add(pre, f2_pack(ct.x, ct.y));
                        }
                        float lo, hi;
                        f2_unpack(pre, lo, hi);
                        float m0, m1;
                        f2_unpack(f2_fma(pre, f2_pack(tanh_approx(lo), tanh_approx(hi)), pre), m0, m1);
                        ow[i] = pack_bf16x2(m0, m1);
                    }
                    o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
                const uint32_t rr = pw * 8 + r * 2 + jj;
                *reinterpret_cast<uint4*>(sA_lane + rr * 128 + ((u ^ (rr & 7)) << 4)) = o;
            }
        };

        const int tstride = 2 * num_pairs;                   // tile distance between two iterations of this CTA
        const int tile0 = 2 * pair + (int)rank;
        {
            const Meta m0 = meta_l1(tile0);
            meta_publish(m0, meta_pos(m0), 0);
        }
        Meta m_next = meta_l1(tile0 + tstride);
        uint4 pa[2], qa[2], pb[2], qb[2];                    // two rounds of gathers: one being evaluated, one in flight
        issue2(0, 0, pa, qa);
        int it = 0;
        for (int tp = pair; tp < num_tp; tp += num_pairs, ++it) {
            const int tile = 2 * tp + (int)rank;
            const int buf = it & 1, slot = it & 1;
            uint8_t* sA_lane = sA + buf * EK_A_BYTES + (lane >> 3) * 16384;   // this lane's 16-byte unit of A row r
            if (lane == 0 && pw == 0) EK_STAMP(it, 0);
            if (lane == 0 && pw == 15) EK_STAMP(it, 9);
            const Pos pos_next = meta_pos(m_next);                            // next tile's positions: in flight during two rounds
            if (lane == 0 && pw == 0) EK_STAMP(it, 1);
            if (it >= 2) mbar_wait_park(&mma_done[buf], ((it - 2) >> 1) & 1);      // the MMA two tiles back has read A[buf]
            if (lane == 0 && pw == 0) EK_STAMP(it, 2);
            issue2(slot, 1, pb, qb);
            finish2(slot, 0, pa, qa, sA_lane);
            issue2(slot, 2, pa, qa);
            finish2(slot, 1, pb, qb, sA_lane);
            meta_publish(m_next, pos_next, slot ^ 1);                         // next tile's metadata -> other slot
            m_next = meta_l1(tile + 2 * tstride);                             // level-1 loads two tiles ahead
            if (lane == 0 && pw == 0) EK_STAMP(it, 12);
            issue2(slot, 3, pb, qb);
            finish2(slot, 2, pa, qa, sA_lane);
            issue2(slot ^ 1, 0, pa, qa);                                      // first round of the NEXT tile (padding rows past the end)
            finish2(slot, 3, pb, qb, sA_lane);
            if (lane == 0 && pw == 0) EK_STAMP(it, 3);
            if (lane == 0 && pw == 15) EK_STAMP(it, 10);
            fence_proxy_async_smem();                 // this thread's rows are visible to the tensor core's proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a_full_leader + (uint32_t)buf * 8);
            if (lane == 0 && pw == 0) EK_STAMP(it, 4);
        }
    } else {
        // ===== epilogue (warps 0-11: TMEM lane quarter q = warp % 4, column group cg = warp / 4 drains chunks [0,4) / [4,10) / [10,16)) =====
        // Group 0 gets the short range: it also waits for the other two partial dot products and writes the gate / head value.
        const int cg = warp >> 2;
        const int q = warp & 3;
        const int trow = q * 32 + lane;
        uint8_t* slab = sSlab + warp * 2048;
        const uint32_t tmem_empty_leader = mapa_shared(smem_u32(tmem_empty), 0);
        int it = 0;
        for (int tp = pair; tp < num_tp; tp += num_pairs, ++it) {
            const int tile = 2 * tp + (int)rank;
            const int buf = it & 1;
            const int e = tile * EK_TILE + trow;
            const bool valid = e < E;
            mbar_wait_park(&mma_done[buf], (it >> 1) & 1);
            if (lane == 0 && warp == 0) EK_STAMP(it, 7);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * EK_H + ((uint32_t)(q * 32) << 16);
            const int row0 = tile * EK_TILE + q * 32;
            float dot;
            if (cg == 2) {
                dot = second ? epilogue_row<kGCL, 10, 6>(c1, d_tmem, slab, &tmap_msg, row0, lane)
                             : epilogue_row<kGCL, 10, 6>(c0, d_tmem, slab, &tmap_msg, row0, lane);
            } else if (cg == 1) {
                dot = second ? epilogue_row<kGCL, 4, 6>(c1, d_tmem, slab, &tmap_msg, row0, lane)
                             : epilogue_row<kGCL, 4, 6>(c0, d_tmem, slab, &tmap_msg, row0, lane);
            } else {
                dot = second ? epilogue_row<kGCL, 0, 4>(c1, d_tmem, slab, &tmap_msg, row0, lane)
                             : epilogue_row<kGCL, 0, 4>(c0, d_tmem, slab, &tmap_msg, row0, lane);
            }
            if (cg != 0) {
                sDot[(buf * 2 + cg - 1) * EK_TILE + trow] = dot;
                // barrier id alternates with the accumulator (a warp may run one tile ahead of its partners)
                asm volatile("bar.arrive %0, %1;" ::"r"(2 + 2 * q + buf), "r"(96) : "memory");
            } else {
                named_bar_sync(2 + 2 * q + buf, 96);
                dot += sDot[(buf * 2) * EK_TILE + trow] + sDot[(buf * 2 + 1) * EK_TILE + trow];
                if (valid) {
                    if (kGCL) g.att[e] = sigmoid_fast(dot + pr.bout) * pr.out_scale;
                    else pr.head_out[e] = pr.out_scale * tanhf(dot);
                }
            }
            if (lane == 0 && warp == 0) EK_STAMP(it, 8);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + (uint32_t)buf * 8);   // accumulator drained (this warp's part)
        }
    }
    if (kGCL && warp < EP_EPI_WARPS && lane == 0) tma_store_wait_all();   // message writes complete before exit
    tc_fence_before_sync();
    cluster_sync_all();                          // no CTA leaves while its peer may still signal it or read its tiles
    if (warp == 0) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace dndm
