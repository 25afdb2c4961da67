// Fused per-edge MLP (GCL edge model / EquivariantUpdate heads), egnn_new.py:31-47 and :96-110.
//
// Algebra: the first Linear of every edge MLP acts on [h_row, h_col, e]; its h-parts are hoisted to the
// nodes (P = W1a h + b1, Q = W1b h, one dense node GEMM), so per edge only
//     a   = SiLU(P[row] + Q[col] + w_r * |x_r - x_c|^2 + w_0 * r0)            (gather + fp32 adds, bf16 out)
//     D   = a . W2^T                                                          (tcgen05, M=128 edges, N=256, K=256)
//     m   = SiLU(D + b2)
//   GCL : att = sigmoid(w_a . m + b_a);  agg[row] += att * m / norm           (deterministic segmented sum)
//   HEAD: out[e] = range * tanh(w5 . m)                                       (coord / cross scalar heads)
// remains.
//
// Persistent, warp-specialised CTA (384 threads, 1 CTA / SM):
//   warps 8-11  producers : gather + first-layer epilogue -> bf16 A tile in SWIZZLE_128B K-major smem; one elected
//                           lane issues the 16 tcgen05.mma (M128 N256 K16) of the tile into TMEM buffer (it & 1)
//   warps 0-3   epilogue group 0 (even tiles of this CTA), warps 4-7 epilogue group 1 (odd tiles):
//                           pass 1 (TMEM -> regs): m, attention dot, m written back to TMEM
//                           pass 2: gated messages staged through smem, column-owner threads form the per-receiver
//                           sums in row order and store them
// W2 (128 KiB bf16) is TMA-loaded once per CTA and stays resident.  The MMA of tile i and the production of tile
// i+1 run under the epilogues of tiles i-1 / i.
//
// Determinism: edges are receiver-sorted (CSR).  Each (receiver, column) sum is accumulated by ONE thread in row
// order (a per-tile segment table gives every receiver's row range); a receiver that continues from the previous
// tile gets its leading partial written to tile_head[tile] and added later, in tile order, by agg_finalize_kernel.
// No atomics anywhere.  P and Q are stored in bf16 (half the gather bytes; the sum and the distance terms are fp32).
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int EK_TILE = 128;      // edges per tile (UMMA M)
constexpr int EK_H = 256;         // hidden size (UMMA N and K)
constexpr int EK_THREADS = 384;
constexpr int EK_W2_BYTES = EK_H * EK_H * 2;            // 131072
constexpr int EK_A_BYTES = EK_TILE * EK_H * 2;          //  65536
constexpr int EK_STAGE_BYTES = 2 * EK_TILE * 32 * 4;    //  32768  (one [128 rows][32 cols] fp32 buffer per epilogue group)
constexpr int EK_MISC_BYTES = 2048 + 512;
constexpr int EK_SMEM_BYTES = EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES + EK_MISC_BYTES;
static_assert(EK_SMEM_BYTES <= 232448, "edge kernel shared memory exceeds 227 KiB");

struct EdgeConsts {         // lives in the kernel-parameter constant bank: warp-uniform reads
    float b2[EK_H];
    float wout[EK_H];
};

struct EdgeProblem {
    const __nv_bfloat16* P;  // [N, ldpq] : W1a h + b1 (row / receiver part), bf16
    const __nv_bfloat16* Q;  // [N, ldpq] : W1b h      (col / sender part), bf16
    const float* w1e;        // [2][256]  : first-layer weights of (radial_now, radial_input)
    float* head_out;         // HEAD: [E] scalar per edge
    float bout;              // GCL: attention bias
    float out_scale;         // GCL: 1/normalization_factor ; HEAD: coords_range
};

struct EdgeGraph {
    const int* erow;         // [E] receiver (sorted)
    const int* ecol;         // [E] sender
    const float* r0;         // [E] |x_r - x_c|^2 of the INPUT coordinates (egnn_new.py:228)
    const float* x;          // [N,3] current coordinates
    const int* n_edges;      // device scalar: number of edges this launch covers
    int ldpq;
    float* agg;              // GCL: [N,256]
    float* tile_head;        // GCL: [tiles,256]
    int debug;               // measurement scaffolding: bit0 skip pass 2, bit1 skip producer math, bit2 skip pass 1
    unsigned long long* timeline;   // measurement scaffolding: [tiles_of_cta0][8] globaltimer stamps (or null)
};

DNDM_DEVICE unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %clock64;" : "=l"(t));   // SM cycles (all stamps of one CTA share the SM clock)
    return t;
}
#define TL(slot) do { if (g.timeline && (!(g.debug & 256) || warp < 8) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && it < 64) g.timeline[it * 8 + (slot)] = gtimer(); } while (0)
DNDM_DEVICE void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
DNDM_DEVICE float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// SiLU / sigmoid through one MUFU.TANH:  x*sigmoid(x) = h + h*tanh(h), h = x/2   (|rel err| ~ 2^-11)
DNDM_DEVICE float silu_fast(float x) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
}
DNDM_DEVICE float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

template <bool kGCL>
__global__ void __launch_bounds__(EK_THREADS, 1)
edge_mlp_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                const __grid_constant__ EdgeConsts c0, const __grid_constant__ EdgeConsts c1,
                EdgeGraph g, EdgeProblem p0, EdgeProblem p1) {
    extern __shared__ __align__(1024) uint8_t smem[];            // SW128 operand tiles need 1024-B alignment
    uint8_t* sW = smem;
    uint8_t* sA = smem + EK_W2_BYTES;
    uint8_t* sStage = smem + EK_W2_BYTES + EK_A_BYTES;           // [2 groups][128 rows][128 B], 16-B units XOR-swizzled
    uint8_t* misc = smem + EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES;
    int* sRow = reinterpret_cast<int*>(misc);                    // [4][128] receiver per tile row (-1 = padding), slot it&3
    uint8_t* sSeg = misc + 2048;                                 // [2][132] first row of every receiver segment (+ sentinel)
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(misc + 2048 + 272);
    uint64_t* mma_done = w_bar + 1;                              // [2]
    uint64_t* tmem_empty = mma_done + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    int* sCont = reinterpret_cast<int*>(tmem_slot + 1);          // [4] tile continues the previous tile's receiver

    const bool second = (blockIdx.y != 0);
    const CUtensorMap* tmap_w = second ? &tmap_w1 : &tmap_w0;
    const EdgeConsts& cc = second ? c1 : c0;
    const EdgeProblem& pr = second ? p1 : p0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = *g.n_edges;
    const int num_tiles = (E + EK_TILE - 1) / EK_TILE;

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(tmap_w);
        mbar_init(w_bar, 1);
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        mbar_init(&tmem_empty[0], 128);
        mbar_init(&tmem_empty[1], 128);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
        // =========================== producers (+ MMA issue) ===========================
        const int pw = warp - 8;
        if (tid == 8 * 32 && (int)blockIdx.x < num_tiles) {
            mbar_arrive_expect_tx(w_bar, EK_W2_BYTES);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) tma_load_2d(sW + kc * 32768, tmap_w, w_bar, kc * 64, 0);
        }
        // lane owns k = 8*lane .. 8*lane+7 of the first-layer pre-activation (one 16-byte bf16 unit)
        float wr[8], w0[8];
        {
            const float4 a = __ldg(reinterpret_cast<const float4*>(pr.w1e + 8 * lane));
            const float4 b = __ldg(reinterpret_cast<const float4*>(pr.w1e + 8 * lane + 4));
            const float4 c = __ldg(reinterpret_cast<const float4*>(pr.w1e + 256 + 8 * lane));
            const float4 d = __ldg(reinterpret_cast<const float4*>(pr.w1e + 256 + 8 * lane + 4));
            wr[0] = a.x; wr[1] = a.y; wr[2] = a.z; wr[3] = a.w; wr[4] = b.x; wr[5] = b.y; wr[6] = b.z; wr[7] = b.w;
            w0[0] = c.x; w0[1] = c.y; w0[2] = c.z; w0[3] = c.w; w0[4] = d.x; w0[5] = d.y; w0[6] = d.z; w0[7] = d.w;
        }
        constexpr uint32_t idesc = make_idesc_bf16_f32(EK_TILE, EK_H);
        const uint32_t kc = lane >> 3, u = lane & 7;
        const uint4* Pb = reinterpret_cast<const uint4*>(pr.P) + lane;          // row stride ldpq/8 uint4
        const uint4* Qb = reinterpret_cast<const uint4*>(pr.Q) + lane;
        const size_t ld4 = (size_t)g.ldpq / 8;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (pw == 0) TL(0);
            const int e = tile * EK_TILE + pw * 32 + lane;
            int my_row = -1, my_col = 0;
            float my_rad = 0.f, my_r0 = 0.f;
            if (e < E) {
                my_row = g.erow[e];
                my_col = g.ecol[e];
                my_r0 = g.r0[e];
                const float dx = g.x[3 * my_row] - g.x[3 * my_col];
                const float dy = g.x[3 * my_row + 1] - g.x[3 * my_col + 1];
                const float dz = g.x[3 * my_row + 2] - g.x[3 * my_col + 2];
                my_rad = dx * dx + dy * dy + dz * dz;
            }
            const int ld_row = my_row < 0 ? 0 : my_row;                          // padding rows load node 0 (discarded)
            if (it >= 1) mbar_wait(&mma_done[buf ^ 1], ((it - 1) >> 1) & 1);     // A smem free again
            if (pw == ((g.debug & 128) ? 1 : 0)) TL(1);
            // row ids live in slot it&3: the epilogue of tile it-4 finished long ago (its TMEM buffer was
            // re-acquired for tile it-2), so production never waits for an epilogue
            sRow[(it & 3) * 128 + pw * 32 + lane] = my_row;
            if (pw == 0 && lane == 0) sCont[it & 3] = (tile > 0) ? (g.erow[tile * EK_TILE - 1] == my_row) : 0;
#pragma unroll 1
            for (int j0 = 0; j0 < ((g.debug & 2) ? 0 : 32); j0 += 8) {
                uint4 pv[8], qv[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {                                 // 16 independent 16-byte gathers in flight
                    const int rj = __shfl_sync(0xffffffffu, ld_row, j0 + jj);
                    const int cj = __shfl_sync(0xffffffffu, my_col, j0 + jj);
                    pv[jj] = __ldg(Pb + (size_t)rj * ld4);
                    qv[jj] = __ldg(Qb + (size_t)cj * ld4);
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float rad = __shfl_sync(0xffffffffu, my_rad, j0 + jj);
                    const float r0v = __shfl_sync(0xffffffffu, my_r0, j0 + jj);
                    const uint32_t pw_[4] = {pv[jj].x, pv[jj].y, pv[jj].z, pv[jj].w};
                    const uint32_t qw_[4] = {qv[jj].x, qv[jj].y, qv[jj].z, qv[jj].w};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[2 * i] = __uint_as_float(pw_[i] << 16) + __uint_as_float(qw_[i] << 16);
                        v[2 * i + 1] = __uint_as_float(pw_[i] & 0xffff0000u) + __uint_as_float(qw_[i] & 0xffff0000u);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = silu_fast(fmaf(w0[i], r0v, fmaf(wr[i], rad, v[i])));
                    uint4 o;
                    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                    const uint32_t r = pw * 32 + j0 + jj;
                    *reinterpret_cast<uint4*>(sA + kc * 16384 + sw128_offset(r, u)) = o;
                }
                if (pw == ((g.debug & 128) ? 1 : 0) && (g.debug & 16)) TL(2 + (j0 >> 3));
            }
            if (pw == 0 && !(g.debug & 16)) TL(2);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            named_bar_sync(1, 128);                   // all four producer warps have written A / sRow
            if (tid == 8 * 32) {
                if (it >= 2) mbar_wait(&tmem_empty[buf], ((it - 2) >> 1) & 1);   // D[buf] drained by its epilogue group
                tc_fence_after_sync();
                if (it == 0) mbar_wait(w_bar, 0);
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * EK_H;
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16(d_tmem, make_kmajor_sw128_desc(a0 + kk * 16384 + k * 32),
                                  make_kmajor_sw128_desc(b0 + kk * 32768 + k * 32), idesc, (kk | k) != 0);
                    }
                }
                umma_commit(&mma_done[buf]);
                if (!(g.debug & 16)) TL(3);
            }
            if (!(g.debug & 64)) __syncwarp();
        }
    } else {
        // =========================== epilogue groups ===========================
        const int grp = warp >> 2;         // handles tiles with (it & 1) == grp
        const int q = warp & 3;            // TMEM lane quarter of this warp
        const int trow = q * 32 + lane;    // tile row owned in passes 1/2
        uint8_t* st = sStage + grp * (EK_TILE * 128);
        uint8_t* seg = sSeg + grp * 132;
        // read offset of column `lane` inside a staged row r: ((lane>>2) ^ (r&7))*16 + (lane&3)*4
        const uint32_t lane_u = lane >> 2, lane_w = (lane & 3) * 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != grp) continue;
            const int* rows = sRow + (it & 3) * 128;
            mbar_wait(&mma_done[grp], (it >> 1) & 1);
            if (q == 0 && !(g.debug & (16 | 256))) TL(4);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)grp * EK_H + ((uint32_t)(q * 32) << 16);
            const int my_node = rows[trow];
            int n_seg = 0;
            bool head0 = false;
            if (kGCL) {
                // receiver segments of the tile: row r starts one if its receiver differs from row r-1's
                // (row 0 always starts segment 0; head0 marks it as the continuation of the previous tile's receiver)
                const int prev = trow > 0 ? rows[trow - 1] : -2;
                const bool start = (trow == 0) || (my_node != prev);
                unsigned sm[4];
                const unsigned mine = __ballot_sync(0xffffffffu, start);
                // exchange the four quarter masks through the segment table area (tiny)
                int before = 0;
                {
                    unsigned* xm = reinterpret_cast<unsigned*>(st);      // stage buffer is idle here
                    if (lane == 0) xm[q] = mine;
                    named_bar_sync(2 + grp, 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) sm[k] = xm[k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < q) before += __popc(sm[k]);
                        n_seg += __popc(sm[k]);
                    }
                }
                if (start) seg[before + __popc(mine & ((1u << lane) - 1u))] = (uint8_t)trow;
                if (trow == 0) seg[n_seg] = 128;
                head0 = sCont[it & 3] != 0;
                named_bar_sync(2 + grp, 128);       // table visible; xm reads done before the stage is reused
            }
            float dot = 0.f;
            // ---- pass 1: m = SiLU(D + b2), dot with wout; GCL keeps m in TMEM for pass 2 ----
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (g.debug & 4) break;
                const int col0 = c * 32;
                uint32_t v[32];
                tmem_ld32(d_tmem + col0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float m = silu_fast(__uint_as_float(v[j]) + cc.b2[col0 + j]);
                    dot = fmaf(m, cc.wout[col0 + j], dot);
                    v[j] = __float_as_uint(m);
                }
                if (kGCL) tmem_st32(d_tmem + col0, v);
            }
            if (q == 0 && !(g.debug & (16 | 256))) TL(5);
            if (!kGCL) {
                if (my_node >= 0) pr.head_out[tile * EK_TILE + trow] = pr.out_scale * tanhf(dot);
            } else {
                tmem_st_wait();
                const float att = sigmoid_fast(dot + pr.bout) * pr.out_scale;
#pragma unroll 1
                for (int c = 0; c < ((g.debug & 1) ? 0 : 8); ++c) {
                    const int col0 = c * 32;
                    uint32_t v[32];
                    const bool tl2 = (g.debug & 256) && q == 1 && c == 0;
                    if (tl2) TL(0);
                    tmem_ld32(d_tmem + col0, v);
                    tmem_ld_wait();
                    if (tl2) TL(1);
#pragma unroll
                    for (int uu = 0; uu < 8; ++uu) {
                        float4 o;
                        o.x = __uint_as_float(v[4 * uu]) * att;     o.y = __uint_as_float(v[4 * uu + 1]) * att;
                        o.z = __uint_as_float(v[4 * uu + 2]) * att; o.w = __uint_as_float(v[4 * uu + 3]) * att;
                        *reinterpret_cast<float4*>(st + trow * 128 + ((uu ^ (trow & 7)) << 4)) = o;
                    }
                    if (tl2) TL(2);
                    named_bar_sync(2 + grp, 128);
                    if (tl2) TL(3);
                    // ---- per-receiver column sums in row order: warp q takes segments q, q+4, ...; lane = column ----
                    for (int s = q; s < n_seg; s += 4) {
                        const int a = seg[s], b = seg[s + 1];
                        const int node = rows[a];
                        if (node < 0) continue;
                        // four interleaved partial sums (rows mod 4), combined in a fixed order: short dependency
                        // chains, all loads of an 8-row group in flight together
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        for (int r = a; r < b; r += 8) {
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int rr = r + i;
                                v[i] = (rr < b) ? *reinterpret_cast<const float*>(st + rr * 128 + ((lane_u ^ (rr & 7)) << 4) + lane_w)
                                                : 0.f;
                            }
                            a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3];
                            a0 += v[4]; a1 += v[5]; a2 += v[6]; a3 += v[7];
                        }
                        const float acc = (a0 + a1) + (a2 + a3);
                        if ((g.debug & 32) && acc != 12345.678f) continue;
                        if (s == 0 && head0) g.tile_head[(size_t)tile * EK_H + col0 + lane] = acc;
                        else g.agg[(size_t)node * EK_H + col0 + lane] = acc;
                    }
                    if (tl2) TL(4);
                    named_bar_sync(2 + grp, 128);      // stage buffer free for the next chunk
                    if (tl2) TL(5);
                    if ((g.debug & 256) && q == 1 && c == 7) TL(6);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&tmem_empty[grp]);
            if (q == 0 && !(g.debug & 256)) TL(6);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace dndm
