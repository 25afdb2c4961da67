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
// order; a receiver that continues from the previous tile gets its leading partial written to tile_head[tile] and
// added later, in tile order, by agg_finalize_kernel.  No atomics anywhere.
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int EK_TILE = 128;      // edges per tile (UMMA M)
constexpr int EK_H = 256;         // hidden size (UMMA N and K)
constexpr int EK_THREADS = 384;
constexpr int EK_W2_BYTES = EK_H * EK_H * 2;            // 131072
constexpr int EK_A_BYTES = EK_TILE * EK_H * 2;          //  65536
constexpr int EK_STAGE_BYTES = 2 * EK_TILE * 32 * 4;    //  32768  (one [128 rows][32 cols] fp32 buffer per epilogue group)
constexpr int EK_MISC_BYTES = 2048;
constexpr int EK_SMEM_BYTES = 1024 + EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES + EK_MISC_BYTES;
static_assert(EK_SMEM_BYTES <= 232448, "edge kernel shared memory exceeds 227 KiB");

struct EdgeConsts {         // lives in the kernel-parameter constant bank: warp-uniform reads
    float b2[EK_H];
    float wout[EK_H];
};

struct EdgeProblem {
    const float* P;          // [N, ldpq] : W1a h + b1 (row / receiver part)
    const float* Q;          // [N, ldpq] : W1b h      (col / sender part)
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
};

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

// number of leading rows of a 32-row quarter that continue the previous row's receiver
DNDM_DEVICE int lead_rows(unsigned start_mask) { return start_mask ? (__ffs(start_mask) - 1) : 32; }

template <bool kGCL>
__global__ void __launch_bounds__(EK_THREADS, 1)
edge_mlp_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                const __grid_constant__ EdgeConsts c0, const __grid_constant__ EdgeConsts c1,
                EdgeGraph g, EdgeProblem p0, EdgeProblem p1) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // keep the pointer derived from the __shared__ array (LDS/STS, not generic LD/ST); 1024-B alignment for SW128
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sA = smem + EK_W2_BYTES;
    float* sStage = reinterpret_cast<float*>(smem + EK_W2_BYTES + EK_A_BYTES);      // [2 groups][128][32] swizzled
    uint8_t* misc = smem + EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES;
    int* sRow = reinterpret_cast<int*>(misc);                    // [2][128] receiver per tile row (-1 = padding)
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(misc + 1024);
    uint64_t* mma_done = w_bar + 1;                              // [2]
    uint64_t* tmem_empty = mma_done + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    int* sCont = reinterpret_cast<int*>(tmem_slot + 1);          // [2] tile continues the previous tile's receiver

    const bool second = (blockIdx.y != 0);
    const CUtensorMap* tmap_w = second ? &tmap_w1 : &tmap_w0;
    const EdgeConsts& cc = second ? c1 : c0;
    const EdgeProblem& pr = second ? p1 : p0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = *g.n_edges;
    const int num_tiles = (E + EK_TILE - 1) / EK_TILE;

    if (tid == 0) {
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
        // lane owns k = 4*lane..+3 and 128+4*lane..+3 of the first-layer pre-activation
        float wr[8], w0[8];
        {
            const float4 a = __ldg(reinterpret_cast<const float4*>(pr.w1e + 4 * lane));
            const float4 b = __ldg(reinterpret_cast<const float4*>(pr.w1e + 128 + 4 * lane));
            const float4 c = __ldg(reinterpret_cast<const float4*>(pr.w1e + 256 + 4 * lane));
            const float4 d = __ldg(reinterpret_cast<const float4*>(pr.w1e + 256 + 128 + 4 * lane));
            wr[0] = a.x; wr[1] = a.y; wr[2] = a.z; wr[3] = a.w; wr[4] = b.x; wr[5] = b.y; wr[6] = b.z; wr[7] = b.w;
            w0[0] = c.x; w0[1] = c.y; w0[2] = c.z; w0[3] = c.w; w0[4] = d.x; w0[5] = d.y; w0[6] = d.z; w0[7] = d.w;
        }
        constexpr uint32_t idesc = make_idesc_bf16_f32(EK_TILE, EK_H);
        const uint32_t kc = lane >> 4, u = (lane & 15) >> 1, half = lane & 1;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (it >= 1) mbar_wait(&mma_done[buf ^ 1], ((it - 1) >> 1) & 1);     // A smem free again
            if (it >= 2) mbar_wait(&tmem_empty[buf], ((it - 2) >> 1) & 1);       // D[buf] and sRow[buf] consumed
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
            sRow[buf * 128 + pw * 32 + lane] = my_row;
            if (pw == 0 && lane == 0) sCont[buf] = (tile > 0) ? (g.erow[tile * EK_TILE - 1] == my_row) : 0;
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                const int rj = __shfl_sync(0xffffffffu, my_row, j);
                const int cj = __shfl_sync(0xffffffffu, my_col, j);
                const float rad = __shfl_sync(0xffffffffu, my_rad, j);
                const float r0v = __shfl_sync(0xffffffffu, my_r0, j);
                float v[8];
                if (rj >= 0) {
                    const float* pp = pr.P + (size_t)rj * g.ldpq + 4 * lane;
                    const float* qq = pr.Q + (size_t)cj * g.ldpq + 4 * lane;
                    const float4 pa = __ldg(reinterpret_cast<const float4*>(pp));
                    const float4 pb = __ldg(reinterpret_cast<const float4*>(pp + 128));
                    const float4 qa = __ldg(reinterpret_cast<const float4*>(qq));
                    const float4 qb = __ldg(reinterpret_cast<const float4*>(qq + 128));
                    v[0] = pa.x + qa.x; v[1] = pa.y + qa.y; v[2] = pa.z + qa.z; v[3] = pa.w + qa.w;
                    v[4] = pb.x + qb.x; v[5] = pb.y + qb.y; v[6] = pb.z + qb.z; v[7] = pb.w + qb.w;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = silu_fast(fmaf(w0[i], r0v, fmaf(wr[i], rad, v[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                }
                const uint32_t r = pw * 32 + j;
                uint2 lo, hi;
                lo.x = pack_bf16x2(v[0], v[1]); lo.y = pack_bf16x2(v[2], v[3]);
                hi.x = pack_bf16x2(v[4], v[5]); hi.y = pack_bf16x2(v[6], v[7]);
                const uint32_t off = sw128_offset(r, u) + half * 8;
                *reinterpret_cast<uint2*>(sA + kc * 16384 + off) = lo;
                *reinterpret_cast<uint2*>(sA + (kc + 2) * 16384 + off) = hi;
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            named_bar_sync(1, 128);                   // all four producer warps have written A / sRow
            if (tid == 8 * 32) {
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
            }
        }
    } else {
        // =========================== epilogue groups ===========================
        const int grp = warp >> 2;         // handles tiles with (it & 1) == grp
        const int q = warp & 3;            // TMEM lane quarter of this warp
        const int trow = q * 32 + lane;    // tile row owned in passes 1/2
        float* st = sStage + grp * (EK_TILE * 32);
        const int* rows = sRow + grp * 128;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != grp) continue;
            mbar_wait(&mma_done[grp], (it >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)grp * EK_H + ((uint32_t)(q * 32) << 16);
            const int my_node = rows[trow];
            float dot = 0.f;
            // ---- pass 1: m = SiLU(D + b2), dot with wout; GCL keeps m in TMEM for pass 2 ----
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
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
            if (!kGCL) {
                if (my_node >= 0) pr.head_out[tile * EK_TILE + trow] = pr.out_scale * tanhf(dot);
            } else {
                tmem_st_wait();
                const float att = sigmoid_fast(dot + pr.bout) * pr.out_scale;
                // segment structure of the tile (warp-uniform): bit i of sm[k] = row 32k+i starts a new receiver
                unsigned sm[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = 32 * k + lane;
                    const int cur = rows[r];
                    const int prev = (r > 0) ? rows[r - 1] : (sCont[grp] ? cur : -2);
                    sm[k] = __ballot_sync(0xffffffffu, cur != prev);
                }
                const unsigned my_sm = sm[q];
                // rows of this quarter before i_start continue a receiver owned by an earlier quarter
                const int i_start = (q == 0) ? 0 : lead_rows(my_sm);
                // rows after this quarter that continue its last receiver
                int ext = 0;
                for (int k = q + 1; k < 4; ++k) {
                    const int l = lead_rows(sm[k]);
                    ext += l;
                    if (l < 32) break;
                }
                const bool tile_head0 = (q == 0) && !(my_sm & 1u);     // leading rows continue the previous TILE
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    const int col0 = c * 32;
                    uint32_t v[32];
                    tmem_ld32(d_tmem + col0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) st[trow * 32 + ((j + trow) & 31)] = __uint_as_float(v[j]) * att;
                    named_bar_sync(2 + grp, 128);
                    // ---- column `lane` over the rows of quarter q (+ continuation rows), in row order ----
                    if (i_start < 32) {
                        float val[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int rr = q * 32 + i;
                            val[i] = st[rr * 32 + ((lane + rr) & 31)];
                        }
                        float acc = 0.f;
                        bool head = tile_head0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i > 0 && ((my_sm >> i) & 1u) && i > i_start) {
                                const int node = rows[q * 32 + i - 1];
                                if (head) g.tile_head[(size_t)tile * EK_H + col0 + lane] = acc;
                                else if (node >= 0) g.agg[(size_t)node * EK_H + col0 + lane] = acc;
                                head = false;
                                acc = 0.f;
                            }
                            if (i >= i_start) acc += val[i];
                        }
                        for (int k = 0; k < ext; ++k) {
                            const int rr = q * 32 + 32 + k;
                            acc += st[rr * 32 + ((lane + rr) & 31)];
                        }
                        const int node = rows[q * 32 + 31 + ext];
                        if (head) g.tile_head[(size_t)tile * EK_H + col0 + lane] = acc;
                        else if (node >= 0) g.agg[(size_t)node * EK_H + col0 + lane] = acc;
                    }
                    named_bar_sync(2 + grp, 128);      // stage buffer free for the next chunk
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&tmem_empty[grp]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace dndm
