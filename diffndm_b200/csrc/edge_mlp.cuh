// Fused per-edge MLP (GCL edge model / EquivariantUpdate heads), egnn_new.py:31-47 and :96-110.
//
// Algebra: the first Linear of every edge MLP acts on [h_row, h_col, e]; its h-parts are hoisted to the
// nodes (P = W1a h + b1, Q = W1b h, one dense node GEMM), so per edge only
//     a   = SiLU(P[row] + Q[col] + w_r * |x_r - x_c|^2 + w_0 * r0)            (gather + fp32 adds, bf16 out)
//     D   = a . W2^T                                                          (tcgen05, M=128 edges, N=256, K=256)
//     m   = SiLU(D + b2)
//   GCL : att = sigmoid(w_a . m + b_a);  agg[row] += att * m / norm           (deterministic segmented sum)
//   HEAD: out[e] = range * tanh(w5 . m)                                       (coord / cross scalar heads)
// remains.  W2 (128 KiB bf16) stays resident in shared memory for the whole persistent CTA, the A tile is
// produced by the threads straight into the SWIZZLE_128B K-major layout, accumulators live in TMEM.
//
// Determinism: edges are receiver-sorted (CSR).  Inside a tile the per-receiver sums are formed in row order by
// one thread per (32-row quarter, column); quarter partials are chained in order; a receiver that continues from
// the previous tile gets its leading partial written to tile_head[tile] and added later, in tile order, by
// agg_finalize_kernel.  No atomics anywhere.
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int EK_TILE = 128;      // edges per tile (UMMA M)
constexpr int EK_H = 256;         // hidden size (UMMA N and K)
constexpr int EK_THREADS = 256;
constexpr int EK_W2_BYTES = EK_H * EK_H * 2;            // 131072
constexpr int EK_A_BYTES = EK_TILE * EK_H * 2;          //  65536
constexpr int EK_STAGE_BYTES = 2 * EK_TILE * 32 * 4;    //  32768  (two column-halves x [128 rows][32 cols])
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

template <bool kGCL>
__global__ void __launch_bounds__(EK_THREADS, 1)
edge_mlp_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                const __grid_constant__ EdgeConsts c0, const __grid_constant__ EdgeConsts c1,
                EdgeGraph g, EdgeProblem p0, EdgeProblem p1) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sA = smem + EK_W2_BYTES;
    float* sStage = reinterpret_cast<float*>(smem + EK_W2_BYTES + EK_A_BYTES);      // [2][128][32] swizzled
    uint8_t* misc = smem + EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES;
    int* sRow = reinterpret_cast<int*>(misc);                    // [128] receiver per tile row (-1 = padding)
    float* sHeads = reinterpret_cast<float*>(misc + 512);        // [2][4][32]
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(misc + 512 + 1024);
    uint64_t* mma_bar = w_bar + 1;                               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);
    int* sCont = reinterpret_cast<int*>(tmem_slot + 1);          // [2] tile-continuation flag per parity

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
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0 && (int)blockIdx.x < num_tiles) {
        mbar_arrive_expect_tx(w_bar, EK_W2_BYTES);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) tma_load_2d(sW + kc * 32768, tmap_w, w_bar, kc * 64, 0);
    }

    // producer constants: lane owns k = 4*lane..+3 and 128+4*lane..+3
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
    const int q = warp & 3;            // TMEM lane quarter of this warp
    const int hf = warp >> 2;          // column half handled in the epilogue
    const int trow = q * 32 + lane;    // tile row owned in the epilogue

    // v1 schedule: produce A -> MMA -> epilogue, tile after tile (the MMA is ~1/4 of a tile's time; overlapping
    // it with the neighbouring epilogue needs double-buffered row ids and is left for the tuning pass).
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        // ---------------- produce A(tile): gather + first-layer epilogue, bf16, SW128 K-major ----------------
        {
            const int e_base = tile * EK_TILE + warp * 16;
            int my_row = -1, my_col = 0;
            float my_rad = 0.f, my_r0 = 0.f;
            if (lane < 16) {
                const int e = e_base + lane;
                if (e < E) {
                    my_row = g.erow[e];
                    my_col = g.ecol[e];
                    my_r0 = g.r0[e];
                    const float dx = g.x[3 * my_row] - g.x[3 * my_col];
                    const float dy = g.x[3 * my_row + 1] - g.x[3 * my_col + 1];
                    const float dz = g.x[3 * my_row + 2] - g.x[3 * my_col + 2];
                    my_rad = dx * dx + dy * dy + dz * dz;
                }
                sRow[warp * 16 + lane] = my_row;
            }
            if (tid == 0) sCont[0] = (tile > 0) ? (g.erow[tile * EK_TILE - 1] == g.erow[tile * EK_TILE]) : 0;
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {
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
                    for (int i = 0; i < 8; ++i) v[i] = silu_f(fmaf(w0[i], r0v, fmaf(wr[i], rad, v[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                }
                const uint32_t r = warp * 16 + j;
                const uint32_t kc = lane >> 4, u = (lane & 15) >> 1, half = lane & 1;
                uint2 lo, hi;
                lo.x = pack_bf16x2(v[0], v[1]); lo.y = pack_bf16x2(v[2], v[3]);
                hi.x = pack_bf16x2(v[4], v[5]); hi.y = pack_bf16x2(v[6], v[7]);
                const uint32_t off = sw128_offset(r, u) + half * 8;
                *reinterpret_cast<uint2*>(sA + kc * 16384 + off) = lo;
                *reinterpret_cast<uint2*>(sA + (kc + 2) * 16384 + off) = hi;
            }
            fence_proxy_async_smem();
        }
        tc_fence_before_sync();
        __syncthreads();
        // ---------------- MMA: D[128 x 256] = A[128 x 256] . W2^T ----------------
        if (tid == 0) {
            tc_fence_after_sync();
            if (it == 0) mbar_wait(w_bar, 0);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_bf16(tmem_base, make_kmajor_sw128_desc(a0 + kc * 16384 + k * 32),
                              make_kmajor_sw128_desc(b0 + kc * 32768 + k * 32), idesc, (kc | k) != 0);
                }
            }
            umma_commit(&mma_bar[0]);
        }
        mbar_wait(&mma_bar[0], it & 1);
        tc_fence_after_sync();

        // ---------------- epilogue ----------------
        const uint32_t d_tmem = tmem_base + ((uint32_t)(q * 32) << 16);
        const int my_node = sRow[trow];
        float part = 0.f;
        // pass 1: m = SiLU(D + b2); partial dot with wout over this warp's column half; GCL keeps m in TMEM
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int col0 = hf * 128 + c * 32;
            uint32_t v[32];
            tmem_ld32(d_tmem + col0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float m = silu_f(__uint_as_float(v[j]) + cc.b2[col0 + j]);
                part = fmaf(m, cc.wout[col0 + j], part);
                v[j] = __float_as_uint(m);
            }
            if (kGCL) tmem_st32(d_tmem + col0, v);
        }
        if (kGCL) tmem_st_wait();
        float* sPart = sStage;                         // [2][128]: the two column-half partial dots per row
        sPart[hf * 128 + trow] = part;
        __syncthreads();
        const float dot = sPart[trow] + sPart[128 + trow];
        __syncthreads();                               // sPart aliases the stage buffer
        if (!kGCL) {
            if (hf == 0 && my_node >= 0) pr.head_out[tile * EK_TILE + trow] = pr.out_scale * tanhf(dot);
        } else {
            const float att = sigmoid_f(dot + pr.bout) * pr.out_scale;
            float* st = sStage + hf * (EK_TILE * 32);
            float* heads = sHeads + hf * 128;
            const int cont0 = sCont[0];
            const int r_begin = q * 32;
            const int first_node = sRow[r_begin];
            const bool head0 = (q == 0) ? (cont0 != 0) : (first_node >= 0 && sRow[r_begin - 1] == first_node);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = hf * 128 + c * 32;
                uint32_t v[32];
                tmem_ld32(d_tmem + col0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) st[trow * 32 + ((j + trow) & 31)] = __uint_as_float(v[j]) * att;
                __syncthreads();
                // segmented column sums in row order: thread (q, lane) owns column `lane` over the rows of quarter q
                bool is_head = head0;
                bool has_head = false;
                float head_val = 0.f, acc = 0.f;
                int cur = first_node;
                for (int i = 0; i < 32; ++i) {
                    const int rr = r_begin + i;
                    const int node = sRow[rr];
                    if (node != cur) {
                        if (cur >= 0) {
                            if (is_head) { head_val = acc; has_head = true; }
                            else g.agg[(size_t)cur * EK_H + col0 + lane] = acc;      // complete inside the quarter
                        }
                        is_head = false;
                        acc = 0.f;
                        cur = node;
                    }
                    if (node >= 0) acc += st[rr * 32 + ((lane + rr) & 31)];
                }
                bool has_tail = false;                 // trailing segment reaching the quarter end
                if (cur >= 0) {
                    if (is_head) { head_val = acc; has_head = true; }                // whole quarter continues
                    else has_tail = true;
                }
                heads[q * 32 + lane] = has_head ? head_val : 0.f;
                __syncthreads();
                // chain the heads of the following quarters, in order, while they continue the same receiver
                if (q == 0 && has_head) {              // leading partial of a receiver that began in an earlier tile
                    float total = head_val;
                    if (!has_tail && sRow[31] == first_node) {
                        for (int qq = 1; qq < 4; ++qq) {
                            if (sRow[qq * 32] != first_node) break;
                            total += heads[qq * 32 + lane];
                            if (sRow[qq * 32 + 31] != first_node) break;
                        }
                    }
                    g.tile_head[(size_t)tile * EK_H + col0 + lane] = total;
                }
                if (has_tail) {
                    float total = acc;
                    for (int qq = q + 1; qq < 4; ++qq) {
                        if (sRow[qq * 32] != cur) break;
                        total += heads[qq * 32 + lane];
                        if (sRow[qq * 32 + 31] != cur) break;
                    }
                    g.agg[(size_t)cur * EK_H + col0 + lane] = total;
                }
                // st / heads are rewritten only after the next chunk's first __syncthreads
            }
        }
        tc_fence_before_sync();
        __syncthreads();                               // sRow / sA / TMEM reusable by the next tile
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace dndm
