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
constexpr int EK_STAGE_BYTES = 2 * EK_TILE * 33 * 4;    //  33792  (one padded [128 rows][33] fp32 buffer per epilogue group)
constexpr int EK_MISC_BYTES = 2048;
constexpr int EK_SMEM_BYTES = EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES + EK_MISC_BYTES;
static_assert(EK_SMEM_BYTES <= 232448, "edge kernel shared memory exceeds 227 KiB");

struct EdgeConsts {         // lives in the kernel-parameter constant bank: warp-uniform reads
    float b2[EK_H];        // HALF of the second-layer bias (the SiLU argument is evaluated as x/2)
    float wout[EK_H];
};

struct EdgeProblem {
    const __nv_bfloat16* P;  // [N, ldpq] : (W1a h + b1)/2 (row / receiver part), bf16
    const __nv_bfloat16* Q;  // [N, ldpq] : (W1b h)/2      (col / sender part), bf16
    const float* w1e;        // [2][256]  : HALF the first-layer weights of (radial_now, radial_input)
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
// SiLU through one MUFU.TANH on the HALVED argument h = x/2:  x*sigmoid(x) = h + h*tanh(h)   (|rel err| ~ 2^-11).
// The 1/2 is folded into the producing linear map on the host (exact: power of two), so callers pass h directly.
DNDM_DEVICE float silu_half(float h) { return fmaf(h, tanh_approx(h), h); }
DNDM_DEVICE float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of burning issue slots
DNDM_DEVICE void mbar_wait_park(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1, %2;\n\t"
        "@P bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(200000u)
        : "memory");
}

// pass 1 of the epilogue for one row: m = SiLU(D + b2) (b2 pre-halved), dot with wout; GCL writes m back to TMEM.
// `cc` must be one of the __grid_constant__ kernel parameters so that b2/wout become constant-bank operands.
template <bool kGCL>
DNDM_DEVICE float epilogue_pass1(const EdgeConsts& cc, uint32_t d_tmem) {
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int col0 = c * 32;
        uint32_t v[32];
        tmem_ld32(d_tmem + col0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float m = silu_half(fmaf(__uint_as_float(v[j]), 0.5f, cc.b2[col0 + j]));
            dot = fmaf(m, cc.wout[col0 + j], dot);
            v[j] = __float_as_uint(m);
        }
        if (kGCL) tmem_st32(d_tmem + col0, v);
    }
    return dot;
}

constexpr int EK_ST_LD = 33;      // padded row stride (floats) of the staging buffer: conflict-free row writes / column reads

template <bool kGCL>
__global__ void __launch_bounds__(EK_THREADS, 1)
edge_mlp_kernel(const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                const __grid_constant__ EdgeConsts c0, const __grid_constant__ EdgeConsts c1,
                EdgeGraph g, EdgeProblem p0, EdgeProblem p1) {
    extern __shared__ __align__(1024) uint8_t smem[];            // SW128 operand tiles need 1024-B alignment
    uint8_t* sW = smem;
    uint8_t* sA = smem + EK_W2_BYTES;
    float* sStage = reinterpret_cast<float*>(smem + EK_W2_BYTES + EK_A_BYTES);   // [2 groups][128 rows][33]
    uint8_t* misc = smem + EK_W2_BYTES + EK_A_BYTES + EK_STAGE_BYTES;
    int* sRow = reinterpret_cast<int*>(misc);                    // [3][128] receiver per tile row (-1 = padding), slot it%3
    uint8_t* sSeg = misc + 1536;                                 // [2][132] first row of every receiver segment (+ sentinel)
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(misc + 1536 + 272);
    uint64_t* mma_done = w_bar + 1;                              // [2]
    uint64_t* tmem_empty = mma_done + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    int* sCont = reinterpret_cast<int*>(tmem_slot + 1);          // [3] tile continues the previous tile's receiver
    unsigned* sMask = reinterpret_cast<unsigned*>(sCont + 3);    // [2][4] per-quarter segment-start masks

    const bool second = (blockIdx.y != 0);
    const CUtensorMap* tmap_w = second ? &tmap_w1 : &tmap_w0;
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
        constexpr uint32_t idesc = make_idesc_bf16_f32(EK_TILE, EK_H);
        const uint4* Pb = reinterpret_cast<const uint4*>(pr.P) + lane;          // row stride ldpq/8 uint4
        const uint4* Qb = reinterpret_cast<const uint4*>(pr.Q) + lane;
        const uint32_t ld4 = (uint32_t)g.ldpq / 8;
        // byte offset of this lane's 16-byte unit inside A row r: kc*16384 + r*128 + ((u ^ (r&7)) << 4)
        uint8_t* sA_lane = sA + (lane >> 3) * 16384;
        const uint32_t u = lane & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
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
            if (it >= 1) mbar_wait_park(&mma_done[buf ^ 1], ((it - 1) >> 1) & 1);   // A smem free again
            // row ids live in slot it%3: the epilogue of tile it-3 has finished (its TMEM buffer was re-acquired
            // for tile it-1, whose MMA completion was awaited above), so production never waits for an epilogue
            const int slot = it % 3;
            sRow[slot * 128 + pw * 32 + lane] = my_row;
            if (pw == 0 && lane == 0) sCont[slot] = (tile > 0) ? (g.erow[tile * EK_TILE - 1] == my_row) : 0;
            float pf[8];                              // P row of the current receiver, unpacked once per receiver
            int cur_r = -1;
#pragma unroll
            for (int i = 0; i < 8; ++i) pf[i] = 0.f;
#pragma unroll 1
            for (int j0 = 0; j0 < 32; j0 += 8) {
                uint4 pv[8], qv[8];
                int rj[8];
                // gathers of the batch, all in flight together; P only when the receiver changes (rows are sorted)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    rj[jj] = __shfl_sync(0xffffffffu, ld_row, j0 + jj);
                    const int cj = __shfl_sync(0xffffffffu, my_col, j0 + jj);
                    const int prev_r = (jj == 0) ? cur_r : rj[jj - 1];
                    if (rj[jj] != prev_r) pv[jj] = __ldg(Pb + (size_t)rj[jj] * ld4);
                    qv[jj] = __ldg(Qb + (size_t)cj * ld4);
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float rad = __shfl_sync(0xffffffffu, my_rad, j0 + jj);
                    const float r0v = __shfl_sync(0xffffffffu, my_r0, j0 + jj);
                    if (rj[jj] != cur_r) {                                   // warp-uniform
                        cur_r = rj[jj];
                        const uint32_t pw_[4] = {pv[jj].x, pv[jj].y, pv[jj].z, pv[jj].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            pf[2 * i] = __uint_as_float(pw_[i] << 16);
                            pf[2 * i + 1] = __uint_as_float(pw_[i] & 0xffff0000u);
                        }
                    }
                    const uint32_t qw_[4] = {qv[jj].x, qv[jj].y, qv[jj].z, qv[jj].w};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[2 * i] = pf[2 * i] + __uint_as_float(qw_[i] << 16);
                        v[2 * i + 1] = pf[2 * i + 1] + __uint_as_float(qw_[i] & 0xffff0000u);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = silu_half(fmaf(w0[i], r0v, fmaf(wr[i], rad, v[i])));
                    uint4 o;
                    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                    const uint32_t r = pw * 32 + j0 + jj;
                    *reinterpret_cast<uint4*>(sA_lane + r * 128 + ((u ^ (r & 7)) << 4)) = o;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            named_bar_sync(1, 128);                   // all four producer warps have written A / sRow
            if (tid == 8 * 32) {
                if (it >= 2) mbar_wait_park(&tmem_empty[buf], ((it - 2) >> 1) & 1);   // D[buf] drained by its epilogue group
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
            __syncwarp();   // re-converge the issuing lane: without it the warp stays split for the whole next tile
        }
    } else {
        // =========================== epilogue groups ===========================
        const int grp = warp >> 2;         // handles tiles with (it & 1) == grp
        const int q = warp & 3;            // TMEM lane quarter of this warp
        const int trow = q * 32 + lane;    // tile row owned in passes 1/2
        float* st = sStage + grp * (EK_TILE * EK_ST_LD);
        uint8_t* seg = sSeg + grp * 132;
        unsigned* xm = sMask + grp * 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != grp) continue;
            const int slot = it % 3;
            const int* rows = sRow + slot * 128;
            mbar_wait_park(&mma_done[grp], (it >> 1) & 1);
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
                const unsigned mine = __ballot_sync(0xffffffffu, start);
                if (lane == 0) xm[q] = mine;
                named_bar_sync(2 + grp, 128);
                int before = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pc = __popc(xm[k]);
                    if (k < q) before += pc;
                    n_seg += pc;
                }
                if (start) seg[before + __popc(mine & ((1u << lane) - 1u))] = (uint8_t)trow;
                if (trow == 0) seg[n_seg] = 128;
                head0 = sCont[slot] != 0;
                named_bar_sync(2 + grp, 128);       // segment table visible
            }
            // ---- pass 1 ----
            const float dot = second ? epilogue_pass1<kGCL>(c1, d_tmem) : epilogue_pass1<kGCL>(c0, d_tmem);
            if (!kGCL) {
                if (my_node >= 0) pr.head_out[tile * EK_TILE + trow] = pr.out_scale * tanhf(dot);
            } else {
                tmem_st_wait();
                const float att = sigmoid_fast(dot + pr.bout) * pr.out_scale;
                float* my_st = st + trow * EK_ST_LD;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    const int col0 = c * 32;
                    uint32_t v[32];
                    tmem_ld32(d_tmem + col0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) my_st[j] = __uint_as_float(v[j]) * att;
                    named_bar_sync(2 + grp, 128);
                    // ---- per-receiver column sums: warp q takes segments q, q+4, ...; lane = column.  Two interleaved
                    //      partial sums (even / odd rows) combined at the end: fixed order, short dependency chains ----
                    for (int s = q; s < n_seg; s += 4) {
                        const int a = seg[s], b = seg[s + 1];
                        const int node = rows[a];
                        if (node < 0) continue;
                        const float* p = st + a * EK_ST_LD + lane;
                        float a0 = 0.f, a1 = 0.f;
                        int r = a;
                        for (; r + 8 <= b; r += 8, p += 8 * EK_ST_LD) {
                            const float v0 = p[0], v1 = p[EK_ST_LD], v2 = p[2 * EK_ST_LD], v3 = p[3 * EK_ST_LD];
                            const float v4 = p[4 * EK_ST_LD], v5 = p[5 * EK_ST_LD], v6 = p[6 * EK_ST_LD], v7 = p[7 * EK_ST_LD];
                            a0 += v0; a1 += v1; a0 += v2; a1 += v3; a0 += v4; a1 += v5; a0 += v6; a1 += v7;
                        }
                        for (; r < b; ++r, p += EK_ST_LD) a0 += p[0];
                        const float acc = a0 + a1;
                        if (s == 0 && head0) g.tile_head[(size_t)tile * EK_H + col0 + lane] = acc;
                        else g.agg[(size_t)node * EK_H + col0 + lane] = acc;
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
