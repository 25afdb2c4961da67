// Shared pieces of the fused per-edge MLP (GCL edge model / EquivariantUpdate heads), egnn_new.py:31-47 and :96-110: the
// arithmetic helpers, the accumulator-drain routine of the epilogue warps and the per-receiver segment reduction.  The
// kernel itself is edge_pair.cuh (one 256-edge tile per CTA pair, tcgen05 cta_group::2).
//
// Algebra: the first Linear of every edge MLP acts on [h_row, h_col, e]; its h-parts are hoisted to the
// nodes (P = W1a h + b1, Q = W1b h, one dense node GEMM), so per edge only
//     a   = SiLU(P[row] + Q[col] + w_r * |x_r - x_c|^2 + w_0 * r0)            (gather + adds, bf16 out)
//     D   = a . W2^T + b2                                                     (tcgen05, N=256, K=256+16)
//     m   = SiLU(D)
//   GCL : att = sigmoid(w_a . m + b_a);  agg[row] += att * m / norm           (deterministic segmented sum)
//   HEAD: out[e] = range * tanh(w5 . m)                                       (coord / cross scalar heads)
// remains.
// The second-layer bias rides on the tensor core: a 17th K=16 MMA step multiplies a constant A slice (columns 0,1 = 1)
// with a B slice holding b2 split into two bf16 terms (hi + lo), so the accumulator already is the SiLU argument
// (W2 and b2 are pre-halved on the host) and the epilogue spends no instruction on it.
//
// The per-receiver sum is a separate streaming kernel (segment_reduce_kernel): edges are receiver-sorted, so the
// messages of a node are one contiguous block that a warp adds up in edge order -- deterministic, no atomics, and it
// emits the bf16 operand of the node MLP directly.  (An in-kernel segmented reduction through shared memory was
// measured first: its staging + barriers made the epilogue 3x longer than the producer.)
// P and Q are stored in bf16 (half the gather bytes).
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int EK_TILE = 128;      // edges per tile (UMMA M)
constexpr int EK_H = 256;         // hidden size (UMMA N and K)
constexpr int EK_A_BYTES = EK_TILE * EK_H * 2;          //  65536  one A tile: 128 edges x 256 first-layer activations, bf16
constexpr int EK_META_BYTES = 16 * 2 * 8 * 16;          //   4096  [16 producer warps][2 tiles][8 edges] x (row, col, radial, radial_input)

struct EdgeConsts {         // lives in the kernel-parameter constant bank: warp-uniform reads
    float b2[EK_H];        // HALF of the second-layer bias (the SiLU argument is evaluated as x/2); fed to the bias MMA step
    float wout[EK_H];
};

struct EdgeProblem {
    const __nv_bfloat16* P;  // [N, ldpq] : (W1a h + b1)/2 (row / receiver part), bf16
    const __nv_bfloat16* Q;  // [N, ldpq] : (W1b h)/2      (col / sender part), bf16
    const float* w1e;        // [2][256]  : HALF the first-layer weights of (radial_now, radial_input)
    const void* etab;        // edge-type embedding folded through the first layer (dynamics.py:118-127): HALF of W1[:, 2H+2:] E[type]
                             // as [3 types][256] bf16 (GCL) / fp32 (HEAD); nullptr without edge types
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
    int n_lig;               // nodes [0, n_lig) are ligand atoms (edge types: 0 ligand-pocket, 1 ligand-ligand, 2 pocket-pocket)
    int ldpq;
    __nv_bfloat16* msg;      // GCL: [E,256] ungated messages m_ij (bf16)
    float* att;              // GCL: [E] attention gate / normalization_factor
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
// Packed bf16x2 arithmetic (one instruction for two channels) for the GCL producers: the first-layer activation is
// rounded to bf16 for the tensor core anyway, so its pre-activation is assembled directly in bf16x2.
DNDM_DEVICE uint32_t bf2_add(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
DNDM_DEVICE uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
DNDM_DEVICE uint32_t bf2_tanh(uint32_t a) {
    uint32_t d;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
DNDM_DEVICE uint32_t bf2_silu_half(uint32_t h) { return bf2_fma(h, bf2_tanh(h), h); }
// Packed fp32 pair arithmetic (FFMA2 on sm_100a): one instruction for two channels at full fp32 precision.
DNDM_DEVICE uint64_t f2_pack(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
DNDM_DEVICE void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DNDM_DEVICE uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// fp32 pair -> bf16x2 for the staged messages.  F2FP.BF16.F32.PACK_AB executes on the XU pipe at the MUFU rate (ncu:
// sm__inst_executed_pipe_xu = MUFU + F2FP at 8 cycles per warp instruction and SM sub-partition), so the 128 packs per
// accumulator row are a fifth of the kernel's XU time.  The alternative (-DDNDM_EP_PACK_ALU: two integer adds + one byte
// permute on the ALU pipe, round to nearest with ties away from zero) was measured SLOWER on B200 (step 1.70 -> 1.74 ms):
// the epilogue warps are bound by their own instruction stream, not by the XU, and the alternative adds two instructions per
// pair.  Even a truncating pack (one PRMT, no XU) changed nothing.
DNDM_DEVICE uint32_t pack_bf16x2_epi(float lo, float hi) {
#ifdef DNDM_EP_PACK_ALU
    const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
    return __byte_perm(a, b, 0x7632);
#else
    return pack_bf16x2(lo, hi);
#endif
}
DNDM_DEVICE uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
DNDM_DEVICE float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// Epilogue of one tile row (one thread = one edge): m = SiLU(D) (D = half pre-activation incl. bias), dot with wout.
// GCL additionally writes m as bf16 to the message buffer (64 contiguous bytes per 32-column chunk).
// `cc` must be one of the __grid_constant__ kernel parameters so that b2/wout become constant-bank operands.
// GCL additionally stages m as bf16 in this warp's shared-memory slab ([32 rows][64 columns], SWIZZLE_128B) and hands
// every finished 64-column quarter to a TMA store into the message buffer.
// `cc` must be one of the __grid_constant__ kernel parameters so that b2/wout become constant-bank operands.
// The warp drains the 16-column chunks [kChunk0, kChunk0 + kNChunks) of its 32 accumulator rows (both even).
template <bool kGCL, int kChunk0, int kNChunks>
DNDM_DEVICE float epilogue_row(const EdgeConsts& cc, uint32_t d_tmem, uint8_t* slab, const CUtensorMap* tmap_msg, int row0,
                               int lane) {
    uint64_t dot2 = 0ull;                              // (even, odd) partial sums of the dot product
    uint8_t* rowp = slab + lane * 64;                  // slab = [32 rows][32 bf16 = 64 B], SWIZZLE_64B
    const uint32_t sw = (lane >> 1) & 3;
    // one 16-column chunk of the accumulator row: SiLU, partial dot, bf16 staging (+ TMA store of every finished 32 columns)
    // (packed fp32 pairs: one FFMA2 for two SiLUs and one for two terms of the dot product -- half the FMA-pipe instructions
    //  of this thread's 128 columns; the even / odd partial sums meet at the end)
    auto chunk = [&](int c, uint32_t (&v)[16]) {
        const int col0 = c * 16;
        float m[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float h0 = __uint_as_float(v[j]), h1 = __uint_as_float(v[j + 1]);
            const uint64_t h2 = f2_pack(h0, h1);
            const uint64_t m2 = f2_fma(h2, f2_pack(tanh_approx(h0), tanh_approx(h1)), h2);
            dot2 = f2_fma(m2, f2_pack(cc.wout[col0 + j], cc.wout[col0 + j + 1]), dot2);
            f2_unpack(m2, m[j], m[j + 1]);
        }
        if (kGCL) {
            if ((c & 1) == 0) {                        // slab reuse: the previous 32-column TMA store has read it
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                uint4 o;
                o.x = pack_bf16x2_epi(m[j], m[j + 1]);     o.y = pack_bf16x2_epi(m[j + 2], m[j + 3]);
                o.z = pack_bf16x2_epi(m[j + 4], m[j + 5]); o.w = pack_bf16x2_epi(m[j + 6], m[j + 7]);
                const uint32_t unit = (c & 1) * 2 + (j >> 3);
                *reinterpret_cast<uint4*>(rowp + ((unit ^ sw) << 4)) = o;
            }
            if ((c & 1) == 1) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(tmap_msg, slab, (c >> 1) * 32, row0);
                    tma_store_commit();
                }
            }
        }
    };
    // the TMEM load of chunk c + 1 is in flight while chunk c is evaluated (tcgen05.wait::ld covers every outstanding load of
    // the thread, so the next load is issued right after the wait)
    uint32_t va[16], vb[16];
    static_assert(kChunk0 % 2 == 0 && kNChunks % 2 == 0, "chunks are stored in pairs of 32 columns");
    tmem_ld16(d_tmem + kChunk0 * 16, va);
#pragma unroll
    for (int cc_ = 0; cc_ < kNChunks; cc_ += 2) {
        const int c = kChunk0 + cc_;                   // 16-column chunk of the 256 channels
        tmem_ld_wait16(va);
        tmem_ld16(d_tmem + (c + 1) * 16, vb);
        chunk(c, va);
        tmem_ld_wait16(vb);
        if (cc_ + 2 < kNChunks) tmem_ld16(d_tmem + (c + 2) * 16, va);
        chunk(c + 1, vb);
    }
    float d0, d1;
    f2_unpack(dot2, d0, d1);
    return d0 + d1;
}

#ifdef DNDM_EK_TRACE
// Development aid (build with DNDM_EXTRA_NVCC_FLAGS=-DDNDM_EK_TRACE): clock64 stamps of CTA 0 (leader of pair 0) of the GCL kernel,
// [iteration][event], read back through dndm_debug_copy(what = 5).  See scripts/ep_timeline.py for the event list.
__device__ unsigned long long g_ek_trace[64 * 16];
#define EK_STAMP(it, ev)                                                                       \
    do {                                                                                       \
        if (kGCL && blockIdx.x == 0 && (it) < 64) g_ek_trace[(it) * 16 + (ev)] = clock64();   \
    } while (0)
#else
#define EK_STAMP(it, ev) do {} while (0)
#endif

// Deterministic per-receiver reduction of the gated messages (replaces unsorted_segment_sum, egnn_new.py:319-335, and
// feeds the node MLP):  agg[n] = sum_{e in row n} att[e] * msg[e]  in CSR (= edge) order, written as the bf16 operand
// half hcat[n][256:512].  One warp per node, lane = 8 columns; the messages of a node are one contiguous block.
__global__ void __launch_bounds__(256)
segment_reduce_kernel(const __nv_bfloat16* __restrict__ msg, const float* __restrict__ att, const int* __restrict__ row_ptr,
                      int n_nodes, __nv_bfloat16* __restrict__ hcat) {
    const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();                               // the node GEMM that follows may set up under this kernel's tail
    if (node >= n_nodes) return;
    const int e0 = row_ptr[node], e1 = row_ptr[node + 1];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const uint4* mp = reinterpret_cast<const uint4*>(msg) + lane;            // row stride 32 uint4
    auto ld_stream = [](const uint4* p) {                                     // read-once data: do not pollute L1
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        return v;
    };
    auto fma_row = [&](const uint4& v, float a) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[2 * i] = fmaf(a, __uint_as_float(w[i] << 16), acc[2 * i]);
            acc[2 * i + 1] = fmaf(a, __uint_as_float(w[i] & 0xffff0000u), acc[2 * i + 1]);
        }
    };
    int e = e0;
    for (; e + 8 <= e1; e += 8) {
        uint4 v[8];
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = ld_stream(mp + (size_t)(e + k) * 32);
            a[k] = __ldg(att + e + k);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) fma_row(v[k], a[k]);
    }
    if (e + 4 <= e1) {
        uint4 v[4];
        float a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = ld_stream(mp + (size_t)(e + k) * 32);
            a[k] = __ldg(att + e + k);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) fma_row(v[k], a[k]);
        e += 4;
    }
    for (; e < e1; ++e) fma_row(ld_stream(mp + (size_t)e * 32), __ldg(att + e));
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(hcat + (size_t)node * 512 + 256 + 8 * lane) = o;
}

}  // namespace dndm
