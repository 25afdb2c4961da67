// Node-level dense GEMM on the 5th-gen tensor cores:
//   C[M, Nout] = epilogue( A[M, K] (bf16, K-major, row stride lda)  x  W[Nout, K]^T (bf16, K-major) )
//
// Persistent, warp-specialised (192 threads, 1 CTA / SM, grid = min(tiles, #SMs)):
//   warp 0    TMA producer : 4-stage ring of (A 128x64, W 128x64) bf16 tiles, SWIZZLE_128B
//   warp 1    MMA issuer   : tcgen05.mma M128 N128 K16 into one of two TMEM accumulators (2 x 128 columns)
//   warps 2-5 epilogue     : tcgen05.ld -> bias / SiLU / residual -> swizzled smem slab (32 rows per warp) -> TMA store
//                            (fp32 and/or bf16 destination); each warp owns its slab, no block-level barrier
// Tiles are walked m-major (tile = m * n_tiles + n) so CTAs working at the same time share the A rows in L2, and
// the epilogue of tile i overlaps the loads and MMAs of tile i+1.
// Used for: node projections (edge / coord / cross first layers hoisted to nodes) and the node MLP.
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_STAGE_BYTES = (GEMM_BM + GEMM_BN) * GEMM_BK * 2;              // 32 KiB
constexpr int GEMM_OUT_BF16_BYTES = 32 * GEMM_BN * 2;    //  8 KiB per epilogue warp: 2 boxes [32 rows][64 bf16]
constexpr int GEMM_OUT_F32_BYTES = 32 * GEMM_BN * 4;     // 16 KiB per epilogue warp: 4 boxes [32 rows][32 f32]
constexpr int GEMM_OUT_BYTES = 4 * (GEMM_OUT_BF16_BYTES + GEMM_OUT_F32_BYTES);
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_OUT_BYTES + 256 /*barriers*/;

struct GemmEpilogue {
    const float* bias;        // [Nout] or nullptr
    int act;                  // 0 none, 1 SiLU
    const float* residual;    // [M, ldr] fp32 or nullptr (may be the fp32 destination: each element is read, then written)
    int ldr;
    int has_f32;              // store fp32 through tmap_o32 at column offset col0_f32
    int col0_f32;
    int has_bf16;             // store bf16 through tmap_o16 at column offset col0_bf16
    int col0_bf16;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_o16, const __grid_constant__ CUtensorMap tmap_o32,
               int M, int K, int a_col0, int n_blk0, int n_tiles, GemmEpilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sOut = smem + GEMM_STAGES * GEMM_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_OUT_BYTES);
    uint64_t* empty_bar = full_bar + GEMM_STAGES;
    uint64_t* acc_full = empty_bar + GEMM_STAGES;     // [2]
    uint64_t* acc_empty = acc_full + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = K / GEMM_BK;
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    const int total = m_tiles * n_tiles;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<2 * GEMM_BN>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int kq = 0;                                            // running k-block counter -> ring slot / phase
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles + n_blk0;
                for (int kb = 0; kb < num_k; ++kb, ++kq) {
                    const int s = kq % GEMM_STAGES;
                    const uint32_t ph = (kq / GEMM_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * GEMM_STAGE_BYTES;
                    uint8_t* sb = sa + GEMM_BM * GEMM_BK * 2;
                    mbar_arrive_expect_tx(&full_bar[s], GEMM_STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, &full_bar[s], a_col0 + kb * GEMM_BK, m_blk * GEMM_BM);
                    tma_load_2d(sb, &tmap_w, &full_bar[s], kb * GEMM_BK, n_blk * GEMM_BN);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BM, GEMM_BN);
            int kq = 0, it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait(&acc_empty[buf], ((it - 2) >> 1) & 1);    // epilogue drained this accumulator
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * GEMM_BN;
                for (int kb = 0; kb < num_k; ++kb, ++kq) {
                    const int s = kq % GEMM_STAGES;
                    const uint32_t ph = (kq / GEMM_STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + s * GEMM_STAGE_BYTES);
                    const uint32_t sb = sa + GEMM_BM * GEMM_BK * 2;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        umma_bf16(d_tmem, make_kmajor_sw128_desc(sa + k * 32), make_kmajor_sw128_desc(sb + k * 32), idesc,
                                  (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);       // frees the smem stage once these MMAs retire
                }
                umma_commit(&acc_full[buf]);          // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps 2..5: warp w owns TMEM lanes [32 (w%4), +32) and its own staging slabs ----
        const int q = warp & 3;
        uint8_t* s16 = sOut + q * GEMM_OUT_BF16_BYTES;                           // 2 boxes [32][64 bf16], SW128
        uint8_t* s32 = sOut + 4 * GEMM_OUT_BF16_BYTES + q * GEMM_OUT_F32_BYTES;  // 4 boxes [32][32 f32],  SW128
        const uint32_t sw = lane & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles + n_blk0;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after_sync();
            const int row0 = m_blk * GEMM_BM + q * 32;
            const long grow = (long)row0 + lane;
            const bool row_ok = grow < M;
            if (it > 0) {                               // the previous tile's TMA stores have read the slabs
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            }
#pragma unroll 1
            for (int c = 0; c < GEMM_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + buf * GEMM_BN + ((uint32_t)(q * 32) << 16) + c * 32, v);
                tmem_ld_wait();
                const int col0 = n_blk * GEMM_BN + c * 32;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (ep.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                        f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                    }
                }
                if (ep.act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
                }
                if (ep.residual && row_ok) {
                    const float* rp = ep.residual + grow * ep.ldr + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 r = *reinterpret_cast<const float4*>(rp + j);
                        f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
                    }
                }
                if (ep.has_f32) {                       // box c: row `lane`, 8 units of 16 B, unit u stored at u ^ (lane & 7)
                    uint8_t* rowp = s32 + c * 4096 + lane * 128;
#pragma unroll
                    for (int uu = 0; uu < 8; ++uu)
                        *reinterpret_cast<float4*>(rowp + ((uu ^ sw) << 4)) =
                            make_float4(f[4 * uu], f[4 * uu + 1], f[4 * uu + 2], f[4 * uu + 3]);
                }
                if (ep.has_bf16) {                      // box c/2: row `lane`, units (c%2)*4 .. +3
                    uint8_t* rowp = s16 + (c >> 1) * 4096 + lane * 128;
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        uint4 pk;
                        pk.x = pack_bf16x2(f[8 * uu], f[8 * uu + 1]);     pk.y = pack_bf16x2(f[8 * uu + 2], f[8 * uu + 3]);
                        pk.z = pack_bf16x2(f[8 * uu + 4], f[8 * uu + 5]); pk.w = pack_bf16x2(f[8 * uu + 6], f[8 * uu + 7]);
                        *reinterpret_cast<uint4*>(rowp + ((((c & 1) * 4 + uu) ^ sw) << 4)) = pk;
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[buf]);               // accumulator drained
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < M) {
                const int ncol = n_blk * GEMM_BN;
                if (ep.has_f32) {
#pragma unroll
                    for (int bx = 0; bx < 4; ++bx) tma_store_2d(&tmap_o32, s32 + bx * 4096, ep.col0_f32 + ncol + bx * 32, row0);
                }
                if (ep.has_bf16) {
#pragma unroll
                    for (int bx = 0; bx < 2; ++bx) tma_store_2d(&tmap_o16, s16 + bx * 4096, ep.col0_bf16 + ncol + bx * 64, row0);
                }
                tma_store_commit();
            }
        }
        if (lane == 0) tma_store_wait_all();            // global writes complete before the kernel ends
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<2 * GEMM_BN>(tmem_base);
}

}  // namespace dndm
