// Node-level dense GEMM on the 5th-gen tensor cores:
//   C[M, Nout] = epilogue( A[M, K] (bf16, K-major, row stride lda)  x  W[Nout, K]^T (bf16, K-major) )
//
// Persistent, warp-specialised (192 threads, 1 CTA / SM, grid = min(tiles, #SMs)):
//   warp 0    TMA producer : 4-stage ring of (A 128x64, W 128x64) bf16 tiles, SWIZZLE_128B
//   warp 1    MMA issuer   : tcgen05.mma M128 N128 K16 into one of two TMEM accumulators (2 x 128 columns)
//   warps 2-5 epilogue     : tcgen05.ld -> bias / SiLU / residual -> fp32 and/or bf16 stores
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
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 256 /*barriers*/;

struct GemmEpilogue {
    const float* bias;        // [Nout] or nullptr
    int act;                  // 0 none, 1 SiLU
    const float* residual;    // [M, ldr] fp32 or nullptr (may alias out_f32)
    int ldr;
    float* out_f32;           // [M, ldc] or nullptr
    int ldc;
    __nv_bfloat16* out_bf16;  // [M, ldcb] or nullptr
    int ldcb;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               int M, int K, int a_col0, int n_blk0, int n_tiles, GemmEpilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
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
        // ---- epilogue warps 2..5: warp w owns TMEM lanes [32 (w%4), +32) ----
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles + n_blk0;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after_sync();
            const long grow = (long)m_blk * GEMM_BM + q * 32 + lane;
            const bool row_ok = grow < M;
#pragma unroll 1
            for (int c = 0; c < GEMM_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + buf * GEMM_BN + ((uint32_t)(q * 32) << 16) + c * 32, v);
                tmem_ld_wait();
                if (row_ok) {
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
                    if (ep.residual) {
                        const float* rp = ep.residual + grow * ep.ldr + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 r = *reinterpret_cast<const float4*>(rp + j);
                            f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
                        }
                    }
                    if (ep.out_f32) {
                        float* op = ep.out_f32 + grow * ep.ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    }
                    if (ep.out_bf16) {
                        __nv_bfloat16* op = ep.out_bf16 + grow * ep.ldcb + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 p;
                            p.x = pack_bf16x2(f[j], f[j + 1]);
                            p.y = pack_bf16x2(f[j + 2], f[j + 3]);
                            p.z = pack_bf16x2(f[j + 4], f[j + 5]);
                            p.w = pack_bf16x2(f[j + 6], f[j + 7]);
                            *reinterpret_cast<uint4*>(op + j) = p;
                        }
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<2 * GEMM_BN>(tmem_base);
}

}  // namespace dndm
