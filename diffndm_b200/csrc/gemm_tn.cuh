// Node-level dense GEMM on the 5th-gen tensor cores:
//   C[M, Nout] = epilogue( A[M, K] (bf16, K-major, row stride lda)  x  W[Nout, K]^T (bf16, K-major) )
// TMA (SWIZZLE_128B) -> 3-stage smem ring -> tcgen05.mma (M=128, N=128, K=16) -> TMEM fp32 -> registers.
// One 128x128 output tile per CTA, 128 threads, 2 CTAs/SM so one CTA's epilogue hides under another's
// mainloop.  Used for: node projections (edge / coord / cross first layers hoisted to nodes), node MLP.
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 128;
constexpr int GEMM_STAGE_BYTES = (GEMM_BM + GEMM_BN) * GEMM_BK * 2;              // 32 KiB
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

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

__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               int M, int K, int a_col0, GemmEpilogue ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + GEMM_STAGES;
    uint64_t* acc_bar = empty_bar + GEMM_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    const int num_k = K / GEMM_BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<GEMM_BN>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % GEMM_STAGES;
                const uint32_t ph = (kb / GEMM_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sa = smem + s * GEMM_STAGE_BYTES;
                uint8_t* sb = sa + GEMM_BM * GEMM_BK * 2;
                mbar_arrive_expect_tx(&full_bar[s], GEMM_STAGE_BYTES);
                tma_load_2d(sa, &tmap_a, &full_bar[s], a_col0 + kb * GEMM_BK, m_blk * GEMM_BM);
                tma_load_2d(sb, &tmap_w, &full_bar[s], kb * GEMM_BK, n_blk * GEMM_BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BM, GEMM_BN);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % GEMM_STAGES;
                const uint32_t ph = (kb / GEMM_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + s * GEMM_STAGE_BYTES);
                const uint32_t sb = sa + GEMM_BM * GEMM_BK * 2;
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    const uint64_t da = make_kmajor_sw128_desc(sa + k * 32);
                    const uint64_t db = make_kmajor_sw128_desc(sb + k * 32);
                    umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[s]);       // frees the smem stage once these MMAs retire
            }
            umma_commit(acc_bar);                  // accumulator complete
        }
        __syncwarp();
    }

    // ---- epilogue: all 4 warps; warp w owns TMEM lanes [32w, 32w+32) ----
    mbar_wait(acc_bar, 0);
    tc_fence_after_sync();
    const int row_in_tile = warp * 32 + lane;
    const long grow = (long)m_blk * GEMM_BM + row_in_tile;
    const bool row_ok = grow < M;
#pragma unroll 1
    for (int c = 0; c < GEMM_BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
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
    __syncthreads();
    if (warp == 1) tmem_dealloc<GEMM_BN>(tmem_base);
}

}  // namespace dndm
