// Radius graph of EGNNDynamics.get_edges (dynamics.py:169-187) as receiver-sorted CSR.
//
// Node order is the reference's: all ligand atoms (sample-major), then all pocket atoms.  Row i lists, in
// ascending global index, the same-sample ligand atoms passing the ligand rule followed by the same-sample
// pocket atoms passing the pocket / interaction rule -- exactly the order torch.where() yields on the block
// adjacency, self loops included.  Distance decisions use fp32 direct differences with one rounding per
// operation ((dx*dx + dy*dy) + dz*dz <= cutoff^2), the arithmetic the oracle pins (oracle/egnn_oracle.py:pair_d2).
//
// v1 scans all same-sample candidates per row (warp per row, ballot compaction keeps the order for free);
// a sample is at most ~750 atoms so this is ~25 candidate chunks per row.
#pragma once
#include "common.cuh"

namespace dndm {

struct GraphParams {
    const float* x;          // [N,3]  (ligand rows first)
    const int* lig_ptr;      // [B+1]
    const int* pok_ptr;      // [B+1]
    const int* node_sample;  // [N]
    int n_lig, n_nodes;
    float cut2_l, cut2_p, cut2_i;   // squared cutoffs; negative = no cutoff (fully connected within sample)
};

DNDM_DEVICE float dist2_rn(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// One warp per row.  kFill=false: deg[i] = number of neighbours.  kFill=true: write col / erow / r0.
template <bool kFill>
__global__ void __launch_bounds__(256)
graph_rows_kernel(GraphParams p, int* deg, const int* row_ptr, int* ecol, int* erow, float* r0, int max_edges) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= p.n_nodes) return;
    const int i = warp;
    const int b = p.node_sample[i];
    const bool i_lig = i < p.n_lig;
    const float xi = p.x[3 * i], yi = p.x[3 * i + 1], zi = p.x[3 * i + 2];
    int out = kFill ? row_ptr[i] : 0;
    int count = 0;
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
        const int beg = part == 0 ? p.lig_ptr[b] : p.n_lig + p.pok_ptr[b];
        const int end = part == 0 ? p.lig_ptr[b + 1] : p.n_lig + p.pok_ptr[b + 1];
        const float cut2 = part == 0 ? (i_lig ? p.cut2_l : p.cut2_i) : (i_lig ? p.cut2_i : p.cut2_p);
        for (int j0 = beg; j0 < end; j0 += 32) {
            const int j = j0 + lane;
            bool ok = false;
            float d2 = 0.f;
            if (j < end) {
                d2 = dist2_rn(xi, yi, zi, p.x[3 * j], p.x[3 * j + 1], p.x[3 * j + 2]);
                ok = (cut2 < 0.f) || (d2 <= cut2);
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (kFill) {
                if (ok) {
                    const int pos = out + __popc(m & ((1u << lane) - 1u));
                    if (pos < max_edges) {
                        ecol[pos] = j;
                        erow[pos] = i;
                        r0[pos] = d2;
                    }
                }
                out += __popc(m);
            } else {
                count += __popc(m);
            }
        }
    }
    if (!kFill && lane == 0) deg[i] = count;
}

// Exclusive scan of deg[0..n) -> row_ptr[0..n] in three small launches (reduce / scan of block sums / rescan):
//   blocks of SCAN_ELEMS elements with coalesced loads.  The middle kernel also publishes scalars[0] = E,
//   scalars[1] = E_lig = row_ptr[n_lig] (set by the last kernel) and raises flag bit 2 when E exceeds capacity.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ELEMS = 2048;          // per block, 8 per thread

DNDM_DEVICE int block_reduce_sum(int v, int* smem_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) smem_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += smem_warp[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums_kernel(const int* __restrict__ deg, int n, int* __restrict__ block_sums) {
    __shared__ int sw[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_ELEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ELEMS / SCAN_THREADS; ++k) {
        const int i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += deg[i];
    }
    const int tot = block_reduce_sum(s, sw);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
scan_offsets_kernel(int* __restrict__ block_sums, int n_blocks, int n, int n_lig, int max_edges, int* __restrict__ row_ptr,
                    int* __restrict__ scalars, unsigned* __restrict__ flags, int slot_total, int slot_lig) {
    // n_blocks <= 1024: one thread per block sum, warp-shuffle scan
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = tid < n_blocks ? block_sums[tid] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = warp_tot[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;
        if (lane == 31) {
            const int E = wi;
            row_ptr[n] = E;
            scalars[slot_total] = min(E, max_edges);
            if (n_lig >= n && slot_lig >= 0) scalars[slot_lig] = min(E, max_edges);   // no pocket rows
            if (E > max_edges) atomicOr(flags, 4u);
        }
    }
    __syncthreads();
    if (tid < n_blocks) block_sums[tid] = warp_tot[warp] + incl - v;     // exclusive offset of each block
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const int* __restrict__ deg, const int* __restrict__ block_offs, int n, int n_lig, int max_edges,
                  int* __restrict__ row_ptr, int* __restrict__ scalars, int slot_lig) {
    // thread t owns 8 consecutive elements; block-level exclusive scan of the per-thread sums
    __shared__ int sw[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * SCAN_ELEMS + tid * 8;
    int d[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        d[k] = (i0 + k < n) ? deg[i0 + k] : 0;
        s += d[k];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sw[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += sw[w];
    int run = block_offs[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (i0 + k < n) {
            row_ptr[i0 + k] = run;
            if (i0 + k == n_lig && slot_lig >= 0) scalars[slot_lig] = min(run, max_edges);      // E_lig = row_ptr[n_lig]
        }
        run += d[k];
    }
}

// ------------------------------------------------------------------------------------------------
// Exact dead-work elimination for the LAST block when the caller discards the pocket output (every conditional
// sampler call site does: `eps, _ = dynamics(...)`): after the last block only ligand rows of h are decoded, and the
// last coordinate update reads h of the ligand atoms and of the pocket atoms that send to a ligand atom.  So the last
// GCL only has to aggregate for receivers in  L u P1,  P1 = {pocket j : exists ligand-row edge (i, j)}.
//   mark_active_kernel : deg_act[n] = deg[n] for n in L u P1, else 0
//   (scan)             : rp_act = exclusive scan of deg_act; scalars[2] = number of kept edges
//   compact_edges_kernel: copies the kept rows of (erow, ecol, r0) to the compacted arrays
// ------------------------------------------------------------------------------------------------
__global__ void mark_active_kernel(const int* __restrict__ ecol, const int* __restrict__ deg, const int* __restrict__ scalars,
                                   int n_lig, int n_nodes, int* __restrict__ deg_act, int pass) {
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (pass == 0) {                      // ligand rows keep their degree, pocket rows start at 0
        for (int n = t; n < n_nodes; n += stride) deg_act[n] = n < n_lig ? deg[n] : 0;
    } else {                              // every pocket sender of a ligand row becomes active (idempotent writes)
        const int e_lig = scalars[1];
        for (int e = t; e < e_lig; e += stride) {
            const int c = ecol[e];
            if (c >= n_lig) deg_act[c] = deg[c];
        }
    }
}

__global__ void __launch_bounds__(256)
compact_edges_kernel(const int* __restrict__ row_ptr, const int* __restrict__ rp_act, const int* __restrict__ erow,
                     const int* __restrict__ ecol, const float* __restrict__ r0, int n_nodes, int* __restrict__ erow_c,
                     int* __restrict__ ecol_c, float* __restrict__ r0_c) {
    const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= n_nodes) return;
    const int a0 = rp_act[node], a1 = rp_act[node + 1];
    const int s0 = row_ptr[node];
    for (int k = lane; k < a1 - a0; k += 32) {
        erow_c[a0 + k] = erow[s0 + k];
        ecol_c[a0 + k] = ecol[s0 + k];
        r0_c[a0 + k] = r0[s0 + k];
    }
}

// Per-sample start offsets from a sorted int64 batch mask (utils.py:145-153 layout) + int32 sample id per node.
__global__ void mask_to_ptr_kernel(const long long* mask, int n, int n_samples, int* ptr, int* node_sample) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_samples) {
        int lo = 0, hi = n;      // first index with mask[idx] >= i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (mask[mid] < (long long)i) lo = mid + 1; else hi = mid;
        }
        ptr[i] = lo;
    }
    if (i < n) node_sample[i] = (int)mask[i];
}

}  // namespace dndm
