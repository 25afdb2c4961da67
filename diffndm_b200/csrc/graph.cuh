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

// ---- pocket-pocket candidate lists --------------------------------------------------------------------------------
// Inside a trajectory the pocket of a sample is a rigid body that only translates (conditional_model.py:529-533), and
// pocket-pocket pairs are 84 % of the edges: scanning all ~330 same-sample pocket atoms for every pocket row, twice per
// call (count, fill), was ~55 us of a 1.65 ms step and mostly exposed (the graph is on the critical path of block 0).  The
// engine therefore keeps, per pocket row, the ascending list of same-sample pocket atoms within cutoff + PP_MARGIN, built
// from the coordinates of some earlier call, and the row kernels run the EXACT test of the current call -- same arithmetic,
// same order -- on that list only.  Nothing is promised by the caller: every call first verifies on the device that the
// batch layout is the one the lists were built for and that every pocket atom still sits, relative to its sample's first
// pocket atom, within PP_TOL of where it sat then (gather_verify_kernel); if not, this call scans everything as before and
// rebuilds the lists afterwards (pp_rebuild_kernel, off the critical path).  A pair within the cutoff now was within
// cutoff + 2 PP_TOL then, so the lists are a superset and the emitted edge set is bit-identical; the translation rounding of
// a 500-step trajectory moves atoms by ~1e-4 A relative to each other.  All of it is device-side: it replays in a CUDA graph.
constexpr int PP_CAP = 64;             // list capacity per row; a denser row has no list (0xFFFF) and scans everything
constexpr float PP_TOL = 0.02f;        // A
constexpr float PP_MARGIN = 0.1f;      // A  (> 2 PP_TOL + rounding)

struct PocketLists {
    float* canon;            // [n_pocket,3] pocket coordinates the lists were built from
    int* ptr;                // [B+1] pok_ptr of that call
    int* meta;               // [0] built, [1] stale (set by verify), [2] n_lig, [3] n_pocket, [4] n_samples of that call,
                             // [5] list length used for the first pocket row by the last call (-1: scanned everything)
    unsigned short* cand;    // [n_pocket][PP_CAP] candidate offsets within the sample's pocket range, ascending
    unsigned short* cnt;     // [n_pocket]
};

// Gathers the coordinates of all nodes ([ligand ; pocket] rows of the two input tensors) into x [N,3] for the graph branch
// and, in the same pass, checks the candidate lists against this call: thread = node.
__global__ void __launch_bounds__(256)
gather_verify_kernel(const float* __restrict__ xh_lig, const float* __restrict__ xh_pok, int ld_lig, int ld_pok, GraphParams p,
                     PocketLists c, int n_samples, float* __restrict__ x) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_pocket = p.n_nodes - p.n_lig;
    if (t < p.n_nodes) {
        const float* src = t < p.n_lig ? xh_lig + (size_t)t * ld_lig : xh_pok + (size_t)(t - p.n_lig) * ld_pok;
        x[3 * t] = src[0]; x[3 * t + 1] = src[1]; x[3 * t + 2] = src[2];
    }
    if (c.meta == nullptr || c.meta[0] == 0) return;
    bool stale = false;
    if (t == 0) stale = c.meta[2] != p.n_lig || c.meta[3] != n_pocket || c.meta[4] != n_samples;
    if (t <= n_samples) stale |= c.ptr[t] != p.pok_ptr[t];
    if (t < n_pocket) {
        const int f = p.pok_ptr[p.node_sample[p.n_lig + t]];    // the sample's first pocket atom (pocket-local index)
        const float* me = xh_pok + (size_t)t * ld_pok;
        const float* first = xh_pok + (size_t)f * ld_pok;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float d = (me[k] - first[k]) - (c.canon[3 * t + k] - c.canon[3 * f + k]);
            stale |= !(fabsf(d) <= PP_TOL);                      // NaN counts as moved
        }
    }
    if (stale) c.meta[1] = 1;
}

// warp per pocket row; runs only when the lists are missing or stale
__global__ void __launch_bounds__(256)
pp_rebuild_kernel(GraphParams p, PocketLists c, int n_samples) {
    if (c.meta[0] != 0 && c.meta[1] == 0) return;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= n_samples) c.ptr[g] = p.pok_ptr[g];
    const int t = g >> 5, lane = threadIdx.x & 31;
    if (t >= p.n_nodes - p.n_lig) return;
    const int j = p.n_lig + t;
    const int b = p.node_sample[j];
    const int beg = p.pok_ptr[b], end = p.pok_ptr[b + 1];
    const float xi = p.x[3 * j], yi = p.x[3 * j + 1], zi = p.x[3 * j + 2];
    const float rc = sqrtf(fmaxf(p.cut2_p, 0.f)) + PP_MARGIN;
    const float cand2 = rc * rc;
    int count = 0;
    for (int j0 = beg; j0 < end; j0 += 32) {
        const int jl = j0 + lane;
        bool ok = false;
        if (jl < end) {
            const float* q = p.x + 3 * (p.n_lig + jl);
            ok = dist2_rn(xi, yi, zi, q[0], q[1], q[2]) <= cand2;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < PP_CAP) c.cand[(size_t)t * PP_CAP + pos] = (unsigned short)(jl - beg);
        }
        count += __popc(m);
    }
    if (lane == 0) c.cnt[t] = (count > PP_CAP || end - beg > 65535 || p.cut2_p < 0.f) ? 0xFFFFu : (unsigned short)count;
    if (lane < 3) c.canon[3 * t + lane] = p.x[3 * j + lane];
}

__global__ void pp_commit_kernel(PocketLists c, int n_lig, int n_pocket, int n_samples) {
    c.meta[0] = 1; c.meta[1] = 0; c.meta[2] = n_lig; c.meta[3] = n_pocket; c.meta[4] = n_samples;
}

// kFill=false: deg[i] = number of neighbours.  kFill=true: write col / erow / r0.
// A ligand row scans its whole sample (~350 candidates): one warp per row.  A pocket row with a candidate list tests ~23
// ligand atoms + ~20 listed pocket atoms: a whole warp per row spent most of its ~200 instructions on per-row set-up (7.4 M +
// 9.0 M warp instructions per call for 35 k rows), so GR_ROWS pocket rows share a warp -- each row's 8 lanes stride its
// candidates, the warp ballot is cut into the row's byte, and the compaction keeps ascending order exactly as before.
constexpr int GR_ROWS = 4;

template <bool kFill>
__global__ void __launch_bounds__(256)
graph_rows_kernel(GraphParams p, PocketLists c, int* __restrict__ deg, const int* __restrict__ row_ptr, int* __restrict__ ecol,
                  int* __restrict__ erow, float* __restrict__ r0, int max_edges) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const float* __restrict__ x = p.x;
    // warps [0, n_lig): one ligand row each; then GR_ROWS pocket rows per warp
    const bool lig_warp = warp < p.n_lig;
    const int L = lig_warp ? 32 : 32 / GR_ROWS;                      // lanes per row (warp-uniform)
    const int g = lane / L, sl = lane % L;
    const int i_raw = lig_warp ? warp : p.n_lig + (warp - p.n_lig) * GR_ROWS + g;
    if ((lig_warp ? warp : p.n_lig + (warp - p.n_lig) * GR_ROWS) >= p.n_nodes) return;
    const bool row_ok = i_raw < p.n_nodes;
    const int i = row_ok ? i_raw : p.n_nodes - 1;                    // idle groups of the last warp: harmless reads, emit nothing
    const int b = p.node_sample[i];
    const bool i_lig = i < p.n_lig;
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    int out = kFill ? row_ptr[i] : 0;
    int count = 0;
    // pocket row with a valid candidate list: the pocket part of the row is the exact test over the list
    int n_list = -1;
    if (!i_lig && c.meta != nullptr && c.meta[0] == 1 && c.meta[1] == 0 && p.cut2_p >= 0.f) {
        const int n = c.cnt[i - p.n_lig];
        if (n != 0xFFFF) n_list = n;
    }
    if (!kFill && row_ok && i == p.n_lig && sl == 0 && c.meta != nullptr) c.meta[5] = n_list;   // introspection: did this call use the lists
    const unsigned row_bits = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
    // candidates k = 0 .. n - 1 of this row, candidate index -> atom through `atom(k)`, ascending
    auto scan = [&](int n, float cut2, auto atom) {
        const int n_max = __reduce_max_sync(0xffffffffu, row_ok ? n : 0);
#pragma unroll 2
        for (int k0 = 0; k0 < n_max; k0 += L) {
            const int k = k0 + sl;
            bool ok = false;
            float d2 = 0.f;
            int j = 0;
            if (row_ok && k < n) {
                j = atom(k);
                d2 = dist2_rn(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2]);
                ok = (cut2 < 0.f) || (d2 <= cut2);
            }
            const unsigned m = (__ballot_sync(0xffffffffu, ok) >> (g * L)) & row_bits;
            if (kFill) {
                if (ok) {
                    const int pos = out + __popc(m & ((1u << sl) - 1u));
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
    };
    const int lb = p.lig_ptr[b], pb = p.n_lig + p.pok_ptr[b];
    scan(p.lig_ptr[b + 1] - lb, i_lig ? p.cut2_l : p.cut2_i, [&](int k) { return lb + k; });
    // rows with and without a list can share a warp, and `scan` is warp-collective: every lane walks through both pocket
    // scans, a row takes part in one of them (the other has n = 0 for it)
    const unsigned short* __restrict__ lst = c.cand + (size_t)(i_lig ? 0 : i - p.n_lig) * PP_CAP;
    scan(n_list >= 0 ? n_list : 0, p.cut2_p, [&](int k) { return pb + (int)lst[k]; });
    scan(n_list >= 0 ? 0 : p.pok_ptr[b + 1] - p.pok_ptr[b], i_lig ? p.cut2_i : p.cut2_p, [&](int k) { return pb + k; });
    if (!kFill && row_ok && sl == 0) deg[i] = count;
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
// The same argument one block earlier: the second-to-last GCL only has to aggregate for  S u senders(S)  (two hops from
// the ligand: ~0.78 E on the benchmark pockets); three hops already reach ~all of a pocket, so it stops there.
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

// One hop further back: the block BEFORE the last one has to produce h for every receiver the last block reads, i.e. for the
// last block's receivers S and all their senders.  pass 0: deg_out = deg_prev (S keeps its rows); pass 1: every sender of a
// row of S -- the columns of the last block's compacted edge list, scalars[slot_prev] of them -- becomes active.
__global__ void mark_senders_kernel(const int* __restrict__ ecol_prev, const int* __restrict__ deg, const int* __restrict__ deg_prev,
                                    const int* __restrict__ scalars, int slot_prev, int n_nodes, int* __restrict__ deg_out, int pass) {
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (pass == 0) {
        for (int n = t; n < n_nodes; n += stride) deg_out[n] = deg_prev[n];
    } else {
        const int n_e = scalars[slot_prev];
        for (int e = t; e < n_e; e += stride) {
            const int c = ecol_prev[e];
            deg_out[c] = deg[c];
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
