// Bond perception for candidate pre-filtering: the numeric part of make_mol_edm
// (analysis/molecule_builder.py:100-113) = get_bond_order_batch (:30-55) on all atom pairs of every molecule of a batch.
//   E[i][j] (i > j) = 3 / 2 / 1 / 0 : 100 * |x_i - x_j| < table_k[type_i][type_j] + margin_k, later rules overwrite
// plus what the host-side filters derive from E: per-atom valence, valence violations, bond-graph components.
// One CTA per molecule (<= BOND_MAX_ATOMS atoms, coordinates / types / adjacency bitmask in shared memory).  Distances are
// fp32 with one rounding per operation (sub, mul, add, sqrt, mul) -- the arithmetic oracle/egnn_oracle.py:bond_orders pins.
#pragma once
#include "common.cuh"

namespace dndm {

constexpr int BOND_MAX_ATOMS = 256;
constexpr int BOND_THREADS = 128;

struct BondTables {
    const float* b1;          // [T,T] single-bond lengths (pm), 0 = no such bond
    const float* b2;
    const float* b3;
    const int* allowed;       // [T] maximum valence per atom type, or nullptr
    int n_types;
    float m1, m2, m3;         // margins (pm)
};

// e_off[b] = sum_{k<b} n_k^2 (bytes of the dense int8 blocks before molecule b); e_off[n_mols] = total
__global__ void bond_offsets_kernel(const int* __restrict__ mol_ptr, int n_mols, long long* __restrict__ e_off) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long acc = 0;
    for (int b = 0; b < n_mols; ++b) {
        e_off[b] = acc;
        const long long n = mol_ptr[b + 1] - mol_ptr[b];
        acc += n * n;
    }
    e_off[n_mols] = acc;
}

__global__ void __launch_bounds__(BOND_THREADS)
bond_orders_kernel(const float* __restrict__ x, int ld_x, const long long* __restrict__ atom_type,
                   const int* __restrict__ mol_ptr, BondTables tb, const long long* __restrict__ e_off,
                   signed char* __restrict__ e_out, long long e_capacity, int* __restrict__ valence_out,
                   int* __restrict__ mol_stats, unsigned* __restrict__ flags) {
    __shared__ float sx[BOND_MAX_ATOMS][3];
    __shared__ int stype[BOND_MAX_ATOMS];
    __shared__ int sval[BOND_MAX_ATOMS];
    __shared__ int slabel[BOND_MAX_ATOMS];
    __shared__ int scount[BOND_MAX_ATOMS];
    __shared__ unsigned sadj[BOND_MAX_ATOMS][BOND_MAX_ATOMS / 32];
    __shared__ int s_bonds, s_viol, s_comp, s_largest;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int a0 = mol_ptr[b], n = mol_ptr[b + 1] - a0;
    if (n > BOND_MAX_ATOMS) {                          // not a drug-sized molecule: report instead of truncating
        if (tid == 0) {
            atomicOr(flags, 8u);
            mol_stats[4 * b] = -1; mol_stats[4 * b + 1] = -1; mol_stats[4 * b + 2] = -1; mol_stats[4 * b + 3] = -1;
        }
        return;
    }
    for (int i = tid; i < n; i += BOND_THREADS) {
        sx[i][0] = x[(size_t)(a0 + i) * ld_x];
        sx[i][1] = x[(size_t)(a0 + i) * ld_x + 1];
        sx[i][2] = x[(size_t)(a0 + i) * ld_x + 2];
        int t = (int)atom_type[a0 + i];
        stype[i] = t < 0 ? 0 : (t >= tb.n_types ? tb.n_types - 1 : t);
        sval[i] = 0;
        slabel[i] = i;
        scount[i] = 0;
        for (int w = 0; w < BOND_MAX_ATOMS / 32; ++w) sadj[i][w] = 0u;
    }
    if (tid == 0) { s_bonds = 0; s_viol = 0; s_comp = 0; s_largest = 0; }
    __syncthreads();
    const long long off = e_off[b];
    const bool store = e_out != nullptr && off + (long long)n * n <= e_capacity;
    if (e_out != nullptr && !store && tid == 0) atomicOr(flags, 4u);     // output capacity exceeded
    int my_bonds = 0;
    for (int p = tid; p < n * n; p += BOND_THREADS) {
        const int i = p / n, j = p - i * n;
        int order = 0;
        if (i > j) {                                   // directed lower triangle, molecule_builder.py:111
            const float dx = __fsub_rn(sx[i][0], sx[j][0]), dy = __fsub_rn(sx[i][1], sx[j][1]),
                        dz = __fsub_rn(sx[i][2], sx[j][2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float d = __fmul_rn(100.0f, __fsqrt_rn(d2));
            const int k = stype[i] * tb.n_types + stype[j];
            if (d < __fadd_rn(tb.b1[k], tb.m1)) order = 1;
            if (d < __fadd_rn(tb.b2[k], tb.m2)) order = 2;
            if (d < __fadd_rn(tb.b3[k], tb.m3)) order = 3;
            if (order) {
                ++my_bonds;
                atomicAdd(&sval[i], order);
                atomicAdd(&sval[j], order);
                atomicOr(&sadj[i][j >> 5], 1u << (j & 31));
                atomicOr(&sadj[j][i >> 5], 1u << (i & 31));
            }
        }
        if (store) e_out[off + p] = (signed char)order;
    }
    if (my_bonds) atomicAdd(&s_bonds, my_bonds);
    __syncthreads();
    // bond-graph components by min-label propagation
    for (int iter = 0; iter < n; ++iter) {
        int changed = 0;
        for (int i = tid; i < n; i += BOND_THREADS) {
            int m = slabel[i];
            for (int w = 0; w < (n + 31) / 32; ++w) {
                unsigned bits = sadj[i][w];
                while (bits) {
                    const int j = w * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    m = min(m, slabel[j]);
                }
            }
            if (m < slabel[i]) { slabel[i] = m; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int i = tid; i < n; i += BOND_THREADS) {
        atomicAdd(&scount[slabel[i]], 1);
        valence_out[a0 + i] = sval[i];
        if (tb.allowed && sval[i] > tb.allowed[stype[i]]) atomicAdd(&s_viol, 1);
    }
    __syncthreads();
    for (int i = tid; i < n; i += BOND_THREADS) {
        if (scount[i] > 0) {
            atomicAdd(&s_comp, 1);
            atomicMax(&s_largest, scount[i]);
        }
    }
    __syncthreads();
    if (tid == 0) {
        mol_stats[4 * b] = s_bonds;
        mol_stats[4 * b + 1] = s_comp;
        mol_stats[4 * b + 2] = s_largest;
        mol_stats[4 * b + 3] = s_viol;
    }
}

}  // namespace dndm
