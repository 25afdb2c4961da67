/* diffndm_b200 -- C ABI of the B200-native DiffNDM / DiffSBDD denoiser engine.
 *
 * Drop-in boundary for ONE path of the reference: the batched EGNN denoising forward and the elementwise
 * sampler step around it.  The reference is pure Python (no FFI of its own); every entry point below names
 * the reference function it replaces (paths relative to the reference repository root).  A maintainer binds
 * these with ctypes (diffndm_b200/engine.py is that binding).
 *
 * Conventions: plain pointers and sizes, no torch types.  Pointers marked DEVICE are CUDA device pointers owned
 * by the caller (torch tensors' data_ptr()); HOST pointers are host memory.  `stream` is a cudaStream_t passed
 * as void* (0 = legacy default stream).  Every function returns 0 on success, a negative DNDM_E* code otherwise;
 * no exceptions cross the ABI.  Device-side conditions (NaN, COM drift, capacity) are reported through a
 * sticky flag word read with dndm_read_flags().  One engine per device per process; calls on one engine must
 * not overlap (the reference is single-threaded).
 */
#ifndef DIFFNDM_B200_H
#define DIFFNDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNDM_OK 0
#define DNDM_EINVAL (-1)    /* bad argument / unsupported configuration */
#define DNDM_ECUDA (-2)     /* CUDA runtime or driver error (see dndm_last_error) */
#define DNDM_ECAPACITY (-3) /* batch exceeds the engine's max_nodes / max_samples */
#define DNDM_EWEIGHTS (-4)  /* missing or mis-shaped weight */

/* sticky device flags (dndm_read_flags) */
#define DNDM_FLAG_NAN 1u          /* NaN in the predicted velocity  -> ValueError, dynamics.py:155-159 */
#define DNDM_FLAG_COM_DRIFT 2u    /* ligand COM of z_t not ~0        -> AssertionError, en_diffusion.py:930-935 */
#define DNDM_FLAG_MOL_TOO_LARGE 8u /* dndm_bond_orders: a molecule has more than 256 atoms (its stats are -1) */
#define DNDM_FLAG_EDGE_OVERFLOW 4u /* more edges than max_edges (results invalid) */

typedef struct DndmEngine DndmEngine;

/* Hyper-parameters of EGNNDynamics (dynamics.py:11-85; configs/crossdock_fullatom_cond.yml:36-51). */
typedef struct DndmConfig {
    int32_t atom_nf;              /* ligand feature width (10) */
    int32_t residue_nf;           /* pocket feature width (10 in full-atom mode) */
    int32_t joint_nf;             /* 128 */
    int32_t hidden_nf;            /* 256 (the only width compiled in this round) */
    int32_t n_layers;             /* 6 EquivariantBlocks, inv_sublayers = 1 */
    float edge_cutoff_ligand;     /* < 0: none (fully connected), dynamics.py:174-175 */
    float edge_cutoff_pocket;     /* 5.0 */
    float edge_cutoff_interaction;/* 5.0 */
    float norm_constant;          /* 1 */
    float normalization_factor;   /* 100 */
    float coords_range;           /* 15 (egnn_new.py:218 passes the undivided value) */
    int32_t max_nodes;            /* capacity: ligand + pocket atoms per call */
    int32_t max_edges;            /* capacity: directed edges incl. self loops per call */
    int32_t max_samples;          /* capacity: batch size per call */
    int32_t device;               /* CUDA device ordinal */
} DndmConfig;

/* One parameter of the reference state_dict `ddpm.dynamics.*`, fp32 HOST memory,
 * row-major [rows, cols] (bias: rows = n, cols = 1). */
typedef struct DndmWeight {
    const char* name;
    const float* data;
    int32_t rows;
    int32_t cols;
} DndmWeight;

/* Library / build information. */
const char* dndm_version(void);
const char* dndm_last_error(void);
/* Number of kernels of this library launched (or captured into a CUDA graph) by this process so far. */
int64_t dndm_launch_count(void);

/* Lifetime.  Replaces EGNNDynamics.__init__ (dynamics.py:11-85): allocates the HBM workspace once. */
int dndm_engine_create(const DndmConfig* cfg, DndmEngine** out);
void dndm_engine_destroy(DndmEngine* e);

/* Replaces nn.Module.load_state_dict for `ddpm.dynamics.*`: packs private bf16 / fp32 device copies
 * (first-layer edge weights split per endpoint, encoder/decoder linears pre-composed).  Must be called again
 * whenever the caller's parameters change.
 * The table is in the shapes of DndmConfig (hidden_nf must be 256, the compiled width; first-layer edge weights
 * [256, 2 * 256 + 2]).  A narrower reference network (hidden_nf 192 / 128) is loaded zero-padded, and the edge-type
 * embedding of dynamics.py:118-127 folded into three optional entries per block,
 *   "egnn.e_block_<i>.{gcl_0.edge_mlp, gcl_equiv.coord_mlp, gcl_equiv.cross_product_mlp}.0.edge_type_bias"  [3, 256]
 *   = W1[:, 2H+2:] E[type]  (type 0 ligand-pocket, 1 ligand-ligand, 2 pocket-pocket),
 * given for every edge MLP of every block or for none; both rewrites are exact and are what the Python binding
 * (diffndm_b200/weights.py: engine_table) does with the reference's state_dict. */
int dndm_engine_load_weights(DndmEngine* e, const DndmWeight* weights, int32_t n_weights);

/* EGNNDynamics.forward (dynamics.py:87-167), update_pocket_coords = False, condition_time = True.
 *   xh_lig    DEVICE [n_lig, 3+atom_nf]      fp32 row-major (coordinates, then features)
 *   xh_pocket DEVICE [n_pocket, 3+residue_nf]
 *   t         DEVICE [t_len]  t_len == n_samples (per-sample time) or 1 (shared)
 *   lig_mask / pocket_mask DEVICE int64, non-decreasing sample ids 0..n_samples-1 (utils.py:145-153)
 *   out_lig   DEVICE [n_lig, 3+atom_nf]      (velocity, decoded features)
 *   out_pocket DEVICE [n_pocket, 3+residue_nf] or NULL (every conditional caller discards it) */
int dndm_egnn_forward(DndmEngine* e, const float* xh_lig, const float* xh_pocket, const float* t, int32_t t_len,
                      const int64_t* lig_mask, const int64_t* pocket_mask, int32_t n_lig, int32_t n_pocket,
                      int32_t n_samples, float* out_lig, float* out_pocket, void* stream);

/* EGNNDynamics.get_edges (dynamics.py:169-187) on its own: builds the receiver-sorted CSR inside the engine and
 * optionally copies it out.  row_ptr DEVICE int32 [n_lig+n_pocket+1] or NULL; col DEVICE int32 [edge_capacity] or
 * NULL.  The edge count is written to *n_edges_host after a stream synchronisation. */
int dndm_radius_graph(DndmEngine* e, const float* xh_lig, const float* xh_pocket, const int64_t* lig_mask,
                      const int64_t* pocket_mask, int32_t n_lig, int32_t n_pocket, int32_t n_samples,
                      int32_t* row_ptr, int32_t* col, int32_t edge_capacity, int32_t* n_edges_host, void* stream);

/* Elementwise core of ConditionalDDPM.sample_p_zs_given_zt (conditional_model.py:524-533) fused with
 * sample_normal_zero_com / remove_mean_batch (:165-186, :1793-1801) and the SPSA update (:801-812):
 *   z_out = z_in * coef[b][0] - coef[b][1] * eps + coef[b][2] * noise  (+ lambda * grad on coordinates)
 *   then the per-sample ligand COM is removed from z_out[:, :3] and from the pocket coordinates.
 * coef DEVICE [n_samples, 3]; eps / noise DEVICE [n_lig, 3+atom_nf] (eps may be NULL when coef[.][1] == 0);
 * grad DEVICE [n_lig, 3] or NULL.  In-place operation (z_out == z_in, pocket_out == pocket_in) is allowed.
 * check_input_com != 0 reproduces assert_mean_zero_with_mask on the INPUT state (conditional_model.py:535, only the
 * true reverse step has it): |sum of a sample's ligand coordinates| >= 1e-2 * max|coordinate| raises DNDM_FLAG_COM_DRIFT.
 * The x0 head, the prior draw, q(z_s|x), the re-noising move and the SPSA update pass 0 (their inputs need not be
 * COM-free). */
int dndm_sampler_step(DndmEngine* e, const float* z_in, const float* eps, const float* noise,
                      const float* xh_pocket_in, const float* coef, const float* grad, float lambda,
                      const int64_t* lig_mask, const int64_t* pocket_mask, int32_t n_lig, int32_t n_pocket,
                      int32_t n_samples, float* z_out, float* xh_pocket_out, int32_t check_input_com, void* stream);

/* Reads and clears the sticky flag word (synchronises `stream`). */
int dndm_read_flags(DndmEngine* e, uint32_t* flags_host, void* stream);

/* Introspection used by tests and the benchmark (not part of the reference surface):
 * copies an internal buffer of the LAST forward to a DEVICE destination.
 *   what: 0 = h [n_nodes,256] fp32, 1 = x [n_nodes,3] fp32 (final coordinates), 2 = row_ptr int32 [n_nodes+1],
 *         3 = col int32 [E], 4 = scalars int32 [8] (E, E_ligand_rows, E of the last block, of the block before it, of the one before that, reserved),
 *         8 = h_0 [n_nodes,256] fp32, the encoder + embedding output of the last forward (kept only while dndm_set_trace is on),
 *         7 = state of the pocket-pocket candidate lists int32 [8] (built, stale, n_lig, n_pocket, n_samples of the call they
 *             describe, list length the last call used for its first pocket row or -1 if it scanned everything)
 * Returns the number of bytes copied (>= 0) or a negative error. */
int64_t dndm_debug_copy(DndmEngine* e, int32_t what, void* dst, int64_t dst_bytes, void* stream);

/* Per-section device timing for the benchmark's roofline line: when on, every forward brackets its sections with
 * CUDA events on the caller's stream (do not enable under graph capture).  dndm_get_profile synchronises, sums the
 * elapsed milliseconds and the number of timed sections per category since the last call, and resets.
 * Categories: 0 fused GCL edge kernel, 1 fused coordinate-head edge kernel, 2 node GEMMs, 3 radius graph,
 * 4 remaining node kernels. */
int dndm_set_profile(DndmEngine* e, int32_t on);

/* Bond perception for candidate pre-filtering -- replaces the numeric part of make_mol_edm / get_bond_order_batch
 * (reference analysis/molecule_builder.py:30-55, 100-113).  All pointers DEVICE.
 *   x [n_atoms, ld_x] fp32 coordinates in Angstrom (first 3 columns), atom_type [n_atoms] int64 decoder indices,
 *   mol_mask [n_atoms] int64 sorted molecule ids 0..n_mols-1, bonds1/2/3 [n_types, n_types] fp32 lengths in pm
 *   (dataset_info['bonds1'] ...), margins in pm (constants.py:17), allowed_valence [n_types] int32 or NULL.
 * Outputs: e_out (or NULL): per molecule a dense row-major int8 [n_b, n_b] block holding the DIRECTED lower-triangular
 *   bond orders (0/1/2/3), blocks concatenated in molecule order (sum n_b^2 bytes; flag bit 2 if e_capacity is too small);
 *   valence_out [n_atoms] int32 (sum of the symmetrised orders); mol_stats [n_mols, 4] int32 =
 *   (n_bonds, n_components of the bond graph, size of the largest component, atoms exceeding allowed_valence). */
int dndm_bond_orders(DndmEngine* e, const float* x, int32_t ld_x, const int64_t* atom_type, const int64_t* mol_mask,
                     int32_t n_atoms, int32_t n_mols, const float* bonds1, const float* bonds2, const float* bonds3,
                     int32_t n_types, float margin1, float margin2, float margin3, const int32_t* allowed_valence,
                     int8_t* e_out, int64_t e_capacity, int32_t* valence_out, int32_t* mol_stats, void* stream);

/* Static batch layout.  Every entry point derives the per-sample offsets from the int64 masks (two small launches).
 * While `on` != 0 the caller promises that the CONTENTS of the mask buffers do not change between calls that pass the
 * same pointers and sizes (the sampling loop: utils.py:145-153 masks are built once per trajectory), and the engine
 * skips that derivation when pointers and sizes repeat.  Switching the option (either way) drops the cached layout. */
int dndm_set_static_masks(DndmEngine* e, int32_t on);
int dndm_get_profile(DndmEngine* e, double* ms_per_category, int32_t* launches_per_category, int32_t n_cat);

/* Trace hook for parity tests: when set (DEVICE, [n_layers, max_trace_nodes, 256] fp32 and
 * [n_layers, max_trace_nodes, 3] fp32, either may be NULL), every forward stores h and x after each block. */
int dndm_set_trace(DndmEngine* e, float* h_trace, float* x_trace, int32_t max_trace_nodes);

/* Self-test of the weight-resident tcgen05 node GEMM (gemm_pair_kernel) with the launch geometry of the forward:
 *   C[M, N] = epilogue( A[M, K] (bf16) x W[N, K]^T (bf16) + bias [+ residual] ),  optional SiLU,
 * (K, bn) in {(256, 256), (512, 128), (256, 128)}: bn output columns per resident weight group, N % bn == 0.  The last
 * n_tail_groups column groups are computed for the first m_tail rows only (the ligand-row-only projections of the merged
 * launch); their other outputs are left untouched.  residual: fp32 [M, 256] or NULL (N == 256).  Outputs: out_f32 fp32
 * [M, N] and / or out_bf16 bf16 [ceil(M / 128) * 128, N] (TMA store), either may be NULL but not both; a_bf16 must be
 * readable for ceil(M / 128) * 128 rows.  All pointers DEVICE. */
int dndm_test_gemm(const void* a_bf16, const void* w_bf16, const float* bias, const float* residual, int32_t act, int32_t M,
                   int32_t N, int32_t K, int32_t bn, int32_t n_tail_groups, int32_t m_tail, float* out_f32, void* out_bf16,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFNDM_B200_H */
