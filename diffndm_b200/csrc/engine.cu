// diffndm_b200 engine: HBM workspace, weight packing, forward orchestration and the C ABI (include/diffndm_b200.h).
#include "../../include/diffndm_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_pair.cuh"
#include "edge_mlp.cuh"
#include "edge_pair.cuh"
#include "graph.cuh"
#include "bonds.cuh"
#include "node_kernels.cuh"

using namespace dndm;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int g_num_sms = 148;
static long long g_launches = 0;   // kernels of this library launched (or captured) so far
#define COUNT_LAUNCH(n) (g_launches += (n))
// Programmatic dependent launch for the weight-resident kernels (edge MLP, node GEMMs): their prologue -- barrier
// init, TMEM allocation, the 64-128 KiB TMA load of the resident weights -- runs under the tail of the previous kernel
// of the stream (see pdl_wait / pdl_trigger in common.cuh).  DNDM_PDL=0 in the environment turns the attribute off.
// DNDM_PP_LISTS=0: every call scans all same-sample pocket atoms for every pocket row (the round-1 path; for A/B)
static bool g_pp_lists = [] { const char* v = getenv("DNDM_PP_LISTS"); return !(v && v[0] == '0'); }();
// DNDM_PRUNE_LEVELS=<n>: how many trailing blocks run on a pruned edge list (1 = the last block only; default 2; maximum 3 --
// three hops from the ligand cover 0.99 E of a 330-atom pocket and 0.85 E of a 600-atom one: measured equal within noise)
static int g_prune_levels = [] { const char* v = getenv("DNDM_PRUNE_LEVELS"); const int n = v ? atoi(v) : 2; return n < 1 ? 1 : n; }();
static bool g_pdl = [] { const char* v = getenv("DNDM_PDL"); return !(v && v[0] == '0'); }();
// GCL producers: all-bf16x2 first-layer pre-activation by default; DNDM_GCL_F32_RADIAL=1 selects the variant with fp32 radial
// terms and activation (measured on the radial stress fixture: same error to two digits, 8 % slower -- see edge_mlp.cuh)
static bool g_bf16_radial = [] { const char* v = getenv("DNDM_GCL_F32_RADIAL"); return !(v && v[0] == '1'); }();
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU_CHECK(x)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (x);                                                                         \
        if (_e != cudaSuccess)                                                                        \
            return set_err(DNDM_ECUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// TMA descriptor creation through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}
static int make_tmap_bf16_box(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                              uint32_t box_rows, CUtensorMapSwizzle swz) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return set_err(DNDM_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(DNDM_ECUDA, "cuTensorMapEncodeTiled (bf16 box) failed with %d", (int)r);
    return DNDM_OK;
}
// bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128-byte swizzle.
static int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return set_err(DNDM_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(DNDM_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// engine state
// ------------------------------------------------------------------------------------------------
struct LayerWeights {
    __nv_bfloat16 *wproj_e, *w2_e, *w2_c, *w2_x, *w3, *w4;             // device bf16
    __nv_bfloat16* wm;                                                  // merged projections [next-edge P|Q ; coord Q ; cross Q ; coord P ; cross P]
    float* bias_m;
    CUtensorMap tm_wm, tm_we0;                                    // 128-row boxes (half a 256-column group per CTA of a pair)
    float *bias_e, *b3, *b4, *w1e_e, *w1e_c, *w1e_x;                   // device fp32
    CUtensorMap tm_w2_e, tm_w2_c, tm_w2_x, tm_w3, tm_w4;
    EdgeConsts c_e, c_c, c_x;                                            // host copies (kernel parameters)
    __nv_bfloat16* et_e = nullptr;                                       // edge-type vectors of the edge model: [3][256] bf16, halved
    float *et_c = nullptr, *et_x = nullptr;                              // ... of the coordinate heads: [3][256] fp32, halved
    float att_bias;
};

struct DndmEngine {
    DndmConfig cfg;
    int num_sms = 148;
    bool weights_loaded = false;
    bool edge_types = false;       // the weight table carries edge-type vectors (dynamics.py:118-127)
    // workspace
    float *x0 = nullptr, *xa = nullptr, *xb = nullptr, *h = nullptr, *att = nullptr;
    __nv_bfloat16* msg = nullptr;  // [E,256] bf16 edge messages of the current block
    __nv_bfloat16* pq = nullptr;   // [N,1536] bf16 node projections: edge P|Q, coord P, cross P, coord Q, cross Q
    float *r0 = nullptr, *phi = nullptr, *psi = nullptr, *pocket_sum = nullptr;
    __nv_bfloat16 *hcat = nullptr, *hid = nullptr;
    int *node_sample = nullptr, *lig_ptr = nullptr, *pok_ptr = nullptr, *deg = nullptr, *row_ptr = nullptr;
    int *ecol = nullptr, *erow = nullptr, *scalars = nullptr, *block_sums = nullptr;
    // compacted graphs of the last PRUNE_LEVELS blocks (exact dead-work elimination, graph.cuh): level k = 0 is the last block
    // (receivers: ligand atoms + their pocket senders), level k the block k before it (level k-1's receivers + all their senders)
    static constexpr int PRUNE_LEVELS = 3;
    int *deg_act[PRUNE_LEVELS] = {}, *rp_act[PRUNE_LEVELS] = {}, *erow_c[PRUNE_LEVELS] = {}, *ecol_c[PRUNE_LEVELS] = {};
    float* r0_c[PRUNE_LEVELS] = {};
    PocketLists pp{};                            // pocket-pocket candidate lists (graph.cuh)
    unsigned* flags = nullptr;
    long long* mol_off = nullptr;                // [max_samples + 1] byte offsets of the per-molecule bond matrices
    // the radius graph only needs coordinates: it is built on a side stream while the main stream encodes the features
    float* xg = nullptr;                         // [N,3] coordinates gathered for the graph branch
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_last = nullptr, ev_join_pp = nullptr;
    // cached batch layout (dndm_set_static_masks)
    bool static_masks = false;
    const int64_t *last_lm = nullptr, *last_pm = nullptr;
    int last_nl = -1, last_np = -1, last_ns = -1;
    CUtensorMap tm_hcat, tm_hid;
    CUtensorMap to_pq32, to_hcat32, to_hid32;    // 32-column bf16 boxes (SWIZZLE_64B) for the weight-resident GEMM
    CUtensorMap to_msg;                          // TMA-store destination of the edge messages (box 32 rows x 64 cols)   // TMA-store destinations of the node GEMMs (32-row boxes)
    // weights
    std::vector<LayerWeights> layers;
    std::vector<void*> weight_allocs;
    EncoderWeights enc_l{}, enc_p{};
    DecoderWeights dec_l{}, dec_p{};
    // last call
    int last_n_lig = 0, last_n_nodes = 0;
    float* x_final = nullptr;
    // trace
    float *h_trace = nullptr, *x_trace = nullptr, *h0_snap = nullptr;
    int max_trace_nodes = 0;
    // per-section CUDA-event profiling (off by default; never used under graph capture)
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Section { int cat; cudaEvent_t a, b; };
    std::vector<Section> sections;
    size_t ev_used = 0;
};

enum { PROF_GCL = 0, PROF_HEAD = 1, PROF_GEMM = 2, PROF_GRAPH = 3, PROF_NODE = 4, PROF_NCAT = 5 };

static cudaEvent_t prof_event(DndmEngine* e) {
    if (e->ev_used == e->ev_pool.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->ev_pool.push_back(ev);
    }
    return e->ev_pool[e->ev_used++];
}
struct ProfScope {
    DndmEngine* e; cudaStream_t st; cudaEvent_t a, b; int cat; bool on;
    ProfScope(DndmEngine* e_, int cat_, cudaStream_t st_) : e(e_), st(st_), cat(cat_), on(e_->profile) {
        if (on) { a = prof_event(e); b = prof_event(e); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(b, st); e->sections.push_back({cat, a, b}); }
    }
};

template <class T>
static int dev_alloc(T** p, size_t n) {
    CU_CHECK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
    return DNDM_OK;
}
#define RET_IF(x)              \
    do {                       \
        int _r = (x);          \
        if (_r != DNDM_OK) return _r; \
    } while (0)

extern "C" const char* dndm_version(void) { return "diffndm_b200 0.3 (sm_100a, tcgen05 cta_group::2 / TMA)"; }
extern "C" const char* dndm_last_error(void) { return g_err; }
extern "C" int64_t dndm_launch_count(void) { return g_launches; }

extern "C" int dndm_engine_create(const DndmConfig* cfg, DndmEngine** out) {
    if (!cfg || !out) return set_err(DNDM_EINVAL, "null argument");
    if (cfg->hidden_nf != EK_H) return set_err(DNDM_EINVAL, "hidden_nf=%d unsupported (compiled for %d)", cfg->hidden_nf, EK_H);
    if (cfg->atom_nf < 1 || cfg->atom_nf > 29 || cfg->residue_nf < 1 || cfg->residue_nf > 29)
        return set_err(DNDM_EINVAL, "atom_nf/residue_nf must be in [1,29]");
    if (2 * cfg->atom_nf > 64 || 2 * cfg->residue_nf > 64) return set_err(DNDM_EINVAL, "encoder hidden width > 64 unsupported");
    if (cfg->max_nodes < 1 || cfg->max_edges < 1 || cfg->max_samples < 1) return set_err(DNDM_EINVAL, "bad capacities");
    CU_CHECK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return set_err(DNDM_ECUDA, "device sm_%d%d is not Blackwell sm_100", prop.major, prop.minor);
    DndmEngine* e = new DndmEngine();
    e->cfg = *cfg;
    e->num_sms = prop.multiProcessorCount;
    // experiment knob (scripts/two_stream_bench.py): size the persistent grids for fewer SMs, so that a second engine's
    // kernels on another stream find free SMs
    if (const char* v = getenv("DNDM_SMS")) { const int n = atoi(v); if (n >= 2 && n <= e->num_sms) e->num_sms = n & ~1; }
    g_num_sms = e->num_sms;
    const size_t N = (size_t)((cfg->max_nodes + 127) / 128) * 128, E = cfg->max_edges, B = cfg->max_samples;
    RET_IF(dev_alloc(&e->x0, N * 3)); RET_IF(dev_alloc(&e->xa, N * 3)); RET_IF(dev_alloc(&e->xb, N * 3));
    RET_IF(dev_alloc(&e->h, N * 256)); RET_IF(dev_alloc(&e->pq, N * 1536));
    RET_IF(dev_alloc(&e->msg, (E + 128) * 256)); RET_IF(dev_alloc(&e->att, E + 128));
    RET_IF(dev_alloc(&e->r0, E)); RET_IF(dev_alloc(&e->phi, E)); RET_IF(dev_alloc(&e->psi, E));
    RET_IF(dev_alloc(&e->pocket_sum, B * 3));
    RET_IF(dev_alloc(&e->hcat, N * 512)); RET_IF(dev_alloc(&e->hid, N * 256));
    RET_IF(dev_alloc(&e->node_sample, N)); RET_IF(dev_alloc(&e->lig_ptr, B + 1)); RET_IF(dev_alloc(&e->pok_ptr, B + 1));
    RET_IF(dev_alloc(&e->deg, N)); RET_IF(dev_alloc(&e->row_ptr, N + 1));
    RET_IF(dev_alloc(&e->ecol, E)); RET_IF(dev_alloc(&e->erow, E + 1)); RET_IF(dev_alloc(&e->scalars, 8)); RET_IF(dev_alloc(&e->block_sums, 1024));
    for (int k = 0; k < DndmEngine::PRUNE_LEVELS; ++k) {
        RET_IF(dev_alloc(&e->deg_act[k], N)); RET_IF(dev_alloc(&e->rp_act[k], N + 1)); RET_IF(dev_alloc(&e->erow_c[k], E + 1));
        RET_IF(dev_alloc(&e->ecol_c[k], E)); RET_IF(dev_alloc(&e->r0_c[k], E));
    }
    RET_IF(dev_alloc(&e->flags, 1));
    RET_IF(dev_alloc(&e->pp.canon, N * 3)); RET_IF(dev_alloc(&e->pp.ptr, B + 1)); RET_IF(dev_alloc(&e->pp.meta, 8));
    RET_IF(dev_alloc(&e->pp.cand, N * PP_CAP)); RET_IF(dev_alloc(&e->pp.cnt, N));
    CU_CHECK(cudaMemset(e->pp.meta, 0, 8 * sizeof(int)));
    RET_IF(dev_alloc(&e->xg, N * 3));
    RET_IF(dev_alloc(&e->mol_off, B + 1));
    CU_CHECK(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_join_last, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&e->ev_join_pp, cudaEventDisableTiming));
    CU_CHECK(cudaMemset(e->flags, 0, 4));
    CU_CHECK(cudaMemset(e->hcat, 0, N * 512 * 2));
    CU_CHECK(cudaMemset(e->hid, 0, N * 256 * 2));
    RET_IF(make_tmap_bf16(&e->tm_hcat, e->hcat, N, 512, 512, WR_BM));
    RET_IF(make_tmap_bf16(&e->tm_hid, e->hid, N, 256, 256, WR_BM));
    RET_IF(make_tmap_bf16_box(&e->to_msg, e->msg, E + 128, 256, 256, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    RET_IF(make_tmap_bf16_box(&e->to_pq32, e->pq, N, 1536, 1536, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    RET_IF(make_tmap_bf16_box(&e->to_hcat32, e->hcat, N, 512, 512, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    RET_IF(make_tmap_bf16_box(&e->to_hid32, e->hid, N, 256, 256, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    CU_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<256, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, WresShape<256, 256>::smem_bytes));
    CU_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WresShape<512, 128>::smem_bytes));
    CU_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<256, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WresShape<256, 128>::smem_bytes));
    CU_CHECK(cudaFuncSetAttribute(edge_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(edge_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(edge_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(edge_pair_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(edge_pair_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM_BYTES));
    CU_CHECK(cudaFuncSetAttribute(edge_pair_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EP_SMEM_BYTES));
    *out = e;
    return DNDM_OK;
}

static void free_weights(DndmEngine* e) {
    for (void* p : e->weight_allocs) cudaFree(p);
    e->weight_allocs.clear();
    e->layers.clear();
    e->weights_loaded = false;
}

extern "C" void dndm_engine_destroy(DndmEngine* e) {
    if (!e) return;
    free_weights(e);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    void* bufs[] = {e->x0, e->xa, e->xb, e->h, e->pq, e->msg, e->att, e->r0, e->phi, e->psi, e->pocket_sum, e->hcat,
                    e->hid, e->node_sample, e->lig_ptr, e->pok_ptr, e->deg, e->row_ptr, e->ecol, e->erow, e->scalars, e->block_sums, e->flags, e->xg, e->mol_off, e->h0_snap, e->pp.canon, e->pp.ptr, e->pp.meta, e->pp.cand, e->pp.cnt};
    for (void* p : bufs) cudaFree(p);
    for (int k = 0; k < DndmEngine::PRUNE_LEVELS; ++k) {
        cudaFree(e->deg_act[k]); cudaFree(e->rp_act[k]); cudaFree(e->erow_c[k]); cudaFree(e->ecol_c[k]); cudaFree(e->r0_c[k]);
    }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->ev_join_last) cudaEventDestroy(e->ev_join_last);
    if (e->ev_join_pp) cudaEventDestroy(e->ev_join_pp);
    if (e->side) cudaStreamDestroy(e->side);
    delete e;
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
struct HostW {
    const float* d;
    int rows, cols;
};
static __nv_bfloat16 f2bf(float f) { return __float2bfloat16_rn(f); }

template <class T>
static int upload(DndmEngine* e, const std::vector<T>& host, T** dev) {
    void* p = nullptr;
    CU_CHECK(cudaMalloc(&p, host.size() * sizeof(T)));
    e->weight_allocs.push_back(p);
    CU_CHECK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = reinterpret_cast<T*>(p);
    return DNDM_OK;
}

extern "C" int dndm_engine_load_weights(DndmEngine* e, const DndmWeight* weights, int32_t n_weights) {
    if (!e || !weights) return set_err(DNDM_EINVAL, "null argument");
    CU_CHECK(cudaSetDevice(e->cfg.device));
    CU_CHECK(cudaDeviceSynchronize());
    free_weights(e);
    e->edge_types = false;
    std::map<std::string, HostW> W;
    for (int i = 0; i < n_weights; ++i) W[weights[i].name] = HostW{weights[i].data, weights[i].rows, weights[i].cols};
    auto need = [&](const std::string& k, int rows, int cols, const float** out) -> int {
        auto it = W.find(k);
        if (it == W.end()) return set_err(DNDM_EWEIGHTS, "missing weight '%s'", k.c_str());
        if (it->second.rows != rows || it->second.cols != cols)
            return set_err(DNDM_EWEIGHTS, "weight '%s' has shape [%d,%d], expected [%d,%d]", k.c_str(), it->second.rows,
                           it->second.cols, rows, cols);
        *out = it->second.d;
        return DNDM_OK;
    };
    const int H = e->cfg.hidden_nf, J = e->cfg.joint_nf, A = e->cfg.atom_nf, R = e->cfg.residue_nf;
    const int DE = 2, KIN = 2 * H + DE;

    // ---- encoder (+ embedding) and decoder (+ embedding_out), pre-composed in fp64 ----
    const float *emb_w, *emb_b, *out_w, *out_b;
    RET_IF(need("egnn.embedding.weight", H, J + 1, &emb_w));
    RET_IF(need("egnn.embedding.bias", H, 1, &emb_b));
    RET_IF(need("egnn.embedding_out.weight", J + 1, H, &out_w));
    RET_IF(need("egnn.embedding_out.bias", J + 1, 1, &out_b));
    auto pack_encoder = [&](const char* pre, int nf, EncoderWeights* ew) -> int {
        const float *w1, *b1, *w2, *b2;
        const int hid = 2 * nf;
        RET_IF(need(std::string(pre) + ".0.weight", hid, nf, &w1));
        RET_IF(need(std::string(pre) + ".0.bias", hid, 1, &b1));
        RET_IF(need(std::string(pre) + ".2.weight", J, hid, &w2));
        RET_IF(need(std::string(pre) + ".2.bias", J, 1, &b2));
        std::vector<float> hw1(w1, w1 + hid * nf), hb1(b1, b1 + hid), wc2((size_t)H * hid), bc(H), wt(H);
        for (int k = 0; k < H; ++k) {
            for (int j = 0; j < hid; ++j) {
                double s = 0;
                for (int m = 0; m < J; ++m) s += (double)emb_w[k * (J + 1) + m] * (double)w2[m * hid + j];
                wc2[(size_t)j * H + k] = (float)s;           // stored transposed [hid][256]: thread k's loads coalesce
            }
            double s = emb_b[k];
            for (int m = 0; m < J; ++m) s += (double)emb_w[k * (J + 1) + m] * (double)b2[m];
            bc[k] = (float)s;
            wt[k] = emb_w[k * (J + 1) + J];
        }
        // one-hot inputs (every pocket atom / residue): the whole encoder is a table lookup.  tb[s][type][k] = the embedding of
        // enc2(SiLU(enc1(v_s e_type))) without the time term, evaluated in fp64 from the reference's own formula, for the two
        // scales a one-hot row arrives in: v = 1 (raw) and v = 1/4 (after ConditionalDDPM.normalize with the normalize_factors
        // [1, 4] of every config, en_diffusion.py:885-900); any other row takes the general path
        std::vector<float> tb((size_t)2 * nf * H);
        std::vector<double> enc(J);
        for (int sc = 0; sc < 2; ++sc) {
            const double v = sc == 0 ? 1.0 : 0.25;
            for (int ty = 0; ty < nf; ++ty) {
                for (int m = 0; m < J; ++m) {
                    double s = b2[m];
                    for (int j = 0; j < hid; ++j) {
                        const double a = (double)b1[j] + v * (double)w1[j * nf + ty];
                        s += (double)w2[m * hid + j] * (a / (1.0 + std::exp(-a)));
                    }
                    enc[m] = s;
                }
                for (int k = 0; k < H; ++k) {
                    double s = emb_b[k];
                    for (int m = 0; m < J; ++m) s += (double)emb_w[k * (J + 1) + m] * enc[m];
                    tb[((size_t)sc * nf + ty) * H + k] = (float)s;
                }
            }
        }
        float *d1, *d2, *d3, *d4, *d5, *d6;
        RET_IF(upload(e, hw1, &d1)); RET_IF(upload(e, hb1, &d2)); RET_IF(upload(e, wc2, &d3));
        RET_IF(upload(e, bc, &d4)); RET_IF(upload(e, wt, &d5)); RET_IF(upload(e, tb, &d6));
        *ew = EncoderWeights{d1, d2, d3, d4, d5, d6, nf, hid};
        return DNDM_OK;
    };
    RET_IF(pack_encoder("atom_encoder", A, &e->enc_l));
    RET_IF(pack_encoder("residue_encoder", R, &e->enc_p));
    auto pack_decoder = [&](const char* pre, int nf, DecoderWeights* dw) -> int {
        const float *w0, *b0, *w2, *b2;
        const int hid = 2 * nf;
        RET_IF(need(std::string(pre) + ".0.weight", hid, J, &w0));
        RET_IF(need(std::string(pre) + ".0.bias", hid, 1, &b0));
        RET_IF(need(std::string(pre) + ".2.weight", nf, hid, &w2));
        RET_IF(need(std::string(pre) + ".2.bias", nf, 1, &b2));
        std::vector<float> wc((size_t)hid * H), bc(hid), hw2(w2, w2 + nf * hid), hb2(b2, b2 + nf);
        for (int j = 0; j < hid; ++j) {
            for (int k = 0; k < H; ++k) {
                double s = 0;
                for (int m = 0; m < J; ++m) s += (double)w0[j * J + m] * (double)out_w[m * H + k];
                wc[(size_t)j * H + k] = (float)s;
            }
            double s = b0[j];
            for (int m = 0; m < J; ++m) s += (double)w0[j * J + m] * (double)out_b[m];
            bc[j] = (float)s;
        }
        float *d1, *d2, *d3, *d4;
        RET_IF(upload(e, wc, &d1)); RET_IF(upload(e, bc, &d2)); RET_IF(upload(e, hw2, &d3)); RET_IF(upload(e, hb2, &d4));
        *dw = DecoderWeights{d1, d2, d3, d4, nf, hid};
        return DNDM_OK;
    };
    RET_IF(pack_decoder("atom_decoder", A, &e->dec_l));
    RET_IF(pack_decoder("residue_decoder", R, &e->dec_p));

    // ---- per-block weights ----
    e->layers.resize(e->cfg.n_layers);
    for (int l = 0; l < e->cfg.n_layers; ++l) {
        LayerWeights& L = e->layers[l];
        const std::string p = "egnn.e_block_" + std::to_string(l) + ".";
        const float *e0w, *e0b, *e2w, *e2b, *n0w, *n0b, *n2w, *n2b, *aw, *ab;
        const float *c0w, *c0b, *c2w, *c2b, *c4w, *x0w, *x0b, *x2w, *x2b, *x4w;
        RET_IF(need(p + "gcl_0.edge_mlp.0.weight", H, KIN, &e0w)); RET_IF(need(p + "gcl_0.edge_mlp.0.bias", H, 1, &e0b));
        RET_IF(need(p + "gcl_0.edge_mlp.2.weight", H, H, &e2w)); RET_IF(need(p + "gcl_0.edge_mlp.2.bias", H, 1, &e2b));
        RET_IF(need(p + "gcl_0.node_mlp.0.weight", H, 2 * H, &n0w)); RET_IF(need(p + "gcl_0.node_mlp.0.bias", H, 1, &n0b));
        RET_IF(need(p + "gcl_0.node_mlp.2.weight", H, H, &n2w)); RET_IF(need(p + "gcl_0.node_mlp.2.bias", H, 1, &n2b));
        RET_IF(need(p + "gcl_0.att_mlp.0.weight", 1, H, &aw)); RET_IF(need(p + "gcl_0.att_mlp.0.bias", 1, 1, &ab));
        RET_IF(need(p + "gcl_equiv.coord_mlp.0.weight", H, KIN, &c0w)); RET_IF(need(p + "gcl_equiv.coord_mlp.0.bias", H, 1, &c0b));
        RET_IF(need(p + "gcl_equiv.coord_mlp.2.weight", H, H, &c2w)); RET_IF(need(p + "gcl_equiv.coord_mlp.2.bias", H, 1, &c2b));
        RET_IF(need(p + "gcl_equiv.coord_mlp.4.weight", 1, H, &c4w));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.0.weight", H, KIN, &x0w));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.0.bias", H, 1, &x0b));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.2.weight", H, H, &x2w));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.2.bias", H, 1, &x2b));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.4.weight", 1, H, &x4w));

        auto split_first = [&](const float* w, std::vector<__nv_bfloat16>& dst, size_t row_off_a, size_t row_off_b) {
            for (int o = 0; o < H; ++o)
                for (int k = 0; k < H; ++k) {
                    // halved (exact): the edge kernel evaluates SiLU(x) = h + h tanh(h) on h = x/2
                    dst[(row_off_a + o) * H + k] = f2bf(0.5f * w[(size_t)o * KIN + k]);
                    dst[(row_off_b + o) * H + k] = f2bf(0.5f * w[(size_t)o * KIN + H + k]);
                }
        };
        auto edge_cols = [&](const float* w) {
            std::vector<float> v(2 * H);
            for (int o = 0; o < H; ++o) {
                v[o] = 0.5f * w[(size_t)o * KIN + 2 * H];          // (half) coefficient of the current radial
                v[H + o] = 0.5f * w[(size_t)o * KIN + 2 * H + 1];  // (half) coefficient of the input radial
            }
            return v;
        };
        auto to_bf = [&](const float* w, size_t n) {
            std::vector<__nv_bfloat16> v(n);
            for (size_t i = 0; i < n; ++i) v[i] = f2bf(w[i]);
            return v;
        };
        // this block's own edge-model projection [P | Q] (only block 0's is launched: later blocks get theirs from the
        // previous block's merged projection)
        std::vector<__nv_bfloat16> pe((size_t)2 * H * H);
        split_first(e0w, pe, 0, H);
        std::vector<float> be(2 * H, 0.f);
        for (int o = 0; o < H; ++o) be[o] = 0.5f * e0b[o];
        RET_IF(upload(e, pe, &L.wproj_e));
        RET_IF(upload(e, be, &L.bias_e));
        RET_IF(upload(e, edge_cols(e0w), &L.w1e_e)); RET_IF(upload(e, edge_cols(c0w), &L.w1e_c));
        RET_IF(upload(e, edge_cols(x0w), &L.w1e_x));
        {   // optional: <mlp>.0.edge_type_bias [3, H] = W1[:, 2H+2:] E[type] (folded by the host binding); all three or none
            auto tb = [&](const char* m) -> const float* {
                auto it = W.find(p + m + ".0.edge_type_bias");
                return (it != W.end() && it->second.rows == 3 && it->second.cols == H) ? it->second.d : nullptr;
            };
            const float *te = tb("gcl_0.edge_mlp"), *tc = tb("gcl_equiv.coord_mlp"), *tx = tb("gcl_equiv.cross_product_mlp");
            if ((te != nullptr) != (tc != nullptr) || (te != nullptr) != (tx != nullptr) || (l > 0 && (te != nullptr) != e->edge_types))
                return set_err(DNDM_EWEIGHTS, "edge_type_bias must be given for every edge MLP of every block or for none");
            e->edge_types = te != nullptr;
            if (te) {
                std::vector<__nv_bfloat16> be(3 * H);
                std::vector<float> bc(3 * H), bx(3 * H);
                for (int i = 0; i < 3 * H; ++i) { be[i] = f2bf(0.5f * te[i]); bc[i] = 0.5f * tc[i]; bx[i] = 0.5f * tx[i]; }
                RET_IF(upload(e, be, &L.et_e)); RET_IF(upload(e, bc, &L.et_c)); RET_IF(upload(e, bx, &L.et_x));
            }
        }
        auto to_bf_half = [&](const float* w, size_t n) {       // edge-MLP second layers: the kernel evaluates SiLU on x/2
            std::vector<__nv_bfloat16> v(n);
            for (size_t i = 0; i < n; ++i) v[i] = f2bf(0.5f * w[i]);
            return v;
        };
        RET_IF(upload(e, to_bf_half(e2w, (size_t)H * H), &L.w2_e)); RET_IF(upload(e, to_bf_half(c2w, (size_t)H * H), &L.w2_c));
        RET_IF(upload(e, to_bf_half(x2w, (size_t)H * H), &L.w2_x));
        RET_IF(upload(e, to_bf(n0w, (size_t)H * 2 * H), &L.w3)); RET_IF(upload(e, to_bf(n2w, (size_t)H * H), &L.w4));
        RET_IF(upload(e, std::vector<float>(n0b, n0b + H), &L.b3)); RET_IF(upload(e, std::vector<float>(n2b, n2b + H), &L.b4));
        for (int o = 0; o < H; ++o) {
            L.c_e.b2[o] = 0.5f * e2b[o]; L.c_e.wout[o] = aw[o];
            L.c_c.b2[o] = 0.5f * c2b[o]; L.c_c.wout[o] = c4w[o];
            L.c_x.b2[o] = 0.5f * x2b[o]; L.c_x.wout[o] = x4w[o];
        }
        L.att_bias = ab[0];
        RET_IF(make_tmap_bf16(&L.tm_w3, L.w3, H, 2 * H, 2 * H, 64));        // boxes of BN/2 weight rows: one CTA's half of a column group
        RET_IF(make_tmap_bf16(&L.tm_w4, L.w4, H, H, H, 64));
        RET_IF(make_tmap_bf16(&L.tm_w2_e, L.w2_e, H, H, H, 128));      // one box = 128 output channels x 64 inputs
        RET_IF(make_tmap_bf16(&L.tm_w2_c, L.w2_c, H, H, H, 128));      // one box = 128 output channels x 64 inputs
        RET_IF(make_tmap_bf16(&L.tm_w2_x, L.w2_x, H, H, H, 128));      // one box = 128 output channels x 64 inputs
    }
    // ---- merged projection weights for the weight-resident GEMM (pq columns: [0,512) edge P|Q of the NEXT block,
    //      [512,1024) coord Q | cross Q, [1024,1536) coord P | cross P of THIS block); all pre-halved ----
    for (int l = 0; l < e->cfg.n_layers; ++l) {
        LayerWeights& L = e->layers[l];
        const std::string p = "egnn.e_block_" + std::to_string(l) + ".";
        const float *c0w, *c0b, *x0w, *x0b, *n0w = nullptr, *n0b = nullptr;
        RET_IF(need(p + "gcl_equiv.coord_mlp.0.weight", H, KIN, &c0w)); RET_IF(need(p + "gcl_equiv.coord_mlp.0.bias", H, 1, &c0b));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.0.weight", H, KIN, &x0w));
        RET_IF(need(p + "gcl_equiv.cross_product_mlp.0.bias", H, 1, &x0b));
        if (l + 1 < e->cfg.n_layers) {
            const std::string pn = "egnn.e_block_" + std::to_string(l + 1) + ".";
            RET_IF(need(pn + "gcl_0.edge_mlp.0.weight", H, KIN, &n0w)); RET_IF(need(pn + "gcl_0.edge_mlp.0.bias", H, 1, &n0b));
        }
        std::vector<__nv_bfloat16> wm((size_t)6 * H * H, f2bf(0.f));     // rows [4H,6H): receiver parts (ligand rows only)
        std::vector<float> bm(6 * H, 0.f);
        for (int o = 0; o < H; ++o)
            for (int k = 0; k < H; ++k) {
                if (n0w) {
                    wm[((size_t)o) * H + k] = f2bf(0.5f * n0w[(size_t)o * KIN + k]);              // next edge, receiver part
                    wm[((size_t)H + o) * H + k] = f2bf(0.5f * n0w[(size_t)o * KIN + H + k]);      // next edge, sender part
                }
                wm[((size_t)2 * H + o) * H + k] = f2bf(0.5f * c0w[(size_t)o * KIN + H + k]);      // coord, sender part
                wm[((size_t)3 * H + o) * H + k] = f2bf(0.5f * x0w[(size_t)o * KIN + H + k]);      // cross, sender part
                wm[((size_t)4 * H + o) * H + k] = f2bf(0.5f * c0w[(size_t)o * KIN + k]);          // coord, receiver part
                wm[((size_t)5 * H + o) * H + k] = f2bf(0.5f * x0w[(size_t)o * KIN + k]);          // cross, receiver part
            }
        for (int o = 0; o < H; ++o) {
            if (n0b) bm[o] = 0.5f * n0b[o];
            bm[4 * H + o] = 0.5f * c0b[o];
            bm[5 * H + o] = 0.5f * x0b[o];
        }
        RET_IF(upload(e, wm, &L.wm)); RET_IF(upload(e, bm, &L.bias_m));
        RET_IF(make_tmap_bf16(&L.tm_wm, L.wm, 6 * H, H, H, 128));
        RET_IF(make_tmap_bf16(&L.tm_we0, L.wproj_e, 2 * H, H, H, 128));
    }
    e->weights_loaded = true;
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
// n_full column groups over M rows followed by n_tail groups over the first M_tail rows (see gemm_pair_kernel); every group
// is served by CTA pairs that stride over its 256-row blocks
template <int kK = 256, int kBN = 256>
static int launch_wres(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to16, int M,
                       int n_full, int g0, int a_col0, const WresEpilogue& ep, int n_tail = 0, int M_tail = 0, bool pdl = false) {
    if (M <= 0) return DNDM_OK;
    if (M_tail <= 0) n_tail = 0;
    const int m_pairs = (M + 2 * WR_BM - 1) / (2 * WR_BM), t_pairs = (M_tail + 2 * WR_BM - 1) / (2 * WR_BM);
    const int slots = g_num_sms / 2 > 0 ? g_num_sms / 2 : 1;      // co-resident CTA pairs
    int pairs_tail = 0;
    if (n_tail > 0) {                                   // share the SMs in proportion to the row blocks
        const double per_pair = (double)(n_full * m_pairs + n_tail * t_pairs) / slots;
        pairs_tail = (int)(t_pairs / per_pair + 0.999);
        if (pairs_tail < 1) pairs_tail = 1;
        if (pairs_tail > t_pairs) pairs_tail = t_pairs;
    }
    int pairs_full = (slots - n_tail * pairs_tail) / n_full;
    if (pairs_full < 1) pairs_full = 1;
    if (pairs_full > m_pairs) pairs_full = m_pairs;
    const int grid = 2 * (n_full * pairs_full + n_tail * pairs_tail);
    CU_CHECK(launch_pdl(pdl, gemm_pair_kernel<kK, kBN>, dim3(grid), dim3(WR_THREADS), WresShape<kK, kBN>::smem_bytes, st, ta, tw, to16,
                        M, M_tail, a_col0, g0, n_full, pairs_full, pairs_tail > 0 ? pairs_tail : 1, ep));
    COUNT_LAUNCH(1);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

static int prepare_batch(DndmEngine* e, const int64_t* lig_mask, const int64_t* pocket_mask, int n_lig, int n_pocket,
                         int n_samples, cudaStream_t st) {
    if (n_lig < 1 || n_pocket < 0 || n_samples < 1) return set_err(DNDM_EINVAL, "empty batch");
    if (n_lig + n_pocket > e->cfg.max_nodes || n_samples > e->cfg.max_samples)
        return set_err(DNDM_ECAPACITY, "batch (%d nodes, %d samples) exceeds engine capacity (%d, %d)", n_lig + n_pocket,
                       n_samples, e->cfg.max_nodes, e->cfg.max_samples);
    if (e->static_masks && lig_mask == e->last_lm && pocket_mask == e->last_pm && n_lig == e->last_nl &&
        n_pocket == e->last_np && n_samples == e->last_ns)
        return DNDM_OK;                               // same buffers, caller-promised unchanged contents
    e->last_lm = lig_mask; e->last_pm = pocket_mask; e->last_nl = n_lig; e->last_np = n_pocket; e->last_ns = n_samples;
    {
        const int n = n_lig > n_samples + 1 ? n_lig : n_samples + 1;
        mask_to_ptr_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(lig_mask), n_lig, n_samples,
                                                            e->lig_ptr, e->node_sample);
    }
    {
        const int n = n_pocket > n_samples + 1 ? n_pocket : n_samples + 1;
        mask_to_ptr_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(pocket_mask), n_pocket,
                                                            n_samples, e->pok_ptr, e->node_sample + n_lig);
    }
    COUNT_LAUNCH(2);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

// exclusive scan of deg[0..n) -> out[0..n]; total -> scalars[slot_total]; out[n_lig] -> scalars[slot_lig] (if >= 0)
static int exclusive_scan(DndmEngine* e, const int* deg, int* out, int n, int n_lig, int slot_total, int slot_lig,
                          cudaStream_t st) {
    const int nb = (n + SCAN_ELEMS - 1) / SCAN_ELEMS;
    if (nb > 1024) return set_err(DNDM_ECAPACITY, "more than %d nodes per call are not supported by the scan", 1024 * SCAN_ELEMS);
    scan_block_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(deg, n, e->block_sums);
    scan_offsets_kernel<<<1, 1024, 0, st>>>(e->block_sums, nb, n, n_lig, e->cfg.max_edges, out, e->scalars, e->flags, slot_total,
                                           slot_lig);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(deg, e->block_sums, n, n_lig, e->cfg.max_edges, out, e->scalars, slot_lig);
    COUNT_LAUNCH(3);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

// scalars: [0] E, [1] E of the ligand rows, [2 + k] E of pruning level k
static int prune_levels(const DndmEngine* e) {
    const int by_env = g_prune_levels, by_depth = e->cfg.n_layers;
    const int n = by_env < by_depth ? by_env : by_depth;
    return n < DndmEngine::PRUNE_LEVELS ? n : DndmEngine::PRUNE_LEVELS;
}

// compacted edge lists of the receivers the last blocks still need
static int build_last_block_edges(DndmEngine* e, int n_lig, int n_nodes, cudaStream_t st) {
    mark_active_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->ecol, e->deg, e->scalars, n_lig, n_nodes, e->deg_act[0], 0);
    mark_active_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->ecol, e->deg, e->scalars, n_lig, n_nodes, e->deg_act[0], 1);
    COUNT_LAUNCH(2);
    for (int k = 0; k < prune_levels(e); ++k) {
        if (k > 0) {               // level k-1's receivers + all their senders
            mark_senders_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->ecol_c[k - 1], e->deg, e->deg_act[k - 1], e->scalars, 1 + k, n_nodes,
                                                                e->deg_act[k], 0);
            mark_senders_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->ecol_c[k - 1], e->deg, e->deg_act[k - 1], e->scalars, 1 + k, n_nodes,
                                                                e->deg_act[k], 1);
            COUNT_LAUNCH(2);
        }
        RET_IF(exclusive_scan(e, e->deg_act[k], e->rp_act[k], n_nodes, n_lig, 2 + k, -1, st));
        compact_edges_kernel<<<(n_nodes * 32 + 255) / 256, 256, 0, st>>>(e->row_ptr, e->rp_act[k], e->erow, e->ecol, e->r0, n_nodes,
                                                                         e->erow_c[k], e->ecol_c[k], e->r0_c[k]);
        COUNT_LAUNCH(1);
    }
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

// xh_lig / xh_pocket: the call's inputs (rows of 3 + atom_nf / 3 + residue_nf floats); x: [N,3] scratch the coordinates are
// gathered into (read by the row kernels and by refresh_pocket_lists)
static int build_graph(DndmEngine* e, const float* xh_lig, const float* xh_pocket, float* x, int n_lig, int n_nodes,
                       int n_samples, cudaStream_t st) {
    GraphParams gp;
    gp.x = x; gp.lig_ptr = e->lig_ptr; gp.pok_ptr = e->pok_ptr; gp.node_sample = e->node_sample;
    gp.n_lig = n_lig; gp.n_nodes = n_nodes;
    auto sq = [](float c) { return c < 0.f ? -1.f : c * c; };
    gp.cut2_l = sq(e->cfg.edge_cutoff_ligand); gp.cut2_p = sq(e->cfg.edge_cutoff_pocket);
    gp.cut2_i = sq(e->cfg.edge_cutoff_interaction);
    const int n_pocket = n_nodes - n_lig;
    const bool lists = n_pocket > 0 && g_pp_lists;
    PocketLists pl = e->pp;
    if (!lists) pl.meta = nullptr;
    {
        const int n = n_nodes > n_samples + 1 ? n_nodes : n_samples + 1;
        gather_verify_kernel<<<(n + 255) / 256, 256, 0, st>>>(xh_lig, xh_pocket, 3 + e->cfg.atom_nf, 3 + e->cfg.residue_nf, gp, pl,
                                                            n_samples, x);
        COUNT_LAUNCH(1);
    }
    const int blocks = ((n_lig + (n_pocket + GR_ROWS - 1) / GR_ROWS) * 32 + 255) / 256;    // a warp per ligand row, per GR_ROWS pocket rows
    graph_rows_kernel<false><<<blocks, 256, 0, st>>>(gp, pl, e->deg, nullptr, nullptr, nullptr, nullptr, e->cfg.max_edges);
    RET_IF(exclusive_scan(e, e->deg, e->row_ptr, n_nodes, n_lig, 0, 1, st));
    graph_rows_kernel<true><<<blocks, 256, 0, st>>>(gp, pl, nullptr, e->row_ptr, e->ecol, e->erow, e->r0, e->cfg.max_edges);
    COUNT_LAUNCH(2);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

// after build_graph (and after whoever waits for the graph has been released): rebuild the pocket-pocket candidate lists if
// this call found them missing or stale, then mark them as describing this call's layout
static int refresh_pocket_lists(DndmEngine* e, const float* x, int n_lig, int n_nodes, int n_samples, cudaStream_t st) {
    const int n_pocket = n_nodes - n_lig;
    if (n_pocket <= 0 || !g_pp_lists) return DNDM_OK;
    GraphParams gp;
    gp.x = x; gp.lig_ptr = e->lig_ptr; gp.pok_ptr = e->pok_ptr; gp.node_sample = e->node_sample;
    gp.n_lig = n_lig; gp.n_nodes = n_nodes;
    auto sq = [](float c) { return c < 0.f ? -1.f : c * c; };
    gp.cut2_l = sq(e->cfg.edge_cutoff_ligand); gp.cut2_p = sq(e->cfg.edge_cutoff_pocket);
    gp.cut2_i = sq(e->cfg.edge_cutoff_interaction);
    const long long threads = (long long)n_pocket * 32 > n_samples + 1 ? (long long)n_pocket * 32 : n_samples + 1;
    pp_rebuild_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(gp, e->pp, n_samples);
    pp_commit_kernel<<<1, 1, 0, st>>>(e->pp, n_lig, n_pocket, n_samples);
    COUNT_LAUNCH(2);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------


extern "C" int dndm_egnn_forward(DndmEngine* e, const float* xh_lig, const float* xh_pocket, const float* t, int32_t t_len,
                                 const int64_t* lig_mask, const int64_t* pocket_mask, int32_t n_lig, int32_t n_pocket,
                                 int32_t n_samples, float* out_lig, float* out_pocket, void* stream) {
    if (!e || !xh_lig || !t || !lig_mask || !out_lig || (n_pocket > 0 && (!xh_pocket || !pocket_mask)))
        return set_err(DNDM_EINVAL, "null argument");
    if (!e->weights_loaded) return set_err(DNDM_EWEIGHTS, "dndm_engine_load_weights has not been called");
    if (t_len != 1 && t_len != n_samples) return set_err(DNDM_EINVAL, "t_len must be 1 or n_samples");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    RET_IF(prepare_batch(e, lig_mask, pocket_mask, n_lig, n_pocket, n_samples, st));
    const int N = n_lig + n_pocket;
    const int A = e->cfg.atom_nf, R = e->cfg.residue_nf;
    const int node_blocks = (N * 32 + 255) / 256;

    const bool prune_last = (out_pocket == nullptr) && n_pocket > 0;
    // ---- graph branch (coordinates only): gather x, pocket sums, radius graph, last-block edge list.  Forked onto the
    //      side stream (also under CUDA-graph capture: the fork/join events become graph edges); with per-section
    //      profiling on it stays on the main stream so that the categories time what they name. ----
    const bool fork = !e->profile;
    const bool pdl = !e->profile;                               // per-section timing wants kernels that do not overlap
    cudaStream_t gs = fork ? e->side : st;
    if (fork) {
        CU_CHECK(cudaEventRecord(e->ev_fork, st));              // after prepare_batch (sample pointers)
        CU_CHECK(cudaStreamWaitEvent(gs, e->ev_fork, 0));
    }
    {
        ProfScope ps(e, PROF_GRAPH, gs);
        RET_IF(build_graph(e, xh_lig, xh_pocket, e->xg, n_lig, N, n_samples, gs));
        if (fork) CU_CHECK(cudaEventRecord(e->ev_join, gs));   // blocks 0..L-2 only need the full graph
        // the compacted edge list is first read by the LAST block: it keeps running on the side stream under block 0
        if (prune_last) RET_IF(build_last_block_edges(e, n_lig, N, gs));
        if (fork && prune_last) CU_CHECK(cudaEventRecord(e->ev_join_last, gs));
        // off the critical path: the main stream only joins it at the very end of the call (so that the side stream is idle
        // -- and, under capture, joined -- when the call returns)
        RET_IF(refresh_pocket_lists(e, e->xg, n_lig, N, n_samples, gs));
        if (fork) CU_CHECK(cudaEventRecord(e->ev_join_pp, gs));
    }

    // ---- feature branch: encoder + embedding, first-layer projections of block 0's edge model (pq columns [0,512)) ----
    {
        ProfScope ps(e, PROF_NODE, st);
        const int lig_ctas = (n_lig + ENC_LIG_NODES_PER_CTA - 1) / ENC_LIG_NODES_PER_CTA;
        const int pok_ctas = (n_pocket + ENC_NODES_PER_CTA - 1) / ENC_NODES_PER_CTA;
        if (2 * A <= 32 && 2 * R <= 32)
            encode_embed_kernel<32><<<lig_ctas + pok_ctas, 256, 0, st>>>(xh_lig, xh_pocket, n_lig, N, 3 + A, 3 + R, t, t_len,
                                                                         e->node_sample, e->enc_l, e->enc_p, lig_ctas, e->x0, e->xa,
                                                                         e->xb, e->h, e->hcat);
        else    // C-alpha pockets: 20 residue types, encoder hidden width 40
            encode_embed_kernel<64><<<lig_ctas + pok_ctas, 256, 0, st>>>(xh_lig, xh_pocket, n_lig, N, 3 + A, 3 + R, t, t_len,
                                                                         e->node_sample, e->enc_l, e->enc_p, lig_ctas, e->x0, e->xa,
                                                                         e->xb, e->h, e->hcat);
        // per-sample sums of the (frozen) pocket coordinates for coord2cross: reads the coordinates the encoder just assembled;
        // on this stream, where the head of the step has slack (the graph branch is the longer one)
        pocket_sum_kernel<<<(n_samples * 32 + 255) / 256, 256, 0, st>>>(e->x0, e->pok_ptr, n_lig, n_samples, e->pocket_sum);
        COUNT_LAUNCH(2);
        if (e->h0_snap && N <= e->max_trace_nodes)
            CU_CHECK(cudaMemcpyAsync(e->h0_snap, e->h, (size_t)N * 256 * 4, cudaMemcpyDeviceToDevice, st));
    }
    float* x_cur = e->xa;
    float* x_next = e->xb;
    const float inv_norm = 1.0f / e->cfg.normalization_factor;
    {
        ProfScope ps(e, PROF_GEMM, st);
        WresEpilogue ep{e->layers[0].bias_e, nullptr, 0, nullptr, 0, 0, 1, 0};
        RET_IF(launch_wres(st, e->tm_hcat, e->layers[0].tm_we0, e->to_pq32, N, 2, 0, 0, ep));
    }
    if (fork) CU_CHECK(cudaStreamWaitEvent(st, e->ev_join, 0));   // join: the edge kernels need the graph
    for (int l = 0; l < e->cfg.n_layers; ++l) {
        LayerWeights& L = e->layers[l];
        // exact dead-work elimination (graph.cuh): the last block aggregates for the ligand rows and their pocket senders only,
        // the block before it for those and all their senders
        // pruning level of this block: 0 for the last block, 1 for the one before it, ...; -1 = all edges
        const int lvl_raw = e->cfg.n_layers - 1 - l;
        const int lvl = (prune_last && lvl_raw < prune_levels(e)) ? lvl_raw : -1;
        const bool pruned = lvl >= 0;
        if (pruned && lvl == prune_levels(e) - 1 && fork) CU_CHECK(cudaStreamWaitEvent(st, e->ev_join_last, 0));   // first pruned block
        // ---- GCL edge model + attention + deterministic aggregation ----
        EdgeGraph g{pruned ? e->erow_c[lvl] : e->erow, pruned ? e->ecol_c[lvl] : e->ecol, pruned ? e->r0_c[lvl] : e->r0, x_cur,
                    e->scalars + (pruned ? 2 + lvl : 0), n_lig, 1536, e->msg, e->att};
        EdgeProblem pe{e->pq, e->pq + 256, L.w1e_e, L.et_e, nullptr, L.att_bias, inv_norm};
        {
            ProfScope ps(e, PROF_GCL, st);
            // PDL: preceded by the merged projection GEMM (block 0) / coord_update (later blocks) on this stream
            auto gcl = e->edge_types ? (g_bf16_radial ? edge_pair_kernel<true, true, true> : edge_pair_kernel<true, false, true>)
                                     : (g_bf16_radial ? edge_pair_kernel<true, true> : edge_pair_kernel<true, false>);
            CU_CHECK(launch_pdl(pdl, gcl, dim3(e->num_sms & ~1, 1), dim3(EP_THREADS), EP_SMEM_BYTES, st, L.tm_w2_e, L.tm_w2_e, e->to_msg,
                                L.c_e, L.c_e, g, pe, pe));
        }
        {
            ProfScope ps(e, PROF_NODE, st);
            segment_reduce_kernel<<<node_blocks, 256, 0, st>>>(e->msg, e->att, pruned ? e->rp_act[lvl] : e->row_ptr, N, e->hcat);
        }
        COUNT_LAUNCH(2);
        // ---- node MLP with residual + node projections: this block's coordinate heads (sender parts for every node, receiver
        //      parts for the ligand rows) and the next block's edge model ----
        const bool has_next = l + 1 < e->cfg.n_layers;
        {
            ProfScope ps(e, PROF_GEMM, st);
            WresEpilogue ep1{L.b3, nullptr, 0, nullptr, 0, 0, 1, 0, 1};          // hid = SiLU(W3 [h | agg] + b3), K = 512
            RET_IF((launch_wres<512, 128>(st, e->tm_hcat, L.tm_w3, e->to_hid32, N, 2, 0, 0, ep1, 0, 0, pdl)));
            WresEpilogue ep2{L.b4, e->h, 256, e->h, 256, 0, 1, 0, 0, e->hcat, 512};   // h += W4 hid + b4 ; bf16 copy -> hcat[:, :256]
            RET_IF((launch_wres<256, 128>(st, e->tm_hid, L.tm_w4, e->to_hcat32, N, 2, 0, 0, ep2, 0, 0, pdl)));
        }
        {
            ProfScope ps(e, PROF_GEMM, st);
            WresEpilogue epm{L.bias_m, nullptr, 0, nullptr, 0, 0, 1, 0};
            // column groups 0-3 over all nodes (0,1 only when a next block exists), groups 4,5 over the ligand rows only
            RET_IF(launch_wres(st, e->tm_hcat, L.tm_wm, e->to_pq32, N, has_next ? 4 : 2, has_next ? 0 : 2, 0, epm, 2, n_lig, pdl));
        }
        // ---- EquivariantUpdate: two scalar heads on the ligand-receiver edges, then the coordinate update ----
        {
            EdgeGraph gh{e->erow, e->ecol, e->r0, x_cur, e->scalars + 1, n_lig, 1536, nullptr, nullptr};
            EdgeProblem pc{e->pq + 1024, e->pq + 512, L.w1e_c, L.et_c, e->phi, 0.f, e->cfg.coords_range};
            EdgeProblem px{e->pq + 1280, e->pq + 768, L.w1e_x, L.et_x, e->psi, 0.f, e->cfg.coords_range};
            const int gx = e->num_sms / 2 > 0 ? e->num_sms / 2 : 1;
            {
                ProfScope ps(e, PROF_HEAD, st);
                // CTA pairs: an even number of CTAs per problem
                CU_CHECK(launch_pdl(pdl, e->edge_types ? edge_pair_kernel<false, true, true> : edge_pair_kernel<false>,
                                    dim3(gx > 1 ? gx & ~1 : 2, 2), dim3(EP_THREADS), EP_SMEM_BYTES, st, L.tm_w2_c, L.tm_w2_x, e->to_msg,
                                    L.c_c, L.c_x, gh, pc, px));
            }
            ProfScope ps2(e, PROF_NODE, st);
            coord_update_kernel<<<(n_lig * 32 + 255) / 256, 256, 0, st>>>(x_cur, x_next, e->row_ptr, e->ecol, e->phi, e->psi,
                                                                          e->node_sample, e->lig_ptr, e->pok_ptr,
                                                                          e->pocket_sum, n_lig, e->cfg.norm_constant, inv_norm);
            COUNT_LAUNCH(2);
            float* tmp = x_cur; x_cur = x_next; x_next = tmp;
        }
        if (e->h_trace && N <= e->max_trace_nodes)
            CU_CHECK(cudaMemcpyAsync(e->h_trace + (size_t)l * e->max_trace_nodes * 256, e->h, (size_t)N * 256 * 4,
                                     cudaMemcpyDeviceToDevice, st));
        if (e->x_trace && N <= e->max_trace_nodes)
            CU_CHECK(cudaMemcpyAsync(e->x_trace + (size_t)l * e->max_trace_nodes * 3, x_cur, (size_t)N * 3 * 4,
                                     cudaMemcpyDeviceToDevice, st));
        CU_CHECK(cudaGetLastError());
    }
    if (fork) CU_CHECK(cudaStreamWaitEvent(st, e->ev_join_pp, 0));
    // ---- embedding_out + decoders + velocity ----
    {
        const int n_out = out_pocket ? N : n_lig;
        ProfScope ps(e, PROF_NODE, st);
        decode_kernel<<<(n_out * 32 + 255) / 256, 256, 0, st>>>(e->h, x_cur, e->x0, n_lig, n_out, 0, e->dec_l, e->dec_p, out_lig,
                                                                out_pocket, e->flags);
        COUNT_LAUNCH(1);
    }
    CU_CHECK(cudaGetLastError());
    e->last_n_lig = n_lig;
    e->last_n_nodes = N;
    e->x_final = x_cur;
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// radius graph on its own
// ------------------------------------------------------------------------------------------------

extern "C" int dndm_radius_graph(DndmEngine* e, const float* xh_lig, const float* xh_pocket, const int64_t* lig_mask,
                                 const int64_t* pocket_mask, int32_t n_lig, int32_t n_pocket, int32_t n_samples,
                                 int32_t* row_ptr, int32_t* col, int32_t edge_capacity, int32_t* n_edges_host, void* stream) {
    if (!e || !xh_lig || !lig_mask) return set_err(DNDM_EINVAL, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    RET_IF(prepare_batch(e, lig_mask, pocket_mask, n_lig, n_pocket, n_samples, st));
    const int N = n_lig + n_pocket;
    RET_IF(build_graph(e, xh_lig, xh_pocket, e->x0, n_lig, N, n_samples, st));
    RET_IF(refresh_pocket_lists(e, e->x0, n_lig, N, n_samples, st));
    int sc[2] = {0, 0};
    CU_CHECK(cudaMemcpyAsync(sc, e->scalars, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    if (n_edges_host) *n_edges_host = sc[0];
    if (row_ptr) CU_CHECK(cudaMemcpyAsync(row_ptr, e->row_ptr, (size_t)(N + 1) * 4, cudaMemcpyDeviceToDevice, st));
    if (col) {
        const int n = sc[0] < edge_capacity ? sc[0] : edge_capacity;
        CU_CHECK(cudaMemcpyAsync(col, e->ecol, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    e->last_n_lig = n_lig;
    e->last_n_nodes = N;
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// sampler step
// ------------------------------------------------------------------------------------------------
extern "C" int dndm_sampler_step(DndmEngine* e, const float* z_in, const float* eps, const float* noise,
                                 const float* xh_pocket_in, const float* coef, const float* grad, float lambda,
                                 const int64_t* lig_mask, const int64_t* pocket_mask, int32_t n_lig, int32_t n_pocket,
                                 int32_t n_samples, float* z_out, float* xh_pocket_out, int32_t check_input_com, void* stream) {
    if (!e || !z_in || !noise || !coef || !lig_mask || !z_out) return set_err(DNDM_EINVAL, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    RET_IF(prepare_batch(e, lig_mask, pocket_mask, n_lig, n_pocket, n_samples, st));
    sampler_step_kernel<<<n_samples, 128, 0, st>>>(z_in, eps ? eps : noise, noise, xh_pocket_in, coef, grad, lambda, e->lig_ptr,
                                                   e->pok_ptr, e->cfg.atom_nf, e->cfg.residue_nf, z_out, xh_pocket_out, e->flags, check_input_com != 0);
    COUNT_LAUNCH(1);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// bond perception (pre-filter of candidate molecules)
// ------------------------------------------------------------------------------------------------
extern "C" int dndm_bond_orders(DndmEngine* e, const float* x, int32_t ld_x, const int64_t* atom_type, const int64_t* mol_mask,
                                int32_t n_atoms, int32_t n_mols, const float* bonds1, const float* bonds2, const float* bonds3,
                                int32_t n_types, float margin1, float margin2, float margin3, const int32_t* allowed_valence,
                                int8_t* e_out, int64_t e_capacity, int32_t* valence_out, int32_t* mol_stats, void* stream) {
    if (!e || !x || !atom_type || !mol_mask || !bonds1 || !bonds2 || !bonds3 || !valence_out || !mol_stats)
        return set_err(DNDM_EINVAL, "null argument");
    if (n_atoms < 1 || n_mols < 1 || n_types < 1 || ld_x < 3) return set_err(DNDM_EINVAL, "bad sizes");
    if (n_atoms > e->cfg.max_nodes || n_mols > e->cfg.max_samples)
        return set_err(DNDM_ECAPACITY, "batch (%d atoms, %d molecules) exceeds engine capacity (%d, %d)", n_atoms, n_mols,
                       e->cfg.max_nodes, e->cfg.max_samples);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    e->last_lm = e->last_pm = nullptr;                       // lig_ptr / node_sample are reused as scratch below
    {
        const int n = n_atoms > n_mols + 1 ? n_atoms : n_mols + 1;
        mask_to_ptr_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(mol_mask), n_atoms, n_mols,
                                                            e->lig_ptr, e->node_sample);
    }
    bond_offsets_kernel<<<1, 32, 0, st>>>(e->lig_ptr, n_mols, e->mol_off);
    BondTables tb{bonds1, bonds2, bonds3, allowed_valence, n_types, margin1, margin2, margin3};
    bond_orders_kernel<<<n_mols, BOND_THREADS, 0, st>>>(x, ld_x, reinterpret_cast<const long long*>(atom_type), e->lig_ptr, tb,
                                                        e->mol_off, reinterpret_cast<signed char*>(e_out), e_capacity,
                                                        valence_out, mol_stats, e->flags);
    COUNT_LAUNCH(3);
    CU_CHECK(cudaGetLastError());
    return DNDM_OK;
}

extern "C" int dndm_read_flags(DndmEngine* e, uint32_t* flags_host, void* stream) {
    if (!e || !flags_host) return set_err(DNDM_EINVAL, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CU_CHECK(cudaMemcpyAsync(flags_host, e->flags, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemsetAsync(e->flags, 0, 4, st));
    CU_CHECK(cudaStreamSynchronize(st));
    return DNDM_OK;
}

extern "C" int64_t dndm_debug_copy(DndmEngine* e, int32_t what, void* dst, int64_t dst_bytes, void* stream) {
    if (!e || !dst) return set_err(DNDM_EINVAL, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const void* src = nullptr;
    int64_t bytes = 0;
    const int64_t N = e->last_n_nodes;
    switch (what) {
        case 0: src = e->h; bytes = N * 256 * 4; break;
        case 1: src = e->x_final ? e->x_final : e->x0; bytes = N * 3 * 4; break;
        case 2: src = e->row_ptr; bytes = (N + 1) * 4; break;
        case 3: {
            int sc[2];
            CU_CHECK(cudaMemcpyAsync(sc, e->scalars, 8, cudaMemcpyDeviceToHost, st));
            CU_CHECK(cudaStreamSynchronize(st));
            src = e->ecol; bytes = (int64_t)sc[0] * 4;
            break;
        }
        case 4: src = e->scalars; bytes = 32; break;
        case 7: src = e->pp.meta; bytes = 32; break;
        case 8:
            if (!e->h0_snap) return set_err(DNDM_EINVAL, "buffer 8 (h_0) exists only while a trace is set");
            src = e->h0_snap; bytes = N * 256 * 4; break;
#ifdef DNDM_EK_TRACE
        case 6: {
            bytes = sizeof(g_wr_trace) < (size_t)dst_bytes ? sizeof(g_wr_trace) : dst_bytes;
            CU_CHECK(cudaMemcpyFromSymbolAsync(dst, g_wr_trace, bytes, 0, cudaMemcpyDeviceToDevice, st));
            return bytes;
        }
        case 5: {
            bytes = sizeof(g_ek_trace) < (size_t)dst_bytes ? sizeof(g_ek_trace) : dst_bytes;
            CU_CHECK(cudaMemcpyFromSymbolAsync(dst, g_ek_trace, bytes, 0, cudaMemcpyDeviceToDevice, st));
            return bytes;
        }
#endif
        default: return set_err(DNDM_EINVAL, "unknown buffer id %d", what);
    }
    if (bytes > dst_bytes) bytes = dst_bytes;
    CU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    return bytes;
}

extern "C" int dndm_set_static_masks(DndmEngine* e, int32_t on) {
    if (!e) return set_err(DNDM_EINVAL, "null argument");
    e->static_masks = on != 0;
    e->last_lm = e->last_pm = nullptr;
    e->last_nl = e->last_np = e->last_ns = -1;
    return DNDM_OK;
}

extern "C" int dndm_set_profile(DndmEngine* e, int32_t on) {
    if (!e) return set_err(DNDM_EINVAL, "null argument");
    e->profile = on != 0;
    e->sections.clear();
    e->ev_used = 0;
    return DNDM_OK;
}

extern "C" int dndm_get_profile(DndmEngine* e, double* ms_per_category, int32_t* launches_per_category, int32_t n_cat) {
    if (!e || !ms_per_category || !launches_per_category) return set_err(DNDM_EINVAL, "null argument");
    CU_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < n_cat; ++i) { ms_per_category[i] = 0.0; launches_per_category[i] = 0; }
    for (auto& s : e->sections) {
        float ms = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, s.a, s.b));
        if (s.cat < n_cat) { ms_per_category[s.cat] += ms; launches_per_category[s.cat] += 1; }
    }
    e->sections.clear();
    e->ev_used = 0;
    return DNDM_OK;
}

extern "C" int dndm_set_trace(DndmEngine* e, float* h_trace, float* x_trace, int32_t max_trace_nodes) {
    if (!e) return set_err(DNDM_EINVAL, "null argument");
    e->h_trace = h_trace; e->x_trace = x_trace; e->max_trace_nodes = max_trace_nodes;
    // while a trace is on, the encoder output h_0 of every forward is kept as well (dndm_debug_copy buffer 8)
    if (e->h0_snap) { cudaFree(e->h0_snap); e->h0_snap = nullptr; }
    if (h_trace && max_trace_nodes > 0) RET_IF(dev_alloc(&e->h0_snap, (size_t)max_trace_nodes * 256));
    return DNDM_OK;
}

// ------------------------------------------------------------------------------------------------
// GEMM self-test: the weight-resident node GEMM exactly as the forward launches it
// ------------------------------------------------------------------------------------------------
extern "C" int dndm_test_gemm(const void* a_bf16, const void* w_bf16, const float* bias, const float* residual, int32_t act,
                              int32_t M, int32_t N, int32_t K, int32_t bn, int32_t n_tail_groups, int32_t m_tail,
                              float* out_f32, void* out_bf16, void* stream) {
    if (!a_bf16 || !w_bf16 || (!out_f32 && !out_bf16)) return set_err(DNDM_EINVAL, "null argument");
    const bool s256 = (K == 256 && bn == 256), s512 = (K == 512 && bn == 128), s128 = (K == 256 && bn == 128);
    if (!(s256 || s512 || s128)) return set_err(DNDM_EINVAL, "supported shapes: (K, BN) = (256, 256), (512, 128), (256, 128)");
    if (M < 1 || N % bn != 0 || n_tail_groups < 0 || n_tail_groups * bn >= N + (n_tail_groups ? 0 : 1) || m_tail > M)
        return set_err(DNDM_EINVAL, "bad sizes");
    if (residual && N != 256) return set_err(DNDM_EINVAL, "the residual epilogue is defined for 256 output columns");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
            g_num_sms = sms;
    }
    const uint64_t rows = (uint64_t)((M + WR_BM - 1) / WR_BM) * WR_BM;      // the caller pads A / out_bf16 to whole row blocks
    CUtensorMap ta, tw, to;
    RET_IF(make_tmap_bf16(&ta, a_bf16, rows, K, K, WR_BM));
    RET_IF(make_tmap_bf16(&tw, w_bf16, N, K, K, bn / 2));
    if (out_bf16) RET_IF(make_tmap_bf16_box(&to, out_bf16, rows, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    else to = ta;
    CU_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<256, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, WresShape<256, 256>::smem_bytes));
    CU_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<512, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WresShape<512, 128>::smem_bytes));
    CU_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<256, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WresShape<256, 128>::smem_bytes));
    WresEpilogue ep{bias, residual, 256, out_f32, N, 0, out_bf16 ? 1 : 0, 0, act, reinterpret_cast<__nv_bfloat16*>(out_bf16), N};
    if (residual && !s128) return set_err(DNDM_EINVAL, "the residual epilogue exists for (K, BN) = (256, 128) only");
    const int n_full = N / bn - n_tail_groups;
    if (s256) return launch_wres<256, 256>(st, ta, tw, to, M, n_full, 0, 0, ep, n_tail_groups, m_tail);
    if (s512) return launch_wres<512, 128>(st, ta, tw, to, M, n_full, 0, 0, ep, n_tail_groups, m_tail);
    return launch_wres<256, 128>(st, ta, tw, to, M, n_full, 0, 0, ep, n_tail_groups, m_tail);
}
