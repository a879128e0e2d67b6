// admm.cu — host side of path 2 behind the C ABI (include/sequnwinder_b200.h): owns the device
// state of one LOne instance, enqueues the kernels of admm_kernels.cuh and runs the outer
// label-consensus loop of LOne.execute (framework/LOne.java:116-204).
//
// The host never waits inside an ADMM iteration: iterations are enqueued ahead of the GPU and a
// small pinned ring of Ctrl snapshots tells the host, a few iterations late, that the device has
// set Ctrl.done; the kernels of the surplus iterations return immediately.
#include "admm_kernels.cuh"
#include "common.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace sqw {

// ------------------------------------------------------------------ NCCL through dlopen
struct NcclUid {
    char internal[128];
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUid *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUid, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib)
                break;
        }
        if (api.lib) {
            api.GetUniqueId = (int (*)(NcclUid *))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void **, int, NcclUid, int))dlsym(api.lib, "ncclCommInitRank");
            api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *,
                                     cudaStream_t))dlsym(api.lib, "ncclAllReduce");
            api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
            api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
        }
    });
    if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce)
        return nullptr;
    return &api;
}

constexpr int kNcclDouble = 8, kNcclSum = 0;

// java.util.HashMap iteration order of keys "Thread0".."Thread{T-1}" (LOne.java:399,411,427):
// the order in which the reference sums the per-thread buffers.
static std::vector<int> java_thread_order(int T)
{
    int cap = 16;
    while (T > (cap * 3) / 4)
        cap *= 2;
    std::vector<int> order;
    for (int b = 0; b < cap; ++b)
        for (int t = 0; t < T; ++t) {
            const std::string key = "Thread" + std::to_string(t);
            uint32_t h = 0;
            for (char ch : key)
                h = 31u * h + (uint32_t)(unsigned char)ch;
            h ^= (h >> 16);
            if ((int)(h & (uint32_t)(cap - 1)) == b)
                order.push_back(t);
        }
    return order;
}

constexpr int kNumPhases = 6;  // x-update launches per ADMM iteration (re-grouping points)
constexpr int kRing = 8;       // Ctrl snapshots in flight
#ifndef SQW_ADMM_FUSED_DEFAULT
#define SQW_ADMM_FUSED_DEFAULT 1
#endif
constexpr int kLookahead = 3;  // iterations the host runs ahead of the snapshot it inspects

}  // namespace sqw

using namespace sqw;

struct sqw_admm {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;

    // params
    bool have_params = false, have_structure = false, have_problem = false, prepared = false;
    bool l2_persist = false;  // this handle set the device's persisting-L2 carve-out
    int cta_budget = 0;       // 0 = all SMs; else at most this many CTAs per x-update launch
    // iteration tail on one GPU: one fused launch unless SQW_ADMM_UNFUSED is set (test hook)
    bool unfused_tail = SQW_ADMM_FUSED_DEFAULT ? getenv("SQW_ADMM_UNFUSED") != nullptr
                                               : getenv("SQW_ADMM_FUSED") == nullptr;
    double ridge = 10.0, pho = 1.7;
    int T = 5, admm_max_itr = 400, nodes_max_itr = 15, debug = 0;

    // comm
    int rank = 0, world = 1;
    void *comm = nullptr;

    // structure (host)
    int NN = 0, num_layers = 0, C_struct = 0;
    std::vector<int> edge_leaf, edge_par, leaf_edge_ptr, edge_sum_order;
    struct Layer {
        std::vector<int> nodes, nbr_ptr, nbr_idx;
        int *d_nodes = nullptr, *d_ptr = nullptr, *d_idx = nullptr;
    };
    std::vector<Layer> layers;  // index = layer number (0 unused)

    // problem
    int64_t n_rows = 0;
    int dim = 0, C = 0;
    std::vector<int> block_ids;

    AdmmDev dev{};
    Ctrl *h_ring = nullptr;  // pinned [kRing]
    cudaEvent_t ev_x0[kRing], ev_x1[kRing], ev_snap[kRing];
    cudaEvent_t ev_a, ev_b, ev_t0, ev_t1;
    // SQW_ADMM_TAIL_PROFILE (multi-GPU): events between the kernels / all-reduces of the iteration
    // tail, so that the time an iteration loses to the collective (incl. the wait for the slowest
    // rank's x-update) is measured, not inferred
    bool tail_prof = false;
    cudaEvent_t ev_tail[kRing][5];
    double tail_ms[6] = {0, 0, 0, 0, 0, 0};   // x-update, pack, all-reduce 1, consensus, all-reduce 2, decide
    // evaluation caps of the x-update phases (cumulative; the last phase runs every block to
    // completion); tuned at cfg-2 (mean 14.7, max 24 evaluations per x-update)
    int caps[kNumPhases] = {8, 12, 16, 20, 26, 0x7fffffff};
    int nphase = kNumPhases;   // launches per x-update (1 when the rank owns a single block)
    int tile_rows = 8;         // rows per tile of the evaluation path in use
    int eval_mode = -1;        // 0 TMA tiles, 1 streaming + W in smem, 2 streaming + W global
    int force_mode = -1;       // test hook (SQW_ADMM_EVAL_MODE)
    size_t xupdate_smem = 0;
    TileSmemLayout layout{};

    // results
    sqw_admm_stats stats{};
    std::vector<int> log_outer, log_itr, log_nfev, admm_iters;
    std::vector<double> log_pho, log_fold, log_primal, log_dual;

    template <typename Tp>
    int dalloc(Tp **p, size_t count, bool zero = true)
    {
        // stream-ordered allocation from the device's default pool (its release threshold is raised
        // in sqw_admm_create): a handle's ~50 buffers come out of memory the previous handle gave
        // back, where cudaMalloc / cudaFree cost the e2e step 0.1 s per training (each cudaFree
        // synchronises the device)
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(Tp);
        SQW_CUDA_CHECK(cudaMallocAsync((void **)p, bytes, stream));
        allocs.push_back((void *)*p);
        if (zero)
            SQW_CUDA_CHECK(cudaMemsetAsync(*p, 0, bytes, stream));
        return SQW_OK;
    }
};

namespace sqw {

typedef void (*XKernel)(AdmmDev, TileSmemLayout, int, int);
typedef void (*EKernel)(AdmmDev, TileSmemLayout, double, double *);

static XKernel xupdate_kernel(int nct, int mode)
{
    switch (nct * 16 + mode) {
    case 16 + 4: return admm_xupdate_kernel<1, 4>;
    case 16 + 6: return admm_xupdate_kernel<1, 6>;
    case 16 + 5: return admm_xupdate_kernel<1, 5>;
    case 16 + 8: return admm_xupdate_kernel<1, 8>;
    case 16 + 9: return admm_xupdate_kernel<1, 9>;
    case 16 + 10: return admm_xupdate_kernel<1, 10>;
    case 32 + 10: return admm_xupdate_kernel<2, 10>;
    case 48 + 10: return admm_xupdate_kernel<3, 10>;
    case 16 + 11: return admm_xupdate_kernel<1, 11>;
    case 16 + 12: return admm_xupdate_kernel<1, 12>;
    case 32 + 11: return admm_xupdate_kernel<2, 11>;
    case 16 + 1: return admm_xupdate_kernel<1, 1>;
    case 16 + 2: return admm_xupdate_kernel<1, 2>;
    case 32 + 4: return admm_xupdate_kernel<2, 4>;
    case 32 + 6: return admm_xupdate_kernel<2, 6>;
    case 32 + 1: return admm_xupdate_kernel<2, 1>;
    case 32 + 2: return admm_xupdate_kernel<2, 2>;
    case 48 + 4: return admm_xupdate_kernel<3, 4>;
    case 48 + 6: return admm_xupdate_kernel<3, 6>;
    case 48 + 1: return admm_xupdate_kernel<3, 1>;
    case 48 + 2: return admm_xupdate_kernel<3, 2>;
    }
    return nullptr;
}

static EKernel eval_kernel(int nct, int mode)
{
    switch (nct * 16 + mode) {
    case 16 + 4: return admm_eval_kernel<1, 4>;
    case 16 + 6: return admm_eval_kernel<1, 6>;
    case 16 + 5: return admm_eval_kernel<1, 5>;
    case 16 + 8: return admm_eval_kernel<1, 8>;
    case 16 + 9: return admm_eval_kernel<1, 9>;
    case 16 + 10: return admm_eval_kernel<1, 10>;
    case 32 + 10: return admm_eval_kernel<2, 10>;
    case 48 + 10: return admm_eval_kernel<3, 10>;
    case 16 + 11: return admm_eval_kernel<1, 11>;
    case 16 + 12: return admm_eval_kernel<1, 12>;
    case 32 + 11: return admm_eval_kernel<2, 11>;
    case 16 + 1: return admm_eval_kernel<1, 1>;
    case 16 + 2: return admm_eval_kernel<1, 2>;
    case 32 + 4: return admm_eval_kernel<2, 4>;
    case 32 + 6: return admm_eval_kernel<2, 6>;
    case 32 + 1: return admm_eval_kernel<2, 1>;
    case 32 + 2: return admm_eval_kernel<2, 2>;
    case 48 + 4: return admm_eval_kernel<3, 4>;
    case 48 + 6: return admm_eval_kernel<3, 6>;
    case 48 + 1: return admm_eval_kernel<3, 1>;
    case 48 + 2: return admm_eval_kernel<3, 2>;
    }
    return nullptr;
}

// class tiles of the kernel instantiation: MODE 11 runs C = 8 n + 1 classes on n DMMA class tiles
static int kernel_nct(const sqw_admm *h) { return h->eval_mode == 11 ? (h->C - 1) / 8 : h->dev.Cp / 8; }
static int mode_threads(int mode) { return mode == 9 ? ModeCfg<9>::NT : kThreads; }
static int mode_ctas_per_sm(int mode) { return mode == 9 ? ModeCfg<9>::MINB : 1; }

static int upload_ints(sqw_admm *h, const std::vector<int> &v, const int **dst)
{
    int *p = nullptr;
    int rc = h->dalloc(&p, v.size(), false);
    if (rc != SQW_OK)
        return rc;
    if (!v.empty())
        SQW_CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice,
                                       h->stream));
    *dst = p;
    return SQW_OK;
}

// CTAs per ADMM block: G while every local block is active (one CTA per SM, all T_local groups
// co-resident) and the cap a re-grouped block may grow to in a later phase. A launch never gives
// a block fewer than min(G, gcap) CTAs.
static void group_sizes(const sqw_admm *h, int n_sm, int T_local, int nb, int &G, int &gcap)
{
    // CTA budget of this handle (sqw_admm_set_cta_budget): several handles with disjoint budgets run
    // their cooperative x-update launches side by side on one GPU (concurrent sweep)
    if (h->cta_budget > 0)
        n_sm = std::min(n_sm, h->cta_budget);
    G = std::max(1, n_sm / T_local);
    const int max_useful = std::max(1, (nb + 7) / 8);
    G = std::min(G, max_useful);
    // measured at cfg-2 (2 500 rows per block; tools/gcap.sh): 64.5 us per evaluation round
    // without a cap, 62.0 with 48, 63.6 with 36, 66.8 with 24. A trip costs ~ rows/G of evaluation
    // plus ~ G of partial-gradient exchange (both scale with C * dim), so the best G grows with
    // the square root of the block's rows: 48 at 2 500 rows, no cap in effect for large blocks.
    // SQW_ADMM_GCAP is the tuning hook.
    const int gcap_auto = std::max(16, (int)(0.96 * std::sqrt((double)nb) + 0.5));
    gcap = getenv("SQW_ADMM_GCAP") ? std::max(1, atoi(getenv("SQW_ADMM_GCAP"))) : gcap_auto;
    if (T_local == 1)
        G = std::min(G, gcap);
}

// Allocate everything that depends on params + structure + problem sizes.
static int prepare(sqw_admm *h)
{
    if (h->prepared)
        return SQW_OK;
    if (!h->have_params || !h->have_structure || !h->have_problem)
        return set_error(SQW_ERR_STATE, "set_structure, set_params and set_problem must precede this call");
    if (h->C_struct != h->C)
        return set_error(SQW_ERR_INVALID_ARG, "structure has %d leaves but num_classes is %d",
                         h->C_struct, h->C);
    AdmmDev &D = h->dev;
    D.C = h->C;
    D.Cp = (h->C + 7) / 8 * 8;
    D.E = (int)h->edge_leaf.size();
    D.NN = h->NN;
    D.npad = D.C * D.dimp;
    D.upad = D.E * D.dimp;
    D.wstride = D.dimp;
    D.world = h->world;
    D.ridge = h->ridge;
    D.n_sm = 148;
    // MODE 9 probe: start delay of the second CTA of every SM (measured: no effect, see plan_blocks)
    D.skew_ns = getenv("SQW_ADMM_SKEW_NS") ? atoi(getenv("SQW_ADMM_SKEW_NS")) : 0;
    const int nct = kernel_nct(h);
    if ((long long)D.Cp * D.dimp * h->dev.T_local > (1ll << 31) - 1 || (long long)D.upad > (1 << 30))
        return set_error(SQW_ERR_UNSUPPORTED, "problem too large for 32-bit coordinate indices");

    cudaDeviceProp prop;
    SQW_CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
    int n_sm = prop.multiProcessorCount;
    D.n_sm = prop.multiProcessorCount;
    if (h->cta_budget > 0)
        n_sm = std::min(n_sm, h->cta_budget);
    const int mode = h->eval_mode;
    const size_t w_bytes = (size_t)D.Cp * D.wstride * sizeof(double);
    if (mode == 4 || mode == 5 || mode == 6 || mode == 8 || mode == 9 || mode == 10 || mode == 11 || mode == 12)
        h->xupdate_smem = h->layout.total;
    else if (mode == 1)
        h->xupdate_smem = w_bytes;
    else
        h->xupdate_smem = 0;
    XKernel xk = xupdate_kernel(nct, mode);
    EKernel ek = eval_kernel(nct, mode);
    // The attribute is per function and process-wide: handles of different shapes (the folds of a
    // cross-validation, two resident optimizers) share the kernels, so it is always raised to the
    // device's opt-in maximum and each launch passes its own size.
    // (dynamic + static shared memory of a kernel may not exceed the opt-in limit)
    for (const void *fn : {(const void *)xk, (const void *)ek}) {
        cudaFuncAttributes fa;
        SQW_CUDA_CHECK(cudaFuncGetAttributes(&fa, fn));
        const size_t dyn_max = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
        if (h->xupdate_smem > dyn_max)
            return set_error(SQW_ERR_UNSUPPORTED, "x-update needs %zu bytes of shared memory (max %zu)",
                             h->xupdate_smem, dyn_max);
        SQW_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
    }
    int per_sm = 0;
    SQW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xk, mode_threads(mode),
                                                                 h->xupdate_smem));
    if (per_sm < mode_ctas_per_sm(mode))
        return set_error(SQW_ERR_UNSUPPORTED, "x-update kernel does not fit on an SM (%d of %d CTAs)",
                         per_sm, mode_ctas_per_sm(mode));
    per_sm = mode_ctas_per_sm(mode);
    // one CTA per SM: G CTAs per ADMM block, all T_local groups co-resident
    if (D.T_local > n_sm * per_sm)
        return set_error(SQW_ERR_UNSUPPORTED, "%d local ADMM blocks exceed %d co-resident CTAs",
                         D.T_local, n_sm * per_sm);
    int G = 1;
    group_sizes(h, prop.multiProcessorCount * per_sm, D.T_local, D.nb, G, D.gcap);
    // a rank that owns a single block has nothing to re-group: one launch, run to completion
    h->nphase = (D.T_local == 1) ? 1 : kNumPhases;
    D.nphase = h->nphase;
    D.G = G;
    D.Gmax = G * D.T_local;

    int rc;
#define TRY(x)            \
    do {                  \
        rc = (x);         \
        if (rc != SQW_OK) \
            return rc;    \
    } while (0)
    TRY(upload_ints(h, h->edge_leaf, &D.edge_leaf));
    TRY(upload_ints(h, h->edge_par, &D.edge_par));
    TRY(upload_ints(h, h->leaf_edge_ptr, &D.leaf_edge_ptr));
    TRY(upload_ints(h, h->edge_sum_order, &D.edge_sum_order));
    TRY(upload_ints(h, h->block_ids, &D.block_ids));
    for (auto &L : h->layers) {
        const int *p;
        TRY(upload_ints(h, L.nodes, &p)); L.d_nodes = (int *)p;
        TRY(upload_ints(h, L.nbr_ptr, &p)); L.d_ptr = (int *)p;
        TRY(upload_ints(h, L.nbr_idx, &p)); L.d_idx = (int *)p;
    }
    const size_t TL = (size_t)D.T_local;
    TRY(h->dalloc(&D.xbar, (size_t)D.npad));
    TRY(h->dalloc(&D.ubar, (size_t)D.upad));
    TRY(h->dalloc(&D.z, (size_t)D.upad));
    TRY(h->dalloc(&D.zold, (size_t)D.upad));
    TRY(h->dalloc(&D.smx, (size_t)D.NN * D.dimp));
    TRY(h->dalloc(&D.smx_old, (size_t)D.NN * D.dimp));
    TRY(h->dalloc(&D.xt, TL * D.Cp * D.dimp));
    TRY(h->dalloc(&D.ut, TL * D.upad));
    TRY(h->dalloc(&D.g, TL * D.npad));
    TRY(h->dalloc(&D.gold, TL * D.npad));
    TRY(h->dalloc(&D.d, TL * D.npad));
    TRY(h->dalloc(&D.wa, TL * D.npad));
    TRY(h->dalloc(&D.S, TL * kLbfgsM * D.npad));
    TRY(h->dalloc(&D.Y, TL * kLbfgsM * D.npad));
    TRY(h->dalloc(&D.R, TL * D.nb * D.Cp));
    TRY(h->dalloc(&D.gpart, TL * D.Gmax * D.npad));
    TRY(h->dalloc(&D.spart, TL * D.Gmax * kNumDots));
    TRY(h->dalloc(&D.bstate, TL));
    TRY(h->dalloc(&D.bar, TL * 32 * kNumPhases));
    TRY(h->dalloc(&D.nfev, TL));
    if (getenv("SQW_ADMM_PROFILE"))
        TRY(h->dalloc(&D.prof, (size_t)D.Gmax * kProfSlots * kNumPhases));
    TRY(h->dalloc(&D.sum1, (size_t)2 * D.npad + D.upad));
    TRY(h->dalloc(&D.sum2, (size_t)D.E));
    TRY(h->dalloc(&D.edge_stats, (size_t)D.E * 4));
    TRY(h->dalloc(&D.leaf_xnorm, (size_t)D.C));
    TRY(h->dalloc(&D.ctrl, 1));
    const size_t cap = (size_t)std::max(1, h->admm_max_itr) * std::max(1, h->nodes_max_itr);
    TRY(h->dalloc(&D.log_outer, cap));
    TRY(h->dalloc(&D.log_itr, cap));
    TRY(h->dalloc(&D.log_nfev, cap * TL));
    TRY(h->dalloc(&D.log_pho, cap));
    TRY(h->dalloc(&D.log_fold, cap));
    TRY(h->dalloc(&D.log_primal, cap));
    TRY(h->dalloc(&D.log_dual, cap));
#undef TRY
    // Optional (SQW_ADMM_L2_PERSIST_MB=<MB>): a persisting access-policy window over the training
    // matrix, hitRatio = persisting bytes / window bytes. Every evaluation re-reads the same X and
    // a cyclic sweep over a working set close to the L2 size thrashes (ncu: 23 % hit rate at
    // cfg-2), but pinning part of it made no measurable difference to the x-update, and the
    // carve-out is a device-wide setting that slows other kernels (the k-mer pass ran 4x slower
    // next to a resident trainer), so it is off unless asked for.
    if (const char *env = D.X ? getenv("SQW_ADMM_L2_PERSIST_MB") : nullptr) {
        const size_t x_bytes = (size_t)D.T_local * D.nb * D.dimp * sizeof(double);
        size_t persist = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)atol(env) << 20);
        if (persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
            h->l2_persist = true;
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            const size_t win = std::min(x_bytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.base_ptr = const_cast<double *>(D.X);
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)win);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess)
                cudaGetLastError();
            if (h->debug)
                fprintf(stderr, "[sqw] L2 persisting window: %zu MB of X (%zu MB), hit ratio %.2f\n",
                        persist >> 20, x_bytes >> 20, attr.accessPolicyWindow.hitRatio);
        } else {
            cudaGetLastError();
        }
    }
    SQW_CUDA_CHECK(cudaMallocHost((void **)&h->h_ring, sizeof(Ctrl) * kRing));
    for (int i = 0; i < kRing; ++i) {
        SQW_CUDA_CHECK(cudaEventCreate(&h->ev_x0[i]));
        SQW_CUDA_CHECK(cudaEventCreate(&h->ev_x1[i]));
        SQW_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_snap[i], cudaEventDisableTiming));
    }
    SQW_CUDA_CHECK(cudaEventCreate(&h->ev_a));
    SQW_CUDA_CHECK(cudaEventCreate(&h->ev_b));
    SQW_CUDA_CHECK(cudaEventCreate(&h->ev_t0));
    SQW_CUDA_CHECK(cudaEventCreate(&h->ev_t1));
    if (h->world > 1 && getenv("SQW_ADMM_TAIL_PROFILE")) {
        for (int i = 0; i < kRing; ++i)
            for (int j = 0; j < 5; ++j)
                SQW_CUDA_CHECK(cudaEventCreate(&h->ev_tail[i][j]));
        h->tail_prof = true;
    }
    h->prepared = true;
    return SQW_OK;
}

static int launch_xupdate(sqw_admm *h)
{
    AdmmDev &D = h->dev;
    XKernel xk = xupdate_kernel(kernel_nct(h), h->eval_mode);
    for (int ph = 0; ph < h->nphase; ++ph) {
        int phase = ph, cap = (ph == h->nphase - 1) ? 0x7fffffff : h->caps[ph];
        void *args[] = {&D, &h->layout, &phase, &cap};
        SQW_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)xk, dim3(D.Gmax), dim3(mode_threads(h->eval_mode)),
                                                   args, h->xupdate_smem, h->stream));
    }
    h->stats.launches += h->nphase - 1;
    return SQW_OK;
}

static int nccl_allreduce(sqw_admm *h, double *buf, size_t count)
{
    if (h->world <= 1)
        return SQW_OK;
    NcclApi *api = nccl_api();
    int rc = api->AllReduce(buf, buf, count, kNcclDouble, kNcclSum, h->comm, h->stream);
    if (rc != 0)
        return set_error(SQW_ERR_NCCL, "ncclAllReduce failed: %s",
                         api->GetErrorString ? api->GetErrorString(rc) : "?");
    return SQW_OK;
}

// One ADMM iteration's worth of launches (LOne.java:608-716), no host synchronisation.
static int enqueue_iteration(sqw_admm *h, int slot)
{
    AdmmDev &D = h->dev;
    int rc;
    SQW_CUDA_CHECK(cudaEventRecord(h->ev_x0[slot], h->stream));
    if ((rc = launch_xupdate(h)) != SQW_OK)
        return rc;
    SQW_CUDA_CHECK(cudaEventRecord(h->ev_x1[slot], h->stream));
    if (h->world == 1 && !h->unfused_tail) {
        // single GPU: pack + consensus + decide in one launch (admm_consensus_fused_kernel)
        admm_consensus_fused_kernel<1><<<D.E, kThreads, 0, h->stream>>>(D);
        SQW_CUDA_CHECK(cudaMemcpyAsync(&h->h_ring[slot], D.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost,
                                       h->stream));
        SQW_CUDA_CHECK(cudaEventRecord(h->ev_snap[slot], h->stream));
        SQW_CUDA_CHECK(cudaGetLastError());
        h->stats.launches += 2;
        return SQW_OK;
    }
    if (h->world > 1 && !h->unfused_tail) {
        // multi-GPU: ONE collective per iteration. pack writes [sum x_t | sum u_t | sum x_t^2], the
        // all-reduce makes them the sums over all T blocks, and one launch does consensus + residuals
        // (from the sums alone) + decision
        const int total3 = 2 * D.npad + D.upad;
        admm_pack_kernel<true><<<(total3 + 255) / 256, 256, 0, h->stream>>>(D);
        if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][0], h->stream);
        if ((rc = nccl_allreduce(h, D.sum1, (size_t)total3)) != SQW_OK)
            return rc;
        if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][1], h->stream);
        admm_consensus_fused_kernel<2><<<D.E, kThreads, 0, h->stream>>>(D);
        if (h->tail_prof)
            for (int j = 2; j < 5; ++j)
                cudaEventRecord(h->ev_tail[slot][j], h->stream);
        SQW_CUDA_CHECK(cudaMemcpyAsync(&h->h_ring[slot], D.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost,
                                       h->stream));
        SQW_CUDA_CHECK(cudaEventRecord(h->ev_snap[slot], h->stream));
        SQW_CUDA_CHECK(cudaGetLastError());
        h->stats.launches += 3;
        return SQW_OK;
    }
    const int total = D.npad + D.upad;
    admm_pack_kernel<false><<<(total + 255) / 256, 256, 0, h->stream>>>(D);
    if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][0], h->stream);
    if ((rc = nccl_allreduce(h, D.sum1, (size_t)total)) != SQW_OK)
        return rc;
    if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][1], h->stream);
    admm_consensus_kernel<<<D.E, kThreads, 0, h->stream>>>(D);  // block_sum() needs kThreads
    if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][2], h->stream);
    if ((rc = nccl_allreduce(h, D.sum2, (size_t)D.E)) != SQW_OK)
        return rc;
    if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][3], h->stream);
    admm_decide_kernel<<<1, kDecideThreads, 0, h->stream>>>(D);
    if (h->tail_prof) cudaEventRecord(h->ev_tail[slot][4], h->stream);
    SQW_CUDA_CHECK(cudaMemcpyAsync(&h->h_ring[slot], D.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost,
                                   h->stream));
    SQW_CUDA_CHECK(cudaEventRecord(h->ev_snap[slot], h->stream));
    SQW_CUDA_CHECK(cudaGetLastError());
    h->stats.launches += 4;
    return SQW_OK;
}

}  // namespace sqw

extern "C" {

int sqw_admm_create(sqw_admm **out)
{
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (!out)
        return set_error(SQW_ERR_INVALID_ARG, "null handle pointer");
    sqw_admm *h = new sqw_admm();
    SQW_CUDA_CHECK(cudaGetDevice(&h->device));
    SQW_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // keep what destroyed handles free in the device's default pool (dalloc): by default a pool
        // hands everything back to the driver at the next synchronisation
        cudaMemPool_t pool = nullptr;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, h->device) != cudaSuccess ||
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess)
            cudaGetLastError();   // older drivers: the allocations still work, only slower
    }
    if (const char *env = getenv("SQW_ADMM_CAPS")) {  // tuning hook (header): comma separated caps
        int i = 0;
        for (const char *q = env; *q && i < kNumPhases - 1; ++i) {
            h->caps[i] = atoi(q);
            while (*q && *q != ',')
                ++q;
            if (*q == ',')
                ++q;
        }
    }
    *out = h;
    return SQW_OK;
}

void sqw_admm_destroy(sqw_admm *h)
{
    if (!h)
        return;
    cudaSetDevice(h->device);
    if (h->stream)
        cudaStreamSynchronize(h->stream);
    if (h->l2_persist) {  // give the device-wide L2 carve-out back
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    // the NCCL communicator is shared between handles and outlives them (see sqw_admm_set_comm)
    for (void *p : h->allocs)
        cudaFreeAsync(p, h->stream);
    if (h->stream)
        cudaStreamSynchronize(h->stream);
    if (h->h_ring) {
        cudaFreeHost(h->h_ring);
        for (int i = 0; i < kRing; ++i) {
            cudaEventDestroy(h->ev_x0[i]);
            cudaEventDestroy(h->ev_x1[i]);
            cudaEventDestroy(h->ev_snap[i]);
        }
        cudaEventDestroy(h->ev_a);
        cudaEventDestroy(h->ev_b);
        cudaEventDestroy(h->ev_t0);
        cudaEventDestroy(h->ev_t1);
        if (h->tail_prof)
            for (int i = 0; i < kRing; ++i)
                for (int j = 0; j < 5; ++j)
                    cudaEventDestroy(h->ev_tail[i][j]);
    }
    if (h->stream)
        cudaStreamDestroy(h->stream);
    delete h;
}

int sqw_admm_set_params(sqw_admm *h, double ridge, double pho, int32_t num_threads,
                        int32_t admm_max_itr, int32_t nodes_max_itr, int32_t debug)
{
    if (!h)
        return set_error(SQW_ERR_INVALID_ARG, "null handle");
    if (h->have_problem)
        return set_error(SQW_ERR_STATE, "set_params must precede set_problem (the row blocking depends on T)");
    if (num_threads < 1 || admm_max_itr < 1 || nodes_max_itr < 1 || !(pho > 0.0))
        return set_error(SQW_ERR_INVALID_ARG, "invalid parameters (threads=%d A=%d S=%d pho=%g)",
                         num_threads, admm_max_itr, nodes_max_itr, pho);
    h->ridge = ridge;
    h->pho = pho;
    h->T = num_threads;
    h->admm_max_itr = admm_max_itr;
    h->nodes_max_itr = nodes_max_itr;
    h->debug = debug;
    h->have_params = true;
    return SQW_OK;
}

int sqw_admm_set_cta_budget(sqw_admm *h, int32_t max_ctas)
{
    if (!h)
        return set_error(SQW_ERR_INVALID_ARG, "null handle");
    if (max_ctas < 0)
        return set_error(SQW_ERR_INVALID_ARG, "negative CTA budget");
    if (h->have_problem)
        return set_error(SQW_ERR_INVALID_ARG, "set the CTA budget before sqw_admm_set_problem");
    h->cta_budget = max_ctas;
    return SQW_OK;
}

int sqw_admm_set_ridge(sqw_admm *h, double ridge)
{
    if (!h)
        return set_error(SQW_ERR_INVALID_ARG, "null handle");
    if (!h->have_params)
        return set_error(SQW_ERR_STATE, "set_params first");
    if (!(ridge >= 0.0))
        return set_error(SQW_ERR_INVALID_ARG, "ridge must be >= 0 (got %g)", ridge);
    h->ridge = ridge;
    h->dev.ridge = ridge;  // the training matrix and the workspace stay resident
    return SQW_OK;
}

int sqw_admm_set_structure(sqw_admm *h, int32_t num_nodes, int32_t num_layers,
                           const int32_t *node_index, const int32_t *is_leaf, const int32_t *layer,
                           const int32_t *parent_ptr, const int32_t *parent_idx,
                           const int32_t *child_ptr, const int32_t *child_idx)
{
    if (!h || num_nodes < 1 || num_layers < 1 || !node_index || !is_leaf || !layer || !parent_ptr ||
        !child_ptr)
        return set_error(SQW_ERR_INVALID_ARG, "invalid structure arguments");
    if (h->prepared)
        return set_error(SQW_ERR_STATE, "structure already in use");
    const int NN = num_nodes;
    // leaves must be 0..C-1 (they index x, Classifier.java:340,389)
    int C = 0;
    for (int e = 0; e < NN; ++e) {
        if (node_index[e] < 0 || node_index[e] >= NN || layer[e] < 0 || layer[e] >= num_layers)
            return set_error(SQW_ERR_INVALID_ARG, "design entry %d out of range", e);
        if (is_leaf[e])
            ++C;
    }
    std::vector<int> entry_of(NN, -1);
    for (int e = 0; e < NN; ++e)
        entry_of[node_index[e]] = e;
    for (int c = 0; c < C; ++c)
        if (entry_of[c] < 0 || !is_leaf[entry_of[c]])
            return set_error(SQW_ERR_INVALID_ARG, "leaf indices must be 0..%d", C - 1);
    h->NN = NN;
    h->num_layers = num_layers;
    h->C_struct = C;
    h->edge_leaf.clear();
    h->edge_par.clear();
    h->leaf_edge_ptr.assign(C + 1, 0);
    for (int c = 0; c < C; ++c) {
        const int e = entry_of[c];
        const int np = parent_ptr[e + 1] - parent_ptr[e];
        if (np > kMaxEdgesPerLeaf)
            return set_error(SQW_ERR_UNSUPPORTED, "leaf %d has %d parents (max %d)", c, np,
                             kMaxEdgesPerLeaf);
        if (np > 0) {
            for (int k = parent_ptr[e]; k < parent_ptr[e + 1]; ++k) {
                if (parent_idx[k] < 0 || parent_idx[k] >= NN)
                    return set_error(SQW_ERR_INVALID_ARG, "parent index out of range");
                h->edge_leaf.push_back(c);
                h->edge_par.push_back(parent_idx[k]);
            }
        } else {  // self edge of a parentless leaf (LOne.java:440,697,846)
            h->edge_leaf.push_back(c);
            h->edge_par.push_back(-1);
        }
        h->leaf_edge_ptr[c + 1] = (int)h->edge_leaf.size();
    }
    const int E = (int)h->edge_leaf.size();
    h->edge_sum_order.resize(E);
    for (int i = 0; i < E; ++i)
        h->edge_sum_order[i] = i;
    std::stable_sort(h->edge_sum_order.begin(), h->edge_sum_order.end(), [&](int a, int b) {
        const long long ka = (long long)h->edge_leaf[a] * NN + (h->edge_par[a] < 0 ? h->edge_leaf[a] : h->edge_par[a]);
        const long long kb = (long long)h->edge_leaf[b] * NN + (h->edge_par[b] < 0 ? h->edge_leaf[b] : h->edge_par[b]);
        return ka < kb;
    });
    // label layers in file order; neighbours = parents then children (LOne.java:229-234)
    h->layers.assign(num_layers, sqw_admm::Layer());
    for (int e = 0; e < NN; ++e) {
        const int l = layer[e];
        if (l == 0)
            continue;
        auto &L = h->layers[l];
        if (L.nbr_ptr.empty())
            L.nbr_ptr.push_back(0);
        L.nodes.push_back(node_index[e]);
        if (is_leaf[e])
            for (int k = parent_ptr[e]; k < parent_ptr[e + 1]; ++k)
                L.nbr_idx.push_back(parent_idx[k]);
        else
            for (int k = child_ptr[e]; k < child_ptr[e + 1]; ++k) {
                if (child_idx[k] < 0 || child_idx[k] >= NN)
                    return set_error(SQW_ERR_INVALID_ARG, "child index out of range");
                L.nbr_idx.push_back(child_idx[k]);
            }
        if ((int)L.nbr_idx.size() - L.nbr_ptr.back() > kMaxFanIn)
            return set_error(SQW_ERR_UNSUPPORTED, "node %d has more than %d neighbours",
                             node_index[e], kMaxFanIn);
        L.nbr_ptr.push_back((int)L.nbr_idx.size());
    }
    h->have_structure = true;
    return SQW_OK;
}

int sqw_nccl_unique_id(uint8_t id_out[128])
{
    NcclApi *api = nccl_api();
    if (!api)
        return set_error(SQW_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    NcclUid id;
    int rc = api->GetUniqueId(&id);
    if (rc != 0)
        return set_error(SQW_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
    memcpy(id_out, id.internal, 128);
    return SQW_OK;
}

int sqw_admm_set_comm(sqw_admm *h, int32_t rank, int32_t world, const uint8_t nccl_id[128])
{
    if (!h || world < 1 || rank < 0 || rank >= world)
        return set_error(SQW_ERR_INVALID_ARG, "invalid rank/world");
    if (h->have_problem)
        return set_error(SQW_ERR_STATE, "set_comm must precede set_problem");
    h->rank = rank;
    h->world = world;
    if (world == 1)
        return SQW_OK;
    NcclApi *api = nccl_api();
    if (!api)
        return set_error(SQW_ERR_NCCL, "libnccl.so.2 could not be loaded");
    // One communicator per ncclUniqueId per process: an id can initialise only one communicator,
    // and every LOne instance of a run (1 + x CV folds) shares it. It lives until process exit.
    struct Cached {
        std::string id;
        int rank, world;
        void *comm;
    };
    static std::vector<Cached> cache;
    static std::mutex cache_mu;
    std::lock_guard<std::mutex> lock(cache_mu);
    const std::string key(reinterpret_cast<const char *>(nccl_id), 128);
    for (const Cached &c : cache)
        if (c.id == key) {
            if (c.rank != rank || c.world != world)
                return set_error(SQW_ERR_INVALID_ARG, "ncclUniqueId reused with a different rank/world");
            h->comm = c.comm;
            return SQW_OK;
        }
    NcclUid id;
    memcpy(id.internal, nccl_id, 128);
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    int rc = api->CommInitRank(&h->comm, world, id, rank);
    if (rc != 0)
        return set_error(SQW_ERR_NCCL, "ncclCommInitRank failed: %s",
                         api->GetErrorString ? api->GetErrorString(rc) : "?");
    cache.push_back(Cached{key, rank, world, h->comm});
    return SQW_OK;
}

// Row blocking shared by both set_problem variants (LOne.java:337-357): block t holds rows
// i = t + T*s, s < floor(N/T); the local blocks are those with t % world == rank, stored in the
// order the reference sums the thread buffers.
static int plan_blocks(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                       bool counts_form = false)
{
    if (!h || n_rows < 1 || dim < 2 || num_classes < 1)
        return set_error(SQW_ERR_INVALID_ARG, "invalid problem sizes");
    if (!h->have_params)
        return set_error(SQW_ERR_STATE, "set_params must precede set_problem");
    if (h->have_problem)
        return set_error(SQW_ERR_STATE, "problem already set");
    if (n_rows / h->T < 1)
        return set_error(SQW_ERR_INVALID_ARG, "fewer rows (%lld) than ADMM blocks (%d)",
                         (long long)n_rows, h->T);
    if (n_rows / h->T > (1 << 30))
        return set_error(SQW_ERR_UNSUPPORTED, "block too large");
    h->n_rows = n_rows;
    h->dim = dim;
    h->C = num_classes;
    h->block_ids.clear();
    for (int t : java_thread_order(h->T))
        if (t % h->world == h->rank)
            h->block_ids.push_back(t);
    if (h->block_ids.empty())
        return set_error(SQW_ERR_INVALID_ARG, "rank %d owns no ADMM block (T=%d, world=%d)",
                         h->rank, h->T, h->world);
    AdmmDev &D = h->dev;
    D.T = h->T;
    D.T_local = (int)h->block_ids.size();
    D.nb = (int)(n_rows / h->T);
    D.dim = dim;
    // Evaluation mode and padded row length (admm_kernels.cuh, admm_eval_tma.cuh):
    //   mode 4 (TMA tiles): dimp = 4 (mod 8) doubles; modes 1/2 (streaming): dimp = 8 (mod 16).
    cudaDeviceProp prop;
    SQW_CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
    const int Cp = (num_classes + 7) / 8 * 8, nct = Cp / 8;
    if (nct > 3)
        return set_error(SQW_ERR_UNSUPPORTED, "more than 24 leaf classes (%d) not supported yet",
                         num_classes);
    int dimp4 = (dim + 3) / 4 * 4;
    if ((dimp4 & 7) == 0)
        dimp4 += 4;
    int dimp8 = (dim + 7) / 8 * 8;
    if (((dimp8 / 8) & 1) == 0)
        dimp8 += 8;
    // static shared memory of the x-update kernel (SolveSmem, tables, profile counters): 5.5 KB + 1 KB
    // reserved by the driver
    const size_t budget = (size_t)prop.sharedMemPerBlockOptin - (SQW_AUG_CACHE ? 7424 : 8192);
    // (one class tile: W lives in registers and the head of the layout is the solver's slice cache)
    TileSmemLayout L = tile_smem_layout(Cp, dimp4, budget, nct == 1 ? kSolveSliceBytes : 0);
    // single-pass tiles: one class tile only (with more the register-resident gradient
    // accumulators of a whole row spill; the column-chunked mode 6 keeps them per chunk)
    const bool ws_ok = nct == 1 && L.NS >= 3 && L.NS <= 4 && (dimp4 + 15) / 16 <= kDmmaWarps * 6 &&
                       dimp4 / 4 <= kDmmaWarps * 24;
    int mode = ws_ok ? 4 : ((size_t)Cp * dimp8 * 8 <= 160 * 1024 ? 1 : 2);
    if (mode == 4 && nct == 1 && (dimp4 / 4 + kDmmaWarps - 1) / kDmmaWarps <= 21)
        mode = 5;  // same kernel with 21 instead of 24 unrolled k-steps per DMMA warp
    TileSmemLayout Lc = chunk_smem_layout(Cp, budget, 0);
    if (mode != 4 && mode != 5 && Lc.NS >= 2)
        mode = 6;  // long rows: column-chunked TMA tiles instead of per-thread global loads
    if (const char *env = getenv("SQW_ADMM_EVAL_MODE")) {  // test hook: force a more general path
        const int m = atoi(env);
        if (m == 6 && Lc.NS >= 2)
            mode = 6;
        if (m == 1 && (size_t)Cp * dimp8 * 8 <= 160 * 1024)
            mode = 1;
        else if (m == 2)
            mode = 2;
    }
    h->eval_mode = mode;
    h->layout = (mode == 6) ? Lc : L;
    D.dimp = (mode == 4 || mode == 5 || mode == 6) ? dimp4 : dimp8;
    // Packed-count form (admm_eval_packed.cuh): one class tile, rows up to 768 columns, and the
    // rows of the busiest CTA of the smallest group a launch can form must fit shared memory next
    // to the partial logits. Otherwise the counts are expanded to the fp64 matrix on the device.
    const bool packed_ok = counts_form && !getenv("SQW_ADMM_NO_PACKED") && !getenv("SQW_ADMM_EVAL_MODE");
    if (packed_ok) {
        // any shape: the streamed, column-chunked packed evaluation (admm_eval_packed_chunk.cuh);
        // replaced below by the resident one where the rows fit shared memory
        const int dimp16 = (dim + 15) / 16 * 16;
        TileSmemLayout Lq = packed_chunk_smem_layout(nct, false, nct == 1 ? kSolveSliceBytes : 0);
        if (Lq.total <= budget) {
            h->eval_mode = 10;
            h->layout = Lq;
            D.dimp = dimp16;
            D.rsq = dimp16;
        }
        // C = 9 or 17 (2^k labels + one outgroup class): one class tile less, the last class as
        // DFMAs next to the DMMAs (MODE 11). SQW_ADMM_PACKED_MODE=10 keeps the padded form.
        const char *force = getenv("SQW_ADMM_PACKED_MODE");
        if ((num_classes == 9 || num_classes == 17) && !(force && atoi(force) == 10)) {
            const int nx = (num_classes - 1) / 8;
            TileSmemLayout Lx = packed_chunk_smem_layout(nx, true, nx == 1 ? kSolveSliceBytes : 0);
            if (Lx.total <= budget) {
                h->eval_mode = 11;
                h->layout = Lx;
                D.dimp = dimp16;
                D.rsq = dimp16;
            }
        }
    }
    if (packed_ok && nct == 1 && dim <= kPackedMaxDim && !(getenv("SQW_ADMM_PACKED_MODE") &&
                                                          atoi(getenv("SQW_ADMM_PACKED_MODE")) == 10)) {
        const int dimp16 = (dim + 15) / 16 * 16;
        const long long tiles = (D.nb + 7) / 8;
        const char *force = getenv("SQW_ADMM_PACKED_MODE");   // test hook: 9
        // MODE 9 (two CTAs of 192 threads per SM, so that one CTA evaluates while the other sits in
        // its solver's exchange chain) is built and tested but OFF: measured at cfg-2 70.6 us per
        // evaluation round against MODE 8's 52.3, with and without a start skew between the two
        // CTAs of an SM - one DMMA warp per SM sub-partition reaches only half the DMMA issue rate
        // (its byte conversions and DMMAs serialise), so a CTA alone on an SM evaluates no faster
        // than two side by side, and the doubled groups lengthen the exchange chain.
        if (D.T_local >= 2 && force && atoi(force) == 9) {
            int G = 1, gcap = 1;
            group_sizes(h, prop.multiProcessorCount * ModeCfg<9>::MINB, D.T_local, D.nb, G, gcap);
            const int gmin = std::min(G, gcap);
            const int rows_cap8 = (int)((tiles + gmin - 1) / gmin) * 8;
            TileSmemLayout Lp = packed_smem_layout(dimp16, rows_cap8,
                                                   (size_t)kSliceVectors * kPackedSliceCap2 * sizeof(double),
                                                   ModeCfg<9>::DW);
            int per_sm = 0;
            cudaFuncAttributes fa;
            const void *fn = (const void *)admm_xupdate_kernel<1, 9>;
            if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess &&
                Lp.total + fa.sharedSizeBytes <= (size_t)prop.sharedMemPerBlockOptin &&
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, ModeCfg<9>::NT, Lp.total) == cudaSuccess &&
                per_sm >= ModeCfg<9>::MINB && (D.Cp * dimp16 + gmin - 1) / gmin <= kPackedSliceCap2 - 15) {
                h->eval_mode = 9;
                h->layout = Lp;
                D.dimp = dimp16;
                D.rsq = dimp16;
            }
            cudaGetLastError();
        }
        if (h->eval_mode != 9) {
            int G = 1, gcap = 1;
            group_sizes(h, prop.multiProcessorCount, D.T_local, D.nb, G, gcap);
            const int gmin = std::min(G, gcap);
            const int rows_cap8 = (int)((tiles + gmin - 1) / gmin) * 8;
            TileSmemLayout Lp = packed_smem_layout(dimp16, rows_cap8, kSolveSliceBytes);
            const bool force12 = force && atoi(force) == 12;   // test hook: panels even where the rows fit
            if (Lp.total <= budget && !force12) {
                h->eval_mode = 8;
                h->layout = Lp;
                D.dimp = dimp16;
                D.rsq = dimp16;
            } else {
                // MODE 12: the CTA's rows pass through the resident form's buffer panel by panel; the
                // panel is as large as shared memory allows (SQW_ADMM_PANEL_ROWS: test hook, smaller)
                // (a row costs its bytes + 8 K-slices of partial logits + weight and class: start the
                // search just above what the budget can hold)
                int cap = (int)std::min<long long>(rows_cap8, (long long)(budget / (size_t)(dimp16 + 524)) / 8 * 8 + 8);
                if (const char *pr = getenv("SQW_ADMM_PANEL_ROWS"))
                    cap = std::min(cap, std::max(8, atoi(pr) / 8 * 8));
                while (cap > 8 && packed_smem_layout(dimp16, cap, kSolveSliceBytes).total > budget)
                    cap -= 8;
                Lp = packed_smem_layout(dimp16, cap, kSolveSliceBytes);
                if (Lp.total <= budget) {
                    h->eval_mode = 12;
                    h->layout = Lp;
                    D.dimp = dimp16;
                    D.rsq = dimp16;
                }
            }
        }
    }
    return SQW_OK;
}

int sqw_admm_set_problem(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                         const double *data, const int32_t *y, const double *w)
{
    int rc = plan_blocks(h, n_rows, dim, num_classes);
    if (rc != SQW_OK)
        return rc;
    if (!data || !y || !w)
        return set_error(SQW_ERR_INVALID_ARG, "null buffer");
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    AdmmDev &D = h->dev;
    const size_t rows = (size_t)D.T_local * D.nb;
    double *dX = nullptr, *dw = nullptr;
    int *dy = nullptr;
    if ((rc = h->dalloc(&dX, rows * D.dimp, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dy, rows, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dw, rows, false)) != SQW_OK) return rc;
    // stage one block at a time in pinned memory
    double *stage = nullptr;
    const size_t blk_elems = (size_t)D.nb * D.dimp;
    SQW_CUDA_CHECK(cudaMallocHost((void **)&stage, blk_elems * sizeof(double)));
    std::vector<int> ys(D.nb);
    std::vector<double> ws(D.nb);
    for (int b = 0; b < D.T_local; ++b) {
        const int t = h->block_ids[b];
        for (int s = 0; s < D.nb; ++s) {
            const int64_t i = (int64_t)t + (int64_t)h->T * s;
            if (y[i] < 0 || y[i] >= num_classes) {
                cudaFreeHost(stage);
                return set_error(SQW_ERR_INVALID_ARG, "class %d of row %lld out of range", y[i],
                                 (long long)i);
            }
            double *dst = stage + (size_t)s * D.dimp;
            memcpy(dst, data + i * dim, sizeof(double) * dim);
            for (int j = dim; j < D.dimp; ++j)
                dst[j] = 0.0;
            ys[s] = y[i];
            ws[s] = w[i];
        }
        cudaError_t e = cudaMemcpyAsync(dX + (size_t)b * blk_elems, stage, blk_elems * sizeof(double),
                                        cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dy + (size_t)b * D.nb, ys.data(), sizeof(int) * D.nb,
                                cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dw + (size_t)b * D.nb, ws.data(), sizeof(double) * D.nb,
                                cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            cudaFreeHost(stage);
            return set_error(SQW_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
        }
    }
    cudaFreeHost(stage);
    D.X = dX;
    D.y = dy;
    D.w = dw;
    h->have_problem = true;
    return SQW_OK;
}

int sqw_admm_set_problem_dev(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                             const double *d_data, const int32_t *d_y, const double *d_w)
{
    int rc = plan_blocks(h, n_rows, dim, num_classes);
    if (rc != SQW_OK)
        return rc;
    if (!d_data || !d_y || !d_w)
        return set_error(SQW_ERR_INVALID_ARG, "null buffer");
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    AdmmDev &D = h->dev;
    const size_t rows = (size_t)D.T_local * D.nb;
    double *dX = nullptr, *dw = nullptr;
    int *dy = nullptr, *dbid = nullptr;
    if ((rc = h->dalloc(&dX, rows * D.dimp, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dy, rows, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dw, rows, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dbid, (size_t)D.T_local, false)) != SQW_OK) return rc;
    SQW_CUDA_CHECK(cudaMemcpyAsync(dbid, h->block_ids.data(), sizeof(int) * D.T_local,
                                   cudaMemcpyHostToDevice, h->stream));
    const int grid = (int)std::min<size_t>(rows, 148 * 16);
    admm_block_rows_kernel<<<grid, 256, 0, h->stream>>>(d_data, d_y, d_w, dim, h->T, dbid,
                                                        D.T_local, D.nb, D.dimp, dX, dy, dw);
    SQW_CUDA_CHECK(cudaGetLastError());
    SQW_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    D.X = dX;
    D.y = dy;
    D.w = dw;
    h->have_problem = true;
    return SQW_OK;
}

}  // extern "C"

// Shared tail of the two count-matrix entry points: d_src holds the counts on the device
// (SrcT = uint8 with the kept features as consecutive columns, or int32 with a column index list).
template <typename SrcT>
static int build_from_counts(sqw_admm *h, const SrcT *d_src, long long src_stride, const int *d_col_idx,
                             const double *d_mean, const double *d_sd, const int *d_y,
                             const double *d_w, int rows_local, cudaStream_t st)
{
    AdmmDev &D = h->dev;
    int rc;
    const size_t rows = (size_t)D.T_local * D.nb;
    double *dw = nullptr, *mean_p = nullptr, *invsd_p = nullptr, *mean_f = nullptr, *sd_f = nullptr;
    int *dy = nullptr, *dbid = nullptr, *dstatus = nullptr;
    if ((rc = h->dalloc(&dy, rows, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dw, rows, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dbid, (size_t)D.T_local, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&dstatus, 1, true)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&mean_p, (size_t)D.dimp, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&invsd_p, (size_t)D.dimp, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&mean_f, (size_t)D.dimp, false)) != SQW_OK) return rc;
    if ((rc = h->dalloc(&sd_f, (size_t)D.dimp, false)) != SQW_OK) return rc;
    // (allocation and memset ran on the handle's stream: order the caller's stream behind them)
    SQW_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    SQW_CUDA_CHECK(cudaMemcpyAsync(dbid, h->block_ids.data(), sizeof(int) * D.T_local,
                                   cudaMemcpyHostToDevice, st));
    admm_packed_stats_kernel<<<(D.dimp + 255) / 256, 256, 0, st>>>(d_mean, d_sd, d_col_idx, D.dim, D.dimp,
                                                                   mean_p, invsd_p, mean_f, sd_f);
    const int grid = (int)std::min<size_t>(rows, 148 * 16);
    if (h->eval_mode == 8 || h->eval_mode == 9 || h->eval_mode == 10 || h->eval_mode == 11 || h->eval_mode == 12) {
        unsigned char *dXq = nullptr;
        if ((rc = h->dalloc(&dXq, rows * D.rsq + 16384, false)) != SQW_OK) return rc;
        admm_gather_counts_kernel<SrcT, true><<<grid, 256, 0, st>>>(
            d_src, src_stride, d_col_idx, d_y, d_w, D.dim, h->T, dbid, D.T_local, D.nb, h->world,
            rows_local, mean_f, sd_f, D.rsq, dXq, dy, dw, dstatus);
        D.Xq = dXq;
        D.cmean = mean_p;
        D.cinvsd = invsd_p;
    } else {
        double *dX = nullptr;
        if ((rc = h->dalloc(&dX, rows * D.dimp, false)) != SQW_OK) return rc;
        admm_gather_counts_kernel<SrcT, false><<<grid, 256, 0, st>>>(
            d_src, src_stride, d_col_idx, d_y, d_w, D.dim, h->T, dbid, D.T_local, D.nb, h->world,
            rows_local, mean_f, sd_f, D.dimp, dX, dy, dw, dstatus);
        D.X = dX;
    }
    SQW_CUDA_CHECK(cudaGetLastError());
    int status = 0;
    SQW_CUDA_CHECK(cudaMemcpyAsync(&status, dstatus, sizeof(int), cudaMemcpyDeviceToHost, st));
    SQW_CUDA_CHECK(cudaStreamSynchronize(st));
    if (status != 0)
        return set_error(SQW_ERR_UNSUPPORTED, "a count exceeds 255: pass the standardised matrix to "
                                              "sqw_admm_set_problem instead");
    D.y = dy;
    D.w = dw;
    h->have_problem = true;
    return SQW_OK;
}

extern "C" {

int sqw_admm_set_problem_counts(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                                const uint8_t *counts, const double *mean, const double *sd,
                                const int32_t *y, const double *w)
{
    int rc = plan_blocks(h, n_rows, dim, num_classes, true);
    if (rc != SQW_OK)
        return rc;
    if (!counts || !mean || !sd || !y || !w)
        return set_error(SQW_ERR_INVALID_ARG, "null buffer");
    for (int64_t i = 0; i < n_rows; ++i)
        if (y[i] < 0 || y[i] >= num_classes)
            return set_error(SQW_ERR_INVALID_ARG, "class %d of row %lld out of range", y[i], (long long)i);
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    // the whole count matrix is 1 byte per element (13 MB at cfg-2): one copy, then the same device
    // gather as the _dev variant; the scratch copies are freed before returning
    const size_t F = (size_t)dim - 1;
    uint8_t *d_src = nullptr;
    double *d_ms = nullptr, *d_w = nullptr;
    int *d_y = nullptr;
    cudaStream_t st = h->stream;
    cudaError_t e = cudaMallocAsync((void **)&d_src, (size_t)n_rows * F, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_ms, sizeof(double) * 2 * dim, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_w, sizeof(double) * n_rows, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_y, sizeof(int) * n_rows, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_src, counts, (size_t)n_rows * F, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ms, mean, sizeof(double) * dim, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ms + dim, sd, sizeof(double) * dim, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, w, sizeof(double) * n_rows, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_y, y, sizeof(int) * n_rows, cudaMemcpyHostToDevice, st);
    rc = (e == cudaSuccess) ? build_from_counts<uint8_t>(h, d_src, (long long)F, nullptr, d_ms, d_ms + dim,
                                                          d_y, d_w, 0, st)
                            : set_error(SQW_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    // (the host buffers are pageable: every copy above has returned from the host's side, and the
    // frees are ordered behind the gather on the stream)
    if (d_src) cudaFreeAsync(d_src, st);
    if (d_ms) cudaFreeAsync(d_ms, st);
    if (d_w) cudaFreeAsync(d_w, st);
    if (d_y) cudaFreeAsync(d_y, st);
    return rc;
}

int sqw_admm_set_problem_counts_dev(sqw_admm *h, int64_t n_rows, int64_t numK, int32_t num_classes,
                                    const int32_t *d_counts, const int32_t *d_keep_idx,
                                    int32_t num_kept, const double *d_mean, const double *d_sd,
                                    const int32_t *d_y, const double *d_w, int32_t rows_local,
                                    void *stream)
{
    if (num_kept < 1 || numK < num_kept)
        return set_error(SQW_ERR_INVALID_ARG, "invalid column plan");
    int rc = plan_blocks(h, n_rows, num_kept + 1, num_classes, true);
    if (rc != SQW_OK)
        return rc;
    if (!d_counts || !d_keep_idx || !d_mean || !d_sd || !d_y || !d_w)
        return set_error(SQW_ERR_INVALID_ARG, "null buffer");
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    return build_from_counts<int32_t>(h, d_counts, (long long)numK, d_keep_idx, d_mean, d_sd, d_y, d_w,
                                      rows_local ? 1 : 0, (cudaStream_t)stream);
}

// host <-> padded device layout helpers
static int upload_padded(sqw_admm *h, const double *src, int rows, int dim, int dimp, double *dst)
{
    std::vector<double> tmp((size_t)rows * dimp, 0.0);
    for (int r = 0; r < rows; ++r)
        memcpy(&tmp[(size_t)r * dimp], src + (size_t)r * dim, sizeof(double) * dim);
    SQW_CUDA_CHECK(cudaMemcpyAsync(dst, tmp.data(), tmp.size() * sizeof(double),
                                   cudaMemcpyHostToDevice, h->stream));
    SQW_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return SQW_OK;
}

static int download_padded(sqw_admm *h, const double *src, int rows, int dim, int dimp, double *dst)
{
    std::vector<double> tmp((size_t)rows * dimp);
    SQW_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), src, tmp.size() * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->stream));
    SQW_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < rows; ++r)
        memcpy(dst + (size_t)r * dim, &tmp[(size_t)r * dimp], sizeof(double) * dim);
    return SQW_OK;
}

static int execute_impl(sqw_admm *h, double *x_inout, double *smx_inout);

int sqw_admm_execute(sqw_admm *h, double *x_inout, double *smx_inout)
{
    if (!h || !x_inout || !smx_inout)
        return set_error(SQW_ERR_INVALID_ARG, "null argument");
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    const int rc = execute_impl(h, x_inout, smx_inout);
    if (rc != SQW_OK && h->stream) {
        // an error path may leave speculatively enqueued iterations behind: let them finish (their
        // kernels return at once when Ctrl.done / fatal is set) before the caller touches the handle
        cudaStreamSynchronize(h->stream);
        cudaGetLastError();
    }
    return rc;
}

static int execute_impl(sqw_admm *h, double *x_inout, double *smx_inout)
{
    int rc = prepare(h);
    if (rc != SQW_OK)
        return rc;
    AdmmDev &D = h->dev;
    cudaStream_t st = h->stream;

    // LOne(xinit, sm_xinit, ...) + initUandZ (LOne.java:80-85,110-114)
    if ((rc = upload_padded(h, x_inout, D.C, D.dim, D.dimp, D.xbar)) != SQW_OK) return rc;
    if ((rc = upload_padded(h, smx_inout, D.NN, D.dim, D.dimp, D.smx)) != SQW_OK) return rc;
    SQW_CUDA_CHECK(cudaMemsetAsync(D.z, 0, sizeof(double) * D.upad, st));
    SQW_CUDA_CHECK(cudaMemsetAsync(D.zold, 0, sizeof(double) * D.upad, st));
    SQW_CUDA_CHECK(cudaMemsetAsync(D.ubar, 0, sizeof(double) * D.upad, st));
    SQW_CUDA_CHECK(cudaMemsetAsync(D.xt, 0, sizeof(double) * (size_t)D.T_local * D.Cp * D.dimp, st));
    Ctrl c0;
    memset(&c0, 0, sizeof c0);
    c0.pho = h->pho;
    c0.fold = 1.0;
    c0.max_itr = h->admm_max_itr;
    c0.log_cap = std::max(1, h->admm_max_itr) * std::max(1, h->nodes_max_itr);
    SQW_CUDA_CHECK(cudaMemcpyAsync(D.ctrl, &c0, sizeof c0, cudaMemcpyHostToDevice, st));
    SQW_CUDA_CHECK(cudaStreamSynchronize(st));

    h->stats = sqw_admm_stats{};
    h->admm_iters.clear();
    double ms_admm = 0.0, ms_eval = 0.0;
    cudaEvent_t ev_t0 = h->ev_t0, ev_t1 = h->ev_t1;
    SQW_CUDA_CHECK(cudaEventRecord(ev_t0, st));
    Ctrl last = c0;
    bool converged = false;
    int outer_done = 0;

    for (int it = 0; it < h->nodes_max_itr; ++it) {
        admm_begin_kernel<<<148, 256, 0, st>>>(D, it, 1);
        h->stats.launches += 1;
        SQW_CUDA_CHECK(cudaEventRecord(h->ev_a, st));
        // ADMMrunner.execute (LOne.java:595-760)
        int enq = 0, seen = 0;
        bool done = false;
        while (!done) {
            if (enq - seen < kLookahead && enq < h->admm_max_itr + kLookahead) {
                if ((rc = enqueue_iteration(h, enq % kRing)) != SQW_OK)
                    return rc;
                ++enq;
                continue;
            }
            const int slot = seen % kRing;
            SQW_CUDA_CHECK(cudaEventSynchronize(h->ev_snap[slot]));
            float ms = 0.f;
            SQW_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_x0[slot], h->ev_x1[slot]));
            ms_eval += ms;
            if (h->tail_prof && !h->h_ring[slot].done) {
                h->tail_ms[0] += ms;
                cudaEvent_t prev = h->ev_x1[slot];
                for (int j = 0; j < 5; ++j) {
                    float t = 0.f;
                    if (cudaEventElapsedTime(&t, prev, h->ev_tail[slot][j]) == cudaSuccess)
                        h->tail_ms[1 + j] += t;
                    prev = h->ev_tail[slot][j];
                }
            }
            last = h->h_ring[slot];
            ++seen;
            if (last.done || last.fatal)
                done = true;
        }
        // drain the speculative iterations (their kernels return at once)
        while (seen < enq) {
            SQW_CUDA_CHECK(cudaEventSynchronize(h->ev_snap[seen % kRing]));
            ++seen;
        }
        SQW_CUDA_CHECK(cudaEventRecord(h->ev_b, st));
        SQW_CUDA_CHECK(cudaEventSynchronize(h->ev_b));
        float ms = 0.f;
        SQW_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_a, h->ev_b));
        ms_admm += ms;
        h->admm_iters.push_back(last.itr);
        h->stats.total_admm_iters += last.itr;
        outer_done = it + 1;
        if (last.fatal)
            break;

        // sm_x[leaf] <- x-bar; updateInternalNodes: odd layers first, then even (LOne.java:208-221)
        admm_copy_leaves_kernel<<<(D.npad + 255) / 256, 256, 0, st>>>(D);
        h->stats.launches += 1;
        for (int pass = 0; pass < 2; ++pass)
            for (int l = (pass == 0 ? 1 : 2); l < h->num_layers; l += 2) {
                auto &L = h->layers[l];
                if (L.nodes.empty())
                    continue;
                const long long work = (long long)L.nodes.size() * (D.dim - 1);
                admm_update_layer_kernel<<<(int)std::min<long long>((work + 255) / 256, 148 * 8), 256,
                                           0, st>>>(D, L.d_nodes, (int)L.nodes.size(), L.d_ptr,
                                                    L.d_idx);
                h->stats.launches += 1;
            }
        admm_outer_check_kernel<<<D.NN, kThreads, 0, st>>>(D);
        h->stats.launches += 1;
        Ctrl cc;
        SQW_CUDA_CHECK(cudaMemcpyAsync(&h->h_ring[0], D.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        SQW_CUDA_CHECK(cudaStreamSynchronize(st));
        cc = h->h_ring[0];
        last = cc;
        // LOne.java:176-180
        if (cc.num_pass == D.NN)
            converged = true;
        else if (cc.num_pass_relaxed == D.NN && it > h->nodes_max_itr / 2)
            converged = true;
        if (converged)
            break;
    }
    SQW_CUDA_CHECK(cudaEventRecord(ev_t1, st));
    SQW_CUDA_CHECK(cudaEventSynchronize(ev_t1));
    float ms_total = 0.f;
    SQW_CUDA_CHECK(cudaEventElapsedTime(&ms_total, ev_t0, ev_t1));

    if ((rc = download_padded(h, D.xbar, D.C, D.dim, D.dimp, x_inout)) != SQW_OK) return rc;
    if ((rc = download_padded(h, D.smx, D.NN, D.dim, D.dimp, smx_inout)) != SQW_OK) return rc;

    if (h->tail_prof && h->stats.total_admm_iters > 0) {
        const double it = (double)h->stats.total_admm_iters;
        fprintf(stderr, "[sqw tail profile] rank %d, %d ADMM iterations, us per iteration: x-update (own blocks) "
                        "%.1f | pack %.1f | all-reduce 1 incl. wait for the slowest rank %.1f | consensus %.1f | "
                        "all-reduce 2 %.1f | decide %.1f\n",
                h->rank, h->stats.total_admm_iters, 1e3 * h->tail_ms[0] / it, 1e3 * h->tail_ms[1] / it,
                1e3 * h->tail_ms[2] / it, 1e3 * h->tail_ms[3] / it, 1e3 * h->tail_ms[4] / it,
                1e3 * h->tail_ms[5] / it);
        for (double &v : h->tail_ms)
            v = 0.0;
    }
    if (D.prof) {
        // thread 0 of CTA 0 (first CTA of the first unfinished block), per phase of the x-update
        std::vector<long long> pr((size_t)D.Gmax * kProfSlots * kNumPhases);
        SQW_CUDA_CHECK(cudaMemcpy(pr.data(), D.prof, pr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const char *names[7] = {"eval", "barrier1", "reduce", "barrier2", "decide", "update", "barrier3"};
        const char *en[8] = {"kernel_epilogue", "tilewait", "pass1", "rr_wait", "kernel_prologue", "pass2",
                             "free_arrive", "combine"};
        const char *sn[8] = {"R:loads+gradient", "R:products+shuffles", "R:sync", "R:warp sums+store",
                             "S:partials", "S:scalar step", "R:partials arrived (probe)",
                             "R:own operands + augmented term (probe)"};
        for (int ph = 0; ph < kNumPhases; ++ph) {
            const long long *q = &pr[(size_t)ph * D.Gmax * kProfSlots];
            if (!q[7])
                continue;
            fprintf(stderr, "[sqw profile] phase %d cta 0: evals %lld;", ph, q[7]);
            for (int i = 0; i < 7; ++i)
                fprintf(stderr, " %s %.0f ns", names[i], (double)q[i] / q[7]);
            fprintf(stderr, "\n[sqw profile]    eval breakdown per eval (ns):");
            for (int i = 0; i < 8; ++i)
                fprintf(stderr, " %s %.0f", en[i], (double)q[8 + i] / q[7]);
            fprintf(stderr, "\n[sqw profile]    solver breakdown per eval (ns):");
            for (int i = 0; i < 8; ++i)
                if (i < 6 || q[16 + i])
                    fprintf(stderr, " %s %.0f", sn[i], (double)q[16 + i] / q[7]);
            fprintf(stderr, "\n");
            if (q[31]) {
                fprintf(stderr, "[sqw profile]    first trip of a launch (%lld launches, ns):", q[31]);
                for (int i = 0; i < 7; ++i)
                    fprintf(stderr, " %s %.0f", names[i], (double)q[24 + i] / q[31]);
                fprintf(stderr, "\n");
            }
        }
        SQW_CUDA_CHECK(cudaMemset(D.prof, 0, pr.size() * sizeof(long long)));
    }
    // logs
    const int rows = last.log_rows;
    h->log_outer.resize(rows); h->log_itr.resize(rows);
    h->log_pho.resize(rows); h->log_fold.resize(rows);
    h->log_primal.resize(rows); h->log_dual.resize(rows);
    h->log_nfev.resize((size_t)rows * D.T_local);
    if (rows > 0) {
        SQW_CUDA_CHECK(cudaMemcpy(h->log_outer.data(), D.log_outer, sizeof(int) * rows, cudaMemcpyDeviceToHost));
        SQW_CUDA_CHECK(cudaMemcpy(h->log_itr.data(), D.log_itr, sizeof(int) * rows, cudaMemcpyDeviceToHost));
        SQW_CUDA_CHECK(cudaMemcpy(h->log_pho.data(), D.log_pho, sizeof(double) * rows, cudaMemcpyDeviceToHost));
        SQW_CUDA_CHECK(cudaMemcpy(h->log_fold.data(), D.log_fold, sizeof(double) * rows, cudaMemcpyDeviceToHost));
        SQW_CUDA_CHECK(cudaMemcpy(h->log_primal.data(), D.log_primal, sizeof(double) * rows, cudaMemcpyDeviceToHost));
        SQW_CUDA_CHECK(cudaMemcpy(h->log_dual.data(), D.log_dual, sizeof(double) * rows, cudaMemcpyDeviceToHost));
        SQW_CUDA_CHECK(cudaMemcpy(h->log_nfev.data(), D.log_nfev, sizeof(int) * (size_t)rows * D.T_local, cudaMemcpyDeviceToHost));
    }
    h->stats.outer_iters = outer_done;
    h->stats.converged = converged ? 1 : 0;
    h->stats.ran_admm = last.ran_admm;
    h->stats.total_evals = last.total_evals;
    h->stats.final_pho = last.pho;
    h->stats.final_fold = last.fold;
    h->stats.seconds_admm = ms_admm * 1e-3;
    h->stats.seconds_total = ms_total * 1e-3;
    h->stats.seconds_eval = ms_eval * 1e-3;
    // algorithmic bytes of one evaluation of one block, two passes over X_b:
    const long long b_eval = 2ll * D.nb * D.dim * 8 + 2ll * D.nb * D.C * 8 + 12ll * D.nb +
                             2ll * D.C * D.dim * 8;
    h->stats.eval_bytes = last.total_evals * b_eval;
    h->stats.eval_mode = h->eval_mode;
    h->stats.ctas_per_block = D.G;
    h->stats.ring_stages = h->layout.NS;
    {   // tiles of the busiest CTA of a group (cta_row_range: whole tiles dealt out evenly)
        const long long tiles = (D.nb + h->tile_rows - 1) / h->tile_rows;
        h->stats.tiles_per_cta = (int)((tiles + D.G - 1) / D.G);
    }
    if (last.fatal)
        return set_error(SQW_ERR_NUMERIC, "NaN/Inf met by the solver (the reference would have thrown, LBFGS.java:792-814)");
    return SQW_OK;
}

int sqw_admm_get_stats(const sqw_admm *h, sqw_admm_stats *out)
{
    if (!h || !out)
        return set_error(SQW_ERR_INVALID_ARG, "null argument");
    *out = h->stats;
    return SQW_OK;
}

int sqw_admm_get_log(const sqw_admm *h, int32_t cap, int32_t *rows_out, int32_t *outer,
                     int32_t *itr, double *pho, double *fold, double *primal, double *dual,
                     int32_t *nfev, int32_t *admm_iters_per_outer)
{
    if (!h || !rows_out)
        return set_error(SQW_ERR_INVALID_ARG, "null argument");
    const int rows = (int)h->log_itr.size();
    *rows_out = rows;
    const int n = std::min(rows, (int)cap);
    const int TL = h->dev.T_local;
    for (int r = 0; r < n; ++r) {
        if (outer) outer[r] = h->log_outer[r];
        if (itr) itr[r] = h->log_itr[r];
        if (pho) pho[r] = h->log_pho[r];
        if (fold) fold[r] = h->log_fold[r];
        if (primal) primal[r] = h->log_primal[r];
        if (dual) dual[r] = h->log_dual[r];
        if (nfev)
            for (int b = 0; b < TL; ++b)
                nfev[(size_t)r * TL + b] = h->log_nfev[(size_t)r * TL + b];
    }
    if (admm_iters_per_outer)
        for (size_t i = 0; i < h->admm_iters.size(); ++i)
            admm_iters_per_outer[i] = h->admm_iters[i];
    return SQW_OK;
}

int sqw_admm_eval(sqw_admm *h, const double *cx, const double *smx, double pho, double *f_out,
                  double *g_out)
{
    if (!h || !cx || !smx || !f_out || !g_out)
        return set_error(SQW_ERR_INVALID_ARG, "null argument");
    SQW_CUDA_CHECK(cudaSetDevice(h->device));
    int rc = prepare(h);
    if (rc != SQW_OK)
        return rc;
    AdmmDev &D = h->dev;
    cudaStream_t st = h->stream;
    SQW_CUDA_CHECK(cudaMemsetAsync(D.xt, 0, sizeof(double) * (size_t)D.T_local * D.Cp * D.dimp, st));
    SQW_CUDA_CHECK(cudaMemsetAsync(D.ut, 0, sizeof(double) * (size_t)D.T_local * D.upad, st));
    SQW_CUDA_CHECK(cudaMemsetAsync(D.z, 0, sizeof(double) * D.upad, st));
    SQW_CUDA_CHECK(cudaMemsetAsync(D.bar, 0, sizeof(unsigned) * 32 * D.T_local * kNumPhases, st));
    for (int b = 0; b < D.T_local; ++b)
        if ((rc = upload_padded(h, cx + (size_t)b * D.C * D.dim, D.C, D.dim, D.dimp,
                                D.xt + (size_t)b * D.Cp * D.dimp)) != SQW_OK)
            return rc;
    if ((rc = upload_padded(h, smx, D.NN, D.dim, D.dimp, D.smx)) != SQW_OK)
        return rc;
    double *d_f = nullptr;
    SQW_CUDA_CHECK(cudaMalloc((void **)&d_f, sizeof(double) * D.T_local));
    SQW_CUDA_CHECK(cudaMemsetAsync(d_f, 0, sizeof(double) * D.T_local, st));
    EKernel ek = eval_kernel(kernel_nct(h), h->eval_mode);
    void *args[] = {&D, &h->layout, &pho, &d_f};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)ek, dim3(D.Gmax), dim3(mode_threads(h->eval_mode)),
                                                args, h->xupdate_smem, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(d_f);
        return set_error(SQW_ERR_CUDA, "eval kernel failed: %s", cudaGetErrorString(e));
    }
    SQW_CUDA_CHECK(cudaMemcpy(f_out, d_f, sizeof(double) * D.T_local, cudaMemcpyDeviceToHost));
    cudaFree(d_f);
    for (int b = 0; b < D.T_local; ++b)
        if ((rc = download_padded(h, D.g + (size_t)b * D.npad, D.C, D.dim, D.dimp,
                                  g_out + (size_t)b * D.C * D.dim)) != SQW_OK)
            return rc;
    return SQW_OK;
}

int sqw_lbfgs_rosenbrock(int32_t n, double *x_inout, double eps, int32_t *nfev_out,
                         int32_t *iters_out, int32_t *iflag_out)
{
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (n < 2 || (n & 1) || !x_inout)
        return set_error(SQW_ERR_INVALID_ARG, "n must be even and >= 2");
    const int G = std::max(1, std::min(8, n / 64));
    double *buf = nullptr;
    unsigned *bar = nullptr;
    int *out = nullptr;
    const size_t nv = (size_t)n;
    const size_t total = nv * (5 + 2 * kLbfgsM) + (size_t)G * kNumDots;
    SQW_CUDA_CHECK(cudaMalloc((void **)&buf, total * sizeof(double)));
    SQW_CUDA_CHECK(cudaMemset(buf, 0, total * sizeof(double)));
    SQW_CUDA_CHECK(cudaMalloc((void **)&bar, 128));
    SQW_CUDA_CHECK(cudaMemset(bar, 0, 128));
    SQW_CUDA_CHECK(cudaMalloc((void **)&out, 2 * sizeof(int)));
    SolveWs ws;
    ws.n = n;
    ws.x = buf;
    ws.g = buf + nv;
    ws.gold = buf + 2 * nv;
    ws.d = buf + 3 * nv;
    ws.wa = buf + 4 * nv;
    ws.S = buf + 5 * nv;
    ws.Y = buf + (5 + kLbfgsM) * nv;
    ws.spart = buf + (5 + 2 * kLbfgsM) * nv;
    ws.G = G;
    ws.cta = 0;
    ws.prof = nullptr;
    ws.slice = nullptr;  // set by the kernel
    SQW_CUDA_CHECK(cudaMemcpy(ws.x, x_inout, nv * sizeof(double), cudaMemcpyHostToDevice));
    void *args[] = {&ws, &bar, &eps, &out};
    cudaError_t e = cudaFuncSetAttribute(lbfgs_rosenbrock_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSolveSliceBytes);
    if (e == cudaSuccess)
        e = cudaLaunchCooperativeKernel((void *)lbfgs_rosenbrock_kernel, dim3(G), dim3(kThreads), args,
                                        kSolveSliceBytes, 0);
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    int res[2] = {0, 0};
    if (e == cudaSuccess)
        e = cudaMemcpy(res, out, sizeof res, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        e = cudaMemcpy(x_inout, ws.x, nv * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    cudaFree(bar);
    cudaFree(out);
    if (e != cudaSuccess)
        return set_error(SQW_ERR_CUDA, "rosenbrock kernel failed: %s", cudaGetErrorString(e));
    if (nfev_out) *nfev_out = res[0];
    if (iters_out) *iters_out = 0;
    if (iflag_out) *iflag_out = (res[1] == 0) ? 0 : (res[1] == 1 ? -1 : -100);
    return SQW_OK;
}

}  // extern "C"
