// Cross-SM exchange probes behind DESIGN.md section 6 (what bounds the solver phases):
//  A. group exchange through L2, the pattern of phase R: every CTA of a group of G stores its own
//     vector of n doubles (st.global.cg), group barrier (red.release / polled ld.acquire), then
//     thread t of the CTA reads coordinate (slice + t) of all G vectors (ld.global.cg) and sums;
//  B. the same reduction inside a thread-block cluster through distributed shared memory: every
//     CTA keeps its vector in shared memory, barrier.cluster, then reads its slice from the peers'
//     shared memory (ld.shared::cluster).
// Prints microseconds per round measured with %globaltimer by thread 0 of CTA 0 (store, barrier,
// loads) and the whole-kernel time per round.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench4 tools/microbench4.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__device__ __forceinline__ long long now_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kThreads = 352;

// ---------------------------------------------------------------- A: through L2
__global__ void __launch_bounds__(kThreads, 1)
l2_exchange(double *part, unsigned *bars, double *out, long long *prof, int n, int G, int rounds)
{
    extern __shared__ double pad[];  // occupies the SM: one CTA per SM
    const int grp = blockIdx.x / G, cta = blockIdx.x % G;
    double *mine = part + ((size_t)grp * G + cta) * n;
    const double *base = part + (size_t)grp * G * n;
    unsigned *ctr = bars + grp * 32;
    const int sl = ((n + G - 1) / G + 15) & ~15;
    const int k0 = min(n, cta * sl), k1 = min(n, k0 + sl);
    unsigned target = 0;
    long long t_store = 0, t_bar = 0, t_load = 0;
    double acc = 0.0;
    if (threadIdx.x == 0)
        pad[0] = 0.0;
    for (int r = 0; r < rounds; ++r) {
        long long t0 = (threadIdx.x == 0) ? now_ns() : 0;
        for (int i = threadIdx.x; i < n; i += kThreads)
            __stcg(mine + i, (double)(r + i + cta));
        long long t1 = (threadIdx.x == 0) ? now_ns() : 0;
        target += (unsigned)G;
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < target);
        }
        __syncthreads();
        long long t2 = (threadIdx.x == 0) ? now_ns() : 0;
        for (int k = k0 + threadIdx.x; k < k1; k += kThreads) {
            double s = 0.0;
            for (int c0 = 0; c0 < G; c0 += 24) {
                double v[24];
#pragma unroll
                for (int j = 0; j < 24; ++j)
                    v[j] = (c0 + j < G) ? __ldcg(base + (size_t)(c0 + j) * n + k) : 0.0;
#pragma unroll
                for (int j = 0; j < 24; ++j)
                    s += v[j];
            }
            acc += s;
        }
        // second barrier so that nobody overwrites a vector that is still being read
        target += (unsigned)G;
        __syncthreads();
        long long t3 = (threadIdx.x == 0) ? now_ns() : 0;
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < target);
            t_store += t1 - t0;
            t_bar += t2 - t1;
            t_load += t3 - t2;
        }
        __syncthreads();
    }
    if (acc == 12345.678)
        out[0] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        prof[0] = t_store;
        prof[1] = t_bar;
        prof[2] = t_load;
    }
}

// ---------------------------------------------------------------- B: through DSMEM in a cluster
template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
dsmem_exchange(double *out, long long *prof, int n, int rounds)
{
    extern __shared__ double vec[];  // [n]
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int sl = ((n + CL - 1) / CL + 15) & ~15;
    const int k0 = min(n, cta * sl), k1 = min(n, k0 + sl);
    long long t_store = 0, t_bar = 0, t_load = 0;
    double acc = 0.0;
    const double *peer[CL];
#pragma unroll
    for (int c = 0; c < CL; ++c)
        peer[c] = cluster.map_shared_rank(vec, c);
    for (int r = 0; r < rounds; ++r) {
        long long t0 = (threadIdx.x == 0) ? now_ns() : 0;
        for (int i = threadIdx.x; i < n; i += kThreads)
            vec[i] = (double)(r + i + cta);
        long long t1 = (threadIdx.x == 0) ? now_ns() : 0;
        cluster.sync();
        long long t2 = (threadIdx.x == 0) ? now_ns() : 0;
        for (int k = k0 + threadIdx.x; k < k1; k += kThreads) {
            double v[CL], s = 0.0;
#pragma unroll
            for (int c = 0; c < CL; ++c)
                v[c] = peer[c][k];
#pragma unroll
            for (int c = 0; c < CL; ++c)
                s += v[c];
            acc += s;
        }
        long long t3 = (threadIdx.x == 0) ? now_ns() : 0;
        cluster.sync();
        if (threadIdx.x == 0) {
            t_store += t1 - t0;
            t_bar += t2 - t1;
            t_load += t3 - t2;
        }
    }
    if (acc == 12345.678)
        out[0] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        prof[0] = t_store;
        prof[1] = t_bar;
        prof[2] = t_load;
    }
}

template <int CL> int run_dsmem(int n, int rounds, double *out, long long *prof)
{
    const size_t smem = (size_t)n * 8;
    CK(cudaFuncSetAttribute(dsmem_exchange<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8)
        CK(cudaFuncSetAttribute(dsmem_exchange<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    const int grid = (144 / CL) * CL;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_exchange<CL>, out, prof, n, 10);
    if (e != cudaSuccess) {
        printf("B. cluster %2d: launch failed: %s\n", CL, cudaGetErrorString(e));
        cudaGetLastError();
        return 0;
    }
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, dsmem_exchange<CL>, out, prof, n, rounds));
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[3];
    CK(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
    printf("B. cluster of %2d CTAs, n = %d doubles in shared memory: per round %.2f us (fill %.2f, cluster barrier %.2f, "
           "DSMEM loads of the slice %.2f)\n",
           CL, n, ms * 1e3 / rounds, h[0] * 1e-3 / rounds, h[1] * 1e-3 / rounds, h[2] * 1e-3 / rounds);
    return 0;
}

int main()
{
    const int n = 5216, rounds = 2000;
    double *part, *out;
    unsigned *bars;
    long long *prof;
    CK(cudaMalloc(&part, (size_t)144 * n * 8));
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&bars, 144 * 32 * 4));
    CK(cudaMalloc(&prof, 64));
    const size_t pad = 150 * 1024;
    CK(cudaFuncSetAttribute(l2_exchange, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
    for (int G : {18, 36, 48, 72, 144}) {
        CK(cudaMemset(bars, 0, 144 * 32 * 4));
        int nn = n, r = 10;
        void *args[] = {&part, &bars, &out, &prof, &nn, &G, &r};
        CK(cudaLaunchCooperativeKernel((void *)l2_exchange, dim3(144), dim3(kThreads), args, pad, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(bars, 0, 144 * 32 * 4));
        r = rounds;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        CK(cudaLaunchCooperativeKernel((void *)l2_exchange, dim3(144), dim3(kThreads), args, pad, 0));
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long h[3];
        CK(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
        printf("A. L2 exchange, 144 CTAs in groups of %3d, n = %d doubles: per round %.2f us (store own vector %.2f, "
               "barrier %.2f, loads of G partials per coordinate %.2f, + closing barrier)\n",
               G, n, ms * 1e3 / rounds, h[0] * 1e-3 / rounds, h[1] * 1e-3 / rounds, h[2] * 1e-3 / rounds);
    }
    if (run_dsmem<2>(n, rounds, out, prof)) return 1;
    if (run_dsmem<4>(n, rounds, out, prof)) return 1;
    if (run_dsmem<8>(n, rounds, out, prof)) return 1;
    if (run_dsmem<16>(n, rounds, out, prof)) return 1;
    return 0;
}
