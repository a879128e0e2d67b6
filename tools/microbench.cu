// microbench.cu — B200 design probes for the ADMM evaluation kernel (not part of the product):
// fp64 FMA vs DMMA throughput, L2/HBM read bandwidth, barrier and launch latencies.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__global__ void dmma884_kernel(double *out, int iters, double a, double b)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void dmma1688_kernel(double *out, int iters, double a, double b)
{
    double c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}

__global__ void read_kernel(const double2 *__restrict__ p, size_t n2, double *out)
{
    double s = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride * 4) {
        double2 v0 = p[i];
        double2 v1 = (i + stride < n2) ? p[i + stride] : make_double2(0, 0);
        double2 v2 = (i + 2 * stride < n2) ? p[i + 2 * stride] : make_double2(0, 0);
        double2 v3 = (i + 3 * stride < n2) ? p[i + 3 * stride] : make_double2(0, 0);
        s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    if (s == 123.456) out[0] = s;
}

__global__ void gridsync_kernel(int iters)
{
    cg::grid_group g = cg::this_grid();
    for (int i = 0; i < iters; ++i) g.sync();
}

// groups of `gsize` consecutive CTAs synchronise on their own monotonically increasing counter
__global__ void groupbar_kernel(unsigned *counters, int gsize, int iters)
{
    const int grp = blockIdx.x / gsize;
    unsigned *ctr = counters + grp * 32;
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            const unsigned target = (unsigned)(i + 1) * gsize;
            while (*((volatile unsigned *)ctr) < target) { }
            __threadfence();
        }
        __syncthreads();
    }
}

__global__ void empty_kernel() {}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d smem/block optin %zu L2 %d MB clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin, prop.l2CacheSize >> 20, prop.clockRate);
    double *out; CK(cudaMalloc(&out, 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    const int nsm = prop.multiProcessorCount;
    {   // DFMA
        int iters = 20000;
        for (int threads : {256, 512, 1024}) {
            dfma_kernel<<<nsm * 2, threads>>>(out, 100, 1.0000001, 1e-9);
            CK(cudaEventRecord(e0)); dfma_kernel<<<nsm * 2, threads>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = 2.0 * 8 * iters * (double)nsm * 2 * threads;
            printf("DFMA  threads=%4d: %.2f TFLOP/s (%.3f ms)\n", threads, flops / ms / 1e9, ms);
        }
    }
    {   // DMMA m8n8k4: 2*8*8*4 = 512 flop per warp instr
        int iters = 20000;
        for (int threads : {256, 512, 1024}) {
            dmma884_kernel<<<nsm * 2, threads>>>(out, 100, 1.0000001, 1e-9);
            CK(cudaEventRecord(e0)); dmma884_kernel<<<nsm * 2, threads>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = 512.0 * 8 * iters * (double)nsm * 2 * threads / 32;
            printf("DMMA m8n8k4  threads=%4d: %.2f TFLOP/s (%.3f ms)\n", threads, flops / ms / 1e9, ms);
        }
        for (int threads : {256, 512, 1024}) {
            dmma1688_kernel<<<nsm * 2, threads>>>(out, 100, 1.0000001, 1e-9);
            CK(cudaEventRecord(e0)); dmma1688_kernel<<<nsm * 2, threads>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = 2.0 * 16 * 8 * 8 * 4 * iters * (double)nsm * 2 * threads / 32;
            printf("DMMA m16n8k8 threads=%4d: %.2f TFLOP/s (%.3f ms)\n", threads, flops / ms / 1e9, ms);
        }
    }
    {   // read bandwidth: 64 MB (L2 resident), 104 MB, 4 GB (HBM)
        for (size_t mb : {(size_t)32, (size_t)64, (size_t)104, (size_t)4096}) {
            size_t bytes = mb << 20; double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
            for (int blocks : {nsm * 2, nsm * 8}) {
                for (int w = 0; w < 3; ++w) read_kernel<<<blocks, 512>>>(buf, bytes / 16, out);
                int reps = mb > 1000 ? 5 : 50;
                CK(cudaEventRecord(e0));
                for (int r = 0; r < reps; ++r) read_kernel<<<blocks, 512>>>(buf, bytes / 16, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("read %5zu MB blocks=%5d: %.1f GB/s (%.2f us per pass)\n", mb, blocks, bytes * (double)reps / ms / 1e6, ms * 1e3 / reps);
            }
            CK(cudaFree(buf));
        }
    }
    {   // grid.sync latency
        int iters = 2000;
        for (int bps : {1, 2}) {
            for (int threads : {256, 1024}) {
                if (bps * threads > 2048) continue;
                void *args[] = {&iters};
                CK(cudaLaunchCooperativeKernel((void *)gridsync_kernel, dim3(nsm * bps), dim3(threads), args, 0, 0));
                CK(cudaEventRecord(e0));
                CK(cudaLaunchCooperativeKernel((void *)gridsync_kernel, dim3(nsm * bps), dim3(threads), args, 0, 0));
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("grid.sync blocks=%d threads=%d: %.3f us each\n", nsm * bps, threads, ms * 1e3 / iters);
            }
        }
    }
    {   // group barrier latency
        unsigned *ctr; CK(cudaMalloc(&ctr, 4096));
        int iters = 2000;
        for (int gsize : {1, 4, 8, 18, 37, 148}) {
            int groups = 148 / gsize; int blocks = groups * gsize;
            for (int threads : {256, 1024}) {
                CK(cudaMemset(ctr, 0, 4096));
                void *args[] = {&ctr, &gsize, &iters};
                CK(cudaEventRecord(e0));
                CK(cudaLaunchCooperativeKernel((void *)groupbar_kernel, dim3(blocks), dim3(threads), args, 0, 0));
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("group barrier gsize=%3d groups=%3d threads=%4d: %.3f us each\n", gsize, groups, threads, ms * 1e3 / iters);
            }
        }
    }
    {   // launch latency
        int n = 2000;
        empty_kernel<<<1, 32>>>(); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); for (int i = 0; i < n; ++i) empty_kernel<<<148, 256>>>(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("plain launch back-to-back: %.3f us each\n", ms * 1e3 / n);
        int iters0 = 0; void *args[] = {&iters0};
        CK(cudaEventRecord(e0)); for (int i = 0; i < n; ++i) CK(cudaLaunchCooperativeKernel((void *)gridsync_kernel, dim3(148), dim3(256), args, 0, 0)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("cooperative launch back-to-back: %.3f us each\n", ms * 1e3 / n);
    }
    return 0;
}
