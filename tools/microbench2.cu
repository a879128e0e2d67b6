// DMMA issue rate vs resident warps per SM and vs number of independent accumulator chains.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
template <int CH>
__global__ void dmma_kernel(double *out, int iters, double a, double b)
{
    double c[CH][2];
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / ((double)iters * CH);
}
template <int CH> int run(int warps, double *out)
{
    const int iters = 4000;
    dmma_kernel<CH><<<148, warps * 32>>>(out, 10, 1.0000001, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); dmma_kernel<CH><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    double flops = 512.0 * CH * iters * 148.0 * warps;
    printf("chains=%2d warps/SM=%2d: %6.2f TFLOP/s, %.1f cycles per DMMA per warp\n", CH, warps, flops / ms / 1e9, h[1]);
    return 0;
}
int main()
{
    double *out; CK(cudaMalloc(&out, 64));
    for (int w : {1, 2, 4, 8, 16, 32}) { run<1>(w, out); }
    for (int w : {1, 2, 4, 8, 16, 32}) { run<2>(w, out); }
    for (int w : {1, 2, 4, 8, 16, 32}) { run<4>(w, out); }
    for (int w : {1, 2, 4, 8, 16, 32}) { run<8>(w, out); }
    return 0;
}
