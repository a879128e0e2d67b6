// DMMA issue rate with distinct A/B operand registers per instruction (no operand reuse).
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
template <int MODE>
__global__ void dmma_kernel(double *out, const double *in, int iters)
{
    double c[8][2], a[8], b[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; a[i] = in[i + threadIdx.x % 7]; b[i] = in[8 + i + threadIdx.x % 5]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double aa = MODE == 0 ? a[0] : a[i];
            const double bb = MODE == 0 ? b[0] : (MODE == 1 ? b[0] : b[i]);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(aa), "d"(bb));
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / ((double)iters * 8);
}
template <int MODE> int run(int warps, double *out, double *in)
{
    const int iters = 4000;
    dmma_kernel<MODE><<<148, warps * 32>>>(out, in, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); dmma_kernel<MODE><<<148, warps * 32>>>(out, in, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    double flops = 512.0 * 8 * iters * 148.0 * warps;
    printf("mode %d (0: same a,b; 1: a varies; 2: a,b vary) warps/SM=%2d: %6.2f TFLOP/s, %.1f cycles per DMMA per warp\n", MODE, warps, flops / ms / 1e9, h[1]);
    return 0;
}
int main()
{
    double *out, *in; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&in, 1024));
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-9 * i; CK(cudaMemcpy(in, h, 1024, cudaMemcpyHostToDevice));
    for (int w : {4, 8, 12}) run<0>(w, out, in);
    for (int w : {4, 8, 12}) run<1>(w, out, in);
    for (int w : {4, 8, 12}) run<2>(w, out, in);
    return 0;
}
