// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/dfma_lat scripts/dfma_lat.cu
// Microbenchmark: FP64 FMA dependent-issue latency and per-SM throughput on this GPU as a function of resident warps
// and independent chains per thread (decides whether the 50-sub-step loop wants occupancy or ILP).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc)
{
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
    const double b = 1.0000001, c = 1e-7;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<1, warps * 32>>>(out, iters, cyc);
    k<ILP><<<1, warps * 32>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d ILP %d : %.2f cycles per DFMA per thread-chain step, %.2f warp-DFMA/cycle/SM\n", warps, ILP,
           (double)h / iters, (double)warps * ILP * iters / h);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 4, 8, 12, 16, 24, 32}) { run<1>(w); run<2>(w); run<4>(w); }
    return 0;
}
