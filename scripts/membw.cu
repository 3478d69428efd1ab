// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/membw scripts/membw.cu
// Microbenchmark: HBM bandwidth of this GPU for read-only, write-only, copy and the return mapping's own mix
// (144 B read + 392 B written per element), 128-bit accesses, grid-stride, buffers far larger than L2.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const double2 *a, size_t n, double *out)
{
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { double2 v = a[i]; s += v.x + v.y; }
    if (s == 1.2345e-300) *out = s;
}
__global__ void k_write(double2 *a, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = make_double2(1., 2.);
}
__global__ void k_copy(const double2 *a, double2 *b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// per "element": 9 double2 read, 24.5 double2 written -> use 18 read / 49 written per pair of elements
__global__ void k_mix(const double2 *a, double2 *b, size_t npairs)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npairs * 49; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(1., 2.);
        if (i % 49 < 18) v = a[(i / 49) * 18 + i % 49];
        b[i] = v;
    }
}
template <class F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 10;
}
int main()
{
    const size_t n = (size_t)1 << 26;   // 64 Mi double2 = 1 GiB
    double2 *a, *b; double *o;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&o, 8);
    cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16);
    for (int bps : {8, 16, 32}) {
        const int g = 148 * bps;
        float t;
        t = timeit([&] { k_read<<<g, 256>>>(a, n, o); });   printf("blocks/SM %2d read  %.0f GB/s\n", bps, n * 16 / t / 1e6);
        t = timeit([&] { k_write<<<g, 256>>>(b, n); });     printf("blocks/SM %2d write %.0f GB/s\n", bps, n * 16 / t / 1e6);
        t = timeit([&] { k_copy<<<g, 256>>>(a, b, n); });   printf("blocks/SM %2d copy  %.0f GB/s (read+write)\n", bps, 2 * n * 16 / t / 1e6);
        const size_t np = n / 49;
        t = timeit([&] { k_mix<<<g, 256>>>(a, b, np); });   printf("blocks/SM %2d mix 144r:392w %.0f GB/s (read+write)\n", bps, np * (18 + 49) * 16 / t / 1e6);
    }
    return 0;
}
