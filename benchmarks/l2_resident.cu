// How fast can a working set that fits the 126 MB L2 be re-read?  (k-means reads one image's 44.5 MB of feature
// planes 20 times: would one-image-at-a-time passes run from L2?)   nvcc -O3 -arch=sm_100a -o build/l2_resident benchmarks/l2_resident.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) read_kernel(const uint4 *p, size_t n, unsigned *sink)
{
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
int main()
{
    const size_t cap = 512ull << 20;
    uint4 *buf; unsigned *sink;
    cudaMalloc(&buf, cap); cudaMalloc(&sink, 4); cudaMemset(buf, 1, cap);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int mbs[] = {8, 16, 32, 45, 64, 90, 110, 128, 192, 256, 512};
    for (int mb : mbs) {
        const size_t n = ((size_t)mb << 20) / 16;
        for (int grid : {148 * 4, 148 * 8}) {
            for (int w = 0; w < 3; ++w) read_kernel<<<grid, 256>>>(buf, n, sink);
            cudaEventRecord(e0);
            const int reps = 20;
            for (int r = 0; r < reps; ++r) read_kernel<<<grid, 256>>>(buf, n, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%4d MB grid %4d: %7.1f us per pass, %7.1f GB/s\n", mb, grid, ms * 1e3 / reps, (double)mb * 1048576.0 * reps / (ms * 1e-3) / 1e9);
        }
    }
    return cudaGetLastError() != cudaSuccess;
}
