// fma_peak.cu — FP32 CUDA-core FMA peak on this GPU: the roofline denominator of the Gabor
// bank (MEASURED_PEAKS.json has HBM and bf16 tensor peaks only).  Prints one JSON line.
//   mode 0: FFMA R, R, R, R          (three register operands)
//   mode 1: FFMA R, R, imm, R        (multiplier is a literal)
//   mode 2: fma.rn.f32x2             (packed pair, sm_100+)
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ACC = 16;

template <int MODE>
__global__ void __launch_bounds__(256) fma_kernel(float *out, const float *in, int iters)
{
    float a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = a * (i + 1);
    if (MODE == 2) {
        unsigned long long a2, b2, p[ACC / 2];
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
        for (int i = 0; i < ACC / 2; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < ACC / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(a2), "l"(b2));
        }
#pragma unroll
        for (int i = 0; i < ACC / 2; ++i) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < ACC; ++i) acc[i] = MODE == 0 ? fmaf(acc[i], a, b) : fmaf(acc[i], 0.999f, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(float *out, const float *in, int blocks, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_kernel<MODE><<<blocks, 256>>>(out, in, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<MODE><<<blocks, 256>>>(out, in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * ACC * 4.0 * iters * 256.0 * blocks;
    return flops / (best * 1e-3) * 1e-12;
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int blocks = prop.multiProcessorCount * 8, iters = 8192;
    float *out, *in;
    cudaMalloc(&out, sizeof(float) * 256 * blocks);
    cudaMalloc(&in, sizeof(float) * 64);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.001f * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const double t0 = run<0>(out, in, blocks, iters), t1 = run<1>(out, in, blocks, iters), t2 = run<2>(out, in, blocks, iters);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d, \"ffma_rrr_tflops\": %.2f, \"ffma_imm_tflops\": %.2f, "
           "\"ffma2_tflops\": %.2f, \"nominal_tflops\": %.2f}\n",
           prop.name, prop.multiProcessorCount, clk / 1000, t0, t1, t2,
           prop.multiProcessorCount * 128.0 * 2.0 * clk * 1e3 * 1e-12);
    return cudaGetLastError() != cudaSuccess;
}
