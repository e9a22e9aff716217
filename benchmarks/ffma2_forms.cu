// ffma2_forms.cu — issue rate of the packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2) by operand form, in the
// register pattern of the filter bank's column pass: 16 accumulator pairs, 8 tap pairs, 8 inputs per block.
//   form 0: acc += W * {x, x}   ptxas folds the duplicated scalar into the FFMA2 ".F32" broadcast operand
//   form 1: acc += W * X        X a genuine register pair
//   form 2: acc += X * {w, w}   the repeated operand is the pair, the varying one a scalar (k-means score loop candidate)
// Prints one JSON line with TFLOP/s per form (2 FLOP per lane-FMA, 2 lanes per FFMA2).
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;

template <int FORM>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int iters)
{
    u64 acc[16], w[16];
    float x[8];
    u64 xp[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(in[i]), "f"(in[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w[i]) : "f"(in[16 + i]), "f"(in[17 + i]));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        x[u] = in[40 + u + (threadIdx.x & 1)];
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xp[u]) : "f"(in[48 + u]), "f"(in[49 + u + (threadIdx.x & 1)]));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const u64 ww = w[(i - u + 8) & 15];
                if (FORM == 2) {   // acc += X * {w, w}: the PAIR operand is the one that repeats (x[u] for 16 FMAs), the scalar varies
                    u64 b;
                    float wl, wh;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(wl), "=f"(wh) : "l"(ww));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(wl));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(xp[u]), "l"(b));
                } else if (FORM == 0) {
                    u64 b;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(x[u]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(ww), "l"(b));
                } else {
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(ww), "l"(xp[u]));
                }
            }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a, b;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i]));
        s += a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int FORM>
double run(float *out, const float *in, int blocks, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<FORM><<<blocks, 256>>>(out, in, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<FORM><<<blocks, 256>>>(out, in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return 2.0 * 2.0 * 128.0 * iters * 256.0 * blocks / (best * 1e-3) * 1e-12;
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int blocks = prop.multiProcessorCount * 8, iters = 2048;
    float *out, *in;
    cudaMalloc(&out, sizeof(float) * 256 * blocks);
    cudaMalloc(&in, sizeof(float) * 128);
    float h[128];
    for (int i = 0; i < 128; ++i) h[i] = 1e-3f * (i % 7) - 2e-3f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const double t0 = run<0>(out, in, blocks, iters), t1 = run<1>(out, in, blocks, iters), t2 = run<2>(out, in, blocks, iters);
    printf("{\"ffma2_scalar_operand_tflops\": %.2f, \"ffma2_pair_operand_tflops\": %.2f, \"ffma2_repeated_pair_varying_scalar_tflops\": %.2f}\n", t0, t1, t2);
    return cudaGetLastError() != cudaSuccess;
}
