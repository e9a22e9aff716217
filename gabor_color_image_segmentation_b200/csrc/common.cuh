// common.cuh — shared helpers for libgcis (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/gcis.h"

namespace gcis {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

int set_error(int code, const char *fmt, ...);

#define GCIS_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return gcis::set_error(GCIS_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,  \
                                   cudaGetErrorString(_e));                                   \
    } while (0)

#define GCIS_LAUNCH_CHECK()                                                                   \
    do {                                                                                      \
        gcis::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return gcis::set_error(GCIS_E_CUDA, "%s:%d kernel launch -> %s", __FILE__,        \
                                   __LINE__, cudaGetErrorString(_e));                         \
    } while (0)

// scipy.ndimage 'reflect' folding (d c b a | a b c d | d c b a), any offset.
__host__ __device__ __forceinline__ int reflect_index(int i, int n)
{
    if (n == 1) return 0;
    int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// Largest dynamic shared-memory size already enabled for one kernel, per device ordinal
// (cudaFuncSetAttribute is a per-device setting; the host side drives one device per process,
// but a process that switches devices must not inherit another device's state).
struct SmemAttrCache {
    size_t v[64] = {};
    size_t &cur()
    {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & 63];
    }
};

// Exact integer moments of a feature plane for the normalisation, per plane GB_STAT_SLOTS int64, with
// r = rint(x 2^16):   [0] sum r   [1] sum of the low 32 bits of the partial sums of r^2   [2] sum of their high bits.
// A thread accumulates a few dozen values (|r| < 2^23 for |x| < 128: sum r fits 32 bits, sum r^2 fits 64), a warp
// flushes its total split in two words, so that the global sums stay exact for images of up to 2^25 pixels.
constexpr int GB_STAT_SLOTS = 3;
#ifdef __CUDACC__
__device__ __forceinline__ void stat_add(float v, int &m1, unsigned long long &m2)
{
    const int r = __float2int_rn(v * 65536.0f);
    m1 += r;
    m2 += (unsigned long long)((long long)r * r);      // one 32 x 32 -> 64 multiply-add
}
__device__ __forceinline__ void stat_flush(long long *dst, long long m1, unsigned long long m2, int lane)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m1 += __shfl_down_sync(0xffffffffu, m1, o);
        m2 += __shfl_down_sync(0xffffffffu, m2, o);
    }
    if (lane == 0 && (m1 | (long long)m2)) {
        unsigned long long *st = reinterpret_cast<unsigned long long *>(dst);
        atomicAdd(st, (unsigned long long)m1);
        atomicAdd(st + 1, m2 & 0xffffffffull);
        atomicAdd(st + 2, m2 >> 32);
    }
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace gcis
