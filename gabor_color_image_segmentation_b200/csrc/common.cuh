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

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace gcis
