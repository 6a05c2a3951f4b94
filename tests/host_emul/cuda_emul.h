// Host emulation of the CUDA execution model for tests (CPU only): one OS thread per CUDA thread of a CTA,
// __syncthreads() = a pthread barrier, __shared__ = process-global storage, CTAs run one after the other.  Enough for
// kernels that use neither warp shuffles nor atomics nor asynchronous copies.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <pthread.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static
#define __constant__
#define __grid_constant__
#define __launch_bounds__(...)

struct Dim3 {
    int x = 0, y = 0, z = 0;
};
static thread_local Dim3 threadIdx;
static Dim3 blockIdx, gridDim, blockDim;
static pthread_barrier_t g_bar;
static float* g_dyn = nullptr;                       // dynamic shared memory of the running CTA
static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline void pdl_enter() {}
struct float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }


template <class F>
static void run_grid(int grid, int nthreads, size_t dyn_floats, F body) {
    std::vector<float> dyn(dyn_floats ? dyn_floats : 1);
    g_dyn = dyn.data();
    gridDim.x = grid;
    blockDim.x = nthreads;
    for (int b = 0; b < grid; ++b) {
        blockIdx.x = b;
        pthread_barrier_init(&g_bar, nullptr, nthreads);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] {
                threadIdx.x = t;
                body();
            });
        for (auto& x : th) x.join();
        pthread_barrier_destroy(&g_bar);
    }
}


static unsigned g_seed = 12345;
static float rnd() {
    g_seed = g_seed * 1664525u + 1013904223u;
    return ((g_seed >> 8) & 0xffff) / 65536.0f - 0.5f;
}

static int check(const char* what, const std::vector<float>& a, const std::vector<float>& b) {
    if (a.size() != b.size()) return 1;
    int bad = 0;
    double mx = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        if (std::memcmp(&a[i], &b[i], 4) != 0) {
            ++bad;
            mx = std::fmax(mx, std::fabs((double)a[i] - b[i]));
        }
        if (!std::isfinite(a[i])) ++bad;
    }
    std::printf("  %-6s %zu values, %d differ (max abs %.3g)\n", what, a.size(), bad, mx);
    return bad;
}
