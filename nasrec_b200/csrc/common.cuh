// Shared helpers for the nasrec_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <time.h>
#include "../../include/nasrec_b200.h"

#define NASREC_VERSION 100

#define CHECK_ARG(cond)            \
    do {                           \
        if (!(cond)) return NASREC_EINVAL; \
    } while (0)

static inline int nasrec_launch_status() {
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

static inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree); result valid in every thread.
// `red` must hold >= 33 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// Predicated loads as inline PTX.  A load inside an `if` ends a basic block, and the compiler then waits for
// it before the next block's loads are issued: an unrolled loop of guarded loads becomes a chain of
// full memory latencies.  These keep such loops branch-free so that all loads are in flight together.
__device__ __forceinline__ float ldg_nc_pred(const float* p, bool pred) {
    float v = 0.f;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "@q ld.global.nc.f32 %0, [%1];\n\t"
        "}\n"
        : "+f"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ float ld_pred(const float* p, bool pred) {     // coherent variant (buffers written by this kernel)
    float v = 0.f;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "@q ld.global.f32 %0, [%1];\n\t"
        "}\n"
        : "+f"(v)
        : "l"(p), "r"((int)pred));
    return v;
}

// ---- programmatic dependent launch -------------------------------------------------------------
// A training step is ~210 small dependent kernels; between two of them the GPU idles for the launch latency of
// the second (~1-2 us).  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_enter(): it lets its successor be scheduled at once (launch_dependents) and then
// waits for its own predecessors to finish and flush (wait) before touching global memory -- so the successor's
// launch, its CTA set-up and any prologue ahead of its wait overlap the tail of the running kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_trigger();
    pdl_wait();
}

// host-side launch accounting (nasrec_host_prof): nanoseconds spent inside cudaLaunchKernelEx and the number of launches
extern long long g_nasrec_launch_ns, g_nasrec_launch_count;
extern long long g_nasrec_launch_total;      // every kernel launch of the library since load (nasrec_host_prof(4))
extern int g_nasrec_host_prof;
// device-side trace (nasrec_host_prof(10 / 11)): CUDA events around EVERY kernel launch of the library on its own stream,
// aggregated by kernel name -- per-kernel time inside the real step (warm L2, both streams), which ncu's serialised
// cold-cache replays cannot give.  The event records sit between launches, so programmatic dependent launch does not
// overlap across them: the traced step is a little slower than the timed one; shares, not absolutes.
extern int g_nasrec_trace;
void nasrec_trace_begin(const void* func, cudaStream_t st);
void nasrec_trace_end(cudaStream_t st);
static inline long long nasrec_now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

// cluster_z > 1: the grid is launched as thread-block clusters of (1, 1, cluster_z) CTAs (grid.z must be a multiple).
template <typename... KArgs, typename... Args>
static inline cudaError_t nasrec_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                                int cluster_z, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cluster_z > 1) {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 1;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = (unsigned)cluster_z;
        cfg.numAttrs = 2;
    }
    ++g_nasrec_launch_total;
    if (g_nasrec_trace) {
        nasrec_trace_begin(reinterpret_cast<const void*>(kernel), st);
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        nasrec_trace_end(st);
        return e;
    }
    if (g_nasrec_host_prof) {
        const long long t0 = nasrec_now_ns();
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        g_nasrec_launch_ns += nasrec_now_ns() - t0;
        ++g_nasrec_launch_count;
        return e;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t nasrec_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ++g_nasrec_launch_total;
    if (g_nasrec_trace) {
        nasrec_trace_begin(reinterpret_cast<const void*>(kernel), st);
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        nasrec_trace_end(st);
        return e;
    }
    if (g_nasrec_host_prof) {
        const long long t0 = nasrec_now_ns();
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        g_nasrec_launch_ns += nasrec_now_ns() - t0;
        ++g_nasrec_launch_count;
        return e;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// library scratch attached with nasrec_set_workspace (gemm.cu); stream-ordered reuse by every kernel family
void nasrec_internal_workspace(float** ws, long long* nfloats);
// per-target accumulate flags for the NEXT nasrec_seg_linear_dgrad / nasrec_sproj_dgrad call (overrides its scalar
// `accumulate`; cleared by that call): lets the op-level backward entry points serve fresh and accumulated gradient
// targets with one launch
void nasrec_internal_set_dgrad_flags(const int* flags);
// deferred LayerNorm parameter gradients (ln.cu): scratch for first-stage partials (null detaches and drops the queue),
// number of recorded reductions, and the batched launch that finishes them
void nasrec_internal_ln_defer_scratch(float* base, long long nfloats);
long long nasrec_internal_ln_pending();
int nasrec_internal_ln_flush(cudaStream_t st);
// fork: returns the side stream ordered after everything issued to `main` so far (or `main` itself when none is attached);
// the work must be joined with nasrec_side_join
cudaStream_t nasrec_internal_fork_side(cudaStream_t main);
// optional second stream on which the op-level backward entry points issue weight-gradient work
cudaStream_t nasrec_internal_side_stream();
void nasrec_internal_set_side_stream(cudaStream_t s);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
