// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 (M = 128, K = 8) as used by the 3xTF32 GEMM: one CTA per SM,
// one thread issues `n` MMAs on fixed operands (garbage data), every `group` MMAs followed by a tcgen05.commit, then waits.
// Reports clocks per MMA.  args (pairs): N <n-tile>  t <1: A from TMEM, 0: A from shared memory>  r <accumulators rotated
// per MMA>  g <MMAs per commit>  c <CTAs>
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_bench tools/mma_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni D;\n\tbra.uni W;\n\tD:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_w(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_w(uint32_t bar) {
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

__global__ void __launch_bounds__(128, 1) k(int N, int ts, int rot, int group, int n, long long* out, int use_elect) {
    extern __shared__ uint8_t raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t bar[16];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(raw + (base - smem_u32(raw)))[i] = 1.0f;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (use_elect == 2 && warp == 0) {
        // every lane of warp 0 runs the issue loop on warp-uniform operands; the election happens inside the asm block
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t bs = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t db = desc128(bs + 16384);
        const uint32_t a0 = tm + 256;
        const uint32_t bar0 = __shfl_sync(0xffffffffu, smem_u32(&bar[0]), 0);
        int nc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < n; i += 12) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                mma_ts_w(tm, a0 + 32 + kk * 8, db + adv, idesc, 1u);
                mma_ts_w(tm, a0 + kk * 8, db + 256 + adv, idesc, 1u);
                mma_ts_w(tm, a0 + kk * 8, db + adv, idesc, 1u);
            }
            if (nc >= 16) mbar_wait(bar0 + 8 * (nc % 16), (uint32_t)((nc - 16) / 16) & 1u);
            commit_w(bar0 + 8 * (nc % 16));
            ++nc;
        }
        for (int c = (nc > 16 ? nc - 16 : 0); c < nc; ++c) mbar_wait(bar0 + 8 * (c % 16), (uint32_t)(c / 16) & 1u);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    }
    uint32_t elected = 0;
    if (warp == 0) asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(elected));
    if (use_elect == 2 ? false : (use_elect ? (warp == 0 && elected) : (threadIdx.x == 0))) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = desc128(base), db = desc128(base + 16384);
        int nc = 0;        // commits issued so far; commit c goes to bar[c % 16], reused only after its previous phase completed
        auto do_commit = [&]() {
            if (nc >= 16) mbar_wait(smem_u32(&bar[nc % 16]), (uint32_t)((nc - 16) / 16) & 1u);
            commit(smem_u32(&bar[nc % 16]));
            ++nc;
        };
        const long long t0 = clock64();
        const uint32_t a0 = tmem + 256;
        for (int i = 0; i < n; i += 12) {          // one "k-tile": 4 k-steps x 3 products, descriptors advance by constants
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                const uint32_t d = tmem + (rot ? (uint32_t)(((i / 12) & 7) * 32) : 0u);
                if (ts) {
                    mma_ts(d, a0 + 32 + kk * 8, db + adv, idesc, 1u);
                    mma_ts(d, a0 + kk * 8, db + 256 + adv, idesc, 1u);
                    mma_ts(d, a0 + kk * 8, db + adv, idesc, 1u);
                } else {
                    mma_ss(d, da + 1024 + adv, db + adv, idesc, 1u);
                    mma_ss(d, da + adv, db + 256 + adv, idesc, 1u);
                    mma_ss(d, da + adv, db + adv, idesc, 1u);
                }
            }
            if (group) do_commit();
        }
        do_commit();
        for (int c = (nc > 16 ? nc - 16 : 0); c < nc; ++c) mbar_wait(smem_u32(&bar[c % 16]), (uint32_t)(c / 16) & 1u);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
    int N = 32, ts = 1, rot = 0, group = 12, n = 1200, ctas = 1, el = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        int v = atoi(argv[i + 1]);
        switch (argv[i][0]) { case 'N': N = v; break; case 't': ts = v; break; case 'r': rot = v; break; case 'g': group = v; break; case 'n': n = v; break; case 'c': ctas = v; break; case 'e': el = v; break; }
    }
    long long* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<<<ctas, 128, 64 * 1024>>>(N, ts, rot, group, n, out, el);
    k<<<ctas, 128, 64 * 1024>>>(N, ts, rot, group, n, out, el);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("elect %d N %3d A-from-%s rot %d group %2d ctas %3d: %6.1f clk per MMA (%lld clk for %d)  [%s]\n", el, N, ts ? "TMEM" : "smem", rot, group, ctas, (double)h / n, h, n,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
