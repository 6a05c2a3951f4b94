// Micro-benchmark: how fast does TMA deliver GEMM operand tiles from L2?  Every CTA loops over `ktiles` k-tiles and
// fetches an A box {kw floats, rowsA} and `nb` B boxes {kw floats, rowsB} per tile through an S-stage mbarrier ring;
// one consumer thread frees each stage as soon as it has landed.  Prints us per launch and the implied per-CTA and
// aggregate rates.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_bench tools/tma_bench.cu
// args (pairs): k ktiles, a rowsA, b rowsB, n boxes of B, s stages, t N-tiles per M-tile, m M-tiles, w k-width (floats)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni D;\n\tbra.uni W;\n\tD:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

struct P { int ktiles, rowsA, rowsB, nb, S, ntile_n, kw; };

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, P p) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int stage_bytes = (p.rowsA + p.nb * p.rowsB) * p.kw * 4;
    const uint32_t bars = base + p.S * stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.S; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (p.S + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int m0 = (blockIdx.x / p.ntile_n) * p.rowsA, n0 = (blockIdx.x % p.ntile_n) * p.rowsB;
    if (threadIdx.x == 0) {
        for (int it = 0; it < p.ktiles; ++it) {
            const int s = it % p.S; const uint32_t ph = (it / p.S) & 1;
            mbar_wait(bars + 8 * (p.S + s), ph ^ 1);
            mbar_expect(bars + 8 * s, stage_bytes);
            uint32_t dst = base + s * stage_bytes;
            tma2d(dst, &ma, bars + 8 * s, it * p.kw, m0);
            dst += p.rowsA * p.kw * 4;
            for (int b = 0; b < p.nb; ++b) { tma2d(dst, &mb, bars + 8 * s, it * p.kw, n0 + b * 4096); dst += p.rowsB * p.kw * 4; }
        }
    } else if (threadIdx.x == 32) {
        for (int it = 0; it < p.ktiles; ++it) {
            const int s = it % p.S; const uint32_t ph = (it / p.S) & 1;
            mbar_wait(bars + 8 * s, ph);
            mbar_arrive(bars + 8 * (p.S + s));
        }
    }
}

int main(int argc, char** argv) {
    P p; p.ktiles = 33; p.rowsA = 128; p.rowsB = 32; p.nb = 2; p.S = 8; p.ntile_n = 32; p.kw = 32;
    int mtiles = 4;
    for (int i = 1; i + 1 < argc; i += 2) {
        int v = atoi(argv[i + 1]);
        switch (argv[i][0]) { case 'k': p.ktiles = v; break; case 'a': p.rowsA = v; break; case 'b': p.rowsB = v; break; case 'n': p.nb = v; break;
                              case 's': p.S = v; break; case 't': p.ntile_n = v; break; case 'm': mtiles = v; break; case 'w': p.kw = v; break; }
    }
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    if (p.kw > 32) { printf("kw > 32 needs several boxes; not supported\n"); return 1; }
    const int K = p.ktiles * p.kw, ld = K + 16;
    float *A, *B;
    cudaMalloc(&A, (size_t)mtiles * p.rowsA * ld * 4); cudaMalloc(&B, (size_t)8192 * ld * 4);
    cudaMemset(A, 0, (size_t)mtiles * p.rowsA * ld * 4); cudaMemset(B, 0, (size_t)8192 * ld * 4);
    CUtensorMap ma, mb;
    cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)mtiles * p.rowsA}, gs[1] = {(cuuint64_t)ld * 4}; cuuint32_t bx[2] = {(cuuint32_t)p.kw, (cuuint32_t)p.rowsA}, es[2] = {1, 1};
    CUtensorMapSwizzle sw = p.kw * 4 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc A failed\n"); return 1; }
    cuuint64_t gdb[2] = {(cuuint64_t)K, 8192}; cuuint32_t bxb[2] = {(cuuint32_t)p.kw, (cuuint32_t)p.rowsB};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, B, gdb, gs, bxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc B failed\n"); return 1; }
    const int stage_bytes = (p.rowsA + p.nb * p.rowsB) * p.kw * 4;
    const int smem = p.S * stage_bytes + 1024 + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int grid = mtiles * p.ntile_n;
    for (int i = 0; i < 3; ++i) k<<<grid, 64, smem>>>(ma, mb, p);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 50;
    for (int i = 0; i < reps; ++i) k<<<grid, 64, smem>>>(ma, mb, p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / reps;
    printf("grid %d ktiles %d rowsA %d rowsB %dx%d kw %d S %d: %.1f us/launch, %.3f us per k-tile per CTA, %.1f GB/s per CTA, %.2f TB/s total (%s)\n", grid, p.ktiles, p.rowsA,
           p.rowsB, p.nb, p.kw, p.S, us, us / p.ktiles, stage_bytes / (us / p.ktiles) / 1e3, (double)grid * p.ktiles * stage_bytes / us / 1e6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
