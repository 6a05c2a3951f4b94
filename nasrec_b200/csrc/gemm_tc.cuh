// tcgen05 (5th-gen tensor core) path of the segment-list GEMM, sm_100a only.
//
// fp32 parity on tensor cores: every fp32 operand is split on the fly into
// hi = rn_tf32(a) and lo = rn_tf32(a - hi); the products hi*hi + lo*hi + hi*lo
// (+ lo*lo in 4-product mode) are accumulated by `tcgen05.mma.kind::tf32` into one
// fp32 accumulator in Tensor Memory.  Operand staging is done by the CTA's own
// producer warps (LDG -> split -> STS into the canonical K-major SWIZZLE_128B layout)
// rather than by TMA because (a) the weights keep the reference's state-dict layout,
// whose row stride (e.g. 1037 floats) is not 16-byte aligned, which cuTensorMap
// rejects, (b) the K axis is a list of segments living in different tensors, and
// (c) every element has to be touched anyway for the hi/lo split.
//
// Warp roles (160 threads): warps 0-3 = producers, then epilogue (tcgen05.ld of their
// own 32 TMEM lanes -> bias/addend -> global); warp 4 = TMEM allocator + single-thread
// MMA issuer.  mbarrier pipeline: full[s] (128 producer arrivals) / empty[s]
// (tcgen05.commit) / done (tcgen05.commit after the last k-tile).
#pragma once
#include "gemm_common.cuh"

namespace nasrec_gemm {

constexpr int TC_BM = 128;          // UMMA M (cta_group::1)
constexpr int TC_BK = 32;           // fp32 elements per k-tile = one 128-byte swizzle row
constexpr int TC_UK = 8;            // UMMA K for kind::tf32 (32 bytes)
constexpr int TC_THREADS = 160;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, canonical value 1) |
// SBO>>4 [32,46) = 1024 B between 8-row groups | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor: c_format F32=1 [4,6) | a_format TF32=2 [7,10) | b_format TF32=2 [10,13)
// | a_major K=0 [15] | b_major K=0 [16] | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
    // hi = a rounded to tf32 (10 explicit mantissa bits), lo = (a - hi) rounded to tf32;
    // a - hi is exact in fp32.  Pre-rounding makes the tensor core's own fp32->tf32
    // conversion (which drops the low 13 bits) a no-op.
    uint32_t u = __float_as_uint(a);
    uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
    hi = __uint_as_float(h);
    float r = a - hi;
    uint32_t l = (__float_as_uint(r) + 0x1000u) & 0xFFFFE000u;
    lo = __uint_as_float(l);
}

// byte offset of the 16-byte chunk (row, kc) inside a [rows][32 fp32] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int kc) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((kc ^ (row & 7)) << 4));
}

template <int ROWS>
__device__ __forceinline__ void tc_load_tile(const View& v, int i0, int I, int k0, int K, uint8_t* s_hi,
                                             uint8_t* s_lo, int tid, bool split) {
    constexpr int CHUNKS = ROWS * 8 / 128;   // 16-byte chunks per producer thread
    float4 val[CHUNKS];
#pragma unroll
    for (int q = 0; q < CHUNKS; ++q) {
        const int c = tid + 128 * q;
        int row, kc;
        if (v.contig_j) {
            row = c >> 3;
            kc = c & 7;
        } else {
            row = c % ROWS;
            kc = c / ROWS;
        }
        const int i = i0 + row, k = k0 + kc * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < I && k < K) {
            const float* p = v.p + voff(v, i, k);
            if (v.contig_j && k + 3 < K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                x = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                x.x = __ldg(p);
                if (k + 1 < K) x.y = __ldg(v.p + voff(v, i, k + 1));
                if (k + 2 < K) x.z = __ldg(v.p + voff(v, i, k + 2));
                if (k + 3 < K) x.w = __ldg(v.p + voff(v, i, k + 3));
            }
        }
        val[q] = x;
    }
#pragma unroll
    for (int q = 0; q < CHUNKS; ++q) {
        const int c = tid + 128 * q;
        int row, kc;
        if (v.contig_j) {
            row = c >> 3;
            kc = c & 7;
        } else {
            row = c % ROWS;
            kc = c / ROWS;
        }
        const uint32_t off = sw128_off(row, kc);
        float4 h, l;
        if (split) {
            split_tf32(val[q].x, h.x, l.x);
            split_tf32(val[q].y, h.y, l.y);
            split_tf32(val[q].z, h.z, l.z);
            split_tf32(val[q].w, h.w, l.w);
            *reinterpret_cast<float4*>(s_lo + off) = l;
        } else {
            h = val[q];
        }
        *reinterpret_cast<float4*>(s_hi + off) = h;
    }
}

template <int BN>
struct TcCfg {
    static constexpr int STAGES = BN == 128 ? 3 : 4;
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;      // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ Batch bt, int nprod) {
    using Cfg = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem_base;

    int z = blockIdx.z, pi = 0;
    for (; pi < bt.nprob; ++pi) {
        const int ns = bt.prob[pi].nsplit;
        if (z < ns) break;
        z -= ns;
    }
    if (pi >= bt.nprob) return;
    const Prob& pr = bt.prob[pi];
    const int split = z;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    if (m0 >= pr.M || n0 >= pr.N) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 1024-byte aligned tile area (SWIZZLE_128B atoms are 8 rows x 128 B)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    uint8_t* tiles = smem_raw + pad;
    uint8_t* bars = tiles + Cfg::STAGES * Cfg::STAGE_BYTES;
    const uint32_t bar_full = smem_u32(bars);                       // STAGES x 8 B
    const uint32_t bar_empty = bar_full + 8 * Cfg::STAGES;
    const uint32_t bar_done = bar_empty + 8 * Cfg::STAGES;

    // k-tile range of this split over the concatenated terms
    int tot = 0;
    for (int t = 0; t < pr.nterm; ++t) tot += (bt.term[pr.term0 + t].K + TC_BK - 1) / TC_BK;
    const int per = (tot + pr.nsplit - 1) / pr.nsplit;
    const int kt_begin = split * per;
    const int kt_end = min(tot, kt_begin + per);
    const int ntiles = max(0, kt_end - kt_begin);

    if (tid == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 128);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    if (warp < 4) {
        // ------------------------------------------------------------ producers
        int it = 0, kt = 0;
        for (int t = 0; t < pr.nterm && ntiles > 0; ++t) {
            const Term& tm = bt.term[pr.term0 + t];
            const int nk = (tm.K + TC_BK - 1) / TC_BK;
            if (kt + nk <= kt_begin) {
                kt += nk;
                continue;
            }
            if (kt >= kt_end) break;
            const int kb = max(0, kt_begin - kt), ke = min(nk, kt_end - kt);
            for (int kk = kb; kk < ke; ++kk, ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
                tc_load_tile<TC_BM>(tm.a, m0, pr.M, kk * TC_BK, tm.K, st, st + Cfg::A_BYTES, tid, nprod > 1);
                tc_load_tile<BN>(tm.b, n0, pr.N, kk * TC_BK, tm.K, st + 2 * Cfg::A_BYTES,
                                 st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, tid, nprod > 1);
                fence_proxy_async_smem();
                mbar_arrive(bar_full + 8 * s);
            }
            kt += nk;
        }
        // ------------------------------------------------------------ epilogue
        const int row = m0 + warp * 32 + lane;
        const int cmask = (1 << pr.c_sh_i) - 1;
        const long long ro = (long long)(row >> pr.c_sh_i) * pr.c_hi_i + (long long)(row & cmask) * pr.c_lo_i +
                             (long long)split * pr.split_stride;
        const long long ro_add = ro - (long long)split * pr.split_stride;
        if (ntiles > 0) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
        }
        // the pipeline buffers are idle now: reuse them as per-warp transpose scratch
        float* scratch = reinterpret_cast<float*>(tiles) + warp * (32 * 33);
        const bool transpose = pr.c_hi_j == 1;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= pr.N) break;
            uint32_t r[32];
            if (ntiles > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (transpose) {
                // C is row-major: go through shared memory so that a warp writes 128 contiguous bytes per row
#pragma unroll
                for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                const int n = n0 + c0 + lane;
                const float bias = (pr.bias && n < pr.N) ? __ldg(pr.bias + n) : 0.f;
                for (int rr = 0; rr < 32; ++rr) {
                    const int m = m0 + warp * 32 + rr;
                    if (m >= pr.M || n >= pr.N) continue;
                    const long long o = (long long)(m >> pr.c_sh_i) * pr.c_hi_i + (long long)(m & cmask) * pr.c_lo_i + n;
                    float v = scratch[rr * 33 + lane] + bias;
                    if (pr.addend) v += pr.addend[o];
                    pr.c[o + (long long)split * pr.split_stride] = v;
                }
                __syncwarp();
            } else if (row < pr.M) {
                // consecutive rows are adjacent in memory (sparse-axis projections): lanes already coalesce
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c0 + j;
                    if (n >= pr.N) break;
                    float v = __uint_as_float(r[j]);
                    if (pr.bias) v += __ldg(pr.bias + n);
                    const long long on = (long long)n * pr.c_hi_j;
                    if (pr.addend) v += pr.addend[ro_add + on];
                    pr.c[ro + on] = v;
                }
            }
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0 && ntiles > 0) {
            const uint32_t idesc = umma_idesc_tf32(BN);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + s * Cfg::STAGE_BYTES);
                const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + Cfg::A_BYTES);
                const uint64_t b_hi = umma_desc(sa + 2 * Cfg::A_BYTES);
                const uint64_t b_lo = umma_desc(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UK; ++k) {
                    const uint64_t adv = (uint64_t)((k * TC_UK * 4) >> 4);     // +32 B per UMMA_K inside the atom
                    const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
                    if (nprod > 1) {
                        // small cross terms first, the dominant hi*hi product last
                        umma_tf32(tmem, a_lo + adv, b_hi + adv, idesc, first);
                        umma_tf32(tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                        if (nprod > 3) umma_tf32(tmem, a_lo + adv, b_lo + adv, idesc, 1u);
                        umma_tf32(tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                    } else {
                        umma_tf32(tmem, a_hi + adv, b_hi + adv, idesc, first);
                    }
                }
                umma_commit(bar_empty + 8 * s);       // frees the stage once these MMAs retire
            }
            umma_commit(bar_done);                    // accumulator complete
        }
        __syncwarp();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
    }
}

template <int BN>
inline int launch_tc_bn(const Batch& bt, int maxM, int maxN, int totz, int nprod, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((maxN + BN - 1) / BN, (maxM + TC_BM - 1) / TC_BM, totz);
    if (grid.y > 65535 || grid.z > 65535) return NASREC_ETOOBIG;
    gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(bt, nprod);
    return (int)cudaGetLastError();
}

inline int launch_tc(const Batch& bt, int maxM, int maxN, int totz, int nprod, cudaStream_t st) {
    // narrow tiles when N is small or when 128-wide tiles would leave most SMs idle
    const long long tiles128 = (long long)((maxN + 127) / 128) * ((maxM + TC_BM - 1) / TC_BM) * totz;
    if (maxN <= 16) return launch_tc_bn<16>(bt, maxM, maxN, totz, nprod, st);
    if (maxN <= 64 || tiles128 < 148) return launch_tc_bn<64>(bt, maxM, maxN, totz, nprod, st);
    return launch_tc_bn<128>(bt, maxM, maxN, totz, nprod, st);
}

}  // namespace nasrec_gemm
