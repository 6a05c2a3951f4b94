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
// Warp roles (288 threads): warps 0-7 = producers; warps 0-3 then run the epilogue (tcgen05.ld of
// their own 32 TMEM lanes -> sum of the round-robin accumulators -> bias/addend -> global);
// warp 8 = TMEM allocator + single-thread MMA issuer.  mbarrier pipeline: full[s] (256 producer
// arrivals) / empty[s] (tcgen05.commit) / done (tcgen05.commit after the last k-tile).
#pragma once
#include "gemm_common.cuh"

namespace nasrec_gemm {

constexpr int TC_BM = 128;          // UMMA M (cta_group::1)
constexpr int TC_BK = 32;           // fp32 elements per k-tile = one 128-byte swizzle row
constexpr int TC_UK = 8;            // UMMA K for kind::tf32 (32 bytes)
constexpr int TC_PRODUCERS = 256;   // warps 0-7 stage operands; warps 0-3 also run the epilogue
constexpr int TC_THREADS = TC_PRODUCERS + 32;

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
    // hi = a rounded to nearest tf32 (10 explicit mantissa bits); lo = a - hi, exact in fp32 and
    // |lo| <= 2^-11 |a|.  The tensor core drops the low 13 mantissa bits of lo itself, an error of
    // at most 2^-21 |a| -- below the fp32 rounding of the product sum it is accumulated into.
    const uint32_t h = (__float_as_uint(a) + 0x1000u) & 0xFFFFE000u;
    hi = __uint_as_float(h);
    lo = a - hi;
}

// byte offset of the 16-byte chunk (row, kc) inside a [rows][32 fp32] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int kc) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((kc ^ (row & 7)) << 4));
}

// Predicated loads as inline PTX: no branches, so all loads of a tile are in flight together.
__device__ __forceinline__ float ldg_pred(const float* p, bool pred) {
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
__device__ __forceinline__ float4 ldg128_pred(const float* p, bool pred) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
        "}\n"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(p), "r"((int)pred));
    return v;
}

__device__ __forceinline__ void store_split4(uint8_t* s_hi, uint8_t* s_lo, uint32_t off, const float4& x, bool split) {
    float4 h, l;
    if (split) {
        split_tf32(x.x, h.x, l.x);
        split_tf32(x.y, h.y, l.y);
        split_tf32(x.z, h.z, l.z);
        split_tf32(x.w, h.w, l.w);
        *reinterpret_cast<float4*>(s_lo + off) = l;
    } else {
        h = x;
    }
    *reinterpret_cast<float4*>(s_hi + off) = h;
}

// One [ROWS x 32] fp32 k-tile travels global -> registers (tile_ldg) -> (hi, lo) -> swizzled shared
// memory (tile_sts), moved by the 256 producer threads.  The two halves are separate so that the
// loads of k-tile t+1 are in flight while k-tile t is split and stored (register double buffering).
// Three block-uniform layouts, each branch-free inside:
//   0 vec16    K-contiguous rows that are 16-byte aligned: one LDG.128 per 16-byte chunk;
//   1 contig_j K-contiguous but unaligned (weights keep the reference's odd row strides): one warp
//              instruction reads one whole 128-byte tile row (lane = k), scalar STS;
//   2 contig_i M/N-contiguous (dgrad/wgrad operands, sparse-axis tensors): lane = row, 4 k per thread.
// PLAIN views (no two-level index) use pointer arithmetic instead of the general offset formula.
template <int ROWS>
struct TileRegs {
    static constexpr int NV = (ROWS * 32 / 256 + 3) / 4 > 0 ? (ROWS * 32 / 256 + 3) / 4 : 1;   // float4 per thread
    float4 v[NV];
    int kind;
};

template <int ROWS, bool PLAIN>
__device__ __forceinline__ void tile_ldg_impl(const View& v, int i0, int I, int k0, int K, int tid,
                                              TileRegs<ROWS>& R) {
    auto at = [&](int i, int k) -> const float* {
        if (PLAIN) return v.p + (long long)i * v.hi_i + (long long)k * v.hi_j;
        return v.p + voff(v, i, k);
    };
    constexpr int NV = TileRegs<ROWS>::NV;
    if (v.contig_j && v.vec16 && k0 + TC_BK <= K) {
        R.kind = 0;
        const int r0 = tid >> 3, kc = tid & 7;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int row = r0 + 32 * q;
            R.v[q] = ldg128_pred(at(i0 + row, k0 + kc * 4), row < ROWS && i0 + row < I);
        }
    } else if (v.contig_j) {
        R.kind = 1;
        constexpr int RPW = (ROWS + 7) / 8;                   // rows per producer warp
        const int w = tid >> 5, lane = tid & 31;
        const bool kok = k0 + lane < K;
        float* f = reinterpret_cast<float*>(R.v);
#pragma unroll
        for (int r = 0; r < NV * 4; ++r) {
            const int row = w * RPW + r;
            f[r] = ldg_pred(at(i0 + row, k0 + lane), r < RPW && kok && row < ROWS && i0 + row < I);
        }
    } else {
        R.kind = 2;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int c = tid + 256 * q, row = c % ROWS, kc = c / ROWS;
            const int i = i0 + row, k = k0 + kc * 4;
            const bool ok = i < I && kc < 8;
            if (PLAIN) {
                const float* p = at(i, k);
                const long long sj = v.hi_j;
                R.v[q].x = ldg_pred(p, ok && k < K);
                R.v[q].y = ldg_pred(p + sj, ok && k + 1 < K);
                R.v[q].z = ldg_pred(p + 2 * sj, ok && k + 2 < K);
                R.v[q].w = ldg_pred(p + 3 * sj, ok && k + 3 < K);
            } else {
                R.v[q].x = ldg_pred(at(i, k), ok && k < K);
                R.v[q].y = ldg_pred(at(i, k + 1), ok && k + 1 < K);
                R.v[q].z = ldg_pred(at(i, k + 2), ok && k + 2 < K);
                R.v[q].w = ldg_pred(at(i, k + 3), ok && k + 3 < K);
            }
        }
    }
}

template <int ROWS>
__device__ __forceinline__ void tile_ldg(const View& v, int i0, int I, int k0, int K, int tid, TileRegs<ROWS>& R) {
    if (v.sh_i == 0 && v.sh_j == 0) tile_ldg_impl<ROWS, true>(v, i0, I, k0, K, tid, R);
    else tile_ldg_impl<ROWS, false>(v, i0, I, k0, K, tid, R);
}

template <int ROWS>
__device__ __forceinline__ void tile_sts(const TileRegs<ROWS>& R, uint8_t* s_hi, uint8_t* s_lo, int tid, bool split) {
    constexpr int NV = TileRegs<ROWS>::NV;
    if (R.kind == 0) {
        const int r0 = tid >> 3, kc = tid & 7;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int row = r0 + 32 * q;
            if (row < ROWS) store_split4(s_hi, s_lo, sw128_off(row, kc), R.v[q], split);
        }
    } else if (R.kind == 1) {
        constexpr int RPW = (ROWS + 7) / 8;
        const int w = tid >> 5, lane = tid & 31;
        const float* f = reinterpret_cast<const float*>(R.v);
#pragma unroll
        for (int r = 0; r < NV * 4; ++r) {
            const int row = w * RPW + r;
            if (r < RPW && row < ROWS) {
                const uint32_t off = sw128_off(row, lane >> 2) + (uint32_t)((lane & 3) << 2);
                float h = f[r], l = 0.f;
                if (split) split_tf32(f[r], h, l);
                *reinterpret_cast<float*>(s_hi + off) = h;
                if (split) *reinterpret_cast<float*>(s_lo + off) = l;
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int c = tid + 256 * q, row = c % ROWS, kc = c / ROWS;
            if (kc < 8) store_split4(s_hi, s_lo, sw128_off(row, kc), R.v[q], split);
        }
    }
}

template <int BN>
struct TcCfg {
    static constexpr int STAGES = BN == 128 ? 3 : 4;     // 192 KB / 192 KB / 160 KB / 144 KB of shared memory
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;      // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    // Round-robin accumulators: k-tile `it` accumulates into TMEM accumulator it % nacc and the
    // epilogue adds the accumulators in fp32 registers (round-to-nearest).  The tensor core's own
    // fp32 accumulation truncates, so its error grows linearly with the number of MMAs chained
    // into one accumulator; spreading K over up to 8 accumulators brings it back to FFMA level.
    static constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;
    static constexpr int NACC_MAX = 512 / ACC_STRIDE < 8 ? 512 / ACC_STRIDE : 8;
    static constexpr int TMEM_COLS = 512;
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
    const int nacc = min(Cfg::NACC_MAX, ntiles);

    if (tid == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS / 32);   // one arrival per producer warp
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    if (warp < 8) {
        // ------------------------------------------------------------ producers
        // flat walk over this split's k-tiles: (term index, k-tile inside the term)
        int t_cur = 0, kk_cur = 0;
        {
            int kt = 0;
            for (; t_cur < pr.nterm; ++t_cur) {
                const int nk = (bt.term[pr.term0 + t_cur].K + TC_BK - 1) / TC_BK;
                if (kt + nk > kt_begin) {
                    kk_cur = kt_begin - kt;
                    break;
                }
                kt += nk;
            }
        }
        TileRegs<TC_BM> ra[2];
        TileRegs<BN> rb[2];
        if (ntiles > 0) {
            const Term& tm = bt.term[pr.term0 + t_cur];
            tile_ldg<TC_BM>(tm.a, m0, pr.M, kk_cur * TC_BK, tm.K, tid, ra[0]);
            tile_ldg<BN>(tm.b, n0, pr.N, kk_cur * TC_BK, tm.K, tid, rb[0]);
        }
#pragma unroll 1
        for (int it = 0; it < ntiles; it += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cur = it + h;
                if (cur < ntiles) {
                    // advance to the next k-tile and put its loads in flight before touching `cur`'s data
                    int t_nxt = t_cur, kk_nxt = kk_cur + 1;
                    if (kk_nxt * TC_BK >= bt.term[pr.term0 + t_cur].K) {
                        ++t_nxt;
                        kk_nxt = 0;
                    }
                    if (cur + 1 < ntiles) {
                        const Term& tn = bt.term[pr.term0 + t_nxt];
                        tile_ldg<TC_BM>(tn.a, m0, pr.M, kk_nxt * TC_BK, tn.K, tid, ra[h ^ 1]);
                        tile_ldg<BN>(tn.b, n0, pr.N, kk_nxt * TC_BK, tn.K, tid, rb[h ^ 1]);
                    }
                    const int s = cur % Cfg::STAGES;
                    const uint32_t ph = (uint32_t)(cur / Cfg::STAGES) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
                    tile_sts<TC_BM>(ra[h], st, st + Cfg::A_BYTES, tid, nprod > 1);
                    tile_sts<BN>(rb[h], st + 2 * Cfg::A_BYTES, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, tid, nprod > 1);
                    fence_proxy_async_smem();          // every writer orders its generic-proxy stores
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * s);
                    t_cur = t_nxt;
                    kk_cur = kk_nxt;
                }
            }
        }
        // ------------------------------------------------------------ epilogue (warps 0-3: TMEM lane quadrants)
        if (warp < 4) {
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
            float r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0.f;
            for (int a = 0; a < nacc; ++a) {
                uint32_t u[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * Cfg::ACC_STRIDE + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]),
                      "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]),
                      "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]),
                      "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] += __uint_as_float(u[j]);
            }
            if (transpose) {
                // C is row-major: go through shared memory so that a warp writes 128 contiguous bytes per row
#pragma unroll
                for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = r[j];
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
                    float v = r[j];
                    if (pr.bias) v += __ldg(pr.bias + n);
                    const long long on = (long long)n * pr.c_hi_j;
                    if (pr.addend) v += pr.addend[ro_add + on];
                    pr.c[ro + on] = v;
                }
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
                const uint32_t tacc = tmem + (uint32_t)((it % nacc) * Cfg::ACC_STRIDE);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UK; ++k) {
                    const uint64_t adv = (uint64_t)((k * TC_UK * 4) >> 4);     // +32 B per UMMA_K inside the atom
                    const uint32_t first = (it >= nacc || k > 0) ? 1u : 0u;    // first touch of this accumulator
                    if (nprod > 1) {
                        // small cross terms first, the dominant hi*hi product last
                        umma_tf32(tacc, a_lo + adv, b_hi + adv, idesc, first);
                        umma_tf32(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
                        if (nprod > 3) umma_tf32(tacc, a_lo + adv, b_lo + adv, idesc, 1u);
                        umma_tf32(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
                    } else {
                        umma_tf32(tacc, a_hi + adv, b_hi + adv, idesc, first);
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
    if (warp == 8) {
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

// CTAs a launch would use with N-tile width bn
inline long long tc_cta_count(const Batch& bt, int bn) {
    long long c = 0;
    for (int i = 0; i < bt.nprob; ++i)
        c += (long long)((bt.prob[i].M + TC_BM - 1) / TC_BM) * ((bt.prob[i].N + bn - 1) / bn) * bt.prob[i].nsplit;
    return c;
}

// widest N tile that still spreads the launch over most of the 148 SMs (one CTA per SM)
inline int tc_pick_bn(const Batch& bt, int maxN) {
    if (maxN <= 16) return 16;
    if (maxN > 64 && tc_cta_count(bt, 128) >= 120) return 128;
    if (maxN > 32 && tc_cta_count(bt, 64) >= 120) return 64;
    return maxN <= 32 ? 32 : (tc_cta_count(bt, 32) > 296 ? 64 : 32);
}

inline int launch_tc(const Batch& bt, int maxM, int maxN, int totz, int nprod, cudaStream_t st) {
    switch (tc_pick_bn(bt, maxN)) {
        case 16: return launch_tc_bn<16>(bt, maxM, maxN, totz, nprod, st);
        case 32: return launch_tc_bn<32>(bt, maxM, maxN, totz, nprod, st);
        case 64: return launch_tc_bn<64>(bt, maxM, maxN, totz, nprod, st);
        default: return launch_tc_bn<128>(bt, maxM, maxN, totz, nprod, st);
    }
}

}  // namespace nasrec_gemm
