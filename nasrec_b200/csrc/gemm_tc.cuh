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
#include <cstdlib>
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

// round to nearest-even bfloat16, kept in an fp32 container (every bf16 value is a tf32 value: the tf32 tensor pipe then
// multiplies exactly what a kind::f16 bf16 MMA would)
__device__ __forceinline__ float round_bf16(float a) {
    const uint32_t u = __float_as_uint(a);
    return __uint_as_float((u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u);
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

__device__ __forceinline__ void store_split4(uint8_t* s_hi, uint8_t* s_lo, uint32_t off, const float4& x, bool split,
                                             bool bf16 = false) {
    float4 h, l;
    if (bf16) {
        h = make_float4(round_bf16(x.x), round_bf16(x.y), round_bf16(x.z), round_bf16(x.w));
    } else if (split) {
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

// One [ROWS x 32] fp32 k-tile travels global -> registers (Stager::load) -> (hi, lo) -> swizzled shared memory
// (Stager::store), moved by the 256 producer threads; the two halves are separate so that the loads of k-tile
// t+1 are in flight while k-tile t is split and stored (register double buffering).
//
// Everything that does not depend on the k-tile is computed once per term: each thread owns NV 16-byte chunks
// (row, 4 consecutive k) of the tile, keeps one global pointer per chunk (rows beyond the operand's extent are
// clamped onto its last row: they only feed output rows/columns that are never stored) and the chunk's swizzled
// shared-memory offset; a k-tile step is `ptr += step`.  Full tiles load without predicates -- one LDG.128 per
// chunk when the view is 16-byte aligned and k-contiguous, else four LDG.32 `se` floats apart (k-contiguous but
// unaligned rows such as the reference's 1037-float weight rows, or M/N-contiguous operands of dgrad/wgrad) --
// and only the last, partial tile of a term takes the predicated path.  Two-level views (sparse-axis tensors) are
// linear in k at tile granularity because their inner extent (16) divides the tile (32).
template <int ROWS>
struct Stager {
    static constexpr int NV = (ROWS * 8 + 255) / 256;
    const float* ptr[NV];
    uint32_t soff[NV];
    long long step, se;
    int kc0, dkc;          // chunk q covers k = 4 * (kc0 + q * dkc) .. +3 of the tile
    int K, k0;             // term length; first k of the next tile to load
    unsigned live;         // bit q: this thread's chunk q exists
    bool vec;

    __device__ __forceinline__ void init(const View& v, int i0, int I, int K_, int k_start, int tid) {
        int row0, drow;
        if (v.contig_j) { row0 = tid >> 3; kc0 = tid & 7; drow = 32; dkc = 0; }
        else { row0 = tid % ROWS; kc0 = tid / ROWS; drow = 0; dkc = 256 / ROWS; }
        step = voff(v, 0, TC_BK) - voff(v, 0, 0);
        se = voff(v, 0, 1) - voff(v, 0, 0);
        vec = v.vec16 != 0;
        K = K_;
        k0 = k_start;
        live = 0;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int row = row0 + q * drow, kc = kc0 + q * dkc;
            const bool ok = row < ROWS && kc < 8;
            if (ok) live |= 1u << q;
            const int r = ok ? row : 0, c = ok ? kc : 0;
            int i = i0 + r;
            i = i < I ? i : I - 1;
            ptr[q] = v.p + voff(v, i, c * 4) + (long long)(k_start / TC_BK) * step;
            soff[q] = sw128_off(r, c);
        }
    }
    __device__ __forceinline__ bool exhausted() const { return k0 >= K; }
    __device__ __forceinline__ void load(float4 (&r)[NV]) {
        if (k0 + TC_BK <= K) {
            if (vec) {
#pragma unroll
                for (int q = 0; q < NV; ++q) r[q] = __ldg(reinterpret_cast<const float4*>(ptr[q]));
            } else {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    r[q].x = __ldg(ptr[q]);
                    r[q].y = __ldg(ptr[q] + se);
                    r[q].z = __ldg(ptr[q] + 2 * se);
                    r[q].w = __ldg(ptr[q] + 3 * se);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int k = k0 + 4 * (kc0 + q * dkc);
                r[q].x = ldg_pred(ptr[q], k < K);
                r[q].y = ldg_pred(ptr[q] + se, k + 1 < K);
                r[q].z = ldg_pred(ptr[q] + 2 * se, k + 2 < K);
                r[q].w = ldg_pred(ptr[q] + 3 * se, k + 3 < K);
            }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) ptr[q] += step;
        k0 += TC_BK;
    }
    __device__ __forceinline__ void store(const float4 (&r)[NV], uint8_t* s_hi, uint8_t* s_lo, bool split, bool bf16) const {
#pragma unroll
        for (int q = 0; q < NV; ++q)
            if (live & (1u << q)) store_split4(s_hi, s_lo, soff[q], r[q], split, bf16);
    }
};

template <int BN>
struct TcCfg {
    // BN <= 64: few stages and 256 TMEM columns, so that two CTAs -- of the same launch, or of the main-stream and
    // side-stream GEMMs that run concurrently -- share an SM and hide each other's prologue/epilogue.  Measured
    // (profiles/r01_native_gemm_full.md): isolated launches get slower (K=1037 fwd 35 -> 50 us with two stages), the
    // training pipelines get faster (fixed-model graph +8 %, KDD B=2048 +7 %) because the GEMMs of the two streams overlap.
    static constexpr int STAGES = BN >= 64 ? 3 : 2;      // 192 KB (BN=128, one CTA per SM) / 144 KB / 81 KB / 73 KB
    static constexpr int PREFETCH = 2;   // k-tiles of global loads kept in registers
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;      // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    // Round-robin accumulators: k-tile `it` accumulates into TMEM accumulator it % nacc and the
    // epilogue adds the accumulators in fp32 registers (round-to-nearest).  The tensor core's own
    // fp32 accumulation truncates, so its error grows linearly with the number of MMAs chained
    // into one accumulator; spreading K over up to 8 accumulators brings it back to FFMA level.
    static constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;
    static constexpr int TMEM_COLS = 256;
    // 128-wide tiles rotate over 2 accumulators only (the TMA kernel shares its 512 TMEM columns with the split A operand):
    // they are chosen only together with split-K that leaves <= 16 k-tiles per CTA (tc_pick_bn), which bounds the chain
    static constexpr int NACC_MAX = TMEM_COLS / ACC_STRIDE < 8 ? TMEM_COLS / ACC_STRIDE : 8;
    static constexpr int CTAS_PER_SM = BN == 128 ? 1 : 2;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<BN>::CTAS_PER_SM) gemm_tc_kernel(const __grid_constant__ Batch bt, int nprod) {
    pdl_trigger();
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
    pdl_wait();          // everything above (problem lookup, barrier init, TMEM allocation) overlapped the predecessor

    if (warp < 8) {
        // ------------------------------------------------------------ producers
        // flat walk over this split's k-tiles: find the term that holds k-tile kt_begin
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
        Stager<TC_BM> sa;
        Stager<BN> sb;
        // register ring of D k-tiles: tile t+D is requested right after tile t has been stored, so D-1 tiles
        // of global loads are in flight while one is split -- the k-loop is otherwise bound by load latency
        // (one CTA per SM, 8 producer warps).
        constexpr int D = Cfg::PREFETCH;
        float4 ra[D][Stager<TC_BM>::NV], rb[D][Stager<BN>::NV];
        auto begin_term = [&](int t, int kk) {
            const Term& tm = bt.term[pr.term0 + t];
            sa.init(tm.a, m0, pr.M, tm.K, kk * TC_BK, tid);
            sb.init(tm.b, n0, pr.N, tm.K, kk * TC_BK, tid);
        };
        auto load_next = [&](float4 (&qa)[Stager<TC_BM>::NV], float4 (&qb)[Stager<BN>::NV]) {
            while (sa.exhausted() && t_cur + 1 < pr.nterm) begin_term(++t_cur, 0);
            sa.load(qa);
            sb.load(qb);
        };
        if (ntiles > 0) begin_term(t_cur, kk_cur);
#pragma unroll
        for (int h = 0; h < D; ++h)
            if (h < ntiles) load_next(ra[h], rb[h]);
#pragma unroll 1
        for (int it = 0; it < ntiles; it += D) {
#pragma unroll
            for (int h = 0; h < D; ++h) {
                const int cur = it + h;
                if (cur < ntiles) {
                    const int s = cur % Cfg::STAGES;
                    const uint32_t ph = (uint32_t)(cur / Cfg::STAGES) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
                    sa.store(ra[h], st, st + Cfg::A_BYTES, nprod > 2, nprod == 2);
                    sb.store(rb[h], st + 2 * Cfg::A_BYTES, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, nprod > 2, nprod == 2);
                    fence_proxy_async_smem();          // every writer orders its generic-proxy stores
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * s);
                    if (cur + D < ntiles) load_next(ra[h], rb[h]);
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
                    if (nprod > 2) {
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
    nasrec_launch(gemm_tc_kernel<BN>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, bt, nprod);
    return (int)cudaGetLastError();
}

// CTAs a launch would use with N-tile width bn
inline long long tc_cta_count(const Prob* prob, int nprob, int bn) {
    long long c = 0;
    for (int i = 0; i < nprob; ++i)
        c += (long long)((prob[i].M + TC_BM - 1) / TC_BM) * ((prob[i].N + bn - 1) / bn) * prob[i].nsplit;
    return c;
}
inline long long tc_cta_count(const Batch& bt, int bn) { return tc_cta_count(bt.prob, bt.nprob, bn); }

// Tile width and split-K: a small cost model fitted to measurements of this kernel family on B200 (profiles/r02_gemm.md).
//   * a CTA's k-loop costs about the same per k-tile whatever the tile width (0.38-0.47 us: TMA issue, conversion of the
//     128 x 32 A tile and twelve MMA issues are per-tile costs; only the B bytes grow with the width), so wide tiles do
//     more work per unit of time -- but a launch of few wide tiles leaves SMs idle and its CTAs walk the whole K range;
//   * split-K shortens that walk.  In the TMA kernel the CTAs of a thread-block cluster share an output tile and sum their
//     partial tiles through distributed shared memory (no workspace, no second launch, ~1 us); the LDG-producer kernel
//     uses the library workspace and a reduction launch with the SAME split count and summation order, so the two kernels
//     agree bit for bit;
//   * a launch costs whole waves of 148 CTAs -- of `sm_budget` CTAs when the caller knows that another stream's GEMMs run
//     concurrently (backward: the weight-gradient GEMMs on the side stream and the dY -> dX chain on the main one each plan
//     for half of the SMs, so that they overlap instead of queueing behind each other's full-GPU launches).
// Constraint from numerics: the tensor core truncates when it accumulates, so the chain of MMAs into one accumulator is
// bounded -- 128-wide tiles have two accumulators and are used only when a CTA walks <= 16 k-tiles.
// NASREC_TC_BN / NASREC_TC_NS force the choice (experiments); NASREC_TILE_POLICY=1 restores the round-1 narrow-tile rule.
constexpr int TC_SM_COUNT = 148;
struct TilePlan {
    int bn, ns;
};
inline int tc_split_for(long long ctas, int ktiles) {
    if (ctas <= 0 || ctas > TC_SM_COUNT / 2) return 1;
    int ns = (int)(TC_SM_COUNT / ctas);
    if (ns > ktiles / 3) ns = ktiles / 3;
    if (ns >= 8) return 8;
    if (ns >= 4) return 4;
    return ns < 2 ? 1 : 2;
}
// kind: 0 forward-like (both operands K-major), 1 dgrad-like (weight planes MN-major: bn / 32 boxes per plane and k-tile),
// 2 wgrad-like (both operands MN-major, the B tile split in shared memory by the converter warps)
template <class KTiles>
inline TilePlan tc_plan(const Prob* prob, int nprob, int maxN, KTiles ktiles_of, bool can_split, int kind, int sm_budget = TC_SM_COUNT) {
    static const int forced_bn = getenv("NASREC_TC_BN") ? atoi(getenv("NASREC_TC_BN")) : 0;
    static const int forced_ns = getenv("NASREC_TC_NS") ? atoi(getenv("NASREC_TC_NS")) : 0;
    static const int policy = getenv("NASREC_TILE_POLICY") ? atoi(getenv("NASREC_TILE_POLICY")) : 3;
    int kt_max = 0, kt_min = 1 << 30;
    for (int p = 0; p < nprob; ++p) {
        const int kt = ktiles_of(p);
        kt_max = kt > kt_max ? kt : kt_max;
        kt_min = kt < kt_min ? kt : kt_min;
    }
    if (policy == 1 && !forced_bn) {
        int bn;
        if (maxN <= 16) bn = 16;
        else if (maxN > 32 && tc_cta_count(prob, nprob, 64) >= 120) bn = 64;
        else bn = maxN <= 32 ? 32 : (tc_cta_count(prob, nprob, 32) > 296 ? 64 : 32);
        return TilePlan{bn, can_split ? tc_split_for(tc_cta_count(prob, nprob, bn), kt_min) : 1};
    }
    TilePlan best{maxN <= 16 ? 16 : 32, 1};
    double best_t = 1e30;
    static const int widths[4] = {16, 32, 64, 128};
    // measured us per k-tile of one CTA (gemm_prof2.py, B200, 3xTF32), by operand kind and tile width
    static const double t_tile[3][4] = {{0.40, 0.42, 0.52, 0.78}, {0.40, 0.42, 0.60, 0.97}, {0.42, 0.42, 0.73, 1.20}};
    // SMs a grid of clusters of 1 / 2 / 4 / 8 one-CTA-per-SM CTAs fills in one wave.  A cluster lives inside one GPC (16-20
    // SMs, not all multiples of the cluster size); measured: 144 CTAs in clusters of 4 take two waves (dgrad M = 512,
    // N = 13 + 1024, K = 1024: 23.9 us at (128, 4) against 15.1 us at (64, 2), profiles/r02_gemm.md)
    static const int capacity[4] = {148, 144, 128, 112};
    const double* tt = t_tile[kind < 0 || kind > 2 ? 2 : kind];
    for (int w = 0; w < 4; ++w) {
        const int bn = widths[w];
        if (forced_bn ? bn != forced_bn : ((bn == 16) != (maxN <= 16) || (bn > 32 && bn >= 2 * maxN))) continue;
        const long long tiles = tc_cta_count(prob, nprob, bn);
        for (int ns = 1, li = 0; ns <= 8; ns *= 2, ++li) {
            if (ns > 1 && (!can_split || kt_min < 2 * ns)) break;
            if (forced_ns && can_split && kt_min >= 2 * forced_ns && ns != forced_ns) continue;
            const int walk = (kt_max + ns - 1) / ns;
            if (bn == 128 && walk > 16 && !forced_bn) continue;
            const double cta = 3.5 + walk * tt[w] + (ns > 1 ? 0.3 + 0.1 * (bn / 32) : 0.0);
            const long long ctas = tiles * ns;
            int cap = capacity[li] < sm_budget ? capacity[li] : sm_budget / ns * ns;
            if (cap < ns) cap = ns;
            const double t = 2.5 + (double)((ctas + cap - 1) / cap) * cta + 0.002 * (double)ctas / TC_SM_COUNT;
            if (t < best_t) {
                best_t = t;
                best = TilePlan{bn, ns};
            }
        }
    }
    return best;
}

inline int launch_tc(const Batch& bt, int bn, int maxM, int maxN, int totz, int nprod, cudaStream_t st) {
    switch (bn) {
        case 16: return launch_tc_bn<16>(bt, maxM, maxN, totz, nprod, st);
        case 32: return launch_tc_bn<32>(bt, maxM, maxN, totz, nprod, st);
        case 64: return launch_tc_bn<64>(bt, maxM, maxN, totz, nprod, st);
        default: return launch_tc_bn<128>(bt, maxM, maxN, totz, nprod, st);
    }
}

}  // namespace nasrec_gemm
