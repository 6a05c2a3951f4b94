// TMA-fed tcgen05 path of the segment-list GEMM (sm_100a).
//
//   warp 9 (3 lanes)   producers: cp.async.bulk.tensor (TMA) of the raw fp32 A tile and of the B tile(s) into shared
//                      memory, one lane per stream (A, B hi, B lo); full[s] counts the landed bytes
//   warps 0-7          converters: A tile shared memory -> registers -> (hi = rn_tf32(x), lo = x - hi) -> TENSOR MEMORY
//                      (tcgen05.st); a B tile that is an activation (weight gradients) is split in place in shared memory
//   warp 8 (one lane)  tcgen05.mma.kind::tf32 issuer: A operand from TMEM, B operand from shared memory
//                      (descriptor), accumulators in TMEM, tcgen05.commit frees the stage
//   warps 0-3          epilogue (as gemm_tc.cuh: round-robin accumulators summed in fp32 registers)
//
// Why A lives in TMEM: the 3xTF32 scheme issues three MMAs per k-step that re-read both operands.  With both in shared
// memory a 128 x 32 x 32 k-tile costs ~60 KB of operand reads + 48 KB of converter traffic + 24 KB of TMA writes against
// the SM's 128 B/clk shared-memory port: ~1000 clk per k-tile where the tensor core needs ~200 (ncu: profiles/r02_gemm.md).
// Writing the split A planes to TMEM removes the converter stores and every A read from that port (B is the small
// operand: BN x 32).
//
// Arithmetic (split, product order, accumulator rotation, k-tile partition) is the one of gemm_tc.cuh, so the two
// kernels agree to the last bit; what changes is how the bytes travel.  Landing layouts of the raw tiles:
//   KM128  K-major, SWIZZLE_128B: 2-D tensor (k, rows); activations / dY in forward and dgrad (A), weight planes (B)
//   MN128  MN-major, 128-byte swizzle with 32-byte atoms (the only MN-major layout kind::tf32 accepts from shared
//          memory): 2-D tensor (rows, k); weight planes in dgrad (B), both operands of wgrad
//   MN3    plain 3-D box (e=16, k=r, b) of a [B, rows, 16] sparse tensor whose rows are m = (b, e) (A only)
//   KM64   K-major, SWIZZLE_64B: 3-D tensor (e=16, rows, b): sparse-axis weight gradient (k = (b, e))
// TMA zero-fills everything outside a tensor's extent, which is what the partial last k-tile of a segment needs.
#pragma once
#include <cuda.h>
#include "gemm_tc.cuh"

namespace nasrec_gemm {

constexpr int TM_MAXMAPS = 20;
constexpr int TM_CONV_WARPS = 8;
constexpr int TM_CONV_THREADS = TM_CONV_WARPS * 32;
constexpr int TM_THREADS = TM_CONV_THREADS + 64;       // + MMA warp + TMA warp
constexpr int TM_SLOTS = 4;                            // TMEM slots of the split A tile (64 columns each)
constexpr int TM_ACC_COLS = 256;                       // TMEM columns [0, 256): accumulators; [256, 512): A slots (hi 32 + lo 32 each)

enum OpKind { OP_KM128 = 0, OP_MN128 = 1, OP_MN3 = 2, OP_KM64 = 3 };

struct OpLayout {
    int kind;               // OpKind (selects the converters' read pattern for A)
    int rank;               // tensor-map rank
    int nbox;               // TMA boxes per tile; box i: coordinate[box_dim] += 32 * i, shared offset i * box_bytes
    int box_dim;
    int box_bytes;
    int rsh[4];             // coordinate[d] = base[d] + (row0 >> rsh[d]) + (k0 >> ksh[d]); shift 31 = no contribution
    int ksh[4];
    uint32_t desc_hi32;     // B only: constant upper word of the shared-memory descriptor (SBO, version, swizzle mode)
    uint32_t desc_lbo;      // B only: LBO >> 4
    int koff[4];            // B only: byte offset of each UMMA_K (8) step of a 32-wide k-tile
    int tile_bytes;         // bytes of one plane of the tile
    int mn_major;           // B only: instruction-descriptor transpose bit
    int convert;            // B only: 1 = lo plane derived in the kernel from the raw tile; 0 = lo plane fetched by TMA
};

struct TTerm {
    int K;
    short a_hi, a_lo, b_hi, b_lo;     // tensor-map indices (a_lo unused: A is always split in the kernel)
    int a_base[4], b_base[4];         // coordinates of (row 0, k 0)
};

// P problems, T terms, MAPS tensor maps.  The everyday launches (one operator direction: <= 16 problems) use the small
// instance; the batched weight-gradient launch (gemm.cu, wgrad_flush: every deferred weight gradient of a backward pass in
// one grid) uses the big one -- kernel parameters are copied per launch, so the common case stays at ~5 KB.
template <int P, int T, int MAPS>
struct TBatchT {
    alignas(64) CUtensorMap maps[MAPS];
    int nprob, nprod;
    int cluster_ns;         // > 1: split-K over a thread-block cluster of this many CTAs along z (DSMEM reduction); else 0/1
    int dbg;                // experiment knob (NASREC_GEMM_DBG): 1 no TMA loads, 2 converters idle, 4 no MMAs, 8 no TMEM-slot wait
    int flat;               // 1: blockIdx.x counts output tiles over ALL problems (tile_end = running totals), so that problems of
                            // very different sizes share a grid without empty CTAs; 0: grid (N tiles, M tiles, problem x split)
    OpLayout la, lb;
    Prob prob[P];
    TTerm term[T];
    int tile_end[P];
};
using TBatch = TBatchT<MAXP, MAXT, TM_MAXMAPS>;
constexpr int BIG_P = 64, BIG_MAPS = 96;
using TBatchBig = TBatchT<BIG_P, BIG_P, BIG_MAPS>;

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// Warp-uniform issue: the WHOLE warp executes these on warp-uniform operands and one lane is elected inside the asm block.
// Measured (tools/mma_bench.cu): a tcgen05.mma issued from an `if (lane == 0)` branch costs ~130 clk whatever its size --
// the compiler cannot prove the operands uniform and wraps every instruction of the uniform datapath in a
// lane-uniformisation loop -- against 26 clk (N = 32) / 65 clk (N = 128) when issued this way.
__device__ __forceinline__ void mbar_expect_tx_w(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_w(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_w(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint32_t bar) {
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar) : "memory");
}


// One TMA box.  Every tensor map of this kernel is encoded with rank 3 (2-D operands get a unit third dimension, gemm.cu
// get_map), so that a single instruction form serves all operand layouts and the producer loop has no rank branches.
__device__ __forceinline__ void tma_load_w(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// One k-tile of the 3xTF32 scheme in ONE asm block: 4 k-steps x (lo*hi, hi*lo, hi*hi) and the commit that frees the
// stage.  Issuing the twelve MMAs separately costs ~20 SASS instructions each (election, moves into uniform registers,
// votes) on a single warp -- ~1200 clk per k-tile, the bottleneck of the k-loop (profiles/r02_gemm.md); here the operands
// enter the uniform datapath once per tile.  bh / bl: low words of the B hi / lo descriptors at k-step 0; k1..k3: the
// (byte offset >> 4) of k-steps 1..3; ta: TMEM column of the A hi plane (lo plane 32 columns further).
__device__ __forceinline__ void umma_tile3_w(uint32_t tacc, uint32_t ta, uint32_t bh, uint32_t bl, uint32_t dhi, uint32_t idesc,
                                            uint32_t first, uint32_t k1, uint32_t k2, uint32_t k3, uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pf, pt;\n\t"
        ".reg .b32 al0, ah1, al1, ah2, al2, ah3, al3, x;\n\t"
        ".reg .b64 h0, h1, h2, h3, l0, l1, l2, l3;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pf, %6, 0;\n\t"
        "setp.eq.b32 pt, %6, %6;\n\t"
        "mov.b64 h0, {%2, %4};\n\t"
        "mov.b64 l0, {%3, %4};\n\t"
        "add.u32 x, %2, %7;\n\tmov.b64 h1, {x, %4};\n\t"
        "add.u32 x, %3, %7;\n\tmov.b64 l1, {x, %4};\n\t"
        "add.u32 x, %2, %8;\n\tmov.b64 h2, {x, %4};\n\t"
        "add.u32 x, %3, %8;\n\tmov.b64 l2, {x, %4};\n\t"
        "add.u32 x, %2, %9;\n\tmov.b64 h3, {x, %4};\n\t"
        "add.u32 x, %3, %9;\n\tmov.b64 l3, {x, %4};\n\t"
        "add.u32 al0, %1, 32;\n\t"
        "add.u32 ah1, %1, 8;\n\tadd.u32 al1, %1, 40;\n\t"
        "add.u32 ah2, %1, 16;\n\tadd.u32 al2, %1, 48;\n\t"
        "add.u32 ah3, %1, 24;\n\tadd.u32 al3, %1, 56;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al0], h0, %5, pf;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], l0, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], h0, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al1], h1, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], l1, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], h1, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al2], h2, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], l2, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], h2, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al3], h3, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], l3, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], h3, %5, pt;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
        "}\n" ::"r"(tacc),
        "r"(ta), "r"(bh), "r"(bl), "r"(dhi), "r"(idesc), "r"(first), "r"(k1), "r"(k2), "r"(k3), "r"(bar)
        : "memory");
}

__device__ __forceinline__ uint64_t tm_desc(const OpLayout& L, uint32_t saddr) {
    return (uint64_t)(((saddr >> 4) & 0x3FFFu) | (L.desc_lbo << 16)) | ((uint64_t)L.desc_hi32 << 32);
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// hi = rn_tf32(x) written back in place, lo = x - hi into the lo plane at the same (swizzled) offset (B tiles of wgrad)
template <int NV, int NT>
__device__ __forceinline__ void convert_tile(uint8_t* hi, uint8_t* lo, int nchunk, int tid, bool bf16) {
    float4 v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = tid + q * NT;
        if (c < nchunk) v[q] = *reinterpret_cast<const float4*>(hi + (size_t)c * 16);
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = tid + q * NT;
        if (c < nchunk) store_split4(hi, lo, (uint32_t)c * 16u, v[q], !bf16, bf16);
    }
}

template <int BN>
struct TmCfg {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;                         // 16 KB raw landing zone
    static constexpr int B_BYTES = (BN < 32 ? 32 : BN) * TC_BK * 4;           // MN-major boxes are 32 wide
    static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
    static constexpr int PW = BN < 32 ? 32 : BN;                               // columns of a partial tile parked in shared memory (cluster split-K)
    static constexpr int STAGES = BN <= 32 ? 8 : (BN == 64 ? 6 : 4);          // 192 KB each: deep TMA prefetch, one CTA per SM
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int ACC_STRIDE = TcCfg<BN>::ACC_STRIDE;
    static constexpr int NACC_MAX = TcCfg<BN>::NACC_MAX;
    static_assert(BN <= 128 && NACC_MAX * ACC_STRIDE <= TM_ACC_COLS, "accumulators and A slots share the 512 TMEM columns");
};

template <int BN, class TB>
__global__ void __launch_bounds__(TM_THREADS, 1) gemm_tma_kernel(const __grid_constant__ TB tb) {
    pdl_trigger();
    using Cfg = TmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem_base;

    // Split-K comes in two forms.  cluster_ns > 1: the launch is a grid of clusters (1, 1, cluster_ns); the CTAs of a cluster
    // share an output tile, each takes a slice of the k-tiles, and the partial tiles are summed through distributed shared
    // memory in split order (no workspace, no reduction launch).  Otherwise a problem may carry its own nsplit with
    // partials going to the caller's workspace (sparse-axis projections, which reduce over the batch).
    const int cns = tb.cluster_ns > 1 ? tb.cluster_ns : 1;
    int z = blockIdx.z, pi = 0, tile_m = blockIdx.y, tile_n = blockIdx.x;
    if (tb.flat) {
        const int t = blockIdx.x;                   // z is the split index (clusters along z)
        while (pi < tb.nprob && t >= tb.tile_end[pi]) ++pi;
        if (pi >= tb.nprob) return;
        const int local = t - (pi ? tb.tile_end[pi - 1] : 0);
        const int ntn = (tb.prob[pi].N + BN - 1) / BN;
        tile_m = local / ntn;
        tile_n = local - tile_m * ntn;
    } else if (cns > 1) {
        pi = z / cns;
        z -= pi * cns;
    } else {
        for (; pi < tb.nprob; ++pi) {
            const int ns = tb.prob[pi].nsplit;
            if (z < ns) break;
            z -= ns;
        }
    }
    if (pi >= tb.nprob) return;
    const Prob& pr = tb.prob[pi];
    const int split = z;
    const int nsplit = cns > 1 ? cns : pr.nsplit;
    const int m0 = tile_m * TC_BM, n0 = tile_n * BN;
    if (m0 >= pr.M || n0 >= pr.N) return;          // the whole cluster leaves together (same tile, same problem)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    uint8_t* tiles = smem_raw + pad;
    uint8_t* bars = tiles + Cfg::STAGES * Cfg::STAGE_BYTES;
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_conv = bar_full + 8 * Cfg::STAGES;
    const uint32_t bar_empty = bar_conv + 8 * Cfg::STAGES;
    const uint32_t bar_done = bar_empty + 8 * Cfg::STAGES;

    int tot = 0;
    for (int t = 0; t < pr.nterm; ++t) tot += (tb.term[pr.term0 + t].K + TC_BK - 1) / TC_BK;
    const int per = (tot + nsplit - 1) / nsplit;
    const int kt_begin = split * per;
    const int kt_end = min(tot, kt_begin + per);
    const int ntiles = max(0, kt_end - kt_begin);
    const int nacc = min(Cfg::NACC_MAX, ntiles);
    const int nprod = tb.nprod;
    // nprod: 1 = plain tf32 single pass, 2 = bf16-rounded operands single pass, 3/4 = tf32 hi/lo split products
    const bool conv_b = nprod > 1 && tb.lb.convert;

    if (tid < 32) {          // one barrier per lane: the 25 initialisations go out together instead of one after the other
        if (tid < Cfg::STAGES) mbar_init(bar_full + 8 * tid, 1);
        else if (tid < 2 * Cfg::STAGES) mbar_init(bar_conv + 8 * (tid - Cfg::STAGES), TM_CONV_WARPS / 2);
        else if (tid < 3 * Cfg::STAGES) mbar_init(bar_empty + 8 * (tid - 2 * Cfg::STAGES), 1);
        else if (tid == 3 * Cfg::STAGES) mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    pdl_wait();

    if (warp == 9) {
        // ------------------------------------------------------------ TMA producer: the whole warp runs warp-uniform code,
        // one elected lane issues; per k-tile: A (raw), B (hi / raw) and, for weights, B lo
        if (ntiles > 0) {
            const OpLayout& LA = tb.la;
            const OpLayout& LB = tb.lb;
            const bool b_lo = nprod > 2 && !LB.convert;                        // the weight's lo plane is fetched, not derived
            const uint32_t tx = (uint32_t)(LA.nbox * LA.box_bytes + LB.nbox * LB.box_bytes * (b_lo ? 2 : 1));
            const int a_nbox = LA.nbox, a_bb = LA.box_bytes, b_nbox = LB.nbox, b_bb = LB.box_bytes;
            const int ar0 = m0 >> LA.rsh[0], ar1 = m0 >> LA.rsh[1], ar2 = m0 >> LA.rsh[2];
            const int br0 = n0 >> LB.rsh[0], br1 = n0 >> LB.rsh[1], br2 = n0 >> LB.rsh[2];
            const int ak0 = LA.ksh[0], ak1 = LA.ksh[1], ak2 = LA.ksh[2], bk0 = LB.ksh[0], bk1 = LB.ksh[1], bk2 = LB.ksh[2];
            const int ad0 = LA.box_dim == 0 ? 32 : 0, ad1 = LA.box_dim == 1 ? 32 : 0;
            const int bd0 = LB.box_dim == 0 ? 32 : 0, bd1 = LB.box_dim == 1 ? 32 : 0;
            int t_cur = 0, kk = 0;
            {
                int kt = 0;
                for (; t_cur < pr.nterm; ++t_cur) {
                    const int nk = (tb.term[pr.term0 + t_cur].K + TC_BK - 1) / TC_BK;
                    if (kt + nk > kt_begin) {
                        kk = kt_begin - kt;
                        break;
                    }
                    kt += nk;
                }
            }
            // Running state: the coordinates of the current k-tile advance by constants (a k-tile is 32 k: +32, +2 (k >> 4)
            // or +0 per dimension), so the loop body is wait / expect / issue with no address arithmetic beyond adds.
            const int ai0 = ak0 == 31 ? 0 : (TC_BK >> ak0), ai1 = ak1 == 31 ? 0 : (TC_BK >> ak1), ai2 = ak2 == 31 ? 0 : (TC_BK >> ak2);
            const int bi0 = bk0 == 31 ? 0 : (TC_BK >> bk0), bi1 = bk1 == 31 ? 0 : (TC_BK >> bk1), bi2 = bk2 == 31 ? 0 : (TC_BK >> bk2);
            int left = 0, a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
            const CUtensorMap *map_a = nullptr, *map_bh = nullptr, *map_bl = nullptr;
            auto load_term = [&](int t, int k_first) {
                const TTerm& tm = tb.term[pr.term0 + t];
                left = (tm.K + TC_BK - 1) / TC_BK - k_first;
                map_a = &tb.maps[tm.a_hi]; map_bh = &tb.maps[tm.b_hi]; map_bl = &tb.maps[tm.b_lo];
                a0 = tm.a_base[0] + ar0 + k_first * ai0; a1 = tm.a_base[1] + ar1 + k_first * ai1; a2 = tm.a_base[2] + ar2 + k_first * ai2;
                b0 = tm.b_base[0] + br0 + k_first * bi0; b1 = tm.b_base[1] + br1 + k_first * bi1; b2 = tm.b_base[2] + br2 + k_first * bi2;
            };
            if (lane == 0)
                for (int t = 0; t < pr.nterm; ++t) {       // warm the descriptor cache
                    const TTerm& tm = tb.term[pr.term0 + t];
                    asm volatile("prefetch.tensormap [%0];" ::"l"(&tb.maps[tm.a_hi]) : "memory");
                    asm volatile("prefetch.tensormap [%0];" ::"l"(&tb.maps[tm.b_hi]) : "memory");
                    if (b_lo) asm volatile("prefetch.tensormap [%0];" ::"l"(&tb.maps[tm.b_lo]) : "memory");
                }
            __syncwarp();
            load_term(t_cur, kk);
            const uint32_t tiles_u32 = smem_u32(tiles);
            const bool no_tma = tb.dbg & 1;
            int s = 0;
            uint32_t ph = 1u;                              // parity the empty barrier of a fresh stage passes at once
#pragma unroll 1
            for (int it = 0; it < ntiles; ++it) {
                while (left == 0) load_term(++t_cur, 0);
                const uint32_t full = bar_full + 8 * s;
                mbar_wait(bar_empty + 8 * s, ph);
                mbar_expect_tx_w(full, no_tma ? 0u : tx);
                if (!no_tma) {
                    const uint32_t dst = tiles_u32 + (uint32_t)(s * Cfg::STAGE_BYTES);
                    for (int i = 0; i < a_nbox; ++i)
                        tma_load_w(dst + (uint32_t)(i * a_bb), map_a, full, a0 + i * ad0, a1 + i * ad1, a2);
                    for (int i = 0; i < b_nbox; ++i) {
                        const uint32_t d = dst + (uint32_t)(Cfg::A_BYTES + i * b_bb);
                        tma_load_w(d, map_bh, full, b0 + i * bd0, b1 + i * bd1, b2);
                        if (b_lo) tma_load_w(d + Cfg::B_BYTES, map_bl, full, b0 + i * bd0, b1 + i * bd1, b2);
                    }
                }
                a0 += ai0; a1 += ai1; a2 += ai2;
                b0 += bi0; b1 += bi1; b2 += bi2;
                --left;
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer: whole warp, warp-uniform operands, elected lane
        if (ntiles > 0) {
            const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t idesc = umma_idesc_tf32(BN) | ((uint32_t)tb.lb.mn_major << 16);       // A from TMEM is K-major
            const uint32_t dlo = tb.lb.desc_lbo << 16;
            const uint64_t dhi = (uint64_t)tb.lb.desc_hi32 << 32;
            const int ko0 = tb.lb.koff[0], ko1 = tb.lb.koff[1], ko2 = tb.lb.koff[2], ko3 = tb.lb.koff[3];
            const uint32_t tiles_u32 = smem_u32(tiles);
            const uint32_t dhi32 = tb.lb.desc_hi32;
            const uint32_t k1 = (uint32_t)ko1 >> 4, k2 = (uint32_t)ko2 >> 4, k3 = (uint32_t)ko3 >> 4;
            const bool no_mma = tb.dbg & 4;
            int s = 0, slot = 0, acc = 0;
            uint32_t ph = 0u, wrapped = 0u;
#pragma unroll 1
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait(bar_full + 8 * s, ph);
                mbar_wait(bar_conv + 8 * s, ph);
                tc_fence_after();
                const uint32_t sb = tiles_u32 + (uint32_t)(s * Cfg::STAGE_BYTES) + Cfg::A_BYTES;
                const uint32_t ta = tm_u + (uint32_t)(TM_ACC_COLS + slot * 64);
                const uint32_t tacc = tm_u + (uint32_t)(acc * Cfg::ACC_STRIDE);
                const uint32_t bar_e = bar_empty + 8 * s;
                if (nprod == 3 && !no_mma) {
                    // hot path: the whole k-tile and its commit in one asm block
                    umma_tile3_w(tacc, ta, ((sb >> 4) & 0x3FFFu) | dlo, (((sb + Cfg::B_BYTES) >> 4) & 0x3FFFu) | dlo, dhi32, idesc, wrapped, k1,
                                 k2, k3, bar_e);
                } else {
                    if (!no_mma)
#pragma unroll
                        for (int k = 0; k < TC_BK / TC_UK; ++k) {
                            const int ko = k == 0 ? ko0 : (k == 1 ? ko1 : (k == 2 ? ko2 : ko3));
                            const uint32_t a_hi = ta + (uint32_t)(k * TC_UK), a_lo = a_hi + 32u;
                            const uint64_t b_hi = (uint64_t)((((sb + ko) >> 4) & 0x3FFFu) | dlo) | dhi;
                            const uint64_t b_lo = (uint64_t)((((sb + Cfg::B_BYTES + ko) >> 4) & 0x3FFFu) | dlo) | dhi;
                            const uint32_t first = (wrapped || k > 0) ? 1u : 0u;
                            if (nprod > 2) {
                                umma_tf32_ts_w(tacc, a_lo, b_hi, idesc, first);
                                umma_tf32_ts_w(tacc, a_hi, b_lo, idesc, 1u);
                                if (nprod > 3) umma_tf32_ts_w(tacc, a_lo, b_lo, idesc, 1u);
                                umma_tf32_ts_w(tacc, a_hi, b_hi, idesc, 1u);
                            } else {
                                umma_tf32_ts_w(tacc, a_hi, b_hi, idesc, first);
                            }
                        }
                    umma_commit_w(bar_e);
                }
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                slot = (slot + 1) & (TM_SLOTS - 1);
                if (++acc == nacc) { acc = 0; wrapped = 1u; }
            }
            umma_commit_w(bar_done);
        }
        __syncwarp();
        tc_fence_before();
    } else {
        // ------------------------------------------------------------ converters (warps 0-7)
        // thread = one A row (TMEM lane 32 * (warp & 3) + lane); warps 0-3 convert the even k-tiles, warps 4-7 the odd ones,
        // so that two tiles' load -> split -> TMEM-store latency chains are in flight at a time
        {
            const int q = warp & 3, g = warp >> 2;
            const int m = q * 32 + lane;
            const int akind = tb.la.kind;
            const int nb = tb.lb.tile_bytes >> 4;
            const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)TM_ACC_COLS;
            const int ctid = tid & 127;
#pragma unroll 1
            for (int it = g; it < ntiles; it += 2) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                const uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
                if (it >= TM_SLOTS && !(tb.dbg & 8)) {
                    // the TMEM slot is free once the MMAs of k-tile it - TM_SLOTS have retired: their commit arrived on that
                    // tile's empty barrier (which cannot complete again before this tile has been converted)
                    const int jt = it - TM_SLOTS;
                    mbar_wait(bar_empty + 8 * (jt % Cfg::STAGES), (uint32_t)(jt / Cfg::STAGES) & 1u);
                    tc_fence_after();
                }
                const uint32_t ta = lane_base + (uint32_t)((it % TM_SLOTS) * 64);
                if (!(tb.dbg & 2))
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float x[16];
                    if (akind == OP_KM128) {
                        // [128 rows][128 B], 16-byte chunks XOR-swizzled with row & 7: 4 chunks of this thread's row
                        const uint8_t* rowp = st + m * 128;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(rowp + ((((h << 2) + c) ^ (m & 7)) << 4));
                            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                        }
                    } else if (akind == OP_MN128) {
                        // 4 boxes of [32 k rows][32 rows x 4 B], 32-byte chunks XOR-swizzled with k & 3
                        const uint8_t* boxp = st + (m >> 5) * 4096 + (m & 7) * 4;
                        const int ch = (m & 31) >> 3;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int k = (h << 4) + j;
                            x[j] = *reinterpret_cast<const float*>(boxp + k * 128 + ((ch ^ (k & 3)) << 5));
                        }
                    } else if (akind == OP_MN3) {
                        // plain [8 samples][32 k rows][16 floats]; row m = (sample m >> 4, lane m & 15)
                        const uint8_t* bp = st + (m >> 4) * 2048 + (m & 15) * 4;
#pragma unroll
                        for (int j = 0; j < 16; ++j) x[j] = *reinterpret_cast<const float*>(bp + ((h << 4) + j) * 64);
                    } else {
                        // OP_KM64: [2 samples][128 rows][64 B], 16-byte chunks XOR-swizzled with (row >> 1) & 3; k = (sample h, e)
                        const uint8_t* rowp = st + h * 8192 + m * 64;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ ((m >> 1) & 3)) << 4));
                            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                        }
                    }
                    if (nprod > 2) {
                        float hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) split_tf32(x[j], hi[j], lo[j]);
                        tmem_st16(ta + (uint32_t)(h * 16), hi);
                        tmem_st16(ta + 32u + (uint32_t)(h * 16), lo);
                    } else if (nprod == 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) x[j] = round_bf16(x[j]);
                        tmem_st16(ta + (uint32_t)(h * 16), x);
                    } else {
                        tmem_st16(ta + (uint32_t)(h * 16), x);
                    }
                }
                if (conv_b) {
                    uint8_t* bh = const_cast<uint8_t*>(st) + Cfg::A_BYTES;
                    convert_tile<(Cfg::B_BYTES / 16 + 127) / 128, 128>(bh, bh + Cfg::B_BYTES, nb, ctid, nprod == 2);
                    fence_proxy_async_smem();
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_conv + 8 * s);
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
                if (cns > 1) {
                    // partial tile -> this CTA's shared memory, rows padded by 4 floats (conflict-free float4 stores)
                    float4* prow = reinterpret_cast<float4*>(reinterpret_cast<float*>(tiles) + (warp * 32 + lane) * (Cfg::PW + 4) + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) prow[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                } else if (transpose) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = r[j];
                    __syncwarp();
                    const int n = n0 + c0 + lane;
                    const float bias = (pr.bias && n < pr.N) ? __ldg(pr.bias + n) : 0.f;
                    if (pr.c_sh_i == 0 && n < pr.N) {
                        // plain row-major output (every linear): one running offset, rows unrolled by 8 so that the addend
                        // loads and the stores of several rows are in flight together (this loop is ~1 us of every launch)
                        const int mbase = m0 + warp * 32;
                        const int rows = min(32, pr.M - mbase);
                        const long long so = (long long)split * pr.split_stride;
                        long long o = (long long)mbase * pr.c_hi_i + n;
                        int rr = 0;
                        for (; rr + 8 <= rows; rr += 8) {
                            float v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) v[u] = scratch[(rr + u) * 33 + lane] + bias;
                            if (pr.addend) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) v[u] += pr.addend[o + (long long)u * pr.c_hi_i];
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) pr.c[o + (long long)u * pr.c_hi_i + so] = v[u];
                            o += 8 * pr.c_hi_i;
                        }
                        for (; rr < rows; ++rr) {
                            float v = scratch[rr * 33 + lane] + bias;
                            if (pr.addend) v += pr.addend[o];
                            pr.c[o + so] = v;
                            o += pr.c_hi_i;
                        }
                    } else {
                        for (int rr = 0; rr < 32; ++rr) {
                            const int mm = m0 + warp * 32 + rr;
                            if (mm >= pr.M || n >= pr.N) continue;
                            const long long o = (long long)(mm >> pr.c_sh_i) * pr.c_hi_i + (long long)(mm & cmask) * pr.c_lo_i + n;
                            float v = scratch[rr * 33 + lane] + bias;
                            if (pr.addend) v += pr.addend[o];
                            pr.c[o + (long long)split * pr.split_stride] = v;
                        }
                    }
                    __syncwarp();
                } else if (row < pr.M) {
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
    }
    if (cns > 1) {
        // ------------------------------------------------------------ cluster split-K: fixed-order sum over the CTAs' partial
        // tiles (split 0, 1, ...: the order of the workspace reduction of the LDG-producer kernel, bit for bit), each CTA
        // finishing a slice of the rows; then bias, accumulate, store
        asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
        if (tid < TM_CONV_THREADS) {
            constexpr int C4 = Cfg::PW / 4;
            const int rows_per = TC_BM / cns;
            const int total = rows_per * C4;
            const uint32_t p_local = smem_u32(tiles);
            const bool vec = ((reinterpret_cast<uintptr_t>(pr.c) & 15) == 0) && ((pr.c_hi_i & 3) == 0) &&
                             (!pr.addend || (reinterpret_cast<uintptr_t>(pr.addend) & 15) == 0);
            for (int i = tid; i < total; i += TM_CONV_THREADS) {
                const int rr = split * rows_per + i / C4, c4 = i % C4;
                const int m = m0 + rr, n = n0 + c4 * 4;
                if (m >= pr.M || n >= pr.N) continue;
                const uint32_t off = p_local + (uint32_t)((rr * (Cfg::PW + 4) + c4 * 4) * 4);
                float4 part[8];
#pragma unroll
                for (int sp = 0; sp < 8; ++sp)
                    if (sp < cns) {
                        uint32_t ra;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(off), "r"(sp));
                        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(part[sp].x), "=f"(part[sp].y), "=f"(part[sp].z), "=f"(part[sp].w)
                                     : "r"(ra)
                                     : "memory");
                    }
                float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int sp = 0; sp < 8; ++sp)
                    if (sp < cns) {
                        v[0] += part[sp].x; v[1] += part[sp].y; v[2] += part[sp].z; v[3] += part[sp].w;
                    }
                const long long o = (long long)m * pr.c_hi_i + n;
                if (vec && n + 3 < pr.N) {
                    if (pr.bias) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] += __ldg(pr.bias + n + j);
                    }
                    if (pr.addend) {
                        const float4 a = *reinterpret_cast<const float4*>(pr.addend + o);
                        v[0] = a.x + v[0]; v[1] = a.y + v[1]; v[2] = a.z + v[2]; v[3] = a.w + v[3];
                    }
                    *reinterpret_cast<float4*>(pr.c + o) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (n + j >= pr.N) break;
                        float w = v[j];
                        if (pr.bias) w += __ldg(pr.bias + n + j);
                        if (pr.addend) w = pr.addend[o + j] + w;
                        pr.c[o + j] = w;
                    }
                }
            }
        }
        // nobody leaves (and frees its shared memory) while a neighbour may still be reading it
        asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int BN, class TB>
inline int launch_tma_bn(const TB& tb, dim3 grid, cudaStream_t st) {
    using Cfg = TmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<BN, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (grid.y > 65535 || grid.z > 65535) return NASREC_ETOOBIG;
    if (tb.cluster_ns > 1) {
        // clusters of 192 KB CTAs need the non-portable opt-in only above 8; ns <= 8 here
        nasrec_launch_cluster(gemm_tma_kernel<BN, TB>, grid, TM_THREADS, Cfg::SMEM_BYTES, st, tb.cluster_ns, tb);
    } else {
        nasrec_launch(gemm_tma_kernel<BN, TB>, grid, TM_THREADS, Cfg::SMEM_BYTES, st, tb);
    }
    return (int)cudaGetLastError();
}

}  // namespace nasrec_gemm
