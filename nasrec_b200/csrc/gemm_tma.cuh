// TMA-fed tcgen05 path of the segment-list GEMM (sm_100a): no operand passes through registers on its way
// from global memory to the tensor core.
//
//   warp 9 (one lane)  producer: cp.async.bulk.tensor (TMA) of the raw fp32 operand tiles into the swizzled
//                      shared-memory layouts the UMMA descriptors name; full[s] counts the landed bytes
//   warps 0-7          converters: fp32 -> (hi = rn_tf32(x), lo = x - hi) in place, shared memory only
//                      (weights arrive pre-split from their hi/lo planes and skip this step)
//   warp 8 (one lane)  tcgen05.mma.kind::tf32 issuer, accumulators in TMEM, tcgen05.commit frees the stage
//   warps 0-3          epilogue (same as gemm_tc.cuh: round-robin accumulators summed in fp32 registers)
//
// Arithmetic (split, product order, accumulator rotation, k-tile partition) is the one of gemm_tc.cuh, so the
// two kernels agree to the last bit; what changes is how the bytes travel.  Operand layouts:
//   KM128  K-major, SWIZZLE_128B: 2-D tensor (k, rows); activations / dY in forward and dgrad, weights forward
//   MN128  MN-major, 128-byte swizzle of 32-byte chunks (the only MN-major layout kind::tf32 accepts): 2-D tensor
//          (rows, k); weights in dgrad, both operands of wgrad
//   MN3R   same shared-memory layout for the rows m = (b, e) of a [B, rows, 16] sparse tensor, whose contiguous runs
//          are only 16 floats: TMA lands the plain 3-D box (e=16, k=r, b) in the lo plane and the converter warps
//          repack it (two samples' lanes per 128-byte row) while they split it
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

enum OpKind { OP_KM128 = 0, OP_MN128 = 1, OP_MN3R = 2, OP_KM64 = 3 };

struct OpLayout {
    int rank;               // tensor-map rank (2, 3 or 4)
    int nbox;               // TMA boxes per tile; box i: coordinate[box_dim] += 32 * i, shared offset i * box_bytes
    int box_dim;
    int box_bytes;
    int rdim, rsh;          // coordinate[rdim] += row0 >> rsh
    int kdim, ksh;          // coordinate[kdim] += k0 >> ksh
    uint32_t desc_hi32;     // constant upper word of the shared-memory descriptor (SBO, version, swizzle mode)
    uint32_t desc_lbo;      // LBO >> 4
    int koff[4];            // byte offset of each UMMA_K (8) step of a 32-wide k-tile
    int tile_bytes;         // bytes of one plane of the tile (= what the converters walk and TMA delivers)
    int mn_major;           // instruction-descriptor transpose bit
    int convert;            // 0: lo plane fetched by TMA; 1: lo plane derived in the kernel from the raw tile;
                            // 2: raw tile lands in the lo plane and is repacked into the MN128 layout while it is split
};

struct TTerm {
    int K;
    short a_hi, a_lo, b_hi, b_lo;     // tensor-map indices (lo unused when the operand is converted in the kernel)
    int a_base[4], b_base[4];         // coordinates of (row 0, k 0)
};

struct TBatch {
    alignas(64) CUtensorMap maps[TM_MAXMAPS];
    int nprob, nprod;
    OpLayout la, lb;
    Prob prob[MAXP];
    TTerm term[MAXT];
};

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

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_tile(const OpLayout& L, const CUtensorMap* map, uint32_t dst, uint32_t bar, const int* base,
                                         int row0, int k0) {
    int c[4] = {base[0], base[1], base[2], base[3]};
    c[L.rdim] += row0 >> L.rsh;
    c[L.kdim] += k0 >> L.ksh;
    for (int i = 0; i < L.nbox; ++i) {
        if (L.rank == 2) tma_load_2d(dst, map, bar, c[0], c[1]);
        else if (L.rank == 3) tma_load_3d(dst, map, bar, c[0], c[1], c[2]);
        else tma_load_4d(dst, map, bar, c[0], c[1], c[2], c[3]);
        dst += L.box_bytes;
        c[L.box_dim] += 32;
    }
}

__device__ __forceinline__ uint64_t tm_desc(const OpLayout& L, uint32_t saddr) {
    return (uint64_t)(((saddr >> 4) & 0x3FFFu) | (L.desc_lbo << 16)) | ((uint64_t)L.desc_hi32 << 32);
}

// hi = rn_tf32(x) written back in place, lo = x - hi into the lo plane at the same (swizzled) offset
template <int NV>
__device__ __forceinline__ void convert_tile(uint8_t* hi, uint8_t* lo, int nchunk, int tid) {
    float4 v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = tid + q * TM_CONV_THREADS;
        if (c < nchunk) v[q] = *reinterpret_cast<const float4*>(hi + (size_t)c * 16);
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = tid + q * TM_CONV_THREADS;
        if (c < nchunk) store_split4(hi, lo, (uint32_t)c * 16u, v[q], true);
    }
}

// MN3R: the landed box is [b (8)][r (32)][16 floats]; chunk c = 4 floats.  Destination: sample pair b>>1 is a 4 KB MN atom,
// k-row r a 128-byte row, (b&1, e>>3) its 32-byte chunk, XOR-swizzled with r&3 (SWIZZLE_128B_BASE32B).
__device__ __forceinline__ void repack_tile(uint8_t* hi, uint8_t* lo, int tid, bool split) {
    constexpr int NV = TC_BM * TC_BK * 4 / 16 / TM_CONV_THREADS;
    float4 v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = *reinterpret_cast<const float4*>(lo + (size_t)(tid + q * TM_CONV_THREADS) * 16);
    asm volatile("bar.sync 1, %0;" ::"n"(TM_CONV_THREADS) : "memory");       // everyone has read the landing zone
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = tid + q * TM_CONV_THREADS;
        const int e4 = c & 3, r = (c >> 2) & 31, b = c >> 7;
        const uint32_t dst = (uint32_t)((b >> 1) * 4096 + r * 128 + (((((b & 1) << 1) | (e4 >> 1)) ^ (r & 3)) << 5) + ((e4 & 1) << 4));
        store_split4(hi, lo, dst, v[q], split);
    }
}

template <int BN>
struct TmCfg {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;                         // 16 KB per plane
    static constexpr int B_BYTES = (BN < 32 ? 32 : BN) * TC_BK * 4;           // MN-major boxes are 32 wide
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = BN <= 32 ? 4 : (BN == 64 ? 4 : 3);          // 160 / 192 / 192 KB
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int ACC_STRIDE = TcCfg<BN>::ACC_STRIDE;
    static constexpr int TMEM_COLS = TcCfg<BN>::TMEM_COLS;
    static constexpr int NACC_MAX = TcCfg<BN>::NACC_MAX;
};

template <int BN>
__global__ void __launch_bounds__(TM_THREADS, 1) gemm_tma_kernel(const __grid_constant__ TBatch tb) {
    pdl_trigger();
    using Cfg = TmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem_base;

    int z = blockIdx.z, pi = 0;
    for (; pi < tb.nprob; ++pi) {
        const int ns = tb.prob[pi].nsplit;
        if (z < ns) break;
        z -= ns;
    }
    if (pi >= tb.nprob) return;
    const Prob& pr = tb.prob[pi];
    const int split = z;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    if (m0 >= pr.M || n0 >= pr.N) return;

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
    const int per = (tot + pr.nsplit - 1) / pr.nsplit;
    const int kt_begin = split * per;
    const int kt_end = min(tot, kt_begin + per);
    const int ntiles = max(0, kt_end - kt_begin);
    const int nacc = min(Cfg::NACC_MAX, ntiles);
    const int nprod = tb.nprod;
    const bool conv_a = (nprod > 1 && tb.la.convert) || tb.la.convert == 2, conv_b = nprod > 1 && tb.lb.convert;

    if (tid == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, TM_CONV_WARPS);
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
    pdl_wait();

    if (warp == 9) {
        // ------------------------------------------------------------ TMA producer (one thread)
        if (lane == 0 && ntiles > 0) {
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
            const bool a_lo = nprod > 1 && !tb.la.convert, b_lo = nprod > 1 && !tb.lb.convert;
            const uint32_t tx = (uint32_t)(tb.la.nbox * tb.la.box_bytes * (a_lo ? 2 : 1) + tb.lb.nbox * tb.lb.box_bytes * (b_lo ? 2 : 1));
            for (int it = 0; it < ntiles; ++it) {
                while (kk * TC_BK >= tb.term[pr.term0 + t_cur].K) { ++t_cur; kk = 0; }
                const TTerm& tm = tb.term[pr.term0 + t_cur];
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, tx);
                const uint32_t st = smem_u32(tiles + s * Cfg::STAGE_BYTES);
                const int k0 = kk * TC_BK;
                tma_tile(tb.la, &tb.maps[tm.a_hi], st + (tb.la.convert == 2 ? Cfg::A_BYTES : 0), full, tm.a_base, m0, k0);
                if (a_lo) tma_tile(tb.la, &tb.maps[tm.a_lo], st + Cfg::A_BYTES, full, tm.a_base, m0, k0);
                tma_tile(tb.lb, &tb.maps[tm.b_hi], st + 2 * Cfg::A_BYTES, full, tm.b_base, n0, k0);
                if (b_lo) tma_tile(tb.lb, &tb.maps[tm.b_lo], st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, full, tm.b_base, n0, k0);
                ++kk;
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0 && ntiles > 0) {
            const uint32_t idesc = umma_idesc_tf32(BN) | ((uint32_t)tb.la.mn_major << 15) | ((uint32_t)tb.lb.mn_major << 16);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                if (conv_a || conv_b) mbar_wait(bar_conv + 8 * s, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + s * Cfg::STAGE_BYTES);
                const uint32_t sb = sa + 2 * Cfg::A_BYTES;
                const uint32_t tacc = tmem + (uint32_t)((it % nacc) * Cfg::ACC_STRIDE);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UK; ++k) {
                    const uint64_t a_hi = tm_desc(tb.la, sa + tb.la.koff[k]);
                    const uint64_t a_lo = tm_desc(tb.la, sa + Cfg::A_BYTES + tb.la.koff[k]);
                    const uint64_t b_hi = tm_desc(tb.lb, sb + tb.lb.koff[k]);
                    const uint64_t b_lo = tm_desc(tb.lb, sb + Cfg::B_BYTES + tb.lb.koff[k]);
                    const uint32_t first = (it >= nacc || k > 0) ? 1u : 0u;
                    if (nprod > 1) {
                        umma_tf32(tacc, a_lo, b_hi, idesc, first);
                        umma_tf32(tacc, a_hi, b_lo, idesc, 1u);
                        if (nprod > 3) umma_tf32(tacc, a_lo, b_lo, idesc, 1u);
                        umma_tf32(tacc, a_hi, b_hi, idesc, 1u);
                    } else {
                        umma_tf32(tacc, a_hi, b_hi, idesc, first);
                    }
                }
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_done);
        }
        __syncwarp();
        tc_fence_before();
    } else {
        // ------------------------------------------------------------ converters (warps 0-7)
        if (conv_a || conv_b) {
            const int na = tb.la.tile_bytes >> 4, nb = tb.lb.tile_bytes >> 4;
#pragma unroll 1
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
                if (tb.la.convert == 2) repack_tile(st, st + Cfg::A_BYTES, tid, nprod > 1);
                else if (conv_a) convert_tile<Cfg::A_BYTES / 16 / TM_CONV_THREADS>(st, st + Cfg::A_BYTES, na, tid);
                if (conv_b)
                    convert_tile<(Cfg::B_BYTES / 16 + TM_CONV_THREADS - 1) / TM_CONV_THREADS>(st + 2 * Cfg::A_BYTES,
                                                                                             st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, nb, tid);
                fence_proxy_async_smem();
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
                if (transpose) {
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
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
    }
}

template <int BN>
inline int launch_tma_bn(const TBatch& tb, int maxM, int maxN, int totz, cudaStream_t st) {
    using Cfg = TmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((maxN + BN - 1) / BN, (maxM + TC_BM - 1) / TC_BM, totz);
    if (grid.y > 65535 || grid.z > 65535) return NASREC_ETOOBIG;
    nasrec_launch(gemm_tma_kernel<BN>, grid, TM_THREADS, Cfg::SMEM_BYTES, st, tb);
    return (int)cudaGetLastError();
}

}  // namespace nasrec_gemm
