// Embedding stem: fused bag-size-1 gather over all F tables, deterministic
// sorted-row gradient reduction and row-wise Adagrad.
//
// Replaces F x nn.Embedding + torch.stack (nasrec/supernet/supernet.py:404-430),
// F x embedding_dense_backward, and the dense torch.optim.Adagrad sweep over
// every table row (nasrec/train_supernet.py:121-123).
//
// HBM-bound byte work: rows are 64 B (16 fp32); 4 lanes move one row as float4,
// so one warp instruction moves 8 rows; indices are read coalesced along F.
#include <cub/block/block_radix_sort.cuh>
#include "common.cuh"

namespace {

// Four (pair, quarter-row) items per thread, a block apart, with every index load, then every row load, then every store
// issued together: the kernel is a chain of three dependent memory latencies (id -> row address -> row), and at the
// evaluation batch it is only ~30 MB, so what counts is how many of those chains a thread keeps in flight.
constexpr int GATHER_UNR = 4;
__global__ void __launch_bounds__(256) emb_gather_kernel(const float* const* __restrict__ tables,
                                                         const int64_t* __restrict__ num_rows,
                                                         const int64_t* __restrict__ idx,
                                                         float* __restrict__ out, long long n_pairs, int F,
                                                         int* err_flag) {
    pdl_enter();
    const long long t0 = blockIdx.x * (long long)(blockDim.x * GATHER_UNR) + threadIdx.x;
    long long pair[GATHER_UNR], row[GATHER_UNR];
    int f[GATHER_UNR];
    bool ok[GATHER_UNR];
#pragma unroll
    for (int u = 0; u < GATHER_UNR; ++u) {
        const long long t = t0 + (long long)u * blockDim.x;
        pair[u] = t >> 2;                   // (b,f) flattened, same order as idx and out
        ok[u] = pair[u] < n_pairs;
        row[u] = ok[u] ? idx[pair[u]] : 0;
        f[u] = (int)(pair[u] % F);
    }
    const float4* src[GATHER_UNR];
    bool bad = false;
#pragma unroll
    for (int u = 0; u < GATHER_UNR; ++u) {
        if (ok[u] && (unsigned long long)row[u] >= (unsigned long long)num_rows[f[u]]) {
            bad = true;
            row[u] = 0;
        }
        src[u] = reinterpret_cast<const float4*>(tables[ok[u] ? f[u] : 0] + row[u] * NASREC_EMB_DIM) + (int)((t0 + (long long)u * blockDim.x) & 3);
    }
    if (bad && err_flag) atomicOr(err_flag, 1);
    float4 v[GATHER_UNR];
#pragma unroll
    for (int u = 0; u < GATHER_UNR; ++u)
        if (ok[u]) v[u] = __ldg(src[u]);
#pragma unroll
    for (int u = 0; u < GATHER_UNR; ++u)
        if (ok[u]) __stcs(reinterpret_cast<float4*>(out + pair[u] * NASREC_EMB_DIM) + (int)((t0 + (long long)u * blockDim.x) & 3), v[u]);
}

// One CTA per table: bitonic sort of (row << 32 | sample) keys in shared memory,
// head flags + block scan -> unique rows, then 16 lanes per unique row sum the
// duplicates in ascending sample order.
// ITEMS == 0: bitonic sort of the keys in shared memory (any power-of-two Bpad, Bpad threads up to 1024).
// ITEMS >= 1: 1024 threads x ITEMS keys in registers, stable LSD radix sort on the row bits only (cub::BlockRadixSort; keys
//             enter in sample order, so equal rows stay in ascending sample order): O(n) per pass instead of the bitonic
//             network's O(n log^2 n) -- what the data-parallel global batch (N x 512 ids per table) needs.
template <int ITEMS>
__global__ void __launch_bounds__(1024) emb_sort_reduce_kernel(const int64_t* __restrict__ idx,
                                                               const int64_t* __restrict__ num_rows, int* err_flag,
                                                               const float* __restrict__ gout, int Bin, int Bpad,
                                                               int F, int64_t* __restrict__ uniq,
                                                               int* __restrict__ nuniq,
                                                               float* __restrict__ row_grad,
                                                               float* __restrict__ sumsq,
                                                               int* __restrict__ seg_scratch) {
    pdl_enter();
    extern __shared__ unsigned long long keys[];
    __shared__ int wsum[32];
    __shared__ float red[34];
    __shared__ int s_total, s_valid;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    // ids outside [0, num_rows[f]) would address rows that do not exist (the reference's nn.Embedding asserts): they are
    // dropped here -- their key sorts behind every valid one -- and reported through err_flag, as the forward gather does
    const unsigned long long nrows = num_rows ? (unsigned long long)num_rows[f] : (1ull << 31);
    int nvalid = 0;
    for (int i = tid; i < Bpad; i += nt) {
        unsigned long long key = ~0ull;
        if (i < Bin) {
            const unsigned long long row = (unsigned long long)idx[(long long)i * F + f];
            if (row < nrows) { key = (row << 32) | (unsigned)i; ++nvalid; }
        }
        keys[i] = key;
    }
    if (nvalid) atomicAdd(&s_valid, nvalid);
    __syncthreads();
    const int B = s_valid;                    // valid keys; the scratch / output strides stay Bin
    if (tid == 0 && B != Bin && err_flag) atomicOr(err_flag, 1);
    if constexpr (ITEMS > 0) {
        using Sort = cub::BlockRadixSort<unsigned long long, 1024, ITEMS>;
        unsigned long long mine[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) mine[j] = keys[tid * ITEMS + j];
        __syncthreads();                      // the key array doubles as cub's scratch from here on
        // row bits only: 2^bits > nrows, so the all-ones key of a dropped / padding entry sorts behind every valid row
        int bits = 1;
        while (bits < 31 && (1ull << bits) <= nrows) ++bits;
        Sort(*reinterpret_cast<typename Sort::TempStorage*>(keys)).Sort(mine, 32, 32 + bits);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) keys[tid * ITEMS + j] = mine[j];
        __syncthreads();
    } else
    for (int k = 2; k <= Bpad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Bpad; i += nt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    // contiguous chunk per thread; count heads, scan, then emit segment starts
    const int chunk = (Bpad + nt - 1) / nt;
    const int beg = tid * chunk, end = min(B, beg + chunk);
    int cnt = 0;
    for (int i = beg; i < end; ++i)
        cnt += (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0;
    // block exclusive scan of cnt
    const int lane = tid & 31, wid = tid >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (nt >> 5) ? wsum[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        wsum[lane] = winc - w;                 // exclusive
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    int pos = wsum[wid] + inc - cnt;
    int* seg = seg_scratch + (long long)f * (Bin + 1);
    for (int i = beg; i < end; ++i)
        if (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) seg[pos++] = i;
    const int U = s_total;
    if (tid == 0) {
        seg[U] = B;
        nuniq[f] = U;
    }
    __syncthreads();
    // 16 lanes per unique row
    const int e = tid & 15;
    float sq = 0.f;
    for (int u = tid >> 4; u < U; u += nt >> 4) {
        const int s0 = seg[u], s1 = seg[u + 1];
        // ascending sample order (fixed summation order); the loads of 8 duplicates are issued together so that a
        // hot row (hundreds of duplicates under Zipf ids) costs one memory latency per 8 samples, not per sample
        float acc = 0.f;
        int i = s0;
        for (; i + 8 <= s1; i += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned b = (unsigned)(keys[i + j] & 0xffffffffu);
                v[j] = __ldg(gout + ((long long)b * F + f) * NASREC_EMB_DIM + e);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j];
        }
        for (; i < s1; ++i) {
            const unsigned b = (unsigned)(keys[i] & 0xffffffffu);
            acc += __ldg(gout + ((long long)b * F + f) * NASREC_EMB_DIM + e);
        }
        row_grad[((long long)f * Bin + u) * NASREC_EMB_DIM + e] = acc;
        if (e == 0) uniq[(long long)f * Bin + u] = (int64_t)(keys[s0] >> 32);
        sq += acc * acc;
    }
    const float tot = block_sum(sq, red);
    if (tid == 0) sumsq[f] = tot;
}

__global__ void __launch_bounds__(256) emb_to_dense_kernel(const int64_t* __restrict__ uniq,
                                                           const int* __restrict__ nuniq,
                                                           const float* __restrict__ row_grad,
                                                           float* const* __restrict__ grad_tables, int B) {
    pdl_enter();
    const int f = blockIdx.y;
    const int U = nuniq[f];
    const int e = threadIdx.x & 15;
    for (int u = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); u < U; u += gridDim.x * (blockDim.x >> 4)) {
        const long long row = uniq[(long long)f * B + u];
        grad_tables[f][row * NASREC_EMB_DIM + e] = row_grad[((long long)f * B + u) * NASREC_EMB_DIM + e];
    }
}

// Four unique rows per 16-lane group (64 per CTA), all of their loads in flight before the first use: per row the kernel is
// a chain id -> (state, weight, gradient) -> two stores, i.e. latency, not bytes, at the sizes of a step.
constexpr int ADAGRAD_UNR = 4;
__global__ void __launch_bounds__(256) emb_adagrad_kernel(const int64_t* __restrict__ uniq,
                                                          const int* __restrict__ nuniq,
                                                          const float* __restrict__ row_grad,
                                                          float* const* __restrict__ tables,
                                                          float* const* __restrict__ states, int B, float lr,
                                                          float eps, const float* __restrict__ clip_coef) {
    pdl_enter();
    const int f = blockIdx.y;
    const int U = nuniq[f];
    const int e = threadIdx.x & 15, g = threadIdx.x >> 4;
    const int u0 = blockIdx.x * (16 * ADAGRAD_UNR) + g;
    if (u0 >= U) return;
    const float coef = clip_coef ? clip_coef[0] : 1.f;
    float* const tab = tables[f];
    float* const sta = states[f];
    long long o[ADAGRAD_UNR];
    bool ok[ADAGRAD_UNR];
#pragma unroll
    for (int j = 0; j < ADAGRAD_UNR; ++j) {
        const int u = u0 + 16 * j;
        ok[j] = u < U;
        o[j] = ok[j] ? uniq[(long long)f * B + u] * NASREC_EMB_DIM + e : 0;
    }
    float gr[ADAGRAD_UNR], st[ADAGRAD_UNR], w[ADAGRAD_UNR];
#pragma unroll
    for (int j = 0; j < ADAGRAD_UNR; ++j) {
        if (!ok[j]) continue;
        gr[j] = row_grad[((long long)f * B + u0 + 16 * j) * NASREC_EMB_DIM + e];
        st[j] = sta[o[j]];
        w[j] = tab[o[j]];
    }
#pragma unroll
    for (int j = 0; j < ADAGRAD_UNR; ++j) {
        if (!ok[j]) continue;
        const float gg = gr[j] * coef;
        const float s2 = st[j] + gg * gg;
        sta[o[j]] = s2;
        tab[o[j]] = w[j] - lr * (gg / (sqrtf(s2) + eps));
    }
}

template <int ITEMS>
int launch_sort_reduce(const int64_t* idx, const int64_t* num_rows, int* err_flag, const float* gout, int B, int Bpad, int F,
                       int64_t* uniq, int* nuniq, float* row_grad, float* sumsq, int* seg_scratch, void* stream) {
    size_t smem = (size_t)Bpad * sizeof(unsigned long long);
    if (ITEMS > 0) {
        const size_t tmp = sizeof(typename cub::BlockRadixSort<unsigned long long, 1024, (ITEMS > 0 ? ITEMS : 1)>::TempStorage);
        if (tmp > smem) smem = tmp;
    }
    static bool attr_set = false;
    if (!attr_set) {
        size_t cap = 16384 * 8;
        if (smem > cap) cap = smem;
        cudaError_t e = cudaFuncSetAttribute(emb_sort_reduce_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int threads = Bpad >= 1024 ? 1024 : (Bpad < 64 ? 64 : Bpad);
    nasrec_launch(emb_sort_reduce_kernel<ITEMS>, F, threads, smem, as_stream(stream), idx, num_rows, err_flag, gout, B, Bpad, F, uniq,
                  nuniq, row_grad, sumsq, seg_scratch);
    return nasrec_launch_status();
}

}  // namespace

extern "C" {

int nasrec_version(int* sm) {
    if (sm) *sm = 100;
    return NASREC_VERSION;
}

int nasrec_emb_gather_fwd(const float* const* tables, const int64_t* num_rows, const int64_t* idx, float* out,
                          int B, int F, int* err_flag, void* stream) {
    CHECK_ARG(tables && num_rows && idx && out && B > 0 && F > 0);
    const long long n_pairs = (long long)B * F;
    nasrec_launch(emb_gather_kernel, cdiv(n_pairs * 4, 256 * GATHER_UNR), 256, 0, as_stream(stream), tables, num_rows, idx, out,
                  n_pairs, F, err_flag);
    return nasrec_launch_status();
}

int nasrec_emb_grad_sort_reduce(const int64_t* idx, const float* gout, int B, int F, int64_t* uniq, int* nuniq,
                                float* row_grad, float* sumsq, int* seg_scratch, void* stream) {
    return nasrec_emb_grad_sort_reduce_checked(idx, nullptr, nullptr, gout, B, F, uniq, nuniq, row_grad, sumsq, seg_scratch, stream);
}

int nasrec_emb_grad_sort_reduce_checked(const int64_t* idx, const int64_t* num_rows, int* err_flag, const float* gout, int B,
                                        int F, int64_t* uniq, int* nuniq, float* row_grad, float* sumsq, int* seg_scratch,
                                        void* stream) {
    CHECK_ARG(idx && gout && uniq && nuniq && row_grad && sumsq && seg_scratch && B > 0 && F > 0);
    if (B > 16384) return NASREC_ETOOBIG;
    int Bpad = 32;
    while (Bpad < B) Bpad <<= 1;
    if (Bpad >= 1024 && Bpad <= 8192) {
        // radix variant: 1024 threads x (Bpad / 1024) keys
        switch (Bpad / 1024) {
            case 1: return launch_sort_reduce<1>(idx, num_rows, err_flag, gout, B, Bpad, F, uniq, nuniq, row_grad, sumsq, seg_scratch, stream);
            case 2: return launch_sort_reduce<2>(idx, num_rows, err_flag, gout, B, Bpad, F, uniq, nuniq, row_grad, sumsq, seg_scratch, stream);
            case 4: return launch_sort_reduce<4>(idx, num_rows, err_flag, gout, B, Bpad, F, uniq, nuniq, row_grad, sumsq, seg_scratch, stream);
            default: return launch_sort_reduce<8>(idx, num_rows, err_flag, gout, B, Bpad, F, uniq, nuniq, row_grad, sumsq, seg_scratch, stream);
        }
    }
    return launch_sort_reduce<0>(idx, num_rows, err_flag, gout, B, Bpad, F, uniq, nuniq, row_grad, sumsq, seg_scratch, stream);
}

int nasrec_emb_grad_to_dense(const int64_t* uniq, const int* nuniq, const float* row_grad,
                             float* const* grad_tables, int B, int F, void* stream) {
    CHECK_ARG(uniq && nuniq && row_grad && grad_tables && B > 0 && F > 0);
    dim3 grid(cdiv(B, 16), F);
    nasrec_launch(emb_to_dense_kernel, grid, 256, 0, as_stream(stream), uniq, nuniq, row_grad, grad_tables, B);
    return nasrec_launch_status();
}

int nasrec_emb_rowwise_adagrad(const int64_t* uniq, const int* nuniq, const float* row_grad, float* const* tables,
                               float* const* states, int B, int F, float lr, float eps, const float* clip_coef,
                               void* stream) {
    CHECK_ARG(uniq && nuniq && row_grad && tables && states && B > 0 && F > 0);
    dim3 grid(cdiv(B, 16 * ADAGRAD_UNR), F);
    nasrec_launch(emb_adagrad_kernel, grid, 256, 0, as_stream(stream), uniq, nuniq, row_grad, tables, states, B, lr, eps,
                                                            clip_coef);
    return nasrec_launch_status();
}

}  // extern "C"
