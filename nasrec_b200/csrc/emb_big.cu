// Sorted-row reduction of the embedding gradient for LARGE batches (multi-CTA, any B): the data-parallel global batch
// (world x B ids per table) and the 16 K-sample KDD batches, where the one-CTA-per-table shared-memory sort of emb.cu
// (<= 16384 ids, O(n log^2 n), 26 of 148 SMs) becomes the longest kernel of the step.
//
//   keys   (f << 27 | row, sample) pairs of all F tables in ONE array; out-of-range ids get the all-ones key
//   sort   cub::DeviceRadixSort (stable LSD): tables in order, rows ascending, samples ascending within a row
//   heads  head flags + inclusive scan -> global unique index g; per-table bases -> output slot f * B + u
//   sum    16 lanes per unique row add its duplicates in ascending sample order; rows with more than LONG duplicates
//          (hot rows of a Zipf batch) are summed by a whole CTA, 16 strided partial sums combined in a fixed order
//   sumsq  one CTA per table over its reduced rows (fixed order)
// Everything is deterministic: the same ids and gradients give the same bits, which is what keeps data-parallel
// replicas identical.  Replaces F x embedding_dense_backward (nn.Embedding, nasrec/supernet/supernet.py:407).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace {

constexpr int ROW_BITS = 27;                 // rows per table < 2^27 (largest shipped table: 22.1 M rows)
constexpr uint32_t ROW_MASK = (1u << ROW_BITS) - 1;
constexpr uint32_t BAD_KEY = 0xFFFFFFFFu;
constexpr int LONG = 256;                    // duplicates above which a row is summed by a whole CTA

__global__ void __launch_bounds__(256) big_keys_kernel(const int64_t* __restrict__ idx, const int64_t* __restrict__ num_rows,
                                                       int B, int F, uint32_t* __restrict__ keys, int* __restrict__ vals,
                                                       int* err_flag) {
    pdl_enter();
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)B * F) return;
    const int b = (int)(t / F), f = (int)(t % F);
    const unsigned long long row = (unsigned long long)idx[t];
    const unsigned long long lim = num_rows ? (unsigned long long)num_rows[f] : (1ull << ROW_BITS);
    uint32_t key = BAD_KEY;
    if (row < lim && row <= ROW_MASK) key = ((uint32_t)f << ROW_BITS) | (uint32_t)row;
    else if (err_flag) atomicOr(err_flag, 1);
    keys[(long long)f * B + b] = key;
    vals[(long long)f * B + b] = b;
}

__global__ void __launch_bounds__(256) big_flags_kernel(const uint32_t* __restrict__ keys, long long n, int* __restrict__ flags) {
    pdl_enter();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    flags[i] = (k != BAD_KEY && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// tbase[f] = number of unique rows in tables < f; tbase[F] = G (all unique rows); also the count of valid keys
__global__ void __launch_bounds__(256) big_tbase_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ rank, long long n,
                                                        int F, int* __restrict__ tbase, int* __restrict__ nvalid) {
    pdl_enter();
    // boundaries: first valid position of each table
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        if (k == BAD_KEY) {
            if (i == 0 || keys[i - 1] != BAD_KEY) *nvalid = (int)i;
            continue;
        }
        const int f = (int)(k >> ROW_BITS);
        if (i == 0 || (int)(keys[i - 1] >> ROW_BITS) != f) tbase[f] = rank[i] - 1;
        if (i == n - 1) *nvalid = (int)n;
    }
}

// tables without a single valid id take the base of the next table; nuniq[f] = tbase[f+1] - tbase[f]
__global__ void big_fixup_kernel(int* __restrict__ tbase, const int* __restrict__ rank, const int* __restrict__ nvalid, int F,
                                 int* __restrict__ nuniq, int* __restrict__ nlong, int* __restrict__ seg_begin) {
    pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nv = *nvalid;
    const int G = nv > 0 ? rank[nv - 1] : 0;
    tbase[F] = G;
    seg_begin[G] = nv;                        // end of the last segment
    for (int f = F - 1; f >= 0; --f)
        if (tbase[f] < 0) tbase[f] = tbase[f + 1];
    for (int f = 0; f < F; ++f) nuniq[f] = tbase[f + 1] - tbase[f];
    *nlong = 0;
}

// per head: output slot, segment begin; segment ends are the next head's begin (seg_begin[G] = number of valid keys)
__global__ void __launch_bounds__(256) big_heads_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ flags,
                                                        const int* __restrict__ rank, const int* __restrict__ tbase,
                                                        long long n, int B, int* __restrict__ seg_begin, int* __restrict__ slot,
                                                        int64_t* __restrict__ uniq) {
    pdl_enter();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const uint32_t k = keys[i];
    const int f = (int)(k >> ROW_BITS);
    const int g = rank[i] - 1;
    const int u = g - tbase[f];
    seg_begin[g] = (int)i;
    slot[g] = f * B + u;
    uniq[(long long)f * B + u] = (int64_t)(k & ROW_MASK);
}

// 16 lanes per unique row; rows longer than LONG are queued for the CTA-wide kernel
__global__ void __launch_bounds__(256) big_sum_kernel(const int* __restrict__ vals, const uint32_t* __restrict__ keys,
                                                      const int* __restrict__ seg_begin, const int* __restrict__ slot,
                                                      const int* __restrict__ tbase, int F_tables, const float* __restrict__ gout, int F,
                                                      float* __restrict__ row_grad, int* __restrict__ long_list, int* __restrict__ nlong) {
    pdl_enter();
    const int G = tbase[F_tables];
    const int e = threadIdx.x & 15;
    for (int g = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); g < G; g += gridDim.x * (blockDim.x >> 4)) {
        const int s0 = seg_begin[g], s1 = seg_begin[g + 1];
        if (s1 - s0 > LONG) {
            if (e == 0) long_list[atomicAdd(nlong, 1)] = g;
            continue;
        }
        const int f = (int)(keys[s0] >> ROW_BITS);
        float acc = 0.f;
        int i = s0;
        for (; i + 8 <= s1; i += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(gout + ((long long)vals[i + j] * F + f) * NASREC_EMB_DIM + e);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j];
        }
        for (; i < s1; ++i) acc += __ldg(gout + ((long long)vals[i] * F + f) * NASREC_EMB_DIM + e);
        row_grad[(long long)slot[g] * NASREC_EMB_DIM + e] = acc;
    }
}

// one CTA per long row: 16 groups of 16 lanes take the duplicates round-robin (each group in ascending sample order),
// the 16 partial sums are added in group order
__global__ void __launch_bounds__(256) big_sum_long_kernel(const int* __restrict__ vals, const uint32_t* __restrict__ keys,
                                                           const int* __restrict__ seg_begin, const int* __restrict__ slot,
                                                           const float* __restrict__ gout, int F, float* __restrict__ row_grad,
                                                           const int* __restrict__ long_list, const int* __restrict__ nlong) {
    pdl_enter();
    __shared__ float part[16][17];
    const int e = threadIdx.x & 15, grp = threadIdx.x >> 4;
    const int n = *nlong;
    for (int li = blockIdx.x; li < n; li += gridDim.x) {
        // the queue was filled in arbitrary order; each row's sum does not depend on it
        const int g = long_list[li];
        const int s0 = seg_begin[g], s1 = seg_begin[g + 1];
        const int f = (int)(keys[s0] >> ROW_BITS);
        float acc = 0.f;
        for (int i = s0 + grp; i < s1; i += 16) acc += __ldg(gout + ((long long)vals[i] * F + f) * NASREC_EMB_DIM + e);
        part[grp][e] = acc;
        __syncthreads();
        if (grp == 0) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) t += part[j][e];
            row_grad[(long long)slot[g] * NASREC_EMB_DIM + e] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) big_sumsq_kernel(const float* __restrict__ row_grad, const int* __restrict__ nuniq, int B,
                                                         float* __restrict__ sumsq) {
    pdl_enter();
    __shared__ float red[34];
    const int f = blockIdx.x;
    const long long n = (long long)nuniq[f] * NASREC_EMB_DIM;
    const float* p = row_grad + (long long)f * B * NASREC_EMB_DIM;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = fmaf(p[i], p[i], acc);
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) sumsq[f] = tot;
}

struct Ws {
    size_t keys_in, keys_out, vals_in, vals_out, flags, rank, seg_begin, slot, tbase, scal, long_list, cub, cub_bytes, total;
};

Ws layout(int B, int F) {
    const size_t n = (size_t)B * F;
    Ws w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    w.keys_in = take(n * 4); w.keys_out = take(n * 4); w.vals_in = take(n * 4); w.vals_out = take(n * 4);
    w.flags = take(n * 4); w.rank = take(n * 4); w.seg_begin = take((n + 1) * 4); w.slot = take(n * 4);
    w.tbase = take((size_t)(F + 2) * 4); w.scal = take(64); w.long_list = take((n / LONG + F + 1) * 4);
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, (int)n);
    w.cub_bytes = a > b ? a : b;
    w.cub = take(w.cub_bytes);
    w.total = o;
    return w;
}

}  // namespace

extern "C" {

int64_t nasrec_emb_grad_sort_reduce_big_ws_bytes(int B, int F) {
    if (B <= 0 || F <= 0 || F > 31) return 0;
    return (int64_t)layout(B, F).total;
}

int nasrec_emb_grad_sort_reduce_big(const int64_t* idx, const int64_t* num_rows, int* err_flag, const float* gout, int B, int F,
                                    int64_t* uniq, int* nuniq, float* row_grad, float* sumsq, void* ws, int64_t ws_bytes,
                                    void* stream) {
    CHECK_ARG(idx && gout && uniq && nuniq && row_grad && sumsq && ws && B > 0 && F > 0 && F <= 31);
    if ((long long)B * F >= (1ll << 31)) return NASREC_ETOOBIG;
    const Ws w = layout(B, F);
    if ((int64_t)w.total > ws_bytes) return NASREC_ENOSPACE;
    cudaStream_t st = as_stream(stream);
    char* base = (char*)ws;
    const long long n = (long long)B * F;
    uint32_t* keys_in = (uint32_t*)(base + w.keys_in);
    uint32_t* keys = (uint32_t*)(base + w.keys_out);
    int* vals_in = (int*)(base + w.vals_in);
    int* vals = (int*)(base + w.vals_out);
    int* flags = (int*)(base + w.flags);
    int* rank = (int*)(base + w.rank);
    int* seg_begin = (int*)(base + w.seg_begin);
    int* slot = (int*)(base + w.slot);
    int* tbase = (int*)(base + w.tbase);
    int* nvalid = (int*)(base + w.scal);
    int* nlong = nvalid + 1;
    int* long_list = (int*)(base + w.long_list);
    const unsigned gn = (unsigned)((n + 255) / 256);
    nasrec_launch(big_keys_kernel, gn, 256, 0, st, idx, num_rows, B, F, keys_in, vals_in, err_flag);
    size_t cb = w.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(base + w.cub, cb, keys_in, keys, vals_in, vals, (int)n, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    nasrec_launch(big_flags_kernel, gn, 256, 0, st, (const uint32_t*)keys, n, flags);
    cb = w.cub_bytes;
    e = cub::DeviceScan::InclusiveSum(base + w.cub, cb, flags, rank, (int)n, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(tbase, 0xFF, (size_t)(F + 2) * 4, st);      // -1: "no valid id seen"
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(nvalid, 0, 8, st);
    if (e != cudaSuccess) return (int)e;
    nasrec_launch(big_tbase_kernel, gn > 592 ? 592u : gn, 256, 0, st, (const uint32_t*)keys, (const int*)rank, n, F, tbase, nvalid);
    nasrec_launch(big_fixup_kernel, 1, 32, 0, st, tbase, (const int*)rank, (const int*)nvalid, F, nuniq, nlong, seg_begin);
    nasrec_launch(big_heads_kernel, gn, 256, 0, st, (const uint32_t*)keys, (const int*)flags, (const int*)rank, (const int*)tbase,
                  n, B, seg_begin, slot, uniq);
    const unsigned gs = (unsigned)((n + 15) / 16 < 2368 ? (n + 15) / 16 : 2368);
    nasrec_launch(big_sum_kernel, gs, 256, 0, st, (const int*)vals, (const uint32_t*)keys, (const int*)seg_begin, (const int*)slot,
                  (const int*)tbase, F, gout, F, row_grad, long_list, nlong);
    const unsigned gl = (unsigned)(n / LONG + 1 < 592 ? n / LONG + 1 : 592);
    nasrec_launch(big_sum_long_kernel, gl, 256, 0, st, (const int*)vals, (const uint32_t*)keys, (const int*)seg_begin,
                  (const int*)slot, gout, F, row_grad, (const int*)long_list, (const int*)nlong);
    nasrec_launch(big_sumsq_kernel, (unsigned)F, 1024, 0, st, (const float*)row_grad, (const int*)nuniq, B, sumsq);
    return nasrec_launch_status();
}

}  // extern "C"
