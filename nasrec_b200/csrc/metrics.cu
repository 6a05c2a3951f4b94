// Evaluation metrics and input transform on the device.
//
//  * nasrec_binary_metrics: accuracy@0.5, ROC-AUC and mean BCE log-loss over the concatenated
//    predictions of an evaluation pass (train_utils.py:158-178), so that EA scoring ships 3 doubles
//    per candidate to the host instead of 1.2 M predictions.  AUC is the Mann-Whitney statistic with
//    ties counted 1/2 (== sklearn.metrics.roc_auc_score), evaluated in exact integer arithmetic:
//      2*P*N*AUC = sum over positives i of ( 2*#{neg: p<p_i} + #{neg: p==p_i} ).
//    Ranking is on p = sigmoid(z) in fp32, as the reference ranks its fp32 sigmoid outputs.
//  * nasrec_input_transform: the reference's per-value Python transform of a raw batch
//    (data_pipes.py:135-175): dense log(max(0,x)+1); categorical int(hex,16) (or -1 when empty)
//    .fmod(N_f-1)+1, i.e. missing -> row 0, ids in [1, N_f-1].
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace {

constexpr int MB = 296;            // partial-sum blocks: 2 per SM, fixed so the reduction order is fixed
constexpr int MT = 256;

struct Partial { double loss; long long correct; };

__global__ void __launch_bounds__(MT) metrics_prepare_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                             long long n, uint32_t* __restrict__ key,
                                                             uint8_t* __restrict__ isneg, Partial* __restrict__ part) {
    pdl_enter();
    __shared__ double sl[MT / 32];
    __shared__ long long sc[MT / 32];
    double loss = 0.0;
    long long correct = 0;
    for (long long i = (long long)blockIdx.x * MT + threadIdx.x; i < n; i += (long long)gridDim.x * MT) {
        const float zi = z[i], yi = y[i];
        const float p = 1.f / (1.f + expf(-zi));
        key[i] = __float_as_uint(p);                       // p in [0,1]: unsigned order == float order
        isneg[i] = yi > 0.5f ? 0 : 1;
        loss += (double)(fmaxf(zi, 0.f) - zi * yi + log1pf(expf(-fabsf(zi))));
        correct += ((p > 0.5f ? 1.f : 0.f) == yi) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        loss += __shfl_xor_sync(0xffffffffu, loss, o);
        correct += __shfl_xor_sync(0xffffffffu, correct, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sl[wid] = loss; sc[wid] = correct; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0; long long c = 0;
        for (int w = 0; w < MT / 32; ++w) { l += sl[w]; c += sc[w]; }
        part[blockIdx.x].loss = l;
        part[blockIdx.x].correct = c;
    }
}

// negpre[i] = #negatives among sorted positions < i (n+1 entries; negpre[n] = N).
__global__ void __launch_bounds__(MT) metrics_pairs_kernel(const uint32_t* __restrict__ key, const uint8_t* __restrict__ isneg,
                                                           const int* __restrict__ negpre, long long n,
                                                           unsigned long long* __restrict__ twice_pairs) {
    pdl_enter();
    unsigned long long acc = 0;
    for (long long i = (long long)blockIdx.x * MT + threadIdx.x; i < n; i += (long long)gridDim.x * MT) {
        if (isneg[i]) continue;
        const uint32_t k = key[i];
        long long lo = 0, hi = i;                          // first position with key == k
        while (lo < hi) { long long m = (lo + hi) >> 1; if (key[m] < k) lo = m + 1; else hi = m; }
        const long long s = lo;
        lo = i + 1; hi = n;                                // first position with key > k
        while (lo < hi) { long long m = (lo + hi) >> 1; if (key[m] <= k) lo = m + 1; else hi = m; }
        const long long below = negpre[s], tied = negpre[lo] - negpre[s];
        acc += (unsigned long long)(2 * below + tied);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(twice_pairs, acc);       // integer adds: order-independent
}

__global__ void metrics_final_kernel(const Partial* __restrict__ part, int nparts, const int* __restrict__ negpre,
                                     long long n, const unsigned long long* __restrict__ twice_pairs,
                                     double* __restrict__ out) {
    pdl_enter();
    double l = 0.0; long long c = 0;
    for (int b = 0; b < nparts; ++b) { l += part[b].loss; c += part[b].correct; }
    const double N = (double)negpre[n], P = (double)n - N;
    out[0] = (double)c / (double)n;
    out[1] = (P > 0 && N > 0) ? (double)(*twice_pairs) / (2.0 * P * N) : nan("");
    out[2] = l / (double)n;
}

struct MetricsWs {
    size_t key_in, key_out, neg_in, neg_out, negpre, part, pairs, cub, total;
    size_t cub_bytes;
};

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

MetricsWs metrics_layout(long long n) {
    MetricsWs w;
    size_t o = 0;
    w.key_in = o;  o += up256(sizeof(uint32_t) * n);
    w.key_out = o; o += up256(sizeof(uint32_t) * n);
    w.neg_in = o;  o += up256(n);
    w.neg_out = o; o += up256(n + 1);
    w.negpre = o;  o += up256(sizeof(int) * (n + 1));
    w.part = o;    o += up256(sizeof(Partial) * MB);
    w.pairs = o;   o += 256;
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint8_t*)nullptr,
                                    (uint8_t*)nullptr, (int)n, 0, 32);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint8_t*)nullptr, (int*)nullptr, (int)n + 1);
    w.cub_bytes = a > b ? a : b;
    w.cub = o;     o += up256(w.cub_bytes);
    w.total = o;
    return w;
}

// ---------------------------------------------------------------- input transform
__device__ __forceinline__ int hex_digit(uint8_t c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

__global__ void __launch_bounds__(256) input_transform_kernel(const float* __restrict__ dense_raw, long long dsb,
                                                              long long dsc, int nd, const uint8_t* __restrict__ hex,
                                                              long long hsb, long long hsf, int width, int F,
                                                              const int64_t* __restrict__ num_rows, long long B,
                                                              float* __restrict__ int_x, int64_t* __restrict__ cat_x,
                                                              int* __restrict__ err_flag) {
    pdl_enter();
    const long long per = (long long)nd + F;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < B * per; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / per;
        const int c = (int)(t - b * per);
        if (c < nd) {                                      // log(max(0, x) + 1), rounded like the reference's two ops
            const float x = dense_raw[b * dsb + c * dsc];
            int_x[b * nd + c] = logf(fmaxf(x, 0.f) + 1.f);
        } else {
            const int f = c - nd;
            const uint8_t* s = hex + b * hsb + f * hsf;
            long long v = 0;
            int len = 0;
            bool bad = false;
            for (int i = 0; i < width; ++i) {
                const uint8_t ch = s[i];
                if (ch == 0) break;
                const int d = hex_digit(ch);
                if (d < 0) { bad = true; break; }
                v = (v << 4) | d;
                ++len;
            }
            if (bad && err_flag) atomicExch(err_flag, 1);
            if (len == 0) v = -1;                          // empty field (data_pipes.py:161)
            const long long m = num_rows[f] - 1;
            cat_x[b * F + f] = (m > 0 ? v % m : 0) + 1;    // C '%' truncates like Tensor.fmod: -1 % m == -1 -> row 0
        }
    }
}

}  // namespace

extern "C" int64_t nasrec_binary_metrics_ws_bytes(int64_t n) {
    if (n <= 0 || n >= (1LL << 31) - 1) return -1;
    return (int64_t)metrics_layout(n).total;
}

extern "C" int nasrec_binary_metrics(const float* logits, const float* y, int64_t n, void* ws, int64_t ws_bytes,
                                     double* out3, void* stream) {
    CHECK_ARG(logits && y && ws && out3 && n > 0 && n < (1LL << 31) - 1);
    const MetricsWs w = metrics_layout(n);
    CHECK_ARG(ws_bytes >= (int64_t)w.total);
    cudaStream_t st = as_stream(stream);
    char* base = (char*)ws;
    uint32_t* key_in = (uint32_t*)(base + w.key_in);
    uint32_t* key_out = (uint32_t*)(base + w.key_out);
    uint8_t* neg_in = (uint8_t*)(base + w.neg_in);
    uint8_t* neg_out = (uint8_t*)(base + w.neg_out);
    int* negpre = (int*)(base + w.negpre);
    Partial* part = (Partial*)(base + w.part);
    unsigned long long* pairs = (unsigned long long*)(base + w.pairs);
    cudaMemsetAsync(pairs, 0, sizeof(unsigned long long), st);
    // the scan runs over n+1 items so that negpre[n] = N; item n's own value is never summed (exclusive)
    nasrec_launch(metrics_prepare_kernel, MB, MT, 0, st, logits, y, n, key_in, neg_in, part);
    size_t cb = w.cub_bytes;
    cub::DeviceRadixSort::SortPairs(base + w.cub, cb, key_in, key_out, neg_in, neg_out, (int)n, 0, 32, st);
    cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(base + w.cub, cb, neg_out, negpre, (int)n + 1, st);
    nasrec_launch(metrics_pairs_kernel, MB, MT, 0, st, key_out, neg_out, negpre, n, pairs);
    nasrec_launch(metrics_final_kernel, 1, 1, 0, st, part, MB, negpre, n, pairs, out3);
    return nasrec_launch_status();
}

extern "C" int nasrec_input_transform(const float* dense_raw, int64_t dense_stride_b, int64_t dense_stride_c, int nd,
                                      const uint8_t* hex, int64_t hex_stride_b, int64_t hex_stride_f, int width, int F,
                                      const int64_t* num_rows, int64_t B, float* int_x, int64_t* cat_x,
                                      int* err_flag, void* stream) {
    CHECK_ARG(B >= 0 && nd >= 0 && F >= 0 && width >= 1 && width <= 15);
    if (B == 0 || nd + F == 0) return 0;
    CHECK_ARG((nd == 0 || (dense_raw && int_x)) && (F == 0 || (hex && num_rows && cat_x)));
    const long long total = B * ((long long)nd + F);
    const int blocks = (int)((total + 255) / 256 < 148LL * 8 ? (total + 255) / 256 : 148 * 8);
    nasrec_launch(input_transform_kernel, blocks, 256, 0, as_stream(stream), dense_raw, dense_stride_b, dense_stride_c, nd, hex, hex_stride_b,
                                                                  hex_stride_f, width, F, num_rows, B, int_x, cat_x,
                                                                  err_flag);
    return nasrec_launch_status();
}
