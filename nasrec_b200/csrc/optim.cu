// Step-body kernels: BCE-with-logits (+gradient), global-norm clip coefficient,
// multi-tensor Adagrad.
//
// Replaces BCEWithLogitsLoss + backward seed (nasrec/utils/train_utils.py:266,283),
// torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0) (:285) and
// torch.optim.Adagrad(eps=1e-2).step() (:286, nasrec/train_supernet.py:121-123),
// which in eager PyTorch are ~10 elementwise launches per parameter tensor.
#include "common.cuh"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

long long g_nasrec_launch_ns = 0, g_nasrec_launch_count = 0, g_nasrec_launch_total = 0;
int g_nasrec_host_prof = 0;
int g_nasrec_trace = 0;

namespace {
struct TraceRec {
    const void* func;
    cudaEvent_t e0, e1;
};
std::vector<TraceRec> g_trace;
size_t g_trace_used = 0;
}  // namespace

void nasrec_trace_begin(const void* func, cudaStream_t st) {
    if (g_trace_used == g_trace.size()) {
        TraceRec r{};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        g_trace.push_back(r);
    }
    g_trace[g_trace_used].func = func;
    cudaEventRecord(g_trace[g_trace_used].e0, st);
}
void nasrec_trace_end(cudaStream_t st) {
    cudaEventRecord(g_trace[g_trace_used].e1, st);
    ++g_trace_used;
}

namespace {

__global__ void __launch_bounds__(1024) bce_kernel(const float* __restrict__ z, const float* __restrict__ y, int B,
                                                   float grad_scale, float* __restrict__ loss,
                                                   float* __restrict__ dz) {
    pdl_enter();
    __shared__ float red[34];
    float acc = 0.f;
    const float invB = 1.f / (float)B;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float v = z[i], t = y[i];
        acc += fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
        if (dz) dz[i] = (1.f / (1.f + expf(-v)) - t) * invB * grad_scale;
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0 && loss) loss[0] = tot * invB;
}

constexpr int MT_MAX = 96;            // tensors per launch
constexpr int MT_CHUNK = 16384;       // elements per CTA

struct SumsqPack {
    int n;
    int pad_;
    const float* g[MT_MAX];
    long long size[MT_MAX];
    int chunk0[MT_MAX + 1];           // first chunk index of tensor i (prefix sum)
};

__global__ void __launch_bounds__(256) sumsq_kernel(const __grid_constant__ SumsqPack pk, float* __restrict__ partial,
                                                    int partial_off) {
    pdl_enter();
    __shared__ float red[34];
    const int c = blockIdx.x;
    int t = 0;
    while (t + 1 < pk.n && pk.chunk0[t + 1] <= c) ++t;
    const long long beg = (long long)(c - pk.chunk0[t]) * MT_CHUNK;
    const long long end = min(pk.size[t], beg + MT_CHUNK);
    const float* g = pk.g[t];
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {          // chunk starts are multiples of 4: 16-byte loads
        const long long n4 = (end - beg) >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(g + beg);
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 v = __ldg(g4 + i);
            acc = fmaf(v.x, v.x, acc);
            acc = fmaf(v.y, v.y, acc);
            acc = fmaf(v.z, v.z, acc);
            acc = fmaf(v.w, v.w, acc);
        }
        for (long long i = beg + (n4 << 2) + threadIdx.x; i < end; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
            const float v = g[i];
            acc = fmaf(v, v, acc);
        }
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) partial[partial_off + c] = tot;
}

__global__ void __launch_bounds__(1024) clip_finalize_kernel(const float* __restrict__ partial, int n_partial,
                                                             const float* __restrict__ extra, int n_extra,
                                                             float max_norm, float* __restrict__ out) {
    pdl_enter();
    __shared__ float red[34];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
    for (int i = threadIdx.x; i < n_extra; i += blockDim.x) acc += extra[i];
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const float norm = sqrtf(tot);
        out[0] = norm;
        out[1] = fminf(1.f, max_norm / (norm + 1e-6f));
    }
}

struct AdagradPack {
    int n;
    int pad_;
    float* w[MT_MAX];
    const float* g[MT_MAX];
    float* s[MT_MAX];
    long long size[MT_MAX];
    int chunk0[MT_MAX + 1];
    // optional pre-split copies of w for the TMA-fed GEMM (gemm_tma.cuh): hi = rn_tf32(w), lo = w - hi, row stride ldp
    float* hi[MT_MAX];
    float* lo[MT_MAX];
    int cols[MT_MAX];
    int ldp[MT_MAX];
    int bf16;               // planes hold rn_bf16(w) (GEMM mode 2) instead of rn_tf32(w)
    int first[MT_MAX];      // plane column of w column c: c + (c >= first ? shift : 0), shift = (4 - first % 4) % 4
};

// bf16 != 0 (GEMM mode 2): hi = rn_bf16(w), what the single-pass bf16 product multiplies (lo is not used then)
__device__ __forceinline__ void plane_split(float w, float& hi, float& lo, int bf16) {
    const uint32_t u = __float_as_uint(w);
    hi = bf16 ? __uint_as_float((u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u) : __uint_as_float((u + 0x1000u) & 0xFFFFE000u);
    lo = w - hi;
}

__global__ void __launch_bounds__(256) adagrad_kernel(const __grid_constant__ AdagradPack pk, float lr, float eps,
                                                      const float* __restrict__ clip_coef) {
    pdl_enter();
    const int c = blockIdx.x;
    int t = 0;
    while (t + 1 < pk.n && pk.chunk0[t + 1] <= c) ++t;
    const long long beg = (long long)(c - pk.chunk0[t]) * MT_CHUNK;
    const long long end = min(pk.size[t], beg + MT_CHUNK);
    const float coef = clip_coef ? clip_coef[0] : 1.f;
    float* w = pk.w[t];
    const float* g = pk.g[t];
    float* s = pk.s[t];
    float* ph = pk.hi[t];
    float* pl = pk.lo[t];
    const unsigned cols = (unsigned)pk.cols[t];
    const long long ldp = pk.ldp[t];
    const unsigned first = (unsigned)pk.first[t], shift = (4u - (first & 3u)) & 3u;
    auto plane = [&](long long i, float wi) {          // flat index of w -> (row, col) of the padded planes
        const unsigned r = (unsigned)i / cols, c = (unsigned)i - r * cols;
        const unsigned pc = c + (c >= first ? shift : 0u);
        float h, l;
        plane_split(wi, h, l, pk.bf16);
        ph[r * ldp + pc] = h;
        pl[r * ldp + pc] = l;
    };
    auto upd = [&](float gi, float& si, float& wi) {
        const float gv = gi * coef;
        const float sv = fmaf(gv, gv, si);
        si = sv;
        wi = wi - lr * (gv / (sqrtf(sv) + eps));
    };
    if (((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0) {
        // three streams in, two out, 16 bytes per access (chunk starts are multiples of 4 elements)
        const long long n4 = (end - beg) >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(g + beg);
        float4* s4 = reinterpret_cast<float4*>(s + beg);
        float4* w4 = reinterpret_cast<float4*>(w + beg);
        // Four 16-byte groups per thread and pass, every load issued before the first use; the (row, column) of a group in the
        // padded planes advances incrementally (one division per thread instead of one per element: the reference layout's
        // row length, e.g. 1037, is not a power of two).  20 B/parameter + 8 B for the planes: ~220 MB per step.
        constexpr int AU = 4;
        unsigned pr0 = 0, pc0 = 0;                         // (row, column in w) of this thread's first element of the pass
        if (ph) {
            const unsigned e0 = (unsigned)(beg + ((long long)threadIdx.x << 2));
            pr0 = e0 / cols;
            pc0 = e0 - pr0 * cols;
        }
        const unsigned step_r = (unsigned)(blockDim.x * 4) / cols, step_c = (unsigned)(blockDim.x * 4) - step_r * cols;
        const bool same_layout = ldp == (long long)cols && shift == 0;
        for (long long i0 = threadIdx.x; i0 < n4; i0 += (long long)blockDim.x * AU) {
            float4 gv[AU], sv[AU], wv[AU];
#pragma unroll
            for (int u = 0; u < AU; ++u) {
                const long long i = i0 + (long long)u * blockDim.x;
                if (i < n4) {
                    gv[u] = __ldg(g4 + i);
                    sv[u] = s4[i];
                    wv[u] = w4[i];
                }
            }
#pragma unroll
            for (int u = 0; u < AU; ++u) {
                const long long i = i0 + (long long)u * blockDim.x;
                if (i >= n4) break;
                upd(gv[u].x, sv[u].x, wv[u].x);
                upd(gv[u].y, sv[u].y, wv[u].y);
                upd(gv[u].z, sv[u].z, wv[u].z);
                upd(gv[u].w, sv[u].w, wv[u].w);
                s4[i] = sv[u];
                w4[i] = wv[u];
                if (ph) {
                    if (same_layout) {      // planes have w's own layout: 16-byte stores
                        const long long e = beg + (i << 2);
                        float4 h4, l4;
                        plane_split(wv[u].x, h4.x, l4.x, pk.bf16);
                        plane_split(wv[u].y, h4.y, l4.y, pk.bf16);
                        plane_split(wv[u].z, h4.z, l4.z, pk.bf16);
                        plane_split(wv[u].w, h4.w, l4.w, pk.bf16);
                        *reinterpret_cast<float4*>(ph + e) = h4;
                        *reinterpret_cast<float4*>(pl + e) = l4;
                    } else {
                        const float vals[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
                        unsigned r = pr0, c = pc0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float h, l;
                            plane_split(vals[q], h, l, pk.bf16);
                            const long long o = (long long)r * ldp + c + (c >= first ? shift : 0u);
                            ph[o] = h;
                            pl[o] = l;
                            if (++c == cols) { c = 0; ++r; }
                        }
                    }
                    pr0 += step_r;
                    pc0 += step_c;
                    if (pc0 >= cols) { pc0 -= cols; ++pr0; }
                }
            }
        }
        for (long long i = beg + (n4 << 2) + threadIdx.x; i < end; i += blockDim.x) {
            upd(g[i], s[i], w[i]);
            if (ph) plane(i, w[i]);
        }
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
            upd(g[i], s[i], w[i]);
            if (ph) plane(i, w[i]);
        }
    }
}

// hi/lo planes of a [rows, cols] weight (row stride ldw) with row stride ldp; pad columns are left alone (zero)
__global__ void __launch_bounds__(256) planes_refresh_kernel(const float* __restrict__ w, long long ldw, int rows, int cols,
                                                             int first, float* __restrict__ hi, float* __restrict__ lo,
                                                             long long ldp, int bf16) {
    pdl_enter();
    const int shift = (4 - (first & 3)) & 3;
    const long long total = (long long)rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        float h, l;
        plane_split(w[r * ldw + c], h, l, bf16);
        const long long o = r * ldp + c + (c >= first ? shift : 0);
        hi[o] = h;
        lo[o] = l;
    }
}

}  // namespace

extern "C" {

int64_t nasrec_host_prof(int what) {
    switch (what) {
        case 0: g_nasrec_host_prof = 0; return 0;
        case 1: g_nasrec_host_prof = 1; g_nasrec_launch_ns = 0; g_nasrec_launch_count = 0; return 0;
        case 2: return g_nasrec_launch_ns;
        case 4: return g_nasrec_launch_total;      // kernels launched by the library since it was loaded (always counted)
        case 10: g_nasrec_trace = 1; g_trace_used = 0; return 0;
        case 11: {            // stop; synchronise; append "name launches total_us" lines to $NASREC_TRACE_FILE; returns launches
            g_nasrec_trace = 0;
            std::map<std::string, std::pair<long long, double>> agg;
            for (size_t i = 0; i < g_trace_used; ++i) {
                cudaEventSynchronize(g_trace[i].e1);
                float ms = 0;
                if (cudaEventElapsedTime(&ms, g_trace[i].e0, g_trace[i].e1) != cudaSuccess) continue;
                const char* name = nullptr;
                if (cudaFuncGetName(&name, g_trace[i].func) != cudaSuccess || !name) name = "?";
                auto& a = agg[name];
                a.first += 1;
                a.second += ms * 1e3;
            }
            if (const char* path = getenv("NASREC_TRACE_FILE")) {
                if (FILE* f = fopen(path, "a")) {
                    for (auto& kv : agg) fprintf(f, "%s %lld %.1f\n", kv.first.c_str(), kv.second.first, kv.second.second);
                    fclose(f);
                }
            }
            cudaGetLastError();
            return (int64_t)g_trace_used;
        }
        default: return g_nasrec_launch_count;
    }
}

int nasrec_bce_fwd_bwd(const float* logits, const float* y, int B, float grad_scale, float* loss, float* dlogits,
                       void* stream) {
    CHECK_ARG(logits && y && B > 0);
    nasrec_launch(bce_kernel, 1, 1024, 0, as_stream(stream), logits, y, B, grad_scale, loss, dlogits);
    return nasrec_launch_status();
}

int64_t nasrec_sumsq_ws_floats(const int64_t* sizes, int n) {
    int64_t c = 0;
    for (int i = 0; i < n; ++i) c += (sizes[i] + MT_CHUNK - 1) / MT_CHUNK;
    return c > 0 ? c : 1;
}

int nasrec_grad_norm_clip(const float* const* grads, const int64_t* sizes, int n, const float* extra_sumsq,
                          int n_extra, float max_norm, float* partial, float* out, void* stream) {
    CHECK_ARG(partial && out && n >= 0 && n_extra >= 0 && (n == 0 || (grads && sizes)));
    cudaStream_t st = as_stream(stream);
    int done = 0, poff = 0;
    while (done < n) {
        SumsqPack pk{};
        int m = 0, chunks = 0;
        for (; done + m < n && m < MT_MAX; ++m) {
            pk.g[m] = grads[done + m];
            pk.size[m] = sizes[done + m];
            pk.chunk0[m] = chunks;
            chunks += (int)((sizes[done + m] + MT_CHUNK - 1) / MT_CHUNK);
        }
        pk.chunk0[m] = chunks;
        pk.n = m;
        if (chunks > 0) {
            nasrec_launch(sumsq_kernel, chunks, 256, 0, st, pk, partial, poff);
            int rc = nasrec_launch_status();
            if (rc) return rc;
        }
        poff += chunks;
        done += m;
    }
    nasrec_launch(clip_finalize_kernel, 1, 1024, 0, st, partial, poff, extra_sumsq, n_extra, max_norm, out);
    return nasrec_launch_status();
}

int nasrec_planes_refresh(const float* W, int64_t ldw, int rows, int cols, int first, float* hi, float* lo, int64_t ldp,
                          void* stream) {
    CHECK_ARG(W && hi && lo && rows > 0 && cols > 0 && first >= 0 && first <= cols && ldw >= cols);
    CHECK_ARG(ldp >= cols + ((4 - (first & 3)) & 3));
    const long long total = (long long)rows * cols;
    long long gx = (total + 255) / 256;
    if (gx > 2368) gx = 2368;
    nasrec_launch(planes_refresh_kernel, (unsigned)gx, 256, 0, as_stream(stream), W, (long long)ldw, rows, cols, first, hi, lo, (long long)ldp,
                  nasrec_get_gemm_mode() == 2 ? 1 : 0);
    return nasrec_launch_status();
}

int nasrec_adagrad_multi(float* const* w, const float* const* grads, float* const* state, const int64_t* sizes,
                         int n, float lr, float eps, const float* clip_coef, void* stream) {
    return nasrec_adagrad_multi_planes(w, grads, state, sizes, n, lr, eps, clip_coef, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int nasrec_adagrad_multi_planes(float* const* w, const float* const* grads, float* const* state, const int64_t* sizes,
                                int n, float lr, float eps, const float* clip_coef, float* const* hi, float* const* lo,
                                const int* cols, const int* first, const int64_t* ldp, void* stream) {
    CHECK_ARG(n >= 0 && (n == 0 || (w && grads && state && sizes)));
    CHECK_ARG(!hi || (lo && cols && first && ldp));
    cudaStream_t st = as_stream(stream);
    int done = 0;
    while (done < n) {
        AdagradPack pk{};
        pk.bf16 = nasrec_get_gemm_mode() == 2 ? 1 : 0;
        int m = 0, chunks = 0;
        for (; done + m < n && m < MT_MAX; ++m) {
            pk.w[m] = w[done + m];
            pk.g[m] = grads[done + m];
            pk.s[m] = state[done + m];
            pk.size[m] = sizes[done + m];
            pk.chunk0[m] = chunks;
            chunks += (int)((sizes[done + m] + MT_CHUNK - 1) / MT_CHUNK);
            if (hi && hi[done + m]) {
                if (sizes[done + m] >= (1ll << 31)) return NASREC_ETOOBIG;
                pk.hi[m] = hi[done + m];
                pk.lo[m] = lo[done + m];
                pk.cols[m] = cols[done + m];
                pk.ldp[m] = (int)ldp[done + m];
                pk.first[m] = first[done + m];
            }
        }
        pk.chunk0[m] = chunks;
        pk.n = m;
        if (chunks > 0) {
            nasrec_launch(adagrad_kernel, chunks, 256, 0, st, pk, lr, eps, clip_coef);
            int rc = nasrec_launch_status();
            if (rc) return rc;
        }
        done += m;
    }
    return 0;
}

}  // extern "C"
