// Per-sample interaction kernels: DotProduct lower-triangle, FactorizationMachine3D,
// SigmoidGating elementwise stage, padded-concat materialisation.
//
// Replaces torch.bmm + tril_indices gather (nasrec/supernet/modules.py:366-383),
// the FM reductions (:736-738), sigmoid/multiply (:580-582) and
// _pad_2Dtensors_if_needed + torch.cat (:403-430).
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int TMAX = 48;   // 1 + P, P <= 45 (round(sqrt(2*1024)))

__device__ __forceinline__ void tril_pair(int r, int& i, int& j) {
    // r = i(i-1)/2 + j, 0 <= j < i  (torch.tril_indices(n, n, -1) row-major order)
    i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)r)) * 0.5f);
    while (i * (i - 1) / 2 > r) --i;
    while ((i + 1) * i / 2 <= r) ++i;
    j = r - i * (i - 1) / 2;
}

__global__ void __launch_bounds__(256) dot_tril_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ y, long long ybs, int P,
                                                           float* __restrict__ R, long long ldr, int B) {
    pdl_enter();
    __shared__ float T[TMAX][17];
    const int Tn = P + 1, NR = Tn * (Tn - 1) / 2;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int t = threadIdx.x; t < Tn * 16; t += blockDim.x) {
            const int i = t >> 4, e = t & 15;
            T[i][e] = i == 0 ? x[(long long)b * ldx + e] : y[(long long)b * ybs + (i - 1) * 16 + e];
        }
        __syncthreads();
        for (int r = threadIdx.x; r < NR; r += blockDim.x) {
            int i, j;
            tril_pair(r, i, j);
            float acc = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) acc = fmaf(T[i][e], T[j][e], acc);
            R[(long long)b * ldr + r] = acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) dot_tril_bwd_kernel(const float* __restrict__ dR, long long ldr,
                                                           const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ y, long long ybs, int P,
                                                           float* __restrict__ dx, long long lddx,
                                                           float* __restrict__ dy, long long dybs, int B) {
    pdl_enter();
    __shared__ float T[TMAX][17];
    __shared__ float G[TMAX * (TMAX - 1) / 2];
    const int Tn = P + 1, NR = Tn * (Tn - 1) / 2;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int t = threadIdx.x; t < Tn * 16; t += blockDim.x) {
            const int i = t >> 4, e = t & 15;
            T[i][e] = i == 0 ? x[(long long)b * ldx + e] : y[(long long)b * ybs + (i - 1) * 16 + e];
        }
        for (int r = threadIdx.x; r < NR; r += blockDim.x) G[r] = dR[(long long)b * ldr + r];
        __syncthreads();
        for (int t = threadIdx.x; t < Tn * 16; t += blockDim.x) {
            const int i = t >> 4, e = t & 15;
            float acc = 0.f;
            const int base = i * (i - 1) / 2;
            for (int j = 0; j < i; ++j) acc = fmaf(G[base + j], T[j][e], acc);
            for (int j = i + 1; j < Tn; ++j) acc = fmaf(G[j * (j - 1) / 2 + i], T[j][e], acc);
            if (i == 0) {
                if (dx) dx[(long long)b * lddx + e] = acc;
            } else if (dy) {
                dy[(long long)b * dybs + (i - 1) * 16 + e] = acc;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ FM
__global__ void __launch_bounds__(256) fm_fwd_kernel(const float* __restrict__ x, long long xbs, int rows,
                                                     float* __restrict__ ix, int B) {
    pdl_enter();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 16) return;
    const int b = t >> 4, e = t & 15;
    const float* xp = x + (long long)b * xbs + e;
    float s = 0.f, q = 0.f;
    for (int r = 0; r < rows; ++r) {
        const float v = xp[r * 16];
        s += v;
        q = fmaf(v, v, q);
    }
    ix[t] = s * s - q;
}

__global__ void __launch_bounds__(256) fm_bwd_kernel(const float* __restrict__ dix, const float* __restrict__ x,
                                                     long long xbs, int rows, const float* __restrict__ dx_in,
                                                     long long dxin_bs, float* __restrict__ dx, long long dxbs,
                                                     int B) {
    pdl_enter();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 16) return;
    const int b = t >> 4, e = t & 15;
    const float* xp = x + (long long)b * xbs + e;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += xp[r * 16];
    const float g2 = 2.f * dix[t];
    float* op = dx + (long long)b * dxbs + e;
    const float* ip = dx_in ? dx_in + (long long)b * dxin_bs + e : nullptr;
    for (int r = 0; r < rows; ++r) {
        const float v = g2 * (s - xp[r * 16]);
        op[r * 16] = ip ? ip[r * 16] + v : v;
    }
}

// Same arithmetic with one CTA per sample: at B = 512 the kernel above is 32 CTAs whose threads each walk `rows`
// strided loads twice, a few loads in flight at a time (14 us, ncu).  Here the CTA stages the sample's rows in shared
// memory with one round of coalesced loads, 16 threads form the column sums in the same row order (so the result is
// bit-identical) and all 256 threads write the rows.  dx_in may be dx (in-place accumulate): an element is read and
// written by the same thread.
constexpr int FM_ROWS_SMEM = 128;
__global__ void __launch_bounds__(256) fm_bwd_rows_kernel(const float* __restrict__ dix, const float* __restrict__ x,
                                                          long long xbs, int rows, const float* dx_in,
                                                          long long dxin_bs, float* dx, long long dxbs, int B) {
    pdl_enter();
    __shared__ float X[FM_ROWS_SMEM * 16];
    __shared__ float S[16];
    const int e = threadIdx.x & 15, rg = threadIdx.x >> 4;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* xb = x + (long long)b * xbs;
        for (int i = threadIdx.x; i < rows * 16; i += 256) X[i] = xb[i];
        __syncthreads();
        if (rg == 0) {
            float s = 0.f;
            for (int r = 0; r < rows; ++r) s += X[r * 16 + e];
            S[e] = s;
        }
        __syncthreads();
        const float s = S[e];
        const float g2 = 2.f * dix[b * 16 + e];
        float* op = dx + (long long)b * dxbs + e;
        const float* ip = dx_in ? dx_in + (long long)b * dxin_bs + e : nullptr;
        for (int r = rg; r < rows; r += 16) {
            const float v = g2 * (s - X[r * 16 + e]);
            op[r * 16] = ip ? ip[r * 16] + v : v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ gating / concat
struct SegPack {
    int nseg;
    int pad_;
    nasrec_seg_t a[NASREC_MAX_SEGS];     // read operand
    nasrec_seg_t d[NASREC_MAX_SEGS];     // write operand (gate bwd only)
    int koff[NASREC_MAX_SEGS];
};

__global__ void gate_fwd_kernel(const float* __restrict__ pre, long long ldp, const __grid_constant__ SegPack sp,
                                float* __restrict__ out, long long ldo, int M) {
    pdl_enter();
    const nasrec_seg_t& s = sp.a[blockIdx.y];
    const int w = (int)s.width, ko = sp.koff[blockIdx.y];
    const long long total = (long long)M * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / w), k = (int)(i % w);
        const float z = pre[(long long)m * ldp + ko + k];
        const float g = 1.f / (1.f + expf(-z));
        out[(long long)m * ldo + ko + k] = g * s.ptr[(long long)m * s.ld + k];
    }
}

__global__ void gate_bwd_kernel(const float* __restrict__ dout, long long ldo, const float* __restrict__ pre,
                                long long ldp, const __grid_constant__ SegPack sp, float* __restrict__ dpre,
                                long long lddp, int M, int accumulate) {
    pdl_enter();
    const nasrec_seg_t& s = sp.a[blockIdx.y];
    const nasrec_seg_t& ds = sp.d[blockIdx.y];
    const int w = (int)s.width, ko = sp.koff[blockIdx.y];
    const long long total = (long long)M * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / w), k = (int)(i % w);
        const float z = pre[(long long)m * ldp + ko + k];
        const float g = 1.f / (1.f + expf(-z));
        const float d = dout[(long long)m * ldo + ko + k];
        const float r = s.ptr[(long long)m * s.ld + k];
        dpre[(long long)m * lddp + ko + k] = d * r * g * (1.f - g);
        if (ds.ptr) {
            float* o = const_cast<float*>(ds.ptr) + (long long)m * ds.ld + k;
            *o = accumulate ? *o + d * g : d * g;
        }
    }
}

__global__ void concat_kernel(const __grid_constant__ SegPack sp, float* __restrict__ out, long long ldo, int M,
                              int accumulate) {
    pdl_enter();
    const nasrec_seg_t& s = sp.a[blockIdx.y];
    const int w = (int)s.width;
    const long long total = (long long)M * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / w), k = (int)(i % w);
        float* o = out + (long long)m * ldo + s.w_off + k;
        const float v = s.ptr[(long long)m * s.ld + k];
        *o = accumulate ? *o + v : v;
    }
}

int fill_pack(SegPack& sp, const nasrec_seg_t* a, const nasrec_seg_t* d, int nseg, long long& maxw) {
    if (!a || nseg <= 0 || nseg > NASREC_MAX_SEGS) return NASREC_EINVAL;
    sp.nseg = nseg;
    int ko = 0;
    maxw = 0;
    for (int s = 0; s < nseg; ++s) {
        if (!a[s].ptr || a[s].width < 0) return NASREC_EINVAL;
        sp.a[s] = a[s];
        if (d) sp.d[s] = d[s];
        else sp.d[s] = nasrec_seg_t{nullptr, 0, 0, 0};
        sp.koff[s] = ko;
        ko += (int)a[s].width;
        if (a[s].width > maxw) maxw = a[s].width;
    }
    return 0;
}

int seg_grid(long long total) {
    long long g = (total + 255) / 256;
    if (g > 1024) g = 1024;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

extern "C" {

int nasrec_dot_tril_fwd(const float* x, int64_t ldx, const float* y, int64_t y_bstride, int P, float* R,
                        int64_t ldr, int B, void* stream) {
    CHECK_ARG(x && y && R && B > 0 && P > 0);
    if (P + 1 > TMAX) return NASREC_ETOOBIG;
    const int grid = B < 148 * 8 ? B : 148 * 8;
    nasrec_launch(dot_tril_fwd_kernel, grid, 256, 0, as_stream(stream), x, ldx, y, y_bstride, P, R, ldr, B);
    return nasrec_launch_status();
}

int nasrec_dot_tril_bwd(const float* dR, int64_t ldr, const float* x, int64_t ldx, const float* y,
                        int64_t y_bstride, int P, float* dx, int64_t lddx, float* dy, int64_t dy_bstride, int B,
                        void* stream) {
    CHECK_ARG(dR && x && y && B > 0 && P > 0);
    if (P + 1 > TMAX) return NASREC_ETOOBIG;
    const int grid = B < 148 * 8 ? B : 148 * 8;
    nasrec_launch(dot_tril_bwd_kernel, grid, 256, 0, as_stream(stream), dR, ldr, x, ldx, y, y_bstride, P, dx, lddx, dy,
                                                             dy_bstride, B);
    return nasrec_launch_status();
}

int nasrec_fm_fwd(const float* x, int64_t x_bstride, int rows, float* ix, int B, void* stream) {
    CHECK_ARG(x && ix && B > 0 && rows > 0);
    nasrec_launch(fm_fwd_kernel, cdiv((long long)B * 16, 256), 256, 0, as_stream(stream), x, x_bstride, rows, ix, B);
    return nasrec_launch_status();
}

int nasrec_fm_bwd(const float* dix, const float* x, int64_t x_bstride, int rows, const float* dx_in,
                  int64_t dxin_bstride, float* dx, int64_t dx_bstride, int B, void* stream) {
    CHECK_ARG(dix && x && dx && B > 0 && rows > 0);
    static const int rows_maxb = [] {                 // NASREC_FM_ROWS_MAXB=0: always the thread-per-column kernel
        const char* e = getenv("NASREC_FM_ROWS_MAXB");
        return e ? atoi(e) : 2048;
    }();
    if (B <= rows_maxb && rows <= FM_ROWS_SMEM)
        nasrec_launch(fm_bwd_rows_kernel, B < 148 * 8 ? B : 148 * 8, 256, 0, as_stream(stream), dix, x, x_bstride, rows,
                      dx_in, dxin_bstride, dx, dx_bstride, B);
    else
        nasrec_launch(fm_bwd_kernel, cdiv((long long)B * 16, 256), 256, 0, as_stream(stream), dix, x, x_bstride, rows, dx_in,
                                                                              dxin_bstride, dx, dx_bstride, B);
    return nasrec_launch_status();
}

int nasrec_gate_fwd(const float* pre, int64_t ldp, const nasrec_seg_t* right, int nseg, float* out, int64_t ldo,
                    int M, void* stream) {
    CHECK_ARG(pre && out && M > 0);
    SegPack sp{};
    long long maxw;
    int rc = fill_pack(sp, right, nullptr, nseg, maxw);
    if (rc) return rc;
    if (maxw == 0) return 0;
    dim3 grid(seg_grid((long long)M * maxw), nseg);
    nasrec_launch(gate_fwd_kernel, grid, 256, 0, as_stream(stream), pre, ldp, sp, out, ldo, M);
    return nasrec_launch_status();
}

int nasrec_gate_bwd(const float* dout, int64_t ldo, const float* pre, int64_t ldp, const nasrec_seg_t* right,
                    const nasrec_seg_t* dright, int nseg, float* dpre, int64_t lddp, int M, int accumulate,
                    void* stream) {
    CHECK_ARG(dout && pre && dpre && M > 0);
    SegPack sp{};
    long long maxw;
    int rc = fill_pack(sp, right, dright, nseg, maxw);
    if (rc) return rc;
    if (maxw == 0) return 0;
    dim3 grid(seg_grid((long long)M * maxw), nseg);
    nasrec_launch(gate_bwd_kernel, grid, 256, 0, as_stream(stream), dout, ldo, pre, ldp, sp, dpre, lddp, M, accumulate);
    return nasrec_launch_status();
}

int nasrec_concat_segs(const nasrec_seg_t* segs, int nseg, float* out, int64_t ldo, int M, int accumulate,
                       void* stream) {
    CHECK_ARG(out && M > 0);
    SegPack sp{};
    long long maxw;
    int rc = fill_pack(sp, segs, nullptr, nseg, maxw);
    if (rc) return rc;
    if (maxw == 0) return 0;
    dim3 grid(seg_grid((long long)M * maxw), nseg);
    nasrec_launch(concat_kernel, grid, 256, 0, as_stream(stream), sp, out, ldo, M, accumulate);
    return nasrec_launch_status();
}

}  // extern "C"
