// LayerNorm (+ReLU) (+prefix mask) epilogues and their backward kernels.
//
// Replaces nn.LayerNorm + activation + `torch.multiply(out, mask)` in
// nasrec/supernet/modules.py:171-181 (FC), :224-230 (EFC), :341,:360,:392-400
// (DotProduct), :492-499 (Sum), :588-593 (SigmoidGating), :649-662 (Transformer
// projection), :741-749 (FM) and supernet.py:1141 (dense->sparse merger).
// The statistics always span the full width N (masked columns included, as in
// the reference: LayerNorm runs before the mask); only the d_out live columns
// are stored.
#include "common.cuh"

namespace {

constexpr int LN_MAXN = 1024;
constexpr int LN_VPT = LN_MAXN / 32;   // values per lane

// ------------------------------------------------------------------ row LN
__global__ void __launch_bounds__(128) ln_fwd_kernel(const float* __restrict__ x, long long ldx, int M, int N,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int relu,
                                                     int d_out, float* __restrict__ y, long long ldy,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int accumulate) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + (long long)row * ldx;
    float v[LN_VPT];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k) {
        const int j = lane + 32 * k;
        v[k] = j < N ? xr[j] : 0.f;
        s += v[k];
    }
    const float mean = warp_sum(s) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k) {
        const int j = lane + 32 * k;
        const float d = j < N ? v[k] - mean : 0.f;
        q += d * d;
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)N + eps);
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
    float* yr = y + (long long)row * ldy;
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k) {
        const int j = lane + 32 * k;
        if (j < d_out) {
            float o = (v[k] - mean) * rstd * gamma[j] + beta[j];
            if (relu) o = fmaxf(o, 0.f);
            yr[j] = accumulate ? yr[j] + o : o;
        }
    }
}

__global__ void __launch_bounds__(128) ln_bwd_kernel(const float* __restrict__ dy, long long lddy, int d_out,
                                                     const float* __restrict__ x, long long ldx, int M, int N,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta,
                                                     const float* __restrict__ mean_in,
                                                     const float* __restrict__ rstd_in, int relu,
                                                     float* __restrict__ dx, long long lddx) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float* xr = x + (long long)row * ldx;
    const float* dyr = dy + (long long)row * lddy;
    float xh[LN_VPT], a[LN_VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k) {
        const int j = lane + 32 * k;
        xh[k] = 0.f;
        a[k] = 0.f;
        if (j < N) {
            xh[k] = (xr[j] - mean) * rstd;
            if (j < d_out) {
                const float g = gamma[j];
                float d = dyr[j];
                if (relu && (xh[k] * g + beta[j]) <= 0.f) d = 0.f;
                a[k] = d * g;
            }
        }
        s1 += a[k];
        s2 += a[k] * xh[k];
    }
    const float c1 = warp_sum(s1) / (float)N;
    const float c2 = warp_sum(s2) / (float)N;
    float* dxr = dx + (long long)row * lddx;
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k) {
        const int j = lane + 32 * k;
        if (j < N) dxr[j] = rstd * (a[k] - c1 - xh[k] * c2);
    }
}

// dgamma[j] = sum_m g[m,j]*xhat[m,j], dbeta[j] = sum_m g[m,j]; one CTA per 32 columns,
// 32x32 threads, fixed-order reduction over the row lanes.
__global__ void __launch_bounds__(1024) ln_param_grad_kernel(const float* __restrict__ dy, long long lddy,
                                                             int d_out, const float* __restrict__ x,
                                                             long long ldx, int M, int N,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ mean_in,
                                                             const float* __restrict__ rstd_in, int relu,
                                                             float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int accumulate) {
    __shared__ float sg[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    float ag = 0.f, ab = 0.f;
    if (j < d_out) {
        const float g = gamma[j], b = beta[j];
        for (int m = ty; m < M; m += 32) {
            const float xh = (x[(long long)m * ldx + j] - mean_in[m]) * rstd_in[m];
            float d = dy[(long long)m * lddy + j];
            if (relu && (xh * g + b) <= 0.f) d = 0.f;
            ag += d * xh;
            ab += d;
        }
    }
    sg[ty][tx] = ag;
    sb[ty][tx] = ab;
    __syncthreads();
    if (ty == 0 && j < N) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            tg += sg[r][tx];
            tb += sb[r][tx];
        }
        dgamma[j] = accumulate ? dgamma[j] + tg : tg;
        dbeta[j] = accumulate ? dbeta[j] + tb : tb;
    }
}

// ------------------------------------------------------------------ LN over P of [B,P,16]
__global__ void __launch_bounds__(256) ln3_fwd_kernel(const float* __restrict__ z, long long zbs, int B, int P,
                                                      const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps, int relu,
                                                      int p_out, float* __restrict__ y, long long ybs,
                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                      int accumulate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 16) return;
    const int b = t >> 4, e = t & 15;
    const float* zp = z + (long long)b * zbs + e;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += zp[p * 16];
    const float mean = s / (float)P;
    float q = 0.f;
    for (int p = 0; p < P; ++p) {
        const float d = zp[p * 16] - mean;
        q += d * d;
    }
    const float rstd = 1.0f / sqrtf(q / (float)P + eps);
    mean_out[t] = mean;
    rstd_out[t] = rstd;
    float* yp = y + (long long)b * ybs + e;
    for (int p = 0; p < p_out; ++p) {
        float o = (zp[p * 16] - mean) * rstd * gamma[p] + beta[p];
        if (relu) o = fmaxf(o, 0.f);
        yp[p * 16] = accumulate ? yp[p * 16] + o : o;
    }
}

__global__ void __launch_bounds__(256) ln3_bwd_kernel(const float* __restrict__ dy, long long dybs, int p_out,
                                                      const float* __restrict__ z, long long zbs, int B, int P,
                                                      const float* __restrict__ gamma,
                                                      const float* __restrict__ beta,
                                                      const float* __restrict__ mean_in,
                                                      const float* __restrict__ rstd_in, int relu,
                                                      float* __restrict__ dz, long long dzbs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 16) return;
    const int b = t >> 4, e = t & 15;
    const float mean = mean_in[t], rstd = rstd_in[t];
    const float* zp = z + (long long)b * zbs + e;
    const float* dp = dy + (long long)b * dybs + e;
    float s1 = 0.f, s2 = 0.f;
    for (int p = 0; p < p_out; ++p) {
        const float xh = (zp[p * 16] - mean) * rstd;
        const float g = gamma[p];
        float d = dp[p * 16];
        if (relu && (xh * g + beta[p]) <= 0.f) d = 0.f;
        const float a = d * g;
        s1 += a;
        s2 += a * xh;
    }
    const float c1 = s1 / (float)P, c2 = s2 / (float)P;
    float* op = dz + (long long)b * dzbs + e;
    for (int p = 0; p < P; ++p) {
        const float xh = (zp[p * 16] - mean) * rstd;
        float a = 0.f;
        if (p < p_out) {
            const float g = gamma[p];
            float d = dp[p * 16];
            if (relu && (xh * g + beta[p]) <= 0.f) d = 0.f;
            a = d * g;
        }
        op[p * 16] = rstd * (a - c1 - xh * c2);
    }
}

// one CTA per p: dgamma[p] = sum_{b,e} g*xhat, dbeta[p] = sum g
__global__ void __launch_bounds__(256) ln3_param_grad_kernel(const float* __restrict__ dy, long long dybs,
                                                             int p_out, const float* __restrict__ z,
                                                             long long zbs, int B, int P,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ mean_in,
                                                             const float* __restrict__ rstd_in, int relu,
                                                             float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int accumulate) {
    __shared__ float red[34];
    const int p = blockIdx.x;
    float ag = 0.f, ab = 0.f;
    if (p < p_out) {
        const float g = gamma[p], bt = beta[p];
        for (int t = threadIdx.x; t < B * 16; t += blockDim.x) {
            const int b = t >> 4, e = t & 15;
            const float xh = (z[(long long)b * zbs + p * 16 + e] - mean_in[t]) * rstd_in[t];
            float d = dy[(long long)b * dybs + p * 16 + e];
            if (relu && (xh * g + bt) <= 0.f) d = 0.f;
            ag += d * xh;
            ab += d;
        }
    }
    const float tg = block_sum(ag, red);
    const float tb = block_sum(ab, red);
    if (threadIdx.x == 0) {
        dgamma[p] = accumulate ? dgamma[p] + tg : tg;
        dbeta[p] = accumulate ? dbeta[p] + tb : tb;
    }
}

// db[p] = sum_{b,e} dz[b,p,e]  (bias of a sparse-axis projection; use_layernorm=False models)
__global__ void __launch_bounds__(256) sproj_bias_grad_kernel(const float* __restrict__ dz, long long dzbs, int B,
                                                              float* __restrict__ db, int accumulate) {
    __shared__ float red[34];
    const int p = blockIdx.x;
    float a = 0.f;
    for (int t = threadIdx.x; t < B * 16; t += blockDim.x) a += dz[(long long)(t >> 4) * dzbs + p * 16 + (t & 15)];
    const float tot = block_sum(a, red);
    if (threadIdx.x == 0) db[p] = accumulate ? db[p] + tot : tot;
}

// ------------------------------------------------------------------ plain activation
__global__ void act_fwd_kernel(const float* __restrict__ x, long long ldx, int M, int N, int relu,
                               float* __restrict__ y, long long ldy, int accumulate) {
    const long long total = (long long)M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i % N);
        float v = x[(long long)m * ldx + n];
        if (relu) v = fmaxf(v, 0.f);
        float* o = y + (long long)m * ldy + n;
        *o = accumulate ? *o + v : v;
    }
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x,
                               long long ldx, int M, int N, int relu, float* __restrict__ dx, long long lddx) {
    const long long total = (long long)M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i % N);
        float d = dy[(long long)m * lddy + n];
        if (relu && x[(long long)m * ldx + n] <= 0.f) d = 0.f;
        dx[(long long)m * lddx + n] = d;
    }
}

__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ x, long long ld, int M, int N,
                                                      float* __restrict__ out, int accumulate) {
    __shared__ float sm[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (j < N)
        for (int m = ty; m < M; m += 32) a += x[(long long)m * ld + j];
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) t += sm[r][tx];
        out[j] = accumulate ? out[j] + t : t;
    }
}

int ew_grid(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

extern "C" {

int nasrec_ln_fwd(const float* x, int64_t ldx, int M, int N, const float* gamma, const float* beta, float eps,
                  int relu, int d_out, float* y, int64_t ldy, float* mean, float* rstd, int accumulate,
                  void* stream) {
    CHECK_ARG(x && gamma && beta && y && mean && rstd && M > 0 && N > 0 && d_out >= 0 && d_out <= N);
    if (N > LN_MAXN) return NASREC_ETOOBIG;
    ln_fwd_kernel<<<cdiv(M, 4), 128, 0, as_stream(stream)>>>(x, ldx, M, N, gamma, beta, eps, relu, d_out, y, ldy,
                                                             mean, rstd, accumulate);
    return nasrec_launch_status();
}

int nasrec_ln_bwd(const float* dy, int64_t lddy, int d_out, const float* x, int64_t ldx, int M, int N,
                  const float* gamma, const float* beta, const float* mean, const float* rstd, int relu,
                  float* dx, int64_t lddx, float* dgamma, float* dbeta, int accumulate_params, void* stream) {
    CHECK_ARG(dy && x && gamma && beta && mean && rstd && M > 0 && N > 0 && d_out >= 0 && d_out <= N);
    if (N > LN_MAXN) return NASREC_ETOOBIG;
    cudaStream_t st = as_stream(stream);
    if (dx) {
        ln_bwd_kernel<<<cdiv(M, 4), 128, 0, st>>>(dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd, relu, dx,
                                                  lddx);
        int rc = nasrec_launch_status();
        if (rc) return rc;
    }
    if (dgamma && dbeta) {
        ln_param_grad_kernel<<<cdiv(N, 32), 1024, 0, st>>>(dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd,
                                                           relu, dgamma, dbeta, accumulate_params);
        return nasrec_launch_status();
    }
    return 0;
}

int nasrec_ln3_fwd(const float* z, int64_t z_bstride, int B, int P, const float* gamma, const float* beta,
                   float eps, int relu, int p_out, float* y, int64_t y_bstride, float* mean, float* rstd,
                   int accumulate, void* stream) {
    CHECK_ARG(z && gamma && beta && y && mean && rstd && B > 0 && P > 0 && p_out >= 0 && p_out <= P);
    ln3_fwd_kernel<<<cdiv((long long)B * 16, 256), 256, 0, as_stream(stream)>>>(
        z, z_bstride, B, P, gamma, beta, eps, relu, p_out, y, y_bstride, mean, rstd, accumulate);
    return nasrec_launch_status();
}

int nasrec_ln3_bwd(const float* dy, int64_t dy_bstride, int p_out, const float* z, int64_t z_bstride, int B,
                   int P, const float* gamma, const float* beta, const float* mean, const float* rstd, int relu,
                   float* dz, int64_t dz_bstride, float* dgamma, float* dbeta, int accumulate_params,
                   void* stream) {
    CHECK_ARG(dy && z && gamma && beta && mean && rstd && B > 0 && P > 0 && p_out >= 0 && p_out <= P);
    cudaStream_t st = as_stream(stream);
    if (dz) {
        ln3_bwd_kernel<<<cdiv((long long)B * 16, 256), 256, 0, st>>>(dy, dy_bstride, p_out, z, z_bstride, B, P,
                                                                     gamma, beta, mean, rstd, relu, dz, dz_bstride);
        int rc = nasrec_launch_status();
        if (rc) return rc;
    }
    if (dgamma && dbeta) {
        ln3_param_grad_kernel<<<P, 256, 0, st>>>(dy, dy_bstride, p_out, z, z_bstride, B, P, gamma, beta, mean, rstd,
                                                 relu, dgamma, dbeta, accumulate_params);
        return nasrec_launch_status();
    }
    return 0;
}

int nasrec_sproj_bias_grad(const float* dZ, int64_t dz_bstride, int P, int B, float* db, int accumulate,
                           void* stream) {
    CHECK_ARG(dZ && db && P > 0 && B > 0);
    sproj_bias_grad_kernel<<<P, 256, 0, as_stream(stream)>>>(dZ, dz_bstride, B, db, accumulate);
    return nasrec_launch_status();
}

int nasrec_act_fwd(const float* x, int64_t ldx, int M, int N, int relu, float* y, int64_t ldy, int accumulate,
                   void* stream) {
    CHECK_ARG(x && y && M > 0 && N > 0);
    act_fwd_kernel<<<ew_grid((long long)M * N), 256, 0, as_stream(stream)>>>(x, ldx, M, N, relu, y, ldy,
                                                                            accumulate);
    return nasrec_launch_status();
}

int nasrec_act_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, int M, int N, int relu, float* dx,
                   int64_t lddx, void* stream) {
    CHECK_ARG(dy && x && dx && M > 0 && N > 0);
    act_bwd_kernel<<<ew_grid((long long)M * N), 256, 0, as_stream(stream)>>>(dy, lddy, x, ldx, M, N, relu, dx,
                                                                            lddx);
    return nasrec_launch_status();
}

int nasrec_colsum(const float* x, int64_t ld, int M, int N, float* out, int accumulate, void* stream) {
    CHECK_ARG(x && out && M > 0 && N > 0);
    colsum_kernel<<<cdiv(N, 32), 1024, 0, as_stream(stream)>>>(x, ld, M, N, out, accumulate);
    return nasrec_launch_status();
}

}  // extern "C"
