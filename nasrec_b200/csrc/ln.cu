// LayerNorm (+ReLU) (+prefix mask) epilogues and their backward kernels.
//
// Replaces nn.LayerNorm + activation + `torch.multiply(out, mask)` in
// nasrec/supernet/modules.py:171-181 (FC), :224-230 (EFC), :341,:360,:392-400
// (DotProduct), :492-499 (Sum), :588-593 (SigmoidGating), :649-662 (Transformer
// projection), :741-749 (FM) and supernet.py:1141 (dense->sparse merger).
// The statistics always span the full width N (masked columns included, as in
// the reference: LayerNorm runs before the mask); only the d_out live columns
// are stored.
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int LN_MAXN = 1024;
constexpr int LN_VPT = LN_MAXN / 32;   // values per lane

// ------------------------------------------------------------------ row LN
// VPT = values per lane: 32 lanes x VPT >= N.  The narrow instances (N <= 64 / 128 / 256) exist because a fixed VPT of 32 makes an
// N = 64 row pay for 32 predicated-off loop bodies per lane where 2 do the work.
template <int VPT>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const float* __restrict__ x, long long ldx, int M, int N,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int relu,
                                                     int d_out, float* __restrict__ y, long long ldy,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int accumulate) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + (long long)row * ldx;
    float* yr = y + (long long)row * ldy;
    // phase 1: every load of the row is issued before anything waits (see ldg_nc_pred)
    float v[VPT], gm[VPT], bt[VPT], yo[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = lane + 32 * k;
        v[k] = ldg_nc_pred(xr + j, j < N);
        gm[k] = ldg_nc_pred(gamma + j, j < d_out);
        bt[k] = ldg_nc_pred(beta + j, j < d_out);
        yo[k] = ld_pred(yr + j, accumulate && j < d_out);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) s += v[k];
    const float mean = warp_sum(s) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = lane + 32 * k;
        const float d = j < N ? v[k] - mean : 0.f;
        q += d * d;
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)N + eps);
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = lane + 32 * k;
        float o = (v[k] - mean) * rstd * gm[k] + bt[k];
        if (relu) o = fmaxf(o, 0.f);
        if (j < d_out) yr[j] = accumulate ? yo[k] + o : o;
    }
}

// dx for every row (persistent warps, fixed row -> warp assignment); when `part` is given, also this
// CTA's partial column sums of dgamma/dbeta (part[cta][0][j], part[cta][1][j]) for the second stage.
template <int VPT>
__global__ void __launch_bounds__(128) ln_bwd_kernel(const float* __restrict__ dy, long long lddy, int d_out,
                                                     const float* __restrict__ x, long long ldx, int M, int N,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta,
                                                     const float* __restrict__ mean_in,
                                                     const float* __restrict__ rstd_in, int relu,
                                                     float* __restrict__ dx, long long lddx,
                                                     float* __restrict__ part) {
    pdl_enter();
    extern __shared__ float sred[];          // [2][4][N] when part != null
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float accg[VPT], accb[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        accg[k] = 0.f;
        accb[k] = 0.f;
    }
    for (int row = blockIdx.x * 4 + w; row < M; row += gridDim.x * 4) {
        const float* xr = x + (long long)row * ldx;
        const float* dyr = dy + (long long)row * lddy;
        // phase 1: all loads in flight together
        const float mean = mean_in[row], rstd = rstd_in[row];
        float xh[VPT], a[VPT], gm[VPT], bt[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = lane + 32 * k;
            xh[k] = ldg_nc_pred(xr + j, j < N);
            a[k] = ldg_nc_pred(dyr + j, j < d_out);
            gm[k] = ldg_nc_pred(gamma + j, j < d_out);
            bt[k] = ldg_nc_pred(beta + j, j < d_out);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = lane + 32 * k;
            xh[k] = j < N ? (xh[k] - mean) * rstd : 0.f;
            float d = a[k];                                      // 0 beyond d_out
            if (relu && (xh[k] * gm[k] + bt[k]) <= 0.f) d = 0.f;
            a[k] = d * gm[k];
            accg[k] = fmaf(d, xh[k], accg[k]);
            accb[k] += d;
            s1 += a[k];
            s2 += a[k] * xh[k];
        }
        const float c1 = warp_sum(s1) / (float)N;
        const float c2 = warp_sum(s2) / (float)N;
        if (dx) {
            float* dxr = dx + (long long)row * lddx;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = lane + 32 * k;
                if (j < N) dxr[j] = rstd * (a[k] - c1 - xh[k] * c2);
            }
        }
    }
    if (!part) return;
    float* sg = sred;
    float* sb = sred + 4 * N;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = lane + 32 * k;
        if (j < N) {
            sg[w * N + j] = accg[k];
            sb[w * N + j] = accb[k];
        }
    }
    __syncthreads();
    float* o = part + (long long)blockIdx.x * 2 * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        o[j] = (sg[j] + sg[N + j]) + (sg[2 * N + j] + sg[3 * N + j]);
        o[N + j] = (sb[j] + sb[N + j]) + (sb[2 * N + j] + sb[3 * N + j]);
    }
}

// ------------------------------------------------------------------ wide rows: four warps per row
// With one warp per row a [512, 1024] LayerNorm is 512 warps of ~1500 dependent instructions each -- 4 resident warps
// per SM, ~6 clk per instruction, ~10 us (ncu, warm caches: warps active 6 %, issue active 14 %, profiles/r02_notes.md).
// For N > 256 a row is split over LNW_WPR = 4 warps (contiguous column quarters, 8 values per lane) and a CTA of 8 warps
// works on two rows at a time; row sums cross the warps through shared memory and are added in a fixed order.
constexpr int LNW_WPR = 4;                       // warps per row
constexpr int LNW_RPC = 8 / LNW_WPR;             // rows per CTA and iteration
constexpr int LNW_VPT = LN_VPT / LNW_WPR;        // values per lane

__device__ __forceinline__ float lnw_row_sum(float v, float (*red)[LNW_WPR], int g, int q, int lane) {
    v = warp_sum(v);
    __syncthreads();                             // the previous use of `red` is over
    if (lane == 0) red[g][q] = v;
    __syncthreads();
    return (red[g][0] + red[g][1]) + (red[g][2] + red[g][3]);
}

__global__ void __launch_bounds__(256) ln_fwd_wide_kernel(const float* __restrict__ x, long long ldx, int M, int N,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, int relu,
                                                          int d_out, float* __restrict__ y, long long ldy,
                                                          float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                          int accumulate) {
    pdl_enter();
    __shared__ float red[LNW_RPC][LNW_WPR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp / LNW_WPR, q = warp % LNW_WPR;
    const int row = blockIdx.x * LNW_RPC + g;
    const bool live = row < M;
    const float* xr = x + (long long)(live ? row : 0) * ldx;
    float* yr = y + (long long)(live ? row : 0) * ldy;
    const int j0 = q * (32 * LNW_VPT) + lane;
    float v[LNW_VPT], gm[LNW_VPT], bt[LNW_VPT], yo[LNW_VPT];
#pragma unroll
    for (int k = 0; k < LNW_VPT; ++k) {
        const int j = j0 + 32 * k;
        v[k] = ldg_nc_pred(xr + j, live && j < N);
        gm[k] = ldg_nc_pred(gamma + j, j < d_out);
        bt[k] = ldg_nc_pred(beta + j, j < d_out);
        yo[k] = ld_pred(yr + j, live && accumulate && j < d_out);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LNW_VPT; ++k) s += v[k];
    const float mean = lnw_row_sum(s, red, g, q, lane) / (float)N;
    float qq = 0.f;
#pragma unroll
    for (int k = 0; k < LNW_VPT; ++k) {
        const int j = j0 + 32 * k;
        const float d = j < N ? v[k] - mean : 0.f;
        qq += d * d;
    }
    const float rstd = 1.0f / sqrtf(lnw_row_sum(qq, red, g, q, lane) / (float)N + eps);
    if (!live) return;
    if (lane == 0 && q == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < LNW_VPT; ++k) {
        const int j = j0 + 32 * k;
        float o = (v[k] - mean) * rstd * gm[k] + bt[k];
        if (relu) o = fmaxf(o, 0.f);
        if (j < d_out) yr[j] = accumulate ? yo[k] + o : o;
    }
}

// dx for every row; with `part`, this CTA's partial column sums of dgamma / dbeta (part[cta][0|1][j]) for the second stage
__global__ void __launch_bounds__(256) ln_bwd_wide_kernel(const float* __restrict__ dy, long long lddy, int d_out,
                                                          const float* __restrict__ x, long long ldx, int M, int N,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ mean_in,
                                                          const float* __restrict__ rstd_in, int relu,
                                                          float* __restrict__ dx, long long lddx,
                                                          float* __restrict__ part) {
    pdl_enter();
    extern __shared__ float sred[];          // [2][LNW_RPC][N] when part != null
    __shared__ float red[LNW_RPC][LNW_WPR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp / LNW_WPR, q = warp % LNW_WPR;
    const int j0 = q * (32 * LNW_VPT) + lane;
    float accg[LNW_VPT], accb[LNW_VPT], gm[LNW_VPT], bt[LNW_VPT];
#pragma unroll
    for (int k = 0; k < LNW_VPT; ++k) {
        accg[k] = 0.f;
        accb[k] = 0.f;
        gm[k] = ldg_nc_pred(gamma + j0 + 32 * k, j0 + 32 * k < d_out);
        bt[k] = ldg_nc_pred(beta + j0 + 32 * k, j0 + 32 * k < d_out);
    }
    for (int r0 = blockIdx.x * LNW_RPC; r0 < M; r0 += gridDim.x * LNW_RPC) {       // uniform trip count per CTA
        const int row = r0 + g;
        const bool live = row < M;
        const float* xr = x + (long long)(live ? row : 0) * ldx;
        const float* dyr = dy + (long long)(live ? row : 0) * lddy;
        const float mean = live ? mean_in[row] : 0.f, rstd = live ? rstd_in[row] : 0.f;
        float xh[LNW_VPT], a[LNW_VPT];
#pragma unroll
        for (int k = 0; k < LNW_VPT; ++k) {
            const int j = j0 + 32 * k;
            xh[k] = ldg_nc_pred(xr + j, live && j < N);
            a[k] = ldg_nc_pred(dyr + j, live && j < d_out);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < LNW_VPT; ++k) {
            const int j = j0 + 32 * k;
            xh[k] = (live && j < N) ? (xh[k] - mean) * rstd : 0.f;
            float d = a[k];                                      // 0 beyond d_out
            if (relu && (xh[k] * gm[k] + bt[k]) <= 0.f) d = 0.f;
            a[k] = d * gm[k];
            accg[k] = fmaf(d, xh[k], accg[k]);
            accb[k] += d;
            s1 += a[k];
            s2 += a[k] * xh[k];
        }
        const float c1 = lnw_row_sum(s1, red, g, q, lane) / (float)N;
        const float c2 = lnw_row_sum(s2, red, g, q, lane) / (float)N;
        if (dx && live) {
            float* dxr = dx + (long long)row * lddx;
#pragma unroll
            for (int k = 0; k < LNW_VPT; ++k) {
                const int j = j0 + 32 * k;
                if (j < N) dxr[j] = rstd * (a[k] - c1 - xh[k] * c2);
            }
        }
    }
    if (!part) return;
    float* sg = sred;
    float* sb = sred + LNW_RPC * N;
#pragma unroll
    for (int k = 0; k < LNW_VPT; ++k) {
        const int j = j0 + 32 * k;
        if (j < N) {
            sg[g * N + j] = accg[k];
            sb[g * N + j] = accb[k];
        }
    }
    __syncthreads();
    float* o = part + (long long)blockIdx.x * 2 * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        o[j] = sg[j] + sg[N + j];
        o[N + j] = sb[j] + sb[N + j];
    }
}

// Second stage of the dgamma/dbeta reduction: part[k][0|1][j], k < nblk.  One CTA per 32 columns; warp r
// sums partial rows r, r+8, ... (coalesced 128-byte reads, 4 independent loads in flight), then the 8
// warp totals are added in a fixed order -- deterministic for a given nblk.
__global__ void __launch_bounds__(256) ln_param_final_kernel(const float* __restrict__ part, int nblk, int N,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             int accumulate) {
    pdl_enter();
    __shared__ float sg[8][33], sb[8][33];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + c;
    float g = 0.f, b = 0.f;
    if (j < N) {
        int k = r;
        for (; k + 24 < nblk; k += 32) {
            const float* p0 = part + (long long)k * 2 * N + j;
            const float g0 = p0[0], b0 = p0[N];
            const float g1 = p0[(long long)16 * N], b1 = p0[(long long)16 * N + N];
            const float g2 = p0[(long long)32 * N], b2 = p0[(long long)32 * N + N];
            const float g3 = p0[(long long)48 * N], b3 = p0[(long long)48 * N + N];
            g += (g0 + g1) + (g2 + g3);
            b += (b0 + b1) + (b2 + b3);
        }
        for (; k < nblk; k += 8) {
            g += part[(long long)k * 2 * N + j];
            b += part[(long long)k * 2 * N + N + j];
        }
    }
    sg[r][c] = g;
    sb[r][c] = b;
    __syncthreads();
    if (r == 0 && j < N) {
        const float gt = ((sg[0][c] + sg[1][c]) + (sg[2][c] + sg[3][c])) + ((sg[4][c] + sg[5][c]) + (sg[6][c] + sg[7][c]));
        const float bt = ((sb[0][c] + sb[1][c]) + (sb[2][c] + sb[3][c])) + ((sb[4][c] + sb[5][c]) + (sb[6][c] + sb[7][c]));
        dgamma[j] = accumulate ? dgamma[j] + gt : gt;
        dbeta[j] = accumulate ? dbeta[j] + bt : bt;
    }
}

__global__ void __launch_bounds__(1024) ln_param_grad_kernel(const float* __restrict__ dy, long long lddy,
                                                             int d_out, const float* __restrict__ x,
                                                             long long ldx, int M, int N,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ mean_in,
                                                             const float* __restrict__ rstd_in, int relu,
                                                             float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int accumulate) {
    pdl_enter();
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
// One warp per sample: lane = (ph, e) with e = lane & 15 the embedding column and ph = lane >> 4 the
// parity of the row it owns (rows ph, ph+2, ...), so every warp load is one contiguous 128-byte line
// (two 64-byte rows) and a column's P <= 64 values live in 32 registers of two lanes.
constexpr int LN3_MAXP = 64;
constexpr int LN3_V = LN3_MAXP / 2;

__global__ void __launch_bounds__(128) ln3_fwd_kernel(const float* __restrict__ z, long long zbs, int B, int P,
                                                      const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps, int relu,
                                                      int p_out, float* __restrict__ y, long long ybs,
                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                      int accumulate) {
    pdl_enter();
    const int lane = threadIdx.x & 31, e = lane & 15, ph = lane >> 4;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* zp = z + (long long)b * zbs + e;
    float* yp = y + (long long)b * ybs + e;
    // phase 1: every load in flight before the first use (see ldg_nc_pred)
    float v[LN3_V], gm[LN3_V], bt[LN3_V], yo[LN3_V];
#pragma unroll
    for (int k = 0; k < LN3_V; ++k) {
        const int p = 2 * k + ph;
        v[k] = ldg_nc_pred(zp + p * 16, p < P);
        gm[k] = ldg_nc_pred(gamma + p, p < p_out);
        bt[k] = ldg_nc_pred(beta + p, p < p_out);
        yo[k] = ld_pred(yp + p * 16, accumulate && p < p_out);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN3_V; ++k) s += v[k];
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    const float mean = s / (float)P;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN3_V; ++k) {
        const int p = 2 * k + ph;
        const float d = p < P ? v[k] - mean : 0.f;
        q += d * d;
    }
    q += __shfl_xor_sync(0xffffffffu, q, 16);
    const float rstd = 1.0f / sqrtf(q / (float)P + eps);
    if (ph == 0) {
        mean_out[b * 16 + e] = mean;
        rstd_out[b * 16 + e] = rstd;
    }
#pragma unroll
    for (int k = 0; k < LN3_V; ++k) {
        const int p = 2 * k + ph;
        float o = (v[k] - mean) * rstd * gm[k] + bt[k];
        if (relu) o = fmaxf(o, 0.f);
        if (p < p_out) yp[p * 16] = accumulate ? yo[k] + o : o;
    }
}

// dz for every sample; when `part` is given, also this CTA's partial sums of dgamma/dbeta
// (part[cta][0][p], part[cta][1][p]) for the fixed-order second stage below.
__global__ void __launch_bounds__(128) ln3_bwd_kernel(const float* __restrict__ dy, long long dybs, int p_out,
                                                      const float* __restrict__ z, long long zbs, int B, int P,
                                                      const float* __restrict__ gamma,
                                                      const float* __restrict__ beta,
                                                      const float* __restrict__ mean_in,
                                                      const float* __restrict__ rstd_in, int relu,
                                                      float* __restrict__ dz, long long dzbs,
                                                      float* __restrict__ part) {
    pdl_enter();
    __shared__ float sg[4][LN3_MAXP], sb[4][LN3_MAXP];
    const int lane = threadIdx.x & 31, e = lane & 15, ph = lane >> 4, w = threadIdx.x >> 5;
    float accg[LN3_V], accb[LN3_V];
#pragma unroll
    for (int k = 0; k < LN3_V; ++k) {
        accg[k] = 0.f;
        accb[k] = 0.f;
    }
    for (int b = blockIdx.x * 4 + w; b < B; b += gridDim.x * 4) {
        const float* zp = z + (long long)b * zbs + e;
        const float* dp = dy + (long long)b * dybs + e;
        // phase 1: all loads in flight together
        const float mean = mean_in[b * 16 + e], rstd = rstd_in[b * 16 + e];
        float xh[LN3_V], a[LN3_V], gm[LN3_V], bt[LN3_V];
#pragma unroll
        for (int k = 0; k < LN3_V; ++k) {
            const int p = 2 * k + ph;
            xh[k] = ldg_nc_pred(zp + p * 16, p < P);
            a[k] = ldg_nc_pred(dp + p * 16, p < p_out);
            gm[k] = ldg_nc_pred(gamma + p, p < p_out);
            bt[k] = ldg_nc_pred(beta + p, p < p_out);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < LN3_V; ++k) {
            const int p = 2 * k + ph;
            xh[k] = p < P ? (xh[k] - mean) * rstd : 0.f;
            float d = a[k];                                      // 0 beyond p_out
            if (relu && (xh[k] * gm[k] + bt[k]) <= 0.f) d = 0.f;
            a[k] = d * gm[k];
            accg[k] = fmaf(d, xh[k], accg[k]);
            accb[k] += d;
            s1 += a[k];
            s2 = fmaf(a[k], xh[k], s2);
        }
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        const float c1 = s1 / (float)P, c2 = s2 / (float)P;
        if (dz) {
            float* op = dz + (long long)b * dzbs + e;
#pragma unroll
            for (int k = 0; k < LN3_V; ++k) {
                const int p = 2 * k + ph;
                if (p < P) op[p * 16] = rstd * (a[k] - c1 - xh[k] * c2);
            }
        }
    }
    if (!part) return;
    // sum over the 16 embedding columns (lanes with equal ph), then over the CTA's 4 warps
#pragma unroll
    for (int k = 0; k < LN3_V; ++k) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            accg[k] += __shfl_xor_sync(0xffffffffu, accg[k], o);
            accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], o);
        }
        if (e == 0) {
            sg[w][2 * k + ph] = accg[k];
            sb[w][2 * k + ph] = accb[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < LN3_MAXP) {
        const int p = threadIdx.x;
        float* o = part + (long long)blockIdx.x * 2 * LN3_MAXP;
        o[p] = (sg[0][p] + sg[1][p]) + (sg[2][p] + sg[3][p]);
        o[LN3_MAXP + p] = (sb[0][p] + sb[1][p]) + (sb[2][p] + sb[3][p]);
    }
}

// ---- four warps per sample (same reasoning as the wide row kernels above: one warp per sample leaves 4 resident warps
// per SM walking ~2000 dependent instructions each).  Warp (g, qw) of a CTA handles sample blockIdx * 2 + g and the rows
// p = 8 k + 2 qw + ph, k < 8; the two row sums of every (sample, e) cross the four warps through shared memory.
constexpr int LN3W_V = LN3_V / 4;
__device__ __forceinline__ float ln3w_sum(float v, float (*red)[4][16], int g, int qw, int e, int ph) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    __syncthreads();
    if (ph == 0) red[g][qw][e] = v;
    __syncthreads();
    return (red[g][0][e] + red[g][1][e]) + (red[g][2][e] + red[g][3][e]);
}

__global__ void __launch_bounds__(256) ln3_fwd_wide_kernel(const float* __restrict__ z, long long zbs, int B, int P,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int relu,
                                                           int p_out, float* __restrict__ y, long long ybs,
                                                           float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                           int accumulate) {
    pdl_enter();
    __shared__ float red[2][4][16];
    const int lane = threadIdx.x & 31, e = lane & 15, ph = lane >> 4, warp = threadIdx.x >> 5;
    const int g = warp >> 2, qw = warp & 3;
    const int b = blockIdx.x * 2 + g;
    const bool live = b < B;
    const float* zp = z + (long long)(live ? b : 0) * zbs + e;
    float* yp = y + (long long)(live ? b : 0) * ybs + e;
    float v[LN3W_V], gm[LN3W_V], bt[LN3W_V], yo[LN3W_V];
#pragma unroll
    for (int k = 0; k < LN3W_V; ++k) {
        const int p = 8 * k + 2 * qw + ph;
        v[k] = ldg_nc_pred(zp + p * 16, live && p < P);
        gm[k] = ldg_nc_pred(gamma + p, p < p_out);
        bt[k] = ldg_nc_pred(beta + p, p < p_out);
        yo[k] = ld_pred(yp + p * 16, live && accumulate && p < p_out);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN3W_V; ++k) s += v[k];
    const float mean = ln3w_sum(s, red, g, qw, e, ph) / (float)P;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN3W_V; ++k) {
        const int p = 8 * k + 2 * qw + ph;
        const float d = p < P ? v[k] - mean : 0.f;
        q += d * d;
    }
    const float rstd = 1.0f / sqrtf(ln3w_sum(q, red, g, qw, e, ph) / (float)P + eps);
    if (!live) return;
    if (ph == 0 && qw == 0) {
        mean_out[b * 16 + e] = mean;
        rstd_out[b * 16 + e] = rstd;
    }
#pragma unroll
    for (int k = 0; k < LN3W_V; ++k) {
        const int p = 8 * k + 2 * qw + ph;
        float o = (v[k] - mean) * rstd * gm[k] + bt[k];
        if (relu) o = fmaxf(o, 0.f);
        if (p < p_out) yp[p * 16] = accumulate ? yo[k] + o : o;
    }
}

__global__ void __launch_bounds__(256) ln3_bwd_wide_kernel(const float* __restrict__ dy, long long dybs, int p_out,
                                                           const float* __restrict__ z, long long zbs, int B, int P,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ mean_in,
                                                           const float* __restrict__ rstd_in, int relu,
                                                           float* __restrict__ dz, long long dzbs,
                                                           float* __restrict__ part) {
    pdl_enter();
    __shared__ float sg[2][LN3_MAXP], sb[2][LN3_MAXP];
    __shared__ float red[2][4][16];
    const int lane = threadIdx.x & 31, e = lane & 15, ph = lane >> 4, warp = threadIdx.x >> 5;
    const int g = warp >> 2, qw = warp & 3;
    float accg[LN3W_V], accb[LN3W_V], gm[LN3W_V], bt[LN3W_V];
#pragma unroll
    for (int k = 0; k < LN3W_V; ++k) {
        const int p = 8 * k + 2 * qw + ph;
        accg[k] = 0.f;
        accb[k] = 0.f;
        gm[k] = ldg_nc_pred(gamma + p, p < p_out);
        bt[k] = ldg_nc_pred(beta + p, p < p_out);
    }
    for (int b0 = blockIdx.x * 2; b0 < B; b0 += gridDim.x * 2) {              // uniform trip count per CTA
        const int b = b0 + g;
        const bool live = b < B;
        const float* zp = z + (long long)(live ? b : 0) * zbs + e;
        const float* dp = dy + (long long)(live ? b : 0) * dybs + e;
        const float mean = live ? mean_in[b * 16 + e] : 0.f, rstd = live ? rstd_in[b * 16 + e] : 0.f;
        float xh[LN3W_V], a[LN3W_V];
#pragma unroll
        for (int k = 0; k < LN3W_V; ++k) {
            const int p = 8 * k + 2 * qw + ph;
            xh[k] = ldg_nc_pred(zp + p * 16, live && p < P);
            a[k] = ldg_nc_pred(dp + p * 16, live && p < p_out);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < LN3W_V; ++k) {
            const int p = 8 * k + 2 * qw + ph;
            xh[k] = (live && p < P) ? (xh[k] - mean) * rstd : 0.f;
            float d = a[k];                                      // 0 beyond p_out
            if (relu && (xh[k] * gm[k] + bt[k]) <= 0.f) d = 0.f;
            a[k] = d * gm[k];
            accg[k] = fmaf(d, xh[k], accg[k]);
            accb[k] += d;
            s1 += a[k];
            s2 = fmaf(a[k], xh[k], s2);
        }
        const float c1 = ln3w_sum(s1, red, g, qw, e, ph) / (float)P;
        const float c2 = ln3w_sum(s2, red, g, qw, e, ph) / (float)P;
        if (dz && live) {
            float* op = dz + (long long)b * dzbs + e;
#pragma unroll
            for (int k = 0; k < LN3W_V; ++k) {
                const int p = 8 * k + 2 * qw + ph;
                if (p < P) op[p * 16] = rstd * (a[k] - c1 - xh[k] * c2);
            }
        }
    }
    if (!part) return;
    // sum over the 16 embedding columns (lanes with equal ph), then over the CTA's two sample groups
#pragma unroll
    for (int k = 0; k < LN3W_V; ++k) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            accg[k] += __shfl_xor_sync(0xffffffffu, accg[k], o);
            accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], o);
        }
        if (e == 0) {
            sg[g][8 * k + 2 * qw + ph] = accg[k];
            sb[g][8 * k + 2 * qw + ph] = accb[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < LN3_MAXP) {
        const int p = threadIdx.x;
        float* o = part + (long long)blockIdx.x * 2 * LN3_MAXP;
        o[p] = sg[0][p] + sg[1][p];
        o[LN3_MAXP + p] = sb[0][p] + sb[1][p];
    }
}

// part[k][0|1][p], k < nblk: warp r sums partial rows r, r+8, ...; fixed-order add of the 8 warp totals.
__global__ void __launch_bounds__(256) ln3_param_final_kernel(const float* __restrict__ part, int nblk, int P,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              int accumulate) {
    pdl_enter();
    __shared__ float sg[8][33], sb[8][33];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + c;
    float g = 0.f, b = 0.f;
    if (p < P) {
        for (int k = r; k < nblk; k += 8) {
            g += part[(long long)k * 2 * LN3_MAXP + p];
            b += part[(long long)k * 2 * LN3_MAXP + LN3_MAXP + p];
        }
    }
    sg[r][c] = g;
    sb[r][c] = b;
    __syncthreads();
    if (r == 0 && p < P) {
        const float gt = ((sg[0][c] + sg[1][c]) + (sg[2][c] + sg[3][c])) + ((sg[4][c] + sg[5][c]) + (sg[6][c] + sg[7][c]));
        const float bt = ((sb[0][c] + sb[1][c]) + (sb[2][c] + sb[3][c])) + ((sb[4][c] + sb[5][c]) + (sb[6][c] + sb[7][c]));
        dgamma[p] = accumulate ? dgamma[p] + gt : gt;
        dbeta[p] = accumulate ? dbeta[p] + bt : bt;
    }
}

// fallback when no workspace is attached: one CTA per p, dgamma[p] = sum_{b,e} g*xhat, dbeta[p] = sum g
__global__ void __launch_bounds__(256) ln3_param_grad_kernel(const float* __restrict__ dy, long long dybs,
                                                             int p_out, const float* __restrict__ z,
                                                             long long zbs, int B, int P,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ mean_in,
                                                             const float* __restrict__ rstd_in, int relu,
                                                             float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int accumulate) {
    pdl_enter();
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
    pdl_enter();
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
    pdl_enter();
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
    pdl_enter();
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
    pdl_enter();
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


// ---- deferred parameter gradients ---------------------------------------------------------------------------------------
// dgamma / dbeta are read only by the optimizer, but their second reduction stage used to be a launch of its own right
// behind every LayerNorm backward -- ~23 launches on the dY -> dX chain of a step.  While a scratch region is attached
// (nasrec_internal_ln_defer_scratch: the step executor hands over a piece of its arena for the duration of a backward
// pass), every LayerNorm backward keeps its first-stage partials in a slot of its own and only records what is left to
// do; nasrec_internal_ln_flush (called by nasrec_wgrad_flush, i.e. once per block, on the side stream when there is one)
// finishes all recorded reductions in ONE launch.  Same fixed-order sums as the per-operator kernels' contract:
// deterministic, independent of scheduling.
struct LnFinalRec {
    const float* part;
    float* dgamma;
    float* dbeta;
    int nblk, N, S, accumulate;          // S: column stride between the two halves of a partial row (N, or LN3_MAXP)
};
constexpr int LNF_MAX = 32;
struct LnFinalBatch {
    int n;
    int pad_;
    LnFinalRec r[LNF_MAX];
};

__global__ void __launch_bounds__(256) ln_param_final_batched_kernel(const __grid_constant__ LnFinalBatch bt) {
    pdl_enter();
    const LnFinalRec& rec = bt.r[blockIdx.y];
    if ((int)blockIdx.x * 32 >= rec.N) return;
    __shared__ float sg[8][33], sb[8][33];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + c;
    float g = 0.f, b = 0.f;
    if (j < rec.N) {
        const float* p0 = rec.part + j;
        const long long row = 2LL * rec.S;
        int k = r;
        for (; k + 8 < rec.nblk; k += 16) {           // two independent loads per half and pass
            const float g0 = p0[k * row], b0 = p0[k * row + rec.S];
            const float g1 = p0[(k + 8) * row], b1 = p0[(k + 8) * row + rec.S];
            g += g0 + g1;
            b += b0 + b1;
        }
        for (; k < rec.nblk; k += 8) {
            g += p0[k * row];
            b += p0[k * row + rec.S];
        }
    }
    sg[r][c] = g;
    sb[r][c] = b;
    __syncthreads();
    if (r == 0 && j < rec.N) {
        const float gt = ((sg[0][c] + sg[1][c]) + (sg[2][c] + sg[3][c])) + ((sg[4][c] + sg[5][c]) + (sg[6][c] + sg[7][c]));
        const float bt2 = ((sb[0][c] + sb[1][c]) + (sb[2][c] + sb[3][c])) + ((sb[4][c] + sb[5][c]) + (sb[6][c] + sb[7][c]));
        rec.dgamma[j] = rec.accumulate ? rec.dgamma[j] + gt : gt;
        rec.dbeta[j] = rec.accumulate ? rec.dbeta[j] + bt2 : bt2;
    }
}

float* g_lnd_base = nullptr;
long long g_lnd_cap = 0, g_lnd_off = 0;
LnFinalBatch g_lnd_q{};

// slot for `need` floats of partials, or null (no scratch attached / full / queue full: the caller finishes at once)
float* lnd_slot(long long need) {
    if (!g_lnd_base || g_lnd_q.n >= LNF_MAX || g_lnd_off + need > g_lnd_cap) return nullptr;
    float* p = g_lnd_base + g_lnd_off;
    g_lnd_off += (need + 63) & ~63LL;
    return p;
}
}  // namespace

void nasrec_internal_ln_defer_scratch(float* base, long long nfloats) {
    static const bool off = getenv("NASREC_LN_DEFER") && atoi(getenv("NASREC_LN_DEFER")) == 0;      // experiment knob
    if (off) base = nullptr;
    g_lnd_base = base;
    g_lnd_cap = base ? nfloats : 0;
    g_lnd_off = 0;
    if (!base) g_lnd_q.n = 0;
}
long long nasrec_internal_ln_pending() { return g_lnd_q.n; }
int nasrec_internal_ln_flush(cudaStream_t st) {
    if (g_lnd_q.n == 0) return 0;
    int maxN = 0;
    for (int i = 0; i < g_lnd_q.n; ++i) maxN = g_lnd_q.r[i].N > maxN ? g_lnd_q.r[i].N : maxN;
    nasrec_launch(ln_param_final_batched_kernel, dim3(cdiv(maxN, 32), g_lnd_q.n), 256, 0, st, g_lnd_q);
    g_lnd_q.n = 0;                      // slots are NOT recycled before the scratch is re-attached: the launch may still read them
    return nasrec_launch_status();
}

extern "C" {

int nasrec_ln_fwd(const float* x, int64_t ldx, int M, int N, const float* gamma, const float* beta, float eps,
                  int relu, int d_out, float* y, int64_t ldy, float* mean, float* rstd, int accumulate,
                  void* stream) {
    CHECK_ARG(x && gamma && beta && y && mean && rstd && M > 0 && N > 0 && d_out >= 0 && d_out <= N);
    if (N > LN_MAXN) return NASREC_ETOOBIG;
    if (N > 256)
        nasrec_launch(ln_fwd_wide_kernel, cdiv(M, LNW_RPC), 256, 0, as_stream(stream), x, ldx, M, N, gamma, beta, eps, relu, d_out, y,
                      ldy, mean, rstd, accumulate);
    else if (N <= 64)
        nasrec_launch(ln_fwd_kernel<2>, cdiv(M, 4), 128, 0, as_stream(stream), x, ldx, M, N, gamma, beta, eps, relu, d_out, y, ldy,
                      mean, rstd, accumulate);
    else if (N <= 128)
        nasrec_launch(ln_fwd_kernel<4>, cdiv(M, 4), 128, 0, as_stream(stream), x, ldx, M, N, gamma, beta, eps, relu, d_out, y, ldy,
                      mean, rstd, accumulate);
    else
        nasrec_launch(ln_fwd_kernel<8>, cdiv(M, 4), 128, 0, as_stream(stream), x, ldx, M, N, gamma, beta, eps, relu, d_out, y, ldy,
                      mean, rstd, accumulate);
    return nasrec_launch_status();
}

int nasrec_ln_bwd(const float* dy, int64_t lddy, int d_out, const float* x, int64_t ldx, int M, int N,
                  const float* gamma, const float* beta, const float* mean, const float* rstd, int relu,
                  float* dx, int64_t lddx, float* dgamma, float* dbeta, int accumulate_params, void* stream) {
    CHECK_ARG(dy && x && gamma && beta && mean && rstd && M > 0 && N > 0 && d_out >= 0 && d_out <= N);
    if (N > LN_MAXN) return NASREC_ETOOBIG;
    cudaStream_t st = as_stream(stream);
    const bool want = dgamma && dbeta;
    if (!dx && !want) return 0;
    float* ws = nullptr;
    long long nws = 0;
    nasrec_internal_workspace(&ws, &nws);
    int grid = cdiv(M, 4);
    if (grid > 296) grid = 296;                    // persistent warps: fixed row -> warp assignment
    // deferred second stage: partials go to a slot of the attached scratch and the final sums join the batched launch
    // (an in-place accumulate may depend on a queued record for the same parameters: finish the queue first)
    float* slot = nullptr;
    if (want && g_lnd_base) {
        if (accumulate_params) {
            const int rc = nasrec_internal_ln_flush(st);
            if (rc) return rc;
        } else {
            slot = lnd_slot((long long)grid * 2 * N);
        }
    }
    if (slot) {
        ws = slot;
        nws = (long long)grid * 2 * N;
    }
    const bool fused = want && ws && nws >= (long long)grid * 2 * N;
    if (dx || fused) {
        if (N > 256) {
            const size_t smem = fused ? (size_t)2 * LNW_RPC * N * sizeof(float) : 0;
            nasrec_launch(ln_bwd_wide_kernel, grid, 256, smem, st, dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd, relu, dx,
                          lddx, fused ? ws : nullptr);
        } else {
            const size_t smem = fused ? (size_t)8 * N * sizeof(float) : 0;
            if (N <= 64)
                nasrec_launch(ln_bwd_kernel<2>, grid, 128, smem, st, dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd, relu, dx,
                              lddx, fused ? ws : nullptr);
            else if (N <= 128)
                nasrec_launch(ln_bwd_kernel<4>, grid, 128, smem, st, dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd, relu, dx,
                              lddx, fused ? ws : nullptr);
            else
                nasrec_launch(ln_bwd_kernel<8>, grid, 128, smem, st, dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd, relu, dx,
                              lddx, fused ? ws : nullptr);
        }
        int rc = nasrec_launch_status();
        if (rc) return rc;
    }
    if (fused && slot) {
        g_lnd_q.r[g_lnd_q.n++] = LnFinalRec{slot, dgamma, dbeta, grid, N, N, 0};
        return 0;
    }
    if (fused) {
        nasrec_launch(ln_param_final_kernel, cdiv(N, 32), 256, 0, st, ws, grid, N, dgamma, dbeta, accumulate_params);
        return nasrec_launch_status();
    }
    if (want) {
        nasrec_launch(ln_param_grad_kernel, cdiv(N, 32), 1024, 0, st, dy, lddy, d_out, x, ldx, M, N, gamma, beta, mean, rstd,
                                                           relu, dgamma, dbeta, accumulate_params);
        return nasrec_launch_status();
    }
    return 0;
}

int nasrec_ln3_fwd(const float* z, int64_t z_bstride, int B, int P, const float* gamma, const float* beta,
                   float eps, int relu, int p_out, float* y, int64_t y_bstride, float* mean, float* rstd,
                   int accumulate, void* stream) {
    CHECK_ARG(z && gamma && beta && y && mean && rstd && B > 0 && P > 0 && p_out >= 0 && p_out <= P);
    if (P > LN3_MAXP) return NASREC_ETOOBIG;
    if (P > 16)
        nasrec_launch(ln3_fwd_wide_kernel, cdiv(B, 2), 256, 0, as_stream(stream), z, z_bstride, B, P, gamma, beta, eps, relu, p_out, y,
                      y_bstride, mean, rstd, accumulate);
    else
        nasrec_launch(ln3_fwd_kernel, cdiv(B, 4), 128, 0, as_stream(stream), z, z_bstride, B, P, gamma, beta, eps, relu, p_out, y,
                                                                 y_bstride, mean, rstd, accumulate);
    return nasrec_launch_status();
}

int nasrec_ln3_bwd(const float* dy, int64_t dy_bstride, int p_out, const float* z, int64_t z_bstride, int B,
                   int P, const float* gamma, const float* beta, const float* mean, const float* rstd, int relu,
                   float* dz, int64_t dz_bstride, float* dgamma, float* dbeta, int accumulate_params,
                   void* stream) {
    CHECK_ARG(dy && z && gamma && beta && mean && rstd && B > 0 && P > 0 && p_out >= 0 && p_out <= P);
    if (P > LN3_MAXP) return NASREC_ETOOBIG;
    cudaStream_t st = as_stream(stream);
    const bool want = dgamma && dbeta;
    if (!dz && !want) return 0;
    float* ws = nullptr;
    long long nws = 0;
    nasrec_internal_workspace(&ws, &nws);
    int grid = cdiv(B, 4);
    if (grid > 148) grid = 148;                    // persistent warps: fixed sample -> warp assignment
    float* slot = nullptr;                         // deferred second stage, as in nasrec_ln_bwd
    if (want && g_lnd_base) {
        if (accumulate_params) {
            const int rc = nasrec_internal_ln_flush(st);
            if (rc) return rc;
        } else {
            slot = lnd_slot((long long)grid * 2 * LN3_MAXP);
        }
    }
    if (slot) {
        ws = slot;
        nws = (long long)grid * 2 * LN3_MAXP;
    }
    const bool fused = want && ws && nws >= (long long)grid * 2 * LN3_MAXP;
    if (dz || fused) {
        if (P > 16)
            nasrec_launch(ln3_bwd_wide_kernel, grid, 256, 0, st, dy, dy_bstride, p_out, z, z_bstride, B, P, gamma, beta, mean, rstd,
                          relu, dz, dz_bstride, fused ? ws : nullptr);
        else
            nasrec_launch(ln3_bwd_kernel, grid, 128, 0, st, dy, dy_bstride, p_out, z, z_bstride, B, P, gamma, beta, mean, rstd, relu,
                                                 dz, dz_bstride, fused ? ws : nullptr);
        int rc = nasrec_launch_status();
        if (rc) return rc;
    }
    if (fused && slot) {
        g_lnd_q.r[g_lnd_q.n++] = LnFinalRec{slot, dgamma, dbeta, grid, P, LN3_MAXP, 0};
        return 0;
    }
    if (fused) {
        nasrec_launch(ln3_param_final_kernel, cdiv(P, 32), 256, 0, st, ws, grid, P, dgamma, dbeta, accumulate_params);
        return nasrec_launch_status();
    }
    if (want) {
        nasrec_launch(ln3_param_grad_kernel, P, 256, 0, st, dy, dy_bstride, p_out, z, z_bstride, B, P, gamma, beta, mean, rstd,
                                                 relu, dgamma, dbeta, accumulate_params);
        return nasrec_launch_status();
    }
    return 0;
}

int nasrec_sproj_bias_grad(const float* dZ, int64_t dz_bstride, int P, int B, float* db, int accumulate,
                           void* stream) {
    CHECK_ARG(dZ && db && P > 0 && B > 0);
    nasrec_launch(sproj_bias_grad_kernel, P, 256, 0, as_stream(stream), dZ, dz_bstride, B, db, accumulate);
    return nasrec_launch_status();
}

int nasrec_act_fwd(const float* x, int64_t ldx, int M, int N, int relu, float* y, int64_t ldy, int accumulate,
                   void* stream) {
    CHECK_ARG(x && y && M > 0 && N > 0);
    nasrec_launch(act_fwd_kernel, ew_grid((long long)M * N), 256, 0, as_stream(stream), x, ldx, M, N, relu, y, ldy,
                                                                            accumulate);
    return nasrec_launch_status();
}

int nasrec_act_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, int M, int N, int relu, float* dx,
                   int64_t lddx, void* stream) {
    CHECK_ARG(dy && x && dx && M > 0 && N > 0);
    nasrec_launch(act_bwd_kernel, ew_grid((long long)M * N), 256, 0, as_stream(stream), dy, lddy, x, ldx, M, N, relu, dx,
                                                                            lddx);
    return nasrec_launch_status();
}

int nasrec_colsum(const float* x, int64_t ld, int M, int N, float* out, int accumulate, void* stream) {
    CHECK_ARG(x && out && M > 0 && N > 0);
    nasrec_launch(colsum_kernel, cdiv(N, 32), 1024, 0, as_stream(stream), x, ld, M, N, out, accumulate);
    return nasrec_launch_status();
}

}  // extern "C"
