// Fused Transformer ("Attention") node core: 8-head self-attention with head_dim 2
// over L <= 64 tokens of width 16, residual + LayerNorm(16), FC-ReLU-FC,
// residual + LayerNorm(16), row mask -- one thread per token, one CTA per sample.
//
// Replaces nn.MultiheadAttention(16, 8, batch_first=True) and the ~25 tiny launches
// around it in nasrec/supernet/modules.py:664-688.  Not tensor-core shaped
// (head_dim = 2): everything lives in registers / shared memory.
//
// Tokens s_live..L-1 are the masked rows of the reference (all-zero inputs).  They
// still take part as keys/values through the in-proj bias and contribute
// exp(score) to every softmax denominator, exactly as in the reference, which
// passes no key-padding mask (modules.py:653-664).
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int E = 16, H = 8, LMAX = 64, ST = 17;     // ST: padded row stride in smem
constexpr int IN_W = 0, IN_B = 768, OUT_W = 816, OUT_B = 1072, LN1_W = 1088, LN1_B = 1104, FC1_W = 1120,
              FC1_B = 1376, FC2_W = 1392, FC2_B = 1648, LN2_W = 1664, LN2_B = 1680, NPARAM = 1696;
static_assert(NPARAM == NASREC_ATTN_PARAMS, "parameter pack size");
constexpr float LN_EPS = 1e-5f;
constexpr int NPTR = 12;
__constant__ const int kOff[NPTR + 1] = {IN_W, IN_B, OUT_W, OUT_B, LN1_W, LN1_B, FC1_W, FC1_B, FC2_W, FC2_B,
                                         LN2_W, LN2_B, NPARAM};
struct AttnPtrs {
    const float* p[NPTR];
};
__device__ __forceinline__ void load_params(float* P, const AttnPtrs& ap, int t, int nt) {
    for (int k = 0; k < NPTR; ++k) {
        const int o = kOff[k], n = kOff[k + 1] - o;
        for (int i = t; i < n; i += nt) P[o + i] = ap.p[k][i];
    }
}
constexpr float SCALE = 0.70710678118654752440f;      // 1/sqrt(head_dim = 2)

__device__ __forceinline__ float score(float q0, float q1, float k0, float k1) {
    return (q0 * k0 + q1 * k1) * SCALE;
}

// y = W x + b for a 16x16 row-major W in shared memory
// The compiler barrier in front of each 16x16 product matters: without it the 256 shared-memory loads of every product
// of a thread's chain are hoisted above the first shared-memory store of the chain (they may alias it) and ~770 values
// are spilled to local memory and read back one per FMA (ptxas: 5 KB of spill traffic per thread, profiles/r02_sass.md).
__device__ __forceinline__ void matvec16(const float* W, const float* b, const float* x, float* y) {
    asm volatile("" ::: "memory");
#pragma unroll
    for (int c = 0; c < E; ++c) {
        float acc = b[c];
#pragma unroll
        for (int e = 0; e < E; ++e) acc = fmaf(W[c * E + e], x[e], acc);
        y[c] = acc;
    }
}

// y = W^T x
__device__ __forceinline__ void matvec16_t(const float* W, const float* x, float* y) {
    asm volatile("" ::: "memory");
#pragma unroll
    for (int e = 0; e < E; ++e) y[e] = 0.f;
#pragma unroll
    for (int c = 0; c < E; ++c)
#pragma unroll
        for (int e = 0; e < E; ++e) y[e] = fmaf(W[c * E + e], x[c], y[e]);
}

__device__ __forceinline__ void ln16(const float* r, const float* g, const float* b, float* xh, float& rstd,
                                     float* y) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s += r[e];
    const float mean = s * (1.f / E);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const float d = r[e] - mean;
        q = fmaf(d, d, q);
    }
    rstd = 1.0f / sqrtf(q * (1.f / E) + LN_EPS);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        xh[e] = (r[e] - mean) * rstd;
        y[e] = xh[e] * g[e] + b[e];
    }
}

// dy -> dr for y = LN(r); returns through dr; gx = dy*xhat (for dgamma)
__device__ __forceinline__ void ln16_bwd(const float* dy, const float* xh, float rstd, const float* g, float* dr) {
    float s1 = 0.f, s2 = 0.f;
    float a[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        a[e] = dy[e] * g[e];
        s1 += a[e];
        s2 = fmaf(a[e], xh[e], s2);
    }
    const float c1 = s1 * (1.f / E), c2 = s2 * (1.f / E);
#pragma unroll
    for (int e = 0; e < E; ++e) dr[e] = rstd * (a[e] - c1 - xh[e] * c2);
}

struct Fwd {
    float x[E], q[E], o[E], m[H], l[H];
    float xh1[E], h1[E], f1[E], xh2[E], y[E];
    float rstd1, rstd2;
};

// qkv for one token; writes k, v rows to shared memory
__device__ __forceinline__ void token_qkv(const float* P, const float* x, float* q, float* Ks, float* Vs, int t) {
#pragma unroll
    for (int c = 0; c < 3 * E; ++c) {
        float acc = P[IN_B + c];
#pragma unroll
        for (int e = 0; e < E; ++e) acc = fmaf(P[IN_W + c * E + e], x[e], acc);
        if (c < E) q[c] = acc;
        else if (c < 2 * E) Ks[t * ST + (c - E)] = acc;
        else Vs[t * ST + (c - 2 * E)] = acc;
    }
}

// everything behind the attention output f.o for one live token: out_proj, residual + LN, FC-ReLU-FC, residual + LN
__device__ __forceinline__ void token_tail(const float* P, Fwd& f) {
    float a[E], r[E];
    matvec16(P + OUT_W, P + OUT_B, f.o, a);
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = a[e] + f.x[e];
    ln16(r, P + LN1_W, P + LN1_B, f.xh1, f.rstd1, f.h1);
    float pre[E], f2[E];
    matvec16(P + FC1_W, P + FC1_B, f.h1, pre);
#pragma unroll
    for (int e = 0; e < E; ++e) f.f1[e] = fmaxf(pre[e], 0.f);
    matvec16(P + FC2_W, P + FC2_B, f.f1, f2);
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = f.h1[e] + f2[e];
    ln16(r, P + LN2_W, P + LN2_B, f.xh2, f.rstd2, f.y);
}

// attention + the two residual/LN stages for one live token (needs Ks/Vs complete)
__device__ __forceinline__ void token_rest(const float* P, const float* Ks, const float* Vs, int L, Fwd& f) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float q0 = f.q[2 * h], q1 = f.q[2 * h + 1];
        float mx = -INFINITY;
        for (int j = 0; j < L; ++j) mx = fmaxf(mx, score(q0, q1, Ks[j * ST + 2 * h], Ks[j * ST + 2 * h + 1]));
        float sum = 0.f, o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < L; ++j) {
            const float p = expf(score(q0, q1, Ks[j * ST + 2 * h], Ks[j * ST + 2 * h + 1]) - mx);
            sum += p;
            o0 = fmaf(p, Vs[j * ST + 2 * h], o0);
            o1 = fmaf(p, Vs[j * ST + 2 * h + 1], o1);
        }
        f.m[h] = mx;
        f.l[h] = sum;
        f.o[2 * h] = o0 / sum;
        f.o[2 * h + 1] = o1 / sum;
    }
    token_tail(P, f);
}

__device__ __forceinline__ void load_row16(const float* p, float* v) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = p4[i];
        v[4 * i] = t.x;
        v[4 * i + 1] = t.y;
        v[4 * i + 2] = t.z;
        v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row16(float* p, const float* v) {
    float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

__global__ void __launch_bounds__(LMAX) attn_fwd_kernel(const float* __restrict__ x, long long xbs, int L,
                                                        int s_live, const __grid_constant__ AttnPtrs ap,
                                                        float* __restrict__ y, long long ybs, int B) {
    pdl_enter();
    __shared__ float P[NPARAM];
    __shared__ float Ks[LMAX * ST], Vs[LMAX * ST];
    const int t = threadIdx.x;
    load_params(P, ap, t, LMAX);
    __syncthreads();
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        Fwd f;
        const bool active = t < L, live = t < s_live;
        if (live) load_row16(x + (long long)b * xbs + t * E, f.x);
        else {
#pragma unroll
            for (int e = 0; e < E; ++e) f.x[e] = 0.f;
        }
        if (active) token_qkv(P, f.x, f.q, Ks, Vs, t);
        __syncthreads();
        if (live) {
            token_rest(P, Ks, Vs, L, f);
            store_row16(y + (long long)b * ybs + t * E, f.y);
        }
        __syncthreads();
    }
}

// dynamic shared memory layout of the backward kernel (floats)
constexpr int S_P = 0;
constexpr int S_KS = S_P + NPARAM;
constexpr int S_VS = S_KS + LMAX * ST;
constexpr int S_QS = S_VS + LMAX * ST;
constexpr int S_DOS = S_QS + LMAX * ST;
constexpr int S_MS = S_DOS + LMAX * ST;          // [LMAX][9]
constexpr int S_LS = S_MS + LMAX * 9;
constexpr int S_DS = S_LS + LMAX * 9;
constexpr int S_X = S_DS + LMAX * 9;
constexpr int S_DA = S_X + LMAX * ST;
constexpr int S_O = S_DA + LMAX * ST;
constexpr int S_DF1 = S_O + LMAX * ST;
constexpr int S_H1 = S_DF1 + LMAX * ST;
constexpr int S_DF2 = S_H1 + LMAX * ST;
constexpr int S_F1 = S_DF2 + LMAX * ST;
constexpr int S_GX1 = S_F1 + LMAX * ST;
constexpr int S_DH1 = S_GX1 + LMAX * ST;
constexpr int S_GX2 = S_DH1 + LMAX * ST;
constexpr int S_DY = S_GX2 + LMAX * ST;
constexpr int S_DQKV = S_DY + LMAX * ST;         // [LMAX][49]
constexpr int S_PG = S_DQKV + LMAX * 49;         // [NPARAM] per-CTA parameter-gradient accumulator
constexpr int S_TOTAL = S_PG + NPARAM;

__device__ __forceinline__ void put16(float* S, int t, const float* v) {
#pragma unroll
    for (int e = 0; e < E; ++e) S[t * ST + e] = v[e];
}

__global__ void __launch_bounds__(LMAX) attn_bwd_kernel(const float* __restrict__ dy, long long dybs,
                                                        const float* __restrict__ x, long long xbs, int L,
                                                        int s_live, const __grid_constant__ AttnPtrs ap,
                                                        float* __restrict__ dx, long long dxbs,
                                                        float* __restrict__ ws, int B) {
    pdl_enter();
    extern __shared__ float sm[];
    float* P = sm + S_P;
    float *Ks = sm + S_KS, *Vs = sm + S_VS, *Qs = sm + S_QS, *DOs = sm + S_DOS;
    float *Ms = sm + S_MS, *Ls = sm + S_LS, *Ds = sm + S_DS;
    float* PG = sm + S_PG;
    float* DQKV = sm + S_DQKV;
    const int t = threadIdx.x;
    load_params(P, ap, t, LMAX);
    for (int i = t; i < NPARAM; i += LMAX) PG[i] = 0.f;
    __syncthreads();
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const bool active = t < L, live = t < s_live;
        Fwd f;
        float kt[E], vt[E];
        if (live) load_row16(x + (long long)b * xbs + t * E, f.x);
        else {
#pragma unroll
            for (int e = 0; e < E; ++e) f.x[e] = 0.f;
        }
        if (active) {
            token_qkv(P, f.x, f.q, Ks, Vs, t);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                kt[e] = Ks[t * ST + e];
                vt[e] = Vs[t * ST + e];
            }
        }
        __syncthreads();
        float dxres[E], dq[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            dxres[e] = 0.f;
            dq[e] = 0.f;
        }
        float dout[E];   // gradient w.r.t. the attention output o (before out_proj)
        float Dh[H];
        if (live) {
            token_rest(P, Ks, Vs, L, f);
            float g[E], dr2[E], df1[E], dh1[E], tmp[E], da[E];
            load_row16(dy + (long long)b * dybs + t * E, g);
            // LN2
            ln16_bwd(g, f.xh2, f.rstd2, P + LN2_W, dr2);
#pragma unroll
            for (int e = 0; e < E; ++e) tmp[e] = g[e] * f.xh2[e];
            put16(sm + S_GX2, t, tmp);
            put16(sm + S_DY, t, g);
            // fc2 / relu / fc1
            matvec16_t(P + FC2_W, dr2, df1);
#pragma unroll
            for (int e = 0; e < E; ++e) df1[e] = f.f1[e] > 0.f ? df1[e] : 0.f;
            matvec16_t(P + FC1_W, df1, dh1);
#pragma unroll
            for (int e = 0; e < E; ++e) dh1[e] += dr2[e];
            put16(sm + S_DF2, t, dr2);
            put16(sm + S_F1, t, f.f1);
            put16(sm + S_DF1, t, df1);
            put16(sm + S_H1, t, f.h1);
            // LN1
            ln16_bwd(dh1, f.xh1, f.rstd1, P + LN1_W, da);
#pragma unroll
            for (int e = 0; e < E; ++e) tmp[e] = dh1[e] * f.xh1[e];
            put16(sm + S_GX1, t, tmp);
            put16(sm + S_DH1, t, dh1);
#pragma unroll
            for (int e = 0; e < E; ++e) dxres[e] = da[e];
            put16(sm + S_DA, t, da);
            put16(sm + S_O, t, f.o);
            put16(sm + S_X, t, f.x);
            // out_proj
            matvec16_t(P + OUT_W, da, dout);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                Dh[h] = dout[2 * h] * f.o[2 * h] + dout[2 * h + 1] * f.o[2 * h + 1];
                Ms[t * 9 + h] = f.m[h];
                Ls[t * 9 + h] = f.l[h];
                Ds[t * 9 + h] = Dh[h];
            }
            put16(Qs, t, f.q);
            put16(DOs, t, dout);
            // pass A: dq_i = scale * sum_j dS_ij k_j
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float q0 = f.q[2 * h], q1 = f.q[2 * h + 1];
                const float inv_l = 1.f / f.l[h];
                float a0 = 0.f, a1 = 0.f;
                for (int j = 0; j < L; ++j) {
                    const float k0 = Ks[j * ST + 2 * h], k1 = Ks[j * ST + 2 * h + 1];
                    const float p = expf(score(q0, q1, k0, k1) - f.m[h]) * inv_l;
                    const float dp = dout[2 * h] * Vs[j * ST + 2 * h] + dout[2 * h + 1] * Vs[j * ST + 2 * h + 1];
                    const float ds = p * (dp - Dh[h]);
                    a0 = fmaf(ds, k0, a0);
                    a1 = fmaf(ds, k1, a1);
                }
                dq[2 * h] = a0 * SCALE;
                dq[2 * h + 1] = a1 * SCALE;
            }
        }
        __syncthreads();
        // pass B: as key/value token t, gather from every live query i
        float dk[E], dv[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            dk[e] = 0.f;
            dv[e] = 0.f;
        }
        if (active) {
            for (int i = 0; i < s_live; ++i) {
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float q0 = Qs[i * ST + 2 * h], q1 = Qs[i * ST + 2 * h + 1];
                    const float d0 = DOs[i * ST + 2 * h], d1 = DOs[i * ST + 2 * h + 1];
                    const float p = expf(score(q0, q1, kt[2 * h], kt[2 * h + 1]) - Ms[i * 9 + h]) / Ls[i * 9 + h];
                    const float dp = d0 * vt[2 * h] + d1 * vt[2 * h + 1];
                    const float ds = p * (dp - Ds[i * 9 + h]) * SCALE;
                    dk[2 * h] = fmaf(ds, q0, dk[2 * h]);
                    dk[2 * h + 1] = fmaf(ds, q1, dk[2 * h + 1]);
                    dv[2 * h] = fmaf(p, d0, dv[2 * h]);
                    dv[2 * h + 1] = fmaf(p, d1, dv[2 * h + 1]);
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                DQKV[t * 49 + e] = dq[e];
                DQKV[t * 49 + E + e] = dk[e];
                DQKV[t * 49 + 2 * E + e] = dv[e];
            }
        }
        if (live && dx) {
            float o[E];
#pragma unroll
            for (int e = 0; e < E; ++e) o[e] = dxres[e];
#pragma unroll
            for (int c = 0; c < E; ++c)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    o[e] = fmaf(P[IN_W + c * E + e], dq[c], o[e]);
                    o[e] = fmaf(P[IN_W + (E + c) * E + e], dk[c], o[e]);
                    o[e] = fmaf(P[IN_W + (2 * E + c) * E + e], dv[c], o[e]);
                }
            store_row16(dx + (long long)b * dxbs + t * E, o);
        }
        __syncthreads();
        // parameter-gradient accumulation: every PG element is owned by exactly one thread
        if (t < 48) {
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.f;
            float bsum = 0.f;
            const float* X = sm + S_X;
            for (int i = 0; i < L; ++i) {
                const float d = DQKV[i * 49 + t];
                bsum += d;
                if (i < s_live) {
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(d, X[i * ST + e], acc[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) PG[IN_W + t * E + e] += acc[e];
            PG[IN_B + t] += bsum;
            // one row of one of the three 16x16 matrices
            const int grp = t >> 4, c = t & 15;
            const float* G = sm + (grp == 0 ? S_DA : (grp == 1 ? S_DF1 : S_DF2));
            const float* A = sm + (grp == 0 ? S_O : (grp == 1 ? S_H1 : S_F1));
            const int wofs = grp == 0 ? OUT_W : (grp == 1 ? FC1_W : FC2_W);
            const int bofs = grp == 0 ? OUT_B : (grp == 1 ? FC1_B : FC2_B);
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.f;
            bsum = 0.f;
            for (int i = 0; i < s_live; ++i) {
                const float d = G[i * ST + c];
                bsum += d;
#pragma unroll
                for (int e = 0; e < E; ++e) acc[e] = fmaf(d, A[i * ST + e], acc[e]);
            }
#pragma unroll
            for (int e = 0; e < E; ++e) PG[wofs + c * E + e] += acc[e];
            PG[bofs + c] += bsum;
        } else {
            const int c = t - 48;
            float g1 = 0.f, b1 = 0.f, g2 = 0.f, b2 = 0.f;
            for (int i = 0; i < s_live; ++i) {
                g1 += sm[S_GX1 + i * ST + c];
                b1 += sm[S_DH1 + i * ST + c];
                g2 += sm[S_GX2 + i * ST + c];
                b2 += sm[S_DY + i * ST + c];
            }
            PG[LN1_W + c] += g1;
            PG[LN1_B + c] += b1;
            PG[LN2_W + c] += g2;
            PG[LN2_B + c] += b2;
        }
        __syncthreads();
    }
    for (int i = t; i < NPARAM; i += LMAX) ws[(long long)blockIdx.x * NPARAM + i] = PG[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Four threads per token (round 2).  One thread per token leaves a sample with two warps walking ~25 000 dependent
// instructions each (backward: 168 registers + 5 KB of spills, 124 us per launch at B = 256).  Three quarters of that
// are the per-head loops over the keys, and heads are independent: thread (t, hq) of a 256-thread CTA owns heads 2*hq
// and 2*hq + 1 of token t in every pass over keys / queries, a quarter of the in-projection outputs and a quarter of
// the dX columns; only the 16-wide tail (out_proj, LayerNorms, FFN and their backward) stays with thread (t, 0).
// Every sum keeps the order of the one-thread-per-token kernels above, which stay selectable (NASREC_ATTN_OLD=1).
constexpr int NT4 = 4 * LMAX;

// in-projection outputs 12*hq .. 12*hq + 11 of token t (q | k | v rows of in_proj_weight)
__device__ __forceinline__ void qkv_quarter(const float* P, const float* x, float* Qs, float* Ks, float* Vs, int t,
                                            int hq) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int c = 12 * hq + i;
        float acc = P[IN_B + c];
#pragma unroll
        for (int e = 0; e < E; ++e) acc = fmaf(P[IN_W + c * E + e], x[e], acc);
        float* dst = c < E ? Qs : (c < 2 * E ? Ks : Vs);
        dst[t * ST + (c & (E - 1))] = acc;
    }
}

// softmax(q k^T / sqrt 2) v for heads 2*hq, 2*hq + 1 of query token t; q4 / o4: [head][2], m2 / l2: row max / sum
__device__ __forceinline__ void heads_fwd(const float* Qs, const float* Ks, const float* Vs, int L, int t, int hq,
                                          float* q4, float* o4, float* m2, float* l2) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * hq + hh;
        const float q0 = Qs[t * ST + 2 * h], q1 = Qs[t * ST + 2 * h + 1];
        float mx = -INFINITY;
        for (int j = 0; j < L; ++j) mx = fmaxf(mx, score(q0, q1, Ks[j * ST + 2 * h], Ks[j * ST + 2 * h + 1]));
        float sum = 0.f, o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < L; ++j) {
            const float p = expf(score(q0, q1, Ks[j * ST + 2 * h], Ks[j * ST + 2 * h + 1]) - mx);
            sum += p;
            o0 = fmaf(p, Vs[j * ST + 2 * h], o0);
            o1 = fmaf(p, Vs[j * ST + 2 * h + 1], o1);
        }
        q4[2 * hh] = q0;
        q4[2 * hh + 1] = q1;
        m2[hh] = mx;
        l2[hh] = sum;
        o4[2 * hh] = o0 / sum;
        o4[2 * hh + 1] = o1 / sum;
    }
}

__global__ void __launch_bounds__(NT4) attn_fwd4_kernel(const float* __restrict__ x, long long xbs, int L, int s_live,
                                                        const __grid_constant__ AttnPtrs ap, float* __restrict__ y,
                                                        long long ybs, int B) {
    pdl_enter();
    __shared__ float P[NPARAM];
    __shared__ float Ks[LMAX * ST], Vs[LMAX * ST], Qs[LMAX * ST], Os[LMAX * ST];
    const int tid = threadIdx.x, t = tid & (LMAX - 1), hq = tid >> 6;
    load_params(P, ap, tid, NT4);
    __syncthreads();
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const bool active = t < L, live = t < s_live;
        Fwd f;
        if (live) load_row16(x + (long long)b * xbs + t * E, f.x);
        else {
#pragma unroll
            for (int e = 0; e < E; ++e) f.x[e] = 0.f;
        }
        if (active) qkv_quarter(P, f.x, Qs, Ks, Vs, t, hq);
        __syncthreads();
        if (live) {
            float q4[4], o4[4], m2[2], l2[2];
            heads_fwd(Qs, Ks, Vs, L, t, hq, q4, o4, m2, l2);
#pragma unroll
            for (int k = 0; k < 4; ++k) Os[t * ST + 4 * hq + k] = o4[k];
        }
        __syncthreads();
        // the tail reads Os, P and registers only; the next sample's first phase writes Qs / Ks / Vs, and Os is not
        // written again before the barrier behind that phase, so no third barrier is needed
        if (live && hq == 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) f.o[e] = Os[t * ST + e];
            token_tail(P, f);
            store_row16(y + (long long)b * ybs + t * E, f.y);
        }
    }
}

__global__ void __launch_bounds__(NT4, 2) attn_bwd4_kernel(const float* __restrict__ dy, long long dybs,
                                                        const float* __restrict__ x, long long xbs, int L, int s_live,
                                                        const __grid_constant__ AttnPtrs ap, float* __restrict__ dx,
                                                        long long dxbs, float* __restrict__ ws, int B) {
    pdl_enter();
    extern __shared__ float sm[];
    float* P = sm + S_P;
    float *Ks = sm + S_KS, *Vs = sm + S_VS, *Qs = sm + S_QS, *DOs = sm + S_DOS;
    float *Ms = sm + S_MS, *Ls = sm + S_LS, *Ds = sm + S_DS;
    float* PG = sm + S_PG;
    float* DQKV = sm + S_DQKV;
    const int tid = threadIdx.x, t = tid & (LMAX - 1), hq = tid >> 6;
    load_params(P, ap, tid, NT4);
    for (int i = tid; i < NPARAM; i += NT4) PG[i] = 0.f;
    __syncthreads();
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const bool active = t < L, live = t < s_live;
        Fwd f;                                        // thread (t, 0) uses all of it, the others only f.x
        if (live) load_row16(x + (long long)b * xbs + t * E, f.x);
        else {
#pragma unroll
            for (int e = 0; e < E; ++e) f.x[e] = 0.f;
        }
        if (active) qkv_quarter(P, f.x, Qs, Ks, Vs, t, hq);
        if (live && hq == 0) put16(sm + S_X, t, f.x);
        __syncthreads();
        // ---- forward attention of my two heads; my token's k / v of those heads for pass B
        float q4[4], o4[4], m2[2], l2[2], k4[4], v4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            k4[k] = active ? Ks[t * ST + 4 * hq + k] : 0.f;
            v4[k] = active ? Vs[t * ST + 4 * hq + k] : 0.f;
            q4[k] = 0.f;
            o4[k] = 0.f;
        }
        m2[0] = m2[1] = 0.f;
        l2[0] = l2[1] = 1.f;
        if (live) {
            heads_fwd(Qs, Ks, Vs, L, t, hq, q4, o4, m2, l2);
#pragma unroll
            for (int k = 0; k < 4; ++k) sm[S_O + t * ST + 4 * hq + k] = o4[k];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                Ms[t * 9 + 2 * hq + hh] = m2[hh];
                Ls[t * 9 + 2 * hq + hh] = l2[hh];
            }
        }
        __syncthreads();
        // ---- the 16-wide tail and its backward, one thread per token
        if (live && hq == 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) f.o[e] = sm[S_O + t * ST + e];
            token_tail(P, f);
            float g[E], dr2[E], df1[E], dh1[E], tmp[E], da[E], dout[E];
            load_row16(dy + (long long)b * dybs + t * E, g);
            // LN2
            ln16_bwd(g, f.xh2, f.rstd2, P + LN2_W, dr2);
#pragma unroll
            for (int e = 0; e < E; ++e) tmp[e] = g[e] * f.xh2[e];
            put16(sm + S_GX2, t, tmp);
            put16(sm + S_DY, t, g);
            // fc2 / relu / fc1
            matvec16_t(P + FC2_W, dr2, df1);
#pragma unroll
            for (int e = 0; e < E; ++e) df1[e] = f.f1[e] > 0.f ? df1[e] : 0.f;
            matvec16_t(P + FC1_W, df1, dh1);
#pragma unroll
            for (int e = 0; e < E; ++e) dh1[e] += dr2[e];
            put16(sm + S_DF2, t, dr2);
            put16(sm + S_F1, t, f.f1);
            put16(sm + S_DF1, t, df1);
            put16(sm + S_H1, t, f.h1);
            // LN1
            ln16_bwd(dh1, f.xh1, f.rstd1, P + LN1_W, da);
#pragma unroll
            for (int e = 0; e < E; ++e) tmp[e] = dh1[e] * f.xh1[e];
            put16(sm + S_GX1, t, tmp);
            put16(sm + S_DH1, t, dh1);
            put16(sm + S_DA, t, da);                  // also the residual part of dX
            // out_proj
            matvec16_t(P + OUT_W, da, dout);
#pragma unroll
            for (int h = 0; h < H; ++h) Ds[t * 9 + h] = dout[2 * h] * f.o[2 * h] + dout[2 * h + 1] * f.o[2 * h + 1];
            put16(DOs, t, dout);
        }
        __syncthreads();
        // ---- pass A: dq_i = scale * sum_j dS_ij k_j for my two heads
        float dq4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dq4[k] = 0.f;
        if (live) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int h = 2 * hq + hh;
                const float q0 = q4[2 * hh], q1 = q4[2 * hh + 1];
                const float d0 = DOs[t * ST + 2 * h], d1 = DOs[t * ST + 2 * h + 1];
                const float Dh = Ds[t * 9 + h];
                const float inv_l = 1.f / l2[hh];
                float a0 = 0.f, a1 = 0.f;
                for (int j = 0; j < L; ++j) {
                    const float k0 = Ks[j * ST + 2 * h], k1 = Ks[j * ST + 2 * h + 1];
                    const float p = expf(score(q0, q1, k0, k1) - m2[hh]) * inv_l;
                    const float dp = d0 * Vs[j * ST + 2 * h] + d1 * Vs[j * ST + 2 * h + 1];
                    const float ds = p * (dp - Dh);
                    a0 = fmaf(ds, k0, a0);
                    a1 = fmaf(ds, k1, a1);
                }
                dq4[2 * hh] = a0 * SCALE;
                dq4[2 * hh + 1] = a1 * SCALE;
            }
        }
        // ---- pass B: as key / value token t, gather from every live query i (my two heads)
        if (active) {
            float dk4[4], dv4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                dk4[k] = 0.f;
                dv4[k] = 0.f;
            }
            for (int i = 0; i < s_live; ++i) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int h = 2 * hq + hh;
                    const float q0 = Qs[i * ST + 2 * h], q1 = Qs[i * ST + 2 * h + 1];
                    const float d0 = DOs[i * ST + 2 * h], d1 = DOs[i * ST + 2 * h + 1];
                    const float p = expf(score(q0, q1, k4[2 * hh], k4[2 * hh + 1]) - Ms[i * 9 + h]) / Ls[i * 9 + h];
                    const float dp = d0 * v4[2 * hh] + d1 * v4[2 * hh + 1];
                    const float ds = p * (dp - Ds[i * 9 + h]) * SCALE;
                    dk4[2 * hh] = fmaf(ds, q0, dk4[2 * hh]);
                    dk4[2 * hh + 1] = fmaf(ds, q1, dk4[2 * hh + 1]);
                    dv4[2 * hh] = fmaf(p, d0, dv4[2 * hh]);
                    dv4[2 * hh + 1] = fmaf(p, d1, dv4[2 * hh + 1]);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                DQKV[t * 49 + 4 * hq + k] = dq4[k];
                DQKV[t * 49 + E + 4 * hq + k] = dk4[k];
                DQKV[t * 49 + 2 * E + 4 * hq + k] = dv4[k];
            }
        }
        __syncthreads();
        // ---- dX columns 4*hq .. 4*hq + 3 of my token: residual + in_proj^T (dq | dk | dv)
        if (live && dx) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = sm[S_DA + t * ST + 4 * hq + k];
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const float dqc = DQKV[t * 49 + c], dkc = DQKV[t * 49 + E + c], dvc = DQKV[t * 49 + 2 * E + c];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = 4 * hq + k;
                    o[k] = fmaf(P[IN_W + c * E + e], dqc, o[k]);
                    o[k] = fmaf(P[IN_W + (E + c) * E + e], dkc, o[k]);
                    o[k] = fmaf(P[IN_W + (2 * E + c) * E + e], dvc, o[k]);
                }
            }
            *reinterpret_cast<float4*>(dx + (long long)b * dxbs + t * E + 4 * hq) = make_float4(o[0], o[1], o[2], o[3]);
        }
        // ---- parameter gradients: every PG element is owned by exactly one thread (three groups of threads)
        if (tid < 48) {                               // one row of in_proj_weight + its bias element
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.f;
            float bsum = 0.f;
            const float* X = sm + S_X;
            for (int i = 0; i < L; ++i) {
                const float d = DQKV[i * 49 + tid];
                bsum += d;
                if (i < s_live) {
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(d, X[i * ST + e], acc[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) PG[IN_W + tid * E + e] += acc[e];
            PG[IN_B + tid] += bsum;
        } else if (tid >= 64 && tid < 112) {          // one row of out_proj / fc1 / fc2 + its bias element
            const int u = tid - 64, grp = u >> 4, c = u & 15;
            const float* G = sm + (grp == 0 ? S_DA : (grp == 1 ? S_DF1 : S_DF2));
            const float* A = sm + (grp == 0 ? S_O : (grp == 1 ? S_H1 : S_F1));
            const int wofs = grp == 0 ? OUT_W : (grp == 1 ? FC1_W : FC2_W);
            const int bofs = grp == 0 ? OUT_B : (grp == 1 ? FC1_B : FC2_B);
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.f;
            float bsum = 0.f;
            for (int i = 0; i < s_live; ++i) {
                const float d = G[i * ST + c];
                bsum += d;
#pragma unroll
                for (int e = 0; e < E; ++e) acc[e] = fmaf(d, A[i * ST + e], acc[e]);
            }
#pragma unroll
            for (int e = 0; e < E; ++e) PG[wofs + c * E + e] += acc[e];
            PG[bofs + c] += bsum;
        } else if (tid >= 128 && tid < 144) {         // the two LayerNorms' gamma / beta
            const int c = tid - 128;
            float g1 = 0.f, b1 = 0.f, g2 = 0.f, b2 = 0.f;
            for (int i = 0; i < s_live; ++i) {
                g1 += sm[S_GX1 + i * ST + c];
                b1 += sm[S_DH1 + i * ST + c];
                g2 += sm[S_GX2 + i * ST + c];
                b2 += sm[S_DY + i * ST + c];
            }
            PG[LN1_W + c] += g1;
            PG[LN1_B + c] += b1;
            PG[LN2_W + c] += g2;
            PG[LN2_B + c] += b2;
        }
        __syncthreads();
    }
    for (int i = tid; i < NPARAM; i += NT4) ws[(long long)blockIdx.x * NPARAM + i] = PG[i];
}

__global__ void attn_param_reduce_kernel(const float* __restrict__ ws, int nblk, float* __restrict__ dparams,
                                         int accumulate) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NPARAM) return;
    float acc = 0.f;
    for (int k = 0; k < nblk; ++k) acc += ws[(long long)k * NPARAM + i];
    dparams[i] = accumulate ? dparams[i] + acc : acc;
}

int attn_grid(int B) {
    const int cap = 148 * 4;
    return B < cap ? B : cap;
}

// NASREC_ATTN_OLD=1: the one-thread-per-token kernels (A/B timing, tests of the two against each other)
bool attn_old() {
    static const bool v = [] {
        const char* e = getenv("NASREC_ATTN_OLD");
        return e && atoi(e) != 0;
    }();
    return v;
}

}  // namespace

extern "C" {

static int fill_ptrs(AttnPtrs& ap, const float* const* params) {
    if (!params) return NASREC_EINVAL;
    for (int k = 0; k < NPTR; ++k) {
        if (!params[k]) return NASREC_EINVAL;
        ap.p[k] = params[k];
    }
    return 0;
}

int nasrec_attn_fwd(const float* x, int64_t x_bstride, int L, int s_live, const float* const* params, float* y,
                    int64_t y_bstride, int B, void* stream) {
    CHECK_ARG(x && y && B > 0 && L > 0 && L <= LMAX && s_live > 0 && s_live <= L);
    AttnPtrs ap;
    if (fill_ptrs(ap, params)) return NASREC_EINVAL;
    // four threads per token at every batch size (measured: 15 vs 31 us at B = 256, 162 vs 193 us at B = 8192, L = 32);
    // NASREC_ATTN_FWD4_MAXB = b keeps the one-thread-per-token kernel above b samples (A/B timing)
    static const int fwd4_maxb = [] {
        const char* e = getenv("NASREC_ATTN_FWD4_MAXB");
        return e ? atoi(e) : (1 << 30);
    }();
    if (attn_old() || B > fwd4_maxb) {
        const int grid = B < 148 * 16 ? B : 148 * 16;
        nasrec_launch(attn_fwd_kernel, grid, LMAX, 0, as_stream(stream), x, x_bstride, L, s_live, ap, y, y_bstride, B);
    } else {
        const int grid = B < 148 * 8 ? B : 148 * 8;
        nasrec_launch(attn_fwd4_kernel, grid, NT4, 0, as_stream(stream), x, x_bstride, L, s_live, ap, y, y_bstride, B);
    }
    return nasrec_launch_status();
}

int64_t nasrec_attn_bwd_ws_floats(int B) { return (int64_t)attn_grid(B) * NPARAM; }

int nasrec_attn_bwd(const float* dy, int64_t dy_bstride, const float* x, int64_t x_bstride, int L, int s_live,
                    const float* const* params, float* dx, int64_t dx_bstride, float* dparams,
                    int accumulate_params, float* ws, int B, void* stream) {
    CHECK_ARG(dy && x && ws && B > 0 && L > 0 && L <= LMAX && s_live > 0 && s_live <= L);
    AttnPtrs ap;
    if (fill_ptrs(ap, params)) return NASREC_EINVAL;
    static bool attr_set = false;
    const size_t smem = (size_t)S_TOTAL * sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_bwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = attn_grid(B);
    cudaStream_t st = as_stream(stream);
    if (attn_old())
        nasrec_launch(attn_bwd_kernel, grid, LMAX, smem, st, dy, dy_bstride, x, x_bstride, L, s_live, ap, dx, dx_bstride, ws, B);
    else
        nasrec_launch(attn_bwd4_kernel, grid, NT4, smem, st, dy, dy_bstride, x, x_bstride, L, s_live, ap, dx, dx_bstride, ws, B);
    int rc = nasrec_launch_status();
    if (rc) return rc;
    if (dparams) {
        nasrec_launch(attn_param_reduce_kernel, cdiv(NPARAM, 256), 256, 0, st, ws, grid, dparams, accumulate_params);
        return nasrec_launch_status();
    }
    return 0;
}

}  // extern "C"
