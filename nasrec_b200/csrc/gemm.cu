// Segment-list SGEMM family (fp32 FFMA path).
//
// One tiled kernel computes, for a small batch of independent problems,
//     C(m,n) = sum_terms sum_k A_t(m,k) * B_t(n,k)  (+ bias[n]) (+ addend(m,n))
// where every operand is a strided "view" with an optional two-level index
// (i -> (i >> sh) * hi + (i & mask) * lo), which is what lets the same kernel
// serve the 2-D linears over zero-padded concats (a K axis made of segments
// living in different tensors) and the projections along the sparse axis of
// [B, rows, 16] tensors without ever materialising a concat or a transpose.
//
// Replaces: F.linear on torch.cat([...zeros...]) in nasrec/supernet/modules.py
// (:171,:223,:340,:359,:385,:489,:578,:584,:648,:740) and supernet.py:598,1140.
#include "gemm_common.cuh"
#include "gemm_tc.cuh"

namespace {
using namespace nasrec_gemm;

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

__device__ __forceinline__ void load_tile(float (*S)[BM + PADM], const View& v, int i0, int I, int k0, int K,
                                          int tid) {
    if (v.contig_j) {
        const int j = tid & 15;
        int i = tid >> 4;
        const bool jok = (k0 + j) < K;
#pragma unroll
        for (int q = 0; q < 4; ++q, i += 16) {
            float val = 0.f;
            if (jok && (i0 + i) < I) val = __ldg(v.p + voff(v, i0 + i, k0 + j));
            S[j][i] = val;
        }
    } else {
        const int i = tid & 63;
        int j = tid >> 6;
        const bool iok = (i0 + i) < I;
#pragma unroll
        for (int q = 0; q < 4; ++q, j += 4) {
            float val = 0.f;
            if (iok && (k0 + j) < K) val = __ldg(v.p + voff(v, i0 + i, k0 + j));
            S[j][i] = val;
        }
    }
}

__global__ void __launch_bounds__(256) gemm64_kernel(const __grid_constant__ Batch bt) {
    pdl_enter();
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Bs[BK][BN + PADM];
    int z = blockIdx.z, pi = 0;
    for (; pi < bt.nprob; ++pi) {
        const int ns = bt.prob[pi].nsplit;
        if (z < ns) break;
        z -= ns;
    }
    if (pi >= bt.nprob) return;
    const Prob& pr = bt.prob[pi];
    const int split = z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= pr.M || n0 >= pr.N) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    int tot = 0;
    for (int t = 0; t < pr.nterm; ++t) tot += (bt.term[pr.term0 + t].K + BK - 1) / BK;
    const int per = (tot + pr.nsplit - 1) / pr.nsplit;
    const int kt_begin = split * per;
    const int kt_end = min(tot, kt_begin + per);
    int kt = 0;
    for (int t = 0; t < pr.nterm; ++t) {
        const Term& tm = bt.term[pr.term0 + t];
        const int nk = (tm.K + BK - 1) / BK;
        if (kt + nk <= kt_begin) {
            kt += nk;
            continue;
        }
        if (kt >= kt_end) break;
        const int kb = max(0, kt_begin - kt), ke = min(nk, kt_end - kt);
        for (int kk = kb; kk < ke; ++kk) {
            const int k0 = kk * BK;
            load_tile(As, tm.a, m0, pr.M, k0, tm.K, tid);
            load_tile(Bs, tm.b, n0, pr.N, k0, tm.K, tid);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < BK; ++j) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[j][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[j][tx * 4]);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
            }
            __syncthreads();
        }
        kt += nk;
    }

    const int cmask = (1 << pr.c_sh_i) - 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty * 4 + r;
        if (m >= pr.M) continue;
        const long long ro = (long long)(m >> pr.c_sh_i) * pr.c_hi_i + (long long)(m & cmask) * pr.c_lo_i;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int n = n0 + tx * 4 + c;
            if (n >= pr.N) continue;
            const long long o = ro + (long long)n * pr.c_hi_j;
            float v = acc[r][c];
            if (pr.bias) v += __ldg(pr.bias + n);
            if (pr.addend) v += pr.addend[o];
            pr.c[o + (long long)split * pr.split_stride] = v;
        }
    }
}

struct RedSeg {
    const float* ws;      // [nsplit][M][N] partial sums
    float* c;             // destination, row stride ldc
    const float* bias;    // indexed by n, or null
    long long ldc;
    int M, N, nsplit, accumulate;
};
struct RedBatch {
    int nseg;
    int pad_;
    RedSeg seg[NASREC_MAX_SEGS];
};

// Deterministic second stage of split-K: fixed-order sum of the partials (+bias, +in-place accumulate).
__global__ void splitk_reduce_kernel(const __grid_constant__ RedBatch rb) {
    pdl_enter();
    const RedSeg& s = rb.seg[blockIdx.y];
    const long long total = (long long)s.M * s.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / s.N), n = (int)(i % s.N);
        float v = 0.f;
        for (int sp = 0; sp < s.nsplit; ++sp) v += s.ws[(long long)sp * total + i];
        if (s.bias) v += __ldg(s.bias + n);
        float* dst = s.c + (long long)m * s.ldc + n;
        *dst = s.accumulate ? (*dst + v) : v;
    }
}

int launch_reduce(const RedBatch& rb, cudaStream_t st) {
    long long maxtot = 0;
    for (int i = 0; i < rb.nseg; ++i) {
        const long long t = (long long)rb.seg[i].M * rb.seg[i].N;
        if (t > maxtot) maxtot = t;
    }
    if (rb.nseg == 0 || maxtot == 0) return 0;
    long long gx = (maxtot + 255) / 256;
    if (gx > 1024) gx = 1024;
    dim3 grid((unsigned)gx, rb.nseg);
    nasrec_launch(splitk_reduce_kernel, grid, 256, 0, st, rb);
    return nasrec_launch_status();
}

float* g_ws = nullptr;          // library workspace for automatic split-K (nasrec_set_workspace)
long long g_ws_floats = 0;
cudaStream_t g_side = nullptr;  // optional second stream for weight-gradient GEMMs (nasrec_set_side_stream)

// Work on the side stream runs concurrently with the main stream: it gets its own half of the workspace.
void ws_region(cudaStream_t st, float** base, long long* n) {
    if (!g_ws) { *base = nullptr; *n = 0; return; }
    // always halves, attached or not, so that split-K decisions (and with them the summation order and the
    // bits of the result) do not depend on whether a side stream is in use
    const long long half = (g_ws_floats / 2) & ~63LL;
    *base = (g_side && st == g_side) ? g_ws + half : g_ws;
    *n = half;
}

int g_gemm_mode = 3;   // 0: fp32 FFMA; 1/3/4: tcgen05 kind::tf32 with 1/3/4 split products (default 3xTF32)

View plain_view(const float* p, long long si, long long sj, int contig_j) {
    View v{};
    v.p = p;
    v.hi_i = si;
    v.hi_j = sj;
    v.contig_j = contig_j;
    return v;
}

void mark_vec16(View& v) {
    const bool jplain = v.sh_j == 0 ? v.hi_j == 1 : (v.lo_j == 1 && v.sh_j >= 2 && (v.hi_j & 3) == 0);
    const bool iok = v.sh_i == 0 ? (v.hi_i & 3) == 0 : ((v.hi_i & 3) == 0 && (v.lo_i & 3) == 0);
    v.vec16 = v.contig_j && jplain && iok && ((reinterpret_cast<uintptr_t>(v.p) & 15) == 0);
}

int launch(Batch& bt, cudaStream_t st) {
    for (int p = 0; p < bt.nprob; ++p)
        for (int t = 0; t < bt.prob[p].nterm; ++t) {
            mark_vec16(bt.term[bt.prob[p].term0 + t].a);
            mark_vec16(bt.term[bt.prob[p].term0 + t].b);
        }
    int maxM = 0, maxN = 0, totz = 0;
    for (int i = 0; i < bt.nprob; ++i) {
        maxM = bt.prob[i].M > maxM ? bt.prob[i].M : maxM;
        maxN = bt.prob[i].N > maxN ? bt.prob[i].N : maxN;
        totz += bt.prob[i].nsplit;
    }
    if (maxM <= 0 || maxN <= 0 || totz <= 0) return 0;
    if (g_gemm_mode != 0) {
        // Skinny launches (few output tiles, long K) leave most SMs idle: split K over several CTAs,
        // partials into the library workspace, fixed-order reduction (deterministic).
        RedBatch rb{};
        const long long ctas = nasrec_gemm::tc_cta_count(bt, nasrec_gemm::tc_pick_bn(bt, maxN));
        float* wsb = nullptr;
        long long wsn = 0;
        ws_region(st, &wsb, &wsn);
        if (wsb && ctas < 64) {
            long long off = 0;
            const int want = (int)(128 / (ctas > 0 ? ctas : 1));
            for (int p = 0; p < bt.nprob; ++p) {
                Prob& pr = bt.prob[p];
                if (pr.nsplit != 1 || pr.c_sh_i != 0 || pr.c_hi_j != 1) continue;
                if (pr.addend && pr.addend != pr.c) continue;
                int ktiles = 0;
                for (int t = 0; t < pr.nterm; ++t) ktiles += (bt.term[pr.term0 + t].K + 31) / 32;
                int ns = want < ktiles / 4 ? want : ktiles / 4;
                if (ns > 16) ns = 16;
                const long long need = (long long)ns * pr.M * pr.N;
                if (ns < 2 || off + need > wsn) continue;
                RedSeg& rs = rb.seg[rb.nseg++];
                rs.ws = wsb + off;
                rs.c = pr.c;
                rs.bias = pr.bias;
                rs.ldc = pr.c_hi_i;
                rs.M = pr.M;
                rs.N = pr.N;
                rs.nsplit = ns;
                rs.accumulate = pr.addend != nullptr;
                pr.c = wsb + off;
                pr.c_hi_i = pr.N;
                pr.bias = nullptr;
                pr.addend = nullptr;
                pr.nsplit = ns;
                pr.split_stride = (long long)pr.M * pr.N;
                off += need;
                totz += ns - 1;
            }
        }
        int rc = nasrec_gemm::launch_tc(bt, maxM, maxN, totz, g_gemm_mode, st);
        if (rc || rb.nseg == 0) return rc;
        return launch_reduce(rb, st);
    }
    dim3 grid(cdiv(maxN, BN), cdiv(maxM, BM), totz);
    if (grid.y > 65535 || grid.z > 65535) return NASREC_ETOOBIG;
    nasrec_launch(gemm64_kernel, grid, 256, 0, st, bt);
    return nasrec_launch_status();
}

bool segs_ok(const nasrec_seg_t* segs, int nseg) {
    if (!segs || nseg <= 0 || nseg > NASREC_MAX_SEGS) return false;
    for (int s = 0; s < nseg; ++s)
        if (!segs[s].ptr || segs[s].width < 0 || segs[s].w_off < 0) return false;
    return true;
}

}  // namespace

void nasrec_internal_workspace(float** ws, long long* nfloats) {   // main-stream region
    ws_region(nullptr, ws, nfloats);
}

cudaStream_t nasrec_internal_side_stream() { return g_side; }
void nasrec_internal_set_side_stream(cudaStream_t s) { g_side = s; }

extern "C" {

int nasrec_set_gemm_mode(int mode) {
    if (mode != 0 && mode != 1 && mode != 3 && mode != 4) return NASREC_EINVAL;
    g_gemm_mode = mode;
    return 0;
}

int nasrec_get_gemm_mode(void) { return g_gemm_mode; }

int nasrec_set_workspace(float* ws, int64_t nfloats) {
    if (nfloats < 0 || (nfloats > 0 && !ws)) return NASREC_EINVAL;
    g_ws = nfloats > 0 ? ws : nullptr;
    g_ws_floats = nfloats;
    return 0;
}

int nasrec_seg_linear_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int n_off, int N,
                          const float* bias, float* C, int64_t ldc, int M, void* stream) {
    CHECK_ARG(segs_ok(segs, nseg) && W && C && M > 0 && N > 0 && n_off >= 0);
    Batch bt{};
    bt.nprob = 1;
    Prob& p = bt.prob[0];
    p.M = M;
    p.N = N;
    p.term0 = 0;
    p.c = C;
    p.c_hi_i = ldc;
    p.c_hi_j = 1;
    p.bias = bias ? bias + n_off : nullptr;
    p.nsplit = 1;
    int nt = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        Term& t = bt.term[nt++];
        t.K = (int)segs[s].width;
        t.a = plain_view(segs[s].ptr, segs[s].ld, 1, 1);
        t.b = plain_view(W + (long long)n_off * ldw + segs[s].w_off, ldw, 1, 1);
    }
    p.nterm = nt;
    return launch(bt, as_stream(stream));
}

int nasrec_seg_linear_dgrad(const float* dC, int64_t ldc, int N, const float* W, int64_t ldw, int n_off,
                            const nasrec_seg_t* dsegs, int nseg, int M, int accumulate, void* stream) {
    CHECK_ARG(segs_ok(dsegs, nseg) && dC && W && M > 0 && N > 0);
    Batch bt{};
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (dsegs[s].width == 0) continue;
        Prob& p = bt.prob[np];
        Term& t = bt.term[np];
        p.M = M;
        p.N = (int)dsegs[s].width;
        p.term0 = np;
        p.nterm = 1;
        p.c = const_cast<float*>(dsegs[s].ptr);
        p.c_hi_i = dsegs[s].ld;
        p.c_hi_j = 1;
        p.addend = accumulate ? dsegs[s].ptr : nullptr;
        p.nsplit = 1;
        t.K = N;
        t.a = plain_view(dC, ldc, 1, 1);
        // B(n = k-column of the segment, k = output row of W): W[(n_off+k)*ldw + w_off + n]
        t.b = plain_view(W + (long long)n_off * ldw + dsegs[s].w_off, 1, ldw, 0);
        ++np;
    }
    bt.nprob = np;
    return launch(bt, as_stream(stream));
}

int nasrec_seg_linear_wgrad(const float* dC, int64_t ldc, int N, const nasrec_seg_t* segs, int nseg, float* dW,
                            int64_t ldw, int n_off, int M, int accumulate, void* stream) {
    CHECK_ARG(segs_ok(segs, nseg) && dC && dW && M > 0 && N > 0);
    Batch bt{};
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        Prob& p = bt.prob[np];
        Term& t = bt.term[np];
        p.M = N;                       // rows of dW
        p.N = (int)segs[s].width;      // columns of dW in this segment
        p.term0 = np;
        p.nterm = 1;
        p.c = dW + (long long)n_off * ldw + segs[s].w_off;
        p.c_hi_i = ldw;
        p.c_hi_j = 1;
        p.addend = accumulate ? p.c : nullptr;
        p.nsplit = 1;
        t.K = M;                       // contraction over the batch
        t.a = plain_view(dC, 1, ldc, 0);                    // A(i = n_out, j = m) = dC[m*ldc + n_out]
        t.b = plain_view(segs[s].ptr, 1, segs[s].ld, 0);    // B(i = k, j = m)     = A_s[m*ld + k]
        ++np;
    }
    bt.nprob = np;
    return launch(bt, as_stream(stream));
}

// ---------------------------------------------------------------- 3-D (sparse axis)
int nasrec_sproj_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int P, const float* bias,
                     float* Z, int64_t z_bstride, int B, void* stream) {
    CHECK_ARG(segs_ok(segs, nseg) && W && Z && B > 0 && P > 0);
    Batch bt{};
    bt.nprob = 1;
    Prob& p = bt.prob[0];
    p.M = B * NASREC_EMB_DIM;          // m = b*16 + e
    p.N = P;
    p.term0 = 0;
    p.c = Z;                           // Z[b*zbs + p*16 + e]
    p.c_hi_i = z_bstride;
    p.c_lo_i = 1;
    p.c_sh_i = 4;
    p.c_hi_j = NASREC_EMB_DIM;
    p.bias = bias;
    p.nsplit = 1;
    int nt = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        Term& t = bt.term[nt++];
        t.K = (int)segs[s].width;
        View a{};                      // A(m=(b,e), k=r) = X[b*ld + r*16 + e]
        a.p = segs[s].ptr;
        a.hi_i = segs[s].ld;
        a.lo_i = 1;
        a.sh_i = 4;
        a.hi_j = NASREC_EMB_DIM;
        a.contig_j = 0;
        t.a = a;
        t.b = plain_view(W + segs[s].w_off, ldw, 1, 1);   // B(n=p, k=r) = W[p*ldw + w_off + r]
    }
    p.nterm = nt;
    return launch(bt, as_stream(stream));
}

int nasrec_sproj_dgrad(const float* dZ, int64_t dz_bstride, int P, const float* W, int64_t ldw,
                       const nasrec_seg_t* dsegs, int nseg, int B, int accumulate, void* stream) {
    CHECK_ARG(segs_ok(dsegs, nseg) && dZ && W && B > 0 && P > 0);
    Batch bt{};
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (dsegs[s].width == 0) continue;
        Prob& p = bt.prob[np];
        Term& t = bt.term[np];
        p.M = B * NASREC_EMB_DIM;
        p.N = (int)dsegs[s].width;     // n = r
        p.term0 = np;
        p.nterm = 1;
        p.c = const_cast<float*>(dsegs[s].ptr);   // dX[b*ld + r*16 + e]
        p.c_hi_i = dsegs[s].ld;
        p.c_lo_i = 1;
        p.c_sh_i = 4;
        p.c_hi_j = NASREC_EMB_DIM;
        p.addend = accumulate ? dsegs[s].ptr : nullptr;
        p.nsplit = 1;
        t.K = P;
        View a{};                      // A(m=(b,e), k=p) = dZ[b*zbs + p*16 + e]
        a.p = dZ;
        a.hi_i = dz_bstride;
        a.lo_i = 1;
        a.sh_i = 4;
        a.hi_j = NASREC_EMB_DIM;
        a.contig_j = 0;
        t.a = a;
        t.b = plain_view(W + dsegs[s].w_off, 1, ldw, 0);   // B(n=r, k=p) = W[p*ldw + w_off + r]
        ++np;
    }
    bt.nprob = np;
    return launch(bt, as_stream(stream));
}

static int sproj_nsplit(int B) {
    int ns = B / 32;
    if (ns < 1) ns = 1;
    if (ns > 32) ns = 32;
    return ns;
}

int64_t nasrec_sproj_wgrad_ws_floats(int P, int64_t total_width, int B) {
    return (int64_t)sproj_nsplit(B) * P * total_width;
}

int nasrec_sproj_wgrad(const float* dZ, int64_t dz_bstride, int P, const nasrec_seg_t* segs, int nseg, float* dW,
                       int64_t ldw, int B, int accumulate, float* ws, void* stream) {
    CHECK_ARG(segs_ok(segs, nseg) && dZ && dW && ws && B > 0 && P > 0);
    const int ns = sproj_nsplit(B);
    Batch bt{};
    RedBatch rb{};
    int np = 0;
    long long wsoff = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        const int w = (int)segs[s].width;
        Prob& p = bt.prob[np];
        Term& t = bt.term[np];
        p.M = P;
        p.N = w;
        p.term0 = np;
        p.nterm = 1;
        p.c = ws + wsoff;              // partial [split][P][w]
        p.c_hi_i = w;
        p.c_hi_j = 1;
        p.nsplit = ns;
        p.split_stride = (long long)P * w;
        t.K = B * NASREC_EMB_DIM;      // k = b*16 + e
        View a{};                      // A(i=p, k=(b,e)) = dZ[b*zbs + p*16 + e]
        a.p = dZ;
        a.hi_i = NASREC_EMB_DIM;
        a.hi_j = dz_bstride;
        a.lo_j = 1;
        a.sh_j = 4;
        a.contig_j = 1;
        t.a = a;
        View b{};                      // B(i=r, k=(b,e)) = X[b*ld + r*16 + e]
        b.p = segs[s].ptr;
        b.hi_i = NASREC_EMB_DIM;
        b.hi_j = segs[s].ld;
        b.lo_j = 1;
        b.sh_j = 4;
        b.contig_j = 1;
        t.b = b;
        rb.seg[np].ws = ws + wsoff;
        rb.seg[np].c = dW + segs[s].w_off;
        rb.seg[np].bias = nullptr;
        rb.seg[np].ldc = ldw;
        rb.seg[np].M = P;
        rb.seg[np].N = w;
        rb.seg[np].nsplit = ns;
        rb.seg[np].accumulate = accumulate;
        wsoff += (long long)ns * P * w;
        ++np;
    }
    bt.nprob = np;
    rb.nseg = np;
    if (np == 0) return 0;
    int rc = launch(bt, as_stream(stream));
    if (rc) return rc;
    return launch_reduce(rb, as_stream(stream));
}

}  // extern "C"
