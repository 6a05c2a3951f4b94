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
#include <cstdio>
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"
#include <cstring>
#include <cstddef>
#include <vector>

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

// Short contractions (total K <= SK_KMAX: the 13 dense features, 16-wide FM / DotProduct projections, 26..64 sparse rows
// or projection channels).  Same contract as gemm64_kernel (Batch of strided views, any operand layout), but the WHOLE K
// range of a 64 x 64 output tile is fetched at once -- every load in flight before the first use, one barrier -- so the
// kernel is one global-memory latency long instead of one per 16-wide k-slab; ~9 KB of registers' worth of loads per
// thread, 35 KB of shared memory, no tensor-memory allocation, several CTAs per SM (a successor launched with
// programmatic dependent launch becomes resident while this one drains).  For these shapes the tensor-core kernels pay
// ~5 us per CTA before their first MMA, times the number of waves: an [8192 x 64 x 64] sparse-axis projection with seven
// gradient targets took 49 us there (profiles/r02_notes.md).
constexpr int SK_KMAX = 64;
__global__ void __launch_bounds__(256) gemm_smallk_kernel(const __grid_constant__ Batch bt) {
    pdl_trigger();
    __shared__ __align__(16) float As[SK_KMAX][BM + PADM];
    __shared__ __align__(16) float Bs[SK_KMAX][BN + PADM];
    __shared__ int kterm[SK_KMAX], kin[SK_KMAX];
    int pi = blockIdx.z;
    if (pi >= bt.nprob) return;
    const Prob& pr = bt.prob[pi];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= pr.M || n0 >= pr.N) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // k -> (term, k inside the term)
    int Ktot = 0;
    for (int t = 0; t < pr.nterm; ++t) Ktot += bt.term[pr.term0 + t].K;
    if (tid < SK_KMAX) {
        int k = tid, t = 0;
        while (t < pr.nterm && k >= bt.term[pr.term0 + t].K) { k -= bt.term[pr.term0 + t].K; ++t; }
        kterm[tid] = t < pr.nterm ? pr.term0 + t : -1;
        kin[tid] = k;
    }
    __syncthreads();
    pdl_wait();
    // operand tiles: thread -> (row i, column k) chosen so that a warp reads along the operand's contiguous axis
    const bool a_kc = bt.term[pr.term0].a.contig_j != 0, b_kc = bt.term[pr.term0].b.contig_j != 0;
    float va[16], vb[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        // 64 x 64 elements = 16 per thread.  k-contiguous: 16 consecutive k per row group (tid & 15 = k low); else tid & 63 = i
        const int ka = a_kc ? ((tid & 15) + 16 * (q & 3)) : ((tid >> 6) + 4 * q);
        const int ia = a_kc ? ((tid >> 4) + 16 * (q >> 2)) : (tid & 63);
        float v = 0.f;
        if (ka < Ktot && m0 + ia < pr.M) {
            const Term& tm = bt.term[kterm[ka]];
            v = __ldg(tm.a.p + voff(tm.a, m0 + ia, kin[ka]));
        }
        va[q] = v;
        const int kb = b_kc ? ((tid & 15) + 16 * (q & 3)) : ((tid >> 6) + 4 * q);
        const int ib = b_kc ? ((tid >> 4) + 16 * (q >> 2)) : (tid & 63);
        v = 0.f;
        if (kb < Ktot && n0 + ib < pr.N) {
            const Term& tm = bt.term[kterm[kb]];
            v = __ldg(tm.b.p + voff(tm.b, n0 + ib, kin[kb]));
        }
        vb[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int ka = a_kc ? ((tid & 15) + 16 * (q & 3)) : ((tid >> 6) + 4 * q);
        const int ia = a_kc ? ((tid >> 4) + 16 * (q >> 2)) : (tid & 63);
        As[ka][ia] = va[q];
        const int kb = b_kc ? ((tid & 15) + 16 * (q & 3)) : ((tid >> 6) + 4 * q);
        const int ib = b_kc ? ((tid >> 4) + 16 * (q >> 2)) : (tid & 63);
        Bs[kb][ib] = vb[q];
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const int kend = Ktot < SK_KMAX ? Ktot : SK_KMAX;
#pragma unroll 4
    for (int j = 0; j < kend; ++j) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[j][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[j][tx * 4]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
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
            pr.c[o] = v;
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

// ---- live GEMM accounting (nasrec_gemm_prof): CUDA events around every GEMM launch of the library on the launching
// stream, with the ALGORITHMIC flops of the launch (2 M N K over the live support) -- what bench.py's roofline reads.
struct GemmProf {
    bool on = false;
    std::vector<cudaEvent_t> e0, e1;
    std::vector<double> flops;
    struct Desc { int M, N, K, nprob, kind, bn, ns, tma; };
    std::vector<Desc> desc;                      // shape and plan of each timed launch (NASREC_GEMM_TRACE dump)
    size_t used = 0;
} g_prof;

template <class ProbT, class KOf>
double launch_flops(const ProbT* prob, int nprob, KOf k_of) {
    double f = 0;
    for (int p = 0; p < nprob; ++p) f += 2.0 * prob[p].M * prob[p].N * (double)k_of(p);
    return f;
}
struct ProfScope {
    bool active;
    cudaStream_t st;
    size_t slot = 0;
    ProfScope(cudaStream_t s, double fl) : active(g_prof.on), st(s) {
        if (!active) return;
        if (g_prof.used == g_prof.e0.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            g_prof.e0.push_back(a);
            g_prof.e1.push_back(b);
            g_prof.flops.push_back(0);
            g_prof.desc.push_back(GemmProf::Desc{});
        }
        slot = g_prof.used++;
        g_prof.flops[slot] = fl;
        cudaEventRecord(g_prof.e0[slot], st);
    }
    ~ProfScope() {
        if (active) cudaEventRecord(g_prof.e1[slot], st);
    }
};

// Contractions of at most g_small_k elements (one or two k-tiles: the 13 dense features, a 16-wide FM / DotProduct
// projection, 26..64 sparse rows) skip the tensor-core kernels: TMEM allocation, tensor maps and the mbarrier pipeline cost
// ~7 us per launch before the first MMA.  OFF by default (0): measured inside the B = 512 step the generic CUDA-core
// kernel below (scalar view loads) takes ~19 us per launch against ~13 us for the tensor-core kernels on the same
// problems (1.57 vs 1.49 ms per step, profiles/r02_notes.md); the switch stays for a kernel that earns it.  fp32-parity
// modes only: the bf16 mode keeps its operand rounding.  NASREC_SMALL_K / nasrec_set_small_k override.
int g_small_k = getenv("NASREC_SMALL_K") ? atoi(getenv("NASREC_SMALL_K")) : 0;
int g_gemm_mode = 3;   // 0: fp32 FFMA; 1/3/4: tcgen05 kind::tf32 with 1/3/4 split products (default 3xTF32); 2: bf16 operands

inline bool small_k_mode() { return g_small_k > 0 && g_gemm_mode >= 3; }
inline long long seg_total(const nasrec_seg_t* segs, int nseg) {
    long long k = 0;
    for (int s = 0; s < nseg; ++s) k += segs[s].width;
    return k;
}

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

const int* g_dgrad_flags = nullptr;   // see nasrec_internal_set_dgrad_flags (common.cuh)
int g_plan_kind = 0;    // operand-layout class of the entry point being served (set on entry): 0 fwd, 1 dgrad, 2 wgrad / sparse-axis

// Tile width + split-K plan of one launch, shared by both tensor-core paths (so that they sum in the same order and agree
// bit for bit): nasrec_gemm::tc_plan.  Split-K applies to plain row-major outputs (optionally accumulated in place).
template <class KTiles>
nasrec_gemm::TilePlan plan_launch(const Prob* prob, int nprob, KTiles ktiles_of, int maxN) {
    bool eligible = true;
    for (int p = 0; p < nprob; ++p) {
        const Prob& pr = prob[p];
        if (pr.nsplit != 1 || pr.c_sh_i != 0 || pr.c_hi_j != 1 || (pr.addend && pr.addend != pr.c)) eligible = false;
    }
    static const int kinds = getenv("NASREC_SPLIT_KINDS") ? atoi(getenv("NASREC_SPLIT_KINDS")) : 7;      // experiment knob
    if (!((kinds >> g_plan_kind) & 1)) eligible = false;
    // SM budget of backward launches (experiment knob; measured: planning the two backward streams for half of the SMs
    // each changes nothing on the B = 512 step, 1.566 vs 1.574 ms).  Must not depend on whether a side stream is attached:
    // the Python engine and the executor have to plan -- and round -- alike.
    static const int bwd_sms = getenv("NASREC_BWD_SMS") ? atoi(getenv("NASREC_BWD_SMS")) : nasrec_gemm::TC_SM_COUNT;
    const int budget = g_plan_kind != 0 ? bwd_sms : nasrec_gemm::TC_SM_COUNT;
    return nasrec_gemm::tc_plan(prob, nprob, maxN, ktiles_of, eligible, g_plan_kind, budget);
}

// LDG-producer kernel: the split CTAs write partial tiles to the library workspace and a fixed-order reduction launch
// follows.  Returns false (plan falls back to ns = 1) when the workspace cannot hold the partials.
bool split_to_workspace(Prob* prob, int nprob, int ns, cudaStream_t st, RedBatch& rb, int& totz) {
    float* wsb = nullptr;
    long long wsn = 0;
    ws_region(st, &wsb, &wsn);
    long long need_all = 0;
    for (int p = 0; p < nprob; ++p) need_all += (long long)ns * prob[p].M * prob[p].N;
    if (!wsb || need_all > wsn || rb.nseg + nprob > NASREC_MAX_SEGS) return false;
    long long off = 0;
    for (int p = 0; p < nprob; ++p) {
        Prob& pr = prob[p];
        const long long need = (long long)ns * pr.M * pr.N;
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
    return true;
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
    ProfScope prof(st, launch_flops(bt.prob, bt.nprob, [&](int p) {
        long long k = 0;
        for (int t = 0; t < bt.prob[p].nterm; ++t) k += bt.term[bt.prob[p].term0 + t].K;
        return k;
    }));
    bool small = small_k_mode();
    for (int p = 0; p < bt.nprob && small; ++p) {
        long long k = 0;
        for (int t = 0; t < bt.prob[p].nterm; ++t) k += bt.term[bt.prob[p].term0 + t].K;
        small = k <= g_small_k;
    }
    if (g_gemm_mode != 0 && !small) {
        RedBatch rb{};
        const nasrec_gemm::TilePlan pl = plan_launch(bt.prob, bt.nprob, [&](int p) {
            int kt = 0;
            for (int t = 0; t < bt.prob[p].nterm; ++t) kt += (bt.term[bt.prob[p].term0 + t].K + 31) / 32;
            return kt;
        }, maxN);
        if (prof.active)
            g_prof.desc[prof.slot] = GemmProf::Desc{bt.prob[0].M, bt.prob[0].N, 0, bt.nprob, g_plan_kind, pl.bn, pl.ns, 0};
        if (pl.ns > 1) split_to_workspace(bt.prob, bt.nprob, pl.ns, st, rb, totz);
        int rc = nasrec_gemm::launch_tc(bt, pl.bn, maxM, maxN, totz, g_gemm_mode, st);
        if (rc || rb.nseg == 0) return rc;
        return launch_reduce(rb, st);
    }
    dim3 grid(cdiv(maxN, BN), cdiv(maxM, BM), totz);
    if (grid.y > 65535 || grid.z > 65535) return NASREC_ETOOBIG;
    bool whole_k = small && totz == bt.nprob;          // short contraction, no caller-imposed split: one-shot kernel
    for (int p = 0; p < bt.nprob && whole_k; ++p) {
        long long k = 0;
        for (int t = 0; t < bt.prob[p].nterm; ++t) {
            k += bt.term[bt.prob[p].term0 + t].K;
            // one layout flag per operand and problem decides the thread -> element mapping
            if (bt.term[bt.prob[p].term0 + t].a.contig_j != bt.term[bt.prob[p].term0].a.contig_j ||
                bt.term[bt.prob[p].term0 + t].b.contig_j != bt.term[bt.prob[p].term0].b.contig_j) whole_k = false;
        }
        if (k > SK_KMAX) whole_k = false;
    }
    if (whole_k) nasrec_launch(gemm_smallk_kernel, grid, 256, 0, st, bt);
    else nasrec_launch(gemm64_kernel, grid, 256, 0, st, bt);
    return nasrec_launch_status();
}

// ---------------------------------------------------------------------------------------------- TMA path
// Host side of gemm_tma.cuh: which launches qualify, the tensor maps, and the per-launch operand layouts.
int g_use_tma = 1;

// Pre-split weight planes (nasrec_set_weight_planes): hi = rn_tf32(W), lo = W - hi, rows 16-byte aligned.
// The reference's state-dict layout keeps row strides such as 1037 floats, which a tensor map cannot describe;
// the planes are what forward and dgrad fetch by TMA instead (and they need no conversion in the kernel).
struct Planes {
    const float* W = nullptr;
    const float* hi = nullptr;
    const float* lo = nullptr;
    long long ldp = 0;
    int rows = 0, cols = 0;
    int first = 0, shift = 0;      // plane column of W column c: c + (c >= first ? shift : 0)
} g_pl;

// A TMA box must start on a 16-byte boundary of the innermost dimension, and the reference layout puts the second
// source of a concat at column nd (13) or F (26): the planes therefore shift every column >= `first` right by
// `shift` (zero pad columns in between) so that all segment starts are multiples of 4 floats.
inline long long plane_col(long long w_off) { return w_off + (w_off >= g_pl.first ? g_pl.shift : 0); }
inline bool plane_seg_ok(long long w_off, long long width) {
    if (width == 0) return true;
    if (w_off < g_pl.first && w_off + width > g_pl.first) return false;
    return (plane_col(w_off) & 3) == 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_encode_tried = false;

bool have_encoder() {
    if (!g_encode_tried) {
        g_encode_tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
        cudaGetLastError();
    }
    return g_encode != nullptr;
}

struct MapSpec {
    const float* base;
    uint64_t dim[4];
    uint64_t stride[3];     // bytes, dims 1..rank-1
    uint32_t box[4];
    uint32_t rank;
    uint32_t swz;           // 0: SWIZZLE_128B, 1: SWIZZLE_64B, 2: SWIZZLE_128B_ATOM_32B, 3: none
};
struct MapEnt {
    MapSpec key;
    CUtensorMap map;
    bool valid;
};
constexpr int MAP_CACHE = 4096;
MapEnt* g_map_cache = nullptr;
long long g_map_hits = 0, g_map_misses = 0, g_tma_launches = 0;

bool get_map(CUtensorMap* out, const MapSpec& sp) {
    if (!g_map_cache) g_map_cache = new MapEnt[MAP_CACHE]();
    uint64_t h = (uint64_t)(uintptr_t)sp.base * 0x9E3779B97F4A7C15ull;
    h ^= (sp.dim[0] * 31 + sp.dim[1]) * 0xC2B2AE3D27D4EB4Full + sp.dim[2] * 0x165667B19E3779F9ull + sp.stride[0] * 13 + sp.box[1] * 7 +
         sp.box[0] + sp.swz;
    MapEnt& e = g_map_cache[(h >> 20) & (MAP_CACHE - 1)];
    if (e.valid && std::memcmp(&e.key, &sp, sizeof(MapSpec)) == 0) {
        *out = e.map;
        ++g_map_hits;
        return true;
    }
    ++g_map_misses;
    cuuint64_t gdim[4] = {sp.dim[0], sp.dim[1], sp.dim[2], sp.dim[3]};
    cuuint64_t gstr[3] = {sp.stride[0], sp.stride[1], sp.stride[2]};
    cuuint32_t box[4] = {sp.box[0], sp.box[1], sp.box[2], sp.box[3]};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    // the kernel issues one instruction form (3-D): 2-D operands get a unit third dimension
    cuuint32_t rank = sp.rank;
    if (rank == 2) {
        rank = 3;
        gdim[2] = 1;
        box[2] = 1;
        gstr[1] = gstr[0] * gdim[1];
    }
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, (void*)sp.base, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sp.swz == 0 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (sp.swz == 1 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : (sp.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    e.key = sp;
    e.map = m;
    e.valid = true;
    *out = m;
    return true;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

using nasrec_gemm::OpLayout;
using nasrec_gemm::TBatch;
using nasrec_gemm::TTerm;
using nasrec_gemm::OP_KM128;
using nasrec_gemm::OP_MN128;
using nasrec_gemm::OP_MN3;
using nasrec_gemm::OP_KM64;

OpLayout make_layout(int kind, int rows, int convert) {
    OpLayout L{};
    L.kind = kind;
    L.convert = convert;
    const uint32_t ver = 1u << 14;
    for (int d = 0; d < 4; ++d) L.rsh[d] = L.ksh[d] = 31;
    switch (kind) {
    case OP_KM128:
        L.rank = 2; L.nbox = 1; L.box_dim = 0; L.box_bytes = rows * 128;
        L.rsh[1] = 0; L.ksh[0] = 0;
        L.desc_hi32 = (1024u >> 4) | ver | (2u << 29); L.desc_lbo = 1;
        for (int j = 0; j < 4; ++j) L.koff[j] = 32 * j;
        L.tile_bytes = rows * 128; L.mn_major = 0;
        break;
    case OP_MN128:      // kind::tf32 accepts MN-major operands only in the 128-byte swizzle with 32-byte atoms: K atom = 4 rows
        L.rank = 2; L.nbox = rows >= 32 ? rows / 32 : 1; L.box_dim = 0; L.box_bytes = 4096;
        L.rsh[0] = 0; L.ksh[1] = 0;
        L.desc_hi32 = (512u >> 4) | ver | (1u << 29); L.desc_lbo = 4096u >> 4;
        for (int j = 0; j < 4; ++j) L.koff[j] = 1024 * j;
        L.tile_bytes = L.nbox * 4096; L.mn_major = 1;
        break;
    case OP_MN3:        // A only; rows = (b, e): one plain box {16 e, 32 k, 8 b}, read by the converters as it lands
        L.rank = 3; L.nbox = 1; L.box_dim = 0; L.box_bytes = 16384;
        L.rsh[2] = 4; L.ksh[1] = 0;
        L.tile_bytes = 16384; L.mn_major = 1;
        break;
    default:            // OP_KM64: k = (b, e): a 32-wide k-tile is 2 samples; one box {16 e, rows, 2 b}
        L.rank = 3; L.nbox = 1; L.box_dim = 0; L.box_bytes = rows * 128;
        L.rsh[1] = 0; L.ksh[2] = 4;
        L.desc_hi32 = (512u >> 4) | ver | (4u << 29); L.desc_lbo = 1;
        L.koff[0] = 0; L.koff[1] = 32; L.koff[2] = rows * 64; L.koff[3] = rows * 64 + 32;
        L.tile_bytes = rows * 128; L.mn_major = 0;
        break;
    }
    return L;
}

void set_box(MapSpec& sp, int kind, int rows) {
    sp.swz = kind == OP_KM128 ? 0 : (kind == OP_KM64 ? 1 : (kind == OP_MN128 ? 2 : 3));
    sp.box[3] = 1;
    switch (kind) {
    case OP_KM128: sp.box[0] = 32; sp.box[1] = (uint32_t)rows; sp.box[2] = 1; break;
    case OP_MN128: sp.box[0] = 32; sp.box[1] = 32; sp.box[2] = 1; break;
    case OP_MN3: sp.box[0] = 16; sp.box[1] = 32; sp.box[2] = 8; break;
    default: sp.box[0] = 16; sp.box[1] = (uint32_t)rows; sp.box[2] = 2; break;
    }
}

MapSpec spec2d(const float* base, uint64_t inner, uint64_t outer, long long ld) {
    MapSpec sp{};
    sp.base = base; sp.rank = 2;
    sp.dim[0] = inner; sp.dim[1] = outer; sp.dim[2] = 1; sp.dim[3] = 1;
    sp.stride[0] = (uint64_t)ld * 4;
    return sp;
}
MapSpec spec3d(const float* base, uint64_t rows, uint64_t B, long long bstride) {      // [B, rows, 16] tensor as (e, row, b)
    MapSpec sp{};
    sp.base = base; sp.rank = 3;
    sp.dim[0] = NASREC_EMB_DIM; sp.dim[1] = rows; sp.dim[2] = B; sp.dim[3] = 1;
    sp.stride[0] = NASREC_EMB_DIM * 4; sp.stride[1] = (uint64_t)bstride * 4;
    return sp;
}
// Work description filled by the entry points before the tile width is known.
template <class TB, int MAPS>
struct TmaJobT {
    TB tb;
    MapSpec spec[MAPS];
    int side[MAPS];     // 0: A operand map, 1: B operand map
    int nmap = 0;
    int a_kind = 0, b_kind = 0, a_conv = 1, b_conv = 0;
    int add(const MapSpec& sp, int which) {
        spec[nmap] = sp;
        side[nmap] = which;
        return nmap++;
    }
    void reset() {
        nmap = 0;
        std::memset(&tb.nprob, 0, sizeof(TB) - offsetof(TB, nprob));
    }
};
using TmaJob = TmaJobT<TBatch, nasrec_gemm::TM_MAXMAPS>;
using TmaJobBig = TmaJobT<nasrec_gemm::TBatchBig, nasrec_gemm::BIG_MAPS>;
TmaJob* g_job = nullptr;       // reusable jobs (host-side scratch; the library is single-threaded per process)
TmaJobBig* g_big_job = nullptr;

TmaJob& fresh_job() {
    if (!g_job) g_job = new TmaJob();
    g_job->reset();
    return *g_job;
}
TmaJobBig& fresh_big_job() {
    if (!g_big_job) g_big_job = new TmaJobBig();
    g_big_job->reset();
    return *g_big_job;
}

template <class Job>
int run_tma(Job& job, cudaStream_t st, RedBatch* extra_rb = nullptr) {
    auto& tb = job.tb;
    int maxM = 0, maxN = 0, totz = 0;
    for (int i = 0; i < tb.nprob; ++i) {
        maxM = tb.prob[i].M > maxM ? tb.prob[i].M : maxM;
        maxN = tb.prob[i].N > maxN ? tb.prob[i].N : maxN;
        totz += tb.prob[i].nsplit;
    }
    if (maxM <= 0 || maxN <= 0 || totz <= 0) return 0;
    ProfScope prof(st, launch_flops(tb.prob, tb.nprob, [&](int p) {
        long long k = 0;
        for (int t = 0; t < tb.prob[p].nterm; ++t) k += tb.term[tb.prob[p].term0 + t].K;
        return k;
    }));
    const nasrec_gemm::TilePlan pl = plan_launch(tb.prob, tb.nprob, [&](int p) {
        int kt = 0;
        for (int t = 0; t < tb.prob[p].nterm; ++t) kt += (tb.term[tb.prob[p].term0 + t].K + 31) / 32;
        return kt;
    }, maxN);
    const int bn = pl.bn;
    if (prof.active) {
        int k = 0;
        for (int t = 0; t < tb.prob[0].nterm; ++t) k += tb.term[tb.prob[0].term0 + t].K;
        g_prof.desc[prof.slot] = GemmProf::Desc{tb.prob[0].M, maxN, k, tb.nprob, g_plan_kind, pl.bn, pl.ns, 1};
    }
    tb.cluster_ns = pl.ns;             // split-K inside thread-block clusters (DSMEM reduction in the kernel)
    if (pl.ns > 1) totz *= pl.ns;      // eligible launches have nsplit == 1 everywhere: z = problem * ns + split
    tb.nprod = g_gemm_mode;
    static const int dbg = getenv("NASREC_GEMM_DBG") ? atoi(getenv("NASREC_GEMM_DBG")) : 0;
    tb.dbg = dbg;
    tb.la = make_layout(job.a_kind, nasrec_gemm::TC_BM, job.a_conv);
    tb.lb = make_layout(job.b_kind, bn, job.b_conv);
    for (int i = 0; i < job.nmap; ++i) {
        set_box(job.spec[i], job.side[i] ? job.b_kind : job.a_kind, job.side[i] ? bn : nasrec_gemm::TC_BM);
        if (!get_map(&tb.maps[i], job.spec[i])) return NASREC_EINVAL;
    }
    dim3 grid((maxN + bn - 1) / bn, (maxM + nasrec_gemm::TC_BM - 1) / nasrec_gemm::TC_BM, totz);
    if (tb.nprob > 1 && !tb.flat) {
        // several problems of different sizes (gradient targets of a dgrad, segments of a wgrad): enumerate the real tiles
        // instead of launching the bounding box of the largest problem for each of them
        bool plain = true;
        for (int i = 0; i < tb.nprob; ++i) plain = plain && tb.prob[i].nsplit == 1;
        if (plain) tb.flat = 1;
    }
    if (tb.flat) {
        // one CTA (or cluster) per REAL output tile: running tile totals per problem; flat launches carry no pre-split problems
        int tot = 0;
        for (int i = 0; i < tb.nprob; ++i) {
            tot += ((tb.prob[i].M + nasrec_gemm::TC_BM - 1) / nasrec_gemm::TC_BM) * ((tb.prob[i].N + bn - 1) / bn);
            tb.tile_end[i] = tot;
        }
        grid = dim3((unsigned)tot, 1, (unsigned)(pl.ns > 1 ? pl.ns : 1));
    }
    int rc;
    switch (bn) {
        case 16: rc = nasrec_gemm::launch_tma_bn<16>(tb, grid, st); break;
        case 32: rc = nasrec_gemm::launch_tma_bn<32>(tb, grid, st); break;
        case 64: rc = nasrec_gemm::launch_tma_bn<64>(tb, grid, st); break;
        default: rc = nasrec_gemm::launch_tma_bn<128>(tb, grid, st); break;
    }
    if (rc) return rc;
    ++g_tma_launches;
    if (extra_rb && extra_rb->nseg) rc = launch_reduce(*extra_rb, st);
    return rc;
}

inline bool tma_on() {
    static const bool off = getenv("NASREC_FORCE_NO_TMA") != nullptr;      // experiment knob (the API switch is nasrec_set_gemm_tma)
    return !off && g_use_tma && g_gemm_mode != 0 && have_encoder();
}
inline bool planes_for(const float* W, long long ldw) {
    return g_pl.W == W && g_pl.hi && (g_gemm_mode <= 2 || g_pl.lo) && g_pl.cols == (int)ldw;
}

constexpr int NOT_TMA = -1000;     // "this launch does not qualify": the caller takes the LDG-producer kernel

int tma_seg_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int n_off, int N, const float* bias,
                float* C, int64_t ldc, int M, cudaStream_t st) {
    if (!tma_on() || !planes_for(W, ldw)) return NOT_TMA;
    int live = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        if (!al16(segs[s].ptr) || (segs[s].ld & 3) || !plane_seg_ok(segs[s].w_off, segs[s].width)) return NOT_TMA;
        ++live;
    }
    if (live + 2 > nasrec_gemm::TM_MAXMAPS || live > MAXT) return NOT_TMA;
    TmaJob& job = fresh_job();
    job.a_kind = OP_KM128; job.a_conv = 1; job.b_kind = OP_KM128; job.b_conv = 0;
    const int bh = job.add(spec2d(g_pl.hi, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1);
    const int bl = g_gemm_mode > 2 ? job.add(spec2d(g_pl.lo, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1) : bh;
    TBatch& tb = job.tb;
    tb.nprob = 1;
    Prob& p = tb.prob[0];
    p.M = M; p.N = N; p.term0 = 0;
    p.c = C; p.c_hi_i = ldc; p.c_hi_j = 1;
    p.bias = bias ? bias + n_off : nullptr;
    p.nsplit = 1;
    int nt = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        TTerm& t = tb.term[nt++];
        t.K = (int)segs[s].width;
        t.a_hi = t.a_lo = (short)job.add(spec2d(segs[s].ptr, segs[s].width, M, segs[s].ld), 0);
        t.b_hi = (short)bh; t.b_lo = (short)bl;
        t.b_base[0] = (int)plane_col(segs[s].w_off); t.b_base[1] = n_off;
    }
    p.nterm = nt;
    return run_tma(job, st);
}

int tma_seg_dgrad(const float* dC, int64_t ldc, int N, const float* W, int64_t ldw, int n_off, const nasrec_seg_t* dsegs,
                  int nseg, int M, int accumulate, cudaStream_t st) {
    if (!tma_on() || !planes_for(W, ldw) || !al16(dC) || (ldc & 3)) return NOT_TMA;
    TmaJob& job = fresh_job();
    job.a_kind = OP_KM128; job.a_conv = 1; job.b_kind = OP_MN128; job.b_conv = 0;
    const int ah = job.add(spec2d(dC, N, M, ldc), 0);
    const int bh = job.add(spec2d(g_pl.hi, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1);
    const int bl = g_gemm_mode > 2 ? job.add(spec2d(g_pl.lo, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1) : bh;
    TBatch& tb = job.tb;
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (dsegs[s].width == 0) continue;
        if (np >= MAXP || !plane_seg_ok(dsegs[s].w_off, dsegs[s].width)) return NOT_TMA;
        Prob& p = tb.prob[np];
        TTerm& t = tb.term[np];
        p.M = M; p.N = (int)dsegs[s].width; p.term0 = np; p.nterm = 1;
        p.c = const_cast<float*>(dsegs[s].ptr); p.c_hi_i = dsegs[s].ld; p.c_hi_j = 1;
        p.addend = (g_dgrad_flags ? g_dgrad_flags[s] : accumulate) ? dsegs[s].ptr : nullptr;
        p.nsplit = 1;
        t.K = N;
        t.a_hi = t.a_lo = (short)ah;
        t.b_hi = (short)bh; t.b_lo = (short)bl;
        t.b_base[0] = (int)plane_col(dsegs[s].w_off); t.b_base[1] = n_off;     // B(n = column of the segment, k = output row)
        ++np;
    }
    tb.nprob = np;
    return run_tma(job, st);
}

// ---- deferred weight gradients (nasrec_wgrad_defer / nasrec_wgrad_flush) -------------------------------------------------
// dW of a linear is read only by the optimizer.  Launched one by one behind their LayerNorm backward, the ~13 dense
// weight-gradient GEMMs of a B = 512 step each pay the ~6 us launch floor, fill the SMs unevenly (one problem per
// launch) and -- on the side stream -- take SMs from the dY -> dX chain exactly while it is the critical path
// (event-timed launches in the step are ~1.6x their isolated duration).  Deferred, they wait in a queue (operands are
// step-lifetime buffers of the caller) and run as ONE grid over all their output tiles when the backward pass is done:
// same kernel, one launch floor, full waves, nothing competing with the chain.
struct PendingWgrad {
    const float* dC;
    int64_t ldc;
    int N;
    nasrec_seg_t segs[NASREC_MAX_SEGS];
    int nseg;
    float* dW;
    int64_t ldw;
    int n_off, M;
};
std::vector<PendingWgrad> g_wq;
bool g_wdefer = false;

int wgrad_flush(cudaStream_t st) {
    size_t i = 0;
    const int saved_kind = g_plan_kind;
    g_plan_kind = 2;
    int rc = 0;
    while (i < g_wq.size() && rc == 0) {
        TmaJobBig& job = fresh_big_job();
        job.a_kind = OP_MN128; job.a_conv = 1; job.b_kind = OP_MN128; job.b_conv = 1;
        auto& tb = job.tb;
        int np = 0;
        const int M = g_wq[i].M;            // a batch shares the contraction length (the batch size of the step)
        for (; i < g_wq.size(); ++i) {
            const PendingWgrad& w = g_wq[i];
            if (w.M != M || np + w.nseg > nasrec_gemm::BIG_P || job.nmap + 1 + w.nseg > nasrec_gemm::BIG_MAPS) break;
            const int ah = job.add(spec2d(w.dC, w.N, w.M, w.ldc), 0);
            for (int s = 0; s < w.nseg; ++s) {
                Prob& p = tb.prob[np];
                TTerm& t = tb.term[np];
                p.M = w.N; p.N = (int)w.segs[s].width; p.term0 = np; p.nterm = 1;
                p.c = w.dW + (long long)w.n_off * w.ldw + w.segs[s].w_off; p.c_hi_i = w.ldw; p.c_hi_j = 1;
                p.addend = nullptr;
                p.nsplit = 1;
                t.K = w.M;
                t.a_hi = t.a_lo = (short)ah;
                t.b_hi = t.b_lo = (short)job.add(spec2d(w.segs[s].ptr, w.segs[s].width, w.M, w.segs[s].ld), 1);
                ++np;
            }
        }
        tb.nprob = np;
        tb.flat = 1;
        rc = np ? run_tma(job, st) : NASREC_EINVAL;
    }
    g_wq.clear();
    g_plan_kind = saved_kind;
    return rc;
}

int tma_seg_wgrad(const float* dC, int64_t ldc, int N, const nasrec_seg_t* segs, int nseg, float* dW, int64_t ldw, int n_off,
                  int M, int accumulate, cudaStream_t st) {
    if (!tma_on() || !al16(dC) || (ldc & 3)) return NOT_TMA;
    int live = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        if (!al16(segs[s].ptr) || (segs[s].ld & 3)) return NOT_TMA;
        ++live;
    }
    if (live + 1 > nasrec_gemm::TM_MAXMAPS || live > MAXP) return NOT_TMA;
    if (g_wdefer && live > 0) {
        // deferred: the weight gradient is consumed only by the optimizer, so it joins the batched launch of wgrad_flush.
        // An in-place accumulate may depend on a queued writer of the same rows: drain the queue first, then run it now.
        if (accumulate) {
            const int rc = wgrad_flush(st);
            if (rc) return rc;
        } else {
            PendingWgrad pw{};
            pw.dC = dC; pw.ldc = ldc; pw.N = N; pw.dW = dW; pw.ldw = ldw; pw.n_off = n_off; pw.M = M;
            for (int s = 0; s < nseg; ++s)
                if (segs[s].width > 0) pw.segs[pw.nseg++] = segs[s];
            g_wq.push_back(pw);
            return 0;
        }
    }
    TmaJob& job = fresh_job();
    job.a_kind = OP_MN128; job.a_conv = 1; job.b_kind = OP_MN128; job.b_conv = 1;
    const int ah = job.add(spec2d(dC, N, M, ldc), 0);          // A(i = output row, k = sample) = dC[k*ldc + i]
    TBatch& tb = job.tb;
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        Prob& p = tb.prob[np];
        TTerm& t = tb.term[np];
        p.M = N; p.N = (int)segs[s].width; p.term0 = np; p.nterm = 1;
        p.c = dW + (long long)n_off * ldw + segs[s].w_off; p.c_hi_i = ldw; p.c_hi_j = 1;
        p.addend = accumulate ? p.c : nullptr;
        p.nsplit = 1;
        t.K = M;
        t.a_hi = t.a_lo = (short)ah;
        t.b_hi = t.b_lo = (short)job.add(spec2d(segs[s].ptr, segs[s].width, M, segs[s].ld), 1);
        ++np;
    }
    tb.nprob = np;
    return run_tma(job, st);
}

int tma_sproj_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int P, const float* bias, float* Z,
                  int64_t z_bstride, int B, cudaStream_t st) {
    if (!tma_on() || !planes_for(W, ldw)) return NOT_TMA;
    int live = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        if (!al16(segs[s].ptr) || (segs[s].ld & 3) || !plane_seg_ok(segs[s].w_off, segs[s].width)) return NOT_TMA;
        ++live;
    }
    if (live + 2 > nasrec_gemm::TM_MAXMAPS || live > MAXT) return NOT_TMA;
    TmaJob& job = fresh_job();
    job.a_kind = OP_MN3; job.a_conv = 1; job.b_kind = OP_KM128; job.b_conv = 0;
    const int bh = job.add(spec2d(g_pl.hi, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1);
    const int bl = g_gemm_mode > 2 ? job.add(spec2d(g_pl.lo, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1) : bh;
    TBatch& tb = job.tb;
    tb.nprob = 1;
    Prob& p = tb.prob[0];
    p.M = B * NASREC_EMB_DIM; p.N = P; p.term0 = 0;
    p.c = Z; p.c_hi_i = z_bstride; p.c_lo_i = 1; p.c_sh_i = 4; p.c_hi_j = NASREC_EMB_DIM;
    p.bias = bias;
    p.nsplit = 1;
    int nt = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        TTerm& t = tb.term[nt++];
        t.K = (int)segs[s].width;
        t.a_hi = t.a_lo = (short)job.add(spec3d(segs[s].ptr, segs[s].width, B, segs[s].ld), 0);
        t.b_hi = (short)bh; t.b_lo = (short)bl;
        t.b_base[0] = (int)plane_col(segs[s].w_off); t.b_base[1] = 0;
    }
    p.nterm = nt;
    return run_tma(job, st);
}

int tma_sproj_dgrad(const float* dZ, int64_t dz_bstride, int P, const float* W, int64_t ldw, const nasrec_seg_t* dsegs,
                    int nseg, int B, int accumulate, cudaStream_t st) {
    if (!tma_on() || !planes_for(W, ldw) || !al16(dZ) || (dz_bstride & 3)) return NOT_TMA;
    TmaJob& job = fresh_job();
    job.a_kind = OP_MN3; job.a_conv = 1; job.b_kind = OP_MN128; job.b_conv = 0;
    const int ah = job.add(spec3d(dZ, P, B, dz_bstride), 0);
    const int bh = job.add(spec2d(g_pl.hi, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1);
    const int bl = g_gemm_mode > 2 ? job.add(spec2d(g_pl.lo, g_pl.cols + g_pl.shift, g_pl.rows, g_pl.ldp), 1) : bh;
    TBatch& tb = job.tb;
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (dsegs[s].width == 0) continue;
        if (np >= MAXP || !plane_seg_ok(dsegs[s].w_off, dsegs[s].width)) return NOT_TMA;
        Prob& p = tb.prob[np];
        TTerm& t = tb.term[np];
        p.M = B * NASREC_EMB_DIM; p.N = (int)dsegs[s].width; p.term0 = np; p.nterm = 1;
        p.c = const_cast<float*>(dsegs[s].ptr);
        p.c_hi_i = dsegs[s].ld; p.c_lo_i = 1; p.c_sh_i = 4; p.c_hi_j = NASREC_EMB_DIM;
        p.addend = (g_dgrad_flags ? g_dgrad_flags[s] : accumulate) ? dsegs[s].ptr : nullptr;
        p.nsplit = 1;
        t.K = P;
        t.a_hi = t.a_lo = (short)ah;
        t.b_hi = (short)bh; t.b_lo = (short)bl;
        t.b_base[0] = (int)plane_col(dsegs[s].w_off); t.b_base[1] = 0;        // B(n = r, k = p) = W[p*ldw + w_off + r]
        ++np;
    }
    tb.nprob = np;
    return run_tma(job, st);
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

void nasrec_internal_set_dgrad_flags(const int* flags) { g_dgrad_flags = flags; }
cudaStream_t nasrec_internal_side_stream() { return g_side; }
void nasrec_internal_set_side_stream(cudaStream_t s) { g_side = s; }

extern "C" {

int nasrec_set_gemm_mode(int mode) {
    if (mode < 0 || mode > 4) return NASREC_EINVAL;
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
    g_plan_kind = 0;
    CHECK_ARG(segs_ok(segs, nseg) && W && C && M > 0 && N > 0 && n_off >= 0);
    if (!(small_k_mode() && seg_total(segs, nseg) <= g_small_k)) {
        const int rc = tma_seg_fwd(segs, nseg, W, ldw, n_off, N, bias, C, ldc, M, as_stream(stream));
        if (rc != NOT_TMA) return rc;
    }
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
    g_plan_kind = 1;
    CHECK_ARG(segs_ok(dsegs, nseg) && dC && W && M > 0 && N > 0);
    if (!(small_k_mode() && N <= g_small_k)) {
        const int rc = tma_seg_dgrad(dC, ldc, N, W, ldw, n_off, dsegs, nseg, M, accumulate, as_stream(stream));
        if (rc != NOT_TMA) return rc;
    }
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
        p.addend = (g_dgrad_flags ? g_dgrad_flags[s] : accumulate) ? dsegs[s].ptr : nullptr;
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
    g_plan_kind = 2;
    CHECK_ARG(segs_ok(segs, nseg) && dC && dW && M > 0 && N > 0);
    {
        const int rc = tma_seg_wgrad(dC, ldc, N, segs, nseg, dW, ldw, n_off, M, accumulate, as_stream(stream));
        if (rc != NOT_TMA) return rc;
    }
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
    g_plan_kind = 2;
    CHECK_ARG(segs_ok(segs, nseg) && W && Z && B > 0 && P > 0);
    if (!(small_k_mode() && seg_total(segs, nseg) <= g_small_k)) {
        const int rc = tma_sproj_fwd(segs, nseg, W, ldw, P, bias, Z, z_bstride, B, as_stream(stream));
        if (rc != NOT_TMA) return rc;
    }
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
    g_plan_kind = 2;
    CHECK_ARG(segs_ok(dsegs, nseg) && dZ && W && B > 0 && P > 0);
    if (!(small_k_mode() && P <= g_small_k)) {
        const int rc = tma_sproj_dgrad(dZ, dz_bstride, P, W, ldw, dsegs, nseg, B, accumulate, as_stream(stream));
        if (rc != NOT_TMA) return rc;
    }
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
        p.addend = (g_dgrad_flags ? g_dgrad_flags[s] : accumulate) ? dsegs[s].ptr : nullptr;
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

// Large batches (B > 1024: more than 64 k-tiles per CTA even at 8 splits): up to 32 splits through the caller's workspace
// and a reduction launch, which keeps the chain of MMAs per accumulator short (numerics, see gemm_tc.cuh).
static int sproj_wgrad_presplit(const float* dZ, int64_t dz_bstride, int P, const nasrec_seg_t* segs, int nseg, float* dW,
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
    bool tma_ok = tma_on() && al16(dZ) && !(dz_bstride & 3) && np + 1 <= nasrec_gemm::TM_MAXMAPS;
    for (int s = 0; s < nseg && tma_ok; ++s)
        if (segs[s].width > 0 && (!al16(segs[s].ptr) || (segs[s].ld & 3))) tma_ok = false;
    if (tma_ok) {
        TmaJob& job = fresh_job();
        job.a_kind = OP_KM64; job.a_conv = 1; job.b_kind = OP_KM64; job.b_conv = 1;
        const int ah = job.add(spec3d(dZ, P, B, dz_bstride), 0);
        int q = 0;
        for (int s = 0; s < nseg; ++s) {
            if (segs[s].width == 0) continue;
            job.tb.prob[q] = bt.prob[q];
            TTerm& t = job.tb.term[q];
            t.K = B * NASREC_EMB_DIM;
            t.a_hi = t.a_lo = (short)ah;
            t.b_hi = t.b_lo = (short)job.add(spec3d(segs[s].ptr, segs[s].width, B, segs[s].ld), 1);
            ++q;
        }
        job.tb.nprob = np;
        return run_tma(job, as_stream(stream), &rb);
    }
    int rc = launch(bt, as_stream(stream));
    if (rc) return rc;
    return launch_reduce(rb, as_stream(stream));
}

int nasrec_sproj_wgrad(const float* dZ, int64_t dz_bstride, int P, const nasrec_seg_t* segs, int nseg, float* dW,
                       int64_t ldw, int B, int accumulate, float* ws, void* stream) {
    // K = (b, e) is long (16 B) and the output small (P x rows): the launch is split over K like any other skinny GEMM --
    // thread-block clusters with a DSMEM reduction in the TMA kernel, library workspace + reduction launch in the
    // LDG-producer kernel (plan_launch).  `ws` is used only by the large-batch path above.
    g_plan_kind = 2;
    static const bool old_path = getenv("NASREC_SPROJ_WGRAD_OLD") != nullptr;      // experiment knob
    if (B > 1024 || old_path) return sproj_wgrad_presplit(dZ, dz_bstride, P, segs, nseg, dW, ldw, B, accumulate, ws, stream);
    CHECK_ARG(segs_ok(segs, nseg) && dZ && dW && B > 0 && P > 0);
    Batch bt{};
    int np = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].width == 0) continue;
        const int w = (int)segs[s].width;
        Prob& p = bt.prob[np];
        Term& t = bt.term[np];
        p.M = P;
        p.N = w;
        p.term0 = np;
        p.nterm = 1;
        p.c = dW + segs[s].w_off;
        p.c_hi_i = ldw;
        p.c_hi_j = 1;
        p.addend = accumulate ? p.c : nullptr;
        p.nsplit = 1;
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
        ++np;
    }
    bt.nprob = np;
    if (np == 0) return 0;
    bool tma_ok = tma_on() && al16(dZ) && !(dz_bstride & 3) && np + 1 <= nasrec_gemm::TM_MAXMAPS;
    for (int s = 0; s < nseg && tma_ok; ++s)
        if (segs[s].width > 0 && (!al16(segs[s].ptr) || (segs[s].ld & 3))) tma_ok = false;
    if (tma_ok) {
        TmaJob& job = fresh_job();
        job.a_kind = OP_KM64; job.a_conv = 1; job.b_kind = OP_KM64; job.b_conv = 1;
        const int ah = job.add(spec3d(dZ, P, B, dz_bstride), 0);
        int q = 0;
        for (int s = 0; s < nseg; ++s) {
            if (segs[s].width == 0) continue;
            job.tb.prob[q] = bt.prob[q];
            TTerm& t = job.tb.term[q];
            t.K = B * NASREC_EMB_DIM;
            t.a_hi = t.a_lo = (short)ah;
            t.b_hi = t.b_lo = (short)job.add(spec3d(segs[s].ptr, segs[s].width, B, segs[s].ld), 1);
            ++q;
        }
        job.tb.nprob = np;
        return run_tma(job, as_stream(stream));
    }
    return launch(bt, as_stream(stream));
}

int nasrec_gemm_prof(int what, double* out3) {
    if (what == 1) { g_prof.on = true; g_prof.used = 0; return 0; }
    g_prof.on = false;
    if (what == 2 && out3) {            // synchronises: total ms, launches, algorithmic flops since the start
        double ms = 0, fl = 0;
        for (size_t i = 0; i < g_prof.used; ++i) {
            cudaEventSynchronize(g_prof.e1[i]);
            float t = 0;
            if (cudaEventElapsedTime(&t, g_prof.e0[i], g_prof.e1[i]) == cudaSuccess) ms += t;
            fl += g_prof.flops[i];
        }
        out3[0] = ms; out3[1] = (double)g_prof.used; out3[2] = fl;
        if (const char* path = getenv("NASREC_GEMM_TRACE")) {       // one line per timed launch: shape, plan, microseconds
            if (FILE* f = fopen(path, "a")) {
                for (size_t i = 0; i < g_prof.used; ++i) {
                    float t = 0;
                    cudaEventElapsedTime(&t, g_prof.e0[i], g_prof.e1[i]);
                    const GemmProf::Desc& d = g_prof.desc[i];
                    fprintf(f, "%d %d %d %d %d %d %d %d %.2f %.0f\n", d.kind, d.M, d.N, d.K, d.nprob, d.bn, d.ns, d.tma, t * 1e3, g_prof.flops[i]);
                }
                fclose(f);
            }
        }
    }
    return 0;
}

int nasrec_gemm_plan(int kind, int M, int N, int K, int nprob, int* bn, int* ns) {
    // host-only: the (tile width, split-K) the planner picks for `nprob` plain row-major problems of this shape
    if (kind < 0 || kind > 2 || M <= 0 || N <= 0 || K <= 0 || nprob <= 0 || nprob > MAXP || !bn || !ns) return NASREC_EINVAL;
    Prob prob[MAXP] = {};
    for (int p = 0; p < nprob; ++p) {
        prob[p].M = M;
        prob[p].N = N;
        prob[p].nsplit = 1;
        prob[p].c_hi_j = 1;
    }
    const int saved = g_plan_kind;
    g_plan_kind = kind;
    const nasrec_gemm::TilePlan pl = plan_launch(prob, nprob, [&](int) { return (K + 31) / 32; }, N);
    g_plan_kind = saved;
    *bn = pl.bn;
    *ns = pl.ns;
    return 0;
}

int nasrec_wgrad_defer(int on) {
    const int old = g_wdefer ? 1 : 0;
    g_wdefer = on != 0;
    if (!g_wdefer) g_wq.clear();          // switching off drops what is still queued: flush first
    return old;
}

int nasrec_wgrad_flush(void* stream) {
    int rc = nasrec_internal_ln_flush(as_stream(stream));      // the other deferred parameter gradients (ln.cu)
    if (rc || g_wq.empty()) return rc;
    return wgrad_flush(as_stream(stream));
}

int64_t nasrec_wgrad_pending(void) { return (int64_t)g_wq.size() + nasrec_internal_ln_pending(); }

int nasrec_set_small_k(int k) {
    const int old = g_small_k;
    g_small_k = k < 0 ? 0 : k;
    return old;
}

int nasrec_set_gemm_tma(int on) {
    g_use_tma = on ? 1 : 0;
    return 0;
}

int nasrec_set_weight_planes(const float* W, const float* hi, const float* lo, int64_t ldp, int rows, int cols, int first) {
    const int shift = (4 - (first & 3)) & 3;
    if (W && (!hi || !al16(hi) || (lo && !al16(lo)) || (ldp & 3) || ldp < cols + shift || rows <= 0 || cols <= 0 || first < 0 ||
              first > cols))
        return NASREC_EINVAL;
    g_pl.W = W; g_pl.hi = hi; g_pl.lo = lo; g_pl.ldp = ldp; g_pl.rows = rows; g_pl.cols = cols;
    g_pl.first = first; g_pl.shift = shift;
    return 0;
}

int64_t nasrec_tensor_map_stats(int which) { return which == 0 ? g_map_hits : (which == 1 ? g_map_misses : g_tma_launches); }

}  // extern "C"
