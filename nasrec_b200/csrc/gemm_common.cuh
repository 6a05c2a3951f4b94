// Shared descriptors of the segment-list GEMM family (FFMA and tcgen05 paths).
#pragma once
#include "common.cuh"

namespace nasrec_gemm {


struct View {
    const float* p;
    long long hi_i, hi_j;
    int lo_i, lo_j, sh_i, sh_j;
    int contig_j;   // 1: consecutive j are adjacent in memory, 0: consecutive i are
    int vec16;      // contig_j and every 4-aligned j chunk of every row is 16-byte aligned
};

__device__ __forceinline__ long long voff(const View& v, int i, int j) {
    return (long long)(i >> v.sh_i) * v.hi_i + (long long)(i & ((1 << v.sh_i) - 1)) * v.lo_i +
           (long long)(j >> v.sh_j) * v.hi_j + (long long)(j & ((1 << v.sh_j) - 1)) * v.lo_j;
}

struct Term {
    View a;   // A(m, k): i = m, j = k
    View b;   // B(n, k): i = n, j = k
    int K;
    int pad_;
};

struct Prob {
    int M, N, term0, nterm;
    float* c;
    const float* addend;          // same layout as c, or null
    long long c_hi_i, c_hi_j;
    int c_lo_i, c_sh_i;
    const float* bias;            // indexed by n, or null
    int nsplit;
    int pad_;
    long long split_stride;
};

constexpr int MAXP = 16, MAXT = 20;
struct Batch {
    int nprob;
    int pad_;
    Prob prob[MAXP];
    Term term[MAXT];
};


}  // namespace nasrec_gemm
