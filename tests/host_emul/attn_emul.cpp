// Host emulation harness for nasrec_b200/csrc/attn.cu (test infrastructure, CPU only).
//
// The device code of attn.cu (everything inside its anonymous namespace, cut out by tests/test_kernels_host_emul.py into
// attn_device_code.inc) is compiled as plain C++: one OS thread per CUDA thread of a CTA, __syncthreads() = a pthread
// barrier, shared memory = process-global arrays, CTAs run one after the other.  The harness runs the
// one-thread-per-token kernels (attn_fwd_kernel / attn_bwd_kernel, validated against the oracle on the GPU) and the
// four-threads-per-token kernels (attn_fwd4_kernel / attn_bwd4_kernel) on the same random inputs and requires
// bit-identical y, dX and per-CTA parameter-gradient partials: the two are meant to perform the same floating-point
// operations in the same order per output.  Build with -ffp-contract=off so the host compiler fuses nothing by itself.
#define NASREC_ATTN_PARAMS 1696
#include "cuda_emul.h"

#include "attn_device_code.inc"
}  // namespace (opened inside the include)

// argv[1] (optional): file that receives case 2 as raw float32: params | x | dy | y | dx | dparams summed over CTAs
int main(int argc, char** argv) {
    struct Case {
        int B, L, s, grid;
        bool with_dx;
    };
    const Case cases[] = {{3, 7, 7, 3, true},   {5, 64, 64, 2, true}, {4, 33, 20, 3, true}, {2, 1, 1, 2, true},
                          {3, 26, 26, 2, false}, {2, 64, 1, 1, true},  {3, 40, 39, 3, true}};
    int failures = 0;
    for (const Case& c : cases) {
        std::printf("case B=%d L=%d s_live=%d grid=%d dx=%d\n", c.B, c.L, c.s, c.grid, (int)c.with_dx);
        std::vector<float> params(NPARAM);
        for (auto& v : params) v = rnd() * 0.8f;
        for (int e = 0; e < E; ++e) {
            params[LN1_W + e] += 1.0f;
            params[LN2_W + e] += 1.0f;
        }
        AttnPtrs ap;
        for (int k = 0; k < NPTR; ++k) ap.p[k] = params.data() + kOff[k];
        const long long xbs = (long long)c.s * E, ybs = xbs + 16, dxbs = xbs;       // y / dy rows live in a wider buffer
        std::vector<float> x(c.B * xbs), dy(c.B * ybs);
        for (auto& v : x) v = rnd() * 2.f;
        for (auto& v : dy) v = rnd();
        std::vector<float> y1(c.B * ybs, 7.f), y4(c.B * ybs, 7.f);
        run_grid(c.grid, LMAX, 0, [&] { attn_fwd_kernel(x.data(), xbs, c.L, c.s, ap, y1.data(), ybs, c.B); });
        run_grid(c.grid, NT4, 0, [&] { attn_fwd4_kernel(x.data(), xbs, c.L, c.s, ap, y4.data(), ybs, c.B); });
        failures += check("y", y1, y4);
        std::vector<float> dx1(c.B * dxbs, 3.f), dx4(c.B * dxbs, 3.f), ws1((size_t)c.grid * NPARAM, 9.f),
            ws4((size_t)c.grid * NPARAM, 9.f);
        run_grid(c.grid, LMAX, S_TOTAL, [&] {
            attn_bwd_kernel(dy.data(), ybs, x.data(), xbs, c.L, c.s, ap, c.with_dx ? dx1.data() : nullptr, dxbs,
                            ws1.data(), c.B);
        });
        run_grid(c.grid, NT4, S_TOTAL, [&] {
            attn_bwd4_kernel(dy.data(), ybs, x.data(), xbs, c.L, c.s, ap, c.with_dx ? dx4.data() : nullptr, dxbs,
                             ws4.data(), c.B);
        });
        failures += check("dx", dx1, dx4);
        failures += check("dparam", ws1, ws4);
        if (argc > 1 && &c == &cases[2]) {
            std::vector<float> dsum(NPARAM, 0.f);
            for (int g = 0; g < c.grid; ++g)
                for (int i = 0; i < NPARAM; ++i) dsum[i] += ws4[(size_t)g * NPARAM + i];
            FILE* f = std::fopen(argv[1], "wb");
            if (!f) return 2;
            const int hdr[4] = {c.B, c.L, c.s, (int)ybs};
            std::fwrite(hdr, sizeof(int), 4, f);
            for (const std::vector<float>* v : {&params, &x, &dy, &y4, &dx4, &dsum}) std::fwrite(v->data(), 4, v->size(), f);
            std::fclose(f);
        }
        // the partials must not be trivially zero
        double nz = 0;
        for (float v : ws4) nz += std::fabs(v);
        if (!(nz > 1e-3)) {
            std::printf("  parameter gradients are all zero\n");
            ++failures;
        }
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
