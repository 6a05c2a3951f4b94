// Host emulation harness for the per-sample interaction kernels of nasrec_b200/csrc/interact.cu (CPU only): DotProduct's
// strict-lower-triangle kernels and the FM reductions against a direct double-precision evaluation of the reference
// formulas (nasrec/supernet/modules.py:366-383: Z = T T^T gathered at tril_indices(T, T, -1) in row-major order;
// :736-738: (sum_r x)^2 - sum_r x^2), forward and backward.
#include "cuda_emul.h"

#include "interact_device_code.inc"
}  // namespace (opened inside the include)

static int close(const char* what, const std::vector<float>& got, const std::vector<double>& want, double tol) {
    double mx = 0, ref = 1e-30;
    for (size_t i = 0; i < got.size(); ++i) {
        mx = std::fmax(mx, std::fabs(got[i] - want[i]));
        ref = std::fmax(ref, std::fabs(want[i]));
        if (!std::isfinite(got[i])) mx = 1e30;
    }
    std::printf("  %-10s %zu values, max abs err %.3g (scale %.3g)\n", what, got.size(), mx, ref);
    return mx <= tol * ref ? 0 : 1;
}

int main() {
    int failures = 0;
    // ---- DotProduct triangle: P + 1 tokens of width 16 (token 0 = the dense projection x, tokens 1..P = y)
    for (int P : {1, 2, 7, 45}) {
        const int B = 3, Tn = P + 1, NR = Tn * (Tn - 1) / 2, grid = 2;
        std::printf("dot_tril P=%d\n", P);
        const long long ldx = 16 + 8, ybs = (long long)P * 16 + 16, ldr = NR + 3;
        std::vector<float> x(B * ldx), y(B * ybs), R(B * ldr, 0.f), dR(B * ldr);
        for (auto& v : x) v = rnd() * 2.f;
        for (auto& v : y) v = rnd() * 2.f;
        for (auto& v : dR) v = rnd();
        run_grid(grid, 256, 0, [&] { dot_tril_fwd_kernel(x.data(), ldx, y.data(), ybs, P, R.data(), ldr, B); });
        auto tok = [&](int b, int i, int e) -> double { return i == 0 ? x[b * ldx + e] : y[b * ybs + (i - 1) * 16 + e]; };
        std::vector<float> got;
        std::vector<double> want;
        for (int b = 0; b < B; ++b) {
            int r = 0;
            for (int i = 1; i < Tn; ++i)
                for (int j = 0; j < i; ++j, ++r) {
                    double acc = 0;
                    for (int e = 0; e < 16; ++e) acc += tok(b, i, e) * tok(b, j, e);
                    want.push_back(acc);
                    got.push_back(R[b * ldr + r]);
                }
        }
        failures += close("R", got, want, 2e-6);
        std::vector<float> dx(B * 16, 0.f), dy(B * (long long)P * 16, 0.f);
        run_grid(grid, 256, 0, [&] {
            dot_tril_bwd_kernel(dR.data(), ldr, x.data(), ldx, y.data(), ybs, P, dx.data(), 16, dy.data(), (long long)P * 16, B);
        });
        got.clear();
        want.clear();
        for (int b = 0; b < B; ++b)
            for (int i = 0; i < Tn; ++i)
                for (int e = 0; e < 16; ++e) {
                    double acc = 0;
                    for (int j = 0; j < Tn; ++j) {
                        if (j == i) continue;
                        const int hi = i > j ? i : j, lo = i > j ? j : i;
                        acc += (double)dR[b * ldr + hi * (hi - 1) / 2 + lo] * tok(b, j, e);
                    }
                    want.push_back(acc);
                    got.push_back(i == 0 ? dx[b * 16 + e] : dy[(long long)b * P * 16 + (i - 1) * 16 + e]);
                }
        failures += close("dT", got, want, 2e-6);
    }
    // ---- FM: ix = (sum_r x)^2 - sum_r x^2 and its backward 2 g (s - x)
    for (int rows : {1, 26, 64, 72}) {
        const int B = 37;
        std::printf("fm rows=%d\n", rows);
        const long long xbs = (long long)rows * 16 + 16;
        std::vector<float> x(B * xbs), ix(B * 16, 0.f), g(B * 16), dx(B * (long long)rows * 16, 0.f);
        for (auto& v : x) v = rnd() * 2.f;
        for (auto& v : g) v = rnd();
        run_grid(cdiv(B * 16, 256), 256, 0, [&] { fm_fwd_kernel(x.data(), xbs, rows, ix.data(), B); });
        run_grid(5, 256, 0, [&] {
            fm_bwd_rows_kernel(g.data(), x.data(), xbs, rows, nullptr, 0, dx.data(), (long long)rows * 16, B);
        });
        std::vector<double> wix, wdx;
        for (int b = 0; b < B; ++b) {
            double s[16] = {0}, q[16] = {0};
            for (int r = 0; r < rows; ++r)
                for (int e = 0; e < 16; ++e) {
                    const double v = x[b * xbs + r * 16 + e];
                    s[e] += v;
                    q[e] += v * v;
                }
            for (int e = 0; e < 16; ++e) wix.push_back(s[e] * s[e] - q[e]);
            for (int r = 0; r < rows; ++r)
                for (int e = 0; e < 16; ++e) wdx.push_back(2.0 * g[b * 16 + e] * (s[e] - x[b * xbs + r * 16 + e]));
        }
        failures += close("ix", ix, wix, 4e-6);
        failures += close("dx", dx, wdx, 4e-6);
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
