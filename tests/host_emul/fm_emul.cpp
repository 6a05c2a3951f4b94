// Host emulation harness for the FactorizationMachine3D backward kernels of nasrec_b200/csrc/interact.cu (CPU only):
// fm_bwd_rows_kernel (one CTA per sample) must equal fm_bwd_kernel (one thread per column) bit for bit, with and
// without the in-place accumulate, over ragged row counts and several samples per CTA.
#include "cuda_emul.h"

#include "interact_device_code.inc"
}  // namespace (opened inside the include)

int main() {
    struct Case {
        int B, rows, grid;
        int mode;          // 0: fresh, 1: accumulate from another buffer, 2: accumulate in place
    };
    const Case cases[] = {{5, 64, 5, 0}, {7, 27, 3, 1}, {4, 72, 2, 2}, {33, 1, 4, 0}, {3, 17, 3, 2}, {2, 16, 1, 1}, {3, 128, 2, 2}};
    int failures = 0;
    for (const Case& c : cases) {
        std::printf("case B=%d rows=%d grid=%d mode=%d\n", c.B, c.rows, c.grid, c.mode);
        const long long xbs = (long long)c.rows * 16 + 32, dbs = (long long)c.rows * 16;
        std::vector<float> x(c.B * xbs), dix(c.B * 16), din(c.B * dbs);
        for (auto& v : x) v = rnd() * 3.f;
        for (auto& v : dix) v = rnd();
        for (auto& v : din) v = rnd();
        std::vector<float> a(c.B * dbs, 5.f), b(c.B * dbs, 5.f);
        if (c.mode == 2) a = b = din;
        const float* ina = c.mode == 0 ? nullptr : (c.mode == 1 ? din.data() : a.data());
        const float* inb = c.mode == 0 ? nullptr : (c.mode == 1 ? din.data() : b.data());
        run_grid(cdiv(c.B * 16, 256), 256, 0,
                 [&] { fm_bwd_kernel(dix.data(), x.data(), xbs, c.rows, ina, dbs, a.data(), dbs, c.B); });
        run_grid(c.grid, 256, 0,
                 [&] { fm_bwd_rows_kernel(dix.data(), x.data(), xbs, c.rows, inb, dbs, b.data(), dbs, c.B); });
        failures += check("dx", a, b);
        double nz = 0;
        for (float v : b) nz += std::fabs(v - 5.f);
        if (!(nz > 1e-3)) ++failures;
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
