// CPU check (no GPU) of the operand arithmetic of the EXPERIMENTAL mixed-kind conv: the very functions the prep kernels
// run (yolo_tf_b200/csrc/y2_mix_prep.cuh), compiled for the host.  For operand sets spanning 24 binades of amax it checks
//   * E16 F16 == E8 F8 4096 and ra = 32, rw = 128 exactly (one accumulator for the three products),
//   * no stored value saturates or overflows its format,
//   * sum_k [X16 W16 + X8 RW8 + RX8 W8] * unscale reproduces sum_k x w to the budget the plan rests on: a K = 2304 dot
//     product to ~1.5e-5 of the typical size of such a sum (median; < 1e-4 always), at least 8x better than the fp16 term
//     alone (~3e-4 median, > 5e-4 worst), and single products to 2^-13.
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <vector>

#include "../../yolo_tf_b200/csrc/y2_mix_prep.cuh"

static double dec8(uint8_t v) { return (double)__half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v, __NV_E4M3))); }
static double nrand() {          // Box-Muller on rand()
    const double u = (rand() + 1.0) / (RAND_MAX + 2.0), v = (rand() + 1.0) / (RAND_MAX + 2.0);
    return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
}

int main() {
    int bad = 0;
    srand(5);
    double worst_prod = 0, worst_dot = 0, worst_dot16 = 0, sum_dot = 0, sum_dot16 = 0;
    int trials = 0;
    for (int trial = 0; trial < 60 && !bad; ++trial) {
        const double sx = std::ldexp(1.0, (trial % 25) - 12) * (1.0 + 0.37 * (trial % 3)), sw = std::ldexp(1.0, -(trial % 11)) * 0.013;
        const int K = 2304;
        std::vector<float> x(K), w(K);
        float ax = 0, aw = 0;
        for (int k = 0; k < K; ++k) {
            double a = nrand(); a = a > 0.1 * a ? a : 0.1 * a;      // leaky-shaped activations
            x[k] = (float)(a * sx); w[k] = (float)(nrand() * sw);
            ax = std::fmax(ax, std::fabs(x[k])); aw = std::fmax(aw, std::fabs(w[k]));
        }
        const y2::MixScales s = y2::mix_scales(ax, aw);
        if (s.E16 * s.F16 != s.E8 * s.F8 * 4096.f || s.ra != 32.f || s.rw != 128.f || s.unscale * s.E16 * s.F16 != 1.0f) {
            printf("FAIL: scales inconsistent (E16 %g E8 %g F16 %g F8 %g ra %g rw %g)\n", s.E16, s.E8, s.F16, s.F8, s.ra, s.rw); bad = 1; break;
        }
        double exact = 0, approx = 0, main_only = 0, sumabs = 0;
        for (int k = 0; k < K; ++k) {
            __half x16, w16; uint8_t x8, rx8, w8, rw8;
            y2::mix_split(x[k], s.E16, s.E8, s.ra, &x16, &x8, &rx8);
            y2::mix_split(w[k], s.F16, s.F8, s.rw, &w16, &w8, &rw8);
            const double X16 = __half2float(x16), W16 = __half2float(w16);
            if (!std::isfinite(X16) || !std::isfinite(W16) || std::fabs(X16) > 32768.0 || std::fabs(W16) > 8192.0 ||
                std::fabs(dec8(x8)) > 256.0 || std::fabs(dec8(w8)) > 256.0 || std::fabs(dec8(rx8)) > 256.0 || std::fabs(dec8(rw8)) > 256.0) {
                printf("FAIL: stored value out of its budgeted range (trial %d k %d)\n", trial, k); bad = 1; break;
            }
            const double p = (X16 * W16 + dec8(x8) * dec8(rw8) + dec8(rx8) * dec8(w8)) * (double)s.unscale;
            const double t = (double)x[k] * (double)w[k];
            exact += t; approx += p; main_only += X16 * W16 * (double)s.unscale; sumabs += std::fabs(t);
            // single products of operands that matter (within 2^-6 of their tensor's amax): 2^-13 relative
            if (std::fabs(x[k]) > ax / 64 && std::fabs(w[k]) > aw / 64) {
                const double e = std::fabs(p - t) / std::fabs(t);
                if (e > worst_prod) worst_prod = e;
            }
        }
        const double scale = sumabs / std::sqrt((double)K);          // typical magnitude of a K-term sum of these products
        worst_dot = std::fmax(worst_dot, std::fabs(approx - exact) / scale);
        worst_dot16 = std::fmax(worst_dot16, std::fabs(main_only - exact) / scale);
        sum_dot += std::fabs(approx - exact) / scale; sum_dot16 += std::fabs(main_only - exact) / scale; ++trials;
    }
    printf("worst single product rel err %.3e (budget 2^-13 = 1.22e-4)\n", worst_prod);
    printf("K=2304 dot product err / scale: all three terms mean %.3e worst %.3e, fp16 term alone mean %.3e worst %.3e\n",
           sum_dot / trials, worst_dot, sum_dot16 / trials, worst_dot16);
    if (worst_prod > 1.2207e-4) { printf("FAIL: single product\n"); bad = 1; }
    if (worst_dot > 1e-4 || sum_dot / trials > 2.5e-5) { printf("FAIL: dot product\n"); bad = 1; }
    if (sum_dot16 < 8 * sum_dot || worst_dot16 < 5e-4) { printf("FAIL: the fp16 term alone should be >= 8x worse (harness not discriminating)\n"); bad = 1; }
    printf(bad ? "MIX PREP CHECK FAILED\n" : "MIX PREP CHECK OK\n");
    return bad;
}
