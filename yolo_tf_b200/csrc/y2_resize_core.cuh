// Image resize on the detection path: detect.py:65 `_image.resize((width, height))` = Pillow's Image.resize with its default
// filter (BICUBIC since Pillow 7; NEAREST before).  The algorithm lives in Pillow (src/libImaging/Resample.c, Geometry.c);
// restated here for 8-bit channels.  Host + device code in one header so that the CPU harness (tests/host/resize_harness.cu)
// runs the very functions the kernels call.
//
//   bicubic (a = -0.5), antialiased when shrinking: per output coordinate a window [xmin, xmin + n) of the input and n weights
//   computed in double, normalised, converted to 22-bit fixed point; out = clip((2^21 + sum pixel * weight) >> 22, 0, 255);
//   horizontal pass into an 8-bit intermediate, then vertical; a pass that does not change the size is skipped.
//   nearest: source index = int(xo), xo = scale / 2 for the first coordinate, += scale for each next (a running double sum).
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

namespace y2 {

static constexpr int RESIZE_PRECISION_BITS = 32 - 8 - 2;
static constexpr int RESIZE_NEAREST = 0, RESIZE_BICUBIC = 3;      // PIL.Image.Resampling codes

static inline double resize_bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static inline int resize_ksize(int in_size, int out_size) {
    double scale = (double)((float)in_size - 0.0f) / out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(2.0 * scale) * 2 + 1;
}
// bounds[2 * xx] = first input index, bounds[2 * xx + 1] = tap count; kk[xx * ksize + x] = fixed-point weight (0 past the count)
static inline int resize_bicubic_tables(int in_size, int out_size, std::vector<int>* bounds, std::vector<int>* kk) {
    const double scale = (double)((float)in_size - 0.0f) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    const double ss = 1.0 / filterscale;
    bounds->assign((size_t)out_size * 2, 0);
    kk->assign((size_t)out_size * ksize, 0);
    std::vector<double> w(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = resize_bicubic_filter((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? w[x] / ww : w[x];
            (*kk)[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << RESIZE_PRECISION_BITS)) : (int)(0.5 + v * (1 << RESIZE_PRECISION_BITS));
        }
        (*bounds)[2 * xx] = xmin;
        (*bounds)[2 * xx + 1] = xmax;
    }
    return ksize;
}
static inline void resize_nearest_table(int in_size, int out_size, std::vector<int>* idx) {
    const double s = (double)in_size / out_size;
    double xo = s * 0.5;
    idx->assign(out_size, 0);
    for (int i = 0; i < out_size; ++i) {
        int v = (int)xo;
        if (v > in_size - 1) v = in_size - 1;
        (*idx)[i] = v;
        xo += s;
    }
}
// one output sample: n taps `stride` bytes apart
__host__ __device__ __forceinline__ uint8_t resize_tap(const uint8_t* p, long long stride, const int* k, int n) {
    int acc = 1 << (RESIZE_PRECISION_BITS - 1);
    for (int x = 0; x < n; ++x) acc += (int)p[(long long)x * stride] * k[x];
    const int v = acc >> RESIZE_PRECISION_BITS;            // arithmetic shift, as Pillow's clip8 lookup
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// element e of the horizontal pass output [in_h][out_w][C] / of the vertical pass output [out_h][out_w][C]
__host__ __device__ __forceinline__ uint8_t resize_h_element(const uint8_t* src, int in_w, int C, int out_w, const int* bounds, const int* kk,
                                                             int ksize, long long e) {
    const int c = (int)(e % C);
    const long long t = e / C;
    const int xx = (int)(t % out_w);
    const long long y = t / out_w;
    return resize_tap(src + (y * in_w + bounds[2 * xx]) * C + c, C, kk + (long long)xx * ksize, bounds[2 * xx + 1]);
}
__host__ __device__ __forceinline__ uint8_t resize_v_element(const uint8_t* src, int w, int C, const int* bounds, const int* kk, int ksize,
                                                             long long e) {
    const long long row = (long long)w * C;
    const long long col = e % row;
    const int yy = (int)(e / row);
    return resize_tap(src + (long long)bounds[2 * yy] * row + col, row, kk + (long long)yy * ksize, bounds[2 * yy + 1]);
}
__host__ __device__ __forceinline__ uint8_t resize_nearest_element(const uint8_t* src, int in_w, int C, int out_w, const int* xidx,
                                                                   const int* yidx, long long e) {
    const int c = (int)(e % C);
    const long long t = e / C;
    return src[((long long)yidx[t / out_w] * in_w + xidx[t % out_w]) * C + c];
}

}  // namespace y2
