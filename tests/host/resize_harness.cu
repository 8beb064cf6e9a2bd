// CPU check (no GPU) of the resize path: the table builders and per-element functions the kernels call
// (yolo_tf_b200/csrc/y2_resize_core.cuh), compiled for the host and driven in the kernels' element order.
// usage: resize_harness in.raw in_h in_w C out_h out_w resample out.raw
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../yolo_tf_b200/csrc/y2_resize_core.cuh"

int main(int argc, char** argv) {
    if (argc != 9) return 2;
    const int in_h = atoi(argv[2]), in_w = atoi(argv[3]), C = atoi(argv[4]), out_h = atoi(argv[5]), out_w = atoi(argv[6]), resample = atoi(argv[7]);
    std::vector<uint8_t> src((size_t)in_h * in_w * C), dst((size_t)out_h * out_w * C);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(src.data(), 1, src.size(), f) != src.size()) return 3;
    fclose(f);
    if (resample == y2::RESIZE_NEAREST) {
        std::vector<int> xi, yi;
        y2::resize_nearest_table(in_w, out_w, &xi);
        y2::resize_nearest_table(in_h, out_h, &yi);
        for (long long e = 0; e < (long long)dst.size(); ++e) dst[e] = y2::resize_nearest_element(src.data(), in_w, C, out_w, xi.data(), yi.data(), e);
    } else {
        const uint8_t* cur = src.data();
        std::vector<uint8_t> tmp;
        int w = in_w;
        if (in_w != out_w) {
            std::vector<int> b, k;
            const int ks = y2::resize_bicubic_tables(in_w, out_w, &b, &k);
            if (ks != y2::resize_ksize(in_w, out_w)) return 4;
            tmp.resize((size_t)in_h * out_w * C);
            for (long long e = 0; e < (long long)tmp.size(); ++e) tmp[e] = y2::resize_h_element(cur, in_w, C, out_w, b.data(), k.data(), ks, e);
            cur = tmp.data(); w = out_w;
        }
        if (in_h != out_h) {
            std::vector<int> b, k;
            const int ks = y2::resize_bicubic_tables(in_h, out_h, &b, &k);
            for (long long e = 0; e < (long long)dst.size(); ++e) dst[e] = y2::resize_v_element(cur, w, C, b.data(), k.data(), ks, e);
        } else {
            for (size_t e = 0; e < dst.size(); ++e) dst[e] = cur[e];
        }
    }
    f = fopen(argv[8], "wb");
    if (!f || fwrite(dst.data(), 1, dst.size(), f) != dst.size()) return 5;
    fclose(f);
    return 0;
}
