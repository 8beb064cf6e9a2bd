// Host build of the NMS float arithmetic the kernels use (yolo_tf_b200/csrc/y2_nms_iou.cuh), driven by tests/test_nms_iou_host.py.
// in : int32 n, float thr, n x 8 floats (box a: xmin ymin xmax ymax, box b: the same)
// out: n x 4 uint32: bits of iou_ref(a, b), iou_hit(quick filter on), iou_hit(filter off), ford(area of a)
#include <stdio.h>
#include <stdlib.h>

#include "../../yolo_tf_b200/csrc/y2_nms_iou.cuh"

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int n = 0;
    float thr = 0.f;
    if (fread(&n, 4, 1, f) != 1 || fread(&thr, 4, 1, f) != 1) return 4;
    float* in = (float*)malloc((size_t)n * 8 * sizeof(float));
    if (fread(in, sizeof(float), (size_t)n * 8, f) != (size_t)n * 8) return 5;
    fclose(f);
    uint32_t* out = (uint32_t*)malloc((size_t)n * 4 * sizeof(uint32_t));
    const float thr_lo = 0.999f * thr;
    for (int i = 0; i < n; ++i) {
        const float* p = in + (size_t)i * 8;
        const float4 a = make_float4(p[0], p[1], p[2], p[3]), b = make_float4(p[4], p[5], p[6], p[7]);
        const float v = y2::iou_ref(a, b);
        memcpy(&out[4 * i], &v, 4);
        out[4 * i + 1] = y2::iou_hit(a, y2::box_area(a), b, y2::box_area(b), thr, thr_lo, thr > 0.0f) ? 1u : 0u;
        out[4 * i + 2] = y2::iou_hit(a, y2::box_area(a), b, y2::box_area(b), thr, thr_lo, false) ? 1u : 0u;
        out[4 * i + 3] = y2::ford(y2::box_area(a));
    }
    f = fopen(argv[2], "wb");
    if (!f) return 6;
    fwrite(out, sizeof(uint32_t), (size_t)n * 4, f);
    fclose(f);
    printf("NMS IOU HARNESS OK %d\n", n);
    return 0;
}
