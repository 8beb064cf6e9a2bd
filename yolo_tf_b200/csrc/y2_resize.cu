// Image resize on the device: detect.py:65 `_image.resize((width, height))`, i.e. Pillow's
// Image.resize with its default filter, 8-bit channels, bit for bit (y2_resize_core.cuh holds the algorithm and is verified
// on the CPU against Pillow itself: tests/host/resize_harness.cu).  HBM-bound: (in + out) bytes plus the 8-bit intermediate.
// UNVERIFIED ON A GPU at the time of writing (round-1 GPU budget spent): the kernels only wrap the per-element functions the
// CPU harness runs; tests/test_gpu_unverified.py holds the GPU test (Y2_EXPERIMENTAL=1).
#include <string.h>

#include "../../include/yolo2_b200.h"
#include "y2_internal.h"
#include "y2_resize_core.cuh"

namespace y2 {

__global__ void resize_h_kernel(const uint8_t* __restrict__ src, int in_w, int C, int out_w, const int* __restrict__ bounds,
                                const int* __restrict__ kk, int ksize, uint8_t* __restrict__ dst, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        dst[e] = resize_h_element(src, in_w, C, out_w, bounds, kk, ksize, e);
}
__global__ void resize_v_kernel(const uint8_t* __restrict__ src, int w, int C, const int* __restrict__ bounds, const int* __restrict__ kk,
                                int ksize, uint8_t* __restrict__ dst, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        dst[e] = resize_v_element(src, w, C, bounds, kk, ksize, e);
}
__global__ void resize_nearest_kernel(const uint8_t* __restrict__ src, int in_w, int C, int out_w, const int* __restrict__ xidx,
                                      const int* __restrict__ yidx, uint8_t* __restrict__ dst, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        dst[e] = resize_nearest_element(src, in_w, C, out_w, xidx, yidx, e);
}

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
struct ResizeLayout { size_t bh, kh, bv, kv, tmp, total; int ksize_h, ksize_v; };
static ResizeLayout resize_layout(int in_h, int in_w, int out_h, int out_w, int C, int resample) {
    ResizeLayout l;
    memset(&l, 0, sizeof(l));
    size_t off = 0;
    if (resample == RESIZE_NEAREST) {
        l.bh = off; off = up256(off + (size_t)out_w * sizeof(int));       // x index table
        l.bv = off; off = up256(off + (size_t)out_h * sizeof(int));       // y index table
    } else {
        l.ksize_h = resize_ksize(in_w, out_w); l.ksize_v = resize_ksize(in_h, out_h);
        l.bh = off; off = up256(off + (size_t)out_w * 2 * sizeof(int));
        l.kh = off; off = up256(off + (size_t)out_w * l.ksize_h * sizeof(int));
        l.bv = off; off = up256(off + (size_t)out_h * 2 * sizeof(int));
        l.kv = off; off = up256(off + (size_t)out_h * l.ksize_v * sizeof(int));
        l.tmp = off; off = up256(off + (size_t)in_h * out_w * C);          // horizontal pass output [in_h][out_w][C]
    }
    l.total = off;
    return l;
}
static inline int grid_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace y2

using namespace y2;

extern "C" {

size_t y2_resize_workspace_bytes(int in_h, int in_w, int out_h, int out_w, int channels, int resample) {
    if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0 || channels <= 0 || (resample != RESIZE_NEAREST && resample != RESIZE_BICUBIC)) return 0;
    return resize_layout(in_h, in_w, out_h, out_w, channels, resample).total;
}

int y2_resize_u8(const uint8_t* src, int in_h, int in_w, int channels, uint8_t* dst, int out_h, int out_w, int resample, void* ws,
                 size_t ws_bytes, void* stream) {
    Y2_REQUIRE(src && dst && ws, "y2_resize_u8: null argument");
    Y2_REQUIRE(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && channels > 0, "y2_resize_u8: bad shape %dx%dx%d -> %dx%d", in_h, in_w, channels,
               out_h, out_w);
    Y2_REQUIRE(resample == RESIZE_NEAREST || resample == RESIZE_BICUBIC, "y2_resize_u8: resample must be 0 (NEAREST) or 3 (BICUBIC), got %d", resample);
    const ResizeLayout l = resize_layout(in_h, in_w, out_h, out_w, channels, resample);
    Y2_REQUIRE(ws_bytes >= l.total && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, "y2_resize_u8: workspace too small or not 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(ws);
    const int C = channels;
    const long long out_total = (long long)out_h * out_w * C;
    // the tables are a few KiB of host memory: pageable copies are staged before cudaMemcpyAsync returns
    if (resample == RESIZE_NEAREST) {
        std::vector<int> xi, yi;
        resize_nearest_table(in_w, out_w, &xi);
        resize_nearest_table(in_h, out_h, &yi);
        Y2_CUDA(cudaMemcpyAsync(base + l.bh, xi.data(), xi.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        Y2_CUDA(cudaMemcpyAsync(base + l.bv, yi.data(), yi.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        resize_nearest_kernel<<<grid_for(out_total), 256, 0, s>>>(src, in_w, C, out_w, reinterpret_cast<const int*>(base + l.bh),
                                                                  reinterpret_cast<const int*>(base + l.bv), dst, out_total);
        Y2_CUDA(cudaGetLastError());
        note_launch();
        return 0;
    }
    const uint8_t* cur = src;
    int w = in_w;
    if (in_w != out_w) {
        std::vector<int> b, k;
        const int ks = resize_bicubic_tables(in_w, out_w, &b, &k);
        Y2_REQUIRE(ks == l.ksize_h, "y2_resize_u8: internal: horizontal tap count mismatch");
        Y2_CUDA(cudaMemcpyAsync(base + l.bh, b.data(), b.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        Y2_CUDA(cudaMemcpyAsync(base + l.kh, k.data(), k.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        uint8_t* out = (in_h != out_h) ? reinterpret_cast<uint8_t*>(base + l.tmp) : dst;
        const long long total = (long long)in_h * out_w * C;
        resize_h_kernel<<<grid_for(total), 256, 0, s>>>(cur, in_w, C, out_w, reinterpret_cast<const int*>(base + l.bh),
                                                        reinterpret_cast<const int*>(base + l.kh), ks, out, total);
        Y2_CUDA(cudaGetLastError());
        note_launch();
        cur = out; w = out_w;
    }
    if (in_h != out_h) {
        std::vector<int> b, k;
        const int ks = resize_bicubic_tables(in_h, out_h, &b, &k);
        Y2_REQUIRE(ks == l.ksize_v, "y2_resize_u8: internal: vertical tap count mismatch");
        Y2_CUDA(cudaMemcpyAsync(base + l.bv, b.data(), b.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        Y2_CUDA(cudaMemcpyAsync(base + l.kv, k.data(), k.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        resize_v_kernel<<<grid_for(out_total), 256, 0, s>>>(cur, w, C, reinterpret_cast<const int*>(base + l.bv),
                                                            reinterpret_cast<const int*>(base + l.kv), ks, dst, out_total);
        Y2_CUDA(cudaGetLastError());
        note_launch();
    } else if (cur == src) {                         // same size: Pillow returns a copy
        Y2_CUDA(cudaMemcpyAsync(dst, src, (size_t)out_total, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

}  // extern "C"
