// Steps either side of the detection path:
//   per_image_standardization  <- utils/preprocess.py:23-25 (detect.py:62: applied to the resized uint8 image cast to float32)
//       out = (x - mean(x)) / max(std(x), 1/sqrt(n)),  mean / population std over ALL n = H*W*3 elements of one image
//   detections                 <- detect.py:72-87: per box index = argmax_c conf (first maximum), kept iff conf[index] > threshold,
//       box scaled from cell units to pixels (xy_min * scale, (xy_max - xy_min) * scale)
// Both are HBM-bound passes; reductions are two-stage with fp64 partials and a fixed-order finish (deterministic).
#include <limits.h>

#include "y2_internal.h"

namespace y2 {

static constexpr int STD_CHUNK = 16384;

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p, size_t i);
template <> __device__ __forceinline__ float ld_as_float<unsigned char>(const unsigned char* p, size_t i) { return (float)__ldg(p + i); }
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p, size_t i) { return __ldg(p + i); }

// PASS 0: partial[b][chunk] = sum x ; PASS 1: partial[b][chunk] = sum (x - mean[b])^2
template <typename T, int PASS>
__global__ void __launch_bounds__(256)
std_partial_kernel(const T* __restrict__ x, size_t n, const float* __restrict__ mean, double* __restrict__ partial) {
    const int b = blockIdx.y, chunks = gridDim.x;
    const size_t lo = (size_t)blockIdx.x * STD_CHUNK, hi = lo + STD_CHUNK < n ? lo + STD_CHUNK : n;
    const T* src = x + (size_t)b * n;
    const float mu = PASS == 1 ? mean[b] : 0.f;
    float s = 0.f;                                           // <= 64 terms per thread in fp32, then fp64
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
        const float v = ld_as_float<T>(src, i);
        if (PASS == 0) s += v; else { const float d = v - mu; s = fmaf(d, d, s); }
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = (double)s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)b * chunks + blockIdx.x] = sm[0];
}

// one warp per image: fixed-order sum of the chunk partials.  PASS 0 -> mean ; PASS 1 -> denom = max(std, 1/sqrt(n))
template <int PASS>
__global__ void std_finish_kernel(const double* __restrict__ partial, int chunks, int B, double n, float* __restrict__ mean,
                                  float* __restrict__ denom) {
    const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    double s = 0.0;
    for (int c = lane; c < chunks; c += 32) s += partial[(size_t)b * chunks + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (PASS == 0) mean[b] = (float)(s / n);
        else denom[b] = fmaxf((float)sqrt(s / n), (float)(1.0 / sqrt(n)));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
std_apply_kernel(const T* __restrict__ x, size_t n, const float* __restrict__ mean, const float* __restrict__ denom, float* __restrict__ out) {
    const int b = blockIdx.y;
    const float mu = mean[b], d = denom[b];
    const T* src = x + (size_t)b * n;
    float* dst = out + (size_t)b * n;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        dst[i] = __fdiv_rn(ld_as_float<T>(src, i) - mu, d);
}

// ---- uint8 fast path (what detect.py:59-65 produces): sum and sum of squares of bytes are EXACT integers, so one pass gives
// the mean and the population variance (var = S2/n - mean^2 in float64: 53 bits hold both terms exactly) -- no second pass over
// the image; 16 pixels per 128-bit load.  partial[b][chunk] = {S1, S2} as doubles (exact below 2^53).
__global__ void __launch_bounds__(256)
std_u8_sums_kernel(const unsigned char* __restrict__ x, size_t n, double* __restrict__ partial) {
    const int b = blockIdx.y, chunks = gridDim.x;
    const size_t lo = (size_t)blockIdx.x * STD_CHUNK, hi = lo + STD_CHUNK < n ? lo + STD_CHUNK : n;
    const unsigned char* src = x + (size_t)b * n;
    unsigned int s1 = 0, s2 = 0;                             // <= 64 bytes per thread: 64 * 255^2 < 2^32
    if ((((size_t)b * n) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        for (size_t i = lo + (size_t)threadIdx.x * 16; i + 16 <= hi; i += 256 * 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
            const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { const unsigned int p = (w[k] >> (8 * j)) & 0xffu; s1 += p; s2 += p * p; }
            }
        }
        const size_t tail = lo + ((hi - lo) / 16) * 16;      // STD_CHUNK % 16 == 0: only the image's last chunk has one
        for (size_t i = tail + threadIdx.x; i < hi; i += 256) { const unsigned int p = __ldg(src + i); s1 += p; s2 += p * p; }
    } else {
        for (size_t i = lo + threadIdx.x; i < hi; i += 256) { const unsigned int p = __ldg(src + i); s1 += p; s2 += p * p; }
    }
    __shared__ unsigned long long sm1[256], sm2[256];
    sm1[threadIdx.x] = s1; sm2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sm1[threadIdx.x] += sm1[threadIdx.x + o]; sm2[threadIdx.x] += sm2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[((size_t)b * chunks + blockIdx.x) * 2 + 0] = (double)sm1[0];
        partial[((size_t)b * chunks + blockIdx.x) * 2 + 1] = (double)sm2[0];
    }
}
__global__ void std_u8_finish_kernel(const double* __restrict__ partial, int chunks, int B, double n, float* __restrict__ mean,
                                     float* __restrict__ denom) {
    const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    double s1 = 0.0, s2 = 0.0;
    for (int c = lane; c < chunks; c += 32) { s1 += partial[((size_t)b * chunks + c) * 2]; s2 += partial[((size_t)b * chunks + c) * 2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) {
        const double mu = s1 / n;
        double var = s2 / n - mu * mu;
        if (var < 0.0) var = 0.0;
        mean[b] = (float)mu;
        denom[b] = fmaxf((float)sqrt(var), (float)(1.0 / sqrt(n)));
    }
}
__global__ void __launch_bounds__(256)
std_u8_apply_kernel(const unsigned char* __restrict__ x, size_t n, const float* __restrict__ mean, const float* __restrict__ denom,
                    float* __restrict__ out) {
    const int b = blockIdx.y;
    const float mu = mean[b], d = denom[b];
    const unsigned char* src = x + (size_t)b * n;
    float* dst = out + (size_t)b * n;
    const size_t n16 = n / 16;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 o;
            o.x = __fdiv_rn((float)(w[k] & 0xffu) - mu, d);
            o.y = __fdiv_rn((float)((w[k] >> 8) & 0xffu) - mu, d);
            o.z = __fdiv_rn((float)((w[k] >> 16) & 0xffu) - mu, d);
            o.w = __fdiv_rn((float)(w[k] >> 24) - mu, d);
            reinterpret_cast<float4*>(dst)[i * 4 + k] = o;
        }
    }
}
static int standardize_run_u8(const unsigned char* x, int B, size_t n, float* out, double* partial, float* mean, float* denom, cudaStream_t s) {
    const int chunks = (int)((n + STD_CHUNK - 1) / STD_CHUNK);
    std_u8_sums_kernel<<<dim3(chunks, B), 256, 0, s>>>(x, n, partial);
    std_u8_finish_kernel<<<(B + 7) / 8, 256, 0, s>>>(partial, chunks, B, (double)n, mean, denom);
    int ab = (int)((n / 16 + 255) / 256);
    if (ab > 148 * 2) ab = 148 * 2;
    if (ab < 1) ab = 1;
    std_u8_apply_kernel<<<dim3(ab, B), 256, 0, s>>>(x, n, mean, denom, out);
    Y2_CUDA(cudaGetLastError());
    for (int i = 0; i < 3; ++i) note_launch();
    return 0;
}

template <typename T>
static int standardize_run(const T* x, int B, size_t n, float* out, double* partial, float* mean, float* denom, cudaStream_t s) {
    const int chunks = (int)((n + STD_CHUNK - 1) / STD_CHUNK);
    dim3 grid(chunks, B);
    std_partial_kernel<T, 0><<<grid, 256, 0, s>>>(x, n, nullptr, partial);
    std_finish_kernel<0><<<(B + 7) / 8, 256, 0, s>>>(partial, chunks, B, (double)n, mean, denom);
    std_partial_kernel<T, 1><<<grid, 256, 0, s>>>(x, n, mean, partial);
    std_finish_kernel<1><<<(B + 7) / 8, 256, 0, s>>>(partial, chunks, B, (double)n, mean, denom);
    int ab = (int)((n + 2047) / 2048);
    if (ab > 148 * 4) ab = 148 * 4;
    std_apply_kernel<T><<<dim3(ab, B), 256, 0, s>>>(x, n, mean, denom, out);
    Y2_CUDA(cudaGetLastError());
    for (int i = 0; i < 5; ++i) note_launch();
    return 0;
}
size_t standardize_workspace_bytes(int B, size_t n) {
    const size_t chunks = (n + STD_CHUNK - 1) / STD_CHUNK;
    return ((size_t)B * chunks * 2 * sizeof(double) + 255) / 256 * 256 + 2 * (((size_t)B * sizeof(float) + 255) / 256 * 256);
}
int standardize_launch(const void* x, int elem_bytes, int B, size_t n, float* out, void* ws, cudaStream_t s) {
    const size_t chunks = (n + STD_CHUNK - 1) / STD_CHUNK;
    char* p = static_cast<char*>(ws);
    double* partial = reinterpret_cast<double*>(p); p += ((size_t)B * chunks * 2 * sizeof(double) + 255) / 256 * 256;
    float* mean = reinterpret_cast<float*>(p); p += ((size_t)B * sizeof(float) + 255) / 256 * 256;
    float* denom = reinterpret_cast<float*>(p);
    // uint8 with 16-byte aligned images: exact integer sums in one pass + vectorised apply; otherwise the generic two-pass path
    if (elem_bytes == 1 && n % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        return standardize_run_u8(static_cast<const unsigned char*>(x), B, n, out, partial, mean, denom, s);
    if (elem_bytes == 1) return standardize_run(static_cast<const unsigned char*>(x), B, n, out, partial, mean, denom, s);
    return standardize_run(static_cast<const float*>(x), B, n, out, partial, mean, denom, s);
}

// ---------------------------------------------------------------------------------------------
// detections (detect.py:72-87), two passes, no host round trip:
//   argmax  : a warp per box over the whole batch (coalesced row read, shuffle argmax with "first maximum" tie-break);
//             class and score are parked UNCOMPACTED at index n of the output arrays.
//   compact : one block per image walks its boxes in chunks of blockDim (box-index order), block-wide exclusive scan of
//             the keep flags (`score > threshold`), ordered append.  A kept box moves from n to pos <= n and every chunk is
//             read into registers before anything of it is written, so the compaction is done in place.
__global__ void __launch_bounds__(256)
detections_argmax_kernel(const float* __restrict__ conf, long long boxes, int C, int* __restrict__ cls, float* __restrict__ score) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long g = warp0; g < boxes; g += nwarps) {
        const float* row = conf + (size_t)g * C;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = __ldg(row + c);
            if (v > best) { best = v; bi = c; }             // ascending c per lane: strict > keeps the first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { cls[g] = bi; score[g] = best; }
    }
}
__global__ void __launch_bounds__(1024)
detections_compact_kernel(const float* __restrict__ xy_min, const float* __restrict__ xy_max, int N, float threshold, float sx, float sy,
                          int* __restrict__ count, int* __restrict__ box, int* __restrict__ cls, float* __restrict__ score,
                          float* __restrict__ xywh) {
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __shared__ int wsum[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += blockDim.x) {
        const int n = n0 + threadIdx.x;
        float sc = 0.f;
        int ci = 0;
        bool k = false;
        if (n < N) {
            sc = score[(size_t)b * N + n];
            ci = cls[(size_t)b * N + n];
            k = sc > threshold;                              // strict, NaN never kept (detect.py:80)
        }
        const uint32_t m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();                                     // every thread has read its (cls, score) and published its warp's count
        int pos = base + __popc(m & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += wsum[w];
        if (k) {
            const size_t o = (size_t)b * N + pos;
            box[o] = n; cls[o] = ci; score[o] = sc;
            const float2 lo = __ldg(reinterpret_cast<const float2*>(xy_min) + (size_t)b * N + n);
            const float2 hi = __ldg(reinterpret_cast<const float2*>(xy_max) + (size_t)b * N + n);
            *reinterpret_cast<float4*>(xywh + o * 4) = make_float4(__fmul_rn(lo.x, sx), __fmul_rn(lo.y, sy), __fmul_rn(__fsub_rn(hi.x, lo.x), sx),
                                                                  __fmul_rn(__fsub_rn(hi.y, lo.y), sy));
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nw; ++w) t += wsum[w]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) count[b] = base;
}
int detections_launch(const float* conf, const float* xy_min, const float* xy_max, int B, int N, int C, float threshold, float sx, float sy,
                      int* count, int* box, int* cls, float* score, float* xywh, cudaStream_t s) {
    Y2_REQUIRE((reinterpret_cast<uintptr_t>(xy_min) & 7) == 0 && (reinterpret_cast<uintptr_t>(xy_max) & 7) == 0 &&
               (reinterpret_cast<uintptr_t>(xywh) & 15) == 0, "detections: box arrays must be 8-byte, xywh 16-byte aligned");
    const long long boxes = (long long)B * N;
    long long blocks = (boxes + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    detections_argmax_kernel<<<(int)blocks, 256, 0, s>>>(conf, boxes, C, cls, score);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    detections_compact_kernel<<<B, 1024, 0, s>>>(xy_min, xy_max, N, threshold, sx, sy, count, box, cls, score, xywh);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// transform_labels (utils/data/__init__.py:112-145), batched: one block per image.  Phase 1 zero-fills the image's slice of
// the six label tensors with 128-bit stores (this is the byte traffic: 4*cells*(C+10) B per image); phase 2 gives each
// object a thread.  numpy's fancy-index assignment lets the LAST object that lands in a cell win mask/coords/offsets
// (class bits accumulate over all of them): a thread writes those only if no later object of the image maps to its cell.
// All arithmetic is float32 in the reference's evaluation order (float32 arrays x Python ints), round-to-nearest, no FMA.
__device__ __forceinline__ int label_cell(const float* __restrict__ c, int cw, int ch, float* ox, float* oy) {
    const float x = __fdiv_rn(__fmul_rn((float)cw, __fadd_rn(c[0], c[2])), 2.0f);      // cell_width * (xmin + xmax) / 2
    const float y = __fdiv_rn(__fmul_rn((float)ch, __fadd_rn(c[1], c[3])), 2.0f);
    const float ix = floorf(x), iy = floorf(y);
    *ox = __fsub_rn(x, ix); *oy = __fsub_rn(y, iy);
    const float fi = __fadd_rn(__fmul_rn(iy, (float)cw), ix);                          // (iy * cell_width + ix).astype(int)
    if (!(fi > -2.0e9f && fi < 2.0e9f)) return INT_MIN;                                // NaN / overflow: out of range below
    return (int)fi;
}
__global__ void __launch_bounds__(256)
transform_labels_kernel(const int* __restrict__ ocls, const float* __restrict__ ocoord, const int* __restrict__ offsets, int classes,
                        int cw, int ch, float* __restrict__ mask, float* __restrict__ prob, float* __restrict__ coords,
                        float* __restrict__ oxy_min, float* __restrict__ oxy_max, float* __restrict__ areas, int* __restrict__ status) {
    const int b = blockIdx.x, cells = cw * ch;
    auto zero = [&](float* base, size_t n) {            // n floats at base (16-byte aligned when n % 4 == 0 per image; else scalar)
        if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
            float4* v = reinterpret_cast<float4*>(base);
            for (size_t i = threadIdx.x; i < n / 4; i += blockDim.x) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (size_t i = threadIdx.x; i < n; i += blockDim.x) base[i] = 0.f;
        }
    };
    zero(mask + (size_t)b * cells, cells);
    zero(prob + (size_t)b * cells * classes, (size_t)cells * classes);
    zero(coords + (size_t)b * cells * 4, (size_t)cells * 4);
    zero(oxy_min + (size_t)b * cells * 2, (size_t)cells * 2);
    zero(oxy_max + (size_t)b * cells * 2, (size_t)cells * 2);
    zero(areas + (size_t)b * cells, cells);
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    const int o0 = offsets[b], n = offsets[b + 1] - o0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const float* c = ocoord + (size_t)(o0 + t) * 4;
        float ox, oy;
        int idx = label_cell(c, cw, ch, &ox, &oy);
        const int k = ocls[o0 + t];
        if (idx < 0 && idx >= -cells) idx += cells;                  // numpy wraps negative indices
        if (idx < 0 || idx >= cells) { atomicOr(&bad, 1); continue; }        // IndexError in the reference
        if (k < -classes || k >= classes) { atomicOr(&bad, 1); continue; }
        prob[((size_t)b * cells + idx) * classes + (k < 0 ? k + classes : k)] = 1.0f;   // every object sets its class bit
        bool last = true;
        for (int j = t + 1; j < n && last; ++j) {
            float tx, ty;
            int ij = label_cell(ocoord + (size_t)(o0 + j) * 4, cw, ch, &tx, &ty);
            if (ij < 0 && ij >= -cells) ij += cells;
            if (ij == idx) last = false;
        }
        const float w = __fsub_rn(c[2], c[0]), h = __fsub_rn(c[3], c[1]);
        const float hw = __fmul_rn(__fdiv_rn(w, 2.0f), (float)cw), hh = __fmul_rn(__fdiv_rn(h, 2.0f), (float)ch);   // w / 2 * cell_width
        const float x0 = __fsub_rn(ox, hw), y0 = __fsub_rn(oy, hh), x1 = __fadd_rn(ox, hw), y1 = __fadd_rn(oy, hh);
        const float dw = __fsub_rn(x1, x0), dh = __fsub_rn(y1, y0);
        if (!(dw >= 0.f) || !(dh >= 0.f)) atomicOr(&bad, 2);         // `assert np.all(wh >= 0)` (:142)
        if (!last) continue;
        const size_t cell = (size_t)b * cells + idx;
        mask[cell] = 1.0f;
        coords[cell * 4 + 0] = ox; coords[cell * 4 + 1] = oy;
        coords[cell * 4 + 2] = __fsqrt_rn(w); coords[cell * 4 + 3] = __fsqrt_rn(h);
        oxy_min[cell * 2 + 0] = x0; oxy_min[cell * 2 + 1] = y0;
        oxy_max[cell * 2 + 0] = x1; oxy_max[cell * 2 + 1] = y1;
        areas[cell] = __fmul_rn(dw, dh);                             // np.multiply.reduce(offset_xy_max - offset_xy_min, -1)
    }
    __syncthreads();
    if (threadIdx.x == 0 && status) status[b] = bad;
}
int transform_labels_launch(const int* ocls, const float* ocoord, const int* offsets, int B, int classes, int cw, int ch, float* mask,
                            float* prob, float* coords, float* oxy_min, float* oxy_max, float* areas, int* status, cudaStream_t s) {
    transform_labels_kernel<<<B, 256, 0, s>>>(ocls, ocoord, offsets, classes, cw, ch, mask, prob, coords, oxy_min, oxy_max, areas, status);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
