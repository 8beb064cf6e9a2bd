// HBM-bound layout kernels around the conv stack: split/merge of bf16 hi/lo planes, weight
// packing, BN folding, 2x2 max-pool on planes, reorg (space-to-depth).  All use 128-bit
// accesses on the contiguous NHWC channel runs and grid-stride loops sized to the SM count.
#include "y2_internal.h"

namespace y2 {

static inline int grid_for(size_t work_items, int threads) {
    size_t blocks = (work_items + threads - 1) / threads;
    const size_t cap = 148 * 8;   // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__device__ __forceinline__ float bf16_lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ---------------------------------------------------------------- split fp32 -> planes
__global__ void split_planes_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                    size_t n4, int f16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        unsigned short h[4], l[4];
        plane_split(v.x, f16, h[0], l[0]); plane_split(v.y, f16, h[1], l[1]);
        plane_split(v.z, f16, h[2], l[2]); plane_split(v.w, f16, h[3], l[3]);
        hi[i] = make_uint2(plane_pack2(h[0], h[1]), plane_pack2(h[2], h[3]));
        lo[i] = make_uint2(plane_pack2(l[0], l[1]), plane_pack2(l[2], l[3]));
    }
}
int split_planes_launch(const float* src, bf16* dst_hi, bf16* dst_lo, size_t n, cudaStream_t s, int f16) {
    Y2_REQUIRE(n % 4 == 0, "split_planes: element count must be a multiple of 4");
    split_planes_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(src),
                                                           reinterpret_cast<uint2*>(dst_hi),
                                                           reinterpret_cast<uint2*>(dst_lo), n / 4, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- power-of-two scale of a weight tensor (fp16 planes)
// amax over the tensor (positive floats order like their bit patterns -> atomicMax on the bits), then k with
// amax * 2^k in [2^13, 2^14): the hi plane keeps 11 significand bits for every weight above amax * 2^-27 and the lo plane
// (the fp16 of the residual) 11 more for every weight above amax * 2^-16; nothing can overflow.
__global__ void amax_bits_kernel(const float* __restrict__ src, size_t n, unsigned int* __restrict__ out_bits) {
    float m = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = fabsf(__ldg(src + i));
        if (v < INFINITY) m = fmaxf(m, v);                   // NaN / inf weights do not steer the scale
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}
__global__ void pow2_from_amax_kernel(float* out2) {
    const float amax = __uint_as_float(reinterpret_cast<unsigned int*>(out2)[0]);
    int k = 0;
    if (amax > 0.f) {
        int e;
        frexpf(amax, &e);                                     // amax = f * 2^e, f in [0.5, 1)
        k = 14 - e;                                           // amax * 2^k in [2^13, 2^14)
        k = max(-100, min(100, k));
    }
    out2[0] = exp2f((float)k);
    out2[1] = exp2f((float)-k);
}
int pow2_scale_launch(const float* src, size_t n, float* out2, cudaStream_t s) {
    Y2_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(float), s));
    amax_bits_kernel<<<grid_for(n, 256), 256, 0, s>>>(src, n, reinterpret_cast<unsigned int*>(out2));
    Y2_CUDA(cudaGetLastError());
    note_launch();
    pow2_from_amax_kernel<<<1, 1, 0, s>>>(out2);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}
__global__ void fold_scale_kernel(const float* scale, const float* factor, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fmul_rn(scale ? scale[i] : 1.0f, factor[0]);          // a power of two: exact
}
int fold_scale_launch(const float* scale, const float* factor, float* out, int n, cudaStream_t s) {
    fold_scale_kernel<<<(n + 127) / 128, 128, 0, s>>>(scale, factor, out, n);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- merge planes -> fp32
__global__ void merge_planes_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, float* __restrict__ dst,
                                    size_t rows, int cols, long long ld, int f16) {
    const int c4 = cols / 4;
    const size_t total = rows * (size_t)c4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / c4;
        const int c = (int)(i - r * c4) * 4;
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + r * ld + c));
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(lo + r * ld + c));
        float4 o;
        o.x = plane_dec((unsigned short)h.x, f16) + plane_dec((unsigned short)l.x, f16);
        o.y = plane_dec((unsigned short)(h.x >> 16), f16) + plane_dec((unsigned short)(l.x >> 16), f16);
        o.z = plane_dec((unsigned short)h.y, f16) + plane_dec((unsigned short)l.y, f16);
        o.w = plane_dec((unsigned short)(h.y >> 16), f16) + plane_dec((unsigned short)(l.y >> 16), f16);
        *reinterpret_cast<float4*>(dst + r * (size_t)cols + c) = o;
    }
}
int merge_planes_launch(const bf16* hi, const bf16* lo, float* dst, size_t rows, int cols, long long ld,
                        cudaStream_t s, int f16) {
    Y2_REQUIRE(cols % 4 == 0 && ld % 4 == 0, "merge_planes: cols and pitch must be multiples of 4");
    merge_planes_kernel<<<grid_for(rows * (size_t)(cols / 4), 256), 256, 0, s>>>(hi, lo, dst, rows, cols, ld, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- weights HWIO fp32 -> [2][cout_pad][tap*Cin] bf16
__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int taps, int cin, int cout,
                                    int cout_pad, int f16, const float* __restrict__ wscale) {
    const float ws = wscale ? __ldg(wscale) : 1.0f;
    unsigned short* o16 = reinterpret_cast<unsigned short*>(out);
    const size_t K = (size_t)taps * cin;
    const size_t total = (size_t)cout_pad * K;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / K;
        const size_t k = i - n * K;            // k = tap*cin + c ; HWIO index = (tap*cin + c)*cout + n
        float v = 0.f;
        if (n < (size_t)cout) v = __fmul_rn(__ldg(w + k * cout + n), ws);
        plane_split(v, f16, o16[i], o16[total + i]);
    }
}
// HWIO [taps][cin][cout] -> HWIO [taps][cin_s][cout_s], zero-filled where the stored channel count is padded
// (tiny: conv0's 16 outputs / conv1's 16 inputs are stored as 32 so the same kernels serve both networks).
__global__ void pad_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int taps, int cin, int cout, int cin_s,
                                   int cout_s) {
    const size_t total = (size_t)taps * cin_s * cout_s;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % cout_s);
        const size_t r = i / cout_s;
        const int c = (int)(r % cin_s);
        const size_t tap = r / cin_s;
        out[i] = (n < cout && c < cin) ? __ldg(w + (tap * cin + c) * cout + n) : 0.f;
    }
}
int pad_weights_launch(const float* w_hwio, float* out, int ksize, int cin, int cout, int cin_s, int cout_s, cudaStream_t s) {
    const size_t total = (size_t)ksize * ksize * cin_s * cout_s;
    pad_weights_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, out, ksize * ksize, cin, cout, cin_s, cout_s);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}
int pack_weights_launch(const float* w_hwio, bf16* wpack, int ksize, int cin, int cout, int cout_pad, cudaStream_t s, int f16,
                        const float* wscale) {
    const size_t total = (size_t)cout_pad * ksize * ksize * cin;
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, s>>>(w_hwio, wpack, ksize * ksize, cin, cout, cout_pad, f16, wscale);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- BN inference fold (TF arithmetic)
// inv = rsqrt(var + eps) * gamma ; scale = inv ; bias = beta - mean * inv   (tf.nn.batch_normalization)
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               float* scale, float* bias, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inv = __fmul_rn(rsqrtf(__fadd_rn(var[i], eps)), gamma ? gamma[i] : 1.0f);
    scale[i] = inv;
    bias[i] = __fsub_rn(beta ? beta[i] : 0.0f, __fmul_rn(mean[i], inv));
}
int bn_fold_launch(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                   float* bias, int n, cudaStream_t s) {
    bn_fold_kernel<<<(n + 127) / 128, 128, 0, s>>>(gamma, beta, mean, var, eps, scale, bias, n);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- 2x2 max-pool on planes (stride 2, or stride 1 SAME)
// One thread = 8 channels of one pooled pixel (16-byte vectors of each plane).  The winner is
// chosen on the exact value hi+lo (16 significand bits, exact in fp32) and its (hi,lo) pair copied.
// stride 1 (tiny's last pool, model/yolo2/inference.py:42): TF SAME pads one row/column at the bottom/right only and
// max-pool ignores padding, so the window of the last row/column is clipped (the clipped taps re-read a valid one).
__global__ void maxpool_planes_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo,
                                      bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int B, int H, int W,
                                      int C, int stride, int f16) {
    const int Ho = H / stride, Wo = W / stride, c8 = C / 8;
    const size_t total = (size_t)B * Ho * Wo * c8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % c8);
        size_t t = i / c8;
        const int xo = (int)(t % Wo);
        t /= Wo;
        const int yo = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const size_t base = (((size_t)b * H + stride * yo) * W + stride * xo) * C + (size_t)cv * 8;
        const size_t ox = (stride * xo + 1 < W) ? (size_t)C : 0, oy = (stride * yo + 1 < H) ? (size_t)W * C : 0;
        uint4 h[4], l[4];
        h[0] = __ldg(reinterpret_cast<const uint4*>(in_hi + base));
        h[1] = __ldg(reinterpret_cast<const uint4*>(in_hi + base + ox));
        h[2] = __ldg(reinterpret_cast<const uint4*>(in_hi + base + oy));
        h[3] = __ldg(reinterpret_cast<const uint4*>(in_hi + base + oy + ox));
        l[0] = __ldg(reinterpret_cast<const uint4*>(in_lo + base));
        l[1] = __ldg(reinterpret_cast<const uint4*>(in_lo + base + ox));
        l[2] = __ldg(reinterpret_cast<const uint4*>(in_lo + base + oy));
        l[3] = __ldg(reinterpret_cast<const uint4*>(in_lo + base + oy + ox));
        uint32_t oh[4], ol[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t* hw0 = reinterpret_cast<const uint32_t*>(&h[0]);
            const uint32_t* lw0 = reinterpret_cast<const uint32_t*>(&l[0]);
            // low half / high half of word w, over the 4 window positions
            float best_a = -INFINITY, best_b = -INFINITY;
            uint32_t ha = 0, la = 0, hb = 0, lb = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t hw = hw0[q * 4 + w], lw = lw0[q * 4 + w];
                const float va = plane_dec((unsigned short)hw, f16) + plane_dec((unsigned short)lw, f16);
                const float vb = plane_dec((unsigned short)(hw >> 16), f16) + plane_dec((unsigned short)(lw >> 16), f16);
                if (va > best_a || q == 0) { best_a = va; ha = hw & 0xFFFFu; la = lw & 0xFFFFu; }
                if (vb > best_b || q == 0) { best_b = vb; hb = hw & 0xFFFF0000u; lb = lw & 0xFFFF0000u; }
            }
            oh[w] = ha | hb;
            ol[w] = la | lb;
        }
        const size_t ob = (((size_t)b * Ho + yo) * Wo + xo) * C + (size_t)cv * 8;
        *reinterpret_cast<uint4*>(out_hi + ob) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<uint4*>(out_lo + ob) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
}
int maxpool_planes_launch(const bf16* in_hi, const bf16* in_lo, bf16* out_hi, bf16* out_lo, int B, int H, int W,
                          int C, cudaStream_t s, int stride, int f16) {
    Y2_REQUIRE(stride == 1 || stride == 2, "maxpool: stride must be 1 or 2");
    Y2_REQUIRE(C % 8 == 0 && H % stride == 0 && W % stride == 0, "maxpool: C%%8, H%%stride, W%%stride must be 0");
    const size_t total = (size_t)B * (H / stride) * (W / stride) * (C / 8);
    maxpool_planes_kernel<<<grid_for(total, 256), 256, 0, s>>>(in_hi, in_lo, out_hi, out_lo, B, H, W, C, stride, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- reorg (space-to-depth)
// model/yolo2/function.py:22-29: out[b, y, x, (dy*s+dx)*C + c] = in[b, s*y+dy, s*x+dx, c].
// Pure permutation of contiguous C-runs, moved as 16-byte vectors.
template <typename V>
__global__ void reorg_kernel(const V* __restrict__ in, V* __restrict__ out, int B, int H, int W, int cvec,
                             int stride, long long out_ld_vec) {
    const int Ho = H / stride, Wo = W / stride;
    const int ss = stride * stride;
    const size_t total = (size_t)B * Ho * Wo * ss * cvec;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % cvec);
        size_t t = i / cvec;
        const int d = (int)(t % ss);
        t /= ss;
        const int xo = (int)(t % Wo);
        t /= Wo;
        const int yo = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int dy = d / stride, dx = d - dy * stride;
        const size_t src = (((size_t)b * H + (size_t)yo * stride + dy) * W + (size_t)xo * stride + dx) * cvec + cv;
        const size_t dst = (((size_t)b * Ho + yo) * Wo + xo) * (size_t)out_ld_vec + (size_t)d * cvec + cv;
        out[dst] = in[src];
    }
}
template <typename V>
static int reorg_dispatch(const void* in, void* out, int B, int H, int W, size_t row_bytes, int stride,
                          size_t out_ld_bytes, cudaStream_t s) {
    const int cvec = (int)(row_bytes / sizeof(V));
    const size_t total = (size_t)B * H * W * cvec;
    reorg_kernel<V><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const V*>(in), reinterpret_cast<V*>(out), B,
                                                        H, W, cvec, stride, (long long)(out_ld_bytes / sizeof(V)));
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}
// Widest vector (16/8/4/2 bytes) that divides the channel run, the output pitch and both base addresses.
// leaky_relu(inputs, alpha) = max(x, alpha * x), float32 -- model/yolo/function.py:21-24 as a standalone op (inside the network it is
// fused into every conv epilogue).  HBM-bound: 8 bytes per element, 128-bit accesses where the pointers allow, grid-stride.
__global__ void __launch_bounds__(256) leaky_relu_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float alpha, int vec4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec4) {
        const size_t n4 = n / 4;
        for (size_t i = t; i < n4; i += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
            reinterpret_cast<float4*>(out)[i] = make_float4(fmaxf(v.x, __fmul_rn(alpha, v.x)), fmaxf(v.y, __fmul_rn(alpha, v.y)),
                                                             fmaxf(v.z, __fmul_rn(alpha, v.z)), fmaxf(v.w, __fmul_rn(alpha, v.w)));
        }
        for (size_t i = n4 * 4 + t; i < n; i += stride) out[i] = fmaxf(in[i], __fmul_rn(alpha, in[i]));
    } else {
        for (size_t i = t; i < n; i += stride) out[i] = fmaxf(in[i], __fmul_rn(alpha, in[i]));
    }
}
int leaky_relu_launch(const float* in, float* out, size_t n, float alpha, cudaStream_t s) {
    Y2_REQUIRE(in && out, "leaky_relu: null argument");
    if (n == 0) return 0;
    const int vec4 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    size_t blocks = (n / (vec4 ? 4 : 1) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks == 0) blocks = 1;
    leaky_relu_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, n, alpha, vec4);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int reorg_launch(const void* in, void* out, int B, int H, int W, int C, int stride, int elem_bytes, long long out_ld,
                 cudaStream_t s) {
    Y2_REQUIRE(stride >= 1 && H % stride == 0 && W % stride == 0, "reorg: H, W must be divisible by stride");
    Y2_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "reorg: elem_bytes must be 2 or 4");
    if ((size_t)B * H * W * C == 0) return 0;
    const size_t row_bytes = (size_t)C * elem_bytes, ld_bytes = (size_t)out_ld * elem_bytes;
    const uintptr_t mix = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | row_bytes | ld_bytes;
    if ((mix & 15) == 0) return reorg_dispatch<uint4>(in, out, B, H, W, row_bytes, stride, ld_bytes, s);
    if ((mix & 7) == 0) return reorg_dispatch<uint2>(in, out, B, H, W, row_bytes, stride, ld_bytes, s);
    if ((mix & 3) == 0) return reorg_dispatch<uint32_t>(in, out, B, H, W, row_bytes, stride, ld_bytes, s);
    return reorg_dispatch<uint16_t>(in, out, B, H, W, row_bytes, stride, ld_bytes, s);
}

}  // namespace y2
