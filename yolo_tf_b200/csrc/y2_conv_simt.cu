// conv0 of Darknet-19 (model/yolo2/inference.py:73, first loop iteration): 3x3, Cin=3 -> Cout=32,
// SAME, + BN (scale/bias) + leaky + 2x2/2 max-pool (:74), fused.  K = 27 is far too small for a
// tensor-core K-block, and the layer is 0.9 % of the FLOPs, so it runs on the CUDA cores in exact
// fp32 (fmaf accumulate).  One thread = one pooled pixel (a 4x4x3 input patch in registers,
// 2x2 conv outputs x 32 channels in two passes of 16); weights are broadcast from shared memory
// as 128-bit loads.  Output goes straight to the bf16 hi/lo planes conv1's TMA reads.
#include "y2_internal.h"

namespace y2 {

__global__ void __launch_bounds__(128)
conv0_pool_kernel(const float* __restrict__ x, const float* __restrict__ w_hwio, const float* __restrict__ scale,
                  const float* __restrict__ bias, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int B, int H,
                  int W) {
    __shared__ __align__(16) float sw[27 * 32];
    __shared__ float ssc[32], sbi[32];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = w_hwio[i];   // HWIO: [(ky*3+kx)*3+c][n]
    if (threadIdx.x < 32) {
        ssc[threadIdx.x] = scale[threadIdx.x];
        sbi[threadIdx.x] = bias[threadIdx.x];
    }
    __syncthreads();
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)B * Ho * Wo;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int xo = (int)(idx % Wo);
        size_t t = idx / Wo;
        const int yo = (int)(t % Ho);
        const int b = (int)(t / Ho);
        float patch[4][4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int yy = 2 * yo - 1 + r;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int xx = 2 * xo - 1 + c;
                const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
                const float* src = x + (((size_t)b * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * 3;
                patch[r][c][0] = ok ? __ldg(src) : 0.f;
                patch[r][c][1] = ok ? __ldg(src + 1) : 0.f;
                patch[r][c][2] = ok ? __ldg(src + 2) : 0.f;
            }
        }
        const size_t obase = (((size_t)b * Ho + yo) * Wo + xo) * 32;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            // packed fp32 FMA (FFMA2, sm_100): two output channels per instruction
            float2 acc2[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int n = 0; n < 8; ++n) acc2[p][n] = make_float2(0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4* wr = reinterpret_cast<const float4*>(&sw[((ky * 3 + kx) * 3 + c) * 32 + half * 16]);
                        float2 wv[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 q = wr[j];
                            wv[2 * j] = make_float2(q.x, q.y);
                            wv[2 * j + 1] = make_float2(q.z, q.w);
                        }
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float v = patch[(p >> 1) + ky][(p & 1) + kx][c];
                            const float2 vv = make_float2(v, v);
#pragma unroll
                            for (int n = 0; n < 8; ++n) acc2[p][n] = __ffma2_rn(vv, wv[n], acc2[p][n]);
                        }
                    }
            float acc[4][16];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    acc[p][2 * n] = acc2[p][n].x;
                    acc[p][2 * n + 1] = acc2[p][n].y;
                }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int n = 0; n < 16; n += 2) {
                float m[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float sc = ssc[half * 16 + n + e], bi = sbi[half * 16 + n + e];
                    float best = -INFINITY;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float tv = acc[p][n + e] * sc + bi;
                        best = fmaxf(best, fmaxf(tv, 0.1f * tv));
                    }
                    m[e] = best;
                }
                const bf16 h0 = __float2bfloat16_rn(m[0]), h1 = __float2bfloat16_rn(m[1]);
                const bf16 l0 = __float2bfloat16_rn(m[0] - __bfloat162float(h0));
                const bf16 l1 = __float2bfloat16_rn(m[1] - __bfloat162float(h1));
                hi[n / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lo[n / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            uint4* dh = reinterpret_cast<uint4*>(out_hi + obase + half * 16);
            uint4* dl = reinterpret_cast<uint4*>(out_lo + obase + half * 16);
            dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Training mode: conv0 WITHOUT BN/pool (batch statistics need the raw output first): z[B,H,W,32] fp32.
__global__ void __launch_bounds__(128)
conv0_raw_kernel(const float* __restrict__ x, const float* __restrict__ w_hwio, float* __restrict__ z, int B, int H, int W) {
    // One thread = two horizontally adjacent output pixels x all 32 channels: every 128-bit weight read from shared memory
    // (warp-broadcast) feeds 8 FMAs instead of 4, which moves the kernel from LDS-bound to FFMA2-bound.
    __shared__ __align__(16) float sw[27 * 32];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = w_hwio[i];
    __syncthreads();
    const int W2 = W / 2;
    const size_t total = (size_t)B * H * W2;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int xx0 = (int)(idx % W2) * 2;
        size_t t = idx / W2;
        const int yy0 = (int)(t % H);
        const int b = (int)(t / H);
        float patch[3][4][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int yy = yy0 - 1 + r, xx = xx0 - 1 + c;
                const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
                const float* src = x + (((size_t)b * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) patch[r][c][ch] = ok ? __ldg(src + ch) : 0.f;
            }
        float2 acc0[16], acc1[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) { acc0[n] = make_float2(0.f, 0.f); acc1[n] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float2 v0 = make_float2(patch[r][c][ch], patch[r][c][ch]);
                    const float2 v1 = make_float2(patch[r][c + 1][ch], patch[r][c + 1][ch]);
                    const float4* wr = reinterpret_cast<const float4*>(&sw[((r * 3 + c) * 3 + ch) * 32]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q = wr[j];
                        const float2 wa = make_float2(q.x, q.y), wb = make_float2(q.z, q.w);
                        acc0[2 * j] = __ffma2_rn(v0, wa, acc0[2 * j]);
                        acc0[2 * j + 1] = __ffma2_rn(v0, wb, acc0[2 * j + 1]);
                        acc1[2 * j] = __ffma2_rn(v1, wa, acc1[2 * j]);
                        acc1[2 * j + 1] = __ffma2_rn(v1, wb, acc1[2 * j + 1]);
                    }
                }
        const size_t pix = ((size_t)b * H + yy0) * W + xx0;
        float4* dst = reinterpret_cast<float4*>(z + pix * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc0[2 * j].x, acc0[2 * j].y, acc0[2 * j + 1].x, acc0[2 * j + 1].y);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[8 + j] = make_float4(acc1[2 * j].x, acc1[2 * j].y, acc1[2 * j + 1].x, acc1[2 * j + 1].y);
    }
}
int conv0_raw_launch(const float* x, const float* w_hwio, float* z, int B, int H, int W, cudaStream_t s) {
    Y2_REQUIRE(W % 2 == 0, "conv0 raw: W must be even");
    const size_t total = (size_t)B * H * (W / 2);          // pixel pairs
    size_t blocks = (total + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    conv0_raw_kernel<<<(int)blocks, 128, 0, s>>>(x, w_hwio, z, B, H, W);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// conv0 weight gradient: dW[tap][cin][cout] = sum over pixels of x[p + tap][cin] * dx[p][cout]  (27 x 32 outputs, 11 M
// pixels at batch 64: a skinny reduction, CUDA cores).  lane = output channel, every lane of a warp walks the SAME
// pixels: the 27 input values of a pixel are warp-broadcast shared-memory reads (128-bit, 4 pixels = 18 consecutive floats
// per image row), dx is one coalesced 64-byte row per plane.  A block stages the three input rows of one image row
// (halo included), its 8 warps split the row's 4-pixel groups.  fp32 partial sums per lane, fp64 per-block partials,
// fixed-order finish (deterministic).
static constexpr int C0W_BLOCKS = 148 * 4;
static constexpr int C0W_MAXW = 1024;
__global__ void __launch_bounds__(256)
conv0_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dx_hi, const bf16* __restrict__ dx_lo,
                   double* __restrict__ partial, int B, int H, int W) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ __align__(16) float c0w_smem[];
    const int pitch = ((W + 2) * 3 + 2 + 3) & ~3;           // floats per staged row (+2: the last 128-bit read overshoots)
    float* sin_ = c0w_smem;                                  // [3][pitch]
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    const int groups = W / 4;
    const long long rows = (long long)B * H;
    for (long long ridx = blockIdx.x; ridx < rows; ridx += gridDim.x) {
        const int yy0 = (int)(ridx % H);
        const int b = (int)(ridx / H);
        __syncthreads();                                     // previous row's readers are done
        for (int e = threadIdx.x; e < 3 * pitch; e += blockDim.x) {
            const int r = e / pitch, rem = e - r * pitch;
            const int px = rem / 3, ch = rem - px * 3;
            const int yy = yy0 - 1 + r, xx = px - 1;
            const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
            sin_[e] = ok ? __ldg(x + (((size_t)b * H + yy) * W + xx) * 3 + ch) : 0.f;
        }
        __syncthreads();
        const size_t row_base = ((size_t)b * H + yy0) * W;
        for (int g = warp; g < groups; g += 8) {
            // dx of the 4 pixels, this lane's channel: hi + lo
            float d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const size_t idx = (row_base + 4 * g + i) * 32 + lane;
                const unsigned short h = __ldg(reinterpret_cast<const unsigned short*>(dx_hi) + idx);
                const unsigned short l = __ldg(reinterpret_cast<const unsigned short*>(dx_lo) + idx);
                d[i] = __uint_as_float((uint32_t)h << 16) + __uint_as_float((uint32_t)l << 16);
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float v[20];                                 // pixels 4g-1 .. 4g+4 of input row r: 18 floats (+2 unused)
                const float4* src = reinterpret_cast<const float4*>(sin_ + r * pitch + 12 * g);
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const float4 t = src[q];
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            acc[(r * 3 + c) * 3 + ch] = fmaf(v[(i + c) * 3 + ch], d[i], acc[(r * 3 + c) * 3 + ch]);
            }
        }
    }
    // block reduction over the 8 warps (same channel = same lane), then one fp64 partial per block
    __syncthreads();
    float* red = c0w_smem;                                   // [8][27][32] floats = 27.6 KB (reuses the staging area)
#pragma unroll
    for (int k = 0; k < 27; ++k) red[(warp * 27 + k) * 32 + lane] = acc[k];
    __syncthreads();
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += (double)red[w * 27 * 32 + i];
        partial[(size_t)blockIdx.x * 27 * 32 + i] = sacc;
    }
}
__global__ void conv0_wgrad_finish_kernel(const double* partial, int nblocks, float* dw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * 32) return;
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 27 * 32 + i];
    dw[i] = (float)s;
}
int conv0_wgrad_launch(const float* x, const bf16* dx_hi, const bf16* dx_lo, float* dw, double* partial, int B, int H, int W,
                       cudaStream_t s) {
    Y2_REQUIRE(W % 4 == 0 && W <= C0W_MAXW, "conv0 wgrad: W must be a multiple of 4 and <= %d (got %d)", C0W_MAXW, W);
    const int pitch = ((W + 2) * 3 + 2 + 3) & ~3;
    size_t smem = (size_t)3 * pitch * sizeof(float);
    if (smem < (size_t)8 * 27 * 32 * sizeof(float)) smem = (size_t)8 * 27 * 32 * sizeof(float);
    conv0_wgrad_kernel<<<C0W_BLOCKS, 256, smem, s>>>(x, dx_hi, dx_lo, partial, B, H, W);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    conv0_wgrad_finish_kernel<<<(27 * 32 + 127) / 128, 128, 0, s>>>(partial, C0W_BLOCKS, dw);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int conv0_pool_launch(const float* x, const float* w_hwio, const float* scale, const float* bias, bf16* out_hi,
                      bf16* out_lo, int B, int H, int W, cudaStream_t s) {
    Y2_REQUIRE(H % 2 == 0 && W % 2 == 0, "conv0: H and W must be even");
    const size_t total = (size_t)B * (H / 2) * (W / 2);
    size_t blocks = (total + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    conv0_pool_kernel<<<(int)blocks, 128, 0, s>>>(x, w_hwio, scale, bias, out_hi, out_lo, B, H, W);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
