// Training-mode kernels around the tcgen05 GEMMs (HBM-bound passes):
//   batch-norm with BATCH statistics, forward and backward  <- slim.batch_norm(is_training=True,
//     decay=0.999, epsilon=1e-5), model/yolo2/inference.py:62-66, and its tf.gradients
//   leaky-ReLU backward (slope 1 for x >= 0, 0.1 below)       <- model/yolo/function.py:21-24
//   2x2 max-pool backward (first maximum wins)                <- slim.layers.max_pool2d, inference.py:69
//   reorg backward (depth-to-space)                           <- model/yolo2/function.py:22-29
//   bias gradient, fp32 -> split-plane conversion with padded pitch, dgrad weight packing
// Reductions over pixels are two-stage and deterministic: per-block fp64 partials, fixed-order finish.
#include "y2_internal.h"

namespace y2 {

static constexpr int RED_THREADS = 256;

static inline int red_blocks(size_t rows, int cols) {
    // each block strides over rows; enough blocks to fill the GPU, few enough to keep the finish cheap
    size_t work = rows * (size_t)((cols + 3) / 4);
    size_t b = (work + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
    if (b > 148 * 2) b = 148 * 2;      // one resident wave of the widest variant (2 blocks / SM); halves the finish traffic
    if (b < 1) b = 1;
    return (int)b;
}
size_t bn_partial_bytes() { return (size_t)148 * 4 * 2 * 3072 * sizeof(double); }   // [blocks][2][C<=3072]

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__device__ __forceinline__ void split4(const float f[4], uint2* hi, uint2* lo) {
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
    *hi = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    *lo = make_uint2(pack2(f[0] - h[0], f[1] - h[1]), pack2(f[2] - h[2], f[3] - h[3]));
}
// the same in either plane format (f16 = 1: fp16 hi / lo, see y2_internal.h)
__device__ __forceinline__ void split4f(const float f[4], uint2* hi, uint2* lo, int f16) {
    unsigned short h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) plane_split(f[j], f16, h[j], l[j]);
    *hi = make_uint2(plane_pack2(h[0], h[1]), plane_pack2(h[2], h[3]));
    *lo = make_uint2(plane_pack2(l[0], l[1]), plane_pack2(l[2], l[3]));
}
__device__ __forceinline__ void merge4f(uint2 h, uint2 l, float v[4], int f16) {
    v[0] = plane_dec((unsigned short)h.x, f16) + plane_dec((unsigned short)l.x, f16);
    v[1] = plane_dec((unsigned short)(h.x >> 16), f16) + plane_dec((unsigned short)(l.x >> 16), f16);
    v[2] = plane_dec((unsigned short)h.y, f16) + plane_dec((unsigned short)l.y, f16);
    v[3] = plane_dec((unsigned short)(h.y >> 16), f16) + plane_dec((unsigned short)(l.y >> 16), f16);
}

// ---------------------------------------------------------------------------------------------
// generic two-quantity column reduction over a [rows][C] fp32 matrix (4 channels per thread):
//   MODE 0 (bn stats):   q1 = z,  q2 = z*z
//   MODE 1 (bn bwd):     q1 = dz, q2 = dz * zhat   with dz = g * leaky'(z*scale+bias), zhat = (z-mean)*inv
//   MODE 2 (bias grad):  q1 = g,  q2 unused
struct RedArgs {
    const float* z;        // [rows][C] (pitch C)
    const float* g;        // [rows][*] pitch ldg (MODE 1, 2)
    long long ldg;
    const float *scale, *bias, *mean, *inv;
    size_t rows;
    int C;
    double* partial;       // [gridDim.x][2][C]
};

template <int MODE>
__global__ void __launch_bounds__(RED_THREADS) col_reduce_kernel(RedArgs a) {
    const int c4 = (a.C + 3) / 4;                 // channel groups of 4
    // thread -> (row lane, channel group): consecutive threads take consecutive channel groups (coalesced)
    const int groups_per_pass = RED_THREADS < c4 ? RED_THREADS : c4;
    const int rows_per_pass = RED_THREADS / groups_per_pass;
    for (int cg0 = 0; cg0 < c4; cg0 += groups_per_pass) {
        const int cg = cg0 + (int)(threadIdx.x % groups_per_pass);
        const int rlane = threadIdx.x / groups_per_pass;
        const bool active = cg < c4 && rlane < rows_per_pass;
        const int c = cg * 4;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
        float sc[4] = {1, 1, 1, 1}, bi[4] = {0, 0, 0, 0}, mu[4] = {0, 0, 0, 0}, iv[4] = {1, 1, 1, 1};
        if (active && MODE == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < a.C) { sc[j] = a.scale[c + j]; bi[j] = a.bias[c + j]; mu[j] = a.mean[c + j]; iv[j] = a.inv[c + j]; }
        }
        int flush = 0;
        if (active) {
            const bool zvec = c + 3 < a.C, gvec = zvec && (a.ldg & 3) == 0;
            auto load_z = [&](size_t r, float (&zv)[4]) {
                if (MODE == 2) return;
                if (zvec) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(a.z + r * a.C + c));
                    zv[0] = t.x; zv[1] = t.y; zv[2] = t.z; zv[3] = t.w;
                } else {
                    for (int j = 0; j < 4; ++j) if (c + j < a.C) zv[j] = __ldg(a.z + r * a.C + c + j);
                }
            };
            auto load_g = [&](size_t r, float (&gv)[4]) {
                if (MODE == 0) return;
                if (gvec) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(a.g + r * a.ldg + c));
                    gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
                } else {
                    for (int j = 0; j < 4; ++j) if (c + j < a.C) gv[j] = __ldg(a.g + r * a.ldg + c + j);
                }
            };
            auto accum = [&](const float (&zv)[4], const float (&gv)[4]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (MODE == 0) { s1[j] += zv[j]; s2[j] += zv[j] * zv[j]; }
                    if (MODE == 1) {
                        const float zb = zv[j] * sc[j] + bi[j];
                        const float dz = gv[j] * (zb >= 0.f ? 1.0f : 0.1f);
                        s1[j] += dz; s2[j] += dz * ((zv[j] - mu[j]) * iv[j]);
                    }
                    if (MODE == 2) s1[j] += gv[j];
                }
                if (++flush == 64) {                  // bound fp32 accumulation length, then widen
#pragma unroll
                    for (int j = 0; j < 4; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
                    flush = 0;
                }
            };
            const size_t rstep = (size_t)gridDim.x * rows_per_pass;
            size_t r = (size_t)blockIdx.x * rows_per_pass + rlane;
            for (; r + 3 * rstep < a.rows; r += 4 * rstep) {  // four rows of loads in flight; same summation order as one by one
                float zz[4][4], gg[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { zz[u][j] = 0.f; gg[u][j] = 0.f; }
                    load_z(r + u * rstep, zz[u]);
                    load_g(r + u * rstep, gg[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) accum(zz[u], gg[u]);
            }
            for (; r + rstep < a.rows; r += 2 * rstep) {
                float z0[4] = {0, 0, 0, 0}, g0[4] = {0, 0, 0, 0}, z1[4] = {0, 0, 0, 0}, g1[4] = {0, 0, 0, 0};
                load_z(r, z0); load_g(r, g0); load_z(r + rstep, z1); load_g(r + rstep, g1);
                accum(z0, g0);
                accum(z1, g1);
            }
            if (r < a.rows) {
                float z0[4] = {0, 0, 0, 0}, g0[4] = {0, 0, 0, 0};
                load_z(r, z0); load_g(r, g0);
                accum(z0, g0);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; }
        }
        // combine the row lanes of this block that share a channel group
        __shared__ double sm1[RED_THREADS][4], sm2[RED_THREADS][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { sm1[threadIdx.x][j] = active ? d1[j] : 0.0; sm2[threadIdx.x][j] = active ? d2[j] : 0.0; }
        __syncthreads();
        if (threadIdx.x < groups_per_pass && cg0 + (int)threadIdx.x < c4) {
            double t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
            for (int rl = 0; rl < rows_per_pass; ++rl)
#pragma unroll
                for (int j = 0; j < 4; ++j) { t1[j] += sm1[rl * groups_per_pass + threadIdx.x][j]; t2[j] += sm2[rl * groups_per_pass + threadIdx.x][j]; }
            const int cc = (cg0 + threadIdx.x) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (cc + j < a.C) {
                    a.partial[((size_t)blockIdx.x * 2 + 0) * a.C + cc + j] = t1[j];
                    a.partial[((size_t)blockIdx.x * 2 + 1) * a.C + cc + j] = t2[j];
                }
        }
        __syncthreads();
    }
}

// ---- finish kernels: a block owns 32 channels (lane = channel: coalesced 256-byte rows of the partial matrix); its
// 32 warps sum interleaved subsets of the per-block partials, then warp 0 combines them in fixed order (deterministic).
static constexpr int FIN_THREADS = 1024;
static constexpr int FIN_WARPS = FIN_THREADS / 32;
__device__ __forceinline__ bool finish_sums(const double* __restrict__ partial, int nblocks, int C, double* s1, double* s2) {
    __shared__ double sm[2][FIN_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double a1 = 0, a2 = 0;
    if (c < C) {
        // this warp's partial blocks: warp, warp + 32, ... ; four loads in flight, combined in a fixed order
        int b = warp;
        for (; b + 3 * FIN_WARPS < nblocks; b += 4 * FIN_WARPS) {
            double u1[4], u2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u1[k] = partial[((size_t)(b + k * FIN_WARPS) * 2) * C + c];
                u2[k] = partial[((size_t)(b + k * FIN_WARPS) * 2 + 1) * C + c];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { a1 += u1[k]; a2 += u2[k]; }
        }
        for (; b < nblocks; b += FIN_WARPS) { a1 += partial[((size_t)b * 2) * C + c]; a2 += partial[((size_t)b * 2 + 1) * C + c]; }
    }
    sm[0][warp][lane] = a1; sm[1][warp][lane] = a2;
    __syncthreads();
    if (warp != 0 || c >= C) return false;
    double t1 = 0, t2 = 0;
#pragma unroll
    for (int w = 0; w < FIN_WARPS; ++w) { t1 += sm[0][w][lane]; t2 += sm[1][w][lane]; }
    *s1 = t1; *s2 = t2;
    return true;
}
// forward: mean/var -> inv, scale, bias; moving averages updated as slim does
// (assign_moving_average: v -= (v - value) * (1 - decay))
__global__ void __launch_bounds__(FIN_THREADS)
bn_stats_finish_kernel(const double* partial, int nblocks, int C, double inv_rows, const float* gamma,
                       const float* beta, float eps, float decay, float* mean, float* inv, float* scale,
                       float* bias, float* moving_mean, float* moving_var) {
    double s1, s2;
    if (!finish_sums(partial, nblocks, C, &s1, &s2)) return;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const double m = s1 * inv_rows;
    double v = s2 * inv_rows - m * m;
    if (v < 0) v = 0;
    const float mf = (float)m, vf = (float)v;
    const float iv = rsqrtf(vf + eps);
    const float sc = iv * gamma[c];
    mean[c] = mf; inv[c] = iv; scale[c] = sc; bias[c] = beta[c] - mf * sc;
    moving_mean[c] -= (moving_mean[c] - mf) * (1.0f - decay);
    moving_var[c] -= (moving_var[c] - vf) * (1.0f - decay);
}
// backward: dbeta = s1, dgamma = s2 (written into the gradient bucket), and the per-channel means
__global__ void __launch_bounds__(FIN_THREADS)
bn_bwd_finish_kernel(const double* partial, int nblocks, int C, double inv_rows, float* dgamma, float* dbeta, float* m1, float* m2) {
    double s1, s2;
    if (!finish_sums(partial, nblocks, C, &s1, &s2)) return;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    dbeta[c] = (float)s1; dgamma[c] = (float)s2;
    m1[c] = (float)(s1 * inv_rows); m2[c] = (float)(s2 * inv_rows);
}
__global__ void __launch_bounds__(FIN_THREADS)
bias_grad_finish_kernel(const double* partial, int nblocks, int C, float* dbias) {
    double s1, s2;
    if (!finish_sums(partial, nblocks, C, &s1, &s2)) return;
    dbias[blockIdx.x * 32 + (threadIdx.x & 31)] = (float)s1;
}

int bn_stats_launch(const float* z, size_t rows, int C, const float* gamma, const float* beta, float eps, float decay,
                    float* mean, float* inv, float* scale, float* bias, float* moving_mean, float* moving_var,
                    double* partial, cudaStream_t s) {
    Y2_REQUIRE(C <= 3072, "bn_stats: too many channels");
    RedArgs a = {};
    a.z = z; a.rows = rows; a.C = C; a.partial = partial;
    const int nb = red_blocks(rows, C);
    col_reduce_kernel<0><<<nb, RED_THREADS, 0, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    bn_stats_finish_kernel<<<(C + 31) / 32, FIN_THREADS, 0, s>>>(partial, nb, C, 1.0 / (double)rows, gamma, beta, eps, decay, mean,
                                                          inv, scale, bias, moving_mean, moving_var);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int bn_bwd_reduce_launch(const float* z, const float* g, long long ldg, size_t rows, int C, const float* scale,
                         const float* bias, const float* mean, const float* inv, float* dgamma, float* dbeta, float* m1,
                         float* m2, double* partial, cudaStream_t s) {
    Y2_REQUIRE(C <= 3072, "bn_bwd: too many channels");
    RedArgs a = {};
    a.z = z; a.g = g; a.ldg = ldg; a.rows = rows; a.C = C; a.partial = partial;
    a.scale = scale; a.bias = bias; a.mean = mean; a.inv = inv;
    const int nb = red_blocks(rows, C);
    col_reduce_kernel<1><<<nb, RED_THREADS, 0, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    bn_bwd_finish_kernel<<<(C + 31) / 32, FIN_THREADS, 0, s>>>(partial, nb, C, 1.0 / (double)rows, dgamma, dbeta, m1, m2);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int bias_grad_launch(const float* g, long long ldg, size_t rows, int C, float* dbias, double* partial, cudaStream_t s) {
    Y2_REQUIRE(C <= 3072, "bias_grad: too many channels");
    RedArgs a = {};
    a.g = g; a.ldg = ldg; a.rows = rows; a.C = C; a.partial = partial;
    const int nb = red_blocks(rows, C);
    col_reduce_kernel<2><<<nb, RED_THREADS, 0, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    bias_grad_finish_kernel<<<(C + 31) / 32, FIN_THREADS, 0, s>>>(partial, nb, C, dbias);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bn apply (forward): y = leaky(z*scale + bias) -> planes.   4 channels per thread.
// Thread -> fixed channel group (tid % c4) and row lane (tid / c4): per-channel parameters live in registers, the row loop
// has no integer division and keeps two rows of loads in flight.  Requires c4 = C/4 to divide 256 (C <= 1024).
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ bias,
                bf16* __restrict__ y_hi, bf16* __restrict__ y_lo, size_t rows, int C, long long ldy, bf16* __restrict__ y16_hi,
                bf16* __restrict__ y16_lo) {
    const int c4 = C / 4;
    const int c = (int)(threadIdx.x % c4) * 4;
    const int rpb = 256 / c4;
    const size_t rstep = (size_t)gridDim.x * rpb;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c)), bi = __ldg(reinterpret_cast<const float4*>(bias + c));
    auto one = [&](size_t r, const float4& t) {
        float f[4] = {t.x * sc.x + bi.x, t.y * sc.y + bi.y, t.z * sc.z + bi.z, t.w * sc.w + bi.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = fmaxf(f[j], 0.1f * f[j]);
        uint2 h, l;
        split4(f, &h, &l);
        *reinterpret_cast<uint2*>(y_hi + r * ldy + c) = h;
        *reinterpret_cast<uint2*>(y_lo + r * ldy + c) = l;
        if (y16_hi) {                                       // the fp16 planes the next forward conv reads (same pitch)
            split4f(f, &h, &l, 1);
            *reinterpret_cast<uint2*>(y16_hi + r * ldy + c) = h;
            *reinterpret_cast<uint2*>(y16_lo + r * ldy + c) = l;
        }
    };
    size_t r = (size_t)blockIdx.x * rpb + threadIdx.x / c4;
    for (; r + rstep < rows; r += 2 * rstep) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(z + r * C + c));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(z + (r + rstep) * C + c));
        one(r, t0);
        one(r + rstep, t1);
    }
    if (r < rows) one(r, __ldg(reinterpret_cast<const float4*>(z + r * C + c)));
}
static inline int ew_grid(size_t items) {
    size_t b = (items + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}
int bn_apply_launch(const float* z, const float* scale, const float* bias, bf16* y_hi, bf16* y_lo, size_t rows, int C,
                    long long ldy, cudaStream_t s, bf16* y16_hi, bf16* y16_lo) {
    Y2_REQUIRE(C % 4 == 0 && ldy % 4 == 0 && 256 % (C / 4) == 0, "bn_apply: C/4 must divide 256 and the pitch be a multiple of 4");
    bn_apply_kernel<<<ew_grid(rows * (size_t)(C / 4) / 2), 256, 0, s>>>(z, scale, bias, y_hi, y_lo, rows, C, ldy, y16_hi, y16_lo);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// bn backward apply: dx = scale * (dz - m1 - zhat * m2) -> planes (operand of dgrad / wgrad)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ g, long long ldg,
                    const float* __restrict__ scale, const float* __restrict__ bias,
                    const float* __restrict__ mean, const float* __restrict__ inv,
                    const float* __restrict__ m1, const float* __restrict__ m2, bf16* __restrict__ dx_hi,
                    bf16* __restrict__ dx_lo, size_t rows, int C) {
    const int c4 = C / 4;
    const int c = (int)(threadIdx.x % c4) * 4;
    const int rpb = 256 / c4;
    const size_t rstep = (size_t)gridDim.x * rpb;
    float sc[4], bi[4], mu[4], iv[4], q1[4], q2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = __ldg(scale + c + j); bi[j] = __ldg(bias + c + j); mu[j] = __ldg(mean + c + j); iv[j] = __ldg(inv + c + j);
        q1[j] = __ldg(m1 + c + j); q2[j] = __ldg(m2 + c + j);
    }
    const bool vec = (ldg & 3) == 0;
    auto load_g = [&](size_t r) -> float4 {
        if (vec) return __ldg(reinterpret_cast<const float4*>(g + r * ldg + c));
        return make_float4(__ldg(g + r * ldg + c), __ldg(g + r * ldg + c + 1), __ldg(g + r * ldg + c + 2), __ldg(g + r * ldg + c + 3));
    };
    auto one = [&](size_t r, const float4& zt, const float4& gt) {
        const float zv[4] = {zt.x, zt.y, zt.z, zt.w}, gv[4] = {gt.x, gt.y, gt.z, gt.w};
        float f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float zb = zv[j] * sc[j] + bi[j];
            const float dz = gv[j] * (zb >= 0.f ? 1.0f : 0.1f);
            const float zh = (zv[j] - mu[j]) * iv[j];
            f[j] = sc[j] * (dz - q1[j] - zh * q2[j]);
        }
        uint2 h, l;
        split4(f, &h, &l);
        *reinterpret_cast<uint2*>(dx_hi + r * C + c) = h;
        *reinterpret_cast<uint2*>(dx_lo + r * C + c) = l;
    };
    size_t r = (size_t)blockIdx.x * rpb + threadIdx.x / c4;
    for (; r + rstep < rows; r += 2 * rstep) {
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(z + r * C + c));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(z + (r + rstep) * C + c));
        const float4 g0 = load_g(r), g1 = load_g(r + rstep);
        one(r, z0, g0);
        one(r + rstep, z1, g1);
    }
    if (r < rows) one(r, __ldg(reinterpret_cast<const float4*>(z + r * C + c)), load_g(r));
}
int bn_bwd_apply_launch(const float* z, const float* g, long long ldg, const float* scale, const float* bias,
                        const float* mean, const float* inv, const float* m1, const float* m2, bf16* dx_hi, bf16* dx_lo,
                        size_t rows, int C, cudaStream_t s) {
    Y2_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0, "bn_bwd_apply: C/4 must divide 256");
    bn_bwd_apply_kernel<<<ew_grid(rows * (size_t)(C / 4) / 2), 256, 0, s>>>(z, g, ldg, scale, bias, mean, inv, m1, m2, dx_hi, dx_lo,
                                                                       rows, C);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// fp32 [rows][C] (pitch ld) -> planes [rows][Cpad] with zero padding (final layer: 425 -> 512)
__global__ void split_planes_pad_kernel(const float* __restrict__ src, long long ld, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                        size_t rows, int C, int Cpad) {
    const size_t total = rows * (size_t)Cpad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / Cpad;
        const int c = (int)(i - r * Cpad);
        const float v = c < C ? __ldg(src + r * ld + c) : 0.f;
        const bf16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}
int split_planes_pad_launch(const float* src, long long ld, bf16* hi, bf16* lo, size_t rows, int C, int Cpad, cudaStream_t s) {
    split_planes_pad_kernel<<<ew_grid(rows * (size_t)Cpad), 256, 0, s>>>(src, ld, hi, lo, rows, C, Cpad);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// max-pool backward: the first maximum of each 2x2 window (row-major window order, strict >) gets the
// gradient.  Values compared are the exact hi+lo activations.  One thread = 4 channels of one window.
__global__ void maxpool_bwd_kernel(const float* __restrict__ gp, long long ldgp, const bf16* __restrict__ y_hi,
                                   const bf16* __restrict__ y_lo, float* __restrict__ g, int B, int H, int W, int C, int f16) {
    const int Ho = H / 2, Wo = W / 2, c4 = C / 4;
    // thread -> fixed channel group; windows walked with 32-bit index arithmetic (c4 divides 256 or is a multiple of it)
    const unsigned windows = (unsigned)B * Ho * Wo;
    const unsigned cpb = c4 < 256 ? c4 : 256;                // channel groups covered by one block
    const unsigned wpb = 256 / cpb;                          // windows per block per pass
    const unsigned cblocks = (c4 + cpb - 1) / cpb;           // blocks along the channel axis
    const unsigned bx = blockIdx.x % cblocks, by = blockIdx.x / cblocks;
    const int cv = (int)(bx * cpb + threadIdx.x % cpb);
    const unsigned wstep = (gridDim.x / cblocks) * wpb;
    for (unsigned w = by * wpb + threadIdx.x / cpb; w < windows && by < gridDim.x / cblocks; w += wstep) {
        const int xo = (int)(w % Wo);
        const unsigned t = w / Wo;
        const int yo = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const size_t prow = ((size_t)b * Ho + yo) * Wo + xo;
        float gpv[4];
        if ((ldgp & 3) == 0) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(gp + prow * ldgp + cv * 4));
            gpv[0] = q.x; gpv[1] = q.y; gpv[2] = q.z; gpv[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) gpv[j] = __ldg(gp + prow * ldgp + cv * 4 + j);
        }
        float v[4][4];
        size_t off[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            off[q] = (((size_t)b * H + 2 * yo + (q >> 1)) * W + 2 * xo + (q & 1)) * C + (size_t)cv * 4;
            const uint2 h = __ldg(reinterpret_cast<const uint2*>(y_hi + off[q]));
            const uint2 l = __ldg(reinterpret_cast<const uint2*>(y_lo + off[q]));
            merge4f(h, l, v[q], f16);
        }
        float o[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int best = 0;
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (v[q][j] > v[best][j]) best = q;
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q][j] = (q == best) ? gpv[j] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(g + off[q]) = make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
    }
}
int maxpool_bwd_launch(const float* gp, long long ldgp, const bf16* y_hi, const bf16* y_lo, float* g, int B, int H, int W,
                       int C, cudaStream_t s, int f16) {
    Y2_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool_bwd: bad shape");
    Y2_REQUIRE(256 % (C / 4) == 0 || (C / 4) % 256 == 0, "maxpool_bwd: C/4 must divide 256 or be a multiple of it");
    Y2_REQUIRE((size_t)B * (H / 2) * (W / 2) < (1ull << 31), "maxpool_bwd: too many windows");
    maxpool_bwd_kernel<<<ew_grid((size_t)B * (H / 2) * (W / 2) * (C / 4)) / ((C / 4 + 255) / 256) * ((C / 4 + 255) / 256), 256, 0, s>>>(gp, ldgp, y_hi, y_lo, g, B, H, W, C, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// 2x2 stride-1 SAME max-pool backward (tiny, model/yolo2/inference.py:42).  Window (wy, wx) covers the pixels
// (wy .. wy+1, wx .. wx+1) clipped to the image (TF pads bottom / right only and the maximum ignores padding: the clipped taps
// re-read a valid pixel and, compared with strict >, never win); its gradient goes to the FIRST maximum in window order.
// Gather form, no atomics: one thread = 4 channels of one INPUT pixel; it re-derives the winner of each of the (up to) four
// windows that contain the pixel and adds the gradients of the ones it wins.
__global__ void maxpool_s1_bwd_kernel(const float* __restrict__ gp, long long ldgp, const bf16* __restrict__ y_hi,
                                      const bf16* __restrict__ y_lo, float* __restrict__ g, int B, int H, int W, int C, int f16) {
    const int c4 = C / 4;
    const size_t total = (size_t)B * H * W * c4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % c4);
        size_t t = i / c4;
        const int x = (int)(t % W);
        t /= W;
        const int y = (int)(t % H);
        const int b = (int)(t / H);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int wy = y - 1; wy <= y; ++wy) {
            if (wy < 0) continue;
            for (int wx = x - 1; wx <= x; ++wx) {
                if (wx < 0) continue;
                const int dy = (wy + 1 < H) ? 1 : 0, dx = (wx + 1 < W) ? 1 : 0;
                float v[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const size_t off = (((size_t)b * H + wy + (q >> 1) * dy) * W + wx + (q & 1) * dx) * C + (size_t)cv * 4;
                    merge4f(__ldg(reinterpret_cast<const uint2*>(y_hi + off)), __ldg(reinterpret_cast<const uint2*>(y_lo + off)), v[q], f16);
                }
                const size_t prow = ((size_t)b * H + wy) * W + wx;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int best = 0;
#pragma unroll
                    for (int q = 1; q < 4; ++q)
                        if (v[q][j] > v[best][j]) best = q;
                    const int py = wy + (best >> 1) * dy, px = wx + (best & 1) * dx;
                    if (py == y && px == x) acc[j] += __ldg(gp + prow * ldgp + cv * 4 + j);
                }
            }
        }
        *reinterpret_cast<float4*>(g + (((size_t)b * H + y) * W + x) * C + (size_t)cv * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}
int maxpool_s1_bwd_launch(const float* gp, long long ldgp, const bf16* y_hi, const bf16* y_lo, float* g, int B, int H, int W, int C,
                          cudaStream_t s, int f16) {
    Y2_REQUIRE(C % 4 == 0, "maxpool_s1_bwd: C must be a multiple of 4");
    maxpool_s1_bwd_kernel<<<ew_grid((size_t)B * H * W * (C / 4)), 256, 0, s>>>(gp, ldgp, y_hi, y_lo, g, B, H, W, C, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// planes [rows][C] -> planes [rows][Cpad] with zero-filled columns (the weight-gradient GEMM reads its gradient operand in
// whole 64-channel atoms; tiny's 32-channel conv1 is the one layer whose dx planes are narrower)
__global__ void repitch_planes_kernel(const bf16* __restrict__ src_hi, const bf16* __restrict__ src_lo, bf16* __restrict__ dst_hi,
                                      bf16* __restrict__ dst_lo, size_t rows, int C, int Cpad) {
    const size_t total = rows * (size_t)Cpad;
    const bf16 zero = __float2bfloat16_rn(0.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / Cpad;
        const int c = (int)(i - r * Cpad);
        dst_hi[i] = c < C ? src_hi[r * C + c] : zero;
        dst_lo[i] = c < C ? src_lo[r * C + c] : zero;
    }
}
int repitch_planes_launch(const bf16* src_hi, const bf16* src_lo, bf16* dst_hi, bf16* dst_lo, size_t rows, int C, int Cpad, cudaStream_t s) {
    repitch_planes_kernel<<<ew_grid(rows * (size_t)Cpad), 256, 0, s>>>(src_hi, src_lo, dst_hi, dst_lo, rows, C, Cpad);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// gradient of a variable whose STORED shape is padded (tiny: conv0's 16 outputs / conv1's 16 inputs live in 32 channels):
// [taps][cin_s][cout_s] -> the variable's own [taps][cin][cout]
__global__ void compact_hwio_kernel(const float* __restrict__ src, float* __restrict__ dst, int taps, int cin_s, int cout_s, int cin, int cout) {
    const size_t total = (size_t)taps * cin * cout;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % cout);
        const size_t r = i / cout;
        const int c = (int)(r % cin);
        const size_t tap = r / cin;
        dst[i] = __ldg(src + (tap * cin_s + c) * cout_s + n);
    }
}
int compact_hwio_launch(const float* src, float* dst, int taps, int cin_s, int cout_s, int cin, int cout, cudaStream_t s) {
    compact_hwio_kernel<<<ew_grid((size_t)taps * cin * cout), 256, 0, s>>>(src, dst, taps, cin_s, cout_s, cin, cout);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Pool layers without a passthrough (conv0, conv1, conv4, conv7): the un-pooled activation and its gradient are never
// materialised.  Everything a 2x2 window needs is a function of the stored raw conv output z: y = leaky(z*scale+bias),
// rounded to the split-plane value hi+lo the inference path would have stored, first maximum in row-major window order
// (strict >) wins -- forward writes the winner's planes, backward routes the pooled gradient to the same pixel.
// One thread = one window x 4 channels; thread -> fixed channel group, 32-bit window arithmetic.
struct PoolWin {
    float z[4][4];      // raw conv output of the 4 window pixels
    float yb[4][4];     // z*scale + bias (sign decides the leaky slope)
    int best[4];        // winning pixel per channel
    float zw[4], ybw[4];   // z and z*scale+bias of the winner
    size_t off[4];      // element offsets of the 4 pixels (row pitch C)
};
// f16: the plane format whose rounding decides the winner (the format the NEXT forward conv reads its input in)
__device__ __forceinline__ void pool_window_load(PoolWin& w, const float* __restrict__ z, const float (&sc)[4], const float (&bi)[4],
                                                 int b, int yo, int xo, int H, int W, int C, int c, int f16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        w.off[q] = (((size_t)b * H + 2 * yo + (q >> 1)) * W + 2 * xo + (q & 1)) * C + c;
        const float4 t = __ldg(reinterpret_cast<const float4*>(z + w.off[q]));
        w.z[q][0] = t.x; w.z[q][1] = t.y; w.z[q][2] = t.z; w.z[q][3] = t.w;
    }
    float v[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w.yb[q][j] = w.z[q][j] * sc[j] + bi[j];
            y[j] = fmaxf(w.yb[q][j], 0.1f * w.yb[q][j]);
        }
        uint2 h, l;
        split4f(y, &h, &l, f16);
        merge4f(h, l, v[q], f16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int best = 0;
        float bv = v[0][j], bz = w.z[0][j], byb = w.yb[0][j];
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            const bool gt = v[q][j] > bv;
            bv = gt ? v[q][j] : bv; bz = gt ? w.z[q][j] : bz; byb = gt ? w.yb[q][j] : byb; best = gt ? q : best;
        }
        w.best[j] = best; w.zw[j] = bz; w.ybw[j] = byb;
    }
}
struct PoolIdx {
    unsigned windows, wstep, w0;
    int c, Ho, Wo;
};
__device__ __forceinline__ PoolIdx pool_index(int B, int H, int W, int C) {
    PoolIdx ix;
    const int c4 = C / 4;
    ix.Ho = H / 2; ix.Wo = W / 2;
    ix.windows = (unsigned)B * ix.Ho * ix.Wo;
    ix.c = (int)(threadIdx.x % c4) * 4;
    const unsigned wpb = 256 / c4;
    ix.wstep = gridDim.x * wpb;
    ix.w0 = blockIdx.x * wpb + threadIdx.x / c4;
    return ix;
}

// forward: z -> pooled planes
__global__ void __launch_bounds__(256)
bn_apply_pool_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ bias,
                     bf16* __restrict__ p_hi, bf16* __restrict__ p_lo, int B, int H, int W, int C, bf16* __restrict__ p16_hi,
                     bf16* __restrict__ p16_lo) {
    const int f16 = p16_hi ? 1 : 0;
    const PoolIdx ix = pool_index(B, H, W, C);
    float sc[4], bi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sc[j] = __ldg(scale + ix.c + j); bi[j] = __ldg(bias + ix.c + j); }
    for (unsigned w = ix.w0; w < ix.windows; w += ix.wstep) {
        const int xo = (int)(w % ix.Wo);
        const unsigned t = w / ix.Wo;
        const int yo = (int)(t % ix.Ho), b = (int)(t / ix.Ho);
        PoolWin pw;
        pool_window_load(pw, z, sc, bi, b, yo, xo, H, W, C, ix.c, f16);
        float yw[4];                                         // the winner's activation; its planes are a pure function of it
#pragma unroll
        for (int j = 0; j < 4; ++j) yw[j] = fmaxf(pw.ybw[j], 0.1f * pw.ybw[j]);
        const size_t po = (size_t)w * C + ix.c;
        uint2 h, l;
        split4(yw, &h, &l);
        *reinterpret_cast<uint2*>(p_hi + po) = h;
        *reinterpret_cast<uint2*>(p_lo + po) = l;
        if (f16) {
            split4f(yw, &h, &l, 1);
            *reinterpret_cast<uint2*>(p16_hi + po) = h;
            *reinterpret_cast<uint2*>(p16_lo + po) = l;
        }
    }
}
int bn_apply_pool_launch(const float* z, const float* scale, const float* bias, bf16* p_hi, bf16* p_lo, int B, int H, int W, int C,
                         cudaStream_t s, bf16* p16_hi, bf16* p16_lo) {
    Y2_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0 && H % 2 == 0 && W % 2 == 0, "bn_apply_pool: bad shape");
    Y2_REQUIRE((size_t)B * (H / 2) * (W / 2) < (1ull << 31), "bn_apply_pool: too many windows");
    bn_apply_pool_kernel<<<ew_grid((size_t)B * (H / 2) * (W / 2) * (C / 4)), 256, 0, s>>>(z, scale, bias, p_hi, p_lo, B, H, W, C, p16_hi, p16_lo);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// backward: APPLY = false -> per-block partial sums of dz and dz*zhat (only the winning pixel of a window has dz != 0);
//           APPLY = true  -> dx = scale * (dz - m1 - zhat * m2) for all 4 pixels -> planes.   gp = pooled gradient, pitch ldgp.
//           gy_out (optional, tests): the un-pooled dL/dy this layer would have received.
template <bool APPLY>
__global__ void __launch_bounds__(256)
bn_bwd_pool_kernel(const float* __restrict__ z, const float* __restrict__ gp, long long ldgp, const float* __restrict__ scale,
                   const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ inv,
                   const float* __restrict__ m1, const float* __restrict__ m2, bf16* __restrict__ dx_hi, bf16* __restrict__ dx_lo,
                   double* __restrict__ partial, float* __restrict__ gy_out, int B, int H, int W, int C, int f16) {
    const PoolIdx ix = pool_index(B, H, W, C);
    const int c4 = C / 4;
    float sc[4], bi[4], mu[4], iv[4], q1[4] = {0, 0, 0, 0}, q2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = __ldg(scale + ix.c + j); bi[j] = __ldg(bias + ix.c + j); mu[j] = __ldg(mean + ix.c + j); iv[j] = __ldg(inv + ix.c + j);
        if (APPLY) { q1[j] = __ldg(m1 + ix.c + j); q2[j] = __ldg(m2 + ix.c + j); }
    }
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
    int flush = 0;
    for (unsigned w = ix.w0; w < ix.windows; w += ix.wstep) {
        const int xo = (int)(w % ix.Wo);
        const unsigned t = w / ix.Wo;
        const int yo = (int)(t % ix.Ho), b = (int)(t / ix.Ho);
        float g[4];
        if ((ldgp & 3) == 0) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(gp + (size_t)w * ldgp + ix.c));
            g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = __ldg(gp + (size_t)w * ldgp + ix.c + j);
        }
        PoolWin pw;
        pool_window_load(pw, z, sc, bi, b, yo, xo, H, W, C, ix.c, f16);
        if (!APPLY) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dz = g[j] * (pw.ybw[j] >= 0.f ? 1.0f : 0.1f);
                s1[j] += dz; s2[j] += dz * ((pw.zw[j] - mu[j]) * iv[j]);
            }
            if (++flush == 16) {                      // the same fp32 run length (64 pixels) as the dense reduction
#pragma unroll
                for (int j = 0; j < 4; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
                flush = 0;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float f[4], gq[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gq[j] = (q == pw.best[j]) ? g[j] : 0.f;
                    const float dz = gq[j] * (pw.yb[q][j] >= 0.f ? 1.0f : 0.1f);
                    const float zh = (pw.z[q][j] - mu[j]) * iv[j];
                    f[j] = sc[j] * (dz - q1[j] - zh * q2[j]);
                }
                uint2 h, l;
                split4(f, &h, &l);
                *reinterpret_cast<uint2*>(dx_hi + pw.off[q]) = h;
                *reinterpret_cast<uint2*>(dx_lo + pw.off[q]) = l;
                if (gy_out) *reinterpret_cast<float4*>(gy_out + pw.off[q]) = make_float4(gq[0], gq[1], gq[2], gq[3]);
            }
        }
    }
    if (!APPLY) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; }
        __shared__ double sm1[256][4], sm2[256][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { sm1[threadIdx.x][j] = d1[j]; sm2[threadIdx.x][j] = d2[j]; }
        __syncthreads();
        if ((int)threadIdx.x < c4) {
            double t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
            for (int rl = 0; rl < 256 / c4; ++rl)
#pragma unroll
                for (int j = 0; j < 4; ++j) { t1[j] += sm1[rl * c4 + threadIdx.x][j]; t2[j] += sm2[rl * c4 + threadIdx.x][j]; }
            const int cc = (int)threadIdx.x * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                partial[((size_t)blockIdx.x * 2 + 0) * C + cc + j] = t1[j];
                partial[((size_t)blockIdx.x * 2 + 1) * C + cc + j] = t2[j];
            }
        }
    }
}
int bn_bwd_pool_launch(const float* z, const float* gp, long long ldgp, int B, int H, int W, int C, const float* scale, const float* bias,
                       const float* mean, const float* inv, float* dgamma, float* dbeta, float* m1, float* m2, bf16* dx_hi,
                       bf16* dx_lo, double* partial, float* gy_out, cudaStream_t s, int f16) {
    Y2_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0 && H % 2 == 0 && W % 2 == 0 && C <= 1024, "bn_bwd_pool: bad shape");
    const size_t windows = (size_t)B * (H / 2) * (W / 2);
    Y2_REQUIRE(windows < (1ull << 31), "bn_bwd_pool: too many windows");
    int nb = red_blocks(windows * 4, C);
    bn_bwd_pool_kernel<false><<<nb, 256, 0, s>>>(z, gp, ldgp, scale, bias, mean, inv, nullptr, nullptr, nullptr, nullptr, partial, nullptr,
                                                  B, H, W, C, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    bn_bwd_finish_kernel<<<(C + 31) / 32, FIN_THREADS, 0, s>>>(partial, nb, C, 1.0 / (double)(windows * 4), dgamma, dbeta, m1, m2);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    bn_bwd_pool_kernel<true><<<ew_grid(windows * (size_t)(C / 4)), 256, 0, s>>>(z, gp, ldgp, scale, bias, mean, inv, m1, m2, dx_hi, dx_lo,
                                                                               nullptr, gy_out, B, H, W, C, f16);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// reorg backward (accumulating): g[b, 2y+dy, 2x+dx, c] += gr[b, y, x, (dy*2+dx)*C + c], gr pitch ldr.
__global__ void reorg_bwd_add_kernel(const float* __restrict__ gr, long long ldr, float* __restrict__ g, int B, int H, int W,
                                     int C) {
    const int Ho = H / 2, Wo = W / 2, c4 = C / 4;
    const size_t total = (size_t)B * H * W * c4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % c4);
        size_t t = i / c4;
        const int x = (int)(t % W);
        t /= W;
        const int y = (int)(t % H);
        const int b = (int)(t / H);
        const int d = (y & 1) * 2 + (x & 1);
        const size_t src = (((size_t)b * Ho + (y >> 1)) * Wo + (x >> 1)) * ldr + (size_t)d * C + cv * 4;
        const float4 a = __ldg(reinterpret_cast<const float4*>(gr + src));
        float4* dst = reinterpret_cast<float4*>(g + ((((size_t)b * H + y) * W + x) * C + cv * 4));
        float4 o = *dst;
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        *dst = o;
    }
}
int reorg_bwd_add_launch(const float* gr, long long ldr, float* g, int B, int H, int W, int C, cudaStream_t s) {
    Y2_REQUIRE(C % 4 == 0 && ldr % 4 == 0, "reorg_bwd: C and pitch must be multiples of 4");
    reorg_bwd_add_kernel<<<ew_grid((size_t)B * H * W * (C / 4)), 256, 0, s>>>(gr, ldr, g, B, H, W, C);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dgrad weights: Wd[n = cin][tap' = 8 - tap][k = cout (padded)] from W HWIO [tap][cin][cout]  (180-degree
// rotated taps, in/out channels swapped), split planes [2][cin_pad][taps*cout_pad].
__global__ void pack_dgrad_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int taps, int cin, int cout,
                                          int cin_pad, int cout_pad) {
    const size_t K = (size_t)taps * cout_pad;
    const size_t total = (size_t)cin_pad * K;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / K;                 // dgrad output channel = forward input channel
        const size_t k = i - n * K;
        const int tapd = (int)(k / cout_pad);   // dgrad tap
        const int co = (int)(k - (size_t)tapd * cout_pad);
        float v = 0.f;
        if (n < (size_t)cin && co < cout) v = __ldg(w + ((size_t)(taps - 1 - tapd) * cin + n) * cout + co);
        const bf16 h = __float2bfloat16_rn(v);
        out[i] = h;
        out[total + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}
int pack_dgrad_weights_launch(const float* w_hwio, bf16* out, int ksize, int cin, int cout, int cin_pad, int cout_pad,
                              cudaStream_t s) {
    const size_t total = (size_t)cin_pad * ksize * ksize * cout_pad;
    pack_dgrad_weights_kernel<<<ew_grid(total), 256, 0, s>>>(w_hwio, out, ksize * ksize, cin, cout, cin_pad, cout_pad);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
