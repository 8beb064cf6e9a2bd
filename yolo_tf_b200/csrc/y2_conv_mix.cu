// EXPERIMENTAL (not on any product path, not used by y2_darknet_forward, never run on a GPU at the time of writing):
// one conv as an implicit GEMM whose products cost 2 MMA-equivalents instead of the 3 bf16 MMAs of conv_tc_kernel.
//
//   x*w ~= X16*W16 + X8*RW8 + RX8*W8       (tests/test_numerics_candidate_fp16_fp8.py)
//   X16 = fp16(x E16)   X8 = e4m3(x E8)   RX8 = e4m3((x E16 - X16) * 4096 E8 / E16)          (weights likewise: F16, F8)
//
// With E16 / E8 = 2^7 and F16 / F8 = 2^5 all three products carry the same power-of-two factor E16 * F16, so ONE fp32 TMEM
// accumulator takes the kind::f16 MMAs of the main term and the kind::f8f6f4 MMAs (half the cycles per product) of the
// two corrections; the epilogue multiplies by 1 / (E16 F16).  Operand bytes per k-element read from shared memory:
// 2 + 1 + 1 on each side instead of 2 + 2 + 2.
//
// Scope of this file: the questions that need hardware -- do the two MMA kinds accumulate into one TMEM tile, do they
// interleave at full rate, what does the accumulator truncation add -- behind a diagnostic entry point (y2_conv2d_mix)
// that converts float32 operands on the fly.  Deliberately minimal: single CTA, linear 128-pixel tiles (im2col TMA),
// whole-K data-parallel tiles, fp32 output, an optional accumulation-chain cap (sub-results summed in the output tile by the
// epilogue), no stream-K / pairs / pooled epilogue.  The conv it stands for is
// the same slim.layers.conv2d (+ folded batch_norm + leaky_relu) of model/yolo2/inference.py:62-69,73-118.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <string.h>

#include "../../include/yolo2_b200.h"
#include "y2_internal.h"
#include "y2_mix_prep.cuh"
#include "y2_ptx.cuh"

namespace y2 {

static constexpr int MX_BLOCK_M = 128;
static constexpr int MX_BK = 64;                 // channels per k-block: 128 B of fp16 (SW128), 64 B of e4m3 (SW64)
static constexpr int MX_THREADS = 256;
static constexpr int MX_EPI_WARP0 = 4;
static constexpr int MX_ACC_COLS = 256;
static constexpr int MX_TMEM_COLS = 512;
static constexpr int MX_SMEM_LIMIT = 227 * 1024;
static constexpr int MX_BAR_BYTES = 256;
static constexpr int MX_A16 = MX_BLOCK_M * MX_BK * 2;      // 16 KiB
static constexpr int MX_A8 = MX_BLOCK_M * MX_BK;           //  8 KiB

struct MixParams {
    int M, N, Cin, ksize, B, H, W;
    int block_n, m_tiles, n_tiles, kblocks, cout_pad, num_stages;
    int terms;               // bit 0: X16*W16, bit 1: X8*RW8, bit 2: RX8*W8 (diagnostics: time / check the terms separately)
    int nchunks;             // accumulation-chain cap: K is cut into nchunks equal chains, each summed from zero in its own TMEM
                             // buffer; the epilogue adds the sub-results in fp32 (round to nearest) in the output tile itself
    int leaky;
    float unscale;           // 1 / (E16 * F16)
    const float* scale;      // [N] folded BN scale (null -> 1)
    const float* bias;       // [N] (null -> 0)
    float* out;              // [M][ldc] float32 (nullable when nchunks == 1 and the split outputs are written)
    long long ldc;
    // optional second output: the result in the storage format the NEXT mixed-kind conv reads (N % 32 == 0, row pitch N)
    __half* o16;             // fp16(y oE16)
    uint8_t* o8;             // e4m3(y oE8)
    uint8_t* or8;            // e4m3((y oE16 - o16) * ora)
    float oE16, oE8, ora;
    unsigned int* amax_out;  // atomicMax of |y| as float bits: what the next layer derives its bound from
};

// same operand form as tcgen05.mma kind::f16; A and B are e4m3 (format 0 / 0 in the instruction descriptor), K = 32
__device__ __forceinline__ void tc_mma_f8_e(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred pe, p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 pe, %7, 0;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
// D = f32, A = B = format 0 (F16 for kind::f16, E4M3 for kind::f8f6f4), both K-major
__host__ __device__ __forceinline__ uint32_t make_idesc_fmt0(uint32_t m, uint32_t n) {
    return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__global__ void __launch_bounds__(MX_THREADS, 1)
conv_mix_kernel(const __grid_constant__ CUtensorMap map_a16, const __grid_constant__ CUtensorMap map_a8,
                const __grid_constant__ CUtensorMap map_ra8, const __grid_constant__ CUtensorMap map_w16,
                const __grid_constant__ CUtensorMap map_rw8, const __grid_constant__ CUtensorMap map_w8, const MixParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b16 = p.block_n * MX_BK * 2, b8 = p.block_n * MX_BK;
    // stage = [A16 | A8 | RA8 | W16 | RW8 | W8]; every piece is a multiple of 1 KiB (block_n % 16 == 0)
    const int stage_bytes = MX_A16 + 2 * MX_A8 + b16 + 2 * b8;
    const int S = p.num_stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a16); tma_prefetch_desc(&map_a8); tma_prefetch_desc(&map_ra8);
        tma_prefetch_desc(&map_w16); tma_prefetch_desc(&map_rw8); tma_prefetch_desc(&map_w8);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, MX_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int KB = p.kblocks;
    const int cblocks = p.Cin / MX_BK;
    const int pad = p.ksize / 2;
    const int hw = p.H * p.W;
    const int tiles = p.m_tiles * p.n_tiles;

    if (warp == 0) {
        // ===================== TMA producer =====================
        const uint32_t leader = elect_one() ? 1u : 0u;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
            const int m0 = mt * MX_BLOCK_M;
            int img = m0 / hw;
            const int rem = m0 - img * hw;
            int y0 = rem / p.W;
            int x0 = rem - y0 * p.W;
            x0 = __shfl_sync(0xffffffffu, x0, 0); y0 = __shfl_sync(0xffffffffu, y0, 0); img = __shfl_sync(0xffffffffu, img, 0);
            const int n0 = __shfl_sync(0xffffffffu, nt * p.block_n, 0);
            int tap = 0, cb = 0;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1u, 0x100u + stage);
                __syncwarp();
                stage = __shfl_sync(0xffffffffu, stage, 0);
                tap = __shfl_sync(0xffffffffu, tap, 0); cb = __shfl_sync(0xffffffffu, cb, 0);
                uint8_t* st = smem + (size_t)stage * stage_bytes;
                const int c0 = cb * MX_BK;
                const int dy = (p.ksize == 3) ? tap / 3 : 0;
                const int dx = (p.ksize == 3) ? tap - dy * 3 : 0;
                const int kcoord = tap * p.Cin + c0;
                mbar_expect_tx_e(leader, &full[stage], (uint32_t)stage_bytes);
                tma_load_im2col_4d_e(leader, st, &map_a16, &full[stage], c0, x0 - pad, y0 - pad, img, (uint16_t)dx, (uint16_t)dy);
                tma_load_im2col_4d_e(leader, st + MX_A16, &map_a8, &full[stage], c0, x0 - pad, y0 - pad, img, (uint16_t)dx, (uint16_t)dy);
                tma_load_im2col_4d_e(leader, st + MX_A16 + MX_A8, &map_ra8, &full[stage], c0, x0 - pad, y0 - pad, img, (uint16_t)dx, (uint16_t)dy);
                uint8_t* sb = st + MX_A16 + 2 * MX_A8;
                tma_load_2d_e(leader, sb, &map_w16, &full[stage], kcoord, n0);
                tma_load_2d_e(leader, sb + b16, &map_rw8, &full[stage], kcoord, n0);
                tma_load_2d_e(leader, sb + b16 + b8, &map_w8, &full[stage], kcoord, n0);
                if (++cb == cblocks) { cb = 0; ++tap; }
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc = make_idesc_fmt0(MX_BLOCK_M, (uint32_t)p.block_n);
        const uint32_t ring = smem_u32(smem);
        const uint32_t h128 = (uint32_t)(make_kmajor_desc(0, 128) >> 32), h64 = (uint32_t)(make_kmajor_desc(0, 64) >> 32);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
          for (int ch = 0; ch < p.nchunks; ++ch) {
            const int kb_lo = (int)((long long)KB * ch / p.nchunks), kb_hi = (int)((long long)KB * (ch + 1) / p.nchunks);
            mbar_wait(&tempty[acc], acc_phase ^ 1u, 0x200u + acc);
            __syncwarp();
            tc_fence_after();
            const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(acc * MX_ACC_COLS), 0);
            uint32_t have = 0;                               // 0 until the first MMA of this chain has been issued
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                mbar_wait(&full[stage], phase, 0x300u + stage);
                __syncwarp();
                tc_fence_after();
                stage = __shfl_sync(0xffffffffu, stage, 0);
                const uint32_t st = ring + (uint32_t)(stage * stage_bytes);
                const uint32_t da16 = (uint32_t)make_kmajor_desc(st, 128);
                const uint32_t da8 = (uint32_t)make_kmajor_desc(st + MX_A16, 64);
                const uint32_t dra8 = (uint32_t)make_kmajor_desc(st + MX_A16 + MX_A8, 64);
                const uint32_t sb = st + MX_A16 + 2 * MX_A8;
                const uint32_t dw16 = (uint32_t)make_kmajor_desc(sb, 128);
                const uint32_t drw8 = (uint32_t)make_kmajor_desc(sb + (uint32_t)b16, 64);
                const uint32_t dw8 = (uint32_t)make_kmajor_desc(sb + (uint32_t)(b16 + b8), 64);
                if (p.terms & 1) {
#pragma unroll
                    for (int k = 0; k < MX_BK / 16; ++k) {   // 16 halves = 32 B per MMA
                        tc_mma_f16_e(leader, d_tmem, da16 + (uint32_t)(k * 2), h128, dw16 + (uint32_t)(k * 2), h128, idesc, have);
                        have = 1u;
                    }
                }
                if (p.terms & 2) {
#pragma unroll
                    for (int k = 0; k < MX_BK / 32; ++k) {   // 32 e4m3 = 32 B per MMA
                        tc_mma_f8_e(leader, d_tmem, da8 + (uint32_t)(k * 2), h64, drw8 + (uint32_t)(k * 2), h64, idesc, have);
                        have = 1u;
                    }
                }
                if (p.terms & 4) {
#pragma unroll
                    for (int k = 0; k < MX_BK / 32; ++k) {
                        tc_mma_f8_e(leader, d_tmem, dra8 + (uint32_t)(k * 2), h64, dw8 + (uint32_t)(k * 2), h64, idesc, have);
                        have = 1u;
                    }
                }
                tc_commit_e(leader, &empty[stage]);
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            tc_commit_e(leader, &tfull[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
    } else if (warp >= MX_EPI_WARP0) {
        // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
        const int q = warp - MX_EPI_WARP0;
        int acc = 0;
        uint32_t acc_phase = 0;
        float amax_local = 0.f;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
            const int n0 = nt * p.block_n;
            const long long row = (long long)mt * MX_BLOCK_M + q * 32 + lane;
            for (int ch = 0; ch < p.nchunks; ++ch) {
                const bool first = ch == 0, last = ch == p.nchunks - 1;
                mbar_wait(&tfull[acc], acc_phase, 0x400u + acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MX_ACC_COLS);
                for (int c = 0; c < p.block_n; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + (uint32_t)c, v);
                    tmem_ld_wait_dep(v);
                    if (row < p.M) {
                        // the running sum of the earlier chains lives in the output tile itself (written and read back by this
                        // thread only); the last chain adds it, then applies scale / bias / leaky
                        float* dst = p.out ? p.out + (size_t)row * p.ldc + n0 + c : nullptr;
                        float t[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + c + j;
                            t[j] = 0.f;
                            if (n < p.N) {
                                float a = __uint_as_float(v[j]) * p.unscale;      // exact: a power of two
                                if (!first) a += dst[j];
                                if (last) {
                                    a = fmaf(a, p.scale ? __ldg(p.scale + n) : 1.0f, p.bias ? __ldg(p.bias + n) : 0.0f);
                                    a = p.leaky ? fmaxf(a, 0.1f * a) : a;
                                }
                                if (dst) dst[j] = a;
                                t[j] = a;
                            }
                        }
                        if (last && p.o16) {                     // (N % 32 == 0: whole chunks only) 64 + 32 + 32 bytes per row
                            uint32_t h[16], q8[8], r8[8];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float a0 = t[2 * j] * p.oE16, a1 = t[2 * j + 1] * p.oE16;
                                const __half2 hh = __floats2half2_rn(a0, a1);
                                const float2 back = __half22float2(hh);
                                h[j] = *reinterpret_cast<const uint32_t*>(&hh);
                                const uint32_t qq = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(t[2 * j] * p.oE8, t[2 * j + 1] * p.oE8), __NV_SATFINITE, __NV_E4M3);
                                const uint32_t rr = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((a0 - back.x) * p.ora, (a1 - back.y) * p.ora), __NV_SATFINITE, __NV_E4M3);
                                if (j & 1) { q8[j >> 1] |= qq << 16; r8[j >> 1] |= rr << 16; }
                                else { q8[j >> 1] = qq; r8[j >> 1] = rr; }
                                amax_local = fmaxf(amax_local, fmaxf(fabsf(t[2 * j]), fabsf(t[2 * j + 1])));
                            }
                            const size_t off = (size_t)row * p.N + n0 + c;
                            uint4* d16 = reinterpret_cast<uint4*>(p.o16 + off);
#pragma unroll
                            for (int j = 0; j < 4; ++j) d16[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                            uint4* d8 = reinterpret_cast<uint4*>(p.o8 + off);
                            uint4* dr = reinterpret_cast<uint4*>(p.or8 + off);
                            d8[0] = make_uint4(q8[0], q8[1], q8[2], q8[3]); d8[1] = make_uint4(q8[4], q8[5], q8[6], q8[7]);
                            dr[0] = make_uint4(r8[0], r8[1], r8[2], r8[3]); dr[1] = make_uint4(r8[4], r8[5], r8[6], r8[7]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
        if (p.amax_out) {                                    // one atomic per warp per kernel, not per tile
            for (int o = 16; o > 0; o >>= 1) amax_local = fmaxf(amax_local, __shfl_xor_sync(0xffffffffu, amax_local, o));
            if (lane == 0) atomicMax(p.amax_out, __float_as_uint(amax_local));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, MX_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ operand preparation
__global__ void mix_amax_kernel(const float* __restrict__ x, size_t n, unsigned int* amax_bits) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(amax_bits, __float_as_uint(m));      // non-negative floats order like their bit patterns
}
// x (fp32) -> X16, X8, RX8 (y2_mix_prep.cuh)
__global__ void mix_prep_act_kernel(const float* __restrict__ x, size_t n, float s16, float s8, float rs, __half* x16, uint8_t* x8,
                                    uint8_t* rx8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        mix_split(x[i], s16, s8, rs, x16 + i, x8 + i, rx8 + i);
}
// w HWIO (fp32) -> [cout_pad][K = tap * Cin + c] (K-major B operand, the layout pack_weights_kernel produces), rows >= cout zero
__global__ void mix_prep_w_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cout_pad, float s16, float s8, float rs,
                                  __half* w16, uint8_t* w8, uint8_t* rw8) {
    const size_t K = (size_t)taps * cin, total = K * cout_pad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / K);
        const size_t k = i - (size_t)n * K;
        mix_split(n < cout ? w[k * cout + n] : 0.f, s16, s8, rs, w16 + i, w8 + i, rw8 + i);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_mx_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_mx_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_mx_encodeTiled mx_encodeTiled = nullptr;
static PFN_mx_encodeIm2col mx_encodeIm2col = nullptr;
static float g_mix_last_ms = 0.f;

static int mx_load_entry_points() {
    if (mx_encodeTiled && mx_encodeIm2col) return 0;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    Y2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    Y2_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    mx_encodeTiled = reinterpret_cast<PFN_mx_encodeTiled>(fn);
    fn = nullptr;
    Y2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    Y2_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col unavailable");
    mx_encodeIm2col = reinterpret_cast<PFN_mx_encodeIm2col>(fn);
    return 0;
}

// im2col map over an NHWC tensor of `esize`-byte elements: 128 pixels x 64 channels per load, zero fill = SAME padding
static int mx_map_act(CUtensorMap* m, void* base, CUtensorMapDataType dt, int esize, int B, int H, int W, int Cin, int ksize) {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * esize, (cuuint64_t)W * Cin * esize, (cuuint64_t)H * W * Cin * esize};
    const int padv = ksize / 2;
    int lower[2] = {-padv, -padv};
    int upper[2] = {padv - (ksize - 1), padv - (ksize - 1)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = mx_encodeIm2col(m, dt, 4, base, dims, strides, lower, upper, (cuuint32_t)MX_BK, (cuuint32_t)MX_BLOCK_M, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, esize == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    Y2_REQUIRE(r == CUDA_SUCCESS, "conv mix: cuTensorMapEncodeIm2col failed (%d), element size %d", (int)r, esize);
    int drv = 0;
    cudaDriverGetVersion(&drv);                      // same driver workaround as tc_conv_plan (tensors below 128 KiB)
    if (drv <= 13010 && (size_t)B * H * W * Cin * esize < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
    return 0;
}
static int mx_map_w(CUtensorMap* m, void* base, CUtensorMapDataType dt, int esize, size_t K, int cout_pad, int block_n) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)K * esize};
    cuuint32_t box[2] = {(cuuint32_t)MX_BK, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = mx_encodeTiled(m, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                esize == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    Y2_REQUIRE(r == CUDA_SUCCESS, "conv mix: cuTensorMapEncodeTiled failed (%d), element size %d", (int)r, esize);
    return 0;
}

static bool g_mix_used = false;      // set by the first y2_conv2d_mix launch: processes that never call it never touch this module
static int conv_mix_check_watchdog_impl() {
    if (!g_mix_used) return 0;
    Watchdog w;
    Y2_CUDA(cudaMemcpyFromSymbol(&w, g_watchdog, sizeof(w)));
    if (!w.fired) return 0;
    Watchdog z;
    memset(&z, 0, sizeof(z));
    cudaMemcpyToSymbol(g_watchdog, &z, sizeof(z));
    set_error("conv mix: barrier watchdog fired (block %u warp %u wait-site 0x%x parity %u)", w.block, w.warp, w.tag, w.parity);
    return -3;
}
int conv_mix_check_watchdog() { return conv_mix_check_watchdog_impl(); }

}  // namespace y2

using namespace y2;

// Shared body: activations already in the storage format (scales fixed by `in_bound` >= their amax), weights converted here.
static int mix_run(const char* who, const __half* x16, const uint8_t* x8, const uint8_t* rx8, float in_bound, int B, int H, int W, int cin,
                   const float* w_hwio, int ksize, int cout, const float* scale, const float* bias, int leaky, float* y, __half* o16,
                   uint8_t* o8, uint8_t* or8, float out_bound, unsigned int* amax_out, int terms, int kcap, int block_n, cudaStream_t s) {
    Y2_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "%s: bad shape B=%d H=%d W=%d cin=%d cout=%d", who, B, H, W, cin, cout);
    Y2_REQUIRE(ksize == 1 || ksize == 3, "%s: ksize must be 1 or 3 (got %d)", who, ksize);
    Y2_REQUIRE(cin % MX_BK == 0, "%s: cin must be a multiple of 64 (got %d)", who, cin);
    Y2_REQUIRE(terms >= 1 && terms <= 7, "%s: terms is a bit mask 1..7 (got %d)", who, terms);
    Y2_REQUIRE(kcap >= 0, "%s: kcap is the longest accumulation chain in k-blocks, 0 = unlimited (got %d)", who, kcap);
    Y2_REQUIRE(block_n == 0 || (block_n % 32 == 0 && block_n >= 32 && block_n <= 256), "%s: block_n %d invalid", who, block_n);
    Y2_REQUIRE(in_bound > 0.f && isfinite(in_bound), "%s: the activations' bound must be positive and finite", who);
    Y2_REQUIRE(!o16 || (o8 && or8 && cout % 32 == 0 && out_bound > 0.f && isfinite(out_bound)),
               "%s: split outputs need all three arrays, cout %% 32 == 0 and a positive finite output bound", who);
    const int taps = ksize * ksize;
    const size_t M = (size_t)B * H * W, K = (size_t)taps * cin;
    Y2_REQUIRE(M < ((size_t)1 << 31), "%s: too many pixels", who);
    int bn = block_n;
    if (bn == 0) {                                   // widest tile that divides the padded channel count evenly (as choose_tiles)
        const int p32 = (cout + 31) / 32 * 32, nt = (p32 + 255) / 256;
        bn = ((p32 + nt - 1) / nt + 31) / 32 * 32;
    }
    const int cout_pad = (cout + bn - 1) / bn * bn;
    const int kblocks = taps * (cin / MX_BK);
    const int nchunks = kcap > 0 ? (kblocks + kcap - 1) / kcap : 1;
    Y2_REQUIRE(y || (o16 && nchunks == 1), "%s: the float32 output may only be omitted with split outputs and a single accumulation chain", who);
    int dev = 0;
    Y2_CUDA(cudaGetDevice(&dev));
    const int num_sms = device_sm_count(dev);
    if (mx_load_entry_points()) return -1;

    unsigned int* amax_d = nullptr;
    __half* w16 = nullptr;
    uint8_t *w8 = nullptr, *rw8 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&amax_d, sizeof(unsigned int)) != cudaSuccess || cudaMalloc(&w16, K * cout_pad * 2) != cudaSuccess ||
            cudaMalloc(&w8, K * cout_pad) != cudaSuccess || cudaMalloc(&rw8, K * cout_pad) != cudaSuccess) { set_error("%s: cudaMalloc failed", who); break; }
        if (cudaMemsetAsync(amax_d, 0, sizeof(unsigned int), s) != cudaSuccess) { set_error("%s: memset failed", who); break; }
        mix_amax_kernel<<<num_sms * 4, 256, 0, s>>>(w_hwio, K * cout, amax_d);
        note_launch();
        float amax_w = 0.f;
        if (cudaMemcpyAsync(&amax_w, amax_d, sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("%s: amax pass failed: %s", who, cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (!(amax_w > 0.f) || !isfinite(amax_w)) { set_error("%s: weights are all zero or not finite", who); break; }
        const MixScales sc = mix_scales(in_bound, amax_w);        // E16 F16 == E8 F8 4096: one accumulator for all three products
        mix_prep_w_kernel<<<num_sms * 8, 256, 0, s>>>(w_hwio, taps, cin, cout, cout_pad, sc.F16, sc.F8, sc.rw, w16, w8, rw8);
        note_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("%s: operand preparation failed to launch", who); break; }

        CUtensorMap ma16, ma8, mra8, mw16, mrw8, mw8;
        if (mx_map_act(&ma16, const_cast<__half*>(x16), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, H, W, cin, ksize)) break;
        if (mx_map_act(&ma8, const_cast<uint8_t*>(x8), CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, H, W, cin, ksize)) break;
        if (mx_map_act(&mra8, const_cast<uint8_t*>(rx8), CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, H, W, cin, ksize)) break;
        if (mx_map_w(&mw16, w16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, K, cout_pad, bn)) break;
        if (mx_map_w(&mrw8, rw8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, K, cout_pad, bn)) break;
        if (mx_map_w(&mw8, w8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, K, cout_pad, bn)) break;

        MixParams p;
        memset(&p, 0, sizeof(p));
        p.M = (int)M; p.N = cout; p.Cin = cin; p.ksize = ksize; p.B = B; p.H = H; p.W = W;
        p.block_n = bn; p.m_tiles = (int)((M + MX_BLOCK_M - 1) / MX_BLOCK_M); p.n_tiles = cout_pad / bn;
        p.kblocks = kblocks; p.cout_pad = cout_pad;
        p.terms = terms; p.leaky = leaky; p.unscale = sc.unscale; p.nchunks = nchunks;
        p.scale = scale; p.bias = bias; p.out = y; p.ldc = cout;
        if (o16) {
            const MixScales so = mix_scales(out_bound, 1.0f);     // only the activation half is used
            p.o16 = o16; p.o8 = o8; p.or8 = or8; p.oE16 = so.E16; p.oE8 = so.E8; p.ora = so.ra; p.amax_out = amax_out;
        }
        const int stage_bytes = MX_A16 + 2 * MX_A8 + bn * MX_BK * 4;
        int stages = (MX_SMEM_LIMIT - 1024 - MX_BAR_BYTES) / stage_bytes;
        if (stages > 8) stages = 8;
        if (stages < 2) { set_error("%s: tile does not fit shared memory", who); break; }
        p.num_stages = stages;
        const int smem_bytes = stages * stage_bytes + 1024 + MX_BAR_BYTES;
        const long long tiles = (long long)p.m_tiles * p.n_tiles;
        const int grid = (int)(tiles < num_sms ? tiles : num_sms);
        if (cudaFuncSetAttribute(conv_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM_LIMIT) != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        g_mix_used = true;
        conv_mix_kernel<<<grid, MX_THREADS, smem_bytes, s>>>(ma16, ma8, mra8, mw16, mrw8, mw8, p);      // warm-up / result
        note_launch();
        cudaEventRecord(e0, s);
        conv_mix_kernel<<<grid, MX_THREADS, smem_bytes, s>>>(ma16, ma8, mra8, mw16, mrw8, mw8, p);      // timed (same output)
        note_launch();
        cudaEventRecord(e1, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("%s: kernel failed: %s", who, cudaGetErrorString(cudaGetLastError())); break; }
        cudaEventElapsedTime(&g_mix_last_ms, e0, e1);
        if (conv_mix_check_watchdog()) break;
        rc = 0;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(amax_d); cudaFree(w16); cudaFree(w8); cudaFree(rw8);
    return rc;
}

extern "C" {

float y2_debug_last_mix_ms(void) { return g_mix_last_ms; }

int y2_mix_split(const float* x, size_t n, float bound, void* x16, void* x8, void* rx8, void* stream) {
    Y2_REQUIRE(x && x16 && x8 && rx8 && n > 0, "y2_mix_split: null argument");
    Y2_REQUIRE(bound > 0.f && isfinite(bound), "y2_mix_split: the bound must be positive and finite");
    int dev = 0;
    Y2_CUDA(cudaGetDevice(&dev));
    const MixScales sc = mix_scales(bound, 1.0f);
    mix_prep_act_kernel<<<device_sm_count(dev) * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, sc.E16, sc.E8, sc.ra, static_cast<__half*>(x16),
                                                                                               static_cast<uint8_t*>(x8), static_cast<uint8_t*>(rx8));
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int y2_conv2d_mix_pre(const void* x16, const void* x8, const void* rx8, float in_bound, int B, int H, int W, int cin, const float* w_hwio,
                      int ksize, int cout, const float* scale, const float* bias, int leaky, float* y, void* o16, void* o8, void* or8,
                      float out_bound, uint32_t* amax_out, int terms, int kcap, int block_n, void* stream) {
    Y2_REQUIRE(x16 && x8 && rx8 && w_hwio, "y2_conv2d_mix_pre: null argument");
    return mix_run("y2_conv2d_mix_pre", static_cast<const __half*>(x16), static_cast<const uint8_t*>(x8), static_cast<const uint8_t*>(rx8), in_bound,
                   B, H, W, cin, w_hwio, ksize, cout, scale, bias, leaky, y, static_cast<__half*>(o16), static_cast<uint8_t*>(o8),
                   static_cast<uint8_t*>(or8), out_bound, amax_out, terms, kcap, block_n, static_cast<cudaStream_t>(stream));
}

int y2_conv2d_mix(const float* x, int B, int H, int W, int cin, const float* w_hwio, int ksize, int cout, const float* scale,
                  const float* bias, int leaky, float* y, int terms, int kcap, int block_n, void* stream) {
    Y2_REQUIRE(x && w_hwio && y, "y2_conv2d_mix: null argument");
    Y2_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "y2_conv2d_mix: bad shape B=%d H=%d W=%d cin=%d cout=%d", B, H, W, cin, cout);
    Y2_REQUIRE(ksize == 1 || ksize == 3, "y2_conv2d_mix: ksize must be 1 or 3 (got %d)", ksize);
    Y2_REQUIRE(cin % MX_BK == 0, "y2_conv2d_mix: cin must be a multiple of 64 (got %d)", cin);
    Y2_REQUIRE(terms >= 1 && terms <= 7, "y2_conv2d_mix: terms is a bit mask 1..7 (got %d)", terms);
    Y2_REQUIRE(kcap >= 0, "y2_conv2d_mix: kcap is the longest accumulation chain in k-blocks, 0 = unlimited (got %d)", kcap);
    Y2_REQUIRE(block_n == 0 || (block_n % 32 == 0 && block_n >= 32 && block_n <= 256), "y2_conv2d_mix: block_n %d invalid", block_n);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    Y2_CUDA(cudaGetDevice(&dev));
    const int num_sms = device_sm_count(dev);
    const size_t n = (size_t)B * H * W * cin;
    unsigned int* amax_d = nullptr;
    __half* x16 = nullptr;
    uint8_t *x8 = nullptr, *rx8 = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&amax_d, sizeof(unsigned int)) != cudaSuccess || cudaMalloc(&x16, n * 2) != cudaSuccess || cudaMalloc(&x8, n) != cudaSuccess ||
            cudaMalloc(&rx8, n) != cudaSuccess) { set_error("y2_conv2d_mix: cudaMalloc failed"); break; }
        if (cudaMemsetAsync(amax_d, 0, sizeof(unsigned int), s) != cudaSuccess) { set_error("y2_conv2d_mix: memset failed"); break; }
        mix_amax_kernel<<<num_sms * 4, 256, 0, s>>>(x, n, amax_d);
        note_launch();
        float amax_x = 0.f;
        if (cudaMemcpyAsync(&amax_x, amax_d, sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("y2_conv2d_mix: amax pass failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (!(amax_x > 0.f) || !isfinite(amax_x)) { set_error("y2_conv2d_mix: activations are all zero or not finite"); break; }
        if (y2_mix_split(x, n, amax_x, x16, x8, rx8, stream)) break;
        rc = mix_run("y2_conv2d_mix", x16, x8, rx8, amax_x, B, H, W, cin, w_hwio, ksize, cout, scale, bias, leaky, y, nullptr, nullptr, nullptr,
                     0.f, nullptr, terms, kcap, block_n, s);
    } while (0);
    cudaFree(amax_d); cudaFree(x16); cudaFree(x8); cudaFree(rx8);
    return rc;
}

}  // extern "C"
