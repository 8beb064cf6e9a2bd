// conv0 (model/yolo2/inference.py:73, first loop iteration; tiny: :35) on the tensor cores: 3x3, Cin = 3 -> 32 stored output
// channels, SAME, + BN (scale/bias) + leaky + 2x2/2 max-pool (:74), fused.
//
// K = 27 is too small and too oddly shaped for TMA (3 channels = 6 bytes per pixel), which is why the first version of this
// layer ran on the CUDA cores (conv0_pool_kernel: 0.22 ms per batch of 32 = 7 % of the inference step at 58 % of the fp32
// FMA rate).  Here the im2col tile is BUILT by ordinary threads: a tile is a 16 x 8 pixel block of one image (128 GEMM rows);
// builder thread t gathers the 3 x 3 x 3 patch of its pixel (27 fp32 values, zero outside the image = SAME padding), splits
// every value into bf16 hi + lo and writes one 64-byte row per plane (K padded 27 -> 32) straight into the 64B-swizzled
// K-major layout tcgen05.mma reads.  One MMA warp then issues 2 k-steps x (hi*hi + hi*lo + lo*hi) with N = 32 into a
// 32-column fp32 TMEM accumulator (double-buffered), and four epilogue warps apply BN, pool the 2x2 windows with the same
// exchange-and-halve shuffles as the fused-pool epilogue of conv_tc_kernel (window partners are lanes r^1 and r^16),
// apply leaky, split and store the pooled bf16 planes conv1's TMA reads.  The weights (2 KiB per plane) sit in shared
// memory for the life of the CTA.  Pipeline: builders -> full[stage] (128 arrivals, fence.proxy.async before each) -> MMA
// warp -> tcgen05.commit -> empty[stage] / tfull[acc] -> epilogue -> tempty[acc].  Every wait sits behind the watchdog.
#include <string.h>

#include "y2_internal.h"
#include "y2_ptx.cuh"

namespace y2 {

static constexpr int C0_THREADS = 288;            // warps 0-3 builders, warp 4 MMA issuer + TMEM owner, warps 5-8 epilogue
static constexpr int C0_STAGES = 4;
static constexpr int C0_A_PLANE = 128 * 64;       // 128 rows x 32 bf16
static constexpr int C0_STAGE = 2 * C0_A_PLANE;   // hi + lo
static constexpr int C0_B_PLANE = 32 * 64;        // 32 output channels x 32 bf16
static constexpr int C0_SMEM = C0_STAGES * C0_STAGE + 2 * C0_B_PLANE + 1024;
static constexpr int C0_TMEM_COLS = 64;           // 2 accumulators x 32 columns

__device__ __forceinline__ void c0_split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Waits of this kernel are short and frequent (a tile every ~1000 cycles per CTA) and five of the nine warps do little else:
// measured with ncu, their mbarrier polling executed as many instructions as the real work and took the issue slots from
// the builders.  Poll with a sleep in between (the 4-stage ring and the double-buffered accumulator absorb the wake-up lag).
__device__ __forceinline__ void c0_wait(uint64_t* bar, uint32_t parity, uint32_t tag) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(100);
        if ((++spins & 255u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(&g_watchdog.fired)) return;
            if (clock64() - t0 > 2000000000LL) {
                watchdog_fire(tag, parity);
                return;
            }
        }
    }
}
__device__ __forceinline__ void c0_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// FAST: tiles that do not touch the image border (all but the outer ring) gather their patch as 3 rows x 9 contiguous floats
// from one base pointer with immediate offsets -- no per-element bounds checks / address selects, which ncu showed to be
// ~40 % of the kernel's instructions (profiles/ncu_r1h_conv0_tc.txt: 127 M warp instructions, as many as the CUDA-core
// kernel); border tiles take the checked path.
template <bool FAST>
__global__ void __launch_bounds__(C0_THREADS, 2)
conv0_tc_pool_kernel(const float* __restrict__ x, const float* __restrict__ w_hwio, const float* __restrict__ scale,
                     const float* __restrict__ bias, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int B, int H, int W) {
    extern __shared__ uint8_t c0_smem_raw[];
    uint8_t* smem = c0_smem_raw + ((1024u - (smem_u32(c0_smem_raw) & 1023u)) & 1023u);
    uint8_t* bsm = smem + C0_STAGES * C0_STAGE;                 // weights: [hi | lo] x 32 rows x 64 B, swizzled like the A rows
    __shared__ uint64_t full[C0_STAGES], empty[C0_STAGES], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float ssc[32], sbi[32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = W / 16, tiles_y = H / 8;
    const int ntiles = tiles_x * tiles_y * B;

    if (threadIdx.x < 32) {
        // weight row n = output channel: k = (ky*3 + kx)*3 + c, HWIO index k*32 + n; K padded to 32 with zeros
        const int n = threadIdx.x;
        ssc[n] = scale[n];
        sbi[n] = bias[n];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = 8 * j + 2 * e;
                const float v0 = k0 < 27 ? __ldg(w_hwio + k0 * 32 + n) : 0.f;
                const float v1 = k0 + 1 < 27 ? __ldg(w_hwio + (k0 + 1) * 32 + n) : 0.f;
                c0_split_pack2(v0, v1, hi[e], lo[e]);
            }
            const uint32_t a = smem_u32(bsm) + (uint32_t)(n * 64 + ((j ^ ((n >> 1) & 3)) << 4));
            c0_st_shared_v4(a, hi[0], hi[1], hi[2], hi[3]);
            c0_st_shared_v4(a + C0_B_PLANE, lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();                               // generic-proxy writes -> visible to the tensor core's reads
    }
    if (warp == 4) {
        if (lane == 0) {
            for (int i = 0; i < C0_STAGES; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, C0_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp < 4) {
        // ===================== builders: one GEMM row (pixel) per thread =====================
        const int t = threadIdx.x;
        const int xx = t & 15, yy = t >> 4;
        const uint32_t row_off = (uint32_t)(t * 64), sw = (uint32_t)((t >> 1) & 3);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int xt = tile % tiles_x;
            const int r2 = tile / tiles_x;
            const int yt = r2 % tiles_y, b = r2 / tiles_y;
            const int px = xt * 16 + xx, py = yt * 8 + yy;
            float v[32];
            if (FAST && xt > 0 && xt < tiles_x - 1 && yt > 0 && yt < tiles_y - 1) {      // (uniform over the CTA)
                // k = (dy*3 + dx)*3 + c = dy*9 + i: the 9 floats of a patch row are contiguous in NHWC
                const float* p0 = x + (((size_t)b * H + (py - 1)) * W + (px - 1)) * 3;
                const size_t rs = (size_t)W * 3;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int i = 0; i < 9; ++i) v[dy * 9 + i] = __ldg(p0 + dy * rs + i);
            } else {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int iy = py + dy - 1;
                const bool rok = iy >= 0 && iy < H;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int ix = px + dx - 1;
                    const bool ok = rok && ix >= 0 && ix < W;
                    const float* src = x + (((size_t)b * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[(dy * 3 + dx) * 3 + c] = ok ? __ldg(src + c) : 0.f;
                }
            }
            }
#pragma unroll
            for (int k = 27; k < 32; ++k) v[k] = 0.f;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) c0_split_pack2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
            c0_wait(&empty[stage], phase ^ 1u, 0xC00u + stage);
            const uint32_t base = smem_u32(smem) + (uint32_t)(stage * C0_STAGE) + row_off;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a = base + ((j ^ sw) << 4);
                c0_st_shared_v4(a, hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                c0_st_shared_v4(a + C0_A_PLANE, lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&full[stage]);
            if (++stage == C0_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc = make_idesc_bf16(128, 32);
        const uint32_t hd = (uint32_t)(make_kmajor_desc(0, 64) >> 32);
        const uint32_t db_hi = (uint32_t)make_kmajor_desc(smem_u32(bsm), 64);
        const uint32_t db_lo = (uint32_t)make_kmajor_desc(smem_u32(bsm) + C0_B_PLANE, 64);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            c0_wait(&tempty[acc], acc_phase ^ 1u, 0xC10u + acc);
            c0_wait(&full[stage], phase, 0xC20u + stage);
            __syncwarp();
            tc_fence_after();
            stage = __shfl_sync(0xffffffffu, stage, 0);
            const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(acc * 32), 0);
            const uint32_t st = smem_u32(smem) + (uint32_t)(stage * C0_STAGE);
            const uint32_t da_hi = (uint32_t)make_kmajor_desc(st, 64);
            const uint32_t da_lo = (uint32_t)make_kmajor_desc(st + C0_A_PLANE, 64);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t koff = (uint32_t)(k * 2);          // 16 bf16 = 32 bytes along K (>> 4)
                tc_mma_f16_e(leader, d_tmem, da_hi + koff, hd, db_hi + koff, hd, idesc, k > 0 ? 1u : 0u);
                tc_mma_f16_e(leader, d_tmem, da_hi + koff, hd, db_lo + koff, hd, idesc, 1u);
                tc_mma_f16_e(leader, d_tmem, da_lo + koff, hd, db_hi + koff, hd, idesc, 1u);
            }
            tc_commit_e(leader, &empty[stage]);
            tc_commit_e(leader, &tfull[acc]);
            if (++stage == C0_STAGES) { stage = 0; phase ^= 1u; }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue: BN -> 2x2 max-pool (shuffles) -> leaky -> hi/lo planes =====================
        const int q = warp & 3;                                  // TMEM lane quarter of this warp
        const int r = q * 32 + lane;                             // GEMM row = pixel (xx, yy) of the 16 x 8 block
        const int xx = r & 15, yy = r >> 4;
        const bool odd_x = (lane & 1) != 0, odd_y = (lane & 16) != 0;
        const int Hp = H / 2, Wp = W / 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int xt = tile % tiles_x;
            const int r2 = tile / tiles_x;
            const int yt = r2 % tiles_y, b = r2 / tiles_y;
            c0_wait(&tfull[acc], acc_phase, 0xC30u + acc);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 32), v);
            tmem_ld_wait_dep(v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);            // the accumulator is in registers: release it at once
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaf(__uint_as_float(v[j]), ssc[j], sbi[j]);
            float a[16], m[8];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float send = odd_x ? f[j] : f[j + 16];
                const float keep = odd_x ? f[j + 16] : f[j];
                a[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float send = odd_y ? a[j] : a[j + 8];
                const float keep = odd_y ? a[j + 8] : a[j];
                const float tmax = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
                m[j] = fmaxf(tmax, 0.1f * tmax);
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c0_split_pack2(m[2 * j], m[2 * j + 1], hi[j], lo[j]);
            const int ppx = xt * 8 + (xx >> 1), ppy = yt * 4 + (yy >> 1);
            const size_t off = (((size_t)b * Hp + ppy) * Wp + ppx) * 32 + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);
            *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C0_TMEM_COLS);
    }
}

bool conv0_tc_applicable(int H, int W) { return H % 8 == 0 && W % 16 == 0; }

int conv0_tc_pool_launch(const float* x, const float* w_hwio, const float* scale, const float* bias, bf16* out_hi, bf16* out_lo, int B,
                         int H, int W, int num_sms, cudaStream_t s, int fast) {
    Y2_REQUIRE(conv0_tc_applicable(H, W), "conv0 (tensor cores): H %% 8 and W %% 16 must be 0");
    static unsigned long long attr_seen = 0;
    if (first_use_on_current_device(attr_seen)) {
        Y2_CUDA(cudaFuncSetAttribute(conv0_tc_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C0_SMEM));
        Y2_CUDA(cudaFuncSetAttribute(conv0_tc_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C0_SMEM));
    }
    const long long ntiles = (long long)(W / 16) * (H / 8) * B;
    Y2_REQUIRE(ntiles < (1ll << 31), "conv0 (tensor cores): too many tiles");
    long long grid = 2LL * num_sms;
    if (grid > ntiles) grid = ntiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(C0_THREADS); cfg.dynamicSmemBytes = C0_SMEM; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_conv_pdl ? 1 : 0;
    if (fast) Y2_CUDA(cudaLaunchKernelEx(&cfg, conv0_tc_pool_kernel<true>, x, w_hwio, scale, bias, out_hi, out_lo, B, H, W));
    else Y2_CUDA(cudaLaunchKernelEx(&cfg, conv0_tc_pool_kernel<false>, x, w_hwio, scale, bias, out_hi, out_lo, B, H, W));
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int conv0_tc_check_watchdog() {
    Watchdog w;
    Y2_CUDA(cudaMemcpyFromSymbol(&w, g_watchdog, sizeof(w)));
    if (!w.fired) return 0;
    Watchdog z;
    memset(&z, 0, sizeof(z));
    cudaMemcpyToSymbol(g_watchdog, &z, sizeof(z));
    set_error("tcgen05 conv0: barrier watchdog fired (block %u warp %u wait-site 0x%x parity %u): pipeline deadlock", w.block, w.warp,
              w.tag, w.parity);
    return -3;
}

}  // namespace y2
