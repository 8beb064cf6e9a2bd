// Weight-gradient GEMM on tcgen05 (training step, BASELINE config 3):
//   dW[tap][cin][cout] = sum over pixels p of  x_in[p + tap][cin] * dx[p][cout]
// i.e. what tf.gradients produces for slim.layers.conv2d weights (model/yolo2/inference.py:73-118,
// train.py:127).  GEMM view: M = (tap, cin) rows, N = cout, K = B*H*W pixels.
//
// Both operands are "MN-major" for the tensor core (the contraction runs over pixels, memory is
// channel-contiguous NHWC): the smem tiles are [64 pixels][64 channels] 128B-swizzled blocks, exactly what
// TMA delivers -- A through the same im2col-mode tensor map as the forward conv (tap shift + zero padding
// for free), B through a tiled map over the dx planes.  Operands are split bf16 planes, 3 MMAs per K-step
// (hi*hi + hi*lo + lo*hi) into one fp32 TMEM accumulator, like the forward kernel.  An M-tile is 128 rows =
// 2 channel atoms of 64 (or 4 atoms of 32 for conv1, whose taps are packed side by side).  Stream-K over the
// (tile, pixel-block) space balances the 148 SMs; partial tiles are handed over through L2 in fixed order.
#include <string.h>

#include <atomic>

#include "y2_internal.h"
#include "y2_ptx.cuh"

namespace y2 {

static constexpr int WG_M = 128;
static constexpr int WG_KPIX = 64;            // pixels per k-block
static constexpr int WG_THREADS = 256;
static constexpr int WG_EPI0 = 4;
static constexpr int WG_TMEM = 512;
static constexpr int WG_ACC = 256;
static constexpr int WG_SMEM = 227 * 1024;

struct WgradParams {
    int P, B, H, W, Cin, Cout, ksize;
    int atom_ch, apt, apc, total_atoms;       // channels per atom (64|32), atoms per tile, atoms per tap, taps*apc
    int m_tiles, n_tiles, block_n, kblocks_total, num_stages;
    int split_tiles, split_chunks;   // K-aligned split (few tiles, huge K): grid = tiles x chunks, CTA i -> (tile i % tiles, k-range i / tiles)
    int dp_tiles, sk_ctas;
    float* dw;                                // [taps][Cin][Cout]
    float* sk_partial;
    float* sk_run;                            // [grid][128][block_n] running sums of capped accumulation chains (CapIter)
    int kcap;                                 // longest tensor-core accumulation chain in k-blocks (0 = unlimited)
    int fmt;                                  // FMT_* bits: element formats of the x (A) and dx (B) planes
    unsigned int* sk_flags;
    unsigned int epoch;
};

struct WgradLaunch {
    CUtensorMap map_x, map_d;
    WgradParams p;
    int grid, smem_bytes;
};

__device__ __forceinline__ void wg_flag_set(unsigned int* f, unsigned int epoch) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__device__ __forceinline__ void wg_flag_wait(const unsigned int* f, unsigned int epoch) {
    const long long t0 = clock64();
    unsigned int v, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v == epoch) return;
        if ((++spins & 63u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(&g_watchdog.fired)) return;
            if (clock64() - t0 > 2000000000LL) { watchdog_fire(0x900u, epoch); return; }
        }
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_e(uint32_t leader, void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %6, 0;\n\t"
        "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(leader)
        : "memory");
}
// MN-major operand: atoms of (row_bytes/2) channels, K rows of row_bytes; LBO = atom stride, SBO = 8-row group.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t row_bytes, uint32_t atom_stride) {
    const uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(atom_stride >> 4) << 16;           // LBO
    d |= static_cast<uint64_t>((8u * row_bytes) >> 4) << 32;      // SBO
    d |= static_cast<uint64_t>(1) << 46;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int a_row = p.atom_ch * 2;                       // bytes per pixel row of an A atom (128 | 64)
    const int a_atom = WG_KPIX * a_row;                    // 8 KiB | 4 KiB
    const int a_plane = p.apt * a_atom;                    // 16 KiB
    const int b_atom = WG_KPIX * 128;                      // 8 KiB
    const int nb = p.block_n / 64;
    const int b_plane = nb * b_atom;
    const int stage_bytes = 2 * (a_plane + b_plane);
    const int S = p.num_stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp: uniform for the compiler

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_d); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, WG_TMEM);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int KB = p.kblocks_total;
    const long long sk_total = (long long)(p.m_tiles * p.n_tiles - p.dp_tiles) * KB;
    const int pad = p.ksize / 2;
    const int hw = p.H * p.W;

    if (warp == 0) {
        // TMA producer: whole warp converged (operands stay in uniform registers), the elected lane issues
        const uint32_t leader = elect_one() ? 1u : 0u;
        int stage = 0;
        uint32_t phase = 0;
        CapIter it;
        if (p.split_chunks) it.it.init_split(p.split_tiles, p.split_chunks, KB); else it.it.init(p.dp_tiles, p.sk_ctas, sk_total, KB);
        it.wrap(p.kcap);
        int tile, kb0, kb1, seg_a, seg_b;
        while (it.next(tile, kb0, kb1, seg_a, seg_b)) {
            const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
            const int atom0 = __shfl_sync(0xffffffffu, mt * p.apt, 0);
            const int ncol0 = __shfl_sync(0xffffffffu, nt * p.block_n, 0);
            const int valid_atoms = min(p.apt, p.total_atoms - atom0);
            const uint32_t tx = 2u * (uint32_t)(valid_atoms * a_atom + b_plane);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int p0 = kb * WG_KPIX;
                const int img = p0 / hw, rem = p0 - img * hw;
                const int y0 = rem / p.W, x0 = rem - y0 * p.W;
                mbar_wait(&empty[stage], phase ^ 1u, 0x600u + stage);
                __syncwarp();
                stage = __shfl_sync(0xffffffffu, stage, 0);
                uint8_t* st = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx_e(leader, &full[stage], tx);
                for (int a = 0; a < valid_atoms; ++a) {
                    const int ga = atom0 + a;
                    const int tap = ga / p.apc, c0 = (ga - tap * p.apc) * p.atom_ch;
                    const int dy = (p.ksize == 3) ? tap / 3 : 0, dx = (p.ksize == 3) ? tap - dy * 3 : 0;
                    tma_load_im2col_4d_e(leader, st + a * a_atom, &map_x, &full[stage], c0, x0 - pad, y0 - pad, img, (uint16_t)dx, (uint16_t)dy);
                    tma_load_im2col_4d_e(leader, st + a_plane + a * a_atom, &map_x, &full[stage], c0, x0 - pad, y0 - pad, img + p.B,
                                         (uint16_t)dx, (uint16_t)dy);
                }
                uint8_t* sb = st + 2 * a_plane;
                for (int j = 0; j < nb; ++j) {
                    tma_load_3d_e(leader, sb + j * b_atom, &map_d, &full[stage], ncol0 + j * 64, p0, 0);
                    tma_load_3d_e(leader, sb + b_plane + j * b_atom, &map_d, &full[stage], ncol0 + j * 64, p0, 1);
                }
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp converged, the elected lane issues.  D = f32, A = B = bf16, both MN-major (bits 15, 16)
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t mn = (1u << 15) | (1u << 16);
        const uint32_t idesc = make_idesc_16(WG_M, (uint32_t)p.block_n, p.fmt & FMT_A_HI, p.fmt & FMT_B_HI) | mn;
        const uint32_t idesc_hl = make_idesc_16(WG_M, (uint32_t)p.block_n, p.fmt & FMT_A_HI, p.fmt & FMT_B_LO) | mn;
        const uint32_t idesc_lh = make_idesc_16(WG_M, (uint32_t)p.block_n, p.fmt & FMT_A_LO, p.fmt & FMT_B_HI) | mn;
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t ha = (uint32_t)(make_mnmajor_desc(0, a_row, a_atom) >> 32);      // high words: constants
        const uint32_t hb = (uint32_t)(make_mnmajor_desc(0, 128, b_atom) >> 32);
        const uint32_t la = (uint32_t)make_mnmajor_desc(0, a_row, a_atom);               // LBO field of the low words
        const uint32_t lb = (uint32_t)make_mnmajor_desc(0, 128, b_atom);
        const uint32_t a_kstep = (uint32_t)(16 * a_row) >> 4, b_kstep = (16u * 128u) >> 4;
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        CapIter it;
        if (p.split_chunks) it.it.init_split(p.split_tiles, p.split_chunks, KB); else it.it.init(p.dp_tiles, p.sk_ctas, sk_total, KB);
        it.wrap(p.kcap);
        int tile, kb0, kb1, seg_a, seg_b;
        while (it.next(tile, kb0, kb1, seg_a, seg_b)) {              // every sub-segment: a fresh accumulator buffer
            mbar_wait(&tempty[acc], acc_phase ^ 1u, 0x700u + acc);
            __syncwarp();
            tc_fence_after();
            const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(acc * WG_ACC), 0);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase, 0x800u + stage);
                __syncwarp();
                tc_fence_after();
                stage = __shfl_sync(0xffffffffu, stage, 0);
                const uint32_t a_hi = smem_base + (uint32_t)(stage * stage_bytes);
                const uint32_t da_hi = la + (a_hi >> 4), da_lo = la + ((a_hi + (uint32_t)a_plane) >> 4);
                const uint32_t db_hi = lb + ((a_hi + 2u * (uint32_t)a_plane) >> 4);
                const uint32_t db_lo = lb + ((a_hi + 2u * (uint32_t)a_plane + (uint32_t)b_plane) >> 4);
#pragma unroll
                for (int k = 0; k < WG_KPIX / 16; ++k) {
                    const uint32_t ak = (uint32_t)k * a_kstep, bk = (uint32_t)k * b_kstep;
                    tc_mma_f16_e(leader, d_tmem, da_hi + ak, ha, db_hi + bk, hb, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    tc_mma_f16_e(leader, d_tmem, da_hi + ak, ha, db_lo + bk, hb, idesc_hl, 1u);
                    tc_mma_f16_e(leader, d_tmem, da_lo + ak, ha, db_hi + bk, hb, idesc_lh, 1u);
                }
                tc_commit_e(leader, &empty[stage]);
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            tc_commit_e(leader, &tfull[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp >= WG_EPI0) {
        const int q = warp - WG_EPI0;
        const int et = threadIdx.x - WG_EPI0 * 32;
        const int r = q * 32 + lane;                        // row inside the M-tile
        int acc = 0;
        uint32_t acc_phase = 0;
        float* my_partial = p.sk_partial + (size_t)blockIdx.x * WG_M * p.block_n + (size_t)r * p.block_n;
        // running sum of this CTA's accumulation chain (see CapIter): private, [32-column chunk][warp][16-byte piece][lane]
        float4* run4 = reinterpret_cast<float4*>(p.sk_run + (size_t)blockIdx.x * WG_M * p.block_n) + (q * 8 * 32 + lane);
        CapIter it;
        if (p.split_chunks) it.it.init_split(p.split_tiles, p.split_chunks, KB); else it.it.init(p.dp_tiles, p.sk_ctas, sk_total, KB);
        it.wrap(p.kcap);
        int tile, kb0, kb1, seg_a, seg_b;
        while (it.next(tile, kb0, kb1, seg_a, seg_b)) {
            const bool first_sub = (kb0 == seg_a), last_sub = (kb1 == seg_b);
            const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
            const int n0 = nt * p.block_n;
            const int ga = mt * p.apt + r / p.atom_ch;
            const bool row_ok = ga < p.total_atoms;
            const int tap = row_ok ? ga / p.apc : 0;
            const int cch = row_ok ? (ga - tap * p.apc) * p.atom_ch + (r % p.atom_ch) : 0;
            float* drow = p.dw + ((size_t)tap * p.Cin + cch) * p.Cout;
            const bool is_head = (seg_a == 0) && last_sub;
            // the other CTAs holding k-ranges of this tile: ids cfirst + j * cstride, j < ncontrib
            int cfirst = blockIdx.x + 1, cstride = 1, ncontrib = 0;
            if (is_head && seg_b < KB) {
                if (p.split_chunks) {
                    cfirst = blockIdx.x + p.split_tiles; cstride = p.split_tiles; ncontrib = p.split_chunks - 1;
                } else {
                    int last_contrib = blockIdx.x;
                    const long long tile_end = (long long)(tile - p.dp_tiles + 1) * KB;
                    while (last_contrib + 1 < p.sk_ctas && sk_total * (last_contrib + 1) / p.sk_ctas < tile_end) ++last_contrib;
                    ncontrib = last_contrib - (int)blockIdx.x;
                }
            }
            mbar_wait(&tfull[acc], acc_phase, 0xA00u + acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * WG_ACC);
            if (!last_sub) {
                // not the end of the chain: add this sub-result to the running sum (round-to-nearest fp32) and release the
                // accumulator; the MMAs of the next sub-segment are already filling the other buffer
                for (int c = 0; c < p.block_n; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + (uint32_t)c, v);
                    tmem_ld_wait();
                    float4* own = run4 + (size_t)(c >> 5) * 4 * 8 * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 t = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                               __uint_as_float(v[4 * j + 3]));
                        if (!first_sub) {
                            const float4 o = __ldcg(own + j * 32);
                            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
                        }
                        __stcg(own + j * 32, t);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
                continue;
            }
            if (p.split_chunks) {
                // K-aligned split: cooperative hand-off.  EVERY CTA of the tile publishes its raw partial, waits for the
                // others, then reduces its own 1/chunks slice of the tile over all partials in chunk order (fixed order =
                // deterministic) and writes that slice of dW.  The adds of a tile are spread over all its CTAs instead of
                // serialised on one (conv3's single 128x64 tile has 148 partials).
                for (int c = 0; c < p.block_n; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + (uint32_t)c, v);
                    tmem_ld_wait();
                    float4* dst = reinterpret_cast<float4*>(my_partial + c);
                    const float4* own = run4 + (size_t)(c >> 5) * 4 * 8 * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 t = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                               __uint_as_float(v[4 * j + 3]));
                        if (!first_sub) {                  // earlier sub-segments of this chain
                            const float4 o = __ldcg(own + j * 32);
                            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
                        }
                        __stcg(dst + j, t);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) wg_flag_set(p.sk_flags + blockIdx.x, p.epoch);
                const int tile_id = (int)blockIdx.x % p.split_tiles, my_chunk = (int)blockIdx.x / p.split_tiles;
                for (int j = et; j < p.split_chunks; j += 128)
                    if (j != my_chunk) wg_flag_wait(p.sk_flags + tile_id + j * p.split_tiles, p.epoch);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int bn4 = p.block_n / 4;
                const int e4 = WG_M * bn4;                                 // float4 elements of the tile
                const int lo = (int)((long long)e4 * my_chunk / p.split_chunks), hi = (int)((long long)e4 * (my_chunk + 1) / p.split_chunks);
                const size_t slot = (size_t)WG_M * p.block_n;
                const float* base = p.sk_partial + (size_t)tile_id * slot;
                const size_t jstride = (size_t)p.split_tiles * slot;
                for (int e = lo + et; e < hi; e += 128) {
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                    int j = 0;
                    for (; j + 8 <= p.split_chunks; j += 8) {
                        float4 t[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) t[u] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)(j + u) * jstride) + e);
#pragma unroll
                        for (int u = 0; u < 8; ++u) { sum.x += t[u].x; sum.y += t[u].y; sum.z += t[u].z; sum.w += t[u].w; }
                    }
                    for (; j < p.split_chunks; ++j) {
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(base + (size_t)j * jstride) + e);
                        sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
                    }
                    const int rr = e / bn4, col = (e - rr * bn4) * 4;
                    const int ga2 = mt * p.apt + rr / p.atom_ch;
                    if (ga2 >= p.total_atoms) continue;
                    const int tap2 = ga2 / p.apc;
                    const int cch2 = (ga2 - tap2 * p.apc) * p.atom_ch + (rr % p.atom_ch);
                    float* d = p.dw + ((size_t)tap2 * p.Cin + cch2) * p.Cout + n0 + col;
                    if ((p.Cout & 3) == 0 && n0 + col + 4 <= p.Cout) {
                        *reinterpret_cast<float4*>(d) = sum;
                    } else {
                        const float sv[4] = {sum.x, sum.y, sum.z, sum.w};
                        for (int u = 0; u < 4; ++u) if (n0 + col + u < p.Cout) d[u] = sv[u];
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
                continue;
            }
            for (int j = 0; j < ncontrib; ++j) {
                if (lane == 0) wg_flag_wait(p.sk_flags + cfirst + j * cstride, p.epoch);
                __syncwarp();
            }
            // Other contributors' partials: staged through the idle smem ring with cp.async in batches of `cap`
            // contributors per 32-column chunk, one (chunk, batch) step ahead of the adds (per-thread row slots, 16-byte
            // pieces XOR-swizzled by row).  A 1x1 layer at 104x104 has ONE output tile and 147 contributors: every
            // batch keeps cap x 16 KiB in flight instead of one exposed L2 round trip per contributor.
            const int cap = (int)(((size_t)S * stage_bytes) / (2u * WG_M * 128u));
            const int nbatch = (ncontrib + cap - 1) / cap;
            auto stage_slot = [&](int buf, int hh) -> uint32_t {
                return smem_u32(smem) + (uint32_t)(((buf * cap + hh) * WG_M + r) * 128);
            };
            auto stage_issue = [&](int c, int bt, int buf) {
                const int h0 = bt * cap, cnt = min(cap, ncontrib - h0);
                for (int hh = 0; hh < cnt; ++hh) {
                    const float* src = p.sk_partial + (size_t)(cfirst + (h0 + hh) * cstride) * WG_M * p.block_n + (size_t)r * p.block_n + c;
                    const uint32_t dst = stage_slot(buf, hh);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)((j ^ (r & 7)) << 4)), "l"(src + 4 * j) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            int step = 0;
            if (is_head && nbatch > 0) stage_issue(0, 0, 0);
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_row + (uint32_t)c, v);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (!first_sub) {                          // earlier sub-segments of this chain
                    const float4* own = run4 + (size_t)(c >> 5) * 4 * 8 * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 o = __ldcg(own + j * 32);
                        f[4 * j] += o.x; f[4 * j + 1] += o.y; f[4 * j + 2] += o.z; f[4 * j + 3] += o.w;
                    }
                }
                if (!is_head) {
                    float4* dst = reinterpret_cast<float4*>(my_partial + c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) __stcg(dst + j, make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    continue;
                }
                for (int bt = 0; bt < nbatch; ++bt, ++step) {           // ascending contributor order: deterministic sum
                    const int buf = step & 1;
                    const bool more = (bt + 1 < nbatch) || (c + 32 < p.block_n);
                    if (more) {
                        if (bt + 1 < nbatch) stage_issue(c, bt + 1, buf ^ 1); else stage_issue(c + 32, 0, buf ^ 1);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                    const int cnt = min(cap, ncontrib - bt * cap);
                    for (int hh = 0; hh < cnt; ++hh) {
                        const uint32_t src = stage_slot(buf, hh);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 t;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                         : "r"(src + (uint32_t)((j ^ (r & 7)) << 4)));
                            f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
                        }
                    }
                }
                if (!row_ok) continue;
                if ((p.Cout & 3) == 0 && n0 + c + 32 <= p.Cout) {
                    float4* dst = reinterpret_cast<float4*>(drow + n0 + c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + c + j < p.Cout) drow[n0 + c + j] = f[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (!is_head) {
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) wg_flag_set(p.sk_flags + blockIdx.x, p.epoch);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, WG_TMEM);
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled wg_encodeTiled = nullptr;
static PFN_encodeIm2col wg_encodeIm2col = nullptr;
static std::atomic<unsigned int> wg_epoch{0};

static int wg_load_entry_points() {
    if (wg_encodeTiled && wg_encodeIm2col) return 0;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    Y2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    Y2_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    wg_encodeTiled = reinterpret_cast<PFN_encodeTiled>(fn);
    fn = nullptr;
    Y2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    Y2_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col unavailable");
    wg_encodeIm2col = reinterpret_cast<PFN_encodeIm2col>(fn);
    return 0;
}

// x_planes: bf16 [2][B][H][W][Cin] (input of the forward conv); dx_planes: bf16 [2][P][dpitch] (dpitch % 64 == 0,
// columns >= Cout are zero); dw: fp32 [k*k][Cin][Cout].  sk_ws as for the forward kernel (zeroed flag page).
int wgrad_tc_run(const bf16* x_planes, int B, int H, int W, int Cin, int ksize, const bf16* dx_planes, int Cout,
                 int dpitch, float* dw, int max_ctas, int num_sms, void* sk_ws, cudaStream_t stream) {
    if (wg_load_entry_points()) return -1;
    Y2_REQUIRE(ksize == 1 || ksize == 3, "wgrad: ksize must be 1 or 3");
    Y2_REQUIRE(Cin % 32 == 0 && dpitch % 64 == 0 && dpitch >= Cout, "wgrad: Cin %% 32, dpitch %% 64 required");
    WgradLaunch L;
    memset(&L, 0, sizeof(L));
    WgradParams& p = L.p;
    const long long P = (long long)B * H * W;
    Y2_REQUIRE(P < (1ll << 31), "wgrad: too many pixels");
    p.P = (int)P; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize;
    p.atom_ch = (Cin % 64 == 0) ? 64 : 32;
    p.apt = WG_M / p.atom_ch;
    p.apc = Cin / p.atom_ch;
    p.total_atoms = ksize * ksize * p.apc;
    p.m_tiles = (p.total_atoms + p.apt - 1) / p.apt;
    const int npad = dpitch;
    const int nt = (npad + 255) / 256;
    int bn = ((npad + nt - 1) / nt + 63) / 64 * 64;
    // few output tiles and a huge K (1x1 layers): narrower N tiles give more tiles, i.e. fewer stream-K contributors per
    // tile to add up in the hand-off (the MMAs of these layers are a few microseconds either way)
    while (ksize == 1 && bn > 64 && (long long)p.m_tiles * ((npad + bn - 1) / bn) < 16) bn = (bn / 2 + 63) / 64 * 64;
    p.block_n = bn;
    p.n_tiles = (npad + bn - 1) / bn;
    p.kblocks_total = (int)((P + WG_KPIX - 1) / WG_KPIX);
    p.dw = dw;
    p.sk_flags = static_cast<unsigned int*>(sk_ws);
    p.sk_partial = reinterpret_cast<float*>(static_cast<char*>(sk_ws) + 4096);
    p.sk_run = p.sk_partial + (size_t)num_sms * WG_M * 256;     // second half of tc_conv_streamk_bytes()
    p.kcap = g_conv_kcap;
    p.fmt = g_wgrad_fmt;
    const int stage_bytes = 2 * (WG_M * WG_KPIX * 2 + bn * WG_KPIX * 2);
    int stages = (WG_SMEM - 1024 - 256) / stage_bytes;
    if (stages > 8) stages = 8;
    Y2_REQUIRE(stages >= 2, "wgrad: tile does not fit shared memory");
    p.num_stages = stages;
    L.smem_bytes = stages * stage_bytes + 1024 + 256;
    choose_schedule((long long)p.m_tiles * p.n_tiles, p.kblocks_total, num_sms, max_ctas, bn / 256.0, 0, &p.dp_tiles, &p.sk_ctas, &L.grid);
    {
        // Few tiles, huge K (the early layers: 3..36 tiles, 10^4 k-blocks): K-aligned split instead of tile-major
        // stream-K.  Tile-major ranges make every tile stream dx (and its taps of x) from HBM on its own -- conv2's
        // wgrad moved 3.5 GB for 0.53 GB of operands; with all CTAs at the same K phase the re-reads hit L2.
        const int tiles = p.m_tiles * p.n_tiles;
        if (max_ctas == 0 && g_sched_override == 0 && tiles * 2 <= num_sms && p.kblocks_total >= 8 * (num_sms / tiles)) {
            p.split_tiles = tiles;
            p.split_chunks = num_sms / tiles;
            p.dp_tiles = 0; p.sk_ctas = 0;
            L.grid = tiles * p.split_chunks;
        }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(2 * B)};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const int padv = ksize / 2;
        int lower[2] = {-padv, -padv};
        int upper[2] = {padv - (ksize - 1), padv - (ksize - 1)};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = wg_encodeIm2col(&L.map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(x_planes), dims, strides,
                                     lower, upper, (cuuint32_t)p.atom_ch, (cuuint32_t)WG_KPIX, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     p.atom_ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeIm2col failed (%d)", (int)r);
        int drv = 0;
        cudaDriverGetVersion(&drv);
        if (drv <= 13010 && (size_t)2 * B * H * W * Cin * 2 < 131072) reinterpret_cast<uint64_t*>(&L.map_x)[1] &= ~(1ull << 21);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)dpitch, (cuuint64_t)P, 2};
        cuuint64_t strides[2] = {(cuuint64_t)dpitch * 2, (cuuint64_t)P * dpitch * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)WG_KPIX, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = wg_encodeTiled(&L.map_d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(dx_planes), dims, strides, box,
                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    static unsigned long long attr_seen = 0;
    if (first_use_on_current_device(attr_seen))
        Y2_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    unsigned int e = wg_epoch.fetch_add(1) + 0x40000001u;     // disjoint from the forward kernel's epochs for a long time
    p.epoch = e;
    wgrad_tc_kernel<<<L.grid, WG_THREADS, L.smem_bytes, stream>>>(L.map_x, L.map_d, L.p);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int wgrad_check_watchdog() {
    Watchdog w;
    Y2_CUDA(cudaMemcpyFromSymbol(&w, g_watchdog, sizeof(w)));
    if (!w.fired) return 0;
    Watchdog z;
    memset(&z, 0, sizeof(z));
    cudaMemcpyToSymbol(g_watchdog, &z, sizeof(z));
    set_error("tcgen05 wgrad: barrier watchdog fired (block %u warp %u wait-site 0x%x): pipeline deadlock", w.block, w.warp, w.tag);
    return -3;
}

}  // namespace y2
