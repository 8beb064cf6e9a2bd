// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05) for sm_100a.
//
// Replaces the slim.layers.conv2d + slim.batch_norm + leaky_relu triple that the reference
// stacks 21 times (model/yolo2/inference.py:62-69,73-117) and the final linear 1x1 conv (:118).
//
// GEMM view: D[M = B*H*W pixels][N = Cout] = A[M][K = taps*Cin] * W[N][K]^T, stride 1, SAME.
//   * A is never materialised: each (tap, 64-channel) K-block of an M-tile is fetched by ONE
//     im2col-mode TMA per plane straight from the NHWC activation (zero fill = SAME padding),
//     landing in the 128B-swizzled K-major layout tcgen05.mma reads.
//   * W tiles come from a pre-packed [Cout][tap][Cin] matrix via tiled TMA.
//   * fp32 parity: operands are split bf16 planes (hi, lo); per K-step the MMA warp issues
//     hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator (lo*lo ~ 2^-32 is dropped).
//   * Persistent CTAs, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (1 thread),
//     warp2 = TMEM allocator, warps4-7 = epilogue (TMEM -> regs -> BN scale/bias + leaky ->
//     bf16 hi/lo split -> global).  Two 256-column TMEM accumulators double-buffer the
//     epilogue against the next tile's MMAs.
//   * Stream-K scheduling: the (tile, k-block) iteration space of a layer is cut into gridDim.x equal
//     contiguous ranges, so every SM gets the same number of MMAs whatever the tile count (the 13x13
//     layers at batch 32 have 172 tiles for 148 SMs).  A range that starts inside a tile produces a
//     raw fp32 partial (one 128 x block_n slot per CTA, L2-resident) and raises a flag; the CTA that
//     owns the tile's first k-block runs last in time, adds the partials in fixed order
//     (deterministic) and applies the epilogue.  No second kernel, no atomics.
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "y2_internal.h"
#include "y2_ptx.cuh"

namespace y2 {

static constexpr int BLOCK_M = 128;
static constexpr int NUM_THREADS = 256;
static constexpr int EPI_WARP0 = 4;
static constexpr int TMEM_COLS = 512;
static constexpr int ACC_COLS = 256;
static constexpr int SMEM_LIMIT = 227 * 1024;
static constexpr int SB_BYTES = 2 * 256 * 4;      // scale/bias staging for one tile
static constexpr int BAR_BYTES = 256;

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stream-K hand-off flags: one word per CTA, value = launch epoch (monotonic, never reset)
__device__ __forceinline__ void flag_set(unsigned int* f, unsigned int epoch) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__device__ __forceinline__ void flag_wait(const unsigned int* f, unsigned int epoch, uint32_t tag) {
    const long long t0 = clock64();
    unsigned int v, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v == epoch) return;
        if ((++spins & 63u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(&g_watchdog.fired)) return;
            if (clock64() - t0 > 2000000000LL) {
                watchdog_fire(tag, epoch);
                return;
            }
        }
    }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read(int pending) {     // pending in {0, 1}
    if (pending == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 1-D bulk copies shared -> global: plain store, and fp32 add performed by the L2 (no read back to the SM)
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32_1d(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
// the same with an L2 eviction-priority hint (the 19 MB of running-sum slots should outlive the operand stream in L2)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_store_1d_hint(void* gdst, uint32_t ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32_1d_hint(void* gdst, uint32_t ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f32 [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc), "r"(bytes), "l"(pol) : "memory");
}

// (x0, x1) -> packed bf16 pairs hi = bf16(x), lo = bf16(x - hi); one packed convert per pair of values.
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);              // .x (low half) = x0
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// PAIR: the kernel runs as clusters of two CTAs (the two SMs of a TPC) that share one tcgen05.mma.cta_group::2 stream:
// a pair owns a 256-row M tile (CTA rank r: rows r*128..), each CTA stages its own A rows and HALF of the B tile, rank 0
// issues M = 256 MMAs whose accumulator rows live in each CTA's own TMEM.  Per k-block a CTA's shared memory then sees
// 64 KiB of TMA writes + 96 KiB of MMA reads instead of 96 + 144 (the main loop was shared-memory-bandwidth bound,
// profiles/README.md).  `worker` (the pair) replaces the CTA in the tile schedule; partial slots and flags stay per CTA.
template <int BK, bool SPLIT3, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_o, const ConvParams p) {
    constexpr int ROW_BYTES = BK * 2;                       // 128 (SW128) or 64 (SW64)
    constexpr int A_TILE = BLOCK_M * ROW_BYTES;
    constexpr int PLANES = SPLIT3 ? 2 : 1;
    constexpr int NCTA = PAIR ? 2 : 1;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space: LDS/STS, not generic)
    uint8_t* smem0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem = smem0 + p.stg_bytes;                    // [store staging | resident weights | ring | scale/bias | barriers]

    const int b_rows = p.block_n / NCTA;                    // rows of the B (weight) tile this CTA stages
    const int b_tile = b_rows * ROW_BYTES;
    const int stage_bytes = (!PAIR && p.halo) ? PLANES * p.halo_plane_bytes : PLANES * (A_TILE + b_tile);
    const int S = p.num_stages;
    uint8_t* ring = smem + p.bres_bytes;                    // halo mode keeps the weights of all taps below the ring
    float* sb = reinterpret_cast<float*>(ring + (size_t)S * stage_bytes);       // [2][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)S * stage_bytes + SB_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    uint64_t* bres = bars + 2 * S + 5;                      // halo mode: resident weights landed

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform (also for the compiler)
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        if (p.tma_store) tma_prefetch_desc(&map_o);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4 * NCTA);               // the epilogue warps of both CTAs release rank 0's accumulator
        }
        mbar_init(bres, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS);
        else tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                            // the peer's barriers exist before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the
    // previous kernel in the stream; nothing below may start before that kernel's memory is visible.  The next
    // kernel may be scheduled onto SMs as soon as this grid's CTAs retire (no-ops without the launch attribute).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int KB = p.kblocks_total;
    const long long sk_total = (long long)(p.m_tiles * p.n_tiles - p.dp_tiles) * KB;   // stream-K part
    const int cblocks = p.Cin / BK;
    const int pad = p.ksize / 2;
    const int hw = p.H * p.W;

    if (warp == 0) {
        // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
        const uint32_t leader = elect_one() ? 1u : 0u;
        int stage = 0;
        uint32_t phase = 0;
        CapIter it;
        it.init(worker, nworkers, p.dp_tiles, p.sk_ctas, sk_total, KB, p.kcap);
        int tile, kb0, kb1, seg_a, seg_b;
        if (!PAIR && BK == 32 && p.halo && !(p.dbg_flags & 1)) {      // all 9 taps of the packed weights: once per CTA
            mbar_expect_tx_e(leader, bres, (uint32_t)(9 * PLANES * b_tile));
            for (int tap = 0; tap < 9; ++tap) {
                tma_load_2d_e(leader, smem + (size_t)(tap * PLANES) * b_tile, &map_w, bres, tap * BK, 0);
                if (SPLIT3) tma_load_2d_e(leader, smem + (size_t)(tap * PLANES + 1) * b_tile, &map_w, bres, tap * BK, p.cout_pad);
            }
        }
        while (it.next(tile, kb0, kb1, seg_a, seg_b)) {
            const int nt = tile / p.m_tiles;
            const int mt = (tile - nt * p.m_tiles) * NCTA + rank;     // this CTA's 128-row tile
            int img, y0, x0;
            if (p.tx) {                                      // spatial tile (tx x ty pixels x tb images)
                const int xt = mt % p.tiles_x, r2 = mt / p.tiles_x;
                x0 = xt * p.tx; y0 = (r2 % p.tiles_y) * p.ty; img = (r2 / p.tiles_y) * p.tb;
            } else {                                         // linear range of 128 pixels
                const int m0 = mt * BLOCK_M;
                img = m0 / hw;
                const int rem = m0 - img * hw;
                y0 = rem / p.W;
                x0 = rem - y0 * p.W;
            }
            // __shfl_sync(.., 0) marks a value warp-uniform for ptxas, so the TMA operands stay in uniform registers
            x0 = __shfl_sync(0xffffffffu, x0, 0); y0 = __shfl_sync(0xffffffffu, y0, 0); img = __shfl_sync(0xffffffffu, img, 0);
            const int n0 = __shfl_sync(0xffffffffu, nt * p.block_n + rank * b_rows, 0);   // pair: this CTA's half of the B tile
            int tap = kb0 / cblocks;
            int cb = kb0 - tap * cblocks;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1u, 0x100u + stage);
                __syncwarp();
                stage = __shfl_sync(0xffffffffu, stage, 0);
                tap = __shfl_sync(0xffffffffu, tap, 0); cb = __shfl_sync(0xffffffffu, cb, 0);
                uint8_t* st = ring + (size_t)stage * stage_bytes;
                if (p.dbg_flags & 1) {                       // diagnostics: no loads, the MMAs read whatever is there
                    if (rank == 0) mbar_arrive_e(leader, &full[stage]);
                } else if (PAIR) {
                    // both CTAs' bytes complete on rank 0's barrier, which expects them all
                    const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
                    const int c0 = cb * BK;
                    const int dy = (p.ksize == 3) ? tap / 3 : 0;
                    const int dx = (p.ksize == 3) ? tap - dy * 3 : 0;
                    if (rank == 0) mbar_expect_tx_e(leader, &full[stage], (uint32_t)(2 * stage_bytes));
                    if (p.tx) {
                        tma_load_4d_pair_e(leader, st, &map_a, fb, c0, x0 + dx - pad, y0 + dy - pad, img);
                        if (SPLIT3) tma_load_4d_pair_e(leader, st + A_TILE, &map_a, fb, c0, x0 + dx - pad, y0 + dy - pad, img + p.B);
                    } else {
                        tma_load_im2col_4d_pair_e(leader, st, &map_a, fb, c0, x0 - pad, y0 - pad, img, (uint16_t)dx, (uint16_t)dy);
                        if (SPLIT3)
                            tma_load_im2col_4d_pair_e(leader, st + A_TILE, &map_a, fb, c0, x0 - pad, y0 - pad, img + p.B,
                                                      (uint16_t)dx, (uint16_t)dy);
                    }
                    uint8_t* sbt = st + PLANES * A_TILE;
                    const int kcoord = tap * p.Cin + c0;
                    tma_load_2d_pair_e(leader, sbt, &map_w, fb, kcoord, n0);
                    if (SPLIT3) tma_load_2d_pair_e(leader, sbt + b_tile, &map_w, fb, kcoord, p.cout_pad + n0);
                } else if (BK == 32 && p.halo) {             // one halo fetch serves all 9 taps of this tile
                    mbar_expect_tx_e(leader, &full[stage], (uint32_t)(PLANES * p.halo_tx_bytes));
                    tma_load_4d_e(leader, st, &map_a, &full[stage], 0, x0 - 1, y0 - 1, img);
                    if (SPLIT3) tma_load_4d_e(leader, st + p.halo_plane_bytes, &map_a, &full[stage], 0, x0 - 1, y0 - 1, img + p.B);
                } else {
                    const int c0 = cb * BK;
                    const int dy = (p.ksize == 3) ? tap / 3 : 0;
                    const int dx = (p.ksize == 3) ? tap - dy * 3 : 0;
                    mbar_expect_tx_e(leader, &full[stage], (uint32_t)stage_bytes);
                    if (p.tx) {      // tile-mode box {BK, tx, ty, tb} shifted by the tap; out-of-image = zero fill = SAME padding
                        tma_load_4d_e(leader, st, &map_a, &full[stage], c0, x0 + dx - pad, y0 + dy - pad, img);
                        if (SPLIT3) tma_load_4d_e(leader, st + A_TILE, &map_a, &full[stage], c0, x0 + dx - pad, y0 + dy - pad, img + p.B);
                    } else {
                        tma_load_im2col_4d_e(leader, st, &map_a, &full[stage], c0, x0 - pad, y0 - pad, img, (uint16_t)dx, (uint16_t)dy);
                        if (SPLIT3)
                            tma_load_im2col_4d_e(leader, st + A_TILE, &map_a, &full[stage], c0, x0 - pad, y0 - pad, img + p.B,
                                                 (uint16_t)dx, (uint16_t)dy);
                    }
                    uint8_t* sbt = st + PLANES * A_TILE;
                    const int kcoord = tap * p.Cin + c0;
                    tma_load_2d_e(leader, sbt, &map_w, &full[stage], kcoord, n0);
                    if (SPLIT3) tma_load_2d_e(leader, sbt + b_tile, &map_w, &full[stage], kcoord, p.cout_pad + n0);
                }
                if (++cb == cblocks) { cb = 0; ++tap; }
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues; pair: rank 0 only) ==========
        const uint32_t leader = elect_one() ? 1u : 0u;
        // three products per k-step: hi*hi, hi*lo, lo*hi -- each with the element formats of ITS two planes
        const uint32_t idesc = make_idesc_16(BLOCK_M * NCTA, (uint32_t)p.block_n, p.fmt & FMT_A_HI, p.fmt & FMT_B_HI);
        const uint32_t idesc_hl = make_idesc_16(BLOCK_M * NCTA, (uint32_t)p.block_n, p.fmt & FMT_A_HI, p.fmt & FMT_B_LO);
        const uint32_t idesc_lh = make_idesc_16(BLOCK_M * NCTA, (uint32_t)p.block_n, p.fmt & FMT_A_LO, p.fmt & FMT_B_HI);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        bool bres_ready = false;
        const uint32_t ring_u32 = smem_u32(ring);
        const uint32_t bres_u32 = smem_u32(smem);
        CapIter it;
        it.init(worker, nworkers, p.dp_tiles, p.sk_ctas, sk_total, KB, p.kcap);
        int tile, kb0, kb1, seg_a, seg_b;
        while (it.next(tile, kb0, kb1, seg_a, seg_b)) {              // every sub-segment: a fresh accumulator buffer
            mbar_wait(&tempty[acc], acc_phase ^ 1u, 0x200u + acc);
            __syncwarp();
            tc_fence_after();
            // __shfl_sync(.., 0) marks a value warp-uniform for ptxas: descriptors are then built in uniform registers
            const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(acc * ACC_COLS), 0);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase, 0x300u + stage);
                __syncwarp();
                tc_fence_after();
                stage = __shfl_sync(0xffffffffu, stage, 0);
                const uint32_t st = ring_u32 + (uint32_t)(stage * stage_bytes);
                if (!PAIR && BK == 32 && p.halo) {
                    if (!bres_ready) {
                        if (!(p.dbg_flags & 1)) mbar_wait(bres, 0u, 0x310u);
                        __syncwarp();
                        tc_fence_after();
                        bres_ready = true;
                    }
                    // A operand of tap (dy, dx) = the 8 x 16 window of the 10 x 18 halo tile starting at (dx, dy):
                    // descriptor start shifted by whole 64-byte rows, 8-row groups one halo row (10 rows) apart.  The
                    // swizzle is a function of the absolute shared-memory address (what TMA wrote), so no base offset.
                    const uint32_t dh_hi = (uint32_t)make_kmajor_desc_ex(st, 64, 640, 0);
                    const uint32_t dh_lo = (uint32_t)make_kmajor_desc_ex(st + (uint32_t)p.halo_plane_bytes, 64, 640, 0);
                    const uint32_t dw_hi = (uint32_t)make_kmajor_desc(bres_u32, 64);
                    const uint32_t dw_lo = (uint32_t)make_kmajor_desc(bres_u32 + (uint32_t)b_tile, 64);
                    const uint32_t ha = (uint32_t)(make_kmajor_desc_ex(0, 64, 640, 0) >> 32);   // high words: constants
                    const uint32_t hb = (uint32_t)(make_kmajor_desc(0, 64) >> 32);
                    const uint32_t wstep = (uint32_t)(PLANES * b_tile) >> 4;
                    if (!(p.dbg_flags & 2)) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint32_t a_off = (uint32_t)(((tap / 3) * 10 + (tap % 3)) * 64 + k * 32) >> 4;
                                const uint32_t w_off = (uint32_t)tap * wstep + (uint32_t)(k * 2);
                                tc_mma_f16_e(leader, d_tmem, dh_hi + a_off, ha, dw_hi + w_off, hb, idesc, (tap > 0 || k > 0) ? 1u : 0u);
                                if (SPLIT3) {
                                    tc_mma_f16_e(leader, d_tmem, dh_hi + a_off, ha, dw_lo + w_off, hb, idesc_hl, 1u);
                                    tc_mma_f16_e(leader, d_tmem, dh_lo + a_off, ha, dw_hi + w_off, hb, idesc_lh, 1u);
                                }
                            }
                        }
                    }
                } else {
                    const uint32_t da_hi = (uint32_t)make_kmajor_desc(st, ROW_BYTES);
                    const uint32_t da_lo = (uint32_t)make_kmajor_desc(st + A_TILE, ROW_BYTES);
                    const uint32_t db_hi = (uint32_t)make_kmajor_desc(st + PLANES * A_TILE, ROW_BYTES);
                    const uint32_t db_lo = (uint32_t)make_kmajor_desc(st + PLANES * A_TILE + (uint32_t)b_tile, ROW_BYTES);
                    const uint32_t hd = (uint32_t)(make_kmajor_desc(0, ROW_BYTES) >> 32);       // high word: constant
                    if (!(p.dbg_flags & 2)) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint32_t koff = (uint32_t)(k * 2);   // 16 bf16 = 32 bytes along K inside the swizzle row (>> 4)
                            if (PAIR) {
                                tc_mma_f16_pair_e(leader, d_tmem, da_hi + koff, hd, db_hi + koff, hd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                                if (SPLIT3) {
                                    tc_mma_f16_pair_e(leader, d_tmem, da_hi + koff, hd, db_lo + koff, hd, idesc_hl, 1u);
                                    tc_mma_f16_pair_e(leader, d_tmem, da_lo + koff, hd, db_hi + koff, hd, idesc_lh, 1u);
                                }
                            } else {
                            tc_mma_f16_e(leader, d_tmem, da_hi + koff, hd, db_hi + koff, hd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            if (SPLIT3) {
                                tc_mma_f16_e(leader, d_tmem, da_hi + koff, hd, db_lo + koff, hd, idesc_hl, 1u);
                                tc_mma_f16_e(leader, d_tmem, da_lo + koff, hd, db_hi + koff, hd, idesc_lh, 1u);
                            }
                            }
                        }
                    }
                }
                if (PAIR) tc_commit_pair_e(leader, &empty[stage]);   // (both CTAs' slots)
                else tc_commit_e(leader, &empty[stage]);             // smem slot reusable once these MMAs retire
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (PAIR) tc_commit_pair_e(leader, &tfull[acc]);         // (both CTAs' epilogues)
            else tc_commit_e(leader, &tfull[acc]);                   // accumulator complete -> epilogue
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
        const int q = warp - EPI_WARP0;
        const int et = threadIdx.x - EPI_WARP0 * 32;       // 0..127
        int acc = 0, sb_nt = -1;
        uint32_t acc_phase = 0;
        uint32_t store_seq = 0;
        const uint32_t stg_base = smem_u32(smem0) + (uint32_t)(q * p.store_bufs * 4096);
        // Partial slots use the same [32-column chunk][warp][16-byte piece][lane] layout as the running sums below: the head
        // thread that collects a partial has the writer's (warp, lane), so every access of a warp is one contiguous 512 B run.
        auto part4 = [&](int slot) -> float4* {
            return reinterpret_cast<float4*>(p.sk_partial + (size_t)slot * BLOCK_M * p.block_n) + (q * 8 * 32 + lane);
        };
        float4* my_part4 = part4((int)blockIdx.x);
        // running sum of a segment's sub-segments: a slot of its own (the partial slot may still hold this CTA's contribution
        // to the previous tile, not yet collected by that tile's head).  Only this thread ever reads what it writes there, so
        // the layout is [32-column chunk][16-byte piece][row]: a warp's 32 rows form one contiguous 512 B run per access
        // instead of 32 lines 1 KiB apart (the main loop is sensitive to every extra L2 request).
        // Layout [32-column chunk][warp][16-byte piece][lane]: 4 KiB per (chunk, warp) = the image of one staging slab, so
        // where slabs exist a sub-result leaves the SM as ONE bulk copy per chunk and is ADDED by the L2
        // (cp.reduce.async.bulk .add.f32): no read-back until the chain's last sub-segment.
        float4* run4 = reinterpret_cast<float4*>(p.sk_run + (size_t)blockIdx.x * BLOCK_M * p.block_n) + (q * 8 * 32 + lane);
        const bool bulk_run = p.store_bufs > 0 && !(p.dbg_flags & 16);
        const bool run_hint = (p.dbg_flags & 32) == 0;       // evict_last on the running-sum slots (measured: -0.6 % of the step)
        const uint64_t run_pol = l2_policy_evict_last();
        CapIter it;
        it.init(worker, nworkers, p.dp_tiles, p.sk_ctas, sk_total, KB, p.kcap);
        const uint32_t tempty_r0 = PAIR ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u, tempty_r1 = PAIR ? mapa_u32(smem_u32(&tempty[1]), 0) : 0u;
        int tile, kb0, kb1, seg_a, seg_b;
        if (p.dbg && et == 0) { p.dbg[blockIdx.x * 4 + 0] = gtime_ns(); p.dbg[blockIdx.x * 4 + 1] = 0; p.dbg[blockIdx.x * 4 + 2] = 0; }
        while (it.next(tile, kb0, kb1, seg_a, seg_b)) {
            if (p.dbg && et == 0 && tile >= p.dp_tiles && p.dbg[blockIdx.x * 4 + 1] == 0) p.dbg[blockIdx.x * 4 + 1] = gtime_ns();
            const int nt = tile / p.m_tiles;
            const int mt = (tile - nt * p.m_tiles) * NCTA + rank;
            const int n0 = nt * p.block_n;
            long long row = (long long)mt * BLOCK_M + q * 32 + lane;
            size_t pool_row = 0;
            int reorg_sub = 0;                               // position of this pixel inside its 2x2 block: dy * 2 + dx
            if (p.tx) {                                      // spatial tile: row r = (image bb, yy, xx) inside the block
                const int r = q * 32 + lane;
                const int xt = mt % p.tiles_x, r2 = mt / p.tiles_x;
                const int xx = r % p.tx, yy = (r / p.tx) % p.ty, bb = r / (p.tx * p.ty);
                const int px = xt * p.tx + xx, py = (r2 % p.tiles_y) * p.ty + yy, pb = (r2 / p.tiles_y) * p.tb + bb;
                row = ((long long)pb * p.H + py) * p.W + px;
                pool_row = ((size_t)pb * (p.H / 2) + py / 2) * (p.W / 2) + px / 2;
                reorg_sub = (py & 1) * 2 + (px & 1);
            }
            const bool row_ok = row < p.M && mt < p.m_tiles128;      // (pair: the odd last 128-row tile has no partner rows)
            const bool first_sub = (kb0 == seg_a), last_sub = (kb1 == seg_b);   // position in this CTA's accumulation chain
            const bool is_head = (seg_a == 0) && last_sub;   // owns the tile's output (written after its last sub-segment)
            // stream-K workers worker+1 .. last_contrib start inside this tile and hold its other k-ranges
            int last_contrib = worker;
            if (is_head && seg_b < KB) {
                const long long tile_end = (long long)(tile - p.dp_tiles + 1) * KB;
                while (last_contrib + 1 < p.sk_ctas && sk_total * (last_contrib + 1) / p.sk_ctas < tile_end) ++last_contrib;
            }

            // stage this tile's scale/bias -- only when the N-tile changes (consecutive tiles share it: m runs fastest),
            // which keeps two barriers and a global-load round trip out of the per-tile critical path of the
            // small-K layers.  The first barrier makes sure the previous tile's readers are done.
            if (is_head && nt != sb_nt) {                   // uniform over the 128 epilogue threads
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = et; i < p.block_n; i += 128) {
                    const int n = n0 + i;
                    sb[i] = (p.scale && n < p.N) ? __ldg(p.scale + n) : 1.0f;
                    sb[256 + i] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.0f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                sb_nt = nt;
            }

            // The other contributors ran at the START of their ranges and are normally long done: their flags are acquired
            // while this CTA's own MMAs are still running, which keeps that round trip out of the exposed tail.
            for (int h = worker + 1; h <= last_contrib; ++h) {       // (pair: the contributor CTA of the same rank)
                if (lane == 0) flag_wait(p.sk_flags + h * NCTA + rank, p.epoch, 0x500u);
                __syncwarp();
            }
            mbar_wait(&tfull[acc], acc_phase, 0x400u + acc);
            tc_fence_after();
            if (p.dbg && et == 0 && last_contrib > worker) p.dbg[blockIdx.x * 4 + 2] = gtime_ns();   // own MMAs done
            if (bulk_run && !first_sub) {
                // order this sub-segment's slot traffic behind the previous one's (store before add, adds in chain order;
                // the last sub-segment reads the finished sum with ordinary loads)
                if (lane == 0) {
                    bulk_wait_all();
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                __syncwarp();
            }
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);
            // Partials of the other contributors are staged through the (now idle: a head segment with contributors
            // is always this CTA's last segment) smem ring with cp.async, one 32-column chunk ahead, so the adds never
            // wait on an L2 round trip per contributor.  Each thread copies and later reads only its own row slot
            // (128 B, 16-byte pieces XOR-swizzled by row against bank conflicts): no cross-thread sync needed.
            const int ncontrib = last_contrib - worker;
            const int max_staged = (int)(((size_t)S * stage_bytes) / (2u * BLOCK_M * 128u));
            const int nstaged = ncontrib < max_staged ? ncontrib : max_staged;
            const int rr = q * 32 + lane;
            auto stage_slot = [&](int buf, int hh) -> uint32_t {
                return smem_u32(ring) + (uint32_t)(((buf * nstaged + hh) * BLOCK_M + rr) * 128);
            };
            auto stage_issue = [&](int c, int buf) {
                for (int hh = 0; hh < nstaged; ++hh) {
                    const float4* src = part4((worker + 1 + hh) * NCTA + rank) + (size_t)(c >> 5) * 4 * 8 * 32;
                    const uint32_t dst = stage_slot(buf, hh);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)((j ^ (rr & 7)) << 4)), "l"(src + j * 32) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            if (nstaged > 0) stage_issue(0, 0);
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row, v);                    // chunk 0 in flight
            for (int c = 0; c < p.block_n; c += 32) {
                tmem_ld_wait_dep(v);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (c + 32 < p.block_n) tmem_ld_32x32b_x32(t_row + (uint32_t)(c + 32), v);   // next chunk streams in behind the math
                float4* own = run4 + (size_t)(c >> 5) * 4 * 8 * 32;
                if (!first_sub && (last_sub || !bulk_run) && !(p.dbg_flags & 4)) {   // running sum of the earlier sub-segments
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = __ldcg(own + j * 32);
                        f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
                    }
                }
                if (!last_sub) {                             // sub-result -> this CTA's private running-sum slot
                    if (bulk_run) {
                        const uint32_t slab = stg_base + (uint32_t)((store_seq % p.store_bufs) * 4096);
                        if (lane == 0) bulk_wait_read(p.store_bufs - 1);   // the copy that last read this slab is done
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(slab + (uint32_t)(j * 512 + lane * 16), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                         __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (run_hint) {
                                if (first_sub) bulk_store_1d_hint(own - lane, slab, 4096, run_pol);
                                else bulk_reduce_add_f32_1d_hint(own - lane, slab, 4096, run_pol);
                            } else {
                                if (first_sub) bulk_store_1d(own - lane, slab, 4096);
                                else bulk_reduce_add_f32_1d(own - lane, slab, 4096);
                            }
                            bulk_commit();
                        }
                        ++store_seq;
                    } else if (!(p.dbg_flags & 8)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) __stcg(own + j * 32, make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    }
                    continue;
                }
                if (!is_head) {                              // raw partial -> this CTA's slot (the tile's head collects it)
                    float4* dst = my_part4 + (size_t)(c >> 5) * 4 * 8 * 32;
                    if (bulk_run) {
                        const uint32_t slab = stg_base + (uint32_t)((store_seq % p.store_bufs) * 4096);
                        if (lane == 0) bulk_wait_read(p.store_bufs - 1);
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(slab + (uint32_t)(j * 512 + lane * 16), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                         __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            bulk_store_1d(dst - lane, slab, 4096);
                            bulk_commit();
                        }
                        ++store_seq;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) __stcg(dst + j * 32, make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    }
                    continue;
                }
                if (nstaged > 0) {
                    const int buf = (c >> 5) & 1;
                    if (c + 32 < p.block_n) {
                        stage_issue(c + 32, buf ^ 1);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                    for (int hh = 0; hh < nstaged; ++hh) {                  // ascending CTA order: deterministic sum
                        const uint32_t src = stage_slot(buf, hh);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 t;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                         : "r"(src + (uint32_t)((j ^ (rr & 7)) << 4)));
                            f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
                        }
                    }
                }
                for (int h = worker + 1 + nstaged; h <= last_contrib; ++h) {      // overflow (tiny-K corner): direct loads
                    const float4* src = part4(h * NCTA + rank) + (size_t)(c >> 5) * 4 * 8 * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = __ldcg(src + j * 32);
                        f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], sb[c + j], sb[256 + c + j]);   // folded BN (1, 0 when absent)
                if (!row_ok && !p.tma_store) continue;       // (the TMA store clips rows >= M itself; all lanes must take part)
                if (p.mode == EPI_PLANES) {
                    if (p.out_hi) {                          // un-pooled tensor: leaky, hi/lo split, 64 B per plane
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float t0 = f[2 * j], t1 = f[2 * j + 1];
                            split_pack2(p.leaky ? fmaxf(t0, 0.1f * t0) : t0, p.leaky ? fmaxf(t1, 0.1f * t1) : t1, hi[j], lo[j]);
                        }
                        if (p.tma_store) {
                            // slab = [plane][32 rows][64 B], SWIZZLE_64B: 16-byte piece j of row r sits at j ^ ((r >> 1) & 3)
                            const uint32_t slab = stg_base + (uint32_t)((store_seq % p.store_bufs) * 4096);
                            if (lane == 0) bulk_wait_read(p.store_bufs - 1);   // the store that last read this slab is done
                            __syncwarp();
                            const uint32_t rowa = slab + (uint32_t)(lane * 64), sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                st_shared_v4(rowa + ((j ^ sw) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                                st_shared_v4(rowa + 2048u + ((j ^ sw) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_3d(&map_o, slab, n0 + c, mt * BLOCK_M + q * 32, 0);
                                bulk_commit();
                            }
                            ++store_seq;
                        } else {
                        // p.reorg (passthrough layer, spatial tiles): the un-pooled output goes straight to its space-to-depth place in
                        // the concat buffer, out[b, y/2, x/2, (dy*2+dx)*N + n] (model/yolo2/function.py:22-29) -- no reorg pass
                        const size_t off = p.reorg ? pool_row * (size_t)p.ldc + (size_t)reorg_sub * p.N + n0 + c : (size_t)row * p.ldc + n0 + c;
                        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + off);
                        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                            dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                        }
                        }
                    }
                    if (p.pool_hi) {
                        // fused 2x2/2 max-pool: the tile is a tx x ty spatial block, so the window partners of row r are
                        // rows r^1 (x) and r^tx (y) = lanes of the same warp.  Exchange-and-halve: after the x step each
                        // lane of a pair keeps 16 of the 32 columns, after the y step 8 -- the four lanes of a window end
                        // with disjoint quarters and each writes 16 B per plane (no idle lanes, 24 shuffles instead of 64).
                        // max runs on the exact fp32 BN outputs; leaky (monotone) is applied to the 8 survivors.
                        const bool odd_x = (lane & 1) != 0, odd_y = (lane & p.tx) != 0;
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
                            const float t = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, p.tx));
                            m[j] = p.leaky ? fmaxf(t, 0.1f * t) : t;
                        }
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) split_pack2(m[2 * j], m[2 * j + 1], hi[j], lo[j]);
                        const size_t off = pool_row * (size_t)p.ldp + n0 + c + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);
                        *reinterpret_cast<uint4*>(p.pool_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(p.pool_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                } else {  // EPI_F32: 128-bit stores when aligned, else masked scalar stores (N = 425, 125)
                    if (p.leaky) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.1f * f[j]);
                    }
                    float* dst = p.out_f32 + (size_t)row * p.ldc + n0 + c;
                    if (p.tma_store) {
                        // slab = [32 rows][128 B], SWIZZLE_128B: 16-byte piece j of row r sits at j ^ (r & 7)
                        const uint32_t slab = stg_base + (uint32_t)((store_seq % p.store_bufs) * 4096);
                        if (lane == 0) bulk_wait_read(p.store_bufs - 1);
                        __syncwarp();
                        const uint32_t rowa = slab + (uint32_t)(lane * 128), sw = (uint32_t)(lane & 7);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(rowa + ((j ^ sw) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                         __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&map_o, slab, n0 + c, mt * BLOCK_M + q * 32);
                            bulk_commit();
                        }
                        ++store_seq;
                    } else if ((p.ldc & 3) == 0 && n0 + c + 32 <= p.N) {
                        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                        for (int j = 0; j < 8; ++j) d4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + c + j < p.N) dst[j] = f[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(acc ? tempty_r1 : tempty_r0);
                else mbar_arrive(&tempty[acc]);
            }
            if (!is_head && last_sub) {                      // publish the partial: all 128 rows written -> flag
                if (bulk_run) {                              // (the bulk copies of this warp have landed)
                    if (lane == 0) {
                        bulk_wait_all();
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    __syncwarp();
                }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) flag_set(p.sk_flags + blockIdx.x, p.epoch);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (p.dbg && et == 0) p.dbg[blockIdx.x * 4 + 3] = gtime_ns();
        if (p.store_bufs && lane == 0) bulk_wait_read(0);   // the slabs have been read out before the CTA (its smem) retires
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();       // neither CTA's shared memory / TMEM goes away while the other can still touch it
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encodeTiled = nullptr;
static PFN_encodeIm2col g_encodeIm2col = nullptr;

static int load_driver_entry_points() {
    if (g_encodeTiled && g_encodeIm2col) return 0;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    Y2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    Y2_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_encodeTiled = reinterpret_cast<PFN_encodeTiled>(fn);
    fn = nullptr;
    Y2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
    Y2_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col not available from the driver");
    g_encodeIm2col = reinterpret_cast<PFN_encodeIm2col>(fn);
    return 0;
}

template <int BK, bool SPLIT3, bool PAIR>
static int launch_inst(const TcConvLaunch& L, cudaStream_t stream) {
    auto kern = conv_tc_kernel<BK, SPLIT3, PAIR>;
    static unsigned long long attr_seen = 0;
    if (first_use_on_current_device(attr_seen))
        Y2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    if (g_conv_pdl || PAIR) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = L.smem_bytes; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (g_conv_pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (PAIR) {                                  // CTA pair = cluster of 2 (the two SMs of one TPC)
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = attr; cfg.numAttrs = na;
        Y2_CUDA(cudaLaunchKernelEx(&cfg, kern, L.map_a, L.map_w, L.map_o, L.p));
    } else {
        kern<<<L.grid, NUM_THREADS, L.smem_bytes, stream>>>(L.map_a, L.map_w, L.map_o, L.p);
    }
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

static std::atomic<unsigned int> g_epoch{0};

// flags page + one partial tile and one running-sum tile per CTA
size_t tc_conv_streamk_bytes(int num_sms) { return 2 * (size_t)num_sms * (BLOCK_M * 256 * sizeof(float)) + 4096; }

int tc_conv_launch(const TcConvLaunch& Lc, cudaStream_t stream) {
    TcConvLaunch L = Lc;
    unsigned int e = g_epoch.fetch_add(1) + 1;
    if (e == 0) e = g_epoch.fetch_add(1) + 1;          // 0 is the "never written" value of a fresh flag buffer
    L.p.epoch = e;
    if (L.p.pair) return L.split3 ? launch_inst<64, true, true>(L, stream) : launch_inst<64, false, true>(L, stream);
    if (L.block_k == 64) return L.split3 ? launch_inst<64, true, false>(L, stream) : launch_inst<64, false, false>(L, stream);
    return L.split3 ? launch_inst<32, true, false>(L, stream) : launch_inst<32, false, false>(L, stream);
}

// CTA pairs that can be co-resident (every worker of the persistent schedule must be: stream-K heads spin on the flags of
// later workers).  Queried once per kernel instantiation with the real shared-memory size.
static int max_active_pairs(int split3, int smem_bytes) {
    static int cached[64][2];
    static unsigned long long seen[2] = {0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 63;
    if (!first_use_on_current_device(seen[split3 ? 1 : 0])) return cached[dev][split3 ? 1 : 0];
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * 74); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e;
    if (split3) {
        cudaFuncSetAttribute(conv_tc_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        e = cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<64, true, true>, &cfg);
    } else {
        cudaFuncSetAttribute(conv_tc_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        e = cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<64, false, true>, &cfg);
    }
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    cached[dev][split3 ? 1 : 0] = n;
    return n;
}

int tc_conv_check_watchdog() {
    Watchdog w;
    Y2_CUDA(cudaMemcpyFromSymbol(&w, g_watchdog, sizeof(w)));
    if (!w.fired) return 0;
    Watchdog z;
    memset(&z, 0, sizeof(z));
    cudaMemcpyToSymbol(g_watchdog, &z, sizeof(z));
    set_error("tcgen05 conv: barrier watchdog fired (block %u warp %u wait-site 0x%x parity %u): pipeline deadlock",
              w.block, w.warp, w.tag, w.parity);
    return -3;
}

// Spatial tiling for the fused max-pool epilogue: tx x ty pixels x tb images = 128 rows, tx a power of two in [2,16]
// (the window partners must be lanes r^1 and r^tx of one warp), exact cover of (W, H, B).
static bool pool_tiling(int B, int H, int W, int* tx, int* ty, int* tb) {
    if ((H & 1) || (W & 1)) return false;
    for (int x = 16; x >= 2; x >>= 1) {
        if (W % x) continue;
        for (int y = 128 / x; y >= 2; y >>= 1) {
            const int b = 128 / (x * y);
            if (x * y * b != 128 || x * y < 4) continue;
            if ((x * y) % 32 != 0 && 32 % (x * y) != 0) continue;     // a warp = whole rows of the block
            if (H % y == 0 && B % b == 0 && (y & 1) == 0) { *tx = x; *ty = y; *tb = b; return true; }
        }
    }
    return false;
}
bool tc_conv_can_fuse_pool(int B, int H, int W) {
    int a, b, c;
    return pool_tiling(B, H, W, &a, &b, &c);
}

int tc_conv_bind_output(TcConvLaunch* L) {
    ConvParams& p = L->p;
    p.tma_store = 0;
    if (p.store_bufs == 0 || p.tx != 0) return 0;
    if (load_driver_entry_points()) return -1;
    cuuint32_t estr[3] = {1, 1, 1};
    if (p.mode == EPI_F32) {
        if (!p.out_f32 || (p.ldc & 3) != 0 || (reinterpret_cast<uintptr_t>(p.out_f32) & 15) != 0) return 0;
        cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)p.M};
        cuuint64_t strides[1] = {(cuuint64_t)p.ldc * 4};
        cuuint32_t box[2] = {32, 32};
        CUresult r = g_encodeTiled(&L->map_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.out_f32, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32 output) failed (%d) M=%d N=%d ldc=%lld", (int)r, p.M, p.N, p.ldc);
        p.tma_store = 1;
    } else {
        if (!p.out_hi || !p.out_lo || p.out_lo <= p.out_hi || (p.ldc & 7) != 0 || (reinterpret_cast<uintptr_t>(p.out_hi) & 15) != 0 ||
            (((p.out_lo - p.out_hi) * 2) & 15) != 0)
            return 0;
        cuuint64_t dims[3] = {(cuuint64_t)p.N, (cuuint64_t)p.M, 2};
        cuuint64_t strides[2] = {(cuuint64_t)p.ldc * 2, (cuuint64_t)(p.out_lo - p.out_hi) * 2};
        cuuint32_t box[3] = {32, 32, 2};
        CUresult r = g_encodeTiled(&L->map_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.out_hi, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (plane output) failed (%d) M=%d N=%d ldc=%lld", (int)r, p.M, p.N, p.ldc);
        p.tma_store = 1;
    }
    return 0;
}

int g_conv_dbg_flags = 0, g_conv_force_halo = 0, g_conv_pdl = 1, g_conv_tma_store = 1, g_conv_force_pair = 0, g_conv_kcap = 32;
int g_conv_fmt = 0, g_wgrad_fmt = 0;

// CTA-pair mode: 64-channel k-blocks (SW128 operands), an even split of the N tile in 8-row swizzle groups, N a multiple of
// 16 (the cta_group::2 MMA shape rule); halo mode keeps the single-CTA kernel.
bool tc_conv_can_pair(int Cin, int block_n, int halo) {
    return !halo && Cin % 64 == 0 && block_n % 16 == 0 && (block_n / 2) % 8 == 0 && block_n >= 32;
}

// Halo mode: 3x3, 32 input channels, a single N tile whose 9 weight taps fit next to the halo ring.
bool tc_conv_can_halo(int B, int H, int W, int Cin, int ksize, int cout_pad, int block_n, int split3) {
    (void)B;
    const int planes = split3 ? 2 : 1;
    return ksize == 3 && Cin == 32 && cout_pad == block_n && (W % 8) == 0 && (H % 16) == 0 &&
           9 * planes * block_n * 64 + 2 * planes * 12288 + 1024 + SB_BYTES + BAR_BYTES <= SMEM_LIMIT;
}

int tc_conv_plan(TcConvLaunch* L, const bf16* in_planes, int B, int H, int W, int Cin, int ksize, const bf16* wpack,
                 int cout, int cout_pad, int block_n, int max_ctas, int split3, int num_sms, void* sk_ws, int fuse_pool,
                 int halo, int pair) {
    if (load_driver_entry_points()) return -1;
    Y2_REQUIRE(ksize == 1 || ksize == 3, "tc conv: ksize must be 1 or 3 (got %d)", ksize);
    Y2_REQUIRE(Cin % 32 == 0, "tc conv: Cin must be a multiple of 32 (got %d)", Cin);
    Y2_REQUIRE(block_n % 32 == 0 && block_n >= 32 && block_n <= 256, "tc conv: block_n %d invalid", block_n);
    Y2_REQUIRE(cout_pad % block_n == 0 && cout_pad >= cout, "tc conv: cout_pad %d vs block_n %d", cout_pad, block_n);
    Y2_REQUIRE((reinterpret_cast<uintptr_t>(in_planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(wpack) & 15) == 0,
               "tc conv: operands must be 16-byte aligned");
    memset(L, 0, sizeof(*L));
    // no co-resident CTA pair on this device / context (cluster launch refused): the single-CTA kernel computes the same thing
    if (pair && max_active_pairs(split3, SMEM_LIMIT) < 1) pair = 0;
    const int BK = (Cin % 64 == 0) ? 64 : 32;
    const int taps = ksize * ksize;
    const long long M = (long long)B * H * W;
    Y2_REQUIRE(M < (1ll << 31), "tc conv: too many pixels");
    ConvParams& p = L->p;
    p.M = (int)M; p.N = cout; p.Cin = Cin; p.ksize = ksize; p.B = B; p.H = H; p.W = W;
    p.block_n = block_n;
    p.fmt = g_conv_fmt;
    p.m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M);
    if (fuse_pool && !halo) {
        Y2_REQUIRE(pool_tiling(B, H, W, &p.tx, &p.ty, &p.tb), "tc conv: no spatial tiling for the fused max-pool at B=%d H=%d W=%d", B, H, W);
        p.tiles_x = W / p.tx; p.tiles_y = H / p.ty;
        p.m_tiles = p.tiles_x * p.tiles_y * (B / p.tb);          // exact cover: same count, different pixel order
    }
    if (halo) {
        Y2_REQUIRE(halo == 1, "tc conv: halo mode %d invalid", halo);
        Y2_REQUIRE(tc_conv_can_halo(B, H, W, Cin, ksize, cout_pad, block_n, split3), "tc conv: halo mode not applicable (B=%d H=%d W=%d Cin=%d k=%d N=%d)", B, H, W, Cin, ksize, block_n);
        p.tx = 8; p.ty = 16; p.tb = 1;                           // 8 x 16 pixel blocks: one 8-row MMA group per pixel row
        p.tiles_x = W / 8; p.tiles_y = H / 16;
        p.m_tiles = p.tiles_x * p.tiles_y * B;
        p.halo = halo;
        p.halo_tx_bytes = 10 * 18 * 64;
        p.halo_plane_bytes = 12288;                              // 1 KiB multiple: swizzle phase identical per plane
        p.bres_bytes = 9 * (split3 ? 2 : 1) * block_n * 64;
    }
    p.m_tiles128 = p.m_tiles;
    if (pair) {
        Y2_REQUIRE(tc_conv_can_pair(Cin, block_n, halo), "tc conv: CTA-pair mode not applicable (Cin=%d block_n=%d halo=%d)", Cin, block_n, halo);
        p.pair = 1;
        p.m_tiles = (p.m_tiles128 + 1) / 2;                      // a pair owns two consecutive 128-row tiles
    }
    p.dbg_flags = g_conv_dbg_flags;
    p.n_tiles = cout_pad / block_n;
    p.kblocks_total = halo ? 1 : taps * (Cin / BK);
    p.cout_pad = cout_pad;
    Y2_REQUIRE(sk_ws && (reinterpret_cast<uintptr_t>(sk_ws) & 15) == 0, "tc conv: stream-K workspace missing/unaligned");
    p.sk_flags = static_cast<unsigned int*>(sk_ws);                         // [num_sms] (first 4 KiB)
    p.sk_partial = reinterpret_cast<float*>(static_cast<char*>(sk_ws) + 4096);
    p.sk_run = p.sk_partial + (size_t)num_sms * BLOCK_M * 256;
    p.kcap = g_conv_kcap;
    const int planes = split3 ? 2 : 1;
    const int b_rows = pair ? block_n / 2 : block_n;            // weight-tile rows staged per CTA
    const int stage_bytes = halo ? planes * p.halo_plane_bytes : planes * (BLOCK_M * BK * 2 + b_rows * BK * 2);
    if (p.tx == 0 && g_conv_tma_store) {          // linear tiles: room for the TMA-store staging slabs (2 per warp if the ring keeps its depth)
        const int avail = SMEM_LIMIT - 1024 - SB_BYTES - BAR_BYTES;
        int st0 = avail / stage_bytes; if (st0 > 8) st0 = 8;
        int st2 = (avail - 32768) / stage_bytes; if (st2 > 8) st2 = 8;
        int st1 = (avail - 16384) / stage_bytes; if (st1 > 8) st1 = 8;
        p.store_bufs = (st2 == st0 || st2 >= 4) ? 2 : ((st1 == st0 || st1 >= 3) ? 1 : 0);
        p.stg_bytes = p.store_bufs * 16384;
    }
    int stages = (SMEM_LIMIT - 1024 - SB_BYTES - BAR_BYTES - p.bres_bytes - p.stg_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    Y2_REQUIRE(stages >= 2, "tc conv: tile does not fit shared memory");
    p.num_stages = stages;
    L->smem_bytes = stages * stage_bytes + 1024 + SB_BYTES + BAR_BYTES + p.bres_bytes + p.stg_bytes;
    L->block_k = BK;
    L->split3 = split3 ? 1 : 0;
    int workers = num_sms;
    if (pair) {
        workers = max_active_pairs(split3, L->smem_bytes);
        Y2_REQUIRE(workers >= 1, "tc conv: no CTA pair can be resident (cluster launch unavailable?)");
        if (workers > num_sms / 2) workers = num_sms / 2;
    }
    choose_schedule((long long)p.m_tiles * p.n_tiles, p.kblocks_total, workers, max_ctas, (block_n / 256.0) * (BK / 64.0),
                    (size_t)planes * cout_pad * taps * Cin * sizeof(bf16), &p.dp_tiles,
                    &p.sk_ctas, &L->grid);
    if (pair) L->grid *= 2;
    if (halo) {                                                  // one k-block per tile: plain data-parallel waves
        p.dp_tiles = p.m_tiles; p.sk_ctas = 0;
        L->grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
        if (max_ctas > 0 && L->grid > max_ctas) L->grid = max_ctas;
    }
    Y2_REQUIRE(L->grid <= 1024, "tc conv: grid too large for the flag page");

    if (halo) {
        // halo map: (C, W, H, N=2B) bf16, TILE mode, box {32, 10, 18, 1}; out-of-image = zero fill = SAME padding
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(2 * B)};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {32u, 10u, 18u, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encodeTiled(&L->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in_planes), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (halo activation) failed (%d) B=%d H=%d W=%d", (int)r, B, H, W);
    } else if (fuse_pool) {
        // activation map for the fused max-pool layers: (C, W, H, N=2B) bf16, TILE mode, box {BK, tx, ty, tb}
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(2 * B)};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)p.tx, (cuuint32_t)p.ty, (cuuint32_t)p.tb};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encodeTiled(&L->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in_planes), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D activation) failed (%d) B=%d H=%d W=%d Cin=%d box=%dx%dx%d",
                   (int)r, B, H, W, Cin, p.tx, p.ty, p.tb);
    } else
    // activation map: (C, W, H, N=2B) bf16, im2col mode, BLOCK_M pixels x BK channels per load
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(2 * B)};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const int padv = ksize / 2;
        int lower[2] = {-padv, -padv};
        int upper[2] = {padv - (ksize - 1), padv - (ksize - 1)};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encodeIm2col(&L->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in_planes), dims,
                                    strides, lower, upper, (cuuint32_t)BK, (cuuint32_t)BLOCK_M, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) for B=%d H=%d W=%d Cin=%d k=%d", (int)r, B,
                   H, W, Cin, ksize);
        // Driver workaround (also applied by CUTLASS): for tensors smaller than 128 KiB old drivers
        // set a bit in the im2col descriptor that must be cleared.
        int drv = 0;
        cudaDriverGetVersion(&drv);
        if (drv <= 13010 && (size_t)2 * B * H * W * Cin * 2 < 131072)
            reinterpret_cast<uint64_t*>(&L->map_a)[1] &= ~(1ull << 21);
    }
    // weight map: (K, 2*cout_pad) bf16 tiled, BK x block_n box
    {
        const cuuint64_t K = (cuuint64_t)taps * Cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)(2 * cout_pad)};
        cuuint64_t strides[1] = {K * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)b_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encodeTiled(&L->map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(wpack), dims,
                                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        Y2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for weights K=%llu cout_pad=%d", (int)r,
                   (unsigned long long)K, cout_pad);
    }
    return 0;
}

}  // namespace y2
