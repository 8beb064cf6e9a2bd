// sm_100a PTX wrappers used by the tcgen05 conv kernels: mbarrier, TMA (tiled + im2col),
// TMEM allocation, tcgen05.mma / commit / ld, fences.  Hand-written inline PTX; no CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace y2 {

// Watchdog: a mis-programmed barrier must never hang the (shared, budgeted) GPU box.
// Every mbarrier wait is bounded; on expiry the first waiter records where it was and raises
// `fired`; every other wait then falls through at once, the kernel drains (with garbage
// results) and the host turns the record into an error. One copy per translation unit.
struct Watchdog {
    unsigned int fired;
    unsigned int block, warp, tag, parity;
};
static __device__ Watchdog g_watchdog;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
static __device__ __noinline__ void watchdog_fire(uint32_t tag, uint32_t parity) {
    if (atomicCAS(&g_watchdog.fired, 0u, 1u) == 0u) {
        g_watchdog.block = blockIdx.x;
        g_watchdog.warp = threadIdx.x >> 5;
        g_watchdog.tag = tag;
        g_watchdog.parity = parity;
        __threadfence();
    }
}
// tag identifies the wait site in the watchdog record.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(&g_watchdog.fired)) return;
            if (clock64() - t0 > 2000000000LL) {   // ~1 s
                watchdog_fire(tag, parity);
                return;
            }
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// im2col-mode load on an NHWC tensor described as (C, W, H, N): `pixels` consecutive output
// positions starting at base pixel (w, h, n), each read at base + (off_w, off_h); out-of-image
// reads are zero-filled (= SAME padding).
__device__ __forceinline__ void tma_load_im2col_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

// ---- warp-converged variants: every lane of the warp executes the call, the lane with `leader` != 0 (= elect_one(),
// evaluated once by the role) issues the operation.  Keeping the role loops converged lets the compiler hold addresses /
// descriptors / coordinates in UNIFORM registers, which is what UTMALDG / UTCHMMA take; issuing from an `if (lane == 0)`
// region instead costs a convergence "waterfall" (ELECT + 5 R2UR + branch) per instruction -- more than a
// 128x64x16 MMA lasts.
__device__ __forceinline__ void mbar_expect_tx_e(uint32_t leader, uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %2, 0;\n\t"
                 "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n"
                 ::"r"(smem_u32(bar)), "r"(bytes), "r"(leader) : "memory");
}
__device__ __forceinline__ void mbar_arrive_e(uint32_t leader, uint64_t* bar) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %1, 0;\n\t"
                 "@pe mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(uint32_t leader, void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_e(uint32_t leader, void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                              int c2, int c3) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %7, 0;\n\t"
        "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_e(uint32_t leader, void* dst, const CUtensorMap* m, uint64_t* bar, int c,
                                                     int w, int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %9, 0;\n\t"
        "@pe cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h), "r"(leader)
        : "memory");
}

// ---------------------------------------------------------------- TMEM / tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16/fp16 in, fp32 accumulate), one CTA.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// Descriptors travel as (lo, hi) 32-bit halves: hi (SBO / version / swizzle mode) is constant per operand kind and all
// address arithmetic happens on lo = (addr >> 4) | LBO << 16 with 32-bit adds (uniform-datapath friendly).
__device__ __forceinline__ void tc_mma_f16_e(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred pe, p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 pe, %7, 0;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint32_t leader, uint64_t* bar) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %1, 0;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                 ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
// ---------------------------------------------------------------- CTA pair (cta_group::2): two SMs of one TPC, one MMA
// The pair is a cluster of 2.  Rank 0 issues every tcgen05.mma for both SMs (M = 256: each CTA's TMEM holds its own 128 rows,
// each CTA's shared memory holds its own A rows and HALF of the B tile); both CTAs' TMA loads complete on rank 0's barrier;
// tcgen05.commit multicasts its arrival to the barrier at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {      // shared::cta address -> shared::cluster address in `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {                                    // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {   // one warp (same warp id) in EACH CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair_e(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred pe, p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 pe, %7, 0;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair_e(uint32_t leader, uint64_t* bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("{\n\t.reg .pred pe;\n\t.reg .b16 mask;\n\tsetp.ne.b32 pe, %1, 0;\n\tmov.b16 mask, 3;\n\t"
                 "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], mask;\n\t}\n"
                 ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
// TMA loads of a pair: destination = this CTA's shared memory, completion bytes -> `bar_cluster` (rank 0's barrier)
__device__ __forceinline__ void tma_load_2d_pair_e(uint32_t leader, void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair_e(uint32_t leader, void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                   int c2, int c3) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %7, 0;\n\t"
        "@pe cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair_e(uint32_t leader, void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c,
                                                          int w, int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %9, 0;\n\t"
        "@pe cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h), "r"(leader)
        : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// Wait for the outstanding tcgen05.ld of this thread; `r` is listed as in/out so that no use of the loaded registers
// can be scheduled above the wait (needed when the load is issued ahead of its consumer).
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- tile schedule (hybrid data-parallel + stream-K)
// Tiles [0, dp_tiles) are processed whole, round-robin over the grid (all CTAs walk K in lockstep, which keeps the
// operand slices they share hot in L2).  The remaining tiles form a stream-K space of (tiles - dp_tiles) * KB
// k-blocks cut into sk_ctas equal contiguous ranges; a range that starts inside a tile yields a partial.
struct SegIter {
    int dp_next, dp_tiles, stride, KB;
    long long cur, end;
    __device__ __forceinline__ void init(int dp_tiles_, int sk_ctas, long long sk_total, int KB_) {
        init_w((int)blockIdx.x, (int)gridDim.x, dp_tiles_, sk_ctas, sk_total, KB_);
    }
    // worker = CTA (or CTA pair) index, nworkers = how many walk the schedule
    __host__ __device__ __forceinline__ void init_w(int worker, int nworkers, int dp_tiles_, int sk_ctas, long long sk_total, int KB_) {
        dp_next = worker; dp_tiles = dp_tiles_; stride = nworkers; KB = KB_;
        if (worker < sk_ctas) {
            cur = sk_total * worker / sk_ctas;
            end = sk_total * (worker + 1) / sk_ctas;
        } else {
            cur = end = 0;
        }
    }
    // K-aligned split: CTA i owns k-range (i / tiles) of tile (i % tiles) -- all CTAs sit at the same K phase of their
    // tiles at the same time, so the operand slices the tiles share are fetched from HBM once and re-read from L2.
    __device__ __forceinline__ void init_split(int tiles, int chunks, int KB_) {
        dp_next = 0; dp_tiles = 0; stride = gridDim.x; KB = KB_;
        const int t = (int)blockIdx.x % tiles, c = (int)blockIdx.x / tiles;
        cur = (long long)t * KB + (long long)KB * c / chunks;
        end = (long long)t * KB + (long long)KB * (c + 1) / chunks;
    }
    __host__ __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
        if (dp_next < dp_tiles) {
            tile = dp_next; kb0 = 0; kb1 = KB; dp_next += stride;
            return true;
        }
        if (cur < end) {
            const int rt = (int)(cur / KB);
            tile = dp_tiles + rt;
            kb0 = (int)(cur - (long long)rt * KB);
            const long long left = end - cur;
            kb1 = kb0 + (int)((long long)(KB - kb0) < left ? (long long)(KB - kb0) : left);
            cur += kb1 - kb0;
            return true;
        }
        return false;
    }
};

// Accumulation-chain cap.  The tensor core adds every MMA into the fp32 TMEM accumulator with truncation, so a long chain
// biases the sum towards zero by ~(chain length) x 2^-24 -- measured: the 13x13 layers at batch 32 (171 k-blocks per
// chain) carried 1.9e-4 to the network output against 4e-5 with 6-k-block chains (tools/diag_layers.py).  Segments of the
// schedule are therefore cut into sub-segments of at most `cap` k-blocks, each accumulated from zero in the next TMEM
// buffer; the epilogue warps (idle during the main loop) add the sub-results with round-to-nearest fp32 adds in the CTA's
// own partial slot.  The producer never notices; the MMA warp only sees more, shorter "segments".
struct CapIter {
    SegIter it;
    int cap, tile_, a, b, pos;
    __host__ __device__ __forceinline__ void init(int worker, int nworkers, int dp_tiles, int sk_ctas, long long sk_total, int KB, int cap_) {
        it.init_w(worker, nworkers, dp_tiles, sk_ctas, sk_total, KB);
        cap = cap_ > 0 ? cap_ : 0x7fffffff;
        a = b = pos = 0; tile_ = 0;
    }
    // for callers that initialise `it` themselves (SegIter::init / init_split)
    __host__ __device__ __forceinline__ void wrap(int cap_) {
        cap = cap_ > 0 ? cap_ : 0x7fffffff;
        a = b = pos = 0; tile_ = 0;
    }
    // [kb0, kb1) = next sub-segment of segment [seg_a, seg_b) of `tile`
    __host__ __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1, int& seg_a, int& seg_b) {
        if (pos >= b) {
            if (!it.next(tile_, a, b)) return false;
            pos = a;
        }
        const int len = b - pos;
        int step = len;
        if (len > cap) {                              // equal pieces, none longer than cap
            const int n = (len + cap - 1) / cap;
            step = (len + n - 1) / n;
        }
        tile = tile_; kb0 = pos; kb1 = pos + step; seg_a = a; seg_b = b;
        pos = kb1;
        return true;
    }
};

// Shared-memory matrix descriptor, K-major operand, rows of `row_bytes` (128 -> SWIZZLE_128B,
// 64 -> SWIZZLE_64B); 8-row groups are `8*row_bytes` apart (SBO). Bit layout as in the PTX ISA
// "matrix descriptor" for tcgen05 (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64)).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
    const uint64_t layout = (row_bytes == 128) ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;                          // LBO (unused for swizzled K-major)
    d |= static_cast<uint64_t>((8u * row_bytes) >> 4) << 32;      // SBO
    d |= static_cast<uint64_t>(1) << 46;                          // descriptor version (sm_100)
    d |= layout << 61;
    return d;
}
// Same, with an explicit 8-row-group pitch (SBO) and matrix base offset: used when the 128 rows of the operand are a
// shifted window of a larger swizzled tile (start address not aligned to the swizzle atom, groups not 8 rows apart).
__device__ __forceinline__ uint64_t make_kmajor_desc_ex(uint32_t smem_addr, uint32_t row_bytes, uint32_t sbo_bytes,
                                                        uint32_t base_offset) {
    const uint64_t layout = (row_bytes == 128) ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_offset & 7u) << 49;
    d |= layout << 61;
    return d;
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M x N tile.
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// The same with the 16-bit operand formats chosen per operand (kind::f16 takes f16 or bf16 for A and for B independently:
// instruction-descriptor fields a_format [7,10) and b_format [10,13), 0 = f16, 1 = bf16).
__host__ __device__ __forceinline__ uint32_t make_idesc_16(uint32_t m, uint32_t n, bool a_f16, bool b_f16) {
    return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace y2
