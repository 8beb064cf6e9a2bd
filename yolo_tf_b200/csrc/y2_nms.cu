// Greedy per-class IoU NMS, bit-exact with the reference's utils/postprocess.py:21-51
// (iou :21-36, non_max_suppress :39-51) as executed under NumPy >= 2 (float32 everywhere).
//
// Reference semantics reproduced exactly:
//   * class c visits boxes in the order of Python's stable descending sort carried across classes
//     = lexicographic (conf[:,c] desc, conf[:,c-1] desc, ..., conf[:,0] desc, index asc) on the
//     ORIGINAL scores;
//   * a box whose class score is > threshold when reached ("live") zeroes the class score of EVERY
//     later box - candidate or not - whose IoU with it is >= threshold_iou; zeroed boxes never act;
//   * IoU uses the float32 operation order of iou() with no FMA contraction (explicit _rn intrinsics; the
//     translation unit is also compiled with -fmad=false).
//
// Three kernels, no host round trip:
//   select : a CTA owns 32 classes of one image.  Candidates (> threshold) are compacted with coalesced 128-bit reads
//            (a lane reads 4 consecutive classes of one box, a warp 4 boxes x 32 classes) into per-class shared-memory
//            lists.  Then one warp per class: (value, index) keys of its candidates are staged in shared memory and
//            sorted by a bitonic network (O(K log^2 K) shared-memory compare-exchanges); runs of EQUAL values -- the only
//            place the lexicographic comparator has to look at earlier class columns, still unmodified because nothing is
//            written here -- are re-ranked exactly.  The candidates' boxes are staged in sorted order and the sequential
//            greedy sweep runs entirely out of shared memory (alive bitmask walked with ffs, lanes = later candidates,
//            the IEEE divide skipped for pairs a conservative bound already rejects).  Outputs: the list of KEPT boxes
//            and a bitmask of the SUPPRESSED candidates per (image, class).  Classes with more than SEL_CAP candidates
//            take the general path (lists in global memory, rank sort).
//   order  : (optional) final permutation of the N boxes = the order of the list the reference returns.
//   apply  : a CTA owns a [128 boxes][32 classes] tile (coalesced 128-bit loads through shared memory); lanes = boxes
//            (4 per lane), work items = (class column, chunk of 32 kept boxes) handed out to the warps dynamically, the
//            chunk's boxes broadcast from shared memory.  A candidate is zeroed iff its bit in the suppressed mask is set;
//            a non-candidate (it sorts after every candidate) is zeroed iff any kept box overlaps it >= threshold_iou.
// Scores are read once by select and once by apply, both coalesced; only tiles that change are written back.
#include "y2_internal.h"
#include "y2_nms_iou.cuh"

namespace y2 {

int g_nms_apply_mode = 0;
int g_nms_select_cg = 0;          // y2_debug_set(13, 0 | 8 | 16 | 32): classes per select CTA (0 = by regime)
static constexpr int NMS_WARPS = 8;
static constexpr int NMS_MAX_N = 8192;     // bitmasks: 256 words per warp
static constexpr int SEL_CAP = 256;        // candidates per class handled out of shared memory

__device__ __forceinline__ float4 load_box(const float* __restrict__ xy_min, const float* __restrict__ xy_max,
                                           size_t i) {
    const float2 lo = __ldg(reinterpret_cast<const float2*>(xy_min) + i);
    const float2 hi = __ldg(reinterpret_cast<const float2*>(xy_max) + i);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// does box i come before box j in class c's visiting order? (i != j)
__device__ __forceinline__ bool precedes(const float* __restrict__ conf_img, int C, int c, int i, float vi, int j,
                                         float vj) {
    if (vi > vj) return true;
    if (vi < vj) return false;
    for (int cc = c - 1; cc >= 0; --cc) {
        const float a = __ldg(conf_img + (size_t)i * C + cc), b = __ldg(conf_img + (size_t)j * C + cc);
        if (a > b) return true;
        if (a < b) return false;
    }
    return i < j;
}
static constexpr int KEPT_HAS_SUPP = 1 << 30;
struct NmsArgs {
    float* conf;
    const float* xy_min;
    const float* xy_max;
    int B, N, C, W;      // W = ceil(N / 32) mask words per (image, class)
    float thr, thr_iou;
    uint16_t* cand;      // [B][C][N] kept list (general path: candidates in index order first)
    uint16_t* sorted;    // [B][C][N] scratch of the general path: candidates in visiting order
    uint32_t* supp;      // [B][C][W] bit n = candidate n of this class was suppressed (written for classes with candidates)
    int* kept_cnt;       // [B][C] kept boxes of the class | KEPT_HAS_SUPP if some candidate of it was suppressed
    int* status;         // [B] nullable: 1 = a reference assert (NaN / xy_min > xy_max) would fire
    int* order_out;      // [B][N] nullable
    uint16_t* area_perm; // [B][N] boxes of the image in ascending order of area (apply's tile order)
    int apply_mode;      // diagnostics: 0 = choose by regime, 1 = chunk items handed out dynamically, 2 = one item per class column
    int cg;              // select: classes per CTA (32, 16 or 8: narrower groups when the batch alone cannot fill the machine)
    int coop;            // select: latency regime -- a class with more than COOP_MIN candidates gets the whole CTA
};

// General path for one (image, class), one warp: any number of candidates, lists in global memory, rank sort.
// alive / supp: per-warp shared-memory bitmasks of NMS_MAX_N bits.
__device__ void select_class_general(const NmsArgs& a, int b, int c, uint32_t* alive, uint32_t* supp, int lane) {
    const float* conf_img = a.conf + (size_t)b * a.N * a.C;
    const float* bmin = a.xy_min + (size_t)b * a.N * 2;
    const float* bmax = a.xy_max + (size_t)b * a.N * 2;
    uint16_t* cand = a.cand + ((size_t)b * a.C + c) * a.N;
    uint16_t* sorted = a.sorted + ((size_t)b * a.C + c) * a.N;
    // compaction in index order (ballot)
    int K = 0;
    for (int n0 = 0; n0 < a.N; n0 += 32) {
        const int n = n0 + lane;
        const bool is = n < a.N && __ldg(conf_img + (size_t)n * a.C + c) > a.thr;
        const uint32_t m = __ballot_sync(0xffffffffu, is);
        if (is) cand[K + __popc(m & ((1u << lane) - 1u))] = (uint16_t)n;
        K += __popc(m);
    }
    __syncwarp();
    for (int j0 = 0; j0 < K; j0 += 32) {
        const int jj = j0 + lane;
        const int j = (jj < K) ? cand[jj] : 0;
        const float vj = (jj < K) ? __ldg(conf_img + (size_t)j * a.C + c) : 0.f;
        int rank = 0;
        for (int ii = 0; ii < K; ++ii) {
            const int i = cand[ii];
            const float vi = __ldg(conf_img + (size_t)i * a.C + c);
            if (jj < K && i != j && precedes(conf_img, a.C, c, i, vi, j, vj)) ++rank;
        }
        if (jj < K) sorted[rank] = (uint16_t)j;
    }
    for (int w = lane; w < (K + 31) / 32; w += 32) alive[w] = 0xffffffffu;
    for (int w = lane; w < a.W; w += 32) supp[w] = 0u;
    __syncwarp();
    int kept = 0;
    for (int r = 0; r < K; ++r) {
        if (!((alive[r >> 5] >> (r & 31)) & 1u)) continue;           // warp-uniform
        const int i = sorted[r];
        const float4 bi = load_box(bmin, bmax, i);
        if (lane == 0) cand[kept] = (uint16_t)i;                     // kept list reuses cand[] (kept <= r)
        ++kept;
        for (int j0 = (r + 1) & ~31; j0 < K; j0 += 32) {
            const int jj = j0 + lane;
            bool kill = false;
            int j = 0;
            if (jj > r && jj < K && ((alive[jj >> 5] >> (jj & 31)) & 1u)) {
                j = sorted[jj];
                kill = iou_ref(bi, load_box(bmin, bmax, j)) >= a.thr_iou;
            }
            const uint32_t km = __ballot_sync(0xffffffffu, kill);
            if (kill) atomicOr(&supp[j >> 5], 1u << (j & 31));
            if (km && lane == 0) alive[j0 >> 5] &= ~km;
            __syncwarp();
        }
    }
    uint32_t* supp_out = a.supp + ((size_t)b * a.C + c) * a.W;
    for (int w = lane; w < a.W; w += 32) supp_out[w] = supp[w];
    if (lane == 0) a.kept_cnt[(size_t)b * a.C + c] = kept | (kept < K ? KEPT_HAS_SUPP : 0);
    __syncwarp();
}

// Greedy sweep over K <= 32 * NW candidates staged in visiting order (box[], shared memory).  The alive mask lives in
// REGISTERS (lane w owns word w) and every lane keeps ITS candidates (jj = 32 * wd + lane) in registers; a live box r is
// tested against all later candidates in one fully unrolled pass whose NW word-iterations are independent (the alive words are
// read before any of them is updated), so the loads / IoU tests / ballots of different words overlap instead of forming one
// dependent chain per word.  Returns the number of kept boxes (their indices in list[]) and the final alive mask.
template <int NW>
__device__ __forceinline__ int nms_sweep(const float4* __restrict__ box, const unsigned long long* __restrict__ key, uint16_t* __restrict__ list,
                                         int K, float thr_iou, float thr_lo, bool quick, int lane, uint32_t& alive_out) {
    uint32_t alive = 0u;
    if (lane < NW && lane * 32 < K) alive = (lane * 32 + 32 <= K) ? 0xffffffffu : ((1u << (K - lane * 32)) - 1u);
    float4 bj[NW];
    float aj[NW];
#pragma unroll
    for (int wd = 0; wd < NW; ++wd) {
        const int jj = wd * 32 + lane;
        bj[wd] = jj < K ? box[jj] : make_float4(0.f, 0.f, 0.f, 0.f);
        aj[wd] = box_area(bj[wd]);
    }
    int kept = 0;
#pragma unroll 1
    for (int g = 0; g < NW; ++g) {
        uint32_t m = __shfl_sync(0xffffffffu, alive, g);
        while (m) {                                                   // warp-uniform
            const int bit = __ffs(m) - 1;
            const int r = g * 32 + bit;
            const float4 bi = box[r];
            const float ai = box_area(bi);
            if (lane == 0) list[kept] = (uint16_t)(0xffff - (int)(key[r] & 0xffffu));   // kept <= r: slot already consumed
            ++kept;
            uint32_t km_mine = 0u;
#pragma unroll
            for (int wd = 0; wd < NW; ++wd) {
                if (wd >= g) {                                        // warp-uniform
                    const int jj = wd * 32 + lane;
                    const uint32_t aw = __shfl_sync(0xffffffffu, alive, wd);
                    const bool kill = jj > r && ((aw >> lane) & 1u) && iou_hit(bi, ai, bj[wd], aj[wd], thr_iou, thr_lo, quick);
                    const uint32_t km = __ballot_sync(0xffffffffu, kill);
                    if (lane == wd) km_mine = km;
                }
            }
            alive &= ~km_mine;
            m = __shfl_sync(0xffffffffu, alive, g) & (bit == 31 ? 0u : (0xffffffffu << (bit + 1)));
        }
    }
    alive_out = alive;
    return kept;
}

// One class with many candidates (COOP_MIN < K <= COOP_CAP) handled by the WHOLE CTA -- the latency-bound regime (few CTAs,
// e.g. a batch of 32 images whose scores put hundreds of candidates into one class): a single warp's greedy sweep is a chain
// of ~300 instructions per kept box, 160 kept boxes long.  Here the 256 threads own up to 4 candidates each: cooperative
// bitonic sort (compare-exchanges spread over the CTA, one __syncthreads per stage), and in the sweep warp w owns the alive
// words w, w+8, w+16, w+24 -- every kept box costs at most four 32-wide IoU tests per warp and one __syncthreads.
// All threads of the CTA call this with the same arguments.  pool: the per-warp areas of the select kernel, free in this phase.
static constexpr int COOP_MIN = 64, COOP_CAP = 1024, COOP_R = COOP_CAP / (NMS_WARPS * 32);
__host__ __device__ inline size_t coop_pool_bytes(int W) { return (size_t)COOP_CAP * (8 + 16 + 2) + (size_t)((W + 3) & ~3) * 4 + 32 * 4; }
__device__ void select_class_coop(const NmsArgs& a, int b, int c, int K, unsigned char* pool, const uint16_t* cand_list, bool quick,
                                  float thr_lo) {
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    constexpr int NT = NMS_WARPS * 32;
    unsigned long long* key = reinterpret_cast<unsigned long long*>(pool);
    float4* box = reinterpret_cast<float4*>(pool + COOP_CAP * 8);
    uint16_t* list = reinterpret_cast<uint16_t*>(pool + COOP_CAP * 24);           // candidates, later the kept list
    uint32_t* supp = reinterpret_cast<uint32_t*>(pool + COOP_CAP * 26);
    uint32_t* alive_sm = supp + ((a.W + 3) & ~3);                                  // [32]
    const float* conf_img = a.conf + (size_t)b * a.N * a.C;
    const float* bmin = a.xy_min + (size_t)b * a.N * 2;
    const float* bmax = a.xy_max + (size_t)b * a.N * 2;
    __shared__ int scan_sm[NMS_WARPS], scan_base;
    if (cand_list) {                                                  // K <= SEL_CAP: the compaction pass recorded all of them
        for (int p = t; p < K; p += NT) list[p] = cand_list[p];
    } else {                                                          // more: re-scan the class column (index order, block scan)
        if (t == 0) scan_base = 0;
        __syncthreads();
        for (int n0 = 0; n0 < a.N; n0 += NT) {
            const int n = n0 + t;
            const bool is = n < a.N && __ldg(conf_img + (size_t)n * a.C + c) > a.thr;
            const uint32_t m = __ballot_sync(0xffffffffu, is);
            if (lane == 0) scan_sm[warp] = __popc(m);
            __syncthreads();
            int pos = scan_base + __popc(m & ((1u << lane) - 1u));
            for (int w = 0; w < warp; ++w) pos += scan_sm[w];
            if (is) list[pos] = (uint16_t)n;
            __syncthreads();
            if (t == 0) { int s = 0; for (int w = 0; w < NMS_WARPS; ++w) s += scan_sm[w]; scan_base += s; }
            __syncthreads();
        }
    }
    for (int w = t; w < a.W; w += NT) supp[w] = 0u;
    __syncthreads();
    int n2 = 128;
    while (n2 < K) n2 <<= 1;
    for (int p = t; p < n2; p += NT) {
        unsigned long long k = 0ull;                                  // padding sorts last (every real key is > 0)
        if (p < K) {
            const int idx = list[p];
            k = ((unsigned long long)ford(__ldg(conf_img + (size_t)idx * a.C + c)) << 32) | (unsigned long long)(0xffffu - idx);
        }
        key[p] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int q = t; q < (n2 >> 1); q += NT) {
                const int lo = ((q & ~(stride - 1)) << 1) | (q & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = key[lo], y = key[hi];
                if ((x < y) == desc) { key[lo] = y; key[hi] = x; }
            }
            __syncthreads();
        }
    }
    if (c > 0) {                                                      // ties descend into earlier class columns
        bool tie = false;
        for (int p = t; p < K; p += NT) {
            const uint32_t v = (uint32_t)(key[p] >> 32);
            tie |= (p > 0 && (uint32_t)(key[p - 1] >> 32) == v) || (p + 1 < K && (uint32_t)(key[p + 1] >> 32) == v);
        }
        if (__syncthreads_or(tie ? 1 : 0)) {
            unsigned long long* tmp = reinterpret_cast<unsigned long long*>(box);
            for (int p = t; p < K; p += NT) {
                const unsigned long long kp = key[p];
                const uint32_t v = (uint32_t)(kp >> 32);
                int s0 = p, e0 = p + 1;
                while (s0 > 0 && (uint32_t)(key[s0 - 1] >> 32) == v) --s0;
                while (e0 < K && (uint32_t)(key[e0] >> 32) == v) ++e0;
                int rank = p;
                if (e0 - s0 > 1) {
                    rank = s0;
                    const int j = 0xffff - (int)(kp & 0xffffu);
                    for (int qq = s0; qq < e0; ++qq) {
                        if (qq == p) continue;
                        const int i = 0xffff - (int)(key[qq] & 0xffffu);
                        if (precedes(conf_img, a.C, c, i, 0.f, j, 0.f)) ++rank;
                    }
                }
                tmp[rank] = kp;
            }
            __syncthreads();
            for (int p = t; p < K; p += NT) key[p] = tmp[p];
            __syncthreads();
        }
    }
    // boxes in visiting order; thread (warp w, lane l) owns the candidates (w + 8 j) * 32 + l, warp w the alive words w + 8 j
    const int words = (K + 31) / 32;
    float4 bj[COOP_R];
    float aj[COOP_R];
    uint32_t alive_w[COOP_R];
#pragma unroll
    for (int j = 0; j < COOP_R; ++j) {
        const int wd = warp + NMS_WARPS * j, p = wd * 32 + lane;
        bj[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < K) { bj[j] = load_box(bmin, bmax, 0xffff - (int)(key[p] & 0xffffu)); box[p] = bj[j]; }
        aj[j] = box_area(bj[j]);
        alive_w[j] = 0u;
        if (wd < words) {
            alive_w[j] = (wd * 32 + 32 <= K) ? 0xffffffffu : ((1u << (K - wd * 32)) - 1u);
            if (lane == 0) alive_sm[wd] = alive_w[j];
        }
    }
    __syncthreads();
    // Greedy sweep, one 32-candidate word per round (round 2: was one kept box per round, i.e. a CTA barrier and two dependent
    // shared-memory round trips per kept box -- 0.3 us each, measured): the warp that owns word g settles it alone (boxes of the word
    // suppress later boxes of the word, strictly in order), publishes the survivors, and every warp then tests its later words
    // against all of them without further synchronisation.  A candidate is still suppressed only by earlier boxes that are
    // themselves alive when their turn comes: the reference's order of events.
    int kept = 0;
    for (int g = 0; g < words; ++g) {                                 // CTA-uniform control flow throughout
        const int ow = g % NMS_WARPS, oj = g / NMS_WARPS;
        if (warp == ow) {
#pragma unroll
            for (int j = 0; j < COOP_R; ++j)
                if (j == oj) {
                    uint32_t al = alive_w[j], todo = al;
                    while (todo) {                                    // warp-uniform
                        const int bit = __ffs(todo) - 1;
                        todo &= todo - 1u;
                        const float4 bi = box[g * 32 + bit];
                        const bool kill = lane > bit && ((al >> lane) & 1u) && iou_hit(bi, box_area(bi), bj[j], aj[j], a.thr_iou, thr_lo, quick);
                        const uint32_t km = __ballot_sync(0xffffffffu, kill);
                        al &= ~km;
                        todo &= ~km;
                    }
                    alive_w[j] = al;
                    if ((al >> lane) & 1u) list[kept + __popc(al & ((1u << lane) - 1u))] = (uint16_t)(0xffff - (int)(key[g * 32 + lane] & 0xffffu));
                    if (lane == 0) alive_sm[g] = al;
                }
        }
        __syncthreads();
        const uint32_t surv = alive_sm[g];
        kept += __popc(surv);
#pragma unroll
        for (int j = 0; j < COOP_R; ++j) {
            const int wd = warp + NMS_WARPS * j;
            if (wd > g && wd < words && alive_w[j]) {                 // warp-uniform
                bool dead = !((alive_w[j] >> lane) & 1u);
                for (uint32_t sv = surv; sv; sv &= sv - 1u) {
                    const float4 bi = box[g * 32 + __ffs(sv) - 1];
                    if (!dead && iou_hit(bi, box_area(bi), bj[j], aj[j], a.thr_iou, thr_lo, quick)) dead = true;
                }
                alive_w[j] &= ~__ballot_sync(0xffffffffu, dead);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < COOP_R; ++j) {                                // suppressed candidates = cleared alive bits
        const int p = (warp + NMS_WARPS * j) * 32 + lane;
        if (p < K && !((alive_w[j] >> lane) & 1u)) {
            const int jx = 0xffff - (int)(key[p] & 0xffffu);
            atomicOr(&supp[jx >> 5], 1u << (jx & 31));
        }
    }
    __syncthreads();
    uint16_t* kept_out = a.cand + ((size_t)b * a.C + c) * a.N;
    for (int p = t; p < kept; p += NT) kept_out[p] = list[p];
    uint32_t* supp_out = a.supp + ((size_t)b * a.C + c) * a.W;
    for (int w = t; w < a.W; w += NT) supp_out[w] = supp[w];
    if (t == 0) a.kept_cnt[(size_t)b * a.C + c] = kept | (kept < K ? KEPT_HAS_SUPP : 0);
    __syncthreads();
}

// dynamic shared memory of nms_select_kernel: per warp {key[SEL_CAP] u64, box[SEL_CAP] float4, supp[Wp] u32, alive[Wp] u32},
// then list[32][SEL_CAP] u16 and cnt[33]; Wp = W rounded up to a multiple of 4
//   key  : (ford(value) << 32) | (0xffff - index): descending key = visiting order up to ties
//   box  : boxes in visiting order (also the scratch of the tie fix)
//   supp : suppressed candidates, by box index;  alive : alive candidates, by rank (general path: up to N ranks)
//   list : per class: candidates as found, later the kept list
__host__ __device__ inline size_t sel_warp_bytes(int W) { return (size_t)SEL_CAP * 24 + 2 * (size_t)((W + 3) & ~3) * 4; }
static size_t sel_smem_bytes(int W) {
    const size_t warps = NMS_WARPS * sel_warp_bytes(W);              // phase B (select_class_coop) re-uses this area as one pool
    return (warps > coop_pool_bytes(W) ? warps : coop_pool_bytes(W)) + 32 * SEL_CAP * sizeof(uint16_t) + 33 * sizeof(int);
}

// Boxes of one image in ascending order of quantised area: a stable counting sort on the top 11 bits of ford(area) (sign,
// exponent, two mantissa bits: buckets 19 % wide), one CTA per image.  apply takes its 128-box tiles in THIS order (see there);
// the order only has to be roughly by area, and it is deterministic (boxes of one bucket stay in index order).  It runs as one
// extra CTA per image inside the select launch (blockIdx.x == gridDim.x - 1), on that kernel's dynamic shared memory.
static constexpr int AREA_BUCKETS = 2048;
static constexpr size_t AREA_ORDER_SMEM = AREA_BUCKETS * sizeof(int) + NMS_MAX_N * sizeof(uint16_t) + 8 * sizeof(int);
static_assert(NMS_WARPS == 8 && AREA_BUCKETS == 8 * NMS_WARPS * 32, "area_order: 8 buckets per thread, 8 warp totals");
static_assert((size_t)NMS_WARPS * SEL_CAP * 24 >= AREA_ORDER_SMEM, "area_order lives in the select kernel's shared memory");
__device__ void area_order(const NmsArgs& a, int b, unsigned char* smem) {
    int* start = reinterpret_cast<int*>(smem);                           // [AREA_BUCKETS]
    uint16_t* bkt = reinterpret_cast<uint16_t*>(smem + AREA_BUCKETS * sizeof(int));      // [N]
    int* wsum = reinterpret_cast<int*>(smem + AREA_BUCKETS * sizeof(int) + NMS_MAX_N * sizeof(uint16_t));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* bmin = a.xy_min + (size_t)b * a.N * 2;
    const float* bmax = a.xy_max + (size_t)b * a.N * 2;
    for (int i = threadIdx.x; i < AREA_BUCKETS; i += 256) start[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < a.N; i += 256) {
        const uint32_t k = ford(box_area(load_box(bmin, bmax, i))) >> 21;
        bkt[i] = (uint16_t)k;
        atomicAdd(&start[k], 1);
    }
    __syncthreads();
    // exclusive prefix sum of the histogram: 8 buckets per thread, warp scan, 8 warp totals
    int loc[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { loc[j] = tot; tot += start[threadIdx.x * 8 + j]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = incl - tot;
    for (int w = 0; w < warp; ++w) base += wsum[w];
#pragma unroll
    for (int j = 0; j < 8; ++j) start[threadIdx.x * 8 + j] = base + loc[j];
    __syncthreads();
    if (warp != 0) return;
    uint16_t* out = a.area_perm + (size_t)b * a.N;
    for (int i0 = 0; i0 < a.N; i0 += 32) {                              // one warp, in index order: stable
        const int i = i0 + lane;
        const uint32_t k = i < a.N ? (uint32_t)bkt[i] : 0xffffffffu;
        const uint32_t same = __match_any_sync(0xffffffffu, k);
        if (i < a.N) out[start[k] + __popc(same & ((1u << lane) - 1u))] = (uint16_t)i;
        __syncwarp();
        if (i < a.N && (same & ((1u << lane) - 1u)) == 0u) start[k] += __popc(same);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(NMS_WARPS * 32) nms_select_kernel(NmsArgs a, int vec4) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    if (blockIdx.x == gridDim.x - 1) { area_order(a, blockIdx.y, sel_smem); return; }      // block-uniform
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = (a.W + 3) & ~3;
    unsigned char* wbase = sel_smem + warp * sel_warp_bytes(a.W);
    unsigned long long* w_key = reinterpret_cast<unsigned long long*>(wbase);
    float4* w_box = reinterpret_cast<float4*>(wbase + SEL_CAP * 8);
    uint32_t* w_supp = reinterpret_cast<uint32_t*>(wbase + SEL_CAP * 24);
    uint32_t* w_alive = w_supp + Wp;
    const size_t area = NMS_WARPS * sel_warp_bytes(a.W) > coop_pool_bytes(a.W) ? NMS_WARPS * sel_warp_bytes(a.W) : coop_pool_bytes(a.W);
    uint16_t (*s_list)[SEL_CAP] = reinterpret_cast<uint16_t (*)[SEL_CAP]>(sel_smem + area);
    int* s_cnt = reinterpret_cast<int*>(sel_smem + area + 32 * SEL_CAP * sizeof(uint16_t));
    const int cg = a.cg;
    const int c0 = blockIdx.x * cg;                     // this CTA: cg (32 | 16 | 8) classes of image b
    const int c_end = min(c0 + cg, a.C);
    const int b = blockIdx.y;
    const float* conf_img = a.conf + (size_t)b * a.N * a.C;
    const float* bmin = a.xy_min + (size_t)b * a.N * 2;
    const float* bmax = a.xy_max + (size_t)b * a.N * 2;

    // 1. compact candidates with COALESCED reads into the per-class shared-memory lists (arbitrary order: step 2 sorts
    //    with a total order).  Counters keep counting past SEL_CAP: such classes take the general path.
    if (threadIdx.x <= 32) s_cnt[threadIdx.x] = 0;        // [32] = next class to hand out
    __syncthreads();
    if (vec4) {                                          // C % 4 == 0, 16-byte aligned rows: lane = (box, 4 classes)
        const int lpr = cg >> 2, rows = 32 / lpr;        // lanes per row (8 | 4 | 2), rows per warp-wide load (4 | 8 | 16)
        const int q = lane % lpr, cl = 4 * q;
        if (c0 + cl < c_end) {
            // four rows of loads in flight per lane (the shared-memory atomics below would otherwise serialise the loads)
            for (int n0 = warp * rows + lane / lpr; n0 < a.N; n0 += NMS_WARPS * rows * 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n0 + u * NMS_WARPS * rows;
                    v[u] = n < a.N ? __ldg(reinterpret_cast<const float4*>(conf_img + (size_t)n * a.C + c0 + cl)) : make_float4(a.thr, a.thr, a.thr, a.thr);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n0 + u * NMS_WARPS * rows;
                    const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (vv[t] > a.thr) {                      // rows past N carry thr itself: never a candidate
                            const int slot = atomicAdd(&s_cnt[cl + t], 1);
                            if (slot < SEL_CAP) s_list[cl + t][slot] = (uint16_t)n;
                        }
                }
            }
        }
    } else {                                             // lane = (box, class): a warp reads cg consecutive classes of 32 / cg boxes
        const int rows = 32 / cg, cl = lane % cg;
        if (c0 + cl < c_end)
            for (int n = warp * rows + lane / cg; n < a.N; n += NMS_WARPS * rows) {
                if (__ldg(conf_img + (size_t)n * a.C + c0 + cl) > a.thr) {
                    const int slot = atomicAdd(&s_cnt[cl], 1);
                    if (slot < SEL_CAP) s_list[cl][slot] = (uint16_t)n;
                }
            }
    }
    __syncthreads();
    // reference asserts fire as soon as one live box is compared with the rest: with any candidate in this CTA's classes
    // every box of the image is checked (once per CTA)
    if (a.status && a.N >= 2) {
        bool any = false;
        for (int cl = 0; cl < 32; ++cl) any |= s_cnt[cl] > 0;
        if (any) {
            bool bad = false;
            for (int n = threadIdx.x; n < a.N; n += NMS_WARPS * 32) {
                const float4 q = load_box(bmin, bmax, n);
                bad |= !(q.x <= q.z) || !(q.y <= q.w);        // also true for NaN
            }
            if (bad) atomicExch(a.status + b, 1);
        }
    }
    const bool quick = a.thr_iou > 0.0f;
    const float thr_lo = __fmul_rn(0.999f, a.thr_iou);
    // Few CTAs in flight (a detection batch, not a sweep over hundreds of images): the machine is idle anyway and the kernel's
    // time is its longest class -- classes with more than COOP_MIN candidates then get the whole CTA (select_class_coop).
    const bool coop = a.coop != 0;
    // one warp per class from here on; classes are handed out dynamically (the candidate counts of real score matrices
    // are very uneven: a few classes hold most of an image's candidates)
    for (;;) {
        int cl = 0;
        if (lane == 0) cl = atomicAdd(&s_cnt[32], 1);
        cl = __shfl_sync(0xffffffffu, cl, 0);
        if (cl >= cg) break;
        const int c = c0 + cl;
        if (c >= a.C) continue;
        const int K = s_cnt[cl];
        if (K == 0) {
            if (lane == 0) a.kept_cnt[(size_t)b * a.C + c] = 0;
            continue;
        }
        if (coop && K > COOP_MIN && K <= COOP_CAP) continue;          // phase B below
        if (K > SEL_CAP) {
            select_class_general(a, b, c, w_alive, w_supp, lane);
            continue;
        }
        // 2. keys -> shared memory, bitonic sort (descending), exact re-ranking of equal-value runs
        uint16_t* list = s_list[cl];
        int n2 = 32;
        while (n2 < K) n2 <<= 1;
        for (int p = lane; p < n2; p += 32) {
            unsigned long long key = 0ull;                            // padding sorts last (every real key is > 0)
            if (p < K) {
                const int idx = list[p];
                key = ((unsigned long long)ford(__ldg(conf_img + (size_t)idx * a.C + c)) << 32) | (unsigned long long)(0xffffu - idx);
            }
            w_key[p] = key;
        }
        for (int w = lane; w < a.W; w += 32) w_supp[w] = 0u;
        __syncwarp();
        for (int size = 2; size <= n2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = lane; t < (n2 >> 1); t += 32) {
                    const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));      // lower index of the t-th pair
                    const int hi = lo + stride;
                    const bool desc = (lo & size) == 0;
                    const unsigned long long x = w_key[lo], y = w_key[hi];
                    if ((x < y) == desc) { w_key[lo] = y; w_key[hi] = x; }
                }
                __syncwarp();
            }
        }
        if (c > 0) {                                                  // ties descend into earlier class columns
            bool tie = false;
            for (int p = lane; p < K; p += 32) {
                const uint32_t v = (uint32_t)(w_key[p] >> 32);
                tie |= (p > 0 && (uint32_t)(w_key[p - 1] >> 32) == v) || (p + 1 < K && (uint32_t)(w_key[p + 1] >> 32) == v);
            }
            if (__any_sync(0xffffffffu, tie)) {
                unsigned long long* tmp = reinterpret_cast<unsigned long long*>(w_box);
                for (int p = lane; p < K; p += 32) {
                    const unsigned long long kp = w_key[p];
                    const uint32_t v = (uint32_t)(kp >> 32);
                    int s = p, e = p + 1;
                    while (s > 0 && (uint32_t)(w_key[s - 1] >> 32) == v) --s;
                    while (e < K && (uint32_t)(w_key[e] >> 32) == v) ++e;
                    int rank = s;
                    if (e - s > 1) {
                        const int j = 0xffff - (int)(kp & 0xffffu);
                        for (int qq = s; qq < e; ++qq) {
                            if (qq == p) continue;
                            const int i = 0xffff - (int)(w_key[qq] & 0xffffu);
                            if (precedes(conf_img, a.C, c, i, 0.f, j, 0.f)) ++rank;       // equal values: earlier columns, then index
                        }
                    } else {
                        rank = p;
                    }
                    tmp[rank] = kp;
                }
                __syncwarp();
                for (int p = lane; p < K; p += 32) w_key[p] = tmp[p];
                __syncwarp();
            }
        }
        // 3. boxes in visiting order, greedy sweep (nms_sweep: alive mask and every lane's candidates in registers)
        for (int p = lane; p < K; p += 32) w_box[p] = load_box(bmin, bmax, 0xffff - (int)(w_key[p] & 0xffffu));
        __syncwarp();
        uint32_t alive = 0u;
        int kept;
        if (K <= 32) kept = nms_sweep<1>(w_box, w_key, list, K, a.thr_iou, thr_lo, quick, lane, alive);
        else if (K <= 64) kept = nms_sweep<2>(w_box, w_key, list, K, a.thr_iou, thr_lo, quick, lane, alive);
        else if (K <= 128) kept = nms_sweep<4>(w_box, w_key, list, K, a.thr_iou, thr_lo, quick, lane, alive);
        else kept = nms_sweep<8>(w_box, w_key, list, K, a.thr_iou, thr_lo, quick, lane, alive);
        // suppressed candidates = the ones whose alive bit was cleared
        for (int p0 = 0; p0 < K; p0 += 32) {
            const uint32_t aw = __shfl_sync(0xffffffffu, alive, p0 >> 5);
            const int p = p0 + lane;
            if (p < K && !((aw >> lane) & 1u)) {
                const int j = 0xffff - (int)(w_key[p] & 0xffffu);
                atomicOr(&w_supp[j >> 5], 1u << (j & 31));
            }
        }
        __syncwarp();
        uint16_t* kept_out = a.cand + ((size_t)b * a.C + c) * a.N;
        for (int p = lane; p < kept; p += 32) kept_out[p] = list[p];
        uint32_t* supp_out = a.supp + ((size_t)b * a.C + c) * a.W;
        for (int w = lane; w < a.W; w += 32) supp_out[w] = w_supp[w];
        if (lane == 0) a.kept_cnt[(size_t)b * a.C + c] = kept | (kept < K ? KEPT_HAS_SUPP : 0);
        __syncwarp();
    }   // class loop
    if (coop) {                                                       // phase B: the heavy classes, one at a time, all 8 warps
        __syncthreads();
        for (int cl = 0; cl < cg; ++cl) {
            const int c = c0 + cl;
            if (c >= a.C) break;
            const int K = s_cnt[cl];
            if (K > COOP_MIN && K <= COOP_CAP) select_class_coop(a, b, c, K, sel_smem, K <= SEL_CAP ? s_list[cl] : nullptr, quick, thr_lo);
        }
    }
}

// Final permutation: rank sort of all N boxes under the class-(C-1) visiting order. One CTA per image.
__global__ void __launch_bounds__(256) nms_order_kernel(NmsArgs a) {
    extern __shared__ float last_col[];          // [N]
    const int b = blockIdx.x;
    const float* conf_img = a.conf + (size_t)b * a.N * a.C;
    const int c = a.C - 1;
    for (int n = threadIdx.x; n < a.N; n += blockDim.x) last_col[n] = __ldg(conf_img + (size_t)n * a.C + c);
    __syncthreads();
    for (int j = threadIdx.x; j < a.N; j += blockDim.x) {
        const float vj = last_col[j];
        int rank = 0;
        for (int i = 0; i < a.N; ++i)
            if (i != j && precedes(conf_img, a.C, c, i, last_col[i], j, vj)) ++rank;
        a.order_out[(size_t)b * a.N + rank] = j;
    }
}

// apply: a CTA owns a [128 boxes][32 classes] tile of one image.  The tile is loaded with coalesced 128-bit loads into
// shared memory; each warp then walks class COLUMNS (lanes = boxes, 4 per lane), so the kept list of the class is uniform
// across the warp and its boxes are broadcast from shared memory -- the IoU loop has no global loads and no divergence.
// Candidates look their fate up in the suppressed mask select wrote.  Only tiles of classes that kept something are
// touched; modified tiles are written back coalesced.
// The 128 boxes of a tile are consecutive in the image's AREA order (area_perm), not in index order -- every row is a
// 128-byte segment of its own either way.  The 32 boxes a warp holds in register slot h then have areas in a narrow
// range [amin_h, amax_h], and since inter <= min(area) and den >= max(area) (1 - 2^-22),
//     IoU(k, box) >= thr  needs  0.998 thr amin_h <= area_k <= amax_h / (0.998 thr):
// two warp-uniform compares rule a kept box out for the whole slot (anchor boxes span three decades of area, so most
// (kept box, slot) pairs go this way).  Conservative: it only ever drops tests whose answer is "no hit".
static constexpr int AP_BOXES = 128, AP_CLASSES = 32, AP_H = AP_BOXES / 32, AP_PF = 8;
static_assert(AP_CLASSES * AP_PF == 256 && AP_CLASSES + AP_BOXES <= 256, "apply's thread roles assume 256 threads");
__global__ void __launch_bounds__(256, 4) nms_apply_kernel(NmsArgs a, int vec4) {
    __shared__ float tile[AP_BOXES][AP_CLASSES + 1];                     // +1: column walks are bank-conflict-free
    __shared__ float4 kbox[8][32];
    __shared__ float2 karea[8][32];                                      // {area, slot mask}
    __shared__ float s_lo[AP_H], s_hi[AP_H];
    __shared__ float4 pf_box[AP_CLASSES][AP_PF];                         // the first kept boxes of every class column, fetched with the tile
    __shared__ int tile_cnt[AP_CLASSES];
    __shared__ int item_start[AP_CLASSES + 1];
    __shared__ int any_kept, dirty, next_item;
    __shared__ uint32_t supp_cols;                                       // class columns with suppressed candidates (phase 1)
    __shared__ float pf_area[AP_PF][AP_CLASSES];
    __shared__ uint16_t s_perm[AP_BOXES];
    __shared__ float4 s_box[AP_BOXES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_tiles = (a.C + AP_CLASSES - 1) / AP_CLASSES;
    const int n_tiles = (a.N + AP_BOXES - 1) / AP_BOXES;
    const int total_tiles = a.B * n_tiles * c_tiles;                     // < 2^31: checked by nms_launch
    const bool quick = a.thr_iou > 0.0f;
    const float thr_lo = __fmul_rn(0.999f, a.thr_iou);
    const float thr_cull = __fmul_rn(0.998f, a.thr_iou);
    // The CTA walks tiles blockIdx.x, + gridDim.x, ... of [B][n_tiles][c_tiles]; the step is decomposed once so that the walk needs no
    // division, and the coordinates of the NEXT tile are known one tile ahead: its kept counts (threads 0..31) and its slice of
    // area_perm (threads 32..159) are fetched while this tile is worked on.
    const int d_ct = (int)gridDim.x % c_tiles, d_nt = ((int)gridDim.x / c_tiles) % n_tiles, d_b = (int)gridDim.x / c_tiles / n_tiles;
    int ct = (int)blockIdx.x % c_tiles, ntile = ((int)blockIdx.x / c_tiles) % n_tiles, b = (int)blockIdx.x / c_tiles / n_tiles;
    int ct2 = 0, ntile2 = 0, b2 = 0;
    auto perm_at = [&](int bb, int nt) -> uint16_t {
        const int r = nt * AP_BOXES + (int)threadIdx.x - AP_CLASSES;
        return (bb < a.B && r < a.N) ? __ldg(a.area_perm + (size_t)bb * a.N + r) : (uint16_t)0;
    };
    auto cnt_at = [&](int bb, int c_tile) -> int {
        const int c = c_tile * AP_CLASSES + (int)threadIdx.x;
        return (bb < a.B && c < a.C) ? __ldg(a.kept_cnt + (size_t)bb * a.C + c) : 0;
    };
    const bool perm_thread = threadIdx.x >= AP_CLASSES && threadIdx.x < AP_CLASSES + AP_BOXES;
    uint16_t perm_next = perm_thread ? perm_at(b, ntile) : (uint16_t)0;
    int cnt_next = threadIdx.x < AP_CLASSES ? cnt_at(b, ct) : 0;
    for (; b < a.B; b = b2, ntile = ntile2, ct = ct2) {
        ct2 = ct + d_ct; ntile2 = ntile + d_nt; b2 = b + d_b;
        if (ct2 >= c_tiles) { ct2 -= c_tiles; ++ntile2; }
        if (ntile2 >= n_tiles) { ntile2 -= n_tiles; ++b2; }
        const int c0 = ct * AP_CLASSES, n0 = ntile * AP_BOXES;
        if (threadIdx.x == 0) { any_kept = 0; dirty = 0; }
        __syncthreads();
        if (threadIdx.x < AP_CLASSES) {
            const int k = cnt_next & ~KEPT_HAS_SUPP;
            const uint32_t has_supp = __ballot_sync(0xffffffffu, cnt_next & KEPT_HAS_SUPP);      // threads 0..31 = warp 0
            cnt_next = cnt_at(b2, ct2);
            tile_cnt[threadIdx.x] = k;
            if (k) any_kept = 1;
            if (lane == 0) supp_cols = has_supp;
        } else if (perm_thread) {
            s_perm[threadIdx.x - AP_CLASSES] = perm_next;
            perm_next = perm_at(b2, ntile2);
        }
        __syncthreads();
        const int ak = any_kept;
        __syncthreads();                                             // everyone has read it before the next tile resets it
        if (!ak) continue;                                           // block-uniform
        float* conf_img = a.conf + (size_t)b * a.N * a.C;
        const float* bmin = a.xy_min + (size_t)b * a.N * 2;
        const float* bmax = a.xy_max + (size_t)b * a.N * 2;
        // thread = (class column, slot): index of the column's slot-th kept box now, its coordinates after the tile's loads are issued
        const int pf_cl = threadIdx.x >> 3, pf_slot = threadIdx.x & (AP_PF - 1);
        const bool pf_on = pf_slot < tile_cnt[pf_cl];
        const int pf_idx = pf_on ? __ldg(a.cand + ((size_t)b * a.C + c0 + pf_cl) * a.N + pf_slot) : 0;
        if (threadIdx.x < AP_BOXES)
            s_box[threadIdx.x] = n0 + threadIdx.x < a.N ? load_box(bmin, bmax, s_perm[threadIdx.x]) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec4) {                                                  // thread = (box row, 4 classes): 8 threads cover a 128-byte row
            const int q = (threadIdx.x & 7) * 4;
            for (int r = threadIdx.x >> 3; r < AP_BOXES; r += 32) {
                const int n = s_perm[r];
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + r < a.N && c0 + q < a.C) v = *reinterpret_cast<const float4*>(conf_img + (size_t)n * a.C + c0 + q);
                tile[r][q] = v.x; tile[r][q + 1] = v.y; tile[r][q + 2] = v.z; tile[r][q + 3] = v.w;
            }
        } else {
            for (int r = warp; r < AP_BOXES; r += 8) {
                const int n = s_perm[r], c = c0 + lane;
                tile[r][lane] = (n0 + r < a.N && c < a.C) ? conf_img[(size_t)n * a.C + c] : 0.f;
            }
        }
        if (pf_on) {
            const float4 q4 = load_box(bmin, bmax, pf_idx);
            pf_box[pf_cl][pf_slot] = q4;
            pf_area[pf_slot][pf_cl] = box_area(q4);
        }
        __syncthreads();
        float4 bn[AP_H];
        float barea[AP_H];
        bool n_ok[AP_H];
#pragma unroll
        for (int h = 0; h < AP_H; ++h) {
            n_ok[h] = n0 + h * 32 + lane < a.N;
            bn[h] = s_box[h * 32 + lane];
            barea[h] = box_area(bn[h]);
        }
        bool wrote = false;
        // phase 1: candidates look their fate up in the suppressed mask (a warp per class column)
        for (int cl = warp; cl < AP_CLASSES; cl += 8) {
            if (!((supp_cols >> cl) & 1u)) continue;                 // warp-uniform: no candidate of this class was suppressed
            const uint32_t* supp = a.supp + ((size_t)b * a.C + c0 + cl) * a.W;
#pragma unroll
            for (int h = 0; h < AP_H; ++h) {
                const int n = s_perm[h * 32 + lane];
                if (n_ok[h] && tile[h * 32 + lane][cl] > a.thr && ((__ldg(supp + (n >> 5)) >> (n & 31)) & 1u)) { tile[h * 32 + lane][cl] = 0.0f; wrote = true; }
            }
        }
        // work items of phase 2 = (class column, chunk of <= 32 kept boxes): the kept counts of real score matrices are very
        // uneven (one class may keep 160 boxes, most keep none), so the chunks -- not the columns -- are handed out to the warps
        // every tile has a CTA of its own (a detection batch): the kernel's time is its slowest warp -> items = chunks of 32 kept
        // boxes, handed out dynamically; many tiles per CTA (a sweep over hundreds of images): throughput regime -> one item per
        // class column (a lane that was hit skips the remaining chunks of its class), static round-robin
        const bool dynamic = a.apply_mode ? a.apply_mode == 1 : total_tiles <= (int)gridDim.x;
        if (warp == 0) {
            // per register slot: the window of kept-box areas that can reach thr_iou against its 32 boxes (all areas when thr_iou <= 0)
            float lo_t = __int_as_float(0x7f800000), hi_t = __int_as_float(0xff800000);
#pragma unroll
            for (int h = 0; h < AP_H; ++h) {
                float mn = n_ok[h] ? barea[h] : __int_as_float(0x7f800000), mx = n_ok[h] ? barea[h] : __int_as_float(0xff800000);
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                const float lo = quick ? __fmul_rn(thr_cull, mn) : __int_as_float(0xff800000);
                const float hi = quick ? __fdiv_rn(mx, thr_cull) : __int_as_float(0x7f800000);
                if (lane == 0) { s_lo[h] = lo; s_hi[h] = hi; }
                lo_t = fminf(lo_t, lo);
                hi_t = fmaxf(hi_t, hi);
            }
            // a class column is worked on only if one of its kept boxes lies in the window of some slot (columns with up to AP_PF
            // kept boxes: decided here from the prefetched areas; the others: when their chunks are staged)
            const int cnt = tile_cnt[lane];
            bool rel = cnt > AP_PF;
            for (int j = 0; j < AP_PF; ++j)
                if (j < cnt) { const float ar = pf_area[j][lane]; rel |= ar >= lo_t && ar <= hi_t; }
            const int chunks = !rel ? 0 : dynamic ? (cnt + 31) >> 5 : 1;
            int incl = chunks;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            item_start[lane] = incl - chunks;
            if (lane == 31) { item_start[32] = incl; next_item = 0; }
        }
        __syncthreads();                                             // phase 1's zeroes are in place, the item table is built
        const int items = item_start[32];
        for (int round = 0;; ++round) {
            int it = warp + 8 * round;
            if (dynamic) {
                if (lane == 0) it = atomicAdd(&next_item, 1);
                it = __shfl_sync(0xffffffffu, it, 0);
            }
            if (it >= items) break;
            const int cl = __popc(__ballot_sync(0xffffffffu, item_start[lane] <= it)) - 1;      // last column that starts at or before it
            const int kb = dynamic ? (it - item_start[cl]) << 5 : 0;                   // the item's range of kept boxes [kb, ke)
            const int ke = dynamic ? min(kb + 32, tile_cnt[cl]) : tile_cnt[cl];
            const uint16_t* kept = a.cand + ((size_t)b * a.C + c0 + cl) * a.N;
            // Non-candidates only: a kept candidate (> threshold) is never touched, a suppressed one is already zero.  Other warps
            // may zero entries of this column while we run: the races are benign (every write is 0.0f, a stale read only costs
            // a redundant test).
            bool skip[AP_H], hit[AP_H];
            bool all_skip = true;
#pragma unroll
            for (int h = 0; h < AP_H; ++h) {
                const float v = tile[h * 32 + lane][cl];
                skip[h] = !n_ok[h] || v > a.thr || __float_as_uint(v) == 0u;      // candidates, and scores that are +0.0 already (-0.0 must become +0.0)
                hit[h] = false;
                all_skip &= skip[h];
            }
            if (__all_sync(0xffffffffu, all_skip)) continue;
            for (int k0 = kb; k0 < ke; k0 += 32) {
                // stage the chunk's kept boxes in shared memory -- only those whose area lies in the window of some slot
                int m = 0;
                float4 kq0 = make_float4(0.f, 0.f, 0.f, 0.f);
                float ka0 = 0.f;
                if (k0 + lane < ke) {
                    kq0 = k0 + lane < AP_PF ? pf_box[cl][k0 + lane] : load_box(bmin, bmax, kept[k0 + lane]);
                    ka0 = box_area(kq0);
#pragma unroll
                    for (int h = 0; h < AP_H; ++h) m |= (ka0 >= s_lo[h] && ka0 <= s_hi[h]) ? 1 << h : 0;
                }
                const uint32_t live = __ballot_sync(0xffffffffu, m != 0);
                if (m) {
                    const int pos = __popc(live & ((1u << lane) - 1u));
                    kbox[warp][pos] = kq0;
                    karea[warp][pos] = make_float2(ka0, __int_as_float(m));
                }
                const int kc = __popc(live);
                __syncwarp();
                for (int t = 0; t < kc; ++t) {
                    const float4 kq = kbox[warp][t];
                    const float2 am = karea[warp][t];
                    const float ka = am.x;
                    const int mt = __float_as_int(am.y);
#pragma unroll
                    for (int h = 0; h < AP_H; ++h)
                        if (((mt >> h) & 1) && !skip[h] && !hit[h]) hit[h] = iou_hit(kq, ka, bn[h], barea[h], a.thr_iou, thr_lo, quick);
                }
                __syncwarp();
            }
#pragma unroll
            for (int h = 0; h < AP_H; ++h)
                if (hit[h]) { tile[h * 32 + lane][cl] = 0.0f; wrote = true; }
        }
        if (wrote) dirty = 1;
        __syncthreads();
        if (dirty) {
            if (vec4) {
                const int q = (threadIdx.x & 7) * 4;
                for (int r = threadIdx.x >> 3; r < AP_BOXES; r += 32) {
                    const int n = s_perm[r];
                    if (n0 + r < a.N && c0 + q < a.C)
                        *reinterpret_cast<float4*>(conf_img + (size_t)n * a.C + c0 + q) = make_float4(tile[r][q], tile[r][q + 1], tile[r][q + 2], tile[r][q + 3]);
                }
            } else {
                for (int r = warp; r < AP_BOXES; r += 8) {
                    const int n = s_perm[r], c = c0 + lane;
                    if (n0 + r < a.N && c < a.C) conf_img[(size_t)n * a.C + c] = tile[r][lane];
                }
            }
        }
        __syncthreads();
    }
}

static inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
size_t nms_workspace_bytes(int B, int N, int C) {
    const size_t lists = al16((size_t)B * C * N * sizeof(uint16_t));
    const size_t masks = al16((size_t)B * C * ((N + 31) / 32) * sizeof(uint32_t));
    return 2 * lists + masks + al16((size_t)B * C * sizeof(int)) + al16((size_t)B * N * sizeof(uint16_t));
}

int nms_launch(float* conf, const float* xy_min, const float* xy_max, int B, int N, int C, float threshold,
               float threshold_iou, int* order_out, int* status_out, void* ws, size_t ws_bytes, cudaStream_t s) {
    Y2_REQUIRE(conf && xy_min && xy_max && ws, "nms: null argument");
    Y2_REQUIRE(B >= 0 && N >= 0 && C >= 0, "nms: negative extent");
    Y2_REQUIRE(N <= NMS_MAX_N, "nms: at most %d boxes per image are supported (got %d)", NMS_MAX_N, N);
    Y2_REQUIRE(ws_bytes >= nms_workspace_bytes(B, N, C), "nms: workspace too small (%zu < %zu)", ws_bytes,
               nms_workspace_bytes(B, N, C));
    Y2_REQUIRE((reinterpret_cast<uintptr_t>(xy_min) & 7) == 0 && (reinterpret_cast<uintptr_t>(xy_max) & 7) == 0,
               "nms: box arrays must be 8-byte aligned");
    Y2_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "nms: workspace must be 16-byte aligned");
    if (status_out) Y2_CUDA(cudaMemsetAsync(status_out, 0, (size_t)B * sizeof(int), s));
    if (B == 0 || N == 0 || C == 0) return 0;
    Y2_REQUIRE(B <= 65535, "nms: batch too large for one launch");
    NmsArgs a;
    a.conf = conf; a.xy_min = xy_min; a.xy_max = xy_max; a.B = B; a.N = N; a.C = C; a.W = (N + 31) / 32;
    a.thr = threshold; a.thr_iou = threshold_iou;
    const size_t lists = al16((size_t)B * C * N * sizeof(uint16_t));
    const size_t masks = al16((size_t)B * C * a.W * sizeof(uint32_t));
    char* base = static_cast<char*>(ws);
    a.cand = reinterpret_cast<uint16_t*>(base);
    a.sorted = reinterpret_cast<uint16_t*>(base + lists);
    a.supp = reinterpret_cast<uint32_t*>(base + 2 * lists);
    a.kept_cnt = reinterpret_cast<int*>(base + 2 * lists + masks);
    a.area_perm = reinterpret_cast<uint16_t*>(base + 2 * lists + masks + al16((size_t)B * C * sizeof(int)));
    a.status = status_out;
    a.order_out = order_out;
    a.apply_mode = g_nms_apply_mode;
    const int vec4 = (C % 4 == 0) && (reinterpret_cast<uintptr_t>(conf) & 15) == 0;
    const size_t smem = sel_smem_bytes(a.W);
    static unsigned long long attr_seen = 0;
    if (first_use_on_current_device(attr_seen))
        Y2_CUDA(cudaFuncSetAttribute(nms_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem_bytes(NMS_MAX_N / 32)));
    // Latency regime (a detection batch): with 32 classes per CTA a batch of 32 images is 96 CTAs, and the kernel's time is the CTA that
    // happens to hold the image's three heaviest classes, one after the other.  Narrower class groups spread them (a CTA reads whole
    // 32-byte sectors either way); a sweep over hundreds of images keeps 32 (fewer, fuller CTAs).
    a.coop = (long long)B * ((C + 31) / 32) <= 2 * 148;
    a.cg = (g_nms_select_cg == 8 || g_nms_select_cg == 16 || g_nms_select_cg == 32) ? g_nms_select_cg : (long long)B * ((C + 7) / 8) <= 3 * 148 ? 8 : (long long)B * ((C + 15) / 16) <= 3 * 148 ? 16 : 32;
    dim3 grid((C + a.cg - 1) / a.cg + 1, B);                             // + one area_order CTA per image
    nms_select_kernel<<<grid, NMS_WARPS * 32, smem, s>>>(a, vec4);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    if (order_out) {
        nms_order_kernel<<<B, 256, (size_t)N * sizeof(float), s>>>(a);
        Y2_CUDA(cudaGetLastError());
        note_launch();
    }
    long long blocks = (long long)B * ((N + AP_BOXES - 1) / AP_BOXES) * ((C + AP_CLASSES - 1) / AP_CLASSES);
    Y2_REQUIRE(blocks < (1ll << 31) - 148 * 8, "nms: B x N x C too large for one launch");
    if (blocks > 148 * 8) blocks = 148 * 8;
    nms_apply_kernel<<<(int)blocks, 256, 0, s>>>(a, vec4);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
