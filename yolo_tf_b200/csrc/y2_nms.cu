// Greedy per-class IoU NMS, bit-exact with the reference's utils/postprocess.py:21-51
// (iou :21-36, non_max_suppress :39-51) as executed under NumPy >= 2 (float32 everywhere).
//
// Reference semantics reproduced exactly:
//   * class c visits boxes in the order of Python's stable descending sort carried across classes
//     = lexicographic (conf[:,c] desc, conf[:,c-1] desc, ..., conf[:,0] desc, index asc) on the
//     ORIGINAL scores;
//   * a box whose class score is > threshold when reached ("live") zeroes the class score of EVERY
//     later box - candidate or not - whose IoU with it is >= threshold_iou; zeroed boxes never act;
//   * IoU uses the float32 operation order of iou() with no FMA contraction (explicit _rn intrinsics).
//
// Three kernels, no host round trip:
//   select : a CTA owns 32 classes of one image.  Candidates (> threshold) are compacted with coalesced reads
//            (a warp reads 32 consecutive classes of one box = 128 bytes; lane = class; slots from a shared-memory
//            counter).  Then one warp per class rank-sorts its candidates with the lexicographic comparator (ties
//            descend into earlier class columns, still unmodified because nothing is written here) and runs the
//            sequential greedy sweep with a shared-memory alive bitmask -> list of KEPT boxes.
//   order  : (optional) final permutation of the N boxes = the order of the list the reference returns.
//   apply  : a CTA owns a [64 boxes][32 classes] tile (coalesced 128-byte rows through shared memory); warps walk
//            class columns with lanes = boxes, the class's kept boxes broadcast from shared memory.  A candidate is
//            zeroed iff it is not kept; a non-candidate (it sorts after every candidate) is zeroed iff any kept box
//            overlaps it >= threshold_iou (disjoint pairs are rejected before the divide when threshold_iou > 0).
// Scores are read once by select and once by apply, both coalesced; only tiles that change are written back.
#include "y2_internal.h"

namespace y2 {

static constexpr int NMS_WARPS = 8;
static constexpr int NMS_MAX_N = 8192;     // alive bitmask: 256 words per warp

__device__ __forceinline__ float iou_ref(float4 a, float4 b) {   // (xmin, ymin, xmax, ymax)
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float iw = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float ih = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float den = fmaxf(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-10f);
    return __fdiv_rn(inter, den);
}
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

struct NmsArgs {
    float* conf;
    const float* xy_min;
    const float* xy_max;
    int B, N, C;
    float thr, thr_iou;
    uint16_t* cand;      // [B][C][N] candidates in index order, overwritten with the kept list
    uint16_t* sorted;    // [B][C][N] scratch: candidates in visiting order
    int* kept_cnt;       // [B][C]
    int* status;         // [B] nullable: 1 = a reference assert (NaN / xy_min > xy_max) would fire
    int* order_out;      // [B][N] nullable
};

__global__ void __launch_bounds__(NMS_WARPS * 32) nms_select_kernel(NmsArgs a) {
    __shared__ uint32_t alive_sm[NMS_WARPS][NMS_MAX_N / 32];
    __shared__ int cnt_sm[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * 32;                     // this CTA: 32 classes of image b
    const int b = blockIdx.y;
    const float* conf_img = a.conf + (size_t)b * a.N * a.C;
    const float* bmin = a.xy_min + (size_t)b * a.N * 2;
    const float* bmax = a.xy_max + (size_t)b * a.N * 2;
    uint32_t* alive = alive_sm[warp];

    // 1. compact candidates with COALESCED reads: a warp reads 32 consecutive classes of one box (128 bytes);
    //    lane = class.  List order is arbitrary (slots from a shared-memory counter); step 2 sorts with a total order.
    if (threadIdx.x < 32) cnt_sm[threadIdx.x] = 0;
    __syncthreads();
    if (c0 + lane < a.C) {
        uint16_t* my_cand = a.cand + ((size_t)b * a.C + c0 + lane) * a.N;
        for (int n = warp; n < a.N; n += NMS_WARPS) {
            if (__ldg(conf_img + (size_t)n * a.C + c0 + lane) > a.thr) my_cand[atomicAdd(&cnt_sm[lane], 1)] = (uint16_t)n;
        }
    }
    __syncthreads();
    for (int cl = warp; cl < 32; cl += NMS_WARPS) {     // one warp per class from here on
    const int c = c0 + cl;
    if (c >= a.C) break;
    uint16_t* cand = a.cand + ((size_t)b * a.C + c) * a.N;
    uint16_t* sorted = a.sorted + ((size_t)b * a.C + c) * a.N;
    const int K = cnt_sm[cl];
    if (K == 0) {
        if (lane == 0) a.kept_cnt[(size_t)b * a.C + c] = 0;
        continue;
    }
    // reference asserts fire as soon as one live box is compared with the rest: every box is checked
    if (a.status && a.N >= 2) {
        bool bad = false;
        for (int n = lane; n < a.N; n += 32) {
            const float4 q = load_box(bmin, bmax, n);
            bad |= !(q.x <= q.z) || !(q.y <= q.w);        // also true for NaN
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(a.status + b, 1);
    }
    // 2. rank sort into visiting order
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
    __syncwarp();
    // 3. greedy sweep in visiting order
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
            if (jj > r && jj < K && ((alive[jj >> 5] >> (jj & 31)) & 1u)) {
                const float4 bj = load_box(bmin, bmax, sorted[jj]);
                kill = iou_ref(bi, bj) >= a.thr_iou;
            }
            const uint32_t km = __ballot_sync(0xffffffffu, kill);
            if (km && lane == 0) alive[j0 >> 5] &= ~km;
            __syncwarp();
        }
    }
    if (lane == 0) a.kept_cnt[(size_t)b * a.C + c] = kept;
    __syncwarp();
    }   // class loop
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

// apply: a CTA owns a [64 boxes][32 classes] tile of one image.  The tile is loaded with coalesced 128-byte rows into
// shared memory; each warp then walks class COLUMNS (lanes = boxes), so the kept list of the class is uniform across
// the warp and its boxes are broadcast from shared memory -- the IoU loop has no global loads and no divergence.
// Only tiles of classes that kept something are touched; modified tiles are written back coalesced.
static constexpr int AP_BOXES = 64, AP_CLASSES = 32;
__global__ void __launch_bounds__(256) nms_apply_kernel(NmsArgs a) {
    __shared__ float tile[AP_BOXES][AP_CLASSES + 1];
    __shared__ float4 kbox[8][32];
    __shared__ float karea[8][32];
    __shared__ int tile_cnt[AP_CLASSES];
    __shared__ int any_kept, dirty;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_tiles = (a.C + AP_CLASSES - 1) / AP_CLASSES;
    const int n_tiles = (a.N + AP_BOXES - 1) / AP_BOXES;
    const long long total_tiles = (long long)a.B * n_tiles * c_tiles;
    for (long long tix = blockIdx.x; tix < total_tiles; tix += gridDim.x) {
        const int ct = (int)(tix % c_tiles);
        const long long r0 = tix / c_tiles;
        const int ntile = (int)(r0 % n_tiles);
        const int b = (int)(r0 / n_tiles);
        const int c0 = ct * AP_CLASSES, n0 = ntile * AP_BOXES;
        if (threadIdx.x == 0) { any_kept = 0; dirty = 0; }
        __syncthreads();
        if (threadIdx.x < AP_CLASSES) {
            const int c = c0 + threadIdx.x;
            const int k = (c < a.C) ? __ldg(a.kept_cnt + (size_t)b * a.C + c) : 0;
            tile_cnt[threadIdx.x] = k;
            if (k) any_kept = 1;
        }
        __syncthreads();
        const int ak = any_kept;
        __syncthreads();                                             // everyone has read it before the next tile resets it
        if (!ak) continue;                                           // block-uniform
        float* conf_img = a.conf + (size_t)b * a.N * a.C;
        // load: row = box, 32 consecutive classes = one 128-byte segment per warp
        for (int r = warp; r < AP_BOXES; r += 8) {
            const int n = n0 + r, c = c0 + lane;
            tile[r][lane] = (n < a.N && c < a.C) ? conf_img[(size_t)n * a.C + c] : 0.f;
        }
        __syncthreads();
        const float* bmin = a.xy_min + (size_t)b * a.N * 2;
        const float* bmax = a.xy_max + (size_t)b * a.N * 2;
        float4 bn[2];
        bool n_ok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * 32 + lane;
            n_ok[h] = n < a.N;
            bn[h] = n_ok[h] ? load_box(bmin, bmax, n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float barea[2] = {__fmul_rn(__fsub_rn(bn[0].z, bn[0].x), __fsub_rn(bn[0].w, bn[0].y)),
                                __fmul_rn(__fsub_rn(bn[1].z, bn[1].x), __fsub_rn(bn[1].w, bn[1].y))};
        bool wrote = false;
        const bool quick = a.thr_iou > 0.0f;
        const float thr_lo = __fmul_rn(0.999f, a.thr_iou);
        for (int cl = warp; cl < AP_CLASSES; cl += 8) {
            const int cnt = tile_cnt[cl];
            if (cnt == 0) continue;                                  // warp-uniform
            const uint16_t* kept = a.cand + ((size_t)b * a.C + c0 + cl) * a.N;
            float v[2];
            bool cand[2], found[2] = {false, false}, hit[2] = {false, false};
#pragma unroll
            for (int h = 0; h < 2; ++h) { v[h] = tile[h * 32 + lane][cl]; cand[h] = v[h] > a.thr; }
            for (int k0 = 0; k0 < cnt; k0 += 32) {
                const int kc = min(32, cnt - k0);
                int kidx = -1;
                if (lane < kc) {
                    kidx = kept[k0 + lane];
                    const float4 kq = load_box(bmin, bmax, kidx);
                    kbox[warp][lane] = kq;
                    karea[warp][lane] = __fmul_rn(__fsub_rn(kq.z, kq.x), __fsub_rn(kq.w, kq.y));
                }
                __syncwarp();
                for (int t = 0; t < kc; ++t) {
                    const float4 kb = kbox[warp][t];
                    const float ka = karea[warp][t];
                    const int ki = __shfl_sync(0xffffffffu, kidx, t);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        found[h] |= (ki == n0 + h * 32 + lane);
                        if (cand[h] || hit[h]) continue;
                        if (quick) {
                            // Same float32 operations as iou_ref up to the divide (bit-identical inter and den), then a
                            // conservative filter: rn(inter/den) >= thr needs inter >= thr*(1-2^-24)*den, so anything
                            // below 0.999*thr*den is certainly no hit and skips the IEEE division.  Disjoint pairs
                            // (inter == 0) fall out here too.  NaNs fail the '<' and take the exact path.
                            const float iw = fmaxf(__fsub_rn(fminf(kb.z, bn[h].z), fmaxf(kb.x, bn[h].x)), 0.0f);
                            const float ih = fmaxf(__fsub_rn(fminf(kb.w, bn[h].w), fmaxf(kb.y, bn[h].y)), 0.0f);
                            const float inter = __fmul_rn(iw, ih);
                            const float den = fmaxf(__fsub_rn(__fadd_rn(ka, barea[h]), inter), 1e-10f);
                            if (inter < __fmul_rn(thr_lo, den)) continue;
                            hit[h] = __fdiv_rn(inter, den) >= a.thr_iou;
                        } else {
                            hit[h] = iou_ref(kb, bn[h]) >= a.thr_iou;
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool zero = n_ok[h] && (cand[h] ? !found[h] : hit[h]);
                if (zero) { tile[h * 32 + lane][cl] = 0.0f; wrote = true; }
            }
        }
        if (wrote) dirty = 1;
        __syncthreads();
        if (dirty) {
            for (int r = warp; r < AP_BOXES; r += 8) {
                const int n = n0 + r, c = c0 + lane;
                if (n < a.N && c < a.C) conf_img[(size_t)n * a.C + c] = tile[r][lane];
            }
        }
        __syncthreads();
    }
}

size_t nms_workspace_bytes(int B, int N, int C) {
    const size_t lists = (size_t)B * C * N * sizeof(uint16_t);
    const size_t a16 = (lists + 15) & ~(size_t)15;
    return 2 * a16 + (((size_t)B * C * sizeof(int)) + 15 & ~(size_t)15);
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
    if (status_out) Y2_CUDA(cudaMemsetAsync(status_out, 0, (size_t)B * sizeof(int), s));
    if (B == 0 || N == 0 || C == 0) return 0;
    Y2_REQUIRE(B <= 65535, "nms: batch too large for one launch");
    NmsArgs a;
    a.conf = conf; a.xy_min = xy_min; a.xy_max = xy_max; a.B = B; a.N = N; a.C = C;
    a.thr = threshold; a.thr_iou = threshold_iou;
    const size_t a16 = ((size_t)B * C * N * sizeof(uint16_t) + 15) & ~(size_t)15;
    a.cand = reinterpret_cast<uint16_t*>(ws);
    a.sorted = reinterpret_cast<uint16_t*>(static_cast<char*>(ws) + a16);
    a.kept_cnt = reinterpret_cast<int*>(static_cast<char*>(ws) + 2 * a16);
    a.status = status_out;
    a.order_out = order_out;
    dim3 grid((C + 31) / 32, B);
    nms_select_kernel<<<grid, NMS_WARPS * 32, 0, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    if (order_out) {
        nms_order_kernel<<<B, 256, (size_t)N * sizeof(float), s>>>(a);
        Y2_CUDA(cudaGetLastError());
    note_launch();
    }
    long long blocks = (long long)B * ((N + AP_BOXES - 1) / AP_BOXES) * ((C + AP_CLASSES - 1) / AP_CLASSES);
    if (blocks > 148 * 8) blocks = 148 * 8;
    nms_apply_kernel<<<(int)blocks, 256, 0, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
