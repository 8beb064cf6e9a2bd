// YOLOv2 head kernels (HBM-bound): anchor-box decode and the fused 4-part loss forward+backward.
//
//   decode  <- Model.__init__        model/yolo2/__init__.py:28-59  (+ calc_cell_xy, model/yolo/__init__.py:29-34)
//   loss    <- Objectives.__init__   model/yolo2/__init__.py:62-94  and the weighting in
//              Builder.create_objectives :114-119; backward is the closed form of d(total)/d(inputs).
//
// Layout: net [B, Hc, Wc, A*(5+C)] float32 viewed as inputs[B, cells, A, 5+C] (k: 0=iou 1=x 2=y 3=w 4=h 5..=class).
#include <algorithm>
#include "y2_internal.h"
#include "../../include/yolo2_b200.h"

namespace y2 {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// decode: a CTA stages G consecutive boxes (G*(5+C) contiguous floats) into shared memory with
// 128-bit loads, one THREAD decodes one box out of its private shared-memory row, results are staged
// and written back with 128-bit stores.
struct DecodeArgs {
    const float* net;
    const float* anchors;     // [A][2] float32
    long long boxes;          // B*cells*A
    int cells, A, C, Wc, Hc, G;
    y2_head_outputs o;
};

__device__ __forceinline__ void copy_out(float* __restrict__ dst, const float* __restrict__ src, int n, int tid,
                                         int nthreads) {
    // dst 16-byte aligned when n%4==0 chunks start on multiples of 4 floats (G%4==0 guarantees it)
    const int n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; i < n4; i += nthreads) d4[i] = s4[i];
    for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
}

__global__ void __launch_bounds__(128) decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int D = 5 + a.C;
    float* s_in = sm;                              // [G][D]
    const int CP = a.C + 4;                        // row pitch of s_conf: +4 words keeps the rows 16-byte aligned and spreads a warp's per-class stores over 8 banks instead of 2
    float* s_conf = s_in + (size_t)a.G * D;        // [G][CP]
    float* s_box = s_conf + (size_t)a.G * CP;      // [G][24]: xmin ymin xmax ymax | iou w h area | x y oxmin oymin | oxmax oymax sqw sqh | sx sy w01 h01 | (cx, cy, anchor as ints)
    for (long long g0 = (long long)blockIdx.x * a.G; g0 < a.boxes; g0 += (long long)gridDim.x * a.G) {
        const int g_cnt = (int)min((long long)a.G, a.boxes - g0);
        // ---- stage in (g0*D*4 bytes is a multiple of 16 because G%4==0)
        {
            const float4* src4 = reinterpret_cast<const float4*>(a.net + g0 * D);
            const int n = g_cnt * D, n4 = n >> 2;
            float4* d4 = reinterpret_cast<float4*>(s_in);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = __ldg(src4 + i);
            for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s_in[i] = __ldg(a.net + g0 * D + i);
        }
        __syncthreads();
        // ---- one THREAD per box (round 2; round 1 ran a warp per box): a box is 5 + C numbers, far too little to keep 32
        // lanes busy -- every shuffle, index computation and scalar box transform was executed 32-fold (444 warp
        // instructions per box).  The box's row is private in shared memory (row pitch 5 + C words: odd for the shipped class
        // counts, so the lanes of a warp hit distinct banks), the class exponentials are written back in place, so the softmax
        // costs one exp per class.  Hardware exp / reciprocal (ex2.approx, rcp.approx: a few ulp; the parity bar is 1e-4).
        if (threadIdx.x < g_cnt) {
            const int g = threadIdx.x;
            float* in = s_in + g * D;
            const long long gi = g0 + g;
            const int n = (int)(gi % ((long long)a.cells * a.A));
            const int cell = n / a.A, an = n - cell * a.A;
            // four independent chains per reduction: the thread's latency, not its instruction count, is what is left
            const int C4 = a.C & ~3;
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            for (int c = 0; c < C4; c += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], in[5 + c + u]);
            }
            for (int c = C4; c < a.C; ++c) m4[0] = fmaxf(m4[0], in[5 + c]);
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < C4; c += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float e = __expf(in[5 + c + u] - mx);
                    in[5 + c + u] = e;
                    s4[u] += e;
                }
            }
            for (int c = C4; c < a.C; ++c) {
                const float e = __expf(in[5 + c] - mx);
                in[5 + c] = e;
                s4[0] += e;
            }
            const float se = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            const float rse = __fdividef(1.0f, se);
            const float iou = __fdividef(1.0f, 1.0f + __expf(-in[0]));
            const float sx = __fdividef(1.0f, 1.0f + __expf(-in[1])), sy = __fdividef(1.0f, 1.0f + __expf(-in[2]));
            const float w = __expf(in[3]) * __ldg(a.anchors + 2 * an), h = __expf(in[4]) * __ldg(a.anchors + 2 * an + 1);
            float* cf = s_conf + g * CP;
            if (a.o.prob) {
                for (int c = 0; c < a.C; ++c) {
                    const float pr = in[5 + c] * rse;
                    cf[c] = iou * pr;
                    a.o.prob[gi * a.C + c] = pr;
                }
            } else {
                const float k = iou;
#pragma unroll 4
                for (int c = 0; c < a.C; ++c) cf[c] = k * (in[5 + c] * rse);
            }
            const float hw = w / 2.0f, hh = h / 2.0f;
            const float oxmin = sx - hw, oymin = sy - hh, oxmax = sx + hw, oymax = sy + hh;
            const float cx = (float)(cell % a.Wc), cy = (float)(cell / a.Wc);
            float4* bx = reinterpret_cast<float4*>(s_box + g * 24);
            bx[0] = make_float4(cx + oxmin, cy + oymin, cx + oxmax, cy + oymax);
            bx[1] = make_float4(iou, w, h, w * h);
            bx[2] = make_float4(cx + sx, cy + sy, oxmin, oymin);
            bx[3] = make_float4(oxmax, oymax, sqrtf(w / (float)a.Wc), sqrtf(h / (float)a.Hc));
            bx[4] = make_float4(sx, sy, w / (float)a.Wc, h / (float)a.Hc);
        }
        __syncthreads();
        // ---- stage out
        if (a.o.conf) {                              // rows of C floats (pitch CP) -> contiguous [g_cnt][C]
            float* dst = a.o.conf + g0 * a.C;
            if ((a.C & 3) == 0) {
                const int c4 = a.C >> 2, n4 = g_cnt * c4;
                for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                    const int r = i / c4, q = i - r * c4;
                    reinterpret_cast<float4*>(dst)[i] = *reinterpret_cast<const float4*>(s_conf + r * CP + 4 * q);
                }
            } else {
                for (int i = threadIdx.x; i < g_cnt * a.C; i += blockDim.x) {
                    const int r = i / a.C;
                    dst[i] = s_conf[r * CP + (i - r * a.C)];
                }
            }
        }
        for (int i = threadIdx.x; i < g_cnt; i += blockDim.x) {
            const float* bx = s_box + (size_t)i * 24;
            const float sx = bx[16], sy = bx[17];
            const long long gi = g0 + i;
            if (a.o.xy_min) *reinterpret_cast<float2*>(a.o.xy_min + gi * 2) = make_float2(bx[0], bx[1]);
            if (a.o.xy_max) *reinterpret_cast<float2*>(a.o.xy_max + gi * 2) = make_float2(bx[2], bx[3]);
            if (a.o.iou) a.o.iou[gi] = bx[4];
            if (a.o.wh) *reinterpret_cast<float2*>(a.o.wh + gi * 2) = make_float2(bx[5], bx[6]);
            if (a.o.areas) a.o.areas[gi] = bx[7];
            if (a.o.xy) *reinterpret_cast<float2*>(a.o.xy + gi * 2) = make_float2(bx[8], bx[9]);
            if (a.o.offset_xy) *reinterpret_cast<float2*>(a.o.offset_xy + gi * 2) = make_float2(sx, sy);
            if (a.o.offset_xy_min) *reinterpret_cast<float2*>(a.o.offset_xy_min + gi * 2) = make_float2(bx[10], bx[11]);
            if (a.o.offset_xy_max) *reinterpret_cast<float2*>(a.o.offset_xy_max + gi * 2) = make_float2(bx[12], bx[13]);
            if (a.o.coords) *reinterpret_cast<float4*>(a.o.coords + gi * 4) = make_float4(sx, sy, bx[14], bx[15]);
            if (a.o.wh01) *reinterpret_cast<float2*>(a.o.wh01 + gi * 2) = make_float2(bx[18], bx[19]);
        }
        __syncthreads();
    }
}

int head_decode_launch(const float* net, int B, int Hc, int Wc, int A, int C, const float* anchors,
                       const y2_head_outputs* outs, cudaStream_t s) {
    Y2_REQUIRE(net && anchors && outs, "head_decode: null argument");
    Y2_REQUIRE(B >= 0 && Hc > 0 && Wc > 0 && A > 0 && C > 0, "head_decode: bad shape");
    Y2_REQUIRE((reinterpret_cast<uintptr_t>(net) & 15) == 0, "head_decode: net must be 16-byte aligned");
    DecodeArgs a;
    a.net = net; a.anchors = anchors; a.cells = Hc * Wc; a.A = A; a.C = C; a.Wc = Wc; a.Hc = Hc;
    a.boxes = (long long)B * a.cells * A;
    a.o = *outs;
    if (a.boxes == 0) return 0;
    const int D = 5 + C;
    // boxes per CTA = threads that decode (the rest only help with the staging copies): 32 .. 128, a multiple of 4
    // (16-byte aligned chunks), enough CTAs to cover the machine at small batches (B = 32: 27 k boxes -> 423 CTAs of 64)
    int G = (int)std::min<long long>(128, std::max<long long>(32, (a.boxes / (148 * 2)) & ~31LL));
    while (G > 4 && (size_t)G * (D + C + 4 + 24) * 4 > 96 * 1024) G -= 4;
    Y2_REQUIRE((size_t)G * (D + C + 4 + 24) * 4 <= 200 * 1024, "head_decode: too many classes (%d)", C);
    a.G = G;
    const size_t smem = (size_t)G * (D + C + 4 + 24) * 4;
    static unsigned long long attr_seen = 0;
    if (first_use_on_current_device(attr_seen))
        Y2_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long blocks = (a.boxes + G - 1) / G;
    if (blocks > 148 * 16) blocks = 148 * 16;
    decode_kernel<<<(int)blocks, 128, smem, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// loss forward + backward, one warp per (image, cell).
struct LossArgs {
    const float* net;
    const float* anchors;
    const float *mask, *prob, *coords, *oxy_min, *oxy_max, *areas;   // labels
    float hp_prob, hp_iou_best, hp_iou_normal, hp_coords;
    float* dnet;             // nullable
    double* partials;        // [grid][4]
    long long ncells;        // B*cells
    int cells, A, C, Wc, Hc;
    float inv_cnt;           // 1 / (B*cells*A)
};

__global__ void __launch_bounds__(256) loss_kernel(LossArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int D = 5 + a.C;
    float s_ib = 0.f, s_in = 0.f, s_co = 0.f, s_pr = 0.f;    // per-warp running sums (lane-partial)
    for (long long ci = (long long)blockIdx.x * nw + warp; ci < a.ncells; ci += (long long)gridDim.x * nw) {
        const float* in = a.net + ci * a.A * D;
        float* gout = a.dnet ? a.dnet + ci * a.A * D : nullptr;
        const float m = __ldg(a.mask + ci);
        // lanes < A: box geometry for anchor = lane
        float sig0 = 0.f, sx = 0.f, sy = 0.f, sqw = 0.f, sqh = 0.f, iou = -INFINITY;
        if (lane < a.A) {
            const float* p = in + lane * D;
            sig0 = sigmoidf_(__ldg(p));
            sx = sigmoidf_(__ldg(p + 1));
            sy = sigmoidf_(__ldg(p + 2));
            const float w = expf(__ldg(p + 3)) * __ldg(a.anchors + 2 * lane);
            const float h = expf(__ldg(p + 4)) * __ldg(a.anchors + 2 * lane + 1);
            sqw = sqrtf(w / (float)a.Wc);
            sqh = sqrtf(h / (float)a.Hc);
            const float hw = w / 2.0f, hh = h / 2.0f;
            const float lox = fmaxf(sx - hw, __ldg(a.oxy_min + ci * 2)), loy = fmaxf(sy - hh, __ldg(a.oxy_min + ci * 2 + 1));
            const float hix = fminf(sx + hw, __ldg(a.oxy_max + ci * 2)), hiy = fminf(sy + hh, __ldg(a.oxy_max + ci * 2 + 1));
            const float iw = fmaxf(hix - lox, 0.f), ih = fmaxf(hiy - loy, 0.f);
            const float inter = iw * ih;
            const float uni = fmaxf(__ldg(a.areas + ci) + w * h - inter, 1e-10f);
            iou = inter / uni;
        }
        const float best = warp_max(iou);
        const float mb = (lane < a.A && iou == best) ? m : 0.f;       // mask_best = mask * (iou == max)
        if (lane < a.A) {
            const float mn = 1.f - mb;
            const float d0 = sig0 - mb;
            s_ib += mb * d0 * d0;
            s_in += mn * d0 * d0;
            const float tx = __ldg(a.coords + ci * 4), ty = __ldg(a.coords + ci * 4 + 1);
            const float tw = __ldg(a.coords + ci * 4 + 2), th = __ldg(a.coords + ci * 4 + 3);
            const float dx = sx - tx, dy = sy - ty, dw = sqw - tw, dh = sqh - th;
            s_co += mb * (dx * dx + dy * dy + dw * dw + dh * dh);
            if (gout) {
                float* g = gout + lane * D;
                const float w_o = (a.hp_iou_best * mb + a.hp_iou_normal * mn) * a.inv_cnt;
                g[0] = 2.f * d0 * sig0 * (1.f - sig0) * w_o;
                const float kc = a.hp_coords * mb * 2.f * a.inv_cnt;
                g[1] = kc * dx * sx * (1.f - sx);
                g[2] = kc * dy * sy * (1.f - sy);
                g[3] = kc * dw * 0.5f * sqw;
                g[4] = kc * dh * 0.5f * sqh;
            }
        }
        // class part, anchor by anchor; only anchors with mask_best != 0 contribute
        for (int an = 0; an < a.A; ++an) {
            const float mba = __shfl_sync(0xffffffffu, mb, an);
            const float* z = in + an * D + 5;
            float* g = gout ? gout + an * D + 5 : nullptr;
            if (mba == 0.f) {
                if (g)
                    for (int c = lane; c < a.C; c += 32) g[c] = 0.f;
                continue;
            }
            float mx = -INFINITY;
            for (int c = lane; c < a.C; c += 32) mx = fmaxf(mx, __ldg(z + c));
            mx = warp_max(mx);
            float se = 0.f;
            for (int c = lane; c < a.C; c += 32) se += expf(__ldg(z + c) - mx);
            se = warp_sum(se);
            float dist = 0.f, qp = 0.f;
            const float kq = a.hp_prob * mba * 2.f * a.inv_cnt;
            for (int c = lane; c < a.C; c += 32) {
                const float p = expf(__ldg(z + c) - mx) / se;
                const float d = p - __ldg(a.prob + ci * a.C + c);
                dist += d * d;
                qp += kq * d * p;
            }
            s_pr += mba * dist;
            if (g) {
                qp = warp_sum(qp);
                for (int c = lane; c < a.C; c += 32) {
                    const float p = expf(__ldg(z + c) - mx) / se;
                    const float d = p - __ldg(a.prob + ci * a.C + c);
                    g[c] = p * (kq * d - qp);
                }
            }
        }
    }
    // block reduce (double) -> partials[block][4]
    __shared__ double red[8][4];
    double v0 = s_ib, v1 = s_in, v2 = s_co, v3 = s_pr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        v3 += __shfl_xor_sync(0xffffffffu, v3, o);
    }
    if (lane == 0) { red[warp][0] = v0; red[warp][1] = v1; red[warp][2] = v2; red[warp][3] = v3; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// objectives[] order: prob, iou_best, iou_normal, coords (= hparam order of the C-ABI)
__global__ void loss_finish_kernel(const double* partials, int nblocks, double inv_cnt, float* objectives) {
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int b = 0; b < nblocks; ++b) t += partials[(size_t)b * 4 + threadIdx.x];   // fixed order: deterministic
        const int dst = threadIdx.x == 0 ? 1 : threadIdx.x == 1 ? 2 : threadIdx.x == 2 ? 3 : 0;   // ib,in,co,pr -> slots
        objectives[dst] = (float)(t * inv_cnt);
    }
}

static int loss_grid(long long ncells) {
    long long blocks = (ncells + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

size_t loss_workspace_bytes(int B, int Hc, int Wc) { return (size_t)loss_grid((long long)B * Hc * Wc) * 4 * sizeof(double); }

int loss_launch(const float* net, int B, int Hc, int Wc, int A, int C, const float* anchors, const float* mask,
                const float* prob, const float* coords, const float* oxy_min, const float* oxy_max,
                const float* areas, const float* hparam, float* objectives, float* dnet, void* ws, size_t ws_bytes,
                cudaStream_t s) {
    Y2_REQUIRE(net && anchors && mask && prob && coords && oxy_min && oxy_max && areas && hparam && objectives && ws,
               "loss: null argument");
    Y2_REQUIRE(A <= 32, "loss: at most 32 anchors per cell (got %d)", A);
    Y2_REQUIRE(B > 0 && Hc > 0 && Wc > 0 && C > 0, "loss: bad shape");
    LossArgs a;
    a.net = net; a.anchors = anchors; a.mask = mask; a.prob = prob; a.coords = coords;
    a.oxy_min = oxy_min; a.oxy_max = oxy_max; a.areas = areas;
    a.hp_prob = hparam[0]; a.hp_iou_best = hparam[1]; a.hp_iou_normal = hparam[2]; a.hp_coords = hparam[3];
    a.dnet = dnet; a.partials = static_cast<double*>(ws);
    a.cells = Hc * Wc; a.A = A; a.C = C; a.Wc = Wc; a.Hc = Hc;
    a.ncells = (long long)B * a.cells;
    const double cnt = (double)a.ncells * A;
    a.inv_cnt = (float)(1.0 / cnt);
    const int grid = loss_grid(a.ncells);
    Y2_REQUIRE(ws_bytes >= (size_t)grid * 4 * sizeof(double), "loss: workspace too small");
    loss_kernel<<<grid, 256, 0, s>>>(a);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    loss_finish_kernel<<<1, 32, 0, s>>>(a.partials, grid, 1.0 / cnt, objectives);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
