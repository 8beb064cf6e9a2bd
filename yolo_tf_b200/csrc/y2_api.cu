// C-ABI of libyolo2_b200.so (include/yolo2_b200.h): network handle, weight loading, forward plan.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/yolo2_b200.h"
#include "y2_internal.h"

namespace y2 {

// ---------------------------------------------------------------- errors
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* last_error() { return g_err.c_str(); }

int g_sched_override = 0;
double g_sched_handoff_kb = 13.0;
static float g_last_conv_ms = 0.f;
static unsigned long long* g_dbg_host = nullptr;
static int g_dbg_n = 0, g_dbg_sched[4] = {0, 0, 0, 0};
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int device_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    return n;
}

// declared in y2_head.cu / y2_nms.cu
int head_decode_launch(const float*, int, int, int, int, int, const float*, const y2_head_outputs*, cudaStream_t);
size_t loss_workspace_bytes(int, int, int);
int loss_launch(const float*, int, int, int, int, int, const float*, const float*, const float*, const float*,
                const float*, const float*, const float*, const float*, float*, float*, void*, size_t, cudaStream_t);
size_t nms_workspace_bytes(int, int, int);
int nms_launch(float*, const float*, const float*, int, int, int, float, float, int*, int*, void*, size_t,
               cudaStream_t);

// ---------------------------------------------------------------- network description
// Darknet-19 + passthrough, model/yolo2/inference.py:70-118; tiny, model/yolo2/inference.py:25-50.
struct LayerDesc {
    int ksize, cin, cout;   // logical = the variable's shape
    int pool;          // max-pool after: 1 = 2x2 stride 2 (inference.py:74,83,96 / :37), 2 = 2x2 stride 1 SAME (tiny, :42)
    int passthrough;   // tapped for reorg (:95)
    int has_bn;
    // stored channel counts: a multiple of 32 (the tcgen05 conv's K granule and the conv0 kernel's width); channels past
    // the logical count hold exact zeros (zero weights, zero scale/bias), so they add nothing to any sum
    int cin_s, cout_s;
    int to_concat;     // output lands in the concat buffer behind the reorg channels (conv19, :116)
    int from_concat;   // input is concat([reorg(passthrough), previous]) (conv20, :117)
};
static inline int store_channels(int c) { return c <= 3 ? c : (c + 31) / 32 * 32; }
static std::vector<LayerDesc> tiny_layers(int classes, int anchors) {
    std::vector<LayerDesc> L;
    int cin = 3, ch = 16;
    auto add = [&](int k, int cout, int pool) {
        L.push_back({k, cin, cout, pool, 0, 1, store_channels(cin), store_channels(cout), 0, 0});
        cin = cout;
    };
    for (int i = 0; i < 5; ++i) { add(3, ch, 1); ch *= 2; }     // :35-39
    add(3, ch, 2); ch *= 2;                                     // :40-44 (max_pool2d stride=1)
    add(3, ch, 0); add(3, ch, 0);                               // :45-47
    L.push_back({1, ch, anchors * (5 + classes), 0, 0, 0, ch, anchors * (5 + classes), 0, 0});   // :48
    return L;
}
static std::vector<LayerDesc> darknet19_layers(int classes, int anchors) {
    std::vector<LayerDesc> L;
    int cin = 3, ch = 32;
    auto add = [&](int k, int cout, int pool, int pt) {
        L.push_back({k, cin, cout, pool, pt, 1, cin, cout, 0, 0});
        cin = cout;
    };
    for (int i = 0; i < 2; ++i) { add(3, ch, 1, 0); ch *= 2; }
    for (int i = 0; i < 2; ++i) { add(3, ch, 0, 0); add(1, ch / 2, 0, 0); add(3, ch, 1, 0); ch *= 2; }
    add(3, ch, 0, 0); add(1, ch / 2, 0, 0); add(3, ch, 0, 0); add(1, ch / 2, 0, 0); add(3, ch, 1, 1);
    ch *= 2;
    add(3, ch, 0, 0); add(1, ch / 2, 0, 0); add(3, ch, 0, 0); add(1, ch / 2, 0, 0);
    add(3, ch, 0, 0); add(3, ch, 0, 0); add(3, ch, 0, 0);
    L.back().to_concat = 1;
    cin = 4 * 512 + ch;                      // concat([reorg(passthrough), net]) :115-116
    add(3, ch, 0, 0);
    L.back().from_concat = 1;
    L.push_back({1, ch, anchors * (5 + classes), 0, 0, 0, ch, anchors * (5 + classes), 0, 0});   // linear + bias :118
    return L;
}

struct LayerState {
    LayerDesc d;
    int cout_pad = 0, block_n = 0;
    bf16* wpack = nullptr;     // [2][cout_pad][k*k*cin] bf16 planes (inference; packed lazily, see ensure_pack)
    bf16* wpack16 = nullptr;   // the same as fp16 planes of w * 2^k (training forward; ensure_pack16)
    float* wsc = nullptr;      // [2 + cout_s]: 2^k, 2^-k, then the epilogue scale of the raw training conv (2^-k per channel)
    bool pack_fresh = false, pack16_fresh = false;
    float* w_f32 = nullptr;    // conv0 only (CUDA-core kernel reads HWIO fp32)
    float* scale = nullptr;    // [cout]  inference fold: gamma * rsqrt(moving_var + eps)
    float* bias = nullptr;     // [cout]  inference fold: beta - moving_mean * scale   (final layer: biases)
    // raw variables (training + re-folding): gamma, beta, moving_mean, moving_variance
    float *gamma = nullptr, *beta = nullptr, *mmean = nullptr, *mvar = nullptr;
    bf16* wpack_dgrad = nullptr;   // [2][cin_pad][k*k*cout_pad]: 180-degree rotated, channels swapped (lazy)
    int dg_cin_pad = 0, dg_cout_pad = 0, dg_block_n = 0;
    bool dgrad_fresh = false;
    bool loaded = false;
};

// training-step state inside the caller's workspace (valid between forward_train and backward)
struct TrainPlan {
    bool valid = false;
    int B = 0, H = 0, W = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    std::vector<int> oh, ow;
    std::vector<float*> z;          // raw conv outputs
    std::vector<bf16*> y, pooled;   // activations (planes); y[19] lives in the concat buffer
    std::vector<bf16*> y16, pooled16;   // the same activations as fp16 planes (f16 mode): what the forward convs read
    bf16* concat16 = nullptr;
    int f16 = 0;
    std::vector<float*> stat;       // per layer: mean, inv, scale, bias, m1, m2 (6 * cout floats)
    bf16* concat = nullptr;
    float *g0 = nullptr, *g1 = nullptr, *gcat = nullptr;
    bf16* dx = nullptr;
    double* red = nullptr;
    float* pad = nullptr;           // scratch for parameter gradients at a padded stored shape (tiny's conv0 / conv1)
    void* streamk = nullptr;
    const float* x = nullptr;
};

struct Plan {
    bool valid = false;
    int B = 0, H = 0, W = 0, precision = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    std::vector<bf16*> act;       // conv output planes per layer (layer 0: pooled output)
    std::vector<bf16*> pooled;    // pooled planes for pool layers (>0)
    std::vector<int> oh, ow;      // conv output spatial size per layer
    bf16* concat = nullptr;
    void* streamk = nullptr;
    std::vector<TcConvLaunch> launch;   // per layer (index 0 unused)
    std::vector<char> fused;            // per layer: max-pool fused into the conv epilogue
    bool reorg_fused = false;           // the passthrough layer's epilogue writes the concat buffer in reorg order itself
};

}  // namespace y2

using namespace y2;

struct y2_handle {
    int device = 0, classes = 0, anchors = 0, num_sms = 148;
    int arch = Y2_ARCH_DARKNET;        // Y2_ARCH_* (y2_create_net)
    int cat_c = 0;                     // channels of the concat buffer (0: the network has none)
    std::vector<LayerState> layers;
    Plan plan;
    TrainPlan tplan;
    int fuse_pool = 1;                 // y2_set_option("fuse_pool")
    int halo = 1;                      // y2_set_option("halo"): halo-tile mode for the 32-channel 3x3 layer (conv1)
    int conv0_tc = 2;                  // y2_set_option("conv0_tc"): conv0 on the tensor cores (SIMT-built im2col tile) instead of the CUDA cores; 2 = + unchecked gather for interior tiles
    int keep_activations = 0;          // y2_set_option("keep_activations"): 1 = every layer's output keeps its own workspace slot (y2_get_activation, tests); 0 = two alternating arenas
    int train_f16 = 1;                 // y2_set_option("train_f16"): training forward on fp16 planes (22 significand bits) with short accumulation chains; 0 = bf16 planes as in inference
    int train_kcap = 16;               // y2_set_option("train_kcap"): accumulation-chain cap of the training forward convs in k-blocks
    int pair = 1;                      // y2_set_option("pair"): CTA-pair (cta_group::2) convs: 0 off, 1 = 3x3 layers with 256-wide N tiles, 2 = every eligible layer
    int probe_layer = -1;              // test hooks (y2_train_probe)
    float *probe_gy = nullptr, *probe_gin = nullptr;
    bool profiling = false;
    std::vector<cudaEvent_t> ev;     // 2 per layer (start, stop) + 2 for the pool/reorg passes of that layer
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void choose_tiles(int cout, int* block_n, int* cout_pad) {
    const int p32 = (int)align_up((size_t)cout, 32);
    const int nt = (p32 + 255) / 256;
    const int bn = (int)align_up((size_t)((p32 + nt - 1) / nt), 32);
    *block_n = bn;
    *cout_pad = bn * nt;
}

extern "C" {

const char* y2_last_error(void) { return y2::last_error(); }
int y2_version(void) { return 100; }

int y2_create(y2_handle** out, int device, int classes, int num_anchors) {
    return y2_create_net(out, device, classes, num_anchors, Y2_ARCH_DARKNET);
}

int y2_create_net(y2_handle** out, int device, int classes, int num_anchors, int arch) {
    Y2_REQUIRE(out, "y2_create: null out");
    Y2_REQUIRE(arch == Y2_ARCH_DARKNET || arch == Y2_ARCH_TINY, "y2_create_net: unknown architecture %d", arch);
    Y2_REQUIRE(classes > 0 && num_anchors > 0, "y2_create: classes and num_anchors must be positive");
    int ndev = 0;
    Y2_CUDA(cudaGetDeviceCount(&ndev));
    Y2_REQUIRE(device >= 0 && device < ndev, "y2_create: device %d not present (%d visible)", device, ndev);
    cudaDeviceProp prop;
    Y2_CUDA(cudaGetDeviceProperties(&prop, device));
    Y2_REQUIRE(prop.major == 10, "y2_create: this library contains sm_100a code only; device %d is sm_%d%d", device,
               prop.major, prop.minor);
    Y2_CUDA(cudaSetDevice(device));
    y2_handle* h = new y2_handle();
    h->device = device; h->classes = classes; h->anchors = num_anchors;
    h->num_sms = prop.multiProcessorCount;
    h->arch = arch;
    auto descs = arch == Y2_ARCH_TINY ? tiny_layers(classes, num_anchors) : darknet19_layers(classes, num_anchors);
    for (size_t i = 0; i < descs.size(); ++i) {
        LayerState s;
        s.d = descs[i];
        if (s.d.from_concat) h->cat_c = s.d.cin_s;
        choose_tiles(s.d.cout_s, &s.block_n, &s.cout_pad);
        const size_t K = (size_t)s.d.ksize * s.d.ksize * s.d.cin_s;
        if (cudaMalloc(&s.w_f32, K * s.d.cout_s * sizeof(float)) != cudaSuccess ||
            (i > 0 && cudaMalloc(&s.wpack, 2 * (size_t)s.cout_pad * K * sizeof(bf16)) != cudaSuccess) ||
            cudaMalloc(&s.scale, s.d.cout_s * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&s.bias, s.d.cout_s * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&s.gamma, 4 * (size_t)s.d.cout_s * sizeof(float)) != cudaSuccess) {
            set_error("y2_create: cudaMalloc failed at layer %zu (%s)", i, cudaGetErrorString(cudaGetLastError()));
            h->layers.push_back(s);            // y2_destroy frees this layer's partial allocations with the earlier layers'
            y2_destroy(h);
            return -2;
        }
        cudaMemset(s.scale, 0, s.d.cout_s * sizeof(float));      // padded channels: scale = bias = 0 -> exact zeros
        cudaMemset(s.bias, 0, s.d.cout_s * sizeof(float));
        // raw BN variables at the STORED channel count; padded channels keep gamma = beta = mean = var = 0, which makes their
        // training-mode scale and bias exactly 0 as well (bn_stats: scale = gamma * inv)
        cudaMemset(s.gamma, 0, 4 * (size_t)s.d.cout_s * sizeof(float));
        s.beta = s.gamma + s.d.cout_s; s.mmean = s.beta + s.d.cout_s; s.mvar = s.mmean + s.d.cout_s;
        h->layers.push_back(s);
    }
    *out = h;
    return 0;
}

void y2_destroy(y2_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& e : h->ev) cudaEventDestroy(e);
    for (auto& s : h->layers) {
        cudaFree(s.wpack); cudaFree(s.wpack16); cudaFree(s.wsc); cudaFree(s.w_f32); cudaFree(s.scale); cudaFree(s.bias); cudaFree(s.gamma); cudaFree(s.wpack_dgrad);
    }
    delete h;
}

int y2_num_layers(const y2_handle* h) {
    Y2_REQUIRE(h, "y2_num_layers: null handle");
    return (int)h->layers.size();
}

int y2_layer_info(const y2_handle* h, int layer, int* ksize, int* cin, int* cout, int* has_bn) {
    Y2_REQUIRE(h && layer >= 0 && layer < (int)h->layers.size(), "y2_layer_info: bad layer %d", layer);
    const LayerDesc& d = h->layers[layer].d;
    if (ksize) *ksize = d.ksize;
    if (cin) *cin = d.cin;
    if (cout) *cout = d.cout;
    if (has_bn) *has_bn = d.has_bn;
    return 0;
}

int y2_load_weights(y2_handle* h, int layer, const float* w_hwio, const float* gamma, const float* beta,
                    const float* moving_mean, const float* moving_variance, const float* bias, void* stream) {
    Y2_REQUIRE(h && layer >= 0 && layer < (int)h->layers.size(), "y2_load_weights: bad layer %d", layer);
    Y2_REQUIRE(w_hwio, "y2_load_weights: null weights");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Y2_CUDA(cudaSetDevice(h->device));
    LayerState& L = h->layers[layer];
    const size_t K = (size_t)L.d.ksize * L.d.ksize * L.d.cin;
    if (L.d.cin_s != L.d.cin || L.d.cout_s != L.d.cout) {
        Y2_REQUIRE(w_hwio != L.w_f32, "y2_load_weights: layer %d stores padded channels; pass the variable itself", layer);
        if (pad_weights_launch(w_hwio, L.w_f32, L.d.ksize, L.d.cin, L.d.cout, L.d.cin_s, L.d.cout_s, s)) return -1;
    } else if (w_hwio != L.w_f32) {
        Y2_CUDA(cudaMemcpyAsync(L.w_f32, w_hwio, K * L.d.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    L.pack_fresh = L.pack16_fresh = L.dgrad_fresh = false;      // re-packed by the first consumer (ensure_pack*)
    if (L.d.has_bn) {
        Y2_REQUIRE(moving_mean && moving_variance, "y2_load_weights: layer %d needs BN statistics", layer);
        const size_t cb = L.d.cout * sizeof(float);
        if (gamma) Y2_CUDA(cudaMemcpyAsync(L.gamma, gamma, cb, cudaMemcpyDeviceToDevice, s));
        else Y2_CUDA(cudaMemsetAsync(L.gamma, 0, cb, s));      // never used: gamma == NULL means "1" below
        if (beta) Y2_CUDA(cudaMemcpyAsync(L.beta, beta, cb, cudaMemcpyDeviceToDevice, s));
        else Y2_CUDA(cudaMemsetAsync(L.beta, 0, cb, s));
        Y2_CUDA(cudaMemcpyAsync(L.mmean, moving_mean, cb, cudaMemcpyDeviceToDevice, s));
        Y2_CUDA(cudaMemcpyAsync(L.mvar, moving_variance, cb, cudaMemcpyDeviceToDevice, s));
        Y2_REQUIRE(gamma, "y2_load_weights: layer %d needs gamma (scale=True in the reference, inference.py:63)", layer);
        if (bn_fold_launch(L.gamma, L.beta, L.mmean, L.mvar, 1e-5f, L.scale, L.bias, L.d.cout, s)) return -1;
    } else {
        Y2_REQUIRE(bias, "y2_load_weights: final layer needs biases");
        Y2_CUDA(cudaMemcpyAsync(L.bias, bias, L.d.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    L.loaded = true;
    return 0;
}

}  // extern "C"

// Packed tensor-core operands of a layer's weights, made on first use after y2_load_weights: the inference plan reads the bf16
// planes, the training forward the fp16 planes (a training step never pays for the bf16 pack and vice versa).
static int ensure_pack(y2_handle* h, int i, cudaStream_t s) {
    LayerState& L = h->layers[i];
    if (i == 0 || L.pack_fresh) return 0;
    if (pack_weights_launch(L.w_f32, L.wpack, L.d.ksize, L.d.cin_s, L.d.cout_s, L.cout_pad, s)) return -1;
    L.pack_fresh = true;
    return 0;
}
static int ensure_pack16(y2_handle* h, int i, cudaStream_t s) {
    LayerState& L = h->layers[i];
    if (i == 0 || L.pack16_fresh) return 0;
    const size_t K = (size_t)L.d.ksize * L.d.ksize * L.d.cin_s;
    if (!L.wpack16) Y2_CUDA(cudaMalloc(&L.wpack16, 2 * (size_t)L.cout_pad * K * sizeof(bf16)));
    if (!L.wsc) Y2_CUDA(cudaMalloc(&L.wsc, (2 + (size_t)L.d.cout_s) * sizeof(float)));
    if (pow2_scale_launch(L.w_f32, K * L.d.cout_s, L.wsc, s)) return -1;
    if (fold_scale_launch(nullptr, L.wsc + 1, L.wsc + 2, L.d.cout_s, s)) return -1;
    if (pack_weights_launch(L.w_f32, L.wpack16, L.d.ksize, L.d.cin_s, L.d.cout_s, L.cout_pad, s, 1, L.wsc)) return -1;
    L.pack16_fresh = true;
    return 0;
}

extern "C" {

// workspace layout is a pure function of (B,H,W): used by both the size query and the planner
struct WsLayout {
    std::vector<size_t> act_off, pool_off;
    std::vector<int> oh, ow;
    size_t concat_off = 0, streamk_off = 0, total = 0;
};
static int layout_workspace(const y2_handle* h, int B, int H, int W, WsLayout* out) {
    Y2_REQUIRE(B > 0 && H > 0 && W > 0, "workspace: bad shape");
    Y2_REQUIRE(H % 32 == 0 && W % 32 == 0, "input size %dx%d is not divisible by the downsampling 32 (utils/__init__.py:52-56)", W, H);
    const int nl = (int)h->layers.size();
    out->act_off.assign(nl, 0); out->pool_off.assign(nl, 0); out->oh.assign(nl, 0); out->ow.assign(nl, 0);
    // A layer's output (un-pooled and / or pooled planes) is read by the NEXT layer only (the passthrough tap is copied into
    // the concat buffer by its own layer), so by default the outputs alternate between two arenas sized for the largest
    // layer: 0.46 GB instead of 1.8 GB at B = 32 / 416^2, 7 GB instead of 30 GB at B = 256 / 608^2.  keep_activations = 1
    // gives every layer its own slot so that y2_get_activation can read any of them after the forward (tests, diagnostics).
    std::vector<size_t> a_bytes(nl, 0), p_bytes(nl, 0);
    int ch = H, cw = W;
    for (int i = 0; i < nl; ++i) {
        const LayerState& L = h->layers[i];
        out->oh[i] = ch; out->ow[i] = cw;
        const size_t M = (size_t)B * ch * cw;
        if (i == nl - 1) break;                    // final layer writes the caller's buffer
        if (i == 0) {                              // conv0 writes its pooled output only (kept in the "act" slot)
            a_bytes[i] = align_up(2 * (M / 4) * L.d.cout_s * sizeof(bf16), 1024);
            ch /= 2; cw /= 2;
            continue;
        }
        // the un-pooled output: not for the layer that writes into the concat buffer (conv19), and -- arena mode -- not for
        // a pool layer whose max-pool is certain to be fused into its conv epilogue (build_plan's condition is a superset)
        const bool surely_fused = !h->keep_activations && L.d.pool == 1 && !L.d.passthrough && h->fuse_pool && tc_conv_can_fuse_pool(B, ch, cw);
        if (!L.d.to_concat && !surely_fused) a_bytes[i] = align_up(2 * M * L.d.cout_s * sizeof(bf16), 1024);
        if (L.d.pool == 1) {
            p_bytes[i] = align_up(2 * (M / 4) * L.d.cout_s * sizeof(bf16), 1024);
            ch /= 2; cw /= 2;
        } else if (L.d.pool == 2) {                // stride 1: same extent
            p_bytes[i] = align_up(2 * M * L.d.cout_s * sizeof(bf16), 1024);
        }
    }
    size_t off = 0;
    if (h->keep_activations) {
        for (int i = 0; i < nl - 1; ++i) {
            out->act_off[i] = a_bytes[i] ? off : (size_t)-1;
            off += a_bytes[i];
            out->pool_off[i] = off;
            off += p_bytes[i];
        }
    } else {
        size_t arena = 0;
        for (int i = 0; i < nl - 1; ++i) arena = std::max(arena, a_bytes[i] + p_bytes[i]);
        for (int i = 0; i < nl - 1; ++i) {
            const size_t base = (i & 1) ? arena : 0;
            out->act_off[i] = a_bytes[i] ? base : (size_t)-1;
            out->pool_off[i] = base + a_bytes[i];
        }
        off = 2 * arena;
    }
    out->concat_off = off;
    off = align_up(off + 2 * (size_t)B * ch * cw * h->cat_c * sizeof(bf16), 1024);
    out->streamk_off = off;
    off = align_up(off + tc_conv_streamk_bytes(h->num_sms), 1024);
    out->total = off;
    return 0;
}

size_t y2_workspace_bytes(const y2_handle* h, int B, int H, int W) {
    if (!h) return 0;
    WsLayout l;
    if (layout_workspace(h, B, H, W, &l)) return 0;
    return l.total;
}

static int build_plan(y2_handle* h, int B, int H, int W, void* ws, size_t ws_bytes, int precision, cudaStream_t s) {
    WsLayout l;
    if (layout_workspace(h, B, H, W, &l)) return -1;
    Y2_REQUIRE(ws && ws_bytes >= l.total, "y2_darknet_forward: workspace too small (%zu < %zu)", ws_bytes, l.total);
    Y2_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "y2_darknet_forward: workspace must be 1024-byte aligned");
    Plan& P = h->plan;
    P = Plan();
    const int nl = (int)h->layers.size();
    char* base = static_cast<char*>(ws);
    P.B = B; P.H = H; P.W = W; P.ws = ws; P.ws_bytes = ws_bytes; P.precision = precision;
    P.act.assign(nl, nullptr); P.pooled.assign(nl, nullptr); P.oh = l.oh; P.ow = l.ow;
    P.launch.resize(nl);
    P.fused.assign(nl, 0);
    P.concat = reinterpret_cast<bf16*>(base + l.concat_off);
    P.streamk = base + l.streamk_off;
    Y2_CUDA(cudaMemsetAsync(P.streamk, 0, 4096, s));          // hand-off flags start at 0 (epochs are >= 1)
    for (int i = 0; i < nl; ++i) {
        if (l.act_off[i] != (size_t)-1 && i != nl - 1) P.act[i] = reinterpret_cast<bf16*>(base + l.act_off[i]);
        if (i > 0 && h->layers[i].d.pool) P.pooled[i] = reinterpret_cast<bf16*>(base + l.pool_off[i]);
    }
    const int cat_c = h->cat_c;
    const bf16* input = P.act[0];
    for (int i = 1; i < nl; ++i) {
        const LayerState& L = h->layers[i];
        const int oh = l.oh[i], ow = l.ow[i];
        const size_t M = (size_t)B * oh * ow;
        if (L.d.from_concat) input = P.concat;     // conv20 reads concat([reorg, conv19])
        TcConvLaunch& T = P.launch[i];
        // max-pool layers: fuse the 2x2/2 pool into the conv epilogue when the batch / extent admit the spatial tiling
        const int halo = (h->halo && tc_conv_can_halo(B, oh, ow, L.d.cin_s, L.d.ksize, L.cout_pad, L.block_n, precision == 0)) ? h->halo : 0;
        const bool fuse = L.d.pool == 1 && h->fuse_pool && (halo || tc_conv_can_fuse_pool(B, oh, ow));
        P.fused[i] = fuse;
        const int pair = (h->pair && tc_conv_can_pair(L.d.cin_s, L.block_n, halo) &&
                          (h->pair >= 2 || (L.d.ksize == 3 && L.block_n == 256))) ? 1 : 0;
        if (tc_conv_plan(&T, input, B, oh, ow, L.d.cin_s, L.d.ksize, L.wpack, L.d.cout_s, L.cout_pad, L.block_n,
                         0, precision == 0, h->num_sms, P.streamk, fuse ? 1 : 0, halo, pair))
            return -1;
        ConvParams& p = T.p;
        if (fuse) {
            p.pool_hi = P.pooled[i];
            p.pool_lo = P.pooled[i] + (M / 4) * L.d.cout_s;
            p.ldp = L.d.cout_s;
        }
        p.scale = L.d.has_bn ? L.scale : nullptr;
        p.bias = L.bias;
        p.leaky = L.d.has_bn ? 1 : 0;
        if (i == nl - 1) {
            p.mode = EPI_F32; p.ldc = L.d.cout;    // out pointer patched per call
        } else if (L.d.to_concat) {
            p.mode = EPI_PLANES; p.ldc = cat_c;
            p.out_hi = P.concat + (cat_c - L.d.cout_s);
            p.out_lo = P.concat + M * cat_c + (cat_c - L.d.cout_s);
        } else {
            p.mode = EPI_PLANES; p.ldc = L.d.cout_s;
            p.out_hi = P.act[i];
            p.out_lo = P.act[i] + M * L.d.cout_s;
            if (fuse && !L.d.passthrough) p.out_hi = p.out_lo = nullptr;   // the un-pooled tensor is never materialised
            // passthrough layer with its pool fused (spatial tiles): the epilogue writes the un-pooled output straight into the
            // concat buffer in reorg order -- the separate reorg pass (29 us at B = 32) and the act slot are not needed.  Tests
            // that read conv12's activation back (keep_activations) keep the separate pass.
            if (fuse && L.d.passthrough && !h->keep_activations && p.tx != 0 && cat_c) {
                p.reorg = 1; p.ldc = cat_c;
                p.out_hi = P.concat;
                p.out_lo = P.concat + (M / 4) * cat_c;
                P.reorg_fused = true;
            }
        }
        if (tc_conv_bind_output(&T)) return -1;
        input = L.d.pool ? P.pooled[i] : P.act[i];
    }
    P.valid = true;
    return 0;
}

int y2_darknet_forward(y2_handle* h, const float* x, int B, int H, int W, float* out, void* ws, size_t ws_bytes,
                       int precision, void* stream) {
    Y2_REQUIRE(h && x && out, "y2_darknet_forward: null argument");
    Y2_REQUIRE(precision == 0 || precision == 1, "y2_darknet_forward: precision must be 0 or 1");
    for (size_t i = 0; i < h->layers.size(); ++i)
        Y2_REQUIRE(h->layers[i].loaded, "y2_darknet_forward: layer %zu has no weights (y2_load_weights)", i);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Y2_CUDA(cudaSetDevice(h->device));
    Plan& P = h->plan;
    if (!P.valid || P.B != B || P.H != H || P.W != W || P.ws != ws || P.precision != precision || P.ws_bytes != ws_bytes)
        if (build_plan(h, B, H, W, ws, ws_bytes, precision, s)) return -1;
    const int nl = (int)h->layers.size();
    for (int i = 1; i < nl; ++i)
        if (ensure_pack(h, i, s)) return -1;
    const LayerState& L0 = h->layers[0];
    if (h->profiling) Y2_CUDA(cudaEventRecord(h->ev[0], s));
    {
        const size_t M0 = (size_t)B * (H / 2) * (W / 2);
        Y2_REQUIRE(L0.d.cout_s == 32 && L0.d.pool == 1, "y2_darknet_forward: conv0 kernel is built for 32 stored output channels + pool");
        if (h->conv0_tc && conv0_tc_applicable(H, W)) {
            if (conv0_tc_pool_launch(x, L0.w_f32, L0.scale, L0.bias, P.act[0], P.act[0] + M0 * L0.d.cout_s, B, H, W, h->num_sms, s, h->conv0_tc >= 2)) return -1;
        } else if (conv0_pool_launch(x, L0.w_f32, L0.scale, L0.bias, P.act[0], P.act[0] + M0 * L0.d.cout_s, B, H, W, s)) return -1;
    }
    if (h->profiling) Y2_CUDA(cudaEventRecord(h->ev[1], s));
    for (int i = 1; i < nl; ++i) {
        const LayerState& L = h->layers[i];
        TcConvLaunch T = P.launch[i];
        if (h->profiling) Y2_CUDA(cudaEventRecord(h->ev[4 * i], s));
        if (i == nl - 1) T.p.out_f32 = out;
        if (tc_conv_launch(T, s)) return -1;
        if (h->profiling) Y2_CUDA(cudaEventRecord(h->ev[4 * i + 1], s));
        const int oh = P.oh[i], ow = P.ow[i];
        const size_t M = (size_t)B * oh * ow;
        if (L.d.passthrough && !P.reorg_fused) {
            // reorg(passthrough) -> concat channels [0, 2048); both planes in one launch (batch 2B)
            if (reorg_launch(P.act[i], P.concat, 2 * B, oh, ow, L.d.cout_s, 2, 2, h->cat_c, s)) return -1;
        }
        if (L.d.pool && !P.fused[i]) {
            const size_t Mp = L.d.pool == 1 ? M / 4 : M;
            if (maxpool_planes_launch(P.act[i], P.act[i] + M * L.d.cout_s, P.pooled[i],
                                      P.pooled[i] + Mp * L.d.cout_s, B, oh, ow, L.d.cout_s, s, L.d.pool == 1 ? 2 : 1))
                return -1;
        }
        if (h->profiling) Y2_CUDA(cudaEventRecord(h->ev[4 * i + 2], s));
    }
    return 0;
}

int y2_set_profiling(y2_handle* h, int enable) {
    Y2_REQUIRE(h, "y2_set_profiling: null handle");
    Y2_CUDA(cudaSetDevice(h->device));
    if (enable && h->ev.empty()) {
        h->ev.resize(4 * h->layers.size());
        for (auto& e : h->ev) Y2_CUDA(cudaEventCreate(&e));
    }
    h->profiling = enable != 0;
    return 0;
}

int y2_get_layer_ms(y2_handle* h, float* conv_ms, float* post_ms) {
    Y2_REQUIRE(h && conv_ms && post_ms, "y2_get_layer_ms: null argument");
    Y2_REQUIRE(h->profiling && !h->ev.empty(), "y2_get_layer_ms: profiling is off");
    const int nl = (int)h->layers.size();
    Y2_CUDA(cudaEventSynchronize(h->ev[4 * (nl - 1) + 2]));
    Y2_CUDA(cudaEventElapsedTime(&conv_ms[0], h->ev[0], h->ev[1]));      // conv0 + pool (CUDA cores)
    post_ms[0] = 0.f;
    for (int i = 1; i < nl; ++i) {
        Y2_CUDA(cudaEventElapsedTime(&conv_ms[i], h->ev[4 * i], h->ev[4 * i + 1]));
        Y2_CUDA(cudaEventElapsedTime(&post_ms[i], h->ev[4 * i + 1], h->ev[4 * i + 2]));
    }
    return 0;
}

unsigned long long y2_launch_count(void) { return g_launches.load(); }

int y2_get_activation(y2_handle* h, int layer, int pooled, float* dst, void* stream) {
    Y2_REQUIRE(h && dst, "y2_get_activation: null argument");
    const Plan& P = h->plan;
    Y2_REQUIRE(P.valid, "y2_get_activation: no forward has run");
    const int nl = (int)h->layers.size();
    Y2_REQUIRE(layer >= 0 && layer < nl - 1, "y2_get_activation: layer %d out of range", layer);
    Y2_REQUIRE(h->keep_activations || h->layers[layer].d.to_concat,
               "y2_get_activation: the layers' outputs share two alternating arenas; y2_set_option(h, \"keep_activations\", 1) before the "
               "forward keeps every layer's output readable");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const LayerDesc& d = h->layers[layer].d;
    const size_t M = (size_t)P.B * P.oh[layer] * P.ow[layer];
    if (layer == 0) {
        Y2_REQUIRE(pooled, "y2_get_activation: conv0 is fused with its max-pool; only the pooled tensor exists");
        return merge_planes_launch(P.act[0], P.act[0] + (M / 4) * d.cout_s, dst, M / 4, d.cout, d.cout_s, s);
    }
    if (pooled) {
        Y2_REQUIRE(d.pool, "y2_get_activation: layer %d has no pool", layer);
        const size_t Mp = d.pool == 1 ? M / 4 : M;
        return merge_planes_launch(P.pooled[layer], P.pooled[layer] + Mp * d.cout_s, dst, Mp, d.cout, d.cout_s, s);
    }
    if (d.to_concat) {
        const int cat_c = h->cat_c;
        return merge_planes_launch(P.concat + (cat_c - d.cout_s), P.concat + M * cat_c + (cat_c - d.cout_s), dst, M, d.cout, cat_c, s);
    }
    Y2_REQUIRE(!(P.fused[layer] && !d.passthrough),
               "y2_get_activation: layer %d's max-pool is fused into its conv epilogue; the un-pooled tensor is not "
               "materialised (y2_set_option(h, \"fuse_pool\", 0) keeps it)", layer);
    return merge_planes_launch(P.act[layer], P.act[layer] + M * d.cout_s, dst, M, d.cout, d.cout_s, s);
}

/* Options: "fuse_pool" (default 1) -- fuse the 2x2 max-pools into the conv epilogues when the shape allows it;
 * "halo" (default 1); "pair" -- CTA-pair convs (see y2_handle::pair). */
int y2_set_option(y2_handle* h, const char* key, int value) {
    Y2_REQUIRE(h && key, "y2_set_option: null argument");
    if (strcmp(key, "fuse_pool") == 0) { h->fuse_pool = value ? 1 : 0; h->plan.valid = false; return 0; }      // changes y2_workspace_bytes in arena mode
    if (strcmp(key, "halo") == 0) { h->halo = value; h->plan.valid = false; return 0; }
    if (strcmp(key, "pair") == 0) { h->pair = value; h->plan.valid = false; return 0; }
    if (strcmp(key, "conv0_tc") == 0) { h->conv0_tc = value; return 0; }
    if (strcmp(key, "keep_activations") == 0) { h->keep_activations = value ? 1 : 0; h->plan.valid = false; return 0; }
    if (strcmp(key, "train_f16") == 0) { h->train_f16 = value ? 1 : 0; h->tplan.valid = false; return 0; }
    if (strcmp(key, "train_kcap") == 0) { Y2_REQUIRE(value >= 0, "y2_set_option: train_kcap must be >= 0"); h->train_kcap = value; return 0; }
    set_error("y2_set_option: unknown option '%s'", key);
    return -1;
}

int y2_conv2d(const float* x, int B, int H, int W, int cin, const float* w_hwio, int ksize, int cout,
              const float* scale, const float* bias, int leaky, float* y, int precision, int block_n, int max_ctas,
              void* stream) {
    Y2_REQUIRE(x && w_hwio && y, "y2_conv2d: null argument");
    Y2_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "y2_conv2d: bad shape B=%d H=%d W=%d cin=%d cout=%d", B, H, W, cin, cout);
    Y2_REQUIRE(ksize == 1 || ksize == 3, "y2_conv2d: ksize must be 1 or 3 (got %d)", ksize);
    Y2_REQUIRE(cin % 32 == 0, "y2_conv2d: cin must be a multiple of 32 (got %d)", cin);
    Y2_REQUIRE(block_n == 0 || (block_n % 32 == 0 && block_n >= 32 && block_n <= 256), "y2_conv2d: block_n %d invalid", block_n);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    Y2_CUDA(cudaGetDevice(&dev));
    const int num_sms = device_sm_count(dev);
    int bn = 0, cpad = 0;
    choose_tiles(cout, &bn, &cpad);
    if (block_n > 0) { bn = block_n; cpad = (int)align_up((size_t)cout, (size_t)bn); }
    const size_t M = (size_t)B * H * W, K = (size_t)ksize * ksize * cin;
    bf16 *xp = nullptr, *wp = nullptr;
    float* wsc = nullptr;
    void* sk = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&xp, 2 * M * cin * sizeof(bf16)) != cudaSuccess || cudaMalloc(&wp, 2 * (size_t)cpad * K * sizeof(bf16)) != cudaSuccess ||
            cudaMalloc(&sk, tc_conv_streamk_bytes(num_sms)) != cudaSuccess) { set_error("y2_conv2d: cudaMalloc failed"); break; }
        if (cudaMemsetAsync(sk, 0, 4096, s) != cudaSuccess) { set_error("y2_conv2d: memset failed"); break; }
        // precision 0 / 1: bf16 planes.  2: fp16 planes (activations as they are, weights pre-scaled by a power of two that the
        // epilogue scale undoes).  3 / 4 (diagnostics of the mixed-format MMAs the training step relies on): 3 = bf16
        // activation planes x fp16 weight planes (dgrad), 4 = fp16 activation planes x bf16 weight planes.
        const int x16 = (precision == 2 || precision == 4) ? 1 : 0, w16 = (precision == 2 || precision == 3) ? 1 : 0;
        if (w16) {
            if (cudaMalloc(&wsc, (2 + (size_t)cout) * sizeof(float)) != cudaSuccess) { set_error("y2_conv2d: cudaMalloc failed"); break; }
            if (pow2_scale_launch(w_hwio, K * cout, wsc, s)) break;
            if (fold_scale_launch(scale, wsc + 1, wsc + 2, cout, s)) break;
            scale = wsc + 2;
        }
        if (split_planes_launch(x, xp, xp + M * cin, M * cin, s, x16)) break;
        if (pack_weights_launch(w_hwio, wp, ksize, cin, cout, cpad, s, w16, w16 ? wsc : nullptr)) break;
        g_conv_fmt = (x16 ? (FMT_A_HI | FMT_A_LO) : 0) | (w16 ? (FMT_B_HI | FMT_B_LO) : 0);
        if (precision >= 2) precision = 0;
        TcConvLaunch T;
        const int halo = (g_conv_force_halo && tc_conv_can_halo(B, H, W, cin, ksize, cpad, bn, precision == 0)) ? g_conv_force_halo : 0;
        const int pair = (g_conv_force_pair && tc_conv_can_pair(cin, bn, halo)) ? 1 : 0;
        if (tc_conv_plan(&T, xp, B, H, W, cin, ksize, wp, cout, cpad, bn, max_ctas, precision == 0, num_sms, sk, 0, halo, pair)) break;
        T.p.scale = scale; T.p.bias = bias; T.p.leaky = leaky;
        T.p.out_f32 = y; T.p.ldc = cout; T.p.mode = EPI_F32;
        if (tc_conv_bind_output(&T)) break;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        unsigned long long* dbg = nullptr;
        if (g_dbg_host) { cudaMalloc(&dbg, (size_t)T.grid * 4 * sizeof(unsigned long long)); T.p.dbg = dbg; }
        if (tc_conv_launch(T, s)) break;                       // warm-up / result
        cudaEventRecord(e0, s);
        if (tc_conv_launch(T, s)) break;                       // timed (same output)
        cudaEventRecord(e1, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("y2_conv2d: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        cudaEventElapsedTime(&g_last_conv_ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (dbg) {
            g_dbg_n = T.grid < 1024 ? T.grid : 1024;
            cudaMemcpy(g_dbg_host, dbg, (size_t)g_dbg_n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            cudaFree(dbg);
            g_dbg_sched[0] = T.p.dp_tiles; g_dbg_sched[1] = T.p.sk_ctas; g_dbg_sched[2] = T.grid; g_dbg_sched[3] = T.p.kblocks_total;
        }
        if (tc_conv_check_watchdog()) break;
        rc = 0;
    } while (0);
    g_conv_fmt = 0;
    cudaFree(xp); cudaFree(wp); cudaFree(sk); cudaFree(wsc);
    return rc;
}

// Diagnostics: schedule override / timing of the last y2_conv2d kernel / per-CTA globaltimer stamps.
int y2_debug_set(int key, double value) {
    if (key == 0) g_sched_override = (int)value;
    else if (key == 1) g_sched_handoff_kb = value;
    else if (key == 3) g_conv_dbg_flags = (int)value;       // ConvParams::dbg_flags of the convs planned from now on
    else if (key == 4) g_conv_force_halo = (int)value;
    else if (key == 13) g_nms_select_cg = (int)value;       // nms_select_kernel: classes per CTA (0 by regime, 8 | 16 | 32)
    else if (key == 11) g_nms_apply_mode = (int)value;      // nms_apply_kernel work items: 0 by regime, 1 dynamic chunks, 2 one per class column
    else if (key == 9) g_conv_fmt = (int)value;             // FMT_* bits of the convs planned from now on (y2_conv2d sets them itself)
    else if (key == 10) g_wgrad_fmt = (int)value;           // FMT_* bits of y2_conv2d_wgrad: 3 = x planes in fp16
    else if (key == 8) g_conv_kcap = (int)value;            // longest tensor-core accumulation chain in k-blocks (0 = unlimited; default 32)
    else if (key == 7) g_conv_force_pair = (int)value;      // y2_conv2d: CTA-pair mode where applicable
    else if (key == 6) g_conv_tma_store = (int)value;       // TMA-store epilogue (default 1)
    else if (key == 5) g_conv_pdl = (int)value;             // programmatic dependent launch of the conv kernels (default 1)      // y2_conv2d: halo mode (1|2) where applicable
    else if (key == 2) { if (value != 0 && !g_dbg_host) g_dbg_host = new unsigned long long[1024 * 4]; if (value == 0) { delete[] g_dbg_host; g_dbg_host = nullptr; } }
    else return -1;
    return 0;
}
float y2_debug_last_conv_ms(void) { return g_last_conv_ms; }
int y2_debug_cta_times(unsigned long long* out, int max_ctas, int* sched4) {
    if (!g_dbg_host) return 0;
    const int n = g_dbg_n < max_ctas ? g_dbg_n : max_ctas;
    memcpy(out, g_dbg_host, (size_t)n * 4 * sizeof(unsigned long long));
    if (sched4) memcpy(sched4, g_dbg_sched, sizeof(g_dbg_sched));
    return n;
}

// Diagnostic twin of y2_conv2d for the weight-gradient GEMM: dw[k][k][cin][cout] from fp32 NHWC x and dy.
int y2_conv2d_wgrad(const float* x, int B, int H, int W, int cin, const float* dy, int ksize, int cout, float* dw,
                    int max_ctas, void* stream) {
    Y2_REQUIRE(x && dy && dw, "y2_conv2d_wgrad: null argument");
    Y2_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "y2_conv2d_wgrad: bad shape B=%d H=%d W=%d cin=%d cout=%d", B, H, W, cin, cout);
    Y2_REQUIRE(ksize == 1 || ksize == 3, "y2_conv2d_wgrad: ksize must be 1 or 3 (got %d)", ksize);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    Y2_CUDA(cudaGetDevice(&dev));
    const int num_sms = device_sm_count(dev);
    const size_t M = (size_t)B * H * W;
    const int dpitch = (int)align_up((size_t)cout, 64);
    bf16 *xp = nullptr, *dp = nullptr;
    void* sk = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&xp, 2 * M * cin * sizeof(bf16)) != cudaSuccess || cudaMalloc(&dp, 2 * M * dpitch * sizeof(bf16)) != cudaSuccess ||
            cudaMalloc(&sk, tc_conv_streamk_bytes(num_sms)) != cudaSuccess) { set_error("y2_conv2d_wgrad: cudaMalloc failed"); break; }
        if (cudaMemsetAsync(sk, 0, 4096, s) != cudaSuccess) { set_error("y2_conv2d_wgrad: memset failed"); break; }
        if (split_planes_launch(x, xp, xp + M * cin, M * cin, s, (g_wgrad_fmt & FMT_A_HI) ? 1 : 0)) break;     // y2_debug_set(10, 3): x planes in fp16
        if (split_planes_pad_launch(dy, cout, dp, dp + M * dpitch, M, cout, dpitch, s)) break;
        if (wgrad_tc_run(xp, B, H, W, cin, ksize, dp, cout, dpitch, dw, max_ctas, num_sms, sk, s)) break;
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("y2_conv2d_wgrad: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (wgrad_check_watchdog()) break;
        rc = 0;
    } while (0);
    cudaFree(xp); cudaFree(dp); cudaFree(sk);
    return rc;
}

int y2_leaky_relu(const float* in, size_t n, float alpha, float* out, void* stream) {
    return leaky_relu_launch(in, out, n, alpha, static_cast<cudaStream_t>(stream));
}

int y2_reorg(const float* in, int B, int H, int W, int C, int stride, float* out, void* stream) {
    Y2_REQUIRE(in && out, "y2_reorg: null argument");
    Y2_REQUIRE(stride >= 1 && H % stride == 0 && W % stride == 0, "y2_reorg: H, W must be divisible by stride");
    return reorg_launch(in, out, B, H, W, C, stride, 4, (long long)C * stride * stride, static_cast<cudaStream_t>(stream));
}

int y2_head_decode(const float* net, int B, int Hc, int Wc, int A, int C, const float* anchors,
                   const y2_head_outputs* outs, void* stream) {
    return head_decode_launch(net, B, Hc, Wc, A, C, anchors, outs, static_cast<cudaStream_t>(stream));
}

size_t y2_loss_workspace_bytes(int B, int Hc, int Wc) { return loss_workspace_bytes(B, Hc, Wc); }

int y2_loss_fwd_bwd(const float* net, int B, int Hc, int Wc, int A, int C, const float* anchors, const float* mask,
                    const float* prob, const float* coords, const float* offset_xy_min, const float* offset_xy_max,
                    const float* areas, const float hparam[4], float* objectives, float* dnet, void* ws,
                    size_t ws_bytes, void* stream) {
    return loss_launch(net, B, Hc, Wc, A, C, anchors, mask, prob, coords, offset_xy_min, offset_xy_max, areas, hparam,
                       objectives, dnet, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

size_t y2_nms_workspace_bytes(int B, int N, int C) { return nms_workspace_bytes(B, N, C); }

int y2_nms(float* conf, const float* xy_min, const float* xy_max, int B, int N, int C, float threshold,
           float threshold_iou, int32_t* order_out, int32_t* status_out, void* ws, size_t ws_bytes, void* stream) {
    return nms_launch(conf, xy_min, xy_max, B, N, C, threshold, threshold_iou, order_out, status_out, ws, ws_bytes,
                      static_cast<cudaStream_t>(stream));
}

/* Returns 0 if no tcgen05 pipeline watchdog fired since the last call (device must be idle). */
int y2_check_async_errors(void) {
    const int a = tc_conv_check_watchdog();
    const int b = wgrad_check_watchdog();
    const int c = conv0_tc_check_watchdog();
    return a ? a : (b ? b : c);
}

}  // extern "C"

#include "y2_train_api.inc"
#include "y2_optim_api.inc"
