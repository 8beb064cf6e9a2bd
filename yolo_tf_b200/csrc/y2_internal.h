// Internal (non-ABI) declarations shared by the translation units of libyolo2_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stddef.h>
#include <stdint.h>

namespace y2 {

// ---- error plumbing (thread-local message behind y2_last_error()) ----
void set_error(const char* fmt, ...);
const char* last_error();
#define Y2_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            y2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)
#define Y2_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            y2::set_error(__VA_ARGS__);  \
            return -1;                   \
        }                                \
    } while (0)

int device_sm_count(int device);
// cudaFuncSetAttribute / occupancy results are PER DEVICE: one bit per device ordinal of the current device; true the first
// time it is asked on that device (ordinals >= 64: always true, i.e. the setup is simply repeated)
static inline bool first_use_on_current_device(unsigned long long& seen) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (seen & (1ull << d)) return false;
    seen |= 1ull << d;
    return true;
}
// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
void note_launch();

// ---- activation storage: "split planes" ----
// A float32 NHWC tensor T[M][C] lives in HBM as two bf16 planes, hi = bf16(T) and
// lo = bf16(T - hi), plane p at base + p*M*ld (ld = row pitch in elements). hi+lo carries 16
// significand bits; the tcgen05 conv multiplies planes pairwise (hi*hi + hi*lo + lo*hi).
typedef __nv_bfloat16 bf16;

// Plane formats of a GEMM's operands, one bit each (1 = fp16, 0 = bf16)
enum : int { FMT_A_HI = 1, FMT_A_LO = 2, FMT_B_HI = 4, FMT_B_LO = 8 };

// ---- plane element formats ----
// The planes of a tensor are 16-bit elements, bf16 (range-safe: float32's exponent; hi + lo = 16 significand bits) or fp16
// (hi + lo = 22 significand bits; range 6e-8 .. 65504: used where the producer bounds the range -- batch-normalised
// activations of the training forward, and weights pre-scaled by a per-layer power of two).  Storage type stays `bf16*`
// (16-bit slots); `f16` says how the bits are read.
#ifdef __CUDACC__
#include <cuda_fp16.h>
__device__ __forceinline__ unsigned short plane_enc(float v, int f16) {
    if (f16) return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)));   // saturate, never inf
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float plane_dec(unsigned short u, int f16) {
    return f16 ? __half2float(__ushort_as_half(u)) : __uint_as_float((uint32_t)u << 16);
}
__device__ __forceinline__ void plane_split(float v, int f16, unsigned short& hi, unsigned short& lo) {
    hi = plane_enc(v, f16);
    lo = plane_enc(v - plane_dec(hi, f16), f16);
}
__device__ __forceinline__ uint32_t plane_pack2(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }
#endif

// ---- tcgen05 implicit-GEMM conv (y2_conv_tc.cu) ----
enum EpilogueMode : int {
    EPI_PLANES = 0,    // y = leaky?(acc*scale+bias) -> bf16 hi/lo planes
    EPI_F32 = 1,       // y = leaky?(acc*scale+bias) -> float32 (final layer / debugging)
};

struct ConvParams {
    int M, N, Cin, ksize, B, H, W;
    int block_n, m_tiles, n_tiles, kblocks_total;
    int dp_tiles, sk_ctas;   // schedule (choose_schedule)
    int tx, ty, tb, tiles_x, tiles_y;   // tx != 0: M-tiles are tx x ty pixel blocks of tb images (fused max-pool layers)
    bf16 *pool_hi, *pool_lo; // fused 2x2/2 max-pool output planes [B][H/2][W/2][N], row pitch ldp
    long long ldp;
    int cout_pad;            // rows per weight plane in the packed weight matrix
    int num_stages;
    int mode;                // EpilogueMode
    int leaky;
    const float* scale;      // [N] (null -> 1)
    const float* bias;       // [N] (null -> 0)
    bf16* out_hi;            // EPI_PLANES (already offset to the first output channel)
    bf16* out_lo;
    float* out_f32;          // EPI_F32
    float* sk_partial;       // stream-K: [grid][128][block_n] raw fp32 partial tiles
    float* sk_run;           // [grid][128][block_n] running sums of capped accumulation chains (head segments)
    unsigned int* sk_flags;  // stream-K: [grid] hand-off flags (value = launch epoch)
    unsigned int epoch;      // set per launch
    unsigned long long* dbg; // optional [grid][4] globaltimer stamps: start, stream-K phase start, hand-off wait start, end
    long long ldc;           // output row pitch, elements
    // halo mode (3x3, Cin = 32, one N tile): the M-tile is an 8 x 16 pixel block whose (8+2) x (16+2) input halo is
    // fetched ONCE per tile (tile-mode TMA) and every tap's A operand is a shifted window of it; the packed weights of
    // all 9 taps stay resident in shared memory (descriptor starts are whole-row shifts, 8-row group pitch = 10 rows).
    int halo, halo_plane_bytes, halo_tx_bytes, bres_bytes;
    // TMA-store epilogue (linear tiles): every epilogue warp stages its 32 rows x 32 columns in a private swizzled
    // shared-memory slab and one lane writes it out with cp.async.bulk.tensor (full lines, off the LSU).  stg_bytes =
    // 4 warps x store_bufs x 4 KiB at the start of dynamic shared memory; tma_store is set by tc_conv_bind_output.
    int tma_store, store_bufs, stg_bytes;
    int dbg_flags;           // diagnostics: 1 = skip TMA loads, 2 = skip MMAs
    // CTA-pair mode (cta_group::2): m_tiles counts 256-row pair tiles, m_tiles128 the 128-row tiles that exist
    int pair, m_tiles128;
    int kcap;                // longest accumulation chain in k-blocks (0 = unlimited), see CapIter
    int reorg;               // spatial tiles only: write the un-pooled output in space-to-depth(2) order (ldc = pitch of that buffer)
    int fmt;                 // FMT_* bits (y2_ptx.cuh): element format of the A (activation) and B (weight) hi / lo planes; 0 = all bf16
};
extern int g_conv_tma_store;   // 1 = TMA-store epilogue where the layout allows it
extern int g_nms_select_cg;            // diagnostics (y2_debug_set key 13): classes per nms_select_kernel CTA (0 = by regime)
extern int g_nms_apply_mode;           // diagnostics (y2_debug_set key 11): work-item scheme of nms_apply_kernel
extern int g_conv_fmt, g_wgrad_fmt;   // FMT_* bits of the GEMMs planned from now on (diagnostic entry points; the network sets them per launch)
extern int g_conv_kcap;        // ConvParams::kcap of the convs planned from now on (default 32)
extern int g_conv_force_pair;  // y2_conv2d: run eligible convs as CTA pairs (diagnostics / tests)
extern int g_conv_dbg_flags, g_conv_force_halo, g_conv_pdl;   // g_conv_pdl: launch convs with programmatic stream serialization

// Hybrid schedule for `tiles` output tiles of `KB` k-blocks on `ctas` persistent CTAs (max_ctas > 0 caps it):
// full waves of whole tiles run data-parallel; the last partial wave of r tiles is either one more
// data-parallel wave (cost KB) or a stream-K split over all CTAs (cost r*KB/ctas + hand-off overhead),
// whichever is cheaper.  max_ctas < 0 forces stream-K for everything (tests).
extern int g_sched_override;          // 0 = cost model, 1 = data-parallel only, 2 = stream-K everything (diagnostics)
extern double g_sched_handoff_kb;
static inline void choose_schedule(long long tiles, int KB, int num_sms, int max_ctas, double kb_weight,
                                   size_t shared_operand_bytes, int* dp_tiles, int* sk_ctas, int* grid) {
    // hand-off (partial write, flag, staged read + finish) ~ 17 us measured = 13 k-block times of a 128x256x64 step;
    // kb_weight = this kernel's k-block cost relative to that (block_n/256 * BK/64)
    const double HANDOFF_KB = g_sched_handoff_kb / (kb_weight > 0.05 ? kb_weight : 0.05);
    if (g_sched_override == 2 && max_ctas == 0) max_ctas = -num_sms;
    // Measured (profiles/probe_sched_r1.json): when the operand every tile shares (the packed weights) stays
    // L2-resident, stream-K over EVERYTHING beats the hybrid (its hand-offs hide behind the other CTAs' MMAs instead
    // of forming a serial tail): conv8 0.120 vs 0.133 ms, conv13 0.118 vs 0.143, conv18 0.209 vs 0.249.  With the
    // 113 MB weights of conv20 it loses (0.866 vs 0.647: every CTA streams a different K phase from DRAM).
    if (g_sched_override == 0 && max_ctas == 0 && tiles < 4LL * num_sms && shared_operand_bytes > 0 &&
        shared_operand_bytes <= ((size_t)48 << 20) && (double)tiles * KB / num_sms >= 2.0 * HANDOFF_KB)
        max_ctas = -num_sms;
    long long G = num_sms;
    const bool force_sk = max_ctas < 0;
    if (max_ctas < 0) max_ctas = -max_ctas;
    if (max_ctas > 0 && max_ctas < G) G = max_ctas;
    long long full = (tiles / G) * G, r = tiles - full;
    if (force_sk) { full = 0; r = tiles; }
    long long sk = 0;
    if (r > 0) {
        sk = G;
        if (sk > r * KB / 4) sk = r * KB / 4;          // never fewer than 4 k-blocks per CTA
        if (sk < 1) sk = 1;
        const double sk_cost = (double)r * KB / (double)sk + HANDOFF_KB;
        if (!force_sk && (sk_cost >= (double)KB || sk <= r || g_sched_override == 1)) { full = tiles; r = 0; sk = 0; }   // plain extra wave
    }
    *dp_tiles = (int)full;
    *sk_ctas = (int)sk;
    long long g = full < G ? full : G;
    if (sk > g) g = sk;
    if (g < 1) g = 1;
    *grid = (int)g;
}

struct TcConvLaunch {
    CUtensorMap map_a;       // im2col map over the input planes (C, W, H, 2B)
    CUtensorMap map_w;       // tiled map over packed weights (K, 2*cout_pad)
    CUtensorMap map_o;       // output map for the TMA-store epilogue (tc_conv_bind_output)
    ConvParams p;
    int grid, smem_bytes, block_k, split3;
};

// Builds tensor maps + launch geometry for one conv.  in_planes: bf16 [2][B][H][W][Cin].
// wpack: bf16 [2][cout_pad][ksize*ksize*Cin] (k = tap*Cin + c).  Returns 0 or <0 (error set).
int tc_conv_plan(TcConvLaunch* L, const bf16* in_planes, int B, int H, int W, int Cin, int ksize,
                 const bf16* wpack, int cout, int cout_pad, int block_n, int max_ctas, int split3, int num_sms,
                 void* streamk_ws, int fuse_pool = 0, int halo = 0, int pair = 0);
// true when the layer can run as CTA pairs (cta_group::2 MMAs, see conv_tc_kernel)
bool tc_conv_can_pair(int Cin, int block_n, int halo);
// Call after the output pointers / pitch / mode of L->p are final: builds the output tensor map and enables the TMA-store
// epilogue when the layout allows it (linear tiles, 16-byte pitches); otherwise the per-thread store path stays.
int tc_conv_bind_output(TcConvLaunch* L);
// true when the layer can run in halo mode (see ConvParams::halo)
bool tc_conv_can_halo(int B, int H, int W, int Cin, int ksize, int cout_pad, int block_n, int split3);
// true when (B, H, W) admits the spatial tiling the fused max-pool epilogue needs
bool tc_conv_can_fuse_pool(int B, int H, int W);
// bytes of the stream-K scratch (flags page + one partial tile per SM); must be zeroed once before first use
size_t tc_conv_streamk_bytes(int num_sms);
int tc_conv_launch(const TcConvLaunch& L, cudaStream_t stream);
// Reads and clears the barrier watchdog; returns 0 if it never fired, else sets the error.
int tc_conv_check_watchdog();

// ---- elementwise / layout kernels (y2_layout.cu) ----
int split_planes_launch(const float* src, bf16* dst_hi, bf16* dst_lo, size_t n, cudaStream_t s, int f16 = 0);
// amax of a device array -> out[0] = 2^k, out[1] = 2^-k with amax * 2^k in [2^13, 2^14) (k = 0 for an all-zero array)
int pow2_scale_launch(const float* src, size_t n, float* out2, cudaStream_t s);
// out[i] = (scale ? scale[i] : 1) * factor[0]   (folds the 2^-k of pre-scaled fp16 weight planes into the epilogue scale)
int fold_scale_launch(const float* scale, const float* factor, float* out, int n, cudaStream_t s);
int merge_planes_launch(const bf16* hi, const bf16* lo, float* dst, size_t rows, int cols, long long ld,
                        cudaStream_t s, int f16 = 0);
// f16 = 1: fp16 planes of w * wscale[0] (wscale: device pointer to the power of two pow2_scale_launch chose; null = 1)
int pack_weights_launch(const float* w_hwio, bf16* wpack, int ksize, int cin, int cout, int cout_pad,
                        cudaStream_t s, int f16 = 0, const float* wscale = nullptr);
int bn_fold_launch(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                   float* scale, float* bias, int n, cudaStream_t s);
int maxpool_planes_launch(const bf16* in_hi, const bf16* in_lo, bf16* out_hi, bf16* out_lo, int B, int H, int W,
                          int C, cudaStream_t s, int stride = 2, int f16 = 0);
int pad_weights_launch(const float* w_hwio, float* out, int ksize, int cin, int cout, int cin_s, int cout_s, cudaStream_t s);
// reorg (space-to-depth 2) on 16-byte vectors; elem_bytes in {2,4}; out row pitch in elements.
int leaky_relu_launch(const float* in, float* out, size_t n, float alpha, cudaStream_t s);
int reorg_launch(const void* in, void* out, int B, int H, int W, int C, int stride, int elem_bytes,
                 long long out_ld, cudaStream_t s);

// ---- training-step kernels (y2_train.cu, y2_wgrad_tc.cu) ----
size_t bn_partial_bytes();
int bn_stats_launch(const float* z, size_t rows, int C, const float* gamma, const float* beta, float eps, float decay,
                    float* mean, float* inv, float* scale, float* bias, float* moving_mean, float* moving_var,
                    double* partial, cudaStream_t s);
// y16_*: optional second copy of y as fp16 planes (same pitch), the input format of the training forward's next conv
int bn_apply_launch(const float* z, const float* scale, const float* bias, bf16* y_hi, bf16* y_lo, size_t rows, int C,
                    long long ldy, cudaStream_t s, bf16* y16_hi = nullptr, bf16* y16_lo = nullptr);
int bn_bwd_reduce_launch(const float* z, const float* g, long long ldg, size_t rows, int C, const float* scale,
                         const float* bias, const float* mean, const float* inv, float* dgamma, float* dbeta, float* m1,
                         float* m2, double* partial, cudaStream_t s);
int bn_bwd_apply_launch(const float* z, const float* g, long long ldg, const float* scale, const float* bias,
                        const float* mean, const float* inv, const float* m1, const float* m2, bf16* dx_hi, bf16* dx_lo,
                        size_t rows, int C, cudaStream_t s);
int bias_grad_launch(const float* g, long long ldg, size_t rows, int C, float* dbias, double* partial, cudaStream_t s);
int split_planes_pad_launch(const float* src, long long ld, bf16* hi, bf16* lo, size_t rows, int C, int Cpad, cudaStream_t s);
int maxpool_bwd_launch(const float* gp, long long ldgp, const bf16* y_hi, const bf16* y_lo, float* g, int B, int H, int W,
                       int C, cudaStream_t s, int f16 = 0);
// pool layers without passthrough: forward z -> pooled planes; backward z + pooled gradient -> dgamma/dbeta + dx planes
int bn_apply_pool_launch(const float* z, const float* scale, const float* bias, bf16* p_hi, bf16* p_lo, int B, int H, int W, int C,
                         cudaStream_t s, bf16* p16_hi = nullptr, bf16* p16_lo = nullptr);
int bn_bwd_pool_launch(const float* z, const float* gp, long long ldgp, int B, int H, int W, int C, const float* scale, const float* bias,
                       const float* mean, const float* inv, float* dgamma, float* dbeta, float* m1, float* m2, bf16* dx_hi,
                       bf16* dx_lo, double* partial, float* gy_out, cudaStream_t s, int f16 = 0);
int maxpool_s1_bwd_launch(const float* gp, long long ldgp, const bf16* y_hi, const bf16* y_lo, float* g, int B, int H, int W, int C,
                          cudaStream_t s, int f16 = 0);
int repitch_planes_launch(const bf16* src_hi, const bf16* src_lo, bf16* dst_hi, bf16* dst_lo, size_t rows, int C, int Cpad, cudaStream_t s);
int compact_hwio_launch(const float* src, float* dst, int taps, int cin_s, int cout_s, int cin, int cout, cudaStream_t s);
int reorg_bwd_add_launch(const float* gr, long long ldr, float* g, int B, int H, int W, int C, cudaStream_t s);
int pack_dgrad_weights_launch(const float* w_hwio, bf16* out, int ksize, int cin, int cout, int cin_pad, int cout_pad,
                              cudaStream_t s);
// dW[k*k][Cin][Cout] = sum_p x_in[p+tap][cin] * dx[p][cout] on tcgen05 (MN-major operands, split planes)
int wgrad_tc_run(const bf16* x_planes, int B, int H, int W, int Cin, int ksize, const bf16* dx_planes, int Cout,
                 int dpitch, float* dw, int max_ctas, int num_sms, void* sk_ws, cudaStream_t stream);
int wgrad_check_watchdog();
// conv0 (Cin = 3) on the CUDA cores: raw forward (training) and weight gradient
int conv0_raw_launch(const float* x, const float* w_hwio, float* z, int B, int H, int W, cudaStream_t s);
int conv0_wgrad_launch(const float* x, const bf16* dx_hi, const bf16* dx_lo, float* dw, double* partial, int B, int H, int W,
                       cudaStream_t s);

// ---- optimizer step on the flat bucket (y2_optim.cu) ----
struct AdamChunk { unsigned long long off; unsigned int count; unsigned int tensor; };   // <= 16384 elements, inside one tensor
int adam_launch(const float* g, float* m, float* v, float* const* params_dev, const unsigned long long* tensor_start_dev,
                const AdamChunk* tab_dev, const int* first_chunk_dev, int nchunks, int ntensors, double* partial, float* scale,
                float alpha, float beta1, float beta2, float eps, float clip, cudaStream_t s);

// ---- pre / post steps (y2_prepost.cu) ----
size_t standardize_workspace_bytes(int B, size_t n);
int standardize_launch(const void* x, int elem_bytes, int B, size_t n, float* out, void* ws, cudaStream_t s);
int detections_launch(const float* conf, const float* xy_min, const float* xy_max, int B, int N, int C, float threshold, float sx, float sy,
                      int* count, int* box, int* cls, float* score, float* xywh, cudaStream_t s);

int transform_labels_launch(const int* ocls, const float* ocoord, const int* offsets, int B, int classes, int cw, int ch, float* mask,
                            float* prob, float* coords, float* oxy_min, float* oxy_max, float* areas, int* status, cudaStream_t s);

// ---- SIMT convs (y2_conv_simt.cu) ----
// conv0: 3x3, Cin=3 -> Cout=32, BN+leaky+2x2 maxpool fused, fp32 in, planes out.
int conv0_pool_launch(const float* x, const float* w_hwio, const float* scale, const float* bias, bf16* out_hi,
                      bf16* out_lo, int B, int H, int W, cudaStream_t s);
// conv0 on the tensor cores (y2_conv0_tc.cu): SIMT-built im2col tiles + tcgen05, same fused BN + leaky + pool and output planes
bool conv0_tc_applicable(int H, int W);
int conv0_tc_pool_launch(const float* x, const float* w_hwio, const float* scale, const float* bias, bf16* out_hi, bf16* out_lo, int B,
                         int H, int W, int num_sms, cudaStream_t s, int fast = 0);
int conv0_tc_check_watchdog();
}  // namespace y2
