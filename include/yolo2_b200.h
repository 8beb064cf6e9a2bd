/* libyolo2_b200.so -- C ABI of the B200-native YOLOv2 / Darknet-19 detection hot path.
 *
 * The reference (ruiminshen/yolo-tf, pure Python + TensorFlow 1.0) has no FFI: its seam is the
 * Python call surface that train.py / detect.py drive.  Each entry point below is what a ctypes
 * binding for that surface calls; the reference interface it replaces is cited per function
 * (paths relative to the reference repository).  The Python mirror lives in yolo_tf_b200/.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller (e.g. torch.Tensor.data_ptr()),
 *     16-byte aligned, float32 NHWC / row-major unless stated; `hparam` and scalars are host values;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is asynchronous
 *     on it; the library never synchronises the device except in y2_create / y2_destroy;
 *   - return 0 on success, < 0 on error; y2_last_error() gives the thread-local message;
 *   - no CPU fallback exists: every call either launches sm_100a kernels or fails.
 */
#ifndef YOLO2_B200_H
#define YOLO2_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct y2_handle y2_handle;

const char* y2_last_error(void);
int y2_version(void);

/* ---- network handle ------------------------------------------------------------------------
 * Replaces the graph + variables that `Builder.__call__` builds through
 * `inference.darknet(net, classes, num_anchors, training, center)` (model/yolo2/__init__.py:109-112,
 * model/yolo2/inference.py:61-120).  One handle per (device, network); not thread-safe per handle. */
int y2_create(y2_handle** out, int device, int classes, int num_anchors);
void y2_destroy(y2_handle* h);

/* The `inference` function is string-dispatched from the INI config (`getattr(inference, config.get(section,
 * 'inference'))`, model/yolo2/__init__.py:107; config/yolo2/{darknet,tiny}-*.ini).  y2_create_net selects it:
 *   Y2_ARCH_DARKNET  `darknet()`  model/yolo2/inference.py:61-120  (y2_create == this)
 *   Y2_ARCH_TINY     `tiny()`     model/yolo2/inference.py:25-50: 16..512-channel 3x3 convs, five 2x2/2 max-pools, one
 *                    2x2 stride-1 SAME max-pool (:42), two 1024-channel 3x3 convs, linear 1x1 conv.  Inference and the
 *                    training step (the reference trains it through the same train.py / TF autodiff).
 * Layers whose channel count is below 32 (tiny conv0: 16) are stored zero-padded to 32; y2_layer_info,
 * y2_load_weights, y2_get_activation and the gradient bucket (y2_param_offsets) speak the logical (variable) shapes;
 * y2_train_get_tensor / y2_train_probe (test hooks) the stored ones. */
#define Y2_ARCH_DARKNET 0
#define Y2_ARCH_TINY 1
int y2_create_net(y2_handle** out, int device, int classes, int num_anchors, int arch);

/* Number of conv layers (22) and the per-layer geometry, in graph order conv0..conv20, conv (final). */
int y2_num_layers(const y2_handle* h);
int y2_layer_info(const y2_handle* h, int layer, int* ksize, int* cin, int* cout, int* has_bn);

/* Load one layer's variables (device pointers).  `w_hwio` is the TF layout [k][k][cin][cout]
 * (`yolo2_darknet/conv{i}/weights`); BN layers take gamma/beta/moving_mean/moving_variance
 * (`.../BatchNorm/*`, eps = 1e-5, inference.py:63) and bias = NULL; the final layer takes `bias`
 * (`yolo2_darknet/conv/biases`, inference.py:118) and NULL BN pointers.  The library folds BN with
 * TF's arithmetic (inv = rsqrt(var+eps)*gamma; y = x*inv + (beta - mean*inv)) and re-packs the
 * weights into split bf16 planes for the tensor cores.  Replaces slim.assign_from_checkpoint_fn
 * (detect.py:104-106). */
int y2_load_weights(y2_handle* h, int layer, const float* w_hwio, const float* gamma, const float* beta,
                    const float* moving_mean, const float* moving_variance, const float* bias, void* stream);

/* Activation workspace needed by y2_darknet_forward for a [B,H,W,3] batch (H, W multiples of 32,
 * the assert of utils/__init__.py:52-56). */
size_t y2_workspace_bytes(const y2_handle* h, int B, int H, int W);

/* x [B,H,W,3] -> out [B,H/32,W/32,A*(5+C)].  Inference mode (BN moving statistics):
 * `builder(image)` / `darknet(..., training=False)`, detect.py:101-102.
 * precision: 0 = split-bf16 x3 (fp32-grade, default), 1 = single bf16 pass (fast, ~1e-2). */
int y2_darknet_forward(y2_handle* h, const float* x, int B, int H, int W, float* out, void* ws, size_t ws_bytes,
                       int precision, void* stream);

/* Per-layer device timing of y2_darknet_forward (CUDA events on the launching stream): conv_ms[i] =
 * conv kernel(s) of layer i (incl. the split-K finishing pass), post_ms[i] = its max-pool / reorg
 * passes; both host arrays of y2_num_layers() floats, valid for the last forward once it finished.
 * y2_launch_count() = kernels launched by this library in this process (all entry points).
 * The reference has no profiler; these feed bench.py's roofline numbers. */
int y2_set_profiling(y2_handle* h, int enable);
int y2_get_layer_ms(y2_handle* h, float* conv_ms, float* post_ms);
unsigned long long y2_launch_count(void);

/* ---- training step (BASELINE config 3) ------------------------------------------------------
 * `builder(batch, training=True)` (train.py:109-110): forward with BATCH statistics
 * (slim.batch_norm(is_training=True): population variance over (B,H,W), moving averages updated with
 * decay 0.999 as slim's UPDATE_OPS do), activations kept in `ws` for the backward.
 * y2_darknet_backward = the tf.gradients part of slim.learning.create_train_op (train.py:127-129):
 * dnet = d(total_loss)/d(out) (from y2_loss_fwd_bwd) -> gradients of every variable in ONE flat float32
 * bucket of y2_param_count() elements, laid out per layer (conv0..conv20, conv) as
 * [weights HWIO | gamma | beta] or [weights | biases] (y2_param_offsets) -- the unit of the single NCCL
 * all-reduce per step.  y2_get_bn_state reads back gamma/beta/moving statistics (device pointers).
 * Numerics: the training forward's convs run on fp16 split planes (22 significand bits, weights pre-scaled by a per-layer
 * power of two) with accumulation chains of "train_kcap" (16) k-blocks, so that the network output stays within 1e-4 of
 * float64 although batch-statistics BN amplifies every layer's error; y2_set_option(h, "train_f16", 0) selects the
 * bf16 planes of the inference path (range-safe, 16 bits, ~2e-4 at the network output). */
size_t y2_train_workspace_bytes(const y2_handle* h, int B, int H, int W);
int y2_darknet_forward_train(y2_handle* h, const float* x, int B, int H, int W, float* out, void* ws, size_t ws_bytes,
                             void* stream);
int y2_darknet_backward(y2_handle* h, const float* dnet, float* flat_grads, void* stream);
size_t y2_param_count(const y2_handle* h);
int y2_param_offsets(const y2_handle* h, int layer, size_t* w_off, size_t* gamma_off, size_t* beta_or_bias_off);
int y2_get_bn_state(y2_handle* h, int layer, float* gamma, float* beta, float* moving_mean, float* moving_variance,
                    void* stream);

/* Optimizer step -- the apply-gradients part of slim.learning.create_train_op (train.py:127-129) with the default
 * optimizer tf.train.AdamOptimizer(lr, beta1, beta2, epsilon) (train.py:70-72, config.ini [optimizer_adam]) on the flat
 * bucket y2_darknet_backward filled (after the all-reduce).  TF-1.0 arithmetic: alpha = lr*sqrt(1-beta2^t)/(1-beta1^t);
 * m += (g-m)*(1-beta1); v += (g*g-v)*(1-beta2); var -= m*alpha/(sqrt(v)+epsilon); clip_norm > 0 first rescales every
 * tensor's gradient by clip*min(rsqrt(sum g*g), 1/clip) (tf.clip_by_norm, --gradient_clip train.py:159).
 * params: HOST array of y2_num_param_tensors() DEVICE pointers in bucket order (per layer weights, gamma, beta | biases),
 * updated in place; m, v: device buffers of y2_param_count() floats (zero before the first step); t = 1, 2, ...
 * The learning-rate schedule (tf.train.exponential_decay, train.py:120) is the caller's: pass the decayed rate. */
int y2_num_param_tensors(const y2_handle* h);
size_t y2_adam_workspace_bytes(const y2_handle* h);
int y2_adam_step(y2_handle* h, const float* flat_grads, float* m, float* v, float* const* params, int ntensors,
                 float learning_rate, float beta1, float beta2, float epsilon, long long t, float clip_norm, void* ws,
                 size_t ws_bytes, void* stream);

/* utils/preprocess.py:23-25 per_image_standardization, batched on the device (detect.py:62 applies it to the resized uint8
 * image cast to float32): out[b] = (x[b] - mean_b) / max(std_b, 1/sqrt(n)), mean and POPULATION std over all
 * n = H*W*3 elements of image b.  x: uint8 (elem_bytes 1; the cast is fused, a quarter of the H2D bytes) or float32 (4). */
size_t y2_standardize_workspace_bytes(int B, size_t n_per_image);
int y2_per_image_standardization(const void* x, int elem_bytes, int B, size_t n_per_image, float* out, void* ws,
                                 size_t ws_bytes, void* stream);

/* detect.py:65 `_image.resize((width, height))`: Pillow's Image.resize on 8-bit channels, bit for bit -- resample 3 = BICUBIC
 * (Pillow's default since 7.0: a = -0.5, antialiased when shrinking, 22-bit fixed-point weights, horizontal pass into an 8-bit
 * intermediate then vertical) or 0 = NEAREST (its default before).  src [in_h][in_w][channels], dst [out_h][out_w][channels],
 * uint8 device pointers.  The algorithm is Pillow's (third-party, unpinned by the reference); the restatement is verified
 * against Pillow itself on the CPU.  NOT YET RUN ON A GPU (written after the round-1 GPU budget was spent). */
size_t y2_resize_workspace_bytes(int in_h, int in_w, int out_h, int out_w, int channels, int resample);
int y2_resize_u8(const uint8_t* src, int in_h, int in_w, int channels, uint8_t* dst, int out_h, int out_w, int resample, void* ws,
                 size_t ws_bytes, void* stream);

/* detect.py:72-87 after non_max_suppress: per box index = argmax_c conf (first maximum), kept iff conf[index] > threshold;
 * kept boxes are appended in box-index order: box[b][i], cls[b][i], score[b][i], xywh[b][i] = (xy_min*scale,
 * (xy_max-xy_min)*scale) with scale = image size / cells (detect.py:72), count[b] = number kept.  All device, [B][N]. */
int y2_detections(const float* conf, const float* xy_min, const float* xy_max, int B, int N, int C, float threshold,
                  float scale_x, float scale_y, int* count, int* box, int* cls, float* score, float* xywh, void* stream);

/* utils/data/__init__.py:112-145 transform_labels, batched on the device (train.py feeds it per image through
 * tf.py_func, utils/data/__init__.py:148-150): ragged object lists -> the six label tensors Objectives consumes.
 * objects_class int32 [T], objects_coord float32 [T][4] = (xmin, ymin, xmax, ymax) normalised to the image, image b owns
 * objects offsets[b] .. offsets[b+1]-1 (offsets int32 [B+1], device).  Outputs (device, fully overwritten):
 * mask [B][cells][1], prob [B][cells][1][classes], coords [B][cells][1][4] = (offset_x, offset_y, sqrt w, sqrt h),
 * offset_xy_min / offset_xy_max [B][cells][1][2], areas [B][cells][1].  One box per cell: the LAST object landing in a cell
 * wins, class bits accumulate (numpy fancy-index assignment).  float32 arithmetic in the reference's order.
 * status (nullable, int32 [B]): bit 0 = an object's cell or class index is out of range (IndexError in the reference; the
 * object is skipped), bit 1 = negative width/height (`assert np.all(wh >= 0)`, :142). */
int y2_transform_labels(const int32_t* objects_class, const float* objects_coord, const int32_t* offsets, int B, int classes,
                        int cell_width, int cell_height, float* mask, float* prob, float* coords, float* offset_xy_min,
                        float* offset_xy_max, float* areas, int32_t* status, void* stream);

/* Test hooks of the training step (per-layer "teacher-forced" backward parity): y2_train_probe arms the next
 * y2_darknet_backward to copy dL/dy of `layer` (dense [M][cout]) and dL/d(input of layer) (dense [M][cin]);
 * y2_train_get_tensor reads saved forward state: kind 0 raw conv output, 1 activation, 2 pooled, 3 concat. */
int y2_train_probe(y2_handle* h, int layer, float* gy_out, float* gin_out);
int y2_train_get_tensor(y2_handle* h, int kind, int layer, float* dst, void* stream);

/* Copy layer `layer`'s post-activation output of the LAST forward (pre-pool) as float32 NHWC into
 * `dst` -- the tensors `yolo2_darknet/conv{i}/...` that the reference exposes by name for summaries
 * (train.py:31-67); used by the per-layer parity tests.  pooled != 0 returns the max-pooled tensor. */
int y2_get_activation(y2_handle* h, int layer, int pooled, float* dst, void* stream);

/* Options.  "fuse_pool" (default 1): fuse the 2x2/2 max-pools (inference.py:74,83,96) into the conv epilogues when
 * the batch and extent admit the spatial tiling (e.g. batch 32 at 416/608); 0 keeps the separate pool pass and
 * materialises the un-pooled activations for y2_get_activation.  "halo" (default 1): run the 32-channel 3x3 layer
 * (conv1, inference.py:75) from one halo tile per output tile with its weights resident in shared memory instead of nine
 * im2col fetches; 0 selects the im2col path (bit-identical results, used by the equivalence tests).
 * "pair" (default 1): CTA-pair (cta_group::2) convs, 0 off / 1 the 3x3 layers with 256-wide N tiles / 2 every eligible layer.
 * "conv0_tc" (default 2): conv0 on the tensor cores.  "keep_activations" (default 0): 1 gives every layer's output its own
 * workspace slot so that y2_get_activation can read it after the forward (changes y2_workspace_bytes); by default the
 * outputs alternate between two arenas.  "train_f16" (default 1), "train_kcap" (default 16): numerics of the training
 * forward, see the training step above. */
int y2_set_option(y2_handle* h, const char* key, int value);

/* One conv (+scale/bias +leaky) on float32 NHWC tensors through the same tcgen05 kernel the
 * network uses (splits operands on the fly).  Diagnostic / test entry point.
 * block_n = 0 picks the tile width; max_ctas = 0 uses every SM (smaller values change how the
 * stream-K scheduler cuts tiles across CTAs -- used by tests to exercise the partial hand-off).
 * precision: 0 / 1 as y2_darknet_forward; 2 = fp16 split planes for both operands (the training forward's format);
 * 3 / 4 = mixed bf16 x fp16 operands, kept as a hardware probe only: tcgen05 kind::f16 faults on them (illegal instruction). */
int y2_conv2d(const float* x, int B, int H, int W, int cin, const float* w_hwio, int ksize, int cout,
              const float* scale, const float* bias, int leaky, float* y, int precision, int block_n, int max_ctas,
              void* stream);

/* Weight gradient of one conv on float32 NHWC tensors through the tcgen05 wgrad kernel the training step uses:
 * dw[k][k][cin][cout] = sum over pixels of x[p + tap][cin] * dy[p][cout] (SAME padding).  Diagnostic / test entry. */
int y2_conv2d_wgrad(const float* x, int B, int H, int W, int cin, const float* dy, int ksize, int cout, float* dw,
                    int max_ctas, void* stream);

/* ---- leaky_relu -- model/yolo/function.py:21-24 (`leaky_relu(inputs, alpha=.1)` = max(x, alpha*x)), float32, n elements; out may
 * alias in.  Inside the network the activation is fused into every conv epilogue; this is the standalone op. */
int y2_leaky_relu(const float* in, size_t n, float alpha, float* out, void* stream);

/* ---- reorg -- model/yolo2/function.py:22-29 (`reorg(net, stride=2)`), float32 NHWC.
 * out[b, y, x, (dy*stride+dx)*C + c] = in[b, stride*y+dy, stride*x+dx, c]. */
int y2_reorg(const float* in, int B, int H, int W, int C, int stride, float* out, void* stream);

/* ---- head decode -- `Model(net, classes, anchors, training)`, model/yolo2/__init__.py:28-59.
 * net [B,Hc,Wc,A*(5+C)]; anchors [A][2] float32 (cell units).  Any output pointer may be NULL.
 * Box index n = cell*A + anchor; all outputs are [B, cells*A, ...]. */
typedef struct y2_head_outputs {
    float* conf;          /* [B,N,C]  iou*prob             (:56) */
    float* xy_min;        /* [B,N,2]  cell units           (:54) */
    float* xy_max;        /* [B,N,2]                       (:55) */
    float* iou;           /* [B,N]    sigmoid(k=0)         (:37) */
    float* prob;          /* [B,N,C]  softmax              (:42) */
    float* wh;            /* [B,N,2]  exp * anchor         (:40) */
    float* areas;         /* [B,N]                         (:43) */
    float* xy;            /* [B,N,2]                       (:53) */
    float* offset_xy;     /* [B,N,2]                       (:38) */
    float* offset_xy_min; /* [B,N,2]                       (:45) */
    float* offset_xy_max; /* [B,N,2]                       (:46) */
    float* coords;        /* [B,N,4]  (off_x, off_y, sqrt(w01), sqrt(h01))  (:49) */
    float* wh01;          /* [B,N,2]                       (:47) */
} y2_head_outputs;
int y2_head_decode(const float* net, int B, int Hc, int Wc, int A, int C, const float* anchors,
                   const y2_head_outputs* outs, void* stream);

/* ---- loss -- `Objectives(model, mask, prob, coords, offset_xy_min, offset_xy_max, areas)`
 * (model/yolo2/__init__.py:62-94) + the weighting of `Builder.create_objectives` (:114-119).
 * Labels are the 6 tensors of utils/data/__init__.py:112-145 batched: mask [B,cells], prob [B,cells,C],
 * coords [B,cells,4], offset_xy_min/max [B,cells,2], areas [B,cells].
 * hparam (HOST) = {prob, iou_best, iou_normal, coords} ([yolo2_hparam], config.ini:98-102).
 * objectives (device, 4 floats, same order): the UNWEIGHTED objectives the reference stores in
 * `builder.objectives[key]`.  dnet (nullable) receives d(sum_k hparam_k * objective_k)/d(net). */
size_t y2_loss_workspace_bytes(int B, int Hc, int Wc);
int y2_loss_fwd_bwd(const float* net, int B, int Hc, int Wc, int A, int C, const float* anchors, const float* mask,
                    const float* prob, const float* coords, const float* offset_xy_min, const float* offset_xy_max,
                    const float* areas, const float hparam[4], float* objectives, float* dnet, void* ws,
                    size_t ws_bytes, void* stream);

/* ---- NMS -- `non_max_suppress(conf, xy_min, xy_max, threshold, threshold_iou)`
 * (utils/postprocess.py:39-51, iou :21-36), batched over images.  conf [B,N,C] is modified IN PLACE
 * exactly as the reference mutates its argument.  order_out (nullable, int32 [B,N]) receives the
 * order of the list the reference returns.  status_out (nullable, int32 [B]) is set to 1 for images
 * on which the reference's asserts (NaN / xy_min > xy_max, postprocess.py:22-27) would fire. */
size_t y2_nms_workspace_bytes(int B, int N, int C);
int y2_nms(float* conf, const float* xy_min, const float* xy_max, int B, int N, int C, float threshold,
           float threshold_iou, int32_t* order_out, int32_t* status_out, void* ws, size_t ws_bytes, void* stream);

/* The tcgen05 pipelines bound every barrier wait; if one ever expires the kernel drains and records
 * it.  Call with the device idle (after a synchronize): 0 = clean, < 0 = a watchdog fired (message
 * in y2_last_error()).  Plays the role of tf.check_numerics-style runtime guards (detect.py:70). */
int y2_check_async_errors(void);

#ifdef __cplusplus
}
#endif
#endif /* YOLO2_B200_H */
