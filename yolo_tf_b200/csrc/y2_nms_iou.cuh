// The float32 arithmetic of the NMS -- utils/postprocess.py:21-36 (`iou`) restated operation by operation, and the order-preserving
// float -> uint32 map of the sort keys.  A header of its own so that tests/host/nms_iou_harness.cu compiles THESE functions for the
// host and checks them against the numpy oracle on the CPU (tests/test_nms_iou_host.py): on the device every operation is an explicit
// round-to-nearest intrinsic (no FMA contraction), on the host the same operations are plain float32 (-ffp-contract=off).
// y2_nms.cu still calls __fmul_rn etc. directly where it needs them; only the functions below are shared.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#include <cuda_runtime.h>

namespace y2 {

#define Y2_NMS_FN __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define Y2_SUB(a, b) __fsub_rn(a, b)
#define Y2_ADD(a, b) __fadd_rn(a, b)
#define Y2_MUL(a, b) __fmul_rn(a, b)
#define Y2_DIV(a, b) __fdiv_rn(a, b)
#define Y2_BITS(v) __float_as_uint(v)
#else
#define Y2_SUB(a, b) ((a) - (b))
#define Y2_ADD(a, b) ((a) + (b))
#define Y2_MUL(a, b) ((a) * (b))
#define Y2_DIV(a, b) ((a) / (b))
static inline uint32_t y2_host_bits(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
#define Y2_BITS(v) y2_host_bits(v)
#endif

Y2_NMS_FN float iou_ref(float4 a, float4 b) {   // (xmin, ymin, xmax, ymax)
    const float a1 = Y2_MUL(Y2_SUB(a.z, a.x), Y2_SUB(a.w, a.y));
    const float a2 = Y2_MUL(Y2_SUB(b.z, b.x), Y2_SUB(b.w, b.y));
    const float iw = fmaxf(Y2_SUB(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float ih = fmaxf(Y2_SUB(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = Y2_MUL(iw, ih);
    const float d0 = Y2_SUB(Y2_ADD(a1, a2), inter);
    const float den = (d0 != d0) ? d0 : fmaxf(d0, 1e-10f);           // np.maximum propagates NaN (inf + inf - inf), fmaxf does not
    return Y2_DIV(inter, den);
}
Y2_NMS_FN float box_area(float4 a) { return Y2_MUL(Y2_SUB(a.z, a.x), Y2_SUB(a.w, a.y)); }
// iou_ref(a, b) >= thr with the areas precomputed.  Same float32 operations as iou_ref up to the divide (bit-identical
// inter and den), then a conservative filter when thr > 0 (quick): rn(inter/den) >= thr needs inter >= thr*(1-2^-24)*den,
// so anything below 0.999*thr*den is certainly no hit and skips the IEEE division; disjoint pairs (inter == 0) fall out
// here too.  NaNs fail the '<' and take the exact path.
Y2_NMS_FN bool iou_hit(float4 a, float aa, float4 b, float ba, float thr, float thr_lo, bool quick) {
    const float iw = fmaxf(Y2_SUB(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float ih = fmaxf(Y2_SUB(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = Y2_MUL(iw, ih);
    const float d0 = Y2_SUB(Y2_ADD(aa, ba), inter);
    if (d0 != d0) return false;                                       // the reference's np.maximum(nan, 1e-10) is nan: nan >= thr_iou is False
    const float den = fmaxf(d0, 1e-10f);
    if (quick && inter < Y2_MUL(thr_lo, den)) return false;
    return Y2_DIV(inter, den) >= thr;
}
// order-preserving map float -> uint32 (a > b  <=>  ford(a) > ford(b) for non-NaN a, b; -0 is canonicalised to +0 first)
Y2_NMS_FN uint32_t ford(float v) {
    const uint32_t u = Y2_BITS(v + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace y2
