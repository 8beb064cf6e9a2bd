// Operand arithmetic of the EXPERIMENTAL mixed-kind conv (y2_conv_mix.cu), host + device so that the CPU harness
// (tests/host/mix_prep_harness.cu) checks the very code the kernels run.
//
//   x*w ~= [ X16*W16 + X8*RW8 + RX8*W8 ] * unscale
//   X16 = fp16(x E16)    X8 = e4m3(x E8)    RX8 = e4m3((x E16 - X16) * ra)      ra = 4096 E8 / E16
//   W16 = fp16(w F16)    W8 = e4m3(w F8)    RW8 = e4m3((w F16 - W16) * rw)      rw = 4096 F8 / F16
//
// E16 F16 == E8 F8 4096, so the three products carry the same power-of-two factor and share one fp32 accumulator.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdint.h>

namespace y2 {

struct MixScales {
    float E16, E8, ra;       // activations
    float F16, F8, rw;       // weights
    float unscale;           // 1 / (E16 F16)
};

// amax_x, amax_w > 0 and finite.  ba, bw = the powers of two at or above them: x E16 <= 2^15 (fp16 max 65504),
// x E8 <= 2^8 (e4m3 max 448), residuals (<= 2^-12 of the value) land at <= 2^8 as well; w F16 <= 2^13.
static inline MixScales mix_scales(float amax_x, float amax_w) {
    const float ba = exp2f(ceilf(log2f(amax_x))), bw = exp2f(ceilf(log2f(amax_w)));
    MixScales s;
    s.E16 = 32768.f / ba; s.E8 = 256.f / ba; s.ra = 4096.f * s.E8 / s.E16;      // = 32
    s.F16 = 8192.f / bw;  s.F8 = 256.f / bw;  s.rw = 4096.f * s.F8 / s.F16;     // = 128
    s.unscale = 1.0f / (s.E16 * s.F16);
    return s;
}

__host__ __device__ __forceinline__ uint8_t mix_to_e4m3(float v) {
    return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}
// one element -> its three stored forms (s16, s8, rs) = (E16, E8, ra) or (F16, F8, rw)
__host__ __device__ __forceinline__ void mix_split(float v, float s16, float s8, float rs, __half* h16, uint8_t* q8, uint8_t* r8) {
    const __half h = __float2half_rn(v * s16);
    *h16 = h;
    *q8 = mix_to_e4m3(v * s8);
    *r8 = mix_to_e4m3((v * s16 - __half2float(h)) * rs);
}

}  // namespace y2
