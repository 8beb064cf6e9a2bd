// Optimizer step on the flat gradient bucket:
//   tf.train.AdamOptimizer(lr, beta1, beta2, epsilon)            <- train.py:70-72 (get_optimizer), config.ini [optimizer_adam]
//   slim.learning.create_train_op(..., clip_gradient_norm=clip)  <- train.py:127-129: per-tensor tf.clip_by_norm, then apply
// TF-1.0 arithmetic restated (training_ops ApplyAdam functor, clip_ops.clip_by_norm):
//   alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)               (host, t = 1, 2, ...)
//   m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + epsilon)
//   clip: g *= clip * min(rsqrt(sum(g*g)), 1/clip) per tensor
// HBM-bound: 28 bytes per parameter (g, m, v, var read; m, v, var written); 67.2 M parameters = 1.88 GB per step.
// The bucket is cut into chunks of <= 16384 elements that never straddle a tensor; one block per chunk.
#include "y2_internal.h"

namespace y2 {

__global__ void __launch_bounds__(256)
adam_sumsq_kernel(const float* __restrict__ g, const AdamChunk* __restrict__ tab, double* __restrict__ partial) {
    const AdamChunk ch = tab[blockIdx.x];
    const float* src = g + ch.off;
    float s = 0.f;
    for (unsigned i = threadIdx.x; i < ch.count; i += 256) { const float t = __ldg(src + i); s = fmaf(t, t, s); }
    __shared__ double sm[256];
    sm[threadIdx.x] = (double)s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {                     // fixed tree: deterministic
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

// one warp per tensor: sum the tensor's chunk partials (lane-strided, then a fixed shuffle tree) -> clip scale
__global__ void adam_clip_scale_kernel(const double* __restrict__ partial, const int* __restrict__ first_chunk, int ntensors, float clip,
                                       float* __restrict__ scale) {
    const int t = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= ntensors) return;
    double s = 0.0;
    for (int c = first_chunk[t] + lane; c < first_chunk[t + 1]; c += 32) s += partial[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        // tf.clip_by_norm: t * clip_norm * minimum(rsqrt(sum(t*t)), 1/clip_norm), float32
        const float l2inv = rsqrtf((float)s);
        scale[t] = clip * fminf(l2inv, 1.0f / clip);
    }
}

__global__ void __launch_bounds__(256)
adam_apply_kernel(const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* const* __restrict__ params,
                  const unsigned long long* __restrict__ tensor_start, const AdamChunk* __restrict__ tab,
                  const float* __restrict__ scale, float alpha, float beta1, float beta2, float eps) {
    const AdamChunk ch = tab[blockIdx.x];
    float* p = params[ch.tensor] + (ch.off - tensor_start[ch.tensor]);
    const float gs = scale ? scale[ch.tensor] : 1.0f;
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const float* gg = g + ch.off;
    float* mm = m + ch.off;
    float* vv = v + ch.off;
    auto one = [&](float gi, float& mi, float& vi, float& pi) {
        gi *= gs;
        mi += (gi - mi) * omb1;
        vi += (gi * gi - vi) * omb2;
        pi -= (mi * alpha) / (sqrtf(vi) + eps);
    };
    const bool vec = ((ch.off & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
    if (vec) {
        const unsigned n4 = ch.count / 4;
        for (unsigned i = threadIdx.x; i < n4; i += 256) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gg) + i);
            float4 m4 = reinterpret_cast<float4*>(mm)[i], v4 = reinterpret_cast<float4*>(vv)[i], p4 = reinterpret_cast<float4*>(p)[i];
            one(g4.x, m4.x, v4.x, p4.x); one(g4.y, m4.y, v4.y, p4.y); one(g4.z, m4.z, v4.z, p4.z); one(g4.w, m4.w, v4.w, p4.w);
            reinterpret_cast<float4*>(mm)[i] = m4; reinterpret_cast<float4*>(vv)[i] = v4; reinterpret_cast<float4*>(p)[i] = p4;
        }
        for (unsigned i = n4 * 4 + threadIdx.x; i < ch.count; i += 256) one(gg[i], mm[i], vv[i], p[i]);
    } else {
        for (unsigned i = threadIdx.x; i < ch.count; i += 256) one(gg[i], mm[i], vv[i], p[i]);
    }
}

int adam_launch(const float* g, float* m, float* v, float* const* params_dev, const unsigned long long* tensor_start_dev,
                const AdamChunk* tab_dev, const int* first_chunk_dev, int nchunks, int ntensors, double* partial, float* scale,
                float alpha, float beta1, float beta2, float eps, float clip, cudaStream_t s) {
    if (clip > 0.f) {
        adam_sumsq_kernel<<<nchunks, 256, 0, s>>>(g, tab_dev, partial);
        Y2_CUDA(cudaGetLastError());
        note_launch();
        adam_clip_scale_kernel<<<(ntensors + 7) / 8, 256, 0, s>>>(partial, first_chunk_dev, ntensors, clip, scale);
        Y2_CUDA(cudaGetLastError());
        note_launch();
    }
    adam_apply_kernel<<<nchunks, 256, 0, s>>>(g, m, v, params_dev, tensor_start_dev, tab_dev, clip > 0.f ? scale : nullptr, alpha, beta1,
                                               beta2, eps);
    Y2_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace y2
