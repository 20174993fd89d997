// Optimizer tail of the training step over the flat fp32 buckets: global gradient-norm clipping
// (torch.nn.utils.clip_grad_norm_, experiment/adaptive_scaling/train.py:468-472) fused into AdamW (torch.optim.AdamW,
// train.py:73-80,287-298,474-478).  Two HBM-bound passes over the parameters' storage: sum of squares of the gradients
// (fp64 accumulation, one atomic per block), then one kernel per bucket that reads param / grad / exp_avg / exp_avg_sq
// and writes param / exp_avg / exp_avg_sq.  The clip coefficient is computed on the device from the sum of squares, so
// the step has no host synchronisation.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
    __shared__ double red[8];
    double acc = 0.0;
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x4 + i);
        acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i] * (double)x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
             float beta1, float beta2, float eps, float weight_decay, float bias_corr1, float inv_sqrt_bias_corr2,
             const double* __restrict__ sumsq, float max_norm, float grad_scale) {
    float coef = grad_scale;
    if (sumsq != nullptr && max_norm > 0.f) {
        const float total = (float)sqrt(*sumsq) * grad_scale;        // norm of the (scaled) gradient
        const float c = max_norm / (total + 1e-6f);                  // clip_grad_norm_: clamped to 1
        coef *= c < 1.f ? c : 1.f;
    }
    const float step = lr / bias_corr1;
    const float decay = 1.f - lr * weight_decay;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bias_corr2 + eps;
        p[i] = p[i] * decay - step * (mi / denom);
    }
}

}  // namespace

extern "C" {

// out[0] += sum_i x[i]^2   (fp64; caller zeroes `out` once per step and passes every bucket)
int vkocr_sumsq_f32(const float* x, long long n, double* out, void* stream) {
    VK_REQUIRE(x && out, VKOCR_BAD_ARGUMENT, "sumsq_f32: null argument");
    VK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, VKOCR_BAD_ALIGN, "sumsq_f32: buffer not 16-byte aligned");
    if (n == 0) return VKOCR_OK;
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, out);
    VK_CHECK_LAUNCH("sumsq_kernel");
    return VKOCR_OK;
}

// One AdamW step (decoupled weight decay, torch.optim.AdamW semantics, amsgrad off) on a flat fp32 range.  The gradient is
// first multiplied by grad_scale and, when `sumsq` (device, the squared global norm of the UNSCALED gradient) is given and
// max_norm > 0, by min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6)).  bias_corr1 = 1 - beta1^t, bias_corr2 = 1 - beta2^t.
int vkocr_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, const double* sumsq,
                     float max_norm, float grad_scale, void* stream) {
    VK_REQUIRE(param && grad && exp_avg && exp_avg_sq, VKOCR_BAD_ARGUMENT, "adamw_step: null argument");
    VK_REQUIRE(bias_corr1 > 0.f && bias_corr2 > 0.f, VKOCR_BAD_ARGUMENT, "adamw_step: bias corrections must be positive");
    if (n == 0) return VKOCR_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    adamw_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bias_corr1, 1.f / sqrtf(bias_corr2), sumsq, max_norm,
        grad_scale);
    VK_CHECK_LAUNCH("adamw_kernel");
    return VKOCR_OK;
}

}  // extern "C"
