// Stand-alone primitive losses of vkit_open_model.loss_function on flat fp32 tensors (pred, gt, optional mask):
//   focal-with-logits (torchvision sigmoid_focal_loss; focal_with_logits.py:18-47), dice (dice.py:17-35),
//   L1 / smooth-L1 (l1.py:19-47), L2 (l2.py:18-34), weight-adaptive heat-map regression
//   (weight_adaptive_heatmap_regression.py:18-33), soft-label cross entropy over a small class axis
//   (cross_entropy_with_logits.py:16-19) and the hard-negative-mining BCE (weighted_bce_with_logits.py:18-54,
//   device-side radix select of the k-th largest negative loss, no host round trip).
// HBM-bound: one grid-stride reduction pass, block partials via warp shuffles, one fp64 atomic per block/quantity,
// a one-thread finalize that leaves the loss and the backward coefficients on the device.
#include "common.cuh"

namespace {

constexpr float EPS = 1e-6f;

enum Kind { K_FOCAL = 0, K_DICE = 1, K_L1 = 2, K_SMOOTH_L1 = 3, K_L2 = 4, K_WAHR = 5 };

__device__ __forceinline__ float bce_logits(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// value and d(value)/d(pred) of the un-reduced loss
__device__ __forceinline__ void pointwise(int kind, float x, float g, float p0, float p1, float* val, float* grad) {
    switch (kind) {
        case K_FOCAL: {
            const float p = sigm(x), ce = bce_logits(x, g);
            const float pt = p * g + (1.f - p) * (1.f - g), omp = 1.f - pt;
            const float at = p0 >= 0.f ? p0 * g + (1.f - p0) * (1.f - g) : 1.f;
            const float w = powf(omp, p1);
            *val = at * ce * w;
            // d/dx [ce * omp^gamma]: dce = p - g ; d(omp)/dx = -(2g-1) p(1-p)
            const float dw = (p1 == 0.f) ? 0.f : p1 * powf(omp, p1 - 1.f) * (-(2.f * g - 1.f)) * p * (1.f - p);
            *grad = at * ((p - g) * w + ce * dw);
            break;
        }
        case K_L1: {
            const float d = x - g;
            *val = fabsf(d);
            *grad = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            break;
        }
        case K_SMOOTH_L1: {
            const float d = x - g, a = fabsf(d);
            if (a < p0) { *val = 0.5f * d * d / p0; *grad = d / p0; }
            else { *val = a - 0.5f * p0; *grad = d > 0.f ? 1.f : -1.f; }
            break;
        }
        case K_L2: {
            const float d = x - g;
            *val = d * d;
            *grad = 2.f * d;
            break;
        }
        case K_WAHR: {
            const float soft = powf(g, p0);
            const float w = soft * (1.f - x) + (1.f - soft) * x;
            const float d = x - g;
            *val = w * d * d;
            *grad = (1.f - 2.f * soft) * d * d + 2.f * w * d;
            break;
        }
        default:
            *val = 0.f; *grad = 0.f;
    }
}

__device__ __forceinline__ void block_accumulate(const float* vals, int n, double* out, float* scratch) {
    for (int i = 0; i < n; ++i) {
        const float s = vk_block_sum(vals[i], scratch);
        if (threadIdx.x == 0 && s != 0.f) atomicAdd(out + i, (double)s);
    }
}

// sums: non-dice [0] sum(loss*mask) [1] sum(mask);  dice [0] sum(p*g) [1] sum(p) [2] sum(g)  (p, g already masked)
__global__ void __launch_bounds__(256)
pw_reduce_kernel(int kind, int pre_sigmoid, const float* __restrict__ pred, const float* __restrict__ gt,
                 const float* __restrict__ mask, long long n, float p0, float p1, double* __restrict__ sums) {
    __shared__ float scratch[33];
    float v[3] = {0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f;
        const float x = pre_sigmoid ? sigm(pred[i]) : pred[i];
        if (kind == K_DICE) {
            const float p = x * m, g = gt[i] * m;
            v[0] += p * g; v[1] += p; v[2] += g;
        } else {
            float val, grad;
            pointwise(kind, x, gt[i], p0, p1, &val, &grad);
            v[0] += val * m; v[1] += m;
        }
    }
    block_accumulate(v, 3, sums, scratch);
}

// coef: [0] loss; non-dice [1] 1/denominator; dice [1] I [2] U
__global__ void pw_finalize_kernel(int kind, const double* __restrict__ sums, double n, int has_mask, float* __restrict__ coef) {
    if (kind == K_DICE) {
        const double I = sums[0], U = sums[1] + sums[2] + (double)EPS;
        coef[0] = (float)(1.0 - 2.0 * I / U);
        coef[1] = (float)I;
        coef[2] = (float)U;
    } else {
        const double den = has_mask ? sums[1] + (double)EPS : n;
        coef[0] = (float)(sums[0] / den);
        coef[1] = (float)(1.0 / den);
        coef[2] = 0.f;
    }
}

__global__ void __launch_bounds__(256)
pw_bwd_kernel(int kind, int pre_sigmoid, const float* __restrict__ pred, const float* __restrict__ gt,
              const float* __restrict__ mask, long long n, float p0, float p1, const float* __restrict__ coef,
              const float* __restrict__ gout, float* __restrict__ dpred) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = mask ? mask[i] : 1.f;
    const float x = pre_sigmoid ? sigm(pred[i]) : pred[i];
    const float chain = pre_sigmoid ? x * (1.f - x) : 1.f;
    const float go = gout[0] * chain;
    if (kind == K_DICE) {
        const float g = gt[i] * m, I = coef[1], U = coef[2];
        dpred[i] = go * (-2.f * (g * U - I) / (U * U)) * m;
    } else {
        float val, grad;
        pointwise(kind, x, gt[i], p0, p1, &val, &grad);
        dpred[i] = go * grad * m * coef[1];
    }
}

// ---- soft-label cross entropy: pred/gt (outer, C, inner), class axis C <= 32, mean over outer*inner ----
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long outer, int C, long long inner,
                 double* __restrict__ sums) {
    __shared__ float scratch[33];
    float v = 0.f;
    const long long total = outer * inner;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long o = i / inner, in = i % inner;
        const float* p = pred + o * C * inner + in;
        const float* g = gt + o * C * inner + in;
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, p[c * inner]);
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(p[c * inner] - mx);
        const float lse = mx + logf(se);
        for (int c = 0; c < C; ++c) v -= g[c * inner] * (p[c * inner] - lse);
    }
    block_accumulate(&v, 1, sums, scratch);
}
__global__ void ce_finalize_kernel(const double* __restrict__ sums, double n, float* __restrict__ coef) {
    coef[0] = (float)(sums[0] / n);
    coef[1] = (float)(1.0 / n);
}
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long outer, int C, long long inner,
              const float* __restrict__ coef, const float* __restrict__ gout, float* __restrict__ dpred) {
    const long long total = outer * inner;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long o = i / inner, in = i % inner;
    const float* p = pred + o * C * inner + in;
    const float* g = gt + o * C * inner + in;
    float* d = dpred + o * C * inner + in;
    float mx = -INFINITY, st = 0.f;
    for (int c = 0; c < C; ++c) { mx = fmaxf(mx, p[c * inner]); st += g[c * inner]; }
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(p[c * inner] - mx);
    const float k = gout[0] * coef[1];
    for (int c = 0; c < C; ++c) d[c * inner] = k * (expf(p[c * inner] - mx) / se * st - g[c * inner]);
}

// ---- hard-negative-mining BCE -------------------------------------------------------------------------------------
// state (unsigned long long[8]): [0] #pos, [1] #neg, [2] k, [3] prefix bits, [4] remaining rank, [5] threshold bits
// hist: 256 unsigned long long.   sums (double[4]): [0] pos loss sum, [1] sum of neg losses > thr, [2] count > thr
__global__ void __launch_bounds__(256)
bce_count_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask, long long n,
                 unsigned long long* __restrict__ state, double* __restrict__ sums) {
    __shared__ float scratch[33];
    float v[3] = {0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f;
        // reference: positive_mask = (gt*mask).byte(), negative_mask = ((1-gt)*mask).byte()  -> truncation to uint8
        const float pm = (float)(unsigned char)(int)(gt[i] * m), nm = (float)(unsigned char)(int)((1.f - gt[i]) * m);
        const float l = bce_logits(pred[i], gt[i]);
        v[0] += pm != 0.f ? 1.f : 0.f;
        v[1] += nm != 0.f ? 1.f : 0.f;
        v[2] += l * pm;
    }
    for (int q = 0; q < 3; ++q) {
        const float s = vk_block_sum(v[q], scratch);
        if (threadIdx.x == 0 && s != 0.f) {
            if (q < 2) atomicAdd(state + q, (unsigned long long)(s + 0.5f));
            else atomicAdd(sums, (double)s);
        }
    }
}
__global__ void bce_pick_k_kernel(unsigned long long* __restrict__ state, float ratio) {
    const double want = rint((double)state[0] * (double)ratio);   // python round() == round-half-even == rint
    unsigned long long k = (unsigned long long)(want < 0 ? 0 : want);
    if (k > state[1]) k = state[1];
    state[2] = k;
    state[3] = 0ull;
    state[4] = k;      // rank (1-based from the top) still to locate
    state[5] = 0ull;
}
// one radix pass over byte `shift/8` of the non-negative float bit patterns, restricted to the current prefix
__global__ void __launch_bounds__(256)
bce_hist_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask, long long n,
                int shift, const unsigned long long* __restrict__ state, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned int prefix = (unsigned int)state[3];
    const unsigned int himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f;
        const float nm = (float)(unsigned char)(int)((1.f - gt[i]) * m);
        if (nm == 0.f) continue;
        const unsigned int bits = __float_as_uint(fmaxf(bce_logits(pred[i], gt[i]) * nm, 0.f));
        if ((bits & himask) == (prefix & himask)) atomicAdd(&sh[(bits >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}
__global__ void bce_scan_kernel(int shift, unsigned long long* __restrict__ state, unsigned long long* __restrict__ hist) {
    unsigned long long rank = state[4];
    unsigned int digit = 0;
    if (rank > 0) {
        for (int d = 255; d >= 0; --d) {
            if (hist[d] >= rank) { digit = (unsigned int)d; break; }
            rank -= hist[d];
        }
    }
    state[3] |= (unsigned long long)digit << shift;
    state[4] = rank;
    if (shift == 0) state[5] = state[3];
    for (int d = 0; d < 256; ++d) hist[d] = 0ull;
}
__global__ void __launch_bounds__(256)
bce_topsum_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask, long long n,
                  const unsigned long long* __restrict__ state, double* __restrict__ sums) {
    __shared__ float scratch[33];
    const unsigned int thr = (unsigned int)state[5];
    float v[2] = {0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f;
        const float nm = (float)(unsigned char)(int)((1.f - gt[i]) * m);
        if (nm == 0.f) continue;
        const float l = fmaxf(bce_logits(pred[i], gt[i]) * nm, 0.f);
        if (__float_as_uint(l) > thr) { v[0] += l; v[1] += 1.f; }
    }
    block_accumulate(v, 2, sums + 1, scratch);
}
// coef: [0] loss, [1] 1/(pos+k+eps), [2] threshold value, [3] weight of elements equal to the threshold
__global__ void bce_finalize_kernel(const unsigned long long* __restrict__ state, const double* __restrict__ sums, float eps,
                                    float* __restrict__ coef) {
    const double k = (double)state[2];
    const float thr = __uint_as_float((unsigned int)state[5]);
    const double ties = k - sums[2];                       // how many threshold-valued elements belong to the top-k
    const double neg = k > 0 ? sums[1] + ties * (double)thr : 0.0;
    const double den = (double)state[0] + k + (double)eps;
    coef[0] = (float)((sums[0] + neg) / den);
    coef[1] = (float)(1.0 / den);
    coef[2] = thr;
    coef[3] = (float)ties;
}
// gradient: positives and the selected negatives get (sigmoid(x) - gt) * weight / den.  Among elements tied at the
// threshold torch.topk picks an arbitrary subset of size `ties`; the first `ties` in index order are used here.
__global__ void __launch_bounds__(256)
bce_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask, long long n,
               const unsigned long long* __restrict__ state, const float* __restrict__ coef, const float* __restrict__ gout,
               unsigned long long* __restrict__ tie_counter, float* __restrict__ dpred) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = mask ? mask[i] : 1.f;
    const float pm = (float)(unsigned char)(int)(gt[i] * m), nm = (float)(unsigned char)(int)((1.f - gt[i]) * m);
    const float x = pred[i];
    const float dl = sigm(x) - gt[i];
    float w = pm;
    if (nm != 0.f && state[2] > 0ull) {
        const unsigned int bits = __float_as_uint(fmaxf(bce_logits(x, gt[i]) * nm, 0.f));
        const unsigned int thr = (unsigned int)state[5];
        if (bits > thr) w += nm;
        else if (bits == thr) {
            const unsigned long long t = atomicAdd(tie_counter, 1ull);
            if ((double)t < (double)coef[3]) w += nm;
        }
    }
    dpred[i] = gout[0] * coef[1] * w * dl;
}

unsigned reduce_blocks(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

// kind: 0 focal (p0 alpha, p1 gamma), 1 dice, 2 L1, 3 smooth-L1 (p0 beta), 4 L2, 5 WAHR (p0 gamma).
// pre_sigmoid != 0: the primitive is applied to sigmoid(pred) and the gradient is chained through the sigmoid.
// sums: 3 doubles zeroed by the caller; coef: 3 floats (coef[0] = loss).
int vkocr_pointwise_loss_fwd(int kind, int pre_sigmoid, const float* pred, const float* gt, const float* mask, long long n, float p0, float p1,
                             double* sums, float* coef, void* stream) {
    VK_REQUIRE(pred && gt && sums && coef, VKOCR_BAD_ARGUMENT, "pointwise_loss_fwd: null argument");
    VK_REQUIRE(kind >= 0 && kind <= 5, VKOCR_BAD_ARGUMENT, "pointwise_loss_fwd: kind %d", kind);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (n > 0) {
        pw_reduce_kernel<<<reduce_blocks(n), 256, 0, s>>>(kind, pre_sigmoid, pred, gt, mask, n, p0, p1, sums);
        VK_CHECK_LAUNCH("pw_reduce_kernel");
    }
    pw_finalize_kernel<<<1, 1, 0, s>>>(kind, sums, (double)n, mask != nullptr, coef);
    VK_CHECK_LAUNCH("pw_finalize_kernel");
    return VKOCR_OK;
}

int vkocr_pointwise_loss_bwd(int kind, int pre_sigmoid, const float* pred, const float* gt, const float* mask, long long n, float p0, float p1,
                             const float* coef, const float* grad_out, float* dpred, void* stream) {
    VK_REQUIRE(pred && gt && coef && grad_out && dpred, VKOCR_BAD_ARGUMENT, "pointwise_loss_bwd: null argument");
    if (n == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    pw_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(kind, pre_sigmoid, pred, gt, mask, n, p0, p1, coef, grad_out, dpred);
    VK_CHECK_LAUNCH("pw_bwd_kernel");
    return VKOCR_OK;
}

// pred/gt: (outer, C, inner) fp32 contiguous; sums: 1 double zeroed; coef: 2 floats.
int vkocr_soft_ce_fwd(const float* pred, const float* gt, long long outer, int C, long long inner, double* sums, float* coef,
                      void* stream) {
    VK_REQUIRE(pred && gt && sums && coef, VKOCR_BAD_ARGUMENT, "soft_ce_fwd: null argument");
    VK_REQUIRE(C >= 1, VKOCR_BAD_SHAPE, "soft_ce_fwd: C %d", C);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = outer * inner;
    if (total > 0) {
        ce_reduce_kernel<<<reduce_blocks(total), 256, 0, s>>>(pred, gt, outer, C, inner, sums);
        VK_CHECK_LAUNCH("ce_reduce_kernel");
    }
    ce_finalize_kernel<<<1, 1, 0, s>>>(sums, (double)total, coef);
    VK_CHECK_LAUNCH("ce_finalize_kernel");
    return VKOCR_OK;
}

int vkocr_soft_ce_bwd(const float* pred, const float* gt, long long outer, int C, long long inner, const float* coef,
                      const float* grad_out, float* dpred, void* stream) {
    VK_REQUIRE(pred && gt && coef && grad_out && dpred, VKOCR_BAD_ARGUMENT, "soft_ce_bwd: null argument");
    const long long total = outer * inner;
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    ce_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pred, gt, outer, C, inner, coef, grad_out, dpred);
    VK_CHECK_LAUNCH("ce_bwd_kernel");
    return VKOCR_OK;
}

// state: 8 x u64 + hist: 256 x u64 (one zeroed 264-element u64 buffer: state first); sums: 4 doubles zeroed; coef: 4 floats.
int vkocr_hard_negative_bce_fwd(const float* pred, const float* gt, const float* mask, long long n, float negative_ratio,
                                float eps, unsigned long long* state_hist, double* sums, float* coef, void* stream) {
    VK_REQUIRE(pred && gt && state_hist && sums && coef, VKOCR_BAD_ARGUMENT, "hard_negative_bce_fwd: null argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long* state = state_hist;
    unsigned long long* hist = state_hist + 8;
    const unsigned blocks = reduce_blocks(n);
    if (n > 0) {
        bce_count_kernel<<<blocks, 256, 0, s>>>(pred, gt, mask, n, state, sums);
        VK_CHECK_LAUNCH("bce_count_kernel");
    }
    bce_pick_k_kernel<<<1, 1, 0, s>>>(state, negative_ratio);
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (n > 0) bce_hist_kernel<<<blocks, 256, 0, s>>>(pred, gt, mask, n, shift, state, hist);
        bce_scan_kernel<<<1, 1, 0, s>>>(shift, state, hist);
    }
    if (n > 0) bce_topsum_kernel<<<blocks, 256, 0, s>>>(pred, gt, mask, n, state, sums);
    bce_finalize_kernel<<<1, 1, 0, s>>>(state, sums, eps, coef);
    VK_CHECK_LAUNCH("hard_negative_bce");
    return VKOCR_OK;
}

// tie_counter: one zeroed u64.
int vkocr_hard_negative_bce_bwd(const float* pred, const float* gt, const float* mask, long long n,
                                const unsigned long long* state_hist, const float* coef, const float* grad_out,
                                unsigned long long* tie_counter, float* dpred, void* stream) {
    VK_REQUIRE(pred && gt && state_hist && coef && grad_out && tie_counter && dpred, VKOCR_BAD_ARGUMENT,
               "hard_negative_bce_bwd: null argument");
    if (n == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bce_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pred, gt, mask, n, state_hist, coef, grad_out, tie_counter, dpred);
    VK_CHECK_LAUNCH("bce_bwd_kernel");
    return VKOCR_OK;
}

}  // extern "C"
