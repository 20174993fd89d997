// Fused adaptive-scaling losses (forward reductions, finalize, backward) on the NCHW fp32 prediction maps.
// Reference: AdaptiveScalingRoughLossFunction / AdaptiveScalingPreciseLossFunction
// (loss_function/adaptive_scaling.py:38-131, 148-346) over the primitives focal (torchvision sigmoid_focal_loss,
// alpha .25 gamma 2), dice, masked smooth-L1, masked L2, soft-label cross entropy.
//
// HBM-bound reductions: every map is read once per pass, block partials via warp shuffles, one fp64 atomic per block
// and quantity.  A 1-thread finalize kernel turns the sums into the scalar loss and the coefficients the backward
// kernels need, so no value ever travels to the host.
#include "common.cuh"

namespace {

constexpr float EPS = 1e-6f;

__device__ __forceinline__ float smooth_l1(float d, float beta) {
    const float a = fabsf(d);
    return a < beta ? 0.5f * d * d / beta : a - 0.5f * beta;
}
__device__ __forceinline__ float smooth_l1_grad(float d, float beta) {
    const float a = fabsf(d);
    return a < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
}
__device__ __forceinline__ float bce_logits(float x, float t) {
    return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void block_accumulate(const float* vals, int n, double* out, float* scratch) {
    for (int i = 0; i < n; ++i) {
        const float s = vk_block_sum(vals[i], scratch);
        if (threadIdx.x == 0 && s != 0.f) atomicAdd(out + i, (double)s);
    }
}

// ---------------------------------------------------------------------------------------------------- rough loss
// sums: [0] focal, [1] sum p*t, [2] sum p, [3] sum t, [4] masked smooth-l1 sum, [5] mask count
__global__ void __launch_bounds__(256)
rough_reduce_kernel(const float* __restrict__ logit, const float* __restrict__ height, const float* __restrict__ gt_mask,
                    const float* __restrict__ gt_score, int B, int H, int W, int up, int left, int CH, int CW, float hmin,
                    float smin, double* __restrict__ sums) {
    __shared__ float scratch[33];
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long total = (long long)B * CH * CW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cx = (int)(i % CW);
        const int cy = (int)((i / CW) % CH);
        const int b = (int)(i / ((long long)CW * CH));
        const long long pi = ((long long)b * H + up + cy) * W + left + cx;
        const float x = logit[pi], h = height[pi], t = gt_mask[i], s = gt_score[i];
        const float p = sigmoidf_precise(x);
        const float ce = bce_logits(x, t);
        const float pt = p * t + (1.f - p) * (1.f - t);
        const float at = 0.25f * t + 0.75f * (1.f - t);
        v[0] += at * ce * (1.f - pt) * (1.f - pt);
        v[1] += p * t;
        v[2] += p;
        v[3] += t;
        const bool m = (h > hmin) && (s > smin) && (t != 0.f);
        if (m) {
            v[4] += smooth_l1(logf(fmaxf(h, hmin)) - logf(fmaxf(s, smin)), 1.f);
            v[5] += 1.f;
        }
    }
    block_accumulate(v, 6, sums, scratch);
}

// coef: [0] loss, [1] focal_factor/n, [2] I, [3] U, [4] dice_factor, [5] l1_factor/(count+eps)
__global__ void rough_finalize_kernel(const double* __restrict__ sums, double n, float ff, float df, float lf, float* __restrict__ coef) {
    const double focal = sums[0] / n;
    const double I = sums[1], U = sums[2] + sums[3] + (double)EPS;
    const double dice = 1.0 - 2.0 * I / U;
    const double l1 = sums[4] / (sums[5] + (double)EPS);
    double loss = 0.0;
    if (ff > 0.f) loss += ff * focal;
    if (df > 0.f) loss += df * dice;
    if (lf > 0.f) loss += lf * l1;
    coef[0] = (float)loss;
    coef[1] = ff > 0.f ? (float)(ff / n) : 0.f;
    coef[2] = (float)I;
    coef[3] = (float)U;
    coef[4] = df > 0.f ? df : 0.f;
    coef[5] = lf > 0.f ? (float)(lf / (sums[5] + (double)EPS)) : 0.f;
}

// writes the full (B,H,W) gradient maps (zero outside the core box)
__global__ void __launch_bounds__(256)
rough_bwd_kernel(const float* __restrict__ logit, const float* __restrict__ height, const float* __restrict__ gt_mask,
                 const float* __restrict__ gt_score, int B, int H, int W, int up, int left, int CH, int CW, float hmin, float smin,
                 const float* __restrict__ coef, const float* __restrict__ gout, float* __restrict__ dlogit,
                 float* __restrict__ dheight) {
    const long long total = (long long)B * H * W;
    const long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= total) return;
    const int xx = (int)(pi % W);
    const int yy = (int)((pi / W) % H);
    const int b = (int)(pi / ((long long)W * H));
    const int cy = yy - up, cx = xx - left;
    float dl = 0.f, dh = 0.f;
    if (cy >= 0 && cy < CH && cx >= 0 && cx < CW) {
        const long long i = ((long long)b * CH + cy) * CW + cx;
        const float g = gout[0];
        const float x = logit[pi], h = height[pi], t = gt_mask[i], s = gt_score[i];
        const float p = sigmoidf_precise(x);
        const float ce = bce_logits(x, t);
        const float pt = p * t + (1.f - p) * (1.f - t);
        const float at = 0.25f * t + 0.75f * (1.f - t);
        const float omp = 1.f - pt;
        const float dfocal = at * ((p - t) * omp * omp - 2.f * ce * omp * p * (1.f - p) * (2.f * t - 1.f));
        const float I = coef[2], U = coef[3];
        const float ddice = -2.f * (t * U - I) / (U * U) * p * (1.f - p);
        dl = g * (coef[1] * dfocal + coef[4] * ddice);
        const bool m = (h > hmin) && (s > smin) && (t != 0.f);
        if (m) dh = g * coef[5] * smooth_l1_grad(logf(fmaxf(h, hmin)) - logf(fmaxf(s, smin)), 1.f) / h;
    }
    dlogit[pi] = dl;
    dheight[pi] = dh;
}

// ---------------------------------------------------------------------------------------------------- precise loss
// dense sums: [0] sum m*(p-s)^2, [1] sum m, [2] sum (1-m)*(p-s)^2, [3] sum (1-m)
__global__ void __launch_bounds__(256)
precise_dense_reduce_kernel(const float* __restrict__ logit, const float* __restrict__ gt_score, const float* __restrict__ gt_mask,
                            int B, int H, int W, int up, int left, int CH, int CW, double* __restrict__ sums) {
    __shared__ float scratch[33];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const long long total = (long long)B * CH * CW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cx = (int)(i % CW);
        const int cy = (int)((i / CW) % CH);
        const int b = (int)(i / ((long long)CW * CH));
        const long long pi = ((long long)b * H + up + cy) * W + left + cx;
        const float p = sigmoidf_precise(logit[pi]);
        const float d = p - gt_score[i], m = gt_mask[i];
        v[0] += m * d * d;
        v[1] += m;
        v[2] += (1.f - m) * d * d;
        v[3] += 1.f - m;
    }
    block_accumulate(v, 4, sums, scratch);
}

// point sums: [4] offset smooth-l1, [5] distance-regulation smooth-l1, [6] soft CE, [7] corner-distance smooth-l1
__global__ void __launch_bounds__(256)
precise_points_reduce_kernel(const float* __restrict__ off, const float* __restrict__ ang, const float* __restrict__ dist,
                             const long long* __restrict__ py, const long long* __restrict__ px,
                             const float* __restrict__ gt_off, const float* __restrict__ gt_ang,
                             const float* __restrict__ gt_dist, int B, int P, int H, int W, float beta, double* __restrict__ sums) {
    __shared__ float scratch[33];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const long long total = (long long)B * P;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / P);
        long long y = py[i], x = px[i];
        if (y < 0) y += H;   // torch advanced indexing wraps negative indices
        if (x < 0) x += W;
        if (y < 0 || y >= H || x < 0 || x >= W) {
            // the reference's advanced indexing raises IndexError here; without a host sync the loud failure is a NaN loss
            // (see precise_finalize_kernel), and nothing is read or (in the backward) written out of bounds
            v[0] = v[1] = v[2] = v[3] = __int_as_float(0x7fc00000);
            continue;
        }
        const long long pix = y * W + x;
        const float o0 = off[((long long)b * 2 + 0) * hw + pix], o1 = off[((long long)b * 2 + 1) * hw + pix];
        v[0] += smooth_l1(o0 - gt_off[i * 2 + 0], beta) + smooth_l1(o1 - gt_off[i * 2 + 1], beta);
        float a[4], d[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[c] = ang[((long long)b * 4 + c) * hw + pix];
            d[c] = dist[((long long)b * 4 + c) * hw + pix];
        }
        v[1] += smooth_l1(sqrtf(o0 * o0 + o1 * o1) - d[0], beta);
        const float mx = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
        const float lse = mx + logf(expf(a[0] - mx) + expf(a[1] - mx) + expf(a[2] - mx) + expf(a[3] - mx));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[2] -= gt_ang[i * 4 + c] * (a[c] - lse);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[3] += smooth_l1(d[c + 1] - gt_dist[i * 3 + c], beta);
    }
    block_accumulate(v, 4, sums + 4, scratch);
}

// f: [0] pos_l2, [1] neg_l2, [2] offset_l1, [3] distance regulation, [4] angle CE, [5] corner distance, [6] loss_factor
// coef: [0] loss, [1] lf*pf/(pos_den+eps), [2] lf*nf/(neg_den+eps)
__global__ void precise_finalize_kernel(const double* __restrict__ sums, double bp, const float* __restrict__ f,
                                        float* __restrict__ coef) {
    double loss = 0.0;
    if (f[0] > 0.f) loss += f[0] * sums[0] / (sums[1] + (double)EPS);
    if (f[1] > 0.f) loss += f[1] * sums[2] / (sums[3] + (double)EPS);
    if (f[2] > 0.f) loss += f[2] * sums[4] / (bp * 2.0);
    if (f[3] > 0.f) loss += f[3] * sums[5] / bp;
    if (f[4] > 0.f) loss += f[4] * sums[6] / bp;
    if (f[5] > 0.f) loss += f[5] * sums[7] / (bp * 3.0);
    loss *= f[6];
    if (sums[4] != sums[4] || sums[5] != sums[5] || sums[6] != sums[6] || sums[7] != sums[7]) loss = sums[4] + sums[5] + sums[6] + sums[7];   // a label point outside the map
    coef[0] = (float)loss;
    coef[1] = f[0] > 0.f ? (float)(f[6] * f[0] / (sums[1] + (double)EPS)) : 0.f;
    coef[2] = f[1] > 0.f ? (float)(f[6] * f[1] / (sums[3] + (double)EPS)) : 0.f;
}

__global__ void __launch_bounds__(256)
precise_dense_bwd_kernel(const float* __restrict__ logit, const float* __restrict__ gt_score, const float* __restrict__ gt_mask,
                         int B, int H, int W, int up, int left, int CH, int CW, const float* __restrict__ coef,
                         const float* __restrict__ gout, float* __restrict__ dlogit) {
    const long long total = (long long)B * H * W;
    const long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= total) return;
    const int xx = (int)(pi % W);
    const int yy = (int)((pi / W) % H);
    const int b = (int)(pi / ((long long)W * H));
    const int cy = yy - up, cx = xx - left;
    float dl = 0.f;
    if (cy >= 0 && cy < CH && cx >= 0 && cx < CW) {
        const long long i = ((long long)b * CH + cy) * CW + cx;
        const float p = sigmoidf_precise(logit[pi]);
        const float m = gt_mask[i];
        dl = gout[0] * 2.f * (p - gt_score[i]) * p * (1.f - p) * (coef[1] * m + coef[2] * (1.f - m));
    }
    dlogit[pi] = dl;
}

// scatter-adds into zero-initialised (B,2|4|4,H,W) gradient maps (duplicate points accumulate)
__global__ void __launch_bounds__(256)
precise_points_bwd_kernel(const float* __restrict__ off, const float* __restrict__ ang, const float* __restrict__ dist,
                          const long long* __restrict__ py, const long long* __restrict__ px,
                          const float* __restrict__ gt_off, const float* __restrict__ gt_ang,
                          const float* __restrict__ gt_dist, int B, int P, int H, int W, float beta, const float* __restrict__ f,
                          const float* __restrict__ gout, float* __restrict__ doff, float* __restrict__ dang,
                          float* __restrict__ ddist) {
    const long long total = (long long)B * P;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long hw = (long long)H * W;
    const int b = (int)(i / P);
    long long y = py[i], x = px[i];
    if (y < 0) y += H;
    if (x < 0) x += W;
    if (y < 0 || y >= H || x < 0 || x >= W) return;   // see the forward: the loss is NaN, nothing is scattered
    const long long pix = y * W + x;
    const float g = gout[0] * f[6];
    const float bp = (float)total;
    const float o0 = off[((long long)b * 2 + 0) * hw + pix], o1 = off[((long long)b * 2 + 1) * hw + pix];
    float a[4], d[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a[c] = ang[((long long)b * 4 + c) * hw + pix];
        d[c] = dist[((long long)b * 4 + c) * hw + pix];
    }
    float g0 = 0.f, g1 = 0.f, gd0 = 0.f;
    if (f[2] > 0.f) {
        const float k = g * f[2] / (bp * 2.f);
        g0 += k * smooth_l1_grad(o0 - gt_off[i * 2 + 0], beta);
        g1 += k * smooth_l1_grad(o1 - gt_off[i * 2 + 1], beta);
    }
    if (f[3] > 0.f) {
        const float nrm = sqrtf(o0 * o0 + o1 * o1);
        const float k = g * f[3] / bp * smooth_l1_grad(nrm - d[0], beta);
        if (nrm > 0.f) {
            g0 += k * o0 / nrm;
            g1 += k * o1 / nrm;
        }
        gd0 -= k;
    }
    atomicAdd(doff + ((long long)b * 2 + 0) * hw + pix, g0);
    atomicAdd(doff + ((long long)b * 2 + 1) * hw + pix, g1);
    atomicAdd(ddist + ((long long)b * 4 + 0) * hw + pix, gd0);
    if (f[4] > 0.f) {
        const float mx = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
        float e[4], se = 0.f, st = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            e[c] = expf(a[c] - mx);
            se += e[c];
            st += gt_ang[i * 4 + c];
        }
        const float k = g * f[4] / bp;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            atomicAdd(dang + ((long long)b * 4 + c) * hw + pix, k * (e[c] / se * st - gt_ang[i * 4 + c]));
    }
    if (f[5] > 0.f) {
        const float k = g * f[5] / (bp * 3.f);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            atomicAdd(ddist + ((long long)b * 4 + c + 1) * hw + pix, k * smooth_l1_grad(d[c + 1] - gt_dist[i * 3 + c], beta));
    }
}

unsigned reduce_blocks(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

// sums: 6 doubles (zeroed by the caller); coef: 6 floats, coef[0] receives the loss.
int vkocr_rough_loss_fwd(const float* logit, const float* height, const float* gt_mask, const float* gt_score, int B, int H, int W,
                         int up, int left, int CH, int CW, float height_min, float score_min, float focal_factor,
                         float dice_factor, float l1_factor, double* sums, float* coef, void* stream) {
    VK_REQUIRE(logit && height && gt_mask && gt_score && sums && coef, VKOCR_BAD_ARGUMENT, "rough_loss_fwd: null argument");
    VK_REQUIRE(up >= 0 && left >= 0 && up + CH <= H && left + CW <= W, VKOCR_BAD_SHAPE, "rough_loss_fwd: core box outside the map");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)B * CH * CW;
    if (total > 0) {
        rough_reduce_kernel<<<reduce_blocks(total), 256, 0, s>>>(logit, height, gt_mask, gt_score, B, H, W, up, left, CH, CW,
                                                                  height_min, score_min, sums);
        VK_CHECK_LAUNCH("rough_reduce_kernel");
    }
    rough_finalize_kernel<<<1, 1, 0, s>>>(sums, (double)total, focal_factor, dice_factor, l1_factor, coef);
    VK_CHECK_LAUNCH("rough_finalize_kernel");
    return VKOCR_OK;
}

int vkocr_rough_loss_bwd(const float* logit, const float* height, const float* gt_mask, const float* gt_score, int B, int H, int W,
                         int up, int left, int CH, int CW, float height_min, float score_min, const float* coef,
                         const float* grad_out, float* dlogit, float* dheight, void* stream) {
    VK_REQUIRE(logit && height && gt_mask && gt_score && coef && grad_out && dlogit && dheight, VKOCR_BAD_ARGUMENT,
               "rough_loss_bwd: null argument");
    const long long total = (long long)B * H * W;
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    rough_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(logit, height, gt_mask, gt_score, B, H, W, up, left, CH, CW,
                                                                      height_min, score_min, coef, grad_out, dlogit, dheight);
    VK_CHECK_LAUNCH("rough_bwd_kernel");
    return VKOCR_OK;
}

// factors: 7 device floats (see precise_finalize_kernel); sums: 8 doubles zeroed by the caller; coef: 3 floats.
int vkocr_precise_loss_fwd(const float* prob, const float* off, const float* ang, const float* dist, const float* gt_score,
                           const float* gt_mask, int B, int H, int W, int up, int left, int CH, int CW, const long long* py,
                           const long long* px, const float* gt_off, const float* gt_ang, const float* gt_dist, int P,
                           float beta, const float* factors, double* sums, float* coef, void* stream) {
    VK_REQUIRE(prob && off && ang && dist && gt_score && gt_mask && factors && sums && coef, VKOCR_BAD_ARGUMENT,
               "precise_loss_fwd: null argument");
    VK_REQUIRE(P == 0 || (py && px && gt_off && gt_ang && gt_dist), VKOCR_BAD_ARGUMENT, "precise_loss_fwd: null label-point argument");
    VK_REQUIRE(up >= 0 && left >= 0 && up + CH <= H && left + CW <= W, VKOCR_BAD_SHAPE, "precise_loss_fwd: core box outside the map");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)B * CH * CW;
    if (total > 0) {
        precise_dense_reduce_kernel<<<reduce_blocks(total), 256, 0, s>>>(prob, gt_score, gt_mask, B, H, W, up, left, CH, CW, sums);
        VK_CHECK_LAUNCH("precise_dense_reduce_kernel");
    }
    const long long pts = (long long)B * P;
    if (pts > 0) {
        precise_points_reduce_kernel<<<reduce_blocks(pts), 256, 0, s>>>(off, ang, dist, py, px, gt_off, gt_ang, gt_dist, B, P, H, W,
                                                                          beta, sums);
        VK_CHECK_LAUNCH("precise_points_reduce_kernel");
    }
    precise_finalize_kernel<<<1, 1, 0, s>>>(sums, (double)pts, factors, coef);
    VK_CHECK_LAUNCH("precise_finalize_kernel");
    return VKOCR_OK;
}

// dprob is fully written; doff/dang/ddist must be zero-initialised by the caller (scatter-add).
int vkocr_precise_loss_bwd(const float* prob, const float* off, const float* ang, const float* dist, const float* gt_score,
                           const float* gt_mask, int B, int H, int W, int up, int left, int CH, int CW, const long long* py,
                           const long long* px, const float* gt_off, const float* gt_ang, const float* gt_dist, int P,
                           float beta, const float* factors, const float* coef, const float* grad_out, float* dprob, float* doff,
                           float* dang, float* ddist, void* stream) {
    VK_REQUIRE(prob && off && ang && dist && gt_score && gt_mask && factors && coef && grad_out && dprob && doff && dang && ddist,
               VKOCR_BAD_ARGUMENT, "precise_loss_bwd: null argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)B * H * W;
    if (total > 0) {
        precise_dense_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(prob, gt_score, gt_mask, B, H, W, up, left, CH, CW,
                                                                                  coef, grad_out, dprob);
        VK_CHECK_LAUNCH("precise_dense_bwd_kernel");
    }
    const long long pts = (long long)B * P;
    if (pts > 0) {
        precise_points_bwd_kernel<<<(unsigned)((pts + 255) / 256), 256, 0, s>>>(off, ang, dist, py, px, gt_off, gt_ang, gt_dist, B, P,
                                                                                 H, W, beta, factors, grad_out, doff, dang, ddist);
        VK_CHECK_LAUNCH("precise_points_bwd_kernel");
    }
    return VKOCR_OK;
}

}  // extern "C"
