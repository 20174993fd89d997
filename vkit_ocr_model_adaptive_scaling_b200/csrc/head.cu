// Tail of the adaptive-scaling heads, fused: LayerNorm(inner) -> exact GELU -> Linear(inner -> out<=4) (-> Softplus),
// reading the 3x3-conv output slice once and writing the NCHW fp32 prediction map.
// Reference: UperNextHead.forward upernext.py:233-248 / FpnHead.forward fpn.py:193-208 (step1 LN+GELU, step2 1x1),
// nn.Softplus on the height / distance heads (adaptive_scaling.py:101,140).
//
// HBM-bound: one warp per pixel, each lane owns NVL 16-byte channel vectors of the slice (the slice is padded to a
// multiple of the vector width; pad channels hold zeros and are masked out of the statistics).
// Backward recomputes LN/GELU from the saved conv output, writes d(conv output) and accumulates all parameter
// gradients (LN gamma/beta, 1x1 weight/bias, conv bias) from per-lane register partials.
#include "common.cuh"

namespace {

constexpr int HT_THREADS = 256;
constexpr int HT_WARPS = HT_THREADS / 32;
constexpr float LN_EPS = 1e-6f;

template <typename T, int NVL, int O>
__global__ void __launch_bounds__(HT_THREADS)
head_tail_fwd_kernel(const T* __restrict__ x, long long ld_x, int inner, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ w2, const float* __restrict__ b2, int softplus,
                     float* __restrict__ out, long long pixels_per_image, long long rows) {
    constexpr int V = VkVec<T>::N;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * HT_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * HT_WARPS;
    float gm[NVL][V], bt[NVL][V], w[O][NVL][V];
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = (lane + 32 * j) * V + i;
            const bool ok = c < inner;
            gm[j][i] = ok ? __ldg(gamma + c) : 0.f;
            bt[j][i] = ok ? __ldg(beta + c) : 0.f;
#pragma unroll
            for (int o = 0; o < O; ++o) w[o][j][i] = ok ? __ldg(w2 + (long long)o * inner + c) : 0.f;
        }
    const float inv = 1.f / inner;
    for (long long r = warp0; r < rows; r += nwarps) {
        const T* xr = x + r * ld_x;
        float f[NVL][V];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NVL; ++j) {
            const int c = (lane + 32 * j) * V;
            if (c < inner) {
                VkVec<T> v;
                v.load(xr + c);
                v.unpack(f[j]);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) f[j][i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < V; ++i) s += (c + i < inner) ? f[j][i] : 0.f;
        }
        const float mean = vk_warp_sum(s) * inv;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = (lane + 32 * j) * V + i;
                const float d = f[j][i] - mean;
                q += (c < inner) ? d * d : 0.f;
            }
        const float rstd = rsqrtf(vk_warp_sum(q) * inv + LN_EPS);
        float dot[O];
#pragma unroll
        for (int o = 0; o < O; ++o) dot[o] = 0.f;
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float g = vk_gelu((f[j][i] - mean) * rstd * gm[j][i] + bt[j][i]);
#pragma unroll
                for (int o = 0; o < O; ++o) dot[o] = fmaf(g, w[o][j][i], dot[o]);   // w == 0 on pad channels
            }
#pragma unroll
        for (int o = 0; o < O; ++o) dot[o] = vk_warp_sum(dot[o]);
        if (lane < O) {
            float v = 0.f;
#pragma unroll
            for (int o = 0; o < O; ++o) v = (lane == o) ? dot[o] : v;
            v += __ldg(b2 + lane);
            if (softplus) v = vk_softplus(v);
            const long long b = r / pixels_per_image, pix = r % pixels_per_image;
            out[(b * O + lane) * pixels_per_image + pix] = v;
        }
    }
}

// Backward.  Per-channel gradient partials (dgamma, dbeta, conv-bias, dW2) live in registers of the lane that owns the
// channel for all rows the warp visits; the LayerNorm / projection parameters are read from shared memory each row so
// that the register budget allows two blocks per SM (the kernel is instruction/latency bound, not HBM bound).
template <typename T, int NVL, int O>
__global__ void __launch_bounds__(HT_THREADS, (NVL * O <= 4) ? 2 : 1)
head_tail_bwd_kernel(const T* __restrict__ x, long long ld_x, int inner, int slice_w, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ w2, int softplus, const float* __restrict__ out,
                     const float* __restrict__ dout, long long pixels_per_image, long long rows, T* __restrict__ dx,
                     long long ld_dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw2,
                     float* __restrict__ db2, float* __restrict__ dbias) {
    constexpr int V = VkVec<T>::N;
    constexpr int CW = 32 * NVL * V;
    extern __shared__ float sm[];
    float* s_gm = sm;                 // [CW]
    float* s_bt = sm + CW;            // [CW]
    float* s_w = sm + 2 * CW;         // [O][CW]
    float* sacc = sm + (2 + O) * CW;  // [(3 + O)][CW] + [O]
    for (int i = threadIdx.x; i < CW; i += blockDim.x) {
        const bool ok = i < inner;
        s_gm[i] = ok ? gamma[i] : 0.f;
        s_bt[i] = ok ? beta[i] : 0.f;
#pragma unroll
        for (int o = 0; o < O; ++o) s_w[o * CW + i] = ok ? w2[(long long)o * inner + i] : 0.f;
    }
    for (int i = threadIdx.x; i < (3 + O) * CW + O; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * HT_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * HT_WARPS;
    float ag[NVL][V], ab[NVL][V], ax[NVL][V], aw[O][NVL][V];
    float adb[O];
#pragma unroll
    for (int o = 0; o < O; ++o) adb[o] = 0.f;
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            ag[j][i] = ab[j][i] = ax[j][i] = 0.f;
#pragma unroll
            for (int o = 0; o < O; ++o) aw[o][j][i] = 0.f;
        }
    const float inv = 1.f / inner;
    for (long long r = warp0; r < rows; r += nwarps) {
        const T* xr = x + r * ld_x;
        float f[NVL][V];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NVL; ++j) {
            const int c = (lane + 32 * j) * V;
            if (c < inner) {
                VkVec<T> v;
                v.load(xr + c);
                v.unpack(f[j]);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) f[j][i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < V; ++i) s += (c + i < inner) ? f[j][i] : 0.f;
        }
        // upstream gradient of the pre-softplus outputs (same value in every lane)
        const long long b = r / pixels_per_image, pix = r % pixels_per_image;
        float dpre[O];
#pragma unroll
        for (int o = 0; o < O; ++o) {
            const long long oi = (b * O + o) * pixels_per_image + pix;
            float d = __ldg(dout + oi);
            if (softplus) {
                const float y = __ldg(out + oi);
                d *= (y > 20.f) ? 1.f : (1.f - __expf(-y));   // sigmoid(pre) = 1 - exp(-softplus(pre))
            }
            dpre[o] = d;
            adb[o] += d;
        }
        const float mean = vk_warp_sum(s) * inv;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = (lane + 32 * j) * V + i;
                const float d = f[j][i] - mean;
                q += (c < inner) ? d * d : 0.f;
            }
        const float rstd = rsqrtf(vk_warp_sum(q) * inv + LN_EPS);
        float s1 = 0.f, s2 = 0.f;
        // after this loop f holds xhat and dz holds d(loss)/d(LN output) per channel
        float dz[NVL][V];
#pragma unroll
        for (int j = 0; j < NVL; ++j) {
            const int c0 = (lane + 32 * j) * V;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = c0 + i;
                const float gmv = s_gm[c];
                const float h = (f[j][i] - mean) * rstd;
                const float z = fmaf(h, gmv, s_bt[c]);
                float g, gp;
                vk_gelu_both(z, &g, &gp);
                float dg = 0.f;
#pragma unroll
                for (int o = 0; o < O; ++o) {
                    dg = fmaf(dpre[o], s_w[o * CW + c], dg);     // w == 0 on pad channels
                    aw[o][j][i] = fmaf(dpre[o], g, aw[o][j][i]);
                }
                const bool ok = c < inner;
                const float d = ok ? dg * gp : 0.f;
                const float xh = ok ? h : 0.f;
                f[j][i] = xh;
                dz[j][i] = d;
                const float dxh = d * gmv;
                s1 += dxh;
                s2 = fmaf(dxh, xh, s2);
                ag[j][i] = fmaf(d, xh, ag[j][i]);
                ab[j][i] += d;
            }
        }
        s1 = vk_warp_sum(s1) * inv;
        s2 = vk_warp_sum(s2) * inv;
        T* dxr = dx + r * ld_dx;
#pragma unroll
        for (int j = 0; j < NVL; ++j) {
            const int c0 = (lane + 32 * j) * V;
            float fo[V];
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = c0 + i;
                const float dxv = (c < inner) ? rstd * (dz[j][i] * s_gm[c] - s1 - f[j][i] * s2) : 0.f;
                fo[i] = dxv;
                ax[j][i] += dxv;
            }
            if (c0 < slice_w) {
                VkVec<T> vo;
                vo.pack(fo);
                vo.store(dxr + c0);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = (lane + 32 * j) * V + i;
            if (c < inner) {
                atomicAdd(&sacc[c], ag[j][i]);
                atomicAdd(&sacc[CW + c], ab[j][i]);
                atomicAdd(&sacc[2 * CW + c], ax[j][i]);
#pragma unroll
                for (int o = 0; o < O; ++o) atomicAdd(&sacc[(3 + o) * CW + c], aw[o][j][i]);
            }
        }
    if (lane == 0) {
#pragma unroll
        for (int o = 0; o < O; ++o) atomicAdd(&sacc[(3 + O) * CW + o], adb[o]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < inner; c += blockDim.x) {
        atomicAdd(dgamma + c, sacc[c]);
        atomicAdd(dbeta + c, sacc[CW + c]);
        atomicAdd(dbias + c, sacc[2 * CW + c]);
        for (int o = 0; o < O; ++o) atomicAdd(dw2 + (long long)o * inner + c, sacc[(3 + o) * CW + c]);
    }
    if (threadIdx.x < O) atomicAdd(db2 + threadIdx.x, sacc[(3 + O) * CW + threadIdx.x]);
}

template <typename T, int NVL>
int launch_fwd(int O, const void* x, long long ld_x, int inner, const float* gamma, const float* beta, const float* w2,
               const float* b2, int softplus, float* out, long long ppi, long long rows, cudaStream_t s) {
    long long blocks = (rows + HT_WARPS - 1) / HT_WARPS;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
#define VK_HT_FWD(OO)                                                                                            \
    head_tail_fwd_kernel<T, NVL, OO><<<(unsigned)blocks, HT_THREADS, 0, s>>>(reinterpret_cast<const T*>(x), ld_x, inner, \
                                                                             gamma, beta, w2, b2, softplus, out, ppi, rows)
    switch (O) {
        case 1: VK_HT_FWD(1); break;
        case 2: VK_HT_FWD(2); break;
        case 3: VK_HT_FWD(3); break;
        case 4: VK_HT_FWD(4); break;
        default: return 1;
    }
#undef VK_HT_FWD
    return 0;
}

template <typename T, int NVL>
int launch_bwd(int O, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
               const float* w2, int softplus, const float* out, const float* dout, long long ppi, long long rows, void* dx,
               long long ld_dx, float* dgamma, float* dbeta, float* dw2, float* db2, float* dbias, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    long long blocks = (rows + HT_WARPS - 1) / HT_WARPS;
    const long long cap = (long long)vkocr_sm_count() * 2;
    if (blocks > cap) blocks = cap;
#define VK_HT_BWD(OO)                                                                                                       \
    head_tail_bwd_kernel<T, NVL, OO><<<(unsigned)blocks, HT_THREADS, ((5 + 2 * OO) * 32 * NVL * V + OO) * sizeof(float), s>>>( \
        reinterpret_cast<const T*>(x), ld_x, inner, slice_w, gamma, beta, w2, softplus, out, dout, ppi, rows,              \
        reinterpret_cast<T*>(dx), ld_dx, dgamma, dbeta, dw2, db2, dbias)
    switch (O) {
        case 1: VK_HT_BWD(1); break;
        case 2: VK_HT_BWD(2); break;
        case 3: VK_HT_BWD(3); break;
        case 4: VK_HT_BWD(4); break;
        default: return 1;
    }
#undef VK_HT_BWD
    return 0;
}

int pick_nvl(int dtype, int slice_w) {
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const int nv = (slice_w / V + 31) / 32;
    return nv <= 1 ? 1 : (nv <= 2 ? 2 : (nv <= 4 ? 4 : 0));
}

}  // namespace

extern "C" {

// x: conv output slice [rows, slice_w] (row stride ld_x), channels >= inner are padding.
// out: fp32 NCHW (rows/pixels_per_image, O, pixels_per_image).
int vkocr_head_tail_fwd(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                        const float* w2, const float* b2, int O, int softplus, float* out, long long pixels_per_image,
                        long long rows, void* stream) {
    VK_REQUIRE(x && gamma && beta && w2 && b2 && out, VKOCR_BAD_ARGUMENT, "head_tail_fwd: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(slice_w % V == 0 && ld_x % V == 0 && slice_w >= inner, VKOCR_BAD_ALIGN, "head_tail_fwd: slice %d ld %lld", slice_w, ld_x);
    VK_REQUIRE(O >= 1 && O <= 4, VKOCR_BAD_SHAPE, "head_tail_fwd: out channels %d (1..4 supported)", O);
    const int nvl = pick_nvl(dtype, slice_w);
    VK_REQUIRE(nvl > 0, VKOCR_BAD_SHAPE, "head_tail_fwd: inner width %d too large", slice_w);
    if (rows == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
#define VK_CALL(NVL) VK_DISPATCH_DTYPE(dtype, T, (rc = launch_fwd<T, NVL>(O, x, ld_x, inner, gamma, beta, w2, b2, softplus, out, pixels_per_image, rows, s)))
    if (nvl == 1) VK_CALL(1);
    else if (nvl == 2) VK_CALL(2);
    else VK_CALL(4);
#undef VK_CALL
    VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_tail_fwd: dispatch failed");
    VK_CHECK_LAUNCH("head_tail_fwd_kernel");
    return VKOCR_OK;
}

// dx: [rows, slice_w] gradient of the conv output slice (pad channels written as zeros).
// dgamma/dbeta/dbias [inner], dw2 [O, inner], db2 [O]: fp32, accumulated into.
int vkocr_head_tail_bwd(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                        const float* w2, int O, int softplus, const float* out, const float* dout, long long pixels_per_image,
                        long long rows, void* dx, long long ld_dx, float* dgamma, float* dbeta, float* dw2, float* db2,
                        float* dbias, void* stream) {
    VK_REQUIRE(x && gamma && beta && w2 && out && dout && dx && dgamma && dbeta && dw2 && db2 && dbias, VKOCR_BAD_ARGUMENT,
               "head_tail_bwd: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(slice_w % V == 0 && ld_x % V == 0 && ld_dx % V == 0 && slice_w >= inner, VKOCR_BAD_ALIGN, "head_tail_bwd: slice %d", slice_w);
    VK_REQUIRE(O >= 1 && O <= 4, VKOCR_BAD_SHAPE, "head_tail_bwd: out channels %d (1..4 supported)", O);
    const int nvl = pick_nvl(dtype, slice_w);
    VK_REQUIRE(nvl > 0, VKOCR_BAD_SHAPE, "head_tail_bwd: inner width %d too large", slice_w);
    if (rows == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
#define VK_CALL(NVL) VK_DISPATCH_DTYPE(dtype, T, (rc = launch_bwd<T, NVL>(O, x, ld_x, inner, slice_w, gamma, beta, w2, softplus, out, dout, pixels_per_image, rows, dx, ld_dx, dgamma, dbeta, dw2, db2, dbias, s)))
    if (nvl == 1) VK_CALL(1);
    else if (nvl == 2) VK_CALL(2);
    else VK_CALL(4);
#undef VK_CALL
    VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_tail_bwd: dispatch failed");
    VK_CHECK_LAUNCH("head_tail_bwd_kernel");
    return VKOCR_OK;
}

}  // extern "C"
