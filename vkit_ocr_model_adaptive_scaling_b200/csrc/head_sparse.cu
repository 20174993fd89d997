// Label-point backward of the convolve-first heads (head_combine.cu).  Three of the four precise heads (corner offset,
// corner angle, corner distance) enter the loss only through their values at the (B, P) label points
// (AdaptiveScalingPreciseLossFunction.get_label_point_feature, loss_function/adaptive_scaling.py:167-179,235-260): their
// d(loss)/d(output) maps are exactly zero everywhere else, so d(loss)/d(conv output) is zero outside <= B*P pixels and the
// dense backward (tail, adjoint, data- and weight-gradient GEMMs over 3.3 M pixels) multiplies zeros.  With G [E, n] the
// gradient rows of the E = B*P label pixels (vkocr_head_tail_bwd_points), the exact same gradients are
//     dW_tap  = G^T . A_tap,   A_tap[e, :] = up(x)[r_e + dy - k/2, s_e + dx - k/2, :]     (vkocr_gather_up_taps + one small GEMM)
//     dX     += sum_tap  adjoint-of-up-and-shift( G . W_tap )                               (one small GEMM + vkocr_scatter_up_taps)
// i.e. two GEMMs with M = E instead of M = B*h*w (6 400 rows instead of 819 200 at B = 32 / 640x640, P = 200).
// Duplicate label points are handled by vkocr_points_claim: the upstream gradient map already holds their SUM at the shared
// pixel, so only the first entry that claims a pixel carries it.
#include "common.cuh"

namespace {

struct Axis {
    int i0, i1;
    float w0, w1;
};
// same arithmetic as head_combine.cu / resample.cu (PyTorch align_corners=False bilinear, floor nearest)
__device__ __forceinline__ Axis hs_axis(int d, int in, int out, int mode) {
    Axis a;
    const float scale = (float)in / (float)out;
    if (mode == 0) {
        float s = scale * (d + 0.5f) - 0.5f;
        if (s < 0.f) s = 0.f;
        int i0 = (int)s;
        if (i0 > in - 1) i0 = in - 1;
        a.i0 = i0;
        a.i1 = i0 + (i0 < in - 1 ? 1 : 0);
        const float l = s - i0;
        a.w0 = 1.f - l;
        a.w1 = l;
    } else {
        int i = (int)floorf(d * scale);
        a.i0 = a.i1 = i < in - 1 ? i : in - 1;
        a.w0 = 1.f;
        a.w1 = 0.f;
    }
    return a;
}
__device__ __forceinline__ float hs_weight_of(int d, int p, int in, int out, int mode) {
    const Axis a = hs_axis(d, in, out, mode);
    return (a.i0 == p ? a.w0 : 0.f) + (a.i1 == p ? a.w1 : 0.f);
}

__global__ void __launch_bounds__(256)
points_claim_kernel(const long long* __restrict__ py, const long long* __restrict__ px, int B, int P, int H, int W, int* __restrict__ owner,
                    int* __restrict__ pix_index) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * P) return;
    long long y = py[e], x = px[e];
    if (y < 0) y += H;   // torch advanced indexing wraps negative indices (as the loss kernels do)
    if (x < 0) x += W;
    int res = -1;
    if (y >= 0 && y < H && x >= 0 && x < W) {
        const int pix = (int)((long long)(e / P) * H * W + y * W + x);
        if (atomicCAS(owner + pix, 0, e + 1) == 0) res = pix;
    }
    pix_index[e] = res;
}

// a[e, tap * C + c] = up_f(x)[r + dy - pad, s + dx - pad, c]  (zero outside the up-sampled grid / for empty entries)
template <typename T>
__global__ void __launch_bounds__(256)
gather_up_taps_kernel(const T* __restrict__ x, long long ld_x, int h, int w, int C, int f, int mode, int ks, const int* __restrict__ pix_index,
                      int E, T* __restrict__ a) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const int taps = ks * ks;
    const long long total = (long long)E * taps * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    const int tap = (int)((idx / CV) % taps);
    const int e = (int)(idx / ((long long)CV * taps));
    const int H = h * f, W = w * f, pad = ks >> 1;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    const int pix = pix_index[e];
    if (pix >= 0) {
        const int b = pix / (H * W), rem = pix - b * (H * W);
        const int Y = rem / W + tap / ks - pad, X = rem % W + tap % ks - pad;
        if (Y >= 0 && Y < H && X >= 0 && X < W) {
            const Axis ay = hs_axis(Y, h, H, mode), ax = hs_axis(X, w, W, mode);
            const T* xb = x + (long long)b * h * w * ld_x + cv * V;
            const int np = mode == 0 ? 2 : 1;
            for (int iy = 0; iy < np; ++iy)
                for (int ix = 0; ix < np; ++ix) {
                    const float ww = (iy ? ay.w1 : ay.w0) * (ix ? ax.w1 : ax.w0);
                    if (ww == 0.f) continue;
                    VkVec<T> v;
                    v.load(xb + ((long long)(iy ? ay.i1 : ay.i0) * w + (ix ? ax.i1 : ax.i0)) * ld_x);
                    float fv[V];
                    v.unpack(fv);
#pragma unroll
                    for (int i = 0; i < V; ++i) acc[i] = fmaf(ww, fv[i], acc[i]);
                }
        }
    }
    VkVec<T> o;
    o.pack(acc);
    o.store(a + ((long long)e * taps + tap) * C + cv * V);
}

__device__ __forceinline__ void hs_atomic_add(float* p, float a, float b) {
    atomicAdd(p, a);
    atomicAdd(p + 1, b);
}
__device__ __forceinline__ void hs_atomic_add(__nv_bfloat16* p, float a, float b) {
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(p), __floats2bfloat162_rn(a, b));
}

// dx[b, p, q, :] += sum_tap R[r + dy - pad, p] C[s + dx - pad, q] u[e, tap, :] for the low-resolution pixels (p, q) whose
// interpolation feeds the k x k window of label pixel e.  One thread per (entry, neighbour, channel pair); the taps are
// summed in registers, so a low-resolution pixel receives ONE atomic add per entry that touches it.
constexpr int HS_NB = 5;   // neighbours per axis that a window of k <= 5 up-sampled rows at factor >= 1 can read (+1 each side)
template <typename T>
__global__ void __launch_bounds__(256)
scatter_up_taps_kernel(const T* __restrict__ u, int h, int w, int C, int f, int mode, int ks, const int* __restrict__ pix_index, int E,
                       T* __restrict__ dx, long long ld_dx) {
    const int CP = C / 2;
    const long long total = (long long)E * HS_NB * HS_NB * CP;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cp = (int)(idx % CP);
    const int nb = (int)((idx / CP) % (HS_NB * HS_NB));
    const int e = (int)(idx / ((long long)CP * HS_NB * HS_NB));
    const int pix = pix_index[e];
    if (pix < 0) return;
    const int H = h * f, W = w * f, pad = ks >> 1, taps = ks * ks;
    const int b = pix / (H * W), rem = pix - b * (H * W);
    const int r = rem / W, s = rem % W;
    // first low-resolution row / column the window can read
    const int ylo = r - pad < 0 ? 0 : r - pad, xlo = s - pad < 0 ? 0 : s - pad;
    const int p = hs_axis(ylo, h, H, mode).i0 + nb / HS_NB, q = hs_axis(xlo, w, W, mode).i0 + nb % HS_NB;
    if (p >= h || q >= w) return;
    float a0 = 0.f, a1 = 0.f;
    bool any = false;
    const T* ue = u + (long long)e * taps * C + 2 * cp;
    for (int dy = 0; dy < ks; ++dy) {
        const int Y = r + dy - pad;
        if (Y < 0 || Y >= H) continue;
        const float wy = hs_weight_of(Y, p, h, H, mode);
        if (wy == 0.f) continue;
        for (int dxx = 0; dxx < ks; ++dxx) {
            const int X = s + dxx - pad;
            if (X < 0 || X >= W) continue;
            const float ww = wy * hs_weight_of(X, q, w, W, mode);
            if (ww == 0.f) continue;
            const T* up = ue + (long long)(dy * ks + dxx) * C;
            a0 = fmaf(ww, vk_to_f32(up[0]), a0);
            a1 = fmaf(ww, vk_to_f32(up[1]), a1);
            any = true;
        }
    }
    if (any) hs_atomic_add(dx + (((long long)b * h + p) * w + q) * ld_dx + 2 * cp, a0, a1);
}

// The same with one 16-byte channel vector per thread (C a multiple of the vector): the interpolation weights of a
// (label pixel, neighbour) -- a dozen divisions and compares -- are computed for 8 channels instead of 2, the tap vectors
// come in as 16-byte loads.  6 400 label pixels x 25 neighbours x 384 channels: 0.52 -> 0.1x ms.
template <typename T>
__global__ void __launch_bounds__(256)
scatter_up_taps_vec_kernel(const T* __restrict__ u, int h, int w, int C, int f, int mode, int ks, const int* __restrict__ pix_index, int E,
                           T* __restrict__ dx, long long ld_dx) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const long long total = (long long)E * HS_NB * HS_NB * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    const int nb = (int)((idx / CV) % (HS_NB * HS_NB));
    const int e = (int)(idx / ((long long)CV * HS_NB * HS_NB));
    const int pix = pix_index[e];
    if (pix < 0) return;
    const int H = h * f, W = w * f, pad = ks >> 1, taps = ks * ks;
    const int b = pix / (H * W), rem = pix - b * (H * W);
    const int r = rem / W, s = rem % W;
    const int ylo = r - pad < 0 ? 0 : r - pad, xlo = s - pad < 0 ? 0 : s - pad;
    const int p = hs_axis(ylo, h, H, mode).i0 + nb / HS_NB, q = hs_axis(xlo, w, W, mode).i0 + nb % HS_NB;
    if (p >= h || q >= w) return;
    float wx[5];
#pragma unroll
    for (int dxx = 0; dxx < 5; ++dxx) {
        const int X = s + dxx - pad;
        wx[dxx] = (dxx < ks && X >= 0 && X < W) ? hs_weight_of(X, q, w, W, mode) : 0.f;
    }
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    bool any = false;
    const T* ue = u + (long long)e * taps * C + cv * V;
    for (int dy = 0; dy < ks; ++dy) {
        const int Y = r + dy - pad;
        if (Y < 0 || Y >= H) continue;
        const float wy = hs_weight_of(Y, p, h, H, mode);
        if (wy == 0.f) continue;
#pragma unroll
        for (int dxx = 0; dxx < 5; ++dxx) {
            const float ww = wy * wx[dxx];
            if (dxx >= ks || ww == 0.f) continue;
            VkVec<T> vec;
            vec.load(ue + (long long)(dy * ks + dxx) * C);
            float fv[V];
            vec.unpack(fv);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = fmaf(ww, fv[i], acc[i]);
            any = true;
        }
    }
    if (any) {
        T* dst = dx + (((long long)b * h + p) * w + q) * ld_dx + cv * V;
#pragma unroll
        for (int i = 0; i < V; i += 2) hs_atomic_add(dst + i, acc[i], acc[i + 1]);
    }
}

}  // namespace

extern "C" {

// py / px: (B, P) int64 label points on the H x W map; owner: B*H*W ints, zeroed by the caller; pix_index[e] receives the
// global pixel index b*H*W + y*W + x of entry e if it is the first entry on that pixel, -1 otherwise (duplicate or outside).
int vkocr_points_claim(const long long* py, const long long* px, int B, int P, int H, int W, int* owner, int* pix_index, void* stream) {
    VK_REQUIRE(py && px && owner && pix_index, VKOCR_BAD_ARGUMENT, "points_claim: null argument");
    VK_REQUIRE((long long)B * H * W < (1LL << 31) && (long long)B * P < (1LL << 31), VKOCR_BAD_SHAPE, "points_claim: map or point list too large");
    if ((long long)B * P == 0) return VKOCR_OK;
    points_claim_kernel<<<(unsigned)(((long long)B * P + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(py, px, B, P, H, W, owner,
                                                                                                                       pix_index);
    VK_CHECK_LAUNCH("points_claim_kernel");
    return VKOCR_OK;
}

// x: [B*h*w, ld_x] low-resolution activations; a: [E, ks*ks*C] (storage dtype) receives, per entry and tap, the
// up-sampled (x factor, mode 0 bilinear / 1 nearest) and tap-shifted input vector of that label pixel.
int vkocr_gather_up_taps(int dtype, const void* x, long long ld_x, int B, int h, int w, int C, int factor, int mode, int ks,
                         const int* pix_index, int E, void* a, void* stream) {
    VK_REQUIRE(x && pix_index && a, VKOCR_BAD_ARGUMENT, "gather_up_taps: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(C % V == 0 && ld_x % V == 0 && factor >= 1 && (mode == 0 || mode == 1) && (ks == 1 || ks == 3 || ks == 5), VKOCR_BAD_SHAPE,
               "gather_up_taps: C %d factor %d mode %d kernel %d", C, factor, mode, ks);
    const long long total = (long long)E * ks * ks * (C / V);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (gather_up_taps_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, h, w, C, factor, mode, ks, pix_index, E, reinterpret_cast<T*>(a))));
    VK_CHECK_LAUNCH("gather_up_taps_kernel");
    return VKOCR_OK;
}

// u: [E, ks*ks*C] per-entry, per-tap gradient vectors w.r.t. the up-sampled input; dx: [B*h*w, ld_dx] is ACCUMULATED into
// (atomic adds in the storage dtype) -- run it after the dense data-gradient GEMM has written dx.
int vkocr_scatter_up_taps(int dtype, const void* u, int B, int h, int w, int C, int factor, int mode, int ks, const int* pix_index, int E,
                          void* dx, long long ld_dx, void* stream) {
    VK_REQUIRE(u && pix_index && dx, VKOCR_BAD_ARGUMENT, "scatter_up_taps: null argument");
    VK_REQUIRE(C % 2 == 0 && ld_dx % 2 == 0 && factor >= 1 && (mode == 0 || mode == 1) && (ks == 1 || ks == 3 || ks == 5), VKOCR_BAD_SHAPE,
               "scatter_up_taps: C %d factor %d mode %d kernel %d", C, factor, mode, ks);
    (void)B;
    const long long total = (long long)E * HS_NB * HS_NB * (C / 2);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    if (C % V == 0 && ld_dx % V == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
        const long long tv = (long long)E * HS_NB * HS_NB * (C / V);
        VK_DISPATCH_DTYPE(dtype, T, (scatter_up_taps_vec_kernel<T><<<(unsigned)((tv + 255) / 256), 256, 0, s>>>(
                                        reinterpret_cast<const T*>(u), h, w, C, factor, mode, ks, pix_index, E, reinterpret_cast<T*>(dx), ld_dx)));
        VK_CHECK_LAUNCH("scatter_up_taps_vec_kernel");
        return VKOCR_OK;
    }
    VK_DISPATCH_DTYPE(dtype, T, (scatter_up_taps_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(u), h, w, C, factor, mode, ks, pix_index, E, reinterpret_cast<T*>(dx), ld_dx)));
    VK_CHECK_LAUNCH("scatter_up_taps_kernel");
    return VKOCR_OK;
}

}  // extern "C"
