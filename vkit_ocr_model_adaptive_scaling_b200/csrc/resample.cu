// Memory-movement kernels of the neck / stem on NHWC activations:
//   * bilinear / nearest up-sampling, forward (overwrite or add into a channel slice) and gather-form backward
//     (F.interpolate(mode='bilinear'|'nearest', align_corners=False): upernext.py:79,178,191,237; fpn.py:125,138,197)
//   * AdaptiveAvgPool2d forward/backward (upernext.py:59-65)
//   * patchify gathers for the stride==kernel convolutions (helper.py:43-58), image and NHWC variants
//   * channel-slice copy (torch.cat replacement: upernext.py:82,197; fpn.py:144)
// All are HBM-bound: one thread per (pixel, 16-byte channel vector), coalesced along C.
#include "common.cuh"

namespace {

struct Axis {
    int i0, i1;
    float w0, w1;
};

// PyTorch area_pixel_compute_source_index(align_corners=False) + clamp, for one destination coordinate.
__device__ __forceinline__ Axis bilinear_axis(int d, int in, int out) {
    const float scale = (float)in / (float)out;
    float s = scale * (d + 0.5f) - 0.5f;
    if (s < 0.f) s = 0.f;
    int i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    const int i1 = i0 + (i0 < in - 1 ? 1 : 0);
    const float l = s - i0;
    Axis a;
    a.i0 = i0; a.i1 = i1; a.w0 = 1.f - l; a.w1 = l;
    return a;
}
__device__ __forceinline__ int nearest_src(int d, int in, int out) {
    const float scale = (float)in / (float)out;
    int i = (int)floorf(d * scale);
    return i < in - 1 ? i : in - 1;
}

template <typename T, typename VT>
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const T* __restrict__ src, long long ld_s, int h, int w, T* __restrict__ dst, long long ld_d, int H, int W,
                    int B, int C, int mode, int accumulate) {
    constexpr int V = VT::N;
    const int CV = C / V;
    const long long total = (long long)B * H * W * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int X = (int)(r % W);
    r /= W;
    const int Y = (int)(r % H);
    const int b = (int)(r / H);
    const int c = cv * V;
    float o[V];
    const T* sb = src + (long long)b * h * w * ld_s + c;
    if (mode == 0) {
        const Axis ay = bilinear_axis(Y, h, H), ax = bilinear_axis(X, w, W);
        VT v00, v01, v10, v11;
        v00.load(sb + ((long long)ay.i0 * w + ax.i0) * ld_s);
        v01.load(sb + ((long long)ay.i0 * w + ax.i1) * ld_s);
        v10.load(sb + ((long long)ay.i1 * w + ax.i0) * ld_s);
        v11.load(sb + ((long long)ay.i1 * w + ax.i1) * ld_s);
        float f00[V], f01[V], f10[V], f11[V];
        v00.unpack(f00); v01.unpack(f01); v10.unpack(f10); v11.unpack(f11);
#pragma unroll
        for (int i = 0; i < V; ++i)
            o[i] = ay.w0 * (ax.w0 * f00[i] + ax.w1 * f01[i]) + ay.w1 * (ax.w0 * f10[i] + ax.w1 * f11[i]);
    } else {
        const int sy = nearest_src(Y, h, H), sx = nearest_src(X, w, W);
        VT v;
        v.load(sb + ((long long)sy * w + sx) * ld_s);
        v.unpack(o);
    }
    T* dp = dst + (((long long)b * H + Y) * W + X) * ld_d + c;
    if (accumulate) {
        VT d;
        d.load(dp);
        float f[V];
        d.unpack(f);
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] += f[i];
    }
    VT out;
    out.pack(o);
    out.store(dp);
}

// Gather-form adjoint: dsrc[b,i,j,:] (+)= sum over destination pixels that read (i,j) of weight * ddst.
template <typename T, typename VT>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const T* __restrict__ ddst, long long ld_d, int H, int W, T* __restrict__ dsrc, long long ld_s, int h, int w,
                    int B, int C, int mode, int accumulate) {
    constexpr int V = VT::N;
    const int CV = C / V;
    const long long total = (long long)B * h * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    r /= w;
    const int i = (int)(r % h);
    const int b = (int)(r / h);
    const int c = cv * V;
    // conservative candidate ranges of destination rows/cols touching source row i / col j
    const float ry = (float)H / (float)h, rx = (float)W / (float)w;
    // a destination row Y reads source rows floor(s), floor(s)+1 with s = (Y + .5) / ry - .5, so source row i is read by
    // the Y with s in (i - 1, i + 1): Y + .5 in ((i - .5) ry, (i + 1.5) ry); one extra row each side absorbs rounding
    // (nearest: the Y with floor(Y / ry) == i lie inside the same range)
    int Ylo = (int)floorf((i - 0.5f) * ry - 0.5f) - 1, Yhi = (int)ceilf((i + 1.5f) * ry - 0.5f) + 1;
    int Xlo = (int)floorf((j - 0.5f) * rx - 0.5f) - 1, Xhi = (int)ceilf((j + 1.5f) * rx - 0.5f) + 1;
    if (Ylo < 0) Ylo = 0;
    if (Xlo < 0) Xlo = 0;
    if (Yhi > H - 1) Yhi = H - 1;
    if (Xhi > W - 1) Xhi = W - 1;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* db = ddst + (long long)b * H * W * ld_d + c;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
        float wy;
        if (mode == 0) {
            const Axis a = bilinear_axis(Y, h, H);
            wy = (a.i0 == i ? a.w0 : 0.f) + (a.i1 == i ? a.w1 : 0.f);
        } else {
            wy = nearest_src(Y, h, H) == i ? 1.f : 0.f;
        }
        if (wy == 0.f) continue;
        for (int X = Xlo; X <= Xhi; ++X) {
            float wx;
            if (mode == 0) {
                const Axis a = bilinear_axis(X, w, W);
                wx = (a.i0 == j ? a.w0 : 0.f) + (a.i1 == j ? a.w1 : 0.f);
            } else {
                wx = nearest_src(X, w, W) == j ? 1.f : 0.f;
            }
            if (wx == 0.f) continue;
            VT v;
            v.load(db + ((long long)Y * W + X) * ld_d);
            float f[V];
            v.unpack(f);
            const float ww = wy * wx;
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] = fmaf(ww, f[e], acc[e]);
        }
    }
    T* sp = dsrc + (((long long)b * h + i) * w + j) * ld_s + c;
    if (accumulate) {
        VT d;
        d.load(sp);
        float f[V];
        d.unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += f[e];
    }
    VT out;
    out.pack(acc);
    out.store(sp);
}

// Separable form of the same adjoint for large scale factors (the PPM / top-down paths of the UperNeXt neck: 160 -> 20,
// 160 -> 40, 20 -> 1..6): the weights factor (wy * wx), so the gradient is first reduced along x into an fp32 workspace
// [B, H, w, C] -- one thread per (row, source column, channel vector), consecutive threads on consecutive channel vectors:
// coalesced, 1.2 M threads at 160 -> 20 -- and then along y.  The one-pass gather has only B*h*w*C/V threads with a
// (2 r + 2)^2-pixel window each: 0.65 TB/s at r = 8, 15 GB/s at 20 -> 1.
__device__ __forceinline__ void adjoint_range(int i, int n_src, int n_dst, int* lo, int* hi) {
    const float r = (float)n_dst / (float)n_src;
    int a = (int)floorf((i - 0.5f) * r - 0.5f) - 1, b = (int)ceilf((i + 1.5f) * r - 0.5f) + 1;
    *lo = a < 0 ? 0 : a;
    *hi = b > n_dst - 1 ? n_dst - 1 : b;
}
__device__ __forceinline__ float adjoint_weight(int d, int i, int n_src, int n_dst, int mode) {
    if (mode == 0) {
        const Axis a = bilinear_axis(d, n_src, n_dst);
        return (a.i0 == i ? a.w0 : 0.f) + (a.i1 == i ? a.w1 : 0.f);
    }
    return nearest_src(d, n_src, n_dst) == i ? 1.f : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample_bwd_x_kernel(const T* __restrict__ ddst, long long ld_d, int H, int W, float* __restrict__ ws, int w, int B, int CV, int mode) {
    constexpr int V = VkVec<T>::N;
    const long long total = (long long)B * H * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    const long long row = r / w;                      // b * H + Y
    int Xlo, Xhi;
    adjoint_range(j, w, W, &Xlo, &Xhi);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* rp = ddst + row * W * ld_d + cv * V;
    for (int X = Xlo; X <= Xhi; ++X) {
        const float wx = adjoint_weight(X, j, w, W, mode);
        if (wx == 0.f) continue;
        VkVec<T> v;
        v.load(rp + (long long)X * ld_d);
        float f[V];
        v.unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = fmaf(wx, f[e], acc[e]);
    }
    float* o = ws + idx * V;
#pragma unroll
    for (int e = 0; e < V; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample_bwd_y_kernel(const float* __restrict__ ws, int H, T* __restrict__ dsrc, long long ld_s, int h, int w, int B, int CV, int mode,
                      int accumulate) {
    constexpr int V = VkVec<T>::N;
    const long long total = (long long)B * h * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    r /= w;
    const int i = (int)(r % h);
    const int b = (int)(r / h);
    int Ylo, Yhi;
    adjoint_range(i, h, H, &Ylo, &Yhi);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
        const float wy = adjoint_weight(Y, i, h, H, mode);
        if (wy == 0.f) continue;
        const float* p = ws + ((((long long)b * H + Y) * w + j) * CV + cv) * V;
#pragma unroll
        for (int e = 0; e < V; e += 4) {
            const float4 f = *reinterpret_cast<const float4*>(p + e);
            acc[e] = fmaf(wy, f.x, acc[e]); acc[e + 1] = fmaf(wy, f.y, acc[e + 1]);
            acc[e + 2] = fmaf(wy, f.z, acc[e + 2]); acc[e + 3] = fmaf(wy, f.w, acc[e + 3]);
        }
    }
    T* sp = dsrc + (((long long)b * h + i) * w + j) * ld_s + cv * V;
    if (accumulate) {
        VkVec<T> d;
        d.load(sp);
        float f[V];
        d.unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += f[e];
    }
    VkVec<T> out;
    out.pack(acc);
    out.store(sp);
}

// ------------------------------------------------------------------------------------------------- exact x2 nearest
// F.interpolate(scale 2, mode='nearest') of the FPN top-down path and head inputs (fpn.py:121-144,197-204):
// dst[2i+p, 2j+q] = src[i, j]; the adjoint sums the 2x2 block.  One thread = one 16-byte channel vector of one source
// pixel (consecutive threads on consecutive channel vectors: every access is a coalesced row segment).
template <typename T>
__global__ void __launch_bounds__(256)
nearest2x_fwd_kernel(const T* __restrict__ src, long long ld_s, int h, int w, T* __restrict__ dst, long long ld_d, int B, int CV,
                     int accumulate) {
    constexpr int V = VkVec<T>::N;
    const long long total = (long long)B * h * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    r /= w;
    const int i = (int)(r % h);
    const long long b = r / h;
    VkVec<T> v;
    v.load(src + ((b * h + i) * w + j) * ld_s + cv * V);
    float f[V];
    if (accumulate) v.unpack(f);
    T* d0 = dst + ((b * 2 * h + 2 * i) * (2LL * w) + 2 * j) * ld_d + cv * V;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            T* dp = d0 + ((long long)p * 2 * w + q) * ld_d;
            if (accumulate) {
                VkVec<T> o;
                o.load(dp);
                float g[V];
                o.unpack(g);
#pragma unroll
                for (int e = 0; e < V; ++e) g[e] += f[e];
                o.pack(g);
                o.store(dp);
            } else {
                v.store(dp);
            }
        }
}

template <typename T>
__global__ void __launch_bounds__(256)
nearest2x_bwd_kernel(const T* __restrict__ ddst, long long ld_d, T* __restrict__ dsrc, long long ld_s, int h, int w, int B, int CV,
                     int accumulate) {
    constexpr int V = VkVec<T>::N;
    const long long total = (long long)B * h * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    r /= w;
    const int i = (int)(r % h);
    const long long b = r / h;
    const T* d0 = ddst + ((b * 2 * h + 2 * i) * (2LL * w) + 2 * j) * ld_d + cv * V;
    VkVec<T> v[4];
    v[0].load(d0);
    v[1].load(d0 + ld_d);
    v[2].load(d0 + 2LL * w * ld_d);
    v[3].load(d0 + (2LL * w + 1) * ld_d);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float f[V];
        v[k].unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += f[e];
    }
    T* sp = dsrc + ((b * h + i) * w + j) * ld_s + cv * V;
    if (accumulate) {
        VkVec<T> o;
        o.load(sp);
        float f[V];
        o.unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += f[e];
    }
    VkVec<T> out;
    out.pack(acc);
    out.store(sp);
}

// ------------------------------------------------------------------------------------------------- exact x2 bilinear
// F.interpolate(scale 2, mode='bilinear', align_corners=False) is the separable 2-tap filter {0.25, 0.75} with the source
// index clamped at the borders (out[2k] = .25 in[k-1] + .75 in[k], out[2k+1] = .75 in[k] + .25 in[k+1]); its adjoint is
// the 4-tap filter {.25, .75, .75, .25} over the gradient rows/cols 2i-1 .. 2i+2 with the GRADIENT index clamped.  These
// carry the 2.5 GB head inputs of the adaptive-scaling heads (upernext.py:237-244) and every top-down step.
// One thread = one 16-byte channel vector of one source column, sliding down ROWS source rows with the horizontally
// filtered rows kept in registers: 3 loads per 4 stores (forward), 8 loads per store (backward, served by L1/L2).
constexpr int UP_ROWS = 8;

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const T* __restrict__ src, long long ld_s, int h, int w, T* __restrict__ dst, long long ld_d, int B, int CV,
                      int accumulate) {
    constexpr int V = VkVec<T>::N;
    const int chunks = (h + UP_ROWS - 1) / UP_ROWS;
    const long long total = (long long)B * chunks * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    r /= w;
    const int ch = (int)(r % chunks);
    const int b = (int)(r / chunks);
    const int c = cv * V;
    const int jm = j > 0 ? j - 1 : 0, jp = j < w - 1 ? j + 1 : w - 1;
    const int i0 = ch * UP_ROWS, i1 = (i0 + UP_ROWS < h) ? i0 + UP_ROWS : h;
    const T* sb = src + (long long)b * h * w * ld_s + c;
    const int W = 2 * w;
    T* db = dst + (long long)b * (2 * h) * W * ld_d + c;
    // he / ho: horizontally filtered source row for the even / odd output column of this source column
    auto hrow = [&](int i, float* he, float* ho) {
        const T* row = sb + (long long)i * w * ld_s;
        VkVec<T> vm, v0, vp;
        vm.load(row + (long long)jm * ld_s);
        v0.load(row + (long long)j * ld_s);
        vp.load(row + (long long)jp * ld_s);
        float fm[V], f0[V], fp[V];
        vm.unpack(fm); v0.unpack(f0); vp.unpack(fp);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            he[e] = fmaf(0.25f, fm[e], 0.75f * f0[e]);
            ho[e] = fmaf(0.25f, fp[e], 0.75f * f0[e]);
        }
    };
    auto emit = [&](int Y, int X, const float* a, const float* bb, float wa, float wb) {
        T* dp = db + ((long long)Y * W + X) * ld_d;
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = fmaf(wa, a[e], wb * bb[e]);
        if (accumulate) {
            VkVec<T> d;
            d.load(dp);
            float f[V];
            d.unpack(f);
#pragma unroll
            for (int e = 0; e < V; ++e) o[e] += f[e];
        }
        VkVec<T> out;
        out.pack(o);
        out.store(dp);
    };
    float pe[V], po[V], ce[V], co[V], ne[V], no[V];
    hrow(i0 > 0 ? i0 - 1 : 0, pe, po);
    hrow(i0, ce, co);
    for (int i = i0; i < i1; ++i) {
        hrow(i < h - 1 ? i + 1 : h - 1, ne, no);
        emit(2 * i, 2 * j, pe, ce, 0.25f, 0.75f);
        emit(2 * i, 2 * j + 1, po, co, 0.25f, 0.75f);
        emit(2 * i + 1, 2 * j, ce, ne, 0.75f, 0.25f);
        emit(2 * i + 1, 2 * j + 1, co, no, 0.75f, 0.25f);
#pragma unroll
        for (int e = 0; e < V; ++e) { pe[e] = ce[e]; po[e] = co[e]; ce[e] = ne[e]; co[e] = no[e]; }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T* __restrict__ ddst, long long ld_d, T* __restrict__ dsrc, long long ld_s, int h, int w, int B, int CV,
                      int accumulate) {
    constexpr int V = VkVec<T>::N;
    const int chunks = (h + UP_ROWS - 1) / UP_ROWS;
    const long long total = (long long)B * chunks * w * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % w);
    r /= w;
    const int ch = (int)(r % chunks);
    const int b = (int)(r / chunks);
    const int c = cv * V;
    const int H = 2 * h, W = 2 * w;
    const int x0 = 2 * j > 0 ? 2 * j - 1 : 0, x1 = 2 * j, x2 = 2 * j + 1, x3 = 2 * j + 2 < W ? 2 * j + 2 : W - 1;
    const int i0 = ch * UP_ROWS, i1 = (i0 + UP_ROWS < h) ? i0 + UP_ROWS : h;
    const T* db = ddst + (long long)b * H * W * ld_d + c;
    T* sb = dsrc + (long long)b * h * w * ld_s + c;
    // g: horizontally filtered gradient row (columns 2j-1 .. 2j+2, clamped)
    auto grow = [&](int Y, float* g) {
        Y = Y < 0 ? 0 : (Y > H - 1 ? H - 1 : Y);
        const T* row = db + (long long)Y * W * ld_d;
        VkVec<T> v0, v1, v2, v3;
        v0.load(row + (long long)x0 * ld_d);
        v1.load(row + (long long)x1 * ld_d);
        v2.load(row + (long long)x2 * ld_d);
        v3.load(row + (long long)x3 * ld_d);
        float f0[V], f1[V], f2[V], f3[V];
        v0.unpack(f0); v1.unpack(f1); v2.unpack(f2); v3.unpack(f3);
#pragma unroll
        for (int e = 0; e < V; ++e) g[e] = fmaf(0.25f, f0[e] + f3[e], 0.75f * (f1[e] + f2[e]));
    };
    float ga[V], gb[V], gc[V], gd[V];
    grow(2 * i0 - 1, ga);
    grow(2 * i0, gb);
    for (int i = i0; i < i1; ++i) {
        grow(2 * i + 1, gc);
        grow(2 * i + 2, gd);
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = fmaf(0.25f, ga[e] + gd[e], 0.75f * (gb[e] + gc[e]));
        T* sp = sb + ((long long)i * w + j) * ld_s;
        if (accumulate) {
            VkVec<T> d;
            d.load(sp);
            float f[V];
            d.unpack(f);
#pragma unroll
            for (int e = 0; e < V; ++e) o[e] += f[e];
        }
        VkVec<T> out;
        out.pack(o);
        out.store(sp);
#pragma unroll
        for (int e = 0; e < V; ++e) { ga[e] = gc[e]; gb[e] = gd[e]; }
    }
}

// AdaptiveAvgPool2d: bin i = [floor(i*in/s), ceil((i+1)*in/s))
__device__ __forceinline__ int bin_lo(int i, int in, int s) { return (i * in) / s; }
__device__ __forceinline__ int bin_hi(int i, int in, int s) { return ((i + 1) * in + s - 1) / s; }

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const T* __restrict__ x, long long ld_x, int H, int W, T* __restrict__ y, long long ld_y, int S, int B, int C) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const long long total = (long long)B * S * S * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % S);
    r /= S;
    const int i = (int)(r % S);
    const int b = (int)(r / S);
    const int c = cv * V;
    const int y0 = bin_lo(i, H, S), y1 = bin_hi(i, H, S), x0 = bin_lo(j, W, S), x1 = bin_hi(j, W, S);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
            VkVec<T> v;
            v.load(x + (((long long)b * H + yy) * W + xx) * ld_x + c);
            float f[V];
            v.unpack(f);
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] += f[e];
        }
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] *= inv;
    VkVec<T> out;
    out.pack(acc);
    out.store(y + (((long long)b * S + i) * S + j) * ld_y + c);
}

// Separable form for large bins (PPM scales 1..6 on a 20x20 .. 64x64 map: a bin holds up to 4096 pixels and there are only
// B * S * S * C / V outputs, so one thread per output is a serial, latency-bound sum: 0.2 TB/s at 2048^2 inference).
// Pass 1 sums every image row over the bin's columns into fp32 ws[B, H, S, C]; pass 2 sums the bin's rows and scales.
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_rows_kernel(const T* __restrict__ x, long long ld_x, int H, int W, float* __restrict__ ws, int S, int B, int CV) {
    constexpr int V = VkVec<T>::N;
    const long long total = (long long)B * H * S * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % S);
    const long long row = r / S;                        // b * H + yy
    const int x0 = bin_lo(j, W, S), x1 = bin_hi(j, W, S);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* rp = x + row * W * ld_x + cv * V;
    for (int xx = x0; xx < x1; ++xx) {
        VkVec<T> v;
        v.load(rp + (long long)xx * ld_x);
        float f[V];
        v.unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += f[e];
    }
    float* o = ws + idx * V;
#pragma unroll
    for (int e = 0; e < V; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_cols_kernel(const float* __restrict__ ws, int H, int W, T* __restrict__ y, long long ld_y, int S, int B, int CV) {
    constexpr int V = VkVec<T>::N;
    const long long total = (long long)B * S * S * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % S);
    r /= S;
    const int i = (int)(r % S);
    const int b = (int)(r / S);
    const int y0 = bin_lo(i, H, S), y1 = bin_hi(i, H, S), x0 = bin_lo(j, W, S), x1 = bin_hi(j, W, S);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int yy = y0; yy < y1; ++yy) {
        const float* p = ws + ((((long long)b * H + yy) * S + j) * CV + cv) * V;
#pragma unroll
        for (int e = 0; e < V; e += 4) {
            const float4 f = *reinterpret_cast<const float4*>(p + e);
            acc[e] += f.x; acc[e + 1] += f.y; acc[e + 2] += f.z; acc[e + 3] += f.w;
        }
    }
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] *= inv;
    VkVec<T> out;
    out.pack(acc);
    out.store(y + (((long long)b * S + i) * S + j) * ld_y + cv * V);
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const T* __restrict__ dy, long long ld_y, int S, T* __restrict__ dx, long long ld_x, int H, int W, int B, int C,
                   int accumulate) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const long long total = (long long)B * H * W * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int xx = (int)(r % W);
    r /= W;
    const int yy = (int)(r % H);
    const int b = (int)(r / H);
    const int c = cv * V;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int i = 0; i < S; ++i) {
        const int y0 = bin_lo(i, H, S), y1 = bin_hi(i, H, S);
        if (yy < y0 || yy >= y1) continue;
        for (int j = 0; j < S; ++j) {
            const int x0 = bin_lo(j, W, S), x1 = bin_hi(j, W, S);
            if (xx < x0 || xx >= x1) continue;
            VkVec<T> v;
            v.load(dy + (((long long)b * S + i) * S + j) * ld_y + c);
            float f[V];
            v.unpack(f);
            const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] = fmaf(inv, f[e], acc[e]);
        }
    }
    T* dp = dx + (((long long)b * H + yy) * W + xx) * ld_x + c;
    if (accumulate) {
        VkVec<T> d;
        d.load(dp);
        float f[V];
        d.unpack(f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += f[e];
    }
    VkVec<T> out;
    out.pack(acc);
    out.store(dp);
}

// image (B,Cin,H,W) fp32 NCHW -> rows [B*Ho*Wo, c_pad], k = (ky*p + kx)*Cin + ch, zero-padded to c_pad
template <typename T>
__global__ void __launch_bounds__(256)
patchify_image_kernel(const float* __restrict__ img, int B, int Cin, int H, int W, int p, T* __restrict__ out, int c_pad) {
    const int Ho = H / p, Wo = W / p;
    const long long total = (long long)B * Ho * Wo * c_pad;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx % c_pad);
    long long m = idx / c_pad;
    const int xo = (int)(m % Wo);
    m /= Wo;
    const int yo = (int)(m % Ho);
    const int b = (int)(m / Ho);
    float v = 0.f;
    if (k < p * p * Cin) {
        const int ch = k % Cin;
        const int kx = (k / Cin) % p;
        const int ky = k / (Cin * p);
        v = img[(((long long)b * Cin + ch) * H + yo * p + ky) * W + xo * p + kx];
    }
    out[idx] = vk_from_f32<T>(v);
}

// The stem's case (3 channels, 4 x 4 patches, convnext.py:106-123): one thread = one patch row (ky) of one patch: three
// coalesced float4 loads (one per colour plane), 12 bf16 values (24 B) written to k = (ky*4 + kx)*3 + ch.
__global__ void __launch_bounds__(256)
patchify_rgb4_kernel(const float* __restrict__ img, int B, int H, int W, __nv_bfloat16* __restrict__ out, int c_pad) {
    const int Ho = H / 4, Wo = W / 4;
    const long long total = (long long)B * Ho * Wo * 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // thread order: xo fastest (coalesced image reads), then ky, then yo, then b
    const int xo = (int)(idx % Wo);
    long long r = idx / Wo;
    const int ky = (int)(r & 3);
    r >>= 2;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const long long plane = (long long)H * W;
    const float* src = img + (long long)b * 3 * plane + (long long)(yo * 4 + ky) * W + xo * 4;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(src + plane));
    const float4 c2 = __ldg(reinterpret_cast<const float4*>(src + 2 * plane));
    const float v[12] = {c0.x, c1.x, c2.x, c0.y, c1.y, c2.y, c0.z, c1.z, c2.z, c0.w, c1.w, c2.w};
    __nv_bfloat16* dst = out + (((long long)b * Ho + yo) * Wo + xo) * c_pad + ky * 12;
    uint2* d2 = reinterpret_cast<uint2*>(dst);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        uint2 raw;
        *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(v[4 * i], v[4 * i + 1]);
        *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(v[4 * i + 2], v[4 * i + 3]);
        d2[i] = raw;
    }
    if (ky == 0)
        for (int k = 48; k < c_pad; k += 4) *reinterpret_cast<uint2*>(dst + k) = make_uint2(0u, 0u);   // K padding
}

// NHWC (B,H,W,C) -> (B,H/2,W/2,4C) with k = (ky*2+kx)*C + c  (dir 0), or the adjoint scatter (dir 1; pixels of an odd
// trailing row/column get zero).  accumulate only applies to dir 1.
template <typename T>
__global__ void __launch_bounds__(256)
space_to_depth2_kernel(T* __restrict__ x, long long ld_x, int B, int H, int W, int C, T* __restrict__ y, long long ld_y, int dir,
                       int accumulate) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const long long total = (long long)B * H * W * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int xx = (int)(r % W);
    r /= W;
    const int yy = (int)(r % H);
    const int b = (int)(r / H);
    const int c = cv * V;
    const int Ho = H / 2, Wo = W / 2;
    const int yo = yy >> 1, xo = xx >> 1;
    const bool inside = yo < Ho && xo < Wo;
    T* xp = x + (((long long)b * H + yy) * W + xx) * ld_x + c;
    T* yp = y + (((long long)b * Ho + yo) * Wo + xo) * ld_y + ((yy & 1) * 2 + (xx & 1)) * C + c;
    if (dir == 0) {
        if (!inside) return;
        VkVec<T> v;
        v.load(xp);
        v.store(yp);
    } else {
        float f[V];
        if (inside) {
            VkVec<T> v;
            v.load(yp);
            v.unpack(f);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = 0.f;
        }
        if (accumulate) {
            VkVec<T> d;
            d.load(xp);
            float g[V];
            d.unpack(g);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] += g[e];
        }
        VkVec<T> out;
        out.pack(f);
        out.store(xp);
    }
}

// dst[r, 0:C] (=|+=) src[r, 0:C]  with independent row strides (channel-slice concat / gradient accumulation)
template <typename T>
__global__ void __launch_bounds__(256)
copy_channels_kernel(const T* __restrict__ src, long long ld_s, T* __restrict__ dst, long long ld_d, long long rows, int C,
                     int accumulate, int vec) {
    constexpr int V = VkVec<T>::N;
    const int CV = vec ? C / V : C;
    const long long total = rows * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    const long long r = idx / CV;
    if (vec) {
        VkVec<T> v;
        v.load(src + r * ld_s + cv * V);
        if (accumulate) {
            VkVec<T> d;
            d.load(dst + r * ld_d + cv * V);
            float f[V], g[V];
            v.unpack(f);
            d.unpack(g);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] += g[e];
            v.pack(f);
        }
        v.store(dst + r * ld_d + cv * V);
    } else {
        float f = vk_to_f32(src[r * ld_s + cv]);
        if (accumulate) f += vk_to_f32(dst[r * ld_d + cv]);
        dst[r * ld_d + cv] = vk_from_f32<T>(f);
    }
}

inline bool vec_ok(int dtype, int C, long long a, long long b) {
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    return C % V == 0 && a % V == 0 && b % V == 0;
}
inline bool ptr16(const void* a, const void* b) {
    return (reinterpret_cast<uintptr_t>(a) % 16 == 0) && (reinterpret_cast<uintptr_t>(b) % 16 == 0);
}

template <typename T, typename VT>
void launch_upsample_fwd(const void* src, long long ld_s, int h, int w, void* dst, long long ld_d, int H, int W, int B, int C,
                         int mode, int accumulate, cudaStream_t s) {
    const long long total = (long long)B * H * W * (C / VT::N);
    upsample_fwd_kernel<T, VT><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const T*>(src), ld_s, h, w, reinterpret_cast<T*>(dst), ld_d, H, W, B, C, mode, accumulate);
}
template <typename T, typename VT>
void launch_upsample_bwd(const void* ddst, long long ld_d, int H, int W, void* dsrc, long long ld_s, int h, int w, int B, int C,
                         int mode, int accumulate, cudaStream_t s) {
    const long long total = (long long)B * h * w * (C / VT::N);
    upsample_bwd_kernel<T, VT><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const T*>(ddst), ld_d, H, W, reinterpret_cast<T*>(dsrc), ld_s, h, w, B, C, mode, accumulate);
}

}  // namespace

extern "C" {

// mode 0 bilinear (align_corners=False), 1 nearest.  dst[b,Y,X,0:C] (=|+=) resample(src)[b,Y,X,0:C]
int vkocr_upsample_fwd(int dtype, const void* src, long long ld_s, int h, int w, void* dst, long long ld_d, int H, int W, int B,
                       int C, int mode, int accumulate, void* stream) {
    VK_REQUIRE(src && dst, VKOCR_BAD_ARGUMENT, "upsample_fwd: null argument");
    VK_REQUIRE(mode == 0 || mode == 1, VKOCR_BAD_ARGUMENT, "upsample_fwd: mode %d", mode);
    VK_REQUIRE(h > 0 && w > 0, VKOCR_BAD_SHAPE, "upsample_fwd: empty source");
    if ((long long)B * H * W * C == 0) return VKOCR_OK;
    const bool vec = vec_ok(dtype, C, ld_s, ld_d) && ptr16(src, dst);   // else: scalar path (odd widths / slice offsets)
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (vec && mode == 0 && H == 2 * h && W == 2 * w) {
        const int V = dtype == VKOCR_F32 ? 4 : 8;
        const long long total = (long long)B * ((h + UP_ROWS - 1) / UP_ROWS) * w * (C / V);
        VK_DISPATCH_DTYPE(dtype, T, (upsample2x_fwd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                        reinterpret_cast<const T*>(src), ld_s, h, w, reinterpret_cast<T*>(dst), ld_d, B, C / V,
                                        accumulate)));
        VK_CHECK_LAUNCH("upsample2x_fwd_kernel");
        return VKOCR_OK;
    }
    if (vec && mode == 1 && H == 2 * h && W == 2 * w) {
        const int V = dtype == VKOCR_F32 ? 4 : 8;
        const long long total = (long long)B * h * w * (C / V);
        VK_DISPATCH_DTYPE(dtype, T, (nearest2x_fwd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                        reinterpret_cast<const T*>(src), ld_s, h, w, reinterpret_cast<T*>(dst), ld_d, B, C / V,
                                        accumulate)));
        VK_CHECK_LAUNCH("nearest2x_fwd_kernel");
        return VKOCR_OK;
    }
    VK_DISPATCH_DTYPE(dtype, T, (vec ? launch_upsample_fwd<T, VkVec<T>>(src, ld_s, h, w, dst, ld_d, H, W, B, C, mode, accumulate, s)
                                     : launch_upsample_fwd<T, VkScalar<T>>(src, ld_s, h, w, dst, ld_d, H, W, B, C, mode, accumulate, s)));
    VK_CHECK_LAUNCH("upsample_fwd_kernel");
    return VKOCR_OK;
}

// dsrc[b,i,j,0:C] (=|+=) adjoint of vkocr_upsample_fwd applied to ddst
int vkocr_upsample_bwd(int dtype, const void* ddst, long long ld_d, int H, int W, void* dsrc, long long ld_s, int h, int w, int B,
                       int C, int mode, int accumulate, void* stream) {
    VK_REQUIRE(ddst && dsrc, VKOCR_BAD_ARGUMENT, "upsample_bwd: null argument");
    VK_REQUIRE(mode == 0 || mode == 1, VKOCR_BAD_ARGUMENT, "upsample_bwd: mode %d", mode);
    if ((long long)B * h * w * C == 0) return VKOCR_OK;
    const bool vec = vec_ok(dtype, C, ld_s, ld_d) && ptr16(ddst, dsrc);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (vec && mode == 0 && H == 2 * h && W == 2 * w) {
        const int V = dtype == VKOCR_F32 ? 4 : 8;
        const long long total = (long long)B * ((h + UP_ROWS - 1) / UP_ROWS) * w * (C / V);
        VK_DISPATCH_DTYPE(dtype, T, (upsample2x_bwd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                        reinterpret_cast<const T*>(ddst), ld_d, reinterpret_cast<T*>(dsrc), ld_s, h, w, B, C / V,
                                        accumulate)));
        VK_CHECK_LAUNCH("upsample2x_bwd_kernel");
        return VKOCR_OK;
    }
    if (vec && mode == 1 && H == 2 * h && W == 2 * w) {
        const int V = dtype == VKOCR_F32 ? 4 : 8;
        const long long total = (long long)B * h * w * (C / V);
        VK_DISPATCH_DTYPE(dtype, T, (nearest2x_bwd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                        reinterpret_cast<const T*>(ddst), ld_d, reinterpret_cast<T*>(dsrc), ld_s, h, w, B, C / V,
                                        accumulate)));
        VK_CHECK_LAUNCH("nearest2x_bwd_kernel");
        return VKOCR_OK;
    }
    VK_DISPATCH_DTYPE(dtype, T, (vec ? launch_upsample_bwd<T, VkVec<T>>(ddst, ld_d, H, W, dsrc, ld_s, h, w, B, C, mode, accumulate, s)
                                     : launch_upsample_bwd<T, VkScalar<T>>(ddst, ld_d, H, W, dsrc, ld_s, h, w, B, C, mode, accumulate, s)));
    VK_CHECK_LAUNCH("upsample_bwd_kernel");
    return VKOCR_OK;
}

// Separable two-pass form of vkocr_upsample_bwd for large scale factors; workspace: fp32 [B, H, w, C] (caller-owned).
// Requires vector-aligned channels / strides.

int vkocr_upsample_bwd_separable(int dtype, const void* ddst, long long ld_d, int H, int W, void* dsrc, long long ld_s, int h, int w,
                                 int B, int C, int mode, int accumulate, float* workspace, void* stream) {
    VK_REQUIRE(ddst && dsrc && workspace, VKOCR_BAD_ARGUMENT, "upsample_bwd_separable: null argument");
    VK_REQUIRE(mode == 0 || mode == 1, VKOCR_BAD_ARGUMENT, "upsample_bwd_separable: mode %d", mode);
    VK_REQUIRE(vec_ok(dtype, C, ld_s, ld_d) && ptr16(ddst, dsrc) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, VKOCR_BAD_ALIGN,
               "upsample_bwd_separable: C %d / strides / pointers not vector aligned", C);
    if ((long long)B * h * w * C == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const long long t1 = (long long)B * H * w * (C / V), t2 = (long long)B * h * w * (C / V);
    VK_DISPATCH_DTYPE(dtype, T, (upsample_bwd_x_kernel<T><<<(unsigned)((t1 + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(ddst), ld_d, H, W, workspace, w, B, C / V, mode)));
    VK_CHECK_LAUNCH("upsample_bwd_x_kernel");
    VK_DISPATCH_DTYPE(dtype, T, (upsample_bwd_y_kernel<T><<<(unsigned)((t2 + 255) / 256), 256, 0, s>>>(
                                    workspace, H, reinterpret_cast<T*>(dsrc), ld_s, h, w, B, C / V, mode, accumulate)));
    VK_CHECK_LAUNCH("upsample_bwd_y_kernel");
    return VKOCR_OK;
}

int vkocr_avgpool_fwd(int dtype, const void* x, long long ld_x, int H, int W, void* y, long long ld_y, int S, int B, int C,
                      void* stream) {
    VK_REQUIRE(x && y && S >= 1, VKOCR_BAD_ARGUMENT, "avgpool_fwd: bad argument");
    VK_REQUIRE(vec_ok(dtype, C, ld_x, ld_y), VKOCR_BAD_ALIGN, "avgpool_fwd: C %d / strides not vector aligned", C);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const long long total = (long long)B * S * S * (C / V);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (avgpool_fwd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, H, W, reinterpret_cast<T*>(y), ld_y, S, B, C)));
    VK_CHECK_LAUNCH("avgpool_fwd_kernel");
    return VKOCR_OK;
}

// The same pooling in two separable passes through a caller-owned fp32 workspace [B, H, S, C] (large bins).
int vkocr_avgpool_fwd_separable(int dtype, const void* x, long long ld_x, int H, int W, void* y, long long ld_y, int S, int B, int C,
                                float* workspace, void* stream) {
    VK_REQUIRE(x && y && workspace && S >= 1, VKOCR_BAD_ARGUMENT, "avgpool_fwd_separable: bad argument");
    VK_REQUIRE(vec_ok(dtype, C, ld_x, ld_y) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, VKOCR_BAD_ALIGN,
               "avgpool_fwd_separable: C %d / strides not vector aligned", C);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const long long t1 = (long long)B * H * S * (C / V), t2 = (long long)B * S * S * (C / V);
    if (t2 == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (avgpool_rows_kernel<T><<<(unsigned)((t1 + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, H, W, workspace, S, B, C / V)));
    VK_CHECK_LAUNCH("avgpool_rows_kernel");
    VK_DISPATCH_DTYPE(dtype, T, (avgpool_cols_kernel<T><<<(unsigned)((t2 + 255) / 256), 256, 0, s>>>(
                                    workspace, H, W, reinterpret_cast<T*>(y), ld_y, S, B, C / V)));
    VK_CHECK_LAUNCH("avgpool_cols_kernel");
    return VKOCR_OK;
}

int vkocr_avgpool_bwd(int dtype, const void* dy, long long ld_y, int S, void* dx, long long ld_x, int H, int W, int B, int C,
                      int accumulate, void* stream) {
    VK_REQUIRE(dy && dx && S >= 1, VKOCR_BAD_ARGUMENT, "avgpool_bwd: bad argument");
    VK_REQUIRE(vec_ok(dtype, C, ld_x, ld_y), VKOCR_BAD_ALIGN, "avgpool_bwd: C %d / strides not vector aligned", C);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const long long total = (long long)B * H * W * (C / V);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (avgpool_bwd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(dy), ld_y, S, reinterpret_cast<T*>(dx), ld_x, H, W, B, C, accumulate)));
    VK_CHECK_LAUNCH("avgpool_bwd_kernel");
    return VKOCR_OK;
}

// img fp32 NCHW (B,Cin,H,W) -> out [B*(H/p)*(W/p), c_pad] storage dtype
int vkocr_patchify_image(int dtype, const float* img, int B, int Cin, int H, int W, int p, void* out, int c_pad, void* stream) {
    VK_REQUIRE(img && out && p >= 1, VKOCR_BAD_ARGUMENT, "patchify_image: bad argument");
    VK_REQUIRE(c_pad >= p * p * Cin, VKOCR_BAD_SHAPE, "patchify_image: c_pad %d < %d", c_pad, p * p * Cin);
    const long long total = (long long)B * (H / p) * (W / p) * c_pad;
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == VKOCR_BF16 && Cin == 3 && p == 4 && W % 4 == 0 && c_pad % 4 == 0 && c_pad >= 48 &&
        (reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
        const long long threads = (long long)B * (H / 4) * (W / 4) * 4;
        patchify_rgb4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(img, B, H, W, reinterpret_cast<__nv_bfloat16*>(out), c_pad);
        VK_CHECK_LAUNCH("patchify_rgb4_kernel");
        return VKOCR_OK;
    }
    VK_DISPATCH_DTYPE(dtype, T, (patchify_image_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    img, B, Cin, H, W, p, reinterpret_cast<T*>(out), c_pad)));
    VK_CHECK_LAUNCH("patchify_image_kernel");
    return VKOCR_OK;
}

// dir 0: y[B,H/2,W/2,4C] = gather(x[B,H,W,C]);  dir 1: x (=|+=) scatter(y)
int vkocr_space_to_depth2(int dtype, void* x, long long ld_x, int B, int H, int W, int C, void* y, long long ld_y, int dir,
                          int accumulate, void* stream) {
    VK_REQUIRE(x && y, VKOCR_BAD_ARGUMENT, "space_to_depth2: null argument");
    VK_REQUIRE(vec_ok(dtype, C, ld_x, ld_y), VKOCR_BAD_ALIGN, "space_to_depth2: C %d / strides not vector aligned", C);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const long long total = (long long)B * H * W * (C / V);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (space_to_depth2_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<T*>(x), ld_x, B, H, W, C, reinterpret_cast<T*>(y), ld_y, dir, accumulate)));
    VK_CHECK_LAUNCH("space_to_depth2_kernel");
    return VKOCR_OK;
}

int vkocr_copy_channels(int dtype, const void* src, long long ld_s, void* dst, long long ld_d, long long rows, int C, int accumulate,
                        void* stream) {
    VK_REQUIRE(src && dst, VKOCR_BAD_ARGUMENT, "copy_channels: null argument");
    if (rows == 0 || C == 0) return VKOCR_OK;
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    const int esz = dtype == VKOCR_F32 ? 4 : 2;
    const int vec = vec_ok(dtype, C, ld_s, ld_d) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
    (void)esz;
    const long long total = rows * (vec ? C / V : C);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (copy_channels_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(src), ld_s, reinterpret_cast<T*>(dst), ld_d, rows, C, accumulate, vec)));
    VK_CHECK_LAUNCH("copy_channels_kernel");
    return VKOCR_OK;
}

}  // extern "C"
