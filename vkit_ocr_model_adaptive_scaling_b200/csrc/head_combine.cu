// "Convolve first, resample after": the x`factor` up-sampling in front of the heads' k x k convolution commutes with the
// convolution's channel mixing, so the contraction runs on the LOW-resolution neck map and only the (cheap, per-channel)
// interpolation + tap shift runs at the output resolution:
//
//   conv_kxk(up_f(x))[r, s] = bias + sum_{dy,dx} up_f(Z_{dy,dx})[r + dy - k/2, s + dx - k/2],   Z_{dy,dx} = x . W[:, :, dy, dx]^T
//
// (terms whose up-sampled coordinate falls outside the H x W grid are the conv's zero padding).  Z = one plain GEMM
// [B h w, C] x [C, k*k * N] on the tensor cores: factor^2 (4x for the default heads) fewer FLOPs than the convolution on
// the up-sampled map, and the (B, C, f h, f w) up-sampled tensor is never built.
// Reference: UperNextHead.forward upernext.py:233-248 (bilinear), FpnHead.forward fpn.py:193-208 (nearest; 5x5 for factors
// in (2, 4]), nn.Softplus adaptive_scaling.py:101,140.
//
// This file holds the two memory-bound halves around that GEMM:
//   vkocr_head_combine_fwd: Z [B h w, k*k * ntot] -> per head: interpolate + shift + sum + conv bias (fp32) -> LayerNorm
//       -> exact GELU -> Linear(inner -> O <= 4) (-> Softplus) -> NCHW fp32 map; optionally the pre-LayerNorm conv
//       output in storage dtype for the backward.
//   vkocr_head_combine_bwd: d(conv output) [B H W, ntot] -> dZ [B h w, k*k * ntot], the exact adjoint (gather form).
// Each has a generic kernel (any factor, k in {1,3,5}, both modes) and a shared-memory-tiled kernel for the hot case
// factor 2, k 3: the interpolation is separable, so a block first reduces the three tap rows of a (rows x columns) tile of
// Z to per-output-row partial sums V in shared memory (every Z element is loaded and converted once per block), then
// every output pixel combines six V entries.
#include "gemm_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr float LN_EPS = 1e-6f;

struct Axis {
    int i0, i1;
    float w0, w1;
};
// PyTorch area_pixel_compute_source_index(align_corners=False) + clamp (same arithmetic as resample.cu)
__device__ __forceinline__ Axis hc_bilinear_axis(int d, int in, int out) {
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
__device__ __forceinline__ int hc_nearest_src(int d, int in, int out) {
    const float scale = (float)in / (float)out;
    int i = (int)floorf(d * scale);
    return i < in - 1 ? i : in - 1;
}
// source rows / weights of up-sampled coordinate d (valid d only); nearest: one row with weight 1
__device__ __forceinline__ Axis hc_axis(int d, int in, int out, int mode) {
    if (mode == 0) return hc_bilinear_axis(d, in, out);
    Axis a;
    a.i0 = a.i1 = hc_nearest_src(d, in, out);
    a.w0 = 1.f; a.w1 = 0.f;
    return a;
}
// weight of source index p in the interpolation of up-sampled coordinate d
__device__ __forceinline__ float hc_weight_of(int d, int p, int in, int out, int mode) {
    const Axis a = hc_axis(d, in, out, mode);
    return (a.i0 == p ? a.w0 : 0.f) + (a.i1 == p ? a.w1 : 0.f);   // nearest: w1 == 0
}

struct Geom {
    int B, h, w, f, mode, ks, ntot;
    int H, W;
    long long ld_z;
};

struct HeadArgs {            // one head of the group
    int col0, inner, softplus;
    const float* bias;       // conv bias [inner]
    const float* gamma;
    const float* beta;
    const float* w2;         // [O, inner]
    const float* b2;         // [O]
    float* out;              // [B, O, H, W] fp32
};

__device__ __forceinline__ float2 hc_warp_sum2(float a, float b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    return make_float2(a, b);
}

// Per-lane copies of the tail's per-channel parameters (lane owns channels (lane + 32 j) V .. + V of the head).
template <int NVL, int V, int O>
struct TailPar {
    float gm[NVL][V], bt[NVL][V], cb[NVL][V], w[O][NVL][V];
    float bias2;
    __device__ __forceinline__ void load(const HeadArgs& hd, int lane) {
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = (lane + 32 * j) * V + i;
                const bool ok = c < hd.inner;
                gm[j][i] = ok ? __ldg(hd.gamma + c) : 0.f;
                bt[j][i] = ok ? __ldg(hd.beta + c) : 0.f;
                cb[j][i] = ok ? __ldg(hd.bias + c) : 0.f;
#pragma unroll
                for (int o = 0; o < O; ++o) w[o][j][i] = ok ? __ldg(hd.w2 + (long long)o * hd.inner + c) : 0.f;
            }
        bias2 = (lane < O) ? __ldg(hd.b2 + lane) : 0.f;
    }
};

// LayerNorm -> GELU -> projection (-> Softplus) of one pixel whose `inner` conv outputs sit across the warp in c (pad
// channels hold exact zeros), written to out[o * ostride] for o < O.  Statistics: one shuffle round of shifted sums.
template <int NVL, int V, int O>
__device__ __forceinline__ void hc_tail(const float (&c)[NVL][V], const TailPar<NVL, V, O>& tp, int lane, int inner, int softplus,
                                        float* out_px, long long ostride) {
    const float x0 = __shfl_sync(0xffffffffu, c[0][0], 0);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float d = c[j][i] - x0;
            s += d;
            q = fmaf(d, d, q);
        }
    const float2 r = hc_warp_sum2(s, q);
    const int npad = NVL * 32 * V - inner;
    const float inv = 1.f / (float)inner;
    const float ss = r.x + (float)npad * x0;
    const float qq = r.y - (float)npad * x0 * x0;
    const float m = ss * inv;                         // mean - x0
    const float mean = x0 + m;
    const float rstd = rsqrtf(fmaxf(fmaf(-m, m, qq * inv), 0.f) + LN_EPS);
    const float shift = -mean * rstd;
    float dot[O];
#pragma unroll
    for (int o = 0; o < O; ++o) dot[o] = 0.f;
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float hh = fmaf(c[j][i], rstd, shift);
            const float g = vk_gelu(fmaf(hh, tp.gm[j][i], tp.bt[j][i]));      // pad channels: gamma = beta = 0 -> gelu(0) = 0
#pragma unroll
            for (int o = 0; o < O; ++o) dot[o] = fmaf(g, tp.w[o][j][i], dot[o]);
        }
#pragma unroll
    for (int o = 0; o < O; ++o) dot[o] = vk_warp_sum(dot[o]);
    if (lane < O) {
        float v = 0.f;
#pragma unroll
        for (int o = 0; o < O; ++o) v = (lane == o) ? dot[o] : v;
        v += tp.bias2;
        if (softplus) v = vk_softplus(v);
        out_px[(long long)lane * ostride] = v;
    }
}

// ------------------------------------------------------------------------------------------------ generic forward
// One warp per output pixel; every tap gathers its (up to four) source vectors of Z.
template <typename T, int NVL, int O>
__global__ void __launch_bounds__(256)
hc_fwd_generic_kernel(const T* __restrict__ z, Geom g, HeadArgs hd, T* __restrict__ conv, long long ld_conv) {
    constexpr int V = VkVec<T>::N;
    const int lane = threadIdx.x & 31;
    TailPar<NVL, V, O> tp;
    tp.load(hd, lane);
    const long long rows = (long long)g.B * g.H * g.W;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int pad = g.ks >> 1;
    const long long ppi = (long long)g.H * g.W;
    for (long long r = warp0; r < rows; r += nwarps) {
        const int b = (int)(r / ppi);
        const int rem = (int)(r - (long long)b * ppi);
        const int y = rem / g.W, x = rem - y * g.W;
        float c[NVL][V];
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) c[j][i] = tp.cb[j][i];
        const T* zb = z + (long long)b * g.h * g.w * g.ld_z + hd.col0;
        for (int dy = 0; dy < g.ks; ++dy) {
            const int Y = y + dy - pad;
            if (Y < 0 || Y >= g.H) continue;
            const Axis ay = hc_axis(Y, g.h, g.H, g.mode);
            for (int dx = 0; dx < g.ks; ++dx) {
                const int X = x + dx - pad;
                if (X < 0 || X >= g.W) continue;
                const Axis ax = hc_axis(X, g.w, g.W, g.mode);
                const T* zt = zb + (long long)(dy * g.ks + dx) * g.ntot;
                const int np = g.mode == 0 ? 2 : 1;
                for (int py = 0; py < np; ++py) {
                    const int p = py ? ay.i1 : ay.i0;
                    const float wy = py ? ay.w1 : ay.w0;
                    for (int px = 0; px < np; ++px) {
                        const int qq = px ? ax.i1 : ax.i0;
                        const float ww = wy * (px ? ax.w1 : ax.w0);
                        if (ww == 0.f) continue;
                        const T* zp = zt + ((long long)p * g.w + qq) * g.ld_z;
#pragma unroll
                        for (int j = 0; j < NVL; ++j) {
                            const int c0 = (lane + 32 * j) * V;
                            if (c0 < hd.inner) {
                                VkVec<T> v;
                                v.load(zp + c0);
                                float f[V];
                                v.unpack(f);
#pragma unroll
                                for (int i = 0; i < V; ++i) c[j][i] = fmaf(ww, f[i], c[j][i]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i)
                if ((lane + 32 * j) * V + i >= hd.inner) c[j][i] = 0.f;
        if (conv) {
            T* cr = conv + r * ld_conv + hd.col0;
#pragma unroll
            for (int j = 0; j < NVL; ++j) {
                const int c0 = (lane + 32 * j) * V;
                if (c0 < hd.inner) {
                    VkVec<T> vo;
                    vo.pack(c[j]);
                    vo.store(cr + c0);
                }
            }
        }
        hc_tail<NVL, V, O>(c, tp, lane, hd.inner, hd.softplus, hd.out + (long long)b * O * ppi + rem, ppi);
    }
}

// ------------------------------------------------------------------------------------------------ generic adjoint
// One thread per (low-res pixel, tap, channel vector): dZ_tap[p, q] = sum_{Y, X} R[Y, p] C[X, q] dconv[Y - dy + pad, X - dx + pad].
template <typename T>
__global__ void __launch_bounds__(256)
hc_bwd_generic_kernel(const T* __restrict__ dc, long long ld_dc, Geom g, int width, T* __restrict__ dz) {
    constexpr int V = VkVec<T>::N;
    const int CV = width / V;
    const int taps = g.ks * g.ks;
    const long long total = (long long)g.B * g.h * g.w * taps * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int tap = (int)(r % taps);
    r /= taps;
    const int q = (int)(r % g.w);
    r /= g.w;
    const int p = (int)(r % g.h);
    const int b = (int)(r / g.h);
    const int dy = tap / g.ks, dx = tap - dy * g.ks;
    const int pad = g.ks >> 1;
    const float ry = (float)g.H / (float)g.h, rx = (float)g.W / (float)g.w;
    int Ylo = (int)floorf((p - 0.5f) * ry - 0.5f) - 1, Yhi = (int)ceilf((p + 1.5f) * ry - 0.5f) + 1;
    int Xlo = (int)floorf((q - 0.5f) * rx - 0.5f) - 1, Xhi = (int)ceilf((q + 1.5f) * rx - 0.5f) + 1;
    if (Ylo < 0) Ylo = 0;
    if (Xlo < 0) Xlo = 0;
    if (Yhi > g.H - 1) Yhi = g.H - 1;
    if (Xhi > g.W - 1) Xhi = g.W - 1;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* db = dc + (long long)b * g.H * g.W * ld_dc + cv * V;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
        const int yy = Y - dy + pad;
        if (yy < 0 || yy >= g.H) continue;
        const float wy = hc_weight_of(Y, p, g.h, g.H, g.mode);
        if (wy == 0.f) continue;
        for (int X = Xlo; X <= Xhi; ++X) {
            const int xx = X - dx + pad;
            if (xx < 0 || xx >= g.W) continue;
            const float wx = hc_weight_of(X, q, g.w, g.W, g.mode);
            if (wx == 0.f) continue;
            VkVec<T> v;
            v.load(db + ((long long)yy * g.W + xx) * ld_dc);
            float f[V];
            v.unpack(f);
            const float ww = wy * wx;
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] = fmaf(ww, f[e], acc[e]);
        }
    }
    VkVec<T> o;
    o.pack(acc);
    o.store(dz + (((long long)b * g.h + p) * g.w + q) * g.ld_z + (long long)tap * g.ntot + cv * V);
}

// ------------------------------------------------------------------------------------------------ factor 2, 3x3: forward
// Up-sampled coordinate Y = 2 i + a + dy - 1 (a = output-row parity) interpolates two source rows at FIXED offsets from i:
//   a = 0: dy 0 -> (i-1, i), dy 1 -> (i-1, i), dy 2 -> (i, i+1);   a = 1: dy 0 -> (i-1, i), dy 1 -> (i, i+1), dy 2 -> (i, i+1)
// with the rows clamped into [0, h) (the clamp of align_corners=False interpolation folds onto the same two rows) and the
// pair's weights taken from the interpolation of Y; a Y outside [0, 2h) is conv padding: weights 0.  Same for columns.
// The kernels are bound by instruction issue, so all per-element arithmetic runs on packed fp32 pairs (FFMA2), the tails of
// the two output pixels of a (row, low-res column) run together (shared V loads, interleaved dependency chains), and the
// warp reductions keep only half of the values per butterfly round.
constexpr int HC_TJ = 8;            // low-res columns per tile
constexpr int HC_QL = HC_TJ + 2;    // columns of V kept per tile (one halo column each side)
__device__ __forceinline__ constexpr int hc_pos(int a, int d, int k) {   // offset of the k-th source row of (parity a, tap d)
    // a=0: (-1,0) (-1,0) (0,1);  a=1: (-1,0) (0,1) (0,1)
    return ((a == 0) ? (d == 2 ? 0 : -1) : (d == 0 ? -1 : 0)) + k;
}
// weights of the two (clamped) source rows of up-sampled coordinate Y = 2 i + a + d - 1
__device__ __forceinline__ float2 hc_pair_weights(int i, int a, int d, int in, int mode) {
    const int out = 2 * in;
    const int Y = 2 * i + a + d - 1;
    if (Y < 0 || Y >= out) return make_float2(0.f, 0.f);
    const Axis ax = hc_axis(Y, in, out, mode);
    int rA = i + hc_pos(a, d, 0), rB = i + hc_pos(a, d, 1);
    rA = rA < 0 ? 0 : (rA > in - 1 ? in - 1 : rA);
    rB = rB < 0 ? 0 : (rB > in - 1 ? in - 1 : rB);
    const float wA = (ax.i0 == rA ? ax.w0 : 0.f) + (ax.i1 == rA ? ax.w1 : 0.f);
    const float wB = (rB != rA) ? ((ax.i0 == rB ? ax.w0 : 0.f) + (ax.i1 == rB ? ax.w1 : 0.f)) : 0.f;
    return make_float2(wA, wB);
}

// Shared-memory channel order of one V / E entry: 16-byte groups are interleaved across the channel vectors
// ([element / 4][vector][element % 4]) so that a warp's float4 accesses are contiguous (no bank conflicts) both when one
// thread per vector writes and when one lane per vector reads.
__device__ __forceinline__ int hc_plane(int cv, int e, int nvec) { return (e >> 2) * (nvec * 4) + cv * 4; }

// 16 bytes of storage -> V/2 fp32 pairs
template <typename T> struct HcPairs;
template <> struct HcPairs<__nv_bfloat16> {
    static constexpr int NP = 4;
    static __device__ __forceinline__ void unpack(const uint4& raw, float2 (&p)[4]) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
    }
    static __device__ __forceinline__ uint4 pack(const float2 (&p)[4]) {
        uint4 raw;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(p[i].x, p[i].y);
        return raw;
    }
};
template <> struct HcPairs<float> {
    static constexpr int NP = 2;
    static __device__ __forceinline__ void unpack(const uint4& raw, float2 (&p)[2]) {
        p[0] = make_float2(__uint_as_float(raw.x), __uint_as_float(raw.y));
        p[1] = make_float2(__uint_as_float(raw.z), __uint_as_float(raw.w));
    }
    static __device__ __forceinline__ uint4 pack(const float2 (&p)[2]) {
        return make_uint4(__float_as_uint(p[0].x), __float_as_uint(p[0].y), __float_as_uint(p[1].x), __float_as_uint(p[1].y));
    }
};
__device__ __forceinline__ uint4 hc_ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Forward-only exact GELU of a pair (vk_gelu with the polynomial packed; one MUFU.EX2 per element)
__device__ __forceinline__ float2 hc_gelu2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 a = make_float2(fminf(ax.x, 9.f), fminf(ax.y, 9.f));
    float2 q = vk_fma2(a, vk_splat2(-3.159132926264e-05f), vk_splat2(7.531542511152e-04f));
    q = vk_fma2(q, a, vk_splat2(-8.015595779370e-03f));
    q = vk_fma2(q, a, vk_splat2(5.328616913828e-02f));
    q = vk_fma2(q, a, vk_splat2(4.588905654064e-01f));
    q = vk_fma2(q, a, vk_splat2(1.151150930355e+00f));
    q = vk_fma2(q, a, vk_splat2(1.f));
    const float2 e = make_float2(vk_ex2(-q.x), vk_ex2(-q.y));
    return vk_fma2(make_float2(-ax.x, -ax.y), e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

// Sum N (power of two, <= 8) per-lane values over the warp, keeping half of the values per butterfly round while more
// than one is left: N - 1 + (5 - log2 N) shuffles instead of 5 N.  On return v[0] of lane L is the total of value
// L >> (5 - log2 N).
template <int N>
__device__ __forceinline__ void hc_reduce_multi(float (&v)[N], int lane) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8, "hc_reduce_multi: N");
    int mask = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = upper ? v[i + n / 2] : v[i];
            const float send = upper ? v[i] : v[i + n / 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
        mask >>= 1;
    }
#pragma unroll
    for (; mask > 0; mask >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
}

template <int O> struct HcPow2 { static constexpr int value = O <= 1 ? 1 : (O <= 2 ? 2 : 4); };

// Per-lane pairs of the tail's per-channel parameters (lane owns channels (lane + 32 j) V .. + V of the head).
// Heads with three or four output maps keep the projection weights in shared memory ([o][vector][pair][lane] float2, a
// conflict-free LDS.64 per use): in registers they are 8 O values per lane next to gamma / beta and the two pixels'
// channels, and the O = 4 instances spilled.
template <int O> struct HcW2Smem { static constexpr bool value = O >= 3; };
template <int NVL, int NP, int O>
struct TailPar2 {
    static constexpr int OR = HcW2Smem<O>::value ? 1 : O;
    float2 gm[NVL][NP], bt[NVL][NP], w[OR][NVL][NP];
    float bias2;      // b2[o] of the (pixel, o) this lane writes after the projection reduce
    __device__ __forceinline__ void load(const HeadArgs& hd, int lane) {
        auto ldc = [&](const float* p, int c) { return c < hd.inner ? __ldg(p + c) : 0.f; };
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const int c = (lane + 32 * j) * NP * 2 + 2 * i;
                gm[j][i] = make_float2(ldc(hd.gamma, c), ldc(hd.gamma, c + 1));
                bt[j][i] = make_float2(ldc(hd.beta, c), ldc(hd.beta, c + 1));
                if constexpr (!HcW2Smem<O>::value) {
#pragma unroll
                    for (int o = 0; o < O; ++o)
                        w[o][j][i] = make_float2(ldc(hd.w2 + (long long)o * hd.inner, c), ldc(hd.w2 + (long long)o * hd.inner, c + 1));
                }
            }
        constexpr int OP = HcPow2<O>::value;
        constexpr int NV = 2 * OP;                    // values of the projection reduce: (pixel, o)
        const int idx = lane / (32 / NV);
        const int o = idx % OP;
        bias2 = o < O ? __ldg(hd.b2 + o) : 0.f;
    }
};

// LayerNorm -> GELU -> projection (-> Softplus) of TWO pixels whose `inner` conv outputs sit across the warp in c0 / c1
// (pad channels hold exact zeros); results to out0[o * ostride] / out1[o * ostride].
template <int NVL, int NP, int O>
__device__ __forceinline__ void hc_tail2(const float2 (&c0)[NVL][NP], const float2 (&c1)[NVL][NP], const TailPar2<NVL, NP, O>& tp,
                                         int lane, int inner, int softplus, float* out0, float* out1, long long ostride,
                                         const float2* __restrict__ sW2) {
    const float x0 = __shfl_sync(0xffffffffu, c0[0][0].x, 0);
    const float x1 = __shfl_sync(0xffffffffu, c1[0][0].x, 0);
    const float2 n0 = vk_splat2(-x0), n1 = vk_splat2(-x1);
    float2 s0 = make_float2(0.f, 0.f), q0 = s0, s1 = s0, q1 = s0;
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const float2 d0 = vk_add2(c0[j][i], n0), d1 = vk_add2(c1[j][i], n1);
            s0 = vk_add2(s0, d0);
            q0 = vk_fma2(d0, d0, q0);
            s1 = vk_add2(s1, d1);
            q1 = vk_fma2(d1, d1, q1);
        }
    float st[4] = {s0.x + s0.y, q0.x + q0.y, s1.x + s1.y, q1.x + q1.y};
    hc_reduce_multi<4>(st, lane);
    const float sA = __shfl_sync(0xffffffffu, st[0], 0), qA = __shfl_sync(0xffffffffu, st[0], 8);
    const float sB = __shfl_sync(0xffffffffu, st[0], 16), qB = __shfl_sync(0xffffffffu, st[0], 24);
    const int npad = NVL * 32 * NP * 2 - inner;
    const float inv = 1.f / (float)inner;
    const float mA = (sA + (float)npad * x0) * inv, mB = (sB + (float)npad * x1) * inv;          // mean - x0
    const float vA = fmaf(-mA, mA, (qA - (float)npad * x0 * x0) * inv), vB = fmaf(-mB, mB, (qB - (float)npad * x1 * x1) * inv);
    const float rA = rsqrtf(fmaxf(vA, 0.f) + LN_EPS), rB = rsqrtf(fmaxf(vB, 0.f) + LN_EPS);
    const float2 rs0 = vk_splat2(rA), sh0 = vk_splat2(-(x0 + mA) * rA), rs1 = vk_splat2(rB), sh1 = vk_splat2(-(x1 + mB) * rB);
    constexpr int OP = HcPow2<O>::value;
    float2 d0[O], d1[O];
#pragma unroll
    for (int o = 0; o < O; ++o) d0[o] = d1[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            // pad channels: gamma = beta = 0 -> gelu(0) = 0
            const float2 g0 = hc_gelu2(vk_fma2(vk_fma2(c0[j][i], rs0, sh0), tp.gm[j][i], tp.bt[j][i]));
            const float2 g1 = hc_gelu2(vk_fma2(vk_fma2(c1[j][i], rs1, sh1), tp.gm[j][i], tp.bt[j][i]));
#pragma unroll
            for (int o = 0; o < O; ++o) {
                float2 wv;
                if constexpr (HcW2Smem<O>::value) wv = sW2[((o * NVL + j) * NP + i) * 32 + lane];
                else wv = tp.w[o][j][i];
                d0[o] = vk_fma2(g0, wv, d0[o]);
                d1[o] = vk_fma2(g1, wv, d1[o]);
            }
        }
    float dv[2 * OP];
#pragma unroll
    for (int o = 0; o < OP; ++o) {
        dv[o] = o < O ? d0[o < O ? o : 0].x + d0[o < O ? o : 0].y : 0.f;
        dv[OP + o] = o < O ? d1[o < O ? o : 0].x + d1[o < O ? o : 0].y : 0.f;
    }
    hc_reduce_multi<2 * OP>(dv, lane);
    constexpr int GROUP = 32 / (2 * OP);
    if ((lane & (GROUP - 1)) == 0) {
        const int idx = lane / GROUP;
        const int o = idx % OP;
        if (o < O) {
            float v = dv[0] + tp.bias2;
            if (softplus) v = vk_softplus(v);
            ((idx / OP) ? out1 : out0)[(long long)o * ostride] = v;
        }
    }
}

// Channel order of a V entry: interleaved planes (hc_plane) for one-lane-per-16-byte-vector readers, plain channel order for
// the half-warp readers below (a lane reads the float4s l16, l16 + 16, l16 + 32 of a 192-channel entry).
template <bool LIN>
__device__ __forceinline__ int hc_voff(int cv, int e, int nvec) { return LIN ? cv * 8 + (e >> 2) * 4 : hc_plane(cv, e, nvec); }

// The tail for heads with inner = 192 and one output map (both rough heads, the precise mask head) with a HALF-warp per
// (row, low-res column): 192 channels are 24 16-byte vectors, so with a lane per vector a quarter of the warp idles and the
// per-pixel reductions serve two pixels; here a lane owns 12 channels (float4s l16 + 16 k of the entry) of the unit's two
// pixels, every lane works, and one round of shuffles serves the four pixels of the warp's two units.
constexpr int HH_K = 3;      // float4s per lane
struct TailParH {
    float2 gm[HH_K][2], bt[HH_K][2], w[HH_K][2];
    float bias2;
    __device__ __forceinline__ void load(const HeadArgs& hd, int l16) {
#pragma unroll
        for (int k = 0; k < HH_K; ++k)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = 4 * (l16 + 16 * k) + 2 * i;
                gm[k][i] = make_float2(__ldg(hd.gamma + c), __ldg(hd.gamma + c + 1));
                bt[k][i] = make_float2(__ldg(hd.beta + c), __ldg(hd.beta + c + 1));
                w[k][i] = make_float2(__ldg(hd.w2 + c), __ldg(hd.w2 + c + 1));
            }
        bias2 = __ldg(hd.b2);
    }
};
__device__ __forceinline__ void hc_tail2_half(const float2 (&c0)[HH_K][2], const float2 (&c1)[HH_K][2], const TailParH& tp, int lane,
                                              int softplus, bool valid, float* out0, float* out1) {
    const int base = lane & 16;
    const float x0 = __shfl_sync(0xffffffffu, c0[0][0].x, base);
    const float x1 = __shfl_sync(0xffffffffu, c1[0][0].x, base);
    const float2 n0 = vk_splat2(-x0), n1 = vk_splat2(-x1);
    float2 s0 = make_float2(0.f, 0.f), q0 = s0, s1 = s0, q1 = s0;
#pragma unroll
    for (int k = 0; k < HH_K; ++k)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float2 d0 = vk_add2(c0[k][i], n0), d1 = vk_add2(c1[k][i], n1);
            s0 = vk_add2(s0, d0);
            q0 = vk_fma2(d0, d0, q0);
            s1 = vk_add2(s1, d1);
            q1 = vk_fma2(d1, d1, q1);
        }
    // four values over 16 lanes, half of them kept per round: lane bit 3 selects the pixel, bit 2 sum / sum of squares
    float t;
    {
        const float a0 = s0.x + s0.y, a1 = q0.x + q0.y, a2 = s1.x + s1.y, a3 = q1.x + q1.y;
        const bool up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
        const float v0 = (up8 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, up8 ? a0 : a2, 8);
        const float v1 = (up8 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, up8 ? a1 : a3, 8);
        t = (up4 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, up4 ? v0 : v1, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
    }
    const float sA = __shfl_sync(0xffffffffu, t, base), qA = __shfl_sync(0xffffffffu, t, base + 4);
    const float sB = __shfl_sync(0xffffffffu, t, base + 8), qB = __shfl_sync(0xffffffffu, t, base + 12);
    constexpr float inv = 1.f / 192.f;
    const float mA = sA * inv, mB = sB * inv;                                                     // mean - x0
    const float rA = rsqrtf(fmaxf(fmaf(-mA, mA, qA * inv), 0.f) + LN_EPS), rB = rsqrtf(fmaxf(fmaf(-mB, mB, qB * inv), 0.f) + LN_EPS);
    const float2 rs0 = vk_splat2(rA), sh0 = vk_splat2(-(x0 + mA) * rA), rs1 = vk_splat2(rB), sh1 = vk_splat2(-(x1 + mB) * rB);
    float2 d0 = make_float2(0.f, 0.f), d1 = d0;
#pragma unroll
    for (int k = 0; k < HH_K; ++k)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float2 g0 = hc_gelu2(vk_fma2(vk_fma2(c0[k][i], rs0, sh0), tp.gm[k][i], tp.bt[k][i]));
            const float2 g1 = hc_gelu2(vk_fma2(vk_fma2(c1[k][i], rs1, sh1), tp.gm[k][i], tp.bt[k][i]));
            d0 = vk_fma2(g0, tp.w[k][i], d0);
            d1 = vk_fma2(g1, tp.w[k][i], d1);
        }
    const bool up8 = (lane & 8) != 0;
    const float p0 = d0.x + d0.y, p1 = d1.x + d1.y;
    float v = (up8 ? p1 : p0) + __shfl_xor_sync(0xffffffffu, up8 ? p0 : p1, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (valid && (lane & 7) == 0) {
        v += tp.bias2;
        if (softplus) v = vk_softplus(v);
        *(up8 ? out1 : out0) = v;
    }
}

template <typename T, int NVL, int O, int HC_R, int THREADS, int MINB, bool HALF = false>
__global__ void __launch_bounds__(THREADS, MINB)
hc_fwd_2x3_kernel(const T* __restrict__ z, Geom g, HeadArgs hd, T* __restrict__ conv, long long ld_conv, int cw, int tiles_j,
                  int tiles_i, long long ntiles) {
    constexpr int V = VkVec<T>::N;
    constexpr int NP = HcPairs<T>::NP;
    extern __shared__ float4 hc_smem4[];
    float* sV = reinterpret_cast<float*>(hc_smem4);                           // [R][2][3][QL][cw]
    float* sBias = sV + HC_R * 2 * 3 * HC_QL * cw;                            // [cw] conv bias in plane order
    float2* sWr = reinterpret_cast<float2*>(sBias + cw);                      // [R][2][3][2] row weights, splatted
    float2* sWc = sWr + HC_R * 2 * 3 * 2;                                     // [TJ][2][3] column-pair weights (wA, wB)
    float2* sW2 = sWc + HC_TJ * 2 * 3;                                        // [O][NVL][NP][32] projection weights (O >= 3 only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cvn = cw / V;                            // channel vectors that hold real channels
    TailPar2<HALF ? 1 : NVL, HALF ? 1 : NP, HALF ? 1 : O> tp;      // (unused with HALF)
    TailParH tph;
    if constexpr (HALF) tph.load(hd, lane & 15);
    else tp.load(hd, lane);
    for (int c = threadIdx.x; c < cw; c += blockDim.x)
        sBias[hc_voff<HALF>(c / V, c % V, cvn) + (c & 3)] = c < hd.inner ? __ldg(hd.bias + c) : 0.f;
    if constexpr (!HALF && HcW2Smem<O>::value) {
        for (int idx = threadIdx.x; idx < O * NVL * NP * 32; idx += blockDim.x) {
            const int ln = idx & 31, i = (idx >> 5) % NP, j = (idx / (32 * NP)) % NVL, o = idx / (32 * NP * NVL);
            const int c = (ln + 32 * j) * NP * 2 + 2 * i;
            const float* wrow = hd.w2 + (long long)o * hd.inner;
            sW2[idx] = make_float2(c < hd.inner ? __ldg(wrow + c) : 0.f, c + 1 < hd.inner ? __ldg(wrow + c + 1) : 0.f);
        }
    }
    const long long ppi = (long long)g.H * g.W;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int jt = (int)(tile % tiles_j);
        const long long t2 = tile / tiles_j;
        const int it = (int)(t2 % tiles_i);
        const int b = (int)(t2 / tiles_i);
        const int i0 = it * HC_R, j0 = jt * HC_TJ;
        if (threadIdx.x < HC_R * 2 * 3) {
            const int d = threadIdx.x % 3, a = (threadIdx.x / 3) & 1, li = threadIdx.x / 6;
            const float2 wv = (i0 + li < g.h) ? hc_pair_weights(i0 + li, a, d, g.h, g.mode) : make_float2(0.f, 0.f);
            sWr[2 * threadIdx.x] = vk_splat2(wv.x);
            sWr[2 * threadIdx.x + 1] = vk_splat2(wv.y);
        } else if (threadIdx.x >= 32 && threadIdx.x < 32 + HC_TJ * 2 * 3) {
            const int t = threadIdx.x - 32;
            const int d = t % 3, b2 = (t / 3) & 1, jl = t / 6;
            sWc[t] = (j0 + jl < g.w) ? hc_pair_weights(j0 + jl, b2, d, g.w, g.mode) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        // ---------------- phase 1: V[li][a][dx][ql][c] = sum_dy (wA Z_{dy,dx}[rowA, q, c] + wB Z_{dy,dx}[rowB, q, c])
        {
            long long rowoff[HC_R + 2];
#pragma unroll
            for (int k = 0; k < HC_R + 2; ++k) {
                int rr = i0 - 1 + k;
                rr = rr < 0 ? 0 : (rr > g.h - 1 ? g.h - 1 : rr);
                rowoff[k] = (long long)rr * g.w * g.ld_z;
            }
            const T* zb = z + (long long)b * g.h * g.w * g.ld_z + hd.col0;
            const int items = HC_QL * 3 * cvn;
            for (int item = threadIdx.x; item < items; item += blockDim.x) {
                const int cv = item % cvn;
                const int t3 = item / cvn;
                const int dx = t3 % 3;
                const int ql = t3 / 3;
                int q = j0 - 1 + ql;
                q = q < 0 ? 0 : (q > g.w - 1 ? g.w - 1 : q);
                const T* zq = zb + (long long)q * g.ld_z + dx * g.ntot + cv * V;
                // all loads first (10 of the 12 (tap row, source row) combinations are read)
                uint4 raw[3][HC_R + 2];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int k = 0; k < HC_R + 2; ++k) {
                        bool used = false;
#pragma unroll
                        for (int li = 0; li < HC_R; ++li)
#pragma unroll
                            for (int a = 0; a < 2; ++a)
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk)
                                    if (hc_pos(a, d, kk) == k - 1 - li) used = true;
                        if (used) raw[d][k] = hc_ldg16(zq + rowoff[k] + (long long)(d * 3) * g.ntot);
                    }
                float2 acc[HC_R][2][NP];
#pragma unroll
                for (int li = 0; li < HC_R; ++li)
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int e = 0; e < NP; ++e) acc[li][a][e] = make_float2(0.f, 0.f);
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int k = 0; k < HC_R + 2; ++k) {
                        bool used = false;
#pragma unroll
                        for (int li = 0; li < HC_R; ++li)
#pragma unroll
                            for (int a = 0; a < 2; ++a)
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk)
                                    if (hc_pos(a, d, kk) == k - 1 - li) used = true;
                        if (!used) continue;
                        float2 f[NP];
                        HcPairs<T>::unpack(raw[d][k], f);
#pragma unroll
                        for (int li = 0; li < HC_R; ++li)
#pragma unroll
                            for (int a = 0; a < 2; ++a)
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk)
                                    if (hc_pos(a, d, kk) == k - 1 - li) {
                                        const float2 wgt = sWr[((li * 2 + a) * 3 + d) * 2 + kk];
#pragma unroll
                                        for (int e = 0; e < NP; ++e) acc[li][a][e] = vk_fma2(wgt, f[e], acc[li][a][e]);
                                    }
                    }
#pragma unroll
                for (int li = 0; li < HC_R; ++li)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        float* dst = sV + ((((li * 2 + a) * 3 + dx) * HC_QL + ql)) * cw;
#pragma unroll
                        for (int e = 0; e < NP; e += 2)
                            *reinterpret_cast<float4*>(dst + hc_voff<HALF>(cv, 2 * e, cvn)) =
                                make_float4(acc[li][a][e].x, acc[li][a][e].y, acc[li][a][e + 1].x, acc[li][a][e + 1].y);
                    }
            }
        }
        __syncthreads();
        if constexpr (HALF) {
            // ---------------- phase 2, half-warp flavour: a warp takes two neighbouring low-res columns of one output row
            const int l16 = lane & 15;
            for (int u2 = warp; u2 < HC_R * HC_TJ; u2 += (blockDim.x >> 5)) {
                const int unit = 2 * u2 + (lane >> 4);
                const int jl = unit % HC_TJ;
                const int la = unit / HC_TJ;               // li * 2 + a
                const int i = i0 + (la >> 1), j = j0 + jl;
                if (i >= g.h || j0 + (jl & ~1) >= g.w) continue;       // warp-uniform: the row or both columns are outside
                const bool valid = j < g.w;
                int qm = j - 1, qp = j + 1;
                qm = qm < 0 ? 0 : qm;
                qp = qp > g.w - 1 ? g.w - 1 : qp;
                int col[3] = {qm - (j0 - 1), jl + 1, qp - (j0 - 1)};
                if (!valid) col[0] = col[1] = col[2] = 0;
                float2 c0[HH_K][2], c1[HH_K][2];
#pragma unroll
                for (int k = 0; k < HH_K; ++k) {
                    const float4 bv = *reinterpret_cast<const float4*>(sBias + 4 * (l16 + 16 * k));
                    c0[k][0] = c1[k][0] = make_float2(bv.x, bv.y);
                    c0[k][1] = c1[k][1] = make_float2(bv.z, bv.w);
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float2 w0 = sWc[(jl * 2 + 0) * 3 + d], w1 = sWc[(jl * 2 + 1) * 3 + d];
                    const float* base = sV + ((la * 3 + d) * HC_QL) * cw;
#pragma unroll
                    for (int pp = 0; pp < 3; ++pp) {
                        float u0 = 0.f, u1 = 0.f;
                        bool r0 = false, r1 = false;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            if (hc_pos(0, d, kk) == pp - 1) { r0 = true; u0 = kk ? w0.y : w0.x; }
                            if (hc_pos(1, d, kk) == pp - 1) { r1 = true; u1 = kk ? w1.y : w1.x; }
                        }
                        if (!r0 && !r1) continue;
                        const float2 u02 = vk_splat2(u0), u12 = vk_splat2(u1);
                        const float* src = base + col[pp] * cw + 4 * l16;
#pragma unroll
                        for (int k = 0; k < HH_K; ++k) {
                            const float4 v = *reinterpret_cast<const float4*>(src + 64 * k);
                            const float2 va = make_float2(v.x, v.y), vb = make_float2(v.z, v.w);
                            if (r0) { c0[k][0] = vk_fma2(u02, va, c0[k][0]); c0[k][1] = vk_fma2(u02, vb, c0[k][1]); }
                            if (r1) { c1[k][0] = vk_fma2(u12, va, c1[k][0]); c1[k][1] = vk_fma2(u12, vb, c1[k][1]); }
                        }
                    }
                }
                const int y = 2 * i + (la & 1), x = 2 * j;
                const long long rem = (long long)y * g.W + x;
                if (conv && valid) {
                    T* cr = conv + ((long long)b * ppi + rem) * ld_conv + hd.col0 + 4 * l16;
#pragma unroll
                    for (int k = 0; k < HH_K; ++k) {
                        uint2 p0, p1;
                        *reinterpret_cast<__nv_bfloat162*>(&p0.x) = __floats2bfloat162_rn(c0[k][0].x, c0[k][0].y);
                        *reinterpret_cast<__nv_bfloat162*>(&p0.y) = __floats2bfloat162_rn(c0[k][1].x, c0[k][1].y);
                        *reinterpret_cast<__nv_bfloat162*>(&p1.x) = __floats2bfloat162_rn(c1[k][0].x, c1[k][0].y);
                        *reinterpret_cast<__nv_bfloat162*>(&p1.y) = __floats2bfloat162_rn(c1[k][1].x, c1[k][1].y);
                        __stcs(reinterpret_cast<uint2*>(cr + 64 * k), p0);
                        __stcs(reinterpret_cast<uint2*>(cr + ld_conv + 64 * k), p1);
                    }
                }
                float* o0 = hd.out + (long long)b * ppi + (valid ? rem : 0);
                hc_tail2_half(c0, c1, tph, lane, hd.softplus, valid, o0, o0 + 1);
            }
        } else
        // ---------------- phase 2: one warp per (row, low-res column): its two output pixels (column parity 0 / 1) together
        for (int unit = warp; unit < HC_R * 2 * HC_TJ; unit += (blockDim.x >> 5)) {
            const int jl = unit % HC_TJ;
            const int la = unit / HC_TJ;               // li * 2 + a
            const int i = i0 + (la >> 1), j = j0 + jl;
            if (i >= g.h || j >= g.w) continue;
            int qm = j - 1, qp = j + 1;
            qm = qm < 0 ? 0 : qm;
            qp = qp > g.w - 1 ? g.w - 1 : qp;
            const int col[3] = {qm - (j0 - 1), jl + 1, qp - (j0 - 1)};
            float2 c0[NVL][NP], c1[NVL][NP];
#pragma unroll
            for (int jj = 0; jj < NVL; ++jj) {
                const int cvi = lane + 32 * jj;
#pragma unroll
                for (int e = 0; e < NP; e += 2) {
                    const float4 bv = cvi < cvn ? *reinterpret_cast<const float4*>(sBias + hc_plane(cvi, 2 * e, cvn)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    c0[jj][e] = c1[jj][e] = make_float2(bv.x, bv.y);
                    c0[jj][e + 1] = c1[jj][e + 1] = make_float2(bv.z, bv.w);
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float2 w0 = sWc[(jl * 2 + 0) * 3 + d], w1 = sWc[(jl * 2 + 1) * 3 + d];
                const float* base = sV + ((la * 3 + d) * HC_QL) * cw;
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) {       // column offsets -1, 0, +1
                    // which of the two pixels read this column for tap d?  (compile-time)
                    float u0 = 0.f, u1 = 0.f;
                    bool r0 = false, r1 = false;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        if (hc_pos(0, d, kk) == pp - 1) { r0 = true; u0 = kk ? w0.y : w0.x; }
                        if (hc_pos(1, d, kk) == pp - 1) { r1 = true; u1 = kk ? w1.y : w1.x; }
                    }
                    if (!r0 && !r1) continue;
                    const float2 u02 = vk_splat2(u0), u12 = vk_splat2(u1);
                    const float* src = base + col[pp] * cw;
#pragma unroll
                    for (int jj = 0; jj < NVL; ++jj) {
                        const int cvi = lane + 32 * jj;
                        if (cvi < cvn) {
#pragma unroll
                            for (int e = 0; e < NP; e += 2) {
                                const float4 v = *reinterpret_cast<const float4*>(src + hc_plane(cvi, 2 * e, cvn));
                                const float2 va = make_float2(v.x, v.y), vb = make_float2(v.z, v.w);
                                if (r0) { c0[jj][e] = vk_fma2(u02, va, c0[jj][e]); c0[jj][e + 1] = vk_fma2(u02, vb, c0[jj][e + 1]); }
                                if (r1) { c1[jj][e] = vk_fma2(u12, va, c1[jj][e]); c1[jj][e + 1] = vk_fma2(u12, vb, c1[jj][e + 1]); }
                            }
                        }
                    }
                }
            }
            // pad channels of the last vector: the weights' pad rows are zero, so Z and the bias are zero there already
            const int y = 2 * i + (la & 1), x = 2 * j;
            const long long rem = (long long)y * g.W + x;
            if (conv) {
                T* cr = conv + ((long long)b * ppi + rem) * ld_conv + hd.col0;
#pragma unroll
                for (int jj = 0; jj < NVL; ++jj) {
                    const int ch0 = (lane + 32 * jj) * V;
                    if (ch0 < hd.inner) {
                        __stcs(reinterpret_cast<uint4*>(cr + ch0), HcPairs<T>::pack(c0[jj]));
                        __stcs(reinterpret_cast<uint4*>(cr + ld_conv + ch0), HcPairs<T>::pack(c1[jj]));
                    }
                }
            }
            float* o0 = hd.out + (long long)b * O * ppi + rem;
            hc_tail2<NVL, NP, O>(c0, c1, tp, lane, hd.inner, hd.softplus, o0, o0 + 1, ppi, sW2);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ factor 2, 3x3: adjoint
// dZ_{dy,dx}[p, q] = sum_{o, o'} R[2p-1+o, p] C[2q-1+o', q] dconv[2p - dy + o, 2q - dx + o'],  o, o' in 0..3 (the four
// up-sampled rows / columns whose interpolation reads source p / q), terms outside the grids dropped.  Separable: a block
// first reduces the rows of a dconv tile to E[pl][dy][column] in shared memory (one thread per (column, channel vector),
// its eight rows prefetched into registers while the previous tile is finished), then one thread per (pixel, channel
// vector) combines six E columns into the nine taps.
constexpr int HB_TQ = 8;                 // low-res columns per tile  -> 2 (TQ + 2) = 20 hi-res columns of dconv
constexpr int HB_R = 2;                  // low-res rows per tile     -> 2 (R + 2)  =  8 hi-res rows of dconv
constexpr int HB_SL = 2 * (HB_TQ + 2);   // 20
constexpr int HB_ROWS = 2 * (HB_R + 2);  // 8
constexpr int HB_CVL = 16;               // channel vectors per chunk
constexpr int HB_THREADS = HB_SL * HB_CVL;   // 320: phase 1 uses all, phase 2 the first R * TQ * CVL = 256

template <typename T>
__global__ void __launch_bounds__(HB_THREADS, 2)
hc_bwd_2x3_kernel(const T* __restrict__ dc, long long ld_dc, Geom g, int width, T* __restrict__ dz, int tiles_q, int tiles_p,
                  int chunks, long long nwork, int pmajor) {
    constexpr int V = VkVec<T>::N;
    constexpr int NP = HcPairs<T>::NP;
    constexpr int CH = HB_CVL * V;
    extern __shared__ float4 hc_smem4[];
    float* sE = reinterpret_cast<float*>(hc_smem4);            // [R][3][SL][CH]
    float* cE = sE + HB_R * 3 * HB_SL * CH;                     // [R][3][4] row coefficients
    float* cC = cE + HB_R * 3 * 4;                              // [TQ][3][4] column coefficients
    const int sl = threadIdx.x / HB_CVL, cvl = threadIdx.x % HB_CVL;
    struct Work { int b, p0, q0, c0; };
    auto decode = [&](long long work) {
        Work wk;
        const int ch = (int)(work % chunks);
        long long t2 = work / chunks;
        int qt, pt;
        if (pmajor) {           // tiles of one column back to back: the blocks of a wave share their halo rows while they are hot in L2
            pt = (int)(t2 % tiles_p);
            t2 /= tiles_p;
            qt = (int)(t2 % tiles_q);
            wk.b = (int)(t2 / tiles_q);
        } else {
            qt = (int)(t2 % tiles_q);
            t2 /= tiles_q;
            pt = (int)(t2 % tiles_p);
            wk.b = (int)(t2 / tiles_p);
        }
        wk.p0 = pt * HB_R; wk.q0 = qt * HB_TQ; wk.c0 = ch * CH;
        return wk;
    };
    uint4 raw[HB_ROWS];
    auto prefetch = [&](const Work& wk) {
        const int s = 2 * (wk.q0 - 1) + sl;
        const int c = wk.c0 + cvl * V;
        const bool ok = s >= 0 && s < g.W && c < width;
        const T* db = dc + ((long long)wk.b * g.H * g.W + (ok ? s : 0)) * ld_dc + (ok ? c : 0);
#pragma unroll
        for (int rr = 0; rr < HB_ROWS; ++rr) {
            const int r = 2 * (wk.p0 - 1) + rr;
            raw[rr] = (ok && r >= 0 && r < g.H) ? hc_ldg16(db + (long long)r * g.W * ld_dc) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    long long work = blockIdx.x;
    Work wk{0, 0, 0, 0};
    if (work < nwork) {
        wk = decode(work);
        prefetch(wk);
    }
    for (; work < nwork; work += gridDim.x) {
        if (threadIdx.x < HB_R * 3 * 4) {
            const int o = threadIdx.x & 3, dy = (threadIdx.x >> 2) % 3, pl = threadIdx.x / 12;
            const int p = wk.p0 + pl, Y = 2 * p - 1 + o, yy = Y - dy + 1;
            float cf = 0.f;
            if (p < g.h && Y >= 0 && Y < g.H && yy >= 0 && yy < g.H) cf = hc_weight_of(Y, p, g.h, g.H, g.mode);
            cE[threadIdx.x] = cf;
        } else if (threadIdx.x >= 32 && threadIdx.x < 32 + HB_TQ * 3 * 4) {
            const int t = threadIdx.x - 32;
            const int o = t & 3, dx = (t >> 2) % 3, ql = t / 12;
            const int q = wk.q0 + ql, X = 2 * q - 1 + o, xx = X - dx + 1;
            float cf = 0.f;
            if (q < g.w && X >= 0 && X < g.W && xx >= 0 && xx < g.W) cf = hc_weight_of(X, q, g.w, g.W, g.mode);
            cC[t] = cf;
        }
        __syncthreads();
        // ---------------- phase 1: E[pl][dy][sl][c] = sum_o cE[pl][dy][o] dconv[2 p - dy + o, s, c]
        {
            float2 acc[HB_R][3][NP];
#pragma unroll
            for (int pl = 0; pl < HB_R; ++pl)
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int e = 0; e < NP; ++e) acc[pl][d][e] = make_float2(0.f, 0.f);
#pragma unroll
            for (int rr = 0; rr < HB_ROWS; ++rr) {
                float2 f[NP];
                HcPairs<T>::unpack(raw[rr], f);
#pragma unroll
                for (int pl = 0; pl < HB_R; ++pl)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int o = rr - 2 * pl - (2 - d);    // r = 2 p - d + o
                        if (o >= 0 && o < 4) {
                            const float2 cf = vk_splat2(cE[(pl * 3 + d) * 4 + o]);
#pragma unroll
                            for (int e = 0; e < NP; ++e) acc[pl][d][e] = vk_fma2(cf, f[e], acc[pl][d][e]);
                        }
                    }
            }
#pragma unroll
            for (int pl = 0; pl < HB_R; ++pl)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float* dst = sE + ((pl * 3 + d) * HB_SL + sl) * CH;
#pragma unroll
                    for (int e = 0; e < NP; e += 2)
                        *reinterpret_cast<float4*>(dst + hc_plane(cvl, 2 * e, HB_CVL)) =
                            make_float4(acc[pl][d][e].x, acc[pl][d][e].y, acc[pl][d][e + 1].x, acc[pl][d][e + 1].y);
                }
        }
        __syncthreads();
        // the next tile's rows travel while this tile's taps are combined and stored
        const Work cur = wk;
        if (work + gridDim.x < nwork) {
            wk = decode(work + gridDim.x);
            prefetch(wk);
        }
        // ---------------- phase 2: dZ_{dy,dx}[p, q, c] = sum_o cC[ql][dx][o] E[pl][dy][2 ql + 2 - dx + o][c]
        if (threadIdx.x < HB_R * HB_TQ * HB_CVL) {
            const int pix = threadIdx.x / HB_CVL;
            const int pl = pix / HB_TQ, ql = pix % HB_TQ;
            const int p = cur.p0 + pl, q = cur.q0 + ql, c = cur.c0 + cvl * V;
            if (p < g.h && q < g.w && c < width) {
                float2 cf[3][4];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int o = 0; o < 4; ++o) cf[dx][o] = vk_splat2(cC[(ql * 3 + dx) * 4 + o]);
                T* dst = dz + (((long long)cur.b * g.h + p) * g.w + q) * g.ld_z + c;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    float2 ev[6][NP];
                    const float* src = sE + ((pl * 3 + dy) * HB_SL + 2 * ql) * CH;
#pragma unroll
                    for (int k = 0; k < 6; ++k)
#pragma unroll
                        for (int e = 0; e < NP; e += 2) {
                            const float4 v = *reinterpret_cast<const float4*>(src + k * CH + hc_plane(cvl, 2 * e, HB_CVL));
                            ev[k][e] = make_float2(v.x, v.y);
                            ev[k][e + 1] = make_float2(v.z, v.w);
                        }
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        float2 acc[NP];
#pragma unroll
                        for (int e = 0; e < NP; ++e) acc[e] = vk_mul2(cf[dx][0], ev[2 - dx][e]);
#pragma unroll
                        for (int o = 1; o < 4; ++o)
#pragma unroll
                            for (int e = 0; e < NP; ++e) acc[e] = vk_fma2(cf[dx][o], ev[2 - dx + o][e], acc[e]);
                        __stcs(reinterpret_cast<uint4*>(dst + (long long)(dy * 3 + dx) * g.ntot), HcPairs<T>::pack(acc));   // write-once stream: keep L2 for the dconv halo rows
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ launchers
template <typename T, int NVL, int O>
int hc_launch_fwd(const void* z, const Geom& g, const HeadArgs& hd, void* conv, long long ld_conv, int algo, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    const long long rows = (long long)g.B * g.H * g.W;
    const bool fast = (algo != 1) && g.f == 2 && g.ks == 3;
    if (fast) {
        const int cvn = (hd.inner + V - 1) / V;
        const int cw = cvn * V;
        int variant = 0;     // 0: 2 rows x 256 threads x 2 blocks/SM; 1: 1 row x 128 threads x 4 blocks/SM; 2: 2 rows x 384 threads x 2 blocks/SM
        if (const char* e = getenv("VKOCR_HC_VARIANT")) variant = atoi(e);
        auto run = [&](auto rtag, auto ttag, auto btag) -> int {
            constexpr int R = decltype(rtag)::value, THREADS = decltype(ttag)::value, MINB = decltype(btag)::value;
            const size_t smem = ((size_t)R * 2 * 3 * HC_QL * cw + cw) * sizeof(float) + (R * 2 * 3 * 2 + HC_TJ * 2 * 3) * sizeof(float2) +
                                (HcW2Smem<O>::value ? (size_t)O * NVL * HcPairs<T>::NP * 32 * sizeof(float2) : 0);
            if (smem > 200 * 1024) return -1;
            const int tiles_j = vk_cdiv(g.w, HC_TJ), tiles_i = vk_cdiv(g.h, R);
            const long long ntiles = (long long)g.B * tiles_i * tiles_j;
            auto kern = hc_fwd_2x3_kernel<T, NVL, O, R, THREADS, MINB>;
            if constexpr (std::is_same<T, __nv_bfloat16>::value && NVL == 1 && O == 1) {
                const bool no_half = getenv("VKOCR_HC_NOHALF") != nullptr;      // A/B switch (tests, tools/profile_combine.py)
                if (hd.inner == 192 && !no_half) kern = hc_fwd_2x3_kernel<T, NVL, O, R, THREADS, MINB, true>;   // half-warp tails
            }
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return 2;
            int per_sm = (int)((220 * 1024) / (smem + 1024));
            if (per_sm > MINB) per_sm = MINB;
            if (per_sm < 1) per_sm = 1;
            long long blocks = (long long)vkocr_sm_count() * per_sm;
            if (blocks > ntiles) blocks = ntiles;
            kern<<<(unsigned)blocks, THREADS, smem, s>>>(reinterpret_cast<const T*>(z), g, hd, reinterpret_cast<T*>(conv), ld_conv, cw, tiles_j,
                                                         tiles_i, ntiles);
            return 0;
        };
        using std::integral_constant;
        int rc;
        if (variant == 1) rc = run(integral_constant<int, 1>{}, integral_constant<int, 128>{}, integral_constant<int, 4>{});
        else if (variant == 2) rc = run(integral_constant<int, 2>{}, integral_constant<int, 384>{}, integral_constant<int, 2>{});
        else rc = run(integral_constant<int, 2>{}, integral_constant<int, 256>{}, integral_constant<int, 2>{});
        if (rc >= 0) return rc;
    }
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    hc_fwd_generic_kernel<T, NVL, O><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const T*>(z), g, hd, reinterpret_cast<T*>(conv), ld_conv);
    return 0;
}

template <typename T, int NVL>
int hc_dispatch_o(int O, const void* z, const Geom& g, const HeadArgs& hd, void* conv, long long ld_conv, int algo, cudaStream_t s) {
    switch (O) {
        case 1: return hc_launch_fwd<T, NVL, 1>(z, g, hd, conv, ld_conv, algo, s);
        case 2: return hc_launch_fwd<T, NVL, 2>(z, g, hd, conv, ld_conv, algo, s);
        case 3: return hc_launch_fwd<T, NVL, 3>(z, g, hd, conv, ld_conv, algo, s);
        case 4: return hc_launch_fwd<T, NVL, 4>(z, g, hd, conv, ld_conv, algo, s);
        default: return 1;
    }
}

}  // namespace

extern "C" {

// z: [B*h*w, ld_z] storage dtype, columns tap * ntot + head * slot + n (tap = dy * ks + dx); conv_bias: [ntot] fp32 in the
// same head * slot + n order.  conv_out (nullable): [B*H*W, ld_conv] storage dtype, columns head * slot + n.
// algo: 0 = pick, 1 = force the generic kernel (tests).
int vkocr_head_combine_fwd(int dtype, const void* z, long long ld_z, int B, int h, int w, int factor, int mode, int ks, int ntot,
                           const float* conv_bias, const VkocrHeadTail* heads, void* conv_out, long long ld_conv, int algo,
                           void* stream) {
    VK_REQUIRE(z && conv_bias && heads, VKOCR_BAD_ARGUMENT, "head_combine_fwd: null argument");
    VK_REQUIRE(dtype == VKOCR_F32 || dtype == VKOCR_BF16, VKOCR_UNSUPPORTED_DTYPE, "head_combine_fwd: dtype %d", dtype);
    VK_REQUIRE(factor >= 1 && factor <= 8 && (mode == 0 || mode == 1) && (ks == 1 || ks == 3 || ks == 5), VKOCR_BAD_SHAPE,
               "head_combine_fwd: factor %d mode %d kernel %d", factor, mode, ks);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(heads->num_heads >= 1 && heads->num_heads <= VKOCR_MAX_HEADS && heads->slot % 8 == 0 && heads->slot <= 256 &&
                   ntot == heads->num_heads * heads->slot,
               VKOCR_BAD_SHAPE, "head_combine_fwd: %d heads x slot %d vs ntot %d", heads->num_heads, heads->slot, ntot);
    VK_REQUIRE(ld_z % V == 0 && ld_z >= (long long)ks * ks * ntot && (reinterpret_cast<uintptr_t>(z) & 15) == 0, VKOCR_BAD_ALIGN,
               "head_combine_fwd: z stride %lld", ld_z);
    VK_REQUIRE(!conv_out || (ld_conv % V == 0 && ld_conv >= ntot && (reinterpret_cast<uintptr_t>(conv_out) & 15) == 0), VKOCR_BAD_ALIGN,
               "head_combine_fwd: conv stride %lld", ld_conv);
    VK_REQUIRE((long long)B * h * factor * w * factor < (1LL << 31), VKOCR_BAD_SHAPE, "head_combine_fwd: too many output pixels");
    if ((long long)B * h * w == 0) return VKOCR_OK;
    Geom g;
    g.B = B; g.h = h; g.w = w; g.f = factor; g.mode = mode; g.ks = ks; g.ntot = ntot;
    g.H = h * factor; g.W = w * factor; g.ld_z = ld_z;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    for (int hi = 0; hi < heads->num_heads; ++hi) {
        VK_REQUIRE(heads->gamma[hi] && heads->beta[hi] && heads->w2[hi] && heads->b2[hi] && heads->out[hi] && heads->inner[hi] >= 1 &&
                       heads->inner[hi] <= heads->slot && heads->out_channels[hi] >= 1 && heads->out_channels[hi] <= 4,
                   VKOCR_BAD_ARGUMENT, "head_combine_fwd: head %d parameters", hi);
        HeadArgs hd;
        hd.col0 = hi * heads->slot;
        hd.inner = heads->inner[hi];
        hd.softplus = heads->softplus[hi];
        hd.bias = conv_bias + hd.col0;
        hd.gamma = heads->gamma[hi]; hd.beta = heads->beta[hi]; hd.w2 = heads->w2[hi]; hd.b2 = heads->b2[hi];
        hd.out = heads->out[hi];
        const int O = heads->out_channels[hi];
        const int nvl = (heads->slot / V + 31) / 32;
        int rc = 1;
        if (dtype == VKOCR_BF16) {
            if (nvl == 1) rc = hc_dispatch_o<__nv_bfloat16, 1>(O, z, g, hd, conv_out, ld_conv, algo, s);
        } else {
            if (nvl == 1) rc = hc_dispatch_o<float, 1>(O, z, g, hd, conv_out, ld_conv, algo, s);
            else if (nvl == 2) rc = hc_dispatch_o<float, 2>(O, z, g, hd, conv_out, ld_conv, algo, s);
        }
        VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_combine_fwd: dispatch failed (%d) for slot %d, O %d", rc, heads->slot, O);
        VK_CHECK_LAUNCH("hc_fwd_kernel");
    }
    return VKOCR_OK;
}

// dconv: [B*H*W, ld_dc] storage dtype, `width` channels (a multiple of the 16-byte vector; pad channels must hold zeros or
// finite values -- they are carried through).  dz: [B*h*w, ld_z], written for every tap and all `width` channels.
int vkocr_head_combine_bwd(int dtype, const void* dconv, long long ld_dc, int B, int h, int w, int factor, int mode, int ks,
                           int width, void* dz, long long ld_z, int algo, void* stream) {
    VK_REQUIRE(dconv && dz, VKOCR_BAD_ARGUMENT, "head_combine_bwd: null argument");
    VK_REQUIRE(dtype == VKOCR_F32 || dtype == VKOCR_BF16, VKOCR_UNSUPPORTED_DTYPE, "head_combine_bwd: dtype %d", dtype);
    VK_REQUIRE(factor >= 1 && factor <= 8 && (mode == 0 || mode == 1) && (ks == 1 || ks == 3 || ks == 5), VKOCR_BAD_SHAPE,
               "head_combine_bwd: factor %d mode %d kernel %d", factor, mode, ks);
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(width % V == 0 && ld_dc % V == 0 && ld_z % V == 0 && ld_dc >= width && ld_z >= (long long)ks * ks * width &&
                   (reinterpret_cast<uintptr_t>(dconv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0,
               VKOCR_BAD_ALIGN, "head_combine_bwd: width %d strides %lld %lld", width, ld_dc, ld_z);
    if ((long long)B * h * w == 0) return VKOCR_OK;
    Geom g;
    g.B = B; g.h = h; g.w = w; g.f = factor; g.mode = mode; g.ks = ks; g.ntot = width;
    g.H = h * factor; g.W = w * factor; g.ld_z = ld_z;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (algo != 1 && factor == 2 && ks == 3) {
        const int CH = HB_CVL * V;
        const int tiles_q = vk_cdiv(w, HB_TQ), tiles_p = vk_cdiv(h, HB_R), chunks = vk_cdiv(width, CH);
        const long long nwork = (long long)B * tiles_p * tiles_q * chunks;
        const size_t smem = ((size_t)HB_R * 3 * HB_SL * CH + HB_R * 3 * 4 + HB_TQ * 3 * 4) * sizeof(float);
        long long blocks = (long long)vkocr_sm_count() * 2;
        if (blocks > nwork) blocks = nwork;
        int pmajor = 1;                                            // VKOCR_HCB_PMAJOR=0: the row-major tile order (A/B, tools/profile_combine.py)
        if (const char* e = getenv("VKOCR_HCB_PMAJOR")) pmajor = atoi(e);
#define VK_HC_BWD(T)                                                                                                             \
    do {                                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(hc_bwd_2x3_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        VK_REQUIRE(e == cudaSuccess, VKOCR_CUDA_ERROR, "head_combine_bwd: smem %zu: %s", smem, cudaGetErrorString(e));           \
        hc_bwd_2x3_kernel<T><<<(unsigned)blocks, HB_THREADS, smem, s>>>(reinterpret_cast<const T*>(dconv), ld_dc, g, width,             \
                                                                 reinterpret_cast<T*>(dz), tiles_q, tiles_p, chunks, nwork,      \
                                                                 pmajor);                                                       \
    } while (0)
        if (dtype == VKOCR_BF16) VK_HC_BWD(__nv_bfloat16);
        else VK_HC_BWD(float);
#undef VK_HC_BWD
        VK_CHECK_LAUNCH("hc_bwd_2x3_kernel");
        return VKOCR_OK;
    }
    const long long total = (long long)B * h * w * ks * ks * (width / V);
    VK_REQUIRE((total + 255) / 256 < (1LL << 31), VKOCR_BAD_SHAPE, "head_combine_bwd: too many elements");
    VK_DISPATCH_DTYPE(dtype, T, (hc_bwd_generic_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(dconv), ld_dc, g, width, reinterpret_cast<T*>(dz))));
    VK_CHECK_LAUNCH("hc_bwd_generic_kernel");
    return VKOCR_OK;
}

}  // extern "C"
