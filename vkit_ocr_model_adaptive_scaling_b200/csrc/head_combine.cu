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
constexpr int HC_TJ = 8;            // low-res columns per tile
constexpr int HC_R = 2;             // low-res rows per tile
constexpr int HC_QL = HC_TJ + 2;    // columns of V kept per tile (one halo column each side)
__device__ __forceinline__ int hc_pos(int a, int d, int k) {   // offset of the k-th source row of (parity a, tap d)
    // a=0: (-1,0) (-1,0) (0,1);  a=1: (-1,0) (0,1) (0,1)
    const int first = (a == 0) ? (d == 2 ? 0 : -1) : (d == 0 ? -1 : 0);
    return first + k;
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

template <typename T, int NVL, int O>
__global__ void __launch_bounds__(256, 2)
hc_fwd_2x3_kernel(const T* __restrict__ z, Geom g, HeadArgs hd, T* __restrict__ conv, long long ld_conv, int cw, int tiles_j,
                  int tiles_i, long long ntiles) {
    constexpr int V = VkVec<T>::N;
    extern __shared__ float4 hc_smem4[];
    float* sV = reinterpret_cast<float*>(hc_smem4);                          // [R][2][3][QL][cw]
    float2* sWr = reinterpret_cast<float2*>(sV + HC_R * 2 * 3 * HC_QL * cw);   // [R][2][3] row-pair weights
    float2* sWc = sWr + HC_R * 2 * 3;                                         // [TJ][2][3] column-pair weights
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TailPar<NVL, V, O> tp;
    tp.load(hd, lane);
    const int cvn = cw / V;                            // channel vectors that hold real channels
    const long long ppi = (long long)g.H * g.W;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int jt = (int)(tile % tiles_j);
        const long long t2 = tile / tiles_j;
        const int it = (int)(t2 % tiles_i);
        const int b = (int)(t2 / tiles_i);
        const int i0 = it * HC_R, j0 = jt * HC_TJ;
        if (threadIdx.x < HC_R * 2 * 3) {
            const int d = threadIdx.x % 3, a = (threadIdx.x / 3) & 1, li = threadIdx.x / 6;
            sWr[threadIdx.x] = (i0 + li < g.h) ? hc_pair_weights(i0 + li, a, d, g.h, g.mode) : make_float2(0.f, 0.f);
        } else if (threadIdx.x >= 32 && threadIdx.x < 32 + HC_TJ * 2 * 3) {
            const int t = threadIdx.x - 32;
            const int d = t % 3, b2 = (t / 3) & 1, jl = t / 6;
            sWc[t] = (j0 + jl < g.w) ? hc_pair_weights(j0 + jl, b2, d, g.w, g.mode) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        // ---------------- phase 1: V[li][a][dx][ql][c] = sum_dy (wA Z_{dy,dx}[rowA, q, c] + wB Z_{dy,dx}[rowB, q, c])
        int rk[HC_R + 2];
#pragma unroll
        for (int k = 0; k < HC_R + 2; ++k) {
            const int rr = i0 - 1 + k;
            rk[k] = rr < 0 ? 0 : (rr > g.h - 1 ? g.h - 1 : rr);
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
            float acc[HC_R][2][V];
#pragma unroll
            for (int li = 0; li < HC_R; ++li)
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[li][a][e] = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const T* zt = zb + (long long)(d * 3 + dx) * g.ntot + cv * V;
#pragma unroll
                for (int k = 0; k < HC_R + 2; ++k) {
                    // which accumulators read source row k (offset k - 1 - li from output row li)?  (compile-time)
                    bool used = false;
#pragma unroll
                    for (int li = 0; li < HC_R; ++li)
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                if (hc_pos(a, d, kk) == k - 1 - li) used = true;
                    if (!used) continue;
                    VkVec<T> v;
                    v.load(zt + ((long long)rk[k] * g.w + q) * g.ld_z);
                    float f[V];
                    v.unpack(f);
#pragma unroll
                    for (int li = 0; li < HC_R; ++li)
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                if (hc_pos(a, d, kk) == k - 1 - li) {
                                    const float2 w2 = sWr[(li * 2 + a) * 3 + d];
                                    const float wgt = kk ? w2.y : w2.x;
#pragma unroll
                                    for (int e = 0; e < V; ++e) acc[li][a][e] = fmaf(wgt, f[e], acc[li][a][e]);
                                }
                }
            }
#pragma unroll
            for (int li = 0; li < HC_R; ++li)
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    float* dst = sV + ((long long)(((li * 2 + a) * 3 + dx) * HC_QL + ql)) * cw;
#pragma unroll
                    for (int e = 0; e < V; e += 4)
                        *reinterpret_cast<float4*>(dst + hc_plane(cv, e, cvn)) =
                            make_float4(acc[li][a][e], acc[li][a][e + 1], acc[li][a][e + 2], acc[li][a][e + 3]);
                }
        }
        __syncthreads();
        // ---------------- phase 2: one warp per output pixel: combine the three tap columns, then the head tail
        for (int pi = warp; pi < HC_R * 2 * HC_TJ * 2; pi += (blockDim.x >> 5)) {
            const int b2 = pi & 1;
            const int jl = (pi >> 1) % HC_TJ;
            const int la = (pi >> 1) / HC_TJ;          // li * 2 + a
            const int li = la >> 1, a = la & 1;
            const int i = i0 + li, j = j0 + jl;
            if (i >= g.h || j >= g.w) continue;
            float c[NVL][V];
#pragma unroll
            for (int jj = 0; jj < NVL; ++jj)
#pragma unroll
                for (int e = 0; e < V; ++e) c[jj][e] = tp.cb[jj][e];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float2 wc = sWc[(jl * 2 + b2) * 3 + d];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const float wgt = kk ? wc.y : wc.x;
                    int q = j + hc_pos(b2, d, kk);
                    q = q < 0 ? 0 : (q > g.w - 1 ? g.w - 1 : q);
                    const int ql = q - (j0 - 1);
                    const float* src = sV + ((long long)((la * 3 + d) * HC_QL + ql)) * cw;
#pragma unroll
                    for (int jj = 0; jj < NVL; ++jj) {
                        const int cvi = lane + 32 * jj;
                        if (cvi < cvn) {
#pragma unroll
                            for (int e = 0; e < V; e += 4) {
                                const float4 v = *reinterpret_cast<const float4*>(src + hc_plane(cvi, e, cvn));
                                c[jj][e] = fmaf(wgt, v.x, c[jj][e]);
                                c[jj][e + 1] = fmaf(wgt, v.y, c[jj][e + 1]);
                                c[jj][e + 2] = fmaf(wgt, v.z, c[jj][e + 2]);
                                c[jj][e + 3] = fmaf(wgt, v.w, c[jj][e + 3]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < NVL; ++jj)
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if ((lane + 32 * jj) * V + e >= hd.inner) c[jj][e] = 0.f;
            const int y = 2 * i + a, x = 2 * j + b2;
            const long long rem = (long long)y * g.W + x;
            if (conv) {
                T* cr = conv + ((long long)b * ppi + rem) * ld_conv + hd.col0;
#pragma unroll
                for (int jj = 0; jj < NVL; ++jj) {
                    const int c0 = (lane + 32 * jj) * V;
                    if (c0 < hd.inner) {
                        VkVec<T> vo;
                        vo.pack(c[jj]);
                        vo.store(cr + c0);
                    }
                }
            }
            hc_tail<NVL, V, O>(c, tp, lane, hd.inner, hd.softplus, hd.out + (long long)b * O * ppi + rem, ppi);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ factor 2, 3x3: adjoint
// dZ_{dy,dx}[p, q] = sum_{o, o'} R[2p-1+o, p] C[2q-1+o', q] dconv[2p - dy + o, 2q - dx + o'],  o, o' in 0..3 (the four
// up-sampled rows / columns whose interpolation reads source p / q), terms outside the grids dropped.  Separable: a block
// first reduces the rows of a dconv tile to E[pl][dy][column] in shared memory, then every (pixel, tap) combines four E's.
constexpr int HB_TQ = 6;                 // low-res columns per tile  -> 2 (TQ + 2) = 16 hi-res columns of dconv
constexpr int HB_R = 2;                  // low-res rows per tile     -> 2 (R + 2)  =  8 hi-res rows of dconv
constexpr int HB_SL = 2 * (HB_TQ + 2);   // 16
constexpr int HB_CVL = 16;               // channel vectors per chunk

template <typename T>
__global__ void __launch_bounds__(256, 2)
hc_bwd_2x3_kernel(const T* __restrict__ dc, long long ld_dc, Geom g, int width, T* __restrict__ dz, int tiles_q, int tiles_p,
                  int chunks, long long nwork) {
    constexpr int V = VkVec<T>::N;
    constexpr int CH = HB_CVL * V;
    extern __shared__ float4 hc_smem4[];
    float* sE = reinterpret_cast<float*>(hc_smem4);            // [R][3][SL][CH]
    float* cE = sE + HB_R * 3 * HB_SL * CH;                     // [R][3][4] row coefficients
    float* cC = cE + HB_R * 3 * 4;                              // [TQ][3][4] column coefficients
    for (long long work = blockIdx.x; work < nwork; work += gridDim.x) {
        const int ch = (int)(work % chunks);
        long long t2 = work / chunks;
        const int qt = (int)(t2 % tiles_q);
        t2 /= tiles_q;
        const int pt = (int)(t2 % tiles_p);
        const int b = (int)(t2 / tiles_p);
        const int p0 = pt * HB_R, q0 = qt * HB_TQ, c0 = ch * CH;
        if (threadIdx.x < HB_R * 3 * 4) {
            const int o = threadIdx.x & 3, dy = (threadIdx.x >> 2) % 3, pl = threadIdx.x / 12;
            const int p = p0 + pl, Y = 2 * p - 1 + o, yy = Y - dy + 1;
            float cf = 0.f;
            if (p < g.h && Y >= 0 && Y < g.H && yy >= 0 && yy < g.H) cf = hc_weight_of(Y, p, g.h, g.H, g.mode);
            cE[threadIdx.x] = cf;
        } else if (threadIdx.x >= 32 && threadIdx.x < 32 + HB_TQ * 3 * 4) {
            const int t = threadIdx.x - 32;
            const int o = t & 3, dx = (t >> 2) % 3, ql = t / 12;
            const int q = q0 + ql, X = 2 * q - 1 + o, xx = X - dx + 1;
            float cf = 0.f;
            if (q < g.w && X >= 0 && X < g.W && xx >= 0 && xx < g.W) cf = hc_weight_of(X, q, g.w, g.W, g.mode);
            cC[t] = cf;
        }
        __syncthreads();
        // ---------------- phase 1: E[pl][dy][sl][c] = sum_o cE[pl][dy][o] dconv[2 p - dy + o, s, c]
        {
            const int sl = threadIdx.x / HB_CVL, cvl = threadIdx.x % HB_CVL;
            const int s = 2 * (q0 - 1) + sl;
            const int c = c0 + cvl * V;
            float acc[HB_R][3][V];
#pragma unroll
            for (int pl = 0; pl < HB_R; ++pl)
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[pl][d][e] = 0.f;
            if (s >= 0 && s < g.W && c < width) {
                const T* db = dc + ((long long)b * g.H * g.W + s) * ld_dc + c;
#pragma unroll
                for (int rr = 0; rr < 2 * (HB_R + 2); ++rr) {
                    const int r = 2 * (p0 - 1) + rr;
                    if (r < 0 || r >= g.H) continue;
                    VkVec<T> v;
                    v.load(db + (long long)r * g.W * ld_dc);
                    float f[V];
                    v.unpack(f);
#pragma unroll
                    for (int pl = 0; pl < HB_R; ++pl)
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const int o = rr - 2 * pl - (2 - d);    // r = 2 p - d + o
                            if (o >= 0 && o < 4) {
                                const float cf = cE[(pl * 3 + d) * 4 + o];
#pragma unroll
                                for (int e = 0; e < V; ++e) acc[pl][d][e] = fmaf(cf, f[e], acc[pl][d][e]);
                            }
                        }
                }
            }
#pragma unroll
            for (int pl = 0; pl < HB_R; ++pl)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float* dst = sE + ((long long)((pl * 3 + d) * HB_SL + sl)) * CH;
#pragma unroll
                    for (int e = 0; e < V; e += 4)
                        *reinterpret_cast<float4*>(dst + hc_plane(cvl, e, HB_CVL)) =
                            make_float4(acc[pl][d][e], acc[pl][d][e + 1], acc[pl][d][e + 2], acc[pl][d][e + 3]);
                }
        }
        __syncthreads();
        // ---------------- phase 2: dZ_{dy,dx}[p, q, c] = sum_o cC[ql][dx][o] E[pl][dy][2 ql + 2 - dx + o][c]
        for (int item = threadIdx.x; item < HB_R * HB_TQ * 9 * HB_CVL; item += blockDim.x) {
            const int cvl = item % HB_CVL;
            int t3 = item / HB_CVL;
            const int tap = t3 % 9;
            t3 /= 9;
            const int ql = t3 % HB_TQ;
            const int pl = t3 / HB_TQ;
            const int p = p0 + pl, q = q0 + ql, c = c0 + cvl * V;
            if (p >= g.h || q >= g.w || c >= width) continue;
            const int dy = tap / 3, dx = tap - dy * 3;
            float acc[V];
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float cf = cC[(ql * 3 + dx) * 4 + o];
                const float* src = sE + ((long long)((pl * 3 + dy) * HB_SL + 2 * ql + 2 - dx + o)) * CH;
#pragma unroll
                for (int e = 0; e < V; e += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(src + hc_plane(cvl, e, HB_CVL));
                    acc[e] = fmaf(cf, v.x, acc[e]);
                    acc[e + 1] = fmaf(cf, v.y, acc[e + 1]);
                    acc[e + 2] = fmaf(cf, v.z, acc[e + 2]);
                    acc[e + 3] = fmaf(cf, v.w, acc[e + 3]);
                }
            }
            VkVec<T> ov;
            ov.pack(acc);
            ov.store(dz + (((long long)b * g.h + p) * g.w + q) * g.ld_z + (long long)tap * g.ntot + c);
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
        const size_t smem = (size_t)HC_R * 2 * 3 * HC_QL * cw * sizeof(float) + (HC_R * 2 * 3 + HC_TJ * 2 * 3) * sizeof(float2);
        if (smem <= 200 * 1024) {
            const int tiles_j = vk_cdiv(g.w, HC_TJ), tiles_i = vk_cdiv(g.h, HC_R);
            const long long ntiles = (long long)g.B * tiles_i * tiles_j;
            cudaError_t e = cudaFuncSetAttribute(hc_fwd_2x3_kernel<T, NVL, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return 2;
            const int per_sm = smem <= 100 * 1024 ? 2 : 1;
            long long blocks = (long long)vkocr_sm_count() * per_sm;
            if (blocks > ntiles) blocks = ntiles;
            hc_fwd_2x3_kernel<T, NVL, O><<<(unsigned)blocks, 256, smem, s>>>(reinterpret_cast<const T*>(z), g, hd, reinterpret_cast<T*>(conv),
                                                                             ld_conv, cw, tiles_j, tiles_i, ntiles);
            return 0;
        }
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
    VK_REQUIRE(heads->num_heads >= 1 && heads->num_heads <= VKOCR_MAX_HEADS && heads->slot % 16 == 0 && heads->slot <= 256 &&
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
        long long blocks = (long long)vkocr_sm_count() * 4;
        if (blocks > nwork) blocks = nwork;
#define VK_HC_BWD(T)                                                                                                             \
    do {                                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(hc_bwd_2x3_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        VK_REQUIRE(e == cudaSuccess, VKOCR_CUDA_ERROR, "head_combine_bwd: smem %zu: %s", smem, cudaGetErrorString(e));           \
        hc_bwd_2x3_kernel<T><<<(unsigned)blocks, 256, smem, s>>>(reinterpret_cast<const T*>(dconv), ld_dc, g, width,             \
                                                                 reinterpret_cast<T*>(dz), tiles_q, tiles_p, chunks, nwork);     \
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
