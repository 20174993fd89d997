// Tail of the adaptive-scaling heads, fused: LayerNorm(inner) -> exact GELU -> Linear(inner -> out<=4) (-> Softplus),
// reading the 3x3-conv output slice once and writing the NCHW fp32 prediction map.
// Reference: UperNextHead.forward upernext.py:233-248 / FpnHead.forward fpn.py:193-208 (step1 LN+GELU, step2 1x1),
// nn.Softplus on the height / distance heads (adaptive_scaling.py:101,140).
//
// One warp per pixel row, each lane owns NVL 16-byte channel vectors of the slice; the slice is padded to a multiple of
// the vector width and pad channels are read as zeros.  The kernels are instruction-bound (about 40 fp32 instructions
// per element in the backward, half of them the GELU), so everything per-row that is not per-element is kept off the
// critical path: the per-channel parameters live in registers, the LayerNorm statistics come from ONE shuffle round of
// shifted sums (sum (x - x0), sum (x - x0)^2: no cancellation, no second sweep), the row -> (image, pixel) split uses a
// float reciprocal instead of a 64-bit division, and each warp works on two rows at a time for memory- and
// instruction-level parallelism.
// Backward recomputes LN/GELU from the saved conv output, writes d(conv output) and accumulates all parameter
// gradients (LN gamma/beta, 1x1 weight/bias, conv bias) from per-lane register partials.
#include "common.cuh"

namespace {

constexpr int HT_THREADS = 256;
constexpr int HT_WARPS = HT_THREADS / 32;
constexpr float LN_EPS = 1e-6f;

__device__ __forceinline__ float2 warp_sum2(float a, float b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    return make_float2(a, b);
}

// row -> (image, pixel) for row < 2^31 without an integer division
__device__ __forceinline__ void split_row(unsigned row, unsigned ppi, float inv_ppi, unsigned* b, unsigned* pix) {
    unsigned q = (unsigned)((float)row * inv_ppi);
    long long rem = (long long)row - (long long)q * ppi;
    if (rem < 0) { --q; rem += ppi; }
    else if (rem >= (long long)ppi) { ++q; rem -= ppi; }
    *b = q;
    *pix = (unsigned)rem;
}

// Row streaming: every thread keeps its vectors of the warp's next RING rows in flight as cp.async copies into a
// shared-memory ring (one commit group per row) and reads back only what it copied itself.  With one 416-byte row per
// warp in flight the kernels were latency-bound at ~1 TB/s; the ring keeps RING rows per warp in flight.
template <int NVL> struct RingDepth { static constexpr int value = NVL == 1 ? 8 : (NVL == 2 ? 4 : 2); };   // 32 KB of vectors per block

template <typename T, int NVL>
__device__ __forceinline__ void ring_issue(uint4* ring, int slot, const T* __restrict__ xr, int lane, int inner) {
    constexpr int V = VkVec<T>::N;
#pragma unroll
    for (int j = 0; j < NVL; ++j) {
        const int c = (lane + 32 * j) * V;
        if (c < inner) vk_cp_async16(ring + (slot * NVL + j) * HT_THREADS + threadIdx.x, xr + c);
    }
}

// The lane's channel vectors of the row in `slot` as fp32, zeros beyond `inner`.
template <typename T, int NVL>
__device__ __forceinline__ void ring_fetch(const uint4* ring, int slot, int lane, int inner, float (&f)[NVL][VkVec<T>::N]) {
    constexpr int V = VkVec<T>::N;
#pragma unroll
    for (int j = 0; j < NVL; ++j) {
        const int c = (lane + 32 * j) * V;
        if (c < inner) {
            VkVec<T> v;
            v.raw = *reinterpret_cast<const decltype(v.raw)*>(ring + (slot * NVL + j) * HT_THREADS + threadIdx.x);
            v.unpack(f[j]);
            if (c + V > inner) {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (c + i >= inner) f[j][i] = 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) f[j][i] = 0.f;
        }
    }
}

// mean / rstd of the `inner` real channels of a row held across the warp (pad channels hold zeros).
template <int NVL, int V>
__device__ __forceinline__ void row_stats(const float (&f)[NVL][V], int inner, int npad, float inv, float* mean, float* rstd) {
    const float x0 = __shfl_sync(0xffffffffu, f[0][0], 0);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float d = f[j][i] - x0;
            s += d;
            q = fmaf(d, d, q);
        }
    const float2 r = warp_sum2(s, q);
    // the zero-valued pad channels contributed (0 - x0) and (0 - x0)^2 each
    const float ss = r.x + (float)npad * x0;
    const float qq = r.y - (float)npad * x0 * x0;
    const float m = ss * inv;                         // mean - x0
    *mean = x0 + m;
    *rstd = rsqrtf(fmaxf(fmaf(-m, m, qq * inv), 0.f) + LN_EPS);
}

template <typename T, int NVL, int O>
__global__ void __launch_bounds__(HT_THREADS)
head_tail_fwd_kernel(const T* __restrict__ x, long long ld_x, int inner, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ w2, const float* __restrict__ b2, int softplus,
                     float* __restrict__ out, unsigned ppi, float inv_ppi, unsigned rows, const int* __restrict__ row_index) {
    // row_index (nullable): x row e is the conv output of pixel row_index[e] (< 0: skip) -- the label-point forward
    constexpr int V = VkVec<T>::N;
    const int lane = threadIdx.x & 31;
    const unsigned warp0 = blockIdx.x * HT_WARPS + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * HT_WARPS;
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
    const int npad = NVL * 32 * V - inner;
    float bias2 = (lane < O) ? __ldg(b2 + lane) : 0.f;
    constexpr int RING = RingDepth<NVL>::value;
    extern __shared__ uint4 ring[];   // [RING][NVL][HT_THREADS]
#pragma unroll 1
    for (int d = 0; d < RING; ++d) {
        const unsigned rr = warp0 + (unsigned)d * nwarps;
        if (rr < rows) ring_issue<T, NVL>(ring, d, x + (long long)rr * ld_x, lane, inner);
        vk_cp_async_commit();
    }
    int slot = 0;
    for (unsigned r = warp0; r < rows; r += nwarps) {
        float f[NVL][V];
        vk_cp_async_wait<RING - 1>();
        ring_fetch<T, NVL>(ring, slot, lane, inner, f);
        {
            const unsigned long long rr = (unsigned long long)r + (unsigned long long)RING * nwarps;
            if (rr < rows) ring_issue<T, NVL>(ring, slot, x + (long long)rr * ld_x, lane, inner);
            vk_cp_async_commit();
            slot = (slot + 1) & (RING - 1);
        }
        float mean, rstd;
        row_stats<NVL, V>(f, inner, npad, inv, &mean, &rstd);
        const float shift = -mean * rstd;
        float dot[O];
#pragma unroll
        for (int o = 0; o < O; ++o) dot[o] = 0.f;
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float h = fmaf(f[j][i], rstd, shift);
                const float g = vk_gelu(fmaf(h, gm[j][i], bt[j][i]));      // pad channels: gelu(0) = 0
#pragma unroll
                for (int o = 0; o < O; ++o) dot[o] = fmaf(g, w[o][j][i], dot[o]);
            }
#pragma unroll
        for (int o = 0; o < O; ++o) dot[o] = vk_warp_sum(dot[o]);
        if (lane < O) {
            float v = 0.f;
#pragma unroll
            for (int o = 0; o < O; ++o) v = (lane == o) ? dot[o] : v;
            v += bias2;
            if (softplus) v = vk_softplus(v);
            const long long pr = row_index ? (long long)__ldg(row_index + r) : (long long)r;
            if (pr >= 0) {
                unsigned b, pix;
                split_row((unsigned)pr, ppi, inv_ppi, &b, &pix);
                out[((long long)b * O + lane) * ppi + pix] = v;
            }
        }
    }
}

// The lane's channel vectors of the row in `slot` as fp32 PAIRS, zeros beyond `inner`.
template <typename T> struct PairUnpack;
template <> struct PairUnpack<float> {
    static __device__ __forceinline__ void run(const float4& raw, float2 (&p)[2]) { p[0] = make_float2(raw.x, raw.y); p[1] = make_float2(raw.z, raw.w); }
};
template <> struct PairUnpack<__nv_bfloat16> {
    static __device__ __forceinline__ void run(const uint4& raw, float2 (&p)[4]) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
    }
};
template <typename T, int NVL>
__device__ __forceinline__ void ring_fetch2(const uint4* ring, int slot, int lane, int inner, float2 (&f)[NVL][VkVec<T>::N / 2]) {
    constexpr int V = VkVec<T>::N;
#pragma unroll
    for (int j = 0; j < NVL; ++j) {
        const int c = (lane + 32 * j) * V;
        if (c < inner) {
            VkVec<T> v;
            v.raw = *reinterpret_cast<const decltype(v.raw)*>(ring + (slot * NVL + j) * HT_THREADS + threadIdx.x);
            PairUnpack<T>::run(v.raw, f[j]);
            if (c + V > inner) {
#pragma unroll
                for (int i = 0; i < V / 2; ++i) {
                    if (c + 2 * i >= inner) f[j][i].x = 0.f;
                    if (c + 2 * i + 1 >= inner) f[j][i].y = 0.f;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < V / 2; ++i) f[j][i] = make_float2(0.f, 0.f);
        }
    }
}

// Backward.  Per-channel gradient partials (dgamma, dbeta, conv-bias, dW2) live in registers of the lane that owns the
// channel for all rows the warp visits.  All per-element arithmetic runs on fp32 PAIRS (FFMA2 / FMUL2 / FADD2): the
// kernel is bound by instruction issue, and a packed instruction takes one issue slot for two channels.
template <typename T, int NVL, int O>
__global__ void __launch_bounds__(HT_THREADS, (NVL * O <= 4) ? 2 : 1)
head_tail_bwd_kernel(const T* __restrict__ x, long long ld_x, int inner, int slice_w, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ w2, int softplus, const float* __restrict__ out,
                     const float* __restrict__ dout, unsigned ppi, float inv_ppi, unsigned rows, T* __restrict__ dx,
                     long long ld_dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw2,
                     float* __restrict__ db2, float* __restrict__ dbias, const int* __restrict__ row_index, int x_compact) {
    // row_index (nullable): entry e reads pixel row row_index[e] (< 0: no pixel, the entry's gradient row is zero) and writes
    // dx row e -- the label-point form used when the upstream gradient is zero outside a short list of pixels
    constexpr int V = VkVec<T>::N;
    constexpr int P = V / 2;
    constexpr int CW = 32 * NVL * V;
    constexpr int RING = RingDepth<NVL>::value;
    extern __shared__ uint4 ring[];   // [RING][NVL][HT_THREADS] vectors, then [RING][HT_WARPS][2 * O] upstream values, then sacc
    float* dring = reinterpret_cast<float*>(ring + RING * NVL * HT_THREADS);
    float* sacc = dring + RING * HT_WARPS * 2 * O;   // [(3 + O)][CW] + [O]
    // O >= 3: the projection weights stay in shared memory (32 registers of them would spill the accumulators)
    constexpr bool W_SMEM = O >= 3;
    float* s_w = sacc + (3 + O) * CW + O + ((3 + O) * CW + O) % 2;   // [O][CW], only when W_SMEM (8-byte aligned)
    for (int i = threadIdx.x; i < (3 + O) * CW + O; i += blockDim.x) sacc[i] = 0.f;
    if (W_SMEM)
        for (int i = threadIdx.x; i < O * CW; i += blockDim.x) {
            const int o = i / CW, c = i % CW;
            s_w[i] = c < inner ? __ldg(w2 + (long long)o * inner + c) : 0.f;
        }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned warp0 = blockIdx.x * HT_WARPS + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * HT_WARPS;
    float2 gm[NVL][P], bt[NVL][P], w[W_SMEM ? 1 : O][NVL][P];
    float2 ag[NVL][P], ab[NVL][P], ax[NVL][P], aw[O][NVL][P];
    float adb[O];
#pragma unroll
    for (int o = 0; o < O; ++o) adb[o] = 0.f;
    auto ldc = [&](const float* p, int c) { return c < inner ? __ldg(p + c) : 0.f; };
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int c = (lane + 32 * j) * V + 2 * i;
            gm[j][i] = make_float2(ldc(gamma, c), ldc(gamma, c + 1));
            bt[j][i] = make_float2(ldc(beta, c), ldc(beta, c + 1));
            ag[j][i] = ab[j][i] = ax[j][i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int o = 0; o < O; ++o) {
                if (!W_SMEM) w[o][j][i] = make_float2(ldc(w2 + (long long)o * inner, c), ldc(w2 + (long long)o * inner, c + 1));
                aw[o][j][i] = make_float2(0.f, 0.f);
            }
        }
    const float inv = 1.f / inner;
    const int npad = CW - inner;
    const int wib = threadIdx.x >> 5;
    // issue one row: the lane's conv vectors, and (lanes < 2*O) the row's upstream gradient / saved output values
    auto issue = [&](unsigned long long rr, int sl) {
        long long pr = (long long)rr;
        if (rr < rows && row_index) pr = __ldg(row_index + rr);
        if (rr < rows && pr >= 0) {
            ring_issue<T, NVL>(ring, sl, x + (x_compact ? (long long)rr : pr) * ld_x, lane, inner);   // x_compact: x holds one row per ENTRY
            if (lane < 2 * O && (softplus || lane < O)) {
                unsigned b, pix;
                split_row((unsigned)pr, ppi, inv_ppi, &b, &pix);
                const int o = lane < O ? lane : lane - O;
                const long long oi = ((long long)b * O + o) * ppi + pix;
                vk_cp_async4(dring + (sl * HT_WARPS + wib) * 2 * O + lane, (lane < O ? dout : out) + oi);
            }
        }
        vk_cp_async_commit();
    };
#pragma unroll 1
    for (int d = 0; d < RING; ++d) issue((unsigned long long)warp0 + (unsigned long long)d * nwarps, d);
    int slot = 0;
    for (unsigned r = warp0; r < rows; r += nwarps) {
        float2 f[NVL][P];
        vk_cp_async_wait<RING - 1>();
        __syncwarp();                                   // the upstream values were copied by other lanes
        const bool live = !row_index || __ldg(row_index + r) >= 0;
        ring_fetch2<T, NVL>(ring, slot, lane, inner, f);
        if (!live) {
#pragma unroll
            for (int j = 0; j < NVL; ++j)
#pragma unroll
                for (int i = 0; i < P; ++i) f[j][i] = make_float2(0.f, 0.f);
        }
        // upstream gradient of the pre-softplus outputs (same value in every lane)
        float2 dpre[O];
#pragma unroll
        for (int o = 0; o < O; ++o) {
            const float* dr = dring + (slot * HT_WARPS + wib) * 2 * O;
            float d = live ? dr[o] : 0.f;
            if (softplus) {
                const float y = live ? dr[O + o] : 0.f;
                d *= (y > 20.f) ? 1.f : (1.f - __expf(-y));   // sigmoid(pre) = 1 - exp(-softplus(pre))
            }
            dpre[o] = vk_splat2(d);
            adb[o] += d;
        }
        __syncwarp();                                   // every lane has read the slot before it is refilled
        issue((unsigned long long)r + (unsigned long long)RING * nwarps, slot);
        slot = (slot + 1) & (RING - 1);
        // LayerNorm statistics: one shuffle round of shifted sums (pad channels hold zeros and are corrected for)
        float mean, rstd;
        {
            const float x0 = __shfl_sync(0xffffffffu, f[0][0].x, 0);
            const float2 nx0 = vk_splat2(-x0);
            float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NVL; ++j)
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const float2 d = vk_add2(f[j][i], nx0);
                    s = vk_add2(s, d);
                    q = vk_fma2(d, d, q);
                }
            const float2 rsum = warp_sum2(s.x + s.y, q.x + q.y);
            const float ss = rsum.x + (float)npad * x0;
            const float qq = rsum.y - (float)npad * x0 * x0;
            const float m = ss * inv;                         // mean - x0
            mean = x0 + m;
            rstd = rsqrtf(fmaxf(fmaf(-m, m, qq * inv), 0.f) + LN_EPS);
        }
        const float2 rstd2 = vk_splat2(rstd), shift2 = vk_splat2(-mean * rstd);
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
        // after this loop f holds xhat and dz holds d(loss)/d(LN output) * gamma per channel.  Pad channels: gamma = beta =
        // w = 0 there, so z = 0, dg = 0 and every product below vanishes without a mask.
        float2 dz[NVL][P];
#pragma unroll
        for (int j = 0; j < NVL; ++j)
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const float2 h = vk_fma2(f[j][i], rstd2, shift2);
                float2 g, gp;
                vk_gelu_both2(vk_fma2(h, gm[j][i], bt[j][i]), &g, &gp);
                float2 dg;
#pragma unroll
                for (int o = 0; o < O; ++o) {
                    const float2 wv = W_SMEM ? *reinterpret_cast<const float2*>(s_w + o * CW + (lane + 32 * j) * V + 2 * i) : w[o][j][i];
                    dg = o == 0 ? vk_mul2(dpre[0], wv) : vk_fma2(dpre[o], wv, dg);
                    aw[o][j][i] = vk_fma2(dpre[o], g, aw[o][j][i]);
                }
                const float2 d = vk_mul2(dg, gp);
                const float2 dxh = vk_mul2(d, gm[j][i]);
                f[j][i] = h;
                dz[j][i] = dxh;
                s1 = vk_add2(s1, dxh);
                s2 = vk_fma2(dxh, h, s2);
                ag[j][i] = vk_fma2(d, h, ag[j][i]);
                ab[j][i] = vk_add2(ab[j][i], d);
            }
        const float2 ss = warp_sum2(s1.x + s1.y, s2.x + s2.y);
        const float m1 = ss.x * inv, m2 = ss.y * inv;
        // dx = rstd * (dz - m1 - xhat * m2) = dz * rstd + (xhat * (-m2 rstd) + (-m1 rstd))
        const float2 cb = vk_splat2(-m1 * rstd), cc = vk_splat2(-m2 * rstd);
        T* dxr = dx + (long long)r * ld_dx;
#pragma unroll
        for (int j = 0; j < NVL; ++j) {
            const int c0 = (lane + 32 * j) * V;
            if (c0 < slice_w) {
                float fo[V];
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    float2 dxv = vk_fma2(dz[j][i], rstd2, vk_fma2(f[j][i], cc, cb));
                    if (c0 + V > inner) {
                        if (c0 + 2 * i >= inner) dxv.x = 0.f;
                        if (c0 + 2 * i + 1 >= inner) dxv.y = 0.f;
                    }
                    fo[2 * i] = dxv.x;
                    fo[2 * i + 1] = dxv.y;
                    ax[j][i] = vk_add2(ax[j][i], dxv);
                }
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
                atomicAdd(&sacc[c], (i & 1) ? ag[j][i / 2].y : ag[j][i / 2].x);
                atomicAdd(&sacc[CW + c], (i & 1) ? ab[j][i / 2].y : ab[j][i / 2].x);
                atomicAdd(&sacc[2 * CW + c], (i & 1) ? ax[j][i / 2].y : ax[j][i / 2].x);
#pragma unroll
                for (int o = 0; o < O; ++o) atomicAdd(&sacc[(3 + o) * CW + c], (i & 1) ? aw[o][j][i / 2].y : aw[o][j][i / 2].x);
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

// Backward for the shape every DENSE call of the training step has (bf16 storage, inner = 192, one output map: both rough
// heads and the precise mask head): a HALF-warp per pixel row.  With a warp per row 8 of the 32 lanes idle (192 channels =
// 24 16-byte vectors) and the per-row work that is not per-element -- two shuffle reductions, the ring bookkeeping, the
// upstream scalars -- is paid once per 6 elements of a lane.  Here every lane owns 12 channels (three 8-byte vectors, 16
// lanes apart: a half-warp covers a 128-byte line per vector) and one round of shuffles serves the two rows of the warp.
constexpr int HH_RING = 4;             // rows in flight per half-warp (12 KB per warp, as in the generic kernel)
constexpr int HH_INNER = 192;
constexpr int HH_NV = 3;               // 8-byte vectors per lane
__device__ __forceinline__ void vk_cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float2 half_sum2(float a, float b) {      // sums over the 16 lanes of a half-warp
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    return make_float2(a, b);
}
__global__ void __launch_bounds__(HT_THREADS, 2)
head_tail_bwd_h192_kernel(const __nv_bfloat16* __restrict__ x, long long ld_x, int slice_w, const float* __restrict__ gamma,
                          const float* __restrict__ beta, const float* __restrict__ w2, int softplus, const float* __restrict__ out,
                          const float* __restrict__ dout, unsigned rows, __nv_bfloat16* __restrict__ dx, long long ld_dx,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw2, float* __restrict__ db2,
                          float* __restrict__ dbias) {
    // one output map: the NCHW fp32 maps out / dout are indexed by the pixel row itself
    extern __shared__ uint4 ring[];
    uint2* ring2 = reinterpret_cast<uint2*>(ring);                               // [HH_RING][HH_NV][HT_THREADS]
    float* dring = reinterpret_cast<float*>(ring2 + HH_RING * HH_NV * HT_THREADS);   // [HH_RING][2 * HT_WARPS][2]: dout, out
    float* sacc = dring + HH_RING * 2 * HT_WARPS * 2;                            // [4][HH_INNER] + [1]
    for (int i = threadIdx.x; i < 4 * HH_INNER + 1; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4, wib = threadIdx.x >> 5;
    const unsigned row0 = (blockIdx.x * HT_WARPS + wib) * 2;
    const unsigned stride = gridDim.x * HT_WARPS * 2;
    float2 gm[HH_NV][2], bt[HH_NV][2], w[HH_NV][2], ag[HH_NV][2], ab[HH_NV][2], ax[HH_NV][2], aw[HH_NV][2];
    float adb = 0.f;
#pragma unroll
    for (int j = 0; j < HH_NV; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = (l16 + 16 * j) * 4 + 2 * i;
            gm[j][i] = make_float2(__ldg(gamma + c), __ldg(gamma + c + 1));
            bt[j][i] = make_float2(__ldg(beta + c), __ldg(beta + c + 1));
            w[j][i] = make_float2(__ldg(w2 + c), __ldg(w2 + c + 1));
            ag[j][i] = ab[j][i] = ax[j][i] = aw[j][i] = make_float2(0.f, 0.f);
        }
    const float inv = 1.f / HH_INNER;
    const int pad_vecs = (slice_w - HH_INNER) / 4;       // zero-filled pad columns of the dx slice, 4 per lane
    float* my_d = dring + (2 * wib + half) * 2;
    auto issue = [&](unsigned long long rb, int sl) {
        const unsigned long long r = rb + (unsigned)half;
        if (r < rows) {
            const __nv_bfloat16* xr = x + (long long)r * ld_x + l16 * 4;
#pragma unroll
            for (int j = 0; j < HH_NV; ++j) vk_cp_async8(ring2 + (sl * HH_NV + j) * HT_THREADS + threadIdx.x, xr + 64 * j);
            if (l16 == 0 || (l16 == 1 && softplus)) vk_cp_async4(my_d + sl * 2 * HT_WARPS * 2 + l16, (l16 == 0 ? dout : out) + r);
        }
        vk_cp_async_commit();
    };
#pragma unroll 1
    for (int d = 0; d < HH_RING; ++d) issue((unsigned long long)row0 + (unsigned long long)d * stride, d);
    int slot = 0;
    for (unsigned rb = row0; rb < rows; rb += stride) {
        const bool live = rb + (unsigned)half < rows;
        vk_cp_async_wait<HH_RING - 1>();
        __syncwarp();                                   // the upstream values were copied by other lanes
        float2 f[HH_NV][2];
#pragma unroll
        for (int j = 0; j < HH_NV; ++j) {
            uint2 raw = make_uint2(0u, 0u);
            if (live) raw = ring2[(slot * HH_NV + j) * HT_THREADS + threadIdx.x];
            f[j][0] = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
            f[j][1] = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
        }
        float d = live ? my_d[slot * 2 * HT_WARPS * 2] : 0.f;
        if (softplus) {
            const float y = live ? my_d[slot * 2 * HT_WARPS * 2 + 1] : 0.f;
            d *= (y > 20.f) ? 1.f : (1.f - __expf(-y));   // sigmoid(pre) = 1 - exp(-softplus(pre))
        }
        adb += d;
        const float2 dpre = vk_splat2(d);
        __syncwarp();                                   // every lane has read the slot before it is refilled
        issue((unsigned long long)rb + (unsigned long long)HH_RING * stride, slot);
        slot = (slot + 1) & (HH_RING - 1);
        float mean, rstd;
        {
            const float x0 = __shfl_sync(0xffffffffu, f[0][0].x, lane & 16);
            const float2 nx0 = vk_splat2(-x0);
            float2 sm = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < HH_NV; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float2 dd = vk_add2(f[j][i], nx0);
                    sm = vk_add2(sm, dd);
                    q = vk_fma2(dd, dd, q);
                }
            const float2 rsum = half_sum2(sm.x + sm.y, q.x + q.y);
            const float m = rsum.x * inv;                     // mean - x0
            mean = x0 + m;
            rstd = rsqrtf(fmaxf(fmaf(-m, m, rsum.y * inv), 0.f) + LN_EPS);
        }
        const float2 rstd2 = vk_splat2(rstd), shift2 = vk_splat2(-mean * rstd);
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
        float2 dz[HH_NV][2];
#pragma unroll
        for (int j = 0; j < HH_NV; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float2 h = vk_fma2(f[j][i], rstd2, shift2);
                float2 g, gp;
                vk_gelu_both2(vk_fma2(h, gm[j][i], bt[j][i]), &g, &gp);
                aw[j][i] = vk_fma2(dpre, g, aw[j][i]);
                const float2 dd = vk_mul2(vk_mul2(dpre, w[j][i]), gp);
                const float2 dxh = vk_mul2(dd, gm[j][i]);
                f[j][i] = h;
                dz[j][i] = dxh;
                s1 = vk_add2(s1, dxh);
                s2 = vk_fma2(dxh, h, s2);
                ag[j][i] = vk_fma2(dd, h, ag[j][i]);
                ab[j][i] = vk_add2(ab[j][i], dd);
            }
        const float2 ss = half_sum2(s1.x + s1.y, s2.x + s2.y);
        // dx = rstd * (dz - m1 - xhat * m2) = dz * rstd + (xhat * (-m2 rstd) + (-m1 rstd))
        const float2 cb = vk_splat2(-ss.x * inv * rstd), cc = vk_splat2(-ss.y * inv * rstd);
        if (live) {
            __nv_bfloat16* dxr = dx + (long long)(rb + (unsigned)half) * ld_dx;
#pragma unroll
            for (int j = 0; j < HH_NV; ++j) {
                const float2 d0 = vk_fma2(dz[j][0], rstd2, vk_fma2(f[j][0], cc, cb));
                const float2 d1 = vk_fma2(dz[j][1], rstd2, vk_fma2(f[j][1], cc, cb));
                ax[j][0] = vk_add2(ax[j][0], d0);
                ax[j][1] = vk_add2(ax[j][1], d1);
                uint2 pk;
                *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(d0.x, d0.y);
                *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(d1.x, d1.y);
                *reinterpret_cast<uint2*>(dxr + (l16 + 16 * j) * 4) = pk;
            }
            if (l16 < pad_vecs) *reinterpret_cast<uint2*>(dxr + HH_INNER + l16 * 4) = make_uint2(0u, 0u);
        }
    }
#pragma unroll
    for (int j = 0; j < HH_NV; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = (l16 + 16 * j) * 4 + 2 * i;
            atomicAdd(&sacc[c], ag[j][i].x);                   atomicAdd(&sacc[c + 1], ag[j][i].y);
            atomicAdd(&sacc[HH_INNER + c], ab[j][i].x);        atomicAdd(&sacc[HH_INNER + c + 1], ab[j][i].y);
            atomicAdd(&sacc[2 * HH_INNER + c], ax[j][i].x);    atomicAdd(&sacc[2 * HH_INNER + c + 1], ax[j][i].y);
            atomicAdd(&sacc[3 * HH_INNER + c], aw[j][i].x);    atomicAdd(&sacc[3 * HH_INNER + c + 1], aw[j][i].y);
        }
    if (l16 == 0) atomicAdd(&sacc[4 * HH_INNER], adb);
    __syncthreads();
    for (int c = threadIdx.x; c < HH_INNER; c += blockDim.x) {
        atomicAdd(dgamma + c, sacc[c]);
        atomicAdd(dbeta + c, sacc[HH_INNER + c]);
        atomicAdd(dbias + c, sacc[2 * HH_INNER + c]);
        atomicAdd(dw2 + c, sacc[3 * HH_INNER + c]);
    }
    if (threadIdx.x == 0) atomicAdd(db2, sacc[4 * HH_INNER]);
}

int launch_bwd_h192(const void* x, long long ld_x, int slice_w, const float* gamma, const float* beta, const float* w2, int softplus,
                    const float* out, const float* dout, long long rows, void* dx, long long ld_dx, float* dgamma, float* dbeta,
                    float* dw2, float* db2, float* dbias, cudaStream_t s) {
    long long blocks = (rows + 2 * HT_WARPS - 1) / (2 * HT_WARPS);
    const long long cap = (long long)vkocr_sm_count() * 2;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)HH_RING * HH_NV * HT_THREADS * 8 + (HH_RING * 2 * HT_WARPS * 2 + 4 * HH_INNER + 1) * sizeof(float);
    head_tail_bwd_h192_kernel<<<(unsigned)blocks, HT_THREADS, smem, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ld_x, slice_w, gamma, beta, w2, softplus, out, dout, (unsigned)rows,
        reinterpret_cast<__nv_bfloat16*>(dx), ld_dx, dgamma, dbeta, dw2, db2, dbias);
    return 0;
}

template <typename T, int NVL>
int launch_fwd(int O, const void* x, long long ld_x, int inner, const float* gamma, const float* beta, const float* w2,
               const float* b2, int softplus, float* out, long long ppi, long long rows, const int* row_index, cudaStream_t s) {
    long long blocks = (rows + HT_WARPS - 1) / HT_WARPS;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const float inv_ppi = 1.f / (float)ppi;
    constexpr int RING = RingDepth<NVL>::value;
#define VK_HT_FWD(OO)                                                                                            \
    head_tail_fwd_kernel<T, NVL, OO><<<(unsigned)blocks, HT_THREADS, RING * NVL * HT_THREADS * 16, s>>>(reinterpret_cast<const T*>(x), ld_x, inner, \
                                                                             gamma, beta, w2, b2, softplus, out,  \
                                                                             (unsigned)ppi, inv_ppi, (unsigned)rows, row_index)
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
               long long ld_dx, float* dgamma, float* dbeta, float* dw2, float* db2, float* dbias, const int* row_index, int x_compact, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    long long blocks = (rows + HT_WARPS - 1) / HT_WARPS;
    const long long cap = (long long)vkocr_sm_count() * ((NVL * O <= 4) ? 2 : 1);
    if (blocks > cap) blocks = cap;
    const float inv_ppi = 1.f / (float)ppi;
    constexpr int RING = RingDepth<NVL>::value;
#define VK_HT_BWD(OO)                                                                                                       \
    cudaFuncSetAttribute(head_tail_bwd_kernel<T, NVL, OO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);         \
    head_tail_bwd_kernel<T, NVL, OO><<<(unsigned)blocks, HT_THREADS,                                                        \
        RING * NVL * HT_THREADS * 16 + (RING * HT_WARPS * 2 * OO + (3 + OO + (OO >= 3 ? OO : 0)) * 32 * NVL * V + OO + 1) * sizeof(float), s>>>(      \
        reinterpret_cast<const T*>(x), ld_x, inner, slice_w, gamma, beta, w2, softplus, out, dout, (unsigned)ppi, inv_ppi,  \
        (unsigned)rows, reinterpret_cast<T*>(dx), ld_dx, dgamma, dbeta, dw2, db2, dbias, row_index, x_compact)
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
    VK_REQUIRE(rows < (1LL << 31) && pixels_per_image >= 1 && pixels_per_image < (1LL << 31), VKOCR_BAD_SHAPE,
               "head_tail_fwd: %lld rows / %lld pixels per image out of range", rows, pixels_per_image);
    const int nvl = pick_nvl(dtype, slice_w);
    VK_REQUIRE(nvl > 0, VKOCR_BAD_SHAPE, "head_tail_fwd: inner width %d too large", slice_w);
    if (rows == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
#define VK_CALL(NVL) VK_DISPATCH_DTYPE(dtype, T, (rc = launch_fwd<T, NVL>(O, x, ld_x, inner, gamma, beta, w2, b2, softplus, out, pixels_per_image, rows, nullptr, s)))
    if (nvl == 1) VK_CALL(1);
    else if (nvl == 2) VK_CALL(2);
    else VK_CALL(4);
#undef VK_CALL
    VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_tail_fwd: dispatch failed");
    VK_CHECK_LAUNCH("head_tail_fwd_kernel");
    return VKOCR_OK;
}

// The same tail over a COMPACT list of conv-output rows: x = [entries, ld_x], entry e belongs to pixel row_index[e] (global
// pixel index b * pixels_per_image + y * W + x; < 0 = skip) of the NCHW fp32 map `out`, which is written at those pixels
// only.  The label-point forward of the precise heads (training: the loss reads these maps at the label points alone).
int vkocr_head_tail_fwd_points(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                               const float* w2, const float* b2, int O, int softplus, float* out, long long pixels_per_image,
                               const int* row_index, long long entries, void* stream) {
    VK_REQUIRE(x && gamma && beta && w2 && b2 && out && row_index, VKOCR_BAD_ARGUMENT, "head_tail_fwd_points: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(slice_w % V == 0 && ld_x % V == 0 && slice_w >= inner, VKOCR_BAD_ALIGN, "head_tail_fwd_points: slice %d ld %lld", slice_w, ld_x);
    VK_REQUIRE(O >= 1 && O <= 4, VKOCR_BAD_SHAPE, "head_tail_fwd_points: out channels %d (1..4 supported)", O);
    VK_REQUIRE(entries < (1LL << 31) && pixels_per_image >= 1 && pixels_per_image < (1LL << 31), VKOCR_BAD_SHAPE,
               "head_tail_fwd_points: %lld entries / %lld pixels per image out of range", entries, pixels_per_image);
    const int nvl = pick_nvl(dtype, slice_w);
    VK_REQUIRE(nvl > 0, VKOCR_BAD_SHAPE, "head_tail_fwd_points: inner width %d too large", slice_w);
    if (entries == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
#define VK_CALL(NVL) VK_DISPATCH_DTYPE(dtype, T, (rc = launch_fwd<T, NVL>(O, x, ld_x, inner, gamma, beta, w2, b2, softplus, out, pixels_per_image, entries, row_index, s)))
    if (nvl == 1) VK_CALL(1);
    else if (nvl == 2) VK_CALL(2);
    else VK_CALL(4);
#undef VK_CALL
    VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_tail_fwd_points: dispatch failed");
    VK_CHECK_LAUNCH("head_tail_fwd_kernel(points)");
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
    VK_REQUIRE(rows < (1LL << 31) && pixels_per_image >= 1 && pixels_per_image < (1LL << 31), VKOCR_BAD_SHAPE,
               "head_tail_bwd: %lld rows / %lld pixels per image out of range", rows, pixels_per_image);
    const int nvl = pick_nvl(dtype, slice_w);
    VK_REQUIRE(nvl > 0, VKOCR_BAD_SHAPE, "head_tail_bwd: inner width %d too large", slice_w);
    if (rows == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
    static const bool generic_only = getenv("VKOCR_HEAD_TAIL_GENERIC") != nullptr;     // A/B switch for tools/kbench.py
    if (dtype == VKOCR_BF16 && inner == HH_INNER && O == 1 && slice_w - HH_INNER <= 64 && ld_x % 4 == 0 && ld_dx % 4 == 0 && !generic_only) {
        launch_bwd_h192(x, ld_x, slice_w, gamma, beta, w2, softplus, out, dout, rows, dx, ld_dx, dgamma, dbeta, dw2, db2, dbias, s);
        VK_CHECK_LAUNCH("head_tail_bwd_h192_kernel");
        return VKOCR_OK;
    }
#define VK_CALL(NVL) VK_DISPATCH_DTYPE(dtype, T, (rc = launch_bwd<T, NVL>(O, x, ld_x, inner, slice_w, gamma, beta, w2, softplus, out, dout, pixels_per_image, rows, dx, ld_dx, dgamma, dbeta, dw2, db2, dbias, nullptr, 0, s)))
    if (nvl == 1) VK_CALL(1);
    else if (nvl == 2) VK_CALL(2);
    else VK_CALL(4);
#undef VK_CALL
    VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_tail_bwd: dispatch failed");
    VK_CHECK_LAUNCH("head_tail_bwd_kernel");
    return VKOCR_OK;
}

// The same backward restricted to a list of pixels: entry e < entries reads the conv / out / dout values of pixel row
// row_index[e] (global pixel index b * pixels_per_image + y * W + x; < 0 = no pixel) and writes the gradient row e of dx
// ([entries, ld_dx], zeros for empty entries).  Used when d(loss)/d(out) is zero outside the label points.  x_compact != 0: x itself
// holds one conv-output row per ENTRY (the label-point forward kept nothing else).
int vkocr_head_tail_bwd_points(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                               const float* w2, int O, int softplus, const float* out, const float* dout, long long pixels_per_image,
                               const int* row_index, long long entries, int x_compact, void* dx, long long ld_dx, float* dgamma,
                               float* dbeta, float* dw2, float* db2, float* dbias, void* stream) {
    VK_REQUIRE(x && gamma && beta && w2 && out && dout && dx && dgamma && dbeta && dw2 && db2 && dbias && row_index, VKOCR_BAD_ARGUMENT,
               "head_tail_bwd_points: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(slice_w % V == 0 && ld_x % V == 0 && ld_dx % V == 0 && slice_w >= inner, VKOCR_BAD_ALIGN, "head_tail_bwd_points: slice %d", slice_w);
    VK_REQUIRE(O >= 1 && O <= 4, VKOCR_BAD_SHAPE, "head_tail_bwd_points: out channels %d (1..4 supported)", O);
    VK_REQUIRE(entries < (1LL << 31) && pixels_per_image >= 1 && pixels_per_image < (1LL << 31), VKOCR_BAD_SHAPE,
               "head_tail_bwd_points: %lld entries / %lld pixels per image out of range", entries, pixels_per_image);
    const int nvl = pick_nvl(dtype, slice_w);
    VK_REQUIRE(nvl > 0, VKOCR_BAD_SHAPE, "head_tail_bwd_points: inner width %d too large", slice_w);
    if (entries == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = 0;
#define VK_CALL(NVL) VK_DISPATCH_DTYPE(dtype, T, (rc = launch_bwd<T, NVL>(O, x, ld_x, inner, slice_w, gamma, beta, w2, softplus, out, dout, pixels_per_image, entries, dx, ld_dx, dgamma, dbeta, dw2, db2, dbias, row_index, x_compact, s)))
    if (nvl == 1) VK_CALL(1);
    else if (nvl == 2) VK_CALL(2);
    else VK_CALL(4);
#undef VK_CALL
    VK_REQUIRE(rc == 0, VKOCR_BAD_SHAPE, "head_tail_bwd_points: dispatch failed");
    VK_CHECK_LAUNCH("head_tail_bwd_kernel(points)");
    return VKOCR_OK;
}

}  // extern "C"
