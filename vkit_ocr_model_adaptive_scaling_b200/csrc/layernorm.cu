// Channels-last LayerNorm (+ optional fused exact GELU) forward and backward, and column sums (bias gradients).
// Reference: helper.ln = nn.LayerNorm(C, eps=1e-6) applied on BHWC (model/helper.py:96-97), followed by GELU in the
// neck/head conv blocks (upernext.py:21-45, fpn.py:21-48).
//
// HBM-bound streaming kernels.  A pixel row is only 192 B .. 1.5 KB, so the fast path maps a row onto a GROUP of G lanes
// (G = 4..32, 1-3 16-byte vectors per lane: C = 96 -> 8 rows per warp instruction) and keeps several row-iterations per
// thread in flight through a cp.async ring in shared memory: one row per warp in flight is latency-bound at ~1 TB/s.
// Statistics: ONE shuffle round over the group of shifted sums (sum (x - x0), sum (x - x0)^2 with x0 the row's first
// element: no cancellation, no second sweep).  Backward also produces the three per-channel reductions the caller
// needs: dgamma += sum_rows dz * xhat,  dbeta += sum_rows dz,  dxsum += sum_rows dx  (= bias gradient of the producer),
// from per-lane register partials merged per block in shared memory.  Odd channel counts / unaligned strides take the
// generic one-warp-per-row kernels at the end of the file.
#include "common.cuh"

namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;

template <typename T>
__device__ __forceinline__ float row_sum(const T* __restrict__ row, int C, bool vec, int lane, float shift, bool square) {
    float s = 0.f;
    if (vec) {
        constexpr int V = VkVec<T>::N;
        for (int c = lane * V; c < C; c += 32 * V) {
            VkVec<T> v;
            v.load(row + c);
            float f[V];
            v.unpack(f);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float d = f[i] - shift;
                s += square ? d * d : d;
            }
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            const float d = vk_to_f32(row[c]) - shift;
            s += square ? d * d : d;
        }
    }
    return vk_warp_sum(s);
}

template <typename T>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_generic_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, long long rows, int C,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
              float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * LN_WARPS;
    constexpr int V = VkVec<T>::N;
    const bool vec = (C % V == 0) && (ld_x % V == 0) && (ld_y % V == 0);
    for (long long r = warp0; r < rows; r += nwarps) {
        const T* xr = x + r * ld_x;
        T* yr = y + r * ld_y;
        const float mean = row_sum(xr, C, vec, lane, 0.f, false) / C;
        const float var = row_sum(xr, C, vec, lane, mean, true) / C;
        const float rstd = rsqrtf(var + eps);
        if (lane == 0 && mean_out) {
            mean_out[r] = mean;
            rstd_out[r] = rstd;
        }
        if (vec) {
            for (int c = lane * V; c < C; c += 32 * V) {
                VkVec<T> v;
                v.load(xr + c);
                float f[V];
                v.unpack(f);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float z = (f[i] - mean) * rstd * __ldg(gamma + c + i) + __ldg(beta + c + i);
                    f[i] = act ? vk_gelu(z) : z;
                }
                v.pack(f);
                v.store(yr + c);
            }
        } else {
            for (int c = lane; c < C; c += 32) {
                float z = (vk_to_f32(xr[c]) - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
                yr[c] = vk_from_f32<T>(act ? vk_gelu(z) : z);
            }
        }
    }
}

// Backward.  Each lane owns fixed channel vectors (lane, lane+32, ...) so its column partials stay in registers over
// all rows the warp visits; they are merged per block in shared memory and flushed with one atomic per channel.
// NV = number of channel vectors per lane (0 selects the generic scalar path for odd C / unaligned strides).
template <typename T, int NV>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const T* __restrict__ dy, long long ld_dy, const T* __restrict__ x, long long ld_x,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* __restrict__ beta, int act, T* __restrict__ dx, long long ld_dx, long long rows, int C,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
    extern __shared__ float sacc[];  // [3][C]
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * LN_WARPS;
    constexpr int V = VkVec<T>::N;
    const float invC = 1.f / C;
    if constexpr (NV > 0) {
        float ag[NV][V], ab[NV][V], ax[NV][V];
        float gm[NV][V], bt[NV][V];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (lane + 32 * j) * V;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                ag[j][i] = ab[j][i] = ax[j][i] = 0.f;
                gm[j][i] = (c < C) ? __ldg(gamma + c + i) : 0.f;
                bt[j][i] = (c < C) ? __ldg(beta + c + i) : 0.f;
            }
        }
        for (long long r = warp0; r < rows; r += nwarps) {
            const T* xr = x + r * ld_x;
            const T* dyr = dy + r * ld_dy;
            T* dxr = dx + r * ld_dx;
            const float mu = mean[r], rs = rstd[r];
            float xh[NV][V], dz[NV][V];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c = (lane + 32 * j) * V;
                if (c < C) {
                    VkVec<T> vx, vd;
                    vx.load(xr + c);
                    vd.load(dyr + c);
                    vx.unpack(xh[j]);
                    vd.unpack(dz[j]);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        xh[j][i] = (xh[j][i] - mu) * rs;
                        if (act) dz[j][i] *= vk_gelu_grad(xh[j][i] * gm[j][i] + bt[j][i]);
                        const float dxh = dz[j][i] * gm[j][i];
                        s1 += dxh;
                        s2 += dxh * xh[j][i];
                    }
                }
            }
            s1 = vk_warp_sum(s1) * invC;
            s2 = vk_warp_sum(s2) * invC;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c = (lane + 32 * j) * V;
                if (c < C) {
                    float fo[V];
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float dxv = rs * (dz[j][i] * gm[j][i] - s1 - xh[j][i] * s2);
                        fo[i] = dxv;
                        ag[j][i] += dz[j][i] * xh[j][i];
                        ab[j][i] += dz[j][i];
                        ax[j][i] += dxv;
                    }
                    VkVec<T> vo;
                    vo.pack(fo);
                    vo.store(dxr + c);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (lane + 32 * j) * V;
            if (c < C) {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    atomicAdd(&sacc[c + i], ag[j][i]);
                    atomicAdd(&sacc[C + c + i], ab[j][i]);
                    atomicAdd(&sacc[2 * C + c + i], ax[j][i]);
                }
            }
        }
    } else {
        for (long long r = warp0; r < rows; r += nwarps) {
            const T* xr = x + r * ld_x;
            const T* dyr = dy + r * ld_dy;
            T* dxr = dx + r * ld_dx;
            const float mu = mean[r], rs = rstd[r];
            float s1 = 0.f, s2 = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float xh = (vk_to_f32(xr[c]) - mu) * rs;
                const float g = __ldg(gamma + c);
                float dz = vk_to_f32(dyr[c]);
                if (act) dz *= vk_gelu_grad(xh * g + __ldg(beta + c));
                s1 += dz * g;
                s2 += dz * g * xh;
            }
            s1 = vk_warp_sum(s1) * invC;
            s2 = vk_warp_sum(s2) * invC;
            for (int c = lane; c < C; c += 32) {
                const float xh = (vk_to_f32(xr[c]) - mu) * rs;
                const float g = __ldg(gamma + c);
                float dz = vk_to_f32(dyr[c]);
                if (act) dz *= vk_gelu_grad(xh * g + __ldg(beta + c));
                const float dxv = rs * (dz * g - s1 - xh * s2);
                dxr[c] = vk_from_f32<T>(dxv);
                atomicAdd(&sacc[c], dz * xh);
                atomicAdd(&sacc[C + c], dz);
                atomicAdd(&sacc[2 * C + c], dxv);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + i, sacc[i]);
        if (dbeta) atomicAdd(dbeta + i, sacc[C + i]);
        if (dxsum) atomicAdd(dxsum + i, sacc[2 * C + i]);
    }
}

template <typename T>
int launch_ln_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, const float* mean, const float* rstd,
                  const float* gamma, const float* beta, int act, void* dx, long long ld_dx, long long rows, int C,
                  float* dgamma, float* dbeta, float* dxsum, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    const bool vec = (C % V == 0) && (ld_x % V == 0) && (ld_dy % V == 0) && (ld_dx % V == 0);
    const int nv = vec ? (C / V + 31) / 32 : 0;
    long long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)vkocr_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)3 * C * sizeof(float);
#define VK_LN_BWD(NVV)                                                                                                   \
    ln_bwd_kernel<T, NVV><<<(unsigned)blocks, LN_THREADS, smem, s>>>(                                                    \
        reinterpret_cast<const T*>(dy), ld_dy, reinterpret_cast<const T*>(x), ld_x, mean, rstd, gamma, beta, act,       \
        reinterpret_cast<T*>(dx), ld_dx, rows, C, dgamma, dbeta, dxsum)
    if (nv == 0 || nv > 8) VK_LN_BWD(0);
    else if (nv == 1) VK_LN_BWD(1);
    else if (nv == 2) VK_LN_BWD(2);
    else if (nv <= 4) VK_LN_BWD(4);
    else VK_LN_BWD(8);
#undef VK_LN_BWD
    return 0;
}

// Column sums of a [rows, C] matrix (bias gradients): out[c] += scale[c]? * sum_r x[r, c]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_generic_kernel(const T* __restrict__ x, long long ld, long long rows, int C, float* __restrict__ out) {
    // block = 32 column-lanes x 8 row-lanes; grid.x tiles columns, grid.y strides rows
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (c < C)
        for (long long r = (long long)blockIdx.y * 8 + ry; r < rows; r += (long long)gridDim.y * 8) s += vk_to_f32(x[r * ld + c]);
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cx];
        atomicAdd(out + c, t);
    }
}


// ------------------------------------------------------------------------------------------------- fast path
constexpr int FT = 256;   // threads per block

template <int G>
__device__ __forceinline__ float2 group_sum2(float a, float b) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    return make_float2(a, b);
}

// ring depth: about 8 KB .. 24 KB of vectors per block
#ifndef VK_LN_DEEP
#define VK_LN_DEEP 1
#endif
template <int NV, int STREAMS> struct LnRing {
    static constexpr int value = VK_LN_DEEP ? ((NV * STREAMS >= 6) ? 2 : ((NV * STREAMS >= 4) ? 4 : ((NV * STREAMS >= 3) ? 4 : 8)))
                                            : ((NV * STREAMS >= 6) ? 2 : ((NV * STREAMS >= 3) ? 3 : 4));
};

// y = LN(x) (* optional GELU).  G lanes per row, NV vectors per lane; R = 32 / G rows per warp iteration.
template <typename T, int G, int NV>
__global__ void __launch_bounds__(FT)
ln_fwd_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, long long rows, int C,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
              float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    constexpr int V = VkVec<T>::N;
    constexpr int R = 32 / G;
    constexpr int D = LnRing<NV, 1>::value;
    extern __shared__ uint4 ln_smem[];
    uint4* ring = ln_smem;                                              // [D][NV][FT]
    float* s_gamma = reinterpret_cast<float*>(ln_smem + D * NV * FT);   // [G * NV * V]
    float* s_beta = s_gamma + G * NV * V;
    const int nvec = C / V;
    for (int i = threadIdx.x; i < G * NV * V; i += FT) {
        s_gamma[i] = i < C ? gamma[i] : 0.f;
        s_beta[i] = i < C ? beta[i] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane & (G - 1), rsub = lane / G;
    const long long it0 = (long long)blockIdx.x * (FT / 32) + (threadIdx.x >> 5);
    const long long nit = (long long)gridDim.x * (FT / 32);
    const long long iters = (rows + R - 1) / R;
    const float invC = 1.f / C;
    const float npad = (float)((G * NV - nvec) * V);

    auto issue = [&](long long it, int slot) {
        const long long r = it * R + rsub;
        if (it < iters && r < rows) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = g + G * j;
                if (v < nvec) vk_cp_async16(ring + (slot * NV + j) * FT + threadIdx.x, x + r * ld_x + v * V);
            }
        }
        vk_cp_async_commit();
    };
#pragma unroll 1
    for (int d = 0; d < D; ++d) issue(it0 + d * nit, d);
    int slot = 0;
    for (long long it = it0; it < iters; it += nit) {
        const long long r = it * R + rsub;
        const bool row_ok = r < rows;
        float f[NV][V];
        vk_cp_async_wait<D - 1>();
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int v = g + G * j;
            if (row_ok && v < nvec) {
                VkVec<T> t;
                t.raw = *reinterpret_cast<const decltype(t.raw)*>(ring + (slot * NV + j) * FT + threadIdx.x);
                t.unpack(f[j]);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) f[j][i] = 0.f;
            }
        }
        issue(it + (long long)D * nit, slot);
        slot = (slot + 1 == D) ? 0 : slot + 1;
        const float x0 = __shfl_sync(0xffffffffu, f[0][0], rsub * G);
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float d = f[j][i] - x0;
                s += d;
                q = fmaf(d, d, q);
            }
        const float2 sq = group_sum2<G>(s, q);
        const float m = (sq.x + npad * x0) * invC;                         // mean - x0 (pad slots hold zeros)
        const float var = fmaxf(fmaf(-m, m, (sq.y - npad * x0 * x0) * invC), 0.f);
        const float mean = x0 + m;
        const float rstd = rsqrtf(var + eps);
        if (row_ok) {
            if (g == 0 && mean_out) {
                mean_out[r] = mean;
                rstd_out[r] = rstd;
            }
            const float shift = -mean * rstd;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = g + G * j;
                if (v < nvec) {
                    float o[V];
#pragma unroll
                    for (int i = 0; i < V; i += 4) {
                        const float4 gm = *reinterpret_cast<const float4*>(s_gamma + v * V + i);
                        const float4 bt = *reinterpret_cast<const float4*>(s_beta + v * V + i);
                        o[i] = fmaf(fmaf(f[j][i], rstd, shift), gm.x, bt.x);
                        o[i + 1] = fmaf(fmaf(f[j][i + 1], rstd, shift), gm.y, bt.y);
                        o[i + 2] = fmaf(fmaf(f[j][i + 2], rstd, shift), gm.z, bt.z);
                        o[i + 3] = fmaf(fmaf(f[j][i + 3], rstd, shift), gm.w, bt.w);
                    }
                    if (act) {
#pragma unroll
                        for (int i = 0; i < V; i += 2) {
                            const float2 g2 = vk_gelu2(make_float2(o[i], o[i + 1]));
                            o[i] = g2.x; o[i + 1] = g2.y;
                        }
                    }
                    VkVec<T> t;
                    t.pack(o);
                    t.store(y + r * ld_y + v * V);
                }
            }
        }
    }
}

// Backward.  Each lane owns the same channel vectors for all rows it visits, so its column partials stay in registers.
template <typename T, int G, int NV>
__global__ void __launch_bounds__(FT, NV == 1 ? 3 : (NV == 2 ? 2 : 1))
ln_bwd_fast_kernel(const T* __restrict__ dy, long long ld_dy, const T* __restrict__ x, long long ld_x,
                   const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int act, T* __restrict__ dx, long long ld_dx, long long rows, int C,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
    constexpr int V = VkVec<T>::N;
    constexpr int R = 32 / G;
    constexpr int D = LnRing<NV, 2>::value;
    constexpr int CW = G * NV * V;
    extern __shared__ uint4 ln_smem[];
    uint4* ring = ln_smem;                                                  // [D][2 * NV][FT]: x vectors, dy vectors
    float* sring = reinterpret_cast<float*>(ln_smem + D * 2 * NV * FT);     // [D][FT / 32][R][2]: mean, rstd
    float* s_gamma = sring + D * (FT / 32) * R * 2;                         // [CW]
    float* s_beta = s_gamma + CW;
    float* sacc = s_beta + CW;                                              // [3][CW]
    const int nvec = C / V;
    for (int i = threadIdx.x; i < CW; i += FT) {
        s_gamma[i] = i < C ? gamma[i] : 0.f;
        s_beta[i] = i < C ? beta[i] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * CW; i += FT) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane & (G - 1), rsub = lane / G;
    const long long it0 = (long long)blockIdx.x * (FT / 32) + wib;
    const long long nit = (long long)gridDim.x * (FT / 32);
    const long long iters = (rows + R - 1) / R;
    const float invC = 1.f / C;

    float gm[NV][V], ag[NV][V], ab[NV][V], ax[NV][V];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            gm[j][i] = s_gamma[(g + G * j) * V + i];
            ag[j][i] = ab[j][i] = ax[j][i] = 0.f;
        }

    auto issue = [&](long long it, int slot) {
        const long long r = it * R + rsub;
        if (it < iters && r < rows) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = g + G * j;
                if (v < nvec) {
                    vk_cp_async16(ring + (slot * 2 * NV + j) * FT + threadIdx.x, x + r * ld_x + v * V);
                    vk_cp_async16(ring + (slot * 2 * NV + NV + j) * FT + threadIdx.x, dy + r * ld_dy + v * V);
                }
            }
            if (g < 2) vk_cp_async4(sring + ((slot * (FT / 32) + wib) * R + rsub) * 2 + g, (g == 0 ? mean : rstd) + r);
        }
        vk_cp_async_commit();
    };
#pragma unroll 1
    for (int d = 0; d < D; ++d) issue(it0 + d * nit, d);
    int slot = 0;
    for (long long it = it0; it < iters; it += nit) {
        const long long r = it * R + rsub;
        const bool row_ok = r < rows;
        float xh[NV][V], dz[NV][V];
        vk_cp_async_wait<D - 1>();
        if (G > 1) __syncwarp();            // mean / rstd were copied by the group's lanes 0 and 1
        const float* sr = sring + ((slot * (FT / 32) + wib) * R + rsub) * 2;
        const float mu = row_ok ? sr[0] : 0.f, rs = row_ok ? sr[1] : 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int v = g + G * j;
            if (row_ok && v < nvec) {
                VkVec<T> tx, td;
                tx.raw = *reinterpret_cast<const decltype(tx.raw)*>(ring + (slot * 2 * NV + j) * FT + threadIdx.x);
                td.raw = *reinterpret_cast<const decltype(td.raw)*>(ring + (slot * 2 * NV + NV + j) * FT + threadIdx.x);
                tx.unpack(xh[j]);
                td.unpack(dz[j]);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) { xh[j][i] = mu; dz[j][i] = 0.f; }   // xhat = 0, dz = 0: contributes nothing
            }
        }
        if (G > 1) __syncwarp();            // every lane has read the slot's statistics before the refill
        issue(it + (long long)D * nit, slot);
        slot = (slot + 1 == D) ? 0 : slot + 1;
        const float shift = -mu * rs;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (act) {      // LN -> GELU blocks of the necks: gelu'(z) on packed pairs (vk_gelu_both2: one MUFU.RCP + one MUFU.EX2 per element)
#pragma unroll
                for (int i = 0; i < V; i += 2) {
                    const float h0 = fmaf(xh[j][i], rs, shift), h1 = fmaf(xh[j][i + 1], rs, shift);
                    const float2 z = make_float2(fmaf(h0, gm[j][i], s_beta[(g + G * j) * V + i]), fmaf(h1, gm[j][i + 1], s_beta[(g + G * j) * V + i + 1]));
                    float2 gv, dg;
                    vk_gelu_both2(z, &gv, &dg);
                    dz[j][i] *= dg.x;
                    dz[j][i + 1] *= dg.y;
                }
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float h = fmaf(xh[j][i], rs, shift);
                float d = dz[j][i];
                xh[j][i] = h;
                ag[j][i] = fmaf(d, h, ag[j][i]);
                ab[j][i] += d;
                const float dxh = d * gm[j][i];
                dz[j][i] = dxh;
                s1 += dxh;
                s2 = fmaf(dxh, h, s2);
            }
        }
        const float2 ss = group_sum2<G>(s1, s2);
        const float m1 = ss.x * invC, m2 = ss.y * invC;
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = g + G * j;
                if (v < nvec) {
                    float o[V];
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        o[i] = rs * (dz[j][i] - m1 - xh[j][i] * m2);
                        ax[j][i] += o[i];
                    }
                    VkVec<T> t;
                    t.pack(o);
                    t.store(dx + r * ld_dx + v * V);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int v = g + G * j;
        if (v < nvec) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                atomicAdd(&sacc[v * V + i], ag[j][i]);
                atomicAdd(&sacc[CW + v * V + i], ab[j][i]);
                atomicAdd(&sacc[2 * CW + v * V + i], ax[j][i]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += FT) {
        if (dgamma) atomicAdd(dgamma + i, sacc[i]);
        if (dbeta) atomicAdd(dbeta + i, sacc[CW + i]);
        if (dxsum) atomicAdd(dxsum + i, sacc[2 * CW + i]);
    }
}

// (G, NV) for nvec 16-byte vectors per row: the fewest lanes per row with at most `max_nv` vectors per lane
inline bool pick_group(int nvec, int max_nv, int* G, int* NV) {
    for (int g = 4; g <= 32; g <<= 1) {
        const int nv = (nvec + g - 1) / g;
        if (nv <= max_nv) { *G = g; *NV = nv; return true; }
    }
    return false;
}

// out[c] += sum_r scale[r / rows_per_group] * x[r, c]; optionally also writes y[r, c] = scale * x[r, c].
// One thread = one 16-byte channel vector, rows strided over the grid; per-block merge in shared memory.
template <typename T>
__global__ void __launch_bounds__(FT)
colsum_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, long long rows, int C,
              const float* __restrict__ scale, int rows_per_group, float* __restrict__ out) {
    constexpr int V = VkVec<T>::N;
    extern __shared__ float cs_acc[];   // [nvec * V]
    const int nvec = C / V;
    for (int i = threadIdx.x; i < C; i += FT) cs_acc[i] = 0.f;
    __syncthreads();
    // threads of a block cover `rpb` consecutive rows x nvec vectors (rpb = FT / nvec, at least 1 vector column set)
    const int rpb = FT / nvec > 0 ? FT / nvec : 1;
    const int v = threadIdx.x % nvec, rloc = threadIdx.x / nvec;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    if (rloc < rpb && nvec <= FT) {
        constexpr int U = 4;   // rows in flight per thread
        const long long stride = (long long)gridDim.x * rpb;
        for (long long r0 = (long long)blockIdx.x * rpb + rloc; r0 < rows; r0 += stride * U) {
            VkVec<T> t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * stride;
                if (r < rows) t[u].load(x + r * ld_x + v * V);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * stride;
                if (r < rows) {
                    float f[V];
                    t[u].unpack(f);
                    if (scale) {
                        const float sc = __ldg(scale + r / rows_per_group);
#pragma unroll
                        for (int i = 0; i < V; ++i) f[i] *= sc;
                        if (y) {
                            VkVec<T> o;
                            o.pack(f);
                            o.store(y + r * ld_y + v * V);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < V; ++i) acc[i] += f[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) atomicAdd(&cs_acc[v * V + i], acc[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += FT) atomicAdd(out + i, cs_acc[i]);
}

}  // namespace

namespace {

#define VK_LN_GROUPS(MACRO)                                                                    \
    if (G == 4 && NV == 1) { MACRO(4, 1); } else if (G == 4 && NV == 2) { MACRO(4, 2); } else if (G == 4 && NV == 3) { MACRO(4, 3); } \
    else if (G == 8 && NV == 1) { MACRO(8, 1); } else if (G == 8 && NV == 2) { MACRO(8, 2); } else if (G == 8 && NV == 3) { MACRO(8, 3); } \
    else if (G == 16 && NV == 1) { MACRO(16, 1); } else if (G == 16 && NV == 2) { MACRO(16, 2); } else if (G == 16 && NV == 3) { MACRO(16, 3); } \
    else if (G == 32 && NV == 1) { MACRO(32, 1); } else if (G == 32 && NV == 2) { MACRO(32, 2); } else { MACRO(32, 3); }

template <typename T>
bool launch_ln_fwd_fast(const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C, const float* gamma,
                        const float* beta, float eps, int act, float* mean, float* rstd, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    if (C % V || ld_x % V || ld_y % V || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
    int G, NV;
    if (!pick_group(C / V, 3, &G, &NV)) return false;
    const int R = 32 / G;
    const long long iters = (rows + R - 1) / R;
    long long blocks = (iters + FT / 32 - 1) / (FT / 32);
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
#define VK_LN_FWD(GG, NN)                                                                                                    \
    do {                                                                                                                     \
        const size_t smem = (size_t)LnRing<NN, 1>::value * NN * FT * 16 + (size_t)2 * GG * NN * V * sizeof(float);          \
        cudaFuncSetAttribute(ln_fwd_kernel<T, GG, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);              \
        ln_fwd_kernel<T, GG, NN><<<(unsigned)blocks, FT, smem, s>>>(reinterpret_cast<const T*>(x), ld_x, reinterpret_cast<T*>(y), \
                                                                     ld_y, rows, C, gamma, beta, eps, act, mean, rstd);      \
    } while (0)
    VK_LN_GROUPS(VK_LN_FWD)
#undef VK_LN_FWD
    return true;
}

template <typename T>
bool launch_ln_bwd_fast(const void* dy, long long ld_dy, const void* x, long long ld_x, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, int act, void* dx, long long ld_dx, long long rows, int C,
                        float* dgamma, float* dbeta, float* dxsum, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    if (C % V || ld_x % V || ld_dy % V || ld_dx % V || (reinterpret_cast<uintptr_t>(x) & 15) ||
        (reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15))
        return false;
    int G, NV;
    // fewest vectors per lane that still fits a 32-lane group: measured fastest for every ConvNeXt width (B200, 640^2)
    if (!pick_group(C / V, 1, &G, &NV) && !pick_group(C / V, 2, &G, &NV) && !pick_group(C / V, 3, &G, &NV)) return false;
    const int R = 32 / G;
    const long long iters = (rows + R - 1) / R;
    long long blocks = (iters + FT / 32 - 1) / (FT / 32);
    long long cap = (long long)vkocr_sm_count() * 4;       // replaced below by the resident block count of the instantiation
    // small maps: every block ends with 3 C shared -> global atomics and starts with a parameter / accumulator prologue, so a
    // warp should see at least ~8 row-iterations (12 800 x 768: 592 -> 200 blocks, 0.059 -> 0.033 ms)
    const long long by_work = iters / ((FT / 32) * 8);
#define VK_LN_BWD_FAST(GG, NN)                                                                                              \
    do {                                                                                                                     \
        constexpr int D = LnRing<NN, 2>::value;                                                                              \
        const size_t smem = (size_t)D * 2 * NN * FT * 16 + ((size_t)D * (FT / 32) * (32 / GG) * 2 + (size_t)5 * GG * NN * V) * sizeof(float); \
        cudaFuncSetAttribute(ln_bwd_fast_kernel<T, GG, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);         \
        static int per_sm = 0;          /* one wave: a grid-stride block beyond the resident ones is a whole second wave */  \
        if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ln_bwd_fast_kernel<T, GG, NN>, FT, smem) != cudaSuccess || per_sm < 1)) per_sm = 1; \
        cap = (long long)vkocr_sm_count() * per_sm;                                                                          \
        if (by_work < cap) cap = by_work > vkocr_sm_count() ? by_work : vkocr_sm_count();                                   \
        if (blocks > cap) blocks = cap;                                                                                      \
        ln_bwd_fast_kernel<T, GG, NN><<<(unsigned)blocks, FT, smem, s>>>(                                                    \
            reinterpret_cast<const T*>(dy), ld_dy, reinterpret_cast<const T*>(x), ld_x, mean, rstd, gamma, beta, act,       \
            reinterpret_cast<T*>(dx), ld_dx, rows, C, dgamma, dbeta, dxsum);                                                 \
    } while (0)
    VK_LN_GROUPS(VK_LN_BWD_FAST)
#undef VK_LN_BWD_FAST
    return true;
}

template <typename T>
bool launch_colsum_fast(const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C, const float* scale,
                        int rows_per_group, float* out, cudaStream_t s) {
    constexpr int V = VkVec<T>::N;
    if (C % V || ld_x % V || (y && ld_y % V) || C / V > FT || (reinterpret_cast<uintptr_t>(x) & 15) ||
        (y && (reinterpret_cast<uintptr_t>(y) & 15)))
        return false;
    const int rpb = FT / (C / V);
    long long blocks = (rows + rpb - 1) / rpb;
    const long long cap = (long long)vkocr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    colsum_kernel<T><<<(unsigned)blocks, FT, (size_t)C * sizeof(float), s>>>(reinterpret_cast<const T*>(x), ld_x,
                                                                              reinterpret_cast<T*>(y), ld_y, rows, C, scale,
                                                                              rows_per_group, out);
    return true;
}

}  // namespace

extern "C" {

// y = LN(x) (optionally GELU(LN(x))).  mean/rstd (fp32 [rows]) are written when non-null (needed by the backward).
int vkocr_layernorm_fwd(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C,
                        const float* gamma, const float* beta, float eps, int act, float* mean, float* rstd, void* stream) {
    VK_REQUIRE(x && y && gamma && beta, VKOCR_BAD_ARGUMENT, "layernorm_fwd: null argument");
    VK_REQUIRE(C >= 1 && ld_x >= C && ld_y >= C, VKOCR_BAD_SHAPE, "layernorm_fwd: C %d ld %lld %lld", C, ld_x, ld_y);
    VK_REQUIRE((mean == nullptr) == (rstd == nullptr), VKOCR_BAD_ARGUMENT, "layernorm_fwd: mean/rstd must come together");
    if (rows == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bool done = false;
    VK_DISPATCH_DTYPE(dtype, T, (done = launch_ln_fwd_fast<T>(x, ld_x, y, ld_y, rows, C, gamma, beta, eps, act, mean, rstd, s)));
    if (!done) {
        long long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
        const long long cap = (long long)vkocr_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        VK_DISPATCH_DTYPE(dtype, T, (ln_fwd_generic_kernel<T><<<(unsigned)blocks, LN_THREADS, 0, s>>>(
                                        reinterpret_cast<const T*>(x), ld_x, reinterpret_cast<T*>(y), ld_y, rows, C, gamma, beta,
                                        eps, act, mean, rstd)));
    }
    VK_CHECK_LAUNCH("ln_fwd_kernel");
    return VKOCR_OK;
}

// dx = d/dx [act(LN(x))] . dy ; dgamma/dbeta/dxsum (fp32 [C], any may be null) are accumulated into.
int vkocr_layernorm_bwd(int dtype, const void* dy, long long ld_dy, const void* x, long long ld_x, const float* mean,
                        const float* rstd, const float* gamma, const float* beta, int act, void* dx, long long ld_dx,
                        long long rows, int C, float* dgamma, float* dbeta, float* dxsum, void* stream) {
    VK_REQUIRE(dy && x && mean && rstd && gamma && beta && dx, VKOCR_BAD_ARGUMENT, "layernorm_bwd: null argument");
    VK_REQUIRE(C >= 1 && C <= 4096, VKOCR_BAD_SHAPE, "layernorm_bwd: C %d", C);
    if (rows == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bool done = false;
    VK_DISPATCH_DTYPE(dtype, T, (done = launch_ln_bwd_fast<T>(dy, ld_dy, x, ld_x, mean, rstd, gamma, beta, act, dx, ld_dx, rows, C,
                                                               dgamma, dbeta, dxsum, s)));
    if (!done)
        VK_DISPATCH_DTYPE(dtype, T, (launch_ln_bwd<T>(dy, ld_dy, x, ld_x, mean, rstd, gamma, beta, act, dx, ld_dx, rows, C, dgamma,
                                                      dbeta, dxsum, s)));
    VK_CHECK_LAUNCH("ln_bwd_kernel");
    return VKOCR_OK;
}

// out[c] += sum_r x[r, c]
int vkocr_colsum(int dtype, const void* x, long long ld, long long rows, int C, float* out, void* stream) {
    VK_REQUIRE(x && out, VKOCR_BAD_ARGUMENT, "colsum: null argument");
    if (rows == 0 || C == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bool done = false;
    VK_DISPATCH_DTYPE(dtype, T, (done = launch_colsum_fast<T>(x, ld, nullptr, 0, rows, C, nullptr, 1, out, s)));
    if (!done) {
        dim3 grid((unsigned)vk_cdiv(C, 32), 1);
        long long gy = (rows + 63) / 64;
        const long long cap = ((long long)vkocr_sm_count() * 8 + grid.x - 1) / grid.x;
        if (gy > cap) gy = cap;
        grid.y = (unsigned)(gy < 1 ? 1 : gy);
        VK_DISPATCH_DTYPE(dtype, T, (colsum_generic_kernel<T><<<grid, 256, 0, s>>>(reinterpret_cast<const T*>(x), ld, rows, C, out)));
    }
    VK_CHECK_LAUNCH("colsum_kernel");
    return VKOCR_OK;
}

// y[r, c] = scale[r / rows_per_group] * x[r, c] and out[c] += sum_r y[r, c]  (stochastic-depth mask applied to the
// incoming gradient of a ConvNeXt layer, convnext.py:41-53, fused with the bias-gradient column sum)
int vkocr_scale_rows_colsum(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C,
                            const float* scale, int rows_per_group, float* out, void* stream) {
    VK_REQUIRE(x && scale && out && rows_per_group > 0, VKOCR_BAD_ARGUMENT, "scale_rows_colsum: bad argument");   // y may be null: sums only
    if (rows == 0 || C == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bool done = false;
    VK_DISPATCH_DTYPE(dtype, T, (done = launch_colsum_fast<T>(x, ld_x, y, ld_y, rows, C, scale, rows_per_group, out, s)));
    VK_REQUIRE(done, VKOCR_BAD_ALIGN, "scale_rows_colsum: C %d / strides must be multiples of 16 bytes and C <= %d vectors", C, FT);
    VK_CHECK_LAUNCH("colsum_kernel");
    return VKOCR_OK;
}

}  // extern "C"
