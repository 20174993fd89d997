// Channels-last LayerNorm (+ optional fused exact GELU) forward and backward.
// Reference: helper.ln = nn.LayerNorm(C, eps=1e-6) applied on BHWC (model/helper.py:96-97), followed by GELU in the
// neck/head conv blocks (upernext.py:21-45, fpn.py:21-48).
//
// HBM-bound: one warp per pixel row, 16-byte vector accesses along C, warp-shuffle reductions.  Each row is read
// from HBM once (the second/third sweeps of the ~1-3 KB row hit L1) and written once.
// Backward also produces the three per-channel reductions the caller needs:
//   dgamma += sum_rows dz * xhat,  dbeta += sum_rows dz,  dxsum += sum_rows dx  (= bias gradient of the producer).
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
ln_fwd_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, long long rows, int C,
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
colsum_kernel(const T* __restrict__ x, long long ld, long long rows, int C, float* __restrict__ out) {
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

}  // namespace

extern "C" {

// y = LN(x) (optionally GELU(LN(x))).  mean/rstd (fp32 [rows]) are written when non-null (needed by the backward).
int vkocr_layernorm_fwd(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C,
                        const float* gamma, const float* beta, float eps, int act, float* mean, float* rstd, void* stream) {
    VK_REQUIRE(x && y && gamma && beta, VKOCR_BAD_ARGUMENT, "layernorm_fwd: null argument");
    VK_REQUIRE(C >= 1 && ld_x >= C && ld_y >= C, VKOCR_BAD_SHAPE, "layernorm_fwd: C %d ld %lld %lld", C, ld_x, ld_y);
    VK_REQUIRE((mean == nullptr) == (rstd == nullptr), VKOCR_BAD_ARGUMENT, "layernorm_fwd: mean/rstd must come together");
    if (rows == 0) return VKOCR_OK;
    long long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)vkocr_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (ln_fwd_kernel<T><<<(unsigned)blocks, LN_THREADS, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, reinterpret_cast<T*>(y), ld_y, rows, C, gamma, beta, eps,
                                    act, mean, rstd)));
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
    VK_DISPATCH_DTYPE(dtype, T, (launch_ln_bwd<T>(dy, ld_dy, x, ld_x, mean, rstd, gamma, beta, act, dx, ld_dx, rows, C, dgamma,
                                                  dbeta, dxsum, s)));
    VK_CHECK_LAUNCH("ln_bwd_kernel");
    return VKOCR_OK;
}

// out[c] += sum_r x[r, c]
int vkocr_colsum(int dtype, const void* x, long long ld, long long rows, int C, float* out, void* stream) {
    VK_REQUIRE(x && out, VKOCR_BAD_ARGUMENT, "colsum: null argument");
    if (rows == 0 || C == 0) return VKOCR_OK;
    dim3 grid((unsigned)vk_cdiv(C, 32), 1);
    long long gy = (rows + 63) / 64;
    const long long cap = ((long long)vkocr_sm_count() * 8 + grid.x - 1) / grid.x;
    if (gy > cap) gy = cap;
    grid.y = (unsigned)(gy < 1 ? 1 : gy);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<grid, 256, 0, s>>>(reinterpret_cast<const T*>(x), ld, rows, C, out)));
    VK_CHECK_LAUNCH("colsum_kernel");
    return VKOCR_OK;
}

}  // extern "C"
