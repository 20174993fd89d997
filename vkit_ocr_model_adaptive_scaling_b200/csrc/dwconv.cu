// Depthwise 7x7 convolution on NHWC activations (reference helper.dconv7x7, model/helper.py:61-73: groups=C, pad 3,
// stride 1, bias), forward, data-gradient and weight-gradient.
//
// HBM-bound.  Forward/data-gradient: each thread produces a 4-pixel strip x one 16-byte channel vector, sweeping the
// 7 input rows once (10 vector loads per row instead of 28), taps in a [49][C] fp32 table.  The data-gradient is the
// same kernel with the taps mirrored (the packer flips them) and an optional `add` operand that fuses the residual
// gradient of ConvNextBlockLayer (convnext.py:56-58).
#include "common.cuh"

namespace {

constexpr int STRIP = 4;

template <typename T>
__global__ void __launch_bounds__(256)
dwconv7_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, int B, int H, int W, int C,
               const float* __restrict__ wt, const float* __restrict__ bias, const T* __restrict__ add, long long ld_add) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const int XS = (W + STRIP - 1) / STRIP;
    const long long total = (long long)B * H * XS * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int xs = (int)(r % XS);
    r /= XS;
    const int yy0 = (int)(r % H);
    const int b = (int)(r / H);
    const int c = cv * V;
    const int x0 = xs * STRIP;

    float acc[STRIP][V];
#pragma unroll
    for (int o = 0; o < STRIP; ++o)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[o][i] = bias ? __ldg(bias + c + i) : 0.f;

    for (int ky = 0; ky < 7; ++ky) {
        const int yy = yy0 + ky - 3;
        if (yy < 0 || yy >= H) continue;
        const T* row = x + ((long long)b * H + yy) * W * ld_x + c;
        VkVec<T> in[STRIP + 6];
#pragma unroll
        for (int j = 0; j < STRIP + 6; ++j) {
            const int xx = x0 + j - 3;
            if (xx >= 0 && xx < W) in[j].load(row + (long long)xx * ld_x);
            else {
                float z[V] = {};
                in[j].pack(z);
            }
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            float w[V];
#pragma unroll
            for (int i = 0; i < V; i += 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wt + (long long)(ky * 7 + kx) * C + c + i));
                w[i] = w4.x; w[i + 1] = w4.y; w[i + 2] = w4.z; w[i + 3] = w4.w;
            }
#pragma unroll
            for (int o = 0; o < STRIP; ++o) {
                float f[V];
                in[o + kx].unpack(f);
#pragma unroll
                for (int i = 0; i < V; ++i) acc[o][i] = fmaf(w[i], f[i], acc[o][i]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < STRIP; ++o) {
        const int xx = x0 + o;
        if (xx >= W) break;
        const long long pix = ((long long)b * H + yy0) * W + xx;
        if (add) {
            VkVec<T> a;
            a.load(add + pix * ld_add + c);
            float f[V];
            a.unpack(f);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[o][i] += f[i];
        }
        VkVec<T> out;
        out.pack(acc[o]);
        out.store(y + pix * ld_y + c);
    }
}

// dW[c][ky*7+kx] += sum_{b,y,x} dy[b,y,x,c] * x[b,y+ky-3,x+kx-3,c]   (the (C,1,7,7) layout of the parameter)
// One thread = one 16-byte channel vector x one kernel row ky (blockIdx.y) x one SEG-pixel segment of an image row: it
// loads the SEG+6 input vectors of the segment (halo included) and the SEG gradient vectors once and does 7*SEG*V FMAs
// on them (each x vector is reused by up to 7 taps from registers), keeping acc[7 kx][V] in registers over all rows the
// block visits.  Partials are merged per block in shared memory, then one atomic per (channel, tap) and block.
template <typename T, int SEG>
__global__ void __launch_bounds__(256, 2)
dwconv7_wgrad_kernel(const T* __restrict__ dy, long long ld_dy, const T* __restrict__ x, long long ld_x, int B, int H, int W,
                     int C, int cv_per_block, float* __restrict__ dw) {
    using VT = VkVec4<T>;
    constexpr int V = VT::N;
    extern __shared__ float red[];  // [7][cv_per_block * V]
    const int CV = C / V;
    const int lane_cv = threadIdx.x % cv_per_block;
    const int sg = threadIdx.x / cv_per_block;
    const int NSG = blockDim.x / cv_per_block;
    const int cv = blockIdx.z * cv_per_block + lane_cv;
    const int ky = blockIdx.y;
    const bool active = cv < CV && sg < NSG;
    const int c = cv * V;
    const int segs = (W + SEG - 1) / SEG;
    float acc[7][V];
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[k][i] = 0.f;
    for (int i = threadIdx.x; i < 7 * cv_per_block * V; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    if (active) {
        const long long rows = (long long)B * H;
        for (long long rr = blockIdx.x; rr < rows; rr += gridDim.x) {
            const int yy = (int)(rr % H);
            const int b = (int)(rr / H);
            const int ys = yy + ky - 3;
            if (ys < 0 || ys >= H) continue;
            const T* dyrow = dy + ((long long)b * H + yy) * W * ld_dy + c;
            const T* xrow = x + ((long long)b * H + ys) * W * ld_x + c;
            for (int s = sg; s < segs; s += NSG) {
                const int x0 = s * SEG;
                VT in[SEG + 6], gd[SEG];
#pragma unroll
                for (int j = 0; j < SEG + 6; ++j) {
                    const int xx = x0 + j - 3;
                    if (xx >= 0 && xx < W) in[j].load(xrow + (long long)xx * ld_x);
                    else {
                        float z[V] = {};
                        in[j].pack(z);
                    }
                }
#pragma unroll
                for (int o = 0; o < SEG; ++o) {
                    const int xx = x0 + o;
                    if (xx < W) gd[o].load(dyrow + (long long)xx * ld_dy);
                    else {
                        float z[V] = {};
                        gd[o].pack(z);
                    }
                }
#pragma unroll
                for (int o = 0; o < SEG; ++o) {
                    float fd[V];
                    gd[o].unpack(fd);
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {
                        float fx[V];
                        in[o + kx].unpack(fx);
#pragma unroll
                        for (int i = 0; i < V; ++i) acc[kx][i] = fmaf(fd[i], fx[i], acc[kx][i]);
                    }
                }
            }
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
            for (int i = 0; i < V; ++i) atomicAdd(&red[(kx * cv_per_block + lane_cv) * V + i], acc[kx][i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 7 * cv_per_block * V; i += blockDim.x) {
        const int kx = i / (cv_per_block * V);
        const int rem = i % (cv_per_block * V);
        const int cc = blockIdx.z * cv_per_block * V + rem;
        if (cc < C) atomicAdd(dw + (long long)cc * 49 + (ky * 7 + kx), red[i]);
    }
}

}  // namespace

extern "C" {

// y[b,y,x,c] = bias[c] + sum_{ky,kx} wt[ky*7+kx][c] * x[b,y+ky-3,x+kx-3,c]  (+ add[b,y,x,c])
// wt is the [49][C] fp32 tap table written by vkocr_pack_dwconv_weight (mirrored for the data gradient).
int vkocr_dwconv7_fwd(int dtype, const void* x, long long ld_x, void* y, long long ld_y, int B, int H, int W, int C,
                      const float* wt, const float* bias, const void* add, long long ld_add, void* stream) {
    VK_REQUIRE(x && y && wt, VKOCR_BAD_ARGUMENT, "dwconv7: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(C % V == 0 && ld_x % V == 0 && ld_y % V == 0 && (!add || ld_add % V == 0), VKOCR_BAD_ALIGN,
               "dwconv7: C %d / strides must be multiples of %d", C, V);
    const long long total = (long long)B * H * ((W + STRIP - 1) / STRIP) * (C / V);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    VK_DISPATCH_DTYPE(dtype, T, (dwconv7_kernel<T><<<blocks, 256, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, reinterpret_cast<T*>(y), ld_y, B, H, W, C, wt, bias,
                                    reinterpret_cast<const T*>(add), ld_add)));
    VK_CHECK_LAUNCH("dwconv7_kernel");
    return VKOCR_OK;
}

// dw (fp32, parameter layout (C,1,7,7)) += correlation of dy with x.
int vkocr_dwconv7_wgrad(int dtype, const void* dy, long long ld_dy, const void* x, long long ld_x, int B, int H, int W, int C,
                        float* dw, void* stream) {
    VK_REQUIRE(dy && x && dw, VKOCR_BAD_ARGUMENT, "dwconv7_wgrad: null argument");
    const int V = 4;   // 4 channels per thread in either storage type (VkVec4)
    VK_REQUIRE(C % V == 0 && ld_x % V == 0 && ld_dy % V == 0, VKOCR_BAD_ALIGN, "dwconv7_wgrad: C %d / strides must be multiples of %d", C, V);
    if ((long long)B * H * W == 0) return VKOCR_OK;
    constexpr int SEG = 8;
    const int CV = C / V;
    // channel vectors per block: a divisor-friendly width so that (almost) all 256 threads carry a segment
    int cvb = CV < 32 ? CV : 32;
    const int zblocks = (CV + cvb - 1) / cvb;
    const int segs = (W + SEG - 1) / SEG;
    int nsg = 256 / cvb;
    if (nsg > segs) nsg = segs;
    const int threads = nsg * cvb;
    long long rows = (long long)B * H;
    long long gx = ((long long)vkocr_sm_count() * 6 + 7 * zblocks - 1) / (7 * zblocks);
    if (gx > rows) gx = rows;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, 7, (unsigned)zblocks);
    const size_t smem = (size_t)7 * cvb * V * sizeof(float);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (dwconv7_wgrad_kernel<T, SEG><<<grid, threads, smem, s>>>(
                                    reinterpret_cast<const T*>(dy), ld_dy, reinterpret_cast<const T*>(x), ld_x, B, H, W, C, cvb,
                                    dw)));
    VK_CHECK_LAUNCH("dwconv7_wgrad_kernel");
    return VKOCR_OK;
}

}  // extern "C"
