// Parameter staging between the reference state_dict layout (fp32 master weights: Linear (out,in), Conv2d OIHW,
// depthwise (C,1,7,7)) and the operand layouts of the kernels, plus small gradient finalisers.
#include "common.cuh"

namespace {

// out[r*o_row + t*o_tap + c] = w[r*s_row + tt*s_tap + c*s_col] * (col_scale ? col_scale[c] : 1),  tt = flip ? taps-1-t : t,
// for r < rows, t < taps, c < cols.  Padding (rows/columns the GEMM tiles expect beyond the logical extents) is the
// caller's zero-initialised buffer; this kernel only writes the logical region.
template <typename T>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, long long s_row, long long s_tap, long long s_col, int rows, int taps, int cols,
                   int flip, const float* __restrict__ col_scale, T* __restrict__ out, long long o_row, long long o_tap) {
    const long long total = (long long)rows * taps * cols;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cols);
    const int t = (int)((idx / cols) % taps);
    const int r = (int)(idx / ((long long)cols * taps));
    const int tt = flip ? taps - 1 - t : t;
    float v = w[r * s_row + tt * s_tap + c * s_col];
    if (col_scale) v *= col_scale[c];
    out[r * o_row + t * o_tap + c] = vk_from_f32<T>(v);
}

// y[n*s_n + t*s_t + c*s_c] += g[(n*T + t)*C + c]   (weight gradient from GEMM order back to the parameter's layout)
__global__ void __launch_bounds__(256)
unpack_grad_kernel(const float* __restrict__ g, int N, int T, int C, float* __restrict__ y, long long s_n, long long s_t,
                   long long s_c) {
    const long long total = (long long)N * T * C;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const int t = (int)((idx / C) % T);
    const long long n = idx / ((long long)C * T);
    y[n * s_n + t * s_t + c * s_c] += g[idx];
}

// y[i0*y0 + i1*y1 + i2*y2] += g[i0*g0 + i1*g1 + i2*g2]: a weight gradient from any GEMM order into the parameter's layout
__global__ void __launch_bounds__(256)
scatter_add_kernel(const float* __restrict__ g, long long g0, long long g1, long long g2, int n0, int n1, int n2,
                   float* __restrict__ y, long long y0, long long y1, long long y2) {
    const long long total = (long long)n0 * n1 * n2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int i2 = (int)(idx % n2);
    const int i1 = (int)((idx / n2) % n1);
    const long long i0 = idx / ((long long)n2 * n1);
    y[i0 * y0 + i1 * y1 + i2 * y2] += g[i0 * g0 + i1 * g1 + i2 * g2];
}

// x[row, col0] = first, x[row, col0 + 1 .. col0 + ncols) = 0: the constant "ones" channel appended to a GEMM operand so that
// the product also delivers a bias gradient (ops.ConvNextLayerFn)
template <typename T>
__global__ void __launch_bounds__(256)
set_columns_kernel(T* __restrict__ x, long long ld, long long rows, int col0, int ncols, float first) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ncols) return;
    const int c = (int)(idx % ncols);
    const long long r = idx / ncols;
    x[r * ld + col0 + c] = vk_from_f32<T>(c == 0 ? first : 0.f);
}

// p[0, head) bytes, then `vecs` 16-byte vectors, then `tail` bytes: all zero
__global__ void __launch_bounds__(256)
zero_kernel(uint8_t* __restrict__ p, long long head, long long vecs, long long tail) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    uint4* v = reinterpret_cast<uint4*>(p + head);
    for (long long i = tid; i < vecs; i += stride) v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < head) p[tid] = 0;
    if (tid < tail) p[head + vecs * 16 + tid] = 0;
}

// ConvNeXt MLP second Linear: S[c,k] = s_scale * sum_p dY[p,c] G[p,k] (G zeroed for dropped samples, s_scale = 1 / p_keep),
// sU[c] = sum_p m_b(p) dY[p,c] (the mask channel's column of the same product)
//   dW2[c,k] += gamma[c] * S[c,k];  dgamma[c] += sum_k W2[c,k] S[c,k] + b2[c] sU[c];  db2[c] += gamma[c] sU[c]
// (out = x + gamma * (G W2^T + b2), convnext.py:35-38,56-58)
__global__ void __launch_bounds__(256)
mlp2_grad_finalize_kernel(const float* __restrict__ S, long long ld_s, float s_scale, const float* __restrict__ sU, long long ld_su,
                          const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ gamma, int C, int K,
                          float* __restrict__ dW2, float* __restrict__ dgamma, float* __restrict__ db2) {
    __shared__ float scratch[33];
    const int c = blockIdx.x;
    const float g = gamma[c];
    float dot = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float s = S[(long long)c * ld_s + k] * s_scale;
        dW2[(long long)c * K + k] += g * s;
        dot = fmaf(W2[(long long)c * K + k], s, dot);
    }
    dot = vk_block_sum(dot, scratch);
    if (threadIdx.x == 0) {
        const float su = sU[(long long)c * ld_su];
        dgamma[c] += dot + b2[c] * su;
        db2[c] += g * su;
    }
}

// y[i] += a[i] * s[i % n_s]  (bias-gradient style finalisers)
__global__ void axpy_kernel(const float* __restrict__ a, float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a[i];
}

// x[row, c] *= scale[row / rows_per_group]   (stochastic-depth mask applied to a gradient, convnext.py:41-53)
template <typename T>
__global__ void __launch_bounds__(256)
scale_rows_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, long long rows, int C,
                  const float* __restrict__ scale, int rows_per_group) {
    const long long total = rows * C;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const long long r = idx / C;
    y[r * ld_y + c] = vk_from_f32<T>(vk_to_f32(x[r * ld_x + c]) * scale[r / rows_per_group]);
}

}  // namespace

extern "C" {

int vkocr_pack_weight(const float* w, long long s_row, long long s_tap, long long s_col, int rows, int taps, int cols, int flip,
                      const float* col_scale, void* out, int out_dtype, long long o_row, long long o_tap, void* stream) {
    VK_REQUIRE(w && out, VKOCR_BAD_ARGUMENT, "pack_weight: null argument");
    VK_REQUIRE(taps >= 1, VKOCR_BAD_SHAPE, "pack_weight: taps %d", taps);
    const long long total = (long long)rows * taps * cols;
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(out_dtype, T, (pack_weight_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                        w, s_row, s_tap, s_col, rows, taps, cols, flip, col_scale, reinterpret_cast<T*>(out), o_row,
                                        o_tap)));
    VK_CHECK_LAUNCH("pack_weight_kernel");
    return VKOCR_OK;
}

int vkocr_mlp2_grad_finalize(const float* S, long long ld_s, float s_scale, const float* sU, long long ld_su, const float* W2,
                             const float* b2, const float* gamma, int C, int K, float* dW2, float* dgamma, float* db2, void* stream) {
    VK_REQUIRE(S && sU && W2 && b2 && gamma && dW2 && dgamma && db2, VKOCR_BAD_ARGUMENT, "mlp2_grad_finalize: null argument");
    if (C == 0) return VKOCR_OK;
    mlp2_grad_finalize_kernel<<<C, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(S, ld_s, s_scale, sU, ld_su, W2, b2, gamma, C, K, dW2,
                                                                                      dgamma, db2);
    VK_CHECK_LAUNCH("mlp2_grad_finalize_kernel");
    return VKOCR_OK;
}

int vkocr_unpack_grad(const float* g, int N, int T, int C, float* y, long long s_n, long long s_t, long long s_c, void* stream) {
    VK_REQUIRE(g && y, VKOCR_BAD_ARGUMENT, "unpack_grad: null argument");
    const long long total = (long long)N * T * C;
    if (total == 0) return VKOCR_OK;
    unpack_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, N, T, C, y, s_n, s_t, s_c);
    VK_CHECK_LAUNCH("unpack_grad_kernel");
    return VKOCR_OK;
}

int vkocr_scatter_add_f32(const float* g, long long g0, long long g1, long long g2, int n0, int n1, int n2, float* y, long long y0,
                          long long y1, long long y2, void* stream) {
    VK_REQUIRE(g && y, VKOCR_BAD_ARGUMENT, "scatter_add_f32: null argument");
    const long long total = (long long)n0 * n1 * n2;
    if (total == 0) return VKOCR_OK;
    scatter_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, g0, g1, g2, n0, n1, n2, y,
                                                                                                             y0, y1, y2);
    VK_CHECK_LAUNCH("scatter_add_kernel");
    return VKOCR_OK;
}

int vkocr_set_columns(int dtype, void* x, long long ld, long long rows, int col0, int ncols, float first, void* stream) {
    VK_REQUIRE(x && ncols >= 1 && col0 >= 0 && ld >= col0 + ncols, VKOCR_BAD_ARGUMENT, "set_columns: bad argument");
    const long long total = rows * ncols;
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (set_columns_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<T*>(x), ld, rows, col0, ncols, first)));
    VK_CHECK_LAUNCH("set_columns_kernel");
    return VKOCR_OK;
}

// Zero fill of caller memory on `stream` (atomic-accumulate workspaces, gradient buckets).  A kernel, not cudaMemsetAsync:
// the driver memset measured 0.8 ms slower per training step (~170 fills, five of them 6..57 MB).
int vkocr_zero(void* p, long long bytes, void* stream) {
    VK_REQUIRE(p || bytes == 0, VKOCR_BAD_ARGUMENT, "zero: null argument");
    if (bytes <= 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* b = reinterpret_cast<uint8_t*>(p);
    const long long head = ((16 - (reinterpret_cast<uintptr_t>(b) & 15)) & 15) < bytes ? ((16 - (reinterpret_cast<uintptr_t>(b) & 15)) & 15) : bytes;
    const long long vecs = (bytes - head) / 16;
    const long long tail = bytes - head - vecs * 16;
    long long blocks = (vecs + 255) / 256;
    const long long cap = (long long)vkocr_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    zero_kernel<<<(unsigned)blocks, 256, 0, s>>>(b, head, vecs, tail);
    VK_CHECK_LAUNCH("zero_kernel");
    return VKOCR_OK;
}

// y += a  (fp32 vectors; gradient accumulation into .grad)
int vkocr_accumulate_f32(const float* a, float* y, long long n, void* stream) {
    VK_REQUIRE(a && y, VKOCR_BAD_ARGUMENT, "accumulate_f32: null argument");
    if (n == 0) return VKOCR_OK;
    axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, y, n);
    VK_CHECK_LAUNCH("axpy_kernel");
    return VKOCR_OK;
}

int vkocr_scale_rows(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C, const float* scale,
                     int rows_per_group, void* stream) {
    VK_REQUIRE(x && y && scale && rows_per_group > 0, VKOCR_BAD_ARGUMENT, "scale_rows: bad argument");
    const long long total = rows * C;
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (scale_rows_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, reinterpret_cast<T*>(y), ld_y, rows, C, scale, rows_per_group)));
    VK_CHECK_LAUNCH("scale_rows_kernel");
    return VKOCR_OK;
}

}  // extern "C"
