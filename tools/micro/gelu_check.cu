// Device check: vk_gelu / vk_gelu_both / vk_gelu_both2 against double-precision erf over a sweep of inputs.
#include "../../vkit_ocr_model_adaptive_scaling_b200/csrc/common.cuh"
#include <cstdio>
#include <cmath>
#include <vector>
__global__ void k(const float* x, float* g1, float* d1, float* g2, float* d2, float* g0, int n) {
    int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 >= n) return;
    vk_gelu_both(x[i], &g1[i], &d1[i]);
    vk_gelu_both(x[i + 1], &g1[i + 1], &d1[i + 1]);
    float2 g, d;
    vk_gelu_both2(make_float2(x[i], x[i + 1]), &g, &d);
    g2[i] = g.x; g2[i + 1] = g.y; d2[i] = d.x; d2[i + 1] = d.y;
    g0[i] = vk_gelu(x[i]); g0[i + 1] = vk_gelu(x[i + 1]);
}
int main() {
    const int n = 1 << 20;
    std::vector<float> hx(n);
    for (int i = 0; i < n; ++i) hx[i] = -12.f + 24.f * i / (n - 1);
    float *x, *g1, *d1, *g2, *d2, *g0;
    cudaMalloc(&x, n * 4); cudaMalloc(&g1, n * 4); cudaMalloc(&d1, n * 4); cudaMalloc(&g2, n * 4); cudaMalloc(&d2, n * 4); cudaMalloc(&g0, n * 4);
    cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 2 / 256, 256>>>(x, g1, d1, g2, d2, g0, n);
    std::vector<float> a(n), b(n), c(n), d(n), e(n);
    cudaMemcpy(a.data(), g1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), d1, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c.data(), g2, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(d.data(), d2, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(e.data(), g0, n * 4, cudaMemcpyDeviceToHost);
    double m[5] = {0, 0, 0, 0, 0}; float at[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n - 1; ++i) {
        const double xx = hx[i], cdf = 0.5 * (1 + erf(xx / sqrt(2.0))), pdf = exp(-xx * xx / 2) / sqrt(2 * M_PI);
        const double ge = xx * cdf, de = cdf + xx * pdf;
        const double errs[5] = {fabs(a[i] - ge), fabs(b[i] - de), fabs(c[i] - ge), fabs(d[i] - de), fabs(e[i] - ge)};
        for (int j = 0; j < 5; ++j) if (errs[j] > m[j]) { m[j] = errs[j]; at[j] = hx[i]; }
    }
    printf("scalar both: gelu %.3e (x=%g)  gelu' %.3e (x=%g)\npacked both: gelu %.3e (x=%g)  gelu' %.3e (x=%g)\nfwd-only gelu: %.3e (x=%g)\n", m[0], at[0], m[1], at[1], m[2], at[2], m[3], at[3], m[4], at[4]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
