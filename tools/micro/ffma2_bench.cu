// Micro-benchmark: fp32 FMA throughput per SM with scalar FFMA versus packed FFMA2 (fma.rn.f32x2), and with a mix of
// integer work issued alongside (does FFMA2 free issue slots?).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float s) {
    float a[16];
    unsigned x = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long v, sv, cv;
                    float2 t = make_float2(a[i], a[i + 1]), s2 = make_float2(s, s), c2 = make_float2(0.5f, 0.5f);
                    v = *reinterpret_cast<unsigned long long*>(&t);
                    sv = *reinterpret_cast<unsigned long long*>(&s2);
                    cv = *reinterpret_cast<unsigned long long*>(&c2);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(sv), "l"(cv));
                    t = *reinterpret_cast<float2*>(&v);
                    a[i] = t.x; a[i + 1] = t.y;
                }
        }
        if (MODE >= 2) {
            // 32 integer ops per 64 FMAs
#pragma unroll
            for (int i = 0; i < 32; ++i) x = (x << 16) ^ (x & 0xffff0000u) ^ (unsigned)i;
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)x;
}

template <int MODE>
void run(const char* name, float* out, int sms) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<sms * 4, 256>>>(out, 100, 0.999f);
    cudaEventRecord(e0);
    bench<MODE><<<sms * 4, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * 4 * 256 * iters * 64;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s fp32  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.9e9);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, sms * 4 * 256 * sizeof(float));
    run<0>("FFMA", out, sms);
    run<1>("FFMA2", out, sms);
    run<2>("FFMA + 0.5 int op / FMA", out, sms);
    run<3>("FFMA2 + 0.5 int op / FMA", out, sms);
    return 0;
}
