// Micro-probe: does the fp32 FMA rate of sm_100a depend on how many DISTINCT register operands an FFMA / FFMA2 reads?
// tools/micro/ffma2_bench.cu measured 128 FMA/clk/SM with two of three operands shared by every instruction (operand reuse
// cache).  The depthwise 7x7 kernels issue  acc[o] += w * in[o + kx]:  the accumulator and the input differ per instruction,
// the tap is shared by 4.  Patterns:  W1 = one multiplier shared by all chains, W4 = shared by 4 consecutive instructions,
// W0 = nothing shared.   Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/ffma2_operands tools/micro/ffma2_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;"
                 : "+l"(dd)
                 : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void ffma1(float& d, const float a, const float b) {
    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(a), "f"(b));
}

// PACKED: FFMA2 on NCH float2 chains, else FFMA on 2*NCH float chains (same FMA count).  SHARE: 0 / 4 / NCH chains per multiplier.
template <bool PACKED, int SHARE>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ src, float* out, int iters) {
    constexpr int NCH = 16;
    float2 acc[NCH], x[NCH], w[NCH];
    for (int i = 0; i < NCH; ++i) {
        acc[i] = make_float2(0.f, 0.f);
        x[i] = make_float2(src[threadIdx.x + 2 * i], src[threadIdx.x + 2 * i + 1]);
        w[i] = make_float2(src[64 + threadIdx.x + 2 * i], src[65 + threadIdx.x + 2 * i]);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int wi = SHARE == 0 ? i : (SHARE == 4 ? ((i / 4 + rep) % (NCH / 4)) * 4 : rep);
                const int xi = (i + rep) % NCH;
                if (PACKED) ffma2(acc[i], w[wi], x[xi]);
                else { ffma1(acc[i].x, w[wi].x, x[xi].x); ffma1(acc[i].y, w[wi].y, x[xi].y); }
            }
    }
    float r = 0.f;
    for (int i = 0; i < NCH; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <bool PACKED, int SHARE>
void run(const char* name, const float* src, float* out, int sms, int blocks_per_sm) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<PACKED, SHARE><<<sms * blocks_per_sm, 256>>>(src, out, 50);
    cudaEventRecord(e0);
    probe<PACKED, SHARE><<<sms * blocks_per_sm, 256>>>(src, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * blocks_per_sm * 256 * iters * 4 * 16 * 2;
    printf("%-44s %d warps/SM %8.3f ms  %6.1f FMA/clk/SM at 1.9 GHz\n", name, blocks_per_sm * 8, ms, fma / (ms * 1e-3) / sms / 1.9e9);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *src, *out;
    cudaMalloc(&src, 4096);
    cudaMemset(src, 0, 4096);
    cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
    for (int bps = 2; bps <= 4; bps += 2) {
        run<false, 16>("FFMA   multiplier shared by all", src, out, sms, bps);
        run<false, 4>("FFMA   multiplier shared by 4", src, out, sms, bps);
        run<false, 0>("FFMA   three distinct operands", src, out, sms, bps);
        run<true, 16>("FFMA2  multiplier shared by all", src, out, sms, bps);
        run<true, 4>("FFMA2  multiplier shared by 4 (dwconv pattern)", src, out, sms, bps);
        run<true, 0>("FFMA2  three distinct operands", src, out, sms, bps);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
