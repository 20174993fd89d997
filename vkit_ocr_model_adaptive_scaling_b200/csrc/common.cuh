// Shared helpers for the sm_100a kernels of the adaptive-scaling hot path.
// Everything here is device/host plumbing: error convention of the C ABI, dtype tags,
// 16-byte vector access, warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

// ---- C-ABI error convention (include/vkocr_b200.h) -------------------------------------------
enum VkocrStatus {
    VKOCR_OK = 0,
    VKOCR_BAD_SHAPE = -1,
    VKOCR_BAD_ALIGN = -2,
    VKOCR_UNSUPPORTED_DTYPE = -3,
    VKOCR_CUDA_ERROR = -4,
    VKOCR_WORKSPACE_TOO_SMALL = -5,
    VKOCR_BAD_ARGUMENT = -6,
};

enum VkocrDtype { VKOCR_F32 = 0, VKOCR_BF16 = 1 };

void vkocr_set_error(const char* fmt, ...);
int vkocr_sm_count();
void vkocr_note_launch();   // bumps the process-wide kernel-launch counter (vkocr_launch_count)

#define VK_FAIL(code, ...)               \
    do {                                 \
        vkocr_set_error(__VA_ARGS__);    \
        return (code);                   \
    } while (0)

#define VK_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) VK_FAIL(code, __VA_ARGS__); \
    } while (0)

#define VK_CHECK_LAUNCH(name)                                                        \
    do {                                                                             \
        vkocr_note_launch();                                                         \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess)                                                      \
            VK_FAIL(VKOCR_CUDA_ERROR, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// Dispatch on a storage dtype tag; T is the storage type (math is always fp32).
#define VK_DISPATCH_DTYPE(dtype, T, ...)                                  \
    do {                                                                  \
        if ((dtype) == VKOCR_F32) { using T = float; __VA_ARGS__; }        \
        else if ((dtype) == VKOCR_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else VK_FAIL(VKOCR_UNSUPPORTED_DTYPE, "unsupported dtype %d", (int)(dtype)); \
    } while (0)

static inline int vk_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar conversions -----------------------------------------------------------------------
__device__ __forceinline__ float vk_to_f32(float v) { return v; }
__device__ __forceinline__ float vk_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T vk_from_f32(float v);
template <> __device__ __forceinline__ float vk_from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 vk_from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- vectors of VEC storage elements = 16 bytes -------------------------------------------------
template <typename T> struct VkVec;
template <> struct VkVec<float> {
    static constexpr int N = 4;
    float4 raw;
    __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
    __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
    __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct VkVec<__nv_bfloat16> {
    static constexpr int N = 8;
    uint4 raw;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void unpack(float* f) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 v = __bfloat1622float2(h[i]);
            f[2 * i] = v.x;
            f[2 * i + 1] = v.y;
        }
    }
    __device__ __forceinline__ void pack(const float* f) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
};

// 8-byte vector of 4 bf16 (for register-hungry kernels that want 4 channels per thread in either storage type)
template <typename T> struct VkVec4;
template <> struct VkVec4<float> : VkVec<float> {};
template <> struct VkVec4<__nv_bfloat16> {
    static constexpr int N = 4;
    uint2 raw;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint2*>(p); }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint2*>(p) = raw; }
    __device__ __forceinline__ void unpack(float* f) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float2 v = __bfloat1622float2(h[i]);
            f[2 * i] = v.x;
            f[2 * i + 1] = v.y;
        }
    }
    __device__ __forceinline__ void pack(const float* f) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
};

// One-element "vector" with the VkVec interface: the generic path for channel counts / strides / slice offsets that are
// not 16-byte aligned.
template <typename T> struct VkScalar {
    static constexpr int N = 1;
    T raw;
    __device__ __forceinline__ void load(const T* p) { raw = *p; }
    __device__ __forceinline__ void store(T* p) const { *p = raw; }
    __device__ __forceinline__ void unpack(float* f) const { f[0] = vk_to_f32(raw); }
    __device__ __forceinline__ void pack(const float* f) { raw = vk_from_f32<T>(f[0]); }
};

// ---- per-thread asynchronous global -> shared copies (LDGSTS) ---------------------------------------------------
// A thread that streams rows keeps a ring of its next rows' vectors in shared memory: the copies are in flight while it
// computes, and it reads back only what it copied itself (no cross-thread synchronisation).  One commit group per row.
__device__ __forceinline__ void vk_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void vk_cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void vk_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void vk_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ float vk_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole block; result valid in every thread. `scratch` needs 33 floats.
__device__ __forceinline__ float vk_block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = vk_warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nwarp ? scratch[lane] : 0.f;
        t = vk_warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---- math ------------------------------------------------------------------------------------------
// Exact (erf) GELU of the reference (helper.py:100-101) and its derivative.  erf is evaluated with Abramowitz & Stegun
// 7.1.26 (absolute error <= 1.5e-7, below fp32 round-off of the surrounding math): straight-line code, one MUFU.RCP
// (refined by one Newton step -- __frcp_rn / "1.f / x" compile to a MUFU plus a conditional slow-path CALL per element,
// which serialises these instruction-bound loops), one MUFU.EX2 and a handful of FMAs.  exp(-x^2/2), needed for
// erf(x/sqrt 2), is also the Gaussian density of the derivative, so gelu and gelu' share all of the work.  The lower
// tail is computed directly (0.5*(1-erf|u|) = 0.5*poly*e), not as 1 - erf, so it keeps its relative accuracy.
__device__ __forceinline__ float vk_rcp(float d) {   // MUFU.RCP: relative error ~1 ulp, no slow path
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(d));
    return t;
}
__device__ __forceinline__ float vk_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 16 instructions (18 with the density): the 0.5 of Phi = 0.5 erfc(.) and the 1/sqrt(2) of the argument are folded into
// the constants; +-inf give exactly 0 / 1.
__device__ __forceinline__ void vk_gelu_parts(float x, float* cdf, float* pdf) {
    const float t = vk_rcp(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.f));
    const float e = vk_ex2(x * x * (-0.5f * 1.44269504088896340736f));   // exp(-x^2 / 2)
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float half_tail = poly * t * e;                   // 0.5 * erfc(|x| / sqrt 2), in [0, 0.5]
    *cdf = 0.5f + copysignf(0.5f - half_tail, x);
    *pdf = 0.39894228040143267794f * e;
}
// Forward-only GELU with ONE transcendental: 0.5 erfc(a / sqrt 2) = 2^-(1 + a P(a)) for a = |x|, P a degree-5 polynomial
// fitted on [0, 9] (minimax of the GELU error: |gelu - exact| <= 1.3e-7, |Phi - exact| <= 7e-7 with fp32 Horner), and
// gelu(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2), which keeps the relative accuracy of the lower tail.  10 instructions,
// one MUFU.EX2: the two-MUFU form above is MUFU-pipe-bound (2 x 8 cycles per warp) wherever the GELU sits in a GEMM
// epilogue.  Beyond the fitted range the argument is clamped (erfc(9 / sqrt 2) ~ 2e-19).
__device__ __forceinline__ float vk_gelu(float x) {
    const float a = fminf(fabsf(x), 9.f);
    float q = fmaf(a, -3.159132926264e-05f, 7.531542511152e-04f);
    q = fmaf(q, a, -8.015595779370e-03f);
    q = fmaf(q, a, 5.328616913828e-02f);
    q = fmaf(q, a, 4.588905654064e-01f);
    q = fmaf(q, a, 1.151150930355e+00f);
    q = fmaf(q, a, 1.f);
    return fmaf(-fabsf(x), vk_ex2(-q), fmaxf(x, 0.f));
}
__device__ __forceinline__ float vk_gelu_grad(float x) {
    float cdf, pdf;
    vk_gelu_parts(x, &cdf, &pdf);
    return fmaf(x, pdf, cdf);
}
// gelu(x) and gelu'(x) together
__device__ __forceinline__ void vk_gelu_both(float x, float* g, float* dg) {
    float cdf, pdf;
    vk_gelu_parts(x, &cdf, &pdf);
    *g = x * cdf;
    *dg = fmaf(x, pdf, cdf);
}
// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) ------------------------------------------------------------
// Two fp32 lanes per instruction: the same FMA-pipe time as two scalar instructions but ONE issue slot.  The streaming
// kernels with ~40 fp32 instructions per element are bound by instruction issue, not by the pipes.
__device__ __forceinline__ float2 vk_fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
          "l"(*reinterpret_cast<const unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 vk_mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(r)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 vk_add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(r)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 vk_splat2(float a) { return make_float2(a, a); }
// gelu and gelu' of a pair: vk_gelu_parts with the polynomial / product work packed (the two MUFUs per lane stay scalar)
__device__ __forceinline__ void vk_gelu_both2(float2 x, float2* g, float2* dg) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 den = vk_fma2(vk_splat2(0.3275911f * 0.70710678118654752440f), ax, vk_splat2(1.f));
    const float2 t = make_float2(vk_rcp(den.x), vk_rcp(den.y));
    const float2 arg = vk_mul2(vk_mul2(x, x), vk_splat2(-0.5f * 1.44269504088896340736f));
    const float2 e = make_float2(vk_ex2(arg.x), vk_ex2(arg.y));
    float2 npoly = vk_fma2(vk_splat2(-0.5f * 1.061405429f), t, vk_splat2(-0.5f * -1.453152027f));   // -(polynomial)
    npoly = vk_fma2(npoly, t, vk_splat2(-0.5f * 1.421413741f));
    npoly = vk_fma2(npoly, t, vk_splat2(-0.5f * -0.284496736f));
    npoly = vk_fma2(npoly, t, vk_splat2(-0.5f * 0.254829592f));
    const float2 up = vk_fma2(npoly, vk_mul2(t, e), vk_splat2(0.5f));                                // 0.5 - half_tail
    const float2 cdf = vk_add2(vk_splat2(0.5f), make_float2(copysignf(up.x, x.x), copysignf(up.y, x.y)));
    const float2 pdf = vk_mul2(vk_splat2(0.39894228040143267794f), e);
    *g = vk_mul2(x, cdf);
    *dg = vk_fma2(x, pdf, cdf);
}
// Forward-only exact GELU of a pair: vk_gelu with the polynomial packed (one MUFU.EX2 per element).  A scalar 3-register FFMA
// issues every other cycle per scheduler (tools/micro/ffma2_operands.cu: 68 FMA/clk/SM), the packed form does two lanes in
// the same slot, so GELU-heavy epilogues and tails are written on pairs.
__device__ __forceinline__ float2 vk_gelu2(float2 x) {
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
__device__ __forceinline__ float vk_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }   // per-pixel maps only (IEEE division)
__device__ __forceinline__ float vk_softplus(float x) {  // beta 1, threshold 20 (torch.nn.Softplus)
    return x > 20.f ? x : log1pf(expf(x));
}
