// Geometry + epilogue description shared by the tcgen05 GEMM (gemm_tc.cu) and the SIMT fp32 GEMM (gemm_simt.cu).
//
// Both kernels compute the same two contractions over NHWC ("pixel-major, channel-contiguous") activations:
//
//   NT  (forward / data-gradient):  D[m, n]      = sum_{tap, c} X[pix(m) + off(tap), c] * Wp[n, tap, c]
//   TN  (weight-gradient):          G[tap, i, j] = sum_{pix}    P[pix, i] * Q[pix + off(tap), j]
//
// ks == 1 degenerates to a plain row-major GEMM (Linear / 1x1 conv / patchified conv).  ks == 3/5 is the
// zero-padded 'same' convolution of reference helper.py:25-40 expressed as an implicit GEMM.
#pragma once
#include "common.cuh"

struct VkocrConvGeom {
    int batch, H, W;       // pixel grid of the activation operand(s); plain GEMM: batch=1, H=1, W=rows
    int ks;                // square kernel size (1, 3, 5); padding = ks/2
    int C;                 // channels contracted per tap (NT) / channels of P (TN)
    long long ld_x;        // pixel stride (elements) of X (NT) / of P (TN)
    int c_pad;             // NT: per-tap K extent of the packed weight (multiple of 64 for the tcgen05 path)
};

struct VkocrEpilogue {
    void* out;             // [rows, ldo] storage dtype (or fp32 when out_f32 != 0)
    long long ldo;
    int out_f32;           // 1: `out` is fp32 regardless of the activation dtype
    int accumulate;        // 1: out += value (fp32 atomics; requires out_f32)
    void* out_pre;         // optional second output, storage dtype: (acc + bias) before the activation; act == 3: gelu'(acc + bias) (fp16 bits when the storage is 16-bit)
    long long ld_pre;
    const float* bias;     // [N] or null
    int act;               // 0 none, 1 exact GELU, 2 multiply by gelu'(aux), 3 exact GELU with out_pre = row_scale * gelu' (and out zeroed where row_scale == 0), 4 multiply by aux
    const float* col_scale;  // [N] or null  (ConvNeXt layer scale, convnext.py:38,56)
    const float* row_scale;  // [rows / rows_per_group] or null (stochastic-depth mask, convnext.py:41-53)
    int rows_per_group;
    const void* residual;  // optional [rows, ld_res] storage dtype, added last (convnext.py:58)
    long long ld_res;
    const void* aux;       // act == 2: value *= gelu'(aux[m, n]); act == 4: value *= aux[m, n] (GELU backward fused into the data-gradient GEMM)
    long long ld_aux;
    // TN only: element (tap, i, j) is accumulated at out[tap*tn_s_tap + i*tn_s_i + j*tn_s_j] (lets the weight gradient
    // land directly in the parameter's own layout, e.g. Conv2d OIHW: s_tap=1, s_i=Cin*taps, s_j=taps)
    long long tn_s_tap, tn_s_i, tn_s_j;
};

// Fused head tail (tcgen05 path only): every N tile of `slot` columns is one head; after the conv bias the epilogue applies
// LayerNorm(inner) -> exact GELU -> Linear(inner -> O <= 4) (-> Softplus) from the fp32 accumulators and writes the NCHW
// fp32 prediction map (UperNextHead.forward upernext.py:233-248 / FpnHead.forward fpn.py:193-208, adaptive_scaling.py:
// 101,140).  The conv output itself is still written to ep.out (storage dtype, needed by the backward) unless ep.out is
// null (inference).
#define VKOCR_MAX_HEADS 4
struct VkocrHeadTail {
    int num_heads;                         // 0: disabled
    int slot;                              // columns per head (multiple of 16, <= 256) == the GEMM's N tile
    long long pixels_per_image;
    const float* gamma[VKOCR_MAX_HEADS];   // LayerNorm weight / bias [inner]
    const float* beta[VKOCR_MAX_HEADS];
    const float* w2[VKOCR_MAX_HEADS];      // [O, inner] row-major
    const float* b2[VKOCR_MAX_HEADS];      // [O]
    float* out[VKOCR_MAX_HEADS];           // [batch, O, H, W] fp32
    int inner[VKOCR_MAX_HEADS];
    int out_channels[VKOCR_MAX_HEADS];
    int softplus[VKOCR_MAX_HEADS];
};

// gelu' side channel (act 3 writes it, act 4 reads it): in 16-bit storage it travels as IEEE fp16 BITS inside the bf16
// buffer -- |gelu'| <= 1.13, and fp16's 11-bit significand keeps the derivative 8x more accurate than a bf16 rounding
// (a bf16-rounded derivative alone raised the end-to-end bf16 gradient error of the TINY model from 1.3e-2 to 2.1e-2).
template <typename T> __device__ __forceinline__ void vk_store_dgelu(T* p, float v);
template <> __device__ __forceinline__ void vk_store_dgelu<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void vk_store_dgelu<__nv_bfloat16>(__nv_bfloat16* p, float v) { *reinterpret_cast<__half*>(p) = __float2half_rn(v); }
template <typename T> __device__ __forceinline__ float vk_load_dgelu(const T* p);
template <> __device__ __forceinline__ float vk_load_dgelu<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float vk_load_dgelu<__nv_bfloat16>(const __nv_bfloat16* p) { return __half2float(*reinterpret_cast<const __half*>(p)); }

// Apply the epilogue to one accumulator value and return the value to store in `out`.
// Side effect: writes out_pre when requested.
template <typename T>
__device__ __forceinline__ float vk_epilogue_value(const VkocrEpilogue& ep, long long m, int n, float acc) {
    float v = acc;
    if (ep.bias) v += __ldg(ep.bias + n);
    if (ep.act == 3) {
        float g, dg;
        vk_gelu_both(v, &g, &dg);
        // with a row scale (stochastic-depth mask m_b in {0, 1/p_keep}) the outputs are prepared for the backward: the
        // derivative side channel carries m_b * gelu', the activation of dropped samples is zeroed (see ConvNextLayerFn)
        const float rs3 = ep.row_scale ? __ldg(ep.row_scale + (m / ep.rows_per_group)) : 1.f;
        if (ep.out_pre) vk_store_dgelu<T>(reinterpret_cast<T*>(ep.out_pre) + m * ep.ld_pre + n, dg * rs3);
        v = rs3 != 0.f ? g : 0.f;
        if (ep.col_scale) v *= __ldg(ep.col_scale + n);
        if (ep.residual) v += vk_to_f32(reinterpret_cast<const T*>(ep.residual)[m * ep.ld_res + n]);
        return v;
    } else {
        if (ep.out_pre) reinterpret_cast<T*>(ep.out_pre)[m * ep.ld_pre + n] = vk_from_f32<T>(v);
        if (ep.act == 1) v = vk_gelu(v);
        else if (ep.act == 2) v *= vk_gelu_grad(vk_to_f32(reinterpret_cast<const T*>(ep.aux)[m * ep.ld_aux + n]));
        else if (ep.act == 4) v *= vk_load_dgelu<T>(reinterpret_cast<const T*>(ep.aux) + m * ep.ld_aux + n);
    }
    if (ep.col_scale) v *= __ldg(ep.col_scale + n);
    if (ep.row_scale) v *= __ldg(ep.row_scale + (m / ep.rows_per_group));
    if (ep.residual) v += vk_to_f32(reinterpret_cast<const T*>(ep.residual)[m * ep.ld_res + n]);
    return v;
}

template <typename T>
__device__ __forceinline__ void vk_epilogue_store(const VkocrEpilogue& ep, long long m, int n, float v) {
    if (ep.out_f32) {
        float* o = reinterpret_cast<float*>(ep.out) + m * ep.ldo + n;
        if (ep.accumulate) atomicAdd(o, v);
        else *o = v;
    } else {
        reinterpret_cast<T*>(ep.out)[m * ep.ldo + n] = vk_from_f32<T>(v);
    }
}

// TN (weight-gradient) store: fp32 atomic accumulate at the caller's strides.
__device__ __forceinline__ void vk_epilogue_store_tn(const VkocrEpilogue& ep, int tap, int i, int j, float v) {
    atomicAdd(reinterpret_cast<float*>(ep.out) + tap * ep.tn_s_tap + i * ep.tn_s_i + j * ep.tn_s_j, v);
}
