// Tensor-side pre/post operations of the rough inference pass, on the device:
//   ingest   uint8 HWC page image(s) -> zero-padded fp32 NCHW network input (pad to a multiple of the backbone stride:
//            inferencing/opt.py:16-41 pad_mat_to_make_divisible; transpose + float: inferencing/adaptive_scaling.py:116-121)
//   postproc mask = sigmoid(logit) >= thr as uint8, height map with the padding region and heights below the minimum
//            zeroed (inferencing/adaptive_scaling.py:145-169)
// so that a page goes H2D as uint8 (4x fewer bytes than fp32) and comes back as a uint8 mask + one fp32 map.
// Precise pass (inferencing/adaptive_scaling.py:326-386, 477-491):
//   precise  sigmoid of the char-prob logits with the padding zeroed, softmax over the 4 corner-angle channels, and the
//            NCHW -> NHWC permutes of the offset / angle / distance maps, one pass over the network outputs
//   peaks    local maxima of the (optionally masked) char-prob map under a size x size maximum filter that reach the
//            positive threshold (scipy.ndimage.maximum_filter, default 'reflect' border = window clamped to the map)
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
ingest_u8_hwc_kernel(const uint8_t* __restrict__ img, int B, int H, int W, float* __restrict__ out, int Hp, int Wp) {
    const long long total = (long long)B * Hp * Wp;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % Wp);
    long long r = idx / Wp;
    const int y = (int)(r % Hp);
    const int b = (int)(r / Hp);
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (y < H && x < W) {
        const uint8_t* p = img + (((long long)b * H + y) * W + x) * 3;
        c0 = (float)p[0]; c1 = (float)p[1]; c2 = (float)p[2];
    }
    const long long plane = (long long)Hp * Wp;
    float* o = out + (long long)b * 3 * plane + (long long)y * Wp + x;
    o[0] = c0;
    o[plane] = c1;
    o[2 * plane] = c2;
}

__global__ void __launch_bounds__(256)
rough_postprocess_kernel(const float* __restrict__ logit, const float* __restrict__ height, int B, int h, int w, int valid_h,
                         int valid_w, float thr, float height_min, uint8_t* __restrict__ mask_out, float* __restrict__ height_out) {
    const long long total = (long long)B * h * w;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % w);
    const int y = (int)((idx / w) % h);
    const bool inside = y < valid_h && x < valid_w;
    const float s = 1.f / (1.f + expf(-logit[idx]));
    mask_out[idx] = (inside && s >= thr) ? 1 : 0;
    const float hv = height[idx];
    height_out[idx] = (inside && !(hv < height_min)) ? hv : 0.f;
}

__global__ void __launch_bounds__(256)
precise_postprocess_kernel(const float* __restrict__ prob_logit, const float* __restrict__ offset, const float* __restrict__ angle,
                           const float* __restrict__ distance, int B, int h, int w, int dist_channels, int valid_h, int valid_w,
                           float* __restrict__ prob_out, float* __restrict__ offset_out, float* __restrict__ angle_out,
                           float* __restrict__ distance_out) {
    const long long plane = (long long)h * w;
    const long long total = (long long)B * plane;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long b = idx / plane, pix = idx - b * plane;
    const int x = (int)(pix % w), y = (int)(pix / w);
    const bool inside = y < valid_h && x < valid_w;
    prob_out[idx] = inside ? 1.f / (1.f + expf(-prob_logit[idx])) : 0.f;
    const float* op = offset + b * 2 * plane + pix;
    offset_out[idx * 2] = op[0];
    offset_out[idx * 2 + 1] = op[plane];
    const float* ap = angle + b * 4 * plane + pix;
    const float a0 = ap[0], a1 = ap[plane], a2 = ap[2 * plane], a3 = ap[3 * plane];
    const float m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
    const float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m), e3 = expf(a3 - m);
    const float inv = 1.f / (e0 + e1 + e2 + e3);
    reinterpret_cast<float4*>(angle_out)[idx] = make_float4(e0 * inv, e1 * inv, e2 * inv, e3 * inv);
    const float* dp = distance + b * dist_channels * plane + pix;
    for (int c = 0; c < dist_channels; ++c) distance_out[idx * dist_channels + c] = dp[c * plane];
}

__global__ void __launch_bounds__(256)
peak_mask_kernel(const float* __restrict__ prob, const uint8_t* __restrict__ char_mask, int B, int h, int w, int before, int after,
                 float thr, uint8_t* __restrict__ peaks) {
    const long long plane = (long long)h * w;
    const long long total = (long long)B * plane;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long b = idx / plane, pix = idx - b * plane;
    const int x = (int)(pix % w), y = (int)(pix / w);
    const float* pm = prob + b * plane;
    const uint8_t* cm = char_mask ? char_mask + b * plane : nullptr;
    auto at = [&](int yy, int xx) { const long long q = (long long)yy * w + xx; return (cm && !cm[q]) ? 0.f : pm[q]; };
    const float v = at(y, x);
    const int y0 = max(y - before, 0), y1 = min(y + after, h - 1), x0 = max(x - before, 0), x1 = min(x + after, w - 1);
    float mx = v;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) mx = fmaxf(mx, at(yy, xx));
    peaks[idx] = (mx == v && !(v < thr)) ? 1 : 0;
}

}  // namespace

extern "C" {

// img: [B, H, W, 3] uint8 (device) -> out: [B, 3, Hp, Wp] fp32, zero beyond (H, W); Hp >= H, Wp >= W.
int vkocr_ingest_image_u8(const void* img, int B, int H, int W, float* out, int Hp, int Wp, void* stream) {
    VK_REQUIRE(img && out, VKOCR_BAD_ARGUMENT, "ingest_image_u8: null argument");
    VK_REQUIRE(Hp >= H && Wp >= W && H >= 0 && W >= 0, VKOCR_BAD_SHAPE, "ingest_image_u8: padded %dx%d < image %dx%d", Hp, Wp, H, W);
    const long long total = (long long)B * Hp * Wp;
    if (total == 0) return VKOCR_OK;
    ingest_u8_hwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint8_t*>(img), B, H, W, out, Hp, Wp);
    VK_CHECK_LAUNCH("ingest_u8_hwc_kernel");
    return VKOCR_OK;
}

// logit / height: [B, 1, h, w] fp32 network outputs; rows >= valid_h and columns >= valid_w are padding.
int vkocr_rough_postprocess(const float* logit, const float* height, int B, int h, int w, int valid_h, int valid_w, float thr,
                            float height_min, void* mask_out, float* height_out, void* stream) {
    VK_REQUIRE(logit && height && mask_out && height_out, VKOCR_BAD_ARGUMENT, "rough_postprocess: null argument");
    const long long total = (long long)B * h * w;
    if (total == 0) return VKOCR_OK;
    rough_postprocess_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        logit, height, B, h, w, valid_h, valid_w, thr, height_min, reinterpret_cast<uint8_t*>(mask_out), height_out);
    VK_CHECK_LAUNCH("rough_postprocess_kernel");
    return VKOCR_OK;
}

// Network outputs of forward_precise (NCHW fp32) -> the arrays the reference hands to its polygon builder: char-prob score
// map (sigmoid, padding zeroed) [B,h,w], up-left offsets [B,h,w,2], corner-angle distribution (softmax) [B,h,w,4], corner
// distances [B,h,w,dist_channels].
int vkocr_precise_postprocess(const float* prob_logit, const float* offset, const float* angle, const float* distance, int B, int h,
                              int w, int dist_channels, int valid_h, int valid_w, float* prob_out, float* offset_out,
                              float* angle_out, float* distance_out, void* stream) {
    VK_REQUIRE(prob_logit && offset && angle && distance && prob_out && offset_out && angle_out && distance_out, VKOCR_BAD_ARGUMENT,
               "precise_postprocess: null argument");
    VK_REQUIRE(dist_channels >= 1 && dist_channels <= 8, VKOCR_BAD_SHAPE, "precise_postprocess: %d distance channels", dist_channels);
    VK_REQUIRE((reinterpret_cast<uintptr_t>(angle_out) & 15) == 0, VKOCR_BAD_ALIGN, "precise_postprocess: angle_out not 16-byte aligned");
    const long long total = (long long)B * h * w;
    if (total == 0) return VKOCR_OK;
    precise_postprocess_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        prob_logit, offset, angle, distance, B, h, w, dist_channels, valid_h, valid_w, prob_out, offset_out, angle_out, distance_out);
    VK_CHECK_LAUNCH("precise_postprocess_kernel");
    return VKOCR_OK;
}

// peaks[b,y,x] = 1 where the (char_mask-ed, optional uint8 [B,h,w]) prob map equals its size x size maximum filter and is
// >= thr.  Window of scipy.ndimage.maximum_filter with origin 0: [-(size/2), size - 1 - size/2] around the pixel.
int vkocr_peak_mask(const float* prob, const void* char_mask, int B, int h, int w, int size, float thr, void* peaks_out, void* stream) {
    VK_REQUIRE(prob && peaks_out, VKOCR_BAD_ARGUMENT, "peak_mask: null argument");
    VK_REQUIRE(size >= 1 && size <= 31, VKOCR_BAD_SHAPE, "peak_mask: filter size %d", size);
    const long long total = (long long)B * h * w;
    if (total == 0) return VKOCR_OK;
    peak_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        prob, reinterpret_cast<const uint8_t*>(char_mask), B, h, w, size / 2, size - 1 - size / 2, thr,
        reinterpret_cast<uint8_t*>(peaks_out));
    VK_CHECK_LAUNCH("peak_mask_kernel");
    return VKOCR_OK;
}

}  // extern "C"
