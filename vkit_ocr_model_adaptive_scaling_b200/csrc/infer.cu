// Tensor-side pre/post operations of the rough inference pass, on the device:
//   ingest   uint8 HWC page image(s) -> zero-padded fp32 NCHW network input (pad to a multiple of the backbone stride:
//            inferencing/opt.py:16-41 pad_mat_to_make_divisible; transpose + float: inferencing/adaptive_scaling.py:116-121)
//   postproc mask = sigmoid(logit) >= thr as uint8, height map with the padding region and heights below the minimum
//            zeroed (inferencing/adaptive_scaling.py:145-169)
// so that a page goes H2D as uint8 (4x fewer bytes than fp32) and comes back as a uint8 mask + one fp32 map.
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

}  // extern "C"
