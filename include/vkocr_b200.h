/* vkocr_b200.h — C ABI of libvkocr_b200.so: the sm_100a (B200) kernels behind the forward / backward / loss hot path
 * of vkit_open_model's adaptive-scaling text-detection network.
 *
 * The reference (vkit-dev/vkit-ocr-model-adaptive-scaling) has no native code and no FFI: its hot path sits directly
 * behind Python classes and dispatches to ATen / cuDNN / cuBLAS.  Each entry point below replaces the stock-PyTorch
 * call sequence of the cited reference lines (paths relative to the reference root, `vkit_open_model/...`).  The
 * Python host layer (`vkit_ocr_model_adaptive_scaling_b200/ops.py`) binds them with ctypes; INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions
 *   - Plain C: raw device pointers, integer extents, strides in ELEMENTS, no torch types.
 *   - The library never allocates, frees or retains caller memory; every launch is asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - Return value: 0 on success, a negative VkocrStatus otherwise; vkocr_last_error() returns a thread-local message.
 *     An unsupported shape / dtype is an error, never a silent fallback.  There is no CPU path.
 *   - `dtype` tags the STORAGE type of activations: 0 = float32, 1 = bfloat16.  Arithmetic is always fp32
 *     (fp32 accumulation in tensor memory for the tcgen05 GEMMs).  Parameters and their gradients are fp32.
 *   - Activations are NHWC: pixel-major, channel-contiguous, with a pixel stride `ld >= C` (elements); 16-byte vector
 *     paths need C and ld to be multiples of 16 bytes / sizeof(storage).
 *   - Re-entrant; safe to call concurrently from several host threads on different streams (PyTorch runs backward on
 *     its own engine thread).
 */
#ifndef VKOCR_B200_H
#define VKOCR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

enum VkocrStatus {
    VKOCR_OK = 0,
    VKOCR_BAD_SHAPE = -1,
    VKOCR_BAD_ALIGN = -2,
    VKOCR_UNSUPPORTED_DTYPE = -3,
    VKOCR_CUDA_ERROR = -4,
    VKOCR_WORKSPACE_TOO_SMALL = -5,
    VKOCR_BAD_ARGUMENT = -6
};

enum VkocrDtype { VKOCR_F32 = 0, VKOCR_BF16 = 1 };

/* ---------------------------------------------------------------------------------------------- runtime basics */
int vkocr_abi_version(void);
const char* vkocr_last_error(void);
int vkocr_device_check(int device); /* 0 iff `device` is compute capability 10.x */
long long vkocr_launch_count(void);  /* kernels launched by this library in this process so far */

/* ------------------------------------------------------------------------- GEMM / implicit-GEMM convolution
 * Replaces: helper.conv1x1 = nn.Linear on BHWC (model/helper.py:18-22), helper.conv3x3 / conv5x5 'same' convolutions
 * (helper.py:25-40), the patchify convolutions pconv2x2 / pconv4x4 (helper.py:43-58) and their autograd backward
 * (aten::addmm / mm / convolution / convolution_backward), including the fused epilogues of ConvNextBlockLayer
 * (model/convnext.py:29-37,56-58: bias, exact GELU, layer scale, stochastic-depth mask, residual add).
 *
 *   NT (forward / data gradient):  D[m, n]      = sum_{tap, c} X[pix(m) + off(tap), c] * Wp[n, tap * c_pad + c]
 *   TN (weight gradient):          G[tap, i, j] += sum_{pix}   P[pix, i] * Q[pix + off(tap), j]
 *
 * bf16 storage -> tcgen05.mma (TMEM accumulators, TMA-fed, persistent, warp-specialised); fp32 storage -> fp32 SIMT.
 * `backend` 1 forces the SIMT kernel for bf16 too (cross-check in the tests). */
typedef struct VkocrConvGeom {
    int batch, H, W;      /* pixel grid of the activation operand(s); plain GEMM: batch = 1, H = 1, W = rows */
    int ks;               /* square kernel size 1, 3 or 5 (zero padding ks/2, stride 1) */
    int C;                /* channels contracted per tap (NT) / channels of P (TN) */
    long long ld_x;       /* pixel stride of X (NT) / of P (TN), elements */
    int c_pad;            /* NT: per-tap K extent of the packed weight, multiple of 64 */
} VkocrConvGeom;

typedef struct VkocrEpilogue {
    void* out;            /* [rows, ldo] storage dtype, or fp32 when out_f32 != 0 */
    long long ldo;
    int out_f32;
    int accumulate;       /* out += value (fp32 atomics; requires out_f32) */
    void* out_pre;        /* optional second output: (acc + bias) before the activation; act 3: gelu'(acc + bias) */
    long long ld_pre;
    const float* bias;    /* [N] or NULL */
    int act;              /* 0 none, 1 exact erf GELU (helper.py:100-101), 2 multiply by gelu'(aux),
                             3 exact GELU with out_pre = row_scale * gelu'(acc + bias) and out zeroed where row_scale == 0
                             (stochastic-depth mask, convnext.py:41-53), 4 multiply by aux */
    const float* col_scale;  /* [N] or NULL — ConvNeXt layer scale (convnext.py:38,56) */
    const float* row_scale;  /* [rows / rows_per_group] or NULL — stochastic-depth mask (convnext.py:41-53) */
    int rows_per_group;
    const void* residual; /* optional [rows, ld_res] storage dtype, added last (convnext.py:58) */
    long long ld_res;
    const void* aux;      /* act == 2 / 4 operand */
    long long ld_aux;
    long long tn_s_tap, tn_s_i, tn_s_j; /* TN: G[tap,i,j] lands at out[tap*s_tap + i*s_i + j*s_j] (e.g. Conv2d OIHW) */
} VkocrEpilogue;

int vkocr_gemm_nt(int dtype, int backend, const void* x, const VkocrConvGeom* g, const void* w_packed, int N,
                  const VkocrEpilogue* ep, void* stream);

/* Fused head group: the 3x3 (or 1x1 / 5x5) conv of all heads that read one neck tensor with each head's
 * LayerNorm -> GELU -> Linear(inner -> out <= 4) (-> Softplus) tail applied in the GEMM epilogue from the fp32
 * accumulators (UperNextHead.forward model/upernext.py:233-248, FpnHead.forward model/fpn.py:193-208, nn.Softplus
 * model/adaptive_scaling.py:101,140).  N = num_heads * slot; head h owns columns [h*slot, h*slot + inner[h]).
 * ep->out (the conv output, needed by the backward) may be NULL for inference. bf16 storage only. */
#define VKOCR_MAX_HEADS 4
typedef struct VkocrHeadTail {
    int num_heads;
    int slot;
    long long pixels_per_image;
    const float* gamma[VKOCR_MAX_HEADS];
    const float* beta[VKOCR_MAX_HEADS];
    const float* w2[VKOCR_MAX_HEADS];
    const float* b2[VKOCR_MAX_HEADS];
    float* out[VKOCR_MAX_HEADS];
    int inner[VKOCR_MAX_HEADS];
    int out_channels[VKOCR_MAX_HEADS];
    int softplus[VKOCR_MAX_HEADS];
} VkocrHeadTail;
int vkocr_gemm_nt_heads(int dtype, const void* x, const VkocrConvGeom* g, const void* w_packed, int N, const VkocrEpilogue* ep,
                        const VkocrHeadTail* heads, void* stream);
int vkocr_gemm_tn(int dtype, int backend, const void* pmat, const VkocrConvGeom* g, const void* qmat, int J, long long ld_q,
                  const VkocrEpilogue* ep, void* stream);

/* ------------------------------------------------------------- heads with up-sampling: convolve first, resample after
 * Replaces F.interpolate(x, scale_factor) -> conv k x k -> LayerNorm -> GELU -> Linear (-> Softplus) of
 * UperNextHead.forward (model/upernext.py:233-248, bilinear) / FpnHead.forward (model/fpn.py:193-208, nearest; 5x5 for
 * factors in (2, 4]) and nn.Softplus (model/adaptive_scaling.py:101,140) for upsampling_factor >= 2.  The channel mixing
 * of the convolution commutes with the per-channel interpolation:
 *     conv(up(x))[r, s] = bias + sum_{dy,dx} up(Z_{dy,dx})[r + dy - k/2, s + dx - k/2],  Z_{dy,dx} = x . W[:, :, dy, dx]^T
 * so the contraction is ONE plain vkocr_gemm_nt on the low-resolution map (N = k*k*ntot columns ordered
 * tap * ntot + head * slot + n; factor^2 fewer FLOPs, the up-sampled tensor never exists) and the interpolation runs on its
 * output:
 * vkocr_head_combine_fwd: z [B*h*w, ld_z] -> per head interpolate + shift + sum + conv bias -> LayerNorm -> GELU ->
 *   Linear(inner -> O <= 4) (-> Softplus) -> heads->out[h] (NCHW fp32 at (factor*h, factor*w)); conv_out (nullable):
 *   [B*H*W, ld_conv] the pre-LayerNorm conv output in storage dtype, which vkocr_head_tail_bwd consumes.
 * vkocr_head_combine_bwd: the exact adjoint, d(conv output) [B*H*W, ld_dc] (`width` = ntot channels) -> dz [B*h*w, ld_z];
 *   the data / weight gradients are then plain vkocr_gemm_nt / vkocr_gemm_tn calls on (dz, x).
 * mode: 0 bilinear (align_corners=False), 1 nearest.  algo: 0 = pick (a shared-memory-tiled kernel for factor 2, 3x3),
 * 1 = force the generic gather kernel (cross-check in the tests). */
int vkocr_head_combine_fwd(int dtype, const void* z, long long ld_z, int B, int h, int w, int factor, int mode, int ks, int ntot,
                           const float* conv_bias, const VkocrHeadTail* heads, void* conv_out, long long ld_conv, int algo,
                           void* stream);
int vkocr_head_combine_bwd(int dtype, const void* dconv, long long ld_dc, int B, int h, int w, int factor, int mode, int ks,
                           int width, void* dz, long long ld_z, int algo, void* stream);

/* ---------------------------------------------------------------------------------------- depthwise 7x7 conv
 * Replaces helper.dconv7x7 (helper.py:61-73; convnext.py:30) forward, data gradient (same kernel, mirrored taps,
 * `add` fuses the residual gradient of convnext.py:58) and weight gradient.  `wt` is the [49][C] fp32 tap table
 * written by vkocr_pack_weight. */
int vkocr_dwconv7_fwd(int dtype, const void* x, long long ld_x, void* y, long long ld_y, int B, int H, int W, int C,
                      const float* wt, const float* bias, const void* add, long long ld_add, void* stream);
int vkocr_dwconv7_wgrad(int dtype, const void* dy, long long ld_dy, const void* x, long long ld_x, int B, int H, int W, int C,
                        float* dw, void* stream);

/* ------------------------------------------------------------------------------- LayerNorm (+ GELU), column sums
 * Replaces helper.ln = nn.LayerNorm(C, eps=1e-6) on BHWC (helper.py:96-97) and the LN -> GELU pair of the neck / head
 * blocks (upernext.py:21-45, fpn.py:21-48), forward and backward; backward also accumulates dgamma, dbeta and the
 * column sum of dx (the bias gradient of the producing conv / Linear).  vkocr_colsum: bias gradients of Linear layers. */
int vkocr_layernorm_fwd(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C,
                        const float* gamma, const float* beta, float eps, int act, float* mean, float* rstd, void* stream);
int vkocr_layernorm_bwd(int dtype, const void* dy, long long ld_dy, const void* x, long long ld_x, const float* mean,
                        const float* rstd, const float* gamma, const float* beta, int act, void* dx, long long ld_dx,
                        long long rows, int C, float* dgamma, float* dbeta, float* dxsum, void* stream);
int vkocr_colsum(int dtype, const void* x, long long ld, long long rows, int C, float* out, void* stream);
/* y = scale[row / rows_per_group] * x (skipped when y is null) and out[c] += sum_rows scale * x[row, c]: the
 * stochastic-depth mask applied to the incoming gradient of a ConvNeXt layer (model/convnext.py:41-53) fused with the
 * bias-gradient column sum. */
int vkocr_scale_rows_colsum(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C,
                            const float* scale, int rows_per_group, float* out, void* stream);

/* ------------------------------------------------------------------------------------ resampling / pooling / gathers
 * vkocr_upsample_*: F.interpolate(mode='bilinear' (0) | 'nearest' (1), align_corners=False) to an explicit size, writing
 *   (or adding) straight into a channel slice — the top-down `+=` (upernext.py:174-182, fpn.py:121-129), the
 *   up-sample + torch.cat of the neck output (upernext.py:189-197, fpn.py:136-144), the PPM (upernext.py:76-82) and the
 *   head's x2 up-sampling (upernext.py:237-244, fpn.py:197-204); `_bwd` is the exact adjoint in gather form.
 * vkocr_avgpool_*: nn.AdaptiveAvgPool2d(S) (upernext.py:59-65), bins [floor(i*in/S), ceil((i+1)*in/S)).
 * vkocr_patchify_image: NCHW fp32 image -> [pixels, p*p*Cin (padded)] rows for the stem GEMM (convnext.py:106-123).
 * vkocr_space_to_depth2: NHWC -> [pixels/4, 4C] rows for pconv2x2 (convnext.py:89-99); dir 1 = adjoint scatter.
 * vkocr_copy_channels: channel-slice copy / accumulate (torch.cat of un-resampled levels). */
int vkocr_upsample_fwd(int dtype, const void* src, long long ld_s, int h, int w, void* dst, long long ld_d, int H, int W, int B,
                       int C, int mode, int accumulate, void* stream);
int vkocr_upsample_bwd(int dtype, const void* ddst, long long ld_d, int H, int W, void* dsrc, long long ld_s, int h, int w, int B,
                       int C, int mode, int accumulate, void* stream);
/* The same adjoint in two separable passes (x, then y) through a caller-owned fp32 workspace [B, H, w, C]: for large
 * scale factors (PPM / top-down paths, upernext.py:59-82,174-197) where the one-pass gather has few threads and large
 * windows.  Workspace: B * H * w * C floats.  Requires 16-byte aligned channel vectors. */
int vkocr_upsample_bwd_separable(int dtype, const void* ddst, long long ld_d, int H, int W, void* dsrc, long long ld_s, int h, int w,
                                 int B, int C, int mode, int accumulate, float* workspace, void* stream);
int vkocr_avgpool_fwd(int dtype, const void* x, long long ld_x, int H, int W, void* y, long long ld_y, int S, int B, int C,
                      void* stream);
/* The same pooling in two separable passes (columns of every row, then rows) through a caller-owned fp32 workspace of
 * B * H * S * C floats: for bins of hundreds of pixels (PPM scales 1..6, upernext.py:59-82). */
int vkocr_avgpool_fwd_separable(int dtype, const void* x, long long ld_x, int H, int W, void* y, long long ld_y, int S, int B, int C,
                                float* workspace, void* stream);
int vkocr_avgpool_bwd(int dtype, const void* dy, long long ld_y, int S, void* dx, long long ld_x, int H, int W, int B, int C,
                      int accumulate, void* stream);
int vkocr_patchify_image(int dtype, const float* img, int B, int Cin, int H, int W, int p, void* out, int c_pad, void* stream);
int vkocr_space_to_depth2(int dtype, void* x, long long ld_x, int B, int H, int W, int C, void* y, long long ld_y, int dir,
                          int accumulate, void* stream);
int vkocr_copy_channels(int dtype, const void* src, long long ld_s, void* dst, long long ld_d, long long rows, int C,
                        int accumulate, void* stream);

/* ----------------------------------------------------------------------------------------------- head tail
 * Replaces the tail of UperNextHead.forward / FpnHead.forward (upernext.py:245-247, fpn.py:205-207): LayerNorm(inner)
 * -> GELU -> Linear(inner -> O <= 4), plus nn.Softplus (adaptive_scaling.py:101,140); writes the NCHW fp32 map.
 * Backward recomputes LN / GELU from the conv output and accumulates every parameter gradient of the tail. */
int vkocr_head_tail_fwd(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                        const float* w2, const float* b2, int O, int softplus, float* out, long long pixels_per_image,
                        long long rows, void* stream);
int vkocr_head_tail_bwd(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                        const float* w2, int O, int softplus, const float* out, const float* dout, long long pixels_per_image,
                        long long rows, void* dx, long long ld_dx, float* dgamma, float* dbeta, float* dw2, float* db2,
                        float* dbias, void* stream);

/* ------------------------------------------------------------------------------- label-point backward of the heads
 * Three of the four precise heads enter the loss only through their values at the (B, P) label points
 * (get_label_point_feature, loss_function/adaptive_scaling.py:167-179,235-260), so their upstream gradient maps are zero
 * elsewhere and the dense backward of model/upernext.py:233-248 / model/fpn.py:193-208 multiplies zeros.  These entry points
 * compute the SAME gradients from the E = B*P label pixels only:
 * vkocr_points_claim: pix_index[e] = b*H*W + y*W + x for the first entry on a pixel, -1 for duplicates / outside points
 *   (the gradient map already holds the sum of duplicates); `owner`: B*H*W ints zeroed by the caller.
 * vkocr_head_tail_bwd_points: vkocr_head_tail_bwd over the listed pixel rows; dx = [E, ld_dx] gradient rows G.
 * vkocr_gather_up_taps: A[e, tap*C + c] = up(x)[r_e + dy - k/2, s_e + dx - k/2, c]; then dW_tap = G^T . A_tap (vkocr_gemm_tn).
 * vkocr_scatter_up_taps: dX += adjoint of up-sample + tap shift applied to U = G . W (vkocr_gemm_nt), atomically. */
int vkocr_points_claim(const long long* py, const long long* px, int B, int P, int H, int W, int* owner, int* pix_index, void* stream);
int vkocr_head_tail_bwd_points(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                               const float* w2, int O, int softplus, const float* out, const float* dout, long long pixels_per_image,
                               const int* row_index, long long entries, int x_compact, void* dx, long long ld_dx, float* dgamma,
                               float* dbeta, float* dw2, float* db2, float* dbias, void* stream);
/* Label-point FORWARD (opt-in, training): conv-output rows of the label pixels only, conv[e, :] = bias + A[e, :] . W^T with A
 * from vkocr_gather_up_taps (one small vkocr_gemm_nt), then this tail writes the NCHW maps at those pixels (the rest of the map
 * is the caller's zero fill).  The loss reads the offset / angle / distance maps at the label points alone
 * (loss_function/adaptive_scaling.py:235-260), so loss and gradients equal the dense forward's. */
int vkocr_head_tail_fwd_points(int dtype, const void* x, long long ld_x, int inner, int slice_w, const float* gamma, const float* beta,
                               const float* w2, const float* b2, int O, int softplus, float* out, long long pixels_per_image,
                               const int* row_index, long long entries, void* stream);
int vkocr_gather_up_taps(int dtype, const void* x, long long ld_x, int B, int h, int w, int C, int factor, int mode, int ks,
                         const int* pix_index, int E, void* a, void* stream);
int vkocr_scatter_up_taps(int dtype, const void* u, int B, int h, int w, int C, int factor, int mode, int ks, const int* pix_index, int E,
                          void* dx, long long ld_dx, void* stream);

/* ------------------------------------------------------------------------------------------------ fused losses
 * vkocr_rough_loss_*: AdaptiveScalingRoughLossFunction.__call__ (loss_function/adaptive_scaling.py:53-131): core-box
 *   crop, 5*focal + 1*dice + 1*masked log-space smooth-L1 (factors are arguments).  sums: 6 zeroed doubles; coef: 6
 *   floats, coef[0] = loss.
 * vkocr_precise_loss_*: AdaptiveScalingPreciseLossFunction.__call__ (:181-346): masked pos / neg L2 of sigmoid(prob) in
 *   the core box + label-point gather (:167-179) terms; `factors` = 7 device floats {pos_l2, neg_l2, offset_l1,
 *   distance_regulation, angle_ce, corner_distance, loss_factor}.  sums: 8 zeroed doubles; coef: 3 floats.  Label points
 *   are (B, P) int64 with negative indices wrapped like torch's advanced indexing; a point outside the map (the reference
 *   raises IndexError) makes the loss NaN and is skipped by both passes -- nothing is read or written out of bounds.
 *   gt_off: the (B, P, 2) offsets as fp32 (the dataset emits int64; the caller converts). */
int vkocr_rough_loss_fwd(const float* logit, const float* height, const float* gt_mask, const float* gt_score, int B, int H,
                         int W, int up, int left, int CH, int CW, float height_min, float score_min, float focal_factor,
                         float dice_factor, float l1_factor, double* sums, float* coef, void* stream);
int vkocr_rough_loss_bwd(const float* logit, const float* height, const float* gt_mask, const float* gt_score, int B, int H,
                         int W, int up, int left, int CH, int CW, float height_min, float score_min, const float* coef,
                         const float* grad_out, float* dlogit, float* dheight, void* stream);
int vkocr_precise_loss_fwd(const float* prob, const float* off, const float* ang, const float* dist, const float* gt_score,
                           const float* gt_mask, int B, int H, int W, int up, int left, int CH, int CW, const long long* py,
                           const long long* px, const float* gt_off, const float* gt_ang, const float* gt_dist, int P,
                           float beta, const float* factors, double* sums, float* coef, void* stream);
int vkocr_precise_loss_bwd(const float* prob, const float* off, const float* ang, const float* dist, const float* gt_score,
                           const float* gt_mask, int B, int H, int W, int up, int left, int CH, int CW, const long long* py,
                           const long long* px, const float* gt_off, const float* gt_ang, const float* gt_dist, int P,
                           float beta, const float* factors, const float* coef, const float* grad_out, float* dprob,
                           float* doff, float* dang, float* ddist, void* stream);

/* ---------------------------------------------------------------------------------------------- primitive losses
 * kind: 0 focal (focal_with_logits.py:18-47; p0 alpha, p1 gamma), 1 dice (dice.py:17-35), 2 L1, 3 smooth-L1 (l1.py:19-47;
 * p0 beta), 4 L2 (l2.py:18-34), 5 WAHR (weight_adaptive_heatmap_regression.py:18-33; p0 gamma); optional mask;
 * pre_sigmoid applies the primitive to sigmoid(pred).  sums: 3 zeroed doubles; coef: 3 floats.
 * vkocr_soft_ce_*: F.cross_entropy with probability targets over axis 1 (cross_entropy_with_logits.py:16-19).
 * vkocr_hard_negative_bce_*: weighted_bce_with_logits.py:18-54 with a device-side radix select instead of the
 *   reference's two host syncs + topk.  state_hist: 264 zeroed u64; sums: 4 zeroed doubles; coef: 4 floats. */
int vkocr_pointwise_loss_fwd(int kind, int pre_sigmoid, const float* pred, const float* gt, const float* mask, long long n,
                             float p0, float p1, double* sums, float* coef, void* stream);
int vkocr_pointwise_loss_bwd(int kind, int pre_sigmoid, const float* pred, const float* gt, const float* mask, long long n,
                             float p0, float p1, const float* coef, const float* grad_out, float* dpred, void* stream);
int vkocr_soft_ce_fwd(const float* pred, const float* gt, long long outer, int C, long long inner, double* sums, float* coef,
                      void* stream);
int vkocr_soft_ce_bwd(const float* pred, const float* gt, long long outer, int C, long long inner, const float* coef,
                      const float* grad_out, float* dpred, void* stream);
int vkocr_hard_negative_bce_fwd(const float* pred, const float* gt, const float* mask, long long n, float negative_ratio,
                                float eps, unsigned long long* state_hist, double* sums, float* coef, void* stream);
int vkocr_hard_negative_bce_bwd(const float* pred, const float* gt, const float* mask, long long n,
                                const unsigned long long* state_hist, const float* coef, const float* grad_out,
                                unsigned long long* tie_counter, float* dpred, void* stream);

/* --------------------------------------------------------------------------------- parameter staging / finalisers
 * The fp32 master parameters keep the reference's state_dict layout (Linear (out,in), Conv2d OIHW, depthwise
 * (C,1,7,7)); vkocr_pack_weight writes the K-major (optionally tap-mirrored / column-scaled) operand copies the GEMMs
 * and the depthwise kernel read; vkocr_unpack_grad / vkocr_accumulate_f32 / vkocr_mlp2_grad_finalize move weight
 * gradients from GEMM order into the parameters' .grad; vkocr_scale_rows applies the stochastic-depth mask to a
 * gradient (convnext.py:41-53). */
int vkocr_pack_weight(const float* w, long long s_row, long long s_tap, long long s_col, int rows, int taps, int cols, int flip,
                      const float* col_scale, void* out, int out_dtype, long long o_row, long long o_tap, void* stream);
int vkocr_unpack_grad(const float* g, int N, int T, int C, float* y, long long s_n, long long s_t, long long s_c, void* stream);
/* S: [C, K] product dY^T G with row stride ld_s, scaled by s_scale on the way in (1 / p_keep when G was stored with the
 * dropped samples zeroed); sU[c * ld_su]: masked column sums of dY (a column of the same product when G carries the mask
 * channel). */
int vkocr_mlp2_grad_finalize(const float* S, long long ld_s, float s_scale, const float* sU, long long ld_su, const float* W2,
                             const float* b2, const float* gamma, int C, int K, float* dW2, float* dgamma, float* db2, void* stream);
int vkocr_accumulate_f32(const float* a, float* y, long long n, void* stream);
/* y[i0*y0 + i1*y1 + i2*y2] += g[i0*g0 + i1*g1 + i2*g2] for i0 < n0, i1 < n1, i2 < n2 (strides in elements): a weight
 * gradient from GEMM order (e.g. the tap-major rows of the head-group product) into the Conv2d OIHW parameter layout. */
int vkocr_scatter_add_f32(const float* g, long long g0, long long g1, long long g2, int n0, int n1, int n2, float* y, long long y0,
                          long long y1, long long y2, void* stream);
int vkocr_scale_rows(int dtype, const void* x, long long ld_x, void* y, long long ld_y, long long rows, int C,
                     const float* scale, int rows_per_group, void* stream);
/* x[row, col0] = first, x[row, col0+1 .. col0+ncols) = 0: the constant column appended to the LayerNorm output so that the
 * weight-gradient GEMM of the MLP up-projection (model/convnext.py:33) also delivers its bias gradient. */
int vkocr_set_columns(int dtype, void* x, long long ld, long long rows, int col0, int ncols, float first, void* stream);
/* Asynchronous zero fill of caller memory on `stream` (atomic-accumulate workspaces, gradient buckets). */
int vkocr_zero(void* p, long long bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ optimizer tail
 * Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step over ~300 tensors
 * (experiment/adaptive_scaling/train.py:73-80,287-298,468-478) by two passes over the flat fp32 buckets.
 * vkocr_sumsq_f32: out[0] += sum x^2 (fp64).  vkocr_adamw_step: AdamW with decoupled weight decay on a flat range; the
 * gradient is scaled by grad_scale and clipped by min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6)) when sumsq is
 * given; bias_corr1/2 = 1 - beta1^t / 1 - beta2^t. */
int vkocr_sumsq_f32(const float* x, long long n, double* out, void* stream);
int vkocr_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, const double* sumsq,
                     float max_norm, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------- rough inference pre/post ops
 * vkocr_ingest_image_u8: uint8 HWC page image(s) -> zero-padded fp32 NCHW network input (inferencing/opt.py:16-41
 * pad_mat_to_make_divisible + the transpose/astype of inferencing/adaptive_scaling.py:116-121).
 * vkocr_rough_postprocess: sigmoid >= thr mask (uint8) and the height map with padding and too-small heights zeroed
 * (inferencing/adaptive_scaling.py:145-169). */
int vkocr_ingest_image_u8(const void* img, int B, int H, int W, float* out, int Hp, int Wp, void* stream);
int vkocr_rough_postprocess(const float* logit, const float* height, int B, int h, int w, int valid_h, int valid_w, float thr,
                            float height_min, void* mask_out, float* height_out, void* stream);

/* ----------------------------------------------------------------------------------- precise inference tensor post-ops
 * vkocr_precise_postprocess: the tensor side of AdaptiveScalingInferencing.precise_infer after forward_precise
 * (inferencing/adaptive_scaling.py:326-386): sigmoid of the char-prob logits with the padding forced to 0, softmax over
 * the 4 corner-angle channels, NCHW -> NHWC permutes of the offset / angle / distance maps.
 * vkocr_peak_mask: the peak picking of precise_build_grouped_polygons (:477-491): mat[~char_mask] = 0,
 * scipy.ndimage.maximum_filter(mat, size) == mat, and mat >= thr. */
int vkocr_precise_postprocess(const float* prob_logit, const float* offset, const float* angle, const float* distance, int B, int h,
                              int w, int dist_channels, int valid_h, int valid_w, float* prob_out, float* offset_out,
                              float* angle_out, float* distance_out, void* stream);
int vkocr_peak_mask(const float* prob, const void* char_mask, int B, int h, int w, int size, float thr, void* peaks_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VKOCR_B200_H */
