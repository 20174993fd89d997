// C-ABI entry points of the dense contractions (include/vkocr_b200.h, section "GEMM / implicit-GEMM convolution").
#include "gemm_common.cuh"

int vkocr_gemm_tc_nt(const void*, const VkocrConvGeom*, const void*, int, const VkocrEpilogue*, const VkocrHeadTail*, cudaStream_t);
int vkocr_gemm_tc_tn(const void*, const VkocrConvGeom*, const void*, int, long long, const VkocrEpilogue*, cudaStream_t);
int vkocr_gemm_simt_nt(int, const void*, const VkocrConvGeom*, const void*, int, const VkocrEpilogue*, cudaStream_t);
int vkocr_gemm_simt_tn(int, const void*, const VkocrConvGeom*, const void*, int, long long, const VkocrEpilogue*, cudaStream_t);

extern "C" {

// backend: 0 = native for the dtype (bf16 -> tcgen05, fp32 -> SIMT fp32), 1 = force the SIMT kernel (tests).
int vkocr_gemm_nt(int dtype, int backend, const void* x, const VkocrConvGeom* g, const void* w_packed, int N,
                  const VkocrEpilogue* ep, void* stream) {
    VK_REQUIRE(x && g && w_packed && ep && ep->out, VKOCR_BAD_ARGUMENT, "gemm_nt: null argument");
    VK_REQUIRE(!(ep->accumulate && !ep->out_f32), VKOCR_BAD_ARGUMENT, "gemm_nt: accumulate needs an fp32 output");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == VKOCR_BF16 && backend == 0) return vkocr_gemm_tc_nt(x, g, w_packed, N, ep, nullptr, s);
    return vkocr_gemm_simt_nt(dtype, x, g, w_packed, N, ep, s);
}

// The conv of a head group with the head tails fused into the epilogue (bf16 / tcgen05 only; fp32 mode runs
// vkocr_gemm_nt + vkocr_head_tail_fwd).  ep->out may be null: the conv output is then not materialised (inference).
int vkocr_gemm_nt_heads(int dtype, const void* x, const VkocrConvGeom* g, const void* w_packed, int N, const VkocrEpilogue* ep,
                        const VkocrHeadTail* heads, void* stream) {
    VK_REQUIRE(x && g && w_packed && ep && heads, VKOCR_BAD_ARGUMENT, "gemm_nt_heads: null argument");
    VK_REQUIRE(dtype == VKOCR_BF16, VKOCR_UNSUPPORTED_DTYPE, "gemm_nt_heads: bf16 storage only (dtype %d)", dtype);
    VK_REQUIRE(heads->num_heads >= 1 && heads->num_heads <= VKOCR_MAX_HEADS && heads->slot % 16 == 0 && heads->slot <= 256 &&
                   N == heads->num_heads * heads->slot,
               VKOCR_BAD_SHAPE, "gemm_nt_heads: %d heads x slot %d vs N %d", heads->num_heads, heads->slot, N);
    for (int h = 0; h < heads->num_heads; ++h)
        VK_REQUIRE(heads->gamma[h] && heads->beta[h] && heads->w2[h] && heads->b2[h] && heads->out[h] && heads->inner[h] >= 1 &&
                       heads->inner[h] <= heads->slot && heads->out_channels[h] >= 1 && heads->out_channels[h] <= 4,
                   VKOCR_BAD_ARGUMENT, "gemm_nt_heads: head %d parameters", h);
    VK_REQUIRE(!ep->out_f32 && !ep->accumulate && !ep->out_pre && ep->act == 0 && !ep->col_scale && !ep->row_scale && !ep->residual,
               VKOCR_BAD_ARGUMENT, "gemm_nt_heads: only the bias epilogue combines with the head tail");
    return vkocr_gemm_tc_nt(x, g, w_packed, N, ep, heads, reinterpret_cast<cudaStream_t>(stream));
}

int vkocr_gemm_tn(int dtype, int backend, const void* pmat, const VkocrConvGeom* g, const void* qmat, int J, long long ld_q,
                  const VkocrEpilogue* ep, void* stream) {
    VK_REQUIRE(pmat && g && qmat && ep && ep->out, VKOCR_BAD_ARGUMENT, "gemm_tn: null argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == VKOCR_BF16 && backend == 0) return vkocr_gemm_tc_tn(pmat, g, qmat, J, ld_q, ep, s);
    return vkocr_gemm_simt_tn(dtype, pmat, g, qmat, J, ld_q, ep, s);
}

}  // extern "C"
