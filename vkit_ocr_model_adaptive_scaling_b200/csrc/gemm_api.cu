// C-ABI entry points of the dense contractions (include/vkocr_b200.h, section "GEMM / implicit-GEMM convolution").
#include "gemm_common.cuh"

int vkocr_gemm_tc_nt(const void*, const VkocrConvGeom*, const void*, int, const VkocrEpilogue*, cudaStream_t);
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
    if (dtype == VKOCR_BF16 && backend == 0) return vkocr_gemm_tc_nt(x, g, w_packed, N, ep, s);
    return vkocr_gemm_simt_nt(dtype, x, g, w_packed, N, ep, s);
}

int vkocr_gemm_tn(int dtype, int backend, const void* pmat, const VkocrConvGeom* g, const void* qmat, int J, long long ld_q,
                  const VkocrEpilogue* ep, void* stream) {
    VK_REQUIRE(pmat && g && qmat && ep && ep->out, VKOCR_BAD_ARGUMENT, "gemm_tn: null argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == VKOCR_BF16 && backend == 0) return vkocr_gemm_tc_tn(pmat, g, qmat, J, ld_q, ep, s);
    return vkocr_gemm_simt_tn(dtype, pmat, g, qmat, J, ld_q, ep, s);
}

}  // extern "C"
