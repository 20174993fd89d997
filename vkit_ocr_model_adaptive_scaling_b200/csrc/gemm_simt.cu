// fp32-accumulate SIMT GEMM with the same operand/epilogue semantics as the tcgen05 kernel (gemm_tc.cu).
// It is the GEMM of the library's fp32 mode (reference parity at rel 1e-4 needs full fp32 products, which the
// tensor cores' bf16/tf32 inputs cannot give) and the on-device cross-check for the tcgen05 path in tests.
#include "gemm_common.cuh"

namespace {

constexpr int TM = 64, TN_ = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
simt_nt_kernel(const T* __restrict__ x, VkocrConvGeom g, const T* __restrict__ w, int N, VkocrEpilogue ep) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN_ + 4];
    const long long M = (long long)g.batch * g.H * g.W;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN_;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int taps = g.ks * g.ks, half = g.ks >> 1;
    const long long kw = (long long)taps * g.c_pad;
    float acc[4][4] = {};

    // each thread gathers 4 A and 4 B elements per K step; (row, k) with k fastest for coalescing
    int a_row[4], a_k[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * 256;
        a_row[e] = idx >> 4;
        a_k[e] = idx & 15;
    }
    // decode pixel coordinates of this thread's A rows once
    int py[4], px[4], pb[4];
    bool pvalid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long m = m0 + a_row[e];
        pvalid[e] = m < M;
        const long long mm = pvalid[e] ? m : 0;
        px[e] = (int)(mm % g.W);
        py[e] = (int)((mm / g.W) % g.H);
        pb[e] = (int)(mm / ((long long)g.W * g.H));
    }

    for (int tap = 0; tap < taps; ++tap) {
        const int dy = tap / g.ks - half, dx = tap % g.ks - half;
        for (int c0 = 0; c0 < g.C; c0 += TK) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c0 + a_k[e];
                float av = 0.f;
                const int yy = py[e] + dy, xx = px[e] + dx;
                if (pvalid[e] && c < g.C && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W)
                    av = vk_to_f32(x[(((long long)pb[e] * g.H + yy) * g.W + xx) * g.ld_x + c]);
                As[a_k[e]][a_row[e]] = av;
                const int n = n0 + a_row[e];
                float bv = 0.f;
                if (n < N && c < g.C) bv = vk_to_f32(w[(long long)n * kw + (long long)tap * g.c_pad + c]);
                Bs[a_k[e]][a_row[e]] = bv;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            const float v = vk_epilogue_value<T>(ep, m, n, acc[i][j]);
            vk_epilogue_store<T>(ep, m, n, v);
        }
    }
}

// G[tap, i, j] += sum_{pix in split} P[pix, i] * Q[pix + off(tap), j]
template <typename T>
__global__ void __launch_bounds__(256)
simt_tn_kernel(const T* __restrict__ pm, VkocrConvGeom g, const T* __restrict__ qm, int J, long long ld_q, int splits,
               VkocrEpilogue ep) {
    __shared__ float Ps[TK][TM + 4];
    __shared__ float Qs[TK][TN_ + 4];
    const int I = g.C;
    const long long M = (long long)g.batch * g.H * g.W;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN_;
    const int tap = blockIdx.z / splits, sp = blockIdx.z % splits;
    const int half = g.ks >> 1;
    const int dy = tap / g.ks - half, dx = tap % g.ks - half;
    const long long chunk = ((M + splits - 1) / splits + TK - 1) / TK * TK;
    const long long p_begin = chunk * sp;
    long long p_end = p_begin + chunk;
    if (p_end > M) p_end = M;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4] = {};
    for (long long p0 = p_begin; p0 < p_end; p0 += TK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            const int kk = idx >> 6, col = idx & 63;   // channel fastest
            const long long pix = p0 + kk;
            float pv = 0.f, qv = 0.f;
            if (pix < p_end) {
                if (i0 + col < I) pv = vk_to_f32(pm[pix * g.ld_x + i0 + col]);
                if (j0 + col < J) {
                    const int xq = (int)(pix % g.W) + dx;
                    const int yq = (int)((pix / g.W) % g.H) + dy;
                    if (xq >= 0 && xq < g.W && yq >= 0 && yq < g.H) {
                        const long long b = pix / ((long long)g.W * g.H);
                        qv = vk_to_f32(qm[((b * g.H + yq) * g.W + xq) * ld_q + j0 + col]);
                    }
                }
            }
            Ps[kk][col] = pv;
            Qs[kk][col] = qv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ps[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Qs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ii = i0 + ty * 4 + i;
        if (ii >= I) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tx * 4 + j;
            if (jj >= J) continue;
            vk_epilogue_store_tn(ep, tap, ii, jj, acc[i][j]);
        }
    }
}

}  // namespace

int vkocr_gemm_simt_nt(int dtype, const void* x, const VkocrConvGeom* g, const void* w_packed, int N, const VkocrEpilogue* ep,
                       cudaStream_t stream) {
    const long long M = (long long)g->batch * g->H * g->W;
    if (M == 0 || N == 0) return VKOCR_OK;
    dim3 grid((unsigned)vk_cdiv(M, TM), (unsigned)vk_cdiv(N, TN_));
    VK_REQUIRE(grid.y <= 65535, VKOCR_BAD_SHAPE, "gemm_simt_nt: N too large");
    VK_DISPATCH_DTYPE(dtype, T, (simt_nt_kernel<T><<<grid, 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(x), *g, reinterpret_cast<const T*>(w_packed), N, *ep)));
    VK_CHECK_LAUNCH("simt_nt_kernel");
    return VKOCR_OK;
}

int vkocr_gemm_simt_tn(int dtype, const void* pmat, const VkocrConvGeom* g, const void* qmat, int J, long long ld_q,
                       const VkocrEpilogue* ep, cudaStream_t stream) {
    VK_REQUIRE(ep->out_f32 && ep->accumulate, VKOCR_BAD_ARGUMENT, "gemm_simt_tn: output must be fp32 accumulate");
    const long long M = (long long)g->batch * g->H * g->W;
    if (M == 0 || J == 0 || g->C == 0) return VKOCR_OK;
    const int taps = g->ks * g->ks;
    const long long tiles = (long long)vk_cdiv(J, TN_) * vk_cdiv(g->C, TM) * taps;
    long long splits = (4LL * vkocr_sm_count() + tiles - 1) / tiles;
    if (splits > (M + 255) / 256) splits = (M + 255) / 256;
    if (splits < 1) splits = 1;
    if (splits * taps > 65535) splits = 65535 / taps;
    dim3 grid((unsigned)vk_cdiv(J, TN_), (unsigned)vk_cdiv(g->C, TM), (unsigned)(taps * splits));
    VK_DISPATCH_DTYPE(dtype, T, (simt_tn_kernel<T><<<grid, 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(pmat), *g, reinterpret_cast<const T*>(qmat), J, ld_q,
                                    (int)splits, *ep)));
    VK_CHECK_LAUNCH("simt_tn_kernel");
    return VKOCR_OK;
}
