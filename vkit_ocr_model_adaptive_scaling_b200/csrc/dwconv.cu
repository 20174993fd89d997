// Depthwise 7x7 convolution on NHWC activations (reference helper.dconv7x7, model/helper.py:61-73: groups=C, pad 3,
// stride 1, bias), forward, data-gradient and weight-gradient.
//
// HBM-bound.  Forward/data-gradient: each thread produces a 4-pixel strip x one 16-byte channel vector, sweeping the
// 7 input rows once (10 vector loads per row instead of 28), taps in a [49][C] fp32 table.  The data-gradient is the
// same kernel with the taps mirrored (the packer flips them) and an optional `add` operand that fuses the residual
// gradient of ConvNextBlockLayer (convnext.py:56-58).
#include "common.cuh"

namespace {

constexpr int STRIP = 4;

template <typename T>
__global__ void __launch_bounds__(256)
dwconv7_generic_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, int B, int H, int W, int C,
               const float* __restrict__ wt, const float* __restrict__ bias, const T* __restrict__ add, long long ld_add) {
    constexpr int V = VkVec<T>::N;
    const int CV = C / V;
    const int XS = (W + STRIP - 1) / STRIP;
    const long long total = (long long)B * H * XS * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int xs = (int)(r % XS);
    r /= XS;
    const int yy0 = (int)(r % H);
    const int b = (int)(r / H);
    const int c = cv * V;
    const int x0 = xs * STRIP;

    float acc[STRIP][V];
#pragma unroll
    for (int o = 0; o < STRIP; ++o)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[o][i] = bias ? __ldg(bias + c + i) : 0.f;

    for (int ky = 0; ky < 7; ++ky) {
        const int yy = yy0 + ky - 3;
        if (yy < 0 || yy >= H) continue;
        const T* row = x + ((long long)b * H + yy) * W * ld_x + c;
        VkVec<T> in[STRIP + 6];
#pragma unroll
        for (int j = 0; j < STRIP + 6; ++j) {
            const int xx = x0 + j - 3;
            if (xx >= 0 && xx < W) in[j].load(row + (long long)xx * ld_x);
            else {
                float z[V] = {};
                in[j].pack(z);
            }
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            float w[V];
#pragma unroll
            for (int i = 0; i < V; i += 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wt + (long long)(ky * 7 + kx) * C + c + i));
                w[i] = w4.x; w[i + 1] = w4.y; w[i + 2] = w4.z; w[i + 3] = w4.w;
            }
#pragma unroll
            for (int o = 0; o < STRIP; ++o) {
                float f[V];
                in[o + kx].unpack(f);
#pragma unroll
                for (int i = 0; i < V; ++i) acc[o][i] = fmaf(w[i], f[i], acc[o][i]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < STRIP; ++o) {
        const int xx = x0 + o;
        if (xx >= W) break;
        const long long pix = ((long long)b * H + yy0) * W + xx;
        if (add) {
            VkVec<T> a;
            a.load(add + pix * ld_add + c);
            float f[V];
            a.unpack(f);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[o][i] += f[i];
        }
        VkVec<T> out;
        out.pack(acc[o]);
        out.store(y + pix * ld_y + c);
    }
}

// dW[c][ky*7+kx] += sum_{b,y,x} dy[b,y,x,c] * x[b,y+ky-3,x+kx-3,c]   (the (C,1,7,7) layout of the parameter)
// One thread = one 16-byte channel vector x one kernel row ky (blockIdx.y) x one SEG-pixel segment of an image row: it
// loads the SEG+6 input vectors of the segment (halo included) and the SEG gradient vectors once and does 7*SEG*V FMAs
// on them (each x vector is reused by up to 7 taps from registers), keeping acc[7 kx][V] in registers over all rows the
// block visits.  Partials are merged per block in shared memory, then one atomic per (channel, tap) and block.
template <typename T, int SEG>
__global__ void __launch_bounds__(256, 2)
dwconv7_wgrad_generic_kernel(const T* __restrict__ dy, long long ld_dy, const T* __restrict__ x, long long ld_x, int B, int H, int W,
                     int C, int cv_per_block, float* __restrict__ dw) {
    using VT = VkVec4<T>;
    constexpr int V = VT::N;
    extern __shared__ float red[];  // [7][cv_per_block * V]
    const int CV = C / V;
    const int lane_cv = threadIdx.x % cv_per_block;
    const int sg = threadIdx.x / cv_per_block;
    const int NSG = blockDim.x / cv_per_block;
    const int cv = blockIdx.z * cv_per_block + lane_cv;
    const int ky = blockIdx.y;
    const bool active = cv < CV && sg < NSG;
    const int c = cv * V;
    const int segs = (W + SEG - 1) / SEG;
    float acc[7][V];
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[k][i] = 0.f;
    for (int i = threadIdx.x; i < 7 * cv_per_block * V; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    if (active) {
        const long long rows = (long long)B * H;
        for (long long rr = blockIdx.x; rr < rows; rr += gridDim.x) {
            const int yy = (int)(rr % H);
            const int b = (int)(rr / H);
            const int ys = yy + ky - 3;
            if (ys < 0 || ys >= H) continue;
            const T* dyrow = dy + ((long long)b * H + yy) * W * ld_dy + c;
            const T* xrow = x + ((long long)b * H + ys) * W * ld_x + c;
            for (int s = sg; s < segs; s += NSG) {
                const int x0 = s * SEG;
                VT in[SEG + 6], gd[SEG];
#pragma unroll
                for (int j = 0; j < SEG + 6; ++j) {
                    const int xx = x0 + j - 3;
                    if (xx >= 0 && xx < W) in[j].load(xrow + (long long)xx * ld_x);
                    else {
                        float z[V] = {};
                        in[j].pack(z);
                    }
                }
#pragma unroll
                for (int o = 0; o < SEG; ++o) {
                    const int xx = x0 + o;
                    if (xx < W) gd[o].load(dyrow + (long long)xx * ld_dy);
                    else {
                        float z[V] = {};
                        gd[o].pack(z);
                    }
                }
#pragma unroll
                for (int o = 0; o < SEG; ++o) {
                    float fd[V];
                    gd[o].unpack(fd);
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {
                        float fx[V];
                        in[o + kx].unpack(fx);
#pragma unroll
                        for (int i = 0; i < V; ++i) acc[kx][i] = fmaf(fd[i], fx[i], acc[kx][i]);
                    }
                }
            }
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
            for (int i = 0; i < V; ++i) atomicAdd(&red[(kx * cv_per_block + lane_cv) * V + i], acc[kx][i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 7 * cv_per_block * V; i += blockDim.x) {
        const int kx = i / (cv_per_block * V);
        const int rem = i % (cv_per_block * V);
        const int cc = blockIdx.z * cv_per_block * V + rem;
        if (cc < C) atomicAdd(dw + (long long)cc * 49 + (ky * 7 + kx), red[i]);
    }
}


// ------------------------------------------------------------------------------------------------- tiled bf16 kernels
// Shared-memory tiled kernels for bf16 storage with C % 32 == 0 (every hot-path shape).  A block is persistent over the
// TW x 8 pixel tiles of one 32-channel block: the (TW+6) x (8+6) halo tile is brought in with 16-byte cp.async copies
// (out-of-image pixels are zero-filled by the copy itself = the conv padding) into one of two buffers while the other is
// being consumed, and all arithmetic reads shared memory.  The kernels are fp32-FMA-pipe-bound (49 FMAs per output element
// against 4 bytes of HBM traffic), so the inner loops are arranged for FMA density: each thread owns 4 channels x a 4 x 2
// pixel patch (forward / data gradient) or 4 channels x one kernel row (weight gradient), converts every bf16 input once
// per use row and issues the arithmetic as packed FFMA2 (two channels per instruction) so that conversions, LDS and
// address arithmetic fit into the issue slots the FMA pipe leaves free.  TW is 40 for the 160/80/40-pixel-wide maps of 640 x 640 training (exact tiling), 20 for 20, 32 otherwise.
constexpr int TH = 8, CB = 32;
constexpr int PIX_B = CB * 2;                       // bytes per pixel in the tile
constexpr int HALO_H = TH + 6;
// row strides: forward -> rows 2 apart land 64 B apart (mod 128), weight gradient -> consecutive rows do: conflict-free LDS.64
__host__ __device__ constexpr int rs_fwd(int tw) { return ((tw + 6) * PIX_B) % 128 == 0 ? (tw + 6) * PIX_B + 32 : (tw + 6) * PIX_B + 96; }
__host__ __device__ constexpr int rs_wg(int tw) { return ((tw + 6) * PIX_B) % 128 == 0 ? (tw + 6) * PIX_B + 64 : (tw + 6) * PIX_B; }

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__device__ __forceinline__ void unpack4(uint2 raw, float* f) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float* f) {
    uint2 raw;
    *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(f[0], f[1]);
    *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(f[2], f[3]);
    return raw;
}

struct TileCoord { int b, y0, x0; };
template <int TW>
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_x, int tiles_y) {
    TileCoord tc;
    const int tx = t % tiles_x;
    t /= tiles_x;
    tc.x0 = tx * TW;
    tc.y0 = (t % tiles_y) * TH;
    tc.b = t / tiles_y;
    return tc;
}

// Packed fp32 FMA (sm_100 FFMA2): d = a * b + d on two lanes.  Same FMA-pipe throughput as two FFMAs but ONE issue slot,
// which is what these kernels are short of (conversions, LDS and address arithmetic share the scheduler with the FMAs).
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(dd)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
// two bf16 channels in one 32-bit word -> an fp32 pair (one shift, one mask; a PRMT for the low half instead of the shift,
// which ptxas turns into IMAD.U32 on the FMA pipe, measured no faster: the kernels are not bound by that pipe's throughput)
__device__ __forceinline__ float2 bf2_to_f2(uint32_t raw) {
    return make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
}

// Column-wise halo loader: thread -> one 16-byte quarter of one halo column, walking down the HALO_H rows with a
// constant address increment (no per-copy index arithmetic).  Threads beyond (TW + 6) * 4 idle.
template <int TW, int RS>
__device__ __forceinline__ void load_halo_cols(uint32_t smem, const __nv_bfloat16* __restrict__ x, long long ld_x, TileCoord tc,
                                               int H, int W, int c0) {
    constexpr int HW = TW + 6;
    if (threadIdx.x < HW * 4) {
        const int part = threadIdx.x & 3, hx = threadIdx.x >> 2;
        const int xx = tc.x0 + hx - 3;
        const bool col_ok = xx >= 0 && xx < W;
        const long long row_step = (long long)W * ld_x;
        const __nv_bfloat16* src = x + (((long long)tc.b * H + (tc.y0 - 3)) * W + (col_ok ? xx : 0)) * ld_x + c0 + part * 8;
        uint32_t dst = smem + (uint32_t)(hx * PIX_B + part * 16);
#pragma unroll
        for (int hy = 0; hy < HALO_H; ++hy) {
            const int yy = tc.y0 + hy - 3;
            const bool ok = col_ok && yy >= 0 && yy < H;
            cp_async16_zfill(dst, ok ? src : x, ok);
            src += row_step;
            dst += RS;
        }
    }
}

// Forward / data gradient.  A block is persistent over the TW x 8 tiles of one 32-channel block and double-buffers the
// halo tile: the cp.async copies of tile t+1 are in flight while tile t is computed from shared memory.
template <int TW>
__global__ void __launch_bounds__(8 * TW, 2)
dwconv7_tile_kernel(const __nv_bfloat16* __restrict__ x, long long ld_x, __nv_bfloat16* __restrict__ y, long long ld_y, int B,
                    int H, int W, int C, const float* __restrict__ wt, const float* __restrict__ bias,
                    const __nv_bfloat16* __restrict__ add, long long ld_add, int tiles_x, int tiles_y, int ntiles) {
    constexpr int RS = rs_fwd(TW);
    constexpr int BUF = HALO_H * RS;
    extern __shared__ __align__(128) uint8_t dw_smem[];
    uint8_t* s_in = dw_smem;                                             // [2][HALO_H][RS]
    float* s_w = reinterpret_cast<float*>(dw_smem + 2 * BUF);            // [49][CB]
    float* s_b = s_w + 49 * CB;                                          // [CB]
    const int c0 = blockIdx.y * CB;
    const uint32_t s_in_a = (uint32_t)__cvta_generic_to_shared(s_in);
    int t = blockIdx.x;
    if (t < ntiles) load_halo_cols<TW, RS>(s_in_a, x, ld_x, tile_coord<TW>(t, tiles_x, tiles_y), H, W, c0);
    vk_cp_async_commit();
    for (int i = threadIdx.x; i < 49 * CB; i += blockDim.x) s_w[i] = __ldg(wt + (long long)(i / CB) * C + c0 + (i % CB));
    if (threadIdx.x < CB) s_b[threadIdx.x] = bias ? __ldg(bias + c0 + threadIdx.x) : 0.f;

    // thread -> 4 channels (cq) x a patch of 4 (x) x 2 (y) pixels; the 4 patches of a warp are stacked in y
    const int cq = threadIdx.x & 7;
    const int sidx = threadIdx.x >> 3;          // 0 .. TW - 1
    const int sy = sidx & 3, sx = sidx >> 2;    // 4 patches in y (8 rows), TW / 4 in x
    const int px = sx * 4, py = sy * 2;
    int buf = 0;
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        const TileCoord tc = tile_coord<TW>(t, tiles_x, tiles_y);
        vk_cp_async_wait<0>();
        __syncthreads();                        // tile t has landed; everybody is done with the other buffer
        if (t + (int)gridDim.x < ntiles)
            load_halo_cols<TW, RS>(s_in_a + (uint32_t)((buf ^ 1) * BUF), x, ld_x, tile_coord<TW>(t + gridDim.x, tiles_x, tiles_y), H, W, c0);
        vk_cp_async_commit();
        float2 acc[2][4][2];
        {
            const float4 bv = *reinterpret_cast<const float4*>(s_b + cq * 4);
#pragma unroll
            for (int oy = 0; oy < 2; ++oy)
#pragma unroll
                for (int o = 0; o < 4; ++o) { acc[oy][o][0] = make_float2(bv.x, bv.y); acc[oy][o][1] = make_float2(bv.z, bv.w); }
        }
        // Halo row py + r feeds output row oy with kernel row ky = r - oy: rows 1..6 feed both output rows, row 0 only
        // the upper and row 7 only the lower one.  The middle rows run as a ROLLED loop: the fully unrolled body (2500
        // instructions, 40 KB) streamed through the instruction cache and left the schedulers half idle.
        const uint8_t* tile = s_in + buf * BUF + py * RS + px * PIX_B + cq * 8;
        const float* wrow = s_w + cq * 4;
        auto load_row = [&](const uint8_t* rowp, float2 (&in)[10][2]) {
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const uint2 raw = *reinterpret_cast<const uint2*>(rowp + j * PIX_B);
                in[j][0] = bf2_to_f2(raw.x);
                in[j][1] = bf2_to_f2(raw.y);
            }
        };
        auto taps = [&](const float* wk, const float2 (&in)[10][2], float2 (&a)[4][2]) {
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4 w = *reinterpret_cast<const float4*>(wk + kx * CB);
                const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    ffma2(a[o][0], w0, in[o + kx][0]);
                    ffma2(a[o][1], w1, in[o + kx][1]);
                }
            }
        };
        {
            float2 in[10][2];
            load_row(tile, in);
            taps(wrow, in, acc[0]);                                   // r = 0: ky = 0 of the upper row
        }
#pragma unroll 1
        for (int r = 1; r < 7; ++r) {
            float2 in[10][2];
            load_row(tile + r * RS, in);
            taps(wrow + r * 7 * CB, in, acc[0]);                      // ky = r
            taps(wrow + (r - 1) * 7 * CB, in, acc[1]);                // ky = r - 1
        }
        {
            float2 in[10][2];
            load_row(tile + 7 * RS, in);
            taps(wrow + 6 * 7 * CB, in, acc[1]);                      // r = 7: ky = 6 of the lower row
        }
        uint2 av[2][4];
        if (add) {
#pragma unroll
            for (int oy = 0; oy < 2; ++oy)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int yy = tc.y0 + py + oy, xx = tc.x0 + px + o;
                    av[oy][o] = make_uint2(0u, 0u);
                    if (yy < H && xx < W)
                        av[oy][o] = __ldg(reinterpret_cast<const uint2*>(add + (((long long)tc.b * H + yy) * W + xx) * ld_add + c0 + cq * 4));
                }
        }
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
            const int yy = tc.y0 + py + oy;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int xx = tc.x0 + px + o;
                if (yy < H && xx < W) {
                    float2 a0 = acc[oy][o][0], a1 = acc[oy][o][1];
                    if (add) {
                        const float2 f0 = bf2_to_f2(av[oy][o].x), f1 = bf2_to_f2(av[oy][o].y);
                        a0.x += f0.x; a0.y += f0.y; a1.x += f1.x; a1.y += f1.y;
                    }
                    uint2 raw;
                    *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(a0.x, a0.y);
                    *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(a1.x, a1.y);
                    *reinterpret_cast<uint2*>(y + (((long long)tc.b * H + yy) * W + xx) * ld_y + c0 + cq * 4) = raw;
                }
            }
        }
    }
}

// Forward / data gradient, second arrangement (TW % 8 == 0): thread -> 4 channels x an 8 x 1 pixel strip.  The 4 x 2 patch
// above reads every tap vector twice (once per output row) and delivers 0.70 shared-memory wavefronts per FFMA2 -- the LSU
// data pipe (one 128-byte wavefront per clock and SM; a 16-byte-per-lane tap load costs 4 of them however much of it is a
// broadcast) runs at 1.2x the FMA pipe's time and bounds the kernel.  Here the 7 tap vectors of a kernel row stay in
// registers for the whole strip and each of the 14 input pixels of the row is loaded and converted once, then applied to
// every output it reaches:  acc[o] += w[j - o] * in[j]  -- 0.50 wavefronts per FFMA2, below the FMA pipe's time.
__host__ __device__ constexpr int rs_strip(int tw) { return ((tw + 6) * PIX_B) % 128 == 0 ? (tw + 6) * PIX_B + 64 : (tw + 6) * PIX_B; }
template <int TW>
__global__ void __launch_bounds__(8 * TW, 2)
dwconv7_strip_kernel(const __nv_bfloat16* __restrict__ x, long long ld_x, __nv_bfloat16* __restrict__ y, long long ld_y, int B,
                     int H, int W, int C, const float* __restrict__ wt, const float* __restrict__ bias,
                     const __nv_bfloat16* __restrict__ add, long long ld_add, int tiles_x, int tiles_y, int ntiles) {
    static_assert(TW % 8 == 0, "8-pixel strips");
    constexpr int RS = rs_strip(TW);             // consecutive rows land 64 B apart (mod 128): the 4 strips of a warp are stacked in y
    constexpr int BUF = HALO_H * RS;
    extern __shared__ __align__(128) uint8_t dw_smem[];
    uint8_t* s_in = dw_smem;                                             // [2][HALO_H][RS]
    float* s_w = reinterpret_cast<float*>(dw_smem + 2 * BUF);            // [49][CB]
    float* s_b = s_w + 49 * CB;                                          // [CB]
    const int c0 = blockIdx.y * CB;
    const uint32_t s_in_a = (uint32_t)__cvta_generic_to_shared(s_in);
    int t = blockIdx.x;
    if (t < ntiles) load_halo_cols<TW, RS>(s_in_a, x, ld_x, tile_coord<TW>(t, tiles_x, tiles_y), H, W, c0);
    vk_cp_async_commit();
    for (int i = threadIdx.x; i < 49 * CB; i += blockDim.x) s_w[i] = __ldg(wt + (long long)(i / CB) * C + c0 + (i % CB));
    if (threadIdx.x < CB) s_b[threadIdx.x] = bias ? __ldg(bias + c0 + threadIdx.x) : 0.f;
    const int cq = threadIdx.x & 7;
    const int sidx = threadIdx.x >> 3;          // 0 .. TW - 1
    const int py = sidx & 7, px = (sidx >> 3) * 8;
    int buf = 0;
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        const TileCoord tc = tile_coord<TW>(t, tiles_x, tiles_y);
        vk_cp_async_wait<0>();
        __syncthreads();                        // tile t has landed; everybody is done with the other buffer
        if (t + (int)gridDim.x < ntiles)
            load_halo_cols<TW, RS>(s_in_a + (uint32_t)((buf ^ 1) * BUF), x, ld_x, tile_coord<TW>(t + gridDim.x, tiles_x, tiles_y), H, W, c0);
        vk_cp_async_commit();
        float2 acc[8][2];
        {
            const float4 bv = *reinterpret_cast<const float4*>(s_b + cq * 4);
#pragma unroll
            for (int o = 0; o < 8; ++o) { acc[o][0] = make_float2(bv.x, bv.y); acc[o][1] = make_float2(bv.z, bv.w); }
        }
        const uint8_t* rowp = s_in + buf * BUF + py * RS + px * PIX_B + cq * 8;
        const float* wk = s_w + cq * 4;
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky, rowp += RS, wk += 7 * CB) {
            float2 w[7][2];
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4 w4 = *reinterpret_cast<const float4*>(wk + kx * CB);
                w[kx][0] = make_float2(w4.x, w4.y);
                w[kx][1] = make_float2(w4.z, w4.w);
            }
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                const uint2 raw = *reinterpret_cast<const uint2*>(rowp + j * PIX_B);
                const float2 i0 = bf2_to_f2(raw.x), i1 = bf2_to_f2(raw.y);
#pragma unroll
                for (int o = (j > 6 ? j - 6 : 0); o <= (j < 7 ? j : 7); ++o) {
                    ffma2(acc[o][0], w[j - o][0], i0);
                    ffma2(acc[o][1], w[j - o][1], i1);
                }
            }
        }
        const int yy = tc.y0 + py;
        if (yy < H) {
            const long long pix0 = ((long long)tc.b * H + yy) * W + tc.x0 + px;
            __nv_bfloat16* yp = y + pix0 * ld_y + c0 + cq * 4;
            const __nv_bfloat16* ap = add ? add + pix0 * ld_add + c0 + cq * 4 : nullptr;
            uint2 av[8];
            if (add) {
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    av[o] = make_uint2(0u, 0u);
                    if (tc.x0 + px + o < W) av[o] = __ldg(reinterpret_cast<const uint2*>(ap + (long long)o * ld_add));
                }
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (tc.x0 + px + o < W) {
                    float2 a0 = acc[o][0], a1 = acc[o][1];
                    if (add) {
                        const float2 f0 = bf2_to_f2(av[o].x), f1 = bf2_to_f2(av[o].y);
                        a0.x += f0.x; a0.y += f0.y; a1.x += f1.x; a1.y += f1.y;
                    }
                    uint2 raw;
                    *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(a0.x, a0.y);
                    *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(a1.x, a1.y);
                    *reinterpret_cast<uint2*>(yp + (long long)o * ld_y) = raw;
                }
            }
        }
    }
}

// Third arrangement: 8 x 2 strips.  The 8 x 1 strip kernel above runs the LSU data pipe (0.50 shared-memory wavefronts per
// FFMA2) and the conversions (two ALU / FMA-pipe integer operations per loaded word -- the `<< 16` is an IMAD.U32 on the
// very pipe the FFMA2s need: ncu has 8 % of all warp instructions there, the heavy FMA pipe 65 % busy with FFMA2 alone
// accounting for 44 points of it) at the FMA pipe's own time.  A thread that owns TWO output rows of its strip applies every
// loaded and converted input pixel to both (kernel row r for the upper output row, r - 1 for the lower one): 8 input rows
// instead of 14 per pair of output rows -- 0.29 wavefronts and 0.25 conversions per FFMA2.  The taps of two kernel rows live
// in registers (wa / wb alternate roles over an unrolled-by-two row loop, so nothing is moved); tiles are TW x TH2 with
// TH2 = 16 (8 for maps whose height is no multiple of 16).
template <int TH2> struct Strip2 { static constexpr int HH = TH2 + 6; };
template <int TW, int TH2>
__device__ __forceinline__ TileCoord tile_coord2(int t, int tiles_x, int tiles_y) {
    TileCoord tc;
    const int tx = t % tiles_x;
    t /= tiles_x;
    tc.x0 = tx * TW;
    tc.y0 = (t % tiles_y) * TH2;
    tc.b = t / tiles_y;
    return tc;
}
template <int TW, int RS, int HH>
__device__ __forceinline__ void load_halo_cols2(uint32_t smem, const __nv_bfloat16* __restrict__ x, long long ld_x, TileCoord tc,
                                                int H, int W, int c0) {
    constexpr int HW = TW + 6;
    for (int q = threadIdx.x; q < HW * 4; q += blockDim.x) {
        const int part = q & 3, hx = q >> 2;
        const int xx = tc.x0 + hx - 3;
        const bool col_ok = xx >= 0 && xx < W;
        const long long row_step = (long long)W * ld_x;
        const __nv_bfloat16* src = x + (((long long)tc.b * H + (tc.y0 - 3)) * W + (col_ok ? xx : 0)) * ld_x + c0 + part * 8;
        uint32_t dst = smem + (uint32_t)(hx * PIX_B + part * 16);
#pragma unroll
        for (int hy = 0; hy < HH; ++hy) {
            const int yy = tc.y0 + hy - 3;
            const bool ok = col_ok && yy >= 0 && yy < H;
            cp_async16_zfill(dst, ok ? src : x, ok);
            src += row_step;
            dst += RS;
        }
    }
}
template <int TW, int TH2>
__global__ void __launch_bounds__(TW * TH2 / 2, (TW * TH2 / 2 <= 128) ? 3 : 2)
dwconv7_strip2_kernel(const __nv_bfloat16* __restrict__ x, long long ld_x, __nv_bfloat16* __restrict__ y, long long ld_y, int B,
                      int H, int W, int C, const float* __restrict__ wt, const float* __restrict__ bias,
                      const __nv_bfloat16* __restrict__ add, long long ld_add, int tiles_x, int tiles_y, int ntiles) {
    static_assert(TW % 8 == 0 && TH2 % 2 == 0, "8 x 2 strips");
    constexpr int HH = TH2 + 6;
    constexpr int RS = rs_fwd(TW);               // rows 2 apart land 64 B apart (mod 128): the 4 strips of a warp are stacked in y
    constexpr int BUF = HH * RS;
    constexpr int NR = TH2 / 2;                  // row pairs per tile
    extern __shared__ __align__(128) uint8_t dw_smem[];
    uint8_t* s_in = dw_smem;                                             // [2][HH][RS]
    float* s_w = reinterpret_cast<float*>(dw_smem + 2 * BUF);            // [49][CB]
    float* s_b = s_w + 49 * CB;                                          // [CB]
    const int c0 = blockIdx.y * CB;
    const uint32_t s_in_a = (uint32_t)__cvta_generic_to_shared(s_in);
    int t = blockIdx.x;
    if (t < ntiles) load_halo_cols2<TW, RS, HH>(s_in_a, x, ld_x, tile_coord2<TW, TH2>(t, tiles_x, tiles_y), H, W, c0);
    vk_cp_async_commit();
    for (int i = threadIdx.x; i < 49 * CB; i += blockDim.x) s_w[i] = __ldg(wt + (long long)(i / CB) * C + c0 + (i % CB));
    if (threadIdx.x < CB) s_b[threadIdx.x] = bias ? __ldg(bias + c0 + threadIdx.x) : 0.f;
    const int cq = threadIdx.x & 7;
    const int sidx = threadIdx.x >> 3;
    const int pr = sidx % NR, px = (sidx / NR) * 8;
    int buf = 0;
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        const TileCoord tc = tile_coord2<TW, TH2>(t, tiles_x, tiles_y);
        vk_cp_async_wait<0>();
        __syncthreads();                        // tile t has landed; everybody is done with the other buffer
        if (t + (int)gridDim.x < ntiles)
            load_halo_cols2<TW, RS, HH>(s_in_a + (uint32_t)((buf ^ 1) * BUF), x, ld_x, tile_coord2<TW, TH2>(t + gridDim.x, tiles_x, tiles_y), H, W, c0);
        vk_cp_async_commit();
        float2 acc[2][8][2];
        {
            const float4 bv = *reinterpret_cast<const float4*>(s_b + cq * 4);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int o = 0; o < 8; ++o) { acc[r][o][0] = make_float2(bv.x, bv.y); acc[r][o][1] = make_float2(bv.z, bv.w); }
        }
        const uint8_t* rowp = s_in + buf * BUF + (2 * pr) * RS + px * PIX_B + cq * 8;
        const float* wk = s_w + cq * 4;
        float2 wa[7][2], wb[7][2];
        auto taps = [&](float2 (&w)[7][2]) {           // the 7 taps of the kernel row wk points at, then on to the next row
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4 w4 = *reinterpret_cast<const float4*>(wk + kx * CB);
                w[kx][0] = make_float2(w4.x, w4.y);
                w[kx][1] = make_float2(w4.z, w4.w);
            }
            wk += 7 * CB;
        };
        // one input row: `up` (nullable at compile time) weighs it into the upper output row, `lo` into the lower one
        auto row = [&](auto has_up, auto has_lo, const float2 (&up)[7][2], const float2 (&lo)[7][2]) {
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                const uint2 raw = *reinterpret_cast<const uint2*>(rowp + j * PIX_B);
                const float2 i0 = bf2_to_f2(raw.x), i1 = bf2_to_f2(raw.y);
#pragma unroll
                for (int o = (j > 6 ? j - 6 : 0); o <= (j < 7 ? j : 7); ++o) {
                    if constexpr (decltype(has_up)::value) {
                        ffma2(acc[0][o][0], up[j - o][0], i0);
                        ffma2(acc[0][o][1], up[j - o][1], i1);
                    }
                    if constexpr (decltype(has_lo)::value) {
                        ffma2(acc[1][o][0], lo[j - o][0], i0);
                        ffma2(acc[1][o][1], lo[j - o][1], i1);
                    }
                }
            }
            rowp += RS;
        };
        using Y = std::true_type;
        using N = std::false_type;
        taps(wa);                               // kernel row 0
        row(Y{}, N{}, wa, wa);                  // input row 0: upper output row only
#pragma unroll 1
        for (int r = 1; r < 7; r += 2) {
            taps(wb);                           // kernel row r
            row(Y{}, Y{}, wb, wa);              // input row r:     upper <- kernel row r,     lower <- kernel row r - 1
            taps(wa);                           // kernel row r + 1
            row(Y{}, Y{}, wa, wb);              // input row r + 1: upper <- kernel row r + 1, lower <- kernel row r
        }
        // The residual operand of the data gradient is read with plain global loads: the upper row's go out before the last
        // input row is applied (its accumulators are final, and wb's registers are free), the lower row's before the upper
        // row is converted and stored -- one exposed DRAM round trip per tile instead of two.
        const long long pix0 = ((long long)tc.b * H + tc.y0 + 2 * pr) * W + tc.x0 + px;
        const __nv_bfloat16* ap = add ? add + pix0 * ld_add + c0 + cq * 4 : nullptr;
        auto fetch = [&](int r, uint2 (&av)[8]) {
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                av[o] = make_uint2(0u, 0u);
                if (add && tc.y0 + 2 * pr + r < H && tc.x0 + px + o < W)
                    av[o] = __ldg(reinterpret_cast<const uint2*>(ap + ((long long)r * W + o) * ld_add));
            }
        };
        auto store = [&](int r, const uint2 (&av)[8]) {
            if (tc.y0 + 2 * pr + r >= H) return;
            __nv_bfloat16* yp = y + (pix0 + (long long)r * W) * ld_y + c0 + cq * 4;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (tc.x0 + px + o < W) {
                    const float2 f0 = bf2_to_f2(av[o].x), f1 = bf2_to_f2(av[o].y);       // zeros without the operand
                    const float2 a0 = acc[r][o][0], a1 = acc[r][o][1];
                    uint2 raw;
                    *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(a0.x + f0.x, a0.y + f0.y);
                    *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(a1.x + f1.x, a1.y + f1.y);
                    *reinterpret_cast<uint2*>(yp + (long long)o * ld_y) = raw;
                }
            }
        };
        uint2 av0[8], av1[8];
        fetch(0, av0);
        row(N{}, Y{}, wa, wa);                  // input row 7: lower output row only, kernel row 6
        fetch(1, av1);
        store(0, av0);
        store(1, av1);
    }
}

// Weight gradient.  Thread -> 4 channels (cq) x one kernel row ky x one tile row; it slides along x keeping the 7 input
// vectors of its window in registers: acc[kx][4] += dy[y][x][4] * in[y + ky][x + kx][4].  Persistent over the tiles of one
// 32-channel block with double-buffered (halo tile of x, tile of dy) stages: the copies of tile t+1 fly while tile t is
// consumed, one block barrier per tile.
constexpr int WG_THREADS = 8 * 7 * TH;          // 448
template <int TW>
__global__ void __launch_bounds__(WG_THREADS, 1)
dwconv7_wgrad_tile_kernel(const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ x,
                          long long ld_x, int B, int H, int W, int C, int tiles_x, int tiles_y, float* __restrict__ dw) {
    constexpr int RS = rs_wg(TW);
    constexpr int STAGE = HALO_H * RS + TH * TW * PIX_B;         // halo tile of x, then the dy tile
    constexpr int HW = TW + 6;
    static_assert(HW * 4 + TW * 4 <= WG_THREADS, "one loader thread per 16-byte column of either tile");
    extern __shared__ __align__(128) uint8_t dw_smem[];
    float* s_acc = reinterpret_cast<float*>(dw_smem + 2 * STAGE);   // [49][CB]
    const int c0 = blockIdx.x * CB;
    const int cq = threadIdx.x & 7;
    const int rest = threadIdx.x >> 3;          // 0..55
    const int ky = rest % 7, row = rest / 7;    // kernel row, tile row of dy
    float2 acc[7][2];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k][0] = acc[k][1] = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < 49 * CB; i += blockDim.x) s_acc[i] = 0.f;
    const int ntiles = B * tiles_y * tiles_x;
    const uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(dw_smem);
    // loaders: threads [0, HW*4) walk the halo columns of x, threads [HW*4, HW*4 + TW*4) the columns of dy
    auto issue = [&](int t, int st) {
        const TileCoord tc = tile_coord<TW>(t, tiles_x, tiles_y);
        const uint32_t base = smem_a + (uint32_t)(st * STAGE);
        load_halo_cols<TW, RS>(base, x, ld_x, tc, H, W, c0);
        const int q = (int)threadIdx.x - HW * 4;
        if (q >= 0 && q < TW * 4) {
            const int part = q & 3, px = q >> 2;
            const int xx = tc.x0 + px;
            const bool col_ok = xx < W;
            const long long row_step = (long long)W * ld_dy;
            const __nv_bfloat16* src = dy + (((long long)tc.b * H + tc.y0) * W + (col_ok ? xx : 0)) * ld_dy + c0 + part * 8;
            uint32_t dst = base + (uint32_t)(HALO_H * RS + px * PIX_B + part * 16);
#pragma unroll
            for (int r = 0; r < TH; ++r) {
                const bool ok = col_ok && tc.y0 + r < H;
                cp_async16_zfill(dst, ok ? src : dy, ok);
                src += row_step;
                dst += TW * PIX_B;
            }
        }
    };
    int t = blockIdx.y, st = 0;
    if (t < ntiles) issue(t, 0);
    vk_cp_async_commit();
    for (; t < ntiles; t += gridDim.y, st ^= 1) {
        vk_cp_async_wait<0>();
        __syncthreads();                        // tile t has landed; everybody is done with the other stage
        if (t + (int)gridDim.y < ntiles) issue(t + gridDim.y, st ^ 1);
        vk_cp_async_commit();
        const uint8_t* s_in = dw_smem + st * STAGE;
        const uint8_t* inrow = s_in + (row + ky) * RS + cq * 8;                       // halo row row + ky  (input row y + ky - 3)
        const uint8_t* dyrow = s_in + HALO_H * RS + row * TW * PIX_B + cq * 8;
        float2 win[7][2];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint2 raw = *reinterpret_cast<const uint2*>(inrow + k * PIX_B);
            win[k + 1][0] = bf2_to_f2(raw.x);
            win[k + 1][1] = bf2_to_f2(raw.y);
        }
#pragma unroll 8
        for (int xx = 0; xx < TW; ++xx) {
#pragma unroll
            for (int k = 0; k < 6; ++k) { win[k][0] = win[k + 1][0]; win[k][1] = win[k + 1][1]; }
            const uint2 raw = *reinterpret_cast<const uint2*>(inrow + (xx + 6) * PIX_B);
            win[6][0] = bf2_to_f2(raw.x);
            win[6][1] = bf2_to_f2(raw.y);
            const uint2 graw = *reinterpret_cast<const uint2*>(dyrow + xx * PIX_B);
            const float2 g0 = bf2_to_f2(graw.x), g1 = bf2_to_f2(graw.y);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                ffma2(acc[k][0], g0, win[k][0]);
                ffma2(acc[k][1], g1, win[k][1]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        atomicAdd(&s_acc[(ky * 7 + k) * CB + cq * 4 + 0], acc[k][0].x);
        atomicAdd(&s_acc[(ky * 7 + k) * CB + cq * 4 + 1], acc[k][0].y);
        atomicAdd(&s_acc[(ky * 7 + k) * CB + cq * 4 + 2], acc[k][1].x);
        atomicAdd(&s_acc[(ky * 7 + k) * CB + cq * 4 + 3], acc[k][1].y);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 49 * CB; i += blockDim.x) atomicAdd(dw + (long long)(c0 + i % CB) * 49 + i / CB, s_acc[i]);
}

inline int pick_tw(int W) {
    static const int forced = getenv("VKOCR_DW_TW") ? atoi(getenv("VKOCR_DW_TW")) : 0;      // experiments
    if (forced && W % forced == 0) return forced;
    return (W % 40 == 0) ? 40 : ((W % 20 == 0 && W < 40) ? 20 : 32);
}
// forward / data gradient: 16-pixel-wide strip tiles where they tile the map exactly (4 blocks of 4 warps per SM: the per-tile
// block barrier costs less with small blocks -- 160 x 160 x 96: 0.250 ms with 40-wide tiles, 0.234 with 16-wide ones)
inline int pick_tw_fwd(int W) {
    const int tw = pick_tw(W);
    static const bool forced = getenv("VKOCR_DW_TW") != nullptr;
    return (!forced && W % 16 == 0) ? 16 : tw;
}

template <int TW>
int launch_dw_strip(const void* x, long long ld_x, void* y, long long ld_y, int B, int H, int W, int C, const float* wt,
                    const float* bias, const void* add, long long ld_add, cudaStream_t s) {
    const int tiles_x = vk_cdiv(W, TW), tiles_y = vk_cdiv(H, TH);
    const long long ntiles = (long long)B * tiles_x * tiles_y;
    const int cblocks = C / CB;
    const int smem = 2 * HALO_H * rs_strip(TW) + (49 * CB + CB) * (int)sizeof(float);
    cudaFuncSetAttribute(dwconv7_strip_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwconv7_strip_kernel<TW>, 8 * TW, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    long long gx = ((long long)vkocr_sm_count() * per_sm) / cblocks;     // persistent: the resident block count, rounded DOWN
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dwconv7_strip_kernel<TW><<<dim3((unsigned)gx, (unsigned)cblocks), 8 * TW, smem, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ld_x, reinterpret_cast<__nv_bfloat16*>(y), ld_y, B, H, W, C, wt, bias,
        reinterpret_cast<const __nv_bfloat16*>(add), ld_add, tiles_x, tiles_y, (int)ntiles);
    return 0;
}

template <int TW, int TH2>
int launch_dw_strip2(const void* x, long long ld_x, void* y, long long ld_y, int B, int H, int W, int C, const float* wt,
                     const float* bias, const void* add, long long ld_add, cudaStream_t s) {
    const int tiles_x = vk_cdiv(W, TW), tiles_y = vk_cdiv(H, TH2);
    const long long ntiles = (long long)B * tiles_x * tiles_y;
    const int cblocks = C / CB;
    const int threads = TW * TH2 / 2;
    const int smem = 2 * (TH2 + 6) * rs_fwd(TW) + (49 * CB + CB) * (int)sizeof(float);
    cudaFuncSetAttribute(dwconv7_strip2_kernel<TW, TH2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwconv7_strip2_kernel<TW, TH2>, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    long long gx = ((long long)vkocr_sm_count() * per_sm) / cblocks;     // persistent: the resident block count, rounded DOWN
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dwconv7_strip2_kernel<TW, TH2><<<dim3((unsigned)gx, (unsigned)cblocks), threads, smem, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ld_x, reinterpret_cast<__nv_bfloat16*>(y), ld_y, B, H, W, C, wt, bias,
        reinterpret_cast<const __nv_bfloat16*>(add), ld_add, tiles_x, tiles_y, (int)ntiles);
    return 0;
}

template <int TW>
int launch_dw_tile(const void* x, long long ld_x, void* y, long long ld_y, int B, int H, int W, int C, const float* wt,
                   const float* bias, const void* add, long long ld_add, cudaStream_t s) {
    const int tiles_x = vk_cdiv(W, TW), tiles_y = vk_cdiv(H, TH);
    const long long ntiles = (long long)B * tiles_x * tiles_y;
    const int cblocks = C / CB;
    const int smem = 2 * HALO_H * rs_fwd(TW) + (49 * CB + CB) * (int)sizeof(float);
    // persistent: as many blocks as are resident at once (2 per SM by registers; shared memory allows it for every TW)
    cudaFuncSetAttribute(dwconv7_tile_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    static int per_sm = 0;      // resident blocks per SM (2 by registers for TW = 40 / 32, 4 for TW = 20)
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwconv7_tile_kernel<TW>, 8 * TW, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    long long gx = ((long long)vkocr_sm_count() * per_sm) / cblocks;     // rounded DOWN: one block too many is a whole second wave
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dwconv7_tile_kernel<TW><<<dim3((unsigned)gx, (unsigned)cblocks), 8 * TW, smem, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ld_x, reinterpret_cast<__nv_bfloat16*>(y), ld_y, B, H, W, C, wt, bias,
        reinterpret_cast<const __nv_bfloat16*>(add), ld_add, tiles_x, tiles_y, (int)ntiles);
    return 0;
}

template <int TW>
int launch_dw_wgrad_tile(const void* dy, long long ld_dy, const void* x, long long ld_x, int B, int H, int W, int C, float* dw,
                         cudaStream_t s) {
    const int tiles_x = vk_cdiv(W, TW), tiles_y = vk_cdiv(H, TH);
    const long long ntiles = (long long)B * tiles_x * tiles_y;
    const int cblocks = C / CB;
    const int smem = 2 * (HALO_H * rs_wg(TW) + TH * TW * PIX_B) + 49 * CB * (int)sizeof(float);
    cudaFuncSetAttribute(dwconv7_wgrad_tile_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // persistent over tiles: exactly the number of blocks that are resident at once, rounded DOWN -- a partial second
    // wave costs a whole block's run time
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwconv7_wgrad_tile_kernel<TW>, WG_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    long long gy = ((long long)vkocr_sm_count() * per_sm) / cblocks;
    if (gy > ntiles) gy = ntiles;
    if (gy < 1) gy = 1;
    dwconv7_wgrad_tile_kernel<TW><<<dim3((unsigned)cblocks, (unsigned)gy), WG_THREADS, smem, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), ld_dy, reinterpret_cast<const __nv_bfloat16*>(x), ld_x, B, H, W, C, tiles_x,
        tiles_y, dw);
    return 0;
}

}  // namespace

extern "C" {

// y[b,y,x,c] = bias[c] + sum_{ky,kx} wt[ky*7+kx][c] * x[b,y+ky-3,x+kx-3,c]  (+ add[b,y,x,c])
// wt is the [49][C] fp32 tap table written by vkocr_pack_dwconv_weight (mirrored for the data gradient).
int vkocr_dwconv7_fwd(int dtype, const void* x, long long ld_x, void* y, long long ld_y, int B, int H, int W, int C,
                      const float* wt, const float* bias, const void* add, long long ld_add, void* stream) {
    VK_REQUIRE(x && y && wt, VKOCR_BAD_ARGUMENT, "dwconv7: null argument");
    const int V = dtype == VKOCR_F32 ? 4 : 8;
    VK_REQUIRE(C % V == 0 && ld_x % V == 0 && ld_y % V == 0 && (!add || ld_add % V == 0), VKOCR_BAD_ALIGN,
               "dwconv7: C %d / strides must be multiples of %d", C, V);
    const long long total = (long long)B * H * ((W + STRIP - 1) / STRIP) * (C / V);
    if (total == 0) return VKOCR_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == VKOCR_BF16 && C % CB == 0 && (long long)B * H * W < (1LL << 30)) {
        const int tw = pick_tw_fwd(W);
        static const bool patch_kernel = getenv("VKOCR_DW_PATCH") != nullptr;      // A/B switch for tools/kbench.py
        int strip2 = 1;                                          // 8 x 2 strips; VKOCR_DW_STRIP2=0 is the A/B switch (tools/kbench.py)
        if (const char* e = getenv("VKOCR_DW_STRIP2")) strip2 = atoi(e);
        if (strip2 && tw == 16 && H % 16 == 0) launch_dw_strip2<16, 16>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (strip2 == 2 && tw == 40 && H % 8 == 0) launch_dw_strip2<40, 8>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);   // measured slower: 10 warps per SM
        else
        if (tw == 40 && !patch_kernel) launch_dw_strip<40>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (tw == 32 && !patch_kernel) launch_dw_strip<32>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (tw == 16) launch_dw_strip<16>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (tw == 8) launch_dw_strip<8>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (tw == 24) launch_dw_strip<24>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (tw == 40) launch_dw_tile<40>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else if (tw == 20) launch_dw_tile<20>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        else launch_dw_tile<32>(x, ld_x, y, ld_y, B, H, W, C, wt, bias, add, ld_add, s);
        VK_CHECK_LAUNCH("dwconv7_tile_kernel");
        return VKOCR_OK;
    }
    const unsigned blocks = (unsigned)((total + 255) / 256);
    VK_DISPATCH_DTYPE(dtype, T, (dwconv7_generic_kernel<T><<<blocks, 256, 0, s>>>(
                                    reinterpret_cast<const T*>(x), ld_x, reinterpret_cast<T*>(y), ld_y, B, H, W, C, wt, bias,
                                    reinterpret_cast<const T*>(add), ld_add)));
    VK_CHECK_LAUNCH("dwconv7_kernel");
    return VKOCR_OK;
}

// dw (fp32, parameter layout (C,1,7,7)) += correlation of dy with x.
int vkocr_dwconv7_wgrad(int dtype, const void* dy, long long ld_dy, const void* x, long long ld_x, int B, int H, int W, int C,
                        float* dw, void* stream) {
    VK_REQUIRE(dy && x && dw, VKOCR_BAD_ARGUMENT, "dwconv7_wgrad: null argument");
    const int V = 4;   // 4 channels per thread in either storage type (VkVec4)
    VK_REQUIRE(C % V == 0 && ld_x % V == 0 && ld_dy % V == 0, VKOCR_BAD_ALIGN, "dwconv7_wgrad: C %d / strides must be multiples of %d", C, V);
    if ((long long)B * H * W == 0) return VKOCR_OK;
    if (dtype == VKOCR_BF16 && C % CB == 0 && (long long)B * H * W < (1LL << 30)) {
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        const int tw = pick_tw(W);
        if (tw == 40) launch_dw_wgrad_tile<40>(dy, ld_dy, x, ld_x, B, H, W, C, dw, st);
        else if (tw == 20) launch_dw_wgrad_tile<20>(dy, ld_dy, x, ld_x, B, H, W, C, dw, st);
        else launch_dw_wgrad_tile<32>(dy, ld_dy, x, ld_x, B, H, W, C, dw, st);
        VK_CHECK_LAUNCH("dwconv7_wgrad_tile_kernel");
        return VKOCR_OK;
    }
    constexpr int SEG = 8;
    const int CV = C / V;
    // channel vectors per block: a divisor-friendly width so that (almost) all 256 threads carry a segment
    int cvb = CV < 32 ? CV : 32;
    const int zblocks = (CV + cvb - 1) / cvb;
    const int segs = (W + SEG - 1) / SEG;
    int nsg = 256 / cvb;
    if (nsg > segs) nsg = segs;
    const int threads = nsg * cvb;
    long long rows = (long long)B * H;
    long long gx = ((long long)vkocr_sm_count() * 6 + 7 * zblocks - 1) / (7 * zblocks);
    if (gx > rows) gx = rows;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, 7, (unsigned)zblocks);
    const size_t smem = (size_t)7 * cvb * V * sizeof(float);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    VK_DISPATCH_DTYPE(dtype, T, (dwconv7_wgrad_generic_kernel<T, SEG><<<grid, threads, smem, s>>>(
                                    reinterpret_cast<const T*>(dy), ld_dy, reinterpret_cast<const T*>(x), ld_x, B, H, W, C, cvb,
                                    dw)));
    VK_CHECK_LAUNCH("dwconv7_wgrad_kernel");
    return VKOCR_OK;
}

}  // extern "C"
