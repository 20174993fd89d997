// tcgen05 / TMEM / TMA GEMM for sm_100a: the tensor-core kernel behind every dense contraction of the
// adaptive-scaling network (Linear / 1x1, patchify convs, 3x3 neck + head convs; forward, data-gradient and
// weight-gradient).  bf16 operands, fp32 accumulation in tensor memory.
//
// One persistent, warp-specialised kernel:
//   warp 0 (1 lane)  TMA producer: 4-D tiled tensor maps over NHWC activations (zero-filled out-of-bounds boxes
//                    give the conv padding for free), 2-D map over the packed weights; SWIZZLE_128B smem tiles.
//   warp 1 (1 lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN (16..256), K=16 per instruction,
//                    accumulators double-buffered in TMEM (2 x 256 columns).
//   warp 2           TMEM allocator.
//   warps 4..15      epilogue: tcgen05.ld -> registers -> bias / GELU / layer-scale / residual -> 32 x 64 tile staged in
//                    shared memory -> coalesced 128-byte-row global stores.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> epilogue).
//
// Modes (see gemm_common.cuh):
//   NT  D[m,n] = sum_{tap,c} X[pix(m)+off(tap), c] * Wp[n, tap, c]       A,B K-major
//   TN  G[tap,i,j] = sum_pix P[pix,i] * Q[pix+off(tap), j]               A,B MN-major, split over pixels, fp32 red.add
#include "gemm_common.cuh"
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements in one 128-byte swizzle row
constexpr int A_BYTES = BM * BK * 2;   // 16 KiB
constexpr int MAX_SMEM = 176 * 1024;  // operand stages
constexpr int EPI_TILE_BYTES = 2048;  // per epilogue warp: 32 rows x 64 B staging tile (row-per-lane in, 8 row segments per store out)
constexpr int NUM_THREADS = 512;         // warps 0-3: TMA / MMA / TMEM alloc / second TMA producer; warps 4-15: epilogue
constexpr int EPI_WARPS = 12;            // three warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int HEAD_PAR_BYTES = 7 * 256 * 4;                 // fused head tail: gamma, beta, conv bias, w2[4] of the CTA's head
constexpr int HEAD_XCH_BYTES = EPI_WARPS * 32 * 6 * 4;      // per-warp partial row statistics (2) and projections (4)
constexpr int PREFETCH_BYTES = EPI_WARPS * EPI_TILE_BYTES;   // second set of staging tiles, carved out of the operand-stage budget
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;

// Division of a 31-bit unsigned by a launch-time constant without a divide: q = (mulhi(n, m) + n) >> s.
// (The unit -> tile decode runs once per tile in every warp; with 64-bit `/` and `%` it cost as many instructions as the
// epilogue of a K = 96 tile itself.)
struct FastDiv {
    uint32_t d, m, s;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.s = s;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.m) + n) >> f.s; }
__device__ __forceinline__ void fd_divmod(uint32_t n, const FastDiv& f, uint32_t* q, uint32_t* r) {
    *q = fd_div(n, f);
    *r = n - *q * f.d;
}

// Compile-time description of a staged epilogue; RT = decide everything at run time.  The kernel is instantiated once
// per kind (template parameter EPI): with every feature decided by run-time tests inside the 32-column chunk loop the
// epilogue warps' code was ~1900 SASS instructions of mostly-untaken branches, and the plain bf16 store of a K = 384 GEMM
// stalled on instruction fetch (ncu: stall_no_inst / branch_resolving on every flag test) -- 5.3 ms of epilogue against
// 3.6 ms of MMA + TMA for [819200, 384] x [384, 7488].
template <int ACT_, bool BIAS_, bool CS_, bool RES_, bool PRE_, bool RT_>
struct EpiKind {
    static constexpr int ACT = ACT_;
    static constexpr bool BIAS = BIAS_, CS = CS_, RES = RES_, PRE = PRE_, RT = RT_;
};
template <int EPI> struct KindOf { using type = EpiKind<0, false, false, false, false, true>; };          // 0: run-time flags
template <> struct KindOf<1> { using type = EpiKind<0, false, false, false, false, false>; };            // plain store
template <> struct KindOf<2> { using type = EpiKind<0, true, false, false, false, false>; };             // + bias
template <> struct KindOf<3> { using type = EpiKind<3, true, false, false, true, false>; };              // bias, GELU, gelu' side channel
template <> struct KindOf<4> { using type = EpiKind<1, true, false, false, false, false>; };             // bias, GELU
template <> struct KindOf<5> { using type = EpiKind<0, true, true, true, false, false>; };               // bias, layer scale, residual
template <> struct KindOf<6> { using type = EpiKind<4, false, false, false, false, false>; };            // * aux (gelu' side channel)
constexpr int NUM_EPI = 7;

struct TcParams {
    int mode;                // 0 NT, 1 TN
    int batch, H, W;
    int BW, BH;              // pixel box (BW*BH == 128 for NT, 64 for TN)
    int bw_shift;            // log2(BW)
    int tiles_x, tiles_y;
    int ks;
    int kb_per_tap;          // NT: 64-wide K blocks per tap
    int N, BN, n_tiles;      // NT: output channels; TN: J (channels of Q)
    int I, i_tiles;          // TN: channels of P
    int splits;              // TN: split of the pixel-tile range
    int num_stages;
    int stage_bytes;
    int skip_tma;            // debug: producers arrive without loading (timing experiments)
    int no_acc_prefetch;     // A/B switch: the epilogue loads each chunk's accumulators only when it gets to them
    int tn_vec;              // TN: J is contiguous and 16-byte aligned in the output -> staged 16-byte reductions
    int tma_store;           // NT staged epilogue: the staging tiles leave as bulk tensor stores (mapO: out, mapP: out_pre)
    int staged_store;        // NT: aligned bf16 outputs leave through the shared-memory staging tiles
    int cta2;                // NT: CTA pairs (cluster of 2, tcgen05 cta_group::2): M = 256 per pair, each CTA stages its 128 pixel rows and HALF of the weight tile
    int prefetch_extra;      // NT staged path: residual / aux tiles are prefetched one chunk ahead (cp.async) into a second set of staging tiles
    int M;                   // NT plain GEMM: number of rows
    long long units;
    int head_mode;           // NT: fused head tail epilogue (ht valid)
    int epi_kind;            // which instance of the kernel (template parameter EPI) matches ep; 0 = the general one
    FastDiv fd_n, fd_x, fd_y, fd_i, fd_taps, fd_rpg;   // n_tiles, tiles_x, tiles_y, i_tiles, ks*ks, rows_per_group
    VkocrEpilogue ep;
    VkocrHeadTail ht;
};

// ------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a broken pipeline traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("vkocr gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, bar, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// Bulk tensor STORE of a staged tile (shared -> global, bulk async-group completion).  Rows / columns outside the tensor are
// clipped by the copy engine.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// ---- CTA-pair (cta_group::2) forms.  The TMA loads of BOTH CTAs signal the leader's full barrier, the leader's MMA
// commit arrives on the barriers of both CTAs (multicast), the follower's epilogue releases the accumulator remotely.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_cta(uint32_t addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
    return ra;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.cta_group::2 [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.cta_group::2 [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {   // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// four consecutive fp32 accumulations as ONE 16-byte reduction (REDG.E.ADD.F32x4)
__device__ __forceinline__ void red_add_v4(float* addr, uint4 v) {
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ void unpack8(uint4 raw, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __bfloat1622float2(h[i]);
        f[2 * i] = v.x;
        f[2 * i + 1] = v.y;
    }
}
// fp16 flavour for the gelu' side channel (see vk_store_dgelu)
__device__ __forceinline__ void unpack8_f16(uint4 raw, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(h[i]);
        f[2 * i] = v.x;
        f[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ uint4 pack8_f16(const float* f) {
    uint4 raw;
    __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return raw;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return raw;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// The same load split in two so that the epilogue can have the NEXT chunk's accumulators in flight while it stages and stores
// the current one.  The wait names all 32 result registers as read-write operands: every consumer depends on the wait, not
// on the load (see the note in tc_ld32 below), and nothing else touches the registers in between.
#define VK_F32RW(r) "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]), "+f"(r[9]), \
    "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]),  \
    "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]),  \
    "+f"(r[30]), "+f"(r[31])
// (the results land in the float variables the arithmetic works on: no register copies between the load and its consumers)
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),
          "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),
          "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld32_wait(float* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : VK_F32RW(r) : : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        // The wait lives in the SAME asm statement as the load: as a separate statement nothing ties it to the 32 result
        // registers, and the compiler may schedule their consumers (in particular non-volatile inline-asm arithmetic)
        // between the asynchronous load and its wait -- observed as a 1.5e-2 -> 2.0e-2 jump of the bf16 gradient error.
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.  Offsets are in 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // layout type: SWIZZLE_128B
    return d;
}

// ------------------------------------------------------------------------------------------------- kernel
// CTA2: compile-time, because every tcgen05 instruction of a kernel must name the same cta_group (and a kernel that uses
// cta_group::2 can only be launched as clusters of two).
template <bool CTA2, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
vkocr_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapP,
                     const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.num_stages;
    const uint32_t epi_base = smem_base + (uint32_t)S * (uint32_t)p.stage_bytes;   // 1024-byte aligned (stage sizes are)
    const uint32_t head_off = (uint32_t)S * (uint32_t)p.stage_bytes + (uint32_t)(EPI_WARPS * EPI_TILE_BYTES);   // from smem_base
    const uint32_t bar_base = smem_base + head_off + (uint32_t)(HEAD_PAR_BYTES + HEAD_XCH_BYTES);
    // barrier layout: full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base address word
    auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * S + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * S + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * S + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), CTA2 ? 2 * EPI_WARPS : EPI_WARPS);      // one arrival per epilogue warp (of both CTAs of a pair)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    constexpr bool cta2 = CTA2;
    uint32_t cta_rank = 0u;
    if constexpr (CTA2) cta_rank = cluster_ctarank();
    // work items: units (one CTA each), or unit PAIRS (two pixel tiles x one N tile) per cluster in pair mode
    const long long w0 = cta2 ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
    const long long wstride = cta2 ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
    const long long wcount = cta2 ? (p.units >> 1) : p.units;
    auto unit_of = [&](long long w) -> long long {
        if (!cta2) return w;
        uint32_t mp, nt;
        fd_divmod((uint32_t)w, p.fd_n, &mp, &nt);
        return (long long)(2u * mp + cta_rank) * p.n_tiles + nt;
    };
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    const int BN = p.BN;
    const int pix_tiles = p.batch * p.tiles_y * p.tiles_x;
    const int half = p.ks >> 1;

    // unit -> coordinates
    struct Unit {
        int n0, nt, b, y0, x0;      // NT
        int tap, i0, kt_begin, kt_end;  // TN
        int num_kb;
    };
    auto decode = [&](long long ul) {
        Unit t;
        const uint32_t u = (uint32_t)ul;          // units < 2^31 (checked on the host)
        if (p.mode == 0) {
            uint32_t nt, mt, tx, ty, b;
            fd_divmod(u, p.fd_n, &mt, &nt);
            fd_divmod(mt, p.fd_x, &mt, &tx);
            fd_divmod(mt, p.fd_y, &b, &ty);
            t.b = (int)b;
            t.nt = (int)nt;
            t.n0 = (int)nt * BN;
            t.x0 = (int)tx * p.BW;
            t.y0 = (int)ty * p.BH;
            t.num_kb = p.ks * p.ks * p.kb_per_tap;
            t.tap = 0; t.i0 = 0; t.kt_begin = 0; t.kt_end = 0;
        } else {
            uint32_t jt, r, it, tap, sp;
            fd_divmod(u, p.fd_n, &r, &jt);
            fd_divmod(r, p.fd_i, &r, &it);
            fd_divmod(r, p.fd_taps, &sp, &tap);
            t.tap = (int)tap;
            t.n0 = (int)jt * BN;
            t.i0 = (int)it * BM;
            t.kt_begin = (int)((long long)pix_tiles * sp / p.splits);
            t.kt_end = (int)((long long)pix_tiles * (sp + 1) / p.splits);
            t.num_kb = t.kt_end - t.kt_begin;
            t.b = 0; t.y0 = 0; t.x0 = 0; t.nt = 0;
        }
        return t;
    };

    if (warp == 3 && lane == 0 && p.mode == 1) {
        // ================================================================ second TMA producer (TN): the Q operand boxes.
        // A cp.async.bulk.tensor costs its issuing thread on the order of 100 cycles; with five boxes per k-block a single
        // producer thread cannot keep up with the MMAs, so the P and Q operands are issued from two threads.  Both wait on
        // the stage's empty barrier; warp 0 performs the full barrier's arrival (expect_tx of the whole stage).
        int s = 0;
        uint32_t ph = 0;
        const int nb = (BN + 63) / 64;
        for (long long w = w0; w < wcount; w += wstride) {
            const long long u = unit_of(w);
            const Unit t = decode(u);
            const int dy = t.tap / p.ks - half, dx = t.tap % p.ks - half;
            int kt = t.kt_begin;
            int tx = kt % p.tiles_x;
            kt /= p.tiles_x;
            int ty = kt % p.tiles_y;
            int b = kt / p.tiles_y;
            for (int kb = 0; kb < t.num_kb; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                const uint32_t sb = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes + A_BYTES;
                const int x0 = tx * p.BW + dx, y0 = ty * p.BH + dy;
                if (!(p.skip_tma & 1))
                    for (int g = 0; g < nb; ++g) tma_load_4d(sb + (uint32_t)g * 8192u, &mapB, full_bar(s), t.n0 + g * 64, x0, y0, b);
                if (++tx == p.tiles_x) {
                    tx = 0;
                    if (++ty == p.tiles_y) { ty = 0; ++b; }
                }
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 0 && lane == 0) {
        // ================================================================ TMA producer
        int s = 0;
        uint32_t ph = 0;
        const uint32_t tx_bytes = (p.mode == 0) ? (uint32_t)(A_BYTES + BN * 128) : (uint32_t)(2 * 8192 + ((BN + 63) / 64) * 8192);
        // All per-k-block coordinates are tracked incrementally: integer divisions here sit on the critical path of a
        // single thread (a handful of dependent divides per k-block costs more than the k-block's MMAs).
        for (long long w = w0; w < wcount; w += wstride) {
            const long long u = unit_of(w);
            const Unit t = decode(u);
            if (p.mode == 0) {
                int c0 = 0, dy = -half, dx = -half;
                for (int kb = 0; kb < t.num_kb; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    if constexpr (CTA2) {
                        // both CTAs' copies complete on the LEADER's full barrier; each brings its pixel rows and its half of
                        // the weight tile
                        const uint32_t fb = cta_rank ? mapa_cta(full_bar(s), 0u) : full_bar(s);
                        if (cta_rank == 0) mbar_expect_tx(full_bar(s), 2u * (uint32_t)(A_BYTES + (BN >> 1) * 128));
                        tma_load_4d_2cta(sa, &mapA, fb, c0, t.x0 + dx, t.y0 + dy, t.b);
                        tma_load_2d_2cta(sa + A_BYTES, &mapB, fb, kb * BK, t.n0 + (int)cta_rank * (BN >> 1));
                    } else {
                        if (p.skip_tma & 1) {
                            mbar_arrive(full_bar(s));
                        } else {
                            mbar_expect_tx(full_bar(s), tx_bytes);
                            tma_load_4d(sa, &mapA, full_bar(s), c0, t.x0 + dx, t.y0 + dy, t.b);
                            tma_load_2d(sa + A_BYTES, &mapB, full_bar(s), kb * BK, t.n0);
                        }
                    }
                    c0 += BK;
                    if (c0 >= p.kb_per_tap * BK) {
                        c0 = 0;
                        if (++dx > half) { dx = -half; ++dy; }
                    }
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            } else {
                int kt = t.kt_begin;
                int tx = kt % p.tiles_x;
                kt /= p.tiles_x;
                int ty = kt % p.tiles_y;
                int b = kt / p.tiles_y;
                for (int kb = 0; kb < t.num_kb; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    const int x0 = tx * p.BW, y0 = ty * p.BH;
                    if (p.skip_tma & 1) {
                        mbar_arrive(full_bar(s));
                    } else {
                        mbar_expect_tx(full_bar(s), tx_bytes);
                        tma_load_4d(sa, &mapA, full_bar(s), t.i0, x0, y0, b);
                        tma_load_4d(sa + 8192, &mapA, full_bar(s), t.i0 + 64, x0, y0, b);
                    }
                    if (++tx == p.tiles_x) {
                        tx = 0;
                        if (++ty == p.tiles_y) { ty = 0; ++b; }
                    }
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && lane == 0 && cta_rank == 0) {
        // ================================================================ MMA issuer (pair mode: the leader CTA only)
        int s = 0;
        uint32_t ph = 0;
        int as = 0;
        uint32_t aph = 0;
        const uint32_t major = (p.mode == 0) ? 0u : 1u;
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (major << 15) | (major << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((cta2 ? 2 * BM : BM) >> 4) << 24);
        const int dbg_major = (p.skip_tma >> 1) & 3;   // debug (timing only, garbage data): bit0 -> A K-major, bit1 -> B K-major
        if (dbg_major & 1) idesc &= ~(1u << 15);
        if (dbg_major & 2) idesc &= ~(1u << 16);
        for (long long w = w0; w < wcount; w += wstride) {
            const long long u = unit_of(w);
            const Unit t = decode(u);
            if (t.num_kb == 0) continue;
            mbar_wait(tempty_bar(as), aph ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * ACC_STRIDE);
            for (int kb = 0; kb < t.num_kb; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                const uint32_t sb = sa + A_BYTES;
                if (p.mode == 0) {
                    // K-major SWIZZLE_128B: 8-row atoms of 1024 B; advance 32 B per K=16 slice inside the atom.
                    const uint64_t ad = make_smem_desc(sa, 1, 64);
                    const uint64_t bd = make_smem_desc(sb, 1, 64);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if constexpr (CTA2) tc_mma2(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                        else if (!(p.skip_tma & 32)) tc_mma(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    }
                } else {
                    // MN-major SWIZZLE_128B: 64-channel groups 8192 B apart (LBO), 8-pixel atoms 1024 B apart (SBO);
                    // one K=16 slice = 2 atoms = 2048 B.
                    const uint64_t ad = (dbg_major & 1) ? make_smem_desc(sa, 1, 64) : make_smem_desc(sa, 512, 64);
                    const uint64_t bd = (dbg_major & 2) ? make_smem_desc(sb, 1, 64) : make_smem_desc(sb, 512, 64);
                    const uint64_t ka = (dbg_major & 1) ? 2 : 128, kbs = (dbg_major & 2) ? 2 : 128;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        if constexpr (!CTA2) tc_mma(tmem_d, ad + ka * (uint64_t)k, bd + kbs * (uint64_t)k, idesc, (kb | k) ? 1u : 0u);
                }
                if constexpr (CTA2) tc_commit2(empty_bar(s));
                else tc_commit(empty_bar(s));
                if (++s == S) { s = 0; ph ^= 1u; }
            }
            if constexpr (CTA2) tc_commit2(tfull_bar(as));
            else tc_commit(tfull_bar(as));
            if (++as == 2) { as = 0; aph ^= 1u; }
        }
    } else if (warp >= 4) {
        // ================================================================ epilogue (TMEM -> registers -> global)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access (warp id % 4)
        const int r = q * 32 + lane;            // accumulator row
        const int cpart = (warp - 4) >> 2;      // which interleaved share of the column chunks
        int as = 0;
        uint32_t aph = 0;
        const VkocrEpilogue& ep = p.ep;
        const int chunks = (BN + 31) / 32;
        int cur_head = -1;
        // bias / layer-scale vectors of the staged NT path live in shared memory (one global read per CTA instead of one
        // L2 round trip per 32-column chunk and pass): [0, 896) bias, [896, 1792) column scale, when N fits
        uint8_t* smem_gen0 = smem_raw + (smem_base - smem_u32(smem_raw));
        float* s_vec = reinterpret_cast<float*>(smem_gen0 + head_off);
        const bool vec_in_smem = p.mode == 0 && p.staged_store && !p.head_mode && p.N <= 896;
        if (vec_in_smem) {
            for (int i = (int)threadIdx.x - 128; i < p.N; i += 32 * EPI_WARPS) {
                s_vec[i] = ep.bias ? __ldg(ep.bias + i) : 0.f;
                s_vec[896 + i] = ep.col_scale ? __ldg(ep.col_scale + i) : 1.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        }
        bool pf_issued = false;                 // the first chunk of the coming tile is already being prefetched
        bool tma_pending = false;               // a bulk store of this warp's staging tile may still be reading it
        const uint32_t pre_base = bar_base + 256u;
        for (long long w = w0; w < wcount; w += wstride) {
            const long long u = unit_of(w);
            const Unit t = decode(u);
            bool row_ok;
            long long row;
            if (p.mode == 0) {
                const int y = t.y0 + r / p.BW, x = t.x0 + r % p.BW;
                row_ok = (y < p.H) && (x < p.W);
                row = ((long long)t.b * p.H + y) * p.W + x;
            } else {
                row_ok = (t.i0 + r) < p.I;
                row = (long long)t.tap * p.I + t.i0 + r;
            }
            if (t.num_kb == 0) continue;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_STRIDE);
            const bool vec_ok = (p.mode == 0) && (!ep.out_f32) && ((ep.ldo & 7) == 0) && (!ep.out_pre || (ep.ld_pre & 7) == 0) &&
                                (!ep.residual || (ep.ld_res & 7) == 0) && (ep.act != 2 || (ep.ld_aux & 7) == 0) && ep.act <= 2;
            auto staged_tile = [&](auto kind) {
                using K = decltype(kind);
                // ---- fast path, one 32-column chunk at a time: each lane finishes its row in registers (bias, GELU / GELU',
                // layer scale, drop-path mask, residual), the warp's 32 x 32 tile is staged in shared memory (64-byte rows,
                // XOR-swizzled: conflict-free both ways) and leaves as coalesced 64-byte row segments, 8 rows per store
                // instruction.  Residual / GELU' operands come in the same way, transposed through the tile.
                const uint32_t wb = epi_base + (uint32_t)(warp - 4) * (uint32_t)EPI_TILE_BYTES;
                const uint32_t my_row = wb + (uint32_t)lane * 64u;
                const uint32_t sw = (uint32_t)((lane >> 1) & 3);
                const int r0 = q * 32;
                const int act = K::RT ? ep.act : K::ACT;
                const bool has_bias = K::RT ? (ep.bias != nullptr) : K::BIAS;
                const bool has_cs = K::RT ? (ep.col_scale != nullptr) : K::CS;
                const bool has_res = K::RT ? (ep.residual != nullptr) : K::RES;
                const bool has_pre = K::RT ? (ep.out_pre != nullptr) : K::PRE;
                const bool uses_extra = has_res || act == 2 || act == 4;
                const __nv_bfloat16* esrc = has_res ? reinterpret_cast<const __nv_bfloat16*>(ep.residual)
                                                    : ((act == 2 || act == 4) ? reinterpret_cast<const __nv_bfloat16*>(ep.aux) : nullptr);
                const long long eld = has_res ? ep.ld_res : ep.ld_aux;
                const float rs = (ep.row_scale && row_ok) ? __ldg(ep.row_scale + fd_div((uint32_t)row, p.fd_rpg)) : 1.f;
                // global pixel index of the 4 staged rows this lane moves (rows 8*i + lane/4 of the warp's quarter)
                int pix[4];
                unsigned okmask = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = r0 + 8 * i + (lane >> 2);
                    const int yy = t.y0 + (rl >> p.bw_shift), xx = t.x0 + (rl & (p.BW - 1));
                    pix[i] = (t.b * p.H + yy) * p.W + xx;
                    if (yy < p.H && xx < p.W) okmask |= 1u << i;
                }
                const int seg = lane & 3;
                const int col_end = (t.n0 + BN < p.N) ? t.n0 + BN : p.N;   // 8-column groups are all-or-nothing
                // Prefetch of the residual / aux tile of chunk cc of tile tt into the warp's second staging tile (same
                // swizzled layout).  A synchronous load here exposes a full DRAM round trip per chunk, which bounded the
                // K <= 384 GEMMs; the copy for the next chunk (or the next tile's first chunk) now flies during this one.
                const uint32_t pwb = pre_base + (uint32_t)(warp - 4) * (uint32_t)EPI_TILE_BYTES;
                auto issue_prefetch = [&](const Unit& tt, int cc) {
                    const int ce = (tt.n0 + BN < p.N) ? tt.n0 + BN : p.N;
                    const int pc = tt.n0 + cc * 32 + seg * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = 8 * i + (lane >> 2);
                        const int rl = r0 + rr;
                        const int yy = tt.y0 + (rl >> p.bw_shift), xx = tt.x0 + (rl & (p.BW - 1));
                        const bool ok = yy < p.H && xx < p.W && pc < ce && cc < chunks;
                        const long long px = (long long)((tt.b * p.H + yy) * p.W + xx);
                        const void* src = ok ? static_cast<const void*>(esrc + px * eld + pc) : static_cast<const void*>(esrc);
                        const int nbytes = ok ? 16 : 0;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(pwb + (uint32_t)rr * 64u + (((uint32_t)seg ^ (uint32_t)((rr >> 1) & 3)) << 4)),
                                     "l"(src), "r"(nbytes) : "memory");
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                const bool prefetch = p.prefetch_extra && esrc != nullptr;
                if (prefetch && !pf_issued) issue_prefetch(t, cpart);
                pf_issued = false;
                // the accumulators of the warp's NEXT chunk of this tile are requested from tensor memory as soon as the
                // current chunk's values are packed, and arrive while the current chunk is staged and stored
                float v[32];
                bool acc_issued = false;
                // pixel coordinates of the warp's first row inside the image (bulk tensor stores)
                const int st_x = t.x0 + (r0 & (p.BW - 1)), st_y = t.y0 + (r0 >> p.bw_shift);
                for (int c = cpart; c < chunks; c += EPI_WARPS / 4) {
                    const int cb = c * 32;
                    const int nb = t.n0 + cb;
                    if (nb >= col_end || (p.skip_tma & 16)) break;   // debug bit 4: epilogue = barrier traffic only
                    if (!acc_issued) tc_ld32_issue(taddr + (uint32_t)cb, v);
                    if (tma_pending && uses_extra && !prefetch) {      // the extra operand is transposed through the same tile
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
                        tma_pending = false;
                    }
                    const int c_next = c + EPI_WARPS / 4;
                    const bool more = !p.no_acc_prefetch && c_next < chunks && t.n0 + c_next * 32 < col_end;
                    const int col = nb + seg * 8;
                    uint4 extra[4];
                    if (uses_extra && prefetch) {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) extra[j] = lds128(pwb + (uint32_t)lane * 64u + (((uint32_t)j ^ sw) << 4));
                        __syncwarp();
                        const int cn = c + EPI_WARPS / 4;
                        if (cn < chunks && t.n0 + cn * 32 < col_end) {
                            issue_prefetch(t, cn);
                        } else if (w + wstride < wcount) {
                            issue_prefetch(decode(unit_of(w + wstride)), cpart);
                            pf_issued = true;
                        }
                    } else if (uses_extra) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = 8 * i + (lane >> 2);
                            uint4 val = make_uint4(0u, 0u, 0u, 0u);
                            if ((okmask & (1u << i)) && col < col_end)
                                val = __ldg(reinterpret_cast<const uint4*>(esrc + (long long)pix[i] * eld + col));
                            sts128(wb + (uint32_t)rr * 64u + (((uint32_t)seg ^ (uint32_t)((rr >> 1) & 3)) << 4), val);
                        }
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) extra[j] = lds128(my_row + (((uint32_t)j ^ sw) << 4));
                        __syncwarp();
                    }
                    // stage the lane's 32 finished values and write the warp's tile out as coalesced row segments
                    auto store_packed = [&](const uint4 (&pk)[4], void* base, long long ld) {
                        if (p.tma_store) {
                            // the tile's layout IS the copy engine's SWIZZLE_64B box (32 columns x the warp's 32 pixels): one
                            // lane hands it over and the warp moves on; the tile is reused once the engine has read it
                            if (tma_pending) {
                                if (lane == 0) bulk_wait_read0();
                                __syncwarp();
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) sts128(my_row + (((uint32_t)j ^ sw) << 4), pk[j]);
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0 && !(p.skip_tma & 8)) {
                                tma_store_4d(base == ep.out ? &mapO : &mapP, wb, nb, st_x, st_y, t.b);
                                bulk_commit();
                            }
                            tma_pending = true;
                            return;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) sts128(my_row + (((uint32_t)j ^ sw) << 4), pk[j]);
                        __syncwarp();
                        if (col < col_end && !(p.skip_tma & 8)) {   // debug bit 3: no global stores
                            __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(base);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int rr = 8 * i + (lane >> 2);
                                if (okmask & (1u << i)) {
                                    const uint4 val = lds128(wb + (uint32_t)rr * 64u + (((uint32_t)seg ^ (uint32_t)((rr >> 1) & 3)) << 4));
                                    *reinterpret_cast<uint4*>(obase + (long long)pix[i] * ld + col) = val;
                                }
                            }
                        }
                        __syncwarp();      // the tile is free again
                    };
                    auto store_tile = [&](const float* vals, void* base, long long ld) {
                        uint4 pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) pk[j] = pack8(vals + 8 * j);
                        store_packed(pk, base, ld);
                    };
                    tc_ld32_wait(v);
                    const bool full = nb + 32 <= p.N;
                    // per-element arithmetic on fp32 PAIRS (FADD2 / FMUL2 / FFMA2): one issue slot for two columns
                    auto add4 = [&](int j, const float4 b) {
                        const float2 a = vk_add2(make_float2(v[j], v[j + 1]), make_float2(b.x, b.y));
                        const float2 c = vk_add2(make_float2(v[j + 2], v[j + 3]), make_float2(b.z, b.w));
                        v[j] = a.x; v[j + 1] = a.y; v[j + 2] = c.x; v[j + 3] = c.y;
                    };
                    auto mul4 = [&](int j, const float4 b) {
                        const float2 a = vk_mul2(make_float2(v[j], v[j + 1]), make_float2(b.x, b.y));
                        const float2 c = vk_mul2(make_float2(v[j + 2], v[j + 3]), make_float2(b.z, b.w));
                        v[j] = a.x; v[j + 1] = a.y; v[j + 2] = c.x; v[j + 3] = c.y;
                    };
                    if (has_bias && vec_in_smem && full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) add4(j, *reinterpret_cast<const float4*>(s_vec + nb + j));
                    } else if (has_bias) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) add4(j, __ldg(reinterpret_cast<const float4*>(ep.bias + nb + j)));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += (nb + j < p.N) ? __ldg(ep.bias + nb + j) : 0.f;
                        }
                    }
                    if (act == 3) {
                        // GELU and its derivative share all transcendental work: the derivative is what the backward
                        // needs (second output), the pre-activation itself is never read again
                        uint4 pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float dg[8];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {      // packed pairs: the epilogue is bound by instruction issue
                                float2 g2, d2;
                                vk_gelu_both2(make_float2(v[8 * j + e], v[8 * j + e + 1]), &g2, &d2);
                                v[8 * j + e] = g2.x; v[8 * j + e + 1] = g2.y;
                                dg[e] = d2.x; dg[e + 1] = d2.y;
                            }
                            if (ep.row_scale) {       // stochastic-depth mask: side channel = m_b * gelu', dropped samples' activation = 0
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    dg[e] *= rs;
                                    v[8 * j + e] = rs != 0.f ? v[8 * j + e] : 0.f;
                                }
                            }
                            pk[j] = pack8_f16(dg);
                        }
                        if (has_pre) store_packed(pk, ep.out_pre, ep.ld_pre);
                    } else {
                        if (has_pre) store_tile(v, ep.out_pre, ep.ld_pre);      // pre-activation copy (acc + bias)
                        if (act == 1) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float2 g2 = vk_gelu2(make_float2(v[j], v[j + 1]));
                                v[j] = g2.x; v[j + 1] = g2.y;
                            }
                        } else if (act == 2) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float f[8];
                                unpack8(extra[j], f);
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[8 * j + e] *= vk_gelu_grad(f[e]);
                            }
                        } else if (act == 4) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float f[8];
                                unpack8_f16(extra[j], f);
                                mul4(8 * j, make_float4(f[0], f[1], f[2], f[3]));
                                mul4(8 * j + 4, make_float4(f[4], f[5], f[6], f[7]));
                            }
                        }
                    }
                    if (has_cs && vec_in_smem && full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) mul4(j, *reinterpret_cast<const float4*>(s_vec + 896 + nb + j));
                    } else if (has_cs) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) mul4(j, __ldg(reinterpret_cast<const float4*>(ep.col_scale + nb + j)));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= (nb + j < p.N) ? __ldg(ep.col_scale + nb + j) : 0.f;
                        }
                    }
                    if (ep.row_scale && act != 3) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) mul4(j, make_float4(rs, rs, rs, rs));
                    }
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float f[8];
                            unpack8(extra[j], f);
                            add4(8 * j, make_float4(f[0], f[1], f[2], f[3]));
                            add4(8 * j + 4, make_float4(f[4], f[5], f[6], f[7]));
                        }
                    }
                    {
                        uint4 pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) pk[j] = pack8(v + 8 * j);
                        acc_issued = more;
                        if (more) tc_ld32_issue(taddr + (uint32_t)(c_next * 32), v);       // the packed copy is all that is still needed
                        store_packed(pk, ep.out, ep.ldo);
                    }
                }
            };
            if constexpr (EPI != 0) {
                staged_tile(typename KindOf<EPI>::type{});
            } else
            if (p.head_mode) {
                // ---- fused head tail (one head per N tile): conv bias -> [conv output stored for the backward] ->
                // LayerNorm over the head's `inner` columns -> exact GELU -> projection to O <= 4 maps (-> Softplus), all
                // from the fp32 accumulators.  The three warps of a TMEM lane quarter split the columns; partial row
                // statistics and partial projections are exchanged through shared memory with a 96-thread named barrier.
                const VkocrHeadTail& ht = p.ht;
                uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
                float* s_par = reinterpret_cast<float*>(smem_gen + head_off);                       // [7][256]
                float* s_xs = reinterpret_cast<float*>(smem_gen + head_off + HEAD_PAR_BYTES);       // [EPI_WARPS][32][2]
                float* s_xd = s_xs + EPI_WARPS * 32 * 2;                                            // [EPI_WARPS][32][4]
                const int head = t.nt;
                const int inner = ht.inner[head], O = ht.out_channels[head];
                if (head != cur_head) {
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // nobody still reads the old parameters
                    for (int c = (int)threadIdx.x - 128; c < 256; c += 32 * EPI_WARPS) {
                        const bool ok = c < inner;
                        s_par[c] = ok ? __ldg(ht.gamma[head] + c) : 0.f;
                        s_par[256 + c] = ok ? __ldg(ht.beta[head] + c) : 0.f;
                        s_par[512 + c] = (ep.bias && c < BN && t.n0 + c < p.N) ? __ldg(ep.bias + t.n0 + c) : 0.f;
#pragma unroll
                        for (int o = 0; o < 4; ++o) s_par[(3 + o) * 256 + c] = (ok && o < O) ? __ldg(ht.w2[head] + (long long)o * inner + c) : 0.f;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                    cur_head = head;
                }
                const int wi = warp - 4;
                const uint32_t wb = epi_base + (uint32_t)wi * (uint32_t)EPI_TILE_BYTES;
                const uint32_t my_row = wb + (uint32_t)lane * 64u;
                const uint32_t sw = (uint32_t)((lane >> 1) & 3);
                const int r0 = q * 32;
                const int seg = lane & 3;
                const bool store_conv = ep.out != nullptr && !(p.skip_tma & 128);   // debug bit 7: no conv output store
                int pix[4];
                unsigned okmask = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = r0 + 8 * i + (lane >> 2);
                    const int yy = t.y0 + (rl >> p.bw_shift), xx = t.x0 + (rl & (p.BW - 1));
                    pix[i] = (t.b * p.H + yy) * p.W + xx;
                    if (yy < p.H && xx < p.W) okmask |= 1u << i;
                }
                const int col_end = (t.n0 + BN < p.N) ? t.n0 + BN : p.N;
                // phase 1: statistics (+ conv output store)
                float rsum = 0.f, rsq = 0.f;
                for (int c = cpart; c < chunks; c += EPI_WARPS / 4) {
                    const int cb = c * 32;
                    float v[32];
                    {
                        uint32_t acc[32];
                        tc_ld32(taddr + (uint32_t)cb, acc);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(s_par + 512 + cb + j);
                            v[j] = __uint_as_float(acc[j]) + b4.x; v[j + 1] = __uint_as_float(acc[j + 1]) + b4.y;
                            v[j + 2] = __uint_as_float(acc[j + 2]) + b4.z; v[j + 3] = __uint_as_float(acc[j + 3]) + b4.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x_ = (cb + j < inner) ? v[j] : 0.f;      // pad columns of the slot hold exact zeros anyway
                        rsum += x_;
                        rsq = fmaf(x_, x_, rsq);
                    }
                    if (store_conv) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) sts128(my_row + (((uint32_t)j ^ sw) << 4), pack8(v + 8 * j));
                        __syncwarp();
                        const int col = t.n0 + cb + seg * 8;
                        if (col < col_end) {
                            __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(ep.out);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int rr = 8 * i + (lane >> 2);
                                if (okmask & (1u << i)) {
                                    const uint4 val = lds128(wb + (uint32_t)rr * 64u + (((uint32_t)seg ^ (uint32_t)((rr >> 1) & 3)) << 4));
                                    *reinterpret_cast<uint4*>(obase + (long long)pix[i] * ep.ldo + col) = val;
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                s_xs[(wi * 32 + lane) * 2] = rsum;
                s_xs[(wi * 32 + lane) * 2 + 1] = rsq;
                asm volatile("bar.sync %0, 96;" ::"r"(2 + q) : "memory");
                float S1 = 0.f, S2 = 0.f;
#pragma unroll
                for (int k = 0; k < EPI_WARPS / 4; ++k) {
                    S1 += s_xs[((q + 4 * k) * 32 + lane) * 2];
                    S2 += s_xs[((q + 4 * k) * 32 + lane) * 2 + 1];
                }
                const float inv = 1.f / (float)inner;
                const float mean = S1 * inv;
                const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, S2 * inv), 0.f) + 1e-6f);
                const float shift = -mean * rstd;
                // phase 2: normalise, GELU, project.  Specialised on the head's output count: the per-column parameters are
                // broadcast shared-memory reads, and shared-memory bandwidth is what the MMA operand fetch of the next tile
                // lives on (a 1-output head reading 4 weight rows cost the GEMM 10 % of its rate).
                float dot[4] = {0.f, 0.f, 0.f, 0.f};
                auto phase2 = [&](auto oc) {
                    constexpr int OO = decltype(oc)::value;
                    for (int c = cpart; c < chunks; c += EPI_WARPS / 4) {
                        const int cb = c * 32;
                        if (cb >= inner || (p.skip_tma & 64)) break;   // debug bit 6: no phase 2
                        uint32_t acc[32];
                        tc_ld32(taddr + (uint32_t)cb, acc);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(s_par + 512 + cb + j);
                            const float4 g4 = *reinterpret_cast<const float4*>(s_par + cb + j);
                            const float4 e4 = *reinterpret_cast<const float4*>(s_par + 256 + cb + j);
                            float ge[4];
                            ge[0] = vk_gelu(fmaf(fmaf(__uint_as_float(acc[j]) + b4.x, rstd, shift), g4.x, e4.x));
                            ge[1] = vk_gelu(fmaf(fmaf(__uint_as_float(acc[j + 1]) + b4.y, rstd, shift), g4.y, e4.y));
                            ge[2] = vk_gelu(fmaf(fmaf(__uint_as_float(acc[j + 2]) + b4.z, rstd, shift), g4.z, e4.z));
                            ge[3] = vk_gelu(fmaf(fmaf(__uint_as_float(acc[j + 3]) + b4.w, rstd, shift), g4.w, e4.w));
#pragma unroll
                            for (int o = 0; o < OO; ++o) {
                                const float4 w4 = *reinterpret_cast<const float4*>(s_par + (3 + o) * 256 + cb + j);   // zero beyond inner
                                dot[o] = fmaf(ge[0], w4.x, dot[o]);
                                dot[o] = fmaf(ge[1], w4.y, dot[o]);
                                dot[o] = fmaf(ge[2], w4.z, dot[o]);
                                dot[o] = fmaf(ge[3], w4.w, dot[o]);
                            }
                        }
                    }
                };
                if (O == 1) phase2(std::integral_constant<int, 1>{});
                else if (O == 2) phase2(std::integral_constant<int, 2>{});
                else phase2(std::integral_constant<int, 4>{});
#pragma unroll
                for (int o = 0; o < 4; ++o) s_xd[(wi * 32 + lane) * 4 + o] = dot[o];
                asm volatile("bar.sync %0, 96;" ::"r"(2 + q) : "memory");
                if (cpart == 0 && row_ok) {
                    const int yy = t.y0 + (r >> p.bw_shift), xx = t.x0 + (r & (p.BW - 1));
                    const long long pin = (long long)yy * p.W + xx;
                    for (int o = 0; o < O; ++o) {
                        float val = __ldg(ht.b2[head] + o);
#pragma unroll
                        for (int k = 0; k < EPI_WARPS / 4; ++k) val += s_xd[((q + 4 * k) * 32 + lane) * 4 + o];
                        if (ht.softplus[head]) val = vk_softplus(val);
                        ht.out[head][((long long)t.b * O + o) * ht.pixels_per_image + pin] = val;
                    }
                }
                // the exchange buffers are rewritten only after the next tile's phase 1, which every warp of the quarter
                // reaches after this point; the cpart-0 reader above is ordered by the next tile's first named barrier
            } else if (p.staged_store) {
                staged_tile(typename KindOf<0>::type{});
            } else
            for (int c = cpart; c < chunks; c += EPI_WARPS / 4) {
                uint32_t acc[32];
                tc_ld32(taddr + (uint32_t)(c * 32), acc);
                if (p.mode == 1 && p.tn_vec) {
                    // weight gradient with contiguous J (the MLP / 1x1 layers): a lane holds 32 columns of ONE row, so scalar
                    // atomics would touch 32 rows per instruction; the chunk goes through the warp's staging tile 16 columns
                    // at a time and leaves as 16-byte reductions, four lanes per 64-byte row segment
                    const uint32_t wb = epi_base + (uint32_t)(warp - 4) * (uint32_t)EPI_TILE_BYTES;
                    const uint32_t my_row = wb + (uint32_t)lane * 64u;
                    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
                    float* obase = reinterpret_cast<float*>(ep.out) + (long long)t.tap * ep.tn_s_tap;
                    const int seg = lane & 3;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            sts128(my_row + (((uint32_t)j ^ sw) << 4),
                                   make_uint4(acc[16 * h + 4 * j], acc[16 * h + 4 * j + 1], acc[16 * h + 4 * j + 2], acc[16 * h + 4 * j + 3]));
                        __syncwarp();
                        const int cl = c * 32 + 16 * h + seg * 4;          // column inside the N tile
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = 8 * i + (lane >> 2);
                            const int irow = t.i0 + q * 32 + rr;
                            if (irow < p.I && t.n0 + cl < p.N && cl < BN && !(p.skip_tma & 256)) {   // debug bit 8: no atomics
                                const uint4 val = lds128(wb + (uint32_t)rr * 64u + (((uint32_t)seg ^ (uint32_t)((rr >> 1) & 3)) << 4));
                                red_add_v4(obase + (long long)irow * ep.tn_s_i + t.n0 + cl, val);
                            }
                        }
                        __syncwarp();
                    }
                    continue;
                }
                if (!row_ok) continue;
                const int nbase = t.n0 + c * 32;
                if (vec_ok && (nbase + 32 <= p.N) && (c * 32 + 32 <= BN) && ((nbase & 7) == 0)) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                    if (ep.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + nbase + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    }
                    if (ep.out_pre) {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out_pre) + row * ep.ld_pre + nbase;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            VkVec<__nv_bfloat16> pk;
                            pk.pack(v + 8 * j);
                            pk.store(o + 8 * j);
                        }
                    }
                    if (ep.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = vk_gelu(v[j]);
                    } else if (ep.act == 2) {
                        const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + row * ep.ld_aux + nbase;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            VkVec<__nv_bfloat16> pk;
                            pk.load(ap + 8 * j);
                            float f[8];
                            pk.unpack(f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[8 * j + e] *= vk_gelu_grad(f[e]);
                        }
                    }
                    if (ep.col_scale) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 s4 = __ldg(reinterpret_cast<const float4*>(ep.col_scale + nbase + j));
                            v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
                        }
                    }
                    if (ep.row_scale) {
                        const float rs = __ldg(ep.row_scale + (row / ep.rows_per_group));
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= rs;
                    }
                    if (ep.residual) {
                        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + row * ep.ld_res + nbase;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            VkVec<__nv_bfloat16> pk;
                            pk.load(rp + 8 * j);
                            float f[8];
                            pk.unpack(f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[8 * j + e] += f[e];
                        }
                    }
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + row * ep.ldo + nbase;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        VkVec<__nv_bfloat16> pk;
                        pk.pack(v + 8 * j);
                        pk.store(o + 8 * j);
                    }
                } else if (p.mode == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nbase + j;
                        if (n < p.N && (c * 32 + j) < BN && !(p.skip_tma & 256)) vk_epilogue_store_tn(ep, t.tap, t.i0 + r, n, __uint_as_float(acc[j]));   // debug bit 8: no atomics
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nbase + j;
                        if (n < p.N && (c * 32 + j) < BN) {
                            const float v = vk_epilogue_value<__nv_bfloat16>(ep, row, n, __uint_as_float(acc[j]));
                            vk_epilogue_store<__nv_bfloat16>(ep, row, n, v);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                               // one arrival per warp (384 same-address arrivals serialise)
                if (cta_rank) mbar_arrive_remote(mapa_cta(tempty_bar(as), 0u));   // pair mode: the leader's MMA thread waits for both
                else mbar_arrive(tempty_bar(as));
            }
            if (++as == 2) { as = 0; aph ^= 1u; }
        }
        if (tma_pending && lane == 0) bulk_wait_all();      // the staging tiles outlive their bulk stores
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CTA2) cluster_sync_all();  // nobody signals the peer's barriers / reads its tiles after this point
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// bf16 tensor map, SWIZZLE_128B, inner box = 64 elements (128 B), zero OOB fill.
int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = get_encode_fn();
    VK_REQUIRE(fn != nullptr, VKOCR_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VK_REQUIRE(r == CUDA_SUCCESS, VKOCR_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u",
               (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return VKOCR_OK;
}

// NHWC activation map: dims (C, W, H, B), pixel stride ld (elements).
int encode_nhwc(CUtensorMap* map, const void* base, int C, int W, int H, int B, long long ld, int bw, int bh) {
    VK_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, VKOCR_BAD_ALIGN, "activation base not 16-byte aligned");
    VK_REQUIRE(((ld * 2) & 15) == 0, VKOCR_BAD_ALIGN, "activation pixel stride %lld not a multiple of 8 elements", ld);
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * W, (uint64_t)ld * 2 * W * H};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
    return encode_map(map, base, 4, dims, str, box);
}

// Output map of the staged epilogue: the same pixel grid, `N` columns, boxes of 32 columns x the 32 pixels of one epilogue warp
// (bw x bh of them: a 32-pixel piece of a tile row, or 32 / BW whole tile rows) in the 64-byte-row swizzle of the staging tiles.
int encode_nhwc_store(CUtensorMap* map, const void* base, int N, int W, int H, int B, long long ld, int BW) {
    const int bw = BW < 32 ? BW : 32;
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * W, (uint64_t)ld * 2 * W * H};
    const uint32_t box[4] = {32, (uint32_t)bw, (uint32_t)(32 / bw), 1};
    return encode_map(map, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

void pick_box(int W, int H, int pixels, int* bw_out, int* bh_out) {
    // choose the BW x BH == pixels box that wastes the fewest out-of-image pixels
    long long best = -1;
    for (int bw = pixels; bw >= 1; bw >>= 1) {
        const int bh = pixels / bw;
        if (bw > 256 || bh > 256) continue;
        const long long cover = (long long)vk_cdiv(W, bw) * bw * (long long)vk_cdiv(H, bh) * bh;
        if (best < 0 || cover < best) {
            best = cover;
            *bw_out = bw;
            *bh_out = bh;
        }
    }
}

typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
template <int... I>
struct KernelTable {
    static TcKernel get(bool cta2, int epi) {
        static const TcKernel one[] = {vkocr_gemm_tc_kernel<false, I>...};
        static const TcKernel two[] = {vkocr_gemm_tc_kernel<true, I>...};
        return cta2 ? two[epi] : one[epi];
    }
};
using Kernels = KernelTable<0, 1, 2, 3, 4, 5, 6>;
static_assert(NUM_EPI == 7, "KernelTable lists every epilogue kind");

int launch(const CUtensorMap& mapA, const CUtensorMap& mapB, TcParams& p, cudaStream_t stream, const CUtensorMap* mapO_in = nullptr,
           const CUtensorMap* mapP_in = nullptr) {
    // output maps of the bulk-store epilogue; the input map stands in where there is none (never dereferenced then)
    const CUtensorMap& mapO = mapO_in ? *mapO_in : mapA;
    const CUtensorMap& mapP = mapP_in ? *mapP_in : mapO;
    static bool attr_set = false;
    const int smem = MAX_SMEM + 1024 + EPI_WARPS * EPI_TILE_BYTES + HEAD_PAR_BYTES + HEAD_XCH_BYTES + 256;
    if (!attr_set) {
        for (int c2 = 0; c2 < 2; ++c2)
            for (int e = 0; e < NUM_EPI; ++e) {
                cudaError_t err = cudaFuncSetAttribute(reinterpret_cast<const void*>(Kernels::get(c2 != 0, e)),
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                VK_REQUIRE(err == cudaSuccess, VKOCR_CUDA_ERROR, "cudaFuncSetAttribute(smem=%d, pair %d, kind %d): %s", smem, c2, e,
                           cudaGetErrorString(err));
            }
        attr_set = true;
    }
    // the specialised instances only hold the staged NT epilogue
    if (!(p.mode == 0 && p.staged_store && !p.head_mode) || p.epi_kind < 0 || p.epi_kind >= NUM_EPI) p.epi_kind = 0;
    if (const char* e = getenv("VKOCR_EPI_GENERIC")) { if (atoi(e)) p.epi_kind = 0; }
    const TcKernel kernel = Kernels::get(p.cta2 != 0, p.epi_kind);
    VK_REQUIRE(p.units < (1LL << 31), VKOCR_BAD_SHAPE, "gemm_tc: %lld work units", p.units);
    p.fd_n = make_fastdiv((uint32_t)p.n_tiles);
    p.fd_x = make_fastdiv((uint32_t)(p.tiles_x > 0 ? p.tiles_x : 1));
    p.fd_y = make_fastdiv((uint32_t)(p.tiles_y > 0 ? p.tiles_y : 1));
    p.fd_i = make_fastdiv((uint32_t)(p.i_tiles > 0 ? p.i_tiles : 1));
    p.fd_taps = make_fastdiv((uint32_t)(p.ks * p.ks));
    p.fd_rpg = make_fastdiv((uint32_t)(p.ep.rows_per_group > 0 ? p.ep.rows_per_group : 1));
    p.num_stages = (MAX_SMEM - (p.prefetch_extra ? PREFETCH_BYTES : 0)) / p.stage_bytes;
    if (p.num_stages > 8) p.num_stages = 8;
    if (p.mode == 1)
        if (const char* e = getenv("VKOCR_TN_STAGES")) p.num_stages = atoi(e) < p.num_stages ? atoi(e) : p.num_stages;
    if (const char* e = getenv("VKOCR_DEBUG_SKIP_TMA")) p.skip_tma = atoi(e);
    p.no_acc_prefetch = getenv("VKOCR_NO_ACC_PREFETCH") != nullptr;
    const long long sms = vkocr_sm_count();
    if (p.cta2) {
        // one cluster of two CTAs per SM pair; in head mode the cluster count is a multiple of the head count so that a
        // cluster keeps its head (and the head's parameters in shared memory) for its whole life
        long long clusters = sms / 2;
        if (clusters > p.units / 2) clusters = p.units / 2;
        if (p.head_mode && clusters >= p.n_tiles) clusters = clusters / p.n_tiles * p.n_tiles;
        if (clusters <= 0) return VKOCR_OK;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * clusters));
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, mapA, mapB, mapO, mapP, p);
        VK_REQUIRE(e == cudaSuccess, VKOCR_CUDA_ERROR, "cudaLaunchKernelEx(cluster 2): %s", cudaGetErrorString(e));
        VK_CHECK_LAUNCH("vkocr_gemm_tc_kernel");
        return VKOCR_OK;
    }
    const int grid = (int)(p.units < sms ? p.units : sms);
    if (grid <= 0) return VKOCR_OK;
    kernel<<<grid, NUM_THREADS, smem, stream>>>(mapA, mapB, mapO, mapP, p);
    VK_CHECK_LAUNCH("vkocr_gemm_tc_kernel");
    return VKOCR_OK;
}

}  // namespace

// NT: D[m,n] = sum_{tap,c} X[pix(m)+off(tap), c] * Wp[n, tap*c_pad + c]   (bf16 in, fp32 accumulate)
int vkocr_gemm_tc_nt(const void* x, const VkocrConvGeom* g, const void* w_packed, int N, const VkocrEpilogue* ep,
                     const VkocrHeadTail* heads, cudaStream_t stream) {
    VK_REQUIRE(g->ks == 1 || g->ks == 3 || g->ks == 5, VKOCR_BAD_SHAPE, "gemm_tc_nt: kernel size %d", g->ks);
    VK_REQUIRE(g->c_pad % BK == 0 && g->c_pad >= g->C, VKOCR_BAD_SHAPE, "gemm_tc_nt: c_pad %d (C %d)", g->c_pad, g->C);
    VK_REQUIRE(N >= 1, VKOCR_BAD_SHAPE, "gemm_tc_nt: N %d", N);
    if ((long long)g->batch * g->H * g->W == 0) return VKOCR_OK;
    TcParams p{};
    p.mode = 0;
    p.batch = g->batch; p.H = g->H; p.W = g->W; p.ks = g->ks;
    pick_box(g->W, g->H, BM, &p.BW, &p.BH);
    for (p.bw_shift = 0; (1 << p.bw_shift) < p.BW; ++p.bw_shift) {}
    p.tiles_x = vk_cdiv(g->W, p.BW);
    p.tiles_y = vk_cdiv(g->H, p.BH);
    p.kb_per_tap = g->c_pad / BK;
    p.N = N;
    // N tile: multiple of 16, <= 256, splitting N as evenly as possible
    const int n_tiles = vk_cdiv(N, 256);
    p.BN = ((vk_cdiv(N, n_tiles) + 15) / 16) * 16;
    // prefer a multiple of 64 (whole 128-byte TMA store rows) when it does not add padded columns: 1152 -> 6 x 192
    for (int cand = 256; cand >= 128 && n_tiles > 1; cand -= 64)
        if ((long long)vk_cdiv(N, cand) * cand <= (long long)vk_cdiv(N, p.BN) * p.BN) { p.BN = cand; break; }
    if (heads) {
        p.BN = heads->slot;          // one head per N tile
        p.head_mode = 1;
        p.ht = *heads;
    }
    p.n_tiles = vk_cdiv(N, p.BN);
    const long long pix_tiles = (long long)p.batch * p.tiles_y * p.tiles_x;
    // CTA pairs (cta_group::2) for the long-K GEMMs: with one CTA per tile every operand byte is written to shared memory
    // once (TMA) and read once (MMA), ~200 B/clk at full MMA rate against the SM's 128 B/clk -- the tensor pipe sat at
    // 62 % (ncu).  A pair shares the weight tile: each CTA stages its 128 pixel rows and HALF of the weight rows.
    p.cta2 = (pix_tiles % 2 == 0) && (p.BN % 16 == 0) && (g->ks * g->ks * (g->c_pad / BK) >= 16) && vkocr_sm_count() % 2 == 0;
    if (const char* e = getenv("VKOCR_CTA2")) {   // 0: never, 2: whenever the shape allows it (experiments)
        if (atoi(e) == 2) p.cta2 = (pix_tiles % 2 == 0) && (p.BN % 16 == 0) && vkocr_sm_count() % 2 == 0;
        else p.cta2 = p.cta2 && atoi(e) != 0;
    }
    p.stage_bytes = A_BYTES + (p.cta2 ? p.BN / 2 : p.BN) * 128;
    p.units = pix_tiles * p.n_tiles;
    p.ep = *ep;
    CUtensorMap mapA, mapB;
    int rc = encode_nhwc(&mapA, x, g->C, g->W, g->H, g->batch, g->ld_x, p.BW, p.BH);
    if (rc) return rc;
    const long long kw = (long long)g->ks * g->ks * g->c_pad;
    const uint64_t dims[2] = {(uint64_t)kw, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)kw * 2};
    const uint32_t box[2] = {64, (uint32_t)(p.cta2 ? p.BN / 2 : p.BN)};
    VK_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, VKOCR_BAD_ALIGN, "packed weight not 16-byte aligned");
    rc = encode_map(&mapB, w_packed, 2, dims, str, box);
    if (rc) return rc;
    // aligned bf16 outputs (every hot-path call) take the staged epilogue; anything else the generic per-element one
    auto aligned = [](const void* ptr, long long ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld & 7) == 0; };
    p.M = g->W;
    p.staged_store = !ep->out_f32 && !ep->accumulate && (N % 8 == 0) && aligned(ep->out, ep->ldo) &&
                     (!ep->out_pre || aligned(ep->out_pre, ep->ld_pre)) && (!ep->residual || aligned(ep->residual, ep->ld_res)) &&
                     ((ep->act != 2 && ep->act != 4) || aligned(ep->aux, ep->ld_aux)) && (long long)g->batch * g->H * g->W < (1LL << 31);
    // short-K GEMMs with a residual / aux operand: one 24 KB set of prefetch tiles comes out of the operand-stage budget
    p.prefetch_extra = p.staged_store && (ep->residual || ep->act == 2 || ep->act == 4) && !p.head_mode &&
                       g->ks * g->ks * p.kb_per_tap <= 12 && (MAX_SMEM - PREFETCH_BYTES) / p.stage_bytes >= 2;
    if (const char* e = getenv("VKOCR_NO_PREFETCH")) { if (atoi(e)) p.prefetch_extra = 0; }
    {
        const bool b = ep->bias != nullptr, cs = ep->col_scale != nullptr, res = ep->residual != nullptr, pre = ep->out_pre != nullptr;
        const bool rsc = ep->row_scale != nullptr;
        p.epi_kind = 0;
        if (ep->act == 0 && !b && !cs && !res && !pre && !rsc) p.epi_kind = 1;
        else if (ep->act == 0 && b && !cs && !res && !pre && !rsc) p.epi_kind = 2;
        else if (ep->act == 3 && b && !cs && !res && pre) p.epi_kind = 3;
        else if (ep->act == 1 && b && !cs && !res && !pre && !rsc) p.epi_kind = 4;
        else if (ep->act == 0 && b && cs && res && !pre) p.epi_kind = 5;
        else if (ep->act == 4 && !b && !cs && !res && !pre && !rsc) p.epi_kind = 6;
    }
    // the staged tiles leave through the copy engine (bulk tensor stores): no per-lane store instructions, clipping for free
    CUtensorMap mapO, mapP;
    // (a box is 32 whole columns: N tiles that are not a multiple of 32 wide would spill into their neighbour's columns; with a
    // residual / aux operand prefetched by cp.async the proxy fence in front of the bulk store would wait for those copies --
    // measured 10 % slower -- so these keep the per-lane stores)
    p.tma_store = p.staged_store && !p.head_mode && !p.prefetch_extra && (p.BN % 32 == 0 || p.n_tiles == 1) &&
                  getenv("VKOCR_NO_TMA_STORE") == nullptr;
    if (p.tma_store) {
        rc = encode_nhwc_store(&mapO, ep->out, N, g->W, g->H, g->batch, ep->ldo, p.BW);
        if (rc) return rc;
        if (ep->out_pre) {
            rc = encode_nhwc_store(&mapP, ep->out_pre, N, g->W, g->H, g->batch, ep->ld_pre, p.BW);
            if (rc) return rc;
        }
    }
    return launch(mapA, mapB, p, stream, p.tma_store ? &mapO : nullptr, (p.tma_store && ep->out_pre) ? &mapP : nullptr);
}

// TN: G[tap,i,j] += sum_pix P[pix,i] * Q[pix+off(tap), j]   (bf16 in, fp32 out accumulated with red.add)
// ep->out must be fp32 and is accumulated into at the tn_s_* strides (caller zeroes it).
//
// Either operand may take the M (128-row tile) side of the MMA: G[tap,i,j] = sum_pix Q[pix,j] * P[pix-off(tap),i] is the
// same contraction with the roles swapped and the tap mirrored.  A tcgen05.mma costs about the same for every N <= ~192,
// so the orientation with the fewest (tile x instruction) products wins, e.g. I=832, J=384: 7x2 tiles of 128x192 (with
// 7 % row padding) versus 3x4 tiles of 128x208 (no padding).
namespace {
int tn_tile_n(int J, int* n_tiles) {
    const int nt = vk_cdiv(J, 256);
    const int bn = ((vk_cdiv(J, nt) + 15) / 16) * 16;
    *n_tiles = vk_cdiv(J, bn);
    return bn;
}
long long tn_cost(int I, int J) {
    int nt;
    const int bn = tn_tile_n(J, &nt);
    return (long long)vk_cdiv(I, BM) * nt * (bn > 176 ? bn : 176);
}
}  // namespace

int vkocr_gemm_tc_tn(const void* pmat, const VkocrConvGeom* g, const void* qmat, int J, long long ld_q,
                     const VkocrEpilogue* ep_in, cudaStream_t stream) {
    VK_REQUIRE(g->ks == 1 || g->ks == 3 || g->ks == 5, VKOCR_BAD_SHAPE, "gemm_tc_tn: kernel size %d", g->ks);
    VK_REQUIRE(ep_in->out_f32 && ep_in->accumulate, VKOCR_BAD_ARGUMENT, "gemm_tc_tn: output must be fp32 accumulate");
    if ((long long)g->batch * g->H * g->W == 0) return VKOCR_OK;
    int I = g->C;
    long long ld_p = g->ld_x;
    VkocrEpilogue ep = *ep_in;
    // measured on B200 (I=832, J=384, 3x3): both orientations run within 2 % of each other once the split fills whole
    // waves, so the swap is only taken when it removes at least a fifth of the tile work
    bool swap = tn_cost(J, I) * 5 < tn_cost(I, J) * 4;
    if (const char* e = getenv("VKOCR_TN_SWAP")) swap = atoi(e) != 0;
    if (swap) {
        const int taps = g->ks * g->ks;
        ep.out = reinterpret_cast<float*>(ep.out) + (long long)(taps - 1) * ep.tn_s_tap;
        ep.tn_s_tap = -ep.tn_s_tap;
        const long long t = ep.tn_s_i;
        ep.tn_s_i = ep.tn_s_j;
        ep.tn_s_j = t;
        const void* tp = pmat; pmat = qmat; qmat = tp;
        const int ti = I; I = J; J = ti;
        const long long tl = ld_p; ld_p = ld_q; ld_q = tl;
    }
    TcParams p{};
    p.mode = 1;
    p.batch = g->batch; p.H = g->H; p.W = g->W; p.ks = g->ks;
    pick_box(g->W, g->H, 64, &p.BW, &p.BH);
    p.tiles_x = vk_cdiv(g->W, p.BW);
    p.tiles_y = vk_cdiv(g->H, p.BH);
    p.I = I;
    p.i_tiles = vk_cdiv(p.I, BM);
    p.N = J;
    p.BN = tn_tile_n(J, &p.n_tiles);
    if (const char* e = getenv("VKOCR_TN_BN")) { p.BN = atoi(e); p.n_tiles = vk_cdiv(J, p.BN); }
    p.stage_bytes = A_BYTES + vk_cdiv(p.BN, 64) * 8192;   // Q arrives in 64-channel boxes
    const long long out_tiles = (long long)p.ks * p.ks * p.i_tiles * p.n_tiles;
    const long long pix_tiles = (long long)p.batch * p.tiles_y * p.tiles_x;
    // Split the pixel range so that the units fill whole waves of the persistent grid: every unit costs the same, so the
    // kernel time is ceil(units / SMs) unit times -- 378 units on 148 SMs would idle 15 % of the machine in its last wave.
    // Candidates: 1..8 waves' worth of splits, keeping >= 32 k-blocks per unit; the fewest splits (fewest fp32 atomics)
    // whose wave efficiency is within 5 % of the best candidate's wins.
    const long long sms = vkocr_sm_count();
    long long max_splits = pix_tiles / 32;           // every unit ends with a 128 x BN tile of fp32 atomics: amortise it
    if (max_splits < 1) max_splits = 1;
    double best_eff = -1.0;
    auto wave_eff = [&](long long cand) {
        const long long units = cand * out_tiles;
        return (double)units / (double)(((units + sms - 1) / sms) * sms);
    };
    for (long long cand = 1; cand <= max_splits && cand * out_tiles <= 8 * sms + out_tiles; ++cand)
        if (wave_eff(cand) > best_eff) best_eff = wave_eff(cand);
    long long splits = 1;
    for (long long cand = 1; cand <= max_splits && cand * out_tiles <= 8 * sms + out_tiles; ++cand)
        if (wave_eff(cand) >= 0.95 * best_eff) { splits = cand; break; }   // fewest splits within 5 % of the best fill
    if (const char* e = getenv("VKOCR_TN_SPLITS")) splits = atoi(e);
    p.splits = (int)splits;
    p.units = out_tiles * splits;
    p.ep = ep;
    p.tn_vec = ep.tn_s_j == 1 && ep.tn_s_i % 4 == 0 && ep.tn_s_tap % 4 == 0 && J % 4 == 0 &&
               (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && getenv("VKOCR_TN_SCALAR_RED") == nullptr;
    CUtensorMap mapA, mapB;
    int rc = encode_nhwc(&mapA, pmat, I, g->W, g->H, g->batch, ld_p, p.BW, p.BH);
    if (rc) return rc;
    rc = encode_nhwc(&mapB, qmat, J, g->W, g->H, g->batch, ld_q, p.BW, p.BH);
    if (rc) return rc;
    p.staged_store = 0;
    return launch(mapA, mapB, p, stream);
}
