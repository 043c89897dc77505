// fpm_mma.cuh -- tcgen05 (5th-gen tensor core) correlation for the large pyramid levels.
//
// Same quantity as fpm_corr_rows_kernel -- the exact s32 dot product of every template row with the
// 7 shifted ROI rows below it (IM_Conv_SIMD, /root/reference/src/TemplateMatcher.cpp:461-483, driven by
// the loop at :496-510) -- shaped as a batched integer GEMM so it runs on the tensor cores:
//
//   for every ROI row y:   D_y[e][n] = sum_x  A_y[e][x] * B_y[n][x]        (u8 x u8 -> s32, exact)
//        A_y[e][x] = S_e[y][x]                      M = 128 evals (candidate x angle) per tile
//        B_y[n][x] = T[y-7+jj][x-c],  n = 8c + jj   N = 64 (8 column shifts x 8 template rows), K = w+6
//   => D_y[e][8c+jj] = rowsum[e][tr = y-7+jj][r = 7-jj][c]     (jj = 0 and c = 7 are padding)
//
// A tiles come straight from the ROI buffer and B tiles from 8 pre-shifted copies of the template,
// both through TMA (cp.async.bulk.tensor, 128-byte swizzle, out-of-bounds rows/columns zero filled);
// tcgen05.mma kind::i8 (a_format = b_format = unsigned 8-bit, S32 accumulate) accumulates the K chunks
// of one ROI row in TMEM; the accumulator of row y is drained by 4 epilogue warps (tcgen05.ld) while
// row y+1 is being multiplied (two TMEM buffers).  The s32 sums never exceed 255*255*(w+6) < 2^31.
//
// Output layout "raw": raw[y][e][64] (one 256-byte line per (ROI row, eval)), consumed in ROI-row
// order by fpm_refine_finalize_kernel, which keeps the reference's float32 accumulation order.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp id % 4), warps 6..9 = window statistics (fpm_mm_stats):
// the 7 shifted row sums / square sums of every ROI row, computed from the A tiles while they are in shared memory.
#pragma once
#include <cuda.h>
#include "fpm_common.cuh"

#define MM_M 128
#define MM_N 64
#define MM_KCHUNK 128                  // bytes of K per TMA box / pipeline stage (one 128B swizzle atom)
#define MM_STAGES 6
#define MM_A_BYTES (MM_M * MM_KCHUNK)  // 16 KB
#define MM_B_BYTES (MM_N * MM_KCHUNK)  // 8 KB
#define MM_STAGE_BYTES (MM_A_BYTES + MM_B_BYTES)
#define MM_THREADS 320
#define MM_TMEM_COLS 128               // two accumulator buffers of 64 columns
#define MM_SMEM_BYTES (MM_STAGES * MM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)

// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): dense, no saturate,
// c_format S32 (2) at [4,6), a/b format UINT8 (0), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
#define MM_IDESC ((2u << 4) | ((uint32_t)(MM_N >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24))

namespace fpm_ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// wait of a warp role that is expected to idle for a while: a spinning warp takes issue slots from the warps that do the
// work (the spin loops of the idle roles were a quarter of all instructions issued by fpm_corr_warp_kernel)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], u8 x u8 -> s32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute SmemDescriptor):
// start>>4 at [0,14), LBO>>4 = 1 at [16,30), SBO>>4 = 64 (8 rows x 128 B) at [32,46), version 1 at [46,48),
// layout SWIZZLE_128B (2) at [61,64)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// 32 lanes x 32 bit, 16 consecutive columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// shared-memory reads of TMA-filled stages: volatile asm, so they stay between the full-barrier wait and the
// empty-barrier arrive no matter what the compiler knows about the pointer
__device__ __forceinline__ uint4 lds128(uint32_t saddr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return (int)v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace fpm_ptx

// Window statistics of the A tiles while they sit in shared memory (warps 6..9 of both tensor-core kernels):
// thread m owns eval row m of the tile, walks its 128-byte rows chunk by chunk (the 128-byte swizzle permutes
// 16-byte chunks inside a row: chunk j of row m lives at chunk j ^ (m & 7)) and produces, per ROI row, the 7
// shifted window sums S_c = sum_{x<tw} roi[y][x+c] and square sums from the base sum over [0, tw) plus the
// 6 head / 6 tail bytes.  kAllRows: store every row (row-split kernel); otherwise store only the first / last
// 6 rows and return the totals over all rows (fused kernel).  Each consumed stage is released on empty_bar.
template <bool kAllRows>
__device__ __forceinline__ void fpm_mm_stats(uint32_t base_u32, uint32_t bar0, int m, int lane, int e, bool live,
                                             int y_begin, int y_end, int rh, int tw, int th, int nk,
                                             int32_t* __restrict__ rowS, int32_t* __restrict__ rowQ,
                                             long long* ts, long long* tq)
{
    using namespace fpm_ptx;
    const uint32_t sw = (uint32_t)(m & 7);
    int it = 0;
    for (int y = y_begin; y < y_end; y++) {
        uint32_t s0 = 0, q0 = 0;
        int hd[6], tl[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { hd[i] = 0; tl[i] = 0; }
        for (int k = 0; k < nk; k++, it++) {
            const int s = it % MM_STAGES;
            const uint32_t ph = (it / MM_STAGES) & 1;
            mbar_wait(bar0 + 8u * s, ph);                                         // full_bar(s)
            const uint32_t rowp = base_u32 + (uint32_t)s * MM_STAGE_BYTES + (uint32_t)m * MM_KCHUNK;
            const int xb = k * MM_KCHUNK;
            // chunks that hold window columns or the 6 tail bytes; every loaded value is consumed before the stage
            // is released (no load may still be in flight when the empty barrier is signalled)
            const int nj = min(MM_KCHUNK / 16, (tw + FPM_ROI_PAD - xb + 15) >> 4);
            for (int j = 0; j < nj; j++) {
                const uint4 v = lds128(rowp + (((uint32_t)j ^ sw) << 4));
                const int rem = tw - (xb + 16 * j);                                 // window bytes in this chunk (may be <= 0)
                if (k == 0 && j == 0) {
                    hd[0] = v.x & 0xff; hd[1] = (v.x >> 8) & 0xff; hd[2] = (v.x >> 16) & 0xff; hd[3] = v.x >> 24;
                    hd[4] = v.y & 0xff; hd[5] = (v.y >> 8) & 0xff;
                }
                if (rem >= 16) {
                    s0 = __dp4a(v.x, 0x01010101u, s0); q0 = __dp4a(v.x, v.x, q0);
                    s0 = __dp4a(v.y, 0x01010101u, s0); q0 = __dp4a(v.y, v.y, q0);
                    s0 = __dp4a(v.z, 0x01010101u, s0); q0 = __dp4a(v.z, v.z, q0);
                    s0 = __dp4a(v.w, 0x01010101u, s0); q0 = __dp4a(v.w, v.w, q0);
                } else {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        const int nb = rem - 4 * jj;
                        const uint32_t mk = nb >= 4 ? w[jj] : (nb <= 0 ? 0u : (w[jj] & (0xffffffffu >> (8 * (4 - nb)))));
                        s0 = __dp4a(mk, 0x01010101u, s0);
                        q0 = __dp4a(mk, mk, q0);
                    }
                    // tail byte i sits at position rem + i of this chunk when that is inside [0, 16)
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        const int pos = rem + i;
                        if (pos >= 0 && pos < 16) {
                            const uint32_t wsel = pos < 8 ? (pos < 4 ? v.x : v.y) : (pos < 12 ? v.z : v.w);
                            tl[i] = (int)__byte_perm(wsel, 0u, 0x4440u | (uint32_t)(pos & 3));
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8u * (MM_STAGES + s));              // empty_bar(s)
        }
        int sc = (int)s0, qc = (int)q0;
        const bool store = live && (kAllRows || y < FPM_ROI_PAD || y >= th);
        int vs[FPM_WSTRIDE], vq[FPM_WSTRIDE];
        vs[FPM_NSHIFT] = 0; vq[FPM_NSHIFT] = 0;
#pragma unroll
        for (int c = 0; c < FPM_NSHIFT; c++) {
            if (c > 0) { sc += tl[c - 1] - hd[c - 1]; qc += tl[c - 1] * tl[c - 1] - hd[c - 1] * hd[c - 1]; }
            if (!kAllRows) { ts[c] += sc; tq[c] += qc; }
            vs[c] = sc; vq[c] = qc;
        }
        if (store) {
            // one aligned 32-byte record per (eval, row): two 128-bit stores each (14 scattered 4-byte stores per row and
            // thread had the statistics warps waiting on the store queue: a third of the kernel's stall samples)
            int4* ps = reinterpret_cast<int4*>(rowS + ((size_t)e * rh + y) * FPM_WSTRIDE);
            int4* pq = reinterpret_cast<int4*>(rowQ + ((size_t)e * rh + y) * FPM_WSTRIDE);
            ps[0] = make_int4(vs[0], vs[1], vs[2], vs[3]); ps[1] = make_int4(vs[4], vs[5], vs[6], vs[7]);
            pq[0] = make_int4(vq[0], vq[1], vq[2], vq[3]); pq[1] = make_int4(vq[4], vq[5], vq[6], vq[7]);
        }
    }
}

// grid: (row_chunks, m_tiles); CTA (bx, by) handles ROI rows [bx*rows_per_cta, ...) of evals [128*by, 128*by+128)
__global__ void __launch_bounds__(MM_THREADS, 1)
fpm_corr_mma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    int n_evals, int e_pad, int rh, int tw, int th, int k_bytes, int rows_per_cta, int32_t* __restrict__ raw,
                    int32_t* __restrict__ rowS, int32_t* __restrict__ rowQ, const int* __restrict__ n_cands_dev, int n_ang)
{
    using namespace fpm_ptx;
    if (n_cands_dev) {                                       // live evals known only on the device; surplus tiles leave at once
        n_evals = min(n_evals, *n_cands_dev * n_ang);
        if ((int)blockIdx.y * MM_M >= n_evals) return;
    }
    extern __shared__ uint8_t mm_smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles
    const uint32_t base_u32 = (smem_u32(mm_smem_raw) + 1023u) & ~1023u;
    uint8_t* base = mm_smem_raw + (base_u32 - smem_u32(mm_smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + MM_STAGES * MM_STAGE_BYTES);
    // barriers: full[0..S), empty[S..2S), tmem_full[2S..2S+2), tmem_empty[2S+2..2S+4)
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (MM_STAGES + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MM_STAGES + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MM_STAGES + 2 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MM_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y_begin = blockIdx.x * rows_per_cta;
    const int y_end = min(rh, y_begin + rows_per_cta);
    const int e0 = blockIdx.y * MM_M;
    const int nk = (k_bytes + MM_KCHUNK - 1) / MM_KCHUNK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < MM_STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1 + 4); }   // MMA commit + 4 statistics warps
        for (int b = 0; b < 2; b++) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), MM_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (y_begin < y_end) {
        if (warp == 0) {
            // ===== TMA producer =====
            if (lane == 0) {
                int it = 0;
                for (int y = y_begin; y < y_end; y++)
                    for (int k = 0; k < nk; k++, it++) {
                        const int s = it % MM_STAGES;
                        const uint32_t ph = (it / MM_STAGES) & 1;
                        mbar_wait(empty_bar(s), ph ^ 1);                 // slot free (first pass returns at once)
                        mbar_expect_tx(full_bar(s), MM_STAGE_BYTES);
                        const uint32_t sa = base_u32 + s * MM_STAGE_BYTES;
                        tma_load_3d(sa, &map_a, full_bar(s), k * MM_KCHUNK, y, e0);                    // 128 evals x 128 B
                        tma_load_3d(sa + MM_A_BYTES, &map_b, full_bar(s), k * MM_KCHUNK, y - 7, 0);    // (8 shifts x 8 rows) x 128 B
                    }
            }
        } else if (warp == 1) {
            // ===== MMA issuer (one thread) =====
            if (lane == 0) {
                int it = 0, yi = 0;
                for (int y = y_begin; y < y_end; y++, yi++) {
                    const int buf = yi & 1;
                    const uint32_t tph = (yi >> 1) & 1;
                    mbar_wait(tempty_bar(buf), tph ^ 1);                 // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + buf * MM_N;
                    for (int k = 0; k < nk; k++, it++) {
                        const int s = it % MM_STAGES;
                        const uint32_t ph = (it / MM_STAGES) & 1;
                        mbar_wait(full_bar(s), ph);                      // TMA bytes have landed
                        tc_fence_after();
                        const uint32_t sa = base_u32 + s * MM_STAGE_BYTES;
                        const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + MM_A_BYTES);
#pragma unroll
                        for (int kk = 0; kk < MM_KCHUNK / 32; kk++)      // K = 32 bytes per instruction
                            mma_i8(d_addr, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), MM_IDESC, (k | kk) ? 1u : 0u);
                        tc_commit(empty_bar(s));                         // frees the smem stage when the MMAs retire
                    }
                    tc_commit(tfull_bar(buf));                           // accumulator of row y complete
                }
            }
        } else if (warp >= 6) {
            // ===== window statistics: warps 6..9 =====
            const int m = (warp - 6) * 32 + lane;
            fpm_mm_stats<true>(base_u32, bar0, m, lane, e0 + m, (e0 + m) < n_evals, y_begin, y_end, rh, tw, th, nk, rowS, rowQ,
                               nullptr, nullptr);
        } else {
            // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
            const int q = warp & 3;
            const int m = q * 32 + lane;                                 // row of the tile = eval e0 + m
            const bool store = (e0 + m) < n_evals;
            int yi = 0;
            for (int y = y_begin; y < y_end; y++, yi++) {
                const int buf = yi & 1;
                const uint32_t tph = (yi >> 1) & 1;
                mbar_wait(tfull_bar(buf), tph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * MM_N;
                uint32_t v[MM_N];
#pragma unroll
                for (int c = 0; c < MM_N / 16; c++) tmem_ld16(taddr + c * 16, v + c * 16);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(buf));             // 4 arrivals (one per epilogue warp)
                if (store) {
                    uint4* o = reinterpret_cast<uint4*>(raw + ((size_t)y * e_pad + e0 + m) * MM_N);
#pragma unroll
                    for (int c = 0; c < MM_N / 4; c++) o[c] = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, MM_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Fused variant: one CTA owns 128 evals for ALL ROI rows, so nothing but the 49 scores' numerators
// and the window statistics leaves the SM.
//   * epilogue warps keep the reference's float32 accumulation chain (MatchTemplate's
//     `result += (float)IM_Conv_SIMD(row)` loop, src/TemplateMatcher.cpp:496-510) in registers:
//     accumulator column 8c+jj of ROI row y is template row tr = y-7+jj of score (r = 7-jj, c), and y
//     ascending is tr ascending, so acc[8c+jj] = fadd(acc[8c+jj], float(D_y[e][8c+jj])) walks exactly the
//     reference's order; rows outside the template come back as 0 (TMA zero fill) and adding +0.0f is exact;
//   * 4 statistics warps read the same A tiles from shared memory (one thread per eval row, swizzle
//     undone per 16-byte chunk) for the window sums: 7 shifted sums / square sums per ROI row, their
//     totals over all rows in 64 bit, and the first / last 6 rows per eval (all the 7x7 windows need).
// Warp roles (320 threads): 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue, 6..9 = statistics.
#define FM_THREADS 320
#define FM_SMEM_BYTES MM_SMEM_BYTES

__global__ void __launch_bounds__(FM_THREADS, 1)
fpm_corr_fused_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      int n_evals, int rh, int tw, int th, int k_bytes, float* __restrict__ numer,
                      int32_t* __restrict__ rowS, int32_t* __restrict__ rowQ,
                      long long* __restrict__ totS, long long* __restrict__ totQ, const int* __restrict__ n_cands_dev, int n_ang)
{
    using namespace fpm_ptx;
    if (n_cands_dev) {
        n_evals = min(n_evals, *n_cands_dev * n_ang);
        if ((int)blockIdx.x * MM_M >= n_evals) return;
    }
    extern __shared__ uint8_t mm_smem_raw[];
    const uint32_t base_u32 = (smem_u32(mm_smem_raw) + 1023u) & ~1023u;
    uint8_t* base = mm_smem_raw + (base_u32 - smem_u32(mm_smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + MM_STAGES * MM_STAGE_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (MM_STAGES + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MM_STAGES + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MM_STAGES + 2 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MM_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * MM_M;
    const int nk = (k_bytes + MM_KCHUNK - 1) / MM_KCHUNK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < MM_STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1 + 4); }   // MMA commit + 4 statistics warps
        for (int b = 0; b < 2; b++) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), MM_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int y = 0; y < rh; y++)
                for (int k = 0; k < nk; k++, it++) {
                    const int s = it % MM_STAGES;
                    const uint32_t ph = (it / MM_STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), MM_STAGE_BYTES);
                    const uint32_t sa = base_u32 + s * MM_STAGE_BYTES;
                    tma_load_3d(sa, &map_a, full_bar(s), k * MM_KCHUNK, y, e0);
                    tma_load_3d(sa + MM_A_BYTES, &map_b, full_bar(s), k * MM_KCHUNK, y - 7, 0);
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int it = 0;
            for (int y = 0; y < rh; y++) {
                const int buf = y & 1;
                const uint32_t tph = (y >> 1) & 1;
                mbar_wait(tempty_bar(buf), tph ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + buf * MM_N;
                for (int k = 0; k < nk; k++, it++) {
                    const int s = it % MM_STAGES;
                    const uint32_t ph = (it / MM_STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sa = base_u32 + s * MM_STAGE_BYTES;
                    const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + MM_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < MM_KCHUNK / 32; kk++)
                        mma_i8(d_addr, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), MM_IDESC, (k | kk) ? 1u : 0u);
                    tc_commit(empty_bar(s));
                }
                tc_commit(tfull_bar(buf));
            }
        }
    } else if (warp < 6) {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4; float chain in registers =====
        const int q = warp & 3;
        const int m = q * 32 + lane;
        float acc[56];
#pragma unroll
        for (int i = 0; i < 56; i++) acc[i] = 0.0f;
        for (int y = 0; y < rh; y++) {
            const int buf = y & 1;
            const uint32_t tph = (y >> 1) & 1;
            mbar_wait(tfull_bar(buf), tph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * MM_N;
            uint32_t v[64];
#pragma unroll
            for (int c = 0; c < 4; c++) tmem_ld16(taddr + c * 16, v + c * 16);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
#pragma unroll
            for (int i = 0; i < 56; i++)
                if (i & 7) acc[i] = __fadd_rn(acc[i], __int2float_rn((int)v[i]));      // jj = 0 is padding
        }
        if (e0 + m < n_evals) {
            float4* o = reinterpret_cast<float4*>(numer + (size_t)(e0 + m) * MM_N);
#pragma unroll
            for (int c = 0; c < 14; c++) o[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
    } else {
        // ===== window statistics: warps 6..9, thread = eval row m of the A tile =====
        const int m = (warp - 6) * 32 + lane;
        const int e = e0 + m;
        const bool live = e < n_evals;
        long long ts[FPM_NSHIFT], tq[FPM_NSHIFT];
#pragma unroll
        for (int c = 0; c < FPM_NSHIFT; c++) { ts[c] = 0; tq[c] = 0; }
        fpm_mm_stats<false>(base_u32, bar0, m, lane, e, live, 0, rh, rh, tw, th, nk, rowS, rowQ, ts, tq);
        if (live) {
#pragma unroll
            for (int c = 0; c < FPM_NSHIFT; c++) { totS[(size_t)e * FPM_NSHIFT + c] = ts[c]; totQ[(size_t)e * FPM_NSHIFT + c] = tq[c]; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, MM_TMEM_COLS);
}

// pre-shifted template copies for the B operand: tsh[c][tr][x] = T[tr][x - c]  (0 outside the template)
__global__ void fpm_shift_template_kernel(const uint8_t* __restrict__ tpl, int tw, int th, int tpitch,
                                          uint8_t* __restrict__ tsh, int bpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, tr = blockIdx.y, c = blockIdx.z;
    if (x >= bpitch) return;
    const int sx = x - c;
    tsh[((size_t)c * th + tr) * bpitch + x] = (sx >= 0 && sx < tw) ? tpl[(size_t)tr * tpitch + sx] : 0;
}
