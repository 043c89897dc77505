// fpm_fused.cuh -- fpm_corr_warp_kernel: getRotatedROI (the ROI warpAffine, src/TemplateMatcher.cpp:1074-1090) fused into the
// producer of the tcgen05 correlation (IM_Conv_SIMD row dots, :461-483, loop :496-510) -- the rotated ROI patches never
// exist in HBM.  Same GEMM as fpm_corr_mma_kernel (fpm_mma.cuh):
//
//   for every ROI row y:   D_y[e][8c + jj] = sum_x  S_e[y][x] * T[y - 7 + jj][x - c]       (u8 x u8 -> s32, exact)
//
// but the A operand (row y of 128 ROI patches) is COMPUTED in the SM: gather warps evaluate OpenCV's fixed-point bilinear
// warpAffine (bit for bit, see fpm_warp_kernel) from a source box staged in shared memory and store the pixels straight
// into the MMA's K-major operand layout; B (8 pre-shifted template copies) still arrives by TMA.
//
// One CTA = one unit = 8 ROI rows x 128 eval slots (42 candidates x 3 angles; slots 126, 127 unused) x all K chunks of
// 128 pixels.  The producer walks "visits" v = (chunk k, candidate c): the three angles of a candidate are anchored at the
// same source point, so ONE source box (<= ~13 KB) covers the 3 x 8 x 128 pixels of a visit.  Pipeline of a visit:
//
//   2 geometry warps: OpenCV's adelta / bdelta / X0 / Y0 tables of the visit (fp64 cvRound, exactly like cv::warpAffine),
//                   exact source box from the tile corners                                        -> geomfull[v % 3]
//   4 copier warps: 4-byte cp.async of the box into box[v % 3] (odd word pitch), out-of-image bytes zero filled
//                   (= BORDER_CONSTANT 0, so the gather needs no predicate)                       -> boxfull[v % 3]
//   prefetch warp : prefetch.global.L2 of the boxes a few visits ahead
//   12 gather warps: warp g owns angle g / 4 and rows 2 (g % 4), 2 (g % 4) + 1; lane l the pixels l, l+32, l+64, l+96:
//                   ~22 instructions per pixel, result bytes stored into A stage r (one stage per ROI row of the unit)
//   after the last candidate of a chunk: fence.proxy.async + arrive                                -> afull
//   MMA thread    : per row r: 4 x tcgen05.mma (K = 32) into TMEM accumulator r (8 x 64 columns = all 512), B from a
//                   3-stage TMA ring; tcgen05.commit frees the B stage / the A stages               -> aempty, tfull
//   4 stats/epilogue warps: window row sums of CCOEFF_Denominator (:570-577) from the A stages (thread = eval row); after
//                   the last chunk, TMEM -> raw[y][slot][64] (consumed by fpm_refine_finalize_kernel)
//
// A stages use the NO-swizzle K-major canonical layout (core matrix = 8 rows x 16 B contiguous; LBO = 144 B between
// K-adjacent core matrices, SBO = 1152 B between 8-row groups): byte x of row m lives at
//   (m >> 3) * 1152 + (x >> 4) * 144 + (m & 7) * 16 + (x & 15)
// so the four pixels of a lane are at constant offsets k * 288 from one address, and the 144-byte LBO puts the two
// half-warps of a store in different banks.
#pragma once
#include "fpm_mma.cuh"

#define FW_M 128
#define FW_CANDS 42                         // candidates per eval tile
#define FW_TILE_EVALS (3 * FW_CANDS)        // 126 of the 128 MMA rows
#define FW_R 8                              // ROI rows per unit = TMEM accumulators
#define FW_A_STAGE (FW_M * MM_KCHUNK)        // 16 KB, K-major 128-byte swizzle like the TMA-fed kernels
#define FW_B_STAGES 3
#define FW_BOX_BYTES 18432
#define FW_T 4                              // geometry table ring
#define FW_GATHER_WARPS 12
#define FW_COPY_WARPS 4
#define FW_GEOM_WARPS 2
#define FW_NBOX 3                            // source boxes in flight
#define FW_PF_AHEAD 6                       // visits the L2 prefetcher runs ahead of the geometry warp
// warps: 0 TMA(B), 1 MMA, geometry, copiers, 1 L2 prefetcher, 4 statistics + epilogue, gather
#define FW_W_GEOM 2
#define FW_W_COPY (FW_W_GEOM + FW_GEOM_WARPS)
#define FW_W_PF (FW_W_COPY + FW_COPY_WARPS)
#define FW_W_EPI (FW_W_PF + 1)
#define FW_W_GATHER (FW_W_EPI + 4)
#define FW_THREADS (32 * (FW_W_GATHER + FW_GATHER_WARPS))
#define FW_JOB_BYTES (FW_TILE_EVALS * 6 * 8)   // inverse matrices of the tile's evals, staged once
#define FW_SMEM_BYTES (FW_R * FW_A_STAGE + FW_B_STAGES * MM_B_BYTES + FW_NBOX * FW_BOX_BYTES + FW_T * 3328 + FW_JOB_BYTES + 1024 + 256)

struct FwTables {                           // geometry of one visit (3 angles of a candidate, 8 rows x 128 columns)
    int ad[3][128], bd[3][128];             // cvRound(M00 * x * 1024), cvRound(M10 * x * 1024)
    int X0[3][FW_R], Y0[3][FW_R];           // cvRound((M01 * y + M02) * 1024) + 16 - (box origin << 10)
    int bx0, by0;                           // box origin in the source level (bx0 aligned down to the copy width; may be < 0)
    int pitch;                              // box row pitch in bytes (a multiple of 4, odd number of words)
    int nrows, nch;                         // box rows, copy chunks per row
    int pad[11];
};
static_assert(sizeof(FwTables) == 3328, "FwTables size");

namespace fpm_ptx {
__device__ __forceinline__ void cp_async_zfill(uint32_t saddr, const void* gmem, int bytes, int vec)
{
    // `vec`-byte copy of which only the first `bytes` come from global memory, the rest is zero filled
    if (vec == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gmem), "r"(bytes) : "memory");
    else if (vec == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(gmem), "r"(bytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48)
__device__ __forceinline__ uint64_t smem_desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ int lds_s32(uint32_t saddr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t saddr, int v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
}  // namespace fpm_ptx

// grid: (row blocks of FW_R rows, eval tiles of FW_CANDS candidates); block: FW_THREADS
// jobs: one FpmWarpJob per eval (eval e = 3 * candidate + angle), matrices already inverted (fpm_refine_prep_kernel)
__global__ void __launch_bounds__(FW_THREADS, 1)
fpm_corr_warp_kernel(const FpmWarpJob* __restrict__ jobs, int n_cands, FpmLevel src, int vec,
                     const __grid_constant__ CUtensorMap map_b, int rh, int tw, int th, int k_bytes, int e_pad,
                     int32_t* __restrict__ raw, int32_t* __restrict__ rowS, int32_t* __restrict__ rowQ, int* __restrict__ err_flag,
                     const int* __restrict__ n_cands_dev)
{
    using namespace fpm_ptx;
    if (n_cands_dev) {                                       // live candidates known only on the device
        n_cands = min(n_cands, *n_cands_dev);
        if ((int)blockIdx.y * FW_CANDS >= n_cands) return;
    }
    extern __shared__ uint8_t fw_smem_raw[];
    const uint32_t base_u32 = (smem_u32(fw_smem_raw) + 1023u) & ~1023u;
    uint8_t* base = fw_smem_raw + (base_u32 - smem_u32(fw_smem_raw));
    // layout: B ring (1024-aligned, SW128) | A stages | boxes | tables | barriers
    const uint32_t b_u32 = base_u32;
    const uint32_t a_u32 = b_u32 + FW_B_STAGES * MM_B_BYTES;
    const uint32_t box_u32 = a_u32 + FW_R * FW_A_STAGE;
    const uint32_t tab_off = FW_B_STAGES * MM_B_BYTES + FW_R * FW_A_STAGE + FW_NBOX * FW_BOX_BYTES;
    FwTables* tabs = reinterpret_cast<FwTables*>(base + tab_off);
    double* jm = reinterpret_cast<double*>(base + tab_off + FW_T * sizeof(FwTables));          // [eval slot][6]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + tab_off + FW_T * sizeof(FwTables) + FW_JOB_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    // barrier indices
    enum { B_FULL = 0, B_EMPTY = FW_B_STAGES, A_FULL = 2 * FW_B_STAGES, A_EMPTY, T_FULL, G_FULL, G_EMPTY = G_FULL + FW_T,
           X_FULL = G_EMPTY + FW_T, X_EMPTY = X_FULL + FW_NBOX, N_BARS = X_EMPTY + FW_NBOX };
    auto bar = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y0 = blockIdx.x * FW_R;
    const int nrows = min(FW_R, rh - y0);
    const int tile = blockIdx.y;
    const int cand0 = tile * FW_CANDS;
    const int ncand = min(FW_CANDS, n_cands - cand0);
    const int nk = (k_bytes + MM_KCHUNK - 1) / MM_KCHUNK;
    const int nvis = nk * ncand;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < FW_B_STAGES; s++) { mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_EMPTY + s), 1); }
        mbar_init(bar(A_FULL), FW_GATHER_WARPS);
        mbar_init(bar(A_EMPTY), 1 + 4);                       // MMA commit + 4 statistics warps
        mbar_init(bar(T_FULL), 1);
        for (int s = 0; s < FW_T; s++) { mbar_init(bar(G_FULL + s), 1); mbar_init(bar(G_EMPTY + s), FW_GATHER_WARPS + FW_COPY_WARPS + 1); }   // gather + copiers + prefetcher
        for (int s = 0; s < FW_NBOX; s++) { mbar_init(bar(X_FULL + s), 32 * FW_COPY_WARPS); mbar_init(bar(X_EMPTY + s), FW_GATHER_WARPS); }
        fence_barrier_init();
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < 3 * ncand * 6; i += FW_THREADS)
        jm[i] = jobs[(size_t)cand0 * 3 + i / 6].m[i % 6];
    __shared__ int src_img[FW_CANDS];
    if (threadIdx.x < ncand) src_img[threadIdx.x] = jobs[(size_t)(cand0 + threadIdx.x) * 3].src_img;
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer of B: (8 shifts x 8 template rows) x 128 B per (chunk, row) =====
        if (lane == 0) {
            int it = 0;
            for (int k = 0; k < nk; k++)
                for (int r = 0; r < nrows; r++, it++) {
                    const int s = it % FW_B_STAGES;
                    mbar_wait_sleep(bar(B_EMPTY + s), ((it / FW_B_STAGES) & 1) ^ 1, 400);
                    mbar_expect_tx(bar(B_FULL + s), MM_B_BYTES);
                    tma_load_3d(b_u32 + s * MM_B_BYTES, &map_b, bar(B_FULL + s), k * MM_KCHUNK, y0 + r - 7, 0);
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0;
            for (int k = 0; k < nk; k++) {
                mbar_wait_sleep(bar(A_FULL), k & 1, 400);      // all 128 rows of the 8 A stages of chunk k are written
                tc_fence_after();
                for (int r = 0; r < nrows; r++, it++) {
                    const int s = it % FW_B_STAGES;
                    mbar_wait(bar(B_FULL + s), (it / FW_B_STAGES) & 1);
                    tc_fence_after();
                    const uint64_t da = smem_desc_sw128(a_u32 + r * FW_A_STAGE);
                    const uint64_t db = smem_desc_sw128(b_u32 + s * MM_B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < MM_KCHUNK / 32; kk++)   // K = 32 bytes per instruction
                        mma_i8(tmem_base + r * MM_N, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), MM_IDESC,
                               (k | kk) ? 1u : 0u);
                    tc_commit(bar(B_EMPTY + s));
                }
                tc_commit(bar(A_EMPTY));                       // the A stages may be overwritten when these MMAs retire
            }
            tc_commit(bar(T_FULL));
        }
    } else if (warp < FW_W_COPY) {
        // ===== geometry: tables + source box of every visit (the warps alternate) =====
        for (int v = warp - FW_W_GEOM; v < nvis; v += FW_GEOM_WARPS) {
            const int k = v / ncand, c = v - k * ncand;
            const int ts = v % FW_T;
            mbar_wait_sleep(bar(G_EMPTY + ts), ((v / FW_T) & 1) ^ 1, 100);
            FwTables& T = tabs[ts];
            const int x0 = k * MM_KCHUNK;
            const int ncols = min(MM_KCHUNK, k_bytes - x0);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double m0 = jm[(c * 3 + j) * 6 + 0], m3 = jm[(c * 3 + j) * 6 + 3];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    // (M * x) * 1024 == M * (x * 1024) bit for bit: scaling by a power of two commutes with rounding
                    const double x1024 = (double)((x0 + lane + 32 * q) * 1024);
                    T.ad[j][lane + 32 * q] = fpm_cvround(m0 * x1024);
                    T.bd[j][lane + 32 * q] = fpm_cvround(m3 * x1024);
                }
            }
            int myX0 = 0, myY0 = 0;
            if (lane < 3 * FW_R) {
                const int j = lane >> 3, r = lane & 7;
                const double* mm = jm + (c * 3 + j) * 6;
                const double y = (double)(y0 + r);
                myX0 = fpm_cvround((mm[1] * y + mm[2]) * 1024.0) + 16;
                myY0 = fpm_cvround((mm[4] * y + mm[5]) * 1024.0) + 16;
            }
            __syncwarp();
            // exact source box: X and Y are sums of a function of x and a function of y, both monotone -> extremes at the
            // corners of the (real) tile; union over the three angles
            int Xmin = 0x7fffffff, Xmax = -0x7fffffff, Ymin = 0x7fffffff, Ymax = -0x7fffffff;
            if (lane < 3 * FW_R && (lane & 7) < nrows && ((lane & 7) == 0 || (lane & 7) == nrows - 1)) {
                const int j = lane >> 3;
                const int xa = T.ad[j][0], xb = T.ad[j][ncols - 1], ya = T.bd[j][0], yb = T.bd[j][ncols - 1];
                Xmin = min(myX0 + xa, myX0 + xb); Xmax = max(myX0 + xa, myX0 + xb);
                Ymin = min(myY0 + ya, myY0 + yb); Ymax = max(myY0 + ya, myY0 + yb);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Xmin = min(Xmin, __shfl_xor_sync(0xffffffffu, Xmin, o)); Xmax = max(Xmax, __shfl_xor_sync(0xffffffffu, Xmax, o));
                Ymin = min(Ymin, __shfl_xor_sync(0xffffffffu, Ymin, o)); Ymax = max(Ymax, __shfl_xor_sync(0xffffffffu, Ymax, o));
            }
            const int sx0 = Xmin >> 10, sx1 = (Xmax >> 10) + 1, sy0 = Ymin >> 10, sy1 = (Ymax >> 10) + 1;
            const int bx0 = sx0 & ~3;                                         // word aligned (two's complement: also for negative coordinates)
            const int by0 = sy0;
            const int nch = (sx1 - bx0) / 4 + 1;                              // words per box row
            // word pitch = +1 (mod 32 banks) when the lanes' line moves in the same direction in x and y, -1 when in opposite
            // directions (see fpm_warp_kernel): the 32 byte gathers of a warp then hit 32 different banks at any angle.  Rows
            // wider than that only occur for nearly horizontal lines, where any odd pitch will do.
            const bool same_dir = jm[(c * 3 + 1) * 6 + 0] * jm[(c * 3 + 1) * 6 + 3] >= 0;
            int pitchw = (same_dir && nch <= 33) ? 33 : ((!same_dir && nch <= 31) ? 31 : (nch | 1));
            if (pitchw * 4 * (sy1 - sy0 + 1) > FW_BOX_BYTES) pitchw = nch | 1;    // tall box of a nearly vertical line: x barely moves
            int brows = sy1 - sy0 + 1;
            if (pitchw * 4 * brows > FW_BOX_BYTES) {                          // cannot happen for a level the host admitted
                if (lane == 0 && err_flag) atomicExch(err_flag, 1);
                brows = FW_BOX_BYTES / (pitchw * 4);
            }
            if (lane < 3 * FW_R) {
                T.X0[lane >> 3][lane & 7] = myX0 - (bx0 << 10);
                T.Y0[lane >> 3][lane & 7] = myY0 - (by0 << 10);
            }
            if (lane == 0) { T.bx0 = bx0; T.by0 = by0; T.pitch = pitchw * 4; T.nrows = brows; T.nch = nch; }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(G_FULL + ts));
        }
    } else if (warp < FW_W_PF) {
        // ===== copiers: source box of every visit into box[v % FW_NBOX] =====
        // 4-byte cp.async (fire and forget: the whole box is in flight at once).  The box keeps an odd WORD pitch -- the only
        // row alignment for which byte gathers along steep lines hit 32 different banks -- which rules out wider copies.
        // A copier lane owns one word column and walks down the rows with constant strides: no division, no dependent
        // address chain (a single warp walking (row, word) pairs was latency bound at ~6000 clocks per box).
        const int sw = src.w, sh = src.h, sp = src.pitch;
        const int cl = (warp - FW_W_COPY) * 32 + lane;         // lane index among the copier lanes
        const int NL = 32 * FW_COPY_WARPS;
        int c = 0, ts = 0, xb = 0;
        uint32_t tph = 0, xph = 0;
        for (int v = 0; v < nvis; v++) {
            mbar_wait_sleep(bar(G_FULL + ts), tph, 60);
            mbar_wait_sleep(bar(X_EMPTY + xb), xph ^ 1, 60);
            const uint32_t t_u32 = smem_u32(&tabs[ts]);
            const int bx0 = lds_s32(t_u32 + offsetof(FwTables, bx0)), by0 = lds_s32(t_u32 + offsetof(FwTables, by0));
            const int pitch = lds_s32(t_u32 + offsetof(FwTables, pitch)), brows = lds_s32(t_u32 + offsetof(FwTables, nrows));
            const int nch = lds_s32(t_u32 + offsetof(FwTables, nch));
            // columns padded to a power of two >= 8: lane -> (row group, word column)
            const int lg = nch <= 8 ? 3 : (nch <= 16 ? 4 : (nch <= 32 ? 5 : 6));
            const int col = cl & ((1 << lg) - 1), rpp = NL >> lg;
            int row = cl >> lg;
            if (col < nch) {
                const int x = bx0 + 4 * col;
                const int nbx = x >= 0 ? max(0, min(4, sw - x)) : 0;
                const uint8_t* __restrict__ simg = src.ptr + (size_t)src_img[c] * src.img_stride;
                const uint32_t dbase = box_u32 + xb * FW_BOX_BYTES + 4 * col;
                // box rows [rlo, rhi) lie inside the image; the others (and columns outside it) are zeros
                const int rlo = min(brows, max(0, -by0)), rhi = max(rlo, min(brows, sh - by0));
                if (nbx == 0 || rlo > 0 || rhi < brows) {
                    for (int r = row; r < brows; r += rpp)
                        if (nbx == 0 || r < rlo || r >= rhi) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dbase + r * pitch), "r"(0) : "memory");
                    __threadfence_block();                     // plain stores before the (cp.async-tracked) arrival below
                }
                if (nbx > 0) {
                    while (row < rlo) row += rpp;
                    const uint8_t* g = simg + (ptrdiff_t)(by0 + row) * sp + x;
                    const ptrdiff_t gstep = (ptrdiff_t)rpp * sp;
                    uint32_t d = dbase + row * pitch;
                    const uint32_t dstep = rpp * pitch;
                    if (nbx == 4) {
#pragma unroll 4
                        for (; row < rhi; row += rpp, g += gstep, d += dstep)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(g) : "memory");
                    } else {
                        for (; row < rhi; row += rpp, g += gstep, d += dstep)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(g), "r"(nbx) : "memory");
                    }
                }
            }
            cp_async_mbar_arrive(bar(X_FULL + xb));
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(G_EMPTY + ts));
            if (++c == ncand) c = 0;
            if (++ts == FW_T) { ts = 0; tph ^= 1; }
            if (++xb == FW_NBOX) { xb = 0; xph ^= 1; }
        }
    } else if (warp == FW_W_PF) {
        // ===== L2 prefetcher: the source boxes of the visits FW_PF_AHEAD ahead of the geometry warp =====
        // (a box fill from HBM takes several visit times and only FW_NBOX boxes fit in shared memory: with the lines already
        //  in L2 the copiers see L2 latency only)
        const int sw = src.w, sh = src.h, sp = src.pitch;
        for (int v = 0; v < nvis; v++) {
            if (v >= FW_PF_AHEAD) {                            // throttle: stay FW_PF_AHEAD visits ahead of the geometry warp
                const int u = v - FW_PF_AHEAD;
                mbar_wait_sleep(bar(G_FULL + u % FW_T), (u / FW_T) & 1, 200);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(G_EMPTY + u % FW_T));
            }
            const int k = v / ncand, c = v - k * ncand;
            const int x0 = k * MM_KCHUNK;
            const int ncols = min(MM_KCHUNK, k_bytes - x0);
            int Xmin = 0x7fffffff, Xmax = -0x7fffffff, Ymin = 0x7fffffff, Ymax = -0x7fffffff;
            if (lane < 12) {
                const int j = lane >> 2;
                const double* mm = jm + (c * 3 + j) * 6;
                const double x = (double)(x0 + ((lane & 1) ? ncols - 1 : 0)), y = (double)(y0 + ((lane & 2) ? nrows - 1 : 0));
                Xmin = Xmax = fpm_cvround((mm[1] * y + mm[2]) * 1024.0) + 16 + fpm_cvround(mm[0] * x * 1024.0);
                Ymin = Ymax = fpm_cvround((mm[4] * y + mm[5]) * 1024.0) + 16 + fpm_cvround(mm[3] * x * 1024.0);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Xmin = min(Xmin, __shfl_xor_sync(0xffffffffu, Xmin, o)); Xmax = max(Xmax, __shfl_xor_sync(0xffffffffu, Xmax, o));
                Ymin = min(Ymin, __shfl_xor_sync(0xffffffffu, Ymin, o)); Ymax = max(Ymax, __shfl_xor_sync(0xffffffffu, Ymax, o));
            }
            const int px0 = max(Xmin >> 10, 0) & ~127, px1 = min((Xmax >> 10) + 1, sw - 1);
            const int py0 = max(Ymin >> 10, 0), py1 = min((Ymax >> 10) + 1, sh - 1);
            const uint8_t* simg = src.ptr + (size_t)src_img[c] * src.img_stride;
            for (int y = py0 + lane; y <= py1; y += 32)
                for (int x = px0; x <= px1; x += 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(simg + (size_t)y * sp + x));
        }
        for (int u = max(0, nvis - FW_PF_AHEAD); u < nvis; u++) {   // the table stages still owe this warp's arrival
            mbar_wait_sleep(bar(G_FULL + u % FW_T), (u / FW_T) & 1, 200);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(G_EMPTY + u % FW_T));
        }
    } else if (warp < FW_W_GATHER) {
        // ===== window statistics while the chunks stream through the A stages, then the epilogue: 4 warps =====
        // thread = eval row m of the A stages = TMEM lane m (a warp may only touch the lane quadrant warp % 4)
        const int q4 = warp & 3;
        const int m = q4 * 32 + lane;
        const bool live = m < 3 * ncand;
        const uint32_t rowoff = (uint32_t)m * MM_KCHUNK, swz = (uint32_t)(m & 7);
        uint32_t s0[FW_R], q0[FW_R];
        unsigned long long hd[FW_R], tl[FW_R];                 // 6 head bytes (x = 0..5) and 6 tail bytes (x = tw..tw+5), packed
#pragma unroll
        for (int r = 0; r < FW_R; r++) { s0[r] = 0; q0[r] = 0; hd[r] = 0; tl[r] = 0; }
        for (int k = 0; k < nk; k++) {
            mbar_wait_sleep(bar(A_FULL), k & 1, 1000);
            const int xb = k * MM_KCHUNK;
            const int nj = min(MM_KCHUNK / 16, (tw + FPM_ROI_PAD - xb + 15) >> 4);
#pragma unroll
            for (int r = 0; r < FW_R; r++) {
                if (r >= nrows) break;
                const uint32_t rp = a_u32 + r * FW_A_STAGE + rowoff;
                for (int j = 0; j < nj; j++) {
                    const uint4 v = lds128(rp + (((uint32_t)j ^ swz) << 4));
                    const int rem = tw - (xb + 16 * j);                             // window bytes in this chunk (may be <= 0)
                    if (k == 0 && j == 0)
                        hd[r] = (unsigned long long)v.x | ((unsigned long long)(v.y & 0xffffu) << 32);
                    if (rem >= 16) {
                        s0[r] = __dp4a(v.x, 0x01010101u, s0[r]); q0[r] = __dp4a(v.x, v.x, q0[r]);
                        s0[r] = __dp4a(v.y, 0x01010101u, s0[r]); q0[r] = __dp4a(v.y, v.y, q0[r]);
                        s0[r] = __dp4a(v.z, 0x01010101u, s0[r]); q0[r] = __dp4a(v.z, v.z, q0[r]);
                        s0[r] = __dp4a(v.w, 0x01010101u, s0[r]); q0[r] = __dp4a(v.w, v.w, q0[r]);
                    } else {
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            const int nb = rem - 4 * jj;
                            const uint32_t mk = nb >= 4 ? w[jj] : (nb <= 0 ? 0u : (w[jj] & (0xffffffffu >> (8 * (4 - nb)))));
                            s0[r] = __dp4a(mk, 0x01010101u, s0[r]);
                            q0[r] = __dp4a(mk, mk, q0[r]);
                        }
                        // tail byte i sits at position rem + i of this chunk when that is inside [0, 16)
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            const int pos = rem + i;
                            if (pos >= 0 && pos < 16) {
                                const uint32_t wsel = pos < 8 ? (pos < 4 ? v.x : v.y) : (pos < 12 ? v.z : v.w);
                                tl[r] |= (unsigned long long)((wsel >> (8 * (pos & 3))) & 0xffu) << (8 * i);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(A_EMPTY));
        }
        if (live) {
            const int e = (cand0 * 3) + m;
#pragma unroll
            for (int r = 0; r < FW_R; r++) {
                if (r >= nrows) break;
                int sc = (int)s0[r], qc = (int)q0[r];
                int32_t* ps = rowS + ((size_t)e * rh + (y0 + r)) * FPM_WSTRIDE;
                int32_t* pq = rowQ + ((size_t)e * rh + (y0 + r)) * FPM_WSTRIDE;
#pragma unroll
                for (int c = 0; c < FPM_NSHIFT; c++) {
                    if (c > 0) {
                        const int t = (int)((tl[r] >> (8 * (c - 1))) & 0xffu), hh = (int)((hd[r] >> (8 * (c - 1))) & 0xffu);
                        sc += t - hh; qc += t * t - hh * hh;
                    }
                    ps[c] = sc; pq[c] = qc;
                }
            }
        }
        // ---- epilogue: TMEM accumulators of the unit's rows -> raw[y][slot][64] ----
        mbar_wait_sleep(bar(T_FULL), 0, 1000);
        tc_fence_after();
        for (int r = 0; r < nrows; r++) {
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + r * MM_N;
            uint32_t vv[MM_N];
#pragma unroll
            for (int cc = 0; cc < MM_N / 16; cc++) tmem_ld16(taddr + cc * 16, vv + cc * 16);
            tmem_ld_wait();
            if (live) {
                uint4* o = reinterpret_cast<uint4*>(raw + ((size_t)(y0 + r) * e_pad + (size_t)tile * FW_M + m) * MM_N);
#pragma unroll
                for (int cc = 0; cc < MM_N / 4; cc++) o[cc] = make_uint4(vv[4 * cc], vv[4 * cc + 1], vv[4 * cc + 2], vv[4 * cc + 3]);
            }
        }
    } else {
        // ===== gather: the last FW_GATHER_WARPS warps =====
        const int g = warp - FW_W_GATHER;
        const int j = g >> 2;                                  // angle of this warp
        const int r0 = 2 * (g & 3);                            // its two rows of the unit
        // visit counters kept incrementally (a division per visit was a fifth of this loop)
        int k = 0, c = 0, ts = 0, xb = 0;
        uint32_t tph = 0, xph = 0;
        const uint32_t a_lane = a_u32 + (uint32_t)(lane & 15) + (uint32_t)r0 * FW_A_STAGE;
        for (int v = 0; v < nvis; v++) {
            if (c == 0) mbar_wait_sleep(bar(A_EMPTY), (k & 1) ^ 1, 100);  // the MMAs and statistics of the previous chunk are done with the A stages
            mbar_wait_sleep(bar(G_FULL + ts), tph, 40);
            mbar_wait_sleep(bar(X_FULL + xb), xph, 40);
            const uint32_t t_u32 = smem_u32(&tabs[ts]);
            const int pitch = lds_s32(t_u32 + offsetof(FwTables, pitch));
            const int ncols = min(MM_KCHUNK, k_bytes - k * MM_KCHUNK);
            int adj[4], bdj[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                adj[q] = lds_s32(t_u32 + offsetof(FwTables, ad) + (j * 128 + lane + 32 * q) * 4);
                bdj[q] = lds_s32(t_u32 + offsetof(FwTables, bd) + (j * 128 + lane + 32 * q) * 4);
            }
            int X0[2], Y0[2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                X0[rr] = lds_s32(t_u32 + offsetof(FwTables, X0) + (j * FW_R + r0 + rr) * 4);
                Y0[rr] = lds_s32(t_u32 + offsetof(FwTables, Y0) + (j * FW_R + r0 + rr) * 4);
            }
            const uint32_t sbox = box_u32 + xb * FW_BOX_BYTES;
            // row m = 3c + j of every A stage; 128-byte swizzle: 16-byte chunk ch of row m lives at chunk ch ^ (m & 7).
            // Pixel q of this lane is byte (lane & 15) of chunk 2q + (lane >> 4).
            const int mrow = 3 * c + j;
            const uint32_t a_row = a_lane + (uint32_t)mrow * MM_KCHUNK;
            uint32_t aoff[4];
#pragma unroll
            for (int q = 0; q < 4; q++) aoff[q] = a_row + ((((uint32_t)(2 * q + (lane >> 4))) ^ (uint32_t)(mrow & 7)) << 4);
            if (ncols == MM_KCHUNK && r0 + 1 < nrows) {
                // full chunk, both rows: 8 independent pixels per lane, no predicates (the loads of all of them are in flight
                // together; a branch per pixel had serialised them)
                int vv[2][4];
#pragma unroll
                for (int rr = 0; rr < 2; rr++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int XX = X0[rr] + adj[q], YY = Y0[rr] + bdj[q];
                        const int fx = XX & 0x3e0, fy = YY & 0x3e0;
                        const uint32_t a = sbox + (uint32_t)((YY >> 10) * pitch + (XX >> 10)), a2 = a + pitch;
                        const int p00 = lds_u8(a), p01 = lds_u8(a + 1), p10 = lds_u8(a2), p11 = lds_u8(a2 + 1);
                        const int top = (p00 << 10) + fx * (p01 - p00);
                        const int dif = (p10 << 10) - top + fx * (p11 - p10);
                        vv[rr][q] = ((top << 10) + (512 << 10) + fy * dif) >> 20;
                    }
#pragma unroll
                for (int rr = 0; rr < 2; rr++)
#pragma unroll
                    for (int q = 0; q < 4; q++) sts_u8(aoff[q] + rr * FW_A_STAGE, vv[rr][q]);
            } else {
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    if (r0 + rr < nrows) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            int val = 0;
                            if (lane + 32 * q < ncols) {
                                const int XX = X0[rr] + adj[q], YY = Y0[rr] + bdj[q];
                                const int fx = XX & 0x3e0, fy = YY & 0x3e0;
                                const uint32_t a = sbox + (uint32_t)((YY >> 10) * pitch + (XX >> 10)), a2 = a + pitch;
                                const int p00 = lds_u8(a), p01 = lds_u8(a + 1), p10 = lds_u8(a2), p11 = lds_u8(a2 + 1);
                                const int top = (p00 << 10) + fx * (p01 - p00);
                                const int dif = (p10 << 10) - top + fx * (p11 - p10);
                                val = ((top << 10) + (512 << 10) + fy * dif) >> 20;
                            }
                            sts_u8(aoff[q] + rr * FW_A_STAGE, val);
                        }
                    }
                }
            }
            const bool last = c == ncand - 1;
            if (last) fence_proxy_async();                     // this thread's A-stage writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(X_EMPTY + xb));
                mbar_arrive(bar(G_EMPTY + ts));
                if (last) mbar_arrive(bar(A_FULL));
            }
            if (++c == ncand) { c = 0; k++; }
            if (++ts == FW_T) { ts = 0; tph ^= 1; }
            if (++xb == FW_NBOX) { xb = 0; xph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}
