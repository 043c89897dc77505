// fpm_pyrdown.cuh -- cv::buildPyramid / cv::pyrDown (src/TemplateMatcher.cpp:55, :124) for sm_100a, round 3 of the kernel.
//
//   5x5 [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, (s + 128) >> 8, out ((w+1)/2, (h+1)/2)           -- bit-exact vs cv2.
//
// fpm_pyrdown_kernel<TWO> produces ONE level (TWO = false) or TWO consecutive levels (TWO = true) per launch.  The
// chain is HBM-bound (1.667 B per source pixel); the old one-level kernel spent 43 thread-instructions per output pixel
// and was issue-bound at 0.6 of the HBM peak, and it re-read level 1 from DRAM to make level 2.  Here:
//
//   * one CTA owns a 128x64 tile of level 1 (and the 64x32 tile of level 2 under it).  Its source rows are staged in
//     shared memory by cp.async (16 / 8 / 4 bytes, whatever the caller's pitch allows); nothing else touches DRAM.
//   * a thread computes an 8-column x 4-row block of level 1 and keeps the last five horizontally filtered source rows
//     in registers (a sliding window), so there is no intermediate buffer and only one barrier per level:
//       - horizontal taps: the 16 centre bytes of a row come with ONE 128-bit shared load, the 2 + 1 halo bytes from the
//         neighbour lanes by shuffle (only strip-edge lanes touch shared memory again); an even output is
//         dp4a(prev,(0,0,1,4)) + dp4a(cur,(6,4,1,0)), an odd one dp4a(cur,(1,4,6,4)) + dp4a(next,(1,0,0,0)): no shifts;
//       - two outputs are packed in one register as 16-bit lanes (sums <= 16*4080 + 128 < 65536, no carry between
//         lanes), the vertical taps are 4 integer ops per pair and the rounded result is byte 1 of each lane, so one
//         PRMT packs four output pixels.  ~11 thread-instructions per level-1 pixel.
//   * TWO: the level-1 block is also written to a 68x144 shared tile (the owned 64x128 + the 2/1 halo rows and columns
//     that level 2 needs, computed redundantly from a slightly larger source tile); halo entries outside the level-1
//     image are replaced by their REFLECT_101 partners, then 128 threads run the same routine on the tile and store
//     level 2.  Level 1 is never read back from DRAM and a 6-level chain is 3 launches.
//
// The phase functions are __host__ __device__ so that tests/pyrdown_emulate.cu can run the exact index arithmetic on
// the CPU (threads one after the other between barriers) against cv2 -- there is no GPU in the build container.
#pragma once
#include <cuda.h>
#include "fpm_common.cuh"
#include <string.h>

#define PD2_R 4                    // level-1 rows per thread
#define PD2_R2 2                   // level-2 rows per thread
#define PD2_TW 128                 // owned level-1 tile
#define PD2_TH 64
#define PD2_G2 8                   // level-2 column groups (64 columns)
#define PD2_S2 (PD2_TH / 2 / PD2_R2)

template <bool TWO> struct Pd2Cfg {
    static constexpr int GOFF = TWO ? 8 : 0;                  // level-1 columns computed left of the owned tile
    static constexpr int ROFF = TWO ? 2 : 0;                  // level-1 rows computed above the owned tile
    static constexpr int NG = TWO ? 18 : 16;                  // 8-column groups
    static constexpr int NS = TWO ? 17 : 16;                  // 4-row strips
    static constexpr int NT = TWO ? 320 : 256;                // threads per CTA
    static constexpr int NCH = NG + 2;                        // 16-byte chunks per staged source row
    static constexpr int IP = 16 * NCH;                       // staged row pitch (bytes)
    static constexpr int IH = 2 * PD2_R * NS + 3;             // staged rows
    static constexpr int L1P = 160;                           // level-1 tile pitch: 8 (shift) + 144 + pad, 16-byte rows
    static constexpr int L1H = PD2_R * NS;
    static constexpr int SMEM = IH * IP + (TWO ? L1H * L1P : 0) + 128;   // + slack to align the TMA destination
};

struct Pd2Args {
    FpmLevel src, d1, d2;
    int vec;                       // 16 / 8 / 4: alignment of src.ptr, src.pitch and src.img_stride; 1 = none
    int st1_vec, st2_vec;          // d1 / d2 rows can take 8-byte stores
};

struct PdRow { uint32_t a, b, c, d; };          // 8 horizontal sums as 4 x (u16, u16)

FPM_HD uint32_t pd_dp4a(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef __CUDA_ARCH__
    return __dp4a(a, b, c);
#else
    for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u);
    return c;
#endif
}

FPM_HD uint32_t pd_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const unsigned long long v = (unsigned long long)a | ((unsigned long long)b << 32);
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 255u) << (8 * k);
    return r;
#endif
}

// horizontal taps of 8 outputs whose centres are bytes 0, 2, .. 14 of w; wm = the word before, wp = the word after
FPM_HD PdRow pd_hsum(uint32_t wm, uint4 w, uint32_t wp)
{
    const uint32_t CA = 0x04010000u, CB = 0x00010406u, CC = 0x04060401u, CD = 0x00000001u;
    const uint32_t e0 = pd_dp4a(wm, CA, pd_dp4a(w.x, CB, 0)), o0 = pd_dp4a(w.x, CC, pd_dp4a(w.y, CD, 0));
    const uint32_t e1 = pd_dp4a(w.x, CA, pd_dp4a(w.y, CB, 0)), o1 = pd_dp4a(w.y, CC, pd_dp4a(w.z, CD, 0));
    const uint32_t e2 = pd_dp4a(w.y, CA, pd_dp4a(w.z, CB, 0)), o2 = pd_dp4a(w.z, CC, pd_dp4a(w.w, CD, 0));
    const uint32_t e3 = pd_dp4a(w.z, CA, pd_dp4a(w.w, CB, 0)), o3 = pd_dp4a(w.w, CC, pd_dp4a(wp, CD, 0));
    PdRow r;
    r.a = pd_prmt(e0, o0, 0x5410); r.b = pd_prmt(e1, o1, 0x5410);
    r.c = pd_prmt(e2, o2, 0x5410); r.d = pd_prmt(e3, o3, 0x5410);
    return r;
}

// one staged row: p = the 16 centre bytes (16-byte aligned).  The halo words come from the neighbour lanes by shuffle.  A
// lane without a neighbour that holds the adjacent chunk (edge: 1 = left, 2 = right; first / last group of a strip, lane
// 0 / 31) reads its halo word itself -- ONE predicated load per row for the left and the right edge lanes together (two
// loads cost 4.4 wavefronts per warp and row: the few active lanes sit in strips whose rows share their banks) and two
// selects, no branch.  No lane has both edges (pd_no_double_edge).  The host build reads both words directly.
FPM_HD PdRow pd_hrow(const uint8_t* p, int edge)
{
    const uint4 w = *reinterpret_cast<const uint4*>(p);
#ifdef __CUDA_ARCH__
    uint32_t wm = __shfl_up_sync(0xffffffffu, w.w, 1), wp = __shfl_down_sync(0xffffffffu, w.x, 1);
    uint32_t v = 0;
    const unsigned ea = (unsigned)__cvta_generic_to_shared(p) + (edge == 1 ? -4 : 16);
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.s32 pe, %2, 0;\n\t@pe ld.shared.u32 %0, [%1];\n\t}" : "+r"(v) : "r"(ea), "r"(edge));
    wm = edge == 1 ? v : wm;
    wp = edge == 2 ? v : wp;
#else
    (void)edge;
    const uint32_t wm = *reinterpret_cast<const uint32_t*>(p - 4), wp = *reinterpret_cast<const uint32_t*>(p + 16);
#endif
    return pd_hsum(wm, w, wp);
}

// with ng groups per strip laid out over the lanes of nt / 32 warps, is there a lane that is both a left and a right edge?
constexpr bool pd_no_double_edge(int ng, int nthreads, int nlive)
{
    for (int t = 0; t < nthreads; t++) {
        const int id = t < nlive ? t : nlive - 1, g = id % ng, lane = t & 31;
        if ((g == 0 || lane == 0) && (g == ng - 1 || lane == 31)) return false;
    }
    return true;
}
static_assert(pd_no_double_edge(18, 320, 18 * 17) && pd_no_double_edge(16, 256, 256) && pd_no_double_edge(PD2_G2, PD2_G2 * PD2_S2, PD2_G2 * PD2_S2),
              "a lane with both halo words missing needs a second load in pd_hrow");

FPM_HD uint32_t pd_vpair(uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t h4)
{
    return h0 + h4 + 0x00800080u + 4u * (h1 + h3) + 6u * h2;
}

// vertical taps of five filtered rows -> 8 output bytes
FPM_HD uint2 pd_vert(const PdRow& h0, const PdRow& h1, const PdRow& h2, const PdRow& h3, const PdRow& h4)
{
    const uint32_t va = pd_vpair(h0.a, h1.a, h2.a, h3.a, h4.a), vb = pd_vpair(h0.b, h1.b, h2.b, h3.b, h4.b);
    const uint32_t vc = pd_vpair(h0.c, h1.c, h2.c, h3.c, h4.c), vd = pd_vpair(h0.d, h1.d, h2.d, h3.d, h4.d);
    return make_uint2(pd_prmt(va, vb, 0x7531), pd_prmt(vc, vd, 0x7531));       // (sum >> 8) of a lane = its byte 1
}

// 8 output bytes to a level row: one 8-byte store when the row allows it, else byte by byte up to the image edge
FPM_HD void pd_store8(uint8_t* row, int c, int w, bool vec8, uint2 o)
{
    if (vec8 && c + 8 <= w) {
        *reinterpret_cast<uint2*>(row + c) = o;
    } else {
        const unsigned long long pk = (unsigned long long)o.x | ((unsigned long long)o.y << 32);
        for (int q = 0; q < 8 && c + q < w; q++) row[c + q] = (uint8_t)(pk >> (8 * q));
    }
}

template <int PB> FPM_HD void pd_copy(uint8_t* smem, const uint8_t* g)
{
#ifdef __CUDA_ARCH__
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if (PB == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
#else
    memcpy(smem, g, PB);
#endif
}

// REFLECT_101 of an index in [-2, n+1] (all that is ever needed), n >= 1
FPM_HD int pd_reflect(int i, int n)
{
    if (n == 1) return 0;
    int c = i < 0 ? -i : i;
    if (c >= n) c = 2 * n - 2 - c;
    return c < 0 ? -c : c;
}

// tile geometry shared by the phases
template <bool TWO> struct Pd2Tile {
    int X1, Y1;                    // owned level-1 tile origin
    int xs, ys;                    // source column / row of staged byte 0 / row 0
    int ng, ns;                    // groups / strips that hold any level-1 pixel somebody needs
    int nrows, nch;                // staged rows / 16-byte chunks per row
    bool rows_inside;              // no staged row needs the reflection
    FPM_HD Pd2Tile(int bx, int by, const Pd2Args& a)
    {
        typedef Pd2Cfg<TWO> C;
        X1 = bx * PD2_TW; Y1 = by * PD2_TH;
        xs = 2 * (X1 - C::GOFF) - 16; ys = 2 * (Y1 - C::ROFF) - 2;
        const int halo = TWO ? 2 : 0;                                       // level-1 pixels past the image that level 2 reads
        const int cols = min(8 * C::NG, a.d1.w + halo - (X1 - C::GOFF));    // needed level-1 columns of the computed area
        const int rows = min(PD2_R * C::NS, a.d1.h + halo - (Y1 - C::ROFF));
        ng = (cols + 7) / 8; ns = (rows + PD2_R - 1) / PD2_R;
        nrows = 2 * PD2_R * ns + 3; nch = ng + 2;
        rows_inside = ys >= 0 && ys + nrows <= a.src.h;
    }
    // source row behind staged row r (rows that no level-1 pixel of the image depends on map to anything valid)
    FPM_HD int src_row(int r, int sh) const { return rows_inside ? ys + r : pd_reflect(max(-2, min(sh + 1, ys + r)), sh); }
};

// ---- phase A: stage the source rows of the tile ------------------------------------------------------------------
// Source pixels that no level-1 pixel of the IMAGE depends on (left of column -2, right of column w+1, likewise rows)
// are not staged at all: they only feed level-1 pixels outside the level-1 image, which phase C overwrites.
//
// 16-byte aligned source: the whole staged tile (IH rows x IP bytes, origin (xs, ys)) is ONE TMA tile load of a
// [batch][h][pitch / 4] u32 tensor (fpm_pyrdown_kernel); rows and columns outside the tensor arrive as zero, the row padding
// between w and pitch as whatever it holds.  This pass then writes the needed bytes that are not image pixels:
// columns -2, -1, w, w+1 of every staged row and the whole rows -2, -1, h, h+1, all through REFLECT_101.  It runs after the
// tile has landed (it overwrites bytes the TMA wrote).
template <bool TWO>
FPM_HD void pd2_stage_fix(int tid, const Pd2Tile<TWO>& t, int bz, const Pd2Args& a, uint8_t* s_in)
{
    typedef Pd2Cfg<TWO> C;
    const int sw = a.src.w, sh = a.src.h;
    const uint8_t* s = a.src.ptr + (size_t)bz * a.src.img_stride;
    const int lo = max(0, -2 - t.xs), hi = min(16 * t.nch, sw + 2 - t.xs);     // needed bytes of a staged row: [lo, hi)
    if (hi <= lo) return;
    const int l1 = min(hi, max(lo, -t.xs)), r0b = max(l1, min(hi, sw - t.xs)); // [lo, l1): columns < 0, [r0b, hi): columns >= w
    const int nl = l1 - lo, nb = nl + (hi - r0b);
    const int ncol = t.nrows * nb;                                             // part 1: the edge columns of every row
    // part 2: staged rows above row 0 and below row h-1 that are needed (source rows -2, -1, h, h+1)
    const int ra0 = max(0, -2 - t.ys), ra1 = min(t.nrows, max(ra0, -t.ys));    // rows above: [ra0, ra1)
    const int rb0 = max(ra1, min(t.nrows, sh - t.ys)), rb1 = min(t.nrows, max(rb0, sh + 2 - t.ys));
    const int na = ra1 - ra0, nrow = (na + (rb1 - rb0)) * (hi - lo);
    for (int i = tid; i < ncol + nrow; i += C::NT) {
        int r, b;
        if (i < ncol) {
            r = i / nb;
            const int k = i - r * nb;
            b = k < nl ? lo + k : r0b + (k - nl);
        } else {
            const int q = (i - ncol) / (hi - lo);
            b = lo + (i - ncol) - q * (hi - lo);
            r = q < na ? ra0 + q : rb0 + (q - na);
        }
        const int sy = pd_reflect(max(-2, min(sh + 1, t.ys + r)), sh);
        s_in[(size_t)r * C::IP + b] = s[(size_t)sy * a.src.pitch + pd_reflect(t.xs + b, sw)];
    }
}

// does the TMA-staged tile need pd2_stage_fix at all?
template <bool TWO>
FPM_HD bool pd2_stage_needs_fix(const Pd2Tile<TWO>& t, const Pd2Args& a)
{
    return !t.rows_inside || t.xs < 0 || t.xs + 16 * t.nch > a.src.w;
}

// pass A, 8- or 4-byte aligned source: one cp.async per piece; a thread keeps its column
template <bool TWO, int PB>
FPM_HD void pd2_stage_pieces(int tid, const Pd2Tile<TWO>& t, int bz, const Pd2Args& a, uint8_t* s_in, int* p_first, int* np)
{
    typedef Pd2Cfg<TWO> C;
    const int sw = a.src.w;
    const uint8_t* s = a.src.ptr + (size_t)bz * a.src.img_stride;
    const int ppr = t.nch * (16 / PB);                      // pieces per row
    *p_first = t.xs < 0 ? (-t.xs + PB - 1) / PB : 0;
    *np = (a.vec >= PB ? min(ppr - 1, (sw - t.xs) / PB - 1) : -1) - *p_first + 1;
    if (*np <= 0) return;
    const int rpp = C::NT / *np;                            // rows per pass
    if (tid >= rpp * *np) return;
    const int pc = *p_first + tid % *np, r0 = tid / *np;
    uint8_t* dst = s_in + (size_t)r0 * C::IP + PB * pc;
    for (int r = r0; r < t.nrows; r += rpp, dst += rpp * C::IP)
        pd_copy<PB>(dst, s + (size_t)t.src_row(r, a.src.h) * a.src.pitch + (t.xs + PB * pc));
}

// pass B: the needed bytes outside the pieces of pass A (columns -2, -1 on the left edge, the ragged tail and columns w, w+1
// on the right edge; everything when the source is not aligned), one byte per thread and turn
template <bool TWO>
FPM_HD void pd2_stage_bytes(int tid, const Pd2Tile<TWO>& t, int bz, const Pd2Args& a, uint8_t* s_in, int PB, int p_first, int np)
{
    typedef Pd2Cfg<TWO> C;
    const int sw = a.src.w, sh = a.src.h;
    const uint8_t* s = a.src.ptr + (size_t)bz * a.src.img_stride;
    const int lo = max(0, -2 - t.xs), hi = min(16 * t.nch, sw + 2 - t.xs);     // needed bytes of a staged row: [lo, hi)
    int l1, r0b;                                                             // [lo, l1) and [r0b, hi) are not covered by pass A
    if (np > 0) { l1 = max(lo, min(PB * p_first, hi)); r0b = max(l1, min(PB * (p_first + np), hi)); }
    else { l1 = hi; r0b = hi; }
    const int nl = l1 - lo, nb = nl + (hi - r0b);
    if (nb <= 0) return;
    for (int i = tid; i < t.nrows * nb; i += C::NT) {
        const int r = i / nb, k = i - r * nb;
        const int b = k < nl ? lo + k : r0b + (k - nl);
        s_in[(size_t)r * C::IP + b] = s[(size_t)t.src_row(r, sh) * a.src.pitch + pd_reflect(t.xs + b, sw)];
    }
}

// ---- phase B: level 1 = 8x4 block per thread --------------------------------------------------------------------
template <bool TWO>
FPM_HD void pd2_level1(int tid, const Pd2Tile<TWO>& t, int bz, const Pd2Args& a, const uint8_t* s_in, uint8_t* s_l1)
{
    typedef Pd2Cfg<TWO> C;
    const bool live = tid < C::NG * C::NS;
    const int id = live ? tid : C::NG * C::NS - 1;          // idle lanes of the last warp mirror a live thread (shuffles)
    const int g = id % C::NG, st = id / C::NG;
    const bool wanted = live && g < t.ng && st < t.ns;
#ifdef __CUDA_ARCH__
    if (!__any_sync(0xffffffffu, wanted)) return;           // the whole warp lies outside the needed area
    const int lane = tid & 31;
    const int edge = ((g == 0 || lane == 0) ? 1 : 0) | ((g == C::NG - 1 || lane == 31) ? 2 : 0);
#else
    if (!wanted) return;
    const int edge = 3;
#endif
    const int c1 = t.X1 - C::GOFF + 8 * g, r1 = t.Y1 - C::ROFF + PD2_R * st;
    const uint8_t* p = s_in + (size_t)(2 * PD2_R * st) * C::IP + 16 * g + 16;
    // rows [jlo, jhi) of the block are stored to the level (the rest is halo for level 2 or lies outside the image)
    const bool own_cols = wanted && c1 >= t.X1 && c1 < t.X1 + PD2_TW && c1 < a.d1.w;
    const int jlo = own_cols ? max(0, t.Y1 - r1) : PD2_R, jhi = min(PD2_R, min(t.Y1 + PD2_TH, a.d1.h) - r1);
    const bool full8 = a.st1_vec != 0 && c1 + 8 <= a.d1.w;
    uint8_t* drow = a.d1.ptr + (size_t)bz * a.d1.img_stride + (long long)r1 * a.d1.pitch;
    uint8_t* trow = s_l1 + (PD2_R * st) * C::L1P + 8 * g + 8;
    PdRow h0 = pd_hrow(p, edge), h1 = pd_hrow(p + C::IP, edge), h2 = pd_hrow(p + 2 * C::IP, edge);
    uint2 o[PD2_R];
#pragma unroll
    for (int j = 0; j < PD2_R; j++) {
        const PdRow h3 = pd_hrow(p + (2 * j + 3) * C::IP, edge), h4 = pd_hrow(p + (2 * j + 4) * C::IP, edge);
        o[j] = pd_vert(h0, h1, h2, h3, h4);
        h0 = h2; h1 = h3; h2 = h4;
    }
    if (TWO && wanted) {
#pragma unroll
        for (int j = 0; j < PD2_R; j++) *reinterpret_cast<uint2*>(trow + j * C::L1P) = o[j];
    }
    if (full8 && jlo == 0 && jhi == PD2_R) {                // the common case: four 8-byte stores, no per-row tests
#pragma unroll
        for (int j = 0; j < PD2_R; j++) *reinterpret_cast<uint2*>(drow + (size_t)j * a.d1.pitch + c1) = o[j];
    } else {
#pragma unroll
        for (int j = 0; j < PD2_R; j++)
            if (j >= jlo && j < jhi) pd_store8(drow + (size_t)j * a.d1.pitch, c1, a.d1.w, full8, o[j]);
    }
}

// ---- phase C (TWO): level-1 tile entries outside the level-1 image <- their REFLECT_101 partners ---------------------
// tile byte of level-1 pixel (r, c) = (r - (Y1-2)) * L1P + (c - (X1-8)) + 8
template <int PART>
FPM_HD void pd2_fix(int tid, int bx, int by, const Pd2Args& a, uint8_t* s_l1)
{
    typedef Pd2Cfg<true> C;
    const int X0 = bx * PD2_TW - C::GOFF, Y0 = by * PD2_TH - C::ROFF;
    const int w1 = a.d1.w, h1 = a.d1.h;
    if (PART == 0) {                                        // columns -2, -1, w1, w1+1: every tile row
        if (tid >= 4 * C::L1H) return;
        const int k = tid / C::L1H, rr = tid % C::L1H;
        const int c = k < 2 ? k - 2 : w1 + (k - 2);
        const int cs = pd_reflect(c, w1);
        if (c - X0 < 0 || c - X0 >= 8 * C::NG || cs - X0 < 0 || cs - X0 >= 8 * C::NG) return;
        s_l1[rr * C::L1P + (c - X0) + 8] = s_l1[rr * C::L1P + (cs - X0) + 8];
    } else {                                                // rows -2, -1, h1, h1+1: whole rows, word by word
        const int nw = 8 * C::NG / 4;
        if (tid >= 4 * nw) return;
        const int k = tid / nw, wd = tid % nw;
        const int r = k < 2 ? k - 2 : h1 + (k - 2);
        const int rs = pd_reflect(r, h1);
        if (r - Y0 < 0 || r - Y0 >= C::L1H || rs - Y0 < 0 || rs - Y0 >= C::L1H) return;
        *reinterpret_cast<uint32_t*>(s_l1 + (r - Y0) * C::L1P + 8 + 4 * wd) = *reinterpret_cast<const uint32_t*>(s_l1 + (rs - Y0) * C::L1P + 8 + 4 * wd);
    }
}

FPM_HD bool pd2_tile_on_border(int bx, int by, const Pd2Args& a)
{
    return bx == 0 || by == 0 || bx * PD2_TW + PD2_TW + 1 > a.d1.w || by * PD2_TH + PD2_TH + 1 > a.d1.h;
}

// ---- phase D (TWO): level 2 = 8x2 block per thread, 128 threads --------------------------------------------------
FPM_HD void pd2_level2(int tid, int bx, int by, int bz, const Pd2Args& a, const uint8_t* s_l1)
{
    typedef Pd2Cfg<true> C;
    if (tid >= PD2_G2 * PD2_S2) return;
    const int g = tid % PD2_G2, st = tid / PD2_G2;
    const int c2 = bx * (PD2_TW / 2) + 8 * g, r2 = by * (PD2_TH / 2) + PD2_R2 * st;
    const bool wanted = c2 < a.d2.w && r2 < a.d2.h;
#ifdef __CUDA_ARCH__
    if (!__any_sync(0xffffffffu, wanted)) return;
    const int lane = tid & 31;
    const int edge = ((g == 0 || lane == 0) ? 1 : 0) | ((g == PD2_G2 - 1 || lane == 31) ? 2 : 0);
#else
    if (!wanted) return;
    const int edge = 3;
#endif
    const uint8_t* p = s_l1 + (2 * PD2_R2 * st) * C::L1P + 16 * g + 16;
    const int jhi = wanted ? min(PD2_R2, a.d2.h - r2) : 0;
    const bool full8 = a.st2_vec != 0 && c2 + 8 <= a.d2.w;
    uint8_t* drow = a.d2.ptr + (size_t)bz * a.d2.img_stride + (size_t)r2 * a.d2.pitch;
    PdRow h0 = pd_hrow(p, edge), h1 = pd_hrow(p + C::L1P, edge), h2 = pd_hrow(p + 2 * C::L1P, edge);
    uint2 o[PD2_R2];
#pragma unroll
    for (int j = 0; j < PD2_R2; j++) {
        const PdRow h3 = pd_hrow(p + (2 * j + 3) * C::L1P, edge), h4 = pd_hrow(p + (2 * j + 4) * C::L1P, edge);
        o[j] = pd_vert(h0, h1, h2, h3, h4);
        h0 = h2; h1 = h3; h2 = h4;
    }
    if (full8 && jhi == PD2_R2) {
#pragma unroll
        for (int j = 0; j < PD2_R2; j++) *reinterpret_cast<uint2*>(drow + (size_t)j * a.d2.pitch + c2) = o[j];
    } else {
#pragma unroll
        for (int j = 0; j < PD2_R2; j++)
            if (j < jhi) pd_store8(drow + (size_t)j * a.d2.pitch, c2, a.d2.w, full8, o[j]);
    }
}

#ifdef __CUDACC__
// tmap: [batch][src.h][src.pitch / 4] u32 tensor of the source level, box (IP / 4, IH, 1), no swizzle, zero fill; only read
// when a.vec >= 16 (TMA needs a 16-byte aligned base and strides)
template <bool TWO>
__global__ void __launch_bounds__(Pd2Cfg<TWO>::NT, 4)
fpm_pyrdown_kernel(Pd2Args a, const __grid_constant__ CUtensorMap tmap)
{
    typedef Pd2Cfg<TWO> C;
    extern __shared__ uint8_t pd_smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    uint8_t* s_in = pd_smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(pd_smem_raw) & 127u)) & 127u);   // TMA destination: 128-byte aligned
    uint8_t* s_l1 = s_in + C::IH * C::IP;
    const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    const Pd2Tile<TWO> t(bx, by, a);
    if (a.vec >= 16) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(C::IH * C::IP) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(s_in)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bar),
                           "r"(t.xs / 4), "r"(t.ys), "r"(bz) : "memory");
        }
        if (tid < 32) {                                     // one warp watches the barrier, the others sleep in bar.sync
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                             : "=r"(done) : "r"(bar) : "memory");
        }
        __syncthreads();
        if (pd2_stage_needs_fix<TWO>(t, a)) {
            pd2_stage_fix<TWO>(tid, t, bz, a, s_in);
            __syncthreads();
        }
    } else {
        int p_first, np;
        if (a.vec >= 8) pd2_stage_pieces<TWO, 8>(tid, t, bz, a, s_in, &p_first, &np);
        else pd2_stage_pieces<TWO, 4>(tid, t, bz, a, s_in, &p_first, &np);
        pd2_stage_bytes<TWO>(tid, t, bz, a, s_in, a.vec >= 8 ? 8 : 4, p_first, np);
        fpm_cp_async_commit();
        fpm_cp_async_wait<0>();
        __syncthreads();
    }
    pd2_level1<TWO>(tid, t, bz, a, s_in, s_l1);
    if (TWO) {
        __syncthreads();
        if (pd2_tile_on_border(bx, by, a)) {
            pd2_fix<0>(tid, bx, by, a, s_l1);
            __syncthreads();
            pd2_fix<1>(tid, bx, by, a, s_l1);
            __syncthreads();
        }
        pd2_level2(tid, bx, by, bz, a, s_l1);
    }
}
#endif
