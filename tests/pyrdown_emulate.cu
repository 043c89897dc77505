// CPU emulation of fpm_pyrdown_kernel (fastest_image_pattern_matching_b200/csrc/fpm_pyrdown.cuh): the kernel's phase
// functions are __host__ __device__; here the threads of every CTA run one after the other between the barriers, so the
// tile / halo / reflection index arithmetic is checked against cv2.pyrDown in the GPU-less container
// (tests/test_pyrdown_emulation.py).  Test infrastructure only.
#include "../fastest_image_pattern_matching_b200/csrc/fpm_pyrdown.cuh"
#include <vector>

template <bool TWO> static void run(const Pd2Args& a, int batch)
{
    typedef Pd2Cfg<TWO> C;
    std::vector<uint8_t> smem_raw(C::SMEM + 64);
    uint8_t* smem = smem_raw.data();
    smem += (16 - (reinterpret_cast<uintptr_t>(smem) & 15)) & 15;
    uint8_t* s_in = smem;
    uint8_t* s_l1 = smem + C::IH * C::IP;
    const int gx = (a.d1.w + PD2_TW - 1) / PD2_TW, gy = (a.d1.h + PD2_TH - 1) / PD2_TH;
    for (int bz = 0; bz < batch; bz++)
        for (int by = 0; by < gy; by++)
            for (int bx = 0; bx < gx; bx++) {
                for (size_t i = 0; i < (size_t)C::SMEM; i++) smem[i] = (uint8_t)(0xA5 + 7 * i);   // "uninitialised" shared memory
                const Pd2Tile<TWO> tile(bx, by, a);
                for (int t = 0; t < C::NT; t++) {
                    int p_first, np;
                    if (a.vec >= 16) {
                        if (t == 0) {                       // the TMA tile load: zero outside the [h][pitch] tensor, padding as it is
                            for (int r = 0; r < C::IH; r++)
                                for (int b = 0; b < C::IP; b++) {
                                    const int x = tile.xs + b, y = tile.ys + r;
                                    s_in[r * C::IP + b] = (x >= 0 && x < a.src.pitch && y >= 0 && y < a.src.h)
                                                              ? a.src.ptr[(size_t)bz * a.src.img_stride + (size_t)y * a.src.pitch + x] : 0;
                                }
                        }
                    } else {
                        if (a.vec >= 8) pd2_stage_pieces<TWO, 8>(t, tile, bz, a, s_in, &p_first, &np);
                        else pd2_stage_pieces<TWO, 4>(t, tile, bz, a, s_in, &p_first, &np);
                        pd2_stage_bytes<TWO>(t, tile, bz, a, s_in, a.vec >= 8 ? 8 : 4, p_first, np);
                    }
                }
                if (a.vec >= 16 && pd2_stage_needs_fix<TWO>(tile, a))
                    for (int t = 0; t < C::NT; t++) pd2_stage_fix<TWO>(t, tile, bz, a, s_in);
                for (int t = 0; t < C::NT; t++) pd2_level1<TWO>(t, tile, bz, a, s_in, s_l1);
                if (TWO) {
                    if (pd2_tile_on_border(bx, by, a)) {
                        for (int t = 0; t < C::NT; t++) pd2_fix<0>(t, bx, by, a, s_l1);
                        for (int t = 0; t < C::NT; t++) pd2_fix<1>(t, bx, by, a, s_l1);
                    }
                    for (int t = 0; t < C::NT; t++) pd2_level2(t, bx, by, bz, a, s_l1);
                }
            }
}

// src: [batch][h][pitch] with the given base alignment; d1/d2 pitches given; vec = 16 / 8 / 4 / 1
extern "C" int pd_emulate(const uint8_t* src, int w, int h, int pitch, long long img_stride, int batch, int vec, uint8_t* d1, int p1,
                          long long s1, uint8_t* d2, int p2, long long s2)
{
    Pd2Args a;
    const int w1 = (w + 1) / 2, h1 = (h + 1) / 2;
    a.src = FpmLevel{const_cast<uint8_t*>(src), w, h, pitch, (size_t)img_stride};
    a.d1 = FpmLevel{d1, w1, h1, p1, (size_t)s1};
    a.d2 = FpmLevel{d2, (w1 + 1) / 2, (h1 + 1) / 2, p2, (size_t)s2};
    a.vec = vec;
    a.st1_vec = (reinterpret_cast<uintptr_t>(d1) % 8 == 0 && p1 % 8 == 0 && s1 % 8 == 0) ? 1 : 0;
    a.st2_vec = (d2 && reinterpret_cast<uintptr_t>(d2) % 8 == 0 && p2 % 8 == 0 && s2 % 8 == 0) ? 1 : 0;
    if (d2) run<true>(a, batch); else run<false>(a, batch);
    return 0;
}
