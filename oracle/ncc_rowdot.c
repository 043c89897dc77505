/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * SSE2 restatement of the reference's only hand-vectorised routine and of the loop that
 * drives it:
 *   IM_Conv_SIMD            /root/reference/src/TemplateMatcher.cpp:461-483
 *   _mm_hsum_epi32          /root/reference/src/TemplateMatcher.cpp:21-26
 *   MatchTemplate SIMD loop /root/reference/src/TemplateMatcher.cpp:490-510
 *
 * Semantics: for every result cell (r, c) a float32 accumulator receives, in template-row
 * order, the exact int32 dot product of template row t_r with source row r + t_r at column c.
 * Pinned against the reference's own source by oracle/build_ref.sh + tests/test_oracle_ref.py.
 *
 * Build: gcc -O3 -msse4.2 -mavx2 -shared -fPIC  (no -ffast-math: the survey verified the
 * reference build with -ffast-math is bit-identical to this strict order).
 */
#include <emmintrin.h>
#include <stdint.h>

static inline int hsum_epi32(__m128i v)
{
    __m128i t = _mm_add_epi32(v, _mm_srli_si128(v, 8));
    t = _mm_add_epi32(t, _mm_srli_si128(t, 4));
    return _mm_cvtsi128_si32(t);
}

/* exact u8 x u8 -> s32 dot product, 16 pixels per step, scalar tail */
int oracle_row_dot(const uint8_t* k, const uint8_t* s, int n)
{
    const __m128i zero = _mm_setzero_si128();
    __m128i acc = zero;
    int blocks = n / 16, i;
    for (i = 0; i < blocks * 16; i += 16) {
        __m128i a = _mm_loadu_si128((const __m128i*)(k + i));
        __m128i b = _mm_loadu_si128((const __m128i*)(s + i));
        __m128i lo = _mm_madd_epi16(_mm_unpacklo_epi8(a, zero), _mm_unpacklo_epi8(b, zero));
        __m128i hi = _mm_madd_epi16(_mm_unpackhi_epi8(a, zero), _mm_unpackhi_epi8(b, zero));
        acc = _mm_add_epi32(acc, _mm_add_epi32(lo, hi));
    }
    int sum = hsum_epi32(acc);
    for (; i < n; ++i) sum += k[i] * s[i];
    return sum;
}

/* src: sh x sw (continuous), tpl: th x tw (continuous), out: (sh-th+1) x (sw-tw+1) float32 */
void oracle_match_template_simd(const uint8_t* src, int sw, int sh,
                                const uint8_t* tpl, int tw, int th, float* out)
{
    int R = sh - th + 1, C = sw - tw + 1;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
            float acc = 0.0f;
            const uint8_t* s = src + (long)r * sw + c;
            const uint8_t* t = tpl;
            for (int tr = 0; tr < th; ++tr, s += sw, t += tw)
                acc = acc + (float)oracle_row_dot(t, s, tw);
            out[(long)r * C + c] = acc;
        }
}

/* per-row exact sums for one cell: rows[tr] (used to pin the CUDA row-sum kernel) */
void oracle_row_sums(const uint8_t* src, int sw, const uint8_t* tpl, int tw, int th,
                     int r, int c, int32_t* rows)
{
    for (int tr = 0; tr < th; ++tr)
        rows[tr] = oracle_row_dot(tpl + (long)tr * tw, src + (long)(r + tr) * sw + c, tw);
}
