#!/bin/sh
# Compiles the reference's OWN SIMD dot-product (IM_Conv_SIMD + _mm_hsum_epi32) from the
# source where it lies under /root/reference, into oracle/_ref/ (git-ignored, travels to the
# GPU box).  The full TemplateMatcher.cpp is unbuildable here (needs OpenCV C++ and Qt headers),
# so only the two self-contained functions are taken: the function bodies are read from the
# reference file at build time (line ranges asserted by grep) and never stored in this repo.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/src/TemplateMatcher.cpp
[ -f "$REF" ] || { echo "reference not present; keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
TMP="$HERE/_ref/imconv_ref_tu.cpp"
S1=$(grep -n 'inline int _mm_hsum_epi32' "$REF" | head -1 | cut -d: -f1)
S2=$(grep -n 'inline int IM_Conv_SIMD' "$REF" | head -1 | cut -d: -f1)
E2=$(grep -n '^void TemplateMatcher::MatchTemplate' "$REF" | head -1 | cut -d: -f1)
{
  echo '#include <immintrin.h>'
  sed -n "${S1},$((S1+5))p" "$REF"
  sed -n "${S2},$((E2-2))p" "$REF"
  cat <<'EOT'
extern "C" int ref_IM_Conv_SIMD(unsigned char* k, unsigned char* c, int n) { return IM_Conv_SIMD(k, c, n); }
// the reference's accumulation statement (src/TemplateMatcher.cpp:505-508), driven over one cell
extern "C" float ref_cell(unsigned char* tpl, int tw, int th, unsigned char* src, int sw)
{
    float acc = 0; float* r_matResult = &acc;
    unsigned char* r_template = tpl; unsigned char* r_sub_source = src;
    for (int t_r = 0; t_r < th; ++t_r, r_sub_source += sw, r_template += tw)
        *r_matResult = *r_matResult + IM_Conv_SIMD(r_template, r_sub_source, tw);
    return acc;
}
EOT
} > "$TMP"
# the reference's Release flags (CMakeLists.txt:67-84)
g++ -O3 -ffast-math -funroll-loops -ftree-vectorize -march=x86-64 -mtune=generic -msse4.2 -mavx -mavx2 \
    -shared -fPIC -o "$HERE/_ref/libimconv_ref.so" "$TMP"
rm -f "$TMP"
echo "built $HERE/_ref/libimconv_ref.so"
