// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// cv_models.cpp -- C++ restatements of the OpenCV routines on the reference's hot path, so that a C++ match() can be
// built and timed in this image (no OpenCV C++ headers / Qt here).  Each model is the arithmetic
// pinned bit-exactly against cv2 4.13 in oracle/models.py + tests/test_oracle.py:
//   cv::pyrDown            (called by cv::buildPyramid, /root/reference/src/TemplateMatcher.cpp:55, :124)
//   cv::getRotationMatrix2D + cv::warpAffine (INTER_LINEAR, BORDER_CONSTANT)   (:163-175, :1082-1089)
//   cv::integral (CV_64F sum and sqsum)                                          (:537)
//   cv::matchTemplate(TM_CCORR) as an exact integer sum (the DFT path of OpenCV is inexact)  (:514)
//   cv::meanStdDev                                                               (:71)
//   cv::RotatedRect(3 pts) / rotatedRectangleIntersection / contourArea + the reference's sortPtWithCenter (:1093-1194):
//     shared, host-compiled, with the device NMS (csrc/fpm_geometry.cuh; differential-tested against cv2, tests/test_abi.py)
// Compiled WITHOUT -ffast-math (OpenCV is a separately built library; only the reference's own sources get its flags).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cv_models.h"
#include "../fastest_image_pattern_matching_b200/csrc/fpm_geometry.cuh"   // host + device geometry (compiled here by nvcc as host code)

namespace cvm {

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

int cv_round(double v) { return (int)lrint(v); }           // round-half-to-even (default rounding mode), like cvRound

void pyr_down(const uint8_t* s, int w, int h, int sp, uint8_t* d, int dp)
{
    const int ow = (w + 1) / 2, oh = (h + 1) / 2;
    // horizontal 5-tap sums of the needed source rows, kept in a ring of 5 rows
    std::vector<int> ring((size_t)5 * ow);
    std::vector<int> ring_row(5, -1000000);
    auto hrow = [&](int sy, int* out) {
        const uint8_t* r = s + (size_t)sy * sp;
        auto edge = [&](int x) {
            const int c = 2 * x;
            out[x] = r[reflect101(c - 2, w)] + 4 * r[reflect101(c - 1, w)] + 6 * r[c] + 4 * r[reflect101(c + 1, w)] + r[reflect101(c + 2, w)];
        };
        const int x_lo = std::min(ow, 1), x_hi = std::max(x_lo, (w - 3) / 2 + 1 > ow ? ow : (w - 3) / 2 + 1);   // interior: 2x-2 >= 0 and 2x+2 < w
        for (int x = 0; x < x_lo; x++) edge(x);
        for (int x = x_lo; x < x_hi; x++) {                                   // branch-free: vectorised by the compiler
            const uint8_t* q = r + 2 * x;
            out[x] = q[-2] + 4 * q[-1] + 6 * q[0] + 4 * q[1] + q[2];
        }
        for (int x = x_hi; x < ow; x++) edge(x);
    };
    for (int y = 0; y < oh; y++) {
        const int* rows[5];
        for (int k = 0; k < 5; k++) {
            const int sy = reflect101(2 * y + k - 2, h);
            int slot = -1;
            for (int q = 0; q < 5; q++) if (ring_row[q] == sy) slot = q;
            if (slot < 0) {
                // evict a slot that this output row does not need
                for (int q = 0; q < 5 && slot < 0; q++) {
                    bool needed = false;
                    for (int kk = 0; kk < 5; kk++) if (ring_row[q] == reflect101(2 * y + kk - 2, h)) needed = true;
                    if (!needed) slot = q;
                }
                hrow(sy, &ring[(size_t)slot * ow]);
                ring_row[slot] = sy;
            }
            rows[k] = &ring[(size_t)slot * ow];
        }
        uint8_t* o = d + (size_t)y * dp;
        for (int x = 0; x < ow; x++)
            o[x] = (uint8_t)((rows[0][x] + 4 * rows[1][x] + 6 * rows[2][x] + 4 * rows[3][x] + rows[4][x] + 128) >> 8);
    }
}

void rotation_matrix(double cx, double cy, double angle_deg, double* m)
{
    const double a = angle_deg * (3.1415926535897932384626433832795 / 180);
    const double alpha = std::cos(a), beta = std::sin(a);
    m[0] = alpha; m[1] = beta;  m[2] = (1 - alpha) * cx - beta * cy;
    m[3] = -beta; m[4] = alpha; m[5] = beta * cx + (1 - alpha) * cy;
}

// forward matrix M (dst = M * src), inverted like cv::warpAffine does without WARP_INVERSE_MAP
void warp_affine(const uint8_t* src, int sw, int sh, int sp, const double* Mf, uint8_t* dst, int dw, int dh, int dp, int border)
{
    double M[6];
    for (int i = 0; i < 6; i++) M[i] = Mf[i];
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    std::vector<int> adelta(dw), bdelta(dw);
    for (int x = 0; x < dw; x++) {
        adelta[x] = cv_round(M[0] * x * 1024.0);
        bdelta[x] = cv_round(M[3] * x * 1024.0);
    }
    std::vector<int> X0v(dh), Y0v(dh);
    for (int y = 0; y < dh; y++) {
        X0v[y] = cv_round((M[1] * y + M[2]) * 1024.0) + 16;
        Y0v[y] = cv_round((M[4] * y + M[5]) * 1024.0) + 16;
    }
    // rows whose taps all lie inside the image take the unchecked loop, walked in 64-column blocks so that the source
    // footprint of a block stays in cache (cv::warpAffine / remap block their work the same way)
    std::vector<char> row_inside(dh);
    for (int y = 0; y < dh; y++) {
        const int X0 = X0v[y], Y0 = Y0v[y];
        const int Xa = (X0 + adelta[0]) >> 10, Xb = (X0 + adelta[dw - 1]) >> 10, Ya = (Y0 + bdelta[0]) >> 10, Yb = (Y0 + bdelta[dw - 1]) >> 10;
        row_inside[y] = std::min(Xa, Xb) >= 0 && std::max(Xa, Xb) < sw - 1 && std::min(Ya, Yb) >= 0 && std::max(Ya, Yb) < sh - 1;
    }
    for (int xb = 0; xb < dw; xb += 64) {
        const int xe = std::min(dw, xb + 64);
        for (int y = 0; y < dh; y++) {
            if (!row_inside[y]) continue;
            const int X0 = X0v[y], Y0 = Y0v[y];
            uint8_t* o = dst + (size_t)y * dp;
            for (int x = xb; x < xe; x++) {
                const int XX = X0 + adelta[x], YY = Y0 + bdelta[x];
                const int ax = (XX >> 5) & 31, ay = (YY >> 5) & 31;
                const uint8_t* p = src + (size_t)(YY >> 10) * sp + (XX >> 10);
                const int p00 = p[0], p01 = p[1], p10 = p[sp], p11 = p[sp + 1];
                const int top = (p00 << 5) + ax * (p01 - p00), bot = (p10 << 5) + ax * (p11 - p10);
                o[x] = (uint8_t)(((top << 5) + ay * (bot - top) + 512) >> 10);
            }
        }
    }
    for (int y = 0; y < dh; y++) {
        if (row_inside[y]) continue;
        const int X0 = X0v[y], Y0 = Y0v[y];
        uint8_t* o = dst + (size_t)y * dp;
        for (int x = 0; x < dw; x++) {
            const int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
            int p00, p01, p10, p11;
            if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
                const uint8_t* p = src + (size_t)sy * sp + sx;
                p00 = p[0]; p01 = p[1]; p10 = p[sp]; p11 = p[sp + 1];
            } else {
                const bool x0 = (unsigned)sx < (unsigned)sw, x1 = (unsigned)(sx + 1) < (unsigned)sw;
                const bool y0 = (unsigned)sy < (unsigned)sh, y1 = (unsigned)(sy + 1) < (unsigned)sh;
                p00 = (x0 && y0) ? src[(size_t)sy * sp + sx] : border;
                p01 = (x1 && y0) ? src[(size_t)sy * sp + sx + 1] : border;
                p10 = (x0 && y1) ? src[(size_t)(sy + 1) * sp + sx] : border;
                p11 = (x1 && y1) ? src[(size_t)(sy + 1) * sp + sx + 1] : border;
            }
            const int top = (p00 << 5) + ax * (p01 - p00), bot = (p10 << 5) + ax * (p11 - p10);
            o[x] = (uint8_t)(((top << 5) + ay * (bot - top) + 512) >> 10);
        }
    }
}

void integral(const uint8_t* s, int w, int h, int sp, double* sum, double* sq)
{
    const int W = w + 1;
    for (int x = 0; x <= w; x++) { sum[x] = 0; sq[x] = 0; }
    for (int y = 0; y < h; y++) {
        const uint8_t* r = s + (size_t)y * sp;
        double* so = sum + (size_t)(y + 1) * W; double* qo = sq + (size_t)(y + 1) * W;
        const double* sa = sum + (size_t)y * W; const double* qa = sq + (size_t)y * W;
        long long rs = 0, rq = 0;                              // exact row prefix (values < 2^53: the f64 results are exact integers)
        so[0] = 0; qo[0] = 0;
        for (int x = 0; x < w; x++) {
            rs += r[x]; rq += (int)r[x] * r[x];
            so[x + 1] = sa[x + 1] + (double)rs;
            qo[x + 1] = qa[x + 1] + (double)rq;
        }
    }
}

void ccorr_exact(const uint8_t* img, int iw, int ih, int ip, const uint8_t* tpl, int tw, int th, int tp, float* out)
{
    const int R = ih - th + 1, C = iw - tw + 1;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) {
            long long acc = 0;
            for (int i = 0; i < th; i++) {
                const uint8_t* a = img + (size_t)(r + i) * ip + c;
                const uint8_t* b = tpl + (size_t)i * tp;
                int row = 0;
                for (int j = 0; j < tw; j++) row += a[j] * b[j];
                acc += row;
            }
            out[(size_t)r * C + c] = (float)acc;
        }
}

void mean_stddev(const uint8_t* s, int w, int h, int sp, double* mean, double* sdv)
{
    unsigned long long S = 0, Q = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) { const unsigned v = s[(size_t)y * sp + x]; S += v; Q += (unsigned long long)v * v; }
    const double scale = 1.0 / ((double)w * h);
    const double m = (double)S * scale;
    const double var = std::max((double)Q * scale - m * m, 0.0);
    *mean = m; *sdv = std::sqrt(var);
}

int rrect_overlap(const float a[5], const float b[5], double max_overlap)
{
    FpmRRect r1{a[0], a[1], a[2], a[3], a[4]}, r2{b[0], b[1], b[2], b[3], b[4]};
    return fpm_rrect_overlap_decision(r1, r2, max_overlap, nullptr, nullptr);
}

void rrect_from3(const float p[6], float out[5])
{
    FpmRRect r = fpm_rrect_from3(p[0], p[1], p[2], p[3], p[4], p[5]);
    out[0] = r.cx; out[1] = r.cy; out[2] = r.w; out[3] = r.h; out[4] = r.angle;
}

// cv::invert(DECOMP_LU) for n > 3 (hal::LU64f on [A | I]) -- same op order as OpenCV's LUImpl
bool lu_inverse(double* A, double* B, int m)
{
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) B[i * m + j] = i == j ? 1.0 : 0.0;
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++) if (std::fabs(A[j * m + i]) > std::fabs(A[k * m + i])) k = j;
        if (std::fabs(A[k * m + i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * m + j], A[k * m + j]);
            for (int j = 0; j < m; j++) std::swap(B[i * m + j], B[k * m + j]);
        }
        const double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            const double alpha = A[j * m + i] * d;
            for (int kk = i + 1; kk < m; kk++) A[j * m + kk] += alpha * A[i * m + kk];
            for (int kk = 0; kk < m; kk++) B[j * m + kk] += alpha * B[i * m + kk];
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < m; j++) {
            double s = B[i * m + j];
            for (int k = i + 1; k < m; k++) s -= A[i * m + k] * B[k * m + j];
            B[i * m + j] = s / A[i * m + i];
        }
    return true;
}

}  // namespace cvm
