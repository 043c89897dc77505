// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// cpu_match.cpp -- "CPU Baseline A" (BASELINE.md 3.2): a C++ restatement of TemplateMatcher::learnPattern / match
// (/root/reference/src/TemplateMatcher.cpp:45-437) built with the reference's own Release flags
// (/root/reference/CMakeLists.txt:67-84: -O3 -ffast-math -funroll-loops -ftree-vectorize -march=x86-64 -mtune=generic
// -msse4.2 -mavx -mavx2), single-threaded like the reference's own loops, with its SSE2 numerator (IM_Conv_SIMD :461-483).
// The OpenCV calls go to oracle/cv_models.cpp (no OpenCV C++ in this image); the top layer uses the exact integer
// TM_CCORR sum instead of OpenCV's DFT path.  Qt TemplateMatcher semantics (angle sign kept, no TargetNum truncation).
// Checked against the Python oracle on the golden cases (tests/test_oracle.py); used as bench.py's CPU arm.
#include <emmintrin.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cv_models.h"

namespace {

inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now().time_since_epoch()).count(); }

const double kTol = 0.0000001, kPi = 3.1415926535897932384626433832795, kD2R = kPi / 180.0, kR2D = 180.0 / kPi;
const int kCandNum = 5;

struct Img {
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    Img() {}
    Img(int w_, int h_) : w(w_), h(h_), px((size_t)w_ * h_) {}
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
};

struct Cand { double ptx, pty, score, angle; int id; bool on_border = false, has_vr = false; double vr[3][3]; };

struct Matcher {
    int max_pos = 70; double max_overlap = 0.0, score = 0.7, tol = 0.0; int mra = 256; bool use_simd = true, subpix = false;
    std::vector<Img> tpl; std::vector<double> mean, norm, inv_area; std::vector<char> equal1; int border = 0; bool learned = false;
    double last_ms = 0;
    double t_pyr = 0, t_warp = 0, t_conv = 0, t_den = 0, t_top = 0;   // per-stage milliseconds of the last match (cpum_stage_ms)
    std::vector<double> isum, isq;          // integral scratch, reused across calls
    Img roi;
};

int top_layer(int w, int h, int min_len)
{
    int top = 0; const int min_area = min_len * min_len; int area = w * h;
    while (area > min_area) { area /= 4; top++; }
    return top;
}

std::vector<Img> pyramid(const Img& l0, int top)
{
    std::vector<Img> p; p.push_back(l0);
    for (int l = 0; l < top; l++) {
        const Img& s = p.back();
        Img d((s.w + 1) / 2, (s.h + 1) / 2);
        cvm::pyr_down(s.px.data(), s.w, s.h, s.w, d.px.data(), d.w);
        p.push_back(std::move(d));
    }
    return p;
}

// ptRotatePt2f, :971-982
void pt_rotate(float px, float py, float ox, float oy, double ang, float* rx, float* ry)
{
    const double dHeight = (double)(oy * 2);
    const double dY1 = dHeight - (double)py, dY2 = dHeight - (double)oy;
    const double c = std::cos(ang), s = std::sin(ang);
    const double dX = ((double)px - (double)ox) * c - (dY1 - (double)oy) * s + (double)ox;
    double dY = ((double)px - (double)ox) * s + (dY1 - (double)oy) * c + dY2;
    dY = -dY + dHeight;
    *rx = (float)dX; *ry = (float)dY;
}

// getBestRotationSize, :901-969
void best_rotation_size(int sw, int sh, int dw, int dh, double ang, int* ow, int* oh)
{
    const double rad = ang * kD2R;
    const float cx = (sw - 1) / 2.0f, cy = (sh - 1) / 2.0f;
    float x[4], y[4];
    pt_rotate(0.f, 0.f, cx, cy, rad, &x[0], &y[0]);
    pt_rotate(0.f, (float)(sh - 1), cx, cy, rad, &x[1], &y[1]);
    pt_rotate((float)(sw - 1), (float)(sh - 1), cx, cy, rad, &x[2], &y[2]);
    pt_rotate((float)(sw - 1), 0.f, cx, cy, rad, &x[3], &y[3]);
    const float topY = std::max(std::max(y[0], y[1]), std::max(y[2], y[3])), botY = std::min(std::min(y[0], y[1]), std::min(y[2], y[3]));
    const float rightX = std::max(std::max(x[0], x[1]), std::max(x[2], x[3])), leftX = std::min(std::min(x[0], x[1]), std::min(x[2], x[3]));
    if (ang > 360) ang -= 360; else if (ang < 0) ang += 360;
    if (std::fabs(std::fabs(ang) - 90) < kTol || std::fabs(std::fabs(ang) - 270) < kTol) { *ow = sh; *oh = sw; return; }
    if (std::fabs(ang) < kTol || std::fabs(std::fabs(ang) - 180) < kTol) { *ow = sw; *oh = sh; return; }
    double a = ang;
    if (a > 0 && a < 90) {} else if (a > 90 && a < 180) a -= 90; else if (a > 180 && a < 270) a -= 180; else if (a > 270 && a < 360) a -= 270;
    const float h1 = (float)(dw * std::sin(a * kD2R) * std::cos(a * kD2R)), h2 = (float)(dh * std::sin(a * kD2R) * std::cos(a * kD2R));
    const int halfH = (int)std::ceil(topY - cy - h1), halfW = (int)std::ceil(rightX - cx - h2);
    int rw = halfW * 2, rh = halfH * 2;
    if ((dw < rw && dh > rh) || (dw > rw && dh < rh) || ((long long)dw * dh > (long long)rw * rh)) {
        rw = (int)((double)(rightX - leftX) + 0.5); rh = (int)((double)(topY - botY) + 0.5);
    }
    *ow = rw; *oh = rh;
}

// IM_Conv_SIMD, :461-483 (+ _mm_hsum_epi32, :21-26)
inline int hsum(__m128i v)
{
    __m128i t = _mm_add_epi32(v, _mm_srli_si128(v, 8));
    t = _mm_add_epi32(t, _mm_srli_si128(t, 4));
    return _mm_cvtsi128_si32(t);
}
inline int conv_simd(const uint8_t* k, const uint8_t* s, int n)
{
    const __m128i zero = _mm_setzero_si128();
    __m128i acc = zero;
    const int blocks = n / 16;
    int i = 0;
    for (; i < blocks * 16; i += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(k + i)), b = _mm_loadu_si128((const __m128i*)(s + i));
        const __m128i lo = _mm_madd_epi16(_mm_unpacklo_epi8(a, zero), _mm_unpacklo_epi8(b, zero));
        const __m128i hi = _mm_madd_epi16(_mm_unpackhi_epi8(a, zero), _mm_unpackhi_epi8(b, zero));
        acc = _mm_add_epi32(acc, _mm_add_epi32(lo, hi));
    }
    int sum = hsum(acc);
    for (; i < n; ++i) sum += k[i] * s[i];
    return sum;
}

// MatchTemplate (:485-525) + CCOEFF_Denominator (:527-598)
void match_template(Matcher& m, const Img& src, int layer, bool use_simd, std::vector<float>& res, int* R, int* C)
{
    const Img& t = m.tpl[layer];
    const int rr = src.h - t.h + 1, cc = src.w - t.w + 1;
    *R = rr; *C = cc;
    res.assign((size_t)rr * cc, 0.f);
    const double tc0 = now_ms();
    if (m.use_simd && use_simd) {
        for (int r = 0; r < rr; r++)
            for (int c = 0; c < cc; c++) {
                float* cell = &res[(size_t)r * cc + c];
                const uint8_t* s = src.row(r) + c; const uint8_t* k = t.px.data();
                for (int tr = 0; tr < t.h; tr++, s += src.w, k += t.w) *cell = *cell + conv_simd(k, s, t.w);
            }
    } else {
        cvm::ccorr_exact(src.px.data(), src.w, src.h, src.w, t.px.data(), t.w, t.h, t.w, res.data());
    }
    const double tc1 = now_ms();
    m.t_conv += tc1 - tc0;
    if (m.equal1[layer]) { std::fill(res.begin(), res.end(), 1.f); return; }
    std::vector<double>& sum = m.isum; std::vector<double>& sq = m.isq;
    if (sum.size() < (size_t)(src.h + 1) * (src.w + 1)) { sum.resize((size_t)(src.h + 1) * (src.w + 1)); sq.resize(sum.size()); }
    cvm::integral(src.px.data(), src.w, src.h, src.w, sum.data(), sq.data());
    const int W = src.w + 1;
    const double tmean = m.mean[layer], tnorm = m.norm[layer], inv_area = m.inv_area[layer];
    for (int i = 0; i < rr; i++)
        for (int j = 0; j < cc; j++) {
            const double* p0 = &sum[(size_t)i * W + j]; const double* q0 = &sq[(size_t)i * W + j];
            double num = res[(size_t)i * cc + j], tt;
            double wndMean2 = 0, wndSum2 = 0;
            tt = p0[0] - p0[t.w] - p0[(size_t)t.h * W] + p0[(size_t)t.h * W + t.w];
            wndMean2 += tt * tt;
            num -= tt * tmean;
            wndMean2 *= inv_area;
            tt = q0[0] - q0[t.w] - q0[(size_t)t.h * W] + q0[(size_t)t.h * W + t.w];
            wndSum2 += tt;
            const double diff2 = std::max(wndSum2 - wndMean2, 0.0);
            if (diff2 <= std::min(0.5, 10 * 1.1920928955078125e-07 * wndSum2)) tt = 0; else tt = std::sqrt(diff2) * tnorm;
            if (std::fabs(num) < tt) num /= tt; else if (std::fabs(num) < tt * 1.125) num = num > 0 ? 1 : -1; else num = 0;
            res[(size_t)i * cc + j] = (float)num;
        }
    m.t_den += now_ms() - tc1;
}

void min_max_loc(const float* m, int R, int C, double* val, int* x, int* y)
{
    float best = m[0]; int bi = 0;
    for (int i = 1; i < R * C; i++) if (m[i] > best) { best = m[i]; bi = i; }
    *val = best; *x = bi % C; *y = bi / C;
}

void paint(float* m, int R, int C, int x, int y, int w, int h)
{
    if (w <= 0 || h <= 0) return;
    const int x0 = std::max(x, 0), y0 = std::max(y, 0), x1 = std::min(x + w, C), y1 = std::min(y + h, R);
    for (int yy = y0; yy < y1; yy++) for (int xx = x0; xx < x1; xx++) m[(size_t)yy * C + xx] = -1.f;
}

// s_BlockMax (Qt flavour), DataStructures.h:118-245
struct BlockMax {
    struct B { int x, y, w, h; double v; int px, py; };
    std::vector<B> blocks; float* mat; int R, C;
    void scan(B& b) { float best = mat[(size_t)b.y * C + b.x]; int bx = b.x, by = b.y;
        for (int yy = b.y; yy < b.y + b.h; yy++) for (int xx = b.x; xx < b.x + b.w; xx++) { const float v = mat[(size_t)yy * C + xx]; if (v > best) { best = v; bx = xx; by = yy; } }
        b.v = best; b.px = bx; b.py = by; }
    void add(int x, int y, int w, int h) { B b{x, y, w, h, 0, 0, 0}; scan(b); blocks.push_back(b); }
    BlockMax(float* m, int R_, int C_, int bw, int bh) : mat(m), R(R_), C(C_) {
        const int ncol = C / bw, nrow = R / bh;
        for (int y = 0; y < nrow; y++) for (int x = 0; x < ncol; x++) add(x * bw, y * bh, bw, bh);
        if (ncol * bw < C) add(ncol * bw, 0, C - ncol * bw, R);
        if (nrow * bh < R && ncol * bw > 0) add(0, nrow * bh, ncol * bw, R - nrow * bh);
        if (ncol * bw < C && nrow * bh < R) add(ncol * bw, nrow * bh, C - ncol * bw, R - nrow * bh);
    }
    void update(int rx, int ry, int rw, int rh) {
        for (B& b : blocks) { const int x0 = std::max(b.x, rx), y0 = std::max(b.y, ry), x1 = std::min(b.x + b.w, rx + rw), y1 = std::min(b.y + b.h, ry + rh);
            if (x1 > x0 && y1 > y0) scan(b); }
    }
    void get(double* v, int* x, int* y) {
        if (blocks.empty()) { *v = -1; *x = -1; *y = -1; return; }
        const B* best = &blocks[0];
        for (size_t i = 1; i < blocks.size(); i++) if (best->v < blocks[i].v) best = &blocks[i];
        *v = best->v; *x = best->px; *y = best->py;
    }
};

// subPixEstimation, :1002-1072
void subpix(const Cand* nw, double step, int best, double* ox, double* oy, double* oa)
{
    double A[27 * 10], S[27];
    const double xm = nw[best].ptx, ym = nw[best].pty, tm = nw[best].angle;
    int row = 0;
    for (int th = 0; th <= 2; th++) for (int y = -1; y <= 1; y++) for (int x = -1; x <= 1; x++) {
        const double dX = xm + x, dY = ym + y, dT = (tm + (th - 1) * step) * kD2R;
        double* a = A + row * 10;
        a[0] = dX * dX; a[1] = dY * dY; a[2] = dT * dT; a[3] = dX * dY; a[4] = dX * dT; a[5] = dY * dT; a[6] = dX; a[7] = dY; a[8] = dT; a[9] = 1.0;
        const Cand& c = nw[best + (th - 1)];
        S[row] = c.has_vr ? c.vr[x + 1][y + 1] : 0.0;
        row++;
    }
    double AtA[100], Inv[100], P[270], Z[10];
    for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) { double s = 0; for (int k = 0; k < 27; k++) s += A[k * 10 + i] * A[k * 10 + j]; AtA[i * 10 + j] = s; }
    if (!cvm::lu_inverse(AtA, Inv, 10)) std::fill(Inv, Inv + 100, 0.0);
    for (int i = 0; i < 10; i++) for (int j = 0; j < 27; j++) { double s = 0; for (int k = 0; k < 10; k++) s += Inv[i * 10 + k] * A[j * 10 + k]; P[i * 27 + j] = s; }
    for (int i = 0; i < 10; i++) { double s = 0; for (int k = 0; k < 27; k++) s += P[i * 27 + k] * S[k]; Z[i] = s; }
    const double S00 = 2 * Z[0], S01 = Z[3], S02 = Z[4], S10 = Z[3], S11 = 2 * Z[1], S12 = Z[5], S20 = Z[4], S21 = Z[5], S22 = 2 * Z[2];
    const double b0 = -Z[6], b1 = -Z[7], b2 = -Z[8];
    double d = S00 * (S11 * S22 - S12 * S21) - S01 * (S10 * S22 - S12 * S20) + S02 * (S10 * S21 - S11 * S20);
    if (d != 0.) {
        d = 1. / d;
        *ox = ((S11 * S22 - S12 * S21) * b0 + (S02 * S21 - S01 * S22) * b1 + (S01 * S12 - S02 * S11) * b2) * d;
        *oy = ((S12 * S20 - S10 * S22) * b0 + (S00 * S22 - S02 * S20) * b1 + (S02 * S10 - S00 * S12) * b2) * d;
        *oa = (((S10 * S21 - S11 * S20) * b0 + (S01 * S20 - S00 * S21) * b1 + (S00 * S11 - S01 * S10) * b2) * d) * kR2D;
    } else { *ox = 0; *oy = 0; *oa = 0; }
}

void corners(double ptx, double pty, double angle, int w, int h, float* lt, float* rt, float* lb, float* rb)
{
    const double ra = -angle * kD2R;
    const float c = (float)std::cos(ra), s = (float)std::sin(ra);
    lt[0] = (float)ptx; lt[1] = (float)pty;
    rt[0] = lt[0] + w * c; rt[1] = lt[1] - w * s;
    lb[0] = lt[0] + h * s; lb[1] = lt[1] + h * c;
    rb[0] = rt[0] + h * s; rb[1] = rt[1] + h * c;
}

bool learn(Matcher& m, const uint8_t* tpl, int w, int h)
{
    if (!tpl || w <= 0 || h <= 0) return false;
    Img t0(w, h); memcpy(t0.px.data(), tpl, (size_t)w * h);
    const int top = top_layer(w, h, (int)std::sqrt((double)m.mra));
    m.tpl = pyramid(t0, top);
    m.mean.clear(); m.norm.clear(); m.inv_area.clear(); m.equal1.clear();
    double mean0, sd0; cvm::mean_stddev(tpl, w, h, w, &mean0, &sd0);
    m.border = mean0 < 128 ? 255 : 0;
    for (const Img& l : m.tpl) {
        const double inv_area = 1.0 / ((double)l.h * l.w);
        double mean, sdv; cvm::mean_stddev(l.px.data(), l.w, l.h, l.w, &mean, &sdv);
        double norm = sdv * sdv;
        m.equal1.push_back(norm < 2.220446049250313e-16);
        norm = std::sqrt(norm); norm /= std::sqrt(inv_area);
        m.inv_area.push_back(inv_area); m.mean.push_back(mean); m.norm.push_back(norm);
    }
    m.learned = true;
    return true;
}

int match(Matcher& m, const uint8_t* srcp, int sw, int sh, double* out, int cap)
{
    const auto t_start = std::chrono::high_resolution_clock::now();
    if (!srcp || sw <= 0 || sh <= 0 || !m.learned) return 0;
    const int t0w = m.tpl[0].w, t0h = m.tpl[0].h;
    if ((t0w < sw && t0h > sh) || (t0w > sw && t0h < sh)) return 0;
    if ((long long)t0w * t0h > (long long)sw * sh) return 0;
    const int top = top_layer(t0w, t0h, (int)std::sqrt((double)m.mra));
    if (top >= (int)m.tpl.size()) return -1;
    Img s0(sw, sh); memcpy(s0.px.data(), srcp, (size_t)sw * sh);         // m_sourceImage = src.clone(), :104
    m.t_pyr = m.t_warp = m.t_conv = m.t_den = m.t_top = 0;
    const double tp0 = now_ms();
    const std::vector<Img> spyr = pyramid(s0, top);
    m.t_pyr = now_ms() - tp0;
    const Img& tp = m.tpl[top];
    const double step_top = std::atan(2.0 / std::max(tp.w, tp.h)) * kR2D;
    std::vector<double> angles;
    if (m.tol < kTol) angles.push_back(0.0);
    else {
        for (double a = 0; a < m.tol + step_top; a += step_top) angles.push_back(a);
        for (double a = -step_top; a > -m.tol - step_top; a -= step_top) angles.push_back(a);
    }
    const Img& ts = spyr[top];
    const float cx = (ts.w - 1) / 2.0f, cy = (ts.h - 1) / 2.0f;
    std::vector<double> layer_score(top + 1); layer_score[0] = m.score;
    for (int l = 1; l <= top; l++) layer_score[l] = layer_score[l - 1] * 0.9;
    const bool by_block = ((ts.w * ts.h) / (tp.w * tp.h) > 500) && m.max_pos > 10;
    std::vector<Cand> cands;
    std::vector<float> res;
    for (double ang : angles) {
        double M[6]; cvm::rotation_matrix(cx, cy, ang, M);
        int bw, bh; best_rotation_size(ts.w, ts.h, tp.w, tp.h, ang, &bw, &bh);
        const float ftx = (bw - 1) / 2.0f - cx, fty = (bh - 1) / 2.0f - cy;
        M[2] += (double)ftx; M[5] += (double)fty;
        Img rot(bw, bh);
        cvm::warp_affine(ts.px.data(), ts.w, ts.h, ts.w, M, rot.px.data(), bw, bh, bw, m.border);
        int R, C; match_template(m, rot, top, false, res, &R, &C);
        const double ov = m.max_overlap;
        auto push = [&](int x, int y, double v) { Cand c; c.ptx = (double)((float)x - ftx); c.pty = (double)((float)y - fty); c.score = v; c.angle = ang; c.id = 0; cands.push_back(c); };
        double v; int x, y;
        if (by_block) {
            BlockMax bm(res.data(), R, C, tp.w, tp.h);
            bm.get(&v, &x, &y);
            if (v < layer_score[top]) continue;
            push(x, y, v);
            for (int j = 0; j < m.max_pos + kCandNum - 1; j++) {
                const int sx = (int)(x - tp.w * (1 - ov)), sy = (int)(y - tp.h * (1 - ov)), rw = (int)(2 * tp.w * (1 - ov)), rh = (int)(2 * tp.h * (1 - ov));
                paint(res.data(), R, C, sx, sy, rw, rh);
                bm.update(sx, sy, rw, rh);
                bm.get(&v, &x, &y);
                if (v < layer_score[top]) break;
                push(x, y, v);
            }
        } else {
            min_max_loc(res.data(), R, C, &v, &x, &y);
            if (v < layer_score[top]) continue;
            push(x, y, v);
            for (int j = 0; j < m.max_pos + kCandNum - 1; j++) {
                paint(res.data(), R, C, (int)(x - tp.w * (1 - ov)), (int)(y - tp.h * (1 - ov)), (int)(2 * tp.w * (1 - ov)), (int)(2 * tp.h * (1 - ov)));
                min_max_loc(res.data(), R, C, &v, &x, &y);
                if (v < layer_score[top]) break;
                push(x, y, v);
            }
        }
    }
    m.t_top = now_ms() - tp0 - m.t_pyr;
    m.t_conv = 0; m.t_den = 0;                               // count the refinement only (the top layer is t_top)
    std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });   // :214 (stable, like the Python oracle)
    for (size_t i = 0; i < cands.size(); i++) cands[i].id = (int)i;

    std::vector<Cand> all;
    const bool one_angle = m.tol < kTol;
    for (Cand cand : cands) {
        float ltx, lty;
        pt_rotate((float)cand.ptx, (float)cand.pty, cx, cy, -cand.angle * kD2R, &ltx, &lty);
        if (top <= 0) { cand.ptx = ltx; cand.pty = lty; all.push_back(cand); continue; }
        bool alive = true;
        for (int layer = top - 1; layer >= 0 && alive; layer--) {
            const Img& tl = m.tpl[layer]; const Img& sl = spyr[layer];
            const double a_step = std::atan(2.0 / std::max(tl.w, tl.h)) * kR2D;
            const int n_ang = one_angle ? 1 : 3;
            double l_angles[3] = {0.0, 0.0, 0.0};
            if (!one_angle) for (int i = 0; i < 3; i++) l_angles[i] = cand.angle + a_step * (i - 1);
            const float scx = (sl.w - 1) / 2.0f, scy = (sl.h - 1) / 2.0f;
            const float l2x = ltx * 2, l2y = lty * 2;
            Cand nw[3]; int best = 0; double big = -1;
            for (int j = 0; j < n_ang; j++) {
                // getRotatedROI, :1074-1090
                float rx, ry; pt_rotate(l2x, l2y, scx, scy, l_angles[j] * kD2R, &rx, &ry);
                double M[6]; cvm::rotation_matrix(scx, scy, l_angles[j], M);
                M[2] -= (double)(rx - 3); M[5] -= (double)(ry - 3);
                Img& roi = m.roi; roi.w = tl.w + 6; roi.h = tl.h + 6;
                if (roi.px.size() < (size_t)roi.w * roi.h) roi.px.resize((size_t)roi.w * roi.h);
                const double tw0 = now_ms();
                cvm::warp_affine(sl.px.data(), sl.w, sl.h, sl.w, M, roi.px.data(), roi.w, roi.h, roi.w, 0);
                m.t_warp += now_ms() - tw0;
                int R, C; match_template(m, roi, layer, true, res, &R, &C);
                double v; int x, y; min_max_loc(res.data(), R, C, &v, &x, &y);
                Cand& p = nw[j]; p = Cand(); p.ptx = x; p.pty = y; p.score = v; p.angle = l_angles[j]; p.id = cand.id;
                if (p.score > big) { best = j; big = p.score; }
                p.on_border = (x == 0 || y == 0 || x == C - 1 || y == R - 1);
                if (!p.on_border) { p.has_vr = true; for (int yy = -1; yy <= 1; yy++) for (int xx = -1; xx <= 1; xx++) p.vr[xx + 1][yy + 1] = res[(size_t)(y + yy) * C + x + xx]; }
            }
            if (nw[best].score < layer_score[layer]) { alive = false; break; }
            if (m.subpix && layer == 0 && !nw[best].on_border && best != 0 && best != 2 && n_ang == 3) {
                double nx, ny, na; subpix(nw, a_step, best, &nx, &ny, &na);
                nw[best].ptx = nx; nw[best].pty = ny; nw[best].angle = na;
            }
            const double new_angle = nw[best].angle;
            float padx, pady; pt_rotate(l2x, l2y, scx, scy, new_angle * kD2R, &padx, &pady);
            padx -= 3.0f; pady -= 3.0f;
            const float qx = (float)(nw[best].ptx + (double)padx), qy = (float)(nw[best].pty + (double)pady);
            float px, py; pt_rotate(qx, qy, scx, scy, -new_angle * kD2R, &px, &py);
            if (layer == 0) { nw[best].ptx = px; nw[best].pty = py; all.push_back(nw[best]); }
            else { cand.angle = new_angle; ltx = px; lty = py; }
        }
    }
    // filterWithScore, :984-1000
    std::stable_sort(all.begin(), all.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });
    for (size_t i = 0; i < all.size(); i++) if (all[i].score < m.score) { all.resize(i); break; }
    // filterWithRotatedRect, :1133-1194
    const int n = (int)all.size();
    std::vector<float> rect((size_t)n * 5);
    std::vector<char> del(n, 0);
    for (int i = 0; i < n; i++) {
        float lt[2], rt[2], lb[2], rb[2];
        corners(all[i].ptx, all[i].pty, all[i].angle, m.tpl[0].w, m.tpl[0].h, lt, rt, lb, rb);
        const float p[6] = {lt[0], lt[1], rt[0], rt[1], rb[0], rb[1]};
        cvm::rrect_from3(p, &rect[(size_t)i * 5]);
    }
    for (int i = 0; i < n - 1; i++) {
        if (del[i]) continue;
        for (int j = i + 1; j < n; j++) {
            if (del[j]) continue;
            if (cvm::rrect_overlap(&rect[(size_t)i * 5], &rect[(size_t)j * 5], m.max_overlap)) del[all[i].score >= all[j].score ? j : i] = 1;
        }
    }
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (del[i]) continue;
        if (k < cap) {
            float lt[2], rt[2], lb[2], rb[2];
            corners(all[i].ptx, all[i].pty, all[i].angle, m.tpl[0].w, m.tpl[0].h, lt, rt, lb, rb);
            double* o = out + (size_t)k * 12;
            o[0] = all[i].score; o[1] = all[i].angle;
            o[2] = (double)((lt[0] + rt[0] + lb[0] + rb[0]) / 4.0f); o[3] = (double)((lt[1] + rt[1] + lb[1] + rb[1]) / 4.0f);
            o[4] = lt[0]; o[5] = lt[1]; o[6] = rt[0]; o[7] = rt[1]; o[8] = rb[0]; o[9] = rb[1]; o[10] = lb[0]; o[11] = lb[1];
        }
        k++;
    }
    m.last_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();
    return k;
}

}  // namespace

extern "C" {
void* cpum_create() { return new Matcher(); }
void cpum_destroy(void* h) { delete static_cast<Matcher*>(h); }
// same parameter ids as include/fpm_b200.h: 0 MaxPositions, 1 MaxOverlap, 2 Score, 3 ToleranceAngle, 4 MinReduceArea, 5 UseSIMD, 6 SubPixel
void cpum_set(void* h, int p, double v)
{
    Matcher& m = *static_cast<Matcher*>(h);
    switch (p) { case 0: m.max_pos = (int)v; break; case 1: m.max_overlap = v; break; case 2: m.score = v; break; case 3: m.tol = v; break;
                 case 4: m.mra = (int)v; break; case 5: m.use_simd = v != 0; break; case 6: m.subpix = v != 0; break; }
}
int cpum_learn(void* h, const uint8_t* tpl, int w, int hgt) { return learn(*static_cast<Matcher*>(h), tpl, w, hgt) ? 0 : -1; }
// out: rows of 12 doubles in the fpm_result field order; returns the number of targets found (may exceed cap)
int cpum_match(void* h, const uint8_t* src, int w, int hgt, double* out, int cap) { return match(*static_cast<Matcher*>(h), src, w, hgt, out, cap); }
double cpum_last_ms(void* h) { return static_cast<Matcher*>(h)->last_ms; }
// stage: 0 pyramid, 1 top-layer sweep, 2 ROI warps, 3 SIMD numerators, 4 integrals + denominators (refinement)
double cpum_stage_ms(void* h, int stage)
{
    const Matcher& m = *static_cast<Matcher*>(h);
    const double v[5] = {m.t_pyr, m.t_top, m.t_warp, m.t_conv, m.t_den};
    return stage >= 0 && stage < 5 ? v[stage] : 0;
}
}
