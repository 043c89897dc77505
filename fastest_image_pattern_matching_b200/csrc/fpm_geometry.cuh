// fpm_geometry.cuh -- rotated-rectangle geometry used by the on-device NMS.
//
// Replaces, for filterWithRotatedRect (/root/reference/src/TemplateMatcher.cpp:1133-1194):
//   cv::RotatedRect(p1,p2,p3)            (:389)
//   cv::rotatedRectangleIntersection     (:1150)
//   sortPtWithCenter                     (:1093-1131, quirks kept on purpose)
//   cv::contourArea                      (:1171)
// All arithmetic is float32/float64 in the same order as the OpenCV 4.x implementation so the
// accept/reject decisions agree with the cv2-backed oracle (differentially tested on the CPU via
// fpm_dbg_rrect_overlap, and on the GPU inside the NMS kernel).
#pragma once
#include "fpm_common.cuh"
#include <float.h>

struct FpmRRect { float cx, cy, w, h, angle; };

enum { FPM_INTERSECT_NONE = 0, FPM_INTERSECT_PARTIAL = 1, FPM_INTERSECT_FULL = 2 };

FPM_HD double fpm_norm2f(float x, float y) { return sqrt((double)x * x + (double)y * y); }

// cv::RotatedRect::RotatedRect(const Point2f&, const Point2f&, const Point2f&) without the
// perpendicularity assertion (the reference would throw there, README.md:137-171).
FPM_HD FpmRRect fpm_rrect_from3(float p1x, float p1y, float p2x, float p2y, float p3x, float p3y)
{
    FpmRRect r;
    r.cx = 0.5f * (p1x + p3x);
    r.cy = 0.5f * (p1y + p3y);
    float v0x = p1x - p2x, v0y = p1y - p2y;
    float v1x = p2x - p3x, v1y = p2y - p3y;
    int wd_i = 0;
    if (fabsf(v1y) < fabsf(v1x)) wd_i = 1;
    float wx = wd_i ? v1x : v0x, wy = wd_i ? v1y : v0y;
    float hx = wd_i ? v0x : v1x, hy = wd_i ? v0y : v1y;
    r.angle = atanf(wy / wx) * 180.0f / (float)FPM_PI;
    r.w = (float)fpm_norm2f(wx, wy);
    r.h = (float)fpm_norm2f(hx, hy);
    return r;
}

// cv::RotatedRect::points
FPM_HD void fpm_rrect_points(const FpmRRect& r, float* px, float* py)
{
    double ang = r.angle * FPM_PI / 180.;
    float b = (float)cos(ang) * 0.5f;
    float a = (float)sin(ang) * 0.5f;
    px[0] = r.cx - a * r.h - b * r.w;
    py[0] = r.cy + b * r.h - a * r.w;
    px[1] = r.cx + a * r.h - b * r.w;
    py[1] = r.cy - b * r.h - a * r.w;
    px[2] = 2 * r.cx - px[0];
    py[2] = 2 * r.cy - py[0];
    px[3] = 2 * r.cx - px[1];
    py[3] = 2 * r.cy - py[1];
}

// static _rotatedRectangleIntersection of OpenCV 4.x (imgproc/src/intersection.cpp)
FPM_HD int fpm_rrect_intersection_shifted(const FpmRRect& rect1, const FpmRRect& rect2,
                                          float* ix, float* iy, int* n_out)
{
    float samePointEps = 1e-6f * fmaxf(rect1.w * rect1.h, rect2.w * rect2.h);
    float vec1x[4], vec1y[4], vec2x[4], vec2y[4];
    float p1x[4], p1y[4], p2x[4], p2y[4];
    fpm_rrect_points(rect1, p1x, p1y);
    fpm_rrect_points(rect2, p2x, p2y);
    int ret = FPM_INTERSECT_FULL;
    int n = 0;

    bool same = true;
    for (int i = 0; i < 4; i++)
        if (fabsf(p1x[i] - p2x[i]) > samePointEps || fabsf(p1y[i] - p2y[i]) > samePointEps) { same = false; break; }
    if (same) {
        for (int i = 0; i < 4; i++) { ix[i] = p1x[i]; iy[i] = p1y[i]; }
        *n_out = 4;
        return FPM_INTERSECT_FULL;
    }
    for (int i = 0; i < 4; i++) {
        vec1x[i] = p1x[(i + 1) % 4] - p1x[i];
        vec1y[i] = p1y[(i + 1) % 4] - p1y[i];
        vec2x[i] = p2x[(i + 1) % 4] - p2x[i];
        vec2y[i] = p2y[(i + 1) % 4] - p2y[i];
    }
    // adapt the epsilon to the smallest dimension of the rects
    for (int i = 0; i < 4; i++) {
        samePointEps = fminf(samePointEps, sqrtf(vec1x[i] * vec1x[i] + vec1y[i] * vec1y[i]));
        samePointEps = fminf(samePointEps, sqrtf(vec2x[i] * vec2x[i] + vec2y[i] * vec2y[i]));
    }
    samePointEps = fmaxf(1e-16f, samePointEps);

    // line test: all 16 edge pairs
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float x21 = p2x[j] - p1x[i];
            float y21 = p2y[j] - p1y[i];
            float vx1 = vec1x[i], vy1 = vec1y[i];
            float vx2 = vec2x[j], vy2 = vec2y[j];
            float normalizationScale = fminf(vx1 * vx1 + vy1 * vy1, vx2 * vx2 + vy2 * vy2);
            normalizationScale = (normalizationScale < 1e-12f) ? 1.f : 1.f / normalizationScale;
            vx1 *= normalizationScale; vy1 *= normalizationScale;
            vx2 *= normalizationScale; vy2 *= normalizationScale;
            const float det = vx2 * vy1 - vx1 * vy2;
            if (fabsf(det) < 1e-12f) continue;
            const float detInvScaled = normalizationScale / det;
            const float t1 = (vx2 * y21 - vy2 * x21) * detInvScaled;
            const float t2 = (vx1 * y21 - vy1 * x21) * detInvScaled;
            if (isinf(t1) || isinf(t2) || isnan(t1) || isnan(t2)) continue;
            if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) {
                ix[n] = p1x[i] + vec1x[i] * t1;
                iy[n] = p1y[i] + vec1y[i] * t1;
                n++;
            }
        }
    if (n > 0) ret = FPM_INTERSECT_PARTIAL;

    // vertices of rect1 inside rect2
    for (int i = 0; i < 4; i++) {
        int posSign = 0, negSign = 0;
        float x = p1x[i], y = p1y[i];
        for (int j = 0; j < 4; j++) {
            float normalizationScale = vec2x[j] * vec2x[j] + vec2y[j] * vec2y[j];
            normalizationScale = (normalizationScale < 1e-12f) ? 1.f : 1.f / normalizationScale;
            float A = -vec2y[j] * normalizationScale;
            float B = vec2x[j] * normalizationScale;
            float C = -(A * p2x[j] + B * p2y[j]);
            float s = A * x + B * y + C;
            if (s >= 0) posSign++; else negSign++;
        }
        if (posSign == 4 || negSign == 4) { ix[n] = p1x[i]; iy[n] = p1y[i]; n++; }
    }
    // vertices of rect2 inside rect1
    for (int i = 0; i < 4; i++) {
        int posSign = 0, negSign = 0;
        float x = p2x[i], y = p2y[i];
        for (int j = 0; j < 4; j++) {
            float normalizationScale = vec1x[j] * vec1x[j] + vec1y[j] * vec1y[j];
            normalizationScale = (normalizationScale < 1e-12f) ? 1.f : 1.f / normalizationScale;
            float A = -vec1y[j] * normalizationScale;
            float B = vec1x[j] * normalizationScale;
            float C = -(A * p1x[j] + B * p1y[j]);
            float s = A * x + B * y + C;
            if (s >= 0) posSign++; else negSign++;
        }
        if (posSign == 4 || negSign == 4) { ix[n] = p2x[i]; iy[n] = p2y[i]; n++; }
    }
    int N = n;
    if (N == 0) { *n_out = 0; return FPM_INTERSECT_NONE; }

    // get rid of duplicated points
    const int Nstride = N;
    float distPt[24 * 24];
    int ptDistRemap[24];
    for (int i = 0; i < N; ++i) {
        float pt0x = ix[i], pt0y = iy[i];
        ptDistRemap[i] = i;
        for (int j = i + 1; j < N;) {
            float dx = ix[j] - pt0x, dy = iy[j] - pt0y;
            float d2 = dx * dx + dy * dy;
            if (d2 <= samePointEps) {
                if (j < N - 1) { ix[j] = ix[N - 1]; iy[j] = iy[N - 1]; }
                N--;
                continue;
            }
            distPt[i * Nstride + j] = d2;
            ++j;
        }
    }
    while (N > 8) {   // still duplicates after the eps threshold: eliminate closest points
        int minI = 0, minJ = 1;
        float minD = distPt[1];
        for (int i = 0; i < N - 1; ++i) {
            const float* pDist = distPt + Nstride * ptDistRemap[i];
            for (int j = i + 1; j < N; ++j) {
                float d = pDist[ptDistRemap[j]];
                if (d < minD) { minD = d; minI = i; minJ = j; }
            }
        }
        (void)minI;
        if (minJ < N - 1) { ix[minJ] = ix[N - 1]; iy[minJ] = iy[N - 1]; ptDistRemap[minJ] = ptDistRemap[N - 1]; }
        N--;
    }
    // order points
    for (int i = 0; i < N - 1; ++i) {
        float diffIx = ix[i + 1] - ix[i], diffIy = iy[i + 1] - iy[i];
        for (int j = i + 2; j < N; ++j) {
            float diffJx = ix[j] - ix[i], diffJy = iy[j] - iy[i];
            if (diffIx * diffJy - diffIy * diffJx < 0) {
                float tx = ix[i + 1], ty = iy[i + 1];
                ix[i + 1] = ix[j]; iy[i + 1] = iy[j];
                ix[j] = tx; iy[j] = ty;
                diffIx = diffJx; diffIy = diffJy;
            }
        }
    }
    *n_out = N;
    return ret;
}

// cv::rotatedRectangleIntersection: shift both rects to their mean centre first
FPM_HD int fpm_rrect_intersection(const FpmRRect& r1, const FpmRRect& r2, float* ix, float* iy, int* n_out)
{
    if (r1.w <= 0 || r1.h <= 0 || r2.w <= 0 || r2.h <= 0) { *n_out = 0; return FPM_INTERSECT_NONE; }
    float acx = (r1.cx + r2.cx) / 2.0f, acy = (r1.cy + r2.cy) / 2.0f;
    FpmRRect s1 = r1, s2 = r2;
    s1.cx -= acx; s1.cy -= acy;
    s2.cx -= acx; s2.cy -= acy;
    int ret = fpm_rrect_intersection_shifted(s1, s2, ix, iy, n_out);
    if (ret != FPM_INTERSECT_NONE)
        for (int i = 0; i < *n_out; ++i) { ix[i] += acx; iy[i] += acy; }
    else
        *n_out = 0;
    return ret;
}

// sortPtWithCenter, src/TemplateMatcher.cpp:1093-1131.  Quirks kept: the "norm" is the squared
// length (:1108), the same-Y branch compares vec1.x - ptCenter.x (:1121).  A NaN key (acos
// domain error) never compares less, so it stays where the stable insertion sort meets it.
FPM_HD void fpm_sort_pt_with_center(float* px, float* py, int n)
{
    float cx = 0, cy = 0;
    for (int i = 0; i < n; i++) { cx += px[i]; cy += py[i]; }
    cx /= n; cy /= n;
    double key[24];
    for (int i = 0; i < n; i++) {
        float vx = px[i] - cx, vy = py[i] - cy;
        float fNorm = vx * vx + vy * vy;
        float fDot = vx;
        if (vy < 0) key[i] = acos((double)(fDot / fNorm)) * FPM_R2D;
        else if (vy > 0) key[i] = 360 - acos((double)(fDot / fNorm)) * FPM_R2D;
        else key[i] = (vx - cx > 0) ? 0 : 180;
    }
    for (int i = 1; i < n; i++) {
        double k = key[i]; float x = px[i], y = py[i];
        int j = i - 1;
        while (j >= 0 && k < key[j]) { key[j + 1] = key[j]; px[j + 1] = px[j]; py[j + 1] = py[j]; j--; }
        key[j + 1] = k; px[j + 1] = x; py[j + 1] = y;
    }
}

// cv::contourArea (float points, not oriented)
FPM_HD double fpm_contour_area(const float* px, const float* py, int n)
{
    if (n == 0) return 0;
    double a00 = 0;
    float prevx = px[n - 1], prevy = py[n - 1];
    for (int i = 0; i < n; i++) {
        a00 += (double)prevx * py[i] - (double)prevy * px[i];
        prevx = px[i]; prevy = py[i];
    }
    a00 *= 0.5;
    return fabs(a00);
}

// One pair decision of filterWithRotatedRect (:1147-1182): returns 1 when the lower-scored
// rect must be deleted.  ratio_out receives dArea / rect1.size.area() (or -1 when not computed).
FPM_HD int fpm_rrect_overlap_decision(const FpmRRect& r1, const FpmRRect& r2, double max_overlap,
                                      int* type_out, double* ratio_out)
{
    float ix[24], iy[24];
    int n = 0;
    int type = fpm_rrect_intersection(r1, r2, ix, iy, &n);
    if (type_out) *type_out = type;
    if (ratio_out) *ratio_out = -1;
    if (type == FPM_INTERSECT_NONE) return 0;
    if (type == FPM_INTERSECT_FULL) return 1;
    if (n < 3) return 0;
    fpm_sort_pt_with_center(ix, iy, n);
    double area = fpm_contour_area(ix, iy, n);
    double ratio = area / (double)(r1.w * r1.h);
    if (ratio_out) *ratio_out = ratio;
    return ratio > max_overlap ? 1 : 0;
}
