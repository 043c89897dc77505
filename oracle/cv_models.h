// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Declarations of oracle/cv_models.cpp (C++ models of the OpenCV calls on the hot path).
#pragma once
#include <cstdint>

namespace cvm {
int cv_round(double v);
void pyr_down(const uint8_t* s, int w, int h, int sp, uint8_t* d, int dp);
void rotation_matrix(double cx, double cy, double angle_deg, double* m6);
void warp_affine(const uint8_t* src, int sw, int sh, int sp, const double* m6_forward, uint8_t* dst, int dw, int dh, int dp, int border);
void integral(const uint8_t* s, int w, int h, int sp, double* sum, double* sqsum);     // (h+1) x (w+1) each
void ccorr_exact(const uint8_t* img, int iw, int ih, int ip, const uint8_t* tpl, int tw, int th, int tp, float* out);
void mean_stddev(const uint8_t* s, int w, int h, int sp, double* mean, double* sdv);
int rrect_overlap(const float a[5], const float b[5], double max_overlap);             // 1: the lower-scored rect must go
void rrect_from3(const float p[6], float out[5]);
bool lu_inverse(double* A, double* B, int m);
}  // namespace cvm
