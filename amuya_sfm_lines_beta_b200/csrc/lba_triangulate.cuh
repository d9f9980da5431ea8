// Two-view DLT triangulation of one image point pair, fp64 with float32 in/out.
//
// Replaces, per endpoint, the reference's  X = cv2.triangulatePoints(P0, P1, pts_v0, pts_v1); X = X[:3] / X[3]
// (/root/reference/src/seq_utils.py:119-121 and :57-59).  OpenCV (pinned 4.5.1.48, requirements.txt:4) builds the
// 4x4 matrix  A[2j] = x_j P_j[2] - P_j[0],  A[2j+1] = y_j P_j[2] - P_j[1]  in double precision, takes cv::SVD(A)
// (one-sided Jacobi) and returns the right singular vector of the smallest singular value converted to the
// depth of the points (float32 in the reference); the reference then divides in float32.  This file restates
// that: Hestenes' one-sided Jacobi on the columns of A with the rotations accumulated in V.
//
// Plain C++ as well as CUDA: the same source is compiled for the host by tests/test_triangulation.py to
// check the arithmetic without a GPU (test infrastructure; the library only ever runs it on the device).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LBA_HD __host__ __device__ __forceinline__
#else
#define LBA_HD inline
#endif

namespace lba {

// P0, P1: row-major 3x4.  x0, x1: pixel coordinates (float32).  Xh: homogeneous point (float32, unit 4-vector up to
// sign), X: Xh[:3] / Xh[3] evaluated in float32.  Returns the number of Jacobi sweeps used.
LBA_HD int triangulate_dlt(const double* P0, const double* P1, float x0, float y0, float x1, float y1, float Xh[4],
                           float X[3]) {
  // At[i][k] = A[k][i]: row i of At is column i of A
  double At[4][4], Vt[4][4], W[4];
  const double xs[4] = {(double)x0, (double)y0, (double)x1, (double)y1};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    At[i][0] = xs[0] * P0[8 + i] - P0[i];
    At[i][1] = xs[1] * P0[8 + i] - P0[4 + i];
    At[i][2] = xs[2] * P1[8 + i] - P1[i];
    At[i][3] = xs[3] * P1[8 + i] - P1[4 + i];
#pragma unroll
    for (int k = 0; k < 4; ++k) Vt[i][k] = i == k ? 1.0 : 0.0;
    W[i] = At[i][0] * At[i][0] + At[i][1] * At[i][1] + At[i][2] * At[i][2] + At[i][3] * At[i][3];
  }
  const double eps = 2.220446049250313e-16 * 10.0;
  int sweeps = 0;
  for (; sweeps < 30; ++sweeps) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        double a = W[i], b = W[j];
        double p = At[i][0] * At[j][0] + At[i][1] * At[j][1] + At[i][2] * At[j][2] + At[i][3] * At[j][3];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2.0;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0.0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2.0);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2.0));
          s = p / (gamma * c * 2.0);
        }
        a = 0.0; b = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double t0 = c * At[i][k] + s * At[j][k];
          const double t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          a += t0 * t0; b += t1 * t1;
          const double v0 = c * Vt[i][k] + s * Vt[j][k];
          const double v1 = -s * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = v0; Vt[j][k] = v1;
        }
        W[i] = a; W[j] = b;
        changed = true;
      }
    }
    if (!changed) break;
  }
  // the column of A with the smallest norm carries the null direction
  int m = 0;
  double wm = W[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (W[i] < wm) { wm = W[i]; m = i; }
  double v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // select without dynamic indexing (keeps Vt in registers on the device)
    v[k] = m == 0 ? Vt[0][k] : (m == 1 ? Vt[1][k] : (m == 2 ? Vt[2][k] : Vt[3][k]));
    Xh[k] = (float)v[k];
  }
  X[0] = Xh[0] / Xh[3];
  X[1] = Xh[1] / Xh[3];
  X[2] = Xh[2] / Xh[3];
  return sweeps;
}

}  // namespace lba
