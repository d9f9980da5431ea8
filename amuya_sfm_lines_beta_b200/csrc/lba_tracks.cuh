// Epipolar verification of one 2D line match, fp64.
//
// Restates the body of verify_duplet (/root/reference/src/utils.py:1877-1911): the endpoints a0, a1 of the
// segment in view v give the epipolar lines F a (:1886-1887); they are intersected with the supporting line of
// the matched segment b0-b1 of view v+1 (line_parametrization :1250, intersection :1279); the match is an inlier
// when one endpoint of b is within threshold * min(|I1 - I0|, |b1 - b0|) of "its" intersection (:1909) and the
// endpoint order is consistent (:1911).  A segment parallel to an epipolar line to within kParallelTol is an
// outlier (the reference raises on an exactly zero determinant and divides by rounding noise next to it).
//
// Plain C++ as well as CUDA: tests compile this header for the host to check the arithmetic without a GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LBA_TRK_HD __host__ __device__ __forceinline__
#else
#define LBA_TRK_HD inline
#endif

namespace lba {

constexpr double kParallelTol = 1e-12;

// F row-major 3x3; a0, a1, b0, b1 pixel coordinates (x, y)
LBA_TRK_HD bool duplet_inlier(const double* F, const double* a0, const double* a1, const double* b0, const double* b1,
                              double threshold) {
  const double A = b0[1] - b1[1];
  const double B = b1[0] - b0[0];
  const double Cc = b0[0] * b1[1] - b1[0] * b0[1];
  const double nAB = hypot(A, B);
  double ix[2], iy[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double* a = e ? a1 : a0;
    const double l0 = F[0] * a[0] + F[1] * a[1] + F[2];
    const double l1 = F[3] * a[0] + F[4] * a[1] + F[5];
    const double l2 = F[6] * a[0] + F[7] * a[1] + F[8];
    const double D = A * l1 - B * l0;
    const double Dx = -Cc * l1 + B * l2;
    const double Dy = -A * l2 + Cc * l0;
    if (!(fabs(D) > kParallelTol * nAB * hypot(l0, l1))) return false;
    ix[e] = Dx / D;
    iy[e] = Dy / D;
  }
  const double len_line = hypot(ix[1] - ix[0], iy[1] - iy[0]);
  const double len_corr = hypot(b1[0] - b0[0], b1[1] - b0[1]);
  const double thr = threshold * (len_line < len_corr ? len_line : len_corr);
  const double d11 = hypot(b0[0] - ix[0], b0[1] - iy[0]);
  const double d22 = hypot(b1[0] - ix[1], b1[1] - iy[1]);
  const double d12 = hypot(b0[0] - ix[1], b0[1] - iy[1]);
  const double d21 = hypot(b1[0] - ix[0], b1[1] - iy[0]);
  if (d11 < thr || d22 < thr)
    if ((d11 <= d22 && d22 < d21) || (d22 < d11 && d11 < d12)) return true;
  return false;
}

}  // namespace lba
