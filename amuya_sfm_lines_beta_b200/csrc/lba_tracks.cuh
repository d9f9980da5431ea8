// Epipolar verification of one 2D line match, fp64.
//
// Restates the body of verify_duplet (/root/reference/src/utils.py:1877-1911): the endpoints a0, a1 of the
// segment in view v give the epipolar lines F a (:1886-1887); they are intersected with the supporting line of
// the matched segment b0-b1 of view v+1 (line_parametrization :1250, intersection :1279); the match is an inlier
// when one endpoint of b is within threshold * min(|I1 - I0|, |b1 - b0|) of "its" intersection (:1909) and the
// endpoint order is consistent (:1911).
//
// The reference's decision on a segment that is (nearly) parallel to an epipolar line is pure rounding: it divides
// by a determinant of a few ulps.  To take the SAME decisions, every operation is rounded exactly as numpy rounds
// it there: scalar products and sums one at a time (no FMA contraction), np.linalg.norm of a 2-vector as
// sqrt(x*x + y*y), and np.dot(F, x) (:1886) as the BLAS kernel evaluates a 3-term row,
// fma(F_k0, x, F_k1 * y) + F_k2 -- the order pinned by the recorded reference flags of all three toy domains
// (tests/golden/trk_*.npz: 2240 matches, 88 of them within 1e-9 of parallel).  Only an exactly zero determinant,
// on which the reference raises (intersection returns False, :1890), is an outlier by rule; `parallel_tol` > 0
// (an opt-in, not the default) additionally drops matches within that relative distance of parallel.
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

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double trk_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double trk_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double trk_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
// host build (tests): compile with -ffp-contract=off
inline double trk_mul(double a, double b) { return a * b; }
inline double trk_add(double a, double b) { return a + b; }
inline double trk_fma(double a, double b, double c) { return fma(a, b, c); }
#endif
LBA_TRK_HD double trk_norm2(double dx, double dy) { return sqrt(trk_add(trk_mul(dx, dx), trk_mul(dy, dy))); }

// F row-major 3x3; a0, a1, b0, b1 pixel coordinates (x, y)
LBA_TRK_HD bool duplet_inlier(const double* F, const double* a0, const double* a1, const double* b0, const double* b1,
                              double threshold, double parallel_tol = 0.0) {
  const double A = trk_add(b0[1], -b1[1]);
  const double B = trk_add(b1[0], -b0[0]);
  const double Cc = trk_add(trk_mul(b0[0], b1[1]), -trk_mul(b1[0], b0[1]));
  double ix[2], iy[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double* a = e ? a1 : a0;
    const double l0 = trk_add(trk_fma(F[0], a[0], trk_mul(F[1], a[1])), F[2]);
    const double l1 = trk_add(trk_fma(F[3], a[0], trk_mul(F[4], a[1])), F[5]);
    const double l2 = trk_add(trk_fma(F[6], a[0], trk_mul(F[7], a[1])), F[8]);
    const double D = trk_add(trk_mul(A, l1), -trk_mul(B, l0));
    const double Dx = trk_add(trk_mul(-Cc, l1), trk_mul(B, l2));
    const double Dy = trk_add(trk_mul(-A, l2), trk_mul(Cc, l0));
    if (D == 0.0) return false;
    if (parallel_tol > 0.0 && !(fabs(D) > parallel_tol * trk_norm2(A, B) * trk_norm2(l0, l1))) return false;
    ix[e] = Dx / D;
    iy[e] = Dy / D;
  }
  const double len_line = trk_norm2(trk_add(ix[1], -ix[0]), trk_add(iy[1], -iy[0]));
  const double len_corr = trk_norm2(trk_add(b1[0], -b0[0]), trk_add(b1[1], -b0[1]));
  const double thr = trk_mul(threshold, len_line < len_corr ? len_line : len_corr);
  const double d11 = trk_norm2(trk_add(b0[0], -ix[0]), trk_add(b0[1], -iy[0]));
  const double d22 = trk_norm2(trk_add(b1[0], -ix[1]), trk_add(b1[1], -iy[1]));
  const double d12 = trk_norm2(trk_add(b0[0], -ix[1]), trk_add(b0[1], -iy[1]));
  const double d21 = trk_norm2(trk_add(b1[0], -ix[0]), trk_add(b1[1], -iy[0]));
  if (d11 < thr || d22 < thr)
    if ((d11 <= d22 && d22 < d21) || (d22 < d11 && d11 < d12)) return true;
  return false;
}

}  // namespace lba
