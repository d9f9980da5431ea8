// Device-side geometry for the line bundle-adjustment hot path (sm_100a, fp64).
//
// Restates, per observation, what the reference computes in Python:
//   project()  /root/reference/src/bundle_adjustment.py:23-74   x = R(w) (X - c), p_hat = x/|x|
//   fun()      /root/reference/src/bundle_adjustment.py:118-135 r = sum_e (lamb p_hat.n)^2 + (p_hat.(n x p_e))^2
//   cv2.Rodrigues (call site :48)                               R = cos t I + (1-cos t) r r^T + sin t [r]x
// plus the analytic Jacobian that replaces scipy's finite differences.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lba {

constexpr double kLamb = 2.0;  // bundle_adjustment.py:120

// Compact per-camera record gathered by the observation kernels: 80 B = 5 x 16 B.
// R and the SO(3) right Jacobian G are rebuilt from (w, A, B, C) in registers
// (40 DFMA) instead of gathering 18 more doubles per observation.
struct __align__(16) CamRec {
  double w[3];
  double c[3];
  double A;  // sin(t)/t
  double B;  // (1-cos t)/t^2
  double C;  // (t-sin t)/t^3
  double pad;
};

__host__ __device__ inline void rodrigues_coeffs(const double w[3], double& A, double& B, double& C) {
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double t = sqrt(t2);
  if (t < 0.1) {
    A = 1.0 + t2 * (-1.0 / 6 + t2 * (1.0 / 120 + t2 * (-1.0 / 5040 + t2 * (1.0 / 362880 + t2 * (-1.0 / 39916800)))));
    B = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800))));
    C = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800))));
  } else {
    const double sh = sin(0.5 * t);
    A = sin(t) / t;
    B = 2.0 * sh * sh / t2;
    C = (t - sin(t)) / (t2 * t);
  }
}

// C = (t - sin t) / t^3 from t^2 and A = sin t / t: what the Jacobian kernel computes per observation instead of
// gathering C with the camera record (the series of rodrigues_coeffs below t = 0.1; above it (1 - A) / t^2 loses
// at most log2(6 / t^2) bits to cancellation, 6e-14 relative at t = 0.1, like (t - sin t) / t^3 itself)
__host__ __device__ inline double coeff_C(double t2, double A) {
  const double ser = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800))));
  return t2 < 0.01 ? ser : (1.0 - A) / t2;
}

// R = I + A [w]x + B [w]x^2,  [w]x^2 = w w^T - t2 I   (row-major 3x3)
__host__ __device__ inline void build_R(const double w[3], double A, double B, double R[9]) {
  const double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  const double xy = B * w[0] * w[1], xz = B * w[0] * w[2], yz = B * w[1] * w[2];
  R[0] = 1.0 - B * (yy + zz); R[1] = xy - A * w[2];       R[2] = xz + A * w[1];
  R[3] = xy + A * w[2];       R[4] = 1.0 - B * (xx + zz); R[5] = yz - A * w[0];
  R[6] = xz - A * w[1];       R[7] = yz + A * w[0];       R[8] = 1.0 - B * (xx + yy);
}

// G = I - B [w]x + C [w]x^2   with d(R d)/dw = -R [d]x G
__host__ __device__ inline void build_G(const double w[3], double B, double C, double G[9]) {
  const double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  const double xy = C * w[0] * w[1], xz = C * w[0] * w[2], yz = C * w[1] * w[2];
  G[0] = 1.0 - C * (yy + zz); G[1] = xy + B * w[2];       G[2] = xz - B * w[1];
  G[3] = xy - B * w[2];       G[4] = 1.0 - C * (xx + zz); G[5] = yz + B * w[0];
  G[6] = xz + B * w[1];       G[7] = yz - B * w[0];       G[8] = 1.0 - C * (xx + yy);
}

// One endpoint of one observation.  Returns e = lamb^2 a^2 + b^2; when WITH_JAC also
// je = de/dX = R^T u and accumulates tq += d x je (so that de/dw = G^T tq, de/dc = -je).
template <bool WITH_JAC>
__device__ __forceinline__ double endpoint_term(const double R[9], const double c[3], const double X[3],
                                                const double n[3], const double me[3], double je[3],
                                                double tq[3]) {
  const double d0 = X[0] - c[0], d1 = X[1] - c[1], d2 = X[2] - c[2];
  const double x0 = R[0] * d0 + R[1] * d1 + R[2] * d2;
  const double x1 = R[3] * d0 + R[4] * d1 + R[5] * d2;
  const double x2 = R[6] * d0 + R[7] * d1 + R[8] * d2;
  const double rho2 = x0 * x0 + x1 * x1 + x2 * x2;
#ifdef __CUDA_ARCH__
  const double inv_rho = rsqrt(rho2);   // one reciprocal-square-root sequence (1 ulp) instead of a square root and a division
#else
  const double inv_rho = 1.0 / sqrt(rho2);
#endif
  const double q0 = x0 * inv_rho, q1 = x1 * inv_rho, q2 = x2 * inv_rho;
  const double a = q0 * n[0] + q1 * n[1] + q2 * n[2];
  const double b = q0 * me[0] + q1 * me[1] + q2 * me[2];
  const double l2 = kLamb * kLamb;
  const double e = l2 * a * a + b * b;
  if (WITH_JAC) {
    const double ga = 2.0 * l2 * a, gb = 2.0 * b;
    const double g0 = ga * n[0] + gb * me[0];
    const double g1 = ga * n[1] + gb * me[1];
    const double g2 = ga * n[2] + gb * me[2];
    const double gp = g0 * q0 + g1 * q1 + g2 * q2;
    const double u0 = (g0 - gp * q0) * inv_rho;
    const double u1 = (g1 - gp * q1) * inv_rho;
    const double u2 = (g2 - gp * q2) * inv_rho;
    je[0] = R[0] * u0 + R[3] * u1 + R[6] * u2;
    je[1] = R[1] * u0 + R[4] * u1 + R[7] * u2;
    je[2] = R[2] * u0 + R[5] * u1 + R[8] * u2;
    tq[0] += d1 * je[2] - d2 * je[1];
    tq[1] += d2 * je[0] - d0 * je[2];
    tq[2] += d0 * je[1] - d1 * je[0];
  }
  return e;
}

// scipy.optimize.least_squares robust losses (scipy/optimize/_lsq/least_squares.py: soft_l1, huber,
// cauchy, arctan): rho(z), rho'(z), rho''(z) for z = (r/f_scale)^2.  loss: 1 soft_l1, 2 huber, 3 cauchy, 4 arctan.
__host__ __device__ inline void robust_rho(int loss, double z, double& r0, double& r1, double& r2) {
  if (loss == 1) {
    const double t = 1.0 + z, st = sqrt(t);
    r0 = 2.0 * (st - 1.0); r1 = 1.0 / st; r2 = -0.5 / (t * st);
  } else if (loss == 2) {
    if (z <= 1.0) { r0 = z; r1 = 1.0; r2 = 0.0; }
    else { const double sz = sqrt(z); r0 = 2.0 * sz - 1.0; r1 = 1.0 / sz; r2 = -0.5 / (z * sz); }
  } else if (loss == 3) {
    const double t = 1.0 + z;
    r0 = log1p(z); r1 = 1.0 / t; r2 = -1.0 / (t * t);
  } else {
    const double t = 1.0 + z * z;
    r0 = atan(z); r1 = 1.0 / t; r2 = -2.0 * z / (t * t);
  }
}

// Packed symmetric 6x6: 21 doubles, upper triangle row-major: (0,0)(0,1)..(0,5)(1,1)..(5,5).
__host__ __device__ constexpr int sym_idx(int a, int b) {
  return a <= b ? (a * (13 - a)) / 2 + (b - a) : (b * (13 - b)) / 2 + (a - b);
}

// In-place inverse of a packed SPD 6x6 via Cholesky (L L^T), returns the number of
// pivots that had to be clamped (0 for a healthy block).
__host__ __device__ inline int spd6_inverse(double S[21]) {
  double L[21];  // lower triangle, row-major packed: L[i][j] at i*(i+1)/2 + j
  int clamped = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = S[sym_idx(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      if (i == j) {
        if (!(s > 0.0)) { s = 1e-300; ++clamped; }
        L[i * (i + 1) / 2 + i] = sqrt(s);
      } else {
        L[i * (i + 1) / 2 + j] = s / L[j * (j + 1) / 2 + j];
      }
    }
  }
  // Linv (lower), in place
  double Li[21];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    Li[i * (i + 1) / 2 + i] = 1.0 / L[i * (i + 1) / 2 + i];
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) s -= L[i * (i + 1) / 2 + k] * Li[k * (k + 1) / 2 + j];
      Li[i * (i + 1) / 2 + j] = s * Li[i * (i + 1) / 2 + i];
    }
  }
  // S^-1 = Linv^T Linv
#pragma unroll
  for (int a = 0; a < 6; ++a) {
#pragma unroll
    for (int b = a; b < 6; ++b) {
      double s = 0.0;
#pragma unroll
      for (int k = b; k < 6; ++k) s += Li[k * (k + 1) / 2 + a] * Li[k * (k + 1) / 2 + b];
      S[sym_idx(a, b)] = s;
    }
  }
  return clamped;
}

// y = S x for packed symmetric S
__host__ __device__ inline void sym6_matvec(const double S[21], const double x[6], double y[6]) {
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) s += S[sym_idx(a, b)] * x[b];
    y[a] = s;
  }
}

}  // namespace lba
