// The 12-parameter camera rows of the reference: camera_params (Nc,12) reshaped to (3,4),
// R = [:, :3] -- nine FREE entries, not a rotation -- and c = [:, 3]  (/root/reference/src/bundle_adjustment.py:43-45,
// produced by input_bundle_adj(rotate_model != "Rodrigues"), src/seq_utils.py:168-169,182-184).  No shipped driver
// uses this mode and its output cannot even be consumed by update_structure_motion, so the problems are the
// reference's toy sizes: the solver for it is the plain one -- analytic residual / Jacobian rows, the FULL normal
// equations H = J^T J assembled densely without atomics (camera blocks by camera, line blocks by line, the 12x6
// coupling block of every observation written directly), Marquardt damping, a dense Cholesky solve (cuSOLVER potrf /
// potrs, loaded on first use) and the same step policy as the main path (lba_policy.cuh, evaluated on the host
// here).  The Schur / PCG machinery of the 6-parameter path is not generalised to 12x12 blocks.
//
//   x = M (X - c),  p = x/|x|,  e = lamb^2 (p.n)^2 + (p.m)^2,  u = de/dx = (g - (g.p) p)/|x|,  g = 2 lamb^2 (p.n) n + 2 (p.m) m
//   de/dM_kj = u_k d_j,  de/dX = M^T u,  de/dc = -M^T u          (row layout [M_k0 M_k1 M_k2 c_k], k = 0..2)
#pragma once
#include "lba_kernels.cuh"

namespace lba {

struct Eval12Args {
  int Mp, ntiles, nc, nl;
  const TileIdx* tidx;
  const double* cst;      // [ntiles][256][9]: n, m0 = n x p0, m1 = n x p1 per slot
  const double* x;        // [12 nc + 6 nl]: camera rows (3x4 row-major), then [X0, X1] per line
  double* r;              // [Mp]
  double* jc;             // [Mp][12] in the row layout of the camera parameters (NULL: residuals only)
  double* jl;             // [Mp][6]
  double* partial; unsigned int* ticket; double* cost_out;
  int loss; double f_scale;
};

template <bool WITH_JAC>
__global__ void __launch_bounds__(kTile) k_eval12(const Eval12Args a) {
  pdl_prologue();
  __shared__ double red[kTile / 32];
  const int tile = blockIdx.x, t = threadIdx.x;
  const TileIdx& ti = a.tidx[tile];
  const int cam = ti.cam[t];
  const size_t i = (size_t)tile * kTile + t;
  double r = 0.0, rho_cost = 0.0;
  double jc[12], jl[6];
#pragma unroll
  for (int k = 0; k < 12; ++k) jc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) jl[k] = 0.0;
  if (cam >= 0) {
    const double* cst = a.cst + ((size_t)tile * kTile + t) * 9;
    double n[3], m[2][3], M[9], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { n[k] = cst[k]; m[0][k] = cst[3 + k]; m[1][k] = cst[6 + k]; }
    const double* cp = a.x + (size_t)cam * 12;
#pragma unroll
    for (int k = 0; k < 3; ++k) { M[3 * k] = cp[4 * k]; M[3 * k + 1] = cp[4 * k + 1]; M[3 * k + 2] = cp[4 * k + 2]; c[k] = cp[4 * k + 3]; }
    const double* X = a.x + (size_t)a.nc * 12 + (size_t)(ti.meta.l0 + ti.ll[t]) * 6;
    const double l2 = kLamb * kLamb;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double d0 = X[3 * e] - c[0], d1 = X[3 * e + 1] - c[1], d2 = X[3 * e + 2] - c[2];
      const double x0 = M[0] * d0 + M[1] * d1 + M[2] * d2;
      const double x1 = M[3] * d0 + M[4] * d1 + M[5] * d2;
      const double x2 = M[6] * d0 + M[7] * d1 + M[8] * d2;
      const double inv_rho = 1.0 / sqrt(x0 * x0 + x1 * x1 + x2 * x2);
      const double q0 = x0 * inv_rho, q1 = x1 * inv_rho, q2 = x2 * inv_rho;
      const double av = q0 * n[0] + q1 * n[1] + q2 * n[2];
      const double bv = q0 * m[e][0] + q1 * m[e][1] + q2 * m[e][2];
      r += l2 * av * av + bv * bv;
      if (WITH_JAC) {
        const double ga = 2.0 * l2 * av, gb = 2.0 * bv;
        const double g0 = ga * n[0] + gb * m[e][0], g1 = ga * n[1] + gb * m[e][1], g2 = ga * n[2] + gb * m[e][2];
        const double gp = g0 * q0 + g1 * q1 + g2 * q2;
        const double u[3] = {(g0 - gp * q0) * inv_rho, (g1 - gp * q1) * inv_rho, (g2 - gp * q2) * inv_rho};
        const double d[3] = {d0, d1, d2};
        double je[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) je[j] = M[j] * u[0] + M[3 + j] * u[1] + M[6 + j] * u[2];   // M^T u
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
          for (int j = 0; j < 3; ++j) jc[4 * k + j] += u[k] * d[j];
          jc[4 * k + 3] -= je[k];
          jl[3 * e + k] = je[k];
        }
      }
    }
    if (a.loss) {   // scipy's robust-loss handling, as in k_eval
      const double rs = r / a.f_scale, z = rs * rs;
      double r0, r1, r2;
      robust_rho(a.loss, z, r0, r1, r2);
      rho_cost = a.f_scale * a.f_scale * r0;
      if (WITH_JAC) {
        double js = r1 + 2.0 * r2 * z;
        js = sqrt(js < 2.220446049250313e-16 ? 2.220446049250313e-16 : js);
        r = r * r1 / js;
#pragma unroll
        for (int k = 0; k < 12; ++k) jc[k] *= js;
#pragma unroll
        for (int k = 0; k < 6; ++k) jl[k] *= js;
      }
    } else {
      rho_cost = r * r;
    }
  }
  a.r[i] = r;
  if (WITH_JAC) {
#pragma unroll
    for (int k = 0; k < 12; ++k) a.jc[i * 12 + k] = jc[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) a.jl[i * 6 + k] = jl[k];
  }
  const double cs[1] = {block_sum<kTile>(rho_cost, red)};
  last_cta_reduce<kTile, 1>(cs, a.partial, a.ticket, a.cost_out, 0.5, red);
}

// H (n x n, column-major == row-major: symmetric, both triangles written) and g = J^T r.
// Camera blocks: one CTA per camera, thread e owns one entry of the 12x12 block (or of g_c) and adds the camera's
// observations in camera order.
__global__ void __launch_bounds__(192) k_h12_cam(int n, const int* __restrict__ cam_ptr, const int* __restrict__ cam_perm,
                                                 const double* __restrict__ jc, const double* __restrict__ r,
                                                 double* __restrict__ H, double* __restrict__ g) {
  pdl_prologue();
  const int c = blockIdx.x, e = threadIdx.x;
  if (e >= 156) return;
  const int j0 = cam_ptr[c], j1 = cam_ptr[c + 1];
  double acc = 0.0;
  if (e < 144) {
    const int a = e / 12, b = e % 12;
    for (int j = j0; j < j1; ++j) { const double* row = jc + (size_t)cam_perm[j] * 12; acc += row[a] * row[b]; }
    H[(size_t)(12 * c + a) * n + 12 * c + b] = acc;
  } else {
    const int a = e - 144;
    for (int j = j0; j < j1; ++j) { const int s = cam_perm[j]; acc += jc[(size_t)s * 12 + a] * r[s]; }
    g[12 * c + a] = acc;
  }
}
// Line blocks: one thread per (line, entry); the observations of a line are consecutive slots.
__global__ void k_h12_line(int n, int nc, int nl, const int* __restrict__ line_slot, const int* __restrict__ line_cnt,
                           const double* __restrict__ jl, const double* __restrict__ r, double* __restrict__ H,
                           double* __restrict__ g) {
  pdl_prologue();
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)nl * 42) return;
  const int l = (int)(q / 42), e = (int)(q % 42);
  const int s0 = line_slot[l], s1 = s0 + line_cnt[l];
  const int off = 12 * nc + 6 * l;
  double acc = 0.0;
  if (e < 36) {
    const int a = e / 6, b = e % 6;
    for (int s = s0; s < s1; ++s) acc += jl[(size_t)s * 6 + a] * jl[(size_t)s * 6 + b];
    H[(size_t)(off + a) * n + off + b] = acc;
  } else {
    const int a = e - 36;
    for (int s = s0; s < s1; ++s) acc += jl[(size_t)s * 6 + a] * r[s];
    g[off + a] = acc;
  }
}
// Coupling blocks: every (camera, line) pair is observed at most once, so the 12 x 6 block is one outer product
// (a pair observed twice would need a sum: `dup` counts such slots and the host refuses the problem).
__global__ void k_h12_cross(int Mp, int n, int nc, const TileIdx* __restrict__ tidx, const double* __restrict__ jc,
                            const double* __restrict__ jl, double* __restrict__ H) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const TileIdx& ti = tidx[i / kTile];
  const int cam = ti.cam[i % kTile];
  if (cam < 0) return;
  const int off = 12 * nc + 6 * (ti.meta.l0 + ti.ll[i % kTile]);
#pragma unroll
  for (int a = 0; a < 12; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const double v = jc[(size_t)i * 12 + a] * jl[(size_t)i * 6 + b];
      H[(size_t)(12 * cam + a) * n + off + b] = v;
      H[(size_t)(off + b) * n + 12 * cam + a] = v;
    }
}
// Marquardt scaling D = running max of diag(H) (zeros -> 1) and the damped copy A = H + lam diag(D)
__global__ void k_damp12(int n, const double* __restrict__ H, double* __restrict__ D, int first, double lam,
                         double* __restrict__ A, int update_scaling) {
  pdl_prologue();
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)n * n) return;
  const int row = (int)(q / n), col = (int)(q % n);
  double v = H[q];
  if (row == col) {
    double d = D[row];
    if (update_scaling) {
      double h = v;
      if (h == 0.0) h = 1.0;
      d = first ? h : fmax(d, h);
      D[row] = d;
    }
    v += lam * d;
  }
  A[q] = v;
}
// rhs = -g
__global__ void k_neg12(int n, const double* __restrict__ g, double* __restrict__ rhs) {
  pdl_prologue();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) rhs[q] = -g[q];
}
// (J delta)_i, r.(J delta), |J delta|^2  (sums[0..1]); single CTA sums |delta|^2, |x|^2, |g|_inf afterwards
__global__ void __launch_bounds__(kTile) k_jd12(int Mp, int nc, const TileIdx* __restrict__ tidx, const double* __restrict__ jc,
                                                const double* __restrict__ jl, const double* __restrict__ r,
                                                const double* __restrict__ delta, double* __restrict__ partial,
                                                unsigned int* ticket, double* __restrict__ out2) {
  pdl_prologue();
  __shared__ double red[kTile / 32];
  const int tile = blockIdx.x, t = threadIdx.x;
  const TileIdx& ti = tidx[tile];
  const int cam = ti.cam[t];
  const size_t i = (size_t)tile * kTile + t;
  double p0 = 0.0, p1 = 0.0;
  if (cam >= 0) {
    const double* dc = delta + (size_t)cam * 12;
    const double* dl = delta + (size_t)nc * 12 + (size_t)(ti.meta.l0 + ti.ll[t]) * 6;
    double jd = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) jd += jc[i * 12 + k] * dc[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) jd += jl[i * 6 + k] * dl[k];
    p0 = r[i] * jd; p1 = jd * jd;
  }
  const double sums[2] = {block_sum<kTile>(p0, red), block_sum<kTile>(p1, red)};
  last_cta_reduce<kTile, 2>(sums, partial, ticket, out2, 1.0, red);
}
__global__ void __launch_bounds__(kTile) k_norms12(int n, const double* __restrict__ delta, const double* __restrict__ x,
                                                   const double* __restrict__ g, double* __restrict__ out3) {
  pdl_prologue();
  __shared__ double red[kTile / 32];
  double d2 = 0.0, x2 = 0.0, gm = 0.0;
  for (int q = threadIdx.x; q < n; q += kTile) {
    d2 += delta[q] * delta[q]; x2 += x[q] * x[q];
    const double ab = fabs(g[q]);
    gm = (ab > gm || ab != ab) ? ab : gm;
  }
  const double s0 = block_sum<kTile>(d2, red), s1 = block_sum<kTile>(x2, red);
  __shared__ double gmax[kTile];
  gmax[threadIdx.x] = gm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int q = 0; q < kTile; ++q) m = (gmax[q] > m || gmax[q] != gmax[q]) ? gmax[q] : m;
    out3[0] = s0; out3[1] = s1; out3[2] = m;
  }
}
// x_new = x + alpha delta, rounded like the oracle (no FMA)
__global__ void k_step12(int n, const double* __restrict__ x, const double* __restrict__ delta, double alpha,
                         double* __restrict__ x_new) {
  pdl_prologue();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) x_new[q] = __dadd_rn(x[q], __dmul_rn(alpha, delta[q]));
}

// project() with [R|c] rows (bundle_adjustment.py:43-63): one thread per row
__global__ void k_project12(long long N, const double* __restrict__ pts, const double* __restrict__ cams, double* __restrict__ out) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* cp = cams + i * 12;
  const double d0 = pts[i * 3] - cp[3], d1 = pts[i * 3 + 1] - cp[7], d2 = pts[i * 3 + 2] - cp[11];
  const double x0 = cp[0] * d0 + cp[1] * d1 + cp[2] * d2;
  const double x1 = cp[4] * d0 + cp[5] * d1 + cp[6] * d2;
  const double x2 = cp[8] * d0 + cp[9] * d1 + cp[10] * d2;
  const double inv = 1.0 / sqrt(x0 * x0 + x1 * x1 + x2 * x2);
  out[i * 3] = x0 * inv; out[i * 3 + 1] = x1 * inv; out[i * 3 + 2] = x2 * inv;
}

}  // namespace lba
