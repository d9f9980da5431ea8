/* lba.h -- C ABI of the B200-native line bundle-adjustment library (liblba.so).
 *
 * Drop-in boundary for the hot path of bgpantojar/amuya_sfm_lines_beta:
 *   /root/reference/src/bundle_adjustment.py::run_bundle_adjustment (:254-349)
 * and the helpers it calls (project :23, fun :78, point_wise_input :194) plus the
 * scipy.optimize.least_squares call (:319) that this library replaces with a
 * Levenberg-Marquardt / Schur / PCG solver running in hand-written sm_100a kernels.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every data pointer may be a host pointer
 *    (pageable or pinned) or a device pointer; copies use cudaMemcpyDefault (UVA).
 *  - Layouts are the reference's own:
 *      camera_params  (Nc,6) f64 row-major, interleaved [w0,c0,w1,c1,w2,c2]
 *                     (seq_utils.py:179-181, bundle_adjustment.py:111)
 *      points_3d      (2*Nl,3) f64 row-major, row 2L+e = endpoint e of 3D line L
 *                     (bundle_adjustment.py:231-233,112)
 *      camera_indices (m,) int64, lines_indices (m,) int64   (seq_utils.py:194-195,203-204)
 *      n, p0, p1      (m,3) f64 row-major: Line.n and Line.points[e].p of each observed
 *                     2D line (domain.py:338-351; read at bundle_adjustment.py:119-130)
 *  - All arithmetic is fp64.  Every function returns LBA_OK (0) or a negative
 *    error code; lba_last_error() returns a thread-local message.
 *  - A context is bound to one CUDA device and one stream; calls are stream-ordered
 *    and synchronise only where a host scalar or host buffer is returned.
 *  - There is no CPU fallback: without a CUDA device lba_create fails.
 */
#ifndef LBA_H_
#define LBA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBA_OK 0
#define LBA_ERR_CUDA (-1)
#define LBA_ERR_ARG (-2)
#define LBA_ERR_STATE (-3)
#define LBA_ERR_UNSUPPORTED (-4)
#define LBA_ERR_NONFINITE (-5) /* "Residuals are not finite in the initial point." (scipy least_squares.py:949) */
#define LBA_ERR_COMM (-6)


typedef struct lba_ctx lba_ctx;

/* Options of lba_lm_solve.  Defaults (lba_default_options) mirror the reference call
 * least_squares(..., ftol=1e-4, method='trf', x_scale='jac') with scipy's default
 * xtol = gtol = 1e-8 (bundle_adjustment.py:319). */
typedef struct lba_options {
  int32_t max_iter;       /* LM iterations: damped solves, each followed by one trial evaluation (two when the
                             first is rejected and retried at a shorter length) */
  int32_t pcg_max_iter;   /* PCG iterations per solve */
  double ftol;            /* stop when dF < ftol*F and ratio > 0.25 (scipy check_termination) */
  double xtol;            /* stop when |step| < xtol*(xtol+|x|) */
  double gtol;            /* stop when |J^T r|_inf < gtol */
  double lambda0;         /* initial damping (Marquardt: J^T J + lambda*D) */
  double lambda_min;
  double lambda_max;
  double pcg_tol;         /* eta: stop PCG when |r|_{M^-1} <= eta |r0|_{M^-1} */
  double min_ratio;       /* accept a step when actual/predicted > min_ratio */
  int32_t verbose;        /* 2: print scipy-style iteration table to stdout */
  int32_t flags;          /* reserved (0) */
  int32_t loss;           /* LBA_LOSS_*: scipy least_squares `loss` (forwarded as loss_f at bundle_adjustment.py:319) */
  int32_t pad;
  double f_scale;         /* scipy `f_scale` (soft margin between inliers and outliers); 1.0 like the reference */
  int32_t max_evals;      /* > 0: stop after this many evaluations, the initial one included (scipy's max_nfev) */
  int32_t max_backtrack;  /* retries along a rejected step at the measured length before the damping is raised (1) */
} lba_options;

/* scipy.optimize.least_squares robust losses, rho(z) with z = (r/f_scale)^2 (scipy/optimize/_lsq/least_squares.py) */
#define LBA_LOSS_LINEAR 0
#define LBA_LOSS_SOFT_L1 1
#define LBA_LOSS_HUBER 2
#define LBA_LOSS_CAUCHY 3
#define LBA_LOSS_ARCTAN 4

#define LBA_TRACE_MAX 512
typedef struct lba_report {
  int32_t iterations;
  int32_t status;         /* 0 max_iter, 1 gtol, 2 ftol, 3 xtol, 4 ftol and xtol */
  int32_t n_jac;          /* residual+Jacobian evaluations (every trial point is evaluated with its Jacobian) */
  int32_t n_res;          /* residual-only evaluations (0) */
  int32_t pcg_iters_total;
  int32_t chol_clamped;   /* number of 6x6 pivots clamped (0 on healthy problems) */
  int32_t n_accepted;     /* accepted trial steps */
  int32_t pad;
  double cost_init;       /* 0.5 * sum r^2 at the input */
  double cost;            /* at the output */
  double seconds;         /* wall time inside lba_lm_solve */
  /* trace: one entry per trial evaluation (first LBA_TRACE_MAX), written by the device decision kernel */
  int32_t trace_len;
  int32_t trace_it[LBA_TRACE_MAX];       /* iteration (damped solve) the trial belongs to */
  int32_t trace_pcg[LBA_TRACE_MAX];
  int32_t trace_accepted[LBA_TRACE_MAX];
  double trace_cost[LBA_TRACE_MAX];      /* cost before the step */
  double trace_cost_new[LBA_TRACE_MAX];  /* cost at the trial point */
  double trace_lambda[LBA_TRACE_MAX];
  double trace_pred[LBA_TRACE_MAX];
  double trace_step_norm[LBA_TRACE_MAX];
  double trace_g_norm[LBA_TRACE_MAX];
  double trace_alpha[LBA_TRACE_MAX];     /* step length of the trial */
} lba_report;

/* Kernel classes for lba_get_profile */
enum {
  LBA_K_RESIDUAL = 0,   /* residual-only pass (trial cost, fun)            */
  LBA_K_JACOBIAN,       /* fused residual + Jacobian + line blocks V_l,g_l */
  LBA_K_CAM_BLOCKS,     /* per-camera U_c, g_c accumulation                */
  LBA_K_LINE_FACTOR,    /* (V_l + lambda D)^-1, V^-1 g_l                   */
  LBA_K_LINE_SCALARS,   /* per-observation Schur weights                   */
  LBA_K_CAM_PRECOND,    /* block-Jacobi blocks + Schur rhs                 */
  LBA_K_PCG_LINE,       /* S*p, line-order half                            */
  LBA_K_PCG_CAM,        /* S*p, camera-order half                          */
  LBA_K_PCG_UPDATE,     /* CG vector update                                */
  LBA_K_BACKSUB,        /* line update + model reduction                   */
  LBA_K_MISC,           /* small helpers (camera records, reductions, axpy) */
  LBA_K_COMM,           /* NCCL collectives                                */
  LBA_K_COUNT
};

const char* lba_last_error(void);
int lba_version(void);
void lba_default_options(lba_options* opts);

/* Create / destroy a context on CUDA device `device`. */
int lba_create(int device, lba_ctx** out);
int lba_destroy(lba_ctx* ctx);
/* Use an existing CUDA stream (cudaStream_t passed as void*); default is a private stream. */
int lba_set_stream(lba_ctx* ctx, void* cuda_stream);
int lba_synchronize(lba_ctx* ctx);
/* Device pointers handed to this library are read on the context's stream.  If they were produced by work still
 * in flight on ANOTHER stream (e.g. the caller's PyTorch stream), call this first: the context's streams then wait
 * for everything enqueued on `cuda_stream` so far.  Calls that return host data synchronise before returning. */
int lba_wait_stream(lba_ctx* ctx, void* cuda_stream);

/* Describe the observations (everything `fun` reads that is not a parameter).
 * Replaces point_wise_input (bundle_adjustment.py:194-250) and the per-call reads of
 * lines_2d[i].n / .points[e].p (:119-130).  Observations may come in any order; the
 * reference emits them sorted by (line, camera) (seq_utils.py:188-195).
 * Nc and Nl are the GLOBAL camera count and the number of 3D lines held by this context. */
int lba_set_problem(lba_ctx* ctx, int64_t m, int64_t n_cameras, int64_t n_lines,
                    const int64_t* camera_indices, const int64_t* lines_indices,
                    const double* n, const double* p0, const double* p1);

/* Integer indexing of point_wise_input (:235-243): camera_indices_pt[2i+e] = camera_indices[i],
 * pt_indices[2i+e] = 2*lines_indices[i]+e, computed on the device from the stored problem.
 * Outputs are (2m,) int64. */
int lba_point_wise_indices(lba_ctx* ctx, int64_t* camera_indices_pt, int64_t* pt_indices);

/* project (:23-74): p_hat for every endpoint-observation, (2m,3) f64, row 2i+e. */
int lba_project(lba_ctx* ctx, const double* camera_params, const double* points_3d, double* proj);

/* Track table of n_views consecutive views ("next" row N2): replaces Tracker.comp_tracker (src/tracker.py:11-96)
 * together with its verify_duplet calls (src/utils.py:1807-1944).
 *   endpoints (n_lines2d,2,2) f64: pixel coordinates of the endpoints of domain.lines2d[id] (the reference indexes
 *       that list by line id, utils.py:1874-1875); NULL: the table made resident by lba_lines2d_set;
 *   match_ptr (n_views,) int64 HOST: the matches of the view pair (v, v+1) are rows [match_ptr[v], match_ptr[v+1])
 *       of matches (M,2) int64 = domain.line_match["<v><v+1>"][1];
 *   F (n_views-1,3,3) f64: fundamental matrix of each pair (relativeCameraMotion utils.py:1483 + F_from_KRt2 :1039);
 *   threshold: 0.1 in the reference (tracker.py:53);  flags_in (M,) u8 or NULL: take these inlier flags instead of
 *       running the epipolar test;
 *   ids, bin (cap_rows,n_views) int64 out, cap_rows >= M: ids[track,view] (0 where absent) and bin[track,view]
 *       (tracker.py:93-95; the reference holds the same integers in float64 arrays); n_new_tracks (n_views-1,) int64
 *       HOST out; *n_tracks HOST out: rows used; flags_out (M,) u8 or NULL: inlier flag per match.
 * The epipolar arithmetic is rounded operation by operation like the reference's numpy code, so the flags equal the
 * reference's also where it divides by rounding noise (segments parallel to an epipolar line within a few ulps); only an
 * exactly zero determinant, on which the reference raises, is an outlier by rule. */
int lba_track_table(lba_ctx* ctx, int32_t n_views, int64_t n_lines2d, const double* endpoints,
                    const int64_t* match_ptr, const int64_t* matches, const double* F, double threshold,
                    const uint8_t* flags_in, int64_t cap_rows, int64_t* ids, int64_t* bin,
                    int64_t* n_new_tracks, int64_t* n_tracks, uint8_t* flags_out);

/* Two-view endpoint triangulation of n_tracks matched line segments ("next" row N3): replaces the per-line
 * loop of triangulate_model / triangulate_model2 (src/seq_utils.py:75-135, :11-72), i.e. for each endpoint
 *     X = cv2.triangulatePoints(P[view0], P[view1], x0, x1);  X = X[:3] / X[3]        (:119-121)
 * with OpenCV's DLT matrix and one-sided Jacobi SVD restated in fp64 and the float32 conversion / division of
 * the reference.  P (n_views,3,4) f64 row-major; view0, view1 (n_tracks,) int64 (view1 may be NULL: view0 + 1,
 * the reference's consecutive pairs); pts0, pts1 (n_tracks,2,2) f32 = [track][endpoint][x,y];
 * X (n_tracks,2,3) f32 out; Xh (n_tracks,2,4) f32 out or NULL (homogeneous, unit 4-vector up to sign);
 * seconds_kernel out or NULL (device time of the kernel).  Independent of lba_set_problem. */
int lba_triangulate_lines(lba_ctx* ctx, int64_t n_tracks, int64_t n_views, const double* P,
                          const int64_t* view0, const int64_t* view1, const float* pts0, const float* pts1,
                          float* X, float* Xh, double* seconds_kernel);

/* ---- "next" row N1: the BA input kept on the device across the incremental loop -------------------------------
 * The reference rebuilds the BA input from Python objects after every registered view (src/main.py:111-133:
 * Tracker.comp_tracker -> triangulate_model -> input_bundle_adj -> run_bundle_adjustment) and re-reads Line.n /
 * Line.points[e].p of every observation.  Here the per-2D-line constants are uploaded once and everything between
 * the track tables and the solve stays on the device.
 *
 * lba_lines2d_set: n, p0, p1 (n_lines2d,3) f64 = Line.n, Line.points[e].p of domain.lines2d[id] (domain.py:338-351);
 *   endpoints (n_lines2d,2,2) f64 pixel coordinates or NULL (needed by lba_triangulate_tracks).  Host or device. */
int lba_lines2d_set(lba_ctx* ctx, int64_t n_lines2d, const double* n, const double* p0, const double* p1,
                    const double* endpoints);
/* The observation list input_bundle_adj derives from the track tables (src/seq_utils.py:186-195): with
 * n_tracks = #rows of bin seen in more than one view, the entries bin[t,v] == 1 of the FIRST n_tracks rows in
 * row-major order give camera_indices = v, lines_indices = t and the 2D line ids[t,v].  ids, bin (n_rows,n_views)
 * int64 as written by lba_track_table (host or device); outputs (cap >= n_rows*n_views) int64, host or device;
 * *m_out, *n_tracks_out HOST. */
int lba_observations_from_tables(lba_ctx* ctx, int64_t n_rows, int32_t n_views, const int64_t* ids, const int64_t* bin,
                                 int64_t cap, int64_t* camera_indices, int64_t* lines_indices, int64_t* line2d_ids,
                                 int64_t* m_out, int64_t* n_tracks_out);
/* lba_set_problem with the observation constants gathered on the device from the resident per-2D-line table:
 * n[i], p0[i], p1[i] = row line2d_ids[i].  Index arrays host or device. */
int lba_set_problem_lines2d(lba_ctx* ctx, int64_t m, int64_t n_cameras, int64_t n_lines, const int64_t* camera_indices,
                            const int64_t* lines_indices, const int64_t* line2d_ids);
/* triangulate_model (src/seq_utils.py:105-131) straight from the DEVICE track table: the new tracks of view pair
 * (v, v+1) are the next n_new_tracks[v] rows of ids (HOST counts, as returned by lba_track_table); their float32
 * endpoints come from the resident endpoint table.  X (n,2,3) f32 and/or points_3d (2n,3) f64 (the float32
 * coordinates widened: the BA's starting point, bundle_adjustment.py:231-233), host or device; *n_tracks_out HOST. */
int lba_triangulate_tracks(lba_ctx* ctx, int32_t n_views, const double* P, const int64_t* ids,
                           const int64_t* n_new_tracks, float* X, double* points_3d, int64_t* n_tracks_out);

/* fun (:78-139): residuals (m,) in the caller's observation order; cost = 0.5*sum r^2
 * (either output may be NULL). */
int lba_residuals(lba_ctx* ctx, const double* camera_params, const double* points_3d,
                  double* residuals, double* cost);

/* Residuals and the analytic Jacobian rows that replace scipy's finite differences.
 * jcam is (m,6) in the reference's interleaved column order [w0,c0,w1,c1,w2,c2]
 * (columns 6*c_i+0..5), jline is (m,6) = d r_i / d [X0,X1] (columns 6*Nc+6*L_i+0..5).
 * Leaves the evaluation resident in the context for the calls below. */
int lba_residuals_jacobian(lba_ctx* ctx, const double* camera_params, const double* points_3d,
                           double* residuals, double* jcam, double* jline);

/* J^T J / J^T r blocks of the resident evaluation.  U (Nc,21) and V (Nl,21) are packed
 * upper triangles, row-major, in the INTERNAL parameter order [w(3),c(3)] / [X0,X1];
 * gc (Nc,6), gl (Nl,6).  With a communicator attached U and gc are all-reduced. */
int lba_normal_equations(lba_ctx* ctx, double* U, double* gc, double* V, double* gl);

/* Solve (J^T J + lambda D) delta = -J^T r for the resident evaluation by eliminating the
 * line blocks (Schur complement) and running block-Jacobi PCG on the reduced camera
 * system; D = diag(J^T J) (zeros -> 1).  dcam (Nc,6) in internal order [w,c], dpts (Nl,6). */
int lba_schur_solve(lba_ctx* ctx, double lambda, double pcg_tol, int32_t pcg_max_iter,
                    double* dcam, double* dpts, int32_t* pcg_iters);

/* The whole LM loop; replaces least_squares(...) at bundle_adjustment.py:319.
 * camera_params / points_3d are read and overwritten with the optimised values
 * (op_camera_params, op_points_3d of :326-329); residuals (m,) receives fun at the
 * solution (:332) in the caller's observation order (may be NULL). */
int lba_lm_solve(lba_ctx* ctx, double* camera_params, double* points_3d, const lba_options* opts,
                 lba_report* report, double* residuals);

/* ---- "next" row N4: the reference's 12-parameter camera rows -----------------------------------------------
 * camera_params (Nc,12) f64 = [R|c] reshaped from (3,4) (bundle_adjustment.py:43-45; input_bundle_adj with
 * rotate_model != "Rodrigues", seq_utils.py:168-169,182-184): the nine entries of R are free parameters.
 * lba_residuals12 = fun, lba_project12 = project (points (N,3), one camera row per point), lba_lm_solve12 = the
 * least_squares call: analytic Jacobian, dense normal equations + Cholesky (cuSOLVER, loaded on first use), the
 * same step policy, options and report as lba_lm_solve.  Meant for the reference's problem sizes
 * (12 Nc + 6 Nl <= 16384, one GPU). */
int lba_residuals12(lba_ctx* ctx, const double* camera_params12, const double* points_3d, double* residuals, double* cost);
int lba_project12(lba_ctx* ctx, int64_t n_rows, const double* points, const double* camera_params12, double* proj);
int lba_lm_solve12(lba_ctx* ctx, double* camera_params12, double* points_3d, const lba_options* opts,
                   lba_report* report, double* residuals);

/* Multi-GPU: one context per rank, observations sharded at 3D-line boundaries, cameras
 * replicated.  Per-camera partial sums are combined with ncclAllReduce on the context's
 * stream.  rank 0 calls lba_comm_unique_id, ships the 128 bytes to the other ranks (e.g.
 * torch.distributed broadcast), then every rank calls lba_comm_init. */
int lba_comm_unique_id(void* id128);
int lba_comm_init(lba_ctx* ctx, const void* id128, int32_t rank, int32_t world_size);
int lba_comm_destroy(lba_ctx* ctx);
/* Optional: one-shot all-reduce over NVLink peer memory instead of ncclAllReduce for the small per-camera
 * vectors.  Every rank allocates a symmetric buffer and exports a 64-byte CUDA IPC handle
 * (lba_comm_peer_local); the handles of all ranks, concatenated in rank order, are then attached
 * (lba_comm_peer_attach).  Payloads larger than half the buffer still go through NCCL. */
int lba_comm_peer_local(lba_ctx* ctx, void* handle64, int64_t bytes);
int lba_comm_peer_attach(lba_ctx* ctx, const void* handles, int32_t rank, int32_t world_size);

/* Instrumentation. */
int64_t lba_launch_count(lba_ctx* ctx);             /* kernels launched by this context so far */
int lba_set_profiling(lba_ctx* ctx, int32_t enable); /* bracket every launch with CUDA events */
/* ms[k], launches[k] for k < LBA_K_COUNT since profiling was enabled. */
int lba_get_profile(lba_ctx* ctx, double* ms, int64_t* launches);
const char* lba_kernel_class_name(int32_t k);
/* Dependent-DFMA microbenchmark: achieved fp64 TFLOP/s of this device (for the FP64 roofline). */
int lba_measure_fp64_peak(lba_ctx* ctx, double* tflops);
/* Streaming copy microbenchmark (GB/s, read+write) for a cross-check of MEASURED_PEAKS.json. */
int lba_measure_copy_bandwidth(lba_ctx* ctx, int64_t bytes, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* LBA_H_ */
