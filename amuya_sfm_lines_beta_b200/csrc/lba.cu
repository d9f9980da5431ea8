// liblba.so -- context, C ABI (include/lba.h) and the Levenberg-Marquardt driver.
//
// Replaces the scipy call at /root/reference/src/bundle_adjustment.py:319
// (least_squares TRF + LSMR + finite differences) with
//   fused residual/Jacobian kernel -> per-line / per-camera J^T J blocks ->
//   Schur elimination of the line blocks -> block-Jacobi PCG on the reduced camera
//   system (matrix-free) -> back-substitution -> trial evaluation -> accept / reject.
// The same algorithm is restated in numpy in oracle/line_ba_oracle.py (test oracle).
#include "../../include/lba.h"
#include "lba_kernels.cuh"
#include "lba_dense12.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

using namespace lba;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(LBA_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define CKR(call)            \
  do {                       \
    int r_ = (call);         \
    if (r_ != LBA_OK) return r_; \
  } while (0)

// ---- NCCL through dlopen: single-GPU use never loads it ----------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return LBA_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(LBA_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(LBA_ERR_COMM, "libnccl is missing required symbols");
  g_nccl.handle = h;
  return LBA_OK;
}

// ---- device buffer ------------------------------------------------------------------------------
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n && p) return LBA_OK;
    release();
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return fail(LBA_ERR_CUDA, "cudaMalloc(%zu B): %s", bytes, cudaGetErrorString(e)); }
    n = count;
    return LBA_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  ~DBuf() { release(); }
};

struct lba_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_order = nullptr;
  int n_sm = 148;

  // problem
  int m = 0, nc = 0, nl = 0, ntiles = 0, Mp = 0, lmax = 8;
  bool have_problem = false, have_eval = false, have_blocks = false, have_scaling = false;
  DBuf<int> cam_perm, cam_ptr, slot_src;
  DBuf<TileIdx> tidx;
  DBuf<int> sort_tmp, cam_sort_tmp, line_tmp;   // device radix-sort keys/values; per-line scratch
  DBuf<long long> idx_stage;     // the caller's int64 index arrays on the device
  DBuf<TileMeta> meta_stage;
  DBuf<unsigned char> cub_tmp;
  DBuf<double> cst, dense_tmp, h2d_stage;   // h2d_stage: grow-only staging for host-side n/p0/p1
  // persistent-grid sizes and dynamic shared memory of the pipelined line-order kernels
  int grid_eval_j = 1, grid_eval_r = 1, grid_lp[3] = {1, 1, 1};
  int smem_eval_j = 0, smem_eval_r = 0, smem_lp[3] = {0, 0, 0};
  bool eval_tab_smem = false;   // camera table resident in shared memory (fits up to ~1.3k cameras)
  int eval_stages_j = 3, eval_stages_r = 3;
  bool eval_z = true;           // Jacobian evaluation with the load pipeline decoupled from the per-line sums (k_eval_z; LBA_EVAL_Z=0: k_eval<true>, the round-1 kernel, kept for A/B runs)

  // state
  DBuf<double> cams, cams_new, pts, pts_new, Gm, Gm_new, api_tmp;
  DBuf<unsigned int> tickets;
  bool pending_first = true;
  int loss = 0;                // robust loss of the running lba_lm_solve (0 = linear)
  double f_scale = 1.0;
  int line_scaling_mode = 2;   // consumed by the next k_line_factor: 0 keep, 1 running max, 2 first
  DBuf<CamRec> recs, recs_new;
  DBuf<double> r, r_new, jl, jl_new, jc, jc_new, w, wk, wr;
  DBuf<double> V, V_new, gl, gl_new, Vinv, y, Dl, dl;
  DBuf<double> U, gc, Dc, Mpart, bpart, Minv, b;
  DBuf<double> cgx, cgr, cgz, cgp, tab, qpart;
  DBuf<double> cost_partial, bs_partial, cam_partial;
  DBuf<Scalars> scal, scal_loc;   // reduced over ranks (read by k_lm_decide and the host) / this rank's sums
  DBuf<LmCtl> ctl;        // device-resident state of the running LM solve (decisions are taken on the device)
  DBuf<LmTrace> trace;
  LmTrace* h_trace = nullptr;   // pinned
  DBuf<CgState> cgst;
  DBuf<GridBar> gridbar;
  DBuf<double> cg_partial;
  int cg_ctas = 1, cam_step_ctas = 1;
  int cg_cluster = 0;      // > 0: the CG vector kernels run as ONE cluster of this many CTAs (<= 1024 cameras, one GPU; LBA_CG_CLUSTER=0: grid kernels)
  int next_cluster = 0;    // cluster size of the next launch (consumed by launch_kernel)
  bool use_pdl = true;     // programmatic dependent launch between kernels (LBA_PDL=0|1|2, see lba_create)
  // CTA sizes of the one-CTA-per-camera kernels: block sums (k_cam_accum / k_cam_grad) and the PCG product
  int cam_threads = 128, pcg_threads = 128, cam_threads_sel = 128;
  DBuf<unsigned long long> gbits, gbits_new;   // gradient inf-norm bits of the two evaluation sets
  DBuf<int> clamped;
  // pinned, device-mapped mirror of the per-iteration scalars (written by k_publish); h_* point into it
  HostMirror* h_mirror = nullptr;
  HostMirror* d_mirror = nullptr;   // the same memory through the device's address
  unsigned long long publish_seq = 0;
  Scalars* h_scal = nullptr;
  CgState* h_cgst = nullptr;
  unsigned long long* h_gbits = nullptr;
  int* h_clamped = nullptr;
  LmCtl* h_ctl = nullptr;

  // comm
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // NVLink peer-memory all-reduce (optional; falls back to NCCL when not attached)
  bool peer_ready = false;
  void* peer_local = nullptr;          // own symmetric buffer (cudaMalloc)
  size_t peer_bytes = 0;
  void* peer_mapped[kMaxPeers] = {nullptr};
  PeerArgs peer_args;
  size_t peer_cams = 0;                     // camera capacity of the per-camera regions
  size_t rgn_off[kRgnCount] = {0}, rgn_cap[kRgnCount] = {0};   // in doubles (cap per slot)
  unsigned long long rgn_seq[kRgnCount] = {0};
  // two-phase (reduce-scatter + all-gather) peer reduce from this many ranks / doubles on;
  // LBA_PEER_TWO_PHASE_MIN=<doubles> overrides the size and drops the rank threshold to 2 (tests), -1 disables
  int peer_rs_world = 4;
  size_t peer_rs_min = 16384;
  DBuf<unsigned long long> peer_trace;   // LBA_PEER_TRACE: device-side wait/total times of the exchanges

  // lines with more than kTile observations (fragment tiles + fix-up kernels)
  int n_long = 0, n_frag = 0;
  DBuf<int> long_line, long_frag0, long_tile0;
  DBuf<double> fragV, fragT, frag_partial;
  LongLines long_lines() const { LongLines L; L.n = n_long; L.line = long_line.p; L.frag0 = long_frag0.p; L.tile0 = long_tile0.p; return L; }

  // N1: per-2D-line tables resident across the incremental loop (lba_lines2d_set)
  long long n_lines2d = 0;
  DBuf<double> l2d_n, l2d_p0, l2d_p1, l2d_end, l2d_stage;

  // second stream + event: the 72 B/observation constants cross PCIe while the index arrays are processed
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_done = nullptr;

  // L2 residency: part of the two streams every PCG iteration re-reads (jl in line order, jc in camera order)
  // is kept in a persisting L2 carve-out; win_* is the access-policy window of the NEXT launch (0 = none)
  size_t l2_persist = 0, l2_max_window = 0;
  const void* win_ptr = nullptr;
  size_t win_bytes = 0;
  float win_ratio = 0.f;
  void window(const void* p, size_t bytes, size_t share) {
    if (!l2_persist || !bytes) return;
    win_ptr = p;
    win_bytes = std::min(bytes, l2_max_window);
    win_ratio = (float)std::min(1.0, (double)share / (double)win_bytes);
  }

  // instrumentation
  int64_t launches = 0;
  bool profiling = false;
  double prof_ms[LBA_K_COUNT] = {0};
  int64_t prof_n[LBA_K_COUNT] = {0};
  int launch_err = 0;

  void pre(int cls) {
    ++launches;
    if (profiling) cudaEventRecord(ev0, stream);
    (void)cls;
  }
  void post(int cls) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && !launch_err) {
      launch_err = 1;
      fail(LBA_ERR_CUDA, "kernel launch (class %d) failed: %s", cls, cudaGetErrorString(e));
    }
    if (profiling) {
      cudaEventRecord(ev1, stream);
      cudaEventSynchronize(ev1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev0, ev1);
      prof_ms[cls] += ms;
      prof_n[cls] += 1;
    }
  }
};

// Every kernel of this library starts with pdl_prologue(), so all of them may be launched with the
// programmatic-stream-serialization attribute (off while profiling: the class events sit between launches).
template <typename... KArgs, typename... Args>
static inline void launch_kernel(lba_ctx* c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[3];
  int na = 0;
  if (c->next_cluster) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)c->next_cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    c->next_cluster = 0;
  }
  if (c->use_pdl && !c->profiling) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (c->win_bytes) {
    at[na].id = cudaLaunchAttributeAccessPolicyWindow;
    at[na].val.accessPolicyWindow.base_ptr = const_cast<void*>(c->win_ptr);
    at[na].val.accessPolicyWindow.num_bytes = c->win_bytes;
    at[na].val.accessPolicyWindow.hitRatio = c->win_ratio;
    at[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++na;
    c->win_bytes = 0;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

#define LAUNCH_DYN(ctx, cls, kern, grid, block, smem, ...)                      \
  do {                                                                          \
    (ctx)->pre(cls);                                                            \
    launch_kernel((ctx), kern, dim3(grid), dim3(block), (smem), __VA_ARGS__);   \
    (ctx)->post(cls);                                                           \
  } while (0)
#define LAUNCH(ctx, cls, kern, grid, block, ...) LAUNCH_DYN(ctx, cls, kern, grid, block, 0, __VA_ARGS__)

// launch a camera-order kernel template with the CTA size chosen for this problem
#define LAUNCH_CAM(ctx, cls, kern, grid, ...)                                        \
  do {                                                                               \
    switch ((ctx)->cam_threads_sel) {                                                \
      case 512: LAUNCH(ctx, cls, kern<512>, grid, 512, __VA_ARGS__); break;          \
      case 256: LAUNCH(ctx, cls, kern<256>, grid, 256, __VA_ARGS__); break;          \
      case 64: LAUNCH(ctx, cls, kern<64>, grid, 64, __VA_ARGS__); break;             \
      case 32: LAUNCH(ctx, cls, kern<32>, grid, 32, __VA_ARGS__); break;             \
      default: LAUNCH(ctx, cls, kern<128>, grid, 128, __VA_ARGS__);                  \
    }                                                                                \
  } while (0)
// same for a camera-order kernel with one more (bool) template argument
#define LAUNCH_CAM2(ctx, cls, kern, flag, grid, ...)                                         \
  do {                                                                                       \
    switch ((ctx)->cam_threads_sel) {                                                        \
      case 512: LAUNCH(ctx, cls, (kern<512, flag>), grid, 512, __VA_ARGS__); break;          \
      case 256: LAUNCH(ctx, cls, (kern<256, flag>), grid, 256, __VA_ARGS__); break;          \
      case 64: LAUNCH(ctx, cls, (kern<64, flag>), grid, 64, __VA_ARGS__); break;             \
      case 32: LAUNCH(ctx, cls, (kern<32, flag>), grid, 32, __VA_ARGS__); break;             \
      default: LAUNCH(ctx, cls, (kern<128, flag>), grid, 128, __VA_ARGS__);                  \
    }                                                                                        \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// a failed launch is reported once, at the API boundary that notices it (the message was recorded by post())
static int take_launch_err(lba_ctx* c) {
  if (c->world > 1) {   // called after a stream synchronisation: did a bounded device-side wait give up?
    unsigned int code = 0;
    if (cudaMemcpyFromSymbol(&code, g_spin_error, sizeof code) == cudaSuccess && code) {
      const unsigned int zero = 0;
      cudaMemcpyToSymbol(g_spin_error, &zero, sizeof zero);
      return fail(LBA_ERR_COMM, "a device-side wait on a peer rank timed out (code %u); results of this call are invalid", code);
    }
  }
  if (!c->launch_err) return LBA_OK;
  c->launch_err = 0;
  return LBA_ERR_CUDA;
}

static const char* kClassNames[LBA_K_COUNT] = {
    "residual", "jacobian", "cam_blocks", "line_factor", "line_scalars", "cam_precond",
    "pcg_line", "pcg_cam", "pcg_update", "backsub", "misc", "comm"};

// ---- C ABI: lifecycle ---------------------------------------------------------------------------
extern "C" const char* lba_last_error(void) { return g_err.c_str(); }
extern "C" int lba_version(void) { return 100; }
extern "C" const char* lba_kernel_class_name(int32_t k) { return (k >= 0 && k < LBA_K_COUNT) ? kClassNames[k] : "?"; }

extern "C" void lba_default_options(lba_options* o) {
  if (!o) return;
  o->max_iter = 50;
  o->pcg_max_iter = 200;
  o->ftol = 1e-4;
  o->xtol = 1e-8;
  o->gtol = 1e-8;
  o->lambda0 = 1e-2;
  o->lambda_min = 1e-6;
  o->lambda_max = 1e12;
  o->pcg_tol = 1e-2;
  o->min_ratio = 1e-4;
  o->verbose = 0;
  o->flags = 0;
  o->loss = LBA_LOSS_LINEAR;
  o->pad = 0;
  o->f_scale = 1.0;
  o->max_evals = 0;
  o->max_backtrack = 1;
}

extern "C" int lba_create(int device, lba_ctx** out) {
  if (!out) return fail(LBA_ERR_ARG, "lba_create: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(LBA_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count) return fail(LBA_ERR_ARG, "device %d out of range (count %d)", device, count);
  CK(cudaSetDevice(device));
  lba_ctx* c = new lba_ctx();
  struct CreateGuard { lba_ctx* c; bool ok = false; ~CreateGuard() { if (!ok) lba_destroy(c); } } guard{c};   // no leak on a failed create
  c->device = device;
  {  // bound of every device-side wait on another CTA / GPU (see SpinGuard)
    const char* e = getenv("LBA_SPIN_TIMEOUT_MS");
    const unsigned long long ns = (e && atoll(e) > 0 ? (unsigned long long)atoll(e) : 10000ULL) * 1000000ULL;
    CK(cudaMemcpyToSymbol(g_spin_timeout_ns, &ns, sizeof ns));
  }
  if (const char* e = getenv("LBA_L2_PERSIST_MB")) {
    int max_persist = 0, max_win = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, device);
    c->l2_persist = std::min((size_t)atoi(e) << 20, (size_t)max_persist);
    c->l2_max_window = (size_t)max_win;
    if (c->l2_persist) CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist));
    if (getenv("LBA_DEBUG_TIMING"))
      fprintf(stderr, "[lba_create] L2 persisting carve-out %zu MB (max %d MB), max window %d MB\n", c->l2_persist >> 20,
              max_persist >> 20, max_win >> 20);
  }
  {  // LBA_PDL=0: plain launches; 1 (default): programmatic stream serialization, dependents released when the
     // CTAs of a kernel exit (+6% LM it/s on config 4); 2: dependents triggered at kernel start (measured 12% slower:
     // the parked CTAs of kernel N+1 take shared memory and L1 away from kernel N)
    const char* e = getenv("LBA_PDL");
    const int mode = e ? atoi(e) : 1;
    c->use_pdl = mode > 0;
    const int early = mode > 1;
    CK(cudaMemcpyToSymbol(g_pdl_early_trigger, &early, sizeof(int)));
  }
  cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device);
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
  CK(cudaHostAlloc(&c->h_mirror, sizeof(HostMirror), cudaHostAllocMapped));
  memset(c->h_mirror, 0, sizeof(HostMirror));
  CK(cudaHostGetDevicePointer(&c->d_mirror, c->h_mirror, 0));
  c->h_scal = &c->h_mirror->scal; c->h_cgst = &c->h_mirror->cg;
  c->h_gbits = &c->h_mirror->gbits; c->h_clamped = &c->h_mirror->clamped; c->h_ctl = &c->h_mirror->ctl;
  CKR(c->scal.alloc(1));
  CKR(c->scal_loc.alloc(1));
  CK(cudaMemsetAsync(c->scal_loc.p, 0, sizeof(Scalars), c->stream));
  CKR(c->ctl.alloc(1));
  CKR(c->trace.alloc(1));
  CK(cudaMemsetAsync(c->ctl.p, 0, sizeof(LmCtl), c->stream));
  CK(cudaHostAlloc(&c->h_trace, sizeof(LmTrace), cudaHostAllocDefault));
  CKR(c->cgst.alloc(1));
  CKR(c->gridbar.alloc(1));
  CKR(c->cg_partial.alloc(2 * kCgMaxCtas));
  CK(cudaMemsetAsync(c->gridbar.p, 0, sizeof(GridBar), c->stream));
  CKR(c->gbits.alloc(1));
  CKR(c->gbits_new.alloc(1));
  CK(cudaMemsetAsync(c->gbits.p, 0, 8, c->stream));
  CK(cudaMemsetAsync(c->gbits_new.p, 0, 8, c->stream));
  CKR(c->clamped.alloc(1));
  CKR(c->tickets.alloc(8));   // 0 eval, 1 backsub, 2 cam step, 3 fused link, 4 two-phase reduce, 5 long lines, 6 dense 12-parameter path
  CK(cudaMemsetAsync(c->tickets.p, 0, 32, c->stream));
  CK(cudaMemsetAsync(c->scal.p, 0, sizeof(Scalars), c->stream));
  CK(cudaMemsetAsync(c->clamped.p, 0, sizeof(int), c->stream));
  guard.ok = true;
  *out = c;
  return LBA_OK;
}

extern "C" int lba_comm_destroy(lba_ctx* ctx) {
  if (ctx && ctx->peer_local) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->peer_trace.p) {
      unsigned long long h[32];
      cudaMemcpy(h, ctx->peer_trace.p, sizeof h, cudaMemcpyDeviceToHost);
      static const char* nm[kRgnCount] = {"S*p (unfused)", "U|gc", "M|b|diag|gc", "scalars", "(results)"};
      for (int g = 0; g < kRgnCount; ++g)
        if (h[g * 4])
          fprintf(stderr, "[lba peer trace r%d] %-14s %6llu calls: wait %7.2f us, total %7.2f us per call\n", ctx->rank, nm[g],
                  h[g * 4], h[g * 4 + 1] * 1e-3 / h[g * 4], h[g * 4 + 2] * 1e-3 / h[g * 4]);
      if (h[24])
        fprintf(stderr, "[lba peer trace r%d] %-14s %6llu calls: wait %7.2f us per call\n", ctx->rank, "S*p (fused)", h[24],
                h[25] * 1e-3 / h[24]);
      ctx->peer_args.trace = nullptr;
    }
    for (int r = 0; r < kMaxPeers; ++r)
      if (ctx->peer_mapped[r] && ctx->peer_mapped[r] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_mapped[r]);
    cudaFree(ctx->peer_local);
    ctx->peer_local = nullptr;
    ctx->peer_ready = false;
    for (int r = 0; r < kMaxPeers; ++r) ctx->peer_mapped[r] = nullptr;
  }
  if (ctx && ctx->comm) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
  }
  return LBA_OK;
}

extern "C" int lba_destroy(lba_ctx* ctx) {
  if (!ctx) return LBA_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  lba_comm_destroy(ctx);
  if (ctx->h_mirror) cudaFreeHost(ctx->h_mirror);
  if (ctx->h_trace) cudaFreeHost(ctx->h_trace);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return LBA_OK;
}

extern "C" int lba_set_stream(lba_ctx* ctx, void* s) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return LBA_OK;
}

// Order the context's stream behind work already enqueued on another stream of the same device (e.g. the PyTorch
// stream that produced a tensor about to be passed in): event record on `s`, wait on the context's stream.
extern "C" int lba_wait_stream(lba_ctx* ctx, void* s) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if ((cudaStream_t)s == ctx->stream) return LBA_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->ev_order, (cudaStream_t)s));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_order, 0));
  if (ctx->copy_stream) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_order, 0));
  return LBA_OK;
}

extern "C" int lba_synchronize(lba_ctx* ctx) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->launch_err) return take_launch_err(ctx);
  return LBA_OK;
}

extern "C" int64_t lba_launch_count(lba_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lba_set_profiling(lba_ctx* ctx, int32_t enable) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  ctx->profiling = enable != 0;
  for (int k = 0; k < LBA_K_COUNT; ++k) { ctx->prof_ms[k] = 0; ctx->prof_n[k] = 0; }
  return LBA_OK;
}

extern "C" int lba_get_profile(lba_ctx* ctx, double* ms, int64_t* launches) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  for (int k = 0; k < LBA_K_COUNT; ++k) {
    if (ms) ms[k] = ctx->prof_ms[k];
    if (launches) launches[k] = ctx->prof_n[k];
  }
  return LBA_OK;
}

// ---- comm -----------------------------------------------------------------------------------------
extern "C" int lba_comm_unique_id(void* id128) {
  if (!id128) return fail(LBA_ERR_ARG, "id is NULL");
  CKR(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(LBA_ERR_COMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  memcpy(id128, &id, 128);
  return LBA_OK;
}

extern "C" int lba_comm_init(lba_ctx* ctx, const void* id128, int32_t rank, int32_t world) {
  if (!ctx || !id128) return fail(LBA_ERR_ARG, "ctx or id is NULL");
  if (world < 1 || rank < 0 || rank >= world) return fail(LBA_ERR_ARG, "bad rank %d / world %d", rank, world);
  CKR(load_nccl());
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
  if (r != ncclSuccess) return fail(LBA_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  ctx->rank = rank;
  ctx->world = world;
  return LBA_OK;
}

static int nccl_allreduce(lba_ctx* c, void* buf, size_t n, ncclDataType_t dt, ncclRedOp_t op, const void* src = nullptr) {
  c->pre(LBA_K_COMM);
  ncclResult_t r = g_nccl.AllReduce(src ? src : buf, buf, n, dt, op, c->comm, c->stream);
  c->post(LBA_K_COMM);
  --c->launches;  // a library collective is not one of our kernels
  if (r != ncclSuccess) return fail(LBA_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return LBA_OK;
}

static inline bool peer_fits(const lba_ctx* c, int region, size_t n) {
  return c->world > 1 && c->peer_ready && n <= c->rgn_cap[region];
}
static inline size_t peer_slot_off(const lba_ctx* c, int region, unsigned long long seq) {
  return c->rgn_off[region] + (size_t)(seq & 1ULL) * c->rgn_cap[region];
}

// A fused producer -> consumer all-reduce on `region`: returns the slot the producer must write and fills
// `pl` (pl.seq == 0 when the peer transport is not active: the caller then uses reduce_finish).
static double* fused_link(lba_ctx* c, int region, size_t n, double* local, PeerLink* pl) {
  memset(pl, 0, sizeof *pl);
  if (!peer_fits(c, region, n)) return local;
  pl->pa = c->peer_args; pl->region = region; pl->seq = ++c->rgn_seq[region];
  pl->slot_off = peer_slot_off(c, region, pl->seq); pl->ticket = c->tickets.p + 3;
  return reinterpret_cast<double*>((unsigned char*)c->peer_local + kPeerDataOff) + pl->slot_off;
}

// Where a producer kernel should write the partial sums that will be all-reduced into `local`:
// straight into the symmetric slot of the NEXT call on that region when the peer path is active.
static double* reduce_target(lba_ctx* c, int region, double* local, size_t n) {
  if (!peer_fits(c, region, n)) return local;
  return reinterpret_cast<double*>((unsigned char*)c->peer_local + kPeerDataOff) + peer_slot_off(c, region, c->rgn_seq[region] + 1);
}

// Finish the all-reduce started by reduce_target: result in `local` on every rank.
static int reduce_finish(lba_ctx* c, int region, double* local, size_t n) {
  if (c->world <= 1 || n == 0) return LBA_OK;
  if (!peer_fits(c, region, n)) return nccl_allreduce(c, local, n, ncclDouble, ncclSum);
  const unsigned long long seq = ++c->rgn_seq[region];
  // one double2 per thread per pass: the loop over a slot is a chain of NVLink round trips (~2.5 us each), so the
  // grid is sized to make it one or two passes (measured at 8 ranks: 30 us with 32 CTAs)
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)2 * c->n_sm, (n / 2 + 511) / 512));
  if (c->world >= c->peer_rs_world && n >= c->peer_rs_min && n <= c->rgn_cap[kRgnR]) {
    // large exchange, many ranks: reduce-scatter + all-gather through the result region
    const unsigned long long rseq = ++c->rgn_seq[kRgnR];
    c->pre(LBA_K_COMM);
    // (all CTAs of this kernel must be co-resident, they wait for each other's ranks: at most one per SM)
    const int grid_rs = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->n_sm, (n / 2 + 1023) / 1024));
    launch_kernel(c, k_peer_reduce_rs, dim3(grid_rs), dim3(512), 0, c->peer_args, region, peer_slot_off(c, region, seq),
                  peer_slot_off(c, kRgnR, rseq), n, seq, rseq, local, c->tickets.p + 4);
    c->post(LBA_K_COMM);
    return LBA_OK;
  }
  c->pre(LBA_K_COMM);
  launch_kernel(c, k_peer_reduce<false>, dim3(grid), dim3(512), 0, c->peer_args, region, peer_slot_off(c, region, seq), n, seq,
                (const double*)nullptr, local, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (const LmCtl*)nullptr);
  c->post(LBA_K_COMM);
  return LBA_OK;
}

// Small all-reduce of a few scalars that live in ordinary device memory (copy variant): buf <- reduce(src), in
// place when src is NULL.
static int allreduce(lba_ctx* c, void* buf, size_t n, ncclDataType_t dt, ncclRedOp_t op, const void* src = nullptr) {
  if (c->world <= 1 || n == 0) return LBA_OK;
  if (!peer_fits(c, kRgnS, n)) return nccl_allreduce(c, buf, n, dt, op, src);
  if (!src) src = buf;
  const unsigned long long seq = ++c->rgn_seq[kRgnS];
  c->pre(LBA_K_COMM);
  if (op == ncclMax)
    launch_kernel(c, k_peer_reduce<true>, dim3(1), dim3(512), 0, c->peer_args, (int)kRgnS, peer_slot_off(c, kRgnS, seq), n, seq,
                  (const double*)src, (double*)buf, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (const LmCtl*)nullptr);
  else
    launch_kernel(c, k_peer_reduce<false>, dim3(1), dim3(512), 0, c->peer_args, (int)kRgnS, peer_slot_off(c, kRgnS, seq), n, seq,
                  (const double*)src, (double*)buf, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (const LmCtl*)nullptr);
  c->post(LBA_K_COMM);
  return LBA_OK;
}
// The per-pass sums are produced into a LOCAL scalar block (scal_loc) and reduced into the one the decision kernel
// and the host read (scal), so that a pass which re-uses sums of an earlier pass (a retry does not repeat the
// back-substitution) never reduces an already reduced value again.  Single rank: the kernels write scal directly.
static inline Scalars* scal_local(lba_ctx* c) { return c->world > 1 ? c->scal_loc.p : c->scal.p; }
// scal.cost .. (n doubles) <- sum over ranks of the local block
static int reduce_scalars(lba_ctx* c, size_t n) {
  return allreduce(c, &c->scal.p->cost, n, ncclDouble, ncclSum, &scal_local(c)->cost);
}
// the end-of-trial scalars (sum, local block -> scal) and the gradient inf-norm bits (max) in ONE exchange.  With the
// peer transport the kernel picks the word of the current evaluation set from the control block; the NCCL fallback
// reduces both.
static int allreduce_scalars(lba_ctx* c, size_t n, const LmCtl* ctl) {
  if (c->world <= 1) return LBA_OK;
  if (n > 8 || !peer_fits(c, kRgnS, 10)) {
    CKR(reduce_scalars(c, n));
    CKR(allreduce(c, c->gbits.p, 1, ncclUint64, ncclMax));
    return allreduce(c, c->gbits_new.p, 1, ncclUint64, ncclMax);
  }
  const unsigned long long seq = ++c->rgn_seq[kRgnS];
  c->pre(LBA_K_COMM);
  launch_kernel(c, k_peer_reduce<false>, dim3(1), dim3(512), 0, c->peer_args, (int)kRgnS, peer_slot_off(c, kRgnS, seq), n, seq,
                (const double*)&scal_local(c)->cost, &c->scal.p->cost, c->gbits.p, c->gbits_new.p, ctl);
  c->post(LBA_K_COMM);
  return LBA_OK;
}

extern "C" int lba_comm_peer_local(lba_ctx* ctx, void* handle64, int64_t bytes) {
  if (!ctx || !handle64 || bytes < 4096) return fail(LBA_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (ctx->peer_local) return fail(LBA_ERR_STATE, "peer buffer already allocated");
  CK(cudaMalloc(&ctx->peer_local, (size_t)bytes));
  CK(cudaMemset(ctx->peer_local, 0, (size_t)bytes));
  ctx->peer_bytes = (size_t)bytes;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->peer_local));
  memcpy(handle64, &h, 64);
  return LBA_OK;
}

extern "C" int lba_comm_peer_attach(lba_ctx* ctx, const void* handles, int32_t rank, int32_t world) {
  if (!ctx || !handles) return fail(LBA_ERR_ARG, "bad argument");
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world) return fail(LBA_ERR_UNSUPPORTED, "peer all-reduce supports 2..%d ranks", kMaxPeers);
  if (!ctx->peer_local) return fail(LBA_ERR_STATE, "call lba_comm_peer_local first");
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ctx->peer_mapped[r] = ctx->peer_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + (size_t)r * 64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&ctx->peer_mapped[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LBA_ERR_COMM, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e)); }
  }
  memset(&ctx->peer_args, 0, sizeof ctx->peer_args);
  for (int r = 0; r < world; ++r) ctx->peer_args.peer[r] = (unsigned char*)ctx->peer_mapped[r];
  ctx->peer_args.rank = rank; ctx->peer_args.world = world;
  if (const char* e = getenv("LBA_PEER_TWO_PHASE_MIN")) {
    const long long v = atoll(e);
    if (v < 0) ctx->peer_rs_world = kMaxPeers + 1;
    else { ctx->peer_rs_world = 2; ctx->peer_rs_min = (size_t)v; }
  }
  if (getenv("LBA_PEER_TRACE")) {
    CKR(ctx->peer_trace.alloc(32));
    CK(cudaMemset(ctx->peer_trace.p, 0, 32 * 8));
    ctx->peer_args.trace = ctx->peer_trace.p;
  }
  // regions: scalars (16), S*p partial (6 Nc), U|gc (27 Nc), M|b|diag U|gc (39 Nc), two-phase results (39 Nc);
  // two slots each
  const size_t total = (ctx->peer_bytes - kPeerDataOff) / 8;
  const size_t ncap = (total - 2 * 16) / (2 * (6 + 27 + 39 + 39));
  ctx->peer_cams = ncap;
  size_t off = 0;
  const size_t caps[kRgnCount] = {6 * ncap, 27 * ncap, 39 * ncap, 16, 39 * ncap};
  for (int g = 0; g < kRgnCount; ++g) { ctx->rgn_off[g] = off; ctx->rgn_cap[g] = caps[g] & ~(size_t)1; off += 2 * ctx->rgn_cap[g]; ctx->rgn_seq[g] = 0; }
  ctx->peer_ready = true;
  return LBA_OK;
}

// ---- problem set-up -------------------------------------------------------------------------------
static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename T>
static int upload(lba_ctx* c, DBuf<T>& dst, const std::vector<T>& src) {
  CKR(dst.alloc(src.size()));
  if (!src.empty()) CK(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return LBA_OK;
}

extern "C" int lba_set_problem(lba_ctx* ctx, int64_t m64, int64_t nc64, int64_t nl64, const int64_t* cam_idx,
                               const int64_t* line_idx, const double* n, const double* p0, const double* p1) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (m64 < 0 || nc64 < 0 || nl64 < 0 || m64 > 0x7fffffffLL / 12 || nc64 > 0x7fffffffLL / 32 || nl64 > 0x7fffffffLL / 32)
    return fail(LBA_ERR_ARG, "problem size out of range (m=%lld Nc=%lld Nl=%lld)", (long long)m64, (long long)nc64, (long long)nl64);
  if (m64 > 0 && (!cam_idx || !line_idx || !n || !p0 || !p1)) return fail(LBA_ERR_ARG, "NULL input array");
  CK(cudaSetDevice(ctx->device));
  const int m = (int)m64, nc = (int)nc64, nl = (int)nl64;
  ctx->have_problem = ctx->have_eval = ctx->have_blocks = ctx->have_scaling = false;
  const bool dbg = getenv("LBA_DEBUG_TIMING") != nullptr;
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[lba_set_problem] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };

  // (1) the caller's index arrays go to the device; range checks, sortedness and the per-line
  //     observation histogram are computed there
  CKR(ctx->idx_stage.alloc((size_t)2 * m + 2));
  long long* d_cam = ctx->idx_stage.p;
  long long* d_line = ctx->idx_stage.p + m;
  const long long* cam_dev = d_cam;
  const long long* line_dev = d_line;
  if (m > 0) {
    if (is_device_ptr(cam_idx)) cam_dev = (const long long*)cam_idx;
    else CK(cudaMemcpyAsync(d_cam, cam_idx, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (is_device_ptr(line_idx)) line_dev = (const long long*)line_idx;
    else CK(cudaMemcpyAsync(d_line, line_idx, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  // the observation constants (n, p0, p1: 72 B per observation, the bulk of the input) start crossing PCIe now,
  // on the copy stream, behind the index arrays; k_setup_constants waits for them with an event
  const double* cst_src[3] = {n, p0, p1};
  const double* cst_dev[3] = {n, p0, p1};
  // whatever way this function returns, the caller's arrays are no longer being read
  struct CopyGuard { cudaStream_t s; ~CopyGuard() { cudaStreamSynchronize(s); } } copy_guard{ctx->copy_stream};
  if (m > 0) {
    size_t need = 0;
    for (int k = 0; k < 3; ++k) if (!is_device_ptr(cst_src[k])) need += (size_t)3 * m;
    CKR(ctx->h2d_stage.alloc(need));
    CK(cudaEventRecord(ctx->copy_done, ctx->stream));            // index copies first on the link
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done, 0));
    size_t off = 0;
    for (int k = 0; k < 3; ++k) {
      if (is_device_ptr(cst_src[k])) continue;
      CK(cudaMemcpyAsync(ctx->h2d_stage.p + off, cst_src[k], (size_t)3 * m * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      cst_dev[k] = ctx->h2d_stage.p + off;
      off += (size_t)3 * m;
    }
    CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
  }
  CKR(ctx->line_tmp.alloc((size_t)3 * nl + 8));   // line_cnt | line_ptr | line_slot
  int* d_line_cnt = ctx->line_tmp.p; int* d_line_ptr = d_line_cnt + nl + 1; int* d_line_slot = d_line_ptr + nl + 1;
  int* d_flags = d_line_slot + nl + 1;
  CK(cudaMemsetAsync(ctx->line_tmp.p, 0, ((size_t)3 * nl + 8) * 4, ctx->stream));
  CKR(ctx->sort_tmp.alloc((size_t)4 * std::max(m, 1)));
  if (m > 0)
    LAUNCH(ctx, LBA_K_MISC, k_index_scan, cdiv(m, 256), 256, m, nc, nl, cam_dev, line_dev, d_line_cnt, d_flags, ctx->sort_tmp.p);
  std::vector<int> line_cnt((size_t)nl + 1, 0);
  int h_flags[3] = {0, 0, 0};
  if (nl > 0) CK(cudaMemcpyAsync(line_cnt.data(), d_line_cnt, (size_t)nl * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_flags, d_flags, 12, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (h_flags[0]) return fail(LBA_ERR_ARG, "camera_indices[%d] out of range [0,%d)", h_flags[0] - 1, nc);
  if (h_flags[1]) return fail(LBA_ERR_ARG, "lines_indices[%d] out of range [0,%d)", h_flags[1] - 1, nl);
  const bool sorted = h_flags[2] == 0;
  lap("index H2D + device scan");

  // (2) per-LINE work on the host: prefix sums and the greedy packing of whole lines into tiles
  std::vector<int> line_ptr((size_t)nl + 1, 0), line_slot((size_t)nl + 1, 0);
  std::vector<TileMeta> meta;
  std::vector<int> long_line, long_frag0, long_tile0;
  int kmax = 0, n_frag = 0;
  {
    TileMeta cur;
    memset(&cur, 0, sizeof cur);
    bool open = false;
    for (int l = 0; l < nl; ++l) {
      const int k = line_cnt[l];
      kmax = std::max(kmax, k);
      line_ptr[(size_t)l + 1] = line_ptr[l] + k;
      if (k > kTile) {
        // a line that does not fit one tile: consecutive fragment tiles of its own (slots stay consecutive)
        if (open) { meta.push_back(cur); open = false; }
        long_line.push_back(l); long_frag0.push_back(n_frag); long_tile0.push_back((int)meta.size());
        line_slot[l] = (int)meta.size() * kTile;
        for (int done = 0; done < k; done += kTile) {
          TileMeta f;
          memset(&f, 0, sizeof f);
          f.l0 = l; f.nl = 1; f.nobs = std::min(kTile, k - done); f.frag = ++n_frag; f.lend[0] = (unsigned short)f.nobs;
          meta.push_back(f);
        }
        continue;
      }
      if (open && (cur.nobs + k > kTile || cur.nl >= kMaxLinesPerTile)) { meta.push_back(cur); open = false; }
      if (!open) { memset(&cur, 0, sizeof cur); cur.l0 = l; open = true; }
      line_slot[l] = (int)meta.size() * kTile + cur.nobs;
      cur.nobs += k;
      cur.lend[cur.nl++] = (unsigned short)std::min(cur.nobs, 65535);
    }
    if (open) meta.push_back(cur);
  }
  long_frag0.push_back(n_frag);
  const int ntiles = (int)meta.size();
  const long long Mp64 = (long long)ntiles * kTile;
  if (Mp64 > 0x7fffffffLL / 12) return fail(LBA_ERR_ARG, "padded observation count out of range");
  const int Mp = (int)Mp64;
  int lmax = 8;
  for (const TileMeta& t : meta) lmax = std::max(lmax, (t.nl + 1) & ~1);   // sizes the per-line parts of the shared-memory stages (16 B granules)
  lap("host tiling (per line)");

  // (3) back on the device: slot assignment of every observation
  ctx->m = m; ctx->nc = nc; ctx->nl = nl; ctx->ntiles = ntiles; ctx->Mp = Mp; ctx->lmax = lmax;
  CKR(ctx->tidx.alloc((size_t)ntiles)); CKR(ctx->slot_src.alloc((size_t)Mp)); CKR(ctx->meta_stage.alloc((size_t)ntiles));
  if (ntiles > 0) {
    CK(cudaMemcpyAsync(ctx->meta_stage.p, meta.data(), (size_t)ntiles * sizeof(TileMeta), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_line_ptr, line_ptr.data(), ((size_t)nl + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_line_slot, line_slot.data(), ((size_t)nl + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, LBA_K_MISC, k_init_tidx, ntiles, kTile, ntiles, ctx->meta_stage.p, ctx->tidx.p);
    CK(cudaMemsetAsync(ctx->slot_src.p, 0xFF, (size_t)Mp * 4, ctx->stream));   // -1 = padding
  }
  ctx->n_long = (int)long_line.size(); ctx->n_frag = n_frag;
  if (ctx->n_long > 0) {
    CKR(upload(ctx, ctx->long_line, long_line)); CKR(upload(ctx, ctx->long_frag0, long_frag0)); CKR(upload(ctx, ctx->long_tile0, long_tile0));
    CKR(ctx->fragV.alloc((size_t)n_frag * 27)); CKR(ctx->fragT.alloc((size_t)n_frag * 6)); CKR(ctx->frag_partial.alloc((size_t)ctx->n_long * 4));
  }
  if (m > 0) {
    const int* d_order = nullptr;
    if (!sorted) {
      // stable sort of the observations by line (the reference emits them sorted: seq_utils.py:190)
      int* keys_in = ctx->sort_tmp.p; int* vals_in = keys_in + m; int* keys_out = vals_in + m; int* vals_out = keys_out + m;
      LAUNCH(ctx, LBA_K_MISC, k_iota, cdiv(m, 256), 256, m, vals_in);
      int end_bit = 1;
      while ((1LL << end_bit) <= nl) ++end_bit;
      size_t tmp_bytes = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, m, 0, end_bit, ctx->stream));
      CKR(ctx->cub_tmp.alloc(tmp_bytes));
      CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, keys_in, keys_out, vals_in, vals_out, m, 0, end_bit, ctx->stream));
      d_order = vals_out;
    }
    LAUNCH(ctx, LBA_K_MISC, k_fill_slots, cdiv(m, 256), 256, m, d_order, cam_dev, line_dev, d_line_ptr, d_line_slot,
           ctx->slot_src.p, ctx->tidx.p);
  }
  // camera order on the device: stable LSD radix sort of (camera, slot) keeps the slot (line) order
  // inside every camera, like the stable counting sort it replaces
  CKR(ctx->cam_perm.alloc((size_t)Mp)); CKR(ctx->cam_ptr.alloc((size_t)nc + 1));
  if (Mp > 0) {
    CKR(ctx->cam_sort_tmp.alloc((size_t)3 * Mp));
    int* keys_in = ctx->cam_sort_tmp.p; int* vals_in = keys_in + Mp; int* keys_out = vals_in + Mp;
    LAUNCH(ctx, LBA_K_MISC, k_cam_keys, cdiv(Mp, 256), 256, Mp, nc, ctx->tidx.p, keys_in, vals_in);
    int end_bit = 1;
    while ((1LL << end_bit) <= nc) ++end_bit;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, ctx->cam_perm.p, Mp, 0, end_bit, ctx->stream));
    CKR(ctx->cub_tmp.alloc(tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, keys_in, keys_out, vals_in, ctx->cam_perm.p, Mp, 0, end_bit, ctx->stream));
    LAUNCH(ctx, LBA_K_MISC, k_cam_index, cdiv(Mp, 256), 256, Mp, m, nc, keys_out, ctx->cam_perm.p, ctx->tidx.p, ctx->cam_ptr.p);
  } else {
    CK(cudaMemsetAsync(ctx->cam_ptr.p, 0, ((size_t)nc + 1) * 4, ctx->stream));
  }
  lap("device slot fill + camera sort (async)");
  // observation constants: AoS (m,3) -> SoA over slots, m_e = n x p_e
  CKR(ctx->cst.alloc((size_t)9 * Mp));
  if (m > 0) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    LAUNCH(ctx, LBA_K_MISC, k_setup_constants, cdiv(Mp, 256), 256, Mp, ctx->slot_src.p, cst_dev[0], cst_dev[1], cst_dev[2], ctx->cst.p);
    CK(cudaStreamSynchronize(ctx->stream));  // the host vectors (tile table, line offsets) go out of scope
  } else {
    CK(cudaStreamSynchronize(ctx->stream));
  }
  lap("constants H2D + transpose + sync");
  // persistent grids: SMs x resident CTAs for each pipelined kernel at this problem's stage size
  {
    auto plan = [&](auto kern, int threads, int smem, int nsub, int* grid) -> int {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int nb = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, (size_t)smem));
      if (nb < 1) return fail(LBA_ERR_UNSUPPORTED, "line-order kernel does not fit on an SM (%d B shared memory)", smem);
      {  // ask for no more shared memory per SM than the resident CTAs need: what is left is L1 for the table gathers
        int smem_sm = 0;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device);
        const long long need = (long long)nb * (smem + 1024);
        if (smem_sm > 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min(100LL, (need * 100 + smem_sm - 1) / smem_sm));
        if (dbg) fprintf(stderr, "[lba_set_problem] plan: %d CTA(s)/SM x %d B shared memory\n", nb, smem);
      }
      *grid = std::max(1, std::min((ntiles + nsub - 1) / nsub, nb * ctx->n_sm));
      return LBA_OK;
    };
    int max_smem = 0;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    // does a kernel variant fit on an SM with this much dynamic shared memory?
    auto fits = [&](auto kern, int threads, int smem) -> bool {
      if (smem > max_smem) return false;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return false; }
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, (size_t)smem) != cudaSuccess) { cudaGetLastError(); return false; }
      if (dbg) fprintf(stderr, "[lba_set_problem] fits? threads %d smem %d -> %d block(s)/SM\n", threads, smem, nb);
      return nb >= 1;
    };
    // preferred: 2 sub-groups per CTA sharing a shared-memory camera table (1 CTA / SM).  Two pipeline stages per
    // sub-group: a third fits (the staging rows live inside the stages) but measured 2.5 % SLOWER on config 4
    // (138.7 vs 135.3 us) -- the mbarrier wait of this kernel is not load latency, and the extra 45 KB of shared
    // memory come out of the L1 that serves its scattered stores; LBA_EVAL_STAGES=3|4 selects them anyway.
    ctx->eval_tab_smem = false;
    int sj = 0, smax = 0;
    const bool no_tab = getenv("LBA_EVAL_TAB") && atoi(getenv("LBA_EVAL_TAB")) == 0;   // force the global-gather variant
    for (int st : {4, 3, 2})
      if (nc > 0 && !smax && !no_tab && fits(k_eval<true, 2, true>, 2 * kTile, eval_smem_bytes(lmax, st, 2, nc)) &&
          fits(k_eval<false, 2, true>, 2 * kTile, eval_smem_bytes(lmax, st, 2, nc))) smax = st;
    if (smax) sj = 2;
    if (const char* e = getenv("LBA_EVAL_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= smax) sj = v; }
    if (sj) {
      ctx->eval_tab_smem = true;
      ctx->eval_stages_j = ctx->eval_stages_r = sj;
      ctx->smem_eval_j = ctx->smem_eval_r = eval_smem_bytes(lmax, sj, 2, nc);
      CKR(plan(k_eval<true, 2, true>, 2 * kTile, ctx->smem_eval_j, 2, &ctx->grid_eval_j));
      CKR(plan(k_eval<false, 2, true>, 2 * kTile, ctx->smem_eval_r, 2, &ctx->grid_eval_r));
    } else {
      ctx->eval_stages_j = ctx->eval_stages_r = 3;
      ctx->smem_eval_j = ctx->smem_eval_r = eval_smem_bytes(lmax, 3, 1, 0);
      CKR(plan(k_eval<true, 1, false>, kTile, ctx->smem_eval_j, 1, &ctx->grid_eval_j));
      CKR(plan(k_eval<false, 1, false>, kTile, ctx->smem_eval_r, 1, &ctx->grid_eval_r));
    }
    ctx->eval_z = !(getenv("LBA_EVAL_Z") && atoi(getenv("LBA_EVAL_Z")) == 0);
    if (ctx->eval_z) {
      if (ctx->eval_tab_smem) {
        if (fits(k_eval_z<2, true>, 2 * kTile, eval_z_smem_bytes(lmax, 2, nc))) {
          ctx->smem_eval_j = eval_z_smem_bytes(lmax, 2, nc);
          CKR(plan(k_eval_z<2, true>, 2 * kTile, ctx->smem_eval_j, 2, &ctx->grid_eval_j));
        } else {
          ctx->eval_z = false;
        }
      } else {
        ctx->smem_eval_j = eval_z_smem_bytes(lmax, 1, 0);
        CKR(plan(k_eval_z<1, false>, kTile, ctx->smem_eval_j, 1, &ctx->grid_eval_j));
      }
    }
    if (dbg) fprintf(stderr, "[lba_set_problem] evaluation kernel: %s, %d stages, %d B shared memory\n",
                     ctx->eval_tab_smem ? "camera table in shared memory" : "camera table gathered from L2", ctx->eval_z ? 2 : ctx->eval_stages_j, ctx->smem_eval_j);
    if (dbg) fprintf(stderr, "[lba_set_problem] Jacobian evaluation: %s\n", ctx->eval_z ? "k_eval_z" : "k_eval<true>");
    {  // one CTA per camera.  The PCG product (6 sums) is HBM-bound on the jc stream and likes 128 threads once
       // there are >= 4 CTAs per SM; the block sums (27-39 sums per camera) pay for their CTA-wide reduction, so
       // they get narrow CTAs (measured, accumulate + finish: 1000 cameras x 2000 obs 31 us at 64 threads vs 35 at 128;
       // 8000 cameras x 250 obs 39 us at 32 threads vs 52 at 128)
      ctx->pcg_threads = nc < 2 * ctx->n_sm ? 512 : (nc < 4 * ctx->n_sm ? 256 : 128);
      const long long per_cam = nc > 0 ? m / nc : 0;
      ctx->cam_threads = nc < 4 * ctx->n_sm ? ctx->pcg_threads : (per_cam >= 512 ? 64 : 32);
      auto knob = [](const char* name, int* v) {
        if (const char* e = getenv(name)) {
          const int x = atoi(e);
          if (x == 32 || x == 64 || x == 128 || x == 256 || x == 512) *v = x;
        }
      };
      knob("LBA_CAM_THREADS", &ctx->cam_threads);
      knob("LBA_PCG_THREADS", &ctx->pcg_threads);
      if (dbg) fprintf(stderr, "[lba_set_problem] camera-kernel CTA sizes: sums %d, PCG %d\n", ctx->cam_threads, ctx->pcg_threads);
    }
    {  // CG vector kernels spin on a grid barrier: the grid must be co-resident
      int nb = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_cg_step, kCgBlock, 0));
      const int want = cdiv((long long)std::max(nc, 1) * 8, kCgBlock);
      ctx->cg_ctas = std::max(1, std::min({want, kCgMaxCtas, std::max(1, nb) * ctx->n_sm}));
      const int ncl = cdiv((long long)std::max(nc, 1) * 8, kCgClBlock);
      ctx->cg_cluster = (ncl <= kCgClMaxCtas && !(getenv("LBA_CG_CLUSTER") && atoi(getenv("LBA_CG_CLUSTER")) == 0)) ? ncl : 0;
      if (dbg) fprintf(stderr, "[lba_set_problem] CG vector kernels: %s\n", ctx->cg_cluster ? "one cluster (single-GPU solves)" : "grid with software barrier");
    }
    ctx->smem_lp[0] = lp_smem_bytes(lmax, kModeMatvec);
    ctx->smem_lp[1] = lp_smem_bytes(lmax, kModeBacksub);
    ctx->smem_lp[2] = lp_smem_bytes(lmax, kModeScalars);
    CKR(plan(k_line_pass<kModeMatvec>, kTile, ctx->smem_lp[0], 1, &ctx->grid_lp[0]));
    CKR(plan(k_line_pass<kModeBacksub>, kTile, ctx->smem_lp[1], 1, &ctx->grid_lp[1]));
    CKR(plan(k_line_pass<kModeScalars>, kTile, ctx->smem_lp[2], 1, &ctx->grid_lp[2]));
  }

  lap("kernel planning");
  // state buffers
  const size_t M = (size_t)Mp, NC = (size_t)nc, NL = (size_t)nl;
  const size_t NG = (size_t)std::max({ctx->grid_eval_j, ctx->grid_eval_r, ctx->grid_lp[1], 1});
  CKR(ctx->cams.alloc(NC * 6)); CKR(ctx->cams_new.alloc(NC * 6)); CKR(ctx->api_tmp.alloc(std::max(NC * 6, (size_t)1)));
  CKR(ctx->pts.alloc(NL * 6 + 2)); CKR(ctx->pts_new.alloc(NL * 6 + 2));
  CKR(ctx->Gm.alloc(NC * 9)); CKR(ctx->Gm_new.alloc(NC * 9)); CKR(ctx->recs.alloc(NC)); CKR(ctx->recs_new.alloc(NC));
  CKR(ctx->r.alloc(M)); CKR(ctx->r_new.alloc(M)); CKR(ctx->jl.alloc(6 * M)); CKR(ctx->jc.alloc(6 * (size_t)m));
  CKR(ctx->jl_new.alloc(6 * M)); CKR(ctx->jc_new.alloc(6 * (size_t)m));
  CKR(ctx->w.alloc(M)); CKR(ctx->wk.alloc(M)); CKR(ctx->wr.alloc(M)); CKR(ctx->dense_tmp.alloc((size_t)m));
  CKR(ctx->V.alloc(NL * kVStride)); CKR(ctx->gl.alloc(NL * 6 + 2)); CKR(ctx->V_new.alloc(NL * kVStride)); CKR(ctx->gl_new.alloc(NL * 6 + 2)); CKR(ctx->Vinv.alloc(NL * kVinvStride + 2));
  CKR(ctx->y.alloc(NL * 6 + 2)); CKR(ctx->Dl.alloc(NL * 6)); CKR(ctx->dl.alloc(NL * 6 + 2));
  CKR(ctx->U.alloc(NC * 27)); CKR(ctx->gc.alloc(NC * 6)); CKR(ctx->Dc.alloc(NC * 6));
  CKR(ctx->Mpart.alloc(NC * 39)); CKR(ctx->Minv.alloc(NC * 21)); CKR(ctx->b.alloc(NC * 6));
  CKR(ctx->cgx.alloc(NC * 6)); CKR(ctx->cgr.alloc(NC * 6)); CKR(ctx->cgz.alloc(NC * 6)); CKR(ctx->cgp.alloc(NC * 6));
  CKR(ctx->tab.alloc(NC * 6)); CKR(ctx->qpart.alloc(NC * 6));
  CKR(ctx->cost_partial.alloc(NG)); CKR(ctx->bs_partial.alloc(NG * 4));
  // k_cam_step: one thread per camera; in a retry pass its CTAs also rebuild the line side of the trial point
  ctx->cam_step_ctas = std::max(cdiv(std::max(nc, 1), kCamStepThreads), std::min(cdiv((long long)nl * 6, kCamStepThreads * 8), 2 * ctx->n_sm));
  CKR(ctx->cam_partial.alloc(2 * (size_t)ctx->cam_step_ctas));
  lap("state buffers");
  ctx->have_problem = true;
  if (ctx->launch_err) return take_launch_err(ctx);
  return LBA_OK;
}

// U and gc (and Mpart/bpart) are stored back to back ([nc][21] then [nc][6]) so that one
// all-reduce covers both.
static inline double* gc_of(lba_ctx* c) { return c->U.p + (size_t)c->nc * 21; }
static inline double* bpart_of(lba_ctx* c) { return c->Mpart.p + (size_t)c->nc * 21; }

// Kernel arguments name BOTH evaluation sets (index 0: the resident one, 1: the "_new" buffers); the kernels
// pick one from the device control block, or set `which` when no LM solve is running (ctl == NULL).
static EvalArgs eval_args(lba_ctx* c, const LmCtl* ctl, int which) {
  EvalArgs a;
  a.Mp = c->Mp; a.ntiles = c->ntiles; a.lmax = c->lmax; a.nc = c->nc; a.nstages = 3;
  a.tidx = c->tidx.p; a.cst = c->cst.p;
  a.recs = dual<const CamRec>(c->recs.p, c->recs_new.p);
  a.pts = dual<const double>(c->pts.p, c->pts_new.p);
  a.r = dual(c->r.p, c->r_new.p); a.jl = dual(c->jl.p, c->jl_new.p); a.jc = dual(c->jc.p, c->jc_new.p);
  a.V = dual(c->V.p, c->V_new.p); a.gl = dual(c->gl.p, c->gl_new.p); a.gbits = dual(c->gbits.p, c->gbits_new.p);
  a.ctl = ctl; a.which = which; a.fragV = c->fragV.p;
  a.cost_partial = c->cost_partial.p;
  a.ticket = c->tickets.p; a.cost_out = &scal_local(c)->cost;
  a.loss = c->loss; a.f_scale = c->f_scale;
  return a;
}

static int load_state(lba_ctx* c, const double* cams_api, const double* pts_api) {
  if (!c->have_problem) return fail(LBA_ERR_STATE, "lba_set_problem has not been called");
  if (!cams_api || !pts_api) return fail(LBA_ERR_ARG, "NULL parameter array");
  CK(cudaSetDevice(c->device));
  if (c->nc > 0) {
    CK(cudaMemcpyAsync(c->api_tmp.p, cams_api, (size_t)c->nc * 48, cudaMemcpyDefault, c->stream));
    LAUNCH(c, LBA_K_MISC, k_cams_from_api, cdiv((long long)c->nc * 6, 256), 256, c->nc, c->api_tmp.p, c->cams.p);
    LAUNCH(c, LBA_K_MISC, k_cam_prepare, cdiv(c->nc, 128), 128, c->nc, c->cams.p, c->recs.p, c->Gm.p);
  }
  if (c->nl > 0) CK(cudaMemcpyAsync(c->pts.p, pts_api, (size_t)c->nl * 48, cudaMemcpyDefault, c->stream));
  c->have_eval = c->have_blocks = false;
  return LBA_OK;
}

static int store_state(lba_ctx* c, double* cams_api, double* pts_api) {
  if (c->nc > 0) {
    LAUNCH(c, LBA_K_MISC, k_cams_to_api, cdiv((long long)c->nc * 6, 256), 256, c->nc, c->cams.p, c->api_tmp.p);
    CK(cudaMemcpyAsync(cams_api, c->api_tmp.p, (size_t)c->nc * 48, cudaMemcpyDefault, c->stream));
  }
  if (c->nl > 0) CK(cudaMemcpyAsync(pts_api, c->pts.p, (size_t)c->nl * 48, cudaMemcpyDefault, c->stream));
  return LBA_OK;
}

// residual-only pass over evaluation set `which` (API calls: ctl == NULL) -> r of that set, scal.cost (local sum)
static int run_residual(lba_ctx* c, const LmCtl* ctl, int which) {
  if (c->ntiles > 0) {
    EvalArgs a = eval_args(c, ctl, which);
    a.nstages = c->eval_stages_r;
    if (c->eval_tab_smem) LAUNCH_DYN(c, LBA_K_RESIDUAL, (k_eval<false, 2, true>), c->grid_eval_r, 2 * kTile, c->smem_eval_r, a);
    else LAUNCH_DYN(c, LBA_K_RESIDUAL, (k_eval<false, 1, false>), c->grid_eval_r, kTile, c->smem_eval_r, a);
  } else {
    CK(cudaMemsetAsync(&scal_local(c)->cost, 0, 8, c->stream));
  }
  return LBA_OK;
}

// fused residual + Jacobian + line blocks (+ line part of the gradient norm) of evaluation set `which` (relative
// to the control block's current set when ctl is given) -> scal.cost and r, jl, jc, V, gl, gbits of that set.
// The gradient-norm word of the set must be zero before: the API path clears it here, inside the LM loop
// k_cam_step clears the trial set's word.
static int run_jacobian(lba_ctx* c, const LmCtl* ctl, int which) {
  if (!ctl) CK(cudaMemsetAsync(which ? c->gbits_new.p : c->gbits.p, 0, 8, c->stream));
  if (c->ntiles > 0) {
    EvalArgs a = eval_args(c, ctl, which);
    a.nstages = c->eval_stages_j;
    if (c->eval_z) {
      if (c->eval_tab_smem) LAUNCH_DYN(c, LBA_K_JACOBIAN, (k_eval_z<2, true>), c->grid_eval_j, 2 * kTile, c->smem_eval_j, a);
      else LAUNCH_DYN(c, LBA_K_JACOBIAN, (k_eval_z<1, false>), c->grid_eval_j, kTile, c->smem_eval_j, a);
    } else if (c->eval_tab_smem) LAUNCH_DYN(c, LBA_K_JACOBIAN, (k_eval<true, 2, true>), c->grid_eval_j, 2 * kTile, c->smem_eval_j, a);
    else LAUNCH_DYN(c, LBA_K_JACOBIAN, (k_eval<true, 1, false>), c->grid_eval_j, kTile, c->smem_eval_j, a);
    if (c->n_long > 0)   // lines with more than kTile observations: add up their fragments' blocks
      LAUNCH(c, LBA_K_JACOBIAN, k_frag_blocks, c->n_long, 32, c->long_lines(), (const double*)c->fragV.p, dual(c->V.p, c->V_new.p),
             dual(c->gl.p, c->gl_new.p), dual(c->gbits.p, c->gbits_new.p), ctl, which);
  } else {
    CK(cudaMemsetAsync(&scal_local(c)->cost, 0, 8, c->stream));
  }
  if (!ctl && which == 0) {
    c->pending_first = !c->have_scaling;
    c->line_scaling_mode = c->pending_first ? 2 : 1;
    c->have_eval = true;
    c->have_blocks = false;
  }
  return LBA_OK;
}

template <typename T>
static inline void swap_buf(DBuf<T>& a, DBuf<T>& b) { std::swap(a.p, b.p); std::swap(a.n, b.n); }

// after an LM solve that ended with set 1 current: make the host's names agree with the device again
static void swap_sets(lba_ctx* c) {
  swap_buf(c->cams, c->cams_new); swap_buf(c->pts, c->pts_new);
  swap_buf(c->recs, c->recs_new); swap_buf(c->Gm, c->Gm_new);
  swap_buf(c->r, c->r_new); swap_buf(c->jl, c->jl_new); swap_buf(c->jc, c->jc_new);
  swap_buf(c->V, c->V_new); swap_buf(c->gl, c->gl_new); swap_buf(c->gbits, c->gbits_new);
}

// camera side of the normal equations (stand-alone API calls): diag(U_c) (or the full blocks) and g_c,
// all-reduced; camera scaling; camera part of the gradient inf-norm
static int run_blocks(lba_ctx* c, bool full) {
  if (c->nc > 0) {
    const size_t n27 = (size_t)c->nc * 27;
    PeerLink pl;
    memset(&pl, 0, sizeof pl);
    if (full) {
      // the complete blocks, reduced into U | gc by a separate reduce
      double* out = reduce_target(c, kRgnU, c->U.p, n27);
      c->cam_threads_sel = c->cam_threads;
      LAUNCH_CAM2(c, LBA_K_CAM_BLOCKS, k_cam_accum, false, c->nc, c->cam_ptr.p, c->cam_perm.p,
                  dual<const double>(c->jc.p, c->jc_new.p), (const double*)nullptr, c->r.p,
                  dual<const double>(nullptr, nullptr), out, out + (size_t)c->nc * 21, (double*)nullptr, c->nc, pl,
                  (const LmCtl*)nullptr);
      CKR(reduce_finish(c, kRgnU, c->U.p, n27));
    } else {
      // diag(U) and g_c only (zero-copy producer + one reduce kernel with the peer transport)
      double* out = reduce_target(c, kRgnU, c->U.p, n27);
      if (out != c->U.p) CK(cudaMemsetAsync(out, 0, n27 * 8, c->stream));   // k_cam_grad writes the diagonal only
      c->cam_threads_sel = c->cam_threads;
      LAUNCH_CAM(c, LBA_K_CAM_BLOCKS, k_cam_grad, c->nc, c->cam_ptr.p, c->cam_perm.p, c->jc.p, c->r.p, out,
                 out + (size_t)c->nc * 21, pl);
      CKR(reduce_finish(c, kRgnU, c->U.p, n27));
    }
    LAUNCH(c, LBA_K_MISC, k_cam_post, cdiv((long long)c->nc * 6, 256), 256, c->nc, c->U.p, gc_of(c), c->gc.p, c->Dc.p,
           c->pending_first ? 1 : 0, c->gbits.p, pl);
  }
  c->have_scaling = true;
  CKR(allreduce(c, c->gbits.p, 1, ncclUint64, ncclMax));
  c->have_blocks = true;
  return LBA_OK;
}

// ---- the damped solve, in three stream-ordered pieces -------------------------------------------
// With `ctl` (the LM loop) the kernels take the damping, the current evaluation set and whether to run at all
// from the device control block; without it (lba_schur_solve) from the arguments.
struct SolvePlan {
  CgArgs ca;
  LinePassArgs la;
  int issued = 0;
};

// (1) Schur elimination set-up for damping lam and the first CG state.  with_grad: the camera gradient and
// Marquardt scaling of the evaluation set ride on the preconditioner pass (one read of jc).
static int solve_setup(lba_ctx* c, const LmCtl* ctl, double lam, double pcg_tol, int pcg_max_iter, SolvePlan* sp, bool with_grad) {
  const int nc = c->nc, nl = c->nl;
  if (nl > 0) {
    LAUNCH(c, LBA_K_LINE_FACTOR, k_line_factor, cdiv(nl, kLfLines), kLfLines, nl, lam, dual<const double>(c->V.p, c->V_new.p),
           c->Dl.p, dual<const double>(c->gl.p, c->gl_new.p), c->Vinv.p, c->y.p, c->clamped.p, c->line_scaling_mode, ctl);
    c->line_scaling_mode = 0;
  }
  LinePassArgs& la = sp->la;
  la.Mp = c->Mp; la.ntiles = c->ntiles; la.lmax = c->lmax; la.tidx = c->tidx.p;
  la.jl = dual<const double>(c->jl.p, c->jl_new.p); la.pts = dual(c->pts.p, c->pts_new.p);
  la.tab = c->tab.p; la.Vinv = c->Vinv.p;
  la.w = c->wk.p; la.w2 = c->wr.p; la.y = c->y.p;
  la.gl = dual<const double>(c->gl.p, c->gl_new.p); la.r = dual<const double>(c->r.p, c->r_new.p);
  la.dl = c->dl.p; la.alpha = 1.0;
  la.partial = c->bs_partial.p; la.ticket = c->tickets.p + 1; la.out4 = &scal_local(c)->rjd; la.st = c->cgst.p; la.ctl = ctl; la.fragT = c->fragT.p;
  if (c->ntiles > 0)
    LAUNCH_DYN(c, LBA_K_LINE_SCALARS, k_line_pass<kModeScalars>, c->grid_lp[2], kTile, c->smem_lp[2], la);
  la.w = c->w.p;
  if (nc > 0) {
    // 27 (39 with the gradient sums) values per camera and a thread-per-camera consumer: a separate,
    // massively parallel reduce kernel (zero-copy producer) beats fusing the peer loads into k_cam_finish
    const size_t nred = (size_t)nc * (with_grad ? 39 : 27);
    PeerLink pl;
    memset(&pl, 0, sizeof pl);
    double* out = reduce_target(c, kRgnM, c->Mpart.p, nred);
    c->cam_threads_sel = c->cam_threads;
    const Dual<const double> jc2 = dual<const double>(c->jc.p, c->jc_new.p);
    if (with_grad) {
      LAUNCH_CAM2(c, LBA_K_CAM_PRECOND, k_cam_accum, true, nc, c->cam_ptr.p, c->cam_perm.p, jc2,
                  (const double*)c->wk.p, c->wr.p, dual<const double>(c->r.p, c->r_new.p), out, out + (size_t)nc * 21,
                  out + (size_t)nc * 27, nc, pl, ctl);
    } else {
      LAUNCH_CAM2(c, LBA_K_CAM_PRECOND, k_cam_accum, false, nc, c->cam_ptr.p, c->cam_perm.p, jc2,
                  (const double*)c->wk.p, c->wr.p, dual<const double>(nullptr, nullptr), out, out + (size_t)nc * 21,
                  (double*)nullptr, nc, pl, ctl);
    }
    CKR(reduce_finish(c, kRgnM, c->Mpart.p, nred));
    LAUNCH(c, LBA_K_CAM_PRECOND, k_cam_finish, cdiv(nc, 128), 128, nc, lam, c->Mpart.p, c->Dc.p, c->gc.p, bpart_of(c),
           with_grad ? (const double*)(c->Mpart.p + (size_t)nc * 27) : (const double*)nullptr,
           dual(c->gbits.p, c->gbits_new.p), c->Minv.p, c->b.p, c->clamped.p, pl, ctl);
    if (with_grad) c->have_scaling = true;   // the caller all-reduces gbits (max) together with its next scalars
  }
  CgArgs& ca = sp->ca;
  ca.nc = nc; ca.lam = lam; ca.tol2 = pcg_tol * pcg_tol; ca.max_iter = pcg_max_iter;
  ca.Dc = c->Dc.p; ca.Minv = c->Minv.p; ca.Gm = dual<const double>(c->Gm.p, c->Gm_new.p);
  ca.cams = dual<const double>(c->cams.p, c->cams_new.p); ca.ctl = ctl; ca.b = c->b.p; ca.qpart = c->qpart.p;
  ca.x = c->cgx.p; ca.r = c->cgr.p; ca.z = c->cgz.p; ca.p = c->cgp.p; ca.tab = c->tab.p; ca.st = c->cgst.p;
  ca.partial = c->cg_partial.p; ca.gb = c->gridbar.p;
  memset(&ca.pl, 0, sizeof ca.pl);
  if (c->cg_cluster && c->world == 1) {
    c->next_cluster = c->cg_cluster;
    LAUNCH(c, LBA_K_PCG_UPDATE, k_cg_begin_cl, c->cg_cluster, kCgClBlock, ca);
  } else {
    LAUNCH(c, LBA_K_PCG_UPDATE, k_cg_begin, c->cg_ctas, kCgBlock, ca);
  }
  sp->issued = 0;
  return LBA_OK;
}

static int fetch_scalars(lba_ctx* c);
static int publish_scalars(lba_ctx* c);
static int wait_scalars(lba_ctx* c);

// (2) n more PCG iterations (kernels return at once when the device-side `done` flag is set)
static int solve_iterate(lba_ctx* c, SolvePlan* sp, int n) {
  for (int q = 0; q < n && sp->issued < sp->ca.max_iter; ++q, ++sp->issued) {
    if (c->profiling) {
      // per-class timing: look at the CG flag before every step, so that no early-exit launch (a few us each)
      // is counted as a PCG kernel and dilutes the averages
      CKR(fetch_scalars(c));
      if (c->h_cgst->done) break;
    }
    if (c->ntiles > 0) {
      c->window(c->jl.p, (size_t)c->ntiles * 6 * kTile * 8, c->l2_persist / 2);
      LAUNCH_DYN(c, LBA_K_PCG_LINE, k_line_pass<kModeMatvec>, c->grid_lp[0], kTile, c->smem_lp[0], sp->la);
      if (c->n_long > 0)
        LAUNCH(c, LBA_K_PCG_LINE, k_frag_matvec, c->n_long, kTile, c->long_lines(), (const TileIdx*)c->tidx.p, sp->la.jl,
               (const double*)c->Vinv.p, (const double*)c->fragT.p, c->w.p, (const CgState*)c->cgst.p, sp->la.ctl);
    }
    PeerLink pl;
    memset(&pl, 0, sizeof pl);
    if (c->nc > 0) {
      const size_t n6 = (size_t)c->nc * 6;
      // fused all-reduce with the peer transport: k_pcg_cam writes this rank's slot and signals,
      // k_cg_step waits and adds the slots of all ranks
      double* out = fused_link(c, kRgnQ, n6, c->qpart.p, &pl);
      c->window(c->jc.p, (size_t)c->m * 48, c->l2_persist / 2);
      c->cam_threads_sel = c->pcg_threads;
      LAUNCH_CAM(c, LBA_K_PCG_CAM, k_pcg_cam, std::min(c->nc, 8 * c->n_sm), c->nc, c->cam_ptr.p, c->cam_perm.p,
                 dual<const double>(c->jc.p, c->jc_new.p), c->w.p, out, c->cgst.p, pl, sp->la.ctl);
      if (!pl.seq) CKR(reduce_finish(c, kRgnQ, c->qpart.p, n6));
    }
    sp->ca.pl = pl;
    if (c->cg_cluster && c->world == 1) {
      c->next_cluster = c->cg_cluster;
      LAUNCH(c, LBA_K_PCG_UPDATE, k_cg_step_cl, c->cg_cluster, kCgClBlock, sp->ca);
    } else {
      LAUNCH(c, LBA_K_PCG_UPDATE, k_cg_step, c->cg_ctas, kCgBlock, sp->ca);
    }
  }
  return LBA_OK;
}

// the trial point on the camera side (and, in a retry pass of the LM loop, on the line side too)
static int launch_cam_step(lba_ctx* c, const LmCtl* ctl) {
  LAUNCH(c, LBA_K_MISC, k_cam_step, c->cam_step_ctas, kCamStepThreads, c->nc, dual(c->cams.p, c->cams_new.p), c->cgx.p,
         dual(c->recs.p, c->recs_new.p), dual(c->Gm.p, c->Gm_new.p), 1.0, dual(c->gbits.p, c->gbits_new.p),
         (size_t)c->nl * 6, dual(c->pts.p, c->pts_new.p), c->dl.p, c->cam_partial.p, c->tickets.p + 2, &c->scal.p->dc2, ctl);
  return LBA_OK;
}

// (3) back-substitution and the candidate point (k_cg_step left the gather table of the solution)
static int solve_finish(lba_ctx* c, SolvePlan* sp) {
  if (c->ntiles > 0) {
    LAUNCH_DYN(c, LBA_K_BACKSUB, k_line_pass<kModeBacksub>, c->grid_lp[1], kTile, c->smem_lp[1], sp->la);
    if (c->n_long > 0)
      LAUNCH(c, LBA_K_BACKSUB, k_frag_backsub, c->n_long, kTile, c->long_lines(), (const TileIdx*)c->tidx.p, sp->la.jl,
             (const double*)c->Vinv.p, sp->la.gl, sp->la.r, sp->la.pts, c->dl.p, (const double*)c->fragT.p, (const double*)c->w.p,
             1.0, c->frag_partial.p, c->tickets.p + 5, sp->la.out4, sp->la.ctl);
  } else {
    if (c->nl > 0) CK(cudaMemcpyAsync(c->pts_new.p, c->pts.p, (size_t)c->nl * 48, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemsetAsync(&scal_local(c)->rjd, 0, 32, c->stream));
  }
  return launch_cam_step(c, sp->la.ctl);
}

// the three pieces with a host check of the CG flag in between (stand-alone lba_schur_solve)
static int run_solve(lba_ctx* c, double lam, double pcg_tol, int pcg_max_iter, int* chunk_hint, int* iters_out,
                     int* breakdown_out) {
  SolvePlan sp;
  CKR(solve_setup(c, nullptr, lam, pcg_tol, pcg_max_iter, &sp, false));
  int chunk = std::max(1, std::min(*chunk_hint, pcg_max_iter));
  while (true) {
    CKR(solve_iterate(c, &sp, chunk));
    CKR(fetch_scalars(c));
    if (c->h_cgst->done || sp.issued >= pcg_max_iter) break;
    chunk = 4;
  }
  *iters_out = c->h_cgst->k;
  *breakdown_out = c->h_cgst->breakdown;
  *chunk_hint = c->h_cgst->k + 1;
  return solve_finish(c, &sp);
}

// The host's look at the device scalars (LM control block, LM scalars, CG state, gradient-norm bits, clamp count):
// k_publish writes them into mapped pinned memory followed by a sequence number; the host spins on that word
// and only falls back to the stream when the device reports an error.
static int publish_scalars(lba_ctx* c) {
  const unsigned long long seq = ++c->publish_seq;
  LAUNCH(c, LBA_K_MISC, k_publish, 1, 32, (const Scalars*)c->scal.p, (const CgState*)c->cgst.p, (const LmCtl*)c->ctl.p,
         dual<const unsigned long long>(c->gbits.p, c->gbits_new.p), (const int*)c->clamped.p, c->d_mirror, seq);
  return LBA_OK;
}
// wait for the newest publication (kernels enqueued behind it keep the GPU busy meanwhile)
static int wait_scalars(lba_ctx* c) {
  const unsigned long long seq = c->publish_seq;
  volatile unsigned long long* flag = &c->h_mirror->seq;
  for (unsigned spins = 0; *flag != seq; ++spins) {
    if ((spins & 0x3ff) == 0x3ff) {
      const cudaError_t e = cudaStreamQuery(c->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) return fail(LBA_ERR_CUDA, "device error while waiting for the iteration scalars: %s", cudaGetErrorString(e));
      if (e == cudaSuccess && *flag != seq) {   // stream drained: the write has landed by now
        CK(cudaStreamSynchronize(c->stream));
        break;
      }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (c->h_mirror->spin_error) {
    const unsigned int code = c->h_mirror->spin_error, zero = 0;
    cudaStreamSynchronize(c->stream);
    cudaMemcpyToSymbol(g_spin_error, &zero, sizeof zero);
    c->h_mirror->spin_error = 0;
    return fail(LBA_ERR_COMM, "a device-side wait timed out (%s%s%s%s): a peer rank stalled or died, or the GPU is shared and a "
                "co-resident grid could not be scheduled; results of this call are invalid",
                (code & kSpinPeerLink) ? "fused peer exchange " : "", (code & kSpinGridBar) ? "grid barrier " : "",
                (code & kSpinPeerReduce) ? "peer all-reduce " : "", (code & kSpinPeerReduceRs) ? "two-phase peer reduce" : "");
  }
  if (c->launch_err) return take_launch_err(c);
  return LBA_OK;
}

static int fetch_scalars(lba_ctx* c) {
  CKR(publish_scalars(c));
  return wait_scalars(c);
}

static int export_residuals(lba_ctx* c, const double* r_dev, double* out) {
  if (!out || c->m == 0) return LBA_OK;
  // slot order -> the caller's observation order
  LAUNCH(c, LBA_K_MISC, k_slots_to_caller, cdiv(c->Mp, 256), 256, c->Mp, c->slot_src.p, r_dev, c->dense_tmp.p);
  CK(cudaMemcpyAsync(out, c->dense_tmp.p, (size_t)c->m * 8, cudaMemcpyDefault, c->stream));
  return LBA_OK;
}

// ---- C ABI: evaluation ---------------------------------------------------------------------------
extern "C" int lba_point_wise_indices(lba_ctx* ctx, int64_t* cam_pt, int64_t* pt_idx) {
  if (!ctx || !ctx->have_problem) return fail(LBA_ERR_STATE, "lba_set_problem has not been called");
  if (!cam_pt || !pt_idx) return fail(LBA_ERR_ARG, "NULL output");
  CK(cudaSetDevice(ctx->device));
  const int m = ctx->m;
  if (m == 0) return LBA_OK;
  DBuf<long long> tmp;
  CKR(tmp.alloc((size_t)4 * m));
  LAUNCH(ctx, LBA_K_MISC, k_point_wise_indices, cdiv(ctx->Mp, 256), 256, ctx->Mp, ctx->slot_src.p, ctx->tidx.p,
         tmp.p, tmp.p + (size_t)2 * m);
  CK(cudaMemcpyAsync(cam_pt, tmp.p, (size_t)2 * m * 8, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(pt_idx, tmp.p + (size_t)2 * m, (size_t)2 * m * 8, cudaMemcpyDefault, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

extern "C" int lba_project(lba_ctx* ctx, const double* cams, const double* pts, double* proj) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (!proj) return fail(LBA_ERR_ARG, "NULL output");
  CKR(load_state(ctx, cams, pts));
  const int m = ctx->m;
  if (m == 0) return LBA_OK;
  DBuf<double> tmp;
  CKR(tmp.alloc((size_t)6 * m));
  ProjectArgs a;
  a.Mp = ctx->Mp; a.slot_src = ctx->slot_src.p; a.tidx = ctx->tidx.p; a.recs = ctx->recs.p; a.pts = ctx->pts.p;
  a.proj = tmp.p;
  LAUNCH(ctx, LBA_K_MISC, k_project, cdiv(ctx->Mp, 256), 256, a);
  CK(cudaMemcpyAsync(proj, tmp.p, (size_t)6 * m * 8, cudaMemcpyDefault, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

// Track table of `n_views` consecutive views (tracker.py:11-96 with verify_duplet, utils.py:1807-1944): the view
// pairs are processed in order on the context's stream, each pair data-parallel, with no host synchronisation
// until the tables are copied out.  match_ptr is read on the host (launch sizes).
extern "C" int lba_track_table(lba_ctx* ctx, int32_t n_views, int64_t n_lines2d, const double* endpoints,
                               const int64_t* match_ptr, const int64_t* matches, const double* F, double threshold,
                               const uint8_t* flags_in, int64_t cap_rows, int64_t* ids, int64_t* bin,
                               int64_t* n_new_tracks, int64_t* n_tracks, uint8_t* flags_out) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (!endpoints && ctx->l2d_end.p) { endpoints = ctx->l2d_end.p; n_lines2d = ctx->n_lines2d; }   // resident table (lba_lines2d_set)
  if (n_views < 2 || n_lines2d < 1) return fail(LBA_ERR_ARG, "need at least two views and one 2D line");
  if (!endpoints || !match_ptr || !F || !ids || !bin || !n_new_tracks || !n_tracks) return fail(LBA_ERR_ARG, "NULL array");
  if (is_device_ptr(match_ptr) || is_device_ptr(n_new_tracks) || is_device_ptr(n_tracks))
    return fail(LBA_ERR_ARG, "match_ptr, n_new_tracks and n_tracks are host arrays");
  const int V = n_views, NP = n_views - 1;
  if (match_ptr[0] != 0) return fail(LBA_ERR_ARG, "match_ptr[0] must be 0");
  for (int v = 0; v < NP; ++v)
    if (match_ptr[v + 1] < match_ptr[v]) return fail(LBA_ERR_ARG, "match_ptr must be non-decreasing");
  const long long M = match_ptr[NP];
  if (M > 0x7fffffffLL || n_lines2d > 0x7fffffffLL) return fail(LBA_ERR_ARG, "too many matches or lines");
  if (M > 0 && !matches) return fail(LBA_ERR_ARG, "NULL matches");
  if (cap_rows < M) return fail(LBA_ERR_ARG, "cap_rows (%lld) must be at least the number of matches (%lld)", (long long)cap_rows, M);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st_ = ctx->stream;
  long long kmax = 0;
  for (int v = 0; v < NP; ++v) kmax = std::max<long long>(kmax, match_ptr[v + 1] - match_ptr[v]);

  DBuf<double> sE, sF; DBuf<long long> sM, d_inl, d_st, d_ids, d_bin, d_nnew; DBuf<unsigned char> sFl, d_present, d_fl8;
  DBuf<int> d_flag, d_pos, d_newflag, d_newpos, d_last, d_bad; DBuf<unsigned char> d_tmp;
  auto on_device = [&](auto& staging, const auto* src, size_t count, auto** out) -> int {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(src)>>;
    if (is_device_ptr(src)) { *out = src; return LBA_OK; }
    CKR(staging.alloc(count));
    CK(cudaMemcpyAsync(staging.p, src, count * sizeof(T), cudaMemcpyHostToDevice, st_));
    *out = staging.p;
    return LBA_OK;
  };
  const double *dE = nullptr, *dF = nullptr; const long long* dM = nullptr; const unsigned char* dFl = nullptr;
  CKR(on_device(sE, endpoints, (size_t)n_lines2d * 4, &dE));
  CKR(on_device(sF, F, (size_t)NP * 9, &dF));
  if (M > 0) CKR(on_device(sM, reinterpret_cast<const long long*>(matches), (size_t)M * 2, &dM));
  if (flags_in && M > 0) CKR(on_device(sFl, flags_in, (size_t)M, &dFl));
  const size_t tab = (size_t)cap_rows * V;
  const bool ids_dev = is_device_ptr(ids), bin_dev = is_device_ptr(bin);
  long long* o_ids = reinterpret_cast<long long*>(ids); long long* o_bin = reinterpret_cast<long long*>(bin);
  if (!ids_dev) { CKR(d_ids.alloc(tab)); o_ids = d_ids.p; }
  if (!bin_dev) { CKR(d_bin.alloc(tab)); o_bin = d_bin.p; }
  CK(cudaMemsetAsync(o_ids, 0, tab * 8, st_));
  CK(cudaMemsetAsync(o_bin, 0, tab * 8, st_));
  CKR(d_flag.alloc((size_t)M)); CKR(d_pos.alloc((size_t)kmax)); CKR(d_newflag.alloc((size_t)kmax)); CKR(d_newpos.alloc((size_t)kmax));
  CKR(d_inl.alloc((size_t)kmax)); CKR(d_present.alloc((size_t)n_lines2d)); CKR(d_last.alloc((size_t)n_lines2d));
  CKR(d_st.alloc(4)); CKR(d_nnew.alloc((size_t)NP)); CKR(d_bad.alloc(1));
  CK(cudaMemsetAsync(d_st.p, 0, 32, st_));
  CK(cudaMemsetAsync(d_nnew.p, 0, (size_t)NP * 8, st_));
  CK(cudaMemsetAsync(d_bad.p, 0, 4, st_));
  size_t tmp_bytes = 0;
  if (kmax > 0) {
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag.p, d_pos.p, (int)kmax, st_));
    CKR(d_tmp.alloc(tmp_bytes));
  }
  const int row_grid = ctx->n_sm * 4;
  for (int v = 0; v < NP; ++v) {
    const long long off = match_ptr[v], k = match_ptr[v + 1] - off;
    if (k == 0) continue;   // n_new_tracks[v] stays 0
    const int grid = cdiv(k, 256);
    int* flag = d_flag.p + off;
    LAUNCH(ctx, LBA_K_MISC, k_tt_verify, grid, 256, k, (long long)n_lines2d, dM + 2 * off, dE, dF + 9 * v, threshold,
           dFl ? dFl + off : (const unsigned char*)nullptr, flag, d_bad.p);
    CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, flag, d_pos.p, (int)k, st_));
    if (v > 0) {
      CK(cudaMemsetAsync(d_present.p, 0, (size_t)n_lines2d, st_));
      CK(cudaMemsetAsync(d_last.p, 0xff, (size_t)n_lines2d * 4, st_));
      LAUNCH(ctx, LBA_K_MISC, k_tt_present, row_grid, 256, (const long long*)d_st.p, V, v, (const long long*)o_ids, d_present.p);
    }
    LAUNCH(ctx, LBA_K_MISC, k_tt_last, grid, 256, k, v, dM + 2 * off, (const int*)flag, (const int*)d_pos.p,
           (const unsigned char*)d_present.p, d_inl.p, d_last.p, d_newflag.p);
    if (v > 0)
      LAUNCH(ctx, LBA_K_MISC, k_tt_extend, row_grid, 256, (const long long*)d_st.p, V, v, o_ids, o_bin, (const int*)d_last.p,
             (const long long*)d_inl.p);
    CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_newflag.p, d_newpos.p, (int)k, st_));
    LAUNCH(ctx, LBA_K_MISC, k_tt_append, grid, 256, k, V, v, (const long long*)d_st.p, dM + 2 * off, (const int*)d_newflag.p,
           (const int*)d_newpos.p, o_ids, o_bin);
    LAUNCH(ctx, LBA_K_MISC, k_tt_advance, 1, 32, k, v, d_st.p, (const int*)d_newflag.p, (const int*)d_newpos.p, d_nnew.p);
  }
  long long rows = 0; int bad = 0;
  CK(cudaMemcpyAsync(&rows, d_st.p, 8, cudaMemcpyDeviceToHost, st_));
  CK(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, st_));
  CK(cudaMemcpyAsync(n_new_tracks, d_nnew.p, (size_t)NP * 8, cudaMemcpyDeviceToHost, st_));
  CK(cudaStreamSynchronize(st_));
  if (ctx->launch_err) return take_launch_err(ctx);
  if (bad) return fail(LBA_ERR_ARG, "%d match(es) name a 2D line outside [0, %lld)", bad, (long long)n_lines2d);
  *n_tracks = rows;
  if (!ids_dev && rows > 0) CK(cudaMemcpyAsync(ids, o_ids, (size_t)rows * V * 8, cudaMemcpyDeviceToHost, st_));
  if (!bin_dev && rows > 0) CK(cudaMemcpyAsync(bin, o_bin, (size_t)rows * V * 8, cudaMemcpyDeviceToHost, st_));
  if (flags_out && M > 0) {
    CKR(d_fl8.alloc((size_t)M));
    LAUNCH(ctx, LBA_K_MISC, k_tt_flags_u8, cdiv(M, 256), 256, M, (const int*)d_flag.p, d_fl8.p);
    CK(cudaMemcpyAsync(flags_out, d_fl8.p, (size_t)M, cudaMemcpyDefault, st_));
  }
  CK(cudaStreamSynchronize(st_));
  return take_launch_err(ctx);
}

// Two-view endpoint triangulation of n_tracks line tracks (seq_utils.py:75-135): independent of the BA problem
// held by the context; uses its stream and device.  `seconds_kernel` (optional) returns the device time of the
// kernel alone (CUDA events on the stream), for the measurement row of DESIGN.md.
extern "C" int lba_triangulate_lines(lba_ctx* ctx, int64_t n_tracks, int64_t n_views, const double* P, const int64_t* view0,
                                     const int64_t* view1, const float* pts0, const float* pts1, float* X, float* Xh,
                                     double* seconds_kernel) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (n_tracks < 0 || n_views < 0 || n_tracks > (1LL << 40)) return fail(LBA_ERR_ARG, "bad sizes (n_tracks=%lld n_views=%lld)", (long long)n_tracks, (long long)n_views);
  if (seconds_kernel) *seconds_kernel = 0.0;
  if (n_tracks == 0) return LBA_OK;
  if (!P || !view0 || !pts0 || !pts1 || !X) return fail(LBA_ERR_ARG, "NULL array");
  if (n_views < 2) return fail(LBA_ERR_ARG, "triangulation needs at least two views");
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)n_tracks;
  // inputs that are not already on the device are staged; outputs are produced on the device and copied out
  DBuf<double> sP; DBuf<long long> sv0, sv1; DBuf<float> sp0, sp1, dX, dXh; DBuf<int> dbad;
  auto on_device = [&](auto& staging, const auto* src, size_t count, auto** out) -> int {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(src)>>;
    if (is_device_ptr(src)) { *out = src; return LBA_OK; }
    CKR(staging.alloc(count));
    CK(cudaMemcpyAsync(staging.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = staging.p;
    return LBA_OK;
  };
  const double* dP = nullptr; const long long *dv0 = nullptr, *dv1 = nullptr; const float *dp0 = nullptr, *dp1 = nullptr;
  CKR(on_device(sP, P, (size_t)n_views * 12, &dP));
  CKR(on_device(sv0, reinterpret_cast<const long long*>(view0), n, &dv0));
  if (view1) CKR(on_device(sv1, reinterpret_cast<const long long*>(view1), n, &dv1));
  CKR(on_device(sp0, pts0, n * 4, &dp0));
  CKR(on_device(sp1, pts1, n * 4, &dp1));
  const bool x_dev = is_device_ptr(X), xh_dev = Xh && is_device_ptr(Xh);
  float* outX = X; float* outXh = Xh;
  if (!x_dev) { CKR(dX.alloc(n * 6)); outX = dX.p; }
  if (Xh && !xh_dev) { CKR(dXh.alloc(n * 8)); outXh = dXh.p; }
  CKR(dbad.alloc(1));
  CK(cudaMemsetAsync(dbad.p, 0, 4, ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  const long long n2 = 2 * (long long)n;
  LAUNCH(ctx, LBA_K_MISC, k_triangulate, cdiv(n2, 256), 256, n2, (int)n_views, dP, dv0, dv1, reinterpret_cast<const float2*>(dp0),
         reinterpret_cast<const float2*>(dp1), outX, reinterpret_cast<float4*>(outXh), dbad.p);
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!x_dev) CK(cudaMemcpyAsync(X, outX, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (Xh && !xh_dev) CK(cudaMemcpyAsync(Xh, outXh, n * 8 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, dbad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (seconds_kernel) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *seconds_kernel = ms * 1e-3;
  }
  if (ctx->launch_err) return take_launch_err(ctx);
  if (bad) return fail(LBA_ERR_ARG, "%d track(s) name a view outside [0, %lld)", bad, (long long)n_views);
  return LBA_OK;
}

// ---- C ABI: BA input resident on the device across the incremental loop ---------------------------
extern "C" int lba_lines2d_set(lba_ctx* ctx, int64_t n_lines2d, const double* n, const double* p0, const double* p1,
                               const double* endpoints) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (n_lines2d < 0 || n_lines2d > 0x7fffffffLL / 4) return fail(LBA_ERR_ARG, "bad number of 2D lines");
  if (n_lines2d > 0 && (!n || !p0 || !p1)) return fail(LBA_ERR_ARG, "NULL array");
  CK(cudaSetDevice(ctx->device));
  const size_t N = (size_t)n_lines2d;
  CKR(ctx->l2d_n.alloc(3 * N)); CKR(ctx->l2d_p0.alloc(3 * N)); CKR(ctx->l2d_p1.alloc(3 * N));
  if (N) {
    CK(cudaMemcpyAsync(ctx->l2d_n.p, n, 24 * N, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(ctx->l2d_p0.p, p0, 24 * N, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(ctx->l2d_p1.p, p1, 24 * N, cudaMemcpyDefault, ctx->stream));
    if (endpoints) {
      CKR(ctx->l2d_end.alloc(4 * N));
      CK(cudaMemcpyAsync(ctx->l2d_end.p, endpoints, 32 * N, cudaMemcpyDefault, ctx->stream));
    }
  }
  if (!endpoints) ctx->l2d_end.release();
  CK(cudaStreamSynchronize(ctx->stream));   // the caller's arrays are no longer read
  ctx->n_lines2d = n_lines2d;
  return take_launch_err(ctx);
}

extern "C" int lba_observations_from_tables(lba_ctx* ctx, int64_t n_rows, int32_t n_views, const int64_t* ids,
                                            const int64_t* bin, int64_t cap, int64_t* cam_idx, int64_t* line_idx,
                                            int64_t* line2d_id, int64_t* m_out, int64_t* n_tracks_out) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (n_rows < 0 || n_views < 1 || (long long)n_rows * n_views > 0x7fffffffLL) return fail(LBA_ERR_ARG, "bad table size");
  if (!m_out || !n_tracks_out) return fail(LBA_ERR_ARG, "NULL output");
  *m_out = 0; *n_tracks_out = 0;
  if (n_rows == 0) return LBA_OK;
  if (!ids || !bin || !cam_idx || !line_idx || !line2d_id) return fail(LBA_ERR_ARG, "NULL array");
  const long long TV = (long long)n_rows * n_views;
  if (cap < TV) return fail(LBA_ERR_ARG, "cap (%lld) must be at least rows x views (%lld)", (long long)cap, TV);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st_ = ctx->stream;
  DBuf<long long> s_ids, s_bin, o_cam, o_line, o_id, d_m; DBuf<int> d_nt, d_flag, d_pos; DBuf<unsigned char> d_tmp;
  const long long *d_ids = reinterpret_cast<const long long*>(ids), *d_bin = reinterpret_cast<const long long*>(bin);
  if (!is_device_ptr(ids)) { CKR(s_ids.alloc(TV)); CK(cudaMemcpyAsync(s_ids.p, ids, TV * 8, cudaMemcpyHostToDevice, st_)); d_ids = s_ids.p; }
  if (!is_device_ptr(bin)) { CKR(s_bin.alloc(TV)); CK(cudaMemcpyAsync(s_bin.p, bin, TV * 8, cudaMemcpyHostToDevice, st_)); d_bin = s_bin.p; }
  long long* outs[3] = {reinterpret_cast<long long*>(cam_idx), reinterpret_cast<long long*>(line_idx), reinterpret_cast<long long*>(line2d_id)};
  DBuf<long long>* stage[3] = {&o_cam, &o_line, &o_id};
  long long* dev_out[3];
  for (int k = 0; k < 3; ++k) {
    if (is_device_ptr(outs[k])) dev_out[k] = outs[k];
    else { CKR(stage[k]->alloc(TV)); dev_out[k] = stage[k]->p; }
  }
  CKR(d_nt.alloc(1)); CKR(d_m.alloc(1)); CKR(d_flag.alloc(TV)); CKR(d_pos.alloc(TV));
  CK(cudaMemsetAsync(d_nt.p, 0, 4, st_));
  LAUNCH(ctx, LBA_K_MISC, k_obs_count, cdiv(n_rows, 256), 256, (long long)n_rows, (int)n_views, d_bin, d_nt.p);
  LAUNCH(ctx, LBA_K_MISC, k_obs_flags, cdiv(TV, 256), 256, (long long)n_rows, (int)n_views, d_bin, (const int*)d_nt.p, d_flag.p);
  size_t tmp_bytes = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag.p, d_pos.p, (int)TV, st_));
  CKR(d_tmp.alloc(tmp_bytes));
  CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_flag.p, d_pos.p, (int)TV, st_));
  LAUNCH(ctx, LBA_K_MISC, k_obs_scatter, cdiv(TV, 256), 256, (long long)n_rows, (int)n_views, d_ids, (const int*)d_flag.p,
         (const int*)d_pos.p, dev_out[0], dev_out[1], dev_out[2], d_m.p);
  long long m = 0; int nt = 0;
  CK(cudaMemcpyAsync(&m, d_m.p, 8, cudaMemcpyDeviceToHost, st_));
  CK(cudaMemcpyAsync(&nt, d_nt.p, 4, cudaMemcpyDeviceToHost, st_));
  CK(cudaStreamSynchronize(st_));
  for (int k = 0; k < 3; ++k)
    if (dev_out[k] != outs[k] && m > 0) CK(cudaMemcpyAsync(outs[k], dev_out[k], (size_t)m * 8, cudaMemcpyDeviceToHost, st_));
  CK(cudaStreamSynchronize(st_));
  *m_out = m; *n_tracks_out = nt;
  return take_launch_err(ctx);
}

extern "C" int lba_set_problem_lines2d(lba_ctx* ctx, int64_t m, int64_t nc, int64_t nl, const int64_t* cam_idx,
                                       const int64_t* line_idx, const int64_t* line2d_id) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (m < 0 || m > 0x7fffffffLL / 12) return fail(LBA_ERR_ARG, "bad observation count");
  if (m > 0 && (!cam_idx || !line_idx || !line2d_id)) return fail(LBA_ERR_ARG, "NULL array");
  if (m > 0 && ctx->n_lines2d == 0) return fail(LBA_ERR_STATE, "lba_lines2d_set has not been called");
  CK(cudaSetDevice(ctx->device));
  const size_t M = (size_t)m;
  CKR(ctx->l2d_stage.alloc(9 * M + 2));
  double *gn = ctx->l2d_stage.p, *gp0 = gn + 3 * M, *gp1 = gp0 + 3 * M;
  if (m > 0) {
    DBuf<long long> s_id; DBuf<int> d_bad;
    const long long* d_id = reinterpret_cast<const long long*>(line2d_id);
    if (!is_device_ptr(line2d_id)) { CKR(s_id.alloc(M)); CK(cudaMemcpyAsync(s_id.p, line2d_id, M * 8, cudaMemcpyHostToDevice, ctx->stream)); d_id = s_id.p; }
    CKR(d_bad.alloc(1));
    CK(cudaMemsetAsync(d_bad.p, 0, 4, ctx->stream));
    LAUNCH(ctx, LBA_K_MISC, k_gather_lines2d, cdiv(3 * (long long)m, 256), 256, (long long)m, ctx->n_lines2d, d_id,
           (const double*)ctx->l2d_n.p, (const double*)ctx->l2d_p0.p, (const double*)ctx->l2d_p1.p, gn, gp0, gp1, d_bad.p);
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(LBA_ERR_ARG, "%d observation(s) name a 2D line outside [0, %lld)", bad, ctx->n_lines2d);
  }
  return lba_set_problem(ctx, m, nc, nl, cam_idx, line_idx, gn, gp0, gp1);
}

extern "C" int lba_triangulate_tracks(lba_ctx* ctx, int32_t n_views, const double* P, const int64_t* ids,
                                      const int64_t* n_new_tracks, float* X, double* points_3d, int64_t* n_tracks_out) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (n_views < 2 || !P || !n_new_tracks || !n_tracks_out) return fail(LBA_ERR_ARG, "bad argument");
  if (is_device_ptr(n_new_tracks)) return fail(LBA_ERR_ARG, "n_new_tracks is a host array");
  if (!ctx->l2d_end.p || ctx->n_lines2d == 0) return fail(LBA_ERR_STATE, "lba_lines2d_set was not given the endpoints");
  long long n = 0;
  for (int v = 0; v < n_views - 1; ++v) { if (n_new_tracks[v] < 0) return fail(LBA_ERR_ARG, "negative track count"); n += n_new_tracks[v]; }
  *n_tracks_out = n;
  if (n == 0) return LBA_OK;
  if (!ids || (!X && !points_3d)) return fail(LBA_ERR_ARG, "NULL array");
  if (!is_device_ptr(ids)) return fail(LBA_ERR_ARG, "ids must be the device table written by lba_track_table");
  CK(cudaSetDevice(ctx->device));
  DBuf<long long> d_v0; DBuf<float> d_p0, d_p1, d_X; DBuf<int> d_bad;
  CKR(d_v0.alloc(n)); CKR(d_p0.alloc(4 * n)); CKR(d_p1.alloc(4 * n)); CKR(d_bad.alloc(1));
  CK(cudaMemsetAsync(d_bad.p, 0, 4, ctx->stream));
  long long row = 0;
  for (int v = 0; v < n_views - 1; ++v) {
    const long long k = n_new_tracks[v];
    if (k > 0)
      LAUNCH(ctx, LBA_K_MISC, k_tri_gather, cdiv(k, 256), 256, k, row, row, (int)n_views, v, ctx->n_lines2d,
             reinterpret_cast<const long long*>(ids), (const double*)ctx->l2d_end.p, d_v0.p, reinterpret_cast<float2*>(d_p0.p),
             reinterpret_cast<float2*>(d_p1.p), d_bad.p);
    row += k;
  }
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bad) return fail(LBA_ERR_ARG, "%d track(s) name a 2D line outside [0, %lld)", bad, ctx->n_lines2d);
  float* Xd = X;
  if (!X || !is_device_ptr(X)) { CKR(d_X.alloc(6 * n)); Xd = d_X.p; }
  CKR(lba_triangulate_lines(ctx, n, n_views, P, reinterpret_cast<const int64_t*>(d_v0.p), nullptr, d_p0.p, d_p1.p, Xd, nullptr, nullptr));
  if (X && Xd != X) CK(cudaMemcpyAsync(X, Xd, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (points_3d) {
    // the BA starts from these float32 coordinates widened to double (bundle_adjustment.py:231-233, :311)
    if (is_device_ptr(points_3d)) {
      LAUNCH(ctx, LBA_K_MISC, k_f32_to_f64, cdiv(6 * n, 256), 256, 6 * n, (const float*)Xd, points_3d);
    } else {
      DBuf<double> d64;
      CKR(d64.alloc(6 * n));
      LAUNCH(ctx, LBA_K_MISC, k_f32_to_f64, cdiv(6 * n, 256), 256, 6 * n, (const float*)Xd, d64.p);
      CK(cudaMemcpyAsync(points_3d, d64.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

extern "C" int lba_residuals(lba_ctx* ctx, const double* cams, const double* pts, double* residuals, double* cost) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  CKR(load_state(ctx, cams, pts));
  CKR(run_residual(ctx, nullptr, 0));
  CKR(reduce_scalars(ctx, 1));
  CKR(export_residuals(ctx, ctx->r.p, residuals));
  CKR(fetch_scalars(ctx));
  if (cost) *cost = ctx->h_scal->cost;
  return LBA_OK;
}

extern "C" int lba_residuals_jacobian(lba_ctx* ctx, const double* cams, const double* pts, double* residuals,
                                      double* jcam, double* jline) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  CKR(load_state(ctx, cams, pts));
  ctx->have_scaling = false;
  CKR(run_jacobian(ctx, nullptr, 0));
  CKR(export_residuals(ctx, ctx->r.p, residuals));
  const int m = ctx->m;
  if (m > 0 && (jcam || jline)) {
    if (!jcam || !jline) return fail(LBA_ERR_ARG, "jcam and jline must both be given or both be NULL");
    DBuf<double> tmp;
    CKR(tmp.alloc((size_t)12 * m));
    LAUNCH(ctx, LBA_K_MISC, k_export_jacobian, cdiv(ctx->Mp, 256), 256, ctx->Mp, ctx->slot_src.p,
           ctx->tidx.p, ctx->jl.p, ctx->jc.p, tmp.p, tmp.p + (size_t)6 * m);
    CK(cudaMemcpyAsync(jcam, tmp.p, (size_t)6 * m * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(jline, tmp.p + (size_t)6 * m, (size_t)6 * m * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

extern "C" int lba_normal_equations(lba_ctx* ctx, double* U, double* gc, double* V, double* gl) {
  if (!ctx || !ctx->have_eval) return fail(LBA_ERR_STATE, "no resident evaluation: call lba_residuals_jacobian first");
  CK(cudaSetDevice(ctx->device));
  CKR(run_blocks(ctx, true));
  if (U && ctx->nc) CK(cudaMemcpyAsync(U, ctx->U.p, (size_t)ctx->nc * 21 * 8, cudaMemcpyDefault, ctx->stream));
  if (gc && ctx->nc) CK(cudaMemcpyAsync(gc, ctx->gc.p, (size_t)ctx->nc * 6 * 8, cudaMemcpyDefault, ctx->stream));
  if (V && ctx->nl) CK(cudaMemcpy2DAsync(V, 21 * 8, ctx->V.p, kVStride * 8, 21 * 8, (size_t)ctx->nl, cudaMemcpyDefault, ctx->stream));   // rows are padded to 22 on the device
  if (gl && ctx->nl) CK(cudaMemcpyAsync(gl, ctx->gl.p, (size_t)ctx->nl * 6 * 8, cudaMemcpyDefault, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

extern "C" int lba_schur_solve(lba_ctx* ctx, double lambda, double pcg_tol, int32_t pcg_max_iter, double* dcam,
                               double* dpts, int32_t* pcg_iters) {
  if (!ctx || !ctx->have_eval) return fail(LBA_ERR_STATE, "no resident evaluation: call lba_residuals_jacobian first");
  if (!(lambda > 0.0)) return fail(LBA_ERR_ARG, "lambda must be positive");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->have_blocks) CKR(run_blocks(ctx, false));
  int hint = 8, iters = 0, brk = 0;
  CKR(run_solve(ctx, lambda, pcg_tol, pcg_max_iter, &hint, &iters, &brk));
  if (pcg_iters) *pcg_iters = iters;
  if (dcam && ctx->nc) CK(cudaMemcpyAsync(dcam, ctx->cgx.p, (size_t)ctx->nc * 48, cudaMemcpyDefault, ctx->stream));
  if (dpts && ctx->nl) {
    // dl = pts_new - pts (pts_new = pts + dl was written by the back-substitution pass); y is free after the solve
    LAUNCH(ctx, LBA_K_MISC, k_sub, cdiv((long long)ctx->nl * 6, 256), 256, (size_t)ctx->nl * 6, ctx->pts_new.p, ctx->pts.p, ctx->y.p);
    CK(cudaMemcpyAsync(dpts, ctx->y.p, (size_t)ctx->nl * 48, cudaMemcpyDefault, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

// ---- C ABI: the LM loop ---------------------------------------------------------------------------
// Every decision of the loop (gain ratio, accept / reject, damping, step length, retry, termination) is taken on the
// device by k_lm_decide; the kernels of a pass read the damping, the step length and which evaluation set is current
// from the control block.  The host only enqueues passes: one host look per pass in this driver (it needs to know
// whether the PCG steps it enqueued were enough).
static int enqueue_trial_tail(lba_ctx* c, const LmCtl* ctl) {
  // trial evaluation WITH its Jacobian into the other evaluation set (it becomes the current one when accepted)
  CKR(run_jacobian(c, ctl, 1));
  // cost, rjd, jd2, dl2, x2_lines are contiguous; the gradient norm of the state the solve started from rides along
  CKR(allreduce_scalars(c, 5, ctl));
  // the decision kernel also posts the pass's scalars to the host mirror (what k_publish does for the other calls)
  const unsigned long long seq = ++c->publish_seq;
  LAUNCH(c, LBA_K_MISC, k_lm_decide, 1, 32, c->ctl.p, (const Scalars*)c->scal.p, (const CgState*)c->cgst.p,
         dual<const unsigned long long>(c->gbits.p, c->gbits_new.p), c->trace.p, (const int*)c->clamped.p, c->d_mirror, seq);
  return LBA_OK;
}

extern "C" int lba_lm_solve(lba_ctx* ctx, double* cams_api, double* pts_api, const lba_options* opts_in,
                            lba_report* rep, double* residuals) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  lba_options o;
  if (opts_in) o = *opts_in; else lba_default_options(&o);
  if (o.max_iter < 0 || o.pcg_max_iter < 1 || !(o.lambda0 > 0.0) || o.loss < 0 || o.loss > LBA_LOSS_ARCTAN || !(o.f_scale > 0.0) ||
      o.max_evals < 0 || o.max_backtrack < 0)
    return fail(LBA_ERR_ARG, "bad options");
  const auto t0 = std::chrono::steady_clock::now();
  lba_report local;
  if (!rep) rep = &local;
  memset(rep, 0, sizeof(*rep));
  lba_ctx* c = ctx;
  CKR(load_state(c, cams_api, pts_api));
  c->have_scaling = false;
  c->loss = o.loss; c->f_scale = o.f_scale;
  struct LossGuard { lba_ctx* c; ~LossGuard() { c->loss = 0; c->f_scale = 1.0; } } loss_guard{c};
  CK(cudaMemsetAsync(c->clamped.p, 0, 4, c->stream));
  // the camera scaling is a running maximum that every solve set-up refreshes (k_cam_finish): start it from zero
  if (c->nc > 0) CK(cudaMemsetAsync(c->Dc.p, 0, (size_t)c->nc * 48, c->stream));
  CKR(run_jacobian(c, nullptr, 0));
  CKR(reduce_scalars(c, 1));
  LmCtl init;
  memset(&init, 0, sizeof init);
  init.ftol = o.ftol; init.xtol = o.xtol; init.gtol = o.gtol; init.lambda_min = o.lambda_min; init.lambda_max = o.lambda_max;
  init.min_ratio = o.min_ratio; init.max_iter = o.max_iter; init.max_evals = o.max_evals; init.pcg_max_iter = o.pcg_max_iter;
  init.max_backtrack = o.max_backtrack; init.lam = o.lambda0;
  LAUNCH(c, LBA_K_MISC, k_lm_init, 1, 32, c->ctl.p, init, (const Scalars*)c->scal.p);
  const LmCtl* ctl = c->ctl.p;
  CKR(fetch_scalars(c));
  rep->cost_init = c->h_ctl->cost;
  if (c->h_ctl->status == -5) return fail(LBA_ERR_NONFINITE, "Residuals are not finite in the initial point.");
  int chunk_hint = 8;
  // The host looks at the device once per pass (it has to know whether the PCG steps it enqueued were enough and how
  // many the next solve will probably need), but the GPU does not wait for that look: the set-up kernels of the NEXT
  // pass -- elimination, Schur weights, preconditioner, first CG state, ~25 % of an iteration -- are enqueued BEHIND
  // the decision kernel before the host reads its result.  They take everything from the control block, and skip
  // themselves when the decision turns out to be "done", "retry the same direction" or "more PCG steps first".
  const bool ahead = !c->profiling && !getenv("LBA_NO_ENQUEUE_AHEAD");
  bool setup_enqueued = false;
  SolvePlan sp;
  while (!c->h_ctl->done) {
    const bool solve = c->h_ctl->need_solve != 0;
    if (solve && !setup_enqueued) CKR(solve_setup(c, ctl, 0.0, o.pcg_tol, o.pcg_max_iter, &sp, true));
    setup_enqueued = false;
    sp.issued = 0;
    int launch = std::max(1, std::min(chunk_hint, o.pcg_max_iter));
    while (true) {
      if (solve) {
        CKR(solve_iterate(c, &sp, launch));
        CKR(solve_finish(c, &sp));
      } else {
        CKR(launch_cam_step(c, ctl));   // same direction, shorter step: both halves of the trial point
      }
      CKR(enqueue_trial_tail(c, ctl));   // ends with the decision kernel, which publishes
      if (ahead) {
        const int issued = sp.issued;
        CKR(solve_setup(c, ctl, 0.0, o.pcg_tol, o.pcg_max_iter, &sp, true));
        sp.issued = issued;
      }
      CKR(wait_scalars(c));
      if (!c->h_ctl->cg_short) { setup_enqueued = ahead; break; }
      launch = 4;   // the prediction was short: more PCG steps, then redo the (cheap to discard) tail
    }
    chunk_hint = c->h_ctl->q_last + 1;
  }
  const LmCtl fin = *c->h_ctl;
  if (fin.cur) swap_sets(c);
  c->have_eval = true; c->have_blocks = false; c->have_scaling = true; c->pending_first = false; c->line_scaling_mode = 1;
  rep->iterations = fin.it;
  rep->status = fin.status;
  rep->cost = fin.cost;
  rep->n_jac = fin.evals;
  rep->n_res = 0;
  rep->n_accepted = fin.n_acc;
  rep->pcg_iters_total = fin.pcg_total;
  rep->chol_clamped = *c->h_clamped;
  CK(cudaMemcpyAsync(c->h_trace, c->trace.p, sizeof(LmTrace), cudaMemcpyDeviceToHost, c->stream));
  CKR(store_state(c, cams_api, pts_api));
  if (c->loss != 0 && residuals) {
    // with a robust loss the resident residuals are the rescaled ones; the caller gets fun itself (res.fun)
    c->loss = 0;
    CKR(run_residual(c, nullptr, 0));
    c->have_eval = false;
  }
  CKR(export_residuals(c, c->r.p, residuals));
  CK(cudaStreamSynchronize(c->stream));
  rep->trace_len = fin.trace_len;
  for (int k = 0; k < fin.trace_len; ++k) {
    const LmTrace& t = *c->h_trace;
    rep->trace_it[k] = t.it[k]; rep->trace_pcg[k] = t.pcg[k]; rep->trace_accepted[k] = t.accepted[k];
    rep->trace_cost[k] = t.cost[k]; rep->trace_cost_new[k] = t.cost_new[k]; rep->trace_lambda[k] = t.lam[k];
    rep->trace_pred[k] = t.pred[k]; rep->trace_step_norm[k] = t.step_norm[k]; rep->trace_g_norm[k] = t.g_norm[k];
    rep->trace_alpha[k] = t.alpha[k];
  }
  if (o.verbose >= 2) {
    // scipy-style table (least_squares verbose=2, the reference prints it at bundle_adjustment.py:319), one row per trial
    printf("   Iteration     Total nfev        Cost      Cost reduction    Step norm     Optimality   \n");
    for (int k = 0; k < rep->trace_len; ++k) {
      const bool acc = rep->trace_accepted[k] != 0;
      printf("   %6d        %6d        %.4e      %9.2e       %9.2e      %9.2e    (lam %.1e alpha %.3f pcg %d %s)\n",
             rep->trace_it[k], k + 2, acc ? rep->trace_cost_new[k] : rep->trace_cost[k],
             acc ? rep->trace_cost[k] - rep->trace_cost_new[k] : 0.0, rep->trace_step_norm[k], rep->trace_g_norm[k],
             rep->trace_lambda[k], rep->trace_alpha[k], rep->trace_pcg[k], acc ? "ok" : "rej");
    }
  }
  rep->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return take_launch_err(c);
}

// ---- C ABI: the 12-parameter [R|c] camera rows (see lba_dense12.cuh) ------------------------------
// cuSOLVER through dlopen, on first use (nothing else in this library needs it)
struct CusolverApi {
  void* handle = nullptr;
  int (*Create)(void**) = nullptr;
  int (*Destroy)(void*) = nullptr;
  int (*SetStream)(void*, cudaStream_t) = nullptr;
  int (*Dpotrf_bufferSize)(void*, int, int, double*, int, int*) = nullptr;
  int (*Dpotrf)(void*, int, int, double*, int, double*, int, int*) = nullptr;
  int (*Dpotrs)(void*, int, int, int, const double*, int, double*, int, int*) = nullptr;
};
static CusolverApi g_cusolver;
static int load_cusolver() {
  if (g_cusolver.handle) return LBA_OK;
  void* h = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcusolver.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("/usr/local/cuda/lib64/libcusolver.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(LBA_ERR_UNSUPPORTED, "the 12-parameter camera mode needs libcusolver (dense Cholesky): %s", dlerror());
  g_cusolver.Create = (decltype(g_cusolver.Create))dlsym(h, "cusolverDnCreate");
  g_cusolver.Destroy = (decltype(g_cusolver.Destroy))dlsym(h, "cusolverDnDestroy");
  g_cusolver.SetStream = (decltype(g_cusolver.SetStream))dlsym(h, "cusolverDnSetStream");
  g_cusolver.Dpotrf_bufferSize = (decltype(g_cusolver.Dpotrf_bufferSize))dlsym(h, "cusolverDnDpotrf_bufferSize");
  g_cusolver.Dpotrf = (decltype(g_cusolver.Dpotrf))dlsym(h, "cusolverDnDpotrf");
  g_cusolver.Dpotrs = (decltype(g_cusolver.Dpotrs))dlsym(h, "cusolverDnDpotrs");
  if (!g_cusolver.Create || !g_cusolver.Destroy || !g_cusolver.SetStream || !g_cusolver.Dpotrf_bufferSize || !g_cusolver.Dpotrf || !g_cusolver.Dpotrs)
    return fail(LBA_ERR_UNSUPPORTED, "libcusolver is missing required symbols");
  g_cusolver.handle = h;
  return LBA_OK;
}

// a (camera, line) pair observed twice would need a sum in the coupling block
__global__ void k_dup12(int Mp, const TileIdx* __restrict__ tidx, const int* __restrict__ line_slot, const int* __restrict__ line_cnt,
                        int* __restrict__ dup) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const TileIdx& ti = tidx[i / kTile];
  const int cam = ti.cam[i % kTile];
  if (cam < 0) return;
  const int l = ti.meta.l0 + ti.ll[i % kTile];
  for (int s = line_slot[l]; s < i; ++s)
    if (tidx[s / kTile].cam[s % kTile] == cam) { atomicAdd(dup, 1); return; }
  (void)line_cnt;
}

struct Dense12 {
  lba_ctx* c;
  int n, Mp;
  DBuf<double> x, x_new, r, r_new, jc, jc_new, jl, jl_new, H, A, g, D, rhs, sums, partial;
  DBuf<int> info, dup;
  DBuf<double> work;
  void* solver = nullptr;
  ~Dense12() { if (solver) g_cusolver.Destroy(solver); }
};

static int eval12(lba_ctx* c, Dense12& d, const double* x, double* r, double* jc, double* jl, bool with_jac, int loss, double f_scale) {
  Eval12Args a;
  a.Mp = c->Mp; a.ntiles = c->ntiles; a.nc = c->nc; a.nl = c->nl; a.tidx = c->tidx.p; a.cst = c->cst.p; a.x = x;
  a.r = r; a.jc = jc; a.jl = jl; a.partial = d.partial.p; a.ticket = c->tickets.p + 6; a.cost_out = d.sums.p;
  a.loss = loss; a.f_scale = f_scale;
  if (c->ntiles == 0) { CK(cudaMemsetAsync(d.sums.p, 0, 8, c->stream)); return LBA_OK; }
  if (with_jac) LAUNCH(c, LBA_K_JACOBIAN, k_eval12<true>, c->ntiles, kTile, a);
  else LAUNCH(c, LBA_K_RESIDUAL, k_eval12<false>, c->ntiles, kTile, a);
  return LBA_OK;
}

static int check12(lba_ctx* ctx, const double* cams, const double* pts) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (!ctx->have_problem) return fail(LBA_ERR_STATE, "lba_set_problem has not been called");
  if (!cams || !pts) return fail(LBA_ERR_ARG, "NULL parameter array");
  if (ctx->world > 1) return fail(LBA_ERR_UNSUPPORTED, "the 12-parameter camera mode runs on one GPU");
  return LBA_OK;
}

extern "C" int lba_residuals12(lba_ctx* ctx, const double* cams12, const double* pts, double* residuals, double* cost) {
  CKR(check12(ctx, cams12, pts));
  CK(cudaSetDevice(ctx->device));
  lba_ctx* c = ctx;
  Dense12 d; d.c = c; d.n = 12 * c->nc + 6 * c->nl; d.Mp = c->Mp;
  CKR(d.x.alloc(d.n)); CKR(d.r.alloc(c->Mp)); CKR(d.sums.alloc(8)); CKR(d.partial.alloc(2 * (size_t)std::max(c->ntiles, 1)));
  CK(cudaMemcpyAsync(d.x.p, cams12, (size_t)c->nc * 96, cudaMemcpyDefault, c->stream));
  CK(cudaMemcpyAsync(d.x.p + 12 * c->nc, pts, (size_t)c->nl * 48, cudaMemcpyDefault, c->stream));
  CKR(eval12(c, d, d.x.p, d.r.p, nullptr, nullptr, false, 0, 1.0));
  CKR(export_residuals(c, d.r.p, residuals));
  double h = 0.0;
  CK(cudaMemcpyAsync(&h, d.sums.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (cost) *cost = h;
  return take_launch_err(c);
}

extern "C" int lba_project12(lba_ctx* ctx, int64_t n_rows, const double* points, const double* cams12, double* proj) {
  if (!ctx) return fail(LBA_ERR_ARG, "ctx is NULL");
  if (n_rows < 0 || (n_rows > 0 && (!points || !cams12 || !proj))) return fail(LBA_ERR_ARG, "bad argument");
  if (n_rows == 0) return LBA_OK;
  CK(cudaSetDevice(ctx->device));
  DBuf<double> dp, dc, dout;
  const size_t N = (size_t)n_rows;
  CKR(dp.alloc(3 * N)); CKR(dc.alloc(12 * N)); CKR(dout.alloc(3 * N));
  CK(cudaMemcpyAsync(dp.p, points, 24 * N, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(dc.p, cams12, 96 * N, cudaMemcpyDefault, ctx->stream));
  LAUNCH(ctx, LBA_K_MISC, k_project12, cdiv(n_rows, 256), 256, (long long)n_rows, (const double*)dp.p, (const double*)dc.p, dout.p);
  CK(cudaMemcpyAsync(proj, dout.p, 24 * N, cudaMemcpyDefault, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return take_launch_err(ctx);
}

extern "C" int lba_lm_solve12(lba_ctx* ctx, double* cams12, double* pts_api, const lba_options* opts_in, lba_report* rep,
                              double* residuals) {
  CKR(check12(ctx, cams12, pts_api));
  lba_options o;
  if (opts_in) o = *opts_in; else lba_default_options(&o);
  if (o.max_iter < 0 || !(o.lambda0 > 0.0) || o.loss < 0 || o.loss > LBA_LOSS_ARCTAN || !(o.f_scale > 0.0) || o.max_evals < 0 || o.max_backtrack < 0)
    return fail(LBA_ERR_ARG, "bad options");
  const auto t0 = std::chrono::steady_clock::now();
  lba_report local;
  if (!rep) rep = &local;
  memset(rep, 0, sizeof(*rep));
  lba_ctx* c = ctx;
  CK(cudaSetDevice(c->device));
  const long long n64 = 12LL * c->nc + 6LL * c->nl;
  if (n64 > 16384) return fail(LBA_ERR_UNSUPPORTED, "the 12-parameter camera mode solves the dense normal equations: %lld parameters exceed its limit of 16384", n64);
  CKR(load_cusolver());
  Dense12 d; d.c = c; d.n = (int)n64; d.Mp = c->Mp;
  const int n = d.n; const size_t M = (size_t)c->Mp, N = (size_t)n;
  CKR(d.x.alloc(N)); CKR(d.x_new.alloc(N)); CKR(d.r.alloc(M)); CKR(d.r_new.alloc(M)); CKR(d.jc.alloc(12 * M)); CKR(d.jc_new.alloc(12 * M));
  CKR(d.jl.alloc(6 * M)); CKR(d.jl_new.alloc(6 * M)); CKR(d.H.alloc(N * N)); CKR(d.A.alloc(N * N)); CKR(d.g.alloc(N)); CKR(d.D.alloc(N));
  CKR(d.rhs.alloc(N)); CKR(d.sums.alloc(8)); CKR(d.partial.alloc(2 * (size_t)std::max(c->ntiles, 1))); CKR(d.info.alloc(1)); CKR(d.dup.alloc(1));
  if (g_cusolver.Create(&d.solver) != 0) return fail(LBA_ERR_CUDA, "cusolverDnCreate failed");
  g_cusolver.SetStream(d.solver, c->stream);
  int lwork = 0;
  if (n > 0 && g_cusolver.Dpotrf_bufferSize(d.solver, 0 /*CUBLAS_FILL_MODE_LOWER*/, n, d.A.p, n, &lwork) != 0) return fail(LBA_ERR_CUDA, "cusolverDnDpotrf_bufferSize failed");
  CKR(d.work.alloc((size_t)std::max(lwork, 1)));
  const int* d_line_cnt = c->line_tmp.p; const int* d_line_slot = d_line_cnt + 2 * ((size_t)c->nl + 1);
  if (c->Mp > 0) {
    CK(cudaMemsetAsync(d.dup.p, 0, 4, c->stream));
    k_dup12<<<cdiv(c->Mp, 256), 256, 0, c->stream>>>(c->Mp, c->tidx.p, d_line_slot, d_line_cnt, d.dup.p);
    int dup = 0;
    CK(cudaMemcpyAsync(&dup, d.dup.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (dup) return fail(LBA_ERR_UNSUPPORTED, "%d observation(s) repeat a (camera, line) pair: not supported with 12-parameter cameras", dup);
  }
  CK(cudaMemcpyAsync(d.x.p, cams12, (size_t)c->nc * 96, cudaMemcpyDefault, c->stream));
  CK(cudaMemcpyAsync(d.x.p + 12 * c->nc, pts_api, (size_t)c->nl * 48, cudaMemcpyDefault, c->stream));

  auto host_scalar = [&](const double* dev, int k, double* out) -> int {
    CK(cudaMemcpyAsync(out, dev, 8 * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return LBA_OK;
  };
  // normal equations of the current evaluation (r, jc, jl): H, g; scaling D updated by the next damp
  auto assemble = [&]() -> int {
    if (n == 0) return LBA_OK;
    CK(cudaMemsetAsync(d.H.p, 0, N * N * 8, c->stream));
    if (c->nc > 0) LAUNCH(c, LBA_K_CAM_BLOCKS, k_h12_cam, c->nc, 192, n, (const int*)c->cam_ptr.p, (const int*)c->cam_perm.p, (const double*)d.jc.p, (const double*)d.r.p, d.H.p, d.g.p);
    if (c->nl > 0) LAUNCH(c, LBA_K_CAM_BLOCKS, k_h12_line, cdiv((long long)c->nl * 42, 256), 256, n, c->nc, c->nl, d_line_slot, d_line_cnt, (const double*)d.jl.p, (const double*)d.r.p, d.H.p, d.g.p);
    if (c->Mp > 0) LAUNCH(c, LBA_K_CAM_BLOCKS, k_h12_cross, cdiv(c->Mp, 256), 256, c->Mp, n, c->nc, (const TileIdx*)c->tidx.p, (const double*)d.jc.p, (const double*)d.jl.p, d.H.p);
    return LBA_OK;
  };

  CKR(eval12(c, d, d.x.p, d.r.p, d.jc.p, d.jl.p, true, o.loss, o.f_scale));
  double cost0 = 0.0;
  CKR(host_scalar(d.sums.p, 1, &cost0));
  rep->cost_init = cost0;
  if (!std::isfinite(cost0)) return fail(LBA_ERR_NONFINITE, "Residuals are not finite in the initial point.");
  CKR(assemble());
  LmCtl L;
  memset(&L, 0, sizeof L);
  L.ftol = o.ftol; L.xtol = o.xtol; L.gtol = o.gtol; L.lambda_min = o.lambda_min; L.lambda_max = o.lambda_max; L.min_ratio = o.min_ratio;
  L.max_iter = o.max_iter; L.max_evals = o.max_evals; L.pcg_max_iter = 1; L.max_backtrack = o.max_backtrack;
  L.lam = o.lambda0; L.nu = 2.0; L.sigma = 1.0; L.alpha = 1.0; L.cost = cost0; L.need_solve = 1; L.evals = 1; L.scaling_mode = 2;
  if (L.max_iter <= 0) L.done = 1;
  LmTrace* tr = c->h_trace;
  bool first_scaling = true, new_eval = true;
  while (!L.done) {
    if (L.need_solve) {
      LAUNCH(c, LBA_K_LINE_FACTOR, k_damp12, cdiv((long long)n * n, 256), 256, n, (const double*)d.H.p, d.D.p, first_scaling ? 1 : 0, L.lam, d.A.p, new_eval ? 1 : 0);
      first_scaling = false; new_eval = false;
      LAUNCH(c, LBA_K_MISC, k_neg12, cdiv(n, 256), 256, n, (const double*)d.g.p, d.rhs.p);
      if (g_cusolver.Dpotrf(d.solver, 0, n, d.A.p, n, d.work.p, lwork, d.info.p) != 0) return fail(LBA_ERR_CUDA, "cusolverDnDpotrf failed");
      if (g_cusolver.Dpotrs(d.solver, 0, n, 1, d.A.p, n, d.rhs.p, n, d.info.p) != 0) return fail(LBA_ERR_CUDA, "cusolverDnDpotrs failed");
      if (c->ntiles > 0)
        LAUNCH(c, LBA_K_BACKSUB, k_jd12, c->ntiles, kTile, c->Mp, c->nc, (const TileIdx*)c->tidx.p, (const double*)d.jc.p, (const double*)d.jl.p,
               (const double*)d.r.p, (const double*)d.rhs.p, d.partial.p, c->tickets.p + 6, d.sums.p + 1);
      LAUNCH(c, LBA_K_MISC, k_norms12, 1, kTile, n, (const double*)d.rhs.p, (const double*)d.x.p, (const double*)d.g.p, d.sums.p + 3);
    }
    LAUNCH(c, LBA_K_MISC, k_step12, cdiv(n, 256), 256, n, (const double*)d.x.p, (const double*)d.rhs.p, L.alpha, d.x_new.p);
    CKR(eval12(c, d, d.x_new.p, d.r_new.p, d.jc_new.p, d.jl_new.p, true, o.loss, o.f_scale));
    LmTrial t;
    CKR(host_scalar(d.sums.p, 1, &t.cost_new));
    double tail[5];   // rjd, jd2, |delta|^2, |x|^2, |g|_inf of the solve this trial belongs to
    CKR(host_scalar(d.sums.p + 1, 5, tail));
    int info = 0;
    CK(cudaMemcpy(&info, d.info.p, 4, cudaMemcpyDeviceToHost));
    if (info != 0) return fail(LBA_ERR_CUDA, "dense Cholesky of the damped normal equations failed (info %d)", info);
    t.rjd = tail[0]; t.jd2 = tail[1]; t.d2 = tail[2]; t.x2 = tail[3]; t.g_norm = tail[4];
    t.pcg_iters = 0; t.pcg_done = 1;
    const int cur_before = L.cur;
    lm_decide(L, t, tr);
    if (L.cur != cur_before) {   // accepted: the trial evaluation becomes the current one
      swap_buf(d.x, d.x_new); swap_buf(d.r, d.r_new); swap_buf(d.jc, d.jc_new); swap_buf(d.jl, d.jl_new);
      CKR(assemble());
      new_eval = true;
    }
  }
  rep->iterations = L.it; rep->status = L.status; rep->cost = L.cost; rep->n_jac = L.evals; rep->n_accepted = L.n_acc;
  rep->trace_len = L.trace_len;
  for (int k = 0; k < L.trace_len; ++k) {
    rep->trace_it[k] = tr->it[k]; rep->trace_pcg[k] = 0; rep->trace_accepted[k] = tr->accepted[k];
    rep->trace_cost[k] = tr->cost[k]; rep->trace_cost_new[k] = tr->cost_new[k]; rep->trace_lambda[k] = tr->lam[k];
    rep->trace_pred[k] = tr->pred[k]; rep->trace_step_norm[k] = tr->step_norm[k]; rep->trace_g_norm[k] = tr->g_norm[k];
    rep->trace_alpha[k] = tr->alpha[k];
  }
  CK(cudaMemcpyAsync(cams12, d.x.p, (size_t)c->nc * 96, cudaMemcpyDefault, c->stream));
  CK(cudaMemcpyAsync(pts_api, d.x.p + 12 * c->nc, (size_t)c->nl * 48, cudaMemcpyDefault, c->stream));
  if (residuals) {
    if (o.loss != 0) CKR(eval12(c, d, d.x.p, d.r.p, nullptr, nullptr, false, 0, 1.0));   // fun itself, not the rescaled residuals
    CKR(export_residuals(c, d.r.p, residuals));
  }
  CK(cudaStreamSynchronize(c->stream));
  c->have_eval = false;
  rep->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return take_launch_err(c);
}

// ---- microbenchmarks --------------------------------------------------------------------------------
extern "C" int lba_measure_fp64_peak(lba_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return fail(LBA_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(ctx->device));
  const int blocks = ctx->n_sm * 8, threads = 256, iters = 20000;
  DBuf<double> out;
  CKR(out.alloc((size_t)blocks * threads));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    ++ctx->launches;
    k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(out.p, iters);
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  *tflops = best;
  return LBA_OK;
}

extern "C" int lba_measure_copy_bandwidth(lba_ctx* ctx, int64_t bytes, double* gbs) {
  if (!ctx || !gbs || bytes < 1024) return fail(LBA_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  const size_t n2 = (size_t)bytes / 16;
  DBuf<double2> a, b;
  CKR(a.alloc(n2)); CKR(b.alloc(n2));
  CK(cudaMemsetAsync(a.p, 0, n2 * 16, ctx->stream));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    ++ctx->launches;
    k_stream_copy<<<ctx->n_sm * 16, 256, 0, ctx->stream>>>(n2, a.p, b.p);
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    best = std::max(best, 2.0 * n2 * 16 / (ms * 1e-3) / 1e9);
  }
  *gbs = best;
  return LBA_OK;
}
