// Hand-written sm_100a kernels for the line bundle-adjustment hot path.
//
// Data layout in HBM (all fp64 / int32, structure-of-arrays over observations)
// -------------------------------------------------------------------------
// LINE ORDER, tile-aligned "slots".  Observations are sorted by 3D line (the reference's own
//   order, seq_utils.py:188-195) and packed into tiles of kTile = 256 slots; a tile holds whole
//   lines only (<= kMaxLinesPerTile), the unused tail slots are padding (camera = -1).  The
//   per-observation data are packed TILE-MAJOR so that everything one kernel needs of a tile is
//   a handful of contiguous, 16 B-aligned blocks, each moved by ONE bulk-async copy (TMA,
//   cp.async.bulk):
//     cst[ntiles][256][9] = per slot n, m0 = n x p0, m1 = n x p1 (18 KB per tile, one 72 B row per slot: the
//     evaluation kernel reuses the row of a slot as its staging row once the constants are in registers);
//     jl[ntiles][6][256]
//     (12 KB per tile);  TileIdx[ntiles] = {line table, local line id, camera, camera-order
//     position} (2.5 KB per tile);  r, w, wk, wr: [Mp] (2 KB per tile)
// CAMERA ORDER (stable counting sort of the same observations by camera), dense m entries:
//     jc[m][6] (one 48 B record per observation, so the scatter from the line-ordered Jacobian
//     kernel is 3 x 16 B stores); cam_perm[m] (camera-order position -> slot); cam_ptr[Nc+1]
// per line:   pts[Nl][6] = [X0,X1]; V[Nl][21]; gl[Nl][6]; Vinv[Nl][22] (21 + pad: 16 B aligned
//             rows); y[Nl][6]; Dl[Nl][6]
// per camera: cams[Nc][6] = [w,c]; CamRec[Nc]; G[Nc][9]; U[Nc][21]; gc[Nc][6]; Dc[Nc][6];
//             Minv[Nc][21]; CG vectors [Nc][6]; tab[Nc][6]
//
// Each residual is a scalar, so a Jacobian row is (jc_i in R^6, jl_i in R^6) and every coupling
// block W_{c,l} = jc_i jl_i^T is rank one: only ONE scalar per observation (s_i, w_i) ever
// crosses between the line-ordered and the camera-ordered halves of a pass.
//
// Execution model
// ---------------
// Line-ordered kernels are persistent CTAs (grid = SMs x resident CTAs) that walk the tiles with
// a 3-stage bulk-async (TMA) -> shared-memory pipeline synchronised by mbarriers: one elected
// thread issues the copies for tile t+3 while the 8 warps work on tile t, so HBM latency is
// covered by bytes in flight rather than by occupancy.  Segmented (per-line) sums run on the
// staged tile in shared memory with warp shuffles; camera-side sums run one CTA per camera over
// the camera-ordered copy with warp-shuffle reductions.  There are no floating-point atomics:
// every sum has a fixed order, so results are bit-reproducible run to run.
#pragma once
#include "lba_geometry.cuh"
#include "lba_triangulate.cuh"
#include "lba_tracks.cuh"
#include "lba_policy.cuh"
#include <stddef.h>
#include <stdint.h>

namespace lba {

constexpr int kTile = 256;             // slots per tile == threads per line-order CTA
constexpr int kMaxLinesPerTile = 128;
// k_line_pass pipeline depth (3 CTAs/SM).  The Schur-weights pass has no sums and no staging rows -- it only streams --
// and spent half of its warp time waiting for tile data with two stages (ncu source view): it gets more.
#ifndef LBA_LP_STAGES_SCALARS
#define LBA_LP_STAGES_SCALARS 3
#endif
__host__ __device__ constexpr int lp_stages(int mode) { return mode == 2 /* kModeScalars */ ? LBA_LP_STAGES_SCALARS : 2; }
constexpr int kVinvStride = 22;        // 21 packed entries + 1 pad -> 176 B rows (16 B aligned)
constexpr int kVStride = 22;           // V_l rows likewise: a tile's rows are one aligned bulk store of the Jacobian kernel

struct __align__(16) TileMeta {
  int l0;      // first global line of the tile
  int nl;      // lines in the tile
  int nobs;    // real observations (slots [0,nobs) are used)
  int frag;    // 0: the tile holds whole lines; f + 1: the tile is fragment f of a line with more than kTile observations
               // (nl == 1, l0 = that line): its per-line sums are partial and go to the fragment buffers
  unsigned short lend[kMaxLinesPerTile];  // exclusive end slot of local line ll (start = lend[ll-1] or 0)
};
static_assert(sizeof(TileMeta) % 16 == 0, "TileMeta must be bulk-copyable");

// Everything integer a tile needs, packed so that ONE bulk copy brings it in (a prefix of it for the
// kernels that do not need the later members).
struct __align__(16) TileIdx {
  TileMeta meta;
  unsigned char ll[kTile];   // local line id of the slot
  int cam[kTile];            // camera of the slot, -1 for padding
  int pos[kTile];            // camera-order position of the slot
};
static_assert(sizeof(TileIdx) % 16 == 0 && offsetof(TileIdx, cam) % 16 == 0 && offsetof(TileIdx, pos) % 16 == 0,
              "TileIdx prefixes must be bulk-copyable");
// tile-major packing of the per-slot double arrays: element (k, slot) of an NK-row array
__host__ __device__ inline size_t tm_index(int nk, int k, size_t slot) {
  return ((slot / kTile) * nk + k) * kTile + (slot % kTile);
}

// The LM loop keeps TWO evaluation sets (parameters, camera records, residuals, Jacobian, line blocks); which one
// is current is decided on the device (LmCtl::cur, flipped by k_lm_decide when a step is accepted), so every
// kernel of the loop gets both pointers and picks one after reading the control block.
template <typename T>
struct Dual {
  T* p[2];
  __host__ __device__ T* sel(int i) const { return i ? p[1] : p[0]; }
};
template <typename T>
__host__ inline Dual<T> dual(T* a, T* b) { Dual<T> d; d.p[0] = a; d.p[1] = b; return d; }
// kernels of the solve part run only in passes that solve; kernels of the trial part until the loop is done
// (cg_short: the trial of the running pass could not be judged yet -- its PCG needs more steps -- so a set-up that
// the host enqueued ahead for the NEXT pass must not touch the solve in progress)
__device__ __forceinline__ bool ctl_skip_solve(const LmCtl* ctl) { return ctl && (ctl->done || !ctl->need_solve || ctl->cg_short); }
// the back-substitution belongs to the pass in progress: it runs again when the PCG was continued (cg_short)
__device__ __forceinline__ bool ctl_skip_backsub(const LmCtl* ctl) { return ctl && (ctl->done || !ctl->need_solve); }
__device__ __forceinline__ bool ctl_skip_trial(const LmCtl* ctl) { return ctl && ctl->done; }
__device__ __forceinline__ int ctl_cur(const LmCtl* ctl) { return ctl ? ctl->cur : 0; }

// ---- bulk-async (TMA) + mbarrier primitives ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Programmatic dependent launch: every kernel waits for its stream predecessor before touching global
// memory, then lets its successor be scheduled, so at most one not-yet-started grid sits resident behind
// the running one and the launch latency of kernel N+1 hides under kernel N.  A no-op for plain launches.
__constant__ int g_pdl_early_trigger = 0;
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (g_pdl_early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy, completion counted in bytes on `bar`; 16 B aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum, returned to every thread.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red /* >= THREADS/32 doubles */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = (l < THREADS / 32) ? red[l] : 0.0;
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// Grid-wide sum without a second launch: every CTA publishes its partial(s), the CTA that takes the
// last ticket adds them up in index order (fixed order => bit-reproducible) and resets the ticket.
template <int THREADS, int NCOLS>
__device__ __forceinline__ void last_cta_reduce(const double (&mine)[NCOLS], double* partial, unsigned int* ticket,
                                                double* out, double scale, double* red) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < NCOLS; ++c) partial[(size_t)blockIdx.x * NCOLS + c] = mine[c];
    __threadfence();
    const unsigned int tk = atomicAdd(ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int c = 0; c < NCOLS; ++c) {
    double acc = 0.0;
    for (int g = threadIdx.x; g < (int)gridDim.x; g += THREADS) acc += __ldcg(partial + (size_t)g * NCOLS + c);
    const double sum = block_sum<THREADS>(acc, red);
    if (threadIdx.x == 0) out[c] = sum * scale;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

__device__ __forceinline__ void load_camrec(const CamRec* __restrict__ recs, int cam, double w[3], double c[3],
                                            double& A, double& B, double& C) {
  const double* p = reinterpret_cast<const double*>(recs + cam);
  const double2 v0 = ldg2(p), v1 = ldg2(p + 2), v2 = ldg2(p + 4), v3 = ldg2(p + 6), v4 = ldg2(p + 8);
  w[0] = v0.x; w[1] = v0.y; w[2] = v1.x;
  c[0] = v1.y; c[1] = v2.x; c[2] = v2.y;
  A = v3.x; B = v3.y; C = v4.x;
}

__device__ __forceinline__ void load6(const double* __restrict__ p, double v[6]) {
  const double2 a = ldg2(p), b = ldg2(p + 2), c = ldg2(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}

// ---- NVLink peer-memory primitives (see the all-reduce section near the end of this file) ----------
constexpr int kMaxPeers = 8;
enum { kRgnQ = 0, kRgnU, kRgnM, kRgnS, kRgnR, kRgnCount };   // kRgnR: results of the two-phase reduce
constexpr int kPeerDataOff = 1024;   // flags: kRgnCount*2*kMaxPeers*8 = 640 B

struct PeerArgs {
  unsigned char* peer[kMaxPeers];   // peer[r] = base of rank r's symmetric buffer as mapped here
  int rank, world;
  unsigned long long* trace;        // optional (LBA_PEER_TRACE): per exchange type {calls, ns waiting for peers, ns total}
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Every device-side wait on another CTA or another GPU is bounded: after g_spin_timeout_ns (LBA_SPIN_TIMEOUT_MS,
// default 10 s) without progress the waiter records the failure in g_spin_error and goes on, so a stalled or dead
// rank cannot hang the surviving GPUs; the host reads the word at its next synchronisation and fails the call with
// LBA_ERR_COMM (the results of that call are garbage).
__device__ unsigned int g_spin_error = 0;
__constant__ unsigned long long g_spin_timeout_ns = 10000000000ULL;
struct SpinGuard {
  unsigned long long t0 = 0;
  unsigned int n = 0;
  __device__ __forceinline__ bool expired(unsigned int code) {
    if ((++n & 0x3ffu) != 0u) return false;
    const unsigned long long t = global_ns();
    if (!t0) { t0 = t; return false; }
    if (t - t0 <= g_spin_timeout_ns) return false;
    atomicOr(&g_spin_error, code);
    return true;
  }
};
enum { kSpinPeerLink = 1u, kSpinGridBar = 2u, kSpinPeerReduce = 4u, kSpinPeerReduceRs = 8u };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_relaxed_sys2(const double* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// A fused producer epilogue / consumer prologue pair for one region of the symmetric buffers:
// the LAST CTA of the producer grid (ticket) stores seq into flag[region][slot][my_rank] of every
// rank; the consumer waits until its own flags of all ranks reach seq.
struct PeerLink {
  PeerArgs pa;
  int region;
  size_t slot_off;          // doubles, inside the data area, of (region, seq & 1)
  unsigned long long seq;   // 0: peer path inactive (single GPU or NCCL transport)
  unsigned int* ticket;
};
// Every CTA orders its writes before its ticket at GPU scope; only the last CTA, which has observed all tickets, pays
// the system-scope fence before it raises the peers' flags (causality order is cumulative: what it has observed
// becomes visible to whoever acquires the flag).
__device__ __forceinline__ void peer_signal_last_cta(const PeerLink& pl) {
  __shared__ int s_last_peer;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last_peer = (atomicAdd(pl.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last_peer) return;
  if (threadIdx.x == 0) { *pl.ticket = 0u; __threadfence_system(); }
  __syncthreads();
  if (threadIdx.x < pl.pa.world) {
    const int fidx = (pl.region * 2 + (int)(pl.seq & 1ULL)) * kMaxPeers;
    st_release_sys(reinterpret_cast<unsigned long long*>(pl.pa.peer[threadIdx.x]) + fidx + pl.pa.rank, pl.seq);
  }
}
__device__ __forceinline__ void peer_wait(const PeerLink& pl) {
  if (threadIdx.x < pl.pa.world) {
    const int fidx = (pl.region * 2 + (int)(pl.seq & 1ULL)) * kMaxPeers;
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(pl.pa.peer[pl.pa.rank]) + fidx + threadIdx.x;
    SpinGuard sg;
    while (ld_acquire_sys(f) < pl.seq) { if (sg.expired(kSpinPeerLink)) break; }
  }
  __syncthreads();
}
// sum over ranks, in rank order, of element e of the region slot
__device__ __forceinline__ double peer_sum(const PeerLink& pl, size_t e) {
  double v[kMaxPeers];
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)   // all W remote loads are in flight before the first add
    v[r] = r < pl.pa.world ? ld_relaxed_sys(reinterpret_cast<const double*>(pl.pa.peer[r] + kPeerDataOff) + pl.slot_off + e) : 0.0;
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) acc += v[r];
  return acc;
}

// ---------------------------------------------------------------------------------
// setup kernels
// ---------------------------------------------------------------------------------

// AoS (m,3) n/p0/p1 in caller order -> cst[Mp][9] rows in slot order, with m_e = n x p_e
// (np.cross at bundle_adjustment.py:126,133, hoisted out of the evaluation loop).
__global__ void k_setup_constants(int Mp, const int* __restrict__ slot_src, const double* __restrict__ n,
                                  const double* __restrict__ p0, const double* __restrict__ p1,
                                  double* __restrict__ cst) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const int src = slot_src[i];
  double* o = cst + (size_t)i * 9;
  if (src < 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = 0.0;
    return;
  }
  const size_t s = (size_t)src;
  const double nx = n[3 * s], ny = n[3 * s + 1], nz = n[3 * s + 2];
  const double ax = p0[3 * s], ay = p0[3 * s + 1], az = p0[3 * s + 2];
  const double bx = p1[3 * s], by = p1[3 * s + 1], bz = p1[3 * s + 2];
  o[0] = nx; o[1] = ny; o[2] = nz;
  o[3] = ny * az - nz * ay; o[4] = nz * ax - nx * az; o[5] = nx * ay - ny * ax;
  o[6] = ny * bz - nz * by; o[7] = nz * bx - nx * bz; o[8] = nx * by - ny * bx;
}

// (Nc,6) interleaved [w0,c0,w1,c1,w2,c2] <-> internal [w,c]
__global__ void k_cams_from_api(int nc, const double* __restrict__ api, double* __restrict__ cams) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * 6) return;
  const int c = t / 6, k = t % 6;  // k = 2*axis + {0: w, 1: centre}
  cams[c * 6 + (k & 1) * 3 + (k >> 1)] = api[t];
}
__global__ void k_cams_to_api(int nc, const double* __restrict__ cams, double* __restrict__ api) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * 6) return;
  const int c = t / 6, k = t % 6;
  api[t] = cams[c * 6 + (k & 1) * 3 + (k >> 1)];
}

__global__ void k_cam_prepare(int nc, const double* __restrict__ cams, CamRec* __restrict__ recs,
                              double* __restrict__ Gm) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  CamRec rec;
  for (int k = 0; k < 3; ++k) { rec.w[k] = cams[c * 6 + k]; rec.c[k] = cams[c * 6 + 3 + k]; }
  rodrigues_coeffs(rec.w, rec.A, rec.B, rec.C);
  rec.pad = 0.0;
  recs[c] = rec;
  if (Gm) {
    double G[9];
    build_G(rec.w, rec.B, rec.C, G);
    for (int k = 0; k < 9; ++k) Gm[c * 9 + k] = G[k];
  }
}

// ---- problem set-up on the device: only per-LINE work (tiling) stays on the host -------------------
// range checks, sortedness and the per-line observation count of the caller's index arrays
__global__ void k_index_scan(int m, int nc, int nl, const long long* __restrict__ cam, const long long* __restrict__ line,
                             int* __restrict__ line_cnt, int* __restrict__ flags /* [0] bad cam, [1] bad line, [2] unsorted */,
                             int* __restrict__ keys /* optional: line as int, for the stable sort of unsorted input */) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long c = cam[i], l = line[i];
  if (c < 0 || c >= nc) { atomicMax(&flags[0], i + 1); return; }
  if (l < 0 || l >= nl) { atomicMax(&flags[1], i + 1); return; }
  if (i > 0 && l < line[i - 1]) flags[2] = 1;
  atomicAdd(&line_cnt[l], 1);
  if (keys) keys[i] = (int)l;
}
__global__ void k_iota(int n, int* __restrict__ v) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
// place every observation into its slot: line-sorted position i (order[i] = caller index, or identity)
__global__ void k_fill_slots(int m, const int* __restrict__ order, const long long* __restrict__ cam,
                             const long long* __restrict__ line, const int* __restrict__ line_ptr,
                             const int* __restrict__ line_slot /* first slot of each line */, int* __restrict__ slot_src,
                             TileIdx* __restrict__ tidx) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int src = order ? order[i] : i;
  const int l = (int)line[src];
  const int slot = line_slot[l] + (i - line_ptr[l]);
  slot_src[slot] = src;
  TileIdx& ti = tidx[slot / kTile];
  ti.cam[slot % kTile] = (int)cam[src];
  ti.ll[slot % kTile] = (unsigned char)(l - ti.meta.l0);
}
// TileIdx records start as {meta from the host, ll = 0, cam = -1, pos = 0}
__global__ void k_init_tidx(int ntiles, const TileMeta* __restrict__ meta, TileIdx* __restrict__ tidx) {
  pdl_prologue();
  const int g = blockIdx.x, t = threadIdx.x;
  if (g >= ntiles) return;
  TileIdx& ti = tidx[g];
  if (t == 0) ti.meta = meta[g];
  ti.ll[t] = 0; ti.cam[t] = -1; ti.pos[t] = 0;
}

// ---- camera order on the device: stable radix sort of (camera, slot), then these two kernels ----
__global__ void k_cam_keys(int Mp, int nc, const TileIdx* __restrict__ tidx, int* __restrict__ keys, int* __restrict__ vals) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const int c = tidx[i / kTile].cam[i % kTile];
  keys[i] = c < 0 ? nc : c;   // padding slots sort behind every camera
  vals[i] = i;
}
// sorted (key, slot) -> cam_pos[slot] = j, cam_ptr[c] = first j with key >= c
__global__ void k_cam_index(int Mp, int m, int nc, const int* __restrict__ keys, const int* __restrict__ slots,
                            TileIdx* __restrict__ tidx, int* __restrict__ cam_ptr) {
  pdl_prologue();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Mp) return;
  const int k = keys[j];
  if (j < m) { const int sl = slots[j]; tidx[sl / kTile].pos[sl % kTile] = j; }
  const int kprev = j > 0 ? keys[j - 1] : -1;
  for (int c = kprev + 1; c <= k && c <= nc; ++c) cam_ptr[c] = j;
  if (j == Mp - 1)
    for (int c = k + 1; c <= nc; ++c) cam_ptr[c] = m;
}

// point_wise_input (:235-243) integer expansion, in the caller's observation order.
__global__ void k_point_wise_indices(int Mp, const int* __restrict__ slot_src, const TileIdx* __restrict__ tidx,
                                     long long* __restrict__ cam_pt, long long* __restrict__ pt_idx) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const int src = slot_src[i];
  if (src < 0) return;
  const size_t o = (size_t)src;
  const TileIdx& ti = tidx[i / kTile];
  const long long c = ti.cam[i % kTile], l = ti.meta.l0 + ti.ll[i % kTile];
  cam_pt[2 * o] = c; cam_pt[2 * o + 1] = c;
  pt_idx[2 * o] = 2 * l; pt_idx[2 * o + 1] = 2 * l + 1;
}

// ---------------------------------------------------------------------------------
// K1/K2/K3 (line side): residuals, Jacobian rows, per-line blocks V_l, g_l
// ---------------------------------------------------------------------------------
struct EvalArgs {
  int Mp, ntiles, lmax;    // lmax: max lines per tile rounded up to 8 (sizes the smem stage)
  int nc, nstages;         // cameras (size of the shared-memory table), pipeline depth (<= kMaxStages)
  const TileIdx* tidx;     // [ntiles]
  const double* cst;       // [ntiles][256][9]
  // the two evaluation sets; the kernel works on set (ctl ? ctl->cur ^ which : which)
  Dual<const CamRec> recs;
  Dual<const double> pts;  // [nl][6]
  Dual<double> r;          // [Mp]
  Dual<double> jl;         // [ntiles][6][256]
  Dual<double> jc;         // [m][6] camera order, one 48 B record per observation
  Dual<double> V;          // [nl][21]
  Dual<double> gl;         // [nl][6]
  Dual<unsigned long long> gbits;  // inf-norm of the gradient, as ordered bits
  const LmCtl* ctl;        // LM loop: control block (NULL for the stand-alone API calls)
  int which;               // 0: the current set, 1: the trial set
  double* fragV;           // [n_fragments][27] partial V_l | g_l of the fragments of long lines (k_frag_blocks adds them up)
  double* cost_partial;    // [gridDim]
  unsigned int* ticket;    // last-CTA reduction ticket
  double* cost_out;        // 0.5 * sum r^2
  int loss;                // 0 linear, else scipy robust loss (cost 0.5 f_scale^2 sum rho((r/f_scale)^2))
  double f_scale;
};

__constant__ int c_sym_a[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__constant__ int c_sym_b[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};

constexpr int kMaxStages = 4;
// 256-thread sub-group barrier (the sub-groups of one CTA run independent tile pipelines)
template <int NSUB>
__device__ __forceinline__ void sub_sync(int sub) {
  if (NSUB == 1) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "r"(kTile) : "memory");
}

// shared-memory stage of k_eval: cst[9][256] | TileIdx | pts[lmax][6]   (three bulk copies per tile)
constexpr int kEvOffIdx = 9 * kTile * 8;
constexpr int kEvOffPts = kEvOffIdx + (int)sizeof(TileIdx);
__host__ __device__ inline int eval_stage_bytes(int lmax) { return (kEvOffPts + lmax * 48 + 127) & ~127; }
// The staging row of a slot for the per-line sums, (jl[6], r), overwrites the slot's own constants row in the stage
// (9 doubles = 72 B: an odd number of 8 B words, conflict-free per half-warp) once the thread holds the constants in
// registers -- no separate staging buffers, so a third pipeline stage fits beside the camera table.
constexpr int kEvRow = 9;
__host__ __device__ inline int eval_sub_bytes(int lmax, int nstages) { return nstages * eval_stage_bytes(lmax); }
__host__ __device__ inline int eval_tab_bytes(int nc) { return (nc * (int)sizeof(CamRec) + 127) & ~127; }
__host__ __device__ inline int eval_smem_bytes(int lmax, int nstages, int nsub, int nc_tab) {
  return eval_tab_bytes(nc_tab) + nsub * eval_sub_bytes(lmax, nstages) + 128;
}

// Fused residual (+ Jacobian + per-line blocks) over the tiles.  A CTA is NSUB independent
// 256-thread sub-groups; with TAB_SMEM they share one shared-memory copy of the camera table
// (the per-observation camera gather is the L1-wavefront hot spot when it goes to global memory).
template <bool WITH_JAC, int NSUB, bool TAB_SMEM>
__global__ void __launch_bounds__(kTile * NSUB, TAB_SMEM ? 1 : 2) k_eval(const EvalArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red[kTile * NSUB / 32];
  __shared__ uint64_t full[NSUB][kMaxStages];
  const int S = a.nstages;
  const int stage_bytes = eval_stage_bytes(a.lmax);
  const int sub = threadIdx.x / kTile, t = threadIdx.x % kTile;
  const int tab_bytes = TAB_SMEM ? eval_tab_bytes(a.nc) : 0;
  unsigned char* sbase = smem + tab_bytes + (size_t)sub * eval_sub_bytes(a.lmax, S);
  // With the Jacobian the only barrier of a tile is the one between staging the rows and the per-line sums; the
  // stage of tile it-1 (its staging rows included) is refilled right after the barrier of tile it -- every warp has
  // passed the sums of tile it-1 by then -- with tile it-1+S, which so has more than a whole tile time to land.
  const bool deferred = WITH_JAC;
  const int first = blockIdx.x * NSUB + sub, stride = gridDim.x * NSUB;
  const int my_n = first < a.ntiles ? (a.ntiles - first + stride - 1) / stride : 0;

  if (t == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[sub][s], 1);
    fence_barrier_init();
  }
  pdl_prologue();
  if (ctl_skip_trial(a.ctl)) return;   // uniform over the grid
  const int set = ctl_cur(a.ctl) ^ a.which;
  const CamRec* g_recs = a.recs.sel(set);
  const double* g_pts = a.pts.sel(set);
  double* g_r = a.r.sel(set);
  double* g_jl = a.jl.sel(set);
  double* g_jc = a.jc.sel(set);
  double* g_V = a.V.sel(set);
  double* g_gl = a.gl.sel(set);
  const CamRec* recs = TAB_SMEM ? reinterpret_cast<const CamRec*>(smem) : g_recs;
  if (TAB_SMEM) {
    const double2* src = reinterpret_cast<const double2*>(g_recs);
    double2* dst = reinterpret_cast<double2*>(smem);
    for (int q = threadIdx.x; q < a.nc * (int)(sizeof(CamRec) / 16); q += kTile * NSUB) dst[q] = __ldg(src + q);
  }
  __syncthreads();

  auto issue = [&](int it, int nl, int l0) {
    const int tile = first + it * stride;
    unsigned char* st = sbase + (size_t)(it % S) * stage_bytes;
    uint64_t* bar = &full[sub][it % S];
    const uint32_t idx_bytes = WITH_JAC ? (uint32_t)sizeof(TileIdx) : (uint32_t)offsetof(TileIdx, pos);
    mbar_expect_tx(bar, 9 * kTile * 8 + idx_bytes + nl * 48);
    bulk_g2s(st, a.cst + (size_t)tile * 9 * kTile, 9 * kTile * 8, bar);
    bulk_g2s(st + kEvOffIdx, a.tidx + tile, idx_bytes, bar);
    if (nl > 0) bulk_g2s(st + kEvOffPts, g_pts + (size_t)l0 * 6, nl * 48, bar);
  };
  // the issuing thread keeps (l0, nl) of the tile it will issue next in registers, one iteration ahead
  int nx_l0 = 0, nx_nl = 0;
  if (t == 0) {
    for (int it = 0; it < S && it < my_n; ++it) {
      const TileMeta* mt = &a.tidx[first + it * stride].meta;
      issue(it, mt->nl, mt->l0);
    }
    if (S < my_n) { const TileMeta* mt = &a.tidx[first + S * stride].meta; nx_l0 = mt->l0; nx_nl = mt->nl; }
  }

  double cost_acc = 0.0, gmax = 0.0;
  for (int it = 0; it < my_n; ++it) {
    const int tile = first + it * stride;
    unsigned char* st = sbase + (size_t)(it % S) * stage_bytes;
    double* s_cst = reinterpret_cast<double*>(st);
    const TileIdx* s_idx = reinterpret_cast<const TileIdx*>(st + kEvOffIdx);
    const int* s_cam = s_idx->cam;
    const int* s_pos = s_idx->pos;
    const unsigned char* s_ll = s_idx->ll;
    const TileMeta* s_meta = &s_idx->meta;
    const double* s_pts = reinterpret_cast<const double*>(st + kEvOffPts);
    double* s_stg = s_cst;   // staging rows alias the constants rows of the same slots
    int cur_l0 = 0, cur_nl = 0;
    {  // prefetch the scalars for the issue after next
      const int nxt = deferred ? it + S : it + S + 1;
      if (t == 0 && nxt < my_n && (!deferred || it >= 1)) {
        const TileMeta* mt = &a.tidx[first + nxt * stride].meta;
        cur_l0 = mt->l0; cur_nl = mt->nl;
      }
    }
    mbar_wait(&full[sub][it % S], (it / S) & 1);

    const int cam = s_cam[t];
    const bool active = cam >= 0;
    const size_t i = (size_t)tile * kTile + t;
    double r = 0.0, rho_cost = 0.0;
    double jl[6] = {0, 0, 0, 0, 0, 0};
    if (active) {
      double n[3], m0[3], m1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        n[k] = s_cst[t * 9 + k];
        m0[k] = s_cst[t * 9 + 3 + k];
        m1[k] = s_cst[t * 9 + 6 + k];
      }
      double w[3], c[3], A, B, C;
      if (TAB_SMEM) {
        const double2* p = reinterpret_cast<const double2*>(recs + cam);
        const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3], v4 = p[4];
        w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; c[0] = v1.y; c[1] = v2.x; c[2] = v2.y;
        A = v3.x; B = v3.y; C = v4.x;
      } else {
        load_camrec(recs, cam, w, c, A, B, C);
      }
      const double* X = s_pts + (int)s_ll[t] * 6;
      double X0[3] = {X[0], X[1], X[2]}, X1[3] = {X[3], X[4], X[5]};
      double R[9];
      build_R(w, A, B, R);
      double tq[3] = {0, 0, 0};
      const double e0 = endpoint_term<WITH_JAC>(R, c, X0, n, m0, jl, tq);
      const double e1 = endpoint_term<WITH_JAC>(R, c, X1, n, m1, jl + 3, tq);
      r = e0 + e1;
      if (a.loss) {
        // scipy's robust-loss handling (construct_loss_function / scale_for_robust_loss_function,
        // _lsq/least_squares.py, _lsq/common.py:720): cost 0.5 fs^2 rho(z); with the Jacobian the row is
        // scaled by sqrt(max(rho' + 2 rho'' z, eps)) and the residual becomes r rho' / that factor
        const double rs = r / a.f_scale;
        const double z = rs * rs;
        double r0, r1, r2;
        robust_rho(a.loss, z, r0, r1, r2);
        rho_cost = a.f_scale * a.f_scale * r0;
        if (WITH_JAC) {
          double js = r1 + 2.0 * r2 * z;
          js = sqrt(js < 2.220446049250313e-16 ? 2.220446049250313e-16 : js);
          r = r * r1 / js;
#pragma unroll
          for (int k = 0; k < 6; ++k) jl[k] *= js;
          tq[0] *= js; tq[1] *= js; tq[2] *= js;
        }
      } else {
        rho_cost = r * r;
      }
      if (WITH_JAC) {
        double G[9];
        build_G(w, B, C, G);
        // d r/d w = G^T tq ;  d r/d c = -(jl0 + jl1): one 48 B record at the camera-order position
        double2* dst = reinterpret_cast<double2*>(g_jc + (size_t)s_pos[t] * 6);
        dst[0] = make_double2(G[0] * tq[0] + G[3] * tq[1] + G[6] * tq[2], G[1] * tq[0] + G[4] * tq[1] + G[7] * tq[2]);
        dst[1] = make_double2(G[2] * tq[0] + G[5] * tq[1] + G[8] * tq[2], -(jl[0] + jl[3]));
        dst[2] = make_double2(-(jl[1] + jl[4]), -(jl[2] + jl[5]));
      }
    }
    g_r[i] = r;
    cost_acc += rho_cost;
    if (WITH_JAC) {
#pragma unroll
      for (int k = 0; k < 6; ++k) g_jl[((size_t)tile * 6 + k) * kTile + t] = jl[k];
      // stage (jl, r) row-major with an odd-word stride for the per-line segmented sums
#pragma unroll
      for (int k = 0; k < 6; ++k) s_stg[t * kEvRow + k] = jl[k];
      s_stg[t * kEvRow + 6] = r;
      if (deferred) fence_proxy_async();
      sub_sync<NSUB>(sub);
      if (deferred && t == 0 && it >= 1) {   // every warp is done with tile it-1: refill its stage
        if (it - 1 + S < my_n) issue(it - 1 + S, nx_nl, nx_l0);
        nx_l0 = cur_l0; nx_nl = cur_nl;
      }
      // per-line blocks V_l = sum jl jl^T (21) and g_l = sum jl r (6): four register-blocked work
      // items per line (rows {0}, {1}, {2,5}, {3,4} of the upper triangle: 7/6/7/7 FMAs per
      // observation on 7/6/5/4 shared-memory loads), fixed summation order
      const int l0 = s_meta->l0, nlg = s_meta->nl, frag = s_meta->frag;
      for (int wi = t; wi < nlg * 4; wi += kTile) {
        const int ll = wi >> 2, part = wi & 3;
        const int s0 = ll ? s_meta->lend[ll - 1] : 0, s1 = s_meta->lend[ll];
        double* Vo = frag ? a.fragV + (size_t)(frag - 1) * 27 : g_V + (size_t)(l0 + ll) * kVStride;
        double* go = frag ? Vo + 21 : g_gl + (size_t)(l0 + ll) * 6;
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        double g0 = 0.0, g1 = 0.0;
        if (part == 0) {
#pragma unroll 4
          for (int s = s0; s < s1; ++s) {
            const double* row = s_stg + s * kEvRow;
            const double x0 = row[0];
#pragma unroll
            for (int b = 0; b < 6; ++b) acc[b] += x0 * row[b];
            acc[6] += x0 * row[6];
          }
#pragma unroll
          for (int b = 0; b < 6; ++b) Vo[b] = acc[b];
          go[0] = acc[6]; g0 = acc[6];
        } else if (part == 1) {
#pragma unroll 4
          for (int s = s0; s < s1; ++s) {
            const double* row = s_stg + s * kEvRow;
            const double x1 = row[1];
#pragma unroll
            for (int b = 1; b < 6; ++b) acc[b] += x1 * row[b];
            acc[6] += x1 * row[6];
          }
#pragma unroll
          for (int b = 1; b < 6; ++b) Vo[5 + b] = acc[b];   // sym_idx(1,b) = 6 + (b-1)
          go[1] = acc[6]; g0 = acc[6];
        } else if (part == 2) {
#pragma unroll 4
          for (int s = s0; s < s1; ++s) {
            const double* row = s_stg + s * kEvRow;
            const double x2 = row[2], x3 = row[3], x4 = row[4], x5 = row[5], rr = row[6];
            acc[0] += x2 * x2; acc[1] += x2 * x3; acc[2] += x2 * x4; acc[3] += x2 * x5; acc[4] += x2 * rr;
            acc[5] += x5 * x5; acc[6] += x5 * rr;
          }
          Vo[11] = acc[0]; Vo[12] = acc[1]; Vo[13] = acc[2]; Vo[14] = acc[3]; go[2] = acc[4];
          Vo[20] = acc[5]; go[5] = acc[6]; g0 = acc[4]; g1 = acc[6];
        } else {
#pragma unroll 4
          for (int s = s0; s < s1; ++s) {
            const double* row = s_stg + s * kEvRow;
            const double x3 = row[3], x4 = row[4], x5 = row[5], rr = row[6];
            acc[0] += x3 * x3; acc[1] += x3 * x4; acc[2] += x3 * x5; acc[3] += x3 * rr;
            acc[4] += x4 * x4; acc[5] += x4 * x5; acc[6] += x4 * rr;
          }
          Vo[15] = acc[0]; Vo[16] = acc[1]; Vo[17] = acc[2]; go[3] = acc[3];
          Vo[18] = acc[4]; Vo[19] = acc[5]; go[4] = acc[6]; g0 = acc[3]; g1 = acc[6];
        }
        if (frag) continue;   // partial gradient of a long line: k_frag_blocks takes its norm
        const double ab0 = fabs(g0), ab1 = fabs(g1);
        gmax = (ab0 > gmax || ab0 != ab0) ? ab0 : gmax;   // NaN propagates
        gmax = (ab1 > gmax || ab1 != ab1) ? ab1 : gmax;
      }
    }
    if (!deferred) {
      fence_proxy_async();
      sub_sync<NSUB>(sub);  // the stage is free: refill it with tile it + S
      if (t == 0) {
        if (it + S < my_n) issue(it + S, nx_nl, nx_l0);
        nx_l0 = cur_l0; nx_nl = cur_nl;
      }
    }
  }
  if (WITH_JAC) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, gmax, o);
      gmax = (other > gmax || other != other) ? other : gmax;
    }
    if ((threadIdx.x & 31) == 0 && gmax != 0.0) atomicMax(a.gbits.sel(set), (unsigned long long)__double_as_longlong(gmax));
  }
  const double cs[1] = {block_sum<kTile * NSUB>(cost_acc, red)};
  last_cta_reduce<kTile * NSUB, 1>(cs, a.cost_partial, a.ticket, a.cost_out, 0.5, red);
}

// ---- helpers of the Jacobian evaluation kernel (k_eval_z below) ---------------------------------------------------
// 16 B / 32 B global stores (naturally aligned); the 256-bit form is sm_100's STG.256
__device__ __forceinline__ void st_global_v2(double* p, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// One half of the 27 sums of a line: rows [P0, P1) of the upper triangle of V_l and the same entries of g_l
// (P0 = 0, P1 = 2: 11 + 2 sums; P0 = 2, P1 = 6: 10 + 4 sums), each sum one FMA chain in slot order.  NV values of V_l
// starting at packed index sym_idx(P0, P0), then P1 - P0 values of g_l.
template <int ROW, int P0, int P1>
__device__ __forceinline__ void line_block_half(const double* __restrict__ s_stg, int s0, int s1,
                                                double (&v)[(P1 - P0) * 6 - (P1 * (P1 - 1) - P0 * (P0 - 1)) / 2], double (&g)[P1 - P0]) {
  constexpr int NV = (P1 - P0) * 6 - (P1 * (P1 - 1) - P0 * (P0 - 1)) / 2;
#pragma unroll
  for (int q = 0; q < NV; ++q) v[q] = 0.0;
#pragma unroll
  for (int q = 0; q < P1 - P0; ++q) g[q] = 0.0;
#pragma unroll 2
  for (int s = s0; s < s1; ++s) {
    const double* row = s_stg + s * ROW;
    double x[6 - P0];
#pragma unroll
    for (int q = 0; q < 6 - P0; ++q) x[q] = row[P0 + q];
    const double rr = row[6];
    int e = 0;
#pragma unroll
    for (int p = P0; p < P1; ++p) {
#pragma unroll
      for (int c = p; c < 6; ++c) { v[e] += x[p - P0] * x[c - P0]; ++e; }
      g[p - P0] += x[p - P0] * rr;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_eval_z: the Jacobian evaluation with the load pipeline decoupled from the per-line sums.
//
// Its predecessor (round 2's first "rotating summing warps" kernel, 112 us on config 4) kept the staging rows of the
// per-line sums INSIDE the load stage, so a stage was refilled only after the owning warp had finished its rows of
// the NEXT tile and then the sums of this one.  Clock counters showed the consequence: 39 % of every warp's cycles
// waiting for tile data although a bulk load takes ~3000 cycles from issue to arrival -- arrival of tile k+2 =
// arrival of tile k + rows + sums + load, whatever the number of stages.  Here
//  * a stage is dead as soon as every warp holds its slot's constants in registers, a few hundred cycles after the
//    tile has arrived, and is refilled right then (seven warps bar.arrive, a rotating issuer warp bar.syncs, keeps
//    the tile's line table in a small ring and issues the bulk copies of tile k+2): two stages suffice;
//  * the staging rows (jl, r: 56 B per slot) have two buffers of their own; the sums of tile k are taken one tile late
//    by two warps (half the 27 sums each), which release the buffer through an mbarrier that the writers of tile k+2
//    wait on, and leave through a small output buffer and the bulk-copy engine (shared -> global);
//  * the shared-memory camera table holds 64 B records (w, c, A, B); C is recomputed per observation.
// Measured on config 4 (ablation builds, same run): the arithmetic is free (replacing it by moves changes nothing);
// without the jc scatter 80 us, without the jl stores 83 us, without both 76 us, all of them at ~4.2 TB/s of DRAM
// traffic -- the kernel sits on the LSU data pipe (60 % busy, two thirds of it shared-memory loads) and the memory
// system together, not on any one of them.  The order of every sum is that of k_eval<true>.
constexpr int kEzRow = 7;        // staging row (jl[6], r): 56 B, an odd number of 8 B words
constexpr int kEzMetaRing = 4;   // line tables of the last four tiles
__host__ __device__ inline int eval_z_stg_bytes() { return kTile * kEzRow * 8; }
constexpr int kEzOutBytes = 32 * (kVStride + 6) * 8;   // V_l | g_l rows of one batch of 32 lines, laid out as in global memory
constexpr int kEzRecBytes = 64;  // shared-memory camera record: w, c, A, B (C is recomputed: the table is what limits the camera count)
__host__ __device__ inline int eval_z_sub_bytes(int lmax) {
  return 2 * eval_stage_bytes(lmax) + 2 * eval_z_stg_bytes() + kEzOutBytes + ((kEzMetaRing * (int)sizeof(TileMeta) + 127) & ~127);
}
__host__ __device__ inline int eval_z_tab_bytes(int nc) { return (nc * kEzRecBytes + 127) & ~127; }
__host__ __device__ inline int eval_z_smem_bytes(int lmax, int nsub, int nc_tab) {
  return eval_z_tab_bytes(nc_tab) + nsub * eval_z_sub_bytes(lmax) + 128;
}
// shared -> global bulk copy (16 B aligned, size % 16 == 0), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NSUB, bool TAB_SMEM>
__global__ void __launch_bounds__(kTile * NSUB, TAB_SMEM ? 1 : 2) k_eval_z(const EvalArgs a) {
  constexpr int kThreads = kTile * NSUB;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red[kThreads / 32];
  __shared__ uint64_t full[NSUB][2], sfree[NSUB][2], ofree[NSUB];
  const int stage_bytes = eval_stage_bytes(a.lmax);
  const int sub = threadIdx.x / kTile, t = threadIdx.x % kTile, warp = t >> 5, lane = t & 31;
  const int tab_bytes = TAB_SMEM ? eval_z_tab_bytes(a.nc) : 0;
  unsigned char* sbase = smem + tab_bytes + (size_t)sub * eval_z_sub_bytes(a.lmax);
  unsigned char* gbase = sbase + 2 * stage_bytes;                       // two staging buffers
  double* obuf = reinterpret_cast<double*>(gbase + 2 * eval_z_stg_bytes());   // V_l | g_l rows on their way out
  TileMeta* mring = reinterpret_cast<TileMeta*>(gbase + 2 * eval_z_stg_bytes() + kEzOutBytes);
  const int first = blockIdx.x * NSUB + sub, stride = gridDim.x * NSUB;
  const int my_n = first < a.ntiles ? (a.ntiles - first + stride - 1) / stride : 0;
  // named barriers of the sub-group, from bar0: + s: stage s consumed;  + 2 + b: rows of staging buffer b staged;
  // + 4 + b: the two summing warps of the tile in buffer b
  const int bar0 = 1 + sub * 6;

  if (t == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full[sub][s], 1); mbar_init(&sfree[sub][s], 2); }
    mbar_init(&ofree[sub], 1);
    fence_barrier_init();
  }
  pdl_prologue();
  if (ctl_skip_trial(a.ctl)) return;   // uniform over the grid
  const int set = ctl_cur(a.ctl) ^ a.which;
  const CamRec* g_recs = a.recs.sel(set);
  const double* g_pts = a.pts.sel(set);
  double* g_r = a.r.sel(set);
  double* g_jl = a.jl.sel(set);
  double* g_jc = a.jc.sel(set);
  double* g_V = a.V.sel(set);
  double* g_gl = a.gl.sel(set);
  if (TAB_SMEM) {   // w, c, A, B of every camera: the first four of the record's five 16 B words
    const double2* src = reinterpret_cast<const double2*>(g_recs);
    double2* dst = reinterpret_cast<double2*>(smem);
    for (int q = threadIdx.x; q < a.nc * 4; q += kThreads) dst[q] = __ldg(src + (q >> 2) * 5 + (q & 3));
  }
  __syncthreads();

  auto issue = [&](int it, int nl, int l0) {
    const int tile = first + it * stride;
    unsigned char* st = sbase + (size_t)(it & 1) * stage_bytes;
    uint64_t* bar = &full[sub][it & 1];
    mbar_expect_tx(bar, 9 * kTile * 8 + (uint32_t)sizeof(TileIdx) + nl * 48);
    bulk_g2s(st, a.cst + (size_t)tile * 9 * kTile, 9 * kTile * 8, bar);
    bulk_g2s(st + kEvOffIdx, a.tidx + tile, (uint32_t)sizeof(TileIdx), bar);
    if (nl > 0) bulk_g2s(st + kEvOffPts, g_pts + (size_t)l0 * 6, nl * 48, bar);
  };
  if (t == 0) {
    for (int it = 0; it < 2 && it < my_n; ++it) {
      const TileMeta* mt = &a.tidx[first + it * stride].meta;
      issue(it, mt->nl, mt->l0);
    }
  }

  double cost_acc = 0.0, gmax = 0.0;
  // Per-line blocks of tile `it` from its staging buffer, by TWO warps (half 0: warp it mod 8, rows 0-1 of V_l and
  // g_l[0..1]; half 1: warp (it + 4) mod 8, rows 2-5 and g_l[2..5]): a summing warp falls behind the others by the time
  // its sums take, and the stage refill two tiles on waits for the last warp to pick up its constants -- half the sums,
  // half the delay.  The buffer is released (mbarrier, two arrivals) as soon as the last batch of lines has been summed.
  // The sums of a line sit in one lane; stored from there every instruction would touch one 128 B line per lane, so both
  // warps lay their entries of the batch's V_l rows (176 B each, as in global memory) and g_l rows (48 B) out in a small
  // buffer and lane 0 of half 0 hands the two blocks to the bulk-copy engine (shared -> global).
  auto line_sums = [&](int it, int half) {
    const int b = it & 1;
    const double* stg = reinterpret_cast<const double*>(gbase + (size_t)b * eval_z_stg_bytes());
    const TileMeta* s_meta = mring + (it & (kEzMetaRing - 1));
    const int pair_id = bar0 + 4 + b;
    named_sync(bar0 + 2 + b, kTile);   // every row of the tile is staged (and its line table is in the ring)
    const int l0 = s_meta->l0, nlg = s_meta->nl, frag = s_meta->frag;
    if (!frag) {
      for (int lb = 0; lb < nlg; lb += 32) {
        const int ll = lb + lane;
        const int nb = nlg - lb < 32 ? nlg - lb : 32;
        const bool live = ll < nlg;
        const int s0 = live && ll ? s_meta->lend[ll - 1] : 0, s1 = live ? s_meta->lend[ll] : 0;
        double* sg = obuf + nb * kVStride;
        auto buffer_free = [&]() {   // the rows of the previous batch / tile have left the buffer
          if (lb == 0) { if (it >= 1) mbar_wait(&ofree[sub], (it - 1) & 1); }
          else {
            if (half == 0) { if (lane == 0) bulk_wait_read(); __syncwarp(); }
            named_sync(pair_id, 64);
          }
        };
        if (half == 0) {
          double v[11], g[2];
          line_block_half<kEzRow, 0, 2>(stg, s0, s1, v, g);
          if (lb + 32 >= nlg) { __syncwarp(); if (lane == 0) mbar_arrive(&sfree[sub][b]); }
          buffer_free();
          if (live) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const double ab = fabs(g[q]);
              gmax = (ab > gmax || ab != ab) ? ab : gmax;   // NaN propagates
            }
#pragma unroll
            for (int q = 0; q < 11; ++q) obuf[lane * kVStride + q] = v[q];
            sg[lane * 6] = g[0]; sg[lane * 6 + 1] = g[1];
          }
        } else {
          double v[10], g[4];
          line_block_half<kEzRow, 2, 6>(stg, s0, s1, v, g);
          if (lb + 32 >= nlg) { __syncwarp(); if (lane == 0) mbar_arrive(&sfree[sub][b]); }
          buffer_free();
          if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const double ab = fabs(g[q]);
              gmax = (ab > gmax || ab != ab) ? ab : gmax;
            }
#pragma unroll
            for (int q = 0; q < 10; ++q) obuf[lane * kVStride + 11 + q] = v[q];
#pragma unroll
            for (int q = 0; q < 4; ++q) sg[lane * 6 + 2 + q] = g[q];
          }
        }
        fence_proxy_async();
        named_sync(pair_id, 64);   // both halves of the batch are in the buffer
        if (half == 0 && lane == 0) {
          bulk_s2g(g_V + (size_t)(l0 + lb) * kVStride, obuf, nb * kVStride * 8);
          bulk_s2g(g_gl + (size_t)(l0 + lb) * 6, sg, nb * 48);
          bulk_commit();
        }
      }
      if (nlg == 0) { __syncwarp(); if (lane == 0) mbar_arrive(&sfree[sub][b]); }
      if (half == 0 && lane == 0) { bulk_wait_read(); mbar_arrive(&ofree[sub]); }
    } else {
      // fragment of a long line: lane k of half 0 owns sum k over the tile's slots (slot order); k_frag_blocks adds the
      // fragments and takes the gradient norm
      if (half == 0 && lane < 27) {
        const int s1 = s_meta->lend[0];
        const int pa = lane < 21 ? c_sym_a[lane] : lane - 21, pb = lane < 21 ? c_sym_b[lane] : 6;
        double acc = 0.0;
        for (int s = 0; s < s1; ++s) acc += stg[s * kEzRow + pa] * stg[s * kEzRow + pb];
        a.fragV[(size_t)(frag - 1) * 27 + lane] = acc;
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(&sfree[sub][b]); if (half == 0) mbar_arrive(&ofree[sub]); }
    }
  };

  int nx_l0 = 0, nx_nl = 0;   // (l0, nl) of the tile this warp will issue next (lane 0)
  if (lane == 0 && warp == 4 && 2 < my_n) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(&a.tidx[first + 2 * stride].meta));
    nx_l0 = v.x; nx_nl = v.y;
  }
  for (int it = 0; it < my_n; ++it) {
    const int tile = first + it * stride;
    const int s = it & 1;
    unsigned char* st = sbase + (size_t)s * stage_bytes;
    const double* s_cst = reinterpret_cast<const double*>(st);
    const TileIdx* s_idx = reinterpret_cast<const TileIdx*>(st + kEvOffIdx);
    const double* s_pts = reinterpret_cast<const double*>(st + kEvOffPts);
    double* stg = reinterpret_cast<double*>(gbase + (size_t)s * eval_z_stg_bytes());
    // the issuer of the NEXT iteration fetches what its bulk copies need now
    if (lane == 0 && warp == ((it + 5) & 7) && it + 3 < my_n) {
      const int2 v = __ldg(reinterpret_cast<const int2*>(&a.tidx[first + (it + 3) * stride].meta));
      nx_l0 = v.x; nx_nl = v.y;
    }
    mbar_wait(&full[sub][s], (it >> 1) & 1);

    // everything this slot needs of the stage, into registers
    const int cam = s_idx->cam[t];
    const size_t i = (size_t)tile * kTile + t;
    int pos = 0;
    double n[3] = {0, 0, 0}, m0[3] = {0, 0, 0}, m1[3] = {0, 0, 0}, X0[3] = {0, 0, 0}, X1[3] = {0, 0, 0};
    if (cam >= 0) {
      pos = s_idx->pos[t];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        n[k] = s_cst[t * 9 + k];
        m0[k] = s_cst[t * 9 + 3 + k];
        m1[k] = s_cst[t * 9 + 6 + k];
      }
      const double* X = s_pts + (int)s_idx->ll[t] * 6;
#pragma unroll
      for (int k = 0; k < 3; ++k) { X0[k] = X[k]; X1[k] = X[3 + k]; }
    }
    if (warp == ((it + 4) & 7)) {
      // issuer: when all eight warps hold their constants, keep the line table and refill the stage with tile it + 2
      named_sync(bar0 + s, kTile);
      if (lane < (int)(sizeof(TileMeta) / 16))
        reinterpret_cast<int4*>(mring + (it & (kEzMetaRing - 1)))[lane] = reinterpret_cast<const int4*>(&s_idx->meta)[lane];
      if (it + 2 < my_n) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) issue(it + 2, nx_nl, nx_l0);
      }
    } else {
      named_arrive(bar0 + s, kTile);
    }

    double r = 0.0, rho_cost = 0.0;
    double jl[6] = {0, 0, 0, 0, 0, 0};
    double jw[3] = {0, 0, 0};
    if (cam >= 0) {
      double w[3], c[3], A, B, C;
      if (TAB_SMEM) {
        const double2* p = reinterpret_cast<const double2*>(smem + (size_t)cam * kEzRecBytes);
        const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
        w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; c[0] = v1.y; c[1] = v2.x; c[2] = v2.y;
        A = v3.x; B = v3.y;
      } else {
        load_camrec(g_recs, cam, w, c, A, B, C);
      }
      C = coeff_C(w[0] * w[0] + w[1] * w[1] + w[2] * w[2], A);   // both variants: the result does not depend on where the table lives
      double R[9];
      build_R(w, A, B, R);
      double tq[3] = {0, 0, 0};
      const double e0 = endpoint_term<true>(R, c, X0, n, m0, jl, tq);
      const double e1 = endpoint_term<true>(R, c, X1, n, m1, jl + 3, tq);
      r = e0 + e1;
      if (a.loss) {   // scipy's robust-loss scaling, as in k_eval
        const double rs = r / a.f_scale;
        const double z = rs * rs;
        double r0, r1, r2;
        robust_rho(a.loss, z, r0, r1, r2);
        rho_cost = a.f_scale * a.f_scale * r0;
        double js = r1 + 2.0 * r2 * z;
        js = sqrt(js < 2.220446049250313e-16 ? 2.220446049250313e-16 : js);
        r = r * r1 / js;
#pragma unroll
        for (int k = 0; k < 6; ++k) jl[k] *= js;
        tq[0] *= js; tq[1] *= js; tq[2] *= js;
      } else {
        rho_cost = r * r;
      }
      double G[9];
      build_G(w, B, C, G);
      jw[0] = G[0] * tq[0] + G[3] * tq[1] + G[6] * tq[2];
      jw[1] = G[1] * tq[0] + G[4] * tq[1] + G[7] * tq[2];
      jw[2] = G[2] * tq[0] + G[5] * tq[1] + G[8] * tq[2];
    }
    cost_acc += rho_cost;
    if (cam >= 0) {   // d r/d w = G^T tq ;  d r/d c = -(jl0 + jl1): one 48 B record at the camera-order position
      // an even position starts a 32 B sector, an odd one ends one: one 32 B and one 16 B store, both naturally aligned
      const double c0 = -(jl[0] + jl[3]), c1 = -(jl[1] + jl[4]), c2 = -(jl[2] + jl[5]);
      double* dst = g_jc + (size_t)pos * 6;
      const bool odd = pos & 1;
      st_global_v2(dst + (odd ? 0 : 4), odd ? jw[0] : c1, odd ? jw[1] : c2);
      st_global_v4(dst + (odd ? 2 : 0), odd ? jw[2] : jw[0], odd ? c0 : jw[1], odd ? c1 : jw[2], odd ? c2 : c0);
    }
    g_r[i] = r;
#pragma unroll
    for (int k = 0; k < 6; ++k) g_jl[((size_t)tile * 6 + k) * kTile + t] = jl[k];
    // staging rows: the buffer was last read by the sums of tile it - 2
    if (it >= 2) mbar_wait(&sfree[sub][s], ((it >> 1) - 1) & 1);
#pragma unroll
    for (int k = 0; k < 6; ++k) stg[t * kEzRow + k] = jl[k];
    stg[t * kEzRow + 6] = r;
    if (warp != (it & 7) && warp != ((it + 4) & 7)) named_arrive(bar0 + 2 + s, kTile);
    if (it >= 1) {
      if (warp == ((it - 1) & 7)) line_sums(it - 1, 0);
      else if (warp == ((it + 3) & 7)) line_sums(it - 1, 1);
    }
  }
  if (my_n >= 1) {
    if (warp == ((my_n - 1) & 7)) line_sums(my_n - 1, 0);
    else if (warp == ((my_n + 3) & 7)) line_sums(my_n - 1, 1);
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, gmax, o);
    gmax = (other > gmax || other != other) ? other : gmax;
  }
  if (lane == 0 && gmax != 0.0) atomicMax(a.gbits.sel(set), (unsigned long long)__double_as_longlong(gmax));
  const double cs[1] = {block_sum<kThreads>(cost_acc, red)};
  last_cta_reduce<kThreads, 1>(cs, a.cost_partial, a.ticket, a.cost_out, 0.5, red);
}

// project(): p_hat rows 2i+e in the caller's observation order (API parity only; cold path)
struct ProjectArgs {
  int Mp;
  const int* slot_src;
  const TileIdx* tidx;
  const CamRec* recs;
  const double* pts;
  double* proj;
};
__global__ void k_project(const ProjectArgs a) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.Mp) return;
  const int src = a.slot_src[i];
  if (src < 0) return;
  const TileIdx& ti = a.tidx[i / kTile];
  const int cam = ti.cam[i % kTile], line = ti.meta.l0 + ti.ll[i % kTile];
  double w[3], c[3], A, B, C;
  load_camrec(a.recs, cam, w, c, A, B, C);
  double X[6];
  load6(a.pts + (size_t)line * 6, X);
  double R[9];
  build_R(w, A, B, R);
  const size_t o = (size_t)src;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double d0 = X[3 * e] - c[0], d1 = X[3 * e + 1] - c[1], d2 = X[3 * e + 2] - c[2];
    const double x0 = R[0] * d0 + R[1] * d1 + R[2] * d2;
    const double x1 = R[3] * d0 + R[4] * d1 + R[5] * d2;
    const double x2 = R[6] * d0 + R[7] * d1 + R[8] * d2;
    const double inv = 1.0 / sqrt(x0 * x0 + x1 * x1 + x2 * x2);
    a.proj[(2 * o + e) * 3 + 0] = x0 * inv;
    a.proj[(2 * o + e) * 3 + 1] = x1 * inv;
    a.proj[(2 * o + e) * 3 + 2] = x2 * inv;
  }
}

// ---------------------------------------------------------------------------------
// K3 (camera side): per-camera sums  S_c = sum_j wa_j jc_j jc_j^T (21),  v_c = sum_j wb_j jc_j (6)
// over the camera-ordered Jacobian copy.  wa == nullptr means weight 1.
// One CTA per camera; 27 register accumulators per thread, warp-shuffle + smem reduce.
// ---------------------------------------------------------------------------------
// diag(U_c) and g_c only (all the LM loop needs from U): 12 accumulators instead of 27
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_cam_grad(const int* __restrict__ cam_ptr,
                                                           const int* __restrict__ cam_perm,
                                                           const double* __restrict__ jc, const double* __restrict__ r,
                                                           double* __restrict__ S, double* __restrict__ v,
                                                           const PeerLink pl) {
  pdl_prologue();
  __shared__ double sm[THREADS / 32][12];
  const int c = blockIdx.x;
  const int j0 = cam_ptr[c], j1 = cam_ptr[c + 1];
  double acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;
#pragma unroll 4
  for (int j = j0 + threadIdx.x; j < j1; j += THREADS) {
    double x[6];
    load6(jc + (size_t)j * 6, x);
    const double b = __ldg(r + cam_perm[j]);
#pragma unroll
    for (int p = 0; p < 6; ++p) { acc[p] += x[p] * x[p]; acc[6 + p] += b * x[p]; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const double s = warp_sum(acc[k]);
    if (l == 0) sm[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < THREADS / 32; ++q) s += sm[q][threadIdx.x];
    if (threadIdx.x < 6) S[(size_t)c * 21 + sym_idx(threadIdx.x, threadIdx.x)] = s;
    else v[(size_t)c * 6 + threadIdx.x - 6] = s;
  }
  if (pl.seq) peer_signal_last_cta(pl);   // fused all-reduce: the consumer (k_cam_post) adds the ranks' slots
}

// Per camera: S = sum a_i jc_i jc_i^T (21, a = 1 when wa is NULL) and v = sum b_i jc_i (6).  With GRAD the same
// pass over jc also gives diag(U) = sum jc_i^2 (6) and g_c = sum r_i jc_i (6) at dg / dg + 6 nc, so that an LM
// iteration reads the camera-order Jacobian once for the Schur blocks, the scaling and the gradient.
template <int THREADS, bool GRAD>
__global__ void __launch_bounds__(THREADS) k_cam_accum(const int* __restrict__ cam_ptr,
                                                            const int* __restrict__ cam_perm,
                                                            const Dual<const double> jc2,
                                                            const double* __restrict__ wa,
                                                            const double* __restrict__ wb,
                                                            const Dual<const double> rr2,
                                                            double* __restrict__ S, double* __restrict__ v,
                                                            double* __restrict__ dg, int nc,
                                                            const PeerLink pl, const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  constexpr int NA = GRAD ? 39 : 27;
  __shared__ double sm[THREADS / 32][NA];
  if (ctl_skip_solve(ctl)) return;
  const int cur = ctl_cur(ctl);
  const double* __restrict__ jc = jc2.sel(cur);
  const double* __restrict__ rr = rr2.sel(cur);
  const int c = blockIdx.x;
  const int j0 = cam_ptr[c], j1 = cam_ptr[c + 1];
  double acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.0;
#pragma unroll 2
  for (int j = j0 + threadIdx.x; j < j1; j += THREADS) {
    const int i = cam_perm[j];
    double x[6];
    load6(jc + (size_t)j * 6, x);
    const double a = wa ? __ldg(wa + i) : 1.0;
    const double b = __ldg(wb + i);
    const double ri = GRAD ? __ldg(rr + i) : 0.0;
    int e = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const double ax = a * x[p];
#pragma unroll
      for (int q = p; q < 6; ++q) acc[e++] += ax * x[q];
    }
#pragma unroll
    for (int p = 0; p < 6; ++p) acc[21 + p] += b * x[p];
    if (GRAD) {
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        acc[27 + p] += x[p] * x[p];
        acc[33 + p] += ri * x[p];
      }
    }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    const double s = warp_sum(acc[k]);
    if (l == 0) sm[w][k] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < NA; k += THREADS) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < THREADS / 32; ++q) s += sm[q][k];
    if (k < 21) S[(size_t)c * 21 + k] = s;
    else if (k < 27) v[(size_t)c * 6 + k - 21] = s;
    else if (k < 33) dg[(size_t)c * 6 + k - 27] = s;
    else dg[(size_t)nc * 6 + (size_t)c * 6 + k - 33] = s;
  }
  if (pl.seq) peer_signal_last_cta(pl);   // fused all-reduce: the consumer (k_cam_finish) adds the ranks' slots
}

// after the all-reduce of U|gc: camera scaling (Marquardt: D = diag(J^T J), zeros -> 1, running max --
// scipy compute_jac_scale, _lsq/common.py:598-610, used by x_scale='jac' at :319) and the camera part of
// the gradient inf-norm, in one launch
__global__ void k_cam_post(int nc, const double* __restrict__ S, const double* __restrict__ gc_in, double* __restrict__ gc_out,
                           double* __restrict__ D, int first, unsigned long long* __restrict__ gbits, const PeerLink pl) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double mx = 0.0;
  if (pl.seq) peer_wait(pl);   // multi-GPU: S | gc live in the symmetric slots of all ranks
  if (t < nc * 6) {
    const int b = t / 6, k = t % 6;
    const size_t ed = (size_t)b * 21 + sym_idx(k, k), eg = (size_t)nc * 21 + t;
    double d = pl.seq ? peer_sum(pl, ed) : S[ed];
    const double g = pl.seq ? peer_sum(pl, eg) : gc_in[t];
    if (d == 0.0) d = 1.0;
    D[t] = first ? d : fmax(D[t], d);
    gc_out[t] = g;
    mx = fabs(g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (other > mx || other != other) ? other : mx;
  }
  if ((threadIdx.x & 31) == 0 && mx != 0.0) atomicMax(gbits, (unsigned long long)__double_as_longlong(mx));
}

// ---------------------------------------------------------------------------------
// K4: Schur elimination pieces
// ---------------------------------------------------------------------------------
// per line: Vinv = (V + lam Dl)^-1, y = Vinv gl.  scaling_mode: 0 keep Dl, 1 Dl = max(Dl, diag V),
// 2 Dl = diag V (first Jacobian); zeros -> 1  (Marquardt scaling, the analogue of scipy's
// compute_jac_scale, _lsq/common.py:598-610, used by x_scale='jac' at bundle_adjustment.py:319).
constexpr int kLfLines = 128;  // lines per CTA of k_line_factor
__global__ void __launch_bounds__(kLfLines) k_line_factor(int nl, double lam, const Dual<const double> V2,
                                                           double* __restrict__ Dl, const Dual<const double> gl2,
                                                           double* __restrict__ Vinv, double* __restrict__ y,
                                                           int* __restrict__ clamped, int scaling_mode,
                                                           const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  if (ctl_skip_solve(ctl)) return;
  if (ctl) { lam = ctl->lam; scaling_mode = ctl->scaling_mode; }
  const double* __restrict__ V = V2.sel(ctl_cur(ctl));
  const double* __restrict__ gl = gl2.sel(ctl_cur(ctl));
  // the per-line records are AoS (21 / 6 / 22 doubles): move them through shared memory so that
  // global traffic is fully coalesced and each thread reads its own row conflict-free (odd strides)
  __shared__ double sV[kLfLines * kVinvStride];
  __shared__ double sD[kLfLines * 6];
  __shared__ double sg[kLfLines * 6];
  const int l0 = blockIdx.x * kLfLines;
  const int n = min(kLfLines, nl - l0);
  const int t = threadIdx.x;
  for (int q = t; q < n * kVStride; q += kLfLines) sV[q] = __ldg(V + (size_t)l0 * kVStride + q);
  for (int q = t; q < n * 6; q += kLfLines) {
    sg[q] = __ldg(gl + (size_t)l0 * 6 + q);
    if (scaling_mode != 2) sD[q] = Dl[(size_t)l0 * 6 + q];
  }
  __syncthreads();
  double S[21], yy[6];
  int cl = 0;
  if (t < n) {
#pragma unroll
    for (int e = 0; e < 21; ++e) S[e] = sV[t * kVStride + e];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double D;
      if (scaling_mode == 0) {
        D = sD[t * 6 + k];
      } else {
        double d = S[sym_idx(k, k)];
        if (d == 0.0) d = 1.0;
        D = scaling_mode == 2 ? d : fmax(sD[t * 6 + k], d);
        sD[t * 6 + k] = D;
      }
      S[sym_idx(k, k)] += lam * D;
    }
    cl = spd6_inverse(S);
    double g[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = sg[t * 6 + k];
    sym6_matvec(S, g, yy);
  }
  if (cl) atomicAdd(clamped, cl);
  __syncthreads();
  if (t < n) {
#pragma unroll
    for (int e = 0; e < 21; ++e) sV[t * kVinvStride + e] = S[e];
    sV[t * kVinvStride + 21] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) sg[t * 6 + k] = yy[k];
  }
  __syncthreads();
  for (int q = t; q < n * kVinvStride; q += kLfLines) Vinv[(size_t)l0 * kVinvStride + q] = sV[q];
  for (int q = t; q < n * 6; q += kLfLines) {
    y[(size_t)l0 * 6 + q] = sg[q];
    if (scaling_mode != 0) Dl[(size_t)l0 * 6 + q] = sD[q];
  }
}

// per camera: Minv = (lam Dc + Mpart)^-1 (block-Jacobi), b = -gc + bpart (Schur rhs).  When the accumulation
// pass also produced diag(U) | g_c (dg != NULL, all-reduced with Mpart | bpart) this kernel first does what
// k_cam_post does for the stand-alone gradient pass: Marquardt scaling Dc = max(Dc, diag U) (zeros -> 1),
// g_c, and the camera part of the gradient inf-norm.
__global__ void __launch_bounds__(128) k_cam_finish(int nc, double lam, const double* __restrict__ Mpart,
                                                     double* __restrict__ Dc, double* __restrict__ gc,
                                                     const double* __restrict__ bpart, const double* __restrict__ dg,
                                                     const Dual<unsigned long long> gbits2,
                                                     double* __restrict__ Minv,
                                                     double* __restrict__ b, int* __restrict__ clamped, const PeerLink pl,
                                                     const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  if (ctl_skip_solve(ctl)) return;
  if (ctl) lam = ctl->lam;
  unsigned long long* gbits = gbits2.sel(ctl_cur(ctl));
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl.seq) peer_wait(pl);   // multi-GPU: Mpart | bpart live in the symmetric slots of all ranks
  double mx = 0.0;
  double D[6], g[6];
  if (c < nc) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      D[k] = Dc[(size_t)c * 6 + k];
      if (dg) {
        double d = dg[(size_t)c * 6 + k];
        if (d == 0.0) d = 1.0;
        D[k] = fmax(D[k], d);
        Dc[(size_t)c * 6 + k] = D[k];
        g[k] = dg[(size_t)nc * 6 + (size_t)c * 6 + k];
        gc[(size_t)c * 6 + k] = g[k];
        const double ab = fabs(g[k]);
        mx = (ab > mx || ab != ab) ? ab : mx;   // NaN propagates
      } else {
        g[k] = gc[(size_t)c * 6 + k];
      }
    }
  }
  if (dg) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, mx, o);
      mx = (other > mx || other != other) ? other : mx;
    }
    if ((threadIdx.x & 31) == 0 && mx != 0.0) atomicMax(gbits, (unsigned long long)__double_as_longlong(mx));
  }
  if (c >= nc) return;
  double S[21];
#pragma unroll
  for (int e = 0; e < 21; ++e) S[e] = pl.seq ? peer_sum(pl, (size_t)c * 21 + e) : Mpart[(size_t)c * 21 + e];
#pragma unroll
  for (int k = 0; k < 6; ++k) S[sym_idx(k, k)] += lam * D[k];
  const int cl = spd6_inverse(S);
  if (cl) atomicAdd(clamped, cl);
#pragma unroll
  for (int e = 0; e < 21; ++e) Minv[(size_t)c * 21 + e] = S[e];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double bp = pl.seq ? peer_sum(pl, (size_t)nc * 21 + (size_t)c * 6 + k) : bpart[(size_t)c * 6 + k];
    b[(size_t)c * 6 + k] = bp - g[k];
  }
}

// ---------------------------------------------------------------------------------
// K5: matrix-free Schur product and PCG
// ---------------------------------------------------------------------------------
struct CgState {
  double rz, rz0, pq;
  int k, done, breakdown, pad;
};

// per camera gather table for a camera-space vector v:  a = G v_w,  h = a x c + v_c, so that
// jc_i . v_c = sum_e jl_e . (a x X_e - h)   (jc is never read in line order).
__device__ __forceinline__ void make_tab_entry(const double* __restrict__ G, const double* __restrict__ cam,
                                               const double v[6], double* __restrict__ tab) {
  const double a0 = G[0] * v[0] + G[1] * v[1] + G[2] * v[2];
  const double a1 = G[3] * v[0] + G[4] * v[1] + G[5] * v[2];
  const double a2 = G[6] * v[0] + G[7] * v[1] + G[8] * v[2];
  const double c0 = cam[3], c1 = cam[4], c2 = cam[5];
  tab[0] = a0; tab[1] = a1; tab[2] = a2;
  tab[3] = a1 * c2 - a2 * c1 + v[3];
  tab[4] = a2 * c0 - a0 * c2 + v[4];
  tab[5] = a0 * c1 - a1 * c0 + v[5];
}

struct LinePassArgs {
  int Mp, ntiles, lmax;
  const TileIdx* tidx;   // [ntiles]
  Dual<const double> jl; // [ntiles][6][256]   (both evaluation sets; the kernel reads the current one)
  Dual<double> pts;      // [nl][6]            (BACKSUB writes the trial point into the other set)
  const double* tab;     // [nc][6]
  const double* Vinv;    // [nl][22]
  double* w;             // MATVEC out [Mp];  SCALARS out: wk
  double* w2;            // SCALARS out: wr
  const double* y;       // SCALARS: [nl][6]
  // BACKSUB only
  Dual<const double> gl; // [nl][6]
  Dual<const double> r;  // [Mp]
  double* dl;            // [nl][6] the full line step (kept for a retry at a shorter length)
  double alpha;          // step length of the trial point (LM loop: taken from ctl)
  double* partial;       // [gridDim][4]: r.jd, jd.jd, |dl|^2, |pts|^2
  unsigned int* ticket;  // BACKSUB: last-CTA reduction
  double* out4;          // BACKSUB: the four sums
  const CgState* st;     // MATVEC: early exit when done
  const LmCtl* ctl;      // LM loop: control block (NULL for the stand-alone API calls)
  double* fragT;         // [n_fragments][6] partial t_l of the fragments of long lines (k_frag_matvec / k_frag_backsub)
};

enum { kModeMatvec = 0, kModeBacksub = 1, kModeScalars = 2 };

// stage: jl[6][256] | TileIdx prefix (meta, ll, cam) | pts[lmax][6] | Vinv[lmax][22] | aux[lmax][6] | r[256]
//   aux = gl (BACKSUB) or y (SCALARS)
constexpr int kLpOffIdx = 6 * kTile * 8;
constexpr int kLpOffPts = kLpOffIdx + (int)offsetof(TileIdx, pos);
__host__ __device__ inline int lp_off_vinv(int lmax) { return kLpOffPts + lmax * 48; }
__host__ __device__ inline int lp_off_aux(int lmax) { return lp_off_vinv(lmax) + lmax * kVinvStride * 8; }
__host__ __device__ inline int lp_off_r(int lmax) { return lp_off_aux(lmax) + lmax * 48; }
__host__ __device__ inline int lp_stage_bytes(int lmax, int mode) {
  const int end = mode == kModeMatvec ? lp_off_aux(lmax) : (mode == kModeScalars ? lp_off_r(lmax) : lp_off_r(lmax) + kTile * 8);
  return (end + 127) & ~127;
}
constexpr int kLpRow = 7;  // staging row stride in doubles (6 + pad): 56 B, conflict-free per half-warp
__host__ __device__ inline int lp_smem_bytes(int lmax, int mode) {
  return lp_stages(mode) * lp_stage_bytes(lmax, mode) + (mode == kModeScalars ? 0 : kTile * kLpRow * 8);
}

// Line-ordered half of S*p (MATVEC), the back-substitution (BACKSUB) and the per-observation
// Schur weights (SCALARS).
//   s_i = jc_i . v_{c_i} (via the table),  t_l = sum_{i in l} jl_i s_i,
//   MATVEC:  u_l = Vinv_l t_l,           w_i = s_i - jl_i . u_l
//   BACKSUB: dl  = -Vinv_l (gl_l + t_l), (J delta)_i = s_i + jl_i . dl
//   SCALARS: wk_i = 1 - jl_i^T Vinv_l jl_i,  wr_i = jl_i . y_l
template <int MODE>
__global__ void __launch_bounds__(kTile, 3) k_line_pass(const LinePassArgs a) {
  constexpr int kLpStages = lp_stages(MODE);
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red[kTile / 32];
  __shared__ uint64_t full[kLpStages];
  const int stage_bytes = lp_stage_bytes(a.lmax, MODE);
  double* s_stg = reinterpret_cast<double*>(smem + (size_t)kLpStages * stage_bytes);  // [256][kLpRow]
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_n = first < a.ntiles ? (a.ntiles - first + stride - 1) / stride : 0;
  const int t = threadIdx.x;
  const int off_vinv = lp_off_vinv(a.lmax), off_aux = lp_off_aux(a.lmax), off_r = lp_off_r(a.lmax);

  if (t == 0) {
    for (int s = 0; s < kLpStages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  pdl_prologue();
  if (MODE == kModeMatvec && a.st->done) return;  // uniform over the grid
  if (MODE == kModeScalars && ctl_skip_solve(a.ctl)) return;
  if (MODE == kModeBacksub && ctl_skip_backsub(a.ctl)) return;
  const int cur = ctl_cur(a.ctl);
  const double* g_jl = a.jl.sel(cur);
  const double* g_pts = a.pts.sel(cur);
  const double* g_gl = a.gl.sel(cur);
  const double* g_r = a.r.sel(cur);
  double* g_pts_new = a.pts.sel(cur ^ 1);
  const double alpha = (MODE == kModeBacksub && a.ctl) ? a.ctl->alpha : a.alpha;
  __syncthreads();

  auto issue = [&](int it, int nl, int l0) {
    const int tile = first + it * stride;
    unsigned char* st = smem + (size_t)(it % kLpStages) * stage_bytes;
    uint64_t* bar = &full[it % kLpStages];
    const uint32_t idx_bytes = MODE == kModeScalars ? (uint32_t)offsetof(TileIdx, cam) : (uint32_t)offsetof(TileIdx, pos);
    uint32_t bytes = 6 * kTile * 8 + idx_bytes + nl * kVinvStride * 8;
    if (MODE != kModeScalars) bytes += nl * 48;
    if (MODE != kModeMatvec) bytes += nl * 48;
    if (MODE == kModeBacksub) bytes += kTile * 8;
    mbar_expect_tx(bar, bytes);
    const size_t o = (size_t)tile * kTile;
    bulk_g2s(st, g_jl + (size_t)tile * 6 * kTile, 6 * kTile * 8, bar);
    bulk_g2s(st + kLpOffIdx, a.tidx + tile, idx_bytes, bar);
    if (nl > 0) {
      if (MODE != kModeScalars) bulk_g2s(st + kLpOffPts, g_pts + (size_t)l0 * 6, nl * 48, bar);
      bulk_g2s(st + off_vinv, a.Vinv + (size_t)l0 * kVinvStride, nl * kVinvStride * 8, bar);
      if (MODE == kModeBacksub) bulk_g2s(st + off_aux, g_gl + (size_t)l0 * 6, nl * 48, bar);
      if (MODE == kModeScalars) bulk_g2s(st + off_aux, a.y + (size_t)l0 * 6, nl * 48, bar);
    }
    if (MODE == kModeBacksub) bulk_g2s(st + off_r, g_r + o, kTile * 8, bar);
  };
  int nx_l0 = 0, nx_nl = 0;
  if (t == 0) {
    for (int it = 0; it < kLpStages && it < my_n; ++it) {
      const TileMeta* mt = &a.tidx[first + it * stride].meta;
      issue(it, mt->nl, mt->l0);
    }
    if (kLpStages < my_n) { const TileMeta* mt = &a.tidx[first + kLpStages * stride].meta; nx_l0 = mt->l0; nx_nl = mt->nl; }
  }

  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;  // BACKSUB partial sums
  for (int it = 0; it < my_n; ++it) {
    const int tile = first + it * stride;
    unsigned char* st = smem + (size_t)(it % kLpStages) * stage_bytes;
    double* s_jl = reinterpret_cast<double*>(st);
    const TileIdx* s_idx = reinterpret_cast<const TileIdx*>(st + kLpOffIdx);
    const int* s_cam = s_idx->cam;
    const unsigned char* s_ll = s_idx->ll;
    const TileMeta* s_meta = &s_idx->meta;
    const double* s_pts = reinterpret_cast<const double*>(st + kLpOffPts);
    // u_l of the tile's lines takes the place of their endpoints in the stage: the endpoints are dead after the first
    // barrier (BACKSUB reads x and writes u at the same index from the same thread), and 3 x this much less shared
    // memory per SM is what keeps the MATVEC pass under the 164 KB carve-out, i.e. leaves the L1 32 KB larger
    double* ul = reinterpret_cast<double*>(st + kLpOffPts);
    const double* s_vinv = reinterpret_cast<const double*>(st + off_vinv);
    const double* s_aux = reinterpret_cast<const double*>(st + off_aux);
    const double* s_r = reinterpret_cast<const double*>(st + off_r);
    int cur_l0 = 0, cur_nl = 0;
    if (t == 0 && it + kLpStages + 1 < my_n) {
      const TileMeta* mt = &a.tidx[first + (it + kLpStages + 1) * stride].meta;
      cur_l0 = mt->l0; cur_nl = mt->nl;
    }
    mbar_wait(&full[it % kLpStages], (it / kLpStages) & 1);

    const size_t i = (size_t)tile * kTile + t;
    const int nobs = s_meta->nobs;
    const bool active = t < nobs;
    const int ll = active ? (int)s_ll[t] : 0;
    double jl[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) jl[k] = s_jl[k * kTile + t];

    if (MODE == kModeScalars) {
      double q = 0.0, d = 0.0;
      if (active) {
        const double* S = s_vinv + ll * kVinvStride;
        int e = 0;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          q += S[e] * jl[p] * jl[p];
          ++e;
#pragma unroll
          for (int s = p + 1; s < 6; ++s) { q += 2.0 * S[e] * jl[p] * jl[s]; ++e; }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) d += jl[k] * s_aux[ll * 6 + k];
      }
      a.w[i] = active ? 1.0 - q : 0.0;
      a.w2[i] = d;
    } else {
      double s = 0.0;
      if (active) {
        const int cam = s_cam[t];
        double tb[6];
        load6(a.tab + (size_t)cam * 6, tb);
        const double* X = s_pts + ll * 6;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double* Xe = X + 3 * e;
          const double v0 = tb[1] * Xe[2] - tb[2] * Xe[1] - tb[3];
          const double v1 = tb[2] * Xe[0] - tb[0] * Xe[2] - tb[4];
          const double v2 = tb[0] * Xe[1] - tb[1] * Xe[0] - tb[5];
          s += jl[3 * e] * v0 + jl[3 * e + 1] * v1 + jl[3 * e + 2] * v2;
        }
      }
      // contributions jl_i s_i staged row-major with an odd-word stride (bank-conflict free)
#pragma unroll
      for (int k = 0; k < 6; ++k) s_stg[t * kLpRow + k] = jl[k] * s;
      __syncthreads();
      // 8 lanes per line: lane b < 6 sums component b, then the 6x6 product by shuffles
      {
        const int sub = t & 7, grp = t >> 3;
        const int nlg = s_meta->nl, l0 = s_meta->l0, frag = s_meta->frag;
        const int rounds = (nlg + (kTile / 8) - 1) / (kTile / 8);
        for (int rd = 0; rd < rounds; ++rd) {
          const int lx = rd * (kTile / 8) + grp;
          const bool have = lx < nlg && sub < 6;
          double tv = 0.0;
          if (have) {
            const int s0 = lx ? s_meta->lend[lx - 1] : 0, s1 = s_meta->lend[lx];
            for (int q = s0; q < s1; ++q) tv += s_stg[q * kLpRow + sub];
            if (frag) {
              // fragment of a long line: only the partial sum; w_i = s_i for now, the fix-up kernel finishes the line
              a.fragT[(size_t)(frag - 1) * 6 + sub] = tv;
              tv = 0.0;
            } else if (MODE == kModeBacksub) {
              tv += s_aux[lx * 6 + sub];
            }
          }
          double u = 0.0;
          const double* S = s_vinv + (have ? lx : 0) * kVinvStride;
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            const double tvb = __shfl_sync(0xffffffffu, tv, (t & 24) + b);
            if (have) u += S[sym_idx(sub < b ? sub : b, sub < b ? b : sub)] * tvb;
          }
          if (have) {
            if (MODE == kModeBacksub && !frag) {
              u = -u;
              const double x = s_pts[lx * 6 + sub];
              // trial point x + alpha dl, rounded like the oracle (no FMA); dl is kept for a retry
              g_pts_new[(size_t)(l0 + lx) * 6 + sub] = __dadd_rn(x, __dmul_rn(alpha, u));
              a.dl[(size_t)(l0 + lx) * 6 + sub] = u;
              p2 += u * u;
              p3 += x * x;
            }
            ul[lx * 6 + sub] = u;
          }
        }
      }
      __syncthreads();
      double d = 0.0;
      if (active) {
#pragma unroll
        for (int k = 0; k < 6; ++k) d += jl[k] * ul[ll * 6 + k];
      }
      if (MODE == kModeMatvec) {
        a.w[i] = active ? s - d : 0.0;
      } else if (s_meta->frag) {
        a.w[i] = s;   // kept for k_frag_backsub: (J delta)_i = s_i + jl_i . dl needs the whole line's dl
      } else {
        const double jd = active ? s + d : 0.0;
        const double rr = active ? s_r[t] : 0.0;
        p0 += rr * jd;
        p1 += jd * jd;
      }
    }
    fence_proxy_async();
    __syncthreads();  // the stage (and ul) are free
    if (t == 0) {
      if (it + kLpStages < my_n) issue(it + kLpStages, nx_nl, nx_l0);
      nx_l0 = cur_l0; nx_nl = cur_nl;
    }
  }
  if (MODE == kModeBacksub) {
    const double sums[4] = {block_sum<kTile>(p0, red), block_sum<kTile>(p1, red), block_sum<kTile>(p2, red),
                            block_sum<kTile>(p3, red)};
    last_cta_reduce<kTile, 4>(sums, a.partial, a.ticket, a.out4, 1.0, red);
  }
}

// ---- lines with more than kTile observations --------------------------------------------------------------------
// Such a line occupies ceil(k / kTile) consecutive FRAGMENT tiles of its own (TileMeta::frag).  The tile kernels
// leave the per-line sums of a fragment partial (fragV, fragT); these kernels, one CTA per long line, add the
// fragments in order and finish what needs the whole line.  (The reference has no limit on the observations of a
// line, /root/reference/src/bundle_adjustment.py:194-250.)
struct LongLines {
  int n;               // long lines
  const int* line;     // [n] global line id
  const int* frag0;    // [n+1] first fragment of each long line (fragments are numbered over all long lines)
  const int* tile0;    // [n] first tile of each long line (its fragments are consecutive tiles)
};

// V_l | g_l of a long line = sum of its fragments' partial blocks; line part of the gradient inf-norm
__global__ void __launch_bounds__(32) k_frag_blocks(const LongLines L, const double* __restrict__ fragV, const Dual<double> V2,
                                                    const Dual<double> gl2, const Dual<unsigned long long> gbits2,
                                                    const LmCtl* __restrict__ ctl, int which) {
  pdl_prologue();
  if (ctl_skip_trial(ctl)) return;
  const int set = ctl_cur(ctl) ^ which;
  const int q = blockIdx.x, t = threadIdx.x;
  const int line = L.line[q];
  double v = 0.0;
  if (t < 27) {
    for (int f = L.frag0[q]; f < L.frag0[q + 1]; ++f) v += fragV[(size_t)f * 27 + t];
    if (t < 21) V2.sel(set)[(size_t)line * kVStride + t] = v;
    else gl2.sel(set)[(size_t)line * 6 + t - 21] = v;
  }
  double g = (t >= 21 && t < 27) ? fabs(v) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, g, o);
    g = (other > g || other != other) ? other : g;
  }
  if (t == 0 && g != 0.0) atomicMax(gbits2.sel(set), (unsigned long long)__double_as_longlong(g));
}

// MATVEC of a long line: t_l = sum of the fragments' partial sums, u_l = Vinv_l t_l, w_i -= jl_i . u_l
// (the tile kernel left w_i = s_i)
__global__ void __launch_bounds__(kTile) k_frag_matvec(const LongLines L, const TileIdx* __restrict__ tidx,
                                                       const Dual<const double> jl2, const double* __restrict__ Vinv,
                                                       const double* __restrict__ fragT, double* __restrict__ w,
                                                       const CgState* __restrict__ st, const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  __shared__ double s_t[6], s_u[6];
  if (st->done) return;
  const double* __restrict__ jl = jl2.sel(ctl_cur(ctl));
  const int q = blockIdx.x, t = threadIdx.x;
  const int line = L.line[q], f0 = L.frag0[q], f1 = L.frag0[q + 1], tile0 = L.tile0[q];
  if (t < 6) {
    double v = 0.0;
    for (int f = f0; f < f1; ++f) v += fragT[(size_t)f * 6 + t];
    s_t[t] = v;
  }
  __syncthreads();
  if (t < 6) {
    const double* S = Vinv + (size_t)line * kVinvStride;
    double u = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) u += S[sym_idx(t < b ? t : b, t < b ? b : t)] * s_t[b];
    s_u[t] = u;
  }
  __syncthreads();
  for (int f = 0; f < f1 - f0; ++f) {
    const int tile = tile0 + f;
    if (t >= tidx[tile].meta.nobs) continue;
    double d = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) d += jl[((size_t)tile * 6 + k) * kTile + t] * s_u[k];
    w[(size_t)tile * kTile + t] -= d;
  }
}

// BACKSUB of a long line: dl = -Vinv_l (gl_l + t_l), the trial point, (J delta)_i = s_i + jl_i . dl and the line's
// share of the four sums (added to `out4` by the last CTA, in line order)
__global__ void __launch_bounds__(kTile) k_frag_backsub(const LongLines L, const TileIdx* __restrict__ tidx,
                                                        const Dual<const double> jl2, const double* __restrict__ Vinv,
                                                        const Dual<const double> gl2, const Dual<const double> r2,
                                                        const Dual<double> pts2, double* __restrict__ dl,
                                                        const double* __restrict__ fragT, const double* __restrict__ s_obs,
                                                        double alpha, double* __restrict__ partial /*[n][4]*/,
                                                        unsigned int* ticket, double* __restrict__ out4,
                                                        const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  __shared__ double s_t[6], s_u[6], red[kTile / 32];
  if (ctl_skip_backsub(ctl)) return;
  const int cur = ctl_cur(ctl);
  if (ctl) alpha = ctl->alpha;
  const double* __restrict__ jl = jl2.sel(cur);
  const double* __restrict__ r = r2.sel(cur);
  const int q = blockIdx.x, t = threadIdx.x;
  const int line = L.line[q], f0 = L.frag0[q], f1 = L.frag0[q + 1], tile0 = L.tile0[q];
  if (t < 6) {
    double v = gl2.sel(cur)[(size_t)line * 6 + t];
    for (int f = f0; f < f1; ++f) v += fragT[(size_t)f * 6 + t];
    s_t[t] = v;
  }
  __syncthreads();
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
  if (t < 6) {
    const double* S = Vinv + (size_t)line * kVinvStride;
    double u = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) u += S[sym_idx(t < b ? t : b, t < b ? b : t)] * s_t[b];
    u = -u;
    const double x = pts2.sel(cur)[(size_t)line * 6 + t];
    pts2.sel(cur ^ 1)[(size_t)line * 6 + t] = __dadd_rn(x, __dmul_rn(alpha, u));
    dl[(size_t)line * 6 + t] = u;
    s_u[t] = u;
    p2 = u * u; p3 = x * x;
  }
  __syncthreads();
  for (int f = 0; f < f1 - f0; ++f) {
    const int tile = tile0 + f;
    if (t >= tidx[tile].meta.nobs) continue;
    double d = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) d += jl[((size_t)tile * 6 + k) * kTile + t] * s_u[k];
    const size_t i = (size_t)tile * kTile + t;
    const double jd = s_obs[i] + d;
    p0 += r[i] * jd;
    p1 += jd * jd;
  }
  const double sums[4] = {block_sum<kTile>(p0, red), block_sum<kTile>(p1, red), block_sum<kTile>(p2, red), block_sum<kTile>(p3, red)};
  // like last_cta_reduce, but ADDED to what the tile kernel already wrote
  __shared__ int s_last;
  if (t == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) partial[(size_t)q * 4 + c] = sums[c];
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (t < 4) {
    double acc = out4[t];
    for (int g = 0; g < (int)gridDim.x; ++g) acc += __ldcg(partial + (size_t)g * 4 + t);
    out4[t] = acc;
  }
  if (t == 0) *ticket = 0u;
}

// Camera-ordered half of S*p: qpart_c = sum_{j in c} jc_j w_{perm[j]}
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_pcg_cam(int nc, const int* __restrict__ cam_ptr,
                                                          const int* __restrict__ cam_perm,
                                                          const Dual<const double> jc2, const double* __restrict__ w,
                                                          double* __restrict__ qpart, const CgState* __restrict__ st,
                                                          const PeerLink pl, const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  __shared__ double sm[THREADS / 32][6];
  if (st->done) return;
  const double* __restrict__ jc = jc2.sel(ctl_cur(ctl));
  for (int c = blockIdx.x; c < nc; c += gridDim.x) {
    const int j0 = cam_ptr[c], j1 = cam_ptr[c + 1];
    double acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int j = j0 + threadIdx.x; j < j1; j += THREADS) {
      const double wi = __ldg(w + cam_perm[j]);
      double x[6];
      load6(jc + (size_t)j * 6, x);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[k] += wi * x[k];
    }
    const int wp = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double s = warp_sum(acc[k]);
      if (l == 0) sm[wp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < THREADS / 32; ++q) s += sm[q][threadIdx.x];
      qpart[(size_t)c * 6 + threadIdx.x] = s;
    }
    __syncthreads();
  }
  // multi-GPU, peer transport: qpart IS this rank's slot of the symmetric buffer; the last CTA tells
  // every rank that it is complete (the consumer, k_cg_step, adds the slots of all ranks itself)
  if (pl.seq) peer_signal_last_cta(pl);
}

struct GridBar;
struct CgArgs {
  int nc;
  double lam, tol2;
  int max_iter;
  const double* Dc;     // [nc][6]
  const double* Minv;   // [nc][21]
  Dual<const double> Gm;     // [nc][9]   (both evaluation sets)
  Dual<const double> cams;   // [nc][6]
  const LmCtl* ctl;     // LM loop: current set and damping come from the control block
  const double* b;      // [nc][6]
  const double* qpart;  // [nc][6] (all-reduced partial S p without the damping term)
  double* x; double* r; double* z; double* p; double* tab;
  CgState* st;
  double* partial;   // [2][kCgMaxCtas] block sums of the grid-wide reductions
  GridBar* gb;
  PeerLink pl;       // seq != 0: S p is the sum over ranks of the peer slots (fused all-reduce)
};

// The CG vector work runs on a small co-resident grid (<= one CTA per SM): every camera owns 8
// consecutive lanes (6 active); the two dot products of a step are grid-wide sums done with a
// software grid barrier -- each CTA publishes its block sum, arrives on a global counter, spins until
// all CTAs have arrived and then adds the partials in CTA order, so every CTA holds the same bits.
constexpr int kCgBlock = 512;
constexpr int kCgMaxCtas = 128;

struct GridBar {
  unsigned long long count;   // monotonically increasing arrival counter
  unsigned long long base;    // value of `count` when the current launch started
};

__device__ __forceinline__ double grid_sum(double v, double* red, double* partial /*[gridDim]*/, GridBar* gb,
                                           unsigned long long target) {
  const double mine = block_sum<kCgBlock>(v, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = mine;
    __threadfence();
    atomicAdd(&gb->count, 1ULL);
    SpinGuard sg;
    while (*reinterpret_cast<volatile unsigned long long*>(&gb->count) < target) { if (sg.expired(kSpinGridBar)) break; }
    __threadfence();
  }
  __syncthreads();
  // one warp adds the published block sums (lane-strided, then the butterfly: the same order in every CTA)
  if (threadIdx.x < 32) {
    double tot = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) tot += __ldcg(partial + b);
    tot = warp_sum(tot);
    if (threadIdx.x == 0) red[0] = tot;
  }
  __syncthreads();
  const double out = red[0];
  __syncthreads();
  return out;
}

// z_k = sum_j Minv[k][j] r_j inside an 8-lane camera group (lane k holds r_k)
__device__ __forceinline__ double group_minv(const double* __restrict__ Minv_c, double rk, int k, bool act) {
  double z = 0.0;
  const int base = (threadIdx.x & 31) & 24;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double rj = __shfl_sync(0xffffffffu, rk, base + j);
    if (act) z += Minv_c[sym_idx(k < j ? k : j, k < j ? j : k)] * rj;
  }
  return z;
}

// lane k of a camera group writes tab[k] for the vector whose component k it holds
__device__ __forceinline__ void group_table(const double* __restrict__ G, const double* __restrict__ cam, double vk, int k,
                                            bool act, double* __restrict__ tab_c) {
  const int base = (threadIdx.x & 31) & 24;
  double v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) v[j] = __shfl_sync(0xffffffffu, vk, base + j);
  if (!act) return;
  const double a0 = G[0] * v[0] + G[1] * v[1] + G[2] * v[2];
  const double a1 = G[3] * v[0] + G[4] * v[1] + G[5] * v[2];
  const double a2 = G[6] * v[0] + G[7] * v[1] + G[8] * v[2];
  const double c0 = cam[3], c1 = cam[4], c2 = cam[5];
  double out;
  switch (k) {
    case 0: out = a0; break;
    case 1: out = a1; break;
    case 2: out = a2; break;
    case 3: out = a1 * c2 - a2 * c1 + v[3]; break;
    case 4: out = a2 * c0 - a0 * c2 + v[4]; break;
    default: out = a0 * c1 - a1 * c0 + v[5]; break;
  }
  tab_c[k] = out;
}

// x = 0, r = b, z = Minv r, p = z, rz0 = r.z ; table(p)
__global__ void __launch_bounds__(kCgBlock) k_cg_begin(const CgArgs a) {
  pdl_prologue();
  __shared__ double red[kCgBlock / 32];
  if (ctl_skip_solve(a.ctl)) return;
  const double* Gm = a.Gm.sel(ctl_cur(a.ctl));
  const double* cams = a.cams.sel(ctl_cur(a.ctl));
  const unsigned long long base = a.gb->base;
  const int gt = blockIdx.x * kCgBlock + threadIdx.x;
  const int k = gt & 7;
  const int groups = gridDim.x * kCgBlock / 8;
  double acc = 0.0;
  for (int c0 = 0; c0 < a.nc; c0 += groups) {
    const int c = c0 + (gt >> 3);
    const bool act = c < a.nc && k < 6;
    const size_t e = (size_t)c * 6 + k;
    const double rk = act ? a.b[e] : 0.0;
    const double zk = group_minv(a.Minv + (size_t)(act ? c : 0) * 21, rk, k, act);
    if (act) {
      a.x[e] = 0.0; a.r[e] = rk; a.z[e] = zk; a.p[e] = zk;
      acc += rk * zk;
    }
    group_table(Gm + (size_t)(act ? c : 0) * 9, cams + (size_t)(act ? c : 0) * 6, zk, k, act, a.tab + (size_t)(act ? c : 0) * 6);
  }
  const double rz = grid_sum(acc, red, a.partial, a.gb, base + gridDim.x);
  if (gt == 0) {
    a.st->rz = rz; a.st->rz0 = rz; a.st->pq = 0.0; a.st->k = 0;
    a.st->breakdown = 0;
    a.st->done = !(rz > 0.0);
    a.gb->base = base + gridDim.x;
  }
}

// One CG step after q = S p is available (q = qpart + lam Dc p).  When the iteration stops, the
// gather table of the SOLUTION x is written instead of that of the next direction, ready for the
// back-substitution pass.
__global__ void __launch_bounds__(kCgBlock) k_cg_step(const CgArgs a) {
  pdl_prologue();
  __shared__ double red[kCgBlock / 32];
  if (a.st->done) return;   // uniform over the whole grid (written by the previous launch)
  const double* Gm = a.Gm.sel(ctl_cur(a.ctl));
  const double* cams = a.cams.sel(ctl_cur(a.ctl));
  const double lam = a.ctl ? a.ctl->lam : a.lam;
  const int max_iter = a.ctl ? a.ctl->pcg_max_iter : a.max_iter;
  const double rz = a.st->rz, rz0 = a.st->rz0;
  const int k_new = a.st->k + 1;
  const unsigned long long base = a.gb->base;
  const int gt = blockIdx.x * kCgBlock + threadIdx.x;
  const int k = gt & 7;
  const int groups = gridDim.x * kCgBlock / 8;
  const bool fused = a.pl.seq != 0ULL;
  const bool tracing = fused && a.pl.pa.trace != nullptr && gt == 0;
  unsigned long long tr0 = 0;
  if (tracing) tr0 = global_ns();
  if (fused) peer_wait(a.pl);   // every rank's partial S p is complete
  if (tracing) {
    atomicAdd(a.pl.pa.trace + 24, 1ULL);
    atomicAdd(a.pl.pa.trace + 25, global_ns() - tr0);
  }
  double acc = 0.0;
  for (int c0 = 0; c0 < a.nc; c0 += groups) {
    const int c = c0 + (gt >> 3);
    if (c < a.nc && k < 6) {
      const size_t e = (size_t)c * 6 + k;
      const double pp = a.p[e];
      const double qsum = fused ? peer_sum(a.pl, e) : a.qpart[e];
      const double q = qsum + lam * a.Dc[e] * pp;
      a.z[e] = q;  // z is free until the preconditioner is applied below; reuse it to keep q
      acc += pp * q;
    }
  }
  const double pq = grid_sum(acc, red, a.partial, a.gb, base + gridDim.x);
  const bool ok = pq > 0.0;
  const double alpha = ok ? rz / pq : 0.0;
  acc = 0.0;
  for (int c0 = 0; c0 < a.nc; c0 += groups) {
    const int c = c0 + (gt >> 3);
    const bool act = c < a.nc && k < 6;
    const size_t e = (size_t)c * 6 + k;
    double rk = 0.0;
    if (act) {
      a.x[e] += alpha * a.p[e];
      rk = a.r[e] - alpha * a.z[e];
      a.r[e] = rk;
    }
    const double zk = group_minv(a.Minv + (size_t)(act ? c : 0) * 21, rk, k, act);
    if (act) {
      a.z[e] = zk;
      acc += rk * zk;
    }
  }
  const double rz_new = grid_sum(acc, red, a.partial + kCgMaxCtas, a.gb, base + 2ULL * gridDim.x);
  const bool stop = !ok || rz_new <= a.tol2 * rz0 || k_new >= max_iter;
  const double beta = stop ? 0.0 : rz_new / rz;
  for (int c0 = 0; c0 < a.nc; c0 += groups) {
    const int c = c0 + (gt >> 3);
    const bool act = c < a.nc && k < 6;
    const size_t e = (size_t)c * 6 + k;
    double vk = 0.0;
    if (act) {
      if (stop) vk = a.x[e];
      else { vk = a.z[e] + beta * a.p[e]; a.p[e] = vk; }
    }
    group_table(Gm + (size_t)(act ? c : 0) * 9, cams + (size_t)(act ? c : 0) * 6, vk, k, act, a.tab + (size_t)(act ? c : 0) * 6);
  }
  // every CTA has read st and gb->base before the second barrier, so thread 0 may publish the new state
  if (gt == 0) {
    a.st->pq = pq;
    a.st->k = k_new;
    a.st->rz = rz_new;
    if (!ok) a.st->breakdown = 1;
    if (stop) a.st->done = 1;
    a.gb->base = base + 2ULL * gridDim.x;
  }
}

// ---- the same two kernels for problems of up to 1024 cameras on one GPU: ONE thread-block cluster ------------------
// 8 CTAs x 1024 threads hold one lane per (camera, component), so every vector element lives in a register from its
// load to its store; the two dot products of a step are block sums exchanged through distributed shared memory and
// added in rank order after a hardware cluster barrier (a few hundred ns, against ~3 us for the software grid barrier
// above: store, fence, atomic, poll, fence, reload).  Same formulas; the block sums partition the cameras differently,
// so the scalars differ from the grid kernels' in the last bits.
constexpr int kCgClBlock = 1024;
constexpr int kCgClMaxCtas = 8;   // portable cluster size

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem(const double* local, unsigned rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}
// sum over the cluster, the same bits in every thread of every CTA (`slot`: one double of shared memory per use)
__device__ __forceinline__ double cluster_sum(double v, double* red, double* slot) {
  const double mine = block_sum<kCgClBlock>(v, red);
  if (threadIdx.x == 0) *slot = mine;
  cluster_sync_all();
  const unsigned lane = threadIdx.x & 31;
  const double part = lane < cluster_nctarank() ? ld_dsmem(slot, lane) : 0.0;
  return warp_sum(part);
}

__global__ void __launch_bounds__(kCgClBlock) k_cg_begin_cl(const CgArgs a) {
  pdl_prologue();
  __shared__ double red[kCgClBlock / 32];
  __shared__ double slot[1];
  if (ctl_skip_solve(a.ctl)) return;   // uniform over the cluster
  const double* Gm = a.Gm.sel(ctl_cur(a.ctl));
  const double* cams = a.cams.sel(ctl_cur(a.ctl));
  const int gt = cluster_ctarank() * kCgClBlock + threadIdx.x;
  const int k = gt & 7, c = gt >> 3;
  const bool act = c < a.nc && k < 6;
  const size_t cc = act ? c : 0, e = cc * 6 + (act ? k : 0);
  const double rk = act ? a.b[e] : 0.0;
  const double zk = group_minv(a.Minv + cc * 21, rk, k, act);
  if (act) { a.x[e] = 0.0; a.r[e] = rk; a.z[e] = zk; a.p[e] = zk; }
  group_table(Gm + cc * 9, cams + cc * 6, zk, k, act, a.tab + cc * 6);
  const double rz = cluster_sum(act ? rk * zk : 0.0, red, &slot[0]);
  if (gt == 0) {
    a.st->rz = rz; a.st->rz0 = rz; a.st->pq = 0.0; a.st->k = 0;
    a.st->breakdown = 0;
    a.st->done = !(rz > 0.0);
  }
  cluster_sync_all();   // no CTA leaves while its slot may still be read
}

__global__ void __launch_bounds__(kCgClBlock) k_cg_step_cl(const CgArgs a) {
  pdl_prologue();
  __shared__ double red[kCgClBlock / 32];
  __shared__ double slot[2];
  if (a.st->done) return;   // uniform over the cluster (written by the previous launch)
  const double* Gm = a.Gm.sel(ctl_cur(a.ctl));
  const double* cams = a.cams.sel(ctl_cur(a.ctl));
  const double lam = a.ctl ? a.ctl->lam : a.lam;
  const int max_iter = a.ctl ? a.ctl->pcg_max_iter : a.max_iter;
  const double rz = a.st->rz, rz0 = a.st->rz0;
  const int k_new = a.st->k + 1;
  const int gt = cluster_ctarank() * kCgClBlock + threadIdx.x;
  const int k = gt & 7, c = gt >> 3;
  const bool act = c < a.nc && k < 6;
  const size_t cc = act ? c : 0, e = cc * 6 + (act ? k : 0);
  // everything the lane needs, loaded once
  const double pp = act ? a.p[e] : 0.0;
  const double q = act ? a.qpart[e] + lam * a.Dc[e] * pp : 0.0;
  double xk = act ? a.x[e] : 0.0;
  double rk = act ? a.r[e] : 0.0;
  double mrow[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) mrow[j] = act ? a.Minv[cc * 21 + sym_idx(k < j ? k : j, k < j ? j : k)] : 0.0;
  const double pq = cluster_sum(pp * q, red, &slot[0]);
  const bool ok = pq > 0.0;
  const double alpha = ok ? rz / pq : 0.0;
  xk += alpha * pp;
  rk = rk - alpha * q;
  double zk = 0.0;
  {
    const int base = (threadIdx.x & 31) & 24;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double rj = __shfl_sync(0xffffffffu, rk, base + j);
      zk += mrow[j] * rj;
    }
  }
  const double rz_new = cluster_sum(act ? rk * zk : 0.0, red, &slot[1]);
  const bool stop = !ok || rz_new <= a.tol2 * rz0 || k_new >= max_iter;
  const double beta = stop ? 0.0 : rz_new / rz;
  const double vk = stop ? xk : zk + beta * pp;
  if (act) {
    a.x[e] = xk; a.r[e] = rk; a.z[e] = zk;
    if (!stop) a.p[e] = vk;
  }
  group_table(Gm + cc * 9, cams + cc * 6, vk, k, act, a.tab + cc * 6);
  cluster_sync_all();   // every CTA has read the CG state; no CTA leaves while its slots may still be read
  if (gt == 0) {
    a.st->pq = pq;
    a.st->k = k_new;
    a.st->rz = rz_new;
    if (!ok) a.st->breakdown = 1;
    if (stop) a.st->done = 1;
  }
}

// ---------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------
struct Scalars {            // device-resident LM scalars, copied to the host once per iteration
  double cost;              // 0.5 sum r^2 of the last evaluation
  double rjd, jd2;          // sum r (J delta), sum (J delta)^2
  double dl2, x2_lines;     // |delta_lines|^2, |pts|^2
  double dc2, x2_cams;      // |delta_cams|^2, |cams|^2
  double gnorm;             // inf-norm of the gradient
  int pcg_iters, pcg_breakdown;
  int clamped, pad;
};

// What the host reads of a running solve, in pinned host memory the device writes directly: one tiny kernel
// posts the values and then a sequence number the host spins on (no copy-engine round trips).
struct HostMirror {
  Scalars scal;
  CgState cg;
  LmCtl ctl;
  unsigned long long gbits;
  int clamped;
  unsigned int spin_error;   // g_spin_error: a bounded device-side wait gave up (LBA_ERR_COMM)
  unsigned long long seq;
};
__global__ void k_publish(const Scalars* __restrict__ s, const CgState* __restrict__ cg, const LmCtl* __restrict__ ctl,
                          const Dual<const unsigned long long> gbits, const int* __restrict__ clamped,
                          HostMirror* __restrict__ h, unsigned long long seq) {
  pdl_prologue();
  if (threadIdx.x == 0) {
    h->scal = *s;
    h->cg = *cg;
    h->ctl = *ctl;
    h->gbits = *gbits.sel(ctl->cur);
    h->clamped = *clamped;
    h->spin_error = g_spin_error;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&h->seq) = seq;
  }
}

// Start of an LM solve: options and the initial state of the control block (the initial evaluation is in set 0
// and its cost, all-reduced, in scal->cost).
__global__ void k_lm_init(LmCtl* __restrict__ L, const LmCtl opts, const Scalars* __restrict__ s) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  LmCtl c = opts;
  c.cost = s->cost;
  c.cur = 0; c.need_solve = 1; c.done = 0; c.status = 0; c.it = 0; c.evals = 1; c.streak = 0; c.bt = 0;
  c.n_acc = 0; c.pcg_total = 0; c.cg_short = 0; c.trace_len = 0; c.scaling_mode = 2; c.q_last = 0; c.pass = 0;
  c.nu = 2.0; c.sigma = 1.0; c.alpha = 1.0; c.g_norm = 0.0;
  if (!pol_finite(c.cost)) { c.done = 1; c.status = -5; }   // "Residuals are not finite in the initial point."
  if (c.max_iter <= 0) c.done = 1;
  *L = c;
}

// The decision after a trial evaluation (lba_policy.cuh): one thread.
// It also posts what the host wants to see of the pass (h != NULL): one launch less per trial than k_publish behind it.
__global__ void k_lm_decide(LmCtl* __restrict__ L, const Scalars* __restrict__ s, const CgState* __restrict__ cg,
                            const Dual<const unsigned long long> gbits, LmTrace* __restrict__ tr,
                            const int* __restrict__ clamped, HostMirror* __restrict__ h, unsigned long long seq) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  LmCtl c = *L;
  LmTrial t;
  t.cost_new = s->cost; t.rjd = s->rjd; t.jd2 = s->jd2;
  t.d2 = __dadd_rn(s->dc2, s->dl2);
  t.x2 = __dadd_rn(s->x2_cams, s->x2_lines);
  t.g_norm = __longlong_as_double((long long)*gbits.sel(c.cur));
  t.pcg_iters = cg->k; t.pcg_done = cg->done;
  lm_decide(c, t, tr);
  *L = c;
  if (h) {
    h->scal = *s;
    h->cg = *cg;
    h->ctl = c;
    h->gbits = *gbits.sel(c.cur);
    h->clamped = *clamped;
    h->spin_error = g_spin_error;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&h->seq) = seq;
  }
}

// Trial point on the camera side: cams_new = cams + alpha dc, its camera records and SO(3) Jacobians, the step /
// parameter norms (grid-wide sums by the last CTA, fixed order), and the reset of the trial set's gradient-norm
// word.  In a RETRY pass (same direction, shorter alpha) the CTAs also rebuild the line side of the trial point
// from the kept step: pts_new = pts + alpha dl.
constexpr int kCamStepThreads = 128;
__global__ void __launch_bounds__(kCamStepThreads) k_cam_step(int nc, const Dual<double> cams2, const double* __restrict__ dc,
                                                               const Dual<CamRec> recs2, const Dual<double> Gm2,
                                                               double alpha, const Dual<unsigned long long> gbits2,
                                                               size_t n_line_params, const Dual<double> pts2,
                                                               const double* __restrict__ dl,
                                                               double* __restrict__ partial, unsigned int* ticket,
                                                               double* __restrict__ out2 /* dc2, x2_cams */,
                                                               const LmCtl* __restrict__ ctl) {
  pdl_prologue();
  __shared__ double red[kCamStepThreads / 32];
  if (ctl_skip_trial(ctl)) return;
  const int cur = ctl_cur(ctl);
  if (ctl) alpha = ctl->alpha;
  const double* __restrict__ cams = cams2.sel(cur);
  double* __restrict__ cams_new = cams2.sel(cur ^ 1);
  CamRec* __restrict__ recs_new = recs2.sel(cur ^ 1);
  double* __restrict__ Gm_new = Gm2.sel(cur ^ 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) *gbits2.sel(cur ^ 1) = 0ULL;
  if (ctl && !ctl->need_solve) {
    const double* __restrict__ pts = pts2.sel(cur);
    double* __restrict__ pts_new = pts2.sel(cur ^ 1);
    for (size_t e = (size_t)blockIdx.x * kCamStepThreads + threadIdx.x; e < n_line_params; e += (size_t)gridDim.x * kCamStepThreads)
      pts_new[e] = __dadd_rn(pts[e], __dmul_rn(alpha, dl[e]));
  }
  double d2 = 0.0, x2 = 0.0;
  const int c = blockIdx.x * kCamStepThreads + threadIdx.x;
  if (c < nc) {
    CamRec rec;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double x = cams[c * 6 + k], d = dc[c * 6 + k];
      const double xn = __dadd_rn(x, __dmul_rn(alpha, d));
      cams_new[c * 6 + k] = xn;
      if (k < 3) rec.w[k] = xn; else rec.c[k - 3] = xn;
      d2 += d * d;
      x2 += x * x;
    }
    rodrigues_coeffs(rec.w, rec.A, rec.B, rec.C);
    rec.pad = 0.0;
    recs_new[c] = rec;
    double G[9];
    build_G(rec.w, rec.B, rec.C, G);
#pragma unroll
    for (int k = 0; k < 9; ++k) Gm_new[c * 9 + k] = G[k];
  }
  const double sums[2] = {block_sum<kCamStepThreads>(d2, red), block_sum<kCamStepThreads>(x2, red)};
  last_cta_reduce<kCamStepThreads, 2>(sums, partial, ticket, out2, 1.0, red);
}

__global__ void k_sub(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  pdl_prologue();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = a[t] - b[t];
}

// slot order -> caller's observation order
__global__ void k_slots_to_caller(int Mp, const int* __restrict__ slot_src, const double* __restrict__ src,
                                  double* __restrict__ dst) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const int o = slot_src[i];
  if (o >= 0) dst[o] = src[i];
}

// Jacobian rows in API layout: jcam (m,6) interleaved [w0,c0,w1,c1,w2,c2], jline (m,6); caller order.
__global__ void k_export_jacobian(int Mp, const int* __restrict__ slot_src, const TileIdx* __restrict__ tidx,
                                  const double* __restrict__ jl, const double* __restrict__ jc,
                                  double* __restrict__ jcam, double* __restrict__ jline) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  const int src = slot_src[i];
  if (src < 0) return;
  const size_t o = (size_t)src;
  const size_t pos = (size_t)tidx[i / kTile].pos[i % kTile];
  for (int k = 0; k < 3; ++k) {
    jcam[o * 6 + 2 * k] = jc[pos * 6 + k];
    jcam[o * 6 + 2 * k + 1] = jc[pos * 6 + 3 + k];
  }
  for (int k = 0; k < 6; ++k) jline[o * 6 + k] = jl[tm_index(6, k, (size_t)i)];
}

// ---------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory (replaces ncclAllReduce for the small, latency-bound
// per-camera vectors).  Every rank owns a "symmetric" buffer mapped into all peers (CUDA IPC):
//   flags[region][2][kMaxPeers] (sequence numbers) | per region: data[2][cap] doubles (ping-pong)
// The PRODUCER kernel of a partial sum (k_pcg_cam, k_cam_grad, k_cam_accum) writes straight into the
// own slot (zero copy).  The reduce kernel of call #seq, slot seq&1: (1) one CTA stores seq into
// flag[region][slot][my_rank] of EVERY rank (release, system scope) -- the partial was written by the
// previous kernel on the stream, so it is complete; (2) every CTA waits until the own flags of all
// ranks reach seq (acquire); (3) the W slots are added in rank order -- the same order on every rank,
// so all ranks end with identical bits.  A slot is rewritten two calls later, after every peer has
// signalled the call in between, i.e. has finished reading it.
// ---------------------------------------------------------------------------------
// slot_off: offset in doubles of (region, slot) inside the data area; n must be even (padded by the host)
// src != nullptr: single-CTA mode that first copies n doubles from local memory into the own slot.
template <bool MAX_U64>
__global__ void __launch_bounds__(512) k_peer_reduce(const PeerArgs pa, int region, size_t slot_off, size_t n,
                                                      unsigned long long seq, const double* __restrict__ src,
                                                      double* __restrict__ dst,
                                                      unsigned long long* maxp = nullptr,
                                                      unsigned long long* maxp1 = nullptr,
                                                      const LmCtl* __restrict__ ctl = nullptr) {
  pdl_prologue();
  if (maxp && ctl && ctl->cur) maxp = maxp1;   // LM loop: the gradient-norm word of the current evaluation set
  const int slot = (int)(seq & 1ULL);
  const int t = threadIdx.x;
  constexpr int kMaxAt = 8;   // maxp (single-CTA mode, n <= 8): one more word, max-reduced as u64, rides at this index
  if (src) {
    double* mine = reinterpret_cast<double*>(pa.peer[pa.rank] + kPeerDataOff) + slot_off;
    for (size_t i = t; i < n; i += blockDim.x) mine[i] = src[i];
    if (maxp && t == 0) mine[kMaxAt] = __longlong_as_double((long long)*maxp);
    __syncthreads();
  }
  const int fidx = (region * 2 + slot) * kMaxPeers;
  const bool tracing = pa.trace != nullptr && blockIdx.x == 0 && t == 0;
  unsigned long long tr0 = 0, tr1 = 0;
  if (tracing) tr0 = global_ns();
  if (blockIdx.x == 0 && t < pa.world)
    st_release_sys(reinterpret_cast<unsigned long long*>(pa.peer[t]) + fidx + pa.rank, seq);
  if (t < pa.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(pa.peer[pa.rank]) + fidx + t;
    SpinGuard sg;
    while (ld_acquire_sys(f) < seq) { if (sg.expired(kSpinPeerReduce)) break; }
  }
  __syncthreads();
  if (tracing) tr1 = global_ns();
  const size_t gsz = (size_t)gridDim.x * blockDim.x;
  for (size_t i2 = (size_t)blockIdx.x * blockDim.x + t; 2 * i2 < n; i2 += gsz) {
    double2 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < pa.world) v[r] = ld_relaxed_sys2(reinterpret_cast<const double*>(pa.peer[r] + kPeerDataOff) + slot_off + 2 * i2);
    double2 acc;
    if (MAX_U64) {
      unsigned long long m0 = 0ULL, m1 = 0ULL;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pa.world) {
          const unsigned long long a0 = (unsigned long long)__double_as_longlong(v[r].x);
          const unsigned long long a1 = (unsigned long long)__double_as_longlong(v[r].y);
          m0 = a0 > m0 ? a0 : m0; m1 = a1 > m1 ? a1 : m1;
        }
      acc.x = __longlong_as_double((long long)m0); acc.y = __longlong_as_double((long long)m1);
    } else {
      acc.x = 0.0; acc.y = 0.0;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pa.world) { acc.x += v[r].x; acc.y += v[r].y; }
    }
    if (2 * i2 + 1 < n) *reinterpret_cast<double2*>(dst + 2 * i2) = acc;
    else dst[2 * i2] = acc.x;
  }
  if (maxp && blockIdx.x == 0 && t == 0) {
    unsigned long long m = 0ULL;
    for (int r = 0; r < pa.world; ++r) {
      const double2 v = ld_relaxed_sys2(reinterpret_cast<const double*>(pa.peer[r] + kPeerDataOff) + slot_off + kMaxAt);
      const unsigned long long a = (unsigned long long)__double_as_longlong(v.x);
      m = a > m ? a : m;
    }
    *maxp = m;
  }
  if (tracing) {
    atomicAdd(pa.trace + region * 4 + 0, 1ULL);
    atomicAdd(pa.trace + region * 4 + 1, tr1 - tr0);
    atomicAdd(pa.trace + region * 4 + 2, global_ns() - tr0);
  }
}

// Two-phase variant for the large per-camera exchanges at >= 4 ranks (reduce-scatter + all-gather over peer
// memory): rank r sums only its 1/W slice of the elements (reading W-1 remote slices instead of W-1 whole
// slots), stores the sums into the result slot of EVERY rank (remote stores are posted), signals, and each
// rank then copies its now complete local result slot to `dst`.  Per rank (W-1)/W * n doubles cross NVLink in
// each direction instead of (W-1) * n inbound.  One rank adds each element, in rank order: identical bits
// everywhere.  The grid (<= 32 CTAs) is co-resident; `ticket` counts the CTAs that finished phase one.
__global__ void __launch_bounds__(512) k_peer_reduce_rs(const PeerArgs pa, int region, size_t slot_off, size_t res_off,
                                                         size_t n, unsigned long long seq, unsigned long long rseq,
                                                         double* __restrict__ dst, unsigned int* __restrict__ ticket) {
  pdl_prologue();
  __shared__ int s_last;
  const int t = threadIdx.x;
  // the result region is shared by the exchanges that use this kernel, so it has its own sequence (rseq)
  const int fidx = (region * 2 + (int)(seq & 1ULL)) * kMaxPeers, ridx = (kRgnR * 2 + (int)(rseq & 1ULL)) * kMaxPeers;
  const bool tracing = pa.trace != nullptr && blockIdx.x == 0 && t == 0;
  unsigned long long tr0 = 0, tr1 = 0;
  if (tracing) tr0 = global_ns();
  if (blockIdx.x == 0 && t < pa.world)
    st_release_sys(reinterpret_cast<unsigned long long*>(pa.peer[t]) + fidx + pa.rank, seq);
  if (t < pa.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(pa.peer[pa.rank]) + fidx + t;
    SpinGuard sg;
    while (ld_acquire_sys(f) < seq) { if (sg.expired(kSpinPeerReduceRs)) break; }
  }
  __syncthreads();
  if (tracing) tr1 = global_ns();
  // phase one: my slice, summed over the ranks' slots, stored to every rank's result slot
  const size_t n2 = (n + 1) / 2, chunk = (n2 + pa.world - 1) / pa.world;
  const size_t lo = (size_t)pa.rank * chunk, hi = lo + chunk < n2 ? lo + chunk : n2;
  const size_t gsz = (size_t)gridDim.x * blockDim.x;
  for (size_t i2 = lo + (size_t)blockIdx.x * blockDim.x + t; i2 < hi; i2 += gsz) {
    double2 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < pa.world) v[r] = ld_relaxed_sys2(reinterpret_cast<const double*>(pa.peer[r] + kPeerDataOff) + slot_off + 2 * i2);
    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < pa.world) { acc.x += v[r].x; acc.y += v[r].y; }
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < pa.world) *reinterpret_cast<double2*>(reinterpret_cast<double*>(pa.peer[r] + kPeerDataOff) + res_off + 2 * i2) = acc;
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (t == 0) *ticket = 0u;
    if (t < pa.world) st_release_sys(reinterpret_cast<unsigned long long*>(pa.peer[t]) + ridx + pa.rank, rseq);
  }
  // phase two: every rank's slice of the result has landed in my result slot
  if (t < pa.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(pa.peer[pa.rank]) + ridx + t;
    SpinGuard sg2;
    while (ld_acquire_sys(f) < rseq) { if (sg2.expired(kSpinPeerReduceRs)) break; }
  }
  __syncthreads();
  const double* res = reinterpret_cast<const double*>(pa.peer[pa.rank] + kPeerDataOff) + res_off;
  for (size_t i2 = (size_t)blockIdx.x * blockDim.x + t; i2 < n2; i2 += gsz) {
    const double2 v = ld_relaxed_sys2(res + 2 * i2);
    if (2 * i2 + 1 < n) *reinterpret_cast<double2*>(dst + 2 * i2) = v;
    else dst[2 * i2] = v.x;
  }
  if (tracing) {
    atomicAdd(pa.trace + region * 4 + 0, 1ULL);
    atomicAdd(pa.trace + region * 4 + 1, tr1 - tr0);
    atomicAdd(pa.trace + region * 4 + 2, global_ns() - tr0);
  }
}

// ---------------------------------------------------------------------------------
// N2: track table (tracker.py:11-96) one view pair at a time, data-parallel inside a pair.
//   ids / bin: [cap][V] int64, zero-filled; st[0] = rows so far, st[1] = inliers of the pair, st[2] = new tracks
//   of the pair.  Per pair v: verify -> exclusive scans (CUB) -> presence table over column v -> last-writer
//   table over the inlier matches -> extension of the existing rows -> ordered append of the new tracks.
// ---------------------------------------------------------------------------------
// flag[j] = 1 when match j of the pair passes verify_duplet; endpoints [n_lines2d][2][2]; line ids range-checked
__global__ void k_tt_verify(long long k, long long n_lines, const long long* __restrict__ matches,
                            const double* __restrict__ endpoints, const double* __restrict__ F, double threshold,
                            const unsigned char* __restrict__ flags_in, int* __restrict__ flag, int* __restrict__ bad) {
  pdl_prologue();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const long long l0 = matches[2 * j], l1 = matches[2 * j + 1];
  if (l0 < 0 || l0 >= n_lines || l1 < 0 || l1 >= n_lines) { atomicAdd(bad, 1); flag[j] = 0; return; }
  if (flags_in) { flag[j] = flags_in[j] ? 1 : 0; return; }
  double Fm[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) Fm[q] = __ldg(F + q);
  const double* a = endpoints + l0 * 4;
  const double* b = endpoints + l1 * 4;
  const double a0[2] = {a[0], a[1]}, a1[2] = {a[2], a[3]}, b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
  flag[j] = duplet_inlier(Fm, a0, a1, b0, b1, threshold) ? 1 : 0;
}
// present[x] = 1 for every value x in column v of the existing rows (the zeros of absent entries included, as in
// the reference's `l in Tr_ij[:, v]`)
__global__ void k_tt_present(const long long* __restrict__ st, int V, int v, const long long* __restrict__ ids,
                             unsigned char* __restrict__ present) {
  pdl_prologue();
  const long long rows = st[0];
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
    present[ids[r * V + v]] = 1;
}
// inlier matches: inl_l1[pos] = second line; last[l0] = largest inlier rank among the matches of a present line
// (the reference applies them in order, so the last one wins); newflag = inlier && line not present
__global__ void k_tt_last(long long k, int v, const long long* __restrict__ matches, const int* __restrict__ flag,
                          const int* __restrict__ pos, const unsigned char* __restrict__ present,
                          long long* __restrict__ inl_l1, int* __restrict__ last, int* __restrict__ newflag) {
  pdl_prologue();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  int nf = 0;
  if (flag[j]) {
    const long long l0 = matches[2 * j];
    inl_l1[pos[j]] = matches[2 * j + 1];
    if (v > 0 && present[l0]) atomicMax(last + l0, pos[j]);
    else nf = 1;
  }
  newflag[j] = nf;
}
__global__ void k_tt_extend(const long long* __restrict__ st, int V, int v, long long* __restrict__ ids,
                            long long* __restrict__ bin, const int* __restrict__ last,
                            const long long* __restrict__ inl_l1) {
  pdl_prologue();
  const long long rows = st[0];
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const int j = last[ids[r * V + v]];
    if (j >= 0) { ids[r * V + v + 1] = inl_l1[j]; bin[r * V + v + 1] = 1; }
  }
}
__global__ void k_tt_append(long long k, int V, int v, const long long* __restrict__ st, const long long* __restrict__ matches,
                            const int* __restrict__ newflag, const int* __restrict__ newpos, long long* __restrict__ ids,
                            long long* __restrict__ bin) {
  pdl_prologue();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k || !newflag[j]) return;
  const long long r = st[0] + newpos[j];
  ids[r * V + v] = matches[2 * j]; ids[r * V + v + 1] = matches[2 * j + 1];
  bin[r * V + v] = 1; bin[r * V + v + 1] = 1;
}
// rows += new tracks of the pair; n_new_tracks[v]; totals come from the last element of the scans
__global__ void k_tt_advance(long long k, int v, long long* __restrict__ st, const int* __restrict__ newflag,
                             const int* __restrict__ newpos, long long* __restrict__ n_new_tracks) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long n_new = k > 0 ? (long long)newpos[k - 1] + newflag[k - 1] : 0;
    n_new_tracks[v] = n_new;
    st[0] += n_new;
  }
}

__global__ void k_tt_flags_u8(long long n, const int* __restrict__ flag, unsigned char* __restrict__ out) {
  pdl_prologue();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = flag[j] ? 1 : 0;
}

// ---------------------------------------------------------------------------------
// N3: two-view endpoint triangulation, one thread per (track, endpoint)  (seq_utils.py:105-127)
// ---------------------------------------------------------------------------------
// P [n_views][12] row-major 3x4; view0/view1 [n] (view1 NULL: view0 + 1, the reference's consecutive pairs);
// pts0/pts1 [n][2][2] float32 = [track][endpoint][x,y]; X [n][2][3] float32; Xh [n][2][4] float32 or NULL.
// FP64-bound (about 2.5 kflop of Jacobi rotations per point against 28 B of traffic); the camera matrices
// (96 B each) are served by L1/L2.  bad[0] counts tracks whose view index is out of range (their X is NaN).
__global__ void __launch_bounds__(256) k_triangulate(long long n2, int n_views, const double* __restrict__ P,
                                                      const long long* __restrict__ view0,
                                                      const long long* __restrict__ view1,
                                                      const float2* __restrict__ pts0, const float2* __restrict__ pts1,
                                                      float* __restrict__ X, float4* __restrict__ Xh,
                                                      int* __restrict__ bad) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 2 * track + endpoint
  if (i >= n2) return;
  const long long t = i >> 1;
  const long long a = view0[t], b = view1 ? view1[t] : a + 1;
  float xh[4], x[3];
  if (a < 0 || a >= n_views || b < 0 || b >= n_views) {
    if ((i & 1) == 0) atomicAdd(bad, 1);
    const float nanv = __int_as_float(0x7fc00000);
    xh[0] = xh[1] = xh[2] = xh[3] = x[0] = x[1] = x[2] = nanv;
  } else {
    double P0[12], P1[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { P0[k] = __ldg(P + a * 12 + k); P1[k] = __ldg(P + b * 12 + k); }
    const float2 q0 = pts0[i], q1 = pts1[i];
    triangulate_dlt(P0, P1, q0.x, q0.y, q1.x, q1.y, xh, x);
  }
  X[i * 3 + 0] = x[0]; X[i * 3 + 1] = x[1]; X[i * 3 + 2] = x[2];
  if (Xh) Xh[i] = make_float4(xh[0], xh[1], xh[2], xh[3]);
}

// ---------------------------------------------------------------------------------
// N1: BA input kept on the device across the incremental loop (seq_utils.py:138-208, main.py:111-133).
// The per-2D-line constants (n, p0, p1 of domain.spherical_representation, domain.py:338-351) and pixel endpoints are
// uploaded ONCE; after every added view the observation list is rebuilt on the device from the track tables and the
// observation constants are gathered from the resident table by 2D line id.
// ---------------------------------------------------------------------------------
// rows of `bin` (T x V int64) seen in more than one view (seq_utils.py:188-189)
__global__ void k_obs_count(long long T, int V, const long long* __restrict__ bin, int* __restrict__ n_tracks) {
  pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  long long s = 0;
  for (int v = 0; v < V; ++v) s += bin[t * V + v];
  if (s > 1) atomicAdd(n_tracks, 1);
}
// the reference takes the FIRST n_tracks rows (bin[:len(cond1)], :190), row-major (track, view) order
__global__ void k_obs_flags(long long T, int V, const long long* __restrict__ bin, const int* __restrict__ n_tracks,
                            int* __restrict__ flag) {
  pdl_prologue();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T * V) return;
  flag[e] = (e / V < *n_tracks && bin[e] == 1) ? 1 : 0;
}
__global__ void k_obs_scatter(long long T, int V, const long long* __restrict__ ids, const int* __restrict__ flag,
                              const int* __restrict__ pos, long long* __restrict__ cam_idx, long long* __restrict__ line_idx,
                              long long* __restrict__ line2d, long long* __restrict__ m_out) {
  pdl_prologue();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T * V) return;
  if (flag[e]) {
    const int p = pos[e];
    cam_idx[p] = e % V; line_idx[p] = e / V; line2d[p] = ids[e];
  }
  if (e == T * V - 1) *m_out = (long long)pos[e] + flag[e];
}
// observation constants of observation i = row line2d[i] of the resident per-2D-line table
__global__ void k_gather_lines2d(long long m, long long n_lines2d, const long long* __restrict__ line2d,
                                 const double* __restrict__ tn, const double* __restrict__ tp0, const double* __restrict__ tp1,
                                 double* __restrict__ n, double* __restrict__ p0, double* __restrict__ p1, int* __restrict__ bad) {
  pdl_prologue();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * m) return;
  const long long i = e / 3, k = e % 3, id = line2d[i];
  if (id < 0 || id >= n_lines2d) { if (k == 0) atomicAdd(bad, 1); n[e] = p0[e] = p1[e] = 0.0; return; }
  n[e] = tn[id * 3 + k]; p0[e] = tp0[id * 3 + k]; p1[e] = tp1[id * 3 + k];
}
// triangulation input of the NEW tracks of view pair (i, i+1), rows [row0, row0 + k) of ids: float32 endpoints of the
// two 2D lines (np.array(..., dtype=np.float32), seq_utils.py:114-117) and the pair's first view
__global__ void k_tri_gather(long long k, long long row0, long long out0, int V, int pair, long long n_lines2d,
                             const long long* __restrict__ ids, const double* __restrict__ endpoints,
                             long long* __restrict__ view0, float2* __restrict__ pts0, float2* __restrict__ pts1,
                             int* __restrict__ bad) {
  pdl_prologue();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const long long l0 = ids[(row0 + j) * V + pair], l1 = ids[(row0 + j) * V + pair + 1];
  const long long o = out0 + j;
  view0[o] = pair;
  const bool ok = l0 >= 0 && l0 < n_lines2d && l1 >= 0 && l1 < n_lines2d;
  if (!ok) atomicAdd(bad, 1);
  for (int e = 0; e < 2; ++e) {
    pts0[o * 2 + e] = ok ? make_float2((float)endpoints[(l0 * 2 + e) * 2], (float)endpoints[(l0 * 2 + e) * 2 + 1]) : make_float2(0.f, 0.f);
    pts1[o * 2 + e] = ok ? make_float2((float)endpoints[(l1 * 2 + e) * 2], (float)endpoints[(l1 * 2 + e) * 2 + 1]) : make_float2(0.f, 0.f);
  }
}
__global__ void k_f32_to_f64(long long n, const float* __restrict__ src, double* __restrict__ dst) {
  pdl_prologue();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) dst[e] = (double)src[e];
}

// ---------------------------------------------------------------------------------
// microbenchmarks for the roofline denominators
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) k_stream_copy(size_t n2, const double2* __restrict__ src,
                                                      double2* __restrict__ dst) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x)
    dst[t] = src[t];
}

}  // namespace lba
