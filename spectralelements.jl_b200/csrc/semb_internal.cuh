// Internal definitions of libsemb (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <functional>
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/semb.h"

#define SEMB_MAXN 17       // largest nr == ns served by the templated strip kernel
// CTA size of the strip kernel: 256 threads (8 warps, 2 per SM sub-partition); 192 for N = 12..14 so that two
// CTAs still fit the 227 KB of shared memory (the staging buffers grow like N^2 per element)
#ifndef SEMB_T9
#define SEMB_T9 256
#endif
constexpr int semb_strip_threads(int n) { return (n >= 12 && n <= 14) ? 192 : (n == 9 ? SEMB_T9 : 256); }
// elements per strip: as many as fit the CTA, with BX*N even so that every staged row is a multiple of 16 bytes
constexpr int semb_strip_bx(int n) {
  return ((semb_strip_threads(n) / n) * n) % 2 == 0 ? semb_strip_threads(n) / n : semb_strip_threads(n) / n - 1;
}
// First element of strip s when Ex elements are dealt to nstrips strips of (nearly) equal width: every CTA of a wave
// then streams the same number of bytes (a fixed width of BX leaves a narrow last strip and BX-wide critical paths).
// For odd N the start is rounded down to an even element so that x0 = e0*N doubles stays 16-byte aligned (bulk copies).
__host__ __device__ inline int semb_strip_e0(int s, int nstrips, int Ex, int N) {
  if (s >= nstrips) return Ex;
  int e = (int)((long long)s * Ex / nstrips);
  if (N & 1) e &= ~1;
  return e;
}
#define SEMB_MAX_RANKS 16

void semb_set_error(const char* fmt, ...);

#define SEMB_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      semb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                              \
      return SEMB_ECUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define SEMB_CHECK_NCCL(expr)                                                         \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != ncclSuccess) {                                                          \
      semb_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(_r)); \
      return SEMB_ENCCL;                                                              \
    }                                                                                 \
  } while (0)

#define SEMB_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      semb_set_error(__VA_ARGS__);   \
      return SEMB_EINVAL;            \
    }                                \
  } while (0)

#define SEMB_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc < 0) return _rc;    \
  } while (0)

// Device-resident PCG / reduction scalars.  One per mesh.
struct SembScal {
  double t;        // sum(r .* h .* mult) of the current residual   (pcg.jl:45)
  double t_prev;   // previous t (the reference recomputes it, pcg.jl:49; same bits)
  double rmax;     // norm(r, Inf)                                    (pcg.jl:36)
  double pap[3];   // partial sum(p .* Ap .* mult): strip kernel, x-seam kernel, y-seam kernel
  double red[4];   // generic reduction results (dot / norm)
  double tol;
  long long iters;
  long long maxiter;
  int done;        // 1 once norm(r,Inf) <= tol or iters == maxiter
  int warned;      // 1 if stopped by maxiter (pcg.jl:39)
  // multi-GPU gathered scalars: slot r = rank r's local contribution
  double xchg_pap[SEMB_MAX_RANKS];
  double xchg_t[2 * SEMB_MAX_RANKS];  // {t_local, rmax_local} per rank
  double xchg_red[2 * SEMB_MAX_RANKS];
  int nranks, rank;
  double pap_total;  // multi-rank: combined sum(p .* Ap .* mult) of the current iteration
  // ---- peer-memory mailbox (multi-GPU, P2P mode): written by the OTHER ranks over NVLink -----------
  // flag_*[r] = epoch of the last value rank r pushed; data slots are double-buffered by epoch parity
  unsigned long long flag_halo[2];  // [0]: row from the lower neighbour, [1]: from the upper neighbour
  unsigned long long flag_pap[SEMB_MAX_RANKS];
  unsigned long long flag_t[SEMB_MAX_RANKS];
  unsigned long long flag_red[SEMB_MAX_RANKS];
  double box_pap[2][SEMB_MAX_RANKS];
  double box_t[2][2 * SEMB_MAX_RANKS];
  double box_red[2][2 * SEMB_MAX_RANKS];
};

// Peer pointers handed to the kernels that exchange data through mapped peer memory (CUDA IPC).
struct P2PArgs {
  int on = 0, nranks = 1, rank = 0;
  unsigned long long epoch = 0;    // epoch of the exchange this kernel takes part in
  unsigned long long epoch_b = 0;  // second exchange in the same kernel (y-seam kernel: halo = epoch, pap = epoch_b)
  SembScal* peer[SEMB_MAX_RANKS] = {nullptr};  // peer[r] = rank r's mailbox (r == rank: the local one)
};

struct semb_ctx {
  std::recursive_mutex mutex;  // one in-flight call per context
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t in_stream = nullptr, out_stream = nullptr;  // H2D / D2H streams of the pipelined host twin
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long launches = 0;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double* flush_buf = nullptr;
  size_t flush_bytes = 0;
  double* d_sync = nullptr;  // 1-double scratch for barrier
  // optional per-kernel timing of the strip kernel (bench.py's roofline leg)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
  size_t prof_used = 0;
};

struct semb_field {
  semb_mesh* mesh = nullptr;
  double* d = nullptr;
};

struct semb_mesh {
  semb_ctx* ctx = nullptr;
  int nr = 0, ns = 0, Ex = 0, Ey = 0;  // Ey global
  int ey0 = 0, ney = 0;                // this rank's slab of element rows
  int perx = 0, pery = 0;
  int nxl = 0, nyl = 0;                // local node counts
  long long pitch = 0;                 // row pitch in doubles (multiple of 16)
  size_t nalloc = 0;                   // doubles per field allocation (pitch * nyl)
  std::vector<double> hDr, hDs, hwr, hws;  // column-major host copies
  double* dDr = nullptr;               // device copies (row-major: D[i*n+k] = D(i,k)) for generic kernels
  double* dDs = nullptr;
  double* arr[SEMB_MESH_ARRAY_COUNT] = {nullptr};
  double* d_wx1d = nullptr;            // separable mult: mult(x,y) = wx1d[x] * wy1d[y] (pitch / nyl entries)
  double* d_wy1d = nullptr;
  bool fast = false;                   // nr == ns in [2, SEMB_MAXN]: templated strip kernel
  bool eo = false;                     // Dr, Ds centro-antisymmetric: even-odd contraction variant
  // launch plan of the strip kernel
  int bx = 0;                          // elements per strip (semb_strip_bx(nr))
  int nstrips = 0, nchunks = 0;
  std::vector<int> h_chunk_r0;         // nchunks+1 element-row offsets
  int* d_chunk_r0 = nullptr;
  unsigned char* d_ystart = nullptr;   // ney+1 flags: element row r starts a y-seam (chunk / periodic / halo)
  std::vector<unsigned char> h_ystart;
  int nxseam = 0;                      // x seams (strip boundaries + periodic wrap): column pairs
  int* d_xseam = nullptr;              // 2*nxseam ints (xa, xb)
  int nyseam = 0;                      // local y seams (row pairs); halo seams handled separately
  int* d_yseam = nullptr;              // 2*nyseam ints (ya, yb)
  int halo_lo = 0, halo_hi = 0;        // 1 if the slab has a neighbour rank below / above
  int rank_lo = -1, rank_hi = -1;
  double* d_halo_lo = nullptr;         // nxl doubles each (received neighbour rows), NCCL path
  double* d_halo_hi = nullptr;
  // P2P path: one IPC-exported allocation per mesh = [SembScal][halo rows: 2 parities x 2 sides x pitch]
  bool p2p = false;
  void* d_mailbox = nullptr;
  double* d_mail_halo = nullptr;       // local halo rows inside the mailbox
  void* peer_mailbox[SEMB_MAX_RANKS] = {nullptr};
  unsigned long long ep_halo = 0, ep_pap = 0, ep_t = 0, ep_red = 0;
  // reductions
  int npartials = 0;
  double* d_partials = nullptr;        // 3 * npartials doubles
  unsigned* d_counters = nullptr;      // 8 tickets
  SembScal* d_scal = nullptr;
  SembScal* h_scal = nullptr;          // pinned mirror
  // pcg state
  semb_field* w_r = nullptr;
  semb_field* w_p = nullptr;
  semb_field* w_Ap = nullptr;
  semb_field* w_h = nullptr;           // h = opM(r) of the preconditioned PCG (written by init / update, staged by the strip kernel)
  semb_field* w_tmp = nullptr;
  semb_field* w_t1 = nullptr;
  semb_field* w_t2 = nullptr;
  semb_field* pcg_x = nullptr;
  semb_pcg_opts pcg_opts;
  bool pcg_active = false;
  bool pcg_keep_h = false;             // preconditioned PCG on the fused path: h kept in w_h
  // operator hook of the device-resident PCG: when set, an iteration is p = h + beta*p, w_Ap = pcg_custom(w_p),
  // sum(p.*Ap.*mult) by the reduction kernel, update -- instead of the fused strip kernel (Stokes Schur operator)
  std::function<int()> pcg_custom;
  std::vector<semb_field*> fields;     // live fields (for leak-free destroy)
  std::vector<semb_field*> host_tmp;   // cached device fields of the *_host twins
};

// Arguments shared by the operator kernels (strip kernel, seam kernels, generic kernels).
struct OpArgs {
  const double* u = nullptr;      // input field (PCG mode: the residual r)
  const double* pold = nullptr;   // PCG mode: previous search direction
  double* pout = nullptr;         // PCG mode: new search direction p = h + beta*pold
  const double* G11 = nullptr;
  const double* G12 = nullptr;
  const double* G22 = nullptr;
  const double* B = nullptr;
  const double* nu_arr = nullptr;
  const double* k_arr = nullptr;
  const double* M_arr = nullptr;
  const double* mult = nullptr;
  double* out = nullptr;
  const double* halo_lo = nullptr;
  const double* halo_hi = nullptr;
  double nu = 1.0, k = 0.0, prec_b0 = 1.0;
  long long pitch = 0;
  int N = 0, Ex = 0, ney = 0, nxl = 0, nyl = 0;
  int perx = 0;
  int gs = 0;          // 1: fused QQ^T + mask; 0: local operator only
  int precond = 0;     // PCG mode: h = (r ./ B) ./ b0
  int mx0 = 0, mx1 = 0, my0 = 0, my1 = 0;  // Dirichlet flags that apply to THIS slab's boundary lines
  const int* chunk_r0 = nullptr;
  int nchunks = 0;
  int chunk0 = 0;               // first chunk of this launch (pipelined host twin: one slab of chunks per launch)
  int y_begin = 0, y_end = 0;   // row range of the x-seam kernel (0, 0 = all rows)
  const unsigned char* ystart = nullptr;
  const int* xseam = nullptr;
  int nxseam = 0;
  const int* yseam = nullptr;
  int nyseam = 0;
  SembScal* scal = nullptr;
  double* partials = nullptr;
  unsigned* counters = nullptr;
  int pcg = 0;         // 1: PCG mode (read scal, fuse p update and dot)
};

// launchers implemented in the .cu files
int semb_launch_strip(semb_ctx* ctx, const OpArgs& a, const double* hDr, const double* hDs, int nstrips,
                      int nchunks, bool pcg, bool massterm, bool eo);
int semb_strip_regs(int N, bool pcg, bool massterm, int* regs, int* smem, int* occ);

// helpers
static inline long long semb_pitch_for(int nxl) { return ((long long)nxl + 15) / 16 * 16; }
