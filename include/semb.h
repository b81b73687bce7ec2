/*
 * semb.h -- C ABI of libsemb.so: the B200-native (sm_100a) implementation of the matrix-free
 * spectral-element operator hot path of vpuri3/SpectralElements.jl.
 *
 * The reference is pure Julia and has NO FFI layer for this path: its "operator
 * API" is the set of exported Julia functions below.  Each entry point here names the reference
 * function (file:line under /root/reference/src) it replaces; INTEGRATION.md shows the Julia
 * `ccall` method a maintainer adds per function so examples (p2d, d2d, cd2d) run unmodified.
 *
 * Conventions
 *   - Every function returns int: 0 = SEMB_OK, <0 = error (message via semb_last_error()),
 *     +1 = SEMB_NOT_CONVERGED (pcg hit maxiter; mirrors pcg.jl:39 "warn and return the iterate").
 *   - No C++ exception crosses the ABI.  Only plain pointers / ints / doubles in signatures.
 *   - Host arrays are column-major (Julia / Fortran order) FP64, shape (nr*Ex) x (ns*Ey_local):
 *     first index = x (contiguous).  The library never retains a host pointer after returning.
 *   - The library owns all device memory.  Internally a field is stored with a padded row pitch
 *     (multiple of 16 doubles); upload/download convert.
 *   - Multi-GPU: one process per GPU (rank).  Element rows are split into contiguous y-slabs;
 *     `Ey` arguments below are GLOBAL element-row counts, host arrays hold the LOCAL slab
 *     (rows [ey0*ns, (ey0+ney)*ns) of the global matrix, see semb_partition).
 *   - There is NO CPU fallback: every compute entry point fails with SEMB_ECUDA without a GPU.
 */
#ifndef SEMB_H
#define SEMB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMB_OK 0
#define SEMB_NOT_CONVERGED 1
#define SEMB_EINVAL (-1)  /* bad argument / size mismatch (Julia: DimensionMismatch / InexactError, ABu.jl:16,26) */
#define SEMB_ECUDA (-2)   /* CUDA runtime error, or no CUDA device */
#define SEMB_ENCCL (-3)   /* NCCL error */
#define SEMB_ENOMEM (-4)
#define SEMB_ESTATE (-5)  /* call sequence error (e.g. comm not initialised) */

typedef struct semb_ctx semb_ctx;     /* one GPU + stream (+ NCCL communicator) */
typedef struct semb_mesh semb_mesh;   /* device copy of Mesh{T} operator data, mesh.jl:25-64 */
typedef struct semb_field semb_field; /* one device-resident (nr*Ex) x (ns*Ey_local) FP64 array */

/* built-in deformation maps for semb_mesh_create_deform (mesh.jl:108 `deform(x,y)`) */
#define SEMB_DEFORM_IDENTITY 0 /* fixU, mesh.jl:6-8 */
#define SEMB_DEFORM_ANNULUS 1  /* geom.jl:40-49, params = {r0, r1, span} */
#define SEMB_DEFORM_WAVY 2     /* bench map: (x+d, y+d), d = a sin(pi x) sin(pi y); params = {a} */

/* selectors for semb_mesh_get (field names of Mesh{T}, mesh.jl:41-62) */
enum semb_mesh_array {
  SEMB_X = 0, SEMB_Y, SEMB_JAC, SEMB_JACI, SEMB_RX, SEMB_RY, SEMB_SX, SEMB_SY,
  SEMB_B, SEMB_BI, SEMB_G11, SEMB_G12, SEMB_G22, SEMB_MULT, SEMB_MESH_ARRAY_COUNT
};

/* ---- library / context ------------------------------------------------------------------ */
int semb_version(void);
const char* semb_last_error(void); /* thread-local message of the last failing call */
/* Create a context on CUDA device `device` (owns a non-default stream + scratch). */
int semb_init(int device, semb_ctx** ctx);
int semb_finalize(semb_ctx* ctx);
int semb_sync(semb_ctx* ctx);                         /* cudaStreamSynchronize of the ctx stream */
int semb_stream(semb_ctx* ctx, void** cuda_stream);   /* the cudaStream_t every kernel is launched on */
/* Device timing helpers (CUDA events on the ctx stream), used by bench.py. */
int semb_timer_start(semb_ctx* ctx);
int semb_timer_stop(semb_ctx* ctx, double* elapsed_ms); /* synchronises, returns ms since start */
/* Number of library kernels launched on this ctx since creation (bench.py's gpu_launches). */
int semb_launch_count(semb_ctx* ctx, long long* n);
/* Writes >L2-size scratch to evict L2 between timed repetitions. */
int semb_flush_l2(semb_ctx* ctx);
/* Per-launch CUDA-event timing of the fused strip kernel (the dominant kernel): enable for up to
 * max_launches launches (0 disables), then read the summed duration and the launch count. */
int semb_profile_enable(semb_ctx* ctx, int max_launches);
int semb_profile_read(semb_ctx* ctx, double* total_ms, int* launches);
/* Page-locked host memory for the *_host twins (bench.py's end-to-end leg). */
int semb_alloc_pinned(size_t bytes, void** p);
int semb_free_pinned(void* p);

/* ---- multi-GPU plumbing (new; the reference is single-process) ------------------ */
/* Contiguous y-slab owned by `rank`: element rows [ey0, ey0+ney). No GPU needed. */
int semb_partition(int Ey, int nranks, int rank, int* ey0, int* ney);
/* Chunk count of the strip-kernel launch plan (grid = strips x chunks) for a slab of `ney` element rows on `slots`
 * resident CTA slots: minimises waves x (element rows of the longest chunk + 1/2).  No reference counterpart (the
 * reference has no launch geometry); exported so the host logic is testable without a GPU. */
int semb_plan_chunks(int nstrips, int ney, int slots, int* nchunks);
/* Neighbour ranks of a slab (periodic wrap included); rank_lo/rank_hi = -1 when absent. No GPU needed. */
int semb_halo_plan(int nranks, int rank, int pery, int* halo_lo, int* halo_hi, int* rank_lo, int* rank_hi);
/* rank 0: fill a 128-byte ncclUniqueId; the host language broadcasts it (torch.distributed / MPI). */
int semb_comm_unique_id(char id[128]);
/* every rank: join the communicator (ncclCommInitRank on the ctx device). */
int semb_comm_init(semb_ctx* ctx, int nranks, int rank, const char id[128]);
int semb_comm_info(semb_ctx* ctx, int* nranks, int* rank);
int semb_comm_barrier(semb_ctx* ctx);                 /* allreduce + stream sync */
int semb_comm_allreduce_max(semb_ctx* ctx, double* v, int n); /* host doubles, in place */

/* ---- 1-D setup helpers (host C++, no GPU) ---------------------------------------------------- */
/* FastGaussQuadrature.gausslobatto(n) as used at mesh.jl:70-71, semmesh.jl:11 */
int semb_gausslobatto(int n, double* z, double* w);
/* derivMat(x), derivMat.jl:9-35; D is n x n column-major */
int semb_deriv_mat(int n, const double* x, double* D);
/* interpMat(xo, xi), interp.jl:10-35; J is no x ni column-major */
int semb_interp_mat(int no, const double* xo, int ni, const double* xi, double* J);
/* semmesh(E, n), semmesh.jl:9-27; z, w have E*n entries */
int semb_semmesh(int E, int n, double* z, double* w);
/* bdfExtK(t; k), time.jl:31-53; t has nt entries, a has k, b has k+1 */
int semb_bdf_ext_k(int nt, const double* t, int k, double* a, double* b);

/* ---- mesh --------------------------------------------------------------------------------------- */
/* Mesh(nr,ns,Ex,Ey,ifperiodic,deform) from deformed coordinates, mesh.jl:66-133: runs jac
 * (jac.jl:24-40) and the B/G11/G12/G22 factors (mesh.jl:114-123) and mult (mesh.jl:94-96) ON DEVICE.
 * x, y: host, local slab, already deformed.  Dr (nr x nr), Ds (ns x ns) column-major; wr, ws weights. */
int semb_mesh_create_xy(semb_ctx* ctx, int nr, int ns, int Ex, int Ey, int perx, int pery,
                        const double* Dr, const double* Ds, const double* wr, const double* ws,
                        const double* x, const double* y, semb_mesh** mesh);
/* Same, but the grid (semmesh.jl + ndgrid.jl) and a built-in deformation are generated on device:
 * nothing of size O(n) touches the host (needed for the 1e8-DOF meshes). */
int semb_mesh_create_deform(semb_ctx* ctx, int nr, int ns, int Ex, int Ey, int perx, int pery,
                            int deform_kind, const double* params, int nparams, semb_mesh** mesh);
/* Mesh from ready-made operator arrays (what a Julia `Mesh` already holds); B may be NULL. */
int semb_mesh_create_arrays(semb_ctx* ctx, int nr, int ns, int Ex, int Ey, int perx, int pery,
                            const double* Dr, const double* Ds, const double* G11, const double* G12,
                            const double* G22, const double* B, semb_mesh** mesh);
int semb_mesh_destroy(semb_mesh* mesh);
/* nxl = nr*Ex, nyl = ns*ney (local), ey0/ney = this rank's slab. Any out pointer may be NULL. */
int semb_mesh_dims(semb_mesh* mesh, int* nr, int* ns, int* Ex, int* Ey, int* nxl, int* nyl, int* ey0, int* ney);
/* Copy one Mesh array (enum semb_mesh_array) to a host nxl x nyl buffer. */
int semb_mesh_get(semb_mesh* mesh, int which, double* host_out);
/* Upload one Mesh array (e.g. the rx, ry, sx, sy a Julia Mesh already holds; needed by semb_grad / semb_advect
 * when the mesh was created with semb_mesh_create_arrays). */
int semb_mesh_set(semb_mesh* mesh, int which, const double* host_in);
/* Dr / Ds as held on the device side (column-major). */
int semb_mesh_get_D(semb_mesh* mesh, double* Dr, double* Ds);
/* generateMask(bc, msh), mesh.jl:149-175: bc = "DDNN" = [xmin,xmax,ymin,ymax]; writes 0/1 doubles. */
int semb_generate_mask(semb_mesh* mesh, const char bc[4], double* host_out);

/* ---- fields ---------------------------------------------------------------------------------------- */
int semb_field_create(semb_mesh* mesh, semb_field** f); /* zero-initialised */
int semb_field_destroy(semb_field* f);
int semb_field_upload(semb_field* f, const double* host);   /* nxl x nyl column-major */
int semb_field_download(semb_field* f, double* host);
int semb_field_fill(semb_field* f, double value);
int semb_field_copy(semb_field* dst, const semb_field* src);
/* Fill with the portable splitmix64 uniform(-1,1) stream, global column-major index. */
int semb_field_fill_random(semb_field* f, uint64_t seed);
/* y = a*x + b*y (pointwise; host-language broadcasts like `u .+= ub`, diffusion.jl:75) */
int semb_field_axpby(double a, const semb_field* x, double b, semb_field* y);
/* Raw device pointer + pitch (in doubles), for zero-copy interop (e.g. torch.from_blob / CuPtr). */
int semb_field_devptr(semb_field* f, void** dptr, long long* pitch);

/* ---- operators on device-resident fields --------------------------------------------------------- */
/* Optional array coefficients: pass NULL to use the scalar.  `out` must not alias `u`. */
/* laplace(u,Dr,Ds,G11,G12,G22), lapl.jl:70-81 == lapl(u,msh), lapl.jl:26-36 (no gs, no mask) */
int semb_lapl(semb_mesh* m, const semb_field* u, semb_field* out);
/* hlmz(u,nu,k,msh), hlmz.jl:12-19: out = nu .* lapl(u) + k .* (B .* u); nu/k scalar or array */
int semb_hlmz(semb_mesh* m, const semb_field* u, const semb_field* nu_arr, double nu,
              const semb_field* k_arr, double k, semb_field* out);
/* mass(u,msh), mass.jl:12-22: out = B .* u */
int semb_mass(semb_mesh* m, const semb_field* u, semb_field* out);
/* gatherScatter(u,msh), gatherScatter.jl:8-21: out = QQ^T u (x pairs first, then y pairs;
 * bitwise equal to the reference's dense product).  Multi-GPU: halo exchange inside. */
int semb_gather_scatter(semb_mesh* m, const semb_field* u, semb_field* out);
/* mask(u,M), mask.jl:10-18: out = M .* u; M = NULL copies (length(M)==0 branch). */
int semb_mask(semb_mesh* m, const semb_field* u, const semb_field* M, semb_field* out);
/* mask(u, generateMask(bc,msh)) without materialising M (mask.jl:14 with mesh.jl:149-175 flags). */
int semb_mask_bc(semb_mesh* m, const semb_field* u, const char bc[4], semb_field* out);
/* The fused unit opLHS(u,dfn), diffusion.jl:36-45 / convectionDiffusion.jl:76-85:
 * out = mask(gatherScatter(hlmz(u,nu,k,msh)), M).  Mask: bc = "DDNN"-style flags (NULL = no mask)
 * or an explicit 0/1 array M_arr (overrides bc).  One fused strip kernel + two seam kernels. */
int semb_oplhs(semb_mesh* m, const semb_field* u, const semb_field* nu_arr, double nu,
               const semb_field* k_arr, double k, const char* bc, const semb_field* M_arr,
               semb_field* out);
/* jac(x,y,Dr,Ds), jac.jl:24-40, as a standalone op on fields (Mesh creation uses the same kernel). */
int semb_jac(semb_mesh* m, const semb_field* x, const semb_field* y, semb_field* J, semb_field* Ji,
             semb_field* rx, semb_field* ry, semb_field* sx, semb_field* sy);
/* sum(a .* b .* mult), pcg.jl:45,52 (deterministic fixed-order reduction; all-reduced over ranks) */
int semb_dot_mult(semb_mesh* m, const semb_field* a, const semb_field* b, double* result);
/* norm(a, Inf), pcg.jl:36 */
int semb_norm_inf(semb_mesh* m, const semb_field* a, double* result);

/* ---- PCG --------------------------------------------------------------------------------------------- */
/* pcg(b, opA; opM, mult, tol, maxiter), pcg.jl:16-60, with opA = the fused opLHS above and
 * opM = identity (diffusion.jl:47-49) or u ./ B ./ b0 (convectionDiffusion.jl:87-91).
 * The whole loop is device-resident (alpha/beta/convergence live in device memory; the host polls a
 * done flag).  x is overwritten (pcg! semantics, pcg.jl:64-79: zero initial guess always).
 * maxiter < 0 => length(b) (global).  iters / resinf (final norm(r,Inf)) may be NULL. */
typedef struct semb_pcg_opts {
  double nu;                /* scalar viscosity (used when nu_arr == NULL) */
  const semb_field* nu_arr; /* array viscosity, diffusion.jl:11,40 */
  double k;                 /* scalar Helmholtz coefficient = bdfB[1] */
  const semb_field* k_arr;  /* array coefficient, examples/poissonNonlin.jl:84,87 */
  const char* bc;           /* "DDDD"-style flags or NULL */
  const semb_field* M_arr;  /* explicit mask array or NULL */
  int precond;              /* 0 = identity, 1 = u ./ B ./ prec_b0, 2 = the mesh's FDM preconditioner (semb_fdm_create) */
  double prec_b0;
  double tol;               /* absolute, on norm(r,Inf); pcg.jl:20 default 1e-8 */
  long long maxiter;        /* <0 => length(b) */
  int check_every;          /* host polls the device done flag every this many iterations (<=0: auto) */
} semb_pcg_opts;
int semb_pcg(semb_mesh* m, const semb_pcg_opts* opts, const semb_field* b, semb_field* x,
             long long* iters, double* resinf);
/* One PCG iteration's worth of kernels `n` times without convergence polling (bench: iterations/s).
 * State must have been initialised by semb_pcg_begin. */
int semb_pcg_begin(semb_mesh* m, const semb_pcg_opts* opts, const semb_field* b, semb_field* x);
int semb_pcg_iterate(semb_mesh* m, int n);
int semb_pcg_status(semb_mesh* m, long long* iters, double* resinf, int* done);

/* ---- FDM Laplacian / Helmholtz preconditioner -----------------------------------------------------
 * The reference holds it as commented-out sketches: lapl_fdm(b,Bi,Sx,Sy,Sxi,Syi,Di) (lapl.jl:105-119) and its set-up from
 * eigen(Ax,Bx), eigen(Ay,By) (examples/p2d_explicit.jl:109-141).  Built here in the form in which it works as the opM of
 * pcg (pcg.jl:37): the same tensor solve on every element extended by one node into its neighbours, combined
 * symmetrically with counting weights (the CPU checker restates it as fdm_schwarz; 5-8x fewer iterations).
 * One per mesh: semb_fdm_create registers it, semb_pcg_opts.precond = 2 uses it, semb_fdm_apply is h = opM(r).
 * Several ranks (y-slabs): create / destroy / apply are collective; needs the peer-memory transport (CUDA IPC), nr >= 4
 * and at least two element rows per rank; the result has the single-rank bits. */
typedef struct semb_fdm semb_fdm;
int semb_fdm_create(semb_mesh* m, const char bc[4], double nu, double k, semb_fdm** out);
int semb_fdm_destroy(semb_mesh* m);
int semb_fdm_apply(semb_mesh* m, const semb_field* r, semb_field* out);
int semb_fdm_apply_host(semb_mesh* m, const double* r, double* out);
/* 1-D reference decomposition behind it (host only, no GPU): A S = B S diag(lam), S' B S = I for A = D' diag(w) D,
 * B = diag(w) extended by one node into neighbours of equal size; kinds: 0 neighbour, 1 Dirichlet, 2 free boundary.
 * S: (n+2) x (n+2) column-major, lam: n+2 (+inf padding). */
int semb_fdm_tables(int n, const double* D, const double* w, int left_kind, int right_kind, double* S, double* lam);

/* ---- implicit diffusion driver, device resident (the caller that defines the fused unit) ---- */
/* Diffusion(bc,msh;Ti,Tf,dt,k), diffusion.jl:20-34, with its Field (mesh.jl:179-195: u, uh[1..k], ub, M) and
 * TimeStepper (time.jl:70-99) kept in HBM across steps.  The user closures setBC!/setForcing!/setVisc!
 * (diffusion.jl:98-100) stay in the host language: between begin_step and finish_step the host uploads
 * whatever they changed into the fields returned by semb_diffusion_field. */
typedef struct semb_diffusion semb_diffusion;
enum semb_diffusion_field_id {
  SEMB_DFN_U = 0, SEMB_DFN_UB, SEMB_DFN_NU, SEMB_DFN_F, SEMB_DFN_RHS, SEMB_DFN_VX, SEMB_DFN_VY, SEMB_DFN_UH0 = 8 /* + i */
};
int semb_diffusion_create(semb_mesh* m, const char bc[4], double Ti, double Tf, double dt, int k, semb_diffusion** d);
int semb_diffusion_destroy(semb_diffusion* d);
int semb_diffusion_field(semb_diffusion* d, int which, semb_field** f);
/* first half of evolve! (diffusion.jl:89-96): updateHist!(fld) (mesh.jl:199-215), time history (mesh.jl:217-224),
 * istep += 1, time[1] += dt, bdfExtK! (time.jl:55-68).  Returns the new time for the host closures. */
int semb_diffusion_begin_step(semb_diffusion* d, double* time, long long* istep);
/* second half (diffusion.jl:102-103): makeRHS! (diffusion.jl:51-65: mass(f) - nu.*lapl(ub) - sum bdfB[1+i].*mass(uh[i]);
 * mask THEN gatherScatter) and solve! (diffusion.jl:67-77: pcg! with opLHS, then u .+= ub). */
int semb_diffusion_finish_step(semb_diffusion* d, double tol, long long* iters, double* resinf);
/* time[k+1], bdfA[k], bdfB[k+1] (time.jl:70-82); any pointer may be NULL */
int semb_diffusion_state(semb_diffusion* d, double* time, double* bdfA, double* bdfB, long long* istep);
/* The opM of the step's solve (pcg.jl:37): kind 0 = the reference's (identity in diffusion.jl:71, opPrecond = u./B./b0 in
 * convectionDiffusion.jl:87-91,118), kind 2 = the FDM preconditioner of nu*lapl + b0*mass (lapl.jl:105-119; registered on the
 * mesh like semb_fdm_create, constant viscosity only -- otherwise the step keeps kind 0).  Opt-in: same solution to the
 * solver tolerance, several times fewer iterations, but not the reference's iteration counts. */
int semb_diffusion_set_precond(semb_diffusion* d, int kind);

/* ---- explicit dealiased convection and the ConvectionDiffusion driver -------------------- */
/* grad(u,msh), grad.jl:15-34: ux = rx.*ur + sx.*us, uy = ry.*ur + sy.*us (needs a mesh built from x,y) */
int semb_grad(semb_mesh* m, const semb_field* u, semb_field* ux, semb_field* uy);
/* advect(T,ux,uy,mshV,mshD,Jr,Js), advect.jl:45-64: gradient on mshV, interpolation of Tx,Ty,ux,uy to the
 * dealiasing mesh mshD (ABu(Js,Jr,.) with Jr = interpMat(mshD.zr,mshV.zr)), pointwise product with mshD.B,
 * projection back (ABu(Js',Jr',.)).  mD = NULL: the un-dealiased form advect(T,ux,uy,msh), advect.jl:27-43.
 * T, ux, uy, out are fields of mV.  mD must have the same Ex, Ey, periodicity (and partition) as mV. */
int semb_advect(semb_mesh* mV, semb_mesh* mD, const semb_field* T, const semb_field* ux, const semb_field* uy,
                semb_field* out);
/* ConvectionDiffusion(name,fld,vx,vy,tstep,mshD,...), convectionDiffusion.jl:31-56: same handle type and
 * begin/finish protocol as semb_diffusion; finish_step runs makeRHS! (convectionDiffusion.jl:93-110: the explicit
 * term exH[i] = -advect(uh[i],vx,vy,...) enters with bdfA[i]) and solve! (:112-122, opM = u./B./bdfB[1]).
 * Extra fields: SEMB_DFN_VX, SEMB_DFN_VY. */
int semb_convdiff_create(semb_mesh* mV, semb_mesh* mD, const char bc[4], double Ti, double Tf, double dt, int k,
                         semb_diffusion** d);

/* ---- Stokes pressure/velocity split ---------------------------------------------------------- */
/* RECONSTRUCTION: the reference's diver.jl / stokes.jl are not executable as shipped (stokes.jl is not included,
 * SpectralElements.jl:53; undefined names at diver.jl:22-27,96 and stokes.jl:114-120).  These entry points follow the
 * docstring math (diver.jl:5-16,35-51,67-72; stokes.jl:5-51).  Deviations from the literal code, all flagged where
 * they occur: gradT applies Dr' along r and Ds' along s (grad.jl:56-60 swaps them); `msh` in diver.jl:22-27 is mshV;
 * `Mvx` in diver.jl:96,101 is the mask of the component, and approxHlmzInv's missing b0 argument (diver.jl:83-84) is
 * the handle's b0; opStokesLHS returns Eq (stokes.jl:120 returns the undefined Eu); the right-hand side is gathered
 * on mshP (stokes.jl:139 names mshV) and the pressure PCG weighs its inner products with mshP.mult (stokes.jl:151
 * names mshV.mult): both live on the pressure mesh. */
/* gradT(u,msh), grad.jl:44-63: ux = Dr'(rx.*u) + Ds'(sx.*u), uy = Dr'(ry.*u) + Ds'(sy.*u), Dr' along r and Ds' along s
 * (grad.jl:56-60 swaps the directions; the docstring :38-42 does not) */
int semb_gradT(semb_mesh* m, const semb_field* u, semb_field* ux, semb_field* uy);
/* approxHlmzInv(u,b0,mshV), diver.jl:92-104: mask(gs( mask(gs(u)) .* Bi ./ b0 )); bc = the component's "DDNN" flags */
int semb_approx_hlmz_inv(semb_mesh* m, const semb_field* u, double b0, const char bc[4], semb_field* out);
/* Stokes(bcVX,bcVY,mshV,mshD,mshP), stokes.jl:75-108, reduced to the pressure system: JrPV = interpMat(mshV.zr,mshP.zr)
 * (:101-102), the velocity masks, b0 (the bdfB[1] of approxHlmzInv, stokes.jl:133-134) and work fields.
 * mshP must share Ex, Ey, periodicity and partition with mshV (pressure order nr-2, examples/semPS.jl:31). */
typedef struct semb_stokes semb_stokes;
int semb_stokes_create(semb_mesh* mV, semb_mesh* mP, const char bcVX[4], const char bcVY[4], double b0, semb_stokes** s);
int semb_stokes_destroy(semb_stokes* s);
/* diver(ux,uy,mshV,Jr,Js), diver.jl:17-31: ABu(Js',Jr', B .* (dx ux + dy uy)) -> field of mshP */
int semb_diver(semb_stokes* s, const semb_field* ux, const semb_field* uy, semb_field* out);
/* diverT(pr,mshV,Jr,Js), diver.jl:53-63: gradT(B .* ABu(Js,Jr,pr)) -> two fields of mshV */
int semb_diverT(semb_stokes* s, const semb_field* pr, semb_field* qx, semb_field* qy);
/* opStokesLHS(q,sks), stokes.jl:110-121: gatherScatter(stokesOp(q), mshP), stokesOp = -DD HH^-1 DD' (diver.jl:73-89) */
int semb_stokes_op(semb_stokes* s, const semb_field* q, semb_field* out);
/* makeStokesRHS!, stokes.jl:128-141: rhs = gatherScatter(diver(vx,vy), mshP) */
int semb_stokes_rhs(semb_stokes* s, const semb_field* vx, const semb_field* vy, semb_field* rhs);
/* solveStokes!, stokes.jl:143-154: pcg(rhs, opStokesLHS; mult = mshP.mult), pcg.jl:16-60.  Returns SEMB_NOT_CONVERGED
 * at maxiter (<0: length(rhs)).  iters / resinf may be NULL. */
int semb_stokes_solve(semb_stokes* s, const semb_field* rhs, semb_field* dp, double tol, long long maxiter,
                      long long* iters, double* resinf);
/* pressureProject!, stokes.jl:159-177: dp = solveStokes; vx += HH^-1 DD'x dp, vy += ..., pr += dp (pr may be NULL) */
int semb_stokes_project(semb_stokes* s, semb_field* vx, semb_field* vy, semb_field* pr, double tol, long long maxiter,
                        long long* iters, double* resinf);

/* ---- host-pointer convenience twins (value semantics of the Julia functions) ------------------- */
/* Each uploads its inputs, runs the device op, downloads `out` (fresh array in Julia). */
int semb_lapl_host(semb_mesh* m, const double* u, double* out);
int semb_hlmz_host(semb_mesh* m, const double* u, const double* nu_arr, double nu,
                   const double* k_arr, double k, double* out);
int semb_mass_host(semb_mesh* m, const double* u, double* out);
int semb_gather_scatter_host(semb_mesh* m, const double* u, double* out);
int semb_mask_host(semb_mesh* m, const double* u, const double* M, double* out);
int semb_oplhs_host(semb_mesh* m, const double* u, const double* nu_arr, double nu,
                    const double* k_arr, double k, const char* bc, const double* M_arr, double* out);
int semb_pcg_host(semb_mesh* m, const semb_pcg_opts* opts_scalars, const double* nu_arr,
                  const double* k_arr, const double* M_arr, const double* b, double* x,
                  long long* iters, double* resinf);
/* ABu(As,Br,u), ABu.jl:9-37, general rectangular blocks.  As is ma x na, Br is mb x nb (column-major);
 * a NULL pointer / zero size is Julia's `[]` (identity).  u is m x n; out is (m*mb/nb) x (n*ma/na).
 * Needs no mesh. */
int semb_abu_host(semb_ctx* ctx, const double* As, int ma, int na, const double* Br, int mb, int nb,
                  const double* u, int m, int n, double* out);

/* ---- explicit-argument operator forms on plain arrays (no Mesh), host twins ---------------------------------- */
/* laplace(u,Dr,Ds,G11,G12,G22), lapl.jl:70-81 (Jr = Js = NULL) and the dealiased laplace(u,Jr,Js,Dr,Ds,G11,G12,G22),
 * lapl.jl:83-103, as examples/p2d_explicit.jl:183 and examples/semPS.jl:168 use them: u is m x n = (nr*Ex) x (ns*Ey);
 * Jr is nrd x nr, Js nsd x ns (column-major); the G factors live on the (m*nrd/nr) x (n*nsd/ns) grid. */
int semb_laplace_host(semb_ctx* ctx, int m, int n, const double* Dr, int nr, const double* Ds, int ns,
                      const double* Jr, int nrd, const double* Js, int nsd, const double* G11, const double* G12,
                      const double* G22, const double* u, double* out);
/* core of mass(u,M,B,Jr,Js,QQtx,QQty,mult), mass.jl:32-50: out = ABu(Js',Jr', B .* ABu(Js,Jr,u)); NULL Jr/Js/B are
 * Julia's `[]` (identity / no weight, mass.jl:38).  The mult hook, gatherScatter and mask that follow are ABu / a.*b. */
int semb_mass_explicit_host(semb_ctx* ctx, int m, int n, const double* Jr, int nrd, int nr, const double* Js, int nsd,
                            int ns, const double* B, const double* u, double* out);
/* out = a .* b on n doubles: mask(u,M), mask.jl:14, and the `d .* mult` hooks (lapl.jl:62, mass.jl:44) for arrays
 * that belong to no Mesh */
int semb_mul_host(semb_ctx* ctx, size_t n, const double* a, const double* b, double* out);

/* ---- diagnostics / tuning hooks (no reference counterpart) ------------------------------------- */
/* registers / shared memory / resident CTAs per SM of the fused strip kernel for polynomial size N */
int semb_strip_kernel_info(int N, int pcg, int massterm, int* regs, int* smem, int* occ);
/* launch plan of the fused operator: strips x chunks, seam counts, fast = templated strip kernel used */
int semb_mesh_plan(semb_mesh* mesh, int* nstrips, int* nchunks, int* nxseam, int* nyseam, int* fast);
/* override the number of y chunks (tests exercise every seam configuration with it) */
int semb_mesh_set_chunks(semb_mesh* mesh, int nchunks);
/* 1 when the interface sums, the halo exchange and the PCG reduction run inside the strip kernel (one launch per
   apply: single rank, or peer memory between ranks), 0 when the separate seam kernels / NCCL path is used */
int semb_mesh_fused_tail(semb_mesh* mesh, int* on);
/* timing instrumentation of the fused tail (builds with -DSEMB_TAIL_TIMING only; SEMB_EINVAL otherwise) */
int semb_mesh_debug_read(semb_mesh* mesh, long long* host, int ncta);
/* CTA rows (grid.y) of the strip kernel; a CTA row marches through one chunk, or two at a rank boundary */
int semb_mesh_groups(semb_mesh* mesh, int* ngroups);
/* Multi-GPU health: SEMB_OK, or SEMB_ENCCL once any kernel of this mesh gave up waiting for a peer rank (bounded
   waits on peer memory, SEMB_PEER_TIMEOUT_MS, default 20 s).  The reference is single-process (no counterpart);
   SURVEY 5 asks that a lost peer surfaces as a status code instead of a hang. */
int semb_mesh_peer_status(semb_mesh* mesh);

#ifdef __cplusplus
}
#endif
#endif /* SEMB_H */
