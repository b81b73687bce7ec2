// Launch wrappers of the streaming kernels in semb_vec.cu (seams, gather-scatter, mask, PCG vector
// updates, reductions, geometry set-up, generic fallbacks).
#pragma once
#include "semb_internal.cuh"

int semb_launch_seam_x(semb_ctx* ctx, const OpArgs& a);
int semb_launch_seam_y(semb_ctx* ctx, const OpArgs& a, int nhalo_lo, int nhalo_hi, bool domask, const P2PArgs& x);
int semb_launch_halo_push(semb_ctx* ctx, const double* out, long long pitch, int nxl, int nyl, double* dst_lo,
                          double* dst_hi, unsigned long long* flag_lo, unsigned long long* flag_hi,
                          unsigned long long epoch, unsigned* counter, const SembScal* scal, int pcg);
int semb_launch_pcg_combine_pap(semb_ctx* ctx, semb_mesh* m);

int semb_launch_gs_x(semb_ctx* ctx, const double* u, double* out, long long pitch, int N, int Ex, int nxl,
                     int nyl, int perx);
int semb_launch_mask(semb_ctx* ctx, const double* u, const double* M, double* out, size_t n);
int semb_launch_axpby(semb_ctx* ctx, double a, const double* x, double b, double* y, size_t n);
int semb_launch_fill(semb_ctx* ctx, double* x, double v, long long pitch, int nxl, int nyl);
int semb_launch_fill_random(semb_ctx* ctx, double* x, long long pitch, int nxl, int nyl, long long gnxl,
                            long long gy0, uint64_t seed);
int semb_launch_mult(semb_ctx* ctx, double* mult, long long pitch, int nr, int ns, int Ex, int Ey, int ey0,
                     int ney, int perx, int pery);
int semb_launch_mask_gen(semb_ctx* ctx, double* M, long long pitch, int nxl, int nyl, int mx0, int mx1, int my0,
                         int my1);
int semb_launch_grid(semb_ctx* ctx, double* x, double* y, long long pitch, int nr, int ns, int Ex, int Ey,
                     int ey0, int ney, const double* d_z0r, const double* d_z0s, int kind, const double* params);
int semb_launch_geom(semb_ctx* ctx, const double* x, const double* y, long long pitch, int nr, int ns, int Ex,
                     int ney, const double* dDr, const double* dDs, const double* d_wr, const double* d_ws,
                     double* J, double* Ji, double* rx, double* ry, double* sx, double* sy, double* B, double* Bi,
                     double* G11, double* G12, double* G22);
// generic (any nr, ns) local operator: two passes through wr/ws temporaries
int semb_launch_generic_local(semb_ctx* ctx, const OpArgs& a, int nr, int ns, const double* dDr,
                              const double* dDs, double* tmp_wr, double* tmp_ws, bool massterm);
// reductions (deterministic): which = 0 dot_mult(a,b,mult), 1 norm_inf(a)
int semb_launch_reduce(semb_ctx* ctx, semb_mesh* m, int which, const double* a, const double* b, const P2PArgs& x,
                       double ref = 0.0);
// PCG vector kernels
// hout (may be NULL): where h = opM(r) is kept when the diagonal preconditioner is on (the strip kernel then stages h)
int semb_launch_pcg_init(semb_ctx* ctx, semb_mesh* m, const double* b, double* x, double* r, double* p, double* hout,
                         int precond, double prec_b0, double tol, long long maxiter, const P2PArgs& xa);
int semb_launch_pcg_update(semb_ctx* ctx, semb_mesh* m, double* x, double* r, const double* p, const double* Ap,
                           double* hout, int precond, double prec_b0, const P2PArgs& xa);
int semb_launch_pcg_dir(semb_ctx* ctx, semb_mesh* m, const double* r, double* p, int precond, double prec_b0);
int semb_launch_mask_dot(semb_ctx* ctx, semb_mesh* m, const OpArgs& a);
int semb_launch_pcg_pack_pap(semb_ctx* ctx, semb_mesh* m);
int semb_launch_pcg_finalize(semb_ctx* ctx, semb_mesh* m, int first);
int semb_launch_reduce_finalize(semb_ctx* ctx, semb_mesh* m, int which);
// generic ABu (ABu.jl:9-37)
int semb_launch_abu_r(semb_ctx* ctx, const double* Br, int mb, int nb, const double* u, int m, int n, long long ldu,
                      double* out, long long ldo);
int semb_launch_abu_s(semb_ctx* ctx, const double* As, int ma, int na, const double* u, int m, int n, long long ldu,
                      double* out, long long ldo);
int semb_launch_grad(semb_ctx* ctx, semb_mesh* m, const double* u, double* ux, double* uy);
int semb_launch_advect_pointwise(semb_ctx* ctx, semb_mesh* m, const double* jux, const double* jtx, const double* juy,
                                 const double* jty, double* out);
// makeRHS! pointwise part (diffusion.jl:55-62): rhs = M .* (B.*f - nu.*lub - sum_i b[i] .* (B.*uh[i]))
int semb_launch_rhs(semb_ctx* ctx, semb_mesh* m, const double* f, const double* nu, const double* lub, int k,
                    const double* const* uh, const double* b, const double* const* adv, const double* a, int mx0,
                    int mx1, int my0, int my1, double* rhs);
int semb_launch_advect_fused(semb_ctx* ctx, semb_mesh* V, semb_mesh* D, const double* T, const double* ux,
                             const double* uy, const double* dJr, const double* dJs, double* out, int* done);
// register-tiled fused dealiased advection for the served (N, M) pairs (semb_advect_tile.cu); nT fields per launch
int semb_launch_advect_tile(semb_ctx* ctx, semb_mesh* V, semb_mesh* D, int nT, const double* const* T, const double* ux,
                            const double* uy, const double* dJr, const double* dJs, double* const* out, int* done);
// Stokes split, element-local kernels (semb_stokes.cu): grad^T (optionally of W .* u), B .* div(ux,uy), (M.*g).*Bi./b0
int semb_launch_gradT(semb_ctx* ctx, semb_mesh* m, const double* u, const double* W, double* ox, double* oy);
int semb_launch_diver_local(semb_ctx* ctx, semb_mesh* m, const double* ux, const double* uy, double* out);
int semb_launch_hinv_mid(semb_ctx* ctx, semb_mesh* m, const double* g, double b0, int mx0, int mx1, int my0, int my1,
                         double* out);
// register-tiled Stokes element kernels (semb_stokes_tile.cu): diverT (transpose = 1) / diver (0) in one launch
int semb_launch_stokes_tile(semb_ctx* ctx, semb_mesh* V, semb_mesh* P, int transpose, const double* in1, const double* in2,
                            double* out1, double* out2, const double* dJr, const double* dJs, double sign, int* done);
int semb_launch_gs_fused(semb_ctx* ctx, semb_mesh* m, const double* u, double* out, int mode, double b0, int mx0, int mx1,
                         int my0, int my1, int stage);
int semb_launch_pcg_set_pap(semb_ctx* ctx, semb_mesh* m);
