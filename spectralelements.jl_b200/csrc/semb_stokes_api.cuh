// C-ABI entry points of the Stokes pressure/velocity split; textually included at the end of semb_api.cu
// (uses its static helpers).  The reference's diver.jl / stokes.jl are not executable as shipped: these
// functions reconstruct gradT, diver, diverT, approxHlmzInv, stokesOp, opStokesLHS, makeStokesRHS, solveStokes and
// pressureProject from the docstring math; the deviations from the literal code are listed in include/semb.h.
#pragma once

struct semb_stokes {
  semb_mesh *V = nullptr, *P = nullptr;
  char bcx[5] = {0}, bcy[5] = {0};
  double b0 = 1.0;
  // JrPV = interpMat(mshV.zr, mshP.zr), JsPV = interpMat(mshV.zs, mshP.zs) (stokes.jl:101-102), column-major, + transposes
  double *dJr = nullptr, *dJs = nullptr, *dJrT = nullptr, *dJsT = nullptr;
  double* mid = nullptr;                                      // mixed-resolution intermediate of the two-pass ABu
  semb_field *v1 = nullptr, *v2 = nullptr, *v3 = nullptr, *v4 = nullptr;  // work fields on mshV
  semb_field *p_Au = nullptr, *p_rhs = nullptr, *p_dp = nullptr;  // on mshP
  semb_field* p_rhs2 = nullptr;  // pressureProject's right-hand side (created at first use)
};

extern "C" int semb_gradT(semb_mesh* m, const semb_field* u, semb_field* ux, semb_field* uy) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "gradT(u)"));
  SEMB_TRY(check_field(m, ux, "gradT(ux)"));
  SEMB_TRY(check_field(m, uy, "gradT(uy)"));
  SEMB_REQUIRE(u != ux && u != uy && ux != uy, "gradT: outputs must not alias");
  SEMB_REQUIRE(m->arr[SEMB_RX] && m->arr[SEMB_SY], "gradT: mesh has no metric terms (create it from x,y)");
  return semb_launch_gradT(m->ctx, m, u->d, nullptr, ux->d, uy->d);
}

// gatherScatter in one pass with a pointwise epilogue (semb_gs_fused_kernel; mode 0 none, 1 (M.*g).*Bi./b0, 2 M.*g); with
// neighbour ranks the slab's boundary rows are completed by the halo exchange before their epilogue.  dst must not alias src.
static int gs_one_pass(semb_mesh* m, const double* src, double* dst, int mode, double b0, const MaskFlags& f) {
  SEMB_TRY(semb_launch_gs_fused(m->ctx, m, src, dst, mode, b0, f.mx0, f.mx1, f.my0, f.my1, 0));
  if (m->halo_lo || m->halo_hi) {
    unsigned long long eph = 0;
    SEMB_TRY(halo_exchange(m, dst, 0, &eph));
    OpArgs y;
    fill_common(m, y);
    y.halo_lo = m->d_halo_lo;
    y.halo_hi = m->d_halo_hi;
    if (m->p2p) {
      y.halo_lo = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 0);
      y.halo_hi = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 1);
    }
    y.out = dst;
    y.nyseam = 0;  // only the received rows
    SEMB_TRY(semb_launch_seam_y(m->ctx, y, m->halo_lo, m->halo_hi, false, p2p_args(m, eph)));
    SEMB_TRY(semb_launch_gs_fused(m->ctx, m, src, dst, mode, b0, f.mx0, f.mx1, f.my0, f.my1, 1));
  }
  return SEMB_OK;
}

extern "C" int semb_approx_hlmz_inv(semb_mesh* m, const semb_field* u, double b0, const char bc[4], semb_field* out) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "approxHlmzInv(u)"));
  SEMB_TRY(check_field(m, out, "approxHlmzInv(out)"));
  SEMB_REQUIRE(u != out, "approxHlmzInv: out must not alias u");
  SEMB_REQUIRE(m->arr[SEMB_BI], "approxHlmzInv: mesh has no Bi");
  SEMB_REQUIRE(b0 != 0.0, "approxHlmzInv: b0 must be non-zero");
  MaskFlags f;
  SEMB_TRY(parse_bc(m, bc, &f));
  if (!m->w_t1) SEMB_TRY(semb_field_create(m, &m->w_t1));
  semb_field* t = m->w_t1;
  SEMB_REQUIRE(t != u && t != out, "approxHlmzInv: argument aliases the mesh work field");
  if (!getenv("SEMB_NO_TILED_STOKES")) {
    // two one-pass kernels (gatherScatter with the pointwise step as its epilogue), same bits as the chain below;
    // with neighbour ranks the slab's boundary rows are completed by the halo exchange before their epilogue
    const int modes[2] = {1, bc ? 2 : 0};
    const double* src[2] = {u->d, t->d};
    double* dst[2] = {t->d, out->d};
    for (int k = 0; k < 2; ++k) SEMB_TRY(gs_one_pass(m, src[k], dst[k], modes[k], b0, f));  // diver.jl:95-98 | :100-101
    return SEMB_OK;
  }
  SEMB_TRY(semb_gather_scatter(m, u, t));                                                            // diver.jl:95
  SEMB_TRY(semb_launch_hinv_mid(m->ctx, m, t->d, b0, f.mx0, f.mx1, f.my0, f.my1, out->d));           // :96-98
  SEMB_TRY(semb_gather_scatter(m, out, t));                                                          // :100
  return bc ? semb_mask_bc(m, t, bc, out) : semb_field_copy(out, t);                                 // :101
}

extern "C" int semb_stokes_destroy(semb_stokes* s) {
  if (!s) return SEMB_OK;
  cudaFree(s->dJr);
  cudaFree(s->dJs);
  cudaFree(s->dJrT);
  cudaFree(s->dJsT);
  cudaFree(s->mid);
  semb_field* fs[] = {s->v1, s->v2, s->v3, s->v4, s->p_Au, s->p_rhs, s->p_dp, s->p_rhs2};
  for (semb_field* f : fs) semb_field_destroy(f);
  delete s;
  return SEMB_OK;
}

static int stokes_create_impl(semb_stokes* s) {
  semb_mesh *V = s->V, *P = s->P;
  auto nodes = [](int n, std::vector<double>& z) {
    std::vector<double> wts(n);
    z.resize(n);
    return semb_gausslobatto(n, z.data(), wts.data());
  };
  std::vector<double> zrV, zsV, zrP, zsP;
  SEMB_TRY(nodes(V->nr, zrV));
  SEMB_TRY(nodes(V->ns, zsV));
  SEMB_TRY(nodes(P->nr, zrP));
  SEMB_TRY(nodes(P->ns, zsP));
  std::vector<double> Jr((size_t)V->nr * P->nr), Js((size_t)V->ns * P->ns), JrT(Jr.size()), JsT(Js.size());
  SEMB_TRY(semb_interp_mat(V->nr, zrV.data(), P->nr, zrP.data(), Jr.data()));
  SEMB_TRY(semb_interp_mat(V->ns, zsV.data(), P->ns, zsP.data(), Js.data()));
  for (int i = 0; i < V->nr; ++i)
    for (int k = 0; k < P->nr; ++k) JrT[k + (size_t)i * P->nr] = Jr[i + (size_t)k * V->nr];
  for (int i = 0; i < V->ns; ++i)
    for (int k = 0; k < P->ns; ++k) JsT[k + (size_t)i * P->ns] = Js[i + (size_t)k * V->ns];
  auto up = [&](const std::vector<double>& h, double** d) -> int {
    SEMB_CHECK_CUDA(cudaMalloc(d, h.size() * sizeof(double)));
    SEMB_CHECK_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return SEMB_OK;
  };
  SEMB_TRY(up(Jr, &s->dJr));
  SEMB_TRY(up(Js, &s->dJs));
  SEMB_TRY(up(JrT, &s->dJrT));
  SEMB_TRY(up(JsT, &s->dJsT));
  const size_t nmid = std::max((size_t)V->pitch * P->nyl, (size_t)P->pitch * V->nyl);
  SEMB_CHECK_CUDA(cudaMalloc(&s->mid, nmid * sizeof(double)));
  if (!V->w_t1) SEMB_TRY(semb_field_create(V, &V->w_t1));  // work field of approxHlmzInv: no allocation while a graph is captured
  semb_field** fv[] = {&s->v1, &s->v2, &s->v3, &s->v4};
  for (semb_field** f : fv) SEMB_TRY(semb_field_create(V, f));
  semb_field** fp[] = {&s->p_Au, &s->p_rhs, &s->p_dp};
  for (semb_field** f : fp) SEMB_TRY(semb_field_create(P, f));
  return SEMB_OK;
}

extern "C" int semb_stokes_create(semb_mesh* mV, semb_mesh* mP, const char bcVX[4], const char bcVY[4], double b0,
                                  semb_stokes** out) {
  SEMB_REQUIRE(mV && mP && out, "semb_stokes_create: null argument");
  *out = nullptr;
  SEMB_ENTER(mV->ctx);
  SEMB_REQUIRE(mP->ctx == mV->ctx && mP->Ex == mV->Ex && mP->Ey == mV->Ey && mP->ney == mV->ney && mP->perx == mV->perx &&
                   mP->pery == mV->pery,
               "Stokes: mshP must match mshV in Ex, Ey, periodicity and partition");
  SEMB_REQUIRE(mV->arr[SEMB_RX] && mV->arr[SEMB_B] && mV->arr[SEMB_BI], "Stokes: mshV has no metric terms / B / Bi");
  SEMB_REQUIRE(b0 != 0.0, "Stokes: b0 must be non-zero");
  MaskFlags f;
  SEMB_TRY(parse_bc(mV, bcVX, &f));
  SEMB_TRY(parse_bc(mV, bcVY, &f));
  semb_stokes* s = new semb_stokes();
  s->V = mV;
  s->P = mP;
  s->b0 = b0;
  if (bcVX) memcpy(s->bcx, bcVX, 4);
  if (bcVY) memcpy(s->bcy, bcVY, 4);
  const int rc = stokes_create_impl(s);
  if (rc < 0) {
    semb_stokes_destroy(s);
    return rc;
  }
  *out = s;
  return SEMB_OK;
}

// ABu(Js,Jr,p): pressure grid -> velocity grid (Br = Jr first, then As = Js, ABu.jl:14-33)
static int stokes_interp_PV(semb_stokes* s, const double* p, double* outV) {
  semb_mesh *V = s->V, *P = s->P;
  semb_ctx* c = V->ctx;
  SEMB_TRY(semb_launch_abu_r(c, s->dJr, V->nr, P->nr, p, P->nxl, P->nyl, P->pitch, s->mid, V->pitch));
  return semb_launch_abu_s(c, s->dJs, V->ns, P->ns, s->mid, V->nxl, P->nyl, V->pitch, outV, V->pitch);
}
// ABu(Js',Jr',v): velocity grid -> pressure grid
static int stokes_interp_VP(semb_stokes* s, const double* v, double* outP) {
  semb_mesh *V = s->V, *P = s->P;
  semb_ctx* c = V->ctx;
  SEMB_TRY(semb_launch_abu_r(c, s->dJrT, P->nr, V->nr, v, V->nxl, V->nyl, V->pitch, s->mid, P->pitch));
  return semb_launch_abu_s(c, s->dJsT, P->ns, V->ns, s->mid, P->nxl, V->nyl, P->pitch, outP, P->pitch);
}

// diver(ux,uy,mshV,Jr,Js), diver.jl:17-31, times `sign`: one register-tiled launch when (nr, nr-2) is served
// (semb_stokes_tile.cu), else the generic chain
static int stokes_diver(semb_stokes* s, const semb_field* ux, const semb_field* uy, semb_field* out, double sign) {
  SEMB_TRY(check_field(s->V, ux, "diver(ux)"));
  SEMB_TRY(check_field(s->V, uy, "diver(uy)"));
  SEMB_TRY(check_field(s->P, out, "diver(out)"));
  SEMB_REQUIRE(ux != s->v4 && uy != s->v4, "diver: argument aliases the work field");
  semb_ctx* c = s->V->ctx;
  int done = 0;
  if (!getenv("SEMB_NO_TILED_STOKES"))
    SEMB_TRY(semb_launch_stokes_tile(c, s->V, s->P, 0, ux->d, uy->d, out->d, nullptr, s->dJr, s->dJs, sign, &done));
  if (done) return SEMB_OK;
  SEMB_TRY(semb_launch_diver_local(c, s->V, ux->d, uy->d, s->v4->d));  // B .* (uxdx + uydy), :22-27
  SEMB_TRY(stokes_interp_VP(s, s->v4->d, out->d));                      // ABu(Js',Jr',Bdiv), :28
  return sign == 1.0 ? SEMB_OK : semb_field_axpby(0.0, out, sign, out);
}

extern "C" int semb_diver(semb_stokes* s, const semb_field* ux, const semb_field* uy, semb_field* out) {
  SEMB_REQUIRE(s, "null Stokes handle");
  SEMB_ENTER(s->V->ctx);
  return stokes_diver(s, ux, uy, out, 1.0);
}

// diverT(pr,mshV,Jr,Js), diver.jl:53-63
extern "C" int semb_diverT(semb_stokes* s, const semb_field* pr, semb_field* qx, semb_field* qy) {
  SEMB_REQUIRE(s, "null Stokes handle");
  SEMB_ENTER(s->V->ctx);
  SEMB_TRY(check_field(s->P, pr, "diverT(pr)"));
  SEMB_TRY(check_field(s->V, qx, "diverT(qx)"));
  SEMB_TRY(check_field(s->V, qy, "diverT(qy)"));
  SEMB_REQUIRE(qx != qy && qx != s->v4 && qy != s->v4, "diverT: outputs must not alias (each other or the work field)");
  int done = 0;
  if (!getenv("SEMB_NO_TILED_STOKES"))
    SEMB_TRY(semb_launch_stokes_tile(s->V->ctx, s->V, s->P, 1, pr->d, nullptr, qx->d, qy->d, s->dJr, s->dJs, 1.0, &done));
  if (done) return SEMB_OK;
  SEMB_TRY(stokes_interp_PV(s, pr->d, s->v4->d));                                                   // Jp, :57
  return semb_launch_gradT(s->V->ctx, s->V, s->v4->d, s->V->arr[SEMB_B], qx->d, qy->d);            // mass, gradT, :58-60
}

// opStokesLHS(q,sks), stokes.jl:110-121 = gatherScatter(stokesOp(q,...), mshP), stokesOp: diver.jl:73-89
extern "C" int semb_stokes_op(semb_stokes* s, const semb_field* q, semb_field* out) {
  SEMB_REQUIRE(s, "null Stokes handle");
  SEMB_ENTER(s->V->ctx);
  SEMB_TRY(check_field(s->P, q, "stokesOp(q)"));
  SEMB_TRY(check_field(s->P, out, "stokesOp(out)"));
  SEMB_REQUIRE(q != out && q != s->p_rhs && out != s->p_rhs, "stokesOp: argument aliases a work field");
  SEMB_TRY(semb_diverT(s, q, s->v1, s->v2));                                  // DD'
  SEMB_TRY(semb_approx_hlmz_inv(s->V, s->v1, s->b0, s->bcx, s->v3));          // HH^-1
  SEMB_TRY(semb_approx_hlmz_inv(s->V, s->v2, s->b0, s->bcy, s->v1));
  SEMB_TRY(stokes_diver(s, s->v3, s->v1, s->p_rhs, -1.0));                    // DD, and "return -Eq" (diver.jl:88)
  if (s->P->fast) return gs_one_pass(s->P, s->p_rhs->d, out->d, 0, 1.0, MaskFlags());  // stokes.jl:118, one pass
  return semb_gather_scatter(s->P, s->p_rhs, out);
}

// makeStokesRHS!, stokes.jl:128-141, for the velocity (vx, vy) to be projected
extern "C" int semb_stokes_rhs(semb_stokes* s, const semb_field* vx, const semb_field* vy, semb_field* rhs) {
  SEMB_REQUIRE(s, "null Stokes handle");
  SEMB_ENTER(s->V->ctx);
  SEMB_TRY(check_field(s->P, rhs, "makeStokesRHS(rhs)"));
  SEMB_REQUIRE(rhs != s->p_Au, "makeStokesRHS: rhs aliases a work field");
  SEMB_TRY(semb_diver(s, vx, vy, s->p_Au));
  return semb_gather_scatter(s->P, s->p_Au, rhs);
}

// solveStokes!, stokes.jl:143-154: pcg(rhs, opStokesLHS; mult = mshP.mult, opM = identity (stokes.jl:123-126)) on the
// device-resident PCG of the pressure mesh (pcg.jl:16-60: scalars, convergence flag and iteration count live in device
// memory; CUDA-graph replay on one rank) with the Schur operator plugged in as its operator hook
extern "C" int semb_stokes_solve(semb_stokes* s, const semb_field* rhs, semb_field* dp, double tol, long long maxiter,
                                 long long* iters, double* resinf) {
  SEMB_REQUIRE(s, "null Stokes handle");
  semb_mesh* P = s->P;
  SEMB_ENTER(P->ctx);
  SEMB_TRY(check_field(P, rhs, "solveStokes(rhs)"));
  SEMB_TRY(check_field(P, dp, "solveStokes(dp)"));
  SEMB_REQUIRE(rhs != dp && rhs != s->p_rhs && dp != s->p_rhs, "solveStokes: rhs / dp alias each other or a work field");
  SEMB_TRY(ensure_tmp(P, &P->w_p));
  SEMB_TRY(ensure_tmp(P, &P->w_Ap));
  semb_pcg_opts o;
  memset(&o, 0, sizeof(o));
  o.nu = 1.0;
  o.tol = tol;
  o.maxiter = maxiter;
  P->pcg_custom = [s, P]() -> int { return semb_stokes_op(s, P->w_p, P->w_Ap); };
  const int rc = semb_pcg(P, &o, rhs, dp, iters, resinf);
  P->pcg_custom = nullptr;
  return rc;
}

// pressureProject!, stokes.jl:159-177: vx, vy, pr are updated in place
extern "C" int semb_stokes_project(semb_stokes* s, semb_field* vx, semb_field* vy, semb_field* pr, double tol,
                                   long long maxiter, long long* iters, double* resinf) {
  SEMB_REQUIRE(s, "null Stokes handle");
  SEMB_ENTER(s->V->ctx);
  SEMB_TRY(check_field(s->V, vx, "pressureProject(vx)"));
  SEMB_TRY(check_field(s->V, vy, "pressureProject(vy)"));
  SEMB_TRY(check_field(s->P, pr, "pressureProject(pr)", true));
  SEMB_TRY(semb_stokes_rhs(s, vx, vy, s->p_rhs));                          // makeStokesRHS!, :162
  // p_rhs is a work field of the operator: the right-hand side lives in its own field, kept on the handle (a
  // cudaMalloc + cudaFree pair per call cost 1 ms of a 34 ms projection)
  if (!s->p_rhs2) SEMB_TRY(semb_field_create(s->P, &s->p_rhs2));
  SEMB_TRY(semb_field_copy(s->p_rhs2, s->p_rhs));
  const int rc = semb_stokes_solve(s, s->p_rhs2, s->p_dp, tol, maxiter, iters, resinf);  // :164
  if (rc < 0) return rc;
  SEMB_TRY(semb_diverT(s, s->p_dp, s->v1, s->v2));                         // :166
  SEMB_TRY(semb_approx_hlmz_inv(s->V, s->v1, s->b0, s->bcx, s->v3));       // :169
  SEMB_TRY(semb_field_axpby(1.0, s->v3, 1.0, vx));                         // :173
  SEMB_TRY(semb_approx_hlmz_inv(s->V, s->v2, s->b0, s->bcy, s->v3));       // :170
  SEMB_TRY(semb_field_axpby(1.0, s->v3, 1.0, vy));                         // :174
  if (pr) SEMB_TRY(semb_field_axpby(1.0, s->p_dp, 1.0, pr));               // :175
  return rc;
}
