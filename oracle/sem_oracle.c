/* CPU ORACLE, second restatement (test infrastructure, NOT product code) -- plain C loops.
 *
 * Independent of oracle/sem_oracle.py (which keeps the reference's GEMM-per-ABu structure in NumPy): the same
 * functions of vpuri3/SpectralElements.jl's matrix-free path written as explicit loops, so that two restatements
 * with different summation structure can be checked against each other (tests/test_oracle_c.py) and a compiled,
 * threaded CPU figure exists next to the NumPy one.  Only tests/ may load the library built from this file.
 *
 * PARITY UNPINNED (as stated in sem_oracle.py): Julia is not installed, the reference ships no golden vectors for this
 * path, and its third-party arithmetic (FastGaussQuadrature, OpenBLAS, Base.sum) is not version-pinned.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference/src).  Arrays are
 * column-major (nxl x nyl), nxl = nr*Ex the contiguous direction, as in the reference (mesh.jl:94).
 * Threading: OpenMP over element columns / rows; every output entry is produced by one thread in a fixed order, so
 * results do not depend on the thread count.  Reductions are pairwise in a fixed order (so_sum3).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads the element loops run on (1 without OpenMP): bench.py reports it as cpu_baseline.cores */
int so_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define SO_PI 3.14159265358979323846

/* ---- FastGaussQuadrature.gausslobatto restated (call sites mesh.jl:70-71, semmesh.jl:11) --------------------- */
static void legendre(int n, double x, double* pn, double* pnm1) {
  double p0 = 1.0, p1 = x;
  if (n == 0) { *pn = 1.0; *pnm1 = 0.0; return; }
  for (int k = 2; k <= n; ++k) {
    const double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1;
    p1 = p2;
  }
  *pn = p1;
  *pnm1 = p0;
}

/* nodes = roots of (1-x^2) P'_{n-1}(x) (Newton from the Chebyshev-Gauss-Lobatto points), w = 2/(n(n-1) P_{n-1}(x)^2) */
int so_gausslobatto(int n, double* z, double* w) {
  if (n < 2) return -1;
  const int N = n - 1;
  for (int i = 0; i < n; ++i) z[i] = -cos(SO_PI * i / N);
  for (int it = 0; it < 100; ++it) {
    double dmax = 0.0;
    for (int i = 1; i < n - 1; ++i) {
      double pn, pm;
      legendre(N, z[i], &pn, &pm);
      const double q = N * (pm - z[i] * pn), dq = -(double)N * (N + 1) * pn;
      const double dx = q / dq;
      z[i] -= dx;
      if (fabs(dx) > dmax) dmax = fabs(dx);
    }
    if (dmax < 1e-16) break;
  }
  z[0] = -1.0;
  z[n - 1] = 1.0;
  for (int i = 0; i < n / 2; ++i) { /* antisymmetry (exact 0 at the centre for odd n) */
    const double a = 0.5 * (z[i] - z[n - 1 - i]);
    z[i] = a;
    z[n - 1 - i] = -a;
  }
  if (n % 2) z[n / 2] = 0.0;
  for (int i = 0; i < n; ++i) {
    double pn, pm;
    legendre(N, z[i], &pn, &pm);
    w[i] = 2.0 / ((double)N * (N + 1) * pn * pn);
  }
  return 0;
}

/* derivMat.jl:9-35: barycentric weights a_i = 1/prod_{j != i}(x_i - x_j); D_ij = a_j / (a_i (x_i - x_j)); the diagonal
 * is the row sum of 1/(x_i - x_j) (derivMat.jl:24-27).  D is n x n column-major. */
void so_derivmat(int n, const double* x, double* D) {
  double* a = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) {
    double p = 1.0;
    for (int j = 0; j < n; ++j)
      if (j != i) p *= (x[i] - x[j]);
    a[i] = 1.0 / p;
  }
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j)
      if (j != i) s += 1.0 / (x[i] - x[j]);
    D[i + i * n] = s;
    for (int j = 0; j < n; ++j)
      if (j != i) D[i + j * n] = a[j] / (a[i] * (x[i] - x[j]));
  }
  free(a);
}

/* interp.jl:10-35: J (no x ni, column-major), J_ij = a_j prod_{k<j}(xo_i - xi_k) prod_{k>j}(xo_i - xi_k) */
void so_interpmat(int no, const double* xo, int ni, const double* xi, double* J) {
  double* a = (double*)malloc(sizeof(double) * ni);
  double* s = (double*)malloc(sizeof(double) * ni);
  double* t = (double*)malloc(sizeof(double) * ni);
  for (int i = 0; i < ni; ++i) {
    double p = 1.0;
    for (int j = 0; j < ni; ++j)
      if (j != i) p *= (xi[i] - xi[j]);
    a[i] = 1.0 / p;
  }
  for (int i = 0; i < no; ++i) {
    const double x = xo[i];
    s[0] = 1.0;
    t[ni - 1] = 1.0;
    for (int j = 1; j < ni; ++j) { /* interp.jl:27-30 */
      s[j] = s[j - 1] * (x - xi[j - 1]);
      t[ni - 1 - j] = t[ni - j] * (x - xi[ni - j]);
    }
    for (int j = 0; j < ni; ++j) J[i + j * no] = a[j] * s[j] * t[j];
  }
  free(a);
  free(s);
  free(t);
}

/* semmesh.jl:9-27: E uniform elements on [-1,1], the n GLL points of each placed affinely */
void so_semmesh(int E, int n, double* z) {
  double* z0 = (double*)malloc(sizeof(double) * n);
  double* w0 = (double*)malloc(sizeof(double) * n);
  so_gausslobatto(n, z0, w0);
  for (int e = 0; e < E; ++e) {
    const double lo = -1.0 + 2.0 * e / E, hi = -1.0 + 2.0 * (e + 1) / E;
    for (int i = 0; i < n; ++i) z[e * n + i] = (hi - lo) * (0.5 * (z0[i] + 1.0)) + lo;
  }
  free(z0);
  free(w0);
}

/* ABu.jl:9-37: out = (As (x)_blk Br) u.  Br (mb x nb) acts on each consecutive nb-row chunk of every column, As (ma x na)
 * on each consecutive na-column chunk (u[:,jj] * As').  A NULL matrix is Julia's `[]` (identity).  u is m x n; out is
 * (m*mb/nb) x (n*ma/na).  Returns -1 where the reference raises InexactError (ABu.jl:16,26). */
int so_abu(const double* As, int ma, int na, const double* Br, int mb, int nb, const double* u, int m, int n,
           double* out) {
  int mo = m;
  if (Br) {
    if (m % nb) return -1;
    mo = m / nb * mb;
  }
  if (As) {
    if (n % na) return -1;
  }
  double* tmp = (double*)malloc(sizeof(double) * (size_t)mo * n);
  if (Br) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j)
      for (int b = 0; b < m / nb; ++b)
        for (int i = 0; i < mb; ++i) {
          double s = 0.0;
          for (int k = 0; k < nb; ++k) s += Br[i + k * mb] * u[(size_t)j * m + b * nb + k];
          tmp[(size_t)j * mo + b * mb + i] = s;
        }
  } else {
    memcpy(tmp, u, sizeof(double) * (size_t)m * n);
  }
  if (As) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < n / na; ++b)
      for (int jo = 0; jo < ma; ++jo)
        for (int i = 0; i < mo; ++i) {
          double s = 0.0;
          for (int k = 0; k < na; ++k) s += tmp[(size_t)(b * na + k) * mo + i] * As[jo + k * ma];
          out[(size_t)(b * ma + jo) * mo + i] = s;
        }
  } else {
    memcpy(out, tmp, sizeof(double) * (size_t)mo * n);
  }
  free(tmp);
  return 0;
}

/* ---- Mesh, mesh.jl:25-133 ------------------------------------------------------------------------------------- */
typedef struct so_mesh {
  int nr, ns, Ex, Ey, perx, pery, nxl, nyl;
  double *zr, *zs, *wr, *ws, *Dr, *Ds;
  double *x, *y, *Jac, *Jaci, *rx, *ry, *sx, *sy, *B, *Bi, *G11, *G12, *G22, *mult;
} so_mesh;

enum { SO_FIXU = 0, SO_ANNULUS = 1, SO_WAVY = 2 };

static double* dalloc(size_t n) { return (double*)calloc(n, sizeof(double)); }

void so_gather_scatter(const so_mesh* m, const double* u, double* out);

/* out = Dr-derivative along r (d = 0) or Ds-derivative along s (d = 1) of a field, element by element
 * (ABu([],Dr,u) / ABu(Ds,[],u), jac.jl:30-33, lapl.jl:72-73); T != 0 applies the transposed matrix (lapl.jl:78) */
static void deriv(const so_mesh* m, const double* u, int d, int T, double* out) {
  const int nr = m->nr, ns = m->ns, nxl = m->nxl;
#pragma omp parallel for schedule(static)
  for (int ey = 0; ey < m->Ey; ++ey)
    for (int ex = 0; ex < m->Ex; ++ex)
      for (int j = 0; j < ns; ++j)
        for (int i = 0; i < nr; ++i) {
          double s = 0.0;
          if (d == 0) {
            const double* row = u + (size_t)(ey * ns + j) * nxl + ex * nr;
            for (int k = 0; k < nr; ++k) s += (T ? m->Dr[k + i * nr] : m->Dr[i + k * nr]) * row[k];
          } else {
            const double* col = u + (size_t)(ey * ns) * nxl + ex * nr + i;
            for (int k = 0; k < ns; ++k) s += (T ? m->Ds[k + j * ns] : m->Ds[j + k * ns]) * col[(size_t)k * nxl];
          }
          out[(size_t)(ey * ns + j) * nxl + ex * nr + i] = s;
        }
}

void so_mesh_free(so_mesh* m) {
  if (!m) return;
  double* all[] = {m->zr, m->zs, m->wr, m->ws, m->Dr, m->Ds, m->x, m->y, m->Jac, m->Jaci, m->rx, m->ry,
                   m->sx, m->sy, m->B, m->Bi, m->G11, m->G12, m->G22, m->mult};
  for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) free(all[i]);
  free(m);
}

/* Mesh(nr,ns,Ex,Ey,ifperiodic,deform), mesh.jl:66-133 (mesh.jl:80 builds Qy with Ex: Ey is used here, flagged in
 * sem_oracle.py as well).  deform: SO_FIXU (mesh.jl:6-8), SO_ANNULUS (geom.jl:40-49), SO_WAVY (bench mesh). */
so_mesh* so_mesh_create(int nr, int ns, int Ex, int Ey, int perx, int pery, int deform) {
  so_mesh* m = (so_mesh*)calloc(1, sizeof(so_mesh));
  m->nr = nr; m->ns = ns; m->Ex = Ex; m->Ey = Ey; m->perx = perx; m->pery = pery;
  const int nxl = m->nxl = nr * Ex, nyl = m->nyl = ns * Ey;
  const size_t n = (size_t)nxl * nyl;
  m->zr = dalloc(nr); m->wr = dalloc(nr); m->zs = dalloc(ns); m->ws = dalloc(ns);
  m->Dr = dalloc((size_t)nr * nr); m->Ds = dalloc((size_t)ns * ns);
  so_gausslobatto(nr, m->zr, m->wr); /* mesh.jl:70-71 */
  so_gausslobatto(ns, m->zs, m->ws);
  so_derivmat(nr, m->zr, m->Dr);     /* mesh.jl:73-74 */
  so_derivmat(ns, m->zs, m->Ds);
  double** f[] = {&m->x, &m->y, &m->Jac, &m->Jaci, &m->rx, &m->ry, &m->sx, &m->sy, &m->B, &m->Bi,
                  &m->G11, &m->G12, &m->G22, &m->mult};
  for (size_t i = 0; i < sizeof(f) / sizeof(f[0]); ++i) *f[i] = dalloc(n);
  /* grid: semmesh + ndgrid (mesh.jl:98-100), then the deformation (mesh.jl:108) */
  double* xe = dalloc(nxl);
  double* ye = dalloc(nyl);
  so_semmesh(Ex, nr, xe);
  so_semmesh(Ey, ns, ye);
  for (int j = 0; j < nyl; ++j)
    for (int i = 0; i < nxl; ++i) {
      const double r = xe[i], s = ye[j];
      double X = r, Y = s;
      if (deform == SO_ANNULUS) { /* geom.jl:40-49, r0 = 0.5, r1 = 1, span = 2 pi */
        const double R = (1.0 - 0.5) / 2 * (r + 1) + 0.5, th = (2 * SO_PI) / 2 * (s + 1) + 0.0;
        X = R * cos(th);
        Y = R * sin(th);
      } else if (deform == SO_WAVY) {
        const double d = 0.1 * sin(SO_PI * r) * sin(SO_PI * s);
        X = r + d;
        Y = s + d;
      }
      m->x[(size_t)j * nxl + i] = X;
      m->y[(size_t)j * nxl + i] = Y;
    }
  free(xe);
  free(ye);
  /* jac.jl:24-40 */
  double *xr = dalloc(n), *xs = dalloc(n), *yr = dalloc(n), *ys = dalloc(n);
  deriv(m, m->x, 0, 0, xr);
  deriv(m, m->x, 1, 0, xs);
  deriv(m, m->y, 0, 0, yr);
  deriv(m, m->y, 1, 0, ys);
  for (size_t q = 0; q < n; ++q) {
    const double J = xr[q] * ys[q] - xs[q] * yr[q], Ji = 1.0 / J;
    m->Jac[q] = J;
    m->Jaci[q] = Ji;
    m->rx[q] = Ji * ys[q];
    m->ry[q] = -Ji * xs[q];
    m->sx[q] = -Ji * yr[q];
    m->sy[q] = Ji * xr[q];
  }
  free(xr); free(xs); free(yr); free(ys);
  /* mesh.jl:114-123: B = Jac .* (wx*wy') with the reference-element weights tiled per element; G11, G12, G22 */
  for (int j = 0; j < nyl; ++j)
    for (int i = 0; i < nxl; ++i) {
      const size_t q = (size_t)j * nxl + i;
      const double B = m->Jac[q] * (m->wr[i % nr] * m->ws[j % ns]);
      const double rx = m->rx[q], ry = m->ry[q], sx = m->sx[q], sy = m->sy[q];
      m->B[q] = B;
      m->Bi[q] = 1.0 / B;
      m->G11[q] = B * (rx * rx + ry * ry);
      m->G12[q] = B * (rx * sx + ry * sy);
      m->G22[q] = B * (sx * sx + sy * sy);
    }
  /* mesh.jl:94-96: mult = 1 ./ gatherScatter(ones) */
  double* ones = dalloc(n);
  for (size_t q = 0; q < n; ++q) ones[q] = 1.0;
  so_gather_scatter(m, ones, m->mult);
  for (size_t q = 0; q < n; ++q) m->mult[q] = 1.0 / m->mult[q];
  free(ones);
  return m;
}

/* which: 0 x, 1 y, 2 Jac, 3 Jaci, 4 rx, 5 ry, 6 sx, 7 sy, 8 B, 9 Bi, 10 G11, 11 G12, 12 G22, 13 mult; 20 Dr, 21 Ds,
 * 22 zr, 23 wr, 24 zs, 25 ws */
const double* so_mesh_array(const so_mesh* m, int which) {
  const double* a[] = {m->x, m->y, m->Jac, m->Jaci, m->rx, m->ry, m->sx, m->sy, m->B, m->Bi, m->G11, m->G12, m->G22,
                       m->mult};
  if (which >= 0 && which < 14) return a[which];
  switch (which) {
    case 20: return m->Dr;
    case 21: return m->Ds;
    case 22: return m->zr;
    case 23: return m->wr;
    case 24: return m->zs;
    case 25: return m->ws;
  }
  return NULL;
}

/* ---- operators ------------------------------------------------------------------------------------------------- */
/* laplace(u,Dr,Ds,G11,G12,G22), lapl.jl:70-81 (no gather-scatter, no mask) */
void so_laplace(const so_mesh* m, const double* u, double* out) {
  const size_t n = (size_t)m->nxl * m->nyl;
  double *ur = dalloc(n), *us = dalloc(n), *wr = dalloc(n), *ws = dalloc(n);
  deriv(m, u, 0, 0, ur);                                    /* lapl.jl:72 */
  deriv(m, u, 1, 0, us);                                    /* lapl.jl:73 */
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < (long long)n; ++q) {
    wr[q] = m->G11[q] * ur[q] + m->G12[q] * us[q];           /* lapl.jl:75 */
    ws[q] = m->G12[q] * ur[q] + m->G22[q] * us[q];           /* lapl.jl:76 */
  }
  deriv(m, wr, 0, 1, ur);                                   /* lapl.jl:78: ABu([],Dr',wr) */
  deriv(m, ws, 1, 1, us);                                   /*             ABu(Ds',[],ws) */
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < (long long)n; ++q) out[q] = ur[q] + us[q];
  free(ur); free(us); free(wr); free(ws);
}

/* mass(u,msh), mass.jl:12-22 */
void so_mass(const so_mesh* m, const double* u, double* out) {
  const size_t n = (size_t)m->nxl * m->nyl;
  for (size_t q = 0; q < n; ++q) out[q] = m->B[q] * u[q];
}

/* hlmz(u,nu,k,msh), hlmz.jl:12-19: Hu = nu .* lapl(u); Hu .+= k .* mass(u); nu_arr / k_arr may be NULL (scalar) */
void so_hlmz(const so_mesh* m, const double* u, const double* nu_arr, double nu, const double* k_arr, double k,
             double* out) {
  const size_t n = (size_t)m->nxl * m->nyl;
  so_laplace(m, u, out);
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < (long long)n; ++q) {
    const double hu = (nu_arr ? nu_arr[q] : nu) * out[q];
    out[q] = hu + (k_arr ? k_arr[q] : k) * (m->B[q] * u[q]);
  }
}

/* gatherScatter(u,msh) = QQtx*u*QQty', gatherScatter.jl:8-21: every interface node receives the sum of its duplicates,
 * x pairs first (Br = QQtx is applied before As = QQty, ABu.jl:14-33), then y pairs; semq.jl:19-22 for the periodic wrap */
void so_gather_scatter(const so_mesh* m, const double* u, double* out) {
  const int nxl = m->nxl, nyl = m->nyl, nr = m->nr, ns = m->ns;
  if (out != u) memcpy(out, u, sizeof(double) * (size_t)nxl * nyl);
  for (int j = 0; j < nyl; ++j) {
    double* row = out + (size_t)j * nxl;
    for (int e = 1; e < m->Ex; ++e) {
      const double s = row[e * nr - 1] + row[e * nr];
      row[e * nr - 1] = s;
      row[e * nr] = s;
    }
    if (m->perx) {
      const double s = row[nxl - 1] + row[0];
      row[0] = s;
      row[nxl - 1] = s;
    }
  }
  for (int e = 1; e < m->Ey; ++e) {
    double *a = out + (size_t)(e * ns - 1) * nxl, *b = out + (size_t)(e * ns) * nxl;
    for (int i = 0; i < nxl; ++i) {
      const double s = a[i] + b[i];
      a[i] = s;
      b[i] = s;
    }
  }
  if (m->pery) {
    double *a = out + (size_t)(nyl - 1) * nxl, *b = out;
    for (int i = 0; i < nxl; ++i) {
      const double s = a[i] + b[i];
      a[i] = s;
      b[i] = s;
    }
  }
}

/* generateMask(bc,msh), mesh.jl:149-175: bc = [xmin,xmax,ymin,ymax], 'D' zeroes the line, a periodic direction is all ones */
void so_generate_mask(const so_mesh* m, const char* bc, double* M) {
  const int nxl = m->nxl, nyl = m->nyl;
  for (int j = 0; j < nyl; ++j)
    for (int i = 0; i < nxl; ++i) {
      double mx = 1.0, my = 1.0;
      if (!m->perx && ((i == 0 && bc[0] == 'D') || (i == nxl - 1 && bc[1] == 'D'))) mx = 0.0;
      if (!m->pery && ((j == 0 && bc[2] == 'D') || (j == nyl - 1 && bc[3] == 'D'))) my = 0.0;
      M[(size_t)j * nxl + i] = mx * my;
    }
}

/* opLHS(u,dfn) = mask(gatherScatter(hlmz(u,nu,bdfB[1],msh)),M), diffusion.jl:36-45: gs THEN mask; M may be NULL (mask.jl:13) */
void so_oplhs(const so_mesh* m, const double* u, const double* nu_arr, double nu, const double* k_arr, double k,
              const double* M, double* out) {
  const size_t n = (size_t)m->nxl * m->nyl;
  so_hlmz(m, u, nu_arr, nu, k_arr, k, out);
  so_gather_scatter(m, out, out);
  if (M)
    for (size_t q = 0; q < n; ++q) out[q] = M[q] * out[q];
}

/* sum(a .* b .* c) with the product order (a*b)*c of pcg.jl:45,49,52 and a fixed pairwise tree (Base.sum is pairwise too,
 * with an unpinned block size: bitwise agreement with Julia is not claimed) */
static double so_sum3(const double* a, const double* b, const double* c, size_t lo, size_t hi) {
  if (hi - lo <= 1024) {
    double s = 0.0;
    for (size_t q = lo; q < hi; ++q) s += (a[q] * b[q]) * c[q];
    return s;
  }
  const size_t mid = lo + (hi - lo) / 2;
  return so_sum3(a, b, c, lo, mid) + so_sum3(a, b, c, mid, hi);
}

/* pcg(b,opA;opM,mult,tol,maxiter), pcg.jl:16-60, with opA = opLHS above and opM = identity (prec_b0 == 0) or
 * u ./ B ./ b0 (convectionDiffusion.jl:87-91).  x0 = 0 (pcg.jl:25); stop when norm(r,Inf) <= tol at the loop top
 * (pcg.jl:36); at k == maxiter the iterate is returned with warned = 1 (pcg.jl:39).  Returns the iteration count. */
long long so_pcg(const so_mesh* m, const double* b, const double* nu_arr, double nu, const double* k_arr, double k,
                 const double* M, double prec_b0, double tol, long long maxiter, double* x, double* resinf, int* warned) {
  const size_t n = (size_t)m->nxl * m->nyl;
  double *r = dalloc(n), *h = dalloc(n), *u = dalloc(n), *Au = dalloc(n);
  if (maxiter < 0) maxiter = (long long)n;
  memset(x, 0, sizeof(double) * n);
  memcpy(r, b, sizeof(double) * n);
  long long it = 0;
  double t_prev = 0.0, rinf = 0.0;
  *warned = 0;
  for (;;) {
    rinf = 0.0;
    for (size_t q = 0; q < n; ++q)
      if (fabs(r[q]) > rinf) rinf = fabs(r[q]);
    if (!(rinf > tol)) break;                                        /* pcg.jl:36 */
    for (size_t q = 0; q < n; ++q) h[q] = prec_b0 != 0.0 ? (r[q] / m->B[q]) / prec_b0 : r[q]; /* pcg.jl:37 */
    if (it == maxiter) { *warned = 1; break; }                       /* pcg.jl:39 */
    ++it;
    const double t = so_sum3(r, h, m->mult, 0, n);                   /* pcg.jl:45 */
    if (it == 1) {
      memcpy(u, h, sizeof(double) * n);
    } else {
      const double beta = t / t_prev;                                /* pcg.jl:49 (t_prev is recomputed there: same value) */
      for (size_t q = 0; q < n; ++q) u[q] = h[q] + beta * u[q];
    }
    so_oplhs(m, u, nu_arr, nu, k_arr, k, M, Au);                     /* pcg.jl:51 */
    const double alpha = t / so_sum3(u, Au, m->mult, 0, n);          /* pcg.jl:52 */
    for (size_t q = 0; q < n; ++q) {
      x[q] = x[q] + alpha * u[q];                                    /* pcg.jl:53 */
      r[q] = r[q] - alpha * Au[q];                                   /* pcg.jl:54 */
    }
    t_prev = t;
  }
  *resinf = rinf;
  free(r); free(h); free(u); free(Au);
  return it;
}

/* ---- explicit dealiased convection --------------------------------------------------------------- */
/* grad(u,msh), grad.jl:15-34: ux = rx.*ur + sx.*us, uy = ry.*ur + sy.*us */
void so_grad(const so_mesh* m, const double* u, double* ux, double* uy) {
  const size_t n = (size_t)m->nxl * m->nyl;
  double *ur = dalloc(n), *us = dalloc(n);
  deriv(m, u, 0, 0, ur);
  deriv(m, u, 1, 0, us);
  for (size_t q = 0; q < n; ++q) {
    ux[q] = m->rx[q] * ur[q] + m->sx[q] * us[q]; /* grad.jl:30 */
    uy[q] = m->ry[q] * ur[q] + m->sy[q] * us[q]; /* grad.jl:31 */
  }
  free(ur);
  free(us);
}

/* one element's tile through J (no x ni, column-major) along r and s: out (no_r x no_s) = Jr * tile * Js', i.e. the
 * element block of ABu(Js,Jr,.) (ABu.jl:14-33: Br = Jr first, then As = Js); T != 0 applies the transposes */
static void tile_interp(const double* Jr, int mr, int nr_, const double* Js, int ms, int ns_, int T, const double* in,
                        size_t ldin, double* out, size_t ldout, double* tmp) {
  const int ir = T ? mr : nr_, or_ = T ? nr_ : mr, is = T ? ms : ns_, os = T ? ns_ : ms;
  for (int j = 0; j < is; ++j)
    for (int i = 0; i < or_; ++i) {
      double s = 0.0;
      for (int k = 0; k < ir; ++k) s += (T ? Jr[k + i * mr] : Jr[i + k * mr]) * in[(size_t)j * ldin + k];
      tmp[(size_t)j * or_ + i] = s;
    }
  for (int j = 0; j < os; ++j)
    for (int i = 0; i < or_; ++i) {
      double s = 0.0;
      for (int k = 0; k < is; ++k) s += tmp[(size_t)k * or_ + i] * (T ? Js[k + j * ms] : Js[j + k * ms]);
      out[(size_t)j * ldout + i] = s;
    }
}

/* advect(T,ux,uy,mshV,mshD,Jr,Js), advect.jl:45-64 with Jr = interpMat(mshD.zr,mshV.zr) (advect.jl:72-73): gradient on
 * mshV, interpolation of Tx, Ty, ux, uy to the dealiasing mesh, pointwise product with mshD.B, projection back with the
 * transposes.  mD == NULL: the un-dealiased form advect(T,ux,uy,msh), advect.jl:27-43.  Element by element. */
int so_advect(const so_mesh* V, const so_mesh* D, const double* T, const double* ux, const double* uy, double* out) {
  const size_t n = (size_t)V->nxl * V->nyl;
  double *Tx = dalloc(n), *Ty = dalloc(n);
  so_grad(V, T, Tx, Ty);
  if (!D) {
    for (size_t q = 0; q < n; ++q) out[q] = (ux[q] * Tx[q] + uy[q] * Ty[q]) * V->B[q]; /* advect.jl:36-37 */
    free(Tx);
    free(Ty);
    return 0;
  }
  if (D->Ex != V->Ex || D->Ey != V->Ey) {
    free(Tx);
    free(Ty);
    return -1;
  }
  const int nr = V->nr, ns = V->ns, mr = D->nr, ms = D->ns;
  double *Jr = dalloc((size_t)mr * nr), *Js = dalloc((size_t)ms * ns);
  so_interpmat(mr, D->zr, nr, V->zr, Jr);
  so_interpmat(ms, D->zs, ns, V->zs, Js);
#pragma omp parallel for schedule(static)
  for (int ey = 0; ey < V->Ey; ++ey) {
    const size_t td = (size_t)mr * ms;
    double* w = (double*)malloc(sizeof(double) * (5 * td + (size_t)mr * (ns > ms ? ns : ms)));
    double *jtx = w, *jty = w + td, *jux = w + 2 * td, *juy = w + 3 * td, *jc = w + 4 * td, *tmp = w + 5 * td;
    for (int ex = 0; ex < V->Ex; ++ex) {
      const size_t oV = (size_t)(ey * ns) * V->nxl + (size_t)ex * nr, oD = (size_t)(ey * ms) * D->nxl + (size_t)ex * mr;
      tile_interp(Jr, mr, nr, Js, ms, ns, 0, Tx + oV, V->nxl, jtx, mr, tmp); /* advect.jl:54-57 */
      tile_interp(Jr, mr, nr, Js, ms, ns, 0, Ty + oV, V->nxl, jty, mr, tmp);
      tile_interp(Jr, mr, nr, Js, ms, ns, 0, ux + oV, V->nxl, jux, mr, tmp);
      tile_interp(Jr, mr, nr, Js, ms, ns, 0, uy + oV, V->nxl, juy, mr, tmp);
      for (int j = 0; j < ms; ++j)
        for (int i = 0; i < mr; ++i) {
          const size_t q = (size_t)j * mr + i;
          jc[q] = (jux[q] * jtx[q] + juy[q] * jty[q]) * D->B[oD + (size_t)j * D->nxl + i]; /* advect.jl:59-60 */
        }
      tile_interp(Jr, mr, nr, Js, ms, ns, 1, jc, mr, out + oV, V->nxl, tmp); /* advect.jl:61: ABu(Js',Jr',JCu) */
    }
    free(w);
  }
  free(Jr);
  free(Js);
  free(Tx);
  free(Ty);
  return 0;
}
