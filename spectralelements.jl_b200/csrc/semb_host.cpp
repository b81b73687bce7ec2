// Host-side 1-D set-up routines of libsemb (no GPU): Gauss-Lobatto-Legendre rule, Lagrange
// derivative / interpolation matrices, the 1-D SEM grid, BDF/EXT coefficients, slab partition.
// They restate (reference /root/reference/src): FastGaussQuadrature.gausslobatto as called at
// mesh.jl:70-71; derivMat.jl:9-35; interp.jl:10-35; semmesh.jl:9-27; time.jl:31-53.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/semb.h"

void semb_set_error(const char* fmt, ...);

namespace {
// Legendre P_n(x) and P_{n-1}(x) via Bonnet's recurrence
void legendre_pair(int n, double x, double* pn, double* pnm1) {
  double a = 1.0, b = x;
  if (n == 0) {
    *pn = 1.0;
    *pnm1 = 0.0;
    return;
  }
  for (int k = 2; k <= n; ++k) {
    const double c = ((2.0 * k - 1.0) * x * b - (k - 1.0) * a) / k;
    a = b;
    b = c;
  }
  *pn = b;
  *pnm1 = a;
}
}  // namespace

extern "C" int semb_gausslobatto(int n, double* z, double* w) {
  if (n < 2 || !z || !w) {
    semb_set_error("semb_gausslobatto: need n >= 2 and non-null outputs");
    return SEMB_EINVAL;
  }
  const int N = n - 1;
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n; ++i) {
    double x = -std::cos(pi * i / N);  // Chebyshev-Gauss-Lobatto start
    if (i > 0 && i < N) {
      for (int it = 0; it < 100; ++it) {
        double pn, pm;
        legendre_pair(N, x, &pn, &pm);
        // q = (1-x^2) P_N' = N (P_{N-1} - x P_N),  q' = -N (N+1) P_N
        const double q = N * (pm - x * pn);
        const double dq = -1.0 * N * (N + 1) * pn;
        const double dx = q / dq;
        x -= dx;
        if (std::fabs(dx) < 1e-16) break;
      }
    }
    z[i] = x;
  }
  z[0] = -1.0;
  z[N] = 1.0;
  for (int i = 0; i < n / 2; ++i) {  // enforce antisymmetry
    const double a = 0.5 * (z[i] - z[N - i]);
    z[i] = a;
    z[N - i] = -a;
  }
  if (n % 2 == 1) z[N / 2] = 0.0;
  for (int i = 0; i < n; ++i) {
    double pn, pm;
    legendre_pair(N, z[i], &pn, &pm);
    w[i] = 2.0 / (N * (N + 1.0) * pn * pn);
  }
  return SEMB_OK;
}

extern "C" int semb_deriv_mat(int n, const double* x, double* D) {
  if (n < 1 || !x || !D) {
    semb_set_error("semb_deriv_mat: bad arguments");
    return SEMB_EINVAL;
  }
  std::vector<double> a(n, 1.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j)
      if (j != i) a[i] *= (x[i] - x[j]);
    a[i] = 1.0 / a[i];  // barycentric weights (derivMat.jl:13-18)
  }
  for (int i = 0; i < n; ++i) {
    double diag = 0.0;  // derivMat.jl:24-27: row sum of 1/(xi-xj)
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      diag += 1.0 / (x[i] - x[j]);
      D[i + (size_t)j * n] = a[j] / (a[i] * (x[i] - x[j]));  // derivMat.jl:31
    }
    D[i + (size_t)i * n] = diag;
  }
  return SEMB_OK;
}

extern "C" int semb_interp_mat(int no, const double* xo, int ni, const double* xi, double* J) {
  if (no < 0 || ni < 0 || (no && !xo) || (ni && !xi) || (no && ni && !J)) {
    semb_set_error("semb_interp_mat: bad arguments");
    return SEMB_EINVAL;
  }
  std::vector<double> a(ni, 1.0), s(ni, 1.0), t(ni, 1.0);
  for (int i = 0; i < ni; ++i) {
    for (int j = 0; j < ni; ++j)
      if (j != i) a[i] *= (xi[i] - xi[j]);
    a[i] = 1.0 / a[i];
  }
  for (int i = 0; i < no; ++i) {
    const double x = xo[i];
    for (int j = 1; j < ni; ++j) {  // interp.jl:27-30 prefix / suffix products
      s[j] = s[j - 1] * (x - xi[j - 1]);
      t[ni - 1 - j] = t[ni - j] * (x - xi[ni - j]);
    }
    for (int j = 0; j < ni; ++j) J[i + (size_t)j * no] = a[j] * s[j] * t[j];
  }
  return SEMB_OK;
}

extern "C" int semb_semmesh(int E, int n, double* z, double* w) {
  if (E < 1 || n < 2 || !z || !w) {
    semb_set_error("semb_semmesh: bad arguments");
    return SEMB_EINVAL;
  }
  std::vector<double> z0(n), w0(n);
  int rc = semb_gausslobatto(n, z0.data(), w0.data());
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    z0[i] = 0.5 * (z0[i] + 1.0);
    w0[i] = 0.5 * w0[i];
  }
  const double step = 2.0 / E;
  for (int e = 0; e < E; ++e) {
    const double ze0 = (double)e * step + -1.0;
    const double ze1 = (e + 1 == E) ? 1.0 : (double)(e + 1) * step + -1.0;
    const double dz = ze1 - ze0;
    for (int i = 0; i < n; ++i) {
      z[e * n + i] = dz * z0[i] + ze0;
      w[e * n + i] = dz * w0[i];
    }
  }
  return SEMB_OK;
}

extern "C" int semb_bdf_ext_k(int nt, const double* t, int k, double* a, double* b) {
  if (nt < 1 || !t || k < 1 || !a || !b) {
    semb_set_error("semb_bdf_ext_k: bad arguments");
    return SEMB_EINVAL;
  }
  std::vector<double> tu;  // unique(t), first-occurrence order (time.jl:32)
  for (int i = 0; i < nt; ++i) {
    bool seen = false;
    for (double v : tu) seen = seen || (v == t[i]);
    if (!seen) tu.push_back(t[i]);
  }
  const int kk = (int)tu.size() - 1;
  std::vector<double> aa(kk > 0 ? kk : 0), D((size_t)(kk + 1) * (kk + 1));
  if (kk > 0) semb_interp_mat(1, &tu[0], kk, &tu[1], aa.data());  // a = interpMat(t1, t0)
  semb_deriv_mat(kk + 1, tu.data(), D.data());                     // b = derivMat(t)[1,:]
  for (int i = 0; i < k; ++i) a[i] = (i < kk) ? aa[i] : 0.0;
  for (int i = 0; i < k + 1; ++i) b[i] = (i < kk + 1) ? D[0 + (size_t)i * (kk + 1)] : 0.0;
  if (kk == 0) a[0] = 1.0;  // steady state, time.jl:48-50
  return SEMB_OK;
}

extern "C" int semb_partition(int Ey, int nranks, int rank, int* ey0, int* ney) {
  if (Ey < 1 || nranks < 1 || rank < 0 || rank >= nranks || nranks > Ey) {
    semb_set_error("semb_partition: need 1 <= nranks <= Ey and 0 <= rank < nranks (Ey=%d nranks=%d rank=%d)", Ey,
                   nranks, rank);
    return SEMB_EINVAL;
  }
  const long long lo = (long long)rank * Ey / nranks, hi = (long long)(rank + 1) * Ey / nranks;
  if (ey0) *ey0 = (int)lo;
  if (ney) *ney = (int)(hi - lo);
  return SEMB_OK;
}

// Chunk count of the strip-kernel launch plan (grid = strips x chunks): small slabs (no more element rows than one wave
// of CTA slots has room for) get one chunk per element row; larger ones the count in [1 wave, 4 waves], with at least two
// element rows per chunk, that minimises
//     waves * (element rows of the longest chunk + 1/2),   waves = ceil(strips * chunks / slots),
// i.e. the critical path in element-row times: every wave pays the rows of its longest chunk plus a pipeline fill/drain
// of about half a row.  Ties go to the larger count (fuller last wave).  Calibrated on chunk sweeps over eight mesh shapes
// (profiles/r01_sweep_chunks_r1l.txt).  Host logic only: mesh_build_plan (semb_api.cu) calls it, tests call it directly.
extern "C" int semb_plan_chunks(int nstrips, int ney, int slots, int* nchunks) {
  if (nstrips < 1 || ney < 1 || slots < 1 || !nchunks) {
    semb_set_error("semb_plan_chunks: need nstrips, ney, slots >= 1 (got %d, %d, %d)", nstrips, ney, slots);
    return SEMB_EINVAL;
  }
  const int lo = std::max(1, slots / nstrips), hi = std::max(lo, 4 * slots / nstrips);
  int best = 1;
  if (ney <= lo) {
    best = ney;
  } else if (ney / 2 <= lo) {
    best = lo;
  } else {
    double best_cost = 1e300;
    for (int nc = lo; nc <= hi && nc <= ney / 2; ++nc) {
      const long long ctas = (long long)nc * nstrips;
      const long long waves = (ctas + slots - 1) / slots;
      const int rows = (ney + nc - 1) / nc;
      const double cost = (double)waves * ((double)rows + 0.5);
      if (cost <= best_cost) {
        best_cost = cost;
        best = nc;
      }
    }
  }
  if (best > ney) best = ney;
  if (best < 1) best = 1;
  *nchunks = best;
  return SEMB_OK;
}

// Halo plan of a y-slab: which neighbour ranks exist below / above (periodic wrap included).  The
// exchange itself (semb_api.cu: halo_exchange) posts, in this order, send(last row -> hi),
// send(first row -> lo), recv(lo), recv(hi): with two ranks and periodic y both neighbours are the
// same peer and the order is what pairs the messages correctly.
extern "C" int semb_halo_plan(int nranks, int rank, int pery, int* halo_lo, int* halo_hi, int* rank_lo,
                              int* rank_hi) {
  if (nranks < 1 || rank < 0 || rank >= nranks) {
    semb_set_error("semb_halo_plan: bad nranks/rank %d/%d", nranks, rank);
    return SEMB_EINVAL;
  }
  const int lo = (rank > 0) || (pery && nranks > 1), hi = (rank < nranks - 1) || (pery && nranks > 1);
  if (halo_lo) *halo_lo = lo;
  if (halo_hi) *halo_hi = hi;
  if (rank_lo) *rank_lo = lo ? (rank - 1 + nranks) % nranks : -1;
  if (rank_hi) *rank_hi = hi ? (rank + 1) % nranks : -1;
  return SEMB_OK;
}

// ---- FDM preconditioner set-up (examples/p2d_explicit.jl:109-141, lapl.jl:105-119) ----------------------
// Generalised eigen-decomposition A S = B S diag(lam), S' B S = I, of the 1-D stiffness / mass pair of ONE reference
// element (half-length 1: A = D' diag(w) D, B = diag(w)) extended by one node into each neighbour of the same size.
// kinds (left, right): 0 = neighbour element (extension node = its first node off the interface, zero beyond it),
// 1 = Dirichlet boundary (the boundary node itself is removed), 2 = free boundary (no extension).
// S is (n+2) x (n+2) column-major with rows [left ext, own 0..n-1, right ext] (zero rows for removed nodes), lam has
// n+2 entries (+inf for the padding modes).  An element of half-length h uses S/sqrt(h), lam/h^2.
// B is diagonal, so the pair reduces to the symmetric problem B^-1/2 A B^-1/2 = Q diag(lam) Q', S = B^-1/2 Q, which the
// cyclic Jacobi method solves to machine precision at these sizes (<= 19).
extern "C" int semb_fdm_tables(int n, const double* D, const double* w, int left, int right, double* S, double* lam) {
  if (!D || !w || !S || !lam || n < 3 || n > 62 || left < 0 || left > 2 || right < 0 || right > 2) {
    semb_set_error("semb_fdm_tables: bad argument (3 <= n <= 62, kinds 0..2)");
    return SEMB_EINVAL;
  }
  const int m = n + 2;
  std::vector<double> A0((size_t)n * n, 0.0), A((size_t)m * m, 0.0), B(m, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += D[k + (size_t)i * n] * w[k] * D[k + (size_t)j * n];  // (D' diag(w) D)(i,j)
      A0[i + (size_t)j * n] = s;
    }
  auto Aat = [&](int i, int j) -> double& { return A[i + (size_t)j * m]; };
  for (int i = 0; i < n; ++i) {
    B[i + 1] += w[i];
    for (int j = 0; j < n; ++j) Aat(i + 1, j + 1) += A0[i + (size_t)j * n];
  }
  std::vector<int> active(m, 1);
  if (left == 0) {  // neighbour's nodes (n-2, n-1) sit on extended indices (0, 1)
    for (int i = 0; i < 2; ++i) {
      B[i] += w[n - 2 + i];
      for (int j = 0; j < 2; ++j) Aat(i, j) += A0[(n - 2 + i) + (size_t)(n - 2 + j) * n];
    }
  } else {
    active[0] = 0;
    if (left == 1) active[1] = 0;
  }
  if (right == 0) {  // neighbour's nodes (0, 1) sit on extended indices (n, n+1)
    for (int i = 0; i < 2; ++i) {
      B[n + i] += w[i];
      for (int j = 0; j < 2; ++j) Aat(n + i, n + j) += A0[i + (size_t)j * n];
    }
  } else {
    active[m - 1] = 0;
    if (right == 1) active[n] = 0;
  }
  std::vector<int> idx;
  for (int i = 0; i < m; ++i)
    if (active[i]) idx.push_back(i);
  const int p = (int)idx.size();
  std::vector<double> C((size_t)p * p), Q((size_t)p * p, 0.0);
  for (int i = 0; i < p; ++i) {
    Q[i + (size_t)i * p] = 1.0;
    for (int j = 0; j < p; ++j) C[i + (size_t)j * p] = Aat(idx[i], idx[j]) / std::sqrt(B[idx[i]] * B[idx[j]]);
  }
  for (int sweep = 0; sweep < 100; ++sweep) {  // cyclic Jacobi
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) (i == j ? diag : off) += C[i + (size_t)j * p] * C[i + (size_t)j * p];
    if (off <= 1e-30 * diag) break;
    for (int a = 0; a < p - 1; ++a)
      for (int b = a + 1; b < p; ++b) {
        const double apq = C[a + (size_t)b * p];
        if (apq == 0.0) continue;
        const double theta = (C[b + (size_t)b * p] - C[a + (size_t)a * p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < p; ++k) {  // columns a, b
          const double ka = C[k + (size_t)a * p], kb = C[k + (size_t)b * p];
          C[k + (size_t)a * p] = c * ka - s * kb;
          C[k + (size_t)b * p] = s * ka + c * kb;
        }
        for (int k = 0; k < p; ++k) {  // rows a, b
          const double ak = C[a + (size_t)k * p], bk = C[b + (size_t)k * p];
          C[a + (size_t)k * p] = c * ak - s * bk;
          C[b + (size_t)k * p] = s * ak + c * bk;
        }
        for (int k = 0; k < p; ++k) {
          const double ka = Q[k + (size_t)a * p], kb = Q[k + (size_t)b * p];
          Q[k + (size_t)a * p] = c * ka - s * kb;
          Q[k + (size_t)b * p] = s * ka + c * kb;
        }
      }
  }
  std::vector<int> order(p);  // ascending eigenvalues (as LAPACK's eigh): deterministic layout
  for (int i = 0; i < p; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return C[x + (size_t)x * p] < C[y + (size_t)y * p]; });
  for (int q = 0; q < m * m; ++q) S[q] = 0.0;
  for (int q = 0; q < m; ++q) lam[q] = INFINITY;
  for (int c = 0; c < p; ++c) {
    const int o = order[c];
    lam[c] = C[o + (size_t)o * p];
    for (int i = 0; i < p; ++i) S[idx[i] + (size_t)c * m] = Q[i + (size_t)o * p] / std::sqrt(B[idx[i]]);
  }
  return SEMB_OK;
}
