/* CPU oracle in C for the SemPy Poisson hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C (OpenMP over elements) restatement of the reference algorithm, used as the timed CPU
 * baseline of bench.py and cross-checked against the NumPy oracle / golden fixtures in
 * tests/test_oracle_c.py.  Nothing under sempy_b200/ links or loads it.
 *
 * Reference lines followed (paths relative to the SemPy checkout):
 *   so_ax        sempy/elliptic.py:9-27 with sempy/gradient.py:6-23 (gradient) and :39-56 (transpose)
 *   so_gs        sempy/mesh.py:444-447 -> gslib gs(x, gs_double, gs_add); stated in
 *                sempy/loopy/loopy_kernels.py:26-39 (sum in segment order, write back to all copies)
 *   so_cg        sempy/elliptic.py:30-73 (precond = 0) and sempy/iterative.py:45-85 written on local
 *                vectors with Minv = dinv (precond = 1), same stop rules
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int so_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ap_e = D^T_r (G_rr p_r + G_rs p_s + G_rt p_t) + D^T_s (...) + D^T_t (...), element by element.
 * G is the reference's (E,3,3,Np) array; node index k*n*n + j*n + i. */
void so_ax(int64_t E, int n, const double* D, const double* G, const double* p, double* ap) {
  const int nn = n * n, np = nn * n;
#pragma omp parallel
  {
    double* pr = (double*)malloc(sizeof(double) * 6 * (size_t)np);
    double *ps = pr + np, *pt = ps + np, *wr = pt + np, *ws = wr + np, *wt = ws + np;
#pragma omp for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
      const double* u = p + e * np;
      const double* g = G + e * 9 * (int64_t)np;
      for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i) {
            double r = 0.0, s = 0.0, t = 0.0;
            for (int m = 0; m < n; ++m) {
              r += u[k * nn + j * n + m] * D[i * n + m]; /* Ur = V D^T        gradient.py:12-13 */
              s += D[j * n + m] * u[k * nn + m * n + i]; /* Us[k] = D V[k]    gradient.py:15-18 */
              t += D[k * n + m] * u[m * nn + j * n + i]; /* Ut = D V          gradient.py:20-21 */
            }
            const int q = k * nn + j * n + i;
            pr[q] = r;
            ps[q] = s;
            pt[q] = t;
          }
      for (int q = 0; q < np; ++q) { /* elliptic.py:21-23 */
        wr[q] = g[0 * np + q] * pr[q] + g[1 * np + q] * ps[q] + g[2 * np + q] * pt[q];
        ws[q] = g[3 * np + q] * pr[q] + g[4 * np + q] * ps[q] + g[5 * np + q] * pt[q];
        wt[q] = g[6 * np + q] * pr[q] + g[7 * np + q] * ps[q] + g[8 * np + q] * pt[q];
      }
      double* out = ap + e * np;
      for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i) {
            double a = 0.0, b = 0.0, c = 0.0;
            for (int m = 0; m < n; ++m) {
              a += wr[k * nn + j * n + m] * D[m * n + i]; /* V D        gradient.py:45-46 */
              b += D[m * n + j] * ws[k * nn + m * n + i]; /* D^T V[k]   gradient.py:48-51 */
              c += D[m * n + k] * wt[m * nn + j * n + i]; /* D^T V      gradient.py:53-54 */
            }
            out[k * nn + j * n + i] = (a + b) + c; /* gradient.py:56 */
          }
    }
    free(pr);
  }
}

/* in-place direct-stiffness sum over nseg segments of local indices */
void so_gs(int64_t nseg, const int64_t* g2l, const int64_t* gstart, double* x) {
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < nseg; ++s) {
    double acc = 0.0;
    for (int64_t q = gstart[s]; q < gstart[s + 1]; ++q) acc += x[g2l[q]];
    for (int64_t q = gstart[s]; q < gstart[s + 1]; ++q) x[g2l[q]] = acc;
  }
}

static double dot3(int64_t n, const double* w, const double* x, const double* y) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t q = 0; q < n; ++q) s += (w ? w[q] * x[q] : x[q]) * y[q];
  return s;
}

/* CG (precond = 0) / Jacobi-PCG (precond = 1, dinv given) on local vectors; returns the iteration count.
 * work: 4*n doubles (r, p, Ap, z). */
int so_cg(int64_t E, int n, const double* D, const double* G, const double* mask, int64_t nseg, const int64_t* g2l,
          const int64_t* gstart, const double* rmult, const double* dinv, int precond, const double* b, double* x,
          double tol, int maxit, double* work) {
  const int64_t N = E * (int64_t)n * n * n;
  double *r = work, *p = work + N, *Ap = work + 2 * N, *z = work + 3 * N;
  const double norm_b = dot3(N, rmult, b, b);
  const double TOL = fmax(tol * tol * norm_b, tol * tol);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < N; ++q) {
    x[q] = 0.0;
    r[q] = b[q];
    z[q] = precond ? dinv[q] * b[q] : b[q];
    p[q] = z[q];
  }
  double rdotr = precond ? dot3(N, rmult, r, z) : norm_b;
  int niter = 0;
  if (!precond && rdotr < 1.0e-20) return 0; /* elliptic.py:43 */
  while (niter < maxit && rdotr > TOL) {
    so_ax(E, n, D, G, p, Ap);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < N; ++q) Ap[q] *= mask[q];        /* mesh.apply_mask, elliptic.py:49 */
    const double pAp = dot3(N, NULL, Ap, p);                  /* before the gather-scatter, :51 */
    const double alpha = rdotr / pAp;
    so_gs(nseg, g2l, gstart, Ap);                             /* :54 */
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < N; ++q) {
      x[q] = x[q] + alpha * p[q];
      r[q] = r[q] - alpha * Ap[q];
      z[q] = precond ? dinv[q] * r[q] : r[q];
    }
    const double rdotr0 = rdotr;
    rdotr = dot3(N, rmult, r, z);
    const double beta = rdotr / rdotr0;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < N; ++q) p[q] = z[q] + beta * p[q];
    ++niter;
  }
  return niter;
}
