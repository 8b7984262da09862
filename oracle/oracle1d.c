/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into libeconpdes_cuda.so).
 *
 * Plain-C restatement of the reference's serial algorithm for 1-D, single-unknown problems
 * (configs C1 and C4), used (a) as a second, independent checker next to the numpy oracle
 * and (b) as the CPU baseline bench.py times on the box's host cores: the numpy oracle is
 * ~15x slower than the Julia reference on this path, a C loop is a fairer stand-in.
 *
 * Follows, function by function (reference root = EconPDEs.jl v3.0.0):
 *   spacing(), differentiate1d()  src/utils/derivatives.jl:39-50, 61-100
 *   sweep()                       pde!, src/pdesolve.jl:436-462
 *   fd_jacobian()                 FiniteDiff.jl 2.x finite_difference_jacobian! (forward, colorvec):
 *                                 f(x) once, then per color: eps = sqrt(eps)*max(1,sqrt(||x1.*mask||)),
 *                                 x1 += eps*mask, f(x1), columns of that color = (f(x1)-f(x))/eps,
 *                                 x1 -= eps*mask     (called at src/finiteschemesolve.jl:184)
 *   tridiag_lu_solve()            Julia's `J \ R` (UMFPACK: LU with partial pivoting), here the
 *                                 tridiagonal special case (LAPACK dgtsv algorithm)
 *   newton()                      NLsolve 4.x newton_ with Static line search + _newtonstep
 *                                 (src/finiteschemesolve.jl:125-147)
 *   solve()                       finiteschemesolve + SERTimeStepController
 *                                 (src/finiteschemesolve.jl:236-275, 343-465)
 * Models: README.md:58-67, examples/asset_pricing/campbell_cochrane.jl:66-84, he_krishnamurthy.jl:45-92, wachter.jl:66-92.
 * Compile with -ffp-contract=off (Julia never fuses a*b+c).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SQRT_EPS 1.4901161193847656e-8

typedef double (*pde_fn)(const double* prm, double x, double v, double v_up, double v_down, double vxx);

static double readme_linear(const double* prm, double x, double v, double v_up, double v_down, double vxx) {
  const double rho = prm[0], kappa = prm[1];
  (void)vxx;
  const double mu = -kappa * (x - 0.5);
  return -(x + mu * (mu >= 0 ? v_up : v_down) - rho * v);
}

static double campbell_cochrane(const double* prm, double s, double p, double ps_up, double ps_down, double pss) {
  const double mu = prm[0], sigma = prm[1], gamma = prm[2], rho = prm[3], kappa_s = prm[4], b = prm[5];
  const double Sbar = sigma * sqrt(gamma / (kappa_s - b / gamma));
  const double sbar = log(Sbar);
  const double lam = 1 / Sbar * sqrt(1 - 2 * (s - sbar)) - 1;
  const double mus = -kappa_s * (s - sbar);
  const double sigs = lam * sigma;
  const double mus_eff = mus + sigs * (sigma - gamma * (sigma + sigs));
  const double ps = (mus_eff >= 0) ? ps_up : ps_down;
  const double sigp = ps / p * sigs;
  const double mup = ps / p * mus + 0.5 * pss / p * (sigs * sigs);
  const double kap = gamma * (sigma + sigs);
  const double r = rho + gamma * mu - (gamma * kappa_s - b) / 2 + b * (sbar - s);
  return -p * (1 / p + mu + mup + sigp * sigma - r - kap * (sigma + sigp));
}

/* examples/asset_pricing/he_krishnamurthy.jl:45-92 (the `@goto start` second pass as a loop) */
static double he_krishnamurthy(const double* prm, double x, double pS, double pSx_up, double pSx_down, double pSxx) {
  const double m = prm[0], lam = prm[1], g = prm[2], sigma = prm[3], rho = prm[4], gamma = prm[5], l = prm[6];
  double pSx = pSx_up, p, px, pxx, aI, sigx, spS, sp, sR, kappa, mux;
  int iter = 0;
  for (;;) {
    p = (1 + l) / (x / pS + (1 - x) * rho);
    px = -(1 + l) / ((x / pS + (1 - x) * rho) * (x / pS + (1 - x) * rho)) * (1 / pS - rho - x / (pS * pS) * pSx);
    pxx = -(1 + l) / ((x / pS + (1 - x) * rho) * (x / pS + (1 - x) * rho)) *
              (-1 / (pS * pS) * pSx - 1 / (pS * pS) * pSx - x / (pS * pS) * pSxx + 2 * x / (pS * pS * pS) * (pSx * pSx)) +
          2 * (1 + l) / ((x / pS + (1 - x) * rho) * (x / pS + (1 - x) * rho) * (x / pS + (1 - x) * rho)) *
              ((1 / pS - rho - x / (pS * pS) * pSx) * (1 / pS - rho - x / (pS * pS) * pSx));
    aI = 1 / (1 - lam * (1 - x));
    if (x * (1 + m) * aI < 1) aI = 1 / (x * (1 + m));
    sigx = x * (aI - 1) * sigma / (1 - x * (aI - 1) * px / p);
    spS = pSx / pS * sigx;
    sp = px / p * sigx;
    sR = sigma + sp;
    kappa = gamma * (aI * sR - spS);
    mux = x * (1 - x) * ((aI - (1 - aI * x) / (1 - x)) * kappa * sR - 1 / pS + rho - (aI - (1 - aI * x) / (1 - x)) * (sR * sR) -
                         l / ((1 - x) * p));
    if (iter == 0 && mux < 0) {
      iter = 1;
      pSx = pSx_down;
      continue;
    }
    break;
  }
  {
    const double mupS = pSx / pS * mux + 0.5 * pSxx / pS * (sigx * sigx);
    const double mup = px / p * mux + 0.5 * pxx / p * (sigx * sigx);
    const double r = 1 / p + g + mup + sigma * sp - kappa * sR;
    const double sCS = kappa / gamma;
    const double muCS = (r - rho) / gamma + (1 + 1 / gamma) / (2 * gamma) * (kappa * kappa);
    return -(1 / pS + muCS + mupS + sCS * spS - r - kappa * (sCS + spS));
  }
}

/* examples/asset_pricing/wachter.jl:66-92; prm[11..13] = mgf(Z, 1), mgf(Z, -gamma), mgf(Z, 1 - gamma) */
static double wachter(const double* prm, double lam, double p, double pl_up, double pl_down, double pll) {
  const double mu = prm[0], sigma = prm[1], lambar = prm[2], kappal = prm[3], nul = prm[4], rho = prm[7], gamma = prm[8],
               psi = prm[9], m1 = prm[11], mg = prm[12], m1g = prm[13];
  const double mul = kappal * (lambar - lam);
  const double sigl = nul * sqrt(lam);
  const double pl = (mul >= 0) ? pl_up : pl_down;
  const double mup = pl / p * mul + 0.5 * pll / p * (sigl * sigl);
  const double sp = pl / p * sigl;
  const double kZc = gamma * sigma;
  const double kZl = (gamma * psi - 1) / (psi - 1) * sp;
  const double eta = lam * (m1 - 1 + mg - 1 - (m1g - 1));
  const double r = rho + mu / psi - (1 + 1 / psi) / 2 * gamma * (sigma * sigma) - lam * (mg - 1) +
                   (1 / psi - gamma) / (1 - gamma) * lam * (m1g - 1) - (gamma * psi - 1) / (2 * (psi - 1)) * (sp * sp);
  return -p * (1 / p + mu + mup + lam * (m1 - 1) - r - kZc * sigma - kZl * sp - eta);
}

static pde_fn model(int id) {
  switch (id) {
    case 0: return readme_linear;
    case 1: return campbell_cochrane;
    case 2: return he_krishnamurthy;
    default: return wachter;
  }
}

typedef struct {
  int n;
  const double* x;
  const double* prm;
  pde_fn f;
  const double* ypost;  /* implicit term: R = G - (ypost - y)/delta */
  double delta;
} ctx_t;

/* G!(ydot, y) fused with G_helper! (src/finiteschemesolve.jl:158-162) */
static void sweep(const ctx_t* c, const double* y, double* out, int implicit) {
  const int n = c->n;
  const double* x = c->x;
  for (int i = 0; i < n; ++i) {
    const int i1 = i > 1 ? i : 1, i0 = i - 1 > 0 ? i - 1 : 0;
    const int j1 = i + 1 < n - 1 ? i + 1 : n - 1, j0 = i < n - 2 ? i : n - 2;
    const double dm = x[i1] - x[i0], dp = x[j1] - x[j0], d = (dm + dp) / 2;
    const double up = i < n - 1 ? (y[i + 1] - y[i]) / dp : 0.0;
    const double dn = i > 0 ? (y[i] - y[i - 1]) / dm : 0.0;
    double dd;
    if (i > 0 && i < n - 1) dd = y[i + 1] / (dp * d) + y[i - 1] / (dm * d) - 2 * y[i] / (dp * dm);
    else if (i == 0) dd = y[1] / (dp * d) + (y[0] - 0.0 * dm) / (dm * d) - 2 * y[0] / (dp * dm);
    else dd = (y[n - 1] + 0.0 * dp) / (dp * d) + y[n - 2] / (dm * d) - 2 * y[n - 1] / (dp * dm);
    double r = c->f(c->prm, x[i], y[i], up, dn, dd);
    if (implicit) r = r - (c->ypost[i] - y[i]) / c->delta;
    out[i] = r;
  }
}

/* colored forward differences into the three diagonals (lo[i] = J[i,i-1], di[i], up[i] = J[i,i+1]) */
static void fd_jacobian(const ctx_t* c, const double* y, double* x1, double* fx, double* fx1, double* lo, double* di,
                        double* up) {
  const int n = c->n;
  memcpy(x1, y, n * sizeof(double));
  sweep(c, y, fx, 1);
  const int ncolors = n < 3 ? n : 3;
  for (int col = 0; col < ncolors; ++col) {
    double ss = 0.0;
    for (int j = col; j < n; j += 3) ss += x1[j] * x1[j];
    const double nrm = sqrt(ss);
    double eps = SQRT_EPS * fabs(sqrt(nrm));
    if (eps < SQRT_EPS) eps = SQRT_EPS;
    for (int j = col; j < n; j += 3) x1[j] = x1[j] + eps;
    sweep(c, x1, fx1, 1);
    for (int j = col; j < n; j += 3) {  /* column j: rows j-1, j, j+1 */
      if (j > 0) up[j - 1] = (fx1[j - 1] - fx[j - 1]) / eps;
      di[j] = (fx1[j] - fx[j]) / eps;
      if (j < n - 1) lo[j + 1] = (fx1[j + 1] - fx[j + 1]) / eps;
    }
    for (int j = col; j < n; j += 3) x1[j] = x1[j] - eps;
  }
}

/* dgtsv: Gaussian elimination with partial pivoting; overwrites the diagonals; returns 1 if singular */
static int tridiag_lu_solve(int n, double* dl, double* d, double* du, double* b, double* du2) {
  /* dl[i] = J[i+1,i] (i<n-1), d[i], du[i] = J[i,i+1]; du2: second super-diagonal created by pivoting */
  memset(du2, 0, n * sizeof(double));
  for (int i = 0; i < n - 1; ++i) {
    if (fabs(d[i]) >= fabs(dl[i])) {
      if (d[i] == 0.0) return 1;
      const double m = dl[i] / d[i];
      d[i + 1] -= m * du[i];
      b[i + 1] -= m * b[i];
      dl[i] = 0.0;
    } else {
      const double m = d[i] / dl[i];
      d[i] = dl[i];
      double t = d[i + 1];
      d[i + 1] = du[i] - m * t;
      if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -m * du2[i]; }
      du[i] = t;
      t = b[i];
      b[i] = b[i + 1];
      b[i + 1] = t - m * b[i + 1];
    }
  }
  if (d[n - 1] == 0.0) return 1;
  b[n - 1] /= d[n - 1];
  if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
  for (int i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
  return 0;
}

typedef struct {
  double* x1; double* fx; double* fx1; double* lo; double* di; double* up; double* dl; double* du; double* p; double* y; double* R; double* du2;
} work_t;

/* implicit_timestep + _newtonstep + newton_: returns nl = ||R(zero)||_2/n (Inf if singular) */
static double newton(ctx_t* c, const double* ypost, double delta, double ftol, int maxit, work_t* w, double* yout,
                     int* iters) {
  const int n = c->n;
  c->ypost = ypost;
  c->delta = delta;
  double* y = w->y;
  memcpy(y, ypost, n * sizeof(double));
  sweep(c, y, w->R, 1);
  int it = 0, stopped = 0, fconv, xconv = 0;
  double mx = 0.0;
  for (int i = 0; i < n; ++i) { if (isnan(w->R[i])) stopped = 1; if (fabs(w->R[i]) > mx) mx = fabs(w->R[i]); }
  fconv = !stopped && mx <= ftol;
  while (!stopped && !fconv && !xconv && it < maxit) {
    ++it;
    fd_jacobian(c, y, w->x1, w->fx, w->fx1, w->lo, w->di, w->up);
    for (int i = 0; i < n - 1; ++i) { w->dl[i] = w->lo[i + 1]; w->du[i] = w->up[i]; }
    memcpy(w->p, w->R, n * sizeof(double));
    if (tridiag_lu_solve(n, w->dl, w->di, w->du, w->p, w->du2)) {   /* SingularException -> (y0, Inf) */
      memcpy(yout, ypost, n * sizeof(double));
      *iters = it;
      return INFINITY;
    }
    xconv = 1;
    for (int i = 0; i < n; ++i) {
      const double xn = y[i] + (-w->p[i]);
      if (xn != y[i]) xconv = 0;
      y[i] = xn;
      if (isnan(xn)) stopped = 1;
    }
    sweep(c, y, w->R, 1);
    mx = 0.0;
    for (int i = 0; i < n; ++i) { if (isnan(w->R[i])) stopped = 1; if (fabs(w->R[i]) > mx) mx = fabs(w->R[i]); }
    fconv = !stopped && mx <= ftol;
  }
  sweep(c, y, w->R, 1);   /* final_residual, src/finiteschemesolve.jl:139-141 */
  double ss = 0.0;
  for (int i = 0; i < n; ++i) ss += w->R[i] * w->R[i];
  memcpy(yout, y, n * sizeof(double));
  *iters = it;
  return sqrt(ss) / n;
}

static double gnorm(const ctx_t* c, const double* y, double* tmp) {
  sweep(c, y, tmp, 0);
  double ss = 0.0;
  for (int i = 0; i < c->n; ++i) ss += tmp[i] * tmp[i];
  return sqrt(ss) / c->n;
}

/* finiteschemesolve with the reference defaults; returns 0, fills outputs */
int oracle1d_solve(int model_id, const double* prm, int n, const double* x, const double* guess, double delta0,
                   int maxiters, double* sol, double* residual_norm, int* n_outer, int* n_newton) {
  const double abstol = SQRT_EPS, inner_abstol = SQRT_EPS, scale = 10.0, min_delta = 1e-9, growth = 10.0;
  const int inner_maxiters = 10;
  ctx_t c = {n, x, prm, model(model_id), NULL, INFINITY};
  double* buf = (double*)malloc(sizeof(double) * n * 14);
  work_t w = {buf, buf + n, buf + 2 * n, buf + 3 * n, buf + 4 * n, buf + 5 * n, buf + 6 * n, buf + 7 * n, buf + 8 * n,
              buf + 9 * n, buf + 10 * n, buf + 13 * n};
  double* ypost = buf + 11 * n;
  double* ytrial = buf + 12 * n;
  memcpy(ypost, guess, n * sizeof(double));
  double r = gnorm(&c, ypost, w.fx);
  *n_outer = 0;
  *n_newton = 0;
  if (isnan(r)) { free(buf); return -2; }
  if (r > abstol) {
    double dl = fabs(delta0), acc = 1.0, cap = INFINITY;
    const double drel = delta0 < 0 ? -1.0 : 1.0;
    int iter = 0;
    while (iter < maxiters && dl >= min_delta && r > abstol) {
      ++iter;
      int its = 0;
      const double nl = newton(&c, ypost, dl * drel, inner_abstol, inner_maxiters, &w, ytrial, &its);
      *n_newton += its;
      const double old = r;
      double rn = gnorm(&c, ytrial, w.fx);
      if (isnan(rn)) rn = INFINITY;
      if (nl <= inner_abstol && rn <= growth * old) {
        if (rn <= old) acc *= scale; else acc = 1.0;          /* accept_step! */
        const double proposed = dl * acc * old / rn;
        const double limit = cap;                              /* maxΔ = Inf */
        if (proposed > limit) { dl = limit; acc = 1.0; } else dl = proposed;
        cap = cap * scale;
        memcpy(ypost, ytrial, n * sizeof(double));
        r = rn;
      } else {                                                 /* reject_step! */
        acc = 1.0;
        cap = dl / 2;
        dl /= 10;
      }
    }
    *n_outer = iter;
  }
  memcpy(sol, ypost, n * sizeof(double));
  *residual_norm = r;
  free(buf);
  return 0;
}

/* B independent problems, row-major [B][n] / [B][nprm], split over `nthreads` pthreads
 * (the image has no libgomp).  The reference itself is single-threaded; a user would run such
 * a sweep with one Julia task per core. */
#include <pthread.h>
typedef struct {
  int model_id, nprm, n, rc;
  long b0, b1;
  const double *prm, *x, *guess;
  double *sol, *rn;
  int *no, *nn;
} job_t;

static void* run_job(void* arg) {
  job_t* j = (job_t*)arg;
  for (long b = j->b0; b < j->b1; ++b) {
    const int r = oracle1d_solve(j->model_id, j->prm + b * j->nprm, j->n, j->x + b * j->n, j->guess + b * j->n, 1.0,
                                 100, j->sol + b * j->n, j->rn + b, j->no + b, j->nn + b);
    if (r != 0) j->rc = r;
  }
  return NULL;
}

int oracle1d_solve_batch(int model_id, int nprm, long B, int n, const double* prm, const double* x,
                         const double* guess, double* sol, double* residual_norm, int* n_outer, int* n_newton,
                         int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  job_t jobs[256];
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) {
    job_t j = {model_id, nprm, n, 0, B * t / nthreads, B * (t + 1) / nthreads, prm, x, guess, sol, residual_norm,
               n_outer, n_newton};
    jobs[t] = j;
    if (nthreads == 1) run_job(&jobs[t]);
    else pthread_create(&th[t], NULL, run_job, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    if (jobs[t].rc != 0) rc = jobs[t].rc;
  }
  return rc;
}
