/*
 * grr_oracle.c -- CPU (IEEE double) restatement of the roadmap-construction hot path of
 * krishauser/GlobalRedundancyResolution.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (globalredundancyresolution_b200/) may import, link or execute this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS
 *   - workspace graphs (oracle/graphs.py): PINNED by the reference's shipped fixture
 *     output/atlas/Rfree_W4000/resolution_graph.pickle (topology + numbering).
 *   - everything in this file that restates Klamp't / KrisLibrary / IKDB arithmetic
 *     (FK, Jacobian, NewtonRoot::GlobalSolve, robot.distance / interpolate, so3 log/exp):
 *     "parity unpinned" -- those libraries are un-vendored, un-pinned third-party
 *     dependencies of the reference (README.md:18-33: "Klamp't 0.7+", IKDB git clone) and
 *     are absent from /root/reference and from this image.  The algorithm restated here is
 *     their published/recalled algorithm; parity is anchored on the
 *     reference's own call sites, which are cited function by function below, and on
 *     closed-form known answers (planar FK, finite-difference Jacobians).
 *
 * Layout conventions of this file: array-of-structs, configs are [N][n] doubles,
 * rotation matrices are ROW-major 3x3.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXJ 32

enum { ORC_FREE = 0, ORC_FIXED = 1, ORC_VARIABLE = 2 };

/* IK status codes (shared with the CUDA path's header by value, not by include). */
enum {
  ORC_OK = 0,            /* max|F| < tol                                  */
  ORC_FAIL_STALL = 1,    /* line search stalled (local min / conv. on x)  */
  ORC_FAIL_CONVX = 2,    /* step below tolx                               */
  ORC_FAIL_MAXIT = 3     /* iteration budget exhausted                    */
};

typedef struct {
  int32_t n;                        /* active revolute joints on the chain base -> goal link */
  int32_t mode;                     /* ORC_FREE / ORC_FIXED / ORC_VARIABLE                   */
  int32_t nlinks_eps;               /* "nlinks" of graph.py:944                              */
  int32_t max_iters;                /* IKSolverParams.numIters (believed 50)                 */
  double tol;                       /* IKSolverParams.tol (1e-3)                             */
  double lambda;                    /* damping of the SVD back-substitution (0.01)           */
  double Rp[ORC_MAXJ][9];           /* TParent rotation preceding joint j (fixed links folded)*/
  double tp[ORC_MAXJ][3];           /* TParent translation preceding joint j                 */
  double axis[ORC_MAXJ][3];         /* local joint axis                                      */
  double qmin[ORC_MAXJ], qmax[ORC_MAXJ];
  double ee_local[3];               /* constrained point in the last active joint's frame    */
  double R_tail[9];                 /* goal-link frame relative to the last active frame     */
  double R_goal[9];                 /* fixed-orientation goal (mode FIXED)                   */
  int32_t spin[ORC_MAXJ];           /* 1: Klamp't `joint spin` (continuous rotation; graph.py:556-561) */
} orc_chain;

int orc_sizeof_chain(void) { return (int)sizeof(orc_chain); }

/* ------------------------------------------------------------------------------------- */
/* small 3x3 helpers                                                                      */
/* ------------------------------------------------------------------------------------- */
static void m3_mul(const double *A, const double *B, double *C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof T);
}
static void m3_mulT(const double *A, const double *B, double *C) { /* C = A * B^T */
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
  memcpy(C, T, sizeof T);
}
static void m3_Tmul(const double *A, const double *B, double *C) { /* C = A^T * B */
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
  memcpy(C, T, sizeof T);
}
static void m3_vec(const double *A, const double *v, double *r) {
  double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  double t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  double t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static void v3_cross(const double *a, const double *b, double *c) {
  double t0 = a[1] * b[2] - a[2] * b[1];
  double t1 = a[2] * b[0] - a[0] * b[2];
  double t2 = a[0] * b[1] - a[1] * b[0];
  c[0] = t0; c[1] = t1; c[2] = t2;
}

/* klampt.math.so3.rotation(axis,angle) -- Rodrigues.  axis must be unit. */
static void so3_rotation(const double *a, double th, double *R) {
  double c = cos(th), s = sin(th), v = 1.0 - c;
  R[0] = c + v * a[0] * a[0];        R[1] = v * a[0] * a[1] - s * a[2]; R[2] = v * a[0] * a[2] + s * a[1];
  R[3] = v * a[1] * a[0] + s * a[2]; R[4] = c + v * a[1] * a[1];        R[5] = v * a[1] * a[2] - s * a[0];
  R[6] = v * a[2] * a[0] - s * a[1]; R[7] = v * a[2] * a[1] + s * a[0]; R[8] = c + v * a[2] * a[2];
}

/* klampt.math.so3.from_moment(w): exp map (graph.py:96,152). */
void orc_so3_exp(const double *w, double *R) {
  double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th < 1e-12) {
    /* first-order: I + [w]x */
    R[0] = 1; R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2]; R[4] = 1; R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0]; R[8] = 1;
    return;
  }
  double a[3] = {w[0] / th, w[1] / th, w[2] / th};
  so3_rotation(a, th, R);
}

/* klampt.math.so3.moment(R): log map, |w| in [0,pi] (graph.py:118,154,209,285).
 * theta is taken from atan2(sin,cos) rather than acos so that the same formula is
 * well-conditioned in single precision on the device. */
void orc_so3_log(const double *R, double *w) {
  double v[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
  double s = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  double th = atan2(s, c);
  if (c > -0.9) {
    double k = (s > 1e-8) ? th / s : 1.0;
    w[0] = v[0] * k; w[1] = v[1] * k; w[2] = v[2] * k;
    return;
  }
  /* near pi: axis from the diagonal, signs from the skew part (or off-diagonals at pi) */
  double omc = 1.0 - c;
  double x2 = (R[0] - c) / omc, y2 = (R[4] - c) / omc, z2 = (R[8] - c) / omc;
  double x = sqrt(x2 > 0 ? x2 : 0), y = sqrt(y2 > 0 ? y2 : 0), z = sqrt(z2 > 0 ? z2 : 0);
  /* choose the largest component as sign anchor */
  if (x >= y && x >= z) {
    if (v[0] < 0) x = -x;
    if ((R[1] + R[3]) * x < 0) y = -y;
    if ((R[2] + R[6]) * x < 0) z = -z;
  } else if (y >= z) {
    if (v[1] < 0) y = -y;
    if ((R[1] + R[3]) * y < 0) x = -x;
    if ((R[5] + R[7]) * y < 0) z = -z;
  } else {
    if (v[2] < 0) z = -z;
    if ((R[2] + R[6]) * z < 0) x = -x;
    if ((R[5] + R[7]) * z < 0) y = -y;
  }
  double nn = sqrt(x * x + y * y + z * z);
  if (nn < 1e-300) nn = 1;
  w[0] = x / nn * th; w[1] = y / nn * th; w[2] = z / nn * th;
}

/* inverse left Jacobian of SO(3): d/dt log(R) = Jl^-1(e) * omega_world, e = log(R). */
static void so3_Jl_inv(const double *e, double *D) {
  double t2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  double g;
  if (t2 < 1e-8) g = 1.0 / 12.0;
  else {
    double th = sqrt(t2), h = 0.5 * th;
    g = (1.0 - h * cos(h) / sin(h)) / t2;
  }
  /* D = I - 0.5 [e]x + g [e]x^2,  [e]x^2 = e e^T - |e|^2 I */
  D[0] = 1 + g * (e[0] * e[0] - t2); D[1] = 0.5 * e[2] + g * e[0] * e[1];  D[2] = -0.5 * e[1] + g * e[0] * e[2];
  D[3] = -0.5 * e[2] + g * e[1] * e[0]; D[4] = 1 + g * (e[1] * e[1] - t2); D[5] = 0.5 * e[0] + g * e[1] * e[2];
  D[6] = 0.5 * e[1] + g * e[2] * e[0];  D[7] = -0.5 * e[0] + g * e[2] * e[1]; D[8] = 1 + g * (e[2] * e[2] - t2);
}

/* ------------------------------------------------------------------------------------- */
/* K: robot model (Klamp't RobotModel)                                        */
/* ------------------------------------------------------------------------------------- */

/* Forward kinematics: T_world[j] = T_world[j-1] * TParent[j] * Rot(axis[j], q[j]).
 * Outputs joint origins o[j], world frames R[j] (after joint j), EE point p, goal-link R. */
static void fk_all(const orc_chain *c, const double *q, double (*o)[3], double (*R)[9], double *p,
                   double *Rlink) {
  double Rw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ow[3] = {0, 0, 0};
  for (int j = 0; j < c->n; j++) {
    double t[3], Rj[9];
    m3_vec(Rw, c->tp[j], t);
    ow[0] += t[0]; ow[1] += t[1]; ow[2] += t[2];
    m3_mul(Rw, c->Rp[j], Rw);
    so3_rotation(c->axis[j], q[j], Rj);
    m3_mul(Rw, Rj, Rw);
    if (o) { o[j][0] = ow[0]; o[j][1] = ow[1]; o[j][2] = ow[2]; }
    if (R) memcpy(R[j], Rw, sizeof Rw);
  }
  double t[3];
  m3_vec(Rw, c->ee_local, t);
  p[0] = ow[0] + t[0]; p[1] = ow[1] + t[1]; p[2] = ow[2] + t[2];
  if (Rlink) m3_mul(Rw, c->R_tail, Rlink);
}

/* public: EE point and goal-link rotation at q */
void orc_fk(const orc_chain *c, const double *q, double *p, double *Rlink) { fk_all(c, q, NULL, NULL, p, Rlink); }

/* [recalled] KrisLibrary math/angle: AngleNormalize into [0, 2pi); AngleDiff(a,b) = a - b wrapped into (-pi, pi];
 * AngleInterp(a,b,u) = AngleNormalize(a + u AngleDiff(b,a)).  Used for 'spin' joints only. */
static double angle_normalize(double a) {
  double an = fmod(a, 2 * M_PI);
  return an < 0 ? an + 2 * M_PI : an;
}
static double angle_diff(double a, double b) {
  double d = angle_normalize(angle_normalize(a) - angle_normalize(b));
  return d > M_PI ? d - 2 * M_PI : d;
}
/* Klamp't robot.distance(a,b): L2 over the DOFs, 'joint normal' DOFs by plain difference, 'joint spin' DOFs by the
 * wrapped angular difference (solver.py:805,827,889; graph.py:946,955,967,971) */
double orc_distance(const orc_chain *c, const double *a, const double *b) {
  double s = 0;
  for (int j = 0; j < c->n; j++) { double d = c->spin[j] ? angle_diff(a[j], b[j]) : a[j] - b[j]; s += d * d; }
  return sqrt(s);
}
/* Klamp't robot.interpolate(a,b,u): per-DOF lerp for 'joint normal' DOFs, shorter arc normalised to [0, 2pi) for
 * 'joint spin' DOFs (graph.py:962) */
static void interp_q(const orc_chain *c, const double *a, const double *b, double u, double *x) {
  for (int j = 0; j < c->n; j++) {
    if (c->spin[j]) { double an = angle_normalize(a[j]); x[j] = angle_normalize(fma(u, angle_diff(b[j], an), an)); }
    else x[j] = a[j] + u * (b[j] - a[j]);
  }
}

/* ------------------------------------------------------------------------------------- */
/* IK task: graph.py:59-103 (IKTask.__init__/setParams), residual = Klamp't IKGoalFunction */
/* ------------------------------------------------------------------------------------- */
typedef struct {
  double x[3];      /* position target                                   */
  double R[9];      /* orientation target (FIXED: chain R_goal; VARIABLE: exp(x[3:6])) */
} orc_target;

/* setIKProblem(x) (graph.py:737-745 -> IKTask.setParams :84-103) */
static void set_target(const orc_chain *c, const double *w, orc_target *t) {
  t->x[0] = w[0]; t->x[1] = w[1]; t->x[2] = w[2];
  if (c->mode == ORC_VARIABLE) orc_so3_exp(w + 3, t->R);
  else memcpy(t->R, c->R_goal, sizeof t->R);
}
static int resid_rows(const orc_chain *c) { return c->mode == ORC_FREE ? 3 : 6; }
static int wdim(const orc_chain *c) { return c->mode == ORC_VARIABLE ? 6 : 3; }

/* F(q): rows 0-2 = p_ee - x ; rows 3-5 = log(R_link * R_goal^T) (world-frame error). */
static void residual(const orc_chain *c, const orc_target *t, const double *q, double *F) {
  double p[3], Rl[9];
  fk_all(c, q, NULL, NULL, p, Rl);
  F[0] = p[0] - t->x[0]; F[1] = p[1] - t->x[1]; F[2] = p[2] - t->x[2];
  if (c->mode != ORC_FREE) {
    double Re[9];
    m3_mulT(Rl, t->R, Re);
    orc_so3_log(Re, F + 3);
  }
}

/* F and J (m x n, row-major J[r*n+j]).  Position column a_j x (p - o_j); rotation column
 * Jl^-1(e) a_j. */
static void residual_jac(const orc_chain *c, const orc_target *t, const double *q, double *F, double *J) {
  double o[ORC_MAXJ][3], R[ORC_MAXJ][9], p[3], Rl[9], D[9];
  int n = c->n;
  fk_all(c, q, o, R, p, Rl);
  F[0] = p[0] - t->x[0]; F[1] = p[1] - t->x[1]; F[2] = p[2] - t->x[2];
  if (c->mode != ORC_FREE) {
    double Re[9];
    m3_mulT(Rl, t->R, Re);
    orc_so3_log(Re, F + 3);
    so3_Jl_inv(F + 3, D);
  }
  for (int j = 0; j < n; j++) {
    double a[3], r[3], col[3];
    m3_vec(R[j], c->axis[j], a);
    r[0] = p[0] - o[j][0]; r[1] = p[1] - o[j][1]; r[2] = p[2] - o[j][2];
    v3_cross(a, r, col);
    J[0 * n + j] = col[0]; J[1 * n + j] = col[1]; J[2 * n + j] = col[2];
    if (c->mode != ORC_FREE) {
      double da[3];
      m3_vec(D, a, da);
      J[3 * n + j] = da[0]; J[4 * n + j] = da[1]; J[5 * n + j] = da[2];
    }
  }
}

void orc_residual_jac(const orc_chain *c, const double *w, const double *q, double *F, double *J) {
  orc_target t; set_target(c, w, &t); residual_jac(c, &t, q, F, J);
}
void orc_residual(const orc_chain *c, const double *w, const double *q, double *F) {
  orc_target t; set_target(c, w, &t); residual(c, &t, q, F);
}

/* ikError(q,x): 2-norm of the residual (graph.py:887-893) */
double orc_ik_error(const orc_chain *c, const double *w, const double *q) {
  double F[6]; orc_target t; set_target(c, w, &t); residual(c, &t, q, F);
  double s = 0; for (int r = 0; r < resid_rows(c); r++) s += F[r] * F[r];
  return sqrt(s);
}

/* ------------------------------------------------------------------------------------- */
/* K3/K4: IKDB IKProblem.solve -> Klamp't IKSolver.solve -> NewtonRoot::GlobalSolve        */
/* (reference call sites solver.py:799,821,884; graph.py:787,790)          */
/* ------------------------------------------------------------------------------------- */
typedef struct { int32_t status, iters, evals; double resid; } orc_ikinfo;

static double maxabs(const double *F, int m) {
  double r = 0; for (int i = 0; i < m; i++) { double a = fabs(F[i]); if (a > r) r = a; } return r;
}

/* p = -J^T (J J^T + lambda^2 I)^-1 F  (== damped SVD back-substitution), Cholesky m x m */
static void damped_step(const double *J, const double *F, int m, int n, double lambda, double *p) {
  double A[36], y[6];
  for (int r = 0; r < m; r++)
    for (int s = 0; s <= r; s++) {
      double acc = 0;
      for (int j = 0; j < n; j++) acc += J[r * n + j] * J[s * n + j];
      A[r * m + s] = acc + (r == s ? lambda * lambda : 0.0);
    }
  /* Cholesky A = L L^T (lower, in place) */
  for (int r = 0; r < m; r++) {
    for (int s = 0; s <= r; s++) {
      double acc = A[r * m + s];
      for (int k = 0; k < s; k++) acc -= A[r * m + k] * A[s * m + k];
      if (r == s) A[r * m + r] = sqrt(acc);
      else A[r * m + s] = acc / A[s * m + s];
    }
  }
  for (int r = 0; r < m; r++) {
    double acc = F[r];
    for (int k = 0; k < r; k++) acc -= A[r * m + k] * y[k];
    y[r] = acc / A[r * m + r];
  }
  for (int r = m - 1; r >= 0; r--) {
    double acc = y[r];
    for (int k = r + 1; k < m; k++) acc -= A[k * m + r] * y[k];
    y[r] = acc / A[r * m + r];
  }
  for (int j = 0; j < n; j++) {
    double acc = 0;
    for (int r = 0; r < m; r++) acc += J[r * n + j] * y[r];
    p[j] = -acc;
  }
}

/* exported for the test that pins the stated equivalence with the SVD damped back-substitution */
void orc_damped_step(const double *J, const double *F, int m, int n, double lambda, double *p) { damped_step(J, F, m, n, lambda, p); }

/* One IK solve from seed x (in/out).  Returns status. */
int orc_ik_solve(const orc_chain *c, const double *w, double *x, orc_ikinfo *info) {
  const int n = c->n, m = resid_rows(c);
  const double tolf = c->tol, tolx = 0.01 * c->tol, stepMax = 10.0, ALF = 1e-4;
  orc_target t; set_target(c, w, &t);
  double F[6], J[6 * ORC_MAXJ], g[ORC_MAXJ], p[ORC_MAXJ], xold[ORC_MAXJ];
  double bmin[ORC_MAXJ], bmax[ORC_MAXJ];
  int evals = 0, iters = 0, status = ORC_FAIL_MAXIT;
  for (int j = 0; j < n; j++) {
    /* revolute joints with range >= 2pi are unbounded */
    if (c->spin[j] || c->qmax[j] - c->qmin[j] >= 2 * M_PI) { bmin[j] = -INFINITY; bmax[j] = INFINITY; }
    else { bmin[j] = c->qmin[j]; bmax[j] = c->qmax[j]; }
    if (x[j] < bmin[j]) x[j] = bmin[j];
    if (x[j] > bmax[j]) x[j] = bmax[j];
  }
  residual(c, &t, x, F); evals++;
  double f = 0; for (int r = 0; r < m; r++) f += F[r] * F[r]; f *= 0.5;
  if (maxabs(F, m) < tolf) { status = ORC_OK; goto done; }
  {
    double xn = 0; for (int j = 0; j < n; j++) xn += x[j] * x[j];
    xn = sqrt(xn);
    const double stpmax = stepMax * (xn > (double)n ? xn : (double)n);
    for (int it = 0; it < c->max_iters; it++) {
      residual_jac(c, &t, x, F, J);
      for (int j = 0; j < n; j++) {
        double acc = 0; for (int r = 0; r < m; r++) acc += J[r * n + j] * F[r];
        g[j] = acc; xold[j] = x[j];
      }
      const double fold = f;
      damped_step(J, F, m, n, c->lambda, p);
      double pn = 0; for (int j = 0; j < n; j++) pn += p[j] * p[j];
      pn = sqrt(pn);
      if (pn > stpmax) for (int j = 0; j < n; j++) p[j] *= stpmax / pn;
      double slope = 0; for (int j = 0; j < n; j++) slope += g[j] * p[j];
      int stalled = 0;
      if (!(slope < 0)) stalled = 1;
      else {
        /* Numerical-Recipes lnsrch with box clamp at every trial */
        double test = 0;
        for (int j = 0; j < n; j++) {
          double tmp = fabs(p[j]) / fmax(fabs(xold[j]), 1.0);
          if (tmp > test) test = tmp;
        }
        const double alamin = tolx / test;
        double alam = 1.0, alam2 = 0, f2 = 0;
        for (;;) {
          for (int j = 0; j < n; j++) {
            double v = xold[j] + alam * p[j];
            if (v < bmin[j]) v = bmin[j];
            if (v > bmax[j]) v = bmax[j];
            x[j] = v;
          }
          residual(c, &t, x, F); evals++;
          f = 0; for (int r = 0; r < m; r++) f += F[r] * F[r]; f *= 0.5;
          if (alam < alamin) { stalled = 1; break; }
          if (f <= fold + ALF * alam * slope) break;
          double tmplam;
          if (alam == 1.0) tmplam = -slope / (2.0 * (f - fold - slope));
          else {
            double rhs1 = f - fold - alam * slope, rhs2 = f2 - fold - alam2 * slope;
            double a = (rhs1 / (alam * alam) - rhs2 / (alam2 * alam2)) / (alam - alam2);
            double b = (-alam2 * rhs1 / (alam * alam) + alam * rhs2 / (alam2 * alam2)) / (alam - alam2);
            if (a == 0.0) tmplam = -slope / (2.0 * b);
            else {
              double disc = b * b - 3.0 * a * slope;
              if (disc < 0.0) tmplam = 0.5 * alam;
              else if (b <= 0.0) tmplam = (-b + sqrt(disc)) / (3.0 * a);
              else tmplam = -slope / (b + sqrt(disc));
            }
            if (tmplam > 0.5 * alam) tmplam = 0.5 * alam;
          }
          alam2 = alam; f2 = f;
          alam = fmax(tmplam, 0.1 * alam);
        }
      }
      iters = it + 1;
      if (stalled) {
        /* x <- xold; F(xold) is not below tolf (checked before), so this is a failure */
        for (int j = 0; j < n; j++) x[j] = xold[j];
        residual(c, &t, x, F);
        status = ORC_FAIL_STALL; goto done;
      }
      if (maxabs(F, m) < tolf) { status = ORC_OK; goto done; }
      double test = 0;
      for (int j = 0; j < n; j++) {
        double tmp = fabs(x[j] - xold[j]) / fmax(fabs(x[j]), 1.0);
        if (tmp > test) test = tmp;
      }
      if (test < tolx) { status = ORC_FAIL_CONVX; goto done; }
    }
    status = ORC_FAIL_MAXIT;
  }
done:
  /* [recalled] robot.NormalizeAngles after the solve: 'spin' joints come back in [0, 2pi) */
  for (int j = 0; j < n; j++) if (c->spin[j]) x[j] = angle_normalize(x[j]);
  if (info) { info->status = status; info->iters = iters; info->evals = evals; info->resid = maxabs(F, m); }
  return status;
}

/* Batch of independent solves; seeds [N][n], targets [N][dw].  OpenMP over instances. */
void orc_ik_solve_batch(const orc_chain *c, int64_t N, const double *targets, const double *seeds,
                        double *out_q, int32_t *status, int32_t *iters, int32_t *evals, double *resid) {
  const int n = c->n, dw = wdim(c);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < N; i++) {
    double x[ORC_MAXJ]; orc_ikinfo info;
    memcpy(x, seeds + i * n, sizeof(double) * n);
    orc_ik_solve(c, targets + i * dw, x, &info);
    memcpy(out_q + i * n, x, sizeof(double) * n);
    status[i] = info.status; iters[i] = info.iters; evals[i] = info.evals; resid[i] = info.resid;
  }
}

/* ------------------------------------------------------------------------------------- */
/* Counter-based RNG shared (by specification, not by code) with the CUDA path:            */
/* Philox4x32-10, key=(seed_lo,seed_hi), counter=(node_uid, sample, dof/4, stream).         */
/* Replaces the reference's unseeded Python `random` / IKSolver.sampleInitial()             */
/* (solver.py:817-821,881-884; SURVEY K3: uniform in [qMin,qMax] on the active DOFs).       */
/* ------------------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t *out) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
  philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

/* Random start for (node_uid, sample): single-precision arithmetic so that CPU and GPU
 * draw bit-identical seeds: q_j = fmaf(u, (float)qmax-(float)qmin, (float)qmin),
 * u = (r>>8)*2^-24.  Unbounded joints sample [-pi,pi]. */
void orc_random_seed(const orc_chain *c, uint64_t seed, uint32_t node_uid, uint32_t sample, double *q) {
  for (int j = 0; j < c->n; j++) {
    uint32_t r[4];
    philox4x32_10(node_uid, sample, (uint32_t)(j / 4), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    float u = (float)(r[j % 4] >> 8) * (1.0f / 16777216.0f);
    float lo = (float)c->qmin[j], hi = (float)c->qmax[j];
    if (c->qmax[j] - c->qmin[j] >= 2 * M_PI) { lo = -3.14159265358979f; hi = 3.14159265358979f; }
    if (c->spin[j]) { lo = 0.0f; hi = 6.28318530717959f; }    /* [recalled] RobotCSpace sampling of a spin joint */
    q[j] = (double)fmaf(u, hi - lo, lo);
  }
}

/* ------------------------------------------------------------------------------------- */
/* S2/S3: parallelQSamplingAtP (solver.py:875-897): Nq random-start solves + greedy dedup  */
/* ------------------------------------------------------------------------------------- */
/* For each node i: raw per-sample results (q_raw [Nw][Nq][n], status/iters [Nw][Nq]) and the
 * kept list (q_kept [Nw][Nq][n] compacted to the front, keep_idx [Nw][Nq] = sample index of
 * each kept config, count [Nw]).  dedup: keep q iff distance(q,q') >= thresh for all kept q'. */
void orc_sample_nodes(const orc_chain *c, int64_t Nw, const double *params, const int32_t *node_uid,
                      int32_t Nq, uint64_t seed, double dedup_thresh, double *q_raw, int32_t *status,
                      int32_t *iters, int32_t *evals, double *q_kept, int32_t *keep_idx, int32_t *count) {
  const int n = c->n, dw = wdim(c);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < Nw; i++) {
    int kept = 0;
    for (int s = 0; s < Nq; s++) {
      double x[ORC_MAXJ]; orc_ikinfo info;
      orc_random_seed(c, seed, (uint32_t)node_uid[i], (uint32_t)s, x);
      orc_ik_solve(c, params + i * dw, x, &info);
      int64_t slot = i * Nq + s;
      if (q_raw) memcpy(q_raw + slot * n, x, sizeof(double) * n);
      if (status) status[slot] = info.status;
      if (iters) iters[slot] = info.iters;
      if (evals) evals[slot] = info.evals;
      if (info.status != ORC_OK) continue;
      int same = 0;
      for (int k = 0; k < kept; k++)
        if (orc_distance(c, x, q_kept + (i * Nq + k) * n) < dedup_thresh) { same = 1; break; }
      if (!same) {
        memcpy(q_kept + (i * Nq + kept) * n, x, sizeof(double) * n);
        if (keep_idx) keep_idx[i * Nq + kept] = s;
        kept++;
      }
    }
    count[i] = kept;
  }
}


/* ------------------------------------------------------------------------------------- */
/* S1: sampleConfigurationSpace with seedFromAdjacent (solver.py:757-835), the sequential   */
/* Gauss-Seidel sweep over nodes in id order.  Python's unseeded `random` is replaced by   */
/* Philox streams: seeded attempt s of node i draws (neighbour, config) from               */
/* ctr=(uid, sample_offset+s, 0, 1): j = adj[r0 % |adj|], q = qlist_j[r1 % |qlist_j|];      */
/* random attempts use the S2 stream (orc_random_seed).                                     */
/* lower_ptr/lower_idx: CSR of each node's neighbours with a smaller id, in increasing id  */
/* order; degree[i] = len(Gw.neighbors(i)).  q_kept/count are in-out (incremental growth:  */
/* existing configs take part in the dedup; solver.py:759).  cap = slots per node.          */
/* ------------------------------------------------------------------------------------- */
/* biased != 0: solver.py:780-785 -- on a roadmap that already has configs, numRandomSamples of a node is 0 if it has     */
/* no configs, else NConfigsPerPoint * configuration_count / (number_of_nodes_with_configurations * len(qlist)) in     */
/* Python 2 integer arithmetic, configuration_count being the running total of the sweep (incremented by every addNode). */
void orc_sample_seeded_b(const orc_chain *c, int64_t Nw, const double *params, const int32_t *node_uid,
                         const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree, int32_t Nq,
                         uint64_t seed, int32_t sample_offset, int32_t seed_from_adjacent, double dedup_thresh,
                         int32_t cap, double *q_kept, int32_t *count, int64_t *stats /* [4]: adj tries, adj ok, rnd tries, rnd ok */,
                         int32_t biased) {
  const int n = c->n, dw = wdim(c);
  int64_t st_adj = 0, st_adj_ok = 0, st_rnd = 0, st_rnd_ok = 0;
  int64_t configuration_count = 0, nodes_with = 0;                                 /* :759-764 */
  for (int64_t i = 0; i < Nw; i++) { configuration_count += count[i]; nodes_with += (count[i] > 0); }
  const int64_t start_node_count = configuration_count;
  for (int64_t i = 0; i < Nw; i++) {
    int adj[64], na = 0;
    for (int k = lower_ptr[i]; k < lower_ptr[i + 1]; k++)
      if (count[lower_idx[k]] > 0 && na < 64) adj[na++] = lower_idx[k];            /* solver.py:776 */
    const double frac = degree[i] > 0 ? (double)na / (double)degree[i] : 0.0;      /* :777 */
    const int numAdjacent = (int)(frac * Nq);                                      /* :778 */
    int64_t numRandom = Nq - numAdjacent;                                          /* :779 */
    if (biased && start_node_count != 0)                                           /* :780-785 */
      numRandom = count[i] == 0 ? 0 : ((int64_t)Nq * configuration_count) / (nodes_with * (int64_t)count[i]);
    int numFailed = 0, kept = count[i];
    if (seed_from_adjacent) {
      for (int s = 0; s < numAdjacent; s++) {                                      /* :794-815 */
        uint32_t r[4];
        philox4x32_10((uint32_t)node_uid[i], (uint32_t)(sample_offset + s), 0u, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const int j = adj[r[0] % (uint32_t)na];
        const int cfg = (int)(r[1] % (uint32_t)count[j]);
        double x[ORC_MAXJ]; orc_ikinfo info;
        memcpy(x, q_kept + ((int64_t)j * cap + cfg) * n, sizeof(double) * n);
        orc_ik_solve(c, params + i * dw, x, &info);
        st_adj++;
        if (info.status != ORC_OK) { numFailed++; continue; }
        st_adj_ok++;
        int same = 0;
        for (int k = 0; k < kept; k++)
          if (orc_distance(c, x, q_kept + (i * cap + k) * n) < dedup_thresh) { same = 1; break; }
        if (!same && kept < cap) { memcpy(q_kept + (i * cap + kept) * n, x, sizeof(double) * n); kept++; configuration_count++; }
      }
    }
    for (int64_t s = 0; s < numRandom + numFailed; s++) {                          /* :817-835 */
      double x[ORC_MAXJ]; orc_ikinfo info;
      orc_random_seed(c, seed, (uint32_t)node_uid[i], (uint32_t)(sample_offset + s), x);
      orc_ik_solve(c, params + i * dw, x, &info);
      st_rnd++;
      if (info.status != ORC_OK) continue;
      st_rnd_ok++;
      int same = 0;
      for (int k = 0; k < kept; k++)
        if (orc_distance(c, x, q_kept + (i * cap + k) * n) < dedup_thresh) { same = 1; break; }
      if (!same && kept < cap) { memcpy(q_kept + (i * cap + kept) * n, x, sizeof(double) * n); kept++; configuration_count++; }
    }
    count[i] = kept;
  }
  if (stats) { stats[0] = st_adj; stats[1] = st_adj_ok; stats[2] = st_rnd; stats[3] = st_rnd_ok; }
}
void orc_sample_seeded(const orc_chain *c, int64_t Nw, const double *params, const int32_t *node_uid,
                       const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree, int32_t Nq,
                       uint64_t seed, int32_t sample_offset, int32_t seed_from_adjacent, double dedup_thresh,
                       int32_t cap, double *q_kept, int32_t *count, int64_t *stats) {
  orc_sample_seeded_b(c, Nw, params, node_uid, lower_ptr, lower_idx, degree, Nq, seed, sample_offset, seed_from_adjacent,
                      dedup_thresh, cap, q_kept, count, stats, 0);
}

/* ------------------------------------------------------------------------------------- */
/* V4: workspaceInterpolate (graph.py:768-777 -> IKTask.interpolate :143-164)              */
/* ------------------------------------------------------------------------------------- */
void orc_workspace_interpolate(const orc_chain *c, const double *wa, const double *wb, double u, double *w) {
  for (int k = 0; k < 3; k++) w[k] = wa[k] + u * (wb[k] - wa[k]);
  if (c->mode == ORC_VARIABLE) {
    double Ra[9], Rb[9], Rrel[9], m[3], Ru[9], R[9];
    orc_so3_exp(wa + 3, Ra); orc_so3_exp(wb + 3, Rb);
    m3_Tmul(Ra, Rb, Rrel);                 /* so3.interpolate: Ra * exp(u log(Ra^T Rb)) */
    orc_so3_log(Rrel, m);
    m[0] *= u; m[1] *= u; m[2] *= u;
    orc_so3_exp(m, Ru);
    m3_mul(Ra, Ru, R);
    orc_so3_log(R, w + 3);
  }
}

/* ------------------------------------------------------------------------------------- */
/* V1: validEdgeLinear (graph.py:938-1006), literal BFS bisection                          */
/* ------------------------------------------------------------------------------------- */
typedef struct {
  int32_t valid;        /* result                                                          */
  int32_t ndivs;        /* graph.py:946                                                    */
  int32_t n_solves;     /* interpolated IK solves executed before return                   */
  int32_t n_iters;      /* Newton iterations summed                                        */
  int32_t n_evals;      /* merit evaluations summed                                        */
  int32_t fail_kind;    /* 0 none, 1 IK failed, 2 leaf > thr, 3 mid-segment > thr*(ib-ia)  */
  double dist;          /* distance(qa,qb)                                                 */
  double max_leaf;      /* largest leaf distance seen (when all solves ran)                */
  double ndivs_margin;  /* |dist/eps - round(dist/eps)|: closeness of Ndivs to a flip      */
} orc_edgeinfo;

typedef struct { int ia, ib; double q0[ORC_MAXJ], q1[ORC_MAXJ]; } bfs_item;

int orc_valid_edge(const orc_chain *c, const double *qa, const double *qb, const double *wa,
                   const double *wb, double epsilon, orc_edgeinfo *info) {
  const int n = c->n;
  const double eps = epsilon * sqrt((double)c->nlinks_eps);      /* graph.py:945 */
  const double dist = orc_distance(c, qa, qb);
  const int Ndivs = (int)ceil(dist / eps);                        /* graph.py:946 */
  const double thr = eps * c->nlinks_eps;                         /* graph.py:947 */
  int valid = 1, fail = 0, n_solves = 0, n_iters = 0, n_evals = 0;
  double max_leaf = 0;
  const int cap = 2 * (Ndivs + 2) + 4;
  bfs_item *queue = (bfs_item *)malloc(sizeof(bfs_item) * cap);
  int head = 0, tail = 0;
  queue[tail].ia = 0; queue[tail].ib = Ndivs + 1;
  memcpy(queue[tail].q0, qa, sizeof(double) * n); memcpy(queue[tail].q1, qb, sizeof(double) * n);
  tail++;
  while (head < tail) {                                            /* graph.py:952 */
    bfs_item it = queue[head++];
    if (it.ib == it.ia + 1) {                                      /* graph.py:954-959 */
      double d = orc_distance(c, it.q0, it.q1);
      if (d > max_leaf) max_leaf = d;
      if (d > thr) { valid = 0; fail = 2; break; }
      continue;
    }
    int im = (it.ia + it.ib) / 2;                                  /* graph.py:960 */
    double u = (double)im / (double)(Ndivs + 1);
    double x[ORC_MAXJ], w[6]; orc_ikinfo ik;
    interp_q(c, qa, qb, u, x);                                     /* graph.py:962: ORIGINAL endpoints */
    orc_workspace_interpolate(c, wa, wb, u, w);
    orc_ik_solve(c, w, x, &ik);                                    /* graph.py:963 -> :779-790 */
    n_solves++; n_iters += ik.iters; n_evals += ik.evals;
    if (ik.status != ORC_OK) { valid = 0; fail = 1; break; }      /* graph.py:964-966 */
    if (orc_distance(c, it.q0, x) > thr * (it.ib - it.ia)) { valid = 0; fail = 3; break; }
    if (orc_distance(c, x, it.q1) > thr * (it.ib - it.ia)) { valid = 0; fail = 3; break; }
    queue[tail].ia = it.ia; queue[tail].ib = im;
    memcpy(queue[tail].q0, it.q0, sizeof(double) * n); memcpy(queue[tail].q1, x, sizeof(double) * n);
    tail++;
    queue[tail].ia = im; queue[tail].ib = it.ib;
    memcpy(queue[tail].q0, x, sizeof(double) * n); memcpy(queue[tail].q1, it.q1, sizeof(double) * n);
    tail++;
  }
  free(queue);
  if (info) {
    info->valid = valid; info->ndivs = Ndivs; info->n_solves = n_solves; info->n_iters = n_iters;
    info->n_evals = n_evals; info->fail_kind = fail; info->dist = dist; info->max_leaf = max_leaf;
    double r = dist / eps; info->ndivs_margin = fabs(r - floor(r + 0.5));
  }
  return valid;
}

/* Closed form of the same predicate: all Ndivs solves succeed and every
 * consecutive distance <= thr.  No early exit; returns the full diagnostics. */
int orc_valid_edge_closed(const orc_chain *c, const double *qa, const double *qb, const double *wa,
                          const double *wb, double epsilon, orc_edgeinfo *info) {
  const int n = c->n;
  const double eps = epsilon * sqrt((double)c->nlinks_eps);
  const double dist = orc_distance(c, qa, qb);
  const int Ndivs = (int)ceil(dist / eps);
  const double thr = eps * c->nlinks_eps;
  int valid = 1, fail = 0, n_iters = 0, n_evals = 0;
  double prev[ORC_MAXJ], max_leaf = 0;
  memcpy(prev, qa, sizeof(double) * n);
  for (int k = 1; k <= Ndivs; k++) {
    double u = (double)k / (double)(Ndivs + 1), x[ORC_MAXJ], w[6]; orc_ikinfo ik;
    interp_q(c, qa, qb, u, x);
    orc_workspace_interpolate(c, wa, wb, u, w);
    orc_ik_solve(c, w, x, &ik);
    n_iters += ik.iters; n_evals += ik.evals;
    if (ik.status != ORC_OK) { valid = 0; if (!fail) fail = 1; }
    double d = orc_distance(c, prev, x);
    if (d > max_leaf) max_leaf = d;
    memcpy(prev, x, sizeof(double) * n);
  }
  double d = orc_distance(c, prev, qb);
  if (Ndivs > 0 && d > max_leaf) max_leaf = d;
  if (Ndivs == 0) max_leaf = d;
  if (max_leaf > thr) { valid = 0; if (!fail) fail = 2; }
  if (info) {
    info->valid = valid; info->ndivs = Ndivs; info->n_solves = Ndivs; info->n_iters = n_iters;
    info->n_evals = n_evals; info->fail_kind = fail; info->dist = dist; info->max_leaf = max_leaf;
    double r = dist / eps; info->ndivs_margin = fabs(r - floor(r + 0.5));
  }
  return valid;
}

/* Batch validEdge over explicit pairs (backs testAndConnect / 'test-multi', solver.py:484-501,2391-2397). */
void orc_valid_edge_batch(const orc_chain *c, int64_t P, const double *qa, const double *qb,
                          const double *wa, const double *wb, double epsilon, int32_t *valid,
                          orc_edgeinfo *infos) {
  const int n = c->n, dw = wdim(c);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < P; i++) {
    orc_edgeinfo e;
    valid[i] = orc_valid_edge(c, qa + i * n, qb + i * n, wa + i * dw, wb + i * dw, epsilon, &e);
    if (infos) infos[i] = e;
  }
}

/* ------------------------------------------------------------------------------------- */
/* E1/E3: connectConfigurationSpaceOnEdge over every workspace edge                         */
/* (solver.py:689-697 inside :700-755).  q_kept [Nw][Nq][n], count [Nw], old_count optional */
/* Output: mask [E][Nq][Nq] bytes (1 = C-space edge (iq,jq) valid, oriented iw<jw),        */
/* and per-workspace-edge counters stats[E][4] = {pairs tested, valid, ik solves, iters}.  */
/* ------------------------------------------------------------------------------------- */
void orc_connect_edges(const orc_chain *c, int64_t E, const int32_t *wedges, const double *params,
                       const double *q_kept, const int32_t *count, const int32_t *old_count, int32_t Nq,
                       double epsilon, uint8_t *mask, int64_t *stats) {
  const int n = c->n, dw = wdim(c);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t e = 0; e < E; e++) {
    int iw = wedges[2 * e], jw = wedges[2 * e + 1];
    int64_t tested = 0, nvalid = 0, solves = 0, its = 0;
    for (int a = 0; a < count[iw]; a++)
      for (int b = 0; b < count[jw]; b++) {
        if (old_count && a < old_count[iw] && b < old_count[jw]) continue;   /* solver.py:694 */
        orc_edgeinfo info;
        int ok = orc_valid_edge(c, q_kept + ((int64_t)iw * Nq + a) * n, q_kept + ((int64_t)jw * Nq + b) * n,
                                params + (int64_t)iw * dw, params + (int64_t)jw * dw, epsilon, &info);
        tested++; nvalid += ok; solves += info.n_solves; its += info.n_iters;
        if (mask) mask[(e * Nq + a) * Nq + b] = (uint8_t)ok;
      }
    if (stats) { stats[4 * e] = tested; stats[4 * e + 1] = nvalid; stats[4 * e + 2] = solves; stats[4 * e + 3] = its; }
  }
}

int orc_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  omp_set_num_threads(t);
#else
  (void)t;
#endif
}
