/*
 * ns_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference hot path: element residual / Jacobian of the stabilized
 * incompressible Navier-Stokes weak forms, plus dolfinx-semantics global assembly and CSR SpMV.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (libnsgpu.so) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in un-vendored third-party packages
 * (fenics-dolfinx 0.9.0, fenics-ffcx 0.9.0, fenics-basix 0.9.0, fenics-ufl 2024.2.0, petsc 3.23.4;
 * /root/reference/environment.yml:37-44,188-189) that are not installed here and the reference
 * ships no tests or golden vectors for this path.  This file restates the published algorithm
 * (UFL operator semantics, basix degree-2 simplex quadrature, FFCx |detJ| scaling, dolfinx
 * assemble_vector / assemble_matrix / apply_lifting / set_bc semantics) anchored on the
 * reference's own call sites.  It is pinned instead against an independent symbolic (sympy)
 * evaluation of the forms as written in the reference (oracle/symbolic_ref.py ->
 * tests/golden/), finite differences, and patch tests (tests/test_oracle.py).
 *
 * Everything here is written in "direct quadrature" style (loop over points, test and trial
 * functions, one term of the weak form per line) on purpose: it is the slow, obviously-correct
 * form, independent of the factorised algebra the CUDA kernels use.
 *
 * Forms restated (file:line are into /root/reference):
 *   flavour 0  G-metric SUPG/PSPG/LSIC Navier-Stokes  NavierStokes/NavierStokesChannelFlow.py:220-251
 *   flavour 1  UGN (h-based) SUPG/PSPG/LSIC NS         LidDrivenFlow/LidDrivenNavierStokesFlow.py:112-143
 *   flavour 2  Stokes  alpha*grad u:grad v + sp*(-p div v + q div u) + beta*h^2 grad p.grad q
 *              (alpha,sp,beta) = (1,+1,0.2)        NavierStokes/NavierStokesChannelFlow.py:160-172
 *                                (nu,+1,1/(12nu))  LidDrivenFlow/LidDrivenNavierStokesFlow.py:86-99
 *                                (1,-1,0)          StokesFlow/DuctStokesFlow.py:188-192
 * Callback sequence restated: NonlinearPDE_SNESProblem.F/.J  NavierStokesChannelFlow.py:51-75.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 10  /* max scalar nodes per cell (P2 tet) */
#define MAXD 34  /* max mixed dofs per cell (P2-P1 tet) */

typedef struct {
  int flavour;   /* 0 gmetric NS, 1 UGN NS, 2 Stokes */
  int gdim;      /* 2 triangle, 3 tet */
  int vdeg;      /* velocity degree 1|2 (pressure always P1) */
  double nu;     /* viscosity 1/Re                         (NavierStokesChannelFlow.py:223) */
  double Ci;     /* G-metric constant, 36 in the reference (NavierStokesChannelFlow.py:237) */
  double alpha, sp, beta; /* Stokes flavour coefficients */
} form_t;

static int n_scalar_nodes(int gdim, int deg) {
  if (deg == 1) return gdim + 1;
  return gdim == 3 ? 10 : 6;
}
int oracle_ndofs_cell(int gdim, int vdeg) { return gdim * n_scalar_nodes(gdim, vdeg) + gdim + 1; }

/* basix degree-2 default (Xiao-Gimbutas) simplex rules: 4-point tet, 3-point triangle.
 * metadata={'quadrature_degree': 2}  NavierStokesChannelFlow.py:222, LidDrivenNavierStokesFlow.py:92 */
static int quadrature(int gdim, double pts[4][3], double wts[4]) {
  if (gdim == 3) {
    const double a = 0.1381966011250105, b = 0.5854101966249685;
    const double P[4][3] = {{a, a, a}, {b, a, a}, {a, b, a}, {a, a, b}};
    for (int q = 0; q < 4; ++q) { for (int i = 0; i < 3; ++i) pts[q][i] = P[q][i]; wts[q] = 1.0 / 24.0; }
    return 4;
  }
  const double P[3][2] = {{1.0 / 6, 1.0 / 6}, {1.0 / 6, 2.0 / 3}, {2.0 / 3, 1.0 / 6}};
  for (int q = 0; q < 3; ++q) { pts[q][0] = P[q][0]; pts[q][1] = P[q][1]; pts[q][2] = 0; wts[q] = 1.0 / 6.0; }
  return 3;
}

/* basix edge numbering of the reference simplex for the P2 edge dofs. */
static const int TET_EDGES[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
static const int TRI_EDGES[3][2] = {{1, 2}, {0, 2}, {0, 1}};

/* Lagrange basis on the physical cell: values N[n], gradients dN[n][j], Hessians d2N[n][j][k].
 * lam[v] are barycentric coordinates at the point, gl[v][j] their (constant) physical gradients. */
static void lagrange(int gdim, int deg, const double lam[4], double gl[4][3],
                     double N[MAXN], double dN[MAXN][3], double d2N[MAXN][3][3]) {
  int nv = gdim + 1;
  memset(d2N, 0, sizeof(double) * MAXN * 9);
  if (deg == 1) {
    for (int v = 0; v < nv; ++v) {
      N[v] = lam[v];
      for (int j = 0; j < 3; ++j) dN[v][j] = (j < gdim) ? gl[v][j] : 0.0;
    }
    return;
  }
  for (int v = 0; v < nv; ++v) {
    N[v] = lam[v] * (2 * lam[v] - 1);
    for (int j = 0; j < 3; ++j) {
      dN[v][j] = (j < gdim) ? (4 * lam[v] - 1) * gl[v][j] : 0.0;
      for (int k = 0; k < 3; ++k) d2N[v][j][k] = (j < gdim && k < gdim) ? 4 * gl[v][j] * gl[v][k] : 0.0;
    }
  }
  int ne = gdim == 3 ? 6 : 3;
  for (int e = 0; e < ne; ++e) {
    int a = gdim == 3 ? TET_EDGES[e][0] : TRI_EDGES[e][0];
    int b = gdim == 3 ? TET_EDGES[e][1] : TRI_EDGES[e][1];
    int n = nv + e;
    N[n] = 4 * lam[a] * lam[b];
    for (int j = 0; j < 3; ++j) {
      dN[n][j] = (j < gdim) ? 4 * (lam[a] * gl[b][j] + lam[b] * gl[a][j]) : 0.0;
      for (int k = 0; k < 3; ++k)
        d2N[n][j][k] = (j < gdim && k < gdim) ? 4 * (gl[a][j] * gl[b][k] + gl[b][j] * gl[a][k]) : 0.0;
    }
  }
}

/* Affine geometry.  J[i][j] = x_{j+1}[i] - x_0[i]; K = J^-1; returns det J (sign kept; the
 * integration scale is |det J| as in FFCx).  coords are 3-padded (dolfinx coordinate_dofs). */
static double geometry(int gdim, const double *x, double K[3][3], double gl[4][3], double *hdiam) {
  double J[3][3] = {{0}};
  for (int i = 0; i < gdim; ++i)
    for (int j = 0; j < gdim; ++j) J[i][j] = x[3 * (j + 1) + i] - x[i];
  double det;
  memset(K, 0, sizeof(double) * 9);
  if (gdim == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    K[0][0] = J[1][1] / det; K[0][1] = -J[0][1] / det;
    K[1][0] = -J[1][0] / det; K[1][1] = J[0][0] / det;
  } else {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    K[0][0] = c00 / det; K[1][0] = c01 / det; K[2][0] = c02 / det;
    K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  }
  /* barycentric gradients: lam_k = xi_{k-1} (k>=1), lam_0 = 1 - sum xi;  d xi_a / d x_j = K[a][j] */
  for (int j = 0; j < 3; ++j) {
    gl[0][j] = 0;
    for (int a = 0; a < gdim; ++a) { gl[a + 1][j] = K[a][j]; gl[0][j] -= K[a][j]; }
  }
  /* ufl.CellDiameter on an affine simplex = max vertex-pair distance */
  double h2 = 0;
  for (int a = 0; a <= gdim; ++a)
    for (int b = a + 1; b <= gdim; ++b) {
      double d2 = 0;
      for (int i = 0; i < gdim; ++i) { double d = x[3 * a + i] - x[3 * b + i]; d2 += d * d; }
      if (d2 > h2) h2 = d2;
    }
  *hdiam = sqrt(h2);
  return det;
}

/* Element residual be[nd] and Jacobian Ae[nd*nd] (row-major, test x trial).  Either may be NULL.
 * Both are ACCUMULATED INTO (UFCx tabulate_tensor convention: caller zeroes).
 * Mixed cell-local dof order: velocity node-major, components interleaved (gdim*n + c), then
 * pressure nodes (gdim*nvn + n) -- mixed_element([P_k^gdim, P1]) NavierStokesChannelFlow.py:128. */
void oracle_element(const form_t *f, const double *x /*(gdim+1) x 3*/, const double *w, double *Ae, double *be) {
  const int gd = f->gdim, nvn = n_scalar_nodes(gd, f->vdeg), npn = gd + 1;
  const int nd = gd * nvn + npn, poff = gd * nvn;
  double K[3][3], gl[4][3], h;
  double det = geometry(gd, x, K, gl, &h);
  double scale = fabs(det);
  double pts[4][3], wts[4];
  int nq = quadrature(gd, pts, wts);

  /* metric tensor G = K^T K  (dxi_dx = inv(Jacobian) * inv(grad(x)) = K; :232-235) */
  double G[3][3] = {{0}}, trG = 0, GG = 0;
  for (int i = 0; i < gd; ++i)
    for (int j = 0; j < gd; ++j) {
      for (int a = 0; a < gd; ++a) G[i][j] += K[a][i] * K[a][j];
    }
  for (int i = 0; i < gd; ++i) { trG += G[i][i]; for (int j = 0; j < gd; ++j) GG += G[i][j] * G[i][j]; }

  for (int q = 0; q < nq; ++q) {
    double lam[4] = {1, 0, 0, 0};
    for (int a = 0; a < gd; ++a) { lam[a + 1] = pts[q][a]; lam[0] -= pts[q][a]; }
    double Nv[MAXN], dNv[MAXN][3], d2Nv[MAXN][3][3], Np[MAXN], dNp[MAXN][3], d2Np[MAXN][3][3];
    lagrange(gd, f->vdeg, lam, gl, Nv, dNv, d2Nv);
    lagrange(gd, 1, lam, gl, Np, dNp, d2Np);
    const double W = wts[q] * scale;

    /* coefficient w at the point */
    double u[3] = {0}, gu[3][3] = {{0}}, Hu[3][3][3] = {{{0}}}, p = 0, gp[3] = {0};
    for (int n = 0; n < nvn; ++n)
      for (int i = 0; i < gd; ++i) {
        double un = w[gd * n + i];
        u[i] += Nv[n] * un;
        for (int j = 0; j < gd; ++j) {
          gu[i][j] += un * dNv[n][j];             /* grad(u)[i][j] = d u_i / d x_j */
          for (int k = 0; k < gd; ++k) Hu[i][j][k] += un * d2Nv[n][j][k];
        }
      }
    for (int n = 0; n < npn; ++n) { p += Np[n] * w[poff + n]; for (int j = 0; j < gd; ++j) gp[j] += w[poff + n] * dNp[n][j]; }
    double divu = 0;
    for (int i = 0; i < gd; ++i) divu += gu[i][i];
    double conv[3] = {0};                          /* dot(u, nabla_grad(u))_c = u_i d_i u_c */
    for (int c = 0; c < gd; ++c) for (int i = 0; i < gd; ++i) conv[c] += u[i] * gu[c][i];

    if (f->flavour == 2) {
      /* ---- Stokes: alpha grad u:grad v + sp(-p div v + q div u) + beta h^2 grad p.grad q ---- */
      const double muT = f->beta * h * h;
      for (int m = 0; m < nvn; ++m) for (int c = 0; c < gd; ++c) {
        int r = gd * m + c;
        if (be) {
          double s = 0;
          for (int j = 0; j < gd; ++j) s += f->alpha * gu[c][j] * dNv[m][j];
          s -= f->sp * p * dNv[m][c];
          be[r] += W * s;
        }
        if (Ae) {
          for (int n = 0; n < nvn; ++n) {
            double s = 0;
            for (int j = 0; j < gd; ++j) s += dNv[n][j] * dNv[m][j];
            Ae[r * nd + gd * n + c] += W * f->alpha * s;
          }
          for (int n = 0; n < npn; ++n) Ae[r * nd + poff + n] += -W * f->sp * Np[n] * dNv[m][c];
        }
      }
      for (int m = 0; m < npn; ++m) {
        int r = poff + m;
        if (be) {
          double s = f->sp * Np[m] * divu;
          for (int j = 0; j < gd; ++j) s += muT * gp[j] * dNp[m][j];
          be[r] += W * s;
        }
        if (Ae) {
          for (int n = 0; n < nvn; ++n) for (int d = 0; d < gd; ++d) Ae[r * nd + gd * n + d] += W * f->sp * Np[m] * dNv[n][d];
          for (int n = 0; n < npn; ++n) {
            double s = 0;
            for (int j = 0; j < gd; ++j) s += dNp[n][j] * dNp[m][j];
            Ae[r * nd + poff + n] += W * muT * s;
          }
        }
      }
      continue;
    }

    /* ---- stabilisation parameters and momentum residual ---- */
    double tau, nuL;                 /* SUPG/PSPG and LSIC parameters */
    double dtau_du[3], dnuL_du[3];   /* Gateaux derivative coefficients: d tau = dtau_du . du */
    double rM[3];                    /* strong momentum residual used by the stabilisation */
    double Gu[3] = {0};
    if (f->flavour == 0) {
      /* tau_SUPS = 1/sqrt(u.Gu + Ci nu^2 G:G)  (:238);  v_LSIC = 1/(tr(G) tau)  (:249) */
      double uGu = 0;
      for (int i = 0; i < gd; ++i) { for (int j = 0; j < gd; ++j) Gu[i] += G[i][j] * u[j]; uGu += u[i] * Gu[i]; }
      tau = 1.0 / sqrt(uGu + f->Ci * f->nu * f->nu * GG);
      nuL = 1.0 / (trG * tau);
      for (int i = 0; i < gd; ++i) { dtau_du[i] = -tau * tau * tau * Gu[i]; dnuL_du[i] = tau * Gu[i] / trG; }
      /* res_M = dot(u, grad(u)) - div(sigma), sigma = 2 nu sym(grad u) - p I  (:240-241)
       * dot(u, grad(u))_j = u_i d_j u_i  (UFL semantics; NOT (u.grad)u) */
      for (int j = 0; j < gd; ++j) {
        double s = 0, divsig = -gp[j];
        for (int i = 0; i < gd; ++i) s += u[i] * gu[i][j];
        for (int k = 0; k < gd; ++k) divsig += f->nu * (Hu[j][k][k] + Hu[k][j][k]);
        rM[j] = s - divsig;
      }
    } else {
      /* UGN parameters LidDrivenNavierStokesFlow.py:123-134 (r = 2) */
      double uu = 0;
      for (int i = 0; i < gd; ++i) uu += u[i] * u[i];
      double un = sqrt(uu);
      double inv1 = (un <= 1e-8) ? 0.0 : 4.0 * uu / (h * h);
      double tau3 = h * h / (4 * f->nu);
      tau = 1.0 / sqrt(inv1 + 1.0 / (tau3 * tau3));
      double ReU = un * h / (2 * f->nu);
      int low = ReU <= 3.0;
      double z = low ? ReU / 3.0 : 1.0;
      nuL = h / 2 * un * z;
      for (int i = 0; i < gd; ++i) {
        dtau_du[i] = (un <= 1e-8) ? 0.0 : -4.0 * tau * tau * tau * u[i] / (h * h);
        /* d|u| := 0 where |u| == 0 (UFL gives 0/0 there; only BC-overwritten rows/cols see it) */
        double dun = (un > 0) ? u[i] / un : 0.0;
        dnuL_du[i] = h / 2 * (dun * z + un * (low ? dun * h / (2 * f->nu) / 3.0 : 0.0));
      }
      /* res = dot(u, nabla_grad(u)) - nu div(sym(grad u)) + grad p  (:140) */
      for (int j = 0; j < gd; ++j) {
        double dsym = 0;
        for (int k = 0; k < gd; ++k) dsym += 0.5 * (Hu[j][k][k] + Hu[k][j][k]);
        rM[j] = conv[j] - f->nu * dsym + gp[j];
      }
    }

    /* ---- velocity test functions v = N_m e_c ---- */
    for (int m = 0; m < nvn; ++m) for (int c = 0; c < gd; ++c) {
      const int r = gd * m + c;
      /* stabilisation test weight T[j]:  gmetric: dot(u, grad(v))_j = u_c d_j N_m   (:247)
       *                                  UGN:     dot(u, nabla_grad(v))_j = delta_jc (u.dN_m) (:141) */
      double T[3] = {0}, udN = 0;
      for (int j = 0; j < gd; ++j) udN += u[j] * dNv[m][j];
      if (f->flavour == 0) for (int j = 0; j < gd; ++j) T[j] = u[c] * dNv[m][j];
      else T[c] = udN;
      double rT = 0;
      for (int j = 0; j < gd; ++j) rT += rM[j] * T[j];
      if (be) {
        double s = conv[c] * Nv[m];                                   /* inner(dot(u,nabla_grad(u)), v) */
        for (int j = 0; j < gd; ++j) s += f->nu * gu[c][j] * dNv[m][j]; /* nu inner(grad u, grad v) */
        s -= p * dNv[m][c];                                            /* -p div v */
        s += tau * rT;                                                 /* SUPG */
        s += nuL * dNv[m][c] * divu;                                   /* LSIC / grad-div */
        be[r] += W * s;
      }
      if (!Ae) continue;
      for (int n = 0; n < nvn; ++n) for (int d = 0; d < gd; ++d) {
        /* trial du = N_n e_d */
        double udNn = 0;
        for (int j = 0; j < gd; ++j) udNn += u[j] * dNv[n][j];
        double dconv = Nv[n] * gu[c][d] + (c == d ? udNn : 0.0);
        double s = dconv * Nv[m];
        if (c == d) for (int j = 0; j < gd; ++j) s += f->nu * dNv[n][j] * dNv[m][j];
        double dtau = dtau_du[d] * Nv[n], dnuL = dnuL_du[d] * Nv[n];
        double drM[3], dT[3] = {0};
        for (int j = 0; j < gd; ++j) {
          double lap = 0;
          for (int k = 0; k < gd; ++k) lap += d2Nv[n][k][k];
          if (f->flavour == 0) {
            /* d(u_i d_j u_i) - div(d sigma)_j */
            drM[j] = Nv[n] * gu[d][j] + u[d] * dNv[n][j] - f->nu * ((j == d ? lap : 0.0) + d2Nv[n][j][d]);
            dT[j] = (c == d) ? Nv[n] * dNv[m][j] : 0.0;
          } else {
            /* d(u_i d_i u_j) - nu div(sym grad du)_j */
            drM[j] = Nv[n] * gu[j][d] + (j == d ? udNn : 0.0) - f->nu * 0.5 * ((j == d ? lap : 0.0) + d2Nv[n][j][d]);
          }
        }
        if (f->flavour == 1) dT[c] = Nv[n] * dNv[m][d];
        double drT = 0, rdT = 0;
        for (int j = 0; j < gd; ++j) { drT += drM[j] * T[j]; rdT += rM[j] * dT[j]; }
        s += dtau * rT + tau * drT + tau * rdT;
        s += dnuL * dNv[m][c] * divu + nuL * dNv[m][c] * dNv[n][d];
        Ae[r * nd + gd * n + d] += W * s;
      }
      for (int n = 0; n < npn; ++n) {
        /* trial dp = Np_n:  d rM = grad Np_n */
        double s = -Np[n] * dNv[m][c];
        double drT = 0;
        for (int j = 0; j < gd; ++j) drT += dNp[n][j] * T[j];
        s += tau * drT;
        Ae[r * nd + poff + n] += W * s;
      }
    }
    /* ---- pressure test functions q = Np_m ---- */
    for (int m = 0; m < npn; ++m) {
      const int r = poff + m;
      double rT = 0;
      for (int j = 0; j < gd; ++j) rT += rM[j] * dNp[m][j];
      if (be) be[r] += W * (Np[m] * divu + tau * rT);
      if (!Ae) continue;
      for (int n = 0; n < nvn; ++n) for (int d = 0; d < gd; ++d) {
        double udNn = 0, lap = 0, drT = 0;
        for (int j = 0; j < gd; ++j) { udNn += u[j] * dNv[n][j]; lap += d2Nv[n][j][j]; }
        for (int j = 0; j < gd; ++j) {
          double drMj;
          if (f->flavour == 0) drMj = Nv[n] * gu[d][j] + u[d] * dNv[n][j] - f->nu * ((j == d ? lap : 0.0) + d2Nv[n][j][d]);
          else drMj = Nv[n] * gu[j][d] + (j == d ? udNn : 0.0) - f->nu * 0.5 * ((j == d ? lap : 0.0) + d2Nv[n][j][d]);
          drT += drMj * dNp[m][j];
        }
        Ae[r * nd + gd * n + d] += W * (Np[m] * dNv[n][d] + dtau_du[d] * Nv[n] * rT + tau * drT);
      }
      for (int n = 0; n < npn; ++n) {
        double s = 0;
        for (int j = 0; j < gd; ++j) s += dNp[n][j] * dNp[m][j];
        Ae[r * nd + poff + n] += W * tau * s;
      }
    }
  }
}

/* convenience wrapper for ctypes */
void oracle_element_py(int flavour, int gdim, int vdeg, double nu, double Ci, double alpha, double sp, double beta,
                       const double *x, const double *w, double *Ae, double *be) {
  form_t f = {flavour, gdim, vdeg, nu, Ci, alpha, sp, beta};
  oracle_element(&f, x, w, Ae, be);
}

/* ------------------------------------------------------------------------------------------
 * Global assembly with dolfinx semantics (call sites
 * NavierStokesChannelFlow.py:51-75).  All arrays are rank-local, dolfinx layout:
 *   x        n_nodes x 3 (mesh.geometry.x)         cells  n_cells x (gdim+1) (mesh.geometry.dofmap)
 *   dofmap   n_cells x nd, block size 1 (W.dofmap.list), local indices, owned dofs first
 *   bc_marker[dof] != 0  <=> dof constrained by some DirichletBC;  bc_value[dof] = g
 * Only the first n_cells_owned cells are integrated.
 * ------------------------------------------------------------------------------------------ */

static inline int64_t find_col(const int64_t *indptr, const int32_t *indices, int32_t row, int32_t col) {
  int64_t lo = indptr[row], hi = indptr[row + 1] - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t c = indices[mid];
    if (c == col) return mid;
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* F of NavierStokesChannelFlow.py:51-67 minus the MPI ghost updates and set_bc (done by caller):
 *   assemble_vector(F, L);  apply_lifting(F, [a], [bc], [x], -1.0)
 * b has n_owned + n_ghost entries and is accumulated into. */
int oracle_assemble_residual(int flavour, int gdim, int vdeg, double nu, double Ci, double alpha, double sp, double beta,
                             const double *x, const int32_t *cells, const int32_t *dofmap, int64_t n_cells_owned,
                             const double *wvec, const uint8_t *bc_marker, const double *bc_value, int lifting,
                             double *b) {
  form_t f = {flavour, gdim, vdeg, nu, Ci, alpha, sp, beta};
  const int nd = oracle_ndofs_cell(gdim, vdeg), nn = gdim + 1;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n_cells_owned; ++c) {
    double xe[12], we[MAXD], be[MAXD], Ae[MAXD * MAXD];
    for (int a = 0; a < nn; ++a) for (int i = 0; i < 3; ++i) xe[3 * a + i] = x[3 * (int64_t)cells[nn * c + a] + i];
    int has_bc = 0;
    for (int k = 0; k < nd; ++k) { int32_t d = dofmap[nd * c + k]; we[k] = wvec[d]; if (bc_marker && bc_marker[d]) has_bc = 1; }
    memset(be, 0, sizeof(be));
    oracle_element(&f, xe, we, NULL, be);
    if (lifting && has_bc) {
      /* apply_lifting: b_e[i] -= Ae[i][j] * alpha * (g_j - x0_j), alpha = -1, x0 = x (:65); Ae un-zeroed */
      memset(Ae, 0, sizeof(double) * nd * nd);
      oracle_element(&f, xe, we, Ae, NULL);
      for (int j = 0; j < nd; ++j) {
        int32_t dj = dofmap[nd * c + j];
        if (!bc_marker[dj]) continue;
        double delta = bc_value[dj] - wvec[dj];
        for (int i = 0; i < nd; ++i) be[i] += Ae[i * nd + j] * delta;
      }
    }
    for (int k = 0; k < nd; ++k) {
#pragma omp atomic
      b[dofmap[nd * c + k]] += be[k];
    }
  }
  return 0;
}

/* J of NavierStokesChannelFlow.py:69-75: zeroEntries (caller), assemble_matrix(J, a, bcs) incl.
 * the +1.0 diagonal per BC object (bc_mult[dof] = number of BC objects holding the owned dof).
 * vals is accumulated into; pattern (indptr int64 / indices int32, sorted) given. */
int oracle_assemble_jacobian(int flavour, int gdim, int vdeg, double nu, double Ci, double alpha, double sp, double beta,
                             const double *x, const int32_t *cells, const int32_t *dofmap, int64_t n_cells_owned,
                             const double *wvec, const uint8_t *bc_marker, const int32_t *bc_mult, int64_t n_owned,
                             const int64_t *indptr, const int32_t *indices, double *vals) {
  form_t f = {flavour, gdim, vdeg, nu, Ci, alpha, sp, beta};
  const int nd = oracle_ndofs_cell(gdim, vdeg), nn = gdim + 1;
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n_cells_owned; ++c) {
    double xe[12], we[MAXD], Ae[MAXD * MAXD];
    for (int a = 0; a < nn; ++a) for (int i = 0; i < 3; ++i) xe[3 * a + i] = x[3 * (int64_t)cells[nn * c + a] + i];
    for (int k = 0; k < nd; ++k) we[k] = wvec[dofmap[nd * c + k]];
    memset(Ae, 0, sizeof(double) * nd * nd);
    oracle_element(&f, xe, we, Ae, NULL);
    for (int i = 0; i < nd; ++i) {
      int32_t di = dofmap[nd * c + i];
      for (int j = 0; j < nd; ++j) {
        int32_t dj = dofmap[nd * c + j];
        /* BC rows (test dofs) and columns (trial dofs) of Ae are zeroed before insertion */
        double v = (bc_marker && (bc_marker[di] || bc_marker[dj])) ? 0.0 : Ae[i * nd + j];
        int64_t pos = find_col(indptr, indices, di, dj);
        if (pos < 0) { err = 1; continue; }
#pragma omp atomic
        vals[pos] += v;
      }
    }
  }
  if (bc_mult)
    for (int64_t d = 0; d < n_owned; ++d)
      if (bc_mult[d] > 0) {
        int64_t pos = find_col(indptr, indices, (int32_t)d, (int32_t)d);
        if (pos < 0) { err = 1; continue; }
        vals[pos] += (double)bc_mult[d];
      }
  return err;
}

/* PETSc MatMult restated for one rank: y = A x over rows [0, n_rows). */
void oracle_spmv(int64_t n_rows, const int64_t *indptr, const int32_t *indices, const double *vals, const double *xv, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; ++r) {
    double s = 0;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) s += vals[k] * xv[indices[k]];
    y[r] = s;
  }
}

/* CSR sparsity pattern = sorted-unique union over cells of dofs x dofs (create_matrix(problem.a),
 * NavierStokes/NavierStokesChannelFlow.py:272), for meshes where the NumPy set-union of oracle.py is too slow.
 * Row-by-row: the cells incident to a row dof (counting sort of the dofmap), their dofs gathered, sorted, made unique.
 * Call once with indices == NULL (fills indptr, returns nnz), then again with the index array. */
static int cmp_i32(const void *a, const void *b) {
  const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

int64_t oracle_build_pattern(int nd, int64_t n_cells, const int32_t *dofmap, int64_t n_rows, int64_t *indptr, int32_t *indices) {
  int64_t *start = (int64_t *)calloc((size_t)n_rows + 2, sizeof(int64_t));
  int32_t *inc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_cells * nd > 0 ? n_cells * nd : 1));
  if (!start || !inc) { free(start); free(inc); return -1; }
  for (int64_t k = 0; k < n_cells * nd; ++k) start[dofmap[k] + 2] += 1;
  for (int64_t r = 0; r < n_rows; ++r) start[r + 2] += start[r + 1];
  for (int64_t c = 0; c < n_cells; ++c)
    for (int j = 0; j < nd; ++j) inc[start[dofmap[c * nd + j] + 1]++] = (int32_t)c;
  /* now start[r] .. start[r+1] are the cells of row r */
  int bad = 0;
  if (!indices) indptr[0] = 0;
#pragma omp parallel
  {
    int cap = 4096;
    int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
#pragma omp for schedule(dynamic, 4096)
    for (int64_t r = 0; r < n_rows; ++r) {
      const int64_t nc = start[r + 1] - start[r];
      if (nc * nd > cap) {
        cap = (int)(nc * nd) * 2;
        free(buf);
        buf = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
      }
      if (!buf) { bad = 1; continue; }
      int n = 0;
      for (int64_t k = start[r]; k < start[r + 1]; ++k)
        for (int j = 0; j < nd; ++j) buf[n++] = dofmap[(int64_t)inc[k] * nd + j];
      qsort(buf, (size_t)n, sizeof(int32_t), cmp_i32);
      int m = 0;
      for (int k = 0; k < n; ++k)
        if (k == 0 || buf[k] != buf[k - 1]) buf[m++] = buf[k];
      if (!indices) indptr[r + 1] = m;
      else {
        if (indptr[r + 1] - indptr[r] != m) { bad = 1; continue; }
        memcpy(indices + indptr[r], buf, sizeof(int32_t) * (size_t)m);
      }
    }
    free(buf);
  }
  free(start); free(inc);
  if (bad) return -1;
  if (!indices)
    for (int64_t r = 0; r < n_rows; ++r) indptr[r + 1] += indptr[r];
  return indptr[n_rows];
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
