// element_generic.cuh -- one ROW of the element Jacobian (and the matching residual entry) of the
// stabilized Navier-Stokes / Stokes weak forms, for any supported element pair, by direct quadrature.
//
// This is the general path of the product: one thread evaluates one (cell, test dof) pair.  It covers
// every form x element combination of the reference (G-metric NS NavierStokesChannelFlow.py:220-251,
// UGN NS LidDrivenNavierStokesFlow.py:112-143, the three Stokes forms) on P1-P1 / P2-P1 triangles and
// tetrahedra.  The factorised P1-P1 tet kernels in element_p1tet.cuh are the optimised special case.
//
// Semantics: affine geometry, integration scale |det J|, basix degree-2 default
// rule (4-point tet / 3-point triangle) for every form (metadata={'quadrature_degree': 2}), UFL operator
// conventions (dot(u, grad(u))_j = u_i d_j u_i in the G-metric stabilisation, nabla_grad in the
// Galerkin convection), exact Gateaux derivative including d tau / d u.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define NS_HD __host__ __device__ __forceinline__
#else
#define NS_HD inline
#endif

namespace nsgpu {

struct FormParams {
  int flavour;  // 0 G-metric NS, 1 UGN NS, 2 Stokes
  double nu, Ci, alpha, sp, beta;
};

template <int GD, int VDEG>
struct ElemTraits {
  static constexpr int NV = GD + 1;                                   // vertices
  static constexpr int NE = (VDEG == 2) ? (GD == 3 ? 6 : 3) : 0;      // edges carrying dofs
  static constexpr int NVN = NV + NE;                                 // velocity scalar nodes
  static constexpr int NPN = NV;                                      // pressure nodes
  static constexpr int POFF = GD * NVN;
  static constexpr int ND = POFF + NPN;
  static constexpr int NQ = GD + 1;                                   // degree-2 rule
  // row groups: dofs living on one mesh entity share their CSR column set
  static constexpr int NENT = NV + NE;
};

// basix reference-cell edge -> vertices
// (tables {2,1,1,0,0,0} / {3,3,2,3,2,1} and {1,0,0} / {2,2,1} packed into nibbles: a run-time edge index then costs two shifts
// instead of a local-memory array)
template <int GD> NS_HD void edge_vertices(int e, int& a, int& b) {
  if (GD == 3) {
    a = (0x000112 >> (4 * e)) & 15; b = (0x123233 >> (4 * e)) & 15;
  } else {
    a = (0x001 >> (4 * e)) & 15; b = (0x122 >> (4 * e)) & 15;
  }
}

// entity (row group) of a cell-local dof: vertex n for its velocity components and pressure, edge e for P2
template <int GD, int VDEG> NS_HD int entity_of_local_dof(int i) {
  using T = ElemTraits<GD, VDEG>;
  return i < T::POFF ? i / GD : i - T::POFF;
}

template <int GD>
struct CellGeom {
  double gl[GD + 1][GD];  // physical gradients of the barycentric coordinates
  double G[GD][GD];       // metric tensor K^T K
  double trG, GG, h, scale;
};

template <int GD> NS_HD void cell_geometry(const double* x /* (GD+1) x 3 */, CellGeom<GD>& g) {
  double J[GD][GD], K[GD][GD];
  for (int i = 0; i < GD; ++i)
    for (int j = 0; j < GD; ++j) J[i][j] = x[3 * (j + 1) + i] - x[i];
  double det;
  if (GD == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    K[0][0] = J[1][1] * id; K[0][1] = -J[0][1] * id;
    K[1][0] = -J[1][0] * id; K[1][1] = J[0][0] * id;
  } else {
    const double c0 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c1 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c2 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
    const double id = 1.0 / det;
    K[0][0] = c0 * id; K[1][0] = c1 * id; K[2][0] = c2 * id;
    K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  }
  g.scale = fabs(det);
  for (int j = 0; j < GD; ++j) {
    double s = 0;
    for (int a = 0; a < GD; ++a) { g.gl[a + 1][j] = K[a][j]; s -= K[a][j]; }
    g.gl[0][j] = s;
  }
  g.trG = 0; g.GG = 0;
  for (int i = 0; i < GD; ++i)
    for (int j = 0; j < GD; ++j) {
      double s = 0;
      for (int a = 0; a < GD; ++a) s += K[a][i] * K[a][j];
      g.G[i][j] = s;
      g.GG += s * s;
      if (i == j) g.trG += s;
    }
  double h2 = 0;
  for (int a = 0; a <= GD; ++a)
    for (int b = a + 1; b <= GD; ++b) {
      double d2 = 0;
      for (int i = 0; i < GD; ++i) { const double d = x[3 * a + i] - x[3 * b + i]; d2 += d * d; }
      h2 = d2 > h2 ? d2 : h2;
    }
  g.h = sqrt(h2);
}

template <int GD> NS_HD void quad_point(int q, double* lam /* GD+1 */, double& wt) {
  if (GD == 3) {
    const double a = 0.1381966011250105, b = 0.5854101966249685;
    // points (a,a,a),(b,a,a),(a,b,a),(a,a,b) in (xi0,xi1,xi2); lam0 = 1 - sum
    lam[1] = (q == 1) ? b : a; lam[2] = (q == 2) ? b : a; lam[3] = (q == 3) ? b : a;
    lam[0] = 1.0 - lam[1] - lam[2] - lam[3];
    wt = 1.0 / 24.0;
  } else {
    lam[1] = (q == 2) ? 2.0 / 3.0 : 1.0 / 6.0;
    lam[2] = (q == 1) ? 2.0 / 3.0 : 1.0 / 6.0;
    lam[0] = 1.0 - lam[1] - lam[2];
    wt = 1.0 / 6.0;
  }
}

// value / gradient of velocity scalar basis function n at barycentric point lam
template <int GD, int VDEG> NS_HD void vbasis(const CellGeom<GD>& g, const double* lam, int n, double& N, double* dN) {
  if (VDEG == 1) {
    N = lam[n];
    for (int j = 0; j < GD; ++j) dN[j] = g.gl[n][j];
  } else if (n <= GD) {
    N = lam[n] * (2.0 * lam[n] - 1.0);
    const double s = 4.0 * lam[n] - 1.0;
    for (int j = 0; j < GD; ++j) dN[j] = s * g.gl[n][j];
  } else {
    int a, b;
    edge_vertices<GD>(n - GD - 1, a, b);
    N = 4.0 * lam[a] * lam[b];
    for (int j = 0; j < GD; ++j) dN[j] = 4.0 * (lam[a] * g.gl[b][j] + lam[b] * g.gl[a][j]);
  }
}

// constant Hessian of velocity scalar basis function n (zero for P1)
template <int GD, int VDEG> NS_HD double vbasis_d2(const CellGeom<GD>& g, int n, int j, int k) {
  if (VDEG == 1) return 0.0;
  if (n <= GD) return 4.0 * g.gl[n][j] * g.gl[n][k];
  int a, b;
  edge_vertices<GD>(n - GD - 1, a, b);
  return 4.0 * (g.gl[a][j] * g.gl[b][k] + g.gl[b][j] * g.gl[a][k]);
}

// Row `row` of the element Jacobian, accumulated into Arow[ND] when WANT_A, and the residual entry,
// accumulated into *brow when WANT_B.  x: vertex coordinates (3-padded), w: the cell's ND coefficients.
template <int GD, int VDEG, bool WANT_A, bool WANT_B>
NS_HD void element_row(const FormParams& f, const double* x, const double* w, int row, double* Arow, double* brow) {
  using T = ElemTraits<GD, VDEG>;
  CellGeom<GD> g;
  cell_geometry<GD>(x, g);
  const bool vtest = row < T::POFF;
  const int m = vtest ? row / GD : row - T::POFF;
  const int c = vtest ? row % GD : 0;

  // constant second-derivative parts (P2 only): visc[j] = sum_k (d_k d_k u_j + d_k d_j u_k)
  double visc[GD];
  for (int j = 0; j < GD; ++j) visc[j] = 0.0;
  if (VDEG == 2 && f.flavour != 2) {
    for (int n = 0; n < T::NVN; ++n)
      for (int j = 0; j < GD; ++j)
        for (int k = 0; k < GD; ++k)
          visc[j] += w[GD * n + j] * vbasis_d2<GD, VDEG>(g, n, k, k) + w[GD * n + k] * vbasis_d2<GD, VDEG>(g, n, j, k);
  }
  double bsum = 0.0;

  for (int q = 0; q < T::NQ; ++q) {
    double lam[GD + 1], wt;
    quad_point<GD>(q, lam, wt);
    const double W = wt * g.scale;

    double u[GD], gu[GD][GD], p = 0.0, gp[GD];
    for (int i = 0; i < GD; ++i) { u[i] = 0.0; gp[i] = 0.0; for (int j = 0; j < GD; ++j) gu[i][j] = 0.0; }
    for (int n = 0; n < T::NVN; ++n) {
      double N, dN[GD];
      vbasis<GD, VDEG>(g, lam, n, N, dN);
      for (int i = 0; i < GD; ++i) {
        const double un = w[GD * n + i];
        u[i] += N * un;
        for (int j = 0; j < GD; ++j) gu[i][j] += un * dN[j];
      }
    }
    for (int n = 0; n < T::NPN; ++n) {
      const double pn = w[T::POFF + n];
      p += lam[n] * pn;
      for (int j = 0; j < GD; ++j) gp[j] += pn * g.gl[n][j];
    }
    double divu = 0.0;
    for (int i = 0; i < GD; ++i) divu += gu[i][i];

    // test function data
    double Nm = 0.0, dNm[GD];
    if (vtest) vbasis<GD, VDEG>(g, lam, m, Nm, dNm);
    else { Nm = lam[m]; for (int j = 0; j < GD; ++j) dNm[j] = g.gl[m][j]; }

    if (f.flavour == 2) {
      const double muT = f.beta * g.h * g.h;
      if (vtest) {
        if (WANT_B) {
          double s = -f.sp * p * dNm[c];
          for (int j = 0; j < GD; ++j) s += f.alpha * gu[c][j] * dNm[j];
          bsum += W * s;
        }
        if (WANT_A) {
          for (int n = 0; n < T::NVN; ++n) {
            double N, dN[GD], s = 0.0;
            vbasis<GD, VDEG>(g, lam, n, N, dN);
            for (int j = 0; j < GD; ++j) s += dN[j] * dNm[j];
            Arow[GD * n + c] += W * f.alpha * s;
          }
          for (int n = 0; n < T::NPN; ++n) Arow[T::POFF + n] -= W * f.sp * lam[n] * dNm[c];
        }
      } else {
        if (WANT_B) {
          double s = f.sp * Nm * divu;
          for (int j = 0; j < GD; ++j) s += muT * gp[j] * dNm[j];
          bsum += W * s;
        }
        if (WANT_A) {
          for (int n = 0; n < T::NVN; ++n) {
            double N, dN[GD];
            vbasis<GD, VDEG>(g, lam, n, N, dN);
            for (int d = 0; d < GD; ++d) Arow[GD * n + d] += W * f.sp * Nm * dN[d];
          }
          for (int n = 0; n < T::NPN; ++n) {
            double s = 0.0;
            for (int j = 0; j < GD; ++j) s += g.gl[n][j] * dNm[j];
            Arow[T::POFF + n] += W * muT * s;
          }
        }
      }
      continue;
    }

    // ---- stabilisation parameters, their u-derivatives, strong momentum residual ----
    double tau, nuL, dtau[GD], dnuL[GD], rM[GD], conv[GD];
    for (int cc = 0; cc < GD; ++cc) { conv[cc] = 0.0; for (int i = 0; i < GD; ++i) conv[cc] += u[i] * gu[cc][i]; }
    if (f.flavour == 0) {
      double Gu[GD], uGu = 0.0;
      for (int i = 0; i < GD; ++i) { Gu[i] = 0.0; for (int j = 0; j < GD; ++j) Gu[i] += g.G[i][j] * u[j]; uGu += u[i] * Gu[i]; }
      tau = 1.0 / sqrt(uGu + f.Ci * f.nu * f.nu * g.GG);
      nuL = 1.0 / (g.trG * tau);
      for (int i = 0; i < GD; ++i) { dtau[i] = -tau * tau * tau * Gu[i]; dnuL[i] = tau * Gu[i] / g.trG; }
      for (int j = 0; j < GD; ++j) {
        double s = gp[j] - f.nu * visc[j];
        for (int i = 0; i < GD; ++i) s += u[i] * gu[i][j];
        rM[j] = s;
      }
    } else {
      double uu = 0.0;
      for (int i = 0; i < GD; ++i) uu += u[i] * u[i];
      const double un = sqrt(uu), h = g.h;
      const bool still = un <= 1e-8;
      const double inv1 = still ? 0.0 : 4.0 * uu / (h * h);
      const double t3 = h * h / (4.0 * f.nu);
      tau = 1.0 / sqrt(inv1 + 1.0 / (t3 * t3));
      const double ReU = un * h / (2.0 * f.nu);
      const bool low = ReU <= 3.0;
      const double z = low ? ReU / 3.0 : 1.0;
      nuL = 0.5 * h * un * z;
      for (int i = 0; i < GD; ++i) {
        dtau[i] = still ? 0.0 : -4.0 * tau * tau * tau * u[i] / (h * h);
        const double dun = un > 0.0 ? u[i] / un : 0.0;   // d|u| := 0 at |u| = 0
        dnuL[i] = 0.5 * h * (dun * z + un * (low ? dun * h / (6.0 * f.nu) : 0.0));
      }
      for (int j = 0; j < GD; ++j) rM[j] = conv[j] - 0.5 * f.nu * visc[j] + gp[j];
    }

    if (vtest) {
      double udNm = 0.0;
      for (int j = 0; j < GD; ++j) udNm += u[j] * dNm[j];
      double Tt[GD], rT = 0.0;
      for (int j = 0; j < GD; ++j) Tt[j] = (f.flavour == 0) ? u[c] * dNm[j] : (j == c ? udNm : 0.0);
      for (int j = 0; j < GD; ++j) rT += rM[j] * Tt[j];
      if (WANT_B) {
        double s = conv[c] * Nm - p * dNm[c] + tau * rT + nuL * dNm[c] * divu;
        for (int j = 0; j < GD; ++j) s += f.nu * gu[c][j] * dNm[j];
        bsum += W * s;
      }
      if (WANT_A) {
        for (int n = 0; n < T::NVN; ++n) {
          double N, dN[GD], udNn = 0.0, dNdN = 0.0, lap = 0.0;
          vbasis<GD, VDEG>(g, lam, n, N, dN);
          for (int j = 0; j < GD; ++j) { udNn += u[j] * dN[j]; dNdN += dN[j] * dNm[j]; lap += vbasis_d2<GD, VDEG>(g, n, j, j); }
          for (int d = 0; d < GD; ++d) {
            double s = (N * gu[c][d] + (c == d ? udNn : 0.0)) * Nm;
            if (c == d) s += f.nu * dNdN;
            double drT = 0.0, rdT = 0.0;
            for (int j = 0; j < GD; ++j) {
              double drM;
              if (f.flavour == 0)
                drM = N * gu[d][j] + u[d] * dN[j] - f.nu * ((j == d ? lap : 0.0) + vbasis_d2<GD, VDEG>(g, n, j, d));
              else
                drM = N * gu[j][d] + (j == d ? udNn : 0.0) - 0.5 * f.nu * ((j == d ? lap : 0.0) + vbasis_d2<GD, VDEG>(g, n, j, d));
              drT += drM * Tt[j];
            }
            if (f.flavour == 0) { if (c == d) for (int j = 0; j < GD; ++j) rdT += rM[j] * N * dNm[j]; }
            else rdT = rM[c] * N * dNm[d];
            s += dtau[d] * N * rT + tau * (drT + rdT);
            s += dnuL[d] * N * dNm[c] * divu + nuL * dNm[c] * dN[d];
            Arow[GD * n + d] += W * s;
          }
        }
        for (int n = 0; n < T::NPN; ++n) {
          double drT = 0.0;
          for (int j = 0; j < GD; ++j) drT += g.gl[n][j] * Tt[j];
          Arow[T::POFF + n] += W * (tau * drT - lam[n] * dNm[c]);
        }
      }
    } else {
      double rT = 0.0;
      for (int j = 0; j < GD; ++j) rT += rM[j] * dNm[j];
      if (WANT_B) bsum += W * (Nm * divu + tau * rT);
      if (WANT_A) {
        for (int n = 0; n < T::NVN; ++n) {
          double N, dN[GD], udNn = 0.0, lap = 0.0;
          vbasis<GD, VDEG>(g, lam, n, N, dN);
          for (int j = 0; j < GD; ++j) { udNn += u[j] * dN[j]; lap += vbasis_d2<GD, VDEG>(g, n, j, j); }
          for (int d = 0; d < GD; ++d) {
            double drT = 0.0;
            for (int j = 0; j < GD; ++j) {
              double drM;
              if (f.flavour == 0)
                drM = N * gu[d][j] + u[d] * dN[j] - f.nu * ((j == d ? lap : 0.0) + vbasis_d2<GD, VDEG>(g, n, j, d));
              else
                drM = N * gu[j][d] + (j == d ? udNn : 0.0) - 0.5 * f.nu * ((j == d ? lap : 0.0) + vbasis_d2<GD, VDEG>(g, n, j, d));
              drT += drM * dNm[j];
            }
            Arow[GD * n + d] += W * (Nm * dN[d] + dtau[d] * N * rT + tau * drT);
          }
        }
        for (int n = 0; n < T::NPN; ++n) {
          double s = 0.0;
          for (int j = 0; j < GD; ++j) s += g.gl[n][j] * dNm[j];
          Arow[T::POFF + n] += W * tau * s;
        }
      }
    }
  }
  if (WANT_B) *brow += bsum;
}

}  // namespace nsgpu
