// element_shared.cuh -- the same direct-quadrature element tensors as element_generic.cuh, split in two stages so
// that the quadrature-point data of a cell is computed ONCE and shared by the cell's row threads:
//   stage 1  point_setup : one call per (cell, quadrature point)  -> PointData (basis tables, fields, stabilisation)
//   stage 2  row_from_points : one call per (cell, test dof)       -> one row of the element Jacobian + residual entry
// element_generic.cuh redoes stage 1 inside every row (34x redundant on P2-P1 tets); the cooperative kernel in
// assemble.cu keeps the PointData of a cell batch in shared memory instead.  Forms / semantics identical
// (G-metric NS NavierStokesChannelFlow.py:220-251, UGN NS LidDrivenNavierStokesFlow.py:112-143, Stokes flavours).
#pragma once
#include "element_generic.cuh"

namespace nsgpu {

template <int GD, int VDEG>
struct PointData {
  using T = ElemTraits<GD, VDEG>;
  double N[T::NVN];          // velocity scalar basis at the point
  double dN[T::NVN][GD];     // ... and its physical gradient
  double lam[GD + 1];        // pressure (P1) basis = barycentric coordinates
  double u[GD], gu[GD][GD], p, gp[GD], divu, conv[GD];
  double tau, nuL, dtau[GD], dnuL[GD], rM[GD];
  double W;                  // quadrature weight * |det J|
  // d rM_j / d u_(n,d): derivative of the strong momentum residual w.r.t. velocity dof (node n, component d) -- independent
  // of the test row, so it is tabulated once per point instead of being rebuilt for each of the ND rows (NS flavours only)
  // (P2 only: with P1 the expression has no Hessian part and is cheaper to form on the fly than to park in shared memory)
  double drM[VDEG == 2 ? T::NVN : 1][GD][GD];
};

template <int GD>
struct CellData {
  double gl[GD + 1][GD];     // gradients of the barycentric coordinates
  double h;                  // CellDiameter
};

// stage 1 (call with the cell's vertex coordinates and coefficients; q = quadrature point index)
template <int GD, int VDEG>
NS_HD void point_setup(const FormParams& f, const double* x, const double* w, int q, PointData<GD, VDEG>& P, CellData<GD>& C) {
  using T = ElemTraits<GD, VDEG>;
  CellGeom<GD> g;
  cell_geometry<GD>(x, g);
  for (int a = 0; a <= GD; ++a)
    for (int j = 0; j < GD; ++j) C.gl[a][j] = g.gl[a][j];
  C.h = g.h;
  double wt;
  quad_point<GD>(q, P.lam, wt);
  P.W = wt * g.scale;
  for (int i = 0; i < GD; ++i) { P.u[i] = 0.0; P.gp[i] = 0.0; for (int j = 0; j < GD; ++j) P.gu[i][j] = 0.0; }
  P.p = 0.0;
  for (int n = 0; n < T::NVN; ++n) {
    vbasis<GD, VDEG>(g, P.lam, n, P.N[n], P.dN[n]);
    for (int i = 0; i < GD; ++i) {
      const double un = w[GD * n + i];
      P.u[i] += P.N[n] * un;
      for (int j = 0; j < GD; ++j) P.gu[i][j] += un * P.dN[n][j];
    }
  }
  for (int n = 0; n < T::NPN; ++n) {
    const double pn = w[T::POFF + n];
    P.p += P.lam[n] * pn;
    for (int j = 0; j < GD; ++j) P.gp[j] += pn * g.gl[n][j];
  }
  P.divu = 0.0;
  for (int i = 0; i < GD; ++i) P.divu += P.gu[i][i];
  for (int c = 0; c < GD; ++c) { P.conv[c] = 0.0; for (int i = 0; i < GD; ++i) P.conv[c] += P.u[i] * P.gu[c][i]; }
  P.tau = 0.0; P.nuL = 0.0;
  for (int i = 0; i < GD; ++i) { P.dtau[i] = 0.0; P.dnuL[i] = 0.0; P.rM[i] = 0.0; }
  if (f.flavour == 2) return;

  // constant second-derivative part (P2 only): visc[j] = sum_k (d_k d_k u_j + d_k d_j u_k)
  double visc[GD];
  for (int j = 0; j < GD; ++j) visc[j] = 0.0;
  if (VDEG == 2) {
    for (int n = 0; n < T::NVN; ++n)
      for (int j = 0; j < GD; ++j)
        for (int k = 0; k < GD; ++k)
          visc[j] += w[GD * n + j] * vbasis_d2<GD, VDEG>(g, n, k, k) + w[GD * n + k] * vbasis_d2<GD, VDEG>(g, n, j, k);
  }
  if (f.flavour == 0) {
    double Gu[GD], uGu = 0.0;
    for (int i = 0; i < GD; ++i) { Gu[i] = 0.0; for (int j = 0; j < GD; ++j) Gu[i] += g.G[i][j] * P.u[j]; uGu += P.u[i] * Gu[i]; }
    P.tau = 1.0 / sqrt(uGu + f.Ci * f.nu * f.nu * g.GG);
    P.nuL = 1.0 / (g.trG * P.tau);
    for (int i = 0; i < GD; ++i) { P.dtau[i] = -P.tau * P.tau * P.tau * Gu[i]; P.dnuL[i] = P.tau * Gu[i] / g.trG; }
    for (int j = 0; j < GD; ++j) {
      double s = P.gp[j] - f.nu * visc[j];
      for (int i = 0; i < GD; ++i) s += P.u[i] * P.gu[i][j];
      P.rM[j] = s;
    }
  } else {
    double uu = 0.0;
    for (int i = 0; i < GD; ++i) uu += P.u[i] * P.u[i];
    const double un = sqrt(uu), h = g.h;
    const bool still = un <= 1e-8;
    const double inv1 = still ? 0.0 : 4.0 * uu / (h * h);
    const double t3 = h * h / (4.0 * f.nu);
    P.tau = 1.0 / sqrt(inv1 + 1.0 / (t3 * t3));
    const double ReU = un * h / (2.0 * f.nu);
    const bool low = ReU <= 3.0;
    const double z = low ? ReU / 3.0 : 1.0;
    P.nuL = 0.5 * h * un * z;
    for (int i = 0; i < GD; ++i) {
      P.dtau[i] = still ? 0.0 : -4.0 * P.tau * P.tau * P.tau * P.u[i] / (h * h);
      const double dun = un > 0.0 ? P.u[i] / un : 0.0;   // d|u| := 0 at |u| = 0
      P.dnuL[i] = 0.5 * h * (dun * z + un * (low ? dun * h / (6.0 * f.nu) : 0.0));
    }
    for (int j = 0; j < GD; ++j) P.rM[j] = P.conv[j] - 0.5 * f.nu * visc[j] + P.gp[j];
  }
  // row-independent derivative table
  if (VDEG == 2)
  for (int n = 0; n < T::NVN; ++n) {
    double udNn = 0.0, lap = 0.0;
    for (int j = 0; j < GD; ++j) { udNn += P.u[j] * P.dN[n][j]; lap += vbasis_d2<GD, VDEG>(g, n, j, j); }
    for (int d = 0; d < GD; ++d)
      for (int j = 0; j < GD; ++j) {
        const double h2 = (j == d ? lap : 0.0) + vbasis_d2<GD, VDEG>(g, n, j, d);
        P.drM[n][d][j] = (f.flavour == 0) ? P.N[n] * P.gu[d][j] + P.u[d] * P.dN[n][j] - f.nu * h2
                                          : P.N[n] * P.gu[j][d] + (j == d ? udNn : 0.0) - 0.5 * f.nu * h2;
      }
  }
}

// constant Hessian entries of velocity basis n from the barycentric gradients
template <int GD, int VDEG> NS_HD double d2_from_gl(const CellData<GD>& C, int n, int j, int k) {
  if (VDEG == 1) return 0.0;
  if (n <= GD) return 4.0 * C.gl[n][j] * C.gl[n][k];
  int a, b;
  edge_vertices<GD>(n - GD - 1, a, b);
  return 4.0 * (C.gl[a][j] * C.gl[b][k] + C.gl[b][j] * C.gl[a][k]);
}

// d rM_j / d u_(n,d) (see PointData::drM)
template <int GD, int VDEG>
NS_HD double drM_entry(const FormParams& f, const PointData<GD, VDEG>& P, int n, int d, int j, double udNn) {
  if (VDEG == 2) return P.drM[n][d][j];
  return (f.flavour == 0) ? P.N[n] * P.gu[d][j] + P.u[d] * P.dN[n][j] : P.N[n] * P.gu[j][d] + (j == d ? udNn : 0.0);
}

// stage 2: contribution of ONE quadrature point to row `row` (accumulated into Arow[ND] when WANT_A, *brow when WANT_B)
template <int GD, int VDEG, bool WANT_A, bool WANT_B>
NS_HD void row_from_point(const FormParams& f, const PointData<GD, VDEG>& P, const CellData<GD>& C, int row, double* Arow, double* brow) {
  using T = ElemTraits<GD, VDEG>;
  const bool vtest = row < T::POFF;
  const int m = vtest ? row / GD : row - T::POFF;
  const int c = vtest ? row % GD : 0;
  const double W = P.W;
  double Nm, dNm[GD];
  if (vtest) { Nm = P.N[m]; for (int j = 0; j < GD; ++j) dNm[j] = P.dN[m][j]; }
  else { Nm = P.lam[m]; for (int j = 0; j < GD; ++j) dNm[j] = C.gl[m][j]; }

  if (f.flavour == 2) {
    const double muT = f.beta * C.h * C.h;
    if (vtest) {
      if (WANT_B) {
        double s = -f.sp * P.p * dNm[c];
        for (int j = 0; j < GD; ++j) s += f.alpha * P.gu[c][j] * dNm[j];
        *brow += W * s;
      }
      if (WANT_A) {
#pragma unroll
        for (int n = 0; n < T::NVN; ++n) {
          double s = 0.0;
          for (int j = 0; j < GD; ++j) s += P.dN[n][j] * dNm[j];
#pragma unroll
          for (int d = 0; d < GD; ++d)
            if (d == c) Arow[GD * n + d] += W * f.alpha * s;     // static register index, runtime predicate
        }
#pragma unroll
        for (int n = 0; n < T::NPN; ++n) Arow[T::POFF + n] -= W * f.sp * P.lam[n] * dNm[c];
      }
    } else {
      if (WANT_B) {
        double s = f.sp * Nm * P.divu;
        for (int j = 0; j < GD; ++j) s += muT * P.gp[j] * dNm[j];
        *brow += W * s;
      }
      if (WANT_A) {
#pragma unroll
        for (int n = 0; n < T::NVN; ++n)
#pragma unroll
          for (int d = 0; d < GD; ++d) Arow[GD * n + d] += W * f.sp * Nm * P.dN[n][d];
#pragma unroll
        for (int n = 0; n < T::NPN; ++n) {
          double s = 0.0;
          for (int j = 0; j < GD; ++j) s += C.gl[n][j] * dNm[j];
          Arow[T::POFF + n] += W * muT * s;
        }
      }
    }
    return;
  }

  const double tau = P.tau, nuL = P.nuL;
  if (vtest) {
    double udNm = 0.0;
    for (int j = 0; j < GD; ++j) udNm += P.u[j] * dNm[j];
    double Tt[GD], rT = 0.0;
    for (int j = 0; j < GD; ++j) Tt[j] = (f.flavour == 0) ? P.u[c] * dNm[j] : (j == c ? udNm : 0.0);
    for (int j = 0; j < GD; ++j) rT += P.rM[j] * Tt[j];
    if (WANT_B) {
      double s = P.conv[c] * Nm - P.p * dNm[c] + tau * rT + nuL * dNm[c] * P.divu;
      for (int j = 0; j < GD; ++j) s += f.nu * P.gu[c][j] * dNm[j];
      *brow += W * s;
    }
    if (WANT_A) {
#pragma unroll
      for (int n = 0; n < T::NVN; ++n) {
        const double N = P.N[n];
        double udNn = 0.0, dNdN = 0.0;
        for (int j = 0; j < GD; ++j) { udNn += P.u[j] * P.dN[n][j]; dNdN += P.dN[n][j] * dNm[j]; }
#pragma unroll
        for (int d = 0; d < GD; ++d) {
          double s = (N * P.gu[c][d] + (c == d ? udNn : 0.0)) * Nm;
          if (c == d) s += f.nu * dNdN;
          double drT = 0.0, rdT = 0.0;
          for (int j = 0; j < GD; ++j) drT += drM_entry<GD, VDEG>(f, P, n, d, j, udNn) * Tt[j];
          if (f.flavour == 0) { if (c == d) for (int j = 0; j < GD; ++j) rdT += P.rM[j] * N * dNm[j]; }
          else rdT = P.rM[c] * N * dNm[d];
          s += P.dtau[d] * N * rT + tau * (drT + rdT);
          s += P.dnuL[d] * N * dNm[c] * P.divu + nuL * dNm[c] * P.dN[n][d];
          Arow[GD * n + d] += W * s;
        }
      }
#pragma unroll
      for (int n = 0; n < T::NPN; ++n) {
        double drT = 0.0;
        for (int j = 0; j < GD; ++j) drT += C.gl[n][j] * Tt[j];
        Arow[T::POFF + n] += W * (tau * drT - P.lam[n] * dNm[c]);
      }
    }
  } else {
    double rT = 0.0;
    for (int j = 0; j < GD; ++j) rT += P.rM[j] * dNm[j];
    if (WANT_B) *brow += W * (Nm * P.divu + tau * rT);
    if (WANT_A) {
#pragma unroll
      for (int n = 0; n < T::NVN; ++n) {
        const double N = P.N[n];
        double udNn = 0.0;
        if (VDEG == 1) for (int j = 0; j < GD; ++j) udNn += P.u[j] * P.dN[n][j];
#pragma unroll
        for (int d = 0; d < GD; ++d) {
          double drT = 0.0;
          for (int j = 0; j < GD; ++j) drT += drM_entry<GD, VDEG>(f, P, n, d, j, udNn) * dNm[j];
          Arow[GD * n + d] += W * (Nm * P.dN[n][d] + P.dtau[d] * N * rT + tau * drT);
        }
      }
#pragma unroll
      for (int n = 0; n < T::NPN; ++n) {
        double s = 0.0;
        for (int j = 0; j < GD; ++j) s += C.gl[n][j] * dNm[j];
        Arow[T::POFF + n] += W * tau * s;
      }
    }
  }
}

}  // namespace nsgpu
