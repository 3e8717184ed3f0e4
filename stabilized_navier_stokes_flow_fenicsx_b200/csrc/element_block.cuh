// element_block.cuh -- the element Jacobian of the stabilized forms as ENTITY BLOCKS, factorised.
//
// element_generic.cuh / element_shared.cuh evaluate one ROW of the element matrix by direct quadrature (~225 fp64
// instructions per 3x3 velocity block and point on P2-P1 tets).  The row-owner kernel (rowown.cu) instead gives one lane
// the block (test entity m) x (trial entity n) of a cell: the velocity-velocity GD x GD block, and where m / n are vertices
// the pressure row / column that belongs to them.  Written per block the point contribution factorises into
//     vv[c][d] = (Nm N) gu[c][d] + u[c] Y[d] + dNm[c] Z[d] + delta_cd * s                               (G-metric form)
// with Y, Z, s built from a handful of dot products of the two basis gradients with point data (~60 instructions).
// Same forms and semantics as element_generic.cuh (NavierStokesChannelFlow.py:220-251 G-metric, LidDrivenNavierStokesFlow.py:
// 112-143 UGN, the Stokes flavours), same quadrature; tests/test_element_block.py compares the two entry by entry on the CPU.
//
// Two stages, like element_shared.cuh:
//   point_record  : one call per (cell, quadrature point) -> PREC doubles of point data (fields, stabilisation parameters)
//   cell_record   : one call per cell -> gradients of the barycentric coordinates, cell diameter
//   entity_block  : one call per (cell, m, n) -> the block, summed over the quadrature points
//   entity_rhs    : one call per (cell, m, row of m) -> residual entry
#pragma once
#include "element_shared.cuh"

namespace nsgpu {

constexpr int PREC = 32;   // doubles per (cell, point) record
constexpr int CREC = 16;   // doubles per cell record
// point record slots
constexpr int PR_U = 0, PR_GU = 3, PR_RM = 12, PR_DTAU = 15, PR_DNUL = 18, PR_TAU = 21, PR_NUL = 22, PR_DIVU = 23, PR_W = 24, PR_P = 25, PR_GP = 26,
              PR_ZD = 29;   // div u * d nu_LSIC / d u (the lean G-metric blocks read the product)
// cell record slots: gl[a][j] at 3 a + j, then
constexpr int CR_H = 12;

// barycentric coordinate a of a point given as lam[] without a dynamically indexed (= local-memory) array access
template <int GD> NS_HD double lam_of(const double* lam, int a) {
  if (GD == 3) return a == 0 ? lam[0] : (a == 1 ? lam[1] : (a == 2 ? lam[2] : lam[3]));
  return a == 0 ? lam[0] : (a == 1 ? lam[1] : lam[2]);
}

// The degree-2 rules (quad_point) have exactly one large barycentric coordinate per point: vertex q of the tetrahedron rule
// (0.5854..., the others 0.1381...), vertex (0, 2, 1)[q] of the triangle rule (2/3, the others 1/6).  One compare and one select
// replace the table.  (lam_0 = 1 - sum of the others in quad_point differs from these constants by at most an ulp.)
template <int GD> struct QuadLam {
  int kq;
  NS_HD explicit QuadLam(int q) : kq(GD == 3 ? q : (q == 0 ? 0 : (q == 1 ? 2 : 1))) {}
  NS_HD double operator()(int a) const {
    return GD == 3 ? (a == kq ? 0.5854101966249685 : 0.1381966011250105) : (a == kq ? 2.0 / 3.0 : 1.0 / 6.0);
  }
};

// Same quantities as point_setup (element_shared.cuh) without its basis / derivative tables: everything stays in registers.
template <int GD, int VDEG>
NS_HD void point_record(const FormParams& f, const double* x, const double* w, int q, double* rec, double* crec) {
  using T = ElemTraits<GD, VDEG>;
  CellGeom<GD> g;
  cell_geometry<GD>(x, g);
  double lam[GD + 1], wt;
  quad_point<GD>(q, lam, wt);
  double u[GD], gu[GD][GD], gp[GD], visc[GD], p = 0.0;
  for (int i = 0; i < GD; ++i) { u[i] = 0.0; gp[i] = 0.0; visc[i] = 0.0; for (int j = 0; j < GD; ++j) gu[i][j] = 0.0; }
#pragma unroll
  for (int n = 0; n < T::NVN; ++n) {
    double N, dN[GD];
    vbasis<GD, VDEG>(g, lam, n, N, dN);
    for (int i = 0; i < GD; ++i) {
      const double un = w[GD * n + i];
      u[i] += N * un;
      for (int j = 0; j < GD; ++j) gu[i][j] += un * dN[j];
    }
    if (VDEG == 2 && f.flavour != 2) {   // visc[j] = sum_k (d_k d_k u_j + d_k d_j u_k)
      for (int j = 0; j < GD; ++j)
        for (int k = 0; k < GD; ++k)
          visc[j] += w[GD * n + j] * vbasis_d2<GD, VDEG>(g, n, k, k) + w[GD * n + k] * vbasis_d2<GD, VDEG>(g, n, j, k);
    }
  }
#pragma unroll
  for (int n = 0; n < T::NPN; ++n) {
    const double pn = w[T::POFF + n];
    p += lam[n] * pn;
    for (int j = 0; j < GD; ++j) gp[j] += pn * g.gl[n][j];
  }
  double divu = 0.0;
  for (int i = 0; i < GD; ++i) divu += gu[i][i];
  double tau = 0.0, nuL = 0.0, dtau[GD], dnuL[GD], rM[GD];
  for (int i = 0; i < GD; ++i) { dtau[i] = 0.0; dnuL[i] = 0.0; rM[i] = 0.0; }
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
  } else if (f.flavour == 1) {
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
    for (int j = 0; j < GD; ++j) {
      double conv = 0.0;
      for (int i = 0; i < GD; ++i) conv += u[i] * gu[j][i];
      rM[j] = conv - 0.5 * f.nu * visc[j] + gp[j];
    }
  }
  for (int k = 0; k < PREC; ++k) rec[k] = 0.0;
  for (int i = 0; i < GD; ++i) {
    rec[PR_U + i] = u[i];
    for (int j = 0; j < GD; ++j) rec[PR_GU + 3 * i + j] = gu[i][j];
    rec[PR_RM + i] = rM[i];
    rec[PR_DTAU + i] = dtau[i];
    rec[PR_DNUL + i] = dnuL[i];
    rec[PR_GP + i] = gp[i];
    rec[PR_ZD + i] = divu * dnuL[i];
  }
  rec[PR_TAU] = tau; rec[PR_NUL] = nuL; rec[PR_DIVU] = divu; rec[PR_W] = wt * g.scale; rec[PR_P] = p;
  if (crec) {
    for (int k = 0; k < CREC; ++k) crec[k] = 0.0;
    for (int a = 0; a <= GD; ++a)
      for (int j = 0; j < GD; ++j) crec[3 * a + j] = g.gl[a][j];
    crec[CR_H] = g.h;
  }
}

// velocity basis function k of the cell written as  N,  dN = ca * A + cb * B  with A = gl[a], B = gl[b]
// (vertex k: a = b = k; edge: its two vertices).  kH: Hessian  d2[j][d] = kH (A[j] B[d] + B[j] A[d]),  laplacian = 2 kH A.B
template <int GD, int VDEG>
struct NodeShape {
  int a, b;
  double kH;
  NS_HD void init(int k) {
    if (VDEG == 1 || k <= GD) { a = b = k; kH = (VDEG == 2) ? 2.0 : 0.0; }
    else { edge_vertices<GD>(k - GD - 1, a, b); kH = 4.0; }
  }
  NS_HD void eval(const QuadLam<GD>& lam, int k, double& N, double& ca, double& cb) const {
    const double la = lam(a), lb = lam(b);
    if (VDEG == 1) { N = la; ca = 1.0; cb = 0.0; }
    else {   // selects, not branches: the lanes of a group mix vertex and edge functions
      const bool v = k <= GD;
      N = v ? la * (2.0 * la - 1.0) : 4.0 * la * lb;
      ca = v ? 4.0 * la - 1.0 : 4.0 * lb;
      cb = v ? 0.0 : 4.0 * la;
    }
  }
};

template <int GD>
struct EntityBlock {
  double vv[GD][GD];   // velocity test rows of m  x  velocity trial columns of n
  double pv[GD];       // pressure test row of m (m a vertex)  x  velocity columns of n
  double vp[GD];       // velocity rows of m  x  pressure column of n (n a vertex)
  double pp;           // pressure row of m x pressure column of n
  double b;            // residual entry of row rsel of m (rsel < GD: velocity component, rsel == GD: pressure dof of vertex m)
};

// Block (m, n) of the element Jacobian (WANT_A) and the residual entry of row rsel of m (WANT_B; it shares the row-side
// quantities of the block), all quadrature points.  recs: NQ point records (stride PREC), crec: cell record.
template <int GD, int VDEG, bool WANT_A = true, bool WANT_B = false, int MV = -1 /* test entity: 1 vertex, 0 edge, -1 decided at run time */>
NS_HD void entity_block(const FormParams& f, const double* recs, const double* crec, int m, int n, EntityBlock<GD>& o, int rsel = 0,
                         int rstride = PREC /* doubles between the point records (the kernel pads them in shared memory) */) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int NV = GD + 1;
  const bool m_vertex = MV < 0 ? m < NV : MV != 0, n_vertex = n < NV;
  for (int c = 0; c < GD; ++c) { o.pv[c] = 0.0; o.vp[c] = 0.0; for (int d = 0; d < GD; ++d) o.vv[c][d] = 0.0; }
  o.pp = 0.0; o.b = 0.0;
  NodeShape<GD, VDEG> Sm, Sn;
  Sm.init(m); Sn.init(n);
  double Am[GD], Bm[GD], A[GD], B[GD], glm[GD], gln[GD];
  for (int j = 0; j < GD; ++j) {
    Am[j] = crec[3 * Sm.a + j]; Bm[j] = crec[3 * Sm.b + j];
    A[j] = crec[3 * Sn.a + j]; B[j] = crec[3 * Sn.b + j];
    glm[j] = m_vertex ? crec[3 * m + j] : 0.0;
    gln[j] = n_vertex ? crec[3 * n + j] : 0.0;
  }
  // point-independent pieces
  double AB = 0.0, glnglm = 0.0, sAp = 0.0, sBp = 0.0;
  for (int j = 0; j < GD; ++j) { AB += A[j] * B[j]; glnglm += gln[j] * glm[j]; sAp += A[j] * glm[j]; sBp += B[j] * glm[j]; }
  const double kH = Sn.kH, lap = 2.0 * kH * AB;
  const double h = crec[CR_H];

#pragma unroll 1
  for (int q = 0; q < T::NQ; ++q) {
    const double* R = recs + rstride * q;
    const QuadLam<GD> lam(q);
    const double W = R[PR_W];
    double Nm, cam, cbm, N, ca, cb, dNm[GD], dN[GD];
    Sm.eval(lam, m, Nm, cam, cbm);
    Sn.eval(lam, n, N, ca, cb);
    const double lam_m = m_vertex ? lam(m) : 0.0, lam_n = n_vertex ? lam(n) : 0.0;
    for (int j = 0; j < GD; ++j) { dNm[j] = cam * Am[j] + cbm * Bm[j]; dN[j] = ca * A[j] + cb * B[j]; }
    double dNdN = 0.0;
    for (int j = 0; j < GD; ++j) dNdN += dN[j] * dNm[j];

    if (f.flavour == 2) {
      const double muT = f.beta * h * h;
      if (WANT_A) {
        for (int c = 0; c < GD; ++c) {
          o.vv[c][c] += W * f.alpha * dNdN;
          o.vp[c] -= W * f.sp * lam_n * dNm[c];
          if (m_vertex) o.pv[c] += W * f.sp * lam_m * dN[c];
        }
        if (m_vertex) o.pp += W * muT * glnglm;
      }
      if (WANT_B) {
        double br = 0.0;
        for (int c = 0; c < GD; ++c) {
          double gud = 0.0;
          for (int j = 0; j < GD; ++j) gud += R[PR_GU + 3 * c + j] * dNm[j];
          const double bc = -f.sp * R[PR_P] * dNm[c] + f.alpha * gud;
          br = rsel == c ? bc : br;
        }
        if (m_vertex) {
          double gpg = 0.0;
          for (int j = 0; j < GD; ++j) gpg += R[PR_GP + j] * glm[j];
          br = rsel == GD ? f.sp * lam_m * R[PR_DIVU] + muT * gpg : br;
        }
        o.b += W * br;
      }
      continue;
    }

    double u[GD], rM[GD], dtau[GD], dnuL[GD], gu[GD][GD];
    for (int i = 0; i < GD; ++i) {
      u[i] = R[PR_U + i]; rM[i] = R[PR_RM + i]; dtau[i] = R[PR_DTAU + i]; dnuL[i] = R[PR_DNUL + i];
      for (int j = 0; j < GD; ++j) gu[i][j] = R[PR_GU + 3 * i + j];
    }
    const double tau = R[PR_TAU], nuL = R[PR_NUL], divu = R[PR_DIVU];
    double udNm = 0.0, udNn = 0.0, rdm = 0.0, sA = 0.0, sB = 0.0, glndNm = 0.0, dNglm = 0.0, rTp = 0.0;
    for (int j = 0; j < GD; ++j) {
      udNm += u[j] * dNm[j]; udNn += u[j] * dN[j]; rdm += rM[j] * dNm[j];
      sA += A[j] * dNm[j]; sB += B[j] * dNm[j];
      glndNm += gln[j] * dNm[j]; dNglm += dN[j] * glm[j]; rTp += rM[j] * glm[j];
    }
    if (WANT_B) {   // residual entry of the selected row: shares u . dNm, rM . dNm, ... with the block
      double br = 0.0;
      for (int c = 0; c < GD; ++c) {
        double gud = 0.0, conv = 0.0;
        for (int j = 0; j < GD; ++j) { gud += gu[c][j] * dNm[j]; conv += u[j] * gu[c][j]; }
        const double rT = (f.flavour == 0) ? u[c] * rdm : rM[c] * udNm;
        const double bc = conv * Nm - R[PR_P] * dNm[c] + tau * rT + nuL * dNm[c] * divu + f.nu * gud;
        br = rsel == c ? bc : br;
      }
      if (m_vertex) br = rsel == GD ? lam_m * divu + tau * rTp : br;
      o.b += W * br;
    }
    if (!WANT_A) continue;
    double Z[GD];
    for (int d = 0; d < GD; ++d) Z[d] = dnuL[d] * N * divu + nuL * dN[d];
    const double NmN = Nm * N;

    if (f.flavour == 0) {
      // velocity rows: vv[c][d] = NmN gu[c][d] + u[c] Y[d] + dNm[c] Z[d] + delta_cd s
      double Y[GD];
      for (int d = 0; d < GD; ++d) {
        double gm = 0.0;
        for (int j = 0; j < GD; ++j) gm += gu[d][j] * dNm[j];
        const double HX = (VDEG == 2) ? lap * dNm[d] + kH * (B[d] * sA + A[d] * sB) : 0.0;
        const double X = N * gm + u[d] * dNdN - f.nu * HX;
        Y[d] = dtau[d] * N * rdm + tau * X;
      }
      const double s = Nm * udNn + f.nu * dNdN + tau * N * rdm;
      for (int c = 0; c < GD; ++c) {
        for (int d = 0; d < GD; ++d) o.vv[c][d] += W * (NmN * gu[c][d] + u[c] * Y[d] + dNm[c] * Z[d] + (c == d ? s : 0.0));
        o.vp[c] += W * (tau * u[c] * glndNm - lam_n * dNm[c]);
      }
      if (m_vertex) {
        for (int d = 0; d < GD; ++d) {
          double gm = 0.0;
          for (int j = 0; j < GD; ++j) gm += gu[d][j] * glm[j];
          const double HX = (VDEG == 2) ? lap * glm[d] + kH * (B[d] * sAp + A[d] * sBp) : 0.0;
          const double X = N * gm + u[d] * dNglm - f.nu * HX;
          o.pv[d] += W * (lam_m * dN[d] + dtau[d] * N * rTp + tau * X);
        }
      }
    } else {
      // UGN: Tt[j] = delta_jc (u . dNm)
      const double k1 = NmN + tau * udNm * N;
      const double s = Nm * udNn + f.nu * dNdN + tau * udNm * udNn;
      for (int c = 0; c < GD; ++c) {
        for (int d = 0; d < GD; ++d) {
          const double Hcd = (VDEG == 2) ? (c == d ? lap : 0.0) + kH * (A[c] * B[d] + B[c] * A[d]) : 0.0;
          o.vv[c][d] += W * (gu[c][d] * k1 + (c == d ? s : 0.0) + rM[c] * N * (dtau[d] * udNm + tau * dNm[d]) - 0.5 * f.nu * tau * udNm * Hcd +
                             dNm[c] * Z[d]);
        }
        o.vp[c] += W * (tau * gln[c] * udNm - lam_n * dNm[c]);
      }
      if (m_vertex) {
        for (int d = 0; d < GD; ++d) {
          double gm = 0.0;
          for (int j = 0; j < GD; ++j) gm += gu[j][d] * glm[j];
          const double HX = (VDEG == 2) ? lap * glm[d] + kH * (B[d] * sAp + A[d] * sBp) : 0.0;
          const double X = N * gm + glm[d] * udNn - 0.5 * f.nu * HX;
          o.pv[d] += W * (lam_m * dN[d] + dtau[d] * N * rTp + tau * X);
        }
      }
    }
    if (m_vertex) o.pp += W * tau * glnglm;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Lean G-metric blocks (flavour 0; NavierStokesChannelFlow.py:220-251): the point contribution to block (m, n) is split into
// a ROW-SIDE record of (m, q) -- everything that does not depend on the trial entity, evaluated once per (cell, m, q) and
// shared by the lanes of the row-owner group -- and a mixed part of ~90 instead of ~225 fp64 instructions per (m, n, q):
//   W vv[c][d] = (N a0) gu[c][d] + u[c] (N aV[d] + k2 u[d] + h1 HXa[d] + h2 HXb[d]) + aD[c] (N ZD[d] + nuL dN[d]) + delta_cd s
//   a0 = W Nm, a1 = W tau, a2 = W tau (rM . dNm), aV = W ((rM . dNm) dtau + tau gu dNm), aD = W dNm, h1|h2 = -nu W tau cam|cbm,
//   k2 = tau (dN . aD), s = a0 (u . dN) + nu (dN . aD) + N a2;  HXa | HXb: the Hessian of N_n applied to the two gradient
//   vectors dNm is a combination of (point independent, VDEG = 2 only).
// The residual entries of m need the row side only: they leave gm_row_side with the record.
constexpr int RSIDE = 16;   // doubles per (entity, point) row-side record
constexpr int RS_A0 = 0, RS_A1 = 1, RS_A2 = 2, RS_H1 = 3, RS_H2 = 4, RS_WL = 5, RS_AV = 6, RS_AD = 9, RS_AP = 12;

// MV: m is a vertex (pressure row).  b (nullable): GD (+ 1) residual contributions of this point, overwritten.
template <int GD, int VDEG, bool MV>
NS_HD void gm_row_side(const FormParams& f, const double* R, const double* crec, int m, int q, double* S, double* b) {
  NodeShape<GD, VDEG> Sm;
  Sm.init(m);
  const QuadLam<GD> lam(q);
  double Nm, cam, cbm, dNm[GD];
  Sm.eval(lam, m, Nm, cam, cbm);
  for (int j = 0; j < GD; ++j) dNm[j] = cam * crec[3 * Sm.a + j] + cbm * crec[3 * Sm.b + j];
  const double W = R[PR_W], tau = R[PR_TAU];
  double rdm = 0.0, gm[GD];
  for (int j = 0; j < GD; ++j) rdm += R[PR_RM + j] * dNm[j];
  for (int d = 0; d < GD; ++d) {
    double g = 0.0;
    for (int j = 0; j < GD; ++j) g += R[PR_GU + 3 * d + j] * dNm[j];
    gm[d] = g;
  }
  const double Wt = W * tau, Wr = W * rdm, a0 = W * Nm, a2 = Wt * rdm;
  S[RS_A0] = a0; S[RS_A1] = Wt; S[RS_A2] = a2; S[RS_H1] = -f.nu * Wt * cam; S[RS_H2] = -f.nu * Wt * cbm;
  double aD[GD];
  for (int d = 0; d < GD; ++d) {
    S[RS_AV + d] = Wr * R[PR_DTAU + d] + Wt * gm[d];
    aD[d] = W * dNm[d];
    S[RS_AD + d] = aD[d];
  }
  double rTp = 0.0, wl = 0.0;
  if (MV) {
    double glm[GD];
    for (int j = 0; j < GD; ++j) { glm[j] = crec[3 * m + j]; rTp += R[PR_RM + j] * glm[j]; }
    wl = W * lam(m);
    const double Wp = W * rTp;
    for (int d = 0; d < GD; ++d) {
      double g = 0.0;
      for (int j = 0; j < GD; ++j) g += R[PR_GU + 3 * d + j] * glm[j];
      S[RS_AP + d] = Wp * R[PR_DTAU + d] + Wt * g;
    }
  }
  S[RS_WL] = wl;
  if (b) {
    const double divu = R[PR_DIVU], kk = R[PR_NUL] * divu - R[PR_P], nuW = f.nu * W;
    for (int c = 0; c < GD; ++c) {
      double conv = 0.0;
      for (int j = 0; j < GD; ++j) conv += R[PR_U + j] * R[PR_GU + 3 * c + j];
      b[c] = a0 * conv + aD[c] * kk + a2 * R[PR_U + c] + nuW * gm[c];
    }
    if (MV) b[GD] = wl * divu + Wt * rTp;
  }
}

// block (m, n) from the NQ row-side records of m (rs, stride RSIDE) and the point records; o.b is not touched
template <int GD, int VDEG, bool MV>
NS_HD void gm_block(const FormParams& f, const double* recs, const double* crec, const double* rs, int m, int n, EntityBlock<GD>& o,
                     int rstride = PREC, int sstride = RSIDE /* doubles between the point / row-side records */) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int NV = GD + 1;
  const bool n_vertex = n < NV;
  NodeShape<GD, VDEG> Sm, Sn;
  Sm.init(m); Sn.init(n);
  double A[GD], B[GD], gln[GD], glm[GD], HXa[GD], HXb[GD], HXp[GD];
  for (int j = 0; j < GD; ++j) {
    A[j] = crec[3 * Sn.a + j]; B[j] = crec[3 * Sn.b + j];
    gln[j] = n_vertex ? crec[3 * n + j] : 0.0;
    glm[j] = MV ? crec[3 * m + j] : 0.0;
    HXa[j] = HXb[j] = HXp[j] = 0.0;
  }
  double glnglm = 0.0;
  for (int j = 0; j < GD; ++j) glnglm += gln[j] * glm[j];
  if (VDEG == 2) {
    double AB = 0.0, AAm = 0.0, BAm = 0.0, ABm = 0.0, BBm = 0.0, sAp = 0.0, sBp = 0.0;
    for (int j = 0; j < GD; ++j) {
      const double am = crec[3 * Sm.a + j], bm = crec[3 * Sm.b + j];
      AB += A[j] * B[j];
      AAm += A[j] * am; BAm += B[j] * am; ABm += A[j] * bm; BBm += B[j] * bm;
      sAp += A[j] * glm[j]; sBp += B[j] * glm[j];
    }
    const double kH = Sn.kH, lap = 2.0 * kH * AB;
    for (int d = 0; d < GD; ++d) {
      HXa[d] = lap * crec[3 * Sm.a + d] + kH * (B[d] * AAm + A[d] * BAm);
      HXb[d] = lap * crec[3 * Sm.b + d] + kH * (B[d] * ABm + A[d] * BBm);
      if (MV) HXp[d] = lap * glm[d] + kH * (B[d] * sAp + A[d] * sBp);
    }
  }
  for (int c = 0; c < GD; ++c) { o.pv[c] = 0.0; o.vp[c] = 0.0; for (int d = 0; d < GD; ++d) o.vv[c][d] = 0.0; }
  o.pp = 0.0;
  double tbar = 0.0;
#pragma unroll 1
  for (int q = 0; q < T::NQ; ++q) {
    const double* R = recs + rstride * q;
    const double* S = rs + sstride * q;
    const QuadLam<GD> lam(q);
    double N, ca, cb, dN[GD], u[GD];
    Sn.eval(lam, n, N, ca, cb);
    const double lam_n = n_vertex ? lam(n) : 0.0;
    double wdd = 0.0, udn = 0.0, wgl = 0.0, dgl = 0.0;
    for (int j = 0; j < GD; ++j) {
      dN[j] = ca * A[j] + cb * B[j];
      u[j] = R[PR_U + j];
      wdd += dN[j] * S[RS_AD + j]; udn += u[j] * dN[j]; wgl += gln[j] * S[RS_AD + j];
      if (MV) dgl += dN[j] * glm[j];
    }
    const double tau = R[PR_TAU], nuL = R[PR_NUL], a0 = S[RS_A0], a1 = S[RS_A1];
    const double k0 = N * a0, k2 = tau * wdd, kv = tau * wgl, kp = a1 * dgl;
    const double s = a0 * udn + f.nu * wdd + N * S[RS_A2];
    double t[GD], Zd[GD];
    for (int d = 0; d < GD; ++d) {
      double td = N * S[RS_AV + d] + k2 * u[d];
      if (VDEG == 2) td += S[RS_H1] * HXa[d] + S[RS_H2] * HXb[d];
      t[d] = td;
      Zd[d] = N * R[PR_ZD + d] + nuL * dN[d];
    }
    for (int c = 0; c < GD; ++c) {
      const double aDc = S[RS_AD + c];
      for (int d = 0; d < GD; ++d) o.vv[c][d] += k0 * R[PR_GU + 3 * c + d] + u[c] * t[d] + aDc * Zd[d];
      o.vv[c][c] += s;
      o.vp[c] += kv * u[c] - lam_n * aDc;
      if (MV) o.pv[c] += S[RS_WL] * dN[c] + N * S[RS_AP + c] + kp * u[c];
    }
    tbar += a1;
  }
  if (MV) {
    if (VDEG == 2) for (int d = 0; d < GD; ++d) o.pv[d] -= f.nu * tbar * HXp[d];
    o.pp = tbar * glnglm;
  }
}

// Residual entry of row r of entity m (r < GD: velocity component r; r == GD: the pressure dof of vertex m), all points.
template <int GD, int VDEG>
NS_HD double entity_rhs(const FormParams& f, const double* recs, const double* crec, int m, int r) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int NV = GD + 1;
  NodeShape<GD, VDEG> Sm;
  Sm.init(m);
  double Am[GD], Bm[GD], glm[GD];
  for (int j = 0; j < GD; ++j) { Am[j] = crec[3 * Sm.a + j]; Bm[j] = crec[3 * Sm.b + j]; glm[j] = m < NV ? crec[3 * m + j] : 0.0; }
  const double h = crec[CR_H];
  double b = 0.0;
#pragma unroll 1
  for (int q = 0; q < T::NQ; ++q) {
    const double* R = recs + PREC * q;
    const QuadLam<GD> lam(q);
    const double W = R[PR_W], tau = R[PR_TAU], nuL = R[PR_NUL], divu = R[PR_DIVU], p = R[PR_P];
    if (r == GD) {   // pressure row
      double rT = 0.0, gpg = 0.0;
      for (int j = 0; j < GD; ++j) { rT += R[PR_RM + j] * glm[j]; gpg += R[PR_GP + j] * glm[j]; }
      const double lam_m = lam(m);
      if (f.flavour == 2) b += W * (f.sp * lam_m * divu + f.beta * h * h * gpg);
      else b += W * (lam_m * divu + tau * rT);
      continue;
    }
    double Nm, cam, cbm, dNm[GD];
    Sm.eval(lam, m, Nm, cam, cbm);
    for (int j = 0; j < GD; ++j) dNm[j] = cam * Am[j] + cbm * Bm[j];
    double gud = 0.0, conv = 0.0, udNm = 0.0, rdm = 0.0;
    for (int j = 0; j < GD; ++j) {
      gud += R[PR_GU + 3 * r + j] * dNm[j];
      conv += R[PR_U + j] * R[PR_GU + 3 * r + j];
      udNm += R[PR_U + j] * dNm[j];
      rdm += R[PR_RM + j] * dNm[j];
    }
    if (f.flavour == 2) { b += W * (-f.sp * p * dNm[r] + f.alpha * gud); continue; }
    const double rT = (f.flavour == 0) ? R[PR_U + r] * rdm : R[PR_RM + r] * udNm;
    b += W * (conv * Nm - p * dNm[r] + tau * rT + nuL * dNm[r] * divu + f.nu * gud);
  }
  return b;
}

}  // namespace nsgpu
