// Test harness only: the factorised entity blocks of csrc/element_block.cuh (what the row-owner GPU kernel evaluates)
// and the direct-quadrature rows of csrc/element_generic.cuh, both compiled with g++, for an entry-by-entry comparison.
#include "../../stabilized_navier_stokes_flow_fenicsx_b200/csrc/element_block.cuh"

using namespace nsgpu;

template <int GD, int VDEG>
static void both(const FormParams& f, const double* x, const double* w, double* A_blocks, double* b_blocks, double* A_rows, double* b_rows) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int ND = T::ND, NV = GD + 1;
  double recs[T::NQ * PREC], crec[CREC];
  for (int q = 0; q < T::NQ; ++q) point_record<GD, VDEG>(f, x, w, q, recs + PREC * q, crec);
  for (int i = 0; i < ND * ND; ++i) { A_blocks[i] = 0.0; A_rows[i] = 0.0; }
  for (int m = 0; m < T::NENT; ++m) {
    for (int n = 0; n < T::NENT; ++n) {
      EntityBlock<GD> B;
      entity_block<GD, VDEG, true, true>(f, recs, crec, m, n, B, n % (GD + 1));
      {   // the residual entry that rides along with the block equals the stand-alone one, with and without the block
        const int r = n % (GD + 1);
        if (r < GD || m < NV) {
          EntityBlock<GD> B2;
          entity_block<GD, VDEG, false, true>(f, recs, crec, m, 0, B2, r);
          const double ref = entity_rhs<GD, VDEG>(f, recs, crec, m, r);
          const double tol = 1e-13 * (fabs(ref) + 1e-300);
          if (fabs(B.b - ref) > tol || fabs(B2.b - ref) > tol) A_blocks[0] = NAN;
        }
      }
      for (int c = 0; c < GD; ++c) {
        for (int d = 0; d < GD; ++d) A_blocks[(GD * m + c) * ND + GD * n + d] = B.vv[c][d];
        if (n < NV) A_blocks[(GD * m + c) * ND + T::POFF + n] = B.vp[c];
        if (m < NV) A_blocks[(T::POFF + m) * ND + GD * n + c] = B.pv[c];
      }
      if (m < NV && n < NV) A_blocks[(T::POFF + m) * ND + T::POFF + n] = B.pp;
    }
    for (int r = 0; r < GD; ++r) b_blocks[GD * m + r] = entity_rhs<GD, VDEG>(f, recs, crec, m, r);
    if (m < NV) b_blocks[T::POFF + m] = entity_rhs<GD, VDEG>(f, recs, crec, m, GD);
  }
  for (int row = 0; row < ND; ++row) {
    b_rows[row] = 0.0;
    element_row<GD, VDEG, true, true>(f, x, w, row, A_rows + row * ND, b_rows + row);
  }
}

// the lean G-metric blocks (gm_row_side + gm_block: what the row-owner kernel runs for flavour 0) against the same rows
template <int GD, int VDEG>
static void lean(const FormParams& f, const double* x, const double* w, double* A_blocks, double* b_blocks) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int ND = T::ND, NV = GD + 1;
  double recs[T::NQ * PREC], crec[CREC], rs[T::NQ * RSIDE];
  for (int q = 0; q < T::NQ; ++q) point_record<GD, VDEG>(f, x, w, q, recs + PREC * q, crec);
  for (int i = 0; i < ND * ND; ++i) A_blocks[i] = 0.0;
  for (int i = 0; i < ND; ++i) b_blocks[i] = 0.0;
  for (int m = 0; m < T::NENT; ++m) {
    const bool mv = m < NV;
    for (int q = 0; q < T::NQ; ++q) {
      double b[GD + 1];
      for (int k = 0; k < RSIDE; ++k) rs[RSIDE * q + k] = NAN;   // unused slots must stay unread
      if (mv) gm_row_side<GD, VDEG, true>(f, recs + PREC * q, crec, m, q, rs + RSIDE * q, b);
      else gm_row_side<GD, VDEG, false>(f, recs + PREC * q, crec, m, q, rs + RSIDE * q, b);
      for (int r = 0; r < GD; ++r) b_blocks[GD * m + r] += b[r];
      if (mv) b_blocks[T::POFF + m] += b[GD];
    }
    for (int n = 0; n < T::NENT; ++n) {
      EntityBlock<GD> B;
      if (mv) gm_block<GD, VDEG, true>(f, recs, crec, rs, m, n, B);
      else gm_block<GD, VDEG, false>(f, recs, crec, rs, m, n, B);
      for (int c = 0; c < GD; ++c) {
        for (int d = 0; d < GD; ++d) A_blocks[(GD * m + c) * ND + GD * n + d] = B.vv[c][d];
        if (n < NV) A_blocks[(GD * m + c) * ND + T::POFF + n] = B.vp[c];
        if (mv) A_blocks[(T::POFF + m) * ND + GD * n + c] = B.pv[c];
      }
      if (mv && n < NV) A_blocks[(T::POFF + m) * ND + T::POFF + n] = B.pp;
    }
  }
}

extern "C" int lean_blocks(int gd, int vdeg, double nu, double Ci, const double* x, const double* w, double* A_blocks, double* b_blocks) {
  const FormParams f{0, nu, Ci, 1.0, 1.0, 0.0};
  if (gd == 3 && vdeg == 1) lean<3, 1>(f, x, w, A_blocks, b_blocks);
  else if (gd == 3 && vdeg == 2) lean<3, 2>(f, x, w, A_blocks, b_blocks);
  else if (gd == 2 && vdeg == 1) lean<2, 1>(f, x, w, A_blocks, b_blocks);
  else if (gd == 2 && vdeg == 2) lean<2, 2>(f, x, w, A_blocks, b_blocks);
  else return -1;
  return 0;
}

extern "C" int block_vs_rows(int gd, int vdeg, int flavour, double nu, double Ci, double alpha, double sp, double beta, const double* x,
                             const double* w, double* A_blocks, double* b_blocks, double* A_rows, double* b_rows) {
  const FormParams f{flavour, nu, Ci, alpha, sp, beta};
  if (gd == 3 && vdeg == 1) both<3, 1>(f, x, w, A_blocks, b_blocks, A_rows, b_rows);
  else if (gd == 3 && vdeg == 2) both<3, 2>(f, x, w, A_blocks, b_blocks, A_rows, b_rows);
  else if (gd == 2 && vdeg == 1) both<2, 1>(f, x, w, A_blocks, b_blocks, A_rows, b_rows);
  else if (gd == 2 && vdeg == 2) both<2, 2>(f, x, w, A_blocks, b_blocks, A_rows, b_rows);
  else return -1;
  return 0;
}
