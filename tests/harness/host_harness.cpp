// Host-compiled harness around the PRODUCT's element headers (csrc/*.cuh are __host__ __device__), so the
// kernel math can be checked against the oracle on machines without a GPU.  Test infrastructure only.
#include "element_generic.cuh"
#include <cstring>

using namespace nsgpu;

template <int GD, int VDEG>
static void run(const FormParams& f, const double* x, const double* w, double* Ae, double* be) {
  using T = ElemTraits<GD, VDEG>;
  for (int r = 0; r < T::ND; ++r) {
    double row[T::ND];
    std::memset(row, 0, sizeof(row));
    double b = 0.0;
    element_row<GD, VDEG, true, true>(f, x, w, r, row, &b);
    for (int j = 0; j < T::ND; ++j) Ae[r * T::ND + j] = row[j];
    be[r] = b;
  }
}

extern "C" int harness_element_generic(int flavour, int gdim, int vdeg, double nu, double Ci, double alpha, double sp, double beta,
                                       const double* x, const double* w, double* Ae, double* be) {
  FormParams f{flavour, nu, Ci, alpha, sp, beta};
  if (gdim == 3 && vdeg == 1) run<3, 1>(f, x, w, Ae, be);
  else if (gdim == 3 && vdeg == 2) run<3, 2>(f, x, w, Ae, be);
  else if (gdim == 2 && vdeg == 1) run<2, 1>(f, x, w, Ae, be);
  else if (gdim == 2 && vdeg == 2) run<2, 2>(f, x, w, Ae, be);
  else return -1;
  return 0;
}

#include "element_p1tet.cuh"

// Full 16x16 element Jacobian / residual rebuilt from the four vertex row-slabs of the factorised kernel
// (each slab is computed with its row vertex rotated to the front, exactly as the CUDA kernel does).
extern "C" int harness_p1tet(double nu, double Ci, const double* x, const double* w, double* Ae, double* be) {
  FormParams f{0, nu, Ci, 1.0, 1.0, 0.0};
  for (int m = 0; m < 4; ++m) {
    int perm[4] = {m, 0, 0, 0};
    for (int k = 0, j = 1; k < 4; ++k) if (k != m) perm[j++] = k;
    double xx[4][3], uu[4][3], pp[4], blk[4][16], fr[4];
    for (int a = 0; a < 4; ++a) {
      for (int i = 0; i < 3; ++i) { xx[a][i] = x[3 * perm[a] + i]; uu[a][i] = w[3 * perm[a] + i]; }
      pp[a] = w[12 + perm[a]];
    }
    p1tet_rowslab<true, true>(f, m == 0, xx, uu, pp, blk, fr);
    for (int r = 0; r < 4; ++r) {
      const int row = r < 3 ? 3 * m + r : 12 + m;
      be[row] = fr[r];
      for (int a = 0; a < 4; ++a)
        for (int d = 0; d < 4; ++d) {
          const int col = d < 3 ? 3 * perm[a] + d : 12 + perm[a];
          Ae[row * 16 + col] = blk[a][4 * r + d];
        }
    }
  }
  return 0;
}

// Same reconstruction through the register-lean variant (scratch and emit are plain arrays on the host).
extern "C" int harness_p1tet2(double nu, double Ci, const double* x, const double* w, double* Ae, double* be) {
  FormParams f{0, nu, Ci, 1.0, 1.0, 0.0};
  for (int m = 0; m < 4; ++m) {
    int perm[4] = {m, 0, 0, 0};
    for (int k = 0, j = 1; k < 4; ++k) if (k != m) perm[j++] = k;
    double xx[4][3], uu[4][3], pp[4], fr[4];
    for (int a = 0; a < 4; ++a) {
      for (int i = 0; i < 3; ++i) { xx[a][i] = x[3 * perm[a] + i]; uu[a][i] = w[3 * perm[a] + i]; }
      pp[a] = w[12 + perm[a]];
    }
    struct Scratch { P1TetPoint q[4]; double after_geometry(double w) const { return w; } void put(int i, const P1TetPoint& p) { q[i] = p; } void get(int i, P1TetPoint& p) const { p = q[i]; } } sc;
    double blks[4][16];
    auto emit = [&](int n, const double (&b)[16]) { for (int k = 0; k < 16; ++k) blks[n][k] = b[k]; };
    p1tet_rowslab2<true, true>(f, m == 0, xx, uu, pp, fr, sc, emit);
    for (int r = 0; r < 4; ++r) {
      const int row = r < 3 ? 3 * m + r : 12 + m;
      be[row] = fr[r];
      for (int a = 0; a < 4; ++a)
        for (int d = 0; d < 4; ++d) {
          const int col = d < 3 ? 3 * perm[a] + d : 12 + perm[a];
          Ae[row * 16 + col] = blks[a][4 * r + d];
        }
    }
  }
  return 0;
}

#include "element_shared.cuh"

template <int GD, int VDEG>
static void run_shared(const FormParams& f, const double* x, const double* w, double* Ae, double* be) {
  using T = ElemTraits<GD, VDEG>;
  PointData<GD, VDEG> P[T::NQ];
  CellData<GD> C;
  for (int q = 0; q < T::NQ; ++q) point_setup<GD, VDEG>(f, x, w, q, P[q], C);
  for (int r = 0; r < T::ND; ++r) {
    double row[T::ND];
    std::memset(row, 0, sizeof(row));
    double b = 0.0;
    for (int q = 0; q < T::NQ; ++q) row_from_point<GD, VDEG, true, true>(f, P[q], C, r, row, &b);
    for (int j = 0; j < T::ND; ++j) Ae[r * T::ND + j] = row[j];
    be[r] = b;
  }
}

// two-stage (point data shared by the rows of a cell) variant of the generic element tensors
extern "C" int harness_element_shared(int flavour, int gdim, int vdeg, double nu, double Ci, double alpha, double sp, double beta,
                                      const double* x, const double* w, double* Ae, double* be) {
  FormParams f{flavour, nu, Ci, alpha, sp, beta};
  if (gdim == 3 && vdeg == 1) run_shared<3, 1>(f, x, w, Ae, be);
  else if (gdim == 3 && vdeg == 2) run_shared<3, 2>(f, x, w, Ae, be);
  else if (gdim == 2 && vdeg == 1) run_shared<2, 1>(f, x, w, Ae, be);
  else if (gdim == 2 && vdeg == 2) run_shared<2, 2>(f, x, w, Ae, be);
  else return -1;
  return 0;
}
