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
