// element_p1tet.cuh -- factorised P1-P1 tetrahedron G-metric kernels (declarations).
#pragma once
#include "common.cuh"

namespace nsgpu {
bool p1tet_fast_available(nsgpu_ctx* ctx);
int p1tet_assemble(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_Fout);
}  // namespace nsgpu
