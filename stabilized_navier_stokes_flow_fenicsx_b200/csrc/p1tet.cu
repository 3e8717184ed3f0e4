// p1tet.cu -- factorised P1-P1 tetrahedron G-metric assembly (placeholder until the fast kernels land).
#include "element_p1tet.cuh"

namespace nsgpu {
bool p1tet_fast_available(nsgpu_ctx*) { return false; }
int p1tet_assemble(nsgpu_ctx* ctx, const double*, bool, bool, double*) {
  set_error(ctx, "fast P1-P1 tet kernel not built");
  return NSGPU_EUNSUPPORTED;
}
}  // namespace nsgpu
