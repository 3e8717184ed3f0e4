// Test harness only: compiles csrc/trace_core.cuh (the streamline integrator the GPU kernels run, all of it
// __host__ __device__) with g++ so that the CPU test suite can compare it with scipy.solve_ivp step for step.
// Not part of the product: libnsgpu.so has no CPU path.
#include "../../stabilized_navier_stokes_flow_fenicsx_b200/csrc/trace_core.cuh"

using namespace nsgpu;

extern "C" {

static TraceField make_field(const double* cmap, const double* cvel, const int64_t* bin_ptr, const int32_t* bin_cells,
                             const double* lo, const double* inv_h, const int32_t* nb, double tol) {
  TraceField F;
  F.cmap = cmap; F.cvel = cvel; F.bin_ptr = bin_ptr; F.bin_cells = bin_cells;
  for (int k = 0; k < 3; ++k) { F.lo[k] = lo[k]; F.inv_h[k] = inv_h[k]; F.nb[k] = nb[k]; }
  F.tol = tol;
  return F;
}

void host_trace_velocity(const double* cmap, const double* cvel, const int64_t* bin_ptr, const int32_t* bin_cells, const double* lo,
                         const double* inv_h, const int32_t* nb, double tol, int64_t n, const double* pts, double* vel, int32_t* cell) {
  const TraceField F = make_field(cmap, cvel, bin_ptr, bin_cells, lo, inv_h, nb, tol);
  for (int64_t i = 0; i < n; ++i) {
    int32_t c = -1;
    tr_velocity(F, 1.0, pts + 3 * i, c, vel + 3 * i);
    cell[i] = c;
  }
}

void host_trace_run(const double* cmap, const double* cvel, const int64_t* bin_ptr, const int32_t* bin_cells, const double* lo,
                    const double* inv_h, const int32_t* nb, double tol, int64_t n, const double* seeds, int reverse, double x_stop,
                    double speed_min, double t_end, double max_step, double rtol, double atol, int64_t max_steps, double* end_xyz,
                    int32_t* status, double* t_final, int32_t* n_steps, int32_t* n_fev) {
  const TraceField F = make_field(cmap, cvel, bin_ptr, bin_cells, lo, inv_h, nb, tol);
  TraceParams P;
  P.dir = reverse ? -1.0 : 1.0; P.x_stop = x_stop; P.x_dir = reverse ? -1.0 : 1.0; P.speed_min = speed_min;
  P.t_end = t_end; P.max_step = max_step; P.rtol = rtol; P.atol = atol; P.max_steps = max_steps;
  for (int64_t i = 0; i < n; ++i) {
    TraceResult R;
    tr_trace(F, P, seeds + 3 * i, R);
    for (int k = 0; k < 3; ++k) end_xyz[3 * i + k] = R.y[k];
    status[i] = R.status; t_final[i] = R.t; n_steps[i] = R.n_steps; n_fev[i] = R.n_fev;
  }
}

}
