// assemble.cu -- global residual / Jacobian assembly with dolfinx semantics
// (assemble_vector + apply_lifting + set_bc and assemble_matrix of
//  NavierStokes/NavierStokesChannelFlow.py:62-67,73-74).
//
// Generic path: one thread per (owned cell, test dof).  The thread evaluates its row of the element
// Jacobian and its residual entry (element_generic.cuh), applies the Dirichlet row/column zeroing and the
// lifting term, and adds into the CSR values through the precomputed entity-relative position map.
#include "common.cuh"
#include "element_p1tet.cuh"
#include "element_shared.cuh"

namespace nsgpu {

template <int GD, int VDEG, bool WANT_J, bool WANT_F>
__global__ void __launch_bounds__(128)
k_assemble_generic(int64_t n_cells, FormParams form, const double* __restrict__ xg, const int32_t* __restrict__ cells,
                   const int32_t* __restrict__ dofmap, const double* __restrict__ wv, const uint8_t* __restrict__ bc_marker,
                   const double* __restrict__ bc_value, const int64_t* __restrict__ indptr, const uint16_t* __restrict__ rel,
                   double* __restrict__ vals, double* __restrict__ F) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int ND = T::ND;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_cells * ND) return;
  const int64_t cell = t / ND;
  const int row = (int)(t - cell * ND);

  double x[3 * (GD + 1)];
#pragma unroll
  for (int a = 0; a <= GD; ++a) {
    const int64_t v = cells[cell * (GD + 1) + a];
#pragma unroll
    for (int i = 0; i < 3; ++i) x[3 * a + i] = xg[3 * v + i];
  }
  const int32_t* dm = dofmap + cell * ND;
  double w[ND];
  bool cell_bc = false;
  for (int k = 0; k < ND; ++k) {
    const int32_t d = dm[k];
    w[k] = wv[d];
    if (bc_marker) cell_bc |= bc_marker[d] != 0;
  }
  const int32_t gi = dm[row];
  const bool need_A = WANT_J || (WANT_F && cell_bc);  // lifting needs the un-zeroed row

  double Arow[ND];
  for (int k = 0; k < ND; ++k) Arow[k] = 0.0;
  double b = 0.0;
  if (need_A) element_row<GD, VDEG, true, WANT_F>(form, x, w, row, Arow, &b);
  else element_row<GD, VDEG, false, WANT_F>(form, x, w, row, Arow, &b);

  if (WANT_F) {
    if (cell_bc) {
      // apply_lifting(F, [a], [bc], [x], -1.0):  b_e[i] += Ae[i][j] (g_j - x_j) over constrained trial dofs j
      for (int j = 0; j < ND; ++j) {
        const int32_t dj = dm[j];
        if (bc_marker[dj]) b += Arow[j] * (bc_value[dj] - w[j]);
      }
    }
    atomicAdd(F + gi, b);
  }
  if (WANT_J) {
    const bool row_bc = bc_marker && bc_marker[gi];
    if (!row_bc) {  // constrained test rows are zeroed: nothing to add
      const int64_t base = indptr[gi];
      const uint16_t* r = rel + (cell * T::NENT + entity_of_local_dof<GD, VDEG>(row)) * ND;
      for (int j = 0; j < ND; ++j) {
        const bool col_bc = cell_bc && bc_marker[dm[j]];
        if (!col_bc) atomicAdd(vals + base + r[j], Arow[j]);
      }
    }
  }
}

// Cooperative variant of the generic path: a CTA takes a batch of cells; the quadrature-point data of each cell
// (basis tables, fields, stabilisation parameters and their derivatives) is computed once by one thread per (cell, point),
// parked in shared memory, and consumed by the cell's ND row threads.  On P2-P1 tetrahedra this removes the 34-fold
// recomputation of the generic kernel.
template <int GD, int VDEG> struct CoopCfg {
  static constexpr int ND = ElemTraits<GD, VDEG>::ND;
  static constexpr int CPB = 256 / ND;                 // cells per block
  static constexpr int NT = 256;
};

template <int GD, int VDEG, bool WANT_J, bool WANT_F>
__global__ void __launch_bounds__(256)
k_assemble_coop(int64_t n_cells, FormParams form, const double* __restrict__ xg, const int32_t* __restrict__ cells,
                const int32_t* __restrict__ dofmap, const double* __restrict__ wv, const uint8_t* __restrict__ bc_marker,
                const double* __restrict__ bc_value, const int64_t* __restrict__ indptr, const uint16_t* __restrict__ rel,
                double* __restrict__ vals, double* __restrict__ F) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int ND = T::ND, NQ = T::NQ, CPB = CoopCfg<GD, VDEG>::CPB, NT = 256;
  __shared__ PointData<GD, VDEG> sP[CPB][NQ];
  __shared__ CellData<GD> sC[CPB];
  __shared__ double sx[CPB][3 * (GD + 1)];
  __shared__ double sw[CPB][ND];
  __shared__ int32_t sdof[CPB][ND];
  __shared__ uint8_t smk[CPB][ND];
  __shared__ int sbc[CPB];
  const int tid = threadIdx.x;
  const int64_t cell0 = (int64_t)blockIdx.x * CPB;
  const int ncell = (int)((n_cells - cell0) < CPB ? (n_cells - cell0) : CPB);

  // stage 0: the batch's dofs, coefficients, Dirichlet flags, vertex coordinates
  if (tid < CPB) sbc[tid] = 0;
  __syncthreads();
  for (int t = tid; t < ncell * ND; t += NT) {
    const int lc = t / ND, k = t - lc * ND;
    const int32_t d = dofmap[(cell0 + lc) * ND + k];
    sdof[lc][k] = d;
    sw[lc][k] = wv[d];
    const uint8_t mk = bc_marker ? bc_marker[d] : 0;
    smk[lc][k] = mk;
    if (mk) sbc[lc] = 1;
  }
  for (int t = tid; t < ncell * (GD + 1) * 3; t += NT) {
    const int lc = t / ((GD + 1) * 3), r = t - lc * (GD + 1) * 3;
    const int a = r / 3, i = r - 3 * a;
    sx[lc][r] = xg[3 * (int64_t)cells[(cell0 + lc) * (GD + 1) + a] + i];
  }
  __syncthreads();
  // stage 1: one thread per (cell, quadrature point)
  for (int t = tid; t < ncell * NQ; t += NT) {
    const int lc = t / NQ, q = t - lc * NQ;
    CellData<GD> C;
    point_setup<GD, VDEG>(form, sx[lc], sw[lc], q, sP[lc][q], C);
    if (q == 0) sC[lc] = C;
  }
  __syncthreads();
  // stage 2: one thread per (cell, test dof)
  const int lc = tid / ND, row = tid - lc * ND;
  if (lc >= ncell) return;
  const bool cell_bc = sbc[lc] != 0;
  const bool need_A = WANT_J || (WANT_F && cell_bc);
  double Arow[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) Arow[k] = 0.0;
  double b = 0.0;
  if (need_A) {
    for (int q = 0; q < NQ; ++q) row_from_point<GD, VDEG, true, WANT_F>(form, sP[lc][q], sC[lc], row, Arow, &b);
  } else {
    for (int q = 0; q < NQ; ++q) row_from_point<GD, VDEG, false, WANT_F>(form, sP[lc][q], sC[lc], row, Arow, &b);
  }
  const int32_t gi = sdof[lc][row];
  if (WANT_F) {
    if (cell_bc) {
      // apply_lifting(F, [a], [bc], [x], -1.0):  b_e[i] += Ae[i][j] (g_j - x_j) over constrained trial dofs j
#pragma unroll
      for (int j = 0; j < ND; ++j)
        if (smk[lc][j]) b += Arow[j] * (bc_value[sdof[lc][j]] - sw[lc][j]);
    }
    atomicAdd(F + gi, b);
  }
  if (WANT_J && !smk[lc][row]) {   // constrained test rows are zeroed: nothing to add
    const int64_t base = indptr[gi];
    const uint16_t* r = rel + ((cell0 + lc) * T::NENT + entity_of_local_dof<GD, VDEG>(row)) * ND;
#pragma unroll
    for (int j = 0; j < ND; ++j)
      if (!smk[lc][j]) atomicAdd(vals + base + r[j], Arow[j]);
  }
}

template <int GD, int VDEG>
static void launch_coop(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_F) {
  constexpr int CPB = CoopCfg<GD, VDEG>::CPB;
  const unsigned grid = (unsigned)ceil_div(ctx->n_cells_owned > 0 ? ctx->n_cells_owned : 1, CPB);
  const uint8_t* mk = ctx->has_bc ? ctx->d_bc_marker : nullptr;
#define NS_LAUNCH(J, F)                                                                                        \
  k_assemble_coop<GD, VDEG, J, F><<<grid, 256, 0, ctx->stream>>>(ctx->n_cells_owned, ctx->form, ctx->d_x, ctx->d_cells, \
      ctx->d_dofmap, d_xin, mk, ctx->d_bc_value, ctx->d_indptr, ctx->d_rel, ctx->d_vals, d_F)
  if (want_J && want_F) NS_LAUNCH(true, true);
  else if (want_J) NS_LAUNCH(true, false);
  else NS_LAUNCH(false, true);
#undef NS_LAUNCH
  ctx->launches += 1;
}

// assemble_matrix's diagonal pass: +1.0 per DirichletBC object holding the owned dof
__global__ void k_bc_diagonal(int64_t n_owned, const int32_t* __restrict__ mult, const int64_t* __restrict__ diag, double* vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_owned && mult[i] > 0 && diag[i] >= 0) vals[diag[i]] += (double)mult[i];
}

// set_bc(F, bc, x, -1.0):  F[dof] = -(g - x[dof]) on owned constrained dofs
__global__ void k_set_bc(int64_t n_owned, const uint8_t* __restrict__ marker, const double* __restrict__ value,
                         const double* __restrict__ xv, double* F) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_owned && marker[i]) F[i] = xv[i] - value[i];
}

template <int GD, int VDEG>
static void launch_generic(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_F) {
  using T = ElemTraits<GD, VDEG>;
  const int bs = 128;
  const int64_t nthreads = ctx->n_cells_owned * T::ND;
  const unsigned grid = (unsigned)ceil_div(nthreads > 0 ? nthreads : 1, bs);
  const uint8_t* mk = ctx->has_bc ? ctx->d_bc_marker : nullptr;
#define NS_LAUNCH(J, F)                                                                                           \
  k_assemble_generic<GD, VDEG, J, F><<<grid, bs, 0, ctx->stream>>>(ctx->n_cells_owned, ctx->form, ctx->d_x, ctx->d_cells, \
      ctx->d_dofmap, d_xin, mk, ctx->d_bc_value, ctx->d_indptr, ctx->d_rel, ctx->d_vals, d_F)
  if (want_J && want_F) NS_LAUNCH(true, true);
  else if (want_J) NS_LAUNCH(true, false);
  else NS_LAUNCH(false, true);
#undef NS_LAUNCH
  ctx->launches += 1;
}

__global__ void k_zero_positions(int64_t n, const int64_t* __restrict__ pos, double* vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) vals[pos[i]] = 0.0;
}

// assemble_matrix's BC diagonal pass on the owned rows (also used by the streamed host path)
int k_bc_diagonal_launch(nsgpu_ctx* ctx) {
  k_bc_diagonal<<<(unsigned)ceil_div(ctx->n_owned, 256), 256, 0, ctx->stream>>>(ctx->n_owned, ctx->d_bc_mult, ctx->d_diag, ctx->d_vals);
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

// NaN / Inf guard on a freshly assembled vector (PETSc's SNES stops with DIVERGED_FNORM_NAN; here the
// assembly call itself reports it).  All ranks see the same verdict: the flag is summed over the communicator.
__global__ void k_nonfinite(int64_t n, const double* __restrict__ v, int* flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && !isfinite(v[i])) *flag = 1;
}

// sync_now = false (device-pointer entry points, which stay asynchronous): the scan is queued and the sticky flag is
// read by the next nsgpu_sync / host-vector call
int check_finite_impl(nsgpu_ctx* ctx, const double* d_v, int64_t n, const char* what, bool sync_now) {
  cudaStream_t s = ctx->stream;
  if (!ctx->d_nonfinite) {
    NS_CUDA(ctx, cudaMalloc(&ctx->d_nonfinite, 2 * sizeof(double)));
    NS_CUDA(ctx, cudaMemsetAsync(ctx->d_nonfinite, 0, 2 * sizeof(double), s));
  }
  if (d_v && n > 0) {
    k_nonfinite<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(n, d_v, ctx->d_nonfinite);
    ctx->launches += 1;
  }
  if (!sync_now) return NSGPU_OK;
  int flag = 0;
  NS_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, s));
  NS_CUDA(ctx, cudaStreamSynchronize(s));
  if (flag) {
    NS_CUDA(ctx, cudaMemsetAsync(ctx->d_nonfinite, 0, 2 * sizeof(double), s));
    set_error(ctx, std::string(what) + ": non-finite entries (NaN / Inf) in the assembled vector");
    return NSGPU_ENONFINITE;
  }
  return NSGPU_OK;
}

// d_xin: n_cols state (halo already refreshed).  d_Fout: n_cols residual (zeroed here; owned part meaningful).
int assemble_impl(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_Fout) {
  cudaStream_t s = ctx->stream;
  if (want_J) ctx->jac_valid = false;   // the resident values are about to change; fuse_fj callers re-validate afterwards
  bool fast = false;
  if (ctx->gdim == 3 && ctx->vdeg == 1 && ctx->form.flavour == NSGPU_FORM_GMETRIC && ctx->kernel_sel != NSGPU_KERNEL_GENERIC)
    fast = p1tet_fast_available(ctx);
  if (ctx->kernel_sel == NSGPU_KERNEL_FAST && !fast) {
    set_error(ctx, "kernel=fast requested but the factorised kernel does not apply to this element/form/numbering");
    return NSGPU_EUNSUPPORTED;
  }
  // J.zeroEntries(): the row-owner kernel writes every entry of every locally assembled row exactly once, so the
  // zero-fill pass is only needed when rows also hold entries that only other ranks contribute to.
  // With several ranks the entries that receive contributions of other ranks' ghost rows (the J.assemble() exchange adds
  // into them, and some of them are entries no local cell touches) are the only ones that have to start from zero: they are
  // exactly the receive positions of the ghost-row plan, a few 10 MB instead of the whole value array.
  // The row-owner kernel of the other element pairs / forms (rowown.cu) writes whole rows, including the entries only other
  // ranks contribute to, so it never needs the zero-fill.
  const bool rowown = !fast && ctx->kernel_sel == NSGPU_KERNEL_AUTO && rowown_available(ctx);
  const bool zero_vals = want_J && !rowown && !(fast && ctx->extra_rows.empty());
  const bool zero_recv_only = zero_vals && fast && ctx->rows.n_neigh > 0 && !ctx->rows.recv_ptr.empty() && ctx->rows.recv_ptr.back() > 0;
  if (zero_recv_only) {
    const int64_t nr = ctx->rows.recv_ptr.back();
    k_zero_positions<<<(unsigned)ceil_div(nr, 256), 256, 0, s>>>(nr, ctx->rows.d_recv_pos, ctx->d_vals);
    ctx->launches += 1;
  } else if (zero_vals) {
    NS_CUDA(ctx, cudaMemsetAsync(ctx->d_vals, 0, sizeof(double) * (ctx->nnz > 0 ? ctx->nnz : 1), s));
  }
  if (want_F) NS_CUDA(ctx, cudaMemsetAsync(d_Fout, 0, sizeof(double) * ctx->n_cols, s));                          // f_local.set(0.0)
  NS_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));

  // Several ranks: the tiles that hold ghost vertices are assembled first; packing and the NCCL send / receive of the ghost
  // rows (J.assemble()) and of the ghost residual entries (F.ghostUpdate(ADD, REVERSE)) then run on a second stream while
  // the interior tiles are assembled, and only the adds wait for both (NavierStokesChannelFlow.py:66, :75).
  const bool split = fast && ctx->nranks > 1 && ctx->overlap && ctx->halo.n_neigh > 0 && p1tet_can_split(ctx);
  if (split) {
    if (!ctx->stream2) {
      int lo = 0, hi = 0;
      NS_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      NS_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
      NS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_x[0], cudaEventDisableTiming));
      NS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_x[1], cudaEventDisableTiming));
    }
    int rc = p1tet_assemble(ctx, d_xin, want_J, want_F, d_Fout, 1, 0);
    if (rc != NSGPU_OK) return rc;
    NS_CUDA(ctx, cudaEventRecord(ctx->ev_x[0], s));
    NS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_x[0], 0));
    if (want_J && (rc = rows_exchange_begin(ctx, ctx->stream2))) return rc;
    if (want_F && (rc = halo_reverse_begin(ctx, d_Fout, ctx->stream2))) return rc;
    NS_CUDA(ctx, cudaEventRecord(ctx->ev_x[1], ctx->stream2));
    if ((rc = p1tet_assemble(ctx, d_xin, want_J, want_F, d_Fout, 2, ctx->sm_reserve))) return rc;
    NS_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
    NS_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_x[1], 0));
    if (want_J) {
      if ((rc = rows_exchange_end(ctx))) return rc;
      if (ctx->has_bc) {
        k_bc_diagonal<<<(unsigned)ceil_div(ctx->n_owned, 256), 256, 0, s>>>(ctx->n_owned, ctx->d_bc_mult, ctx->d_diag, ctx->d_vals);
        ctx->launches += 1;
      }
    }
    if (want_F) {
      if ((rc = halo_reverse_end(ctx, d_Fout))) return rc;
      if (ctx->has_bc) {
        k_set_bc<<<(unsigned)ceil_div(ctx->n_owned, 256), 256, 0, s>>>(ctx->n_owned, ctx->d_bc_marker, ctx->d_bc_value, d_xin, d_Fout);
        ctx->launches += 1;
      }
    }
    NS_CUDA(ctx, cudaGetLastError());
    return NSGPU_OK;
  }
  if (fast) {
    int rc = p1tet_assemble(ctx, d_xin, want_J, want_F, d_Fout);
    if (rc != NSGPU_OK) return rc;
  } else if (rowown) {
    ctx->last_kernel = "rowown";
    const int rc = rowown_assemble(ctx, d_xin, want_J, want_F, d_Fout);
    if (rc != NSGPU_OK) return rc;
  } else {
    const int key = ctx->gdim * 10 + ctx->vdeg;
    const bool coop = ctx->kernel_sel != NSGPU_KERNEL_GENERIC;   // AUTO: cooperative (shared point data); GENERIC: thread per row
    ctx->last_kernel = coop ? "generic_coop" : "generic_row";
    switch (key) {
      case 31: coop ? launch_coop<3, 1>(ctx, d_xin, want_J, want_F, d_Fout) : launch_generic<3, 1>(ctx, d_xin, want_J, want_F, d_Fout); break;
      case 32: coop ? launch_coop<3, 2>(ctx, d_xin, want_J, want_F, d_Fout) : launch_generic<3, 2>(ctx, d_xin, want_J, want_F, d_Fout); break;
      case 21: coop ? launch_coop<2, 1>(ctx, d_xin, want_J, want_F, d_Fout) : launch_generic<2, 1>(ctx, d_xin, want_J, want_F, d_Fout); break;
      case 22: coop ? launch_coop<2, 2>(ctx, d_xin, want_J, want_F, d_Fout) : launch_generic<2, 2>(ctx, d_xin, want_J, want_F, d_Fout); break;
      default: set_error(ctx, "unsupported element"); return NSGPU_EUNSUPPORTED;
    }
  }
  NS_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
  NS_CUDA(ctx, cudaGetLastError());

  if (want_J) {
    int rc = rows_exchange_add(ctx);   // J.assemble(): ghost rows -> owners
    if (rc != NSGPU_OK) return rc;
    if (ctx->has_bc) {
      k_bc_diagonal<<<(unsigned)ceil_div(ctx->n_owned, 256), 256, 0, s>>>(ctx->n_owned, ctx->d_bc_mult, ctx->d_diag, ctx->d_vals);
      ctx->launches += 1;
    }
  }
  if (want_F) {
    int rc = halo_reverse_add(ctx, d_Fout);   // F.ghostUpdate(ADD, REVERSE)
    if (rc != NSGPU_OK) return rc;
    if (ctx->has_bc) {
      k_set_bc<<<(unsigned)ceil_div(ctx->n_owned, 256), 256, 0, s>>>(ctx->n_owned, ctx->d_bc_marker, ctx->d_bc_value, d_xin, d_Fout);
      ctx->launches += 1;
    }
  }
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

}  // namespace nsgpu
