// spmv.cu -- CSR sparse matrix-vector product y = J x (PETSc MatMult inside KSPTFQMR,
// NavierStokes/NavierStokesChannelFlow.py:282-283).  Rows of the mixed P1-P1 matrix hold ~60 entries,
// P2-P1 ~100-400: a sub-warp of LPR lanes walks one row with coalesced value/column loads and gathers x
// through the read-only path; partial sums are combined with shuffles.
#include "common.cuh"
#include "element_p1tet.cuh"

namespace nsgpu {

template <int LPR>
__global__ void __launch_bounds__(256)
k_spmv(int64_t n_rows, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
       const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = t / LPR;
  const int lane = (int)(t % LPR);
  double s = 0.0;
  if (row < n_rows) {
    const int64_t b = indptr[row], e = indptr[row + 1];
    for (int64_t k = b + lane; k < e; k += LPR) s += vals[k] * __ldg(x + indices[k]);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
  if (row < n_rows && lane == 0) y[row] = s;
}

int spmv_impl(nsgpu_ctx* ctx, const double* d_x, double* d_y) {
  const int64_t n = ctx->n_owned;   // owned rows only: ghost rows were shipped to their owners
  const double avg = n > 0 ? (double)ctx->nnz / (double)ctx->n_rows : 0.0;
  const int bs = 256;
  cudaStream_t s = ctx->stream;
  NS_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
  if (ctx->kernel_sel != NSGPU_KERNEL_GENERIC && p1tet_spmv(ctx, d_x, d_y)) {
    // vertex-blocked P1-P1 matrix: one column index per 4x4 block (p1tet.cu)
    ctx->last_spmv = "spmv_block4";
  } else {
  ctx->last_spmv = "spmv_csr";
  if (avg > 48) k_spmv<16><<<(unsigned)ceil_div(n * 16, bs), bs, 0, s>>>(n, ctx->d_indptr, ctx->d_indices, ctx->d_vals, d_x, d_y);
  else if (avg > 24) k_spmv<8><<<(unsigned)ceil_div(n * 8, bs), bs, 0, s>>>(n, ctx->d_indptr, ctx->d_indices, ctx->d_vals, d_x, d_y);
  else k_spmv<4><<<(unsigned)ceil_div(n * 4, bs), bs, 0, s>>>(n, ctx->d_indptr, ctx->d_indices, ctx->d_vals, d_x, d_y);
  }
  NS_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

}  // namespace nsgpu
