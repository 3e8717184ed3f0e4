// p1tet.cu -- atomics-free assembly of the P1-P1 tetrahedron G-metric Jacobian + residual.
//
// Why not "one thread per cell + atomicAdd": every cell adds 256 fp64 values into the CSR matrix, each
// matrix entry receives ~6.4 of them.  Measured on B200 (tools/microbench.cu): fp64 RED peaks at ~235 G/s
// even when perfectly sector-coalesced, i.e. >= 55 ms for the 50 M-cell duct, 5x the fp64-pipe time of the
// element algebra itself.  The reduction therefore has to happen on the SM.
//
// Design ("row-owner gather"): the unit of work is an INCIDENCE (vertex A, cell c containing A).
//   * A CTA owns a TILE of consecutive vertices (row groups); all incidences of those vertices, <= CAP.
//   * Phase A -- one thread per incidence evaluates, with the factorised algebra of element_p1tet.cuh, the
//     4 x 16 row slab of the element Jacobian that belongs to its vertex (four 4x4 blocks) and the four
//     residual entries, applies Dirichlet lifting / row / column zeroing, and parks the result in shared
//     memory (512 + 32 B per incidence).
//   * Phase B -- one thread per (vertex, row, neighbour slot) sums the <= deg(A) parked blocks that target
//     that slot and writes the finished 32-byte piece of the CSR row with plain stores.  Rows of a vertex
//     are written by neighbouring threads: full-sector, coalesced, write-once.
// No atomics, no colouring, bitwise reproducible; every matrix entry is written exactly once per assembly
// (no zero-fill pass needed on a single rank).  Cost: the cell-level part of the algebra is recomputed by
// the four incidences of a cell (~3.7 k instead of ~2.5 k DFMA per cell).
//
// The plan (incidence lists, tiles, slot maps, output positions) is built once per pattern, on the device,
// from the entity-level pair list the pattern builder already sorted.
#include <cub/cub.cuh>

#include "common.cuh"
#include "element_p1tet.cuh"

struct nsgpu_p1tet_plan {
  int64_t n_inc = 0, n_ent = 0, n_tiles = 0;
  int cap = 0, maxdeg = 0;
  uint32_t* d_inc_cell = nullptr;   // [n_inc] cell * 4 + local vertex, sorted by row vertex
  uint32_t* d_inc_slot = nullptr;   // [n_inc] 4 x uint8: neighbour slot of the cell's 4 vertices (rotated order)
  int64_t* d_inc_ptr = nullptr;     // [n_ent + 1]
  int64_t* d_slot_ptr = nullptr;    // [n_ent + 1] prefix of neighbour counts
  int64_t* d_rowpos = nullptr;      // [n_ent * 4] CSR start of the vertex's 4 rows
  int32_t* d_rowdof = nullptr;      // [n_ent * 4] the vertex's 4 dofs
  int64_t* d_tile_ent = nullptr;    // [n_tiles + 1]
  uint8_t* d_cell_bc = nullptr;     // [n_cells] cell touches a Dirichlet dof
  bool bc_dirty = true;
};

namespace nsgpu {

constexpr int P1_CAP = 256;   // incidences (= phase-A threads) per CTA

// ------------------------------------------------------------------------------------------ plan kernels
__global__ void k_inc_keys(int64_t n_cells, const int32_t* __restrict__ dofmap, uint64_t* keys) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_cells * 4) return;
  const int64_t cell = t >> 2;
  const int m = (int)(t & 3);
  keys[t] = ((uint64_t)(uint32_t)dofmap[cell * 16 + 3 * m] << 32) | (uint32_t)t;
}

struct HiWord {
  __host__ __device__ uint32_t operator()(const uint64_t& k) const { return (uint32_t)(k >> 32); }
};

__global__ void k_inc_fill(int64_t n_inc, const uint64_t* __restrict__ keys, const int32_t* __restrict__ dofmap,
                           const uint64_t* __restrict__ pairs, const int64_t* __restrict__ pfirst,
                           const int64_t* __restrict__ plast, uint32_t* inc_cell, uint32_t* inc_slot, int* bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_inc) return;
  const uint32_t A = (uint32_t)(keys[i] >> 32), cm = (uint32_t)(keys[i] & 0xffffffffu);
  const int64_t cell = cm >> 2;
  const int m = cm & 3;
  inc_cell[i] = cm;
  const int64_t lo0 = pfirst[A], hi0 = plast[A];
  uint32_t packed = 0;
  int pos = 1;
  for (int k = 0; k < 4; ++k) {
    const int a = (k == m) ? 0 : pos++;                     // rotated position of original local vertex k
    const uint32_t B = (uint32_t)dofmap[cell * 16 + 3 * k];
    int64_t lo = lo0, hi = hi0 - 1, found = -1;
    while (lo <= hi) {
      const int64_t mid = (lo + hi) >> 1;
      const uint32_t v = (uint32_t)(pairs[mid] & 0xffffffffu);
      if (v == B) { found = mid - lo0; break; }
      if (v < B) lo = mid + 1; else hi = mid - 1;
    }
    if (found < 0 || found > 255) { *bad = 1; found = 0; }
    packed |= (uint32_t)found << (8 * a);
  }
  inc_slot[i] = packed;
}

__global__ void k_ent_info(int64_t n_ent, const uint32_t* __restrict__ ent_leader, const int32_t* __restrict__ members,
                           const int64_t* __restrict__ pfirst, const int64_t* __restrict__ plast,
                           const int64_t* __restrict__ indptr, int64_t* nslots, int64_t* rowpos, int32_t* rowdof) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > n_ent) return;
  if (e == n_ent) { nslots[e] = 0; return; }
  const uint32_t A = ent_leader[e];
  nslots[e] = plast[A] - pfirst[A];
  for (int c = 0; c < 4; ++c) {
    const int32_t d = members[(int64_t)A * KMAX + c];
    rowdof[e * 4 + c] = d;
    rowpos[e * 4 + c] = indptr[d];
  }
}

__global__ void k_tiles(int64_t n_tiles, int64_t n_ent, int64_t capeff, const int64_t* __restrict__ inc_ptr, int64_t* tile_ent) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  const int64_t target = t * capeff;   // first entity whose incidence offset is >= target
  int64_t lo = 0, hi = n_ent;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (inc_ptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  tile_ent[t] = lo;
}

__global__ void k_cell_bc(int64_t n_cells, const int32_t* __restrict__ dofmap, const uint8_t* __restrict__ marker, uint8_t* cell_bc) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  uint8_t f = 0;
  for (int k = 0; k < 16; ++k) f |= marker[dofmap[c * 16 + k]];
  cell_bc[c] = f;
}

// ------------------------------------------------------------------------------------------ the kernel
template <bool WANT_J, bool WANT_F>
__global__ void __launch_bounds__(P1_CAP, 1)
k_p1tet_tiles(FormParams form, const double* __restrict__ xg, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap,
              const double* __restrict__ wv, const uint8_t* __restrict__ bc_marker, const double* __restrict__ bc_value,
              const uint8_t* __restrict__ cell_bc, const uint32_t* __restrict__ inc_cell, const uint32_t* __restrict__ inc_slot,
              const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ slot_ptr, const int64_t* __restrict__ rowpos,
              const int32_t* __restrict__ rowdof, const int64_t* __restrict__ tile_ent, double* __restrict__ vals, double* __restrict__ F) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: stageJ [16][CAP] double4 | stageF [CAP] double4 | slots [CAP] u32 | incp [CAP+1] i32 | slotp [CAP+1] i32
  double4* stageJ = reinterpret_cast<double4*>(smem_raw);
  double4* stageF = stageJ + (WANT_J ? 16 * P1_CAP : 0);
  uint32_t* slots = reinterpret_cast<uint32_t*>(stageF + P1_CAP);
  int* incp = reinterpret_cast<int*>(slots + P1_CAP);
  int* slotp = incp + (P1_CAP + 1);

  const int tid = threadIdx.x;
  const int64_t e0 = tile_ent[blockIdx.x], e1 = tile_ent[blockIdx.x + 1];
  const int nent = (int)(e1 - e0);
  if (nent <= 0) return;
  const int64_t i0 = inc_ptr[e0];
  const int ninc = (int)(inc_ptr[e1] - i0);
  const int64_t s0 = slot_ptr[e0];
  for (int k = tid; k <= nent; k += P1_CAP) {
    incp[k] = (int)(inc_ptr[e0 + k] - i0);
    slotp[k] = (int)(slot_ptr[e0 + k] - s0);
  }

  // ---------------- phase A: one incidence per thread ----------------
  if (tid < ninc) {
    const uint32_t cm = inc_cell[i0 + tid];
    const int64_t cell = cm >> 2;
    const int m = cm & 3;
    slots[tid] = inc_slot[i0 + tid];
    // rotated local vertex order: row vertex first, the others ascending
    int perm[4];
    perm[0] = m;
    perm[1] = (m == 0) ? 1 : 0;
    perm[2] = (m <= 1) ? 2 : 1;
    perm[3] = (m <= 2) ? 3 : 2;
    double x[4][3], u[4][3], p[4];
    int32_t dof[4][4];
    const int4 cv = *reinterpret_cast<const int4*>(cells + cell * 4);
    const int vtx[4] = {cv.x, cv.y, cv.z, cv.w};
    const int4* dmr = reinterpret_cast<const int4*>(dofmap + cell * 16);
    const int4 d0 = dmr[0], d1 = dmr[1], d2 = dmr[2], d3 = dmr[3];
    const int dl[16] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      // select without dynamic register indexing
      int v = vtx[0], q0 = dl[0], q1 = dl[1], q2 = dl[2], q3 = dl[12];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (perm[a] == k) { v = vtx[k]; q0 = dl[3 * k]; q1 = dl[3 * k + 1]; q2 = dl[3 * k + 2]; q3 = dl[12 + k]; }
      dof[a][0] = q0; dof[a][1] = q1; dof[a][2] = q2; dof[a][3] = q3;
      x[a][0] = xg[3 * (int64_t)v]; x[a][1] = xg[3 * (int64_t)v + 1]; x[a][2] = xg[3 * (int64_t)v + 2];
      u[a][0] = wv[q0]; u[a][1] = wv[q1]; u[a][2] = wv[q2]; p[a] = wv[q3];
    }
    double blk[4][16], fr[4];
    const bool has_bc = cell_bc && cell_bc[cell];
    if (WANT_J || !has_bc) p1tet_rowslab<WANT_J, WANT_F>(form, m == 0, x, u, p, blk, fr);
    else p1tet_rowslab<true, WANT_F>(form, m == 0, x, u, p, blk, fr);   // residual only, but lifting needs the Jacobian rows

    if (has_bc) {
      // Dirichlet handling at element level (assemble_matrix / apply_lifting semantics)
      bool rowbc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) rowbc[r] = bc_marker[dof[0][r]] != 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int32_t dj = dof[a][d];
          if (bc_marker[dj]) {
            const double delta = bc_value[dj] - ((d < 3) ? u[a][d] : p[a]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (WANT_F) fr[r] += blk[a][4 * r + d] * delta;   // lifting with the un-zeroed entry
              if (WANT_J) blk[a][4 * r + d] = 0.0;               // constrained trial column
            }
          }
        }
      if (WANT_J) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (rowbc[r]) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int d = 0; d < 4; ++d) blk[a][4 * r + d] = 0.0;   // constrained test row
          }
      }
    }
    if (WANT_J) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          stageJ[(a * 4 + r) * P1_CAP + tid] = make_double4(blk[a][4 * r], blk[a][4 * r + 1], blk[a][4 * r + 2], blk[a][4 * r + 3]);
    }
    if (WANT_F) stageF[tid] = make_double4(fr[0], fr[1], fr[2], fr[3]);
  }
  __syncthreads();

  // ---------------- phase B: one (vertex, row, slot) piece per thread ----------------
  if (WANT_J) {
    const int nitems = 4 * slotp[nent];
    for (int item = tid; item < nitems; item += P1_CAP) {
      // entity le with 4*slotp[le] <= item < 4*slotp[le+1]
      int lo = 0, hi = nent - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (4 * slotp[mid] <= item) lo = mid; else hi = mid - 1;
      }
      const int le = lo;
      const int ns = slotp[le + 1] - slotp[le];
      const int rem = item - 4 * slotp[le];
      const int r = rem / ns, s = rem - r * ns;
      const uint32_t pat = (uint32_t)s * 0x01010101u;
      double4 acc = make_double4(0.0, 0.0, 0.0, 0.0);
      for (int ii = incp[le]; ii < incp[le + 1]; ++ii) {
        const uint32_t eq = __vcmpeq4(slots[ii], pat);
        if (eq) {
          const int a = (__ffs(eq) - 1) >> 3;
          const double4 v = stageJ[(a * 4 + r) * P1_CAP + ii];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      double* dst = vals + rowpos[(e0 + le) * 4 + r] + 4 * s;
      __stcs(reinterpret_cast<double2*>(dst), make_double2(acc.x, acc.y));
      __stcs(reinterpret_cast<double2*>(dst) + 1, make_double2(acc.z, acc.w));
    }
  }
  if (WANT_F) {
    for (int item = tid; item < 4 * nent; item += P1_CAP) {
      const int le = item >> 2, r = item & 3;
      double acc = 0.0;
      for (int ii = incp[le]; ii < incp[le + 1]; ++ii) {
        const double4 v = stageF[ii];
        acc += (r == 0) ? v.x : (r == 1) ? v.y : (r == 2) ? v.z : v.w;
      }
      F[rowdof[(e0 + le) * 4 + r]] = acc;
    }
  }
}

static size_t smem_bytes(bool want_J) {
  return (want_J ? 16 * P1_CAP * sizeof(double4) : 0) + P1_CAP * sizeof(double4) + P1_CAP * sizeof(uint32_t) + 2 * (P1_CAP + 1) * sizeof(int);
}

// ------------------------------------------------------------------------------------------ host side
void p1tet_free(nsgpu_ctx* ctx) {
  nsgpu_p1tet_plan* P = ctx->p1plan;
  if (!P) return;
  cudaFree(P->d_inc_cell); cudaFree(P->d_inc_slot); cudaFree(P->d_inc_ptr); cudaFree(P->d_slot_ptr);
  cudaFree(P->d_rowpos); cudaFree(P->d_rowdof); cudaFree(P->d_tile_ent); cudaFree(P->d_cell_bc);
  delete P;
  ctx->p1plan = nullptr;
}

void p1tet_mark_bc_dirty(nsgpu_ctx* ctx) {
  if (ctx->p1plan) ctx->p1plan->bc_dirty = true;
}

bool p1tet_fast_available(nsgpu_ctx* ctx) {
  if (ctx->gdim != 3 || ctx->vdeg != 1 || !ctx->pattern_built || !ctx->rows_presorted || !ctx->d_pairs) return false;
  if (ctx->n_cells_owned >= ((int64_t)1 << 29)) return false;
  if (!ctx->p1plan) {
    if (p1tet_build_plan(ctx) != NSGPU_OK) return false;
  }
  return ctx->p1plan != nullptr && ctx->p1plan->n_tiles >= 0;
}

static inline unsigned g256(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, 256); }

int p1tet_build_plan(nsgpu_ctx* ctx) {
  cudaStream_t s = ctx->stream;
  p1tet_free(ctx);
  nsgpu_p1tet_plan* P = new nsgpu_p1tet_plan();
  const int64_t n_inc = ctx->n_cells_owned * 4;
  P->n_inc = n_inc;
  P->cap = P1_CAP;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  uint32_t* d_leader = nullptr;
  int64_t *d_cnt = nullptr, *d_nrun = nullptr, *d_nslots = nullptr, *d_max = nullptr;
  void* d_tmp = nullptr;
  int* d_flag = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_leader); cudaFree(d_cnt); cudaFree(d_nrun); cudaFree(d_nslots);
    cudaFree(d_max); cudaFree(d_tmp); cudaFree(d_flag);
  };
#define PL_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      set_error(ctx, std::string("p1tet plan: ") + #call + ": " + cudaGetErrorString(e__));        \
      cleanup();                                                                                   \
      ctx->p1plan = P; p1tet_free(ctx);                                                            \
      return NSGPU_ECUDA;                                                                          \
    }                                                                                              \
  } while (0)
  if (n_inc == 0) { P->n_tiles = 0; ctx->p1plan = P; return NSGPU_OK; }

  PL_CUDA(cudaMalloc(&d_keys, sizeof(uint64_t) * n_inc));
  PL_CUDA(cudaMalloc(&d_keys2, sizeof(uint64_t) * n_inc));
  k_inc_keys<<<g256(n_inc), 256, 0, s>>>(ctx->n_cells_owned, ctx->d_dofmap, d_keys);
  int key_bits = 32;
  while (key_bits > 1 && !((uint64_t)(ctx->n_dofs - 1) >> (key_bits - 1))) --key_bits;
  size_t tb = 0;
  PL_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_keys, d_keys2, n_inc, 0, 32 + key_bits, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tb, d_keys, d_keys2, n_inc, 0, 32 + key_bits, s));
  cudaFree(d_tmp); d_tmp = nullptr;
  cudaFree(d_keys); d_keys = nullptr;

  // run-length encode the row vertices -> entity list + incidence counts
  PL_CUDA(cudaMalloc(&d_leader, sizeof(uint32_t) * n_inc));
  PL_CUDA(cudaMalloc(&d_cnt, sizeof(int64_t) * (n_inc + 1)));
  PL_CUDA(cudaMalloc(&d_nrun, sizeof(int64_t)));
  cub::TransformInputIterator<uint32_t, HiWord, const uint64_t*> hi_it(d_keys2, HiWord());
  tb = 0;
  PL_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb, hi_it, d_leader, d_cnt, d_nrun, n_inc, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceRunLengthEncode::Encode(d_tmp, tb, hi_it, d_leader, d_cnt, d_nrun, n_inc, s));
  int64_t n_ent = 0;
  PL_CUDA(cudaMemcpyAsync(&n_ent, d_nrun, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  PL_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_tmp); d_tmp = nullptr;
  P->n_ent = n_ent;
  PL_CUDA(cudaMemsetAsync(d_cnt + n_ent, 0, sizeof(int64_t), s));

  PL_CUDA(cudaMalloc(&d_max, sizeof(int64_t)));
  tb = 0;
  PL_CUDA(cub::DeviceReduce::Max(nullptr, tb, d_cnt, d_max, n_ent, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceReduce::Max(d_tmp, tb, d_cnt, d_max, n_ent, s));
  int64_t maxdeg = 0;
  PL_CUDA(cudaMemcpyAsync(&maxdeg, d_max, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  PL_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_tmp); d_tmp = nullptr;
  P->maxdeg = (int)maxdeg;
  if (maxdeg > P1_CAP / 2) {   // pathological vertex degree: keep the generic path
    cleanup();
    ctx->p1plan = P; p1tet_free(ctx);
    return NSGPU_EUNSUPPORTED;
  }

  PL_CUDA(cudaMalloc(&P->d_inc_ptr, sizeof(int64_t) * (n_ent + 1)));
  tb = 0;
  PL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt, P->d_inc_ptr, n_ent + 1, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_cnt, P->d_inc_ptr, n_ent + 1, s));
  PL_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_tmp); d_tmp = nullptr;

  // per-incidence cell / slot words
  PL_CUDA(cudaMalloc(&P->d_inc_cell, sizeof(uint32_t) * n_inc));
  PL_CUDA(cudaMalloc(&P->d_inc_slot, sizeof(uint32_t) * n_inc));
  PL_CUDA(cudaMalloc(&d_flag, sizeof(int)));
  PL_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
  k_inc_fill<<<g256(n_inc), 256, 0, s>>>(n_inc, d_keys2, ctx->d_dofmap, ctx->d_pairs, ctx->d_pair_first, ctx->d_pair_last,
                                         P->d_inc_cell, P->d_inc_slot, d_flag);

  // per-entity output info
  PL_CUDA(cudaMalloc(&d_nslots, sizeof(int64_t) * (n_ent + 1)));
  PL_CUDA(cudaMalloc(&P->d_slot_ptr, sizeof(int64_t) * (n_ent + 1)));
  PL_CUDA(cudaMalloc(&P->d_rowpos, sizeof(int64_t) * n_ent * 4));
  PL_CUDA(cudaMalloc(&P->d_rowdof, sizeof(int32_t) * n_ent * 4));
  k_ent_info<<<g256(n_ent + 1), 256, 0, s>>>(n_ent, d_leader, ctx->d_members, ctx->d_pair_first, ctx->d_pair_last, ctx->d_indptr,
                                             d_nslots, P->d_rowpos, P->d_rowdof);
  tb = 0;
  PL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_nslots, P->d_slot_ptr, n_ent + 1, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_nslots, P->d_slot_ptr, n_ent + 1, s));
  PL_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_tmp); d_tmp = nullptr;

  // tiles: entity e belongs to tile floor(inc_ptr[e] / capeff)
  const int64_t capeff = P1_CAP - maxdeg + 1;
  P->n_tiles = ceil_div(n_inc, capeff);
  PL_CUDA(cudaMalloc(&P->d_tile_ent, sizeof(int64_t) * (P->n_tiles + 1)));
  k_tiles<<<g256(P->n_tiles + 1), 256, 0, s>>>(P->n_tiles, n_ent, capeff, P->d_inc_ptr, P->d_tile_ent);
  PL_CUDA(cudaMalloc(&P->d_cell_bc, ctx->n_cells_owned > 0 ? ctx->n_cells_owned : 1));
  int bad = 0;
  PL_CUDA(cudaMemcpyAsync(&bad, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  PL_CUDA(cudaStreamSynchronize(s));
  PL_CUDA(cudaGetLastError());
  ctx->launches += 12;
  cleanup();
  if (bad) {   // a vertex has more than 255 neighbours: generic path
    ctx->p1plan = P; p1tet_free(ctx);
    return NSGPU_EUNSUPPORTED;
  }
  P->bc_dirty = true;
  ctx->p1plan = P;

  PL_CUDA(cudaFuncSetAttribute(k_p1tet_tiles<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(true)));
  PL_CUDA(cudaFuncSetAttribute(k_p1tet_tiles<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(true)));
  PL_CUDA(cudaFuncSetAttribute(k_p1tet_tiles<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(false)));
  return NSGPU_OK;
#undef PL_CUDA
}

int p1tet_assemble(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_Fout) {
  nsgpu_p1tet_plan* P = ctx->p1plan;
  if (!P) { set_error(ctx, "p1tet plan missing"); return NSGPU_EINVAL; }
  cudaStream_t s = ctx->stream;
  if (P->n_tiles == 0) return NSGPU_OK;
  if (ctx->has_bc && P->bc_dirty) {
    k_cell_bc<<<g256(ctx->n_cells_owned), 256, 0, s>>>(ctx->n_cells_owned, ctx->d_dofmap, ctx->d_bc_marker, P->d_cell_bc);
    P->bc_dirty = false;
    ctx->launches += 1;
  }
  const uint8_t* cbc = ctx->has_bc ? P->d_cell_bc : nullptr;
#define P1_LAUNCH(J, F)                                                                                             \
  k_p1tet_tiles<J, F><<<(unsigned)P->n_tiles, P1_CAP, smem_bytes(J), s>>>(ctx->form, ctx->d_x, ctx->d_cells, ctx->d_dofmap, d_xin, \
      ctx->d_bc_marker, ctx->d_bc_value, cbc, P->d_inc_cell, P->d_inc_slot, P->d_inc_ptr, P->d_slot_ptr, P->d_rowpos,   \
      P->d_rowdof, P->d_tile_ent, ctx->d_vals, d_Fout)
  if (want_J && want_F) P1_LAUNCH(true, true);
  else if (want_J) P1_LAUNCH(true, false);
  else P1_LAUNCH(false, true);
#undef P1_LAUNCH
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

}  // namespace nsgpu
