// rowown.cu -- atomics-free "row-owner" assembly for every element pair / form the factorised P1-P1 tet kernel does not
// cover: P2-P1 tetrahedra (the reference's second pair: LidDrivenNavierStokesFlow.py:46, StokesFlow/DuctStokesFlow.py:147-154),
// UGN triangles (LidDrivenNavierStokesFlow.py:112-146), the Stokes operators.  Same dolfinx semantics as assemble.cu
// (assemble_matrix with element-level Dirichlet row / column zeroing, assemble_vector + apply_lifting).
//
// Who computes what.  All dofs of a mesh entity (a vertex: GD velocity components + pressure; a P2 edge: GD velocity
// components) share their CSR column set, so an entity's rows are owned by one GROUP of NENT lanes (NENT = entities per
// cell = 10 on P2-P1 tets; three groups per warp).  The group walks the entity's incident cells; in each cell lane n evaluates the
// block (entity, n-th entity of the cell) -- GD x GD velocity entries plus the pressure row / column where the two are
// vertices -- with the factorised formulas of element_block.cuh and adds it into a shared-memory copy of the entity's rows at the
// positions the pattern build tabulated (rel).  Lanes of a group hit distinct columns, cells are taken one after the other,
// so there are no atomics and the sums are bitwise reproducible.  When the last cell is done the rows are written to the CSR
// value array with plain stores, every entry exactly once (no zero-fill pass), and the residual entries likewise.
//
// The point data (fields, stabilisation parameters and their derivatives at the quadrature points) is computed once per
// (cell, point) by a pre-pass and read through L1/L2 by the ~NENT groups that visit the cell: 1 KB per cell instead of the
// element matrix (9 KB per P2-P1 cell) ever leaving the SM.
#include <cub/cub.cuh>

#include "common.cuh"
#include "element_block.cuh"

namespace nsgpu {

struct RowOwnPlan {
  int64_t n_inc = 0, n_ent[2] = {0, 0};
  int lmax[2] = {0, 0};
  uint32_t* d_inc = nullptr;        // incidences (cell << 4 | local entity), grouped by entity
  int4* d_ent_hdr = nullptr;        // per entity two int4 (vertex entities first; inside a class by falling incidence count, so that the
                                    //  groups of a warp walk equally many cells): {first incidence lo, hi, incidences, row length}, {row dofs}
  double* d_prec = nullptr;         // [cell][NQ point records of PREC doubles | cell record of CREC doubles]: one contiguous piece per cell
  uint8_t* d_cellbc = nullptr;      // cell holds a constrained dof
  bool unsupported = false;
  bool attr_set = false;
};

template <int GD, int VDEG>
__global__ void k_ro_keys(int64_t n_cells, const int32_t* __restrict__ dofmap, uint64_t* keys) {
  using T = ElemTraits<GD, VDEG>;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_cells * T::NENT) return;
  const int64_t cell = t / T::NENT;
  const int m = (int)(t - cell * T::NENT);
  const uint64_t leader = (uint32_t)dofmap[cell * T::ND + GD * m];
  const uint64_t cls = m <= GD ? 0 : 1;
  keys[t] = (cls << 63) | (leader << 32) | ((uint64_t)cell << 4) | (uint64_t)m;
}

__global__ void k_ro_heads(int64_t n, const uint64_t* __restrict__ keys, int64_t* head) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  head[i] = (i < n && (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32))) ? 1 : 0;
}

// entity starts, incidence words, class counts, longest row per class
__global__ void k_ro_scatter(int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ head, const int64_t* __restrict__ pos,
                             const int64_t* __restrict__ indptr, int64_t* ent_start, uint32_t* inc, unsigned long long* n_cls0, int* lmax) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { ent_start[pos[n]] = n; return; }
  inc[i] = (uint32_t)(keys[i] & 0xffffffffu);
  if (head[i]) {
    ent_start[pos[i]] = i;
    const int cls = (int)(keys[i] >> 63);
    const int64_t leader = (int64_t)((keys[i] >> 32) & 0x7fffffffu);
    if (cls == 0) atomicAdd(n_cls0, 1ull);
    atomicMax(lmax + cls, (int)(indptr[leader + 1] - indptr[leader]));
  }
}

// entity order: class, then falling incidence count, then leader (= the order of the first sort)
__global__ void k_ro_key2(int64_t n_ent, const int64_t* __restrict__ start, const uint64_t* __restrict__ keys, uint64_t* key2) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_ent) return;
  const int64_t cnt = start[e + 1] - start[e];
  const uint64_t cls = keys[start[e]] >> 63;
  key2[e] = (cls << 60) | ((uint64_t)(255 - (cnt > 255 ? 255 : cnt)) << 48) | (uint64_t)e;
}

__global__ void k_ro_perm(int64_t n_ent, const uint64_t* __restrict__ key2, const int64_t* __restrict__ start, const uint32_t* __restrict__ inc,
                          const int32_t* __restrict__ dofmap, const int64_t* __restrict__ indptr, int gd, int nd, int poff, int4* hdr) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_ent) return;
  const int64_t e = (int64_t)(key2[i] & ((uint64_t(1) << 48) - 1));
  const int64_t s0 = start[e];
  const uint32_t w0 = inc[s0];
  const int32_t* dm = dofmap + (int64_t)(w0 >> 4) * nd;
  const int m0 = (int)(w0 & 15u);
  int gi[4] = {0, 0, 0, 0};
  for (int r = 0; r < gd; ++r) gi[r] = dm[gd * m0 + r];
  if (m0 <= gd) gi[gd] = dm[poff + m0];
  hdr[2 * i] = make_int4((int)(s0 & 0xffffffff), (int)(s0 >> 32), (int)(start[e + 1] - s0), (int)(indptr[gi[0] + 1] - indptr[gi[0]]));
  hdr[2 * i + 1] = make_int4(gi[0], gi[1], gi[2], gi[3]);
}

// pre-pass: point records and cell records (element_block.cuh) of every owned cell
template <int GD, int VDEG>
__global__ void __launch_bounds__(128)
k_ro_points(int64_t n_cells, FormParams form, const double* __restrict__ xg, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap,
            const double* __restrict__ wv, const uint8_t* __restrict__ marker, double* __restrict__ prec, uint8_t* __restrict__ cellbc) {
  using T = ElemTraits<GD, VDEG>;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_cells * T::NQ) return;
  const int64_t cell = t / T::NQ;
  const int q = (int)(t - cell * T::NQ);
  double x[3 * (GD + 1)], w[T::ND];
  for (int a = 0; a <= GD; ++a) {
    const int64_t v = cells[cell * (GD + 1) + a];
    for (int i = 0; i < 3; ++i) x[3 * a + i] = xg[3 * v + i];
  }
  bool bc = false;
  for (int k = 0; k < T::ND; ++k) {
    const int32_t d = dofmap[cell * T::ND + k];
    w[k] = wv[d];
    if (marker) bc |= marker[d] != 0;
  }
  double rec[PREC], cr[CREC];
  point_record<GD, VDEG>(form, x, w, q, rec, cr);
  double* o = prec + cell * (T::NQ * PREC + CREC) + q * PREC;
  for (int k = 0; k < PREC; ++k) o[k] = rec[k];
  if (q == 0) {
    double* oc = prec + cell * (T::NQ * PREC + CREC) + T::NQ * PREC;
    for (int k = 0; k < CREC; ++k) oc[k] = cr[k];
    cellbc[cell] = bc ? 1 : 0;
  }
}

constexpr int RO_PAD = 4;   // doubles between the per-group copies of the staged records (bank spreading)
constexpr int RO_RPAD = 2;  // doubles between consecutive point / row-side records of a group: lanes 0 .. NQ-1 read record q = lane in the row-side phase

struct RowOwnArgs {
  FormParams form;
  int64_t e0, e1;
  const int4* ent_hdr;
  const uint32_t* inc;
  const int32_t* dofmap;
  const uint16_t* rel;
  const int64_t* indptr;
  const double* prec;       // per cell: NQ point records, then the cell record
  const uint8_t* cellbc;
  const uint8_t* marker;     // nullptr without Dirichlet conditions
  const double* bc_value;
  const double* wv;
  double* vals;
  double* F;
  int lstride;               // doubles per accumulator row (>= longest row of the class)
};

__device__ __forceinline__ void ro_cp16(void* dst_shared, const void* src_global) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void ro_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ro_cp_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// LEAN: the G-metric form through gm_row_side / gm_block (element_block.cuh): lanes 0 .. NQ-1 of a group evaluate the row-side
// records of (entity, point) once per cell -- the residual entries come with them --, every lane then runs the short mixed part.
template <int GD, int VDEG, bool VCLASS, bool WANT_J, bool WANT_F, bool LEAN>
// (P2 vertex entities: the 4 x ~210 accumulator rows per group limit a CTA to two warps and an SM to ~8 warps, so the register budget can be 255)
__global__ void __launch_bounds__(128, VDEG == 2 ? (VCLASS ? 2 : 3) : 4)
k_rowown(RowOwnArgs a) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int NENT = T::NENT, ND = T::ND, NV = GD + 1, NQ = T::NQ, POFF = T::POFF;
  constexpr int GPW = 32 / NENT;               // groups per warp
  constexpr int R = VCLASS ? GD + 1 : GD;      // rows of the entity
  constexpr int SREC = PREC + RO_RPAD, SRS = RSIDE + RO_RPAD;   // record strides in shared memory (lane q of a group works on record q)
  constexpr int STG = NQ * SREC + CREC;        // doubles of one staged cell: its point records and its cell record
  // the lanes of a group read the same staged word (broadcast), the groups of a warp their own copies: 4 doubles of padding per
  // group put the three copies into different 16-byte bank groups (unpadded strides are multiples of 128 bytes: 3-way conflicts)
  constexpr int STGP = 2 * STG + RO_PAD, RSP = NQ * SRS + RO_PAD;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gw = lane / NENT, n = lane - gw * NENT;
  const bool active = gw < GPW;
  const int G = nwarps * GPW;
  const int g = warp * GPW + (active ? gw : 0);
  double* acc = smem + (size_t)g * R * a.lstride;
  double* red = smem + (size_t)G * R * a.lstride + (size_t)g * NENT * 4;
  double* stage = smem + (size_t)G * R * a.lstride + (size_t)G * NENT * 4 + (size_t)g * STGP;   // double buffer of the group
  double* rsd = smem + (size_t)G * R * a.lstride + (size_t)G * NENT * 4 + (size_t)G * STGP + (size_t)g * RSP;   // row-side records (LEAN)
  const unsigned FULL = 0xffffffffu;
  const unsigned gmask = active ? (((1u << NENT) - 1u) << (gw * NENT)) : 0u;
  // the group's lanes copy the records of one cell (16-byte pieces, cp.async) into stage buffer b
  auto fetch = [&](int64_t cell, int b) {
    double* dst = stage + b * STG;
    const double* src = a.prec + cell * (NQ * PREC + CREC);
#pragma unroll
    for (int j = 0; j < ((NQ * PREC + CREC) / 2 + NENT - 1) / NENT; ++j) {
      const int k = n + j * NENT;                      // 16-byte piece k of the cell's records; record k / (PREC / 2) is shifted by its padding
      if (k < (NQ * PREC + CREC) / 2) ro_cp16(dst + 2 * k + min(k / (PREC / 2), NQ) * RO_RPAD, src + 2 * k);
    }
  };

  if (WANT_J) {
    for (int k = threadIdx.x; k < G * R * a.lstride; k += blockDim.x) smem[k] = 0.0;
    __syncthreads();
  }

  // entity headers are fetched one round ahead
  const int64_t first = a.e0 + (int64_t)blockIdx.x * G, stride = (int64_t)gridDim.x * G;
  int4 h0 = make_int4(0, 0, 0, 0), h1 = make_int4(0, 0, 0, 0);
  if (active && first + g < a.e1) { h0 = a.ent_hdr[2 * (first + g)]; h1 = a.ent_hdr[2 * (first + g) + 1]; }
  for (int64_t base = first; base < a.e1; base += stride) {
    const int64_t ent = base + g;
    const bool has = active && ent < a.e1;
    const int64_t s0 = has ? ((int64_t)(uint32_t)h0.x | ((int64_t)h0.y << 32)) : 0;
    const int cnt = has ? h0.z : 0;
    const int L = has ? h0.w : 0;
    int32_t gi[R];
    gi[0] = h1.x; gi[1] = h1.y;
    if (R > 2) gi[2] = h1.z;
    if (R > 3) gi[R - 1] = h1.w;
    if (active && ent + stride < a.e1) { h0 = a.ent_hdr[2 * (ent + stride)]; h1 = a.ent_hdr[2 * (ent + stride) + 1]; }
    const int maxcnt = __reduce_max_sync(FULL, cnt);
    const int maxL = __reduce_max_sync(FULL, L);
    bool rmk[R];
    int64_t rstart[R];       // CSR starts of the entity's rows: asked for now, needed when the last cell is done
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rmk[r] = (has && a.marker) ? a.marker[gi[r]] != 0 : false;
      rstart[r] = (WANT_J && has) ? a.indptr[gi[r]] : 0;
    }
    double lift[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lift[r] = 0.0;
    double brow = 0.0;
    double brow4[R];          // LEAN: residual contributions of this lane's quadrature point, all rows of the entity
#pragma unroll
    for (int r = 0; r < R; ++r) brow4[r] = 0.0;

    // the entity's incidence words: up to 3 NENT of them live in the group's registers and are handed round by shuffle
    uint32_t iw0 = 0, iw1 = 0, iw2 = 0;
    if (n < cnt) iw0 = a.inc[s0 + n];
    if (n + NENT < cnt) iw1 = a.inc[s0 + n + NENT];
    if (n + 2 * NENT < cnt) iw2 = a.inc[s0 + n + 2 * NENT];
    auto word = [&](int k) -> uint32_t {     // incidence k of this lane's entity; k is warp-uniform
      const int j = k / NENT, src = (active ? gw : 0) * NENT + (k - j * NENT);
      const uint32_t v = __shfl_sync(FULL, j == 0 ? iw0 : (j == 1 ? iw1 : iw2), src);
      return (j < 3 || k >= cnt) ? v : a.inc[s0 + k];
    };
    uint32_t wnext = word(0);
    if (cnt > 0) fetch((int64_t)(wnext >> 4), 0);
    ro_cp_commit();
    for (int it = 0; it < maxcnt; ++it) {
      const uint32_t wd = wnext;
      wnext = word(it + 1);
      if (it + 1 < cnt) fetch((int64_t)(wnext >> 4), (it + 1) & 1);   // the next cell's records travel while this one is evaluated
      ro_cp_commit();
      ro_cp_wait1();
      __syncwarp();
      if (it < cnt) {
        const int64_t cell = wd >> 4;
        const int m = (int)(wd & 15u);
        const double* pr = stage + (it & 1) * STG;
        const double* cr = pr + NQ * SREC;
        const bool cbc = a.marker && a.cellbc[cell];
        if (LEAN) {
          if (n < NQ) {
            double bq[R];
            gm_row_side<GD, VDEG, VCLASS>(a.form, pr + SREC * n, cr, m, n, rsd + SRS * n, WANT_F ? bq : nullptr);
            if (WANT_F) {
#pragma unroll
              for (int r = 0; r < R; ++r) brow4[r] += bq[r];
            }
          }
          __syncwarp(gmask);
        }
        if (WANT_J || cbc) {
          const int32_t* dm = a.dofmap + cell * ND;
          const uint16_t* rp = a.rel + (cell * NENT + m) * ND;
          int ov[GD], op = 0;     // positions of this lane's columns in the entity's rows: asked for before the block is evaluated
          if (WANT_J) {
#pragma unroll
            for (int d = 0; d < GD; ++d) ov[d] = rp[GD * n + d];
            if (n < NV) op = rp[POFF + n];
          }
          EntityBlock<GD> B;
          if (LEAN) gm_block<GD, VDEG, VCLASS>(a.form, pr, cr, rsd, m, n, B, SREC, SRS);
          else {
            entity_block<GD, VDEG, true, WANT_F, VCLASS ? 1 : 0>(a.form, pr, cr, m, n, B, n, SREC);
            if (WANT_F && n < R) brow += B.b;
          }
          bool mk[GD], mkp = false;
#pragma unroll
          for (int d = 0; d < GD; ++d) mk[d] = false;
          if (cbc) {
#pragma unroll
            for (int d = 0; d < GD; ++d) {
              const int32_t cd = dm[GD * n + d];
              mk[d] = a.marker[cd] != 0;
              if (WANT_F && mk[d]) {     // apply_lifting: b -= A[:, bc] (x - g), un-zeroed column of this cell
                const double dg = a.bc_value[cd] - a.wv[cd];
#pragma unroll
                for (int c = 0; c < GD; ++c) lift[c] += B.vv[c][d] * dg;
                if (VCLASS) lift[R - 1] += B.pv[d] * dg;
              }
            }
            if (n < NV) {
              const int32_t cp = dm[POFF + n];
              mkp = a.marker[cp] != 0;
              if (WANT_F && mkp) {
                const double dg = a.bc_value[cp] - a.wv[cp];
#pragma unroll
                for (int c = 0; c < GD; ++c) lift[c] += B.vp[c] * dg;
                if (VCLASS) lift[R - 1] += B.pp * dg;
              }
            }
          }
          if (WANT_J) {
            // The (row, column) positions of one lane are pairwise distinct, and so are the columns of different lanes: all loads of a
            // batch are issued before its first store (written as read-modify-writes one after the other the compiler has to keep
            // every load behind the previous store: 16 dependent shared-memory round trips per cell).
            double old[R][GD];
#pragma unroll
            for (int d = 0; d < GD; ++d)
#pragma unroll
              for (int r = 0; r < R; ++r) old[r][d] = acc[r * a.lstride + ov[d]];
#pragma unroll
            for (int d = 0; d < GD; ++d) {
              if (mk[d]) continue;
#pragma unroll
              for (int c = 0; c < GD; ++c)
                if (!rmk[c]) acc[c * a.lstride + ov[d]] = old[c][d] + B.vv[c][d];
              if (VCLASS && !rmk[R - 1]) acc[(R - 1) * a.lstride + ov[d]] = old[R - 1][d] + B.pv[d];
            }
            if (n < NV && !mkp) {
              double oldp[R];
#pragma unroll
              for (int r = 0; r < R; ++r) oldp[r] = acc[r * a.lstride + op];
#pragma unroll
              for (int c = 0; c < GD; ++c)
                if (!rmk[c]) acc[c * a.lstride + op] = oldp[c] + B.vp[c];
              if (VCLASS && !rmk[R - 1]) acc[(R - 1) * a.lstride + op] = oldp[R - 1] + B.pp;
            }
          }
        }
        else if (WANT_F && !LEAN) {   // residual only, no constrained dof in the cell: the row-side part of the block routine
          EntityBlock<GD> B;
          entity_block<GD, VDEG, false, true, VCLASS ? 1 : 0>(a.form, pr, cr, m, n < R ? n : 0, B, n, SREC);
          if (n < R) brow += B.b;
        }
      }
      __syncwarp();
    }

    if (WANT_F) {
      if (has) {
#pragma unroll
        for (int r = 0; r < R; ++r) red[n * 4 + r] = lift[r] + (LEAN ? (n < NQ ? brow4[r] : 0.0) : (n == r ? brow : 0.0));
      }
      __syncwarp();
      if (has && n < R) {
        double s = 0.0;
        for (int k = 0; k < NENT; ++k) s += red[k * 4 + n];
        int32_t row = gi[0];
#pragma unroll
        for (int r = 1; r < R; ++r) row = (n == r) ? gi[r] : row;
        a.F[row] = s;
      }
      __syncwarp();
    }
    if (WANT_J) {
      for (int k = n; k < maxL; k += NENT) {
        if (k < L) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            a.vals[rstart[r] + k] = acc[r * a.lstride + k];
            acc[r * a.lstride + k] = 0.0;
          }
        }
      }
      __syncwarp();
    }
  }
}

void rowown_free(nsgpu_ctx* ctx) {
  RowOwnPlan* P = static_cast<RowOwnPlan*>(ctx->rowown_plan);
  if (!P) return;
  cudaFree(P->d_inc); cudaFree(P->d_ent_hdr); cudaFree(P->d_prec); cudaFree(P->d_cellbc);
  delete P;
  ctx->rowown_plan = nullptr;
}

template <int GD, int VDEG>
static int rowown_build(nsgpu_ctx* ctx) {
  using T = ElemTraits<GD, VDEG>;
  rowown_free(ctx);
  RowOwnPlan* P = new RowOwnPlan();
  ctx->rowown_plan = P;
  cudaStream_t s = ctx->stream;
  const int64_t nc = ctx->n_cells_owned, n = nc * T::NENT;
  if (nc <= 0 || nc >= (int64_t(1) << 28)) { P->unsupported = true; return NSGPU_OK; }
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  int64_t *d_head = nullptr, *d_pos = nullptr, *d_start = nullptr;
  uint64_t *d_k2 = nullptr, *d_k2s = nullptr;
  unsigned long long* d_cnt = nullptr;
  int* d_lmax = nullptr;
  void* d_tmp = nullptr;
  auto cleanup = [&]() { cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_head); cudaFree(d_pos); cudaFree(d_cnt); cudaFree(d_lmax); cudaFree(d_tmp); cudaFree(d_start); cudaFree(d_k2); cudaFree(d_k2s); };
#define RO_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      set_error(ctx, std::string("rowown plan: " #call ": ") + cudaGetErrorString(e__));           \
      cleanup(); rowown_free(ctx);                                                                 \
      return NSGPU_ECUDA;                                                                          \
    }                                                                                              \
  } while (0)
  RO_CUDA(cudaMalloc(&d_keys, sizeof(uint64_t) * n));
  RO_CUDA(cudaMalloc(&d_keys2, sizeof(uint64_t) * n));
  RO_CUDA(cudaMalloc(&d_head, sizeof(int64_t) * (n + 1)));
  RO_CUDA(cudaMalloc(&d_pos, sizeof(int64_t) * (n + 1)));
  RO_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long)));
  RO_CUDA(cudaMalloc(&d_lmax, 2 * sizeof(int)));
  RO_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
  RO_CUDA(cudaMemsetAsync(d_lmax, 0, 2 * sizeof(int), s));
  k_ro_keys<GD, VDEG><<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(nc, ctx->d_dofmap, d_keys);
  size_t tmp_bytes = 0;
  RO_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_keys2, n, 0, 64, s));
  RO_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
  RO_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys2, n, 0, 64, s));
  k_ro_heads<<<(unsigned)ceil_div(n + 1, 256), 256, 0, s>>>(n, d_keys2, d_head);
  cudaFree(d_tmp); d_tmp = nullptr;
  RO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_head, d_pos, n + 1, s));
  RO_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
  RO_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_head, d_pos, n + 1, s));
  int64_t n_ent = 0;
  RO_CUDA(cudaMemcpyAsync(&n_ent, d_pos + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RO_CUDA(cudaStreamSynchronize(s));
  int rc;
  RO_CUDA(cudaMalloc(&d_start, sizeof(int64_t) * (n_ent + 1)));
  RO_CUDA(cudaMalloc(&d_k2, sizeof(uint64_t) * n_ent));
  RO_CUDA(cudaMalloc(&d_k2s, sizeof(uint64_t) * n_ent));
  if ((rc = dev_alloc(ctx, &P->d_ent_hdr, 2 * n_ent)) || (rc = dev_alloc(ctx, &P->d_inc, n)) || (rc = dev_alloc(ctx, &P->d_prec, nc * (T::NQ * PREC + CREC))) ||
      (rc = dev_alloc(ctx, &P->d_cellbc, nc))) { cleanup(); rowown_free(ctx); return rc; }
  k_ro_scatter<<<(unsigned)ceil_div(n + 1, 256), 256, 0, s>>>(n, d_keys2, d_head, d_pos, ctx->d_indptr, d_start, P->d_inc, d_cnt, d_lmax);
  k_ro_key2<<<(unsigned)ceil_div(n_ent, 256), 256, 0, s>>>(n_ent, d_start, d_keys2, d_k2);
  cudaFree(d_tmp); d_tmp = nullptr;
  RO_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_k2, d_k2s, n_ent, 0, 62, s));
  RO_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
  RO_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_k2, d_k2s, n_ent, 0, 62, s));
  k_ro_perm<<<(unsigned)ceil_div(n_ent, 256), 256, 0, s>>>(n_ent, d_k2s, d_start, P->d_inc, ctx->d_dofmap, ctx->d_indptr, GD, T::ND, T::POFF, P->d_ent_hdr);
  unsigned long long n0 = 0;
  RO_CUDA(cudaMemcpyAsync(&n0, d_cnt, sizeof(n0), cudaMemcpyDeviceToHost, s));
  RO_CUDA(cudaMemcpyAsync(P->lmax, d_lmax, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  RO_CUDA(cudaStreamSynchronize(s));
  RO_CUDA(cudaGetLastError());
  ctx->launches += 5;
  cleanup();
#undef RO_CUDA
  P->n_inc = n;
  P->n_ent[0] = (int64_t)n0;
  P->n_ent[1] = n_ent - (int64_t)n0;
  return NSGPU_OK;
}

// shared memory of a launch: accumulator rows + reduction scratch per group
static size_t ro_smem(int groups, int rows, int lstride, int nent, int nq) {
  return sizeof(double) * ((size_t)groups * rows * lstride + (size_t)groups * nent * 4 + (size_t)groups * (2 * (nq * (PREC + RO_RPAD) + CREC) + RO_PAD) +
                           (size_t)groups * (nq * (RSIDE + RO_RPAD) + RO_PAD));
}

template <int GD, int VDEG, bool VCLASS>
static int rowown_launch_class(nsgpu_ctx* ctx, RowOwnPlan* P, RowOwnArgs a, bool want_J, bool want_F) {
  using T = ElemTraits<GD, VDEG>;
  constexpr int GPW = 32 / T::NENT, R = VCLASS ? GD + 1 : GD;
  const int cls = VCLASS ? 0 : 1;
  if (P->n_ent[cls] == 0) return NSGPU_OK;
  a.e0 = VCLASS ? 0 : P->n_ent[0];
  a.e1 = a.e0 + P->n_ent[cls];
  a.lstride = want_J ? (P->lmax[cls] + 3) & ~3 : 0;
  // warps per CTA: the choice that keeps most warps resident on an SM (227 KB of shared memory; the P2 vertex class is shared-memory
  // bound: ~30 KB per warp of accumulator rows and staged records)
  int warps = 4, best = 0;
  for (int w = 4; w >= 1; w >>= 1) {
    const size_t need = ro_smem(w * GPW, R, a.lstride, T::NENT, T::NQ) + 1024;
    const int regcap = VDEG == 2 ? (VCLASS ? 8 : 12) : 16;          // warps the register file holds at the launch bounds of k_rowown
    const int resident = std::min((int)std::min<size_t>(227 * 1024 / need, 32) * w, regcap);
    if (4 * resident > 5 * best) { best = resident; warps = w; }   // a smaller CTA only for >= 25 % more resident warps (measured: near-ties favour the larger CTA)
  }
  const size_t smem = ro_smem(warps * GPW, R, a.lstride, T::NENT, T::NQ);
  const int G = warps * GPW;
  const int64_t need = ceil_div(P->n_ent[cls], G);
  const int64_t cap = (int64_t)ctx->n_sms * 16;
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  // the lean G-metric path is compiled for tetrahedra (the reference's G-metric form lives on the 3-D channel / duct meshes)
  constexpr bool CAN_LEAN = GD == 3;
  const bool lean = CAN_LEAN && a.form.flavour == NSGPU_FORM_GMETRIC && ctx->rowown_lean;
#define RO_LAUNCH1(J, F, L)                                                                                                              \
  do {                                                                                                                                  \
    if (smem > 48 * 1024)                                                                                                               \
      NS_CUDA(ctx, cudaFuncSetAttribute(k_rowown<GD, VDEG, VCLASS, J, F, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    k_rowown<GD, VDEG, VCLASS, J, F, L><<<grid, warps * 32, smem, ctx->stream>>>(a);                                                    \
  } while (0)
#define RO_LAUNCH(J, F)                                  \
  do {                                                   \
    if (lean) RO_LAUNCH1(J, F, CAN_LEAN);                \
    else RO_LAUNCH1(J, F, false);                        \
  } while (0)
  if (want_J && want_F) RO_LAUNCH(true, true);
  else if (want_J) RO_LAUNCH(true, false);
  else RO_LAUNCH(false, true);
#undef RO_LAUNCH
#undef RO_LAUNCH1
  ctx->launches += 1;
  return NSGPU_OK;
}

template <int GD, int VDEG>
static int rowown_run(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_Fout) {
  using T = ElemTraits<GD, VDEG>;
  RowOwnPlan* P = static_cast<RowOwnPlan*>(ctx->rowown_plan);
  const int64_t nc = ctx->n_cells_owned;
  const uint8_t* mk = ctx->has_bc ? ctx->d_bc_marker : nullptr;
  k_ro_points<GD, VDEG><<<(unsigned)ceil_div(nc * T::NQ, 128), 128, 0, ctx->stream>>>(nc, ctx->form, ctx->d_x, ctx->d_cells, ctx->d_dofmap, d_xin, mk,
                                                                                      P->d_prec, P->d_cellbc);
  ctx->launches += 1;
  RowOwnArgs a;
  a.form = ctx->form;
  a.e0 = a.e1 = 0;
  a.ent_hdr = P->d_ent_hdr; a.inc = P->d_inc; a.dofmap = ctx->d_dofmap; a.rel = ctx->d_rel; a.indptr = ctx->d_indptr;
  a.prec = P->d_prec; a.cellbc = P->d_cellbc; a.marker = mk; a.bc_value = ctx->d_bc_value; a.wv = d_xin;
  a.vals = ctx->d_vals; a.F = d_Fout; a.lstride = 0;
  int rc = rowown_launch_class<GD, VDEG, true>(ctx, P, a, want_J, want_F);
  if (rc == NSGPU_OK && VDEG == 2) rc = rowown_launch_class<GD, VDEG, false>(ctx, P, a, want_J, want_F);
  return rc;
}

// does the row-owner kernel apply (builds the plan on first use)?
bool rowown_available(nsgpu_ctx* ctx) {
  // option rowown: 0 never, 1 (default) for P2-P1 spaces -- measured on B200 the cooperative kernel with atomics is still the
  // faster one for the small P1-P1 blocks (UGN triangles 914 vs 480 Mcells/s, P1-P1 Stokes tets 290 vs 200) --, 2 always
  if (!ctx->rowown || (ctx->rowown == 1 && ctx->vdeg != 2) || !ctx->pattern_built || !ctx->d_rel) return false;
  RowOwnPlan* P = static_cast<RowOwnPlan*>(ctx->rowown_plan);
  if (!P) {
    int rc = NSGPU_EUNSUPPORTED;
    switch (ctx->gdim * 10 + ctx->vdeg) {
      case 31: rc = rowown_build<3, 1>(ctx); break;
      case 32: rc = rowown_build<3, 2>(ctx); break;
      case 21: rc = rowown_build<2, 1>(ctx); break;
      case 22: rc = rowown_build<2, 2>(ctx); break;
    }
    if (rc != NSGPU_OK) return false;
    P = static_cast<RowOwnPlan*>(ctx->rowown_plan);
    // the rows of one entity must fit the shared-memory accumulators of a one-warp CTA, else the cooperative kernel stays
    const int gpw = 32 / ctx->nent;
    if (P && (ro_smem(gpw, ctx->gdim + 1, (P->lmax[0] + 3) & ~3, ctx->nent, ctx->gdim + 1) > 200 * 1024 ||
              ro_smem(gpw, ctx->gdim, (P->lmax[1] + 3) & ~3, ctx->nent, ctx->gdim + 1) > 200 * 1024))
      P->unsupported = true;
  }
  return P && !P->unsupported;
}

int rowown_assemble(nsgpu_ctx* ctx, const double* d_xin, bool want_J, bool want_F, double* d_Fout) {
  switch (ctx->gdim * 10 + ctx->vdeg) {
    case 31: return rowown_run<3, 1>(ctx, d_xin, want_J, want_F, d_Fout);
    case 32: return rowown_run<3, 2>(ctx, d_xin, want_J, want_F, d_Fout);
    case 21: return rowown_run<2, 1>(ctx, d_xin, want_J, want_F, d_Fout);
    case 22: return rowown_run<2, 2>(ctx, d_xin, want_J, want_F, d_Fout);
  }
  set_error(ctx, "rowown: unsupported element");
  return NSGPU_EUNSUPPORTED;
}

}  // namespace nsgpu
