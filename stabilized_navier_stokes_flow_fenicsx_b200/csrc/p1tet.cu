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
//   * Phase B -- one thread per (vertex, neighbour slot, row) walks the precomputed list of parked blocks
//     that target that slot (6.4 on average), sums them and writes the finished 32-byte piece of the CSR
//     row with plain streaming stores: full-sector, write-once.
// No atomics, no colouring, bitwise reproducible; every matrix entry of a locally assembled row is written
// exactly once per assembly (no zero-fill pass needed on a single rank).  Cost: the cell-level part of the
// algebra is recomputed by the four incidences of a cell (~3.7 k instead of ~2.5 k DFMA per cell).
//
// The plan (incidence lists, tiles, gather lists, output positions) is built once per pattern, on the
// device, from the entity-level pair list the pattern builder already sorted.
#include <cub/cub.cuh>

#include "common.cuh"
#include "element_p1tet.cuh"

struct TileHdr {
  int64_t e0, i0, s0;
  int nent, ninc, nslots, pad;
};

struct nsgpu_p1tet_plan {
  int64_t n_inc = 0, n_ent = 0, n_tiles = 0, n_slots = 0;
  int cap = 0, maxdeg = 0;
  uint32_t* d_inc_cell = nullptr;   // [n_inc] cell * 4 + local vertex, sorted by row vertex
  int4* d_inc_vtx = nullptr;        // [n_inc] geometry vertex ids, row vertex first (rotated order)
  int4* d_inc_lead = nullptr;       // [n_inc] first dof of the 4 vertices, same order
  uint8_t* d_src = nullptr;         // [4 n_inc] gather lists: (incidence-within-vertex << 2 | block), grouped by slot
  uint8_t* d_slot_start = nullptr;  // [n_slots + n_ent] per vertex: ns + 1 list offsets inside its 4 * deg items
  int64_t* d_inc_ptr = nullptr;     // [n_ent + 1]
  int64_t* d_slot_ptr = nullptr;    // [n_ent + 1] prefix of neighbour counts
  int64_t* d_rowpos = nullptr;      // [n_ent * 4] CSR start of the vertex's 4 rows
  int32_t* d_rowdof = nullptr;      // [n_ent * 4] the vertex's 4 dofs
  int64_t* d_tile_ent = nullptr;    // [n_tiles + 1]
  TileHdr* d_tile_hdr = nullptr;    // [n_tiles]
  int2* d_ent_rel = nullptr;        // [n_ent] (incidence offset, slot offset) relative to the tile start
  uint8_t* d_cell_bc = nullptr;     // [n_cells] cell touches a Dirichlet dof
  bool contiguous = false;          // every vertex's dofs are (first dof) + 0,1,2,3 and first dof is even
  bool bc_dirty = true;
};

namespace nsgpu {

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

__global__ void k_ent_info(int64_t n_ent, const uint32_t* __restrict__ ent_leader, const int32_t* __restrict__ members,
                           const int64_t* __restrict__ pfirst, const int64_t* __restrict__ plast,
                           const int64_t* __restrict__ indptr, int64_t* nslots, int64_t* rowpos, int32_t* rowdof, int* not_contig) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > n_ent) return;
  if (e == n_ent) { nslots[e] = 0; return; }
  const uint32_t A = ent_leader[e];
  nslots[e] = plast[A] - pfirst[A];
  bool ok = (A & 1u) == 0;
  for (int c = 0; c < 4; ++c) {
    const int32_t d = members[(int64_t)A * KMAX + c];
    rowdof[e * 4 + c] = d;
    rowpos[e * 4 + c] = indptr[d];
    ok = ok && d == (int32_t)A + c;
  }
  if (!ok) *not_contig = 1;
}

// per incidence: rotated vertex / leader quadruples and the four gather-list keys (global slot << 8 | code)
__global__ void k_inc_fill(int64_t n_inc, int64_t n_ent, const uint64_t* __restrict__ keys, const int32_t* __restrict__ cells,
                           const int32_t* __restrict__ dofmap, const uint64_t* __restrict__ pairs, const int64_t* __restrict__ pfirst,
                           const int64_t* __restrict__ plast, const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ slot_ptr,
                           uint32_t* inc_cell, int4* inc_vtx, int4* inc_lead, uint64_t* item_keys, int* slot_cnt, int* bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_inc) return;
  const uint32_t A = (uint32_t)(keys[i] >> 32), cm = (uint32_t)(keys[i] & 0xffffffffu);
  const int64_t cell = cm >> 2;
  const int m = cm & 3;
  inc_cell[i] = cm;
  // entity index: last e with inc_ptr[e] <= i
  int64_t lo = 0, hi = n_ent - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (inc_ptr[mid] <= i) lo = mid; else hi = mid - 1;
  }
  const int64_t e = lo;
  const int64_t rel = i - inc_ptr[e];
  const int64_t lo0 = pfirst[A], hi0 = plast[A];
  int v[4] = {0, 0, 0, 0}, ld[4] = {0, 0, 0, 0};
  int pos = 1;
  for (int k = 0; k < 4; ++k) {
    const int a = (k == m) ? 0 : pos++;                     // rotated position of original local vertex k
    const uint32_t B = (uint32_t)dofmap[cell * 16 + 3 * k];
    v[a] = cells[cell * 4 + k];
    ld[a] = (int)B;
    int64_t l2 = lo0, h2 = hi0 - 1, found = -1;
    while (l2 <= h2) {
      const int64_t mid = (l2 + h2) >> 1;
      const uint32_t pv = (uint32_t)(pairs[mid] & 0xffffffffu);
      if (pv == B) { found = mid - lo0; break; }
      if (pv < B) l2 = mid + 1; else h2 = mid - 1;
    }
    if (found < 0 || rel > 63) { *bad = 1; found = 0; }
    const int64_t gslot = slot_ptr[e] + found;
    item_keys[4 * i + a] = ((uint64_t)gslot << 8) | (uint64_t)(((rel & 63) << 2) | a);
    atomicAdd(slot_cnt + gslot, 1);
  }
  inc_vtx[i] = make_int4(v[0], v[1], v[2], v[3]);
  inc_lead[i] = make_int4(ld[0], ld[1], ld[2], ld[3]);
}

__global__ void k_item_bytes(int64_t n, const uint64_t* __restrict__ keys, uint8_t* src) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n) src[j] = (uint8_t)(keys[j] & 0xffu);
}

__global__ void k_slot_start(int64_t n_ent, const int64_t* __restrict__ slot_ptr, const int* __restrict__ slot_cnt, uint8_t* slot_start, int* bad) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_ent) return;
  const int64_t g0 = slot_ptr[e], g1 = slot_ptr[e + 1];
  int run = 0;
  uint8_t* out = slot_start + g0 + e;
  for (int64_t g = g0; g < g1; ++g) {
    out[g - g0] = (uint8_t)run;
    run += slot_cnt[g];
  }
  out[g1 - g0] = (uint8_t)run;
  if (run > 252) *bad = 1;
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

__global__ void k_tile_hdr(int64_t n_tiles, const int64_t* __restrict__ tile_ent, const int64_t* __restrict__ inc_ptr,
                           const int64_t* __restrict__ slot_ptr, TileHdr* hdr) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  TileHdr h;
  h.e0 = tile_ent[t];
  const int64_t e1 = tile_ent[t + 1];
  h.i0 = inc_ptr[h.e0]; h.s0 = slot_ptr[h.e0];
  h.nent = (int)(e1 - h.e0); h.ninc = (int)(inc_ptr[e1] - h.i0); h.nslots = (int)(slot_ptr[e1] - h.s0); h.pad = 0;
  hdr[t] = h;
}

__global__ void k_ent_rel(int64_t n_tiles, const TileHdr* __restrict__ hdr, const int64_t* __restrict__ inc_ptr,
                          const int64_t* __restrict__ slot_ptr, int2* ent_rel) {
  const int64_t t = blockIdx.x;
  const TileHdr h = hdr[t];
  for (int k = threadIdx.x; k < h.nent; k += blockDim.x)
    ent_rel[h.e0 + k] = make_int2((int)(inc_ptr[h.e0 + k] - h.i0), (int)(slot_ptr[h.e0 + k] - h.s0));
}

__global__ void k_cell_bc(int64_t n_cells, const int32_t* __restrict__ dofmap, const uint8_t* __restrict__ marker, uint8_t* cell_bc) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  uint8_t f = 0;
  for (int k = 0; k < 16; ++k) f |= marker[dofmap[c * 16 + k]];
  cell_bc[c] = f;
}

// ------------------------------------------------------------------------------------------ the kernel
template <int CAP> struct TileSmem {
  // stageJ [4 blocks][CAP][4 rows] double4 (row index swizzled by incidence) | stageF [CAP] double4 |
  // rowpos [CAP][4] i64 | rowdof [CAP][4] i32 | rel [CAP+1] int2 | src [CAP] u32 | slot_start [5 CAP + 8] u8 |
  // ent_of_slot [4 CAP] u8
  static constexpr size_t stageJ = 16 * (size_t)CAP * sizeof(double4);
  static constexpr size_t stageF = (size_t)CAP * sizeof(double4);
  static constexpr size_t tables = 32 * CAP + 16 * CAP + 8 * (CAP + 2) + 4 * CAP + (5 * CAP + 16) + 4 * CAP;
  static constexpr size_t bytes(bool want_J) { return (want_J ? stageJ : 0) + stageF + tables; }
};

template <int CAP, int MINB, bool WANT_J, bool WANT_F>
__global__ void __launch_bounds__(CAP, MINB)
k_p1tet_tiles(FormParams form, const double* __restrict__ xg, const double* __restrict__ wv, const int32_t* __restrict__ members,
              const bool contiguous, const uint8_t* __restrict__ bc_marker, const double* __restrict__ bc_value,
              const uint8_t* __restrict__ cell_bc, const uint32_t* __restrict__ inc_cell, const int4* __restrict__ inc_vtx,
              const int4* __restrict__ inc_lead, const uint32_t* __restrict__ src, const uint8_t* __restrict__ slot_start,
              const int2* __restrict__ ent_rel, const int64_t* __restrict__ rowpos, const int4* __restrict__ rowdof,
              const TileHdr* __restrict__ tile_hdr, double* __restrict__ vals, double* __restrict__ F) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double4* stageJ = reinterpret_cast<double4*>(smem_raw);
  double4* stageF = stageJ + (WANT_J ? 16 * CAP : 0);
  int64_t* s_rowpos = reinterpret_cast<int64_t*>(stageF + CAP);
  int4* s_rowdof = reinterpret_cast<int4*>(s_rowpos + 4 * CAP);
  int2* s_rel = reinterpret_cast<int2*>(s_rowdof + CAP);
  uint32_t* s_src = reinterpret_cast<uint32_t*>(s_rel + (CAP + 2));
  uint8_t* s_ss = reinterpret_cast<uint8_t*>(s_src + CAP);
  uint8_t* ent_of_slot = s_ss + (5 * CAP + 16);

  const int tid = threadIdx.x;
  const TileHdr h = tile_hdr[blockIdx.x];
  if (h.nent <= 0) return;

  // ---- issue every table load of the tile up front: their latency hides behind phase A ----
  const bool has_inc = tid < h.ninc, has_ent = tid < h.nent;
  uint32_t r_src = 0;
  int2 r_rel = make_int2(h.ninc, h.nslots), r_rel_next = make_int2(h.ninc, h.nslots);
  int64_t r_rowpos[4] = {0, 0, 0, 0};
  int4 r_rowdof = make_int4(0, 0, 0, 0);
  uint8_t r_ss[5] = {0, 0, 0, 0, 0};
  const int n_ss = h.nslots + h.nent + 1;
  if (WANT_J) {
    if (has_inc) r_src = src[h.i0 + tid];
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (tid + k * CAP < n_ss) r_ss[k] = slot_start[h.s0 + h.e0 + tid + k * CAP];
  }
  if (has_ent) {
    r_rel = ent_rel[h.e0 + tid];
    if (tid + 1 < h.nent) r_rel_next = ent_rel[h.e0 + tid + 1];
    r_rowdof = rowdof[h.e0 + tid];
    if (WANT_J) {
      const longlong2* rp = reinterpret_cast<const longlong2*>(rowpos + 4 * (h.e0 + tid));
      const longlong2 p01 = rp[0], p23 = rp[1];
      r_rowpos[0] = p01.x; r_rowpos[1] = p01.y; r_rowpos[2] = p23.x; r_rowpos[3] = p23.y;
    }
  }

  // ---------------- phase A: one incidence per thread ----------------
  if (has_inc) {
    const int4 vt = inc_vtx[h.i0 + tid];
    const int4 ld = inc_lead[h.i0 + tid];
    const uint32_t cm = inc_cell[h.i0 + tid];
    const int vtx[4] = {vt.x, vt.y, vt.z, vt.w};
    const int lead[4] = {ld.x, ld.y, ld.z, ld.w};
    double x[4][3], u[4][3], p[4];
    int dof[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double* xp = xg + 3 * (int64_t)vtx[a];
      x[a][0] = xp[0]; x[a][1] = xp[1]; x[a][2] = xp[2];
      if (contiguous) {
        const double2* wp = reinterpret_cast<const double2*>(wv + lead[a]);
        const double2 w01 = wp[0], w23 = wp[1];
        u[a][0] = w01.x; u[a][1] = w01.y; u[a][2] = w23.x; p[a] = w23.y;
#pragma unroll
        for (int c = 0; c < 4; ++c) dof[a][c] = lead[a] + c;
      } else {
        const int4 mem = reinterpret_cast<const int4*>(members)[lead[a]];
        dof[a][0] = mem.x; dof[a][1] = mem.y; dof[a][2] = mem.z; dof[a][3] = mem.w;
        u[a][0] = wv[mem.x]; u[a][1] = wv[mem.y]; u[a][2] = wv[mem.z]; p[a] = wv[mem.w];
      }
    }
    double blk[4][16], fr[4];
    const bool row_is_origin = (cm & 3u) == 0;
    const bool has_bc = cell_bc && cell_bc[cm >> 2];
    if (WANT_J || !has_bc) p1tet_rowslab<WANT_J, WANT_F>(form, row_is_origin, x, u, p, blk, fr);
    else p1tet_rowslab<true, WANT_F>(form, row_is_origin, x, u, p, blk, fr);   // residual only, but lifting needs the Jacobian rows

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
      const int sw = tid & 3;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          stageJ[(a * CAP + tid) * 4 + (r ^ sw)] = make_double4(blk[a][4 * r], blk[a][4 * r + 1], blk[a][4 * r + 2], blk[a][4 * r + 3]);
    }
    if (WANT_F) stageF[tid] = make_double4(fr[0], fr[1], fr[2], fr[3]);
  }
  // park the tables
  if (WANT_J) {
    if (has_inc) s_src[tid] = r_src;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (tid + k * CAP < n_ss) s_ss[tid + k * CAP] = r_ss[k];
  }
  if (has_ent) {
    s_rel[tid] = r_rel;
    s_rowdof[tid] = r_rowdof;
    if (WANT_J) {
#pragma unroll
      for (int r = 0; r < 4; ++r) s_rowpos[4 * tid + r] = r_rowpos[r];
      for (int k = r_rel.y; k < r_rel_next.y; ++k) ent_of_slot[k] = (uint8_t)tid;
    }
  }
  if (tid == 0) s_rel[h.nent] = make_int2(h.ninc, h.nslots);
  __syncthreads();

  // ---------------- phase B ----------------
  if (WANT_F) {
    // residual: (vertex, row) sums over the vertex's incidences, four lanes per sum
    const int nF = 16 * h.nent;
    const double* sf = reinterpret_cast<const double*>(stageF);
    for (int base = 0; base < nF; base += CAP) {
      const int item = base + tid;
      const int part = item & 3, r = (item >> 2) & 3, le = item >> 4;
      double acc = 0.0;
      if (le < h.nent)
        for (int ii = s_rel[le].x + part; ii < s_rel[le + 1].x; ii += 4) acc += sf[4 * ii + r];
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (le < h.nent && part == 0) {
        const int4 rd = s_rowdof[le];
        F[(r == 0) ? rd.x : (r == 1) ? rd.y : (r == 2) ? rd.z : rd.w] = acc;
      }
    }
  }
  if (WANT_J) {
    // Jacobian: one (vertex, slot, row) piece per thread
    const int nitems = 4 * h.nslots;
    const uint8_t* srcb = reinterpret_cast<const uint8_t*>(s_src);
    for (int item = tid; item < nitems; item += CAP) {
      const int ls = item >> 2, r = item & 3;
      const int le = ent_of_slot[ls];
      const int2 rel = s_rel[le];
      const int s = ls - rel.y;
      const uint8_t* ss = s_ss + ls + le;
      const int jb = ss[0], je = ss[1];
      const uint8_t* sp = srcb + 4 * rel.x;
      double4 acc = make_double4(0.0, 0.0, 0.0, 0.0);
      for (int j = jb; j < je; ++j) {
        const int code = sp[j];
        const int ii = rel.x + (code >> 2), a = code & 3;
        const double4 v = stageJ[(a * CAP + ii) * 4 + (r ^ (ii & 3))];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      double* dst = vals + s_rowpos[4 * le + r] + 4 * s;
      __stcs(reinterpret_cast<double2*>(dst), make_double2(acc.x, acc.y));
      __stcs(reinterpret_cast<double2*>(dst) + 1, make_double2(acc.z, acc.w));
    }
  }
}

// ------------------------------------------------------------------------------------------ quad-lane kernel
// Same tile / staging / phase-B machinery; phase A runs FOUR lanes per incidence (p1tet_quad), so a tile of CAPI
// incidences is a CTA of 4*CAPI threads with ~1/2 the registers per thread.
template <int CAPI, int MINB, bool WANT_J, bool WANT_F>
__global__ void __launch_bounds__(4 * CAPI, MINB)
k_p1tet_quad(FormParams form, const double* __restrict__ xg, const double* __restrict__ wv, const int32_t* __restrict__ members,
             const bool contiguous, const uint8_t* __restrict__ bc_marker, const double* __restrict__ bc_value,
             const uint8_t* __restrict__ cell_bc, const uint32_t* __restrict__ inc_cell, const int4* __restrict__ inc_vtx,
             const int4* __restrict__ inc_lead, const uint32_t* __restrict__ src, const uint8_t* __restrict__ slot_start,
             const int2* __restrict__ ent_rel, const int64_t* __restrict__ rowpos, const int4* __restrict__ rowdof,
             const TileHdr* __restrict__ tile_hdr, double* __restrict__ vals, double* __restrict__ F) {
  constexpr int CAP = CAPI;
  constexpr int NT = 4 * CAPI;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double4* stageJ = reinterpret_cast<double4*>(smem_raw);
  double4* stageF = stageJ + (WANT_J ? 16 * CAP : 0);
  int64_t* s_rowpos = reinterpret_cast<int64_t*>(stageF + CAP);
  int4* s_rowdof = reinterpret_cast<int4*>(s_rowpos + 4 * CAP);
  int2* s_rel = reinterpret_cast<int2*>(s_rowdof + CAP);
  uint32_t* s_src = reinterpret_cast<uint32_t*>(s_rel + (CAP + 2));
  uint8_t* s_ss = reinterpret_cast<uint8_t*>(s_src + CAP);
  uint8_t* ent_of_slot = s_ss + (5 * CAP + 16);

  const int tid = threadIdx.x;
  const TileHdr h = tile_hdr[blockIdx.x];
  if (h.nent <= 0) return;
  const int inc = tid >> 2, j = tid & 3;
  const bool has_inc = inc < h.ninc;

  // ---- tile tables: every thread moves a few entries global -> shared (latency overlaps phase A) ----
  const int n_ss = h.nslots + h.nent + 1;
  if (WANT_J) {
    if (tid < h.ninc) s_src[tid] = src[h.i0 + tid];
    for (int k = tid; k < n_ss; k += NT) s_ss[k] = slot_start[h.s0 + h.e0 + k];
  }
  if (tid < h.nent) {
    const int2 rel = ent_rel[h.e0 + tid];
    const int2 reln = (tid + 1 < h.nent) ? ent_rel[h.e0 + tid + 1] : make_int2(h.ninc, h.nslots);
    s_rel[tid] = rel;
    s_rowdof[tid] = rowdof[h.e0 + tid];
    if (WANT_J) {
      const longlong2* rp = reinterpret_cast<const longlong2*>(rowpos + 4 * (h.e0 + tid));
      const longlong2 p01 = rp[0], p23 = rp[1];
      s_rowpos[4 * tid] = p01.x; s_rowpos[4 * tid + 1] = p01.y; s_rowpos[4 * tid + 2] = p23.x; s_rowpos[4 * tid + 3] = p23.y;
      for (int k = rel.y; k < reln.y; ++k) ent_of_slot[k] = (uint8_t)tid;
    }
  }
  if (tid == 0) s_rel[h.nent] = make_int2(h.ninc, h.nslots);

  // ---------------- phase A: four lanes per incidence ----------------
  {
    // lanes of idle quads (inc >= ninc) replay the tile's first incidence so that the quad shuffles stay warp-uniform
    const int64_t gi = h.i0 + (has_inc ? inc : 0);
    const int4 vt = inc_vtx[gi];
    const int4 ld = inc_lead[gi];
    const uint32_t cm = inc_cell[gi];
    const int vtx[4] = {vt.x, vt.y, vt.z, vt.w};
    const int lead[4] = {ld.x, ld.y, ld.z, ld.w};
    double x[4][3], u[4][3], p[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double* xp = xg + 3 * (int64_t)vtx[a];
      x[a][0] = xp[0]; x[a][1] = xp[1]; x[a][2] = xp[2];
      if (contiguous) {
        const double2* wp = reinterpret_cast<const double2*>(wv + lead[a]);
        const double2 w01 = wp[0], w23 = wp[1];
        u[a][0] = w01.x; u[a][1] = w01.y; u[a][2] = w23.x; p[a] = w23.y;
      } else {
        const int4 mem = reinterpret_cast<const int4*>(members)[lead[a]];
        u[a][0] = wv[mem.x]; u[a][1] = wv[mem.y]; u[a][2] = wv[mem.z]; p[a] = wv[mem.w];
      }
    }
    double blk[16], fr[4];
    const bool row_is_origin = (cm & 3u) == 0;
    const bool has_bc = cell_bc && cell_bc[cm >> 2];
    // lifting needs the Jacobian rows even in a residual-only pass; BC cells are rare, so the branch is cheap
    if (WANT_J) p1tet_quad<true, WANT_F>(form, row_is_origin, j, x, u, p, blk, fr);
    else if (__any_sync(0xffffffffu, has_bc)) p1tet_quad<true, WANT_F>(form, row_is_origin, j, x, u, p, blk, fr);
    else p1tet_quad<false, WANT_F>(form, row_is_origin, j, x, u, p, blk, fr);

    double lift[4] = {0.0, 0.0, 0.0, 0.0};
    if (has_bc) {
      // Dirichlet handling at element level (assemble_matrix / apply_lifting semantics)
      const int lj = (j == 0) ? lead[0] : (j == 1) ? lead[1] : (j == 2) ? lead[2] : lead[3];
      int cd[4], rd[4];
      if (contiguous) {
#pragma unroll
        for (int d = 0; d < 4; ++d) { cd[d] = lj + d; rd[d] = lead[0] + d; }
      } else {
        const int4 mc = reinterpret_cast<const int4*>(members)[lj];
        const int4 mr = reinterpret_cast<const int4*>(members)[lead[0]];
        cd[0] = mc.x; cd[1] = mc.y; cd[2] = mc.z; cd[3] = mc.w;
        rd[0] = mr.x; rd[1] = mr.y; rd[2] = mr.z; rd[3] = mr.w;
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (bc_marker[cd[d]]) {
          const double delta = bc_value[cd[d]] - wv[cd[d]];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (WANT_F) lift[r] += blk[4 * r + d] * delta;   // lifting with the un-zeroed entry
            blk[4 * r + d] = 0.0;                              // constrained trial column
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (bc_marker[rd[r]]) {
#pragma unroll
          for (int d = 0; d < 4; ++d) blk[4 * r + d] = 0.0;    // constrained test row
        }
    }
    if (WANT_F) {
#pragma unroll
      for (int r = 0; r < 4; ++r) fr[r] += quad_sum(lift[r]);
    }
    if (has_inc) {
      if (WANT_J) {
        const int sw = inc & 3;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          stageJ[(j * CAP + inc) * 4 + (r ^ sw)] = make_double4(blk[4 * r], blk[4 * r + 1], blk[4 * r + 2], blk[4 * r + 3]);
      }
      if (WANT_F && j == 0) stageF[inc] = make_double4(fr[0], fr[1], fr[2], fr[3]);
    }
  }
  __syncthreads();

  // ---------------- phase B ----------------
  if (WANT_F) {
    const int nF = 16 * h.nent;
    const double* sf = reinterpret_cast<const double*>(stageF);
    for (int base = 0; base < nF; base += NT) {
      const int item = base + tid;
      const int part = item & 3, r = (item >> 2) & 3, le = item >> 4;
      double acc = 0.0;
      if (le < h.nent)
        for (int ii = s_rel[le].x + part; ii < s_rel[le + 1].x; ii += 4) acc += sf[4 * ii + r];
      acc = quad_sum(acc);
      if (le < h.nent && part == 0) {
        const int4 rd = s_rowdof[le];
        F[(r == 0) ? rd.x : (r == 1) ? rd.y : (r == 2) ? rd.z : rd.w] = acc;
      }
    }
  }
  if (WANT_J) {
    const int nitems = 4 * h.nslots;
    const uint8_t* srcb = reinterpret_cast<const uint8_t*>(s_src);
    for (int item = tid; item < nitems; item += NT) {
      const int ls = item >> 2, r = item & 3;
      const int le = ent_of_slot[ls];
      const int2 rel = s_rel[le];
      const int s = ls - rel.y;
      const uint8_t* ss = s_ss + ls + le;
      const int jb = ss[0], je = ss[1];
      const uint8_t* sp = srcb + 4 * rel.x;
      double4 acc = make_double4(0.0, 0.0, 0.0, 0.0);
      for (int q = jb; q < je; ++q) {
        const int code = sp[q];
        const int ii = rel.x + (code >> 2), a = code & 3;
        const double4 v = stageJ[(a * CAP + ii) * 4 + (r ^ (ii & 3))];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      double* dst = vals + s_rowpos[4 * le + r] + 4 * s;
      __stcs(reinterpret_cast<double2*>(dst), make_double2(acc.x, acc.y));
      __stcs(reinterpret_cast<double2*>(dst) + 1, make_double2(acc.z, acc.w));
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
void p1tet_free(nsgpu_ctx* ctx) {
  nsgpu_p1tet_plan* P = ctx->p1plan;
  if (!P) return;
  cudaFree(P->d_inc_cell); cudaFree(P->d_inc_vtx); cudaFree(P->d_inc_lead); cudaFree(P->d_src); cudaFree(P->d_slot_start);
  cudaFree(P->d_inc_ptr); cudaFree(P->d_slot_ptr); cudaFree(P->d_rowpos); cudaFree(P->d_rowdof); cudaFree(P->d_tile_ent);
  cudaFree(P->d_cell_bc); cudaFree(P->d_tile_hdr); cudaFree(P->d_ent_rel);
  delete P;
  ctx->p1plan = nullptr;
}

void p1tet_mark_bc_dirty(nsgpu_ctx* ctx) {
  if (ctx->p1plan) ctx->p1plan->bc_dirty = true;
}

static bool use_quad(nsgpu_ctx* ctx) { return ctx->lanes == 4; }
static int plan_cap(nsgpu_ctx* ctx) { return use_quad(ctx) ? ctx->threads / 4 : ctx->threads; }   // incidences per tile
// launch-bounds pairing: 256 -> 1 CTA/SM, 192 -> 1, 128 -> 2 (all at the full 255-register budget)

bool p1tet_fast_available(nsgpu_ctx* ctx) {
  // valid (lanes, threads) pairs: 1 x {64,128,192,256}; 4 x {256,384,512}
  if (ctx->lanes == 4 && ctx->threads < 256) ctx->threads = 256;
  if (ctx->lanes == 1 && ctx->threads > 256) ctx->threads = 256;
  if (ctx->gdim != 3 || ctx->vdeg != 1 || !ctx->pattern_built || !ctx->rows_presorted || !ctx->d_pairs) return false;
  if (ctx->n_cells_owned >= ((int64_t)1 << 29)) return false;
  if (ctx->p1plan && ctx->p1plan->cap != plan_cap(ctx)) p1tet_free(ctx);
  if (!ctx->p1plan) {
    if (p1tet_build_plan(ctx) != NSGPU_OK) return false;
  }
  return ctx->p1plan != nullptr;
}

static inline unsigned g256(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, 256); }

template <int CAP, int MINB>
static cudaError_t set_smem_attr() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_p1tet_tiles<CAP, MINB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<CAP>::bytes(true)))) return e;
  if ((e = cudaFuncSetAttribute(k_p1tet_tiles<CAP, MINB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<CAP>::bytes(true)))) return e;
  return cudaFuncSetAttribute(k_p1tet_tiles<CAP, MINB, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<CAP>::bytes(false));
}

template <int CAPI, int MINB>
static cudaError_t set_smem_attr_quad() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_p1tet_quad<CAPI, MINB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<CAPI>::bytes(true)))) return e;
  if ((e = cudaFuncSetAttribute(k_p1tet_quad<CAPI, MINB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<CAPI>::bytes(true)))) return e;
  return cudaFuncSetAttribute(k_p1tet_quad<CAPI, MINB, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<CAPI>::bytes(false));
}

int p1tet_build_plan(nsgpu_ctx* ctx) {
  cudaStream_t s = ctx->stream;
  p1tet_free(ctx);
  nsgpu_p1tet_plan* P = new nsgpu_p1tet_plan();
  const int64_t n_inc = ctx->n_cells_owned * 4;
  const int CAPV = plan_cap(ctx);
  P->n_inc = n_inc;
  P->cap = CAPV;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr, *d_items = nullptr, *d_items2 = nullptr;
  uint32_t* d_leader = nullptr;
  int64_t *d_cnt = nullptr, *d_nrun = nullptr, *d_nslots = nullptr, *d_max = nullptr;
  int* d_slot_cnt = nullptr;
  void* d_tmp = nullptr;
  int* d_flag = nullptr;   // [0] bad, [1] not contiguous
  auto cleanup = [&]() {
    cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_items); cudaFree(d_items2); cudaFree(d_leader); cudaFree(d_cnt);
    cudaFree(d_nrun); cudaFree(d_nslots); cudaFree(d_max); cudaFree(d_slot_cnt); cudaFree(d_tmp); cudaFree(d_flag);
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
  PL_CUDA(cudaStreamSynchronize(s));
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
  if (maxdeg > 63 || maxdeg > CAPV / 2) {   // pathological vertex degree: keep the generic path
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

  // per-entity output info and neighbour-slot prefix
  PL_CUDA(cudaMalloc(&d_flag, 2 * sizeof(int)));
  PL_CUDA(cudaMemsetAsync(d_flag, 0, 2 * sizeof(int), s));
  PL_CUDA(cudaMalloc(&d_nslots, sizeof(int64_t) * (n_ent + 1)));
  PL_CUDA(cudaMalloc(&P->d_slot_ptr, sizeof(int64_t) * (n_ent + 1)));
  PL_CUDA(cudaMalloc(&P->d_rowpos, sizeof(int64_t) * n_ent * 4));
  PL_CUDA(cudaMalloc(&P->d_rowdof, sizeof(int32_t) * n_ent * 4));
  k_ent_info<<<g256(n_ent + 1), 256, 0, s>>>(n_ent, d_leader, ctx->d_members, ctx->d_pair_first, ctx->d_pair_last, ctx->d_indptr,
                                             d_nslots, P->d_rowpos, P->d_rowdof, d_flag + 1);
  tb = 0;
  PL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_nslots, P->d_slot_ptr, n_ent + 1, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_nslots, P->d_slot_ptr, n_ent + 1, s));
  int64_t n_slots = 0;
  PL_CUDA(cudaMemcpyAsync(&n_slots, P->d_slot_ptr + n_ent, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  PL_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_tmp); d_tmp = nullptr;
  P->n_slots = n_slots;

  // per-incidence words + gather-list items
  PL_CUDA(cudaMalloc(&P->d_inc_cell, sizeof(uint32_t) * n_inc));
  PL_CUDA(cudaMalloc(&P->d_inc_vtx, sizeof(int4) * n_inc));
  PL_CUDA(cudaMalloc(&P->d_inc_lead, sizeof(int4) * n_inc));
  PL_CUDA(cudaMalloc(&d_items, sizeof(uint64_t) * 4 * n_inc));
  PL_CUDA(cudaMalloc(&d_slot_cnt, sizeof(int) * (n_slots + 1)));
  PL_CUDA(cudaMemsetAsync(d_slot_cnt, 0, sizeof(int) * (n_slots + 1), s));
  k_inc_fill<<<g256(n_inc), 256, 0, s>>>(n_inc, n_ent, d_keys2, ctx->d_cells, ctx->d_dofmap, ctx->d_pairs, ctx->d_pair_first,
                                         ctx->d_pair_last, P->d_inc_ptr, P->d_slot_ptr, P->d_inc_cell, P->d_inc_vtx, P->d_inc_lead,
                                         d_items, d_slot_cnt, d_flag);
  PL_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_keys2); d_keys2 = nullptr;
  // sort items by (global slot, code): the low bytes in sorted order are the gather lists
  int slot_bits = 1;
  while (slot_bits < 56 && ((uint64_t)n_slots >> slot_bits)) ++slot_bits;
  PL_CUDA(cudaMalloc(&d_items2, sizeof(uint64_t) * 4 * n_inc));
  tb = 0;
  PL_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_items, d_items2, 4 * n_inc, 0, 8 + slot_bits, s));
  PL_CUDA(cudaMalloc(&d_tmp, tb));
  PL_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tb, d_items, d_items2, 4 * n_inc, 0, 8 + slot_bits, s));
  PL_CUDA(cudaMalloc(&P->d_src, 4 * n_inc));
  k_item_bytes<<<g256(4 * n_inc), 256, 0, s>>>(4 * n_inc, d_items2, P->d_src);
  PL_CUDA(cudaMalloc(&P->d_slot_start, n_slots + n_ent + 1));
  k_slot_start<<<g256(n_ent), 256, 0, s>>>(n_ent, P->d_slot_ptr, d_slot_cnt, P->d_slot_start, d_flag);

  // tiles: entity e belongs to tile floor(inc_ptr[e] / capeff)
  const int64_t capeff = CAPV - maxdeg + 1;
  P->n_tiles = ceil_div(n_inc, capeff);
  PL_CUDA(cudaMalloc(&P->d_tile_ent, sizeof(int64_t) * (P->n_tiles + 1)));
  k_tiles<<<g256(P->n_tiles + 1), 256, 0, s>>>(P->n_tiles, n_ent, capeff, P->d_inc_ptr, P->d_tile_ent);
  PL_CUDA(cudaMalloc(&P->d_tile_hdr, sizeof(TileHdr) * P->n_tiles));
  PL_CUDA(cudaMalloc(&P->d_ent_rel, sizeof(int2) * (n_ent + 1)));
  k_tile_hdr<<<g256(P->n_tiles), 256, 0, s>>>(P->n_tiles, P->d_tile_ent, P->d_inc_ptr, P->d_slot_ptr, P->d_tile_hdr);
  k_ent_rel<<<(unsigned)P->n_tiles, 64, 0, s>>>(P->n_tiles, P->d_tile_hdr, P->d_inc_ptr, P->d_slot_ptr, P->d_ent_rel);
  PL_CUDA(cudaMalloc(&P->d_cell_bc, ctx->n_cells_owned > 0 ? ctx->n_cells_owned : 1));
  int flags[2] = {0, 0};
  PL_CUDA(cudaMemcpyAsync(flags, d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  PL_CUDA(cudaStreamSynchronize(s));
  PL_CUDA(cudaGetLastError());
  ctx->launches += 16;
  cleanup();
  if (flags[0]) {   // a vertex with too many neighbours / incidences for the byte-packed lists: generic path
    ctx->p1plan = P; p1tet_free(ctx);
    return NSGPU_EUNSUPPORTED;
  }
  P->contiguous = flags[1] == 0;
  P->bc_dirty = true;
  ctx->p1plan = P;
  cudaError_t e = cudaSuccess;
  if (use_quad(ctx)) e = CAPV == 128 ? set_smem_attr_quad<128, 1>() : (CAPV == 96 ? set_smem_attr_quad<96, 1>() : set_smem_attr_quad<64, 2>());
  else e = CAPV == 256 ? set_smem_attr<256, 1>() : (CAPV == 192 ? set_smem_attr<192, 1>() : (CAPV == 128 ? set_smem_attr<128, 2>() : set_smem_attr<64, 4>()));
  if (e != cudaSuccess) { set_error(ctx, std::string("p1tet plan: smem attribute: ") + cudaGetErrorString(e)); p1tet_free(ctx); return NSGPU_ECUDA; }
  return NSGPU_OK;
#undef PL_CUDA
}

template <int CAP, int MINB>
static void launch_tiles(nsgpu_ctx* ctx, nsgpu_p1tet_plan* P, const double* d_xin, bool want_J, bool want_F, double* d_Fout, const uint8_t* cbc) {
  cudaStream_t s = ctx->stream;
#define P1_LAUNCH(J, F)                                                                                              \
  k_p1tet_tiles<CAP, MINB, J, F><<<(unsigned)P->n_tiles, CAP, TileSmem<CAP>::bytes(J), s>>>(ctx->form, ctx->d_x, d_xin,   \
      ctx->d_members, P->contiguous, ctx->d_bc_marker, ctx->d_bc_value, cbc, P->d_inc_cell, P->d_inc_vtx, P->d_inc_lead,  \
      reinterpret_cast<const uint32_t*>(P->d_src), P->d_slot_start, P->d_ent_rel, P->d_rowpos,                            \
      reinterpret_cast<const int4*>(P->d_rowdof), P->d_tile_hdr, ctx->d_vals, d_Fout)
  if (want_J && want_F) P1_LAUNCH(true, true);
  else if (want_J) P1_LAUNCH(true, false);
  else P1_LAUNCH(false, true);
#undef P1_LAUNCH
}

template <int CAPI, int MINB>
static void launch_quad(nsgpu_ctx* ctx, nsgpu_p1tet_plan* P, const double* d_xin, bool want_J, bool want_F, double* d_Fout, const uint8_t* cbc) {
  cudaStream_t s = ctx->stream;
#define P1Q_LAUNCH(J, F)                                                                                             \
  k_p1tet_quad<CAPI, MINB, J, F><<<(unsigned)P->n_tiles, 4 * CAPI, TileSmem<CAPI>::bytes(J), s>>>(ctx->form, ctx->d_x, d_xin,   \
      ctx->d_members, P->contiguous, ctx->d_bc_marker, ctx->d_bc_value, cbc, P->d_inc_cell, P->d_inc_vtx, P->d_inc_lead,  \
      reinterpret_cast<const uint32_t*>(P->d_src), P->d_slot_start, P->d_ent_rel, P->d_rowpos,                            \
      reinterpret_cast<const int4*>(P->d_rowdof), P->d_tile_hdr, ctx->d_vals, d_Fout)
  if (want_J && want_F) P1Q_LAUNCH(true, true);
  else if (want_J) P1Q_LAUNCH(true, false);
  else P1Q_LAUNCH(false, true);
#undef P1Q_LAUNCH
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
  if (use_quad(ctx)) {
    if (P->cap == 128) launch_quad<128, 1>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
    else if (P->cap == 96) launch_quad<96, 1>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
    else launch_quad<64, 2>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
  } else if (P->cap == 256) launch_tiles<256, 1>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
  else if (P->cap == 192) launch_tiles<192, 1>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
  else if (P->cap == 128) launch_tiles<128, 2>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
  else launch_tiles<64, 4>(ctx, P, d_xin, want_J, want_F, d_Fout, cbc);
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

}  // namespace nsgpu
