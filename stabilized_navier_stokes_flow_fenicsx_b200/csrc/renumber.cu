// renumber.cu -- internal vertex-blocked dof numbering behind the C ABI.
//
// The reference hands over W.dofmap.list of functionspace(msh, mixed_element([P1^3, P1]))
// (NavierStokes/NavierStokesChannelFlow.py:128-129): block size 1, the four dofs of a mesh vertex carry whatever
// numbers dolfinx's graph reordering gave them.  The factorised row-owner kernels and the 4x4-block SpMV want the
// four dofs of a vertex at 4 e + (u_x, u_y, u_z, p) with spatially compact runs of vertices.  So the library keeps
// its OWN numbering below the ABI:
//   * entity (vertex) order = owned vertices sorted by the Morton code of their coordinates, then ghost vertices;
//   * internal dof = 4 * (entity rank) + component; column ghosts (>= n_dofs) keep their index;
//   * everything below nsgpu.cu (pattern, plans, kernels, halo lists, Krylov vectors) only ever sees internal indices;
//   * the ABI translates: x / b in, F / y / x out (one gather pass each), Dirichlet / halo / pattern-entry index lists
//     at set-up, and the CSR pattern + values are exported in the CALLER's numbering (sorted local columns, i.e. exactly
//     dolfinx's la::SparsityPattern order) on request.
// A caller numbering that already is vertex-blocked (the synthetic ducts) is used as it is: no permutation, no passes.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"
#include "element_p1tet.cuh"

namespace nsgpu {

static inline unsigned g256(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, 256); }

// ---------------------------------------------------------------------------------------------- numbering
// entity tables of the P1-P1 tetrahedron space from every local cell: leader (= u_x dof) of each dof, the four member dofs
// and the geometry vertex of each leader
__global__ void k_rn_entities(int64_t n_cells, const int32_t* __restrict__ dofmap, const int32_t* __restrict__ cells, int32_t* leader,
                              int32_t* members, int32_t* evert) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_cells * 4) return;
  const int64_t cell = t >> 2;
  const int m = (int)(t & 3);
  const int32_t* dm = dofmap + cell * 16;
  const int32_t A = dm[3 * m];
  const int32_t d[4] = {A, dm[3 * m + 1], dm[3 * m + 2], dm[12 + m]};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    leader[d[c]] = A;
    members[(int64_t)A * 4 + c] = d[c];
  }
  evert[A] = cells[cell * 4 + m];
}

// flags[0]: a dof belongs to no cell; [1]: an entity mixes owned and ghost dofs; [2]: the numbering is not vertex-blocked
__global__ void k_rn_check(int64_t n_dofs, int64_t n_owned, const int32_t* __restrict__ leader, const int32_t* __restrict__ members, int* flags,
                           uint8_t* is_leader) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= n_dofs) return;
  const int32_t A = leader[d];
  is_leader[d] = (A == (int32_t)d) ? 1 : 0;
  if (A < 0) { flags[0] = 1; return; }
  if ((d < n_owned) != ((int64_t)A < n_owned)) flags[1] = 1;
  if (A == (int32_t)d) {
    const bool ghost = d >= n_owned;
    const int64_t base = ghost ? n_owned : 0;
    bool ok = ((d - base) & 3) == 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) ok = ok && members[d * 4 + c] == (int32_t)d + c;
    if (!ok) flags[2] = 1;
  }
}

__device__ __forceinline__ uint64_t spread21(uint64_t v) {   // bits of v (21) to every third position
  v &= 0x1fffffULL;
  v = (v | (v << 32)) & 0x1f00000000ffffULL;
  v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
  v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
  v = (v | (v << 2)) & 0x1249249249249249ULL;
  return v;
}

// sort key of an entity: ghost flag on top, then the 63-bit Morton code of the vertex (mode 2) or the leader dof (mode 1)
__global__ void k_rn_keys(int64_t n_ent, int64_t n_owned, const int32_t* __restrict__ ent_leader, const int32_t* __restrict__ evert,
                          const double* __restrict__ xg, double lo0, double lo1, double lo2, double inv_scale, int morton, uint64_t* keys) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_ent) return;
  const int32_t A = ent_leader[e];
  uint64_t k;
  if (morton) {
    const double* p = xg + 3 * (int64_t)evert[A];
    const double s = 2097151.0;   // 2^21 - 1
    const uint64_t q0 = (uint64_t)fmin(fmax((p[0] - lo0) * inv_scale, 0.0) * s, s);
    const uint64_t q1 = (uint64_t)fmin(fmax((p[1] - lo1) * inv_scale, 0.0) * s, s);
    const uint64_t q2 = (uint64_t)fmin(fmax((p[2] - lo2) * inv_scale, 0.0) * s, s);
    k = spread21(q0) | (spread21(q1) << 1) | (spread21(q2) << 2);
  } else {
    k = (uint64_t)(uint32_t)A;
  }
  keys[e] = k | ((int64_t)A >= n_owned ? (1ULL << 63) : 0ULL);
}

__global__ void k_rn_assign(int64_t n_ent, int64_t n_owned, int64_t n_own_ent, const int32_t* __restrict__ order, const int32_t* __restrict__ members,
                            int32_t* perm, int32_t* iperm) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_ent) return;
  const int32_t A = order[e];
  const int64_t base = e < n_own_ent ? 4 * e : n_owned + 4 * (e - n_own_ent);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int32_t d = members[(int64_t)A * 4 + c];
    perm[d] = (int32_t)(base + c);
    iperm[base + c] = d;
  }
}

__global__ void k_rn_map_inplace(int64_t n, const int32_t* __restrict__ perm, int32_t* a) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = perm[a[i]];
}

void renumber_free(nsgpu_ctx* ctx) {
  cudaFree(ctx->d_perm); cudaFree(ctx->d_iperm); cudaFree(ctx->d_indptr_c); cudaFree(ctx->d_indices_c); cudaFree(ctx->d_vals_c);
  cudaFree(ctx->d_px); cudaFree(ctx->d_pF);
  ctx->d_perm = ctx->d_iperm = nullptr; ctx->d_indptr_c = nullptr; ctx->d_indices_c = nullptr; ctx->d_vals_c = nullptr;
  ctx->d_px = ctx->d_pF = nullptr;
  ctx->h_perm.clear();
  ctx->caller_pattern_built = false;
}

// called at the end of nsgpu_set_space with the caller's dofmap in ctx->d_dofmap; may rewrite it in internal numbering
int renumber_build(nsgpu_ctx* ctx) {
  renumber_free(ctx);
  if (ctx->renumber == 0 || ctx->gdim != 3 || ctx->vdeg != 1 || ctx->n_cells_total == 0) return NSGPU_OK;
  cudaStream_t s = ctx->stream;
  const int64_t n = ctx->n_dofs, n_owned = ctx->n_owned;
  int32_t *d_leader = nullptr, *d_members = nullptr, *d_evert = nullptr, *d_ent = nullptr, *d_ent2 = nullptr;
  uint8_t* d_isl = nullptr;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  int* d_flags = nullptr;
  int64_t* d_cnt = nullptr;
  void* d_tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_leader); cudaFree(d_members); cudaFree(d_evert); cudaFree(d_ent); cudaFree(d_ent2); cudaFree(d_isl); cudaFree(d_keys);
    cudaFree(d_keys2); cudaFree(d_flags); cudaFree(d_cnt); cudaFree(d_tmp);
  };
#define RN_CUDA(call)                                                                             \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      set_error(ctx, std::string("renumber: ") + #call + ": " + cudaGetErrorString(e__));         \
      cleanup();                                                                                  \
      renumber_free(ctx);                                                                         \
      return NSGPU_ECUDA;                                                                         \
    }                                                                                             \
  } while (0)
  RN_CUDA(cudaMalloc(&d_leader, sizeof(int32_t) * n));
  RN_CUDA(cudaMalloc(&d_members, sizeof(int32_t) * n * 4));
  RN_CUDA(cudaMalloc(&d_evert, sizeof(int32_t) * n));
  RN_CUDA(cudaMalloc(&d_isl, n));
  RN_CUDA(cudaMalloc(&d_flags, 4 * sizeof(int)));
  RN_CUDA(cudaMemsetAsync(d_leader, 0xff, sizeof(int32_t) * n, s));
  RN_CUDA(cudaMemsetAsync(d_flags, 0, 4 * sizeof(int), s));
  k_rn_entities<<<g256(ctx->n_cells_total * 4), 256, 0, s>>>(ctx->n_cells_total, ctx->d_dofmap, ctx->d_cells, d_leader, d_members, d_evert);
  k_rn_check<<<g256(n), 256, 0, s>>>(n, n_owned, d_leader, d_members, d_flags, d_isl);
  int flags[4] = {0, 0, 0, 0};
  RN_CUDA(cudaMemcpyAsync(flags, d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  RN_CUDA(cudaStreamSynchronize(s));
  ctx->launches += 2;
  if (flags[0] || flags[1]) { cleanup(); return NSGPU_OK; }        // not entity-consistent: keep the caller's numbering (generic kernels)
  const bool blocked = flags[2] == 0 && n_owned % 4 == 0;
  if (blocked && ctx->renumber != 2) { cleanup(); return NSGPU_OK; }   // already vertex-blocked: identity

  // compact the leaders (ascending), sort them by (ghost, key)
  RN_CUDA(cudaMalloc(&d_ent, sizeof(int32_t) * n));
  RN_CUDA(cudaMalloc(&d_ent2, sizeof(int32_t) * n));
  RN_CUDA(cudaMalloc(&d_cnt, sizeof(int64_t)));
  cub::CountingInputIterator<int32_t> ids(0);
  size_t tb = 0;
  RN_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids, d_isl, d_ent, d_cnt, n, s));
  RN_CUDA(cudaMalloc(&d_tmp, tb));
  RN_CUDA(cub::DeviceSelect::Flagged(d_tmp, tb, ids, d_isl, d_ent, d_cnt, n, s));
  int64_t n_ent = 0;
  RN_CUDA(cudaMemcpyAsync(&n_ent, d_cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RN_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_tmp); d_tmp = nullptr;
  if (n_ent * 4 != n || n_owned % 4 != 0) { cleanup(); return NSGPU_OK; }   // dofs outside vertex entities: keep the caller's numbering
  RN_CUDA(cudaMalloc(&d_keys, sizeof(uint64_t) * n_ent));
  RN_CUDA(cudaMalloc(&d_keys2, sizeof(uint64_t) * n_ent));
  const double ext = std::max(std::max(ctx->bbox_hi[0] - ctx->bbox_lo[0], ctx->bbox_hi[1] - ctx->bbox_lo[1]), ctx->bbox_hi[2] - ctx->bbox_lo[2]);
  k_rn_keys<<<g256(n_ent), 256, 0, s>>>(n_ent, n_owned, d_ent, d_evert, ctx->d_x, ctx->bbox_lo[0], ctx->bbox_lo[1], ctx->bbox_lo[2],
                                        ext > 0.0 ? 1.0 / ext : 0.0, ctx->renumber_order != 1, d_keys);
  tb = 0;
  RN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys, d_keys2, d_ent, d_ent2, n_ent, 0, 64, s));
  RN_CUDA(cudaMalloc(&d_tmp, tb));
  RN_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys2, d_ent, d_ent2, n_ent, 0, 64, s));
  RN_CUDA(cudaMalloc(&ctx->d_perm, sizeof(int32_t) * n));
  RN_CUDA(cudaMalloc(&ctx->d_iperm, sizeof(int32_t) * n));
  k_rn_assign<<<g256(n_ent), 256, 0, s>>>(n_ent, n_owned, n_owned / 4, d_ent2, d_members, ctx->d_perm, ctx->d_iperm);
  k_rn_map_inplace<<<g256(ctx->n_cells_total * 16), 256, 0, s>>>(ctx->n_cells_total * 16, ctx->d_perm, ctx->d_dofmap);
  ctx->h_perm.resize((size_t)n);
  RN_CUDA(cudaMemcpyAsync(ctx->h_perm.data(), ctx->d_perm, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  RN_CUDA(cudaStreamSynchronize(s));
  RN_CUDA(cudaGetLastError());
  ctx->launches += 6;
  cleanup();
  return NSGPU_OK;
#undef RN_CUDA
}

// Column ghosts (dofs of no local cell that appear in owned rows through other ranks' ghost rows, nsgpu_set_col_ghosts) arrive
// in the caller's order -- with a dolfinx numbering the four dofs of such a vertex are scattered.  Internally they are grouped:
// column ghost (entity rank e, slot s) -> n_dofs + 4 e + s, entities in the order of their first dof.  The permutation arrays are
// extended to n_cols entries (identity when the ghosts are not complete 4-dof entities) and the entity tables are translated.
int renumber_extend_cols(nsgpu_ctx* ctx, int64_t n_extra, const int32_t* leader_local, const int32_t* slot, const int32_t* size,
                         std::vector<int32_t>& o_leader, std::vector<int32_t>& o_slot, std::vector<int32_t>& o_size) {
  o_leader.assign(leader_local, leader_local + n_extra);
  o_slot.assign(slot, slot + n_extra);
  o_size.assign(size, size + n_extra);
  if (!ctx->d_perm) return NSGPU_OK;
  const int64_t n = ctx->n_dofs, nc = n + n_extra;
  ctx->h_perm.resize((size_t)nc);
  for (int64_t k = 0; k < n_extra; ++k) ctx->h_perm[(size_t)(n + k)] = (int32_t)(n + k);
  // entity ranks in the order of the leaders; usable only if every entity is a complete vertex (slots 0..3 once each)
  std::vector<int32_t> leaders(leader_local, leader_local + n_extra);
  std::sort(leaders.begin(), leaders.end());
  leaders.erase(std::unique(leaders.begin(), leaders.end()), leaders.end());
  bool ok = n_extra % 4 == 0 && (int64_t)leaders.size() * 4 == n_extra;
  std::vector<uint8_t> seen((size_t)n_extra, 0);
  std::vector<int32_t> ext((size_t)n_extra, 0);
  for (int64_t k = 0; k < n_extra && ok; ++k) {
    const int64_t e = std::lower_bound(leaders.begin(), leaders.end(), leader_local[k]) - leaders.begin();
    if (size[k] != 4 || slot[k] < 0 || slot[k] > 3) { ok = false; break; }
    const int64_t pos = 4 * e + slot[k];
    if (seen[(size_t)pos]) { ok = false; break; }
    seen[(size_t)pos] = 1;
    ext[(size_t)k] = (int32_t)(n + pos);
  }
  if (ok)
    for (int64_t k = 0; k < n_extra; ++k) {
      ctx->h_perm[(size_t)(n + k)] = ext[(size_t)k];
      const int64_t pos = ext[(size_t)k] - n;
      o_leader[(size_t)pos] = (int32_t)(n + 4 * (pos / 4));
      o_slot[(size_t)pos] = (int32_t)(pos % 4);
      o_size[(size_t)pos] = 4;
    }
  // device copies of the extended permutation and its inverse
  std::vector<int32_t> h_iperm((size_t)nc);
  int32_t *d_p = nullptr, *d_ip = nullptr;
  NS_CUDA(ctx, cudaMalloc(&d_p, sizeof(int32_t) * (size_t)nc));
  NS_CUDA(ctx, cudaMalloc(&d_ip, sizeof(int32_t) * (size_t)nc));
  for (int64_t d = 0; d < nc; ++d) h_iperm[(size_t)ctx->h_perm[(size_t)d]] = (int32_t)d;
  NS_CUDA(ctx, h2d_sync(ctx, d_p, ctx->h_perm.data(), sizeof(int32_t) * (size_t)nc));
  NS_CUDA(ctx, h2d_sync(ctx, d_ip, h_iperm.data(), sizeof(int32_t) * (size_t)nc));
  cudaFree(ctx->d_perm); cudaFree(ctx->d_iperm);
  ctx->d_perm = d_p; ctx->d_iperm = d_ip;
  return NSGPU_OK;
}

// ---------------------------------------------------------------------------------------------- vectors
__global__ void k_perm_in(int64_t n, int64_t n_tot, const int32_t* __restrict__ iperm, const double* __restrict__ src, double* __restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // internal index
  if (i < n) dst[i] = src[iperm[i]];
  else if (i < n_tot) dst[i] = src[i];
}
__global__ void k_perm_out(int64_t n, int64_t n_tot, const int32_t* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // caller index
  if (d < n) dst[d] = src[perm[d]];
  else if (d < n_tot) dst[d] = src[d];
}

// dst (internal numbering) <- src (caller numbering): entries [0, n_perm) permuted, [n_perm, n_tot) copied (column ghosts)
int perm_in(nsgpu_ctx* ctx, const double* d_src, double* d_dst, int64_t n_perm, int64_t n_tot) {
  k_perm_in<<<g256(n_tot), 256, 0, ctx->stream>>>(n_perm, n_tot, ctx->d_iperm, d_src, d_dst);
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}
int perm_out(nsgpu_ctx* ctx, const double* d_src, double* d_dst, int64_t n_perm, int64_t n_tot) {
  k_perm_out<<<g256(n_tot), 256, 0, ctx->stream>>>(n_perm, n_tot, ctx->d_perm, d_src, d_dst);
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

int perm_work(nsgpu_ctx* ctx) {   // staging vectors of the host / device entry points (n_cols each)
  const int64_t n = ctx->n_cols > 0 ? ctx->n_cols : 1;
  if (ctx->d_px && ctx->perm_work_n >= n) return NSGPU_OK;
  cudaFree(ctx->d_px); cudaFree(ctx->d_pF); ctx->d_px = ctx->d_pF = nullptr;
  NS_CUDA(ctx, cudaMalloc(&ctx->d_px, sizeof(double) * n));
  NS_CUDA(ctx, cudaMalloc(&ctx->d_pF, sizeof(double) * n));
  ctx->perm_work_n = n;
  return NSGPU_OK;
}

// ---------------------------------------------------------------------------------------------- pattern / values in the caller's numbering
__global__ void k_cp_len(int64_t n_rows, const int32_t* __restrict__ perm, const int64_t* __restrict__ indptr_i, int64_t* len) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d > n_rows) return;
  if (d == n_rows) { len[d] = 0; return; }
  const int64_t ri = perm[d];
  len[d] = indptr_i[ri + 1] - indptr_i[ri];
}

// one thread per caller row: translate the internal row's columns and sort them (rows are short)
__global__ void k_cp_fill(int64_t n_rows, int64_t n_dofs, const int32_t* __restrict__ perm, const int32_t* __restrict__ iperm,
                          const int64_t* __restrict__ indptr_i, const int32_t* __restrict__ indices_i, const int64_t* __restrict__ indptr_c,
                          int32_t* __restrict__ indices_c) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= n_rows) return;
  const int64_t bi = indptr_i[perm[d]], b = indptr_c[d], len = indptr_c[d + 1] - b;
  for (int64_t k = 0; k < len; ++k) {
    const int32_t ci = indices_i[bi + k];
    const int32_t v = ci < n_dofs ? iperm[ci] : ci;
    int64_t j = k - 1;
    while (j >= 0 && indices_c[b + j] > v) { indices_c[b + j + 1] = indices_c[b + j]; --j; }
    indices_c[b + j + 1] = v;
  }
}

__device__ __forceinline__ int64_t row_find(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t col) {
  const int64_t b = lo;
  --hi;
  while (lo <= hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t c = idx[mid];
    if (c == col) return mid - b;
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// values between the two orders, LPR lanes per caller row.  EXPORT: vals_c <- vals_i, else vals_i <- vals_c
template <bool EXPORT>
__global__ void k_cp_values(int64_t n_rows, int64_t n_dofs, const int32_t* __restrict__ perm, const int32_t* __restrict__ iperm,
                            const int64_t* __restrict__ indptr_i, const int32_t* __restrict__ indices_i, const int64_t* __restrict__ indptr_c,
                            const int32_t* __restrict__ indices_c, double* __restrict__ vals_i, double* __restrict__ vals_c) {
  constexpr int LPR = 8;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t d = t / LPR;
  const int lane = (int)(t % LPR);
  if (d >= n_rows) return;
  const int64_t bi = indptr_i[perm[d]], b = indptr_c[d], e = indptr_c[d + 1];
  for (int64_t k = lane; k < e - b; k += LPR) {
    const int32_t ci = indices_i[bi + k];
    const int32_t v = ci < n_dofs ? iperm[ci] : ci;
    const int64_t j = row_find(indices_c, b, e, v);
    if (j < 0) continue;
    if (EXPORT) vals_c[b + j] = vals_i[bi + k]; else vals_i[bi + k] = vals_c[b + j];
  }
}

int ensure_caller_pattern(nsgpu_ctx* ctx) {
  if (!ctx->d_perm || ctx->caller_pattern_built) return NSGPU_OK;
  cudaStream_t s = ctx->stream;
  const int64_t n = ctx->n_rows;
  int64_t* d_len = nullptr;
  void* d_tmp = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &ctx->d_indptr_c, n + 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_indices_c, ctx->nnz))) return rc;
  NS_CUDA(ctx, cudaMalloc(&d_len, sizeof(int64_t) * (n + 1)));
  k_cp_len<<<g256(n + 1), 256, 0, s>>>(n, ctx->d_perm, ctx->d_indptr, d_len);
  size_t tb = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, d_len, ctx->d_indptr_c, n + 1, s);
  if (e == cudaSuccess) e = cudaMalloc(&d_tmp, tb);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_len, ctx->d_indptr_c, n + 1, s);
  if (e == cudaSuccess) {
    k_cp_fill<<<(unsigned)ceil_div(n > 0 ? n : 1, 64), 64, 0, s>>>(n, ctx->n_cols, ctx->d_perm, ctx->d_iperm, ctx->d_indptr, ctx->d_indices,
                                                                    ctx->d_indptr_c, ctx->d_indices_c);
    e = cudaStreamSynchronize(s);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d_len); cudaFree(d_tmp);
  ctx->launches += 3;
  if (e != cudaSuccess) { set_error(ctx, std::string("caller-order pattern: ") + cudaGetErrorString(e)); return NSGPU_ECUDA; }
  ctx->caller_pattern_built = true;
  return NSGPU_OK;
}

// d_dst (caller CSR order, nnz doubles on the device) <- the resident values
int export_values(nsgpu_ctx* ctx, double* d_dst) {
  int rc = ensure_caller_pattern(ctx);
  if (rc) return rc;
  k_cp_values<true><<<g256(ctx->n_rows * 8), 256, 0, ctx->stream>>>(ctx->n_rows, ctx->n_cols, ctx->d_perm, ctx->d_iperm, ctx->d_indptr, ctx->d_indices,
                                                                      ctx->d_indptr_c, ctx->d_indices_c, ctx->d_vals, d_dst);
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

int import_values(nsgpu_ctx* ctx, double* d_src) {
  int rc = ensure_caller_pattern(ctx);
  if (rc) return rc;
  k_cp_values<false><<<g256(ctx->n_rows * 8), 256, 0, ctx->stream>>>(ctx->n_rows, ctx->n_cols, ctx->d_perm, ctx->d_iperm, ctx->d_indptr, ctx->d_indices,
                                                                       ctx->d_indptr_c, ctx->d_indices_c, ctx->d_vals, d_src);
  ctx->launches += 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

int caller_vals_buffer(nsgpu_ctx* ctx) {
  if (ctx->d_vals_c) return NSGPU_OK;
  NS_CUDA(ctx, cudaMalloc(&ctx->d_vals_c, sizeof(double) * (size_t)(ctx->nnz > 0 ? ctx->nnz : 1)));
  return NSGPU_OK;
}

// CSR positions of the caller-order pattern -> positions of the same entries in the internal value array
__global__ void k_cp_positions(int64_t n, int64_t n_rows, int64_t n_dofs, const int32_t* __restrict__ perm, const int64_t* __restrict__ indptr_i,
                               const int32_t* __restrict__ indices_i, const int64_t* __restrict__ indptr_c, const int32_t* __restrict__ indices_c,
                               const int64_t* __restrict__ pos_c, int64_t* __restrict__ pos_i) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t p = pos_c[t];
  int64_t lo = 0, hi = n_rows - 1;   // last row with indptr_c[row] <= p
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (indptr_c[mid] <= p) lo = mid; else hi = mid - 1;
  }
  const int32_t c = indices_c[p];
  const int32_t ci = c < n_dofs ? perm[c] : c;
  const int64_t ri = perm[lo];
  const int64_t j = row_find(indices_i, indptr_i[ri], indptr_i[ri + 1], ci);
  pos_i[t] = j < 0 ? -1 : indptr_i[ri] + j;
}

int translate_positions(nsgpu_ctx* ctx, int64_t n, const int64_t* h_pos_c, int64_t* d_pos_i) {
  if (n <= 0) return NSGPU_OK;
  int rc = ensure_caller_pattern(ctx);
  if (rc) return rc;
  int64_t* d_pc = nullptr;
  NS_CUDA(ctx, cudaMalloc(&d_pc, sizeof(int64_t) * n));
  cudaError_t e = cudaMemcpyAsync(d_pc, h_pos_c, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    k_cp_positions<<<g256(n), 256, 0, ctx->stream>>>(n, ctx->n_rows, ctx->n_cols, ctx->d_perm, ctx->d_indptr, ctx->d_indices, ctx->d_indptr_c,
                                                    ctx->d_indices_c, d_pc, d_pos_i);
    e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_pc);
  ctx->launches += 1;
  if (e != cudaSuccess) { set_error(ctx, std::string("translate_positions: ") + cudaGetErrorString(e)); return NSGPU_ECUDA; }
  return NSGPU_OK;
}

}  // namespace nsgpu
