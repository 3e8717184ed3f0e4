// ilu.cu -- multicolour block ILU(0) preconditioner for the device-resident KSPTFQMR.
//
// The reference leaves the preconditioner of snes_ksp_type = 'tfqmr' at PETSc's default (NavierStokes/NavierStokesChannelFlow.py:
// 282-291): ILU(0) on one rank, block Jacobi with ILU(0) on each rank's diagonal block under mpirun.  This is that class of
// preconditioner in a form a GPU can run: the incomplete factorisation works on the 4x4 vertex blocks of the P1-P1 matrix (the
// velocity components and the pressure of a vertex are eliminated together) in a multicolour elimination order -- vertices of one
// colour share no matrix entry, so a colour is factorised, and later substituted, by one kernel launch with one thread group
// per vertex.  Couplings to other ranks' vertices are dropped (block Jacobi over the ranks, as PETSc does).
//
//   colouring   Jones-Plassmann with hashed priorities on the vertex graph of the owned rows (deterministic), <= 64 colours
//   order       elimination order = (colour, vertex)
//   factorise   for colour c, vertex i:  for every neighbour k with colour < c, in elimination order:
//                   L_ik = A_ik U_kk^-1 ;  A_ij -= L_ik U_kj  for the blocks (i, j) of the pattern with j after k    [ILU(0)]
//               then U_ii^-1 by Gauss-Jordan with partial pivoting
//   apply       z = U^-1 L^-1 r: one launch per colour forwards (unit lower factor), one per colour backwards
//
// Block access is the vertex-blocked view the block SpMV uses (p1tet.cu): per vertex the list of neighbour vertices
// (ctx->d_pairs) and the CSR starts of its four rows, blocks of four contiguous values per row.
#include <cub/cub.cuh>

#include "common.cuh"
#include "element_p1tet.cuh"

namespace nsgpu {

struct IluPlan {
  int64_t nv = 0;                  // owned vertices
  int n_colours = 0;
  int32_t* d_colour = nullptr;     // per owned vertex
  int32_t* d_order = nullptr;      // vertices sorted by (colour, vertex)
  int4* d_orec = nullptr;          // per position of the order: (vertex, neighbours, first pair lo, hi): one load instead of order -> pair0 / ns
  // packed factor (streaming sweeps): the off-diagonal blocks in elimination order -- per position of the order the L blocks (lower
  // colour) then the U blocks (higher colour) of its vertex, the column vertex inline -- refilled from d_lu after every factorisation
  int4* d_erec = nullptr;          // per position: (vertex, nL | nU << 16, first packed block lo, hi)
  uint32_t* d_emap = nullptr;      // packed block -> pair index (its source in d_lu)
  uint32_t* d_ecol = nullptr;      // packed block -> first dof of the column vertex
  double* d_lue = nullptr;         // 16 doubles per packed block
  int64_t n_packed = 0;
  int max_ns = 0;                  // longest block row (the 16-lane factorisation kernel takes rows of at most 16 blocks)
  std::vector<int64_t> cstart;     // n_colours + 1 offsets into d_order
  double* d_lu = nullptr;          // 16 doubles per (vertex, neighbour) pair, row-major 4x4, indexed like ctx->d_pairs
  double* d_dinv = nullptr;        // 16 doubles per vertex: U_ii^-1
  uint8_t* d_nbc = nullptr;        // per (vertex, neighbour) pair: colour of the neighbour (255: other rank's vertex, 254: the vertex itself)
  bool unsupported = false;
};

__device__ __forceinline__ uint32_t ilu_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void k_ilu_check(int64_t nv, const int32_t* __restrict__ rowdof, int* bad) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < nv && rowdof[4 * e] != (int32_t)(4 * e)) *bad = 1;
}

// one Jones-Plassmann round: an uncoloured vertex whose priority beats every uncoloured neighbour takes the lowest free colour
__global__ void k_ilu_colour(int64_t nv, int64_t n_owned, const int64_t* __restrict__ pair0, const int32_t* __restrict__ ns,
                             const uint64_t* __restrict__ pairs, const int32_t* __restrict__ colour, int32_t* colour_out, unsigned long long* left,
                             int* too_many) {
  // decisions are taken on the colours at the start of the round (colour), results go to colour_out: the colouring does not
  // depend on the order in which the threads run
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nv || colour[e] >= 0) return;
  const uint32_t pe = ilu_hash((uint32_t)e);
  unsigned long long used = 0;
  const int64_t p0 = pair0[e];
  const int n = ns[e];
  for (int s = 0; s < n; ++s) {
    const int64_t B = (int64_t)(pairs[p0 + s] & 0xffffffffu);
    if (B >= n_owned) continue;
    const int64_t k = B >> 2;
    if (k == e) continue;
    const int ck = colour[k];
    if (ck >= 0) { used |= 1ull << ck; continue; }
    const uint32_t pk = ilu_hash((uint32_t)k);
    if (pk > pe || (pk == pe && k > e)) { atomicAdd(left, 1ull); return; }   // a neighbour goes first
  }
  int c = 0;
  while (c < 64 && ((used >> c) & 1ull)) ++c;
  if (c >= 64) { *too_many = 1; c = 63; }
  colour_out[e] = c;
}

__global__ void k_ilu_keys(int64_t nv, const int32_t* __restrict__ colour, uint64_t* keys) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < nv) keys[e] = ((uint64_t)(uint32_t)colour[e] << 32) | (uint64_t)e;
}

__global__ void k_ilu_order(int64_t nv, const uint64_t* __restrict__ keys, int32_t* order, unsigned long long* count) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nv) return;
  order[i] = (int32_t)(keys[i] & 0xffffffffu);
  atomicAdd(count + (keys[i] >> 32), 1ull);
}

// neighbour colours next to the neighbour lists: the substitution sweeps read them in a stream instead of gathering colour[]
__global__ void k_ilu_nbc(int64_t nv, int64_t n_owned, const int64_t* __restrict__ pair0, const int32_t* __restrict__ ns,
                          const uint64_t* __restrict__ pairs, const int32_t* __restrict__ colour, uint8_t* __restrict__ nbc) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t e = t >> 4;
  if (e >= nv) return;
  const int64_t p0 = pair0[e];
  const int n = ns[e];
  for (int s = (int)(t & 15); s < n; s += 16) {
    const int64_t B = (int64_t)(pairs[p0 + s] & 0xffffffffu);
    nbc[p0 + s] = B >= n_owned ? 255 : ((B >> 2) == e ? 254 : (uint8_t)colour[B >> 2]);
  }
}

// the sweeps' per-vertex record in elimination order
__global__ void k_ilu_orec(int64_t nv, const int32_t* __restrict__ order, const int64_t* __restrict__ pair0, const int32_t* __restrict__ ns, int4* __restrict__ orec) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= nv) return;
  const int32_t i = order[idx];
  const int64_t p0 = pair0[i];
  orec[idx] = make_int4(i, ns[i], (int)(p0 & 0xffffffffLL), (int)(p0 >> 32));
}

// LU blocks <- the owned x owned blocks of the assembled Jacobian
__global__ void k_ilu_load(int64_t nv, int64_t n_owned, const int64_t* __restrict__ pair0, const int32_t* __restrict__ ns,
                           const uint64_t* __restrict__ pairs, const int64_t* __restrict__ rowpos, const double* __restrict__ vals, double* __restrict__ lu) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t e = t >> 4;
  const int lane = (int)(t & 15);
  if (e >= nv) return;
  const int64_t p0 = pair0[e];
  const int n = ns[e];
  for (int s = lane; s < n; s += 16) {
    const bool own = (int64_t)(pairs[p0 + s] & 0xffffffffu) < n_owned;
    double* o = lu + 16 * (p0 + s);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double* v = vals + rowpos[4 * e + c] + 4 * s;
#pragma unroll
      for (int d = 0; d < 4; ++d) o[4 * c + d] = own ? v[d] : 0.0;
    }
  }
}

__device__ __forceinline__ void blk_load(const double* p, double (&A)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) A[r][c] = p[4 * r + c];
}
__device__ __forceinline__ void blk_store(double* p, const double (&A)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) p[4 * r + c] = A[r][c];
}

__device__ bool blk_inverse(double (&A)[4][4], double (&B)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) B[r][c] = (r == c) ? 1.0 : 0.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int piv = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r > k && fabs(A[r][k]) > best) { best = fabs(A[r][k]); piv = r; }
    if (best == 0.0 || !isfinite(best)) { ok = false; break; }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r == piv && piv != k) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { double t = A[k][c]; A[k][c] = A[r][c]; A[r][c] = t; t = B[k][c]; B[k][c] = B[r][c]; B[r][c] = t; }
      }
    const double ip = 1.0 / A[k][k];
#pragma unroll
    for (int c = 0; c < 4; ++c) { A[k][c] *= ip; B[k][c] *= ip; }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r != k) {
        const double f = A[r][k];
#pragma unroll
        for (int c = 0; c < 4; ++c) { A[r][c] -= f * A[k][c]; B[r][c] -= f * B[k][c]; }
      }
  }
  if (!ok) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) B[r][c] = (r == c) ? 1.0 : 0.0;
  }
  return ok;
}

// factorise the vertices order[i0 .. i1) (one colour).  One thread per vertex.
__global__ void __launch_bounds__(64)
k_ilu_factor(int64_t i0, int64_t i1, int cc, int64_t n_owned, const int32_t* __restrict__ order, const int32_t* __restrict__ colour,
             const int64_t* __restrict__ pair0, const int32_t* __restrict__ ns, const uint64_t* __restrict__ pairs, double* lu, double* dinv,
             int* singular) {
  const int64_t idx = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= i1) return;
  const int64_t i = order[idx];
  const int64_t p0 = pair0[i];
  const int n = ns[i];
  int sdiag = -1;
  // eliminate the earlier neighbours in elimination order (colour, vertex): repeated selection of the smallest unprocessed one
  long long last = -1;    // key of the neighbour processed last
  for (;;) {
    long long best = -1;
    int sb = -1;
    for (int s = 0; s < n; ++s) {
      const int64_t B = (int64_t)(pairs[p0 + s] & 0xffffffffu);
      if (B >= n_owned) continue;
      const int64_t k = B >> 2;
      if (k == i) { sdiag = s; continue; }
      const int ck = colour[k];
      if (ck >= cc) continue;
      const long long key = ((long long)ck << 32) | k;
      if (key > last && (best < 0 || key < best)) { best = key; sb = s; }
    }
    if (sb < 0) break;
    last = best;
    const int64_t k = best & 0xffffffffLL;
    const int ck = (int)(best >> 32);
    // L_ik = A_ik U_kk^-1
    double Aik[4][4], Dk[4][4], L[4][4];
    blk_load(lu + 16 * (p0 + sb), Aik);
    blk_load(dinv + 16 * k, Dk);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) L[r][c] = Aik[r][0] * Dk[0][c] + Aik[r][1] * Dk[1][c] + Aik[r][2] * Dk[2][c] + Aik[r][3] * Dk[3][c];
    blk_store(lu + 16 * (p0 + sb), L);
    // A_ij -= L_ik U_kj for the pattern blocks (i, j) with j after k
    const int64_t q0 = pair0[k];
    const int nk = ns[k];
    for (int t = 0; t < n; ++t) {
      const int64_t Bj = (int64_t)(pairs[p0 + t] & 0xffffffffu);
      if (Bj >= n_owned) continue;
      const int64_t j = Bj >> 2;
      if (j == k) continue;
      const int cj = j == i ? cc : colour[j];
      if (cj < ck || (cj == ck && j < k)) continue;        // j before k: that block is an L block already final
      // (k, j) in row k?  neighbour lists are sorted by first dof
      int lo = 0, hi = nk - 1, u = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int64_t Bm = (int64_t)(pairs[q0 + mid] & 0xffffffffu);
        if (Bm == Bj) { u = mid; break; }
        if (Bm < Bj) lo = mid + 1; else hi = mid - 1;
      }
      if (u < 0) continue;
      double U[4][4], A[4][4];
      blk_load(lu + 16 * (q0 + u), U);
      blk_load(lu + 16 * (p0 + t), A);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) A[r][c] -= L[r][0] * U[0][c] + L[r][1] * U[1][c] + L[r][2] * U[2][c] + L[r][3] * U[3][c];
      blk_store(lu + 16 * (p0 + t), A);
    }
  }
  double D[4][4], Di[4][4];
  if (sdiag >= 0) blk_load(lu + 16 * (p0 + sdiag), D);
  else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) D[r][c] = (r == c) ? 1.0 : 0.0;
  }
  if (!blk_inverse(D, Di)) *singular = 1;
  blk_store(dinv + 16 * i, Di);
}

// The same factorisation with sixteen lanes per vertex (rows of at most 16 blocks): lane t keeps block (i, slot t) in registers from the
// first elimination step to the last; per step the lanes agree on the next earlier neighbour k (minimum of the (colour, vertex) keys
// above the last one), the lane that holds A_ik turns it into L_ik = A_ik U_kk^-1 and hands it round through shared memory, and every
// lane whose column comes after k looks (k, j) up in row k and updates its block.  Same operations per block in the same order as
// k_ilu_factor (bitwise the same factor); the dependent loads per vertex drop from (earlier neighbours x slots x search steps) to
// (earlier neighbours x search steps).
__global__ void __launch_bounds__(128)
k_ilu_factor16(int64_t i0, int64_t i1, int cc, int64_t n_owned, const int32_t* __restrict__ order, const int32_t* __restrict__ colour,
               const int64_t* __restrict__ pair0, const int32_t* __restrict__ ns, const uint64_t* __restrict__ pairs, double* lu, double* dinv,
               int* singular) {
  __shared__ double sL[8][16];
  const int grp = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const unsigned gm = 0xffffu << (16 * (grp & 1));
  const int64_t idx = i0 + (int64_t)blockIdx.x * 8 + grp;
  const bool valid = idx < i1;
  const int64_t i = valid ? order[idx] : 0;
  const int64_t p0 = valid ? pair0[i] : 0;
  const int n = valid ? ns[i] : 0;
  const long long NONE = 0x7fffffffffffffffLL;
  int64_t B = -1, kv = -1;
  int ck = 255;
  double A[4][4];
  bool owned = false;
  if (lane < n) {
    B = (int64_t)(pairs[p0 + lane] & 0xffffffffu);
    owned = B < n_owned;
    if (owned) {
      kv = B >> 2;
      ck = kv == i ? cc : colour[kv];
      blk_load(lu + 16 * (p0 + lane), A);
    }
  }
  const bool isdiag = owned && kv == i;
  const long long key = (owned && !isdiag && ck < cc) ? (((long long)ck << 32) | kv) : NONE;
  long long last = -1;
  for (;;) {
    long long best = key > last ? key : NONE;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const long long other = __shfl_xor_sync(gm, best, o, 16);
      best = other < best ? other : best;
    }
    if (best == NONE) break;                                  // uniform over the group
    last = best;
    const int64_t k = best & 0xffffffffLL;
    const int ckk = (int)(best >> 32);
    if (key == best) {                                        // L_ik = A_ik U_kk^-1
      double Dk[4][4], L[4][4];
      blk_load(dinv + 16 * k, Dk);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) L[r][c] = A[r][0] * Dk[0][c] + A[r][1] * Dk[1][c] + A[r][2] * Dk[2][c] + A[r][3] * Dk[3][c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { A[r][c] = L[r][c]; sL[grp][4 * r + c] = L[r][c]; }
    }
    __syncwarp(gm);
    if (owned && key != best) {
      const int cj = isdiag ? cc : ck;
      if (!(cj < ckk || (cj == ckk && kv < k))) {             // j after k: (k, j) in row k?  neighbour lists are sorted by first dof
        const int64_t q0 = pair0[k];
        int lo = 0, hi = ns[k] - 1, u = -1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const int64_t Bm = (int64_t)(pairs[q0 + mid] & 0xffffffffu);
          if (Bm == B) { u = mid; break; }
          if (Bm < B) lo = mid + 1; else hi = mid - 1;
        }
        if (u >= 0) {
          double U[4][4];
          blk_load(lu + 16 * (q0 + u), U);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              A[r][c] -= sL[grp][4 * r] * U[0][c] + sL[grp][4 * r + 1] * U[1][c] + sL[grp][4 * r + 2] * U[2][c] + sL[grp][4 * r + 3] * U[3][c];
        }
      }
    }
    __syncwarp(gm);
  }
  if (owned) blk_store(lu + 16 * (p0 + lane), A);
  const unsigned anyd = __ballot_sync(gm, isdiag) & gm;
  if (isdiag) {
    double Di[4][4];
    if (!blk_inverse(A, Di)) *singular = 1;
    blk_store(dinv + 16 * i, Di);
  } else if (valid && anyd == 0 && lane == 0) {               // no diagonal block in the pattern: identity, as in k_ilu_factor
    double Di[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) Di[r][c] = (r == c) ? 1.0 : 0.0;
    blk_store(dinv + 16 * i, Di);
  }
}

__global__ void k_ilu_maxns(int64_t nv, const int32_t* __restrict__ ns, int* out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < nv) atomicMax(out, ns[e]);
}

// one colour of the forward (LOWER: z_i -= sum over earlier neighbours L_ik z_k) or backward (z_i = U_ii^-1 (z_i - sum over later
// neighbours U_ij z_j)) substitution, in place.  Sixteen lanes per vertex.
// WIDE (32-byte aligned factor and vector): four lanes per block -- lane (slot, r) loads row r of the block of neighbour slot, slot + 4, ...
// with one 256-bit load, so that one load instruction of the four lanes asks for the whole 128-byte block at once -- and the row
// sums are reduced over the four slots.  Otherwise: lane s takes neighbour s with scalar loads.
template <bool LOWER, bool WIDE>
__global__ void __launch_bounds__(256)
k_ilu_sweep(int64_t i0, int64_t i1, int cc, const int4* __restrict__ orec, const uint8_t* __restrict__ nbc,
            const uint64_t* __restrict__ pairs, const double* __restrict__ lu, const double* __restrict__ dinv, double* z) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t idx = i0 + (t >> 4);
  const int lane = (int)(t & 15);
  int64_t i = 0;
  if (WIDE) {
    const int r = lane & 3, slot = lane >> 2;
    double acc = 0.0;
    if (idx < i1) {
      const int4 rec = orec[idx];
      i = rec.x;
      const int n = rec.y;
      const int64_t p0 = (int64_t)(uint32_t)rec.z | ((int64_t)rec.w << 32);
#pragma unroll 2
      for (int s = slot; s < n; s += 4) {
        const int ck = nbc[p0 + s];
        const uint64_t pw = pairs[p0 + s];          // asked for together with the colour byte: one memory round trip less on the chain
        if (LOWER ? ck >= cc : (ck <= cc || ck >= 254)) continue;
        const int64_t B = (int64_t)(pw & 0xffffffffu);
        const double4 m = ld256_stream(lu + 16 * (p0 + s) + 4 * r);
        const double4 zz = ld256(z + B);
        acc += m.x * zz.x + m.y * zz.y + m.z * zz.z + m.w * zz.w;
      }
    }
    acc += __shfl_down_sync(0xffffffffu, acc, 8, 16);
    acc += __shfl_down_sync(0xffffffffu, acc, 4, 16);
    // lanes 0..3 of the group hold the row sums; y = z_i - sums
    double y = 0.0;
    if (idx < i1 && slot == 0) y = z[4 * i + r] - acc;
    if (LOWER) {
      if (idx < i1 && slot == 0) z[4 * i + r] = y;
    } else {
      const double y0 = __shfl_sync(0xffffffffu, y, 0, 16), y1 = __shfl_sync(0xffffffffu, y, 1, 16);
      const double y2 = __shfl_sync(0xffffffffu, y, 2, 16), y3 = __shfl_sync(0xffffffffu, y, 3, 16);
      if (idx < i1 && slot == 0) {
        const double4 d = ld256_nc(dinv + 16 * i + 4 * r);
        z[4 * i + r] = d.x * y0 + d.y * y1 + d.z * y2 + d.w * y3;
      }
    }
    return;
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (idx < i1) {
    const int4 rec = orec[idx];
    i = rec.x;
    const int n = rec.y;
    const int64_t p0 = (int64_t)(uint32_t)rec.z | ((int64_t)rec.w << 32);
    for (int s = lane; s < n; s += 16) {
      const int ck = nbc[p0 + s];
      const uint64_t pw = pairs[p0 + s];
      if (LOWER ? ck >= cc : (ck <= cc || ck >= 254)) continue;
      const int64_t B = (int64_t)(pw & 0xffffffffu);
      const double* M = lu + 16 * (p0 + s);
      const double z0 = z[B], z1 = z[B + 1], z2 = z[B + 2], z3 = z[B + 3];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += M[4 * r] * z0 + M[4 * r + 1] * z1 + M[4 * r + 2] * z2 + M[4 * r + 3] * z3;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc[r] += __shfl_down_sync(0xffffffffu, acc[r], o, 16);
  }
  if (idx < i1 && lane == 0) {
    double y[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) y[r] = z[4 * i + r] - acc[r];
    if (LOWER) {
#pragma unroll
      for (int r = 0; r < 4; ++r) z[4 * i + r] = y[r];
    } else {
      const double* D = dinv + 16 * i;
#pragma unroll
      for (int r = 0; r < 4; ++r) z[4 * i + r] = D[4 * r] * y[0] + D[4 * r + 1] * y[1] + D[4 * r + 2] * y[2] + D[4 * r + 3] * y[3];
    }
  }
}

// ---- packed factor: counts, fill, refill, sweeps
__global__ void k_ilu_ecount(int64_t nv, const int4* __restrict__ orec, const uint8_t* __restrict__ nbc, const int32_t* __restrict__ colour,
                             int64_t* __restrict__ cnt) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx > nv) return;
  if (idx == nv) { cnt[idx] = 0; return; }
  const int4 rec = orec[idx];
  const int64_t p0 = (int64_t)(uint32_t)rec.z | ((int64_t)rec.w << 32);
  int m = 0;
  for (int s = 0; s < rec.y; ++s) m += nbc[p0 + s] < 254 ? 1 : 0;      // 254: the vertex itself, 255: another rank's vertex
  cnt[idx] = m;
}
__global__ void k_ilu_efill(int64_t nv, const int4* __restrict__ orec, const uint8_t* __restrict__ nbc, const int32_t* __restrict__ colour,
                            const uint64_t* __restrict__ pairs, const int64_t* __restrict__ eoff, int4* __restrict__ erec, uint32_t* __restrict__ emap,
                            uint32_t* __restrict__ ecol) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= nv) return;
  const int4 rec = orec[idx];
  const int64_t p0 = (int64_t)(uint32_t)rec.z | ((int64_t)rec.w << 32);
  const int cc = colour[rec.x];
  const int64_t q0 = eoff[idx];
  int64_t q = q0;
  int nL = 0, nU = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int s = 0; s < rec.y; ++s) {
      const int ck = nbc[p0 + s];
      if (ck >= 254 || (pass == 0 ? ck >= cc : ck <= cc)) continue;
      emap[q] = (uint32_t)(p0 + s);
      ecol[q] = (uint32_t)(pairs[p0 + s] & 0xffffffffu);
      ++q;
      if (pass == 0) ++nL; else ++nU;
    }
  erec[idx] = make_int4(rec.x, nL | (nU << 16), (int)(q0 & 0xffffffffLL), (int)(q0 >> 32));
}
// packed blocks <- factor (four lanes per block, one 32-byte row each)
__global__ void k_ilu_pack(int64_t n_packed, const uint32_t* __restrict__ emap, const double* __restrict__ lu, double* __restrict__ lue) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t e = t >> 2;
  const int r = (int)(t & 3);
  if (e >= n_packed) return;
  st256(lue + 16 * e + 4 * r, ld256_stream(lu + 16 * (int64_t)emap[e] + 4 * r));
}
// one colour of the substitution on the packed factor: the blocks of the launch are one contiguous stream
template <bool LOWER>
__global__ void __launch_bounds__(256)
k_ilu_sweep_packed(int64_t i0, int64_t i1, const int4* __restrict__ erec, const uint32_t* __restrict__ ecol, const double* __restrict__ lue,
                   const double* __restrict__ dinv, double* z) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t idx = i0 + (t >> 4);
  const int lane = (int)(t & 15), r = lane & 3, slot = lane >> 2;
  int64_t i = 0;
  double acc = 0.0;
  if (idx < i1) {
    const int4 rec = erec[idx];
    i = rec.x;
    const int nL = rec.y & 0xffff, nU = (unsigned)rec.y >> 16;
    const int64_t q0 = ((int64_t)(uint32_t)rec.z | ((int64_t)rec.w << 32)) + (LOWER ? 0 : nL);
    const int cnt = LOWER ? nL : nU;
#pragma unroll 2
    for (int k = slot; k < cnt; k += 4) {
      const uint32_t B = ecol[q0 + k];
      const double4 m = ld256_stream(lue + 16 * (q0 + k) + 4 * r);
      const double4 zz = ld256(z + B);
      acc += m.x * zz.x + m.y * zz.y + m.z * zz.z + m.w * zz.w;
    }
  }
  acc += __shfl_down_sync(0xffffffffu, acc, 8, 16);
  acc += __shfl_down_sync(0xffffffffu, acc, 4, 16);
  double y = 0.0;
  if (idx < i1 && slot == 0) y = z[4 * i + r] - acc;
  if (LOWER) {
    if (idx < i1 && slot == 0) z[4 * i + r] = y;
  } else {
    const double y0 = __shfl_sync(0xffffffffu, y, 0, 16), y1 = __shfl_sync(0xffffffffu, y, 1, 16);
    const double y2 = __shfl_sync(0xffffffffu, y, 2, 16), y3 = __shfl_sync(0xffffffffu, y, 3, 16);
    if (idx < i1 && slot == 0) {
      const double4 d = ld256_nc(dinv + 16 * i + 4 * r);
      z[4 * i + r] = d.x * y0 + d.y * y1 + d.z * y2 + d.w * y3;
    }
  }
}

void ilu_free(nsgpu_ctx* ctx) {
  IluPlan* P = static_cast<IluPlan*>(ctx->ilu);
  if (!P) return;
  cudaFree(P->d_colour); cudaFree(P->d_order); cudaFree(P->d_lu); cudaFree(P->d_dinv); cudaFree(P->d_nbc); cudaFree(P->d_orec);
  cudaFree(P->d_erec); cudaFree(P->d_emap); cudaFree(P->d_ecol); cudaFree(P->d_lue);
  delete P;
  ctx->ilu = nullptr;
}

static inline unsigned g256(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, 256); }

// colouring and elimination order (once per pattern)
static int ilu_plan(nsgpu_ctx* ctx, const P1BlockView& V) {
  ilu_free(ctx);
  IluPlan* P = new IluPlan();
  ctx->ilu = P;
  cudaStream_t s = ctx->stream;
  const int64_t nv = ctx->n_owned / 4;
  P->nv = nv;
  int* d_flag = nullptr;
  unsigned long long* d_cnt = nullptr;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  void* d_tmp = nullptr;
  auto cleanup = [&]() { cudaFree(d_flag); cudaFree(d_cnt); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_tmp); };
#define IL_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      set_error(ctx, std::string("ilu: " #call ": ") + cudaGetErrorString(e__));                   \
      cleanup(); ilu_free(ctx);                                                                    \
      return NSGPU_ECUDA;                                                                          \
    }                                                                                              \
  } while (0)
  IL_CUDA(cudaMalloc(&d_flag, 2 * sizeof(int)));
  IL_CUDA(cudaMalloc(&d_cnt, 65 * sizeof(unsigned long long)));
  IL_CUDA(cudaMemsetAsync(d_flag, 0, 2 * sizeof(int), s));
  if (ctx->n_owned % 4 != 0 || nv <= 0 || nv > V.n_ent || nv >= (int64_t(1) << 31)) { P->unsupported = true; cleanup(); return NSGPU_OK; }
  k_ilu_check<<<g256(nv), 256, 0, s>>>(nv, V.rowdof, d_flag);
  int rc;
  if ((rc = dev_alloc(ctx, &P->d_colour, nv)) || (rc = dev_alloc(ctx, &P->d_order, nv))) { cleanup(); ilu_free(ctx); return rc; }
  IL_CUDA(cudaMemsetAsync(P->d_colour, 0xff, sizeof(int32_t) * nv, s));
  int32_t* d_prev = reinterpret_cast<int32_t*>(P->d_order);   // round-start snapshot (the order array is filled later)
  for (int round = 0; round < 1000; ++round) {
    IL_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
    IL_CUDA(cudaMemcpyAsync(d_prev, P->d_colour, sizeof(int32_t) * nv, cudaMemcpyDeviceToDevice, s));
    k_ilu_colour<<<g256(nv), 256, 0, s>>>(nv, ctx->n_owned, V.pair0, V.ns, ctx->d_pairs, d_prev, P->d_colour, d_cnt, d_flag + 1);
    unsigned long long left = 0;
    IL_CUDA(cudaMemcpyAsync(&left, d_cnt, sizeof(left), cudaMemcpyDeviceToHost, s));
    IL_CUDA(cudaStreamSynchronize(s));
    ctx->launches += 1;
    if (left == 0) break;
  }
  int flags[2] = {0, 0};
  IL_CUDA(cudaMemcpyAsync(flags, d_flag, sizeof(flags), cudaMemcpyDeviceToHost, s));
  IL_CUDA(cudaStreamSynchronize(s));
  if (flags[0] || flags[1]) { P->unsupported = true; cleanup(); return NSGPU_OK; }   // not the blocked numbering / more than 64 colours
  IL_CUDA(cudaMalloc(&d_keys, sizeof(uint64_t) * nv));
  IL_CUDA(cudaMalloc(&d_keys2, sizeof(uint64_t) * nv));
  k_ilu_keys<<<g256(nv), 256, 0, s>>>(nv, P->d_colour, d_keys);
  size_t tmp_bytes = 0;
  IL_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_keys2, nv, 0, 40, s));
  IL_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
  IL_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys2, nv, 0, 40, s));
  IL_CUDA(cudaMemsetAsync(d_cnt, 0, 65 * sizeof(unsigned long long), s));
  k_ilu_order<<<g256(nv), 256, 0, s>>>(nv, d_keys2, P->d_order, d_cnt);
  unsigned long long cnt[65];
  IL_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  IL_CUDA(cudaStreamSynchronize(s));
  IL_CUDA(cudaGetLastError());
  ctx->launches += 3;
  P->cstart.assign(1, 0);
  for (int c = 0; c < 64; ++c) {
    if (cnt[c] == 0) break;
    P->cstart.push_back(P->cstart.back() + (int64_t)cnt[c]);
  }
  P->n_colours = (int)P->cstart.size() - 1;
  if (P->cstart.back() != nv) { P->unsupported = true; cleanup(); return NSGPU_OK; }
  if ((rc = dev_alloc(ctx, &P->d_lu, 16 * ctx->n_pairs)) || (rc = dev_alloc(ctx, &P->d_dinv, 16 * nv)) || (rc = dev_alloc(ctx, &P->d_nbc, ctx->n_pairs)) ||
      (rc = dev_alloc(ctx, &P->d_orec, nv))) {
    cleanup(); ilu_free(ctx); return rc;
  }
  k_ilu_nbc<<<g256(nv * 16), 256, 0, s>>>(nv, ctx->n_owned, V.pair0, V.ns, ctx->d_pairs, P->d_colour, P->d_nbc);
  k_ilu_orec<<<g256(nv), 256, 0, s>>>(nv, P->d_order, V.pair0, V.ns, P->d_orec);
  ctx->launches += 2;
  IL_CUDA(cudaGetLastError());
  IL_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
  k_ilu_maxns<<<g256(nv), 256, 0, s>>>(nv, V.ns, d_flag);
  IL_CUDA(cudaMemcpyAsync(&P->max_ns, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  IL_CUDA(cudaStreamSynchronize(s));
  ctx->launches += 1;
  // packed factor tables (optional: without the memory for a second copy of the blocks the sweeps walk d_lu through the neighbour lists)
  if (ctx->ilu_packed) {
    int64_t *d_ecnt = nullptr, *d_eoff = nullptr;
    void* d_tmp2 = nullptr;
    bool ok = cudaMalloc(&d_ecnt, sizeof(int64_t) * (nv + 1)) == cudaSuccess && cudaMalloc(&d_eoff, sizeof(int64_t) * (nv + 1)) == cudaSuccess;
    if (ok) {
      k_ilu_ecount<<<g256(nv + 1), 256, 0, s>>>(nv, P->d_orec, P->d_nbc, P->d_colour, d_ecnt);
      size_t tb = 0;
      ok = cub::DeviceScan::ExclusiveSum(nullptr, tb, d_ecnt, d_eoff, nv + 1, s) == cudaSuccess && cudaMalloc(&d_tmp2, tb) == cudaSuccess &&
           cub::DeviceScan::ExclusiveSum(d_tmp2, tb, d_ecnt, d_eoff, nv + 1, s) == cudaSuccess;
    }
    int64_t np = 0;
    if (ok) ok = cudaMemcpyAsync(&np, d_eoff + nv, sizeof(int64_t), cudaMemcpyDeviceToHost, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
    if (ok && np > 0 && np < (int64_t(1) << 32) && ctx->n_pairs < (int64_t(1) << 32)) {
      ok = cudaMalloc(&P->d_erec, sizeof(int4) * nv) == cudaSuccess && cudaMalloc(&P->d_emap, sizeof(uint32_t) * np) == cudaSuccess &&
           cudaMalloc(&P->d_ecol, sizeof(uint32_t) * np) == cudaSuccess && cudaMalloc(&P->d_lue, sizeof(double) * 16 * np) == cudaSuccess;
      if (ok) {
        k_ilu_efill<<<g256(nv), 256, 0, s>>>(nv, P->d_orec, P->d_nbc, P->d_colour, ctx->d_pairs, d_eoff, P->d_erec, P->d_emap, P->d_ecol);
        ok = cudaStreamSynchronize(s) == cudaSuccess && cudaGetLastError() == cudaSuccess;
        ctx->launches += 2;
        P->n_packed = np;
      }
    } else ok = false;
    if (!ok) {   // not fatal: fall back to the unpacked sweeps
      cudaGetLastError();
      cudaFree(P->d_erec); cudaFree(P->d_emap); cudaFree(P->d_ecol); cudaFree(P->d_lue);
      P->d_erec = nullptr; P->d_emap = nullptr; P->d_ecol = nullptr; P->d_lue = nullptr; P->n_packed = 0;
    }
    cudaFree(d_ecnt); cudaFree(d_eoff); cudaFree(d_tmp2);
  }
  cleanup();
#undef IL_CUDA
  return NSGPU_OK;
}

// factorise the Jacobian now resident in ctx->d_vals.  NSGPU_EUNSUPPORTED when the vertex-blocked view does not exist.
int ilu_factor(nsgpu_ctx* ctx) {
  P1BlockView V;
  if (!p1tet_block_view(ctx, &V)) { set_error(ctx, "pc = ILU needs the vertex-blocked P1-P1 tet layout (block SpMV layout)"); return NSGPU_EUNSUPPORTED; }
  IluPlan* P = static_cast<IluPlan*>(ctx->ilu);
  if (!P) {
    const int rc = ilu_plan(ctx, V);
    if (rc) return rc;
    P = static_cast<IluPlan*>(ctx->ilu);
  }
  if (P->unsupported) { set_error(ctx, "pc = ILU: the numbering is not vertex-blocked or the vertex graph needs more than 64 colours"); return NSGPU_EUNSUPPORTED; }
  cudaStream_t s = ctx->stream;
  int* d_sing = nullptr;
  NS_CUDA(ctx, cudaMalloc(&d_sing, sizeof(int)));
  cudaMemsetAsync(d_sing, 0, sizeof(int), s);
  k_ilu_load<<<g256(P->nv * 16), 256, 0, s>>>(P->nv, ctx->n_owned, V.pair0, V.ns, ctx->d_pairs, V.rowpos, ctx->d_vals, P->d_lu);
  for (int c = 0; c < P->n_colours; ++c) {
    const int64_t i0 = P->cstart[c], i1 = P->cstart[c + 1];
    if (P->max_ns <= 16 && ctx->ilu_factor16)
      k_ilu_factor16<<<(unsigned)ceil_div(i1 - i0, 8), 128, 0, s>>>(i0, i1, c, ctx->n_owned, P->d_order, P->d_colour, V.pair0, V.ns, ctx->d_pairs, P->d_lu,
                                                                      P->d_dinv, d_sing);
    else
      k_ilu_factor<<<(unsigned)ceil_div(i1 - i0, 64), 64, 0, s>>>(i0, i1, c, ctx->n_owned, P->d_order, P->d_colour, V.pair0, V.ns, ctx->d_pairs, P->d_lu,
                                                                     P->d_dinv, d_sing);
  }
  ctx->launches += 1 + P->n_colours;
  if (P->d_lue) {
    k_ilu_pack<<<g256(P->n_packed * 4), 256, 0, s>>>(P->n_packed, P->d_emap, P->d_lu, P->d_lue);
    ctx->launches += 1;
  }
  int sing = 0;
  cudaError_t e = cudaMemcpyAsync(&sing, d_sing, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d_sing);
  if (e != cudaSuccess) { set_error(ctx, std::string("ilu factorisation: ") + cudaGetErrorString(e)); return NSGPU_ECUDA; }
  if (sing) { set_error(ctx, "pc = ILU: a singular 4x4 pivot block (replaced by the identity)"); }   // not fatal: PETSc would shift; the block is skipped
  return NSGPU_OK;
}

// d_z (owned entries) = U^-1 L^-1 d_r.  d_z may alias d_r.
int ilu_apply(nsgpu_ctx* ctx, const double* d_r, double* d_z) {
  IluPlan* P = static_cast<IluPlan*>(ctx->ilu);
  P1BlockView V;
  if (!P || P->unsupported || !P->d_lu || !p1tet_block_view(ctx, &V)) { set_error(ctx, "ilu_apply: no factorisation"); return NSGPU_EINVAL; }
  cudaStream_t s = ctx->stream;
  if (d_z != d_r) NS_CUDA(ctx, cudaMemcpyAsync(d_z, d_r, sizeof(double) * (size_t)ctx->n_owned, cudaMemcpyDeviceToDevice, s));
  const bool wide = (reinterpret_cast<uintptr_t>(d_z) & 31) == 0 && (reinterpret_cast<uintptr_t>(P->d_lu) & 31) == 0;   // 256-bit loads
  if (wide && P->d_lue) {
    for (int c = 1; c < P->n_colours; ++c) {
      const int64_t i0 = P->cstart[c], i1 = P->cstart[c + 1];
      k_ilu_sweep_packed<true><<<g256((i1 - i0) * 16), 256, 0, s>>>(i0, i1, P->d_erec, P->d_ecol, P->d_lue, P->d_dinv, d_z);
    }
    for (int c = P->n_colours - 1; c >= 0; --c) {
      const int64_t i0 = P->cstart[c], i1 = P->cstart[c + 1];
      k_ilu_sweep_packed<false><<<g256((i1 - i0) * 16), 256, 0, s>>>(i0, i1, P->d_erec, P->d_ecol, P->d_lue, P->d_dinv, d_z);
    }
    ctx->launches += 2 * P->n_colours - 1;
    NS_CUDA(ctx, cudaGetLastError());
    return NSGPU_OK;
  }
  for (int c = 1; c < P->n_colours; ++c) {
    const int64_t i0 = P->cstart[c], i1 = P->cstart[c + 1];
    if (wide) k_ilu_sweep<true, true><<<g256((i1 - i0) * 16), 256, 0, s>>>(i0, i1, c, P->d_orec, P->d_nbc, ctx->d_pairs, P->d_lu, P->d_dinv, d_z);
    else k_ilu_sweep<true, false><<<g256((i1 - i0) * 16), 256, 0, s>>>(i0, i1, c, P->d_orec, P->d_nbc, ctx->d_pairs, P->d_lu, P->d_dinv, d_z);
  }
  for (int c = P->n_colours - 1; c >= 0; --c) {
    const int64_t i0 = P->cstart[c], i1 = P->cstart[c + 1];
    if (wide) k_ilu_sweep<false, true><<<g256((i1 - i0) * 16), 256, 0, s>>>(i0, i1, c, P->d_orec, P->d_nbc, ctx->d_pairs, P->d_lu, P->d_dinv, d_z);
    else k_ilu_sweep<false, false><<<g256((i1 - i0) * 16), 256, 0, s>>>(i0, i1, c, P->d_orec, P->d_nbc, ctx->d_pairs, P->d_lu, P->d_dinv, d_z);
  }
  ctx->launches += 2 * P->n_colours - 1;
  NS_CUDA(ctx, cudaGetLastError());
  return NSGPU_OK;
}

int ilu_colours(nsgpu_ctx* ctx, int32_t* h_colour, int32_t* n_colours) {
  IluPlan* P = static_cast<IluPlan*>(ctx->ilu);
  if (!P || P->unsupported || !P->d_colour) { set_error(ctx, "ilu_colours: no factorisation"); return NSGPU_EINVAL; }
  if (n_colours) *n_colours = P->n_colours;
  if (h_colour) {
    NS_CUDA(ctx, cudaMemcpyAsync(h_colour, P->d_colour, sizeof(int32_t) * (size_t)P->nv, cudaMemcpyDeviceToHost, ctx->stream));
    NS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return NSGPU_OK;
}

}  // namespace nsgpu
