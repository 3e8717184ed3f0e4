"""TEST INFRASTRUCTURE ONLY -- textbook block ILU(0) (Saad, Iterative Methods for Sparse Linear Systems, alg. 10.4, on 4x4 blocks) in a
given elimination order, NumPy.  It is the checker for csrc/ilu.cu: PETSc's default preconditioner for the reference's
snes_ksp_type = 'tfqmr' (NavierStokes/NavierStokesChannelFlow.py:282-291) is ILU(0); the library factorises the 4x4 vertex blocks in
a multicolour order, and this file restates exactly that factorisation for a vertex-blocked CSR matrix (dof = 4 vertex + component)."""
import numpy as np


def block_ilu0(indptr, indices, vals, colour):
    nv = len(colour)
    rank = np.empty(nv, dtype=np.int64)                      # position in the elimination order (colour, vertex)
    order = np.lexsort((np.arange(nv), colour))
    rank[order] = np.arange(nv)
    blocks = [dict() for _ in range(nv)]                     # blocks[i][j] = 4x4
    for i in range(nv):
        for r in range(4):
            row = 4 * i + r
            for p in range(indptr[row], indptr[row + 1]):
                c = indices[p]
                if c >= 4 * nv:
                    continue                                 # other ranks' columns are dropped (block Jacobi over the ranks)
                j = c // 4
                blocks[i].setdefault(j, np.zeros((4, 4)))[r, c % 4] = vals[p]
    dinv = [None] * nv
    for i in order:
        Bi = blocks[i]
        for k in sorted((k for k in Bi if rank[k] < rank[i]), key=lambda k: rank[k]):
            L = Bi[k] @ dinv[k]
            Bi[k] = L
            Bk = blocks[k]
            for j in Bi:
                if j != k and rank[j] > rank[k] and j in Bk:
                    Bi[j] = Bi[j] - L @ Bk[j]
        dinv[i] = np.linalg.inv(Bi[i])
    return blocks, dinv, order, rank


def apply(blocks, dinv, order, rank, r):
    z = np.array(r, dtype=np.float64).reshape(-1, 4).copy()
    for i in order:
        for k, L in blocks[i].items():
            if rank[k] < rank[i]:
                z[i] -= L @ z[k]
    for i in order[::-1]:
        acc = z[i].copy()
        for j, U in blocks[i].items():
            if rank[j] > rank[i]:
                acc -= U @ z[j]
        z[i] = dinv[i] @ acc
    return z.reshape(-1)
