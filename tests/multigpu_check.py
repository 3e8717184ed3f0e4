"""Launched by torchrun (one rank per GPU): distributed GPU assembly / residual / MatMult on the slab-partitioned
duct against the serial CPU oracle.  Exit code 0 = parity within 1e-12 (relative to the largest entry)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _local_perm(n_owned, n_ghost, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.permutation(n_owned), n_owned + rng.permutation(n_ghost)]).astype(np.int64)   # old local -> new local


def _permuted(part, parts_of_all_ranks):
    """The same partition behind a dolfinx-like numbering in which the dofs of a vertex are scattered: on every rank the owned
    dofs are permuted among the owned and the ghosts among the ghosts, and -- as in a dolfinx index map -- the GLOBAL index of
    an owned dof is (the rank's offset + its local index), so the global numbering changes with it.  Returns the new
    partition and, per new local dof, its global index in the ORIGINAL numbering (for the comparison with the serial oracle)."""
    import copy
    new_global = {}
    for q, pq in enumerate(parts_of_all_ranks):                     # every rank can reproduce every rank's permutation
        perm_q = _local_perm(pq.n_owned, pq.n_ghost, 100 + q)
        new_global[q] = (pq.local_to_global[: pq.n_owned], pq.global_offset + perm_q[: pq.n_owned])
    n_total = sum(len(v[0]) for v in new_global.values())
    old2new = np.empty(n_total, dtype=np.int64)
    for og, ng in new_global.values():
        old2new[og] = ng
    n_owned, n_ghost = part.n_owned, part.n_ghost
    perm = _local_perm(n_owned, n_ghost, 100 + part.rank)
    p = copy.copy(part)
    p.dofmap = perm[part.dofmap].astype(np.int32)
    l2g_old = np.empty_like(part.local_to_global); l2g_old[perm] = part.local_to_global
    p.local_to_global = old2new[l2g_old]
    p.ghost_global = p.local_to_global[n_owned:]
    go = np.empty_like(part.ghost_owner); go[perm[n_owned:] - n_owned] = part.ghost_owner
    p.ghost_owner = go
    w = np.empty_like(part.w); w[perm] = part.w
    p.w = w
    p.bcs = [(perm[np.asarray(d)].astype(np.int32), v) for d, v in part.bcs]
    return p, l2g_old, old2new


def main():
    from oracle import oracle
    from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
    from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
    from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler
    n_cross, n_long = 6, 20
    comm = D.Comm.from_env()
    rank, size = comm.rank, comm.size
    part0 = D.duct_partition(n_cross, n_long, rank, size)
    # (kernel, overlap, permuted): auto kernel with the overlapped exchanges, the same with serial exchanges, the generic
    # kernel, and the auto kernel behind a dolfinx-like numbering (dofs of a vertex scattered: the library renumbers internally)
    for kernel, overlap, permuted in ((0, 1, False), (0, 0, False), (1, 1, False), (0, 1, True)):
        if permuted:
            part, l2g_ser, old2new = _permuted(part0, [D.duct_partition(n_cross, n_long, q, size) for q in range(size)])
        else:
            part, l2g_ser, old2new = part0, part0.local_to_global, None
        asm = NSAssembler(part.x, part.cells, part.dofmap, vdeg=1, n_dofs_owned=part.n_owned, n_dofs_ghost=part.n_ghost,
                          n_cells_owned=part.n_cells_owned, device=int(os.environ.get("LOCAL_RANK", 0)))
        asm.set_form(flavour=0, nu=0.1)
        asm.set_bcs(part.bcs)
        asm.set_option("kernel", kernel)
        asm.set_option("overlap", overlap)
        D.attach(asm, part, comm)
        plans = D.finish_pattern_exchange(asm, part, comm)
        n_owned, n_dofs = part.n_owned, part.n_owned + part.n_ghost
        # serial oracle
        m = M.duct_mesh(n_cross, n_long); sp = M.mixed_space(m, 1)
        w, bcs = M.duct_state(sp), M.duct_bcs(sp)
        marker, value, mult = oracle.bc_arrays(sp.n_dofs, [b[0] for b in bcs], [b[1] for b in bcs])
        gp, gi = oracle.build_pattern(sp.dofmap, sp.n_dofs)
        form = oracle.Form(0, 3, 1, 0.1)
        gv = oracle.assemble_jacobian(form, m.x, m.cells, sp.dofmap, w, gp, gi, marker, mult)
        gF = oracle.assemble_residual(form, m.x, m.cells, sp.dofmap, w, marker, value)
        oracle.set_bc(gF, [b[0] for b in bcs], [b[1] for b in bcs], w)
        import scipy.sparse as sps
        A = sps.csr_matrix((gv, gi, gp), shape=(sp.n_dofs,) * 2)
        l2g = l2g_ser                                                # local dof -> global dof of the serial oracle
        colg_new = np.concatenate([part.local_to_global, plans.col_ghost_global]) if plans is not None else part.local_to_global
        if old2new is None:
            colg = colg_new
        else:                                                        # column ghosts arrive in the renumbered global numbering
            new2old = np.empty_like(old2new); new2old[old2new] = np.arange(len(old2new))
            colg = new2old[colg_new]
        # ghosts deliberately wrong on input: the forward halo must refresh them
        xin = part.w.copy()
        xin[n_owned:] = 1e30
        asm.set_values(np.full(asm.nnz, 1e30))        # poison: every entry must be written or explicitly zeroed by the assembly
        vals, F = asm.jacobian_residual(xin)
        ip = np.empty(n_dofs + 1, dtype=np.int64); ix = np.empty(asm.nnz, dtype=np.int32)
        asm._check(asm.lib.nsgpu_get_pattern(asm.ctx, ip.ctypes.data, ix.ctypes.data), "get_pattern")
        worst = 0.0
        for r in range(n_owned):
            seg = slice(ip[r], ip[r + 1])
            gcols = colg[ix[seg]]
            ref = A.getrow(int(l2g[r]))
            order = np.argsort(gcols)
            assert np.array_equal(gcols[order], ref.indices), f"rank {rank}: pattern of row {r} differs"
            worst = max(worst, np.abs(vals[seg][order] - ref.data).max())
        assert worst <= 1e-12 * np.abs(gv).max(), f"rank {rank} kernel {kernel}: J err {worst}"
        eF = np.abs(F[:n_owned] - gF[l2g[:n_owned]]).max()
        assert eF <= 1e-12 * np.abs(gF).max(), f"rank {rank} kernel {kernel}: F err {eF}"
        xv = np.random.default_rng(5).standard_normal(sp.n_dofs)
        xl = np.zeros(n_dofs); xl[:n_owned] = xv[l2g[:n_owned]]
        y = asm.mult(xl)
        ey = np.abs(y - (A @ xv)[l2g[:n_owned]]).max()
        assert ey <= 1e-12 * np.abs(A @ xv).max(), f"rank {rank} kernel {kernel}: MatMult err {ey}"
        # KSPTFQMR on the distributed Jacobian (dot products over ncclAllReduce) against a serial direct solve
        import scipy.sparse.linalg as spla
        x_ref = spla.spsolve(A.tocsc(), gF)
        xs, info = asm.tfqmr(F[:n_owned], rtol=1e-12, max_it=4000, pc=4)
        ex = np.abs(xs - x_ref[l2g[:n_owned]]).max()
        assert ex <= 1e-8 * np.abs(x_ref).max(), f"rank {rank} kernel {kernel}: TFQMR err {ex} {info}"
        ilu_its = -1
        if kernel == 0:     # the same solve with the multicolour block ILU(0) of each rank's diagonal block (block Jacobi over the ranks)
            xi, info_i = asm.tfqmr(F[:n_owned], rtol=1e-12, max_it=4000, pc=5)
            exi = np.abs(xi - x_ref[l2g[:n_owned]]).max()
            assert exi <= 1e-8 * np.abs(x_ref).max(), f"rank {rank}: TFQMR + ILU err {exi} {info_i}"
            ilu_its = info_i["its"]
        asm.set_option("stream_host", 0)
        asm.jacobian_residual(xin)
        kname = asm.last_kernel_name()
        if kernel == 0:
            assert kname == "p1tet_ws" and asm.last_spmv_name() == "spmv_block4", (kname, asm.last_spmv_name())
        print(f"rank {rank}/{size} kernel {kernel} ({kname}, {asm.last_spmv_name()}) overlap {overlap} permuted {int(permuted)}: "
              f"J {worst:.2e} F {eF:.2e} Jx {ey:.2e} tfqmr {ex:.2e} in {info['its']} its (ILU: {ilu_its}) "
              f"(n_owned {n_owned}, ghosts {part.n_ghost}, col ghosts {asm.n_cols - n_dofs})", flush=True)
        asm.close()
    comm.close()


if __name__ == "__main__":
    main()
