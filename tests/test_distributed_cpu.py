"""world_size-2 (and 3) gloo tests of the multi-rank host logic on CPU: slab partition, pattern shipment
(SparsityPattern.finalize), column ghosts, vector halo and ghost-row exchange plans.  The device is replaced
by an oracle-backed pattern/assembly provider; the plans are executed with torch.distributed send/recv and the
result is compared with the serial oracle on the unpartitioned mesh."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleProvider:
    """Stand-in for NSAssembler: builds local patterns with NumPy and serves rows of them."""

    def __init__(self, part):
        self.part = part
        self.n_dofs = part.n_owned + part.n_ghost

    def build_pattern(self, extra_rows=None, extra_cols=None, colx_leader=None, colx_slot=None, colx_size=None):
        from oracle import oracle
        nx = 0 if colx_leader is None else len(colx_leader)
        self.n_cols = self.n_dofs + nx
        dm = self.part.dofmap[: self.part.n_cells_owned].astype(np.int64)
        nd = dm.shape[1]
        rows = np.repeat(dm, nd, axis=1).ravel()
        cols = np.tile(dm, (1, nd)).ravel()
        if extra_rows is not None and len(extra_rows):
            rows = np.concatenate([rows, np.asarray(extra_rows, dtype=np.int64)])
            cols = np.concatenate([cols, np.asarray(extra_cols, dtype=np.int64)])
        keys = np.unique(rows * self.n_cols + cols)
        r, c = keys // self.n_cols, keys % self.n_cols
        indptr = np.zeros(self.n_dofs + 1, dtype=np.int64)
        np.add.at(indptr, r + 1, 1)
        self.indptr, self.indices = np.cumsum(indptr), c.astype(np.int32)

    def get_rows(self, rows):
        rows = np.asarray(rows, dtype=np.int64)
        start = self.indptr[rows]
        lens = self.indptr[rows + 1] - start
        ptr = np.concatenate([[0], np.cumsum(lens)])
        idx = np.concatenate([self.indices[s:s + n] for s, n in zip(start, lens)]) if len(rows) else np.zeros(0, np.int32)
        return start, ptr, idx.astype(np.int32)


def _exchange_values(comm, neigh, sptr, sbuf, rptr):
    """what csrc/halo.cu does with grouped ncclSend/ncclRecv, on gloo"""
    out = np.zeros(int(rptr[-1]))
    reqs = []
    for k, o in enumerate(neigh):
        ns, nr = int(sptr[k + 1] - sptr[k]), int(rptr[k + 1] - rptr[k])
        if ns:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(sbuf[sptr[k]:sptr[k + 1]])), dst=int(o)))
    for k, o in enumerate(neigh):
        nr = int(rptr[k + 1] - rptr[k])
        if nr:
            t = torch.zeros(nr, dtype=torch.float64)
            dist.recv(t, src=int(o))
            out[rptr[k]:rptr[k + 1]] = t.numpy()
    for r in reqs:
        r.wait()
    return out


def _worker(rank, size, port, n_cross, n_long, errs):
    try:
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        from oracle import oracle
        from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
        from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
        comm = D.Comm.from_env("gloo")
        part = D.duct_partition(n_cross, n_long, rank, size)
        prov = OracleProvider(part)
        plans = D.build_plans(part, comm, prov)
        n_owned, n_dofs = part.n_owned, part.n_owned + part.n_ghost
        form = oracle.Form(0, 3, 1, 0.1)

        # serial reference on the whole duct
        m = M.duct_mesh(n_cross, n_long); sp = M.mixed_space(m, 1)
        w, bcs = M.duct_state(sp), M.duct_bcs(sp)
        marker, value, mult = oracle.bc_arrays(sp.n_dofs, [b[0] for b in bcs], [b[1] for b in bcs])
        gp, gi = oracle.build_pattern(sp.dofmap, sp.n_dofs)
        gv = oracle.assemble_jacobian(form, m.x, m.cells, sp.dofmap, w, gp, gi, marker, mult)
        gF = oracle.assemble_residual(form, m.x, m.cells, sp.dofmap, w, marker, value)
        oracle.set_bc(gF, [b[0] for b in bcs], [b[1] for b in bcs], w)
        import scipy.sparse as sps
        A = sps.csr_matrix((gv, gi, gp), shape=(sp.n_dofs,) * 2)

        # the partition reproduces the global state / BCs on its dofs
        l2g = part.local_to_global
        np.testing.assert_allclose(part.w, w[l2g], rtol=0, atol=1e-15)

        # forward halo: owners' values reach the ghosts (incl. the column ghosts)
        neigh, sptr, sidx, rptr, ridx = plans.halo
        xloc = np.zeros(plans.n_cols)
        xloc[:n_owned] = w[l2g[:n_owned]]
        got = _exchange_values(comm, neigh, sptr, xloc[sidx], rptr)
        xloc[ridx] = got
        colg = np.concatenate([l2g, plans.col_ghost_global])
        np.testing.assert_array_equal(xloc, w[colg])

        # local assembly (owned cells, rows = owned + ghost), then J.assemble(): ghost rows -> owners
        lmarker, lvalue, lmult = oracle.bc_arrays(n_dofs, [b[0] for b in part.bcs], [b[1] for b in part.bcs])
        lv = oracle.assemble_jacobian(form, part.x, part.cells, part.dofmap, xloc[:n_dofs], prov.indptr, prov.indices, lmarker, None,
                                      n_owned=n_owned, n_cells_owned=part.n_cells_owned)
        rn, sp_ptr, sp_pos, rp_ptr, rp_pos = plans.rows
        got = _exchange_values(comm, rn, sp_ptr, lv[sp_pos], rp_ptr)
        np.add.at(lv, rp_pos, got)
        diag_rows = np.nonzero(lmult[:n_owned] > 0)[0]
        for d in diag_rows:                                   # assemble_matrix's BC diagonal, owned dofs only
            seg = prov.indices[prov.indptr[d]:prov.indptr[d + 1]]
            lv[prov.indptr[d] + np.searchsorted(seg, d)] += lmult[d]
        # compare the owned rows with the serial matrix (pattern bit-exact after mapping columns to global)
        for r in range(0, n_owned, max(1, n_owned // 400)):
            seg = slice(prov.indptr[r], prov.indptr[r + 1])
            gcols = colg[prov.indices[seg]]
            ref = A.getrow(int(l2g[r]))
            order = np.argsort(gcols)
            np.testing.assert_array_equal(gcols[order], ref.indices)
            np.testing.assert_allclose(lv[seg][order], ref.data, rtol=0, atol=1e-13 * np.abs(gv).max())

        # residual: reverse halo-add, then set_bc on owned dofs
        lF = oracle.assemble_residual(form, part.x, part.cells, part.dofmap, xloc[:n_dofs], lmarker, lvalue, n_cells_owned=part.n_cells_owned)
        lFx = np.zeros(plans.n_cols); lFx[:n_dofs] = lF
        got = _exchange_values(comm, neigh, rptr, lFx[ridx], sptr)      # roles swapped
        np.add.at(lFx, sidx, got)
        oracle.set_bc(lFx, [b[0] for b in part.bcs], [b[1] for b in part.bcs], xloc, n_owned=n_owned)
        np.testing.assert_allclose(lFx[:n_owned], gF[l2g[:n_owned]], rtol=0, atol=1e-13 * np.abs(gF).max())

        # MatMult with the extended column space
        yv = np.array([lv[prov.indptr[r]:prov.indptr[r + 1]] @ xloc[prov.indices[prov.indptr[r]:prov.indptr[r + 1]]] for r in range(n_owned)])
        np.testing.assert_allclose(yv, (A @ w)[l2g[:n_owned]], rtol=0, atol=1e-12 * np.abs(A @ w).max())
        comm.close()
    except Exception:
        import traceback
        errs.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("size,n_cross,n_long", [(2, 3, 4), (3, 2, 7)])
def test_slab_partition_plans_reproduce_serial_assembly(size, n_cross, n_long):
    from oracle import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    errs = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + size
    procs = [ctx.Process(target=_worker, args=(r, size, port, n_cross, n_long, errs)) for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errs.empty():
        msgs.append(errs.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append((-1, "worker hung"))
    assert not msgs, "\n".join(f"rank {r}:\n{m}" for r, m in msgs)
    assert all(p.exitcode == 0 for p in procs)


def test_slab_ranges_balanced():
    from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
    for n, s in ((512, 8), (7, 3), (40, 6)):
        r = D.slab_ranges(n, s)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
