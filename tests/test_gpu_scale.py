"""Oracle parity at the sizes the benchmark runs (-m gpu): every entry on M (3.0 M cells), and on L (50.3 M cells,
nnz = 94.5 % of 2^31, int64 row positions) the rows of > 1000 sampled vertices including the last ones of the matrix.
An index overflow at L, or a tile / gather-list error that only shows beyond the small meshes of test_gpu_parity.py,
fails here.  Tolerance 1e-12 relative to the largest magnitude of the compared array (BASELINE.json north_star)."""
import numpy as np
import pytest

from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

pytestmark = pytest.mark.gpu
RTOL = 1e-12
NU, CI = 0.1, 36.0


def test_every_entry_on_M_matches_the_oracle(oracle):
    part = D.duct_partition(50, 200, 0, 1)
    n = part.n_owned
    form = oracle.Form(gdim=3, vdeg=1, flavour=0, nu=NU, Ci=CI)
    marker, value, mult = oracle.bc_arrays(n, [b[0] for b in part.bcs], [np.broadcast_to(b[1], (len(b[0]),)) for b in part.bcs])
    indptr, indices = oracle.build_pattern_c(part.dofmap, n)
    vals = oracle.assemble_jacobian(form, part.x, part.cells, part.dofmap, part.w, indptr, indices, marker, mult)
    F = oracle.assemble_residual(form, part.x, part.cells, part.dofmap, part.w, marker, value)
    F[marker != 0] = part.w[marker != 0] - value[marker != 0]                      # set_bc(F, bc, x, -1.0)
    asm = NSAssembler(part.x, part.cells, part.dofmap, vdeg=1)
    asm.set_form(flavour=0, nu=NU, Ci=CI); asm.set_bcs(part.bcs)
    gp, gi = asm.create_matrix()
    np.testing.assert_array_equal(gp, indptr)
    np.testing.assert_array_equal(gi, indices)
    asm.set_option("stream_host", 0)
    gv, gF = asm.jacobian_residual(part.w)
    assert asm.last_kernel_name() == "p1tet_ws"
    assert np.abs(gv - vals).max() <= RTOL * np.abs(vals).max()
    assert np.abs(gF - F).max() <= RTOL * np.abs(F).max()
    xv = np.random.default_rng(2).standard_normal(n)
    y = asm.mult(xv)
    yo = oracle.spmv(indptr, indices, vals, xv)
    assert np.abs(y - yo).max() <= RTOL * np.abs(yo).max()
    asm.close()


def _incident_cells(verts, n_cross, n_long):
    """cells of the structured duct (box-major, first in-plane axis fastest, 6 tets per box) around the given vertices"""
    s0, s1 = n_cross + 1, (n_cross + 1) ** 2
    i0, i1, i2 = verts % s0, (verts // s0) % s0, verts // s1
    out = []
    for d0 in (-1, 0):
        for d1 in (-1, 0):
            for d2 in (-1, 0):
                b0, b1, b2 = i0 + d0, i1 + d1, i2 + d2
                ok = (b0 >= 0) & (b0 < n_cross) & (b1 >= 0) & (b1 < n_cross) & (b2 >= 0) & (b2 < n_long)
                box = (b0 + n_cross * (b1 + n_cross * b2))[ok]
                out.append((6 * box[:, None] + np.arange(6)[None, :]).ravel())
    return np.unique(np.concatenate(out))


def test_sampled_rows_on_L_match_the_oracle(oracle):
    n_cross, n_long = 128, 512
    part = D.duct_partition(n_cross, n_long, 0, 1)
    nv = part.x.shape[0]
    n = part.n_owned
    assert n == 4 * nv == 34147332
    rng = np.random.default_rng(77)
    verts = np.unique(np.concatenate([rng.integers(0, nv, 1100), np.arange(nv - 40, nv), np.arange(0, 10),        # last / first rows
                                      (n_cross + 1) ** 2 * rng.integers(0, n_long + 1, 40) + rng.integers(0, (n_cross + 1) ** 2, 40)]))
    cand = _incident_cells(verts, n_cross, n_long)
    cells_g = part.cells[cand]
    keep = np.isin(cells_g, verts).any(axis=1)                                       # only cells that touch a sampled vertex
    cells_g = cells_g[keep]
    uniq, inv = np.unique(cells_g, return_inverse=True)
    cells_s = inv.reshape(cells_g.shape).astype(np.int32)
    x_s = part.x[uniq]
    dofmap_s = np.concatenate([(4 * cells_s[:, :, None] + np.arange(3)[None, None, :]).reshape(len(cells_s), 12), 4 * cells_s + 3], axis=1).astype(np.int32)
    np.testing.assert_array_equal(part.dofmap[cand][keep][:, :12].reshape(-1, 4, 3)[:, :, 0], 4 * cells_g)     # the duct's numbering is 4 v + c
    gdofs = (4 * uniq[:, None].astype(np.int64) + np.arange(4)[None, :]).ravel()                              # sub-mesh dof -> global dof
    marker, value, mult = oracle.bc_arrays(n, [b[0] for b in part.bcs], [np.broadcast_to(b[1], (len(b[0]),)) for b in part.bcs])
    form = oracle.Form(gdim=3, vdeg=1, flavour=0, nu=NU, Ci=CI)
    w_s = part.w[gdofs]
    ip_s, ix_s = oracle.build_pattern(dofmap_s, len(gdofs))
    vals_s = oracle.assemble_jacobian(form, x_s, cells_s, dofmap_s, w_s, ip_s, ix_s, marker[gdofs], mult[gdofs])
    F_s = oracle.assemble_residual(form, x_s, cells_s, dofmap_s, w_s, marker[gdofs], value[gdofs])
    mk = marker[gdofs] != 0
    F_s[mk] = w_s[mk] - value[gdofs][mk]

    asm = NSAssembler(part.x, part.cells, part.dofmap, vdeg=1)
    asm.set_form(flavour=0, nu=NU, Ci=CI); asm.set_bcs(part.bcs)
    asm.create_matrix(fetch=False)
    assert asm.nnz == 2029817872
    x_dev, F_dev = asm.dev_alloc(8 * asm.n_cols), asm.dev_alloc(8 * asm.n_cols)
    asm.h2d(x_dev, part.w)
    asm.jacobian_residual_dev(x_dev, True, F_dev)
    asm.sync()
    assert asm.last_kernel_name() == "p1tet_ws"
    gF = np.empty(n); asm.d2h(gF, F_dev)
    vdev = asm.values_dev()
    lv = np.searchsorted(uniq, verts)                                               # sampled vertices in sub-mesh numbering
    rows_s = (4 * lv[:, None] + np.arange(4)[None, :]).ravel()
    rows_g = gdofs[rows_s].astype(np.int32)
    start, ptr, idx = asm.get_rows(rows_g)
    assert start.max() > 0.9 * 2 ** 31                                               # the sample reaches the end of the value array
    scale = np.abs(vals_s).max()
    worst = 0.0
    for k, (rs, rg) in enumerate(zip(rows_s, rows_g)):
        cols_o = gdofs[ix_s[ip_s[rs]: ip_s[rs + 1]]]
        np.testing.assert_array_equal(idx[ptr[k]: ptr[k + 1]], cols_o)              # pattern row, bit-exact (both sorted)
        gv = asm.get_value_range(start[k], ptr[k + 1] - ptr[k], vdev)
        worst = max(worst, np.abs(gv - vals_s[ip_s[rs]: ip_s[rs + 1]]).max())
    assert worst <= RTOL * scale
    assert np.abs(gF[rows_g] - F_s[rows_s]).max() <= RTOL * max(np.abs(F_s[rows_s]).max(), np.abs(gF).max())
    asm.close()
