"""Standalone .msh ingestion: what gmshio.model_to_mesh + locate_dofs_topological give the drivers
(NavierStokes/NavierStokesChannelFlow.py:107-147), without gmsh or dolfinx."""
import numpy as np
import pytest

from stabilized_navier_stokes_flow_fenicsx_b200 import gmsh_io as G
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M

MARK = {"inlet_1": 1, "inlet_2": 2, "outlet": 3, "wall": 4}     # image2gmsh3D.py:435-438


def _tagged_duct(n_cross=3, n_long=5):
    m = M.duct_mesh(n_cross, n_long)
    bf = G.boundary_facets(m)
    c = m.x[bf].mean(axis=1)
    tags = np.full(len(bf), MARK["wall"], dtype=np.int32)
    inlet = np.abs(c[:, 0]) < 1e-12
    tags[inlet & (np.hypot(c[:, 1], c[:, 2]) < 0.25)] = MARK["inlet_1"]      # an inner and an outer inlet region, like the two-stream inlets
    tags[inlet & (np.hypot(c[:, 1], c[:, 2]) >= 0.25)] = MARK["inlet_2"]
    tags[np.abs(c[:, 0] - 4.0) < 1e-12] = MARK["outlet"]
    return m, bf, tags


def test_msh_round_trip_and_tagged_facets(tmp_path):
    m, bf, tags = _tagged_duct()
    names = {(2, v): k for k, v in MARK.items()}
    path = tmp_path / "ChannelMesh.msh"
    G.write_msh(path, m, bf, tags, names)
    m2, ctags, f2, ft2, names2 = G.read_msh(path)
    assert m2.gdim == 3 and m2.n_cells == m.n_cells and np.all(ctags == 1)
    np.testing.assert_array_equal(m2.cells, m.cells)
    np.testing.assert_allclose(m2.x, m.x, rtol=0, atol=0)
    assert names2[(2, 4)] == "wall" and names2[(3, 1)] == "fluid"
    key = lambda f, t: sorted((tuple(sorted(r)), int(v)) for r, v in zip(f.tolist(), t.tolist()))
    assert key(f2, ft2) == key(bf, tags)
    # every boundary facet is tagged exactly once, the tagged area is the duct's surface
    a, b, c = (m2.x[f2[:, k]] for k in range(3))
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    assert abs(area[ft2 == MARK["outlet"]].sum() - 1.0) < 1e-12 and abs(area.sum() - (2 * 1.0 + 4 * 4.0)) < 1e-12


@pytest.mark.parametrize("vdeg", [1, 2])
def test_facet_dofs_match_the_geometric_locators(tmp_path, vdeg):
    """locate_dofs_topological on the wall / outlet markers gives the same dof sets as mesh.duct_bcs finds geometrically."""
    m, bf, tags = _tagged_duct()
    path = tmp_path / "duct.msh"
    G.write_msh(path, m, bf, tags)
    m2, _, f2, ft2, _ = G.read_msh(path)
    sp = M.mixed_space(m2, vdeg)
    wall_geo, inlet_geo, outlet_geo = (b[0] for b in M.duct_bcs(sp))
    np.testing.assert_array_equal(G.facet_dofs(sp, f2, ft2, MARK["wall"]), np.sort(wall_geo))
    np.testing.assert_array_equal(G.facet_dofs(sp, f2, ft2, MARK["outlet"], sub="pressure"), np.sort(outlet_geo))
    inlet = np.union1d(G.facet_dofs(sp, f2, ft2, MARK["inlet_1"]), G.facet_dofs(sp, f2, ft2, MARK["inlet_2"]))
    np.testing.assert_array_equal(inlet, np.sort(inlet_geo))


def test_triangle_mesh_and_bad_files(tmp_path):
    m = M.create_rectangle_tris(4, 3)
    bf = G.boundary_facets(m)
    tags = np.where(np.abs(m.x[bf].mean(axis=1)[:, 1] - 1.0) < 1e-12, 2, 1).astype(np.int32)     # lid = 2, rest = 1
    path = tmp_path / "cavity.msh"
    G.write_msh(path, m, bf, tags)
    m2, _, f2, ft2, _ = G.read_msh(path)
    assert m2.gdim == 2 and m2.n_cells == 24 and (ft2 == 2).sum() == 4
    sp = M.mixed_space(m2, 1)
    lid = G.facet_dofs(sp, f2, ft2, 2)
    assert np.all(np.abs(sp.dof_x[lid, 1] - 1.0) < 1e-12) and len(lid) == 2 * 5
    bad = tmp_path / "old.msh"
    bad.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError, match="MSH 4"):
        G.read_msh(bad)
