"""Synthetic meshes, dof maps, states and boundary conditions in dolfinx array layout (host, NumPy).

The reference gets these arrays from dolfinx (``mesh.geometry.x``, ``mesh.geometry.dofmap``,
``W.dofmap.list``, ``locate_dofs_topological``; NavierStokes/NavierStokesChannelFlow.py:107-147,
LidDrivenFlow/LidDrivenNavierStokesFlow.py:29-77).  dolfinx is not part of the hot path being
replaced, and is not installable here, so for the BASELINE configurations that are *synthetic by
definition* (structured-tet ducts, the create_rectangle cavity) this module emits arrays with the same
layout contract:

* geometry ``x``: (n_nodes, 3) float64, 3-padded; ``cells``: (n_cells, gdim+1) int32;
* ``dofmap``: (n_cells, ndofs_cell) int32, block size 1, cell-local order = velocity node-major with
  interleaved components, then pressure nodes (``mixed_element([P_k^gdim, P1])``, :128);
* Dirichlet conditions as an ordered list of (dofs, values) pairs, one per ``dirichletbc`` object,
  overlaps preserved (a dof on the wall/inlet rim is held by two objects, :134-140).

Global dof numbering of the synthetic spaces: all dofs of a mesh entity are contiguous
(vertex v -> (gdim+1)*v + [u_0..u_{gdim-1}, p]; P2 edge e -> (gdim+1)*Nv + gdim*e + c).
"""
from dataclasses import dataclass, field

import numpy as np

# dolfinx create_box: each box -> 6 tets sharing the v0-v7 diagonal; v_k lexicographic, first axis fastest
_BOX_TETS = np.array([[0, 1, 3, 7], [0, 1, 7, 5], [0, 5, 7, 4], [0, 3, 2, 7], [0, 6, 4, 7], [0, 2, 6, 7]], dtype=np.int64)
# basix reference-cell edge -> vertex pairs (P2 edge dof order)
TET_EDGES = np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]])
TRI_EDGES = np.array([[1, 2], [0, 2], [0, 1]])


@dataclass
class Mesh:
    gdim: int
    x: np.ndarray            # (n_nodes, 3) float64
    cells: np.ndarray        # (n_cells, gdim+1) int32
    shape: tuple = ()        # structured box counts along the logical axes (fastest first)
    meta: dict = field(default_factory=dict)

    @property
    def n_cells(self):
        return self.cells.shape[0]

    @property
    def n_vertices(self):
        return self.x.shape[0]


@dataclass
class Space:
    """Mixed P_k^gdim x P1 space in dolfinx layout (bs = 1)."""
    mesh: Mesh
    vdeg: int
    dofmap: np.ndarray       # (n_cells, ndofs_cell) int32
    n_dofs: int
    dof_x: np.ndarray        # (n_dofs, 3) coordinates of each dof's node
    dof_comp: np.ndarray     # (n_dofs,) int8: 0..gdim-1 velocity component, gdim = pressure
    edges: np.ndarray = None  # (n_edges, 2) vertex pairs for P2

    @property
    def ndofs_cell(self):
        return self.dofmap.shape[1]


def create_box_tets(n, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0), axes=(0, 1, 2)):
    """Structured tet mesh: n[0] x n[1] x n[2] boxes along logical axes (first fastest), 6 tets per box,
    cell order box-major.  ``axes[k]`` is the physical coordinate the k-th logical axis maps to."""
    n0, n1, n2 = (int(v) for v in n)
    s0, s1 = n0 + 1, (n0 + 1) * (n1 + 1)
    nv = s1 * (n2 + 1)
    i0, i1, i2 = np.meshgrid(np.arange(n0 + 1), np.arange(n1 + 1), np.arange(n2 + 1), indexing="ij")
    idx = (i0 + s0 * i1 + s1 * i2).ravel()
    x = np.zeros((nv, 3))
    for k, (ik, nk) in enumerate(((i0, n0), (i1, n1), (i2, n2))):
        ax = axes[k]
        x[idx, ax] = p0[ax] + (p1[ax] - p0[ax]) * ik.ravel() / nk
    b0, b1, b2 = np.meshgrid(np.arange(n0), np.arange(n1), np.arange(n2), indexing="ij")
    # box-major order with the first logical axis fastest
    order = np.argsort((b0 + n0 * (b1 + n1 * b2)).ravel(), kind="stable")
    base = (b0 + s0 * b1 + s1 * b2).ravel()[order].astype(np.int32)
    corner = np.array([(k & 1) + s0 * ((k >> 1) & 1) + s1 * ((k >> 2) & 1) for k in range(8)], dtype=np.int32)
    cells = (base[:, None, None] + corner[_BOX_TETS][None, :, :]).reshape(-1, 4)
    return Mesh(3, x, np.ascontiguousarray(cells, dtype=np.int32), (n0, n1, n2), {"kind": "box_tets", "axes": tuple(axes)})


def duct_mesh(n_cross, n_long, length=4.0):
    """BASELINE synthetic duct [0,length] x [-1/2,1/2]^2: n_cross x n_cross x n_long boxes, duct axis
    (physical x) = slowest logical axis so that x-slabs are contiguous index ranges."""
    m = create_box_tets((n_cross, n_cross, n_long), p0=(0.0, -0.5, -0.5), p1=(length, 0.5, 0.5), axes=(1, 2, 0))
    m.meta.update(kind="duct", n_cross=n_cross, n_long=n_long, length=length)
    return m


def create_rectangle_tris(nx, ny, p0=(0.0, 0.0), p1=(1.0, 1.0)):
    """dolfinx create_rectangle(..., CellType.triangle), right diagonal: each square -> 2 triangles
    sharing the (v0, v3) diagonal (LidDrivenNavierStokesFlow.py:29-30)."""
    s0 = nx + 1
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    idx = (ix + s0 * iy).ravel()
    x = np.zeros(((nx + 1) * (ny + 1), 3))
    x[idx, 0] = p0[0] + (p1[0] - p0[0]) * ix.ravel() / nx
    x[idx, 1] = p0[1] + (p1[1] - p0[1]) * iy.ravel() / ny
    bx, by = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    order = np.argsort((bx + nx * by).ravel(), kind="stable")
    v0 = (bx + s0 * by).ravel()[order]
    v1, v2, v3 = v0 + 1, v0 + s0, v0 + s0 + 1
    cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], 1).reshape(-1, 3)
    return Mesh(2, x, cells.astype(np.int32), (nx, ny), {"kind": "rectangle_tris"})


def mixed_space(mesh, vdeg=1):
    """Mixed P_vdeg^gdim x P1 dof map, dolfinx cell-local ordering, entity-contiguous global numbering."""
    gd = mesh.gdim
    nv = mesh.n_vertices
    nc = mesh.cells.shape[0]
    bs = gd + 1
    if vdeg == 1:
        nvn = gd + 1
        dm = np.empty((nc, gd * nvn + gd + 1), dtype=np.int32)
        for n in range(nvn):
            base = bs * mesh.cells[:, n]
            for c in range(gd):
                dm[:, gd * n + c] = base + c
            dm[:, gd * nvn + n] = base + gd
        n_dofs = bs * nv
        dof_x = np.repeat(mesh.x, bs, axis=0)
        dof_comp = np.tile(np.arange(bs, dtype=np.int8), nv)
        return Space(mesh, 1, dm, n_dofs, dof_x, dof_comp)
    cells = mesh.cells.astype(np.int64)
    ledges = TET_EDGES if gd == 3 else TRI_EDGES
    ne_l = len(ledges)
    pairs = np.sort(cells[:, ledges], axis=2)                      # (nc, ne_l, 2)
    keys = pairs[..., 0] * nv + pairs[..., 1]
    ukeys, inv = np.unique(keys.ravel(), return_inverse=True)
    eid = inv.reshape(nc, ne_l)
    edges = np.stack([ukeys // nv, ukeys % nv], 1)
    n_edges = len(ukeys)
    nvn = gd + 1 + ne_l
    dm = np.empty((nc, gd * nvn + gd + 1), dtype=np.int64)
    for n in range(gd + 1):
        for c in range(gd):
            dm[:, gd * n + c] = bs * cells[:, n] + c
        dm[:, gd * nvn + n] = bs * cells[:, n] + gd
    for e in range(ne_l):
        for c in range(gd):
            dm[:, gd * (gd + 1 + e) + c] = bs * nv + gd * eid[:, e] + c
    n_dofs = bs * nv + gd * n_edges
    dof_x = np.concatenate([np.repeat(mesh.x, bs, axis=0), np.repeat(0.5 * (mesh.x[edges[:, 0]] + mesh.x[edges[:, 1]]), gd, axis=0)])
    dof_comp = np.concatenate([np.tile(np.arange(bs, dtype=np.int8), nv), np.tile(np.arange(gd, dtype=np.int8), n_edges)])
    return Space(mesh, 2, dm.astype(np.int32), n_dofs, dof_x, dof_comp, edges)


# ----------------------------------------------------------------------------------------- states / BCs
def duct_state(space, seed=1234, noise=1e-3):
    """BASELINE.md section 3 state: u = (1.5(1-4y^2)(1-4z^2)(1+0.1 sin 2 pi x), 0.05 sin 2 pi y,
    0.05 sin 2 pi z), p = 4 - x, plus noise * N(0,1) from default_rng(seed)."""
    X, c = space.dof_x, space.dof_comp
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    w = np.where(c == 0, 1.5 * (1 - 4 * y * y) * (1 - 4 * z * z) * (1 + 0.1 * np.sin(2 * np.pi * x)),
        np.where(c == 1, 0.05 * np.sin(2 * np.pi * y),
        np.where(c == 2, 0.05 * np.sin(2 * np.pi * z), 4.0 - x)))
    return w + noise * np.random.default_rng(seed).standard_normal(space.n_dofs)


def duct_bcs(space, length=4.0, tol=1e-12):
    """[wall, inlet, outlet] mirroring NavierStokesChannelFlow.py:127-147: no-slip on the four side
    walls, Dirichlet inlet profile on x = 0 (rim vertices shared with the wall object), p = 0 on
    x = length.  Returns a list of (dofs int32, values float64)."""
    X, c = space.dof_x, space.dof_comp
    gd = space.mesh.gdim
    vel = c < gd
    on_wall = (np.abs(np.abs(X[:, 1]) - 0.5) < tol) | (np.abs(np.abs(X[:, 2]) - 0.5) < tol)
    wall = np.nonzero(vel & on_wall)[0]
    inlet = np.nonzero(vel & (np.abs(X[:, 0]) < tol))[0]
    g_in = np.where(c[inlet] == 0, 1.5 * (1 - 4 * X[inlet, 1] ** 2) * (1 - 4 * X[inlet, 2] ** 2), 0.0)
    outlet = np.nonzero((c == gd) & (np.abs(X[:, 0] - length) < tol))[0]
    return [(wall.astype(np.int32), np.zeros(len(wall))), (inlet.astype(np.int32), g_in),
            (outlet.astype(np.int32), np.zeros(len(outlet)))]


def cavity_bcs(space, tol=1e-12):
    """[noslip, lid, pressure pin] of LidDrivenNavierStokesFlow.py:57-77 on the unit square."""
    X, c = space.dof_x, space.dof_comp
    vel = c < 2
    noslip = np.nonzero(vel & ((np.abs(X[:, 0]) < tol) | (np.abs(X[:, 0] - 1) < tol) | (np.abs(X[:, 1]) < tol)))[0]
    lid = np.nonzero(vel & (np.abs(X[:, 1] - 1) < tol))[0]
    g_lid = np.where(c[lid] == 0, 1.0, 0.0)
    pin = np.nonzero((c == 2) & (np.abs(X[:, 0]) < tol) & (np.abs(X[:, 1]) < tol))[0]
    return [(noslip.astype(np.int32), np.zeros(len(noslip))), (lid.astype(np.int32), g_lid),
            (pin.astype(np.int32), np.zeros(len(pin)))]


def cavity_state(space, seed=1234, noise=1e-3):
    """Smooth recirculating field + noise for the lid-driven cavity parity cases."""
    X, c = space.dof_x, space.dof_comp
    x, y = X[:, 0], X[:, 1]
    w = np.where(c == 0, np.sin(np.pi * x) ** 2 * np.sin(2 * np.pi * y) * y,
        np.where(c == 1, -np.sin(2 * np.pi * x) * np.sin(np.pi * y) ** 2 * 0.5, 0.1 * np.cos(np.pi * x) * y))
    return w + noise * np.random.default_rng(seed).standard_normal(space.n_dofs)


# ----------------------------------------------------------------------------------------- DFG 2D-1 (Schaefer-Turek) domain
def dfg_cylinder_mesh(h_far=0.02, n_cyl=96, growth=1.18):
    """Triangulation of the DFG 2D-1 benchmark domain of NavierStokes/Validation_Flow (dfg_pillar_2D.geo: channel
    [0, 2.2] x [0, 0.41], cylinder of radius 0.05 at (0.2, 0.2)) without gmsh: points on the cylinder, on concentric rings
    that grow geometrically away from it, and on a regular background lattice, triangulated with scipy's Delaunay; the
    triangles inside the cylinder are dropped.  Cells are counter-clockwise."""
    from scipy.spatial import Delaunay
    L, H, cx, cy, r = 2.2, 0.41, 0.2, 0.2, 0.05
    pts = []
    th = 2 * np.pi * np.arange(n_cyl) / n_cyl
    pts.append(np.stack([cx + r * np.cos(th), cy + r * np.sin(th)], 1))
    ring_r, dr, k = r, 2 * np.pi * r / n_cyl, 0
    while True:
        dr = min(dr * growth, h_far)
        ring_r += dr
        if ring_r + 0.6 * h_far > min(cx, cy, H - cy):
            break
        k += 1
        n = max(12, int(round(2 * np.pi * ring_r / dr)))
        t = 2 * np.pi * (np.arange(n) + 0.5 * (k % 2)) / n
        pts.append(np.stack([cx + ring_r * np.cos(t), cy + ring_r * np.sin(t)], 1))
    r_out = ring_r - dr
    nx, ny = int(round(L / h_far)), int(round(H / h_far))
    gx, gy = np.meshgrid(np.linspace(0, L, nx + 1), np.linspace(0, H, ny + 1), indexing="ij")
    gx = gx.copy()
    gx[:, 1::2][1:-1] += 0.5 * L / nx                                    # staggered rows: near-equilateral triangles
    g = np.stack([gx.ravel(), gy.ravel()], 1)
    g = g[(g[:, 0] <= L + 1e-12)]
    keep = np.hypot(g[:, 0] - cx, g[:, 1] - cy) > r_out + 0.7 * h_far
    pts.append(g[keep])
    P = np.concatenate(pts)
    tri = Delaunay(P)
    T = tri.simplices
    cen = P[T].mean(axis=1)
    inside = (np.hypot(cen[:, 0] - cx, cen[:, 1] - cy) < r * np.cos(np.pi / n_cyl)) | (T < n_cyl).all(axis=1)   # the first n_cyl points are the cylinder
    T = T[~inside]
    a, b, c = P[T[:, 0]], P[T[:, 1]], P[T[:, 2]]
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    T = T[np.abs(area) > 1e-14]
    area = area[np.abs(area) > 1e-14]
    T[area < 0] = T[area < 0][:, [0, 2, 1]]
    used = np.unique(T)
    remap = -np.ones(len(P), dtype=np.int64); remap[used] = np.arange(len(used))
    x = np.zeros((len(used), 3)); x[:, :2] = P[used]
    return Mesh(2, x, remap[T].astype(np.int32), (), {"kind": "dfg2d", "L": L, "H": H, "cx": cx, "cy": cy, "r": r, "n_cyl": n_cyl})


def dfg_bcs(space, tol=1e-9):
    """bc = [bcu_inflow, bcu_walls, bcu_obstacle] of DFG_2D_Validation.py:64-97: parabolic inflow 4 * 0.3 * y (0.41 - y) / 0.41^2,
    no-slip walls and cylinder, do-nothing outflow.  Also returns the velocity dofs on the cylinder (for drag / lift)."""
    X, c = space.dof_x, space.dof_comp
    mt = space.mesh.meta
    vel = c < 2
    inflow = np.nonzero(vel & (np.abs(X[:, 0]) < tol))[0]
    g_in = np.where(c[inflow] == 0, 4.0 * 0.3 * X[inflow, 1] * (mt["H"] - X[inflow, 1]) / mt["H"] ** 2, 0.0)
    walls = np.nonzero(vel & ((np.abs(X[:, 1]) < tol) | (np.abs(X[:, 1] - mt["H"]) < tol)))[0]
    on_cyl = np.abs(np.hypot(X[:, 0] - mt["cx"], X[:, 1] - mt["cy"]) - mt["r"]) < 1e-6
    obstacle = np.nonzero(vel & on_cyl)[0]
    bcs = [(inflow.astype(np.int32), g_in), (walls.astype(np.int32), np.zeros(len(walls))), (obstacle.astype(np.int32), np.zeros(len(obstacle)))]
    return bcs, obstacle
