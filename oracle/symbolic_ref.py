#!/usr/bin/env python
"""symbolic_ref.py -- independent symbolic (sympy) evaluation of the reference's weak forms.

TEST INFRASTRUCTURE (oracle side).  Generates the golden element tensors committed under
tests/golden/ that pin oracle/ns_oracle.c.  Nothing here is imported by the product.

Why: the reference's arithmetic is done by UFL + FFCx + basix (un-vendored, not installable
here), and the reference has no tests.  To avoid pinning the hand-derived Jacobian of
ns_oracle.c against itself, this script re-derives everything mechanically:

  * a ~60-line "mini-UFL" on sympy matrices implements the operator semantics UFL documents
    (grad(u)[i,j] = d u_i/d x_j, nabla_grad = transpose, dot = contract last/first index,
    div(T)_i = d T_ij/d x_j, inner = full contraction, derivative = exact Gateaux derivative);
  * the forms are typed in below so that they read line-for-line like the reference scripts
      gmetric : NavierStokes/NavierStokesChannelFlow.py:220-251
      ugn     : LidDrivenFlow/LidDrivenNavierStokesFlow.py:112-143
      stokes_channel : NavierStokesChannelFlow.py:160-172
      stokes_lid     : LidDrivenNavierStokesFlow.py:86-99
      stokes_duct    : StokesFlow/DuctStokesFlow.py:188-192
  * the coefficient w is a genuine polynomial function of the physical coordinate x with
    symbolic dofs, so second derivatives (div(sigma) for P2) come out of sympy.diff, and the
    Jacobian is sympy.diff of the residual with respect to each dof -- no hand algebra;
  * integration = basix degree-2 default rule (4-pt tet / 3-pt triangle), scale |det J|.

Run:  python oracle/symbolic_ref.py          (rewrites tests/golden/element_golden.npz)
Arithmetic is done with 70-digit mpmath floats and rounded to float64 at the very end.
"""
import os
import sys
import numpy as np
import sympy as sp

PREC = 70
X = sp.symbols("x0 x1 x2", real=True)


# ----------------------------------------------------------------------------- mini-UFL
def _as_matrix(a):
    return a if isinstance(a, sp.MatrixBase) else sp.Matrix([[a]])


def grad(f, gdim):
    """UFL grad: appends a derivative index.  scalar -> vector, vector -> matrix [i,j]=d f_i/d x_j."""
    if isinstance(f, sp.MatrixBase):
        assert f.shape[1] == 1
        return sp.Matrix(f.shape[0], gdim, lambda i, j: sp.diff(f[i], X[j]))
    return sp.Matrix(gdim, 1, lambda j, _: sp.diff(f, X[j]))


def nabla_grad(f, gdim):
    """UFL nabla_grad: derivative index first.  vector -> matrix [i,j]=d f_j/d x_i."""
    g = grad(f, gdim)
    return g.T if isinstance(f, sp.MatrixBase) else g


def div(f, gdim):
    """UFL div: contracts the LAST index.  vector -> scalar, matrix -> vector_i = d T_ij/d x_j."""
    if f.shape[1] == 1:
        return sum(sp.diff(f[j], X[j]) for j in range(gdim))
    return sp.Matrix(f.shape[0], 1, lambda i, _: sum(sp.diff(f[i, j], X[j]) for j in range(gdim)))


def dot(a, b):
    """UFL dot: contract last index of a with first index of b."""
    a, b = _as_matrix(a), _as_matrix(b)
    if a.shape[1] == 1 and b.shape[1] == 1:          # vector . vector
        return sum(a[i] * b[i] for i in range(a.shape[0]))
    if a.shape[1] == 1:                               # vector . matrix -> vector_j = a_i B_ij
        return sp.Matrix(b.shape[1], 1, lambda j, _: sum(a[i] * b[i, j] for i in range(a.shape[0])))
    return a * b                                      # matrix . vector / matrix . matrix


def inner(a, b):
    a, b = _as_matrix(a), _as_matrix(b)
    assert a.shape == b.shape
    return sum(a[i, j] * b[i, j] for i in range(a.shape[0]) for j in range(a.shape[1]))


def sym(a):
    return (a + a.T) / 2


def tr(a):
    return a.trace()


# ----------------------------------------------------------------------------- cell + elements
TET_EDGES = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]   # basix reference-tet edge numbering
TRI_EDGES = [(1, 2), (0, 2), (0, 1)]


class Cell:
    def __init__(self, coords):
        coords = np.asarray(coords, dtype=np.float64)
        self.gdim = gd = coords.shape[0] - 1
        self.xv = [[sp.Rational(float(coords[a, i])) for i in range(gd)] for a in range(gd + 1)]
        # Jacobian(msh): J[i][j] = x_{j+1}[i] - x_0[i]
        self.J = sp.Matrix(gd, gd, lambda i, j: self.xv[j + 1][i] - self.xv[0][i])
        self.detJ = self.J.det()
        K = self.J.inv()
        xs = sp.Matrix(gd, 1, lambda i, _: X[i] - self.xv[0][i])
        xi = K * xs                                   # reference coordinates as affine functions of x
        self.lam = [1 - sum(xi)] + [xi[a] for a in range(gd)]
        self.h = sp.sqrt(max(sum((self.xv[a][i] - self.xv[b][i]) ** 2 for i in range(gd))
                             for a in range(gd + 1) for b in range(a + 1, gd + 1)))  # CellDiameter

    def lagrange(self, deg):
        lam, gd = self.lam, self.gdim
        if deg == 1:
            return list(lam)
        edges = TET_EDGES if gd == 3 else TRI_EDGES
        return [l * (2 * l - 1) for l in lam] + [4 * lam[a] * lam[b] for a, b in edges]

    def quadrature(self):
        if self.gdim == 3:
            s5 = sp.sqrt(5)
            a, b = (5 - s5) / 20, (5 + 3 * s5) / 20
            pts = [(a, a, a), (b, a, a), (a, b, a), (a, a, b)]
            return pts, [sp.Rational(1, 24)] * 4
        pts = [(sp.Rational(1, 6), sp.Rational(1, 6)), (sp.Rational(1, 6), sp.Rational(2, 3)),
               (sp.Rational(2, 3), sp.Rational(1, 6))]
        return pts, [sp.Rational(1, 6)] * 3

    def physical_point(self, xi):
        gd = self.gdim
        return [self.xv[0][i] + sum(self.J[i, a] * xi[a] for a in range(gd)) for i in range(gd)]


# ----------------------------------------------------------------------------- the forms, as written
def form_gmetric(cell, u, p, v, q, nu, Ci=36.0):
    """NavierStokesChannelFlow.py:222-251, typed as in the reference."""
    gd = cell.gdim
    x = sp.Matrix(gd, 1, lambda i, _: X[i])
    dxi_dy = cell.J.inv()                                   # inv(Jacobian(msh))
    dxi_dx = dxi_dy * grad(x, gd).inv()                     # dxi_dy * inv(grad(x))
    G = dxi_dx.T * dxi_dx
    Ci = sp.Rational(Ci)
    tau_SUPS = 1 / sp.sqrt(inner(u, G * u) + Ci * (nu ** 2) * inner(G, G))
    sigma = 2 * nu * sym(grad(u, gd)) - p * sp.eye(gd)
    res_M = dot(u, grad(u, gd)) - div(sigma, gd)
    a = inner(dot(u, nabla_grad(u, gd)), v)
    a += nu * inner(grad(u, gd), grad(v, gd))
    a -= p * div(v, gd)
    a += q * div(u, gd)
    a += inner(tau_SUPS * res_M, dot(u, grad(v, gd)) + grad(q, gd))
    v_LSIC = 1 / (tr(G) * tau_SUPS)
    res_C = div(u, gd)
    a += v_LSIC * div(v, gd) * res_C
    return a


def form_ugn(cell, u, p, v, q, nu):
    """LidDrivenNavierStokesFlow.py:123-143, typed as in the reference (r = 2)."""
    gd = cell.gdim
    h = cell.h
    r = 2
    u_norm = sp.sqrt(dot(u, u))
    tau_SUNG1 = h / (2 * u_norm)
    inv_tau_SUNG1 = sp.Piecewise((0, u_norm <= sp.Float("1e-8")), (1 / (tau_SUNG1 ** r), True))
    tau_SUNG3 = h * h / (4 * nu)
    tau_SUPG = (inv_tau_SUNG1 + 1 / (tau_SUNG3 ** r)) ** sp.Rational(-1, r)
    Re_UGN = u_norm * h / (2 * nu)
    z = sp.Piecewise((Re_UGN / 3, Re_UGN <= 3), (1, True))
    tau_LSIC = h / 2 * u_norm * z
    a = inner(dot(u, nabla_grad(u, gd)), v)
    a += nu * inner(grad(u, gd), grad(v, gd))
    a -= p * div(v, gd)
    a += q * div(u, gd)
    res = dot(u, nabla_grad(u, gd)) - nu * div(sym(grad(u, gd)), gd) + grad(p, gd)
    a += tau_SUPG * inner(dot(u, nabla_grad(v, gd)), res)
    a += tau_SUPG * inner(grad(q, gd), res)
    a += tau_LSIC * div(v, gd) * div(u, gd)
    return a


def form_stokes(cell, u, p, v, q, alpha, sp_sign, beta):
    """alpha grad u:grad v + sp(-p div v + div u q) + beta h^2 grad p.grad q
    (NavierStokesChannelFlow.py:168-170; LidDrivenNavierStokesFlow.py:93-96; DuctStokesFlow.py:191)."""
    gd = cell.gdim
    mu_T = beta * cell.h * cell.h
    a = alpha * inner(grad(u, gd), grad(v, gd))
    a -= sp_sign * p * div(v, gd)
    a += sp_sign * div(u, gd) * q
    a += mu_T * inner(grad(p, gd), grad(q, gd))
    return a


# ----------------------------------------------------------------------------- element tensors
def _setup(kind, coords, vdeg, nu, kw):
    cell = Cell(coords)
    gd = cell.gdim
    phi_v, phi_p = cell.lagrange(vdeg), cell.lagrange(1)
    nvn, npn = len(phi_v), len(phi_p)
    nd = gd * nvn + npn
    nu = sp.Rational(float(nu))

    def integrand(u, p, v, q):
        if kind == "gmetric":
            return form_gmetric(cell, u, p, v, q, nu, kw.get("Ci", 36.0))
        if kind == "ugn":
            return form_ugn(cell, u, p, v, q, nu)
        if kind == "stokes":
            return form_stokes(cell, u, p, v, q, sp.Rational(float(kw["alpha"])), int(kw["sp"]),
                               sp.Rational(float(kw["beta"])))
        raise ValueError(kind)

    tests = []
    for m in range(nvn):
        for c in range(gd):
            v = sp.zeros(gd, 1)
            v[c] = phi_v[m]
            tests.append((v, sp.Integer(0)))
    for m in range(npn):
        tests.append((sp.zeros(gd, 1), phi_p[m]))
    return cell, gd, phi_v, phi_p, nvn, npn, nd, integrand, tests


def _integrate(cell, expr):
    """basix degree-2 rule, scale |det J| (FFCx)."""
    gd = cell.gdim
    pts, wts = cell.quadrature()
    tot = 0
    for xi, wt in zip(pts, wts):
        xq = cell.physical_point(xi)
        tot += sp.N(wt, PREC) * expr.subs({X[k]: sp.N(xq[k], PREC) for k in range(gd)})
    return tot * sp.N(sp.Abs(cell.detJ), PREC)


def element_tensors(kind, coords, w, vdeg, nu, method="fd", **kw):
    """Return (be, Ae) as float64 arrays for one cell.  Cell-local dof order: velocity node-major
    with interleaved components, then pressure (mixed_element([P_k^gdim, P1])).

    method="symbolic": dofs are sympy symbols, Ae = sympy.diff(be_i, w_j) (exact; minutes per P1 cell).
    method="fd": dofs are 70-digit numbers; the residual is still evaluated mechanically by the
      mini-UFL (all x-derivatives by sympy.diff), and Ae[:, j] is the central difference
      (F(w + eps e_j) - F(w - eps e_j)) / 2 eps with eps = 1e-25 in 70-digit arithmetic: truncation
      ~1e-50, i.e. exact to float64.  Both methods agree to the last float64 digit on P1 cells
      (tests/test_oracle.py::test_symbolic_methods_agree)."""
    cell, gd, phi_v, phi_p, nvn, npn, nd, integrand, tests = _setup(kind, coords, vdeg, nu, kw)

    if method == "symbolic":
        U = sp.symbols(f"w0:{nd}", real=True)
        u = sp.Matrix(gd, 1, lambda i, _: sum(U[gd * n + i] * phi_v[n] for n in range(nvn)))
        p = sum(U[gd * nvn + n] * phi_p[n] for n in range(npn))
        wvals = {U[k]: sp.Float(sp.Rational(float(w[k])), PREC) for k in range(nd)}
        be, Ae = np.zeros(nd), np.zeros((nd, nd))
        for i, (v, q) in enumerate(tests):
            Fi = _integrate(cell, integrand(u, p, v, q))
            be[i] = float(sp.N(Fi.subs(wvals), PREC))
            for j in range(nd):
                Ae[i, j] = float(sp.N(sp.diff(Fi, U[j]).subs(wvals), PREC))
        return be, Ae

    def residual(wv):
        u = sp.Matrix(gd, 1, lambda i, _: sp.expand(sum(wv[gd * n + i] * phi_v[n] for n in range(nvn))))
        p = sp.expand(sum(wv[gd * nvn + n] * phi_p[n] for n in range(npn)))
        return [_integrate(cell, integrand(u, p, v, q)) for v, q in tests]

    w0 = [sp.Float(sp.Rational(float(w[k])), PREC) for k in range(nd)]
    eps = sp.Float("1e-25", PREC)
    be = np.array([float(f) for f in residual(w0)])
    Ae = np.zeros((nd, nd))
    for j in range(nd):
        wp, wm = list(w0), list(w0)
        wp[j] = w0[j] + eps
        wm[j] = w0[j] - eps
        Fp, Fm = residual(wp), residual(wm)
        Ae[:, j] = [float((a - b) / (2 * eps)) for a, b in zip(Fp, Fm)]
    return be, Ae


# ----------------------------------------------------------------------------- golden cases
def golden_cases():
    rng = np.random.default_rng(20261018)
    cases = []

    def rand_cell(gd, scale):
        ref = np.vstack([np.zeros(gd), np.eye(gd)])
        return (ref + 0.25 * rng.standard_normal((gd + 1, gd))) * scale + rng.standard_normal(gd)

    # G-metric, P1-P1 tets (config 3/5), three cells incl. one with negative det J
    for k, nu in enumerate([0.1, 1.0 / 40, 1.0 / 70]):
        x = rand_cell(3, 0.05 * (k + 1))
        if k == 2:
            x = x[[0, 2, 1, 3]]
        cases.append(dict(name=f"gmetric_p1p1_tet_{k}", kind="gmetric", vdeg=1, nu=nu, x=x,
                          w=rng.standard_normal(16) * np.r_[np.ones(12), 3 * np.ones(4)]))
    # G-metric, P2-P1 tets (config 4)
    for k, nu in enumerate([0.1, 1.0 / 50]):
        cases.append(dict(name=f"gmetric_p2p1_tet_{k}", kind="gmetric", vdeg=2, nu=nu, x=rand_cell(3, 0.1),
                          w=rng.standard_normal(34)))
    # UGN, P1-P1 triangles (config 2): low-Re_UGN branch, high branch
    x = rand_cell(2, 1.0 / 64)
    cases.append(dict(name="ugn_p1p1_tri_lowRe", kind="ugn", vdeg=1, nu=1.0 / 10, x=x, w=0.3 * rng.standard_normal(9)))
    cases.append(dict(name="ugn_p1p1_tri_highRe", kind="ugn", vdeg=1, nu=1.0 / 4000, x=x * 8, w=rng.standard_normal(9) + 1.0))
    cases.append(dict(name="ugn_p2p1_tri", kind="ugn", vdeg=2, nu=1.0 / 100, x=rand_cell(2, 0.1), w=rng.standard_normal(15)))
    # Stokes flavours
    cases.append(dict(name="stokes_channel_p1p1_tet", kind="stokes", vdeg=1, nu=1.0, alpha=1.0, sp=1, beta=0.2,
                      x=rand_cell(3, 0.1), w=rng.standard_normal(16)))
    cases.append(dict(name="stokes_lid_p1p1_tri", kind="stokes", vdeg=1, nu=0.01, alpha=0.01, sp=1, beta=1.0 / (12 * 0.01),
                      x=rand_cell(2, 0.1), w=rng.standard_normal(9)))
    cases.append(dict(name="stokes_duct_p2p1_tet", kind="stokes", vdeg=2, nu=1.0, alpha=1.0, sp=-1, beta=0.0,
                      x=rand_cell(3, 0.1), w=rng.standard_normal(34)))
    return cases


def _run_case(c):
    kw = {k: c[k] for k in ("alpha", "sp", "beta") if k in c}
    be, Ae = element_tensors(c["kind"], c["x"], c["w"], c["vdeg"], c["nu"], **kw)
    return c["name"], be, Ae


def main():
    import multiprocessing as mp
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "element_golden.npz")
    only = sys.argv[1:]
    data = {}
    if only and os.path.exists(out):
        data = dict(np.load(out))
    cases = [c for c in golden_cases() if not only or c["name"] in only]
    by_name = {c["name"]: c for c in cases}
    with mp.Pool(min(len(cases), os.cpu_count() or 1)) as pool:
        for n, be, Ae in pool.imap_unordered(_run_case, cases):
            c = by_name[n]
            xpad = np.zeros((c["x"].shape[0], 3))
            xpad[:, : c["x"].shape[1]] = c["x"]
            data[n + "/x"], data[n + "/w"], data[n + "/be"], data[n + "/Ae"] = xpad, c["w"], be, Ae
            data[n + "/meta"] = np.array([{"gmetric": 0, "ugn": 1, "stokes": 2}[c["kind"]], c["x"].shape[1], c["vdeg"]], dtype=np.int64)
            data[n + "/params"] = np.array([c["nu"], 36.0, c.get("alpha", 0.0), c.get("sp", 0.0), c.get("beta", 0.0)])
            print(n, "done  |be|=%.6e |Ae|=%.6e" % (np.linalg.norm(be), np.linalg.norm(Ae)), flush=True)
    np.savez_compressed(out, **data)
    print("wrote", os.path.normpath(out))


if __name__ == "__main__":
    main()
