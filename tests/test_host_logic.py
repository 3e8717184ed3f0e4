"""Host-side mirror of the reference's SNES seam, exercised on the CPU with a stand-in assembler (no GPU, no libnsgpu):
NonlinearPDE_SNESProblem keeps the callback signatures of NavierStokes/NavierStokesChannelFlow.py:40-75 and must
behave like them -- copy x into the state mirror, fill F in place (ghost part zero), hand J its CSR values."""
import ast
import os

import numpy as np

from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NonlinearPDE_SNESProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeAssembler:
    """Same attributes / methods the mirror uses; residual = 2x - 1, Jacobian values = 7, 8, 9, ... (recognisable)."""
    def __init__(self, n_owned=6, n_ghost=2, nnz=11):
        self.n_owned, self.n_ghost, self.n_dofs, self.nnz = n_owned, n_ghost, n_owned + n_ghost, nnz
        self.options, self.calls = {}, []

    def set_option(self, name, value):
        self.options[name] = value

    def create_matrix(self):
        return np.arange(self.n_dofs + 1, dtype=np.int64), np.arange(self.nnz, dtype=np.int32)

    def residual(self, x):
        self.calls.append("F")
        return 2.0 * np.asarray(x) - 1.0

    def jacobian(self, x):
        self.calls.append("J")
        return 7.0 + np.arange(self.nnz)


def test_snes_callbacks_follow_the_reference_sequence():
    asm = _FakeAssembler()
    u = np.zeros(asm.n_dofs)
    prob = NonlinearPDE_SNESProblem(asm, u=u)
    assert asm.options.get("fuse_fj") == 1                      # F-then-J at the same iterate: one assembly pass
    x = np.linspace(0.0, 1.0, asm.n_dofs)
    F = np.full(asm.n_dofs, np.nan)
    prob.F(None, x, F)
    np.testing.assert_array_equal(u, x)                          # x.copy(self.u.x.petsc_vec)
    np.testing.assert_allclose(F[: asm.n_owned], 2.0 * x[: asm.n_owned] - 1.0)
    assert np.all(F[asm.n_owned:] == 0.0)                        # ghost part of F after the reverse scatter
    J = np.zeros(asm.nnz)
    prob.J(None, x, J, None)
    np.testing.assert_array_equal(J, 7.0 + np.arange(asm.nnz))
    assert asm.calls == ["F", "J"]
    assert NonlinearPDE_SNESProblem(_FakeAssembler(), fuse=False).asm.options["fuse_fj"] == 0


def test_snes_jacobian_into_a_petsc_like_matrix():
    class Mat:
        def zeroEntries(self): self.zeroed = True
        def setValuesCSR(self, ip, ix, v): self.csr = (np.array(ip), np.array(ix), np.array(v))
        def assemble(self): self.assembled = True
    asm = _FakeAssembler(n_owned=4, n_ghost=1, nnz=5)
    prob = NonlinearPDE_SNESProblem(asm)
    A = Mat()
    prob.J(None, np.zeros(asm.n_dofs), A, None)
    ip, ix, v = A.csr
    assert A.zeroed and A.assembled and ip.dtype == np.int32    # PetscInt of the pinned build is 32-bit
    assert len(ip) == asm.n_owned + 1 and len(ix) == ip[-1] == len(v)   # owned rows only: ghost rows were shipped to their owners


def test_tools_parse_without_their_optional_dependencies():
    for name in ("compare_with_dolfinx.py", "sweep.py", "bench_configs.py", "e2e_chunks.py", "spmv_sweep.py", "ncu_summary.py"):
        ast.parse(open(os.path.join(ROOT, "tools", name)).read())
