"""Host-side mirror of the reference's SNES seam, exercised on the CPU with a stand-in assembler (no GPU, no libnsgpu):
NonlinearPDE_SNESProblem keeps the callback signatures of NavierStokes/NavierStokesChannelFlow.py:40-75 and must
behave like them -- copy x into the state mirror, fill F in place (ghost part zero), hand J its CSR values."""
import ast
import os

import numpy as np

import pytest

from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler, NonlinearPDE_SNESProblem, NonlinearProblem, newton_solve
from stabilized_navier_stokes_flow_fenicsx_b200.distributed import Comm

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
    with pytest.raises(ValueError, match="global column"):      # local column indices must not reach setValuesCSR once ghosts exist
        NonlinearPDE_SNESProblem(asm).J(None, np.zeros(asm.n_dofs), Mat(), None)
    prob = NonlinearPDE_SNESProblem(asm, local_to_global=100 + np.arange(asm.n_dofs))
    A = Mat()
    prob.J(None, np.zeros(asm.n_dofs), A, None)
    ip, ix, v = A.csr
    assert A.zeroed and A.assembled and ip.dtype == np.int32    # PetscInt of the pinned build is 32-bit
    assert len(ip) == asm.n_owned + 1 and len(ix) == ip[-1] == len(v)   # owned rows only: ghost rows were shipped to their owners
    np.testing.assert_array_equal(ix, 100 + np.arange(len(ix)))          # global columns

    class LocalMat(Mat):                                         # dolfinx's create_matrix installs local-to-global maps
        def setValuesLocalCSR(self, ip, ix, v): self.local = (np.array(ip), np.array(ix), np.array(v))
    B = LocalMat()
    NonlinearPDE_SNESProblem(asm).J(None, np.zeros(asm.n_dofs), B, None)
    np.testing.assert_array_equal(B.local[1], np.arange(len(B.local[1])))


def test_row_pointers_beyond_int32_are_refused():
    class Mat:
        def zeroEntries(self): pass
        def setValuesLocalCSR(self, ip, ix, v): raise AssertionError("must not be reached")
        def assemble(self): pass
    class Big(_FakeAssembler):
        def create_matrix(self):
            return np.array([0, 2 ** 31 + 5, 2 ** 31 + 6], dtype=np.int64), np.zeros(1, dtype=np.int32)
    asm = Big(n_owned=1, n_ghost=1, nnz=1)
    with pytest.raises(OverflowError, match="32-bit"):
        NonlinearPDE_SNESProblem(asm).J(None, np.zeros(2), Mat(), None)


def test_nonlinear_problem_adapter_and_newton_iteration():
    """dolfinx NonlinearProblem shape (LidDrivenNavierStokesFlow.py:150-169): form(x), F(x, b), J(x, A) driving the
    incremental-criterion Newton iteration.  The stand-in problem is F(x) = x^2 - c (J diagonal)."""
    class Quad(_FakeAssembler):
        def __init__(self):
            super().__init__(n_owned=5, n_ghost=0, nnz=5)
            self.c = np.array([1.0, 4.0, 9.0, 16.0, 25.0])
        def create_matrix(self):
            return np.arange(6, dtype=np.int64), np.arange(5, dtype=np.int32)
        def residual(self, x):
            self.calls.append("F"); return np.asarray(x) ** 2 - self.c
        def jacobian(self, x):
            self.calls.append("J"); return 2.0 * np.asarray(x)
    asm = Quad()
    u = np.zeros(5)
    prob = NonlinearProblem(asm, u=u)
    assert asm.options.get("fuse_fj") == 1
    x = np.full(5, 3.0)
    its, ok = newton_solve(prob, x, lambda A, b: b / A, rtol=1e-12)
    assert ok and its < 12
    np.testing.assert_allclose(x, [1.0, 2.0, 3.0, 4.0, 5.0], rtol=1e-12)
    assert asm.calls[:4] == ["F", "J", "F", "J"]               # F then J at every iterate
    assert np.abs(u - x).max() < 1e-3                           # form() keeps the state mirror in step (one iterate behind at most)


def test_state_vector_lengths_are_checked_before_the_c_abi_sees_them():
    """the C ABI reads n_owned + n_ghost doubles from x_local; a short array must never reach it."""
    asm = NSAssembler.__new__(NSAssembler)                       # no GPU: only the host-side checks are exercised
    asm.n_owned, asm.n_ghost, asm.n_dofs = 6, 2, 8
    np.testing.assert_array_equal(asm._state(np.arange(8.0)), np.arange(8.0))
    padded = asm._state(np.arange(6.0))                          # Vec.array: owned entries only -> padded, ghosts come from the halo
    assert padded.size == 8 and np.all(padded[6:] == 0.0)
    with pytest.raises(ValueError, match="expected n_owned"):
        asm._state(np.arange(5.0))
    with pytest.raises(ValueError, match="at least 8"):
        asm._out(np.zeros(7), 8, "residual output")
    asm.ctx = None


def test_comm_over_an_mpi4py_like_communicator():
    class FakeMPI:                                               # the five calls the facade makes, single rank
        def Get_rank(self): return 0
        def Get_size(self): return 1
        def Barrier(self): pass
    c = Comm.from_mpi4py(FakeMPI())
    assert (c.rank, c.size) == (0, 1) and c.max(3.5) == 3.5 and c.sum(7) == 7
    assert c.bcast_bytes(b"abc", 3) == b"abc" and c.exchange({}) == {}
    c.barrier(); c.close()

    class FakeMPI2(FakeMPI):                                     # rank 0 of 2 with canned answers: checks the call shapes
        def Get_size(self): return 2
        def allgather(self, v): return [v, 2 * v]
        def bcast(self, data, root=0): return data
        def alltoall(self, lst): return [None, np.array([5, 6], dtype=np.int64)]
    c2 = Comm.from_mpi4py(FakeMPI2())
    assert c2.max(2.0) == 4.0 and c2.sum(3) == 9
    out = c2.exchange({1: np.array([1, 2, 3])})
    assert list(out) == [1] and list(out[1]) == [5, 6]


def test_tools_parse_without_their_optional_dependencies():
    for name in ("compare_with_dolfinx.py", "sweep.py", "bench_configs.py", "e2e_chunks.py", "spmv_sweep.py", "ncu_summary.py"):
        ast.parse(open(os.path.join(ROOT, "tools", name)).read())
