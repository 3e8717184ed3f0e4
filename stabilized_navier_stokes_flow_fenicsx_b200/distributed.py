"""Cell-partitioned multi-GPU plumbing (host side): x-slab partition of the synthetic duct in dolfinx
owned+ghost layout, vector halo plans, ghost-row (J.assemble) plans, and the thin communicator wrapper.

One process per GPU.  ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests) is used only for
the set-up exchanges and for barriers / max-over-ranks timing; the per-iteration halo traffic runs inside
libnsgpu.so over its own NCCL communicator (csrc/halo.cu).

Mirrors: ``x.ghostUpdate(INSERT, FORWARD)``, ``F.ghostUpdate(ADD, REVERSE)``, ``J.assemble()``
(NavierStokes/NavierStokesChannelFlow.py:57-60, :66, :75) and dolfinx ``SparsityPattern.finalize``.
"""
import os
from dataclasses import dataclass, field

import numpy as np

from . import mesh as M


# ------------------------------------------------------------------------------------------------ comm
class Comm:
    """Minimal communicator facade: single process, or torch.distributed (nccl / gloo)."""

    def __init__(self, rank=0, size=1, dist=None, device=None):
        self.rank, self.size, self.dist, self.device = rank, size, dist, device

    @staticmethod
    def single():
        return Comm()

    @staticmethod
    def from_env(backend=None):
        import torch
        import torch.distributed as dist
        rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local = int(os.environ.get("LOCAL_RANK", rank))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        device = None
        if backend == "nccl":
            torch.cuda.set_device(local)
            device = torch.device("cuda", local)
        if not dist.is_initialized():
            dist.init_process_group(backend=backend, rank=rank, world_size=size)
        return Comm(rank, size, dist, device)

    def _t(self, arr):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.device) if self.device is not None else t

    def barrier(self):
        if self.size > 1:
            self.dist.barrier()

    def _reduce(self, v, op):
        if self.size == 1:
            return v
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._reduce(v, self.dist.ReduceOp.MAX) if self.size > 1 else v

    def sum(self, v):
        if self.size == 1:
            return v
        import torch
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def bcast_bytes(self, data, n, root=0):
        """Broadcast n bytes (bytes object on root, ignored elsewhere)."""
        if self.size == 1:
            return data
        import torch
        buf = np.frombuffer(data, dtype=np.uint8).copy() if self.rank == root else np.zeros(n, dtype=np.uint8)
        t = self._t(buf)
        self.dist.broadcast(t, src=root)
        return t.cpu().numpy().tobytes()

    def exchange(self, send):
        """Sparse all-to-all of int64 arrays: ``send`` maps peer rank -> array; returns peer -> array."""
        if self.size == 1:
            return {}
        import torch
        counts = np.zeros(self.size, dtype=np.int64)
        for p, a in send.items():
            counts[p] = len(a)
        all_counts = [torch.zeros(self.size, dtype=torch.int64, device=self.device) for _ in range(self.size)]
        self.dist.all_gather(all_counts, self._t(counts))
        all_counts = np.stack([c.cpu().numpy() for c in all_counts])      # [src][dst]
        out = {}
        # pairwise ordered send/recv (works on both nccl and gloo)
        for src in range(self.size):
            for dst in range(self.size):
                n = int(all_counts[src, dst])
                if n == 0 or src == dst:
                    continue
                if self.rank == src:
                    self.dist.send(self._t(np.asarray(send[dst], dtype=np.int64)), dst=dst)
                elif self.rank == dst:
                    t = torch.zeros(n, dtype=torch.int64, device=self.device)
                    self.dist.recv(t, src=src)
                    out[src] = t.cpu().numpy()
        return out

    def close(self):
        if self.size > 1 and self.dist.is_initialized():
            self.dist.barrier()
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ partition
@dataclass
class Partition:
    """Rank-local arrays in dolfinx layout (owned dofs/cells first, ghosts after)."""
    rank: int
    size: int
    x: np.ndarray
    cells: np.ndarray
    dofmap: np.ndarray
    n_cells_owned: int
    n_owned: int
    n_ghost: int
    global_offset: int                 # first global dof owned by this rank
    ghost_global: np.ndarray           # (n_ghost,) global dof index of each ghost
    ghost_owner: np.ndarray            # (n_ghost,) owning rank
    local_to_global: np.ndarray        # (n_owned + n_ghost,)
    w: np.ndarray                      # state, local layout
    bcs: list
    meta: dict = field(default_factory=dict)

    def owned_nnz(self, asm):
        """nnz held in rows this rank owns (for whole-job nnz bookkeeping)."""
        return asm.owned_nnz()


def slab_ranges(n_long, size):
    """Balanced contiguous box-layer ranges along the duct axis."""
    base, rem = divmod(n_long, size)
    starts = [r * base + min(r, rem) for r in range(size + 1)]
    return [(starts[r], starts[r + 1]) for r in range(size)]


def duct_partition(n_cross, n_long, rank, size, length=4.0, seed=1234, noise=1e-3):
    """x-slab partition of the structured duct.  Shared interface planes are owned by the lower
    rank; there are no ghost cells (GhostMode.none, as for gmsh-imported meshes)."""
    k0, k1 = slab_ranges(n_long, size)[rank]
    nl = k1 - k0
    if nl <= 0:
        raise ValueError("more ranks than box layers")
    plane = (n_cross + 1) ** 2
    x0 = length * k0 / n_long
    local = M.create_box_tets((n_cross, n_cross, nl), p0=(x0, -0.5, -0.5), p1=(length * k1 / n_long, 0.5, 0.5), axes=(1, 2, 0))
    # local vertex v = plane index (k - k0) * plane + in-plane index; plane k0 is a ghost plane for rank > 0
    nvl = local.n_vertices
    if rank == 0:
        perm = np.arange(nvl, dtype=np.int64)                 # all owned
        n_owned_v = nvl
    else:
        perm = np.concatenate([np.arange(plane, nvl), np.arange(plane)])   # new order: owned planes, then ghost plane
        n_owned_v = nvl - plane
    inv = np.empty(nvl, dtype=np.int64)
    inv[perm] = np.arange(nvl)
    xl = local.x[perm]
    cells = inv[local.cells].astype(np.int32)
    mloc = M.Mesh(3, xl, cells, local.shape, dict(local.meta, kind="duct_slab"))
    sp = M.mixed_space(mloc, 1)
    # global vertex ids (same numbering as the unpartitioned duct_mesh)
    gv = (k0 * plane + perm).astype(np.int64)
    l2g = (4 * gv[:, None] + np.arange(4)[None, :]).ravel()
    n_owned, n_ghost = 4 * n_owned_v, 4 * (nvl - n_owned_v)
    global_offset = 0 if rank == 0 else 4 * (k0 + 1) * plane
    ghost_global = l2g[n_owned:]
    ghost_owner = np.full(n_ghost, rank - 1, dtype=np.int32)
    # state: analytic part from coordinates, noise indexed by global dof so that every partition sees the same field
    w = M.duct_state(sp, seed=seed, noise=0.0)
    if noise:
        ntot = 4 * plane * (n_long + 1)
        w = w + noise * np.random.default_rng(seed).standard_normal(ntot)[l2g]
    bcs = M.duct_bcs(sp, length=length)
    return Partition(rank, size, xl, cells, sp.dofmap, mloc.n_cells, n_owned, n_ghost, global_offset, ghost_global, ghost_owner,
                     l2g, w, bcs, {"k0": k0, "k1": k1, "plane": plane})


# ------------------------------------------------------------------------------------------------ plans
def attach(asm, part, comm):
    """Create the library's NCCL communicator and install the vector-halo plan."""
    if comm.size == 1:
        return
    raise NotImplementedError("multi-GPU plans are installed by a later milestone of this round")


def finish_pattern_exchange(asm, part, comm):
    if comm.size == 1:
        return
    raise NotImplementedError
