"""Cell partition and ghost-dof halo exchange (multi-GPU, one process per GPU).

Stands in for what the reference gets from DMPlex distribution with a one-facet overlap, PetscSF halo
updates and MPI reductions (`mpiexec -n N python example.py`).  Cells are partitioned;
rows are owned with their cell; every cut facet is evaluated by both owners, each keeping its own rows,
so no matrix entry is ever communicated.  Per residual / Jacobian / SpMV the only exchange is the state
of the ghost cells (grouped send/recv, NCCL over NVLink on the GPU box, gloo in the CPU tests).

Local numbering: owned cells first (in global order), then ghost cells (by owner rank, then global id).
Ghost rows exist in every local vector and matrix but are empty, so all arrays of a rank share one
layout with field stride `ncell_total * n`.
"""
import numpy as np

from .mesh import TensorMesh, _box_markers


class LocalMesh:
    """The part of a mesh one rank sees: owned + ghost cells, duck-typing `mesh.Mesh` for the engine."""

    def __init__(self, ext_mesh, order, num_owned, global_ids, ghost_owner, name=None, global_num_cells=None):
        # `order`: indices into ext_mesh cells, owned first then ghosts
        self.dim = ext_mesh.dim
        self.name = name or ext_mesh.name
        self.layers = ext_mesh.layers
        self.num_owned = int(num_owned)
        self.num_cells = int(len(order))                  # owned + ghost: host-side sizing covers both
        self.global_ids = np.asarray(global_ids, dtype=np.int64)          # per local cell
        self.ghost_owner = np.asarray(ghost_owner, dtype=np.int64)        # per ghost cell (local index - num_owned)
        self.global_num_cells = global_num_cells
        inv = -np.ones(ext_mesh.num_cells, dtype=np.int64)
        inv[order] = np.arange(len(order))
        self._coords = ext_mesh.cell_coords()[order]
        owned = order[:num_owned]
        nb = ext_mesh.nbr_cell[owned]
        loc = np.where(nb >= 0, inv[np.maximum(nb, 0)], -1)
        if np.any((nb >= 0) & (loc < 0)):
            raise ValueError("an owned cell has a neighbour that is neither owned nor ghost")
        self.nbr_cell = loc
        self.nbr_lf = ext_mesh.nbr_lf[owned]
        self.nbr_orient = ext_mesh.nbr_orient[owned]
        self.ext_marker = ext_mesh.ext_marker[owned]
        self.coords = ext_mesh.coords
        self.num_interior_facets = int((self.nbr_cell >= 0).sum())       # facets seen from owned cells
        self.num_exterior_facets = int((self.nbr_cell < 0).sum())

    def topological_dimension(self):
        return self.dim

    def geometric_dimension(self):
        return self.dim

    @property
    def cell_set(self):
        class _S:
            size = self.num_owned
        return _S

    def cell_coords(self):
        return np.ascontiguousarray(self._coords)

    def connectivity(self, marker_class):
        nc, nlf = self.nbr_cell.shape
        nbr = self.nbr_cell.copy()
        lut = np.zeros(int(self.ext_marker.max()) + 2 if nc else 1, dtype=np.int64)
        for m, k in marker_class.items():
            if 0 <= int(m) < len(lut):
                lut[int(m)] = k
        cls = lut[self.ext_marker]
        bnd = nbr < 0
        nbr[bnd] = -(1 + cls[bnd])
        code = (self.nbr_lf | (self.nbr_orient << 3)).astype(np.uint8)
        big = np.iinfo(np.int64).max
        allc = np.concatenate([np.arange(nc)[:, None], np.where(bnd, big, self.nbr_cell)], axis=1)
        rank = np.argsort(np.argsort(allc, axis=1, kind="stable"), axis=1, kind="stable")
        return (nbr.astype(np.int32), code, rank[:, 1:].astype(np.uint8), rank[:, 0].astype(np.uint8),
                (1 + (~bnd).sum(axis=1)).astype(np.uint8))


class HaloPlan:
    """Who sends which owned cells to whom, and where received cells land (local cell indices)."""

    def __init__(self, rank, size, send, recv):
        self.rank, self.size = rank, size
        self.send = send          # {peer: int64 array of local owned cell ids, sorted by global id}
        self.recv = recv          # {peer: int64 array of local ghost cell ids, sorted by global id}

    def peers(self):
        return sorted(set(self.send) | set(self.recv))


def _halo_plan(local, owner_of_ext_nbr, rank, size):
    """Build the exchange lists of a LocalMesh; `owner_of_ext_nbr(local ghost index) -> owner rank`."""
    no = local.num_owned
    recv = {}
    if len(local.ghost_owner) == 0:
        return HaloPlan(rank, size, {}, {})
    for peer in np.unique(local.ghost_owner):
        g = np.nonzero(local.ghost_owner == peer)[0] + no
        recv[int(peer)] = g[np.argsort(local.global_ids[g], kind="stable")]
    send = {}
    nb = local.nbr_cell
    is_ghost = nb >= no
    for peer in recv:
        mask = np.zeros(no, dtype=bool)
        gmask = is_ghost & (local.ghost_owner[np.maximum(nb - no, 0)] == peer)
        mask |= gmask.any(axis=1)
        s = np.nonzero(mask)[0]
        send[peer] = s[np.argsort(local.global_ids[s], kind="stable")]
    return HaloPlan(rank, size, send, recv)


def graph_partition(mesh, nparts):
    """Part id of every cell from the CELL GRAPH (cells = vertices, interior facets = edges), by recursive
    graph bisection: cells are ordered by the difference of their graph distances to two mutually far cells and
    cut where half of them (scaled to the number of parts on each side) lie on one side; both halves are split
    again.  Stands in for the graph partitioner behind the reference's DMPlex distribution; it
    uses nothing but connectivity, so it applies to gmsh-read unstructured meshes as well as to tensor meshes."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import shortest_path
    nc = mesh.num_cells
    nb = mesh.nbr_cell
    rows = np.repeat(np.arange(nc), nb.shape[1])
    cols = nb.reshape(-1)
    keep = cols >= 0
    G = sp.csr_matrix((np.ones(int(keep.sum()), dtype=np.int8), (rows[keep], cols[keep])), shape=(nc, nc))
    parts = np.zeros(nc, dtype=np.int64)

    def dist(sub, start):
        d = shortest_path(sub, method="D", directed=False, unweighted=True, indices=[start])[0]
        d[~np.isfinite(d)] = d[np.isfinite(d)].max() + 1.0            # other components: farthest
        return d

    def split(cells, first, count):
        if count == 1 or len(cells) == 0:
            parts[cells] = first
            return
        sub = G[cells][:, cells]
        # two mutually far cells a, b (pseudo-peripheral pair); cells are ordered by d(a, .) - d(b, .), so that
        # BOTH sides of the cut are "half spaces" of the graph metric (a plain BFS ball leaves a remainder that
        # may fall apart)
        a = int(np.argmax(dist(sub, 0)))
        da = dist(sub, a)
        bb = int(np.argmax(da))
        db = dist(sub, bb)
        order = np.lexsort((np.arange(len(cells)), da, da - db))
        left = count // 2
        cut = int(round(len(cells) * left / float(count)))
        split(cells[order[:cut]], first, left)
        split(cells[order[cut:]], first + left, count - left)
    split(np.arange(nc), 0, int(nparts))
    return parts


def partition_mesh(mesh, rank, size, edges=None, parts=None):
    """Rank `rank`'s part of an existing (global) mesh; returns (LocalMesh, HaloPlan).

    parts: part id per cell (e.g. from `graph_partition`); default: contiguous ranges of cell numbers (`edges`),
    which on a tile-numbered mesh are unions of tiles."""
    nc = mesh.num_cells
    if parts is None:
        if edges is None:
            edges = np.linspace(0, nc, size + 1).astype(np.int64)
        parts = np.searchsorted(edges, np.arange(nc), side="right") - 1
    parts = np.asarray(parts, dtype=np.int64)
    owned = np.nonzero(parts == rank)[0]
    nb = mesh.nbr_cell[owned]
    cand = nb[nb >= 0]
    ghosts = np.unique(cand[parts[cand] != rank])
    owner = parts[ghosts]
    o = np.lexsort((ghosts, owner))
    ghosts, owner = ghosts[o], owner[o]
    order = np.concatenate([owned, ghosts])
    local = LocalMesh(mesh, order, len(owned), order, owner, global_num_cells=nc)
    local.halo_plan = _halo_plan(local, None, rank, size)
    return local, local.halo_plan


def strip_partition(axes, rank, size, name="strip", tile=None):
    """Rank `rank`'s part of the tensor mesh over `axes`, split into strips along the first direction,
    built WITHOUT the global mesh (weak-scaling benchmarks): the strip plus one ghost layer per side.
    tile: the owned cells are numbered tile by tile (`mesh.tile_order`), as `UnitSquareMesh(..., tile=)` does."""
    axes = [np.asarray(a, dtype=float) for a in axes]
    nx = len(axes[0]) - 1
    edges = np.linspace(0, nx, size + 1).astype(np.int64)
    i0, i1 = int(edges[rank]), int(edges[rank + 1])
    lo, hi = max(i0 - 1, 0), min(i1 + 1, nx)
    sub = [axes[0][lo:hi + 1]] + axes[1:]
    marker = _box_markers([a[0] for a in axes], [a[-1] for a in axes])
    ext = TensorMesh(sub, name=name, marker_of_facet=marker)
    per_col = int(np.prod([len(a) - 1 for a in axes[1:]]))
    col = np.arange(ext.num_cells) // per_col + lo                 # global column of every ext cell
    owned_mask = (col >= i0) & (col < i1)
    gid = (col * per_col + np.arange(ext.num_cells) % per_col).astype(np.int64)
    owned = np.nonzero(owned_mask)[0]
    oshape = tuple([i1 - i0] + [len(a) - 1 for a in axes[1:]])
    operm = None
    if tile is not None and all(n % int(tile) == 0 for n in oshape):
        from .mesh import tile_order
        operm = tile_order(oshape, tile)               # owned cells are lexicographic in ext order
        owned = owned[operm]
    ghosts = np.nonzero(~owned_mask)[0]
    gowner = np.where(col[ghosts] < i0, rank - 1, rank + 1)
    o = np.lexsort((gid[ghosts], gowner))
    ghosts, gowner = ghosts[o], gowner[o]
    order = np.concatenate([owned, ghosts])
    local = LocalMesh(ext, order, len(owned), gid[order], gowner, name=name, global_num_cells=nx * per_col)
    local.halo_plan = _halo_plan(local, None, rank, size)
    # the owned cells are a tensor mesh of their own, numbered lexicographically: a rank-local multigrid hierarchy
    # (the V-cycle as the block of PETSc's bjacobi) can be built on it
    local.owned_axes = [axes[0][i0:i1 + 1]] + axes[1:]
    local.global_axes = axes                 # the coarse space shared by all ranks lives on the global vertex grid
    local.cell_perm = operm
    if operm is not None:
        local.block_hint = int(tile) ** len(axes)
    return local, local.halo_plan


class HaloExchange:
    """Ghost-cell state exchange for vectors laid out [nf][ncell_total * n].  Device side: ONE pack launch gathers
    the messages of all peers into contiguous slices of one send buffer, grouped NCCL send / recv move the slices
    (torch.distributed P2P), ONE unpack launch scatters the receive buffer into the ghost entries
    (`ech_halo_pack` / `ech_halo_unpack`).  Without a CUDA device (gloo tests) the same slices are packed with torch
    indexing."""

    def __init__(self, plan, nf, nloc, ncell_total, device):
        import torch
        self.plan, self.nf, self.nloc, self.ntot = plan, nf, nloc, ncell_total
        self.device = device
        self.peers = plan.peers()
        self._sslice, self._rslice = {}, {}
        sidx, ridx = [], []
        so = ro = 0
        for peer in self.peers:
            if peer in plan.send:
                ix = self._dof_index(plan.send[peer])
                self._sslice[peer] = (so, so + len(ix))
                so += len(ix)
                sidx.append(ix)
            if peer in plan.recv:
                ix = self._dof_index(plan.recv[peer])
                self._rslice[peer] = (ro, ro + len(ix))
                ro += len(ix)
                ridx.append(ix)
        cat = lambda parts: torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64)
        self._sidx, self._ridx = cat(sidx).to(device), cat(ridx).to(device)
        self._sbuf = torch.empty(so, dtype=torch.float64, device=device)
        self._rbuf = torch.empty(ro, dtype=torch.float64, device=device)
        self.bytes_per_exchange = 8 * ro
        self._lib = None
        if torch.device(device).type == "cuda":
            from . import capi
            self._lib = capi.lib()

    def _dof_index(self, cells):
        import torch
        cells = np.asarray(cells, dtype=np.int64)
        dof = (cells[:, None] * self.nloc + np.arange(self.nloc)[None, :]).reshape(-1)
        idx = (np.arange(self.nf)[:, None] * (self.ntot * self.nloc) + dof[None, :]).reshape(-1)
        return torch.from_numpy(idx)

    def exchange(self, u):
        """Fill the ghost-cell entries of `u` with the owners' values."""
        import torch
        import torch.distributed as dist
        if self.plan.size == 1 or not self.peers:
            return
        if self._lib is not None:
            from . import capi
            from .capi import ptr
            st = capi.C.c_void_p(torch.cuda.current_stream().cuda_stream)
            capi.check(self._lib.ech_halo_pack(self._sidx.numel(), ptr(self._sidx), ptr(u), ptr(self._sbuf), st))
        else:
            torch.index_select(u, 0, self._sidx, out=self._sbuf)
        ops = []
        for peer in self.peers:
            if peer in self._sslice:
                a, b = self._sslice[peer]
                ops.append(dist.P2POp(dist.isend, self._sbuf[a:b], peer))
            if peer in self._rslice:
                a, b = self._rslice[peer]
                ops.append(dist.P2POp(dist.irecv, self._rbuf[a:b], peer))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if self._lib is not None:
            capi.check(self._lib.ech_halo_unpack(self._ridx.numel(), ptr(self._ridx), ptr(self._rbuf), ptr(u), st))
        else:
            u.index_copy_(0, self._ridx, self._rbuf)
