"""Aggregation map of the optional coarse level of the linear solver (host logic, no device code).

The reference offers multilevel option sets (`gmg`, `amg`; echemfem/solver.py:1190-1213) besides `lu`
and `ilu`.  On the device path `lu` is served by GMRES + block-Jacobi/ILU(0); for elliptic blocks that
alone needs O(N) Krylov iterations, so a coarse level can be added: dofs are binned by their
coordinates into boxes, one coarse dof per (field, non-empty box), piecewise-constant prolongation
(the nodal bases are partitions of unity, so the prolongated coarse vector is constant per box).
The Galerkin product, restriction and prolongation are CUDA kernels (`ech_coarse_*`).
"""
import numpy as np


def box_aggregates(coords, nf, target):
    """coords: (n, dim) dof coordinates of ONE field; returns (agg[nf * n] int32, nc).

    About `target` coarse dofs in total; boxes per direction follow the aspect ratio of the bounding box.
    Empty boxes get no coarse dof (numbers are compressed)."""
    x = np.asarray(coords, dtype=float)
    if x.ndim == 1:
        x = x[:, None]
    n, dim = x.shape
    if n == 0:
        return np.zeros(0, dtype=np.int32), 0
    nb = max(1.0, (float(target) / float(nf)) ** (1.0 / dim))
    lo, hi = x.min(axis=0), x.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    aspect = span / np.prod(span) ** (1.0 / dim)
    nbs = np.maximum(1, np.round(nb * aspect).astype(np.int64))
    idx = np.minimum((nbs * (x - lo) / span).astype(np.int64), nbs - 1)
    lin = idx[:, 0]
    for k in range(1, dim):
        lin = lin * nbs[k] + idx[:, k]
    _, box = np.unique(lin, return_inverse=True)
    nagg = int(box.max()) + 1
    agg = np.concatenate([box + f * nagg for f in range(nf)]).astype(np.int32)
    return agg, nagg * nf


def box_aggregates_global(coords, nf, target, lo, hi):
    """Like `box_aggregates`, on a box grid fixed by the GLOBAL bounding box (lo, hi), so that every rank of a
    partitioned mesh numbers the boxes identically.  Returns (agg[nf * n] int32, nc, nbs): numbers are NOT
    compressed (nc = nf * number of boxes); the caller removes globally empty boxes."""
    x = np.asarray(coords, dtype=float)
    if x.ndim == 1:
        x = x[:, None]
    n, dim = x.shape
    lo, hi = np.asarray(lo, dtype=float), np.asarray(hi, dtype=float)
    nb = max(1.0, (float(target) / float(nf)) ** (1.0 / dim))
    span = np.where(hi > lo, hi - lo, 1.0)
    aspect = span / np.prod(span) ** (1.0 / dim)
    nbs = np.maximum(1, np.round(nb * aspect).astype(np.int64))
    idx = np.clip((nbs * (x - lo) / span).astype(np.int64), 0, nbs - 1)
    lin = idx[:, 0]
    for k in range(1, dim):
        lin = lin * nbs[k] + idx[:, k]
    nbox = int(np.prod(nbs))
    agg = np.concatenate([lin + f * nbox for f in range(nf)]).astype(np.int32)
    return agg, nbox * nf, nbs
