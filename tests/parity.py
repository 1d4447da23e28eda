"""Entry-level tolerance of the parity tests (north_star: 1e-12 relative in FP64).

    |a - b| <= 1e-12 * max(|a|, |b|, tau_block)

`tau_block` is the max-abs of the ELEMENT BLOCK the entry belongs to (pure relative error is meaningless for
entries that cancel to ~0, a global scale hides blocks that are orders of magnitude smaller than the penalty-
or conductivity-scaled ones).  "Element block" is meant literally: the element-level (cell / facet macro-element)
blocks whose sum the assembled block is -- an assembled block may cancel to far below them (advection cell term
against upwind facet term), and rounding error, in the oracle's complex step as much as in the kernels, scales
with what was added.  Where the oracle is at hand it supplies the sum of absolute element contributions per
entry (`Discretization.jacobian(u, return_scale=True)`); without it the assembled block itself is the scale.

* Jacobian, discontinuous spaces: block = (row field, row cell, column field, column cell), the n x n entries one
  pair of cells contributes for one pair of fields;
* Jacobian, continuous spaces: assembled entries are sums over the cells around a node, so the block is
  (row field, row node, column field): the entries of one row inside one field-pair segment;
* residual: one scale per field -- the larger of the field's assembled max-abs and, when the oracle is at hand,
  the max-abs of the ELEMENT vectors (cell, facet halves) whose sum the residual is
  (`Discretization.residual_scale`): near a solution the assembled entries cancel to far below the terms that
  were added, and rounding error scales with the terms, not with what is left of them.

No extra safety factor on the relative part.  Rounding floor: on top of it an entry is allowed ROW_ULPS units in
the last place of the largest magnitude in its matrix row (vector: of the field scale).  A small block can carry
rounding residue of a large neighbour: the upwind block across a facet whose normal is orthogonal to the velocity
receives eps * |vel| (the computed normal component is 1e-16 instead of 0) next to diffusion-sized entries 1e5
times smaller.  No implementation, the reference included, can agree with another below that floor.
"""
import numpy as np

RTOL = 1e-12
ROW_ULPS = 32
EPS = np.finfo(float).eps


def _group_max(keys, mag):
    """max of `mag` over equal `keys` (int64), broadcast back to the entries."""
    uniq, inv = np.unique(keys, return_inverse=True)
    gmax = np.zeros(len(uniq))
    np.maximum.at(gmax, inv, mag)
    return gmax[inv]


def csr_block_keys(indptr, indices, nf, nscalar, unit, continuous=False):
    """Block key of every CSR entry.  `nscalar`: rows per field; `unit`: dofs per cell (DG) -- ignored for CG."""
    nrows = len(indptr) - 1
    rows = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(indptr))
    cols = np.asarray(indices, dtype=np.int64)
    fi, ri = rows // nscalar, rows % nscalar
    fj, cj = cols // nscalar, cols % nscalar
    if continuous:
        return (fi * nscalar + ri) * nf + fj
    nunit = (nscalar + unit - 1) // unit
    return ((fi * nunit + ri // unit) * nf + fj) * nunit + cj // unit


def assert_csr_close(a, b, indptr, indices, nf, nscalar, unit, what, continuous=False, rtol=RTOL, scale=None):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, what
    mag = np.maximum(np.abs(a), np.abs(b))
    if scale is not None:
        mag = np.maximum(mag, np.abs(scale))
    tau = _group_max(csr_block_keys(indptr, indices, nf, nscalar, unit, continuous), mag)
    err = np.abs(a - b)
    lens = np.diff(indptr)
    rowmax = np.zeros(len(lens))
    nz = lens > 0
    rowmax[nz] = np.maximum.reduceat(mag, np.asarray(indptr[:-1])[nz])
    floor = ROW_ULPS * EPS * np.repeat(rowmax, lens)
    bad = err > rtol * tau + floor
    if bad.any():
        k = int(np.argmax(np.where(tau > 0, err / np.maximum(tau, 1e-300), 0.0)))
        raise AssertionError("%s: %d of %d entries beyond %.0e * tau_block; worst entry %d: %.17g vs %.17g "
                             "(err %.3e, tau_block %.3e, ratio %.3e)"
                             % (what, int(bad.sum()), len(a), rtol, k, a[k], b[k], err[k], tau[k],
                                err[k] / max(tau[k], 1e-300)))


def assert_vec_close(a, b, nf, what, rtol=RTOL, scale=None):
    """Residual vectors, field-blocked: one scale per field (`scale[f]`: element-level scale from the oracle)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape and len(a) % nf == 0, what
    A, B = a.reshape(nf, -1), b.reshape(nf, -1)
    for f in range(nf):
        tau = max(np.abs(A[f]).max(), np.abs(B[f]).max(), 0.0 if scale is None else float(scale[f]))
        err = np.abs(A[f] - B[f]).max()
        assert err <= rtol * tau, "%s field %d: max abs err %.3e vs field scale %.3e (ratio %.3e)" % (
            what, f, err, tau, err / max(tau, 1e-300))


def worst_ratio_csr(a, b, indptr, indices, nf, nscalar, unit, continuous=False):
    """max over entries of err / tau_block (diagnostic)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    mag = np.maximum(np.abs(a), np.abs(b))
    tau = _group_max(csr_block_keys(indptr, indices, nf, nscalar, unit, continuous), mag)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(tau > 0, np.abs(a - b) / tau, 0.0)
    return float(r.max()) if len(r) else 0.0


# -- convenience wrappers over a NewtonEngine ------------------------------------------------------------------
def check_F(eng, F, F_ref, what, rtol=RTOL, scale=None):
    assert_vec_close(F, F_ref, eng.nf, what, rtol, scale)


def check_J(eng, vals, J_ref, what, rtol=RTOL):
    """`J_ref`: scipy CSR matrix on the same (bit-exact) pattern as the engine's, or the pair
    `disc.jacobian(u, return_scale=True)`."""
    scale = None
    if isinstance(J_ref, tuple):
        J_ref, scale = J_ref
    assert_csr_close(vals, J_ref.data, J_ref.indptr, J_ref.indices, eng.nf, eng.N, eng.nloc, what,
                     continuous=(eng.family == "CG"), rtol=rtol, scale=scale)
