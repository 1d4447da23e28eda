"""Residual / Jacobian assembly of the oracle (numpy + scipy.sparse).

Restates what Firedrake/PyOP2/PETSc do below `NonlinearVariationalSolver`
(echemfem/solver.py:714-725): loops over cells, interior facets and exterior
facets with `degree = p+1` Gauss-Legendre rules (solver.py:214-251), gather by
the cell->node map, insertion into a field-blocked AIJ matrix.
The Jacobian is the exact Gateaux derivative of the residual, obtained by
complex-step differentiation of the local residuals.
"""
import numpy as np
import scipy.sparse as sp

from . import fem


class Discretization:
    def __init__(self, mesh, problem):
        self.mesh = mesh
        self.prob = problem
        d = mesh.dim
        p = problem.p
        self.d = d
        self.p = p
        fam = "DG" if problem.family == "DG" else "CG"
        self.nodes1d = fem.nodes_1d(p, fam, problem.variant)
        self.n = (p + 1) ** d
        self.nf = problem.num_fields
        # reference node coordinates (n, d)
        g = np.meshgrid(*([self.nodes1d] * d), indexing="ij")
        self.ref_nodes = np.stack([a.ravel() for a in g], axis=1)
        self.Xc = mesh.cell_vertex_coords()                    # (nc, 2^d, d)
        self._build_dofmap()
        self.ndof_scalar = int(self.cell_nodes.max()) + 1
        self.ndof = self.ndof_scalar * self.nf
        # quadrature
        self.xi_c, self.w_c = fem.cell_quadrature(p, d)
        self.t_f, self.w_f = fem.facet_quadrature(p, d)
        self.phi_c, self.dphi_c = fem.tabulate(self.nodes1d, self.xi_c)
        self._geometry_constants()
        self._mask = None

    # -- dof map ----------------------------------------------------------
    def _build_dofmap(self):
        nc = self.mesh.num_cells
        if self.prob.family == "DG":
            self.cell_nodes = np.arange(nc * self.n).reshape(nc, self.n)
            self.node_coords = self._map_points(np.arange(nc), self.ref_nodes)[0].reshape(-1, self.d)
            return
        xn = self._map_points(np.arange(nc), self.ref_nodes)[0]   # (nc, n, d)
        flat = xn.reshape(-1, self.d)
        scale = np.abs(flat).max() + 1.0
        key = np.round(flat / scale * 2 ** 40).astype(np.int64)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        inv = inv.ravel()
        self.cell_nodes = inv.reshape(nc, self.n)
        coords = np.zeros((uniq.shape[0], self.d))
        coords[inv] = flat
        self.node_coords = coords

    # -- geometry ---------------------------------------------------------
    def _map_points(self, cells, xi):
        """Physical points, Jacobians for reference points xi (npts, d) in given cells.

        Returns x (ne, npts, d), J (ne, npts, d, d) with J[.., a, b] = dx_a/dxi_b.
        """
        N, dN = fem.tabulate(np.array([0.0, 1.0]), xi)          # (npts, 2^d), (npts, 2^d, d)
        X = self.Xc[cells]                                      # (ne, 2^d, d)
        # edge form (sum N = 1, sum dN = 0): differences of neighbouring vertices are exact, so J keeps full
        # relative accuracy for small cells far from the origin (graded meshes) instead of cancelling
        E = X[:, 1:] - X[:, :1]
        x = X[:, :1] + np.einsum("qv,eva->eqa", N[:, 1:], E)
        J = np.einsum("qvb,eva->eqab", dN[:, 1:], E)
        return x, J

    def _geometry_constants(self):
        d = self.d
        nc = self.mesh.num_cells
        xi, w = fem.cell_quadrature(3, d)                      # 3 points/direction: exact for multilinear maps
        _, J = self._map_points(np.arange(nc), xi)
        self.cell_volume = np.einsum("q,eq->e", w, np.abs(np.linalg.det(J)))
        self.facet_area = np.zeros((nc, 2 * d))
        t, wf = fem.facet_quadrature(3, d)
        for lf in range(2 * d):
            xi_f = fem.facet_to_cell(lf, t, d)
            _, J = self._map_points(np.arange(nc), xi_f)
            _, ds = self._normal(J, lf)
            self.facet_area[:, lf] = np.einsum("q,eq->e", wf, ds)
        X = self.Xc
        diff = X[:, :, None, :] - X[:, None, :, :]
        self.cell_diameter = np.sqrt((diff ** 2).sum(-1)).max(axis=(1, 2))

    @staticmethod
    def _normal(J, lf):
        """Outward unit normal and surface measure on local facet lf."""
        k, s = divmod(lf, 2)
        detJ = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
        g = Jinv[..., k, :] * (2 * s - 1)                      # J^{-T} e_k, outward
        nrm = np.sqrt((g ** 2).sum(-1))
        n = g / nrm[..., None]
        ds = np.abs(detJ) * nrm
        return n, ds

    # -- field evaluation -------------------------------------------------
    def _local(self, u, cells):
        """(ne, nf, n) local dofs of the mixed vector u."""
        idx = self.cell_nodes[cells]                           # (ne, n)
        return np.stack([u[f * self.ndof_scalar + idx] for f in range(self.nf)], axis=1)

    def _eval(self, Uloc, cells, xi):
        x, J = self._map_points(cells, xi)
        phi, dphi = fem.tabulate(self.nodes1d, xi)
        Jinv = np.linalg.inv(J)
        gphi = np.einsum("qbk,eqka->eqba", dphi, Jinv)         # physical gradients (ne, q, n, d)
        return x, J, phi, gphi

    @staticmethod
    def _fields(Uloc, phi, gphi):
        nf = Uloc.shape[1]
        u = [np.einsum("eb,qb->eq", Uloc[:, f], phi) for f in range(nf)]
        gu = [np.einsum("eb,eqba->eqa", Uloc[:, f], gphi) for f in range(nf)]
        return u, gu

    def _aux(self):
        """Applied potential as the reference sees it (solver.py:373-380)."""
        ua = self.prob.physical_params.get("U_app")
        if ua is None:
            return None
        if callable(ua):
            return ua(self.node_coords)                        # nodal interpolant in Vu
        return float(ua)

    # -- local residuals --------------------------------------------------
    def _cell_local(self, Uloc, cells):
        x, J, phi, gphi = self._eval(Uloc, cells, self.xi_c)
        w = self.w_c[None, :] * np.abs(np.linalg.det(J))
        u, gu = self._fields(Uloc, phi, gphi)
        geo = {"cell_volume": self.cell_volume[cells][:, None],
               "cell_diameter": self.cell_diameter[cells][:, None]}
        f0, f1 = self.prob.cell(x, u, gu, geo)
        tr = getattr(self.prob, "transient", None)
        if tr is not None:
            # backward Euler accumulation term (C - C_old)/dt v dx of TransientEchemSolver (solver.py:2693)
            Uold = self._local(tr["u_old"], cells)
            for i in range(self.prob.num_mass):
                f0[i] = f0[i] + (u[i] - np.einsum("eb,qb->eq", Uold[:, i], phi)) / tr["dt"]
        # element-level magnitude of analytically cancelling pointwise sums (Problem.cancel0), for the tolerance
        c0 = getattr(self.prob, "cancel0", None)
        self._cell_cancel = None
        if c0 is not None and any(np.any(c) for c in c0):
            aw, ap = np.abs(w), np.abs(phi)
            self._cell_cancel = np.stack([np.einsum("eq,eq,qb->eb", aw, np.real(c), ap) +
                                          1j * np.einsum("eq,eq,qb->eb", aw, np.imag(c), ap) for c in c0], axis=1)
        return self._test(f0, f1, w, phi, gphi)

    def _test(self, f0, f1, w, phi, gphi):
        R = []
        for i in range(self.nf):
            r = np.einsum("eq,eq,qb->eb", w, f0[i], phi) + np.einsum("eq,eqa,eqba->eb", w, f1[i], gphi)
            R.append(r)
        return np.stack(R, axis=1)                             # (ne, nf, n)

    def _facet_points(self, c0, l0, c1, l1):
        """Reference points of the facet quadrature in both cells, grouped by (l0, l1, orientation)."""
        d = self.d
        xi0 = fem.facet_to_cell(l0, self.t_f, d)
        B = self.mesh.neighbour_facet_map(c0, l0, c1, l1)
        if d == 1:
            t1 = self.t_f
        else:
            N, _ = fem.tabulate(np.array([0.0, 1.0]), self.t_f)
            t1 = N @ B
        xi1 = fem.facet_to_cell(l1, t1, d)
        return xi0, xi1

    def _int_groups(self):
        """Interior facets grouped by identical reference configuration."""
        if getattr(self, "_igroups", None) is not None:
            return self._igroups
        groups = {}
        IF = self.mesh.int_facets
        for f in range(IF.shape[0]):
            c0, c1, l0, l1 = IF[f]
            B = self.mesh.neighbour_facet_map(c0, l0, c1, l1)
            key = (int(l0), int(l1), tuple(B.ravel().astype(int)))
            groups.setdefault(key, []).append(f)
        out = []
        for key, fl in groups.items():
            fl = np.array(fl)
            c0, c1, l0, l1 = IF[fl[0]]
            xi0, xi1 = self._facet_points(c0, l0, c1, l1)
            out.append((fl, int(l0), int(l1), xi0, xi1))
        self._igroups = out
        return out

    def _int_local(self, Up, Um, fl, l0, l1, xi0, xi1):
        IF = self.mesh.int_facets[fl]
        cp_, cm_ = IF[:, 0], IF[:, 1]
        x, Jp, phip, gphip = self._eval(Up, cp_, xi0)
        _, Jm, phim, gphim = self._eval(Um, cm_, xi1)
        n, ds = self._normal(Jp, l0)
        w = self.w_f[None, :] * ds
        up, gup = self._fields(Up, phip, gphip)
        um, gum = self._fields(Um, phim, gphim)
        geop = {"cell_volume": self.cell_volume[cp_][:, None], "facet_area": self.facet_area[cp_, l0][:, None]}
        geom = {"cell_volume": self.cell_volume[cm_][:, None], "facet_area": self.facet_area[cm_, l1][:, None]}
        f0p, f1p, f0m, f1m = self.prob.interior_facet(x, n, up, gup, um, gum, geop, geom)
        Rp = self._test(f0p, f1p, w, phip, gphip)
        Rm = self._test(f0m, f1m, w, phim, gphim)
        return Rp, Rm

    def _ext_groups(self):
        if getattr(self, "_egroups", None) is not None:
            return self._egroups
        EF = self.mesh.ext_facets
        mk = self.mesh.ext_marker
        out = []
        for lf in range(2 * self.d):
            for m in np.unique(mk):
                sel = np.nonzero((EF[:, 1] == lf) & (mk == m))[0]
                if len(sel):
                    out.append((sel, lf, int(m)))
        self._egroups = out
        return out

    def _ext_local(self, Uloc, sel, lf, marker, aux_nodal):
        cells = self.mesh.ext_facets[sel, 0]
        xi = fem.facet_to_cell(lf, self.t_f, self.d)
        x, J, phi, gphi = self._eval(Uloc, cells, xi)
        n, ds = self._normal(J, lf)
        w = self.w_f[None, :] * ds
        u, gu = self._fields(Uloc, phi, gphi)
        geo = {"cell_volume": self.cell_volume[cells][:, None], "facet_area": self.facet_area[cells, lf][:, None]}
        aux = {}
        if aux_nodal is not None:
            if np.isscalar(aux_nodal):
                aux["U_app"] = aux_nodal
            else:
                aux["U_app"] = np.einsum("eb,qb->eq", aux_nodal[self.cell_nodes[cells]], phi)
        f0, f1 = self.prob.exterior_facet(marker, x, n, u, gu, geo, aux)
        return self._test(f0, f1, w, phi, gphi)

    # -- global residual --------------------------------------------------
    def _rows(self, cells):
        """(ne, nf, n) global row numbers."""
        idx = self.cell_nodes[cells]
        return np.stack([f * self.ndof_scalar + idx for f in range(self.nf)], axis=1)

    def residual(self, u):
        u = np.asarray(u, dtype=float)
        F = np.zeros(self.ndof)
        nc = self.mesh.num_cells
        cells = np.arange(nc)
        R = self._cell_local(self._local(u, cells), cells)
        np.add.at(F, self._rows(cells).ravel(), R.ravel())
        if self.prob.family == "DG":
            for fl, l0, l1, xi0, xi1 in self._int_groups():
                IF = self.mesh.int_facets[fl]
                Rp, Rm = self._int_local(self._local(u, IF[:, 0]), self._local(u, IF[:, 1]), fl, l0, l1, xi0, xi1)
                np.add.at(F, self._rows(IF[:, 0]).ravel(), Rp.ravel())
                np.add.at(F, self._rows(IF[:, 1]).ravel(), Rm.ravel())
        aux = self._aux()
        for sel, lf, m in self._ext_groups():
            c = self.mesh.ext_facets[sel, 0]
            R = self._ext_local(self._local(u, c), sel, lf, m, aux)
            np.add.at(F, self._rows(c).ravel(), R.ravel())
        return F

    def residual_scale(self, u):
        """Per-field max-abs of the ELEMENT vectors (cell, interior-facet halves, exterior facets) whose sum is the
        assembled residual: the `tau_block` of residual entries.  Near a solution the assembled entries
        cancel to far below the terms that were added, and rounding error scales with the terms."""
        u = np.asarray(u, dtype=float)
        tau = np.zeros(self.nf)

        def see(R):
            if R.shape[0]:
                np.maximum(tau, np.abs(R).max(axis=(0, 2)), out=tau)
        cells = np.arange(self.mesh.num_cells)
        see(self._cell_local(self._local(u, cells), cells))
        if self._cell_cancel is not None:
            see(np.real(self._cell_cancel))
        if self.prob.family == "DG":
            for fl, l0, l1, xi0, xi1 in self._int_groups():
                IF = self.mesh.int_facets[fl]
                Rp, Rm = self._int_local(self._local(u, IF[:, 0]), self._local(u, IF[:, 1]), fl, l0, l1, xi0, xi1)
                see(Rp)
                see(Rm)
        aux = self._aux()
        for sel, lf, m in self._ext_groups():
            c = self.mesh.ext_facets[sel, 0]
            see(self._ext_local(self._local(u, c), sel, lf, m, aux))
        return tau

    # -- local Jacobians by complex step ----------------------------------
    def _cs(self, fun, U, h=1e-30, cancel=None):
        """d fun / d U for U (ne, nf, n) -> (ne, nf, n, nf, n) [row eq, row node, col field, col node].
        cancel: list that receives the matching magnitudes of analytically cancelling sums (cell terms only)."""
        ne, nf, n = U.shape
        K = np.zeros((ne, nf, n, nf, n))
        Kc = np.zeros((ne, nf, n, nf, n)) if cancel is not None else None
        Uc = U.astype(complex)
        for j in range(nf):
            for b in range(n):
                Uc[:, j, b] += 1j * h
                K[:, :, :, j, b] = np.imag(fun(Uc)) / h
                if Kc is not None and self._cell_cancel is not None:
                    Kc[:, :, :, j, b] = np.imag(self._cell_cancel) / h
                Uc[:, j, b] = U[:, j, b]
        if cancel is not None:
            cancel.append(Kc)
        return K

    def element_matrices(self, u):
        """All local matrices: dict with 'cell', 'int' [(fl, Kpp, Kpm, Kmp, Kmm)], 'ext' [(sel, K)]."""
        u = np.asarray(u, dtype=float)
        nc = self.mesh.num_cells
        cells = np.arange(nc)
        out = {}
        U = self._local(u, cells)
        canc = []
        out["cell"] = self._cs(lambda V: self._cell_local(V, cells), U, cancel=canc)
        out["cell_cancel"] = canc[0]
        out["int"] = []
        if self.prob.family == "DG":
            for fl, l0, l1, xi0, xi1 in self._int_groups():
                IF = self.mesh.int_facets[fl]
                Up = self._local(u, IF[:, 0])
                Um = self._local(u, IF[:, 1])
                ne, nf, n = Up.shape
                h = 1e-30
                Kpp = np.zeros((ne, nf, n, nf, n)); Kmp = np.zeros_like(Kpp)
                Kpm = np.zeros_like(Kpp); Kmm = np.zeros_like(Kpp)
                Upc = Up.astype(complex); Umc = Um.astype(complex)
                for j in range(nf):
                    for b in range(n):
                        Upc[:, j, b] += 1j * h
                        Rp, Rm = self._int_local(Upc, Umc, fl, l0, l1, xi0, xi1)
                        Kpp[:, :, :, j, b] = np.imag(Rp) / h
                        Kmp[:, :, :, j, b] = np.imag(Rm) / h
                        Upc[:, j, b] = Up[:, j, b]
                        Umc[:, j, b] += 1j * h
                        Rp, Rm = self._int_local(Upc, Umc, fl, l0, l1, xi0, xi1)
                        Kpm[:, :, :, j, b] = np.imag(Rp) / h
                        Kmm[:, :, :, j, b] = np.imag(Rm) / h
                        Umc[:, j, b] = Um[:, j, b]
                out["int"].append((fl, Kpp, Kpm, Kmp, Kmm))
        out["ext"] = []
        aux = self._aux()
        for sel, lf, m in self._ext_groups():
            c = self.mesh.ext_facets[sel, 0]
            U = self._local(u, c)
            K = self._cs(lambda V: self._ext_local(V, sel, lf, m, aux), U)
            out["ext"].append((sel, K))
        return out

    # -- structural block mask ----------------------------
    def block_mask(self, nsamples=2, seed=0):
        """(own[nf,nf], nbr[nf,nf]) booleans: which field blocks the Jacobian form has.

        Structural dependence is probed numerically: a block is present iff its
        exact derivative is non-zero for some random positive state.
        """
        if self._mask is not None:
            return self._mask
        rng = np.random.default_rng(seed)
        nf = self.nf
        own = np.eye(nf, dtype=bool)
        nbr = np.zeros((nf, nf), dtype=bool)
        for _ in range(nsamples):
            u = 0.5 + rng.random(self.ndof)
            em = self.element_matrices(u)

            def nz(K):
                return np.abs(K).max(axis=(0, 2, 4)) > 0 if K.shape[0] else np.zeros((nf, nf), bool)
            own |= nz(em["cell"])
            for _, Kpp, Kpm, Kmp, Kmm in em["int"]:
                own |= nz(Kpp) | nz(Kmm)
                nbr |= nz(Kpm) | nz(Kmp)
            for _, K in em["ext"]:
                own |= nz(K)
        own |= nbr
        self._mask = (own, nbr)
        return self._mask

    # -- sparsity pattern (field-blocked AIJ, sorted columns; Q5) ---------
    def pattern(self):
        own, nbr = self.block_mask()
        nc = self.mesh.num_cells
        n, nf, N = self.n, self.nf, self.ndof_scalar
        rows = []
        cols = []
        cn = self.cell_nodes
        for i in range(nf):
            for j in range(nf):
                if own[i, j]:
                    r = (i * N + cn)[:, :, None] + np.zeros((1, 1, n), dtype=np.int64)
                    c = (j * N + cn)[:, None, :] + np.zeros((1, n, 1), dtype=np.int64)
                    rows.append(r.ravel()); cols.append(c.ravel())
                if nbr[i, j] and self.prob.family == "DG":
                    IF = self.mesh.int_facets
                    for a, b in ((0, 1), (1, 0)):
                        r = (i * N + cn[IF[:, a]])[:, :, None] + np.zeros((1, 1, n), dtype=np.int64)
                        c = (j * N + cn[IF[:, b]])[:, None, :] + np.zeros((1, n, 1), dtype=np.int64)
                        rows.append(r.ravel()); cols.append(c.ravel())
        rows = np.concatenate(rows)
        cols = np.concatenate(cols)
        P = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(self.ndof, self.ndof))
        P.sum_duplicates()
        P.sort_indices()
        return P.indptr.astype(np.int64), P.indices.astype(np.int32)

    def jacobian(self, u, return_scale=False):
        """CSR Jacobian on the fixed pattern (explicit zeros kept).

        return_scale: also return, on the same pattern, the sum of the ABSOLUTE values of the element-level
        contributions (cell matrix + facet macro-element blocks) to every entry.  Its block-wise maximum is the
        `tau_block` of SURVEY 8c: assembled entries may cancel to far below the element blocks that were added
        (advection cell term against upwind facet term), and rounding error scales with those."""
        own, nbr = self.block_mask()
        em = self.element_matrices(u)
        nf, n = self.nf, self.n
        rows = []; cols = []; vals = []

        def add(rc, cc, K, mask):
            R = self._rows(rc)                                   # (ne, nf, n)
            C = self._rows(cc)
            for i in range(nf):
                for j in range(nf):
                    blk = K[:, i, :, j, :]
                    if not mask[i, j]:
                        assert not np.any(blk), "non-zero entries outside the allocated pattern"
                        continue
                    r = R[:, i, :, None] + np.zeros((1, 1, n), dtype=np.int64)
                    c = C[:, j, None, :] + np.zeros((1, n, 1), dtype=np.int64)
                    rows.append(r.ravel()); cols.append(c.ravel()); vals.append(blk.ravel())
        cells = np.arange(self.mesh.num_cells)
        add(cells, cells, em["cell"], own)
        for fl, Kpp, Kpm, Kmp, Kmm in em["int"]:
            IF = self.mesh.int_facets[fl]
            add(IF[:, 0], IF[:, 0], Kpp, own)
            add(IF[:, 0], IF[:, 1], Kpm, nbr)
            add(IF[:, 1], IF[:, 0], Kmp, nbr)
            add(IF[:, 1], IF[:, 1], Kmm, own)
        for sel, K in em["ext"]:
            c = self.mesh.ext_facets[sel, 0]
            add(c, c, K, own)
        indptr, indices = self.pattern()
        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.ndof, self.ndof))
        A.sum_duplicates()
        J = _on_pattern(A, indptr, indices, self.ndof)
        if not return_scale:
            return J
        if np.any(em["cell_cancel"]):
            # magnitudes of analytically cancelling pointwise sums (Problem.cancel0): part of what was added
            add(cells, cells, np.where(own[None, :, None, :, None], em["cell_cancel"], 0.0), own)
        S = sp.csr_matrix((np.abs(np.concatenate(vals)), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.ndof, self.ndof))
        S.sum_duplicates()
        return J, _on_pattern(S, indptr, indices, self.ndof).data

    # -- helpers ----------------------------------------------------------
    def interpolate(self, f):
        """Nodal interpolant of a float-or-callable into the scalar space."""
        if callable(f):
            return np.asarray(f(self.node_coords), dtype=float)
        return float(f) * np.ones(self.ndof_scalar)

    def field(self, u, f):
        return u[f * self.ndof_scalar:(f + 1) * self.ndof_scalar]

    def l2_error(self, uf, exact, nq=None):
        """sqrt(int (u_h - exact)^2): `errornorm(exact, u_h)` of the reference tests."""
        d = self.d
        x1, w1 = fem.gauss_legendre(nq or self.p + 3)
        g = np.meshgrid(*([x1] * d), indexing="ij")
        xi = np.stack([a.ravel() for a in g], axis=1)
        wg = np.meshgrid(*([w1] * d), indexing="ij")
        w = np.ones_like(wg[0])
        for a in wg:
            w = w * a
        w = w.ravel()
        cells = np.arange(self.mesh.num_cells)
        x, J = self._map_points(cells, xi)
        phi, _ = fem.tabulate(self.nodes1d, xi)
        uh = np.einsum("eb,qb->eq", uf[self.cell_nodes], phi)
        err = uh - exact(x)
        return np.sqrt(np.einsum("q,eq,eq->", w, np.abs(np.linalg.det(J)), err ** 2))


def _on_pattern(A, indptr, indices, n):
    """Values of A laid out on the CSR pattern (indptr, indices); zeros where A has no entry."""
    A = A.tocsr()
    A.sort_indices()
    vals = np.zeros(len(indices))
    rows_p = np.repeat(np.arange(n), np.diff(indptr))
    key_p = rows_p.astype(np.int64) * n + indices
    rows_a = np.repeat(np.arange(n), np.diff(A.indptr))
    key_a = rows_a.astype(np.int64) * n + A.indices
    pos = np.searchsorted(key_p, key_a)
    assert np.all(key_p[pos] == key_a), "entry outside the pattern"
    vals[pos] = A.data
    return sp.csr_matrix((vals, indices, indptr), shape=(n, n))
