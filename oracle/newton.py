"""Newton driver of the oracle.

Restates the PETSc configuration the reference requests (echemfem/solver.py:
938-957, 1173-1189): SNES `newtonls` with the `l2` line search, rtol = atol =
1e-16, max_it 50, divergence tolerance 1e6 -- so convergence is in practice
declared by PETSc's default step tolerance stol = 1e-8 -- and a
direct LU solve (`pc_type="lu"`, the default of every example) of each
linearised system.  Also the potential pre-solve of `setup_solver`
(solver.py:645-663): Newton on the potential block with concentrations frozen
at the initial guess, starting from U_app / 2, snes_rtol 1e-8, max_it 20.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def l2_linesearch(resid, x, y, f0norm2):
    """One iteration of PETSc's SNESLineSearchL2: secant step on d||F||^2/dlambda
    from ||F||^2 at lambda in {0, 1/2, 1}; returns (lambda, F(x - lambda y))."""
    lam, lam_old = 1.0, 0.0
    lam_mid = 0.5 * (lam + lam_old)
    fm = resid(x - lam_mid * y)
    f1 = resid(x - lam * y)
    n_mid = fm @ fm
    n_1 = f1 @ f1
    dl = lam - lam_old
    d_new = (3.0 * n_1 - 4.0 * n_mid + f0norm2) / dl
    d_old = (-3.0 * f0norm2 + 4.0 * n_mid - n_1) / dl
    d2 = (d_new - d_old) / dl
    if d2 == 0.0 or not np.isfinite(d2):
        return lam, f1
    lam_new = lam - d_new / d2
    if not np.isfinite(lam_new) or lam_new <= 1e-12:
        lam_new = 0.5 * (lam + lam_old)
    lam_new = min(lam_new, 1e8)
    return lam_new, resid(x - lam_new * y)


def newton(resid, jac, x0, rtol=1e-16, atol=1e-16, stol=1e-8, max_it=50, dtol=1e6,
           linesearch="l2", fixed=None, monitor=None):
    """Solve resid(x) = 0.  `fixed`: boolean mask of strongly constrained dofs (CG Dirichlet)."""
    x = x0.copy()
    F = resid(x)
    if fixed is not None:
        F[fixed] = 0.0
    f0 = np.linalg.norm(F)
    fn = f0
    its = 0
    hist = [fn]
    if monitor:
        monitor(0, fn)
    reason = "max_it"
    if fn < atol:
        return x, {"its": 0, "reason": "atol", "history": hist}
    while its < max_it:
        J = jac(x).tocsr()
        if fixed is not None:
            keep = sp.diags((~fixed).astype(float))
            J = keep @ J @ keep + sp.diags(fixed.astype(float))
        y = spla.spsolve(J.tocsc(), F)

        def r(z):
            g = resid(z)
            if fixed is not None:
                g[fixed] = 0.0
            return g
        if linesearch == "l2":
            lam, Fn = l2_linesearch(r, x, y, fn * fn)
        else:
            lam, Fn = 1.0, r(x - y)
        x = x - lam * y
        F = Fn
        fn = np.linalg.norm(F)
        its += 1
        hist.append(fn)
        if monitor:
            monitor(its, fn)
        if fn < atol:
            reason = "atol"; break
        if fn <= rtol * f0:
            reason = "rtol"; break
        if lam * np.linalg.norm(y) < stol * np.linalg.norm(x):
            reason = "stol"; break
        if fn > dtol * f0:
            raise RuntimeError("oracle Newton diverged")
    return x, {"its": its, "reason": reason, "history": hist}


def solve_problem(disc, initial_guess=True, initial_solve=True, monitor=None, u0=None):
    """`setup_solver()` + `solve()` of the reference (solver.py:528-727) on the oracle."""
    prob = disc.prob
    N = disc.ndof_scalar
    u = np.zeros(disc.ndof) if u0 is None else u0.copy()
    fixed, gvals = dirichlet_dofs(disc)
    if initial_guess and u0 is None:
        for i, k in enumerate(prob.idx_c):                        # solver.py:571-582
            u[i * N:(i + 1) * N] = disc.interpolate(prob.conc_params[k]["bulk"])
        if prob.has_potential:
            iU = prob.i_Ul
            ua = prob.physical_params.get("U_app")
            nU = 1
            U0 = np.zeros(N)
            if prob.i_Us is not None:                              # solver.py:609-643: both potentials
                nU = 2
                U0 = np.concatenate([np.zeros(N), disc.interpolate(ua)])
            elif ua is not None:
                U0 = disc.interpolate(ua) / 2.0                    # solver.py:646-648
            if initial_solve:                                      # solver.py:649-662
                sl = slice(iU * N, (iU + nU) * N)
                fx = None if fixed is None else fixed[sl]

                def rU(U):
                    v = u.copy(); v[sl] = U
                    return disc.residual(v)[sl]

                def jU(U):
                    v = u.copy(); v[sl] = U
                    return disc.jacobian(v)[sl, sl]
                if fx is not None:
                    U0[fx] = gvals[sl][fx]
                U0, _ = newton(rU, jU, U0, rtol=1e-8, atol=1e-50, stol=1e-8, max_it=20, fixed=fx)
            u[iU * N:(iU + nU) * N] = U0
    if fixed is not None:
        u[fixed] = gvals[fixed]
    return newton(disc.residual, disc.jacobian, u, fixed=fixed, monitor=monitor)


def dirichlet_dofs(disc):
    """Strong DirichletBC dofs (CG only; on DG spaces a DirichletBC constrains nothing)."""
    prob = disc.prob
    if prob.family == "DG":
        return None, None
    mesh = disc.mesh
    N = disc.ndof_scalar
    fixed = np.zeros(disc.ndof, dtype=bool)
    vals = np.zeros(disc.ndof)
    from . import fem
    for field, ids, spec in prob.dirichlet_sets():
        for f, (c, lf) in enumerate(mesh.ext_facets):
            if mesh.ext_marker[f] not in ids:
                continue
            k, s = divmod(int(lf), 2)
            on = np.nonzero(np.abs(disc.ref_nodes[:, k] - s) < 1e-14)[0]
            nodes = disc.cell_nodes[c, on]
            x = disc.node_coords[nodes]
            if spec[0] == "bulk":
                b = prob.conc_params[spec[1]]["bulk"]
                v = b(x) if callable(b) else b * np.ones(len(nodes))
            elif spec[0] == "gas":
                b = prob.conc_params[spec[1]]["gas"]
                v = b(x) if callable(b) else b * np.ones(len(nodes))
            elif spec[0] == "U_app":
                b = prob.physical_params["U_app"]
                v = b(x) if callable(b) else b * np.ones(len(nodes))
            else:
                v = np.zeros(len(nodes))
            fixed[field * N + nodes] = True
            vals[field * N + nodes] = v
    return fixed, vals
