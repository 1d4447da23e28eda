"""Manufactured-solution cases of the reference's tests, as numpy closures for the oracle.

Restates tests/test_advection_diffusion_migration.py:6-64 of the reference
(fields C1 = C2 = cos x + sin y + 3, U = sin x + cos y + 3, Poiseuille
velocity, D = (0.5, 1), z = (2, -2), F = R = T = 1); the forcing terms
f_k = div((vel - D_k z_k grad U) C_k) - div(D_k grad C_k) are expanded by hand.
"""
import numpy as np


def C1ex(x):
    return np.cos(x[..., 0]) + np.sin(x[..., 1]) + 3.0


def Uex(x):
    return np.sin(x[..., 0]) + np.cos(x[..., 1]) + 3.0


def make_vel(vavg, d=2):
    def vel(x):
        v = np.zeros(x.shape, dtype=float)
        y = x[..., 1]
        v[..., 0] = 6.0 * vavg * y * (1.0 - y)
        return v
    return vel


def forcing(vavg, D1=0.5, D2=1.0, z1=2.0, z2=-2.0):
    def f(u, x):
        X, Y = x[..., 0], x[..., 1]
        C = np.cos(X) + np.sin(Y) + 3.0
        Cx, Cy = -np.sin(X), np.cos(Y)
        lapC = -np.cos(X) - np.sin(Y)
        Ux, Uy = np.cos(X), -np.sin(Y)
        lapU = -np.sin(X) - np.cos(Y)
        vx = 6.0 * vavg * Y * (1.0 - Y)
        out = []
        for D, z in ((D1, z1), (D2, z2)):
            out.append(vx * Cx - D * z * (Cx * Ux + Cy * Uy + C * lapU) - D * lapC)
        return out
    return f


def adm_problem(vavg, family="DG", p=1, D1=0.5, D2=1.0, variant="spectral", markers=None):
    """AdvectionDiffusionMigrationSolver of the reference test, for the oracle."""
    from oracle.forms import Problem
    conc_params = [
        {"name": "C1", "diffusion coefficient": D1, "z": 2.0, "bulk": C1ex},
        {"name": "C2", "diffusion coefficient": D2, "z": -2.0, "bulk": C1ex},
    ]
    physical_params = {
        "flow": ["diffusion", "advection", "migration", "electroneutrality"],
        "F": 1.0, "R": 1.0, "T": 1.0, "U_app": Uex,
        "bulk reaction": forcing(vavg, D1, D2),
    }
    if markers is None:
        markers = {"applied": (1, 2, 3, 4), "bulk dirichlet": (1, 2, 3, 4)}
    return Problem(conc_params, physical_params, markers, p=p, family=family,
                   vel=make_vel(vavg), variant=variant)
