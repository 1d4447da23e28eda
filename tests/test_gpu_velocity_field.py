"""`self.vel` given as a DISCRETE field (what `FlowSolver.vel` is in the reference, echemfem/flow_solver.py:114;
examples/simple_flow_battery.py:98-114) instead of a symbolic expression: the components become data coefficients.
A velocity that the nodal space represents exactly (linear in x, y) must give the residual and Jacobian of the
symbolic velocity, and both must match the oracle."""
import numpy as np
import pytest
import torch

import mms_cases
import parity
from echemfem_b200.fd import (SpatialCoordinate, Constant, as_vector, VectorFunctionSpace, VectorFunction)
from echemfem_b200.workloads import MMSSolver


class LinearVelocitySolver(MMSSolver):
    def __init__(self, N, discrete, family="DG"):
        self.discrete = discrete
        super().__init__(N, 1.0, family=family)

    def set_velocity(self):
        x, y = SpatialCoordinate(self.mesh)
        expr = as_vector([0.3 + 1.91 * y - 0.2 * x, 0.4 * x + 0.1])
        if not self.discrete:
            self.vel = expr
            return
        u = VectorFunction(VectorFunctionSpace(self.mesh, self.V), name="Velocity")       # flow_solver.py:114
        u.interpolate(expr)
        self.vel = u


def _vel(x):
    v = np.zeros(x.shape, dtype=x.dtype)
    v[..., 0] = 0.3 + 1.91 * x[..., 1] - 0.2 * x[..., 0]
    v[..., 1] = 0.4 * x[..., 0] + 0.1
    return v


def test_velocity_field_cpu_registration():
    """Host side only: the components are registered as auxiliary fields and the weak form compiles."""
    from echemfem_b200.compiler import CompiledForm
    s = LinearVelocitySolver(3, True)
    assert len(s._aux) == 3 and s.velocity_field.sub(0)._symbol == ("a", 1)        # U_app, vel_x, vel_y
    cf = CompiledForm(s.Form, 2, len(s._aux), 2, 1, "DG")
    assert "p.a[1]" in cf.header or "a1" in cf.header


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["DG", "CG"])
def test_discrete_velocity_matches_symbolic_and_oracle(family):
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from echemfem_b200.newton import NewtonEngine
    N = 5
    prob = mms_cases.adm_problem(1.0, family=family)
    prob.vel = _vel

    def forcing(u, x):            # div((vel - D z grad Uex) Cex) - div(D grad Cex) with THIS velocity (div vel = -0.2)
        X, Y = x[..., 0], x[..., 1]
        C = np.cos(X) + np.sin(Y) + 3.0
        Cx, Cy = -np.sin(X), np.cos(Y)
        lapC = -np.cos(X) - np.sin(Y)
        Ux, Uy = np.cos(X), -np.sin(Y)
        lapU = -np.sin(X) - np.cos(Y)
        v = _vel(x)
        adv = v[..., 0] * Cx + v[..., 1] * Cy - 0.2 * C
        return [adv - D * z * (Cx * Ux + Cy * Uy + C * lapU) - D * lapC for D, z in ((0.5, 2.0), (1.0, -2.0))]
    prob.physical_params["bulk reaction"] = forcing
    disc = Discretization(unit_square_mesh(N), prob)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert])
    Fo, Jo = disc.residual(u), disc.jacobian(u, return_scale=True)
    out = {}
    for discrete in (False, True):
        s = LinearVelocitySolver(N, discrete, family=family)
        eng = NewtonEngine(s)
        s.u.tensor.copy_(torch.from_numpy(u).cuda())
        F = torch.zeros_like(s.u.tensor)
        eng.residual_jacobian(s.u.tensor, out=F)
        out[discrete] = (F.cpu().numpy(), eng.vals.cpu().numpy())
        parity.check_F(eng, out[discrete][0], Fo, "vel discrete=%s residual" % discrete, scale=disc.residual_scale(u))
        parity.check_J(eng, out[discrete][1], Jo, "vel discrete=%s jacobian" % discrete)
