"""Golden vectors of the oracle for the five BASELINE.json configurations at toy sizes.

The reference itself cannot run offline (Firedrake / PETSc missing), so these fixtures are NOT
reference outputs: they freeze the oracle (`oracle/`, the numpy restatement of echemfem/solver.py) so that
 (a) a change of the oracle that alters any residual / Jacobian entry is caught on the CPU, and
 (b) the CUDA path is compared with stored numbers as well as with the live oracle.
Regenerate with  `python tests/golden/make_golden.py`  (writes tests/golden/*.npz).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def cases():
    """name -> (Discretization, state u)."""
    import mms_cases
    import product_cases as pc
    from oracle.mesh import unit_square_mesh, tensor_mesh
    from oracle.assemble import Discretization
    out = {}
    # configs[1]: MMS DG Q1
    d = Discretization(unit_square_mesh(4), mms_cases.adm_problem(1.0))
    x = d.node_coords
    out["cfg2_mms_dg_q1"] = (d, np.concatenate([mms_cases.C1ex(x), mms_cases.Uex(x) + 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1])]))
    # configs[0]: Bortels 2D
    d = Discretization(mms_cases.bortels_oracle_mesh(3, 4, 3, 3), mms_cases.bortels_problem())
    x = d.node_coords
    pert = 1e-2 * np.sin(3 * x[:, 0] + 5 * x[:, 1])
    out["cfg1_bortels_2d"] = (d, np.concatenate([1.0 + pert, 1.0 - 0.5 * pert, 0.1 + pert]))
    # configs[2]: Gupta CG
    d = Discretization(mms_cases.gupta_oracle_mesh(4, 2, 2), mms_cases.gupta_problem(SUPG=True))
    x = d.node_coords
    ph = np.sin(7e2 * x[:, 0] + 3e3 * x[:, 1])
    bulk = [b for _, _, b, _ in pc.GUPTA_SPECIES[:4]]
    out["cfg3_gupta_cg_supg"] = (d, np.concatenate([b * (1.0 + 0.05 * ph * (1 + 0.1 * k)) for k, b in enumerate(bulk)] + [1e-3 * ph]))
    # configs[3]: Bortels extruded 3D
    d = Discretization(mms_cases.bortels_oracle_mesh(3, 3, 3, 3, layers=2, hz=6.0), mms_cases.bortels_problem(hz=6.0))
    x = d.node_coords
    pert = 1e-2 * np.sin(3 * x[:, 0] + 5 * x[:, 1] + 0.7 * x[:, 2])
    out["cfg4_bortels_3d"] = (d, np.concatenate([1.0 + pert, 1.0 - 0.5 * pert, 0.1 + pert]))
    # configs[4]: catalyst layer
    d = Discretization(tensor_mesh([np.linspace(0.0, 1.0, 5), np.linspace(0.0, 0.5, 3)]), mms_cases.catalyst_layer_problem(1))
    x = d.node_coords
    ph = np.sin(5 * x[:, 0] + 3 * x[:, 1])
    bulk = [b for _, _, _, b, _, _ in pc.CL_SPECIES]
    out["cfg5_catalyst_layer"] = (d, np.concatenate([b * (1.0 + 0.05 * ph * (1 + 0.1 * k)) for k, b in enumerate(bulk)] +
                                                    [-0.02 * (1 + ph), -0.3 + 0.02 * ph]))
    return out


def main():
    for name, (d, u) in cases().items():
        # Fscale / Jscale: element-level scales of the entries (the tau_block of tests/parity.py)
        J, Jscale = d.jacobian(u, return_scale=True)
        indptr, indices = d.pattern()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), u=u, F=d.residual(u), indptr=indptr.astype(np.int64),
                            indices=indices.astype(np.int32), data=J.data, Fscale=d.residual_scale(u), Jscale=Jscale)
        print(name, "ndof", d.ndof, "nnz", J.nnz)


if __name__ == "__main__":
    main()
