"""Oracle solution of the reference's examples/bortels_threeion_nondim.py AT FULL SIZE (297 x 49 graded cells, DG Q1,
174 636 dofs; BASELINE.json configs[0]): Newton solve of the numpy oracle (oracle/newton.py: bulk initial guess,
potential pre-solve, newtonls + l2 line search, direct sparse solves).  Stores every STRIDE-th dof of each field plus
the per-field norms, against which tests/test_gpu_reference_examples.py checks the solution the UNMODIFIED example
script produces on the GPU.  Not a reference output (Firedrake is not installable here): an oracle output.
Run:  python tests/golden/make_bortels_example.py   (minutes)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
STRIDE = 97


def main():
    import mms_cases
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    t0 = time.time()
    disc = Discretization(mms_cases.bortels_oracle_mesh(100, 100, 100, 50), mms_cases.bortels_problem())
    u, info = solve_problem(disc)
    Ns = disc.ndof_scalar
    U = u.reshape(disc.nf, Ns)
    np.savez_compressed(os.path.join(HERE, "example_bortels_threeion_nondim.npz"), stride=STRIDE,
                        samples=U[:, ::STRIDE], norms=np.linalg.norm(U, axis=1), ndof_scalar=Ns,
                        node_coords_samples=disc.node_coords[::STRIDE])
    print("bortels example: %d dofs, info %s, %.1f s" % (disc.ndof, info, time.time() - t0))


if __name__ == "__main__":
    main()
