"""CPU oracle for the EchemFEM Newton-step hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
algorithm the reference (LLNL/echemfem, `echemfem/solver.py`) delegates to
Firedrake/PETSc.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product package `echemfem_b200` never does.

PARITY UNPINNED at entry level: the reference ships no golden vectors, stored
matrices or expected iteration counts (its tests are manufactured-solution
convergence inequalities only) and Firedrake/PETSc are not installable in the
build container.  The oracle is therefore pinned only by the reference's own
acceptance criteria (the MMS error ratios of `tests/test_*.py`, restated in
`tests/test_oracle_mms.py`) and by internal consistency checks (complex-step
Jacobian vs finite differences, +/- restriction symmetry).  Conventions that
live in Firedrake rather than in the reference (nodal variant, quadrature
family, dof ordering) sit behind explicit switches, see `oracle/fem.py`.
"""
