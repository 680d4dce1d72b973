"""CPU oracle for the MuPoPP per-timestep hot path.  TEST INFRASTRUCTURE ONLY.

This package is an fp64 NumPy/SciPy restatement of the weak forms that the
reference (`/root/reference/src/modules/mupopp.py`) hands to DOLFIN 2017.1 and
of the DOLFIN behaviours those forms rely on (mesh/dof numbering, symmetric
Dirichlet elimination, direct LU solve).  DOLFIN/FFC/PETSc are third-party,
un-vendored and not installable here, so the arithmetic is
restated from the UFL in mupopp.py (file:line cited on every function) and from
DOLFIN 2017.1's published conventions.

PARITY STATUS: the reference ships no golden vectors, known-answer tests or
asserted tolerances for this path ("parity unpinned" for field VALUES).  What is
pinned: the UFC dof numbering of P2-vector/triangle spaces against the
reference's own fixture `src/run/fenics_demo/dolfin_fine_velocity.xml.gz`
(tests/golden/fenics_demo_dofmap.npz, tests/test_oracle_golden.py) and the
analytical known answers of SURVEY.md section 4.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`mupopp_b200/`) never imports it and has no CPU fallback.
"""
