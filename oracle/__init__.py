"""CPU oracle for the JOTS implicit conduction solve  --  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the algorithm the reference
(j-signorelli/jots, C++ on top of MFEM/hypre) runs for its conduction solve:
full sparse assembly of mass / diffusion / Newton-Jacobian matrices, SpMV based
CG / GMRES / FGMRES with Jacobi or Chebyshev smoothing, Newton, backward Euler.

It is NOT the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker / the timed CPU baseline.  The product path
(``jots_b200`` -> ``libjots_b200.so``) never imports, links or executes it.

Parity status
-------------
The arithmetic of the reference's hot path lives in MFEM 4.7 / hypre, which are
NOT vendored in /root/reference and are absent from this image, so the reference
cannot be compiled or run here.  The reference ships no golden vectors; its tests
pin results only through analytic solutions with tolerances
(tests/test_helper.hpp.in:107-159, :261-376).  This oracle is pinned against all
of those (tests/test_oracle_analytic.py: Reinert B1/B2/B3 linearized + nonlinear
at 1e-8, Danish model 1 Dirichlet + Neumann at 3e-5, using the reference's own
error functional, tests/test_helper.hpp.in:134).  Below those tolerances the
oracle is checked against things that do not depend on it: exact integrals of the
Lagrange basis (tests/test_oracle_exact.py -- on affine hexes of order <= 3 MFEM's
default rule integrates the benchmarked forms exactly, so the discrete operator IS
the exact Galerkin integral; on quads the (p + 1)-point Gauss sum, which is not),
the defining minimisation problems of CG / GMRES / FGMRES and SciPy's CG
(tests/test_oracle_krylov.py), hypre's degree-one Chebyshev polynomial, and the
expected convergence orders.  What no test here can establish is that MFEM 4.7 /
hypre really use those rules, stopping criteria and nodal-interpolation semantics
(restated in the module docstrings): against the reference
itself **parity is unpinned**.
"""
