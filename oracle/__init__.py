"""CPU oracle for the MPC hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product path
(``mpc4quantum_b200``) fails loudly if ``libm4q.so`` is missing.

Parity status (see DESIGN.md "Oracle"):

* rows pinned by the reference's own code run verbatim through
  ``oracle/refshim.py`` in the build container (discretisation,
  linearisation, line search, the ``mpc()`` loop, lift/proj) -> the numpy
  restatement in ``oracle/restate.py`` is checked against them to ~1e-13 and
  the outputs are committed as ``tests/golden/*.npz``;
* QP leaf (``optimize.quad_program``: cvxpy 1.1.13 -> OSQP 0.6.2.post0) and
  plant leaf (``QExperiment.simulate``: qutip 4.6.2 ``mesolve``) live in
  third-party packages that are absent from this image and cannot be
  installed: **parity unpinned** for those two leaves; the oracle restates
  their published mathematics exactly (strictly convex QP solved to machine
  precision with an independent KKT certificate; ``expm`` conjugation).
"""
