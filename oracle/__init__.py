"""CPU oracle for the CMCD bridge hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + torch-CPU) of the reference's algorithm
for the path named in BASELINE.json (``/root/reference/src`` -- cited file:line in
every function).  It exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` can check and time the
hand-written CUDA path.  Nothing in ``cmcd_b200/`` may import it.

Parity pin status
-----------------
The reference is pure JAX; JAX is not installable in this image, the reference has no
tests, golden vectors or fixtures of its own, so the reference itself cannot be run
here.  What *is* pinned (tests/test_oracle_prng.py, tests/golden/):

* threefry2x32 against the Random123 known-answer vectors;
* ``split`` / ``normal`` against the values printed in JAX's public documentation
  (``split(PRNGKey(0))``, ``normal(PRNGKey(0),(1,))``, ``normal(PRNGKey(42),())``,
  ``normal(PRNGKey(0),(3,))``);
* analytic pins: ln Z = 0 for gmm / many_gmm / funnel, CAIS == ULA when the drift
  network outputs zero, K=0 ULA == MFVI bound.

Everything downstream of the PRNG (targets, networks, bridge steps, gradients) is a
restatement with **parity unpinned** against a running reference (none can run here).
"""
