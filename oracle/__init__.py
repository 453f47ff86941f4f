"""CPU oracle for the SCP hot path of ahmadgazar/centroidal-MPC.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.  The CUDA path must fail loudly
when its extension is missing instead of falling back to anything here.

PARITY UNPINNED.  The reference ships no tests, golden vectors or recorded
outputs for this path, and it cannot run in this environment (jax, osqp,
pinocchio are absent).  The oracle is therefore a
restatement, in numpy/scipy float64, of

  * the reference's own code (each function cites the file:line it follows,
    relative to /root/reference), and
  * the published OSQP algorithm (Stellato et al., "OSQP: an operator
    splitting solver for quadratic programs", 2020; OSQP 0.6.x C sources as
    remembered: the reference imports ``osqp`` un-pinned, setup.py:1-7).

It is cross-checked against an independent solver (scipy's bundled HiGHS QP)
and a KKT certificate in tests/test_oracle_qp.py, which is the strongest pin
available here.
"""
