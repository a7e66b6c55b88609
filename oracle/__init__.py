"""CPU oracle for the BoundPlanner geometry hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with NumPy/SciPy, the algorithm that the
reference (Thieso/BoundPlanner, pure Python) runs on its hot path:

* ``convex_set_finder`` -- bound_planner/BoundPlanner/ConvexSetFinder.py
* ``obstacles``         -- bound_planner/BoundPlanner/BoundPlanner.py:126-152,
                           bound_planner/utils/util_functions.py:66-79,119-133
* ``set_graph``         -- bound_planner/BoundPlanner/BoundPlanner.py:774-798
* ``fk_iiwa14``         -- bound_planner/RobotModel/RobotModel.py:146-211 +
                           bound_planner/RobotModel/iiwa.urdf
* ``casadi_blob``       -- decoder / VM for the reference's *.ca FK functions
* ``reduce_ineqs``      -- bound_planner/utils/util_functions.py:82-88 (cddlib)
* ``planner_graph``     -- BoundPlanner.check_intersection / projection QP
                           (BoundPlanner.py:745-772, :842-864) and the planner loop's
                           rejection / duplicate / shortest-path steps (:459-478,
                           :505-512, :434 -- the latter is networkx itself)
  (``convex_set_finder`` also covers find_set_around_line / mvie_socp_fixed_r,
  :242-307 / :564-588, and general polytope obstacles: ragged row / vertex counts,
  ``closest_points_segment_polytopes`` for the segment QP of :52-99)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the *checker* or the
*timed CPU baseline*.  Nothing under ``boundplanner_b200/`` imports it: the
product path fails loudly when the CUDA library is missing.

PARITY PINNING.  The reference's arithmetic for this path lives in third-party
solvers that are absent from /root/reference and not installable here
(casadi 3.6.7 -> OSQP / qpOASES, cvxpy 1.6.3 -> Clarabel, pycddlib 3.0.2,
pin 3.4.0; see requirements.txt:1-9).  The reference ships no tests, fixtures or
golden vectors.  What IS pinned:

* ``set_graph.set_intersection`` executes the reference's own call verbatim
  (scipy.optimize.linprog / HiGHS, BoundPlanner.py:779-784) -- this row is
  pinned against the real third-party solver, run here.
* ``fk_iiwa14`` is pinned against the reference's own serialized CasADi
  functions (fk_pos.ca, fk_pos_col_{0..5}.ca, hom_trans.ca, jacobian.ca -- the
  files RobotModel.py:158,179,209,229 load), decoded and evaluated by
  ``oracle/casadi_blob.py`` without CasADi; golden vectors generated from those
  blobs are committed as tests/golden/fk_reference_blobs.npz (script:
  tests/golden/make_golden.py).  Agreement: <= 7e-16.
* closest-point QPs (OSQP / qpOASES in the reference) and the MVIE SOCP
  (Clarabel) are restated with exact solvers (active-set enumeration, barrier
  Newton + KKT check).  For these rows **parity is unpinned** against the true
  third-party solvers; they are checked against analytic known answers, against
  independent SciPy solvers (SLSQP / trust-constr) and through KKT residuals.
"""
