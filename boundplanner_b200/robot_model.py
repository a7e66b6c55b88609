"""Drop-in for the numeric (Pinocchio) branch of the reference's ``RobotModel``
(bound_planner/RobotModel/RobotModel.py:146-231), backed by the FK kernel.

Only the kinematic queries the MPC loop makes on NumPy inputs are mirrored
(BoundMPC.py:480-481, MPCNode.py:38,118): ``fk_pos``, ``fk_pos_col``, ``fk``,
``hom_transform_endeffector``, ``jacobian_fk``, ``djacobian_fk``,
``forward_kinematics(q, dq)`` (:70-77), ``velocity_ee`` / ``acceleration_ee`` /
``omega_ee`` (:253-267), plus batched forms.  The symbolic CasADi branch (used
inside the OCP) stays in the reference.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as R

from . import geometry as geo

Q_LIM_UPPER = np.array([2.9670597283903604, 2.0943951023931953, 2.9670597283903604, 2.0943951023931953,
                        2.9670597283903604, 2.0943951023931953, 3.0543261909900763])   # iiwa.urdf:27-124


class RobotModel:
    def __init__(self):
        self.col_joint_sizes = [0.09, 0.12, 0.09, 0.10, 0.07, 0.09, 0.075]      # RobotModel.py:37
        self.q_lim_upper = Q_LIM_UPPER.copy()
        self.q_lim_lower = -Q_LIM_UPPER
        self.dq_lim_upper = 10.0 * np.ones(7)                                     # velocity="10" in the URDF
        self.dq_lim_lower = -self.dq_lim_upper
        self.tau_lim_lower = [-320, -320, -176, -176, -110, -40, -40]
        self.tau_lim_upper = [320, 320, 176, 176, 110, 40, 40]
        self.u_max = 35
        self.u_min = -35

    def get_robot_limits(self):
        return (self.q_lim_upper, self.q_lim_lower, self.dq_lim_upper, self.dq_lim_lower, self.tau_lim_upper,
                self.tau_lim_lower, self.u_max, self.u_min)

    @staticmethod
    def _numeric(q):
        if not isinstance(q, np.ndarray):
            raise NotImplementedError("boundplanner_b200.RobotModel replaces the numeric (np.ndarray) branch only; "
                                      "symbolic CasADi inputs stay with the reference RobotModel")
        return np.ascontiguousarray(q, dtype=np.float64).reshape(1, 7)

    # ---- batched device forms -------------------------------------------
    @staticmethod
    def fk_batch(q, want_pose=False, want_jacobian=False):
        """q [B,7] (host or CUDA tensor) -> (p_ee [B,3], p_col [B,7,3], T_ee, jac) CUDA tensors."""
        return geo.fk_iiwa14(q, want_pose=want_pose, want_jacobian=want_jacobian)

    # ---- reference surface ------------------------------------------------
    def fk_pos(self, q):
        return geo.fk_iiwa14(self._numeric(q))[0][0].cpu().numpy()

    def fk_pos_col(self, q, i):
        return geo.fk_iiwa14(self._numeric(q))[1][0, i].cpu().numpy()

    def hom_transform_endeffector(self, q):
        return geo.fk_iiwa14(self._numeric(q), want_pose=True)[2][0].cpu().numpy()

    def fk(self, q):
        h = self.hom_transform_endeffector(q)
        m = np.zeros(6)
        m[:3] = h[:3, 3]
        m[3:] = R.from_matrix(h[:3, :3]).as_rotvec()
        return m

    def jacobian_fk(self, q):
        return geo.fk_iiwa14(self._numeric(q), want_jacobian=True)[3][0].cpu().numpy()

    def djacobian_fk(self, q, dq):
        """RobotModel.py:233-251: time variation of the LOCAL_WORLD_ALIGNED frame Jacobian along dq."""
        dq = np.ascontiguousarray(dq, dtype=np.float64).reshape(1, 7)
        return geo.fk_kinematics(self._numeric(q), dq)[2][0].cpu().numpy()

    def forward_kinematics(self, q, dq):
        """RobotModel.py:70-77 -> (p_robot [p, rotvec], jac_ee [6,7], djac_ee [6,7]); one kernel launch and one
        D2H copy (the reference makes three Pinocchio passes)."""
        dq = np.ascontiguousarray(dq, dtype=np.float64).reshape(1, 7)
        T, J, dJ = geo.fk_kinematics(self._numeric(q), dq)
        h = T[0].cpu().numpy()
        m = np.zeros(6)
        m[:3] = h[:3, 3]
        m[3:] = R.from_matrix(h[:3, :3]).as_rotvec()
        return m, J[0].cpu().numpy(), dJ[0].cpu().numpy()

    @staticmethod
    def forward_kinematics_batch(q, dq):
        """q, dq [B,7] -> (T_ee [B,4,4], jac [B,6,7], djac [B,6,7]) CUDA tensors."""
        return geo.fk_kinematics(q, dq)

    def acceleration_ee(self, q, dq, ddq):
        """RobotModel.py:258-262"""
        dq = np.ascontiguousarray(dq, dtype=np.float64).reshape(7)
        T, J, dJ = geo.fk_kinematics(self._numeric(q), dq.reshape(1, 7))
        return dJ[0].cpu().numpy() @ dq + J[0].cpu().numpy() @ np.asarray(ddq, dtype=np.float64).reshape(7)

    def velocity_ee(self, q, dq):
        return (self.jacobian_fk(q) @ dq)[:3]

    def omega_ee(self, q, dq):
        return (self.jacobian_fk(q) @ dq)[3:]
