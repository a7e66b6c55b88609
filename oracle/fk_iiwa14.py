"""Oracle (TEST INFRASTRUCTURE): iiwa14 forward kinematics.

Restates what the reference's numeric branch obtains from Pinocchio
(bound_planner/RobotModel/RobotModel.py:146-231) by walking the kinematic
chain of bound_planner/RobotModel/iiwa.urdf (joint origins at :22-147).
URDF convention: a joint's fixed transform is  T = Trans(xyz) * Rz(yaw) Ry(pitch) Rx(roll),
followed by a rotation about the joint's own z axis by q_k (all axes are 0 0 1).

Pinned against the reference's serialized CasADi functions (fk_pos.ca,
fk_pos_col_*.ca, hom_trans.ca, jacobian.ca) through oracle/casadi_blob.py ->
tests/golden/fk_golden.npz.
"""
from __future__ import annotations

import numpy as np

HALF_PI = 1.5707963267948966
PI = 3.141592653589793

# (xyz, rpy) of joint_1..joint_7, iiwa.urdf:25,40,55,70,85,107,122
JOINT_ORIGINS = [
    ((0.0, 0.0, 0.1525), (0.0, 0.0, 0.0)),
    ((0.0, 0.0, 0.2075), (HALF_PI, 0.0, PI)),
    ((0.0, 0.2325, 0.0), (HALF_PI, 0.0, PI)),
    ((0.0, 0.0, 0.1875), (HALF_PI, 0.0, 0.0)),
    ((0.0, 0.2125, 0.0), (-HALF_PI, PI, 0.0)),
    ((0.0, 0.0, 0.1875), (HALF_PI, 0.0, 0.0)),
    ((0.0, 0.0796, 0.0), (-HALF_PI, PI, 0.0)),
]
# fixed frames: (parent joint index 1-based, xyz, rpy)
LINK4_COL = (4, (0.0, 0.3, 0.0), (0.0, 0.0, 0.0))           # iiwa.urdf:92-97
EE = (7, (0.0, 0.0, 0.21), (0.0, -1.575, -1.575))           # iiwa.urdf:134-138 (1.575, not pi/2)
EE_COL = (7, (0.0, 0.0, 0.13), (0.0, 0.0, 0.0))             # iiwa.urdf:142-147

Q_LOWER = np.array([-2.9670597283903604, -2.0943951023931953, -2.9670597283903604,
                    -2.0943951023931953, -2.9670597283903604, -2.0943951023931953,
                    -3.0543261909900763])
Q_UPPER = -Q_LOWER
COL_JOINT_SIZES = [0.09, 0.12, 0.09, 0.10, 0.07, 0.09, 0.075]  # RobotModel.py:37


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _hom(xyz, rpy):
    t = np.eye(4)
    t[:3, :3] = _rpy(*rpy)
    t[:3, 3] = xyz
    return t


def _rotz(q):
    t = np.eye(4)
    c, s = np.cos(q), np.sin(q)
    t[:2, :2] = [[c, -s], [s, c]]
    return t


def joint_frames(q):
    """World placement of joint frames 1..7 (Pinocchio's data.oMi[1..7])."""
    t = np.eye(4)                                            # world_joint: no origin (iiwa.urdf:9-13)
    out = []
    for k in range(7):
        t = t @ _hom(*JOINT_ORIGINS[k]) @ _rotz(q[k])
        out.append(t.copy())
    return out


def hom_transform_endeffector(q):
    """RobotModel.py:197-211 -> data.oMf[end_effector_link].homogeneous"""
    return joint_frames(q)[EE[0] - 1] @ _hom(EE[1], EE[2])


def fk_pos(q):
    """RobotModel.py:146-160"""
    return hom_transform_endeffector(q)[:3, 3]


def fk_pos_col_all(q):
    """(7,3): joint_3..joint_7 origins, link4_col_link, end_effector_col_link
    (RobotModel.py:27-35, :162-181)."""
    fr = joint_frames(q)
    pts = [fr[k][:3, 3] for k in range(2, 7)]
    pts.append((fr[LINK4_COL[0] - 1] @ _hom(LINK4_COL[1], LINK4_COL[2]))[:3, 3])
    pts.append((fr[EE_COL[0] - 1] @ _hom(EE_COL[1], EE_COL[2]))[:3, 3])
    return np.array(pts)


def fk_pos_col(q, i):
    return fk_pos_col_all(q)[i]


def fk(q):
    """[p, rotvec] of the end effector, RobotModel.py:183-195."""
    from scipy.spatial.transform import Rotation as R

    h = hom_transform_endeffector(q)
    m = np.zeros(6)
    m[:3] = h[:3, 3]
    m[3:] = R.from_matrix(h[:3, :3]).as_rotvec()
    return m


def jacobian_fk(q):
    """6x7 geometric Jacobian of end_effector_link, LOCAL_WORLD_ALIGNED
    (RobotModel.py:213-231): column k = [z_k x (p_ee - p_k); z_k]."""
    fr = joint_frames(q)
    p_ee = fk_pos(q)
    jac = np.zeros((6, 7))
    for k in range(7):
        z = fr[k][:3, 2]
        jac[:3, k] = np.cross(z, p_ee - fr[k][:3, 3])
        jac[3:, k] = z
    return jac


def djacobian_fk(q, dq):
    """Time derivative of the LOCAL_WORLD_ALIGNED frame Jacobian for joint velocities dq
    (RobotModel.py:233-251, pin.getFrameJacobianTimeVariation): with z_k, o_k the axis / origin of joint k,
        d/dt z_k = w_k x z_k,            w_k  = sum_{i<k} z_i dq_i      (angular velocity of the parent link)
        d/dt o_k = sum_{i<k} dq_i z_i x (o_k - o_i),   v_ee = sum_i dq_i z_i x (p_ee - o_i)
        dJ[:3,k] = dz_k x (p_ee - o_k) + z_k x (v_ee - do_k),   dJ[3:,k] = dz_k.
    The reference ships no djacobian.ca; this is pinned to the directional derivative of its jacobian.ca
    (tests/golden/fk_reference_blobs.npz, key djacobian)."""
    fr = joint_frames(q)
    p_ee = fk_pos(q)
    z = [f[:3, 2] for f in fr]
    o = [f[:3, 3] for f in fr]
    v_ee = sum(dq[i] * np.cross(z[i], p_ee - o[i]) for i in range(7))
    dj = np.zeros((6, 7))
    w = np.zeros(3)
    for k in range(7):
        dz = np.cross(w, z[k])
        do = sum((dq[i] * np.cross(z[i], o[k] - o[i]) for i in range(k)), np.zeros(3))
        dj[:3, k] = np.cross(dz, p_ee - o[k]) + np.cross(z[k], v_ee - do)
        dj[3:, k] = dz
        w = w + z[k] * dq[k]
    return dj


def forward_kinematics(q, dq):
    """RobotModel.py:70-77 -> (fk(q), jacobian_fk(q), djacobian_fk(q, dq))"""
    return fk(q), jacobian_fk(q), djacobian_fk(q, dq)


def acceleration_ee(q, dq, ddq):
    """RobotModel.py:258-262"""
    return djacobian_fk(q, dq) @ dq + jacobian_fk(q) @ ddq
