// bp_fk.cuh -- K7: iiwa14 forward kinematics, one thread per configuration.
//
// Replaces the numeric (Pinocchio) branch of the reference's RobotModel:
//   fk_pos                     RobotModel.py:146-160
//   fk_pos_col                 RobotModel.py:162-181  (col_ids :27-35)
//   hom_transform_endeffector  RobotModel.py:197-211
//   jacobian_fk                RobotModel.py:213-231  (LOCAL_WORLD_ALIGNED)
// Chain constants from bound_planner/RobotModel/iiwa.urdf:22-147.  Every joint
// rotates about its own z axis; the fixed joint rotations are rpy multiples of
// pi/2, i.e. signed axis permutations (the 6e-17 residues cos(pi/2) leaves in
// Pinocchio's matrices are dropped -- far below the 1e-6 parity tolerance).
// joint_ee uses rpy = (0, -1.575, -1.575) (urdf:137; 1.575, NOT pi/2).
#pragma once
#include "bp_math.cuh"

// Rz(-1.575) * Ry(-1.575), row-major
#define BP_EE_R00 1.7670764329018813e-05
#define BP_EE_R01 0.9999911645788031
#define BP_EE_R02 0.004203623683574308
#define BP_EE_R10 0.004203623683574308
#define BP_EE_R11 -0.0042036608246882635
#define BP_EE_R12 0.9999823292356709
#define BP_EE_R20 0.9999911645788031
#define BP_EE_R21 0.0
#define BP_EE_R22 -0.0042036608246882635

struct BpFrame {
  double X[3], Y[3], Z[3], p[3];
};

// rotate the frame about its own z by q
BP_HD void bp_rotz(BpFrame& f, double q) {
  double s, c;
#ifdef __CUDA_ARCH__
  sincos(q, &s, &c);
#else
  s = sin(q); c = cos(q);
#endif
  for (int k = 0; k < 3; ++k) {
    double x = f.X[k], y = f.Y[k];
    f.X[k] = c * x + s * y;
    f.Y[k] = c * y - s * x;
  }
}

// fixed rotation C = rpy(pi/2, 0, pi) == rpy(-pi/2, pi, 0): columns (-X, Z, Y)
BP_HD void bp_fix_a(BpFrame& f) {
  for (int k = 0; k < 3; ++k) {
    double y = f.Y[k];
    f.X[k] = -f.X[k];
    f.Y[k] = f.Z[k];
    f.Z[k] = y;
  }
}

// fixed rotation C = rpy(pi/2, 0, 0): columns (X, Z, -Y)
BP_HD void bp_fix_b(BpFrame& f) {
  for (int k = 0; k < 3; ++k) {
    double y = f.Y[k];
    f.Y[k] = f.Z[k];
    f.Z[k] = -y;
  }
}

// q[7] -> p_ee[3], p_col[7*3] (joint_3..joint_7 origins, link4_col_link,
// end_effector_col_link), optional T_ee[16] (row-major 4x4), optional
// jac[6*7] (row-major; rows 0-2 linear, 3-5 angular), optional djac[6*7] = d/dt jac for the joint
// velocities dq[7] (djacobian_fk, RobotModel.py:233-251: pin.getFrameJacobianTimeVariation,
// LOCAL_WORLD_ALIGNED): with z_k, o_k the axis / origin of joint k,
//   dz_k = w_k x z_k, w_k = sum_{i<k} z_i dq_i;  do_k = sum_{i<k} dq_i z_i x (o_k - o_i);
//   v_ee = sum_i dq_i z_i x (p_ee - o_i);  djac[:3,k] = dz_k x (p_ee - o_k) + z_k x (v_ee - do_k);  djac[3:,k] = dz_k.
BP_HD void bp_cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

BP_HD void bp_fk_iiwa14(const double* q, double* p_ee, double* p_col, double* T_ee, double* jac,
                        const double* dq = nullptr, double* djac = nullptr) {
  BpFrame f;
  f.X[0] = 1; f.X[1] = 0; f.X[2] = 0;
  f.Y[0] = 0; f.Y[1] = 1; f.Y[2] = 0;
  f.Z[0] = 0; f.Z[1] = 0; f.Z[2] = 1;
  f.p[0] = 0; f.p[1] = 0; f.p[2] = 0;
  double zax[7][3], org[7][3];
#define BP_SAVE(K)                                                           \
  for (int k = 0; k < 3; ++k) { zax[K][k] = f.Z[k]; org[K][k] = f.p[k]; }
  // joint_1: xyz (0,0,0.1525), rpy 0                         (urdf:25)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.1525 * f.Z[k];
  BP_SAVE(0)
  bp_rotz(f, q[0]);
  // joint_2: xyz (0,0,0.2075), rpy (pi/2,0,pi)               (urdf:40)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.2075 * f.Z[k];
  bp_fix_a(f);
  BP_SAVE(1)
  bp_rotz(f, q[1]);
  // joint_3: xyz (0,0.2325,0), rpy (pi/2,0,pi)               (urdf:55)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.2325 * f.Y[k];
  bp_fix_a(f);
  BP_SAVE(2)
  for (int k = 0; k < 3; ++k) p_col[0 + k] = f.p[k];
  bp_rotz(f, q[2]);
  // joint_4: xyz (0,0,0.1875), rpy (pi/2,0,0)                (urdf:70)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.1875 * f.Z[k];
  bp_fix_b(f);
  BP_SAVE(3)
  for (int k = 0; k < 3; ++k) p_col[3 + k] = f.p[k];
  bp_rotz(f, q[3]);
  // link4_col: parent link_4, xyz (0,0.3,0)                  (urdf:92-97)
  for (int k = 0; k < 3; ++k) p_col[15 + k] = f.p[k] + 0.3 * f.Y[k];
  // joint_5: xyz (0,0.2125,0), rpy (-pi/2,pi,0)              (urdf:85)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.2125 * f.Y[k];
  bp_fix_a(f);
  BP_SAVE(4)
  for (int k = 0; k < 3; ++k) p_col[6 + k] = f.p[k];
  bp_rotz(f, q[4]);
  // joint_6: xyz (0,0,0.1875), rpy (pi/2,0,0)                (urdf:107)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.1875 * f.Z[k];
  bp_fix_b(f);
  BP_SAVE(5)
  for (int k = 0; k < 3; ++k) p_col[9 + k] = f.p[k];
  bp_rotz(f, q[5]);
  // joint_7: xyz (0,0.0796,0), rpy (-pi/2,pi,0)              (urdf:122)
  for (int k = 0; k < 3; ++k) f.p[k] += 0.0796 * f.Y[k];
  bp_fix_a(f);
  BP_SAVE(6)
  for (int k = 0; k < 3; ++k) p_col[12 + k] = f.p[k];
  bp_rotz(f, q[6]);
#undef BP_SAVE
  // end_effector_col: parent link_7, xyz (0,0,0.13)          (urdf:142-147)
  for (int k = 0; k < 3; ++k) p_col[18 + k] = f.p[k] + 0.13 * f.Z[k];
  // joint_ee: parent link_7, xyz (0,0,0.21), rpy (0,-1.575,-1.575)   (urdf:134-138)
  double pe[3];
  for (int k = 0; k < 3; ++k) pe[k] = f.p[k] + 0.21 * f.Z[k];
  p_ee[0] = pe[0]; p_ee[1] = pe[1]; p_ee[2] = pe[2];
  if (T_ee) {
    for (int k = 0; k < 3; ++k) {
      T_ee[4 * k + 0] = f.X[k] * BP_EE_R00 + f.Y[k] * BP_EE_R10 + f.Z[k] * BP_EE_R20;
      T_ee[4 * k + 1] = f.X[k] * BP_EE_R01 + f.Y[k] * BP_EE_R11 + f.Z[k] * BP_EE_R21;
      T_ee[4 * k + 2] = f.X[k] * BP_EE_R02 + f.Y[k] * BP_EE_R12 + f.Z[k] * BP_EE_R22;
      T_ee[4 * k + 3] = pe[k];
    }
    T_ee[12] = 0.0; T_ee[13] = 0.0; T_ee[14] = 0.0; T_ee[15] = 1.0;
  }
  if (jac) {
    for (int j = 0; j < 7; ++j) {
      double r0 = pe[0] - org[j][0], r1 = pe[1] - org[j][1], r2 = pe[2] - org[j][2];
      jac[0 * 7 + j] = zax[j][1] * r2 - zax[j][2] * r1;
      jac[1 * 7 + j] = zax[j][2] * r0 - zax[j][0] * r2;
      jac[2 * 7 + j] = zax[j][0] * r1 - zax[j][1] * r0;
      jac[3 * 7 + j] = zax[j][0];
      jac[4 * 7 + j] = zax[j][1];
      jac[5 * 7 + j] = zax[j][2];
    }
  }
  if (djac && dq) {
    double v_ee[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < 7; ++i) {
      const double r[3] = {pe[0] - org[i][0], pe[1] - org[i][1], pe[2] - org[i][2]};
      double c[3];
      bp_cross3(zax[i], r, c);
      for (int k = 0; k < 3; ++k) v_ee[k] += dq[i] * c[k];
    }
    double w[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < 7; ++j) {
      double dz[3], dob[3] = {0.0, 0.0, 0.0};
      bp_cross3(w, zax[j], dz);
      for (int i = 0; i < j; ++i) {
        const double r[3] = {org[j][0] - org[i][0], org[j][1] - org[i][1], org[j][2] - org[i][2]};
        double c[3];
        bp_cross3(zax[i], r, c);
        for (int k = 0; k < 3; ++k) dob[k] += dq[i] * c[k];
      }
      const double r[3] = {pe[0] - org[j][0], pe[1] - org[j][1], pe[2] - org[j][2]};
      const double dv[3] = {v_ee[0] - dob[0], v_ee[1] - dob[1], v_ee[2] - dob[2]};
      double c1[3], c2[3];
      bp_cross3(dz, r, c1);
      bp_cross3(zax[j], dv, c2);
      for (int k = 0; k < 3; ++k) {
        djac[k * 7 + j] = c1[k] + c2[k];
        djac[(3 + k) * 7 + j] = dz[k];
        w[k] += zax[j][k] * dq[j];
      }
    }
  }
}
