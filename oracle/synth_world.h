// ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product path.
//
// Double-precision CPU definition of the synthetic world used by the
// BASELINE.json configs ("synthetic 7-DOF arm", "sphere-FK", "sphere/box
// obstacles").  The reference takes validity / FK as opaque user callbacks
// (ompl::base::StateValidityChecker::isValid at pr_ompl/src/RRTConnect.cpp:198,
// mFkFunc at frechet/src/NNFrechet.cpp:295,372), so this world has no
// counterpart in /root/reference: this file IS its specification
//  and the CUDA kernels are judged against it.
//
//   arm      : dof revolute joints, standard DH:  T_i = Rz(q_i+theta0_i) Tz(d_i) Tx(a_i) Rx(alpha_i)
//              (cos/sin of alpha are given directly so +-90deg is exact)
//   spheres  : S link spheres, sphere s rigidly attached to frame link[s]
//              (0 = base frame, i = frame after joint i), local centre + radius
//   obstacles: spheres (centre, radius) and axis-aligned boxes (centre, half extents)
//   clearance(q) = min over (link sphere, obstacle) of
//                    |c - o| - (r + R)                         sphere obstacle
//                    |max(|c - o| - h, 0)| - r                 box obstacle
//   isValid(q)   = clearance(q) > 0
#ifndef ORACLE_SYNTH_WORLD_H
#define ORACLE_SYNTH_WORLD_H
#include <cmath>
#include <limits>
#include <vector>
namespace oracle {
struct SynthWorld {
  int dof = 0;
  std::vector<double> a, d, ca, sa, theta0;      // per joint
  std::vector<int> sphere_link;                  // per link sphere
  std::vector<double> sphere_local;              // S x 4 (x,y,z,r)
  std::vector<double> obs_sphere;                // Ns x 4 (x,y,z,R)
  std::vector<double> obs_box;                   // Nb x 6 (cx,cy,cz,hx,hy,hz)

  int numLinkSpheres() const { return (int)sphere_link.size(); }
  int numObsSpheres() const { return (int)(obs_sphere.size() / 4); }
  int numObsBoxes() const { return (int)(obs_box.size() / 6); }

  // frames[(i)*12 .. ] = 3x4 row-major [R|p] of frame i, i = 0..dof (frame 0 = identity)
  void frames(const double *q, double *F) const {
    double *T = F;
    T[0] = 1; T[1] = 0; T[2] = 0; T[3] = 0;
    T[4] = 0; T[5] = 1; T[6] = 0; T[7] = 0;
    T[8] = 0; T[9] = 0; T[10] = 1; T[11] = 0;
    for (int i = 0; i < dof; ++i) {
      const double th = q[i] + theta0[i];
      const double ct = std::cos(th), st = std::sin(th);
      const double cA = ca[i], sA = sa[i];
      // local DH matrix
      const double L[12] = {ct, -st * cA, st * sA,  a[i] * ct,
                            st, ct * cA,  -ct * sA, a[i] * st,
                            0.0, sA,      cA,       d[i]};
      const double *P = F + 12 * i;
      double *N = F + 12 * (i + 1);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
          double v = P[4 * r + 0] * L[c] + P[4 * r + 1] * L[4 + c] + P[4 * r + 2] * L[8 + c];
          if (c == 3)
            v += P[4 * r + 3];
          N[4 * r + c] = v;
        }
      }
    }
  }
  // end-effector pose (frame dof) as 3x4 row-major
  void fk(const double *q, double *pose12) const {
    std::vector<double> F(12 * (dof + 1));
    frames(q, F.data());
    for (int i = 0; i < 12; ++i)
      pose12[i] = F[12 * dof + i];
  }
  // world centres of the link spheres, S x 3
  void linkSpheres(const double *q, double *centres) const {
    std::vector<double> F(12 * (dof + 1));
    frames(q, F.data());
    for (int s = 0; s < numLinkSpheres(); ++s) {
      const double *T = F.data() + 12 * sphere_link[s];
      const double *l = sphere_local.data() + 4 * s;
      for (int r = 0; r < 3; ++r)
        centres[3 * s + r] = T[4 * r + 0] * l[0] + T[4 * r + 1] * l[1] + T[4 * r + 2] * l[2] + T[4 * r + 3];
    }
  }
  double clearance(const double *q) const {
    const int S = numLinkSpheres();
    std::vector<double> c(3 * S);
    linkSpheres(q, c.data());
    double best = std::numeric_limits<double>::infinity();
    for (int s = 0; s < S; ++s) {
      const double r = sphere_local[4 * s + 3];
      const double *p = c.data() + 3 * s;
      for (int o = 0; o < numObsSpheres(); ++o) {
        const double *ob = obs_sphere.data() + 4 * o;
        const double dx = p[0] - ob[0], dy = p[1] - ob[1], dz = p[2] - ob[2];
        const double cl = std::sqrt(dx * dx + dy * dy + dz * dz) - (r + ob[3]);
        if (cl < best)
          best = cl;
      }
      for (int o = 0; o < numObsBoxes(); ++o) {
        const double *ob = obs_box.data() + 6 * o;
        double ex = std::fabs(p[0] - ob[0]) - ob[3];
        double ey = std::fabs(p[1] - ob[1]) - ob[4];
        double ez = std::fabs(p[2] - ob[2]) - ob[5];
        ex = ex > 0.0 ? ex : 0.0;
        ey = ey > 0.0 ? ey : 0.0;
        ez = ez > 0.0 ? ez : 0.0;
        const double cl = std::sqrt(ex * ex + ey * ey + ez * ez) - r;
        if (cl < best)
          best = cl;
      }
    }
    return best;
  }
  bool isValid(const double *q) const { return clearance(q) > 0.0; }
};
} // namespace oracle
#endif
