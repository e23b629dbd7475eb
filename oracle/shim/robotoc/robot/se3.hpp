// oracle/shim/robotoc/robot/se3.hpp -- TEST INFRASTRUCTURE.  Stand-in for the reference's se3.hpp (an alias of
// pinocchio::SE3, which is not in this image): just enough of a rigid transform for ContactStatus / ImpactStatus to hold
// contact placements.  None of the hot-path arithmetic that is compared with the oracle goes through this type.
#ifndef ROBOTOC_SE3_HPP_
#define ROBOTOC_SE3_HPP_

#include "Eigen/Core"

namespace robotoc {

class SE3 {
 public:
  SE3() : R_(Eigen::Matrix3d::Identity()), p_(Eigen::Vector3d::Zero()) {}
  SE3(const Eigen::Matrix3d& R, const Eigen::Vector3d& p) : R_(R), p_(p) {}
  static SE3 Identity() { return SE3(); }
  const Eigen::Matrix3d& rotation() const { return R_; }
  const Eigen::Vector3d& translation() const { return p_; }
  Eigen::Matrix3d& rotation() { return R_; }
  Eigen::Vector3d& translation() { return p_; }
  bool isApprox(const SE3& o, double prec = 1e-12) const { return R_.isApprox(o.R_, prec) && (p_ - o.p_).norm() <= prec * (1.0 + p_.norm()); }
  friend std::ostream& operator<<(std::ostream& os, const SE3& m) { return os << "R =\n" << m.R_ << "\np = " << m.p_.transpose(); }

 private:
  Eigen::Matrix3d R_;
  Eigen::Vector3d p_;
};

}  // namespace robotoc

#endif  // ROBOTOC_SE3_HPP_
