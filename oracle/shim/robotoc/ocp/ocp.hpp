// oracle/shim/robotoc/ocp/ocp.hpp -- TEST INFRASTRUCTURE.  Stand-in for robotoc::OCP (include/robotoc/ocp/ocp.hpp): the
// Riccati sources read only `robot`, `N` and `reserved_num_discrete_events` from it (src/riccati/riccati_recursion.cpp:9-16);
// cost / constraints / contact sequence (the rest of the real struct) are outside the hot path.
#ifndef ROBOTOC_OCP_HPP_
#define ROBOTOC_OCP_HPP_

#include "robotoc/robot/robot.hpp"

namespace robotoc {

struct OCP {
  Robot robot;
  double T = 0;
  int N = 0;
  int reserved_num_discrete_events = 0;
};

}  // namespace robotoc

#endif  // ROBOTOC_OCP_HPP_
