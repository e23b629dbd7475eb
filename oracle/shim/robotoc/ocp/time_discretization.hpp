// oracle/shim/robotoc/ocp/time_discretization.hpp -- TEST INFRASTRUCTURE.  Stand-in for robotoc::TimeDiscretization as
// the Riccati sources see it (src/riccati/riccati_recursion.cpp:35-131: size(), operator[], grid()): a list of the
// reference's own GridInfo (include/robotoc/ocp/grid_info.hpp, used unmodified).  The producer
// (TimeDiscretization::discretize, which needs ContactSequence) is restated by tests/schedule_fixture.py instead.
#ifndef ROBOTOC_TIME_DISCRETIZATION_HPP_
#define ROBOTOC_TIME_DISCRETIZATION_HPP_

#include <vector>

#include "robotoc/ocp/grid_info.hpp"

namespace robotoc {

class TimeDiscretization {
 public:
  TimeDiscretization() {}
  explicit TimeDiscretization(const std::vector<GridInfo>& grid) : grid_(grid) {}
  int size() const { return int(grid_.size()); }
  const GridInfo& operator[](const int i) const { return grid_[size_t(i)]; }
  const GridInfo& grid(const int i) const { return grid_[size_t(i)]; }
  const GridInfo& front() const { return grid_.front(); }
  const GridInfo& back() const { return grid_.back(); }
  std::vector<GridInfo>& grids() { return grid_; }

 private:
  std::vector<GridInfo> grid_;
};

}  // namespace robotoc

#endif  // ROBOTOC_TIME_DISCRETIZATION_HPP_
