#pragma once
#include <cstdint>
#include <vector>
#include <boost/shared_ptr.hpp>
#include <pcl/point_types.h>
namespace pcl {
template <class PointT> class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
  PointCloud() : width(0), height(0), is_dense(true) {}
  std::vector<PointT> points;
  uint32_t width, height;
  bool is_dense;
  size_t size() const { return points.size(); }
  void push_back(const PointT& p) { points.push_back(p); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (uint32_t)points.size();
    height = 1;
    return *this;
  }
};
}
