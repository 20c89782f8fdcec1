#pragma once
#include <string>
#include <pcl/point_cloud.h>
namespace pcl { namespace io {
template <class PointT> int savePCDFileASCII(const std::string&, const PointCloud<PointT>&) { return -1; }
} }
