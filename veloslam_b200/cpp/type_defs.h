// type_defs.h -- data model of the drop-in facade (mirrors the reference's type_defs.h).
//
// Same names, fields and operators as /root/reference/type_defs.h:86-176, with the
// third-party types replaced by dependency-free equivalents:
//   boost::posix_time::ptime  -> ptime  (int64 microseconds since the Unix epoch + not_a_date_time)
//   boost::shared_ptr         -> std::shared_ptr (vs_shared_ptr alias)
//   Eigen::Affine3d           -> Affine3d ([L | t], row-major)
//   pcl::PointXYZI/PointCloud -> layout-compatible PODs
#ifndef VELOSLAM_B200_TYPE_DEFS_H
#define VELOSLAM_B200_TYPE_DEFS_H

#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "FrameArena.h"

#define HDL_NUM_ROT_ANGLES 36001
#define HDL_LASER_PER_FIRING 32
#define HDL_MAX_NUM_LASERS 64
#define HDL_FIRING_PER_PKT 12
#define HDL_MAX_PTS_PER_LASER 2200
#define HDL_META_EXT_NAME ".hdlmeta"
#define INS_META_EXT_NAME ".insmeta"

#define TO_RADIUS(degree) ((degree) * M_PI / 180)
#define TO_DEGREE(radius) ((radius) * 180 / M_PI)

template <class T> using vs_shared_ptr = std::shared_ptr<T>;

// Microsecond-resolution time point; default-constructed == not_a_date_time.
struct time_duration {
  int64_t us;
  time_duration() : us(0) {}
  time_duration(int64_t h, int64_t m, int64_t s, int64_t frac = 0)
      : us(((h * 60 + m) * 60 + s) * 1000000ll + frac) {}
  static time_duration microseconds(int64_t v) { time_duration d; d.us = v; return d; }
  int64_t total_microseconds() const { return us; }
  int64_t total_milliseconds() const { return us / 1000; }
};
struct ptime {
  int64_t us;
  ptime() : us(std::numeric_limits<int64_t>::min()) {}
  explicit ptime(int64_t microseconds_since_epoch) : us(microseconds_since_epoch) {}
  bool is_special() const { return us == std::numeric_limits<int64_t>::min(); }
  int64_t microseconds() const { return us; }
  bool operator==(const ptime& o) const { return us == o.us; }
  bool operator!=(const ptime& o) const { return us != o.us; }
  bool operator<(const ptime& o) const { return us < o.us; }
  bool operator>(const ptime& o) const { return us > o.us; }
  bool operator<=(const ptime& o) const { return us <= o.us; }
  bool operator>=(const ptime& o) const { return us >= o.us; }
  time_duration operator-(const ptime& o) const { return time_duration::microseconds(us - o.us); }
  ptime operator+(const time_duration& d) const { return ptime(us + d.us); }
  ptime operator-(const time_duration& d) const { return ptime(us - d.us); }
};
std::string to_iso_string(const ptime& t);   // "YYYYMMDDTHHMMSS[.ffffff]"
bool from_iso_string(const std::string& s, ptime* out);

// [L | t], the part of Eigen::Affine3d the path uses
struct Affine3d {
  double L[3][3];
  double t[3];
  double operator()(int i, int j) const { return j < 3 ? L[i][j] : t[i]; }
};

struct PoseTransform {
  double T[3];  // ENU metres
  double R[3];  // roll, pitch, yaw(azimuth), degrees
  double V[3];  // ENU m/s
  ptime timestamp;
  uint16_t week_number;
  uint32_t milliseconds;
  uint32_t week_number_pos;
  double seconds_pos;  // -1 == not a valid pose (reference type_defs.cxx:56)

  PoseTransform();
  // component-wise on T, R, V; timestamp is not handled (reference type_defs.h:99-131)
  PoseTransform operator+(PoseTransform delta) const;
  PoseTransform operator*(double ratio) const;
  PoseTransform operator-(PoseTransform delta) const;
  // Ry(R0) . Rx(R1) . Rz(R2), translation T (reference type_defs.h:134-146)
  Affine3d getMatrix() const;
};

// p' = L . p + t, rows summed left to right (reference type_defs.h:160-166)
void transformPoint(double pt0[3], const Affine3d& transform);

struct PointMeta {
  unsigned short azimuth;
  float distance;
  unsigned char intensityFlag;
  unsigned char distanceFlag;
  unsigned char flags;
};
static_assert(sizeof(PointMeta) == 12, "PointMeta is what the GPU writes (vs_layout_frames)");
// HDLFrame::pointsMeta[row]: a std::vector whose storage may be a slice of the frame's arena
typedef vs::AdoptableVector<PointMeta> PointMetaVector;

namespace pcl {
struct PointXYZI {
  float x, y, z;
  float intensity;
};
static_assert(sizeof(PointXYZI) == 16, "PointXYZI is what the GPU writes (vs_layout_frames)");
template <class PointT> struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT> > Ptr;
  // PCL's own type here is std::vector<PointT, Eigen::aligned_allocator<PointT>>: also a vector
  // with a non-default allocator.  This one can adopt a slice of the frame's page-locked arena.
  vs::AdoptableVector<PointT> points;
  uint32_t width = 0, height = 0;
  size_t size() const { return points.size(); }
};
}  // namespace pcl

#endif
