// TransformManager.h -- pose timeline + interpolation with the reference's interface
// (/root/reference/TransformManager.h:80-123).  The timeline is kept as a sorted TimeLine and
// handed to the GPU as an immutable snapshot per batch (vs_set_poses).
#ifndef VELOSLAM_B200_TRANSFORMMANAGER_H
#define VELOSLAM_B200_TRANSFORMMANAGER_H

#include <cstdint>
#include <memory>
#include <ostream>
#include <mutex>
#include <string>
#include <vector>

#include "TimeLine.h"
#include "type_defs.h"

class TransformManager {
 public:
  TransformManager();
  virtual ~TransformManager();

  int getNumberOfTransforms();
  void clearTransforms();
  // best practice is to add poses in time order; out-of-order inserts and duplicate
  // timestamps (overwrite) are handled like the reference's TimeLine::addData
  void addTransform(std::shared_ptr<PoseTransform> trans);

  bool loadFromMetaFile(std::string filename, bool clearOldData = false);
  // 8 columns per line: x y yaw roll pitch v sec usec (yaw negated, radians -> degrees;
  // reference TransformManager.cxx:95-125)
  bool loadFromTxtFile(std::string filename, bool clearOldData = false);
  bool writeToMetaFile(const std::string& filename);
  // one .insmeta record (also what INSSource appends to its output file while receiving)
  static void writePoseRecord(std::ostream& os, const PoseTransform& p);

  // Linear interpolation of T, R (Euler degrees) and V between the bracketing samples; with
  // one sample: velocity extrapolation that leaves the pose flagged invalid; empty: false
  // (reference TransformManager.cxx:149-177)
  bool interpolateTransform(ptime& t, PoseTransform* xform);

  void setOriginLLH(const double LLH[3]);

  // --- B200 path: immutable snapshot of the timeline ---------------------------------------
  // monotonically increasing; bumped by every mutation
  uint64_t version();
  // sorted times (microseconds) and n x 9 doubles (T, R, V)
  void snapshot(std::vector<int64_t>* t_us, std::vector<double>* trv);
  // The part of the timeline interpolateTransform() can touch for any time in [tmin, tmax]: the
  // brackets clamp(lower_bound(t), 1, N-1) - 1 .. clamp(lower_bound(t), 1, N-1) of both ends and
  // everything between them (so extrapolation past either end of the timeline still sees its
  // two end samples).  What a batch of packets uploads instead of the whole history.
  void snapshotWindow(int64_t tmin_us, int64_t tmax_us, std::vector<int64_t>* t_us,
                      std::vector<double>* trv);

 protected:
  TimeLine<PoseTransform> transforms;

 private:
  TransformManager(const TransformManager&);
  void operator=(const TransformManager&);
  std::mutex mutex_;
  uint64_t version_;
  double originLLH[3];
};

#endif
