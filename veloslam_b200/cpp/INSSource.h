// INSSource.h -- online INS source of the drop-in facade with the reference's interface
// (/root/reference/INSSource.h:17-40, INSSource.cxx:217-326).
//
// One thread receives the NovAtel datagrams (UDP port 6777); an INSPVA message becomes a
// PoseTransform -- ENU position about the configured ECEF origin (CoordiTran llh2enu), Euler
// angles, velocity, TimeSolver timestamp -- that is appended to the output .insmeta file and
// added to the TransformManager (whose snapshot the GPU path uploads per batch).  Other
// message ids are ignored, as in the reference.  A whole INS log is converted on the GPU by
// vs_poses_from_ins instead.
#ifndef VELOSLAM_B200_INSSOURCE_H
#define VELOSLAM_B200_INSSOURCE_H

#include <cstdint>
#include <memory>
#include <string>

#include "TimeSolver.h"
#include "TransformManager.h"

enum PackageType { INSPVA = 508, RAWINS = 325, BESTGPSPOS = 423 };  // reference type_defs.h:34-36

class INSSource {
 public:
  INSSource(int port = 6777);
  virtual ~INSSource();

  void start();
  void stop();
  void setOutputFile(const std::string& filename);
  void setTransformManager(std::shared_ptr<TransformManager> mgr);
  void setTimeSolver(std::shared_ptr<TimeSolver> solver);
  void setOrigin(double org[3]);

  // ---- not in the reference ----------------------------------------------------------------
  uint64_t posesReceived() const;
  bool isRunning() const;
  // INSSource.cxx:300-326 for one record (what the receive thread calls)
  std::shared_ptr<PoseTransform> calcTransform(InsPVA const* data);

 private:
  INSSource(const INSSource&);
  void operator=(const INSSource&);
  class vsInternal;
  vsInternal* internal_;
};

#endif
