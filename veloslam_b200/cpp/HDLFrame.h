// HDLFrame.h -- result container of the drop-in facade (reference HDLFrame.h:13-83): one point
// list and one meta list per laser, the raw packets of the rotation, the car pose of the
// frame's first packet, and the file position / skip count that locate it in a pcap file.
#ifndef VELOSLAM_B200_HDLFRAME_H
#define VELOSLAM_B200_HDLFRAME_H

#include <atomic>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "type_defs.h"

struct HDLFrame {
  HDLFrame() : isInMemory(false), isOnHardDrive(false), count(0), fileStartPos(0), skips(0) {
    carpose = std::shared_ptr<PoseTransform>(new PoseTransform);
  }
  ptime timestamp;
  std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> points;
  std::vector<std::shared_ptr<PointMetaVector> > pointsMeta;
  std::vector<std::pair<ptime, std::string> > packets;
  std::shared_ptr<PoseTransform> carpose;
  bool isInMemory;
  bool isOnHardDrive;
  std::atomic<unsigned char> count;   // intrusive reference count of end users (reference HDLFrame.cxx:211-219)
  ptime filenameTime;
  int64_t fileStartPos;  // byte offset of the frame's first packet record (fpos_t in the reference)
  uint8_t skips;

  size_t numberOfPoints() const {
    size_t n = 0;
    for (const auto& c : points)
      if (c) n += c->points.size();
    return n;
  }
  // release points, meta and packets, keep the file locator (reference HDLFrame.cxx:142-158)
  void clear() {
    std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr>().swap(points);
    std::vector<std::shared_ptr<PointMetaVector> >().swap(pointsMeta);
    std::vector<std::pair<ptime, std::string> >().swap(packets);
    isInMemory = false;
  }
};

inline void intrusive_ptr_add_ref(HDLFrame* p) { ++p->count; }
inline void intrusive_ptr_release(HDLFrame* p) {
  unsigned char c = p->count.load();
  while (c != 0 && !p->count.compare_exchange_weak(c, (unsigned char)(c - 1))) {
  }
}

#endif
