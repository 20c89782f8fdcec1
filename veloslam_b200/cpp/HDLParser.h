// HDLParser.h -- drop-in facade with the reference's HDLParser interface
// (/root/reference/HDLParser.h:84-147) over the B200 C ABI (include/veloslam_b200.h).
//
// Same method names, argument meaning and error behaviour as the reference; the differences
// forced by the missing third-party libraries are: boost::shared_ptr -> std::shared_ptr,
// boost::posix_time::ptime -> ptime (type_defs.h), fpos_t -> int64_t byte offset.
//
// processHDLPacket() appends the packet to a pinned host ring; the ring is decoded on the GPU
// (segmentation + pose + decode/deskew kernels) when it is full, or when getAllFrames() is
// called and the buffered packets contain an azimuth decrease -- the only event that can
// close a frame -- so callers that poll getAllFrames() after every packet (HDLSource.cxx:
// 209-225) see each frame right after the packet that closes it, as with the reference.
// There is no CPU decode path: without a CUDA device every decode call reports an error.
#ifndef VELOSLAM_B200_HDLPARSER_H
#define VELOSLAM_B200_HDLPARSER_H

#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "HDLFrame.h"
#include "TransformManager.h"
#include "type_defs.h"

class HDLManager;

class HDLParser {
 public:
  enum DualFlag {
    DUAL_DISTANCE_NEAR = 0x1,
    DUAL_DISTANCE_FAR = 0x2,
    DUAL_INTENSITY_HIGH = 0x4,
    DUAL_INTENSITY_LOW = 0x8,
    DUAL_DOUBLED = 0xf,
    DUAL_DISTANCE_MASK = 0x3,
    DUAL_INTENSITY_MASK = 0xc,
  };

  HDLParser();
  ~HDLParser();

  const std::string& getDirName();
  void setDirName(const std::string& filename);

  const std::string& getCorrectionsFile();
  // reads the db.xml calibration; an unreadable file prints "Invalid sensor configuration
  // file" and leaves the parser unchanged (reference HDLParser.cxx:458-475)
  void setCorrectionsFile(const std::string& correctionsFile);

  void setNumberOfTrailingFrames(int numberTrailing);
  void setLaserSelection(int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int,
                         int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int,
                         int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int,
                         int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int);
  void setLaserSelection(int LaserSelection[64]);
  void getLaserSelection(int LaserSelection[64]);
  void getVerticalCorrections(double LaserAngles[64]);

  unsigned int getDualReturnFilter() const;
  void setDualReturnFilter(unsigned int);
  void setPointsSkip(int);
  void setCropReturns(int);
  void setCropInside(int);
  void setCropRegion(double[6]);
  void setCropRegion(double, double, double, double, double, double);

  int getNumberOfChannels();

  // offline index: one entry per rotation with fileStartPos / skips / timestamp
  // (touchOnly: only make sure the file is named after its first packet's time)
  std::vector<std::shared_ptr<HDLFrame> > readFrameInformation(const std::string& name,
                                                              bool touchOnly = false);
  void setHDLManager(std::shared_ptr<HDLManager> p);

  // decode one rotation from `filename` starting at byte offset startPos, skipping `skip`
  // firing blocks of the first packet
  bool getFrame(std::shared_ptr<HDLFrame>& dest, const std::string& filename, int64_t& startPos,
                const int& skip);
  void processHDLPacket(unsigned char* data, unsigned int bytesReceived, ptime t);
  // The reference returns the deque by value (HDLParser.h:135) and its consumer asks for it
  // after every packet (HDLSource.cxx:220-222: getAllFrames().size(), getAllFrames().back()):
  // two heap allocations per packet for an empty deque.  A reference to the parser's own list
  // keeps those call sites and `std::deque<...> fr = p.getAllFrames()` compiling unchanged and
  // makes the per-packet poll free; the list changes with the next processHDLPacket /
  // clearAllFrames, so callers that keep frames copy them (as the reference's do).
  const std::deque<std::shared_ptr<HDLFrame> >& getAllFrames();
  void clearAllFrames();
  std::shared_ptr<HDLFrame> createHDLFrame();

  std::shared_ptr<TransformManager> getTransformMgr() const;
  void setTransformMgr(std::shared_ptr<TransformManager> mgr);
  int getApplyTransform();
  void setApplyTransform(int apply);

  // ---- B200 controls (not in the reference) -------------------------------------------------
  void setDevice(int cudaDevice);          // before the first packet; default 0
  // Optional, for online use: create the CUDA context, launch every kernel once and page-lock
  // `frames` frame arenas now, so that the first rotations of a live stream do not pay for it
  // (call after setCorrectionsFile; 140 000 points ~ one 10 Hz HDL-64E rotation).
  bool prepare(int frames = 0, size_t pointsPerFrame = 140000);
  void setBatchPackets(int maxPackets);    // capacity of the pinned ring; default 4096
  void setStorePackets(bool store);        // keep raw packets inside HDLFrame::packets; default on
  void setFetchMeta(bool fetch);           // fill HDLFrame::pointsMeta (12 of the 28 bytes per point
                                           // that cross PCIe); default on
  // Throughput mode for replay through the per-packet API (before the first packet; default off):
  // two packet rings and two GPU result slots, so that filling the next ring, the host -> device
  // copy + kernels of the current batch and the device -> host copy of the previous one overlap.
  // getAllFrames() then returns frames when their batch has come back -- up to two rings
  // (2 x setBatchPackets packets) after the packet that closed them -- instead of forcing the
  // GPU round trip at every rotation; flush() drains everything.
  void setPipelined(bool on);
  void flush();                            // decode whatever is buffered now, wait for all of it
  // Keep a whole packet file (vtkPacketFileWriter format, fixed 1264-byte records) resident in
  // HBM: readFrameInformation() of that file becomes one segmentation pass on the GPU and
  // getFrame() decodes rotations straight out of HBM instead of re-reading the file
  // (HDLManager::loadOffline / prepareFrame).  false: not such a file, nothing changes.
  bool loadRecording(const std::string& pcapfile);
  // One shard of a recording: records [firstRecord, firstRecord + nRecords) only (nRecords < 0:
  // to the end of the file).  readFrameInformation() then indexes that range and getFrame()
  // serves the frames that start inside it, both in whole-file positions
  // (HDLManager::setDevices: one parser and one range per GPU).
  bool loadRecordingRange(const std::string& pcapfile, int64_t firstRecord, int64_t nRecords);
  void recordingRange(int64_t* firstRecord, int64_t* nRecords) const;
  void unloadRecording();
  bool hasRecording(const std::string& pcapfile) const;
  const std::string& lastError() const;    // empty when the last GPU call succeeded
  // where the host time of the pipelined mode went (seconds since construction), as JSON
  std::string pipelineStats() const;
  void resetPipelineStats();

 protected:
  void unloadData();

  std::string correctionsFile;
  std::string dirName;

  class vsInternal;
  vsInternal* internal_;

 private:
  HDLParser(const HDLParser&);
  void operator=(const HDLParser&);
};

// Names of the reference's earlier revisions of the same classes.
typedef HDLParser HDLReader;

#endif
