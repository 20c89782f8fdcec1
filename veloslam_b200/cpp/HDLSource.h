// HDLSource.h -- online packet source of the drop-in facade with the reference's interface
// (/root/reference/HDLSource.h:37-101, HDLSource.cxx:197-528).
//
// A receiver thread reads the sensor's UDP datagrams (port 2368) into a ring of packet slots; a
// consumer thread stamps each packet with TimeSolver::calcTimestamp(gpsTimestamp) and hands it
// to HDLParser::processHDLPacket, which batches it into the pinned ring the GPU decodes from;
// every frame the parser closes goes to HDLManager::addFrame (HDLSource.cxx:209-225).
// POSIX sockets and std::thread instead of boost::asio; no packet is copied into a heap string.
#ifndef VELOSLAM_B200_HDLSOURCE_H
#define VELOSLAM_B200_HDLSOURCE_H

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "HDLParser.h"
#include "TimeSolver.h"
#include "TransformManager.h"

class HDLManager;

class HDLSource {
 public:
  HDLSource(int _port = 2368);
  virtual ~HDLSource();

  void start();
  void stop();

  const std::string& getCorrectionsFile();
  void setCorrectionsFile(const std::string& correctionsFile);
  void setLaserSelection(int LaserSelection[64]);
  void getLaserSelection(int LaserSelection[64]);
  void setCropReturns(int);
  void setCropInside(int);
  void setCropRegion(double[6]);
  void setCropRegion(double, double, double, double, double, double);
  void getVerticalCorrections(double LaserAngles[64]);
  unsigned int getDualReturnFilter() const;
  void setDualReturnFilter(unsigned int);
  int getNumberOfChannels();

  void setHDLManager(HDLManager* hp);
  void setTimeSolver(std::shared_ptr<TimeSolver> solver);
  void setTransformManager(std::shared_ptr<TransformManager> mgr);
  std::shared_ptr<HDLParser> getHDLParser();

  // ---- not in the reference ----------------------------------------------------------------
  // packets received / dropped because the ring was full / handed to the parser so far
  void getCounters(uint64_t* received, uint64_t* dropped, uint64_t* consumed) const;
  // per frame handed to the manager, microseconds: from the closing packet entering the parser,
  // and from that packet's arrival on the socket
  void getFrameLatencies(std::vector<double>* processUs, std::vector<double>* fromArrivalUs) const;
  // deliver (payload, length, time) here instead of the parser (tests, custom consumers)
  void setPacketCallback(std::function<void(const unsigned char*, unsigned int, ptime)> cb);
  bool isRunning() const;

 protected:
  int sensorPort;
  std::string CorrectionsFile;

 private:
  HDLSource(const HDLSource&);
  void operator=(const HDLSource&);
  class vsInternal;
  vsInternal* internal_;
};

#endif
